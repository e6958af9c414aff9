// rv_register.cu -- K2: depth -> colour registration as a z-buffer scatter.
//
// Replaces the vendor calls AlignFilter(align_to_stream=COLOR_STREAM).process
// (femto_bolt_code/scripts/better_three_capture.py:169,188) and
// rs.align(rs.stream.color).process (realsense_d415i/capture_scripts/capture_aligned_all.py:75,197).
// Contract: SURVEY.md Appendix B.3 (librealsense align z16 -> other), float32 geometry with one
// rounding per operation (the library is built with --fmad=false), integer result.
//
// Each depth pixel maps its two half-pixel corners into the colour image and takes
// atomicMin over the covered rectangle on packed 64-bit keys (raw_depth << 32 | source index):
// the minimum depth wins and, on equal depth, the lowest source index -- exactly the result
// of the sequential reference loop.  Frames are processed in chunks whose key planes
// (8 B per colour pixel) stay resident in the 126 MB L2 between the scatter and the resolve.
#include "rv_common.cuh"

namespace {

struct CamF {
  float fx, fy, ppx, ppy;
  float k[5];
  int model;
  int width, height;
};

CamF to_camf(const RvCam &c) {
  CamF f;
  f.fx = (float)c.fx;
  f.fy = (float)c.fy;
  f.ppx = (float)c.cx;
  f.ppy = (float)c.cy;
  for (int i = 0; i < 5; ++i) f.k[i] = (float)c.dist[i];
  f.model = c.model;
  f.width = c.width;
  f.height = c.height;
  return f;
}

struct RegArgs {
  const uint16_t *depth;
  unsigned long long *keys;  // winners requested: (raw depth << 32) | source index
  unsigned int *keys32;      // depth only: raw depth, 0xffffffff = empty
  uint16_t *out;
  int32_t *winner;
  CamF dcam, ccam;
  float R[9], t[3];
  float depth_units;
  int frames;  // frames in this chunk
};

__device__ __forceinline__ void deproject_f32(float pt[3], const CamF &in, float px, float py, float depth) {
  float x = (px - in.ppx) / in.fx;
  float y = (py - in.ppy) / in.fy;
  if (in.model == RV_DIST_INVERSE_BROWN_CONRADY) {
    const float r2 = x * x + y * y;
    const float f = 1.0f + in.k[0] * r2 + in.k[1] * r2 * r2 + in.k[4] * r2 * r2 * r2;
    const float ux = x * f + 2.0f * in.k[2] * x * y + in.k[3] * (r2 + 2.0f * x * x);
    const float uy = y * f + 2.0f * in.k[3] * x * y + in.k[2] * (r2 + 2.0f * y * y);
    x = ux;
    y = uy;
  } else if (in.model == RV_DIST_BROWN_CONRADY) {
    const float xo = x, yo = y;
    for (int i = 0; i < 10; ++i) {
      const float r2 = x * x + y * y;
      const float icdist = 1.0f / (1.0f + ((in.k[4] * r2 + in.k[1]) * r2 + in.k[0]) * r2);
      const float dx = 2.0f * in.k[2] * x * y + in.k[3] * (r2 + 2.0f * x * x);
      const float dy = 2.0f * in.k[3] * x * y + in.k[2] * (r2 + 2.0f * y * y);
      x = (xo - dx) * icdist;
      y = (yo - dy) * icdist;
    }
  }
  pt[0] = depth * x;
  pt[1] = depth * y;
  pt[2] = depth;
}

__device__ __forceinline__ void project_f32(float pix[2], const CamF &in, const float p[3]) {
  float x = p[0] / p[2], y = p[1] / p[2];
  if (in.model == RV_DIST_MODIFIED_BROWN_CONRADY || in.model == RV_DIST_INVERSE_BROWN_CONRADY) {
    const float r2 = x * x + y * y;
    const float f = 1.0f + in.k[0] * r2 + in.k[1] * r2 * r2 + in.k[4] * r2 * r2 * r2;
    x *= f;
    y *= f;
    const float dx = x + 2.0f * in.k[2] * x * y + in.k[3] * (r2 + 2.0f * x * x);
    const float dy = y + 2.0f * in.k[3] * x * y + in.k[2] * (r2 + 2.0f * y * y);
    x = dx;
    y = dy;
  } else if (in.model == RV_DIST_BROWN_CONRADY) {
    const float r2 = x * x + y * y;
    const float f = 1.0f + in.k[0] * r2 + in.k[1] * r2 * r2 + in.k[4] * r2 * r2 * r2;
    const float xf = x * f, yf = y * f;
    const float dx = xf + 2.0f * in.k[2] * x * y + in.k[3] * (r2 + 2.0f * x * x);
    const float dy = yf + 2.0f * in.k[3] * x * y + in.k[2] * (r2 + 2.0f * y * y);
    x = dx;
    y = dy;
  }
  pix[0] = x * in.fx + in.ppx;
  pix[1] = y * in.fy + in.ppy;
}

// (int)(p + 0.5f); NaN or |.| >= 2^30 rejects the pixel (undefined in the C original)
__device__ __forceinline__ bool round_pix(float p, int &out) {
  const float q = p + 0.5f;
  if (!(fabsf(q) < 1073741824.0f)) return false;
  out = (int)q;  // cvt.rzi: truncation toward zero like the C cast
  return true;
}

__device__ __forceinline__ bool map_corner(const RegArgs &a, float px, float py, float d, int &ix, int &iy) {
  float pt[3], q[3], pix[2];
  deproject_f32(pt, a.dcam, px, py, d);
  q[0] = a.R[0] * pt[0] + a.R[3] * pt[1] + a.R[6] * pt[2] + a.t[0];
  q[1] = a.R[1] * pt[0] + a.R[4] * pt[1] + a.R[7] * pt[2] + a.t[1];
  q[2] = a.R[2] * pt[0] + a.R[5] * pt[1] + a.R[8] * pt[2] + a.t[2];
  project_f32(pix, a.ccam, q);
  return round_pix(pix[0], ix) && round_pix(pix[1], iy);
}

template <bool kWinner>
__global__ void __launch_bounds__(256) k_reg_scatter(const RegArgs a) {
  const int Wd = a.dcam.width, Hd = a.dcam.height;
  const int Wc = a.ccam.width, Hc = a.ccam.height;
  const long long Pd = (long long)Wd * Hd;
  const long long total = Pd * a.frames;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const uint32_t z = __ldg(a.depth + i);
    if (!z) continue;
    const int f = (int)(i / Pd);
    const int src = (int)(i - (long long)f * Pd);
    const int dy = src / Wd, dx = src - dy * Wd;
    const float d = (float)z * a.depth_units;
    int x0, y0, x1, y1;
    if (!map_corner(a, (float)dx - 0.5f, (float)dy - 0.5f, d, x0, y0)) continue;
    if (!map_corner(a, (float)dx + 0.5f, (float)dy + 0.5f, d, x1, y1)) continue;
    if (x0 < 0 || y0 < 0 || x1 >= Wc || y1 >= Hc) continue;
    if (kWinner) {
      const unsigned long long key = ((unsigned long long)z << 32) | (unsigned int)src;
      unsigned long long *kf = a.keys + (long long)f * Wc * Hc;
      for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) atomicMin(kf + (long long)y * Wc + x, key);
    } else {
      unsigned int *kf = a.keys32 + (long long)f * Wc * Hc;
      for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) atomicMin(kf + (long long)y * Wc + x, z);
    }
  }
}

// The resolve pass also puts the key plane back to "empty", so the next chunk of frames needs no memset.
__global__ void __launch_bounds__(256) k_reg_resolve32(unsigned int *__restrict__ keys, long long n, uint16_t *__restrict__ out) {
  // eight colour pixels per thread: two 16-byte key loads, one 16-byte depth store
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long octs = n >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < octs; i += stride) {
    uint4 *kp = reinterpret_cast<uint4 *>(keys + 8 * i);
    const uint4 k0 = kp[0], k1 = kp[1];
    const uint4 e = make_uint4(~0u, ~0u, ~0u, ~0u);
    kp[0] = e;
    kp[1] = e;
    uint4 o;
    o.x = (k0.x == ~0u ? 0u : k0.x) | ((k0.y == ~0u ? 0u : k0.y) << 16);
    o.y = (k0.z == ~0u ? 0u : k0.z) | ((k0.w == ~0u ? 0u : k0.w) << 16);
    o.z = (k1.x == ~0u ? 0u : k1.x) | ((k1.y == ~0u ? 0u : k1.y) << 16);
    o.w = (k1.z == ~0u ? 0u : k1.z) | ((k1.w == ~0u ? 0u : k1.w) << 16);
    *reinterpret_cast<uint4 *>(out + 8 * i) = o;
  }
  const long long tail0 = octs << 3;
  for (long long i = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned int k = keys[i];
    keys[i] = ~0u;
    out[i] = k == ~0u ? 0 : (uint16_t)k;
  }
}

__global__ void __launch_bounds__(256) k_reg_resolve(unsigned long long *__restrict__ keys, long long n,
                                                     uint16_t *__restrict__ out, int32_t *__restrict__ winner) {
  // two colour pixels per thread: one 16-byte key load, one 32-bit depth store
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long pairs = n >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
    const ulonglong2 k = *reinterpret_cast<const ulonglong2 *>(keys + 2 * i);
    *reinterpret_cast<ulonglong2 *>(keys + 2 * i) = make_ulonglong2(~0ull, ~0ull);
    const bool e0 = k.x == ~0ull, e1 = k.y == ~0ull;
    const uint32_t z0 = e0 ? 0u : (uint32_t)(k.x >> 32), z1 = e1 ? 0u : (uint32_t)(k.y >> 32);
    *reinterpret_cast<uint32_t *>(out + 2 * i) = z0 | (z1 << 16);
    if (winner) {
      int2 w;
      w.x = e0 ? -1 : (int)(uint32_t)k.x;
      w.y = e1 ? -1 : (int)(uint32_t)k.y;
      *reinterpret_cast<int2 *>(winner + 2 * i) = w;
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long k = keys[n - 1];
    keys[n - 1] = ~0ull;
    out[n - 1] = k == ~0ull ? 0 : (uint16_t)(k >> 32);
    if (winner) winner[n - 1] = k == ~0ull ? -1 : (int)(uint32_t)k;
  }
}

constexpr int kChunkFrames = 8;  // 8 x 7.4 MB of keys at 720p: L2-resident

}  // namespace

extern "C" {

size_t rv_register_workspace_bytes(int B, int Hc, int Wc) {
  if (B <= 0 || Hc <= 0 || Wc <= 0) return 16;
  const int chunk = B < kChunkFrames ? B : kChunkFrames;
  return (size_t)chunk * Hc * Wc * sizeof(unsigned long long);
}

int rv_register_depth_to_color(rv_ctx *ctx, const uint16_t *d_depth, int B, const RvCam *depth_cam,
                               const RvCam *color_cam, const float *R_colmajor, const float *t, float depth_units,
                               uint16_t *d_out, int32_t *d_winner, void *d_ws, size_t ws_bytes, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (!depth_cam || !color_cam || !R_colmajor || !t) RV_FAIL(ctx, RV_EINVAL, "rv_register: null camera / extrinsics");
  if (B < 0 || depth_cam->width <= 0 || depth_cam->height <= 0 || color_cam->width <= 0 || color_cam->height <= 0)
    RV_FAIL(ctx, RV_EINVAL, "rv_register: bad shape");
  if (B == 0) return RV_OK;
  if (!d_depth || !d_out) RV_FAIL(ctx, RV_EINVAL, "rv_register: null image pointer");
  const long long Pc = (long long)color_cam->width * color_cam->height;
  const long long Pd = (long long)depth_cam->width * depth_cam->height;
  if (Pd > 0x7fffffffll || Pc > 0x7fffffffll) RV_FAIL(ctx, RV_EINVAL, "rv_register: image too large");
  if (!d_ws || ws_bytes < (size_t)Pc * 8) RV_FAIL(ctx, RV_EWORKSPACE, "rv_register: workspace %zu < %zu", ws_bytes, (size_t)Pc * 8);
  if (!rv_aligned(d_ws, 16)) RV_FAIL(ctx, RV_EALIGN, "rv_register: workspace must be 16-byte aligned");
  if (!rv_aligned(d_out, 16) || (d_winner && !rv_aligned(d_winner, 8)))
    RV_FAIL(ctx, RV_EALIGN, "rv_register: out must be 16-byte and winner 8-byte aligned");
  if ((Pc & 1) && B > 1) RV_FAIL(ctx, RV_EINVAL, "rv_register: odd colour pixel count needs B == 1");
  cudaStream_t st = (cudaStream_t)stream;
  const bool want_winner = d_winner != nullptr;
  const size_t key_bytes = want_winner ? 8 : 4;  // depth-only registration packs nothing but the raw depth
  long long chunk = (long long)(ws_bytes / ((size_t)Pc * key_bytes));
  if (chunk > kChunkFrames * (want_winner ? 1 : 2)) chunk = kChunkFrames * (want_winner ? 1 : 2);
  if (chunk > B) chunk = B;
  if ((Pc & 7) && B > 1 && !want_winner) RV_FAIL(ctx, RV_EINVAL, "rv_register: colour pixel count must be a multiple of 8 for B > 1");

  RegArgs a;
  memset(&a, 0, sizeof(a));
  a.dcam = to_camf(*depth_cam);
  a.ccam = to_camf(*color_cam);
  for (int i = 0; i < 9; ++i) a.R[i] = R_colmajor[i];
  for (int i = 0; i < 3; ++i) a.t[i] = t[i];
  a.depth_units = depth_units;
  a.keys = reinterpret_cast<unsigned long long *>(d_ws);
  a.keys32 = reinterpret_cast<unsigned int *>(d_ws);
  const int max_blocks = ctx->sm_count * 8;
  // one memset for the whole call: every resolve pass leaves the plane empty for the next chunk
  RV_CUDA(ctx, cudaMemsetAsync(d_ws, 0xff, (size_t)chunk * Pc * key_bytes, st));
  for (long long f0 = 0; f0 < B; f0 += chunk) {
    const int nf = (int)((B - f0) < chunk ? (B - f0) : chunk);
    a.depth = d_depth + f0 * Pd;
    a.frames = nf;
    long long blocks = (Pd * nf + 255) / 256;
    if (blocks > max_blocks) blocks = max_blocks;
    if (want_winner) k_reg_scatter<true><<<(int)blocks, 256, 0, st>>>(a);
    else k_reg_scatter<false><<<(int)blocks, 256, 0, st>>>(a);
    RV_LAUNCHED(ctx);
    const long long n = Pc * nf;
    if (want_winner) {
      blocks = ((n >> 1) + 255) / 256;
      if (blocks > max_blocks) blocks = max_blocks;
      if (blocks < 1) blocks = 1;
      k_reg_resolve<<<(int)blocks, 256, 0, st>>>(a.keys, n, d_out + f0 * Pc, d_winner + f0 * Pc);
    } else {
      blocks = ((n >> 3) + 255) / 256;
      if (blocks > max_blocks) blocks = max_blocks;
      if (blocks < 1) blocks = 1;
      k_reg_resolve32<<<(int)blocks, 256, 0, st>>>(a.keys32, n, d_out + f0 * Pc);
    }
    RV_LAUNCHED(ctx);
  }
  return RV_OK;
}

}  // extern "C"
