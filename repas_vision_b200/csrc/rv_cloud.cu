// rv_cloud.cu -- K3 pose transform + merge, K4 hash voxel grid, PLY record packing.
//
// Reference semantics (paths relative to the reference checkout):
//   geometry.transform(T)        femto_bolt_code/scripts/final_view_with_cad.py:333,
//                                realsense_d415i/vis_tool/vis_tool_april_tag_pose_validaiton.py:239-245
//   pcd.voxel_down_sample(v)     femto_bolt_code/scripts/mpa_icp_export.py:44,174 (+16 more call sites)
//   o3d.io.write_point_cloud     femto_bolt_code/scripts/create_masked_ply.py:177
// Arithmetic: Open3D 0.19 PointCloud::Transform / VoxelDownSample / PLY writer as restated in
// SURVEY.md Appendix B.1 (the wheel is not part of the reference checkout).
#include "rv_common.cuh"

namespace {

// ------------------------------------------------------------------ transform + merge
struct XformArgs {
  const void *in;
  void *out;
  long long in_stride, out_stride, n, out_offset;
  double T[16];
  double *bounds;  // 6 doubles or null
  int has_color;
};

template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

__device__ __forceinline__ void block_bounds_commit(double lo[3], double hi[3], double *bounds) {
  __shared__ double s_lo[3][8], s_hi[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = warp_min(lo[a]);
    hi[a] = warp_max(hi[a]);
    if (lane == 0) {
      s_lo[a][warp] = lo[a];
      s_hi[a][warp] = hi[a];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    double l = s_lo[a][0], h = s_hi[a][0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      l = s_lo[a][w] < l ? s_lo[a][w] : l;
      h = s_hi[a][w] > h ? s_hi[a][w] : h;
    }
    rv_atomic_min_f64(bounds + a, l);
    rv_atomic_max_f64(bounds + 3 + a, h);
  }
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) k_transform(const XformArgs a) {
  const InT *in = reinterpret_cast<const InT *>(a.in);
  OutT *out = reinterpret_cast<OutT *>(a.out) + a.out_offset;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    const double x = (double)in[i], y = (double)in[a.in_stride + i], z = (double)in[2 * a.in_stride + i];
    const double *T = a.T;
    // Open3D Transform: q = T [p,1]; p' = q.head<3>() / q(3).  ((T0*x + T1*y) + T2*z) + T3, one rounding per op
    const double qx = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
    const double qy = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
    const double qz = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
    const double qw = ((T[12] * x + T[13] * y) + T[14] * z) + T[15];
    double px = qx, py = qy, pz = qz;
    if (qw != 1.0) {
      px = qx / qw;
      py = qy / qw;
      pz = qz / qw;
    }
    const OutT ox = (OutT)px, oy = (OutT)py, oz = (OutT)pz;
    out[i] = ox;
    out[a.out_stride + i] = oy;
    out[2 * a.out_stride + i] = oz;
    if (a.has_color) {
      out[3 * a.out_stride + i] = (OutT)in[3 * a.in_stride + i];
      out[4 * a.out_stride + i] = (OutT)in[4 * a.in_stride + i];
      out[5 * a.out_stride + i] = (OutT)in[5 * a.in_stride + i];
    }
    if (a.bounds) {  // bounds of the STORED merged cloud
      const double sx = (double)ox, sy = (double)oy, sz = (double)oz;
      lo[0] = sx < lo[0] ? sx : lo[0];
      lo[1] = sy < lo[1] ? sy : lo[1];
      lo[2] = sz < lo[2] ? sz : lo[2];
      hi[0] = sx > hi[0] ? sx : hi[0];
      hi[1] = sy > hi[1] ? sy : hi[1];
      hi[2] = sz > hi[2] ? sz : hi[2];
    }
  }
  if (a.bounds) block_bounds_commit(lo, hi, a.bounds);
}

__global__ void k_bounds_init(double *b) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  if (threadIdx.x < 3) b[threadIdx.x] = inf;
  else if (threadIdx.x < 6) b[threadIdx.x] = -inf;
}

template <typename T>
__global__ void __launch_bounds__(256) k_bounds(const T *__restrict__ in, long long stride_in, long long n, double *bounds) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double v = (double)in[a * stride_in + i];
      lo[a] = v < lo[a] ? v : lo[a];
      hi[a] = v > hi[a] ? v : hi[a];
    }
  }
  block_bounds_commit(lo, hi, bounds);
}

// --------------------------------------------------------------------- voxel grid
// Workspace: [header 256 B][table: capacity x 64-byte slots]
struct VoxHeader {
  double bounds[6];
  unsigned long long n_out;  // voxel counter (compaction)
  int error;                 // 1: a voxel index does not fit 21 bits / extent check failed
  int pad;
};
struct __align__(64) VoxSlot {
  unsigned long long key;  // 0 = empty, else packed(ix,iy,iz) + 1
  unsigned int count;
  unsigned int pad;
  double sum[6];
};
static_assert(sizeof(VoxSlot) == 64, "slot must be one 64-byte record");

struct VoxArgs {
  const void *in;
  long long in_stride, n;
  int has_color;
  double voxel, rvoxel;
  const double *bounds;  // device
  VoxHeader *hdr;
  VoxSlot *table;
  unsigned long long cap_mask;
};

__device__ __forceinline__ unsigned long long vox_hash(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

template <typename T>
__global__ void __launch_bounds__(256) k_voxel_insert(const VoxArgs a) {
  const T *in = reinterpret_cast<const T *>(a.in);
  const double half = a.voxel * 0.5;
  const double ox = a.bounds[0] - half, oy = a.bounds[1] - half, oz = a.bounds[2] - half;
  {
    // Open3D: voxel_size * INT_MAX < max extent -> error; we additionally need 21-bit indices
    double ext = a.bounds[3] - a.bounds[0];
    const double ey = a.bounds[4] - a.bounds[1], ez = a.bounds[5] - a.bounds[2];
    ext = ey > ext ? ey : ext;
    ext = ez > ext ? ez : ext;
    if (!(ext / a.voxel + 1.0 < 2097152.0)) {
      if (blockIdx.x == 0 && threadIdx.x == 0) a.hdr->error = 1;
      return;
    }
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    const double x = (double)in[i], y = (double)in[a.in_stride + i], z = (double)in[2 * a.in_stride + i];
    // key = floor((p - origin) / voxel), IEEE division via the exact-reciprocal helper
    const long long ix = (long long)floor(rv_div(x - ox, a.voxel, a.rvoxel));
    const long long iy = (long long)floor(rv_div(y - oy, a.voxel, a.rvoxel));
    const long long iz = (long long)floor(rv_div(z - oz, a.voxel, a.rvoxel));
    const unsigned long long key =
        (((unsigned long long)ix & 0x1fffff) << 42 | ((unsigned long long)iy & 0x1fffff) << 21 | ((unsigned long long)iz & 0x1fffff)) + 1ull;
    unsigned long long h = vox_hash(key) & a.cap_mask;
    for (;;) {
      VoxSlot *s = a.table + h;
      unsigned long long cur = s->key;
      if (cur == 0) cur = atomicCAS(&s->key, 0ull, key);
      if (cur == 0 || cur == key) {
        atomicAdd(&s->count, 1u);
        atomicAdd(&s->sum[0], x);
        atomicAdd(&s->sum[1], y);
        atomicAdd(&s->sum[2], z);
        if (a.has_color) {
          atomicAdd(&s->sum[3], (double)in[3 * a.in_stride + i]);
          atomicAdd(&s->sum[4], (double)in[4 * a.in_stride + i]);
          atomicAdd(&s->sum[5], (double)in[5 * a.in_stride + i]);
        }
        break;
      }
      h = (h + 1) & a.cap_mask;
    }
  }
}

struct VoxOutArgs {
  VoxHeader *hdr;
  const VoxSlot *table;
  unsigned long long capacity;
  void *out;
  long long out_stride, out_capacity;
  int32_t *keys;
  int32_t *counts;
  long long *m;
  int has_color;
};

template <typename OutT>
__global__ void __launch_bounds__(256) k_voxel_emit(const VoxOutArgs a) {
  OutT *out = reinterpret_cast<OutT *>(a.out);
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // whole warps iterate together so the ballot below is always full
  const long long cap_round = (long long)((a.capacity + 31) & ~31ull);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cap_round; i += stride) {
    const bool live = (unsigned long long)i < a.capacity && a.table[i].key != 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, live);
    if (!bal) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&a.hdr->n_out, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!live) continue;
    const long long o = (long long)(base + __popc(bal & rv_lanemask_lt()));
    if (o >= a.out_capacity) continue;
    const VoxSlot s = a.table[i];
    const double c = (double)s.count;
    out[o] = (OutT)(s.sum[0] / c);
    out[a.out_stride + o] = (OutT)(s.sum[1] / c);
    out[2 * a.out_stride + o] = (OutT)(s.sum[2] / c);
    if (a.has_color) {
      out[3 * a.out_stride + o] = (OutT)(s.sum[3] / c);
      out[4 * a.out_stride + o] = (OutT)(s.sum[4] / c);
      out[5 * a.out_stride + o] = (OutT)(s.sum[5] / c);
    }
    if (a.keys) {
      const unsigned long long k = s.key - 1ull;
      a.keys[o] = (int32_t)((k >> 42) & 0x1fffff);
      a.keys[a.out_capacity + o] = (int32_t)((k >> 21) & 0x1fffff);
      a.keys[2 * a.out_capacity + o] = (int32_t)(k & 0x1fffff);
    }
    if (a.counts) a.counts[o] = (int32_t)s.count;
  }
}

__global__ void k_voxel_finish(const VoxHeader *hdr, long long *m) {
  *m = hdr->error ? -1ll : (long long)hdr->n_out;
}

unsigned long long vox_capacity(long long n) {
  unsigned long long c = 1024;
  while (c < (unsigned long long)n * 2ull) c <<= 1;
  return c;
}

// --------------------------------------------------------------------- PLY records
template <typename InT, typename CoordT>
__global__ void __launch_bounds__(256) k_pack_ply(const InT *__restrict__ in, long long stride_in, long long n,
                                                  int has_color, int color_255, uint8_t *__restrict__ rec) {
  constexpr int kRec = 3 * (int)sizeof(CoordT) + 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint8_t buf[kRec];
    CoordT xyz[3] = {(CoordT)in[i], (CoordT)in[stride_in + i], (CoordT)in[2 * stride_in + i]};
    memcpy(buf, xyz, sizeof(xyz));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double v = has_color ? (double)in[(3 + c) * stride_in + i] : 0.0;
      if (!color_255) {
        // Open3D: (uint8_t) round(min(1, max(0, c)) * 255)
        v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
        v = round(v * 255.0);
      } else {
        v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
        v = round(v);
      }
      buf[sizeof(xyz) + c] = (uint8_t)v;
    }
    uint8_t *dst = rec + i * kRec;
#pragma unroll
    for (int k = 0; k < kRec; ++k) dst[k] = buf[k];
  }
}

int grid_for(const rv_ctx *ctx, long long n, int per_sm = 8) {
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)ctx->sm_count * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" {

int rv_bounds_init(rv_ctx *ctx, double *d_bounds, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (!d_bounds) RV_FAIL(ctx, RV_EINVAL, "rv_bounds_init: null pointer");
  k_bounds_init<<<1, 32, 0, (cudaStream_t)stream>>>(d_bounds);
  RV_LAUNCHED(ctx);
  return RV_OK;
}

int rv_transform_merge(rv_ctx *ctx, int n_views, const void *const *d_in, const int64_t *in_plane_stride,
                       const int64_t *n, const double *T, int in_dtype, int has_color, void *d_out,
                       int64_t out_plane_stride, int out_dtype, double *d_bounds, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (n_views < 0 || (n_views > 0 && (!d_in || !in_plane_stride || !n || !T)))
    RV_FAIL(ctx, RV_EINVAL, "rv_transform_merge: null argument");
  if ((in_dtype != RV_F32 && in_dtype != RV_F64) || (out_dtype != RV_F32 && out_dtype != RV_F64))
    RV_FAIL(ctx, RV_EINVAL, "rv_transform_merge: bad dtype");
  long long total = 0;
  for (int v = 0; v < n_views; ++v) {
    if (n[v] < 0 || in_plane_stride[v] < n[v]) RV_FAIL(ctx, RV_EINVAL, "rv_transform_merge: bad n / stride for view %d", v);
    total += n[v];
  }
  if (total > out_plane_stride) RV_FAIL(ctx, RV_ECAPACITY, "rv_transform_merge: out_plane_stride %lld < %lld points",
                                        (long long)out_plane_stride, total);
  if (total > 0 && !d_out) RV_FAIL(ctx, RV_EINVAL, "rv_transform_merge: null output");
  cudaStream_t st = (cudaStream_t)stream;
  long long off = 0;
  for (int v = 0; v < n_views; ++v) {
    if (n[v] == 0) continue;
    if (!d_in[v]) RV_FAIL(ctx, RV_EINVAL, "rv_transform_merge: view %d is null", v);
    XformArgs a;
    memset(&a, 0, sizeof(a));
    a.in = d_in[v];
    a.out = d_out;
    a.in_stride = in_plane_stride[v];
    a.out_stride = out_plane_stride;
    a.n = n[v];
    a.out_offset = off;
    memcpy(a.T, T + 16 * v, sizeof(a.T));
    a.bounds = d_bounds;
    a.has_color = has_color ? 1 : 0;
    const int grid = grid_for(ctx, n[v]);
    if (in_dtype == RV_F32 && out_dtype == RV_F32) k_transform<float, float><<<grid, 256, 0, st>>>(a);
    else if (in_dtype == RV_F32) k_transform<float, double><<<grid, 256, 0, st>>>(a);
    else if (out_dtype == RV_F32) k_transform<double, float><<<grid, 256, 0, st>>>(a);
    else k_transform<double, double><<<grid, 256, 0, st>>>(a);
    RV_LAUNCHED(ctx);
    off += n[v];
  }
  return RV_OK;
}

size_t rv_voxel_workspace_bytes(int64_t n) {
  if (n < 0) n = 0;
  return 256 + (size_t)vox_capacity(n) * sizeof(VoxSlot);
}

int rv_voxel_downsample(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int in_dtype, int has_color,
                        double voxel_size, const double *d_bounds, void *d_out, int64_t out_plane_stride, int out_dtype,
                        int64_t out_capacity, int32_t *d_keys, int32_t *d_counts_out, int64_t *d_m, void *d_ws,
                        size_t ws_bytes, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (!(voxel_size > 0.0)) RV_FAIL(ctx, RV_EINVAL, "rv_voxel_downsample: voxel_size <= 0");
  if (n < 0 || in_plane_stride < n || !d_m) RV_FAIL(ctx, RV_EINVAL, "rv_voxel_downsample: bad n / stride / m");
  if ((in_dtype != RV_F32 && in_dtype != RV_F64) || (out_dtype != RV_F32 && out_dtype != RV_F64))
    RV_FAIL(ctx, RV_EINVAL, "rv_voxel_downsample: bad dtype");
  if (out_capacity < 0 || out_plane_stride < out_capacity) RV_FAIL(ctx, RV_EINVAL, "rv_voxel_downsample: bad output capacity");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    RV_CUDA(ctx, cudaMemsetAsync(d_m, 0, sizeof(int64_t), st));
    return RV_OK;
  }
  if (!d_in || (out_capacity > 0 && !d_out)) RV_FAIL(ctx, RV_EINVAL, "rv_voxel_downsample: null cloud pointer");
  const size_t need = rv_voxel_workspace_bytes(n);
  if (!d_ws || ws_bytes < need) RV_FAIL(ctx, RV_EWORKSPACE, "rv_voxel_downsample: workspace %zu < %zu", ws_bytes, need);
  if (!rv_aligned(d_ws, 64)) RV_FAIL(ctx, RV_EALIGN, "rv_voxel_downsample: workspace must be 64-byte aligned");
  const unsigned long long cap = vox_capacity(n);
  VoxHeader *hdr = reinterpret_cast<VoxHeader *>(d_ws);
  VoxSlot *table = reinterpret_cast<VoxSlot *>(reinterpret_cast<char *>(d_ws) + 256);
  RV_CUDA(ctx, cudaMemsetAsync(d_ws, 0, need, st));
  const double *bounds = d_bounds;
  if (!bounds) {
    k_bounds_init<<<1, 32, 0, st>>>(hdr->bounds);
    RV_LAUNCHED(ctx);
    const int g = grid_for(ctx, n);
    if (in_dtype == RV_F32) k_bounds<float><<<g, 256, 0, st>>>(reinterpret_cast<const float *>(d_in), in_plane_stride, n, hdr->bounds);
    else k_bounds<double><<<g, 256, 0, st>>>(reinterpret_cast<const double *>(d_in), in_plane_stride, n, hdr->bounds);
    RV_LAUNCHED(ctx);
    bounds = hdr->bounds;
  }
  VoxArgs a;
  memset(&a, 0, sizeof(a));
  a.in = d_in;
  a.in_stride = in_plane_stride;
  a.n = n;
  a.has_color = has_color ? 1 : 0;
  a.voxel = voxel_size;
  a.rvoxel = 1.0 / voxel_size;
  a.bounds = bounds;
  a.hdr = hdr;
  a.table = table;
  a.cap_mask = cap - 1;
  {
    const int g = grid_for(ctx, n);
    if (in_dtype == RV_F32) k_voxel_insert<float><<<g, 256, 0, st>>>(a);
    else k_voxel_insert<double><<<g, 256, 0, st>>>(a);
    RV_LAUNCHED(ctx);
  }
  VoxOutArgs o;
  memset(&o, 0, sizeof(o));
  o.hdr = hdr;
  o.table = table;
  o.capacity = cap;
  o.out = d_out;
  o.out_stride = out_plane_stride;
  o.out_capacity = out_capacity;
  o.keys = d_keys;
  o.counts = d_counts_out;
  o.has_color = has_color ? 1 : 0;
  {
    const int g = grid_for(ctx, (long long)cap);
    if (out_dtype == RV_F32) k_voxel_emit<float><<<g, 256, 0, st>>>(o);
    else k_voxel_emit<double><<<g, 256, 0, st>>>(o);
    RV_LAUNCHED(ctx);
  }
  k_voxel_finish<<<1, 1, 0, st>>>(hdr, reinterpret_cast<long long *>(d_m));
  RV_LAUNCHED(ctx);
  return RV_OK;
}

int rv_pack_ply_records(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int in_dtype, int has_color,
                        int color_scale, int coord_dtype, uint8_t *d_records, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (n < 0 || in_plane_stride < n) RV_FAIL(ctx, RV_EINVAL, "rv_pack_ply_records: bad n / stride");
  if ((in_dtype != RV_F32 && in_dtype != RV_F64) || (coord_dtype != RV_F32 && coord_dtype != RV_F64))
    RV_FAIL(ctx, RV_EINVAL, "rv_pack_ply_records: bad dtype");
  if (n == 0) return RV_OK;
  if (!d_in || !d_records) RV_FAIL(ctx, RV_EINVAL, "rv_pack_ply_records: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(ctx, n);
  const int c255 = color_scale == RV_COLOR_255;
  if (in_dtype == RV_F32 && coord_dtype == RV_F32)
    k_pack_ply<float, float><<<g, 256, 0, st>>>(reinterpret_cast<const float *>(d_in), in_plane_stride, n, has_color, c255, d_records);
  else if (in_dtype == RV_F32)
    k_pack_ply<float, double><<<g, 256, 0, st>>>(reinterpret_cast<const float *>(d_in), in_plane_stride, n, has_color, c255, d_records);
  else if (coord_dtype == RV_F32)
    k_pack_ply<double, float><<<g, 256, 0, st>>>(reinterpret_cast<const double *>(d_in), in_plane_stride, n, has_color, c255, d_records);
  else
    k_pack_ply<double, double><<<g, 256, 0, st>>>(reinterpret_cast<const double *>(d_in), in_plane_stride, n, has_color, c255, d_records);
  RV_LAUNCHED(ctx);
  return RV_OK;
}

}  // extern "C"
