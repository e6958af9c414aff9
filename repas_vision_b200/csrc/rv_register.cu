// rv_register.cu -- K2: depth -> colour registration as a z-buffer scatter.
//
// Replaces the vendor calls AlignFilter(align_to_stream=COLOR_STREAM).process
// (femto_bolt_code/scripts/better_three_capture.py:169,188) and
// rs.align(rs.stream.color).process (realsense_d415i/capture_scripts/capture_aligned_all.py:75,197).
// Contract: SURVEY.md Appendix B.3 (librealsense align z16 -> other), float32 geometry with one
// rounding per operation (the library is built with --fmad=false), integer result.
//
// Result = the sequential reference loop: over the rectangle each depth pixel covers, the minimum raw
// depth wins and, on equal depth, the lowest source index.  Two kernels, no global atomics on pixels:
//
//  k_reg_rects  one thread per depth pixel: both half-pixel corners through the float32 geometry
//               (the expensive part, done exactly once per pixel), the covered colour rectangle stored
//               as 8 bytes; every 32-pixel row segment ("block") also stores its bounding box and
//               appends itself to the list of each 64x32 colour tile the box touches.
//  k_reg_tile   one CTA per colour tile: a z-buffer of the tile in shared memory, filled with
//               shared-memory atomicMin from the blocks on the tile's list (second sweep for the
//               winners: lowest source index among the pixels that hold the minimum depth), then the
//               tile is written out with 16-byte stores.  A tile owns its pixels, so there is no key
//               plane in HBM, no memset and no resolve pass.
//
// Exact by construction: a tile whose list overflowed its fixed capacity scans every block's
// bounding box instead.  The rectangle records and lists of a chunk of frames stay in the L2.
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
  const uint16_t *depth;  // first frame of the chunk
  uint2 *rects;           // [frames][Pd]   x0 | y0 << 16, x1 | y1 << 16; .x = 0xffffffff: covers nothing
  uint2 *bbox;            // [frames][nblk] same packing, union over the block's pixels
  unsigned int *tcount;   // [frames][ntiles]
  unsigned int *tlist;    // [frames][ntiles][kListCap] blocks as depth row << 11 | block in the row
  uint16_t *out;
  int32_t *winner;
  CamF dcam, ccam;
  float R[9], t[3];
  float depth_units;
  int frames;  // frames in this chunk
  int segs;    // 32-pixel blocks per depth row
  int nblk;    // blocks per frame
  int tiles_x, tiles_y;
  float *tables;   // xn0[Wd], xn1[Wd], yn0[Hd], yn1[Hd] (k_reg_tables)
  int use_tables;  // depth camera without distortion: normalised corner coordinates come from per-column / per-row tables
  int vec_ok;      // rows of the colour image start 16-byte aligned: whole-vector stores
};

constexpr int kTW = 64, kTH = 32;   // colour tile owned by one CTA of k_reg_tile
constexpr int kTileCells = kTW * kTH;
constexpr int kListCap = 256;       // blocks per tile list; ~55 expected for 640x480 -> 1280x720
constexpr unsigned int kNone = 0xffffffffu;

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


// corner with normalised depth-camera coordinates (x, y) at depth d -> rounded colour pixel
// (kPlainColour: the colour camera has no distortion model, so the projection is the bare pinhole form of project_f32)
template <bool kPlainColour>
__device__ __forceinline__ bool map_norm(const RegArgs &a, float x, float y, float d, int &ix, int &iy) {
  float pt[3] = {d * x, d * y, d}, q[3], pix[2];
  q[0] = a.R[0] * pt[0] + a.R[3] * pt[1] + a.R[6] * pt[2] + a.t[0];
  q[1] = a.R[1] * pt[0] + a.R[4] * pt[1] + a.R[7] * pt[2] + a.t[1];
  q[2] = a.R[2] * pt[0] + a.R[5] * pt[1] + a.R[8] * pt[2] + a.t[2];
  if (kPlainColour) {
    pix[0] = (q[0] / q[2]) * a.ccam.fx + a.ccam.ppx;
    pix[1] = (q[1] / q[2]) * a.ccam.fy + a.ccam.ppy;
  } else {
    project_f32(pix, a.ccam, q);
  }
  return round_pix(pix[0], ix) && round_pix(pix[1], iy);
}

__device__ __forceinline__ bool map_corner(const RegArgs &a, float px, float py, float d, int &ix, int &iy) {
  float pt[3];
  deproject_f32(pt, a.dcam, px, py, 1.0f);  // normalised coordinates: 1 * x is exact, so d * x below is the same product
  return map_norm<false>(a, pt[0], pt[1], d, ix, iy);
}

// ---------------------------------------------------------------- kernel 0: normalised corner coordinates
// xn0[Wd], xn1[Wd], yn0[Hd], yn1[Hd]: the two operations rs2_deproject_pixel_to_point performs per corner, (p - pp) / f,
// done once per column / row instead of once per pixel (depth camera without distortion)
__global__ void __launch_bounds__(256) k_reg_tables(const RegArgs a) {
  const int Wd = a.dcam.width, Hd = a.dcam.height;
  float *const xn0 = a.tables, *const xn1 = xn0 + Wd, *const yn0 = xn1 + Wd, *const yn1 = yn0 + Hd;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Wd; i += gridDim.x * blockDim.x) {
    xn0[i] = (((float)i - 0.5f) - a.dcam.ppx) / a.dcam.fx;
    xn1[i] = (((float)i + 0.5f) - a.dcam.ppx) / a.dcam.fx;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Hd; i += gridDim.x * blockDim.x) {
    yn0[i] = (((float)i - 0.5f) - a.dcam.ppy) / a.dcam.fy;
    yn1[i] = (((float)i + 0.5f) - a.dcam.ppy) / a.dcam.fy;
  }
}

// ---------------------------------------------------------------- kernel 1: rectangles, block boxes, tile lists
// grid (depth rows, frames); the warps of a CTA walk the 32-pixel blocks of one row.
// kPlain: neither camera has a distortion model (tables for the depth side, bare pinhole projection): straight-line code
template <bool kPlain>
__global__ void __launch_bounds__(128) k_reg_rects(const RegArgs a) {
  const int Wd = a.dcam.width, Hd = a.dcam.height;
  const int Wc = a.ccam.width, Hc = a.ccam.height;
  const int dy = blockIdx.x, f = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const uint32_t fbase = (uint32_t)f * (uint32_t)(Wd * Hd);  // chunk-relative: frames * Pd < 2^32 is checked by the host
  const uint16_t *const drow = a.depth + fbase + (uint32_t)dy * Wd;
  uint2 *const rrow = a.rects + fbase + (uint32_t)dy * Wd;
  const bool tab = kPlain || a.use_tables;
  const float y0n = tab ? __ldg(a.tables + 2 * Wd + dy) : 0.f;
  const float y1n = tab ? __ldg(a.tables + 2 * Wd + Hd + dy) : 0.f;
  for (int sgm = threadIdx.x >> 5; sgm < a.segs; sgm += blockDim.x >> 5) {
    const int dx = sgm * 32 + lane;
    uint2 r = make_uint2(kNone, kNone);
    if (dx < Wd) {
      const uint32_t z = __ldg(drow + dx);
      if (z) {
        const float d = (float)z * a.depth_units;
        int x0, y0, x1, y1;
        bool ok;
        if (kPlain) {
          const bool ok0 = map_norm<true>(a, __ldg(a.tables + dx), y0n, d, x0, y0);
          const bool ok1 = map_norm<true>(a, __ldg(a.tables + Wd + dx), y1n, d, x1, y1);
          ok = ok0 && ok1;
        } else if (a.use_tables) {
          ok = map_norm<false>(a, __ldg(a.tables + dx), y0n, d, x0, y0) &&
               map_norm<false>(a, __ldg(a.tables + Wd + dx), y1n, d, x1, y1);
        } else {
          ok = map_corner(a, (float)dx - 0.5f, (float)dy - 0.5f, d, x0, y0) &&
               map_corner(a, (float)dx + 0.5f, (float)dy + 0.5f, d, x1, y1);
        }
        // the reference rejects a pixel with any corner outside the image; an empty loop range covers nothing
        if (ok && x0 >= 0 && y0 >= 0 && x1 < Wc && y1 < Hc && x0 <= x1 && y0 <= y1)
          r = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)x1 | ((uint32_t)y1 << 16));
      }
      rrow[dx] = r;
    }
    // bounding box of the block, then one list entry per colour tile it touches
    const bool live = r.x != kNone;
    const uint32_t bx0 = __reduce_min_sync(0xffffffffu, live ? (r.x & 0xffffu) : 0xffffu);
    const bool any = bx0 != 0xffffu;  // warp-uniform
    const int blk = dy * a.segs + sgm;
    uint2 bb = make_uint2(kNone, kNone);
    if (any) {
      const uint32_t by0 = __reduce_min_sync(0xffffffffu, live ? (r.x >> 16) : 0xffffu);
      const uint32_t bx1 = __reduce_max_sync(0xffffffffu, live ? (r.y & 0xffffu) : 0u);
      const uint32_t by1 = __reduce_max_sync(0xffffffffu, live ? (r.y >> 16) : 0u);
      bb = make_uint2(bx0 | (by0 << 16), bx1 | (by1 << 16));
      const int tx0 = (int)bx0 / kTW, tx1 = (int)bx1 / kTW, ty0 = (int)by0 / kTH, ty1 = (int)by1 / kTH;
      // lanes as an 8 x 4 patch of tiles: one step covers the box for anything but extreme up-scaling
      const unsigned int entry = ((unsigned int)dy << 11) | (unsigned int)sgm;
      const uint32_t tile0 = (uint32_t)f * a.tiles_y * a.tiles_x;
      if (tx1 - tx0 < 8 && ty1 - ty0 < 4) {
        const int tx = tx0 + (lane & 7), ty = ty0 + (lane >> 3);
        if (tx <= tx1 && ty <= ty1) {
          const uint32_t tile = tile0 + ty * a.tiles_x + tx;
          const unsigned int pos = atomicAdd(a.tcount + tile, 1u);
          if (pos < (unsigned int)kListCap) a.tlist[tile * kListCap + pos] = entry;
        }
      } else {
#pragma unroll 1
        for (int oy = ty0; oy <= ty1; oy += 4)
#pragma unroll 1
          for (int ox = tx0; ox <= tx1; ox += 8) {
            const int tx = ox + (lane & 7), ty = oy + (lane >> 3);
            if (tx <= tx1 && ty <= ty1) {
              const uint32_t tile = tile0 + ty * a.tiles_x + tx;
              const unsigned int pos = atomicAdd(a.tcount + tile, 1u);
              if (pos < (unsigned int)kListCap) a.tlist[tile * kListCap + pos] = entry;
            }
          }
      }
    }
    if (lane == 0) a.bbox[(uint32_t)f * a.nblk + blk] = bb;
  }
}

// ---------------------------------------------------------------- kernel 2: one colour tile per CTA
template <bool kSecond>
__device__ __forceinline__ void splat(uint2 r, unsigned int z, unsigned int srci, int X0, int Y0, unsigned int *zs,
                                      unsigned int *ws) {
  if (r.x == kNone) return;
  const int x0 = max((int)(r.x & 0xffffu) - X0, 0), y0 = max((int)(r.x >> 16) - Y0, 0);
  const int x1 = min((int)(r.y & 0xffffu) - X0, kTW - 1), y1 = min((int)(r.y >> 16) - Y0, kTH - 1);
  if (x0 > x1) return;
  const int w = x1 - x0;  // columns - 1; 1 or 2 for the usual 1.5x up-scaling
  for (int y = y0; y <= y1; ++y) {
    unsigned int *const zr = zs + y * kTW + x0;
    unsigned int *const wr = ws + y * kTW + x0;
    if (!kSecond) {
      atomicMin(zr, z);
      if (w >= 1) atomicMin(zr + 1, z);
      if (w >= 2) atomicMin(zr + 2, z);
      for (int x = 3; x <= w; ++x) atomicMin(zr + x, z);
    } else {
      if (zr[0] == z) atomicMin(wr, srci);
      if (w >= 1 && zr[1] == z) atomicMin(wr + 1, srci);
      if (w >= 2 && zr[2] == z) atomicMin(wr + 2, srci);
      for (int x = 3; x <= w; ++x)
        if (zr[x] == z) atomicMin(wr + x, srci);
    }
  }
}

// grid (colour tiles, frames)
template <bool kWinner>
__global__ void __launch_bounds__(256) k_reg_tile(const RegArgs a) {
  __shared__ __align__(16) unsigned int zs[kTileCells];
  __shared__ __align__(16) unsigned int ws[kWinner ? kTileCells : 4];
  const int Wc = a.ccam.width, Hc = a.ccam.height, Wd = a.dcam.width;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f = blockIdx.y, tile = blockIdx.x;
  const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
  const int X0 = tx * kTW, Y0 = ty * kTH;
  const uint32_t job = (uint32_t)f * (a.tiles_x * a.tiles_y) + tile;
  const uint32_t fbase = (uint32_t)f * (uint32_t)(Wd * a.dcam.height);
  const uint16_t *const dfr = a.depth + fbase;
  const uint2 *const rfr = a.rects + fbase;
  {
    const uint4 e = make_uint4(kNone, kNone, kNone, kNone);
    for (int i = threadIdx.x; i < kTileCells / 4; i += 256) {
      reinterpret_cast<uint4 *>(zs)[i] = e;
      if (kWinner) reinterpret_cast<uint4 *>(ws)[i] = e;
    }
  }
  const unsigned int n = a.tcount[job];
  const unsigned int *list = a.tlist + job * kListCap;
  const bool listed = n <= (unsigned int)kListCap;
  __syncthreads();
#pragma unroll 1
  for (int pass = 0; pass < (kWinner ? 2 : 1); ++pass) {
    if (listed) {
      // four list entries per warp and step: all their loads are in flight before the first shared-memory atomic
      for (int i0 = warp; i0 < (int)n; i0 += 32) {
        uint2 r[4];
        unsigned int z[4], srci[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          r[k] = make_uint2(kNone, kNone);
          z[k] = 0;
          srci[k] = 0;
          const int i = i0 + 8 * k;
          if (i < (int)n) {
            const unsigned int e = __ldg(list + i);  // row << 11 | block in the row
            const int by = (int)(e >> 11);
            const int dx = (int)(e & 2047u) * 32 + lane;
            if (dx < Wd) {
              srci[k] = (unsigned int)(by * Wd + dx);
              r[k] = __ldg(rfr + srci[k]);
              z[k] = __ldg(dfr + srci[k]);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (pass == 0) splat<false>(r[k], z[k], srci[k], X0, Y0, zs, ws);
          else splat<true>(r[k], z[k], srci[k], X0, Y0, zs, ws);
        }
      }
    } else {  // overflowed list: every block of the frame whose box touches the tile
      for (int blk = warp; blk < a.nblk; blk += 8) {
        const uint2 bb = a.bbox[(uint32_t)f * a.nblk + blk];
        if (bb.x == kNone) continue;
        if ((int)(bb.x & 0xffffu) > X0 + kTW - 1 || (int)(bb.y & 0xffffu) < X0 || (int)(bb.x >> 16) > Y0 + kTH - 1 ||
            (int)(bb.y >> 16) < Y0)
          continue;
        const int by = blk / a.segs;
        const int dx = (blk - by * a.segs) * 32 + lane;
        if (dx >= Wd) continue;
        const unsigned int srci = (unsigned int)(by * Wd + dx);
        if (pass == 0) splat<false>(rfr[srci], dfr[srci], srci, X0, Y0, zs, ws);
        else splat<true>(rfr[srci], dfr[srci], srci, X0, Y0, zs, ws);
      }
    }
    __syncthreads();
  }
  // write the tile: eight colour pixels per thread
  const int row = threadIdx.x >> 3, col = (threadIdx.x & 7) * 8;
  const int Y = Y0 + row, X = X0 + col;
  if (Y < Hc && X < Wc) {
    const unsigned int *zp = zs + row * kTW + col;
    const long long o = ((long long)f * Hc + Y) * Wc + X;
    if (a.vec_ok && X + 7 < Wc) {
      const uint4 k0 = *reinterpret_cast<const uint4 *>(zp), k1 = *reinterpret_cast<const uint4 *>(zp + 4);
      uint4 v;
      v.x = (k0.x == kNone ? 0u : k0.x) | ((k0.y == kNone ? 0u : k0.y) << 16);
      v.y = (k0.z == kNone ? 0u : k0.z) | ((k0.w == kNone ? 0u : k0.w) << 16);
      v.z = (k1.x == kNone ? 0u : k1.x) | ((k1.y == kNone ? 0u : k1.y) << 16);
      v.w = (k1.z == kNone ? 0u : k1.z) | ((k1.w == kNone ? 0u : k1.w) << 16);
      *reinterpret_cast<uint4 *>(a.out + o) = v;
      if (kWinner) {
        const int4 w0 = *reinterpret_cast<const int4 *>(ws + row * kTW + col);
        const int4 w1 = *reinterpret_cast<const int4 *>(ws + row * kTW + col + 4);
        *reinterpret_cast<int4 *>(a.winner + o) = w0;  // 0xffffffff is already -1
        *reinterpret_cast<int4 *>(a.winner + o + 4) = w1;
      }
    } else {
      for (int j = 0; j < 8 && X + j < Wc; ++j) {
        a.out[o + j] = zp[j] == kNone ? 0 : (uint16_t)zp[j];
        if (kWinner) a.winner[o + j] = (int)ws[row * kTW + col + j];
      }
    }
  }
}

struct RegLayout {
  size_t rects, bbox, tcount, tlist, per_frame;
};
size_t reg_tables_bytes(int Wd, int Hd) { return (((size_t)(Wd + Hd) * 2 * sizeof(float)) + 255) & ~(size_t)255; }
RegLayout reg_layout(long long Pd, int nblk, int ntiles) {
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  RegLayout l;
  l.rects = up((size_t)Pd * 8);
  l.bbox = up((size_t)nblk * 8);
  l.tcount = up((size_t)ntiles * 4);
  l.tlist = up((size_t)ntiles * kListCap * 4);
  l.per_frame = l.rects + l.bbox + l.tcount + l.tlist;
  return l;
}

constexpr int kChunkFrames = 16;  // 16 x ~3 MB of rectangles + lists (640x480 -> 720p): L2-resident between the two kernels

}  // namespace

extern "C" {

size_t rv_register_workspace_bytes(int B, int Hd, int Wd, int Hc, int Wc) {
  if (B <= 0 || Hd <= 0 || Wd <= 0 || Hc <= 0 || Wc <= 0) return 256;
  const int chunk = B < kChunkFrames ? B : kChunkFrames;
  const int nblk = Hd * ((Wd + 31) / 32);
  const int ntiles = ((Wc + kTW - 1) / kTW) * ((Hc + kTH - 1) / kTH);
  return reg_tables_bytes(Wd, Hd) + reg_layout((long long)Hd * Wd, nblk, ntiles).per_frame * chunk;
}

int rv_register_depth_to_color(rv_ctx *ctx, const uint16_t *d_depth, int B, const RvCam *depth_cam,
                               const RvCam *color_cam, const float *R_colmajor, const float *t, float depth_units,
                               uint16_t *d_out, int32_t *d_winner, void *d_ws, size_t ws_bytes, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (!depth_cam || !color_cam || !R_colmajor || !t) RV_FAIL(ctx, RV_EINVAL, "rv_register: null camera / extrinsics");
  if (B < 0 || depth_cam->width <= 0 || depth_cam->height <= 0 || color_cam->width <= 0 || color_cam->height <= 0)
    RV_FAIL(ctx, RV_EINVAL, "rv_register: bad shape");
  if (color_cam->width > 65535 || color_cam->height > 65535) RV_FAIL(ctx, RV_EINVAL, "rv_register: colour image side > 65535");
  if (B == 0) return RV_OK;
  if (!d_depth || !d_out) RV_FAIL(ctx, RV_EINVAL, "rv_register: null image pointer");
  const int Wd = depth_cam->width, Hd = depth_cam->height, Wc = color_cam->width, Hc = color_cam->height;
  const long long Pc = (long long)Wc * Hc;
  const long long Pd = (long long)Wd * Hd;
  if (Pd > 0x7fffffffll || Pc > 0x7fffffffll) RV_FAIL(ctx, RV_EINVAL, "rv_register: image too large");
  const int segs = (Wd + 31) / 32, nblk = Hd * segs;
  const int tiles_x = (Wc + kTW - 1) / kTW, tiles_y = (Hc + kTH - 1) / kTH, ntiles = tiles_x * tiles_y;
  const RegLayout L = reg_layout(Pd, nblk, ntiles);
  const size_t tab = reg_tables_bytes(Wd, Hd);
  if (!d_ws || ws_bytes < tab + L.per_frame)
    RV_FAIL(ctx, RV_EWORKSPACE, "rv_register: workspace %zu < %zu", ws_bytes, tab + L.per_frame);
  if (!rv_aligned(d_ws, 16)) RV_FAIL(ctx, RV_EALIGN, "rv_register: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const bool want_winner = d_winner != nullptr;
  long long chunk = (long long)((ws_bytes - tab) / L.per_frame);
  if (chunk > kChunkFrames) chunk = kChunkFrames;
  if (chunk > B) chunk = B;
  if (chunk > 65535) chunk = 65535;                                 // gridDim.y
  while (chunk > 1 && chunk * Pd > 0xffffffffll) chunk >>= 1;       // 32-bit pixel offsets inside a chunk
  if (Wd > 65535 || Hd > (1 << 21) - 1) RV_FAIL(ctx, RV_EINVAL, "rv_register: depth image too large");

  RegArgs a;
  memset(&a, 0, sizeof(a));
  a.dcam = to_camf(*depth_cam);
  a.ccam = to_camf(*color_cam);
  for (int i = 0; i < 9; ++i) a.R[i] = R_colmajor[i];
  for (int i = 0; i < 3; ++i) a.t[i] = t[i];
  a.depth_units = depth_units;
  a.segs = segs;
  a.nblk = nblk;
  a.tiles_x = tiles_x;
  a.tiles_y = tiles_y;
  unsigned char *w = reinterpret_cast<unsigned char *>(d_ws);
  a.tables = reinterpret_cast<float *>(w);
  w += tab;
  a.rects = reinterpret_cast<uint2 *>(w);
  a.bbox = reinterpret_cast<uint2 *>(w + L.rects * chunk);
  a.tcount = reinterpret_cast<unsigned int *>(w + (L.rects + L.bbox) * chunk);
  a.tlist = reinterpret_cast<unsigned int *>(w + (L.rects + L.bbox + L.tcount) * chunk);
  a.use_tables = depth_cam->model == RV_DIST_NONE;
  a.vec_ok = (Wc % 8 == 0) && rv_aligned(d_out, 16) && (!want_winner || rv_aligned(d_winner, 16));
  if (a.use_tables) {
    k_reg_tables<<<(Wd + Hd + 255) / 256, 256, 0, st>>>(a);
    RV_LAUNCHED(ctx);
  }
  int rect_threads = 32 * (segs < 4 ? segs : 4);  // 4 warps per depth row; 5 blocks each at 640 columns
  for (long long f0 = 0; f0 < B; f0 += chunk) {
    const int nf = (int)((B - f0) < chunk ? (B - f0) : chunk);
    a.depth = d_depth + f0 * Pd;
    a.out = d_out + f0 * Pc;
    a.winner = want_winner ? d_winner + f0 * Pc : nullptr;
    a.frames = nf;
    RV_CUDA(ctx, cudaMemsetAsync(a.tcount, 0, (size_t)nf * ntiles * sizeof(unsigned int), st));
    if (a.use_tables && color_cam->model == RV_DIST_NONE) k_reg_rects<true><<<dim3((unsigned)Hd, (unsigned)nf), rect_threads, 0, st>>>(a);
    else k_reg_rects<false><<<dim3((unsigned)Hd, (unsigned)nf), rect_threads, 0, st>>>(a);
    RV_LAUNCHED(ctx);
    if (want_winner) k_reg_tile<true><<<dim3((unsigned)ntiles, (unsigned)nf), 256, 0, st>>>(a);
    else k_reg_tile<false><<<dim3((unsigned)ntiles, (unsigned)nf), 256, 0, st>>>(a);
    RV_LAUNCHED(ctx);
  }
  return RV_OK;
}

}  // extern "C"
