// rv_deproject.cu -- K1: depth units, ray table, fused deprojection + masks + stream
// compaction, and the stand-alone cloud filter.
//
// Reference semantics (paths relative to the reference checkout):
//   create_masked_pointcloud     femto_bolt_code/scripts/create_masked_ply.py:56-107
//   depth_to_meters              femto_bolt_code/scripts/better_three_capture.py:118-125
//   distance mask                realsense_d415i/capture_scripts/distance_masking_on_ply.py:12-19
//   Z clip                       femto_bolt_code/scripts/view_point_cloud.py:109-116
//   AABB crop                    femto_bolt_code/scripts/april_tag_bg_removal_pl.py:450-455
//   rs2_deproject_pixel_to_point SURVEY.md Appendix B.3 (ray table)
//
// Data layout: a frame is a flat array of P = H*W pixels.  A tile is 2048 consecutive
// pixels; warp w of a 256-thread CTA owns pixels [256w, 256w+256) of the tile and walks
// them 32 at a time, so every warp-wide load and every compacted store is one contiguous
// run.  Ordered compaction chains the tiles of a frame with a decoupled look-back
// (rv_common.cuh); tiles are handed out by an atomic ticket so a tile's predecessors
// are always already running.
#include "rv_common.cuh"
#include "rv_deproject_args.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kIters = 8;
constexpr int kTile = kThreads * kIters;  // 2048 pixels
static_assert(kTile == kGenericTilePx, "rv_deproject_args.cuh sizes the workspace from this tile");


template <typename T>
__device__ __forceinline__ T color_value(uint32_t k, int color_255);
template <>
__device__ __forceinline__ float color_value<float>(uint32_t k, int color_255) {
  const float kf = (float)k;
  return color_255 ? kf : rv_divf(kf, 255.0f, 0.003921568859368563f /* RN_f32(1/255) */);
}
template <>
__device__ __forceinline__ double color_value<double>(uint32_t k, int color_255) {
  const double kd = (double)k;
  return color_255 ? kd : rv_div(kd, 255.0, 0.00392156862745098 /* RN(1/255) */);
}

template <typename T>
__device__ __forceinline__ T quiet_nan();
template <>
__device__ __forceinline__ float quiet_nan<float>() {
  return __int_as_float(0x7fc00000);
}
template <>
__device__ __forceinline__ double quiet_nan<double>() {
  return __longlong_as_double(0x7ff8000000000000ll);
}

// MODE: RvDeprojectMode.  DK: RvDepthKind.
template <typename OutT, int DK, int MODE>
__global__ void __launch_bounds__(kThreads) k_deproject(const DeprojArgs a) {
  constexpr bool kPacked = MODE == RV_MODE_COMPACT_PACKED;
  constexpr bool kOrdered = MODE == RV_MODE_COMPACT_ORDERED || kPacked;
  constexpr bool kCompact = kOrdered || MODE == RV_MODE_COMPACT_UNORDERED;
  __shared__ uint32_t s_warp_tot[kWarps];
  __shared__ int s_tile;
  __shared__ unsigned long long s_base;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t lt = rv_lanemask_lt();
  OutT *const out = reinterpret_cast<OutT *>(a.out);

  for (int round = 0;; ++round) {
    int tile;
    if (kOrdered) {
      if (threadIdx.x == 0) s_tile = (int)atomicAdd(a.ticket, 1u);
      __syncthreads();
      tile = s_tile;
    } else {
      tile = blockIdx.x + round * gridDim.x;
    }
    if (tile >= a.total_tiles) break;
    const int b = tile / a.tiles_per_frame;
    const int t = tile - b * a.tiles_per_frame;
    const int p0 = t * kTile + warp * (32 * kIters) + lane;  // first pixel of this lane
    const long long fpix = (long long)b * a.P;               // frame offset in pixels

    // ---------------- loads: everything this lane needs, issued before any use
    uint32_t draw[kIters];
    uint32_t craw[kIters];
    uint32_t mraw[kIters];
#pragma unroll
    for (int j = 0; j < kIters; ++j) {
      const int p = p0 + j * 32;
      draw[j] = 0;
      craw[j] = 0;
      mraw[j] = 255;
      if (p < a.P) {
        const long long g = fpix + p;
        if (DK == RV_DEPTH_U16)
          draw[j] = __ldg(reinterpret_cast<const uint16_t *>(a.depth) + g);
        else
          draw[j] = __float_as_uint(__ldg(reinterpret_cast<const float *>(a.depth) + g));
        if (a.use_mask) mraw[j] = __ldg(a.mask + g);
        if (a.bgr && a.color_nv12) {
          // NV12 frame: H rows of luma, then H/2 rows of interleaved U,V (one pair per 2 x 2 pixels)
          const uint8_t *f = a.bgr + (long long)b * (a.P + (a.P >> 1));
          const int pv = p / a.W, pu = p - pv * a.W;
          const uint8_t *uv = f + a.P + (long long)(pv >> 1) * a.W + (pu & ~1);
          craw[j] = rv_nv12_pixel_bgr(__ldg(f + p), __ldg(uv), __ldg(uv + 1));
        } else if (a.bgr) {
          const uint8_t *c = a.bgr + 3 * g;
          craw[j] = (uint32_t)__ldg(c) | ((uint32_t)__ldg(c + 1) << 8) | ((uint32_t)__ldg(c + 2) << 16);
        }
      }
    }

    // ---------------- geometry + predicates
    int u = p0 % a.W;
    int v = p0 / a.W;
    OutT xs[kIters], ys[kIters], zs[kIters];
    uint32_t ballots[kIters];
    uint32_t warp_total = 0;
#pragma unroll
    for (int j = 0; j < kIters; ++j) {
      const int p = p0 + j * 32;
      float z32;
      double z64;
      bool ok;
      if (DK == RV_DEPTH_U16) {
        const float df = (float)draw[j];
        if (a.unit_rule == RV_UNIT_MUL_F32) {
          z32 = df * a.unit_scale_f;
          z64 = (double)z32;
        } else if (a.unit_rule == RV_UNIT_DIV_F32) {
          z32 = rv_divf(df, a.unit_scale_f, a.unit_rcp_f);
          z64 = (double)z32;
        } else {
          z64 = rv_div((double)draw[j], a.unit_scale, a.unit_rcp);
          z32 = (float)z64;
        }
        ok = draw[j] != 0;
      } else {
        z32 = __uint_as_float(draw[j]);
        z64 = (double)z32;
        ok = (z32 > 0.0f) && (z32 < __int_as_float(0x7f800000));  // finite and positive
      }
      ok = ok && (p < a.P);
      if (a.use_mask) ok = ok && (a.invert_mask ? (mraw[j] == 0) : (mraw[j] != 0));
      if (a.use_trunc) ok = ok && !(z32 >= a.trunc_f);

      double x64, y64;
      if (a.geom_f32) {  // the SDKs' float32 form: z * ((u - ppx) / fx), IEEE division, one rounding per operation
        x64 = (double)(z32 * rv_divf((float)u - a.cx_f, a.fx_f, a.rfx_f));
        y64 = (double)(z32 * rv_divf((float)v - a.cy_f, a.fy_f, a.rfy_f));
      } else if (a.rays) {
        double2 r = make_double2(0.0, 0.0);
        if (p < a.P) r = __ldg(a.rays + p);
        x64 = z64 * r.x;
        y64 = z64 * r.y;
      } else {
        x64 = rv_div(((double)u - a.cx) * z64, a.fx, a.rfx);
        y64 = rv_div(((double)v - a.cy) * z64, a.fy, a.rfy);
      }
      const OutT xo = (OutT)x64, yo = (OutT)y64, zo = (OutT)z64;
      if (a.use_zclip | a.use_radius | a.use_aabb) {
        // predicates on the STORED values (exact up-casts for float32 storage)
        const double X = (double)xo, Y = (double)yo, Z = (double)zo;
        if (a.use_zclip) ok = ok && (Z >= a.z_min) && (Z <= a.z_max);
        if (a.use_radius) ok = ok && (((X * X + Y * Y) + Z * Z) < a.r2_thresh);
        if (a.use_aabb)
          ok = ok && (X >= a.amin[0]) && (X <= a.amax[0]) && (Y >= a.amin[1]) && (Y <= a.amax[1]) &&
               (Z >= a.amin[2]) && (Z <= a.amax[2]);
      }
      xs[j] = xo;
      ys[j] = yo;
      zs[j] = zo;
      ballots[j] = __ballot_sync(0xffffffffu, ok);
      warp_total += __popc(ballots[j]);
      if (a.valid && p < a.P) a.valid[fpix + p] = ok ? 1 : 0;
      u += 32;
      while (u >= a.W) {
        u -= a.W;
        ++v;
      }
    }

    // ---------------- tile totals, tile base
    if (lane == 0) s_warp_tot[warp] = warp_total;
    __syncthreads();
    uint32_t warp_excl = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = s_warp_tot[w];
      warp_excl += (w < warp) ? c : 0u;
      tile_total += c;
    }
    unsigned long long base = 0;
    if (kOrdered) {
      if (warp == 0) {
        // per-frame chains restart at every frame; the packed chain runs through the whole batch
        const uint32_t excl = rv_lookback(a.status, tile, kPacked ? tile : t, tile_total);
        if (lane == 0) {
          s_base = excl;
          if (kPacked) {
            if (t == 0) a.counts[b] = excl;  // exclusive offset of frame b
            if (tile == a.total_tiles - 1) a.counts[a.B] = (unsigned long long)excl + tile_total;
          } else if (t == a.tiles_per_frame - 1) {
            a.counts[b] = (unsigned long long)excl + tile_total;
          }
        }
      }
      __syncthreads();
      base = s_base;
    } else if (MODE == RV_MODE_COMPACT_UNORDERED) {
      if (threadIdx.x == 0) s_base = atomicAdd(a.counts + b, (unsigned long long)tile_total);
      __syncthreads();
      base = s_base;
    } else {
      if (threadIdx.x == 0 && tile_total) atomicAdd(a.counts + b, (unsigned long long)tile_total);
    }

    // ---------------- stores
    const long long fout = kPacked ? 0ll : (long long)b * a.frame_stride;
    const unsigned long long cap = kPacked ? (unsigned long long)a.plane_stride : (unsigned long long)a.frame_stride;
    unsigned long long run = base + warp_excl;
#pragma unroll
    for (int j = 0; j < kIters; ++j) {
      const int p = p0 + j * 32;
      const bool ok = (ballots[j] >> lane) & 1u;
      if (kCompact) {
        const unsigned long long pos = run + __popc(ballots[j] & lt);
        run += __popc(ballots[j]);
        if (ok && pos < cap) {
          OutT *o = out + fout + pos;
          o[0] = xs[j];
          o[a.plane_stride] = ys[j];
          o[2 * a.plane_stride] = zs[j];
          if (a.bgr && a.color_packed) {
            reinterpret_cast<uint32_t *>(o + 3 * a.plane_stride)[0] = __byte_perm(craw[j], 0u, 0x4012);  // bytes r,g,b,0
          } else if (a.bgr) {
            o[3 * a.plane_stride] = color_value<OutT>((craw[j] >> 16) & 255u, a.color_255);
            o[4 * a.plane_stride] = color_value<OutT>((craw[j] >> 8) & 255u, a.color_255);
            o[5 * a.plane_stride] = color_value<OutT>(craw[j] & 255u, a.color_255);
          }
          if (a.src_index) a.src_index[fout + pos] = p;
        }
      } else if (p < a.P && p < a.frame_stride) {
        OutT *o = out + fout + p;
        const OutT bad = (MODE == RV_MODE_DENSE_NAN) ? quiet_nan<OutT>() : (OutT)0;
        o[0] = ok ? xs[j] : bad;
        o[a.plane_stride] = ok ? ys[j] : bad;
        o[2 * a.plane_stride] = ok ? zs[j] : bad;
        if (a.bgr && a.color_packed) {
          reinterpret_cast<uint32_t *>(o + 3 * a.plane_stride)[0] = ok ? __byte_perm(craw[j], 0u, 0x4012) : 0u;
        } else if (a.bgr) {
          o[3 * a.plane_stride] = ok ? color_value<OutT>((craw[j] >> 16) & 255u, a.color_255) : (OutT)0;
          o[4 * a.plane_stride] = ok ? color_value<OutT>((craw[j] >> 8) & 255u, a.color_255) : (OutT)0;
          o[5 * a.plane_stride] = ok ? color_value<OutT>(craw[j] & 255u, a.color_255) : (OutT)0;
        }
        if (a.src_index) a.src_index[fout + p] = ok ? p : -1;
      }
    }
    if (!kOrdered) __syncthreads();  // s_warp_tot / s_base are reused next round
  }
}

// ------------------------------------------------------------------ depth units
__global__ void __launch_bounds__(256) k_depth_to_meters(const uint16_t *__restrict__ d, long long n, int rule,
                                                         double scale, double rcp, float scale_f, float rcp_f,
                                                         float *__restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t r = d[i];
    float z;
    if (rule == RV_UNIT_MUL_F32)
      z = (float)r * scale_f;
    else if (rule == RV_UNIT_DIV_F32)
      z = rv_divf((float)r, scale_f, rcp_f);
    else
      z = (float)rv_div((double)r, scale, rcp);
    out[i] = z;
  }
}

// -------------------------------------------------------------------- ray table
// One-off per camera: float64 rs2_deproject_pixel_to_point normalised ray per pixel.
__global__ void __launch_bounds__(256) k_ray_table(RvCam cam, double2 *__restrict__ table) {
  const int n = cam.width * cam.height;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int u = i % cam.width, v = i / cam.width;
  double x = ((double)u - cam.cx) / cam.fx;
  double y = ((double)v - cam.cy) / cam.fy;
  const double k1 = cam.dist[0], k2 = cam.dist[1], p1 = cam.dist[2], p2 = cam.dist[3], k3 = cam.dist[4];
  if (cam.model == RV_DIST_INVERSE_BROWN_CONRADY) {
    const double r2 = x * x + y * y;
    const double f = 1.0 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2;
    const double ux = x * f + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
    const double uy = y * f + 2.0 * p2 * x * y + p1 * (r2 + 2.0 * y * y);
    x = ux;
    y = uy;
  } else if (cam.model == RV_DIST_BROWN_CONRADY) {
    const double xo = x, yo = y;
    for (int it = 0; it < 10; ++it) {
      const double r2 = x * x + y * y;
      const double icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2);
      const double dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
      const double dy = 2.0 * p2 * x * y + p1 * (r2 + 2.0 * y * y);
      x = (xo - dx) * icdist;
      y = (yo - dy) * icdist;
    }
  }
  table[i] = make_double2(x, y);
}

// ------------------------------------------------------------- cloud filter (a8/a9)
// Ordered compaction of an existing cloud in two streaming passes with no chain between tiles: k_filter_count (predicates
// only: the three coordinate planes, or just the mask bytes) leaves one count per 2048-point chunk, k_filter_scan turns the
// counts into offsets, k_filter_write re-evaluates the predicate and places the kept points.  The coordinates are read
// twice (12 of 24 + 24 * kept bytes per point extra), in exchange both passes run at streaming bandwidth; the single-pass
// decoupled look-back this replaced spent its time in the 16 k-hop chain of a 32 M-point cloud (0.22 of the HBM peak).
struct FilterArgs {
  const void *in;
  void *out;
  long long in_stride, out_stride, n;
  unsigned long long *count;
  unsigned int *chunk_count;         // [chunks]
  unsigned long long *chunk_offset;  // [chunks]
  long long chunks;
  int has_color;
  double z_min, z_max, r2_thresh, amin[3], amax[3];
  int use_zclip, use_radius, use_aabb;
  const uint8_t *keep;   // optional per-point mask (rv_select_by_mask)
  long long *index_out;  // optional: source index of every kept point
};
constexpr int kFilterChunk = 2048;  // points per warp and chunk: eight rounds of eight 32-point groups

constexpr int kFilterBatch = 8;  // 32-point groups whose loads are issued together (24 coordinate loads in flight per lane)

// predicates of kFilterBatch groups starting at point i0 (lane's first point): all loads first, then the decisions
template <typename T>
__device__ __forceinline__ void filter_batch(const FilterArgs &a, const T *in, long long i0, T (&x)[kFilterBatch],
                                             T (&y)[kFilterBatch], T (&z)[kFilterBatch], bool (&ok)[kFilterBatch]) {
  const bool need_xyz = (a.use_zclip | a.use_radius | a.use_aabb) != 0;
  uint8_t m[kFilterBatch];
#pragma unroll
  for (int j = 0; j < kFilterBatch; ++j) {
    const long long i = i0 + j * 32;
    const bool in_range = i < a.n;
    x[j] = y[j] = z[j] = (T)0;
    m[j] = 1;
    if (in_range) {
      if (need_xyz) {
        x[j] = in[i];
        y[j] = in[a.in_stride + i];
        z[j] = in[2 * a.in_stride + i];
      }
      if (a.keep) m[j] = a.keep[i];
    }
    ok[j] = in_range;
  }
#pragma unroll
  for (int j = 0; j < kFilterBatch; ++j) {
    bool k = ok[j] && m[j] != 0;
    if (need_xyz) {
      const double X = (double)x[j], Y = (double)y[j], Z = (double)z[j];
      if (a.use_zclip) k = k && (Z >= a.z_min) && (Z <= a.z_max);
      if (a.use_radius) k = k && (((X * X + Y * Y) + Z * Z) < a.r2_thresh);
      if (a.use_aabb)
        k = k && (X >= a.amin[0]) && (X <= a.amax[0]) && (Y >= a.amin[1]) && (Y <= a.amax[1]) && (Z >= a.amin[2]) &&
            (Z <= a.amax[2]);
    }
    ok[j] = k;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_filter_count(const FilterArgs a) {
  const T *in = reinterpret_cast<const T *>(a.in);
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < a.chunks; c += warps) {
    uint32_t cnt = 0;
    for (int r = 0; r < kFilterChunk / (32 * kFilterBatch); ++r) {
      T x[kFilterBatch], y[kFilterBatch], z[kFilterBatch];
      bool ok[kFilterBatch];
      filter_batch(a, in, c * kFilterChunk + r * 32 * kFilterBatch + lane, x, y, z, ok);
#pragma unroll
      for (int j = 0; j < kFilterBatch; ++j) cnt += __popc(__ballot_sync(0xffffffffu, ok[j]));
    }
    if (lane == 0) a.chunk_count[c] = cnt;
  }
}

// exclusive prefix of the chunk counts: one CTA walks the counts 1024 at a time (coalesced), carrying the running total
__global__ void __launch_bounds__(1024) k_filter_scan(const FilterArgs a) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (long long c0 = 0; c0 < a.chunks; c0 += 1024) {
    const long long c = c0 + threadIdx.x;
    const unsigned long long v = c < a.chunks ? a.chunk_count[c] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += up;
      }
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned long long carry = s_carry;
    const unsigned long long before = carry + (warp ? s_warp[warp - 1] : 0ull) + incl - v;
    if (c < a.chunks) a.chunk_offset[c] = before;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *a.count = s_carry;
}

template <typename T>
__global__ void __launch_bounds__(256) k_filter_write(const FilterArgs a) {
  const T *in = reinterpret_cast<const T *>(a.in);
  T *out = reinterpret_cast<T *>(a.out);
  const int lane = threadIdx.x & 31;
  const uint32_t lt = rv_lanemask_lt();
  const bool have_xyz = (a.use_zclip | a.use_radius | a.use_aabb) != 0;  // filter_batch loaded the coordinates
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < a.chunks; c += warps) {
    if (a.chunk_count[c] == 0) continue;  // warp-uniform
    unsigned long long run = a.chunk_offset[c];
    for (int r = 0; r < kFilterChunk / (32 * kFilterBatch); ++r) {
      T x[kFilterBatch], y[kFilterBatch], z[kFilterBatch];
      bool ok[kFilterBatch];
      const long long i0 = c * kFilterChunk + r * 32 * kFilterBatch + lane;
      filter_batch(a, in, i0, x, y, z, ok);
      // everything still to be read for the kept points of the batch is requested before the first store
      T cr[kFilterBatch], cg[kFilterBatch], cb[kFilterBatch];
#pragma unroll
      for (int j = 0; j < kFilterBatch; ++j) {
        const long long i = i0 + j * 32;
        cr[j] = cg[j] = cb[j] = (T)0;
        if (ok[j]) {
          if (!have_xyz) {
            x[j] = in[i];
            y[j] = in[a.in_stride + i];
            z[j] = in[2 * a.in_stride + i];
          }
          if (a.has_color) {
            cr[j] = in[3 * a.in_stride + i];
            cg[j] = in[4 * a.in_stride + i];
            cb[j] = in[5 * a.in_stride + i];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kFilterBatch; ++j) {
        const uint32_t bal = __ballot_sync(0xffffffffu, ok[j]);
        if (ok[j]) {
          const unsigned long long pos = run + __popc(bal & lt);
          out[pos] = x[j];
          out[a.out_stride + pos] = y[j];
          out[2 * a.out_stride + pos] = z[j];
          if (a.has_color) {
            out[3 * a.out_stride + pos] = cr[j];
            out[4 * a.out_stride + pos] = cg[j];
            out[5 * a.out_stride + pos] = cb[j];
          }
          if (a.index_out) a.index_out[pos] = i0 + j * 32;
        }
        run += __popc(bal);
      }
    }
  }
}

// shared launcher of rv_filter_cloud / rv_select_by_mask
template <typename T>
static void filter_launch(rv_ctx *ctx, const FilterArgs &a, cudaStream_t st) {
  const int per_cta = 256 / 32;
  long long blocks = (a.chunks + per_cta - 1) / per_cta;
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  k_filter_count<T><<<(int)blocks, 256, 0, st>>>(a);
  ctx->launches++;
  k_filter_scan<<<1, 1024, 0, st>>>(a);
  ctx->launches++;
  k_filter_write<T><<<(int)blocks, 256, 0, st>>>(a);
}

// Largest double s with RN(sqrt(s)) < r  <=>  s < T: T is the smallest double whose
// correctly rounded square root reaches r.  sqrt is monotone, so a short search around r*r finds it.
double radius_threshold(double r) {
  if (!(r > 0.0)) return 0.0;
  double t = r * r;
  while (sqrt(t) >= r) t = nextafter(t, 0.0);
  while (sqrt(t) < r) t = nextafter(t, INFINITY);
  return t;  // first value whose sqrt is >= r
}

// smallest float >= d / largest float <= d (NaN stays NaN: every comparison with it is false, as in float64)
float float_at_least(double d) {
  float f = (float)d;
  if ((double)f < d) f = nextafterf(f, INFINITY);
  return f;
}
float float_at_most(double d) {
  float f = (float)d;
  if ((double)f > d) f = nextafterf(f, -INFINITY);
  return f;
}

template <typename OutT, int DK>
void launch_mode(rv_ctx *ctx, const DeprojArgs &a, int mode, cudaStream_t st) {
#define RV_GO(M)                                                                                   \
  {                                                                                                \
    auto k = k_deproject<OutT, DK, M>;                                                             \
    const int grid = rv_persistent_grid(ctx, k, kThreads, 0, a.total_tiles);                       \
    k<<<grid, kThreads, 0, st>>>(a);                                                               \
  }
  switch (mode) {
    case RV_MODE_COMPACT_ORDERED: RV_GO(RV_MODE_COMPACT_ORDERED) break;
    case RV_MODE_COMPACT_UNORDERED: RV_GO(RV_MODE_COMPACT_UNORDERED) break;
    case RV_MODE_COMPACT_PACKED: RV_GO(RV_MODE_COMPACT_PACKED) break;
    case RV_MODE_DENSE_ZERO: RV_GO(RV_MODE_DENSE_ZERO) break;
    default: RV_GO(RV_MODE_DENSE_NAN) break;
  }
#undef RV_GO
}

}  // namespace

extern "C" {

int rv_depth_to_meters(rv_ctx *ctx, const uint16_t *d_depth, int64_t n, int unit_rule, double unit_scale, float *d_out,
                       rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (n < 0 || (n > 0 && (!d_depth || !d_out))) RV_FAIL(ctx, RV_EINVAL, "rv_depth_to_meters: null pointer or n<0");
  if (unit_rule < RV_UNIT_MUL_F32 || unit_rule > RV_UNIT_DIV_F64 || !(unit_scale > 0.0))
    RV_FAIL(ctx, RV_EINVAL, "rv_depth_to_meters: bad unit rule / scale");
  if (n == 0) return RV_OK;
  const float sf = (float)unit_scale;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  k_depth_to_meters<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_depth, n, unit_rule, unit_scale, 1.0 / unit_scale, sf,
                                                                  1.0f / sf, d_out);
  RV_LAUNCHED(ctx);
  return RV_OK;
}

int rv_build_ray_table(rv_ctx *ctx, const RvCam *cam, double *d_table, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (!cam || !d_table || cam->width <= 0 || cam->height <= 0) RV_FAIL(ctx, RV_EINVAL, "rv_build_ray_table: bad camera");
  if (!rv_aligned(d_table, 16)) RV_FAIL(ctx, RV_EALIGN, "rv_build_ray_table: table must be 16-byte aligned");
  const int n = cam->width * cam->height;
  k_ray_table<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*cam, reinterpret_cast<double2 *>(d_table));
  RV_LAUNCHED(ctx);
  return RV_OK;
}

size_t rv_deproject_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 128;
  const long long P = (long long)H * W;
  const long long tiles = (P + kMinTilePx - 1) / kMinTilePx * B;
  return 128 + (size_t)tiles * 8;
}

int rv_deproject_mask(rv_ctx *ctx, const void *d_depth, const uint8_t *d_bgr, const uint8_t *d_mask,
                      const double *d_ray_table, int B, int H, int W, const RvDeprojectParams *p, void *d_out,
                      int64_t plane_stride, int64_t frame_stride, uint8_t *d_valid, int32_t *d_src_index,
                      int64_t *d_counts, void *d_ws, size_t ws_bytes, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (!p) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: params is null");
  if (B < 0 || H <= 0 || W <= 0 || (long long)H * W > 0x7fffffffll)
    RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: bad shape B=%d H=%d W=%d", B, H, W);
  if (B == 0) return RV_OK;
  if (!d_depth || !d_out || !d_counts) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: depth/out/counts must be non-null");
  if (p->depth_kind != RV_DEPTH_U16 && p->depth_kind != RV_DEPTH_F32_METERS)
    RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: bad depth_kind %d", p->depth_kind);
  if (p->out_dtype != RV_F32 && p->out_dtype != RV_F64) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: bad out_dtype");
  if (p->mode < RV_MODE_COMPACT_ORDERED || p->mode > RV_MODE_COMPACT_PACKED) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: bad mode");
  if (p->kernel_select < RV_KERNEL_AUTO || p->kernel_select > RV_KERNEL_TMA) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: bad kernel_select");
  if (p->depth_kind == RV_DEPTH_U16 &&
      (p->unit_rule < RV_UNIT_MUL_F32 || p->unit_rule > RV_UNIT_DIV_F64 || !(p->unit_scale > 0.0)))
    RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: bad unit rule / scale");
  if (p->use_seg_mask && !d_mask) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: use_seg_mask set but mask is null");
  if (p->color_scale < RV_COLOR_UNIT || p->color_scale > RV_COLOR_PACKED8) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: bad color_scale");
  if (p->color_scale == RV_COLOR_PACKED8 && p->out_dtype != RV_F32)
    RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: RV_COLOR_PACKED8 needs float32 output planes (a colour word per point)");
  if (p->color_format != RV_COLORFMT_BGR8 && p->color_format != RV_COLORFMT_NV12) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: bad color_format");
  if (p->geometry != RV_GEOM_REFERENCE_F64 && p->geometry != RV_GEOM_SDK_F32) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: bad geometry");
  if (p->geometry == RV_GEOM_SDK_F32 && p->cam.model != RV_DIST_NONE && p->cam.model != RV_DIST_MODIFIED_BROWN_CONRADY)
    RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: RV_GEOM_SDK_F32 is restated for pinhole deprojection only");
  if (p->geometry == RV_GEOM_SDK_F32 && p->kernel_select == RV_KERNEL_TMA)
    RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: RV_GEOM_SDK_F32 runs on the generic kernel");
  if (p->color_format == RV_COLORFMT_NV12 && d_bgr && ((H & 1) || (W & 1)))
    RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: NV12 frames need even H and W");
  if (p->cam.model != RV_DIST_NONE && p->cam.model != RV_DIST_MODIFIED_BROWN_CONRADY && !d_ray_table)
    RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: distorted camera needs a ray table (rv_build_ray_table)");
  if (!(p->cam.fx != 0.0) || !(p->cam.fy != 0.0)) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: fx/fy must be non-zero");
  const long long P = (long long)H * W;
  const bool dense = p->mode == RV_MODE_DENSE_ZERO || p->mode == RV_MODE_DENSE_NAN;
  const bool packed = p->mode == RV_MODE_COMPACT_PACKED;
  if (packed ? plane_stride <= 0 : (frame_stride <= 0 || plane_stride < (int64_t)B * frame_stride))
    RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: plane_stride %lld < B*frame_stride", (long long)plane_stride);
  if (packed && (long long)B * H * W > 0xffffffffll) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: packed batches hold < 2^32 pixels");
  if (dense && frame_stride < P) RV_FAIL(ctx, RV_ECAPACITY, "rv_deproject_mask: dense mode needs frame_stride >= H*W");
  if (d_ray_table && !rv_aligned(d_ray_table, 16)) RV_FAIL(ctx, RV_EALIGN, "rv_deproject_mask: ray table alignment");
  const size_t need = rv_deproject_workspace_bytes(B, H, W);
  // every mode takes the workspace: the ordered modes keep their prefix chains in it and the TMA pipeline hands
  // out tiles through its ticket counter
  if (!d_ws || ws_bytes < need) RV_FAIL(ctx, RV_EWORKSPACE, "rv_deproject_mask: workspace %zu < %zu", ws_bytes, need);
  if (!rv_aligned(d_ws, 8)) RV_FAIL(ctx, RV_EALIGN, "rv_deproject_mask: workspace alignment");
  cudaStream_t st = (cudaStream_t)stream;

  DeprojArgs a;
  memset(&a, 0, sizeof(a));
  a.depth = d_depth;
  a.bgr = d_bgr;
  a.mask = p->use_seg_mask ? d_mask : nullptr;
  const bool use_rays = p->cam.model != RV_DIST_NONE && p->cam.model != RV_DIST_MODIFIED_BROWN_CONRADY;
  a.rays = use_rays ? reinterpret_cast<const double2 *>(d_ray_table) : nullptr;
  a.out = d_out;
  a.valid = d_valid;
  a.src_index = d_src_index;
  a.counts = reinterpret_cast<unsigned long long *>(d_counts);
  a.ticket = reinterpret_cast<unsigned int *>(d_ws);
  a.status = d_ws ? reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(d_ws) + 128) : nullptr;
  a.B = B;
  a.H = H;
  a.W = W;
  a.P = (int)P;
  if ((P + kMinTilePx - 1) / kMinTilePx * (long long)B > 0x7fffffffll) RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: batch too large");
  a.plane_stride = plane_stride;
  a.frame_stride = frame_stride;
  a.cx = p->cam.cx;
  a.cy = p->cam.cy;
  a.fx = p->cam.fx;
  a.fy = p->cam.fy;
  a.rfx = 1.0 / p->cam.fx;
  a.rfy = 1.0 / p->cam.fy;
  a.unit_rule = p->unit_rule;
  a.unit_scale = p->unit_scale;
  a.unit_rcp = 1.0 / p->unit_scale;
  a.unit_scale_f = (float)p->unit_scale;
  a.unit_rcp_f = 1.0f / a.unit_scale_f;
  a.trunc_f = (float)p->depth_trunc;
  a.z_min = p->z_min;
  a.z_max = p->z_max;
  a.r2_thresh = radius_threshold(p->r_max);
  for (int i = 0; i < 3; ++i) {
    a.amin[i] = p->aabb_min[i];
    a.amax[i] = p->aabb_max[i];
  }
  a.use_mask = p->use_seg_mask ? 1 : 0;
  a.invert_mask = p->invert_mask ? 1 : 0;
  a.use_trunc = p->use_depth_trunc ? 1 : 0;
  a.use_zclip = p->use_zclip ? 1 : 0;
  a.use_radius = p->use_radius ? 1 : 0;
  a.use_aabb = p->use_aabb ? 1 : 0;
  a.color_255 = p->color_scale == RV_COLOR_255;
  a.color_packed = (d_bgr && p->color_scale == RV_COLOR_PACKED8) ? 1 : 0;
  a.color_nv12 = (d_bgr && p->color_format == RV_COLORFMT_NV12) ? 1 : 0;
  a.geom_f32 = p->geometry == RV_GEOM_SDK_F32 ? 1 : 0;
  a.cx_f = (float)p->cam.cx;
  a.cy_f = (float)p->cam.cy;
  a.fx_f = (float)p->cam.fx;
  a.fy_f = (float)p->cam.fy;
  a.rfx_f = 1.0f / a.fx_f;
  a.rfy_f = 1.0f / a.fy_f;
  a.zmin_f = float_at_least(a.z_min);
  a.zmax_f = float_at_most(a.z_max);
  for (int i = 0; i < 3; ++i) {
    a.amin_f[i] = float_at_least(a.amin[i]);
    a.amax_f[i] = float_at_most(a.amax[i]);
  }
  a.fast_radius = a.r2_thresh > 1e-30 && a.r2_thresh < 1e30;
  a.r2_lo_f = a.fast_radius ? float_at_most(a.r2_thresh * (1.0 - 9.5367431640625e-07)) : 0.0f;   // 2^-20 band
  a.r2_hi_f = a.fast_radius ? float_at_least(a.r2_thresh * (1.0 + 9.5367431640625e-07)) : 0.0f;
  // smallest raw depth whose z alone already fails the float32 sphere test (z*z >= r2_hi): monotone in d, bisection
  a.d_cand = 65536u;
  if (a.use_radius && a.fast_radius && p->depth_kind == RV_DEPTH_U16) {
    auto z_of = [&](unsigned d) -> float {
      if (a.unit_rule == RV_UNIT_MUL_F32) return (float)d * a.unit_scale_f;
      if (a.unit_rule == RV_UNIT_DIV_F32) return (float)d / a.unit_scale_f;
      return (float)((double)d / a.unit_scale);
    };
    unsigned lo = 0, hi = 65536;  // invariant: beyond(lo) false (or lo == 0), beyond(hi) true (or hi == 65536)
    while (hi - lo > 1) {
      const unsigned mid = (lo + hi) / 2;
      const float z = z_of(mid);
      if (z * z >= a.r2_hi_f) hi = mid; else lo = mid;
    }
    a.d_cand = hi;
  }
  const bool fast_ok = rv_deproject_fast_eligible(a, p->mode);
  if (p->kernel_select == RV_KERNEL_TMA && !fast_ok)
    RV_FAIL(ctx, RV_EINVAL, "rv_deproject_mask: RV_KERNEL_TMA requested but the inputs are not eligible "
                            "(H*W %% 16 == 0, W >= 32, 16-byte aligned inputs)");
  const bool use_fast = fast_ok && p->kernel_select != RV_KERNEL_GENERIC && !a.geom_f32;
  // COMPACT_UNORDERED asks for less than COMPACT_ORDERED delivers: on the fast path it simply gets the ordered kernel
  const int mode = (use_fast && p->mode == RV_MODE_COMPACT_UNORDERED) ? RV_MODE_COMPACT_ORDERED : p->mode;
  const bool ordered = mode == RV_MODE_COMPACT_ORDERED || packed;
  const int tile_px = use_fast ? kFastTilePx : kTile;
  a.tiles_per_frame = (int)((P + tile_px - 1) / tile_px);
  a.total_tiles = a.tiles_per_frame * B;

  if (ordered) {
    RV_CUDA(ctx, cudaMemsetAsync(d_ws, 0, 128 + (size_t)a.total_tiles * 8, st));  // ticket counter + one status word per tile
  } else {
    RV_CUDA(ctx, cudaMemsetAsync(d_ws, 0, 128, st));  // ticket counter
    RV_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)B * sizeof(int64_t), st));
  }
  if (use_fast) {
    RV_CUDA(ctx, rv_deproject_fast_launch(ctx, a, mode, p->out_dtype, p->depth_kind, st));
  } else if (p->out_dtype == RV_F32) {
    if (p->depth_kind == RV_DEPTH_U16)
      launch_mode<float, RV_DEPTH_U16>(ctx, a, p->mode, st);
    else
      launch_mode<float, RV_DEPTH_F32_METERS>(ctx, a, p->mode, st);
  } else {
    if (p->depth_kind == RV_DEPTH_U16)
      launch_mode<double, RV_DEPTH_U16>(ctx, a, p->mode, st);
    else
      launch_mode<double, RV_DEPTH_F32_METERS>(ctx, a, p->mode, st);
  }
  RV_LAUNCHED(ctx);
  return RV_OK;
}

size_t rv_filter_workspace_bytes(int64_t n) {
  if (n <= 0) return 128;
  const size_t chunks = (size_t)((n + kFilterChunk - 1) / kFilterChunk);
  return 128 + ((chunks * 4 + 127) & ~(size_t)127) + chunks * 8;
}

int rv_filter_cloud(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int dtype, int has_color,
                    const RvDeprojectParams *p, void *d_out, int64_t out_plane_stride, int64_t *d_count, int64_t *d_index,
                    void *d_ws, size_t ws_bytes, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (!p || !d_count) RV_FAIL(ctx, RV_EINVAL, "rv_filter_cloud: null params/count");
  if (n < 0 || in_plane_stride < n) RV_FAIL(ctx, RV_EINVAL, "rv_filter_cloud: bad n / stride");
  // the kept count is only known on the device: every plane must be able to take all n points
  if (out_plane_stride < n) RV_FAIL(ctx, RV_ECAPACITY, "rv_filter_cloud: out_plane_stride %lld < n %lld", (long long)out_plane_stride, (long long)n);
  if (dtype != RV_F32 && dtype != RV_F64) RV_FAIL(ctx, RV_EINVAL, "rv_filter_cloud: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    RV_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    return RV_OK;
  }
  if (!d_in || !d_out) RV_FAIL(ctx, RV_EINVAL, "rv_filter_cloud: null cloud pointer");
  const size_t need = rv_filter_workspace_bytes(n);
  if (!d_ws || ws_bytes < need) RV_FAIL(ctx, RV_EWORKSPACE, "rv_filter_cloud: workspace %zu < %zu", ws_bytes, need);
  FilterArgs a;
  memset(&a, 0, sizeof(a));
  a.in = d_in;
  a.out = d_out;
  a.in_stride = in_plane_stride;
  a.out_stride = out_plane_stride;
  a.n = n;
  a.count = reinterpret_cast<unsigned long long *>(d_count);
  a.chunks = (n + kFilterChunk - 1) / kFilterChunk;
  a.chunk_count = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(d_ws) + 128);
  a.chunk_offset = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(d_ws) + 128 + (((size_t)a.chunks * 4 + 127) & ~(size_t)127));
  a.has_color = has_color ? 1 : 0;
  a.z_min = p->z_min;
  a.z_max = p->z_max;
  a.r2_thresh = radius_threshold(p->r_max);
  for (int i = 0; i < 3; ++i) {
    a.amin[i] = p->aabb_min[i];
    a.amax[i] = p->aabb_max[i];
  }
  a.use_zclip = p->use_zclip ? 1 : 0;
  a.use_radius = p->use_radius ? 1 : 0;
  a.use_aabb = p->use_aabb ? 1 : 0;
  a.index_out = reinterpret_cast<long long *>(d_index);
  if (dtype == RV_F32) filter_launch<float>(ctx, a, st);
  else filter_launch<double>(ctx, a, st);
  RV_LAUNCHED(ctx);
  return RV_OK;
}

int rv_select_by_mask(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int dtype, int has_color,
                      const uint8_t *d_keep, void *d_out, int64_t out_plane_stride, int64_t *d_count, int64_t *d_index,
                      void *d_ws, size_t ws_bytes, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (!d_keep || !d_count) RV_FAIL(ctx, RV_EINVAL, "rv_select_by_mask: null mask/count");
  if (n < 0 || in_plane_stride < n) RV_FAIL(ctx, RV_EINVAL, "rv_select_by_mask: bad n / stride");
  if (out_plane_stride < n) RV_FAIL(ctx, RV_ECAPACITY, "rv_select_by_mask: out_plane_stride %lld < n %lld", (long long)out_plane_stride, (long long)n);
  if (dtype != RV_F32 && dtype != RV_F64) RV_FAIL(ctx, RV_EINVAL, "rv_select_by_mask: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    RV_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    return RV_OK;
  }
  if (!d_in || !d_out) RV_FAIL(ctx, RV_EINVAL, "rv_select_by_mask: null cloud pointer");
  const size_t need = rv_filter_workspace_bytes(n);
  if (!d_ws || ws_bytes < need) RV_FAIL(ctx, RV_EWORKSPACE, "rv_select_by_mask: workspace %zu < %zu", ws_bytes, need);
  FilterArgs a;
  memset(&a, 0, sizeof(a));
  a.in = d_in;
  a.out = d_out;
  a.in_stride = in_plane_stride;
  a.out_stride = out_plane_stride;
  a.n = n;
  a.count = reinterpret_cast<unsigned long long *>(d_count);
  a.chunks = (n + kFilterChunk - 1) / kFilterChunk;
  a.chunk_count = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(d_ws) + 128);
  a.chunk_offset = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(d_ws) + 128 + (((size_t)a.chunks * 4 + 127) & ~(size_t)127));
  a.has_color = has_color ? 1 : 0;
  a.keep = d_keep;
  a.index_out = reinterpret_cast<long long *>(d_index);
  if (dtype == RV_F32) filter_launch<float>(ctx, a, st);
  else filter_launch<double>(ctx, a, st);
  RV_LAUNCHED(ctx);
  return RV_OK;
}

}  // extern "C"
