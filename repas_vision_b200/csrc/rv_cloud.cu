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
constexpr int kXformViews = 8;  // views per launch (blockIdx.y); the four_pose_captures fusion is one launch
struct XformView {
  const void *in;
  long long in_stride, n, out_offset;
  double T[16];
};
struct XformArgs {
  XformView view[kXformViews];
  void *out;
  long long out_stride;
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
__global__ void __launch_bounds__(256) k_transform(const __grid_constant__ XformArgs args) {
  const XformView &a = args.view[blockIdx.y];
  const InT *__restrict__ in = reinterpret_cast<const InT *>(a.in);  // the merged cloud never overlaps a view
  OutT *__restrict__ out = reinterpret_cast<OutT *>(args.out) + a.out_offset;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 4
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
    out[args.out_stride + i] = oy;
    out[2 * args.out_stride + i] = oz;
    if (args.has_color) {
      out[3 * args.out_stride + i] = (OutT)in[3 * a.in_stride + i];
      out[4 * args.out_stride + i] = (OutT)in[4 * a.in_stride + i];
      out[5 * args.out_stride + i] = (OutT)in[5 * a.in_stride + i];
    }
    if (args.bounds) {  // bounds of the STORED merged cloud
      const double sx = (double)ox, sy = (double)oy, sz = (double)oz;
      lo[0] = sx < lo[0] ? sx : lo[0];
      lo[1] = sy < lo[1] ? sy : lo[1];
      lo[2] = sz < lo[2] ? sz : lo[2];
      hi[0] = sx > hi[0] ? sx : hi[0];
      hi[1] = sy > hi[1] ? sy : hi[1];
      hi[2] = sz > hi[2] ? sz : hi[2];
    }
  }
  if (args.bounds) block_bounds_commit(lo, hi, args.bounds);
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

// min / max / sum of the coordinates in one pass (get_min_bound, get_max_bound, get_center)
template <typename T>
__global__ void __launch_bounds__(256) k_cloud_stats(const T *__restrict__ in, long long stride_in, long long n, double *stats) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf}, sum[3] = {0.0, 0.0, 0.0};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double v = (double)in[a * stride_in + i];
      lo[a] = v < lo[a] ? v : lo[a];
      hi[a] = v > hi[a] ? v : hi[a];
      sum[a] += v;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum[a] += __shfl_xor_sync(0xffffffffu, sum[a], o);
    if ((threadIdx.x & 31) == 0) atomicAdd(stats + 6 + a, sum[a]);
  }
  block_bounds_commit(lo, hi, stats);
}

__global__ void k_stats_init(double *b) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  if (threadIdx.x < 3) b[threadIdx.x] = inf;
  else if (threadIdx.x < 6) b[threadIdx.x] = -inf;
  else if (threadIdx.x < 9) b[threadIdx.x] = 0.0;
}

// --------------------------------------------------------------------- voxel grid
// Workspace: [header 256 B][per-part counters: 2048 x 8 B][keys: capacity x 8 B][record index: capacity x 4 B][records: n x 64 B][list: n x 4 B]
//
// The hash table holds only 8-byte keys (1.5 n slots, any size: the hash is mapped with a multiply-high), so clearing it
// and probing it touch an eighth of the bytes a table of full records would and the whole table stays in the L2; the record
// index of each slot lives in a parallel array that is written by the slot's creator and never cleared.  Points arrive in
// pixel order, so consecutive points usually share a voxel: each warp first folds runs of equal keys with a segmented
// shuffle reduction, and only the head of a run goes to memory.  A run head
//   * writes its partial sums to the record with its own point index (plain stores, no allocation counter),
//   * tries to claim the slot of its key with ONE compare-and-swap: the winner's record becomes the voxel's record
//     ("creator"); a head that finds the key already there is a "joiner",
//   * creators and joiners are appended to the two ends of the CTA's own stretch of the list (the cloud is cut into one
//     contiguous part per CTA), positions from shared-memory counters: no global counter anywhere.
// k_voxel_merge then adds every joiner's record into its voxel's record with float64 atomics (the kernel boundary is
// the only ordering the scheme needs: no fences, no spinning on another thread's publication), and k_voxel_emit writes
// part p's creators behind the creators of parts 0..p-1 (a 2048-entry prefix sum per CTA).
struct VoxHeader {
  double bounds[6];
  int error;  // 1: a voxel index does not fit 21 bits / extent check failed
  int pad;
};
constexpr int kVoxMaxParts = 2048;
constexpr size_t kVoxHead = 256 + (size_t)kVoxMaxParts * 8;  // header + per-part {creators, joiners}
struct __align__(64) VoxAcc {
  unsigned long long key;
  unsigned int count;
  unsigned int pad;
  double sum[6];
};
static_assert(sizeof(VoxAcc) == 64, "record layout");

struct VoxArgs {
  const void *in;
  long long in_stride, n;
  int has_color;
  double voxel, rvoxel;
  const double *bounds;  // device
  VoxHeader *hdr;
  unsigned long long *keys;  // 0 = empty, else packed(ix,iy,iz) + 1
  unsigned int *slot_rec;    // record index of the slot's creator (valid where keys[] != 0 once k_voxel_insert is done)
  VoxAcc *acc;
  unsigned int *list;  // part p owns list[p * span, ...): creators from its front, joiners from its back (record indices)
  uint2 *parts;        // per part {creators, joiners}
  long long span;      // points per part, a multiple of 32
  unsigned int cap;
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
  const int lane = threadIdx.x & 31;
  __shared__ unsigned int s_create, s_join;
  if (threadIdx.x == 0) s_create = s_join = 0;
  __syncthreads();
  const long long p0 = (long long)blockIdx.x * a.span;                  // this CTA's part of the cloud
  const long long p1 = p0 + a.span < a.n ? p0 + a.span : a.n;           // (empty when p0 >= n)
  for (long long i = p0 + threadIdx.x; i < p0 + a.span && i - lane < p1; i += blockDim.x) {  // whole warps: full shuffles
    const bool valid = i < p1;
    double v[6] = {0, 0, 0, 0, 0, 0};
    unsigned long long key = 0;  // lanes past the end form one run that is never written
    if (valid) {
      v[0] = (double)in[i], v[1] = (double)in[a.in_stride + i], v[2] = (double)in[2 * a.in_stride + i];
      if (a.has_color) {
        v[3] = (double)in[3 * a.in_stride + i];
        v[4] = (double)in[4 * a.in_stride + i];
        v[5] = (double)in[5 * a.in_stride + i];
      }
      // key = floor((p - origin) / voxel), IEEE division via the exact-reciprocal helper
      const long long ix = (long long)floor(rv_div(v[0] - ox, a.voxel, a.rvoxel));
      const long long iy = (long long)floor(rv_div(v[1] - oy, a.voxel, a.rvoxel));
      const long long iz = (long long)floor(rv_div(v[2] - oz, a.voxel, a.rvoxel));
      key = (((unsigned long long)ix & 0x1fffff) << 42 | ((unsigned long long)iy & 0x1fffff) << 21 | ((unsigned long long)iz & 0x1fffff)) + 1ull;
    }
    // ---- fold runs of equal keys inside the warp (segmented reduction towards the first lane of each run)
    const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane == 0 || key != prev;
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const uint32_t after = lane == 31 ? 0u : (heads >> (lane + 1));
    const int run = after ? __ffs(after) : 32 - lane;  // lanes from this one to the end of its run
    const int ncol = a.has_color ? 6 : 3;
    const int longest = __reduce_max_sync(0xffffffffu, run);  // steps needed: log2 of the longest run in the warp (often 2-4 points)
#pragma unroll 1
    for (int d = 1; d < longest; d <<= 1) {
      const bool take = d < run;  // lane + d still belongs to this lane's run
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        if (c < ncol) {
          const double o = __shfl_down_sync(0xffffffffu, v[c], d);
          if (take) v[c] += o;
        }
      }
    }
    // ---- one record and one slot probe per run
    const bool lead = head && valid;
    bool created = false;
    if (lead) {
      VoxAcc *acc = a.acc + i;
      acc->key = key;
      acc->count = (unsigned int)run;
#pragma unroll
      for (int c = 0; c < 6; ++c) acc->sum[c] = v[c];
      unsigned int h = __umulhi((unsigned int)(vox_hash(key) >> 32), a.cap);  // uniform over [0, cap)
      for (;;) {
        const unsigned long long cur = atomicCAS(a.keys + h, 0ull, key);  // the table is at most two thirds full: usually empty
        if (cur == 0) {
          a.slot_rec[h] = (unsigned int)i;  // read by the next kernel
          created = true;
          break;
        }
        if (cur == key) {
          acc->pad = h;  // joiner: where its voxel's slot is, so the merge pass does not probe again
          break;
        }
        if (++h == a.cap) h = 0;
      }
    }
    // ---- creators to the front of the part's stretch of the list, joiners to its back
    const uint32_t cb = __ballot_sync(0xffffffffu, created);
    const uint32_t jb = __ballot_sync(0xffffffffu, lead && !created);
    unsigned int cbase = 0, jbase = 0;
    if (lane == 0) {
      if (cb) cbase = atomicAdd(&s_create, (unsigned int)__popc(cb));
      if (jb) jbase = atomicAdd(&s_join, (unsigned int)__popc(jb));
    }
    cbase = __shfl_sync(0xffffffffu, cbase, 0);
    jbase = __shfl_sync(0xffffffffu, jbase, 0);
    if (created) a.list[p0 + cbase + __popc(cb & rv_lanemask_lt())] = (unsigned int)i;
    else if (lead) a.list[p1 - 1 - (long long)(jbase + __popc(jb & rv_lanemask_lt()))] = (unsigned int)i;
  }
  __syncthreads();
  if (threadIdx.x == 0) a.parts[blockIdx.x] = make_uint2(s_create, s_join);
}

// every joiner's record is added to the record of its voxel's creator; the joiners of all parts are dealt evenly to
// the threads of the grid (prefix sum of the per-part counts in shared memory, binary search per joiner)
__global__ void __launch_bounds__(256) k_voxel_merge(const VoxArgs a, int n_parts) {
  __shared__ unsigned int s_pre[kVoxMaxParts + 1];  // exclusive prefix of the joiner counts
  __shared__ unsigned int s_warp[8];
  // block-wide exclusive scan, eight parts per thread
  unsigned int v[8], sum = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int q = threadIdx.x * 8 + k;
    v[k] = q < n_parts ? a.parts[q].y : 0u;
    sum += v[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned int wbase = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if (w < warp) wbase += s_warp[w];
  unsigned int run = wbase + incl - sum;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s_pre[threadIdx.x * 8 + k] = run;
    run += v[k];
  }
  if (threadIdx.x == 255) s_pre[kVoxMaxParts] = run;
  __syncthreads();
  const unsigned int total = s_pre[kVoxMaxParts];
  const unsigned int stride = gridDim.x * blockDim.x;
  for (unsigned int j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
    int lo = 0, hi = kVoxMaxParts;  // last part whose prefix is <= j
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_pre[mid] <= j) lo = mid; else hi = mid;
    }
    const long long p0 = (long long)lo * a.span;
    const long long p1 = p0 + a.span < a.n ? p0 + a.span : a.n;
    const VoxAcc r = a.acc[a.list[p1 - 1 - (long long)(j - s_pre[lo])]];
    VoxAcc *acc = a.acc + a.slot_rec[r.pad];
    atomicAdd(&acc->count, r.count);
    atomicAdd(&acc->sum[0], r.sum[0]);
    atomicAdd(&acc->sum[1], r.sum[1]);
    atomicAdd(&acc->sum[2], r.sum[2]);
    if (a.has_color) {
      atomicAdd(&acc->sum[3], r.sum[3]);
      atomicAdd(&acc->sum[4], r.sum[4]);
      atomicAdd(&acc->sum[5], r.sum[5]);
    }
  }
}

struct VoxOutArgs {
  const VoxHeader *hdr;
  const VoxAcc *acc;
  const unsigned int *list;
  const uint2 *parts;
  long long span;
  void *out;
  long long out_stride, out_capacity;
  int32_t *keys;
  int32_t *counts;
  long long *m;
  int has_color;
};

// CTA p writes the voxels part p created, behind those of parts 0..p-1: mean = sum / count (one IEEE division for
// 1 / count, then the exact-quotient correction of rv_div per component)
template <typename OutT>
__global__ void __launch_bounds__(256) k_voxel_emit(const VoxOutArgs a) {
  OutT *out = reinterpret_cast<OutT *>(a.out);
  __shared__ unsigned long long s_red[2][8];
  unsigned long long before = 0, total = 0;
  for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
    const unsigned long long c = a.parts[q].x;
    total += c;
    if (q < (int)blockIdx.x) before += c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    before += __shfl_xor_sync(0xffffffffu, before, o);
    total += __shfl_xor_sync(0xffffffffu, total, o);
  }
  if ((threadIdx.x & 31) == 0) s_red[0][threadIdx.x >> 5] = before, s_red[1][threadIdx.x >> 5] = total;
  __syncthreads();
  before = total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) before += s_red[0][w], total += s_red[1][w];
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.m = a.hdr->error ? -1ll : (long long)total;
  if (a.hdr->error) return;
  const unsigned int mine = a.parts[blockIdx.x].x;
  const unsigned int *list = a.list + (long long)blockIdx.x * a.span;
  for (unsigned int k = threadIdx.x; k < mine; k += blockDim.x) {
    const long long o = (long long)(before + k);
    if (o >= a.out_capacity) break;
    const VoxAcc s = a.acc[list[k]];
    const double c = (double)s.count, rc = 1.0 / c;
    out[o] = (OutT)rv_div(s.sum[0], c, rc);
    out[a.out_stride + o] = (OutT)rv_div(s.sum[1], c, rc);
    out[2 * a.out_stride + o] = (OutT)rv_div(s.sum[2], c, rc);
    if (a.has_color) {
      out[3 * a.out_stride + o] = (OutT)rv_div(s.sum[3], c, rc);
      out[4 * a.out_stride + o] = (OutT)rv_div(s.sum[4], c, rc);
      out[5 * a.out_stride + o] = (OutT)rv_div(s.sum[5], c, rc);
    }
    if (a.keys) {
      const unsigned long long kk = s.key - 1ull;
      a.keys[o] = (int32_t)((kk >> 42) & 0x1fffff);
      a.keys[a.out_capacity + o] = (int32_t)((kk >> 21) & 0x1fffff);
      a.keys[2 * a.out_capacity + o] = (int32_t)(kk & 0x1fffff);
    }
    if (a.counts) a.counts[o] = (int32_t)s.count;
  }
}

unsigned long long vox_capacity(long long n) {  // 1.5 slots per point, a multiple of 32
  unsigned long long c = ((unsigned long long)n * 3ull / 2ull + 31ull) & ~31ull;
  return c < 1024 ? 1024 : c;
}

// --------------------------------------------------------------------- PLY records
// every warp gathers the 32 records of a step in shared memory and writes them out as 16-byte pieces (32 records are
// 480 or 864 bytes: a multiple of 16); byte-wise stores of 15-byte records reach the L2 as partial-sector writes and ran at
// an eighth of the bandwidth
template <typename InT, typename CoordT>
__global__ void __launch_bounds__(256) k_pack_ply(const InT *__restrict__ in, long long stride_in, long long n,
                                                  int has_color, int color_255, uint8_t *__restrict__ rec) {
  constexpr int kRec = 3 * (int)sizeof(CoordT) + 3;
  __shared__ __align__(16) uint8_t s_rec[8][32 * kRec];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool wide = (reinterpret_cast<uintptr_t>(rec) & 15) == 0;
  const long long warps = (long long)gridDim.x * 8;
  const long long groups = (n + 31) >> 5;
  for (long long g = (long long)blockIdx.x * 8 + warp; g < groups; g += warps) {
    const long long i = g * 32 + lane;
    uint8_t *buf = &s_rec[warp][lane * kRec];
    if (i < n) {
      CoordT xyz[3] = {(CoordT)in[i], (CoordT)in[stride_in + i], (CoordT)in[2 * stride_in + i]};
      uint8_t raw[3 * sizeof(CoordT)];
      memcpy(raw, xyz, sizeof(xyz));
#pragma unroll
      for (int k = 0; k < (int)sizeof(xyz); ++k) buf[k] = raw[k];
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
    }
    __syncwarp();
    const long long left = n - g * 32;
    const int bytes = (int)(left < 32 ? left : 32) * kRec;
    uint8_t *dst = rec + g * 32 * kRec;
    if (wide && bytes == 32 * kRec) {
      for (int k = lane; k < 32 * kRec / 16; k += 32)
        reinterpret_cast<uint4 *>(dst)[k] = reinterpret_cast<const uint4 *>(s_rec[warp])[k];
    } else {
      for (int k = lane; k < bytes; k += 32) dst[k] = s_rec[warp][k];
    }
    __syncwarp();
  }
}

struct UnpackArgs {
  const uint8_t *rec;
  long long n, out_stride;
  int record_bytes;
  int xyz_off[3], rgb_off[3];
  int has_color;
};

template <typename CoordT, typename OutT>
__global__ void __launch_bounds__(256) k_unpack_ply(const UnpackArgs a, OutT *__restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    const uint8_t *r = a.rec + i * a.record_bytes;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint8_t b[sizeof(CoordT)];  // records are packed: fields are not aligned
#pragma unroll
      for (int k = 0; k < (int)sizeof(CoordT); ++k) b[k] = r[a.xyz_off[c] + k];
      CoordT v;
      memcpy(&v, b, sizeof(CoordT));
      out[c * a.out_stride + i] = (OutT)v;
    }
    if (a.has_color) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double k = (double)r[a.rgb_off[c]];
        out[(3 + c) * a.out_stride + i] = (OutT)rv_div(k, 255.0, 0.00392156862745098);  // k / 255.0 in float64, as Open3D
      }
    }
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

int rv_cloud_stats(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int dtype, double *d_stats,
                   rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (!d_stats || n < 0 || in_plane_stride < n) RV_FAIL(ctx, RV_EINVAL, "rv_cloud_stats: bad n / stride / null output");
  if (dtype != RV_F32 && dtype != RV_F64) RV_FAIL(ctx, RV_EINVAL, "rv_cloud_stats: bad dtype");
  if (n > 0 && !d_in) RV_FAIL(ctx, RV_EINVAL, "rv_cloud_stats: null cloud pointer");
  cudaStream_t st = (cudaStream_t)stream;
  k_stats_init<<<1, 32, 0, st>>>(d_stats);
  RV_LAUNCHED(ctx);
  if (n == 0) return RV_OK;
  const int g = grid_for(ctx, n);
  if (dtype == RV_F32) k_cloud_stats<float><<<g, 256, 0, st>>>(reinterpret_cast<const float *>(d_in), in_plane_stride, n, d_stats);
  else k_cloud_stats<double><<<g, 256, 0, st>>>(reinterpret_cast<const double *>(d_in), in_plane_stride, n, d_stats);
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
  for (int v0 = 0; v0 < n_views; v0 += kXformViews) {
    XformArgs a;
    memset(&a, 0, sizeof(a));
    a.out = d_out;
    a.out_stride = out_plane_stride;
    a.bounds = d_bounds;
    a.has_color = has_color ? 1 : 0;
    int nv = 0;
    long long nmax = 0;
    for (int v = v0; v < n_views && v < v0 + kXformViews; ++v) {
      if (n[v] > 0) {
        if (!d_in[v]) RV_FAIL(ctx, RV_EINVAL, "rv_transform_merge: view %d is null", v);
        XformView &w = a.view[nv++];
        w.in = d_in[v];
        w.in_stride = in_plane_stride[v];
        w.n = n[v];
        w.out_offset = off;
        memcpy(w.T, T + 16 * v, sizeof(w.T));
        nmax = n[v] > nmax ? n[v] : nmax;
      }
      off += n[v];
    }
    if (nv == 0) continue;
    const dim3 grid((unsigned)grid_for(ctx, nmax, nv > 1 ? 4 : 8), (unsigned)nv);
    if (in_dtype == RV_F32 && out_dtype == RV_F32) k_transform<float, float><<<grid, 256, 0, st>>>(a);
    else if (in_dtype == RV_F32) k_transform<float, double><<<grid, 256, 0, st>>>(a);
    else if (out_dtype == RV_F32) k_transform<double, float><<<grid, 256, 0, st>>>(a);
    else k_transform<double, double><<<grid, 256, 0, st>>>(a);
    RV_LAUNCHED(ctx);
  }
  return RV_OK;
}

size_t rv_voxel_workspace_bytes(int64_t n) {
  if (n < 0) n = 0;
  return kVoxHead + (size_t)vox_capacity(n) * 12 + (size_t)n * sizeof(VoxAcc) + (((size_t)n * 4 + 63) & ~(size_t)63);
}

int rv_voxel_downsample(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int in_dtype, int has_color,
                        double voxel_size, const double *d_bounds, void *d_out, int64_t out_plane_stride, int out_dtype,
                        int64_t out_capacity, int32_t *d_keys, int32_t *d_counts_out, int64_t *d_m, void *d_ws,
                        size_t ws_bytes, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (!(voxel_size > 0.0)) RV_FAIL(ctx, RV_EINVAL, "rv_voxel_downsample: voxel_size <= 0");
  if (n >= 0xa0000000ll) RV_FAIL(ctx, RV_EINVAL, "rv_voxel_downsample: more than 2.6e9 points");  // 1.5 n slots in 32 bits
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
  uint2 *parts = reinterpret_cast<uint2 *>(reinterpret_cast<char *>(d_ws) + 256);
  char *w = reinterpret_cast<char *>(d_ws) + kVoxHead;
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(w);
  unsigned int *slot_rec = reinterpret_cast<unsigned int *>(w + cap * 8);
  VoxAcc *acc = reinterpret_cast<VoxAcc *>(w + cap * 12);  // cap % 32 == 0: 64-byte aligned
  unsigned int *list = reinterpret_cast<unsigned int *>(acc + n);
  RV_CUDA(ctx, cudaMemsetAsync(d_ws, 0, kVoxHead + cap * 8, st));  // header, counters, keys; everything else is written, not cleared
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
  a.keys = keys;
  a.slot_rec = slot_rec;
  a.acc = acc;
  a.list = list;
  a.cap = (unsigned int)cap;
  int parts_n;
  {
    auto kf = k_voxel_insert<float>;
    auto kd = k_voxel_insert<double>;
    // parts of at least 1024 points, up to 2048 of them: two to three waves of small CTAs keep the tail of the kernel short
    parts_n = (int)((n + 1023) / 1024 < kVoxMaxParts ? (n + 1023) / 1024 : kVoxMaxParts);
    if (parts_n < 1) parts_n = 1;
    a.span = (((n + parts_n - 1) / parts_n) + 31) & ~31ll;
    a.parts = parts;
    if (in_dtype == RV_F32) kf<<<parts_n, 256, 0, st>>>(a);
    else kd<<<parts_n, 256, 0, st>>>(a);
  }
  RV_LAUNCHED(ctx);
  k_voxel_merge<<<grid_for(ctx, n / 2 + 1, 8), 256, 0, st>>>(a, parts_n);  // joiners < points; grid-stride over the real count
  RV_LAUNCHED(ctx);
  VoxOutArgs o;
  memset(&o, 0, sizeof(o));
  o.hdr = hdr;
  o.acc = acc;
  o.list = list;
  o.parts = parts;
  o.span = a.span;
  o.out = d_out;
  o.out_stride = out_plane_stride;
  o.out_capacity = out_capacity;
  o.keys = d_keys;
  o.counts = d_counts_out;
  o.m = reinterpret_cast<long long *>(d_m);
  o.has_color = has_color ? 1 : 0;
  if (out_dtype == RV_F32) k_voxel_emit<float><<<parts_n, 256, 0, st>>>(o);
  else k_voxel_emit<double><<<parts_n, 256, 0, st>>>(o);
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
  const int g = grid_for(ctx, n);  // 8 warps per CTA, 32 records per warp and step
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

int rv_unpack_ply_records(rv_ctx *ctx, const uint8_t *d_records, int64_t n, int record_bytes, const int32_t *xyz_offset,
                          int coord_dtype, const int32_t *rgb_offset, void *d_out, int64_t out_plane_stride,
                          int out_dtype, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (n < 0 || out_plane_stride < n || record_bytes <= 0 || !xyz_offset) RV_FAIL(ctx, RV_EINVAL, "rv_unpack_ply_records: bad n / stride / record");
  if ((coord_dtype != RV_F32 && coord_dtype != RV_F64) || (out_dtype != RV_F32 && out_dtype != RV_F64))
    RV_FAIL(ctx, RV_EINVAL, "rv_unpack_ply_records: bad dtype");
  const int cw = coord_dtype == RV_F32 ? 4 : 8;
  for (int c = 0; c < 3; ++c) {
    if (xyz_offset[c] < 0 || xyz_offset[c] + cw > record_bytes) RV_FAIL(ctx, RV_EINVAL, "rv_unpack_ply_records: coordinate offset outside the record");
    if (rgb_offset && (rgb_offset[c] < 0 || rgb_offset[c] >= record_bytes)) RV_FAIL(ctx, RV_EINVAL, "rv_unpack_ply_records: colour offset outside the record");
  }
  if (n == 0) return RV_OK;
  if (!d_records || !d_out) RV_FAIL(ctx, RV_EINVAL, "rv_unpack_ply_records: null pointer");
  UnpackArgs a;
  memset(&a, 0, sizeof(a));
  a.rec = d_records;
  a.n = n;
  a.out_stride = out_plane_stride;
  a.record_bytes = record_bytes;
  for (int c = 0; c < 3; ++c) {
    a.xyz_off[c] = xyz_offset[c];
    a.rgb_off[c] = rgb_offset ? rgb_offset[c] : 0;
  }
  a.has_color = rgb_offset ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(ctx, n);
  if (coord_dtype == RV_F32 && out_dtype == RV_F32) k_unpack_ply<float, float><<<g, 256, 0, st>>>(a, reinterpret_cast<float *>(d_out));
  else if (coord_dtype == RV_F32) k_unpack_ply<float, double><<<g, 256, 0, st>>>(a, reinterpret_cast<double *>(d_out));
  else if (out_dtype == RV_F32) k_unpack_ply<double, float><<<g, 256, 0, st>>>(a, reinterpret_cast<float *>(d_out));
  else k_unpack_ply<double, double><<<g, 256, 0, st>>>(a, reinterpret_cast<double *>(d_out));
  RV_LAUNCHED(ctx);
  return RV_OK;
}

}  // extern "C"
