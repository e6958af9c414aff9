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
// Hash grid without per-point records (round 1 wrote a 64-byte record per run of points and read it back twice: 6.1 x the
// algorithmic DRAM traffic).  Workspace:
//   [header 256 B][per-part counters 2048 x 8 B][long-voxel bits: cap / 8 B][table: cap x 16 B {key, chain head}][list: n x 8 B][joiner slots: n x 4 B]
//   [head masks: ceil(n / 32) x 4 B][long-voxel pool: (n / 9 + 1) x 64 B][fusion only: transformed xyz, n x 24 B]
// k_vox_insert  reads the coordinates only (RV_VOX_GROUPS 32-point groups per warp and step, their loads and first-probe
//               compare-and-swaps issued before anything is consumed).  Points arrive in pixel order, so consecutive points
//               usually share a voxel: a run of equal keys inside a 32-point group is represented by its first point (the
//               "head"; the ballot of heads is kept per group, 1 bit per point).  A head claims its key's slot with ONE
//               compare-and-swap ("creator") or finds it taken ("joiner"); both append (point index, slot) to their part's
//               stretch of the list (creators from the front, joiners from the back; positions from shared-memory counters:
//               the cloud is cut into one contiguous part per CTA, no global counter anywhere); a joiner additionally links
//               itself into the slot's chain with one atomic exchange.
// k_vox_emit    one thread per creator = per voxel: it gathers the voxel's runs (its own and the chain's), sorts the run heads
//               by point index when there are at most eight of them, and adds the points one by one in float64 -- for those
//               voxels exactly the index-order sums of Open3D's loop, so the means are bit-identical to the oracle's.  Points
//               are re-read from the input (L2-resident: the insert just streamed it) and, for a fusion, re-transformed.
//               mean = sum / count by the exact-quotient helper.  Part p's voxels go behind those of parts 0..p-1.
// A voxel made of more than kVoxHeads runs (a coarse grid, or a surface seen at close range) is finished by k_vox_long instead:
// a thread walking a long chain would serialise the gather, so every joiner of such a voxel adds its run into a pool record
// with float64 atomics, in parallel, and k_vox_long_final divides.  Both return at once when no voxel is long.
//
// The pose transform of a fusion (K3) is fused in: the views are read where they lie, p' = T p is rounded to the cloud's
// storage type exactly as rv_transform_merge stores it, and the merged cloud is never written.
constexpr int kVoxViews = 8;
struct VoxView {
  const void *in;
  long long stride, n, first;  // first: index of the view's first point in the merged order
  double T[16];
};
struct VoxHeader {
  double bounds[6];
  int error;  // 1: a voxel index does not fit 21 bits / extent check failed
  unsigned int n_long;
};
constexpr int kVoxMaxParts = 2048;
constexpr size_t kVoxHead = 256 + (size_t)kVoxMaxParts * 8;  // header + per-part {creators, joiners}
constexpr unsigned int kVoxNil = 0xffffffffu;
constexpr unsigned int kVoxLongTag = 0x80000000u;
constexpr int kVoxHeads = 8;  // run heads of a voxel that k_vox_emit gathers itself (sorted, summed in index order)
// Both kernels are chains of dependent L2 round trips, so resident warps are what hides them: measured on the four-view
// 5 mm fusion cloud (tools/k4_sweep.sh), one group per warp step at six CTAs per SM and 256-thread emit CTAs at eight per SM
// took 134 us where four groups (69 registers, three CTAs per SM) and 128-thread emit CTAs took 193 us.
#ifndef RV_VOX_GROUPS
#define RV_VOX_GROUPS 1
#endif
#ifndef RV_VOX_INSERT_OCC
#define RV_VOX_INSERT_OCC 6
#endif
#ifndef RV_VOX_EMIT_THREADS
#define RV_VOX_EMIT_THREADS 256
#endif
#ifndef RV_VOX_EMIT_OCC
#define RV_VOX_EMIT_OCC 8
#endif
struct __align__(64) VoxLong {
  unsigned long long key;
  unsigned int count;
  unsigned int out;  // output position of the voxel
  double sum[6];
};
static_assert(sizeof(VoxLong) == 64, "pool record layout");

struct __align__(16) VoxSlot {
  unsigned long long key;  // 0 = empty, else packed(ix,iy,iz) + 1
  unsigned int chain;      // 0 = no joiner yet; list position + 1 of the most recent joiner; kVoxLongTag | pool index once long
  unsigned int pad;
};

struct VoxArgs {
  VoxView view[kVoxViews];
  int n_views, identity, has_color;
  long long n;
  double voxel, rvoxel;
  const double *bounds;  // device
  VoxHeader *hdr;
  VoxSlot *slots;            // the table: key and chain head of a voxel in one 16-byte record (one sector for both)
  uint2 *list;               // part p owns list[p * span, ...): creators (point index, slot) from the front, joiners
                             // (point index, list position of the joiner linked before it or kVoxNil) from the back
  unsigned int *jslot;       // per list position of a joiner: its slot (read by k_vox_long only)
  unsigned int *longmap;     // one bit per slot: the voxel went to the pool (k_vox_long's joiners look here first: 0.3 MB
                             // that stays in the L2, where the table is 41 MB of random reads)
  unsigned int *headmask;    // per 32-point group: bit l set = point l starts a run
  VoxLong *pool;
  void *mxyz;    // fusion only: the transformed coordinates (three planes of n, the cloud's element type), written once by the
                 // bounds pass so that insert and emit read plain values instead of transforming every point again
  uint2 *parts;  // per part {creators, joiners}
  long long span;  // points per part, a multiple of 128
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

// view of merged point i (views are few: a short scan)
__device__ __forceinline__ int vox_view_of(const VoxArgs &a, long long i) {
  int v = 0;
#pragma unroll 1
  for (int q = 1; q < a.n_views; ++q)
    if (i >= a.view[q].first) v = q;
  return v;
}

// coordinates of merged point i as they are STORED in the cloud's element type T (after the pose transform of its view)
template <typename T>
__device__ __forceinline__ void vox_xyz_src(const VoxArgs &a, long long i, double &x, double &y, double &z);

// kFused is a compile-time flag: the emit kernel holds eight copies of this read and lost a quarter of its speed (62 -> 84
// us) when the choice was a run-time branch
template <typename T, bool kFused>
__device__ __forceinline__ void vox_xyz(const VoxArgs &a, long long i, double &x, double &y, double &z) {
  if (kFused) {  // fusion: transformed once by the bounds pass
    const T *m = reinterpret_cast<const T *>(a.mxyz);
    x = (double)m[i];
    y = (double)m[a.n + i];
    z = (double)m[2 * a.n + i];
    return;
  }
  vox_xyz_src<T>(a, i, x, y, z);
}

template <typename T>
__device__ __forceinline__ void vox_xyz_src(const VoxArgs &a, long long i, double &x, double &y, double &z) {
  const int v = a.n_views > 1 ? vox_view_of(a, i) : 0;
  const VoxView &w = a.view[v];
  const T *in = reinterpret_cast<const T *>(w.in);
  const long long k = i - w.first;
  x = (double)in[k];
  y = (double)in[w.stride + k];
  z = (double)in[2 * w.stride + k];
  if (!a.identity) {  // Open3D Transform, one rounding per operation (k_transform), then the store's rounding
    const double *M = w.T;
    const double qx = ((M[0] * x + M[1] * y) + M[2] * z) + M[3];
    const double qy = ((M[4] * x + M[5] * y) + M[6] * z) + M[7];
    const double qz = ((M[8] * x + M[9] * y) + M[10] * z) + M[11];
    const double qw = ((M[12] * x + M[13] * y) + M[14] * z) + M[15];
    double px = qx, py = qy, pz = qz;
    if (qw != 1.0) {
      px = qx / qw;
      py = qy / qw;
      pz = qz / qw;
    }
    x = (double)(T)px;
    y = (double)(T)py;
    z = (double)(T)pz;
  }
}

template <typename T>
__device__ __forceinline__ void vox_rgb(const VoxArgs &a, long long i, double &r, double &g, double &b) {
  const int v = a.n_views > 1 ? vox_view_of(a, i) : 0;
  const VoxView &w = a.view[v];
  const T *in = reinterpret_cast<const T *>(w.in);
  const long long k = i - w.first;
  r = (double)in[3 * w.stride + k];
  g = (double)in[4 * w.stride + k];
  b = (double)in[5 * w.stride + k];
}

// bounds of the (transformed, stored) cloud: the grid's origin is min_bound - voxel / 2
template <typename T>
__global__ void __launch_bounds__(256) k_vox_bounds(const VoxArgs a, double *bounds) {
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
  const long long stride = (long long)gridDim.x * blockDim.x;
  T *m = reinterpret_cast<T *>(a.mxyz);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    double p[3];
    vox_xyz_src<T>(a, i, p[0], p[1], p[2]);
    if (m) {
      m[i] = (T)p[0];
      m[a.n + i] = (T)p[1];
      m[2 * a.n + i] = (T)p[2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      lo[c] = p[c] < lo[c] ? p[c] : lo[c];
      hi[c] = p[c] > hi[c] ? p[c] : hi[c];
    }
  }
  block_bounds_commit(lo, hi, bounds);
}

constexpr int kVoxGroups = RV_VOX_GROUPS;  // 32-point groups per warp and step

template <typename T, bool kFused>
__global__ void __launch_bounds__(256, RV_VOX_INSERT_OCC) k_vox_insert(const VoxArgs a) {
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
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = rv_lanemask_lt();
  __shared__ unsigned int s_create, s_join;
  if (threadIdx.x == 0) s_create = s_join = 0;
  __syncthreads();
  const long long p0 = (long long)blockIdx.x * a.span;         // this CTA's part of the cloud
  const long long p1 = p0 + a.span < a.n ? p0 + a.span : a.n;  // (empty when p0 >= n)
  for (long long w0 = p0 + warp * (32 * kVoxGroups); w0 < p1; w0 += 8 * 32 * kVoxGroups) {
    // ---- every load of the step first, then the keys
    double px[kVoxGroups], py[kVoxGroups], pz[kVoxGroups];
    bool valid[kVoxGroups];
#pragma unroll
    for (int g = 0; g < kVoxGroups; ++g) {
      const long long i = w0 + 32 * g + lane;
      valid[g] = i < p1;
      px[g] = py[g] = pz[g] = 0.0;
      if (valid[g]) vox_xyz<T, kFused>(a, i, px[g], py[g], pz[g]);
    }
    unsigned long long key[kVoxGroups];
    uint32_t heads[kVoxGroups];
    bool lead[kVoxGroups];
#pragma unroll
    for (int g = 0; g < kVoxGroups; ++g) {
      key[g] = 0;  // lanes past the end form one run that is never written
      if (valid[g]) {
        // key = floor((p - origin) / voxel), IEEE division via the exact-reciprocal helper
        const long long ix = (long long)floor(rv_div(px[g] - ox, a.voxel, a.rvoxel));
        const long long iy = (long long)floor(rv_div(py[g] - oy, a.voxel, a.rvoxel));
        const long long iz = (long long)floor(rv_div(pz[g] - oz, a.voxel, a.rvoxel));
        key[g] = (((unsigned long long)ix & 0x1fffff) << 42 | ((unsigned long long)iy & 0x1fffff) << 21 | ((unsigned long long)iz & 0x1fffff)) + 1ull;
      }
      // runs of equal keys inside the group: only the first point of a run goes to the table
      const unsigned long long prev = __shfl_up_sync(0xffffffffu, key[g], 1);
      const bool head = lane == 0 || key[g] != prev;
      heads[g] = __ballot_sync(0xffffffffu, head);
      lead[g] = head && valid[g];
      if (lane == 0 && valid[g]) a.headmask[(w0 >> 5) + g] = heads[g];
    }
    // ---- one slot probe per run: the first compare-and-swap of all groups is in flight before any result is looked at
    unsigned int h[kVoxGroups];
    unsigned long long cur[kVoxGroups];
#pragma unroll
    for (int g = 0; g < kVoxGroups; ++g) {
      h[g] = 0;
      cur[g] = 0;
      if (lead[g]) {
        h[g] = __umulhi((unsigned int)(vox_hash(key[g]) >> 32), a.cap);  // uniform over [0, cap)
        cur[g] = atomicCAS(&a.slots[h[g]].key, 0ull, key[g]);             // the table is at most two thirds full: usually empty
      }
    }
#pragma unroll
    for (int g = 0; g < kVoxGroups; ++g) {
      bool created = false;
      if (lead[g]) {
        while (cur[g] != 0ull && cur[g] != key[g]) {  // another voxel's slot: linear probing
          if (++h[g] == a.cap) h[g] = 0;
          cur[g] = atomicCAS(&a.slots[h[g]].key, 0ull, key[g]);
        }
        created = cur[g] == 0ull;
      }
      // ---- creators to the front of the part's stretch of the list, joiners to its back
      const uint32_t cb = __ballot_sync(0xffffffffu, created);
      const uint32_t jb = __ballot_sync(0xffffffffu, lead[g] && !created);
      unsigned int cbase = 0, jbase = 0;
      if (lane == 0) {
        if (cb) cbase = atomicAdd(&s_create, (unsigned int)__popc(cb));
        if (jb) jbase = atomicAdd(&s_join, (unsigned int)__popc(jb));
      }
      cbase = __shfl_sync(0xffffffffu, cbase, 0);
      jbase = __shfl_sync(0xffffffffu, jbase, 0);
      const unsigned int i32 = (unsigned int)(w0 + 32 * g + lane);
      if (created) {
        a.list[p0 + cbase + __popc(cb & lt)] = make_uint2(i32, h[g]);
      } else if (lead[g]) {
        const unsigned int pos = (unsigned int)(p1 - 1 - (long long)(jbase + __popc(jb & lt)));
        const unsigned int before = atomicExch(&a.slots[h[g]].chain, pos + 1u);  // link behind whoever joined this voxel before
        a.list[pos] = make_uint2(i32, before ? before - 1u : kVoxNil);
        a.jslot[pos] = h[g];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) a.parts[blockIdx.x] = make_uint2(s_create, s_join);
}

struct VoxOutArgs {
  void *out;
  long long out_stride, out_capacity;
  int32_t *keys;
  int32_t *counts;
  long long *m;
};

// length of the run that starts at point i: up to the next head of its 32-point group
__device__ __forceinline__ int vox_run(const VoxArgs &a, unsigned int i) {
  const uint32_t m = a.headmask[i >> 5];
  const int l = (int)(i & 31u);
  const uint32_t after = l == 31 ? 0u : (m >> (l + 1));
  return after ? __ffs(after) : 32 - l;
}

struct VoxSum {
  double s[6];
  unsigned int count;
};

// add the points of the run starting at i, one by one (index order inside the run)
// (kept inline on purpose: as a called function -- eight calls per voxel instead of eight copies of the loop -- the emit
// kernel ran 2.6 x slower, 351 vs 134 us on the 5 mm fusion cloud)
template <typename T, bool kFused>
__device__ __forceinline__ void vox_add_run(const VoxArgs &a, unsigned int i, VoxSum &acc) {
  const int run = vox_run(a, i);
  for (int r = 0; r < run; ++r) {
    double x, y, z;
    vox_xyz<T, kFused>(a, (long long)i + r, x, y, z);
    acc.s[0] += x;
    acc.s[1] += y;
    acc.s[2] += z;
    if (a.has_color) {
      double cr, cg, cb;
      vox_rgb<T>(a, (long long)i + r, cr, cg, cb);
      acc.s[3] += cr;
      acc.s[4] += cg;
      acc.s[5] += cb;
    }
  }
  acc.count += (unsigned int)run;
}

template <typename OutT>
__device__ __forceinline__ void vox_store(const VoxOutArgs &o, int has_color, long long pos, const double s[6], unsigned int count,
                                          unsigned long long key) {
  if (pos >= o.out_capacity) return;
  OutT *out = reinterpret_cast<OutT *>(o.out);
  const double c = (double)count, rc = 1.0 / c;
  out[pos] = (OutT)rv_div(s[0], c, rc);
  out[o.out_stride + pos] = (OutT)rv_div(s[1], c, rc);
  out[2 * o.out_stride + pos] = (OutT)rv_div(s[2], c, rc);
  if (has_color) {
    out[3 * o.out_stride + pos] = (OutT)rv_div(s[3], c, rc);
    out[4 * o.out_stride + pos] = (OutT)rv_div(s[4], c, rc);
    out[5 * o.out_stride + pos] = (OutT)rv_div(s[5], c, rc);
  }
  if (o.keys) {
    const unsigned long long kk = key - 1ull;
    o.keys[pos] = (int32_t)((kk >> 42) & 0x1fffff);
    o.keys[o.out_capacity + pos] = (int32_t)((kk >> 21) & 0x1fffff);
    o.keys[2 * o.out_capacity + pos] = (int32_t)(kk & 0x1fffff);
  }
  if (o.counts) o.counts[pos] = (int32_t)count;
}

// CTA p finishes the voxels part p created, behind those of parts 0..p-1
constexpr int kVoxEmitThreads = RV_VOX_EMIT_THREADS;

template <typename T, typename OutT, bool kFused>
__global__ void __launch_bounds__(kVoxEmitThreads, RV_VOX_EMIT_OCC) k_vox_emit(const VoxArgs a, const VoxOutArgs o) {
  __shared__ unsigned long long s_red[2][kVoxEmitThreads / 32];
  unsigned long long before = 0, total = 0;
  for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
    const unsigned long long c = a.parts[q].x;
    total += c;
    if (q < (int)blockIdx.x) before += c;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    before += __shfl_xor_sync(0xffffffffu, before, s);
    total += __shfl_xor_sync(0xffffffffu, total, s);
  }
  if ((threadIdx.x & 31) == 0) s_red[0][threadIdx.x >> 5] = before, s_red[1][threadIdx.x >> 5] = total;
  __syncthreads();
  before = total = 0;
#pragma unroll
  for (int w = 0; w < kVoxEmitThreads / 32; ++w) before += s_red[0][w], total += s_red[1][w];
  if (blockIdx.x == 0 && threadIdx.x == 0) *o.m = a.hdr->error ? -1ll : (long long)total;
  if (a.hdr->error) return;
  const unsigned int mine = a.parts[blockIdx.x].x;
  const uint2 *list = a.list + (long long)blockIdx.x * a.span;
  for (unsigned int k = threadIdx.x; k < mine; k += blockDim.x) {
    const long long pos = (long long)(before + k);
    const uint2 me = list[k];  // (first point, slot)
    // ---- the voxel's run heads: its creator and up to seven joiners in registers, sorted by point index so that the
    // points are added in index order (the order of Open3D's loop); whatever a longer chain holds follows in chain order
    unsigned int idx[8];
    idx[0] = me.x;
    const uint4 slot = *reinterpret_cast<const uint4 *>(a.slots + me.y);  // key and chain head in one load
    const unsigned long long key = ((unsigned long long)slot.y << 32) | slot.x;
    unsigned int link = slot.z ? slot.z - 1u : kVoxNil;
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      idx[q] = kVoxNil;
      if (link != kVoxNil) {
        const uint2 e = a.list[link];  // (point, the joiner before it)
        idx[q] = e.x;
        link = e.y;
      }
    }
#define RV_CX(i, j)                                     \
  {                                                     \
    const unsigned int lo_ = min(idx[i], idx[j]);       \
    const unsigned int hi_ = max(idx[i], idx[j]);       \
    idx[i] = lo_;                                       \
    idx[j] = hi_;                                       \
  }
    if (idx[1] != kVoxNil && link == kVoxNil) {  // 19-comparator network for eight keys (empty entries sort last)
      RV_CX(0, 1) RV_CX(2, 3) RV_CX(4, 5) RV_CX(6, 7)
      RV_CX(0, 2) RV_CX(1, 3) RV_CX(4, 6) RV_CX(5, 7)
      RV_CX(1, 2) RV_CX(5, 6) RV_CX(0, 4) RV_CX(3, 7)
      RV_CX(1, 5) RV_CX(2, 6)
      RV_CX(1, 4) RV_CX(3, 6)
      RV_CX(2, 4) RV_CX(3, 5)
      RV_CX(3, 4)
    }
#undef RV_CX
    // pool slots for the long voxels of this warp step: one atomic per warp (a coarse grid makes every voxel long, and
    // same-address atomics serialise)
    const bool is_long = link != kVoxNil;
    const unsigned int active = __activemask();
    const unsigned int lm = __ballot_sync(active, is_long);
    unsigned int qbase = 0;
    if (lm) {
      const int leader = __ffs(lm) - 1;
      if ((int)(threadIdx.x & 31) == leader) qbase = atomicAdd(&a.hdr->n_long, (unsigned int)__popc(lm));
      qbase = __shfl_sync(active, qbase, leader);
    }
    if (is_long) {
      // more than eight runs: hand the voxel to the atomic path (its joiners add themselves in parallel in k_vox_long).
      // (Letting the thread gather runs 9..16 itself, so that a fine grid has no long voxel and k_vox_long returns at once
      // instead of walking the joiners for 10 us, made this kernel 29 us slower: 62 -> 91 us.  So did gathering all points
      // of a voxel of up to four points in one go -- every load in flight together, added in index order afterwards --:
      // the 48-80 registers that takes cost more resident warps than the shorter chain gives back, 146-163 vs 132 us.)
      const unsigned int q = qbase + __popc(lm & rv_lanemask_lt());
      VoxLong *rec = a.pool + q;
      rec->key = key;
      rec->out = (unsigned int)pos;
      VoxSum own;
#pragma unroll
      for (int c = 0; c < 6; ++c) own.s[c] = 0.0;
      own.count = 0;
      vox_add_run<T, kFused>(a, me.x, own);
      rec->count = own.count;
#pragma unroll
      for (int c = 0; c < 6; ++c) rec->sum[c] = own.s[c];
      a.slots[me.y].chain = kVoxLongTag | q;
      atomicOr(a.longmap + (me.y >> 5), 1u << (me.y & 31u));
      continue;
    }
    VoxSum acc;
#pragma unroll
    for (int c = 0; c < 6; ++c) acc.s[c] = 0.0;
    acc.count = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (idx[q] != kVoxNil) vox_add_run<T, kFused>(a, idx[q], acc);
    vox_store<OutT>(o, a.has_color, pos, acc.s, acc.count, key);
  }
}

// every joiner of a long voxel adds its run into the voxel's pool record (float64 atomics); nothing to do without long voxels
template <typename T, bool kFused>
__global__ void __launch_bounds__(256) k_vox_long(const VoxArgs a, int n_parts) {
  if (a.hdr->error || a.hdr->n_long == 0) return;
  for (int p = blockIdx.x; p < n_parts; p += gridDim.x) {
    const long long p0 = (long long)p * a.span;
    const long long p1 = p0 + a.span < a.n ? p0 + a.span : a.n;
    const unsigned int joiners = a.parts[p].y;
    for (unsigned int j = threadIdx.x; j < joiners; j += blockDim.x) {
      const long long at = p1 - 1 - (long long)j;
      const unsigned int slot = a.jslot[at];
      if (!((a.longmap[slot >> 5] >> (slot & 31u)) & 1u)) continue;
      const uint2 e = a.list[at];
      const unsigned int v = a.slots[slot].chain;
      VoxSum s;
#pragma unroll
      for (int c = 0; c < 6; ++c) s.s[c] = 0.0;
      s.count = 0;
      vox_add_run<T, kFused>(a, e.x, s);
      VoxLong *rec = a.pool + (v & ~kVoxLongTag);
      atomicAdd(&rec->count, s.count);
      atomicAdd(&rec->sum[0], s.s[0]);
      atomicAdd(&rec->sum[1], s.s[1]);
      atomicAdd(&rec->sum[2], s.s[2]);
      if (a.has_color) {
        atomicAdd(&rec->sum[3], s.s[3]);
        atomicAdd(&rec->sum[4], s.s[4]);
        atomicAdd(&rec->sum[5], s.s[5]);
      }
    }
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256) k_vox_long_final(const VoxArgs a, const VoxOutArgs o) {
  if (a.hdr->error) return;
  const unsigned int n_long = a.hdr->n_long;
  for (unsigned int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_long; q += gridDim.x * blockDim.x) {
    const VoxLong rec = a.pool[q];
    vox_store<OutT>(o, a.has_color, (long long)rec.out, rec.sum, rec.count, rec.key);
  }
}

unsigned long long vox_capacity(long long n) {  // 1.5 slots per point, a multiple of 32
  unsigned long long c = ((unsigned long long)n * 3ull / 2ull + 31ull) & ~31ull;
  return c < 1024 ? 1024 : c;
}

size_t vox_align(size_t b) { return (b + 255) & ~(size_t)255; }

struct VoxLayout {
  size_t longmap, slots, list, jslot, masks, pool, mxyz, total;
};
VoxLayout vox_layout(long long n) {
  VoxLayout L;
  const unsigned long long cap = vox_capacity(n);
  size_t off = kVoxHead;
  L.longmap = off;
  off += vox_align((size_t)(cap / 32) * 4);  // cap is a multiple of 32
  L.slots = off;
  off += vox_align((size_t)cap * sizeof(VoxSlot));
  L.list = off;
  off += vox_align((size_t)n * 8);
  L.jslot = off;
  off += vox_align((size_t)n * 4);
  L.masks = off;
  off += vox_align((size_t)((n + 31) / 32) * 4);
  L.pool = off;
  off += vox_align((size_t)(n / (kVoxHeads + 1) + 1) * sizeof(VoxLong));  // a long voxel has more than kVoxHeads run heads
  L.mxyz = off;
  off += vox_align((size_t)n * 24);  // a fusion's transformed coordinates (float64 at most); unused by rv_voxel_downsample
  L.total = off;
  return L;
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
  return vox_layout(n).total;
}

// shared body of rv_voxel_downsample (one view, no transform) and rv_fuse_voxel (K3 fused in)
static int vox_run_all(rv_ctx *ctx, const char *who, int n_views, const void *const *d_in, const int64_t *in_plane_stride,
                       const int64_t *n, const double *T, int in_dtype, int has_color, double voxel_size, const double *d_bounds,
                       void *d_out, int64_t out_plane_stride, int out_dtype, int64_t out_capacity, int32_t *d_keys,
                       int32_t *d_counts_out, int64_t *d_m, void *d_ws, size_t ws_bytes, cudaStream_t st) {
  if (!(voxel_size > 0.0)) RV_FAIL(ctx, RV_EINVAL, "%s: voxel_size <= 0", who);
  if (n_views < 1 || n_views > kVoxViews || !d_in || !in_plane_stride || !n || !d_m)
    RV_FAIL(ctx, RV_EINVAL, "%s: 1..%d views and non-null arrays are required", who, kVoxViews);
  if ((in_dtype != RV_F32 && in_dtype != RV_F64) || (out_dtype != RV_F32 && out_dtype != RV_F64)) RV_FAIL(ctx, RV_EINVAL, "%s: bad dtype", who);
  if (out_capacity < 0 || out_plane_stride < out_capacity) RV_FAIL(ctx, RV_EINVAL, "%s: bad output capacity", who);
  VoxArgs a;
  memset(&a, 0, sizeof(a));
  long long total = 0;
  int nv = 0;
  for (int v = 0; v < n_views; ++v) {
    if (n[v] < 0 || in_plane_stride[v] < n[v]) RV_FAIL(ctx, RV_EINVAL, "%s: bad n / stride for view %d", who, v);
    if (n[v] == 0) continue;
    if (!d_in[v]) RV_FAIL(ctx, RV_EINVAL, "%s: view %d is null", who, v);
    VoxView &w = a.view[nv++];
    w.in = d_in[v];
    w.stride = in_plane_stride[v];
    w.n = n[v];
    w.first = total;
    if (T) memcpy(w.T, T + 16 * v, sizeof(w.T));
    total += n[v];
  }
  if (total >= 0x7fffffffll) RV_FAIL(ctx, RV_EINVAL, "%s: more than 2^31 points", who);  // list positions and the long tag share 32 bits
  if (total == 0) {
    RV_CUDA(ctx, cudaMemsetAsync(d_m, 0, sizeof(int64_t), st));
    return RV_OK;
  }
  if (out_capacity > 0 && !d_out) RV_FAIL(ctx, RV_EINVAL, "%s: null output", who);
  const VoxLayout L = vox_layout(total);
  if (!d_ws || ws_bytes < L.total) RV_FAIL(ctx, RV_EWORKSPACE, "%s: workspace %zu < %zu", who, ws_bytes, L.total);
  if (!rv_aligned(d_ws, 64)) RV_FAIL(ctx, RV_EALIGN, "%s: workspace must be 64-byte aligned", who);
  const unsigned long long cap = vox_capacity(total);
  char *w = reinterpret_cast<char *>(d_ws);
  a.n_views = nv;
  a.identity = T ? 0 : 1;
  a.has_color = has_color ? 1 : 0;
  a.n = total;
  a.voxel = voxel_size;
  a.rvoxel = 1.0 / voxel_size;
  a.hdr = reinterpret_cast<VoxHeader *>(w);
  a.parts = reinterpret_cast<uint2 *>(w + 256);
  a.longmap = reinterpret_cast<unsigned int *>(w + L.longmap);
  a.slots = reinterpret_cast<VoxSlot *>(w + L.slots);
  a.list = reinterpret_cast<uint2 *>(w + L.list);
  a.jslot = reinterpret_cast<unsigned int *>(w + L.jslot);
  a.headmask = reinterpret_cast<unsigned int *>(w + L.masks);
  a.pool = reinterpret_cast<VoxLong *>(w + L.pool);
  a.mxyz = a.identity ? nullptr : (void *)(w + L.mxyz);
  a.cap = (unsigned int)cap;
  // header, part counters, long-voxel bits and the table (empty keys, no joiners) are cleared; nothing else needs initialising
  RV_CUDA(ctx, cudaMemsetAsync(d_ws, 0, L.slots + (size_t)cap * sizeof(VoxSlot), st));
  const double *bounds = d_bounds;
  if (!bounds) {
    k_bounds_init<<<1, 32, 0, st>>>(a.hdr->bounds);
    RV_LAUNCHED(ctx);
    const int g = grid_for(ctx, total);
    if (in_dtype == RV_F32) k_vox_bounds<float><<<g, 256, 0, st>>>(a, a.hdr->bounds);  // a fusion also stores p' = T p here
    else k_vox_bounds<double><<<g, 256, 0, st>>>(a, a.hdr->bounds);
    RV_LAUNCHED(ctx);
    bounds = a.hdr->bounds;
  }
  a.bounds = bounds;
  // parts of at least 1024 points, up to 2048 of them: two to three waves of small CTAs keep the tail of the kernel short
  int parts_n = (int)((total + 1023) / 1024 < kVoxMaxParts ? (total + 1023) / 1024 : kVoxMaxParts);
  if (parts_n < 1) parts_n = 1;
  a.span = (((total + parts_n - 1) / parts_n) + 127) & ~127ll;
  const bool fused = a.mxyz != nullptr;
  if (in_dtype == RV_F32) {
    if (fused) k_vox_insert<float, true><<<parts_n, 256, 0, st>>>(a);
    else k_vox_insert<float, false><<<parts_n, 256, 0, st>>>(a);
  } else {
    if (fused) k_vox_insert<double, true><<<parts_n, 256, 0, st>>>(a);
    else k_vox_insert<double, false><<<parts_n, 256, 0, st>>>(a);
  }
  RV_LAUNCHED(ctx);
  VoxOutArgs o;
  memset(&o, 0, sizeof(o));
  o.out = d_out;
  o.out_stride = out_plane_stride;
  o.out_capacity = out_capacity;
  o.keys = d_keys;
  o.counts = d_counts_out;
  o.m = reinterpret_cast<long long *>(d_m);
#define RV_EMIT(TI, TO)                                                                        \
  {                                                                                            \
    if (fused) k_vox_emit<TI, TO, true><<<parts_n, kVoxEmitThreads, 0, st>>>(a, o);            \
    else k_vox_emit<TI, TO, false><<<parts_n, kVoxEmitThreads, 0, st>>>(a, o);                 \
  }
  if (in_dtype == RV_F32 && out_dtype == RV_F32) RV_EMIT(float, float)
  else if (in_dtype == RV_F32) RV_EMIT(float, double)
  else if (out_dtype == RV_F32) RV_EMIT(double, float)
  else RV_EMIT(double, double)
#undef RV_EMIT
  RV_LAUNCHED(ctx);
  // voxels with very long chains (none on a fine grid: both kernels then return at once)
  const int lg = parts_n < ctx->sm_count * 4 ? parts_n : ctx->sm_count * 4;
  if (in_dtype == RV_F32) {
    if (fused) k_vox_long<float, true><<<lg, 256, 0, st>>>(a, parts_n);
    else k_vox_long<float, false><<<lg, 256, 0, st>>>(a, parts_n);
  } else {
    if (fused) k_vox_long<double, true><<<lg, 256, 0, st>>>(a, parts_n);
    else k_vox_long<double, false><<<lg, 256, 0, st>>>(a, parts_n);
  }
  RV_LAUNCHED(ctx);
  if (out_dtype == RV_F32) k_vox_long_final<float><<<8, 256, 0, st>>>(a, o);
  else k_vox_long_final<double><<<8, 256, 0, st>>>(a, o);
  RV_LAUNCHED(ctx);
  return RV_OK;
}

int rv_voxel_downsample(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int in_dtype, int has_color,
                        double voxel_size, const double *d_bounds, void *d_out, int64_t out_plane_stride, int out_dtype,
                        int64_t out_capacity, int32_t *d_keys, int32_t *d_counts_out, int64_t *d_m, void *d_ws,
                        size_t ws_bytes, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (n < 0 || in_plane_stride < n || !d_m) RV_FAIL(ctx, RV_EINVAL, "rv_voxel_downsample: bad n / stride / m");
  if (n > 0 && !d_in) RV_FAIL(ctx, RV_EINVAL, "rv_voxel_downsample: null cloud pointer");
  const void *views[1] = {d_in};
  const int64_t strides[1] = {in_plane_stride}, ns[1] = {n};
  return vox_run_all(ctx, "rv_voxel_downsample", 1, views, strides, ns, nullptr, in_dtype, has_color, voxel_size, d_bounds, d_out,
                     out_plane_stride, out_dtype, out_capacity, d_keys, d_counts_out, d_m, d_ws, ws_bytes, (cudaStream_t)stream);
}

int rv_fuse_voxel(rv_ctx *ctx, int n_views, const void *const *d_in, const int64_t *in_plane_stride, const int64_t *n,
                  const double *T, int in_dtype, int has_color, double voxel_size, void *d_out, int64_t out_plane_stride,
                  int out_dtype, int64_t out_capacity, int32_t *d_keys, int32_t *d_counts_out, int64_t *d_m, void *d_ws,
                  size_t ws_bytes, rv_stream stream) {
  if (!ctx) return RV_EINVAL;
  RvDeviceGuard dev_guard(ctx);
  if (!T) RV_FAIL(ctx, RV_EINVAL, "rv_fuse_voxel: the poses are required (rv_voxel_downsample takes a cloud as it is)");
  return vox_run_all(ctx, "rv_fuse_voxel", n_views, d_in, in_plane_stride, n, T, in_dtype, has_color, voxel_size, nullptr, d_out,
                     out_plane_stride, out_dtype, out_capacity, d_keys, d_counts_out, d_m, d_ws, ws_bytes, (cudaStream_t)stream);
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
