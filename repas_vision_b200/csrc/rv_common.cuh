// rv_common.cuh -- shared device/host helpers for the sm_100a kernels.
//
// Numerics policy (DESIGN.md "Numerics"): the whole library is compiled with
// --fmad=false, so `a*b+c` rounds twice exactly like numpy / plain C; fused
// multiply-adds appear only where written explicitly (fma(), fmaf()) inside the
// exact-division helpers below.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/repas_vision.h"

struct rv_ctx {
  int device;
  int sm_count;
  int64_t launches;
  char err[512];
};

#define RV_FAIL(ctx, code, ...)                               \
  do {                                                        \
    if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
    return (code);                                            \
  } while (0)

#define RV_CUDA(ctx, expr)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) RV_FAIL(ctx, RV_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

#define RV_LAUNCHED(ctx)                                                                        \
  do {                                                                                          \
    (ctx)->launches++;                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) RV_FAIL(ctx, RV_ECUDA, "kernel launch: %s", cudaGetErrorString(e__)); \
  } while (0)

// Entry points launch on the context's device whatever the caller's current device is, and restore it.
struct RvDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit RvDeviceGuard(const rv_ctx *ctx) {
    if (ctx && cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device) switched = cudaSetDevice(ctx->device) == cudaSuccess;
  }
  ~RvDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

static inline bool rv_aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// ---------------------------------------------------------------- exact division
// Correctly rounded a/b from a precomputed correctly rounded reciprocal rb = RN(1/b)
// (computed on the host): q0 = a*rb, then two residual corrections with fused
// multiply-adds (Markstein / Cornea et al.).  5 ops instead of the ~30-instruction
// IEEE division subroutine; bit-identical to `a / b` for the finite, normal operands
// this path sees (pixel offsets x depths over focal lengths, bytes over 255, raw depth
// over 1000, coordinates over the voxel size).  tests/test_gpu_cloud.py::test_exact_division_helper_against_numpy
// checks it against numpy (all 65 536 raw depths x several focal lengths, all 256 bytes).
__device__ __forceinline__ double rv_div(double a, double b, double rb) {
  double q = a * rb;
  double r = fma(-q, b, a);
  q = fma(r, rb, q);
  r = fma(-q, b, a);
  return fma(r, rb, q);
}
__device__ __forceinline__ float rv_divf(float a, float b, float rb) {
  float q = a * rb;
  float r = fmaf(-q, b, a);
  q = fmaf(r, rb, q);
  r = fmaf(-q, b, a);
  return fmaf(r, rb, q);
}

// ------------------------------------------------------------------- memory ops
__device__ __forceinline__ void rv_st_relaxed(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long rv_ld_relaxed(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ uint32_t rv_lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ uint32_t rv_warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------ decoupled look-back (one chain per frame)
// status word: bits 63..62 = flag (0 not ready, 1 tile aggregate, 2 inclusive prefix), low 32 bits = value.
#define RV_ST_AGG (1ull << 62)
#define RV_ST_PREFIX (2ull << 62)

// Called by every lane of ONE warp.  `tile` indexes status[]; the tile has `n_pred`
// predecessors in its chain (tile-1 ... tile-n_pred).  Publishes this tile's aggregate,
// walks back over the predecessors 32 at a time and returns the exclusive prefix.
__device__ __forceinline__ uint32_t rv_lookback(unsigned long long *status, int tile, int n_pred, uint32_t aggregate) {
  const int lane = threadIdx.x & 31;
  if (n_pred == 0) {
    if (lane == 0) rv_st_relaxed(status + tile, RV_ST_PREFIX | aggregate);
    return 0;
  }
  if (lane == 0) rv_st_relaxed(status + tile, RV_ST_AGG | aggregate);
  uint32_t excl = 0;
  int look = tile - 1;
  int remaining = n_pred;
  for (;;) {
    unsigned long long s = RV_ST_PREFIX;  // lanes past the chain head read as "prefix 0"
    if (lane < remaining) {
      s = rv_ld_relaxed(status + (look - lane));
      while ((s >> 62) == 0) {  // predecessor still computing in another CTA: poll gently, the issue slots are needed
        __nanosleep(256);
        s = rv_ld_relaxed(status + (look - lane));
      }
    }
    const uint32_t pmask = __ballot_sync(0xffffffffu, (s >> 62) == 2);
    const int first = pmask ? (__ffs(pmask) - 1) : 32;
    excl += rv_warp_sum(lane <= first ? (uint32_t)s : 0u);
    if (pmask) break;
    look -= 32;
    remaining -= 32;
  }
  if (lane == 0) rv_st_relaxed(status + tile, RV_ST_PREFIX | (unsigned long long)(excl + aggregate));
  return excl;
}

// ------------------------------------------------------------ double atomic min / max
__device__ __forceinline__ void rv_atomic_min_f64(double *addr, double v) {
  unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
  unsigned long long old = *a;
  while (v < __longlong_as_double((long long)old)) {
    unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void rv_atomic_max_f64(double *addr, double v) {
  unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
  unsigned long long old = *a;
  while (v > __longlong_as_double((long long)old)) {
    unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

// host-side: grid for a persistent kernel = min(work items, SMs x resident CTAs)
template <typename K>
static inline int rv_persistent_grid(const rv_ctx *ctx, K kernel, int block, size_t smem, long long work_items) {
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem) != cudaSuccess || occ < 1) occ = 1;
  long long g = (long long)ctx->sm_count * occ;
  if (work_items < g) g = work_items;
  if (g < 1) g = 1;
  return (int)g;
}
