// rv_deproject_tma.cu -- K1 fast path: the fused deprojection + masks + ordered compaction kernel as a
// TMA-fed persistent pipeline (sm_100a).
//
// Same contract and arithmetic as k_deproject in rv_deproject.cu (reference semantics cited there:
// create_masked_ply.py:56-107, better_three_capture.py:118-125, distance_masking_on_ply.py:12-19,
// view_point_cloud.py:109-116, april_tag_bg_removal_pl.py:450-455); what changes is how bytes move:
//
//  * one producer warp per CTA takes tile tickets and issues 1-D bulk async copies (cp.async.bulk ->
//    UBLKCP, the TMA engine) of the tile's depth / BGR / mask bytes into a kStages-deep shared-memory
//    ring guarded by full/empty mbarriers, so the next tile's loads are in flight while this one is
//    computed.  The ring is kept shallow (2 stages, ticket taken when the stage frees) on purpose: a
//    ticket held in a deep queue delays the prefix chain of its frame for every CTA behind it (measured:
//    3 stages + ticket prefetch 420 k frames/s, 2 stages without 500 k at 1024 frames per launch);
//  * eight compute warps read the tile from shared memory (lane + 32 j ownership, so every ballot is a
//    contiguous run of pixels).  A group of 32 consecutive pixels with no candidate (holes, or depths
//    already beyond the distance mask) costs ~10 instructions and no float64 work;
//  * kept points are compacted into a per-warp shared-memory staging area (x, y, z and the packed BGR
//    bytes), which frees the registers during the arithmetic and lets the input stage return to the
//    producer before the prefix is known; once the tile's base offset is known each warp drains its
//    run with fully coalesced stores (colours are converted to k/255 there, at 100 % lane use);
//  * tiles are handed out frame-interleaved (ticket -> tile t of frame ticket % B), so the B per-frame
//    prefix chains advance in parallel and a tile's predecessor finished a whole round earlier: its
//    inclusive prefix is fetched by ONE load issued before the tile's arithmetic ("early peek"); the
//    decoupled look-back of rv_common.cuh remains as the fallback when the peek finds no prefix yet;
//  * the radius / z-clip / AABB decisions are taken on float32 values with thresholds rounded so that
//    the decision equals the float64 predicate of the oracle bit for bit (exact float64 re-evaluation
//    inside a 2^-20 band around the sphere);
//  * a warp whose 256-pixel run holds no candidate at all (holes, background beyond the distance mask: four
//    runs out of five on the benchmark workload) finds that out with ONE 16-byte shared-memory load per lane
//    and six packed 16-bit min/add instructions, posts a zero total and goes straight to the next tile: it
//    neither computes pixel coordinates nor waits for the other warps.  The run totals are exchanged through
//    an mbarrier (arrive by everybody, wait only by the warps that have points to place and by warp 0,
//    which publishes the tile's prefix), not through a CTA barrier;
//  * colour arrives as BGR8 or as the camera's NV12 frames (luma row + the interleaved chroma rows of the
//    tile's row segments, fetched by the same bulk copies); NV12 is converted per KEPT pixel with the
//    integer arithmetic of cv2.cvtColor, so the BGR image never exists in memory
//    (better_three_capture.py:101-106,159);
//  * colours leave as k/255 floats (three planes) or, RV_COLOR_PACKED8, as one plane of r,g,b,0 bytes.
//
// Eligibility (checked by rv_deproject_fast_eligible): H*W % 16 == 0, W >= 32, 16-byte aligned inputs; NV12 frames
// additionally W % 16 == 0 and W >= 256 (a COMPACT_UNORDERED request is served by the ordered variant).
// Everything else runs k_deproject.
#include "rv_common.cuh"
#include "rv_deproject_args.cuh"

namespace {

constexpr int kCW = RV_K1_CW;            // compute warps
constexpr int kCT = kCW * 32;           // compute threads
constexpr int kThreadsT = kCT + 32;     // + producer warp
constexpr int kItersT = RV_K1_ITERS;
constexpr int kWarpPx = 32 * kItersT;   // 256 pixels per warp per tile
constexpr int kTileT = kCT * kItersT;   // 2048 pixels
#ifndef RV_K1_STAGES
#define RV_K1_STAGES 2
#endif
#ifndef RV_K1_TICKET_AHEAD
#define RV_K1_TICKET_AHEAD 0
#endif
#ifndef RV_K1_TICKET_BATCH
#define RV_K1_TICKET_BATCH 4  // consecutive tickets taken per atomic: the shared counter is touched once per BATCH tiles
#endif
constexpr int kStages = RV_K1_STAGES;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// mbarrier operations on 32-bit shared-window addresses (computed once per thread, outside the tile loop)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Waiting threads suspend in hardware for up to `hint_ns` and are woken by the completing arrive, instead of
// burning issue slots on a poll loop (a plain try_wait returned within a few ns here, and 1 + 8 warps per CTA
// polling cost ~20 % of all issued instructions).
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
#ifndef RV_K1_POLL_NS
#define RV_K1_POLL_NS 0  // experiment: an explicit sleep between polls (0 = none: try_wait's own suspension only)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    if (RV_K1_POLL_NS) __nanosleep(RV_K1_POLL_NS);
  }
}
// 1-D bulk copy global -> shared, completion reported to an mbarrier in bytes (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// shared-memory loads by 32-bit shared-window address.  The per-warp base addresses are computed once and made opaque to
// the compiler (it otherwise rebuilds them from the thread index and the CTA's shared window on every tile: ~25 of the ~85
// instructions a warp spends on a tile that holds nothing for it).  volatile: never moved across the mbarrier waits.
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// n / d for 0 <= n < 2^24 (exact int -> float) with a precomputed float reciprocal and one fix-up; callers fall
// back to the integer divide above that range
__device__ __forceinline__ int fast_div(int n, int d, float rcp, bool small) {
  if (!small) return n / d;
  int q = __float2int_rz(__int2float_rn(n) * rcp);
  const int r = n - q * d;
  if (r < 0) --q;
  if (r >= d) ++q;
  return q;
}

// k / 255 with ONE residual correction: exhaustively exact for k = 0..255 in both precisions
// (tests/test_gpu_cloud.py::test_exact_division_helper_against_numpy walks all 256 bytes).
template <typename T>
__device__ __forceinline__ T unit_color(uint32_t k, bool color_255);
template <>
__device__ __forceinline__ float unit_color<float>(uint32_t k, bool color_255) {
  const float kf = (float)k;
  if (color_255) return kf;
  const float c = 0.003921568859368563f;
  const float q = kf * c;
  return fmaf(fmaf(-q, 255.0f, kf), c, q);
}
template <>
__device__ __forceinline__ double unit_color<double>(uint32_t k, bool color_255) {
  const double kd = (double)k;
  if (color_255) return kd;
  const double c = 0.00392156862745098;
  const double q = kd * c;
  return fma(fma(-q, 255.0, kd), c, q);
}

template <typename T>
__device__ __forceinline__ T qnan();
template <>
__device__ __forceinline__ float qnan<float>() {
  return __int_as_float(0x7fc00000);
}
template <>
__device__ __forceinline__ double qnan<double>() {
  return __longlong_as_double(0x7ff8000000000000ll);
}

template <typename OutT, int DK, bool kGen>
struct Layout {
  static constexpr int kDepthB = DK == RV_DEPTH_U16 ? 2 : 4;
  // colour bytes of a stage: 3 per pixel for BGR8; NV12 uses the first 2 (one luma byte, one byte of its U,V pair).
  // The segmentation-mask bytes and the source-index column exist only in the run-time-flag variant
  static constexpr int kStageBytes = kTileT * (kDepthB + 3 + (kGen ? 1 : 0));
  // per-warp staging: x, y, z (OutT), packed colour (u32) [, source index inside the tile (u16)]
  static constexpr int kWarpStage = kWarpPx * (3 * (int)sizeof(OutT) + 4 + (kGen ? 2 : 0));
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kSmem = kRingBytes + kCW * kWarpStage;
  // resident CTAs per SM the shared memory allows (227 KB usable, 1.4 KB static + 1 KB reserved per CTA)
  static constexpr int kOccSmem = (227 * 1024) / (kSmem + 2560);
#ifndef RV_K1_WARPS_PER_SM
#define RV_K1_WARPS_PER_SM 32
#endif
  static constexpr int kOccCap = RV_K1_WARPS_PER_SM / kCW;  // 32 compute warps per SM: 4 CTAs of 8 warps (or 8 CTAs of 4)
  static constexpr int kOcc = kOccSmem > kOccCap ? kOccCap : kOccSmem;
};

// k / 255 in float32, correctly rounded for every byte: 1/255 split into hi + lo so that fma(k, hi, k * lo) carries
// ~48 bits (tests/test_gpu_cloud.py::test_exact_division_helper_against_numpy walks all 256 bytes)
__device__ __forceinline__ float unit_color_f(float kf) {
  const float hi = 0.003921568859368563f;       // RN(1/255)
  const float lo = -2.319175823606301e-10f;    // RN(1/255 - hi)
  return fmaf(kf, hi, kf * lo);
}

// store to a global address + compile-time byte offset (keeps the address arithmetic out of the unrolled drain)
template <int OFF>
__device__ __forceinline__ void st_global(unsigned long long addr, float v) {
  asm volatile("st.global.f32 [%0+%1], %2;" ::"l"(addr), "n"(OFF), "f"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void st_global(unsigned long long addr, double v) {
  asm volatile("st.global.f64 [%0+%1], %2;" ::"l"(addr), "n"(OFF), "d"(v) : "memory");
}
__device__ __forceinline__ void st_global_u32(unsigned long long addr, uint32_t v) {
  asm volatile("st.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}

// SPEC selects how much of the predicate set is compiled in:
//   0  uint16/float depth with the MUL_F32 unit rule, colour, validity only
//   1  the same plus the radius mask (the canopy / BASELINE configuration)
//   2  everything, decided by the run-time flags of DeprojArgs
//   3  as 1 with the DIV_F32 unit rule (f32(d) / 1000: the RealSense canopy scripts)
// CF (SPEC != 2 only; the run-time-flag variant reads a.color_packed / a.color_nv12): bit 0 = RV_COLOR_PACKED8 output,
// bit 1 = NV12 input.
template <typename OutT, int DK, int MODE, int SPEC, int CF>
__global__ void __launch_bounds__(kThreadsT, (Layout<OutT, DK, SPEC == 2>::kOcc)) k_deproject_tma(const DeprojArgs a) {
  constexpr bool kPacked = MODE == RV_MODE_COMPACT_PACKED;
  constexpr bool kOrdered = MODE == RV_MODE_COMPACT_ORDERED || kPacked;
  constexpr bool kF32 = sizeof(OutT) == 4;
  constexpr bool kGen = SPEC == 2;
  using L = Layout<OutT, DK, kGen>;
  constexpr int kDepthB = L::kDepthB;

  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  // Per-tile exchange slots are eight deep (tile it uses slot it & 7): a warp that finds nothing in its run posts its zero
  // and moves on without waiting, so it can be up to kStages - 1 tiles ahead of the slowest warp of the CTA (the input ring
  // holds it there: a stage is refilled only after all eight warps released it).  kStages <= 7.
  __shared__ __align__(8) uint64_t tot_bar[8];                // the warps' run totals of a tile are posted
  __shared__ __align__(8) uint64_t base_bar[8];               // slow path: warp 0 posts the tile base
  __shared__ int4 s_info[kStages];                            // {status index or -1, frame, tile in frame, pixels in tile}
  __shared__ __align__(16) uint32_t s_tot[8][kCW];            // kept points of the tile's eight warp runs
  __shared__ __align__(8) unsigned long long s_peek[8];       // predecessor status word seen by warp 0
  __shared__ uint32_t s_base[8];
  static_assert(kStages >= 2 && kStages <= 7, "exchange slots are eight deep");

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar), tot_a = smem_u32(tot_bar),
                 base_a = smem_u32(base_bar);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_a + 8 * s, 1);
      mbar_init(empty_a + 8 * s, kCW);
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      mbar_init(tot_a + 8 * s, kCW);
      mbar_init(base_a + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int tpf = a.tiles_per_frame;
  const int nB = a.B;
  const int P = a.P;
  const bool has_bgr = kGen ? (a.bgr != nullptr) : true;
  const bool has_mask = kGen ? (a.use_mask != 0) : false;
  const bool nv12 = kGen ? (a.color_nv12 != 0) : ((CF & 2) != 0);
  const bool pk8 = kGen ? (a.color_packed != 0) : ((CF & 1) != 0);
  const bool small_idx = a.total_tiles < (1 << 24) && P < (1 << 24);
  const uint32_t ring_a = smem_u32(smem);

  // ============================================================ producer warp
  if (warp == kCW) {
    if (lane == 0) {
      const float rcpB = 1.0f / (float)nB;
      const float rcpT = 1.0f / (float)tpf;
      const float rcpWp = 1.0f / (float)a.W;
      // tickets are requested one tile ahead so the atomic's round trip hides behind the wait for a free stage
      int ticket = RV_K1_TICKET_AHEAD ? (int)atomicAdd(a.ticket, 1u) : 0;
      int left = 0;  // tickets still unused from the last batch
      for (int it = 0;; ++it) {
        const int s = it % kStages;
        int next = 0;
        if (RV_K1_TICKET_AHEAD) next = ticket < a.total_tiles ? (int)atomicAdd(a.ticket, 1u) : ticket;
        mbar_wait(empty_a + 8 * s, ((it / kStages) & 1) ^ 1);
        if (!RV_K1_TICKET_AHEAD) {
          // (packed mode is ONE chain through the batch: consecutive tickets depend on each other, so they are taken singly)
          constexpr int kBatch = kPacked ? 1 : RV_K1_TICKET_BATCH;
          if (left == 0) {
            ticket = (int)atomicAdd(a.ticket, (unsigned int)kBatch);
            left = kBatch;
          } else {
            ++ticket;
          }
          --left;
        }
        if (ticket >= a.total_tiles) {
          s_info[s] = make_int4(-1, 0, 0, 0);
          mbar_arrive(full_a + 8 * s);
          break;
        }
        int b, t;
        if (kPacked) {  // one chain through the whole batch: frame-major
          b = fast_div(ticket, tpf, rcpT, small_idx);
          t = ticket - b * tpf;
        } else {  // frame-interleaved: consecutive tickets belong to different per-frame chains
          t = fast_div(ticket, nB, rcpB, small_idx);
          b = ticket - t * nB;
        }
        const int px0 = t * kTileT;
        const int npx = min(kTileT, P - px0);
        s_info[s] = make_int4(b * tpf + t, b, t, npx);
        const long long g = (long long)b * P + px0;
        const uint32_t st = ring_a + (uint32_t)(s * L::kStageBytes);
        const uint32_t fb = full_a + 8 * s;
        uint32_t bytes = (uint32_t)npx * kDepthB;
        if (has_bgr) bytes += (uint32_t)npx * (nv12 ? 2u : 3u);
        if (has_mask) bytes += (uint32_t)npx;
        mbar_arrive_expect_tx(fb, bytes);
        bulk_load(st, reinterpret_cast<const unsigned char *>(a.depth) + g * kDepthB, (uint32_t)npx * kDepthB, fb);
        if (has_bgr) {
          if (nv12) {
            // luma of the tile: one run; chroma: the tile's row segments, each a run of the interleaved U,V row v / 2.
            // Laid out by pixel index inside the tile, so pixel li finds its pair at (li & ~1), (li | 1).
            const unsigned char *frame = a.bgr + (long long)b * (P + (P >> 1));
            bulk_load(st + kTileT * kDepthB, frame + px0, (uint32_t)npx, fb);
            int v = fast_div(px0, a.W, rcpWp, small_idx);
            int u = px0 - v * a.W;
            for (int done = 0; done < npx;) {
              const int len = min(a.W - u, npx - done);
              bulk_load(st + kTileT * (kDepthB + 1) + done, frame + P + (long long)(v >> 1) * a.W + u, (uint32_t)len, fb);
              done += len;
              u = 0;
              ++v;
            }
          } else {
            bulk_load(st + kTileT * kDepthB, a.bgr + g * 3, (uint32_t)npx * 3, fb);
          }
        }
        if (kGen && has_mask) bulk_load(st + kTileT * (kDepthB + 3), a.mask + g, (uint32_t)npx, fb);
        if (RV_K1_TICKET_AHEAD) ticket = next;
      }
    }
    return;
  }

  // ============================================================ compute warps
  const uint32_t lt = (1u << lane) - 1u;
  const float inf_f = __int_as_float(0x7f800000);
  const int W = a.W;
  const bool wide = W >= kWarpPx;  // a warp's 256-pixel run then crosses at most one row boundary
  const float rcpW = 1.0f / (float)W;
  const double cx = a.cx, cy = a.cy, fx = a.fx, fy = a.fy, rfx = a.rfx, rfy = a.rfy;
  const float unit_f = a.unit_scale_f;
  const long long ps = a.plane_stride;
  const bool use_radius = kGen ? (a.use_radius != 0) : (SPEC == 1 || SPEC == 3);
  const bool fast_radius = kGen ? (a.fast_radius != 0) : true;
  const bool color_255 = kGen ? (a.color_255 != 0) : false;
  const int unit_rule = kGen ? a.unit_rule : (SPEC == 3 ? RV_UNIT_DIV_F32 : RV_UNIT_MUL_F32);
  // raw depths in [1, dcand) are candidates; dcand folds "z alone is already beyond the sphere" into an integer compare
  const uint32_t dcand_m1 = ((kF32 && use_radius && fast_radius) ? a.d_cand : 65536u) - 1u;
  // the whole-run rejection test needs nothing but the raw depths; the per-pixel validity image (kGen) and the dense
  // modes have something to write for every pixel, so they walk the groups
  const bool quick = DK == RV_DEPTH_U16 && kOrdered && !(kGen && a.valid);

  // this warp's staging area: room for its whole 256-pixel run
  unsigned char *const wst = smem + L::kRingBytes + (size_t)warp * L::kWarpStage;
  OutT *const sx = reinterpret_cast<OutT *>(wst);
  OutT *const sy = sx + kWarpPx;
  OutT *const sz = sy + kWarpPx;
  uint32_t *const sc = reinterpret_cast<uint32_t *>(sz + kWarpPx);
  uint16_t *const si = reinterpret_cast<uint16_t *>(sc + kWarpPx);  // kGen only

  unsigned long long peek_nxt = 0;  // warp 0 lane 0: status word of the NEXT tile's predecessor, requested a tile early
  int peek_nxt_tile = -1;
  const int w0 = warp * kWarpPx;  // each warp owns 256 consecutive pixels of the tile: its kept points are ONE run of the output
  // this lane's addresses inside stage 0: its eight raw depths of the whole-run test, its pixel of group 0
  uint32_t q_a = ring_a + 2u * (uint32_t)(w0 + 8 * lane);
  uint32_t g_a = ring_a + (uint32_t)kDepthB * (uint32_t)(w0 + lane);
  uint32_t info_a = smem_u32(s_info), full_r = full_a, empty_r = empty_a, tot_r = tot_a;
#ifndef RV_K1_OPAQUE
#define RV_K1_OPAQUE 1
#endif
#if RV_K1_OPAQUE
  asm volatile("" : "+r"(q_a), "+r"(g_a), "+r"(info_a), "+r"(full_r), "+r"(empty_r), "+r"(tot_r));
#endif

  for (int it = 0;; ++it) {
    const int s = it % kStages;
    mbar_wait(full_r + 8 * s, (it / kStages) & 1);
    const uint4 info = lds_v4(info_a + 16 * s);
    const int tile = (int)info.x;  // index into status[]
    if (tile < 0) break;
    const int b = (int)info.y, t = (int)info.z, npx = (int)info.w;
    const uint32_t so = (uint32_t)(s * L::kStageBytes);
    const int n_pred = kPacked ? tile : t;
    const int px0 = t * kTileT;
    unsigned char *st = smem + (size_t)s * L::kStageBytes;
    const uint8_t *s_col = st + kTileT * kDepthB;
    const uint8_t *s_msk = st + kTileT * (kDepthB + 3);

    // early peek (warp 0, lane 0): the predecessor's status word.  The request for the NEXT tile is issued here as well
    // when that tile's descriptor has already landed, which hides its L2 round trip behind a whole tile of work.
    if (kOrdered && warp == 0 && lane == 0) {
      unsigned long long peek = 0;
      if (n_pred > 0) {
        const bool have = peek_nxt_tile == tile && (peek_nxt >> 62) == 2;  // a word without a prefix is stale by now
        peek = have ? peek_nxt : rv_ld_relaxed(a.status + tile - 1);
      }
      const int sn = (it + 1) % kStages;
      peek_nxt_tile = -1;
      if (mbar_try_wait(full_a + 8 * sn, ((it + 1) / kStages) & 1)) {
        const int4 nx = s_info[sn];
        if (nx.x >= 0 && (kPacked ? nx.x : nx.z) > 0) {
          peek_nxt = rv_ld_relaxed(a.status + nx.x - 1);
          peek_nxt_tile = nx.x;
        }
      }
      s_peek[it & 7] = peek;  // read by the warps that drain, after the totals barrier
    }

    // a partial last tile: zero the depth of this warp's own pixels beyond the frame so the loops need no bounds test
    if (npx < kTileT) {
#pragma unroll
      for (int j = 0; j < kItersT; ++j) {
        const int li = w0 + j * 32 + lane;
        if (li >= npx) {
          if (DK == RV_DEPTH_U16) reinterpret_cast<uint16_t *>(st)[li] = 0;
          else reinterpret_cast<float *>(st)[li] = 0.0f;
        }
      }
      __syncwarp();
    }

    // ---- whole-run rejection: eight raw depths per lane in one load, min over (d - 1) mod 2^16 with packed 16-bit ops
    bool any_cand = true;
    if (quick) {
      const uint4 q = lds_v4(q_a + so);
      const uint32_t m2 = __vminu2(__vminu2(__vsub2(q.x, 0x00010001u), __vsub2(q.y, 0x00010001u)),
                                   __vminu2(__vsub2(q.z, 0x00010001u), __vsub2(q.w, 0x00010001u)));
      any_cand = __any_sync(0xffffffffu, min(m2 & 0xffffu, m2 >> 16) < dcand_m1);
    }

    uint32_t run = 0;  // kept points of this warp so far (warp-uniform)
    if (any_cand) {
      const int p0 = px0 + w0 + lane;  // this lane's pixel in group j = 0
      const int v0 = fast_div(p0, W, rcpW, small_idx);
      const int u0 = p0 - v0 * W;
#pragma unroll
      for (int j = 0; j < kItersT; ++j) {
        const int li = w0 + j * 32 + lane;  // index inside the tile
        float z32 = 0.0f;
        double z64 = 0.0;
        uint32_t draw = 0;
        bool ok;
        if (DK == RV_DEPTH_U16) {
          draw = lds_u16(g_a + so + (uint32_t)(j * 32 * kDepthB));
          ok = (draw - 1u) < dcand_m1;  // draw != 0 && draw < dcand
        } else {
          z32 = lds_f32(g_a + so + (uint32_t)(j * 32 * kDepthB));
          ok = (z32 > 0.0f) && (z32 < inf_f);
          if (kF32 && use_radius && fast_radius) ok = ok && (z32 * z32 < a.r2_hi_f);
        }
        if (kGen && has_mask) {
          const uint32_t m = s_msk[li];
          ok = ok && (a.invert_mask ? (m == 0) : (m != 0));
        }

        OutT xo = (OutT)0, yo = (OutT)0;
        if (__any_sync(0xffffffffu, ok)) {  // warp-uniform: 32 consecutive holes / far pixels cost no geometry
          if (DK == RV_DEPTH_U16) {
            const float df = (float)draw;
            if (unit_rule == RV_UNIT_MUL_F32) {
              z32 = df * unit_f;
            } else if (unit_rule == RV_UNIT_DIV_F32) {
              z32 = rv_divf(df, unit_f, a.unit_rcp_f);
            } else {
              z64 = rv_div((double)draw, a.unit_scale, a.unit_rcp);
              z32 = (float)z64;
            }
          }
          if (kGen && a.use_trunc) ok = ok && !(z32 >= a.trunc_f);
          int uj = u0 + j * 32, vj = v0;
          if (wide) {
            if (uj >= W) {
              uj -= W;
              ++vj;
            }
          } else {
            const int p = p0 + j * 32;
            vj = fast_div(p, W, rcpW, small_idx);
            uj = p - vj * W;
          }
          if (!(DK == RV_DEPTH_U16 && unit_rule == RV_UNIT_DIV_F64)) z64 = (double)z32;
          double x64, y64;
          if (kGen && a.rays) {  // distorted camera: the normalised ray of every pixel comes from the (L2-resident) table
            double2 r = make_double2(0.0, 0.0);
            if (li < npx) r = __ldg(a.rays + px0 + li);
            x64 = z64 * r.x;
            y64 = z64 * r.y;
          } else {
            x64 = rv_div(((double)uj - cx) * z64, fx, rfx);
            y64 = rv_div(((double)vj - cy) * z64, fy, rfy);
          }
          xo = (OutT)x64;
          yo = (OutT)y64;
          if (kF32) {
            // float32 storage: compare the stored floats against thresholds rounded toward the kept side;
            // identical to the float64 predicate on their exact up-casts
            const float xf = (float)xo, yf = (float)yo;
            if (kGen) {
              if (a.use_zclip) ok = ok && (z32 >= a.zmin_f) && (z32 <= a.zmax_f);
              if (a.use_aabb)
                ok = ok && (xf >= a.amin_f[0]) && (xf <= a.amax_f[0]) && (yf >= a.amin_f[1]) && (yf <= a.amax_f[1]) &&
                     (z32 >= a.amin_f[2]) && (z32 <= a.amax_f[2]);
            }
            if (use_radius) {
              bool in;
              if (fast_radius) {
                const float sf = fmaf(z32, z32, fmaf(yf, yf, xf * xf));
                in = sf <= a.r2_lo_f;
                if (sf > a.r2_lo_f && sf < a.r2_hi_f) {  // inside the 2^-20 band: the float64 sum decides
                  const double X = (double)xf, Y = (double)yf, Z = (double)z32;
                  in = ((X * X + Y * Y) + Z * Z) < a.r2_thresh;
                }
              } else {
                const double X = (double)xf, Y = (double)yf, Z = (double)z32;
                in = ((X * X + Y * Y) + Z * Z) < a.r2_thresh;
              }
              ok = ok && in;
            }
          } else {
            const double X = (double)xo, Y = (double)yo, Z = z64;
            if (kGen) {
              if (a.use_zclip) ok = ok && (Z >= a.z_min) && (Z <= a.z_max);
              if (a.use_aabb)
                ok = ok && (X >= a.amin[0]) && (X <= a.amax[0]) && (Y >= a.amin[1]) && (Y <= a.amax[1]) && (Z >= a.amin[2]) &&
                     (Z <= a.amax[2]);
            }
            if (use_radius) ok = ok && (((X * X + Y * Y) + Z * Z) < a.r2_thresh);
          }
          const uint32_t bal = __ballot_sync(0xffffffffu, ok);
          // ---- stage the kept lanes behind the warp's earlier groups (compact modes)
          if (kOrdered && ok) {
            const uint32_t pos = run + __popc(bal & lt);
            sx[pos] = xo;
            sy[pos] = yo;
            sz[pos] = kF32 ? (OutT)z32 : (OutT)z64;
            if (has_bgr) {
              if (nv12) {
                sc[pos] = rv_nv12_pixel_bgr(s_col[li], s_col[kTileT + (li & ~1)], s_col[kTileT + (li | 1)]);
              } else {
                const uint8_t *c = s_col + 3 * li;
                sc[pos] = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16);
              }
            }
            if (kGen) si[pos] = (uint16_t)li;
          }
          run += __popc(bal);
        } else {
          ok = false;
        }
        if (kGen && a.valid && li < npx) a.valid[(long long)b * P + px0 + li] = ok ? 1 : 0;
        if (!kOrdered) {  // dense modes fill every lane's slot
          const OutT bad = (MODE == RV_MODE_DENSE_NAN) ? qnan<OutT>() : (OutT)0;
          const int pos = j * 32 + lane;
          sx[pos] = ok ? xo : bad;
          sy[pos] = ok ? yo : bad;
          sz[pos] = ok ? (kF32 ? (OutT)z32 : (OutT)z64) : bad;
          uint32_t cpk = 0;
          if (has_bgr && ok) {
            if (nv12) {
              cpk = rv_nv12_pixel_bgr(s_col[li], s_col[kTileT + (li & ~1)], s_col[kTileT + (li | 1)]);
            } else {
              const uint8_t *c = s_col + 3 * li;
              cpk = (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16);
            }
          }
          sc[pos] = cpk;
          if (kGen) si[pos] = ok ? (uint16_t)li : (uint16_t)0xffffu;
        }
      }
    }
    // the input stage is no longer needed: hand it back to the producer before the prefix is resolved (the run total is
    // posted first: the stage release is what lets the other warps, and this one, run ahead)
    if (npx < kTileT) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // our zero fill vs the next bulk copy
    __syncwarp();
    if (lane == 0) {
      if (kOrdered) {
        s_tot[it & 7][warp] = run;
        mbar_arrive(tot_r + 8 * (it & 7));
      }
      mbar_arrive(empty_r + 8 * s);
    }

    uint32_t base = 0, off = 0;
    if (kOrdered) {
      // ---------------- the eight run totals.  Everybody posts; only the warps that have points to place wait for the
      // others, and warp 0, which publishes the tile's prefix for the frame's chain.  (Letting the warp that posted last
      // publish instead, so that nobody waits without need, measured 1.3 % slower: one more shared-memory atomic per warp and
      // tile, and the early peek stays with warp 0 anyway.)
      const bool publisher = warp == 0;
      if (run == 0 && !publisher) continue;  // warp-uniform
      mbar_wait(tot_r + 8 * (it & 7), (uint32_t)(it >> 3) & 1u);
      const uint32_t t8 = lane < kCW ? s_tot[it & 7][lane] : 0u;
      const uint32_t tile_total = __reduce_add_sync(0xffffffffu, t8);
      off = __reduce_add_sync(0xffffffffu, lane < warp ? t8 : 0u);

      // ---------------- tile base
      bool hit = false;
#ifdef RV_K1_NO_CHAIN  // timing experiment only (wrong offsets): what the kernel would cost without the prefix chain
      if (false) {
#else
      if (n_pred > 0) {
#endif
        const unsigned long long pk = s_peek[it & 7];
        hit = (pk >> 62) == 2;
        if (hit) {
          base = (uint32_t)pk;
        } else if (publisher) {
          base = rv_lookback(a.status, tile, n_pred, tile_total);  // publishes the aggregate, then the prefix
          if (lane == 0) {
            s_base[it & 7] = base;
            mbar_arrive(base_a + 8 * (it & 7));
          }
        } else {
          // the predecessor is still in flight in another CTA: sleep until the publisher has walked the chain
          mbar_wait(base_a + 8 * (it & 7), (uint32_t)(it >> 3) & 1u);
          base = s_base[it & 7];
        }
      }
      if (publisher && lane == 0) {
        if (n_pred == 0 || hit) {
          rv_st_relaxed(a.status + tile, RV_ST_PREFIX | (unsigned long long)(base + tile_total));
          mbar_arrive(base_a + 8 * (it & 7));  // keep the slot's phase in step with the iteration count
        }
        if (kPacked) {
          if (t == 0) a.counts[b] = base;
          if (tile == a.total_tiles - 1) a.counts[nB] = (unsigned long long)base + tile_total;
        } else if (t == tpf - 1) {
          a.counts[b] = (unsigned long long)base + tile_total;
        }
      }
    } else {
      if (lane == 0 && run) atomicAdd(a.counts + b, (unsigned long long)run);
    }

    // ---------------- drain: the warp's packed run -> its contiguous place in every plane, all lanes busy
    {
      const long long fout = kPacked ? 0ll : (long long)b * a.frame_stride;
      const unsigned long long cap64 = kPacked ? (unsigned long long)ps : (unsigned long long)a.frame_stride;
      const uint32_t cap = cap64 > 0xffffffffull ? 0xffffffffu : (uint32_t)cap64;
      OutT *const o0 = reinterpret_cast<OutT *>(a.out) + fout;
      const uint32_t n = kOrdered ? run : (uint32_t)max(0, min(kWarpPx, npx - w0));
      const uint32_t g0 = kOrdered ? base + off : (uint32_t)(px0 + w0);
      if (n != 0) {
        // plane pointers once per tile; the unrolled iterations below then address with immediates only
        const unsigned long long pb = (unsigned long long)ps * sizeof(OutT);
        const unsigned long long ox = (unsigned long long)__cvta_generic_to_global(o0 + g0 + lane);
        const unsigned long long oy = ox + pb, oz = oy + pb, orr = oz + pb, og = orr + pb, ob = og + pb;
        const uint32_t room = g0 < cap ? cap - g0 : 0u;  // output slots left in this frame (capacity overflow is reported)
        const uint32_t m = n < room ? n : room;
#pragma unroll
        for (int j = 0; j < kItersT; ++j) {
          if ((uint32_t)(j * 32) >= m) break;  // warp-uniform
          const int k = j * 32 + lane;
          constexpr int kStep = 32 * (int)sizeof(OutT);
          if ((uint32_t)k < m) {
            st_global<0>(ox + j * kStep, sx[k]);
            st_global<0>(oy + j * kStep, sy[k]);
            st_global<0>(oz + j * kStep, sz[k]);
            if (has_bgr) {
              const uint32_t c = sc[k];
              if (kF32 && pk8) {
                st_global_u32(orr + j * kStep, __byte_perm(c, 0u, 0x4012));  // b,g,r,0 -> r,g,b,0
              } else if (kF32 && !kGen) {
                st_global<0>(orr + j * kStep, (OutT)unit_color_f((float)((c >> 16) & 255u)));
                st_global<0>(og + j * kStep, (OutT)unit_color_f((float)((c >> 8) & 255u)));
                st_global<0>(ob + j * kStep, (OutT)unit_color_f((float)(c & 255u)));
              } else {
                st_global<0>(orr + j * kStep, unit_color<OutT>((c >> 16) & 255u, color_255));
                st_global<0>(og + j * kStep, unit_color<OutT>((c >> 8) & 255u, color_255));
                st_global<0>(ob + j * kStep, unit_color<OutT>(c & 255u, color_255));
              }
            }
            if (kGen && a.src_index) {
              const uint32_t idx = si[k];
              a.src_index[fout + g0 + k] = (!kOrdered && idx == 0xffffu) ? -1 : px0 + (int)idx;
            }
          }
        }
      }
    }
    __syncwarp();  // the staging area is rewritten by the next tile
  }
}

// SPEC 0/1/3 cover the plain configurations; anything else falls to the run-time-flag variant
int pick_spec(const DeprojArgs &a, int depth_kind) {
  const bool plain = a.bgr && !a.rays && !a.use_mask && !a.use_trunc && !a.use_zclip && !a.use_aabb && !a.valid && !a.src_index &&
                     !a.color_255;
  if (!plain) return 2;
  const bool mul = depth_kind != RV_DEPTH_U16 || a.unit_rule == RV_UNIT_MUL_F32;
  const bool div = depth_kind == RV_DEPTH_U16 && a.unit_rule == RV_UNIT_DIV_F32;
  if (mul && !a.use_radius) return 0;
  if (mul && a.fast_radius) return 1;
  if (div && a.use_radius && a.fast_radius) return 3;  // the canopy configuration: f32(d) / 1000 and a distance mask
  return 2;
}

template <typename OutT, int DK, int MODE, int SPEC, int CF>
cudaError_t launch_one(const rv_ctx *ctx, const DeprojArgs &a, cudaStream_t st) {
  const size_t smem = (size_t)Layout<OutT, DK, SPEC == 2>::kSmem;
  auto k = k_deproject_tma<OutT, DK, MODE, SPEC, CF>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = rv_persistent_grid(ctx, k, kThreadsT, smem, a.total_tiles);
  k<<<grid, kThreadsT, smem, st>>>(a);
  return cudaSuccess;
}

// the colour variants (NV12 in, packed bytes out) are compiled for the configurations the host pipeline runs: uint16
// depth, float32 storage, the compact modes, plain predicates; every other combination takes the run-time-flag variant
template <typename OutT, int DK, int MODE, int SPEC>
cudaError_t launch_cf(const rv_ctx *ctx, const DeprojArgs &a, cudaStream_t st) {
  constexpr bool kHasCf = sizeof(OutT) == 4 && DK == RV_DEPTH_U16 &&
                          (MODE == RV_MODE_COMPACT_ORDERED || MODE == RV_MODE_COMPACT_PACKED) && SPEC != 2;
  const int cf = (a.color_packed ? 1 : 0) | (a.color_nv12 ? 2 : 0);
  if (cf == 0 || SPEC == 2) return launch_one<OutT, DK, MODE, SPEC, 0>(ctx, a, st);
  if (kHasCf) {
    if (cf == 1) return launch_one<OutT, DK, MODE, kHasCf ? SPEC : 2, kHasCf ? 1 : 0>(ctx, a, st);
    if (cf == 2) return launch_one<OutT, DK, MODE, kHasCf ? SPEC : 2, kHasCf ? 2 : 0>(ctx, a, st);
    return launch_one<OutT, DK, MODE, kHasCf ? SPEC : 2, kHasCf ? 3 : 0>(ctx, a, st);
  }
  return launch_one<OutT, DK, MODE, 2, 0>(ctx, a, st);
}

template <typename OutT, int DK, int MODE>
cudaError_t launch_spec(const rv_ctx *ctx, const DeprojArgs &a, int spec, cudaStream_t st) {
  if (spec == 0) return launch_cf<OutT, DK, MODE, 0>(ctx, a, st);
  if (spec == 1) return launch_cf<OutT, DK, MODE, 1>(ctx, a, st);
  if (spec == 3 && DK == RV_DEPTH_U16) return launch_cf<OutT, DK, MODE, 3>(ctx, a, st);
  return launch_cf<OutT, DK, MODE, 2>(ctx, a, st);
}

template <typename OutT, int DK>
cudaError_t launch_tma(const rv_ctx *ctx, const DeprojArgs &a, int mode, cudaStream_t st) {
  const int spec = pick_spec(a, DK);
  switch (mode) {
    case RV_MODE_COMPACT_ORDERED: return launch_spec<OutT, DK, RV_MODE_COMPACT_ORDERED>(ctx, a, spec, st);
    case RV_MODE_COMPACT_PACKED: return launch_spec<OutT, DK, RV_MODE_COMPACT_PACKED>(ctx, a, spec, st);
    case RV_MODE_DENSE_ZERO: return launch_spec<OutT, DK, RV_MODE_DENSE_ZERO>(ctx, a, spec, st);
    default: return launch_spec<OutT, DK, RV_MODE_DENSE_NAN>(ctx, a, spec, st);
  }
}

}  // namespace

bool rv_deproject_fast_eligible(const DeprojArgs &a, int mode) {
  (void)mode;
  if (a.W < 32 || (a.P % 16) != 0) return false;
  if (!rv_aligned(a.depth, 16) || (a.bgr && !rv_aligned(a.bgr, 16)) || (a.mask && !rv_aligned(a.mask, 16))) return false;
  if (a.bgr && a.color_nv12 && ((a.W % 16) != 0 || a.W < 256)) return false;  // chroma row segments as 16-byte bulk copies
  return true;
}

cudaError_t rv_deproject_fast_launch(const rv_ctx *ctx, const DeprojArgs &a, int mode, int out_dtype, int depth_kind,
                                     cudaStream_t st) {
  if (out_dtype == RV_F32) {
    if (depth_kind == RV_DEPTH_U16) return launch_tma<float, RV_DEPTH_U16>(ctx, a, mode, st);
    return launch_tma<float, RV_DEPTH_F32_METERS>(ctx, a, mode, st);
  }
  if (depth_kind == RV_DEPTH_U16) return launch_tma<double, RV_DEPTH_U16>(ctx, a, mode, st);
  return launch_tma<double, RV_DEPTH_F32_METERS>(ctx, a, mode, st);
}
