// rv_deproject_args.cuh -- kernel argument block shared by the two K1 kernels
// (k_deproject in rv_deproject.cu: generic; k_deproject_tma in rv_deproject_tma.cu: TMA-fed fast path).
#pragma once
#include "rv_common.cuh"

struct DeprojArgs {
  const void *depth;
  const uint8_t *bgr;
  const uint8_t *mask;
  const double2 *rays;
  void *out;
  uint8_t *valid;
  int32_t *src_index;
  unsigned long long *counts;
  unsigned long long *status;  // ws + 128 B
  unsigned int *ticket;        // ws
  int B, H, W, P;
  int tiles_per_frame, total_tiles;
  long long plane_stride, frame_stride;
  double cx, cy, fx, fy, rfx, rfy;
  double unit_scale, unit_rcp;
  float unit_scale_f, unit_rcp_f;
  float trunc_f;
  int unit_rule;
  double z_min, z_max;
  double r2_thresh;  // keep iff (x*x+y*y)+z*z < r2_thresh  <=>  sqrt(.) < r_max
  double amin[3], amax[3];
  int use_mask, invert_mask, use_trunc, use_zclip, use_radius, use_aabb;
  int color_255;
  // float32 forms of the cloud predicates (exactly equivalent on float32 storage; rv_deproject_tma.cu)
  float zmin_f, zmax_f;        // smallest float >= z_min, largest float <= z_max
  float amin_f[3], amax_f[3];  // same rounding for the box
  float r2_lo_f, r2_hi_f;      // below lo: inside for sure; at or above hi: outside for sure; between: float64 decides
  int fast_radius;
  unsigned int d_cand;  // uint16 depths in [1, d_cand) can still be inside the sphere; 65536 when there is no radius mask
};

// fast path (rv_deproject_tma.cu): RV_K1_CW compute warps per CTA, 32 * RV_K1_ITERS pixels each per tile
#ifndef RV_K1_CW
#define RV_K1_CW 8
#endif
#ifndef RV_K1_ITERS
#define RV_K1_ITERS 8  // 32-pixel groups per warp and tile
#endif
constexpr int kFastTilePx = RV_K1_CW * 32 * RV_K1_ITERS;
constexpr int kGenericTilePx = 2048;
constexpr int kMinTilePx = kFastTilePx < kGenericTilePx ? kFastTilePx : kGenericTilePx;  // sizes the workspace

bool rv_deproject_fast_eligible(const DeprojArgs &a, int mode);
cudaError_t rv_deproject_fast_launch(const rv_ctx *ctx, const DeprojArgs &a, int mode, int out_dtype, int depth_kind,
                                     cudaStream_t st);
