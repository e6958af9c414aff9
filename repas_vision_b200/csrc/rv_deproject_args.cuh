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
  int color_packed;  // RV_COLOR_PACKED8: plane 3 holds the bytes r,g,b,0 of every point (float32 output only)
  int color_nv12;    // `bgr` points at NV12 frames ([H*3/2, W] bytes each)
  int geom_f32;      // RV_GEOM_SDK_F32: x = z * ((u - ppx) / fx) in float32 (generic kernel only)
  float cx_f, cy_f, fx_f, fy_f, rfx_f, rfy_f;
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

// cv2.cvtColor(COLOR_YUV2BGR_NV12): OpenCV's fixed-point ITU-R BT.601 conversion with 20-bit coefficients (the arithmetic
// rv_misc.cu checks bit for bit against cv2).  Returns the bytes b | g << 8 | r << 16.
__device__ __forceinline__ uint32_t rv_nv12_pixel_bgr(uint32_t Y, uint32_t U, uint32_t V) {
  const int u = (int)U - 128, v = (int)V - 128;
  const int ruv = (1 << 19) + 1673527 * v;
  const int guv = (1 << 19) - 852492 * v - 409993 * u;
  const int buv = (1 << 19) + 2116026 * u;
  const int y = max(0, (int)Y - 16) * 1220542;
  const uint32_t b = (uint32_t)min(max((y + buv) >> 20, 0), 255);
  const uint32_t g = (uint32_t)min(max((y + guv) >> 20, 0), 255);
  const uint32_t r = (uint32_t)min(max((y + ruv) >> 20, 0), 255);
  return b | (g << 8) | (r << 16);
}

bool rv_deproject_fast_eligible(const DeprojArgs &a, int mode);
cudaError_t rv_deproject_fast_launch(const rv_ctx *ctx, const DeprojArgs &a, int mode, int out_dtype, int depth_kind,
                                     cudaStream_t st);
