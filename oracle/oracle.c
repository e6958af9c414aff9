/*
 * oracle.c -- CPU oracle loops (plain C) for the parts of the hot path whose
 * reference arithmetic is a per-element loop inside a vendor library.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  The product never calls it.
 *
 * PARITY UNPINNED for both functions: librealsense (pyrealsense2==2.56.5.9235) and
 * Open3D (open3d==0.19.0) are pip dependencies of the reference
 * (realsense_d415i/requirements.txt, femto_bolt_code/requirements.txt), not part of
 * its checkout, and the reference stores no expected outputs for these calls.  The
 * published algorithms are restated from SURVEY.md Appendix B.1 / B.3; call sites:
 *   rs.align(rs.stream.color).process   realsense_d415i/capture_scripts/capture_aligned_all.py:75,197
 *   AlignFilter(COLOR_STREAM).process   femto_bolt_code/scripts/better_three_capture.py:169,188
 *   pcd.voxel_down_sample(voxel)        femto_bolt_code/scripts/mpa_icp_export.py:44,174
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: every float op rounds once).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_DIST_NONE = 0, ORC_DIST_BROWN_CONRADY = 1, ORC_DIST_INVERSE_BROWN_CONRADY = 2, ORC_DIST_MODIFIED_BROWN_CONRADY = 3 };

typedef struct orc_cam {
  float fx, fy, ppx, ppy;
  float coeffs[5]; /* k1 k2 p1 p2 k3 */
  int32_t model;
  int32_t width, height;
} orc_cam;

/* rs2_deproject_pixel_to_point, float32 */
static void deproject(float point[3], const orc_cam *in, float px, float py, float depth) {
  float x = (px - in->ppx) / in->fx;
  float y = (py - in->ppy) / in->fy;
  const float *c = in->coeffs;
  if (in->model == ORC_DIST_INVERSE_BROWN_CONRADY) {
    float r2 = x * x + y * y;
    float f = 1 + c[0] * r2 + c[1] * r2 * r2 + c[4] * r2 * r2 * r2;
    float ux = x * f + 2 * c[2] * x * y + c[3] * (r2 + 2 * x * x);
    float uy = y * f + 2 * c[3] * x * y + c[2] * (r2 + 2 * y * y);
    x = ux;
    y = uy;
  } else if (in->model == ORC_DIST_BROWN_CONRADY) {
    float xo = x, yo = y;
    for (int i = 0; i < 10; i++) {
      float r2 = x * x + y * y;
      float icdist = (float)1 / (float)(1 + ((c[4] * r2 + c[1]) * r2 + c[0]) * r2);
      float dx = 2 * c[2] * x * y + c[3] * (r2 + 2 * x * x);
      float dy = 2 * c[3] * x * y + c[2] * (r2 + 2 * y * y);
      x = (xo - dx) * icdist;
      y = (yo - dy) * icdist;
    }
  }
  point[0] = depth * x;
  point[1] = depth * y;
  point[2] = depth;
}

/* rs2_project_point_to_pixel, float32 */
static void project(float pixel[2], const orc_cam *in, const float p[3]) {
  float x = p[0] / p[2], y = p[1] / p[2];
  const float *c = in->coeffs;
  if (in->model == ORC_DIST_MODIFIED_BROWN_CONRADY || in->model == ORC_DIST_INVERSE_BROWN_CONRADY) {
    float r2 = x * x + y * y;
    float f = 1 + c[0] * r2 + c[1] * r2 * r2 + c[4] * r2 * r2 * r2;
    x *= f;
    y *= f;
    float dx = x + 2 * c[2] * x * y + c[3] * (r2 + 2 * x * x);
    float dy = y + 2 * c[3] * x * y + c[2] * (r2 + 2 * y * y);
    x = dx;
    y = dy;
  } else if (in->model == ORC_DIST_BROWN_CONRADY) {
    float r2 = x * x + y * y;
    float f = 1 + c[0] * r2 + c[1] * r2 * r2 + c[4] * r2 * r2 * r2;
    float xf = x * f, yf = y * f;
    float dx = xf + 2 * c[2] * x * y + c[3] * (r2 + 2 * x * x);
    float dy = yf + 2 * c[3] * x * y + c[2] * (r2 + 2 * y * y);
    x = dx;
    y = dy;
  }
  pixel[0] = x * in->fx + in->ppx;
  pixel[1] = y * in->fy + in->ppy;
}

/* rs2_transform_point_to_point: rotation stored column-major */
static void xform(float to[3], const float R[9], const float t[3], const float from[3]) {
  to[0] = R[0] * from[0] + R[3] * from[1] + R[6] * from[2] + t[0];
  to[1] = R[1] * from[0] + R[4] * from[1] + R[7] * from[2] + t[1];
  to[2] = R[2] * from[0] + R[5] * from[1] + R[8] * from[2] + t[2];
}

/* (int)(p + 0.5f) with the undefined range made explicit: NaN or |p| >= 2^30 rejects */
static int round_pix(float p, int *out) {
  float q = p + 0.5f;
  if (!(fabsf(q) < 1073741824.0f)) return 0;
  *out = (int)q;
  return 1;
}

/* align z16 depth to the other (colour) camera: one frame.
 * out [Hc*Wc] u16 (0 = empty); winner [Hc*Wc] i32 source index or -1 (may be NULL). */
int orc_register_z16(const uint16_t *depth, int Hd, int Wd, const orc_cam *dcam, const orc_cam *ccam,
                     const float R[9], const float t[3], float depth_units, uint16_t *out, int32_t *winner) {
  const int Wc = ccam->width, Hc = ccam->height;
  memset(out, 0, (size_t)Wc * Hc * sizeof(uint16_t));
  if (winner)
    for (int64_t i = 0; i < (int64_t)Wc * Hc; i++) winner[i] = -1;
  for (int dy = 0; dy < Hd; dy++) {
    for (int dx = 0; dx < Wd; dx++) {
      const int src = dy * Wd + dx;
      const uint16_t z = depth[src];
      if (!z) continue;
      const float d = (float)z * depth_units;
      float pt[3], q[3], pix[2];
      int x0, y0, x1, y1;
      deproject(pt, dcam, (float)dx - 0.5f, (float)dy - 0.5f, d);
      xform(q, R, t, pt);
      project(pix, ccam, q);
      if (!round_pix(pix[0], &x0) || !round_pix(pix[1], &y0)) continue;
      deproject(pt, dcam, (float)dx + 0.5f, (float)dy + 0.5f, d);
      xform(q, R, t, pt);
      project(pix, ccam, q);
      if (!round_pix(pix[0], &x1) || !round_pix(pix[1], &y1)) continue;
      if (x0 < 0 || y0 < 0 || x1 >= Wc || y1 >= Hc) continue;
      for (int y = y0; y <= y1; y++)
        for (int x = x0; x <= x1; x++) {
          const int64_t o = (int64_t)y * Wc + x;
          if (out[o] == 0 || z < out[o]) { /* strict: lowest source index keeps ties */
            out[o] = z;
            if (winner) winner[o] = src;
          }
        }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ voxel grid */
typedef struct {
  int32_t k[3];
  int32_t used;
  int64_t idx; /* dense output slot, insertion order */
} slot_t;

static uint64_t hash3(const int32_t k[3]) {
  uint64_t h = (uint32_t)k[0] * 0x9E3779B97F4A7C15ull;
  h ^= ((uint32_t)k[1] + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
  h ^= ((uint32_t)k[2] + 0x165667B1ull) * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
  return h;
}

/* Open3D PointCloud::VoxelDownSample.  pts / cols: [n,3] float64 AoS (cols may be NULL).
 * Outputs in first-insertion order: keys [cap,3] i32, cent [cap,3], col [cap,3], cnt [cap].
 * Returns M, or -1 (voxel_size <= 0), -2 (voxel too small for the extent), -3 (cap too small). */
int64_t orc_voxel_down_sample(const double *pts, const double *cols, int64_t n, double voxel, int32_t *keys,
                              double *cent, double *col, int64_t *cnt, int64_t cap) {
  if (!(voxel > 0.0)) return -1;
  if (n == 0) return 0;
  double mn[3] = {pts[0], pts[1], pts[2]}, mx[3] = {pts[0], pts[1], pts[2]};
  for (int64_t i = 1; i < n; i++)
    for (int a = 0; a < 3; a++) {
      double v = pts[3 * i + a];
      if (v < mn[a]) mn[a] = v;
      if (v > mx[a]) mx[a] = v;
    }
  double ext = mx[0] - mn[0];
  if (mx[1] - mn[1] > ext) ext = mx[1] - mn[1];
  if (mx[2] - mn[2] > ext) ext = mx[2] - mn[2];
  if (voxel * 2147483647.0 < ext) return -2;
  const double org[3] = {mn[0] - voxel * 0.5, mn[1] - voxel * 0.5, mn[2] - voxel * 0.5};
  uint64_t tsz = 16;
  while (tsz < (uint64_t)n * 2) tsz <<= 1;
  slot_t *tab = (slot_t *)calloc(tsz, sizeof(slot_t));
  if (!tab) return -4;
  int64_t m = 0;
  for (int64_t i = 0; i < n; i++) {
    int32_t k[3];
    for (int a = 0; a < 3; a++) k[a] = (int32_t)floor((pts[3 * i + a] - org[a]) / voxel);
    uint64_t h = hash3(k) & (tsz - 1);
    while (tab[h].used && (tab[h].k[0] != k[0] || tab[h].k[1] != k[1] || tab[h].k[2] != k[2])) h = (h + 1) & (tsz - 1);
    if (!tab[h].used) {
      if (m >= cap) {
        free(tab);
        return -3;
      }
      tab[h].used = 1;
      tab[h].k[0] = k[0];
      tab[h].k[1] = k[1];
      tab[h].k[2] = k[2];
      tab[h].idx = m;
      for (int a = 0; a < 3; a++) {
        keys[3 * m + a] = k[a];
        cent[3 * m + a] = 0.0;
        if (cols) col[3 * m + a] = 0.0;
      }
      cnt[m] = 0;
      m++;
    }
    const int64_t j = tab[h].idx;
    for (int a = 0; a < 3; a++) {
      cent[3 * j + a] += pts[3 * i + a]; /* index order, float64 */
      if (cols) col[3 * j + a] += cols[3 * i + a];
    }
    cnt[j]++;
  }
  for (int64_t j = 0; j < m; j++)
    for (int a = 0; a < 3; a++) {
      cent[3 * j + a] /= (double)cnt[j];
      if (cols) col[3 * j + a] /= (double)cnt[j];
    }
  free(tab);
  return m;
}

/* ------------------------------------------------------ masked deprojection (C) */
/* The float64 loop form of create_masked_pointcloud (create_masked_ply.py:74-100) plus the
 * distance mask of distance_masking_on_ply.py:12-19, one frame, u16 depth with the
 * MUL_F32 unit rule.  Used as a compiled single-thread CPU baseline beside the numpy port.
 * pts/cols: [cap,3] float64.  Returns the number of kept points. */
int64_t orc_deproject_masked(const uint16_t *depth, const uint8_t *bgr, const uint8_t *mask, int H, int W, double fx,
                             double fy, double cx, double cy, float scale, int use_radius, double r_max, double *pts,
                             double *cols) {
  int64_t m = 0;
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      const int64_t i = (int64_t)v * W + u;
      if (mask && !mask[i]) continue;
      if (!depth[i]) continue;
      const double z = (double)((float)depth[i] * scale);
      const double x = ((double)u - cx) * z / fx;
      const double y = ((double)v - cy) * z / fy;
      if (use_radius && !(sqrt((x * x + y * y) + z * z) < r_max)) continue;
      pts[3 * m + 0] = x;
      pts[3 * m + 1] = y;
      pts[3 * m + 2] = z;
      cols[3 * m + 0] = bgr[3 * i + 2] / 255.0;
      cols[3 * m + 1] = bgr[3 * i + 1] / 255.0;
      cols[3 * m + 2] = bgr[3 * i + 0] / 255.0;
      m++;
    }
  return m;
}
