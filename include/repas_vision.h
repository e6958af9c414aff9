/*
 * repas_vision.h -- C ABI of the B200-native RGB-D -> point-cloud hot path.
 *
 * The reference (blanklavender/repas-vision) is a folder of Python scripts with
 * no FFI of its own; the boundary it offers is a set of Python call shapes.  Each
 * entry point below names the reference call site(s) it replaces (paths are
 * relative to the reference checkout).  The Python host in
 * repas_vision_b200/ binds this header with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer named d_* is DEVICE memory owned by the caller (PyTorch's
 *    caching allocator in this repo); the library allocates nothing on the hot
 *    path and keeps no pointer after the call returns;
 *  - every call is asynchronous and ordered on `stream` (a cudaStream_t passed
 *    as void*); results are ready when the stream reaches that point;
 *  - every call returns an int status (RV_OK == 0).  rv_last_error(ctx) holds
 *    the text of the last failure on that context.  No exception crosses the ABI;
 *  - clouds are SoA: six planes x,y,z,r,g,b of `plane_stride` elements each,
 *    element type float32 or float64 (RvDType);
 *  - images are row-major, pixel (u = column, v = row), pixel centres at
 *    integer coordinates (create_masked_ply.py:77,94-95).
 */
#ifndef REPAS_VISION_H
#define REPAS_VISION_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RV_ABI_VERSION 2

typedef struct rv_ctx rv_ctx;
typedef void *rv_stream; /* cudaStream_t */

/* ---- status codes ------------------------------------------------------- */
enum RvStatus {
  RV_OK = 0,
  RV_EINVAL = 1,    /* bad shape / dtype / enum / null pointer               */
  RV_ECAPACITY = 2, /* an output capacity was too small (counts still valid) */
  RV_ECUDA = 3,     /* a CUDA runtime call failed; see rv_last_error          */
  RV_EALIGN = 4,    /* pointer or stride violates the stated alignment        */
  RV_EWORKSPACE = 5 /* workspace smaller than rv_*_workspace_bytes said       */
};

/* ---- element types ------------------------------------------------------ */
enum RvDType { RV_F32 = 0, RV_F64 = 1 };

/* ---- camera model --------------------------------------------------------
 * Intrinsics JSON on disk: {fx,fy,cx,cy,width,height[,dist_coeffs]}
 * (create_masked_ply.py:27-43, april_tag_detector_solvepnp.py:51-66) or the
 * RealSense form {fx,fy,ppx,ppy,coeffs,distortion_model}
 * (vis_tool_april_tag_pose_validaiton.py:38-47).  dist = (k1,k2,p1,p2,k3).   */
enum RvDistortion {
  RV_DIST_NONE = 0,
  RV_DIST_BROWN_CONRADY = 1,         /* forward model; deprojection inverts it iteratively */
  RV_DIST_INVERSE_BROWN_CONRADY = 2, /* closed form on deprojection                        */
  RV_DIST_MODIFIED_BROWN_CONRADY = 3 /* forward model on projection only                   */
};

typedef struct RvCam {
  double fx, fy, cx, cy;
  double dist[5];
  int32_t model; /* RvDistortion */
  int32_t width;
  int32_t height;
  int32_t reserved;
} RvCam;

/* ---- depth unit rules --------------------------------
 * MUL_F32 : f32(u16) * f32(scale)    better_three_capture.py:118-125, april_tag_bg_removal_pl.py:286-288
 * DIV_F32 : f32(u16) / f32(scale)    custom_reader.py:39-40, Open3D RGBDImage (depth_scale = 1000)
 * DIV_F64 : f64(u16) / scale         canopy_return.py:312-315
 * `scale` is 0.001 for MUL_F32 and 1000.0 for the DIV rules in the reference. */
enum RvUnitRule { RV_UNIT_MUL_F32 = 0, RV_UNIT_DIV_F32 = 1, RV_UNIT_DIV_F64 = 2 };

enum RvDepthKind { RV_DEPTH_U16 = 0, RV_DEPTH_F32_METERS = 1 };

/* ---- output modes of the deprojection kernel ------------------------------ */
enum RvDeprojectMode {
  RV_MODE_COMPACT_ORDERED = 0,   /* row-major order of valid pixels, == numpy `[valid]` (create_masked_ply.py:89-100) */
  RV_MODE_COMPACT_UNORDERED = 1, /* any order of tiles (generic kernel: arrival order; TMA kernel: row-major like ORDERED)    */
  RV_MODE_DENSE_ZERO = 2,        /* one record per pixel, zeros where invalid (rs.pointcloud / Orbbec RGB_POINT)     */
  RV_MODE_DENSE_NAN = 3,         /* one record per pixel, NaN xyz where invalid (Open3D project_valid_depth_only=False) */
  RV_MODE_COMPACT_PACKED = 4     /* COMPACT_ORDERED with the frames of the batch back to back: frame b occupies
                                    [offsets[b], offsets[b+1]) of every plane; d_counts receives the B+1 offsets and
                                    frame_stride is ignored (one contiguous device->host copy per plane)            */
};

enum RvColorScale {
  RV_COLOR_UNIT = 0,   /* u8 / 255.0 in [0,1]   (create_masked_ply.py:100)             */
  RV_COLOR_255 = 1,    /* raw 0..255 as floats (Orbbec RGB_POINT, better_three_capture.py:235-237) */
  RV_COLOR_PACKED8 = 2 /* the colour bytes themselves: ONE plane (plane 3) of 32-bit words holding the bytes r,g,b,0 in
                          memory order (what a PLY `uchar red green blue` record stores, create_masked_ply.py:177), so a
                          point costs 16 instead of 24 bytes on its way to the host; float32 output only; the host expands
                          k / 255.0 lazily */
};

/* which arithmetic produces x and y */
enum RvGeometry {
  RV_GEOM_REFERENCE_F64 = 0, /* x = (u - cx) * z / fx in float64, multiply then divide, rounded once to the output type:
                                the reference's numpy statement (create_masked_ply.py:94-95) */
  RV_GEOM_SDK_F32 = 1        /* x = z * ((u - ppx) / fx) in float32 throughout: rs2_deproject_pixel_to_point as rs.pointcloud
                                (capture_aligned_all.py:209-216) and Orbbec's PointCloudFilter (better_three_capture.py:235-237)
                                evaluate it; pinhole cameras (RV_DIST_NONE) only; served by the generic kernel.  Differs from
                                the reference form by at most one float32 ulp on the captured frames */
};

/* layout of the colour image handed to rv_deproject_mask */
enum RvColorFormat {
  RV_COLORFMT_BGR8 = 0, /* [B,H,W,3] uint8, cv2 / SDK bgr8 */
  RV_COLORFMT_NV12 = 1  /* [B,H*3/2,W] uint8 as the camera delivers it (better_three_capture.py:101-106,159): H rows of
                           luma, H/2 rows of interleaved U,V; converted per kept pixel with the integer arithmetic of
                           cv2.cvtColor(COLOR_YUV2BGR_NV12), see rv_nv12_to_bgr.  H and W must be even */
};

/* which K1 kernel runs: AUTO picks the TMA-fed pipeline when H*W % 16 == 0, W >= 32, the inputs are 16-byte
 * aligned; otherwise the generic kernel.  Both produce identical bytes (tests/test_gpu_deproject.py runs every case
 * through both); a COMPACT_UNORDERED request on the TMA path gets the ordered result, which satisfies it. */
enum RvKernelSelect { RV_KERNEL_AUTO = 0, RV_KERNEL_GENERIC = 1, RV_KERNEL_TMA = 2 };

typedef struct RvDeprojectParams {
  RvCam cam;            /* intrinsics of the grid the depth lives on (the colour camera after alignment) */
  int32_t depth_kind;   /* RvDepthKind */
  int32_t unit_rule;    /* RvUnitRule, used when depth_kind == RV_DEPTH_U16 */
  double unit_scale;    /* 0.001 or 1000.0, see RvUnitRule */
  int32_t use_seg_mask; /* 1: d_mask is read; keep mask>0 (or mask==0 when invert_mask) create_masked_ply.py:79-83 */
  int32_t invert_mask;
  int32_t use_depth_trunc; /* 1: z32 >= (float)depth_trunc -> invalid (Open3D create_from_color_and_depth) */
  int32_t use_zclip;       /* 1: keep z_min <= Z <= z_max, inclusive (view_point_cloud.py:109-116) */
  double depth_trunc;
  double z_min, z_max;
  int32_t use_radius; /* 1: keep ||p||_2 < r_max, strict (distance_masking_on_ply.py:12-16) */
  int32_t use_aabb;   /* 1: keep aabb_min <= p <= aabb_max per axis, inclusive (april_tag_bg_removal_pl.py:450-455) */
  double r_max;
  double aabb_min[3];
  double aabb_max[3];
  int32_t mode;        /* RvDeprojectMode */
  int32_t out_dtype;   /* RvDType of the six output planes */
  int32_t color_scale; /* RvColorScale */
  int32_t kernel_select; /* RvKernelSelect */
  int32_t color_format;  /* RvColorFormat of d_bgr */
  int32_t geometry;      /* RvGeometry */
} RvDeprojectParams;

/* ---- context -------------------------------------------------------------- */
int rv_abi_version(void);
/* sizeof(RvCam) / sizeof(RvDeprojectParams) as compiled, so a foreign-language binding can verify its layout */
int rv_sizeof_cam(void);
int rv_sizeof_deproject_params(void);
const char *rv_build_info(void); /* "sm_100a nvcc 12.9 ..." */
int rv_create(int device, rv_ctx **out_ctx);
int rv_destroy(rv_ctx *ctx);
const char *rv_last_error(const rv_ctx *ctx);
const char *rv_status_string(int status);
int rv_sm_count(const rv_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t rv_launch_count(const rv_ctx *ctx);

/* ---- a1/a2: depth units -----------------------------------------------------
 * replaces depth_to_meters (better_three_capture.py:118-125) and the inline
 * `astype(float32)/1000.0` conversions (custom_reader.py:39-40).
 * d_depth: n uint16 ; d_out: n float32 (DIV_F64 results are rounded to f32). */
int rv_depth_to_meters(rv_ctx *ctx, const uint16_t *d_depth, int64_t n, int unit_rule, double unit_scale,
                       float *d_out, rv_stream stream);

/* ---- a6: depth -> colour registration ---------------------------------------
 * replaces AlignFilter(align_to_stream=COLOR_STREAM).process
 * (better_three_capture.py:169,188; april_tag_detector_ToF.py:169,183) and
 * rs.align(rs.stream.color).process (capture_aligned_all.py:75,197;
 * canopy_return.py:442,453).  Semantics: SURVEY Appendix B.3 (librealsense
 * align z16 -> other): each depth pixel's two half-pixel corners are
 * deprojected, moved by (R,t), projected into the colour camera and rounded;
 * the covered rectangle takes the minimum raw depth (z-buffer).  float32
 * geometry, no fused multiply-add.
 *  d_depth      [B,Hd,Wd] uint16
 *  R (9, COLUMN-major as librealsense stores it) and t (3): host pointers
 *  d_out        [B,Hc,Wc] uint16, 0 where no depth pixel lands
 *  d_winner     [B,Hc,Wc] int32 source index (y*Wd+x) of the winning depth pixel,
 *               lowest index on ties, -1 where empty; may be NULL
 *  d_ws         workspace of rv_register_workspace_bytes(B,Hd,Wd,Hc,Wc) bytes, 16-B aligned: per frame of a
 *               16-frame chunk the 8-byte colour rectangle of every depth pixel, the bounding box of every
 *               32-pixel row segment and one fixed-capacity segment list per 64x32 colour tile (a workspace
 *               that holds at least one frame's share is accepted; the batch is then walked in smaller chunks) */
size_t rv_register_workspace_bytes(int B, int Hd, int Wd, int Hc, int Wc);
int rv_register_depth_to_color(rv_ctx *ctx, const uint16_t *d_depth, int B, const RvCam *depth_cam,
                               const RvCam *color_cam, const float *R_colmajor, const float *t, float depth_units,
                               uint16_t *d_out, int32_t *d_winner, void *d_ws, size_t ws_bytes, rv_stream stream);

/* ---- distortion ray table -----------------------------------------------------
 * Per-pixel normalised ray (xn, yn) of a distorted camera, so the deprojection
 * kernel multiplies instead of iterating (rs2_deproject_pixel_to_point semantics,
 * SURVEY Appendix B.3, in float64).  d_table: [H,W,2] float64. */
int rv_build_ray_table(rv_ctx *ctx, const RvCam *cam, double *d_table, rv_stream stream);

/* ---- a3/a5/a7/a8/a9: deprojection + masks + stream compaction -------------------
 * replaces create_masked_pointcloud (create_masked_ply.py:56-107), the SDK
 * clouds PointCloudFilter.process (better_three_capture.py:235-237) and
 * rs.pointcloud.calculate (capture_aligned_all.py:209-216) in the dense modes,
 * and fuses the cloud predicates of distance_masking_on_ply.py:12-19,
 * view_point_cloud.py:109-116 and april_tag_bg_removal_pl.py:450-455.
 *  d_depth   [B,H,W] uint16 or float32 (params->depth_kind)
 *  d_bgr     [B,H,W,3] uint8, BGR order (cv2 / SDK bgr8), or [B,H*3/2,W] NV12 when params->color_format says so;
 *            may be NULL (no colour planes written)
 *  d_mask    [B,H,W] uint8 segmentation mask, or NULL
 *  d_ray_table  [H,W,2] float64 from rv_build_ray_table, required iff cam.model != RV_DIST_NONE
 *  d_out     six planes (x,y,z,r,g,b) -- four (x,y,z,packed rgb) with RV_COLOR_PACKED8 --, plane p of frame b starts at element
 *            p*plane_stride + b*frame_stride; compact modes write counts[b] elements per
 *            plane, dense modes H*W.  frame_stride is the per-frame capacity.
 *  d_valid   [B,H,W] uint8 (1 = kept) or NULL
 *  d_src_index  same addressing as one plane, int32 source pixel index of each output point, or NULL
 *  d_counts  [B] int64 kept points per frame (written in every mode); [B+1] exclusive offsets in COMPACT_PACKED
 *  d_ws      rv_deproject_workspace_bytes(B,H,W) bytes
 * Points beyond frame_stride are not written; counts[b] still reports the true
 * number so the caller can detect RV_ECAPACITY after synchronising. */
size_t rv_deproject_workspace_bytes(int B, int H, int W);
int rv_deproject_mask(rv_ctx *ctx, const void *d_depth, const uint8_t *d_bgr, const uint8_t *d_mask,
                      const double *d_ray_table, int B, int H, int W, const RvDeprojectParams *params, void *d_out,
                      int64_t plane_stride, int64_t frame_stride, uint8_t *d_valid, int32_t *d_src_index,
                      int64_t *d_counts, void *d_ws, size_t ws_bytes, rv_stream stream);

/* ---- a8/a9 on an existing cloud -------------------------------------------------
 * the stand-alone form of the predicates above (distance_masking_on_ply.py,
 * view_point_cloud.py --z-min/--z-max, AABB crop): ordered compaction of an
 * n-point SoA cloud.  Only use_zclip/use_radius/use_aabb and their values are read
 * from params.  d_count: one int64.  d_index: NULL or [n] int64, the source index of every kept
 * point (what select_by_index takes; used to carry normals along).  The kept count is only
 * known on the device, so out_plane_stride must be >= n (RV_ECAPACITY otherwise); the same
 * holds for rv_select_by_mask. */
size_t rv_filter_workspace_bytes(int64_t n);
int rv_filter_cloud(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int dtype, int has_color,
                    const RvDeprojectParams *params, void *d_out, int64_t out_plane_stride, int64_t *d_count,
                    int64_t *d_index, void *d_ws, size_t ws_bytes, rv_stream stream);

/* ---- a11: 4x4 pose transform + merge ------------------------------------------
 * replaces geometry.transform(T) (final_view_with_cad.py:333,
 * vis_tool_april_tag_pose_validaiton.py:239-245) and PointCloud `+`
 * (concatenation) for the four_pose_captures fusion.
 * For view v: p' = (T_v [p,1])[:3] / (T_v [p,1])[3] in float64, colours copied.
 *  T            host, n_views x 16 doubles, row-major 4x4 (np.loadtxt layout, 6dof_icp_export.py:55-70)
 *  d_in[v]      host array of device pointers to each view's six-plane cloud
 *  n[v], in_plane_stride[v]  host arrays
 *  d_out        merged cloud, view v starts at sum(n[0..v-1])
 *  d_bounds     6 doubles (min xyz, max xyz) of the merged cloud, or NULL.  Must be
 *               pre-initialised by rv_bounds_init when several calls accumulate. */
int rv_bounds_init(rv_ctx *ctx, double *d_bounds, rv_stream stream);
/* get_min_bound / get_max_bound / get_center of a cloud (april_tag_bg_removal_pl.py:425-431 reads the bounds of every
 * cloud it crops): d_stats receives 9 doubles, min xyz, max xyz, sum xyz (centre = sum / n), all in float64. */
int rv_cloud_stats(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int dtype, double *d_stats,
                   rv_stream stream);
int rv_transform_merge(rv_ctx *ctx, int n_views, const void *const *d_in, const int64_t *in_plane_stride,
                       const int64_t *n, const double *T, int in_dtype, int has_color, void *d_out,
                       int64_t out_plane_stride, int out_dtype, double *d_bounds, rv_stream stream);

/* ---- a12: voxel grid downsampling -----------------------------------------------
 * replaces PointCloud.voxel_down_sample(voxel_size) (17 call sites, e.g.
 * mpa_icp_export.py:44,174; create_masked_ply.py:164; view_point_cloud.py:120).
 * Semantics: SURVEY Appendix B.1 -- origin = min_bound - voxel/2, key =
 * floor((p - origin)/voxel) per axis in float64, output = per-voxel mean of
 * points and colours, order unspecified.
 *  d_in         six-plane cloud (three planes when has_color == 0)
 *  d_bounds     6 doubles min/max of the cloud, or NULL (computed internally)
 *  d_out        up to out_capacity voxels, six planes
 *  d_keys       [3, out_capacity] int32 voxel indices, or NULL
 *  d_counts_out [out_capacity] int32 points per voxel, or NULL
 *  d_m          one int64: number of voxels (true number even if > out_capacity)
 *  d_ws         rv_voxel_workspace_bytes(n) bytes, 64-B aligned: 1.5 n sixteen-byte table slots (hash key + chain head)
 *               and one bit per slot,
 *               one 8-byte list entry and one 4-byte slot index per point, one bit per point of run heads, a small pool
 *               for voxels with very long chains (about 44 bytes per point) and room for a fusion's transformed
 *               coordinates (24 bytes per point, untouched by this call); the table is initialised by the call
 * Each voxel index must fit 21 bits (extent/voxel < 2^21); otherwise d_m is set to -1.  n < 2^31.
 * Voxels made of at most eight runs of consecutive points (practically all of a 5 mm grid over camera clouds) are summed
 * point by point in index order in float64: their means equal the sequential Open3D loop bit for bit; voxels of more runs
 * are summed by float64 atomics (1e-15 relative). */
size_t rv_voxel_workspace_bytes(int64_t n);
int rv_voxel_downsample(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int in_dtype,
                        int has_color, double voxel_size, const double *d_bounds, void *d_out,
                        int64_t out_plane_stride, int out_dtype, int64_t out_capacity, int32_t *d_keys,
                        int32_t *d_counts_out, int64_t *d_m, void *d_ws, size_t ws_bytes, rv_stream stream);

/* ---- a11 + a12 in one call: the four_pose_captures fusion -------------------------
 * pcd_i.transform(T_i) for every view, `+`, voxel_down_sample(voxel_size) (SURVEY Appendix D.4;
 * final_view_with_cad.py:333, mpa_icp_export.py:174) without building the merged six-plane cloud:
 * the views are read where they lie; the pass that finds the merged cloud's bounds stores
 * p' = T_v p (rounded to in_dtype exactly as rv_transform_merge would store it) in the
 * workspace, the colours are read from the views when a voxel is summed, and the voxel grid
 * (origin = min bound of the merged cloud - voxel/2) is built from those values.
 * Same results as rv_transform_merge followed by
 * rv_voxel_downsample.  Up to 8 views; arguments as in those two calls; workspace
 * rv_voxel_workspace_bytes(sum of n). */
int rv_fuse_voxel(rv_ctx *ctx, int n_views, const void *const *d_in, const int64_t *in_plane_stride, const int64_t *n,
                  const double *T, int in_dtype, int has_color, double voxel_size, void *d_out, int64_t out_plane_stride,
                  int out_dtype, int64_t out_capacity, int32_t *d_keys, int32_t *d_counts_out, int64_t *d_m, void *d_ws,
                  size_t ws_bytes, rv_stream stream);

/* ---- a13 (next): PLY vertex records on device -----------------------
 * packs an SoA cloud into the binary little-endian vertex records that
 * o3d.io.write_point_cloud emits (create_masked_ply.py:177): xyz as float32 or
 * float64 followed by uchar red,green,blue = round(clamp(c,0,1)*255).
 * d_records: n * (3*sizeof(coord) + 3) bytes. */
int rv_pack_ply_records(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int in_dtype,
                        int has_color, int color_scale, int coord_dtype, uint8_t *d_records, rv_stream stream);

/* the inverse, for o3d.io.read_point_cloud (view_point_cloud.py:104 and 30 more call sites): the vertex element of a
 * binary little-endian PLY, uploaded as it lies in the file, unpacked on the device into SoA planes.
 *  d_records     n records of record_bytes each (any other vertex properties are skipped)
 *  xyz_offset    byte offsets of x, y, z inside a record; coord_dtype RV_F32 / RV_F64 as stored in the file
 *  rgb_offset    byte offsets of the uchar red, green, blue, or NULL: colours become k / 255.0 (Open3D's conversion)
 *  d_out         three or six planes of out_dtype, out_plane_stride elements apart */
int rv_unpack_ply_records(rv_ctx *ctx, const uint8_t *d_records, int64_t n, int record_bytes, const int32_t *xyz_offset,
                          int coord_dtype, const int32_t *rgb_offset, void *d_out, int64_t out_plane_stride,
                          int out_dtype, rv_stream stream);

/* ---- next, SURVEY 8f-4: statistical outlier removal ------------------------------
 * replaces pcd.remove_statistical_outlier(nb_neighbors=20, std_ratio=2.0) (create_masked_ply.py:168-170); semantics of
 * Open3D 0.19 PointCloud::RemoveStatisticalOutliers: avg[i] = mean of sqrt(squared distance) over the k nearest
 * neighbours of point i INCLUDING itself (float64), cloud_mean / std_dev over the avg > 0 summed in index order with
 * Bessel's correction, keep i <=> 0 < avg[i] < cloud_mean + std_ratio * std_dev.
 *  rv_knn_mean_distance         exact k-nearest search on a hash grid (1 <= k <= 64); d_mean [n] float64
 *  rv_statistical_outlier_mask  d_keep [n] uint8; d_stats 520 doubles of scratch whose first four are the results:
 *                               cloud_mean, std_dev, threshold, points counted (parallel sums; redone in index order
 *                               only when a point sits within rounding distance of the threshold, so d_keep always
 *                               equals the sequential definition)
 *  rv_select_by_mask            ordered compaction of the cloud by a uint8 mask; d_index (NULL or [n] int64) receives the
 *                               source index of every kept point (Open3D's `ind`); workspace rv_filter_workspace_bytes(n)
 *  rv_estimate_normals          pcd.estimate_normals(KDTreeSearchParamHybrid(radius, max_nn)) (create_masked_ply.py:173) and,
 *                               with camera_location != NULL, pcd.orient_normals_towards_camera_location (:174): per point the
 *                               up to max_nn nearest neighbours closer than radius, covariance from the cumulants in
 *                               neighbour order, unit eigenvector of the smallest eigenvalue (Open3D's non-iterative
 *                               FastEigen3x3), (0,0,1) with fewer than three neighbours; d_normals three float64 planes;
 *                               workspace rv_knn_workspace_bytes(n)
 *  rv_orient_normals            the stand-alone orient_normals_towards_camera_location */
size_t rv_knn_workspace_bytes(int64_t n);
int rv_orient_normals(rv_ctx *ctx, const void *d_xyz, int64_t plane_stride, int64_t n, int dtype, double *d_normals,
                      int64_t normal_stride, const double *camera_location, rv_stream stream);
int rv_estimate_normals(rv_ctx *ctx, const void *d_xyz, int64_t plane_stride, int64_t n, int dtype, double radius, int max_nn,
                        const double *camera_location, double *d_normals, int64_t normal_stride, void *d_ws, size_t ws_bytes,
                        rv_stream stream);
int rv_knn_mean_distance(rv_ctx *ctx, const void *d_xyz, int64_t plane_stride, int64_t n, int dtype, int k, double *d_mean,
                         void *d_ws, size_t ws_bytes, rv_stream stream);
int rv_statistical_outlier_mask(rv_ctx *ctx, const double *d_mean, int64_t n, double std_ratio, uint8_t *d_keep,
                                double *d_stats, rv_stream stream);
int rv_select_by_mask(rv_ctx *ctx, const void *d_in, int64_t in_plane_stride, int64_t n, int dtype, int has_color,
                      const uint8_t *d_keep, void *d_out, int64_t out_plane_stride, int64_t *d_count, int64_t *d_index,
                      void *d_ws, size_t ws_bytes, rv_stream stream);

/* ---- next, SURVEY 8f-4: ICP correspondence search ------------------------------------
 * replaces the per-iteration work of o3d.pipelines.registration.registration_icp(source, target, max_dist, init,
 * TransformationEstimationPointToPlane(), ICPConvergenceCriteria(...)) (mpa_icp_export.py:187-197, 6dof_icp_export.py:134-144,
 * mpa_icp.py:159, icp_cad_model.py:90-94,271-275); semantics of Open3D 0.19 Registration.cpp
 * GetRegistrationResultAndCorrespondences + TransformationEstimationPointToPlane / PointToPoint::ComputeTransformation:
 *  rv_nn_index_build   hash-grid index of the target cloud (the KD-tree of the reference), built once per registration;
 *                      workspace rv_knn_workspace_bytes(n), which then IS the index
 *  rv_nn_search        d_nearest[i] = index of the target point nearest to query i with squared distance strictly below
 *                      max_distance^2 (KDTreeFlann::SearchHybrid(p, max_distance, 1)), -1 without one; equal distances
 *                      go to the lower index
 *  rv_icp_sums         d_sums (rv_icp_sums_bytes() bytes; the first 32 doubles are the result, summed per block in a
 *                      fixed order): [0] correspondences, [1] sum of squared distances;
 *                      point_to_plane: [2] sum r^2, [3..8] J^T r, [9..29] upper triangle of J^T J (row-major), with
 *                      r = (s - t) . n_t and J = [s x n_t, n_t];
 *                      point-to-point: [2] sum |s|^2, [3..5] sum s, [6..8] sum t, [9..17] sum t_a s_b (a row-major)
 * With these three calls the 6x6 solve / Umeyama step and the convergence test run on the host
 * (repas_vision_b200/registration.py: the point-to-point estimation and `device_loop=False`); rv_icp_iterate below keeps
 * the point-to-plane loop on the device. */
int rv_nn_index_build(rv_ctx *ctx, const void *d_xyz, int64_t plane_stride, int64_t n, int dtype, double max_distance,
                      void *d_ws, size_t ws_bytes, rv_stream stream);
int rv_nn_search(rv_ctx *ctx, const void *d_index_ws, size_t ws_bytes, int64_t n_indexed, const void *d_query_xyz,
                 int64_t query_stride, int64_t n_query, int dtype, double max_distance, int32_t *d_nearest, rv_stream stream);
size_t rv_icp_sums_bytes(void);
int rv_icp_sums(rv_ctx *ctx, int point_to_plane, const void *d_source_xyz, int64_t source_stride, int64_t n_source,
                int source_dtype, const void *d_target_xyz, int64_t target_stride, int64_t n_target, int target_dtype,
                const double *d_target_normals, int64_t normal_stride, const int32_t *d_nearest, double *d_sums,
                rv_stream stream);

/* the whole point-to-plane loop on the device: rv_icp_begin writes the state (accumulated transformation = T_init, the
 * stopping rule's thresholds), rv_icp_iterate queues `steps` iterations -- each: pcd.Transform(update) on the caller's
 * working copy of the source coordinates IN PLACE (done by the search kernel on its way in), nearest search (bounded from
 * the start by the previous evaluation's match, which d_nearest must therefore still hold when `first` is 0; the result
 * is the same), sums, then ONE thread's worth of host logic on the device: fitness / inlier RMSE, |delta fitness| <
 * relative_fitness && |delta rmse| < relative_rmse, the 6 x 6 solve (L D L^T in registers; elimination with partial
 * pivoting when a pivot is not positive, and a vanishing pivot there leaves the update at identity) and transformation =
 * update * transformation -- preceded, when `first` is set, by the evaluation of the initial alignment.  Once the loop has
 * stopped every queued kernel returns at once, so the host reads the state back once per few iterations instead of
 * once per iteration.  The state (rv_icp_state_bytes() = 448 bytes): double T[16], update[16], fitness, inlier_rmse,
 * relative_fitness, relative_rmse, iterations, max_iteration, n_source, evaluated, 8 reserved; then int32 done.
 * d_nearest holds the correspondences of the last evaluation; d_sums as in rv_icp_sums. */
size_t rv_icp_state_bytes(void);
int rv_icp_begin(rv_ctx *ctx, void *d_state, const double *T_init, int max_iteration, double relative_fitness,
                 double relative_rmse, int64_t n_source, rv_stream stream);
int rv_icp_iterate(rv_ctx *ctx, void *d_state, int first, int steps, void *d_work_xyz, int64_t work_stride, int64_t n_source,
                   int work_dtype, const void *d_index_ws, size_t ws_bytes, int64_t n_target, const void *d_target_xyz,
                   int64_t target_stride, int target_dtype, const double *d_target_normals, int64_t normal_stride,
                   double max_distance, int32_t *d_nearest, double *d_sums, rv_stream stream);

/* ---- a2 (next): windowed median depth ------------------------------
 * replaces get_depth_at_pixel (canopy_return.py:279-317) and median_depth
 * (final_view.py:132-141): median of the non-zero raw depths in a win x win
 * window clipped to the image; numpy median (mean of the two middle values for
 * even counts); result in raw units as float64, NaN when the window holds no
 * valid depth.
 *  d_depth [H,W] uint16 ; d_uv [n,2] int32 (x,y) ; d_out [n] float64 */
int rv_median_depth_window(rv_ctx *ctx, const uint16_t *d_depth, int H, int W, const int32_t *d_uv, int64_t n,
                           int window, double *d_out, rv_stream stream);

/* ---- (next): NV12 -> BGR ---------------------------------------------
 * replaces cv2.cvtColor(nv12, COLOR_YUV2BGR_NV12) in frame_to_bgr_image
 * (better_three_capture.py:101-106).  d_nv12: [B, H*3/2, W] uint8; d_bgr: [B,H,W,3]. */
int rv_nv12_to_bgr(rv_ctx *ctx, const uint8_t *d_nv12, int B, int H, int W, uint8_t *d_bgr, rv_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* REPAS_VISION_H */
