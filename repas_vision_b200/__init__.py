"""repas_vision_b200 -- B200-native (sm_100a) RGB-D -> point-cloud hot path of blanklavender/repas-vision.

Depth->colour registration, masked deprojection with stream compaction, 4x4 pose transform + merge,
hash-grid voxel downsampling and PLY records run as hand-written CUDA kernels in librepasvision.so behind
the C ABI of include/repas_vision.h; this package keeps the reference scripts' Python call shapes on top.
There is no CPU fallback: compute entry points raise RuntimeError without the library or a Blackwell GPU.
"""
from . import _lib
from .calibration import (Camera, load_camera, load_color_intrinsics, load_extrinsics, load_intrinsics,
                          load_intrinsics_json, load_transform_matrix, read_depth_to_color_extrinsics, scale_intrinsics)
from .cloud import (AxisAlignedBoundingBox, CloudBatch, KDTreeSearchParamHybrid, PointCloud, create_from_rgbd_image, create_masked_pointcloud, depth_to_meters,
                    deproject_batch, deproject_pixel_to_point, fuse_views, get_depth_at_pixel, median_depth_windows,
                    merge, nv12_to_bgr, register_depth_to_color)
from .ply import read_point_cloud, write_point_cloud
from . import registration
from .registration import (ICPConvergenceCriteria, RegistrationResult, TransformationEstimationPointToPlane,
                           TransformationEstimationPointToPoint, evaluate_registration, registration_icp)
from .pose import (invert_rigid, pose_from_tag_corners, solve_pnp_with_best_obj_order, to_4x4, world_from_camera)

__version__ = "0.1.0"


def library_path() -> str:
    return _lib.SO_PATH


def launch_count(device: int = 0) -> int:
    """Kernels launched by this package on `device` since its context was created."""
    return _lib.context(device).launches
