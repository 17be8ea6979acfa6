"""POD layouts of the C ABI (include/ses3d.h) as numpy dtypes and ctypes structures.

They mirror the reference's ROS messages (person_msgs/msg/*.msg): Keypoint2D (24 B),
Person2D (428 B), KeypointWithCovariance (80 B), PersonCov (1768 B), and the
FUSION_BODY_PARTS slot map (skeleton_3d/include/skeleton_3d/fusion_body_parts.h:4-25).
"""
import ctypes as C

import numpy as np

NUM_KEYPOINTS = 17
NUM_FUSION_KEYPOINTS = 21

keypoint2d_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("score", "<f4"), ("cov", "<f4", (3,))])
person2d_dtype = np.dtype([("score", "<f4"), ("keypoints", keypoint2d_dtype, (NUM_KEYPOINTS,)), ("bbox", "<f4", (4,))])
keypoint_cov_dtype = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("score", "<f4"), ("pad_", "<f4"),
                               ("cov", "<f8", (6,))])
person_cov_dtype = np.dtype([("id", "<u4"), ("score", "<f4"), ("keypoints", keypoint_cov_dtype, (NUM_FUSION_KEYPOINTS,)),
                             ("bbox_center", "<f8", (7,)), ("bbox_size", "<f8", (3,))])
camera_dtype = np.dtype([("T_cam_base", "<f8", (12,)), ("fx", "<f8"), ("fy", "<f8"), ("cx", "<f8"), ("cy", "<f8"),
                         ("Tx", "<f8"), ("Ty", "<f8"), ("width", "<u4"), ("height", "<u4")])

ellipsoid_dtype = np.dtype([("qw", "<f8"), ("qx", "<f8"), ("qy", "<f8"), ("qz", "<f8"), ("sx", "<f8"), ("sy", "<f8"),
                            ("sz", "<f8")])   # ses3d_ellipsoid: marker orientation + scale
MARKERS_SKELETON3D, MARKERS_POSE_PRIOR = 0, 1
MARKER_MAX_SEGMENTS = 22

assert keypoint2d_dtype.itemsize == 24
assert person2d_dtype.itemsize == 428
assert keypoint_cov_dtype.itemsize == 80
assert person_cov_dtype.itemsize == 1768
assert camera_dtype.itemsize == 152

# joint-order maps, S3D:139-145 == REP:47-53
KP2FUSION_SIMPLE = (0, 16, 15, 18, 17, 5, 2, 6, 3, 7, 4, 12, 9, 13, 10, 14, 11)
KP2FUSION_H36M = (0, 19, 1, 20, 8, 5, 2, 6, 3, 7, 4, 12, 9, 13, 10, 14, 11)

POSE_SIMPLE, POSE_H36M = 0, 1
PRECISION_FP32, PRECISION_FP64 = 0, 1
HOST_BUFFERS, DEVICE_BUFFERS = 0, 1

OK, E_INVALID, E_CUDA, E_CAPACITY, E_NOMEM = 0, -1, -2, -3, -4


class Params(C.Structure):
    """ses3d_params: the reference's parameters and constants (S3D:43-64, 149, 1095-1099)."""
    _fields_ = [("pose_method", C.c_int32), ("precision", C.c_int32), ("lm_refine", C.c_int32),
                ("lm_max_iters", C.c_int32), ("min_num_valid_keypoints", C.c_int32),
                ("triangulation_threshold", C.c_float), ("max_epipolar_error", C.c_double),
                ("reproj_error_max_acceptable", C.c_double), ("max_joint_dist_to_root", C.c_double),
                ("merge_dist_thresh", C.c_double), ("limb_cov_offset_sigma", C.c_double)]


def default_params(**overrides) -> Params:
    p = Params(POSE_SIMPLE, PRECISION_FP32, 0, 10, 9, 0.30, 0.050, 0.050, 2.0, 0.20, 0.075)
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(f"ses3d_params has no field {k!r}")
        setattr(p, k, v)
    return p


class AssocDump(C.Structure):
    _fields_ = [("hyp_of", C.c_void_p), ("n_hyp", C.c_void_p), ("n_hungarian", C.c_void_p)]


class SynthConfig(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_people", C.c_int32), ("p_max", C.c_int32), ("dropout", C.c_float),
                ("noise_px", C.c_float), ("area", C.c_float * 4), ("min_separation", C.c_float),
                ("min_visible", C.c_int32), ("frames_per_sequence", C.c_int32), ("step_m", C.c_float)]


def make_cameras(T_cam_base, fx=1000.0, fy=1000.0, cx=640.0, cy=360.0, Tx=0.0, Ty=0.0, width=1280, height=720):
    """Build a camera array from [n][3][4] (or [n][12]) base->camera transforms.

    Intrinsic defaults are the SURVEY 8(d) assumption (the reference ships no intrinsics; S3D:59 says f ~ 1000)."""
    T = np.asarray(T_cam_base, dtype=np.float64).reshape(-1, 12)
    cams = np.zeros(T.shape[0], dtype=camera_dtype)
    cams["T_cam_base"] = T
    for name, val in (("fx", fx), ("fy", fy), ("cx", cx), ("cy", cy), ("Tx", Tx), ("Ty", Ty)):
        cams[name] = val
    cams["width"] = width
    cams["height"] = height
    return cams


class PriorParams(C.Structure):
    """ses3d_prior_params: constants of pose_prior_mult_node.cpp (PRI:39-66) + gtsam's default LM parameters."""
    _fields_ = [("pose_method", C.c_int32), ("normalize_by_height", C.c_int32), ("min_num_obs_track", C.c_int32),
                ("lm_max_iterations", C.c_int32), ("min_score", C.c_float), ("pad_", C.c_float),
                ("pred_noise_sigma", C.c_double), ("default_res_sigma", C.c_double), ("avg_delay", C.c_double),
                ("root_sigma_factor", C.c_double), ("t_max_unobserved", C.c_double), ("dist_threshold", C.c_double),
                ("merge_dist_thresh", C.c_double), ("lm_lambda_initial", C.c_double), ("lm_lambda_factor", C.c_double),
                ("lm_lambda_upper_bound", C.c_double), ("lm_relative_error_tol", C.c_double),
                ("lm_absolute_error_tol", C.c_double), ("lm_min_model_fidelity", C.c_double)]


def default_prior_params(**overrides) -> PriorParams:
    p = PriorParams(POSE_SIMPLE, 0, 10, 100, 0.10, 0.0, 0.12, 0.10, 0.10, 100.0, 1.0, 5.0, 0.20, 1e-5, 10.0, 1e5, 1e-5,
                    1e-5, 1e-3)
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(f"ses3d_prior_params has no field {k!r}")
        setattr(p, k, v)
    return p
