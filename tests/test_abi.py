"""CPU tests of the drop-in boundary: libses3d.so loads and exports every symbol include/ses3d.h declares, POD
layouts have the person_msgs sizes, defaults equal the reference constants, the product fails loudly without a
GPU, and the synthetic generator is deterministic. No compute calls are made on the library here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from smartedgesensor3dhumanpose_b200 import lib as L
from smartedgesensor3dhumanpose_b200 import layouts, rigs, synth

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "ses3d.h").read_text()
    declared = set(re.findall(r"\b(ses3d_[a-z0-9_]+)\s*\(", header))
    declared -= {"ses3d_handle_s"}
    assert declared, "no declarations parsed"
    lib = L.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"libses3d.so does not export {missing}"
    assert set(L.EXPORTS) <= declared
    assert b"sm_100a" in lib.ses3d_version()


def test_pod_layouts_match_person_msgs():
    assert layouts.keypoint2d_dtype.itemsize == 24        # Keypoint2D.msg: 3 + 3 float32
    assert layouts.person2d_dtype.itemsize == 428         # 4 + 17*24 + 16
    assert layouts.keypoint_cov_dtype.itemsize == 80      # Point(24) + score(4) + pad(4) + 6 float64
    assert layouts.person_cov_dtype.itemsize == 1768      # 8 + 21*80 + 56 + 24
    assert layouts.person_cov_dtype.fields["keypoints"][1] == 8
    assert layouts.person2d_dtype.fields["bbox"][1] == 412
    assert C.sizeof(layouts.Params) == 64 and C.sizeof(layouts.SynthConfig) == 56
    assert layouts.KP2FUSION_SIMPLE == (0, 16, 15, 18, 17, 5, 2, 6, 3, 7, 4, 12, 9, 13, 10, 14, 11)   # S3D:139-142
    assert layouts.KP2FUSION_H36M == (0, 19, 1, 20, 8, 5, 2, 6, 3, 7, 4, 12, 9, 13, 10, 14, 11)       # S3D:143-145


def test_default_params_are_the_reference_constants():
    p = layouts.Params()
    L.load().ses3d_default_params(C.byref(p))
    q = layouts.default_params()
    for name, _ in layouts.Params._fields_:
        assert getattr(p, name) == getattr(q, name), name
    assert p.min_num_valid_keypoints == 9 and abs(p.triangulation_threshold - 0.30) < 1e-7      # S3D:57-58
    assert (p.max_epipolar_error, p.reproj_error_max_acceptable) == (0.050, 0.050)              # S3D:59-60
    assert (p.max_joint_dist_to_root, p.merge_dist_thresh, p.limb_cov_offset_sigma) == (2.0, 0.20, 0.075)
    assert p.lm_refine == 0 and p.precision == layouts.PRECISION_FP32 and p.pose_method == layouts.POSE_SIMPLE


def test_invalid_arguments_are_rejected_without_touching_the_gpu():
    lib = L.load()
    h = C.c_void_p()
    cams = rigs.ring4()
    assert lib.ses3d_create(1, cams.ctypes.data, None, 0, C.byref(h)) == layouts.E_INVALID       # < 2 cameras (S3D:1133)
    assert b"camera" in lib.ses3d_last_error_string()
    assert lib.ses3d_create(4, None, None, 0, C.byref(h)) == layouts.E_INVALID
    assert lib.ses3d_triangulate_batch(None, 1, 1, None, None, 1, None, None, None, 0, None) == layouts.E_INVALID
    assert lib.ses3d_destroy(None) == 0


def test_product_fails_loudly_without_a_gpu():
    """No CPU fallback: on a machine without CUDA the constructor must raise, not silently compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from smartedgesensor3dhumanpose_b200 import api
    with pytest.raises(L.Ses3dError) as ei:
        api.GeometryPipeline(rigs.ring4())
    assert ei.value.code == layouts.E_CUDA and "no CPU path" in str(ei.value)


def test_product_never_imports_the_oracle():
    """The oracle and the serial test build are checkers: no product source may import, include or load them."""
    pkg = ROOT / "smartedgesensor3dhumanpose_b200"
    for src in pkg.glob("*.py"):
        for line in src.read_text().splitlines():
            code = line.split("#")[0]
            assert not re.search(r"\b(import|from)\s+(oracle|tests)\b", code), f"{src.name}: {line}"
            assert "libses3d_oracle" not in code and "libhostsim" not in code, f"{src.name}: {line}"
    for src in (pkg / "csrc").iterdir():
        if src.is_file():
            for line in src.read_text(errors="ignore").splitlines():
                if line.lstrip().startswith("#include"):
                    assert "oracle" not in line and "hostsim" not in line, f"{src.name}: {line}"


def test_synthetic_generator_is_deterministic_and_counter_based():
    cams = rigs.ring8()
    cfg = synth.synth_config(seed=5, n_people=4, dropout=0.05, area=rigs.AREAS["ring8"])
    a = synth.synth_frames(cams, cfg, 64)
    b = synth.synth_frames(cams, cfg, 64)
    assert a["persons"].tobytes() == b["persons"].tobytes()
    # counter based: frames [32,64) generated alone are identical to the tail of [0,64)
    c = synth.synth_frames(cams, cfg, 32, first_frame=32)
    assert c["persons"].tobytes() == a["persons"][32:].tobytes() and np.array_equal(c["n_persons"], a["n_persons"][32:])
    # sanity of the content
    kp = a["persons"]["keypoints"]
    live = np.arange(4)[None, None, :] < a["n_persons"][:, :, None]
    assert a["n_persons"].max() <= 4 and live.any()
    assert ((kp["score"][live] >= 0) & (kp["score"][live] <= 1)).all()
    assert np.all(a["gt_id"][live] >= 0) and np.all(a["gt_id"][~live] == -1)
    other = synth.synth_frames(cams, synth.synth_config(seed=6, n_people=4, dropout=0.05, area=rigs.AREAS["ring8"]), 64)
    assert other["persons"].tobytes() != a["persons"].tobytes()


def test_rig_fixture_matches_the_reference_launch_file():
    cams = rigs.hall16()
    T = cams["T_cam_base"].reshape(16, 3, 4)
    centres = np.array([-t[:, :3].T @ t[:, 3] for t in T])
    assert np.allclose(centres.min(0), [-11.79, -6.89, 2.254], atol=0.01)      # SURVEY 4 fixture recipe
    assert np.allclose(centres.max(0), [4.556, 6.667, 2.800], atol=0.01)
    for t in T:
        assert np.allclose(t[:, :3] @ t[:, :3].T, np.eye(3), atol=1e-9)
    # cam_1: base -> cam_1 translation/quaternion straight from cameras_extrinsics.launch:2
    assert np.allclose(centres[0], [1.5499999523162842, 3.0099990367889404, 2.6500000953674316], atol=1e-12)


def test_prior_defaults_and_no_cpu_path():
    """ses3d_prior_*: defaults are pose_prior's constants (PRI:39-66) and gtsam's default LM parameters; the handle
    cannot be created without a GPU."""
    p = layouts.PriorParams()
    L.load().ses3d_prior_default_params(C.byref(p))
    q = layouts.default_prior_params()
    for name, _ in layouts.PriorParams._fields_:
        assert getattr(p, name) == getattr(q, name), name
    assert (p.min_num_obs_track, p.dist_threshold, p.merge_dist_thresh, p.t_max_unobserved) == (10, 5.0, 0.20, 1.0)
    assert (p.pred_noise_sigma, p.default_res_sigma, p.avg_delay, p.root_sigma_factor) == (0.12, 0.10, 0.10, 100.0)
    assert abs(p.min_score - 0.10) < 1e-8
    assert (p.lm_lambda_initial, p.lm_lambda_factor, p.lm_lambda_upper_bound, p.lm_max_iterations) == (1e-5, 10.0, 1e5, 100)
    assert C.sizeof(layouts.PriorParams) == 128   # 6 x 4 + 13 x 8
    h = C.c_void_p()
    assert L.load().ses3d_prior_create(C.byref(p), 0, 8, 0, C.byref(h)) == layouts.E_INVALID
    assert L.load().ses3d_prior_create(C.byref(p), 1, 65, 0, C.byref(h)) == layouts.E_INVALID
    import torch
    if not torch.cuda.is_available():
        assert L.load().ses3d_prior_create(C.byref(p), 1, 8, 0, C.byref(h)) == layouts.E_CUDA
        assert b"no CPU path" in L.load().ses3d_last_error_string()
