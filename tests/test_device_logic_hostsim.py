"""CPU tests of the CUDA library's *device algorithms* (csrc/*_core.h) instantiated serially by the test-only
build tests/hostsim: association indices bit-exact against the oracle, joints within the north_star tolerances,
reprojection bit-exact. This is not a product CPU path (libses3d.so has none); the same comparisons run on the
GPU through the C ABI in test_gpu_parity.py."""
import numpy as np
import pytest

from oracle.binding import Oracle
from smartedgesensor3dhumanpose_b200.layouts import PRECISION_FP64, default_params
from tests import helpers
from tests.hostsim.binding import HostSim


def _pair(name, n_frames, params=None, outliers=0.0, **over):
    fr = helpers.make_workload(name, n_frames, **over)
    if outliers:
        helpers.inject_outliers(fr, outliers)
    params = params or default_params()
    ro = Oracle(fr["cameras"], params).triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"], n_threads=4)
    rh = HostSim(fr["cameras"], params).triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    assert ro["status"] == 0 and rh["status"] == 0
    return fr, ro, rh


def test_camera_tables_bit_identical():
    from smartedgesensor3dhumanpose_b200 import rigs
    for rig in ("ring4", "hall16", "crowd64"):
        cams = rigs.RIGS[rig]()
        Po, Fo = Oracle(cams).tables()
        Ph, Fh = HostSim(cams).tables()
        assert np.array_equal(Po, Ph) and np.array_equal(Fo, Fh)


@pytest.mark.parametrize("name,n_frames", [("cfg1_ring4x1", 300), ("cfg2_hall16x6", 200), ("cfg3_hall16x6_dropout", 200),
                                           ("cfg5_ring8x4", 200), ("dense_ring16x6", 40), ("cfg4_crowd64x20", 2)])
def test_association_bit_exact_and_joints(name, n_frames):
    fr, ro, rh = _pair(name, n_frames)
    assert np.array_equal(ro["hyp_of"], rh["hyp_of"])
    assert np.array_equal(ro["n_hyp"], rh["n_hyp"]) and np.array_equal(ro["n_hungarian"], rh["n_hungarian"])
    helpers.compare_persons3d(ro, rh, 1e-3)


def test_fp64_mode():
    fr, ro, rh = _pair("cfg3_hall16x6_dropout", 150, params=default_params(precision=PRECISION_FP64))
    assert np.array_equal(ro["hyp_of"], rh["hyp_of"])
    helpers.compare_persons3d(ro, rh, 1e-4, cov_rtol=1e-6, score_tol=1e-6)


MARGIN_EPS = 1e-4


@pytest.mark.parametrize("name,n_frames", [("cfg5_ring8x4", 400), ("dense_ring16x6", 60)])
def test_outlier_rejection_branches(name, n_frames):
    """Frames whose closest branch decision is inside the eps-band around its threshold are excluded (the oracle's
    'margin' diagnostic); every other frame must agree completely (see tests/test_gpu_parity.py)."""
    fr = helpers.make_workload(name, n_frames, h_max=40)
    helpers.inject_outliers(fr, 0.06)
    ro = Oracle(fr["cameras"]).triangulate_batch(fr["persons"], fr["n_persons"], 40, n_threads=4, diag=True)
    rh = HostSim(fr["cameras"]).triangulate_batch(fr["persons"], fr["n_persons"], 40)
    assert np.array_equal(ro["hyp_of"], rh["hyp_of"])
    keep = ro["margin"] >= MARGIN_EPS
    assert (~keep).sum() <= n_frames // 20
    sub = lambda r: dict(persons3d=r["persons3d"][keep], n_out=r["n_out"][keep])
    helpers.compare_persons3d(sub(ro), sub(rh), 1e-3, cov_rtol=1e-2)


@pytest.mark.parametrize("first", [6958, 24017, 27427, 36792])
def test_far_points_are_resolved_exactly(first):
    """Soak offenders of round 1 (cfg3): a joint triangulated hundreds of metres away. The device algorithm re-solves
    such joints in the oracle's operation order (tri_core.h::exact_weighted_resolve): bit-identical position and score."""
    fr = helpers.make_workload("cfg3_hall16x6_dropout", 2, first_frame=first, h_max=40)
    ro = Oracle(fr["cameras"]).triangulate_batch(fr["persons"], fr["n_persons"], 40)
    rh = HostSim(fr["cameras"]).triangulate_batch(fr["persons"], fr["n_persons"], 40)
    helpers.compare_persons3d(ro, rh, 1e-3)
    ka, kb = ro["persons3d"]["keypoints"], rh["persons3d"]["keypoints"]
    far = (ka["score"] > 0) & (ka["x"] ** 2 + ka["y"] ** 2 + ka["z"] ** 2 > 21.0 ** 2)
    assert far.any()
    for c in ("x", "y", "z", "score"):
        assert np.array_equal(ka[c][far], kb[c][far])


def test_lm_refinement():
    fr, ro, rh = _pair("cfg3_hall16x6_dropout", 150, params=default_params(lm_refine=1))
    helpers.compare_persons3d(ro, rh, 1e-3, cov_rtol=5e-2)


@pytest.mark.parametrize("cam_tile", [0, 1, 3, 5])
def test_reprojection_bit_exact_for_every_camera_tile(cam_tile):
    fr, ro, rh = _pair("cfg2_hall16x6", 60)
    po = Oracle(fr["cameras"]).reproject_batch(ro["persons3d"], ro["n_out"])
    ph = HostSim(fr["cameras"]).reproject_batch(ro["persons3d"], ro["n_out"], cam_tile=cam_tile)
    st = helpers.compare_persons2d(po, ph, px_tol=0.0)
    assert st["n_persons"] > 0


def test_edge_cases():
    fr = helpers.make_workload("cfg5_ring8x4", 32)
    persons, n_persons = fr["persons"].copy(), fr["n_persons"].copy()
    n_persons[0] = 0
    n_persons[1] = 0; n_persons[1, 3] = 2
    n_persons[2, :4] = 0
    persons["keypoints"]["score"][3, 0] = 0.1
    persons["keypoints"]["score"][4] = 0.3
    persons["keypoints"]["cov"][5] = 0.0                  # zero 2-D covariance -> NaN 3-D covariance (S3D:473-475)
    ro = Oracle(fr["cameras"]).triangulate_batch(persons, n_persons, 40)
    rh = HostSim(fr["cameras"]).triangulate_batch(persons, n_persons, 40)
    assert np.array_equal(ro["hyp_of"], rh["hyp_of"]) and np.array_equal(ro["n_out"], rh["n_out"])
    assert ro["n_out"][0] == 0 and ro["n_out"][1] == 0 and ro["n_hyp"][4] == 32
    keep = np.ones(32, bool); keep[5] = False
    helpers.compare_persons3d(dict(persons3d=ro["persons3d"][keep], n_out=ro["n_out"][keep]),
                              dict(persons3d=rh["persons3d"][keep], n_out=rh["n_out"][keep]), 1e-3)
    a = ro["persons3d"][5, :ro["n_out"][5]]["keypoints"]
    b = rh["persons3d"][5, :rh["n_out"][5]]["keypoints"]
    assert np.array_equal(np.isnan(a["cov"]), np.isnan(b["cov"]))      # the NaN trap is reproduced, not hidden


def test_capacity_overflow_reported():
    fr = helpers.make_workload("cfg5_ring8x4", 8)
    rh = HostSim(fr["cameras"]).triangulate_batch(fr["persons"], fr["n_persons"], 2)
    assert rh["status"] == -3


@pytest.mark.parametrize("prm", [dict(pose_method=1), dict(max_epipolar_error=0.045), dict(pose_method=1, precision=1),
                                 dict(merge_dist_thresh=0.8, max_joint_dist_to_root=1.0)])
def test_parameter_variants(prm):
    """h36m skeleton tables (S3D:111-145), the demo launch file's max_epi_dist = 0.045, and thresholds that make the
    merge (S3D:984-996) and root-distance (S3D:937-953) branches fire."""
    fr, ro, rh = _pair("cfg5_ring8x4", 120, params=default_params(**prm))
    assert np.array_equal(ro["hyp_of"], rh["hyp_of"])
    helpers.compare_persons3d(ro, rh, 1e-4 if prm.get("precision") else 1e-3)


def test_merge_of_close_skeletons():
    """Two detections of the same person that the association failed to join are merged when closer than 0.20 m."""
    fr = helpers.make_workload("cfg5_ring8x4", 60)
    prm = default_params(max_epipolar_error=0.0005)      # almost nothing associates across > 2 cameras -> many duplicates
    ro = Oracle(fr["cameras"], prm).triangulate_batch(fr["persons"], fr["n_persons"], 40, n_threads=4)
    rh = HostSim(fr["cameras"], prm).triangulate_batch(fr["persons"], fr["n_persons"], 40)
    assert np.array_equal(ro["hyp_of"], rh["hyp_of"]) and np.array_equal(ro["n_out"], rh["n_out"])
    assert (ro["n_hyp"] > 4).any()
    helpers.compare_persons3d(ro, rh, 1e-3, cov_rtol=5e-2)


def test_reprojection_prefilter_is_exact_near_borders_and_behind_cameras():
    """The single-precision "certainly outside the image" pre-test (reproj_core.h) must never change a record:
    skeletons scattered all over (and outside) the hall, near the image borders, behind and very close to cameras,
    with covariances from millimetres to metres, NaNs and a non-SPD covariance - bitwise equal to the oracle."""
    from smartedgesensor3dhumanpose_b200.layouts import KP2FUSION_SIMPLE, person_cov_dtype
    from smartedgesensor3dhumanpose_b200.sequences import TEMPLATE
    rng = np.random.default_rng(77)
    for rig in ("hall16", "ring8"):
        cams = helpers.rigs.RIGS[rig]()
        F, H = 120, 6
        p3 = np.zeros((F, H), person_cov_dtype)
        n3 = np.full(F, H, np.int32)
        fus = np.array(KP2FUSION_SIMPLE)
        pos = rng.uniform(-12, 12, (F, H, 1, 3)) * [1, 1, 0.2]
        X = TEMPLATE[None, None] * rng.uniform(0.3, 3.0, (F, H, 1, 1)) + pos
        kp = p3["keypoints"]
        for i, c in enumerate("xyz"):
            v = np.zeros((F, H, 21)); v[..., fus] = X[..., i]; kp[c] = v
        sc = np.zeros((F, H, 21), np.float32); sc[..., fus] = rng.uniform(0.05, 1.0, (F, H, 17)); kp["score"] = sc
        A = rng.normal(0, 1, (F, H, 21, 3, 3)) * 10.0 ** rng.uniform(-3, 0.3, (F, H, 21, 1, 1))
        S = A @ np.swapaxes(A, -1, -2) + 1e-12 * np.eye(3)
        kp["cov"] = np.stack([S[..., 0, 0], S[..., 0, 1], S[..., 0, 2], S[..., 1, 1], S[..., 1, 2], S[..., 2, 2]], -1)
        p3["keypoints"] = kp
        p3["keypoints"]["x"][3, 1, 5] = np.nan
        p3["keypoints"]["cov"][4, 2, 6] = [1e-4, 5e-4, 0, 1e-4, 0, 1e-4]      # not positive definite -> NaN sigma points
        p3["keypoints"]["cov"][5, 0, 2] = 0.0
        po = Oracle(cams).reproject_batch(p3, n3)
        ph = HostSim(cams).reproject_batch(p3, n3)
        assert np.array_equal(po["n_out"], ph["n_out"])
        live = np.arange(H)[None, None, :] < po["n_out"][:, :, None]
        assert po["persons2d"][live].tobytes() == ph["persons2d"][live].tobytes()
        assert po["n_out"].sum() > 100


@pytest.mark.parametrize("name,n_frames", [("cfg4_crowd64x20", 3), ("cfg2_hall16x6", 80), ("cfg3_hall16x6_dropout", 80),
                                           ("dense_ring16x6", 30), ("cfg5_ring8x4", 80), ("cfg1_ring4x1", 60)])
def test_big_rig_association_path_bit_exact(name, n_frames):
    """Rigs that do not fit shared memory keep the normalised keypoints in global scratch; every association index and
    record must be identical to the shared-memory path and to the oracle."""
    from tests.hostsim import binding
    fr = helpers.make_workload(name, n_frames)
    ro = Oracle(fr["cameras"], ref_hungarian=True).triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    flat = HostSim(fr["cameras"]).triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    binding.set_big_rig_path(True)
    try:
        tiled = HostSim(fr["cameras"]).triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    finally:
        binding.set_big_rig_path(False)
    for key in ("hyp_of", "n_hyp", "n_hungarian", "n_out"):
        assert np.array_equal(ro[key], tiled[key]), key
        assert np.array_equal(flat[key], tiled[key]), key
    assert flat["persons3d"].tobytes() == tiled["persons3d"].tobytes()
