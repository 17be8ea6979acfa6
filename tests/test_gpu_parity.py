"""GPU parity tests: the CUDA library, called through the C ABI, against the CPU oracle.

Association indices: bit-exact. 3-D joints: within 1e-3 m (FP32 mode) / 1e-4 m (FP64 mode) of the
oracle (north_star tolerance). Reprojection: bit-exact (FP64 path compiled without FMA contraction)."""
import numpy as np
import pytest

from oracle.binding import Oracle
from smartedgesensor3dhumanpose_b200 import api
from smartedgesensor3dhumanpose_b200.layouts import PRECISION_FP64, default_params
from tests import helpers

pytestmark = pytest.mark.gpu

POS_TOL_FP32 = 1e-3  # metres, north_star
POS_TOL_FP64 = 1e-4


def _run_pair(name, n_frames, params=None, outliers=0.0, **over):
    fr = helpers.make_workload(name, n_frames, **over)
    if outliers:
        helpers.inject_outliers(fr, outliers)
    params = params or default_params()
    orc = Oracle(fr["cameras"], params, ref_hungarian=True)
    gpu = api.GeometryPipeline(fr["cameras"], params)
    ro = orc.triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"], n_threads=8)
    rg = gpu.triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    assert ro["status"] == 0
    return fr, orc, gpu, ro, rg


@pytest.mark.parametrize("name,n_frames", [("cfg1_ring4x1", 2000), ("cfg2_hall16x6", 1500), ("cfg3_hall16x6_dropout", 1500),
                                           ("cfg5_ring8x4", 1500), ("dense_ring16x6", 300), ("cfg4_crowd64x20", 96)])
def test_association_bit_exact_and_joints_fp32(name, n_frames):
    fr, orc, gpu, ro, rg = _run_pair(name, n_frames)
    assert np.array_equal(gpu.tables()[1], orc.tables()[1]), "fundamental matrices differ"
    assert np.array_equal(ro["hyp_of"], rg["hyp_of"]), "association indices differ"
    assert np.array_equal(ro["n_hyp"], rg["n_hyp"])
    assert np.array_equal(ro["n_hungarian"], rg["n_hungarian"])
    st = helpers.compare_persons3d(ro, rg, POS_TOL_FP32)
    assert st["n_joints"] > 0


@pytest.mark.parametrize("name,n_frames", [("cfg3_hall16x6_dropout", 800), ("cfg5_ring8x4", 800)])
def test_joints_fp64_mode(name, n_frames):
    fr, orc, gpu, ro, rg = _run_pair(name, n_frames, params=default_params(precision=PRECISION_FP64))
    assert np.array_equal(ro["hyp_of"], rg["hyp_of"])
    helpers.compare_persons3d(ro, rg, POS_TOL_FP64, cov_rtol=1e-6, score_tol=1e-6)


MARGIN_EPS = 1e-4  # relative half-width of the band around a branch threshold inside which two float SVDs may disagree


@pytest.mark.parametrize("name,n_frames", [("cfg5_ring8x4", 1500), ("dense_ring16x6", 300), ("cfg2_hall16x6", 1500)])
def test_outlier_rejection_branches(name, n_frames):
    """Gross 2-D outliers exercise the 3-view epipolar and >=4-view leave-one-out branches (S3D:748-838).

    Branch decisions compare floating-point results with thresholds (S3D:748/793 err > 0.05, S3D:775 d < bestDist,
    S3D:813 e_sub vs best / 0.9 err, S3D:943, 964, 988). The oracle reports for every frame how close its closest
    decision came to its threshold (relative); frames inside the band MARGIN_EPS are excluded - two correct float
    implementations may take different sides there - and the excluded fraction is printed. EVERY other frame must
    agree: same persons, same joints, positions within 1e-3 m, scores within 2e-5, covariances within 1 %."""
    fr = helpers.make_workload(name, n_frames, h_max=40)
    helpers.inject_outliers(fr, 0.06)
    orc = Oracle(fr["cameras"], ref_hungarian=True)
    gpu = api.GeometryPipeline(fr["cameras"])
    ro = orc.triangulate_batch(fr["persons"], fr["n_persons"], 40, n_threads=8, diag=True)
    rg = gpu.triangulate_batch(fr["persons"], fr["n_persons"], 40)
    assert np.array_equal(ro["hyp_of"], rg["hyp_of"])
    keep = ro["margin"] >= MARGIN_EPS
    excluded = int((~keep).sum())
    print(f"{name}: {excluded} of {n_frames} frames ({100.0 * excluded / n_frames:.2f} %) inside the eps-band {MARGIN_EPS:g}")
    assert excluded <= n_frames // 20
    sub = lambda r: dict(persons3d=r["persons3d"][keep], n_out=r["n_out"][keep])
    st = helpers.compare_persons3d(sub(ro), sub(rg), POS_TOL_FP32, cov_rtol=1e-2)
    assert st["n_joints"] > 1000


def _outlier_chunk(name, f0, n, seed_chunk=20000, frac=0.05):
    """Frames [f0, f0+n) of the soak's outlier runs: scripts/soak_parity.py injects outliers per 20000-frame chunk with
    the chunk's first frame as seed, so the chunk is rebuilt and sliced."""
    c0 = f0 // seed_chunk * seed_chunk
    fr = helpers.make_workload(name, seed_chunk, first_frame=c0, h_max=40)
    helpers.inject_outliers(fr, frac, seed=c0)
    sl = slice(f0 - c0, f0 - c0 + n)
    return dict(fr, persons=fr["persons"][sl].copy(), n_persons=fr["n_persons"][sl].copy())


@pytest.mark.parametrize("name,f0,n,outliers", [
    ("cfg3_hall16x6_dropout", 6900, 120, False), ("cfg3_hall16x6_dropout", 24000, 60, False),
    ("cfg3_hall16x6_dropout", 27400, 60, False), ("cfg3_hall16x6_dropout", 36760, 60, False),
    ("cfg5_ring8x4", 4650, 120, True), ("cfg5_ring8x4", 17300, 300, True), ("cfg5_ring8x4", 18100, 100, True),
    ("cfg5_ring8x4", 19740, 80, True), ("cfg5_ring8x4", 40480, 60, True), ("cfg5_ring8x4", 53400, 60, True)])
def test_soak_offender_slices_have_zero_frames_over_tolerance(name, f0, n, outliers):
    """Fixed slices of the parity soak that contain the frames in which round 1's FP32 path left the 1e-3 m tolerance
    (cfg3: frames 6958, 24017, 27427, 36792; ring8 + 5 % outliers: 4714, 17544, 18149, 19780, 40502, 53434, ...): joints
    hundreds of metres away from mismatched / nearly parallel views, and high-residual joints. These are now re-solved
    in the oracle's operation order (tri_core.h::exact_weighted_resolve), so outside the eps-band NO frame may differ."""
    if outliers:
        fr = _outlier_chunk(name, f0, n)
    else:
        fr = helpers.make_workload(name, n, first_frame=f0, h_max=40)
    orc = Oracle(fr["cameras"], ref_hungarian=True)
    gpu = api.GeometryPipeline(fr["cameras"])
    ro = orc.triangulate_batch(fr["persons"], fr["n_persons"], 40, n_threads=8, diag=True)
    rg = gpu.triangulate_batch(fr["persons"], fr["n_persons"], 40)
    assert np.array_equal(ro["hyp_of"], rg["hyp_of"])
    keep = ro["margin"] >= MARGIN_EPS
    assert keep.sum() >= n - max(4, n // 20), f"{int((~keep).sum())} of {n} frames inside the eps-band"
    sub = lambda r: dict(persons3d=r["persons3d"][keep], n_out=r["n_out"][keep])
    helpers.compare_persons3d(sub(ro), sub(rg), POS_TOL_FP32, cov_rtol=1e-2)
    # joints beyond the far-point radius (20 m) are exact, not merely within tolerance
    ka, kb = ro["persons3d"]["keypoints"][keep], rg["persons3d"]["keypoints"][keep]
    far = (ka["score"] > 0) & (ka["x"] ** 2 + ka["y"] ** 2 + ka["z"] ** 2 > 21.0 ** 2)
    for c in "xyz":
        assert np.array_equal(ka[c][far], kb[c][far])
    assert np.array_equal(ka["score"][far], kb["score"][far])


def test_lm_refinement_matches_oracle():
    prm = default_params(lm_refine=1)
    fr, orc, gpu, ro, rg = _run_pair("cfg3_hall16x6_dropout", 600, params=prm)
    assert np.array_equal(ro["hyp_of"], rg["hyp_of"])
    helpers.compare_persons3d(ro, rg, POS_TOL_FP32, cov_rtol=5e-2)


@pytest.mark.parametrize("name,n_frames", [("cfg2_hall16x6", 800), ("cfg5_ring8x4", 500), ("cfg4_crowd64x20", 48)])
def test_reprojection_bit_exact(name, n_frames):
    fr, orc, gpu, ro, rg = _run_pair(name, n_frames)
    # feed the ORACLE's 3-D persons to both, so the comparison isolates the reprojection stage
    po = orc.reproject_batch(ro["persons3d"], ro["n_out"])
    pg = gpu.reproject_batch(ro["persons3d"], ro["n_out"])
    st = helpers.compare_persons2d(po, pg, px_tol=0.0)
    assert st["n_persons"] > 0


def test_process_batch_chains_both_stages():
    fr, orc, gpu, ro, rg = _run_pair("cfg2_hall16x6", 700)
    full = gpu.process_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    assert np.array_equal(full["n_out3d"], rg["n_out"])
    live = np.arange(fr["h_max"])[None, :] < rg["n_out"][:, None]
    assert full["persons3d"][live].tobytes() == rg["persons3d"][live].tobytes()
    two = gpu.reproject_batch(rg["persons3d"], rg["n_out"])
    helpers.compare_persons2d(two, dict(persons2d=full["persons2d"], n_out=full["n_out2d"]), px_tol=0.0)


def test_edge_cases_empty_and_ragged():
    fr = helpers.make_workload("cfg5_ring8x4", 64)
    persons, n_persons = fr["persons"].copy(), fr["n_persons"].copy()
    n_persons[0] = 0                      # no detections at all
    n_persons[1] = 0; n_persons[1, 3] = 2  # a single camera with detections -> empty output (S3D:557-560)
    n_persons[2, :4] = 0                  # leading cameras empty: seeding moves on (S3D:567-586)
    persons["keypoints"]["score"][3, 0] = 0.1  # first camera has only invalid persons
    persons["keypoints"]["score"][4] = 0.3     # scores exactly at the threshold (>= vs > asymmetry S3D:321,354)
    orc = Oracle(fr["cameras"], ref_hungarian=True)
    gpu = api.GeometryPipeline(fr["cameras"])
    ro = orc.triangulate_batch(persons, n_persons, 40)   # frame 4: no shared joint -> every detection its own hypothesis
    rg = gpu.triangulate_batch(persons, n_persons, 40)
    assert ro["status"] == 0 and ro["n_hyp"][4] == 32
    assert np.array_equal(ro["hyp_of"], rg["hyp_of"])
    assert ro["n_out"][0] == 0 and ro["n_out"][1] == 0 and rg["n_out"][0] == 0 and rg["n_out"][1] == 0
    helpers.compare_persons3d(ro, rg, POS_TOL_FP32)
    # zero frames is a no-op
    z = gpu.triangulate_batch(persons[:0], n_persons[:0], 40)
    assert z["n_out"].shape == (0,)


def test_capacity_overflow_is_reported():
    from smartedgesensor3dhumanpose_b200.lib import Ses3dError
    fr = helpers.make_workload("cfg5_ring8x4", 32)
    gpu = api.GeometryPipeline(fr["cameras"])
    with pytest.raises(Ses3dError) as ei:
        gpu.triangulate_batch(fr["persons"], fr["n_persons"], 2)
    assert ei.value.code == -3


def test_reference_style_single_frame_api():
    fr = helpers.make_workload("cfg2_hall16x6", 4)
    orc = Oracle(fr["cameras"], ref_hungarian=True)
    node = api.Skeleton3D(fr["cameras"])
    rep = api.PoseReprojection(fr["cameras"])
    for f in range(4):
        people = [fr["persons"][f, c, :fr["n_persons"][f, c]] for c in range(16)]
        persons3d = node.triangulate_persons(people)
        ro = orc.triangulate_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], 32)
        assert len(persons3d) == ro["n_out"][0]
        helpers.compare_persons3d(ro, dict(persons3d=np.pad(persons3d, (0, 32 - len(persons3d)))[None], n_out=ro["n_out"]),
                                  POS_TOL_FP32)
        per_cam = rep.fused_skeleton_callback(persons3d)
        po = orc.reproject_batch(np.pad(persons3d, (0, 32 - len(persons3d)))[None], ro["n_out"])
        assert [len(p) for p in per_cam] == po["n_out"][0].tolist()


def test_device_buffer_path_and_device_generator():
    import torch
    from smartedgesensor3dhumanpose_b200 import lib as _lib, synth
    from smartedgesensor3dhumanpose_b200.layouts import person2d_dtype, person_cov_dtype
    import ctypes as C
    fr = helpers.make_workload("cfg5_ring8x4", 512)
    cams, n_frames, p_max, h_max = fr["cameras"], 512, fr["persons"].shape[2], fr["h_max"]
    cfg = synth.synth_config(seed=5, n_people=4, dropout=0.05, area=(-2, -2, 2, 2))
    dev = torch.device("cuda:0")
    d_persons = torch.empty(fr["persons"].nbytes, dtype=torch.uint8, device=dev)
    d_np = torch.empty((n_frames, 8), dtype=torch.int32, device=dev)
    rc = _lib.load().ses3d_synth_frames_device(8, cams.ctypes.data, C.byref(cfg), 0, n_frames, d_persons.data_ptr(),
                                               d_np.data_ptr(), None, None)
    assert rc == 0
    # host and device generators are bit-identical
    assert d_persons.cpu().numpy().tobytes() == fr["persons"].tobytes()
    assert np.array_equal(d_np.cpu().numpy(), fr["n_persons"])
    gpu = api.GeometryPipeline(cams)
    d_out = torch.zeros(n_frames * h_max * person_cov_dtype.itemsize, dtype=torch.uint8, device=dev)
    d_nout = torch.zeros(n_frames, dtype=torch.int32, device=dev)
    gpu.triangulate_device(n_frames, p_max, h_max, d_persons.data_ptr(), d_np.data_ptr(), d_out.data_ptr(),
                           d_nout.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    host = gpu.triangulate_batch(fr["persons"], fr["n_persons"], h_max)
    assert np.array_equal(d_nout.cpu().numpy(), host["n_out"])
    got = d_out.cpu().numpy().view(person_cov_dtype).reshape(n_frames, h_max)
    live = np.arange(h_max)[None, :] < host["n_out"][:, None]
    assert got[live].tobytes() == host["persons3d"][live].tobytes()


@pytest.mark.parametrize("case", [c[0] for c in __import__("scripts.make_golden", fromlist=["CASES"]).CASES])
def test_gpu_against_committed_golden_vectors(case):
    """The committed fixtures (tests/golden, made by scripts/make_golden.py from the oracle + the reference's
    verbatim Hungarian.cpp) do not need the oracle or /root/reference at run time."""
    import hashlib
    from pathlib import Path
    import scripts.make_golden as mg
    g = np.load(Path(__file__).resolve().parent / "golden" / "golden_v1.npz")
    name, workload, n_frames, outliers, prm, first = next(c for c in mg.CASES if c[0] == case)
    fr = helpers.make_workload(workload, n_frames, first_frame=first, h_max=mg.H_MAX)
    if outliers:
        helpers.inject_outliers(fr, outliers, seed=7)
    assert hashlib.sha256(fr["persons"].tobytes() + fr["n_persons"].tobytes()).digest() == g[f"{name}/input_sha256"].tobytes()
    gpu = api.GeometryPipeline(fr["cameras"], default_params(**prm))
    r = gpu.triangulate_batch(fr["persons"], fr["n_persons"], mg.H_MAX)
    assert np.array_equal(r["hyp_of"], g[f"{name}/hyp_of"])
    assert np.array_equal(r["n_hungarian"], g[f"{name}/n_hungarian"]) and np.array_equal(r["n_hyp"], g[f"{name}/n_hyp"])
    # positions are checked in every case, outlier cases included; only frames whose closest branch decision lies inside
    # the eps-band recorded with the fixture are exempt (none in the committed fixture: asserted)
    keep = g[f"{name}/margin"] >= mg.MARGIN_EPS
    assert keep.all(), "a golden frame sits inside the eps-band: pick another seed for the fixture"
    assert np.array_equal(r["n_out"], g[f"{name}/n_out"])
    live = np.arange(mg.H_MAX)[None, :] < r["n_out"][:, None]
    kp = r["persons3d"]["keypoints"][live]
    tol = POS_TOL_FP64 if prm.get("precision") else POS_TOL_FP32
    d = np.linalg.norm(np.stack([kp["x"], kp["y"], kp["z"]], -1) - g[f"{name}/xyz"], axis=-1)
    assert np.array_equal(kp["score"] > 0, g[f"{name}/score"] > 0) and d[kp["score"] > 0].max() <= tol
    assert np.abs(kp["score"] - g[f"{name}/score"]).max() <= 2e-5


def test_nan_trap_is_reproduced_not_hidden():
    """Zero 2-D covariance -> NaN 3-D covariance in the reference (S3D:473-475); the library must not invent numbers."""
    fr = helpers.make_workload("cfg5_ring8x4", 8)
    persons = fr["persons"].copy()
    persons["keypoints"]["cov"][5] = 0.0
    orc = Oracle(fr["cameras"], ref_hungarian=True)
    gpu = api.GeometryPipeline(fr["cameras"])
    ro = orc.triangulate_batch(persons, fr["n_persons"], fr["h_max"])
    rg = gpu.triangulate_batch(persons, fr["n_persons"], fr["h_max"])
    assert np.array_equal(ro["n_out"], rg["n_out"]) and ro["n_out"][5] > 0
    a = ro["persons3d"][5, :ro["n_out"][5]]["keypoints"]
    b = rg["persons3d"][5, :rg["n_out"][5]]["keypoints"]
    assert np.isnan(a["cov"]).any() and np.array_equal(np.isnan(a["cov"]), np.isnan(b["cov"]))


def test_multi_chunk_batches_and_reuse_of_the_handle():
    """More frames than one internal device chunk (16384) and than one host chunk; second call reuses scratch."""
    fr = helpers.make_workload("cfg1_ring4x1", 20000)
    gpu = api.GeometryPipeline(fr["cameras"])
    a = gpu.process_batch(fr["persons"], fr["n_persons"], 4)
    b = gpu.process_batch(fr["persons"], fr["n_persons"], 4)
    assert a["persons3d"].tobytes() == b["persons3d"].tobytes() and a["persons2d"].tobytes() == b["persons2d"].tobytes()
    ro = Oracle(fr["cameras"], ref_hungarian=True).triangulate_batch(fr["persons"], fr["n_persons"], 4, n_threads=8)
    helpers.compare_persons3d(ro, dict(persons3d=a["persons3d"], n_out=a["n_out3d"]), POS_TOL_FP32)


def test_ragged_batch_call_equals_padded_call():
    """ses3d_process_batch_ragged: dense records in/out, same results as the padded call, several chunks."""
    from smartedgesensor3dhumanpose_b200.layouts import person2d_dtype, person_cov_dtype
    fr = helpers.make_workload("cfg2_hall16x6", 3000)
    gpu = api.GeometryPipeline(fr["cameras"])
    h_max, p_max = fr["h_max"], fr["persons"].shape[2]
    pad = gpu.process_batch(fr["persons"], fr["n_persons"], h_max)
    dense_in = api.to_ragged(fr["persons"], fr["n_persons"])
    t3e, t2e = int(pad["n_out3d"].sum()), int(pad["n_out2d"].sum())
    out3d = np.zeros(t3e + 5, person_cov_dtype)
    out2d = np.zeros(t2e + 5, person2d_dtype)
    n3 = np.zeros(3000, np.int32)
    n2 = np.zeros((3000, 16), np.int32)
    t3, t2 = gpu.process_batch_ragged(dense_in, fr["n_persons"], p_max, h_max, out3d, n3, out2d, n2)
    assert (t3, t2) == (t3e, t2e)
    assert np.array_equal(n3, pad["n_out3d"]) and np.array_equal(n2, pad["n_out2d"])
    live3 = np.arange(h_max)[None, :] < n3[:, None]
    assert out3d[:t3].tobytes() == pad["persons3d"][live3].tobytes()
    live2 = np.arange(h_max)[None, None, :] < n2[:, :, None]
    assert out2d[:t2].tobytes() == pad["persons2d"][live2].tobytes()
    back = api.from_ragged(out3d, n3, h_max, person_cov_dtype)
    assert back[live3].tobytes() == pad["persons3d"][live3].tobytes()
    # too-small output capacity is reported, not overrun
    from smartedgesensor3dhumanpose_b200.lib import Ses3dError
    with pytest.raises(Ses3dError) as ei:
        gpu.process_batch_ragged(dense_in, fr["n_persons"], p_max, h_max, out3d[:10], n3, out2d, n2)
    assert ei.value.code == -3


@pytest.mark.parametrize("shape", [(1, 1), (3, 3), (6, 6), (4, 7), (7, 4), (20, 20), (24, 20), (20, 31), (44, 20), (70, 33)])
def test_device_munkres_matches_reference_indices(shape):
    """The warp-cooperative Munkres against the reference's verbatim Hungarian.cpp, including heavily tied matrices."""
    from oracle import binding as ob
    from smartedgesensor3dhumanpose_b200 import rigs
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    costs = rng.random((200,) + shape)
    costs[::3][rng.random((len(costs[::3]),) + shape) < 0.6] = 1e6
    costs[::5] = np.round(costs[::5] * 4) / 4
    costs[7] = 0.0
    gpu = api.GeometryPipeline(rigs.ring4())
    got = gpu.munkres_batch(costs)
    solver = ob.ref_munkres if ob.REF_HUNGARIAN_PATH.exists() else ob.munkres
    for i in range(len(costs)):
        want, _ = solver(costs[i])
        assert np.array_equal(got[i], want), (shape, i)


def test_non_finite_inputs_terminate():
    """NaN / inf keypoints (undefined behaviour in the reference) must not hang the GPU; other frames are unaffected."""
    fr = helpers.make_workload("cfg5_ring8x4", 64)
    persons = fr["persons"].copy()
    persons["keypoints"]["x"][3] = np.nan
    persons["keypoints"]["y"][7, 2] = np.inf
    persons["keypoints"]["score"][9, 1, 0, :] = np.nan
    gpu = api.GeometryPipeline(fr["cameras"])
    rg = gpu.process_batch(persons, fr["n_persons"], 40)
    ro = Oracle(fr["cameras"], ref_hungarian=False).triangulate_batch(fr["persons"], fr["n_persons"], 40)
    ok = np.ones(64, bool); ok[[3, 7, 9]] = False
    helpers.compare_persons3d(dict(persons3d=ro["persons3d"][ok], n_out=ro["n_out"][ok]),
                              dict(persons3d=rg["persons3d"][ok], n_out=rg["n_out3d"][ok]), POS_TOL_FP32)


@pytest.mark.parametrize("prm", [dict(pose_method=1), dict(max_epipolar_error=0.045), dict(pose_method=1, precision=1),
                                 dict(merge_dist_thresh=0.8, max_joint_dist_to_root=1.0), dict(max_epipolar_error=0.0005)])
def test_parameter_variants(prm):
    """h36m tables, the demo's max_epi_dist, and thresholds that fire the merge / root-distance branches."""
    fr, orc, gpu, ro, rg = _run_pair("cfg5_ring8x4", 400, params=default_params(**prm), h_max=40)
    assert np.array_equal(ro["hyp_of"], rg["hyp_of"])
    helpers.compare_persons3d(ro, rg, POS_TOL_FP64 if prm.get("precision") else POS_TOL_FP32, cov_rtol=5e-2)
    po = orc.reproject_batch(ro["persons3d"], ro["n_out"])
    pg = gpu.reproject_batch(ro["persons3d"], ro["n_out"])
    helpers.compare_persons2d(po, pg, px_tol=0.0)


def test_device_calls_are_stream_ordered_and_overflow_is_reported_by_check():
    """ses3d.h: device-buffer calls enqueue on the caller's stream and return without a host synchronisation. Two
    batches are enqueued back to back on a side stream (no sync in between) and must both equal the host-path results;
    ses3d_check() then reports OK. A call whose frames exceed h_max returns OK at once and the overflow is raised by
    check() (or by the next call)."""
    import torch
    from smartedgesensor3dhumanpose_b200.layouts import person_cov_dtype, person2d_dtype
    from smartedgesensor3dhumanpose_b200.lib import Ses3dError
    dev = torch.device("cuda:0")
    gpu = api.GeometryPipeline(helpers.make_workload("cfg5_ring8x4", 1)["cameras"])
    batches = [helpers.make_workload("cfg5_ring8x4", 3000, first_frame=f0) for f0 in (0, 3000)]
    h_max, p_max, C_ = batches[0]["h_max"], batches[0]["persons"].shape[2], 8
    want = [gpu.process_batch(b["persons"], b["n_persons"], h_max) for b in batches]
    side = torch.cuda.Stream(device=dev)
    outs = []
    with torch.cuda.stream(side):
        for b in batches:
            d_in = torch.from_numpy(b["persons"].view(np.uint8).reshape(-1)).to(dev, non_blocking=True)
            d_n = torch.from_numpy(b["n_persons"]).to(dev, non_blocking=True)
            o3 = torch.zeros(3000 * h_max * person_cov_dtype.itemsize, dtype=torch.uint8, device=dev)
            n3 = torch.zeros(3000, dtype=torch.int32, device=dev)
            o2 = torch.zeros(3000 * C_ * h_max * person2d_dtype.itemsize, dtype=torch.uint8, device=dev)
            n2 = torch.zeros(3000 * C_, dtype=torch.int32, device=dev)
            gpu.process_device(3000, p_max, h_max, d_in.data_ptr(), d_n.data_ptr(), o3.data_ptr(), n3.data_ptr(),
                               o2.data_ptr(), n2.data_ptr(), stream=side.cuda_stream)
            outs.append((d_in, d_n, o3, n3, o2, n2))
    gpu.check()                               # waits for both batches; no overflow
    for w, (_, _, o3, n3, o2, n2) in zip(want, outs):
        assert np.array_equal(n3.cpu().numpy(), w["n_out3d"])
        got3 = o3.cpu().numpy().view(person_cov_dtype).reshape(3000, h_max)
        live = np.arange(h_max)[None, :] < w["n_out3d"][:, None]
        assert got3[live].tobytes() == w["persons3d"][live].tobytes()
        assert np.array_equal(n2.cpu().numpy().reshape(3000, C_), w["n_out2d"])
        got2 = o2.cpu().numpy().view(person2d_dtype).reshape(3000, C_, h_max)
        live2 = np.arange(h_max)[None, None, :] < w["n_out2d"][:, :, None]
        assert got2[live2].tobytes() == w["persons2d"][live2].tobytes()
    # overflow: h_max = 2 is too small for four people
    d_in, d_n = outs[0][0], outs[0][1]
    o3 = torch.zeros(3000 * 2 * person_cov_dtype.itemsize, dtype=torch.uint8, device=dev)
    n3 = torch.zeros(3000, dtype=torch.int32, device=dev)
    gpu.triangulate_device(3000, p_max, 2, d_in.data_ptr(), d_n.data_ptr(), o3.data_ptr(), n3.data_ptr(),
                           stream=side.cuda_stream)   # returns OK: nothing has been waited for
    with pytest.raises(Ses3dError) as ei:
        gpu.check()
    assert ei.value.code == -3
    gpu.check()                               # reported once, then clear


def test_ragged_call_direct_and_staged_modes_agree():
    """Host outputs leave through the copy engine (staged, the default); with SES3D_RAGGED_DIRECT=1 the pack kernels
    write pinned outputs themselves (mapped memory): same bytes either way."""
    import os
    import torch
    from smartedgesensor3dhumanpose_b200.layouts import person2d_dtype, person_cov_dtype
    fr = helpers.make_workload("cfg2_hall16x6", 5000)
    h_max, p_max = fr["h_max"], fr["persons"].shape[2]
    dense_in = api.to_ragged(fr["persons"], fr["n_persons"])
    res = []
    for pinned in (False, True):
        if pinned:
            os.environ["SES3D_RAGGED_DIRECT"] = "1"   # read once, at ses3d_create
        try:
            gpu = api.GeometryPipeline(fr["cameras"])
        finally:
            os.environ.pop("SES3D_RAGGED_DIRECT", None)
        def buf(n, dt):
            if not pinned:
                return np.zeros(n, dt)
            t = torch.zeros(n * dt.itemsize, dtype=torch.uint8).pin_memory()
            keep.append(t)
            return t.numpy().view(dt)
        keep = []
        out3d, out2d = buf(5000 * 8, person_cov_dtype), buf(5000 * 16 * 4, person2d_dtype)
        n3, n2 = np.zeros(5000, np.int32), np.zeros((5000, 16), np.int32)
        t3, t2 = gpu.process_batch_ragged(dense_in, fr["n_persons"], p_max, h_max, out3d, n3, out2d, n2)
        res.append((t3, t2, out3d[:t3].tobytes(), out2d[:t2].tobytes(), n3.copy(), n2.copy()))
    a, b = res
    assert a[0] == b[0] and a[1] == b[1] and a[0] > 0 and a[1] > 0
    assert a[2] == b[2] and a[3] == b[3] and np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5])
    # capacity error in direct mode: nothing beyond the capacity is written
    t = torch.zeros(100 * person_cov_dtype.itemsize + 64, dtype=torch.uint8).pin_memory()
    t[-64:] = 0x5A
    small3 = t.numpy()[:-64].view(person_cov_dtype)
    t2_ = torch.zeros(5000 * 16 * 4 * person2d_dtype.itemsize, dtype=torch.uint8).pin_memory()
    from smartedgesensor3dhumanpose_b200.lib import Ses3dError
    with pytest.raises(Ses3dError) as ei:
        gpu.process_batch_ragged(dense_in, fr["n_persons"], p_max, h_max, small3, np.zeros(5000, np.int32),
                                 t2_.numpy().view(person2d_dtype), np.zeros((5000, 16), np.int32))
    assert ei.value.code == -3 and bool((t[-64:] == 0x5A).all())


def test_single_process_multi_device_entry_matches_single_device():
    """ses3d_create_multi / ses3d_multi_process_batch(_ragged): frames sharded over the device list from one process
    (the same device listed twice stands in for two GPUs on a one-GPU box)."""
    import torch
    from smartedgesensor3dhumanpose_b200.layouts import person2d_dtype, person_cov_dtype
    fr = helpers.make_workload("cfg5_ring8x4", 2001)
    one = api.GeometryPipeline(fr["cameras"])
    h_max, p_max = fr["h_max"], fr["persons"].shape[2]
    want = one.process_batch(fr["persons"], fr["n_persons"], h_max)
    n_dev = torch.cuda.device_count()
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    multi = api.MultiPipeline(fr["cameras"], devices=devices)
    assert multi.n_devices == len(devices)
    got = multi.process_batch(fr["persons"], fr["n_persons"], h_max, dump=True)
    assert np.array_equal(got["n_out3d"], want["n_out3d"]) and np.array_equal(got["n_out2d"], want["n_out2d"])
    assert got["persons3d"].tobytes() == want["persons3d"].tobytes()
    assert got["persons2d"].tobytes() == want["persons2d"].tobytes()
    assert np.array_equal(got["hyp_of"], one.triangulate_batch(fr["persons"], fr["n_persons"], h_max)["hyp_of"])
    dense_in = api.to_ragged(fr["persons"], fr["n_persons"])
    out3d, out2d = np.zeros(2001 * 6, person_cov_dtype), np.zeros(2001 * 8 * 5, person2d_dtype)   # per-device slices need slack
    n3, n2 = np.zeros(2001, np.int32), np.zeros((2001, 8), np.int32)
    seg3, seg2 = multi.process_batch_ragged(dense_in, fr["n_persons"], p_max, h_max, out3d, n3, out2d, n2)
    assert np.array_equal(n3, want["n_out3d"]) and np.array_equal(n2, want["n_out2d"])
    live3 = np.arange(h_max)[None, :] < n3[:, None]
    live2 = np.arange(h_max)[None, None, :] < n2[:, :, None]
    cat3 = np.concatenate([out3d[s:s + c] for s, c in seg3])
    cat2 = np.concatenate([out2d[s:s + c] for s, c in seg2])
    assert cat3.tobytes() == want["persons3d"][live3].tobytes()
    assert cat2.tobytes() == want["persons2d"][live2].tobytes()
