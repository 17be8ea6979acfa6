"""GPU parity of the pose_prior stage: ses3d_prior_run through the C ABI against the CPU oracle
(oracle/pose_prior_oracle.cpp with the reference's verbatim Hungarian.cpp). Track assignment, ids, publication
counts and scores: exact. Fused / predicted joints: within 1e-6 m (FP64 everywhere; the only difference is the
elimination order — tree on the GPU, dense Cholesky in the oracle); covariances within 1e-5 relative."""
import numpy as np
import pytest

from oracle.binding import PriorOracle
from smartedgesensor3dhumanpose_b200 import api
from smartedgesensor3dhumanpose_b200.layouts import default_prior_params, person_cov_dtype
from smartedgesensor3dhumanpose_b200.sequences import synth_person_sequences
from tests.test_pose_prior import compare_runs

pytestmark = pytest.mark.gpu

POS_TOL = 1e-6   # metres
COV_RTOL = 1e-5


@pytest.mark.parametrize("kw", [dict(), dict(joint_dropout=0.3, person_dropout=0.1), dict(n_people=8, noise_m=0.03),
                                dict(pose_method=1), dict(normalize_by_height=1), dict(min_num_obs_track=0),
                                dict(pose_method=1, normalize_by_height=1, joint_dropout=0.2)])
def test_prior_matches_oracle(kw):
    kw = dict(kw)
    pkw = {k: kw.pop(k) for k in ("normalize_by_height", "min_num_obs_track") if k in kw}
    if "pose_method" in kw:
        pkw["pose_method"] = kw["pose_method"]
    S, T = 24, 60
    seq = synth_person_sequences(S, T, kw.pop("n_people", 5), seed=21, **kw)
    prm = default_prior_params(**pkw)
    ro = PriorOracle(prm, S, ref_hungarian=True).run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"],
                                                      n_threads=8)
    gpu = api.PriorTracker(prm, S)
    rg = gpu.run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"])
    worst = compare_runs(ro, rg, POS_TOL, COV_RTOL)
    assert ro["n_out"].sum() > S * 10
    assert gpu.launch_count >= 2
    print("max joint deviation", worst)


def test_prior_track_tables_match_oracle():
    S, T = 6, 45
    seq = synth_person_sequences(S, T, 4, seed=22, person_dropout=0.15)
    o = PriorOracle(default_prior_params(), S, ref_hungarian=True)
    o.run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"])
    g = api.PriorTracker(default_prior_params(), S)
    g.run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"])
    for s in range(S):
        io, no = o.tracks(s)
        ig, ng = g.tracks(s)
        assert np.array_equal(io, ig) and np.array_equal(no, ng)


def test_prior_streaming_and_node_mirror():
    """One message per call (the ROS-shim use case) == the whole sequence in one launch; PosePrior mirrors the node."""
    seq = synth_person_sequences(1, 40, 3, seed=23)
    prm = default_prior_params()
    whole = api.PriorTracker(prm, 1).run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"])
    node = api.PosePrior(prm, h_max=seq["h_max"])
    for t in range(40):
        n = seq["n_persons"][0, t]
        fused, pred, delay = node.skeleton_callback(seq["persons"][0, t, :n], int(seq["stamp_ns"][0, t]),
                                                    seq["fb_delay"][0, t])
        k = whole["n_out"][0, t]
        assert len(fused) == k and len(pred) == k
        assert fused.tobytes() == whole["fused"][0, t, :k].tobytes()
        assert pred.tobytes() == whole["pred"][0, t, :k].tobytes()
        assert np.float32(delay) == whole["pred_delay"][0, t]
    node.reset()
    fused, _, _ = node.skeleton_callback(seq["persons"][0, 0, :seq["n_persons"][0, 0]], int(seq["stamp_ns"][0, 0]))
    assert len(fused) == 0


def test_prior_results_do_not_depend_on_the_launch_shape():
    """Few streams run one detection per warp on up to six warps, many streams three detections per warp on two
    (kernels_prior.cu::launch_prior); SES3D_PRIOR_GROUP / SES3D_PRIOR_WARPS force a shape. Same bytes either way."""
    import os
    seq = synth_person_sequences(5, 24, 7, seed=77, joint_dropout=0.15, person_dropout=0.1, h_max=10)
    prm = default_prior_params(min_num_obs_track=3)
    ref = api.PriorTracker(prm, 5).run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"])   # 1 x 6
    assert ref["n_out"].sum() > 0
    for group, warps in [(3, 2), (2, 3), (1, 4), (3, 1)]:
        os.environ["SES3D_PRIOR_GROUP"], os.environ["SES3D_PRIOR_WARPS"] = str(group), str(warps)
        try:
            trk = api.PriorTracker(prm, 5)
        finally:
            del os.environ["SES3D_PRIOR_GROUP"], os.environ["SES3D_PRIOR_WARPS"]
        r = trk.run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"])
        for k in ("fused", "pred", "n_out", "pred_delay", "track_of"):
            assert r[k].tobytes() == ref[k].tobytes(), (group, warps, k)
        trk.close()


def test_prior_device_buffers_and_chain():
    """Device-resident call (torch only provides the memory) fed by the triangulation stage's output layout."""
    import torch
    S, T = 64, 32
    seq = synth_person_sequences(S, T, 4, seed=24)
    prm = default_prior_params()
    host = api.PriorTracker(prm, S).run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"])
    dev = torch.device("cuda:0")
    H, C = seq["h_max"], seq["n_cams"]
    tb = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
    d_p, d_n, d_s, d_f = tb(seq["persons"]), tb(seq["n_persons"]), tb(seq["stamp_ns"]), tb(seq["fb_delay"])
    rec = person_cov_dtype.itemsize
    d_fused = torch.zeros(S * T * H * rec, dtype=torch.uint8, device=dev)
    d_pred = torch.zeros_like(d_fused)
    d_nout = torch.zeros(S * T, dtype=torch.int32, device=dev)
    d_delay = torch.zeros(S * T, dtype=torch.float32, device=dev)
    g = api.PriorTracker(prm, S)
    st = torch.cuda.current_stream().cuda_stream
    g.run_device(S, T, H, d_p.data_ptr(), d_n.data_ptr(), d_s.data_ptr(), C, d_f.data_ptr(), d_fused.data_ptr(),
                 d_pred.data_ptr(), d_nout.data_ptr(), d_delay.data_ptr(), 0, st)
    torch.cuda.synchronize()
    n_out = d_nout.cpu().numpy().reshape(S, T)
    assert np.array_equal(n_out, host["n_out"])
    fused = d_fused.cpu().numpy().view(person_cov_dtype).reshape(S, T, H)
    live = np.arange(H)[None, None, :] < n_out[:, :, None]
    assert fused[live].tobytes() == host["fused"][live].tobytes()
    assert g.last_kernel_ms() > 0


def test_prior_capacity_error():
    seq = synth_person_sequences(1, 3, 5, seed=25, person_dropout=0.0)
    g = api.PriorTracker(default_prior_params(), 1, max_tracks=3)
    from smartedgesensor3dhumanpose_b200.lib import Ses3dError
    with pytest.raises(Ses3dError) as e:
        g.run(seq["persons"], seq["n_persons"], seq["stamp_ns"], None)
    assert e.value.code == -3


def test_demo_chain_on_gpu():
    """2-D detections -> skeleton_3d -> pose_prior -> pose_reprojection through the C ABI against the oracle chain."""
    from oracle.binding import Oracle
    from tests import helpers
    from tests.test_demo_chain import check_chain_parity, check_chain_physics
    S, T = 6, 28
    fr = helpers.make_sequence_workload("hall16", S, T, 4)
    prm = default_prior_params()
    ref = helpers.run_demo_chain(Oracle(fr["cameras"], ref_hungarian=True), PriorOracle(prm, S, ref_hungarian=True), fr)
    dev = helpers.run_demo_chain(api.GeometryPipeline(fr["cameras"]), api.PriorTracker(prm, S), fr)
    check_chain_parity(fr, ref, dev)
    check_chain_physics(fr, dev)


@pytest.mark.parametrize("case", [c[0] for c in __import__("scripts.make_golden_prior", fromlist=["CASES"]).CASES])
def test_prior_against_committed_golden_vectors(case):
    """The committed fixtures (tests/golden/golden_prior_v1.npz) through the C ABI: nothing here needs /root/reference."""
    from tests.test_pose_prior import _prior_golden
    mg, g = _prior_golden()
    c = next(x for x in mg.CASES if x[0] == case)
    seq, r = mg.run_case(c, make=lambda prm, S: api.PriorTracker(prm, S))
    mg.compare(g, case, seq, r, POS_TOL, COV_RTOL)


def test_prior_ragged_call_equals_padded():
    """ses3d_prior_run_ragged (dense records in and out) == ses3d_prior_run on the same streams."""
    S, T = 12, 30
    seq = synth_person_sequences(S, T, 5, seed=26, person_dropout=0.2)
    prm = default_prior_params(min_num_obs_track=3)
    H = seq["h_max"]
    padded = api.PriorTracker(prm, S).run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"])
    live_in = np.arange(H)[None, None, :] < seq["n_persons"][:, :, None]
    dense_in = np.ascontiguousarray(seq["persons"][live_in])
    cap = int(seq["n_persons"].sum())
    fused = np.zeros(cap, person_cov_dtype)
    pred = np.zeros(cap, person_cov_dtype)
    n_out, delay, total = api.PriorTracker(prm, S).run_ragged(dense_in, seq["n_persons"], seq["stamp_ns"], H, fused, pred,
                                                              seq["fb_delay"])
    assert np.array_equal(n_out, padded["n_out"]) and np.array_equal(delay, padded["pred_delay"])
    assert total == padded["n_out"].sum() and total > S * 10
    live = np.arange(H)[None, None, :] < padded["n_out"][:, :, None]
    assert fused[:total].tobytes() == padded["fused"][live].tobytes()
    assert pred[:total].tobytes() == padded["pred"][live].tobytes()
    # capacity error: an output array that is too small
    from smartedgesensor3dhumanpose_b200.lib import Ses3dError
    with pytest.raises(Ses3dError) as e:
        api.PriorTracker(prm, S).run_ragged(dense_in, seq["n_persons"], seq["stamp_ns"], H, fused[:5], pred[:5], seq["fb_delay"])
    assert e.value.code == -3
