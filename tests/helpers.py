"""Shared test helpers: workloads and result comparison."""
import numpy as np

from smartedgesensor3dhumanpose_b200 import rigs, synth

from smartedgesensor3dhumanpose_b200.workloads import CONFIGS, make_workload  # noqa: E402,F401  (re-exported)


def inject_outliers(fr, frac=0.05, seed=0, shift_px=150.0):
    """Move a fraction of the confident keypoints far away (gross outliers -> rejection branches S3D:748-838)."""
    rng = np.random.default_rng(seed)
    kp = fr["persons"]["keypoints"]
    m = (kp["score"] >= 0.5) & (rng.random(kp["score"].shape) < frac)
    kp["x"][m] += shift_px * rng.choice([-1.0, 1.0], size=m.sum()).astype(np.float32)
    kp["y"][m] += shift_px * rng.choice([-1.0, 1.0], size=m.sum()).astype(np.float32)
    return int(m.sum())


def compare_persons3d(ref, got, pos_tol, cov_rtol=1e-2, score_tol=2e-5):
    """ref/got: dicts with persons3d [F][H], n_out [F]. Returns a stats dict; raises AssertionError on mismatch."""
    assert np.array_equal(ref["n_out"], got["n_out"]), "number of output persons differs"
    F, H = ref["persons3d"].shape
    live = np.arange(H)[None, :] < ref["n_out"][:, None]
    a = ref["persons3d"]["keypoints"][live]
    b = got["persons3d"]["keypoints"][live]
    pa, pb = a["score"] > 0, b["score"] > 0
    assert np.array_equal(pa, pb), "set of triangulated joints differs"
    d = np.sqrt((a["x"] - b["x"]) ** 2 + (a["y"] - b["y"]) ** 2 + (a["z"] - b["z"]) ** 2)[pa]
    ds = np.abs(a["score"] - b["score"])[pa]
    ca, cb = a["cov"][pa], b["cov"][pa]
    scale = np.abs(ca).max(axis=-1, keepdims=True) + 1e-30
    dc = (np.abs(ca - cb) / scale).max(axis=-1) if len(ca) else np.zeros(0)
    stats = dict(n_joints=int(pa.sum()), max_pos=float(d.max()) if d.size else 0.0,
                 max_score=float(ds.max()) if ds.size else 0.0, max_cov_rel=float(dc.max()) if dc.size else 0.0)
    assert stats["max_pos"] <= pos_tol, f"joint position differs by {stats['max_pos']:.3e} m (tol {pos_tol})"
    assert stats["max_score"] <= score_tol, f"score differs by {stats['max_score']:.3e}"
    assert stats["max_cov_rel"] <= cov_rtol, f"covariance differs by {stats['max_cov_rel']:.3e} (relative)"
    return stats


def compare_persons2d(ref, got, px_tol=0.0):
    """Reprojection outputs: counts exact; keypoints exact when px_tol == 0 else within px_tol pixels."""
    assert np.array_equal(ref["n_out"], got["n_out"]), "per-camera person counts differ"
    F, C, H = ref["persons2d"].shape
    live = np.arange(H)[None, None, :] < ref["n_out"][:, :, None]
    a, b = ref["persons2d"][live], got["persons2d"][live]
    if px_tol == 0.0:
        assert a.tobytes() == b.tobytes(), "reprojected Person2D records differ bitwise"
        return dict(n_persons=int(live.sum()), max_px=0.0)
    ka, kb = a["keypoints"], b["keypoints"]
    assert np.array_equal(ka["score"] > 0, kb["score"] > 0)
    d = max(np.abs(ka["x"] - kb["x"]).max(initial=0), np.abs(ka["y"] - kb["y"]).max(initial=0))
    assert d <= px_tol, f"reprojected keypoints differ by {d} px"
    return dict(n_persons=int(live.sum()), max_px=float(d))


def make_sequence_workload(rig, n_sequences, n_frames, n_people, seed=7, fps=30.0, step_m=1.0 / 30.0, **over):
    """Temporally coherent 2-D detections (synthetic generator in sequence mode): frames [q T, (q+1) T) are stream q.
    Returns the frame arrays plus the per-message stamps of the demo chain skeleton_3d -> pose_prior -> reprojection."""
    cams = rigs.RIGS[rig]()
    cfg = synth.synth_config(seed=seed, n_people=n_people, dropout=over.get("dropout", 0.0),
                             noise_px=over.get("noise_px", 2.0), area=over.get("area", rigs.AREAS[rig]),
                             frames_per_sequence=n_frames, step_m=step_m)
    fr = synth.synth_frames(cams, cfg, n_sequences * n_frames)
    fr["cameras"] = cams
    fr["h_max"] = over.get("h_max", max(8, 2 * n_people + 4))
    t = 1000.0 + np.arange(n_frames) / fps
    fr["stamp_ns"] = np.broadcast_to(np.round(t * 1e9).astype(np.int64), (n_sequences, n_frames)).copy()
    fr["n_sequences"], fr["n_frames_per_sequence"] = n_sequences, n_frames
    return fr


def run_demo_chain(tri, prior, fr):
    """The demo wiring (pose_prior/launch/pose_triangulate_demo.launch): persons_3d -> pose_prior ->
    persons3d_fused_pred -> pose_reprojection. tri needs triangulate_batch / reproject_batch, prior needs run."""
    S, T, H = fr["n_sequences"], fr["n_frames_per_sequence"], fr["h_max"]
    r3 = tri.triangulate_batch(fr["persons"], fr["n_persons"], H)
    rp = prior.run(r3["persons3d"].reshape(S, T, H), r3["n_out"].reshape(S, T), fr["stamp_ns"], None)
    r2 = tri.reproject_batch(rp["pred"].reshape(S * T, H), rp["n_out"].reshape(S * T))
    return r3, rp, r2
