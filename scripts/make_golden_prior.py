#!/usr/bin/env python
"""Generate tests/golden/golden_prior_v1.npz: frozen outputs of the pose_prior oracle (oracle/pose_prior_oracle.cpp with
the reference's verbatim Hungarian.cpp from oracle/_ref) on fixed-seed synthetic message streams. The reference has no
golden vectors and gtsam is absent, so these pin the oracle against regressions and give the GPU tests a
fixture that does not need /root/reference. Run in the build container:  python scripts/make_golden_prior.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle.binding import PriorOracle  # noqa: E402
from smartedgesensor3dhumanpose_b200.layouts import default_prior_params  # noqa: E402
from smartedgesensor3dhumanpose_b200.sequences import synth_person_sequences  # noqa: E402

CASES = [  # name, streams, messages, people, generator kwargs, parameter overrides
    ("simple", 2, 20, 4, {}, {}),
    ("dropouts", 2, 22, 5, dict(joint_dropout=0.3, person_dropout=0.15), {}),
    ("h36m", 1, 18, 3, dict(pose_method=1), dict(pose_method=1)),
    ("norm_height", 1, 18, 3, {}, dict(normalize_by_height=1)),
    ("publish_all", 1, 8, 7, dict(h_max=10), dict(min_num_obs_track=0)),
]
OUT = ROOT / "tests" / "golden" / "golden_prior_v1.npz"


def inputs(case):
    name, S, T, P, gkw, pkw = case
    seq = synth_person_sequences(S, T, P, seed=1000 + sum(map(ord, name)), **gkw)
    return seq, default_prior_params(**pkw)


def run_case(case, make=PriorOracle):
    seq, prm = inputs(case)
    impl = make(prm, case[1], ref_hungarian=True) if make is PriorOracle else make(prm, case[1])
    return seq, impl.run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"])


def pack(seq, r):
    H = r["fused"].shape[-1]
    live = np.arange(H)[None, None, :] < r["n_out"][:, :, None]
    out = {"input_sha256": np.frombuffer(hashlib.sha256(seq["persons"].tobytes() + seq["stamp_ns"].tobytes()).digest(), np.uint8),
           "n_out": r["n_out"].astype(np.int16), "track_of": r["track_of"].astype(np.int16), "pred_delay": r["pred_delay"]}
    for key in ("fused", "pred"):
        rec = r[key][live]
        kp = rec["keypoints"]
        out[f"{key}_id"] = rec["id"].astype(np.int16)
        out[f"{key}_xyz"] = np.stack([kp["x"], kp["y"], kp["z"]], -1)
        out[f"{key}_score"] = kp["score"]
        if key == "fused":   # the predicted covariance is the fused one + pred_noise_sigma^2 on the diagonal
            out[f"{key}_cov"] = kp["cov"]
        else:
            out["pred_cov_minus_fused"] = kp["cov"] - r["fused"][live]["keypoints"]["cov"]
    return out


def compare(golden, name, seq, r, pos_tol, cov_rtol):
    """Assert that a run reproduces the stored vectors of case `name`."""
    got = pack(seq, r)
    g = {k.split("/", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/")}
    assert np.array_equal(g["input_sha256"], got["input_sha256"]), "synthetic input changed: regenerate the golden file"
    for k in ("n_out", "track_of", "pred_delay", "fused_id", "pred_id", "fused_score", "pred_score"):
        assert np.array_equal(g[k], got[k]), k
    worst = 0.0
    for key in ("fused", "pred"):
        d = np.abs(g[f"{key}_xyz"] - got[f"{key}_xyz"]).max(initial=0)
        assert d <= pos_tol, f"{name}/{key}: joints differ by {d} m"
        worst = max(worst, d)
    scale = np.abs(g["fused_cov"]).max(axis=-1, keepdims=True) + 1e-30
    dc = (np.abs(g["fused_cov"] - got["fused_cov"]) / scale).max(initial=0)
    assert dc <= cov_rtol, f"{name}: covariances differ by {dc} (relative)"
    assert np.abs(g["pred_cov_minus_fused"] - got["pred_cov_minus_fused"]).max(initial=0) <= 1e-8   # stored as float32
    return worst


def main():
    out = {}
    for case in CASES:
        seq, r = run_case(case)
        for k, v in pack(seq, r).items():
            out[f"{case[0]}/{k}"] = v
        print(case[0], "published", int(r["n_out"].sum()))
    OUT.parent.mkdir(exist_ok=True)
    out = {k: (v.astype(np.float32) if k.endswith("pred_cov_minus_fused") else v) for k, v in out.items()}
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
