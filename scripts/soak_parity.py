#!/usr/bin/env python
"""Large-sample parity soak: GPU (through the C ABI) against the CPU oracle on many frames per config, to catch
rare association flips or near-threshold branch differences. Prints one JSON line per run."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle.binding import Oracle  # noqa: E402
from smartedgesensor3dhumanpose_b200 import api  # noqa: E402
from smartedgesensor3dhumanpose_b200.layouts import default_params  # noqa: E402
from tests import helpers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--eps", type=float, default=1e-4, help="relative half-width of the branch-threshold band")
ap.add_argument("--threads", type=int, default=0)
a = ap.parse_args()
import os  # noqa: E402
THREADS = a.threads or (os.cpu_count() or 8)
RUNS = [("cfg2_hall16x6", 200000, 0.0, {}), ("cfg3_hall16x6_dropout", 200000, 0.0, {}), ("cfg5_ring8x4", 200000, 0.0, {}),
        ("cfg1_ring4x1", 400000, 0.0, {}), ("dense_ring16x6", 20000, 0.0, {}), ("cfg2_hall16x6", 60000, 0.05, {}),
        ("cfg5_ring8x4", 60000, 0.05, {}), ("cfg2_hall16x6", 60000, 0.0, {"precision": 1})]
for name, n, outl, prm in RUNS:
    n = max(1000, int(n * a.scale))
    t0 = time.time()
    tot = dict(frames=0, assoc_mismatch_frames=0, count_mismatch_frames=0, joint_set_mismatch_frames=0, pos_fail_frames=0,
               max_pos=0.0, max_cov_rel=0.0, joints=0, eps_band_frames=0, mismatch_frames_outside_band=0,
               score_fail_frames_outside_band=0, far_joints=0, far_joints_not_bit_exact=0, offenders=[])
    chunk = 20000
    params = default_params(**prm)
    tol = 1e-4 if prm.get("precision") else 1e-3
    gpu = orc = None
    for f0 in range(0, n, chunk):
        nf = min(chunk, n - f0)
        fr = helpers.make_workload(name, nf, first_frame=f0, h_max=40)
        if outl:
            helpers.inject_outliers(fr, outl, seed=f0)
        if gpu is None:
            gpu, orc = api.GeometryPipeline(fr["cameras"], params), Oracle(fr["cameras"], params, ref_hungarian=True)
        rg = gpu.triangulate_batch(fr["persons"], fr["n_persons"], 40)
        ro = orc.triangulate_batch(fr["persons"], fr["n_persons"], 40, n_threads=THREADS, diag=True)
        tot["frames"] += nf
        tot["assoc_mismatch_frames"] += int((ro["hyp_of"] != rg["hyp_of"]).reshape(nf, -1).any(1).sum())
        cm = ro["n_out"] != rg["n_out"]
        tot["count_mismatch_frames"] += int(cm.sum())
        live = (np.arange(40)[None, :] < np.minimum(ro["n_out"], rg["n_out"])[:, None]) & ~cm[:, None]
        ka, kb = ro["persons3d"]["keypoints"], rg["persons3d"]["keypoints"]
        pa, pb = (ka["score"] > 0) & live[..., None], (kb["score"] > 0) & live[..., None]
        tot["joint_set_mismatch_frames"] += int((pa != pb).reshape(nf, -1).any(1).sum())
        both = pa & pb
        d = np.sqrt((ka["x"] - kb["x"]) ** 2 + (ka["y"] - kb["y"]) ** 2 + (ka["z"] - kb["z"]) ** 2)
        d = np.where(both, d, 0.0)
        tot["pos_fail_frames"] += int((d > tol).reshape(nf, -1).any(1).sum())
        # joints whose own UT covariance says "metres of uncertainty" (two nearly parallel rays) are at the limit of
        # FP32 in the reference as well: count separately the failures that exceed 1e-5 of the joint's 1-sigma
        sig = np.sqrt(np.maximum(ka["cov"][..., 0] + ka["cov"][..., 3] + ka["cov"][..., 5], 0.0))
        tot["pos_fail_frames_beyond_1e-5_sigma"] = tot.get("pos_fail_frames_beyond_1e-5_sigma", 0) + int(
            ((d > tol) & (d > 1e-5 * sig)).reshape(nf, -1).any(1).sum())
        ok = d <= tol
        tot["max_pos"] = max(tot["max_pos"], float(d[ok].max(initial=0.0)))
        scale = np.abs(ka["cov"]).max(-1) + 1e-30
        dc = np.where(both & ok, np.abs(ka["cov"] - kb["cov"]).max(-1) / scale, 0.0)
        tot["max_cov_rel"] = max(tot["max_cov_rel"], float(np.nanmax(dc, initial=0.0)))
        tot["joints"] += int(both.sum())
        # explicit eps-band: frames whose closest branch decision (S3D:748/775/793/813/943/964/988) is within --eps of its
        # threshold may legitimately differ between two float implementations; every other frame must agree
        band = ro["margin"] < a.eps
        ds = np.where(both, np.abs(ka["score"] - kb["score"]), 0.0)
        bad = cm | (pa != pb).reshape(nf, -1).any(1) | (d > tol).reshape(nf, -1).any(1)
        sbad = (ds > 2e-5).reshape(nf, -1).any(1)
        tot["eps_band_frames"] += int(band.sum())
        tot["mismatch_frames_outside_band"] += int((bad & ~band).sum())
        tot["score_fail_frames_outside_band"] += int((sbad & ~bad & ~band).sum())
        far = both & (ka["x"] ** 2 + ka["y"] ** 2 + ka["z"] ** 2 > 21.0 ** 2) & ~band[:, None, None]
        tot["far_joints"] += int(far.sum())
        tot["far_joints_not_bit_exact"] += int((far & ((ka["x"] != kb["x"]) | (ka["y"] != kb["y"]) | (ka["z"] != kb["z"]))).sum())
        for f in np.nonzero((bad | sbad) & ~band)[0][:20]:
            tot["offenders"].append(dict(frame=int(f0 + f), max_pos=float(d[f].max()), max_score=float(ds[f].max()),
                                         margin=float(ro["margin"][f]), count_mismatch=bool(cm[f])))
    tot.update(config=name, outliers=outl, params=prm, seconds=round(time.time() - t0, 1), pos_tol=tol)
    print(json.dumps(tot), flush=True)
