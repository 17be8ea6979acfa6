#!/usr/bin/env python
"""CPU-only parity soak: the device algorithms compiled for the host (tests/hostsim, serial team) against the CPU
oracle, many frames per config, frame-parallel over worker processes. Lists every frame whose joints differ by more
than the tolerance together with the oracle's conditioning / branch-margin diagnostics, so that offenders found here
(or on the GPU by scripts/soak_parity.py) can be promoted into fixed-slice tests.

    python scripts/soak_hostsim.py --config cfg3_hall16x6_dropout --frames 200000 [--outliers 0.05] [--procs 8]
"""
import argparse
import json
import sys
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def run_slice(args):
    name, f0, nf, outl, prm, tol, variant = args
    from oracle.binding import Oracle
    from smartedgesensor3dhumanpose_b200.layouts import default_params
    from tests import helpers
    from tests.hostsim.binding import HostSim
    fr = helpers.make_workload(name, nf, first_frame=f0, h_max=40)
    if outl:
        helpers.inject_outliers(fr, outl, seed=f0)
    params = default_params(**prm)
    sim = HostSim(fr["cameras"], params)
    orc = Oracle(fr["cameras"], params, ref_hungarian=True, svd_variant=variant)
    rg = sim.triangulate_batch(fr["persons"], fr["n_persons"], 40)
    ro = orc.triangulate_batch(fr["persons"], fr["n_persons"], 40, diag=True)
    out = dict(frames=nf, assoc=int((ro["hyp_of"] != rg["hyp_of"]).reshape(nf, -1).any(1).sum()), offenders=[])
    cm = ro["n_out"] != rg["n_out"]
    live = (np.arange(40)[None, :] < np.minimum(ro["n_out"], rg["n_out"])[:, None]) & ~cm[:, None]
    ka, kb = ro["persons3d"]["keypoints"], rg["persons3d"]["keypoints"]
    pa, pb = (ka["score"] > 0) & live[..., None], (kb["score"] > 0) & live[..., None]
    setm = (pa != pb).reshape(nf, -1).any(1)
    both = pa & pb
    d = np.sqrt((ka["x"] - kb["x"]) ** 2 + (ka["y"] - kb["y"]) ** 2 + (ka["z"] - kb["z"]) ** 2)
    d = np.where(both, d, 0.0)
    ds = np.where(both, np.abs(ka["score"] - kb["score"]), 0.0)
    dmax = d.reshape(nf, -1).max(1)
    rr = np.sqrt(ka["x"] ** 2 + ka["y"] ** 2 + ka["z"] ** 2).reshape(nf, -1)
    r_at = rr[np.arange(nf), d.reshape(nf, -1).argmax(1)]
    smax = ds.reshape(nf, -1).max(1)
    scale = np.abs(ka["cov"]).max(-1) + 1e-30
    dc = np.where(both, np.abs(ka["cov"] - kb["cov"]).max(-1) / scale, 0.0)
    cmax = np.nan_to_num(dc.reshape(nf, -1).max(1))
    bad = cm | setm | (dmax > tol) | (smax > 2e-5)
    for f in np.nonzero(bad)[0]:
        out["offenders"].append(dict(frame=int(f0 + f), count_mismatch=bool(cm[f]), set_mismatch=bool(setm[f]),
                                     max_pos=float(dmax[f]), dist_of_worst_joint=float(r_at[f]), max_score=float(smax[f]), margin=float(ro["margin"][f]),
                                     cond=float(ro["cond"][f])))
    out["max_pos_ok"] = float(dmax[~bad].max(initial=0.0))
    out["max_cov_ok"] = float(cmax[~bad].max(initial=0.0))
    out["joints"] = int(both.sum())
    out["cond_hist"] = np.histogram(ro["cond"], bins=[0, 10, 30, 100, 300, 1000, 3000, 1e4, 1e9])[0].tolist()
    out["margin_lt_1e-4"] = int((ro["margin"] < 1e-4).sum())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3_hall16x6_dropout")
    ap.add_argument("--frames", type=int, default=20000)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--outliers", type=float, default=0.0)
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--chunk", type=int, default=2000)
    ap.add_argument("--fp64", action="store_true")
    ap.add_argument("--lm", action="store_true")
    ap.add_argument("--svd-variant", type=int, default=0)
    a = ap.parse_args()
    prm = {}
    if a.fp64:
        prm["precision"] = 1
    if a.lm:
        prm["lm_refine"] = 1
    tol = 1e-4 if a.fp64 else 1e-3
    # chunks aligned like scripts/soak_parity.py (20000-frame chunks seeded by their first frame) when outliers are
    # injected, so that frame numbers are comparable between the two scripts
    jobs = [(a.config, f0, min(a.chunk, a.first + a.frames - f0), a.outliers, prm, tol, a.svd_variant)
            for f0 in range(a.first, a.first + a.frames, a.chunk)]
    tot = dict(config=a.config, frames=0, assoc=0, joints=0, offenders=[], max_pos_ok=0.0, max_cov_ok=0.0,
               cond_hist=None, margin_lt_1e_4=0)
    with ProcessPoolExecutor(a.procs) as ex:
        for r in ex.map(run_slice, jobs):
            tot["frames"] += r["frames"]
            tot["assoc"] += r["assoc"]
            tot["joints"] += r["joints"]
            tot["offenders"] += r["offenders"]
            tot["max_pos_ok"] = max(tot["max_pos_ok"], r["max_pos_ok"])
            tot["max_cov_ok"] = max(tot["max_cov_ok"], r["max_cov_ok"])
            tot["margin_lt_1e_4"] += r["margin_lt_1e-4"]
            tot["cond_hist"] = r["cond_hist"] if tot["cond_hist"] is None else [x + y for x, y in zip(tot["cond_hist"], r["cond_hist"])]
    tot["n_offenders"] = len(tot["offenders"])
    print(json.dumps(tot))


if __name__ == "__main__":
    main()
