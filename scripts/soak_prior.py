#!/usr/bin/env python
"""Large-sample parity soak of the pose_prior stage: GPU (through the C ABI) against the CPU oracle (reference's verbatim
Hungarian.cpp) on long message streams with people entering / leaving, drop-outs and time gaps, to catch rare
assignment flips, track life-cycle differences or LM branch flips. Prints one JSON line per run."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle.binding import PriorOracle  # noqa: E402
from smartedgesensor3dhumanpose_b200 import api  # noqa: E402
from smartedgesensor3dhumanpose_b200.layouts import default_prior_params  # noqa: E402
from smartedgesensor3dhumanpose_b200.sequences import synth_person_sequences  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
a = ap.parse_args()
RUNS = [  # label, streams, messages, people, generator kwargs, params
    ("default", 400, 150, 5, {}, {}),
    ("dropouts + people leaving", 400, 150, 6, dict(joint_dropout=0.3, person_dropout=0.2), {}),
    ("noisy (5 cm), crowded area", 300, 120, 8, dict(noise_m=0.05, area=3.0, h_max=10), {}),
    ("h36m, height-normalised", 300, 120, 4, dict(pose_method=1), dict(pose_method=1, normalize_by_height=1)),
]
for label, S, T, P, gkw, pkw in RUNS:
    S = max(8, int(S * a.scale))
    t0 = time.time()
    seq = synth_person_sequences(S, T, P, seed=4242, **gkw)
    rng = np.random.default_rng(1)
    gap = rng.integers(20, T - 5, S)                      # a 1.5 s gap somewhere in every stream: tracks are pruned
    for s in range(S):
        seq["stamp_ns"][s, gap[s]:] += int(1.5e9)
    prm = default_prior_params(**pkw)
    ro = PriorOracle(prm, S, ref_hungarian=True).run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"],
                                                      n_threads=16)
    rg = api.PriorTracker(prm, S).run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"])
    H = ro["fused"].shape[-1]
    msgs = S * T
    cm = (ro["n_out"] != rg["n_out"])
    tm = (ro["track_of"] != rg["track_of"]).any(-1)
    live = (np.arange(H)[None, None, :] < np.minimum(ro["n_out"], rg["n_out"])[:, :, None]) & ~(cm | tm)[:, :, None]
    ka, kb = ro["fused"][live]["keypoints"], rg["fused"][live]["keypoints"]
    ids_differ = int((ro["fused"][live]["id"] != rg["fused"][live]["id"]).sum())
    m = (ka["score"] > 0) & (kb["score"] > 0)
    d = np.sqrt(sum((ka[c] - kb[c]) ** 2 for c in "xyz"))[m]
    pa, pb = ro["pred"][live]["keypoints"], rg["pred"][live]["keypoints"]
    dp = np.sqrt(sum((pa[c] - pb[c]) ** 2 for c in "xyz"))[m]
    scale = np.abs(ka["cov"]).max(axis=-1) + 1e-30
    dc = (np.abs(ka["cov"] - kb["cov"]).max(axis=-1) / scale)[m]
    print(json.dumps(dict(run=label, streams=S, messages=msgs, fits=int(seq["n_persons"].sum()),
                          published=int(ro["n_out"].sum()), count_mismatch_messages=int(cm.sum()),
                          track_assignment_mismatch_messages=int(tm.sum()), id_mismatch_records=ids_differ,
                          joint_set_mismatch=int(((ka["score"] > 0) != (kb["score"] > 0)).sum()),
                          joints=int(m.sum()), fused_max_m=float(d.max(initial=0)), fused_over_1e6_m=int((d > 1e-6).sum()),
                          pred_max_m=float(dp.max(initial=0)), cov_max_rel=float(dc.max(initial=0)),
                          seconds=round(time.time() - t0, 1))), flush=True)
