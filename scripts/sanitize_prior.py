#!/usr/bin/env python
"""Small pose_prior run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_prior.py
Covers track creation / pruning / merging, drop-outs, more detections than one fit group, both pose methods."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from smartedgesensor3dhumanpose_b200 import api  # noqa: E402
from smartedgesensor3dhumanpose_b200.layouts import default_prior_params  # noqa: E402
from smartedgesensor3dhumanpose_b200.sequences import synth_person_sequences  # noqa: E402

for pm, nh, people in [(0, 0, 7), (1, 1, 3)]:
    seq = synth_person_sequences(3, 14, people, seed=41 + pm, joint_dropout=0.2, person_dropout=0.15, pose_method=pm,
                                 h_max=10)
    seq["stamp_ns"][2, 8:] += int(3e9)          # a gap: tracks of stream 2 are pruned and re-created
    trk = api.PriorTracker(default_prior_params(pose_method=pm, normalize_by_height=nh, min_num_obs_track=2), 3)
    r = trk.run(seq["persons"], seq["n_persons"], seq["stamp_ns"], seq["fb_delay"])
    print("pose_method", pm, "published", int(r["n_out"].sum()), "tracks", [len(trk.tracks(s)[0]) for s in range(3)])
    trk.close()
    # ragged call (pack / unpack kernels, pipelined slots)
    from smartedgesensor3dhumanpose_b200.layouts import person_cov_dtype
    H = seq["h_max"]
    live = np.arange(H)[None, None, :] < seq["n_persons"][:, :, None]
    dense = np.ascontiguousarray(seq["persons"][live])
    fused, pred = np.zeros(len(dense), person_cov_dtype), np.zeros(len(dense), person_cov_dtype)
    trk = api.PriorTracker(default_prior_params(pose_method=pm, normalize_by_height=nh, min_num_obs_track=2), 3)
    n_out, _, total = trk.run_ragged(dense, seq["n_persons"], seq["stamp_ns"], H, fused, pred, seq["fb_delay"])
    assert total == r["n_out"].sum() and np.array_equal(n_out, r["n_out"])
    trk.close()
    # streaming: one message per call (the contiguous small-call record) equals the batched launch
    trk = api.PriorTracker(default_prior_params(pose_method=pm, normalize_by_height=nh, min_num_obs_track=2), 3)
    for t in range(seq["persons"].shape[1]):
        rs = trk.run(seq["persons"][:, t:t + 1], seq["n_persons"][:, t:t + 1], seq["stamp_ns"][:, t:t + 1],
                     seq["fb_delay"][:, t:t + 1])
        assert rs["fused"][:, 0].tobytes() == r["fused"][:, t].tobytes()
    trk.close()
# visualisation kernel on fused skeletons
from smartedgesensor3dhumanpose_b200 import rigs  # noqa: E402
pipe = api.GeometryPipeline(rigs.ring8())
m = pipe.markers_batch(r["fused"].reshape(-1, r["fused"].shape[-1]), r["n_out"].reshape(-1), 1)
print("markers: segments", int(m["n_segments"].sum()))
print("sanitize prior done")
