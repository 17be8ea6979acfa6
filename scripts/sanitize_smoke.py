#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py
Covers the padded, ragged and device-generator paths, outlier branches (incl. far joints: the exact re-solves), FP64
mode, the crowd rig (global scratch, sliced pair list, CTA-per-frame rounds), the multi-device entry, markers and the
overlay renderer, and single-frame host calls (captured CUDA graph, low-latency launch shapes)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from smartedgesensor3dhumanpose_b200 import api  # noqa: E402
from smartedgesensor3dhumanpose_b200.layouts import default_params, person2d_dtype, person_cov_dtype  # noqa: E402
from tests import helpers  # noqa: E402

for name, n, prm, outl in [("cfg2_hall16x6", 24, {}, 0.06), ("cfg5_ring8x4", 16, {"precision": 1, "lm_refine": 1}, 0.0),
                           ("cfg3_hall16x6_dropout", 48, {}, 0.0), ("cfg4_crowd64x20", 2, {}, 0.0)]:
    fr = helpers.make_workload(name, n, h_max=40 if outl else None) if outl else helpers.make_workload(name, n)
    if outl:
        helpers.inject_outliers(fr, outl)
    pipe = api.GeometryPipeline(fr["cameras"], default_params(**prm))
    r = pipe.process_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    dense = api.to_ragged(fr["persons"], fr["n_persons"])
    C = fr["persons"].shape[1]
    o3 = np.zeros(int(r["n_out3d"].sum()) + 4, person_cov_dtype)
    o2 = np.zeros(int(r["n_out2d"].sum()) + 4, person2d_dtype)
    t3, t2 = pipe.process_batch_ragged(dense, fr["n_persons"], fr["persons"].shape[2], fr["h_max"], o3,
                                       np.zeros(n, np.int32), o2, np.zeros((n, C), np.int32))
    print(name, "persons3d", int(r["n_out3d"].sum()), "ragged totals", t3, t2)
    if name in ("cfg2_hall16x6", "cfg5_ring8x4"):
        # single-frame host calls: eager + capture, then graph replays (low-latency launch shapes: sliced pair list,
        # one warp per camera in the reprojection), all three stage masks
        for f in [0, 1, 2, 3, 1]:
            one = pipe.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], fr["h_max"])
            assert one["persons3d"][0].tobytes() == r["persons3d"][f].tobytes()
            assert one["persons2d"][0].tobytes() == r["persons2d"][f].tobytes()
            t = pipe.triangulate_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], fr["h_max"], dump=False)
            q = pipe.reproject_batch(t["persons3d"], t["n_out"])
            assert q["persons2d"][0].tobytes() == r["persons2d"][f].tobytes()
        print(name, "single-frame graph replays ok")
    if name == "cfg2_hall16x6":
        mk = pipe.markers_batch(r["persons3d"], r["n_out3d"], style=0)
        img = pipe.overlay_batch(fr["persons"][0], fr["n_persons"][0], 320, 240)
        print("markers", int(mk["n_segments"].sum()), "overlay", img.shape, int((img != 255).any(-1).sum()))
    pipe.close()
print("sanitize smoke done")
