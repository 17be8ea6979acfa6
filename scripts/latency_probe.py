#!/usr/bin/env python
"""Single-frame call latency (the ROS-shim use case, n_frames = 1): wall-clock p50/p90 of the host-buffer calls and
the device time of each kernel."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from smartedgesensor3dhumanpose_b200 import api  # noqa: E402
from tests import helpers  # noqa: E402

import os

fr = helpers.make_workload("cfg2_hall16x6", 512)
h_max = fr["h_max"]
# the round-1 shape of the call for comparison: eager launches, throughput launch shapes
os.environ["SES3D_FRAME_GRAPH"] = "0"
os.environ["SES3D_LATENCY_FRAMES"] = "0"
old = api.GeometryPipeline(fr["cameras"])
os.environ["SES3D_LATENCY_FRAMES"] = "-1"
eager = api.GeometryPipeline(fr["cameras"])
del os.environ["SES3D_FRAME_GRAPH"], os.environ["SES3D_LATENCY_FRAMES"]
pipe = api.GeometryPipeline(fr["cameras"])
for name, fn in [("eager launches, throughput shapes: process_batch(1)", lambda f: old.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], h_max)),
                 ("eager launches, latency shapes: process_batch(1)", lambda f: eager.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], h_max)),
                 ("triangulate_batch(1)", lambda f: pipe.triangulate_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], h_max, dump=False)),
                 ("process_batch(1)", lambda f: pipe.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], h_max))]:
    lat = []
    for f in range(512):
        t0 = time.perf_counter()
        fn(f)
        lat.append(time.perf_counter() - t0)
    lat = np.array(lat[32:]) * 1e6
    print(f"{name}: p50 {np.median(lat):.1f} us  p90 {np.percentile(lat, 90):.1f} us  min {lat.min():.1f} us")
for label, p in (("throughput shapes", old), ("latency shapes", pipe)):
    p.set_profiling(True)
    acc = {}
    for f in range(64):
        p.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], h_max)
        for k, v in p.last_kernel_ms().items():
            acc.setdefault(k, []).append(v * 1e3)
    print(f"kernel device time per single-frame call, {label} (us, median):",
          {k: round(float(np.median(v)), 1) for k, v in acc.items()})
