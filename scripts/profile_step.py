#!/usr/bin/env python
"""Device-resident steps only (same launch shapes as bench.py's timed `value` leg), for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:k_triangulate -s 2 -c 1 -o out python scripts/profile_step.py
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from smartedgesensor3dhumanpose_b200 import api  # noqa: E402
from smartedgesensor3dhumanpose_b200.layouts import person2d_dtype, person_cov_dtype  # noqa: E402
from tests import helpers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2_hall16x6")
ap.add_argument("--frames", type=int, default=16384)
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
fr = helpers.make_workload(a.workload, a.frames)
pipe = api.GeometryPipeline(fr["cameras"])
B, C, PM, h_max = a.frames, fr["persons"].shape[1], fr["persons"].shape[2], fr["h_max"]
dev = torch.device("cuda:0")
d_persons = torch.from_numpy(fr["persons"].view(np.uint8).reshape(-1)).to(dev)
d_np = torch.from_numpy(fr["n_persons"]).to(dev)
d3 = torch.zeros(B * h_max * person_cov_dtype.itemsize, dtype=torch.uint8, device=dev)
dn3 = torch.zeros(B, dtype=torch.int32, device=dev)
d2 = torch.zeros(B * C * h_max * person2d_dtype.itemsize, dtype=torch.uint8, device=dev)
dn2 = torch.zeros(B * C, dtype=torch.int32, device=dev)
for _ in range(a.steps):
    pipe.process_device(B, PM, h_max, d_persons.data_ptr(), d_np.data_ptr(), d3.data_ptr(), dn3.data_ptr(), d2.data_ptr(),
                        dn2.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("done", pipe.launch_count)
