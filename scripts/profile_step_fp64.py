#!/usr/bin/env python
"""Device-resident steps of config 3 in FP64 mode (LM refinement on) for ncu captures of k_triangulate<double>."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from smartedgesensor3dhumanpose_b200 import api  # noqa: E402
from smartedgesensor3dhumanpose_b200.layouts import default_params, person_cov_dtype  # noqa: E402
from tests import helpers  # noqa: E402

B = 8192
fr = helpers.make_workload("cfg3_hall16x6_dropout", B)
pipe = api.GeometryPipeline(fr["cameras"], default_params(precision=1, lm_refine=1))
C, PM, h_max = fr["persons"].shape[1], fr["persons"].shape[2], fr["h_max"]
dev = torch.device("cuda:0")
d_persons = torch.from_numpy(fr["persons"].view(np.uint8).reshape(-1)).to(dev)
d_np = torch.from_numpy(fr["n_persons"]).to(dev)
d3 = torch.zeros(B * h_max * person_cov_dtype.itemsize, dtype=torch.uint8, device=dev)
dn3 = torch.zeros(B, dtype=torch.int32, device=dev)
for _ in range(3):
    pipe.triangulate_device(B, PM, h_max, d_persons.data_ptr(), d_np.data_ptr(), d3.data_ptr(), dn3.data_ptr())
torch.cuda.synchronize()
print("done")
