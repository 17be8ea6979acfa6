#!/usr/bin/env python
"""Device-resident throughput of the reference's whole demo chain on one B200
(pose_prior/launch/pose_triangulate_demo.launch): 2-D detections -> skeleton_3d (associate, triangulate, finalize)
-> pose_prior (track, fit, predict) -> pose_reprojection of the predicted skeletons.

    python scripts/bench_chain.py [--streams 512 --frames 32 --rig hall16 --people 6 --steps 10 --warmup 3]

Input: temporally coherent synthetic streams (generator in sequence mode), frames [qT, (q+1)T) = stream q. All buffers
stay in HBM; one step = S x T frames through the three library calls on one CUDA stream; the trackers are reset
between steps (not timed). Prints one JSON line; rank-0-only, single GPU."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from bench import ClockSampler  # noqa: E402
from smartedgesensor3dhumanpose_b200 import api  # noqa: E402
from smartedgesensor3dhumanpose_b200.layouts import default_prior_params, person2d_dtype, person_cov_dtype  # noqa: E402
from tests import helpers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=512)
ap.add_argument("--frames", type=int, default=32)
ap.add_argument("--rig", default="hall16")
ap.add_argument("--people", type=int, default=6)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()

S, T = a.streams, a.frames
F = S * T
fr = helpers.make_sequence_workload(a.rig, S, T, a.people)
C, PM, H = fr["persons"].shape[1], fr["persons"].shape[2], fr["h_max"]
dev = torch.device("cuda:0")
tb = lambda x: torch.from_numpy(np.ascontiguousarray(x).view(np.uint8).reshape(-1)).to(dev)
d_in, d_nin, d_stamp = tb(fr["persons"]), tb(fr["n_persons"]), tb(fr["stamp_ns"])
rec3, rec2 = person_cov_dtype.itemsize, person2d_dtype.itemsize
d_3d = torch.zeros(F * H * rec3, dtype=torch.uint8, device=dev)
d_n3d = torch.zeros(F, dtype=torch.int32, device=dev)
d_fused = torch.zeros_like(d_3d)
d_pred = torch.zeros_like(d_3d)
d_npub = torch.zeros(F, dtype=torch.int32, device=dev)
d_2d = torch.zeros(F * C * H * rec2, dtype=torch.uint8, device=dev)
d_n2d = torch.zeros(F * C, dtype=torch.int32, device=dev)
pipe = api.GeometryPipeline(fr["cameras"])
pipe.reserve(F, PM, H)
prior = api.PriorTracker(default_prior_params(), S)
torch.cuda.set_stream(torch.cuda.Stream())
st = torch.cuda.current_stream().cuda_stream


def step():
    pipe.triangulate_device(F, PM, H, d_in.data_ptr(), d_nin.data_ptr(), d_3d.data_ptr(), d_n3d.data_ptr(), stream=st)
    prior.run_device(S, T, H, d_3d.data_ptr(), d_n3d.data_ptr(), d_stamp.data_ptr(), 0, 0, d_fused.data_ptr(),
                     d_pred.data_ptr(), d_npub.data_ptr(), 0, 0, st)
    pipe.reproject_device(F, H, d_pred.data_ptr(), d_npub.data_ptr(), d_2d.data_ptr(), d_n2d.data_ptr(), stream=st)


for _ in range(a.warmup):
    prior.reset()
    step()
torch.cuda.synchronize()
l0 = pipe.launch_count + prior.launch_count
times = []
with ClockSampler(0) as clk:
    time.sleep(0.6)
    for _ in range(a.steps):
        prior.reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    time.sleep(0.3)
launches = (pipe.launch_count + prior.launch_count - l0 - a.steps) // a.steps
ms = float(np.mean(times))
n3 = int(d_n3d.sum().item())
npub = int(d_npub.sum().item())
n2 = int(d_n2d.sum().item())
print(json.dumps({
    "metric": "demo_chain_frames_per_sec", "value": F / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1, "steps": a.steps,
    "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "f32 (skeleton_3d) / f64 (pose_prior, reprojection)",
    "data": "synthetic",
    "config": {"workload": f"demo chain {a.rig} x {a.people}, sequence mode", "streams": S, "frames_per_stream": T,
               "cameras": C, "h_max": H, "stages": "associate+triangulate+finalize -> pose_prior -> reproject(pred)",
               "persons3d_per_step": n3, "fused_published_per_step": npub, "reprojected_persons2d_per_step": n2,
               "l2_policy": f"inputs larger than L2 ({d_in.numel() / 2**20:.0f} MiB in)"},
    "clocks": clk.summary(), "gpu_launches": int(launches)}))
