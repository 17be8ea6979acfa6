#!/usr/bin/env python
"""Run every BASELINE.json config (SURVEY 8(d) configs 1-5 + the dense 16-view variant) on one GPU:
parity against the oracle on a sample, device-resident throughput, per-kernel times. Writes one JSON object
per config to stdout / --out. Usage (GPU box):  python scripts/run_configs.py --out gpurun_out/configs.json"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from oracle.binding import Oracle  # noqa: E402
from smartedgesensor3dhumanpose_b200 import api  # noqa: E402
from smartedgesensor3dhumanpose_b200.layouts import default_params, person2d_dtype, person_cov_dtype  # noqa: E402
from tests import helpers  # noqa: E402

RUNS = [  # label, workload, frames, params, parity sample
    ("cfg1 4cam x 1 person", "cfg1_ring4x1", 65536, {}, 2000),
    ("cfg2 hall16 x 6", "cfg2_hall16x6", 16384, {}, 1000),
    ("cfg2-dense ring16 x 6 (16 views/joint)", "dense_ring16x6", 8192, {}, 200),
    ("cfg3 hall16 x 6, 30% dropout, LM, FP32", "cfg3_hall16x6_dropout", 16384, {"lm_refine": 1}, 1000),
    ("cfg3 hall16 x 6, 30% dropout, LM, FP64", "cfg3_hall16x6_dropout", 16384, {"lm_refine": 1, "precision": 1}, 1000),
    ("cfg4 crowd 64cam x 20", "cfg4_crowd64x20", 512, {}, 8),
    ("cfg5 ring8 x 4", "cfg5_ring8x4", 32768, {}, 1000),
]


def run(label, workload, n_frames, prm, n_parity, steps=5):
    fr = helpers.make_workload(workload, n_frames)
    params = default_params(**prm)
    pipe = api.GeometryPipeline(fr["cameras"], params)
    C, PM, h_max = fr["persons"].shape[1], fr["persons"].shape[2], fr["h_max"]
    # parity on a sample
    sub = {k: v[:n_parity] for k, v in fr.items() if isinstance(v, np.ndarray) and v.shape[:1] == (n_frames,)}
    ro = Oracle(fr["cameras"], params, ref_hungarian=True).triangulate_batch(sub["persons"], sub["n_persons"], h_max, n_threads=16)
    rg = pipe.triangulate_batch(sub["persons"], sub["n_persons"], h_max)
    idx_exact = bool(np.array_equal(ro["hyp_of"], rg["hyp_of"]) and np.array_equal(ro["n_hungarian"], rg["n_hungarian"]))
    tol = 1e-4 if prm.get("precision") else 1e-3
    try:
        st = helpers.compare_persons3d(ro, rg, tol, cov_rtol=5e-2)
        parity = dict(ok=True, **st)
    except AssertionError as e:
        parity = dict(ok=False, error=str(e))
    dev = torch.device("cuda:0")
    d_persons = torch.from_numpy(fr["persons"].view(np.uint8).reshape(-1)).to(dev)
    d_np = torch.from_numpy(fr["n_persons"]).to(dev)
    d3 = torch.zeros(n_frames * h_max * person_cov_dtype.itemsize, dtype=torch.uint8, device=dev)
    dn3 = torch.zeros(n_frames, dtype=torch.int32, device=dev)
    d2 = torch.zeros(n_frames * C * h_max * person2d_dtype.itemsize, dtype=torch.uint8, device=dev)
    dn2 = torch.zeros(n_frames * C, dtype=torch.int32, device=dev)
    st_ = torch.cuda.current_stream()

    def step():
        pipe.process_device(n_frames, PM, h_max, d_persons.data_ptr(), d_np.data_ptr(), d3.data_ptr(), dn3.data_ptr(),
                            d2.data_ptr(), dn2.data_ptr(), stream=st_.cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    pipe.set_profiling(True)
    step()
    kms = pipe.last_kernel_ms()
    pipe.set_profiling(False)
    n3 = dn3.cpu().numpy()
    out3 = d3.cpu().numpy().view(person_cov_dtype).reshape(n_frames, h_max)
    live = np.arange(h_max)[None, :] < n3[:, None]
    joints = int(((out3["keypoints"]["score"] > 0) & live[..., None]).sum())
    return dict(label=label, workload=workload, frames=n_frames, params=prm, cameras=C, p_max=PM, h_max=h_max,
                association_bit_exact=idx_exact, parity=parity, ms_per_step=ms, frames_per_sec=n_frames / ms * 1e3,
                joints_per_sec=joints / ms * 1e3, joints_per_frame=joints / n_frames, kernel_ms=kms,
                hungarian_per_frame=float(rg["n_hungarian"].mean()), persons_out_per_frame=float(n3.mean()))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    res = []
    for r in RUNS:
        t0 = time.time()
        out = run(*r)
        out["wall_s"] = time.time() - t0
        print(json.dumps(out), flush=True)
        res.append(out)
    if a.out:
        Path(a.out).write_text(json.dumps(res, indent=1))
