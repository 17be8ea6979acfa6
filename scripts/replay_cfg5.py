#!/usr/bin/env python
"""BASELINE config 5: offline batch replay of N (default 1e7) synthetic frames, 8 cameras x 4 people, frame-sharded
across the ranks of one box. Frames are generated ON THE DEVICE chunk by chunk with the counter-based generator
(shipping 1e7 frames = 137 GB over PCIe would dwarf the compute), processed device-resident
(associate + triangulate + finalize + reproject), and only summary statistics are reduced at the end.

    python scripts/replay_cfg5.py --frames 10000000                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/replay_cfg5.py --frames 10000000
"""
import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from smartedgesensor3dhumanpose_b200 import api, lib, rigs, sharding, synth  # noqa: E402
from smartedgesensor3dhumanpose_b200.layouts import person2d_dtype, person_cov_dtype  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=10_000_000)
ap.add_argument("--chunk", type=int, default=65536)
a = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
cams = rigs.ring8()
cfg = synth.synth_config(seed=5, n_people=4, dropout=0.05, area=rigs.AREAS["ring8"])
Cn, PM, h_max, B = 8, 4, 12, a.chunk
lo, hi = sharding.shard_range(a.frames, rank, world)
pipe = api.GeometryPipeline(cams, device=local)
L = lib.load()
d_persons = torch.empty(B * Cn * PM * person2d_dtype.itemsize, dtype=torch.uint8, device=dev)
d_np = torch.empty(B * Cn, dtype=torch.int32, device=dev)
d3 = torch.zeros(B * h_max * person_cov_dtype.itemsize, dtype=torch.uint8, device=dev)
dn3 = torch.zeros(B, dtype=torch.int32, device=dev)
d2 = torch.zeros(B * Cn * h_max * person2d_dtype.itemsize, dtype=torch.uint8, device=dev)
dn2 = torch.zeros(B * Cn, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream()
gen_ms = proc_ms = 0.0
persons_out = joints = reproj = 0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t_all0 = torch.cuda.Event(enable_timing=True); t_all1 = torch.cuda.Event(enable_timing=True)
t_all0.record()
for f0 in range(lo, hi, B):
    nf = min(B, hi - f0)
    ev[0].record()
    rc = L.ses3d_synth_frames_device(Cn, cams.ctypes.data, C.byref(cfg), f0, nf, d_persons.data_ptr(), d_np.data_ptr(), None,
                                     st.cuda_stream)
    assert rc == 0
    ev[1].record()
    pipe.process_device(nf, PM, h_max, d_persons.data_ptr(), d_np.data_ptr(), d3.data_ptr(), dn3.data_ptr(), d2.data_ptr(),
                        dn2.data_ptr(), stream=st.cuda_stream)
    ev[2].record()
    torch.cuda.synchronize()
    gen_ms += ev[0].elapsed_time(ev[1]); proc_ms += ev[1].elapsed_time(ev[2])
    comp = sharding.compact_torch(d3[:nf * h_max * person_cov_dtype.itemsize], nf, h_max)
    live = (torch.arange(h_max, device=dev)[None, :] < dn3[:nf, None])
    joints += int(((comp[..., 3] > 0) & live[..., None]).sum().item())
    persons_out += int(dn3[:nf].sum().item())
    reproj += int(dn2[:nf * Cn].sum().item())
t_all1.record()
torch.cuda.synchronize()
wall_ms = t_all0.elapsed_time(t_all1)
stats = torch.tensor([hi - lo, persons_out, joints, reproj], dtype=torch.float64, device=dev)
times = torch.tensor([gen_ms, proc_ms, wall_ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(stats)                       # the only collective: summary statistics
    dist.all_reduce(times, op=dist.ReduceOp.MAX)
if rank == 0:
    f, p, j, r = stats.tolist()
    g, pr, w = times.tolist()
    print(json.dumps({"config": "cfg5 ring8 x 4, 5% dropout, seed 5", "frames": int(f), "n_gpus": world, "chunk": B,
                      "input": "generated on device per chunk (counter-based Philox generator, bit-identical to the host one)",
                      "persons3d": int(p), "joints": int(j), "reprojected_persons2d": int(r),
                      "generate_ms_max_rank": g, "process_ms_max_rank": pr, "wall_ms_max_rank": w,
                      "frames_per_sec_processing": f / (pr * 1e-3), "joints_per_sec_processing": j / (pr * 1e-3),
                      "frames_per_sec_incl_generation_and_stats": f / (w * 1e-3),
                      "result_gather": "summary statistics only (all_reduce of 4 doubles); full PersonCov stays on the device"}))
if world > 1:
    dist.destroy_process_group()
