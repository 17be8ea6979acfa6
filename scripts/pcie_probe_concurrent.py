"""Host<->device link ceiling with ALL ranks copying at once (launch with torchrun, one rank per GPU): what the box
gives N processes that move bench.py's e2e traffic concurrently. Each rank binds to its GPU's NUMA node first, like
bench.py does. Rank 0 prints one JSON line: per-rank and aggregate GB/s for upload only, download only, both ways, and
the bench's own mix (269 MB up + 401 MB down per step) with the step time that mix alone would take."""
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from smartedgesensor3dhumanpose_b200 import lib as _lib  # noqa: E402

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
node = _lib.bind_thread_to_device_numa(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

UP, DOWN = 269_499_592, 400_815_088   # bytes per 16 384-frame step of the ragged call (profiles/bench_r02.json)
hu = torch.empty(UP, dtype=torch.uint8).pin_memory()
hd = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
du = torch.empty(UP, dtype=torch.uint8, device="cuda")
dd = torch.empty(DOWN, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    if world > 1:   # the slowest rank defines the step
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    return dt


def up():
    with torch.cuda.stream(s1):
        du.copy_(hu, non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        hd.copy_(dd, non_blocking=True)


def both():
    up()
    down()


t_up, t_down, t_both = timed(up), timed(down), timed(both)
nodes = [node]
if world > 1:
    g = [None] * world
    dist.all_gather_object(g, node)
    nodes = g
if rank == 0:
    print(json.dumps({
        "n_gpus": world, "numa_node_per_rank": nodes,
        "upload_only_gbs_per_gpu": UP / t_up / 1e9, "download_only_gbs_per_gpu": DOWN / t_down / 1e9,
        "mix_ms_per_step": t_both * 1e3,
        "mix_gbs_per_gpu": {"up": UP / t_both / 1e9, "down": DOWN / t_both / 1e9},
        "mix_aggregate_gbs": (UP + DOWN) * world / t_both / 1e9,
        "mix_frames_per_sec_ceiling": 16384 * world / t_both,
        "note": "slowest rank; all ranks copy concurrently; mix = bench.py's e2e bytes per 16 384-frame step"}))
if world > 1:
    dist.destroy_process_group()
