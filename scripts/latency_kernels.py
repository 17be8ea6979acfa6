"""Per-kernel device time of single-frame host calls (CUPTI via torch.profiler): which kernel the per-message latency
of the ROS nodes sits in. Diagnostic, run on a GPU box:  python scripts/latency_kernels.py [workload]"""
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    from torch.profiler import ProfilerActivity, profile
    from smartedgesensor3dhumanpose_b200 import api, workloads
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_hall16x6"
    fr = workloads.make_workload(name, 128)
    pipe = api.GeometryPipeline(fr["cameras"], device=0)
    h_max = fr["h_max"]
    for f in range(16):
        pipe.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], h_max)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for f in range(16, 80):
            pipe.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], h_max)
        torch.cuda.synchronize()
    dur = defaultdict(list)
    spans = []
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                 key=lambda e: e.time_range.start)
    for e in evs:
        dur[e.name.split("(")[0].split("<")[0]].append(e.time_range.elapsed_us())
    n_calls = 64
    for k, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:40s} n/call {len(v) / n_calls:5.2f}  median {np.median(v):7.2f} us  sum/call {sum(v) / n_calls:7.2f} us")
    # busy span per call: first event start to last event end of each group of events (calls are serialised)
    starts = [e.time_range.start for e in evs]
    ends = [e.time_range.end for e in evs]
    per = len(evs) // n_calls
    if per * n_calls == len(evs):
        sp = [ends[i * per + per - 1] - starts[i * per] for i in range(n_calls)]
        print(f"device span per call (first activity start -> last activity end): median {np.median(sp):.1f} us")


if __name__ == "__main__":
    main()
