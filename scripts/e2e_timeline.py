"""Timeline of one host-buffer ragged call (ses3d_process_batch_ragged) from CUPTI via torch.profiler: per kernel /
memcpy kind the busy time and span, copy-engine utilisation, and the critical gaps. Diagnostic, run on a GPU box:

    python scripts/e2e_timeline.py [--frames 16384] [--workload cfg2_hall16x6] [--out gpurun_out/tl.json]
"""
import argparse
import json
import sys
import time
from collections import defaultdict
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def union_ms(iv):
    iv = sorted(iv)
    tot, cur_s, cur_e = 0.0, None, None
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        tot += cur_e - cur_s
    return tot / 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16384)
    ap.add_argument("--workload", default="cfg2_hall16x6")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--window", default="", help="lo,hi in ms: list every GPU activity that starts in the window (stderr)")
    a = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    from smartedgesensor3dhumanpose_b200 import api, workloads
    from smartedgesensor3dhumanpose_b200.layouts import person2d_dtype, person_cov_dtype

    B = a.frames
    fr = workloads.make_workload(a.workload, B)
    cams, h_max = fr["cameras"], fr["h_max"]
    C, PM = fr["persons"].shape[1], fr["persons"].shape[2]
    pipe = api.GeometryPipeline(cams, device=0)
    pipe.reserve(B, PM, h_max)

    def pinned(arr):
        t = torch.from_numpy(arr.view(np.uint8).reshape(-1)).pin_memory()
        return t, t.numpy().view(arr.dtype).reshape(arr.shape)

    pad = pipe.process_batch(fr["persons"], fr["n_persons"], h_max)
    tot3, tot2 = int(pad["n_out3d"].sum()), int(pad["n_out2d"].sum())
    keep = []
    t, dense_in = pinned(api.to_ragged(fr["persons"], fr["n_persons"])); keep.append(t)
    t, n_in = pinned(fr["n_persons"]); keep.append(t)
    t, d3 = pinned(np.zeros(tot3 + 16, person_cov_dtype)); keep.append(t)
    t, d2 = pinned(np.zeros(tot2 + 16, person2d_dtype)); keep.append(t)
    t, n3 = pinned(np.zeros(B, np.int32)); keep.append(t)
    t, n2 = pinned(np.zeros((B, C), np.int32)); keep.append(t)

    def step():
        pipe.process_batch_ragged(dense_in, n_in, PM, h_max, d3, n3, d2, n2)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    wall = []
    for _ in range(a.reps):
        t0 = time.perf_counter(); step(); wall.append((time.perf_counter() - t0) * 1e3)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        t0 = time.perf_counter(); step(); torch.cuda.synchronize(); prof_wall = (time.perf_counter() - t0) * 1e3
    copies = []
    if a.out:
        chrome = a.out.replace(".json", "_chrome.json")
        prof.export_chrome_trace(chrome)
        # per-copy list (bytes are only in the chrome trace): start, duration, size, rate - big copies only
        tr = json.load(open(chrome))
        cp = [e for e in tr.get("traceEvents", []) if e.get("cat") == "gpu_memcpy" and e.get("args", {}).get("bytes", 0) >= 1 << 20]
        if cp:
            t0c = min(e["ts"] for e in tr["traceEvents"] if e.get("cat") in ("gpu_memcpy", "kernel", "gpu_memset"))
            for e in sorted(cp, key=lambda e: e["ts"]):
                kind = "HtoD" if "HtoD" in e["name"] else "DtoH" if "DtoH" in e["name"] else "other"
                copies.append({"kind": kind, "start_ms": round((e["ts"] - t0c) / 1e3, 3), "dur_ms": round(e["dur"] / 1e3, 3),
                               "mb": round(e["args"]["bytes"] / 1e6, 2), "gbs": round(e["args"]["bytes"] / e["dur"] / 1e3, 1),
                               "stream": e["args"].get("stream")})
        if a.window:
            lo, hi = [float(x) for x in a.window.split(",")]
            for e in sorted(tr["traceEvents"], key=lambda e: e.get("ts", 0)):
                if e.get("cat") in ("gpu_memcpy", "kernel", "gpu_memset") and lo <= (e["ts"] - t0c) / 1e3 <= hi:
                    print("%8.3f %7.3f s%-3s %s %s" % ((e["ts"] - t0c) / 1e3, e["dur"] / 1e3, e["args"].get("stream"),
                                                     e["name"].split("(")[0].replace("void ses3d::", "").replace("ses3d::", "")[:28],
                                                     ("%.1fMB" % (e["args"]["bytes"] / 1e6)) if "bytes" in e.get("args", {}) else
                                                     ("grid %s" % (e["args"].get("grid"),))), file=sys.stderr)
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    by = defaultdict(list)
    t_min = min(e.time_range.start for e in evs)
    t_max = max(e.time_range.end for e in evs)
    for e in evs:
        name = e.name
        if "Memcpy" in name:
            name = "memcpy " + ("HtoD" if "HtoD" in name else "DtoH" if "DtoH" in name else "DtoD" if "DtoD" in name else name)
        elif "Memset" in name:
            name = "memset"
        else:
            name = name.split("(")[0].split("<")[0].replace("ses3d::", "")
        by[name].append((e.time_range.start, e.time_range.end))
    res = {"frames": B, "workload": a.workload, "wall_ms": wall, "profiled_wall_ms": prof_wall,
           "gpu_span_ms": (t_max - t_min) / 1e3, "rows": {}}
    kernels = []
    for name, iv in sorted(by.items(), key=lambda kv: -sum(e - s for s, e in kv[1])):
        res["rows"][name] = {"n": len(iv), "sum_ms": sum(e - s for s, e in iv) / 1e3, "busy_union_ms": union_ms(iv),
                             "first_start_ms": (min(s for s, _ in iv) - t_min) / 1e3,
                             "last_end_ms": (max(e for _, e in iv) - t_min) / 1e3}
        if not name.startswith("mem"):
            kernels += iv
    res["all_kernels_busy_union_ms"] = union_ms(kernels)
    res["copies_over_1mb"] = copies
    print(json.dumps(res, indent=1))
    if a.out:
        Path(a.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
