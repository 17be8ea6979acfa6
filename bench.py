#!/usr/bin/env python
"""bench.py — headline benchmark of the multi-view geometry hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames B] [--workload NAME]

A step = one pass of association + triangulation (+UT covariance) + plausibility/merge + reprojection
over one batch of B synthetic frames per GPU (the reference's 16-camera hall rig,
6 people). `value` = joints triangulated / s over the whole job with inputs resident in HBM; `e2e` = the
same metric through the public host-buffer call (pinned host memory, H2D + D2H inside the timed region).
`--impl reference` times the CPU oracle (a port of the reference's algorithm + the reference's verbatim
Hungarian.cpp) on the host cores instead. One JSON line on stdout (rank 0); `extra` holds time-boxed measurements of
the other BASELINE.json configs (value, parity against the oracle, roofline fraction).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CPU_LABEL = "port"
METRIC = "joints_triangulated_per_sec"
UNIT = "joints/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=16384, help="frames per GPU per step")
    ap.add_argument("--workload", default="cfg2_hall16x6")
    ap.add_argument("--ref-frames", type=int, default=0, help="frames per step of the CPU reference arm (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary measurements (other configs, latency)")
    ap.add_argument("--extra-seconds", type=float, default=12.0, help="time box per secondary config")
    return ap.parse_args()


# ------------------------------------------------------------------ flop / byte model
def algorithmic_work(fr, res, thr=0.30):
    """Per-frame algorithmic flops and bytes of a batch, from the survey's cost model applied to the
    batch's actual view counts: epipolar pair 50; weighted DLT 88n+1003; unweighted 80n+1003; reprojection
    error 32n; per joint T_w + R + (4n+1) T_u; reprojection 250 per (joint, camera). Bytes: Person2D 428 B in,
    PersonCov 1684 B out, reprojected Person2D 428 B out (wire layouts, each byte once)."""
    persons, n_persons, hyp_of = fr["persons"], fr["n_persons"], res["hyp_of"]
    F, C, PM = hyp_of.shape
    score = persons["keypoints"]["score"]                      # [F][C][PM][17]
    slot = np.arange(PM)[None, None, :] < n_persons[:, :, None]
    valid_kp = (score >= thr) & slot[..., None]
    assigned = hyp_of >= 0
    H = int(hyp_of.max()) + 1 if assigned.any() else 0
    # views per (frame, hypothesis, joint)
    n_views = np.zeros((F, max(H, 1), 17), np.int32)
    n_obs = np.zeros((F, max(H, 1)), np.int32)
    f_idx = np.broadcast_to(np.arange(F)[:, None, None], hyp_of.shape)
    np.add.at(n_obs, (f_idx[assigned], hyp_of[assigned]), 1)
    np.add.at(n_views, (f_idx[assigned], hyp_of[assigned]), valid_kp[assigned].astype(np.int32))
    n = n_views[(n_obs >= 2)[..., None] & (n_views >= 2)].astype(np.float64)
    tri_flops = ((88 * n + 1003) + 32 * n + (4 * n + 1) * (80 * n + 1003)).sum()
    # association: every (hypothesis observation, detection) pair of later cameras, joints valid in both;
    # approximated by the pairs of valid detections in different cameras (upper bound of what calcCost visits
    # is data dependent; this counts each unordered cross-camera detection pair once, as the cost matrix does)
    strict = (score > thr) & slot[..., None]
    det_valid = (valid_kp.sum(-1) > 8)
    v = (strict & det_valid[..., None]).astype(np.float64)     # [F][C][PM][17]
    per_cam = v.sum(2)                                          # [F][C][17] valid joint count per camera
    tot = per_cam.sum(1)
    pairs = ((tot ** 2 - (per_cam ** 2).sum(1)) / 2).sum()      # sum over joints of cross-camera pairs
    assoc_flops = 50.0 * pairs
    live = np.arange(res["persons3d"].shape[1])[None, :] < res["n_out"][:, None]
    joints_out = int(((res["persons3d"]["keypoints"]["score"] > 0) & live[..., None]).sum())
    rep_flops = 250.0 * joints_out * C
    bytes_in = float(n_persons.sum()) * 428
    bytes_out3d = float(res["n_out"].sum()) * 1684
    bytes_out2d = float(res.get("n_out2d_total", 0)) * 428
    return dict(tri_flops_per_frame=tri_flops / F, assoc_flops_per_frame=assoc_flops / F,
                reproj_flops_per_frame=rep_flops / F, joints_per_frame=joints_out / F,
                bytes_per_frame=(bytes_in + bytes_out3d + bytes_out2d) / F,
                mean_views_per_joint=float(n.mean()) if n.size else 0.0,
                mean_detections_per_camera=float(n_persons.mean()))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, sm_load, mx, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                if float(r[7]) >= 20.0:
                    sm_load.append(float(r[0]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        under = sm_load or sm
        return {"sm_mhz": float(np.median(under)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": len(sm_load)}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)),
                    sm_sustained_mhz=float(d.get("clocks_under_load", {}).get("sm_mhz_median", 1327.0)), src="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, sm_sustained_mhz=1327.0, src="fallback")


def cpu_reference_run(fr, n_frames, n_threads, steps=1, warmup=0):
    """Time the CPU oracle (triangulation + reprojection) on n_frames frames; returns (joints/s, frames/s, ms/step)."""
    from oracle.binding import REF_HUNGARIAN_PATH, Oracle
    orc = Oracle(fr["cameras"], ref_hungarian=REF_HUNGARIAN_PATH.exists())
    persons, n_persons = fr["persons"][:n_frames], fr["n_persons"][:n_frames]
    times, joints = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        r = orc.triangulate_batch(persons, n_persons, fr["h_max"], n_threads=n_threads)
        orc.reproject_batch(r["persons3d"], r["n_out"], n_threads=n_threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            joints = r["n_joints"]
    t = float(np.mean(times))
    # kind: the node itself cannot be built here (ROS / Eigen absent), so this is a PORT of the algorithm; only the
    # Munkres solver inside it is the reference's own code when oracle/_ref is present
    global CPU_LABEL
    CPU_LABEL = "port+reference-hungarian" if orc.ref_hungarian else "port"
    return joints / t, n_frames / t, 1e3 * t, "port"


def workload_config(name, B, fr, stages="associate+triangulate(+UT covariance)+finalize+reproject"):
    """The `config` object both arms print: it names the workload, nothing measured goes in here."""
    from smartedgesensor3dhumanpose_b200 import workloads
    rig, people, dropout, seed = workloads.CONFIGS[name]
    C, PM = fr["persons"].shape[1], fr["persons"].shape[2]
    return {"workload": name, "rig": rig, "cameras": int(C), "people": people, "dropout": dropout, "seed": seed,
            "p_max": int(PM), "h_max": int(fr["h_max"]), "frames_per_step_per_gpu": int(B), "stages": stages,
            "l2_policy": "inputs of one step are larger than L2 (no flush needed)",
            "sharding": "frames across ranks, no data-path collective; final gather timed separately"}


# key, workload, frames per step, parameter overrides
EXTRA_RUNS = [
    ("dense_ring16x6", "dense_ring16x6", 4096, {}),
    ("cfg3_fp32_lm", "cfg3_hall16x6_dropout", 16384, {"lm_refine": 1}),
    ("cfg3_fp64_lm", "cfg3_hall16x6_dropout", 16384, {"lm_refine": 1, "precision": 1}),
    ("cfg4_crowd64x20", "cfg4_crowd64x20", 512, {}),
    ("cfg5_ring8x4", "cfg5_ring8x4", 16384, {}),
]


def run_extra(workload, B, prm, seconds, peaks, device):
    """One secondary config, bounded to ~`seconds`: frames resident in HBM, full path, CUDA events on the launching
    stream; parity of a sample against the CPU oracle (association bit-exact; joints within tolerance outside the
    branch-threshold eps-band, tests/test_gpu_parity.py); roofline of the dominant kernel as in the headline."""
    import torch
    from oracle.binding import REF_HUNGARIAN_PATH, Oracle
    from smartedgesensor3dhumanpose_b200 import api, workloads
    from smartedgesensor3dhumanpose_b200.layouts import default_params, person2d_dtype, person_cov_dtype
    t_start = time.perf_counter()
    params = default_params(**prm)
    fp64 = bool(prm.get("precision"))
    fr = workloads.make_workload(workload, B)
    cams, h_max = fr["cameras"], fr["h_max"]
    C, PM = fr["persons"].shape[1], fr["persons"].shape[2]
    dev = torch.device(f"cuda:{device}")
    pipe = api.GeometryPipeline(cams, params, device=device)
    d_in = torch.from_numpy(fr["persons"].view(np.uint8).reshape(-1)).to(dev)
    d_n = torch.from_numpy(fr["n_persons"]).to(dev)
    d3 = torch.zeros(B * h_max * person_cov_dtype.itemsize, dtype=torch.uint8, device=dev)
    n3 = torch.zeros(B, dtype=torch.int32, device=dev)
    d2 = torch.zeros(B * C * h_max * person2d_dtype.itemsize, dtype=torch.uint8, device=dev)
    n2 = torch.zeros(B * C, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        pipe.process_device(B, PM, h_max, d_in.data_ptr(), d_n.data_ptr(), d3.data_ptr(), n3.data_ptr(), d2.data_ptr(),
                            n2.data_ptr(), stream=stream.cuda_stream)

    step()
    pipe.check()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); step(); e1.record(stream); torch.cuda.synchronize()
    one = max(e0.elapsed_time(e1), 1e-3)
    steps = int(max(3, min(20, 0.35 * seconds * 1e3 / one)))
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out3d = d3.cpu().numpy().view(person_cov_dtype).reshape(B, h_max)
    n3h = n3.cpu().numpy()
    live = np.arange(h_max)[None, :] < n3h[:, None]
    joints = int(((out3d["keypoints"]["score"] > 0) & live[..., None]).sum())
    pipe.set_profiling(True)
    step()
    kms = pipe.last_kernel_ms()
    pipe.set_profiling(False)
    # parity sample
    n_par = int(min(B, {"cfg4_crowd64x20": 6}.get(workload, 192)))
    sub = dict(persons=fr["persons"][:n_par], n_persons=fr["n_persons"][:n_par])
    orc = Oracle(cams, params, ref_hungarian=REF_HUNGARIAN_PATH.exists())
    ro = orc.triangulate_batch(sub["persons"], sub["n_persons"], h_max, n_threads=os.cpu_count() or 1, diag=True)
    rg = pipe.triangulate_batch(sub["persons"], sub["n_persons"], h_max)
    keep = ro["margin"] >= 1e-4
    tol = 1e-4 if fp64 else 1e-3
    assoc_ok = bool(np.array_equal(ro["hyp_of"], rg["hyp_of"]) and np.array_equal(ro["n_hungarian"], rg["n_hungarian"]))
    ka, kb = ro["persons3d"]["keypoints"][keep], rg["persons3d"]["keypoints"][keep]
    same_sets = bool(np.array_equal(ro["n_out"][keep], rg["n_out"][keep]) and np.array_equal(ka["score"] > 0, kb["score"] > 0))
    dmax = 0.0
    if same_sets:
        m = ka["score"] > 0
        dd = np.sqrt((ka["x"] - kb["x"]) ** 2 + (ka["y"] - kb["y"]) ** 2 + (ka["z"] - kb["z"]) ** 2)[m]
        dmax = float(dd.max(initial=0.0))
    # roofline of the dominant kernel (same flop model as the headline)
    res = dict(rg)
    rp = pipe.reproject_batch(rg["persons3d"], rg["n_out"])
    res["n_out2d_total"] = int(rp["n_out"].sum())
    work = algorithmic_work(dict(sub), res)
    dom = max(kms, key=kms.get)
    dom_flops = {"triangulate": work["tri_flops_per_frame"], "associate": work["assoc_flops_per_frame"],
                 "reproject": work["reproj_flops_per_frame"], "finalize": 0.0}[dom] * B
    pipe_peak = 148 * 128 * (1 if fp64 and dom == "triangulate" else 2) * peaks["sm_max_mhz"] * 1e6 / 1e12
    achieved = dom_flops / (kms[dom] * 1e-3) / 1e12
    out = {"workload": workload, "params": prm, "frames_per_step": B, "ms_per_step": ms, "steps": steps,
           "value": joints / (ms * 1e-3), "unit": UNIT, "frames_per_sec": B / (ms * 1e-3),
           "joints_per_frame": joints / B, "mean_views_per_joint": work["mean_views_per_joint"],
           "kernel_ms_per_step": kms,
           "parity": {"ok": bool(assoc_ok and same_sets and dmax <= tol), "association_bit_exact": assoc_ok,
                      "same_persons_and_joints": same_sets, "max_joint_dev_m": dmax, "tolerance_m": tol,
                      "frames_checked": int(keep.sum()), "frames_in_eps_band": int((~keep).sum()),
                      "n_hungarian_per_frame": float(ro["n_hungarian"].mean())},
           "roofline": {"kernel": f"k_{dom}", "bound": "fp64" if fp64 and dom == "triangulate" else "fp32",
                        "achieved": achieved, "peak": pipe_peak, "unit": "TFLOP/s", "frac": achieved / pipe_peak,
                        "algorithmic_mflop_per_frame": {"associate": work["assoc_flops_per_frame"] / 1e6,
                                                        "triangulate": work["tri_flops_per_frame"] / 1e6,
                                                        "reproject": work["reproj_flops_per_frame"] / 1e6}},
           "seconds": round(time.perf_counter() - t_start, 1)}
    pipe.close()
    return out


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from smartedgesensor3dhumanpose_b200 import workloads as helpers   # CONFIGS / make_workload live in the package
    rig, people, dropout, seed = helpers.CONFIGS[a.workload]

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        pilot = helpers.make_workload(a.workload, 256)
        _, fps, _, _ = cpu_reference_run(pilot, 256, cores)
        n = a.ref_frames or int(min(32768, max(256, fps * 8.0)))  # ~8 s of CPU work per step
        fr = helpers.make_workload(a.workload, n)
        jps, fps, ms, kind = cpu_reference_run(fr, n, cores, steps=a.steps, warmup=min(a.warmup, 1))
        line = {"metric": METRIC, "value": jps, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "frames_per_sec": fps,
                "config": workload_config(a.workload, a.frames, fr),
                "reference_frames_per_step": n,
                "cpu_baseline": {"value": jps, "unit": UNIT, "cores": cores, "kind": kind, "label": CPU_LABEL,
                                 "sample": f"{n} frames/step of {a.workload}, frame-parallel over {cores} threads"},
                "e2e": {"value": jps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    from smartedgesensor3dhumanpose_b200 import api
    from smartedgesensor3dhumanpose_b200.layouts import person2d_dtype, person_cov_dtype

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    # host placement: this rank's thread (and therefore the pinned buffers it allocates next) on the GPU's NUMA node
    from smartedgesensor3dhumanpose_b200 import lib as _libmod
    numa_node = _libmod.bind_thread_to_device_numa(local_rank)
    B = a.frames
    fr = helpers.make_workload(a.workload, B, first_frame=rank * B)
    cams, h_max = fr["cameras"], fr["h_max"]
    C, PM = fr["persons"].shape[1], fr["persons"].shape[2]
    pipe = api.GeometryPipeline(cams, device=local_rank)
    pipe.reserve(B, PM, h_max)

    def pinned(arr):
        t = torch.from_numpy(arr.view(np.uint8).reshape(-1)).pin_memory()
        return t, t.numpy().view(arr.dtype).reshape(arr.shape)

    # host buffers (pinned) and device-resident copies
    tp, persons_h = pinned(fr["persons"])
    tn, n_persons_h = pinned(fr["n_persons"])
    t3, out3d_h = pinned(np.zeros((B, h_max), person_cov_dtype))
    tn3, n3d_h = pinned(np.zeros(B, np.int32))
    t2, out2d_h = pinned(np.zeros((B, C, h_max), person2d_dtype))
    tn2, n2d_h = pinned(np.zeros((B, C), np.int32))
    d_persons, d_np = tp.to(dev), tn.to(dev)
    d_out3d = torch.zeros(t3.numel(), dtype=torch.uint8, device=dev)
    d_n3d = torch.zeros(B, dtype=torch.int32, device=dev)
    d_out2d = torch.zeros(t2.numel(), dtype=torch.uint8, device=dev)
    d_n2d = torch.zeros(B * C, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        pipe.process_device(B, PM, h_max, d_persons.data_ptr(), d_np.data_ptr(), d_out3d.data_ptr(), d_n3d.data_ptr(),
                            d_out2d.data_ptr(), d_n2d.data_ptr(), stream=stream.cuda_stream)

    bufs = dict(persons3d=out3d_h, n_out3d=n3d_h, persons2d=out2d_h, n_out2d=n2d_h)

    def step_host_padded():
        pipe.process_batch(persons_h, n_persons_h, h_max, bufs=bufs)

    # ragged host call (the public batch API for message-like, variable-length lists): dense pinned buffers
    pad0 = pipe.process_batch(persons_h, n_persons_h, h_max, bufs=bufs)
    tot3, tot2 = int(pad0["n_out3d"].sum()), int(pad0["n_out2d"].sum())
    tdi, dense_in_h = pinned(api.to_ragged(fr["persons"], fr["n_persons"]))
    td3, dense3_h = pinned(np.zeros(tot3 + 16, person_cov_dtype))
    td2, dense2_h = pinned(np.zeros(tot2 + 16, person2d_dtype))

    def step_host():
        pipe.process_batch_ragged(dense_in_h, n_persons_h, PM, h_max, dense3_h, n3d_h, dense2_h, n2d_h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record(stream)
        for i in range(steps):
            fn()
            evs[i + 1].record(stream)
        barrier()
        per = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(steps)])
        total = evs[0].elapsed_time(evs[-1])
        if world > 1:
            t = torch.tensor([total], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total, per

    # joints per step (count once from a real run; identical every step since the batch is fixed)
    step_device()
    torch.cuda.synchronize()
    n3d = d_n3d.cpu().numpy()
    out3d = d_out3d.cpu().numpy().view(person_cov_dtype).reshape(B, h_max)
    live = np.arange(h_max)[None, :] < n3d[:, None]
    joints_step = int(((out3d["keypoints"]["score"] > 0) & live[..., None]).sum())
    jt = torch.tensor([joints_step], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(jt)
    joints_all = int(jt.item())

    clk = ClockSampler(local_rank)
    clk.__enter__()
    time.sleep(0.6)   # let nvidia-smi start sampling before the timed region
    launches0 = pipe.launch_count
    total_ms, per_ms = timed(step_device, a.steps, a.warmup)
    launches = (pipe.launch_count - launches0) * a.steps // (a.steps + a.warmup)
    value = joints_all * a.steps / (total_ms * 1e-3)
    frames_ps = B * world * a.steps / (total_ms * 1e-3)

    e2e_ms, e2e_per = timed(step_host, a.steps, a.warmup)
    e2e_value = joints_all * a.steps / (e2e_ms * 1e-3)
    h2d = int(dense_in_h.nbytes + fr["n_persons"].nbytes)
    d2h = int(tot3 * person_cov_dtype.itemsize + tot2 * person2d_dtype.itemsize + n3d_h.nbytes + n2d_h.nbytes)
    e2e_pad_ms, _ = timed(step_host_padded, max(2, a.steps // 2), 1)
    e2e_pad_steps = max(2, a.steps // 2)
    clk.__exit__(None, None, None)

    # per-kernel device time (CUDA events on the launching stream, inside the library), same steps
    pipe.set_profiling(True)
    ksum = {}
    for _ in range(a.steps):
        step_device()
        for k, v in pipe.last_kernel_ms().items():
            ksum[k] = ksum.get(k, 0.0) + v
    pipe.set_profiling(False)
    kernel_ms = {k: v / a.steps for k, v in ksum.items()}

    # final result gather (compact xyz+score per joint), timed separately - the only collective
    gather_ms = None
    if world > 1:
        from smartedgesensor3dhumanpose_b200 import sharding
        compact = sharding.compact_torch(d_out3d, B, h_max)
        sharding.gather_compact(compact[:1].contiguous(), dst=0)   # connection set-up is not part of the gather
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        parts = sharding.gather_compact(compact, dst=0, shapes=[tuple(compact.shape)] * world)
        t1.record()
        barrier()
        gather_ms = t0.elapsed_time(t1)
        if rank == 0:
            assert len(parts) == world and all(p.shape == compact.shape for p in parts)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel
    peaks = load_peaks()
    sample = min(B, 2048)
    sub = {k: (v[:sample] if isinstance(v, np.ndarray) and v.shape[:1] == (B,) else v) for k, v in fr.items()}
    res = pipe.triangulate_batch(sub["persons"], sub["n_persons"], h_max)
    rp = pipe.reproject_batch(res["persons3d"], res["n_out"])
    res["n_out2d_total"] = int(rp["n_out"].sum())
    work = algorithmic_work(sub, res)
    fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12          # TFLOP/s at max clock
    fp32_sustained = 148 * 128 * 2 * peaks["sm_sustained_mhz"] * 1e6 / 1e12
    dom = max(kernel_ms, key=kernel_ms.get)
    dom_flops = {"triangulate": work["tri_flops_per_frame"], "associate": work["assoc_flops_per_frame"],
                 "reproject": work["reproj_flops_per_frame"], "finalize": 0.0}[dom] * B
    n_launch = max(1, -(-B // 16384))
    achieved = dom_flops / (kernel_ms[dom] * 1e-3) / 1e12
    step_ms = total_ms / a.steps
    roofline = {"bound": "fp32", "kernel": f"k_{dom}", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved / fp32_peak, "peak_source": f"148 SM x 128 lanes x 2 x sm_max_mhz ({peaks['src']} MEASURED_PEAKS.json)",
                "frac_of_sustained_clock_peak": achieved / fp32_sustained,
                "launches_per_step": n_launch, "kernel_ms_per_step": kernel_ms, "kernel_share_of_step": kernel_ms[dom] / sum(kernel_ms.values()),
                "algorithmic_flops_per_frame": {"associate": work["assoc_flops_per_frame"], "triangulate": work["tri_flops_per_frame"],
                                                "reproject": work["reproj_flops_per_frame"]},
                "whole_step_tflops": (work["tri_flops_per_frame"] + work["assoc_flops_per_frame"] + work["reproj_flops_per_frame"]) * B / (step_ms * 1e-3) / 1e12,
                "hbm": {"achieved": work["bytes_per_frame"] * B / (step_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": work["bytes_per_frame"] * B / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        "algorithmic_bytes_per_frame": work["bytes_per_frame"]},
                "traffic": None}
    # the FMA pipe measured live on this GPU (micro-benchmark inside the library): the formula above assumes the
    # maximum clock, the micro-benchmark shows what the pipe sustains on this box
    try:
        from smartedgesensor3dhumanpose_b200 import lib as _l
        m = _l.measure_fma_peak(local_rank, fp64=False)
        roofline["peak_measured_fma_microbench"] = m
        roofline["frac_of_measured_fma_peak"] = achieved / m if m > 0 else None
    except Exception as e:   # diagnostics only
        roofline["peak_measured_fma_microbench"] = None
        roofline["peak_measured_error"] = str(e)
    # DRAM traffic of the dominant kernel per launch from the committed ncu --set full capture of the same launch
    # shape (scripts/profile_step.py, profiles/<round>_ncu_full_summary.json); null if no matching capture
    try:
        cap = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        ent = cap.get(a.workload, {}).get(f"k_{dom}")
        if ent and int(ent["frames_per_launch"]) == min(B, 16384):
            roofline["traffic"] = {"dram_bytes_per_launch": ent["dram_bytes_read"] + ent["dram_bytes_write"],
                                   "dram_bytes_read": ent["dram_bytes_read"], "dram_bytes_write": ent["dram_bytes_write"],
                                   "source": ent["source"]}
    except (OSError, ValueError, KeyError):
        pass

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "frames_per_sec": frames_ps,
            "p50_frame_latency_us": float(np.median(per_ms)) * 1e3 / B,
            "config": workload_config(a.workload, B, fr),
            "workload_stats": {"joints_per_frame": work["joints_per_frame"], "mean_views_per_joint": work["mean_views_per_joint"],
                               "mean_detections_per_camera": work["mean_detections_per_camera"],
                               "h2d_mib_per_step": h2d / 2**20, "d2h_mib_per_step": d2h / 2**20,
                               "host_numa_node_rank0": numa_node,
                               "note": "hall rig with f = 1000 px: a camera sees 2.4 of the 6 people and a joint 4.4 views; "
                                       "the 16-view case of the survey's flop model is extra.dense_ring16x6"},
            "roofline": roofline, "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / a.steps, "frames_per_sec": B * world * a.steps / (e2e_ms * 1e-3),
                    "call": "ses3d_process_batch_ragged, pinned host buffers, occupied records only",
                    "padded_call": {"call": "ses3d_process_batch ([F][C][p_max] in, [F][h_max] + [F][C][h_max] out)",
                                    "value": joints_all * e2e_pad_steps / (e2e_pad_ms * 1e-3),
                                    "frames_per_sec": B * world * e2e_pad_steps / (e2e_pad_ms * 1e-3),
                                    "h2d_bytes_per_step": int(fr["persons"].nbytes + fr["n_persons"].nbytes),
                                    "d2h_bytes_per_step": int(out3d_h.nbytes + n3d_h.nbytes + out2d_h.nbytes + n2d_h.nbytes)}},
            "gpu_launches": int(launches), "gather_ms": gather_ms}

    if world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        _, fps, _, _ = cpu_reference_run(fr, 256, cores)
        n = int(min(B, max(256, fps * 12.0)))
        jps, fps, ms, kind = cpu_reference_run(fr, n, cores)
        _, fps1, _, _ = cpu_reference_run(fr, min(n, max(64, int(fps / cores * 3))), 1)
        line["cpu_baseline"] = {"value": jps, "unit": UNIT, "cores": cores, "kind": kind, "label": CPU_LABEL,
                                "frames_per_sec": fps,
                                "single_thread_frames_per_sec": fps1,
                                "single_thread_mean_frame_latency_ms": 1e3 / fps1 if fps1 > 0 else None,
                                "sample": f"first {n} frames of the step's batch, frame-parallel over {cores} threads ({ms:.0f} ms)"}
    if world == 1 and not a.no_extra:
        # the other BASELINE.json configs, time-boxed: device-resident throughput, parity against the oracle on a
        # sample of the same frames, roofline fraction of the dominant kernel
        del d_persons, d_np, d_out3d, d_n3d, d_out2d, d_n2d
        torch.cuda.empty_cache()
        line["extra"] = {}
        for key, wl, frames, prm in EXTRA_RUNS:
            try:
                line["extra"][key] = run_extra(wl, frames, prm, a.extra_seconds, peaks, local_rank)
            except Exception as e:   # a failing secondary run must not hide the headline
                line["extra"][key] = {"error": f"{type(e).__name__}: {e}"}
        # single-frame call latency (the ROS-shim use case), host buffers, wall clock
        lat = []
        for f in range(200):
            t0 = time.perf_counter()
            pipe.process_batch(persons_h[f:f + 1], n_persons_h[f:f + 1], h_max)
            lat.append(time.perf_counter() - t0)
        line["single_frame_call_p50_us"] = float(np.median(lat[20:])) * 1e6
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
