#!/usr/bin/env python
"""Turn the ncu captures / bench lines under gpurun_out/ into the committed summaries under profiles/.
usage: python scripts/summarise_profiles.py r01"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
P.mkdir(exist_ok=True)

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "lts__t_bytes.sum", "l1tex__t_bytes.sum"]

# launch list
src = G / f"launches_{tag}.csv"
if src.exists():
    rows = list(csv.reader(open(src)))
    hdr, keep = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            keep.append((d["ID"], d["Kernel Name"].split("(")[0], d["Grid Size"], d["Block Size"], d["Metric Value"]))
    with open(P / f"{tag}_launches.csv", "w") as f:
        f.write("id,kernel,grid,block,gpu__time_duration.sum_ns\n")
        for k in keep:
            f.write(",".join(f'"{x}"' for x in k) + "\n")
    per = {}
    for k in keep[len(keep) // 2:]:
        per.setdefault(k[1], []).append(float(k[4]))
    tot = sum(sum(v) for v in per.values())
    with open(P / f"{tag}_launch_shares.txt", "w") as f:
        f.write(f"kernel shares of the step (second half of the {len(keep)} captured launches; ncu --metrics "
                "gpu__time_duration.sum --clock-control none; cold-cache serialised times: compare shares)\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{100 * sum(v) / tot:6.2f}%  n={len(v):3d}  mean {sum(v) / len(v) / 1e3:9.1f} us  {k}\n")

summ = {}
for rep in sorted(G.glob(f"prof_k_*_{tag}.ncu-rep")):
    name = rep.stem.replace("prof_", "").replace(f"_{tag}", "")
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    summ[name] = {m: f"{vals[hdr.index(m)]} {units[hdr.index(m)]}".strip() for m in METRICS if m in hdr}
    src_csv = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--print-source", "cuda,sass", "--csv"],
                             capture_output=True, text=True).stdout
    tmp = G / f"{name}_src.csv"
    tmp.write_text(src_csv)
    lines = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_lines.py"), str(tmp), "25"], capture_output=True, text=True).stdout
    (P / f"{tag}_{name}_hot_lines.txt").write_text(lines)
if summ:
    (P / f"{tag}_ncu_full_summary.json").write_text(json.dumps(summ, indent=1))
    # per-launch DRAM traffic for bench.py's roofline.traffic (captures made with scripts/profile_step.py defaults)
    def to_bytes(txt):
        v, u = txt.split()[0], txt.split()[1] if len(txt.split()) > 1 else "byte"
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    workload = sys.argv[3] if len(sys.argv) > 3 else "cfg2_hall16x6"
    traffic = {workload: {k: {"dram_bytes_read": to_bytes(v["dram__bytes_read.sum"]), "dram_bytes_write": to_bytes(v["dram__bytes_write.sum"]),
                              "frames_per_launch": frames, "source": f"profiles/{tag}_ncu_full_summary.json"}
                          for k, v in summ.items() if "dram__bytes_read.sum" in v}}
    if "k_prior" in traffic[workload]:   # captured with scripts/bench_prior.py --profile-only (2048 streams x 32 messages)
        kp = traffic[workload].pop("k_prior")
        kp["frames_per_launch"] = 2048 * 32
        traffic["pose_prior"] = {"k_prior": kp}
    (P / "ncu_traffic.json").write_text(json.dumps(traffic, indent=1))
for name in (f"bench_{tag}.json", f"bench_{tag}_reference.json", f"bench_{tag}_dense.json", f"configs_{tag}.json"):
    if (G / name).exists():
        (P / name).write_text((G / name).read_text())
print("profiles:", sorted(p.name for p in P.iterdir()))
