#!/usr/bin/env python
"""Turn the round-2 record run under gpurun_out/ (scripts/gpu_r02_final.sh) into the committed summaries under profiles/."""
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
summ = {}
for k in ["k_triangulate", "k_finproj", "k_pairs", "k_rounds"]:
    rows = list(csv.reader(open(G / f"{tag}_{k}_raw.csv")))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    summ[k] = {m: f"{vals[hdr.index(m)]} {units[hdr.index(m)]}".strip() for m in METRICS if m in hdr}
    summ[k]["kernel"] = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else k
    lines = subprocess.run([sys.executable, str(ROOT / "scripts/ncu_lines.py"), str(G / f"{tag}_{k}_src.csv"), "25"], capture_output=True, text=True).stdout
    stalls = subprocess.run([sys.executable, str(ROOT / "scripts/ncu_stalls.py"), str(G / f"{tag}_{k}_src.csv"), "stall_no_inst", "8"], capture_output=True, text=True).stdout
    (P / f"{tag}_{k}_hot_lines.txt").write_text(lines + "\n" + stalls)
(P / f"{tag}_ncu_full_summary.json").write_text(json.dumps(summ, indent=1))


def to_bytes(txt):
    v, u = txt.split()[0], txt.split()[1] if len(txt.split()) > 1 else "byte"
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


tr = json.loads((P / "ncu_traffic.json").read_text())
w = tr.setdefault("cfg2_hall16x6", {})
for k, v in summ.items():
    w[k] = {"dram_bytes_read": to_bytes(v["dram__bytes_read.sum"]), "dram_bytes_write": to_bytes(v["dram__bytes_write.sum"]),
            "frames_per_launch": 16384, "source": f"profiles/{tag}_ncu_full_summary.json"}
w["k_reproject"] = dict(w["k_finproj"], note="fused finalize + reproject kernel k_finproj of the process calls")
w["k_associate"] = {"dram_bytes_read": w["k_pairs"]["dram_bytes_read"] + w["k_rounds"]["dram_bytes_read"],
                    "dram_bytes_write": w["k_pairs"]["dram_bytes_write"] + w["k_rounds"]["dram_bytes_write"],
                    "frames_per_launch": 16384, "source": f"profiles/{tag}_ncu_full_summary.json", "note": "k_pairs + k_rounds"}
(P / "ncu_traffic.json").write_text(json.dumps(tr, indent=1))
for k, v in summ.items():
    print(k, v["gpu__time_duration.sum"], "issue", v["smsp__issue_active.avg.pct_of_peak_sustained_active"], "lanes",
          v["smsp__thread_inst_executed_per_inst_executed.ratio"], "fma", v["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"],
          "fp64", v["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"], "warps", v["sm__warps_active.avg.pct_of_peak_sustained_active"],
          "dram r/w", v["dram__bytes_read.sum"], v["dram__bytes_write.sum"])
for f in (f"bench_{tag}.json", f"bench_{tag}_reference.json", f"bench_prior_{tag}.json", f"bench_chain_{tag}.json", f"latency_{tag}.txt"):
    if (G / f).exists():
        shutil.copy(G / f, P / f)
for src, dst in ((f"{tag}_soak.jsonl", f"{tag}_soak.jsonl"), (f"{tag}_tl.json", f"{tag}_e2e_timeline.json"),
                 (f"{tag}_latency_kernels.txt", f"{tag}_latency_kernels.txt"), (f"{tag}_link_n1.json", f"{tag}_link_n1.json"),
                 (f"{tag}_link_pieces.txt", f"{tag}_link_pieces.txt")):
    if (G / src).exists():
        shutil.copy(G / src, P / dst)
rows = list(csv.reader(open(G / f"launches_{tag}.csv")))
hdr, keep = None, []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        keep.append((d["ID"], d["Kernel Name"].split("(")[0], d["Grid Size"], d["Block Size"], d["Metric Value"]))
(P / f"{tag}_launches.csv").write_text("id,kernel,grid,block,gpu__time_duration.sum_ns\n" + "\n".join(",".join(f'"{x}"' for x in k) for k in keep) + "\n")
per = {}
for k in keep[len(keep) // 2:]:
    if "ses3d" in k[1]:
        per.setdefault(k[1], []).append(float(k[4]))
tot = sum(sum(v) for v in per.values())
out = [f"kernel shares of the step (second half of the {len(keep)} captured launches of scripts/profile_step.py; ncu --metrics "
       "gpu__time_duration.sum --clock-control none; cold-cache serialised times: compare shares)"]
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"{100 * sum(v) / tot:6.2f}%  n={len(v):3d}  mean {sum(v) / len(v) / 1e3:9.1f} us  {k}")
(P / f"{tag}_launch_shares.txt").write_text("\n".join(out) + "\n")
print("\n".join(out))
