import sys, json
sys.path.insert(0, "/root/repo")
import scripts.run_configs as rc
r = rc.run(*rc.RUNS[5])
print(json.dumps({k: r[k] for k in ("label", "ms_per_step", "frames_per_sec", "kernel_ms", "parity", "association_bit_exact")}))
