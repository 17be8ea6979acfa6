"""Run selected secondary configs of bench.py (run_extra) and print their key numbers: python scripts/extra_probe.py cfg3_fp64_lm ..."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

peaks = bench.load_peaks()
for key, wl, frames, prm in bench.EXTRA_RUNS:
    if sys.argv[1:] and key not in sys.argv[1:]:
        continue
    r = bench.run_extra(wl, frames, prm, 8.0, peaks, 0)
    print(key, json.dumps({"ms": round(r["ms_per_step"], 3), "kms": {k: round(v, 3) for k, v in r["kernel_ms_per_step"].items()},
                           "parity": r["parity"]["ok"], "dev": r["parity"]["max_joint_dev_m"], "frac": round(r["roofline"]["frac"], 3)}))
