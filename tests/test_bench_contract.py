"""CPU test of the bench.py contract for the reference arm (the GPU arm needs a device): one JSON line with the
required keys; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-frames", "64", "--workload", "cfg5_ring8x4"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "joints_triangulated_per_sec" and d["value"] > 0
    assert d["config"]["workload"] == "cfg5_ring8x4" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    # the node itself cannot be built here: the arm is a port whose Munkres solver is the reference's verbatim file
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["label"] in ("port+reference-hungarian", "port")
    # the reference arm must not map the product library: its inputs come from libses3d_synth.so
    assert "frames_per_step_per_gpu" in d["config"] and "p_max" in d["config"] and "h_max" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # ranks other than 0 exit quietly
    r1 = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference"], capture_output=True, text=True,
                        timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
