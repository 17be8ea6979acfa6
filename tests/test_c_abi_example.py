"""The public header is valid C99 and a plain-C client links and runs against libses3d.so (examples/abi_smoke.c).
On a machine without a GPU the program checks the host-only entry points and that ses3d_create refuses to run."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "smartedgesensor3dhumanpose_b200"


def _build(tmp_path):
    exe = tmp_path / "abi_smoke"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{ROOT / 'include'}",
                    str(ROOT / "examples" / "abi_smoke.c"), f"-L{PKG}", "-lses3d", f"-Wl,-rpath,{PKG}", "-lm", "-o", str(exe)],
                   check=True)
    return exe


def test_header_is_c99_and_c_client_runs(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ses3d" in r.stdout


@pytest.mark.gpu
def test_c_client_single_frame_call_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "persons3d = 1" in r.stdout
    assert "pose_prior: track 0, 12 observations" in r.stdout
    assert "single-frame replays: identical records" in r.stdout
    assert "identical record" in r.stdout and "overlay:" in r.stdout and "overlay: 0 coloured" not in r.stdout
