"""Compile the shim nodes under ros_shim/src against the stub ROS headers (tests/ros_stub/include) and libses3d.so.
TEST INFRASTRUCTURE ONLY: proves the node sources compile unchanged and lets the replay tests run them."""
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
PKG = ROOT / "smartedgesensor3dhumanpose_b200"
SHIM = ROOT / "ros_shim"
OUT = HERE / "bin"
NODES = ["skeleton_3d_ses3d_node", "pose_reproj_ses3d_node", "pose_prior_ses3d_node"]


def _stale(out, deps):
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def build(force=False, syntax_only=False):
    cxx = shutil.which("g++") or shutil.which("c++")
    if not cxx:
        raise RuntimeError("g++ not found")
    OUT.mkdir(exist_ok=True)
    deps = list((HERE / "include").rglob("*.h")) + list((SHIM / "include").rglob("*.h")) + [HERE / "harness.cpp", ROOT / "include" / "ses3d.h", Path(__file__)]
    outs = {}
    for node in NODES:
        src = SHIM / "src" / f"{node}.cpp"
        exe = OUT / node
        outs[node] = exe
        common = [cxx, "-std=c++14", "-O1", "-Wall", "-Wextra", "-Wno-unused-parameter", f"-I{HERE / 'include'}",
                  f"-I{SHIM / 'include'}", f"-I{ROOT / 'include'}"]
        if syntax_only:
            subprocess.run(common + ["-fsyntax-only", str(src)], check=True)
            continue
        if not force and not _stale(exe, deps + [src, PKG / "libses3d.so"]):
            continue
        obj = OUT / f"{node}.o"
        subprocess.run(common + ["-Dmain=node_main", "-c", str(src), "-o", str(obj)], check=True)
        subprocess.run(common + [str(HERE / "harness.cpp"), str(obj), "-o", str(exe), f"-L{PKG}", "-lses3d",
                                 f"-Wl,-rpath,{PKG}", "-lpthread"], check=True)
    return outs


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, syntax_only="--syntax-only" in sys.argv))
