"""Build tests/hostsim/libhostsim.so (g++, no CUDA): the device algorithms run serially on the CPU. Test-only."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
CSRC = ROOT / "smartedgesensor3dhumanpose_b200" / "csrc"
LIB = HERE / "libhostsim.so"


def build(force=False):
    deps = [HERE / "hostsim.cpp", CSRC / "host_setup.cpp"] + list(CSRC.glob("*.h")) + [ROOT / "include" / "ses3d.h"]
    if not force and LIB.exists() and all(d.stat().st_mtime <= LIB.stat().st_mtime for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           f"-I{ROOT / 'include'}", f"-I{CSRC}", "-o", str(LIB), str(HERE / "hostsim.cpp"), str(CSRC / "host_setup.cpp")]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(True))
