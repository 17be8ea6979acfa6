"""In-tree build of the C-ABI library libses3d.so for sm_100a (nvcc, no cmake).

    python -m smartedgesensor3dhumanpose_b200.build [--force] [--verbose]

The association / finalize / reproject kernels are compiled with -fmad=false (bit-exact with the
reference's x86-64 arithmetic); host code with -ffp-contract=off. The .so stays in the package
directory so that it travels to the GPU box with the repo snapshot.
"""
import argparse
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
BUILD = PKG / "build"
LIB = PKG / "libses3d.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", f"-I{INCLUDE}", f"-I{CSRC}", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
          "--expt-relaxed-constexpr"]
UNITS = [
    ("kernels_exact.cu", ["-fmad=false"]),
    ("kernels_tri.cu", []),
    ("synth_device.cu", ["-fmad=false"]),
    ("kernels_pack.cu", []),
    ("kernels_prior.cu", []),
    ("kernels_peak.cu", []),
    ("kernels_markers.cu", []),
    ("kernels_overlay.cu", []),
    ("prior_api.cpp", []),
    ("api.cpp", []),
    ("multi.cpp", []),
    ("host_setup.cpp", []),
    ("synth.cpp", []),
    ("frame_assembler.cpp", []),
    ("wire.cpp", []),
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return exe


def _stale(out: Path, deps):
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps if d.exists())


SYNTH_LIB = PKG / "libses3d_synth.so"


def build_synth(force=False, verbose=False):
    """The host-side synthetic frame generator as its own small library (g++, no CUDA): test / bench input source.
    Kept apart from libses3d.so so that a process which only needs inputs (e.g. bench.py --impl reference, which
    times the CPU oracle) never maps the product library."""
    deps = [CSRC / "synth.cpp", CSRC / "synth.h", INCLUDE / "ses3d.h", Path(__file__)]
    if not force and not _stale(SYNTH_LIB, deps):
        return SYNTH_LIB
    cxx = shutil.which("g++") or shutil.which("c++")
    if not cxx:
        raise RuntimeError("g++ not found: the synthetic generator library cannot be built")
    cmd = [cxx, "-O3", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", f"-I{INCLUDE}",
           f"-I{CSRC}", "-o", str(SYNTH_LIB), str(CSRC / "synth.cpp"), "-lpthread"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return SYNTH_LIB


def build(force=False, verbose=False, ptxas_info=False):
    build_synth(force, verbose)
    headers = list(CSRC.glob("*.h")) + list(INCLUDE.glob("*.h")) + [Path(__file__)]
    sources = [CSRC / name for name, _ in UNITS if (CSRC / name).exists()]
    if not force and not ptxas_info and not _stale(LIB, sources + headers):
        return LIB      # up to date (e.g. the prebuilt library shipped to the GPU box without the object files)
    nvcc = _nvcc()
    BUILD.mkdir(exist_ok=True)
    objs = []
    for name, extra in UNITS:
        src = CSRC / name
        if not src.exists():
            continue
        obj = BUILD / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, "-c", str(src), "-o", str(obj)] + ARCH + COMMON + extra
            if ptxas_info and src.suffix == ".cu":
                cmd += ["-Xptxas", "-v"]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.run(cmd, check=True)
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB)] + [str(o) for o in objs] + ARCH + ["-Xcompiler", "-fPIC", "-lpthread"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--ptxas-info", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose, a.ptxas_info))
