#!/usr/bin/env python
"""A/B builds of one kernel translation unit: the same library with different -D switches, written next to this script
as _variants/libses3d_<tag>.so (select one with SES3D_LIB=...). Development tool, not part of the product build.
    python scripts/build_variants.py J1P1U0:-DSES_COLD_JACOBI=1,-DSES_COLD_PATHS=1 J0P0U0:-DSES_COLD_JACOBI=0,-DSES_COLD_PATHS=0
    python scripts/build_variants.py --unit kernels_exact.cu M1:-DSES_COLD_MUNKRES=1
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from smartedgesensor3dhumanpose_b200 import build as b  # noqa: E402

b.build()
out = Path(__file__).resolve().parent / "_variants"
out.mkdir(exist_ok=True)
nvcc = b._nvcc()
procs = []
args = sys.argv[1:]
unit = "kernels_tri.cu"
if args and args[0] == "--unit":
    unit, args = args[1], args[2:]
extra = dict(b.UNITS)[unit]
stem = unit.rsplit(".", 1)[0]
for spec in args:
    tag, flags = spec.split(":")
    obj = out / f"{stem}_{tag}.o"
    cmd = [nvcc, "-c", str(b.CSRC / unit), "-o", str(obj)] + b.ARCH + b.COMMON + extra + flags.split(",")
    procs.append((tag, obj, subprocess.Popen(cmd)))
for tag, obj, p in procs:
    if p.wait() != 0:
        raise SystemExit(f"variant {tag} failed")
    objs = [str(obj if o.name == f"{stem}.o" else o) for o in sorted(b.BUILD.glob("*.o"))]
    lib = out / f"libses3d_{tag}.so"
    subprocess.run([nvcc, "-shared", "-o", str(lib)] + objs + b.ARCH + ["-Xcompiler", "-fPIC", "-lpthread"], check=True)
    print(lib)
