#!/usr/bin/env python
"""A/B builds of kernels_tri.cu: the same library with different -D switches, written next to this script as
_variants/libses3d_<tag>.so (select one with SES3D_LIB=...). Development tool, not part of the product build.
    python scripts/build_variants.py J1P1U0:-DSES_COLD_JACOBI=1,-DSES_COLD_PATHS=1 J0P0U0:-DSES_COLD_JACOBI=0,-DSES_COLD_PATHS=0
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
for spec in sys.argv[1:]:
    tag, flags = spec.split(":")
    obj = out / f"kernels_tri_{tag}.o"
    cmd = [nvcc, "-c", str(b.CSRC / "kernels_tri.cu"), "-o", str(obj)] + b.ARCH + b.COMMON + flags.split(",")
    procs.append((tag, obj, subprocess.Popen(cmd)))
for tag, obj, p in procs:
    if p.wait() != 0:
        raise SystemExit(f"variant {tag} failed")
    objs = [str(obj if o.name == "kernels_tri.o" else o) for o in sorted(b.BUILD.glob("*.o"))]
    lib = out / f"libses3d_{tag}.so"
    subprocess.run([nvcc, "-shared", "-o", str(lib)] + objs + b.ARCH + ["-Xcompiler", "-fPIC", "-lpthread"], check=True)
    print(lib)
