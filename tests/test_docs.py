"""Keeps the documentation honest where that can be checked mechanically (no GPU needed)."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "smartedgesensor3dhumanpose_b200" / "csrc"


def test_every_tuning_switch_read_by_the_library_is_documented():
    """Each environment variable the C++ / CUDA sources read (getenv / env_int) has a row in INTEGRATION.md section 4."""
    read = set()
    for f in list(CSRC.glob("*.cpp")) + list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")):
        if f.name == "synth.cpp":   # the stand-alone workload generator is not the product library
            continue
        read |= set(re.findall(r'(?:getenv|env_int)\(\s*"(SES3D_[A-Z0-9_]+)"', f.read_text()))
    assert len(read) >= 20, sorted(read)
    doc = (ROOT / "INTEGRATION.md").read_text()
    missing = sorted(v for v in read if v not in doc)
    assert not missing, f"undocumented tuning switches: {missing}"


def test_header_entry_points_are_named_in_the_integration_guide():
    """Every function include/ses3d.h declares is mentioned in INTEGRATION.md or DESIGN.md (the reference-side binding a
    maintainer adds must be able to find it)."""
    hdr = (ROOT / "include" / "ses3d.h").read_text()
    names = set(re.findall(r"\b(ses3d_[a-z0-9_]+)\s*\(", hdr))
    text = (ROOT / "INTEGRATION.md").read_text() + (ROOT / "DESIGN.md").read_text() + (ROOT / "README.md").read_text()

    def expand(m):   # ses3d_wire_{decode,encode}_{a,b} and ses3d_x[_ragged] shorthands of the tables
        import itertools
        parts = re.split(r"\{([^{}]*)\}", m.group(0))
        alts = [p.split(",") if i % 2 else [p] for i, p in enumerate(parts)]
        return " ".join("".join(c) for c in itertools.product(*alts))
    text = re.sub(r"ses3d_[a-z0-9_]*(?:\{[^{}]*\}[a-z0-9_]*)+", expand, text)
    text = re.sub(r"(ses3d_[a-z0-9_]+)\[(_[a-z0-9_]+)\]", r"\1 \1\2", text)
    missing = sorted(n for n in names if n not in text and not any(n.startswith(p) and p + "*" in text for p in
                                                                    ("ses3d_prior_", "ses3d_assembler_", "ses3d_multi_")))
    assert not missing, f"entry points never mentioned in the docs: {missing}"
