"""Single-frame calls on the eager path, for `ncu -k regex:k_rounds -s N -c M python scripts/latency_one.py`."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["SES3D_FRAME_GRAPH"] = "0"
from smartedgesensor3dhumanpose_b200 import api, workloads  # noqa: E402

fr = workloads.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2_hall16x6", 40)
pipe = api.GeometryPipeline(fr["cameras"], device=0)
for f in range(40):
    pipe.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], fr["h_max"])
