"""Single-frame host calls replay a captured CUDA graph (api.cpp::run_frame_graph): the per-message call of the ROS
nodes. The graph path must give exactly the bytes the batched / eager path gives, survive buffer growth in between
(captured device addresses go stale), shape changes, and report capacity overflows on every call."""
import os

import numpy as np
import pytest

from smartedgesensor3dhumanpose_b200 import api
from tests import helpers

pytestmark = pytest.mark.gpu


def _same(a, b):
    return a.tobytes() == b.tobytes()


def _single_calls(pipe, fr, frames, h_max):
    out = []
    for f in frames:
        r = pipe.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], h_max)
        out.append({k: v.copy() for k, v in r.items()})
    return out


def test_graph_replays_equal_the_batched_call():
    fr = helpers.make_workload("cfg2_hall16x6", 600)
    h_max = fr["h_max"]
    pipe = api.GeometryPipeline(fr["cameras"])
    batch = pipe.process_batch(fr["persons"], fr["n_persons"], h_max)
    n0 = pipe.launch_count
    singles = _single_calls(pipe, fr, range(40), h_max)     # call 0 eager + capture, 1.. replays
    per_call = (pipe.launch_count - n0) / 40
    assert per_call == pytest.approx(round(per_call)) and per_call >= 4   # replays are counted as their kernels
    for f, r in enumerate(singles):
        for k in ("persons3d", "n_out3d", "persons2d", "n_out2d"):
            assert _same(r[k][0], batch[k][f]), (f, k)
    # a big call in between grows / moves the slot buffers: the stale graph must not be replayed
    big = helpers.make_workload("cfg2_hall16x6", 9000, seed=3)
    pipe.process_batch(big["persons"], big["n_persons"], h_max)
    for f, r in zip(range(40, 80), _single_calls(pipe, fr, range(40, 80), h_max)):
        for k in ("persons3d", "n_out3d", "persons2d", "n_out2d"):
            assert _same(r[k][0], batch[k][f]), (f, k)
    # another h_max: re-captured for the new shape
    batch2 = pipe.process_batch(fr["persons"][:32], fr["n_persons"][:32], h_max + 5)
    for f, r in enumerate(_single_calls(pipe, fr, range(32), h_max + 5)):
        for k in ("persons3d", "n_out3d", "persons2d", "n_out2d"):
            assert _same(r[k][0], batch2[k][f]), (f, k)


def test_graph_path_of_the_separate_stages_and_eager_switch():
    fr = helpers.make_workload("cfg5_ring8x4", 64)
    h_max = fr["h_max"]
    pipe = api.GeometryPipeline(fr["cameras"])
    tri = pipe.triangulate_batch(fr["persons"], fr["n_persons"], h_max, dump=False)
    rep = pipe.reproject_batch(tri["persons3d"], tri["n_out"])
    os.environ["SES3D_FRAME_GRAPH"] = "0"
    try:
        eager = api.GeometryPipeline(fr["cameras"])
    finally:
        del os.environ["SES3D_FRAME_GRAPH"]
    for f in list(range(24)) + [3, 3, 60]:
        for p in (pipe, eager):
            t = p.triangulate_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], h_max, dump=False)
            assert _same(t["persons3d"][0], tri["persons3d"][f]) and t["n_out"][0] == tri["n_out"][f]
            r = p.reproject_batch(t["persons3d"], t["n_out"])
            assert _same(r["persons2d"][0], rep["persons2d"][f]) and _same(r["n_out"][0], rep["n_out"][f])
    # 3-D output not wanted (NULL out3d): the replay skips that copy, the 2-D records are the same
    both = pipe.process_batch(fr["persons"], fr["n_persons"], h_max)
    for f in (7, 8, 9, 7):
        r = pipe.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], h_max, want_3d=False)
        assert r["persons3d"] is None and r["n_out3d"][0] == both["n_out3d"][f]
        assert _same(r["persons2d"][0], both["persons2d"][f]) and _same(r["n_out2d"][0], both["n_out2d"][f])
    # a call that asks for the association dump stays on the eager path and still agrees
    d = pipe.triangulate_batch(fr["persons"][5:6], fr["n_persons"][5:6], h_max, dump=True)
    assert _same(d["persons3d"][0], tri["persons3d"][5]) and d["n_hyp"][0] >= d["n_out"][0]


def test_graph_path_reports_every_overflow_and_recovers():
    from smartedgesensor3dhumanpose_b200.lib import Ses3dError
    fr = helpers.make_workload("cfg5_ring8x4", 8)
    pipe = api.GeometryPipeline(fr["cameras"])
    ok = pipe.process_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    assert ok["n_out3d"].max() > 2
    f = int(np.argmax(ok["n_out3d"]))
    empty = np.zeros_like(fr["persons"][:1])
    none = np.zeros_like(fr["n_persons"][:1])
    for _ in range(3):   # an empty frame: eager, then captured, then replayed at h_max = 2
        r = pipe.process_batch(empty, none, 2)
        assert r["n_out3d"][0] == 0
    for _ in range(3):   # the replay overflows: reported each time, never sticky
        with pytest.raises(Ses3dError) as ei:
            pipe.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], 2)
        assert ei.value.code == -3
        r = pipe.process_batch(empty, none, 2)
        assert r["n_out3d"][0] == 0
    r = pipe.process_batch(fr["persons"][f:f + 1], fr["n_persons"][f:f + 1], fr["h_max"])
    assert _same(r["persons3d"][0], ok["persons3d"][f])
