"""CPU tests of the person_msgs wire (de)serialiser: C++ against an independent struct-based
implementation of the ROS 1 encoding, round trips, and malformed-buffer handling."""
import numpy as np
import pytest

from oracle import wire_ref
from smartedgesensor3dhumanpose_b200 import wire
from smartedgesensor3dhumanpose_b200.layouts import person2d_dtype, person_cov_dtype
from tests import helpers


def _p2d_to_ref(p):
    return dict(score=float(p["score"]), bbox=tuple(float(v) for v in p["bbox"]),
                keypoints=[(float(k["x"]), float(k["y"]), float(k["score"]), *[float(c) for c in k["cov"]]) for k in p["keypoints"]])


def test_person2dlist_matches_struct_reference_and_round_trips():
    fr = helpers.make_workload("cfg2_hall16x6", 4)
    for f in range(4):
        for c in range(16):
            persons = fr["persons"][f, c, :fr["n_persons"][f, c]]
            stamp = 1_700_000_000_123_456_789 + f * 33_000_000 + c
            data = wire.encode_person2dlist(persons, stamp, frame_id=f"cam_{c + 1}_color_optical_frame", fb_delay=0.1, seq=f)
            ref = wire_ref.encode_person2dlist([_p2d_to_ref(p) for p in persons], stamp, f"cam_{c + 1}_color_optical_frame", 0.1, f)
            assert data == ref
            assert len(data) == 16 + len(f"cam_{c + 1}_color_optical_frame") + 8 + len(persons) * 432
            d = wire.decode_person2dlist(data)
            assert (d["seq"], d["stamp_ns"], d["frame_id"]) == (f, stamp, f"cam_{c + 1}_color_optical_frame")
            assert abs(d["fb_delay"] - 0.1) < 1e-7 and d["persons"].tobytes() == persons.tobytes()
            r = wire_ref.decode_person2dlist(data)
            assert len(r["persons"]) == len(persons)


def test_personcovlist_matches_struct_reference_and_round_trips():
    from tests.hostsim.binding import HostSim
    fr = helpers.make_workload("cfg5_ring8x4", 6)
    res = HostSim(fr["cameras"]).triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    for f in range(6):
        persons = res["persons3d"][f, :res["n_out"][f]].copy()
        persons["id"] = np.arange(len(persons))
        ts = 1_700_000_000_000_000_000 + np.arange(8) * 1_000_003
        fb = np.linspace(0.05, 0.15, 8).astype(np.float32)
        data = wire.encode_personcovlist(persons, int(ts[3]), ts, fb, frame_id="base", seq=7)
        ref = wire_ref.encode_personcovlist(
            [dict(id=int(p["id"]), score=float(p["score"]),
                  keypoints=[(k["x"], k["y"], k["z"], float(k["score"]), *k["cov"]) for k in p["keypoints"]],
                  bbox_center=tuple(p["bbox_center"]), bbox_size=tuple(p["bbox_size"])) for p in persons],
            int(ts[3]), [int(t) for t in ts], [float(v) for v in fb], "base", 7)
        assert data == ref
        assert len(data) == 16 + 4 + 4 + 8 * 8 + 4 + 8 * 4 + 4 + len(persons) * 1688
        d = wire.decode_personcovlist(data)
        assert d["frame_id"] == "base" and np.array_equal(d["ts_per_cam_ns"], ts) and np.array_equal(d["fb_delay_per_cam"], fb)
        got, want = d["persons"], persons
        for name in ("id", "score", "bbox_center", "bbox_size"):
            assert np.array_equal(got[name], want[name])
        for name in ("x", "y", "z", "score", "cov"):
            assert np.array_equal(got["keypoints"][name], want["keypoints"][name])


def test_malformed_buffers_are_rejected():
    fr = helpers.make_workload("cfg5_ring8x4", 1)
    data = wire.encode_person2dlist(fr["persons"][0, 0, :fr["n_persons"][0, 0]], 5, "x")
    for cut in (0, 3, 15, 20, len(data) - 1):
        with pytest.raises(ValueError):
            wire.decode_person2dlist(data[:cut])
    # a Person2D with 16 keypoints cannot be held by the fixed-size POD
    bad = wire_ref.encode_person2dlist([dict(score=1.0, keypoints=[(0.0,) * 6] * 16, bbox=(0, 0, 0, 0))], 5, "x")
    with pytest.raises(ValueError):
        wire.decode_person2dlist(bad)
    # a PersonCov with another keypoint count decodes to an empty skeleton (skipped by the reference, REP:166-169)
    odd = wire_ref.encode_personcovlist([dict(id=3, score=0.5, keypoints=[(1.0, 2.0, 3.0, 0.9) + (0.1,) * 6] * 5,
                                              bbox_center=(0,) * 7, bbox_size=(0,) * 3)], 9, [1, 2], [0.1, 0.2])
    d = wire.decode_personcovlist(odd)
    assert len(d["persons"]) == 1 and d["persons"][0]["id"] == 3 and not d["persons"][0]["keypoints"]["score"].any()


def test_empty_lists():
    d = wire.decode_person2dlist(wire.encode_person2dlist(np.zeros(0, person2d_dtype), 123456789, "cam"))
    assert len(d["persons"]) == 0 and d["stamp_ns"] == 123456789
    d = wire.decode_personcovlist(wire.encode_personcovlist(np.zeros(0, person_cov_dtype), 1, [], []))
    assert len(d["persons"]) == 0 and len(d["ts_per_cam_ns"]) == 0
