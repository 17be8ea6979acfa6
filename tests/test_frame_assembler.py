"""CPU tests of the frame assembler: C++ (libses3d.so, host only) against the pure-Python
restatement of the reference's synchroniser + worker gating on randomised message streams, and against
properties of the policy."""
import numpy as np
import pytest

from oracle.frame_assembler_ref import RefAssembler
from smartedgesensor3dhumanpose_b200.assembler import FrameAssembler, pack_frames
from smartedgesensor3dhumanpose_b200.layouts import person2d_dtype


def _stream(rng, n_cams, n_ticks, period_ns=33_333_333, jitter_ns=4_000_000, drop=0.05, late=0.03, burst=0.0):
    """Camera messages at ~30 Hz with per-camera phase, jitter, random drops, occasional late (stale) stamps, and
    arrival order = stamp + transport delay."""
    msgs = []
    phase = rng.integers(0, period_ns // 2, n_cams)
    for t in range(n_ticks):
        for c in range(n_cams):
            if rng.random() < drop:
                continue
            stamp = 1_000_000_000 + t * period_ns + int(phase[c]) + int(rng.integers(-jitter_ns, jitter_ns + 1))
            if rng.random() < late:
                stamp -= int(rng.integers(60_000_000, 150_000_000))
            arrival = stamp + int(rng.integers(1_000_000, 30_000_000)) + (int(rng.integers(0, 200_000_000)) if rng.random() < burst else 0)
            msgs.append((arrival, c, max(stamp, 1)))
    msgs.sort()
    return [(c, s) for _, c, s in msgs]


@pytest.mark.parametrize("n_cams,seed,kw", [(4, 0, {}), (16, 1, {}), (16, 2, dict(drop=0.2, late=0.1)), (8, 3, dict(burst=0.1)),
                                            (5, 4, dict(jitter_ns=30_000_000)), (16, 5, dict(drop=0.0, late=0.0, jitter_ns=0)),
                                            (3, 6, dict(drop=0.3, late=0.2, burst=0.2))])
def test_cpp_assembler_matches_python_restatement(n_cams, seed, kw):
    rng = np.random.default_rng(seed)
    stream = _stream(rng, n_cams, 400, **kw)
    a, r = FrameAssembler(n_cams), RefAssembler(n_cams)
    got, want = [], []
    for mid, (cam, stamp) in enumerate(stream):
        frames = a.add(cam, stamp, payload=mid)
        n_ref = r.add(cam, stamp, mid)
        assert len(frames) == n_ref
        got += frames
    want = r.ready
    assert len(got) == len(want) and len(got) > 50
    for g, w in zip(got, want):
        assert g["payloads"] == w["ids"] and g["stamps_ns"].tolist() == w["stamps_ns"]
        assert g["blank"].tolist() == w["blank"] and g["pivot"] == w["pivot"]
    assert a.stats() == r.stats


def test_frames_are_time_consistent_and_use_each_message_once():
    rng = np.random.default_rng(7)
    n_cams = 16
    a = FrameAssembler(n_cams)
    seen, last_pivot = set(), 0
    n_frames = 0
    for mid, (cam, stamp) in enumerate(_stream(rng, n_cams, 300)):
        for fr in a.add(cam, stamp, payload=(cam, mid)):
            n_frames += 1
            for c, p in enumerate(fr["payloads"]):
                assert p[0] == c and p not in seen          # one message per camera, never reused
                seen.add(p)
            pivot_stamp = fr["stamps_ns"][fr["pivot"]]
            assert pivot_stamp == fr["stamps_ns"].max() and pivot_stamp > last_pivot     # S3D:1029-1046
            last_pivot = pivot_stamp
            lag = (pivot_stamp - fr["stamps_ns"]) * 1e-9
            assert np.array_equal(fr["blank"], lag > 0.067)                               # S3D:1049-1057
    st = a.stats()
    assert st["emitted"] == n_frames and n_frames > 150


def test_perfectly_synchronous_cameras_give_one_frame_per_tick():
    a = FrameAssembler(4)
    out = []
    for t in range(50):
        for c in range(4):
            out += a.add(c, 1_000_000_000 + t * 40_000_000, payload=(t, c))
    ticks = [fr["payloads"][0][0] for fr in out]
    assert ticks == sorted(ticks) and len(set(ticks)) == len(ticks) >= 47
    for fr in out:
        assert len({p[0] for p in fr["payloads"]}) == 1 and not fr["blank"].any()


def test_default_config_is_the_reference_setup():
    a = FrameAssembler(16)
    assert (a.cfg.queue_size, a.cfg.inter_message_lower_bound_ns, a.cfg.age_penalty, a.cfg.max_sync_diff_s) == (5, 20_000_000, 2.0, 0.067)
    assert FrameAssembler(4).cfg.queue_size == 3                      # std::max(3u, 1 + 4/4)  S3D:1219


def test_pack_frames_blanks_lagging_cameras():
    people = np.zeros(2, person2d_dtype)
    people["score"] = [0.5, 0.9]
    frames = [dict(payloads=[people, people[:1], None], blank=np.array([False, True, False]))]
    persons, n_persons = pack_frames(frames, 3, 4)
    assert n_persons.tolist() == [[2, 0, 0]] and persons[0, 0, 1]["score"] == np.float32(0.9)
