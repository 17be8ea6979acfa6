"""CPU test of the live path around the hot path (SURVEY 8 f1 + f2): serialised per-camera Person2DList messages ->
wire decode -> frame assembler -> batch arrays -> (device algorithms, serial test build) -> PersonCovList wire bytes."""
import numpy as np

from smartedgesensor3dhumanpose_b200 import wire
from smartedgesensor3dhumanpose_b200.assembler import FrameAssembler, pack_frames
from tests import helpers
from tests.hostsim.binding import HostSim


def test_wire_to_assembler_to_batch_reproduces_direct_processing():
    n_frames, n_cams = 40, 8
    fr = helpers.make_workload("cfg5_ring8x4", n_frames)
    rng = np.random.default_rng(0)
    # one serialised message per (frame, camera); cameras are synchronous up to 2 ms, messages arrive shuffled
    msgs = []
    for f in range(n_frames):
        for c in range(n_cams):
            stamp = 2_000_000_000 + f * 40_000_000 + int(rng.integers(0, 2_000_000))
            body = wire.encode_person2dlist(fr["persons"][f, c, :fr["n_persons"][f, c]], stamp, f"cam_{c + 1}", seq=f)
            msgs.append((stamp + int(rng.integers(0, 15_000_000)), c, body))
    msgs.sort(key=lambda m: m[0])
    asm = FrameAssembler(n_cams)
    frames = []
    for _, c, body in msgs:
        d = wire.decode_person2dlist(body)
        frames += asm.add(c, d["stamp_ns"], payload=d["persons"])
    assert len(frames) >= n_frames - 2 and asm.stats()["blanked_cameras"] == 0
    persons, n_persons = pack_frames(frames, n_cams, fr["persons"].shape[2])
    seqs = [int((f["stamps_ns"].min() - 2_000_000_000) // 40_000_000) for f in frames]
    assert seqs == sorted(seqs)
    sim = HostSim(fr["cameras"])
    got = sim.triangulate_batch(persons, n_persons, fr["h_max"])
    want = sim.triangulate_batch(fr["persons"][seqs], fr["n_persons"][seqs], fr["h_max"])
    assert np.array_equal(got["n_out"], want["n_out"]) and got["persons3d"].tobytes() == want["persons3d"].tobytes()
    # results go back out as PersonCovList bodies
    f0 = frames[0]
    body = wire.encode_personcovlist(got["persons3d"][0, :got["n_out"][0]], int(f0["stamps_ns"][f0["pivot"]]),
                                     f0["stamps_ns"], np.full(n_cams, 0.1, np.float32))
    back = wire.decode_personcovlist(body)
    assert len(back["persons"]) == got["n_out"][0] and back["stamp_ns"] == int(f0["stamps_ns"].max())
