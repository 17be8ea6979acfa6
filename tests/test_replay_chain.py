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


def test_full_demo_replay_over_the_wire():
    """The demo wiring end to end on recorded-like message bodies: per-camera Person2DList bytes -> assembler ->
    skeleton_3d -> PersonCovList bytes (persons_3d) -> pose_prior -> PersonCovList bytes (persons3d_fused_pred, with the
    predicted delay in fb_delay_per_cam, PRI:531) -> pose_reprojection -> per-camera Person2DList bytes. Every hop
    goes through the wire codec; the result equals processing the arrays directly."""
    from smartedgesensor3dhumanpose_b200.layouts import default_prior_params
    from tests.hostsim.binding import PriorHostSim
    T, n_cams = 24, 8
    fr = helpers.make_sequence_workload("ring8", 1, T, 3)
    sim = HostSim(fr["cameras"])
    prm = default_prior_params(min_num_obs_track=3)
    # reference: arrays straight through
    direct = helpers.run_demo_chain(sim, PriorHostSim(prm, 1), fr)
    # replay: one message per (frame, camera), then one PersonCovList per frame between the nodes
    asm = FrameAssembler(n_cams)
    prior = PriorHostSim(prm, 1)
    H, PM = fr["h_max"], fr["persons"].shape[2]
    n_checked = 0
    for f in range(T):
        frames = []
        for c in range(n_cams):
            stamp = int(fr["stamp_ns"][0, f])
            body = wire.encode_person2dlist(fr["persons"][f, c, :fr["n_persons"][f, c]], stamp, f"cam_{c + 1}", seq=f)
            d = wire.decode_person2dlist(body)
            frames += asm.add(c, d["stamp_ns"], payload=d["persons"])
        assert len(frames) == (1 if f > 0 else 0) or f == 0   # synchronous cameras: a frame closes when the next one starts
        for fm in frames:
            persons, n_persons = pack_frames([fm], n_cams, PM)
            r3 = sim.triangulate_batch(persons, n_persons, H)
            stamp = int(fm["stamps_ns"][fm["pivot"]])
            body3 = wire.encode_personcovlist(r3["persons3d"][0, :r3["n_out"][0]], stamp, fm["stamps_ns"],
                                              np.full(n_cams, -1.0, np.float32))
            m3 = wire.decode_personcovlist(body3)
            buf = np.zeros((1, 1, H), r3["persons3d"].dtype)
            buf[0, 0, :len(m3["persons"])] = m3["persons"]
            rp = prior.run(buf, np.array([[len(m3["persons"])]], np.int32), np.array([[m3["stamp_ns"]]], np.int64),
                           m3["fb_delay_per_cam"][None, None, :])
            k = int(rp["n_out"][0, 0])
            body_pred = wire.encode_personcovlist(rp["pred"][0, 0, :k], m3["stamp_ns"], m3["ts_per_cam_ns"],
                                                  np.full(n_cams, rp["pred_delay"][0, 0], np.float32))
            mp = wire.decode_personcovlist(body_pred)
            assert np.allclose(mp["fb_delay_per_cam"], 0.1)            # no delay measured -> g_avg_delay
            p3 = np.zeros((1, H), r3["persons3d"].dtype)
            p3[0, :k] = mp["persons"]
            r2 = sim.reproject_batch(p3, np.array([k], np.int32))
            t = int(round((m3["stamp_ns"] - int(fr["stamp_ns"][0, 0])) * 30 / 1e9))
            assert np.array_equal(r2["n_out"][0], direct[2]["n_out"][t])
            for c in range(n_cams):
                n2 = r2["n_out"][0, c]
                back = wire.decode_person2dlist(wire.encode_person2dlist(r2["persons2d"][0, c, :n2], m3["stamp_ns"], f"cam_{c + 1}"))
                assert back["persons"].tobytes() == direct[2]["persons2d"][t, c, :n2].tobytes()
                n_checked += n2
    assert n_checked > 50
