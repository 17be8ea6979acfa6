"""CPU test of the N > 1 path: frames sharded across ranks with no data-path collective and one final gather
(SURVEY 8(e)), world_size 2 over gloo. Each rank generates its own shard with the counter-based generator and
runs the test-only serial instantiation of the device algorithms; rank 0 checks the gathered result against a
single-process run over all frames."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smartedgesensor3dhumanpose_b200 import sharding
from tests import helpers
from tests.hostsim.binding import HostSim

N_FRAMES, WORLD = 37, 2   # ragged split: 18 + 19


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(N_FRAMES, rank, world)
    fr = helpers.make_workload("cfg5_ring8x4", hi - lo, first_frame=lo)
    r = HostSim(fr["cameras"]).triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    compact = torch.from_numpy(sharding.compact_numpy(r["persons3d"], r["n_out"]))
    parts = sharding.gather_compact(compact, dst=0)
    if rank == 0:
        q.put(torch.cat(parts).numpy())
    else:
        assert parts is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 37, 1000):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_gloo_sharded_run_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, WORLD, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    fr = helpers.make_workload("cfg5_ring8x4", N_FRAMES)
    r = HostSim(fr["cameras"]).triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    assert np.array_equal(gathered, sharding.compact_numpy(r["persons3d"], r["n_out"]))


def test_compact_torch_matches_numpy():
    fr = helpers.make_workload("cfg5_ring8x4", 16)
    r = HostSim(fr["cameras"]).triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    # zero the dead slots like the numpy version does before comparing (the raw buffer keeps stale records there)
    p3 = r["persons3d"].copy()
    dead = np.arange(p3.shape[1])[None, :] >= r["n_out"][:, None]
    p3[dead] = np.zeros((), p3.dtype)
    raw = torch.from_numpy(p3.view(np.uint8).reshape(-1).copy())
    got = sharding.compact_torch(raw, p3.shape[0], p3.shape[1]).numpy()
    assert np.array_equal(got, sharding.compact_numpy(p3, r["n_out"]))


# ---- pose_prior: message streams sharded across ranks (independent trackers), final gather of the fused skeletons
N_STREAMS, N_MSG = 5, 14   # ragged split: 2 + 3 streams


def _prior_worker(rank, world, port, q):
    from smartedgesensor3dhumanpose_b200.layouts import default_prior_params
    from smartedgesensor3dhumanpose_b200.sequences import synth_person_sequences
    from tests.hostsim.binding import PriorHostSim
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(N_STREAMS, rank, world)
    seq = synth_person_sequences(N_STREAMS, N_MSG, 3, seed=51)          # every rank cuts its streams out of the same set
    prm = default_prior_params(min_num_obs_track=2)
    r = PriorHostSim(prm, hi - lo).run(seq["persons"][lo:hi], seq["n_persons"][lo:hi], seq["stamp_ns"][lo:hi],
                                       seq["fb_delay"][lo:hi])
    H = r["fused"].shape[-1]
    compact = torch.from_numpy(sharding.compact_numpy(r["fused"].reshape(-1, H), r["n_out"].reshape(-1)))
    parts = sharding.gather_compact(compact, dst=0)
    if rank == 0:
        q.put(torch.cat(parts).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_prior_streams_equal_single_process():
    from smartedgesensor3dhumanpose_b200.layouts import default_prior_params
    from smartedgesensor3dhumanpose_b200.sequences import synth_person_sequences
    from tests.hostsim.binding import PriorHostSim
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_prior_worker, args=(r, WORLD, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    seq = synth_person_sequences(N_STREAMS, N_MSG, 3, seed=51)
    r = PriorHostSim(default_prior_params(min_num_obs_track=2), N_STREAMS).run(seq["persons"], seq["n_persons"],
                                                                               seq["stamp_ns"], seq["fb_delay"])
    H = r["fused"].shape[-1]
    want = sharding.compact_numpy(r["fused"].reshape(-1, H), r["n_out"].reshape(-1))
    assert gathered.shape == want.shape and np.array_equal(gathered, want)
    assert (want[..., 3] > 0).sum() > 100
