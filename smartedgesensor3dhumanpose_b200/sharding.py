"""Frame sharding across the GPUs of one box.

Frames are independent (triangulate_persons reads one frame plus immutable camera tables), so
rank g of G gets the contiguous range [g*N/G, (g+1)*N/G) and the data path needs no collective.
The only exchange is one final gather of compact results (xyz + score per joint) to rank 0,
`gather_compact`, which works on NCCL (CUDA tensors) and gloo (CPU tensors, used by the tests).
"""
import numpy as np

from .layouts import NUM_FUSION_KEYPOINTS, person_cov_dtype


def shard_range(n_frames, rank, world):
    """Contiguous frame range of `rank`: [lo, hi)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (n_frames * rank) // world, (n_frames * (rank + 1)) // world


def compact_numpy(persons3d, n_out):
    """[F][H] PersonCov -> float32 [F][H][21][4] (x, y, z, score); slots beyond n_out are zero."""
    persons3d = np.asarray(persons3d, dtype=person_cov_dtype)
    F, H = persons3d.shape
    kp = persons3d["keypoints"]
    out = np.zeros((F, H, NUM_FUSION_KEYPOINTS, 4), np.float32)
    live = (np.arange(H)[None, :] < np.asarray(n_out)[:, None])[..., None]
    out[..., 0] = np.where(live, kp["x"], 0)
    out[..., 1] = np.where(live, kp["y"], 0)
    out[..., 2] = np.where(live, kp["z"], 0)
    out[..., 3] = np.where(live, kp["score"], 0)
    return out


def compact_torch(raw_u8, n_frames, h_max):
    """Same as compact_numpy on a raw device buffer (uint8 tensor holding [F][H] PersonCov records)."""
    import torch
    rec = raw_u8.view(torch.float64).view(n_frames, h_max, person_cov_dtype.itemsize // 8)
    kp = rec[:, :, 1:1 + NUM_FUSION_KEYPOINTS * 10].reshape(n_frames, h_max, NUM_FUSION_KEYPOINTS, 10)
    xyz = kp[..., 0:3].float()
    score = kp[..., 3].contiguous().view(torch.float32).view(n_frames, h_max, NUM_FUSION_KEYPOINTS, 2)[..., 0:1]
    return torch.cat([xyz, score], dim=-1).contiguous()


def gather_compact(compact, dst=0, shapes=None):
    """Gather every rank's compact tensor on `dst` (list ordered by rank) - the path's only collective.
    `shapes` (one per rank) skips the shape exchange when the caller already knows them."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [compact]
    world, rank = dist.get_world_size(), dist.get_rank()
    if shapes is None:
        shapes = [None] * world
        dist.all_gather_object(shapes, tuple(compact.shape))
    if dist.get_backend() == "nccl":
        # NCCL gather needs equal sizes: pad the frame dimension to the largest shard
        fmax = max(s[0] for s in shapes)
        pad = torch.zeros((fmax,) + tuple(compact.shape[1:]), dtype=compact.dtype, device=compact.device)
        pad[:compact.shape[0]] = compact
        outl = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, outl, dst=dst)
        return [o[:s[0]] for o, s in zip(outl, shapes)] if rank == dst else None
    outl = [torch.empty(s, dtype=compact.dtype) for s in shapes] if rank == dst else None
    if rank == dst:
        outl[dst] = compact
        for r in range(world):
            if r != dst:
                dist.recv(outl[r], src=r)
        return outl
    dist.send(compact, dst=dst)
    return None
