"""Named synthetic workloads = BASELINE.json configs 1..5 plus the dense 16-view variant. Shared by
bench.py, the scripts and the tests; inputs come from the stand-alone generator library (synth.py)."""
from . import rigs, synth

# name -> (rig, people, dropout, seed)
CONFIGS = {
    "cfg1_ring4x1": ("ring4", 1, 0.0, 1),
    "cfg2_hall16x6": ("hall16", 6, 0.0, 2),
    "cfg3_hall16x6_dropout": ("hall16", 6, 0.30, 3),
    "cfg4_crowd64x20": ("crowd64", 20, 0.10, 4),
    "cfg5_ring8x4": ("ring8", 4, 0.05, 5),
    "dense_ring16x6": ("ring16", 6, 0.0, 6),
}


def make_workload(name, n_frames, first_frame=0, **over):
    """Frames [first_frame, first_frame + n_frames) of a named workload: dict(persons, n_persons, gt_id, gt_joints,
    cameras, h_max)."""
    rig, people, dropout, seed = CONFIGS[name]
    cams = rigs.RIGS[rig]()
    cfg = synth.synth_config(seed=over.get("seed", seed), n_people=people, dropout=over.get("dropout", dropout),
                             noise_px=over.get("noise_px", 2.0), area=rigs.AREAS[rig])
    fr = synth.synth_frames(cams, cfg, n_frames, first_frame)
    fr["cameras"] = cams
    fr["h_max"] = over.get("h_max", max(8, 2 * people + 4))
    return fr
