"""Synthetic frame generator — host side of csrc/synth.h through the C ABI."""
import ctypes as C

import numpy as np

from pathlib import Path

from .layouts import SynthConfig, camera_dtype, person2d_dtype

SYNTH_LIB_PATH = Path(__file__).resolve().parent / "libses3d_synth.so"
_synth = None


def _load():
    """libses3d_synth.so: the host generator only (g++-built, no CUDA), separate from the product library."""
    global _synth
    if _synth is None:
        if not SYNTH_LIB_PATH.exists():
            raise FileNotFoundError(f"{SYNTH_LIB_PATH} not built: run `python -m smartedgesensor3dhumanpose_b200.build`")
        L = C.CDLL(str(SYNTH_LIB_PATH))
        L.ses3d_synth_frames.argtypes = [C.c_int32, C.c_void_p, C.POINTER(SynthConfig), C.c_int64, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _synth = L
    return _synth


def synth_config(seed, n_people, p_max=None, dropout=0.0, noise_px=2.0, area=(-2, -2, 2, 2), min_separation=0.6,
                 min_visible=5, frames_per_sequence=0, step_m=1.0 / 30.0):
    """frames_per_sequence = T > 0: frames [qT, (q+1)T) are one scene whose people walk step_m per frame."""
    return SynthConfig(seed, n_people, p_max if p_max is not None else n_people, dropout, noise_px,
                       (C.c_float * 4)(*[float(a) for a in area]), min_separation, min_visible, frames_per_sequence,
                       step_m)


def synth_frames(cameras, cfg: SynthConfig, n_frames, first_frame=0, want_gt=True):
    """Generate frames on the host. Returns dict(persons [F][C][p_max], n_persons [F][C], gt_id, gt_joints)."""
    cams = np.ascontiguousarray(cameras, dtype=camera_dtype)
    n_cams = len(cams)
    persons = np.zeros((n_frames, n_cams, cfg.p_max), person2d_dtype)
    n_persons = np.zeros((n_frames, n_cams), np.int32)
    gt_id = np.full((n_frames, n_cams, cfg.p_max), -1, np.int32) if want_gt else None
    gt_joints = np.zeros((n_frames, cfg.n_people, 17, 3), np.float32) if want_gt else None
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = _load().ses3d_synth_frames(n_cams, p(cams), C.byref(cfg), first_frame, n_frames, p(persons), p(n_persons),
                                    p(gt_id), p(gt_joints))
    if rc != 0:
        raise ValueError(f"ses3d_synth_frames failed with status {rc} (bad configuration)")
    return dict(persons=persons, n_persons=n_persons, gt_id=gt_id, gt_joints=gt_joints)
