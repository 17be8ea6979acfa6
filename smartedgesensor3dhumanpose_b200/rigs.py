"""Camera rigs used by tests and benchmarks.

`hall16()` is the reference's real 16-camera rig (numbers from
pose_prior/launch/cameras_extrinsics.launch:2-18, stored in data/rig16_hall.json by
scripts/make_rig_fixture.py); the ring rigs are synthetic.
"""
import json
from pathlib import Path

import numpy as np

from .layouts import make_cameras

_DATA = Path(__file__).resolve().parent / "data"


def look_at(eye, target, up=(0.0, 0.0, 1.0)):
    """Row-major 3x4 [R|t] base->camera optical frame (x right, y down, z forward)."""
    eye = np.asarray(eye, dtype=np.float64)
    z = np.asarray(target, dtype=np.float64) - eye
    z /= np.linalg.norm(z)
    x = np.cross(z, np.asarray(up, dtype=np.float64))
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])
    return np.concatenate([R, (-R @ eye)[:, None]], axis=1)


def ring(n_cams, radius, height, target=(0.0, 0.0, 1.0), phase=0.0, **intr):
    T = [look_at((radius * np.cos(phase + 2 * np.pi * k / n_cams), radius * np.sin(phase + 2 * np.pi * k / n_cams), height),
                 target) for k in range(n_cams)]
    return make_cameras(np.stack(T), **intr)


def ring4():
    """Config 1: 4-camera ring, radius 5 m, height 2.5 m, looking at (0,0,1)."""
    return ring(4, 5.0, 2.5)


def ring8():
    """Config 5: 8-camera ring."""
    return ring(8, 6.0, 2.6)


def hall16():
    """Config 2/3: the reference's 16-camera hall rig."""
    d = json.loads((_DATA / "rig16_hall.json").read_text())
    return make_cameras(np.array([c["T_cam_base"] for c in d["cameras"]]))


def ring16():
    """Dense variant of config 2: 16-camera ring in which every camera sees every person (n = 16 views per
    joint, the case the SURVEY 8(d) flop model J(16) describes)."""
    return ring(16, 7.0, 2.6, target=(0.0, 0.0, 1.0))


def crowd64():
    """Config 4: two rings of 32 cameras (radii 8 m / 11 m, heights 2.5 m / 4 m)."""
    a = ring(32, 8.0, 2.5)
    b = ring(32, 11.0, 4.0, phase=np.pi / 32)
    return np.concatenate([a, b])


# floor areas people are placed in (x0, y0, x1, y1), chosen inside each rig's field of view
AREAS = {"ring4": (-1.5, -1.5, 1.5, 1.5), "ring8": (-2.0, -2.0, 2.0, 2.0), "hall16": (-9.0, -4.5, 2.0, 4.5),
         "ring16": (-1.5, -1.5, 1.5, 1.5), "crowd64": (-5.0, -5.0, 5.0, 5.0)}
RIGS = {"ring4": ring4, "ring8": ring8, "hall16": hall16, "ring16": ring16, "crowd64": crowd64}
