#!/usr/bin/env python
"""Turn the reference's static-transform launch file into a JSON data fixture.

Reads  /root/reference/pose_prior/launch/cameras_extrinsics.launch:2-18  (args: x y z qx qy qz qw
parent child; `base -> cam_1` and `cam_1 -> cam_k`) and writes the row-major 3x4 [R|t] that maps a
base-frame point into each camera's optical frame (what lookupTransform(target=cam, source=base)
returns at S3D:166-167), cameras ordered cam_1..cam_16 as in pose_triangulate_demo.launch:6.

Only the numbers are taken (the rig is data, not code). Run in the build container:
    python scripts/make_rig_fixture.py
"""
import json
import re
import sys
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/pose_prior/launch/cameras_extrinsics.launch")
DST = Path(__file__).resolve().parents[1] / "smartedgesensor3dhumanpose_b200" / "data" / "rig16_hall.json"


def quat_to_R(qx, qy, qz, qw):
    n = np.sqrt(qx * qx + qy * qy + qz * qz + qw * qw)
    qx, qy, qz, qw = qx / n, qy / n, qz / n, qw / n
    return np.array([
        [1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
        [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
        [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])


def main():
    text = SRC.read_text()
    poses = {}
    for m in re.finditer(r'args="([^"]+)"', text):
        a = m.group(1).split()
        vals = [float(v) for v in a[:7]]
        parent, child = a[7].strip("/"), a[8].strip("/")
        T = np.eye(4)
        T[:3, :3] = quat_to_R(*vals[3:7])
        T[:3, 3] = vals[:3]
        poses[child] = (parent, T)
    base_cam1 = poses["cam_1_color_optical_frame"][1]
    cams = []
    for k in range(1, 17):
        parent, T = poses[f"cam_{k}_color_optical_frame"]
        pose = T if parent == "base" else base_cam1 @ T       # pose of cam_k in base
        T_cam_base = np.linalg.inv(pose)[:3, :]                # base -> cam_k
        cams.append({"name": f"cam_{k}", "T_cam_base": [float(v) for v in T_cam_base.reshape(-1)]})
    out = {"source": "pose_prior/launch/cameras_extrinsics.launch:2-18 (reference repo), via scripts/make_rig_fixture.py",
           "note": "T_cam_base = row-major 3x4 [R|t], base -> camera optical frame",
           "cameras": cams}
    DST.write_text(json.dumps(out, indent=1))
    centres = np.array([-np.array(c["T_cam_base"]).reshape(3, 4)[:, :3].T @ np.array(c["T_cam_base"]).reshape(3, 4)[:, 3] for c in cams])
    print("wrote", DST, "centres span", centres.min(0), centres.max(0), file=sys.stderr)


if __name__ == "__main__":
    main()
