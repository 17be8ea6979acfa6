"""Scenario files for the stub ROS runtime (tests/ros_stub/include/ros_stub/core.h) and readers for what a node
published. TEST INFRASTRUCTURE ONLY. The CameraInfo encoder and the MarkerArray decoder below are written against the
ROS 1 .msg definitions, independently of the C++ stub's (de)serialisers."""
import struct

import numpy as np

STRING, BOOL, DOUBLE, INT, STRINGS = range(5)


def _s(x):
    b = x.encode()
    return struct.pack("<I", len(b)) + b


def write_scenario(path, params, tf, msgs):
    """params: {name: value} ('~' prefix for private ones); tf: [(target, source, t[3], q_xyzw[4])];
    msgs: [(deliver_ns, phase, topic, bytes)] delivered in list order within a phase."""
    out = [b"RSTB", struct.pack("<I", 1), struct.pack("<I", len(params))]
    for k, v in params.items():
        out.append(_s(k))
        if isinstance(v, bool):
            out.append(struct.pack("<BB", BOOL, int(v)))
        elif isinstance(v, str):
            out.append(struct.pack("<B", STRING) + _s(v))
        elif isinstance(v, float):
            out.append(struct.pack("<Bd", DOUBLE, v))
        elif isinstance(v, int):
            out.append(struct.pack("<Bq", INT, v))
        else:
            out.append(struct.pack("<BI", STRINGS, len(v)) + b"".join(_s(x) for x in v))
    out.append(struct.pack("<I", len(tf)))
    for target, source, t, q in tf:
        out.append(_s(target) + _s(source) + struct.pack("<3d", *t) + struct.pack("<4d", *q))
    out.append(struct.pack("<I", len(msgs)))
    for deliver_ns, phase, topic, body in msgs:
        out.append(struct.pack("<qi", int(deliver_ns), int(phase)) + _s(topic) + struct.pack("<I", len(body)) + body)
    with open(path, "wb") as f:
        f.write(b"".join(out))


def read_output(path):
    """-> (published [(topic, bytes)], log lines)"""
    data = open(path, "rb").read()
    off = 0

    def u32():
        nonlocal off
        v = struct.unpack_from("<I", data, off)[0]
        off += 4
        return v

    def blob():
        nonlocal off
        n = u32()
        b = data[off:off + n]
        off += n
        return b

    pubs = []
    for _ in range(u32()):
        topic = blob().decode()
        pubs.append((topic, blob()))
    log = [blob().decode() for _ in range(u32())]
    return pubs, log


def header(seq, stamp_ns, frame_id):
    return struct.pack("<III", seq, stamp_ns // 1_000_000_000, stamp_ns % 1_000_000_000) + _s(frame_id)


def encode_camera_info(cam, frame_id, stamp_ns=0):
    """sensor_msgs/CameraInfo of one ses3d_camera record (numpy camera_dtype element): plumb-bob model without
    distortion, K = R-rectified P."""
    fx, fy, cx, cy, Tx, Ty = (float(cam[k]) for k in ("fx", "fy", "cx", "cy", "Tx", "Ty"))
    K = [fx, 0, cx, 0, fy, cy, 0, 0, 1]
    R = [1, 0, 0, 0, 1, 0, 0, 0, 1]
    P = [fx, 0, cx, Tx, 0, fy, cy, Ty, 0, 0, 1, 0]
    D = [0.0] * 5
    return (header(0, stamp_ns, frame_id) + struct.pack("<II", int(cam["height"]), int(cam["width"])) + _s("plumb_bob") +
            struct.pack("<I", len(D)) + struct.pack(f"<{len(D)}d", *D) + struct.pack("<9d", *K) + struct.pack("<9d", *R) +
            struct.pack("<12d", *P) + struct.pack("<II", 0, 0) + struct.pack("<IIIIB", 0, 0, 0, 0, 0))


def rotation_to_quaternion(Rm):
    """xyzw quaternion of a rotation matrix (Shepperd)."""
    Rm = np.asarray(Rm, np.float64)
    t = np.trace(Rm)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        w, x, y, z = 0.25 * s, (Rm[2, 1] - Rm[1, 2]) / s, (Rm[0, 2] - Rm[2, 0]) / s, (Rm[1, 0] - Rm[0, 1]) / s
    else:
        i = int(np.argmax(np.diag(Rm)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + Rm[i, i] - Rm[j, j] - Rm[k, k]) * 2
        q = [0.0, 0.0, 0.0]
        q[i] = 0.25 * s
        q[j] = (Rm[j, i] + Rm[i, j]) / s
        q[k] = (Rm[k, i] + Rm[i, k]) / s
        w = (Rm[k, j] - Rm[j, k]) / s
        x, y, z = q
    return [float(x), float(y), float(z), float(w)]


def quaternion_to_matrix(q):
    """Eigen's Quaternion::toRotationMatrix (what tf2::transformToEigen applies)."""
    x, y, z, w = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz, txx, txy, txz, tyy, tyz, tzz = tx * w, ty * w, tz * w, tx * x, ty * x, tz * x, ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy], [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def decode_marker_array(data):
    """visualization_msgs/MarkerArray -> list of dicts (ROS 1 Marker.msg field order)."""
    off = 0

    def take(fmt):
        nonlocal off
        v = struct.unpack_from("<" + fmt, data, off)
        off += struct.calcsize("<" + fmt)
        return v

    def string():
        nonlocal off
        n = take("I")[0]
        s = data[off:off + n].decode()
        off += n
        return s

    out = []
    for _ in range(take("I")[0]):
        m = {}
        m["seq"], sec, nsec = take("III")
        m["stamp_ns"] = sec * 1_000_000_000 + nsec
        m["frame_id"] = string()
        m["ns"] = string()
        m["id"], m["type"], m["action"] = take("iii")
        m["position"] = take("3d")
        m["orientation_xyzw"] = take("4d")
        m["scale"] = take("3d")
        m["color"] = take("4f")
        ls, lns = take("ii")
        m["lifetime"] = ls + 1e-9 * lns
        m["frame_locked"] = take("B")[0]
        m["points"] = np.array([take("3d") for _ in range(take("I")[0])]).reshape(-1, 3)
        m["colors"] = np.array([take("4f") for _ in range(take("I")[0])]).reshape(-1, 4)
        m["text"] = string()
        m["mesh_resource"] = string()
        m["mesh_use_embedded_materials"] = take("B")[0]
        out.append(m)
    assert off == len(data)
    return out
