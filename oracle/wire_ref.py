"""Independent `struct`-based encoder/decoder of the ROS 1 wire format of person_msgs/Person2DList and
PersonCovList (person_msgs/msg/*.msg). TEST INFRASTRUCTURE ONLY: checker for csrc/wire.cpp."""
import struct


def _header(seq, stamp_ns, frame_id):
    f = frame_id.encode()
    return struct.pack("<IIII", seq, stamp_ns // 10**9, stamp_ns % 10**9, len(f)) + f


def encode_person2dlist(persons, stamp_ns, frame_id="", fb_delay=0.0, seq=0):
    """persons: list of dict(score, keypoints=[(x, y, score, cxx, cxy, cyy)] * n, bbox=(x0, y0, x1, y1))"""
    out = _header(seq, stamp_ns, frame_id) + struct.pack("<fI", fb_delay, len(persons))
    for p in persons:
        out += struct.pack("<fI", p["score"], len(p["keypoints"]))
        for kp in p["keypoints"]:
            out += struct.pack("<6f", *kp)
        out += struct.pack("<4f", *p["bbox"])
    return out


def decode_person2dlist(data):
    seq, sec, nsec, flen = struct.unpack_from("<IIII", data, 0)
    off = 16
    frame_id = data[off:off + flen].decode(); off += flen
    fb, n = struct.unpack_from("<fI", data, off); off += 8
    persons = []
    for _ in range(n):
        score, nk = struct.unpack_from("<fI", data, off); off += 8
        kps = [struct.unpack_from("<6f", data, off + 24 * k) for k in range(nk)]; off += 24 * nk
        bbox = struct.unpack_from("<4f", data, off); off += 16
        persons.append(dict(score=score, keypoints=kps, bbox=bbox))
    assert off == len(data)
    return dict(seq=seq, stamp_ns=sec * 10**9 + nsec, frame_id=frame_id, fb_delay=fb, persons=persons)


def encode_personcovlist(persons, stamp_ns, ts_per_cam_ns, fb_delay_per_cam, frame_id="base", seq=0):
    """persons: list of dict(id, score, keypoints=[(x, y, z, score, c0..c5)] * n, bbox_center=(7), bbox_size=(3))"""
    out = _header(seq, stamp_ns, frame_id) + struct.pack("<I", len(ts_per_cam_ns))
    for t in ts_per_cam_ns:
        out += struct.pack("<II", t // 10**9, t % 10**9)
    out += struct.pack("<I", len(fb_delay_per_cam)) + struct.pack(f"<{len(fb_delay_per_cam)}f", *fb_delay_per_cam)
    out += struct.pack("<I", len(persons))
    for p in persons:
        out += struct.pack("<IfI", p["id"], p["score"], len(p["keypoints"]))
        for kp in p["keypoints"]:
            out += struct.pack("<3df6d", *kp)
        out += struct.pack("<7d", *p["bbox_center"]) + struct.pack("<3d", *p["bbox_size"])
    return out
