"""ROS-free reader/writer of the person_msgs wire format (ROS 1 serialisation) on top of the C ABI."""
import ctypes as C

import numpy as np

from . import lib as _lib
from .layouts import person2d_dtype, person_cov_dtype


def encode_person2dlist(persons, stamp_ns, frame_id="", fb_delay=0.0, seq=0):
    persons = np.ascontiguousarray(persons, dtype=person2d_dtype).reshape(-1)
    L = _lib.load()
    args = (seq, int(stamp_ns), frame_id.encode(), float(fb_delay), persons.ctypes.data, len(persons))
    n = L.ses3d_wire_encode_person2dlist(*args, None, 0)
    buf = np.zeros(n, np.uint8)
    L.ses3d_wire_encode_person2dlist(*args, buf.ctypes.data, n)
    return buf.tobytes()


def decode_person2dlist(data, cap=64):
    L = _lib.load()
    buf = np.frombuffer(data, np.uint8)
    seq, stamp, fb = C.c_uint32(), C.c_int64(), C.c_float()
    fid = C.create_string_buffer(256)
    persons = np.zeros(cap, person2d_dtype)
    n = L.ses3d_wire_decode_person2dlist(buf.ctypes.data, len(buf), C.byref(seq), C.byref(stamp), fid, 256, C.byref(fb),
                                         persons.ctypes.data, cap)
    if n < 0:
        raise ValueError("malformed person_msgs/Person2DList")
    if n > cap:
        return decode_person2dlist(data, cap=n)
    return dict(seq=seq.value, stamp_ns=stamp.value, frame_id=fid.value.decode(), fb_delay=fb.value, persons=persons[:n].copy())


def encode_personcovlist(persons, stamp_ns, ts_per_cam_ns, fb_delay_per_cam, frame_id="base", seq=0):
    persons = np.ascontiguousarray(persons, dtype=person_cov_dtype).reshape(-1)
    ts = np.ascontiguousarray(ts_per_cam_ns, dtype=np.int64)
    fb = np.ascontiguousarray(fb_delay_per_cam, dtype=np.float32)
    L = _lib.load()
    args = (seq, int(stamp_ns), frame_id.encode(), len(ts), ts.ctypes.data, fb.ctypes.data, persons.ctypes.data, len(persons))
    n = L.ses3d_wire_encode_personcovlist(*args, None, 0)
    buf = np.zeros(n, np.uint8)
    L.ses3d_wire_encode_personcovlist(*args, buf.ctypes.data, n)
    return buf.tobytes()


def decode_personcovlist(data, cap=64, cam_cap=256):
    L = _lib.load()
    buf = np.frombuffer(data, np.uint8)
    seq, stamp, n_cams = C.c_uint32(), C.c_int64(), C.c_int32()
    fid = C.create_string_buffer(256)
    ts = np.zeros(cam_cap, np.int64)
    fb = np.zeros(cam_cap, np.float32)
    persons = np.zeros(cap, person_cov_dtype)
    n = L.ses3d_wire_decode_personcovlist(buf.ctypes.data, len(buf), C.byref(seq), C.byref(stamp), fid, 256, ts.ctypes.data,
                                          fb.ctypes.data, cam_cap, C.byref(n_cams), persons.ctypes.data, cap)
    if n < 0:
        raise ValueError("malformed person_msgs/PersonCovList")
    if n > cap:
        return decode_personcovlist(data, cap=n, cam_cap=cam_cap)
    return dict(seq=seq.value, stamp_ns=stamp.value, frame_id=fid.value.decode(), ts_per_cam_ns=ts[:n_cams.value].copy(),
                fb_delay_per_cam=fb[:n_cams.value].copy(), persons=persons[:n].copy())
