"""Frame assembler: host-side mirror of the reference's approximate-time synchroniser + worker gating
(my_message_filters/sync_policies/approximate_time_vec.h, S3D:1029-1057, 1218-1223) on top of the C ABI,
and the batcher that packs emitted frames into the [n_frames][C][p_max] arrays of the batch calls."""
import ctypes as C

import numpy as np

from . import lib as _lib
from .layouts import person2d_dtype


class AssemblerConfig(C.Structure):
    _fields_ = [("n_cams", C.c_int32), ("queue_size", C.c_uint32), ("inter_message_lower_bound_ns", C.c_int64),
                ("age_penalty", C.c_double), ("max_interval_ns", C.c_int64), ("max_sync_diff_s", C.c_double)]


class FrameAssembler:
    """add(cam, stamp_ns, payload) -> list of ready frames; a frame = (payloads[C], stamps[C], blank[C], pivot)."""

    def __init__(self, n_cams, **overrides):
        self._L = _lib.load()
        self.cfg = AssemblerConfig()
        _lib.check(self._L.ses3d_assembler_default_config(n_cams, C.byref(self.cfg)))
        for k, v in overrides.items():
            setattr(self.cfg, k, v)
        h = C.c_void_p()
        _lib.check(self._L.ses3d_assembler_create(C.byref(self.cfg), C.byref(h)))
        self._h, self.n_cams = h, n_cams
        self._payloads, self._next_id = {}, 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.ses3d_assembler_destroy(self._h)
            self._h = None

    __del__ = close

    def add(self, cam, stamp_ns, payload=None):
        mid = self._next_id
        self._next_id += 1
        self._payloads[mid] = payload
        n = self._L.ses3d_assembler_add(self._h, cam, int(stamp_ns), mid)
        if n < 0:
            raise ValueError(f"ses3d_assembler_add -> {n}")
        frames = []
        ids = np.zeros(self.n_cams, np.int64)
        stamps = np.zeros(self.n_cams, np.int64)
        blank = np.zeros(self.n_cams, np.uint8)
        pivot = C.c_int32(-1)
        while self._L.ses3d_assembler_pop(self._h, ids.ctypes.data, stamps.ctypes.data, blank.ctypes.data, C.byref(pivot)) == 1:
            frames.append(dict(ids=ids.copy(), payloads=[self._payloads.get(int(i)) for i in ids], stamps_ns=stamps.copy(),
                               blank=blank.astype(bool), pivot=int(pivot.value)))
        if len(self._payloads) > 64 * self.n_cams:   # forget payloads that can no longer be referenced
            keep = sorted(self._payloads)[-32 * self.n_cams:]
            self._payloads = {k: self._payloads[k] for k in keep}
        return frames

    def stats(self):
        st = np.zeros(5, np.int64)
        _lib.check(self._L.ses3d_assembler_stats(self._h, st.ctypes.data))
        return dict(zip(("emitted", "skipped_backwards", "blanked_cameras", "dropped_messages", "signalled"), st.tolist()))


def pack_frames(frames, n_cams, p_max):
    """Emitted frames (payload = array of Person2D per camera) -> persons [F][C][p_max], n_persons [F][C];
    blanked cameras and missing payloads contribute zero persons (the reference's dummy message, S3D:1052-1054)."""
    persons = np.zeros((len(frames), n_cams, p_max), person2d_dtype)
    n_persons = np.zeros((len(frames), n_cams), np.int32)
    for f, fr in enumerate(frames):
        for c in range(n_cams):
            plist = fr["payloads"][c]
            if fr["blank"][c] or plist is None:
                continue
            plist = np.asarray(plist, dtype=person2d_dtype).reshape(-1)[:p_max]
            persons[f, c, :len(plist)] = plist
            n_persons[f, c] = len(plist)
    return persons, n_persons


def mailbox_replay(t_ready_ns, busy_ns):
    """The node's 1-slot latest-wins mailbox (S3D:999-1025) replayed: which frames a worker that needs busy_ns[i] for
    frame i processes when frame i reaches the slot at t_ready_ns[i]. Returns (taken uint8 [n], t_start_ns int64 [n])."""
    t = np.ascontiguousarray(t_ready_ns, dtype=np.int64)
    b = np.ascontiguousarray(busy_ns, dtype=np.int64)
    if t.shape != b.shape or t.ndim != 1:
        raise ValueError("t_ready_ns and busy_ns must be 1-D arrays of the same length")
    taken = np.zeros(len(t), np.uint8)
    start = np.full(len(t), -1, np.int64)
    _lib.check(min(0, _lib.load().ses3d_mailbox_replay(len(t), t.ctypes.data, b.ctypes.data, taken.ctypes.data, start.ctypes.data)))
    return taken, start
