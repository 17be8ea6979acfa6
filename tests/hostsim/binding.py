"""ctypes binding of tests/hostsim/libhostsim.so — the CUDA library's device algorithms run
serially on the CPU. Test-only; see hostsim.cpp."""
import ctypes as C

import numpy as np

from smartedgesensor3dhumanpose_b200.layouts import camera_dtype, default_params, person2d_dtype, person_cov_dtype

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(_build.build()))
        L.hostsim_create.restype = C.c_void_p
        L.hostsim_create.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]
        L.hostsim_destroy.argtypes = [C.c_void_p]
        L.hostsim_get_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hostsim_triangulate_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hostsim_reproject_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        L.hostsim_set_big_rig_path.argtypes = [C.c_int32]
        L.hostsim_markers.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 4
        L.hostsim_prior_create.restype = C.c_void_p
        L.hostsim_prior_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.hostsim_prior_destroy.argtypes = [C.c_void_p]
        L.hostsim_prior_run.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 3 + [C.c_int32] + \
            [C.c_void_p] * 6 + [C.c_int32]
        L.hostsim_prior_get_tracks.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostSim:
    def __init__(self, cameras, params=None):
        self.cameras = np.ascontiguousarray(cameras, dtype=camera_dtype)
        self.params = params if params is not None else default_params()
        self.n_cams = len(self.cameras)
        self._h = lib().hostsim_create(self.n_cams, _p(self.cameras), C.byref(self.params))
        if not self._h:
            raise ValueError("hostsim_create failed")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hostsim_destroy(self._h)
            self._h = None

    def tables(self):
        P = np.zeros((self.n_cams, 12), np.float32)
        F = np.zeros((self.n_cams * (self.n_cams - 1) // 2, 9), np.float32)
        lib().hostsim_get_tables(self._h, _p(P), _p(F))
        return P, F

    def triangulate_batch(self, persons, n_persons, h_max):
        persons = np.ascontiguousarray(persons, dtype=person2d_dtype)
        n_frames, n_cams, p_max = persons.shape
        n_persons = np.ascontiguousarray(n_persons, dtype=np.int32).reshape(n_frames, n_cams)
        out = np.zeros((n_frames, h_max), person_cov_dtype)
        n_out = np.zeros(n_frames, np.int32)
        hyp_of = np.full((n_frames, n_cams, p_max), -1, np.int32)
        n_hyp = np.zeros(n_frames, np.int32)
        n_hung = np.zeros(n_frames, np.int32)
        rc = lib().hostsim_triangulate_batch(self._h, n_frames, p_max, _p(persons), _p(n_persons), h_max, _p(out),
                                             _p(n_out), _p(hyp_of), _p(n_hyp), _p(n_hung))
        return dict(status=rc, persons3d=out, n_out=n_out, hyp_of=hyp_of, n_hyp=n_hyp, n_hungarian=n_hung)

    def reproject_batch(self, persons3d, n_persons3d, cam_tile=0):
        persons3d = np.ascontiguousarray(persons3d, dtype=person_cov_dtype)
        n_frames, h_max = persons3d.shape
        n_persons3d = np.ascontiguousarray(n_persons3d, dtype=np.int32)
        out = np.zeros((n_frames, self.n_cams, h_max), person2d_dtype)
        n_out = np.zeros((n_frames, self.n_cams), np.int32)
        lib().hostsim_reproject_batch(self._h, n_frames, h_max, cam_tile, _p(persons3d), _p(n_persons3d), _p(out),
                                      _p(n_out))
        return dict(persons2d=out, n_out=n_out)


def set_big_rig_path(on):
    """Run the association the way the kernel does for rigs whose frame does not fit shared memory."""
    lib().hostsim_set_big_rig_path(int(on))


def _markers(call, persons3d, n_out, style):
    from smartedgesensor3dhumanpose_b200.layouts import ellipsoid_dtype
    persons3d = np.ascontiguousarray(persons3d, dtype=person_cov_dtype)
    F, H = persons3d.shape
    n_out = np.ascontiguousarray(n_out, dtype=np.int32).reshape(F)
    ell = np.zeros((F, H, 21), ellipsoid_dtype)
    seg = np.zeros((F, H, 22, 2, 3), np.float64)
    n_seg = np.zeros((F, H), np.int32)
    slot = np.zeros((F, H, 22), np.int8)
    call(F, H, persons3d, n_out, style, ell, seg, n_seg, slot)
    return dict(ellipsoids=ell, segments=seg, n_segments=n_seg, segment_slot=slot)


def hostsim_markers(sim, persons3d, n_out, style=0):
    return _markers(lambda F, H, p, n, st, e, s, ns, sl: lib().hostsim_markers(sim._h, F, H, _p(p), _p(n), st, _p(e), _p(s),
                                                                              _p(ns), _p(sl)), persons3d, n_out, style)


class PriorHostSim:
    """prior_core.h (the device algorithm of ses3d_prior_run) instantiated with the serial team."""

    def __init__(self, params=None, n_sequences=1, max_tracks=32, group=6):
        from smartedgesensor3dhumanpose_b200.layouts import default_prior_params
        self.params = params if params is not None else default_prior_params()
        self.max_tracks = max_tracks
        self.group = group   # detections fitted together by one warp on the GPU
        self._h = lib().hostsim_prior_create(C.byref(self.params), n_sequences, max_tracks)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hostsim_prior_destroy(self._h)
            self._h = None

    def run(self, persons, n_persons, stamp_ns, fb_delay=None):
        persons = np.ascontiguousarray(persons, dtype=person_cov_dtype)
        S, T, H = persons.shape
        n_persons = np.ascontiguousarray(n_persons, dtype=np.int32).reshape(S, T)
        stamp_ns = np.ascontiguousarray(stamp_ns, dtype=np.int64).reshape(S, T)
        n_cams = 0
        if fb_delay is not None:
            fb_delay = np.ascontiguousarray(fb_delay, dtype=np.float32)
            n_cams = fb_delay.shape[-1]
        fused = np.zeros((S, T, H), person_cov_dtype)
        pred = np.zeros((S, T, H), person_cov_dtype)
        n_out = np.zeros((S, T), np.int32)
        pred_delay = np.zeros((S, T), np.float32)
        track_of = np.full((S, T, H), -1, np.int32)
        rc = lib().hostsim_prior_run(self._h, S, T, H, _p(persons), _p(n_persons), _p(stamp_ns), n_cams, _p(fb_delay),
                                     _p(fused), _p(pred), _p(n_out), _p(pred_delay), _p(track_of), self.group)
        if rc != 0:
            raise RuntimeError(f"hostsim_prior_run -> {rc}")
        return dict(fused=fused, pred=pred, n_out=n_out, pred_delay=pred_delay, track_of=track_of)

    def tracks(self, sequence=0):
        ids = np.zeros(self.max_tracks, np.int32)
        nobs = np.zeros(self.max_tracks, np.int32)
        n = lib().hostsim_prior_get_tracks(self._h, sequence, _p(ids), _p(nobs))
        return ids[:n].copy(), nobs[:n].copy()
