"""Host-side mirror of the reference's entry points on top of the C ABI (libses3d.so).

    Skeleton3D.triangulate_persons(people)        <- triangulate_persons(), S3D:525-997 (called S3D:1069)
    PoseReprojection.fused_skeleton_callback(msg) <- fusedSkeletonCallback(), REP:139-235
    PosePrior.skeleton_callback(msg)              <- skeletonCallback(), pose_prior_mult_node.cpp:505-921

plus the batched calls the benchmark uses (`GeometryPipeline.*_batch`). ROS messages are
replaced by numpy structured arrays with the person_msgs layouts (layouts.py). All compute
happens in the CUDA library; there is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import lib as _lib
from .layouts import (DEVICE_BUFFERS, HOST_BUFFERS, AssocDump, camera_dtype, default_params, default_prior_params,
                      person2d_dtype, person_cov_dtype)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class GeometryPipeline:
    """One handle = one camera rig on one GPU (ses3d_create: the set-up of S3D:1184-1214 / REP:272-279)."""

    def __init__(self, cameras, params=None, device=0):
        self._L = _lib.load()
        self.cameras = np.ascontiguousarray(cameras, dtype=camera_dtype)
        self.n_cams = len(self.cameras)
        self.params = params if params is not None else default_params()
        self.device = device
        h = C.c_void_p()
        _lib.check(self._L.ses3d_create(self.n_cams, _p(self.cameras), C.byref(self.params), device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.ses3d_destroy(self._h)
            self._h = None

    __del__ = close

    # ------------------------------------------------------------------ tables
    def tables(self):
        P = np.zeros((self.n_cams, 12), np.float32)
        F = np.zeros((self.n_cams * (self.n_cams - 1) // 2, 9), np.float32)
        _lib.check(self._L.ses3d_get_tables(self._h, _p(P), _p(F)))
        return P, F

    def munkres_batch(self, costs):
        """[n][rows][cols] cost matrices -> [n][rows] assignments through the device Munkres (diagnostics)."""
        costs = np.asarray(costs, dtype=np.float64)
        n, rows, cols = costs.shape
        cm = np.ascontiguousarray(np.transpose(costs, (0, 2, 1)))   # column-major per problem
        out = np.zeros((n, rows), np.int32)
        _lib.check(self._L.ses3d_munkres_batch(self._h, n, rows, cols, _p(cm), _p(out)))
        return out

    def reserve(self, n_frames, p_max, h_max):
        _lib.check(self._L.ses3d_reserve(self._h, n_frames, p_max, h_max))

    @property
    def launch_count(self):
        return int(self._L.ses3d_launch_count(self._h))

    def set_profiling(self, on=True):
        _lib.check(self._L.ses3d_set_profiling(self._h, int(on)))

    def last_kernel_ms(self):
        ms = np.zeros(4, np.float32)
        _lib.check(self._L.ses3d_last_kernel_ms(self._h, _p(ms)))
        return dict(zip(("associate", "triangulate", "finalize", "reproject"), ms.tolist()))

    # ------------------------------------------------------- host-buffer calls
    def triangulate_batch(self, persons, n_persons, h_max, dump=True, out=None, n_out=None):
        persons = np.ascontiguousarray(persons, dtype=person2d_dtype)
        n_frames, n_cams, p_max = persons.shape
        if n_cams != self.n_cams:
            raise ValueError(f"persons has {n_cams} cameras, the rig has {self.n_cams}")
        n_persons = np.ascontiguousarray(n_persons, dtype=np.int32).reshape(n_frames, n_cams)
        out = np.zeros((n_frames, h_max), person_cov_dtype) if out is None else out
        n_out = np.zeros(n_frames, np.int32) if n_out is None else n_out
        res = dict(persons3d=out, n_out=n_out)
        d = None
        if dump:
            res["hyp_of"] = np.full((n_frames, n_cams, p_max), -1, np.int32)
            res["n_hyp"] = np.zeros(n_frames, np.int32)
            res["n_hungarian"] = np.zeros(n_frames, np.int32)
            d = AssocDump(res["hyp_of"].ctypes.data, res["n_hyp"].ctypes.data, res["n_hungarian"].ctypes.data)
        _lib.check(self._L.ses3d_triangulate_batch(self._h, n_frames, p_max, _p(persons), _p(n_persons), h_max, _p(out),
                                                   _p(n_out), C.byref(d) if d else None, HOST_BUFFERS, None))
        return res

    def reproject_batch(self, persons3d, n_persons3d, out=None, n_out=None):
        persons3d = np.ascontiguousarray(persons3d, dtype=person_cov_dtype)
        n_frames, h_max = persons3d.shape
        n_persons3d = np.ascontiguousarray(n_persons3d, dtype=np.int32).reshape(n_frames)
        out = np.zeros((n_frames, self.n_cams, h_max), person2d_dtype) if out is None else out
        n_out = np.zeros((n_frames, self.n_cams), np.int32) if n_out is None else n_out
        _lib.check(self._L.ses3d_reproject_batch(self._h, n_frames, h_max, _p(persons3d), _p(n_persons3d), _p(out),
                                                 _p(n_out), HOST_BUFFERS, None))
        return dict(persons2d=out, n_out=n_out)

    def process_batch(self, persons, n_persons, h_max, want_3d=True, bufs=None):
        """association + triangulation + reprojection chained on the GPU (host buffers in/out)."""
        persons = np.ascontiguousarray(persons, dtype=person2d_dtype)
        n_frames, n_cams, p_max = persons.shape
        n_persons = np.ascontiguousarray(n_persons, dtype=np.int32).reshape(n_frames, n_cams)
        b = bufs or {}
        out3d = b["persons3d"] if "persons3d" in b else (np.zeros((n_frames, h_max), person_cov_dtype) if want_3d else None)
        n3d = b["n_out3d"] if "n_out3d" in b else np.zeros(n_frames, np.int32)
        out2d = b["persons2d"] if "persons2d" in b else np.zeros((n_frames, n_cams, h_max), person2d_dtype)
        n2d = b["n_out2d"] if "n_out2d" in b else np.zeros((n_frames, n_cams), np.int32)
        _lib.check(self._L.ses3d_process_batch(self._h, n_frames, p_max, _p(persons), _p(n_persons), h_max, _p(out3d),
                                               _p(n3d), _p(out2d), _p(n2d), None, HOST_BUFFERS, None))
        return dict(persons3d=out3d, n_out3d=n3d, persons2d=out2d, n_out2d=n2d)

    def process_batch_ragged(self, persons_dense, n_persons, p_max, h_max, out3d, n_out3d, out2d, n_out2d):
        """Ragged form (only occupied records cross PCIe): persons_dense = all detections back to back
        (frame-major, camera-major), n_persons [F][C] run lengths; out3d / out2d are caller-provided dense record
        arrays (capacity = their length). Returns (total3d, total2d)."""
        n_persons = np.ascontiguousarray(n_persons, dtype=np.int32)
        n_frames = n_persons.shape[0]
        t3, t2 = C.c_int64(0), C.c_int64(0)
        _lib.check(self._L.ses3d_process_batch_ragged(self._h, n_frames, p_max, _p(persons_dense), _p(n_persons), h_max,
                                                      _p(out3d), len(out3d), _p(n_out3d), _p(out2d), len(out2d),
                                                      _p(n_out2d), C.byref(t3), C.byref(t2), HOST_BUFFERS))
        return t3.value, t2.value

    def markers_batch(self, persons3d, n_persons3d, style=0):
        """Numeric content of the rviz markers: covariance ellipsoids [F][H][21] (setMarkerPose
        S3D:279-310) and skeleton LINE_LIST segments [F][H][22][2][3]; style 0 = skeleton_3d, 1 = pose_prior."""
        from .layouts import ellipsoid_dtype
        persons3d = np.ascontiguousarray(persons3d, dtype=person_cov_dtype)
        F, H = persons3d.shape
        n_persons3d = np.ascontiguousarray(n_persons3d, dtype=np.int32).reshape(F)
        ell = np.zeros((F, H, 21), ellipsoid_dtype)
        seg = np.zeros((F, H, 22, 2, 3), np.float64)
        n_seg = np.zeros((F, H), np.int32)
        slot = np.zeros((F, H, 22), np.int8)
        _lib.check(self._L.ses3d_markers_batch(self._h, F, H, _p(persons3d), _p(n_persons3d), style, _p(ell), _p(seg),
                                               _p(n_seg), _p(slot), HOST_BUFFERS, None))
        return dict(ellipsoids=ell, segments=seg, n_segments=n_seg, segment_slot=slot)

    def overlay_batch(self, persons, n_persons, width=640, height=480):
        """The overlay images of person_msgs/scripts/pose2D_plot_node.py: persons [N][p_max] Person2D,
        n_persons [N] -> uint8 [N][height][width][3] (rgb8, white background)."""
        persons = np.ascontiguousarray(persons, dtype=person2d_dtype)
        N, p_max = persons.shape
        n_persons = np.ascontiguousarray(n_persons, dtype=np.int32).reshape(N)
        rgb = np.zeros((N, height, width, 3), np.uint8)
        _lib.check(self._L.ses3d_overlay_batch(self._h, N, p_max, _p(persons), _p(n_persons), width, height, _p(rgb),
                                               HOST_BUFFERS, None))
        return rgb

    # ----------------------------------------------------- device-buffer calls
    # Arguments are raw device addresses (e.g. torch_tensor.data_ptr()) on this handle's GPU. These calls are
    # STREAM-ORDERED: they enqueue on `stream` (0 = the legacy default stream) and return without waiting. A capacity
    # overflow is raised by check() or by the next call on the handle.
    def check(self):
        """Wait for outstanding stream-ordered work of this handle and raise if a frame exceeded h_max (ses3d_check)."""
        _lib.check(self._L.ses3d_check(self._h))

    def triangulate_device(self, n_frames, p_max, h_max, persons_ptr, n_persons_ptr, out_ptr, n_out_ptr, stream=0,
                           hyp_of_ptr=0, n_hyp_ptr=0, n_hung_ptr=0):
        d = AssocDump(hyp_of_ptr or None, n_hyp_ptr or None, n_hung_ptr or None)
        _lib.check(self._L.ses3d_triangulate_batch(self._h, n_frames, p_max, persons_ptr, n_persons_ptr, h_max, out_ptr,
                                                   n_out_ptr, C.byref(d), DEVICE_BUFFERS, stream or None))

    def reproject_device(self, n_frames, h_max, persons3d_ptr, n_persons3d_ptr, out_ptr, n_out_ptr, stream=0):
        _lib.check(self._L.ses3d_reproject_batch(self._h, n_frames, h_max, persons3d_ptr, n_persons3d_ptr, out_ptr,
                                                 n_out_ptr, DEVICE_BUFFERS, stream or None))

    def process_device(self, n_frames, p_max, h_max, persons_ptr, n_persons_ptr, out3d_ptr, n_out3d_ptr, out2d_ptr,
                       n_out2d_ptr, stream=0):
        _lib.check(self._L.ses3d_process_batch(self._h, n_frames, p_max, persons_ptr, n_persons_ptr, h_max, out3d_ptr,
                                               n_out3d_ptr, out2d_ptr, n_out2d_ptr, None, DEVICE_BUFFERS,
                                               stream or None))


class MultiPipeline:
    """One rig on several GPUs from one process (ses3d_create_multi): frames are cut into contiguous ranges, one per
    device, each driven by its own host thread bound to the GPU's NUMA node; no data-path collective."""

    def __init__(self, cameras, params=None, devices=None):
        self._L = _lib.load()
        self.cameras = np.ascontiguousarray(cameras, dtype=camera_dtype)
        self.n_cams = len(self.cameras)
        self.params = params if params is not None else default_params()
        dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        _lib.check(self._L.ses3d_create_multi(self.n_cams, _p(self.cameras), C.byref(self.params),
                                              0 if dev is None else len(dev), _p(dev), C.byref(h)))
        self._h = h
        self.n_devices = int(self._L.ses3d_multi_device_count(h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.ses3d_multi_destroy(self._h)
            self._h = None

    __del__ = close

    def reserve(self, n_frames_per_device, p_max, h_max):
        for i in range(self.n_devices):
            _lib.check(self._L.ses3d_reserve(self._L.ses3d_multi_handle(self._h, i), n_frames_per_device, p_max, h_max))

    def process_batch(self, persons, n_persons, h_max, bufs=None, dump=False):
        persons = np.ascontiguousarray(persons, dtype=person2d_dtype)
        n_frames, n_cams, p_max = persons.shape
        n_persons = np.ascontiguousarray(n_persons, dtype=np.int32).reshape(n_frames, n_cams)
        b = bufs or {}
        out3d = b["persons3d"] if "persons3d" in b else np.zeros((n_frames, h_max), person_cov_dtype)
        n3d = b["n_out3d"] if "n_out3d" in b else np.zeros(n_frames, np.int32)
        out2d = b["persons2d"] if "persons2d" in b else np.zeros((n_frames, n_cams, h_max), person2d_dtype)
        n2d = b["n_out2d"] if "n_out2d" in b else np.zeros((n_frames, n_cams), np.int32)
        res = dict(persons3d=out3d, n_out3d=n3d, persons2d=out2d, n_out2d=n2d)
        d = None
        if dump:
            res["hyp_of"] = np.full((n_frames, n_cams, p_max), -1, np.int32)
            res["n_hyp"] = np.zeros(n_frames, np.int32)
            res["n_hungarian"] = np.zeros(n_frames, np.int32)
            d = AssocDump(res["hyp_of"].ctypes.data, res["n_hyp"].ctypes.data, res["n_hungarian"].ctypes.data)
        _lib.check(self._L.ses3d_multi_process_batch(self._h, n_frames, p_max, _p(persons), _p(n_persons), h_max, _p(out3d),
                                                     _p(n3d), _p(out2d), _p(n2d), C.byref(d) if d else None))
        return res

    def process_batch_ragged(self, persons_dense, n_persons, p_max, h_max, out3d, n_out3d, out2d, n_out2d):
        """Dense records in; dense records out as one segment per device. Returns (seg3d, seg2d), each [n_devices][2] =
        (first record index, record count) of the device's segment in out3d / out2d."""
        n_persons = np.ascontiguousarray(n_persons, dtype=np.int32)
        n_frames = n_persons.shape[0]
        seg3 = np.zeros((self.n_devices, 2), np.int64)
        seg2 = np.zeros((self.n_devices, 2), np.int64)
        _lib.check(self._L.ses3d_multi_process_batch_ragged(self._h, n_frames, p_max, _p(persons_dense), _p(n_persons),
                                                            h_max, _p(out3d), len(out3d), _p(n_out3d), _p(out2d),
                                                            len(out2d), _p(n_out2d), _p(seg3), _p(seg2)))
        return seg3, seg2


def to_ragged(persons, n_persons):
    """[F][C][p_max] padded detections -> dense record array in frame-major, camera-major order."""
    persons = np.asarray(persons, dtype=person2d_dtype)
    live = np.arange(persons.shape[2])[None, None, :] < np.asarray(n_persons)[:, :, None]
    return np.ascontiguousarray(persons[live])


def from_ragged(dense, counts, cap, dtype):
    """dense records + run lengths counts[...] -> padded array counts.shape + (cap,)."""
    counts = np.asarray(counts)
    out = np.zeros(counts.shape + (cap,), dtype)
    live = np.arange(cap).reshape((1,) * counts.ndim + (cap,)) < counts[..., None]
    out[live] = dense[:int(counts.sum())]
    return out


def _pack_frame(people, n_cams):
    """list (per camera) of Person2D arrays -> ([1][C][p_max] array, [1][C] counts)."""
    if len(people) != n_cams:
        raise ValueError(f"expected one Person2DList per camera ({n_cams}), got {len(people)}")  # assert S3D:534
    p_max = max(1, max(len(p) for p in people))
    persons = np.zeros((1, n_cams, p_max), person2d_dtype)
    n_persons = np.zeros((1, n_cams), np.int32)
    for c, plist in enumerate(people):
        plist = np.asarray(plist, dtype=person2d_dtype).reshape(-1)
        persons[0, c, :len(plist)] = plist
        n_persons[0, c] = len(plist)
    return persons, n_persons


class Skeleton3D:
    """skeleton_3d node body: per-frame triangulate_persons (S3D:525-997) with n_frames = 1."""

    def __init__(self, cameras, params=None, device=0, h_max=32):
        self.pipe = GeometryPipeline(cameras, params, device)
        self.h_max = h_max

    def triangulate_persons(self, people):
        """people: one array of Person2D per camera (a Person2DList each). Returns the PersonCov array
        (persons3d_msg.persons); fewer than two cameras with detections gives an empty list (S3D:557-560)."""
        persons, n_persons = _pack_frame(people, self.pipe.n_cams)
        r = self.pipe.triangulate_batch(persons, n_persons, self.h_max, dump=False)
        return r["persons3d"][0, :r["n_out"][0]].copy()


class PoseReprojection:
    """pose_reprojection node body: fusedSkeletonCallback (REP:139-235) with n_frames = 1."""

    def __init__(self, cameras, params=None, device=0):
        self.pipe = GeometryPipeline(cameras, params, device)

    def fused_skeleton_callback(self, persons3d):
        """persons3d: PersonCov array (PersonCovList.persons). Returns one Person2D array per camera."""
        persons3d = np.asarray(persons3d, dtype=person_cov_dtype).reshape(1, -1)
        n = persons3d.shape[1]
        if n == 0:
            return [np.zeros(0, person2d_dtype) for _ in range(self.pipe.n_cams)]
        r = self.pipe.reproject_batch(persons3d, np.array([n], np.int32))
        return [r["persons2d"][0, c, :r["n_out"][0, c]].copy() for c in range(self.pipe.n_cams)]


class PriorTracker:
    """pose_prior for n_sequences independent message streams on one GPU (ses3d_prior_*): tracking, skeleton-model
    fit, marginal covariances and prediction of pose_prior_mult_node.cpp:505-921. State persists between calls."""

    def __init__(self, params=None, n_sequences=1, max_tracks=32, device=0):
        self._L = _lib.load()
        self.params = params if params is not None else default_prior_params()
        self.n_sequences, self.max_tracks, self.device = n_sequences, max_tracks, device
        h = C.c_void_p()
        _lib.check(self._L.ses3d_prior_create(C.byref(self.params), n_sequences, max_tracks, device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.ses3d_prior_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        """reset(), PRI:182-189."""
        _lib.check(self._L.ses3d_prior_reset(self._h))

    @property
    def launch_count(self):
        return int(self._L.ses3d_prior_launch_count(self._h))

    def last_kernel_ms(self):
        ms = C.c_float(0)
        _lib.check(self._L.ses3d_prior_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def tracks(self, sequence=0):
        ids = np.zeros(self.max_tracks, np.int32)
        nobs = np.zeros(self.max_tracks, np.int32)
        n = self._L.ses3d_prior_get_tracks(self._h, sequence, _p(ids), _p(nobs))
        if n < 0:
            _lib.check(n)
        return ids[:n].copy(), nobs[:n].copy()

    def run(self, persons, n_persons, stamp_ns, fb_delay=None, want_track_of=True, out=None):
        """persons [S][T][h_max] PersonCov, n_persons [S][T], stamp_ns [S][T] int64, fb_delay [S][T][n_cams] float32
        or None. Host buffers in and out; `out` may hold caller-provided (e.g. pinned) "fused" / "pred" arrays."""
        persons = np.ascontiguousarray(persons, dtype=person_cov_dtype)
        S, T, H = persons.shape
        n_persons = np.ascontiguousarray(n_persons, dtype=np.int32).reshape(S, T)
        stamp_ns = np.ascontiguousarray(stamp_ns, dtype=np.int64).reshape(S, T)
        n_cams = 0
        if fb_delay is not None:
            fb_delay = np.ascontiguousarray(fb_delay, dtype=np.float32).reshape(S, T, -1)
            n_cams = fb_delay.shape[-1]
        out = out or {}
        fused = out["fused"] if "fused" in out else np.zeros((S, T, H), person_cov_dtype)
        pred = out["pred"] if "pred" in out else np.zeros((S, T, H), person_cov_dtype)
        n_out = np.zeros((S, T), np.int32)
        pred_delay = np.zeros((S, T), np.float32)
        track_of = np.full((S, T, H), -1, np.int32) if want_track_of else None
        _lib.check(self._L.ses3d_prior_run(self._h, S, T, H, _p(persons), _p(n_persons), _p(stamp_ns), n_cams,
                                           _p(fb_delay), _p(fused), _p(pred), _p(n_out), _p(pred_delay), _p(track_of),
                                           HOST_BUFFERS, None))
        return dict(fused=fused, pred=pred, n_out=n_out, pred_delay=pred_delay, track_of=track_of)

    def run_ragged(self, persons_dense, n_persons, stamp_ns, h_max, fused_dense, pred_dense, fb_delay=None):
        """Ragged form (only occupied records cross PCIe): persons_dense = all input records back to back
        (stream-major, message-major), n_persons [S][T] run lengths; fused_dense / pred_dense are caller-provided record
        arrays (capacity = their length). Returns (n_out [S][T], pred_delay [S][T], total)."""
        n_persons = np.ascontiguousarray(n_persons, dtype=np.int32)
        S, T = n_persons.shape
        stamp_ns = np.ascontiguousarray(stamp_ns, dtype=np.int64).reshape(S, T)
        n_cams = 0
        if fb_delay is not None:
            fb_delay = np.ascontiguousarray(fb_delay, dtype=np.float32).reshape(S, T, -1)
            n_cams = fb_delay.shape[-1]
        n_out = np.zeros((S, T), np.int32)
        pred_delay = np.zeros((S, T), np.float32)
        total = C.c_int64(0)
        _lib.check(self._L.ses3d_prior_run_ragged(self._h, S, T, h_max, _p(persons_dense), _p(n_persons), _p(stamp_ns),
                                                  n_cams, _p(fb_delay), _p(fused_dense), _p(pred_dense),
                                                  min(len(fused_dense), len(pred_dense)), _p(n_out), _p(pred_delay),
                                                  C.byref(total)))
        return n_out, pred_delay, total.value

    def run_device(self, n_sequences, n_frames, h_max, persons_ptr, n_persons_ptr, stamp_ptr, n_cams, fb_delay_ptr,
                   fused_ptr, pred_ptr, n_out_ptr, pred_delay_ptr=0, track_of_ptr=0, stream=0):
        """Raw device addresses on this handle's GPU; stream-ordered."""
        _lib.check(self._L.ses3d_prior_run(self._h, n_sequences, n_frames, h_max, persons_ptr, n_persons_ptr, stamp_ptr,
                                           n_cams, fb_delay_ptr or None, fused_ptr, pred_ptr, n_out_ptr,
                                           pred_delay_ptr or None, track_of_ptr or None, DEVICE_BUFFERS, stream or None))


class PosePrior:
    """pose_prior node body: skeletonCallback (PRI:505-921) for one message at a time (n_sequences = n_frames = 1)."""

    def __init__(self, params=None, device=0, h_max=32, max_tracks=32):
        self.tracker = PriorTracker(params, 1, max_tracks, device)
        self.h_max = h_max

    def skeleton_callback(self, persons3d, stamp_ns, fb_delay_per_cam=None):
        """persons3d: PersonCov array (PersonCovList.persons), stamp_ns: header.stamp, fb_delay_per_cam: float array.
        Returns (persons3d_fused, persons3d_fused_pred, fb_delay) as published on PRI:906-907."""
        persons3d = np.asarray(persons3d, dtype=person_cov_dtype).reshape(-1)
        n = len(persons3d)
        if n > self.h_max:
            raise ValueError(f"{n} persons exceed h_max = {self.h_max}")
        buf = np.zeros((1, 1, self.h_max), person_cov_dtype)
        buf[0, 0, :n] = persons3d
        fb = None if fb_delay_per_cam is None else np.asarray(fb_delay_per_cam, np.float32).reshape(1, 1, -1)
        r = self.tracker.run(buf, np.array([[n]], np.int32), np.array([[stamp_ns]], np.int64), fb, want_track_of=False)
        k = int(r["n_out"][0, 0])
        return r["fused"][0, 0, :k].copy(), r["pred"][0, 0, :k].copy(), float(r["pred_delay"][0, 0])

    def reset(self):
        self.tracker.reset()
