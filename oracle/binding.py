"""ctypes binding of the CPU oracle (oracle/libses3d_oracle.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from smartedgesensor3dhumanpose_b200.layouts import (Params, camera_dtype, default_params, person2d_dtype,
                                                     person_cov_dtype)

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libses3d_oracle.so"
REF_HUNGARIAN_PATH = HERE / "_ref" / "libref_hungarian.so"


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    srcs = [HERE / "ses3d_oracle.cpp", HERE / "pose_prior_oracle.cpp", HERE.parent / "include" / "ses3d.h"]
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(HERE), "-s"], check=True)
    elif not REF_HUNGARIAN_PATH.exists() and Path("/root/reference/skeleton_3d/src/Hungarian.cpp").exists():
        subprocess.run(["make", "-C", str(HERE), "-s", "ref"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int32, C.c_void_p, C.POINTER(Params)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_use_ref_hungarian.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_get_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_munkres.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_triangulate_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_int32]
        L.oracle_triangulate_batch_ex.argtypes = L.oracle_triangulate_batch.argtypes + [C.c_void_p]
        L.oracle_set_svd_variant.argtypes = [C.c_void_p, C.c_int32]
        L.oracle_triangulate_point_v.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_reproject_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int32]
        L.oracle_triangulate_point.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                               C.c_void_p]
        L.oracle_lm_refine.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.oracle_ut_covariance.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU restatement of triangulate_persons (S3D:525-997) + fusedSkeletonCallback (REP:139-235)."""

    def __init__(self, cameras, params=None, ref_hungarian=False, svd_variant=0):
        self.cameras = np.ascontiguousarray(cameras, dtype=camera_dtype)
        self.params = params if params is not None else default_params()
        self.n_cams = len(self.cameras)
        self._h = lib().oracle_create(self.n_cams, _p(self.cameras), C.byref(self.params))
        if not self._h:
            raise ValueError("oracle_create failed")
        # 0 = one-sided Hestenes Jacobi (primary oracle), 1 = Eigen 3.3 JacobiSVD restatement (S3D:456)
        lib().oracle_set_svd_variant(self._h, int(svd_variant))
        self.ref_hungarian = False
        if ref_hungarian:
            if not REF_HUNGARIAN_PATH.exists():
                raise FileNotFoundError(REF_HUNGARIAN_PATH)
            rc = lib().oracle_use_ref_hungarian(self._h, str(REF_HUNGARIAN_PATH).encode())
            if rc != 0:
                raise RuntimeError(f"oracle_use_ref_hungarian -> {rc}")
            self.ref_hungarian = True

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def tables(self):
        P = np.zeros((self.n_cams, 12), np.float32)
        F = np.zeros((self.n_cams * (self.n_cams - 1) // 2, 9), np.float32)
        lib().oracle_get_tables(self._h, _p(P), _p(F))
        return P, F

    def triangulate_batch(self, persons, n_persons, h_max, n_threads=1, diag=False):
        """diag=True adds per-frame 'margin' (smallest relative distance of a floating-point branch decision of the
        frame to its threshold) and 'cond' (largest sigma_1/sigma_3 of a weighted DLT system of the frame)."""
        persons = np.ascontiguousarray(persons, dtype=person2d_dtype)
        n_frames, n_cams, p_max = persons.shape
        assert n_cams == self.n_cams
        n_persons = np.ascontiguousarray(n_persons, dtype=np.int32).reshape(n_frames, n_cams)
        out = np.zeros((n_frames, h_max), person_cov_dtype)
        n_out = np.zeros(n_frames, np.int32)
        hyp_of = np.full((n_frames, n_cams, p_max), -1, np.int32)
        n_hyp = np.zeros(n_frames, np.int32)
        n_hung = np.zeros(n_frames, np.int32)
        n_joints = np.zeros(1, np.int64)
        dg = np.zeros((n_frames, 2), np.float64) if diag else None
        rc = lib().oracle_triangulate_batch_ex(self._h, n_frames, p_max, _p(persons), _p(n_persons), h_max, _p(out),
                                               _p(n_out), _p(hyp_of), _p(n_hyp), _p(n_hung), _p(n_joints), n_threads,
                                               _p(dg))
        r = dict(status=rc, persons3d=out, n_out=n_out, hyp_of=hyp_of, n_hyp=n_hyp, n_hungarian=n_hung,
                 n_joints=int(n_joints[0]))
        if diag:
            r["margin"], r["cond"] = dg[:, 0].copy(), dg[:, 1].copy()
        return r

    def reproject_batch(self, persons3d, n_persons3d, n_threads=1):
        persons3d = np.ascontiguousarray(persons3d, dtype=person_cov_dtype)
        n_frames, h_max = persons3d.shape
        n_persons3d = np.ascontiguousarray(n_persons3d, dtype=np.int32)
        out = np.zeros((n_frames, self.n_cams, h_max), person2d_dtype)
        n_out = np.zeros((n_frames, self.n_cams), np.int32)
        lib().oracle_reproject_batch(self._h, n_frames, h_max, _p(persons3d), _p(n_persons3d), _p(out), _p(n_out),
                                     n_threads)
        return dict(persons2d=out, n_out=n_out)


def munkres(cost):
    """Restated Munkres on a (rows, cols) matrix; returns (assignment[rows], total cost)."""
    cost = np.asarray(cost, dtype=np.float64)
    r, c = cost.shape
    cm = np.asfortranarray(cost).ravel(order="F").copy()
    a = np.zeros(r, np.int32)
    total = np.zeros(1, np.float64)
    lib().oracle_munkres(_p(a), _p(total), _p(cm), r, c)
    return a, float(total[0])


_ref_lib = None


def ref_munkres(cost):
    """The reference's verbatim Hungarian.cpp (oracle/_ref)."""
    global _ref_lib
    if _ref_lib is None:
        build()
        _ref_lib = C.CDLL(str(REF_HUNGARIAN_PATH))
        _ref_lib.ref_hungarian_assignmentoptimal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    cost = np.asarray(cost, dtype=np.float64)
    r, c = cost.shape
    cm = np.asfortranarray(cost).ravel(order="F").copy()
    a = np.zeros(r, np.int32)
    total = np.zeros(1, np.float64)
    _ref_lib.ref_hungarian_assignmentoptimal(_p(a), _p(total), _p(cm), r, c)
    return a, float(total[0])


def triangulate_point(P, pts, weighted=True, use_double=False):
    P = np.ascontiguousarray(P, np.float64).reshape(-1, 12)
    pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    X = np.zeros(3)
    e = np.zeros(1)
    lib().oracle_triangulate_point(len(P), _p(P), _p(pts), int(weighted), int(use_double), _p(X), _p(e))
    return X, float(e[0])


def triangulate_point_v(P, pts, weighted=True, use_double=False, svd_variant=0):
    """Single DLT solve with the chosen SVD variant; returns (X, reprojection error, singular values descending)."""
    P = np.ascontiguousarray(P, np.float64).reshape(-1, 12)
    pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    X, e, sv = np.zeros(3), np.zeros(1), np.zeros(4)
    lib().oracle_triangulate_point_v(len(P), _p(P), _p(pts), int(weighted), int(use_double), int(svd_variant), _p(X),
                                     _p(e), _p(sv))
    return X, float(e[0]), sv


def lm_refine(P, pts, X0, max_iters=10):
    P = np.ascontiguousarray(P, np.float64).reshape(-1, 12)
    pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    X = np.array(X0, np.float64).copy()
    lib().oracle_lm_refine(len(P), _p(P), _p(pts), max_iters, _p(X))
    return X


def ut_covariance(P, pts, cov2d, mean):
    P = np.ascontiguousarray(P, np.float64).reshape(-1, 12)
    pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    cov2d = np.ascontiguousarray(cov2d, np.float64).reshape(-1, 3)
    mean = np.ascontiguousarray(mean, np.float64)
    cov = np.zeros(9)
    lib().oracle_ut_covariance(len(P), _p(P), _p(pts), _p(cov2d), _p(mean), _p(cov))
    return cov.reshape(3, 3)


# ------------------------------------------------------------------------------------------------ pose_prior
class PriorOracle:
    """CPU restatement of pose_prior's skeletonCallback (pose_prior_mult_node.cpp:505-921), oracle/pose_prior_oracle.cpp."""

    def __init__(self, params=None, n_sequences=1, ref_hungarian=False):
        from smartedgesensor3dhumanpose_b200.layouts import PriorParams, default_prior_params
        L = lib()
        L.prior_oracle_create.restype = C.c_void_p
        L.prior_oracle_create.argtypes = [C.POINTER(PriorParams), C.c_int32, C.c_char_p]
        L.prior_oracle_destroy.argtypes = [C.c_void_p]
        L.prior_oracle_reset.argtypes = [C.c_void_p]
        L.prior_oracle_run.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 3 + [C.c_int32] + \
            [C.c_void_p] * 6 + [C.c_int32]
        L.prior_oracle_get_tracks.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.prior_oracle_stats.argtypes = [C.c_void_p, C.c_void_p]
        self._L = L
        self.params = params if params is not None else default_prior_params()
        self.n_sequences = n_sequences
        so = None
        if ref_hungarian:
            if not REF_HUNGARIAN_PATH.exists():
                raise FileNotFoundError(REF_HUNGARIAN_PATH)
            so = str(REF_HUNGARIAN_PATH).encode()
        self._h = L.prior_oracle_create(C.byref(self.params), n_sequences, so)
        if not self._h:
            raise ValueError("prior_oracle_create failed")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.prior_oracle_destroy(self._h)
            self._h = None

    def reset(self):
        self._L.prior_oracle_reset(self._h)

    def run(self, persons, n_persons, stamp_ns, fb_delay=None, n_threads=1):
        """persons [S][T][h_max], n_persons [S][T], stamp_ns [S][T], fb_delay [S][T][n_cams] or None."""
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
        rc = self._L.prior_oracle_run(self._h, S, T, H, _p(persons), _p(n_persons), _p(stamp_ns), n_cams, _p(fb_delay),
                                      _p(fused), _p(pred), _p(n_out), _p(pred_delay), _p(track_of), n_threads)
        if rc != 0:
            raise RuntimeError(f"prior_oracle_run -> {rc}")
        return dict(fused=fused, pred=pred, n_out=n_out, pred_delay=pred_delay, track_of=track_of)

    def tracks(self, sequence=0):
        ids = np.zeros(1024, np.int32)
        nobs = np.zeros(1024, np.int32)
        n = self._L.prior_oracle_get_tracks(self._h, sequence, _p(ids), _p(nobs))
        return ids[:n].copy(), nobs[:n].copy()

    def stats(self):
        out = np.zeros(3, np.int64)
        self._L.prior_oracle_stats(self._h, _p(out))
        return dict(fits=int(out[0]), lm_outer=int(out[1]), lm_inner=int(out[2]))
