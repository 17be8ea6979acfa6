"""Visualisation content: covariance ellipsoids (setMarkerPose S3D:279-310 == PRI:237-254) and the
skeleton LINE_LIST segments (S3D:898-916; addJointToSkeleton PRI:273-382). Eigen is absent and eigenvector signs are
arbitrary, so the ellipsoid is pinned by its invariants against numpy; the segment lists against a line-by-line
python statement of the two reference loops. CPU: serial host instantiation of csrc/markers_core.h."""
import numpy as np
import pytest

from oracle.binding import Oracle, PriorOracle
from smartedgesensor3dhumanpose_b200.layouts import KP2FUSION_SIMPLE, default_prior_params
from tests import helpers
from tests.hostsim.binding import HostSim, hostsim_markers

PARENT_SIMPLE = [-1, 0, 0, 1, 2, 0, 0, 5, 6, 7, 8, 5, 6, 11, 12, 13, 14]   # EdgeTPU_BodyParts_Simple::kpParent S3D:100


def quat_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def check_ellipsoids(persons3d, n_out, ell):
    n = 0
    for f in range(persons3d.shape[0]):
        for p in range(persons3d.shape[1]):
            for k in range(21):
                kp, e = persons3d[f, p]["keypoints"][k], ell[f, p, k]
                if p >= n_out[f] or not kp["score"] > 0:
                    assert not any(e[name] for name in e.dtype.names)
                    continue
                c = kp["cov"]
                S = np.array([[c[0], c[1], c[2]], [c[1], c[3], c[4]], [c[2], c[4], c[5]]])
                if not np.isfinite(S).all():
                    continue
                lam = np.linalg.eigvalsh(S)
                scale = np.array([e["sx"], e["sy"], e["sz"]])
                assert np.allclose(scale, 2 * 2.7955 * np.sqrt(lam), rtol=1e-7, atol=1e-12)
                q = np.array([e["qw"], e["qx"], e["qy"], e["qz"]])
                assert abs(np.linalg.norm(q) - 1) < 1e-9
                R = quat_to_R(q)
                assert abs(np.linalg.det(R) - 1) < 1e-9
                assert np.allclose(R @ np.diag(lam) @ R.T, S, rtol=0, atol=1e-7 * np.abs(S).max())
                n += 1
    return n


def ref_segments_skeleton3d(person):
    """S3D:861-916 for pose_method "simple"."""
    kp = person["keypoints"]
    pts, idx = [], [-1] * 17
    for k in range(17):
        s = KP2FUSION_SIMPLE[k]
        if not kp[s]["score"] > 0:
            continue
        joint = (kp[s]["x"], kp[s]["y"], kp[s]["z"])
        par = PARENT_SIMPLE[k]
        if par >= 0 and idx[par] != -1 and idx[par] < len(pts):
            pts.append(pts[idx[par]])
        else:
            pts.append(joint)
        pts.append(joint)
        idx[k] = len(pts) - 1
    return np.array(pts).reshape(-1, 2, 3)


def ref_segments_pose_prior(person):
    """addJointToSkeleton PRI:273-382 called for the fused joints in ascending slot order (PRI:770, 836)."""
    kp = person["keypoints"]
    pts, idx = [], [-1] * 21
    NOSE, NECK, MIDHIP, HEAD, BELLY = 0, 1, 8, 19, 20

    def ok(s):
        return idx[s] != -1 and idx[s] < len(pts)

    for s in range(21):
        if not kp[s]["score"] > 0:
            continue
        j = (kp[s]["x"], kp[s]["y"], kp[s]["z"])
        if s == NOSE:
            pts += [j, j]
        elif s == HEAD or s in (15, 16, NECK):
            pts += [pts[idx[NOSE]] if ok(NOSE) else j, j]
        elif s in (3, 4, 6, 7, 10, 11, 13, 14):
            pts += [pts[idx[s - 1]] if ok(s - 1) else j, j]
        elif s in (2, 5, MIDHIP):
            pts += [pts[idx[NECK]] if ok(NECK) else (pts[idx[NOSE]] if ok(NOSE) else j), j]
        elif s == BELLY:
            pts += [pts[idx[NECK]] if ok(NECK) else j, j]
            pts += [pts[idx[MIDHIP]] if ok(MIDHIP) else j, j]
        elif s in (9, 12):
            pts += [pts[idx[MIDHIP]] if ok(MIDHIP) else (pts[idx[NECK]] if ok(NECK) else (pts[idx[s - 7]] if ok(s - 7) else j)), j]
        elif s in (17, 18):
            pts += [pts[idx[s - 2]] if ok(s - 2) else j, j]
        idx[s] = len(pts) - 1
    return np.array(pts).reshape(-1, 2, 3)


def _fused_workload():
    fr = helpers.make_sequence_workload("ring8", 2, 16, 3, dropout=0.1)
    orc = Oracle(fr["cameras"], ref_hungarian=True)
    r3 = orc.triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    S, T, H = 2, 16, fr["h_max"]
    rp = PriorOracle(default_prior_params(min_num_obs_track=2), S).run(r3["persons3d"].reshape(S, T, H),
                                                                       r3["n_out"].reshape(S, T), fr["stamp_ns"], None)
    return fr, r3, rp["fused"].reshape(S * T, H), rp["n_out"].reshape(S * T)


def check_all(run, fr, r3, fused, n_fused):
    m = run(r3["persons3d"], r3["n_out"], 0)
    assert check_ellipsoids(r3["persons3d"], r3["n_out"], m["ellipsoids"]) > 500
    n = 0
    for f in range(r3["persons3d"].shape[0]):
        for p in range(r3["n_out"][f]):
            want = ref_segments_skeleton3d(r3["persons3d"][f, p])
            assert m["n_segments"][f, p] == len(want)
            assert np.array_equal(m["segments"][f, p, :len(want)], want)
            n += len(want)
        assert not m["n_segments"][f, r3["n_out"][f]:].any()
    assert n > 500
    m = run(fused, n_fused, 1)
    assert check_ellipsoids(fused, n_fused, m["ellipsoids"]) > 200
    n = 0
    for f in range(fused.shape[0]):
        for p in range(n_fused[f]):
            want = ref_segments_pose_prior(fused[f, p])
            assert m["n_segments"][f, p] == len(want)
            assert np.array_equal(m["segments"][f, p, :len(want)], want)
            n += len(want)
    assert n > 200


def test_markers_device_algorithm():
    fr, r3, fused, n_fused = _fused_workload()
    sim = HostSim(fr["cameras"])
    check_all(lambda p, n, style: hostsim_markers(sim, p, n, style), fr, r3, fused, n_fused)


def test_ellipsoid_of_degenerate_covariances():
    """Isotropic, diagonal, rank-deficient and NaN covariances: finite input gives a valid right-handed frame."""
    fr, r3, _, _ = _fused_workload()
    p3 = r3["persons3d"].copy()
    c = p3["keypoints"]["cov"]
    c[0, 0, 0] = [4e-4, 0, 0, 4e-4, 0, 4e-4]
    c[0, 0, 5] = [9e-4, 0, 0, 1e-4, 0, 4e-4]
    c[0, 0, 6] = [1e-4, 1e-4, 0, 1e-4, 0, 0]
    c[0, 0, 7] = np.nan
    p3["keypoints"]["cov"] = c
    m = hostsim_markers(HostSim(fr["cameras"]), p3, r3["n_out"], 0)
    assert check_ellipsoids(p3, r3["n_out"], m["ellipsoids"]) > 100
    e = m["ellipsoids"][0, 0, 5]
    assert np.allclose(sorted([e["sx"], e["sy"], e["sz"]]), 2 * 2.7955 * np.array([0.01, 0.02, 0.03]))


@pytest.mark.gpu
def test_markers_on_gpu():
    from smartedgesensor3dhumanpose_b200 import api
    fr, r3, fused, n_fused = _fused_workload()
    pipe = api.GeometryPipeline(fr["cameras"])
    check_all(pipe.markers_batch, fr, r3, fused, n_fused)
