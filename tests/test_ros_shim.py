"""SURVEY 8 f2 (+ f1 mailbox, f4 marker assembly): the shim nodes under ros_shim/ — the files a maintainer drops into the
catkin workspace in place of the three reference nodes — compiled UNCHANGED against a stub ROS runtime (tests/ros_stub;
ROS itself is absent from this image) and run on recorded-like message streams. What the nodes publish must equal what
the library returns for the same frames when called directly, record for record.

CPU suite: the node sources compile and link; without a GPU they fail loudly (no CPU path). GPU suite: the replays."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from smartedgesensor3dhumanpose_b200 import rigs, wire
from smartedgesensor3dhumanpose_b200.layouts import KP2FUSION_SIMPLE, camera_dtype
from tests import helpers
from tests.ros_stub import build as stub_build
from tests.ros_stub import scenario as sc

ROOT = Path(__file__).resolve().parents[1]
KP2FUSION = list(KP2FUSION_SIMPLE)
T0 = 2_000_000_000
DT = 40_000_000   # 25 Hz


@pytest.fixture(scope="module")
def nodes():
    return stub_build.build()


def _gpu_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def rig_through_tf(cams):
    """The rig as the nodes see it: extrinsics travel as (translation, quaternion) tf transforms; returns the tf
    entries and the camera table rebuilt from them with Eigen's quaternion -> matrix formula (what
    ses3d_ros::make_camera does), so the direct calls use bit-identical tables."""
    cams = np.array(cams, dtype=camera_dtype)
    tf, names = [], []
    for i in range(len(cams)):
        T = np.asarray(cams[i]["T_cam_base"], np.float64).reshape(3, 4)
        q = sc.rotation_to_quaternion(T[:, :3])
        name = f"cam_{i + 1}"
        names.append(name)
        tf.append((name + "_color_optical_frame", "base", [float(x) for x in T[:, 3]], q))
        R = sc.quaternion_to_matrix(q)
        cams[i]["T_cam_base"] = np.concatenate([R, T[:, 3:4]], axis=1).reshape(-1)
    return cams, tf, names


def camera_info_msgs(cams, names):
    return [(0, 0, f"{n}/color/camera_info", sc.encode_camera_info(cams[i], n + "_color_optical_frame"))
            for i, n in enumerate(names)]


def run_node(exe, tmp_path, params, tf, msgs, tag):
    scn, out = tmp_path / f"{tag}.scn", tmp_path / f"{tag}.out"
    sc.write_scenario(scn, params, tf, msgs)
    r = subprocess.run([str(exe), str(scn), str(out)], capture_output=True, text=True, timeout=300)
    pubs, log = sc.read_output(out) if out.exists() else ([], [])
    return r.returncode, pubs, log


# ---------------------------------------------------------------------------------------------- CPU suite
def test_shim_nodes_compile_against_the_stub_ros(nodes):
    for name, exe in nodes.items():
        assert exe.exists(), name
    # every ROS-facing name the reference nodes use appears in the shim sources (topics, node names, parameters)
    src = {p.name: p.read_text() for p in (ROOT / "ros_shim" / "src").glob("*.cpp")}
    for needle in ("skeleton_singlePerson_3d", "human_pose_estimation/persons3d", "human_pose_estimation/skeleton3d_vis",
                   "/human_joints", "max_epi_dist", "vis_cov", "pose_method", "cameras"):
        assert needle in src["skeleton_3d_ses3d_node.cpp"], needle
    for needle in ("multi_skeleton_reprojection", "human_pose_estimation/persons3d_fused_pred", "/skel_pred"):
        assert needle in src["pose_reproj_ses3d_node.cpp"], needle
    for needle in ("pose_prior", "human_pose_estimation/persons3d_fused", "human_pose_estimation/skeleton3d_fused",
                   "norm_height"):
        assert needle in src["pose_prior_ses3d_node.cpp"], needle
    # no reference-side heavy dependency is left in the shim
    for text in list(src.values()) + [(ROOT / "ros_shim" / "include" / "ses3d_ros" / "convert.h").read_text()]:
        includes = [l for l in text.splitlines() if l.lstrip().startswith("#include")]
        for banned in ("Eigen", "gtsam", "image_geometry", "Hungarian", "omp.h", "cv_bridge", "opencv"):
            assert not any(banned in l for l in includes), banned


@pytest.mark.skipif(_gpu_present(), reason="needs a box without a GPU")
def test_shim_node_fails_loudly_without_a_gpu(nodes, tmp_path):
    cams, tf, names = rig_through_tf(rigs.ring4())
    rc, pubs, log = run_node(nodes["skeleton_3d_ses3d_node"], tmp_path, {"~cameras": names}, tf,
                             camera_info_msgs(cams, names), "nogpu")
    assert rc != 0 and not pubs
    assert any("ses3d_create" in l and "no CPU path" in l for l in log), log


def test_mailbox_model_matches_the_python_restatement():
    """ses3d_mailbox_replay: the 1-slot latest-wins mailbox between the synchroniser callback and the worker
    (S3D:999-1025) as a deterministic replay; against oracle/frame_assembler_ref.py::mailbox_replay."""
    from oracle.frame_assembler_ref import mailbox_replay
    from smartedgesensor3dhumanpose_b200.assembler import mailbox_replay as lib_replay
    rng = np.random.default_rng(5)
    for trial in range(200):
        n = int(rng.integers(1, 60))
        t_ready = np.cumsum(rng.integers(0, 50_000_000, n)).astype(np.int64)
        busy = rng.integers(1_000_000, 120_000_000, n).astype(np.int64)
        if trial % 3 == 0:
            busy[:] = int(rng.integers(1_000_000, 90_000_000))
        want_taken, want_start = mailbox_replay(t_ready.tolist(), busy.tolist())
        taken, start = lib_replay(t_ready, busy)
        assert taken.tolist() == want_taken and start[taken.astype(bool)].tolist() == [s for s, t in zip(want_start, want_taken) if t]
        assert taken[-1] == 1            # the newest frame is never lost
    # a worker faster than the frame period drops nothing; a slow one keeps every k-th frame
    t = np.arange(20, dtype=np.int64) * DT
    assert lib_replay(t, np.full(20, DT // 2, np.int64))[0].all()
    slow = lib_replay(t, np.full(20, int(2.5 * DT), np.int64))[0]
    assert slow.sum() < 10 and slow[0] == 1 and slow[-1] == 1


# ---------------------------------------------------------------------------------------------- GPU suite
def person2d_stream(fr, names, jitter_ns=1_500_000, seed=0):
    """One Person2DList per (frame, camera), delivered in stamp order; fb_delay distinct per camera."""
    rng = np.random.default_rng(seed)
    F, C = fr["n_persons"].shape
    msgs, stamps = [], np.zeros((F, C), np.int64)
    for f in range(F):
        for c in range(C):
            stamp = T0 + f * DT + int(rng.integers(0, jitter_ns))
            stamps[f, c] = stamp
            body = wire.encode_person2dlist(fr["persons"][f, c, :fr["n_persons"][f, c]], stamp, names[c] + "_color_optical_frame",
                                            fb_delay=0.01 * (c + 1), seq=f)
            msgs.append((stamp + int(rng.integers(0, 8_000_000)), 1, f"{names[c]}/human_joints", body))
    msgs.sort(key=lambda m: m[0])
    return msgs, stamps


@pytest.mark.gpu
def test_skeleton_3d_node_replay_equals_direct_calls(nodes, tmp_path):
    from smartedgesensor3dhumanpose_b200 import api
    F = 30
    fr = helpers.make_workload("cfg5_ring8x4", F)
    cams, tf, names = rig_through_tf(fr["cameras"])
    msgs, stamps = person2d_stream(fr, names)
    params = {"~cameras": names, "~lossless": True, "~vis_cov": True, "~max_epi_dist": 0.05, "~h_max": fr["h_max"]}
    rc, pubs, log = run_node(nodes["skeleton_3d_ses3d_node"], tmp_path, params, tf, camera_info_msgs(cams, names) + msgs, "s3d")
    assert rc == 0, log
    lists = [wire.decode_personcovlist(b) for t, b in pubs if t == "human_pose_estimation/persons3d"]
    vis = [sc.decode_marker_array(b) for t, b in pubs if t == "human_pose_estimation/skeleton3d_vis"]
    assert len(lists) >= F - 2
    pivot_stamp = stamps.max(axis=1)
    pipe = api.GeometryPipeline(cams)
    direct = pipe.triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    seen, n_people, vi = [], 0, 0
    for m in lists:
        f = int(np.nonzero(pivot_stamp == m["stamp_ns"])[0][0])   # header = the pivot camera's header (S3D:1060)
        seen.append(f)
        assert m["frame_id"] == "base" and m["ts_per_cam_ns"].tolist() == stamps[f].tolist()
        assert np.allclose(m["fb_delay_per_cam"], 0.01 * (np.arange(len(names)) + 1))
        n = int(direct["n_out"][f])
        assert len(m["persons"]) == n
        assert m["persons"].tobytes() == direct["persons3d"][f, :n].tobytes()   # record for record, bit for bit
        n_people += n
        if n:   # marker message: a LINE_LIST and a SPHERE_LIST per person plus the covariance spheres (S3D:688-715, 885-916)
            mk = vis[vi]
            vi += 1
            want = pipe.markers_batch(direct["persons3d"][f:f + 1], direct["n_out"][f:f + 1], style=0)
            lines = [x for x in mk if x["ns"] == "joints"]
            spheres = [x for x in mk if x["ns"] == "joint_spheres"]
            covs = [x for x in mk if x["ns"] == "joint_cov_3d"]
            assert len(lines) == n and len(spheres) == n
            for p in range(n):
                ns = int(want["n_segments"][0, p])
                assert lines[p]["type"] == 5 and lines[p]["points"].shape == (2 * ns, 3) and lines[p]["lifetime"] == 2.0
                assert np.array_equal(lines[p]["points"].reshape(ns, 2, 3), want["segments"][0, p, :ns])
                kp = direct["persons3d"][f, p]["keypoints"]
                assert spheres[p]["type"] == 7 and len(spheres[p]["points"]) == int((kp["score"][KP2FUSION] > 0).sum())
            n_cov = sum(int(((direct["persons3d"][f, p]["keypoints"]["score"] > 0)[[s for s in KP2FUSION if s < 15]]).sum())
                        for p in range(n))
            assert len(covs) == n_cov and all(abs(c["color"][3] - 0.5) < 1e-7 and c["lifetime"] == 5.0 for c in covs)
            e = want["ellipsoids"][0]
            for c in covs:
                p, k = divmod(c["id"], 21)
                slot = KP2FUSION[k]
                assert c["scale"] == (e[p, slot]["sx"], e[p, slot]["sy"], e[p, slot]["sz"])
    assert seen == sorted(seen) and len(set(seen)) == len(seen) and n_people > 0 and vi == len(vis)
    pipe.close()


@pytest.mark.gpu
def test_skeleton_3d_node_latest_wins_mailbox_only_drops_frames(nodes, tmp_path):
    from smartedgesensor3dhumanpose_b200 import api
    F = 40
    fr = helpers.make_workload("cfg5_ring8x4", F)
    cams, tf, names = rig_through_tf(fr["cameras"])
    msgs, stamps = person2d_stream(fr, names, seed=1)
    rc, pubs, log = run_node(nodes["skeleton_3d_ses3d_node"], tmp_path, {"~cameras": names, "~h_max": fr["h_max"]}, tf,
                             camera_info_msgs(cams, names) + msgs, "s3d_lw")
    assert rc == 0, log
    lists = [wire.decode_personcovlist(b) for t, b in pubs if t == "human_pose_estimation/persons3d"]
    pipe = api.GeometryPipeline(cams)
    direct = pipe.triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    pivot_stamp = stamps.max(axis=1)
    seen = [int(np.nonzero(pivot_stamp == m["stamp_ns"])[0][0]) for m in lists]
    # the replay floods the node, so frames are overwritten in the slot - but what is processed is processed correctly,
    # in order, and the newest synchronised frame always gets through (S3D:999-1025)
    assert 1 <= len(seen) <= F and seen == sorted(set(seen)) and seen[-1] >= F - 2
    for m, f in zip(lists, seen):
        n = int(direct["n_out"][f])
        assert m["persons"].tobytes() == direct["persons3d"][f, :n].tobytes()
    pipe.close()


@pytest.mark.gpu
def test_reprojection_node_replay_equals_direct_calls(nodes, tmp_path):
    from smartedgesensor3dhumanpose_b200 import api
    F = 20
    fr = helpers.make_workload("cfg5_ring8x4", F)
    cams, tf, names = rig_through_tf(fr["cameras"])
    pipe = api.GeometryPipeline(cams)
    r3 = pipe.triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    C = len(names)
    ts = T0 + np.arange(F)[:, None] * DT + np.arange(C)[None, :] * 1000
    msgs = []
    for f in range(F):
        body = wire.encode_personcovlist(r3["persons3d"][f, :r3["n_out"][f]], int(ts[f].max()), ts[f], np.full(C, 0.08, np.float32), seq=f)
        msgs.append((int(ts[f].max()), 1, "human_pose_estimation/persons3d_fused_pred", body))
    # a message in the wrong frame is refused (REP:140-143)
    msgs.append((int(ts[-1].max()) + DT, 1, "human_pose_estimation/persons3d_fused_pred",
                 wire.encode_personcovlist(r3["persons3d"][0, :1], int(ts[-1].max()) + DT, ts[-1], np.zeros(C, np.float32), frame_id="map")))
    rc, pubs, log = run_node(nodes["pose_reproj_ses3d_node"], tmp_path, {"~cameras": names}, tf, camera_info_msgs(cams, names) + msgs, "rep")
    assert rc == 0, log
    assert any("not given in" in l for l in log)
    assert len(pubs) == F * C                      # one list per camera per message, empty ones included (REP:233-234)
    want = pipe.reproject_batch(r3["persons3d"], r3["n_out"])
    n_rec = 0
    for i, (topic, body) in enumerate(pubs):
        f, c = divmod(i, C)
        assert topic == f"{names[c]}/skel_pred"
        m = wire.decode_person2dlist(body)
        assert m["stamp_ns"] == ts[f, c] and abs(m["fb_delay"] - 0.08) < 1e-7 and m["frame_id"] == names[c] + "_color_optical_frame"
        n = int(want["n_out"][f, c])
        assert len(m["persons"]) == n and m["persons"].tobytes() == want["persons2d"][f, c, :n].tobytes()
        n_rec += n
    assert n_rec > 0
    pipe.close()


@pytest.mark.gpu
def test_pose_prior_node_replay_equals_direct_calls(nodes, tmp_path):
    from smartedgesensor3dhumanpose_b200 import api
    T = 30
    fr = helpers.make_sequence_workload("ring8", 1, T, 3)
    pipe = api.GeometryPipeline(fr["cameras"])
    r3 = pipe.triangulate_batch(fr["persons"], fr["n_persons"], fr["h_max"])
    C = len(fr["cameras"])
    msgs = []
    fb = np.full(C, 0.09, np.float32)
    for f in range(T):
        stamp = int(fr["stamp_ns"][0, f])
        body = wire.encode_personcovlist(r3["persons3d"][f, :r3["n_out"][f]], stamp, np.full(C, stamp), fb, seq=f)
        msgs.append((stamp, 1, "human_pose_estimation/persons3d", body))
    rc, pubs, log = run_node(nodes["pose_prior_ses3d_node"], tmp_path, {"~vis_cov": False}, [], msgs, "pri")
    assert rc == 0, log
    fused = [wire.decode_personcovlist(b) for t, b in pubs if t == "human_pose_estimation/persons3d_fused"]
    pred = [wire.decode_personcovlist(b) for t, b in pubs if t == "human_pose_estimation/persons3d_fused_pred"]
    mk = [sc.decode_marker_array(b) for t, b in pubs if t == "human_pose_estimation/skeleton3d_fused"]
    assert len(fused) == T and len(pred) == T and len(mk) == T
    node = api.PosePrior(h_max=max(int(r3["n_out"].max()), 1))
    n_pub = 0
    for f in range(T):
        n = int(r3["n_out"][f])
        # the node sizes h_max per message; the tracker's results do not depend on the padding
        wf, wp, delay = node.skeleton_callback(r3["persons3d"][f, :n], int(fr["stamp_ns"][0, f]), fb)
        assert fused[f]["persons"].tobytes() == wf.tobytes() and pred[f]["persons"].tobytes() == wp.tobytes()
        assert np.allclose(fused[f]["fb_delay_per_cam"], delay) and len(fused[f]["fb_delay_per_cam"]) == C   # PRI:531
        assert len([m for m in mk[f] if m["ns"] == "fused_skeleton"]) == len(wf)
        n_pub += len(wf)
    assert n_pub > 0
    pipe.close()
