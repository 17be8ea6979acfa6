"""Synthetic PersonCovList *sequences* for the pose_prior stage: temporally coherent people.

Test / bench input source, not part of the reference. People walk on smooth trajectories; every frame holds
their noisy 3-D skeletons in FUSION_BODY_PARTS slots exactly as skeleton_3d publishes them ("simple": 17 of the
21 slots, S3D:139-141; Neck/MidHip/Head/Belly stay empty), with SPD covariances, detector-like scores, joint and
person drop-outs, shuffled person order, 30 Hz stamps with jitter and per-camera feedback delays.
"""
import numpy as np

from .layouts import KP2FUSION_H36M, KP2FUSION_SIMPLE, POSE_SIMPLE, person_cov_dtype

# upright COCO-17 figure (same as csrc/synth.h::template_joint), x forward / y left / z up
TEMPLATE = np.array([
    [0.100, 0.000, 1.6400], [0.075, 0.030, 1.671225], [0.075, -0.030, 1.671225], [-0.005, 0.085, 1.651225],
    [-0.005, -0.085, 1.651225], [0.000, 0.175, 1.4744], [0.000, -0.175, 1.4744], [0.000, 0.225, 1.1989],
    [0.000, -0.225, 1.1989], [0.060, 0.225, 0.9562], [0.060, -0.225, 0.9562], [0.000, 0.135, 0.9760],
    [0.000, -0.135, 0.9760], [0.000, 0.135, 0.5260], [0.000, -0.135, 0.5260], [0.000, 0.135, 0.0800],
    [0.000, -0.135, 0.0800]])

# H36M-17 detector order (S3D:111-128): joints 1..4 are Head, Neck, Belly, Root instead of eyes / ears
TEMPLATE_H36M = TEMPLATE.copy()
TEMPLATE_H36M[1] = [0.060, 0.000, 1.7400]    # Head
TEMPLATE_H36M[2] = [0.000, 0.000, 1.4744]    # Neck (between the shoulders)
TEMPLATE_H36M[3] = [0.000, 0.000, 1.2190]    # Belly
TEMPLATE_H36M[4] = [0.000, 0.000, 0.9760]    # Root / MidHip (between the hips)


def synth_person_sequences(n_sequences, n_frames, n_people, seed=0, h_max=None, pose_method=POSE_SIMPLE, n_cams=16,
                           noise_m=0.015, joint_dropout=0.05, person_dropout=0.02, fps=30.0, jitter_s=0.002,
                           area=6.0, speed=1.0, shuffle=True, t0_s=1000.0):
    """Returns dict(persons [S][T][h_max], n_persons [S][T], stamp_ns [S][T], fb_delay [S][T][n_cams],
    gt_person [S][T][h_max] = generator person index of every emitted record)."""
    rng = np.random.default_rng(seed)
    S, T, P = n_sequences, n_frames, n_people
    h_max = h_max or max(8, P + 2)
    fus = np.array(KP2FUSION_SIMPLE if pose_method == POSE_SIMPLE else KP2FUSION_H36M)
    tpl = TEMPLATE if pose_method == POSE_SIMPLE else TEMPLATE_H36M
    # smooth trajectories: heading random walk, constant speed
    pos0 = rng.uniform(-area / 2, area / 2, (S, P, 2))
    head = rng.uniform(0, 2 * np.pi, (S, P, 1)) + np.cumsum(rng.normal(0, 0.03, (S, P, T)), axis=-1)
    v = speed * rng.uniform(0.3, 1.2, (S, P, 1))
    dt = 1.0 / fps
    step = np.stack([np.cos(head), np.sin(head)], -1) * (v * dt)[..., None]          # [S][P][T][2]
    pos = pos0[:, :, None, :] + np.cumsum(step, axis=2)
    c, s = np.cos(head), np.sin(head)                                                 # body faces the heading
    X = np.empty((S, P, T, 17, 3))
    X[..., 0] = pos[..., None, 0] + c[..., None] * tpl[:, 0] - s[..., None] * tpl[:, 1]
    X[..., 1] = pos[..., None, 1] + s[..., None] * tpl[:, 0] + c[..., None] * tpl[:, 1]
    X[..., 2] = tpl[:, 2]
    # arm / leg swing so that bones move a little
    phase = rng.uniform(0, 2 * np.pi, (S, P, 1)) + 2 * np.pi * 1.8 * dt * np.arange(T)
    swing = 0.10 * np.sin(phase)
    for k, sign in ((9, 1), (10, -1), (15, -1), (16, 1), (13, -0.5), (14, 0.5), (7, 0.5), (8, -0.5)):
        X[..., k, 0] += sign * swing * c
        X[..., k, 1] += sign * swing * s
    X += rng.normal(0, noise_m, X.shape)
    # covariances: SPD, roughly noise_m^2 with correlated axes
    A = rng.normal(0, 1, (S, P, T, 17, 3, 3)) * 0.35 + np.eye(3)
    cov = (A @ np.swapaxes(A, -1, -2)) * noise_m ** 2
    score = rng.uniform(0.5, 1.0, (S, P, T, 17)).astype(np.float32)
    drop = rng.random((S, P, T, 17)) < joint_dropout
    present = rng.random((S, P, T)) >= person_dropout

    # all records [S][T][P], then per message: shuffle, drop absent people, pad to h_max
    full = np.zeros((S, T, P), person_cov_dtype)
    keep = np.transpose(~drop, (0, 2, 1, 3))                                  # [S][T][P][17]
    Xt = np.transpose(X, (0, 2, 1, 3, 4))
    ct = np.transpose(cov, (0, 2, 1, 3, 4, 5))
    sc = np.transpose(score, (0, 2, 1, 3)) * keep
    kp = full["keypoints"]
    for name, axis in (("x", 0), ("y", 1), ("z", 2)):
        v = np.zeros((S, T, P, 21))
        v[..., fus] = Xt[..., axis] * keep
        kp[name] = v
    v = np.zeros((S, T, P, 21), np.float32)
    v[..., fus] = sc
    kp["score"] = v
    c6 = np.stack([ct[..., 0, 0], ct[..., 0, 1], ct[..., 0, 2], ct[..., 1, 1], ct[..., 1, 2], ct[..., 2, 2]], -1)
    v = np.zeros((S, T, P, 21, 6))
    v[..., fus, :] = c6 * keep[..., None]
    kp["cov"] = v
    full["keypoints"] = kp
    full["score"] = (sc.sum(-1) / np.maximum(keep.sum(-1), 1)).astype(np.float32)
    rank = np.argsort(rng.random((S, T, P)), axis=-1) if shuffle else np.broadcast_to(np.arange(P), (S, T, P))
    pres = np.transpose(present, (0, 2, 1))                                   # [S][T][P]
    key = ~np.take_along_axis(pres, rank, -1)
    # position list: people in `rank` order, absent ones last
    pos = np.take_along_axis(rank, np.argsort(key, axis=-1, kind="stable"), -1)
    n_persons = np.minimum(pres.sum(-1), h_max).astype(np.int32)
    take = min(P, h_max)
    persons = np.zeros((S, T, h_max), person_cov_dtype)
    persons[:, :, :take] = np.take_along_axis(full, pos[:, :, :take], axis=2)
    live = np.arange(h_max)[None, None, :] < n_persons[:, :, None]
    persons[~live] = np.zeros((), person_cov_dtype)
    gt = np.full((S, T, h_max), -1, np.int32)
    gt[:, :, :take] = pos[:, :, :take]
    gt[~live] = -1
    persons["id"] = np.where(live, np.arange(h_max)[None, None, :], 0)
    t = t0_s + np.arange(T) * dt + rng.normal(0, jitter_s, (S, T))
    t = np.maximum.accumulate(t, axis=1)
    stamp_ns = np.round(t * 1e9).astype(np.int64)
    fb = rng.uniform(0.05, 0.15, (S, T, n_cams)).astype(np.float32)
    fb[rng.random((S, T, n_cams)) < 0.2] = -1.0
    return dict(persons=persons, n_persons=n_persons, stamp_ns=stamp_ns, fb_delay=fb, gt_person=gt, h_max=h_max,
                n_cams=n_cams)
