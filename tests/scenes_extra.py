"""BASELINE config 2 — VINS-style sliding window: fixed extrinsic pose + 11 x (pose, speed-bias)
=> P = 171, 10 EdgeImu, ~1000 inverse-depth features with Cauchy(1) reprojection edges, and a 156-dim
marginalisation prior extended by 15.

The IMU pre-integration constants and the prior are produced by the UNMODIFIED reference
(IntegrationBase::push_back and Problem::Marginalize through oracle/_ref/libref17.so), so this generator only runs
where oracle/_ref is built; the resulting scene is committed as tests/golden/window_v17_scene.npz and loaded from
there on the GPU box.  Motion model: /root/reference/.../17-vins-initialization/simulator/src/imu.cpp:76-117,
extrinsics .../simulator/src/param.cpp:11-16, noise .../vins-mono/config/vio_simulation.yaml:60-63,79.
"""
import ctypes as C
import importlib

import numpy as np

from tests import refshim

capi = importlib.import_module("visual-inertial-odometry_b200").capi
_dp = C.POINTER(C.c_double)
NOISE = np.array([0.2687, 7.07e-6, 0.2121, 7.07e-7])  # ACC_N ACC_W GYR_N GYR_W
G = np.array([0.0, 0.0, 9.81])
R_BC = np.array([[0.0, 0, -1], [-1, 0, 0], [0, 1, 0]])
T_BC = np.array([0.05, 0.04, 0.03])


def _d(a):
    return a.ctypes.data_as(_dp)


def euler2R(e):
    r, p, y = e
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, sy * sr + cy * cr * sp],
                     [sy * cp, cy * cr + sy * sr * sp, sp * sy * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def motion(t):
    ex, ey, z, K1, K = 15.0, 20.0, 1.0, 10.0, np.pi / 10
    pos = np.array([ex * np.cos(K * t) + 5, ey * np.sin(K * t) + 5, z * np.sin(K1 * K * t) + 5])
    dp = np.array([-K * ex * np.sin(K * t), K * ey * np.cos(K * t), z * K1 * K * np.cos(K1 * K * t)])
    ddp = np.array([-K * K * ex * np.cos(K * t), -K * K * ey * np.sin(K * t), -z * K1 * K1 * K * K * np.sin(K1 * K * t)])
    eul = np.array([0.1 * np.cos(t), 0.2 * np.sin(t), K * t])
    deul = np.array([-0.1 * np.sin(t), 0.2 * np.cos(t), K])
    Rwb = euler2R(eul)
    cr, sr, cp, sp = np.cos(eul[0]), np.sin(eul[0]), np.cos(eul[1]), np.sin(eul[1])
    E = np.array([[1, 0, -sp], [0, cr, sr * cp], [0, -sr, cr * cp]])
    gyro = E @ deul
    acc = Rwb.T @ (ddp - np.array([0, 0, -9.81]))
    return pos, Rwb, dp, gyro, acc


def R2q(R):
    """xyzw, trace-based conversion"""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        return np.array([(R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s, w])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    q = np.zeros(4)
    q[i] = 0.5 * s
    s = 0.5 / s
    q[3] = (R[k, j] - R[j, k]) * s
    q[j] = (R[j, i] + R[i, j]) * s
    q[k] = (R[k, i] + R[i, k]) * s
    return q


def expm_so3(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def preintegrate(dt, acc, gyr):
    L = C.CDLL(refshim.os.path.join(refshim.REF_DIR, "libref17.so"))
    n = acc.shape[0]
    dts = np.full(n, dt)
    z3 = np.zeros(3)
    sum_dt = C.c_double()
    dp, dq, dv = np.zeros(3), np.zeros(4), np.zeros(3)
    jac, cov = np.zeros((15, 15)), np.zeros((15, 15))
    acc, gyr = np.ascontiguousarray(acc), np.ascontiguousarray(gyr)
    rc = L.ref17_preintegrate(n, _d(dts), _d(acc), _d(gyr), _d(z3), _d(z3), _d(NOISE), C.byref(sum_dt), _d(dp), _d(dq), _d(dv),
                              _d(jac), _d(cov))
    assert rc == 0
    return sum_dt.value, dp, dq, dv, jac, cov


def _base_scene(frames, states, imu_edges, feats, rng):
    """frames: list of global keyframe ids in this window; builds ext + (pose, sb) per frame."""
    s = capi.Scene()
    nf = len(frames)
    pose = np.zeros((nf + 1, 7))
    pose[0, :3] = T_BC
    pose[0, 3:] = R2q(R_BC)
    sb = np.zeros((nf, 9))
    for k, f in enumerate(frames):
        pose[k + 1] = states["pose"][f]
        sb[k] = states["sb"][f]
    s.pose, s.speedbias = pose, sb
    s.pose_fixed = np.zeros(nf + 1, np.uint8)
    s.pose_fixed[0] = 1
    s.ext_pose = 0
    order = [0]
    for k in range(nf):
        order += [k + 1, ~k]
    s.pclass_order = np.array(order, np.int32)
    s.gravity = G.copy()
    idx = {f: k for k, f in enumerate(frames)}
    m = dict(pose_i=[], sb_i=[], pose_j=[], sb_j=[], sum_dt=[], delta_p=[], delta_q=[], delta_v=[], lin_ba=[], lin_bg=[],
             jacobian=[], covariance=[])
    for (fa, fb), pre in imu_edges.items():
        if fa in idx and fb in idx:
            m["pose_i"].append(idx[fa] + 1); m["sb_i"].append(idx[fa]); m["pose_j"].append(idx[fb] + 1); m["sb_j"].append(idx[fb])
            m["sum_dt"].append(pre[0]); m["delta_p"].append(pre[1]); m["delta_q"].append(pre[2]); m["delta_v"].append(pre[3])
            m["lin_ba"].append(np.zeros(3)); m["lin_bg"].append(np.zeros(3)); m["jacobian"].append(pre[4].ravel())
            m["covariance"].append(pre[5].ravel())
    s.imu = {k: np.array(v) for k, v in m.items()}
    lm, pi, pj, pti, ptj, invd = [], [], [], [], [], []
    for l, (host, obs, lam0) in enumerate(feats):
        invd.append(lam0)
        for f, xy in obs[1:]:
            lm.append(l); pi.append(idx[host] + 1); pj.append(idx[f] + 1)
            pti.append([obs[0][1][0], obs[0][1][1], 1.0]); ptj.append(xy)
    s.inv_depth = np.array(invd)
    s.rp_landmark, s.rp_pose_i, s.rp_pose_j = np.array(lm, np.int32), np.array(pi, np.int32), np.array(pj, np.int32)
    s.rp_pts_i, s.rp_pts_j = np.array(pti).reshape(-1, 3), np.array(ptj).reshape(-1, 2)
    s.rp_info = (460.0 / 1.5) ** 2
    s.rp_loss, s.rp_loss_delta = capi.LOSS_CAUCHY, 1.0
    s.storage = capi.STORAGE_DENSE
    return s


def _make_features(n, hosts, lens, frames_gt, rng):
    """n features: host frame, observing frames host..host+len-1, point 5-15 m in front of the host camera."""
    feats = []
    for h, ln in zip(hosts, lens):
        Rwb, twb = frames_gt[h]
        Rwc, twc = Rwb @ R_BC, Rwb @ T_BC + twb
        depth = rng.uniform(5, 15)
        pc = np.array([rng.uniform(-0.4, 0.4) * depth, rng.uniform(-0.4, 0.4) * depth, depth])
        pw = Rwc @ pc + twc
        obs = []
        for f in range(h, h + ln):
            Rb, tb = frames_gt[f]
            pcf = (Rb @ R_BC).T @ (pw - (Rb @ T_BC + tb))
            if pcf[2] < 0.5:
                break
            obs.append((f, pcf[:2] / pcf[2] + rng.normal(0, 1.0 / 460.0, 2)))
        if len(obs) >= 2:
            feats.append((h, obs, (1.0 / depth) * (1.0 + rng.normal(0, 0.1))))
    return feats


def marginalize_ref(A, marg_pose=1, marg_sb=0):
    """Problem::Marginalize of the unmodified v17 backend on scene A -> dict(H, b, err, jt_inv) of dimension P - 15."""
    Lr = C.CDLL(refshim.os.path.join(refshim.REF_DIR, "libref17.so"))
    gA, keep = A.to_c()
    pr0, k2 = refshim._prior(A)
    P = A.P
    dim = C.c_int32()
    Hm, bm, em, jm = np.zeros((P, P)), np.zeros(P), np.zeros(P), np.zeros((P, P))
    rc = Lr.ref17_marginalize(C.byref(gA), C.byref(pr0), marg_pose, marg_sb, P, C.byref(dim), _d(Hm), _d(bm), _d(em), _d(jm))
    assert rc == 0, rc
    n = dim.value
    return dict(H=Hm.ravel()[:n * n].reshape(n, n).copy(), b=bm[:n].copy(), err=em[:n].copy(),
                jt_inv=jm.ravel()[:n * n].reshape(n, n).copy())


def window_scene(seed=2, n_feat=1000, return_marg_window=False):
    if not refshim.available(17):
        raise RuntimeError("window_scene needs oracle/_ref/libref17.so (IntegrationBase + Marginalize of the reference)")
    rng = np.random.default_rng(seed)
    dt, nkf, t0, kf_dt = 0.005, 12, 1.0, 0.2
    frames_gt, states = {}, {"pose": {}, "sb": {}}
    ba0, bg0 = rng.normal(0, 0.01, 3), rng.normal(0, 0.001, 3)  # slowly varying biases (random walk ACC_W, GYR_W)
    for k in range(nkf):
        pos, Rwb, vel, _, _ = motion(t0 + kf_dt * k)
        frames_gt[k] = (Rwb, pos)
        Rn = Rwb @ expm_so3(rng.normal(0, 0.01, 3))
        states["pose"][k] = np.concatenate([pos + rng.normal(0, 0.03, 3), R2q(Rn)])
        states["sb"][k] = np.concatenate([vel + rng.normal(0, 0.05, 3), ba0 + rng.normal(0, 3e-6, 3), bg0 + rng.normal(0, 3e-7, 3)])
    imu_edges = {}
    for k in range(nkf - 1):
        ts = t0 + kf_dt * k + dt * np.arange(int(round(kf_dt / dt)) + 1)
        acc, gyr = np.zeros((len(ts), 3)), np.zeros((len(ts), 3))
        for i, t in enumerate(ts):
            _, _, _, g_, a_ = motion(t)
            gyr[i] = g_ + rng.normal(0, 0.015 / np.sqrt(dt), 3)
            acc[i] = a_ + rng.normal(0, 0.019 / np.sqrt(dt), 3)
        imu_edges[(k, k + 1)] = preintegrate(dt, acc, gyr)
    # window A: frames 0..10, IMU edge 0->1 only, landmarks hosted in frame 0 (MargOldFrame, estimator.cpp:693-829)
    na = 120
    featsA = _make_features(na, [0] * na, rng.integers(2, 8, na), frames_gt, rng)
    A = _base_scene(list(range(11)), states, {(0, 1): imu_edges[(0, 1)]}, featsA, rng)
    Lr = C.CDLL(refshim.os.path.join(refshim.REF_DIR, "libref17.so"))
    gA, keep = A.to_c()
    pr0 = refshim.RefPrior()
    dim = C.c_int32()
    Hm, bm, em, jm = np.zeros((171, 171)), np.zeros(171), np.zeros(171), np.zeros((171, 171))
    rc = Lr.ref17_marginalize(C.byref(gA), C.byref(pr0), 1, 0, 171, C.byref(dim), _d(Hm), _d(bm), _d(em), _d(jm))
    assert rc == 0 and dim.value == 156, (rc, dim.value)
    n = dim.value
    Hp = Hm.ravel()[:n * n].reshape(n, n)
    bp, ep = bm[:n], em[:n]
    Jp = jm.ravel()[:n * n].reshape(n, n)
    # window B: frames 1..11 ; prior extended by 15 (ExtendHessiansPriorSize(15), estimator.cpp:1029-1033)
    hosts = rng.integers(0, 8, n_feat)
    lens = np.array([rng.integers(2, 11 - h + 1) for h in hosts])
    featsB = _make_features(n_feat, hosts + 1, lens, frames_gt, rng)
    B = _base_scene(list(range(1, 12)), states, imu_edges, featsB, rng)
    P = B.P
    assert P == 171
    H = np.zeros((P, P)); H[:n, :n] = Hp
    b = np.zeros(P); b[:n] = bp
    B.prior = dict(H=H, b=b, err=ep.copy(), jt_inv=Jp.copy())
    if return_marg_window:
        return B, A
    return B


def xyz_scene(kind):
    """Deterministic VertexPointXYZ / EdgeReprojectionXYZ scenes behind tests/golden/*xyz*.npz and mixed_*.npz: the
    TestMonoBA scene with (some of) its landmarks re-parameterised as world points (scenes.to_xyz)."""
    vio = importlib.import_module("visual-inertial-odometry_b200")
    sc = vio.scenes
    if kind == "xyz_v15":
        return sc.to_xyz(sc.monoba(6, 40), noise=0.01, seed=1)
    if kind == "xyz_v17_cauchy":
        s = sc.to_xyz(sc.monoba(6, 40, with_ext=True), noise=0.01, seed=1)
        s.rp_loss, s.rp_loss_delta, s.rp_info = capi.LOSS_CAUCHY, 1.0, 100.0
        return s
    if kind == "mixed_v17":
        m = np.zeros(40, bool)
        m[::2] = True
        s = sc.to_xyz(sc.monoba(6, 40, with_ext=True), mask=m, noise=0.01, seed=2)
        s.rp_loss, s.rp_loss_delta, s.rp_info = capi.LOSS_CAUCHY, 1.0, 100.0
        return s
    if kind == "xyz_v17_solve":
        return sc.to_xyz(sc.monoba(20, 300, with_ext=True), noise=0.02, seed=3)
    if kind == "mixed_v17_solve":
        m = np.zeros(300, bool)
        m[100:250] = True
        return sc.to_xyz(sc.monoba(20, 300, with_ext=True), mask=m, noise=0.02, seed=4)
    if kind == "xyz_v15_solve":
        return sc.to_xyz(sc.monoba(20, 300), noise=0.02, seed=5)
    raise ValueError(kind)


FIXED_LM = (3, 17, 18)
FIXED_PT = (0, 5, 39)


def fixed_scene(kind):
    """Scenes with FIXED landmark-class vertices behind tests/golden/fixed*_6x40_v17_lin.npz (make_golden_fixedlm.py)."""
    vio = importlib.import_module("visual-inertial-odometry_b200")
    sc = vio.scenes
    f = np.zeros(40, np.uint8)
    if kind == "lm":
        s = sc.monoba(6, 40, with_ext=True)
        f[list(FIXED_LM)] = 1
        s.landmark_fixed = f
        return s
    if kind == "pt":
        s = sc.to_xyz(sc.monoba(6, 40, with_ext=True), noise=0.01, seed=1)
        f[list(FIXED_PT)] = 1
        s.point_fixed = f
        return s
    raise ValueError(kind)


def extfree_scene(n_pose=6, n_feat=40):
    """TestMonoBA scene in the v17 4-vertex form with the extrinsic VertexPose NOT fixed (ESTIMATE_EXTRINSIC=1): every
    EdgeReprojection contributes its 4th Jacobian (A17/src/backend/edge_reprojection.cc:97-103)."""
    vio = importlib.import_module("visual-inertial-odometry_b200")
    s = vio.scenes.monoba(n_pose, n_feat, with_ext=True)
    s.pose_fixed = np.zeros(s.pose.shape[0], np.uint8)
    q = np.array([0.01, -0.02, 0.015, 1.0])
    s.pose[0, :3] = [0.05, -0.02, 0.01]
    s.pose[0, 3:] = q / np.linalg.norm(q)
    s.rp_loss, s.rp_loss_delta, s.rp_info = capi.LOSS_CAUCHY, 1.0, 100.0
    return s
