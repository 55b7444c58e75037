"""Synthetic scenes for BASELINE.json's configs (wrappers over csrc/scene_gen.cc, libvio_scenes.so)."""
import ctypes as C
import os

import numpy as np

from . import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvio_scenes.so")
_lib = None


def _L():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run __graft_entry__.build()")
        _lib = C.CDLL(LIB_PATH)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _alloc(n_pose, n_lm, n_e):
    s = capi.Scene()
    s.pose = np.zeros((n_pose, 7))
    s.pose_fixed = np.zeros(n_pose, np.uint8)
    s.pose_gt = np.zeros((n_pose, 7))
    s.inv_depth = np.zeros(n_lm)
    s.inv_depth_gt = np.zeros(n_lm)
    s.rp_landmark = np.zeros(n_e, np.int32)
    s.rp_pose_i = np.zeros(n_e, np.int32)
    s.rp_pose_j = np.zeros(n_e, np.int32)
    s.rp_pts_i = np.zeros((n_e, 3))
    s.rp_pts_j = np.zeros((n_e, 2))
    s.sp_pose = np.zeros(2, np.int32)
    s.sp_p = np.zeros((2, 3))
    s.sp_q = np.zeros((2, 4))
    s.sp_info = np.zeros((2, 36))
    return s


def _fill_args(s):
    d, i, b = C.c_double, C.c_int32, C.c_uint8
    return [_p(s.pose, d), _p(s.pose_fixed, b), _p(s.pose_gt, d), _p(s.inv_depth, d), _p(s.inv_depth_gt, d),
            _p(s.rp_landmark, i), _p(s.rp_pose_i, i), _p(s.rp_pose_j, i), _p(s.rp_pts_i, d), _p(s.rp_pts_j, d),
            _p(s.sp_pose, i), _p(s.sp_p, d), _p(s.sp_q, d), _p(s.sp_info, d)]


def monoba(pose_nums=20, feature_nums=300, with_ext=False, prior_weight=1e4):
    """BASELINE config 1: the reference's TestMonoBA scene (A15/app/TestMonoBA.cpp), draw for draw.

    with_ext=True adds the fixed identity extrinsic VertexPose (pose 0) the v17 4-vertex edge needs.
    """
    L = _L()
    n_pose, n_lm, n_e = C.c_int32(), C.c_int32(), C.c_int64()
    rc = L.vio_scene_monoba_sizes(pose_nums, feature_nums, int(with_ext), C.byref(n_pose), C.byref(n_lm), C.byref(n_e))
    if rc:
        raise ValueError("bad monoba sizes")
    s = _alloc(n_pose.value, n_lm.value, n_e.value)
    L.vio_scene_monoba_fill(pose_nums, feature_nums, int(with_ext), C.c_double(prior_weight), *_fill_args(s))
    s.ext_pose = 0 if with_ext else -1
    return s


def ring(n_cam=1000, n_landmark=100000, k_obs=11, with_ext=False, seed=4, prior_weight=1e4):
    """BASELINE configs 4/5: cameras on a circle, K observations per landmark."""
    L = _L()
    n_pose, n_e = C.c_int32(), C.c_int64()
    rc = L.vio_scene_ring_sizes(n_cam, n_landmark, k_obs, int(with_ext), C.byref(n_pose), C.byref(n_e))
    if rc:
        raise ValueError("bad ring sizes")
    s = _alloc(n_pose.value, n_landmark, n_e.value)
    L.vio_scene_ring_fill(n_cam, n_landmark, k_obs, int(with_ext), C.c_uint64(seed), C.c_double(prior_weight),
                          *_fill_args(s))
    s.ext_pose = 0 if with_ext else -1
    return s


def nullspace(draw_order=1):
    """The reference's hessian_nullspace_test scene (A14 = 14-sliding-window/src/hessian_nullspace_test.cpp:45-93) as a
    graph of VertexPose + VertexPointXYZ + EdgeReprojectionXYZ: 10 cameras, 20 world points, every camera sees every
    point, exact observations (zero residual), identity extrinsics, no priors.  J^T J of this graph has the singular
    values the reference publishes (A14/README.md:125-149), nullspace dimension 7."""
    L = _L()
    s = capi.Scene()
    s.pose = np.zeros((10, 7))
    pts = np.zeros((20, 3))
    L.vio_scene_nullspace_fill(int(draw_order), _p(s.pose, C.c_double), _p(pts, C.c_double))
    s.point_xyz = pts
    rx_point, rx_pose, rx_obs = [], [], []
    for m in range(20):
        for n in range(10):
            pc = _quat_R(s.pose[n, 3:7]).T @ (pts[m] - s.pose[n, :3])
            rx_point.append(m); rx_pose.append(n); rx_obs.append(pc[:2] / pc[2])
    s.rx_point = np.asarray(rx_point, np.int32)
    s.rx_pose = np.asarray(rx_pose, np.int32)
    s.rx_obs = np.asarray(rx_obs, np.float64)
    return s


CONFIGS = {
    # name: (factory, kwargs)
    "config1_monoba_20x300": (monoba, dict(pose_nums=20, feature_nums=300)),
    "config4_ba_1k_100k": (ring, dict(n_cam=1000, n_landmark=100000, k_obs=11, seed=4)),
    "config5_ba_10k_1m": (ring, dict(n_cam=10000, n_landmark=1000000, k_obs=11, seed=5)),
}


def _quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def to_xyz(scene, mask=None, noise=0.0, seed=0):
    """Re-parameterise inverse-depth landmarks as VertexPointXYZ world points observed through EdgeReprojectionXYZ
    (A15/backend/edge_reprojection.cc:113-163): every observation of a converted landmark - the host's included -
    becomes a 2-vertex [X_w, T_i] factor.  mask selects the landmarks to convert (default all); the others stay
    inverse-depth landmarks, so mixed graphs can be built.  noise perturbs the points (metres, Gaussian)."""
    from .capi import Scene
    scene._norm()
    L = scene.inv_depth.shape[0]
    mask = np.ones(L, bool) if mask is None else np.asarray(mask, bool)
    d = scene.export()
    if scene.ext_pose >= 0:
        ex = scene.pose[scene.ext_pose]
        tic, Ric = ex[:3], _quat_R(ex[3:7])
    else:
        tic, Ric = np.asarray(scene.t_ic, float), _quat_R(scene.q_ic)
    lm, pi, pj = scene.rp_landmark, scene.rp_pose_i, scene.rp_pose_j
    conv = mask[lm]
    new_id = -np.ones(L, np.int64)
    new_id[mask] = np.arange(mask.sum())
    keep_id = -np.ones(L, np.int64)
    keep_id[~mask] = np.arange((~mask).sum())
    pts = np.zeros((int(mask.sum()), 3))
    rx_point, rx_pose, rx_obs = [], [], []
    seen = set()
    for e in np.nonzero(conv)[0]:
        l, h = int(lm[e]), int(pi[e])
        if l not in seen:
            seen.add(l)
            pc = scene.rp_pts_i[e] / scene.inv_depth[l]
            pb = Ric @ pc + tic
            pts[new_id[l]] = _quat_R(scene.pose[h, 3:7]) @ pb + scene.pose[h, :3]
            rx_point.append(new_id[l]); rx_pose.append(h); rx_obs.append(scene.rp_pts_i[e, :2] / scene.rp_pts_i[e, 2])
        rx_point.append(new_id[l]); rx_pose.append(int(pj[e])); rx_obs.append(scene.rp_pts_j[e, :2])
    if noise > 0:
        pts += np.random.default_rng(seed).normal(0, noise, pts.shape)
    k = ~conv
    d["inv_depth"] = scene.inv_depth[~mask]
    d["rp_landmark"] = keep_id[lm[k]].astype(np.int32)
    for key in ("rp_pose_i", "rp_pose_j", "rp_pts_i", "rp_pts_j"):
        d[key] = getattr(scene, key)[k]
    d["point_xyz"] = pts
    d["rx_point"] = np.asarray(rx_point, np.int32)
    d["rx_pose"] = np.asarray(rx_pose, np.int32)
    d["rx_obs"] = np.asarray(rx_obs, np.float64).reshape(-1, 2)
    d.pop("inv_depth_gt", None)
    return Scene.from_dict(d)
