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


CONFIGS = {
    # name: (factory, kwargs)
    "config1_monoba_20x300": (monoba, dict(pose_nums=20, feature_nums=300)),
    "config4_ba_1k_100k": (ring, dict(n_cam=1000, n_landmark=100000, k_obs=11, seed=4)),
    "config5_ba_10k_1m": (ring, dict(n_cam=10000, n_landmark=1000000, k_obs=11, seed=5)),
}
