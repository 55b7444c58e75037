"""ctypes access to tests/libhost_emul.so: the device per-landmark bodies compiled for the host (unit-test aid)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "libhost_emul.so")
_dp = C.POINTER(C.c_double)
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def hessian(scene):
    g, keep = scene.to_c()
    n = scene.P + scene.inv_depth.shape[0]
    H, b = np.zeros((n, n)), np.zeros(n)
    rc = lib().emul_hessian(C.byref(g), _d(H), _d(b))
    assert rc == 0, rc
    return H, b


def schur(scene):
    g, keep = scene.to_c()
    P = scene.P
    S, bS = np.zeros((P, P)), np.zeros(P)
    rc = lib().emul_linearize(C.byref(g), 1, _d(S), None, _d(bS), None, None, None, None, None)
    assert rc == 0, rc
    return S, bS


def chi2(scene):
    g, keep = scene.to_c()
    out = C.c_double()
    rc = lib().emul_chi2(C.byref(g), C.byref(out))
    assert rc == 0, rc
    return out.value


def update_pose(pose, dx6, sign=1.0):
    pose = np.ascontiguousarray(pose, np.float64)
    dx6 = np.ascontiguousarray(dx6, np.float64)
    out = np.zeros_like(pose)
    lib().emul_update_pose.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp]
    rc = lib().emul_update_pose(pose.shape[0], _d(pose), _d(dx6), sign, _d(out))
    assert rc == 0
    return out
