"""ctypes access to oracle/_ref/libref15.so / libref17.so — the UNMODIFIED reference backends.

Checker only.  Imported from tests/ (and tests/golden/make_golden.py); never from the package.
"""
import ctypes as C
import importlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
capi = importlib.import_module("visual-inertial-odometry_b200").capi
_dp = C.POINTER(C.c_double)


class RefPrior(C.Structure):
    _fields_ = [("dim", C.c_int32), ("H", _dp), ("b", _dp), ("err_dim", C.c_int32), ("err", _dp), ("jt_inv", _dp)]


class RefResult(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("chi2_trace", C.c_double * capi.TRACE_MAX),
                ("lambda_trace", C.c_double * capi.TRACE_MAX), ("chi2_final", C.c_double),
                ("lambda_final", C.c_double), ("ms_solve", C.c_double), ("ms_hessian", C.c_double)]


def available(ver):
    return os.path.exists(os.path.join(REF_DIR, f"libref{ver}.so"))


_libs = {}


def _lib(ver):
    if ver not in _libs:
        _libs[ver] = C.CDLL(os.path.join(REF_DIR, f"libref{ver}.so"))
    return _libs[ver]


def _d(a):
    return a.ctypes.data_as(_dp)


def _prior(scene):
    pr = RefPrior()
    keep = []
    if scene.prior is not None:
        H = np.ascontiguousarray(scene.prior["H"], np.float64)
        b = np.ascontiguousarray(scene.prior["b"], np.float64)
        pr.dim, pr.H, pr.b = b.shape[0], _d(H), _d(b)
        keep += [H, b]
        err = scene.prior.get("err")
        if err is not None and len(err):
            err = np.ascontiguousarray(err, np.float64)
            jt = np.ascontiguousarray(scene.prior["jt_inv"], np.float64)
            pr.err_dim, pr.err, pr.jt_inv = err.shape[0], _d(err), _d(jt)
            keep += [err, jt]
    return pr, keep


def _m(scene):
    """landmark dimension: inverse depths + 3 per VertexPointXYZ"""
    scene._norm()
    return scene.inv_depth.shape[0] + 3 * scene.point_xyz.shape[0]


def hessian(ver, scene):
    g, keep = scene.to_c()
    pr, k2 = _prior(scene)
    n = scene.P + _m(scene)
    H, b = np.zeros((n, n)), np.zeros(n)
    P, M = C.c_int32(), C.c_int32()
    rc = getattr(_lib(ver), f"ref{ver}_hessian")(C.byref(g), C.byref(pr), _d(H), _d(b), C.byref(P), C.byref(M))
    assert rc == 0, rc
    assert P.value == scene.P and M.value == _m(scene)
    return H, b


def init(ver, scene):
    g, keep = scene.to_c()
    pr, k2 = _prior(scene)
    chi, lam = C.c_double(), C.c_double()
    rc = getattr(_lib(ver), f"ref{ver}_init")(C.byref(g), C.byref(pr), C.byref(chi), C.byref(lam))
    assert rc == 0, rc
    return chi.value, lam.value


def step(ver, scene, lam):
    g, keep = scene.to_c()
    pr, k2 = _prior(scene)
    P = scene.P
    n = P + _m(scene)
    S, bS, dx = np.zeros((P, P)), np.zeros(P), np.zeros(n)
    rc = getattr(_lib(ver), f"ref{ver}_step")(C.byref(g), C.byref(pr), C.c_double(lam), _d(S), _d(bS), _d(dx))
    assert rc == 0, rc
    return S, bS, dx


def solve(ver, scene, iterations):
    g, keep = scene.to_c()
    pr, k2 = _prior(scene)
    pose = np.zeros_like(scene.pose)
    sb = np.zeros_like(scene.speedbias)
    invd = np.zeros_like(scene.inv_depth)
    pts = np.zeros_like(scene.point_xyz)
    res = RefResult()
    bpo = np.zeros(max(scene.P, 1))
    epo = np.zeros(max(scene.P, 1))
    rc = getattr(_lib(ver), f"ref{ver}_solve_points")(C.byref(g), C.byref(pr), iterations, _d(pose), _d(sb) if sb.size else None,
                                                      _d(invd), _d(pts), _d(bpo), _d(epo), C.byref(res))
    assert rc == 0, rc
    out = dict(pose=pose, speedbias=sb, inv_depth=invd, point_xyz=pts, iterations=res.iterations,
               chi2_trace=np.array(res.chi2_trace[:res.iterations]), lambda_trace=np.array(res.lambda_trace[:res.iterations]),
               chi2_final=res.chi2_final, lambda_final=res.lambda_final, ms_solve=res.ms_solve, ms_hessian=res.ms_hessian,
               b_prior=bpo, err_prior=epo)
    return out


def sparse_solve(scene, iterations, fixed_iterations=False):
    """Problem::Solve restated with block-sparse containers around the reference's own Edge / Vertex code
    (oracle/ref_sparse17.cpp): the at-scale CPU baseline and parity target.  v17, inverse-depth landmarks, fixed ext vertex."""
    L = _lib(17)
    g, keep = scene.to_c()
    pose = np.zeros_like(scene.pose)
    invd = np.zeros_like(scene.inv_depth)
    res = RefResult()
    timing = np.zeros(4)
    L.ref17_sparse_solve.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p,
                                     C.POINTER(C.c_double)]
    rc = L.ref17_sparse_solve(C.byref(g), iterations, int(bool(fixed_iterations)), _d(pose), _d(invd), C.byref(res), _d(timing))
    assert rc == 0, rc
    n = min(res.iterations, capi.TRACE_MAX)
    return dict(pose=pose, inv_depth=invd, iterations=res.iterations, chi2_trace=np.array(res.chi2_trace[:n]),
                lambda_trace=np.array(res.lambda_trace[:n]), chi2_final=res.chi2_final, lambda_final=res.lambda_final,
                t_linearize=timing[0], t_solve=timing[1], t_chi2=timing[2], t_total=timing[3])
