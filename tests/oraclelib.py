"""ctypes access to oracle/liboracle.so — the plain-C CPU restatement (checker only; tests/bench baseline)."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle.so")
capi = importlib.import_module("visual-inertial-odometry_b200").capi
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class OrcPrior(C.Structure):
    _fields_ = [("dim", C.c_int32), ("H", _dp), ("b", _dp), ("err_dim", C.c_int32), ("err", _dp), ("jt_inv", _dp)]


class OrcResult(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("linearizations", C.c_int32), ("trial_steps", C.c_int32),
                ("pcg_iterations", C.c_int64), ("chi2_initial", C.c_double), ("chi2_final", C.c_double),
                ("lambda_initial", C.c_double), ("lambda_final", C.c_double), ("ms_total", C.c_double),
                ("ms_hessian", C.c_double), ("chi2_trace", C.c_double * capi.TRACE_MAX),
                ("lambda_trace", C.c_double * capi.TRACE_MAX)]


_lib = None


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_solve_linear.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_double, C.c_int, _dp, _dp, _dp, C.POINTER(C.c_int64)]
        _lib.orc_linearize_bsr.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp, _dp, _dp, C.c_int64, C.c_int64]
        _lib.orc_linearize_sample.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _dp]
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _prior(scene):
    pr = OrcPrior()
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


def hessian(scene, flavour):
    g, keep = scene.to_c()
    pr, k2 = _prior(scene)
    n = scene.P + scene.inv_depth.shape[0] + 3 * scene.point_xyz.shape[0]
    H, b = np.zeros((n, n)), np.zeros(n)
    rc = lib().orc_make_hessian(C.byref(g), C.byref(pr), flavour, _d(H), _d(b))
    assert rc == 0, rc
    return H, b


def chi2(scene, flavour):
    g, keep = scene.to_c()
    pr, k2 = _prior(scene)
    out = C.c_double()
    rc = lib().orc_chi2(C.byref(g), C.byref(pr), flavour, C.byref(out))
    assert rc == 0, rc
    return out.value


def solve_linear(H, b, P, lam, solver, n_point=0):
    """n_point: number of VertexPointXYZ landmarks (3x3 blocks) at the end of the landmark range"""
    n = H.shape[0]
    M1 = n - P - 3 * n_point
    S, bS, dx = np.zeros((P, P)), np.zeros(P), np.zeros(n)
    it = C.c_int64()
    H = np.ascontiguousarray(H)
    b = np.ascontiguousarray(b)
    lib().orc_solve_linear_blocks.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _dp, _dp, _dp,
                                              C.POINTER(C.c_int64)]
    rc = lib().orc_solve_linear_blocks(_d(H), _d(b), P, M1, n_point, lam, solver, _d(S), _d(bS), _d(dx), C.byref(it))
    assert rc == 0, rc
    return S, bS, dx, it.value


def solve(scene, iterations, opts):
    g, keep = scene.to_c()
    pr, k2 = _prior(scene)
    pose = np.zeros_like(scene.pose)
    sb = np.zeros_like(scene.speedbias)
    invd = np.zeros_like(scene.inv_depth)
    pts = np.zeros_like(scene.point_xyz)
    bpo = np.zeros(max(scene.P, 1))
    epo = np.zeros(max(scene.P, 1))
    res = OrcResult()
    rc = lib().orc_solve_points(C.byref(g), C.byref(pr), iterations, C.byref(opts), _d(pose), _d(sb) if sb.size else None,
                                _d(invd), _d(pts), _d(bpo), _d(epo), C.byref(res))
    assert rc == 0, rc
    n = min(res.iterations, capi.TRACE_MAX)
    return dict(pose=pose, speedbias=sb, inv_depth=invd, point_xyz=pts, iterations=res.iterations, chi2_trace=np.array(res.chi2_trace[:n]),
                lambda_trace=np.array(res.lambda_trace[:n]), chi2_final=res.chi2_final, lambda_final=res.lambda_final,
                ms_total=res.ms_total, ms_hessian=res.ms_hessian, linearizations=res.linearizations,
                pcg_iterations=res.pcg_iterations, b_prior=bpo, err_prior=epo)


def linearize_bsr(scene, rowptr, col, lm_begin=0, lm_end=None):
    g, keep = scene.to_c()
    L = scene.inv_depth.shape[0]
    lm_end = L if lm_end is None else lm_end
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    val = np.zeros((col.shape[0], 6, 6))
    bS = np.zeros(scene.P)
    Hll, bl = np.zeros(L), np.zeros(L)
    rc = lib().orc_linearize_bsr(C.byref(g), rowptr.ctypes.data_as(_ip), col.ctypes.data_as(_ip), _d(val), _d(bS), _d(Hll),
                                 _d(bl), lm_begin, lm_end)
    assert rc == 0, rc
    return val, bS, Hll, bl


def linearize_sample(scene, lm_begin, lm_end):
    g, keep = scene.to_c()
    cs = C.c_double()
    rc = lib().orc_linearize_sample(C.byref(g), lm_begin, lm_end, C.byref(cs))
    assert rc == 0, rc
    return cs.value


def preintegrate(dt, acc, gyr, ba, bg, noise):
    """orc_preintegrate on one segment -> (sum_dt, dp, dq, dv, jac[225], cov[225])."""
    L = lib()
    dt, acc, gyr = (np.ascontiguousarray(x, np.float64) for x in (dt, acc, gyr))
    ba, bg, noise = (np.ascontiguousarray(x, np.float64) for x in (ba, bg, noise))
    sd = C.c_double()
    dp, dq, dv, jac, cov = np.zeros(3), np.zeros(4), np.zeros(3), np.zeros(225), np.zeros(225)
    L.orc_preintegrate.argtypes = [C.c_int32, _dp, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_double), _dp, _dp, _dp, _dp, _dp]
    rc = L.orc_preintegrate(dt.shape[0], _d(dt), _d(acc), _d(gyr), _d(ba), _d(bg), _d(noise), C.byref(sd), _d(dp), _d(dq), _d(dv),
                            _d(jac), _d(cov))
    assert rc == 0, rc
    return sd.value, dp, dq, dv, jac, cov


def marginalize(scene, marg_pose, marg_sb):
    """orc_marginalize -> dict(dim, H, b, err, jt_inv) like Problem.marginalize"""
    g, keep = scene.to_c()
    pr, k2 = _prior(scene)
    n = scene.P
    H, b, err, Jt = np.zeros((n, n)), np.zeros(n), np.zeros(n), np.zeros((n, n))
    dim = C.c_int32()
    L = lib()
    L.orc_marginalize.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), _dp, _dp, _dp, _dp]
    rc = L.orc_marginalize(C.byref(g), C.byref(pr), marg_pose, marg_sb, C.byref(dim), _d(H), _d(b), _d(err), _d(Jt))
    assert rc == 0, rc
    k = dim.value
    return dict(dim=k, H=H.ravel()[:k * k].reshape(k, k).copy(), b=b[:k].copy(), err=err[:k].copy(),
                jt_inv=Jt.ravel()[:k * k].reshape(k, k).copy())


def nullspace_hessian(scene):
    """numpy restatement of hessian_nullspace_test.cpp:95-140 (14-sliding-window/src): H = J^T J over pose(6) + XYZ(3)
    reprojection blocks of the `scenes.nullspace()` graph, in the reference's own parameterisation
    (jacobian_Ci = [J_uv Rcw | J_uv hat(Pc)], jacobian_Pj = J_uv Rcw, fx = fy = 1)."""
    qR = importlib.import_module("visual-inertial-odometry_b200").scenes._quat_R
    N, M = scene.pose.shape[0], scene.point_xyz.shape[0]
    H = np.zeros((6 * N + 3 * M, 6 * N + 3 * M))
    for m in range(M):
        Pw = scene.point_xyz[m]
        for n in range(N):
            Rcw = qR(scene.pose[n, 3:7]).T
            x, y, z = Rcw @ (Pw - scene.pose[n, :3])
            Juv = np.array([[1 / z, 0, -x / z ** 2], [0, 1 / z, -y / z ** 2]])
            JP = Juv @ Rcw
            JC = np.hstack([Juv @ Rcw, Juv @ np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])])
            a, b = slice(6 * n, 6 * n + 6), slice(6 * N + 3 * m, 6 * N + 3 * m + 3)
            H[a, a] += JC.T @ JC
            H[a, b] += JC.T @ JP
            H[b, a] += JP.T @ JC
            H[b, b] += JP.T @ JP
    return H


def nullspace_golden():
    """The 120 singular values printed by the unmodified reference binary (6 digits), tests/golden/."""
    path = os.path.join(ROOT, "tests", "golden", "hessian_nullspace_singular_values.txt")
    with open(path) as f:
        return np.array([float(ln.split(":")[1]) for ln in f if ":" in ln and ln.strip()[0].isdigit()])
