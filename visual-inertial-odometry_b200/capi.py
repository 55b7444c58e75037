"""ctypes binding of include/vio_b200.h — the same C-ABI the C++ `myslam::backend::Problem` mirror uses.

There is no Python or CPU implementation behind this module: every call goes to libvio_b200.so
(CUDA, sm_100a).  If the library is missing or no GPU is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvio_b200.so")

VIO_OK = 0
VIO_ERR_INVALID, VIO_ERR_CUDA, VIO_ERR_UNSUPPORTED, VIO_ERR_EMPTY, VIO_ERR_NO_DEVICE, VIO_ERR_STATE = 1, 2, 3, 4, 5, 6
ERR_NAMES = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "UNSUPPORTED", 4: "EMPTY", 5: "NO_DEVICE", 6: "STATE"}
LM_V15, LM_V17 = 0, 1
SOLVER_AUTO, SOLVER_DENSE_CHOL, SOLVER_REF_PCG, SOLVER_BLOCK_PCG, SOLVER_BLOCK_PCG_2L, SOLVER_BLOCK_CHOL = 0, 1, 2, 3, 4, 5
SOLVER_BCR = 6
LOSS_TRIVIAL, LOSS_HUBER, LOSS_CAUCHY, LOSS_TUKEY = 0, 1, 2, 3
STORAGE_AUTO, STORAGE_DENSE, STORAGE_BSR = 0, 1, 2
TRACE_MAX = 256

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


class VioGraph(C.Structure):
    _fields_ = [
        ("n_pose", C.c_int32), ("pose", _dp), ("pose_fixed", _bp),
        ("n_speedbias", C.c_int32), ("speedbias", _dp), ("speedbias_fixed", _bp),
        ("pclass_order", _ip),
        ("n_landmark", C.c_int32), ("inv_depth", _dp),
        ("n_reproj", C.c_int64), ("rp_landmark", _ip), ("rp_pose_i", _ip), ("rp_pose_j", _ip),
        ("rp_pts_i", _dp), ("rp_pts_j", _dp), ("rp_info", C.c_double), ("rp_loss", C.c_int32),
        ("rp_loss_delta", C.c_double), ("ext_pose", C.c_int32), ("q_ic", C.c_double * 4), ("t_ic", C.c_double * 3),
        ("n_se3prior", C.c_int32), ("sp_pose", _ip), ("sp_p", _dp), ("sp_q", _dp), ("sp_info", _dp),
        ("n_imu", C.c_int32), ("imu_pose_i", _ip), ("imu_sb_i", _ip), ("imu_pose_j", _ip), ("imu_sb_j", _ip),
        ("imu_sum_dt", _dp), ("imu_delta_p", _dp), ("imu_delta_q", _dp), ("imu_delta_v", _dp),
        ("imu_lin_ba", _dp), ("imu_lin_bg", _dp), ("imu_jacobian", _dp), ("imu_covariance", _dp),
        ("gravity", C.c_double * 3),
        ("storage", C.c_int32),
        ("n_point", C.c_int32), ("reserved_xyz", C.c_int32), ("point_xyz", _dp),
        ("n_reproj_xyz", C.c_int64), ("rx_point", _ip), ("rx_pose", _ip), ("rx_obs", _dp),
        ("landmark_fixed", _bp), ("point_fixed", _bp),
    ]


class VioLmOpts(C.Structure):
    _fields_ = [
        ("flavour", C.c_int32), ("solver", C.c_int32), ("verbose", C.c_int32), ("pcg_max_iter", C.c_int32),
        ("pcg_tol", C.c_double), ("fixed_iterations", C.c_int32), ("warm_start", C.c_int32),
    ]


class VioStats(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("linearizations", C.c_int32), ("trial_steps", C.c_int32),
        ("accepted_steps", C.c_int32), ("pcg_iterations", C.c_int64),
        ("chi2_initial", C.c_double), ("chi2_final", C.c_double),
        ("lambda_initial", C.c_double), ("lambda_final", C.c_double),
        ("ms_total", C.c_double), ("ms_linearize", C.c_double), ("ms_reduced_solve", C.c_double),
        ("ms_backsub_update", C.c_double), ("ms_chi2", C.c_double),
        ("n_trace", C.c_int32), ("solver_used", C.c_int32),
        ("chi2_trace", C.c_double * TRACE_MAX), ("lambda_trace", C.c_double * TRACE_MAX),
    ]


class VioDims(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("M", C.c_int32), ("n_pose_blocks", C.c_int32), ("storage", C.c_int32),
        ("nnz_blocks", C.c_int64), ("n_reproj", C.c_int64), ("n_groups", C.c_int32), ("reserved", C.c_int32),
    ]


class VioBatchItem(C.Structure):
    _fields_ = [
        ("graph", C.POINTER(VioGraph)), ("prior_dim", C.c_int32), ("err_dim", C.c_int32),
        ("H_prior", _dp), ("b_prior", _dp), ("err_prior", _dp), ("Jt_prior_inv", _dp),
        ("pose_out", _dp), ("speedbias_out", _dp), ("inv_depth_out", _dp),
        ("stats", C.POINTER(VioStats)), ("rc", C.c_int32), ("reserved", C.c_int32),
    ]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)

# every symbol include/vio_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "vio_create", "vio_destroy", "vio_last_error", "vio_version", "vio_struct_size", "vio_device_count", "vio_set_graph", "vio_get_dims",
    "vio_set_allreduce", "vio_set_shard", "vio_set_prior", "vio_get_prior", "vio_set_vertices", "vio_get_vertices",
    "vio_solve", "vio_linearize", "vio_chi2", "vio_solve_step", "vio_apply_step", "vio_rollback_step",
    "vio_get_hessian", "vio_get_schur", "vio_get_schur_bsr", "vio_get_delta", "vio_get_b", "vio_get_landmark_diag",
    "vio_get_kernel_ms", "vio_launch_count", "vio_measure_fp64_peak", "vio_dense_accumulate", "vio_dense_chi2",
    "vio_dense_solve", "vio_dense_get", "vio_solve_batched", "vio_solve_batched_lockstep", "vio_lockstep_release", "vio_get_coarse", "vio_preintegrate", "vio_get_solver_ms", "vio_set_points", "vio_get_points", "vio_get_point_system", "vio_marginalize",
    "vio_nccl_unique_id", "vio_nccl_init", "vio_set_nccl_comm", "vio_get_owned_landmarks", "vio_p2p_enabled",
]

_lib = None


def lib():
    """Load libvio_b200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.vio_version.restype = C.c_char_p
        L.vio_struct_size.restype = C.c_size_t
        L.vio_struct_size.argtypes = [C.c_int]
        for which, cls in enumerate((VioGraph, VioLmOpts, VioStats, VioDims)):
            if L.vio_struct_size(which) != C.sizeof(cls):
                raise RuntimeError(f"ABI mismatch: {cls.__name__} is {C.sizeof(cls)} bytes here, "
                                   f"{L.vio_struct_size(which)} in libvio_b200.so")
        L.vio_last_error.restype = C.c_char_p
        L.vio_last_error.argtypes = [C.c_void_p]
        L.vio_launch_count.restype = C.c_int64
        L.vio_launch_count.argtypes = [C.c_void_p]
        L.vio_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.vio_destroy.argtypes = [C.c_void_p]
        L.vio_destroy.restype = None
        L.vio_set_graph.argtypes = [C.c_void_p, C.POINTER(VioGraph)]
        L.vio_get_dims.argtypes = [C.c_void_p, C.POINTER(VioDims)]
        L.vio_set_allreduce.argtypes = [C.c_void_p, ALLREDUCE_FN, C.c_void_p]
        L.vio_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.vio_nccl_unique_id.argtypes = [C.c_void_p]
        L.vio_nccl_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vio_set_nccl_comm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.vio_p2p_enabled.argtypes = [C.c_void_p]
        L.vio_p2p_enabled.restype = C.c_int
        L.vio_set_prior.argtypes = [C.c_void_p, C.c_int32, _dp, _dp, C.c_int32, _dp, _dp]
        L.vio_get_prior.argtypes = [C.c_void_p, _dp, _dp]
        L.vio_set_vertices.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.vio_get_vertices.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.vio_solve.argtypes = [C.c_void_p, C.c_int32, C.POINTER(VioLmOpts), C.POINTER(VioStats)]
        L.vio_linearize.argtypes = [C.c_void_p, C.POINTER(VioLmOpts)]
        L.vio_chi2.argtypes = [C.c_void_p, C.POINTER(VioLmOpts), _dp]
        L.vio_solve_step.argtypes = [C.c_void_p, C.POINTER(VioLmOpts), C.c_double, C.POINTER(C.c_int64)]
        L.vio_apply_step.argtypes = [C.c_void_p, C.POINTER(VioLmOpts)]
        L.vio_rollback_step.argtypes = [C.c_void_p, C.POINTER(VioLmOpts)]
        L.vio_get_hessian.argtypes = [C.c_void_p, C.POINTER(VioLmOpts), _dp, _dp]
        L.vio_get_schur.argtypes = [C.c_void_p, _dp, _dp]
        L.vio_get_schur_bsr.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp]
        L.vio_get_delta.argtypes = [C.c_void_p, _dp, _dp]
        L.vio_get_b.argtypes = [C.c_void_p, _dp, _dp]
        L.vio_get_landmark_diag.argtypes = [C.c_void_p, _dp]
        L.vio_get_kernel_ms.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_int64)]
        L.vio_measure_fp64_peak.argtypes = [C.c_int, _dp]
        _lib = L
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _b(a):
    return None if a is None else a.ctypes.data_as(_bp)


class Scene:
    """Flat description of a graph: numpy arrays with the exact layout of `vio_graph`."""

    def __init__(self):
        self.pose = np.zeros((0, 7))
        self.pose_fixed = None
        self.speedbias = np.zeros((0, 9))
        self.speedbias_fixed = None
        self.pclass_order = None
        self.inv_depth = np.zeros(0)
        self.rp_landmark = np.zeros(0, np.int32)
        self.rp_pose_i = np.zeros(0, np.int32)
        self.rp_pose_j = np.zeros(0, np.int32)
        self.rp_pts_i = np.zeros((0, 3))
        self.rp_pts_j = np.zeros((0, 2))
        self.rp_info = 1.0
        self.rp_loss = LOSS_TRIVIAL
        self.rp_loss_delta = 1.0
        self.ext_pose = -1
        self.q_ic = np.array([0.0, 0.0, 0.0, 1.0])
        self.t_ic = np.zeros(3)
        self.sp_pose = np.zeros(0, np.int32)
        self.sp_p = np.zeros((0, 3))
        self.sp_q = np.zeros((0, 4))
        self.sp_info = np.zeros((0, 36))
        self.imu = None  # dict of arrays, see to_c
        # VertexPointXYZ landmarks + EdgeReprojectionXYZ observations
        self.point_xyz = np.zeros((0, 3))
        self.rx_point = np.zeros(0, np.int32)
        self.rx_pose = np.zeros(0, np.int32)
        self.rx_obs = np.zeros((0, 2))
        self.landmark_fixed = None  # uint8 per inverse-depth landmark / per point (Vertex::SetFixed), None = none fixed
        self.point_fixed = None
        self.gravity = np.array([0.0, 0.0, 9.81])
        self.storage = STORAGE_AUTO
        # generator extras (not part of the graph)
        self.pose_gt = None
        self.inv_depth_gt = None
        self.prior = None  # dict(H, b, err, jt_inv) for v17 windows

    def _norm(self):
        c = np.ascontiguousarray
        self.pose = c(self.pose, np.float64).reshape(-1, 7)
        self.speedbias = c(self.speedbias, np.float64).reshape(-1, 9)
        self.inv_depth = c(self.inv_depth, np.float64)
        for k in ("rp_landmark", "rp_pose_i", "rp_pose_j", "sp_pose"):
            setattr(self, k, c(getattr(self, k), np.int32))
        self.rp_pts_i = c(self.rp_pts_i, np.float64).reshape(-1, 3)
        self.rp_pts_j = c(self.rp_pts_j, np.float64).reshape(-1, 2)
        self.sp_p = c(self.sp_p, np.float64).reshape(-1, 3)
        self.sp_q = c(self.sp_q, np.float64).reshape(-1, 4)
        self.sp_info = c(self.sp_info, np.float64).reshape(-1, 36)
        self.point_xyz = c(self.point_xyz, np.float64).reshape(-1, 3)
        self.rx_point = c(self.rx_point, np.int32)
        self.rx_pose = c(self.rx_pose, np.int32)
        self.rx_obs = c(self.rx_obs, np.float64).reshape(-1, 2)
        if self.pose_fixed is not None:
            self.pose_fixed = c(self.pose_fixed, np.uint8)
        if self.speedbias_fixed is not None:
            self.speedbias_fixed = c(self.speedbias_fixed, np.uint8)
        if self.pclass_order is not None:
            self.pclass_order = c(self.pclass_order, np.int32)
        if self.landmark_fixed is not None:
            self.landmark_fixed = c(self.landmark_fixed, np.uint8)
        if self.point_fixed is not None:
            self.point_fixed = c(self.point_fixed, np.uint8)

    def to_c(self):
        """-> (VioGraph, keepalive list).  Arrays are borrowed: keep `self` alive during the call."""
        self._norm()
        g = VioGraph()
        g.n_pose = self.pose.shape[0]
        g.pose = _d(self.pose)
        g.pose_fixed = _b(self.pose_fixed)
        g.n_speedbias = self.speedbias.shape[0]
        g.speedbias = _d(self.speedbias)
        g.speedbias_fixed = _b(self.speedbias_fixed)
        g.pclass_order = _i(self.pclass_order)
        g.n_landmark = self.inv_depth.shape[0]
        g.inv_depth = _d(self.inv_depth)
        g.n_reproj = self.rp_landmark.shape[0]
        g.rp_landmark = _i(self.rp_landmark)
        g.rp_pose_i = _i(self.rp_pose_i)
        g.rp_pose_j = _i(self.rp_pose_j)
        g.rp_pts_i = _d(self.rp_pts_i)
        g.rp_pts_j = _d(self.rp_pts_j)
        g.rp_info = float(self.rp_info)
        g.rp_loss = int(self.rp_loss)
        g.rp_loss_delta = float(self.rp_loss_delta)
        g.ext_pose = int(self.ext_pose)
        for k in range(4):
            g.q_ic[k] = float(self.q_ic[k])
        for k in range(3):
            g.t_ic[k] = float(self.t_ic[k])
            g.gravity[k] = float(self.gravity[k])
        g.n_se3prior = self.sp_pose.shape[0]
        g.sp_pose = _i(self.sp_pose)
        g.sp_p = _d(self.sp_p)
        g.sp_q = _d(self.sp_q)
        g.sp_info = _d(self.sp_info)
        keep = []
        if self.imu is not None and len(self.imu["pose_i"]) > 0:
            m = self.imu
            for k in ("pose_i", "sb_i", "pose_j", "sb_j"):
                m[k] = np.ascontiguousarray(m[k], np.int32)
            for k in ("sum_dt", "delta_p", "delta_q", "delta_v", "lin_ba", "lin_bg", "jacobian", "covariance"):
                m[k] = np.ascontiguousarray(m[k], np.float64)
            g.n_imu = len(m["pose_i"])
            g.imu_pose_i, g.imu_sb_i, g.imu_pose_j, g.imu_sb_j = _i(m["pose_i"]), _i(m["sb_i"]), _i(m["pose_j"]), _i(m["sb_j"])
            g.imu_sum_dt, g.imu_delta_p, g.imu_delta_q, g.imu_delta_v = _d(m["sum_dt"]), _d(m["delta_p"]), _d(m["delta_q"]), _d(m["delta_v"])
            g.imu_lin_ba, g.imu_lin_bg = _d(m["lin_ba"]), _d(m["lin_bg"])
            g.imu_jacobian, g.imu_covariance = _d(m["jacobian"]), _d(m["covariance"])
            keep.append(m)
        g.storage = int(self.storage)
        g.n_point = self.point_xyz.shape[0]
        g.point_xyz = _d(self.point_xyz)
        g.n_reproj_xyz = self.rx_point.shape[0]
        g.rx_point, g.rx_pose, g.rx_obs = _i(self.rx_point), _i(self.rx_pose), _d(self.rx_obs)
        g.landmark_fixed = _b(self.landmark_fixed)
        g.point_fixed = _b(self.point_fixed)
        return g, keep

    @property
    def P(self):
        return 6 * self.pose.shape[0] + 9 * self.speedbias.shape[0]

    _ARRAYS = ("pose", "pose_fixed", "speedbias", "speedbias_fixed", "pclass_order", "inv_depth", "rp_landmark",
               "rp_pose_i", "rp_pose_j", "rp_pts_i", "rp_pts_j", "q_ic", "t_ic", "sp_pose", "sp_p", "sp_q", "sp_info",
               "gravity", "pose_gt", "inv_depth_gt", "point_xyz", "rx_point", "rx_pose", "rx_obs", "landmark_fixed", "point_fixed")
    _SCALARS = ("rp_info", "rp_loss", "rp_loss_delta", "ext_pose", "storage")

    def export(self):
        """-> dict of numpy arrays (np.savez) holding the whole scene, including IMU constants and prior."""
        self._norm()
        d = {}
        for k in self._ARRAYS:
            v = getattr(self, k)
            if v is not None:
                d[k] = np.asarray(v)
        for k in self._SCALARS:
            d[k] = np.asarray(getattr(self, k))
        if self.imu is not None:
            for k, v in self.imu.items():
                d["imu_" + k] = np.asarray(v)
        if self.prior is not None:
            for k, v in self.prior.items():
                if v is not None:
                    d["prior_" + k] = np.asarray(v)
        return d

    @classmethod
    def from_dict(cls, d):
        s = cls()
        for k in cls._ARRAYS:
            if k in d:
                setattr(s, k, np.array(d[k]))
        for k in cls._SCALARS:
            if k in d:
                v = d[k].item() if hasattr(d[k], "item") else d[k]
                setattr(s, k, v)
        imu = {k[4:]: np.array(d[k]) for k in d if k.startswith("imu_")}
        s.imu = imu or None
        pr = {k[6:]: np.array(d[k]) for k in d if k.startswith("prior_")}
        s.prior = pr or None
        s._norm()
        return s


class VioError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vio_b200 error {code} ({ERR_NAMES.get(code, '?')}): {msg}")
        self.code = code


def make_opts(flavour=LM_V17, solver=SOLVER_AUTO, verbose=0, pcg_max_iter=0, pcg_tol=0.0, fixed_iterations=0,
              warm_start=0):
    o = VioLmOpts()
    o.flavour, o.solver, o.verbose = flavour, solver, verbose
    o.pcg_max_iter, o.pcg_tol, o.fixed_iterations, o.warm_start = pcg_max_iter, pcg_tol, fixed_iterations, warm_start
    return o


class Problem:
    """Thin RAII wrapper over a `vio_problem` handle."""

    def __init__(self, device=0, stream=None):
        self._L = lib()
        self._h = C.c_void_p()
        rc = self._L.vio_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h))
        if rc != VIO_OK:
            raise VioError(rc, "vio_create failed (no CUDA device? the product path has no CPU fallback)")
        self._cb = None
        self.scene = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.vio_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != VIO_OK:
            raise VioError(rc, self._L.vio_last_error(self._h).decode())

    def set_shard(self, rank, world):
        self._ck(self._L.vio_set_shard(self._h, rank, world))

    def owned_landmarks(self):
        """Indices (into the scene's landmark arrays) of the landmarks this handle / rank owns."""
        n = C.c_int64()
        self._L.vio_get_owned_landmarks.argtypes = [C.c_void_p, _ip, C.c_int64, C.POINTER(C.c_int64)]
        self._ck(self._L.vio_get_owned_landmarks(self._h, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), np.int32)
        self._ck(self._L.vio_get_owned_landmarks(self._h, out.ctypes.data_as(_ip), n.value, C.byref(n)))
        return out[:n.value]

    def nccl_init(self, rank, world, unique_id):
        """Native NCCL path: create the communicator inside libvio_b200.so (collective over all ranks) and set the shard.
        unique_id: the 128 bytes of nccl_unique_id() made on rank 0 and shipped to every rank by the caller."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self._L.vio_nccl_init(self._h, rank, world, buf))

    def p2p_enabled(self):
        """True when the small all-reduces of the distributed solve go through the NVLink peer-memory mailbox (vio_p2p.cuh)."""
        return bool(self._L.vio_p2p_enabled(self._h))

    def set_allreduce(self, pyfunc):
        """pyfunc(dev_ptr:int, count:int, stream:int) -> 0 on success."""
        self._cb = ALLREDUCE_FN(lambda ptr, n, st, user: int(pyfunc(ptr, n, st) or 0))
        self._ck(self._L.vio_set_allreduce(self._h, self._cb, None))

    def set_graph(self, scene):
        g, keep = scene.to_c()
        self.scene = scene
        self._ck(self._L.vio_set_graph(self._h, C.byref(g)))
        del keep
        if scene.prior is not None:
            pr = scene.prior
            self.set_prior(pr["H"], pr["b"], pr.get("err"), pr.get("jt_inv"))

    def set_prior(self, H, b, err=None, jt_inv=None):
        H = np.ascontiguousarray(H, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        ed = 0
        if err is not None and len(err) > 0:
            err = np.ascontiguousarray(err, np.float64)
            jt_inv = np.ascontiguousarray(jt_inv, np.float64)
            ed = err.shape[0]
        self._ck(self._L.vio_set_prior(self._h, b.shape[0], _d(H), _d(b), ed, _d(err) if ed else None, _d(jt_inv) if ed else None))

    def get_prior(self):
        d = self.dims()
        b = np.zeros(d.P)
        err = np.zeros(max(d.P - 15, 0))
        self._ck(self._L.vio_get_prior(self._h, _d(b), _d(err)))
        return b, err

    def dims(self):
        d = VioDims()
        self._ck(self._L.vio_get_dims(self._h, C.byref(d)))
        return d

    def solve(self, iterations, opts=None):
        st = VioStats()
        self._ck(self._L.vio_solve(self._h, iterations, C.byref(opts) if opts is not None else None, C.byref(st)))
        return st

    def linearize(self, opts=None):
        self._ck(self._L.vio_linearize(self._h, C.byref(opts) if opts is not None else None))

    def chi2(self, opts=None):
        out = C.c_double()
        self._ck(self._L.vio_chi2(self._h, C.byref(opts) if opts is not None else None, C.byref(out)))
        return out.value

    def solve_step(self, lam, opts=None):
        it = C.c_int64()
        self._ck(self._L.vio_solve_step(self._h, C.byref(opts) if opts is not None else None, lam, C.byref(it)))
        return it.value

    def apply_step(self, opts=None):
        self._ck(self._L.vio_apply_step(self._h, C.byref(opts) if opts is not None else None))

    def rollback_step(self, opts=None):
        self._ck(self._L.vio_rollback_step(self._h, C.byref(opts) if opts is not None else None))

    def set_vertices(self, pose=None, speedbias=None, inv_depth=None):
        a = [None if x is None else np.ascontiguousarray(x, np.float64) for x in (pose, speedbias, inv_depth)]
        self._ck(self._L.vio_set_vertices(self._h, _d(a[0]), _d(a[1]), _d(a[2])))

    def get_vertices(self):
        s = self.scene
        pose = np.zeros_like(s.pose)
        sb = np.zeros_like(s.speedbias)
        invd = np.array(s.inv_depth, copy=True)
        self._ck(self._L.vio_get_vertices(self._h, _d(pose), _d(sb) if sb.size else None, _d(invd) if invd.size else None))
        return pose, sb, invd

    def get_points(self):
        xyz = np.zeros_like(self.scene.point_xyz)
        if xyz.size:
            self._ck(self._L.vio_get_points(self._h, _d(xyz)))
        return xyz

    def set_points(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float64)
        self._ck(self._L.vio_set_points(self._h, _d(xyz)))

    def get_point_system(self):
        n = self.scene.point_xyz.shape[0]
        H, b, dx = np.zeros((n, 3, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        if n:
            self._ck(self._L.vio_get_point_system(self._h, _d(H), _d(b), _d(dx)))
        return H, b, dx

    def get_hessian(self, opts=None):
        d = self.dims()
        n = d.P + d.M + 3 * self.scene.point_xyz.shape[0]
        H = np.zeros((n, n))
        b = np.zeros(n)
        self._ck(self._L.vio_get_hessian(self._h, C.byref(opts) if opts is not None else None, _d(H), _d(b)))
        return H, b

    def get_schur(self):
        d = self.dims()
        S = np.zeros((d.P, d.P))
        bS = np.zeros(d.P)
        self._ck(self._L.vio_get_schur(self._h, _d(S), _d(bS)))
        return S, bS

    def get_schur_bsr(self):
        d = self.dims()
        rowptr = np.zeros(d.n_pose_blocks + 1, np.int32)
        col = np.zeros(d.nnz_blocks, np.int32)
        val = np.zeros((d.nnz_blocks, 6, 6))
        bS = np.zeros(d.P)
        self._ck(self._L.vio_get_schur_bsr(self._h, _i(rowptr), _i(col), _d(val), _d(bS)))
        return rowptr, col, val, bS

    def get_delta(self):
        d = self.dims()
        dp = np.zeros(d.P)
        dl = np.zeros(d.M)
        self._ck(self._L.vio_get_delta(self._h, _d(dp), _d(dl) if d.M else None))
        return dp, dl

    def get_coarse(self):
        nc, ma = C.c_int32(), C.c_int32()
        self._ck(self._L.vio_get_coarse(self._h, C.byref(nc), C.byref(ma), None, None))
        nb = self.dims().n_pose_blocks
        A = np.zeros((nc.value, nc.value))
        Z = np.zeros((nb, 6, 7))
        self._ck(self._L.vio_get_coarse(self._h, None, None, _d(A), _d(Z)))
        return ma.value, A, Z

    def get_b(self):
        d = self.dims()
        bp = np.zeros(d.P)
        bl = np.zeros(d.M)
        self._ck(self._L.vio_get_b(self._h, _d(bp), _d(bl) if d.M else None))
        return bp, bl

    def marginalize(self, marg_pose, marg_sb):
        d = self.dims()
        n = d.P
        H = np.zeros((n, n))
        b = np.zeros(n)
        err = np.zeros(n)
        Jt = np.zeros((n, n))
        dim = C.c_int32()
        self._L.vio_marginalize.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), _dp, _dp, _dp, _dp]
        self._ck(self._L.vio_marginalize(self._h, marg_pose, marg_sb, C.byref(dim), _d(H), _d(b), _d(err), _d(Jt)))
        k = dim.value
        return dict(dim=k, H=H.ravel()[:k * k].reshape(k, k).copy(), b=b[:k].copy(), err=err[:k].copy(),
                    jt_inv=Jt.ravel()[:k * k].reshape(k, k).copy())

    def kernel_ms(self):
        ms = C.c_double()
        n = C.c_int64()
        self._ck(self._L.vio_get_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def solver_ms(self):
        """-> dict: PCG kernel ms per launch, launches, iterations, coarse refresh ms and count (last solve)"""
        a, c, it = C.c_double(), C.c_double(), C.c_double()
        na, nc = C.c_int64(), C.c_int64()
        self._ck(self._L.vio_get_solver_ms(self._h, C.byref(a), C.byref(na), C.byref(it), C.byref(c), C.byref(nc)))
        return dict(pcg_ms=a.value, pcg_launches=na.value, pcg_iterations=it.value, coarse_ms=c.value, coarse_refreshes=nc.value)

    def launch_count(self):
        return int(self._L.vio_launch_count(self._h))


def nccl_unique_id():
    """128-byte ncclUniqueId from the libnccl.so.2 the library loaded (call on rank 0, broadcast to the other ranks)."""
    buf = C.create_string_buffer(128)
    rc = lib().vio_nccl_unique_id(buf)
    if rc != VIO_OK:
        raise VioError(rc, "vio_nccl_unique_id (libnccl.so.2 not loadable?)")
    return buf.raw


def solve_batched(scenes, iterations, opts=None, device=0, n_workers=16, lockstep=False, max_chunk=0):
    """BASELINE config 3: many independent problems through vio_solve_batched (one handle per worker thread) or, with
    lockstep=True, vio_solve_batched_lockstep (the whole batch as one packed graph).

    Returns (list of dicts with pose / speedbias / inv_depth / stats per scene, wall seconds of the call)."""
    import time
    n = len(scenes)
    items = (VioBatchItem * n)()
    keep, outs = [], []
    for i, s in enumerate(scenes):
        g, k = s.to_c()
        gp = C.pointer(g)
        st = VioStats()
        pose = np.zeros_like(s.pose)
        sb = np.zeros_like(s.speedbias)
        invd = np.array(s.inv_depth, copy=True)
        it = items[i]
        it.graph = gp
        if s.prior is not None:
            H = np.ascontiguousarray(s.prior["H"], np.float64)
            b = np.ascontiguousarray(s.prior["b"], np.float64)
            it.prior_dim, it.H_prior, it.b_prior = b.shape[0], _d(H), _d(b)
            keep += [H, b]
            err = s.prior.get("err")
            if err is not None and len(err):
                err = np.ascontiguousarray(err, np.float64)
                jt = np.ascontiguousarray(s.prior["jt_inv"], np.float64)
                it.err_dim, it.err_prior, it.Jt_prior_inv = err.shape[0], _d(err), _d(jt)
                keep += [err, jt]
        it.pose_out = _d(pose)
        it.speedbias_out = _d(sb) if sb.size else None
        it.inv_depth_out = _d(invd) if invd.size else None
        it.stats = C.pointer(st)
        keep += [g, gp, k, s]
        outs.append(dict(pose=pose, speedbias=sb, inv_depth=invd, stats=st))
    L = lib()
    L.vio_solve_batched.argtypes = [C.c_int, C.c_int32, C.POINTER(VioBatchItem), C.c_int64, C.c_int32, C.POINTER(VioLmOpts)]
    L.vio_solve_batched_lockstep.argtypes = [C.c_int, C.POINTER(VioBatchItem), C.c_int64, C.c_int32, C.POINTER(VioLmOpts), C.c_int32]
    t0 = time.perf_counter()
    if lockstep:
        rc = L.vio_solve_batched_lockstep(device, items, n, iterations, C.byref(opts) if opts is not None else None, max_chunk)
    else:
        rc = L.vio_solve_batched(device, n_workers, items, n, iterations, C.byref(opts) if opts is not None else None)
    dt = time.perf_counter() - t0
    if rc != VIO_OK:
        raise VioError(rc, "vio_solve_batched: item errors " + str([items[i].rc for i in range(n) if items[i].rc][:5]))
    return outs, dt


class VioImuSegments(C.Structure):
    _fields_ = [("n_segments", C.c_int32), ("reserved", C.c_int32), ("seg_ptr", C.POINTER(C.c_int32)),
                ("dt", _dp), ("acc", _dp), ("gyr", _dp), ("ba", _dp), ("bg", _dp),
                ("acc_n", C.c_double), ("acc_w", C.c_double), ("gyr_n", C.c_double), ("gyr_w", C.c_double)]


def preintegrate(seg_ptr, dt, acc, gyr, ba, bg, noise, device=0):
    """IntegrationBase::push_back over a batch of IMU segments on the device (vio_preintegrate).
    noise = (ACC_N, ACC_W, GYR_N, GYR_W).  -> dict with the EdgeImu constants per segment."""
    seg_ptr = np.ascontiguousarray(seg_ptr, np.int32)
    n = seg_ptr.shape[0] - 1
    dt, acc, gyr = (np.ascontiguousarray(x, np.float64) for x in (dt, acc, gyr))
    ba, bg = np.ascontiguousarray(ba, np.float64), np.ascontiguousarray(bg, np.float64)
    s = VioImuSegments()
    s.n_segments = n
    s.seg_ptr = seg_ptr.ctypes.data_as(C.POINTER(C.c_int32))
    s.dt, s.acc, s.gyr, s.ba, s.bg = _d(dt), _d(acc), _d(gyr), _d(ba), _d(bg)
    s.acc_n, s.acc_w, s.gyr_n, s.gyr_w = (float(x) for x in noise)
    out = dict(sum_dt=np.zeros(n), delta_p=np.zeros((n, 3)), delta_q=np.zeros((n, 4)), delta_v=np.zeros((n, 3)),
               jacobian=np.zeros((n, 225)), covariance=np.zeros((n, 225)))
    L = lib()
    L.vio_preintegrate.argtypes = [C.c_int, C.POINTER(VioImuSegments), _dp, _dp, _dp, _dp, _dp, _dp]
    rc = L.vio_preintegrate(device, C.byref(s), _d(out["sum_dt"]), _d(out["delta_p"]), _d(out["delta_q"]), _d(out["delta_v"]),
                            _d(out["jacobian"]), _d(out["covariance"]))
    if rc != VIO_OK:
        raise VioError(rc, "vio_preintegrate")
    return out


def measure_fp64_peak(device=0):
    out = C.c_double()
    rc = lib().vio_measure_fp64_peak(device, C.byref(out))
    if rc != VIO_OK:
        raise VioError(rc, "vio_measure_fp64_peak")
    return out.value


def device_count():
    return int(lib().vio_device_count())
