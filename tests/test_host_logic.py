"""Host-side logic without a GPU.  Lock-step batch (vio_solve_batched_lockstep): the per-item packs merged by
PackedMerge must equal pack_graph's own batch mode on the caller-concatenated graph, and MakeHessian + Schur run with
the device bodies over the merged pack (tests/host_emul.cu) must give every window's reduced system bit for bit."""
import ctypes as C
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import emul  # noqa: E402

pytestmark = pytest.mark.skipif(not emul.available(), reason="tests/libhost_emul.so not built (__graft_entry__.build())")


def _window(vio):
    path = os.path.join(ROOT, "tests", "golden", "window_v17_scene.npz")
    return vio.Scene.from_dict(dict(np.load(path)))


def _items(vio, ragged):
    base = _window(vio)
    rng = np.random.default_rng(2)
    out = []
    for k in range(4):
        d = base.export()
        d.pop("imu_pose_i", None)  # the emulation covers the reprojection factors; drop IMU edges and the prior
        for key in [x for x in d if x.startswith("imu_") or x.startswith("prior_")]:
            d.pop(key)
        if ragged and k in (1, 3):
            keep_l = d["inv_depth"].shape[0] - 11 * k
            m = d["rp_landmark"] < keep_l
            for key in ("rp_landmark", "rp_pose_i", "rp_pose_j", "rp_pts_i", "rp_pts_j"):
                d[key] = d[key][m]
            d["inv_depth"] = d["inv_depth"][:keep_l]
        s = vio.Scene.from_dict(d)
        s.pose[1:, :3] += rng.normal(0, 0.01, (s.pose.shape[0] - 1, 3))
        s.inv_depth *= 1.0 + rng.normal(0, 0.02, s.inv_depth.shape[0])
        out.append(s)
    return out


def _concat(vio, items):
    s0 = items[0]
    Cn, NSB = s0.pose.shape[0], s0.speedbias.shape[0]
    for s in items:
        s._norm()
    d = s0.export()
    d["pose"] = np.vstack([s.pose for s in items])
    d["pose_fixed"] = np.concatenate([s.pose_fixed for s in items])
    d["speedbias"] = np.vstack([s.speedbias for s in items])
    d["speedbias_fixed"] = np.concatenate([np.atleast_1d(s.speedbias_fixed) if s.speedbias_fixed is not None
                                           else np.zeros(NSB, np.uint8) for s in items])
    order = []
    for k, s in enumerate(items):
        o = np.asarray(s.pclass_order, np.int64)
        order.append(np.where(o >= 0, o + k * Cn, ~((~o) + k * NSB)))
    d["pclass_order"] = np.concatenate(order).astype(np.int32)
    loff = np.cumsum([0] + [s.inv_depth.shape[0] for s in items])
    d["inv_depth"] = np.concatenate([s.inv_depth for s in items])
    d["rp_landmark"] = np.concatenate([s.rp_landmark + loff[k] for k, s in enumerate(items)]).astype(np.int32)
    d["rp_pose_i"] = np.concatenate([s.rp_pose_i + k * Cn for k, s in enumerate(items)]).astype(np.int32)
    d["rp_pose_j"] = np.concatenate([s.rp_pose_j + k * Cn for k, s in enumerate(items)]).astype(np.int32)
    d["rp_pts_i"] = np.vstack([s.rp_pts_i for s in items])
    d["rp_pts_j"] = np.vstack([s.rp_pts_j for s in items])
    return vio.Scene.from_dict(d)


@pytest.mark.parametrize("ragged", [False, True])
def test_merged_pack_equals_batch_pack_and_linearises_per_window(ragged):
    vio = importlib.import_module("visual-inertial-odometry_b200")
    items = _items(vio, ragged)
    cat = _concat(vio, items)
    B = len(items)
    P = items[0].P
    gs = [s.to_c() for s in items]
    arr = (C.POINTER(vio.capi.VioGraph) * B)(*[C.pointer(g) for g, _ in gs])
    gc, keep = cat.to_c()
    nd = C.c_int(-1)
    S = np.zeros((B, P, P))
    bS = np.zeros((B, P))
    L = emul.lib()
    L.emul_merge_check.argtypes = [C.POINTER(C.POINTER(vio.capi.VioGraph)), C.c_int, C.POINTER(vio.capi.VioGraph),
                                   C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rc = L.emul_merge_check(arr, B, C.byref(gc), C.byref(nd), S.ctypes.data_as(C.POINTER(C.c_double)),
                            bS.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    assert nd.value == 0
    for k, s in enumerate(items):
        Sk, bk = emul.schur(s)
        assert np.array_equal(S[k], Sk)
        assert np.array_equal(bS[k], bk)


@pytest.mark.parametrize("nb,kind", [(40, "ring"), (60, "random")])
def test_block_cholesky_symbolic_factorisation(nb, kind):
    """Host logic of VIO_SOLVER_BLOCK_CHOL: the symbolic factorisation (fill pattern + update map, csrc/vio_bchol.h)
    driven by a CPU restatement of the device loops solves random SPD block-sparse systems - a band with wrap-around
    border (camera ring) and an irregular pattern - to rounding."""
    rng = np.random.default_rng(nb)
    pat = [set([i]) for i in range(nb)]
    if kind == "ring":
        for i in range(nb):
            for d in range(1, 5):
                j = (i + d) % nb
                pat[i].add(j); pat[j].add(i)
    else:
        for _ in range(3 * nb):
            i, j = rng.integers(0, nb, 2)
            pat[i].add(int(j)); pat[j].add(int(i))
    A = np.zeros((6 * nb, 6 * nb))
    for i in range(nb):
        for j in pat[i]:
            if j > i:
                B = rng.normal(size=(6, 6))
                A[6 * i:6 * i + 6, 6 * j:6 * j + 6] = B
                A[6 * j:6 * j + 6, 6 * i:6 * i + 6] = B.T
    A += np.diag(np.abs(A).sum(1) + 1.0)  # diagonally dominant => SPD
    rowptr, col, val = [0], [], []
    for i in range(nb):
        for j in sorted(pat[i]):
            col.append(j)
            val.append(A[6 * i:6 * i + 6, 6 * j:6 * j + 6].copy())
        rowptr.append(len(col))
    rowptr, col = np.array(rowptr, np.int32), np.array(col, np.int32)
    val = np.ascontiguousarray(np.array(val))
    b = rng.normal(size=6 * nb)
    x = np.zeros(6 * nb)
    nnzL = C.c_longlong()
    L = emul.lib()
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.emul_bchol_solve.argtypes = [C.c_int, ip, ip, dp, C.c_double, dp, dp, C.POINTER(C.c_longlong)]
    lam = 0.37
    rc = L.emul_bchol_solve(nb, rowptr.ctypes.data_as(ip), col.ctypes.data_as(ip), val.ctypes.data_as(dp), lam,
                            b.ctypes.data_as(dp), x.ctypes.data_as(dp), C.byref(nnzL))
    assert rc == 0
    ref = np.linalg.solve(A + lam * np.eye(6 * nb), b)
    assert np.abs(x - ref).max() <= 1e-11 * np.abs(ref).max()
    assert nnzL.value >= (len(col) + nb) // 2  # at least the lower triangle of A


def test_packer_rejects_what_the_device_path_does_not_cover():
    """vio_set_graph's checks (csrc/vio_pack.h), exercised through the CPU harness: bad indices are VIO_ERR_INVALID,
    graphs outside the documented preconditions are VIO_ERR_UNSUPPORTED - never a silent wrong answer."""
    vio = importlib.import_module("visual-inertial-odometry_b200")
    capi = vio.capi
    INVALID, UNSUPPORTED = capi.VIO_ERR_INVALID, capi.VIO_ERR_UNSUPPORTED

    def rc_of(scene):
        g, keep = scene.to_c()
        P = scene.P
        S, bS = np.zeros((max(P, 1), max(P, 1))), np.zeros(max(P, 1))
        dp = C.POINTER(C.c_double)
        return emul.lib().emul_linearize(C.byref(g), 1, S.ctypes.data_as(dp), None, bS.ctypes.data_as(dp), None, None, None, None, None)

    ok = vio.scenes.monoba(4, 30)
    assert rc_of(ok) == 0
    s = vio.scenes.monoba(4, 30)
    s.rp_pose_j[3] = 99                                   # pose index out of range
    assert rc_of(s) == INVALID
    s = vio.scenes.monoba(4, 30)
    s.rp_landmark[0] = -1
    assert rc_of(s) == INVALID
    s = vio.scenes.monoba(4, 30)
    k = int(np.nonzero(s.rp_landmark == s.rp_landmark[5])[0][-1])
    s.rp_pose_i[k] = (s.rp_pose_i[k] + 1) % 4              # edges of one landmark disagree on the host pose
    assert rc_of(s) == UNSUPPORTED
    s = _window(vio)
    s.storage = capi.STORAGE_BSR                          # block-sparse storage with speed-bias vertices
    assert rc_of(s) == UNSUPPORTED
    s = vio.scenes.monoba(4, 30, with_ext=True)
    s.pose_fixed[0] = 0                                   # free extrinsic vertex: fine on a dense single problem ...
    assert rc_of(s) == 0
    s.storage = capi.STORAGE_BSR                          # ... but not with block-sparse storage
    assert rc_of(s) == UNSUPPORTED
    s = vio.scenes.to_xyz(vio.scenes.monoba(4, 30))
    s.rx_point[2] = 10 ** 6
    assert rc_of(s) == INVALID


def test_reference_scenes_are_grouped_for_the_shared_memory_kernel():
    """The production linearise kernel needs every landmark in a group (same host, <= 22 pose slots, shared-memory
    budget); the packer must manage that for the reference-shaped scenes, otherwise the slow generic kernel would run."""
    vio = importlib.import_module("visual-inertial-odometry_b200")

    def info(scene):
        g, keep = scene.to_c()
        ok, ng, nt, sm = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
        assert emul.lib().emul_pack_info(C.byref(g), C.byref(ok), C.byref(ng), C.byref(nt), C.byref(sm)) == 0
        return ok.value, ng.value, nt.value, sm.value

    for s, min_groups in ((vio.scenes.monoba(20, 300), 3), (vio.scenes.monoba(20, 300, with_ext=True), 3), (_window(vio), 8),
                          (vio.scenes.ring(n_cam=200, n_landmark=4000, k_obs=11, seed=1), 40)):
        ok, ng, nt, sm = info(s)
        assert ok == 1 and ng >= min_groups, (ok, ng)
        assert 128 <= nt <= 320 and sm <= 200 * 1024
    # a landmark observed twice by the same pose cannot be grouped: falls back to the generic kernel, still packs
    s = vio.scenes.monoba(6, 40)
    k = int(np.nonzero(s.rp_landmark == 3)[0][1])
    s.rp_pose_j[k] = s.rp_pose_j[int(np.nonzero(s.rp_landmark == 3)[0][0])]
    assert info(s)[0] == 0


def _band_system(nb, w, closed, iso, rng):
    """Random SPD block matrix: pose blocks `iso` are isolated (diagonal only), the others form a chain (closed: ring)
    with half bandwidth w in chain order."""
    chain = [i for i in range(nb) if i not in iso]
    nc = len(chain)
    pat = [set([i]) for i in range(nb)]
    for a in range(nc):
        for d in range(1, w + 1):
            c = a + d
            if c >= nc:
                if not closed:
                    continue
                c -= nc
            i, j = chain[a], chain[c]
            if i != j:
                pat[i].add(j); pat[j].add(i)
    A = np.zeros((6 * nb, 6 * nb))
    for i in range(nb):
        for j in pat[i]:
            if j > i:
                B = rng.normal(size=(6, 6))
                A[6 * i:6 * i + 6, 6 * j:6 * j + 6] = B
                A[6 * j:6 * j + 6, 6 * i:6 * i + 6] = B.T
    A += np.diag(np.abs(A).sum(1) + 1.0)
    rowptr, col, val = [0], [], []
    for i in range(nb):
        for j in sorted(pat[i]):
            col.append(j)
            val.append(A[6 * i:6 * i + 6, 6 * j:6 * j + 6].copy())
        rowptr.append(len(col))
    return A, np.array(rowptr, np.int32), np.array(col, np.int32), np.ascontiguousarray(np.array(val))


@pytest.mark.parametrize("nb,w,closed,iso", [(40, 4, True, ()), (41, 4, True, (0,)), (37, 3, True, (5, 20)), (36, 4, False, ()),
                                             (130, 10, True, (0,)), (24, 8, True, ()), (12, 4, True, ()), (101, 10, True, (0,)),
                                             (64, 2, True, ()), (65, 2, False, (64,)), (250, 5, True, ()), (30, 11, True, ())])
def test_block_cyclic_reduction_plan_solves_band_systems(nb, w, closed, iso):
    """Host logic of VIO_SOLVER_BCR (csrc/vio_bcr.h): the node partition, level schedule, coupling bookkeeping and item
    dependencies, interpreted on the CPU with dense loops (tests/host_emul.cu::emul_bcr_solve), solve random SPD block
    band systems - rings and open chains, odd and even node counts at every level, ragged last nodes, isolated (fixed)
    pose blocks - to rounding."""
    rng = np.random.default_rng(1000 * nb + w)
    A, rowptr, col, val = _band_system(nb, w, closed, set(iso), rng)
    b = rng.normal(size=6 * nb)
    x = np.zeros(6 * nb)
    L = emul.lib()
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.emul_bcr_solve.argtypes = [C.c_int, ip, ip, dp, C.c_double, dp, dp, ip]
    info = np.zeros(8, np.int32)
    lam = 0.21
    rc = L.emul_bcr_solve(nb, rowptr.ctypes.data_as(ip), col.ctypes.data_as(ip), val.ctypes.data_as(dp), lam,
                          b.ctypes.data_as(dp), x.ctypes.data_as(dp), info.ctypes.data_as(ip))
    nc = nb - len(iso)
    if (nc // w) < 3 or 6 * (-(-nc // (nc // w)) + (-(-nc // (nc // w))) % 2) > 72:
        assert rc == importlib.import_module("visual-inertial-odometry_b200").capi.VIO_ERR_UNSUPPORTED
        return
    assert rc == 0, rc
    ref = np.linalg.solve(A + lam * np.eye(6 * nb), b)
    assert np.abs(x - ref).max() <= 1e-11 * np.abs(ref).max()
    n, wq, M, levels, items, slots = info[:6]
    assert wq == w and n == nc // w and M % 12 == 0
    assert levels == int(np.ceil(np.log2(n))) + 1


def test_block_cyclic_reduction_rejects_other_patterns():
    """Patterns that are not a narrow cyclic block band (random long-range couplings, a band wider than the shared-memory
    tile) are refused, so VIO_SOLVER_AUTO keeps the PCG for them."""
    rng = np.random.default_rng(7)
    capi = importlib.import_module("visual-inertial-odometry_b200").capi
    L = emul.lib()
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.emul_bcr_solve.argtypes = [C.c_int, ip, ip, dp, C.c_double, dp, dp, ip]
    for nb, w in ((60, 13), (200, 40)):
        A, rowptr, col, val = _band_system(nb, w, True, set(), rng)
        b, x = rng.normal(size=6 * nb), np.zeros(6 * nb)
        rc = L.emul_bcr_solve(nb, rowptr.ctypes.data_as(ip), col.ctypes.data_as(ip), val.ctypes.data_as(dp), 0.1,
                              b.ctypes.data_as(dp), x.ctypes.data_as(dp), None)
        assert rc == capi.VIO_ERR_UNSUPPORTED


@pytest.mark.parametrize("nb,w,closed,iso,world", [(80, 4, True, (), 2), (81, 4, True, (0,), 2), (130, 10, True, (0,), 2), (160, 5, True, (), 4),
                                                   (163, 5, True, (7,), 4), (240, 5, True, (), 8), (250, 5, True, (), 8), (96, 3, False, (), 4),
                                                   (90, 3, True, (), 3), (64, 4, True, (), 8)])
def test_block_cyclic_reduction_multi_rank_plan(nb, w, closed, iso, world):
    """Host logic of the multi-GPU reduced solve (csrc/vio_bcr.h: BcrDistPlan): every rank eliminates the open chain of its
    own nodes between two pinned interface nodes, the interface system is summed over the ranks and solved, the interiors
    are back-substituted.  Played rank after rank on the CPU (tests/host_emul.cu::emul_bcr_dist_solve) it solves the same
    random SPD band systems as the single-rank plan, with the interface nodes' blocks split between neighbouring ranks."""
    rng = np.random.default_rng(77 * nb + world)
    A, rowptr, col, val = _band_system(nb, w, closed, set(iso), rng)
    b = rng.normal(size=6 * nb)
    x = np.zeros(6 * nb)
    L = emul.lib()
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.emul_bcr_dist_solve.argtypes = [C.c_int, ip, ip, dp, C.c_double, dp, C.c_int, dp]
    lam = 0.17
    rc = L.emul_bcr_dist_solve(nb, rowptr.ctypes.data_as(ip), col.ctypes.data_as(ip), val.ctypes.data_as(dp), lam,
                               b.ctypes.data_as(dp), world, x.ctypes.data_as(dp))
    nc = nb - len(iso)
    if (nc // w) < 2 * world:
        assert rc == importlib.import_module("visual-inertial-odometry_b200").capi.VIO_ERR_UNSUPPORTED
        return
    assert rc == 0, rc
    ref = np.linalg.solve(A + lam * np.eye(6 * nb), b)
    assert np.abs(x - ref).max() <= 1e-11 * np.abs(ref).max()
