"""Debug probe (GPU box): two-level PCG pieces against numpy on a 1000-camera ring."""
import importlib, os, sys
import numpy as np
import scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vio = importlib.import_module("visual-inertial-odometry_b200")
ncam = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
s = vio.scenes.ring(n_cam=ncam, n_landmark=20 * ncam, k_obs=11, seed=9)
s.storage = vio.capi.STORAGE_BSR
p = vio.Problem(); p.set_graph(s)
o1 = vio.make_opts(flavour=vio.capi.LM_V17, solver=vio.capi.SOLVER_BLOCK_PCG, pcg_tol=1e-6)
o2 = vio.make_opts(flavour=vio.capi.LM_V17, solver=vio.capi.SOLVER_BLOCK_PCG_2L, pcg_tol=1e-6)
p.linearize(o1)
rowptr, col, val, bS = p.get_schur_bsr()
nb = len(rowptr) - 1; n = 6 * nb
lam = 1e-9 * np.abs(val).max()
it1 = p.solve_step(lam, o1)
it2 = p.solve_step(lam, o2)
print("gpu iters plain", it1, "two-level", it2)
ma, Ainv, Z = p.get_coarse()
nc = Ainv.shape[0]; na = nc // 7
print("nc", nc, "ma", ma, "nb", nb)
S = sp.bsr_matrix((val, col, rowptr), shape=(n, n)).tocsr()
A = (S + lam * sp.identity(n)).tocsr()
# Z as sparse with the device's aggregates: aggregate of block i from the device layout (grid CTAs x apc)
grid = nc // 7  # apc folded in
# recover aggregate of each block: consecutive chunks; device: c = i // brc, al = min(apc-1, (i % brc)//ma)
import math
dims = p.dims()
# device layout: grid CTAs x apc aggregates; recover (grid, apc) from na and ma
cands = [(g, na // g) for g in range(1, 149) if na % g == 0]
agg = None
for g, apc in cands:
    brc = (nb + g - 1) // g
    if (brc + apc - 1) // apc == ma and g == min(g, (nb + 7) // 8):
        agg = np.array([(i // brc) * apc + min(apc - 1, (i % brc) // ma) for i in range(nb)])
print("layout grid", g, "apc", apc)
rows, cols, vals = [], [], []
for i in range(nb):
    for x in range(6):
        for m in range(7):
            if Z[i, x, m] != 0:
                rows.append(6 * i + x); cols.append(7 * agg[i] + m); vals.append(Z[i, x, m])
Zs = sp.csr_matrix((vals, (rows, cols)), shape=(n, nc))
Ac = (Zs.T @ (A @ Zs)).toarray()
dz = np.where(np.diag(Ac) <= 0)[0]; Ac[dz, dz] = 1
Ai = np.linalg.inv(Ac)
print("cond Ac %.3g" % np.linalg.cond(Ac), "Ainv rel err", np.abs(Ainv - Ai).max() / np.abs(Ai).max(), "asym", np.abs(Ainv - Ainv.T).max() / np.abs(Ainv).max())
print("||Ainv Ac - I||", np.abs(Ainv @ Ac - np.eye(nc)).max(), "numpy:", np.abs(Ai @ Ac - np.eye(nc)).max())
diag_idx = np.array([rowptr[i] + np.searchsorted(col[rowptr[i]:rowptr[i + 1]], i) for i in range(nb)])
Dinv = np.linalg.inv(val[diag_idx] + lam * np.eye(6))
bj = lambda r: np.einsum('kij,kj->ki', Dinv, r.reshape(nb, 6)).ravel()
def pcg(M, tol=1e-6, maxit=3000):
    x = np.zeros(n); r = bS.copy(); z = M(r); pp = z.copy(); rz = r @ z; bb = np.sqrt(bS @ bS)
    for it in range(1, maxit + 1):
        w = A @ pp; al = rz / (pp @ w); x += al * pp; r -= al * w
        if np.sqrt(r @ r) <= tol * bb: return it
        z = M(r); rzn = r @ z; pp = z + (rzn / rz) * pp; rz = rzn
    return maxit
print("numpy iters: two-level (numpy inverse)", pcg(lambda r: bj(r) + Zs @ (Ai @ (Zs.T @ r))), " two-level (device inverse)", pcg(lambda r: bj(r) + Zs @ (Ainv @ (Zs.T @ r))))
