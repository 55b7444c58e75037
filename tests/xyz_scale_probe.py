"""Probe (GPU box): VertexPointXYZ landmarks at BASELINE config-4 size - the ring scene re-parameterised as world points,
block-sparse reduced system, two-level PCG.  Prints timings; not collected by pytest."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vio = importlib.import_module("visual-inertial-odometry_b200")
n_cam, n_lm = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 100000
s0 = vio.scenes.ring(n_cam=n_cam, n_landmark=n_lm, k_obs=11, seed=4)
t0 = time.time()
s = vio.scenes.to_xyz(s0, noise=0.01, seed=1)
print("to_xyz %.1f s: %d points, %d XYZ edges" % (time.time() - t0, s.point_xyz.shape[0], s.rx_point.shape[0]))
s.storage = vio.capi.STORAGE_BSR
for name, src in (("inverse depth", s0), ("xyz", s)):
    src.storage = vio.capi.STORAGE_BSR
    p = vio.Problem()
    t0 = time.time(); p.set_graph(src); t_pack = time.time() - t0
    opts = vio.make_opts(flavour=vio.capi.LM_V17, fixed_iterations=1)
    p.solve(3, opts)
    p.set_vertices(pose=src.pose, inv_depth=src.inv_depth if src.inv_depth.size else None)
    if src.point_xyz.shape[0]:
        p.set_points(src.point_xyz)
    t0 = time.time(); st = p.solve(10, opts); dt = time.time() - t0
    print("%-14s pack %.2f s, Solve(10) %.1f ms (%.2f ms per LM iteration), chi2 %.6g -> %.6g, pcg its %d, lin ms %.3f" %
          (name, t_pack, 1e3 * dt, 1e2 * dt, st.chi2_initial, st.chi2_final, st.pcg_iterations, p.kernel_ms()[0]))
