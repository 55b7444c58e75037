"""Solve the config-2 window once (profiling helper for ncu launch lists; not a pytest)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vio = importlib.import_module("visual-inertial-odometry_b200")
s = vio.Scene.from_dict(dict(np.load(os.path.join(ROOT, "tests", "golden", "window_v17_scene.npz"))))
p = vio.Problem()
opts = vio.make_opts(flavour=vio.capi.LM_V17)
for rep in range(3):
    p.set_graph(s)
    t0 = time.perf_counter()
    st = p.solve(10, opts)
    dt = time.perf_counter() - t0
print(f"window solve: {st.iterations} iterations, wall {dt*1e3:.2f} ms, device {st.ms_total:.2f} ms, linearise {st.ms_linearize:.3f} ms, chi2 {st.chi2_final:.6g}")
