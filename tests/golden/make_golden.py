"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference for `make -C oracle ref`):
    python tests/golden/make_golden.py
The scenes themselves are regenerated deterministically by the tests (visual-inertial-odometry_b200.scenes);
only reference OUTPUTS are stored.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
vio = importlib.import_module("visual-inertial-odometry_b200")
from tests import refshim  # noqa: E402
from tests.scenes_extra import marginalize_ref, window_scene  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
capi = vio.capi


def lin(ver, scene, name):
    H, b = refshim.hessian(ver, scene)
    chi, lam = refshim.init(ver, scene)
    S, bS, dx = refshim.step(ver, scene, lam)
    np.savez_compressed(os.path.join(OUT, name), H=H, b=b, chi2=chi, lam=lam, S=S, bS=bS, dx=dx)


def sol(ver, scene, iters, name):
    r = refshim.solve(ver, scene, iters)
    np.savez_compressed(os.path.join(OUT, name), pose=r["pose"], speedbias=r["speedbias"], inv_depth=r["inv_depth"],
                        iterations=r["iterations"], chi2_trace=r["chi2_trace"], lambda_trace=r["lambda_trace"],
                        chi2_final=r["chi2_final"], lambda_final=r["lambda_final"], b_prior=r["b_prior"],
                        err_prior=r["err_prior"])


def main():
    lin(15, vio.scenes.monoba(3, 20), "monoba_3x20_v15_lin.npz")
    lin(17, vio.scenes.monoba(3, 20, with_ext=True), "monoba_3x20_v17_lin.npz")
    s = vio.scenes.monoba(6, 40, with_ext=True)
    s.rp_loss, s.rp_loss_delta, s.rp_info = capi.LOSS_CAUCHY, 1.0, 100.0
    lin(17, s, "monoba_6x40_v17_cauchy_lin.npz")
    s.rp_loss = capi.LOSS_HUBER
    lin(17, s, "monoba_6x40_v17_huber_lin.npz")
    s.rp_loss, s.rp_loss_delta = capi.LOSS_TUKEY, 10.0
    lin(17, s, "monoba_6x40_v17_tukey_lin.npz")
    sol(15, vio.scenes.monoba(20, 300), 10, "monoba_20x300_v15_solve10.npz")
    sol(15, vio.scenes.monoba(20, 300), 100, "monoba_20x300_v15_solve100.npz")
    sol(17, vio.scenes.monoba(20, 300, with_ext=True), 100, "monoba_20x300_v17_solve.npz")
    w, wA = window_scene(seed=2, return_marg_window=True)
    # Problem::Marginalize: window A (frames 0..10, IMU edge 0->1, landmarks hosted in frame 0) -> 156-dim prior
    np.savez_compressed(os.path.join(OUT, "windowA_v17_scene.npz"), **wA.export())
    np.savez_compressed(os.path.join(OUT, "windowA_v17_marg.npz"), **marginalize_ref(wA))
    # second step of the chain: window B WITH its prior, marginalise its oldest frame again
    np.savez_compressed(os.path.join(OUT, "windowB_v17_marg.npz"), **marginalize_ref(w))
    lin(17, w, "window_v17_lin.npz")
    sol(17, w, 10, "window_v17_solve10.npz")
    np.savez_compressed(os.path.join(OUT, "window_v17_scene.npz"), **w.export())
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
