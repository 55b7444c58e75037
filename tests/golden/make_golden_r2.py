"""Round-2 golden vectors from the UNMODIFIED reference (oracle/_ref); run in the build container:
    python tests/golden/make_golden_r2.py

  hessian_nullspace_singular_values.txt   stdout of the UNMODIFIED oracle/_ref/hessian_nullspace_test binary
                                         (14-sliding-window/src/hessian_nullspace_test.cpp, 6 printed digits): the
                                         known answer for J^T J of pose(6) + XYZ(3) reprojection blocks.
  monoba_6x40_v17_huber_inlier_lin.npz   HuberLoss with delta so large that every edge is an inlier: the one Huber
                                         regime in which RobustInfo (A17/src/backend/edge.cc:39-74) is free of the
                                         rho1 + 2 rho2 e2 == 0 tie, so H is pinned as well as b and chi2.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
vio = importlib.import_module("visual-inertial-odometry_b200")
from tests import refshim  # noqa: E402
from tests.golden.make_golden import lin  # noqa: E402

capi = vio.capi
OUT_DIR = os.path.dirname(os.path.abspath(__file__))
HUBER_INLIER_DELTA = 1e4


def main():
    import subprocess
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "hessian_nullspace_test")], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(OUT_DIR, "hessian_nullspace_singular_values.txt"), "w") as f:
        f.write(out)
    s = vio.scenes.monoba(6, 40, with_ext=True)
    s.rp_loss, s.rp_loss_delta, s.rp_info = capi.LOSS_HUBER, HUBER_INLIER_DELTA, 100.0
    lin(17, s, "monoba_6x40_v17_huber_inlier_lin.npz")
    # sanity: the same scene without a loss function must give the same H (Huber == trivial for inliers)
    s2 = vio.scenes.monoba(6, 40, with_ext=True)
    s2.rp_info = 100.0
    H, b = refshim.hessian(17, s)
    H2, b2 = refshim.hessian(17, s2)
    print("huber-inlier vs trivial loss: max|dH|/max|H| =", np.abs(H - H2).max() / np.abs(H2).max())


if __name__ == "__main__":
    main()
