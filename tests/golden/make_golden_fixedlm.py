"""Golden vectors for FIXED landmark-class vertices, from the UNMODIFIED reference (oracle/_ref); run in the build container:
    python tests/golden/make_golden_fixedlm.py

  fixedlm_6x40_v17_lin.npz   TestMonoBA-style scene (6 cameras, 40 inverse-depth landmarks, fixed extrinsic vertex) with
                             landmarks 3, 17 and 18 fixed: H and b of the unmodified Problem::MakeHessian, which skips the
                             Jacobian blocks of every fixed vertex (A17/src/backend/problem.cc:325,340).
  fixedpt_6x40_v17_lin.npz   the same with VertexPointXYZ landmarks, points 0, 5 and 39 fixed.
Only H and b are pinned: the reference's own Schur step inverts the zero H_mm block of a fixed landmark
(problem.cc:421-425 -> inf / NaN), so S, the step and Solve have no reference value (see include/vio_b200.h).
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
vio = importlib.import_module("visual-inertial-odometry_b200")
from tests import refshim  # noqa: E402

OUT_DIR = os.path.dirname(os.path.abspath(__file__))
from tests.scenes_extra import fixed_scene  # noqa: E402


def main():
    for name, kind in (("fixedlm_6x40_v17_lin.npz", "lm"), ("fixedpt_6x40_v17_lin.npz", "pt")):
        s = fixed_scene(kind)
        H, b = refshim.hessian(17, s)
        chi, lam = refshim.init(17, s)
        np.savez_compressed(os.path.join(OUT_DIR, name), H=H, b=b, chi2=chi, lam=lam)
        P = s.P
        print(name, "H", H.shape, "zero rows:", [int(i) for i in np.where(np.abs(H).sum(1) == 0)[0] if i >= P][:12], "chi2", chi, "lam", lam)


if __name__ == "__main__":
    main()
