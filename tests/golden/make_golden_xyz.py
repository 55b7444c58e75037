"""Golden vectors for VertexPointXYZ / EdgeReprojectionXYZ graphs from the UNMODIFIED reference backends (oracle/_ref).
Scenes are regenerated deterministically by the tests (tests/scenes_extra.xyz_scene); only reference outputs are stored.

    python tests/golden/make_golden_xyz.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refshim  # noqa: E402
from tests.golden.make_golden import OUT, lin  # noqa: E402
from tests.scenes_extra import xyz_scene  # noqa: E402


def sol(ver, scene, iters, name):
    r = refshim.solve(ver, scene, iters)
    np.savez_compressed(os.path.join(OUT, name), pose=r["pose"], inv_depth=r["inv_depth"], point_xyz=r["point_xyz"],
                        iterations=r["iterations"], chi2_trace=r["chi2_trace"], lambda_trace=r["lambda_trace"],
                        chi2_final=r["chi2_final"], lambda_final=r["lambda_final"])


def main():
    lin(15, xyz_scene("xyz_v15"), "xyz_6x40_v15_lin.npz")
    lin(17, xyz_scene("xyz_v17_cauchy"), "xyz_6x40_v17_cauchy_lin.npz")
    lin(17, xyz_scene("mixed_v17"), "mixed_6x40_v17_lin.npz")
    sol(17, xyz_scene("xyz_v17_solve"), 20, "xyz_20x300_v17_solve.npz")
    sol(17, xyz_scene("mixed_v17_solve"), 20, "mixed_20x300_v17_solve.npz")
    sol(15, xyz_scene("xyz_v15_solve"), 10, "xyz_20x300_v15_solve10.npz")
    print("xyz golden vectors written to", OUT)


if __name__ == "__main__":
    main()
