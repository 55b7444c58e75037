"""Golden vectors for the 4-vertex EdgeReprojection with a FREE extrinsic vertex, from the unmodified v17 backend.

    python tests/golden/make_golden_extfree.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden.make_golden import lin, sol  # noqa: E402
from tests.scenes_extra import extfree_scene  # noqa: E402

if __name__ == "__main__":
    lin(17, extfree_scene(6, 40), "extfree_6x40_v17_lin.npz")
    sol(17, extfree_scene(20, 300), 20, "extfree_20x300_v17_solve.npz")
    print("ext-free golden vectors written")
