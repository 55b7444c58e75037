"""BAL fixture + golden vectors from the reference side; run in the build container:
    python tests/golden/make_golden_bal.py

  bal_fixture.txt       a small problem in the BAL text format of the reference's g2o assignment
                        (07-backend-optimization/01-bal-g2o): 8 cameras, 40 points, 240 observations with 0.3 px noise,
                        points and cameras perturbed (so that there is something to optimise).
  bal_fixture_ref.npz   what the reference's OWN reader (src/bal.cpp BALProblem, compiled unmodified into
                        oracle/_ref/bal_ref next to oracle/ref_bal.cpp) parsed from that file - indices, observations,
                        cameras, points - and the pixel its camera model (src/bal_g2o.cpp:25-42, 94-109: Sophus
                        SO3d::exp, P = R X + t, p = -P.xy / P.z, radial distortion) predicts for every observation.
"""
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
vio = importlib.import_module("visual-inertial-odometry_b200")
from tests.test_bal_io import _synthetic_bal  # noqa: E402

OUT_DIR = os.path.dirname(os.path.abspath(__file__))


def main():
    b = _synthetic_bal(vio, n_cam=8, n_pt=40, seed=11, noise_px=0.3)
    rng = np.random.default_rng(12)
    b["points"] = b["points"] + rng.normal(0, 0.05, b["points"].shape)
    b["cameras"][2:, 3:6] += rng.normal(0, 0.02, (b["cameras"].shape[0] - 2, 3))
    path = os.path.join(OUT_DIR, "bal_fixture.txt")
    vio.bal.write_bal(path, b)
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "bal_ref"), path], capture_output=True, text=True, check=True).stdout
    ci, pi, obs, pred, cams, pts = [], [], [], [], [], []
    for line in out.splitlines():
        t = line.split()
        if not t:
            continue
        if t[0] == "OBS":
            ci.append(int(t[1])); pi.append(int(t[2])); obs.append([float(t[3]), float(t[4])]); pred.append([float(t[5]), float(t[6])])
        elif t[0] == "CAM":
            cams.append([float(x) for x in t[1:]])
        elif t[0] == "PT":
            pts.append([float(x) for x in t[1:]])
    np.savez_compressed(os.path.join(OUT_DIR, "bal_fixture_ref.npz"), cam_index=np.array(ci, np.int32), pt_index=np.array(pi, np.int32),
                        obs=np.array(obs), pred=np.array(pred), cameras=np.array(cams), points=np.array(pts))
    print("observations", len(ci), "cameras", len(cams), "points", len(pts), "rms px", np.sqrt(((np.array(pred) - np.array(obs)) ** 2).sum(1).mean()))


if __name__ == "__main__":
    main()
