"""BAL problem I/O: file round trip, the BAL camera model against the EdgeReprojectionXYZ mapping (CPU,
through the oracle), and - on a GPU - a small synthetic BAL problem solved by the backend."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _synthetic_bal(vio, n_cam=12, n_pt=200, seed=3, noise_px=0.3):
    """cameras on an arc looking at a point cloud, BAL conventions (camera looks down -z), every point seen by 6 cameras"""
    bal = vio.bal
    rng = np.random.default_rng(seed)
    cams = np.zeros((n_cam, 9))
    pts = np.stack([rng.uniform(-3, 3, n_pt), rng.uniform(-2, 2, n_pt), rng.uniform(-1, 1, n_pt)], 1)
    for i in range(n_cam):
        ang = -0.5 + i / (n_cam - 1)
        Rwc = bal._rodrigues(np.array([0.0, ang, 0.0]))             # camera-to-world
        C = np.array([8.0 * np.sin(ang), 0.1 * i, 8.0 * np.cos(ang)])  # centre; the camera's -z axis points at the origin
        R = Rwc.T
        cams[i, :3] = bal._log_so3(R)
        cams[i, 3:6] = -R @ C
        cams[i, 6:] = [520.0 + 5 * i, -1e-2, 1e-4]
    ci, pi, obs = [], [], []
    for k in range(n_pt):
        for c in rng.choice(n_cam, 6, replace=False):
            P = bal._rodrigues(cams[c, :3]) @ pts[k] + cams[c, 3:6]
            p = -P[:2] / P[2]
            r2 = p @ p
            ci.append(c); pi.append(k)
            obs.append(cams[c, 6] * (1 + cams[c, 7] * r2 + cams[c, 8] * r2 * r2) * p + rng.normal(0, noise_px, 2))
    return dict(cam_index=np.array(ci, np.int32), pt_index=np.array(pi, np.int32), obs=np.array(obs), cameras=cams, points=pts)


def test_bal_roundtrip_and_camera_model(tmp_path):
    vio = importlib.import_module("visual-inertial-odometry_b200")
    from tests import oraclelib as orc
    b = _synthetic_bal(vio, noise_px=0.0)
    path = str(tmp_path / "problem.txt")
    vio.bal.write_bal(path, b)
    b2 = vio.bal.read_bal(path)
    for k in b:
        assert np.array_equal(b[k], b2[k]), k
    assert vio.bal.bal_reprojection_error(b2) < 1e-9
    # the mapping onto EdgeReprojectionXYZ reproduces the BAL residual: zero noise => zero chi2 in the backend's factor
    s = vio.bal.bal_to_scene(b2)
    assert s.point_xyz.shape == (200, 3) and s.rx_obs.shape[0] == 1200
    assert orc.chi2(s, vio.capi.LM_V15) < 1e-16
    # and with noise the backend's chi2 is the BAL squared error in normalised units
    bn = _synthetic_bal(vio, noise_px=0.5)
    sn = vio.bal.bal_to_scene(bn)
    f = bn["cameras"][bn["cam_index"], 6]
    chi = orc.chi2(sn, vio.capi.LM_V15)
    rms_norm = np.sqrt(chi / len(f))
    assert abs(rms_norm * f.mean() - vio.bal.bal_reprojection_error(bn)) < 0.05 * vio.bal.bal_reprojection_error(bn)
    # scene -> BAL -> scene keeps the estimates
    b3 = vio.bal.scene_to_bal(sn, sn.point_xyz, sn.pose, f=500.0)
    s3 = vio.bal.bal_to_scene(b3)
    assert np.abs(s3.pose - sn.pose).max() < 1e-9 and np.abs(s3.rx_obs - sn.rx_obs).max() < 1e-9


@pytest.mark.gpu
def test_bal_problem_solved_on_device():
    """a perturbed synthetic BAL problem: the backend (XYZ landmarks, v17 LM, exact reduced solve) brings the pixel RMS
    back to the noise level and agrees with the CPU oracle"""
    vio = importlib.import_module("visual-inertial-odometry_b200")
    from tests import oraclelib as orc
    b = _synthetic_bal(vio, noise_px=0.3)
    rng = np.random.default_rng(9)
    b["points"] = b["points"] + rng.normal(0, 0.05, b["points"].shape)
    b["cameras"][2:, 3:6] += rng.normal(0, 0.02, (b["cameras"].shape[0] - 2, 3))
    rms0 = vio.bal.bal_reprojection_error(b)
    s = vio.bal.bal_to_scene(b, fix_first=2)
    opts = vio.make_opts(flavour=vio.capi.LM_V17)
    p = vio.Problem()
    p.set_graph(s)
    st = p.solve(15, opts)
    pose, _, _ = p.get_vertices()
    pts = p.get_points()
    out = vio.bal.scene_to_bal(s, pts, pose)
    out["cameras"][:, 6:] = b["cameras"][:, 6:]
    out["obs"] = b["obs"]
    rms1 = vio.bal.bal_reprojection_error(out)
    assert rms0 > 5.0 and rms1 < 0.5, (rms0, rms1)
    ref = orc.solve(s, 15, opts)
    assert st.iterations == ref["iterations"]
    assert abs(st.chi2_final - ref["chi2_final"]) <= 1e-6 * ref["chi2_final"]
    assert np.abs(pts - ref["point_xyz"]).max() <= 1e-6 * np.abs(ref["point_xyz"]).max()


GOLD = os.path.join(ROOT, "tests", "golden")


def test_bal_fixture_against_the_reference_reader_and_camera_model():
    """tests/golden/bal_fixture.txt through bal.py against the reference side (tests/golden/bal_fixture_ref.npz, written
    by make_golden_bal.py from oracle/_ref/bal_ref = the reference's own BALProblem reader, bal.cpp, unmodified, plus its
    camera model restated from bal_g2o.cpp:25-42, 94-109 with the reference's Sophus): the same indices, observations,
    cameras and points are parsed, and the predicted pixel of every observation agrees to 1e-9 px."""
    vio = importlib.import_module("visual-inertial-odometry_b200")
    g = np.load(os.path.join(GOLD, "bal_fixture_ref.npz"))
    b = vio.bal.read_bal(os.path.join(GOLD, "bal_fixture.txt"))
    assert np.array_equal(b["cam_index"], g["cam_index"]) and np.array_equal(b["pt_index"], g["pt_index"])
    for k in ("obs", "cameras", "points"):
        assert np.array_equal(b[k], g[k]), k  # both sides parse the same decimal strings
    pred = vio.bal.bal_project(b)
    assert np.abs(pred - g["pred"]).max() <= 1e-9
    rms_ref = np.sqrt(((g["pred"] - g["obs"]) ** 2).sum(1).mean())
    assert abs(vio.bal.bal_reprojection_error(b) - rms_ref) <= 1e-9
    # the scene mapping keeps that residual: chi2 of the backend's factor = sum |undistorted residual|^2
    from tests import oraclelib as orc
    s = vio.bal.bal_to_scene(b)
    f = b["cameras"][b["cam_index"], 6]
    rms_norm = np.sqrt(orc.chi2(s, vio.capi.LM_V15) / len(f))
    assert abs(rms_norm * f.mean() - rms_ref) < 0.05 * rms_ref


@pytest.mark.gpu
def test_bal_fixture_solved_on_device():
    """the committed BAL fixture on the device: pixel RMS (the reference-side camera model) goes from 5 px to the
    0.3 px noise level; estimates agree with the CPU oracle"""
    vio = importlib.import_module("visual-inertial-odometry_b200")
    from tests import oraclelib as orc
    b = vio.bal.read_bal(os.path.join(GOLD, "bal_fixture.txt"))
    rms0 = vio.bal.bal_reprojection_error(b)
    s = vio.bal.bal_to_scene(b, fix_first=2)
    opts = vio.make_opts(flavour=vio.capi.LM_V17)
    p = vio.Problem()
    p.set_graph(s)
    st = p.solve(15, opts)
    pose, _, _ = p.get_vertices()
    pts = p.get_points()
    out = vio.bal.scene_to_bal(s, pts, pose)
    out["cameras"][:, 6:] = b["cameras"][:, 6:]
    out["obs"] = b["obs"]
    rms1 = vio.bal.bal_reprojection_error(out)
    assert rms0 > 4.0 and rms1 < 0.5, (rms0, rms1)
    ref = orc.solve(s, 15, opts)
    assert st.iterations == ref["iterations"]
    assert abs(st.chi2_final - ref["chi2_final"]) <= 1e-6 * ref["chi2_final"]
    assert np.abs(pts - ref["point_xyz"]).max() <= 1e-6 * np.abs(ref["point_xyz"]).max()
