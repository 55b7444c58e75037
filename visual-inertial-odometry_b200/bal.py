"""BAL ("Bundle Adjustment in the Large") problem I/O.

The reference reads BAL text files in its g2o assignment (workspace/assignments/07-backend-optimization/01-bal-g2o/
include/bal.hpp:4-91, src/bal.cpp): header `num_cameras num_points num_observations`, then one line per observation
`camera_index point_index x y` (pixels, origin at the image centre), then 9 numbers per camera (angle-axis r, translation t,
focal length f, radial distortion k1, k2) and 3 per point, one number per line.  The BAL camera model is
    P = R(r) X + t,   p = -P.xy / P.z,   pixel = f (1 + k1 |p|^2 + k2 |p|^4) p.

This module maps such a problem onto the backend's own factor, EdgeReprojectionXYZ over VertexPointXYZ / VertexPose
(A15/backend/edge_reprojection.cc:113-163), with the INTRINSICS HELD FIXED at the file's values (the backend has no
intrinsics vertex): pixels are undistorted into normalised image coordinates once, the camera pose becomes the
body-to-world VertexPose (R^T, -R^T t) and the BAL convention "camera looks down -z" becomes the constant extrinsic
rotation R_ic = diag(1, -1, -1), t_ic = 0.  With that, r = p_c.xy / p_c.z - obs is exactly the BAL residual divided by
f (1 + k1 |p|^2 + ...), i.e. the same minimiser up to the per-observation weight.
"""
import numpy as np

from .capi import Scene, STORAGE_AUTO


def _rodrigues(r):
    th = np.linalg.norm(r)
    K = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


def _log_so3(R):
    c = np.clip((np.trace(R) - 1) / 2, -1.0, 1.0)
    th = np.arccos(c)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    if th < 1e-9:
        return w
    return w * th / np.sin(th)


def _R2q(R):
    """rotation matrix -> quaternion xyzw (positive w)"""
    t = np.trace(R)
    if t > 0:
        s = 0.5 / np.sqrt(t + 1.0)
        q = np.array([(R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s, 0.25 / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q = np.zeros(4)
        q[i] = 0.5 * s
        s = 0.5 / s
        q[3] = (R[k, j] - R[j, k]) * s
        q[j] = (R[j, i] + R[i, j]) * s
        q[k] = (R[k, i] + R[i, k]) * s
    return q if q[3] >= 0 else -q


def _q2R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def read_bal(path):
    """-> dict(cam_index, pt_index, obs[n,2] pixels, cameras[nc,9], points[np,3])"""
    with open(path) as f:
        tok = f.read().split()
    nc, npt, no = int(tok[0]), int(tok[1]), int(tok[2])
    a = np.array(tok[3:3 + 4 * no], dtype=np.float64).reshape(no, 4)
    rest = np.array(tok[3 + 4 * no:3 + 4 * no + 9 * nc + 3 * npt], dtype=np.float64)
    if rest.size != 9 * nc + 3 * npt:
        raise ValueError("truncated BAL file")
    return dict(cam_index=a[:, 0].astype(np.int32), pt_index=a[:, 1].astype(np.int32), obs=a[:, 2:4].copy(),
                cameras=rest[:9 * nc].reshape(nc, 9).copy(), points=rest[9 * nc:].reshape(npt, 3).copy())


def write_bal(path, bal):
    """the inverse of read_bal (same layout as BALProblem::WriteToFile, bal.cpp)"""
    nc, npt, no = bal["cameras"].shape[0], bal["points"].shape[0], bal["obs"].shape[0]
    with open(path, "w") as f:
        f.write(f"{nc} {npt} {no}\n")
        for c, p, (x, y) in zip(bal["cam_index"], bal["pt_index"], bal["obs"]):
            f.write(f"{int(c)} {int(p)} {x:.17g} {y:.17g}\n")
        for v in bal["cameras"].ravel():
            f.write(f"{v:.17g}\n")
        for v in bal["points"].ravel():
            f.write(f"{v:.17g}\n")


def undistort(pix, f, k1, k2, iters=20):
    """pixel (BAL, centred) -> normalised p with  pixel = f (1 + k1 |p|^2 + k2 |p|^4) p  (fixed-point iteration)"""
    d = pix / f
    p = d.copy()
    for _ in range(iters):
        r2 = (p * p).sum(-1, keepdims=True)
        p = d / (1.0 + k1 * r2 + k2 * r2 * r2)
    return p


def bal_to_scene(bal, rp_info=1.0, fix_first=2):
    """BAL problem -> Scene of VertexPose cameras + VertexPointXYZ points + EdgeReprojectionXYZ observations (intrinsics
    fixed, see the module docstring).  fix_first cameras are held fixed (gauge)."""
    cams, pts = bal["cameras"], bal["points"]
    nc = cams.shape[0]
    s = Scene()
    pose = np.zeros((nc, 7))
    for i in range(nc):
        R = _rodrigues(cams[i, :3])
        pose[i, :3] = -R.T @ cams[i, 3:6]
        pose[i, 3:] = _R2q(R.T)
    s.pose = pose
    s.pose_fixed = np.zeros(nc, np.uint8)
    s.pose_fixed[:fix_first] = 1
    s.point_xyz = pts.copy()
    ci = bal["cam_index"]
    p = undistort(bal["obs"], cams[ci, 6:7], cams[ci, 7:8], cams[ci, 8:9])
    # BAL: p = -P.xy / P.z with the camera looking down -z;  ours: p_c = R_ic^T p_b with R_ic = diag(1, -1, -1):
    # p_c.xy / p_c.z = (P.x, -P.y) / (-P.z) = (-P.x / P.z, P.y / P.z) = (p.x, -p.y)
    s.rx_obs = np.stack([p[:, 0], -p[:, 1]], 1)
    s.rx_point = bal["pt_index"].astype(np.int32)
    s.rx_pose = ci.astype(np.int32)
    s.q_ic = np.array([1.0, 0.0, 0.0, 0.0])  # 180 degrees about x
    s.t_ic = np.zeros(3)
    s.ext_pose = -1
    s.rp_info = rp_info
    s.storage = STORAGE_AUTO
    return s


def scene_to_bal(scene, points, poses, f=500.0, k1=0.0, k2=0.0):
    """estimates (poses[n,7] body-to-world, points[m,3]) of a scene built by bal_to_scene -> BAL dict with pixel
    observations re-distorted with (f, k1, k2); used to write results back in the dataset's own format"""
    nc = poses.shape[0]
    cams = np.zeros((nc, 9))
    for i in range(nc):
        Rwb = _q2R(poses[i, 3:7])
        R = Rwb.T
        cams[i, :3] = _log_so3(R)
        cams[i, 3:6] = -R @ poses[i, :3]
        cams[i, 6:] = [f, k1, k2]
    p = np.stack([scene.rx_obs[:, 0], -scene.rx_obs[:, 1]], 1)
    r2 = (p * p).sum(-1, keepdims=True)
    pix = f * (1 + k1 * r2 + k2 * r2 * r2) * p
    return dict(cam_index=scene.rx_pose.copy(), pt_index=scene.rx_point.copy(), obs=pix, cameras=cams, points=points.copy())


def bal_project(bal):
    """predicted pixel of every observation with the file's own camera model (VertexPoseAndIntrinsics::project,
    01-bal-g2o/src/bal_g2o.cpp:94-109) -> [n_obs, 2]"""
    cams, pts = bal["cameras"], bal["points"]
    out = np.zeros((len(bal["obs"]), 2))
    for i, (c, k) in enumerate(zip(bal["cam_index"], bal["pt_index"])):
        P = _rodrigues(cams[c, :3]) @ pts[k] + cams[c, 3:6]
        p = -P[:2] / P[2]
        r2 = p @ p
        out[i] = cams[c, 6] * (1 + cams[c, 7] * r2 + cams[c, 8] * r2 * r2) * p
    return out


def bal_reprojection_error(bal):
    """RMS BAL residual in pixels with the file's own camera model (independent of the scene conversion)"""
    cams, pts = bal["cameras"], bal["points"]
    e2 = 0.0
    for c, k, o in zip(bal["cam_index"], bal["pt_index"], bal["obs"]):
        P = _rodrigues(cams[c, :3]) @ pts[k] + cams[c, 3:6]
        p = -P[:2] / P[2]
        r2 = p @ p
        e2 += ((cams[c, 6] * (1 + cams[c, 7] * r2 + cams[c, 8] * r2 * r2) * p - o) ** 2).sum()
    return np.sqrt(e2 / len(bal["obs"]))
