"""GPU test of the drop-in boundary: the reference's own driver source (15-vio-backend/app/TestMonoBA.cpp, compiled
UNMODIFIED against include/backend/*.h + libvio_backend.so + libvio_b200.so) runs on the GPU and prints the same
estimates as the reference's own binary (oracle/_ref/test_mono_ba15)."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "build", "test_mono_ba_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "test_mono_ba15")


def _parse(out):
    opt = [float(x) for x in re.findall(r"opt\s+(-?[0-9.]+)", out)]
    cams = [[float(v) for v in m.split()] for m in re.findall(r"optimized:\s+(-?[0-9.]+\s+-?[0-9.]+\s+-?[0-9.]+)", out)]
    iters = len(re.findall(r"^iter: ", out, flags=re.M))
    chi0 = float(re.search(r"iter: 0 , chi= ([0-9.e+-]+)", out).group(1))
    prior = re.search(r"after marg-+\s+([0-9.e+-]+)\s+([0-9.e+-]+)\s+([0-9.e+-]+)\s+([0-9.e+-]+)", out)
    return np.array(opt), np.array(cams), iters, chi0, [float(prior.group(k)) for k in range(1, 5)]


def test_unmodified_reference_driver_links_and_matches():
    if not (os.path.exists(OURS) and os.path.exists(REF)):
        pytest.skip("drop-in demo binaries not built (need /root/reference at build time)")
    ours = subprocess.run([OURS], capture_output=True, text=True, timeout=300)
    assert ours.returncode == 0, ours.stderr[-2000:]
    ref = subprocess.run([REF], capture_output=True, text=True, timeout=300)
    o_opt, o_cam, o_it, o_chi0, o_prior = _parse(ours.stdout)
    r_opt, r_cam, r_it, r_chi0, r_prior = _parse(ref.stdout)
    assert len(o_opt) == len(r_opt) == 20 and o_cam.shape == r_cam.shape == (3, 3)
    assert o_it == r_it
    assert abs(o_chi0 - r_chi0) <= 1e-4 * r_chi0
    # printed with 4 decimals (std::fixed, precision 4)
    assert np.abs(o_opt - r_opt).max() <= 2e-4
    assert np.abs(o_cam - r_cam).max() <= 2e-4
    assert np.allclose(o_prior, r_prior, atol=1e-4)
    assert np.allclose(o_prior, [26.5306, -8.1633, -8.1633, 10.2041], atol=1e-4)


@pytest.mark.parametrize("ours,flav", [("curve_fitting17_b200", 17), ("curve_fitting15_b200", 15)])
def test_unmodified_curve_fitting_driver(ours, flav):
    """GENERIC_PROBLEM with user-defined host Vertex/Edge subclasses: the reference's CurveFitting drivers, compiled
    unmodified against the drop-in headers.  The v17 backend converges to 0.941841 2.09467 0.965537 (probe of the
    unmodified binary); both of our flavours use the damped system (the v15 binary's generic branch is
    broken upstream and returns 0 0 0)."""
    exe = os.path.join(ROOT, "build", ours)
    ref = os.path.join(ROOT, "oracle", "_ref", "curve_fitting17")
    if not os.path.exists(exe):
        pytest.skip("drop-in demo binaries not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    if flav == 17:
        m = re.search(r"we got these parameters :\s*\n\s*([0-9.e+-]+)\s+([0-9.e+-]+)\s+([0-9.e+-]+)", out.stdout)
        got = np.array([float(m.group(k)) for k in (1, 2, 3)])
    else:  # the v15 driver prints the parameter vector as a column after the two timing lines
        tail = out.stdout.split("makeHessian cost:")[-1].split()
        got = np.array([float(x) for x in tail[-3:]])
    assert np.abs(got - np.array([0.941841, 2.09467, 0.965537])).max() <= (2e-5 if flav == 17 else 2e-3)
    if flav == 17 and os.path.exists(ref):
        r = subprocess.run([ref], capture_output=True, text=True, timeout=300)
        it_ref = len(re.findall(r"^iter: ", r.stdout, flags=re.M))
        it_ours = len(re.findall(r"^iter: ", out.stdout, flags=re.M))
        assert it_ours == it_ref
        chi_ref = [float(x) for x in re.findall(r"^iter: \d+ , chi= ([0-9.e+-]+)", r.stdout, flags=re.M)]
        chi_ours = [float(x) for x in re.findall(r"^iter: \d+ , chi= ([0-9.e+-]+)", out.stdout, flags=re.M)]
        assert np.allclose(chi_ours, chi_ref, rtol=1e-4)


def test_xyz_driver_matches_reference_backend():
    """VertexPointXYZ / EdgeReprojectionXYZ through the C++ drop-in: tests/xyz_ba_driver.cc (written for this repo; it
    uses only the API shared with the reference) built against the UNMODIFIED v15 backend and against
    include/backend + libvio_backend.so prints the same estimates.  The v15 flavour solves the reduced system with
    the reference's inexact PCG, so the comparison is to 1e-3."""
    ours = os.path.join(ROOT, "build", "xyz_ba_b200")
    ref = os.path.join(ROOT, "oracle", "_ref", "xyz_ba_ref15")
    if not (os.path.exists(ours) and os.path.exists(ref)):
        pytest.skip("xyz driver binaries not built (need /root/reference at build time)")
    o = subprocess.run([ours], capture_output=True, text=True, timeout=300)
    assert o.returncode == 0, o.stderr[-2000:]
    r = subprocess.run([ref], capture_output=True, text=True, timeout=300)

    def parse(out):
        rows = re.findall(r"^(?:cam|point) \d+ : (-?[0-9.]+) (-?[0-9.]+) (-?[0-9.]+)", out, flags=re.M)
        chi = [float(x) for x in re.findall(r"^iter: \d+ , chi= ([0-9.e+-]+)", out, flags=re.M)]
        return np.array(rows, float), np.array(chi)
    vo, co = parse(o.stdout)
    vr, cr = parse(r.stdout)
    assert vo.shape == vr.shape == (20, 3)
    assert len(co) == len(cr) and np.allclose(co, cr, rtol=1e-3)
    assert np.abs(vo - vr).max() <= 2e-3


def test_frame_stream_driver_matches_reference_backend():
    """A fresh Problem per frame, the way the VINS estimator drives the backend (estimator.cpp:902-1037): the drop-in
    hands its device handle back to a pool when a Problem dies and the next Problem reuses it ('reuse
    previous window's device buffers').  tests/frame_stream_driver.cc built against the UNMODIFIED v15 backend and
    against the drop-in prints the same estimates for five consecutive frames of three different window shapes, a
    repeated frame prints exactly what it printed the first time (nothing leaks from one Problem into the next), and
    only the first frame pays for creating a handle."""
    ours = os.path.join(ROOT, "build", "frame_stream_b200")
    ref = os.path.join(ROOT, "oracle", "_ref", "frame_stream_ref15")
    if not (os.path.exists(ours) and os.path.exists(ref)):
        pytest.skip("frame stream driver binaries not built (need /root/reference at build time)")
    o = subprocess.run([ours], capture_output=True, text=True, timeout=300)
    assert o.returncode == 0, o.stderr[-2000:]
    r = subprocess.run([ref], capture_output=True, text=True, timeout=300)

    def parse(out):
        rows = re.findall(r"^frame (\d+) (?:cam|lm) \d+ : (.*)$", out, flags=re.M)
        per = {}
        for f, vals in rows:
            per.setdefault(int(f), []).extend(float(x) for x in vals.split())
        return {k: np.array(v) for k, v in per.items()}
    po, pr = parse(o.stdout), parse(r.stdout)
    assert sorted(po) == sorted(pr) == [0, 1, 2, 3, 4]
    for f in range(5):
        assert po[f].shape == pr[f].shape
        assert np.abs(po[f] - pr[f]).max() <= 2e-3  # v15 flavour: inexact reference PCG
    # the final RED.F64 flush into S is order-dependent at the 1e-16 level and the v15 flavour's inexact PCG (stopped at
    # |r| <= 1e-6 |b|) turns that into ~1e-5 run-to-run differences; a leak between Problems would be gross
    assert np.abs(po[2] - po[0]).max() <= 1e-4 and np.abs(po[3] - po[1]).max() <= 1e-4
    ms = [float(x) for x in re.findall(r"landmarks\): ([0-9.]+) ms", o.stderr)]
    assert len(ms) == 5
    print("frame wall times (ms):", ms)
    assert ms[2] < ms[0]  # frame 0 creates the CUDA context and the handle; frame 2 (same shape) reuses both
