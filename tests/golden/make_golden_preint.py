"""Golden vectors for IMU pre-integration from the UNMODIFIED reference (oracle/_ref/libref17.so ->
IntegrationBase::push_back, A17/include/factor/integration_base.h).  Run in the container that holds /root/reference:

    python tests/golden/make_golden_preint.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refshim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)


def _d(a):
    return a.ctypes.data_as(_dp)


def main():
    L = C.CDLL(os.path.join(refshim.REF_DIR, "libref17.so"))
    rng = np.random.default_rng(17)
    lengths = [1, 2, 11, 21, 40, 200]
    noise = np.array([0.08, 0.00004, 0.004, 2.0e-6])  # ACC_N ACC_W GYR_N GYR_W (the EuRoC config values)
    seg_ptr = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)
    ns = int(seg_ptr[-1])
    dt = 0.005 * (1.0 + 0.1 * rng.standard_normal(ns))
    t = np.cumsum(dt)
    # a smooth motion + sensor noise: gravity-loaded accelerometer, slow rotation
    acc = np.stack([0.8 * np.sin(1.3 * t), 0.5 * np.cos(0.7 * t), 9.81 + 0.3 * np.sin(2.1 * t)], 1) + 0.05 * rng.standard_normal((ns, 3))
    gyr = np.stack([0.3 * np.cos(0.9 * t), -0.2 * np.sin(1.7 * t), 0.4 * np.sin(0.5 * t)], 1) + 0.01 * rng.standard_normal((ns, 3))
    nseg = len(lengths)
    ba = 0.05 * rng.standard_normal((nseg, 3))
    bg = 0.01 * rng.standard_normal((nseg, 3))
    out = dict(seg_ptr=seg_ptr, dt=dt, acc=acc, gyr=gyr, ba=ba, bg=bg, noise=noise,
               sum_dt=np.zeros(nseg), delta_p=np.zeros((nseg, 3)), delta_q=np.zeros((nseg, 4)), delta_v=np.zeros((nseg, 3)),
               jacobian=np.zeros((nseg, 225)), covariance=np.zeros((nseg, 225)))
    for k in range(nseg):
        a, b = seg_ptr[k], seg_ptr[k + 1]
        sd = C.c_double()
        d, ac, gy = np.ascontiguousarray(dt[a:b]), np.ascontiguousarray(acc[a:b]), np.ascontiguousarray(gyr[a:b])
        rc = L.ref17_preintegrate(int(b - a), _d(d), _d(ac), _d(gy), _d(np.ascontiguousarray(ba[k])), _d(np.ascontiguousarray(bg[k])),
                                  _d(noise), C.byref(sd), _d(out["delta_p"][k]), _d(out["delta_q"][k]), _d(out["delta_v"][k]),
                                  _d(out["jacobian"][k]), _d(out["covariance"][k]))
        assert rc == 0
        out["sum_dt"][k] = sd.value
    np.savez_compressed(os.path.join(OUT, "preint_v17.npz"), **out)
    print("wrote preint_v17.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
