"""visual-inertial-odometry_b200 — B200-native backend::Problem least-squares hot path.

The product is libvio_b200.so (hand-written sm_100a CUDA behind the C-ABI of include/vio_b200.h);
this package only holds the ctypes binding (`capi`), the scene generators (`scenes`), BAL dataset I/O (`bal`) and
the landmark-sharding glue for torch.distributed (`dist`).
"""
from . import bal, capi, scenes  # noqa: F401
from .capi import Problem, Scene, make_opts  # noqa: F401

__all__ = ["bal", "capi", "scenes", "Problem", "Scene", "make_opts"]
