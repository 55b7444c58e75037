"""CPU test of the N>1 host logic with world_size 2 on gloo: landmark shards are disjoint, cover the graph, and
the all-reduced partial reduced systems (computed per shard by the oracle) equal the single-rank system."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    import importlib
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    vio = importlib.import_module("visual-inertial-odometry_b200")
    from tests import oraclelib as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = vio.scenes.ring(n_cam=24, n_landmark=240, k_obs=5, seed=11)
    L = s.inv_depth.shape[0]
    eptr = np.concatenate([[0], np.cumsum(np.bincount(s.rp_landmark, minlength=L))])
    cuts = importlib.import_module("visual-inertial-odometry_b200.dist").shard_ranges(eptr, world)
    # full block pattern (ring: every camera pair within k_obs-1) = dense here for simplicity
    C = s.pose.shape[0]
    rowptr = np.arange(C + 1, dtype=np.int32) * C
    col = np.tile(np.arange(C, dtype=np.int32), C)
    val, bS, Hll, bl = orc.linearize_bsr(s, rowptr, col, cuts[rank], cuts[rank + 1])
    t = torch.from_numpy(np.concatenate([val.ravel(), bS]))
    dist.all_reduce(t)
    if rank == 0:
        full_val, full_bS, _, _ = orc.linearize_bsr(s, rowptr, col, 0, L)
        ref = np.concatenate([full_val.ravel(), full_bS])
        out["err"] = float(np.abs(t.numpy() - ref).max() / np.abs(ref).max())
        out["cuts"] = cuts
        out["L"] = L
    dist.barrier()
    dist.destroy_process_group()


def test_landmark_shards_sum_to_full_system():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["err"] <= 1e-12
    cuts = out["cuts"]
    assert cuts[0] == 0 and cuts[-1] == out["L"] and cuts[1] > 0 and cuts[1] < out["L"]


def test_shard_ranges_balance():
    import importlib
    sys.path.insert(0, ROOT)
    d = importlib.import_module("visual-inertial-odometry_b200.dist")
    eptr = np.arange(0, 1001) * 10
    cuts = d.shard_ranges(eptr, 8)
    assert cuts[0] == 0 and cuts[-1] == 1000
    sizes = np.diff(cuts)
    assert sizes.min() >= 124 and sizes.max() <= 126
