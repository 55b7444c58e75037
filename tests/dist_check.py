"""Multi-GPU parity check (run under torchrun on a GPU box; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py

Every rank solves the same ring scene landmark-sharded (vio_set_shard + NCCL all-reduce hook); rank 0 also solves it
alone on one GPU and compares: reduced system after the all-reduce, chi2 trace, final poses and landmarks.
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    vio = importlib.import_module("visual-inertial-odometry_b200")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s = vio.scenes.ring(n_cam=300, n_landmark=30000, k_obs=11, seed=21)
    s.storage = vio.capi.STORAGE_BSR
    opts = vio.make_opts(flavour=vio.capi.LM_V17, solver=vio.capi.SOLVER_BLOCK_PCG_2L, pcg_tol=1e-10, fixed_iterations=1)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        p = vio.Problem(device=local, stream=stream.cuda_stream)
        p.set_shard(rank, world)
        p.set_allreduce(importlib.import_module("visual-inertial-odometry_b200.dist").make_allreduce_hook())
        p.set_graph(s)
        p.linearize(opts)
        S, bS = p.get_schur()
        st = p.solve(6, opts)
        pose, _, invd_local = p.get_vertices()
        # landmarks: every rank only owns its shard; gather by summing the owned entries
        d = p.dims()
        own = np.zeros(len(s.inv_depth))
        full = np.array(s.inv_depth, copy=True)
        changed = invd_local != full
        t = torch.from_numpy(np.where(changed, invd_local, 0.0)).cuda()
        c = torch.from_numpy(changed.astype(np.float64)).cuda()
        dist.all_reduce(t)
        dist.all_reduce(c)
        torch.cuda.synchronize()
        invd = np.where(c.cpu().numpy() > 0, t.cpu().numpy(), full)
    ok = True
    if rank == 0:
        q = vio.Problem(device=local)
        q.set_graph(s)
        q.linearize(opts)
        S1, bS1 = q.get_schur()
        st1 = q.solve(6, opts)
        pose1, _, invd1 = q.get_vertices()
        eS = np.abs(S - S1).max() / np.abs(S1).max()
        eb = np.linalg.norm(bS - bS1) / np.linalg.norm(bS1)
        tr, tr1 = np.array(st.chi2_trace[:st.n_trace]), np.array(st1.chi2_trace[:st1.n_trace])
        et = np.abs(tr - tr1).max() / np.abs(tr1).max()
        ep = np.abs(pose - pose1).max()
        el = np.abs(invd - invd1).max()
        print(f"world={world}: S rel {eS:.2e}, bS rel {eb:.2e}, chi2 trace rel {et:.2e}, pose abs {ep:.2e}, "
              f"landmark abs {el:.2e}, chi2 {st.chi2_final:.6g} vs {st1.chi2_final:.6g}, shard landmarks {d.reserved}")
        ok = eS <= 1e-12 and eb <= 1e-11 and et <= 1e-6 and ep <= 1e-6 and el <= 1e-6
        print("DIST_CHECK", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
