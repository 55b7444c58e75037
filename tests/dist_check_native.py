"""Multi-GPU parity check of the NATIVE path (run under torchrun on a box with >= 2 GPUs; tests/test_gpu_multi.py launches it):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_check_native.py

Every rank solves the same camera ring with node-range landmark shards: NCCL communicator created inside libvio_b200.so
(vio_nccl_init), distributed block cyclic reduction, exchange steps on the NVLink peer-memory kernel (csrc/vio_p2p.cuh),
trial step replayed as a CUDA graph.  Rank 0 also solves the scene alone and compares: reduced system (summed over the
ranks), chi2 trace, iteration counts, final poses and the landmarks every rank owns.  A second pass repeats the sharded
solve with VIO_B200_NO_P2P=1 (NCCL all-reduces, plain launches) - both must give the same answer.
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sharded_solve(vio, s, opts, rank, world, local, want_p2p):
    vdist = importlib.import_module("visual-inertial-odometry_b200.dist")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        p = vio.Problem(device=local, stream=stream.cuda_stream)
        vdist.init_native_nccl(p, rank, world)
        assert p.p2p_enabled() == want_p2p, (p.p2p_enabled(), want_p2p)
        p.set_graph(s)
        p.linearize(opts)
        rowptr, col, val, bS = p.get_schur_bsr()  # collective: the sum over the ranks
        st = p.solve(8, opts)
        pose, _, invd_local = p.get_vertices()
        owned = p.owned_landmarks()
        full = np.array(s.inv_depth, copy=True)
        mask = np.zeros(len(full))
        mask[owned] = 1.0
        t = torch.from_numpy(np.where(mask > 0, invd_local, 0.0)).cuda()
        c = torch.from_numpy(mask).cuda()
        dist.all_reduce(t)
        dist.all_reduce(c)
        torch.cuda.synchronize()
        assert float(c.min()) == 1.0 and float(c.max()) == 1.0, "every landmark must be owned by exactly one rank"
        invd = t.cpu().numpy()
    return st, val, bS, pose, invd


def main():
    vio = importlib.import_module("visual-inertial-odometry_b200")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s = vio.scenes.ring(n_cam=600, n_landmark=60000, k_obs=11, seed=21)
    s.storage = vio.capi.STORAGE_BSR
    opts = vio.make_opts(flavour=vio.capi.LM_V17)
    ok = True
    results = []
    for want_p2p in (True, False):
        if want_p2p:
            os.environ.pop("VIO_B200_NO_P2P", None)
        else:
            os.environ["VIO_B200_NO_P2P"] = "1"
        results.append(sharded_solve(vio, s, opts, rank, world, local, want_p2p))
    if rank == 0:
        q = vio.Problem(device=local)
        q.set_graph(s)
        q.linearize(opts)
        _, _, val1, bS1 = q.get_schur_bsr()
        st1 = q.solve(8, opts)
        pose1, _, invd1 = q.get_vertices()
        tr1 = np.array(st1.chi2_trace[:st1.n_trace])
        for name, (st, val, bS, pose, invd) in zip(("p2p+graph", "nccl"), results):
            tr = np.array(st.chi2_trace[:st.n_trace])
            # the single-GPU tap holds the upper block triangle only when the cyclic reduction is the solver: compare the blocks both have
            m = (val1 != 0) & (val != 0)
            eS = np.abs(val - val1)[m].max() / np.abs(val1).max()
            eb = np.linalg.norm(bS - bS1) / np.linalg.norm(bS1)
            same_len = len(tr) == len(tr1)
            et = np.abs(tr - tr1).max() / np.abs(tr1).max() if same_len else np.inf
            ep = np.abs(pose - pose1).max()
            el = np.abs(invd - invd1).max()
            print(f"world={world} {name}: S rel {eS:.2e}, bS rel {eb:.2e}, chi2 trace rel {et:.2e}, pose abs {ep:.2e}, landmark abs {el:.2e}, "
                  f"iterations {st.iterations} vs {st1.iterations}, solver {st.solver_used}")
            ok = ok and eS <= 1e-9 and eb <= 1e-9 and et <= 1e-6 and ep <= 1e-6 and el <= 1e-6 and st.iterations == st1.iterations
            ok = ok and st.solver_used == vio.capi.SOLVER_BCR
        print("DIST_CHECK_NATIVE", "PASS" if ok else "FAIL")
    flag = torch.tensor([1.0 if ok else 0.0]).cuda()
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if float(flag.item()) == 1.0 else 1)


if __name__ == "__main__":
    main()
