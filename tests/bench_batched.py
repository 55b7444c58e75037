"""BASELINE config 3 measurement (run on a GPU box): N perturbed copies of the config-2 window through
vio_solve_batched; prints one JSON line with problems/s and summed edges/s, and the CPU oracle's rate beside it.

    python tests/bench_batched.py --n 4096 --workers 16
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


_BASE = None


def _cpu_worker(arg):
    k, n = arg
    vio = importlib.import_module("visual-inertial-odometry_b200")
    from tests import oraclelib as orc
    base = vio.Scene.from_dict(dict(np.load(os.path.join(ROOT, "tests", "golden", "window_v17_scene.npz"))))
    rng = np.random.default_rng(100 + k)
    opts = vio.make_opts(flavour=vio.capi.LM_V17)
    for _ in range(n):
        s = vio.Scene.from_dict(base.export())
        s.pose[1:, :3] += rng.normal(0, 0.01, (s.pose.shape[0] - 1, 3))
        orc.solve(s, 10, opts)
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--workers", type=int, default=16)
    ap.add_argument("--cpu-n", type=int, default=4)
    ap.add_argument("--lockstep", action="store_true", help="vio_solve_batched_lockstep instead of the thread pool")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--cpu-procs", type=int, default=0, help="also run the CPU oracle in this many processes (0 = skip)")
    args = ap.parse_args()
    vio = importlib.import_module("visual-inertial-odometry_b200")
    from tests import oraclelib as orc
    base = vio.Scene.from_dict(dict(np.load(os.path.join(ROOT, "tests", "golden", "window_v17_scene.npz"))))
    rng = np.random.default_rng(3)
    distinct = []
    for k in range(min(args.n, 64)):
        s = vio.Scene.from_dict(base.export())
        s.pose[1:, :3] += rng.normal(0, 0.01, (s.pose.shape[0] - 1, 3))
        s.inv_depth *= 1.0 + rng.normal(0, 0.02, s.inv_depth.shape[0])
        distinct.append(s)
    scenes = [distinct[i % len(distinct)] for i in range(args.n)]
    opts = vio.make_opts(flavour=vio.capi.LM_V17)
    kw = dict(n_workers=args.workers, lockstep=args.lockstep, max_chunk=args.chunk)
    # warm-up (contexts, allocations); the lock-step entry keeps its handle + staging, so warm it at the measured size
    vio.capi.solve_batched(scenes if args.lockstep else scenes[:args.workers], 10, opts, **kw)
    outs, dt = vio.capi.solve_batched(scenes, 10, opts, **kw)
    E = int(base.rp_landmark.shape[0])
    iters = sum(o["stats"].iterations for o in outs)
    t0 = time.perf_counter()
    for s in distinct[:args.cpu_n]:
        orc.solve(s, 10, opts)
    t_cpu = (time.perf_counter() - t0) / args.cpu_n
    cpu_all = None
    if args.cpu_procs > 0:
        # the same oracle in N independent processes (the reference itself is single-threaded): aggregate problems/s
        import multiprocessing as mp
        per = max(1, args.cpu_n)
        with mp.get_context("fork").Pool(args.cpu_procs) as pool:
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(k, per) for k in range(args.cpu_procs)])
            cpu_all = args.cpu_procs * per / (time.perf_counter() - t0)
    lat = np.array([o["stats"].ms_total for o in outs])
    print(json.dumps({"metric": "batched_windows_per_sec", "value": args.n / dt, "unit": "problems/s", "n_problems": args.n,
                      "workers": args.workers, "lockstep": bool(args.lockstep), "chunk": args.chunk, "wall_s": dt, "lm_iterations_total": iters,
                      "edges_per_sec": E * iters / dt, "edges_per_problem": E, "P": base.P,
                      "cpu_oracle_port_problems_per_sec_1core": 1.0 / t_cpu,
                      "cpu_oracle_port_problems_per_sec_all_procs": cpu_all, "cpu_procs": args.cpu_procs,
                      "device_ms_per_problem_p50": float(np.percentile(lat, 50)), "device_ms_per_problem_p95": float(np.percentile(lat, 95)),
                      "latency_note": "vio_stats.ms_total per item: thread pool = that problem's own Solve on its stream; "
                                      "lock-step = the chunk's LM loop (all items of a chunk finish together)",
                      "note": "each problem: pack + H2D + Solve(10) + D2H through vio_solve_batched" + ("_lockstep" if args.lockstep else "")}))


if __name__ == "__main__":
    main()
