#!/usr/bin/env python
"""bench.py — LM hot-path throughput on synthetic BA (BASELINE.json: "reprojection edges/sec + LM iters/sec").

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, landmark-sharded when N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, bounded sample

A "step" is one Levenberg-Marquardt iteration of backend::Problem::Solve over the whole scene: reduced solve +
back-substitution + state update + chi2 pass + (accepted) re-linearisation with MakeHessian + Schur.
`value` = E x K / t for one vio_solve(K) with all inputs resident in HBM (CUDA events, max over ranks);
`e2e` = the same call with HOST state buffers: vio_set_vertices (pinned H2D) -> vio_solve(K) -> vio_get_vertices (D2H),
graph resident; `dropin` = a complete Problem::Solve(K) on a fresh problem: vio_set_graph (pack + upload of the edge
records) -> vio_solve(K) -> vio_get_vertices.  At N > 1 rank 0 also re-solves unsharded and reports `parity_vs_n1`.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "visual-inertial-odometry_b200"

WORKLOADS = {
    # BASELINE.json configs[4] / configs[3]; ring generator of SURVEY.md §8(d)
    "config5_ba_10k_cams_1m_landmarks_10m_obs": dict(n_cam=10000, n_landmark=1000000, k_obs=11, seed=5),
    "config4_ba_1k_cams_100k_landmarks_1m_obs": dict(n_cam=1000, n_landmark=100000, k_obs=11, seed=4),
    "tiny_ba_100_cams_10k_landmarks": dict(n_cam=100, n_landmark=10000, k_obs=11, seed=3),
}
DEFAULT_WORKLOAD = "config5_ba_10k_cams_1m_landmarks_10m_obs"
# BASELINE.json configs[0..2]: single problems that fit one CTA's worth of reduced system (dense storage, P = 120 / 126 / 171)
SMALL_WORKLOADS = {
    "config1_monoba_20x300_v15": dict(kind="monoba", ver=15),   # TestMonoBA as committed: v15 LM + reference PCG (missing first step)
    "config1_monoba_20x300_v17": dict(kind="monoba", ver=17),   # the same scene, v17 LM + exact reduced solve, fixed identity extrinsic vertex
    "config2_vins_window": dict(kind="window", ver=17),          # 11 keyframes, 1000 features, 10 IMU edges, 171-dim marginalisation prior
    "config3_batched_4096_windows": dict(kind="batch", ver=17, n=4096),  # config 2 from seeds 2..4097, one lock-step batch call
}
METRIC = "reprojection_edges_per_sec"
UNIT = "edges/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False

    def run(self):
        # NVML in-process (a sample every few ms: the timed region is ~0.1 s); nvidia-smi subprocess as the fallback
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append([str(self.gpu), str(sm), str(mx), "", hex(r)] +
                                    ["Active" if r & bits[n] else "Not Active" for n in
                                     ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
                time.sleep(0.005)
            return
        except Exception:
            pass
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm, reasons, mx = [], set(), 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx = max(mx, float(s[2]))
                for n, v in zip(names, s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_traffic(workload, kernel):
    """dram bytes per launch from the committed ncu --set full capture (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)[workload][kernel]
        return d["dram_read_bytes"] + d["dram_write_bytes"]
    except Exception:
        return None


def alg_bytes_linearize(E, L, C, nnzb):
    """SURVEY.md §8(d) per-pass figure for linearise+Schur: 52 E + 24 L + 56 C + 8 (nnz(S) + P)."""
    return 52.0 * E + 24.0 * L + 56.0 * C + 8.0 * (36.0 * nnzb + 6.0 * C)


def alg_flops_linearize(E, k_obs):
    """400 linearise + 416 JtWJ/Jtr + Schur (n(n+1)+3n)/(K-1), n = 6K, per edge."""
    n = 6.0 * k_obs
    return E * (400.0 + 416.0 + (n * (n + 1) + 3 * n) / (k_obs - 1))


def sub_scene(vio, s, lm_begin, lm_end, with_ext):
    """Landmarks [lm_begin, lm_end) of a ring scene with their edges and the poses they touch (renumbered)."""
    e0 = int(np.searchsorted(s.rp_landmark, lm_begin, side="left"))
    e1 = int(np.searchsorted(s.rp_landmark, lm_end, side="left"))
    poses = np.unique(np.concatenate([s.rp_pose_i[e0:e1], s.rp_pose_j[e0:e1]]))
    remap = -np.ones(s.pose.shape[0], np.int64)
    base = 1 if with_ext else 0
    remap[poses] = np.arange(len(poses)) + base
    t = vio.Scene()
    t.pose = s.pose[poses]
    if with_ext:
        t.pose = np.vstack([np.array([[0, 0, 0, 0, 0, 0, 1.0]]), t.pose])
        t.pose_fixed = np.zeros(t.pose.shape[0], np.uint8)
        t.pose_fixed[0] = 1
        t.ext_pose = 0
    t.inv_depth = s.inv_depth[lm_begin:lm_end].copy()
    t.rp_landmark = (s.rp_landmark[e0:e1] - lm_begin).astype(np.int32)
    t.rp_pose_i = remap[s.rp_pose_i[e0:e1]].astype(np.int32)
    t.rp_pose_j = remap[s.rp_pose_j[e0:e1]].astype(np.int32)
    t.rp_pts_i = s.rp_pts_i[e0:e1].copy()
    t.rp_pts_j = s.rp_pts_j[e0:e1].copy()
    # gauge: SE3 priors on the first two cameras of the sample, like the generator's cameras 0 and 1
    t.sp_pose = np.array([base, base + 1], np.int32)
    t.sp_p = s.pose_gt[poses[:2], :3].copy()
    t.sp_q = s.pose_gt[poses[:2], 3:].copy()
    t.sp_info = np.tile((np.eye(6) * 1e4).reshape(1, 36), (2, 1))
    return t


def cpu_reference_rate(vio, scene, steps, warmup, target_s=12.0, full=False):
    """The reference's own CPU path, one core (the reference has no active threading).

    oracle/_ref present -> `sparse17`: Problem::Solve restated with block-sparse containers around the reference's OWN
    compiled Edge / Vertex code (oracle/ref_sparse17.cpp; reproduces the unmodified Problem::Solve to 1e-13 where that
    fits in memory) - on the whole workload (full=True, the reference arm) or on the landmarks of its first cameras
    (the bounded cpu_baseline leg).  Otherwise the plain-C oracle port's linearisation pass.
    """
    from tests import oraclelib as orc
    from tests import refshim
    L = scene.inv_depth.shape[0]
    out = {"cores": 1}
    if refshim.available(17) and hasattr(refshim, "sparse_solve"):
        if full:
            sub, n_lm = scene, L
            if sub.ext_pose < 0:
                sub = sub_scene(vio, scene, 0, L, with_ext=True)
        else:
            # ~4 us per edge and pass, three passes per LM iteration: size the sample for about target_s seconds
            n_lm = int(min(L, max(2000, target_s / (2 * 3 * 4e-6) / 10)))
            sub = sub_scene(vio, scene, 0, n_lm, with_ext=True)
        E_s = sub.rp_landmark.shape[0]
        K = max(1, min(int(steps), 2))
        r = refshim.sparse_solve(sub, K, fixed_iterations=True)
        its = max(int(r["iterations"]), 1)
        # steady-state time per LM iteration: the initial MakeHessian + chi2 (1 of its+1 linearisations / chi2 passes) left out
        t = (r["t_total"] - r["t_linearize"] / (its + 1) - r["t_chi2"] / (its + 1)) / its
        out.update(kind="reference", value=E_s / t, unit=UNIT, ms_per_step=t * 1e3, steps_run=its, same_config=bool(full),
                   breakdown_s={"linearize": r["t_linearize"], "reduced_solve_backsub": r["t_solve"], "chi2": r["t_chi2"], "total": r["t_total"]},
                   sample=(f"reference Edge/Vertex code + block-sparse containers + Eigen SimplicialLDLT (oracle/ref_sparse17.cpp), "
                           f"Solve({K}) on {'the whole workload' if full else f'landmarks [0,{n_lm}) of the workload'} "
                           f"({E_s} edges, {sub.pose.shape[0]} poses); edges/s = E / steady-state seconds per LM iteration"))
    else:
        orc.lib()
        n_cal = min(20000, L)
        t0 = time.perf_counter()
        orc.linearize_sample(scene, 0, n_cal)
        t_cal = time.perf_counter() - t0
        n_lm = int(min(L, max(n_cal, n_cal * (target_s / max(steps, 1)) / max(t_cal, 1e-6))))
        e1 = int(np.searchsorted(scene.rp_landmark, n_lm, side="left"))
        times = []
        for k in range(steps):
            t0 = time.perf_counter()
            orc.linearize_sample(scene, 0, n_lm)
            times.append(time.perf_counter() - t0)
            if sum(times) > target_s:
                break
        t = float(np.mean(times))
        out.update(kind="port", value=e1 / t, unit=UNIT, ms_per_step=t * 1e3, steps_run=len(times), same_config=False,
                   sample=f"plain-C oracle port, MakeHessian+Schur pass over landmarks [0,{n_lm}) ({e1} edges); reduced "
                          f"solve not included")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vio = importlib.import_module(PKG)
    wl = WORKLOADS[args.workload]
    scene = vio.scenes.ring(**wl)
    r = cpu_reference_rate(vio, scene, args.steps, min(args.warmup, 1), target_s=60.0, full=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps_run"], "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, **wl},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                         "same_config": r.get("same_config"), "breakdown_s": r.get("breakdown_s")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": host_info(),
    }
    emit(line)


SOLVERS = {"auto": "SOLVER_AUTO", "bcr": "SOLVER_BCR", "two_level": "SOLVER_BLOCK_PCG_2L", "block_jacobi": "SOLVER_BLOCK_PCG",
           "block_cholesky": "SOLVER_BLOCK_CHOL"}
SOLVER_NAMES = {0: "auto", 1: "dense_cholesky", 2: "reference_pcg", 3: "block_pcg_6x6_block_jacobi", 4: "block_pcg_6x6_two_level",
                5: "block_sparse_cholesky", 6: "block_cyclic_reduction"}


def parity_vs_unsharded(vio, capi, scene, opts, steps, sharded, stream):
    """Rank 0 only: the same Solve(K) on ONE GPU, unsharded, compared with what the N-rank run produced."""
    import torch
    with torch.cuda.stream(stream):
        q = vio.Problem(device=torch.cuda.current_device(), stream=stream.cuda_stream)
        q.set_graph(scene)
        q.linearize(opts)
        _, _, val1, bS1 = q.get_schur_bsr()
        st1 = q.solve(steps, opts)
        pose1, _, invd1 = q.get_vertices()
    n = min(st1.n_trace, len(sharded["trace"]))
    tr1, trN = np.array(st1.chi2_trace[:n]), np.array(sharded["trace"][:n])
    out = {
        "S_rel": float(np.abs(sharded["val"] - val1).max() / np.abs(val1).max()),
        "bS_rel": float(np.linalg.norm(sharded["bS"] - bS1) / np.linalg.norm(bS1)),
        "chi2_trace_rel": float(np.abs(trN / tr1 - 1.0).max()) if n else None,
        "chi2_final_rel": float(abs(sharded["chi2_final"] / st1.chi2_final - 1.0)),
        "pose_abs": float(np.abs(sharded["pose"] - pose1).max()),
        "inv_depth_abs_owned": float(np.abs(sharded["invd"][sharded["owned"]] - invd1[sharded["owned"]]).max()),
        "iterations_equal": bool(st1.iterations == sharded["iterations"]),
        "tolerances": {"S_rel": 1e-9, "chi2": 1e-6, "pose_abs": 1e-6},
    }
    out["ok"] = bool(out["S_rel"] <= 1e-9 and out["bS_rel"] <= 1e-9 and out["chi2_final_rel"] <= 1e-6 and
                     (out["chi2_trace_rel"] is None or out["chi2_trace_rel"] <= 1e-6) and out["pose_abs"] <= 1e-6 and
                     out["inv_depth_abs_owned"] <= 1e-6 and out["iterations_equal"])
    return out


def run_ours(args):
    import torch
    vio = importlib.import_module(PKG)
    capi = vio.capi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = WORKLOADS[args.workload]
    t0 = time.perf_counter()
    scene = vio.scenes.ring(**wl)
    t_gen = time.perf_counter() - t0
    E = int(scene.rp_landmark.shape[0])
    L = int(scene.inv_depth.shape[0])
    C = int(scene.pose.shape[0])
    scene.storage = capi.STORAGE_BSR

    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        p = vio.Problem(device=local_rank, stream=stream.cuda_stream)
        if world > 1:
            vdist = importlib.import_module(PKG + ".dist")
            if args.collective == "native":
                vdist.init_native_nccl(p, rank, world)  # NCCL communicator inside libvio_b200.so, no Python in the loop
            else:
                p.set_shard(rank, world)
                p.set_allreduce(vdist.make_allreduce_hook())
        t0 = time.perf_counter()
        p.set_graph(scene)
        t_pack = time.perf_counter() - t0
        d = p.dims()
        nnzb = int(d.nnz_blocks)
        E_local = int(d.n_reproj)

        def barrier():
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        def max_over_ranks(x):
            if dist is None:
                return x
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        solver = getattr(capi, SOLVERS[args.solver])
        cold = vio.make_opts(flavour=capi.LM_V17, solver=solver, fixed_iterations=1, pcg_max_iter=args.pcg_max_iter)
        # ---- HBM-resident timing: the reference's Solve(K) from the perturbed initial state, K LM iterations on the
        # natural damping schedule (lambda0 = 1e-5 max diag, shrinking with every accepted step).  Warm-up = the same
        # Solve for W iterations (builds the per-graph solver tables), then the initial state is restored (untimed).
        st_w = p.solve(args.warmup, cold)
        p.set_vertices(pose=scene.pose, inv_depth=scene.inv_depth)
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
        launches0 = p.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        st = p.solve(args.steps, cold)
        ev1.record(stream)
        barrier()
        launches = p.launch_count() - launches0
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        lin_ms, lin_n = p.kernel_ms()  # CUDA events around the linearise kernel on the handle's stream
        sol = p.solver_ms()      # the same around every reduced-solve launch (and coarse-preconditioner refresh)
        lin_ms = max_over_ranks(lin_ms)
        sharded = None
        if world > 1 and not args.no_parity:
            pose_n, _, invd_n = p.get_vertices()
            # landmarks this rank owns (the others keep their initial values in get_vertices)
            owned = np.zeros(L, bool)
            owned[p.owned_landmarks()] = True
            sharded = {"pose": pose_n, "invd": invd_n, "owned": owned, "trace": list(st.chi2_trace[:st.n_trace]),
                       "chi2_final": st.chi2_final, "iterations": st.iterations}
        # ---- end to end through the C-ABI with host state buffers -------------------------------------------
        pose_h = torch.empty((C, 7), dtype=torch.float64).pin_memory()
        invd_h = torch.empty((L,), dtype=torch.float64).pin_memory()
        pose_np, invd_np = pose_h.numpy(), invd_h.numpy()
        pose_np[:] = scene.pose
        invd_np[:] = scene.inv_depth
        # the user-facing call on a resident graph: state from pinned host memory -> Solve(K) -> state back, twice
        e2e_calls = 2
        p.set_vertices(pose=pose_np, inv_depth=invd_np)
        p.solve(1, cold)
        p.get_vertices()
        barrier()
        t0 = time.perf_counter()
        e2e_iters = 0
        for k in range(e2e_calls):
            p.set_vertices(pose=pose_np, inv_depth=invd_np)
            st_e = p.solve(args.steps, cold)
            e2e_iters += st_e.iterations
            po, _, iv = p.get_vertices()
        barrier()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        # ---- the drop-in call: what one Problem::Solve(K) on a fresh Problem costs - graph packing + upload of the edge
        # records from host arrays (vio_set_graph), Solve(K), read-back of the estimates
        barrier()
        t0 = time.perf_counter()
        p.set_graph(scene)
        t_sg = time.perf_counter() - t0
        st_d = p.solve(args.steps, cold)
        p.get_vertices()
        barrier()
        t_dropin = max_over_ranks(time.perf_counter() - t0)
        if sharded is not None:
            p.set_vertices(pose=scene.pose, inv_depth=scene.inv_depth)
            p.linearize(cold)
            _, _, sharded["val"], sharded["bS"] = p.get_schur_bsr()
        if sampler:
            sampler.stop_flag = True
            sampler.join(timeout=2)
        fp64_peak = capi.measure_fp64_peak(local_rank) if rank == 0 else None
        parity = None
        if rank == 0 and sharded is not None:
            parity = parity_vs_unsharded(vio, capi, scene, cold, args.steps, sharded, stream)

    if rank == 0:
        value = E * st.iterations / (ms * 1e-3)
        hbm_peak, peak_src = measured_peaks()
        E_k, L_k = E_local, int(d.reserved)
        bytes_alg = alg_bytes_linearize(E_k, L_k, C, nnzb)
        flops_alg = alg_flops_linearize(E_k, wl["k_obs"])
        ach_gbs = bytes_alg / (lin_ms * 1e-3) / 1e9 if lin_ms > 0 else None
        ach_tf = flops_alg / (lin_ms * 1e-3) / 1e12 if lin_ms > 0 else None
        solver_used = SOLVER_NAMES.get(int(st.solver_used), str(st.solver_used))
        red_t = sol["pcg_ms"] * sol["pcg_launches"] * 1e-3   # seconds inside the timed reduced-solve launches
        hbm_frac = (ach_gbs / hbm_peak) if ach_gbs else None
        fp64_frac = (ach_tf / fp64_peak) if (ach_tf and fp64_peak) else None
        tr_e, tr_s = measured_traffic(args.workload, "k_lin_edges"), measured_traffic(args.workload, "k_schur_groups")
        roof_lin = {"kernel": "k_lin_edges + k_schur_groups (MakeHessian + Schur: edge kernel, then the DMMA Schur kernel; one CTA per "
                              "landmark group each, timed together)",
                    "bound": "fp64" if (fp64_frac and hbm_frac and fp64_frac > hbm_frac) else "hbm",
                    "achieved": ach_tf if (fp64_frac and hbm_frac and fp64_frac > hbm_frac) else ach_gbs,
                    "peak": fp64_peak if (fp64_frac and hbm_frac and fp64_frac > hbm_frac) else hbm_peak,
                    "unit": "TFLOP/s" if (fp64_frac and hbm_frac and fp64_frac > hbm_frac) else "GB/s",
                    "frac": max(x for x in (hbm_frac, fp64_frac) if x is not None) if (hbm_frac or fp64_frac) else None,
                    "hbm_frac": hbm_frac, "fp64_frac": fp64_frac,
                    "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": bytes_alg},
                    "fp64": {"achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                             "algorithmic_flops_per_launch": flops_alg,
                             "peak_source": "measured DFMA micro-benchmark (vio_measure_fp64_peak) in this run"},
                    "traffic": (tr_e + tr_s) if (world == 1 and tr_e and tr_s) else None,
                    "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                    "kernel_ms": lin_ms, "kernel_launches_timed": int(lin_n),
                    "note": "SURVEY 8(d): achieved := max(bytes_alg/t/BW_peak, flops_alg/t/FP64_peak); MakeHessian + Schur is "
                            "FP64-pipe bound (21 FLOP/B against a ridge of ~5.5); traffic = both kernels (the H_lp rows are written "
                            "by the first and read back by the second and by back-substitution)"}
        if int(st.solver_used) == capi.SOLVER_BCR:
            w = wl["k_obs"] - 1
            n_nodes, M = C // w, 6 * (w + (w & 1))
            flops_exec = n_nodes * 15.0 * M ** 3
            flops_min = 6.0 * C * (6 * (w + 1)) ** 2
            roof_red = {"kernel": "k_bcr_run (block cyclic reduction, persistent work-queue kernel; + k_bcr_load / k_bcr_finish)",
                        "bound": "fp64", "achieved": flops_exec / red_t * sol["pcg_launches"] / 1e12 if red_t > 0 else None,
                        "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": (flops_exec / red_t * sol["pcg_launches"] / 1e12 / fp64_peak) if (red_t > 0 and fp64_peak) else None,
                        "executed_flops_per_solve": flops_exec, "banded_cholesky_flops_per_solve": flops_min,
                        "nodes": n_nodes, "node_dim": M, "levels": int(np.ceil(np.log2(max(n_nodes, 2)))) + 1,
                        "kernel_ms": sol["pcg_ms"], "kernel_launches_timed": sol["pcg_launches"],
                        "note": "exact solve; the dependency chain of log2(nodes) levels (one M x M Cholesky + 6 dense M^3 products "
                                "each) bounds it, not the FP64 pipe: the upper levels keep only a few SMs busy"}
        else:
            pcg_gbs = (8.0 * 36 * nnzb * sol["pcg_iterations"] / red_t / 1e9) if red_t > 0 else None
            roof_red = {"kernel": "k_bpcg_persistent (6x6 block PCG, one cooperative launch per solve)", "bound": "hbm",
                        "achieved": pcg_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": (pcg_gbs / hbm_peak) if pcg_gbs else None,
                        "traffic": measured_traffic(args.workload, "k_bpcg_persistent") if (world == 1 and args.solver == "two_level") else None,
                        "peak_source": peak_src, "algorithmic_bytes_per_iteration": 8.0 * 36 * nnzb, "iterations": sol["pcg_iterations"],
                        "kernel_ms": sol["pcg_ms"], "kernel_launches_timed": sol["pcg_launches"],
                        "us_per_iteration": (1e3 * sol["pcg_ms"] * sol["pcg_launches"] / sol["pcg_iterations"]) if sol["pcg_iterations"] else None,
                        "coarse_refresh_ms": sol["coarse_ms"], "coarse_refreshes": sol["coarse_refreshes"]}
        share = {"linearize_and_schur": lin_ms * lin_n / ms if ms > 0 else None,
                 "reduced_solve": sol["pcg_ms"] * sol["pcg_launches"] / ms if ms > 0 else None,
                 "coarse_refresh": sol["coarse_ms"] * sol["coarse_refreshes"] / ms if ms > 0 else None}
        dominant_is_lin = (share["linearize_and_schur"] or 0) >= (share["reduced_solve"] or 0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": int(st.iterations),
            "warmup": int(args.warmup), "ms_per_step": ms / max(st.iterations, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "lm_iters_per_sec": st.iterations / (ms * 1e-3),
            # SURVEY 8(d) metric (i): E x linearisations / time inside the linearise + JtWJ + Schur kernel alone
            "linearise_edges_per_sec": (E_local * world) / (lin_ms * 1e-3) if lin_ms > 0 else None,
            "config": {"workload": args.workload, **wl, "edges": E, "landmarks": L, "cameras": C,
                       "lm_flavour": "v17", "reduced_solver": solver_used, "reduced_solver_requested": args.solver,
                       "schedule": "Solve(K) from the perturbed initial state (natural LM damping schedule)",
                       "parallelism": f"landmark_shard{world}" + ("+distributed_reduced_solve" if (world > 1 and st.solver_used == capi.SOLVER_BCR and os.environ.get("VIO_B200_SHARD_LEGACY") is None) else ""), "collective": ((("nvlink_peer_memory_kernel+nccl" if p.p2p_enabled() else "nccl") if args.collective == "native" else "hook") if world > 1 else None), "l2": "inputs (>=520 MB of edge records) larger than L2",
                       "scene_gen_s": round(t_gen, 2), "pack_upload_s": round(t_pack, 2)},
            "lm": {"trial_steps": int(st.trial_steps), "accepted": int(st.accepted_steps),
                   "linearizations": int(st.linearizations), "pcg_iterations": int(st.pcg_iterations),
                   "chi2_start": st.chi2_trace[0] if st.n_trace else None, "chi2_final": st.chi2_final,
                   "warmup_chi2_initial": st_w.chi2_initial},
            # the kernel that dominates the step (kernel_share_of_step)
            "roofline": roof_lin if dominant_is_lin else roof_red,
            "roofline_linearize": roof_lin,
            "roofline_reduced_solve": roof_red,
            "kernel_share_of_step": share,
            "e2e": {"value": E * e2e_iters / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int((pose_np.nbytes + invd_np.nbytes) // max(1, args.steps)),
                    "d2h_bytes_per_step": int((pose_np.nbytes + invd_np.nbytes) // max(1, args.steps)), "steps": int(e2e_iters),
                    "calls": e2e_calls, "h2d_bytes_per_call": int(pose_np.nbytes + invd_np.nbytes),
                    "d2h_bytes_per_call": int(pose_np.nbytes + invd_np.nbytes),
                    "call": "graph resident; per call: vio_set_vertices(pinned host state) -> vio_solve(K) -> vio_get_vertices(host); "
                            "bytes_per_step = bytes_per_call / K"},
            # one complete Problem::Solve(K) on a FRESH problem: edge records from host arrays every call
            "dropin": {"value": E * st_d.iterations / t_dropin, "unit": UNIT, "seconds_per_call": t_dropin,
                       "set_graph_s": t_sg, "steps": int(st_d.iterations),
                       "h2d_bytes_per_call": int(52 * E + 24 * L + 56 * C),
                       "call": "vio_set_graph(host edge arrays: pack + upload) -> vio_solve(K) -> vio_get_vertices(host)"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary() if sampler else None,
        }
        if parity is not None:
            line["parity_vs_n1"] = parity
        if world == 1 and not args.no_cpu:
            cb = cpu_reference_rate(vio, scene, steps=3, warmup=1, target_s=12.0)
            line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"],
                                    "sample": cb["sample"], "ms_per_step": cb["ms_per_step"], "breakdown_s": cb.get("breakdown_s")}
            line["host"] = host_info()
        emit(line)
        if parity is not None and not parity["ok"]:
            sys.stderr.write("parity_vs_n1 FAILED: %s\n" % json.dumps(parity))
            if dist is not None:
                dist.barrier()
                dist.destroy_process_group()
            sys.exit(3)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The ONE JSON line of this run, on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)



def host_info():
    model = None
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    model = ln.split(":", 1)[1].strip()
                    break
    except Exception:
        pass
    return {"nproc": os.cpu_count(), "cpu_model": model, "affinity": sorted(os.sched_getaffinity(0))[:4] + (["..."] if len(os.sched_getaffinity(0)) > 4 else [])}


def _gen_window(seed):
    from tests.scenes_extra import window_scene
    return window_scene(seed=seed).export()


def small_scene(vio, wl, n_batch):
    """Scenes of BASELINE configs 1-3: TestMonoBA draw for draw; the VINS-style window built through the
    unmodified reference's IntegrationBase + Marginalize (oracle/_ref/libref17.so, which travels with the repo)."""
    if wl["kind"] == "monoba":
        return [vio.scenes.monoba(20, 300, with_ext=(wl["ver"] == 17))]
    from tests.scenes_extra import window_scene
    if wl["kind"] == "window":
        return [window_scene(seed=2)]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 32)) as pool:
        dicts = pool.map(_gen_window, range(2, 2 + n_batch), chunksize=8)
    return [vio.Scene.from_dict(d) for d in dicts]


def run_small(args):
    """Configs 1-3: one (or 4096) sliding-window sized problems.  A step = one LM iteration of Problem::Solve."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # replicas only: these problems are too small to shard
    vio = importlib.import_module(PKG)
    capi = vio.capi
    wl = SMALL_WORKLOADS[args.workload]
    ver = wl["ver"]
    n_batch = (args.batch_n or wl.get("n", 1)) if wl["kind"] == "batch" else 1
    t0 = time.perf_counter()
    scenes = small_scene(vio, wl, n_batch)
    t_gen = time.perf_counter() - t0
    s0 = scenes[0]
    E_all = int(sum(sc.rp_landmark.shape[0] for sc in scenes))
    K = int(args.steps)
    flav = capi.LM_V15 if ver == 15 else capi.LM_V17
    opts = vio.make_opts(flavour=flav, fixed_iterations=1)
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": 1, "warmup": int(args.warmup), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "problems": len(scenes), "edges": E_all, "P": int(s0.P),
                       "landmarks_per_problem": int(s0.inv_depth.shape[0]), "lm_flavour": f"v{ver}", "scene_gen_s": round(t_gen, 2),
                       "l2": "problem resident in L2 (single sliding-window sized graph: the reference's own operating point)",
                       "parallelism": "single_problem" if len(scenes) == 1 else "lockstep_batch"},
            "host": host_info()}
    if args.impl == "reference":
        # the UNMODIFIED reference backend (oracle/_ref) on the same scene(s), one core
        from tests import refshim
        if not refshim.available(ver):
            emit({"impl": "reference", "unavailable": "oracle/_ref/libref%d.so not built" % ver})
            return
        sample = scenes[:max(1, min(len(scenes), 4))]
        for _ in range(min(args.warmup, 1)):
            refshim.solve(ver, sample[0], K)
        t0 = time.perf_counter()
        its = 0
        for sc in sample:
            its += refshim.solve(ver, sc, K)["iterations"]
        dt = time.perf_counter() - t0
        E_s = sum(sc.rp_landmark.shape[0] for sc in sample) / len(sample)
        val = E_s * its / dt
        line.update({"impl": "reference", "value": val, "steps": int(its), "ms_per_step": 1e3 * dt / max(its, 1),
                     "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference",
                                      "sample": f"unmodified v{ver} backend::Problem::Solve({K}) on {len(sample)} of the {len(scenes)} problem(s) of this workload"},
                     "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(line)
        return
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(0)
    sampler = ClockSampler(0)
    if wl["kind"] == "batch":
        kw = dict(lockstep=True, max_chunk=0)
        for _ in range(2):
            capi.solve_batched(scenes, K, opts, **kw)   # warm-up at the measured size (the entry keeps its handle + staging)
        sampler.start()
        reps, dts, iters, lat = 3, [], 0, []
        for _ in range(reps):
            outs, dt = capi.solve_batched(scenes, K, opts, **kw)
            dts.append(dt)
            iters = sum(o["stats"].iterations for o in outs)
            lat = np.array([o["stats"].ms_total for o in outs])
        sampler.stop_flag = True
        sampler.join(timeout=2)
        dt = float(np.median(dts))
        e_it = sum(sc.rp_landmark.shape[0] * o["stats"].iterations for sc, o in zip(scenes, outs))
        # device part: the chunks' LM loops (every item reports its chunk's loop time; chunks of 1024 run one after the other)
        dev_s = float(np.sum(np.unique(lat))) * 1e-3
        line.update({"value": e_it / dev_s if dev_s > 0 else None, "steps": int(iters), "ms_per_step": 1e3 * dt / max(iters, 1),
                     "problems_per_sec": len(scenes) / dt, "lm_iters_per_sec": iters / dt,
                     "e2e": {"value": e_it / dt, "unit": UNIT, "seconds_per_call": dt,
                             "h2d_bytes_per_step": int(52 * E_all // max(K, 1)), "d2h_bytes_per_step": int(8 * sum(sc.inv_depth.shape[0] + 16 * sc.pose.shape[0] for sc in scenes) // max(K, 1)),
                             "call": "vio_solve_batched_lockstep(host graphs): pack + H2D + Solve(K) per window + D2H, two-slot pipeline"},
                     "latency_ms": {"p50": float(np.percentile(lat, 50)), "p95": float(np.percentile(lat, 95)),
                                    "note": "per item = its chunk's LM loop (all windows of a chunk finish together)"},
                     "value_note": "value = sum(edges x LM iterations) / device time of the chunks' LM loops; e2e = the same over the wall time of the call",
                     "roofline": {"kernel": "k_chol_batch (one CTA per window, 171 x 171 Cholesky in shared memory)", "bound": "fp64", "achieved": None,
                                  "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                                  "note": "latency bound: P column steps x 3 barriers per window; see DESIGN.md section 4"},
                     "gpu_launches": None, "clocks": sampler.summary()})
    else:
        sc = s0
        p = vio.Problem(device=0)
        p.set_graph(sc)
        for _ in range(max(args.warmup, 3)):
            p.set_vertices(pose=sc.pose, speedbias=sc.speedbias if sc.speedbias.size else None, inv_depth=sc.inv_depth)
            if sc.prior is not None:
                p.set_prior(sc.prior["H"], sc.prior["b"], sc.prior.get("err"), sc.prior.get("jt_inv"))
            p.solve(K, opts)
        sampler.start()
        reps, ms_dev, iters, t_e2e = 20, 0.0, 0, 0.0
        launches0 = p.launch_count()
        for _ in range(reps):
            t0 = time.perf_counter()
            p.set_vertices(pose=sc.pose, speedbias=sc.speedbias if sc.speedbias.size else None, inv_depth=sc.inv_depth)
            if sc.prior is not None:
                p.set_prior(sc.prior["H"], sc.prior["b"], sc.prior.get("err"), sc.prior.get("jt_inv"))
            st = p.solve(K, opts)
            p.get_vertices()
            t_e2e += time.perf_counter() - t0
            ms_dev += st.ms_total
            iters += st.iterations
        launches = p.launch_count() - launches0
        t0 = time.perf_counter()
        q = vio.Problem(device=0)
        q.set_graph(sc)
        st_d = q.solve(K, opts)
        q.get_vertices()
        t_drop = time.perf_counter() - t0
        sampler.stop_flag = True
        sampler.join(timeout=2)
        E = int(sc.rp_landmark.shape[0])
        lin_ms, lin_n = p.kernel_ms()
        fp64_peak = capi.measure_fp64_peak(0)
        hbm_peak, peak_src = measured_peaks()
        flops = E * (400.0 + 416.0) + 2.0 * sc.inv_depth.shape[0] * sc.P * sc.P  # linearise + JtWJ per edge, Schur per landmark (dense S)
        line.update({"value": E * iters / (ms_dev * 1e-3), "steps": int(iters // reps), "ms_per_step": ms_dev / max(iters, 1),
                     "lm_iters_per_sec": iters / (ms_dev * 1e-3), "solve_ms": ms_dev / reps, "solver": SOLVER_NAMES.get(int(st.solver_used)),
                     "lm": {"chi2_initial": st.chi2_initial, "chi2_final": st.chi2_final, "iterations": int(st.iterations), "trial_steps": int(st.trial_steps)},
                     "e2e": {"value": E * iters / t_e2e, "unit": UNIT, "seconds_per_call": t_e2e / reps,
                             "h2d_bytes_per_step": int(8 * (sc.pose.size + sc.speedbias.size + sc.inv_depth.size) // max(K, 1)),
                             "d2h_bytes_per_step": int(8 * (sc.pose.size + sc.speedbias.size + sc.inv_depth.size) // max(K, 1)),
                             "call": "vio_set_vertices (+ vio_set_prior) -> vio_solve(K) -> vio_get_vertices, graph resident"},
                     "dropin": {"value": E * st_d.iterations / t_drop, "unit": UNIT, "seconds_per_call": t_drop,
                                "call": "vio_create -> vio_set_graph -> vio_solve(K) -> vio_get_vertices"},
                     "roofline": {"kernel": "k_lin_edges + k_schur_groups", "bound": "fp64", "achieved": (flops / (lin_ms * 1e-3) / 1e12) if lin_ms > 0 else None,
                                  "peak": fp64_peak, "unit": "TFLOP/s",
                                  "frac": (flops / (lin_ms * 1e-3) / 1e12 / fp64_peak) if (lin_ms > 0 and fp64_peak) else None, "traffic": None,
                                  "kernel_ms": lin_ms, "hbm_peak": hbm_peak,
                                  "note": "a single window fills a handful of the 148 SMs: launch latency and the one-CTA dense Cholesky bound "
                                          "the step, not a roofline (DESIGN.md section 4)"},
                     "gpu_launches": int(launches // reps), "clocks": sampler.summary()})
    if not args.no_cpu:
        from tests import refshim
        if refshim.available(ver):
            sample = scenes[:1]
            t0 = time.perf_counter()
            its = refshim.solve(ver, sample[0], K)["iterations"]
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample[0].rp_landmark.shape[0] * its / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                                    "ms_per_solve": 1e3 * dt,
                                    "sample": f"unmodified v{ver} backend::Problem::Solve({K}) on problem 0 of this workload (same scene, same iterations)"}
    emit(line)


def main():
    # stdout carries exactly one JSON line: native libraries (NCCL prints its version banner with printf at communicator
    # creation) get stderr as their fd 1 for the whole run, the JSON goes to the saved descriptor
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS) + list(SMALL_WORKLOADS))
    ap.add_argument("--batch-n", type=int, default=0, help="config 3: number of windows (default 4096)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--solver", "--pcg", dest="solver", default="auto", choices=list(SOLVERS),
                    help="reduced solver on the block-sparse S: auto (= block cyclic reduction on a camera ring), bcr, block PCG "
                         "with the two-level or the plain block-Jacobi preconditioner, or the block-sparse Cholesky")
    ap.add_argument("--collective", default="native", choices=["native", "hook"],
                    help="N > 1: native = NCCL called from inside libvio_b200.so (vio_nccl_init); hook = torch.distributed callback")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the unsharded re-solve on rank 0 (parity_vs_n1)")
    ap.add_argument("--pcg-max-iter", type=int, default=0, help="cap PCG iterations (profiling runs only; 0 = 2P like the reference)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload in SMALL_WORKLOADS:
        run_small(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
