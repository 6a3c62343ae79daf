#!/usr/bin/env python
"""bench.py -- candidate inverse-iteration steps/s (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]              the CUDA path (this repo)
  python bench.py --impl reference [...]                            the reference's CPU algorithm (oracle port)
  torchrun --nproc-per-node N ... bench.py --gpus N ...             one rank per GPU (driver launches it this way)

A "step" = one generation of the hot path (AMS:574-576) over the rank's candidate shard: Rayleigh quotient, fused
H build, batched LU solve, mix + normalise, residual, the alpha/state scalars, and (N > 1) the per-generation
all-gather.  Workload = K3-c128: dense complex128 n = 4096, 128 live candidates PER GPU (weak scaling), frozen
population (no spawn / prune).  value = candidate-steps/s with the vectors resident in HBM;
e2e = the same through step_population() on host candidate objects (H2D + D2H of every vector every step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DEFAULT, C_DEFAULT = 4096, 128
FP64_PEAK_TFLOPS = 37.1      # measured on this pool's B200: register-resident DMMA loop, profiles/fp64_peak_r01.txt
METRIC = "candidate inverse-iter steps/s at n=4096 c128"


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 8:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def lu_gemm_algorithmic(n, cand, nb=128, group=4, leaf=64):
    """(launches, flops, bytes) of the GEMM launches of ONE batched LU generation, following the schedule of
    maus_lu_solve (csrc/maus_api.cu): per launch bytes = C read (beta = 1) + C write + A + B, complex128."""
    launches, flops, byts = 0, 0.0, 0.0

    def g(M, N, K, beta):
        nonlocal launches, flops, byts
        if M <= 0 or N <= 0:
            return
        launches += 1
        flops += 8.0 * M * N * K * cand
        byts += 16.0 * cand * (M * N * (1 + beta) + M * K + K * N)
    k0 = 0
    while k0 < n:
        kend = min(n, k0 + group * nb)
        nc_out = n + 1 - kend
        kp = k0
        while kp < kend:
            jb = min(nb, kend - kp)

            def block(kb, w):                      # recursive halving of the panel down to `leaf`-wide cluster panels
                if w <= leaf:
                    return
                wl = ((w // 2 + leaf - 1) // leaf) * leaf
                wr, km = w - wl, kb + wl
                block(kb, wl)
                g(wl, wr, wl, 0)
                if n - km > 0:
                    g(n - km, wr, wl, 1)
                block(km, wr)
            block(kp, jb)
            kq = kp + jb
            if kp > k0 and nc_out > 0:
                g(jb, nc_out, kp - k0, 1)
            g(jb, n + 1 - kq, jb, 0)
            if n - kq > 0 and kend - kq > 0:
                g(n - kq, kend - kq, jb, 1)
            kp = kq
        if n - kend > 0 and nc_out > 0:
            g(n - kend, nc_out, kend - k0, 1)
        k0 = kend
    return launches, flops, byts


# DRAM traffic of the GEMM launches measured with ncu (dram__bytes_read.sum + dram__bytes_write.sum over the 151 LU launches
# + 2 batched A*V launches of one generation at n = 4096 with 16 candidates: profiles/launches_r01_final.csv), per candidate
NCU_GEMM_DRAM_BYTES_PER_CANDIDATE = ((27.326e9 + 12.371e9) / 16.0 if os.environ.get("MAUS_GEMM_3M", "1") != "0"
                                     else (33.565e9 + 13.110e9) / 16.0)
NCU_GEMM_TRAFFIC_SOURCE = ("profiles/launches_r01_3m.csv" if os.environ.get("MAUS_GEMM_3M", "1") != "0"
                           else "profiles/launches_r01_final.csv")


def vector_alpha_update(alpha, resid, prev):
    """AMS:306-316 vectorised over the shard (frozen population: the convergence stop AMS:318-331 is not applied)."""
    a = alpha.copy()
    act = prev > 1e-10
    good = act & (resid < prev * 0.9)
    bad = act & ~good & (resid > prev * 1.5) & (prev > 1e-5)
    rest = act & ~good & ~bad
    a[good] = np.minimum(a[good] * 1.1, 1.0)
    a[bad] = np.maximum(a[bad] * 0.5, 1e-6)
    a[rest] = np.maximum(a[rest] * 0.95, 1e-6)
    return a


# =====================================================================================================================
def run_b200(args):
    import torch
    import adaptive_matrix_solver_b200 as pkg
    from adaptive_matrix_solver_b200 import _abi
    from adaptive_matrix_solver_b200.dist import Shard, gather_energy_and_best
    from adaptive_matrix_solver_b200.workloads import k2_matrix, initial_vectors
    from adaptive_matrix_solver_b200.constants import PSI_EPSILON_BASE, psi_magnitude

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libmaus_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    shard = Shard(rank, world, dev if world > 1 else None)

    n, C_ = args.n, args.candidates
    A = k2_matrix(n, seed=20260)
    V0 = initial_vectors(C_, n, seed=20260 + 1000 * rank)
    eng = pkg.MausEngine(local)
    eng.set_matrix(A)
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    base_psi = PSI_EPSILON_BASE * 1.0
    psi0 = np.full(C_, complex(psi_magnitude(base_psi, 0, 0)).real)
    ids = np.arange(C_, dtype=np.uint64) + np.uint64(rank * C_)

    # ------------------------------------------------------------------ resident leg ("value")
    eng.upload_vectors(V0)
    state = {"alpha": np.full(C_, 0.01), "prev": np.full(C_, np.inf), "gen": 0}

    def resident_step():
        g = state["gen"]
        keys = (ids << np.uint64(32)) | np.uint64((g & 0xffffff) << 8)
        out = eng.step(_abi.EIGENVALUE, state["alpha"], psi0, V=None, rng_key=keys, method=_abi.METHOD_LU)
        state["alpha"] = vector_alpha_update(state["alpha"], out["resid"], state["prev"])
        state["prev"] = out["resid"].copy()
        state["gen"] = g + 1
        if world > 1:
            # per-generation exchange: candidate energies + the arg-min-residual eigenpair
            k = int(np.argmin(out["resid"]))
            best = eng.download_vector_range(k, 1)
            state["global"] = gather_energy_and_best(shard, out["resid"], out["lam"], best, best_index=0)
        return out

    for _ in range(args.warmup):
        resident_step()
    shard.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.profile_reset(True)
    l0 = eng.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = resident_step()
    e1.record(stream)
    shard.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    launches = eng.launches - l0
    prof = eng.profile_read()
    breakdown = eng.profile_breakdown()
    eng.profile_reset(False)
    clocks = sampler.stop() if rank == 0 else None
    # device time never exceeds host wall here (every step ends with a stream sync); report the max over ranks
    elapsed_s = shard.all_reduce_max(max(dev_ms / 1e3, 0.0))
    value = world * C_ * args.steps / elapsed_s

    # ------------------------------------------------------------------ e2e leg (host candidate objects)
    np.random.seed(20260 + rank)
    cands = [pkg.Candidate(A, pkg.ProblemType.EIGENVALUE, n, initial_lambda=0j, initial_v=V0[i].copy()) for i in range(C_)]
    strat = dict(overall_psi_aggression_factor=1.0, max_psi_retries=25, current_convergence_threshold=0.0)
    know = dict(local_solver_preference="direct_solve", is_sparse_problem=False, is_hermitian=False)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(args.warmup, 2)):
        pkg.step_population(cands, A, None, strat, know, eng)
    shard.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pkg.step_population(cands, A, None, strat, know, eng)
    torch.cuda.synchronize()
    e2e_s = shard.all_reduce_max(time.perf_counter() - t0)
    e2e_value = world * C_ * e2e_steps / e2e_s
    vec_bytes = C_ * n * 16

    # ------------------------------------------------------------------ roofline of the dominant kernel
    gemm_s = prof["lu_gemm_ms"] / 1e3
    achieved = prof["lu_gemm_flops"] / gemm_s / 1e12 if gemm_s > 0 else 0.0
    n_launch, alg_flops, alg_bytes = lu_gemm_algorithmic(n, C_)
    per_launch = max(1, prof["lu_gemm_launches"] // args.steps)
    use_3m = os.environ.get("MAUS_GEMM_3M", "1") != "0"
    exec_ratio = 0.75 if use_3m else 1.0          # 3M: 6 instead of 8 real flops per complex multiply-add reach the DMMA pipe
    roofline = {"kernel": ("zgemm3m_dmma_kernel" if use_3m else "zgemm_dmma_kernel") + " (LU trailing update + U12 solve)",
                "bound": "tensor",
                "achieved": round(achieved, 3), "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": round(achieved / FP64_PEAK_TFLOPS, 4),
                "note": ("achieved = ALGORITHMIC flops (8 per complex multiply-add, SURVEY.md 8d) / kernel time; the kernel forms every "
                         "complex product from three real products (3M), so only 0.75 x that rate executes on the FP64 tensor pipe: "
                         "frac can exceed 1, pipe_frac is the hardware utilisation") if use_3m else None,
                "executed_tflops": round(achieved * exec_ratio, 3),
                "pipe_frac": round(achieved * exec_ratio / FP64_PEAK_TFLOPS, 4),
                "traffic": round(NCU_GEMM_DRAM_BYTES_PER_CANDIDATE * C_ / per_launch) if n == 4096 else None,
                "traffic_unit": "bytes per launch (ncu dram read+write, average over the LU GEMM launches of a generation; "
                                "measured at 16 candidates and scaled by the candidate count, " + NCU_GEMM_TRAFFIC_SOURCE + ")",
                "algorithmic_flops_per_launch": round(alg_flops / n_launch), "algorithmic_bytes_per_launch": round(alg_bytes / n_launch),
                "peak_source": "own measurement (FP64 DMMA, profiles/fp64_peak_r01.txt); MEASURED_PEAKS.json has no FP64 entry",
                "share_of_step": round(gemm_s / (dev_ms / 1e3), 4) if dev_ms > 0 else None,
                "launches": prof["lu_gemm_launches"],
                "whole_step_frac_of_fp64_peak": round((8.0 / 3.0 * n ** 3 * C_ * args.steps) / (dev_ms / 1e3) / 1e12
                                                      / FP64_PEAK_TFLOPS, 4)}

    # ------------------------------------------------------------------ CPU baseline (rank 0, bounded sample)
    cpu = cpu_baseline(n, args.cpu_sample) if (rank == 0 and args.cpu_sample > 0) else None

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": "candidate-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(elapsed_s / args.steps * 1e3, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)",
                "data": "synthetic",
                "config": {"workload": f"K3-c128: dense complex128 non-Hermitian eigen, n={n}, {C_} live candidates per GPU, "
                                       f"direct (LU) path, Psi on, frozen population", "n": n, "candidates_per_gpu": C_,
                           "parallelism": f"candidate-sharded x{world}, A replicated",
                           "l2": "per-step working set (candidates x 256 MiB LU workspaces) >> 126 MB L2"},
                "e2e": {"value": round(e2e_value, 3), "unit": "candidate-steps/s", "h2d_bytes_per_step": vec_bytes + 40 * C_,
                        "d2h_bytes_per_step": vec_bytes + 48 * C_, "steps": e2e_steps,
                        "api": "step_population(candidates, M, b, strat_params, problem_knowledge, engine)"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "step_breakdown_ms": {k: round(v["ms"] / args.steps, 2) for k, v in breakdown.items() if v["launches"]},
                "wall_s_timed": round(wall, 3), "min_residual": float(np.min(out["resid"]))}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# =====================================================================================================================
def _oracle_sample(n, n_cand, threads):
    """time n_cand oracle candidate-steps at order n with `threads` BLAS threads; returns seconds"""
    import warnings
    from threadpoolctl import threadpool_limits
    from oracle import maus_oracle as mo
    from adaptive_matrix_solver_b200.workloads import k2_matrix, initial_vectors
    A = _oracle_sample.cache.get(n)
    if A is None:
        A = _oracle_sample.cache[n] = k2_matrix(n, seed=20260)
    V0 = initial_vectors(n_cand, n, seed=20260)
    strat = dict(overall_psi_aggression_factor=1.0, max_psi_retries=25, current_convergence_threshold=0.0)
    know = dict(local_solver_preference="direct_solve", is_sparse_problem=False, is_hermitian=False)
    cands = []
    for i in range(n_cand):
        c = mo.CandState(problem_type=mo.EIGENVALUE, N=n)
        c.v_k = V0[i].copy(); c.lambda_k = 0j
        cands.append(c)
    np.random.seed(1)
    with warnings.catch_warnings(), threadpool_limits(limits=threads):
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        for c in cands:
            mo.candidate_step(c, A, None, strat, know)
        return time.perf_counter() - t0


_oracle_sample.cache = {}


def cpu_baseline(n, n_cand):
    cores = os.cpu_count() or 1
    best = None
    for th in sorted({max(1, cores // 2), cores}):
        s = _oracle_sample(n, n_cand, th)
        v = n_cand / s
        if best is None or v > best[0]:
            best = (v, th)
    import scipy
    return {"value": round(best[0], 4), "unit": "candidate-steps/s", "cores": best[1], "kind": "port",
            "sample": f"{n_cand} oracle candidate-steps (oracle/maus_oracle.py: numpy {np.__version__} / scipy "
                      f"{scipy.__version__} zgesv) at n={n}, best of BLAS threads {{cores/2, cores}}, host has {cores} cores"}


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm for the path.  The reference is a single Python file that
    cannot travel to the GPU box, so this times the oracle port (numpy/scipy, the same LAPACK zgesv the reference
    calls at AMS:59) on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n
    cores = os.cpu_count() or 1
    per_step = max(1, args.cpu_sample // 2)
    for _ in range(min(args.warmup, 1)):
        _oracle_sample(n, 1, cores)
    t = 0.0
    for _ in range(args.steps):
        t += _oracle_sample(n, per_step, cores)
    value = per_step * args.steps / t
    import scipy
    sample = (f"each step = {per_step} oracle candidate-steps at n={n} (numpy {np.__version__} / scipy {scipy.__version__}), "
              f"{cores} BLAS threads")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": "candidate-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t / args.steps * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
            "config": {"workload": f"K3-c128: dense complex128 non-Hermitian eigen, n={n}, direct path, bounded sample",
                       "n": n, "candidates_per_gpu": args.candidates},
            "cpu_baseline": {"value": round(value, 4), "unit": "candidate-steps/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": round(value, 4), "unit": "candidate-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--candidates", type=int, default=C_DEFAULT, help="live candidates per GPU")
    ap.add_argument("--cpu-sample", type=int, default=4, help="oracle candidate-steps timed for cpu_baseline (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                       # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
