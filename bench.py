#!/usr/bin/env python
"""bench.py -- candidate inverse-iteration steps/s (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]              the CUDA path (this repo)
  python bench.py --impl reference [...]                            the reference's CPU algorithm (oracle port)
  torchrun --nproc-per-node N ... bench.py --gpus N ...             one rank per GPU (driver launches it this way)

A "step" = one generation of the hot path (AMS:574-576) over the rank's candidate shard: Rayleigh quotient, fused
H build, batched LU solve, mix + normalise, residual, the alpha/state scalars, and (N > 1) the per-generation
all-gather.  Workload = K3-c128: dense complex128 n = 4096, 128 live candidates PER GPU (weak scaling), frozen
population (no spawn / prune).  value = candidate-steps/s with the vectors resident in HBM;
e2e = the same through step_population() (N > 1: step_population_sharded(), all-gather included) on host candidate
objects (H2D + D2H of every vector every step).

Outside the timed region the same JSON line also carries (each bounded to a few seconds):
  parity_sample     one candidate of an extra generation re-computed by the oracle (lambda, v up to phase, residual)
  time_to_residual  BASELINE metric, second half: 256 candidates (config 3) stepped with the full alpha / state logic until the
                    first residual < 1e-10, against the oracle's seconds per candidate-step on the host cores
  strong_256        config 3 as worded: 256 candidates in total, 256 / N per GPU
  k2 / k4 / k5      the other BASELINE configurations on one GPU, each with the roofline of its dominant kernel (N = 1 only)
  k5_rowshard       config 5 as worded, matrix row-sharded over the N GPUs (N > 1 only)
"""
import argparse
import hashlib
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
HBM_PEAK_FALLBACK_GBS = 6524.9
METRIC = "candidate inverse-iter steps/s at n=4096 c128"


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def hbm_peak():
    return float(measured_peaks().get("hbm_gbs", HBM_PEAK_FALLBACK_GBS))


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


def gemm_traffic_record():
    """DRAM traffic of the LU GEMM launches as ncu measured it (profiles/ncu_traffic.json, written by
    profiles/summarize_launches.py from the committed launch list).  The record carries the sha of the kernel sources it was
    taken on; a mismatch with the sources of THIS build is reported as stale instead of being passed off as a measurement."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None
    rec["stale"] = rec.get("kernel_sources_sha16") != kernel_sources_sha()
    return rec


def kernel_sources_sha():
    h = hashlib.sha256()
    for f in ("zgemm.cu", "zgemm.cuh", "lu.cu", "lu.cuh"):
        try:
            h.update(open(os.path.join(ROOT, "adaptive-matrix-solver_b200", "csrc", f), "rb").read())
        except Exception:
            pass
    return h.hexdigest()[:16]


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


K3_STRAT = dict(overall_psi_aggression_factor=1.0, max_psi_retries=25, current_convergence_threshold=0.0)
K3_KNOW = dict(local_solver_preference="direct_solve", is_sparse_problem=False, is_hermitian=False)


def guarded(fn):
    """extras never take the headline down: a failure becomes {"error": ...} in the line"""
    def run(*a, **k):
        t0 = time.perf_counter()
        try:
            out = fn(*a, **k)
        except Exception as e:  # noqa: BLE001
            out = {"error": f"{type(e).__name__}: {e}"[:300]}
        if isinstance(out, dict):
            out["wall_s"] = round(time.perf_counter() - t0, 2)
        return out
    return run


# =====================================================================================================================
class ResidentLoop:
    """The benchmark loop on device-resident vectors: fused step, alpha rule, (N > 1) per-generation exchange."""

    def __init__(self, eng, shard, C_, rank, psi0, _abi, gather):
        self.eng, self.shard, self.C, self.abi, self.gather = eng, shard, C_, _abi, gather
        self.ids = np.arange(C_, dtype=np.uint64) + np.uint64(rank * C_)
        self.psi0 = psi0
        self.state = {"alpha": np.full(C_, 0.01), "prev": np.full(C_, np.inf), "gen": 0}

    def step(self):
        st = self.state
        g = st["gen"]
        keys = (self.ids << np.uint64(32)) | np.uint64((g & 0xffffff) << 8)
        out = self.eng.step(self.abi.EIGENVALUE, st["alpha"], self.psi0, V=None, rng_key=keys, method=self.abi.METHOD_LU)
        st["alpha"] = vector_alpha_update(st["alpha"], out["resid"], st["prev"])
        st["prev"] = out["resid"].copy()
        st["gen"] = g + 1
        if self.shard.world > 1:
            # per-generation exchange: candidate energies + the arg-min-residual eigenpair
            k = int(np.argmin(out["resid"]))
            best = self.eng.download_vector_range(k, 1)
            st["global"] = self.gather(self.shard, out["resid"], out["lam"], best, best_index=0)
        return out


def timed_resident(loop, shard, stream, torch, steps, warmup):
    for _ in range(warmup):
        loop.step()
    shard.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    out = None
    for _ in range(steps):
        out = loop.step()
    e1.record(stream)
    shard.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    return shard.all_reduce_max(max(dev_ms / 1e3, 0.0)), wall, dev_ms, out


def run_b200(args):
    import torch
    import adaptive_matrix_solver_b200 as pkg
    from adaptive_matrix_solver_b200 import _abi
    from adaptive_matrix_solver_b200.dist import Shard, gather_energy_and_best, step_population_sharded
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
    n, C_ = args.n, args.candidates
    A = k2_matrix(n, seed=20260)
    V0 = initial_vectors(C_, n, seed=20260 + 1000 * rank)
    eng = pkg.MausEngine(local)
    if world > 1:
        eng.enable_row_sharding(rank, world)      # NCCL communicator of the context: maus_gather + the row-sharded operator
    shard = Shard(rank, world, dev if world > 1 else None, engine=eng if world > 1 else None)
    eng.set_matrix(A)
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    base_psi = PSI_EPSILON_BASE * 1.0
    psi0 = np.full(C_, complex(psi_magnitude(base_psi, 0, 0)).real)

    # ------------------------------------------------------------------ resident leg ("value")
    eng.upload_vectors(V0)
    loop = ResidentLoop(eng, shard, C_, rank, psi0, _abi, gather_energy_and_best)
    for _ in range(args.warmup):
        loop.step()
    shard.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.profile_reset(True)
    l0 = eng.launches
    elapsed_s, wall, dev_ms, out = timed_resident(loop, shard, stream, torch, args.steps, 0)
    launches = eng.launches - l0
    prof = eng.profile_read()
    breakdown = eng.profile_breakdown()
    eng.profile_reset(False)
    clocks = sampler.stop() if rank == 0 else None
    value = world * C_ * args.steps / elapsed_s

    # ------------------------------------------------------------------ parity sample (outside the timed region)
    parity = parity_sample(eng, loop, A, n, rank) if args.extras != "none" else None

    # ------------------------------------------------------------------ e2e leg (host candidate objects)
    # N = 1: step_population on this GPU's candidates.  N > 1: every rank holds the WHOLE population (N x C candidates) and
    # calls the multi-GPU drop-in, step_population_sharded: it steps the candidates it owns and all-gathers record + vector
    # of every candidate (dist.py) -- the collective is inside the timed region.
    np.random.seed(20260)
    Vall = V0 if world == 1 else np.concatenate([initial_vectors(C_, n, seed=20260 + 1000 * r) for r in range(world)])
    cands = [pkg.Candidate(A, pkg.ProblemType.EIGENVALUE, n, initial_lambda=0j, initial_v=Vall[i].copy())
             for i in range(world * C_)]
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_step():
        if world == 1:
            pkg.step_population(cands, A, None, K3_STRAT, K3_KNOW, eng)
        else:
            step_population_sharded(cands, A, None, K3_STRAT, K3_KNOW, eng, shard)
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    shard.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = shard.all_reduce_max(time.perf_counter() - t0)
    e2e_value = world * C_ * e2e_steps / e2e_s
    vec_bytes = C_ * n * 16
    replicas = None
    if world > 1:
        # multi-rank parity observed by the driver: after the sharded generations every rank must hold the SAME population
        # (bit for bit: each candidate is computed by exactly one rank and all-gathered), and rank 0's copy of a candidate that
        # ANOTHER rank stepped must satisfy the step's invariants (unit norm, residual consistent with a host recomputation)
        import hashlib
        h = hashlib.sha256()
        for c in cands:
            h.update(np.ascontiguousarray(c.v_k).tobytes()); h.update(np.complex128(c.lambda_k).tobytes())
            h.update(np.float64(c.residual_k).tobytes())
        dig = float(int.from_bytes(h.digest()[:6], "little"))
        same = shard.all_reduce_max(dig) == -shard.all_reduce_max(-dig)
        other = cands[(rank + 1) % world]                      # owned by the next rank (live candidate i -> rank i mod G)
        r_host = float(np.linalg.norm(A @ other.v_k - other.lambda_k * other.v_k))
        replicas = {"identical_on_all_ranks": bool(same), "foreign_candidate_unit_norm_err": abs(float(np.linalg.norm(other.v_k)) - 1.0),
                    "foreign_candidate_residual": float(other.residual_k), "foreign_candidate_residual_recomputed": r_host}
    del cands

    # ------------------------------------------------------------------ roofline of the dominant kernel
    gemm_s = prof["lu_gemm_ms"] / 1e3
    achieved = prof["lu_gemm_flops"] / gemm_s / 1e12 if gemm_s > 0 else 0.0
    n_launch, alg_flops, alg_bytes = lu_gemm_algorithmic(n, C_)
    per_launch = max(1, prof["lu_gemm_launches"] // args.steps)
    use_3m = os.environ.get("MAUS_GEMM_3M", "1") != "0"
    exec_ratio = 0.75 if use_3m else 1.0          # 3M: 6 instead of 8 real flops per complex multiply-add reach the DMMA pipe
    tr = gemm_traffic_record()
    traffic = None
    if tr and n == 4096 and use_3m and not tr.get("stale"):
        traffic = round(tr["dram_bytes_per_candidate"] * C_ / per_launch)
    roofline = {"kernel": ("zgemm3m_dmma_kernel" if use_3m else "zgemm_dmma_kernel") + " (LU trailing update + U12 solve)",
                "bound": "tensor",
                "achieved": round(achieved, 3), "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": round(achieved / FP64_PEAK_TFLOPS, 4),
                "note": ("achieved = ALGORITHMIC flops (8 per complex multiply-add, SURVEY.md 8d) / kernel time; the kernel forms every "
                         "complex product from three real products (3M), so only 0.75 x that rate executes on the FP64 tensor pipe: "
                         "frac can exceed 1, pipe_frac is the hardware utilisation") if use_3m else None,
                "executed_tflops": round(achieved * exec_ratio, 3),
                "pipe_frac": round(achieved * exec_ratio / FP64_PEAK_TFLOPS, 4),
                "traffic": traffic,
                "traffic_source": None if tr is None else {k: tr.get(k) for k in ("source", "commit", "candidates_measured",
                                                                                  "kernel_sources_sha16", "stale")},
                "traffic_unit": "bytes per launch (ncu dram read+write, average over the LU GEMM launches of a generation, scaled "
                                "by the candidate count); null when the record was taken on other kernel sources than this build",
                "algorithmic_flops_per_launch": round(alg_flops / n_launch), "algorithmic_bytes_per_launch": round(alg_bytes / n_launch),
                "peak_source": "own measurement (FP64 DMMA, profiles/fp64_peak_r01.txt); MEASURED_PEAKS.json has no FP64 entry",
                "share_of_step": round(gemm_s / (dev_ms / 1e3), 4) if dev_ms > 0 else None,
                "launches": prof["lu_gemm_launches"],
                "whole_step_frac_of_fp64_peak": round((8.0 / 3.0 * n ** 3 * C_ * args.steps) / (dev_ms / 1e3) / 1e12
                                                      / FP64_PEAK_TFLOPS, 4)}

    # ------------------------------------------------------------------ extras (all outside the timed regions above)
    extras = {}
    if args.extras != "none":
        cpu_step_s = None
        if rank == 0 and args.cpu_sample > 0:
            cpu = cpu_baseline(n, args.cpu_sample)
            cpu_step_s = 1.0 / cpu["value"] if cpu and cpu.get("value") else None
        else:
            cpu = None
        extras["strong_256"] = strong_256(pkg, eng, shard, stream, torch, A, n, rank, world, _abi, gather_energy_and_best, psi0[0])
        extras["time_to_residual"] = time_to_residual(pkg, eng, shard, torch, A, n, world, step_population_sharded, cpu, cpu_step_s)
        if world == 1:
            extras["k2"] = bench_k2(pkg, eng, stream, torch, _abi)
            extras["k4"] = bench_k4(pkg, eng, stream, torch, _abi)
            extras["k5"] = bench_k5(pkg, eng, stream, torch, _abi)
        else:
            extras["k5_rowshard"] = bench_k5_rowshard(pkg, eng, shard, torch, _abi, rank, world)
    else:
        cpu = cpu_baseline(n, args.cpu_sample) if (rank == 0 and args.cpu_sample > 0) else None

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": "candidate-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(elapsed_s / args.steps * 1e3, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)",
                "data": "synthetic",
                "config": k3_config(n, C_, world),
                "e2e": {"value": round(e2e_value, 3), "unit": "candidate-steps/s", "h2d_bytes_per_step": vec_bytes + 40 * C_,
                        "d2h_bytes_per_step": vec_bytes + 48 * C_, "steps": e2e_steps,
                        "api": ("step_population(candidates, M, b, strat_params, problem_knowledge, engine)" if world == 1 else
                                "step_population_sharded(candidates, M, b, strat_params, problem_knowledge, engine, shard): "
                                "per-generation all-gather of record + vector of every candidate inside the timed region"),
                        "allgather_bytes_per_rank_per_step": 0 if world == 1 else C_ * (14 + 2 * n) * 8,
                        "sharded_replicas": replicas},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "step_breakdown_ms": {k: round(v["ms"] / args.steps, 2) for k, v in breakdown.items() if v["launches"]},
                "wall_s_timed": round(wall, 3), "min_residual": float(np.min(out["resid"])),
                "parity_sample": parity}
        line.update(extras)
        print(json.dumps(line), flush=True)
    shard.barrier()
    eng.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def k3_config(n, C_, world):
    return {"workload": f"K3-c128: dense complex128 non-Hermitian eigen, n={n}, {C_} live candidates per GPU, "
                        f"direct (LU) path, Psi on, frozen population", "n": n, "candidates_per_gpu": C_,
            "parallelism": f"candidate-sharded x{world}, A replicated",
            "l2": "per-step working set (candidates x 256 MiB LU workspaces) >> 126 MB L2"}


# =====================================================================================================================
@guarded
def parity_sample(eng, loop, A, n, rank):
    """One more generation of the benched loop, with candidate k's state saved before it; the oracle (the reference's
    numpy / LAPACK step, AMS:264-299) recomputes that candidate on the host.  Tolerance = north_star's 1e-10 relative with the
    4e-13 ||A|| floor of tests/parity.py."""
    import warnings
    from oracle import maus_oracle as mo
    k = 3 if loop.C > 3 else 0
    v_before = eng.download_vector_range(k, 1)[0].copy()
    alpha = float(loop.state["alpha"][k]); prev = float(loop.state["prev"][k])
    out = loop.step()
    v_after = eng.download_vector_range(k, 1)[0]
    o = mo.CandState(problem_type=mo.EIGENVALUE, N=n)
    o.v_k = v_before; o.lambda_k = 0j; o.alpha_local_step = alpha; o.residual_k = prev
    prng = np.random.default_rng(1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mo.candidate_step(o, A, None, K3_STRAT, K3_KNOW, rand=lambda *s: prng.random(s))
    floor = 4e-13 * float(np.abs(A).sum(axis=1).max())
    lam_err = abs(complex(out["lam"][k]) - complex(o.lambda_k))
    res_err = abs(float(out["resid"][k]) - float(o.residual_k))
    ph = np.vdot(v_after, o.v_k); ph = ph / abs(ph) if abs(ph) > 0 else 1.0
    vec_err = float(np.abs(v_after * ph - o.v_k).max() / np.abs(o.v_k).max())
    ok = (lam_err <= 1e-10 * abs(o.lambda_k) + floor and res_err <= 1e-10 * float(o.residual_k) + floor and vec_err <= 1e-9)
    return {"status": "ok" if ok else "MISMATCH", "candidate": int(k), "generation": int(loop.state["gen"]),
            "lambda_abs_err": lam_err, "residual_abs_err": res_err, "vector_err_up_to_phase": vec_err,
            "oracle": "oracle/maus_oracle.py candidate_step (numpy / scipy zgesv), same v, alpha, lambda rule",
            "tolerance": f"1e-10 relative + {floor:.2e} (4e-13 ||A||_inf); vectors 1e-9 after phase alignment"}


@guarded
def strong_256(pkg, eng, shard, stream, torch, A, n, rank, world, _abi, gather, psi_value, total=256, steps=2, warmup=1):
    """BASELINE config 3 as worded: 256 candidates in total, 256 / N per GPU, same resident loop as the headline."""
    Cs = total // world
    eng.set_matrix(A)
    from adaptive_matrix_solver_b200.workloads import initial_vectors
    eng.upload_vectors(initial_vectors(Cs, n, seed=777 + rank))
    loop = ResidentLoop(eng, shard, Cs, rank, np.full(Cs, psi_value), _abi, gather)
    elapsed_s, _, _, out = timed_resident(loop, shard, stream, torch, steps, warmup)
    return {"value": round(total * steps / elapsed_s, 3), "unit": "candidate-steps/s", "ms_per_step": round(elapsed_s / steps * 1e3, 3),
            "candidates_total": total, "candidates_per_gpu": Cs, "steps": steps, "scaling": "strong",
            "frac_of_fp64_peak_whole_step": round((8.0 / 3.0 * n ** 3 * Cs * steps) / elapsed_s / 1e12 / FP64_PEAK_TFLOPS, 4)}


@guarded
def time_to_residual(pkg, eng, shard, torch, A, n, world, step_population_sharded, cpu, cpu_step_s, total=256, tol=1e-10,
                     max_gens=14, want_distinct=8):
    """Second half of the BASELINE metric / the north_star target run: n = 4096, 256 candidates over the N GPUs, full
    alpha / state / convergence logic through the drop-in (step_population[_sharded]) until the first residual < 1e-10, and on
    until 8 DISTINCT eigenpairs have converged."""
    from adaptive_matrix_solver_b200.workloads import initial_vectors
    V0 = initial_vectors(total, n, seed=20260)
    strat = dict(K3_STRAT, current_convergence_threshold=tol)
    np.random.seed(1)
    warm = [pkg.Candidate(A, pkg.ProblemType.EIGENVALUE, n, initial_lambda=0j, initial_v=V0[i].copy()) for i in range(2 * world)]
    step_population_sharded(warm, A, None, K3_STRAT, K3_KNOW, eng, shard)
    np.random.seed(1)
    cands = [pkg.Candidate(A, pkg.ProblemType.EIGENVALUE, n, initial_lambda=0j, initial_v=V0[i].copy()) for i in range(total)]
    State = pkg.Candidate.State
    shard.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    gens, first, t_first, g_first = 0, None, None, None
    distinct, t_eight = [], None
    while gens < max_gens and t_eight is None:
        step_population_sharded(cands, A, None, strat, K3_KNOW, eng, shard)
        gens += 1
        conv = [c for c in cands if c.state == State.CONVERGED and c.residual_k < tol]
        if conv and first is None:
            torch.cuda.synchronize()
            first = min(conv, key=lambda c: c.residual_k); t_first = time.perf_counter() - t0; g_first = gens
        # distinct converged eigenpairs by the reference's similarity rule (AMS:435-436): eigenvalues apart or |<v_i, v_j>| <= 0.999
        distinct = []
        for c in conv:
            if all(abs(c.lambda_k - d.lambda_k) > 1e-5 + 1e-6 * abs(d.lambda_k) or abs(np.vdot(c.v_k, d.v_k)) <= 0.999 for d in distinct):
                distinct.append(c)
        if len(distinct) >= want_distinct:
            torch.cuda.synchronize()
            t_eight = time.perf_counter() - t0
    torch.cuda.synchronize()
    gpu_s = shard.all_reduce_max(t_first if t_first is not None else time.perf_counter() - t0)
    out = {"gpu_s": round(gpu_s, 3), "generations": g_first if g_first is not None else gens, "reached": first is not None,
           "candidates_total": total, "candidates_per_gpu": total // world, "tol": tol, "n": n,
           "api": "step_population_sharded" if world > 1 else "step_population",
           "distinct_target": want_distinct, "distinct_reached": len(distinct),
           "gpu_s_to_distinct": None if t_eight is None else round(shard.all_reduce_max(t_eight), 3),
           "generations_to_distinct": gens if t_eight is not None else None,
           "converged_after": sum(c.state == State.CONVERGED for c in cands)}
    if first is not None:
        true_res = float(np.linalg.norm(A @ first.v_k - first.lambda_k * first.v_k))
        out.update({"first_residual": float(first.residual_k), "first_residual_recomputed_on_host": true_res,
                    "first_lambda": [float(np.real(first.lambda_k)), float(np.imag(first.lambda_k))]})
    if distinct:
        out["max_true_residual_of_distinct"] = max(float(np.linalg.norm(A @ c.v_k - c.lambda_k * c.v_k)) for c in distinct)
    gens = out["generations"]
    if cpu_step_s:
        out.update({"cpu_s_per_step": round(cpu_step_s, 3), "cores": cpu.get("cores"),
                    "cpu_extrapolated_s": round(cpu_step_s * gens * total, 1),
                    "cpu_extrapolation": "generations x candidates x measured oracle seconds per candidate-step: the reference steps its "
                                         "population sequentially (AMS:574-576)",
                    "cpu_single_candidate_s": round(cpu_step_s * gens, 2)})
    return out


def _profiled_steps(eng, stream, torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    eng.profile_reset(True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        out = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    bd = eng.profile_breakdown()
    eng.profile_reset(False)
    return e0.elapsed_time(e1) / 1e3, bd, out


@guarded
def bench_k2(pkg, eng, stream, torch, _abi, n=1024, C_=64, steps=10):
    """BASELINE config 2: dense non-Hermitian eigen, n = 1024, 64 candidates, Psi on, direct path, one GPU."""
    from adaptive_matrix_solver_b200.workloads import k2_matrix, initial_vectors
    A = k2_matrix(n, seed=20260 + n)
    eng.set_matrix(A)
    eng.upload_vectors(initial_vectors(C_, n, seed=5))
    alpha, psi = np.full(C_, 0.05), np.full(C_, 1e-20)
    keys = np.arange(C_, dtype=np.uint64) << np.uint64(32)
    s, bd, out = _profiled_steps(eng, stream, torch,
                                 lambda: eng.step(_abi.EIGENVALUE, alpha, psi, V=None, rng_key=keys, method=_abi.METHOD_LU), steps, 3)
    g = bd["lu_gemm"]; p = bd["panel"]
    dom = max(bd.items(), key=lambda kv: kv[1]["ms"])
    gemm_tf = g["work"] / (g["ms"] / 1e3) / 1e12 if g["ms"] > 0 else 0.0
    return {"value": round(C_ * steps / s, 1), "unit": "candidate-steps/s", "ms_per_step": round(s / steps * 1e3, 3),
            "workload": f"K2: dense non-Hermitian eigen n={n}, {C_} candidates, direct path", "dominant_kernel": dom[0],
            "dominant_share": round(dom[1]["ms"] / 1e3 / s, 3),
            "roofline": {"kernel": "zgemm3m_dmma_kernel", "bound": "tensor", "achieved": round(gemm_tf, 2), "peak": FP64_PEAK_TFLOPS,
                         "unit": "TFLOP/s", "frac": round(gemm_tf / FP64_PEAK_TFLOPS, 4),
                         "whole_step_frac_of_fp64_peak": round((8.0 / 3.0 * n ** 3 * C_ * steps) / s / 1e12 / FP64_PEAK_TFLOPS, 4)},
            "breakdown_ms": {k: round(v["ms"] / steps, 3) for k, v in bd.items() if v["launches"]},
            "max_residual": float(np.max(out["resid"])), "panel_ms": round(p["ms"] / steps, 3)}


@guarded
def bench_k4(pkg, eng, stream, torch, _abi, n=8192, C_=128):
    """BASELINE config 4: ill-conditioned dense Ax = b, n = 8192, 128 candidates, Psi on, GMRES(20) x 50 with the Jacobi
    preconditioner for the stuck half (AMS:61-90)."""
    from adaptive_matrix_solver_b200.workloads import k4_system
    A, b = k4_system(n)
    eng.set_matrix(A); eng.set_rhs(b)
    rng = np.random.default_rng(0)
    X0 = rng.standard_normal((C_, n)) + 1j * rng.standard_normal((C_, n))
    alpha = np.full(C_, 0.25)
    psi = np.full(C_, 1e-19)
    jac = (np.arange(C_) % 2).astype(np.uint8)
    keys = np.arange(C_, dtype=np.uint64) << np.uint64(32)

    def one():
        eng.upload_vectors(X0)
        return eng.step(_abi.SOLVE_LINEAR_SYSTEM, alpha, psi, V=None, rng_key=keys, method=_abi.METHOD_GMRES, use_jacobi=jac)
    s, bd, out = _profiled_steps(eng, stream, torch, one, 1, 1)
    mg = bd["matvec_gemm"]
    tf = mg["work"] / (mg["ms"] / 1e3) / 1e12 if mg["ms"] > 0 else 0.0
    it = out["iters"]
    inner_total = int(it.max())
    return {"value": round(C_ / s, 1), "unit": "candidate-steps/s", "ms_per_step": round(s * 1e3, 2),
            "workload": f"K4: dense Ax=b n={n} cond~1e9, {C_} candidates, GMRES(20)x50, Jacobi for half", "status_ok": int((out["status"] == 0).sum()),
            "inner_iters_jacobi": int(it[jac == 1].max()), "inner_iters_plain": int(it[jac == 0].max()),
            "ms_per_inner_iteration": round(s * 1e3 / max(1, inner_total), 3),
            "dominant_kernel": "batched A*V on the DMMA GEMM (matvec_gemm)", "dominant_share": round(mg["ms"] / 1e3 / s, 3),
            "roofline": {"kernel": "zgemm3m_dmma_kernel / zgemm_dmma_kernel as A*[v_1..v_C]", "bound": "tensor", "achieved": round(tf, 2),
                         "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": round(tf / FP64_PEAK_TFLOPS, 4),
                         "launches": mg["launches"]}}


@guarded
def bench_k5(pkg, eng, stream, torch, _abi, n=1_000_000, C_=8):
    """BASELINE config 5 on one GPU (matrix replicated): sparse CSC linear system, n = 1M, ~21 nnz / row, GMRES / SpMM path."""
    from adaptive_matrix_solver_b200.workloads import k5_sparse
    A = k5_sparse(n)
    rng = np.random.default_rng(8)
    b = rng.random(n) + 1j * rng.random(n)
    eng.set_matrix(A); eng.set_rhs(b)
    X0 = rng.random((C_, n)) + 1j * rng.random((C_, n))
    alpha, psi = np.full(C_, 0.5), np.full(C_, 5e-19)

    def one():
        eng.upload_vectors(X0)
        return eng.step(_abi.SOLVE_LINEAR_SYSTEM, alpha, psi, V=None, rng_key=None, method=_abi.METHOD_GMRES)
    s, bd, out = _profiled_steps(eng, stream, torch, one, 1, 1)
    mv = bd["matvec"]
    gbs = mv["work"] / (mv["ms"] / 1e3) / 1e9 if mv["ms"] > 0 else 0.0
    peak = hbm_peak()
    return {"value": round(C_ / s, 1), "unit": "candidate-steps/s", "ms_per_step": round(s * 1e3, 2), "nnz": int(A.nnz),
            "workload": f"K5: sparse CSC Ax=b n={n}, ~{A.nnz // n} nnz/row, {C_} candidates, GMRES(20)x50, matrix replicated",
            "status_ok": int((out["status"] == 0).sum()), "inner_iters": int(out["iters"].max()),
            "dominant_kernel": "csr_spmm (packed gathers)", "dominant_share": round(mv["ms"] / 1e3 / s, 3),
            "roofline": {"kernel": "csr_spmm_packed_kernel", "bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(gbs / peak, 4), "launches": mv["launches"],
                         "bytes": "algorithmic: matrix (nnz x 20 B + rowptr) once per pass of packed candidates + 32 n B per candidate"}}


@guarded
def bench_k5_rowshard(pkg, eng, shard, torch, _abi, rank, world, n=1_000_000, C_per_gpu=8):
    """BASELINE config 5 as worded: the matrix ROW-SHARDED over the N GPUs (rank r owns n / N rows and that slice of every vector),
    N x 8 linear solves done jointly; beside it the default mode (matrix replicated, 8 solves per GPU) on the same box."""
    from adaptive_matrix_solver_b200.workloads import k5_sparse
    A = k5_sparse(n)
    psi = np.full(C_per_gpu, 5e-19); zero = np.zeros(C_per_gpu, dtype=complex)
    rng = np.random.default_rng(100 + rank)
    RHS = rng.random((C_per_gpu, n)) + 1j * rng.random((C_per_gpu, n)); RHS /= np.linalg.norm(RHS, axis=1, keepdims=True)
    eng.set_matrix(A)
    eng.solve_shifted(zero, psi, rng_key=None, method=_abi.METHOD_GMRES, RHS=RHS, want_x=False)
    shard.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    _, st_a, it_a = eng.solve_shifted(zero, psi, rng_key=None, method=_abi.METHOD_GMRES, RHS=RHS, want_x=False)
    dt_a = shard.all_reduce_max(time.perf_counter() - t0)
    op = eng.enable_row_sharding(rank, world)
    op.set_matrix(A)
    Call = C_per_gpu * world
    rng = np.random.default_rng(100)
    Rloc = np.empty((Call, op.nloc), dtype=np.complex128)
    for c in range(Call):                                   # every rank draws the same full vectors, keeps its slice
        v = rng.random(n) + 1j * rng.random(n)
        Rloc[c] = (v / np.linalg.norm(v))[op.row0:op.row0 + op.nloc]
    psi_all = np.full(Call, 5e-19); zero_all = np.zeros(Call, dtype=complex)
    op.gmres(zero_all, psi_all, Rloc, want_x=False)
    shard.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    _, st_b, it_b = op.gmres(zero_all, psi_all, Rloc, want_x=False)      # like the replicated leg: the solutions stay on the device
    dt_b = shard.all_reduce_max(time.perf_counter() - t0)
    Xl, _, _ = op.gmres(zero_all[:1], psi_all[:1], Rloc[:1])             # untimed: candidate 0 again, for the residual check
    # residual of candidate 0 on the local rows (needs the full solution: gathered through the operator's own matvec)
    Yl = op.matvec(Xl[:1])
    r2 = float(np.linalg.norm(Yl[0] - Rloc[0]) ** 2)
    r2 = shard.all_reduce_sum(r2)
    return {"value": round(Call / dt_b, 1), "unit": "candidate-solves/s", "seconds": round(dt_b, 4), "candidates_total": Call,
            "mode": op.mode_description(), "inner_iters": sorted(set(it_b.tolist())), "status_ok": int((st_b == 0).sum()),
            "rel_residual_c0": float(np.sqrt(r2)),
            "replicated": {"value": round(Call / dt_a, 1), "seconds": round(dt_a, 4), "inner_iters": sorted(set(it_a.tolist())),
                           "status_ok_rank0": int((st_a == 0).sum())},
            "rowsharded_over_replicated": round(dt_a / dt_b, 3)}


# =====================================================================================================================
def _oracle_sample(n, n_cand, threads):
    """time n_cand oracle candidate-steps at order n with `threads` BLAS threads (None = library default); returns seconds"""
    import contextlib
    import warnings
    from threadpoolctl import threadpool_limits
    from oracle import maus_oracle as mo
    from adaptive_matrix_solver_b200.workloads import k2_matrix, initial_vectors
    A = _oracle_sample.cache.get(n)
    if A is None:
        A = _oracle_sample.cache[n] = k2_matrix(n, seed=20260)
    V0 = initial_vectors(n_cand, n, seed=20260)
    cands = []
    for i in range(n_cand):
        c = mo.CandState(problem_type=mo.EIGENVALUE, N=n)
        c.v_k = V0[i].copy(); c.lambda_k = 0j
        cands.append(c)
    np.random.seed(1)
    limit = threadpool_limits(limits=threads) if threads else contextlib.nullcontext()
    with warnings.catch_warnings(), limit:
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        for c in cands:
            mo.candidate_step(c, A, None, K3_STRAT, K3_KNOW)
        return time.perf_counter() - t0


_oracle_sample.cache = {}


def _thread_grid(cores):
    return sorted({t for t in (1, 2, 4, 8, cores) if t <= cores})


def cpu_baseline(n, n_cand):
    """The reference's CPU path (oracle port) on the box's host cores, SURVEY.md 8d hygiene: OPENBLAS threads swept over
    {1, 2, 4, 8, all} -- at a reduced order (n = 1024) so that the sweep stays within seconds -- then the K3 sample itself
    (n_cand candidate-steps at order n) with the best count of the sweep, with all cores, and with the library default."""
    import scipy
    from threadpoolctl import threadpool_info
    cores = os.cpu_count() or 1
    sweep_n = min(1024, n)
    sweep = {}
    for th in _thread_grid(cores):
        _oracle_sample(sweep_n, 1, th)
        sweep[str(th)] = round(2 / _oracle_sample(sweep_n, 2, th), 3)
    best_th = int(max(sweep, key=lambda k: sweep[k]))
    runs = {}
    for th in sorted({best_th, cores}):
        runs[str(th)] = round(n_cand / _oracle_sample(n, n_cand, th), 4)
    default = round(n_cand / _oracle_sample(n, n_cand, None), 4)
    best = max(runs.items(), key=lambda kv: kv[1])
    if default > best[1]:
        best = ("default", default)
    info = [{k: d.get(k) for k in ("user_api", "internal_api", "version", "num_threads", "threading_layer")} for d in threadpool_info()]
    return {"value": best[1], "unit": "candidate-steps/s", "cores": cores if best[0] == "default" else int(best[0]), "kind": "port",
            "default_threads_value": default, "by_threads_at_n": runs,
            "thread_sweep": {"n": sweep_n, "candidate_steps_per_s": sweep},
            "host_cores": cores, "threadpool_info": info,
            "sample": f"{n_cand} oracle candidate-steps (oracle/maus_oracle.py: numpy {np.__version__} / scipy "
                      f"{scipy.__version__} zgesv) at n={n} per thread setting; thread counts {{1,2,4,8,all}} swept at n={sweep_n}, "
                      f"best of the sweep + all cores + library default re-timed at n={n}"}


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
              f"{cores} BLAS threads (all host cores)")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": "candidate-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t / args.steps * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
            "config": k3_config(n, args.candidates, max(1, args.gpus)),
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
    ap.add_argument("--cpu-sample", type=int, default=2, help="oracle candidate-steps timed per thread setting for cpu_baseline (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--extras", default="all", choices=["all", "none"],
                    help="none: headline + e2e + cpu_baseline only (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                       # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
