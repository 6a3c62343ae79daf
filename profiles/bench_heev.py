"""Device Hermitian eigensolver (heev.cu) against the host LAPACK call the reference makes (scipy.linalg.eigh, AMS:161).
python profiles/bench_heev.py [n ...] -> JSON lines"""
import json
import os
import sys
import time

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adaptive_matrix_solver_b200 as pkg                      # noqa: E402

eng = pkg.MausEngine(0)
for n in [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]:
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H = (G + G.conj().T) / np.sqrt(n)
    H[np.arange(n), np.arange(n)] += np.linspace(-3, 3, n)
    eng.heev(H[:64, :64].copy())                                  # module load
    t0 = time.perf_counter(); w, E = eng.heev(H); t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter(); wr, Er = sla.eigh(H); t_cpu = time.perf_counter() - t0
    res = float(np.linalg.norm(H @ E - E * w) / np.linalg.norm(H))
    print(json.dumps(dict(n=n, gpu_s=round(t_gpu, 3), sweeps=eng.heev_info["sweeps"], off_ratio=eng.heev_info["off_ratio"],
                          host_eigh_s=round(t_cpu, 3), host_cores=os.cpu_count(), speedup=round(t_cpu / t_gpu, 2),
                          max_eig_abs_err=float(np.abs(w - wr).max()), rel_residual=res,
                          hbm_bytes_per_sweep=64.0 * n ** 3, achieved_tbs=round(64.0 * n ** 3 * eng.heev_info["sweeps"] / t_gpu / 1e12, 2))),
          flush=True)
eng.close()
