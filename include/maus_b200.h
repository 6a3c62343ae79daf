/*
 * maus_b200.h -- C ABI of libmaus_b200.so: the B200 (sm_100a) implementation of ONE hot path of
 * Kier73/Adaptive-Matrix-Solver ("MAUS"): the per-candidate Psi-regularised shifted inverse-iteration step.
 *
 * The reference (AMS = Adaptive_Matrix_Solver_0.1.py) is pure Python and has no FFI; the seams this library
 * sits behind are Python names:
 *   Seam A  InverseIterateSolver.solve                 AMS:39-104   -> maus_solve_shifted (C = 1)
 *   Seam B  the per-candidate loop of MAUS_Solver.evolve AMS:574-576 -> maus_step (all live candidates at once)
 * Each entry point below cites the reference lines it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain C, no exceptions; every function returns 0 on success, a negative MAUS_E_* code on failure;
 *     maus_last_error() returns the text of the last failure of that context.
 *   - complex128 = two consecutive doubles (re, im), exactly numpy's layout.  All host buffers are caller-owned,
 *     C-contiguous; "V[C][n]" means candidate c's vector is the n complex numbers starting at V + 2*n*c.
 *   - one context = one GPU = one host thread (one process per GPU; multi-GPU = candidate sharding with one all-gather per
 *     generation (maus_gather; host logic in adaptive-matrix-solver_b200/dist.py) or the row-sharded operator (maus_rs_*)).
 *   - there is NO CPU fallback: without a CUDA device maus_create fails with MAUS_E_CUDA.
 */
#ifndef MAUS_B200_H
#define MAUS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct maus_ctx maus_ctx;

/* error codes */
#define MAUS_OK            0
#define MAUS_E_ARG        -1   /* bad argument (NULL pointer, size, unsupported combination) */
#define MAUS_E_CUDA       -2   /* CUDA runtime error, text in maus_last_error */
#define MAUS_E_STATE      -3   /* call order (e.g. step before set_dense) */
#define MAUS_E_NOMEM      -4   /* device workspace does not fit */

/* per-candidate status words written by the solve / step entry points (0 = success).  They let the host keep
 * the reference's retry ladder (AMS:43, 98-104) with identical semantics:
 *   ZERO_PIVOT   <-> scipy.linalg.solve raising LinAlgError (exactly singular), AMS:59/98
 *   NONFINITE    <-> "Solution vector not finite" ValueError, AMS:94-95
 *   GMRES_NOCONV <-> gmres info != 0 -> LinAlgError, AMS:90
 *   V_COLLAPSED  <-> ||v|| < 1e-10 before the step, AMS:259 (host re-initialises from its RNG and re-submits)
 *   MIX_COLLAPSED<-> ||(1-a)v + a x|| <= 1e-10 after the mix, AMS:283 (host replaces v from its RNG)
 *   SKIPPED      <-> candidate was masked out of this call */
#define MAUS_ST_OK             0
#define MAUS_ST_ZERO_PIVOT     1
#define MAUS_ST_NONFINITE      2
#define MAUS_ST_GMRES_NOCONV   3
#define MAUS_ST_V_COLLAPSED    4
#define MAUS_ST_MIX_COLLAPSED  5
#define MAUS_ST_SKIPPED        6

/* problem types, AMS:10-13 */
#define MAUS_EIGENVALUE          1
#define MAUS_SOLVE_LINEAR_SYSTEM 2

/* solver methods, AMS:35-36 ('direct_solve' / 'iterative_gmres') */
#define MAUS_METHOD_LU     0
#define MAUS_METHOD_GMRES  1

/* matrix slots: 0 = current_matrix_A of the step (AMS:145), 1 = the candidate's ctor-time problem_matrix used by
 * the residual (AMS:118, 295).  Slot 1 defaults to slot 0 until set. */
#define MAUS_SLOT_CURRENT  0
#define MAUS_SLOT_CTOR     1

/* ---- context ---------------------------------------------------------------------------------------------- */
int         maus_create(maus_ctx** out, int device);
int         maus_destroy(maus_ctx* ctx);
const char* maus_last_error(maus_ctx* ctx);
/* cap on device workspace for the batched LU (bytes); 0 = default (60% of free memory). */
int         maus_set_workspace_limit(maus_ctx* ctx, int64_t bytes);
/* device id, SM count, bytes of device memory currently held by the context */
int         maus_info(maus_ctx* ctx, int* device, int* sm_count, int64_t* bytes_held);

/* page-locked host buffers for the host<->device copies of maus_step / maus_upload_vectors (optional; any host
 * pointer works, pinned memory makes the copies asynchronous and full-speed) */
void*       maus_alloc_pinned(int64_t bytes);
void        maus_free_pinned(void* p);

/* ---- problem upload (replaces the numpy / scipy.sparse objects self.M, self.b of AMS:342-346) --------------- */
/* dense n x n complex128, row-major (numpy C order).  Kept on the device both row-major (matvec) and
 * column-major (LU). */
int maus_set_dense(maus_ctx* ctx, int slot, int64_t n, const double* A_rowmajor);
/* sparse CSC as scipy.sparse.csc_matrix holds it (AMS:358); converted once to CSR on the device. */
int maus_set_csc(maus_ctx* ctx, int slot, int64_t n, int64_t nnz, const int64_t* colptr, const int64_t* rowidx,
                 const double* vals);
/* attach the dense form (row-major, same order) of the SPARSE matrix resident in slot 0: matvecs keep the CSR copy, the batched LU
 * becomes available as the direct-solve fallback of the retry ladder (the reference falls back to SuperLU, AMS:57, 99-102). */
int maus_add_dense_form(maus_ctx* ctx, int64_t n, const double* A_rowmajor);
/* right-hand side b of SOLVE_LINEAR_SYSTEM (AMS:346) */
int maus_set_rhs(maus_ctx* ctx, const double* b);

/* ---- resident candidate vectors --------------------------------------------------------------------------- */
/* copy C candidate vectors (v_k for eigen, x_k for linear; AMS:120) to / from the device-resident population. */
int maus_upload_vectors(maus_ctx* ctx, int64_t C, const double* V);
int maus_download_vectors(maus_ctx* ctx, int64_t C, double* V);
/* `count` resident vectors starting at candidate `first` (e.g. the arg-min-residual eigenvector of a generation) */
int maus_download_vector_range(maus_ctx* ctx, int64_t first, int64_t count, double* V);

/* ---- granular pieces (the host keeps the control flow of AMS:39-104 / 145-331) ---------------------------- */
/* Rayleigh quotient lambda_c = <v,Av>/<v,v> (lambda = 0 when |<v,v>| < 1e-12) and <v,v>, AMS:264-268.
 * V == NULL uses the resident vectors. */
int maus_rq(maus_ctx* ctx, int64_t C, const double* V, double* lambda_out, double* vnorm2_out);

/* x_c = (A - sigma_c I + psi_c I + R_c)^-1 rhs_c for C candidates, AMS:44-59, 61-97 (one attempt of the ladder):
 *   sigma  [C] complex : lambda_k for eigen (AMS:270), 0 for linear (AMS:274)
 *   psi    [C] real    : AMS:44 (its imaginary part is identically 0)
 *   rng_key[C] uint64  : counter-based (Philox4x32-10) key of the dense perturbation R_c of AMS:49; the host
 *                        passes (candidate id, generation, attempt) packed into 64 bits.  Sparse: R = 0 (AMS:47).
 *   method             : MAUS_METHOD_LU (AMS:59) or MAUS_METHOD_GMRES (AMS:61-90, rtol 1e-8, restart 20, maxiter 50)
 *   use_jacobi [C]     : request the Jacobi preconditioner of AMS:64-86 (host passes stuck_counter > 1); the
 *                        finite / |d| > 1e-12 tests are done on the device.  May be NULL.
 *   RHS                : [C][n], or [n] shared when rhs_shared != 0 (AMS:271 / 275); NULL = resident vectors
 *                        (eigen) or the uploaded b (linear, with rhs_shared != 0)
 *   X_out  [C][n]      : may be NULL (result stays on the device for maus_mix_residual)
 *   status_out [C], iters_out [C] (gmres inner iterations; 0 for LU) */
int maus_solve_shifted(maus_ctx* ctx, int64_t C, const double* sigma, const double* psi, const uint64_t* rng_key,
                       int method, const uint8_t* use_jacobi, const double* RHS, int rhs_shared,
                       double* X_out, int32_t* status_out, int32_t* iters_out);

/* debug / parity: same as the LU branch above for ONE candidate but with the perturbation R supplied by the
 * host (n x n complex128 row-major, e.g. drawn from np.random exactly as AMS:49 does). */
int maus_solve_with_R(maus_ctx* ctx, const double* sigma, const double* psi, const double* R_rowmajor,
                      const double* rhs, double* x_out, int32_t* status_out);

/* damped mix + normalise (AMS:280-285) and residual (AMS:295-299) on the resident vectors and the resident
 * result of the last maus_solve_shifted:
 *   eigen : v <- (1-a) v + a x ; nv = ||v|| ; v <- v/nv if nv > 1e-10 ; r = || A_res v - lambda_old v ||
 *   linear: x <- (1-a) x + a x_new ; r = || A_res x - b ||
 * skip[c] != 0 leaves candidate c untouched (failed solve; status SKIPPED).  res_slot selects A_res. */
int maus_mix_residual(maus_ctx* ctx, int64_t C, int problem_type, const double* alpha, const double* lambda_old,
                      const uint8_t* skip, int res_slot, double* V_out, double* resid_out, double* mixnorm_out,
                      int32_t* status_out);

/* residual only (AMS:295-299) for the resident vectors -- used after the host replaced collapsed vectors. */
int maus_residual(maus_ctx* ctx, int64_t C, int problem_type, const double* V, const double* lambda,
                  int res_slot, double* resid_out);

/* ---- fused fast path: one generation of AMS:574-576 for C candidates, all attempts = 0 -------------------- */
/* rq -> solve -> mix -> normalise -> residual.  V_io == NULL keeps everything resident (no H2D / D2H of vectors).
 * Candidates whose status != 0 are left unchanged (the host runs the ladder for them with the granular calls). */
int maus_step(maus_ctx* ctx, int64_t C, int problem_type, int method, double* V_io, const double* alpha,
              const double* psi, const uint64_t* rng_key, const uint8_t* use_jacobi, int res_slot,
              double* lambda_out, double* resid_out, double* mixnorm_out, int32_t* status_out, int32_t* iters_out);

/* ---- SVD power-sweep branch (AMS:227-255 sweep, AMS:300-301 residual) ----------------------- */
/* rectangular rows x cols complex128 matrix, row-major (the SVD problem matrix of AMS:343) */
int maus_svd_set_matrix(maus_ctx* ctx, int64_t rows, int64_t cols, const double* A_rowmajor);
/* one sweep for C candidates: u <- A v / ||A v||, v <- A^H u / ||A^H u||, sigma = max of the two norms, residual =
 * ||A v - sigma u|| + ||A^H u - sigma v||.  U_io [C][rows], V_io [C][cols] are updated in place.
 * status: V_COLLAPSED = ||v|| < 1e-10 on entry (AMS:229), MIX_COLLAPSED = ||u|| < 1e-10 after the first half sweep
 * (AMS:236); such candidates are left for the host's exception branch (AMS:249-255). */
int maus_svd_step(maus_ctx* ctx, int64_t C, double* U_io, double* V_io, double* sigma_out, double* resid_out,
                  int32_t* status_out);
/* residual only (AMS:301) for host-replaced (u, v, sigma) */
int maus_svd_residual(maus_ctx* ctx, int64_t C, const double* U, const double* V, const double* sigma, double* resid_out);

/* ---- dedup / pruning similarity ------------------------------------------------------------------ */
/* G[i][j] = <v_i, v_j> = np.vdot(v_i, v_j) for C host vectors V [C][n] (complex128); G_out [C][C] row-major complex128.
 * Replaces the O(C^2) host vdots of the reference's converged-solution dedup and survivor pruning
 * (Adaptive_Matrix_Solver_0.1.py:436, 450, 515, 520) by one device pass. */
int maus_gram(maus_ctx* ctx, int64_t C, int64_t n, const double* V, double* G_out);

/* ---- set-up diagnostics (MAUS_Solver._diagnose_matrix_initial, Adaptive_Matrix_Solver_0.1.py:374-404) ----- */
/* On the dense matrix resident in slot 0: number of non-zero entries (np.count_nonzero, AMS:381) and the two np.allclose tests
 * of AMS:384-385 -- is_hermitian = all(isclose(A, A^H)), is_complex_symmetric = all(isclose(A, A^T)) with numpy's element
 * test |a - b| <= atol + rtol |b| (numpy defaults: rtol 1e-5, atol 1e-8). */
int maus_diag_dense(maus_ctx* ctx, double rtol, double atol, int64_t* nonzeros, int32_t* is_hermitian,
                    int32_t* is_complex_symmetric);
/* 2-norm condition number estimate replacing np.linalg.cond (AMS:400, a full SVD): sigma_max by `power_iters` power sweeps on
 * A^H A (HBM-bound matvecs), sigma_min by `inverse_iters` inverse sweeps through the batched LU (A^H y = x, A z = y).  Both are
 * one-sided (sigma_max from below, sigma_min from above): sigma_max / sigma_min <= cond_2(A).  sigma_min = 0 and lu_status != 0
 * when the factorisation met a zero pivot / non-finite solve (numerically singular).  `start`: optional start vector [n]
 * complex128.  Uses the population buffers as scratch: call it before the candidate vectors are uploaded. */
int maus_cond2_estimate(maus_ctx* ctx, int power_iters, int inverse_iters, const double* start, double* sigma_max,
                        double* sigma_min, int32_t* lu_status);

/* ---- Hermitian shortcut --------------------------------------------------------------------------- */
/* Eigendecomposition of a dense Hermitian matrix (replaces scipy.linalg.eigh -> LAPACK zheevd at
 * Adaptive_Matrix_Solver_0.1.py:161): cyclic two-sided Jacobi, round-robin parallel ordering, on the device.  A_rowmajor [n][n]
 * complex128 (the LOWER triangle is used, like eigh's default); w_out [n] ascending; E_rowmajor_out [n][n] (column j = unit
 * eigenvector of w_out[j], numpy layout; may be NULL).  max_sweeps <= 0 = 30.  off_ratio_out = ||offdiag||_F / ||A||_F reached.
 * Returns MAUS_E_STATE when the sweeps did not converge (the caller then behaves as AMS:182-185: "Falling back"). */
int maus_heev(maus_ctx* ctx, int64_t n, const double* A_rowmajor, int max_sweeps, double* w_out, double* E_rowmajor_out,
              int32_t* sweeps_out, double* off_ratio_out);
/* P[c][i] = <e_i, v_c> for the m eigenvectors E = [e_0 .. e_{m-1}] of sla.eigh and C candidate vectors: the similarity scores
 * |v^H E| of Adaptive_Matrix_Solver_0.1.py:165 for the whole population as one tensor-pipe GEMM.  Ec = conj(E) in C order
 * ([n][m] complex128), V [C][n], P_out [C][m]. */
int maus_project(maus_ctx* ctx, int64_t n, int64_t m, const double* Ec, int64_t C, const double* V, double* P_out);

/* ---- multi-GPU: NCCL communicator of the context + row-sharded sparse operator (BASELINE config 5) -------- */
/* One process per GPU.  rank 0 creates the id, the host layer broadcasts the 128 bytes (torch.distributed or anything else);
 * NCCL is bound at run time from `libpath` = the libnccl.so.2 the process already uses (e.g. torch's), so libmaus_b200.so has
 * no link-time dependency on it. */
int maus_nccl_unique_id(const char* libpath, char* out128);
int maus_dist_init(maus_ctx* ctx, const char* libpath, int rank, int world, const char* id128);
/* rank / world of the context; peer_memory = 1 when the row-sharded operator runs on NVLink peer memory (cudaIpc-mapped
 * symmetric segments: fused pack + all-gather stores, one-kernel reductions), 0 when it fell back to NCCL calls */
int maus_dist_info(maus_ctx* ctx, int* rank, int* world, int* peer_memory);
/* candidate-sharded mode: the per-generation all-gather of SURVEY.md 8e (candidate records + vectors, or energies + the best
 * eigenpair): `count` doubles from every rank, recv_all [world][count]; NCCL all-gather on the context's stream. */
int maus_gather(maus_ctx* ctx, const double* send, int64_t count, double* recv_all);

/* Row-sharded operator: every rank owns n / world consecutive rows of A (CSR slice, GLOBAL column indices) and the same slice
 * of every vector.  rowptr has nrows + 1 entries (any base); rowptr / colidx / vals point at the START of the full arrays'
 * slice, i.e. rowptr[i] indexes colidx / vals directly. */
int maus_set_csr_rowblock(maus_ctx* ctx, int64_t n, int64_t row0, int64_t nrows, const int64_t* rowptr,
                          const int64_t* colidx, const double* vals);
/* right-hand side of SOLVE_LINEAR_SYSTEM for the row-sharded operator (full vector; the rank keeps its slice) */
int maus_rs_set_rhs(maus_ctx* ctx, const double* b_full);
/* Y_local[c] = (A V[c])(local rows); V_local, Y_local: [C][nrows] */
int maus_rs_matvec(maus_ctx* ctx, int64_t C, const double* V_local, double* Y_local);
/* x_c = (A - sigma_c I + psi_c I)^-1 rhs_c with the batched GMRES of maus_solve_shifted on the row-sharded operator;
 * RHS_local / X_local_out are the local slices [C][nrows]; status / iters are identical on every rank */
int maus_rs_gmres(maus_ctx* ctx, int64_t C, const double* sigma, const double* psi, const uint8_t* use_jacobi,
                  const double* RHS_local, double* X_local_out, int32_t* status_out, int32_t* iters_out);
/* maus_step on the row-sharded operator (Seam B for config 5 as worded): every rank passes the SAME C candidates; V_full_io
 * [C][n] full-length host vectors (v_k / x_k), updated in place on every rank.  phases: 1 Rayleigh quotient (AMS:264-270) |
 * 2 solve (AMS:44-97, GMRES; sparse: R = 0) | 4 mix + normalise (AMS:280-285) | 8 residual (AMS:295-299); 15 = one generation,
 * 14 with sigma_in = one attempt of the retry ladder (AMS:98-103) for a candidate whose first try failed, 8 = residual only
 * (after the host replaced a vector).  Without phase 1, sigma_in [C] complex supplies the shifts = stale lambdas (eigen). */
#define MAUS_RS_PHASE_RQ       1
#define MAUS_RS_PHASE_SOLVE    2
#define MAUS_RS_PHASE_MIX      4
#define MAUS_RS_PHASE_RESIDUAL 8
int maus_rs_step(maus_ctx* ctx, int64_t C, int problem_type, int phases, double* V_full_io, const double* alpha,
                 const double* psi, const uint8_t* use_jacobi, const double* sigma_in, double* lambda_out, double* resid_out,
                 double* mixnorm_out, int32_t* status_out, int32_t* iters_out);

/* debug / parity: C = beta*C + s*A*B on column-major complex128 host matrices (A: M x K, B: K x N, C: M x N,
 * `batch` of each, densely packed) through the tensor-pipe kernel (use_dmma = 1; 2 = the three-real-product
 * variant used by the LU trailing updates; 3 = its 128 x 32 tile shape used for skinny batched A*V) or the plain FP64-FMA
 * kernel (use_dmma = 0). */
int maus_debug_zgemm(maus_ctx* ctx, int M, int N, int K, int batch, const double* A, const double* B, double* C,
                     int beta, int negate, int use_dmma);

/* ---- instrumentation --------------------------------------------------------------------------------------- */
/* number of kernels this context launched since creation (bench.py's gpu_launches) */
int64_t maus_launch_count(maus_ctx* ctx);
/* device-time (ms, CUDA events on the context's stream) and launch count of the LU trailing-update kernel and
 * the batched matvec kernel accumulated since the last reset -- the roofline numbers of bench.py */
int maus_profile_reset(maus_ctx* ctx, int enable);
int maus_profile_read(maus_ctx* ctx, double* lu_gemm_ms, int64_t* lu_gemm_launches, double* lu_gemm_flops,
                      double* matvec_ms, int64_t* matvec_launches, double* matvec_bytes);
/* per-kernel-family accumulators (device ms, launches, flops or bytes) for the step breakdown */
#define MAUS_PROF_LU_GEMM     0   /* zgemm_dmma_kernel inside the LU (flops) */
#define MAUS_PROF_MATVEC      1   /* gemv_rowmajor_kernel / csr_spmm_kernel (bytes) */
#define MAUS_PROF_PANEL       2   /* lu_panel_kernel (flops) */
#define MAUS_PROF_TRTRI       3
#define MAUS_PROF_BACKSOLVE   4   /* lu_backsolve_kernel (bytes) */
#define MAUS_PROF_BUILD       5   /* lu_build_aug_kernel (bytes) */
#define MAUS_PROF_PERMUTE     6
#define MAUS_PROF_MATVEC_GEMM 7   /* zgemm_dmma_kernel as batched A*V (flops) */
#define MAUS_PROF_VEC         8   /* rq_finish / mix_normalise / residual_finish families (algorithmic bytes) */
#define MAUS_PROF_KINDS       9
int maus_profile_read_kind(maus_ctx* ctx, int kind, double* ms, int64_t* launches, double* work);
/* the context's CUDA stream as a void* (cudaStream_t) so a host layer can order its own work after it */
void* maus_stream(maus_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MAUS_B200_H */
