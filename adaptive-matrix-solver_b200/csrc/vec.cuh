// vec.cuh -- fused reduction / elementwise kernels of the candidate step (Rayleigh quotient, mix + normalise,
// residual norms; AMS:264-268, 280-285, 295-299) and layout helpers.
#pragma once
#include "common.cuh"

// out (n x n column-major) = transpose-free re-layout of in (n x n row-major)
cudaError_t vec_rowmajor_to_colmajor(const cplx* in_rm, cplx* out_cm, int n, cudaStream_t stream);

// lambda_c = <v_c, y_c> / <v_c, v_c>  (0 when |<v,v>| < 1e-12), vnorm2_c = <v_c, v_c>; status V_COLLAPSED when
// sqrt(<v,v>) < 1e-10.  V, Y: [C][n].
// `scratch` (vec_scratch_doubles(C) doubles, may be null): with it, vectors of n >= 32768 are reduced by many CTAs per
// candidate (two-level, fixed-order partial sums) instead of one -- the n ~ 1e6 sparse configurations
size_t vec_scratch_doubles(long long C);
// scratch_cap > 0: the candidate capacity the scratch buffer was sized (and zero-initialised) for -- the block-done counters sit
// behind the partials of that many candidates, and the last block of a candidate folds the final reduction into the same launch
cudaError_t vec_rq_finish(const cplx* V, const cplx* Y, int n, int C, cplx* lambda, double* vnorm2, int* status,
                          double* scratch, cudaStream_t stream, int scratch_cap = 0);

// eigen : v <- (1-a) v + a x ; nv = ||v||_2 ; v /= nv when nv > 1e-10 else status MIX_COLLAPSED (v left unnormalised)
// linear: v <- (1-a) v + a x ; nv = ||v||_2 (reported only)
// candidates with status[c] != 0 on entry are left untouched.
cudaError_t vec_mix_normalise(cplx* V, const cplx* X, int n, int C, int problem_type, const double* alpha,
                              double* mixnorm, int* status, double* scratch, cudaStream_t stream);

// eigen : r_c = || y_c - lambda_c v_c ||_2       linear: r_c = || y_c - b ||_2
cudaError_t vec_residual_finish(const cplx* V, const cplx* Y, int n, int C, int problem_type, const cplx* lambda,
                                const cplx* b, double* resid, double* scratch, cudaStream_t stream, int scratch_cap = 0);

// Y[c] = A_rowmajor * V[c]: HBM-bound batched matvec, one warp per matrix row, CB candidates per pass
cudaError_t vec_gemv_rowmajor(const cplx* A_rm, const cplx* V, long long ldv, cplx* Y, long long ldy, int n, int C,
                              cudaStream_t stream);

// diag[i] = A[i][i]; *amax = max_ij (|re| + |im|)  (amax must be zero-initialised)
cudaError_t vec_diag_amax(const cplx* A_rm, int n, cplx* diag, double* amax, cudaStream_t stream);

// Y[c] (nrows) = A (nrows x ncols, row-major) * V[c] (ncols): rectangular variant for the SVD sweep
cudaError_t vec_gemv_rect(const cplx* A_rm, int nrows, int ncols, const cplx* V, long long ldv, cplx* Y, long long ldy, int C,
                          cudaStream_t stream);

// G[i][j] = <v_i, v_j> = sum_k conj(v_i[k]) v_j[k] for C vectors of length n ([C][n]); G is [C][C] row-major
cudaError_t vec_gram(const cplx* V, int n, int C, cplx* G, cudaStream_t stream);

// explicit halves of the multi-block reductions for row-sharded vectors (rowshard.cu).  scratch: [C][VEC_PART_MAXBLK][4] doubles;
// a `part` kernel fills blocks 0 .. nblk-1 of every candidate, the caller combines them over the ranks (component-wise sum or
// max, see rowshard.cu) and the `final` / `apply` kernel reads them back.  The rare scaled-norm recomputation of the final
// kernels (plain sum of squares out of range) only sees the local slice: not supported for sharded vectors.
constexpr int VEC_PART_MAXBLK = 64;
int vec_part_blocks(long long n);
cudaError_t vec_rq_part(const cplx* V, const cplx* Y, int n, int C, double* scratch, int nblk, cudaStream_t stream);          // comps 0,1,2: sums
cudaError_t vec_rq_final(const double* scratch, int nblk, int C, cplx* lambda, double* vnorm2, int* status, cudaStream_t stream);
cudaError_t vec_mix_part(cplx* V, const cplx* X, int n, int C, const double* alpha, const int* status, double* scratch, int nblk,
                         cudaStream_t stream);                                                                                 // comp 0: max (-1 = skipped), 1: sum
cudaError_t vec_mix_apply(cplx* V, int n, int C, int problem_type, double* mixnorm, int* status, const double* scratch, int nblk,
                          cudaStream_t stream);
cudaError_t vec_res_part(const cplx* V, const cplx* Y, int n, int C, int problem_type, const cplx* lambda, const cplx* b,
                         double* scratch, int nblk, cudaStream_t stream);                                                      // comp 0: max, 1: sum, 3: max
cudaError_t vec_res_final(const cplx* V, const cplx* Y, int n, int C, int problem_type, const cplx* lambda, const cplx* b,
                          const double* scratch, int nblk, double* resid, cudaStream_t stream);
