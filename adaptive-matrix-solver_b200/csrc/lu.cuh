// lu.cuh -- batched complex128 LU with partial pivoting on augmented systems [H | rhs] (replaces the LAPACK zgesv
// behind scipy.linalg.solve at AMS:59).
#pragma once
#include "common.cuh"

constexpr int LU_NB = 128;          // outer block (panel width)
constexpr int LU_MAX_PAIRS = 2 * LU_NB;
constexpr int LU_MAX_N = 8192;      // rows a panel cluster (8 CTAs x 512 threads x 2 rows) can own

// per-candidate row-permutation record written by the panel kernel, consumed by lu_permute_rows
struct LuPairs {
    int count;
    int dst[LU_MAX_PAIRS];
    int src[LU_MAX_PAIRS];
};

// W: batch matrices, column-major, ld = n, n+1 columns (last column = right-hand side), strideW elements apart.
// H_b = Acm + (psi_b - sigma_b) I + R_b  with R_b from Philox (key_b) or R_host (batch must be 1) or none.
// conj_in: the entries of Acm are conjugated while they are read (Acm = the row-major copy + conj_in = A^H, diag.cu).
cudaError_t lu_build_aug(cplx* W, long long strideW, int n, int batch, const cplx* Acm, const cplx* sigma,
                         const double* psi, const unsigned long long* keys, const cplx* R_cm,
                         const cplx* rhs, long long rhs_stride, cudaStream_t stream, int conj_in = 0);

// one panel step k0: factor W[k0:n, k0:k0+jb] with implicit partial pivoting, emit the row permutation.
cudaError_t lu_panel(cplx* W, long long strideW, int n, int k0, int jb, int batch, LuPairs* pairs, int* info,
                     cudaStream_t stream);
// apply the panel's row permutation (rows relative to k0) to columns [colstart, n]: the panel, the trailing matrix, the
// rhs column and -- for the second panel of a pair -- the previous panel's L21 whose update is still pending
cudaError_t lu_permute_rows(cplx* W, long long strideW, int n, int k0, int colstart, int batch, const LuPairs* pairs,
                            cudaStream_t stream);
// Linv_b = inverse of the unit-lower-triangular L11 block (jb x jb, ld = LU_NB)
cudaError_t lu_trtri(const cplx* W, long long strideW, int n, int k0, int jb, int batch, cplx* Linv,
                     cudaStream_t stream);
// x_b = U_b^-1 y_b (y = last column of W after the factorisation); writes X[b][n]; flags non-finite results and
// zero pivots into status (MAUS_ST_*), leaves other status words untouched.
cudaError_t lu_backsolve(const cplx* W, long long strideW, int n, int batch, const int* info, cplx* X,
                         int* status, cudaStream_t stream);
