// spmv.cuh -- CSR SpMM: Y[c] = A * V[c] for C candidate vectors ([C][n] layout).
#pragma once
#include "common.cuh"
// n rows (local rows of a row block), ncols = length of the vectors; pack_ws: optional [ncols][4] scratch -- with it, 2..4
// candidates are gathered from an interleaved copy (whole L2 sectors per matrix entry), bit-identical results
cudaError_t csr_spmm(const long long* rowptr, const int* colidx, const cplx* vals, const cplx* V, long long ldv, cplx* Y,
                     long long ldy, long long n, long long ncols, int C, cplx* pack_ws, cudaStream_t stream);
