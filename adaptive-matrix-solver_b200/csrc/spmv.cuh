// spmv.cuh -- CSR SpMM: Y[c] = A * V[c] for C candidate vectors ([C][n] layout).
#pragma once
#include "common.cuh"
cudaError_t csr_spmm(const long long* rowptr, const int* colidx, const cplx* vals, const cplx* V, long long ldv, cplx* Y,
                     long long ldy, long long n, int C, cudaStream_t stream);
