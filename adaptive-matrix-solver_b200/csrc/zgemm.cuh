// zgemm.cuh -- batched complex128 GEMM on the FP64 tensor pipe (DMMA) with TMA-bulk staged panels.
#pragma once
#include "common.cuh"

struct ZgemmParams {
    const cplx* A; long long lda, strideA;   // M x K, column-major
    const cplx* B; long long ldb, strideB;   // K x N, column-major
    cplx* C;       long long ldc, strideC;   // M x N, column-major
    int M, N, K, batch;
    int beta;      // 0: C = s*A*B        1: C = C + s*A*B
    int negate;    // s = -1 when set, else +1
    int algo3m;    // 1: three-real-product complex arithmetic (6 instead of 8 flops per complex FMA; LU updates)
    int tile_n;    // 3M kernel only: 0 / 48 = 128 x 48 CTA tiles, 32 = 128 x 32 tiles (skinny products whose N is a multiple of 32)
};

// Tensor-pipe kernel (any M, N, K >= 1).  In-place use (C == B) is supported for M <= 128: it is routed to the 128-row tile configuration so one CTA owns all rows of its columns.
cudaError_t zgemm_dmma_launch(const ZgemmParams& p, cudaStream_t stream);
// Plain FP64-FMA kernel, independent code path used by the tests to cross-check the tensor-pipe kernel.
cudaError_t zgemm_simple_launch(const ZgemmParams& p, cudaStream_t stream);
