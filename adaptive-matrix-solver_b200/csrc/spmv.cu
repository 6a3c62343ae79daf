// spmv.cu -- CSR sparse matrix x candidate block (SpMM) for the sparse GMRES path (AMS:47, 57 replaced by GMRES;
// the reference's matrix is scipy CSC, converted once to CSR at upload so that rows are contiguous).
#include "spmv.cuh"

namespace {

// 8 lanes per matrix row (rows have ~20 non-zeros in the K5 workload), 4 rows per warp, CB candidates per pass.
// Values / column indices are read in contiguous 128 B / 32 B segments per row; candidate entries are gathered.
constexpr int SP_NT = 256, SP_LANES = 8;
template <int CB>
__global__ void __launch_bounds__(SP_NT) csr_spmm_kernel(const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                                                         const cplx* __restrict__ vals, const cplx* __restrict__ V,
                                                         long long ldv, cplx* __restrict__ Y, long long ldy, long long n,
                                                         int c0, int ncand) {
    const long long row = ((long long)blockIdx.x * SP_NT + threadIdx.x) / SP_LANES;
    const int sub = threadIdx.x & (SP_LANES - 1);
    cplx acc[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) acc[c] = cmake(0.0, 0.0);
    if (row < n) {
        const long long k1 = rowptr[row + 1];
        for (long long k = rowptr[row] + sub; k < k1; k += SP_LANES) {
            const cplx a = __ldcs(&vals[k]);
            const int j = __ldcs(&colidx[k]);
#pragma unroll
            for (int c = 0; c < CB; ++c)
                if (c < ncand) cfma(acc[c], a, __ldg(&V[(long long)(c0 + c) * ldv + j]));
        }
    }
#pragma unroll
    for (int c = 0; c < CB; ++c) {
#pragma unroll
        for (int o = SP_LANES / 2; o > 0; o >>= 1) {
            acc[c].x += __shfl_xor_sync(0xffffffffu, acc[c].x, o);
            acc[c].y += __shfl_xor_sync(0xffffffffu, acc[c].y, o);
        }
        if (sub == 0 && row < n && c < ncand) Y[(long long)(c0 + c) * ldy + row] = acc[c];
    }
}

}  // namespace

cudaError_t csr_spmm(const long long* rowptr, const int* colidx, const cplx* vals, const cplx* V, long long ldv, cplx* Y,
                     long long ldy, long long n, int C, cudaStream_t stream) {
    const long long threads = n * SP_LANES;
    const unsigned grid = (unsigned)((threads + SP_NT - 1) / SP_NT);
    for (int c0 = 0; c0 < C; c0 += 4) {
        const int nc = (C - c0 < 4) ? (C - c0) : 4;
        if (nc == 1) csr_spmm_kernel<1><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, V, ldv, Y, ldy, n, c0, nc);
        else if (nc == 2) csr_spmm_kernel<2><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, V, ldv, Y, ldy, n, c0, nc);
        else csr_spmm_kernel<4><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, V, ldv, Y, ldy, n, c0, nc);
    }
    return cudaGetLastError();
}
