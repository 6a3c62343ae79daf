// spmv.cu -- CSR sparse matrix x candidate block (SpMM) for the sparse GMRES path (AMS:47, 57 replaced by GMRES;
// the reference's matrix is scipy CSC, converted once to CSR at upload so that rows are contiguous).
#include "spmv.cuh"

namespace {

// 8 lanes per matrix row (rows have ~20 non-zeros in the K5 workload), 4 rows per warp, CB candidates per pass.
// Values / column indices are read in contiguous 128 B / 32 B segments per row; candidate entries are gathered.
constexpr int SP_NT = 256, SP_LANES = 8;
template <int CB, int SP_U, int MINB>
__global__ void __launch_bounds__(SP_NT, MINB) csr_spmm_kernel(const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                                                         const cplx* __restrict__ vals, const cplx* __restrict__ V,
                                                         long long ldv, cplx* __restrict__ Y, long long ldy, long long n,
                                                         int c0, int ncand) {
    const long long row = ((long long)blockIdx.x * SP_NT + threadIdx.x) / SP_LANES;
    const int sub = threadIdx.x & (SP_LANES - 1);
    cplx acc[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) acc[c] = cmake(0.0, 0.0);
    if (row < n) {
        // SP_U entries per lane are fetched before any of them is used: value + index loads of a whole row (<= 24
        // entries) are in flight together, then all gathers -- two dependent memory round trips per row instead of six
        const long long k1 = rowptr[row + 1];
        for (long long kb = rowptr[row] + sub; kb < k1; kb += SP_LANES * SP_U) {
            cplx a[SP_U]; int j[SP_U];
#pragma unroll
            for (int u = 0; u < SP_U; ++u) {
                const long long k = kb + u * SP_LANES;
                const bool ok = k < k1;
                a[u] = ok ? __ldcs(&vals[k]) : cmake(0.0, 0.0);
                j[u] = ok ? __ldcs(&colidx[k]) : -1;           // -1: no entry (nothing is gathered, 0 * NaN cannot occur)
            }
            cplx v[SP_U][CB];
#pragma unroll
            for (int u = 0; u < SP_U; ++u)
#pragma unroll
                for (int c = 0; c < CB; ++c)
                    v[u][c] = (c < ncand && j[u] >= 0) ? __ldg(&V[(long long)(c0 + c) * ldv + j[u]]) : cmake(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < SP_U; ++u)
#pragma unroll
                for (int c = 0; c < CB; ++c) cfma(acc[c], a[u], v[u][c]);
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

// ---- packed gathers ------------------------------------------------------------------------------------------------
// With several candidates the SpMM is bound by L2 SECTOR traffic, not by HBM: a gathered 16 B entry moves a 32 B sector,
// and CB separate vectors mean CB half-used sectors per matrix entry.  The CB vectors are therefore first interleaved
// into P[j][CB] (one streaming pass, 32 n CB bytes), so that one matrix entry gathers CB * 16 contiguous bytes = whole
// sectors (CB = 2: 1 sector, CB = 4: 2 sectors instead of 4).  Accumulation order per candidate is unchanged, so the
// results are bit-identical to the unpacked kernel.
template <int CB>
__global__ void __launch_bounds__(256) spmm_pack_kernel(const cplx* __restrict__ V, long long ldv, cplx* __restrict__ P,
                                                        long long ncols, int c0, int ncand) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
#pragma unroll
    for (int c = 0; c < CB; ++c) P[j * CB + c] = (c < ncand) ? __ldcs(&V[(long long)(c0 + c) * ldv + j]) : cmake(0.0, 0.0);
}

template <int CB> struct PackedRow;
template <> struct PackedRow<2> {
    cplx v[2];
    __device__ __forceinline__ void load(const cplx* p) {
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0].x), "=d"(v[0].y), "=d"(v[1].x), "=d"(v[1].y) : "l"(p));
    }
};
template <> struct PackedRow<4> {
    cplx v[4];
    __device__ __forceinline__ void load(const cplx* p) {
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0].x), "=d"(v[0].y), "=d"(v[1].x), "=d"(v[1].y) : "l"(p));
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[2].x), "=d"(v[2].y), "=d"(v[3].x), "=d"(v[3].y) : "l"(p + 2));
    }
};

template <int CB, int SP_U>
__global__ void __launch_bounds__(SP_NT) csr_spmm_packed_kernel(const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                                                                const cplx* __restrict__ vals, const cplx* __restrict__ P,
                                                                cplx* __restrict__ Y, long long ldy, long long n, int c0, int ncand) {
    const long long row = ((long long)blockIdx.x * SP_NT + threadIdx.x) / SP_LANES;
    const int sub = threadIdx.x & (SP_LANES - 1);
    cplx acc[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) acc[c] = cmake(0.0, 0.0);
    if (row < n) {
        const long long k1 = rowptr[row + 1];
        for (long long kb = rowptr[row] + sub; kb < k1; kb += SP_LANES * SP_U) {
            cplx a[SP_U]; int j[SP_U];
#pragma unroll
            for (int u = 0; u < SP_U; ++u) {
                const long long k = kb + u * SP_LANES;
                const bool ok = k < k1;
                a[u] = ok ? __ldcs(&vals[k]) : cmake(0.0, 0.0);
                j[u] = ok ? __ldcs(&colidx[k]) : -1;
            }
            PackedRow<CB> v[SP_U];
#pragma unroll
            for (int u = 0; u < SP_U; ++u) {
                if (j[u] >= 0) v[u].load(P + (long long)j[u] * CB);
                else {
#pragma unroll
                    for (int c = 0; c < CB; ++c) v[u].v[c] = cmake(0.0, 0.0);
                }
            }
#pragma unroll
            for (int u = 0; u < SP_U; ++u)
#pragma unroll
                for (int c = 0; c < CB; ++c) cfma(acc[c], a[u], v[u].v[c]);
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
                     long long ldy, long long n, long long ncols, int C, cplx* pack_ws, cudaStream_t stream) {
    const long long threads = n * SP_LANES;
    const unsigned grid = (unsigned)((threads + SP_NT - 1) / SP_NT);
    const unsigned pgrid = (unsigned)((ncols + 255) / 256);
    for (int c0 = 0; c0 < C; c0 += 4) {
        const int nc = (C - c0 < 4) ? (C - c0) : 4;
        // one candidate: latency bound on the dependent index -> gather chain, so occupancy beats unrolling (SP_U = 1, <= 32
        // registers, 8 CTAs / SM; measured 0.132 vs 0.140 ms at n = 1M)
        if (nc == 1) csr_spmm_kernel<1, 1, 8><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, V, ldv, Y, ldy, n, c0, nc);
        else if (!pack_ws) {
            if (nc == 2) csr_spmm_kernel<2, 3, 1><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, V, ldv, Y, ldy, n, c0, nc);
            else csr_spmm_kernel<4, 3, 1><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, V, ldv, Y, ldy, n, c0, nc);
        } else if (nc == 2) {
            spmm_pack_kernel<2><<<pgrid, 256, 0, stream>>>(V, ldv, pack_ws, ncols, c0, nc);
            csr_spmm_packed_kernel<2, 3><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, pack_ws, Y, ldy, n, c0, nc);
        } else {
            spmm_pack_kernel<4><<<pgrid, 256, 0, stream>>>(V, ldv, pack_ws, ncols, c0, nc);
            csr_spmm_packed_kernel<4, 3><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, pack_ws, Y, ldy, n, c0, nc);
        }
    }
    return cudaGetLastError();
}
