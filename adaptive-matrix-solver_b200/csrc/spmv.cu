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
                                                        long long ncols, int c0, int ctotal) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    P += (long long)blockIdx.y * ncols * CB;               // blockIdx.y = group of CB candidates
    c0 += (int)blockIdx.y * CB;
    const int ncand = min(CB, ctotal - c0);
#pragma unroll
    for (int c = 0; c < CB; ++c) P[j * CB + c] = (c < ncand) ? __ldcs(&V[(long long)(c0 + c) * ldv + j]) : cmake(0.0, 0.0);
}

// Gather kernel for the interleaved copy.  A gathered entry costs one L1 wavefront per distinct 128 B line a warp instruction
// touches, whatever its width (ncu, round 1: the kernel that let ONE lane fetch all CB candidates of an entry was l1tex-bound at
// two 32 B loads per entry per lane).  Here the CB candidates of an entry are fetched by CB NEIGHBOURING lanes, 16 B each: a warp
// instruction covers 32 / CB entries with one fully used CB * 16 B segment per entry, and every lane carries one accumulator
// (~32 registers: 64 warps per SM keep the index -> gather chains of 64 rows in flight).
//   lane = CB * sub + c : sub = 0..7 walks the row's entries sub, sub + 8, ... exactly like the unpacked kernel, then the same
//   xor tree over sub -- so every candidate's sum is accumulated in the SAME order as in csr_spmm_kernel (bit-identical results).
template <int CB, int SP_U>
__global__ void __launch_bounds__(SP_NT) csr_spmm_packed_kernel(const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                                                                const cplx* __restrict__ vals, const cplx* __restrict__ P,
                                                                long long p_gstride, cplx* __restrict__ Y, long long ldy,
                                                                long long n, int c0, int ctotal) {
    constexpr int LPR = SP_LANES * CB;                     // lanes per row
    // blockIdx.y = group of CB candidates: its interleaved copy starts p_gstride elements after the previous group's
    P += (long long)blockIdx.y * p_gstride;
    c0 += (int)blockIdx.y * CB;
    const int ncand = min(CB, ctotal - c0);
    const long long row = ((long long)blockIdx.x * SP_NT + threadIdx.x) / LPR;
    const int l = threadIdx.x % LPR, sub = l / CB, c = l % CB;
    cplx acc = cmake(0.0, 0.0);
    if (row < n) {
        const long long k1 = rowptr[row + 1];
        for (long long kb = rowptr[row] + sub; kb < k1; kb += SP_LANES * SP_U) {
            cplx a[SP_U]; int j[SP_U];
#pragma unroll
            for (int u = 0; u < SP_U; ++u) {
                const long long k = kb + u * SP_LANES;
                const bool ok = k < k1;
                a[u] = ok ? __ldcs(&vals[k]) : cmake(0.0, 0.0);       // the CB lanes of an entry read the same address (broadcast)
                j[u] = ok ? __ldcs(&colidx[k]) : -1;
            }
            cplx v[SP_U];
#pragma unroll
            for (int u = 0; u < SP_U; ++u) v[u] = (j[u] >= 0) ? __ldg(&P[(long long)j[u] * CB + c]) : cmake(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < SP_U; ++u) cfma(acc, a[u], v[u]);
        }
    }
#pragma unroll
    for (int o = SP_LANES / 2; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o * CB);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o * CB);
    }
    if (sub == 0 && row < n && c < ncand) Y[(long long)(c0 + c) * ldy + row] = acc;
}

}  // namespace

cudaError_t csr_spmm_packed4(const long long* rowptr, const int* colidx, const cplx* vals, const cplx* P, long long p_gstride,
                             cplx* Y, long long ldy, long long n, int c0, int ctotal, int groups, cudaStream_t stream) {
    if (groups <= 0 || n <= 0) return cudaSuccess;
    const long long threads = n * SP_LANES * 4;
    dim3 grid((unsigned)((threads + SP_NT - 1) / SP_NT), (unsigned)groups);
    csr_spmm_packed_kernel<4, 3><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, P, p_gstride, Y, ldy, n, c0, ctotal);
    return cudaGetLastError();
}

cudaError_t csr_spmm(const long long* rowptr, const int* colidx, const cplx* vals, const cplx* V, long long ldv, cplx* Y,
                     long long ldy, long long n, long long ncols, int C, cplx* pack_ws, cudaStream_t stream) {
    const long long threads = n * SP_LANES;
    const unsigned grid = (unsigned)((threads + SP_NT - 1) / SP_NT);
    const unsigned pgrid = (unsigned)((ncols + 255) / 256);
    if (C <= 0) return cudaSuccess;
    // one candidate: latency bound on the dependent index -> gather chain, so occupancy beats unrolling (SP_U = 1, <= 32
    // registers, 8 CTAs / SM; measured 0.132 vs 0.140 ms at n = 1M)
    if (C == 1) { csr_spmm_kernel<1, 1, 8><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, V, ldv, Y, ldy, n, 0, 1); return cudaGetLastError(); }
    if (!pack_ws) {
        for (int c0 = 0; c0 < C; c0 += 4) {
            const int nc = (C - c0 < 4) ? (C - c0) : 4;
            if (nc == 1) csr_spmm_kernel<1, 1, 8><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, V, ldv, Y, ldy, n, c0, nc);
            else if (nc == 2) csr_spmm_kernel<2, 3, 1><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, V, ldv, Y, ldy, n, c0, nc);
            else csr_spmm_kernel<4, 3, 1><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, V, ldv, Y, ldy, n, c0, nc);
        }
        return cudaGetLastError();
    }
    // pack_ws holds ceil(C / 4) groups of [ncols][4]: ALL candidates are interleaved in one pass and multiplied in one launch
    // (blockIdx.y = group); two candidates use the half-width layout
    if (C == 2) {
        spmm_pack_kernel<2><<<dim3(pgrid, 1), 256, 0, stream>>>(V, ldv, pack_ws, ncols, 0, C);
        const unsigned grid2 = (unsigned)((threads * 2 + SP_NT - 1) / SP_NT);
        csr_spmm_packed_kernel<2, 3><<<dim3(grid2, 1), SP_NT, 0, stream>>>(rowptr, colidx, vals, pack_ws, 0, Y, ldy, n, 0, C);
        return cudaGetLastError();
    }
    const int groups = (C + 3) / 4;
    spmm_pack_kernel<4><<<dim3(pgrid, groups), 256, 0, stream>>>(V, ldv, pack_ws, ncols, 0, C);
    return csr_spmm_packed4(rowptr, colidx, vals, pack_ws, ncols * 4, Y, ldy, n, 0, C, groups, stream);
}

// number of complex elements of the interleaved copy csr_spmm needs for C candidates
size_t csr_spmm_pack_elems(long long ncols, int C) { return C <= 1 ? 0 : (size_t)ncols * 4 * (size_t)((C + 3) / 4); }
