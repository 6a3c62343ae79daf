// spmv.cu -- CSR sparse matrix x candidate block (SpMM) for the sparse GMRES path (AMS:47, 57 replaced by GMRES;
// the reference's matrix is scipy CSC, converted once to CSR at upload so that rows are contiguous).
#include <cstdlib>
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
// The row chain rowptr -> (index, value) -> gather -> reduce is three dependent memory round trips; with one row per lane group
// at a time the kernel was bound by that latency (ncu, round 2: DRAM 31 %, L2 31 %, l1tex 58 % busy at 63 % occupancy).  The
// lane groups therefore walk their rows grid-strided and SOFTWARE-PIPELINED: while the gathers of row r are in flight, the
// (index, value) loads of the group's next row and the rowptr pair of the row after that are already issued.
// MINB = CTAs per SM the register allocation is held to (the kernel is bound by the latency of its dependent loads: warps in
// flight matter more than look-ahead depth -- two rows of look-ahead at 114 registers measured 0.48 ms, one row at 64 registers
// 0.29 ms for 4 candidates at n = 1M)
// ONECHUNK: the host knows that no row has more than SP_LANES * SP_U entries (max_row_len, found when the CSR is built): the
// loop over further chunks is compiled out.  The kernel is bound by instruction issue + latency at ~0.55 IPC per sub-partition
// (ncu, round 2b: 171 instructions per row with the chunk loop in place and 0.30 ms at n = 1M, 4 candidates; 132 and 0.21 ms
// without it -- profiles/microbench/spmm_ablate.cu)
template <int CB, int SP_U, int MINB, bool ONECHUNK = false>
__global__ void __launch_bounds__(SP_NT, MINB) csr_spmm_packed_kernel(const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                                                                const cplx* __restrict__ vals, const cplx* __restrict__ P,
                                                                long long p_gstride, cplx* __restrict__ Y, long long ldy,
                                                                long long n, int c0, int ctotal) {
    constexpr int LPR = SP_LANES * CB;                     // lanes per row
    constexpr int GPB = SP_NT / LPR;                       // lane groups (rows in flight) per CTA
    // blockIdx.y = group of CB candidates: its interleaved copy starts p_gstride elements after the previous group's
    P += (long long)blockIdx.y * p_gstride;
    c0 += (int)blockIdx.y * CB;
    const int ncand = min(CB, ctotal - c0);
    const int l = threadIdx.x % LPR, sub = l / CB, c = l % CB;
    const long long stride = (long long)gridDim.x * GPB;
    long long row = (long long)blockIdx.x * GPB + threadIdx.x / LPR;
    auto load_entries = [&](long long k0, long long k1, cplx* a, int* j) {
#pragma unroll
        for (int u = 0; u < SP_U; ++u) {
            const long long k = k0 + sub + u * SP_LANES;
            const bool ok = k < k1;
            a[u] = ok ? __ldcs(&vals[k]) : cmake(0.0, 0.0);           // the CB lanes of an entry read the same address (broadcast)
            j[u] = ok ? __ldcs(&colidx[k]) : -1;                      // -1: no entry (nothing is gathered, 0 * NaN cannot occur)
        }
    };
    long long k0 = 0, k1 = 0, k0n = 0, k1n = 0;
    cplx an[SP_U]; int jn[SP_U];
    if (row < n) { k0 = rowptr[row]; k1 = rowptr[row + 1]; }
    load_entries(k0, k1, an, jn);
    long long rown = row + stride;
    if (rown < n) { k0n = rowptr[rown]; k1n = rowptr[rown + 1]; }
    // warp-uniform trip count (a warp holds 32 / LPR lane groups): a group past its last row runs empty iterations -- its
    // ranges are empty, so it loads and gathers nothing -- and takes part in the full-mask shuffles
    while (__any_sync(0xffffffffu, row < n)) {
        cplx a[SP_U]; int j[SP_U];
#pragma unroll
        for (int u = 0; u < SP_U; ++u) { a[u] = an[u]; j[u] = jn[u]; }
        cplx v[SP_U];
#pragma unroll
        for (int u = 0; u < SP_U; ++u) v[u] = (j[u] >= 0) ? __ldg(&P[(long long)j[u] * CB + c]) : cmake(0.0, 0.0);
        // next row of this lane group: its entries, and the rowptr pair of the row after it
        const long long kc0 = k0, kc1 = k1, rown2 = rown + stride;
        long long k0nn = 0, k1nn = 0;
        load_entries(k0n, k1n, an, jn);                              // empty range (0, 0) when there is no next row
        if (rown2 < n) { k0nn = rowptr[rown2]; k1nn = rowptr[rown2 + 1]; }
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < SP_U; ++u) cfma(acc, a[u], v[u]);
        // rows with more than SP_LANES * SP_U entries: the remaining chunks, same order as before, not pipelined
        if (!ONECHUNK)
            for (long long kb = kc0 + SP_LANES * SP_U; kb < kc1; kb += SP_LANES * SP_U) {
                load_entries(kb, kc1, a, j);
#pragma unroll
                for (int u = 0; u < SP_U; ++u) v[u] = (j[u] >= 0) ? __ldg(&P[(long long)j[u] * CB + c]) : cmake(0.0, 0.0);
#pragma unroll
                for (int u = 0; u < SP_U; ++u) cfma(acc, a[u], v[u]);
            }
#pragma unroll
        for (int o = SP_LANES / 2; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o * CB);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o * CB);
        }
        if (sub == 0 && c < ncand && row < n) Y[(long long)(c0 + c) * ldy + row] = acc;
        row = rown; rown = rown2; k0 = k0n; k1 = k1n; k0n = k0nn; k1n = k1nn;
    }
}

// Eight candidates per pass: P[j][8], one matrix entry gathers one whole 128-byte line, fetched by 8 neighbouring lanes.  A warp
// instruction then covers 4 entries with 4 fully used lines (the 4-candidate layout: 8 entries, 8 half lines -- the kernel is bound
// by L1 wavefronts and L2 sectors per gathered byte, not by DRAM), and the matrix is streamed once per 8 candidates.  One warp walks
// one row at a time: lane = 8 * s + c, s = 0..3.  Lane s keeps TWO accumulators, for the entries = s and = s + 4 (mod 8), adds them
// and then runs the xor tree over s: exactly the association of the 8-lane kernels (whose first tree step adds lanes k and k ^ 4),
// so the results stay bit-identical to csr_spmm_kernel.  Rows are walked grid-strided and software-pipelined like above.
template <int SP_U, int MINB>
__global__ void __launch_bounds__(SP_NT, MINB) csr_spmm_packed8_kernel(const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                                                                 const cplx* __restrict__ vals, const cplx* __restrict__ P,
                                                                 long long p_gstride, cplx* __restrict__ Y, long long ldy,
                                                                 long long n, int c0, int ctotal) {
    constexpr int CB = 8, SUBS = 4, NE = 2 * SP_U;         // NE entries per lane and chunk: u even -> residue s, u odd -> residue s + 4
    constexpr int GPB = SP_NT / 32;                        // rows in flight per CTA
    P += (long long)blockIdx.y * p_gstride;
    c0 += (int)blockIdx.y * CB;
    const int ncand = min(CB, ctotal - c0);
    const int lane = threadIdx.x & 31, s = lane >> 3, c = lane & 7;
    const long long stride = (long long)gridDim.x * GPB;
    long long row = (long long)blockIdx.x * GPB + (threadIdx.x >> 5);
    auto load_entries = [&](long long k0, long long k1, cplx* a, int* j) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const long long k = k0 + s + e * SUBS;          // e = 2u: residue s, e = 2u + 1: residue s + 4 (mod 8)
            const bool ok = k < k1;
            a[e] = ok ? __ldcs(&vals[k]) : cmake(0.0, 0.0);
            j[e] = ok ? __ldcs(&colidx[k]) : -1;
        }
    };
    long long k0 = 0, k1 = 0, k0n = 0, k1n = 0;
    cplx an[NE]; int jn[NE];
    if (row < n) { k0 = rowptr[row]; k1 = rowptr[row + 1]; }
    load_entries(k0, k1, an, jn);
    long long rown = row + stride;
    if (rown < n) { k0n = rowptr[rown]; k1n = rowptr[rown + 1]; }
    while (row < n) {                                       // warp-uniform: the whole warp works on one row
        cplx a[NE]; int j[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) { a[e] = an[e]; j[e] = jn[e]; }
        cplx v[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) v[e] = (j[e] >= 0) ? __ldg(&P[(long long)j[e] * CB + c]) : cmake(0.0, 0.0);
        const long long kc0 = k0, kc1 = k1, rown2 = rown + stride;
        long long k0nn = 0, k1nn = 0;
        load_entries(k0n, k1n, an, jn);
        if (rown2 < n) { k0nn = rowptr[rown2]; k1nn = rowptr[rown2 + 1]; }
        cplx acc0 = cmake(0.0, 0.0), acc1 = cmake(0.0, 0.0);
#pragma unroll
        for (int e = 0; e < NE; e += 2) { cfma(acc0, a[e], v[e]); cfma(acc1, a[e + 1], v[e + 1]); }
        for (long long kb = kc0 + SUBS * NE; kb < kc1; kb += SUBS * NE) {
            load_entries(kb, kc1, a, j);
#pragma unroll
            for (int e = 0; e < NE; ++e) v[e] = (j[e] >= 0) ? __ldg(&P[(long long)j[e] * CB + c]) : cmake(0.0, 0.0);
#pragma unroll
            for (int e = 0; e < NE; e += 2) { cfma(acc0, a[e], v[e]); cfma(acc1, a[e + 1], v[e + 1]); }
        }
        cplx acc = cmake(acc0.x + acc1.x, acc0.y + acc1.y);
#pragma unroll
        for (int o = SUBS / 2; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o * CB);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o * CB);
        }
        if (s == 0 && c < ncand) Y[(long long)(c0 + c) * ldy + row] = acc;
        row = rown; rown = rown2; k0 = k0n; k1 = k1n; k0n = k0nn; k1n = k1nn;
    }
}

// persistent-style grid for the pipelined kernel: exactly the CTAs that are resident at once (occupancy query per instantiation),
// every lane group walks many rows
template <int CB, int SP_U, int MINB, bool ONECHUNK = false>
static unsigned spmm_pipe_grid(long long n) {
    static int per_sm = 0;
    if (!per_sm) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_spmm_packed_kernel<CB, SP_U, MINB, ONECHUNK>, SP_NT, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    }
    const long long groups_per_block = SP_NT / (SP_LANES * CB);
    const long long need = (n + groups_per_block - 1) / groups_per_block;
    const long long cap = (long long)MAUS_SM_COUNT_B200 * per_sm;
    return (unsigned)(need < cap ? need : cap);
}

template <int SP_U, int MINB>
static unsigned spmm_pipe8_grid(long long n) {
    static int per_sm = 0;
    if (!per_sm) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_spmm_packed8_kernel<SP_U, MINB>, SP_NT, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    }
    const long long rows_per_block = SP_NT / 32;
    const long long need = (n + rows_per_block - 1) / rows_per_block;
    const long long cap = (long long)MAUS_SM_COUNT_B200 * per_sm;
    return (unsigned)(need < cap ? need : cap);
}

}  // namespace

cudaError_t csr_spmm_packed4(const long long* rowptr, const int* colidx, const cplx* vals, const cplx* P, long long p_gstride,
                             cplx* Y, long long ldy, long long n, int c0, int ctotal, int groups, int max_row_len, cudaStream_t stream) {
    if (groups <= 0 || n <= 0) return cudaSuccess;
    // Measured alternatives that did not beat this kernel (round 2, 4 candidates, n = 1M, 0.29 - 0.33 ms): two rows of look-ahead
    // (114 registers, 0.48 ms), 48 / 40-register builds with spills (0.49 / 0.82 ms), a shared-memory staged matrix stream with six
    // gathers in flight per lane (commit 109fb77, 0.37 ms), an L2 persisting window over the interleaved copy (commit 1638896, 0.32 - 0.51 ms).
    // Round 2b (profiles/microbench/l2_gather.cu, spmm_ablate.cu): the bare L2 gather of the same 1.34 GB takes 0.096 ms, so the
    // kernel is NOT at the L2 limit; it is bound by bytes in flight per SM x instruction issue, and compiling the chunk loop out
    // for matrices whose rows fit one chunk is worth 0.30 -> 0.21 ms.
    if (max_row_len > 0 && max_row_len <= SP_LANES * 3) {
        dim3 grid(spmm_pipe_grid<4, 3, 4, true>(n), (unsigned)groups);
        csr_spmm_packed_kernel<4, 3, 4, true><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, P, p_gstride, Y, ldy, n, c0, ctotal);
        return cudaGetLastError();
    }
    dim3 grid(spmm_pipe_grid<4, 3, 4>(n), (unsigned)groups);
    csr_spmm_packed_kernel<4, 3, 4><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, P, p_gstride, Y, ldy, n, c0, ctotal);
    return cudaGetLastError();
}

cudaError_t csr_spmm(const long long* rowptr, const int* colidx, const cplx* vals, const cplx* V, long long ldv, cplx* Y,
                     long long ldy, long long n, long long ncols, int C, cplx* pack_ws, int max_row_len, cudaStream_t stream) {
    const bool one = max_row_len > 0 && max_row_len <= SP_LANES * 3;
    const long long threads = n * SP_LANES;
    const unsigned grid = (unsigned)((threads + SP_NT - 1) / SP_NT);
    const unsigned pgrid = (unsigned)((ncols + 255) / 256);
    if (C <= 0) return cudaSuccess;
    // one candidate: latency bound on the dependent index -> gather chain, so occupancy beats unrolling (SP_U = 1, <= 32
    // registers, 8 CTAs / SM; measured 0.132 vs 0.140 ms at n = 1M)
    if (C == 1) {
        static int pipe1 = -1;               // MAUS_SPMM_PIPE1=0: the one-row-at-a-time kernel (A/B measurements)
        if (pipe1 < 0) { const char* e = getenv("MAUS_SPMM_PIPE1"); pipe1 = e ? (atoi(e) != 0) : 1; }
        if (pipe1 && one) csr_spmm_packed_kernel<1, 3, 4, true><<<dim3(spmm_pipe_grid<1, 3, 4, true>(n), 1), SP_NT, 0, stream>>>(rowptr, colidx, vals, V, 0, Y, ldy, n, 0, 1);
        else if (pipe1) csr_spmm_packed_kernel<1, 3, 4><<<dim3(spmm_pipe_grid<1, 3, 4>(n), 1), SP_NT, 0, stream>>>(rowptr, colidx, vals, V, 0, Y, ldy, n, 0, 1);
        else csr_spmm_kernel<1, 1, 8><<<grid, SP_NT, 0, stream>>>(rowptr, colidx, vals, V, ldv, Y, ldy, n, 0, 1);
        return cudaGetLastError();
    }
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
        if (one) csr_spmm_packed_kernel<2, 3, 4, true><<<dim3(spmm_pipe_grid<2, 3, 4, true>(n), 1), SP_NT, 0, stream>>>(rowptr, colidx, vals, pack_ws, 0, Y, ldy, n, 0, C);
        else csr_spmm_packed_kernel<2, 3, 4><<<dim3(spmm_pipe_grid<2, 3, 4>(n), 1), SP_NT, 0, stream>>>(rowptr, colidx, vals, pack_ws, 0, Y, ldy, n, 0, C);
        return cudaGetLastError();
    }
    // full groups of 8 candidates through the 8-wide layout, the rest (<= 7) in groups of 4 behind them in the same buffer
    static int pack8 = -1;               // MAUS_SPMM_PACK8=0: 4-wide groups only (A/B measurements); 2 / 3: register allocation variants
    if (pack8 < 0) { const char* e = getenv("MAUS_SPMM_PACK8"); pack8 = e ? atoi(e) : 1; }
    // measured (round 2, profiles/spmm_pack8_sweep.py, 8 / 16 candidates): -13 % at n = 125 000 and 250 000, -5 % / +2 % at 500 000,
    // +5 % at 1 000 000 -- the 8-wide copy (128 B per column) must stay L2-resident next to the matrix stream, so it is used
    // while all its groups together take at most 40 MB; above that the 4-wide groups (64 MB at n = 1M) are faster
    const bool fits8 = (size_t)ncols * 128u * (size_t)(C / 8) <= (size_t)40 << 20;
    const int g8 = (pack8 && fits8) ? C / 8 : 0;
    const int rest = C - 8 * g8;
    if (g8 > 0) {
        spmm_pack_kernel<8><<<dim3(pgrid, g8), 256, 0, stream>>>(V, ldv, pack_ws, ncols, 0, 8 * g8);
        if (pack8 == 2) csr_spmm_packed8_kernel<3, 3><<<dim3(spmm_pipe8_grid<3, 3>(n), g8), SP_NT, 0, stream>>>(rowptr, colidx, vals, pack_ws, ncols * 8, Y, ldy, n, 0, 8 * g8);
        else if (pack8 == 3) csr_spmm_packed8_kernel<3, 4><<<dim3(spmm_pipe8_grid<3, 4>(n), g8), SP_NT, 0, stream>>>(rowptr, colidx, vals, pack_ws, ncols * 8, Y, ldy, n, 0, 8 * g8);
        else csr_spmm_packed8_kernel<3, 2><<<dim3(spmm_pipe8_grid<3, 2>(n), g8), SP_NT, 0, stream>>>(rowptr, colidx, vals, pack_ws, ncols * 8, Y, ldy, n, 0, 8 * g8);
        if (cudaGetLastError() != cudaSuccess) return cudaGetLastError();
    }
    if (rest == 0) return cudaGetLastError();
    cplx* ws4 = pack_ws + (size_t)ncols * 8 * g8;
    if (rest == 1) {
        if (one) csr_spmm_packed_kernel<1, 3, 4, true><<<dim3(spmm_pipe_grid<1, 3, 4, true>(n), 1), SP_NT, 0, stream>>>(rowptr, colidx, vals, V + (long long)(8 * g8) * ldv, 0,
                                                                                                        Y + (long long)(8 * g8) * ldy, ldy, n, 0, 1);
        else csr_spmm_packed_kernel<1, 3, 4><<<dim3(spmm_pipe_grid<1, 3, 4>(n), 1), SP_NT, 0, stream>>>(rowptr, colidx, vals, V + (long long)(8 * g8) * ldv, 0,
                                                                                                 Y + (long long)(8 * g8) * ldy, ldy, n, 0, 1);
        return cudaGetLastError();
    }
    const int groups = (rest + 3) / 4;
    spmm_pack_kernel<4><<<dim3(pgrid, groups), 256, 0, stream>>>(V + (long long)(8 * g8) * ldv, ldv, ws4, ncols, 0, rest);
    return csr_spmm_packed4(rowptr, colidx, vals, ws4, ncols * 4, Y + (long long)(8 * g8) * ldy, ldy, n, 0, rest, groups, max_row_len, stream);
}

// number of complex elements of the interleaved copy csr_spmm needs for C candidates
size_t csr_spmm_pack_elems(long long ncols, int C) { return C <= 1 ? 0 : (size_t)ncols * 4 * (size_t)((C + 3) / 4); }
