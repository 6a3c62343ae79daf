// spmv.cu -- CSR sparse matrix x candidate block (SpMM) for the sparse GMRES path (AMS:47, 57 replaced by GMRES;
// the reference's matrix is scipy CSC, converted once to CSR at upload so that rows are contiguous).
#include <cstdlib>
#include <cstdint>
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
template <int CB, int SP_U, int MINB>
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

// ---- staged variant: the matrix stream goes through SHARED memory -------------------------------------------------------------
// In the kernel above every in-flight row holds its (value, index) pairs in registers -- four times over, once per candidate
// lane -- which caps the warps per SM at 32 and with them the gathers in flight.  Here a CTA walks blocks of ST_RB rows; the
// contiguous value / index ranges of a block are fetched by two bulk async copies (TMA, no registers, two stages deep), the
// lanes read them with LDS and only the gathers remain as exposed global round trips: ~45 registers, five CTAs per SM, six
// gathers in flight per lane.  Blocks with more than ST_CAP entries (rows far longer than the K5 family's ~21) take the same
// arithmetic from global memory.  Per-candidate summation order unchanged (bit-identical to the other kernels).
constexpr int ST_RB = 32, ST_CAP = 1024, ST_NT = 256;
struct __align__(16) SpmmStage {
    cplx vals[ST_CAP];
    int idx[ST_CAP + 8];
    long long rp[ST_RB + 1];
    int direct, pad;
};
__device__ __forceinline__ uint32_t sp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sp_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sp_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sp_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sp_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sp_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(sp_smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void sp_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sp_smem_u32(dst)), "l"(src), "r"(bytes), "r"(sp_smem_u32(bar)) : "memory");
}

template <int CB, int MINB>
__global__ void __launch_bounds__(ST_NT, MINB) csr_spmm_staged_kernel(const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                                                                      const cplx* __restrict__ vals, const cplx* __restrict__ P,
                                                                      long long p_gstride, cplx* __restrict__ Y, long long ldy,
                                                                      long long n, int c0, int ctotal) {
    constexpr int LPR = SP_LANES * CB;                     // lanes per row
    constexpr int RPW = 32 / LPR;                          // rows a warp works on at the same time
    constexpr int NW = ST_NT / 32;
    static_assert(ST_RB % (NW * RPW) == 0, "rows of a block must split evenly over the warps");
    __shared__ SpmmStage st[2];
    __shared__ __align__(8) uint64_t bar[2];
    P += (long long)blockIdx.y * p_gstride;                // blockIdx.y = group of CB candidates
    c0 += (int)blockIdx.y * CB;
    const int ncand = min(CB, ctotal - c0);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int l = lane % LPR, sub = l / CB, c = l % CB, rsel = lane / LPR;
    const long long nblocks = (n + ST_RB - 1) / ST_RB;
    if (t == 0) { sp_mbar_init(&bar[0], 1); sp_mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    // warp 0 stages block b into stage s: the rowptr slice (already in its registers), then the two bulk copies
    auto issue = [&](long long b, int s, long long rp_lane, long long rp_last) {
        if (b >= nblocks) return;
        const long long r0 = b * ST_RB;
        const int nr = (int)min((long long)ST_RB, n - r0);
        if (lane <= nr) st[s].rp[lane] = rp_lane;
        if (lane == 0 && nr == ST_RB) st[s].rp[ST_RB] = rp_last;
        __syncwarp();
        if (lane == 0) {
            const long long k0 = st[s].rp[0], k1 = st[s].rp[nr];
            const long long ks = k0 & ~3LL;                                  // index copy starts 16-byte aligned
            const long long E = k1 - k0;
            const bool direct = E > ST_CAP;
            st[s].direct = direct ? 1 : 0;
            uint32_t bv = 0, bi = 0;
            if (!direct && E > 0) { bv = (uint32_t)(E * sizeof(cplx)); bi = (uint32_t)(((k1 - ks + 3) & ~3LL) * sizeof(int)); }
            sp_mbar_expect_tx(&bar[s], bv + bi);
            if (bv) { sp_bulk_g2s(st[s].vals, vals + k0, bv, &bar[s]); sp_bulk_g2s(st[s].idx, colidx + ks, bi, &bar[s]); }
        }
    };
    auto load_rp = [&](long long b, long long& rp_lane, long long& rp_last) {
        rp_lane = 0; rp_last = 0;
        if (b >= nblocks) return;
        const long long r0 = b * ST_RB;
        const int nr = (int)min((long long)ST_RB, n - r0);
        if (lane <= nr) rp_lane = rowptr[r0 + lane];
        if (lane == 0 && nr == ST_RB) rp_last = rowptr[r0 + ST_RB];
    };
    long long b = blockIdx.x;
    if (warp == 0) {
        long long a0, a1;
        load_rp(b, a0, a1); issue(b, 0, a0, a1);
        load_rp(b + gridDim.x, a0, a1); issue(b + gridDim.x, 1, a0, a1);
    }
    uint32_t phase[2] = {0, 0};
    for (int s = 0; b < nblocks; b += gridDim.x, s ^= 1) {
        long long nx0 = 0, nx1 = 0;
        if (warp == 0) load_rp(b + 2LL * gridDim.x, nx0, nx1);               // in flight behind this block's gathers
        sp_mbar_wait(&bar[s], phase[s]); phase[s] ^= 1;
        const SpmmStage& S = st[s];
        const long long r0 = b * ST_RB;
        const long long kbase = S.rp[0], ibase = kbase & ~3LL;
        const bool direct = S.direct != 0;
        // each warp: rows (warp * RPW + rsel) + i * NW * RPW of the block, two at a time -> 2 * 3 gathers in flight per lane
#pragma unroll 1
        for (int i = 0; i < ST_RB / (NW * RPW); i += 2) {
            cplx acc[2]; long long ra[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                acc[h] = cmake(0.0, 0.0);
                const int lr = warp * RPW + rsel + (i + h) * NW * RPW;
                ra[h] = (i + h < ST_RB / (NW * RPW)) ? r0 + lr : n;
            }
            long long k0r[2], k1r[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool live = ra[h] < n;
                k0r[h] = live ? S.rp[ra[h] - r0] : 0;
                k1r[h] = live ? S.rp[ra[h] - r0 + 1] : 0;
            }
            // first chunk of both rows: indices -> gathers issued together, then values and FMAs
            int j[2][3]; cplx v[2][3];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const long long k = k0r[h] + sub + u * SP_LANES;
                    j[h][u] = (k < k1r[h]) ? (direct ? __ldcs(&colidx[k]) : S.idx[k - ibase]) : -1;
                }
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int u = 0; u < 3; ++u) v[h][u] = (j[h][u] >= 0) ? __ldg(&P[(long long)j[h][u] * CB + c]) : cmake(0.0, 0.0);
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const long long k = k0r[h] + sub + u * SP_LANES;
                    if (j[h][u] >= 0) cfma(acc[h], direct ? __ldcs(&vals[k]) : S.vals[k - kbase], v[h][u]);
                    else cfma(acc[h], cmake(0.0, 0.0), v[h][u]);             // same operation sequence as the register kernels
                }
            // rows longer than 24 entries: remaining chunks, same order
#pragma unroll
            for (int h = 0; h < 2; ++h)
                for (long long kb = k0r[h] + SP_LANES * 3; kb < k1r[h]; kb += SP_LANES * 3) {
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const long long k = kb + sub + u * SP_LANES;
                        j[h][u] = (k < k1r[h]) ? (direct ? __ldcs(&colidx[k]) : S.idx[k - ibase]) : -1;
                    }
#pragma unroll
                    for (int u = 0; u < 3; ++u) v[h][u] = (j[h][u] >= 0) ? __ldg(&P[(long long)j[h][u] * CB + c]) : cmake(0.0, 0.0);
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const long long k = kb + sub + u * SP_LANES;
                        cfma(acc[h], (j[h][u] >= 0) ? (direct ? __ldcs(&vals[k]) : S.vals[k - kbase]) : cmake(0.0, 0.0), v[h][u]);
                    }
                }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int o = SP_LANES / 2; o > 0; o >>= 1) {
                    acc[h].x += __shfl_xor_sync(0xffffffffu, acc[h].x, o * CB);
                    acc[h].y += __shfl_xor_sync(0xffffffffu, acc[h].y, o * CB);
                }
                if (sub == 0 && c < ncand && ra[h] < n) Y[(long long)(c0 + c) * ldy + ra[h]] = acc[h];
            }
        }
        __syncthreads();                                                     // everybody is done with stage s
        if (warp == 0) issue(b + 2LL * gridDim.x, s, nx0, nx1);
    }
}

// persistent-style grid for the pipelined kernel: exactly the CTAs that are resident at once (occupancy query per instantiation),
// every lane group walks many rows
template <int CB, int SP_U, int MINB>
static unsigned spmm_pipe_grid(long long n) {
    static int per_sm = 0;
    if (!per_sm) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_spmm_packed_kernel<CB, SP_U, MINB>, SP_NT, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    }
    const long long groups_per_block = SP_NT / (SP_LANES * CB);
    const long long need = (n + groups_per_block - 1) / groups_per_block;
    const long long cap = (long long)MAUS_SM_COUNT_B200 * per_sm;
    return (unsigned)(need < cap ? need : cap);
}

}  // namespace

// L2 residency of the interleaved copy.  The gathers hit P at random; one group of P at n = 1M is 64 MB, the matrix streams
// 440 MB through the same L2 and (ncu, round 2) ~30 % of the gathered sectors missed and became random DRAM reads.  With
// MAUS_SPMM_L2PERSIST (fraction of P marked persisting, e.g. 0.6) the groups are launched one by one, each with an access-policy
// window over its own copy: persisting for P, streaming for everything else.
static float spmm_l2_persist_ratio() {
    static float ratio = -1.f;
    if (ratio < 0.f) {
        const char* e = getenv("MAUS_SPMM_L2PERSIST");
        ratio = e ? (float)atof(e) : 0.f;
        if (ratio > 0.f) {
            int dev = 0; cudaGetDevice(&dev);
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0 ||
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize) != cudaSuccess) {
                cudaGetLastError(); ratio = 0.f;
            }
        }
    }
    return ratio;
}

template <class Launch>
static cudaError_t spmm_launch_groups(const cplx* P, long long p_gstride, int groups, cudaStream_t stream, Launch launch) {
    const float ratio = spmm_l2_persist_ratio();
    if (ratio <= 0.f || p_gstride <= 0) { launch(0, groups); return cudaGetLastError(); }
    int dev = 0; cudaGetDevice(&dev);
    int max_win = 0; cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    for (int g = 0; g < groups; ++g) {
        cudaStreamAttrValue attr = {};
        attr.accessPolicyWindow.base_ptr = const_cast<cplx*>(P + (long long)g * p_gstride);
        size_t bytes = (size_t)p_gstride * sizeof(cplx);
        if (max_win > 0 && bytes > (size_t)max_win) bytes = (size_t)max_win;
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio = ratio > 1.f ? 1.f : ratio;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        launch(g, 1);
    }
    cudaStreamAttrValue off = {};
    off.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &off);
    return cudaGetLastError();
}

cudaError_t csr_spmm_packed4(const long long* rowptr, const int* colidx, const cplx* vals, const cplx* P, long long p_gstride,
                             cplx* Y, long long ldy, long long n, int c0, int ctotal, int groups, cudaStream_t stream) {
    if (groups <= 0 || n <= 0) return cudaSuccess;
    static int staged = -1;                  // MAUS_SPMM_STAGED=1: the shared-memory staged kernel (measured slower: 0.37 vs 0.33 ms)
    if (staged < 0) { const char* e = getenv("MAUS_SPMM_STAGED"); staged = e ? atoi(e) : 0; }
    if (staged) {
        static int per_sm = 0;
        if (!per_sm && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_spmm_staged_kernel<4, 4>, ST_NT, 0) != cudaSuccess || per_sm < 1)) per_sm = 2;
        const long long nblocks = (n + ST_RB - 1) / ST_RB, cap = (long long)MAUS_SM_COUNT_B200 * per_sm;
        const unsigned gx = (unsigned)(nblocks < cap ? nblocks : cap);
        return spmm_launch_groups(P, p_gstride, groups, stream, [&](int g0, int ng) {
            csr_spmm_staged_kernel<4, 4><<<dim3(gx, (unsigned)ng), ST_NT, 0, stream>>>(rowptr, colidx, vals, P + (long long)g0 * p_gstride, p_gstride,
                                                                                    Y, ldy, n, c0 + 4 * g0, ctotal);
        });
    }
    const unsigned gx = spmm_pipe_grid<4, 3, 4>(n);
    return spmm_launch_groups(P, p_gstride, groups, stream, [&](int g0, int ng) {
        csr_spmm_packed_kernel<4, 3, 4><<<dim3(gx, (unsigned)ng), SP_NT, 0, stream>>>(rowptr, colidx, vals, P + (long long)g0 * p_gstride, p_gstride,
                                                                                    Y, ldy, n, c0 + 4 * g0, ctotal);
    });
}

cudaError_t csr_spmm(const long long* rowptr, const int* colidx, const cplx* vals, const cplx* V, long long ldv, cplx* Y,
                     long long ldy, long long n, long long ncols, int C, cplx* pack_ws, cudaStream_t stream) {
    const long long threads = n * SP_LANES;
    const unsigned grid = (unsigned)((threads + SP_NT - 1) / SP_NT);
    const unsigned pgrid = (unsigned)((ncols + 255) / 256);
    if (C <= 0) return cudaSuccess;
    // one candidate: latency bound on the dependent index -> gather chain, so occupancy beats unrolling (SP_U = 1, <= 32
    // registers, 8 CTAs / SM; measured 0.132 vs 0.140 ms at n = 1M)
    if (C == 1) {
        static int pipe1 = -1;               // MAUS_SPMM_PIPE1=0: the one-row-at-a-time kernel (A/B measurements)
        if (pipe1 < 0) { const char* e = getenv("MAUS_SPMM_PIPE1"); pipe1 = e ? (atoi(e) != 0) : 1; }
        if (pipe1) csr_spmm_packed_kernel<1, 3, 4><<<dim3(spmm_pipe_grid<1, 3, 4>(n), 1), SP_NT, 0, stream>>>(rowptr, colidx, vals, V, 0, Y, ldy, n, 0, 1);
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
        csr_spmm_packed_kernel<2, 3, 4><<<dim3(spmm_pipe_grid<2, 3, 4>(n), 1), SP_NT, 0, stream>>>(rowptr, colidx, vals, pack_ws, 0, Y, ldy, n, 0, C);
        return cudaGetLastError();
    }
    const int groups = (C + 3) / 4;
    spmm_pack_kernel<4><<<dim3(pgrid, groups), 256, 0, stream>>>(V, ldv, pack_ws, ncols, 0, C);
    return csr_spmm_packed4(rowptr, colidx, vals, pack_ws, ncols * 4, Y, ldy, n, 0, C, groups, stream);
}

// number of complex elements of the interleaved copy csr_spmm needs for C candidates
size_t csr_spmm_pack_elems(long long ncols, int C) { return C <= 1 ? 0 : (size_t)ncols * 4 * (size_t)((C + 3) / 4); }
