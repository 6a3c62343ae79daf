// lu.cu -- batched complex128 LU with partial pivoting for sm_100a (replaces LAPACK zgesv behind AMS:59).
//
// Right-looking blocked LU on AUGMENTED systems W_b = [H_b | rhs_b] (column-major, ld = n, n+1 columns): the
// right-hand side rides along as column n, so after the last step it holds y = L^-1 P rhs and only the
// back-substitution with U remains; L is never needed again, which is why the row permutation is applied to the
// columns at and right of the current outer block only.  The schedule (two-level blocking: 4 panels per outer block,
// bulk trailing update with K = 512) lives in maus_lu_solve (maus_api.cu); the kernels of one panel step are:
//
// Per panel (width 128):
//   lu_panel        one thread-block CLUSTER per candidate (up to 8 CTAs x 512 threads); every thread owns one or
//                   two rows of the panel and keeps a 16 (8) column inner block of them in registers.  Pivot search
//                   (BLAS izamax metric |re|+|im|, first maximum) = REDUX on integer keys inside a warp -> CTA -> the
//                   CTAs' records exchanged with st.async into every CTA's shared memory, completing transaction bytes
//                   on the receivers' mbarriers (no cluster barrier in the column loop).  Pivoting is IMPLICIT inside the
//                   panel (rows are marked, not moved); the net permutation is emitted as <= 256 (dst, src) pairs.
//   lu_permute_rows applies those pairs to every column >= the outer block's first column in one parallel pass (no
//                   sequential swap chain).
//   lu_trtri        inverse of the unit-lower 128 x 128 diagonal block, so that the triangular solve for U12 and the
//                   trailing update are both plain GEMMs on the FP64 tensor pipe (zgemm.cu).
// After the last panel: lu_backsolve (one CTA per candidate) or lu_backsolve_cluster (2 - 8 CTAs per candidate, small batches).
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include "lu.cuh"
#include "../../include/maus_b200.h"

namespace cg = cooperative_groups;

namespace {

// ------------------------------------------------------------------------------------------------------------
// H build: fused shift + Psi regulariser (AMS:44-52, 270) while copying the shared matrix into each workspace
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lu_build_aug_kernel(cplx* W, long long strideW, int n, int batch,
                                                           const cplx* __restrict__ Acm, const cplx* __restrict__ sigma,
                                                           const double* __restrict__ psi,
                                                           const unsigned long long* __restrict__ keys,
                                                           const cplx* __restrict__ R_cm, const cplx* __restrict__ rhs,
                                                           long long rhs_stride, int always_draw, int conj_in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= n) return;
    const long long off = i + (long long)j * n;
    if (j == n) {
        for (int b = 0; b < batch; ++b) W[b * strideW + off] = rhs[b * rhs_stride + i];
        return;
    }
    cplx a = Acm[off];
    if (conj_in) a.y = -a.y;
    for (int b = 0; b < batch; ++b) {
        // same association as the reference: T = A - lambda*I (AMS:270); reg = psi*I + perturb (AMS:50); H = T + reg (AMS:52)
        cplx h = a;
        const double ps = psi[b];
        if (i == j) { h.x -= sigma[b].x; h.y -= sigma[b].y; }
        cplx reg = cmake((i == j) ? ps : 0.0, 0.0);
        if (R_cm) { reg.x += R_cm[off].x; reg.y += R_cm[off].y; }
        else if (keys) {
            // |r.x|, |r.y| <= 0.075*psi.  If |reg| + 0.075*psi is below half an ulp (2^-54 |t|) of BOTH components of
            // T = A - sigma*I, then fl(T + fl(reg + r)) == T == fl(T + reg) whatever r is: the Philox draw cannot change
            // a single bit and is skipped (exact, not an approximation; on the diagonal reg = psi).
            const double bound = (fabs(reg.x) + 0.075 * fabs(ps)) * 18014398509481984.0;      // * 2^54
            if (always_draw || !(bound < fabs(h.x) && bound < fabs(h.y))) {
                cplx r = psi_perturbation(keys[b], (uint32_t)i, (uint32_t)j, ps); reg.x += r.x; reg.y += r.y;
            }
        }
        h.x += reg.x; h.y += reg.y;
        W[b * strideW + off] = h;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Panel factorisation
// ------------------------------------------------------------------------------------------------------------
constexpr int PANEL_MAXC = 8;

// MAUS_PANEL_PROF (profiling build only, `make PROF=1`): per-phase clock64() sums of thread 0 of cluster rank 0 of candidate 0
#ifdef MAUS_PANEL_PROF
__device__ unsigned long long g_panel_prof[16];
#define PP_DECL __shared__ long long pp_acc[16]; long long pp_last = 0; const bool pp_on = (blockIdx.x == 0 && threadIdx.x == 0); \
    if (pp_on) { for (int i_ = 0; i_ < 16; ++i_) pp_acc[i_] = 0; pp_last = clock64(); }
#define PP(i) do { if (pp_on) { long long t_ = clock64(); pp_acc[i] += t_ - pp_last; pp_last = t_; } } while (0)
#define PP_FLUSH do { if (pp_on) { for (int i_ = 0; i_ < 16; ++i_) atomicAdd(&g_panel_prof[i_], (unsigned long long)pp_acc[i_]); atomicAdd(&g_panel_prof[15], 1ULL); } } while (0)
#else
#define PP_DECL
#define PP(i) do { } while (0)
#define PP_FLUSH do { } while (0)
#endif

// One 16-byte-granular record per CTA and column: the CTA's best row, published into every CTA of the cluster.
template <int IB>
struct __align__(16) PanelSlot {
    unsigned long long key;   // pivot key of the CTA's best live row (pivot_key below); 0 = the CTA has no live row
    int row;                  // panel-relative row index
    int pad;
    cplx recip;               // 1 / pivot
    cplx data[IB];            // that row's window (columns j .. j + IB - 1 of the inner block)
};

// Pivot order without FP64 compares: |re| + |im| >= +0 (NaN mapped to +inf), and non-negative doubles order like their bit
// patterns.  key = bits + 1 so that 0 means "no live row"; key == 1 is an exact zero pivot.
__device__ __forceinline__ unsigned long long pivot_key(cplx a) {
    double v = cabs1(a);
    if (v != v) v = INFINITY;
    return (unsigned long long)__double_as_longlong(v) + 1ULL;
}
// 1 / a with ONE division (Smith's formula in crecip costs three dependent FP64 divisions, ~440 cycles on the owner's critical
// path of every column; this one ~200): scale by an exact power of two so that |a|^2 can neither overflow nor underflow,
// then conj(a') / |a'|^2.  Exponent fields 0 (zero / subnormal) and 2047 (inf / nan) take the robust path.
__device__ __forceinline__ cplx pivot_recip(cplx a) {
    const double mx = fmax(fabs(a.x), fabs(a.y));
    const int e = (int)((__double_as_longlong(mx) >> 52) & 0x7ff);
    if (e == 0 || e >= 2046) return crecip(a);
    const double sc = __longlong_as_double((long long)(2046 - e) << 52);       // 2^(1023 - e): max(|x'|, |y'|) in [1, 2)
    const double xs = a.x * sc, ys = a.y * sc;
    const double inv = 1.0 / fma(xs, xs, ys * ys);
    return cmake((xs * inv) * sc, (-ys * inv) * sc);
}
// warp-wide (max key, then min row) with three REDUX instead of five dependent shuffle rounds of 64-bit values
__device__ __forceinline__ void warp_best(unsigned long long& key, int& row) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    bool c = hi == mh;
    const unsigned ml = __reduce_max_sync(0xffffffffu, c ? lo : 0u);
    c = c && lo == ml;
    const unsigned mr = __reduce_min_sync(0xffffffffu, c ? (unsigned)row : 0x7fffffffu);
    key = ((unsigned long long)mh << 32) | ml; row = (int)mr;
}
__device__ __forceinline__ uint32_t pn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pn_mapa(uint32_t addr, int cta) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta)); return r;
}
// 16-byte store into a peer CTA's shared memory that also counts 16 bytes on the peer's mbarrier (SASS STAS.128): data and
// "it has arrived" travel together, no separate flag or fence
__device__ __forceinline__ void pn_st_async16(uint32_t dst, double2 v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
                 ::"r"(dst), "d"(v.x), "d"(v.y), "r"(bar) : "memory");
}
// default semantics (acquire at CTA scope) as in every TMA / cluster-barrier consumer: the records arrive in THIS CTA's shared
// memory through the barrier's own transaction count; acquire.cluster would add an L1 invalidation (CCTL.IVALL) per column
__device__ __forceinline__ void pn_bar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t a = pn_smem_u32(bar);
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}

// MINB = CTAs per SM the register allocation is held to, UW = columns per step of the in-leaf rank-IB update (the update keeps
// 2 UW columns of prefetch + UW in flight per row: UW = 4 needs ~128 registers)
//
// Column step (the kernel is bound by the LATENCY of this chain, 2 x jb x panels times per factorisation; round-2 clocks:
// 4 900 cycles per column with shuffle reductions on doubles, a serial slot scan and cg::cluster.sync):
//   thread best -> warp_best (REDUX) -> smem -> CTA barrier -> warp_best over the warps' results
//   -> the owner warp sends the CTA's record (key, row, 1/pivot, window) to EVERY CTA of the cluster with st.async, which
//      completes transaction bytes on the receiver's mbarrier: the exchange is the synchronisation, there is no cluster
//      barrier in the column loop; records and mbarriers are double-buffered by column parity
//   -> wait on the own mbarrier -> warp_best over the <= 8 records -> scale, park the multiplier in shared memory (lbuf),
//      rank-1 update of the register window.
template <int R, int IB, int PANEL_NT, int MINB, int UW>
__global__ void __launch_bounds__(PANEL_NT, MINB)
lu_panel_kernel(cplx* W, long long strideW, int n, int k0, int jb, LuPairs* pairs, int* info) {
    cg::cluster_group cluster = cg::this_cluster();
    const int NC = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / NC;
    const int T = NC * PANEL_NT;
    const int tg = rank * PANEL_NT + threadIdx.x;
    const int m = n - k0;
    const int ld = n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = PANEL_NT / 32;
    constexpr int S16 = (int)(sizeof(PanelSlot<IB>) / 16);     // 16-byte pieces of a record
    static_assert(NW <= 32 && PANEL_MAXC <= 32, "second-level reductions run inside one warp");
    cplx* P = W + (long long)b * strideW + (long long)k0 * ld + k0;   // P[r + c*ld]

    extern __shared__ __align__(16) unsigned char panel_dyn[];
    cplx* lbuf = reinterpret_cast<cplx*>(panel_dyn);          // multipliers of the current inner block: [R][IB][PANEL_NT]
    __shared__ PanelSlot<IB> slots[2][PANEL_MAXC];
    __shared__ PanelSlot<IB> pub;                             // owner lane -> owner warp
    __shared__ __align__(8) uint64_t xbar[2];
    __shared__ unsigned long long wkey[2][NW];
    __shared__ int wrow[2][NW];
    __shared__ cplx L11[IB][IB + 1];
    __shared__ cplx U12[IB][LU_NB];
    __shared__ int piv[LU_NB];
    __shared__ unsigned char is_piv[LU_NB];

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pn_smem_u32(&xbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pn_smem_u32(&xbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();      // every CTA's mbarriers exist before the first remote transaction can arrive

    int row[R];
    bool valid[R], done[R];
#pragma unroll
    for (int q = 0; q < R; ++q) { row[q] = tg + q * T; valid[q] = row[q] < m; done[q] = false; }
    bool zero_seen = false;
    int zero_col = 0;
    PP_DECL

    for (int ib0 = 0; ib0 < jb; ib0 += IB) {
        const int ibw = min(IB, jb - ib0);
        // a[q][i] always holds column (current j) + i of row q: after every column the window is shifted left by one,
        // so the column loop below is a ROLLED loop with constant register indices (the unrolled version was 480 KB of
        // SASS and instruction-fetch bound)
        cplx a[R][IB];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const bool live = valid[q] && !done[q];
#pragma unroll
            for (int i = 0; i < IB; ++i)
                a[q][i] = (live && i < ibw) ? P[row[q] + (long long)(ib0 + i) * ld] : cmake(0.0, 0.0);
        }
        PP(0);
#pragma unroll 1
        for (int j = 0; j < ibw; ++j) {
            const int col = ib0 + j, buf = col & 1;
            const uint32_t par = (uint32_t)((col >> 1) & 1);
            // ---- pivot search: thread -> warp -> CTA.  The "row" that travels through the min-reductions carries its owner in
            // the low bits (row << 2 | q inside a warp, then << 5 | warp inside the CTA, row << 4 | CTA rank in the cluster):
            // rows are unique, so the order is the row order, and nobody has to divide by the cluster's thread count ----
            unsigned long long key = 0ULL; int myw = 0x7fffffff;
#pragma unroll
            for (int q = 0; q < R; ++q)
                if (valid[q] && !done[q]) {
                    const unsigned long long kq = pivot_key(a[q][0]);
                    const int wq = (row[q] << 2) | q;
                    if (kq > key || (kq == key && wq < myw)) { key = kq; myw = wq; }
                }
            int bw = myw;
            warp_best(key, bw);
            if (lane == 0) { wkey[buf][warp] = key; wrow[buf][warp] = (key != 0ULL) ? ((bw << 5) | warp) : 0x7fffffff; }
            // this column's transactions: one record from every CTA of the cluster (armed before anybody can wait on it)
            if (threadIdx.x == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                             ::"r"(pn_smem_u32(&xbar[buf])), "r"((uint32_t)(NC * sizeof(PanelSlot<IB>))) : "memory");
            __syncthreads();
            PP(1);
            unsigned long long ck = (lane < NW) ? wkey[buf][lane] : 0ULL;
            int cw = (lane < NW) ? wrow[buf][lane] : 0x7fffffff;
            warp_best(ck, cw);
            // ---- the warp that owns the CTA's best row (warp 0 when the CTA has no live row) sends the record ----
            {
                const bool have = ck != 0ULL;
                const int owner_warp = have ? (cw & 31) : 0;
                if (warp == owner_warp) {
                    const int ow = cw >> 5;                              // (row << 2) | q of the CTA's best row
                    if (have ? (myw == ow) : (lane == 0)) {
                        const int owner_q = have ? (ow & 3) : 0;
                        pub.key = ck; pub.row = have ? (((ow >> 2) << 4) | rank) : 0x7fffffff; pub.pad = 0;
#pragma unroll
                        for (int i = 0; i < IB; ++i) {
                            cplx sel = a[0][i];
#pragma unroll
                            for (int q = 1; q < R; ++q) if (owner_q == q) sel = a[q][i];
                            pub.data[i] = sel;
                            if (i == 0) pub.recip = have ? pivot_recip(sel) : cmake(0.0, 0.0);
                        }
                    }
                    __syncwarp();
                    const double2* src = reinterpret_cast<const double2*>(&pub);
                    const uint32_t slot0 = pn_smem_u32(&slots[buf][rank]), bar0 = pn_smem_u32(&xbar[buf]);
                    for (int idx = lane; idx < NC * S16; idx += 32) {
                        const int d = idx / S16, e = idx - d * S16;
                        pn_st_async16(pn_mapa(slot0 + 16u * (uint32_t)e, d), src[e], pn_mapa(bar0, d));
                    }
                    // pub is rewritten by the next column's owner after the next CTA barrier, which this warp reaches
                    // only after the loads above have returned
                }
            }
            PP(2);
            pn_bar_wait(&xbar[buf], par);
            PP(3);
            unsigned long long wk = (lane < NC) ? slots[buf][lane].key : 0ULL;
            int ww = (lane < NC) ? slots[buf][lane].row : 0x7fffffff;
            warp_best(wk, ww);
            const int wi = (wk != 0ULL) ? (ww & 15) : 0;                 // CTA whose record won
            const int wr = (wk != 0ULL) ? (ww >> 4) : 0x7fffffff;        // the pivot row
            const PanelSlot<IB>& ws = slots[buf][wi];
            if (threadIdx.x == 0) piv[col] = wr;
            const bool zero = wk <= 1ULL;                                 // no live row at all, or the best |a| is exactly 0
            if (zero && !zero_seen) { zero_seen = true; zero_col = k0 + col + 1; }
            const cplx rc = ws.recip;
#pragma unroll
            for (int q = 0; q < R; ++q)
                if (valid[q] && !done[q]) {
                    cplx* prow = P + row[q] + (long long)col * ld;
                    if (row[q] == wr) {
                        done[q] = true;                               // my window holds U(j, j..): final values
#pragma unroll
                        for (int i = 0; i < IB; ++i)
                            if (j + i < ibw) prow[(long long)i * ld] = a[q][i];
                    } else {
                        cplx l = zero ? a[q][0] : cmul(a[q][0], rc);
                        prow[0] = l;                                   // multiplier, stays in its physical row
                        if (zero) l = cmake(0.0, 0.0);
                        lbuf[(q * IB + j) * PANEL_NT + threadIdx.x] = l;   // ... and a copy for the block update below
#pragma unroll
                        for (int i = 0; i + 1 < IB; ++i) { cplx x = a[q][i + 1]; cfms(x, l, ws.data[i + 1]); a[q][i] = x; }
                        a[q][IB - 1] = cmake(0.0, 0.0);
                    }
                }
            PP(4);
        }
        const int rest = jb - (ib0 + ibw);
        if (rest > 0) {
            cluster.sync();     // multipliers / U entries written in the column loop are visible cluster-wide
            // L11 = multipliers of the block's pivot rows w.r.t. the earlier pivots of the block
            for (int idx = threadIdx.x; idx < IB * IB; idx += PANEL_NT) {
                const int jj = idx / IB, i = idx % IB;
                if (i < jj && jj < ibw) L11[jj][i] = __ldcg(&P[piv[ib0 + jj] + (long long)(ib0 + i) * ld]);
            }
            __syncthreads();
            PP(5);
            // (i) U12 block = L11^-1 * (pivot rows, remaining panel columns); every CTA computes its own copy.  All loads of
            // a column are issued before the first dependent FMA (one L2 round trip instead of ibw)
            if ((int)threadIdx.x < rest) {
                const int c = threadIdx.x;
                const long long coff = (long long)(ib0 + ibw + c) * ld;
                cplx xs[IB];
#pragma unroll
                for (int j = 0; j < IB; ++j) xs[j] = (j < ibw) ? __ldcg(&P[piv[ib0 + j] + coff]) : cmake(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < IB; ++j) {
                    if (j < ibw) {
                        cplx x = xs[j];
#pragma unroll
                        for (int i = 0; i < j; ++i) cfms(x, L11[j][i], xs[i]);
                        xs[j] = x;
                        U12[j][c] = x;
                    }
                }
            }
            PP(6);
            cluster.sync();   // every CTA has read the pivot rows before rank 0 overwrites them with U12
            PP(7);
            if (rank == 0 && (int)threadIdx.x < rest) {
                const long long coff = (long long)(ib0 + ibw + threadIdx.x) * ld;
                for (int j = 0; j < ibw; ++j) P[piv[ib0 + j] + coff] = U12[j][threadIdx.x];
            }
            // (ii) rank-ibw update of the remaining panel columns for the rows that are still live.  The R rows of a thread
            // go through the loop TOGETHER, UW columns at a time: R * UW * 2 independent FMA chains (a DFMA result is ready
            // after ~21 cycles) and every U12 entry fetched from shared memory serves R rows; the multipliers are read from
            // lbuf as they are needed instead of occupying 4 IB registers per row; the next UW columns of every row are
            // prefetched while the current ones are updated (PFD groups deep: the rows come from L2 / HBM, ~1 us away)
            {
                constexpr int PFD = (R == 1) ? 2 : 1;
                bool lv[R];
                cplx* prow[R];
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    lv[q] = valid[q] && !done[q];
                    prow[q] = P + row[q] + (long long)(ib0 + ibw) * ld;
                    if (ibw < IB)      // ragged last block: the multipliers of the missing columns are zero
                        for (int j = ibw; j < IB; ++j) lbuf[(q * IB + j) * PANEL_NT + threadIdx.x] = cmake(0.0, 0.0);
                }
                const int rest4 = (rest / UW) * UW;
                cplx nx[PFD][R][UW];
#pragma unroll
                for (int d = 0; d < PFD; ++d)
#pragma unroll
                    for (int q = 0; q < R; ++q)
#pragma unroll
                        for (int u = 0; u < UW; ++u)
                            nx[d][q][u] = (lv[q] && d * UW < rest4) ? prow[q][(long long)(d * UW + u) * ld] : cmake(0.0, 0.0);
#pragma unroll 1
                for (int c = 0; c < rest4; c += UW) {
                    cplx x[R][UW];
#pragma unroll
                    for (int q = 0; q < R; ++q)
#pragma unroll
                        for (int u = 0; u < UW; ++u) {
                            x[q][u] = nx[0][q][u];
#pragma unroll
                            for (int d = 0; d + 1 < PFD; ++d) nx[d][q][u] = nx[d + 1][q][u];
                        }
                    if (c + PFD * UW < rest4) {
#pragma unroll
                        for (int q = 0; q < R; ++q)
#pragma unroll
                            for (int u = 0; u < UW; ++u)
                                if (lv[q]) nx[PFD - 1][q][u] = prow[q][(long long)(c + PFD * UW + u) * ld];
                    }
#pragma unroll
                    for (int j = 0; j < IB; ++j) {
                        cplx uj[UW];
#pragma unroll
                        for (int u = 0; u < UW; ++u) uj[u] = U12[j][c + u];
#pragma unroll
                        for (int q = 0; q < R; ++q) {
                            const cplx lj = lbuf[(q * IB + j) * PANEL_NT + threadIdx.x];
#pragma unroll
                            for (int u = 0; u < UW; ++u) cfms(x[q][u], lj, uj[u]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < R; ++q)
                        if (lv[q]) {
#pragma unroll
                            for (int u = 0; u < UW; ++u) prow[q][(long long)(c + u) * ld] = x[q][u];
                        }
                }
                for (int c = rest4; c < rest; ++c)
#pragma unroll
                    for (int q = 0; q < R; ++q)
                        if (lv[q]) {
                            cplx x = prow[q][(long long)c * ld];
#pragma unroll
                            for (int j = 0; j < IB; ++j) cfms(x, lbuf[(q * IB + j) * PANEL_NT + threadIdx.x], U12[j][c]);
                            prow[q][(long long)c * ld] = x;
                        }
            }
            // Updates written in (ii) are read by other CTAs in the next block's step (i): the cluster barrier at the top of
            // that step (release / acquire) orders them, and the reads use ld.cg.
            __syncthreads();   // U12 / L11 are rewritten by the next block
            PP(8);
        }
    }
    __syncthreads();
    PP_FLUSH;
    // ---- net row permutation of this panel (relative to k0) ----
    if (rank == 0) {
        if (threadIdx.x < LU_NB) is_piv[threadIdx.x] = 0;
        __syncthreads();
        if ((int)threadIdx.x < jb && piv[threadIdx.x] < jb) is_piv[piv[threadIdx.x]] = 1;
        __syncthreads();
        if (threadIdx.x == 0) {
            LuPairs* pr = pairs + b;
            int cnt = 0, d = 0;   // d walks the non-pivot rows of the top block in increasing order
            for (int j = 0; j < jb; ++j) {
                const int pj = piv[j];
                if (pj != j) { pr->dst[cnt] = j; pr->src[cnt] = pj; ++cnt; }
                if (pj >= jb) {
                    while (is_piv[d]) ++d;
                    pr->dst[cnt] = pj; pr->src[cnt] = d; ++cnt; ++d;
                }
            }
            pr->count = cnt;
            if (zero_seen && info[b] == 0) info[b] = zero_col;
        }
    }
    cluster.sync();   // no CTA may exit while a peer could still address its shared memory
}

// ------------------------------------------------------------------------------------------------------------
// Row permutation of all columns >= k0
// ------------------------------------------------------------------------------------------------------------
constexpr int PERM_COLS = 8;
__global__ void __launch_bounds__(256) lu_permute_rows_kernel(cplx* W, long long strideW, int n, int k0, int colstart,
                                                              const LuPairs* __restrict__ pairs) {
    __shared__ cplx tmp[LU_MAX_PAIRS * PERM_COLS];
    __shared__ int sdst[LU_MAX_PAIRS], ssrc[LU_MAX_PAIRS];
    const int b = blockIdx.y;
    const int cnt = pairs[b].count;
    if (cnt == 0) return;
    for (int q = threadIdx.x; q < cnt; q += blockDim.x) { sdst[q] = pairs[b].dst[q]; ssrc[q] = pairs[b].src[q]; }
    __syncthreads();
    const int col0 = colstart + blockIdx.x * PERM_COLS;
    const int ncol = min(PERM_COLS, n + 1 - col0);
    cplx* Wb = W + (long long)b * strideW + k0;
    const int total = cnt * ncol;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int c = idx / cnt, q = idx - c * cnt;
        tmp[idx] = Wb[ssrc[q] + (long long)(col0 + c) * n];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int c = idx / cnt, q = idx - c * cnt;
        Wb[sdst[q] + (long long)(col0 + c) * n] = tmp[idx];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Inverse of the unit-lower-triangular diagonal block
// ------------------------------------------------------------------------------------------------------------
// thread c owns column c of X = L11^-1; X is kept packed (row p holds columns 0..p) in shared memory.  Rows of L11 are
// streamed through a ring of TRTRI_PF row buffers with cp.async so the global-load latency of row r + TRTRI_PF - 1
// hides behind the arithmetic of rows r .. r + TRTRI_PF - 2 (the one-row-ahead version was latency-bound: 353 us).
//
// The row recurrence is sequential (two CTA barriers per row), so the block is split into 4 (or 2) diagonal blocks of h rows:
//   phase A  inverts all diagonal blocks SIMULTANEOUSLY (thread column c works on row rl of its own block in the same
//            iteration): h - 1 instead of jb - 1 sequential steps;
//   phase B  fills the off-diagonal blocks level by level, X21 = -X22 (L21 X11): triangular block products spread over the
//            threads (T = L21 X11 goes into X21's slots, then X21 row by row from the bottom up, in place).
constexpr int TRTRI_PF = 8;
constexpr int TRTRI_SPLIT = 4;     // lanes sharing one column's dot product (critical path / 4)
__global__ void __launch_bounds__(LU_NB * TRTRI_SPLIT) lu_trtri_kernel(const cplx* __restrict__ W, long long strideW, int n, int k0,
                                                         int jb, cplx* __restrict__ Linv) {
    extern __shared__ __align__(16) unsigned char trtri_smem[];
    cplx* X = reinterpret_cast<cplx*>(trtri_smem);                 // jb*(jb+1)/2
    cplx* rowbuf = X + (LU_NB * (LU_NB + 1)) / 2;                  // TRTRI_PF x LU_NB ring of rows of L
    const int b = blockIdx.x, c = threadIdx.x / TRTRI_SPLIT, sp = threadIdx.x % TRTRI_SPLIT;
    const cplx* L = W + (long long)b * strideW + (long long)k0 * n + k0;   // L[r + p*n]
    // diagonal blocks inverted simultaneously: 4 (jb a multiple of 4, >= 64), 2 (jb even, >= 32) or 1
    const int nblk = (jb >= 64 && (jb & 3) == 0) ? 4 : ((jb >= 32 && (jb & 1) == 0) ? 2 : 1);
    const int h = jb / nblk;                                       // diagonal block size
    const int base = (c < jb) ? (c / h) * h : jb;                  // first row / column of this thread's diagonal block
    if (c < jb && sp == 0) X[(c * (c + 1)) / 2 + c] = cmake(1.0, 0.0);
    // ring slot entry c holds L[base + rl][c]: row rl of the diagonal block that column c belongs to
    auto fetch_row = [&](int rl) {
        const int r = base + rl;
        if (rl < h && r < jb && c < r && sp == 0) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&rowbuf[(rl % TRTRI_PF) * LU_NB + c]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(L + r + (long long)c * n) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int rl = 1; rl < TRTRI_PF; ++rl) fetch_row(rl);
    for (int rl = 1; rl < h; ++rl) {
        fetch_row(rl + TRTRI_PF - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(TRTRI_PF - 1) : "memory");     // row rl has landed (for this thread)
        __syncthreads();                                                             // ... and for every thread
        const cplx* lr = rowbuf + (rl % TRTRI_PF) * LU_NB;
        const int r = base + rl;
        {
            cplx acc = cmake(0.0, 0.0);
            if (c < r && r < jb)
                for (int p = c + sp; p < r; p += TRTRI_SPLIT) cfms(acc, lr[p], X[(p * (p + 1)) / 2 + c]);
#pragma unroll
            for (int o = TRTRI_SPLIT / 2; o > 0; o >>= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            }
            if (c < r && r < jb && sp == 0) X[(r * (r + 1)) / 2 + c] = acc;
        }
        __syncthreads();       // X row complete; ring slot rl % PF may be refilled by the next fetch
    }
    // ---- phase B: off-diagonal blocks, level by level: X21 = -X22 * (L21 * X11) for the pair of hh x hh blocks at offset o ----
    // executed by the nt threads [t0, t0 + nt) (warp-aligned); both __syncthreads are reached by every thread of the CTA.
    // T = L21 X11 goes into its own buffer Tb (hh x (hh + 1), aliases the row ring of phase A), so the second product is a plain
    // triangular matrix product spread over all threads (round 2: it ran in place, row by row from the bottom up -- hh sequential
    // steps of an 8-lane dot product, 35 k of the kernel's 165 k cycles at jb = 128)
    auto pair_product = [&](int o, int hh, int t0, int nt, cplx* Tb) {
        const int tl = (int)threadIdx.x - t0;
        const bool mine = tl >= 0 && tl < nt;
        const int ldt = hh + 1;
        // T[i][j] = sum_{p = j}^{hh-1} L21[i][p] X11[p][j]; thread tl: row i = tl % hh, columns j = tl / hh + u * (nt / hh)
        if (mine) {
            const int i = tl % hh;
            const cplx* l21 = L + (o + hh + i) + (long long)o * n;     // L21[i][p] = l21[p * n] (consecutive threads, consecutive rows)
            for (int j = tl / hh; j < hh; j += nt / hh) {
                cplx acc = cmake(0.0, 0.0);
#pragma unroll 4
                for (int p = j; p < hh; ++p) cfma(acc, __ldg(&l21[(long long)p * n]), X[((o + p) * (o + p + 1)) / 2 + o + j]);
                Tb[i * ldt + j] = acc;
            }
        }
        __syncthreads();
        // X21[i][j] = -sum_{q <= i} X22[i][q] T[q][j]; thread tl: column j = tl % hh (T and X21 contiguous over the warp, X22[i][q] a
        // broadcast), rows i = tl / hh + u * (nt / hh); two accumulators halve the dependent FMA chain
        if (mine) {
            const int j = tl % hh;
            for (int i = tl / hh; i < hh; i += nt / hh) {
                const int ri = o + hh + i;
                const cplx* x22 = X + (ri * (ri + 1)) / 2 + o + hh;              // X22[i][q], q <= i
                cplx a0 = cmake(0.0, 0.0), a1 = cmake(0.0, 0.0);
                int q = 0;
                for (; q + 1 <= i; q += 2) { cfma(a0, x22[q], Tb[q * ldt + j]); cfma(a1, x22[q + 1], Tb[(q + 1) * ldt + j]); }
                if (q <= i) cfma(a0, x22[q], Tb[q * ldt + j]);
                X[(ri * (ri + 1)) / 2 + o + j] = cmake(-(a0.x + a1.x), -(a0.y + a1.y));
            }
        }
        __syncthreads();
    };
    {
        const int NT = LU_NB * TRTRI_SPLIT;
        cplx* Tb = rowbuf;                                               // phase A is over: the row ring is free
        if (nblk == 4) {
            // level 1: the two pairs of h-blocks side by side (half of the CTA each), level 2: the pair of 2h-blocks
            const int upper = ((int)threadIdx.x < NT / 2) ? 0 : 1;        // one call site: every thread passes the same barriers
            pair_product(upper * 2 * h, h, upper * (NT / 2), NT / 2, Tb + upper * (h * (h + 1)));
            pair_product(0, 2 * h, 0, NT, Tb);
        } else if (nblk == 2) {
            pair_product(0, h, 0, NT, Tb);
        }
    }
    cplx* out = Linv + (long long)b * LU_NB * LU_NB;
    for (int r = sp; r < jb; r += TRTRI_SPLIT)
        if (c < jb) out[r + c * LU_NB] = (c <= r) ? X[(r * (r + 1)) / 2 + c] : cmake(0.0, 0.0);
}

// ------------------------------------------------------------------------------------------------------------
// Back substitution with U (one CTA per candidate; the right-hand side lives in shared memory)
// ------------------------------------------------------------------------------------------------------------
constexpr int BS_NT = 512;
constexpr int BS_BLK = 32;
__global__ void __launch_bounds__(BS_NT) lu_backsolve_kernel(const cplx* __restrict__ W, long long strideW, int n,
                                                             const int* __restrict__ info, cplx* __restrict__ X,
                                                             int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char bs_smem[];
    cplx* y = reinterpret_cast<cplx*>(bs_smem);          // n
    cplx* D = y + n;                                     // BS_BLK x (BS_BLK+1)
    __shared__ int bad;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const cplx* Wb = W + (long long)b * strideW;
    for (int i = tid; i < n; i += BS_NT) y[i] = Wb[i + (long long)n * n];
    if (tid == 0) bad = 0;
    __syncthreads();
    const int nblk = (n + BS_BLK - 1) / BS_BLK;
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int r0 = kb * BS_BLK, bs = min(BS_BLK, n - r0);
        for (int idx = tid; idx < bs * bs; idx += BS_NT) {
            int i = idx % bs, j = idx / bs;
            D[i * (BS_BLK + 1) + j] = Wb[(r0 + i) + (long long)(r0 + j) * n];
        }
        __syncthreads();
        if (warp == 0) {
            cplx yi = (lane < bs) ? y[r0 + lane] : cmake(0.0, 0.0);
            for (int j = bs - 1; j >= 0; --j) {
                cplx yj = cmake(__shfl_sync(0xffffffffu, yi.x, j), __shfl_sync(0xffffffffu, yi.y, j));
                cplx xj = cdiv(yj, D[j * (BS_BLK + 1) + j]);
                if (lane == j) yi = xj;
                else if (lane < j) cfms(yi, D[lane * (BS_BLK + 1) + j], xj);
            }
            if (lane < bs) y[r0 + lane] = yi;
        }
        __syncthreads();
        for (int i = tid; i < r0; i += BS_NT) {
            cplx acc = y[i];
            const cplx* u = Wb + i + (long long)r0 * n;
            if (bs == BS_BLK) {
                // all 32 loads of the row segment in flight before the first FMA (the kernel is bound by HBM latency x
                // bytes in flight per SM: one CTA per SM streams the whole upper triangle)
                cplx uu[BS_BLK];
#pragma unroll
                for (int j = 0; j < BS_BLK; ++j) uu[j] = __ldcs(&u[(long long)j * n]);
#pragma unroll
                for (int j = 0; j < BS_BLK; ++j) cfms(acc, uu[j], y[r0 + j]);
            } else {
                for (int j = 0; j < bs; ++j) cfms(acc, __ldg(&u[(long long)j * n]), y[r0 + j]);
            }
            y[i] = acc;
        }
        __syncthreads();
    }
    int mybad = 0;
    for (int i = tid; i < n; i += BS_NT) {
        cplx v = y[i];
        if (!cfinite(v)) mybad = 1;
        X[(long long)b * n + i] = v;
    }
    if (mybad) bad = 1;
    __syncthreads();
    if (tid == 0) {
        if (status[b] == 0) {
            if (info[b] != 0) status[b] = MAUS_ST_ZERO_PIVOT;
            else if (bad) status[b] = MAUS_ST_NONFINITE;
        }
    }
}

// Back substitution for SMALL batches: one CTA per candidate streams the whole upper triangle (n^2 / 2 entries) through a single
// SM, ~38 GB/s, 3.3 - 3.6 ms at n = 4096 however few candidates there are.  Here a CLUSTER of NC CTAs shares one candidate:
// the 32-row blocks are owned block-cyclically (block kb -> CTA kb % NC), every CTA keeps the y entries of its own blocks in
// shared memory, the owner of block kb solves its 32 x 32 triangle and writes x_kb into every CTA's shared memory (DSMEM), one
// cluster barrier per block step, then every CTA subtracts U[own rows, block kb] x_kb from its y.  Same arithmetic per row as
// lu_backsolve_kernel except that the diagonal is applied as a reciprocal (one division per diagonal entry, computed by 32
// lanes at once instead of a Smith division inside the 32-step dependent chain).
__global__ void __launch_bounds__(BS_NT) lu_backsolve_cluster_kernel(const cplx* __restrict__ W, long long strideW, int n,
                                                                     const int* __restrict__ info, cplx* __restrict__ X,
                                                                     int* __restrict__ status) {
    cg::cluster_group cluster = cg::this_cluster();
    const int NC = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) unsigned char bs_smem[];
    cplx* yloc = reinterpret_cast<cplx*>(bs_smem);        // [owned blocks][32]
    __shared__ cplx xb[2][BS_BLK];
    __shared__ cplx D[2][BS_BLK][BS_BLK + 1];              // diagonal blocks: the one being solved and the owner's next one
    __shared__ int bad[PANEL_MAXC];                        // rank 0 collects one flag per CTA
    const int b = blockIdx.x / NC, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const cplx* Wb = W + (long long)b * strideW;
    const int nblk = (n + BS_BLK - 1) / BS_BLK;
    const int nown = (nblk - rank + NC - 1) / NC;          // blocks rank, rank + NC, ...
    for (int idx = tid; idx < nown * BS_BLK; idx += BS_NT) {
        const int i = ((idx / BS_BLK) * NC + rank) * BS_BLK + (idx % BS_BLK);
        yloc[idx] = (i < n) ? Wb[i + (long long)n * n] : cmake(0.0, 0.0);
    }
    // the diagonal blocks are final before this kernel starts: the owner fetches its next one while it solves the current one
    auto fetch_D = [&](int kq, int slot, int t0, int nt) {
        if (kq < 0) return;
        const int q0 = kq * BS_BLK, qs = min(BS_BLK, n - q0);
        for (int idx = tid - t0; idx >= 0 && idx < qs * qs; idx += nt) {
            const int i = idx % qs, j = idx / qs;
            D[slot][i][j] = Wb[(q0 + i) + (long long)(q0 + j) * n];
        }
    };
    int dcur = 0;
    if (nown > 0) fetch_D((nown - 1) * NC + rank, 0, 0, BS_NT);
    int mybad = 0;
    __syncthreads();
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int r0 = kb * BS_BLK, bs = min(BS_BLK, n - r0), buf = kb & 1;
        if (kb % NC == rank) {
            if (warp == 0) {
                cplx yi = (lane < bs) ? yloc[(kb / NC) * BS_BLK + lane] : cmake(0.0, 0.0);
                const cplx rinv = (lane < bs) ? pivot_recip(D[dcur][lane][lane]) : cmake(0.0, 0.0);
                for (int j = bs - 1; j >= 0; --j) {
                    const cplx yj = cmake(__shfl_sync(0xffffffffu, yi.x, j), __shfl_sync(0xffffffffu, yi.y, j));
                    const cplx rj = cmake(__shfl_sync(0xffffffffu, rinv.x, j), __shfl_sync(0xffffffffu, rinv.y, j));
                    const cplx xj = cmul(yj, rj);
                    if (lane == j) yi = xj;
                    else if (lane < j) cfms(yi, D[dcur][lane][j], xj);
                }
                if (lane < bs) {
                    if (!cfinite(yi)) mybad = 1;
                    X[(long long)b * n + r0 + lane] = yi;
                    for (int d = 0; d < NC; ++d) *cluster.map_shared_rank(&xb[buf][lane], d) = yi;
                }
            } else {
                fetch_D(kb - NC, dcur ^ 1, 32, BS_NT - 32);
            }
            dcur ^= 1;
        }
        cluster.sync();          // x_kb has landed everywhere; everybody has finished reading the buffer it replaces (step kb + 2)
        // own row blocks above block kb: 16 of them per pass, one warp each, one row per lane
        for (int ob = warp; ob * NC + rank < kb; ob += BS_NT / 32) {
            const int i = (ob * NC + rank) * BS_BLK + lane;             // < r0 <= n - 1
            cplx acc = yloc[ob * BS_BLK + lane];
            const cplx* u = Wb + i + (long long)r0 * n;
            if (bs == BS_BLK) {
                cplx uu[BS_BLK];
#pragma unroll
                for (int j = 0; j < BS_BLK; ++j) uu[j] = __ldcs(&u[(long long)j * n]);
#pragma unroll
                for (int j = 0; j < BS_BLK; ++j) cfms(acc, uu[j], xb[buf][j]);
            } else {
                for (int j = 0; j < bs; ++j) cfms(acc, __ldg(&u[(long long)j * n]), xb[buf][j]);
            }
            yloc[ob * BS_BLK + lane] = acc;
        }
        __syncthreads();         // the owner of block kb - 1 reads its y entries next; D is rewritten
    }
    // one flag per CTA (any lane of any owned block saw a non-finite x), collected by rank 0
    const int cta_bad = __syncthreads_or(mybad);
    if (tid == 0) *cluster.map_shared_rank(&bad[rank], 0) = cta_bad ? 1 : 0;
    cluster.sync();
    if (rank == 0 && tid == 0) {
        int any = 0;
        for (int d = 0; d < NC; ++d) any |= bad[d];
        if (status[b] == 0) {
            if (info[b] != 0) status[b] = MAUS_ST_ZERO_PIVOT;
            else if (any) status[b] = MAUS_ST_NONFINITE;
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
#ifdef MAUS_PANEL_PROF
extern "C" int maus_debug_panel_prof(unsigned long long* out16, int reset) {
    if (out16 && cudaMemcpyFromSymbol(out16, g_panel_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {0}; if (cudaMemcpyToSymbol(g_panel_prof, z, sizeof(z)) != cudaSuccess) return -1; }
    return 0;
}
#endif

cudaError_t lu_build_aug(cplx* W, long long strideW, int n, int batch, const cplx* Acm, const cplx* sigma,
                         const double* psi, const unsigned long long* keys, const cplx* R_cm, const cplx* rhs,
                         long long rhs_stride, cudaStream_t stream, int conj_in) {
    dim3 grid((n + 255) / 256, n + 1, 1);
    static int always_draw = -1;      // MAUS_PHILOX_ALWAYS=1 disables the exact skip (used by the bit-identity test)
    if (always_draw < 0) { const char* e = getenv("MAUS_PHILOX_ALWAYS"); always_draw = (e && atoi(e)) ? 1 : 0; }
    lu_build_aug_kernel<<<grid, 256, 0, stream>>>(W, strideW, n, batch, Acm, sigma, psi, keys, R_cm, rhs, rhs_stride, always_draw, conj_in);
    return cudaGetLastError();
}

template <int R, int IB, int PANEL_NT, int MINB, int UW>
static cudaError_t launch_panel(cplx* W, long long strideW, int n, int k0, int jb, int batch, LuPairs* pairs, int* info,
                                int nc, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(batch * nc, 1, 1);
    cfg.blockDim = dim3(PANEL_NT, 1, 1);
    const size_t dyn = (size_t)R * IB * PANEL_NT * sizeof(cplx);       // lbuf
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(lu_panel_kernel<R, IB, PANEL_NT, MINB, UW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, lu_panel_kernel<R, IB, PANEL_NT, MINB, UW>, W, strideW, n, k0, jb, pairs, info);
}

cudaError_t lu_panel(cplx* W, long long strideW, int n, int k0, int jb, int batch, LuPairs* pairs, int* info,
                     cudaStream_t stream) {
    const int m = n - k0;
    if (m > LU_MAX_N) return cudaErrorInvalidValue;
    // Rows per thread R: fewer, fatter CTAs cost less SM-time per panel (the column loop is latency-bound), more CTAs
    // shorten a single panel.  Batches of >= 12 candidates are SM-time bound -> R = 2 (measured: -23 % panel time at 64
    // candidates); small batches are latency bound -> R = 1.  R = 4 spills and was not faster.
    static int force_r = -1;
    if (force_r < 0) { const char* e = getenv("MAUS_PANEL_R"); force_r = e ? atoi(e) : 0; }
    // configurations (rows per cluster = NC * NT * R, NC <= 8):
    //   A: NT 512, R 1, IB 16 -> <= 4096 rows, lowest latency (small batches)
    //   B: NT 256, R 2, IB 8  -> <= 4096 rows, two CTAs of different clusters share an SM so one cluster's exchange
    //                            latency hides behind the other's arithmetic (large batches are SM-time bound)
    //   C: NT 512, R 2, IB 8  -> <= 8192 rows
    // (round 2 also measured 512 x 1 rows at 64 registers and 256 x 2 at 80 registers / three CTAs per SM: 82.7 / 65.7 ms
    // against 55.9 ms for B at 128 candidates -- spills; removed)
    int mode = (batch >= 12) ? 1 : 0;
    if (force_r == 1) mode = 0; else if (force_r == 2) mode = 2; else if (force_r == 3) mode = 1;
    if (m > PANEL_MAXC * 512) mode = 2;
    const int rows_per_cta = (mode == 2) ? 1024 : 512;
    // cluster size = exactly the CTAs the panel's rows need (1..8, not rounded up to a power of two: a CTA without rows
    // would still occupy its half SM for the whole column loop)
    static int pow2 = -1;
    if (pow2 < 0) { const char* e = getenv("MAUS_PANEL_POW2"); pow2 = (e && atoi(e)) ? 1 : 0; }
    int need = (m + rows_per_cta - 1) / rows_per_cta;
    int nc = need < 1 ? 1 : need;
    if (pow2) { nc = 1; while (nc < need) nc <<= 1; }
    if (mode == 0) return launch_panel<1, 16, 512, 1, 4>(W, strideW, n, k0, jb, batch, pairs, info, nc, stream);
    if (mode == 1) return launch_panel<2, 8, 256, 2, 4>(W, strideW, n, k0, jb, batch, pairs, info, nc, stream);
    return launch_panel<2, 8, 512, 1, 4>(W, strideW, n, k0, jb, batch, pairs, info, nc, stream);
}

cudaError_t lu_permute_rows(cplx* W, long long strideW, int n, int k0, int colstart, int batch, const LuPairs* pairs,
                            cudaStream_t stream) {
    const int ncols = n + 1 - colstart;
    dim3 grid((ncols + PERM_COLS - 1) / PERM_COLS, batch, 1);
    lu_permute_rows_kernel<<<grid, 256, 0, stream>>>(W, strideW, n, k0, colstart, pairs);
    return cudaGetLastError();
}

cudaError_t lu_trtri(const cplx* W, long long strideW, int n, int k0, int jb, int batch, cplx* Linv, cudaStream_t stream) {
    // packed X + the larger of the phase-A row ring and the phase-B T buffer (64 x 65 at jb = 128)
    const size_t smem = ((size_t)(LU_NB * (LU_NB + 1)) / 2 + (size_t)std::max(TRTRI_PF * LU_NB, (LU_NB / 2) * (LU_NB / 2 + 1))) * sizeof(cplx);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(lu_trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    lu_trtri_kernel<<<batch, LU_NB * TRTRI_SPLIT, smem, stream>>>(W, strideW, n, k0, jb, Linv);
    return cudaGetLastError();
}

cudaError_t lu_backsolve(const cplx* W, long long strideW, int n, int batch, const int* info, cplx* X, int* status,
                         cudaStream_t stream) {
    // a cluster of CTAs per candidate: the largest power of two that keeps the grid within two CTAs per SM
    static int use_cluster = -1;            // MAUS_BACKSOLVE_CLUSTER=0: one CTA per candidate always (A/B measurements)
    if (use_cluster < 0) { const char* e = getenv("MAUS_BACKSOLVE_CLUSTER"); use_cluster = e ? (atoi(e) != 0) : 1; }
    int nc = 1;
    while (use_cluster && nc < PANEL_MAXC && batch * nc * 2 <= 2 * MAUS_SM_COUNT_B200) nc *= 2;     // up to two CTAs per SM (measured: 128 candidates x 2: 4.17 -> 2.92 ms, x 4: 3.5)
    static int force_nc = -1;               // MAUS_BACKSOLVE_NC=k: k CTAs per candidate whatever the batch (A/B measurements)
    if (force_nc < 0) { const char* e = getenv("MAUS_BACKSOLVE_NC"); force_nc = e ? atoi(e) : 0; }
    if (force_nc >= 1 && force_nc <= PANEL_MAXC) nc = force_nc;
    if (nc > 1 && n >= 8 * BS_BLK) {
        const int nblk = (n + BS_BLK - 1) / BS_BLK;
        const size_t smem = (size_t)((nblk + nc - 1) / nc) * BS_BLK * sizeof(cplx);
        static size_t attr_smem_c = 0;
        if (smem > attr_smem_c) {
            cudaError_t e = cudaFuncSetAttribute(lu_backsolve_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            attr_smem_c = smem;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(batch * nc, 1, 1);
        cfg.blockDim = dim3(BS_NT, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = nc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, lu_backsolve_cluster_kernel, W, strideW, n, info, X, status);
    }
    const size_t smem = ((size_t)n + BS_BLK * (BS_BLK + 1)) * sizeof(cplx);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(lu_backsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_smem = smem;
    }
    lu_backsolve_kernel<<<batch, BS_NT, smem, stream>>>(W, strideW, n, info, X, status);
    return cudaGetLastError();
}
