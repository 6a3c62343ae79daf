// zgemm.cu -- batched complex128 GEMM for sm_100a.
//
// The only dense contraction of the MAUS hot path: the LU trailing update C -= L21*U12 (AMS:59 -> LAPACK zgetrf)
// and the batched matvecs A*V of the Rayleigh quotient / residual (AMS:268, 297).
//
// Design (B200): tcgen05 has no f64 kind, so the FP64 tensor pipe is reached with warp-level
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; 37.1 TFLOP/s measured register-resident, profiles/fp64_peak_r01.txt).
// Two kernels share the TMA-bulk / mbarrier plumbing:
//
//   zgemm_dmma_kernel    conventional complex product as ONE real product of twice the size: A interleaved (re,im) along K,
//                        B expanded on the fly to [[br, bi], [-bi, br]], accumulator interleaved (re,im).  CTA tile 128 x 64,
//                        8 consumer warps (64 x 16 each) + 1 producer warp.  Batched matvecs A*V, SVD products.
//   zgemm3m_dmma_kernel  (further down) three real products per complex product (3M), 25 % fewer DMMA instructions;
//                        CTA tile 128 x 48, asynchronous TMA epilogue.  All LU trailing updates.
//
// Common: the kernel is persistent (one CTA per SM walks the tile list, m-tile fastest so that concurrently running CTAs
// share the B / U12 columns in L2); a producer warp stages the operands with cp.async.bulk (TMA bulk copies, SASS UBLKCP)
// into shared-memory rings guarded by full/empty mbarriers and runs ahead across tile boundaries (measured: the bulk-copy
// path costs ~35 ns per copy per SM, so copy COUNT, not bytes, bounds the feed -> few, large slabs); column strides in
// shared memory are padded so that the fragment loads are bank-conflict free; numpy's complex128 layout end to end.
#include <cstdlib>
#include "zgemm.cuh"

namespace {

constexpr int TN = 64;
constexpr int KC = 16, STAGES = 3;           // A slabs: 16 complex k per stage (16 bulk copies of TM*16 B)
constexpr int KCB = 32;                      // B slabs: 32 complex k per stage (64 bulk copies of 512 B) -- fewer, larger copies
constexpr int LDSB = KCB + 2;                // complex elements between consecutive n-columns of the B slab
constexpr int B_STAGE = TN * LDSB;
// Tile configurations.  AB = m8 row-blocks per warp (warp tile 8*AB x 16 complex); NWM x 4 consumer warps cover TM x 64.
//   <128, 8>: 8 consumer warps, 1 CTA / SM, B double-buffered        <128, 4>: 16 consumer warps
//   < 64, 8>: 4 consumer warps, 2 CTAs / SM (one CTA's C-tile epilogue and pipeline bubbles hide behind the other's MMAs)
template <int TM_, int AB> struct Cfg {
    static constexpr int TM = TM_;
    static constexpr int NWM = TM / (8 * AB);
    static constexpr int NCONS = NWM * 4;
    static constexpr int NTHREADS = (NCONS + 1) * 32;
    static constexpr int CTAS_PER_SM = (TM == 64) ? 2 : 1;
    static constexpr int BSTAGES = (TM == 64) ? 1 : 2;      // 2 CTAs / SM must stay under 113 KB each
    static constexpr int LDSA = TM + 4;      // complex elements between consecutive k-columns of the A slab
    static constexpr int A_STAGE = KC * LDSA;
    static constexpr int NBAR = 2 * STAGES + 2 * BSTAGES;
    static constexpr size_t SMEM_BYTES = (size_t)(STAGES * A_STAGE + BSTAGES * B_STAGE) * sizeof(cplx) + NBAR * sizeof(uint64_t);
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// L2 policies: the A / B panels are re-read by many tiles (keep), the C tiles stream through once (do not keep)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ cplx ld_stream(const cplx* p, uint64_t policy) {
    cplx v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(policy) : "memory");
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int TM, int AB>
__global__ void __launch_bounds__(Cfg<TM, AB>::NTHREADS, Cfg<TM, AB>::CTAS_PER_SM) zgemm_dmma_kernel(ZgemmParams p) {
    using CF = Cfg<TM, AB>;
    constexpr int NCONS = CF::NCONS, BSTAGES = CF::BSTAGES, LDSA = CF::LDSA, A_STAGE = CF::A_STAGE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* sA = reinterpret_cast<cplx*>(smem_raw);
    cplx* sB = sA + STAGES * A_STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + BSTAGES * B_STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* bfull = empty + STAGES;
    uint64_t* bempty = bfull + BSTAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KT = (p.K + KC - 1) / KC;                 // A slabs per tile
    constexpr int APB = KCB / KC;                       // A slabs per B slab
    const int tiles_m = (p.M + TM - 1) / TM, tiles_n = (p.N + TN - 1) / TN;
    const long long ntiles = (long long)tiles_m * tiles_n * p.batch;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS); }
        for (int s = 0; s < BSTAGES; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], NCONS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // PERSISTENT: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (m-tile fastest so that the CTAs running
    // at the same time share the B / U12 columns in L2).  The producer runs ahead across tile boundaries, so the
    // pipeline fill of the next tile is not exposed; the epilogue stores drain behind the next K loop.
    if (warp == NCONS) {
        // ---------------- producer warp: TMA bulk copies, one per contiguous column segment ----------------
        long long it = 0, ib = 0;
        const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int mt = (int)(tile % tiles_m), nt = (int)((tile / tiles_m) % tiles_n), bz = (int)(tile / ((long long)tiles_m * tiles_n));
            const int m0 = mt * TM, n0 = nt * TN;
            const cplx* A = p.A + (long long)bz * p.strideA;
            const cplx* B = p.B + (long long)bz * p.strideB;
            const int rows_valid = min(TM, p.M - m0), cols_valid = min(TN, p.N - n0);
            if (p.beta) {
                // warm L2 with the C tile of the NEXT tile of this CTA (and of the first one): the consumers initialise
                // their accumulators from it at tile start, so the load must be an L2 hit by then
                for (long long tn = (tile == blockIdx.x ? tile : tile + gridDim.x); tn < ntiles && tn <= tile + gridDim.x; tn += gridDim.x) {
                    const int mt2 = (int)(tn % tiles_m), nt2 = (int)((tn / tiles_m) % tiles_n), bz2 = (int)(tn / ((long long)tiles_m * tiles_n));
                    const cplx* C2 = p.C + (long long)bz2 * p.strideC + mt2 * TM + (long long)(nt2 * TN) * p.ldc;
                    const int rv = min(TM, p.M - mt2 * TM), cvn = min(TN, p.N - nt2 * TN);
                    for (int j = lane; j < cvn; j += 32)
                        asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(C2 + (long long)j * p.ldc),
                                     "r"((uint32_t)(rv * sizeof(cplx))), "l"(stream) : "memory");
                }
            }
            for (int kt = 0; kt < KT; ++kt, ++it) {
                if (kt % APB == 0) {
                    const int sb = (int)(ib % BSTAGES), kb0 = kt * KC;
                    const int kvb = min(KCB, p.K - kb0);
                    if (ib >= BSTAGES) mbar_wait(&bempty[sb], (uint32_t)(((ib / BSTAGES) - 1) & 1));
                    if (lane == 0) mbar_expect_tx(&bfull[sb], (uint32_t)(cols_valid * kvb * sizeof(cplx)));
                    __syncwarp();
                    cplx* b_s = sB + sb * B_STAGE;
                    for (int j = lane; j < cols_valid; j += 32)
                        bulk_g2s(b_s + j * LDSB, B + kb0 + (long long)(n0 + j) * p.ldb, (uint32_t)(kvb * sizeof(cplx)), &bfull[sb], keep);
                    ++ib;
                }
                const int s = (int)(it % STAGES), k0 = kt * KC;
                const int kv = min(KC, p.K - k0);
                if (it >= STAGES) mbar_wait(&empty[s], (uint32_t)(((it / STAGES) - 1) & 1));
                if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)(kv * rows_valid * sizeof(cplx)));
                __syncwarp();
                if (lane < kv)
                    bulk_g2s(sA + s * A_STAGE + lane * LDSA, A + m0 + (long long)(k0 + lane) * p.lda,
                             (uint32_t)(rows_valid * sizeof(cplx)), &full[s], keep);
            }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int wm = warp >> 2, wn = warp & 3;       // NWM x 4 warps
    const int g = lane >> 2, t = lane & 3;         // fragment coordinates
    // B~ = [[br, bi], [-bi, br]]: this lane holds B~[kk = t][nn = g] -> component (t ^ g) & 1, negative iff (t odd, g even)
    const int comp = (t ^ g) & 1;
    const long long sflip = (((((t & 1) && !(g & 1)) ? 1 : 0) ^ (p.negate ? 1 : 0)) ? 1LL : 0LL) << 63;   // sign-bit flip
    const int kh = t >> 1;                          // which of the two complex k of a k4 step
    const int arow = wm * (8 * AB) + g;
    const int bcol = wn * 16 + (g >> 1);
    long long it = 0, ib = 0;
    const uint64_t stream = l2_policy_evict_first();
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int mt = (int)(tile % tiles_m), nt = (int)((tile / tiles_m) % tiles_n), bz = (int)(tile / ((long long)tiles_m * tiles_n));
        const int m0 = mt * TM, n0 = nt * TN;
        cplx* C = p.C + (long long)bz * p.strideC;
        double acc[AB][4][2];
        const int crow = m0 + wm * (8 * AB) + g;       // + 8*qa
        const int ccol = n0 + wn * 16 + t;             // + 4*qb
        if (p.beta) {
            // accumulators start from C (L2 hits: the producer prefetched this tile while the previous one was computed)
#pragma unroll
            for (int qa = 0; qa < AB; ++qa)
#pragma unroll
                for (int qb = 0; qb < 4; ++qb) {
                    int r = crow + 8 * qa, c = ccol + 4 * qb;
                    cplx v = (r < p.M && c < p.N) ? ld_stream(&C[r + (long long)c * p.ldc], stream) : cmake(0.0, 0.0);
                    acc[qa][qb][0] = v.x; acc[qa][qb][1] = v.y;
                }
        } else {
#pragma unroll
            for (int qa = 0; qa < AB; ++qa)
#pragma unroll
                for (int qb = 0; qb < 4; ++qb) { acc[qa][qb][0] = 0.0; acc[qa][qb][1] = 0.0; }
        }
        const double* b_s = nullptr;
        int sb = 0;
        for (int kt = 0; kt < KT; ++kt, ++it) {
            if (kt % APB == 0) {
                sb = (int)(ib % BSTAGES);
                mbar_wait(&bfull[sb], (uint32_t)((ib / BSTAGES) & 1));
                b_s = reinterpret_cast<const double*>(sB + sb * B_STAGE);
                ++ib;
            }
            const int s = (int)(it % STAGES);
            const int kv = min(KC, p.K - kt * KC);
            const int kboff = (kt % APB) * KC;       // offset of this A slab inside the B slab
            mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
            const double* a_s = reinterpret_cast<const double*>(sA + s * A_STAGE);
            if (kv == KC) {
#pragma unroll
                for (int ks = 0; ks < KC / 2; ++ks) {
                    const int kc = 2 * ks + kh;
                    double af[AB], bf[4];
#pragma unroll
                    for (int qa = 0; qa < AB; ++qa) af[qa] = a_s[2 * (kc * LDSA + arow + 8 * qa) + (t & 1)];
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb)
                        bf[qb] = __longlong_as_double(__double_as_longlong(b_s[2 * ((bcol + 4 * qb) * LDSB + kboff + kc) + comp]) ^ sflip);
#pragma unroll
                    for (int qa = 0; qa < AB; ++qa)
#pragma unroll
                        for (int qb = 0; qb < 4; ++qb) dmma884(acc[qa][qb][0], acc[qa][qb][1], af[qa], bf[qb]);
                }
            } else {
                // K tail: complex k >= kv contribute exact zeros (both fragments are cleared in registers)
                for (int ks = 0; 2 * ks < kv; ++ks) {
                    const int kc = 2 * ks + kh;
                    const bool ok = kc < kv;
                    double af[AB], bf[4];
#pragma unroll
                    for (int qa = 0; qa < AB; ++qa) af[qa] = ok ? a_s[2 * (kc * LDSA + arow + 8 * qa) + (t & 1)] : 0.0;
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb)
                        bf[qb] = ok ? __longlong_as_double(__double_as_longlong(b_s[2 * ((bcol + 4 * qb) * LDSB + kboff + kc) + comp]) ^ sflip) : 0.0;
#pragma unroll
                    for (int qa = 0; qa < AB; ++qa)
#pragma unroll
                        for (int qb = 0; qb < 4; ++qb) dmma884(acc[qa][qb][0], acc[qa][qb][1], af[qa], bf[qb]);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty[s]);
                if (kt % APB == APB - 1 || kt == KT - 1) mbar_arrive(&bempty[sb]);
            }
        }
#pragma unroll
        for (int qa = 0; qa < AB; ++qa)
#pragma unroll
            for (int qb = 0; qb < 4; ++qb) {
                int r = crow + 8 * qa, c = ccol + 4 * qb;
                if (r < p.M && c < p.N) st_stream(&C[r + (long long)c * p.ldc], cmake(acc[qa][qb][0], acc[qa][qb][1]), stream);
            }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 3M variant (LU trailing updates).  (Ar + iAi)(Br + iBi) with THREE real products instead of four:
//   P1 = Ar Br,  P2 = Ai Bi,  P3 = (Ar + Ai)(Br + Bi)      Re = P1 - P2,  Im = P3 - P1 - P2
// (the "3M" / Karatsuba complex product of ZGEMM3M; normwise as stable as the conventional product, Higham 1992).
// It executes 6 instead of 8 real flops per complex multiply-add on the same DMMA pipe, i.e. 25 % fewer tensor
// instructions for the same algorithmic work.  Operands stay interleaved complex in shared memory: one LDS.128 fetches
// (re, im) of a fragment element, the sums are formed in registers (2 DADD per 18 DMMA).
//
//   CTA tile 128 x 48 complex, 8 consumer warps (4 x 2), each 32 x 24 complex = 4 x 3 DMMA tiles x 3 products
//   (72 f64 accumulators) + 1 producer warp (same TMA bulk-copy rings as above).  Column strides 130 / 36 complex make
//   the 128-bit fragment loads conflict free per quarter warp.  C is added in the epilogue (accumulators start at 0).
namespace m3 {
constexpr int TM = 128;
constexpr int QA = 4, NWM = 4, NWN = 2;          // QB (m8n8 tiles per warp along n) is a Pipe parameter: CTA tile 128 x (16 QB)
// Register budget: the register file is split per SM sub-partition (16 K registers each), so a 9th warp would cap every
// thread at 168 registers.  The CTA is launched as three warpgroups (2 consumer + 1 producer warpgroup of which one warp
// works) at 168 registers, then the consumers grow to 232 and the producer warpgroup shrinks to 40 (setmaxnreg):
// per sub-partition 2 x 32 x 232 + 32 x 40 = 16 128 <= 16 384.
constexpr int NCONS = NWM * NWN, NTHREADS = (NCONS + 4) * 32;
constexpr int REG_CONSUMER = 232, REG_PRODUCER = 40;
static_assert(NCONS == 8, "two consumer warpgroups");
static_assert(NWM * QA * 8 == TM, "warp grid must cover the CTA tile");
// pipeline shape: A slabs of KC complex k (KC bulk copies of TM*16 B), B slabs of KCB complex k (TN copies of KCB*16 B);
// QB = 3: 128 x 48 tiles (LU trailing updates), QB = 2: 128 x 32 tiles (skinny batched A*V: N = 64 / 128 candidates are whole
// multiples of 32 but leave a 16-wide remainder tile at 48)
template <int KC_, int STAGES_, int KCB_, int BSTAGES_, int QB_ = 3> struct Pipe {
    static constexpr int KC = KC_, STAGES = STAGES_, KCB = KCB_, BSTAGES = BSTAGES_, QB = QB_, TN = NWN * QB_ * 8;
    static constexpr int LDSA = TM + 2, LDSB = KCB + 4;        // strides = 2 / 4 (mod 8) complex: conflict-free LDS.128
    static constexpr int A_STAGE = KC * LDSA, B_STAGE = TN * LDSB;
    // result staging tile [TN][LDSC]: column stride = 1 (mod 4) complex makes the fragment-order STS.128 conflict free
    static constexpr int LDSC = TM + 1, C_STAGE = TN * LDSC;
    static constexpr int NBAR = 2 * STAGES + 2 * BSTAGES + 2;
    static constexpr size_t SMEM_BYTES = (size_t)(STAGES * A_STAGE + BSTAGES * B_STAGE + C_STAGE) * sizeof(cplx) + NBAR * sizeof(uint64_t);
    static_assert(KC <= 32 && KCB % KC == 0 && KC % 4 == 0, "slab shapes");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};
}  // namespace m3

template <class PP>
__global__ void __launch_bounds__(m3::NTHREADS, 1) zgemm3m_dmma_kernel(ZgemmParams p) {
    constexpr int TM = m3::TM, TN = PP::TN, KC = PP::KC, STAGES = PP::STAGES, KCB = PP::KCB, BSTAGES = PP::BSTAGES, LDSA = PP::LDSA,
                  LDSB = PP::LDSB, A_STAGE = PP::A_STAGE, B_STAGE = PP::B_STAGE, QA = m3::QA, QB = PP::QB, NWN = m3::NWN,
                  NCONS = m3::NCONS, LDSC = PP::LDSC, C_STAGE = PP::C_STAGE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* sA = reinterpret_cast<cplx*>(smem_raw);
    cplx* sB = sA + STAGES * A_STAGE;
    cplx* sC = sB + BSTAGES * B_STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(sC + C_STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* bfull = empty + STAGES;
    uint64_t* bempty = bfull + BSTAGES;
    uint64_t* cfull = bempty + BSTAGES;          // consumers -> storer: the staged result tile is complete
    uint64_t* cempty = cfull + 1;                // storer -> consumers: the staging tile may be overwritten

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KT = (p.K + KC - 1) / KC;
    constexpr int APB = KCB / KC;
    const int tiles_m = (p.M + TM - 1) / TM, tiles_n = (p.N + TN - 1) / TN;
    const long long ntiles = (long long)tiles_m * tiles_n * p.batch;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS); }
        for (int s = 0; s < BSTAGES; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], NCONS); }
        mbar_init(cfull, NCONS); mbar_init(cempty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= NCONS) {
        // ---------------- producer warpgroup: a loader warp and a storer warp ----------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(m3::REG_PRODUCER));
        if (warp == NCONS + 1) {
            // storer: the consumers park the finished tile in sC and go on with the next tile; this warp writes it out with
            // TMA bulk operations, one per column -- a plain store (beta = 0) or an f64 add performed by the L2 (C += tile),
            // so C never travels to the SM and no thread waits for a global round trip
            long long tl = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
                const int mt = (int)(tile % tiles_m), nt = (int)((tile / tiles_m) % tiles_n), bz = (int)(tile / ((long long)tiles_m * tiles_n));
                const int m0 = mt * TM, n0 = nt * TN;
                const int rows_valid = min(TM, p.M - m0), cols_valid = min(TN, p.N - n0);
                cplx* Cg = p.C + (long long)bz * p.strideC + m0 + (long long)n0 * p.ldc;
                mbar_wait(cfull, (uint32_t)(tl & 1));
                for (int j = lane; j < cols_valid; j += 32) {
                    if (p.beta)
                        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                                     ::"l"(Cg + (long long)j * p.ldc), "r"(smem_u32(sC + j * LDSC)), "r"((uint32_t)(rows_valid * sizeof(cplx))) : "memory");
                    else
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                     ::"l"(Cg + (long long)j * p.ldc), "r"(smem_u32(sC + j * LDSC)), "r"((uint32_t)(rows_valid * sizeof(cplx))) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // sC has been read
                __syncwarp();
                if (lane == 0) mbar_arrive(cempty);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");               // all writes performed
            return;
        }
        if (warp != NCONS) return;
        long long it = 0, ib = 0;
        const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int mt = (int)(tile % tiles_m), nt = (int)((tile / tiles_m) % tiles_n), bz = (int)(tile / ((long long)tiles_m * tiles_n));
            const int m0 = mt * TM, n0 = nt * TN;
            const cplx* A = p.A + (long long)bz * p.strideA;
            const cplx* B = p.B + (long long)bz * p.strideB;
            const int rows_valid = min(TM, p.M - m0), cols_valid = min(TN, p.N - n0);
            for (int kt = 0; kt < KT; ++kt, ++it) {
                if (kt % APB == 0) {
                    const int sb = (int)(ib % BSTAGES), kb0 = kt * KC;
                    const int kvb = min(KCB, p.K - kb0);
                    if (ib >= BSTAGES) mbar_wait(&bempty[sb], (uint32_t)(((ib / BSTAGES) - 1) & 1));
                    if (lane == 0) mbar_expect_tx(&bfull[sb], (uint32_t)(cols_valid * kvb * sizeof(cplx)));
                    __syncwarp();
                    cplx* b_s = sB + sb * B_STAGE;
                    for (int j = lane; j < cols_valid; j += 32)
                        bulk_g2s(b_s + j * LDSB, B + kb0 + (long long)(n0 + j) * p.ldb, (uint32_t)(kvb * sizeof(cplx)), &bfull[sb], keep);
                    ++ib;
                }
                const int s = (int)(it % STAGES), k0 = kt * KC;
                const int kv = min(KC, p.K - k0);
                if (it >= STAGES) mbar_wait(&empty[s], (uint32_t)(((it / STAGES) - 1) & 1));
                if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)(kv * rows_valid * sizeof(cplx)));
                __syncwarp();
                if (lane < kv)
                    bulk_g2s(sA + s * A_STAGE + lane * LDSA, A + m0 + (long long)(k0 + lane) * p.lda,
                             (uint32_t)(rows_valid * sizeof(cplx)), &full[s], keep);
            }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(m3::REG_CONSUMER));
    const int wm = warp / NWN, wn = warp % NWN;
    const int g = lane >> 2, t = lane & 3;         // DMMA fragment coordinates: A[g][t], B[t][g], C[g][2t, 2t+1]
    const long long sflip = (p.negate ? 1LL : 0LL) << 63;
    const int arow = wm * (8 * QA) + g;
    const int bcol = wn * (8 * QB) + g;
    long long it = 0, ib = 0, tl = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        double p1[QA][QB][2], p2[QA][QB][2], p3[QA][QB][2];
#pragma unroll
        for (int qa = 0; qa < QA; ++qa)
#pragma unroll
            for (int qb = 0; qb < QB; ++qb)
#pragma unroll
                for (int j = 0; j < 2; ++j) { p1[qa][qb][j] = 0.0; p2[qa][qb][j] = 0.0; p3[qa][qb][j] = 0.0; }
        // K loop.  The fragments of a k4 step (4 + 3 LDS.128) are loaded one step ahead of the 36 DMMAs that use them, also
        // ACROSS slab boundaries: the last step of a slab first waits for the next slab's barrier and loads its first
        // fragments, then issues its own DMMAs and only then releases its slab -- no LDS / barrier latency is exposed
        // between slabs (needs >= 2 stages of A and of B).
        struct Frag { double2 a[QA]; double2 b[QB]; };
        const double2 *a_s = nullptr, *b_sk = nullptr;      // current A slab; current B slab advanced to this A slab's k offset
        int s = 0, sb = 0;
        auto enter_slab = [&](int kt) {
            if (kt % APB == 0) {
                sb = (int)(ib % BSTAGES);
                mbar_wait(&bfull[sb], (uint32_t)((ib / BSTAGES) & 1));
                ++ib;
            }
            s = (int)(it % STAGES);
            mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
            ++it;
            a_s = reinterpret_cast<const double2*>(sA + s * A_STAGE);
            b_sk = reinterpret_cast<const double2*>(sB + sb * B_STAGE) + (kt % APB) * KC;
        };
        auto load_frag = [&](Frag& f, int ks) {
            const int kc = 4 * ks + t;                       // this lane's complex k of the k4 step
#pragma unroll
            for (int qb = 0; qb < QB; ++qb) f.b[qb] = b_sk[(bcol + 8 * qb) * LDSB + kc];
#pragma unroll
            for (int qa = 0; qa < QA; ++qa) f.a[qa] = a_s[kc * LDSA + arow + 8 * qa];
        };
        auto mma_frag = [&](const Frag& f) {
            double br[QB], bi[QB], bs[QB];
#pragma unroll
            for (int qb = 0; qb < QB; ++qb) {
                br[qb] = __longlong_as_double(__double_as_longlong(f.b[qb].x) ^ sflip);
                bi[qb] = __longlong_as_double(__double_as_longlong(f.b[qb].y) ^ sflip);
                bs[qb] = br[qb] + bi[qb];
            }
#pragma unroll
            for (int qa = 0; qa < QA; ++qa) {
                const double as = f.a[qa].x + f.a[qa].y;
#pragma unroll
                for (int qb = 0; qb < QB; ++qb) {
                    dmma884(p1[qa][qb][0], p1[qa][qb][1], f.a[qa].x, br[qb]);
                    dmma884(p2[qa][qb][0], p2[qa][qb][1], f.a[qa].y, bi[qb]);
                    dmma884(p3[qa][qb][0], p3[qa][qb][1], as, bs[qb]);
                }
            }
        };
        auto release_slab = [&](int s_rel, int sb_rel, bool rel_b) {
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty[s_rel]);
                if (rel_b) mbar_arrive(&bempty[sb_rel]);
            }
        };
        constexpr int NS = KC / 4;                           // k4 steps per full slab (even)
        static_assert(NS % 2 == 0, "fragment ping-pong needs an even number of steps per slab");
        Frag f0, f1;
        if (KT > 0) {
            enter_slab(0);
            if (p.K >= KC) load_frag(f0, 0);
        }
        for (int kt = 0; kt < KT; ++kt) {
            const int kv = min(KC, p.K - kt * KC);
            const bool rel_b = (kt % APB == APB - 1 || kt == KT - 1);
            if (kv == KC) {
                // full slab: f0 holds step 0 (loaded by the previous slab's last step or before the loop)
#pragma unroll
                for (int ks = 0; ks < NS; ks += 2) {
                    load_frag(f1, ks + 1);
                    mma_frag(f0);
                    if (ks + 2 < NS) {
                        load_frag(f0, ks + 2);
                        mma_frag(f1);
                    } else {
                        const int s_rel = s, sb_rel = sb;
                        if (kt + 1 < KT) {
                            enter_slab(kt + 1);
                            if (p.K - (kt + 1) * KC >= KC) load_frag(f0, 0);
                        }
                        mma_frag(f1);
                        release_slab(s_rel, sb_rel, rel_b);
                    }
                }
            } else {
                // K tail (always the last slab of the tile): complex k >= kv contribute exact zeros
                for (int ks = 0; 4 * ks < kv; ++ks) {
                    const bool ok = 4 * ks + t < kv;
                    Frag f;
                    if (ok) load_frag(f, ks);
                    else {
#pragma unroll
                        for (int qb = 0; qb < QB; ++qb) f.b[qb] = make_double2(0.0, 0.0);
#pragma unroll
                        for (int qa = 0; qa < QA; ++qa) f.a[qa] = make_double2(0.0, 0.0);
                    }
                    mma_frag(f);
                }
                release_slab(s, sb, rel_b);
            }
        }
        // epilogue: Re = P1 - P2, Im = (P3 - P1) - P2 go to the staging tile (fragment order, conflict-free STS.128); the
        // storer warp adds them to / stores them over C asynchronously while this warp starts the next tile
        if (tl >= 1) mbar_wait(cempty, (uint32_t)((tl - 1) & 1));
        {
            cplx* dst = sC + (wn * (8 * QB) + 2 * t) * LDSC + wm * (8 * QA) + g;
#pragma unroll
            for (int qa = 0; qa < QA; ++qa)
#pragma unroll
                for (int qb = 0; qb < QB; ++qb)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        dst[(8 * qb + j) * LDSC + 8 * qa] = cmake(p1[qa][qb][j] - p2[qa][qb][j], (p3[qa][qb][j] - p1[qa][qb][j]) - p2[qa][qb][j]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the bulk (async proxy) reads
        __syncwarp();
        if (lane == 0) mbar_arrive(cfull);
        ++tl;
    }
}

// Independent FP64-FMA implementation (16 x 16 tiles) -- test cross-check only.
__global__ void __launch_bounds__(256) zgemm_simple_kernel(ZgemmParams p) {
    __shared__ cplx sa[16][17], sb[16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int r = blockIdx.x * 16 + tx, c = blockIdx.y * 16 + ty, bz = blockIdx.z;
    const cplx* A = p.A + (long long)bz * p.strideA;
    const cplx* B = p.B + (long long)bz * p.strideB;
    cplx* C = p.C + (long long)bz * p.strideC;
    cplx acc = cmake(0.0, 0.0);
    for (int k0 = 0; k0 < p.K; k0 += 16) {
        int ka = k0 + ty, kb = k0 + tx;
        sa[ty][tx] = (r < p.M && ka < p.K) ? A[r + (long long)ka * p.lda] : cmake(0.0, 0.0);
        sb[ty][tx] = (c < p.N && kb < p.K) ? B[kb + (long long)c * p.ldb] : cmake(0.0, 0.0);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) cfma(acc, sa[k][tx], sb[ty][k]);
        __syncthreads();
    }
    if (r < p.M && c < p.N) {
        cplx* dst = &C[r + (long long)c * p.ldc];
        if (p.negate) acc = cmake(-acc.x, -acc.y);
        if (p.beta) acc = cadd(acc, *dst);
        *dst = acc;
    }
}

}  // namespace

template <int TM, int AB>
static cudaError_t launch_cfg(const ZgemmParams& p, cudaStream_t stream) {
    using CF = Cfg<TM, AB>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(zgemm_dmma_kernel<TM, AB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const long long ntiles = (long long)((p.M + TM - 1) / TM) * ((p.N + TN - 1) / TN) * p.batch;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = MAUS_SM_COUNT_B200; }
    const long long slots = (long long)sms * CF::CTAS_PER_SM;
    const unsigned grid = (unsigned)(ntiles < slots ? ntiles : slots);     // persistent: CTAS_PER_SM CTAs per SM
    zgemm_dmma_kernel<TM, AB><<<grid, CF::NTHREADS, CF::SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

template <class PP>
static cudaError_t launch_3m_cfg(const ZgemmParams& p, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(zgemm3m_dmma_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PP::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const long long ntiles = (long long)((p.M + m3::TM - 1) / m3::TM) * ((p.N + PP::TN - 1) / PP::TN) * p.batch;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = MAUS_SM_COUNT_B200; }
    const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);          // persistent: one CTA per SM
    zgemm3m_dmma_kernel<PP><<<grid, m3::NTHREADS, PP::SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

static cudaError_t launch_3m(const ZgemmParams& p, cudaStream_t stream) {
    if (p.tile_n == 32) return launch_3m_cfg<m3::Pipe<16, 2, 32, 2, 2>>(p, stream);
    static int cfg = -1;
    if (cfg < 0) { const char* e = getenv("MAUS_3M_CFG"); cfg = e ? atoi(e) : 0; }
    switch (cfg) {
        case 1: return launch_3m_cfg<m3::Pipe<8, 4, 32, 2>>(p, stream);
        case 2: return launch_3m_cfg<m3::Pipe<8, 4, 16, 4>>(p, stream);
        default: return launch_3m_cfg<m3::Pipe<16, 2, 32, 2>>(p, stream);
    }
}

static int g_zgemm_cfg = -1;
void zgemm_set_config(int cfg) { g_zgemm_cfg = cfg; }

cudaError_t zgemm_dmma_launch(const ZgemmParams& p, cudaStream_t stream) {
    if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return cudaSuccess;
    if (p.K <= 0 && p.beta) return cudaSuccess;          // nothing to accumulate: C = C
    if (g_zgemm_cfg < 0) {
        const char* e = getenv("MAUS_GEMM_CFG");
        g_zgemm_cfg = e ? atoi(e) : 0;
    }
    if (p.algo3m) return launch_3m(p, stream);              // 128-row tiles: also valid for the in-place U12 solve
    if ((const void*)p.C == (const void*)p.B) return launch_cfg<128, 8>(p, stream);   // in-place (U12 = L11^-1 A12): one row tile must own all rows
    if (g_zgemm_cfg == 1) return launch_cfg<128, 4>(p, stream);
    if (g_zgemm_cfg == 3) return launch_cfg<64, 8>(p, stream);   // 64 x 64 tiles, 2 CTAs / SM
    return launch_cfg<128, 8>(p, stream);                    // default: 128 x 64 tiles, 8 consumer warps, 1 CTA / SM
}

cudaError_t zgemm_simple_launch(const ZgemmParams& p, cudaStream_t stream) {
    if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return cudaSuccess;
    dim3 grid((p.M + 15) / 16, (p.N + 15) / 16, p.batch);
    zgemm_simple_kernel<<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
}
