// zgemm.cu -- batched complex128 GEMM for sm_100a.
//
// The only dense contraction of the MAUS hot path: the LU trailing update C -= L21*U12 (AMS:59 -> LAPACK zgetrf)
// and the batched matvecs A*V of the Rayleigh quotient / residual (AMS:268, 297).
//
// Design (B200): tcgen05 has no f64 kind, so the FP64 tensor pipe is reached with warp-level
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; 37.1 TFLOP/s measured register-resident, profiles/fp64_peak_r01.txt).
// A complex product is computed as ONE real product of twice the size: with A interleaved (re,im) along K and
// B expanded on the fly to [[br, bi], [-bi, br]], the accumulator comes out interleaved (re,im) as well, so
// global memory keeps numpy's complex128 layout end to end.
//
//   CTA tile 128 x 64 complex, 8 consumer warps (2 x 4), each 64 x 16 complex = 8 x 4 DMMA tiles (64 f64 accum),
//   + 1 producer warp that stages K-slabs of 16 complex with cp.async.bulk (TMA bulk copies, SASS UBLKCP) into a
//   4-deep shared-memory ring guarded by full/empty mbarriers.  Column strides in shared memory are padded
//   (132 / 18 complex) so that both fragment loads are bank-conflict free.
#include "zgemm.cuh"

namespace {

constexpr int TM = 128, TN = 64, KC = 16, STAGES = 4;
constexpr int LDSA = TM + 4;     // complex elements between consecutive k-columns of the A slab
constexpr int LDSB = KC + 2;     // complex elements between consecutive n-columns of the B slab
constexpr int A_STAGE = KC * LDSA;           // complex elements
constexpr int B_STAGE = TN * LDSB;
constexpr int NCONS = 8;                     // consumer warps
constexpr int NTHREADS = (NCONS + 1) * 32;
constexpr size_t SMEM_BYTES = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(cplx) + 2 * STAGES * sizeof(uint64_t);

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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(NTHREADS, 1) zgemm_dmma_kernel(ZgemmParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* sA = reinterpret_cast<cplx*>(smem_raw);
    cplx* sB = sA + STAGES * A_STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);
    uint64_t* empty = full + STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN, bz = blockIdx.z;
    const cplx* A = p.A + (long long)bz * p.strideA;
    const cplx* B = p.B + (long long)bz * p.strideB;
    cplx* C = p.C + (long long)bz * p.strideC;
    const int KT = (p.K + KC - 1) / KC;
    const int rows_valid = min(TM, p.M - m0), cols_valid = min(TN, p.N - n0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCONS) {
        // ---------------- producer warp: TMA bulk copies, one per contiguous column segment ----------------
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES, k0 = kt * KC;
            const int kv = min(KC, p.K - k0);
            if (kt >= STAGES) mbar_wait(&empty[s], ((kt / STAGES) - 1) & 1);
            if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)((kv * rows_valid + cols_valid * kv) * sizeof(cplx)));
            __syncwarp();
            cplx* a_s = sA + s * A_STAGE;
            cplx* b_s = sB + s * B_STAGE;
            if (lane < kv)
                bulk_g2s(a_s + lane * LDSA, A + m0 + (long long)(k0 + lane) * p.lda,
                         (uint32_t)(rows_valid * sizeof(cplx)), &full[s]);
            for (int j = lane; j < cols_valid; j += 32)
                bulk_g2s(b_s + j * LDSB, B + k0 + (long long)(n0 + j) * p.ldb, (uint32_t)(kv * sizeof(cplx)), &full[s]);
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int wm = warp >> 2, wn = warp & 3;       // 2 x 4 warps
    const int g = lane >> 2, t = lane & 3;         // fragment coordinates
    double acc[8][4][2];
    const int crow = m0 + wm * 64 + g;             // + 8*qa
    const int ccol = n0 + wn * 16 + t;             // + 4*qb
    if (p.beta) {
#pragma unroll
        for (int qa = 0; qa < 8; ++qa)
#pragma unroll
            for (int qb = 0; qb < 4; ++qb) {
                int r = crow + 8 * qa, c = ccol + 4 * qb;
                cplx v = cmake(0.0, 0.0);
                if (r < p.M && c < p.N) v = C[r + (long long)c * p.ldc];
                acc[qa][qb][0] = v.x; acc[qa][qb][1] = v.y;
            }
    } else {
#pragma unroll
        for (int qa = 0; qa < 8; ++qa)
#pragma unroll
            for (int qb = 0; qb < 4; ++qb) { acc[qa][qb][0] = 0.0; acc[qa][qb][1] = 0.0; }
    }
    // B~ = [[br, bi], [-bi, br]]: this lane holds B~[kk = t][nn = g] -> component (t ^ g) & 1, negative iff (t odd, g even)
    const int comp = (t ^ g) & 1;
    const double sgn = ((((t & 1) && !(g & 1)) ? 1 : 0) ^ (p.negate ? 1 : 0)) ? -1.0 : 1.0;
    const int kh = t >> 1;                          // which of the two complex k of a k4 step
    const int arow = wm * 64 + g;
    const int bcol = wn * 16 + (g >> 1);

    for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % STAGES;
        const int kv = min(KC, p.K - kt * KC);
        mbar_wait(&full[s], (kt / STAGES) & 1);
        const double* a_s = reinterpret_cast<const double*>(sA + s * A_STAGE);
        const double* b_s = reinterpret_cast<const double*>(sB + s * B_STAGE);
        if (kv == KC) {
#pragma unroll
            for (int ks = 0; ks < KC / 2; ++ks) {
                const int kc = 2 * ks + kh;
                double af[8], bf[4];
#pragma unroll
                for (int qa = 0; qa < 8; ++qa) af[qa] = a_s[2 * (kc * LDSA + arow + 8 * qa) + (t & 1)];
#pragma unroll
                for (int qb = 0; qb < 4; ++qb) bf[qb] = b_s[2 * ((bcol + 4 * qb) * LDSB + kc) + comp] * sgn;
#pragma unroll
                for (int qa = 0; qa < 8; ++qa)
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb) dmma884(acc[qa][qb][0], acc[qa][qb][1], af[qa], bf[qb]);
            }
        } else {
            // K tail: complex k >= kv contribute exact zeros (both fragments are cleared in registers)
            for (int ks = 0; 2 * ks < kv; ++ks) {
                const int kc = 2 * ks + kh;
                const bool ok = kc < kv;
                double af[8], bf[4];
#pragma unroll
                for (int qa = 0; qa < 8; ++qa) af[qa] = ok ? a_s[2 * (kc * LDSA + arow + 8 * qa) + (t & 1)] : 0.0;
#pragma unroll
                for (int qb = 0; qb < 4; ++qb) bf[qb] = ok ? b_s[2 * ((bcol + 4 * qb) * LDSB + kc) + comp] * sgn : 0.0;
#pragma unroll
                for (int qa = 0; qa < 8; ++qa)
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb) dmma884(acc[qa][qb][0], acc[qa][qb][1], af[qa], bf[qb]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
#pragma unroll
    for (int qa = 0; qa < 8; ++qa)
#pragma unroll
        for (int qb = 0; qb < 4; ++qb) {
            int r = crow + 8 * qa, c = ccol + 4 * qb;
            if (r < p.M && c < p.N) C[r + (long long)c * p.ldc] = cmake(acc[qa][qb][0], acc[qa][qb][1]);
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

cudaError_t zgemm_dmma_launch(const ZgemmParams& p, cudaStream_t stream) {
    if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return cudaSuccess;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(zgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid((p.M + TM - 1) / TM, (p.N + TN - 1) / TN, p.batch);
    if (p.K <= 0) {   // nothing to accumulate: C = beta*C
        if (p.beta) return cudaSuccess;
    }
    zgemm_dmma_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t zgemm_simple_launch(const ZgemmParams& p, cudaStream_t stream) {
    if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return cudaSuccess;
    dim3 grid((p.M + 15) / 16, (p.N + 15) / 16, p.batch);
    zgemm_simple_kernel<<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
}
