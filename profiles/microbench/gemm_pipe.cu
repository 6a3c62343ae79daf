// Microbenchmark: progressively add the pieces of the real zgemm kernel around the 99%-efficient inner loop. Scratch.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s @%d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)
typedef double2 cplx;
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t ph) {
    uint32_t d; do { asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(d) : "r"(s32(b)), "r"(ph) : "memory"); } while (!d);
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)) : "memory");
}
constexpr int KC = 8, LDSA = 132, KCB = 64, LDSB = 66, STAGES = 4, AB = 8, NCONS = 8;
constexpr int A_STAGE = KC * LDSA, B_STAGE = 64 * LDSB;
// MODE 0: consumers only, no barriers.  1: + producer/consumer mbarrier handshake (no copies).  2: + A/B bulk copies.
// 3: + epilogue C load/add/store (global).   4: mode 3 but no copies (handshake + epilogue).
template <int MODE>
__global__ void __launch_bounds__(288, 1) k_pipe(const cplx* __restrict__ gA, const cplx* __restrict__ gB, cplx* gC, int tiles, int KT, long long ldc, double* out) {
    extern __shared__ __align__(128) unsigned char raw[];
    cplx* sA = (cplx*)raw; cplx* sB = sA + STAGES * A_STAGE;
    uint64_t* full = (uint64_t*)(sB + 2 * B_STAGE); uint64_t* empty = full + STAGES; uint64_t* bfull = empty + STAGES; uint64_t* bempty = bfull + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < STAGES * A_STAGE + 2 * B_STAGE; i += blockDim.x) sA[i] = make_double2(1e-3 * (i % 97), -2e-3 * (i % 89));
    if (threadIdx.x == 0) { for (int s = 0; s < STAGES; s++) { mb_init(&full[s], 1); mb_init(&empty[s], NCONS); } for (int s = 0; s < 2; s++) { mb_init(&bfull[s], 1); mb_init(&bempty[s], NCONS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    constexpr int APB = KCB / KC;
    if (warp == NCONS) {
        if (MODE == 0) return;
        long long it = 0, ib = 0;
        for (int tile = 0; tile < tiles; ++tile) {
            const cplx* A = gA + (size_t)(blockIdx.x * 7 + tile) % 61 * 128; const cplx* B = gB + (size_t)(blockIdx.x * 3 + tile) % 53 * 64 * 4096;
            for (int kt = 0; kt < KT; ++kt, ++it) {
                if (kt % APB == 0) {
                    int sb = ib % 2; if (ib >= 2) mb_wait(&bempty[sb], ((ib / 2) - 1) & 1);
                    if (MODE == 2 || MODE == 3) { if (lane == 0) mb_expect(&bfull[sb], 64 * KCB * 16); __syncwarp();
                        for (int j = lane; j < 64; j += 32) bulk(sB + sb * B_STAGE + j * LDSB, B + kt * KC + (size_t)j * 4096, KCB * 16, &bfull[sb]); }
                    else if (lane == 0) mb_arrive(&bfull[sb]);
                    ++ib;
                }
                int s = it % STAGES; if (it >= STAGES) mb_wait(&empty[s], ((it / STAGES) - 1) & 1);
                if (MODE == 2 || MODE == 3) { if (lane == 0) mb_expect(&full[s], KC * 128 * 16); __syncwarp();
                    if (lane < KC) bulk(sA + s * A_STAGE + lane * LDSA, A + (size_t)(kt * KC + lane) * 8192, 128 * 16, &full[s]); }
                else if (lane == 0) mb_arrive(&full[s]);
            }
        }
        return;
    }
    const int g = lane >> 2, t = lane & 3, wm = warp >> 2, wn = warp & 3;
    const int arow = wm * 64 + g, bcol = wn * 16 + (g >> 1), comp = (t ^ g) & 1, kh = t >> 1;
    const long long sflip = ((long long)((t & 1) && !(g & 1))) << 63;
    long long it = 0, ib = 0; double chk = 0;
    for (int tile = 0; tile < tiles; ++tile) {
        double acc[AB][4][2];
        for (int a = 0; a < AB; a++) for (int b = 0; b < 4; b++) { acc[a][b][0] = 0; acc[a][b][1] = 0; }
        const double* b_s = (const double*)sB; int sb = 0;
        for (int kt = 0; kt < KT; ++kt, ++it) {
            if (MODE != 0 && kt % APB == 0) { sb = ib % 2; mb_wait(&bfull[sb], (ib / 2) & 1); b_s = (const double*)(sB + sb * B_STAGE); ++ib; }
            int s = it % STAGES; if (MODE != 0) mb_wait(&full[s], (it / STAGES) & 1);
            const double* a_s = (const double*)(sA + s * A_STAGE); const int kboff = (kt % APB) * KC;
#pragma unroll
            for (int ks = 0; ks < KC / 2; ++ks) {
                const int kc = 2 * ks + kh; double af[AB], bf[4];
#pragma unroll
                for (int q = 0; q < AB; ++q) af[q] = a_s[2 * (kc * LDSA + arow + 8 * q) + (t & 1)];
#pragma unroll
                for (int q = 0; q < 4; ++q) bf[q] = __longlong_as_double(__double_as_longlong(b_s[2 * ((bcol + 4 * q) * LDSB + kboff + kc) + comp]) ^ sflip);
#pragma unroll
                for (int qa = 0; qa < AB; ++qa)
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb) dmma(acc[qa][qb][0], acc[qa][qb][1], af[qa], bf[qb]);
            }
            if (MODE != 0) { __syncwarp(); if (lane == 0) { mb_arrive(&empty[s]); if (kt % APB == APB - 1 || kt == KT - 1) mb_arrive(&bempty[sb]); } }
        }
        if (MODE >= 3) {
            cplx* C = gC + ((size_t)(blockIdx.x + 148 * (tile % 8)) * 64) * ldc;     // distinct 128 x 64 tiles
#pragma unroll
            for (int qa = 0; qa < AB; ++qa)
#pragma unroll
                for (int qb = 0; qb < 4; ++qb) {
                    cplx* d = &C[wm * 64 + g + 8 * qa + (size_t)(wn * 16 + t + 4 * qb) * ldc];
                    cplx c0 = *d; *d = make_double2(acc[qa][qb][0] + c0.x, acc[qa][qb][1] + c0.y);
                }
        } else { for (int a = 0; a < AB; a++) for (int b = 0; b < 4; b++) chk += acc[a][b][0] + acc[a][b][1]; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = chk;
}
template <int MODE> int run(const char* name, cplx* gA, cplx* gB, cplx* gC, double* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int tiles = 200, KT = 16; size_t smem = (STAGES * A_STAGE + 2 * B_STAGE) * 16 + 128;
    CK(cudaFuncSetAttribute(k_pipe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k_pipe<MODE><<<148, 288, smem>>>(gA, gB, gC, tiles, KT, 4096, out); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double fl = 148.0 * tiles * 8.0 * 128 * 64 * 128;
    printf("%-58s %.3f ms  %.2f TFLOP/s  (%.1f us/tile)\n", name, best, fl / best * 1e-9, best * 1e3 / tiles);
    return 0;
}
int main() {
    cplx *gA, *gB, *gC; double* out;
    CK(cudaMalloc(&gA, (size_t)8192 * 256 * 16)); CK(cudaMalloc(&gB, (size_t)4096 * 64 * 64 * 16)); CK(cudaMalloc(&gC, (size_t)4096 * 64 * 148 * 8 * 16)); CK(cudaMalloc(&out, 148 * 288 * 8));
    CK(cudaMemset(gA, 0, (size_t)8192 * 256 * 16)); CK(cudaMemset(gB, 0, (size_t)4096 * 64 * 64 * 16)); CK(cudaMemset(gC, 0, (size_t)4096 * 64 * 148 * 8 * 16));
    run<0>("0 consumers only", gA, gB, gC, out);
    run<1>("1 + mbarrier handshake with producer warp", gA, gB, gC, out);
    run<2>("2 + A/B bulk copies (L2 resident)", gA, gB, gC, out);
    run<4>("4 handshake + epilogue C load/add/store, no copies", gA, gB, gC, out);
    run<3>("3 everything", gA, gB, gC, out);
    return 0;
}
