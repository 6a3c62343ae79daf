// Microbenchmark of the zgemm consumer inner loop variants (smem-resident operands, no TMA). Scratch.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s @%d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
constexpr int KC = 8, LDSA = 132, LDSB = 10, LDSA2 = 130;
// V=1: LDS.64 per fragment (current kernel).  V=2: register-resident fragments.  V=3: LDS.128 paired mapping.
template <int V, int AB, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k_inner(double* out, int iters) {
    extern __shared__ double2 sm[];
    double2* sA = sm; double2* sB = sm + KC * 136;
    for (int i = threadIdx.x; i < KC * 136 + 64 * LDSB; i += blockDim.x) sm[i] = make_double2(1e-3 * i, -2e-3 * i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    double acc[AB][4][2];
    for (int a = 0; a < AB; a++) for (int b = 0; b < 4; b++) { acc[a][b][0] = 0; acc[a][b][1] = 0; }
    const int arow = wm * 8 * AB + g, bcol = wn * 16 + (g >> 1), comp = (t ^ g) & 1, kh = t >> 1;
    const long long sflip = ((long long)((t & 1) && !(g & 1))) << 63;
    const double* a_s = (const double*)sA; const double* b_s = (const double*)sB;
    double af0[AB], bf0[4];
    for (int q = 0; q < AB; q++) af0[q] = a_s[2 * (kh * LDSA + arow + 8 * q) + (t & 1)];
    for (int q = 0; q < 4; q++) bf0[q] = b_s[2 * ((bcol + 4 * q) * LDSB + kh) + comp];
    for (int it = 0; it < iters; ++it) {
        if (V == 1) {
#pragma unroll
            for (int ks = 0; ks < KC / 2; ++ks) {
                const int kc = 2 * ks + kh;
                double af[AB], bf[4];
#pragma unroll
                for (int q = 0; q < AB; ++q) af[q] = a_s[2 * (kc * LDSA + arow + 8 * q) + (t & 1)];
#pragma unroll
                for (int q = 0; q < 4; ++q) bf[q] = __longlong_as_double(__double_as_longlong(b_s[2 * ((bcol + 4 * q) * LDSB + kc) + comp]) ^ sflip);
#pragma unroll
                for (int qa = 0; qa < AB; ++qa)
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb) dmma(acc[qa][qb][0], acc[qa][qb][1], af[qa], bf[qb]);
            }
        } else if (V == 2) {
#pragma unroll
            for (int ks = 0; ks < KC / 2; ++ks)
#pragma unroll
                for (int qa = 0; qa < AB; ++qa)
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb) dmma(acc[qa][qb][0], acc[qa][qb][1], af0[qa], bf0[qb]);
        } else {
            // paired mapping: lane t handles complex k = 4p + t; step 0 uses re(A), step 1 uses im(A)
#pragma unroll
            for (int p = 0; p < KC / 4; ++p) {
                const int kc = 4 * p + t;
                double2 av[AB], bv[4];
#pragma unroll
                for (int q = 0; q < AB; ++q) av[q] = sA[kc * LDSA2 + arow + 8 * q];
#pragma unroll
                for (int q = 0; q < 4; ++q) bv[q] = sB[(bcol + 4 * q) * LDSB + kc];
                double b0[4], b1[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    b0[q] = (g & 1) ? bv[q].y : bv[q].x;
                    b1[q] = (g & 1) ? bv[q].x : -bv[q].y;
                }
#pragma unroll
                for (int qa = 0; qa < AB; ++qa)
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb) dmma(acc[qa][qb][0], acc[qa][qb][1], av[qa].x, b0[qb]);
#pragma unroll
                for (int qa = 0; qa < AB; ++qa)
#pragma unroll
                    for (int qb = 0; qb < 4; ++qb) dmma(acc[qa][qb][0], acc[qa][qb][1], av[qa].y, b1[qb]);
            }
        }
    }
    double s = 0; for (int a = 0; a < AB; a++) for (int b = 0; b < 4; b++) s += acc[a][b][0] + acc[a][b][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V, int AB, int NW> int run(const char* name, double* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000; size_t smem = (KC * 136 + 64 * LDSB) * 16;
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k_inner<V, AB, NW><<<148, NW * 32, smem>>>(out, iters); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fl = 148.0 * NW * iters * (KC / 2) * AB * 4 * 512.0;
    printf("%-40s %.3f ms  %.2f TFLOP/s\n", name, best, fl / best * 1e-9);
    return 0;
}
int main() {
    double* out; CK(cudaMalloc(&out, 148 * 1024 * 8));
    run<1, 8, 8>("V1 LDS.64  8 warps 64x16", out);
    run<1, 4, 16>("V1 LDS.64 16 warps 32x16", out);
    run<2, 8, 8>("V2 regs    8 warps 64x16", out);
    run<2, 4, 16>("V2 regs   16 warps 32x16", out);
    run<3, 8, 8>("V3 LDS.128 8 warps 64x16", out);
    run<3, 4, 16>("V3 LDS.128 16 warps 32x16", out);
    return 0;
}
