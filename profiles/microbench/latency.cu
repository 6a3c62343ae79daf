// Dependent-chain latencies of the instructions on the LU panel's per-column critical path (B200).  Scratch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu && ./latency
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>
typedef double2 cplx;
__device__ __forceinline__ cplx crecip_smith(cplx a) {
    if (fabs(a.x) >= fabs(a.y)) { double r = a.y / a.x, d = a.x + a.y * r; return make_double2(1.0 / d, -r / d); }
    else { double r = a.x / a.y, d = a.x * r + a.y; return make_double2(r / d, -1.0 / d); }
}
template <int OP>
__global__ void k(long long* out, double* sink, int iters, int T) {
    __shared__ double sm[64];
    double x = 1.0 + threadIdx.x * 1e-9, y = 0.5;
    unsigned u = threadIdx.x + 7; int r = threadIdx.x;
    sm[threadIdx.x & 63] = x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) { x = fma(x, y, 0.25); }
        else if (OP == 1) { x = x + y; }
        else if (OP == 2) { u = __reduce_max_sync(0xffffffffu, u) + (unsigned)i; }
        else if (OP == 3) { x = __shfl_xor_sync(0xffffffffu, x, 1) + 0.0; x = __longlong_as_double(__double_as_longlong(x) ^ 1); }
        else if (OP == 4) { x = (x > y) ? y : x; y = __longlong_as_double(__double_as_longlong(y) + 1); }
        else if (OP == 5) { x = 1.0 / x; }
        else if (OP == 6) { cplx c = crecip_smith(make_double2(x, y)); x = c.x; y = c.y; }
        else if (OP == 7) { r = (r + 100000) / T + (r % T); }
        else if (OP == 8) { __syncthreads(); }
        else if (OP == 9) { x = sm[((int)__double_as_longlong(x)) & 63]; }
        else if (OP == 10) { long long t = clock64(); r += (int)t; }
        else if (OP == 11) { u = __shfl_xor_sync(0xffffffffu, u, 1) + 1u; }
        else if (OP == 12) { unsigned b = __ballot_sync(0xffffffffu, u & 1); u = u + __ffs(b); }
        else if (OP == 13) { double rr; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(x)); rr = fma(fma(-x, rr, 1.0), rr, rr); rr = fma(fma(-x, rr, 1.0), rr, rr); x = rr; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = x + y + u + r;
}
int main() {
    long long* out; double* sink; cudaMalloc(&out, 8); cudaMalloc(&sink, 8 * 1024);
    const char* names[] = {"DFMA", "DADD", "REDUX.max.u32", "SHFL f64 (2 x SHFL) + LOP", "DSETP+SEL + IADD64", "f64 divide 1/x", "crecip (Smith)", "int div + mod by runtime T",
                           "__syncthreads (256 threads)", "LDS dependent", "clock64", "SHFL u32 + IADD", "ballot + ffs", "rcp.approx.f64 + 2 Newton"};
    const int iters = 2000;
    auto run = [&](auto kern, int i, int nt) {
        kern<<<1, nt>>>(out, sink, iters, 2048); cudaDeviceSynchronize();
        kern<<<1, nt>>>(out, sink, iters, 2048); cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("%-32s %7.1f cycles per iteration (%d threads)\n", names[i], (double)h / iters, nt);
    };
    run(k<0>, 0, 32); run(k<1>, 1, 32); run(k<2>, 2, 32); run(k<3>, 3, 32); run(k<4>, 4, 32); run(k<5>, 5, 32); run(k<6>, 6, 32);
    run(k<7>, 7, 32); run(k<8>, 8, 256); run(k<8>, 8, 512); run(k<9>, 9, 32); run(k<10>, 10, 32); run(k<11>, 11, 32); run(k<12>, 12, 32); run(k<13>, 13, 32);
    return 0;
}
