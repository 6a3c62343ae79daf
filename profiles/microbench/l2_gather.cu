// Random-gather bandwidth out of L2 (B200): the ceiling of the batched SpMM's x-gathers (spmv.cu).  Scratch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather l2_gather.cu && ./l2_gather
// A table of `seg` bytes per index (64 B = 4 candidates, 128 B = 8) and `mb` megabytes is gathered at uniformly random indices
// (precomputed index stream, read coalesced like colidx); seg / 16 neighbouring lanes fetch one segment, 16 B each, `U` gathers
// in flight per lane.  Reported: gathered bytes / time (the index stream adds 4 B per segment on top).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
template <int SEG, int U>
__global__ void __launch_bounds__(256) gather(const double2* __restrict__ tab, const int* __restrict__ idx, long long nidx, double2* out) {
    constexpr int LPS = SEG / 16;                          // lanes per segment
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long g = t / LPS; const int l = (int)(t % LPS);
    const long long ngroups = (long long)gridDim.x * blockDim.x / LPS;
    double2 acc = make_double2(0.0, 0.0);
    for (long long k = g * U; k + U <= nidx; k += ngroups * U) {
        int j[U]; double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) j[u] = __ldcs(&idx[k + u]);
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(&tab[(long long)j[u] * LPS + l]);
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
    }
    if (acc.x == 123.456) out[t] = acc;
}
template <int SEG, int U>
static void run(int mb, long long nidx, const char* tag) {
    const long long nseg = (long long)mb * 1048576 / SEG;
    double2* tab; int* idx; double2* out;
    cudaMalloc(&tab, (size_t)mb << 20); cudaMemset(tab, 0, (size_t)mb << 20);
    std::vector<int> h(nidx);
    unsigned long long s = 88172645463325252ULL;
    for (long long i = 0; i < nidx; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % (unsigned long long)nseg); }
    cudaMalloc(&idx, nidx * 4); cudaMemcpy(idx, h.data(), nidx * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 148 * 8 * 256 * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) gather<SEG, U><<<148 * 8, 256>>>(tab, idx, nidx, out);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int rep = 0; rep < reps; ++rep) gather<SEG, U><<<148 * 8, 256>>>(tab, idx, nidx, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("{\"segment_bytes\": %d, \"table_mb\": %d, \"gathers_in_flight_per_lane\": %d, \"segments\": %lld, \"ms\": %.4f, \"gathered_gbs\": %.1f, \"note\": \"%s\"}\n",
           SEG, mb, U, nidx, ms, (double)nidx * SEG / (ms * 1e-3) / 1e9, tag);
    cudaFree(tab); cudaFree(idx); cudaFree(out);
}
int main() {
    const long long nidx = 21000000;       // = nnz of the K5 matrix
    run<64, 3>(16, nidx, "4 candidates, n = 250k");
    run<64, 3>(64, nidx, "4 candidates, n = 1M (the K5 pass)");
    run<64, 6>(64, nidx, "same, 6 gathers in flight");
    run<128, 3>(32, nidx, "8 candidates, n = 250k");
    run<128, 3>(128, nidx, "8 candidates, n = 1M: the table exceeds the 126 MB L2");
    run<16, 3>(16, nidx, "1 candidate, n = 1M");
    return 0;
}
