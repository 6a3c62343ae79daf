// Ablation of the 4-candidate packed SpMM (spmv.cu: csr_spmm_packed_kernel<4, 3, 4>) on the K5 shape: which part of the row
// loop costs the factor 3.4 between the kernel (0.33 ms) and the bare L2 gather of the same 1.34 GB (0.096 ms, l2_gather.cu)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o spmm_ablate spmm_ablate.cu && ./spmm_ablate
// n = 1M rows, exactly 21 entries per row at uniformly random columns (or the K5 structure from dump_k5_csr.py as argv[1]),
// P[j][4] interleaved candidates.  Scratch.  Measured (B200, round 2b, profiles/r02/spmm_ablate_r02c.jsonl):
//   production shape 0.213 ms | no value stream 0.166 | no reduction / store 0.214 | neither 0.148 | not pipelined, 8 CTAs/SM 0.223
//   lean row loop k2 (107 instead of 132 instructions per row) 0.214 - 0.227 | k3 = L2 prefetch of the entries two rows ahead 0.258
//   L2 state (P rewritten / evicted before every launch) 0.218 / 0.226 -- i.e. the loop sits on a latency floor of ~0.21 ms that
//   neither fewer instructions nor deeper prefetch moves; the library kernel WITH its chunk loop (171 instructions) took 0.30 ms.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
typedef double2 cplx;
__device__ __forceinline__ void cfma(cplx& a, cplx b, cplx c) {
    a.x = fma(b.x, c.x, a.x); a.x = fma(-b.y, c.y, a.x); a.y = fma(b.x, c.y, a.y); a.y = fma(b.y, c.x, a.y);
}
constexpr int NT = 256, SUBS = 8, CB = 4, U = 3;
// VALS: stream the matrix values; REDUCE: xor tree + store per row (else one store at the end); PIPE: next row's entries prefetched
template <bool VALS, bool REDUCE, bool PIPE, int MINB>
__global__ void __launch_bounds__(NT, MINB) k(const long long* __restrict__ rowptr, const int* __restrict__ colidx, const cplx* __restrict__ vals,
                                             const cplx* __restrict__ P, cplx* __restrict__ Y, long long n) {
    constexpr int LPR = SUBS * CB, GPB = NT / LPR;
    const int l = threadIdx.x % LPR, sub = l / CB, c = l % CB;
    const long long stride = (long long)gridDim.x * GPB;
    long long row = (long long)blockIdx.x * GPB + threadIdx.x / LPR;
    auto load_entries = [&](long long k0, long long k1, cplx* a, int* j) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long kk = k0 + sub + u * SUBS;
            const bool ok = kk < k1;
            a[u] = (ok && VALS) ? __ldcs(&vals[kk]) : make_double2(ok ? 1.0 : 0.0, 0.0);
            j[u] = ok ? __ldcs(&colidx[kk]) : -1;
        }
    };
    cplx tot = make_double2(0.0, 0.0);
    if (PIPE) {
        long long k0 = 0, k1 = 0, k0n = 0, k1n = 0;
        cplx an[U]; int jn[U];
        if (row < n) { k0 = rowptr[row]; k1 = rowptr[row + 1]; }
        load_entries(k0, k1, an, jn);
        long long rown = row + stride;
        if (rown < n) { k0n = rowptr[rown]; k1n = rowptr[rown + 1]; }
        while (__any_sync(0xffffffffu, row < n)) {
            cplx a[U]; int j[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { a[u] = an[u]; j[u] = jn[u]; }
            cplx v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = (j[u] >= 0) ? __ldg(&P[(long long)j[u] * CB + c]) : make_double2(0.0, 0.0);
            const long long rown2 = rown + stride;
            long long k0nn = 0, k1nn = 0;
            load_entries(k0n, k1n, an, jn);
            if (rown2 < n) { k0nn = rowptr[rown2]; k1nn = rowptr[rown2 + 1]; }
            cplx acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < U; ++u) cfma(acc, a[u], v[u]);
            if (REDUCE) {
#pragma unroll
                for (int o = SUBS / 2; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o * CB); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o * CB); }
                if (sub == 0 && row < n) Y[(long long)c * n + row] = acc;
            } else { tot.x += acc.x; tot.y += acc.y; }
            row = rown; rown = rown2; k0 = k0n; k1 = k1n; k0n = k0nn; k1n = k1nn;
        }
    } else {
        for (; row < n; row += stride) {
            const long long k0 = rowptr[row], k1 = rowptr[row + 1];
            cplx a[U]; int j[U]; cplx v[U];
            load_entries(k0, k1, a, j);
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = (j[u] >= 0) ? __ldg(&P[(long long)j[u] * CB + c]) : make_double2(0.0, 0.0);
            cplx acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < U; ++u) cfma(acc, a[u], v[u]);
            if (REDUCE) {
#pragma unroll
                for (int o = SUBS / 2; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o * CB); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o * CB); }
                if (sub == 0) Y[(long long)c * n + row] = acc;
            } else { tot.x += acc.x; tot.y += acc.y; }
        }
    }
    if (!REDUCE && tot.x == 123.456) Y[threadIdx.x] = tot;
}
// Leaner row loop for matrices whose rows fit one chunk (<= 24 entries): one warp per row, per-row base pointers + constant
// lane offsets instead of 64-bit index arithmetic per entry, two register sets used alternately (no copies), no vote
struct RowSet { cplx a[U]; int j[U]; };
template <int MINB>
__global__ void __launch_bounds__(NT, MINB) k2(const long long* __restrict__ rowptr, const int* __restrict__ colidx, const cplx* __restrict__ vals,
                                              const cplx* __restrict__ P, cplx* __restrict__ Y, int n) {
    const int lane = threadIdx.x & 31, sub = lane >> 2, c = lane & 3;
    const int stride = gridDim.x * (NT / 32);
    int row = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    const cplx* Pc = P + c;
    cplx* Yc = Y + (long long)c * n;
    auto fetch = [&](long long k0, int len, RowSet& s) {
        const cplx* vb = vals + k0 + sub;
        const int* cb = colidx + k0 + sub;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool ok = sub + u * SUBS < len;
            s.a[u] = ok ? __ldcs(vb + u * SUBS) : make_double2(0.0, 0.0);
            s.j[u] = ok ? __ldcs(cb + u * SUBS) : -1;
        }
    };
    auto finish = [&](const RowSet& s, const cplx* v, int r) {
        cplx acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < U; ++u) cfma(acc, s.a[u], v[u]);
#pragma unroll
        for (int o = SUBS / 2; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o * CB); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o * CB); }
        if (sub == 0) Yc[r] = acc;
    };
    RowSet A, B;
    long long k0 = 0, k1 = 0, k0n = 0, k1n = 0;
    if (row < n) { k0 = rowptr[row]; k1 = rowptr[row + 1]; }
    fetch(k0, (int)(k1 - k0), A);
    if (row + stride < n) { k0n = rowptr[row + stride]; k1n = rowptr[row + stride + 1]; }
    // step(cur, nxt): gathers of the current row, entries of the next one, row pointers of the one after it
    auto step = [&](RowSet& cur, RowSet& nxt) {
        cplx v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (cur.j[u] >= 0) ? __ldg(Pc + (long long)cur.j[u] * CB) : make_double2(0.0, 0.0);
        fetch(k0n, (int)(k1n - k0n), nxt);
        const int r2 = row + 2 * stride;
        k0n = 0; k1n = 0;
        if (r2 < n) { k0n = rowptr[r2]; k1n = rowptr[r2 + 1]; }
        finish(cur, v, row);
        row += stride;
    };
    while (row < n) {
        step(A, B);
        if (row >= n) break;
        step(B, A);
    }
}
// production shape + L2 prefetch of the entries two rows ahead (the register prefetch one row ahead then hits L2 instead of HBM)
template <int MINB>
__global__ void __launch_bounds__(NT, MINB) k3(const long long* __restrict__ rowptr, const int* __restrict__ colidx, const cplx* __restrict__ vals,
                                              const cplx* __restrict__ P, cplx* __restrict__ Y, long long n) {
    constexpr int LPR = SUBS * CB, GPB = NT / LPR;
    const int l = threadIdx.x % LPR, sub = l / CB, c = l % CB;
    const long long stride = (long long)gridDim.x * GPB;
    long long row = (long long)blockIdx.x * GPB + threadIdx.x / LPR;
    auto load_entries = [&](long long k0, long long k1, cplx* a, int* j) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long kk = k0 + sub + u * SUBS;
            const bool ok = kk < k1;
            a[u] = ok ? __ldcs(&vals[kk]) : make_double2(0.0, 0.0);
            j[u] = ok ? __ldcs(&colidx[kk]) : -1;
        }
    };
    long long k0n = 0, k1n = 0, k0nn = 0, k1nn = 0;
    cplx an[U]; int jn[U];
    { long long k0 = 0, k1 = 0; if (row < n) { k0 = rowptr[row]; k1 = rowptr[row + 1]; } load_entries(k0, k1, an, jn); }
    long long rown = row + stride, rown2 = row + 2 * stride;
    if (rown < n) { k0n = rowptr[rown]; k1n = rowptr[rown + 1]; }
    if (rown2 < n) { k0nn = rowptr[rown2]; k1nn = rowptr[rown2 + 1]; }
    while (__any_sync(0xffffffffu, row < n)) {
        cplx a[U]; int j[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { a[u] = an[u]; j[u] = jn[u]; }
        cplx v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (j[u] >= 0) ? __ldg(&P[(long long)j[u] * CB + c]) : make_double2(0.0, 0.0);
        // L2 prefetch of the entries two rows ahead: 4 lanes cover the (<= 384 B of values, <= 96 B of indices) of that row
        if (l < 3 && k0nn + l * 8 < k1nn) asm volatile("prefetch.global.L2 [%0];" ::"l"(vals + k0nn + l * 8));
        if (l == 3 && k0nn < k1nn) asm volatile("prefetch.global.L2 [%0];" ::"l"(colidx + k0nn));
        const long long rown3 = rown2 + stride;
        long long k0n3 = 0, k1n3 = 0;
        load_entries(k0n, k1n, an, jn);
        if (rown3 < n) { k0n3 = rowptr[rown3]; k1n3 = rowptr[rown3 + 1]; }
        cplx acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < U; ++u) cfma(acc, a[u], v[u]);
#pragma unroll
        for (int o = SUBS / 2; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o * CB); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o * CB); }
        if (sub == 0 && row < n) Y[(long long)c * n + row] = acc;
        row = rown; rown = rown2; rown2 = rown3; k0n = k0nn; k1n = k1nn; k0nn = k0n3; k1nn = k1n3;
    }
}
template <int MINB>
static void run3(const char* tag, const long long* rp, const int* ci, const cplx* va, const cplx* P, cplx* Y, long long n) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int per_sm = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3<MINB>, NT, 0);
    const unsigned grid = 148u * per_sm;
    for (int r = 0; r < 3; ++r) k3<MINB><<<grid, NT>>>(rp, ci, va, P, Y, n);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) k3<MINB><<<grid, NT>>>(rp, ci, va, P, Y, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("{\"variant\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.4f}\n", tag, per_sm, ms);
}
template <int MINB>
static void run2(const char* tag, const long long* rp, const int* ci, const cplx* va, const cplx* P, cplx* Y, long long n) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int per_sm = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2<MINB>, NT, 0);
    const unsigned grid = 148u * per_sm;
    for (int r = 0; r < 3; ++r) k2<MINB><<<grid, NT>>>(rp, ci, va, P, Y, (int)n);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) k2<MINB><<<grid, NT>>>(rp, ci, va, P, Y, (int)n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("{\"variant\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.4f}\n", tag, per_sm, ms);
}
template <bool VALS, bool REDUCE, bool PIPE, int MINB>
static void run(const char* tag, const long long* rp, const int* ci, const cplx* va, const cplx* P, cplx* Y, long long n, int ctas_per_sm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int per_sm = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k<VALS, REDUCE, PIPE, MINB>, NT, 0);
    if (ctas_per_sm > 0 && ctas_per_sm < per_sm) per_sm = ctas_per_sm;
    const unsigned grid = 148u * per_sm;
    for (int r = 0; r < 3; ++r) k<VALS, REDUCE, PIPE, MINB><<<grid, NT>>>(rp, ci, va, P, Y, n);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) k<VALS, REDUCE, PIPE, MINB><<<grid, NT>>>(rp, ci, va, P, Y, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("{\"variant\": \"%s\", \"values\": %d, \"reduce_store\": %d, \"pipelined\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f}\n", tag, (int)VALS, (int)REDUCE, (int)PIPE, per_sm, ms);
}
int main(int argc, char** argv) {
    long long n = 1000000; const int per = 21; long long nnz = n * per;
    std::vector<long long> rp; std::vector<int> ci;
    if (argc > 1) {
        // CSR structure dumped by profiles/microbench/dump_k5_csr.py: int64 n, int64 nnz, int64 rowptr[n + 1], int32 colidx[nnz]
        FILE* f = fopen(argv[1], "rb");
        if (!f) { printf("cannot open %s\n", argv[1]); return 1; }
        if (fread(&n, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1) return 1;
        rp.resize(n + 1); ci.resize(nnz);
        if (fread(rp.data(), 8, n + 1, f) != (size_t)(n + 1) || fread(ci.data(), 4, nnz, f) != (size_t)nnz) return 1;
        fclose(f);
        printf("{\"matrix\": \"%s\", \"n\": %lld, \"nnz\": %lld}\n", argv[1], n, nnz);
    } else {
        rp.resize(n + 1); ci.resize(nnz);
        unsigned long long s = 88172645463325252ULL;
        for (long long i = 0; i <= n; ++i) rp[i] = i * per;
        for (long long i = 0; i < nnz; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; ci[i] = (int)(s % (unsigned long long)n); }
    }
    long long* d_rp; int* d_ci; cplx *d_va, *d_P, *d_Y;
    cudaMalloc(&d_rp, (n + 1) * 8); cudaMalloc(&d_ci, nnz * 4); cudaMalloc(&d_va, nnz * 16); cudaMalloc(&d_P, n * 64); cudaMalloc(&d_Y, n * 64);
    cudaMemcpy(d_rp, rp.data(), (n + 1) * 8, cudaMemcpyHostToDevice); cudaMemcpy(d_ci, ci.data(), nnz * 4, cudaMemcpyHostToDevice);
    {   // random (non-zero) values and vectors, as in the real run
        std::vector<cplx> hv(nnz), hp(n * 4);
        unsigned long long t = 1234567ULL;
        auto rnd = [&]() { t ^= t << 13; t ^= t >> 7; t ^= t << 17; return (double)(t >> 11) * (1.0 / 9007199254740992.0) - 0.5; };
        for (long long i = 0; i < nnz; ++i) hv[i] = make_double2(rnd(), rnd());
        for (long long i = 0; i < n * 4; ++i) hp[i] = make_double2(rnd(), rnd());
        cudaMemcpy(d_va, hv.data(), nnz * 16, cudaMemcpyHostToDevice); cudaMemcpy(d_P, hp.data(), n * 64, cudaMemcpyHostToDevice);
    }
    run<true, true, true, 4>("production shape (64 registers, 4 CTAs/SM)", d_rp, d_ci, d_va, d_P, d_Y, n, 0);
    run<false, true, true, 4>("no value stream", d_rp, d_ci, d_va, d_P, d_Y, n, 0);
    run<true, false, true, 4>("no per-row reduction / store", d_rp, d_ci, d_va, d_P, d_Y, n, 0);
    run<false, false, true, 4>("neither", d_rp, d_ci, d_va, d_P, d_Y, n, 0);
    run<true, true, false, 8>("not pipelined, 8 CTAs/SM", d_rp, d_ci, d_va, d_P, d_Y, n, 0);
    run<false, false, false, 8>("not pipelined, no values, no reduction, 8 CTAs/SM", d_rp, d_ci, d_va, d_P, d_Y, n, 0);
    run<true, true, true, 2>("production shape, register cap lifted (2 CTAs/SM)", d_rp, d_ci, d_va, d_P, d_Y, n, 0);
    run3<4>("production shape + L2 prefetch two rows ahead (k3), 4 CTAs/SM", d_rp, d_ci, d_va, d_P, d_Y, n);
    run3<3>("k3, 3 CTAs/SM", d_rp, d_ci, d_va, d_P, d_Y, n);
    run2<4>("lean row loop (k2), 4 CTAs/SM", d_rp, d_ci, d_va, d_P, d_Y, n);
    run2<5>("lean row loop (k2), 5 CTAs/SM", d_rp, d_ci, d_va, d_P, d_Y, n);
    run2<6>("lean row loop (k2), 6 CTAs/SM", d_rp, d_ci, d_va, d_P, d_Y, n);
    // L2 state as in the solver: (a) P rewritten right before every SpMM (the pack kernel), (b) 256 MB of other traffic between
    // two SpMMs (the vector passes of the GMRES step); every SpMM launch timed with its own event pair
    {
        cplx* d_other; cudaMalloc(&d_other, (size_t)256 << 20);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int mode = 0; mode < 3; ++mode) {
            float tot = 0.f;
            for (int r = 0; r < 12; ++r) {
                if (mode == 1 || mode == 2) cudaMemsetAsync(d_other, r, (size_t)256 << 20);
                if (mode == 2) cudaMemcpyAsync(d_P, d_other, n * 64, cudaMemcpyDeviceToDevice);     // P freshly written (dirty in L2)
                cudaEventRecord(e0);
                k<true, true, true, 4><<<148 * 4, NT>>>(d_rp, d_ci, d_va, d_P, d_Y, n);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (r >= 2) tot += ms;
            }
            printf("{\"variant\": \"production shape, one event pair per launch, %s\", \"ms\": %.4f}\n",
                   mode == 0 ? "back to back" : (mode == 1 ? "256 MB written elsewhere before each launch (P evicted)" : "P rewritten before each launch"), tot / 10);
        }
    }
    return 0;
}
