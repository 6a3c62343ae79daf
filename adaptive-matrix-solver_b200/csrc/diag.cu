// diag.cu -- set-up diagnostics of the reference's MAUS_Solver._diagnose_matrix_initial (AMS:374-404) on the device
// (SURVEY.md 8f-4).  As shipped the reference runs, on the host and before the first step, two np.allclose passes
// (Hermitian / complex-symmetric tests), np.count_nonzero and np.linalg.cond -- a FULL SVD, O(n^3) with a large constant,
// minutes at n = 8192 -- only to pick one of three strategies from the thresholds cond > 1e12 / > 1e6 (AMS:405-413).
//
// Here, on the dense matrix already resident for the candidate steps:
//   maus_diag_dense     one pass over both resident layouts (row-major a_ij and column-major a_ji are both coalesced):
//                       non-zero count, isclose(A, A^H) and isclose(A, A^T) element tests with numpy's formula;
//   maus_cond2_estimate 2-norm condition number sigma_max / sigma_min: power iteration on A^H A with the HBM-bound batched
//                       matvec kernels (A v on the row-major copy, A^H u = conj(A^T conj(u)) on the column-major copy), and
//                       inverse iteration on A^H A through the batched LU of the hot path (solve A^H y = x, then A z = y).
// Both iterations converge from below (sigma_max) / above (sigma_min), i.e. the estimate is a lower bound of cond_2 that is
// tight to a few per cent after the default iteration counts -- enough to reproduce the reference's three-way decision away
// from the thresholds; the host layer (diagnostics.py) states the tolerance.
#include <vector>
#include "ctx.cuh"
#include "vec.cuh"

namespace {

constexpr int DG_NT = 256;

__device__ __forceinline__ bool dg_isclose(cplx a, cplx b, double rtol, double atol) {
    // numpy.isclose(a, b): finite values |a - b| <= atol + rtol |b|; non-finite values only when equal; NaN never
    const bool fa = isfinite(a.x) && isfinite(a.y), fb = isfinite(b.x) && isfinite(b.y);
    if (fa && fb) return hypot(a.x - b.x, a.y - b.y) <= atol + rtol * hypot(b.x, b.y);
    return a.x == b.x && a.y == b.y;
}

// out[0] = non-zeros, out[1] = entries violating isclose(A, A^H), out[2] = entries violating isclose(A, A^T)
__global__ void __launch_bounds__(DG_NT) dg_scan_kernel(const cplx* __restrict__ rm, const cplx* __restrict__ cm, long long n,
                                                        double rtol, double atol, unsigned long long* out) {
    const long long total = n * n;
    unsigned long long nz = 0, bad_h = 0, bad_s = 0;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const cplx a = rm[idx];            // A[i][j], idx = i n + j
        const cplx t = cm[idx];            // cm[j + i n] = A[j][i]
        if (a.x != 0.0 || a.y != 0.0) ++nz;
        if (!dg_isclose(a, cmake(t.x, -t.y), rtol, atol)) ++bad_h;
        if (!dg_isclose(a, t, rtol, atol)) ++bad_s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nz += __shfl_xor_sync(0xffffffffu, nz, o); bad_h += __shfl_xor_sync(0xffffffffu, bad_h, o); bad_s += __shfl_xor_sync(0xffffffffu, bad_s, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], nz); atomicAdd(&out[1], bad_h); atomicAdd(&out[2], bad_s); }
}

// nrm[0] = ||v||_2 (scaled like dznrm2) ; v <- v / ||v|| (conjugated first when conj_in) ; one CTA
__global__ void __launch_bounds__(1024) dg_normalise_kernel(cplx* __restrict__ v, long long n, int conj_in, double* nrm) {
    __shared__ double sh[32];
    __shared__ double s_amax, s_nv;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    double amax = 0.0;
    for (long long i = t; i < n; i += blockDim.x) amax = fmax(amax, fmax(fabs(v[i].x), fabs(v[i].y)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) sh[warp] = amax;
    __syncthreads();
    if (t == 0) { double m = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, sh[w]); s_amax = m; }
    __syncthreads();
    amax = s_amax;
    double ss = 0.0;
    if (amax > 0.0 && isfinite(amax)) {
        const double inv = 1.0 / amax;
        for (long long i = t; i < n; i += blockDim.x) { const double p = v[i].x * inv, q = v[i].y * inv; ss = fma(p, p, ss); ss = fma(q, q, ss); }
    }
    ss = warp_sum(ss);
    __syncthreads();
    if (lane == 0) sh[warp] = ss;
    __syncthreads();
    if (t == 0) {
        double s = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
        s_nv = (amax > 0.0 && isfinite(amax)) ? amax * sqrt(s) : amax;
        nrm[0] = s_nv;
    }
    __syncthreads();
    const double nv = s_nv;
    if (nv > 0.0 && isfinite(nv))
        for (long long i = t; i < n; i += blockDim.x) { cplx a = v[i]; v[i] = cmake(a.x / nv, (conj_in ? -a.y : a.y) / nv); }
}

}  // namespace

extern "C" int maus_diag_dense(maus_ctx* ctx, double rtol, double atol, int64_t* nonzeros, int32_t* is_hermitian,
                               int32_t* is_complex_symmetric) {
    if (!ctx) return MAUS_E_ARG;
    MatrixSlot& s = ctx->slot[0];
    if (ctx->n <= 0 || !s.dense) return maus_fail(ctx, MAUS_E_STATE, "maus_diag_dense: a dense matrix must be resident (maus_set_dense)");
    cudaSetDevice(ctx->device);
    MausNvtxRange range("maus.diag.scan");
    unsigned long long* d = nullptr;
    MAUS_CUDA(ctx, cudaMalloc(&d, 3 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), ctx->stream);
    unsigned long long h[3] = {0, 0, 0};
    if (e == cudaSuccess) {
        dg_scan_kernel<<<MAUS_SM_COUNT_B200 * 8, DG_NT, 0, ctx->stream>>>(s.rm, s.cm, ctx->n, rtol, atol, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    ctx->launches += 1;
    if (e != cudaSuccess) return maus_fail(ctx, MAUS_E_CUDA, "maus_diag_dense", e);
    if (nonzeros) *nonzeros = (int64_t)h[0];
    if (is_hermitian) *is_hermitian = h[1] == 0 ? 1 : 0;
    if (is_complex_symmetric) *is_complex_symmetric = h[2] == 0 ? 1 : 0;
    return MAUS_OK;
}

extern "C" int maus_cond2_estimate(maus_ctx* ctx, int power_iters, int inverse_iters, const double* start /* [n] complex or NULL */,
                                   double* sigma_max, double* sigma_min, int32_t* lu_status) {
    if (!ctx || power_iters < 1 || inverse_iters < 0) return maus_fail(ctx, MAUS_E_ARG, "maus_cond2_estimate: bad argument");
    MatrixSlot& s = ctx->slot[0];
    if (ctx->n <= 0 || !s.dense) return maus_fail(ctx, MAUS_E_STATE, "maus_cond2_estimate: a dense matrix must be resident (maus_set_dense)");
    cudaSetDevice(ctx->device);
    int rc = maus_ensure_population(ctx, 2); if (rc) return rc;
    MausNvtxRange range("maus.diag.cond");
    const long long n = ctx->n;
    cudaStream_t st = ctx->stream;
    cplx *v = ctx->V, *u = ctx->Y, *w = ctx->X;             // three length-n scratch vectors of the population buffers
    double* dn = ctx->vnorm2;                               // device scalar
    std::vector<cplx> hv((size_t)n);
    if (start) memcpy(hv.data(), start, (size_t)n * sizeof(cplx));
    else for (long long i = 0; i < n; ++i) {                // fixed quasi-random start (no zero component, no structure)
        const double a = 0.5 + 0.5 * sin(12.9898 * (double)(i + 1)), b = cos(78.233 * (double)(i + 1));
        hv[(size_t)i] = cmake(a, 0.37 * b);
    }
    MAUS_CUDA(ctx, cudaMemcpyAsync(v, hv.data(), (size_t)n * sizeof(cplx), cudaMemcpyHostToDevice, st));
    dg_normalise_kernel<<<1, 1024, 0, st>>>(v, n, 0, dn);
    double smax = 0.0, smin = 0.0;
    // ---- sigma_max: v <- A^H (A v) normalised; the last ||A^H u|| with ||u|| = 1 is the estimate ----
    for (int it = 0; it < power_iters; ++it) {
        MAUS_CUDA(ctx, vec_gemv_rowmajor(s.rm, v, n, u, n, (int)n, 1, st));                 // u = A v
        dg_normalise_kernel<<<1, 1024, 0, st>>>(u, n, 1, dn);                              // u = conj(u / ||u||)
        MAUS_CUDA(ctx, vec_gemv_rowmajor(s.cm, u, n, v, n, (int)n, 1, st));                 // v = A^T conj(u)
        dg_normalise_kernel<<<1, 1024, 0, st>>>(v, n, 1, dn);                              // v = conj(.) / ||.|| = A^H u / ||A^H u||
        ctx->launches += 4;
    }
    MAUS_CUDA(ctx, cudaMemcpyAsync(&smax, dn, sizeof(double), cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaStreamSynchronize(st));
    // ---- sigma_min: x <- (A^H A)^-1 x = A^-1 (A^-H x) through the batched LU of the hot path (no Psi: sigma = psi = 0) ----
    int status = 0;
    if (inverse_iters > 0) {
        MAUS_CUDA(ctx, cudaMemsetAsync(ctx->sigma, 0, sizeof(cplx), st));
        MAUS_CUDA(ctx, cudaMemsetAsync(ctx->psi, 0, sizeof(double), st));
        MAUS_CUDA(ctx, cudaMemcpyAsync(v, hv.data(), (size_t)n * sizeof(cplx), cudaMemcpyHostToDevice, st));
        dg_normalise_kernel<<<1, 1024, 0, st>>>(v, n, 0, dn);
        double mu = 0.0;
        for (int it = 0; it < inverse_iters && status == 0; ++it) {
            MAUS_CUDA(ctx, cudaMemsetAsync(ctx->status, 0, sizeof(int), st));
            ctx->lu_conj_transpose = true;
            rc = maus_lu_solve(ctx, 1, ctx->sigma, ctx->psi, nullptr, nullptr, v, n, w, ctx->status);     // w = A^-H v
            ctx->lu_conj_transpose = false;
            if (rc) return rc;
            rc = maus_lu_solve(ctx, 1, ctx->sigma, ctx->psi, nullptr, nullptr, w, n, v, ctx->status);     // v = A^-1 w
            if (rc) return rc;
            dg_normalise_kernel<<<1, 1024, 0, st>>>(v, n, 0, dn);
            ctx->launches += 1;
            MAUS_CUDA(ctx, cudaMemcpyAsync(&mu, dn, sizeof(double), cudaMemcpyDeviceToHost, st));
            MAUS_CUDA(ctx, cudaMemcpyAsync(&status, ctx->status, sizeof(int), cudaMemcpyDeviceToHost, st));
            MAUS_CUDA(ctx, cudaStreamSynchronize(st));
        }
        // ||(A^H A)^-1 x|| -> 1 / sigma_min^2 ; a zero pivot or a non-finite solve means numerically singular
        smin = (status != 0 || !(mu > 0.0) || !isfinite(mu)) ? 0.0 : 1.0 / sqrt(mu);
    }
    if (sigma_max) *sigma_max = smax;
    if (sigma_min) *sigma_min = smin;
    if (lu_status) *lu_status = status;
    return MAUS_OK;
}
