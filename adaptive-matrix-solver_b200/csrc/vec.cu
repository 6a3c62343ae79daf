// vec.cu -- HBM-bound kernels of the candidate step: batched matvec, Rayleigh quotient, mix + normalise, residual.
#include "vec.cuh"
#include "../../include/maus_b200.h"

namespace {

constexpr int RED_NT = 512;

__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];   // fixed order -> deterministic
    return r;
}
__device__ __forceinline__ double block_max(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = fmax(r, sh[w]);
    return r;
}

__global__ void __launch_bounds__(1024) transpose_kernel(const cplx* __restrict__ in_rm, cplx* __restrict__ out_cm, int n) {
    __shared__ cplx tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    int r = by + threadIdx.y, c = bx + threadIdx.x;            // in_rm[r][c]
    if (r < n && c < n) tile[threadIdx.y][threadIdx.x] = in_rm[(long long)r * n + c];
    __syncthreads();
    int orow = by + threadIdx.x, ocol = bx + threadIdx.y;        // out_cm[orow + ocol*n] = in_rm[orow][ocol]
    if (orow < n && ocol < n) out_cm[orow + (long long)ocol * n] = tile[threadIdx.x][threadIdx.y];
}

__global__ void __launch_bounds__(RED_NT) rq_finish_kernel(const cplx* __restrict__ V, const cplx* __restrict__ Y, int n,
                                                           cplx* lambda, double* vnorm2, int* status) {
    __shared__ double sh[RED_NT / 32];
    const int c = blockIdx.x;
    const cplx* v = V + (long long)c * n;
    const cplx* y = Y + (long long)c * n;
    double nr = 0.0, ni = 0.0, d = 0.0;
    for (int i = threadIdx.x; i < n; i += RED_NT) {
        cplx a = v[i], b = y[i];
        nr = fma(a.x, b.x, nr); nr = fma(a.y, b.y, nr);     // conj(a) * b
        ni = fma(a.x, b.y, ni); ni = fma(-a.y, b.x, ni);
        d = fma(a.x, a.x, d); d = fma(a.y, a.y, d);
    }
    nr = block_sum(nr, sh); ni = block_sum(ni, sh); d = block_sum(d, sh);
    if (threadIdx.x == 0) {
        vnorm2[c] = d;
        lambda[c] = (fabs(d) < 1e-12) ? cmake(0.0, 0.0) : cmake(nr / d, ni / d);     // AMS:265-268
        if (status && sqrt(d) < 1e-10) status[c] = MAUS_ST_V_COLLAPSED;              // AMS:259
    }
}

__global__ void __launch_bounds__(RED_NT) mix_normalise_kernel(cplx* __restrict__ V, const cplx* __restrict__ X, int n,
                                                               int problem_type, const double* __restrict__ alpha,
                                                               double* mixnorm, int* status) {
    __shared__ double sh[RED_NT / 32];
    const int c = blockIdx.x;
    if (status[c] != 0) { if (threadIdx.x == 0 && mixnorm) mixnorm[c] = 0.0; return; }
    cplx* v = V + (long long)c * n;
    const cplx* x = X + (long long)c * n;
    const double a = alpha[c], oma = 1.0 - a;               // AMS:280 / 285 (alpha has zero imaginary part)
    double amax = 0.0;
    for (int i = threadIdx.x; i < n; i += RED_NT) {
        cplx vi = v[i], xi = x[i];
        cplx m = cmake(oma * vi.x + a * xi.x, oma * vi.y + a * xi.y);
        v[i] = m;
        amax = fmax(amax, fmax(fabs(m.x), fabs(m.y)));
    }
    amax = block_max(amax, sh);
    // scaled 2-norm (overflow-safe like BLAS dznrm2 behind np.linalg.norm, AMS:281)
    double ss = 0.0;
    if (amax > 0.0 && isfinite(amax)) {
        const double inv = 1.0 / amax;
        for (int i = threadIdx.x; i < n; i += RED_NT) {
            cplx m = v[i];
            double p = m.x * inv, q = m.y * inv;
            ss = fma(p, p, ss); ss = fma(q, q, ss);
        }
    }
    ss = block_sum(ss, sh);
    const double nv = (amax > 0.0 && isfinite(amax)) ? amax * sqrt(ss) : amax;
    if (threadIdx.x == 0 && mixnorm) mixnorm[c] = nv;
    if (problem_type == MAUS_EIGENVALUE) {
        if (nv > 1e-10) {                                    // AMS:282
            for (int i = threadIdx.x; i < n; i += RED_NT) { cplx m = v[i]; v[i] = cmake(m.x / nv, m.y / nv); }
        } else if (threadIdx.x == 0) {
            status[c] = MAUS_ST_MIX_COLLAPSED;               // AMS:283: host draws the replacement vector
        }
    }
}

__global__ void __launch_bounds__(RED_NT) residual_finish_kernel(const cplx* __restrict__ V, const cplx* __restrict__ Y,
                                                                 int n, int problem_type, const cplx* __restrict__ lambda,
                                                                 const cplx* __restrict__ b, double* resid) {
    __shared__ double sh[RED_NT / 32];
    const int c = blockIdx.x;
    const cplx* v = V + (long long)c * n;
    const cplx* y = Y + (long long)c * n;
    const cplx lam = (problem_type == MAUS_EIGENVALUE) ? lambda[c] : cmake(0.0, 0.0);
    double amax = 0.0;
    for (int i = threadIdx.x; i < n; i += RED_NT) {
        cplx r = y[i];
        if (problem_type == MAUS_EIGENVALUE) cfms(r, lam, v[i]); else r = csub(r, b[i]);
        amax = fmax(amax, fmax(fabs(r.x), fabs(r.y)));
    }
    amax = block_max(amax, sh);
    double ss = 0.0;
    const bool ok = amax > 0.0 && isfinite(amax);
    if (ok) {
        const double inv = 1.0 / amax;
        for (int i = threadIdx.x; i < n; i += RED_NT) {
            cplx r = y[i];
            if (problem_type == MAUS_EIGENVALUE) cfms(r, lam, v[i]); else r = csub(r, b[i]);
            double p = r.x * inv, q = r.y * inv;
            ss = fma(p, p, ss); ss = fma(q, q, ss);
        }
    }
    ss = block_sum(ss, sh);
    if (threadIdx.x == 0) {
        double out = ok ? amax * sqrt(ss) : amax;
        // a NaN anywhere must surface as NaN like np.linalg.norm does (fmax drops NaNs)
        resid[c] = out;
    }
}
// NaN propagation helper: fmax ignores NaN, so detect them explicitly
__global__ void __launch_bounds__(RED_NT) nan_scan_kernel(const cplx* __restrict__ Y, const cplx* __restrict__ V, int n,
                                                          double* resid) {
    __shared__ int bad;
    const int c = blockIdx.x;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    int mybad = 0;
    for (int i = threadIdx.x; i < n; i += RED_NT) {
        cplx a = Y[(long long)c * n + i], b = V[(long long)c * n + i];
        if (a.x != a.x || a.y != a.y || b.x != b.x || b.y != b.y) mybad = 1;
    }
    if (mybad) bad = 1;
    __syncthreads();
    if (threadIdx.x == 0 && bad) resid[c] = nan("");
}

// ---- HBM-bound batched matvec: Y[c] = A_rm * V[c] -------------------------------------------------------------
// One warp per matrix row (coalesced 512 B row segments, warp-shuffle dot-product reduction); CB candidate vectors
// are staged through shared memory in chunks so the matrix is streamed from HBM exactly once per CB candidates.
constexpr int GV_JC = 512;
// CB candidates per pass, RPW rows per warp (the staged vector entries are reused across rows: half the LDS traffic at
// RPW = 2), GV_NT threads (128-thread CTAs keep the grid at several waves when RPW = 2 is used below n = 8192)
template <int CB, int RPW, int GV_NT>
__global__ void __launch_bounds__(GV_NT) gemv_rowmajor_kernel(const cplx* __restrict__ A, const cplx* __restrict__ V,
                                                              long long ldv, cplx* __restrict__ Y, long long ldy, int nrows,
                                                              int n, int c0, int ncand) {
    // A is nrows x n row-major; Y[c] (nrows) = A * V[c] (n)
    __shared__ cplx sV[CB][GV_JC];
    constexpr int GV_ROWS = RPW * (GV_NT / 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * GV_ROWS + warp * RPW;
    cplx acc[RPW][CB];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int c = 0; c < CB; ++c) acc[r][c] = cmake(0.0, 0.0);
    for (int j0 = 0; j0 < n; j0 += GV_JC) {
        const int jc = min(GV_JC, n - j0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < CB * jc; idx += GV_NT) {
            int c = idx / jc, j = idx - c * jc;
            sV[c][j] = (c < ncand) ? V[(long long)(c0 + c) * ldv + j0 + j] : cmake(0.0, 0.0);
        }
        __syncthreads();
        const cplx* arow[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) arow[r] = A + (long long)min(row0 + r, nrows - 1) * n + j0;   // clamped: extra rows are not stored
        int j = lane;
        for (; j + 96 < jc; j += 128) {
            cplx a[RPW][4];
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                a[r][0] = __ldcs(&arow[r][j]); a[r][1] = __ldcs(&arow[r][j + 32]);
                a[r][2] = __ldcs(&arow[r][j + 64]); a[r][3] = __ldcs(&arow[r][j + 96]);
            }
#pragma unroll
            for (int c = 0; c < CB; ++c) {
                const cplx v0 = sV[c][j], v1 = sV[c][j + 32], v2 = sV[c][j + 64], v3 = sV[c][j + 96];
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    cfma(acc[r][c], a[r][0], v0); cfma(acc[r][c], a[r][1], v1);
                    cfma(acc[r][c], a[r][2], v2); cfma(acc[r][c], a[r][3], v3);
                }
            }
        }
        for (; j < jc; j += 32) {
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const cplx a0 = __ldcs(&arow[r][j]);
#pragma unroll
                for (int c = 0; c < CB; ++c) cfma(acc[r][c], a0, sV[c][j]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int row = row0 + r;
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            cplx s = warp_sum(acc[r][c]);
            if (lane == 0 && row < nrows && c < ncand) Y[(long long)(c0 + c) * ldy + row] = s;
        }
    }
}

__global__ void __launch_bounds__(256) diag_amax_kernel(const cplx* __restrict__ A, int n, cplx* __restrict__ diag, double* amax) {
    const long long total = (long long)n * n;
    double m = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        cplx a = A[i];
        double v = fabs(a.x) + fabs(a.y);
        if (v == v) m = fmax(m, v);
        if (i / n == i % n) diag[i / n] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0 && isfinite(m))
        atomicMax(reinterpret_cast<unsigned long long*>(amax), (unsigned long long)__double_as_longlong(m));   // m >= 0: bit order = value order
}

}  // namespace

cudaError_t vec_diag_amax(const cplx* A_rm, int n, cplx* diag, double* amax, cudaStream_t stream) {
    diag_amax_kernel<<<MAUS_SM_COUNT_B200 * 4, 256, 0, stream>>>(A_rm, n, diag, amax);
    return cudaGetLastError();
}

cudaError_t vec_rowmajor_to_colmajor(const cplx* in_rm, cplx* out_cm, int n, cudaStream_t stream) {
    dim3 grid((n + 31) / 32, (n + 31) / 32), block(32, 32);
    transpose_kernel<<<grid, block, 0, stream>>>(in_rm, out_cm, n);
    return cudaGetLastError();
}

cudaError_t vec_rq_finish(const cplx* V, const cplx* Y, int n, int C, cplx* lambda, double* vnorm2, int* status,
                          cudaStream_t stream) {
    if (C <= 0) return cudaSuccess;
    rq_finish_kernel<<<C, RED_NT, 0, stream>>>(V, Y, n, lambda, vnorm2, status);
    return cudaGetLastError();
}

cudaError_t vec_mix_normalise(cplx* V, const cplx* X, int n, int C, int problem_type, const double* alpha,
                              double* mixnorm, int* status, cudaStream_t stream) {
    if (C <= 0) return cudaSuccess;
    mix_normalise_kernel<<<C, RED_NT, 0, stream>>>(V, X, n, problem_type, alpha, mixnorm, status);
    return cudaGetLastError();
}

cudaError_t vec_residual_finish(const cplx* V, const cplx* Y, int n, int C, int problem_type, const cplx* lambda,
                                const cplx* b, double* resid, cudaStream_t stream) {
    if (C <= 0) return cudaSuccess;
    residual_finish_kernel<<<C, RED_NT, 0, stream>>>(V, Y, n, problem_type, lambda, b, resid);
    nan_scan_kernel<<<C, RED_NT, 0, stream>>>(Y, V, n, resid);
    return cudaGetLastError();
}

template <int RPW, int GV_NT>
static void gemv_launch(const cplx* A_rm, const cplx* V, long long ldv, cplx* Y, long long ldy, int nrows, int n, int C,
                        cudaStream_t stream) {
    const int rows = RPW * (GV_NT / 32);
    const int grid = (nrows + rows - 1) / rows;
    for (int c0 = 0; c0 < C; c0 += 4) {
        const int nc = (C - c0 < 4) ? (C - c0) : 4;
        if (nc == 1) gemv_rowmajor_kernel<1, RPW, GV_NT><<<grid, GV_NT, 0, stream>>>(A_rm, V, ldv, Y, ldy, nrows, n, c0, nc);
        else if (nc == 2) gemv_rowmajor_kernel<2, RPW, GV_NT><<<grid, GV_NT, 0, stream>>>(A_rm, V, ldv, Y, ldy, nrows, n, c0, nc);
        else gemv_rowmajor_kernel<4, RPW, GV_NT><<<grid, GV_NT, 0, stream>>>(A_rm, V, ldv, Y, ldy, nrows, n, c0, nc);
    }
}

cudaError_t vec_gemv_rowmajor(const cplx* A_rm, const cplx* V, long long ldv, cplx* Y, long long ldy, int n, int C,
                              cudaStream_t stream) {
    // grid sized so that several waves of CTAs cover the 148 SMs: 8 rows per CTA below n = 8192, 16 above
    if (n >= 8192) gemv_launch<2, 256>(A_rm, V, ldv, Y, ldy, n, n, C, stream);
    else if (C >= 2 && n >= 2048) gemv_launch<2, 128>(A_rm, V, ldv, Y, ldy, n, n, C, stream);   // several candidates: LDS-bound at RPW = 1
    else gemv_launch<1, 256>(A_rm, V, ldv, Y, ldy, n, n, C, stream);
    return cudaGetLastError();
}

cudaError_t vec_gemv_rect(const cplx* A_rm, int nrows, int ncols, const cplx* V, long long ldv, cplx* Y, long long ldy, int C,
                          cudaStream_t stream) {
    if (nrows >= 8192) gemv_launch<2, 256>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream);
    else if (C >= 2 && nrows >= 2048) gemv_launch<2, 128>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream);
    else gemv_launch<1, 256>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream);
    return cudaGetLastError();
}
