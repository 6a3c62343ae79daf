// vec.cu -- HBM-bound kernels of the candidate step: batched matvec, Rayleigh quotient, mix + normalise, residual.
#include <cstdlib>
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

// ---- multi-block variants for long vectors (n >= VMB_MIN_N: the sparse configurations, n ~ 1e6) ------------------------------
// One CTA per candidate cannot stream a 16 MB vector at HBM speed.  Here a candidate's vector is split over up to VMB_MAXBLK
// CTAs; every pass leaves per-block partials in `scratch` ([C][VMB_MAXBLK][4] doubles) and the consumer re-reduces them in a
// FIXED order, so the results do not depend on scheduling.  Same formulas (scaled 2-norms, thresholds) as the one-CTA kernels.
constexpr int VMB_MAXBLK = 64, VMB_MIN_N = 32768;
__device__ __forceinline__ void vmb_range(int n, int nblk, int blk, int& i0, int& i1) {
    const int per = ((n + nblk - 1) / nblk + 1) & ~1;
    i0 = min(n, blk * per);
    i1 = min(n, i0 + per);
}
__device__ __forceinline__ double vmb_sum(const double* part, int nblk, int k) {
    double s = 0.0;
    for (int q = 0; q < nblk; ++q) s += __ldcg(&part[q * 4 + k]);      // L2: written by other blocks (of this launch, when fused)
    return s;
}

// The 2-norms are accumulated UNSCALED (six FMAs per element, nothing else on the FP64 pipe); the scaled (dznrm2-like,
// overflow-safe) sum of the one-CTA kernels -- with its maximum and NaN scan -- is only recomputed when the plain sum left the
// safe range (never for normalised vectors).
__device__ __forceinline__ bool vmb_ss_safe(double ss) { return isfinite(ss) && ss > 1e-290; }

// block-done counters live behind the partials: scratch[C * VMB_MAXBLK * 4 + c] (zero-initialised, reset by the last block)
__device__ __forceinline__ unsigned int* vmb_counter(double* scratch, int C, int c) {
    return reinterpret_cast<unsigned int*>(scratch + (long long)C * VMB_MAXBLK * 4 + c);
}
// true in every thread of the LAST block of candidate c to finish (the partials of all blocks are then visible to it)
__device__ __forceinline__ bool vmb_last_block(double* scratch, int C, int c) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int* cnt = vmb_counter(scratch, C, c);
        const bool last = atomicAdd(cnt, 1u) == gridDim.x - 1;
        if (last) *cnt = 0;
        s_last = last ? 1 : 0;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}
__device__ __forceinline__ void rq_final_body(const double* part, int nblk, int c, cplx* lambda, double* vnorm2, int* status) {
    const double nr = vmb_sum(part, nblk, 0), ni = vmb_sum(part, nblk, 1), d = vmb_sum(part, nblk, 2);
    vnorm2[c] = d;
    lambda[c] = (fabs(d) < 1e-12) ? cmake(0.0, 0.0) : cmake(nr / d, ni / d);     // AMS:265-268
    if (status && sqrt(d) < 1e-10) status[c] = MAUS_ST_V_COLLAPSED;              // AMS:259
}

// fuse_C > 0: the last block of a candidate to finish also does the final reduction (fixed order over the blocks, whichever
// block it is): one launch instead of part + final
__global__ void __launch_bounds__(RED_NT) rq_part_kernel(const cplx* __restrict__ V, const cplx* __restrict__ Y, int n, double* scratch,
                                                         int fuse_C, cplx* lambda, double* vnorm2, int* status) {
    __shared__ double sh[RED_NT / 32];
    const int c = blockIdx.y;
    int i0, i1; vmb_range(n, gridDim.x, blockIdx.x, i0, i1);
    const cplx* v = V + (long long)c * n;
    const cplx* y = Y + (long long)c * n;
    double nr = 0.0, ni = 0.0, d = 0.0;
    for (int i = i0 + threadIdx.x; i < i1; i += 4 * RED_NT) {
        cplx a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = i + u * RED_NT;
            a[u] = (k < i1) ? __ldcs(&v[k]) : cmake(0.0, 0.0);
            b[u] = (k < i1) ? __ldcs(&y[k]) : cmake(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            nr = fma(a[u].x, b[u].x, nr); nr = fma(a[u].y, b[u].y, nr);     // conj(a) * b
            ni = fma(a[u].x, b[u].y, ni); ni = fma(-a[u].y, b[u].x, ni);
            d = fma(a[u].x, a[u].x, d); d = fma(a[u].y, a[u].y, d);
        }
    }
    nr = block_sum(nr, sh); ni = block_sum(ni, sh); d = block_sum(d, sh);
    if (threadIdx.x == 0) {
        double* o = scratch + ((long long)c * VMB_MAXBLK + blockIdx.x) * 4;
        o[0] = nr; o[1] = ni; o[2] = d;
    }
    if (fuse_C > 0 && vmb_last_block(scratch, fuse_C, c) && threadIdx.x == 0)
        rq_final_body(scratch + (long long)c * VMB_MAXBLK * 4, gridDim.x, c, lambda, vnorm2, status);
}
__global__ void rq_final_kernel(const double* __restrict__ scratch, int nblk, int C, cplx* lambda, double* vnorm2, int* status) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    rq_final_body(scratch + (long long)c * VMB_MAXBLK * 4, nblk, c, lambda, vnorm2, status);
}

// pass 1: v <- (1-a) v + a x; per block: max |component| -> part[0], plain sum of squares -> part[1]
__global__ void __launch_bounds__(RED_NT) mix_part_kernel(cplx* __restrict__ V, const cplx* __restrict__ X, int n,
                                                          const double* __restrict__ alpha, const int* __restrict__ status,
                                                          double* scratch) {
    __shared__ double sh[RED_NT / 32];
    const int c = blockIdx.y;
    if (status[c] != 0) return;
    int i0, i1; vmb_range(n, gridDim.x, blockIdx.x, i0, i1);
    cplx* v = V + (long long)c * n;
    const cplx* x = X + (long long)c * n;
    const double a = alpha[c], oma = 1.0 - a;
    double ss = 0.0;
    for (int i = i0 + threadIdx.x; i < i1; i += 4 * RED_NT) {
        cplx vi[4], xi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = i + u * RED_NT;
            vi[u] = (k < i1) ? v[k] : cmake(0.0, 0.0);
            xi[u] = (k < i1) ? __ldcs(&x[k]) : cmake(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = i + u * RED_NT;
            const cplx m = cmake(oma * vi[u].x + a * xi[u].x, oma * vi[u].y + a * xi[u].y);
            if (k < i1) v[k] = m;
            ss = fma(m.x, m.x, ss); ss = fma(m.y, m.y, ss);
        }
    }
    ss = block_sum(ss, sh);
    if (threadIdx.x == 0) {
        double* o = scratch + ((long long)c * VMB_MAXBLK + blockIdx.x) * 4;
        o[0] = 0.0; o[1] = ss;       // o[0] >= 0: "not skipped" (mix_preset); the maximum is found by the rare slow path itself
    }
}
// scaled 2-norm of a whole vector by ONE block (rare fallback when the plain sum of squares over- / underflowed)
__device__ double vmb_scaled_norm_slow(const cplx* __restrict__ v, int n, double* sh) {
    double amax = 0.0;
    bool nan_seen = false;
    for (int i = threadIdx.x; i < n; i += RED_NT) {
        const cplx m = v[i];
        if (m.x != m.x || m.y != m.y) nan_seen = true;
        amax = fmax(amax, fmax(fabs(m.x), fabs(m.y)));
    }
    amax = block_max(amax, sh);
    const double anynan = block_max(nan_seen ? 1.0 : 0.0, sh);
    if (anynan > 0.0) return nan("");
    if (!(amax > 0.0) || !isfinite(amax)) return amax;
    const double inv = 1.0 / amax;
    double ss = 0.0;
    for (int i = threadIdx.x; i < n; i += RED_NT) {
        const cplx m = v[i];
        const double p = m.x * inv, q = m.y * inv;
        ss = fma(p, p, ss); ss = fma(q, q, ss);
    }
    ss = block_sum(ss, sh);
    return amax * sqrt(ss);
}
// pass 2: norm from the partials, normalisation (eigen) / collapse status.  Skipped candidates are recognised by the
// preset part[0] = -1 (status[c] itself is updated by block 0 while other blocks of this pass may still be starting).
__global__ void __launch_bounds__(RED_NT) mix_apply_kernel(cplx* __restrict__ V, int n, int problem_type, double* mixnorm,
                                                           int* status, const double* __restrict__ scratch) {
    __shared__ double sh[RED_NT / 32];
    const int c = blockIdx.y;
    const double* part = scratch + (long long)c * VMB_MAXBLK * 4;
    if (part[0] < 0.0) { if (blockIdx.x == 0 && threadIdx.x == 0 && mixnorm) mixnorm[c] = 0.0; return; }
    int i0, i1; vmb_range(n, gridDim.x, blockIdx.x, i0, i1);
    cplx* v = V + (long long)c * n;
    // plain sum of squares; outside its safe range (zero vector, over- / underflow, inf / nan) every block recomputes the same
    // scaled (dznrm2-like) value from the whole vector
    const double ss = vmb_sum(part, gridDim.x, 1);
    const double nv = vmb_ss_safe(ss) ? sqrt(ss) : ((ss != ss) ? ss : vmb_scaled_norm_slow(v, n, sh));
    if (blockIdx.x == 0 && threadIdx.x == 0 && mixnorm) mixnorm[c] = nv;
    if (problem_type == MAUS_EIGENVALUE) {
        if (nv > 1e-10) {                                    // AMS:282
            for (int i = i0 + threadIdx.x; i < i1; i += RED_NT) { cplx m = v[i]; v[i] = cmake(m.x / nv, m.y / nv); }
        } else if (blockIdx.x == 0 && threadIdx.x == 0) {
            status[c] = MAUS_ST_MIX_COLLAPSED;               // AMS:283
        }
    }
}
__global__ void mix_preset_kernel(const int* __restrict__ status, int C, double* scratch) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) scratch[(long long)c * VMB_MAXBLK * 4] = (status[c] != 0) ? -1.0 : 0.0;
}

// residual, one pass: r = y - lambda v (eigen) / y - b (linear); max |component|, plain sum of squares, NaN flag of y and v
// final value of candidate c from the partials of nblk blocks; executed by a whole block (the rare scaled recomputation walks the vector)
__device__ void res_final_body(const cplx* __restrict__ V, const cplx* __restrict__ Y, int n, int problem_type, const cplx* __restrict__ lambda,
                               const cplx* __restrict__ b, const double* part, int nblk, int c, double* resid, double* sh) {
    const double ss = vmb_sum(part, nblk, 1);
    double out;
    if (vmb_ss_safe(ss)) out = sqrt(ss);
    else if (ss != ss) out = ss;          // NaN in y or v (np.linalg.norm propagates it); decided on the combined sum, so that
                                          // the ranks of a row-sharded vector agree
    else {
        // rare (zero residual, over- / underflow, inf / nan): this block alone walks the vector: NaN test and maximum, then the
        // scaled sum.  np.linalg.norm propagates NaN (fmax would drop it).
        const cplx* v = V + (long long)c * n;
        const cplx* y = Y + (long long)c * n;
        const bool eig = problem_type == MAUS_EIGENVALUE;
        const cplx lam = eig ? lambda[c] : cmake(0.0, 0.0);
        double amax = 0.0, bad = 0.0;
        for (int i = threadIdx.x; i < n; i += RED_NT) {
            cplx r = y[i];
            const cplx w = v[i];
            if (r.x != r.x || r.y != r.y || w.x != w.x || w.y != w.y) bad = 1.0;
            if (eig) cfms(r, lam, w); else r = csub(r, b[i]);
            if (r.x != r.x || r.y != r.y) bad = 1.0;
            amax = fmax(amax, fmax(fabs(r.x), fabs(r.y)));
        }
        amax = block_max(amax, sh); bad = block_max(bad, sh);
        if (bad > 0.0) out = nan("");
        else if (!(amax > 0.0) || !isfinite(amax)) out = amax;
        else {
            const double inv = 1.0 / amax;
            double s2 = 0.0;
            for (int i = threadIdx.x; i < n; i += RED_NT) {
                cplx r = y[i];
                if (eig) cfms(r, lam, v[i]); else r = csub(r, b[i]);
                const double p = r.x * inv, q = r.y * inv;
                s2 = fma(p, p, s2); s2 = fma(q, q, s2);
            }
            s2 = block_sum(s2, sh);
            out = amax * sqrt(s2);
        }
    }
    if (threadIdx.x == 0) resid[c] = out;
}

__global__ void __launch_bounds__(RED_NT) res_part_kernel(const cplx* __restrict__ V, const cplx* __restrict__ Y, int n, int problem_type,
                                                          const cplx* __restrict__ lambda, const cplx* __restrict__ b, double* scratch,
                                                          int fuse_C, double* resid) {
    __shared__ double sh[RED_NT / 32];
    const int c = blockIdx.y;
    int i0, i1; vmb_range(n, gridDim.x, blockIdx.x, i0, i1);
    const cplx* v = V + (long long)c * n;
    const cplx* y = Y + (long long)c * n;
    const bool eig = problem_type == MAUS_EIGENVALUE;
    const cplx lam = eig ? lambda[c] : cmake(0.0, 0.0);
    // one pass, six FMAs per element: r = y - lambda v (eigen) / y - b (linear) and the plain sum of squares.  A NaN in y or v
    // reaches the sum through r, an overflow makes it inf, a zero residual leaves 0: all three send the final step to its slow path
    double ss = 0.0;
    for (int i = i0 + threadIdx.x; i < i1; i += 4 * RED_NT) {
        cplx r[4], w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = i + u * RED_NT;
            r[u] = (k < i1) ? __ldcs(&y[k]) : cmake(0.0, 0.0);
            w[u] = (k < i1) ? (eig ? v[k] : b[k]) : cmake(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (eig) cfms(r[u], lam, w[u]); else r[u] = csub(r[u], w[u]);
            ss = fma(r[u].x, r[u].x, ss); ss = fma(r[u].y, r[u].y, ss);
        }
    }
    ss = block_sum(ss, sh);
    if (threadIdx.x == 0) {
        double* o = scratch + ((long long)c * VMB_MAXBLK + blockIdx.x) * 4;
        o[1] = ss;
    }
    if (fuse_C > 0 && vmb_last_block(scratch, fuse_C, c))
        res_final_body(V, Y, n, problem_type, lambda, b, scratch + (long long)c * VMB_MAXBLK * 4, gridDim.x, c, resid, sh);
}
__global__ void __launch_bounds__(RED_NT) res_final_kernel(const cplx* __restrict__ V, const cplx* __restrict__ Y, int n, int problem_type,
                                                           const cplx* __restrict__ lambda, const cplx* __restrict__ b,
                                                           const double* __restrict__ scratch, int nblk, double* resid) {
    __shared__ double sh[RED_NT / 32];
    const int c = blockIdx.x;
    res_final_body(V, Y, n, problem_type, lambda, b, scratch + (long long)c * VMB_MAXBLK * 4, nblk, c, resid, sh);
}

__host__ inline int vmb_blocks(int n) {
    const int want = (n + RED_NT * 8 - 1) / (RED_NT * 8);      // >= 8 elements per thread
    return want < 1 ? 1 : (want > VMB_MAXBLK ? VMB_MAXBLK : want);
}
// Blocks per candidate when C candidates are reduced in one launch: these kernels run for 40 - 200 us, so a partly filled last
// wave of CTAs costs up to a third of the run time (8 candidates x 64 blocks = 512 CTAs = 3.46 per SM: the SMs that got four set
// the time).  Among the counts between half the cap and the cap, take the one whose CTA total fills whole waves best (a
// deterministic function of n and C, like the summation order it implies).
__host__ inline int vmb_blocks(int n, int C) {
    const int cap = vmb_blocks(n);
    if (cap < 8 || C <= 0) return cap;
    int best = cap; double best_eff = 0.0;
    for (int nb = cap; nb >= (cap + 1) / 2; --nb) {
        const long long ctas = (long long)nb * C;
        const long long waves = (ctas + MAUS_SM_COUNT_B200 - 1) / MAUS_SM_COUNT_B200;
        const double eff = (double)ctas / (double)(waves * MAUS_SM_COUNT_B200);
        if (eff >= 0.97) return nb;                        // the largest count that is balanced to 3 %
        if (eff > best_eff) { best_eff = eff; best = nb; }
    }
    return best;
}

// ---- Gram matrix of converged candidates (dedup tests |<v_i, v_j>| > 0.999, AMS:436, 450, 515, 520) ----------------------
// G[i][j] = sum_k conj(v_i[k]) v_j[k] (= np.vdot(v_i, v_j)).  One CTA per (i, group of GR_J columns j): v_i is read once per
// group, deterministic block reduction.  C is a few hundred at most and the vectors are L2 resident (C n 16 B).
constexpr int GR_J = 4, GR_NT = 256;
__global__ void __launch_bounds__(GR_NT) gram_kernel(const cplx* __restrict__ V, int n, int C, cplx* __restrict__ G) {
    __shared__ double sh[GR_NT / 32];
    const int i = blockIdx.x, j0 = blockIdx.y * GR_J;
    const cplx* vi = V + (long long)i * n;
    double ar[GR_J], ai[GR_J];
#pragma unroll
    for (int q = 0; q < GR_J; ++q) { ar[q] = 0.0; ai[q] = 0.0; }
    for (int k = threadIdx.x; k < n; k += GR_NT) {
        const cplx a = vi[k];
#pragma unroll
        for (int q = 0; q < GR_J; ++q) {
            const int j = min(j0 + q, C - 1);
            const cplx b = V[(long long)j * n + k];
            ar[q] = fma(a.x, b.x, ar[q]); ar[q] = fma(a.y, b.y, ar[q]);      // conj(a) * b
            ai[q] = fma(a.x, b.y, ai[q]); ai[q] = fma(-a.y, b.x, ai[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < GR_J; ++q) {
        const double sr = block_sum(ar[q], sh), si = block_sum(ai[q], sh);
        if (threadIdx.x == 0 && j0 + q < C) G[(long long)i * C + j0 + q] = cmake(sr, si);
    }
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

cudaError_t vec_gram(const cplx* V, int n, int C, cplx* G, cudaStream_t stream) {
    if (C <= 0) return cudaSuccess;
    gram_kernel<<<dim3(C, (C + GR_J - 1) / GR_J), GR_NT, 0, stream>>>(V, n, C, G);
    return cudaGetLastError();
}

// per-block partials [C][VMB_MAXBLK][4] + one block-done counter per candidate (must be zero-initialised once)
size_t vec_scratch_doubles(long long C) { return (size_t)C * VMB_MAXBLK * 4 + (size_t)C; }

cudaError_t vec_rq_finish(const cplx* V, const cplx* Y, int n, int C, cplx* lambda, double* vnorm2, int* status,
                          double* scratch, cudaStream_t stream, int scratch_cap) {
    if (C <= 0) return cudaSuccess;
    if (scratch && n >= VMB_MIN_N) {
        const int nblk = vmb_blocks(n, C);
        rq_part_kernel<<<dim3(nblk, C), RED_NT, 0, stream>>>(V, Y, n, scratch, scratch_cap > 0 ? scratch_cap : 0, lambda, vnorm2, status);
        if (scratch_cap <= 0) rq_final_kernel<<<(C + 127) / 128, 128, 0, stream>>>(scratch, nblk, C, lambda, vnorm2, status);
        return cudaGetLastError();
    }
    rq_finish_kernel<<<C, RED_NT, 0, stream>>>(V, Y, n, lambda, vnorm2, status);
    return cudaGetLastError();
}

cudaError_t vec_mix_normalise(cplx* V, const cplx* X, int n, int C, int problem_type, const double* alpha,
                              double* mixnorm, int* status, double* scratch, cudaStream_t stream) {
    if (C <= 0) return cudaSuccess;
    if (scratch && n >= VMB_MIN_N) {
        const int nblk = vmb_blocks(n, C);
        mix_preset_kernel<<<(C + 127) / 128, 128, 0, stream>>>(status, C, scratch);
        mix_part_kernel<<<dim3(nblk, C), RED_NT, 0, stream>>>(V, X, n, alpha, status, scratch);
        mix_apply_kernel<<<dim3(nblk, C), RED_NT, 0, stream>>>(V, n, problem_type, mixnorm, status, scratch);
        return cudaGetLastError();
    }
    mix_normalise_kernel<<<C, RED_NT, 0, stream>>>(V, X, n, problem_type, alpha, mixnorm, status);
    return cudaGetLastError();
}

cudaError_t vec_residual_finish(const cplx* V, const cplx* Y, int n, int C, int problem_type, const cplx* lambda,
                                const cplx* b, double* resid, double* scratch, cudaStream_t stream, int scratch_cap) {
    if (C <= 0) return cudaSuccess;
    if (scratch && n >= VMB_MIN_N) {
        const int nblk = vmb_blocks(n, C);
        res_part_kernel<<<dim3(nblk, C), RED_NT, 0, stream>>>(V, Y, n, problem_type, lambda, b, scratch, scratch_cap > 0 ? scratch_cap : 0, resid);
        if (scratch_cap <= 0) res_final_kernel<<<C, RED_NT, 0, stream>>>(V, Y, n, problem_type, lambda, b, scratch, nblk, resid);
        return cudaGetLastError();
    }
    residual_finish_kernel<<<C, RED_NT, 0, stream>>>(V, Y, n, problem_type, lambda, b, resid);
    nan_scan_kernel<<<C, RED_NT, 0, stream>>>(Y, V, n, resid);
    return cudaGetLastError();
}

// ---- explicit halves of the multi-block reductions (row-sharded vectors, rowshard.cu): every rank runs the `part` kernel on
// its slice, the per-block partials in `scratch` are combined over the ranks, then the `final` / `apply` kernel runs ----
int vec_part_blocks(long long n) { return vmb_blocks((int)n); }
cudaError_t vec_rq_part(const cplx* V, const cplx* Y, int n, int C, double* scratch, int nblk, cudaStream_t stream) {
    rq_part_kernel<<<dim3(nblk, C), RED_NT, 0, stream>>>(V, Y, n, scratch, 0, nullptr, nullptr, nullptr);
    return cudaGetLastError();
}
cudaError_t vec_rq_final(const double* scratch, int nblk, int C, cplx* lambda, double* vnorm2, int* status, cudaStream_t stream) {
    rq_final_kernel<<<(C + 127) / 128, 128, 0, stream>>>(scratch, nblk, C, lambda, vnorm2, status);
    return cudaGetLastError();
}
cudaError_t vec_mix_part(cplx* V, const cplx* X, int n, int C, const double* alpha, const int* status, double* scratch, int nblk,
                         cudaStream_t stream) {
    mix_preset_kernel<<<(C + 127) / 128, 128, 0, stream>>>(status, C, scratch);
    mix_part_kernel<<<dim3(nblk, C), RED_NT, 0, stream>>>(V, X, n, alpha, status, scratch);
    return cudaGetLastError();
}
cudaError_t vec_mix_apply(cplx* V, int n, int C, int problem_type, double* mixnorm, int* status, const double* scratch, int nblk,
                          cudaStream_t stream) {
    mix_apply_kernel<<<dim3(nblk, C), RED_NT, 0, stream>>>(V, n, problem_type, mixnorm, status, scratch);
    return cudaGetLastError();
}
cudaError_t vec_res_part(const cplx* V, const cplx* Y, int n, int C, int problem_type, const cplx* lambda, const cplx* b,
                         double* scratch, int nblk, cudaStream_t stream) {
    res_part_kernel<<<dim3(nblk, C), RED_NT, 0, stream>>>(V, Y, n, problem_type, lambda, b, scratch, 0, nullptr);
    return cudaGetLastError();
}
cudaError_t vec_res_final(const cplx* V, const cplx* Y, int n, int C, int problem_type, const cplx* lambda, const cplx* b,
                          const double* scratch, int nblk, double* resid, cudaStream_t stream) {
    res_final_kernel<<<C, RED_NT, 0, stream>>>(V, Y, n, problem_type, lambda, b, scratch, nblk, resid);
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

// Rows per CTA = RPW * warps.  At n = 4096 a fixed 8 rows per CTA gives 512 CTAs = 3.46 per SM: the SMs that got four set the
// time (0.865 of a balanced grid).  The warp count (4 .. 8) is therefore chosen so that the CTAs fill whole waves of the 148 SMs
// as well as possible (n = 4096, two rows per warp: 7 warps -> 293 CTAs = 1.98 per SM).
template <int RPW>
static void gemv_dispatch(const cplx* A_rm, const cplx* V, long long ldv, cplx* Y, long long ldy, int nrows, int ncols, int C,
                          cudaStream_t stream) {
    int best_w = 8; double best_eff = -1.0;
    for (int w = 8; w >= 4; --w) {
        const long long ctas = (nrows + RPW * w - 1) / (RPW * w);
        const long long per_sm = (ctas + MAUS_SM_COUNT_B200 - 1) / MAUS_SM_COUNT_B200;
        const double eff = (double)ctas / (double)(per_sm * MAUS_SM_COUNT_B200);
        if (eff > best_eff + 0.02) { best_eff = eff; best_w = w; }      // prefer the larger CTA unless a smaller one is clearly better
    }
    switch (best_w) {
        case 4: gemv_launch<RPW, 128>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream); break;
        case 5: gemv_launch<RPW, 160>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream); break;
        case 6: gemv_launch<RPW, 192>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream); break;
        case 7: gemv_launch<RPW, 224>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream); break;
        default: gemv_launch<RPW, 256>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream); break;
    }
}

cudaError_t vec_gemv_rect(const cplx* A_rm, int nrows, int ncols, const cplx* V, long long ldv, cplx* Y, long long ldy, int C,
                          cudaStream_t stream) {
    // two rows per warp from 2048 rows with several candidates (LDS-bound at one row per warp) and from 8192 rows always
    // (four rows per warp were measured in round 2: 0.32 / 0.47 of the HBM peak at n = 4096 against 0.61 -- register pressure
    // costs more occupancy than the halved shared-memory traffic returns).  One candidate: wave-balanced CTA size (n = 8192:
    // 0.91 -> 0.97 of the HBM peak); several candidates keep the small CTAs -- every CTA re-stages the vectors chunk by chunk,
    // and few large CTAs per SM expose those bubbles (balanced sizes measured 0.61 -> 0.58 at n = 4096, 0.76 -> 0.68 at 8192)
    if (C == 1) {
        if (nrows >= 8192) gemv_dispatch<2>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream);
        else gemv_dispatch<1>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream);
    } else if (nrows >= 8192) gemv_launch<2, 256>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream);
    else if (nrows >= 2048) gemv_launch<2, 128>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream);
    else gemv_launch<1, 256>(A_rm, V, ldv, Y, ldy, nrows, ncols, C, stream);
    return cudaGetLastError();
}

cudaError_t vec_gemv_rowmajor(const cplx* A_rm, const cplx* V, long long ldv, cplx* Y, long long ldy, int n, int C,
                              cudaStream_t stream) {
    return vec_gemv_rect(A_rm, n, n, V, ldv, Y, ldy, C, stream);
}
