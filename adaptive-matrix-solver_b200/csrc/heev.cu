// heev.cu -- dense Hermitian eigensolver on the device: the factorisation behind the reference's Hermitian shortcut
// (AMS:155-186 calls scipy.linalg.eigh -> LAPACK zheevd once per candidate per step).
//
// Algorithm: cyclic two-sided Jacobi with the round-robin ("tournament") parallel ordering.  One step applies n / 2 disjoint
// complex plane rotations J = prod J_(p,q) at once:  A <- J^H A J,  V <- V J.  With disjoint pairs the new value of every
// element depends only on the 2 x 2 block A[{p,q}, {r,s}] spanned by the pair of its row and the pair of its column, so ONE
// kernel updates the whole matrix in place, one thread per block, no intermediate copy, and with the round-robin order the
// four streams a warp touches (rows p / q, columns r ascending / s descending) are contiguous.  n - 1 steps visit every index
// pair once (one sweep); sweeps repeat until the off-diagonal Frobenius norm is below a few ulps of ||A||_F (quadratic
// convergence, 6 - 10 sweeps).  Jacobi is chosen over tridiagonalisation + divide and conquer because every step is a
// perfectly regular, HBM / L2-streaming pass (B200: 6.5 TB/s) with no panel factorisations or host round trips, and because
// it delivers the eigenvectors to full relative accuracy.  Work per sweep: 64 n^3 B of traffic (A and V read + written once
// per step).
#include <vector>
#include <algorithm>
#include <cfloat>
#include "ctx.cuh"

namespace {

struct HeevWs {
    long long n = 0;
    cplx *A = nullptr, *V = nullptr, *out = nullptr;
    double4* rot = nullptr;          // per pair: (c, s.re, s.im, unused)
    double* red = nullptr;           // [2]: off-diagonal and total sums of squares
    int* perm = nullptr;
    size_t bytes = 0;
};

// pair k of step t in the round-robin ordering of m (even) players: player m - 1 stays, the others rotate
__device__ __forceinline__ void hv_pair(int m, int t, int k, int& p, int& q) {
    if (k == 0) { p = m - 1; q = t; }
    else { p = (t + k) % (m - 1); q = (t - k + (m - 1)) % (m - 1); }
}

// work copy from the LOWER triangle (LAPACK zheevd's default, scipy eigh(lower=True)); imaginary parts of the diagonal dropped
__global__ void hv_init_kernel(const cplx* __restrict__ in, cplx* __restrict__ A, cplx* __restrict__ V, long long n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const long long i = idx / n, j = idx % n;
    cplx a;
    if (i > j) a = in[idx];
    else if (i < j) { const cplx t = in[j * n + i]; a = cmake(t.x, -t.y); }
    else a = cmake(in[idx].x, 0.0);
    A[idx] = a;
    V[idx] = cmake(i == j ? 1.0 : 0.0, 0.0);
}

// rotation of every pair of step t from its 2 x 2 diagonal block [[alpha, beta], [conj(beta), gamma]]:
// J = [[c, s], [-conj(s), c]], c real, chosen so that (J^H A J)_pq = 0 (the smaller-angle root, |t| <= 1)
__global__ void hv_rot_kernel(const cplx* __restrict__ A, long long n, int m, int t, double4* __restrict__ rot) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m / 2) return;
    int p, q; hv_pair(m, t, k, p, q);
    double c = 1.0; cplx s = cmake(0.0, 0.0);
    if (p < n && q < n) {
        const double alpha = A[(long long)p * n + p].x, gamma = A[(long long)q * n + q].x;
        const cplx beta = A[(long long)p * n + q];
        const double ab = hypot(beta.x, beta.y);
        if (ab > 0.0 && ab > 1e-300 && ab >= 1.0e-19 * (fabs(alpha) + fabs(gamma))) {
            const double tau = (gamma - alpha) / (2.0 * ab);
            const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + tt * tt);
            const double sr = tt * c;
            s = cmake(sr * beta.x / ab, sr * beta.y / ab);
        }
    }
    rot[k] = make_double4(c, s.x, s.y, 0.0);
}

// A <- J^H A J: thread (k1, k2) owns the 2 x 2 block rows {p, q} of pair k1, columns {r, s} of pair k2
__global__ void __launch_bounds__(128) hv_apply_a_kernel(cplx* __restrict__ A, long long n, int m, int t, const double4* __restrict__ rot) {
    const int k2 = blockIdx.x * blockDim.x + threadIdx.x, k1 = blockIdx.y;
    if (k2 >= m / 2) return;
    int p, q, r, s; hv_pair(m, t, k1, p, q); hv_pair(m, t, k2, r, s);
    const double4 RP = rot[k1], RQ = rot[k2];
    const bool vp = p < n, vq = q < n, vr = r < n, vs = s < n;
    const cplx z = cmake(0.0, 0.0);
    cplx b00 = (vp && vr) ? A[(long long)p * n + r] : z, b01 = (vp && vs) ? A[(long long)p * n + s] : z;
    cplx b10 = (vq && vr) ? A[(long long)q * n + r] : z, b11 = (vq && vs) ? A[(long long)q * n + s] : z;
    const double cP = RP.x, cQ = RQ.x;
    const cplx sP = cmake(RP.y, RP.z), sQ = cmake(RQ.y, RQ.z), sPc = cmake(RP.y, -RP.z), sQc = cmake(RQ.y, -RQ.z);
    // T = J_P^H B : row 0 = c B0 - s B1 ; row 1 = conj(s) B0 + c B1
    cplx t00 = cscale(b00, cP), t01 = cscale(b01, cP), t10 = cscale(b10, cP), t11 = cscale(b11, cP);
    cfms(t00, sP, b10); cfms(t01, sP, b11); cfma(t10, sPc, b00); cfma(t11, sPc, b01);
    // B' = T J_Q : column 0 = c T0 - conj(s) T1 ; column 1 = s T0 + c T1
    cplx n00 = cscale(t00, cQ), n10 = cscale(t10, cQ), n01 = cscale(t01, cQ), n11 = cscale(t11, cQ);
    cfms(n00, sQc, t01); cfms(n10, sQc, t11); cfma(n01, sQ, t00); cfma(n11, sQ, t10);
    if (k1 == k2) { n01 = z; n10 = z; n00.y = 0.0; n11.y = 0.0; }      // the annihilated pair: exact zeros, real diagonal
    if (vp && vr) A[(long long)p * n + r] = n00;
    if (vp && vs) A[(long long)p * n + s] = n01;
    if (vq && vr) A[(long long)q * n + r] = n10;
    if (vq && vs) A[(long long)q * n + s] = n11;
}

// V <- V J: thread (i, k2) owns V[i][{r, s}]
__global__ void __launch_bounds__(128) hv_apply_v_kernel(cplx* __restrict__ V, long long n, int m, int t, const double4* __restrict__ rot) {
    const int k2 = blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = blockIdx.y;
    if (k2 >= m / 2) return;
    int r, s; hv_pair(m, t, k2, r, s);
    if (r >= n || s >= n) return;                                       // a pair with the padding index: identity
    const double4 RQ = rot[k2];
    const double cQ = RQ.x;
    const cplx sQ = cmake(RQ.y, RQ.z), sQc = cmake(RQ.y, -RQ.z);
    const cplx v0 = V[i * n + r], v1 = V[i * n + s];
    cplx n0 = cscale(v0, cQ), n1 = cscale(v1, cQ);
    cfms(n0, sQc, v1); cfma(n1, sQ, v0);
    V[i * n + r] = n0; V[i * n + s] = n1;
}

// red[0] = sum_{i != j} |a_ij|^2, red[1] = sum |a_ij|^2 (must be zeroed before)
__global__ void __launch_bounds__(256) hv_offnorm_kernel(const cplx* __restrict__ A, long long n, double* red) {
    double off = 0.0, tot = 0.0;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += (long long)gridDim.x * blockDim.x) {
        const cplx a = A[idx];
        const double v = a.x * a.x + a.y * a.y;
        tot += v;
        if (idx / n != idx % n) off += v;
    }
    off = warp_sum(off); tot = warp_sum(tot);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&red[0], off); atomicAdd(&red[1], tot); }
}

__global__ void hv_diag_kernel(const cplx* __restrict__ A, long long n, double* __restrict__ w) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = A[i * n + i].x;
}

// out[i][j] = V[i][perm[j]] (eigenvectors as columns, ascending eigenvalues, row-major like numpy)
__global__ void hv_gather_kernel(const cplx* __restrict__ V, const int* __restrict__ perm, cplx* __restrict__ out, long long n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const long long i = idx / n, j = idx % n;
    out[idx] = V[i * n + perm[j]];
}

}  // namespace

void maus_heev_free(maus_ctx* ctx) {
    HeevWs* ws = (HeevWs*)ctx->heev;
    if (!ws) return;
    cudaFree(ws->A); cudaFree(ws->V); cudaFree(ws->out); cudaFree(ws->rot); cudaFree(ws->red); cudaFree(ws->perm);
    ctx->bytes_held -= (long long)ws->bytes;
    delete ws;
    ctx->heev = nullptr;
}

extern "C" int maus_heev(maus_ctx* ctx, int64_t n, const double* A_rowmajor, int max_sweeps, double* w_out, double* E_rowmajor_out,
                         int32_t* sweeps_out, double* off_ratio_out) {
    if (!ctx || !A_rowmajor || !w_out || n <= 0 || n > 46340) return maus_fail(ctx, MAUS_E_ARG, "maus_heev: bad argument");
    cudaSetDevice(ctx->device);
    if (max_sweeps <= 0) max_sweeps = 30;
    cudaStream_t st = ctx->stream;
    HeevWs* ws = (HeevWs*)ctx->heev;
    if (!ws || ws->n != n) {
        MAUS_CUDA(ctx, cudaStreamSynchronize(st));
        maus_heev_free(ctx);
        ws = new HeevWs();
        ws->n = n;
        ctx->heev = ws;
        const size_t mat = (size_t)n * n * sizeof(cplx);
        cudaError_t e = cudaMalloc(&ws->A, mat);
        if (e == cudaSuccess) e = cudaMalloc(&ws->V, mat);
        if (e == cudaSuccess) e = cudaMalloc(&ws->out, mat);
        if (e == cudaSuccess) e = cudaMalloc(&ws->rot, (size_t)(n / 2 + 1) * sizeof(double4));
        if (e == cudaSuccess) e = cudaMalloc(&ws->red, 2 * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&ws->perm, (size_t)n * sizeof(int));
        if (e != cudaSuccess) { maus_heev_free(ctx); return maus_fail(ctx, MAUS_E_NOMEM, "maus_heev: workspace", e); }
        ws->bytes = 3 * mat + (size_t)(n / 2 + 1) * sizeof(double4) + 16 + (size_t)n * 4;
        ctx->bytes_held += (long long)ws->bytes;
    }
    MausNvtxRange range("maus.heev");
    const size_t mat = (size_t)n * n * sizeof(cplx);
    const unsigned gall = (unsigned)(((long long)n * n + 255) / 256);
    MAUS_CUDA(ctx, cudaMemcpyAsync(ws->out, A_rowmajor, mat, cudaMemcpyHostToDevice, st));
    hv_init_kernel<<<gall, 256, 0, st>>>(ws->out, ws->A, ws->V, n);
    const int m = (int)((n + 1) & ~1LL);                  // even number of players; index n (if any) is padding
    const int half = m / 2;
    const dim3 ga((unsigned)((half + 127) / 128), (unsigned)half), gv((unsigned)((half + 127) / 128), (unsigned)n);
    double red[2] = {0.0, 0.0};
    int sweeps = 0;
    double ratio = 0.0, prev_ratio = 1.0e300;
    for (; sweeps < max_sweeps; ) {
        MAUS_CUDA(ctx, cudaMemsetAsync(ws->red, 0, 2 * sizeof(double), st));
        hv_offnorm_kernel<<<MAUS_SM_COUNT_B200 * 4, 256, 0, st>>>(ws->A, n, ws->red);
        MAUS_CUDA(ctx, cudaMemcpyAsync(red, ws->red, sizeof red, cudaMemcpyDeviceToHost, st));
        MAUS_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->launches += 1;
        if (!(red[1] > 0.0) || !std::isfinite(red[1])) { ratio = red[1] > 0.0 ? NAN : 0.0; break; }      // zero matrix / non-finite input
        ratio = std::sqrt(red[0] / red[1]);
        // converged: off-diagonal mass at the rounding floor -- a few ulps of ||A||_F, or (large n: the floor grows like
        // sqrt(n) eps) no longer shrinking once it is far below the accuracy asked of the eigenpairs
        if (ratio <= 4.0 * DBL_EPSILON || (ratio <= 1.0e-13 && ratio > 0.5 * prev_ratio)) break;
        prev_ratio = ratio;
        if (n > 1)
            for (int t = 0; t < m - 1; ++t) {
                hv_rot_kernel<<<(unsigned)((half + 127) / 128), 128, 0, st>>>(ws->A, n, m, t, ws->rot);
                hv_apply_a_kernel<<<ga, 128, 0, st>>>(ws->A, n, m, t, ws->rot);
                hv_apply_v_kernel<<<gv, 128, 0, st>>>(ws->V, n, m, t, ws->rot);
            }
        ctx->launches += 3LL * (m - 1);
        ++sweeps;
    }
    MAUS_CUDA(ctx, cudaGetLastError());
    // eigenvalues = the diagonal, ascending like eigh; eigenvector j = column perm[j] of V
    std::vector<double> w((size_t)n);
    double* dw = reinterpret_cast<double*>(ws->rot);      // reuse: n doubles fit in (n / 2 + 1) double4
    hv_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws->A, n, dw);
    MAUS_CUDA(ctx, cudaMemcpyAsync(w.data(), dw, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<int> perm((size_t)n);
    for (long long i = 0; i < n; ++i) perm[(size_t)i] = (int)i;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return w[(size_t)a] < w[(size_t)b]; });
    for (long long i = 0; i < n; ++i) w_out[i] = w[(size_t)perm[(size_t)i]];
    if (E_rowmajor_out) {
        MAUS_CUDA(ctx, cudaMemcpyAsync(ws->perm, perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
        hv_gather_kernel<<<gall, 256, 0, st>>>(ws->V, ws->perm, ws->out, n);
        MAUS_CUDA(ctx, cudaMemcpyAsync(E_rowmajor_out, ws->out, mat, cudaMemcpyDeviceToHost, st));
        MAUS_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->launches += 1;
    }
    if (sweeps_out) *sweeps_out = sweeps;
    if (off_ratio_out) *off_ratio_out = ratio;
    if (!(ratio <= 1.0e-12)) return maus_fail(ctx, MAUS_E_STATE, "maus_heev: Jacobi sweeps did not converge (non-finite or non-Hermitian input?)");
    return MAUS_OK;
}
