// gmres.cu -- batched restarted GMRES for the candidate population (replaces scipy.sparse.linalg.gmres at AMS:89).
//
// Mirrors scipy 1.18.1 `_isolve/iterative.py:gmres` as the reference calls it (rtol 1e-8 through the tol->rtol shim,
// restart = min(20, n), maxiter = 50 restart cycles, x0 = b, left preconditioning by M): modified Gram-Schmidt in the
// same order, breakdown test h1 <= eps*h0, Givens rotations (LAPACK zlartg convention), inner stop on the
// PRECONDITIONED residual with the adaptive `ptol` of gh-8400, pseudo-solve of the Hessenberg system, the true
// residual recomputed every cycle, info = 0 iff ||b - Hx|| <= atol.
//
// B200 design: all candidates advance in lock-step (a candidate that left its inner loop waits, masked, for the
// cycle end), so every operator application is ONE batched product with the shared matrix -- a DMMA GEMM A*[v_1..v_C]
// (dense; FP64 tensor pipe) or one CSR SpMM (sparse; the matrix is streamed from HBM once per 4 candidates) -- and
// the per-candidate shift/regulariser enters as  H_c z = A z + (psi_c - sigma_c) z (+ R_c z only when it is above
// rounding, see perturb gate).  Dot products / norms use warp-shuffle + block reductions with per-block partials
// that the consumer kernel re-sums in fixed order (deterministic, no atomics).  The tiny Hessenberg / Givens / ptol
// logic runs in one thread per candidate on the device; the host only reads two counters per inner iteration.
#include <vector>
#include <algorithm>
#include <cfloat>
#include <functional>
#include <cooperative_groups.h>
#include "ctx.cuh"
#include "gmres.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int GM_RESTART = 20;
constexpr int GM_MAXITER = 50;
constexpr int GM_NT = 256;
constexpr int GM_MAXBLK = 64;     // partial-sum blocks per candidate

struct GmresCand {
    // Hessenberg data, scipy layout: h[col][k] = H(k, col)
    cplx h[GM_RESTART][GM_RESTART + 1];
    cplx givens[GM_RESTART][2];
    cplx S[GM_RESTART + 1];
    cplx y[GM_RESTART];
    cplx coef;             // MGS coefficient to apply to w in the next pass
    cplx shift;            // psi - sigma
    double psi;
    double bnrm2, atol, Mb_nrm2, ptol, ptol_max_factor, presid, rnorm, h0;
    double scale;          // 1/norm applied to the vector being stored
    int col;               // current inner column
    int last_col;          // column at which the inner loop stopped
    int inner_iter;
    int breakdown;
    int inner_active;      // still inside the Arnoldi loop of this cycle
    int active;            // not finished overall
    int info;
    int jac_on;            // Jacobi preconditioner valid and requested
    int perturb;           // apply the random perturbation R_c in matvecs
    int masked;            // candidate was masked out on entry (status != 0)
};

struct GmresWs {
    long long n = 0, C = 0;
    int m = 0, m_cap = 0, nblk = 0;
    cplx *Vk = nullptr, *w = nullptr, *z = nullptr, *x = nullptr, *r = nullptr, *minv = nullptr;
    cplx* partial = nullptr;      // [C][GM_MAXBLK]
    GmresCand* cand = nullptr;
    int* counters = nullptr;      // [0] = inner_active count, [1] = active count
    int* jacbad = nullptr;        // [C]
    int* active_idx = nullptr;    // [C] indices of the candidates still iterating (matvec compaction)
    cplx *vc = nullptr, *zc = nullptr;   // [C][n] compacted matvec input / output
    int* host_poll = nullptr;            // pinned [2]: inner-active counts read back with a lag of one iteration
    cudaEvent_t poll_ev[2] = {nullptr, nullptr};
    size_t bytes = 0;
};

__device__ __forceinline__ cplx block_sum_c(cplx v, cplx* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    cplx r = cmake(0.0, 0.0);
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { r.x += sh[q].x; r.y += sh[q].y; }
    return r;
}
__device__ __forceinline__ cplx sum_partials(const cplx* partial, int b, int nblk) {
    cplx s = cmake(0.0, 0.0);
    for (int q = 0; q < nblk; ++q) { cplx p = partial[b * GM_MAXBLK + q]; s.x += p.x; s.y += p.y; }
    return s;
}
__device__ __forceinline__ void chunk_range(long long n, int nblk, int blk, long long& i0, long long& i1) {
    long long per = (n + nblk - 1) / nblk;
    i0 = (long long)blk * per;
    i1 = i0 + per < n ? i0 + per : n;
}

// ---- matvec compaction ---------------------------------------------------------------------------------------------------
// Candidates leave the iteration at different times (a Jacobi-preconditioned one after a handful of steps, a plain one after
// hundreds: BASELINE config 4).  The matvec is the only kernel that cannot skip a finished candidate by itself (a GEMM / SpMM
// over all C columns), so once a quarter of the batch is done the still-active columns are gathered into a compact block.
__global__ void gm_list_active_kernel(const GmresCand* __restrict__ cand, int C, int* __restrict__ idx) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int a = 0;
    for (int b = 0; b < C; ++b) if (cand[b].active) idx[a++] = b;      // increasing order: deterministic
}
__global__ void gm_gather_kernel(const cplx* __restrict__ src, long long ld, const int* __restrict__ idx, cplx* __restrict__ dst,
                                 long long n, int nblk) {
    const int a = blockIdx.y;
    long long i0, i1; chunk_range(n, nblk, blockIdx.x, i0, i1);
    const cplx* s = src + (long long)idx[a] * ld;
    cplx* d = dst + (long long)a * n;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) d[i] = s[i];
}
__global__ void gm_scatter_kernel(const cplx* __restrict__ src, const int* __restrict__ idx, cplx* __restrict__ dst, long long ld,
                                  long long n, int nblk) {
    const int a = blockIdx.y;
    long long i0, i1; chunk_range(n, nblk, blockIdx.x, i0, i1);
    const cplx* s = src + (long long)a * n;
    cplx* d = dst + (long long)idx[a] * ld;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) d[i] = s[i];
}

// ---- setup ------------------------------------------------------------------------------------------------------
// Jacobi: d_i = A_ii - sigma + psi (+ R_ii);  minv = 1/d when requested; validity (all finite, |d| > 1e-12) -> jacbad
__global__ void gm_precond_kernel(const cplx* __restrict__ diagA, long long n, const cplx* __restrict__ sigma,
                                  const double* __restrict__ psi, const unsigned long long* __restrict__ keys,
                                  const unsigned char* __restrict__ jac, cplx* __restrict__ minv, int* jacbad, int nblk,
                                  long long row0) {
    const int b = blockIdx.y;
    long long i0, i1; chunk_range(n, nblk, blockIdx.x, i0, i1);
    const bool want = jac && jac[b];
    int bad = 0;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        cplx mi = cmake(1.0, 0.0);
        if (want) {
            cplx d = diagA[i];
            d.x -= sigma[b].x; d.y -= sigma[b].y;           // T = A - sigma I (AMS:270)
            cplx reg = cmake(psi[b], 0.0);                    // + psi I (+ R_ii), AMS:47-52
            if (keys) { cplx r = psi_perturbation(keys[b], (uint32_t)(row0 + i), (uint32_t)(row0 + i), psi[b]); reg.x += r.x; reg.y += r.y; }
            d.x += reg.x; d.y += reg.y;
            mi = crecip(d);                                   // AMS:70
            const double ad = hypot(d.x, d.y);
            if (!cfinite(mi) || !(ad > 1e-12)) bad = 1;       // AMS:72
        }
        minv[(long long)b * n + i] = mi;
    }
    if (bad) atomicOr(&jacbad[b], 1);
}

__global__ void gm_init_kernel(GmresCand* cand, const cplx* __restrict__ sigma, const double* __restrict__ psi,
                               const unsigned char* __restrict__ jac, const int* __restrict__ jacbad,
                               const int* __restrict__ status, int C, double perturb_gate, int have_keys) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= C) return;
    GmresCand& g = cand[b];
    g.shift = cmake(psi[b] - sigma[b].x, -sigma[b].y);
    g.psi = psi[b];
    g.jac_on = (jac && jac[b] && !jacbad[b]) ? 1 : 0;
    g.perturb = (have_keys && psi[b] * 0.15 > perturb_gate) ? 1 : 0;
    g.active = (status[b] == 0) ? 1 : 0;
    g.masked = g.active ? 0 : 1;
    g.inner_active = 0; g.info = 0; g.inner_iter = 0; g.breakdown = 0; g.col = 0; g.last_col = 0;
    g.ptol_max_factor = 1.0; g.presid = 0.0; g.rnorm = 0.0; g.coef = cmake(0.0, 0.0); g.scale = 1.0;
}

// partial sums of |rhs|^2 and |M rhs|^2 ; also x = rhs (x0 = b, AMS:61)
__global__ void gm_norms_b_kernel(const cplx* __restrict__ rhs, long long rhs_stride, const cplx* __restrict__ minv,
                                  const GmresCand* __restrict__ cand, long long n, cplx* __restrict__ x, cplx* partial,
                                  int nblk) {
    __shared__ cplx sh[GM_NT / 32];
    const int b = blockIdx.y;
    long long i0, i1; chunk_range(n, nblk, blockIdx.x, i0, i1);
    const bool jac = cand[b].jac_on;
    double s1 = 0.0, s2 = 0.0;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        cplx v = rhs[(long long)b * rhs_stride + i];
        x[(long long)b * n + i] = v;
        s1 = fma(v.x, v.x, s1); s1 = fma(v.y, v.y, s1);
        cplx mv = jac ? cmul(minv[(long long)b * n + i], v) : v;
        s2 = fma(mv.x, mv.x, s2); s2 = fma(mv.y, mv.y, s2);
    }
    cplx t = block_sum_c(cmake(s1, s2), sh);
    if (threadIdx.x == 0) partial[b * GM_MAXBLK + blockIdx.x] = t;
}

__global__ void gm_after_norms_kernel(GmresCand* cand, const cplx* partial, int C, int nblk) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= C) return;
    GmresCand& g = cand[b];
    if (!g.active) return;
    cplx t = sum_partials(partial, b, nblk);
    g.bnrm2 = sqrt(t.x);
    g.Mb_nrm2 = sqrt(t.y);
    g.atol = fmax(0.0, 1e-8 * g.bnrm2);                     // _get_atol_rtol, rtol = 1e-8 (AMS:89)
    if (g.bnrm2 == 0.0) { g.active = 0; g.info = 0; return; }   // "if bnrm2 == 0: return b, 0"
    g.ptol = g.Mb_nrm2 * fmin(1.0, g.atol / g.bnrm2);
}

// r = rhs - (z + shift*x)   (z = A x already computed) ; partial |r|^2
__global__ void gm_residual_kernel(const cplx* __restrict__ rhs, long long rhs_stride, const cplx* __restrict__ z,
                                   const cplx* __restrict__ x, const GmresCand* __restrict__ cand, long long n,
                                   cplx* __restrict__ r, cplx* partial, int nblk) {
    __shared__ cplx sh[GM_NT / 32];
    const int b = blockIdx.y;
    if (!cand[b].active) return;
    long long i0, i1; chunk_range(n, nblk, blockIdx.x, i0, i1);
    const cplx sft = cand[b].shift;
    double s = 0.0;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const long long o = (long long)b * n + i;
        cplx hx = z[o];
        cfma(hx, sft, x[o]);
        cplx rv = csub(rhs[(long long)b * rhs_stride + i], hx);
        r[o] = rv;
        s = fma(rv.x, rv.x, s); s = fma(rv.y, rv.y, s);
    }
    cplx t = block_sum_c(cmake(s, 0.0), sh);
    if (threadIdx.x == 0) partial[b * GM_MAXBLK + blockIdx.x] = t;
}

// decisions after the true residual: first = 1 -> the pre-loop test "if norm(r) < atol: return x, 0"
__global__ void gm_cycle_end_kernel(GmresCand* cand, const cplx* partial, int C, int nblk, int first, int last_cycle,
                                    int* counters) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= C) return;
    GmresCand& g = cand[b];
    if (!g.active) return;
    const double rnorm = sqrt(sum_partials(partial, b, nblk).x);
    g.rnorm = rnorm;
    const double eps = DBL_EPSILON;
    if (first) {
        if (rnorm < g.atol) { g.active = 0; g.info = 0; }
    } else {
        bool stop = false;
        if (rnorm <= g.atol) stop = true;
        else if (g.breakdown) stop = true;
        else if (g.presid <= g.ptol) g.ptol_max_factor = fmax(eps, 0.25 * g.ptol_max_factor);
        else g.ptol_max_factor = fmin(1.0, 1.5 * g.ptol_max_factor);
        if (!stop) g.ptol = g.presid * fmin(g.ptol_max_factor, g.atol / rnorm);
        if (stop || last_cycle) { g.active = 0; g.info = (rnorm <= g.atol) ? 0 : GM_MAXITER; }
    }
    if (g.active) atomicAdd(&counters[1], 1);
}

// cycle start: w = M r ; partial |w|^2
__global__ void gm_start_kernel(const cplx* __restrict__ r, const cplx* __restrict__ minv, const GmresCand* __restrict__ cand,
                                long long n, cplx* __restrict__ w, cplx* partial, int nblk) {
    __shared__ cplx sh[GM_NT / 32];
    const int b = blockIdx.y;
    if (!cand[b].active) return;
    long long i0, i1; chunk_range(n, nblk, blockIdx.x, i0, i1);
    const bool jac = cand[b].jac_on;
    double s = 0.0;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const long long o = (long long)b * n + i;
        cplx v = jac ? cmul(minv[o], r[o]) : r[o];
        w[o] = v;
        s = fma(v.x, v.x, s); s = fma(v.y, v.y, s);
    }
    cplx t = block_sum_c(cmake(s, 0.0), sh);
    if (threadIdx.x == 0) partial[b * GM_MAXBLK + blockIdx.x] = t;
}

__global__ void gm_start_scalar_kernel(GmresCand* cand, const cplx* partial, int C, int nblk, int* counters) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= C) return;
    GmresCand& g = cand[b];
    if (!g.active) { g.inner_active = 0; return; }
    const double tmp = sqrt(sum_partials(partial, b, nblk).x);
    for (int k = 0; k <= GM_RESTART; ++k) g.S[k] = cmake(0.0, 0.0);
    for (int c = 0; c < GM_RESTART; ++c)
        for (int k = 0; k <= GM_RESTART; ++k) g.h[c][k] = cmake(0.0, 0.0);
    g.S[0] = cmake(tmp, 0.0);
    g.scale = 1.0 / tmp;                                   // v[0] *= (1 / tmp)
    g.breakdown = 0; g.col = 0; g.inner_active = 1;
    atomicAdd(&counters[0], 1);
}

// Vk[slot] = w * scale   (for inner-active candidates)
__global__ void gm_store_kernel(const cplx* __restrict__ w, const GmresCand* __restrict__ cand, long long n, int m,
                                cplx* __restrict__ Vk, int slot_is_col_plus1, int nblk, int require_inner) {
    const int b = blockIdx.y;
    const GmresCand& g = cand[b];
    if (!g.active || (require_inner && !g.inner_active)) return;
    const int slot = slot_is_col_plus1 ? g.col + 1 : 0;
    long long i0, i1; chunk_range(n, nblk, blockIdx.x, i0, i1);
    const double sc = g.scale;
    cplx* dst = Vk + ((long long)b * (m + 1) + slot) * n;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        cplx v = w[(long long)b * n + i];
        dst[i] = cmake(v.x * sc, v.y * sc);
    }
}

// after the batched matvec z = A v[col]:  w = M (z + shift v[col] [+ R v[col]]) ; partial |w|^2 (h0)
__global__ void gm_post_matvec_kernel(const cplx* __restrict__ z, const cplx* __restrict__ Vk, const cplx* __restrict__ minv,
                                      const GmresCand* __restrict__ cand, long long n, int m, int col, cplx* __restrict__ w,
                                      cplx* partial, int nblk) {
    __shared__ cplx sh[GM_NT / 32];
    const int b = blockIdx.y;
    const GmresCand& g = cand[b];
    if (!g.active || !g.inner_active) return;
    long long i0, i1; chunk_range(n, nblk, blockIdx.x, i0, i1);
    const cplx* v = Vk + ((long long)b * (m + 1) + col) * n;
    const bool jac = g.jac_on;
    const cplx sft = g.shift;
    double s = 0.0;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const long long o = (long long)b * n + i;
        cplx av = z[o];
        cfma(av, sft, v[i]);
        cplx wv = jac ? cmul(minv[o], av) : av;
        w[o] = wv;
        s = fma(wv.x, wv.x, s); s = fma(wv.y, wv.y, s);
    }
    cplx t = block_sum_c(cmake(s, 0.0), sh);
    if (threadIdx.x == 0) partial[b * GM_MAXBLK + blockIdx.x] = t;
}

// z[c] += R_c v   (dense perturbation of AMS:49, generated on the fly; only launched when a candidate's psi makes it
// larger than rounding).  One warp per row.
__global__ void gm_perturb_matvec_kernel(cplx* __restrict__ z, long long ldz, const cplx* __restrict__ v, long long ldv,
                                         const GmresCand* __restrict__ cand, const unsigned long long* __restrict__ keys,
                                         long long n) {
    const int b = blockIdx.y;
    if (!cand[b].perturb || !cand[b].active) return;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const double ps = cand[b].psi;
    const unsigned long long key = keys[b];
    cplx acc = cmake(0.0, 0.0);
    for (long long j = lane; j < n; j += 32) cfma(acc, psi_perturbation(key, (uint32_t)row, (uint32_t)j, ps), v[(long long)b * ldv + j]);
    acc = warp_sum(acc);
    if (lane == 0) { cplx* d = &z[(long long)b * ldz + row]; d->x += acc.x; d->y += acc.y; }
}

// h0 = |w| recorded ; coef = 0 before the first MGS pass
__global__ void gm_h0_kernel(GmresCand* cand, const cplx* partial, int C, int nblk) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= C) return;
    GmresCand& g = cand[b];
    if (!g.active || !g.inner_active) return;
    g.h0 = sqrt(sum_partials(partial, b, nblk).x);
    g.coef = cmake(0.0, 0.0);
}

// MGS pass k:  w -= coef * v[k-1] (k > 0) ; partial <v[k], w>     (k == col+1: final pass -> partial |w|^2)
__global__ void gm_mgs_pass_kernel(cplx* __restrict__ w, const cplx* __restrict__ Vk, const GmresCand* __restrict__ cand,
                                   long long n, int m, int k, int final_pass, cplx* partial, int nblk) {
    __shared__ cplx sh[GM_NT / 32];
    const int b = blockIdx.y;
    const GmresCand& g = cand[b];
    if (!g.active || !g.inner_active) return;
    long long i0, i1; chunk_range(n, nblk, blockIdx.x, i0, i1);
    const cplx cf = g.coef;
    const cplx* vprev = Vk + ((long long)b * (m + 1) + (k - 1)) * n;
    const cplx* vk = Vk + ((long long)b * (m + 1) + k) * n;
    cplx acc = cmake(0.0, 0.0);
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const long long o = (long long)b * n + i;
        cplx wv = w[o];
        if (k > 0) { cfms(wv, cf, vprev[i]); w[o] = wv; }
        if (final_pass) { acc.x = fma(wv.x, wv.x, acc.x); acc.x = fma(wv.y, wv.y, acc.x); }
        else cfma_conj(acc, vk[i], wv);                    // vdot(v[k], w) = sum conj(v) * w
    }
    cplx t = block_sum_c(acc, sh);
    if (threadIdx.x == 0) partial[b * GM_MAXBLK + blockIdx.x] = t;
}

__global__ void gm_mgs_scalar_kernel(GmresCand* cand, const cplx* partial, int C, int nblk, int k) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= C) return;
    GmresCand& g = cand[b];
    if (!g.active || !g.inner_active) return;
    cplx t = sum_partials(partial, b, nblk);
    g.h[g.col][k] = t;
    g.coef = t;
}

// LAPACK zlartg: [c s; -conj(s) c] [f; g] = [r; 0], c real
__device__ void zlartg_dev(cplx f, cplx g, double& c, cplx& s, cplx& r) {
    if (g.x == 0.0 && g.y == 0.0) { c = 1.0; s = cmake(0.0, 0.0); r = f; return; }
    const double f1 = hypot(f.x, f.y), g1 = hypot(g.x, g.y);
    if (f1 == 0.0) { c = 0.0; s = cmake(g.x / g1, -g.y / g1); r = cmake(g1, 0.0); return; }
    const double d = hypot(f1, g1);
    c = f1 / d;
    const cplx fu = cmake(f.x / f1, f.y / f1);             // f / |f|
    const cplx gc = cmake(g.x / d, -g.y / d);              // conj(g) / d
    s = cmul(fu, gc);
    r = cmake(fu.x * d, fu.y * d);
}

// end of an inner iteration (scipy lines 772-805): h1, breakdown, Givens, presid, loop exit test.  The column h[col][0..col] is
// already in g.h; returns true when the candidate leaves the inner loop after storing v[col + 1] once more.
__device__ bool gm_hess_update(GmresCand& g, double h1, int m) {
    const int col = g.col;
    g.h[col][col + 1] = cmake(h1, 0.0);
    if (h1 <= DBL_EPSILON * g.h0) { g.h[col][col + 1] = cmake(0.0, 0.0); g.breakdown = 1; g.scale = 1.0; }
    else g.scale = 1.0 / h1;
    for (int k = 0; k < col; ++k) {
        const cplx c = g.givens[k][0], s = g.givens[k][1];
        const cplx n0 = g.h[col][k], n1 = g.h[col][k + 1];
        cplx a = cmul(c, n0); cfma(a, s, n1);                              // c*n0 + s*n1
        cplx bb = cmul(c, n1); cplx sc = cmake(-s.x, s.y); cfma(bb, sc, n0);   // -conj(s)*n0 + c*n1
        g.h[col][k] = a; g.h[col][k + 1] = bb;
    }
    double c; cplx s, mag;
    zlartg_dev(g.h[col][col], g.h[col][col + 1], c, s, mag);
    g.givens[col][0] = cmake(c, 0.0); g.givens[col][1] = s;
    g.h[col][col] = mag; g.h[col][col + 1] = cmake(0.0, 0.0);
    const cplx tmp = cmul(cmake(-s.x, s.y), g.S[col]);                     // -conj(s) * S[col]
    g.S[col] = cmake(c * g.S[col].x, c * g.S[col].y);
    g.S[col + 1] = tmp;
    g.presid = hypot(tmp.x, tmp.y);
    g.inner_iter += 1;
    g.last_col = col;
    return g.presid <= g.ptol || g.breakdown || col + 1 >= m;
}
__global__ void gm_hess_kernel(GmresCand* cand, const cplx* partial, int C, int nblk, int m, int* counters) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= C) return;
    GmresCand& g = cand[b];
    if (!g.active || !g.inner_active) return;
    const double h1 = sqrt(sum_partials(partial, b, nblk).x);
    if (gm_hess_update(g, h1, m)) g.inner_active = 2;                      // 2 = store v[col+1] once more, then leave
    else atomicAdd(&counters[0], 1);
}

// ---- whole inner iteration in ONE launch for vectors that fit a thread-block cluster (n <= 32768: the dense configurations) ----
// After the batched matvec z = A v[col], a cluster of up to 8 CTAs owns one candidate and keeps its slice of w in REGISTERS:
//   w = M (z + shift v[col]) ; h0 = |w| ; for k <= col: h[col][k] = <v[k], w>, w -= h[col][k] v[k] ; h1 = |w| ;
//   Hessenberg / Givens / stopping logic (CTA 0, one thread) ; v[col + 1] = w * scale ; advance.
// Same arithmetic sequence as the multi-launch path (modified Gram-Schmidt in scipy's order), but every v[k] is read once
// instead of three vector passes per k, and the 2 (col + 1) + 6 launches of the multi-launch path become one.  Dot products:
// warp shuffle -> CTA (fixed order) -> distributed shared memory, one cluster barrier per reduction, summed in CTA-rank order
// by every CTA (deterministic).
constexpr int GA_NT = 256, GA_MAXC = 8;
template <int EPT>
__global__ void __launch_bounds__(GA_NT) gm_arnoldi_cluster_kernel(const cplx* __restrict__ z, cplx* __restrict__ Vk,
                                                                   const cplx* __restrict__ minv, GmresCand* cand, long long n,
                                                                   int m, int col, int* counters) {
    cg::cluster_group cluster = cg::this_cluster();
    const int NC = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int b = blockIdx.x / NC;
    GmresCand& g = cand[b];
    if (!g.active || !g.inner_active) return;                  // uniform over the cluster: nobody reaches a cluster barrier
    __shared__ cplx slots[2][GA_MAXC];
    __shared__ cplx wsum[GA_NT / 32];
    __shared__ double s_scale;
    __shared__ int s_leave;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int red = 0;
    auto cluster_sum = [&](cplx v) -> cplx {
        v = warp_sum(v);
        if (lane == 0) wsum[warp] = v;
        __syncthreads();
        if (t == 0) {
            cplx tot = cmake(0.0, 0.0);
            for (int q = 0; q < GA_NT / 32; ++q) { tot.x += wsum[q].x; tot.y += wsum[q].y; }
            for (int d = 0; d < NC; ++d) *cluster.map_shared_rank(&slots[red & 1][rank], d) = tot;
        }
        cluster.sync();
        cplx r = cmake(0.0, 0.0);
        for (int d = 0; d < NC; ++d) { const cplx p = slots[red & 1][d]; r.x += p.x; r.y += p.y; }
        ++red;
        return r;
    };
    const long long stride = (long long)NC * GA_NT, i0 = (long long)rank * GA_NT + t;
    const cplx* zb = z + (long long)b * n;
    const cplx* mb = minv + (long long)b * n;
    cplx* vb = Vk + (long long)b * (m + 1) * n;
    const bool jac = g.jac_on;
    const cplx sft = g.shift;
    cplx w[EPT];
    double s = 0.0;
    {
        const cplx* v = vb + (long long)col * n;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const long long i = i0 + e * stride;
            cplx wv = cmake(0.0, 0.0);
            if (i < n) {
                cplx av = zb[i];
                cfma(av, sft, v[i]);
                wv = jac ? cmul(mb[i], av) : av;
            }
            w[e] = wv;
            s = fma(wv.x, wv.x, s); s = fma(wv.y, wv.y, s);
        }
    }
    const double h0 = sqrt(cluster_sum(cmake(s, 0.0)).x);
    for (int k = 0; k <= col; ++k) {
        const cplx* vk = vb + (long long)k * n;
        cplx vv[EPT];
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const long long i = i0 + e * stride;
            vv[e] = (i < n) ? vk[i] : cmake(0.0, 0.0);
            cfma_conj(acc, vv[e], w[e]);                        // vdot(v[k], w) = sum conj(v) * w
        }
        const cplx hk = cluster_sum(acc);
        if (rank == 0 && t == 0) g.h[col][k] = hk;
#pragma unroll
        for (int e = 0; e < EPT; ++e) cfms(w[e], hk, vv[e]);
    }
    double s2 = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) { s2 = fma(w[e].x, w[e].x, s2); s2 = fma(w[e].y, w[e].y, s2); }
    const double h1 = sqrt(cluster_sum(cmake(s2, 0.0)).x);
    if (rank == 0 && t == 0) {
        g.h0 = h0;
        const bool leave = gm_hess_update(g, h1, m);
        const double sc = g.scale;
        for (int d = 0; d < NC; ++d) { *cluster.map_shared_rank(&s_scale, d) = sc; *cluster.map_shared_rank(&s_leave, d) = leave ? 1 : 0; }
    }
    cluster.sync();
    const double sc = s_scale;
    {
        cplx* vn = vb + (long long)(col + 1) * n;              // v[col + 1] = w * scale (stored also when leaving: scipy line 806)
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const long long i = i0 + e * stride;
            if (i < n) vn[i] = cmake(w[e].x * sc, w[e].y * sc);
        }
    }
    if (rank == 0 && t == 0) {
        if (s_leave) g.inner_active = 0;
        else { g.col = col + 1; atomicAdd(&counters[0], 1); }
    }
}

__global__ void gm_advance_kernel(GmresCand* cand, int C) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= C) return;
    GmresCand& g = cand[b];
    if (!g.active) return;
    if (g.inner_active == 2) g.inner_active = 0;
    else if (g.inner_active == 1) g.col += 1;
}

// Hessenberg back substitution with pseudo-solve (scipy lines 812-823)
__global__ void gm_solve_y_kernel(GmresCand* cand, int C) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= C) return;
    GmresCand& g = cand[b];
    if (!g.active) return;
    const int col = g.last_col;
    if (g.h[col][col].x == 0.0 && g.h[col][col].y == 0.0) g.S[col] = cmake(0.0, 0.0);
    for (int k = 0; k <= col; ++k) g.y[k] = g.S[k];
    for (int k = col; k > 0; --k) {
        if (g.y[k].x != 0.0 || g.y[k].y != 0.0) {
            g.y[k] = cdiv(g.y[k], g.h[k][k]);
            const cplx t = g.y[k];
            for (int i = 0; i < k; ++i) cfms(g.y[i], t, g.h[k][i]);
        }
    }
    if (g.y[0].x != 0.0 || g.y[0].y != 0.0) g.y[0] = cdiv(g.y[0], g.h[0][0]);
}

// x += y @ v[:col+1]
__global__ void gm_update_x_kernel(cplx* __restrict__ x, const cplx* __restrict__ Vk, const GmresCand* __restrict__ cand,
                                   long long n, int m, int nblk) {
    const int b = blockIdx.y;
    const GmresCand& g = cand[b];
    if (!g.active) return;
    long long i0, i1; chunk_range(n, nblk, blockIdx.x, i0, i1);
    const int col = g.last_col;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        cplx acc = cmake(0.0, 0.0);
        for (int k = 0; k <= col; ++k) cfma(acc, g.y[k], Vk[((long long)b * (m + 1) + k) * n + i]);
        cplx* d = &x[(long long)b * n + i];
        d->x += acc.x; d->y += acc.y;
    }
}

__global__ void gm_finish_kernel(const GmresCand* __restrict__ cand, const cplx* __restrict__ x, long long n, cplx* __restrict__ X,
                                 int* status, int* iters, int nblk) {
    __shared__ int bad;
    const int b = blockIdx.y;
    if (cand[b].masked) return;                  // masked on entry
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    long long i0, i1; chunk_range(n, nblk, blockIdx.x, i0, i1);
    int mybad = 0;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        cplx v = x[(long long)b * n + i];
        X[(long long)b * n + i] = v;
        if (!cfinite(v)) mybad = 1;
    }
    if (mybad) bad = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) iters[b] = cand[b].inner_iter;
        if (cand[b].info != 0) atomicMax(&status[b], MAUS_ST_GMRES_NOCONV);   // AMS:90 (checked before AMS:94)
        else if (bad) atomicCAS(&status[b], 0, MAUS_ST_NONFINITE);            // AMS:94-95
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
void maus_gmres_free(maus_ctx* ctx) {
    GmresWs* ws = (GmresWs*)ctx->gmres;
    if (!ws) return;
    cudaFree(ws->Vk); cudaFree(ws->w); cudaFree(ws->z); cudaFree(ws->x); cudaFree(ws->r); cudaFree(ws->minv);
    cudaFree(ws->partial); cudaFree(ws->cand); cudaFree(ws->counters); cudaFree(ws->jacbad); cudaFree(ws->active_idx); cudaFree(ws->vc); cudaFree(ws->zc);
    if (ws->host_poll) cudaFreeHost(ws->host_poll);
    for (auto& e : ws->poll_ev) if (e) cudaEventDestroy(e);
    ctx->bytes_held -= (long long)ws->bytes;
    delete ws;
    ctx->gmres = nullptr;
}

static int gmres_ensure(maus_ctx* ctx, long long n, long long C, GmresWs** out) {
    GmresWs* ws = (GmresWs*)ctx->gmres;
    if (ws && ws->n == n && ws->C >= C) { *out = ws; return MAUS_OK; }
    if (ws) { MAUS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); maus_gmres_free(ctx); }
    ws = new GmresWs();
    ws->n = n; ws->C = std::max<long long>(C, 4);
    ws->m = ws->m_cap = GM_RESTART;          // Krylov slots are sized for the full restart length
    ws->nblk = (int)std::max<long long>(1, std::min<long long>(GM_MAXBLK, (n + 4 * GM_NT - 1) / (4 * GM_NT)));
    const size_t vec = (size_t)ws->C * n * sizeof(cplx);
    cudaError_t e;
    size_t total = 0;
#define GM_ALLOC(p, bytes) do { e = cudaMalloc((void**)&(p), (bytes)); if (e != cudaSuccess) { ctx->gmres = ws; maus_gmres_free(ctx); return maus_fail(ctx, MAUS_E_NOMEM, "GMRES workspace", e); } total += (bytes); } while (0)
    GM_ALLOC(ws->Vk, vec * (ws->m + 1));
    GM_ALLOC(ws->w, vec); GM_ALLOC(ws->z, vec); GM_ALLOC(ws->x, vec); GM_ALLOC(ws->r, vec); GM_ALLOC(ws->minv, vec);
    GM_ALLOC(ws->partial, (size_t)ws->C * GM_MAXBLK * sizeof(cplx));
    GM_ALLOC(ws->cand, (size_t)ws->C * sizeof(GmresCand));
    GM_ALLOC(ws->counters, 2 * sizeof(int));
    GM_ALLOC(ws->jacbad, (size_t)ws->C * sizeof(int));
    GM_ALLOC(ws->active_idx, (size_t)ws->C * sizeof(int));
    // compaction buffers are part of the workspace: allocating them lazily could fail on ONE rank of a row-sharded solve
    // and make the ranks issue different numbers of collectives
    GM_ALLOC(ws->vc, vec); GM_ALLOC(ws->zc, vec);
#undef GM_ALLOC
    if (cudaHostAlloc((void**)&ws->host_poll, 2 * sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&ws->poll_ev[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ws->poll_ev[1], cudaEventDisableTiming) != cudaSuccess) {
        ctx->gmres = ws; ws->bytes = total; ctx->bytes_held += (long long)total; maus_gmres_free(ctx);
        return maus_fail(ctx, MAUS_E_NOMEM, "GMRES poll buffers");
    }
    ws->bytes = total;
    ctx->bytes_held += (long long)total;
    ctx->gmres = ws;
    *out = ws;
    return MAUS_OK;
}

static int launch_arnoldi_cluster(maus_ctx* ctx, GmresWs* ws, long long n, int m, int col, long long C, cudaStream_t st) {
    int ept = 1;
    while ((n + ept - 1) / ept > (long long)GA_MAXC * GA_NT) ept *= 2;       // n <= 32768 -> ept <= 16
    const int nc = (int)(((n + ept - 1) / ept + GA_NT - 1) / GA_NT);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(C * nc), 1, 1);
    cfg.blockDim = dim3(GA_NT, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cplx* z = ws->z; cplx* Vk = ws->Vk; const cplx* minv = ws->minv; GmresCand* cand = ws->cand; int* counters = ws->counters;
    cudaError_t e;
    switch (ept) {
        case 1: e = cudaLaunchKernelEx(&cfg, gm_arnoldi_cluster_kernel<1>, z, Vk, minv, cand, n, m, col, counters); break;
        case 2: e = cudaLaunchKernelEx(&cfg, gm_arnoldi_cluster_kernel<2>, z, Vk, minv, cand, n, m, col, counters); break;
        case 4: e = cudaLaunchKernelEx(&cfg, gm_arnoldi_cluster_kernel<4>, z, Vk, minv, cand, n, m, col, counters); break;
        case 8: e = cudaLaunchKernelEx(&cfg, gm_arnoldi_cluster_kernel<8>, z, Vk, minv, cand, n, m, col, counters); break;
        default: e = cudaLaunchKernelEx(&cfg, gm_arnoldi_cluster_kernel<16>, z, Vk, minv, cand, n, m, col, counters); break;
    }
    MAUS_CUDA(ctx, e);
    return MAUS_OK;
}

// The solver proper, written against GmresOperator (gmres.cuh): `n` is the LOCAL vector length, op.matvec applies the
// shared matrix to C vectors, op.reduce_partials / op.flag_sync make partial sums / flags global in row-sharded mode.
int gmres_core(maus_ctx* ctx, const GmresOperator& op, long long C, const cplx* sigma, const double* psi,
               const unsigned long long* keys, const unsigned char* use_jacobi, const cplx* rhs, long long rhs_stride, cplx* X,
               int* status, int* iters, double max_psi_host) {
    GmresWs* ws = nullptr;
    const long long n = op.nloc;
    MausNvtxRange nvtx_range("maus.gmres");
    int rc = gmres_ensure(ctx, n, C, &ws); if (rc) return rc;
    // restart = min(20, GLOBAL order) (scipy), not the local slice length
    ws->m = (int)std::min<long long>(GM_RESTART, std::min<long long>(op.nglobal, ws->m_cap));
    const int m = ws->m, nblk = ws->nblk;
    cudaStream_t st = ctx->stream;
    const dim3 gridv(nblk, (unsigned)C), gridc((unsigned)((C + 127) / 128));
    // the random perturbation only matters once 0.15*psi exceeds rounding of the matvec (~1e-17 * max|a_ij|)
    const double gate = 1e-17 * std::max(op.amax, 1e-300);
    const bool any_perturb = op.dense && keys && (0.15 * max_psi_host > gate);
    // one-launch Arnoldi step for vectors that fit a cluster (MAUS_GMRES_CLUSTER=0: the multi-launch path, A/B measurements)
    static int cluster_env = -1;
    if (cluster_env < 0) { const char* e = getenv("MAUS_GMRES_CLUSTER"); cluster_env = e ? (atoi(e) != 0) : 1; }
    const bool use_cluster = cluster_env && !op.reduce_partials && n <= (long long)GA_MAXC * GA_NT * 16;
    int host_counters[2];
    // row-sharded mode: every rank holds a slice of each vector, so the per-block partial sums are combined over the ranks
    // (one kernel over NVLink peer memory, rowshard.cu) before the scalar kernels consume them
    auto sync_partials = [&]() -> int {
        if (!op.reduce_partials) return MAUS_OK;
        return op.reduce_partials(ws->partial, GM_MAXBLK, nblk, C);
    };
#define GM_SYNC() do { if ((rc = sync_partials())) return rc; } while (0)

    MAUS_CUDA(ctx, cudaMemsetAsync(ws->jacbad, 0, (size_t)C * sizeof(int), st));
    gm_precond_kernel<<<gridv, GM_NT, 0, st>>>(op.diag, n, sigma, psi, keys, use_jacobi, ws->minv, ws->jacbad, nblk, op.row0);
    if (op.flag_sync && (rc = op.flag_sync(ws->jacbad, C))) return rc;
    gm_init_kernel<<<gridc, 128, 0, st>>>(ws->cand, sigma, psi, use_jacobi, ws->jacbad, status, (int)C, gate, keys != nullptr);
    gm_norms_b_kernel<<<gridv, GM_NT, 0, st>>>(rhs, rhs_stride, ws->minv, ws->cand, n, ws->x, ws->partial, nblk);
    GM_SYNC();
    gm_after_norms_kernel<<<gridc, 128, 0, st>>>(ws->cand, ws->partial, (int)C, nblk);
    ctx->launches += 4;

    long long n_compact = 0;          // > 0: ws->active_idx lists that many still-active candidates, matvecs run on them only
    auto matvec = [&](const cplx* v, long long ldv) -> int {
        int r2;
        if (n_compact > 0) {
            const dim3 gridg(nblk, (unsigned)n_compact);
            gm_gather_kernel<<<gridg, GM_NT, 0, st>>>(v, ldv, ws->active_idx, ws->vc, n, nblk);
            r2 = op.matvec(ws->vc, n, ws->zc, n, n_compact);
            if (r2) return r2;
            gm_scatter_kernel<<<gridg, GM_NT, 0, st>>>(ws->zc, ws->active_idx, ws->z, n, n, nblk);
            ctx->launches += 2;
        } else {
            r2 = op.matvec(v, ldv, ws->z, n, C);
            if (r2) return r2;
        }
        if (any_perturb) {
            // launched only if some candidate's psi puts R above rounding; per-candidate test inside the kernel
            const int wpb = 8;
            dim3 g((unsigned)((n + wpb - 1) / wpb), (unsigned)C);
            gm_perturb_matvec_kernel<<<g, wpb * 32, 0, st>>>(ws->z, n, v, ldv, ws->cand, keys, n);
            ctx->launches += 1;
        }
        return MAUS_OK;
    };

    // r = b - H x0 ; pre-loop convergence test
    if ((rc = matvec(ws->x, n))) return rc;
    gm_residual_kernel<<<gridv, GM_NT, 0, st>>>(rhs, rhs_stride, ws->z, ws->x, ws->cand, n, ws->r, ws->partial, nblk);
    GM_SYNC();
    MAUS_CUDA(ctx, cudaMemsetAsync(ws->counters, 0, 2 * sizeof(int), st));
    gm_cycle_end_kernel<<<gridc, 128, 0, st>>>(ws->cand, ws->partial, (int)C, nblk, 1, 0, ws->counters);
    ctx->launches += 2;
    MAUS_CUDA(ctx, cudaMemcpyAsync(host_counters, ws->counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaStreamSynchronize(st));
    int n_active = host_counters[1];

    for (int outer = 0; outer < GM_MAXITER && n_active > 0; ++outer) {
        MAUS_CUDA(ctx, cudaMemsetAsync(ws->counters, 0, 2 * sizeof(int), st));
        gm_start_kernel<<<gridv, GM_NT, 0, st>>>(ws->r, ws->minv, ws->cand, n, ws->w, ws->partial, nblk);
        GM_SYNC();
        gm_start_scalar_kernel<<<gridc, 128, 0, st>>>(ws->cand, ws->partial, (int)C, nblk, ws->counters);
        gm_store_kernel<<<gridv, GM_NT, 0, st>>>(ws->w, ws->cand, n, m, ws->Vk, 0, nblk, 1);
        ctx->launches += 3;
        // The host only needs the number of candidates still inside their Arnoldi loop to END the loop early.  It is read with a
        // lag of one iteration (the copy of iteration j is waited for after iteration j + 1 has been queued), so the stream
        // never drains between iterations; the kernels of a surplus iteration find no inner-active candidate and do nothing.
        int pending = -1;                                     // slot of host_poll that is in flight
        bool all_left = false;
        for (int col = 0; col < m && !all_left; ++col) {
            if ((rc = matvec(ws->Vk + (long long)col * n, (long long)(m + 1) * n))) return rc;
            MAUS_CUDA(ctx, cudaMemsetAsync(ws->counters, 0, sizeof(int), st));
            if (use_cluster) {
                if ((rc = launch_arnoldi_cluster(ctx, ws, n, m, col, C, st))) return rc;
                ctx->launches += 1;
            } else {
                gm_post_matvec_kernel<<<gridv, GM_NT, 0, st>>>(ws->z, ws->Vk, ws->minv, ws->cand, n, m, col, ws->w, ws->partial, nblk);
                GM_SYNC();
                gm_h0_kernel<<<gridc, 128, 0, st>>>(ws->cand, ws->partial, (int)C, nblk);
                for (int k = 0; k <= col; ++k) {
                    gm_mgs_pass_kernel<<<gridv, GM_NT, 0, st>>>(ws->w, ws->Vk, ws->cand, n, m, k, 0, ws->partial, nblk);
                    GM_SYNC();
                    gm_mgs_scalar_kernel<<<gridc, 128, 0, st>>>(ws->cand, ws->partial, (int)C, nblk, k);
                }
                gm_mgs_pass_kernel<<<gridv, GM_NT, 0, st>>>(ws->w, ws->Vk, ws->cand, n, m, col + 1, 1, ws->partial, nblk);
                GM_SYNC();
                gm_hess_kernel<<<gridc, 128, 0, st>>>(ws->cand, ws->partial, (int)C, nblk, m, ws->counters);
                gm_store_kernel<<<gridv, GM_NT, 0, st>>>(ws->w, ws->cand, n, m, ws->Vk, 1, nblk, 1);
                gm_advance_kernel<<<gridc, 128, 0, st>>>(ws->cand, (int)C);
                ctx->launches += 6 + 2 * (col + 1);
            }
            const int slot = col & 1;
            MAUS_CUDA(ctx, cudaMemcpyAsync(&ws->host_poll[slot], ws->counters, sizeof(int), cudaMemcpyDeviceToHost, st));
            MAUS_CUDA(ctx, cudaEventRecord(ws->poll_ev[slot], st));
            if (pending >= 0) {
                MAUS_CUDA(ctx, cudaEventSynchronize(ws->poll_ev[pending]));
                if (ws->host_poll[pending] == 0) all_left = true;          // every candidate had left one iteration ago
            }
            pending = slot;
        }
        gm_solve_y_kernel<<<gridc, 128, 0, st>>>(ws->cand, (int)C);
        gm_update_x_kernel<<<gridv, GM_NT, 0, st>>>(ws->x, ws->Vk, ws->cand, n, m, nblk);
        if ((rc = matvec(ws->x, n))) return rc;
        gm_residual_kernel<<<gridv, GM_NT, 0, st>>>(rhs, rhs_stride, ws->z, ws->x, ws->cand, n, ws->r, ws->partial, nblk);
        GM_SYNC();
        MAUS_CUDA(ctx, cudaMemsetAsync(ws->counters, 0, 2 * sizeof(int), st));
        gm_cycle_end_kernel<<<gridc, 128, 0, st>>>(ws->cand, ws->partial, (int)C, nblk, 0, outer == GM_MAXITER - 1, ws->counters);
        ctx->launches += 4;
        MAUS_CUDA(ctx, cudaMemcpyAsync(host_counters, ws->counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        MAUS_CUDA(ctx, cudaStreamSynchronize(st));
        n_active = host_counters[1];
        // compaction pays once a quarter of the batch has finished (and changes the passes of 4 of the SpMM / the GEMM width)
        if (n_active > 0 && n_active * 4 <= C * 3 && n_active != n_compact) {
            gm_list_active_kernel<<<1, 1, 0, st>>>(ws->cand, (int)C, ws->active_idx);
            ctx->launches += 1;
            n_compact = n_active;
        }
    }
    gm_finish_kernel<<<gridv, GM_NT, 0, st>>>(ws->cand, ws->x, n, X, status, iters, nblk);
    ctx->launches += 1;
    if (op.flag_sync) {
        // a non-finite entry on ANY rank's slice fails the candidate everywhere (AMS:94-95)
        if ((rc = op.flag_sync(status, C))) return rc;
    }
    MAUS_CUDA(ctx, cudaGetLastError());
    return MAUS_OK;
#undef GM_SYNC
}

// replicated matrix on one GPU (the default path)
int maus_gmres_solve(maus_ctx* ctx, long long C, const cplx* sigma, const double* psi, const unsigned long long* keys,
                     const unsigned char* use_jacobi, const cplx* rhs, long long rhs_stride, cplx* X, int* status,
                     int* iters, double max_psi_host) {
    MatrixSlot& s = ctx->slot[0];
    GmresOperator op;
    op.nloc = ctx->n; op.nglobal = ctx->n; op.row0 = 0;
    op.diag = s.diag; op.amax = s.amax; op.dense = s.dense && !s.sparse;
    op.matvec = [ctx](const cplx* v, long long ldv, cplx* z, long long ldz, long long Cn) { return maus_apply_matrix(ctx, 0, v, ldv, z, ldz, Cn); };
    if (s.sparse) keys = nullptr;                          // AMS:47: the sparse regulariser has no random part
    return gmres_core(ctx, op, C, sigma, psi, keys, use_jacobi, rhs, rhs_stride, X, status, iters, max_psi_host);
}
