// svd.cu -- SVD power-sweep branch of the candidate step (AMS:227-255, residual AMS:300-301): the first "next" row of
// SURVEY.md section 8f.  u = A v / ||A v||, v = A^H u / ||A^H u||, sigma = max of the two norms, residual
// ||A v - sigma u|| + ||A^H u - sigma v||.  Never calls the inverse-iteration solver; it is two batched matvecs per
// sweep plus one for the residual, with the rectangular matrix kept in four layouts so that both A and A^H products are
// row-streaming GEMVs (<= 8 candidates, HBM-bound) or plain DMMA GEMMs (more candidates).
#include <algorithm>
#include "ctx.cuh"
#include "vec.cuh"

struct SvdWs {
    long long rows = 0, cols = 0, Ccap = 0;
    cplx *A_rm = nullptr, *A_cm = nullptr, *AH_rm = nullptr, *AH_cm = nullptr;
    cplx *U = nullptr, *V = nullptr, *TU = nullptr, *TV = nullptr;
    double *sig1 = nullptr, *sigma = nullptr, *resid = nullptr;
    int* status = nullptr;
    size_t bytes = 0;
};

namespace {
constexpr int SV_NT = 256;

__global__ void __launch_bounds__(1024) svd_relayout_kernel(const cplx* __restrict__ A_rm, int rows, int cols, cplx* __restrict__ A_cm,
                                                            cplx* __restrict__ AH_rm, cplx* __restrict__ AH_cm) {
    __shared__ cplx tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int r = by + threadIdx.y, c = bx + threadIdx.x;            // A_rm[r][c]
    if (r < rows && c < cols) {
        const cplx a = A_rm[(long long)r * cols + c];
        tile[threadIdx.y][threadIdx.x] = a;
        AH_cm[(long long)r * cols + c] = cmake(a.x, -a.y);           // (A^H)(c, r), column-major cols x rows: c + r*cols
    }
    __syncthreads();
    const int orow = by + threadIdx.x, ocol = bx + threadIdx.y;      // element A[orow][ocol]
    if (orow < rows && ocol < cols) {
        const cplx a = tile[threadIdx.x][threadIdx.y];
        A_cm[orow + (long long)ocol * rows] = a;                     // column-major rows x cols
        AH_rm[(long long)ocol * rows + orow] = cmake(a.x, -a.y);     // (A^H) row-major cols x rows
    }
}

__device__ __forceinline__ double block_norm(const cplx* __restrict__ y, long long len, double* sh) {
    // overflow-safe 2-norm like BLAS dznrm2 (np.linalg.norm)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double amax = 0.0;
    for (long long i = threadIdx.x; i < len; i += SV_NT) amax = fmax(amax, fmax(fabs(y[i].x), fabs(y[i].y)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    __syncthreads();
    if (lane == 0) sh[warp] = amax;
    __syncthreads();
    amax = 0.0;
    for (int w = 0; w < SV_NT / 32; ++w) amax = fmax(amax, sh[w]);
    if (!(amax > 0.0) || !isfinite(amax)) return amax;
    const double inv = 1.0 / amax;
    double ss = 0.0;
    for (long long i = threadIdx.x; i < len; i += SV_NT) { double p = y[i].x * inv, q = y[i].y * inv; ss = fma(p, p, ss); ss = fma(q, q, ss); }
    ss = warp_sum(ss);
    __syncthreads();
    if (lane == 0) sh[warp] = ss;
    __syncthreads();
    ss = 0.0;
    for (int w = 0; w < SV_NT / 32; ++w) ss += sh[w];
    return amax * sqrt(ss);
}

// step 1: ||v|| < 1e-10 -> V_COLLAPSED (AMS:229)
__global__ void __launch_bounds__(SV_NT) svd_check_v_kernel(const cplx* __restrict__ V, long long cols, int* status) {
    __shared__ double sh[SV_NT / 32];
    const int c = blockIdx.x;
    const double nv = block_norm(V + (long long)c * cols, cols, sh);
    if (threadIdx.x == 0) status[c] = (nv < 1e-10) ? MAUS_ST_V_COLLAPSED : MAUS_ST_OK;
}
// step 3: sigma1 = ||A v|| ; u = (A v) / (sigma1 > 1e-10 ? sigma1 : 1) ; ||u|| < 1e-10 -> U collapsed (AMS:233-236)
__global__ void __launch_bounds__(SV_NT) svd_make_u_kernel(const cplx* __restrict__ TU, cplx* __restrict__ U, long long rows,
                                                           double* sig1, int* status) {
    __shared__ double sh[SV_NT / 32];
    const int c = blockIdx.x;
    if (status[c] != 0) return;
    const cplx* y = TU + (long long)c * rows;
    const double s1 = block_norm(y, rows, sh);
    const double den = (s1 > 1e-10) ? s1 : 1.0;
    for (long long i = threadIdx.x; i < rows; i += SV_NT) U[(long long)c * rows + i] = cmake(y[i].x / den, y[i].y / den);
    if (threadIdx.x == 0) { sig1[c] = s1; if (s1 / den < 1e-10) status[c] = MAUS_ST_MIX_COLLAPSED; }
}
// step 5: n2 = ||A^H u|| ; sigma = max(sigma1, n2) ; v = (A^H u) / (n2 > 1e-10 ? n2 : 1)   (AMS:240-242)
__global__ void __launch_bounds__(SV_NT) svd_make_v_kernel(const cplx* __restrict__ TV, cplx* __restrict__ V, long long cols,
                                                           const double* __restrict__ sig1, double* sigma, const int* __restrict__ status) {
    __shared__ double sh[SV_NT / 32];
    const int c = blockIdx.x;
    if (status[c] != 0) return;
    const cplx* y = TV + (long long)c * cols;
    const double n2 = block_norm(y, cols, sh);
    const double den = (n2 > 1e-10) ? n2 : 1.0;
    for (long long i = threadIdx.x; i < cols; i += SV_NT) V[(long long)c * cols + i] = cmake(y[i].x / den, y[i].y / den);
    if (threadIdx.x == 0) sigma[c] = fmax(sig1[c], n2);
}
// residual = ||A v - sigma u|| + ||A^H u - sigma v||  (AMS:301); AV = A v (rows), AHU = A^H u (cols)
__global__ void __launch_bounds__(SV_NT) svd_residual_kernel(cplx* __restrict__ AV, const cplx* __restrict__ U, long long rows,
                                                             cplx* __restrict__ AHU, const cplx* __restrict__ V, long long cols,
                                                             const double* __restrict__ sigma, double* resid,
                                                             const int* __restrict__ status) {
    __shared__ double sh[SV_NT / 32];
    const int c = blockIdx.x;
    if (status && status[c] != 0) return;
    const double s = sigma[c];
    cplx* a = AV + (long long)c * rows; const cplx* u = U + (long long)c * rows;
    for (long long i = threadIdx.x; i < rows; i += SV_NT) a[i] = cmake(a[i].x - s * u[i].x, a[i].y - s * u[i].y);
    cplx* b = AHU + (long long)c * cols; const cplx* v = V + (long long)c * cols;
    for (long long i = threadIdx.x; i < cols; i += SV_NT) b[i] = cmake(b[i].x - s * v[i].x, b[i].y - s * v[i].y);
    __syncthreads();
    const double r1 = block_norm(a, rows, sh);
    const double r2 = block_norm(b, cols, sh);
    if (threadIdx.x == 0) resid[c] = r1 + r2;
}
}  // namespace

void maus_svd_free(maus_ctx* ctx) {
    SvdWs* w = (SvdWs*)ctx->svd;
    if (!w) return;
    cudaFree(w->A_rm); cudaFree(w->A_cm); cudaFree(w->AH_rm); cudaFree(w->AH_cm);
    cudaFree(w->U); cudaFree(w->V); cudaFree(w->TU); cudaFree(w->TV);
    cudaFree(w->sig1); cudaFree(w->sigma); cudaFree(w->resid); cudaFree(w->status);
    ctx->bytes_held -= (long long)w->bytes;
    delete w;
    ctx->svd = nullptr;
}

static int svd_ensure_cands(maus_ctx* ctx, SvdWs* w, long long C) {
    if (C <= w->Ccap) return MAUS_OK;
    MAUS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(w->U); cudaFree(w->V); cudaFree(w->TU); cudaFree(w->TV);
    cudaFree(w->sig1); cudaFree(w->sigma); cudaFree(w->resid); cudaFree(w->status);
    const long long cap = std::max<long long>(C, 8);
    MAUS_CUDA(ctx, cudaMalloc(&w->U, (size_t)cap * w->rows * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMalloc(&w->TU, (size_t)cap * w->rows * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMalloc(&w->V, (size_t)cap * w->cols * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMalloc(&w->TV, (size_t)cap * w->cols * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMalloc(&w->sig1, (size_t)cap * 8)); MAUS_CUDA(ctx, cudaMalloc(&w->sigma, (size_t)cap * 8));
    MAUS_CUDA(ctx, cudaMalloc(&w->resid, (size_t)cap * 8)); MAUS_CUDA(ctx, cudaMalloc(&w->status, (size_t)cap * 4));
    w->Ccap = cap;
    return MAUS_OK;
}

extern "C" int maus_svd_set_matrix(maus_ctx* ctx, int64_t rows, int64_t cols, const double* A_rowmajor) {
    if (!ctx || !A_rowmajor || rows <= 0 || cols <= 0 || rows > 0x7fffffffLL || cols > 0x7fffffffLL)
        return maus_fail(ctx, MAUS_E_ARG, "maus_svd_set_matrix: bad argument");
    cudaSetDevice(ctx->device);
    MAUS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    maus_svd_free(ctx);
    SvdWs* w = new SvdWs();
    ctx->svd = w;
    w->rows = rows; w->cols = cols;
    const size_t bytes = (size_t)rows * cols * sizeof(cplx);
    MAUS_CUDA(ctx, cudaMalloc(&w->A_rm, bytes)); MAUS_CUDA(ctx, cudaMalloc(&w->A_cm, bytes));
    MAUS_CUDA(ctx, cudaMalloc(&w->AH_rm, bytes)); MAUS_CUDA(ctx, cudaMalloc(&w->AH_cm, bytes));
    w->bytes = 4 * bytes; ctx->bytes_held += (long long)w->bytes;
    MAUS_CUDA(ctx, cudaMemcpyAsync(w->A_rm, A_rowmajor, bytes, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 32);
    svd_relayout_kernel<<<grid, block, 0, ctx->stream>>>(w->A_rm, (int)rows, (int)cols, w->A_cm, w->AH_rm, w->AH_cm);
    ctx->launches += 1;
    MAUS_CUDA(ctx, cudaGetLastError());
    MAUS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MAUS_OK;
}

// Y[c] (orows) = Op * X[c] (ocols), Op given row-major (GEMV) and column-major (GEMM)
static int svd_apply(maus_ctx* ctx, const cplx* Op_rm, const cplx* Op_cm, long long orows, long long ocols, const cplx* X, cplx* Y,
                     long long C) {
    if (C <= 8) {
        int h = prof_begin(ctx, MAUS_PROF_MATVEC, (double)((C + 3) / 4) * 16.0 * orows * ocols + 16.0 * (orows + ocols) * C);
        MAUS_CUDA(ctx, vec_gemv_rect(Op_rm, (int)orows, (int)ocols, X, ocols, Y, orows, (int)C, ctx->stream));
        prof_end(ctx, h);
        ctx->launches += (C + 3) / 4;
    } else {
        ZgemmParams p = {};
        p.A = Op_cm; p.lda = orows; p.strideA = 0;
        p.B = X; p.ldb = ocols; p.strideB = 0;
        p.C = Y; p.ldc = orows; p.strideC = 0;
        p.M = (int)orows; p.N = (int)C; p.K = (int)ocols; p.batch = 1; p.beta = 0; p.negate = 0;
        int h = prof_begin(ctx, MAUS_PROF_MATVEC_GEMM, 8.0 * orows * (double)ocols * C);
        MAUS_CUDA(ctx, zgemm_dmma_launch(p, ctx->stream));
        prof_end(ctx, h);
        ctx->launches += 1;
    }
    return MAUS_OK;
}

extern "C" int maus_svd_step(maus_ctx* ctx, int64_t C, double* U_io, double* V_io, double* sigma_out, double* resid_out,
                             int32_t* status_out) {
    if (!ctx || C <= 0 || !U_io || !V_io) return maus_fail(ctx, MAUS_E_ARG, "maus_svd_step: bad argument");
    SvdWs* w = (SvdWs*)ctx->svd;
    if (!w) return maus_fail(ctx, MAUS_E_STATE, "maus_svd_step: maus_svd_set_matrix first");
    cudaSetDevice(ctx->device);
    int rc = svd_ensure_cands(ctx, w, C); if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const long long rows = w->rows, cols = w->cols;
    MAUS_CUDA(ctx, cudaMemcpyAsync(w->U, U_io, (size_t)C * rows * sizeof(cplx), cudaMemcpyHostToDevice, st));
    MAUS_CUDA(ctx, cudaMemcpyAsync(w->V, V_io, (size_t)C * cols * sizeof(cplx), cudaMemcpyHostToDevice, st));
    svd_check_v_kernel<<<(unsigned)C, SV_NT, 0, st>>>(w->V, cols, w->status);
    if ((rc = svd_apply(ctx, w->A_rm, w->A_cm, rows, cols, w->V, w->TU, C))) return rc;             // temp_u = A v
    svd_make_u_kernel<<<(unsigned)C, SV_NT, 0, st>>>(w->TU, w->U, rows, w->sig1, w->status);
    if ((rc = svd_apply(ctx, w->AH_rm, w->AH_cm, cols, rows, w->U, w->TV, C))) return rc;           // temp_v = A^H u
    svd_make_v_kernel<<<(unsigned)C, SV_NT, 0, st>>>(w->TV, w->V, cols, w->sig1, w->sigma, w->status);
    if ((rc = svd_apply(ctx, w->A_rm, w->A_cm, rows, cols, w->V, w->TU, C))) return rc;             // A v_new
    svd_residual_kernel<<<(unsigned)C, SV_NT, 0, st>>>(w->TU, w->U, rows, w->TV, w->V, cols, w->sigma, w->resid, w->status);
    ctx->launches += 4;
    MAUS_CUDA(ctx, cudaGetLastError());
    MAUS_CUDA(ctx, cudaMemcpyAsync(U_io, w->U, (size_t)C * rows * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaMemcpyAsync(V_io, w->V, (size_t)C * cols * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    if (sigma_out) MAUS_CUDA(ctx, cudaMemcpyAsync(sigma_out, w->sigma, (size_t)C * 8, cudaMemcpyDeviceToHost, st));
    if (resid_out) MAUS_CUDA(ctx, cudaMemcpyAsync(resid_out, w->resid, (size_t)C * 8, cudaMemcpyDeviceToHost, st));
    if (status_out) MAUS_CUDA(ctx, cudaMemcpyAsync(status_out, w->status, (size_t)C * 4, cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaStreamSynchronize(st));
    return MAUS_OK;
}

extern "C" int maus_svd_residual(maus_ctx* ctx, int64_t C, const double* U, const double* V, const double* sigma, double* resid_out) {
    if (!ctx || C <= 0 || !U || !V || !sigma || !resid_out) return maus_fail(ctx, MAUS_E_ARG, "maus_svd_residual: bad argument");
    SvdWs* w = (SvdWs*)ctx->svd;
    if (!w) return maus_fail(ctx, MAUS_E_STATE, "maus_svd_residual: maus_svd_set_matrix first");
    cudaSetDevice(ctx->device);
    int rc = svd_ensure_cands(ctx, w, C); if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const long long rows = w->rows, cols = w->cols;
    MAUS_CUDA(ctx, cudaMemcpyAsync(w->U, U, (size_t)C * rows * sizeof(cplx), cudaMemcpyHostToDevice, st));
    MAUS_CUDA(ctx, cudaMemcpyAsync(w->V, V, (size_t)C * cols * sizeof(cplx), cudaMemcpyHostToDevice, st));
    MAUS_CUDA(ctx, cudaMemcpyAsync(w->sigma, sigma, (size_t)C * 8, cudaMemcpyHostToDevice, st));
    if ((rc = svd_apply(ctx, w->A_rm, w->A_cm, rows, cols, w->V, w->TU, C))) return rc;
    if ((rc = svd_apply(ctx, w->AH_rm, w->AH_cm, cols, rows, w->U, w->TV, C))) return rc;
    svd_residual_kernel<<<(unsigned)C, SV_NT, 0, st>>>(w->TU, w->U, rows, w->TV, w->V, cols, w->sigma, w->resid, nullptr);
    ctx->launches += 1;
    MAUS_CUDA(ctx, cudaGetLastError());
    MAUS_CUDA(ctx, cudaMemcpyAsync(resid_out, w->resid, (size_t)C * 8, cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaStreamSynchronize(st));
    return MAUS_OK;
}
