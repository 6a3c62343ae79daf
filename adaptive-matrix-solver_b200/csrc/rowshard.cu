// rowshard.cu -- row-sharded sparse operator for BASELINE config #5 ("sparse CSC n = 1M ... row-sharded across 8 B200"),
// SURVEY.md section 8e.  Every rank owns a contiguous block of rows of A (CSR slice, global column indices) and the
// matching slice of every candidate vector.  A matvec all-gathers its input vector over NVLink (NCCL), the batched
// GMRES (gmres.cu, same code as the replicated path) all-reduces each dot product / norm.  As SURVEY.md predicts this
// regime is communication-bound and slower than replicating the 0.44 GB matrix and sharding the candidates; it exists
// because the configuration names it, and both are reported side by side (profiles/README_r01.md).
//
// NCCL is bound at run time (dlopen of the libnccl the process already uses -- torch's bundled copy) so that
// libmaus_b200.so has no link-time dependency on it.
#include <dlfcn.h>
#include <cstring>
#include <nccl.h>
#include <vector>
#include <algorithm>
#include "ctx.cuh"
#include "gmres.cuh"
#include "spmv.cuh"

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_bind(const char* libpath) {
    if (g_nccl.handle) return 0;
    void* h = nullptr;
    if (libpath && libpath[0]) h = dlopen(libpath, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return -1;
#define BIND(field, name) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); if (!g_nccl.field) return -1
    BIND(GetUniqueId, "ncclGetUniqueId"); BIND(CommInitRank, "ncclCommInitRank"); BIND(CommDestroy, "ncclCommDestroy");
    BIND(AllGather, "ncclAllGather"); BIND(AllReduce, "ncclAllReduce"); BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd"); BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    g_nccl.handle = h;
    return 0;
}

struct RowShard {
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    long long n = 0, row0 = 0, nloc = 0, nnz = 0;
    long long* rowptr = nullptr; int* colidx = nullptr; cplx* vals = nullptr; cplx* diag = nullptr;
    double amax = 0.0;
    cplx* xfull = nullptr; long long xcap = 0;      // [C][n] gathered input of the matvec
    cplx* pack = nullptr;                           // [n][4] interleaved copy for the SpMM gathers
    cplx *V = nullptr, *X = nullptr, *Y = nullptr, *sigma = nullptr; double* psi = nullptr; unsigned char* jac = nullptr;
    int *status = nullptr, *iters = nullptr; long long Ccap = 0;
};

#define MAUS_NCCL(ctx, call)                                                                             \
    do {                                                                                                 \
        ncclResult_t _r = (call);                                                                        \
        if (_r != ncclSuccess) return maus_fail(ctx, MAUS_E_CUDA, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : #call); \
    } while (0)

void maus_rowshard_free(maus_ctx* ctx) {
    RowShard* rs = (RowShard*)ctx->rowshard;
    if (!rs) return;
    cudaFree(rs->rowptr); cudaFree(rs->colidx); cudaFree(rs->vals); cudaFree(rs->diag); cudaFree(rs->xfull); cudaFree(rs->pack);
    cudaFree(rs->V); cudaFree(rs->X); cudaFree(rs->Y); cudaFree(rs->sigma); cudaFree(rs->psi); cudaFree(rs->jac);
    cudaFree(rs->status); cudaFree(rs->iters);
    if (rs->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(rs->comm);
    delete rs;
    ctx->rowshard = nullptr;
}

extern "C" int maus_nccl_unique_id(const char* libpath, char* out128) {
    if (!out128 || nccl_bind(libpath)) return MAUS_E_ARG;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return MAUS_E_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return MAUS_OK;
}

extern "C" int maus_dist_init(maus_ctx* ctx, const char* libpath, int rank, int world, const char* id128) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return maus_fail(ctx, MAUS_E_ARG, "maus_dist_init: bad argument");
    if (nccl_bind(libpath)) return maus_fail(ctx, MAUS_E_STATE, "maus_dist_init: cannot bind libnccl.so.2");
    cudaSetDevice(ctx->device);
    maus_rowshard_free(ctx);
    RowShard* rs = new RowShard();
    ctx->rowshard = rs;
    rs->rank = rank; rs->world = world;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    MAUS_NCCL(ctx, g_nccl.CommInitRank(&rs->comm, world, id, rank));
    return MAUS_OK;
}

extern "C" int maus_set_csr_rowblock(maus_ctx* ctx, int64_t n, int64_t row0, int64_t nrows, const int64_t* rowptr,
                                     const int64_t* colidx, const double* vals) {
    RowShard* rs = ctx ? (RowShard*)ctx->rowshard : nullptr;
    if (!rs) return maus_fail(ctx, MAUS_E_STATE, "maus_set_csr_rowblock: maus_dist_init first");
    if (!rowptr || n <= 0 || nrows <= 0 || row0 < 0 || row0 + nrows > n || n > 0x7fffffffLL)
        return maus_fail(ctx, MAUS_E_ARG, "maus_set_csr_rowblock: bad argument");
    if (nrows * rs->world != n || row0 != nrows * rs->rank)
        return maus_fail(ctx, MAUS_E_ARG, "maus_set_csr_rowblock: equal row blocks required (n % world == 0, row0 = rank * n / world)");
    cudaSetDevice(ctx->device);
    const long long nnz = rowptr[nrows] - rowptr[0];
    std::vector<long long> rp((size_t)nrows + 1);
    for (long long i = 0; i <= nrows; ++i) rp[(size_t)i] = rowptr[i] - rowptr[0];
    std::vector<int> ci((size_t)nnz);
    std::vector<cplx> dg((size_t)nrows, cmake(0.0, 0.0));
    double amax = 0.0;
    for (long long i = 0; i < nrows; ++i)
        for (long long k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const long long j = colidx[k];
            if (j < 0 || j >= n) return maus_fail(ctx, MAUS_E_ARG, "maus_set_csr_rowblock: column index out of range");
            ci[(size_t)(k - rowptr[0])] = (int)j;
            const cplx z = cmake(vals[2 * k], vals[2 * k + 1]);
            if (j == row0 + i) { dg[(size_t)i].x += z.x; dg[(size_t)i].y += z.y; }
            amax = std::max(amax, std::fabs(z.x) + std::fabs(z.y));
        }
    cudaFree(rs->rowptr); cudaFree(rs->colidx); cudaFree(rs->vals); cudaFree(rs->diag);
    MAUS_CUDA(ctx, cudaMalloc(&rs->rowptr, (size_t)(nrows + 1) * 8));
    MAUS_CUDA(ctx, cudaMalloc(&rs->colidx, std::max<size_t>((size_t)nnz * 4, 16)));
    MAUS_CUDA(ctx, cudaMalloc(&rs->vals, std::max<size_t>((size_t)nnz * sizeof(cplx), 16)));
    MAUS_CUDA(ctx, cudaMalloc(&rs->diag, (size_t)nrows * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMemcpy(rs->rowptr, rp.data(), (size_t)(nrows + 1) * 8, cudaMemcpyHostToDevice));
    if (nnz) {
        MAUS_CUDA(ctx, cudaMemcpy(rs->colidx, ci.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice));
        MAUS_CUDA(ctx, cudaMemcpy(rs->vals, vals + 2 * rowptr[0], (size_t)nnz * sizeof(cplx), cudaMemcpyHostToDevice));
    }
    MAUS_CUDA(ctx, cudaMemcpy(rs->diag, dg.data(), (size_t)nrows * sizeof(cplx), cudaMemcpyHostToDevice));
    rs->n = n; rs->row0 = row0; rs->nloc = nrows; rs->nnz = nnz; rs->amax = amax;
    return MAUS_OK;
}

static int rs_ensure(maus_ctx* ctx, RowShard* rs, long long C) {
    if (C <= rs->Ccap) return MAUS_OK;
    MAUS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(rs->xfull); cudaFree(rs->pack); cudaFree(rs->V); cudaFree(rs->X); cudaFree(rs->Y); cudaFree(rs->sigma); cudaFree(rs->psi);
    cudaFree(rs->jac); cudaFree(rs->status); cudaFree(rs->iters);
    // a failing cudaMalloc below returns early: no pointer may stay dangling for maus_rowshard_free
    rs->xfull = rs->pack = rs->V = rs->X = rs->Y = rs->sigma = nullptr; rs->psi = nullptr; rs->jac = nullptr;
    rs->status = rs->iters = nullptr; rs->Ccap = 0;
    const long long cap = std::max<long long>(C, 4);
    MAUS_CUDA(ctx, cudaMalloc(&rs->xfull, (size_t)cap * rs->n * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMalloc(&rs->pack, (size_t)4 * rs->n * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMalloc(&rs->V, (size_t)cap * rs->nloc * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMalloc(&rs->X, (size_t)cap * rs->nloc * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMalloc(&rs->Y, (size_t)cap * rs->nloc * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMalloc(&rs->sigma, (size_t)cap * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMalloc(&rs->psi, (size_t)cap * 8));
    MAUS_CUDA(ctx, cudaMalloc(&rs->jac, (size_t)cap));
    MAUS_CUDA(ctx, cudaMalloc(&rs->status, (size_t)cap * 4));
    MAUS_CUDA(ctx, cudaMalloc(&rs->iters, (size_t)cap * 4));
    rs->Ccap = cap;
    return MAUS_OK;
}

// z[c] (local rows) = A_local * allgather(v[c]); one NCCL all-gather per candidate vector, grouped
static int rs_matvec(maus_ctx* ctx, RowShard* rs, const cplx* v, long long ldv, cplx* z, long long ldz, long long C) {
    cudaStream_t st = ctx->stream;
    int h = prof_begin(ctx, MAUS_PROF_MATVEC, (double)((C + 3) / 4) * (20.0 * rs->nnz + 8.0 * (rs->nloc + 1)) + 16.0 * (rs->n + rs->nloc) * C);
    MAUS_NCCL(ctx, g_nccl.GroupStart());
    for (long long c = 0; c < C; ++c)
        MAUS_NCCL(ctx, g_nccl.AllGather(v + c * ldv, rs->xfull + c * rs->n, (size_t)rs->nloc * 2, ncclDouble, rs->comm, st));
    MAUS_NCCL(ctx, g_nccl.GroupEnd());
    MAUS_CUDA(ctx, csr_spmm(rs->rowptr, rs->colidx, rs->vals, rs->xfull, rs->n, z, ldz, rs->nloc, rs->n, (int)C, rs->pack, st));
    prof_end(ctx, h);
    ctx->launches += (C + 3) / 4;
    return MAUS_OK;
}

extern "C" int maus_rs_matvec(maus_ctx* ctx, int64_t C, const double* V_local, double* Y_local) {
    RowShard* rs = ctx ? (RowShard*)ctx->rowshard : nullptr;
    if (!rs || !rs->rowptr) return maus_fail(ctx, MAUS_E_STATE, "maus_rs_matvec: row block not set");
    if (C <= 0 || !V_local || !Y_local) return maus_fail(ctx, MAUS_E_ARG, "maus_rs_matvec: bad argument");
    cudaSetDevice(ctx->device);
    int rc = rs_ensure(ctx, rs, C); if (rc) return rc;
    cudaStream_t st = ctx->stream;
    MAUS_CUDA(ctx, cudaMemcpyAsync(rs->V, V_local, (size_t)C * rs->nloc * sizeof(cplx), cudaMemcpyHostToDevice, st));
    if ((rc = rs_matvec(ctx, rs, rs->V, rs->nloc, rs->Y, rs->nloc, C))) return rc;
    MAUS_CUDA(ctx, cudaMemcpyAsync(Y_local, rs->Y, (size_t)C * rs->nloc * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaStreamSynchronize(st));
    return MAUS_OK;
}

// x_c = (A - sigma_c I + psi_c I)^-1 rhs_c by the batched GMRES on the row-sharded operator; RHS / X are local slices
extern "C" int maus_rs_gmres(maus_ctx* ctx, int64_t C, const double* sigma, const double* psi, const uint8_t* use_jacobi,
                             const double* RHS_local, double* X_local_out, int32_t* status_out, int32_t* iters_out) {
    RowShard* rs = ctx ? (RowShard*)ctx->rowshard : nullptr;
    if (!rs || !rs->rowptr) return maus_fail(ctx, MAUS_E_STATE, "maus_rs_gmres: row block not set");
    if (C <= 0 || !sigma || !psi || !RHS_local) return maus_fail(ctx, MAUS_E_ARG, "maus_rs_gmres: bad argument");
    cudaSetDevice(ctx->device);
    int rc = rs_ensure(ctx, rs, C); if (rc) return rc;
    cudaStream_t st = ctx->stream;
    MAUS_CUDA(ctx, cudaMemcpyAsync(rs->V, RHS_local, (size_t)C * rs->nloc * sizeof(cplx), cudaMemcpyHostToDevice, st));
    MAUS_CUDA(ctx, cudaMemcpyAsync(rs->sigma, sigma, (size_t)C * sizeof(cplx), cudaMemcpyHostToDevice, st));
    MAUS_CUDA(ctx, cudaMemcpyAsync(rs->psi, psi, (size_t)C * 8, cudaMemcpyHostToDevice, st));
    if (use_jacobi) MAUS_CUDA(ctx, cudaMemcpyAsync(rs->jac, use_jacobi, (size_t)C, cudaMemcpyHostToDevice, st));
    else MAUS_CUDA(ctx, cudaMemsetAsync(rs->jac, 0, (size_t)C, st));
    MAUS_CUDA(ctx, cudaMemsetAsync(rs->status, 0, (size_t)C * 4, st));
    MAUS_CUDA(ctx, cudaMemsetAsync(rs->iters, 0, (size_t)C * 4, st));
    GmresOperator op;
    op.nloc = rs->nloc; op.nglobal = rs->n; op.row0 = rs->row0; op.diag = rs->diag; op.amax = rs->amax; op.dense = false;
    op.matvec = [ctx, rs](const cplx* v, long long ldv, cplx* z, long long ldz, long long Cn) { return rs_matvec(ctx, rs, v, ldv, z, ldz, Cn); };
    op.reduce_sync = [ctx, rs](cplx* red, long long Cn) -> int {
        MAUS_NCCL(ctx, g_nccl.AllReduce(red, red, (size_t)Cn * 2, ncclDouble, ncclSum, rs->comm, ctx->stream));
        return MAUS_OK;
    };
    op.flag_sync = [ctx, rs](int* flags, long long Cn) -> int {
        MAUS_NCCL(ctx, g_nccl.AllReduce(flags, flags, (size_t)Cn, ncclInt32, ncclMax, rs->comm, ctx->stream));
        return MAUS_OK;
    };
    double max_psi = 0.0;
    for (long long c = 0; c < C; ++c) max_psi = std::max(max_psi, std::fabs(psi[c]));
    if ((rc = gmres_core(ctx, op, C, rs->sigma, rs->psi, nullptr, rs->jac, rs->V, rs->nloc, rs->X, rs->status, rs->iters, max_psi)))
        return rc;
    if (X_local_out) MAUS_CUDA(ctx, cudaMemcpyAsync(X_local_out, rs->X, (size_t)C * rs->nloc * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    if (status_out) MAUS_CUDA(ctx, cudaMemcpyAsync(status_out, rs->status, (size_t)C * 4, cudaMemcpyDeviceToHost, st));
    if (iters_out) MAUS_CUDA(ctx, cudaMemcpyAsync(iters_out, rs->iters, (size_t)C * 4, cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaStreamSynchronize(st));
    return MAUS_OK;
}
