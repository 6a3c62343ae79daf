// ctx.cuh -- the opaque context behind the C ABI (include/maus_b200.h) and internal helpers shared by the .cu files.
#pragma once
#include <string>
#include <vector>
#include "common.cuh"
#include "lu.cuh"
#include "zgemm.cuh"
#include "../../include/maus_b200.h"

struct MatrixSlot {
    bool dense = false, sparse = false;
    cplx* rm = nullptr;          // n x n row-major (as uploaded)   -- matvec
    cplx* cm = nullptr;          // n x n column-major               -- LU / GEMM
    // CSR (converted from the uploaded CSC)
    long long nnz = 0;
    int max_row = 0;             // longest CSR row (0 = unknown): selects the single-chunk SpMM kernels (spmv.cu)
    long long* rowptr = nullptr; // n+1
    int* colidx = nullptr;       // nnz
    cplx* vals = nullptr;        // nnz
    cplx* diag = nullptr;        // n, diagonal of the matrix (Jacobi preconditioner, AMS:67)
    cplx* pack = nullptr;        // groups of [n][4] interleaved candidate vectors for the SpMM gathers (spmv.cu)
    size_t pack_elems = 0;
    double amax = 0.0;           // max |a_ij| (cabs1), used to gate the sub-ulp Psi perturbation in matvec-only paths
};

struct ProfAccum {
    bool enabled = false;
    std::vector<cudaEvent_t> ev;     // pairs
    std::vector<int> kind;           // MAUS_PROF_* (include/maus_b200.h)
    std::vector<long long> tag;      // optional per-launch tag (GEMM shape), printed when MAUS_GEMM_LOG is set
    size_t used = 0;
    double ms[MAUS_PROF_KINDS] = {0};
    long long launches[MAUS_PROF_KINDS] = {0};
    double work[MAUS_PROF_KINDS] = {0};            // flops (GEMM kinds) / bytes (HBM kinds)
};

struct maus_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    long long n = 0;
    MatrixSlot slot[2];
    bool slot1_set = false;
    cplx* b = nullptr;
    bool b_set = false;

    // resident population (capacity Ccap)
    long long Ccap = 0;
    cplx *V = nullptr, *X = nullptr, *Y = nullptr;
    cplx *lambda = nullptr, *sigma = nullptr;
    double *psi = nullptr, *alpha = nullptr, *vnorm2 = nullptr, *resid = nullptr, *mixnorm = nullptr;
    double* vscratch = nullptr;    // per-block partials of the multi-block vector reductions (vec.cu, long vectors)
    unsigned long long* keys = nullptr;
    int *status = nullptr, *iters = nullptr, *info = nullptr;
    unsigned char *skip = nullptr, *jac = nullptr;

    // LU workspace
    cplx* W = nullptr;
    long long Wbytes = 0;
    int Wbatch = 0;
    LuPairs* pairs = nullptr;
    cplx* Linv = nullptr;
    cplx* Rcm = nullptr;             // debug: host-supplied perturbation (column-major)
    long long ws_limit = 0;
    bool lu_conj_transpose = false;  // maus_lu_solve factors A^H instead of A (condition estimate, diag.cu)

    // GMRES workspace (gmres.cu)
    void* gmres = nullptr;
    // SVD branch (svd.cu): rectangular matrix in four layouts + candidate buffers
    void* svd = nullptr;
    // row-sharded sparse operator + NCCL communicator (rowshard.cu)
    void* rowshard = nullptr;
    // Hermitian eigensolver workspace (heev.cu)
    void* heev = nullptr;

    long long launches = 0;
    long long bytes_held = 0;
    ProfAccum prof;
};

int maus_fail(maus_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess);
#define MAUS_CUDA(ctx, call)                                                     \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess) return maus_fail(ctx, MAUS_E_CUDA, #call, _e);    \
    } while (0)

cudaError_t maus_dev_alloc(maus_ctx* ctx, void** p, size_t bytes);
void maus_dev_free(maus_ctx* ctx, void* p, size_t bytes);
int maus_ensure_population(maus_ctx* ctx, long long C);

// NVTX ranges (kernel families; no-ops without an attached tool)
void maus_nvtx_push(const char* name);
void maus_nvtx_pop();
struct MausNvtxRange {
    explicit MausNvtxRange(const char* name) { maus_nvtx_push(name); }
    ~MausNvtxRange() { maus_nvtx_pop(); }
};

// profiling brackets (event timing is a no-op unless enabled; each pair is also an NVTX range)
int prof_begin(maus_ctx* ctx, int kind, double work);
void prof_end(maus_ctx* ctx, int handle);
void prof_tag(maus_ctx* ctx, int handle, int M, int N, int K, int batch);

// Y[c] = A(slot) * V[c] for C candidates: dense -> DMMA GEMM (C > 8) or HBM-bound GEMV; sparse -> CSR SpMM
int maus_apply_matrix(maus_ctx* ctx, int slot, const cplx* V, long long ldv, cplx* Y, long long ldy, long long C);

// batched LU solve of C systems (chunked to the workspace): X[c] = (A - sigma_c I + psi_c I + R_c)^-1 rhs_c
int maus_lu_solve(maus_ctx* ctx, long long C, const cplx* sigma, const double* psi, const unsigned long long* keys,
                  const cplx* Rcm, const cplx* rhs, long long rhs_stride, cplx* X, int* status);

// batched GMRES (gmres.cu)
int maus_gmres_solve(maus_ctx* ctx, long long C, const cplx* sigma, const double* psi, const unsigned long long* keys,
                     const unsigned char* use_jacobi, const cplx* rhs, long long rhs_stride, cplx* X, int* status,
                     int* iters, double max_psi_host);
void maus_gmres_free(maus_ctx* ctx);
void maus_svd_free(maus_ctx* ctx);
void maus_rowshard_free(maus_ctx* ctx);
void maus_heev_free(maus_ctx* ctx);
