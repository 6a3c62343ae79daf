// gmres.cuh -- operator abstraction of the batched GMRES (gmres.cu): the same solver runs on a replicated matrix
// (one GPU owns whole vectors) or on a row-sharded matrix (rowshard.cu: every rank owns a slice of each vector, the
// matvec all-gathers its input, partial sums are all-reduced with NCCL).
#pragma once
#include <functional>
#include "ctx.cuh"

struct GmresOperator {
    long long nloc = 0;        // local vector length (= n when replicated)
    long long nglobal = 0;     // global order (restart = min(20, nglobal))
    long long row0 = 0;        // global index of the first local row
    const cplx* diag = nullptr;   // diagonal of the matrix, local rows
    double amax = 0.0;
    bool dense = false;
    // z[c] (ldz apart) = A * v[c] (ldv apart) for C vectors; v / z are LOCAL slices in row-sharded mode
    std::function<int(const cplx* v, long long ldv, cplx* z, long long ldz, long long C)> matvec;
    // row-sharded mode only: partial [C][maxblk] complex holds nblk per-block partial sums per candidate; replace block 0 by the
    // sum over blocks AND ranks (same value, same rounding on every rank), zero the other blocks
    std::function<int(cplx* partial, int maxblk, int nblk, long long C)> reduce_partials;
    // row-sharded mode only: max-combine C device ints over the ranks, in place
    std::function<int(int* flags, long long C)> flag_sync;
};

int gmres_core(maus_ctx* ctx, const GmresOperator& op, long long C, const cplx* sigma, const double* psi,
               const unsigned long long* keys, const unsigned char* use_jacobi, const cplx* rhs, long long rhs_stride, cplx* X,
               int* status, int* iters, double max_psi_host);
