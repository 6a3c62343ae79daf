// gmres.cu -- batched restarted GMRES (placeholder until the Arnoldi kernels land)
#include "ctx.cuh"
int maus_gmres_solve(maus_ctx* ctx, long long, const cplx*, const double*, const unsigned long long*, const unsigned char*,
                     const cplx*, long long, cplx*, int*, int*) {
    return maus_fail(ctx, MAUS_E_ARG, "GMRES path not built");
}
void maus_gmres_free(maus_ctx*) {}
