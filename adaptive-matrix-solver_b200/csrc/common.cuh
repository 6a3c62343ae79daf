// common.cuh -- shared device helpers for libmaus_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

typedef double2 cplx;   // complex128 = (re, im), numpy layout

#define MAUS_SM_COUNT_B200 148

__host__ __device__ __forceinline__ cplx cmake(double r, double i) { cplx z; z.x = r; z.y = i; return z; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return cmake(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// a += b*c
__device__ __forceinline__ void cfma(cplx& a, cplx b, cplx c) {
    a.x = fma(b.x, c.x, a.x); a.x = fma(-b.y, c.y, a.x);
    a.y = fma(b.x, c.y, a.y); a.y = fma(b.y, c.x, a.y);
}
// a -= b*c
__device__ __forceinline__ void cfms(cplx& a, cplx b, cplx c) {
    a.x = fma(-b.x, c.x, a.x); a.x = fma(b.y, c.y, a.x);
    a.y = fma(-b.x, c.y, a.y); a.y = fma(-b.y, c.x, a.y);
}
// a += conj(b)*c
__device__ __forceinline__ void cfma_conj(cplx& a, cplx b, cplx c) {
    a.x = fma(b.x, c.x, a.x); a.x = fma(b.y, c.y, a.x);
    a.y = fma(b.x, c.y, a.y); a.y = fma(-b.y, c.x, a.y);
}
__device__ __forceinline__ cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
__device__ __forceinline__ double cabs1(cplx a) { return fabs(a.x) + fabs(a.y); }   // BLAS izamax metric
__device__ __forceinline__ double cabs2(cplx a) { return fma(a.x, a.x, a.y * a.y); }
// robust complex reciprocal (Smith)
__device__ __forceinline__ cplx crecip(cplx a) {
    if (fabs(a.x) >= fabs(a.y)) {
        double r = a.y / a.x, d = a.x + a.y * r;
        return cmake(1.0 / d, -r / d);
    } else {
        double r = a.x / a.y, d = a.x * r + a.y;
        return cmake(r / d, -1.0 / d);
    }
}
// robust complex division a / b (Smith)
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    if (fabs(b.x) >= fabs(b.y)) {
        double r = b.y / b.x, d = b.x + b.y * r;
        return cmake((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    } else {
        double r = b.x / b.y, d = b.x * r + b.y;
        return cmake((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
}
__device__ __forceinline__ bool cfinite(cplx a) { return isfinite(a.x) && isfinite(a.y); }

// ---- Philox4x32-10 counter RNG (Salmon et al. 2011) for the dense Psi perturbation of AMS:49 -----------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// two uniforms in [0,1) with 53 random bits each (same construction as MT19937's genrand_res53)
__device__ __forceinline__ void philox_uniform2(uint64_t key, uint32_t i, uint32_t j, double& u1, double& u2) {
    uint32_t o[4];
    philox4x32_10(i, j, 0x4d415553u /* "MAUS" */, 0u, (uint32_t)key, (uint32_t)(key >> 32), o);
    u1 = ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) * (1.0 / 9007199254740992.0);
    u2 = ((double)(o[2] >> 5) * 67108864.0 + (double)(o[3] >> 6)) * (1.0 / 9007199254740992.0);
}
// R_ij of AMS:49-50 without the psi*I term:  0.15 * psi * ((U1 - 0.5) + i (U2 - 0.5))
__device__ __forceinline__ cplx psi_perturbation(uint64_t key, uint32_t i, uint32_t j, double psi) {
    double u1, u2;
    philox_uniform2(key, i, j, u1, u2);
    double s = psi * 0.15;
    return cmake((u1 - 0.5) * s, (u2 - 0.5) * s);
}

// ---- warp / block reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ cplx warp_sum(cplx v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}
