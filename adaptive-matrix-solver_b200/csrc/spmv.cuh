// spmv.cuh -- CSR SpMM: Y[c] = A * V[c] for C candidate vectors ([C][n] layout).
#pragma once
#include "common.cuh"
// n rows (local rows of a row block), ncols = length of the vectors; pack_ws: optional scratch of csr_spmm_pack_elems(ncols, C)
// elements -- with it, the candidates are gathered from interleaved copies [ncols][4] (whole L2 sectors per matrix entry, one
// L1 wavefront per entry and 4 candidates), bit-identical results.  max_row_len: the longest row of the CSR block (0 = unknown):
// rows of at most 24 entries take a leaner instantiation of the packed kernels
cudaError_t csr_spmm(const long long* rowptr, const int* colidx, const cplx* vals, const cplx* V, long long ldv, cplx* Y,
                     long long ldy, long long n, long long ncols, int C, cplx* pack_ws, int max_row_len, cudaStream_t stream);
// the packed kernel alone: P holds `groups` interleaved copies [ncols][4], p_gstride elements apart (row-sharded operator: the
// copies are filled by the peers over NVLink, rowshard.cu); candidate c0 + 4 g + c of ctotal is written to Y[(c0 + 4 g + c) * ldy]
cudaError_t csr_spmm_packed4(const long long* rowptr, const int* colidx, const cplx* vals, const cplx* P, long long p_gstride,
                             cplx* Y, long long ldy, long long n, int c0, int ctotal, int groups, int max_row_len, cudaStream_t stream);
size_t csr_spmm_pack_elems(long long ncols, int C);
