// rowshard.cu -- row-sharded sparse operator for BASELINE config #5 ("sparse CSC n = 1M ... row-sharded across 8 B200"),
// SURVEY.md section 8e.  Every rank owns a contiguous block of rows of A (CSR slice, global column indices) and the
// matching slice of every candidate vector; the batched GMRES (gmres.cu, same code as the replicated path) runs on it.
//
// B200 design (NVLink 5 / NVSwitch peer memory, no NCCL call on the solve path):
//   * every rank maps one symmetric device segment of every peer (cudaIpc), holding flags, reduction slots and the
//     gathered, INTERLEAVED matvec input x (groups of [n][4], the layout the SpMM gathers from, spmv.cu);
//   * all-gather fused into its producer: rs_push_pack_kernel reads the local slices of the C vectors once, interleaves them in
//     shared memory and STORES the 4 KB pieces straight into every peer's copy over NVLink (one kernel = pack + all-gather);
//     the last CTA releases a sequence flag on every peer, the consumer side spins on its LOCAL flags;
//   * dot products / norms: ONE kernel per reduction (rs_allreduce_kernel) collapses the per-block partials, writes the C
//     values into a slot on every peer, exchanges flags and sums the G contributions in RANK ORDER -- every rank gets
//     bit-identical results (same iteration counts, same control flow), a few microseconds instead of an NCCL launch.
// NCCL (bound at run time from the libnccl the process already uses, no link dependency) only sets the segment up (handle
// exchange), gathers the full vectors for the host write-back once per generation, backs maus_gather, and remains as the
// fallback transport (MAUS_RS_NCCL=1, or when peer access is unavailable).
#include <dlfcn.h>
#include <cstring>
#include <cstdlib>
#include <cstdint>
#include <nccl.h>
#include <vector>
#include <algorithm>
#include "ctx.cuh"
#include "gmres.cuh"
#include "spmv.cuh"
#include "vec.cuh"

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

constexpr int RS_MAXG = 16;            // ranks of one NVSwitch domain
constexpr int RS_RED_SLOTS = 4;        // a rank is at most one reduction ahead of the slowest peer; 4 slots leave margin
constexpr int RS_RED_COMP = 4;         // doubles per candidate and reduction
constexpr int RS_PUSH_ROWS = 64;       // rows per CTA of the push kernel: 64 x 4 x 16 B = one 4 KB piece per destination
constexpr long long RS_SPIN_LIMIT = 40000000000LL;  // ~20 s of SM clocks: a lost peer sets the error flag instead of hanging the GPU

// flag words at the head of the segment (unsigned long long each)
constexpr int RS_F_XREADY = 0;                          // [2][RS_MAXG]  push of buffer b by rank r complete (sequence number)
constexpr int RS_F_XDONE = 2 * RS_MAXG;                 // [RS_MAXG]     rank r finished reading its copy for matvec seq
constexpr int RS_F_RED = 3 * RS_MAXG;                   // [RS_RED_SLOTS][RS_MAXG]
constexpr int RS_F_WORDS = (3 + RS_RED_SLOTS) * RS_MAXG;
constexpr size_t RS_FLAG_BYTES = 4096;
static_assert(RS_F_WORDS * 8 <= (int)RS_FLAG_BYTES, "flag area");

struct RowShard {
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    long long n = 0, row0 = 0, nloc = 0, nnz = 0;
    int max_row = 0;          // longest row of the local CSR block (selects the single-chunk SpMM kernels)
    long long* rowptr = nullptr; int* colidx = nullptr; cplx* vals = nullptr; cplx* diag = nullptr;
    double amax = 0.0;
    cplx* xfull = nullptr;                          // [C][n]: NCCL transport of the matvec input / full vectors for the host write-back
    cplx* pack = nullptr; size_t pack_elems = 0;    // NCCL transport: interleaved copy of xfull
    cplx *V = nullptr, *X = nullptr, *Y = nullptr, *sigma = nullptr, *lambda = nullptr, *b = nullptr;
    double *psi = nullptr, *alpha = nullptr, *vnorm2 = nullptr, *resid = nullptr, *mixnorm = nullptr, *scratch = nullptr;
    unsigned char* jac = nullptr;
    int *status = nullptr, *iters = nullptr; long long Ccap = 0;
    bool b_set = false;
    // peer-memory transport
    bool p2p = false, p2p_tried = false;
    unsigned char* seg = nullptr; size_t seg_bytes = 0;
    void* peer[RS_MAXG] = {nullptr};
    void** d_peer = nullptr;
    size_t off_red = 0, off_x = 0, xbuf_bytes = 0;
    long long seg_C = 0;                            // candidates the segment was sized for
    unsigned long long seq_x = 0, seq_red = 0;
    unsigned int* d_counter = nullptr;              // last-CTA-done counter of the push kernel
    int* d_err = nullptr;                           // set by a spin loop that hit RS_SPIN_LIMIT
    // gather buffers of maus_gather
    double *gsend = nullptr, *grecv = nullptr; size_t gcap = 0;
};

#define MAUS_NCCL(ctx, call)                                                                             \
    do {                                                                                                 \
        ncclResult_t _r = (call);                                                                        \
        if (_r != ncclSuccess) return maus_fail(ctx, MAUS_E_CUDA, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : #call); \
    } while (0)

// ------------------------------------------------------------------------------------------------------------------
// device side of the peer-memory transport
// ------------------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// spin until *flag >= want; gives up (and latches the error flag) after RS_SPIN_LIMIT clocks or when an earlier wait failed
__device__ __forceinline__ void spin_until(const unsigned long long* flag, unsigned long long want, int* err) {
    if (ld_acquire_sys(flag) >= want) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < want) {
        if (*(volatile int*)err) return;
        if (clock64() - t0 > RS_SPIN_LIMIT) { atomicExch(err, 1); return; }
        __nanosleep(64);
    }
}
__device__ __forceinline__ unsigned long long* seg_flags(void* base) { return reinterpret_cast<unsigned long long*>(base); }

// pack + all-gather in one kernel: the local slices of the C vectors -> interleaved [group][row0 + i][4] in EVERY rank's segment.
// A bounded grid (two CTAs per SM) walks (row tile, group) items: 64 rows x 4 candidates are read coalesced, interleaved in
// shared memory in destination order, and ONE thread per destination hands the 4 KB piece to the TMA (cp.async.bulk shared ->
// global, peer address): full-size NVLink writes issued by the copy engine of the SM instead of 16-byte stores by every thread.
// (Measured first version, per-thread stores: ~530 GB/s of NVLink egress per rank at 8 GPUs.)  Overlapping the pushes of later
// groups with the SpMM of earlier ones on a second stream was measured too (commit 9124dc1): 109.7 vs 107.0 ms per 64 solves --
// both sides want the SMs' memory pipes, nothing is hidden -- so the matvec stays push -> wait -> SpMM -> done on one stream.
__global__ void __launch_bounds__(256) rs_push_pack_kernel(const cplx* __restrict__ v, long long ldv, int C, long long nloc,
                                                           long long row0, long long n, void* const* __restrict__ peer,
                                                           size_t off_xbuf, int buf, int groups, int rank, int G,
                                                           unsigned long long seq, unsigned int* counter, int* err) {
    __shared__ __align__(128) cplx tile[2][RS_PUSH_ROWS * 4];         // destination order: element (i, c) at i * 4 + c
    __shared__ int is_last;
    const int t = threadIdx.x;
    // the copy being overwritten was last read by matvec seq - 2 (two buffers): every peer must have finished that one
    if (seq > 2 && t < G) spin_until(seg_flags(peer[rank]) + RS_F_XDONE + t, seq - 2, err);
    __syncthreads();
    const long long tiles = (nloc + RS_PUSH_ROWS - 1) / RS_PUSH_ROWS, items = tiles * groups;
    int par = 0;
    for (long long it = blockIdx.x; it < items; it += gridDim.x, par ^= 1) {
        const int g = (int)(it / tiles);
        const long long i0 = (it % tiles) * RS_PUSH_ROWS;
        const int rows = (int)min((long long)RS_PUSH_ROWS, nloc - i0);
        // the bulk stores that read tile[par] two iterations ago must have finished READING it
        if (t < G) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        {
            const int c = t / RS_PUSH_ROWS, i = t % RS_PUSH_ROWS;      // 64 consecutive rows of one candidate: 1 KB coalesced
            const int cand = g * 4 + c;
            tile[par][i * 4 + c] = (cand < C && i < rows) ? v[(long long)cand * ldv + i0 + i] : cmake(0.0, 0.0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk (async proxy) reads
        __syncthreads();
        if (t < G) {
            const int d = (rank + t) % G;                              // every rank starts with its own copy: spreads the NVLink ports
            cplx* dst = reinterpret_cast<cplx*>(static_cast<unsigned char*>(peer[d]) + off_xbuf) + ((long long)g * n + row0 + i0) * 4;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(dst), "r"((uint32_t)__cvta_generic_to_shared(&tile[par][0])), "r"((uint32_t)(rows * 4 * sizeof(cplx))) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (t < G) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");          // all bulk writes of this thread performed
        asm volatile("fence.proxy.async;" ::: "memory");                   // ... and ordered before the generic-proxy flag store below
    }
    __threadfence_system();
    __syncthreads();
    if (t == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (is_last) {
        if (t == 0) *counter = 0;
        __threadfence_system();
        if (t < G) st_release_sys(seg_flags(peer[t]) + RS_F_XREADY + buf * RS_MAXG + rank, seq);
    }
}

// consumer side of the all-gather: wait until every rank's piece of buffer `buf` has arrived
__global__ void rs_wait_x_kernel(void* const* __restrict__ peer, int buf, int rank, int G, unsigned long long seq, int* err) {
    if ((int)threadIdx.x < G) spin_until(seg_flags(peer[rank]) + RS_F_XREADY + buf * RS_MAXG + threadIdx.x, seq, err);
}
// after the SpMM: tell every peer this rank no longer reads the copy of matvec `seq`
__global__ void rs_done_x_kernel(void* const* __restrict__ peer, int rank, int G, unsigned long long seq) {
    if ((int)threadIdx.x < G) st_release_sys(seg_flags(peer[threadIdx.x]) + RS_F_XDONE + rank, seq);
}

// One-kernel all-reduce of per-candidate records over the ranks.  part: [C] records, cand_stride doubles apart, each with nblk
// blocks of blk_stride doubles holding ncomp components; component k is combined with max when bit k of maxmask is set, else
// summed; components whose bit in skipmask is set are not exchanged (written as 0).  neg0_skips: a record whose first value is
// negative marks a skipped candidate (vec_mix_part) and travels as (-1, 0, ...).  Result -> block 0 of every record, neutral
// zeros -> the other blocks (the consumers re-reduce all nblk blocks).  Single CTA.
__global__ void __launch_bounds__(256) rs_allreduce_kernel(double* __restrict__ part, long long cand_stride, int blk_stride, int nblk,
                                                           int ncomp, unsigned maxmask, unsigned skipmask, int neg0_skips, int C,
                                                           void* const* __restrict__ peer, size_t off_red, long long segC,
                                                           int rank, int G, unsigned long long seq, int* err) {
    const int t = threadIdx.x;
    const int slot = (int)(seq % RS_RED_SLOTS);
    for (int c = t; c < C; c += blockDim.x) {
        double v[RS_RED_COMP] = {0.0, 0.0, 0.0, 0.0};
        double* rec = part + (long long)c * cand_stride;
        if (neg0_skips && rec[0] < 0.0) v[0] = -1.0;
        else
            for (int k = 0; k < ncomp; ++k) {
                if ((skipmask >> k) & 1u) continue;
                double a = rec[k];
                for (int q = 1; q < nblk; ++q) {
                    const double bq = rec[(long long)q * blk_stride + k];
                    a = ((maxmask >> k) & 1u) ? fmax(a, bq) : a + bq;
                }
                v[k] = a;
            }
        for (int j = 0; j < G; ++j) {
            const int d = (rank + j) % G;
            double* dst = reinterpret_cast<double*>(static_cast<unsigned char*>(peer[d]) + off_red)
                          + (((long long)slot * RS_MAXG + rank) * segC + c) * RS_RED_COMP;
            dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; dst[3] = v[3];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (t < G) st_release_sys(seg_flags(peer[t]) + RS_F_RED + slot * RS_MAXG + rank, seq);
    if (t < G) spin_until(seg_flags(peer[rank]) + RS_F_RED + slot * RS_MAXG + t, seq, err);
    __syncthreads();
    const double* mine = reinterpret_cast<const double*>(static_cast<unsigned char*>(peer[rank]) + off_red)
                         + (long long)slot * RS_MAXG * segC * RS_RED_COMP;
    for (int c = t; c < C; c += blockDim.x) {
        double* rec = part + (long long)c * cand_stride;
        for (int k = 0; k < ncomp; ++k) {
            double a = 0.0;
            if (!((skipmask >> k) & 1u)) {
                a = ld_volatile_f64(mine + ((long long)0 * segC + c) * RS_RED_COMP + k);
                for (int r = 1; r < G; ++r) {                          // rank order: identical on every rank
                    const double br = ld_volatile_f64(mine + ((long long)r * segC + c) * RS_RED_COMP + k);
                    a = ((maxmask >> k) & 1u) ? fmax(a, br) : a + br;
                }
            }
            rec[k] = a;
            for (int q = 1; q < nblk; ++q) rec[(long long)q * blk_stride + k] = 0.0;
        }
    }
}

// max-combine C ints over the ranks (Jacobi validity, non-finite status): same protocol, values travel as doubles
__global__ void __launch_bounds__(256) rs_allreduce_int_max_kernel(int* __restrict__ flags, int C, void* const* __restrict__ peer,
                                                                   size_t off_red, long long segC, int rank, int G,
                                                                   unsigned long long seq, int* err) {
    const int t = threadIdx.x;
    const int slot = (int)(seq % RS_RED_SLOTS);
    for (int c = t; c < C; c += blockDim.x) {
        const double v = (double)flags[c];
        for (int j = 0; j < G; ++j) {
            const int d = (rank + j) % G;
            double* dst = reinterpret_cast<double*>(static_cast<unsigned char*>(peer[d]) + off_red)
                          + (((long long)slot * RS_MAXG + rank) * segC + c) * RS_RED_COMP;
            dst[0] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (t < G) st_release_sys(seg_flags(peer[t]) + RS_F_RED + slot * RS_MAXG + rank, seq);
    if (t < G) spin_until(seg_flags(peer[rank]) + RS_F_RED + slot * RS_MAXG + t, seq, err);
    __syncthreads();
    const double* mine = reinterpret_cast<const double*>(static_cast<unsigned char*>(peer[rank]) + off_red)
                         + (long long)slot * RS_MAXG * segC * RS_RED_COMP;
    for (int c = t; c < C; c += blockDim.x) {
        double a = ld_volatile_f64(mine + (long long)c * RS_RED_COMP);
        for (int r = 1; r < G; ++r) a = fmax(a, ld_volatile_f64(mine + ((long long)r * segC + c) * RS_RED_COMP));
        flags[c] = (int)a;
    }
}

// NCCL transport: records are collapsed to [C][4] doubles, all-reduced in two calls (sum part / max part), expanded again
__global__ void rs_collapse4_kernel(const double* __restrict__ part, long long cand_stride, int blk_stride, int nblk, int ncomp,
                                    unsigned maxmask, unsigned skipmask, int neg0_skips, int C, double* __restrict__ sums,
                                    double* __restrict__ maxs) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double* rec = part + (long long)c * cand_stride;
    for (int k = 0; k < RS_RED_COMP; ++k) { sums[c * RS_RED_COMP + k] = 0.0; maxs[c * RS_RED_COMP + k] = -1.0e300; }
    if (neg0_skips && rec[0] < 0.0) { maxs[c * RS_RED_COMP] = -1.0; return; }
    for (int k = 0; k < ncomp; ++k) {
        if ((skipmask >> k) & 1u) continue;
        double a = rec[k];
        const bool mx = (maxmask >> k) & 1u;
        for (int q = 1; q < nblk; ++q) { const double bq = rec[(long long)q * blk_stride + k]; a = mx ? fmax(a, bq) : a + bq; }
        if (mx) maxs[c * RS_RED_COMP + k] = a; else sums[c * RS_RED_COMP + k] = a;
    }
}
__global__ void rs_expand4_kernel(double* __restrict__ part, long long cand_stride, int blk_stride, int nblk, int ncomp,
                                  unsigned maxmask, unsigned skipmask, int C, const double* __restrict__ sums,
                                  const double* __restrict__ maxs) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double* rec = part + (long long)c * cand_stride;
    for (int k = 0; k < ncomp; ++k) {
        double a = 0.0;
        if (!((skipmask >> k) & 1u)) a = ((maxmask >> k) & 1u) ? maxs[c * RS_RED_COMP + k] : sums[c * RS_RED_COMP + k];
        rec[k] = a;
        for (int q = 1; q < nblk; ++q) rec[(long long)q * blk_stride + k] = 0.0;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
// Tear-down order matters: a segment must not be freed while a peer still has it mapped.  close = unmap the peers' segments,
// free = release the own one; the collective variant puts a barrier before each phase (nobody writes any more / nobody maps any more).
static void rs_p2p_close_peers(RowShard* rs) {
    for (int r = 0; r < rs->world && r < RS_MAXG; ++r)
        if (r != rs->rank && rs->peer[r]) cudaIpcCloseMemHandle(rs->peer[r]);
    for (int r = 0; r < RS_MAXG; ++r) rs->peer[r] = nullptr;
}
static void rs_p2p_free_own(RowShard* rs) {
    cudaFree(rs->seg); rs->seg = nullptr; rs->seg_bytes = 0;
    cudaFree(rs->d_peer); rs->d_peer = nullptr;
    rs->p2p = false; rs->seg_C = 0; rs->seq_x = rs->seq_red = 0;
}
static void rs_p2p_release(RowShard* rs) { rs_p2p_close_peers(rs); rs_p2p_free_own(rs); }

static int rs_barrier(maus_ctx* ctx, RowShard* rs) {
    if (rs->world <= 1) return MAUS_OK;
    ncclResult_t r = g_nccl.AllReduce(rs->d_err, rs->d_err, 1, ncclInt32, ncclMax, rs->comm, ctx->stream);
    if (r != ncclSuccess) return maus_fail(ctx, MAUS_E_CUDA, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "ncclAllReduce");
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return maus_fail(ctx, MAUS_E_CUDA, "rs_barrier", e);
    return MAUS_OK;
}
// collective: every rank calls it at the same point
static int rs_p2p_release_collective(maus_ctx* ctx, RowShard* rs) {
    int rc = rs_barrier(ctx, rs); if (rc) return rc;          // nobody still writes into a peer's segment
    rs_p2p_close_peers(rs);
    if ((rc = rs_barrier(ctx, rs))) return rc;                // nobody still maps this rank's segment
    rs_p2p_free_own(rs);
    return MAUS_OK;
}

void maus_rowshard_free(maus_ctx* ctx) {
    RowShard* rs = (RowShard*)ctx->rowshard;
    if (!rs) return;
    cudaFree(rs->rowptr); cudaFree(rs->colidx); cudaFree(rs->vals); cudaFree(rs->diag); cudaFree(rs->xfull); cudaFree(rs->pack);
    cudaFree(rs->V); cudaFree(rs->X); cudaFree(rs->Y); cudaFree(rs->sigma); cudaFree(rs->lambda); cudaFree(rs->b); cudaFree(rs->psi);
    cudaFree(rs->alpha); cudaFree(rs->vnorm2); cudaFree(rs->resid); cudaFree(rs->mixnorm); cudaFree(rs->scratch);
    cudaFree(rs->jac); cudaFree(rs->status); cudaFree(rs->iters); cudaFree(rs->gsend); cudaFree(rs->grecv);
    rs_p2p_release(rs);
    cudaFree(rs->d_counter); cudaFree(rs->d_err);
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
    MAUS_CUDA(ctx, cudaMalloc(&rs->d_counter, sizeof(unsigned int)));
    MAUS_CUDA(ctx, cudaMalloc(&rs->d_err, sizeof(int)));
    MAUS_CUDA(ctx, cudaMemset(rs->d_counter, 0, sizeof(unsigned int)));
    MAUS_CUDA(ctx, cudaMemset(rs->d_err, 0, sizeof(int)));
    return MAUS_OK;
}

extern "C" int maus_dist_info(maus_ctx* ctx, int* rank, int* world, int* peer_memory) {
    RowShard* rs = ctx ? (RowShard*)ctx->rowshard : nullptr;
    if (!rs) return maus_fail(ctx, MAUS_E_STATE, "maus_dist_info: maus_dist_init first");
    if (rank) *rank = rs->rank;
    if (world) *world = rs->world;
    if (peer_memory) *peer_memory = rs->p2p ? 1 : 0;
    return MAUS_OK;
}

// per-generation exchange of the candidate-sharded mode: all-gather `count` doubles per rank on the context's
// stream (NCCL over NVLink), host buffers in and out
extern "C" int maus_gather(maus_ctx* ctx, const double* send, int64_t count, double* recv_all) {
    RowShard* rs = ctx ? (RowShard*)ctx->rowshard : nullptr;
    if (!rs || !rs->comm) return maus_fail(ctx, MAUS_E_STATE, "maus_gather: maus_dist_init first");
    if (!send || !recv_all || count <= 0) return maus_fail(ctx, MAUS_E_ARG, "maus_gather: bad argument");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    if ((size_t)count > rs->gcap) {
        MAUS_CUDA(ctx, cudaStreamSynchronize(st));
        cudaFree(rs->gsend); cudaFree(rs->grecv); rs->gsend = rs->grecv = nullptr; rs->gcap = 0;
        MAUS_CUDA(ctx, cudaMalloc(&rs->gsend, (size_t)count * 8));
        MAUS_CUDA(ctx, cudaMalloc(&rs->grecv, (size_t)count * 8 * rs->world));
        rs->gcap = (size_t)count;
    }
    MausNvtxRange range("maus.gather");
    MAUS_CUDA(ctx, cudaMemcpyAsync(rs->gsend, send, (size_t)count * 8, cudaMemcpyHostToDevice, st));
    MAUS_NCCL(ctx, g_nccl.AllGather(rs->gsend, rs->grecv, (size_t)count, ncclDouble, rs->comm, st));
    MAUS_CUDA(ctx, cudaMemcpyAsync(recv_all, rs->grecv, (size_t)count * 8 * rs->world, cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaStreamSynchronize(st));
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
    long long longest = 0;
    for (long long i = 0; i < nrows; ++i) longest = std::max<long long>(longest, (long long)(rowptr[i + 1] - rowptr[i]));
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
    MAUS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (rs->n != n || rs->nloc != nrows) {
        // a different order: everything sized by n / nloc is rebuilt on demand (collectively: every rank sets its block)
        cudaFree(rs->xfull); cudaFree(rs->pack); cudaFree(rs->V); cudaFree(rs->X); cudaFree(rs->Y); cudaFree(rs->b);
        rs->xfull = rs->pack = rs->V = rs->X = rs->Y = rs->b = nullptr; rs->pack_elems = 0; rs->Ccap = 0; rs->b_set = false;
        if (rs->seg) {
            int rc = rs_p2p_release_collective(ctx, rs); if (rc) return rc;
            rs->p2p_tried = false;
        }
    }
    cudaFree(rs->rowptr); cudaFree(rs->colidx); cudaFree(rs->vals); cudaFree(rs->diag);
    rs->rowptr = nullptr; rs->colidx = nullptr; rs->vals = nullptr; rs->diag = nullptr;
    MAUS_CUDA(ctx, cudaMalloc(&rs->rowptr, (size_t)(nrows + 1) * 8));
    MAUS_CUDA(ctx, cudaMalloc(&rs->colidx, (size_t)nnz * 4 + 16));          // + 16: the staged SpMM copies whole 16-byte units
    MAUS_CUDA(ctx, cudaMalloc(&rs->vals, std::max<size_t>((size_t)nnz * sizeof(cplx), 16)));
    MAUS_CUDA(ctx, cudaMalloc(&rs->diag, (size_t)nrows * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMemcpy(rs->rowptr, rp.data(), (size_t)(nrows + 1) * 8, cudaMemcpyHostToDevice));
    if (nnz) {
        MAUS_CUDA(ctx, cudaMemcpy(rs->colidx, ci.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice));
        MAUS_CUDA(ctx, cudaMemcpy(rs->vals, vals + 2 * rowptr[0], (size_t)nnz * sizeof(cplx), cudaMemcpyHostToDevice));
    }
    MAUS_CUDA(ctx, cudaMemcpy(rs->diag, dg.data(), (size_t)nrows * sizeof(cplx), cudaMemcpyHostToDevice));
    rs->n = n; rs->row0 = row0; rs->nloc = nrows; rs->nnz = nnz; rs->amax = amax; rs->max_row = (int)std::min<long long>(longest, 0x7fffffffLL);
    return MAUS_OK;
}

// (re)build the symmetric segment for C candidates and map every peer's copy.  Collective: all ranks call it with the same C.
static int rs_p2p_setup(maus_ctx* ctx, RowShard* rs, long long C) {
    cudaStream_t st = ctx->stream;
    if (rs->seg) { int rc = rs_p2p_release_collective(ctx, rs); if (rc) return rc; }
    rs->p2p_tried = true;
    const char* force = getenv("MAUS_RS_NCCL");
    if (force && atoi(force)) { rs->p2p = false; return MAUS_OK; }
    if (rs->world > RS_MAXG) { rs->p2p = false; return MAUS_OK; }
    const long long groups = (C + 3) / 4;
    rs->seg_C = groups * 4;
    rs->off_red = RS_FLAG_BYTES;
    const size_t red_bytes = (size_t)RS_RED_SLOTS * RS_MAXG * rs->seg_C * RS_RED_COMP * sizeof(double);
    rs->off_x = (rs->off_red + red_bytes + 4095) & ~(size_t)4095;
    rs->xbuf_bytes = (size_t)groups * rs->n * 4 * sizeof(cplx);
    rs->seg_bytes = rs->off_x + 2 * rs->xbuf_bytes;
    int ok = 1;
    if (cudaMalloc(&rs->seg, rs->seg_bytes) != cudaSuccess) { cudaGetLastError(); rs->seg = nullptr; ok = 0; }
    if (ok && cudaMemsetAsync(rs->seg, 0, rs->off_x, st) != cudaSuccess) ok = 0;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok && rs->world > 1 && cudaIpcGetMemHandle(&mine, rs->seg) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    std::vector<cudaIpcMemHandle_t> all((size_t)rs->world);
    std::vector<int> oks((size_t)rs->world, 1);
    if (rs->world > 1) {
        // exchange (handle, ok) through NCCL: 64 + 4 bytes per rank
        struct Rec { cudaIpcMemHandle_t h; int ok; int pad[3]; };
        static_assert(sizeof(Rec) == 80, "record size");
        Rec rec; rec.h = mine; rec.ok = ok; rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
        Rec *dsend = nullptr, *drecv = nullptr;
        MAUS_CUDA(ctx, cudaMalloc(&dsend, sizeof(Rec)));
        MAUS_CUDA(ctx, cudaMalloc(&drecv, sizeof(Rec) * rs->world));
        MAUS_CUDA(ctx, cudaMemcpyAsync(dsend, &rec, sizeof(Rec), cudaMemcpyHostToDevice, st));
        MAUS_NCCL(ctx, g_nccl.AllGather(dsend, drecv, sizeof(Rec), ncclChar, rs->comm, st));
        std::vector<Rec> recs((size_t)rs->world);
        MAUS_CUDA(ctx, cudaMemcpyAsync(recs.data(), drecv, sizeof(Rec) * rs->world, cudaMemcpyDeviceToHost, st));
        MAUS_CUDA(ctx, cudaStreamSynchronize(st));
        cudaFree(dsend); cudaFree(drecv);
        for (int r = 0; r < rs->world; ++r) { all[(size_t)r] = recs[(size_t)r].h; oks[(size_t)r] = recs[(size_t)r].ok; }
    } else {
        MAUS_CUDA(ctx, cudaStreamSynchronize(st));
    }
    for (int r = 0; r < rs->world; ++r) if (!oks[(size_t)r]) ok = 0;
    if (ok) {
        rs->peer[rs->rank] = rs->seg;
        for (int r = 0; r < rs->world && ok; ++r) {
            if (r == rs->rank) continue;
            if (cudaIpcOpenMemHandle(&rs->peer[r], all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError(); rs->peer[r] = nullptr; ok = 0;
            }
        }
    }
    // every rank must take the same transport: agree on the minimum
    if (rs->world > 1) {
        int* dflag = nullptr;
        MAUS_CUDA(ctx, cudaMalloc(&dflag, sizeof(int)));
        MAUS_CUDA(ctx, cudaMemcpyAsync(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
        MAUS_NCCL(ctx, g_nccl.AllReduce(dflag, dflag, 1, ncclInt32, ncclMin, rs->comm, st));
        MAUS_CUDA(ctx, cudaMemcpyAsync(&ok, dflag, sizeof(int), cudaMemcpyDeviceToHost, st));
        MAUS_CUDA(ctx, cudaStreamSynchronize(st));
        cudaFree(dflag);
    }
    if (!ok) return rs_p2p_release_collective(ctx, rs);                  // every rank takes this branch: NCCL transport
    MAUS_CUDA(ctx, cudaMalloc(&rs->d_peer, sizeof(void*) * RS_MAXG));
    MAUS_CUDA(ctx, cudaMemcpy(rs->d_peer, rs->peer, sizeof(void*) * RS_MAXG, cudaMemcpyHostToDevice));
    rs->p2p = true;
    rs->seq_x = rs->seq_red = 0;
    return MAUS_OK;
}

static int rs_ensure(maus_ctx* ctx, RowShard* rs, long long C) {
    if (C > rs->Ccap) {
        MAUS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(rs->xfull); cudaFree(rs->V); cudaFree(rs->X); cudaFree(rs->Y); cudaFree(rs->sigma); cudaFree(rs->lambda);
        cudaFree(rs->psi); cudaFree(rs->alpha); cudaFree(rs->vnorm2); cudaFree(rs->resid); cudaFree(rs->mixnorm); cudaFree(rs->scratch);
        cudaFree(rs->jac); cudaFree(rs->status); cudaFree(rs->iters);
        // a failing cudaMalloc below returns early: no pointer may stay dangling for maus_rowshard_free
        rs->xfull = rs->V = rs->X = rs->Y = rs->sigma = rs->lambda = nullptr;
        rs->psi = rs->alpha = rs->vnorm2 = rs->resid = rs->mixnorm = rs->scratch = nullptr; rs->jac = nullptr;
        rs->status = rs->iters = nullptr; rs->Ccap = 0;
        const long long cap = std::max<long long>(C, 4);
        MAUS_CUDA(ctx, cudaMalloc(&rs->xfull, (size_t)cap * rs->n * sizeof(cplx)));
        MAUS_CUDA(ctx, cudaMalloc(&rs->V, (size_t)cap * rs->nloc * sizeof(cplx)));
        MAUS_CUDA(ctx, cudaMalloc(&rs->X, (size_t)cap * rs->nloc * sizeof(cplx)));
        MAUS_CUDA(ctx, cudaMalloc(&rs->Y, (size_t)cap * rs->nloc * sizeof(cplx)));
        MAUS_CUDA(ctx, cudaMalloc(&rs->sigma, (size_t)cap * sizeof(cplx)));
        MAUS_CUDA(ctx, cudaMalloc(&rs->lambda, (size_t)cap * sizeof(cplx)));
        MAUS_CUDA(ctx, cudaMalloc(&rs->psi, (size_t)cap * 8));
        MAUS_CUDA(ctx, cudaMalloc(&rs->alpha, (size_t)cap * 8));
        MAUS_CUDA(ctx, cudaMalloc(&rs->vnorm2, (size_t)cap * 8));
        MAUS_CUDA(ctx, cudaMalloc(&rs->resid, (size_t)cap * 8));
        MAUS_CUDA(ctx, cudaMalloc(&rs->mixnorm, (size_t)cap * 8));
        MAUS_CUDA(ctx, cudaMalloc(&rs->scratch, vec_scratch_doubles(cap) * 8));
        MAUS_CUDA(ctx, cudaMalloc(&rs->jac, (size_t)cap));
        MAUS_CUDA(ctx, cudaMalloc(&rs->status, (size_t)cap * 4));
        MAUS_CUDA(ctx, cudaMalloc(&rs->iters, (size_t)cap * 4));
        rs->Ccap = cap;
    }
    if (!rs->p2p_tried || (rs->p2p && C > rs->seg_C)) {
        int rc = rs_p2p_setup(ctx, rs, std::max<long long>(C, rs->Ccap)); if (rc) return rc;
    }
    if (!rs->p2p) {
        const size_t need = csr_spmm_pack_elems(rs->n, (int)std::max<long long>(C, 2));
        if (need > rs->pack_elems) {
            MAUS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(rs->pack); rs->pack = nullptr; rs->pack_elems = 0;
            MAUS_CUDA(ctx, cudaMalloc(&rs->pack, need * sizeof(cplx)));
            rs->pack_elems = need;
        }
    }
    return MAUS_OK;
}

static int rs_check_err(maus_ctx* ctx, RowShard* rs) {
    if (!rs->p2p) return MAUS_OK;
    int e = 0;
    MAUS_CUDA(ctx, cudaMemcpy(&e, rs->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) return maus_fail(ctx, MAUS_E_STATE, "row-sharded operator: a peer-memory wait timed out (a rank died or left the collective)");
    return MAUS_OK;
}

// z[c] (local rows) = A_local * x[c], x = the C full-length vectors whose local slices are v[c]
static int rs_matvec(maus_ctx* ctx, RowShard* rs, const cplx* v, long long ldv, cplx* z, long long ldz, long long C) {
    cudaStream_t st = ctx->stream;
    const int groups = (int)((C + 3) / 4);
    int h = prof_begin(ctx, MAUS_PROF_MATVEC, (double)groups * (20.0 * rs->nnz + 8.0 * (rs->nloc + 1)) + 16.0 * (rs->n + rs->nloc) * C);
    if (rs->p2p) {
        const unsigned long long seq = ++rs->seq_x;
        const int buf = (int)(seq & 1);
        const size_t off = rs->off_x + (size_t)buf * rs->xbuf_bytes;
        const long long items = ((rs->nloc + RS_PUSH_ROWS - 1) / RS_PUSH_ROWS) * groups;
        const unsigned gpush = (unsigned)std::min<long long>(items, 2LL * MAUS_SM_COUNT_B200);
        rs_push_pack_kernel<<<gpush, 256, 0, st>>>(v, ldv, (int)C, rs->nloc, rs->row0, rs->n, rs->d_peer, off, buf, groups, rs->rank,
                                                   rs->world, seq, rs->d_counter, rs->d_err);
        rs_wait_x_kernel<<<1, 32, 0, st>>>(rs->d_peer, buf, rs->rank, rs->world, seq, rs->d_err);
        const cplx* P = reinterpret_cast<const cplx*>(rs->seg + off);
        MAUS_CUDA(ctx, csr_spmm_packed4(rs->rowptr, rs->colidx, rs->vals, P, rs->n * 4, z, ldz, rs->nloc, 0, (int)C, groups, rs->max_row, st));
        rs_done_x_kernel<<<1, 32, 0, st>>>(rs->d_peer, rs->rank, rs->world, seq);
        MAUS_CUDA(ctx, cudaGetLastError());
        ctx->launches += 4;
    } else {
        ncclResult_t r1 = g_nccl.GroupStart();
        ncclResult_t r2 = ncclSuccess;
        for (long long c = 0; c < C && r2 == ncclSuccess; ++c)
            r2 = g_nccl.AllGather(v + c * ldv, rs->xfull + c * rs->n, (size_t)rs->nloc * 2, ncclDouble, rs->comm, st);
        ncclResult_t r3 = g_nccl.GroupEnd();                                   // the group is closed on every path
        MAUS_NCCL(ctx, r1); MAUS_NCCL(ctx, r2); MAUS_NCCL(ctx, r3);
        MAUS_CUDA(ctx, csr_spmm(rs->rowptr, rs->colidx, rs->vals, rs->xfull, rs->n, z, ldz, rs->nloc, rs->n, (int)C,
                                C > 1 ? rs->pack : nullptr, rs->max_row, st));
        ctx->launches += 2;
    }
    prof_end(ctx, h);
    return MAUS_OK;
}

// combine per-candidate records over the ranks (see rs_allreduce_kernel)
static int rs_reduce_records(maus_ctx* ctx, RowShard* rs, double* part, long long cand_stride, int blk_stride, int nblk, int ncomp,
                             unsigned maxmask, unsigned skipmask, int neg0_skips, long long C) {
    cudaStream_t st = ctx->stream;
    if (rs->p2p) {
        const unsigned long long seq = ++rs->seq_red;
        rs_allreduce_kernel<<<1, 256, 0, st>>>(part, cand_stride, blk_stride, nblk, ncomp, maxmask, skipmask, neg0_skips, (int)C,
                                               rs->d_peer, rs->off_red, rs->seg_C, rs->rank, rs->world, seq, rs->d_err);
        MAUS_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
        return MAUS_OK;
    }
    // NCCL transport: xfull doubles as the staging area ([C][4] sums, [C][4] maxima)
    double* sums = reinterpret_cast<double*>(rs->xfull);
    double* maxs = sums + C * RS_RED_COMP;
    const unsigned g = (unsigned)((C + 127) / 128);
    rs_collapse4_kernel<<<g, 128, 0, st>>>(part, cand_stride, blk_stride, nblk, ncomp, maxmask, skipmask, neg0_skips, (int)C, sums, maxs);
    MAUS_NCCL(ctx, g_nccl.AllReduce(sums, sums, (size_t)C * RS_RED_COMP, ncclDouble, ncclSum, rs->comm, st));
    if (maxmask) MAUS_NCCL(ctx, g_nccl.AllReduce(maxs, maxs, (size_t)C * RS_RED_COMP, ncclDouble, ncclMax, rs->comm, st));
    rs_expand4_kernel<<<g, 128, 0, st>>>(part, cand_stride, blk_stride, nblk, ncomp, maxmask, skipmask, (int)C, sums, maxs);
    MAUS_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return MAUS_OK;
}

static int rs_reduce_flags(maus_ctx* ctx, RowShard* rs, int* flags, long long C) {
    cudaStream_t st = ctx->stream;
    if (rs->p2p) {
        const unsigned long long seq = ++rs->seq_red;
        rs_allreduce_int_max_kernel<<<1, 256, 0, st>>>(flags, (int)C, rs->d_peer, rs->off_red, rs->seg_C, rs->rank, rs->world, seq, rs->d_err);
        MAUS_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
        return MAUS_OK;
    }
    MAUS_NCCL(ctx, g_nccl.AllReduce(flags, flags, (size_t)C, ncclInt32, ncclMax, rs->comm, st));
    return MAUS_OK;
}

static GmresOperator rs_operator(maus_ctx* ctx, RowShard* rs) {
    GmresOperator op;
    op.nloc = rs->nloc; op.nglobal = rs->n; op.row0 = rs->row0; op.diag = rs->diag; op.amax = rs->amax; op.dense = false;
    op.matvec = [ctx, rs](const cplx* v, long long ldv, cplx* z, long long ldz, long long Cn) { return rs_matvec(ctx, rs, v, ldv, z, ldz, Cn); };
    // GMRES partial sums: [C][GM_MAXBLK] complex = records of 2 components, both summed
    op.reduce_partials = [ctx, rs](cplx* partial, int maxblk, int nblk, long long Cn) -> int {
        return rs_reduce_records(ctx, rs, reinterpret_cast<double*>(partial), 2LL * maxblk, 2, nblk, 2, 0u, 0u, 0, Cn);
    };
    op.flag_sync = [ctx, rs](int* flags, long long Cn) -> int { return rs_reduce_flags(ctx, rs, flags, Cn); };
    return op;
}

static RowShard* rs_ready(maus_ctx* ctx, const char* who) {
    RowShard* rs = ctx ? (RowShard*)ctx->rowshard : nullptr;
    if (!rs || !rs->rowptr) { maus_fail(ctx, MAUS_E_STATE, who); return nullptr; }
    cudaSetDevice(ctx->device);
    return rs;
}

extern "C" int maus_rs_matvec(maus_ctx* ctx, int64_t C, const double* V_local, double* Y_local) {
    RowShard* rs = rs_ready(ctx, "maus_rs_matvec: row block not set");
    if (!rs) return MAUS_E_STATE;
    if (C <= 0 || !V_local || !Y_local) return maus_fail(ctx, MAUS_E_ARG, "maus_rs_matvec: bad argument");
    int rc = rs_ensure(ctx, rs, C); if (rc) return rc;
    cudaStream_t st = ctx->stream;
    MAUS_CUDA(ctx, cudaMemcpyAsync(rs->V, V_local, (size_t)C * rs->nloc * sizeof(cplx), cudaMemcpyHostToDevice, st));
    if ((rc = rs_matvec(ctx, rs, rs->V, rs->nloc, rs->Y, rs->nloc, C))) return rc;
    MAUS_CUDA(ctx, cudaMemcpyAsync(Y_local, rs->Y, (size_t)C * rs->nloc * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaStreamSynchronize(st));
    return rs_check_err(ctx, rs);
}

// x_c = (A - sigma_c I + psi_c I)^-1 rhs_c by the batched GMRES on the row-sharded operator; RHS / X are local slices
extern "C" int maus_rs_gmres(maus_ctx* ctx, int64_t C, const double* sigma, const double* psi, const uint8_t* use_jacobi,
                             const double* RHS_local, double* X_local_out, int32_t* status_out, int32_t* iters_out) {
    RowShard* rs = rs_ready(ctx, "maus_rs_gmres: row block not set");
    if (!rs) return MAUS_E_STATE;
    if (C <= 0 || !sigma || !psi || !RHS_local) return maus_fail(ctx, MAUS_E_ARG, "maus_rs_gmres: bad argument");
    int rc = rs_ensure(ctx, rs, C); if (rc) return rc;
    cudaStream_t st = ctx->stream;
    MAUS_CUDA(ctx, cudaMemcpyAsync(rs->V, RHS_local, (size_t)C * rs->nloc * sizeof(cplx), cudaMemcpyHostToDevice, st));
    MAUS_CUDA(ctx, cudaMemcpyAsync(rs->sigma, sigma, (size_t)C * sizeof(cplx), cudaMemcpyHostToDevice, st));
    MAUS_CUDA(ctx, cudaMemcpyAsync(rs->psi, psi, (size_t)C * 8, cudaMemcpyHostToDevice, st));
    if (use_jacobi) MAUS_CUDA(ctx, cudaMemcpyAsync(rs->jac, use_jacobi, (size_t)C, cudaMemcpyHostToDevice, st));
    else MAUS_CUDA(ctx, cudaMemsetAsync(rs->jac, 0, (size_t)C, st));
    MAUS_CUDA(ctx, cudaMemsetAsync(rs->status, 0, (size_t)C * 4, st));
    MAUS_CUDA(ctx, cudaMemsetAsync(rs->iters, 0, (size_t)C * 4, st));
    GmresOperator op = rs_operator(ctx, rs);
    double max_psi = 0.0;
    for (long long c = 0; c < C; ++c) max_psi = std::max(max_psi, std::fabs(psi[c]));
    if ((rc = gmres_core(ctx, op, C, rs->sigma, rs->psi, nullptr, rs->jac, rs->V, rs->nloc, rs->X, rs->status, rs->iters, max_psi)))
        return rc;
    if (X_local_out) MAUS_CUDA(ctx, cudaMemcpyAsync(X_local_out, rs->X, (size_t)C * rs->nloc * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    if (status_out) MAUS_CUDA(ctx, cudaMemcpyAsync(status_out, rs->status, (size_t)C * 4, cudaMemcpyDeviceToHost, st));
    if (iters_out) MAUS_CUDA(ctx, cudaMemcpyAsync(iters_out, rs->iters, (size_t)C * 4, cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaStreamSynchronize(st));
    return rs_check_err(ctx, rs);
}

extern "C" int maus_rs_set_rhs(maus_ctx* ctx, const double* b_full) {
    RowShard* rs = rs_ready(ctx, "maus_rs_set_rhs: row block not set");
    if (!rs) return MAUS_E_STATE;
    if (!b_full) return maus_fail(ctx, MAUS_E_ARG, "maus_rs_set_rhs: bad argument");
    if (!rs->b) MAUS_CUDA(ctx, cudaMalloc(&rs->b, (size_t)rs->nloc * sizeof(cplx)));
    MAUS_CUDA(ctx, cudaMemcpyAsync(rs->b, b_full + 2 * rs->row0, (size_t)rs->nloc * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream));
    MAUS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rs->b_set = true;
    return MAUS_OK;
}

// One generation (or one ladder attempt, or a residual only) of AMS:574-576 on the row-sharded operator for C candidates that
// EVERY rank passes identically.  phases: 1 Rayleigh quotient (AMS:264-270) | 2 solve (AMS:44-97, GMRES) | 4 mix + normalise
// (AMS:280-285) | 8 residual (AMS:295-299).  V_full_io [C][n] host vectors: the rank's slices are uploaded, and after the call
// every rank holds the full updated vectors (one NCCL all-gather).  Without phase 1 the shifts / stale lambdas come from
// `sigma_in` [C] complex.
extern "C" int maus_rs_step(maus_ctx* ctx, int64_t C, int problem_type, int phases, double* V_full_io, const double* alpha,
                            const double* psi, const uint8_t* use_jacobi, const double* sigma_in, double* lambda_out,
                            double* resid_out, double* mixnorm_out, int32_t* status_out, int32_t* iters_out) {
    RowShard* rs = rs_ready(ctx, "maus_rs_step: row block not set");
    if (!rs) return MAUS_E_STATE;
    if (C <= 0 || !V_full_io) return maus_fail(ctx, MAUS_E_ARG, "maus_rs_step: bad argument");
    if (problem_type != MAUS_EIGENVALUE && problem_type != MAUS_SOLVE_LINEAR_SYSTEM) return maus_fail(ctx, MAUS_E_ARG, "maus_rs_step: problem type");
    const bool eigen = problem_type == MAUS_EIGENVALUE;
    const bool do_rq = (phases & 1) && eigen, do_solve = phases & 2, do_mix = phases & 4, do_res = phases & 8;
    if ((do_solve || do_mix) && (!alpha || !psi)) return maus_fail(ctx, MAUS_E_ARG, "maus_rs_step: alpha / psi required");
    if (eigen && !do_rq && !sigma_in) return maus_fail(ctx, MAUS_E_ARG, "maus_rs_step: sigma required without the Rayleigh-quotient phase");
    if (!eigen && !rs->b_set) return maus_fail(ctx, MAUS_E_STATE, "maus_rs_step: rhs not set");
    int rc = rs_ensure(ctx, rs, C); if (rc) return rc;
    MausNvtxRange range("maus.rowshard.step");
    cudaStream_t st = ctx->stream;
    const long long n = rs->n, nl = rs->nloc;
    const int nblk = vec_part_blocks(nl);
    const long long cs = (long long)VEC_PART_MAXBLK * 4;          // scratch record stride (doubles)
    MAUS_CUDA(ctx, cudaMemcpy2DAsync(rs->V, (size_t)nl * sizeof(cplx), reinterpret_cast<const cplx*>(V_full_io) + rs->row0,
                                     (size_t)n * sizeof(cplx), (size_t)nl * sizeof(cplx), (size_t)C, cudaMemcpyHostToDevice, st));
    if (alpha) MAUS_CUDA(ctx, cudaMemcpyAsync(rs->alpha, alpha, (size_t)C * 8, cudaMemcpyHostToDevice, st));
    if (psi) MAUS_CUDA(ctx, cudaMemcpyAsync(rs->psi, psi, (size_t)C * 8, cudaMemcpyHostToDevice, st));
    if (use_jacobi) MAUS_CUDA(ctx, cudaMemcpyAsync(rs->jac, use_jacobi, (size_t)C, cudaMemcpyHostToDevice, st));
    else MAUS_CUDA(ctx, cudaMemsetAsync(rs->jac, 0, (size_t)C, st));
    MAUS_CUDA(ctx, cudaMemsetAsync(rs->status, 0, (size_t)C * 4, st));
    MAUS_CUDA(ctx, cudaMemsetAsync(rs->iters, 0, (size_t)C * 4, st));
    MAUS_CUDA(ctx, cudaMemsetAsync(rs->mixnorm, 0, (size_t)C * 8, st));
    if (do_rq) {
        if ((rc = rs_matvec(ctx, rs, rs->V, nl, rs->Y, nl, C))) return rc;
        MAUS_CUDA(ctx, vec_rq_part(rs->V, rs->Y, (int)nl, (int)C, rs->scratch, nblk, st));
        if ((rc = rs_reduce_records(ctx, rs, rs->scratch, cs, 4, nblk, 3, 0u, 0u, 0, C))) return rc;
        MAUS_CUDA(ctx, vec_rq_final(rs->scratch, nblk, (int)C, rs->lambda, rs->vnorm2, rs->status, st));
        MAUS_CUDA(ctx, cudaMemcpyAsync(rs->sigma, rs->lambda, (size_t)C * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
        ctx->launches += 2;
    } else if (eigen) {
        MAUS_CUDA(ctx, cudaMemcpyAsync(rs->sigma, sigma_in, (size_t)C * sizeof(cplx), cudaMemcpyHostToDevice, st));
        MAUS_CUDA(ctx, cudaMemcpyAsync(rs->lambda, sigma_in, (size_t)C * sizeof(cplx), cudaMemcpyHostToDevice, st));
    } else {
        MAUS_CUDA(ctx, cudaMemsetAsync(rs->sigma, 0, (size_t)C * sizeof(cplx), st));
        MAUS_CUDA(ctx, cudaMemsetAsync(rs->lambda, 0, (size_t)C * sizeof(cplx), st));
    }
    if (do_solve) {
        GmresOperator op = rs_operator(ctx, rs);
        double max_psi = 0.0;
        for (long long c = 0; c < C; ++c) max_psi = std::max(max_psi, std::fabs(psi[c]));
        const cplx* rhs = eigen ? rs->V : rs->b;
        if ((rc = gmres_core(ctx, op, C, rs->sigma, rs->psi, nullptr, rs->jac, rhs, eigen ? nl : 0, rs->X, rs->status, rs->iters, max_psi)))
            return rc;
    }
    if (do_mix) {
        MAUS_CUDA(ctx, vec_mix_part(rs->V, rs->X, (int)nl, (int)C, rs->alpha, rs->status, rs->scratch, nblk, st));
        if ((rc = rs_reduce_records(ctx, rs, rs->scratch, cs, 4, nblk, 2, 1u, 0u, 1, C))) return rc;
        MAUS_CUDA(ctx, vec_mix_apply(rs->V, (int)nl, (int)C, problem_type, rs->mixnorm, rs->status, rs->scratch, nblk, st));
        ctx->launches += 3;
    }
    if (do_res) {
        if ((rc = rs_matvec(ctx, rs, rs->V, nl, rs->Y, nl, C))) return rc;
        MAUS_CUDA(ctx, vec_res_part(rs->V, rs->Y, (int)nl, (int)C, problem_type, rs->lambda, rs->b, rs->scratch, nblk, st));
        if ((rc = rs_reduce_records(ctx, rs, rs->scratch, cs, 4, nblk, 2, 0x0u, 0x1u, 0, C))) return rc;      // component 1: sum of squares
        MAUS_CUDA(ctx, vec_res_final(rs->V, rs->Y, (int)nl, (int)C, problem_type, rs->lambda, rs->b, rs->scratch, nblk, rs->resid, st));
        ctx->launches += 2;
    }
    // write-back: every rank receives the full vectors (bulk transfer, once per generation: NCCL over NVLink)
    if (do_mix || do_solve) {
        ncclResult_t r1 = g_nccl.GroupStart();
        ncclResult_t r2 = ncclSuccess;
        for (long long c = 0; c < C && r2 == ncclSuccess; ++c)
            r2 = g_nccl.AllGather(rs->V + c * nl, rs->xfull + c * n, (size_t)nl * 2, ncclDouble, rs->comm, st);
        ncclResult_t r3 = g_nccl.GroupEnd();
        MAUS_NCCL(ctx, r1); MAUS_NCCL(ctx, r2); MAUS_NCCL(ctx, r3);
        MAUS_CUDA(ctx, cudaMemcpyAsync(V_full_io, rs->xfull, (size_t)C * n * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    }
    if (lambda_out) MAUS_CUDA(ctx, cudaMemcpyAsync(lambda_out, rs->lambda, (size_t)C * sizeof(cplx), cudaMemcpyDeviceToHost, st));
    if (resid_out) MAUS_CUDA(ctx, cudaMemcpyAsync(resid_out, rs->resid, (size_t)C * 8, cudaMemcpyDeviceToHost, st));
    if (mixnorm_out) MAUS_CUDA(ctx, cudaMemcpyAsync(mixnorm_out, rs->mixnorm, (size_t)C * 8, cudaMemcpyDeviceToHost, st));
    if (status_out) MAUS_CUDA(ctx, cudaMemcpyAsync(status_out, rs->status, (size_t)C * 4, cudaMemcpyDeviceToHost, st));
    if (iters_out) MAUS_CUDA(ctx, cudaMemcpyAsync(iters_out, rs->iters, (size_t)C * 4, cudaMemcpyDeviceToHost, st));
    MAUS_CUDA(ctx, cudaStreamSynchronize(st));
    return rs_check_err(ctx, rs);
}
