"""Row-sharded sparse operator: BASELINE config #5 as worded ("sparse CSC ... row-sharded across 8 B200").

Rank r owns rows [r*n/G, (r+1)*n/G) of A (a CSR slice with global column indices) and the same slice of every candidate
vector.  The library all-gathers the input of each SpMM over NVLink and all-reduces every GMRES dot product (NCCL, bound
at run time inside libmaus_b200: ``maus_dist_init``).  The solver is the SAME batched GMRES as the replicated path
(AMS:61-90 -> scipy gmres control flow); only the operator differs.  One process per GPU; the 128-byte NCCL id travels
through ``torch.distributed`` (any backend) or a caller-supplied broadcast.
"""
import ctypes as C
import glob
import os

import numpy as np

from ._abi import MausError

_c128 = np.complex128


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def nccl_library_path():
    """libnccl.so.2 that torch itself loads (so that both users share one copy); '' = let the loader search"""
    try:
        import nvidia.nccl
        base = list(getattr(nvidia.nccl, "__path__", [])) or [os.path.dirname(nvidia.nccl.__file__)]
        for b in base:
            hits = sorted(glob.glob(os.path.join(b, "lib", "libnccl.so*")))
            if hits:
                return hits[0]
    except Exception:
        pass
    return ""


def row_block(n, rank, world):
    """(row0, nrows) of the block a rank owns; equal blocks are required (the all-gather is not ragged)"""
    if n % world:
        raise ValueError(f"row sharding needs n % world == 0 (n={n}, world={world}); pad the operator")
    nloc = n // world
    return rank * nloc, nloc


def csr_row_block(A, rank, world):
    """CSR slice (rowptr[nloc+1] rebased to 0, global colidx, values) of a scipy.sparse matrix for one rank"""
    Ar = A.tocsr()
    Ar.sort_indices()
    n = Ar.shape[0]
    if Ar.shape[0] != Ar.shape[1]:
        raise ValueError("square matrix required")
    row0, nloc = row_block(n, rank, world)
    lo, hi = int(Ar.indptr[row0]), int(Ar.indptr[row0 + nloc])
    rowptr = np.ascontiguousarray(Ar.indptr[row0:row0 + nloc + 1], dtype=np.int64) - lo
    colidx = np.ascontiguousarray(Ar.indices[lo:hi], dtype=np.int64)
    vals = np.ascontiguousarray(Ar.data[lo:hi], dtype=_c128)
    return n, row0, nloc, rowptr, colidx, vals


class RowShardedOperator:
    """One rank's part of the row-sharded operator, living in ``engine``'s context."""

    def __init__(self, engine, rank=0, world=1, broadcast=None):
        """``broadcast(bytes_or_None) -> bytes``: returns rank 0's argument on every rank.  Default: torch.distributed
        (``broadcast_object_list``) when world > 1."""
        self.engine, self.rank, self.world = engine, int(rank), int(world)
        self._lib = engine._lib
        lib = nccl_library_path().encode()
        ident = None
        if self.rank == 0:
            buf = C.create_string_buffer(128)
            rc = self._lib.maus_nccl_unique_id(lib, buf)
            if rc != 0:
                raise MausError(f"maus_nccl_unique_id failed (rc={rc}): libnccl.so.2 not loadable")
            ident = buf.raw
        if self.world > 1:
            if broadcast is None:
                import torch.distributed as dist
                box = [ident]
                dist.broadcast_object_list(box, src=0)
                ident = box[0]
            else:
                ident = broadcast(ident)
        engine._check(self._lib.maus_dist_init(engine._h, lib, self.rank, self.world, ident))
        self.n = self.row0 = self.nloc = 0
        self._matrix = None

    def info(self):
        r, w, p = C.c_int32(), C.c_int32(), C.c_int32()
        self.engine._check(self._lib.maus_dist_info(self.engine._h, C.byref(r), C.byref(w), C.byref(p)))
        return dict(rank=r.value, world=w.value, peer_memory=bool(p.value))

    def mode_description(self):
        pm = self.info()["peer_memory"]
        return (f"matrix row-sharded x{self.world}; " +
                ("NVLink peer memory: pack + all-gather fused into one store kernel, one-kernel rank-ordered reductions"
                 if pm else "NCCL transport: all-gather per SpMM, all-reduce per dot product"))

    def ensure_matrix(self, A):
        """upload this rank's row block of A unless the same object is already resident"""
        if self._matrix is not A:
            self.set_matrix(A)
            self._matrix = A

    def gather(self, send, pinned=False):
        """all-gather a float64 array of identical length from every rank (NCCL all-gather on the context's stream,
        ``maus_gather``); returns [world][len].  ``pinned``: the result lands in one of TWO page-locked buffers that are
        reused alternately (no page faults of a fresh 8 world n C-byte array every generation, D2H at the pinned rate); the
        caller may keep views into a result until the call after the next one."""
        send = np.ascontiguousarray(send, dtype=np.float64).ravel()
        if pinned:
            bufs = getattr(self, "_gather_bufs", None)
            need = self.world * send.size
            if bufs is None or bufs[0].size < need:
                bufs = self._gather_bufs = [self.engine.pinned_empty((need,), dtype=np.float64) for _ in range(2)]
                self._gather_turn = 0
            out = bufs[self._gather_turn][:need].reshape(self.world, send.size)
            self._gather_turn ^= 1
        else:
            out = np.empty((self.world, send.size), dtype=np.float64)
        self.engine._check(self._lib.maus_gather(self.engine._h, _dp(send), send.size, _dp(out)))
        return out

    def set_rhs(self, b):
        b = np.ascontiguousarray(b, dtype=_c128)
        if b.shape != (self.n,):
            raise ValueError("b must be the full right-hand side [n]")
        self.engine._check(self._lib.maus_rs_set_rhs(self.engine._h, _dp(b)))

    PH_RQ, PH_SOLVE, PH_MIX, PH_RESIDUAL = 1, 2, 4, 8

    def step(self, problem_type, V, alpha=None, psi=None, use_jacobi=None, sigma=None, phases=15):
        """``maus_rs_step``: one generation (phases 15), one ladder attempt (14 with sigma) or a residual (8 with sigma) for C
        candidates that every rank passes identically; V [C][n] full-length, updated in place on every rank."""
        if V.dtype != _c128 or not V.flags.c_contiguous or V.ndim != 2 or V.shape[1] != self.n:
            raise ValueError("V must be a C-contiguous complex128 [C][n] array (updated in place)")
        C_ = V.shape[0]
        f64 = lambda a: None if a is None else np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)      # noqa: E731
        alpha, psi = f64(alpha), f64(psi)
        jac = None if use_jacobi is None else np.ascontiguousarray(np.atleast_1d(use_jacobi), dtype=np.uint8)
        sig = None if sigma is None else np.ascontiguousarray(np.atleast_1d(sigma), dtype=_c128)
        out = dict(lam=np.empty(C_, dtype=_c128), resid=np.empty(C_, dtype=np.float64), mixnorm=np.empty(C_, dtype=np.float64),
                   status=np.empty(C_, dtype=np.int32), iters=np.empty(C_, dtype=np.int32))
        self.engine._check(self._lib.maus_rs_step(
            self.engine._h, C_, int(problem_type), int(phases), _dp(V), _dp(alpha), _dp(psi), _ip(jac, C.c_uint8), _dp(sig),
            _dp(out["lam"]), _dp(out["resid"]), _dp(out["mixnorm"]), _ip(out["status"], C.c_int32), _ip(out["iters"], C.c_int32)))
        return out

    def set_matrix(self, A):
        """A: the FULL scipy.sparse matrix (every rank slices its own rows) -- or use ``set_row_block`` directly"""
        return self.set_row_block(*csr_row_block(A, self.rank, self.world))

    def set_row_block(self, n, row0, nloc, rowptr, colidx, vals):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        colidx = np.ascontiguousarray(colidx, dtype=np.int64)
        vals = np.ascontiguousarray(vals, dtype=_c128)
        self.engine._check(self._lib.maus_set_csr_rowblock(self.engine._h, int(n), int(row0), int(nloc),
                                                           _ip(rowptr, C.c_int64), _ip(colidx, C.c_int64), _dp(vals)))
        self.n, self.row0, self.nloc = int(n), int(row0), int(nloc)
        self._matrix = None

    def local(self, full):
        """slice [..., row0:row0+nloc] of full-length vector(s)"""
        return np.ascontiguousarray(np.asarray(full)[..., self.row0:self.row0 + self.nloc], dtype=_c128)

    def matvec(self, V_local):
        V_local = np.ascontiguousarray(V_local, dtype=_c128)
        if V_local.ndim != 2 or V_local.shape[1] != self.nloc:
            raise ValueError("V_local must be [C][nloc]")
        Y = np.empty_like(V_local)
        self.engine._check(self._lib.maus_rs_matvec(self.engine._h, V_local.shape[0], _dp(V_local), _dp(Y)))
        return Y

    def gmres(self, sigma, psi, RHS_local, use_jacobi=None, want_x=True):
        """x_c = (A - sigma_c I + psi_c I)^-1 rhs_c; returns (X_local [C][nloc], status [C], inner iterations [C])"""
        sigma = np.ascontiguousarray(np.atleast_1d(sigma), dtype=_c128)
        C_ = sigma.shape[0]
        psi = np.ascontiguousarray(np.atleast_1d(psi), dtype=np.float64)
        RHS_local = np.ascontiguousarray(RHS_local, dtype=_c128)
        if RHS_local.shape != (C_, self.nloc) or psi.shape != (C_,):
            raise ValueError("sigma [C], psi [C], RHS_local [C][nloc] expected")
        jac = None if use_jacobi is None else np.ascontiguousarray(np.atleast_1d(use_jacobi), dtype=np.uint8)
        X = np.empty((C_, self.nloc), dtype=_c128) if want_x else None
        status = np.empty(C_, dtype=np.int32)
        iters = np.empty(C_, dtype=np.int32)
        self.engine._check(self._lib.maus_rs_gmres(self.engine._h, C_, _dp(sigma), _dp(psi), _ip(jac, C.c_uint8),
                                                   _dp(RHS_local), _dp(X), _ip(status, C.c_int32), _ip(iters, C.c_int32)))
        return X, status, iters
