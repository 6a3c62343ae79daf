"""MausEngine -- numpy-facing wrapper of one libmaus_b200 context (one GPU).

Everything numeric happens in the CUDA library; this class only marshals numpy arrays through the C ABI
(include/maus_b200.h).  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import MausError

_c128 = np.complex128


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _as_c128(a, shape=None):
    a = np.ascontiguousarray(a, dtype=_c128)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class MausEngine:
    """One GPU context.  ``set_matrix`` / ``set_rhs`` upload the problem once; the candidate vectors can stay
    resident between steps (``resident=True`` paths) or travel with every call."""

    def __init__(self, device=0, workspace_limit_bytes=0):
        self._lib = _abi.load_library()
        h = C.c_void_p()
        rc = self._lib.maus_create(C.byref(h), int(device))
        if rc != 0 or not h:
            raise MausError(f"maus_create(device={device}) failed (rc={rc}): no usable sm_100a GPU -- "
                            f"libmaus_b200 has no CPU fallback")
        self._h = h
        self.device = int(device)
        self.n = 0
        self.is_sparse = False
        self.has_dense_form = False         # the batched LU can run on slot 0 (dense matrix, or sparse + attached dense form)
        self.generation = 0
        self.matrix_epoch = [0, 0]          # uploads per slot; population._MatrixCache compares it (shared-engine safety)
        self.rowshard = None                # RowShardedOperator of this context (rowshard.py), set by enable_row_sharding
        self._pinned = []
        self._close_hooks = []              # callables run before the page-locked buffers are released (dist.py: detach views)
        if workspace_limit_bytes:
            self._check(self._lib.maus_set_workspace_limit(self._h, int(workspace_limit_bytes)))

    # -- plumbing ---------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self._lib.maus_last_error(self._h)
            raise MausError(f"libmaus_b200 error {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_h", None):
            for hook in getattr(self, "_close_hooks", []):
                hook()
            self._close_hooks = []
            for p in self._pinned:
                self._lib.maus_free_pinned(p)
            self._pinned = []
            self._lib.maus_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def enable_row_sharding(self, rank, world, broadcast=None):
        """Multi-GPU set-up of the context (one process per GPU): NCCL communicator + the row-sharded sparse operator.  After
        this call ``step_population`` runs sparse GMRES problems with the matrix row-sharded over the ranks (BASELINE config 5
        as worded) and ``dist.Shard`` can use ``maus_gather`` for the per-generation exchange."""
        from .rowshard import RowShardedOperator
        if self.rowshard is None or self.rowshard.world != int(world) or self.rowshard.rank != int(rank):
            self.rowshard = RowShardedOperator(self, rank, world, broadcast)
        return self.rowshard

    def pinned_empty(self, shape, dtype=_c128):
        """numpy array backed by page-locked host memory (for the e2e path)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self._lib.maus_alloc_pinned(max(nbytes, 16))
        if not p:
            raise MausError("maus_alloc_pinned failed")
        self._pinned.append(p)
        buf = (C.c_char * max(nbytes, 16)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def staging(self, shape):
        """Reusable page-locked [C][n] complex128 buffer for the per-step host<->device copies of the candidate vectors."""
        need = int(np.prod(shape))
        buf = getattr(self, "_staging", None)
        if buf is None or buf.size < need:
            buf = self._staging = self.pinned_empty((max(need, 1),))
        return buf[:need].reshape(shape)

    @property
    def stream_ptr(self):
        """cudaStream_t of the context (int), e.g. for torch.cuda.ExternalStream"""
        return int(self._lib.maus_stream(self._h) or 0)

    @property
    def launches(self):
        return int(self._lib.maus_launch_count(self._h))

    def info(self):
        d, s, b = C.c_int32(), C.c_int32(), C.c_int64()
        self._check(self._lib.maus_info(self._h, C.byref(d), C.byref(s), C.byref(b)))
        return dict(device=d.value, sm_count=s.value, bytes_held=b.value)

    def profile_reset(self, enable=True):
        self._check(self._lib.maus_profile_reset(self._h, 1 if enable else 0))

    def profile_read(self):
        a, b, c, d, e, f = C.c_double(), C.c_int64(), C.c_double(), C.c_double(), C.c_int64(), C.c_double()
        self._check(self._lib.maus_profile_read(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e),
                                                C.byref(f)))
        return dict(lu_gemm_ms=a.value, lu_gemm_launches=b.value, lu_gemm_flops=c.value, matvec_ms=d.value,
                    matvec_launches=e.value, matvec_bytes=f.value)

    PROF_KINDS = ("lu_gemm", "matvec", "panel", "trtri", "backsolve", "build", "permute", "matvec_gemm", "vec")

    def profile_breakdown(self):
        """{kind: dict(ms, launches, work)} accumulated since profile_reset(True)"""
        out = {}
        for k, name in enumerate(self.PROF_KINDS):
            ms, ln, wk = C.c_double(), C.c_int64(), C.c_double()
            self._check(self._lib.maus_profile_read_kind(self._h, k, C.byref(ms), C.byref(ln), C.byref(wk)))
            out[name] = dict(ms=ms.value, launches=ln.value, work=wk.value)
        return out

    # -- problem ----------------------------------------------------------------------------------------------
    def set_matrix(self, A, slot=_abi.SLOT_CURRENT):
        """A: numpy 2-D array (any dtype, coerced like AMS:343) or scipy.sparse matrix (kept sparse, AMS:342)."""
        try:
            import scipy.sparse as sp
            sparse = sp.issparse(A)
        except Exception:  # pragma: no cover
            sparse = False
        if sparse:
            Ac = A.tocsc()
            Ac.sort_indices()
            n = Ac.shape[0]
            if Ac.shape[0] != Ac.shape[1]:
                raise ValueError("square matrix required")
            indptr = np.ascontiguousarray(Ac.indptr, dtype=np.int64)
            indices = np.ascontiguousarray(Ac.indices, dtype=np.int64)
            data = _as_c128(Ac.data)
            self._check(self._lib.maus_set_csc(self._h, int(slot), n, int(Ac.nnz),
                                               indptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                               indices.ctypes.data_as(C.POINTER(C.c_int64)), _dp(data)))
            if slot == _abi.SLOT_CURRENT:
                self.is_sparse = True
                self.has_dense_form = False
        else:
            A = _as_c128(A)
            if A.ndim != 2 or A.shape[0] != A.shape[1]:
                raise ValueError("square matrix required")
            n = A.shape[0]
            self._check(self._lib.maus_set_dense(self._h, int(slot), n, _dp(A)))
            if slot == _abi.SLOT_CURRENT:
                self.is_sparse = False
                self.has_dense_form = True
        if slot == _abi.SLOT_CURRENT:
            self.n = n
            self.matrix_epoch[1] += 1       # a new slot-0 matrix may change n, which drops slot 1 on the device
        self.matrix_epoch[slot] += 1

    def add_dense_form(self, A_dense):
        """keep the sparse matrix of slot 0 for the matvecs and attach its dense form for the batched LU (ladder fallback)"""
        A_dense = _as_c128(A_dense, (self.n, self.n))
        self._check(self._lib.maus_add_dense_form(self._h, self.n, _dp(A_dense)))
        self.has_dense_form = True

    def set_rhs(self, b):
        b = _as_c128(b, (self.n,))
        self._check(self._lib.maus_set_rhs(self._h, _dp(b)))

    # -- resident vectors -------------------------------------------------------------------------------------
    def upload_vectors(self, V):
        V = _as_c128(V)
        if V.ndim != 2 or V.shape[1] != self.n:
            raise ValueError("V must be [C][n]")
        self._check(self._lib.maus_upload_vectors(self._h, V.shape[0], _dp(V)))

    def download_vectors(self, C_, out=None):
        out = np.empty((C_, self.n), dtype=_c128) if out is None else out
        self._check(self._lib.maus_download_vectors(self._h, int(C_), _dp(out)))
        return out

    def download_vector_range(self, first, count=1):
        out = np.empty((count, self.n), dtype=_c128)
        self._check(self._lib.maus_download_vector_range(self._h, int(first), int(count), _dp(out)))
        return out

    # -- granular pieces --------------------------------------------------------------------------------------
    def rq(self, V=None, C_=None):
        if V is not None:
            V = _as_c128(V)
            C_ = V.shape[0]
        lam = np.empty(C_, dtype=_c128)
        vn2 = np.empty(C_, dtype=np.float64)
        self._check(self._lib.maus_rq(self._h, int(C_), _dp(V), _dp(lam), _dp(vn2)))
        return lam, vn2

    def solve_shifted(self, sigma, psi, rng_key=None, method=_abi.METHOD_LU, use_jacobi=None, RHS=None,
                      rhs_shared=False, want_x=True):
        sigma = _as_c128(np.atleast_1d(sigma))
        C_ = sigma.shape[0]
        psi = np.ascontiguousarray(np.atleast_1d(psi), dtype=np.float64)
        keys = None if rng_key is None else np.ascontiguousarray(np.atleast_1d(rng_key), dtype=np.uint64)
        jac = None if use_jacobi is None else np.ascontiguousarray(np.atleast_1d(use_jacobi), dtype=np.uint8)
        if RHS is not None:
            RHS = _as_c128(RHS)
        X = np.empty((C_, self.n), dtype=_c128) if want_x else None
        status = np.empty(C_, dtype=np.int32)
        iters = np.empty(C_, dtype=np.int32)
        self._check(self._lib.maus_solve_shifted(
            self._h, C_, _dp(sigma), _dp(psi),
            None if keys is None else keys.ctypes.data_as(C.POINTER(C.c_uint64)), int(method),
            None if jac is None else jac.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(RHS), 1 if rhs_shared else 0,
            _dp(X), status.ctypes.data_as(C.POINTER(C.c_int32)), iters.ctypes.data_as(C.POINTER(C.c_int32))))
        return X, status, iters

    def solve_with_R(self, sigma, psi, R, rhs):
        sigma = _as_c128(np.atleast_1d(sigma))
        psi = np.ascontiguousarray(np.atleast_1d(psi), dtype=np.float64)
        R = _as_c128(R, (self.n, self.n))
        rhs = _as_c128(rhs, (self.n,))
        x = np.empty(self.n, dtype=_c128)
        st = np.empty(1, dtype=np.int32)
        self._check(self._lib.maus_solve_with_R(self._h, _dp(sigma), _dp(psi), _dp(R), _dp(rhs), _dp(x),
                                                st.ctypes.data_as(C.POINTER(C.c_int32))))
        return x, int(st[0])

    def mix_residual(self, problem_type, alpha, lambda_old=None, skip=None, res_slot=_abi.SLOT_CURRENT,
                     want_v=True):
        alpha = np.ascontiguousarray(np.atleast_1d(alpha), dtype=np.float64)
        C_ = alpha.shape[0]
        lam = None if lambda_old is None else _as_c128(np.atleast_1d(lambda_old))
        sk = None if skip is None else np.ascontiguousarray(np.atleast_1d(skip), dtype=np.uint8)
        V = np.empty((C_, self.n), dtype=_c128) if want_v else None
        resid = np.empty(C_, dtype=np.float64)
        mixn = np.empty(C_, dtype=np.float64)
        status = np.empty(C_, dtype=np.int32)
        self._check(self._lib.maus_mix_residual(
            self._h, C_, int(problem_type), _dp(alpha), _dp(lam),
            None if sk is None else sk.ctypes.data_as(C.POINTER(C.c_uint8)), int(res_slot), _dp(V), _dp(resid),
            _dp(mixn), status.ctypes.data_as(C.POINTER(C.c_int32))))
        return V, resid, mixn, status

    def residual(self, problem_type, V=None, lam=None, C_=None, res_slot=_abi.SLOT_CURRENT):
        if V is not None:
            V = _as_c128(V)
            C_ = V.shape[0]
        lam = None if lam is None else _as_c128(np.atleast_1d(lam))
        resid = np.empty(C_, dtype=np.float64)
        self._check(self._lib.maus_residual(self._h, int(C_), int(problem_type), _dp(V), _dp(lam), int(res_slot),
                                            _dp(resid)))
        return resid

    def debug_zgemm(self, A, B, Cm, beta=0, negate=False, use_dmma=True):
        """Parity hook: batched column-major complex GEMM; A [batch][K][M] (i.e. column-major M x K), etc.
        use_dmma: 0 plain FP64-FMA kernel, 1 tensor-pipe kernel (4 real products), 2 tensor-pipe 3M kernel of the LU."""
        A = _as_c128(A); B = _as_c128(B); Cm = _as_c128(Cm).copy()
        batch, K, M = A.shape
        _, N, K2 = B.shape
        assert K2 == K and Cm.shape == (batch, N, M)
        self._check(self._lib.maus_debug_zgemm(self._h, M, N, K, batch, _dp(A), _dp(B), _dp(Cm), int(beta),
                                               1 if negate else 0, int(use_dmma)))
        return Cm

    def project(self, Ec, V):
        """P[c] = E^H v_c for the rows of V ([C][n]) -- the similarity scores of the Hermitian shortcut (AMS:165) as one
        tensor-pipe GEMM.  ``Ec`` = conj(E) in C order ([n][m], i.e. E^H stored column-major); returns [C][m]."""
        Ec = _as_c128(Ec); V = _as_c128(V)
        n, m = Ec.shape
        if V.ndim != 2 or V.shape[1] != n:
            raise ValueError("V must be [C][n]")
        out = np.empty((V.shape[0], m), dtype=_c128)
        self._check(self._lib.maus_project(self._h, n, m, _dp(Ec), V.shape[0], _dp(V), _dp(out)))
        return out

    def heev(self, A, max_sweeps=0, vectors=True):
        """(w, E) = eigendecomposition of the dense Hermitian matrix A on the device (Jacobi; replaces sla.eigh of AMS:161):
        w ascending float64 [n], E [n][n] complex128 with unit eigenvectors as columns.  Raises MausError when the sweeps do
        not converge."""
        A = _as_c128(A)
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise ValueError("square matrix required")
        n = A.shape[0]
        w = np.empty(n, dtype=np.float64)
        E = np.empty((n, n), dtype=_c128) if vectors else None
        sw, off = C.c_int32(), C.c_double()
        self._check(self._lib.maus_heev(self._h, n, _dp(A), int(max_sweeps), _dp(w), _dp(E), C.byref(sw), C.byref(off)))
        self.heev_info = dict(sweeps=sw.value, off_ratio=off.value)
        return (w, E) if vectors else w

    def gram(self, V):
        """G[i][j] = np.vdot(V[i], V[j]) for the rows of V ([C][n] complex128) in one device pass (dedup similarity tests)."""
        V = _as_c128(V)
        if V.ndim != 2:
            raise ValueError("V must be [C][n]")
        G = np.empty((V.shape[0], V.shape[0]), dtype=_c128)
        if V.shape[0]:
            self._check(self._lib.maus_gram(self._h, V.shape[0], V.shape[1], _dp(V), _dp(G)))
        return G

    # -- set-up diagnostics (AMS:374-404) -----------------------------------------------------------------------
    def diag_dense(self, rtol=1e-5, atol=1e-8):
        """(non-zeros, is_hermitian, is_complex_symmetric) of the resident dense matrix: np.count_nonzero + the two np.allclose
        tests of AMS:381-385 in one device pass."""
        nz, h, s_ = C.c_int64(), C.c_int32(), C.c_int32()
        self._check(self._lib.maus_diag_dense(self._h, float(rtol), float(atol), C.byref(nz), C.byref(h), C.byref(s_)))
        return int(nz.value), bool(h.value), bool(s_.value)

    def cond2_estimate(self, power_iters=40, inverse_iters=8, start=None):
        """(sigma_max, sigma_min, lu_status) of the resident dense matrix (replaces the full SVD behind np.linalg.cond, AMS:400)"""
        a, b, st = C.c_double(), C.c_double(), C.c_int32()
        start = None if start is None else _as_c128(start, (self.n,))
        self._check(self._lib.maus_cond2_estimate(self._h, int(power_iters), int(inverse_iters), _dp(start), C.byref(a),
                                                  C.byref(b), C.byref(st)))
        return a.value, b.value, int(st.value)

    # -- SVD power-sweep branch (AMS:227-255, 300-301) -------------------------------------------------------
    def svd_set_matrix(self, A):
        A = _as_c128(A)
        if A.ndim != 2:
            raise ValueError("2-D matrix required")
        self.svd_shape = A.shape
        self._check(self._lib.maus_svd_set_matrix(self._h, A.shape[0], A.shape[1], _dp(A)))

    def svd_step(self, U, V):
        """U [C][rows], V [C][cols] complex128 C-contiguous, updated in place. Returns dict(sigma, resid, status)."""
        C_ = U.shape[0]
        rows, cols = self.svd_shape
        if U.dtype != _c128 or V.dtype != _c128 or not U.flags.c_contiguous or not V.flags.c_contiguous \
                or U.shape != (C_, rows) or V.shape != (C_, cols):
            raise ValueError("U [C][rows] / V [C][cols] must be C-contiguous complex128")
        sigma = np.empty(C_, dtype=np.float64); resid = np.empty(C_, dtype=np.float64); status = np.empty(C_, dtype=np.int32)
        self._check(self._lib.maus_svd_step(self._h, C_, _dp(U), _dp(V), _dp(sigma), _dp(resid),
                                            status.ctypes.data_as(C.POINTER(C.c_int32))))
        return dict(sigma=sigma, resid=resid, status=status)

    def svd_residual(self, U, V, sigma):
        U = _as_c128(U); V = _as_c128(V)
        sigma = np.ascontiguousarray(np.atleast_1d(sigma), dtype=np.float64)
        resid = np.empty(U.shape[0], dtype=np.float64)
        self._check(self._lib.maus_svd_residual(self._h, U.shape[0], _dp(U), _dp(V), _dp(sigma), _dp(resid)))
        return resid

    # -- fused generation step --------------------------------------------------------------------------------
    def step(self, problem_type, alpha, psi, V=None, rng_key=None, method=_abi.METHOD_LU, use_jacobi=None,
             res_slot=_abi.SLOT_CURRENT, out=None):
        """One attempt-0 generation for C candidates (AMS:574-576 fast path).  ``V`` ([C][n] complex128) travels
        host->device->host when given and is updated IN PLACE; ``V=None`` keeps the vectors resident.
        Returns dict(lam, resid, mixnorm, status, iters)."""
        alpha = np.ascontiguousarray(np.atleast_1d(alpha), dtype=np.float64)
        C_ = alpha.shape[0]
        psi = np.ascontiguousarray(np.atleast_1d(psi), dtype=np.float64)
        keys = None if rng_key is None else np.ascontiguousarray(np.atleast_1d(rng_key), dtype=np.uint64)
        jac = None if use_jacobi is None else np.ascontiguousarray(np.atleast_1d(use_jacobi), dtype=np.uint8)
        if V is not None:
            if V.dtype != _c128 or not V.flags.c_contiguous or V.shape != (C_, self.n):
                raise ValueError("V must be a C-contiguous complex128 [C][n] array (updated in place)")
        if out is None:
            out = dict(lam=np.empty(C_, dtype=_c128), resid=np.empty(C_, dtype=np.float64),
                       mixnorm=np.empty(C_, dtype=np.float64), status=np.empty(C_, dtype=np.int32),
                       iters=np.empty(C_, dtype=np.int32))
        self._check(self._lib.maus_step(
            self._h, C_, int(problem_type), int(method), _dp(V), _dp(alpha), _dp(psi),
            None if keys is None else keys.ctypes.data_as(C.POINTER(C.c_uint64)),
            None if jac is None else jac.ctypes.data_as(C.POINTER(C.c_uint8)), int(res_slot),
            _dp(out["lam"]), _dp(out["resid"]), _dp(out["mixnorm"]),
            out["status"].ctypes.data_as(C.POINTER(C.c_int32)), out["iters"].ctypes.data_as(C.POINTER(C.c_int32))))
        self.generation += 1
        return out
