"""Candidate-population sharding across the GPUs of one box.

Candidates never read each other inside a step (AMS:574-576 touches only ``self`` + shared read-only inputs), so the
population shards with NO data-path collective: live candidate i -> rank ``i mod G``, the matrix is replicated, every
rank runs the fused CUDA step on its own shard.  The only exchange is one all-gather per generation of the updated
candidate records (+ vectors), which feeds the host-side diagnostics / pruning of the reference
(``_update_global_diagnostics`` AMS:424-475, ``_manage_candidates`` AMS:504-549) that every rank replays identically.

One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np

from . import _abi
from .population import step_population

# float64 record per candidate exchanged each generation
REC_FIELDS = ("lambda_re", "lambda_im", "residual", "prev_residual", "alpha_re", "alpha_is_complex", "stuck", "retries",
              "resets", "w", "state", "hist_added", "lambda_is_real", "rng_drawn")
NREC = len(REC_FIELDS)
assert NREC % 2 == 0          # the vector part of a gathered row starts on a complex128 boundary


class Shard:
    def __init__(self, rank=0, world=1, device=None, engine=None):
        """``engine``: a MausEngine set up with ``enable_row_sharding`` -- the per-generation exchange then runs through the
        library's own ``maus_gather`` (NCCL all-gather on the context's stream) instead of torch.distributed."""
        self.rank, self.world, self.device, self.engine = int(rank), int(world), device, engine

    def owned(self, live_count):
        """indices (into the live list) this rank steps: round-robin, re-balanced every generation"""
        return list(range(self.rank, live_count, self.world))

    # ---- collectives ----------------------------------------------------------------------------------------
    def _dist(self):
        import torch.distributed as dist
        return dist

    def all_gather_rows(self, local, max_rows, reuse=False):
        """all-gather a ragged [rows_r][cols] float64 block; returns list of per-rank arrays (padded rows dropped by
        the caller, who knows each rank's count from ``owned``).  ``reuse``: the result may live in a recycled page-locked
        buffer that stays valid until the call after the next one (``RowShardedOperator.gather``)."""
        import torch
        if self.world == 1:
            return [local]
        cols = local.shape[1]
        buf = np.zeros((max_rows, cols), dtype=np.float64)
        buf[:local.shape[0]] = local
        rs = getattr(self.engine, "rowshard", None) if self.engine is not None else None
        if rs is not None and rs.world == self.world:
            out = rs.gather(buf, pinned=reuse).reshape(self.world, max_rows, cols)
            return [out[r] for r in range(self.world)]
        dist = self._dist()
        t = torch.from_numpy(buf)
        if self.device is not None:
            t = t.to(self.device, non_blocking=False)
        out = torch.empty((self.world, max_rows, cols), dtype=torch.float64, device=t.device)
        try:
            dist.all_gather_into_tensor(out.view(-1), t.view(-1))
        except (RuntimeError, AttributeError):
            parts = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(parts, t)
            out = torch.stack(parts, 0)
        out = out.cpu().numpy()
        return [out[r] for r in range(self.world)]

    def all_reduce_max(self, value):
        import torch
        if self.world == 1:
            return float(value)
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self._dist().all_reduce(t, op=self._dist().ReduceOp.MAX)
        return float(t.item())

    def all_reduce_sum(self, value):
        import torch
        if self.world == 1:
            return float(value)
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self._dist().all_reduce(t, op=self._dist().ReduceOp.SUM)
        return float(t.item())

    def barrier(self):
        if self.world > 1:
            self._dist().barrier()


def _rng_fingerprint():
    """cheap identity of the two global host RNG streams the reference draws re-initialisations from (AMS:130-135, 260, 283)"""
    import random
    st = np.random.get_state()
    return (int(st[2]), int(st[1][0]), int(st[1][-1]), hash(random.getstate()[1][-3:]))


def _record(c, hist_before, rng_drawn=False):
    al = c.alpha_local_step
    lam = complex(c.lambda_k) if c.lambda_k is not None else complex(0.0, 0.0)
    lam_real = c.lambda_k is not None and not isinstance(c.lambda_k, (complex, np.complexfloating))
    return [lam.real, lam.imag, float(c.residual_k), float(c.prev_residual), complex(al).real,
            1.0 if isinstance(al, (complex, np.complexfloating)) else 0.0, float(c.stuck_counter),
            float(c.local_psi_retries_needed), float(c.num_resets), float(c.w_k), float(c.state.value),
            float(len(c.residual_history) - hist_before), 1.0 if lam_real else 0.0, 1.0 if rng_drawn else 0.0]


def _apply_record(c, rec, vec, eigen, State):
    if eigen:
        # the Hermitian shortcut stores a real float64 eigenvalue (AMS:171), the inverse-iteration branch a complex128
        c.lambda_k = np.float64(rec[0]) if rec[12] else np.complex128(complex(rec[0], rec[1]))
        c.v_k = vec
    else:
        c.x_k = vec
    c.residual_k = np.float64(rec[2])
    c.prev_residual = rec[3]
    c.alpha_local_step = np.complex128(rec[4]) if rec[5] else float(rec[4])
    c.stuck_counter = int(rec[6])
    c.local_psi_retries_needed = int(rec[7])
    c.num_resets = int(rec[8])
    c.w_k = float(rec[9])
    c.state = State(int(rec[10]))
    for _ in range(int(rec[11])):
        c.param_history.append(c.get_current_solution_params())
        c.residual_history.append(c.residual_k)


def step_population_sharded(candidates, M, b, strat_params, problem_knowledge, engine, shard):
    """Sharded equivalent of the loop AMS:574-576: each rank steps the live candidates it owns on its GPU, then one
    all-gather brings every replica of the population up to date.  Returns the number of live candidates."""
    if shard.world == 1:
        return step_population(candidates, M, b, strat_params, problem_knowledge, engine)
    if not candidates:
        return 0
    State = type(candidates[0]).State
    live = [c for c in candidates if c.state not in (State.CONVERGED, State.RETIRED)]
    if not live:
        return 0
    if live[0].problem_type.value not in (_abi.EIGENVALUE, _abi.SOLVE_LINEAR_SYSTEM):
        # SVD / other branches: not sharded (every rank steps its own replica of the whole population)
        return step_population(candidates, M, b, strat_params, problem_knowledge, engine)
    if _uses_row_sharding(live, M, problem_knowledge, engine):
        # the MATRIX is sharded, not the population: every rank passes all live candidates to the collective row-sharded step
        # and receives the full updated vectors; statuses (and with them the host RNG draws) are identical on every rank
        return step_population(candidates, M, b, strat_params, problem_knowledge, engine)
    n = live[0].N_diag
    eigen = live[0].problem_type.value == _abi.EIGENVALUE
    # candidates whose vector is a view into a recycled gather buffer and that have left the live set through host logic
    # (pruned / retired by the reference's _manage_candidates) keep their vector for good: give them their own copy now
    views = getattr(shard, "_views", None)
    if views is None:
        views = shard._views = {}
        hooks = getattr(engine, "_close_hooks", None)
        if hooks is not None:
            # the page-locked gather buffers die with the engine: every candidate still looking into one gets its own copy first
            hooks.append(lambda: _detach_views(shard))
    if views:
        live_ids = {id(c) for c in live}
        for key in [k for k in views if k not in live_ids]:
            c = views.pop(key)
            if eigen and c.v_k is not None:
                c.v_k = np.array(c.v_k, copy=True)
            elif not eigen and c.x_k is not None:
                c.x_k = np.array(c.x_k, copy=True)
    counts = [len(range(r, len(live), shard.world)) for r in range(shard.world)]
    mine = [live[i] for i in shard.owned(len(live))]
    hist_before = [len(c.residual_history) for c in mine]
    rng_before = _rng_fingerprint()
    if mine:
        step_population(mine, M, b, strat_params, problem_knowledge, engine)
    rng_drawn = _rng_fingerprint() != rng_before
    # record + vector travel together as one float64 row: [NREC | 2n]
    local = np.zeros((len(mine), NREC + 2 * n), dtype=np.float64)
    for k, c in enumerate(mine):
        local[k, :NREC] = _record(c, hist_before[k], rng_drawn)
        vec = c.v_k if eigen else c.x_k
        local[k, NREC:] = np.ascontiguousarray(vec, dtype=np.complex128).view(np.float64)
    gathered = shard.all_gather_rows(local, max(counts), reuse=True)
    done = (State.CONVERGED.value, State.RETIRED.value)
    for r in range(shard.world):
        if r == shard.rank:
            continue
        for k, i in enumerate(range(r, len(live), shard.world)):
            row = gathered[r][k]
            # a VIEW into the gathered block (16-byte aligned at column NREC): no per-candidate copy of the 16 n bytes.  The
            # block is one of two recycled page-locked buffers: a live candidate replaces its vector in the next generation,
            # i.e. before this buffer's turn comes again; a candidate that just left the live set keeps its vector for good
            # and gets its own copy.
            vec = row[NREC:].view(np.complex128)
            if int(row[10]) in done:
                vec = vec.copy()
                views.pop(id(live[i]), None)
            else:
                views[id(live[i])] = live[i]
            _apply_record(live[i], row[:NREC], vec, eigen, State)
    for c in mine:
        views.pop(id(c), None)               # stepped here: the vector is this rank's own array again
    _resync_host_rng(gathered, counts)
    return len(live)


def _detach_views(shard):
    """give every candidate whose vector is a view into a recycled gather buffer its own copy (engine shutdown)"""
    views = getattr(shard, "_views", None) or {}
    for c in list(views.values()):
        if getattr(c, "v_k", None) is not None and c.v_k.base is not None:
            c.v_k = np.array(c.v_k, copy=True)
        if getattr(c, "x_k", None) is not None and c.x_k.base is not None:
            c.x_k = np.array(c.x_k, copy=True)
    views.clear()


def _resync_host_rng(gathered, counts):
    """Re-initialisations (AMS:260, 283, 293) draw from the OWNING rank's global numpy / random streams only, so after the first
    failure or collapse the replicas' streams differ and the reference's ``_manage_candidates`` (AMS:525-549: ``random.choice``,
    ``np.random.rand``) would spawn different candidates per rank.  When any rank drew this generation, every rank re-seeds
    both streams from a digest of the gathered records -- data all ranks hold bit-identically -- so the replicas stay equal."""
    drew = any(counts[r] and gathered[r][:counts[r], REC_FIELDS.index("rng_drawn")].any() for r in range(len(gathered)))
    if not drew:
        return False
    import hashlib
    import random
    h = hashlib.sha256()
    for r in range(len(gathered)):
        h.update(np.ascontiguousarray(gathered[r][:counts[r], :NREC]).tobytes())
    seed = int.from_bytes(h.digest()[:8], "little")
    np.random.seed(seed % (2 ** 32))
    random.seed(seed)
    return True


def _uses_row_sharding(live, M, problem_knowledge, engine):
    """mirror of the routing test in population.step_population"""
    from .constants import LU_MAX_N
    from .population import _is_sparse
    rs = getattr(engine, "rowshard", None)
    pref = problem_knowledge.get('local_solver_preference', 'direct_solve')
    hermitian_eigen = live[0].problem_type.value == _abi.EIGENVALUE and bool(problem_knowledge.get('is_hermitian', False))
    return (rs is not None and not hermitian_eigen and _is_sparse(M) and (pref == 'iterative_gmres' or live[0].N_diag > LU_MAX_N)
            and all(c.problem_matrix is M for c in live))


def gather_energy_and_best(shard, resid, lam, vectors, best_index=None):
    """The per-generation exchange of the benchmark loop: all-gather (residual, lambda) of every candidate and the
    arg-min-residual eigenpair.  resid [C] f64, lam [C] c128, vectors [C][n] c128 (host).  Returns
    (all_resid [G*C], all_lam [G*C], best_vec [n], best_rank)."""
    C_ = resid.shape[0]
    n = vectors.shape[1]
    k = (int(np.argmin(resid)) if C_ else 0) if best_index is None else int(best_index)
    local = np.zeros((1, 3 * C_ + 2 * n), dtype=np.float64)
    local[0, :C_] = resid
    local[0, C_:3 * C_] = np.ascontiguousarray(lam, dtype=np.complex128).view(np.float64)
    local[0, 3 * C_:] = vectors[k].view(np.float64)
    rows = shard.all_gather_rows(local, 1)
    all_resid = np.concatenate([r[0, :C_] for r in rows])
    all_lam = np.concatenate([r[0, C_:3 * C_].copy().view(np.complex128) for r in rows])
    best_rank = int(np.argmin([r[0, :C_].min() for r in rows]))
    best_vec = rows[best_rank][0, 3 * C_:].copy().view(np.complex128)
    return all_resid, all_lam, best_vec, best_rank
