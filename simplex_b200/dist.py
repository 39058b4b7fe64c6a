"""One wide LP over the GPUs of one box (BASELINE configs[3]): A column-sharded,
B^-1 row-sharded, one process per GPU. torch.distributed is only plumbing here — it carries the
64-byte IPC handles of the exchange windows, the one-off sum of the crash-basis diagonal, a
barrier and the final gather of x_B. Every per-pivot exchange (two arg-min all-reduces, the
entering column and the pivot row) is done by the CUDA kernels themselves over NVLink
(simplex_b200/csrc/sb200_sharded.cuh).

The reference has no counterpart: it is single-threaded."""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import Sb200Error, TERM_NAMES
from .solver import DUALS, ENGINES, PRICING, _dp, _lp

DUAL_BLOCK_ROWS = 32   # csrc/sb200_fused.cuh
MAX_WORLD = 8          # csrc/sb200_sharded.cuh


def _guards_at_close(L, ctx):
    """SB200_GUARD=1: count the damaged guard zones of a context that is about to be destroyed."""
    if not _abi.GUARD_CHECK_AT_CLOSE:
        return 0
    n, bad = C.c_int64(0), C.c_int64(0)
    if L.sb200_debug_check_guards(ctx, C.byref(n), C.byref(bad)) != _abi.OK:
        return 0
    _abi.GUARD_STATS["contexts"] += 1
    _abi.GUARD_STATS["buffers"] += int(n.value)
    _abi.GUARD_STATS["damaged"] += int(bad.value)
    return int(bad.value)


def shard_plan(m, N, world):
    """The partition the library uses (sb200_load_shard): cyclic columns {rank + k*world} (keeps
    structural and slack columns balanced), row blocks of round_up(ceil(m/world), 32). Returns a
    list of dicts, one per rank; the columns of rank r are range(col0, N, col_stride)."""
    if not 1 <= world <= MAX_WORLD:
        raise ValueError("world must be 1..%d" % MAX_WORLD)
    rpr = -(-(-(-m // world)) // DUAL_BLOCK_ROWS) * DUAL_BLOCK_ROWS
    plan = []
    for r in range(world):
        row0 = min(m, r * rpr)
        plan.append({"rank": r, "col0": r, "col_stride": world, "ncols": (-(-(N - r) // world)) if N > r else 0,
                     "row0": row0, "nrows": max(0, min(rpr, m - row0)), "rows_per_rank": rpr})
    return plan


def owner_of_column(col, N, world):
    return col % world


def owner_of_row(row, m, world):
    rpr = -(-(-(-m // world)) // DUAL_BLOCK_ROWS) * DUAL_BLOCK_ROWS
    return row // rpr


def combine_candidates(cands):
    """Host restatement of the cross-rank arg-min the kernels perform: lexicographic minimum of
    (key, cls, idx) over candidates with idx >= 0 (ties to the lowest position / row)."""
    best = None
    for key, cls, idx in cands:
        if idx < 0:
            continue
        if best is None or (key, cls, idx) < best:
            best = (key, cls, idx)
    return best if best is not None else (0.0, 0, -1)


# ------------------------------------------------------------------ torch.distributed plumbing
def _host_group(group):
    """Collectives on host bytes need a CPU-capable backend; under NCCL make a gloo twin."""
    import torch.distributed as dist
    if dist.get_backend(group) == "gloo":
        return group
    if not hasattr(_host_group, "_twin"):
        _host_group._twin = dist.new_group(backend="gloo")
    return _host_group._twin


def exchange_handles(handle64, group=None):
    """All-gather of the 64-byte window handles. Returns world*64 bytes, rank-ordered."""
    import torch
    import torch.distributed as dist
    g = _host_group(group)
    world = dist.get_world_size(g)
    mine = torch.frombuffer(bytearray(handle64), dtype=torch.uint8).clone()
    assert mine.numel() == 64
    out = [torch.zeros(64, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(out, mine, group=g)
    return b"".join(bytes(t.numpy().tobytes()) for t in out)


def sum_over_ranks(vec, group=None):
    """Element-wise sum of a float64 host vector over the ranks (each entry is non-zero on exactly
    one rank in our use, so the sum is exact)."""
    import torch
    import torch.distributed as dist
    g = _host_group(group)
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=g)
    return t.numpy()


def worst_rc(rc, group=None):
    """MAX of an sb200 return code over the ranks: a rank whose call failed (say ERR_UNSUPPORTED from
    sb200_shard_prepare) must not leave its peers blocked in the next collective or inside the kernel's exchanges —
    every rank learns that somebody failed and raises."""
    import torch
    import torch.distributed as dist
    g = _host_group(group)
    t = torch.tensor([int(rc)], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=g)
    return int(t.item())


def gather_rows(local, plan, group=None):
    """Concatenate the row-sharded x_B of all ranks (rank order = row order)."""
    import torch
    import torch.distributed as dist
    g = _host_group(group)
    world = dist.get_world_size(g)
    width = max(p["nrows"] for p in plan)
    buf = torch.zeros(width, dtype=torch.float64)
    buf[:local.size] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    out = [torch.zeros(width, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(out, buf, group=g)
    return np.concatenate([out[r].numpy()[:plan[r]["nrows"]] for r in range(world)])


class ShardedSimplex:
    """One rank of the sharded solver (one process per GPU)."""

    def __init__(self, rank, world, pricing="most_neg", device=None, eps=1e-7, iter_limit=10 ** 9,
                 trace_capacity=0, chunk_iters=256, stream=None, group=None):
        L = _abi.lib()
        o = _abi.default_opts()
        o.device = rank if device is None else device
        o.pricing = PRICING[pricing] if isinstance(pricing, str) else int(pricing)
        o.engine = ENGINES["persistent"]
        o.dual_mode = DUALS["update"]
        o.eps = eps
        o.iter_limit = iter_limit
        o.trace_capacity = trace_capacity
        o.chunk_iters = chunk_iters
        o.stream = stream
        o.rank, o.world = rank, world
        self.iter_limit = iter_limit
        self._L, self.rank, self.world, self.group = L, rank, world, group
        self._ctx = C.c_void_p()
        rc = L.sb200_create(C.byref(self._ctx), C.byref(o))
        if rc != _abi.OK:
            raise Sb200Error(rc, L.sb200_last_error(None).decode())
        self.m = self.N = 0
        self.plan = None

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            damaged = _guards_at_close(self._L, self._ctx)
            self._L.sb200_destroy(self._ctx)
            self._ctx = C.c_void_p()
            if damaged:
                raise Sb200Error(2, "%d device buffer(s) of this context have a damaged guard zone" % damaged)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != _abi.OK:
            raise Sb200Error(rc, self._L.sb200_last_error(self._ctx).decode())

    def _check_all(self, rc, what):
        """Collective form of _check: every rank raises when any rank failed."""
        worst = worst_rc(rc, self.group) if self.world > 1 else rc
        if rc != _abi.OK:
            raise Sb200Error(rc, self._L.sb200_last_error(self._ctx).decode())
        if worst != _abi.OK:
            raise Sb200Error(worst, "%s failed on another rank" % what)

    def load_shard(self, m, N, A_cols, b, c, ub=None):
        """A_cols: this rank's block of columns, (m, ncols) Fortran order (see shard_plan). ub: None or the N upper bounds
        (replicated; DBL_MAX = none): bounded ratio test and bound flips on the shards."""
        self.plan = shard_plan(m, N, self.world)
        me = self.plan[self.rank]
        A_cols = np.asfortranarray(A_cols, dtype=np.float64)
        assert A_cols.shape == (m, me["ncols"]), (A_cols.shape, me)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        self._check_all(self._L.sb200_load_shard(self._ctx, m, N, me["col0"], me["ncols"],
                                                 A_cols.ctypes.data_as(C.c_void_p), m, b.ctypes.data_as(C.c_void_p),
                                                 c.ctypes.data_as(C.c_void_p), 0), "sb200_load_shard")
        self.m, self.N = m, N
        # exchange the window handles and map the peers
        h = (C.c_ubyte * 64)()
        self._check_all(self._L.sb200_shard_window_handle(self._ctx, h), "sb200_shard_window_handle")
        if self.world > 1:
            allh = exchange_handles(bytes(h), self.group)
            buf = (C.c_ubyte * (64 * self.world)).from_buffer_copy(allh)
            self._check_all(self._L.sb200_shard_connect(self._ctx, buf), "sb200_shard_connect")
        if ub is not None:
            ub = np.ascontiguousarray(ub, dtype=np.float64)
            assert ub.size == N
            self._check_all(self._L.sb200_shard_set_bounds(self._ctx, _dp(ub)), "sb200_shard_set_bounds")

    def set_costs(self, c):
        """Phase I -> Phase II: new costs for the first len(c) columns, the trailing (artificial) ones are dropped;
        A stays where it is on every rank. Collective."""
        c = np.ascontiguousarray(c, dtype=np.float64)
        self._check_all(self._L.sb200_shard_set_costs(self._ctx, c.size, _dp(c)), "sb200_shard_set_costs")
        self.N = c.size

    def _owned_basic_columns(self):
        n = C.c_int64(0)
        self._check(self._L.sb200_shard_get_basic_columns(self._ctx, None, None, 0, C.byref(n)))
        cap = max(int(n.value), 1)
        cols = np.zeros((cap, self.m))
        pos = np.zeros(cap, dtype=np.int64)
        self._check(self._L.sb200_shard_get_basic_columns(self._ctx, _dp(cols), _lp(pos), cap, C.byref(n)))
        return cols[:int(n.value)], pos[:int(n.value)]

    def reinvert(self):
        """Collective. B^-1 and x_B of the CURRENT basis from scratch: every rank contributes the basic columns it owns,
        every rank inverts the whole basis matrix with the single-GPU Gauss-Jordan (same bits everywhere) and keeps its
        rows. Used for a general start basis and for re-inversion between runs."""
        cols, pos = self._owned_basic_columns()
        AB = np.zeros((self.m, self.m), order="F")
        AB[:, pos] = cols.T
        if self.world > 1:
            AB = np.asfortranarray(sum_over_ranks(AB.ravel(order="F"), self.group).reshape((self.m, self.m), order="F"))
        self._check_all(self._L.sb200_shard_invert(self._ctx, _dp(AB)), "sb200_shard_invert")

    def simplex(self, basis, carry_over=False, general=False, reinvert_period=0, hygiene_tol=None, hygiene_period=0, b=None):
        """Collective: every rank calls it with the same lists. Returns loop statistics with the
        objective summed over the ranks; basis holds the final (global) lists. carry_over: start from the basis
        the previous call ended on, with its row-sharded B^-1 and x_B (the Phase-II call after set_costs).
        general: any start basis (inverted as reinvert() does); the default expects a unit (slack / artificial) basis.
        reinvert_period = P > 0: B^-1 and x_B are rebuilt every P iterations; hygiene_tol (with b, the right-hand side):
        every hygiene_period iterations the primal residual is measured over the ranks and a re-inversion follows only
        when max|b - A_B x_B| / (1 + max|b|) exceeds it."""
        segment, between_segments, self.hygiene = None, None, {"checks": 0, "refreshes": 0, "worst": 0.0}
        if reinvert_period > 0:
            segment = reinvert_period

            def between_segments():
                self.reinvert()
                self.hygiene["refreshes"] += 1
        elif hygiene_tol is not None and hygiene_period > 0:
            assert b is not None, "hygiene_tol needs the right-hand side b"
            segment = hygiene_period
            scale = 1.0 + float(np.abs(b).max())

            def between_segments():
                r = self.primal_residual(b) / scale
                self.hygiene["checks"] += 1
                self.hygiene["worst"] = max(self.hygiene["worst"], r)
                if r > hygiene_tol:
                    self.reinvert()
                    self.hygiene["refreshes"] += 1
        if general:
            self._check_all(self._L.sb200_shard_prepare_general(self._ctx, _lp(basis.basic), basis.basic.size,
                                                                _lp(basis.non_basic), basis.non_basic.size),
                            "sb200_shard_prepare_general")
            self.reinvert()
        elif carry_over:
            self._check_all(self._L.sb200_shard_continue(self._ctx, _lp(basis.basic), basis.basic.size, _lp(basis.non_basic),
                                                         basis.non_basic.size), "sb200_shard_continue")
        else:
            diag = np.zeros(self.m)
            self._check_all(self._L.sb200_shard_prepare(self._ctx, _lp(basis.basic), basis.basic.size, _lp(basis.non_basic),
                                                        basis.non_basic.size, _dp(diag)), "sb200_shard_prepare")
            if self.world > 1:
                diag = sum_over_ranks(diag, self.group)
            # (the collective check doubles as the barrier behind which every window is reset before anyone publishes)
            self._check_all(self._L.sb200_shard_finish_prepare(self._ctx, _dp(diag)), "sb200_shard_finish_prepare")
        res = _abi.Result()
        loop_s = 0.0
        if segment is None:
            self._check_all(self._L.sb200_shard_run(self._ctx, _lp(basis.basic), _lp(basis.non_basic), C.byref(res)),
                            "sb200_shard_run")
            loop_s = res.loop_seconds
        else:
            # the run in segments of `segment` iterations; between two segments the hook decides about a re-inversion
            limit = 0
            while True:
                limit = min(self.iter_limit, limit + segment)
                self._check_all(self._L.sb200_shard_resume(self._ctx, limit), "sb200_shard_resume")
                self._check_all(self._L.sb200_shard_run(self._ctx, _lp(basis.basic), _lp(basis.non_basic), C.byref(res)),
                                "sb200_shard_run")
                loop_s += res.loop_seconds
                if TERM_NAMES.get(res.term) != "iter_limit" or res.iterations >= self.iter_limit:
                    break
                between_segments()
            res.loop_seconds = loop_s
        obj = np.array([res.objective])
        if self.world > 1:
            obj = sum_over_ranks(obj, self.group)
        return {"term": TERM_NAMES.get(res.term, res.term), "iterations": res.iterations, "pivots": res.pivots,
                "flips": res.bound_flips, "objective": float(obj[0]), "loop_seconds": res.loop_seconds}

    def xB_local(self):
        row0, nrows = C.c_int64(0), C.c_int64(0)
        self._check(self._L.sb200_shard_get_xB_local(self._ctx, None, C.byref(row0), C.byref(nrows)))
        out = np.zeros(max(int(nrows.value), 1))
        self._check(self._L.sb200_shard_get_xB_local(self._ctx, _dp(out), C.byref(row0), C.byref(nrows)))
        return out[:int(nrows.value)], int(row0.value)

    def xB(self):
        loc, _ = self.xB_local()
        if self.world == 1:
            return loc
        return gather_rows(loc, self.plan, self.group)

    def primal_residual(self, b):
        """max_i |b_i - (A_B x_B)_i| of the basis the last run ended on (collective): every rank multiplies the
        basic columns it owns with the gathered x_B, the host adds the ranks up. b: the right-hand side."""
        x = np.ascontiguousarray(self.xB(), dtype=np.float64)
        part = np.zeros(self.m)
        self._check_all(self._L.sb200_shard_residual_partial(self._ctx, _dp(x), _dp(part)), "sb200_shard_residual_partial")
        if self.world > 1:
            part = sum_over_ranks(part, self.group)
        return float(np.abs(np.asarray(b, dtype=np.float64) - part).max())

    def trace(self):
        n = C.c_int64(0)
        self._check(self._L.sb200_get_trace(self._ctx, None, 0, C.byref(n)))
        cap = max(int(n.value), 1)
        buf = (_abi.Pivot * cap)()
        self._check(self._L.sb200_get_trace(self._ctx, buf, cap, C.byref(n)))
        arr = np.frombuffer(buf, dtype=np.int32).reshape(cap, 4)
        return arr[:min(int(n.value), cap)].copy(), int(n.value)

    def phase_cycles(self):
        out = np.zeros(8, dtype=np.int64)
        self._check(self._L.sb200_get_phase_cycles(self._ctx, _lp(out)))
        return out

    def launch_count(self):
        return int(self._L.sb200_launch_count(self._ctx))



class LocalShardedSimplex:
    """Every rank of the sharded solver in ONE process: one context and, while the loop runs, one host thread per
    rank (sb200_shard_connect_local). The ranks may sit on different GPUs (peer access) or SHARE one GPU, each on
    its share of the SMs — which is how the independence of the pivot sequence from the number of ranks is
    checked on a one-GPU box. No torch.distributed anywhere: the host-side sums are plain numpy."""

    def __init__(self, world, devices=None, pricing="most_neg", eps=1e-7, iter_limit=10 ** 9, trace_capacity=0,
                 chunk_iters=256):
        if not 1 <= world <= MAX_WORLD:
            raise ValueError("world must be 1..%d" % MAX_WORLD)
        self._L = _abi.lib()
        self.world = world
        self.iter_limit = iter_limit
        self.devices = list(devices) if devices is not None else [0] * world
        self._ctxs = []
        for r in range(world):
            o = _abi.default_opts()
            o.device = self.devices[r]
            o.pricing = PRICING[pricing] if isinstance(pricing, str) else int(pricing)
            o.engine = ENGINES["persistent"]
            o.dual_mode = DUALS["update"]
            o.eps, o.iter_limit, o.trace_capacity, o.chunk_iters = eps, iter_limit, trace_capacity, chunk_iters
            o.rank, o.world = r, world
            ctx = C.c_void_p()
            rc = self._L.sb200_create(C.byref(ctx), C.byref(o))
            if rc != _abi.OK:
                self.close()
                raise Sb200Error(rc, self._L.sb200_last_error(None).decode())
            self._ctxs.append(ctx)
        self.m = self.N = 0
        self.plan = None

    def close(self):
        damaged = 0
        for ctx in self._ctxs:
            if ctx.value:
                damaged += _guards_at_close(self._L, ctx) or 0
                self._L.sb200_destroy(ctx)
        self._ctxs = []
        if damaged:
            raise Sb200Error(2, "%d device buffer(s) of these contexts have a damaged guard zone" % damaged)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, r, rc):
        if rc != _abi.OK:
            raise Sb200Error(rc, "rank %d: %s" % (r, self._L.sb200_last_error(self._ctxs[r]).decode()))

    def load(self, A, b, c, ub=None):
        """A: the whole (m, N) tableau; rank r gets its cyclic column set A[:, r::world]. ub: None or N upper bounds."""
        m, N = A.shape
        self.plan = shard_plan(m, N, self.world)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        for r, ctx in enumerate(self._ctxs):
            cols = np.asfortranarray(A[:, r::self.world], dtype=np.float64)
            me = self.plan[r]
            assert cols.shape == (m, me["ncols"])
            self._check(r, self._L.sb200_load_shard(ctx, m, N, me["col0"], me["ncols"], cols.ctypes.data_as(C.c_void_p), m,
                                                    b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), 0))
        arr = (C.c_void_p * self.world)(*[ctx.value for ctx in self._ctxs])
        for r, ctx in enumerate(self._ctxs):
            self._check(r, self._L.sb200_shard_connect_local(ctx, arr, self.world))
        if ub is not None:
            ub = np.ascontiguousarray(ub, dtype=np.float64)
            assert ub.size == N
            for r, ctx in enumerate(self._ctxs):
                self._check(r, self._L.sb200_shard_set_bounds(ctx, _dp(ub)))
        self.m, self.N = m, N

    def set_costs(self, c):
        c = np.ascontiguousarray(c, dtype=np.float64)
        for r, ctx in enumerate(self._ctxs):
            self._check(r, self._L.sb200_shard_set_costs(ctx, c.size, _dp(c)))
        self.N = c.size

    def primal_residual(self, b):
        x = np.ascontiguousarray(self.xB(), dtype=np.float64)
        tot = np.zeros(self.m)
        for r, ctx in enumerate(self._ctxs):
            part = np.zeros(self.m)
            self._check(r, self._L.sb200_shard_residual_partial(ctx, _dp(x), _dp(part)))
            tot += part
        return float(np.abs(np.asarray(b, dtype=np.float64) - tot).max())

    def reinvert(self):
        """B^-1 and x_B of the current basis from scratch on every rank (see ShardedSimplex.reinvert)."""
        AB = np.zeros((self.m, self.m), order="F")
        for r, ctx in enumerate(self._ctxs):
            n = C.c_int64(0)
            self._check(r, self._L.sb200_shard_get_basic_columns(ctx, None, None, 0, C.byref(n)))
            cap = max(int(n.value), 1)
            cols, pos = np.zeros((cap, self.m)), np.zeros(cap, dtype=np.int64)
            self._check(r, self._L.sb200_shard_get_basic_columns(ctx, _dp(cols), _lp(pos), cap, C.byref(n)))
            k = int(n.value)
            AB[:, pos[:k]] = cols[:k].T
        for r, ctx in enumerate(self._ctxs):
            self._check(r, self._L.sb200_shard_invert(ctx, _dp(AB)))

    def simplex(self, basis, carry_over=False, general=False, reinvert_period=0, hygiene_tol=None, hygiene_period=0, b=None):
        """See ShardedSimplex.simplex; here every rank's sb200_shard_run runs on its own host thread."""
        import threading
        self.hygiene = {"checks": 0, "refreshes": 0, "worst": 0.0}
        if general:
            for r, ctx in enumerate(self._ctxs):
                self._check(r, self._L.sb200_shard_prepare_general(ctx, _lp(basis.basic), basis.basic.size,
                                                                   _lp(basis.non_basic), basis.non_basic.size))
            self.reinvert()
        elif carry_over:
            for r, ctx in enumerate(self._ctxs):
                self._check(r, self._L.sb200_shard_continue(ctx, _lp(basis.basic), basis.basic.size, _lp(basis.non_basic),
                                                            basis.non_basic.size))
        else:
            diag = np.zeros(self.m)
            for r, ctx in enumerate(self._ctxs):
                part = np.zeros(self.m)
                self._check(r, self._L.sb200_shard_prepare(ctx, _lp(basis.basic), basis.basic.size, _lp(basis.non_basic),
                                                           basis.non_basic.size, _dp(part)))
                diag += part          # every entry is non-zero on exactly one rank
            for r, ctx in enumerate(self._ctxs):
                self._check(r, self._L.sb200_shard_finish_prepare(ctx, _dp(diag)))
        results = [_abi.Result() for _ in self._ctxs]
        lists = [(basis.basic.copy(), basis.non_basic.copy()) for _ in self._ctxs]

        def run_all():
            rcs = [None] * self.world

            def run(r):
                rcs[r] = self._L.sb200_shard_run(self._ctxs[r], _lp(lists[r][0]), _lp(lists[r][1]), C.byref(results[r]))

            threads = [threading.Thread(target=run, args=(r,)) for r in range(self.world)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            for r in range(self.world):
                self._check(r, rcs[r])

        segment = reinvert_period if reinvert_period > 0 else (hygiene_period if hygiene_tol is not None else 0)
        loop_s = 0.0
        if segment <= 0:
            run_all()
            loop_s = max(x.loop_seconds for x in results)
        else:
            limit = 0
            while True:
                limit = min(self.iter_limit, limit + segment)
                for r, ctx in enumerate(self._ctxs):
                    self._check(r, self._L.sb200_shard_resume(ctx, limit))
                run_all()
                loop_s += max(x.loop_seconds for x in results)
                if TERM_NAMES.get(results[0].term) != "iter_limit" or results[0].iterations >= self.iter_limit:
                    break
                if reinvert_period > 0:
                    self.reinvert()
                    self.hygiene["refreshes"] += 1
                else:
                    rr = self.primal_residual(b) / (1.0 + float(np.abs(b).max()))
                    self.hygiene["checks"] += 1
                    self.hygiene["worst"] = max(self.hygiene["worst"], rr)
                    if rr > hygiene_tol:
                        self.reinvert()
                        self.hygiene["refreshes"] += 1
        for r in range(1, self.world):
            assert np.array_equal(lists[r][0], lists[0][0]) and np.array_equal(lists[r][1], lists[0][1]), "ranks disagree"
            assert (results[r].term, results[r].iterations, results[r].pivots) == \
                   (results[0].term, results[0].iterations, results[0].pivots), "ranks disagree"
        basis.basic[:], basis.non_basic[:] = lists[0]
        res = results[0]
        return {"term": TERM_NAMES.get(res.term, res.term), "iterations": res.iterations, "pivots": res.pivots,
                "flips": res.bound_flips, "objective": float(sum(x.objective for x in results)), "loop_seconds": loop_s}

    def xB(self):
        parts = []
        for r, ctx in enumerate(self._ctxs):
            row0, nrows = C.c_int64(0), C.c_int64(0)
            self._check(r, self._L.sb200_shard_get_xB_local(ctx, None, C.byref(row0), C.byref(nrows)))
            out = np.zeros(max(int(nrows.value), 1))
            self._check(r, self._L.sb200_shard_get_xB_local(ctx, _dp(out), C.byref(row0), C.byref(nrows)))
            parts.append(out[:int(nrows.value)])
        return np.concatenate(parts)

    def trace(self, rank=0):
        ctx = self._ctxs[rank]
        n = C.c_int64(0)
        self._check(rank, self._L.sb200_get_trace(ctx, None, 0, C.byref(n)))
        cap = max(int(n.value), 1)
        buf = (_abi.Pivot * cap)()
        self._check(rank, self._L.sb200_get_trace(ctx, buf, cap, C.byref(n)))
        arr = np.frombuffer(buf, dtype=np.int32).reshape(cap, 4)
        return arr[:min(int(n.value), cap)].copy(), int(n.value)

    def phase_cycles(self, rank=0):
        out = np.zeros(8, dtype=np.int64)
        self._check(rank, self._L.sb200_get_phase_cycles(self._ctxs[rank], _lp(out)))
        return out

    def flip_flags(self, rank=0):
        out = np.zeros(self.N, dtype=np.uint8)
        self._check(rank, self._L.sb200_get_flip_flags(self._ctxs[rank], out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def b(self, rank=0):
        out = np.zeros(self.m)
        self._check(rank, self._L.sb200_get_b(self._ctxs[rank], _dp(out)))
        return out
