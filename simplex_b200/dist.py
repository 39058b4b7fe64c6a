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
        self._L, self.rank, self.world, self.group = L, rank, world, group
        self._ctx = C.c_void_p()
        rc = L.sb200_create(C.byref(self._ctx), C.byref(o))
        if rc != _abi.OK:
            raise Sb200Error(rc, L.sb200_last_error(None).decode())
        self.m = self.N = 0
        self.plan = None

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._L.sb200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != _abi.OK:
            raise Sb200Error(rc, self._L.sb200_last_error(self._ctx).decode())

    def load_shard(self, m, N, A_cols, b, c):
        """A_cols: this rank's block of columns, (m, ncols) Fortran order (see shard_plan)."""
        self.plan = shard_plan(m, N, self.world)
        me = self.plan[self.rank]
        A_cols = np.asfortranarray(A_cols, dtype=np.float64)
        assert A_cols.shape == (m, me["ncols"]), (A_cols.shape, me)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        self._check(self._L.sb200_load_shard(self._ctx, m, N, me["col0"], me["ncols"],
                                             A_cols.ctypes.data_as(C.c_void_p), m, b.ctypes.data_as(C.c_void_p),
                                             c.ctypes.data_as(C.c_void_p), 0))
        self.m, self.N = m, N
        # exchange the window handles and map the peers
        h = (C.c_ubyte * 64)()
        self._check(self._L.sb200_shard_window_handle(self._ctx, h))
        if self.world > 1:
            allh = exchange_handles(bytes(h), self.group)
            buf = (C.c_ubyte * (64 * self.world)).from_buffer_copy(allh)
            self._check(self._L.sb200_shard_connect(self._ctx, buf))

    def simplex(self, basis):
        """Collective: every rank calls it with the same lists. Returns loop statistics with the
        objective summed over the ranks; basis holds the final (global) lists."""
        import torch.distributed as dist
        diag = np.zeros(self.m)
        self._check(self._L.sb200_shard_prepare(self._ctx, _lp(basis.basic), basis.basic.size, _lp(basis.non_basic),
                                                basis.non_basic.size, _dp(diag)))
        if self.world > 1:
            diag = sum_over_ranks(diag, self.group)
        self._check(self._L.sb200_shard_finish_prepare(self._ctx, _dp(diag)))
        if self.world > 1:
            dist.barrier(group=_host_group(self.group))   # every window is reset before anyone publishes
        res = _abi.Result()
        self._check(self._L.sb200_shard_run(self._ctx, _lp(basis.basic), _lp(basis.non_basic), C.byref(res)))
        obj = np.array([res.objective])
        if self.world > 1:
            obj = sum_over_ranks(obj, self.group)
        return {"term": TERM_NAMES.get(res.term, res.term), "iterations": res.iterations, "pivots": res.pivots,
                "objective": float(obj[0]), "loop_seconds": res.loop_seconds}

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
