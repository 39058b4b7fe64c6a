"""Host-side Python mirror of the seam the reference exposes as `Simplex::simplex(Basis&)`
(reference src/Simplex.h:25) on top of the C ABI (include/simplex_b200.h). Names follow the
reference: a `Basis` holds the ordered basic / non-basic column lists (src/Basis.h:18-34),
`DeviceSimplex.simplex(basis)` mutates it in place and returns the loop statistics.

Everything numerical happens in libsimplex_b200.so on the GPU. There is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import (DUAL_RECOMPUTE, DUAL_UPDATE, ENGINE_PERSISTENT, ENGINE_STAGED, ENGINE_STREAMING, PRICE_FIRST_ELIGIBLE,
                   PRICE_MOST_NEGATIVE, TERM_NAMES, Sb200Error)

PRICING = {"first": PRICE_FIRST_ELIGIBLE, "stock": PRICE_FIRST_ELIGIBLE, "bland": PRICE_FIRST_ELIGIBLE,
           "most_neg": PRICE_MOST_NEGATIVE, "dantzig": PRICE_MOST_NEGATIVE}
ENGINES = {"persistent": ENGINE_PERSISTENT, "staged": ENGINE_STAGED, "streaming": ENGINE_STREAMING}
DUALS = {"recompute": DUAL_RECOMPUTE, "update": DUAL_UPDATE}


class Basis:
    """Ordered index lists, as simplex::Basis (reference src/Basis.h:18-34). Order matters:
    entering candidates are keyed by position in `non_basic`, leaving ones by row in `basic`."""

    def __init__(self, basic, non_basic):
        self.basic = np.ascontiguousarray(basic, dtype=np.int64).copy()
        self.non_basic = np.ascontiguousarray(non_basic, dtype=np.int64).copy()

    def copy(self):
        return Basis(self.basic, self.non_basic)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _lp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class DeviceSimplex:
    """One solver context on one GPU (sb200_ctx)."""

    def __init__(self, pricing="first", engine="persistent", dual="update", device=0, eps=1e-7,
                 iter_limit=10 ** 9, trace_capacity=0, chunk_iters=64, dual_refresh=0, stream=None,
                 reinvert_period=0, hygiene_tol=None, hygiene_period=None, hygiene_rows=None):
        L = _abi.lib()
        o = _abi.default_opts()
        o.device = device
        o.pricing = PRICING[pricing] if isinstance(pricing, str) else int(pricing)
        o.engine = ENGINES[engine] if isinstance(engine, str) else int(engine)
        o.dual_mode = DUALS[dual] if isinstance(dual, str) else int(dual)
        o.dual_refresh = dual_refresh
        o.reinvert_period = reinvert_period
        if hygiene_tol is not None:
            o.hygiene_tol = hygiene_tol
        if hygiene_period is not None:
            o.hygiene_period = hygiene_period
        if hygiene_rows is not None:
            o.hygiene_rows = hygiene_rows
        o.eps = eps
        o.iter_limit = iter_limit
        o.trace_capacity = trace_capacity
        o.chunk_iters = chunk_iters
        o.stream = stream
        self._L = L
        self._ctx = C.c_void_p()
        rc = L.sb200_create(C.byref(self._ctx), C.byref(o))
        if rc != _abi.OK:
            raise Sb200Error(rc, L.sb200_last_error(None).decode())
        self.m = self.N = 0
        self._keep = None

    # -- lifecycle
    def check_guards(self):
        """(buffers, damaged): the guard zones around every device buffer of this context (out-of-bounds writes)."""
        n, bad = C.c_int64(0), C.c_int64(0)
        self._check(self._L.sb200_debug_check_guards(self._ctx, C.byref(n), C.byref(bad)))
        return int(n.value), int(bad.value)

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            damaged = None
            if _abi.GUARD_CHECK_AT_CLOSE:
                try:
                    _abi.GUARD_STATS["contexts"] += 1
                    n, damaged = self.check_guards()
                    _abi.GUARD_STATS["buffers"] += n
                    _abi.GUARD_STATS["damaged"] += damaged
                except Sb200Error:
                    damaged = None
            self._L.sb200_destroy(self._ctx)
            self._ctx = C.c_void_p()
            if damaged:
                raise Sb200Error(2, "%d device buffer(s) of this context have a damaged guard zone (out-of-bounds write)" % damaged)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != _abi.OK:
            raise Sb200Error(rc, self._L.sb200_last_error(self._ctx).decode())

    # -- data in
    def load(self, A, b, c, ub=None):
        """A: (m, N) array; the column-major copy handed to the device is what the reference's
        loop reads as m_lp.matrix_A (post-scaling). ub: None or per-column upper bounds."""
        A = np.asfortranarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        m, N = A.shape
        if b.size != m or c.size != N:
            raise ValueError("b must have m entries and c N entries")
        ubp = None
        if ub is not None:
            ub = np.ascontiguousarray(ub, dtype=np.float64)
            if ub.size != N:
                raise ValueError("ub must have N entries")
            ubp = _dp(ub)
        self._check(self._L.sb200_load(self._ctx, m, N, _dp(A), m, _dp(b), _dp(c), ubp))
        self.m, self.N = m, N

    def load_device(self, m, N, dA_ptr, lda, db_ptr, dc_ptr, dub_ptr=None):
        self._check(self._L.sb200_load_device(self._ctx, m, N, dA_ptr, lda, db_ptr, dc_ptr, dub_ptr))
        self.m, self.N = m, N

    def set_costs(self, c):
        c = np.ascontiguousarray(c, dtype=np.float64)
        self._check(self._L.sb200_set_costs(self._ctx, c.size, _dp(c)))
        self.N = c.size

    # -- the seam
    def simplex(self, basis):
        """Simplex::simplex(Basis&): runs to optimal / unbounded / iteration limit on the GPU,
        leaves the final lists in `basis`. Returns a dict of loop statistics."""
        res = _abi.Result()
        self._check(self._L.sb200_run(self._ctx, _lp(basis.basic), basis.basic.size, _lp(basis.non_basic),
                                      basis.non_basic.size, C.byref(res)))
        return {"term": TERM_NAMES.get(res.term, res.term), "iterations": res.iterations, "pivots": res.pivots,
                "flips": res.bound_flips, "objective": res.objective, "loop_seconds": res.loop_seconds}

    # -- read-out
    def _vec(self, fn, n, dtype=np.float64):
        out = np.zeros(n, dtype=dtype)
        ptr = _dp(out) if dtype == np.float64 else out.ctypes.data_as(C.POINTER(C.c_uint8))
        self._check(fn(self._ctx, ptr))
        return out

    def xB(self):
        return self._vec(self._L.sb200_get_xB, self.m)

    def y(self):
        return self._vec(self._L.sb200_get_y, self.m)

    def d(self):
        return self._vec(self._L.sb200_get_d, self.m)

    def b(self):
        return self._vec(self._L.sb200_get_b, self.m)

    def flip_flags(self):
        return self._vec(self._L.sb200_get_flip_flags, self.N, dtype=np.uint8)

    def Binv(self):
        out = np.zeros((self.m, self.m))
        self._check(self._L.sb200_get_Binv(self._ctx, _dp(out)))
        return out

    def trace(self):
        n = C.c_int64(0)
        self._check(self._L.sb200_get_trace(self._ctx, None, 0, C.byref(n)))
        cap = max(int(n.value), 1)
        buf = (_abi.Pivot * cap)()
        self._check(self._L.sb200_get_trace(self._ctx, buf, cap, C.byref(n)))
        arr = np.frombuffer(buf, dtype=np.int32).reshape(cap, 4)
        return arr[:min(int(n.value), cap)].copy(), int(n.value)

    def primal_residual(self):
        """max_i |b_i - (A_B x_B)_i| of the resident basis (numerical hygiene, off the hot path)."""
        v = C.c_double(0.0)
        self._check(self._L.sb200_primal_residual(self._ctx, C.byref(v)))
        return float(v.value)

    def inverse_residual(self, rows=8, first=0):
        """max over `rows` evenly spread rows i of |e_i^T B^-1 A_B - e_i^T|_inf (numerical hygiene)."""
        v = C.c_double(0.0)
        self._check(self._L.sb200_inverse_residual(self._ctx, rows, first, C.byref(v)))
        return float(v.value)

    def hygiene_stats(self):
        """(checks, re-inversions they triggered, worst residual seen) of the last simplex() call."""
        a, b, w = C.c_int64(0), C.c_int64(0), C.c_double(0.0)
        self._check(self._L.sb200_hygiene_stats(self._ctx, C.byref(a), C.byref(b), C.byref(w)))
        return {"checks": int(a.value), "refreshes": int(b.value), "worst": float(w.value)}

    def reinversion_count(self):
        return int(self._L.sb200_reinversion_count(self._ctx))

    def inverse_carried(self):
        """True if the last simplex() / prepare() carried B^-1 and x_B over from the run that ended on its start basis."""
        return bool(self._L.sb200_inverse_carried(self._ctx))

    def launch_count(self):
        return int(self._L.sb200_launch_count(self._ctx))

    def engine_used(self):
        """Loop kernel of the last simplex() call: staged / persistent / fused / resident."""
        return _abi.ENGINE_USED_NAMES[int(self._L.sb200_engine_used(self._ctx))]

    def phase_cycles(self):
        """Persistent engine: cycles CTA 0 spent in (price, barrier, ftran, barrier, update, barrier,
        dual) summed over the iterations of the last run, and the iteration count."""
        out = np.zeros(8, dtype=np.int64)
        self._check(self._L.sb200_get_phase_cycles(self._ctx, _lp(out)))
        return out

    # -- single-phase entry points (one kernel each)
    def prepare(self, basis):
        self._check(self._L.sb200_prepare(self._ctx, _lp(basis.basic), basis.basic.size, _lp(basis.non_basic),
                                          basis.non_basic.size))

    def step_dual(self):
        self._check(self._L.sb200_step_dual(self._ctx))

    def step_price(self):
        pos, val = C.c_int64(-1), C.c_double(0.0)
        self._check(self._L.sb200_step_price(self._ctx, C.byref(pos), C.byref(val)))
        return int(pos.value), float(val.value)

    def step_ftran(self, pos):
        self._check(self._L.sb200_step_ftran(self._ctx, pos))

    def step_ratio(self, pos):
        row, theta = C.c_int64(-1), C.c_double(0.0)
        self._check(self._L.sb200_step_ratio(self._ctx, pos, C.byref(row), C.byref(theta)))
        return int(row.value), float(theta.value)

    def step_update(self):
        self._check(self._L.sb200_step_update(self._ctx))

    # -- batched mode
    def run_batched(self, A, b, c, basic, nonbasic, ub=None):
        """A: (batch, N, m) — each LP column-major; b: (batch, m); c: (batch, N);
        basic: (batch, m) int32; nonbasic: (batch, N-m) int32 (both updated in place).
        ub: None or (batch, N) finite upper bounds (DBL_MAX = none): bounded ratio test and bound flips; the result
        then also has "flip" (batch, N) and "flips" (batch,)."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        batch, N, m = A.shape
        assert basic.dtype == np.int32 and nonbasic.dtype == np.int32
        assert basic.shape == (batch, m) and nonbasic.shape == (batch, N - m)
        xB = np.zeros((batch, m))
        obj = np.zeros(batch)
        term = np.zeros(batch, dtype=np.int32)
        its = np.zeros(batch, dtype=np.int32)
        sec = C.c_double(0.0)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        if ub is not None:
            ub = np.ascontiguousarray(ub, dtype=np.float64)
            assert ub.shape == (batch, N)
            flip = np.zeros((batch, N), dtype=np.uint8)
            flips = np.zeros(batch, dtype=np.int32)
            self._check(self._L.sb200_run_batched_bounded(self._ctx, batch, m, N, vp(A), vp(b), vp(c), vp(ub), vp(basic),
                                                          vp(nonbasic), vp(xB), vp(obj), vp(term), vp(its), vp(flip), vp(flips),
                                                          0, C.byref(sec)))
            return {"xB": xB, "objective": obj, "term": term, "iterations": its, "seconds": sec.value, "flip": flip,
                    "flips": flips}
        self._check(self._L.sb200_run_batched(self._ctx, batch, m, N, vp(A), vp(b), vp(c), vp(basic), vp(nonbasic),
                                              vp(xB), vp(obj), vp(term), vp(its), 0, C.byref(sec)))
        return {"xB": xB, "objective": obj, "term": term, "iterations": its, "seconds": sec.value}


    def run_batched_device(self, batch, m, N, dA, db, dc, dbasic, dnonbasic):
        """Same as run_batched with everything already in HBM: arguments are torch CUDA tensors
        (A [batch,N,m] f64, b, c f64, lists int32, updated in place); results stay on the device."""
        import torch
        dev = dA.device
        xB = torch.empty((batch, m), dtype=torch.float64, device=dev)
        obj = torch.empty(batch, dtype=torch.float64, device=dev)
        term = torch.empty(batch, dtype=torch.int32, device=dev)
        its = torch.empty(batch, dtype=torch.int32, device=dev)
        sec = C.c_double(0.0)
        for t, dt in ((dA, torch.float64), (db, torch.float64), (dc, torch.float64), (dbasic, torch.int32),
                      (dnonbasic, torch.int32)):
            assert t.is_cuda and t.dtype == dt and t.is_contiguous()
        vp = lambda t: C.c_void_p(t.data_ptr())
        self._check(self._L.sb200_run_batched(self._ctx, batch, m, N, vp(dA), vp(db), vp(dc), vp(dbasic), vp(dnonbasic),
                                              vp(xB), vp(obj), vp(term), vp(its), 1, C.byref(sec)))
        return {"xB": xB, "objective": obj, "term": term, "iterations": its, "seconds": sec.value}


def algorithmic_bytes_per_pivot(m, n_nonbasic):
    """8*(m*N_nb + 4*m^2)."""
    return float(_abi.lib().sb200_algorithmic_bytes_per_pivot(m, n_nonbasic))
