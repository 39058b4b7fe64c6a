"""ctypes view of the host library libsimplex_b200_host.so (simplex_b200/host/*.cpp): the text
model, the .lp reader, presolve and the two-phase control of the reference (reference
src/Parser.cpp, src/Presolve.cpp, src/InitialBasis.cpp, src/Simplex.cpp:22-45,181-245,356-373),
with the per-iteration loop on the GPU.

    with LpSession() as s:
        s.read(open("model.lp").read())      # parse + presolve
        s.solve()                            # Phase I / Phase II on the device
        s.objective(), s.solution(), s.write("sol_test.xml")

The step methods (open_phase_one / close_phase_one / initialize_tableau / absorb_optimum) are
the host halves of Simplex::solve() around the device loop; the parity tests drive them one at
a time. Nothing here computes a pivot on the CPU.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libsimplex_b200_host.so")
CLI_PATH = os.path.join(_HERE, "simplex")
_LIB = None

STATUS = {0: "not_solved", 1: "optimal", 2: "unbounded", 3: "infeasible"}
READ_OK, READ_EXCEPTION, READ_FATAL, READ_PRESOLVE_INFEASIBLE = 0, 1, 2, 3


class LpError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(msg)
        self.rc = rc


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise ImportError("%s is missing: run simplex_b200/host/build.sh" % HOST_LIB_PATH)
        L = C.CDLL(HOST_LIB_PATH)
        vp, i64p, dp = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double)
        sigs = {
            "sbh_open": (vp, []), "sbh_close": (None, [vp]), "sbh_error": (C.c_char_p, [vp]),
            "sbh_read": (C.c_int, [vp, C.c_char_p, C.c_int]),
            "sbh_build_synth": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_int64, C.c_uint64]), "sbh_saw_end": (C.c_int, [vp]),
            "sbh_set_options": (None, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]),
            "sbh_open_phase_one": (C.c_int, [vp, i64p]), "sbh_close_phase_one": (C.c_int, [vp]),
            "sbh_initialize_tableau": (C.c_int, [vp]),
            "sbh_dims": (None, [vp, i64p, i64p, i64p, i64p, C.POINTER(C.c_int)]),
            "sbh_copy_tableau": (None, [vp, dp, dp, dp, dp]), "sbh_copy_basis": (None, [vp, i64p, i64p]),
            "sbh_set_basis": (None, [vp, i64p, C.c_int64, i64p, C.c_int64]),
            "sbh_absorb_optimum": (C.c_int, [vp, dp, C.POINTER(C.c_uint8)]),
            "sbh_solve": (C.c_int, [vp]), "sbh_status": (C.c_int, [vp]), "sbh_objective": (C.c_double, [vp]),
            "sbh_sense": (C.c_char, [vp]), "sbh_write": (C.c_int, [vp, C.c_char_p]),
            "sbh_solution_table": (C.c_int64, [vp, C.c_char_p, C.c_int64]),
            "sbh_names_table": (C.c_int64, [vp, C.c_char_p, C.c_int64]),
            "sbh_shifts_table": (C.c_int64, [vp, C.c_char_p, C.c_int64]),
            "sbh_calls": (C.c_int64, [vp, i64p, C.c_int64]),
            "sbh_dense_solve": (C.c_int, [C.c_int64, C.c_int64, dp, C.c_char_p, dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int64, C.POINTER(C.c_int), dp, dp, dp, i64p, i64p, dp, C.c_char_p, C.c_int64]),
        }
        for name, (res, args) in sigs.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        L._sbh_symbols = sorted(sigs)
        _LIB = L
    return _LIB


def _i64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _f64(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class LpSession:
    """One LinearProgram + Simplex + InitialBasis of the host library."""

    def __init__(self, pricing="first", device=0, engine="persistent", dual="update", chunk_iters=256):
        self._L = lib()
        self._h = C.c_void_p(self._L.sbh_open())
        self.set_options(pricing=pricing, device=device, engine=engine, dual=dual, chunk_iters=chunk_iters)

    def set_options(self, pricing="first", device=0, engine="persistent", dual="update", chunk_iters=256):
        p = {"first": 0, "stock": 0, "most_neg": 1, "dantzig": 1}[pricing]
        e = {"persistent": 0, "staged": 1, "streaming": 2}[engine]
        d = {"recompute": 0, "update": 1}[dual]
        self._L.sbh_set_options(self._h, device, p, e, d, chunk_iters)

    def close(self):
        if self._h is not None and self._h.value:
            self._L.sbh_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise LpError(rc, self._L.sbh_error(self._h).decode())

    # -- model in
    def read(self, text, presolve=True, check=True):
        rc = self._L.sbh_read(self._h, text.encode(), 1 if presolve else 0)
        if check:
            self._check(rc)
        return rc

    def build_synth(self, m, n, n_eq, seed):
        """G(m,n,n_eq,seed) through add_variable/add_constraint + presolve."""
        self._check(self._L.sbh_build_synth(self._h, m, n, n_eq, seed))

    def error(self):
        return self._L.sbh_error(self._h).decode()

    def saw_end(self):
        return bool(self._L.sbh_saw_end(self._h))

    # -- host halves of Simplex::solve()
    def open_phase_one(self):
        n = C.c_int64(0)
        self._check(self._L.sbh_open_phase_one(self._h, C.byref(n)))
        return int(n.value)

    def close_phase_one(self):
        self._check(self._L.sbh_close_phase_one(self._h))

    def initialize_tableau(self):
        self._check(self._L.sbh_initialize_tableau(self._h))

    def dims(self):
        r, c, nb, nn = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        bounded = C.c_int()
        self._L.sbh_dims(self._h, C.byref(r), C.byref(c), C.byref(nb), C.byref(nn), C.byref(bounded))
        return {"rows": r.value, "cols": c.value, "n_basic": nb.value, "n_non_basic": nn.value, "bounded": bool(bounded.value)}

    def tableau(self):
        """(A[m,N] Fortran order, b, c, ub) exactly as the next call of the seam would upload them."""
        d = self.dims()
        m, N = d["rows"], d["cols"]
        A = np.zeros((m, N), order="F")
        b, c, ub = np.zeros(m), np.zeros(N), np.zeros(N)
        self._L.sbh_copy_tableau(self._h, _f64(A), _f64(b), _f64(c), _f64(ub))
        return A, b, c, ub

    def basis(self):
        d = self.dims()
        basic = np.zeros(d["n_basic"], dtype=np.int64)
        non_basic = np.zeros(d["n_non_basic"], dtype=np.int64)
        self._L.sbh_copy_basis(self._h, _i64(basic), _i64(non_basic))
        return basic, non_basic

    def set_basis(self, basic, non_basic):
        basic = np.ascontiguousarray(basic, dtype=np.int64)
        non_basic = np.ascontiguousarray(non_basic, dtype=np.int64)
        self._L.sbh_set_basis(self._h, _i64(basic), basic.size, _i64(non_basic), non_basic.size)

    def absorb_optimum(self, x_B, flags=None):
        x_B = np.ascontiguousarray(x_B, dtype=np.float64)
        fp = None
        if flags is not None:
            flags = np.ascontiguousarray(flags, dtype=np.uint8)
            fp = flags.ctypes.data_as(C.POINTER(C.c_uint8))
        self._check(self._L.sbh_absorb_optimum(self._h, _f64(x_B), fp))

    # -- the whole flow (GPU)
    def solve(self):
        self._check(self._L.sbh_solve(self._h))

    def status(self):
        return STATUS[self._L.sbh_status(self._h)]

    def objective(self):
        return float(self._L.sbh_objective(self._h))

    def sense(self):
        return self._L.sbh_sense(self._h).decode()

    def write(self, path):
        self._check(self._L.sbh_write(self._h, os.fsencode(path)))

    def _table(self, fn):
        need = fn(self._h, None, 0)
        buf = C.create_string_buffer(int(need))
        fn(self._h, buf, need)
        return buf.value.decode()

    def solution(self):
        out = {}
        for line in self._table(self._L.sbh_solution_table).splitlines():
            name, value = line.split("\t")
            out[name] = float(value)
        return out

    def names(self):
        """(variable names by ascending id, row labels in tableau order)"""
        text = self._table(self._L.sbh_names_table)
        head, _, tail = text.partition("#rows\n")
        ids = [ln.split("\t") for ln in head.splitlines()]
        rows = tail.split("\n")[:-1] if tail else []
        return {int(i): n for i, n in ids}, rows

    def shifts(self):
        return {int(ln.split("\t")[0]): float(ln.split("\t")[1]) for ln in self._table(self._L.sbh_shifts_table).splitlines()}

    def calls(self):
        buf = np.zeros(8 * 16, dtype=np.int64)
        n = int(self._L.sbh_calls(self._h, _i64(buf), 16))
        keys = ("M", "N", "term", "iterations", "pivots", "bound_flips", "loop_us", "resident")
        return [dict(zip(keys, buf[8 * k:8 * k + 8].tolist())) for k in range(min(n, 16))]


def solve_lp_text(text, pricing="first", device=0):
    """Parse, presolve and solve on the GPU; returns (status, objective, solution dict)."""
    with LpSession(pricing=pricing, device=device) as s:
        s.read(text)
        s.solve()
        return s.status(), s.objective(), s.solution()


def solve_dense(a, types, rhs, c, ub=None, pricing="first", device=0, engine="persistent", dual="update", chunk_iters=256):
    """Dense fast path of the host library (simplex_b200/host/dense_solver.h): min c.x subject to
    a[i].x (types[i] in '<', '>', '=') rhs[i], 0 <= x <= ub, through the reference's two-phase flow
    with both loops on the GPU, without the string-keyed model. a: (m, n). Returns a dict."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    m, n = a.shape
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    tb = "".join(types).encode() if not isinstance(types, (bytes, bytearray)) else bytes(types)
    assert len(tb) == m and rhs.size == m and c.size == n
    ubp = None
    if ub is not None:
        ub = np.ascontiguousarray(ub, dtype=np.float64)
        assert ub.size == n
        ubp = _f64(ub)
    status, n_calls = C.c_int(0), C.c_int64(0)
    objective, seconds = C.c_double(0.0), C.c_double(0.0)
    x_scaled, x = np.zeros(n), np.zeros(n)
    calls = np.zeros(16, dtype=np.int64)
    err = C.create_string_buffer(512)
    rc = lib().sbh_dense_solve(m, n, _f64(a), tb, _f64(rhs), _f64(c), ubp, device,
                               {"first": 0, "stock": 0, "most_neg": 1, "dantzig": 1}[pricing],
                               {"persistent": 0, "staged": 1, "streaming": 2}[engine], {"recompute": 0, "update": 1}[dual], chunk_iters,
                               C.byref(status), C.byref(objective), _f64(x_scaled), _f64(x), _i64(calls), C.byref(n_calls),
                               C.byref(seconds), err, 512)
    if rc != 0:
        raise LpError(rc, err.value.decode())
    keys = ("M", "N", "term", "iterations", "pivots", "bound_flips", "loop_us", "resident")
    return {"status": STATUS[status.value], "objective": objective.value, "x_scaled": x_scaled, "x": x,
            "calls": [dict(zip(keys, calls[8 * k:8 * k + 8].tolist())) for k in range(min(int(n_calls.value), 2))],
            "seconds_host": seconds.value}
