"""TEST INFRASTRUCTURE — ctypes view of oracle/liboracle.so (the C restatement of the reference's
hot path, oracle/simplex_oracle.c). Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product (simplex_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TERM = {0: "running", 1: "optimal", 2: "unbounded", 3: "iter_limit"}
RULE_FIRST, RULE_MOST_NEG = 0, 1
RULES = {"stock": RULE_FIRST, "first": RULE_FIRST, "dantzig": RULE_MOST_NEG, "most_neg": RULE_MOST_NEG}


class Result(C.Structure):
    _fields_ = [("term", C.c_int32), ("pad", C.c_int32), ("iterations", C.c_int64), ("pivots", C.c_int64),
                ("flips", C.c_int64), ("objective", C.c_double)]


def build():
    """Compile the restatement (and the reference binaries when /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp, lp, u8p, i32p = (C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.POINTER(C.c_int32))
        sig = [C.c_int64, C.c_int64, dp, dp, dp, dp, lp, lp, C.c_int, C.c_double, C.c_int64, dp, dp, u8p, i32p,
               C.c_int64, C.POINTER(Result)]
        for f in (L.oracle_simplex_lu, L.oracle_simplex_pfi):
            f.argtypes = sig
            f.restype = C.c_int
        L.oracle_rescale.argtypes = [C.c_int64, C.c_int64, dp, dp, dp]
        L.oracle_rescale.restype = None
        L.oracle_synth.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_uint64, dp, dp, dp]
        L.oracle_synth.restype = None
        L.oracle_synth_tableau.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int, dp, dp, dp, lp, lp]
        L.oracle_synth_tableau.restype = C.c_int64
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _lp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def simplex(A, b, c, basic, nonbasic, ub=None, rule=RULE_FIRST, eps=1e-7, iter_limit=10**9, variant="lu",
            trace_cap=1 << 20):
    """Run the restated Simplex::simplex(Basis&). A is (m, N), any layout; copies are made, the
    mutated copies (bound flips) are returned. Returns a dict."""
    m, N = A.shape
    Af = np.array(A, dtype=np.float64, order="F", copy=True)
    bb = np.array(b, dtype=np.float64, copy=True)
    cc = np.array(c, dtype=np.float64, copy=True)
    bas = np.array(basic, dtype=np.int64, copy=True)
    non = np.array(nonbasic, dtype=np.int64, copy=True)
    assert bas.size == m and non.size == N - m
    ubp = None
    if ub is not None:
        ubb = np.array(ub, dtype=np.float64, copy=True)
        ubp = _dp(ubb)
    x = np.zeros(m)
    y = np.zeros(m)
    flip = np.zeros(N, dtype=np.uint8)
    trace = np.zeros((trace_cap, 4), dtype=np.int32)
    res = Result()
    fn = lib().oracle_simplex_lu if variant == "lu" else lib().oracle_simplex_pfi
    rc = fn(m, N, _dp(Af), _dp(bb), _dp(cc), ubp, _lp(bas), _lp(non), int(rule), float(eps), int(iter_limit), _dp(x),
            _dp(y), flip.ctypes.data_as(C.POINTER(C.c_uint8)), trace.ctypes.data_as(C.POINTER(C.c_int32)), trace_cap,
            C.byref(res))
    assert rc == 0
    return {"term": TERM[res.term], "iterations": res.iterations, "pivots": res.pivots, "flips": res.flips,
            "objective": res.objective, "x": x, "y": y, "flip": flip, "basic": bas, "nonbasic": non,
            "trace": trace[:min(res.pivots, trace_cap)].copy(), "A": Af, "b": bb, "c": cc}


def synth(m, n, n_eq, seed):
    """G(m,n,n_eq,seed): returns (c[n], a[m,n] row-major, rhs[m])."""
    c = np.zeros(n)
    a = np.zeros((m, n))
    rhs = np.zeros(m)
    lib().oracle_synth(m, n, n_eq, seed, _dp(c), _dp(a), _dp(rhs))
    return c, a, rhs


def synth_tableau(m, n, n_eq, seed, phase1):
    """Dense standard-form tableau of G as the reference flow builds it (scaled). Returns
    (A[m,N] Fortran order, b, c, basic, nonbasic); the lists are the crash basis when it is
    defined by this phase alone (phase1 or n_eq == 0), else None."""
    N = n + (m - n_eq) + (n_eq if phase1 else 0)
    A = np.zeros((m, N), order="F")
    b = np.zeros(m)
    c = np.zeros(N)
    basic = np.zeros(m, dtype=np.int64)
    nonbasic = np.zeros(N - m, dtype=np.int64)
    got = lib().oracle_synth_tableau(m, n, n_eq, seed, 1 if phase1 else 0, _dp(A), _dp(b), _dp(c), _lp(basic),
                                     _lp(nonbasic))
    assert got == N
    if phase1 or n_eq == 0:
        return A, b, c, basic, nonbasic
    return A, b, c, None, None


def phase2_lists_from_phase1(basic1, nonbasic1, n_struct_slack):
    """InitialBasis.cpp:85-121: artificial ids (>= n_struct_slack) are deleted in place from the
    pivot-permuted Phase-I lists (Basis.cpp:25-35). Raises if one is still basic (:97-114)."""
    if any(j >= n_struct_slack for j in basic1):
        raise RuntimeError("artificial variable in the basis")
    return np.array(basic1, dtype=np.int64), np.array([j for j in nonbasic1 if j < n_struct_slack], dtype=np.int64)
