"""Shared helpers of the parity tests (test infrastructure)."""
import numpy as np

DBL_MAX = np.finfo(np.float64).max

# the two Phase-I calls whose vertices are degenerate (b contains 0): ties are
# decided by rounding noise, so only the objective is pinned for them (north_star: "same pivot
# count and final basis on non-degenerate instances").
DEGENERATE_CALLS = {("test", 0), ("test_presolve_2", 0)}


def call_inputs(call):
    """(A[m,N] Fortran, b, c, ub-or-None, basic, nonbasic) of one recorded Simplex::simplex() call."""
    m, N = call["M"], call["N"]
    A = np.array(call["A"], dtype=np.float64).reshape((N, m)).T.copy(order="F")
    b = np.array(call["b"], dtype=np.float64)
    c = np.array(call["c"], dtype=np.float64)
    ub = np.array(call["ub"], dtype=np.float64)
    return A, b, c, ub, np.array(call["basic_in"], dtype=np.int64), np.array(call["nonbasic_in"], dtype=np.int64)


def has_bounds(rec):
    """m_has_UBS (LinearProgram.cpp:151): true iff the file has a finite upper bound."""
    txt = rec["lp_text"].lower()
    return "bounds" in txt


def rel_close(a, b, tol=1e-9):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))))
