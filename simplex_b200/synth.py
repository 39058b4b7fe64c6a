"""Synthetic dense feasible LP G(m, n, n_eq, seed) of BASELINE.json's configs,
built by the host C++ library (simplex_b200/host/dense_tableau.cpp) straight into the dense
standard-form tableau the hot loop consumes, i.e. what the reference's
LinearProgram::initialize_tableau + rescale_matrix would hand to Simplex::simplex()
(reference src/LinearProgram.cpp:64-97, 171-195)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libsimplex_b200_host.so")
        if not os.path.exists(path):
            raise ImportError("%s is missing: run simplex_b200/host/build.sh" % path)
        L = C.CDLL(path)
        dp, lp = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.sbh_synth_num_cols.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int]
        L.sbh_synth_num_cols.restype = C.c_int64
        L.sbh_synth_tableau.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int, dp, dp, dp, lp, lp,
                                        C.c_int64, C.c_int64, C.c_int64]
        L.sbh_synth_tableau.restype = C.c_int64
        L.sbh_synth_raw.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_uint64, dp, dp, dp]
        L.sbh_synth_raw.restype = None
        L.sbh_synth_batch.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_uint64, dp, dp, dp, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.c_int]
        L.sbh_synth_batch.restype = None
        _LIB = L
    return _LIB


def num_cols(m, n, n_eq, phase1):
    return int(_lib().sbh_synth_num_cols(m, n, n_eq, 1 if phase1 else 0))


def tableau(m, n, n_eq=0, seed=1234, phase1=False, col0=0, ncols=None, out=None, col_stride=1):
    """Returns (A, b, c, basic, nonbasic). A is (m, ncols) in Fortran order (column-major, the
    layout of the reference's matrix_A); with col0/ncols/col_stride only the columns col0,
    col0+col_stride, ... are built (column shards of the multi-GPU mode: cyclic, stride = world). basic/nonbasic are the crash basis when this phase
    defines it (phase1 or n_eq == 0), else None. `out` may be a preallocated (pinned) buffer."""
    N = num_cols(m, n, n_eq, phase1)
    if ncols is None:
        ncols = -(-(N - col0) // col_stride)
    A = out if out is not None else np.zeros((m, ncols), order="F")
    assert A.shape == (m, ncols) and A.flags.f_contiguous
    b = np.zeros(m)
    c = np.zeros(N)
    basic = np.zeros(m, dtype=np.int64)
    nonbasic = np.zeros(N - m, dtype=np.int64)
    dp, lp = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    got = _lib().sbh_synth_tableau(m, n, n_eq, seed, 1 if phase1 else 0, A.ctypes.data_as(dp), b.ctypes.data_as(dp),
                                   c.ctypes.data_as(dp), basic.ctypes.data_as(lp), nonbasic.ctypes.data_as(lp), col0,
                                   ncols, col_stride)
    assert got == N
    if phase1 or n_eq == 0:
        return A, b, c, basic, nonbasic
    return A, b, c, None, None


def phase2_lists(basic1, nonbasic1, n_struct_slack):
    """Phase hand-off of the reference (src/InitialBasis.cpp:85-121): the artificial columns
    (ids >= n_struct_slack) are deleted in place from the pivot-permuted Phase-I lists. Raises
    like the reference (:97-114) if one of them is still basic."""
    if any(int(j) >= n_struct_slack for j in basic1):
        raise RuntimeError("artificial variable in the basis.")
    return (np.array(basic1, dtype=np.int64),
            np.array([j for j in nonbasic1 if j < n_struct_slack], dtype=np.int64))


def batch(count, m, n, seed0=1234, threads=0, out=None):
    """`count` independent LPs G(m, n, 0, seed0 + k) in the layout of sb200_run_batched:
    (A[count, N, m], b[count, m], c[count, N], basic[count, m] int32, nonbasic[count, n] int32)."""
    N = n + m
    A = out if out is not None else np.zeros((count, N, m))
    assert A.shape == (count, N, m) and A.flags.c_contiguous
    b = np.zeros((count, m))
    c = np.zeros((count, N))
    basic = np.zeros((count, m), dtype=np.int32)
    nonbasic = np.zeros((count, n), dtype=np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    _lib().sbh_synth_batch(count, m, n, seed0, A.ctypes.data_as(dp), b.ctypes.data_as(dp), c.ctypes.data_as(dp),
                           basic.ctypes.data_as(ip), nonbasic.ctypes.data_as(ip), threads)
    return A, b, c, basic, nonbasic


def raw(m, n, n_eq=0, seed=1234):
    """The LP G(m, n, n_eq, seed) itself: (c[n], a[m, n] row-major, rhs[m], types) with rows
    0..n_eq-1 equalities and the others '<='; minimise c.x, x >= 0."""
    c = np.zeros(n)
    a = np.zeros((m, n))
    rhs = np.zeros(m)
    dp = C.POINTER(C.c_double)
    _lib().sbh_synth_raw(m, n, n_eq, seed, c.ctypes.data_as(dp), a.ctypes.data_as(dp), rhs.ctypes.data_as(dp))
    return c, a, rhs, ["="] * n_eq + ["<"] * (m - n_eq)
