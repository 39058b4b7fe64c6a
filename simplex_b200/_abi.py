"""ctypes binding of the C ABI in include/simplex_b200.h (libsimplex_b200.so, built in-tree by
simplex_b200/csrc/build.sh). Product code: it never imports anything from oracle/ and it fails
loudly when the CUDA library is missing or no GPU is present — there is no CPU fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB200_LIB_PATH") or os.path.join(_HERE, "libsimplex_b200.so")   # (override: kernel-variant probes)

ABI_VERSION = 2
OK = 0
RC_NAMES = {0: "OK", 1: "ERR_INVALID", 2: "ERR_CUDA", 3: "ERR_NOMEM", 4: "ERR_STATE", 5: "ERR_BASIS_SIZE",
            6: "ERR_SINGULAR", 7: "ERR_UNSUPPORTED"}
PRICE_FIRST_ELIGIBLE, PRICE_MOST_NEGATIVE = 0, 1
ENGINE_PERSISTENT, ENGINE_STAGED, ENGINE_STREAMING = 0, 1, 2
DUAL_RECOMPUTE, DUAL_UPDATE = 0, 1
ENGINE_USED_NAMES = {0: "none", 1: "staged", 2: "persistent", 3: "fused", 4: "resident"}
TERM_NAMES = {0: "running", 1: "optimal", 2: "unbounded", 3: "iter_limit"}


class Opts(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("pricing", C.c_int32), ("engine", C.c_int32),
                ("dual_mode", C.c_int32), ("dual_refresh", C.c_int32), ("reinvert_period", C.c_int32),
                ("trace_capacity", C.c_int32), ("eps", C.c_double), ("iter_limit", C.c_int64),
                ("chunk_iters", C.c_int64), ("stream", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32),
                ("hygiene_tol", C.c_double), ("hygiene_period", C.c_int32), ("hygiene_rows", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("term", C.c_int32), ("reserved", C.c_int32), ("iterations", C.c_int64), ("pivots", C.c_int64),
                ("bound_flips", C.c_int64), ("objective", C.c_double), ("loop_seconds", C.c_double)]


class Pivot(C.Structure):
    _fields_ = [("leaving_col", C.c_int32), ("row", C.c_int32), ("entering_col", C.c_int32), ("pos", C.c_int32)]


# every symbol include/simplex_b200.h declares: (name, restype, argtypes)
_dp, _lp, _vp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_void_p
_i32p, _u8p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
SYMBOLS = [
    ("sb200_default_opts", None, [C.POINTER(Opts)]),
    ("sb200_create", C.c_int, [C.POINTER(_vp), C.POINTER(Opts)]),
    ("sb200_destroy", None, [_vp]),
    ("sb200_last_error", C.c_char_p, [_vp]),
    ("sb200_abi_version", C.c_int, []),
    ("sb200_load", C.c_int, [_vp, C.c_int64, C.c_int64, _dp, C.c_int64, _dp, _dp, _dp]),
    ("sb200_load_device", C.c_int, [_vp, C.c_int64, C.c_int64, _vp, C.c_int64, _vp, _vp, _vp]),
    ("sb200_set_costs", C.c_int, [_vp, C.c_int64, _dp]),
    ("sb200_run", C.c_int, [_vp, _lp, C.c_int64, _lp, C.c_int64, C.POINTER(Result)]),
    ("sb200_get_xB", C.c_int, [_vp, _dp]),
    ("sb200_get_y", C.c_int, [_vp, _dp]),
    ("sb200_get_d", C.c_int, [_vp, _dp]),
    ("sb200_get_flip_flags", C.c_int, [_vp, _u8p]),
    ("sb200_get_Binv", C.c_int, [_vp, _dp]),
    ("sb200_get_b", C.c_int, [_vp, _dp]),
    ("sb200_get_trace", C.c_int, [_vp, C.POINTER(Pivot), C.c_int64, _lp]),
    ("sb200_prepare", C.c_int, [_vp, _lp, C.c_int64, _lp, C.c_int64]),
    ("sb200_step_dual", C.c_int, [_vp]),
    ("sb200_step_price", C.c_int, [_vp, _lp, _dp]),
    ("sb200_step_ftran", C.c_int, [_vp, C.c_int64]),
    ("sb200_step_ratio", C.c_int, [_vp, C.c_int64, _lp, _dp]),
    ("sb200_step_update", C.c_int, [_vp]),
    ("sb200_run_batched", C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                    C.c_int, _dp]),
    ("sb200_run_batched_bounded", C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                            _vp, _vp, _vp, C.c_int, _dp]),
    ("sb200_shard_window_handle", C.c_int, [_vp, _vp]),
    ("sb200_shard_connect", C.c_int, [_vp, _vp]),
    ("sb200_shard_connect_local", C.c_int, [_vp, C.POINTER(_vp), C.c_int64]),
    ("sb200_load_shard", C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, C.c_int64, _vp, _vp,
                                   C.c_int]),
    ("sb200_shard_prepare", C.c_int, [_vp, _lp, C.c_int64, _lp, C.c_int64, _dp]),
    ("sb200_shard_finish_prepare", C.c_int, [_vp, _dp]),
    ("sb200_shard_run", C.c_int, [_vp, _lp, _lp, C.POINTER(Result)]),
    ("sb200_shard_get_xB_local", C.c_int, [_vp, _dp, _lp, _lp]),
    ("sb200_shard_set_bounds", C.c_int, [_vp, _dp]),
    ("sb200_shard_set_costs", C.c_int, [_vp, C.c_int64, _dp]),
    ("sb200_shard_continue", C.c_int, [_vp, _lp, C.c_int64, _lp, C.c_int64]),
    ("sb200_shard_residual_partial", C.c_int, [_vp, _dp, _dp]),
    ("sb200_shard_prepare_general", C.c_int, [_vp, _lp, C.c_int64, _lp, C.c_int64]),
    ("sb200_shard_get_basic_columns", C.c_int, [_vp, _dp, _lp, C.c_int64, _lp]),
    ("sb200_shard_invert", C.c_int, [_vp, _dp]),
    ("sb200_shard_resume", C.c_int, [_vp, C.c_int64]),
    ("sb200_primal_residual", C.c_int, [_vp, _dp]),
    ("sb200_inverse_residual", C.c_int, [_vp, C.c_int64, C.c_int64, _dp]),
    ("sb200_hygiene_stats", C.c_int, [_vp, _lp, _lp, _dp]),
    ("sb200_reinversion_count", C.c_int64, [_vp]),
    ("sb200_debug_check_guards", C.c_int, [_vp, _lp, _lp]),
    ("sb200_algorithmic_bytes_per_pivot", C.c_double, [C.c_int64, C.c_int64]),
    ("sb200_launch_count", C.c_int64, [_vp]),
    ("sb200_inverse_carried", C.c_int, [_vp]),
    ("sb200_get_phase_cycles", C.c_int, [_vp, _lp]),
    ("sb200_engine_used", C.c_int, [_vp]),
]

_LIB = None
# SB200_GUARD=1: every DeviceSimplex / ShardedSimplex checks the guard zones of its device buffers when it is
# closed and raises if a kernel wrote outside a buffer (tests/conftest.py switches it on for the GPU suite)
GUARD_CHECK_AT_CLOSE = os.environ.get("SB200_GUARD", "0") not in ("", "0")
GUARD_STATS = {"contexts": 0, "buffers": 0, "damaged": 0}


class Sb200Error(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__("sb200 %s: %s" % (RC_NAMES.get(rc, rc), msg))
        self.rc = rc


def lib():
    """Load the CUDA library. Raises if it has not been built: the product has no other path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run simplex_b200/csrc/build.sh (or __graft_entry__.build()); "
                              "simplex_b200 has no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        if L.sb200_abi_version() != ABI_VERSION:
            raise ImportError("libsimplex_b200.so ABI %d != binding %d" % (L.sb200_abi_version(), ABI_VERSION))
        _LIB = L
    return _LIB


def default_opts():
    o = Opts()
    lib().sb200_default_opts(C.byref(o))
    return o
