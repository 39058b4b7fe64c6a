"""CPU-side checks: the C-ABI library loads and exports every symbol include/simplex_b200.h
declares, fails loudly without a GPU, and the product-side generator equals the oracle's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "simplex_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_A-Z]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from simplex_b200 import _abi
    lib = _abi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "libsimplex_b200.so does not export %s" % name
    bound = {n for n, _, _ in _abi.SYMBOLS}
    assert bound == set(declared), (bound ^ set(declared))
    assert lib.sb200_abi_version() == _abi.ABI_VERSION


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from simplex_b200 import DeviceSimplex, Sb200Error
    with pytest.raises(Sb200Error) as ei:
        DeviceSimplex()
    assert "no CUDA device" in str(ei.value)


def test_product_never_imports_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "simplex_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "simplex_oracle" not in txt, os.path.join(dirpath, f)
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), os.path.join(dirpath, f)


def test_algorithmic_bytes_figure():
    from simplex_b200 import algorithmic_bytes_per_pivot
    # config 3 moves exactly 1 GiB per pivot
    assert algorithmic_bytes_per_pivot(4096, 16384) == 1073741824.0
    assert algorithmic_bytes_per_pivot(1024, 1792) == 8.0 * (1024 * 1792 + 4 * 1024 ** 2)


@pytest.mark.parametrize("m,n,n_eq,phase1", [(8, 16, 0, False), (16, 32, 4, True), (16, 32, 4, False),
                                             (64, 128, 16, True), (200, 300, 50, False)])
def test_host_generator_equals_oracle_generator(m, n, n_eq, phase1):
    from oracle import oracle
    from simplex_b200 import synth
    A, b, c, bas, non = synth.tableau(m, n, n_eq, 1234, phase1)
    Ao, bo, co, baso, nono = oracle.synth_tableau(m, n, n_eq, 1234, phase1)
    assert np.array_equal(A, Ao) and np.array_equal(b, bo) and np.array_equal(c, co)
    if bas is not None:
        assert bas.tolist() == baso.tolist() and non.tolist() == nono.tolist()


def test_host_generator_column_shards():
    from simplex_b200 import synth
    A, b, c, _, _ = synth.tableau(48, 100, 0, 9)
    for col0, ncols in ((0, 37), (37, 63), (100, 48), (90, 30)):
        As, bs, cs, _, _ = synth.tableau(48, 100, 0, 9, col0=col0, ncols=ncols)
        assert np.array_equal(As, A[:, col0:col0 + ncols]) and np.array_equal(bs, b) and np.array_equal(cs, c)
    for world in (2, 3, 8):       # cyclic shards of the multi-GPU mode
        for r in range(world):
            As, bs, cs, _, _ = synth.tableau(48, 100, 0, 9, col0=r, col_stride=world)
            assert np.array_equal(As, A[:, r::world]) and np.array_equal(bs, b) and np.array_equal(cs, c)


def test_phase2_lists_drop_artificials_in_place():
    from simplex_b200 import synth
    bas, non = synth.phase2_lists([3, 1, 4], [7, 0, 6, 2, 5], 6)
    assert bas.tolist() == [3, 1, 4] and non.tolist() == [0, 2, 5]
    with pytest.raises(RuntimeError):
        synth.phase2_lists([3, 6, 4], [7, 0, 1, 2, 5], 6)


def test_binding_constants_match_the_header():
    """The integer values the ctypes binding uses for the enums of include/simplex_b200.h (return codes,
    pricing rules, engines, dual modes, terminations, the engine_used report) are the header's."""
    import re
    from simplex_b200 import _abi
    text = open(os.path.join(ROOT, "include", "simplex_b200.h")).read()
    enum = {name: int(val) for name, val in re.findall(r"\b(SB200_[A-Z0-9_]+)\s*=\s*(-?\d+)", text)}
    assert enum["SB200_OK"] == _abi.OK
    for val, name in _abi.RC_NAMES.items():
        assert enum["SB200_" + name] == val, name
    assert (enum["SB200_PRICE_FIRST_ELIGIBLE"], enum["SB200_PRICE_MOST_NEGATIVE"]) == \
        (_abi.PRICE_FIRST_ELIGIBLE, _abi.PRICE_MOST_NEGATIVE)
    assert (enum["SB200_ENGINE_PERSISTENT"], enum["SB200_ENGINE_STAGED"], enum["SB200_ENGINE_STREAMING"]) == \
        (_abi.ENGINE_PERSISTENT, _abi.ENGINE_STAGED, _abi.ENGINE_STREAMING)
    assert (enum["SB200_DUAL_RECOMPUTE"], enum["SB200_DUAL_UPDATE"]) == (_abi.DUAL_RECOMPUTE, _abi.DUAL_UPDATE)
    for val, name in _abi.TERM_NAMES.items():
        assert enum["SB200_TERM_" + name.upper()] == val, name
    for val, name in _abi.ENGINE_USED_NAMES.items():
        assert enum["SB200_ENGINE_USED_" + name.upper()] == val, name
