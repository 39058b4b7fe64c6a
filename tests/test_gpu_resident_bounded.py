"""Finite upper bounds on the shared-memory-resident engine (k_resident_bounded, sb200_resident.cuh): the bounded
ratio test, flips of the entering variable and substituted leaving variables of the reference
(src/Simplex.cpp:116-179, src/LinearProgram.cpp:154-169) with the tableau on chip. Every CTA keeps its own copy of
the sign flags; the columns are rewritten behind the kernel (k_apply_negations). Checked against the golden traces
of the reference's bounded fixtures, the LU oracle and, bit for bit, the bounded fused kernel."""
import numpy as np
import pytest

from helpers import DBL_MAX, DEGENERATE_CALLS, call_inputs, has_bounds, rel_close
from oracle import oracle
from test_gpu_parity import _bounded_synth, _run

pytestmark = pytest.mark.gpu


def test_bounded_resident_can_be_switched_off(monkeypatch):
    """SB200_RESIDENT_BOUNDED=0 keeps a bounded LP on the bounded fused kernel."""
    from simplex_b200 import Basis, DeviceSimplex
    monkeypatch.setenv("SB200_RESIDENT_BOUNDED", "0")
    A, b, c, ub, basic, nonbasic = _bounded_synth(64, 128, 3, 0.3)
    with DeviceSimplex(pricing="most_neg", dual="update") as s:
        s.load(A, b, c, ub)
        assert s.simplex(Basis(basic, nonbasic))["term"] == "optimal" and s.engine_used() == "fused"


def _same_as_oracle(got, want, tag):
    assert got["engine_used"] == "resident", tag
    assert (got["term"], got["iterations"], got["flips"]) == (want["term"], want["iterations"], want["flips"]), tag
    assert got["trace"].tolist() == want["trace"].tolist(), tag
    assert got["flip"].tolist() == want["flip"].tolist(), tag
    assert got["basic"].tolist() == want["basic"].tolist() and got["nonbasic"].tolist() == want["nonbasic"].tolist(), tag
    if want["term"] == "optimal":
        assert rel_close(got["x"], want["x"]) and rel_close(got["objective"], want["objective"]), tag


def test_bounded_fixtures_on_the_resident_engine(golden_fixtures):
    """The reference's bounded fixtures (8 of its 14 files): every recorded Simplex::simplex() call walks the
    reference's own pivot trace on k_resident_bounded, one launch and a launch every 3 iterations."""
    n = 0
    for name, rec in golden_fixtures.items():
        if not has_bounds(rec):
            continue
        for k, call in enumerate(rec["calls"]):
            A, b, c, ub, basic, nonbasic = call_inputs(call)
            for chunk in (64, 3):
                got = _run(A, b, c, basic, nonbasic, ub=ub, pricing="first", dual="update", chunk_iters=chunk)
                assert got["engine_used"] == "resident", (name, k)
                assert got["term"] == ("unbounded" if call["unbounded"] else "optimal"), (name, k)
                if (name, k) in DEGENERATE_CALLS:
                    continue
                assert got["iterations"] == call["iterations"], (name, k, chunk)
                assert got["trace"].tolist() == call["pivots"], (name, k, chunk)
                assert got["basic"].tolist() == call["basic_out"], (name, k, chunk)
            n += 1
    assert n >= 8


@pytest.mark.parametrize("rule", ["most_neg", "first"])
@pytest.mark.parametrize("shape", [(16, 32, 0.5), (64, 128, 0.3), (96, 200, 0.6), (150, 90, 0.4), (256, 512, 0.3),
                                   (513, 700, 0.4)])
def test_bounded_resident_engine_matches_oracle_and_fused(shape, rule):
    """Synthetic bounded LPs: the LU oracle's iteration count, pivot trace, flip flags, lists, x_B, objective; b as
    the reference leaves it; and the bits of the bounded fused kernel (x_B, y, B^-1 are the same operations in the
    same order), with one launch and with a launch every 5 iterations (flags written out and cleared at every exit)."""
    from simplex_b200 import Basis, DeviceSimplex
    m, n, tight = shape
    A, b, c, ub, basic, nonbasic = _bounded_synth(m, n, 31 + m, tight)
    want = oracle.simplex(A, b, c, basic, nonbasic, ub=ub, rule=oracle.RULES[rule], variant="lu")
    assert want["term"] == "optimal" and want["flips"] > 0
    state = {}
    for engine, chunk in (("persistent", 64), ("persistent", 5), ("streaming", 64), ("streaming", 5)):
        with DeviceSimplex(pricing=rule, engine=engine, dual="update", chunk_iters=chunk, trace_capacity=1 << 16) as s:
            s.load(A, b, c, ub)
            basis = Basis(basic, nonbasic)
            r = s.simplex(basis)
            tr, _ = s.trace()
            got = dict(r, trace=tr, flip=s.flip_flags(), basic=basis.basic, nonbasic=basis.non_basic, x=s.xB(),
                       engine_used=s.engine_used())
            if engine == "persistent":
                _same_as_oracle(got, want, (engine, chunk))
            else:
                assert s.engine_used() == "fused"
            assert rel_close(s.b(), want["b"])
            state[(engine, chunk)] = (s.xB(), s.y(), s.Binv(), s.b(), tr.tolist(), s.flip_flags().tolist())
    for chunk in (64, 5):  # (the duals are solved afresh at every launch: like is compared with like)
        got, ref = state[("persistent", chunk)], state[("streaming", chunk)]
        for a, f in zip(got[:4], ref[:4]):
            assert np.array_equal(a, f), chunk
        assert got[4:] == ref[4:], chunk


def test_bounded_resident_engine_leaves_the_tableau_as_the_reference_does():
    """A second run from the final basis on the same context (columns rewritten behind the first run, unit-column
    table still valid, flags clear) is optimal at once; primal residual and sampled rows of B^-1 A_B on the REWRITTEN
    columns are clean."""
    from simplex_b200 import Basis, DeviceSimplex
    A, b, c, ub, basic, nonbasic = _bounded_synth(96, 200, 5, 0.6)
    want = oracle.simplex(A, b, c, basic, nonbasic, ub=ub, rule=oracle.RULE_MOST_NEG, variant="lu")
    with DeviceSimplex(pricing="most_neg", dual="update", chunk_iters=7) as s:
        s.load(A, b, c, ub)
        basis = Basis(basic, nonbasic)
        r = s.simplex(basis)
        assert s.engine_used() == "resident" and r["flips"] == want["flips"] > 0
        assert rel_close(s.b(), want["b"])
        assert s.primal_residual() < 1e-9 and s.inverse_residual(16, 3) < 1e-9
        x1 = s.xB()
        r2 = s.simplex(basis)
        assert (r2["term"], r2["iterations"], r2["pivots"], r2["flips"]) == ("optimal", 1, 0, 0)
        assert rel_close(s.xB(), x1)


@pytest.mark.parametrize("rule", ["most_neg", "first"])
def test_bounded_unit_columns_on_the_resident_engine(rule):
    """Unit structural columns with small bounds (priced from the slice caches as -(c_j - v y_r) while their flag is
    up, substituted like dense ones; the unit table entry is rewritten behind every launch: a launch every 3
    iterations), against the LU oracle."""
    from simplex_b200 import synth
    m, n, k = 48, 80, 24
    A0, b, c0, basic, nonbasic0 = synth.tableau(m, n, 0, 314)
    rng = np.random.default_rng(7)
    D = np.zeros((m, k))
    rows = rng.choice(m, size=k, replace=False)
    D[rows, np.arange(k)] = rng.uniform(0.3, 1.0, k)
    A = np.asfortranarray(np.hstack([A0, D]))
    c = np.concatenate([c0, -rng.uniform(1.0, 2.0, k)])
    N = A.shape[1]
    ub = np.full(N, DBL_MAX)
    ub[:n] = rng.uniform(0.5, 30.0, n)
    ub[N - k:] = rng.uniform(0.05, 0.5, k)
    nonbasic = np.concatenate([nonbasic0, np.arange(N - k, N)])
    want = oracle.simplex(A, b, c, basic, nonbasic, ub=ub, rule=oracle.RULES[rule], variant="lu")
    assert want["term"] == "optimal" and want["flips"] > 0 and want["flip"][N - k:].sum() > 0
    for chunk in (3, 64):
        got = _run(A, b, c, basic, nonbasic, ub=ub, pricing=rule, dual="update", chunk_iters=chunk)
        _same_as_oracle(got, want, chunk)


def test_fuzz_bounded_small_shapes_on_the_resident_engine():
    """Seeded random bounded shapes (3 <= m <= 140: fewer rows than CTAs, ragged leading dimensions, Phase-I tableaux
    with artificial unit columns, bounds on slack columns), both rules, a launch every 11 iterations."""
    from simplex_b200 import synth
    rng = np.random.default_rng(4242)
    n_runs = 0
    for case in range(40):
        m = int(rng.integers(3, 141))
        n = int(rng.integers(2, 221))
        n_eq = int(rng.integers(0, max(1, m // 3))) if case % 3 == 0 else 0
        A, b, c, basic, nonbasic = synth.tableau(m, n, n_eq, 5000 + case, phase1=(n_eq > 0))
        N = A.shape[1]
        ub = np.full(N, DBL_MAX)
        ub[:n] = np.where(rng.random(n) < 0.5, rng.uniform(0.01, 0.4, n), rng.uniform(1.0, 40.0, n))
        ns = min(N - n, m)
        pick = rng.random(ns) < 0.5
        ub[n:n + ns] = np.where(pick, np.maximum(b[:ns], 0.05) * rng.uniform(1.2, 3.0, ns), DBL_MAX)
        rule = "most_neg" if case % 2 == 0 else "first"
        want = oracle.simplex(A, b, c, basic, nonbasic, ub=ub, rule=oracle.RULES[rule], variant="lu")
        got = _run(A, b, c, basic, nonbasic, ub=ub, pricing=rule, dual="update", chunk_iters=11)
        assert got["engine_used"] == "resident" and got["term"] == want["term"], (case, m, n, n_eq, rule)
        if want["x"].min() < 1e-9 and want["term"] == "optimal" and n_eq > 0:
            continue      # degenerate Phase-I vertex: ties are decided by rounding noise
        _same_as_oracle(got, want, (case, m, n, n_eq, rule))
        n_runs += 1
    assert n_runs >= 28, n_runs


def test_bounded_resident_at_the_two_phase_shape():
    """m=1024, n=2048 with finite bounds (the shape BASELINE configs[1] is quoted on), 600 iterations: the resident
    bounded kernel walks the bounded fused kernel's pivots and flips to the same bits, and is the faster of the two."""
    import time
    from simplex_b200 import Basis, DeviceSimplex
    A, b, c, ub, basic, nonbasic = _bounded_synth(1024, 2048, 1234, 0.4)
    out, secs = [], {}
    for engine in ("persistent", "streaming"):
        with DeviceSimplex(pricing="most_neg", engine=engine, dual="update", iter_limit=600, trace_capacity=600,
                           chunk_iters=256) as s:
            s.load(A, b, c, ub)
            basis = Basis(basic, nonbasic)
            r = s.simplex(basis)
            tr, _ = s.trace()
            assert s.engine_used() == ("resident" if engine == "persistent" else "fused")
            assert r["term"] == "iter_limit" and r["flips"] > 0 and r["pivots"] + r["flips"] == 600
            assert s.inverse_residual(16, 3) < 1e-9 and s.primal_residual() < 1e-9
            out.append((r["pivots"], r["flips"], tr.tolist(), s.flip_flags().tolist(), basis.basic.tolist(),
                        s.xB().tobytes(), s.y().tobytes()))
            secs[engine] = r["loop_seconds"] / 600
    assert out[0] == out[1]
    print("bounded m=1024 n=2048: resident %.2f us/iteration, fused %.2f us/iteration"
          % (1e6 * secs["persistent"], 1e6 * secs["streaming"]))


# ---- the stream-columns variant of the resident kernel (rows of B^-1 on chip, dense columns priced out of L2)

@pytest.mark.parametrize("bounded", [False, True])
@pytest.mark.parametrize("shape", [(5, 7, 0), (96, 200, 24), (300, 640, 0), (48, 6000, 8)])
def test_resident_stream_columns_is_bit_identical(shape, bounded, monkeypatch):
    """k_resident_stream / k_resident_stream_bounded forced on small shapes (SB200_RES_STREAM_COLS=2): the pricing dots
    read the columns from global memory with the lane partition and summation order of the on-chip ones, so trace,
    lists, x_B, y and B^-1 are the bits of the fused kernel; one launch and a launch every 7 iterations."""
    from simplex_b200 import Basis, DeviceSimplex, synth
    m, n, n_eq = shape
    A, b, c, basic, nonbasic = synth.tableau(m, n, n_eq, 77, phase1=(n_eq > 0))
    ub = None
    if bounded:
        rng = np.random.default_rng(m)
        ub = np.full(A.shape[1], DBL_MAX)
        ub[:n] = np.where(rng.random(n) < 0.4, rng.uniform(0.01, 0.4, n), rng.uniform(1.0, 40.0, n))
    monkeypatch.setenv("SB200_RES_STREAM_COLS", "2")
    for chunk in (7, 4096):
        state = {}
        for engine in ("persistent", "streaming"):
            with DeviceSimplex(pricing="most_neg", engine=engine, dual="update", chunk_iters=chunk, trace_capacity=1 << 16) as s:
                s.load(A, b, c, ub)
                basis = Basis(basic, nonbasic)
                r = s.simplex(basis)
                tr, _ = s.trace()
                assert s.engine_used() == ("resident" if engine == "persistent" else "fused")
                state[engine] = ((r["term"], r["iterations"], r["pivots"], r["flips"], r["objective"], tr.tolist(),
                                  basis.basic.tolist(), basis.non_basic.tolist(), s.flip_flags().tolist()),
                                 (s.xB(), s.y(), s.Binv(), s.b()))
        assert state["persistent"][0] == state["streaming"][0], chunk
        for a, f in zip(state["persistent"][1], state["streaming"][1]):
            assert np.array_equal(a, f), chunk


def test_resident_stream_columns_at_a_mid_size_shape():
    """m=1536, n=3072: 28 rows + columns of 12 KB per SM do not fit the shared memory, the 11 rows alone do, and A
    (57 MB) lives in L2: the default engine is the stream-columns resident kernel. 1500 pivots, bit-identical to the
    fused kernel, and faster."""
    from simplex_b200 import Basis, DeviceSimplex, synth
    A, b, c, basic, nonbasic = synth.tableau(1536, 3072, 0, 5)
    out, secs = {}, {}
    for engine in ("persistent", "streaming"):
        with DeviceSimplex(pricing="most_neg", engine=engine, dual="update", iter_limit=1500, trace_capacity=1500,
                           chunk_iters=512) as s:
            s.load(A, b, c)
            basis = Basis(basic, nonbasic)
            r = s.simplex(basis)
            tr, _ = s.trace()
            assert s.engine_used() == ("resident" if engine == "persistent" else "fused")
            assert r["term"] == "iter_limit" and r["pivots"] == 1500
            assert s.inverse_residual(16, 3) < 1e-9 and s.primal_residual() < 1e-9
            out[engine] = (tr.tolist(), basis.basic.tolist(), s.xB().tobytes(), s.y().tobytes())
            secs[engine] = r["loop_seconds"] / 1500
    assert out["persistent"] == out["streaming"]
    print("m=1536 n=3072: resident (columns from L2) %.2f us/pivot, fused %.2f us/pivot"
          % (1e6 * secs["persistent"], 1e6 * secs["streaming"]))
