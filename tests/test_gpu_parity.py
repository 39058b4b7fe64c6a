"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's golden
vectors. Everything here needs a GPU: run with `-m gpu` on the B200 box."""
import os

import numpy as np
import pytest

from helpers import DEGENERATE_CALLS, call_inputs, has_bounds, rel_close
from oracle import oracle

pytestmark = pytest.mark.gpu

# ("persistent", "update") picks the shared-memory-resident kernel whenever the tableau fits on chip
# (every shape below m ~ 1100); ("streaming", "update") is the same engine pinned to the HBM-streaming
# fused kernel, which is what the headline shape runs.
ENGINE_DUAL = [("staged", "recompute"), ("staged", "update"), ("persistent", "recompute"), ("persistent", "update"),
               ("streaming", "update")]


def _device_solver(**kw):
    from simplex_b200 import DeviceSimplex
    return DeviceSimplex(**kw)


def _run(A, b, c, basic, nonbasic, ub=None, **kw):
    from simplex_b200 import Basis
    with _device_solver(trace_capacity=1 << 16, **kw) as s:
        s.load(A, b, c, ub)
        basis = Basis(basic, nonbasic)
        res = s.simplex(basis)
        res["basic"], res["nonbasic"] = basis.basic, basis.non_basic
        res["x"] = s.xB()
        res["trace"], res["n_trace"] = s.trace()
        res["flip"] = s.flip_flags()
        res["launches"] = s.launch_count()
        res["engine_used"] = s.engine_used()
    return res


@pytest.mark.parametrize("engine,dual", ENGINE_DUAL)
def test_fixture_calls_match_reference(golden_fixtures, engine, dual):
    """All recorded Simplex::simplex() calls of the reference's 14 fixtures (+ Klee-Minty n=10,
    1023 pivots): same status, iteration count, pivot trace, final lists and flip flags; x_B within
    1e-9 of the reference's LU-based oracle. Degenerate calls: status only."""
    for name, rec in golden_fixtures.items():
        bounded = has_bounds(rec)
        for k, call in enumerate(rec["calls"]):
            A, b, c, ub, basic, nonbasic = call_inputs(call)
            ubx = ub if bounded else None
            got = _run(A, b, c, basic, nonbasic, ub=ubx, pricing="first", engine=engine, dual=dual)
            want_term = "unbounded" if call["unbounded"] else "optimal"
            assert got["term"] == want_term, (name, k)
            assert got["launches"] > 0
            if bounded and dual == "update" and engine != "staged":
                # bounded ratio test + bound flips on the fast engines
                assert got["engine_used"] == ("fused" if engine == "streaming" else "resident"), (name, k)
            if (name, k) in DEGENERATE_CALLS:
                continue
            assert got["iterations"] == call["iterations"], (name, k)
            assert got["trace"].tolist() == call["pivots"], (name, k)
            assert got["basic"].tolist() == call["basic_out"], (name, k)
            assert got["nonbasic"].tolist() == call["nonbasic_out"], (name, k)
            ref = oracle.simplex(A, b, c, basic, nonbasic, ub=ubx, rule=oracle.RULE_FIRST, variant="lu")
            assert got["flip"].tolist() == ref["flip"].tolist(), (name, k)
            if want_term == "optimal":
                assert rel_close(got["x"], ref["x"]), (name, k)
                assert rel_close(got["objective"], ref["objective"]), (name, k)


def _two_phase_gpu(rec, engine, dual, resident):
    """Phase I then Phase II as the reference's InitialBasis / Simplex::solve sequence them."""
    from simplex_b200 import Basis, synth
    m, n, n_eq, seed, rule = rec["m"], rec["n"], rec["n_eq"], rec["seed"], rec["rule"]
    out = []
    with _device_solver(pricing=rule, engine=engine, dual=dual, trace_capacity=1 << 16) as s:
        if n_eq > 0:
            A1, b1, c1, basic, nonbasic = synth.tableau(m, n, n_eq, seed, phase1=True)
            s.load(A1, b1, c1)
            basis = Basis(basic, nonbasic)
            r = s.simplex(basis)
            r["trace"], _ = s.trace()
            r["basic"], r["nonbasic"] = basis.basic.copy(), basis.non_basic.copy()
            out.append(r)
            basic, nonbasic = synth.phase2_lists(basis.basic, basis.non_basic, n + (m - n_eq))
            A2, b2, c2, _, _ = synth.tableau(m, n, n_eq, seed, phase1=False)
            if resident:
                assert np.array_equal(A2, A1[:, :A2.shape[1]]) and np.array_equal(b2, b1)
                s.set_costs(c2)       # A stays in HBM; artificial columns are a trailing block
            else:
                s.load(A2, b2, c2)
        else:
            A2, b2, c2, basic, nonbasic = synth.tableau(m, n, n_eq, seed, phase1=False)
            s.load(A2, b2, c2)
        basis = Basis(basic, nonbasic)
        r = s.simplex(basis)
        r["trace"], _ = s.trace()
        r["basic"], r["nonbasic"] = basis.basic.copy(), basis.non_basic.copy()
        r["x"] = s.xB()
        out.append(r)
    return out


@pytest.mark.parametrize("engine,dual", ENGINE_DUAL)
def test_synth_traces_match_reference(golden_synth, engine, dual):
    """Synthetic LPs G(m,n,n_eq,seed) up to m=512 for both pricing rules: pivot sequence,
    iteration count and final basis identical to the reference, objective and structural x
    within 1e-9 relative (north_star's parity bar)."""
    n_checked = 0
    for rec in golden_synth:
        res = _two_phase_gpu(rec, engine, dual, resident=(rec["m"] % 64 == 0))
        tag = (rec["m"], rec["n"], rec["n_eq"], rec["seed"], rec["rule"])
        assert len(res) == len(rec["calls"]), tag
        for k, (r, call) in enumerate(zip(res, rec["calls"])):
            assert r["term"] == "optimal", (tag, k)
            assert r["iterations"] == call["iterations"], (tag, k)
            assert r["trace"].tolist() == call["pivots"], (tag, k)
            assert r["basic"].tolist() == call["basic_out"], (tag, k)
            assert r["nonbasic"].tolist() == call["nonbasic_out"], (tag, k)
        last = res[-1]
        assert rel_close(last["objective"], rec["objective"]), tag
        x = np.zeros(rec["n"])
        for i, col in enumerate(last["basic"]):
            if col < rec["n"]:
                x[col] = last["x"][i]
        assert rel_close(x, rec["x_struct"]), tag
        n_checked += 1
    assert n_checked >= 20


def test_single_phase_kernels_against_numpy():
    """Each kernel of one pivot through its own C-ABI entry point (K0 prepare, K1 dual,
    K2 pricing + arg-min, K3 FTRAN, K4 ratio, K5 rank-1 update) against numpy fp64."""
    from simplex_b200 import Basis, synth
    m, n = 96, 200
    A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 99)
    for pricing in ("first", "most_neg"):
        with _device_solver(pricing=pricing, engine="staged", dual="recompute") as s:
            s.load(A, b, c)
            basis = Basis(basic, nonbasic)
            s.prepare(basis)
            bas, non = basis.basic.copy(), basis.non_basic.copy()
            Binv = np.linalg.inv(A[:, bas])
            for it in range(12):
                s.step_dual()
                y = c[bas] @ Binv
                assert np.allclose(s.y(), y, rtol=1e-12, atol=1e-13)
                sred = c[non] - A[:, non].T @ y
                pos, val = s.step_price()
                if pricing == "first":
                    elig = np.nonzero(sred < -1e-7)[0]
                    want = int(elig[0]) if elig.size else -1
                else:
                    want = int(np.argmin(sred)) if sred.min() < -1e-7 else -1
                assert pos == want
                if pos < 0:
                    break
                assert abs(val - sred[pos]) < 1e-12
                s.step_ftran(pos)
                d = Binv @ A[:, non[pos]]
                assert np.allclose(s.d(), d, rtol=1e-12, atol=1e-13)
                x = Binv @ b
                assert np.allclose(s.xB(), x, rtol=1e-11, atol=1e-12)
                ratios = np.where(d > 1e-7, x / np.where(d > 1e-7, d, 1.0), np.inf)
                row, theta = s.step_ratio(pos)
                assert row == int(np.argmin(ratios))
                assert abs(theta - ratios[row]) <= 1e-12 * max(1.0, abs(theta))
                s.step_update()
                s.step_update()  # second call must be a no-op
                bas[row], non[pos] = non[pos], bas[row]
                Binv = np.linalg.inv(A[:, bas])
                assert np.allclose(s.Binv(), Binv, rtol=1e-9, atol=1e-10)


def test_terminations_and_errors():
    from simplex_b200 import Basis, DeviceSimplex, Sb200Error, synth
    A, b, c, basic, nonbasic = synth.tableau(32, 64, 0, 5)
    for engine in ("staged", "persistent"):
        # iteration limit: the reference throws after the loop (Simplex.cpp:347-350)
        with DeviceSimplex(pricing="most_neg", engine=engine, iter_limit=7, chunk_iters=4) as s:
            s.load(A, b, c)
            r = s.simplex(Basis(basic, nonbasic))
            assert r["term"] == "iter_limit" and r["iterations"] == 7 and r["pivots"] == 7
        # unbounded: min -x1 with a row that does not bound it
        Au = np.array([[-1.0, 1.0, 1.0, 0.0], [0.0, 1.0, 0.0, 1.0]])
        with DeviceSimplex(engine=engine) as s:
            s.load(Au, np.array([1.0, 2.0]), np.array([-1.0, 0.0, 0.0, 0.0]))
            r = s.simplex(Basis([2, 3], [0, 1]))
            assert r["term"] == "unbounded"
        # wrong basis size: the reference throws std::string (Simplex.cpp:260-263)
        with DeviceSimplex(engine=engine) as s:
            s.load(A, b, c)
            with pytest.raises(Sb200Error) as ei:
                s.simplex(Basis(basic[:-1], np.append(nonbasic, basic[-1])))
            assert "Basis size must be equal to M=32" in str(ei.value)
        # already optimal start: one iteration, zero pivots
        with DeviceSimplex(engine=engine) as s:
            s.load(A, b, -c)
            r = s.simplex(Basis(basic, nonbasic))
            assert r["term"] == "optimal" and r["iterations"] == 1 and r["pivots"] == 0


def test_general_start_basis_is_inverted_on_device():
    """Phase II normally starts from a non-trivial basis: K0 must invert it on the device."""
    from simplex_b200 import Basis, DeviceSimplex, synth
    m, n = 48, 96
    A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 3)
    ref = oracle.simplex(A, b, c, basic, nonbasic, rule=oracle.RULE_MOST_NEG, iter_limit=20, variant="lu")
    bas, non = ref["basic"], ref["nonbasic"]  # a basis 20 pivots into the solve
    full = oracle.simplex(A, b, c, bas, non, rule=oracle.RULE_MOST_NEG, variant="lu")
    for engine in ("staged", "persistent"):
        with DeviceSimplex(pricing="most_neg", engine=engine, trace_capacity=4096) as s:
            s.load(A, b, c)
            basis = Basis(bas, non)
            r = s.simplex(basis)
            tr, _ = s.trace()
            assert r["term"] == "optimal" and r["iterations"] == full["iterations"]
            assert tr.tolist() == full["trace"].tolist()
            assert rel_close(r["objective"], full["objective"])


def _batch_of(m, n, batch, seed0):
    from simplex_b200 import synth
    A, b, c, basic, nonbasic = synth.batch(batch, m, n, seed0)
    cols = [synth.tableau(m, n, 0, seed0 + k) for k in range(batch)]
    for k in range(batch):
        assert np.array_equal(A[k].T, cols[k][0])
    return A, b, c, basic, nonbasic, cols


@pytest.mark.parametrize("kernel", ["fast", "fast_nocache", "fast_hybrid", "generic"])
@pytest.mark.parametrize("m,n,rule", [(64, 128, "most_neg"), (64, 128, "first"), (48, 80, "most_neg"), (17, 40, "first"),
                                      (32, 200, "most_neg")])
def test_batched_matches_oracle(kernel, m, n, rule, monkeypatch):
    """K7: independent small LPs, one CTA each (config 5 shape m=64, n=128 and smaller, padded
    shapes), both kernels (registers-resident fast kernel / generic shared-memory kernel), both
    entering rules: every LP walks the oracle's pivot sequence to the same basis and x_B."""
    from simplex_b200 import DeviceSimplex
    if kernel == "generic":
        monkeypatch.setenv("SB200_BATCHED_GENERIC", "1")
    if kernel == "fast_hybrid":
        monkeypatch.setenv("SB200_BATCHED_HYBRID", "1")    # unit columns implicit even when all columns would fit
    if kernel == "fast_nocache":
        monkeypatch.setenv("SB200_BATCHED_NOCACHE", "1")   # more than one pricing trip per pivot is also covered by n=200
    batch = 24
    A, b, c, basic, nonbasic, cols = _batch_of(m, n, batch, 1234)
    orule = oracle.RULE_MOST_NEG if rule == "most_neg" else oracle.RULE_FIRST
    want = [oracle.simplex(Ak, bk, ck, bas, non, rule=orule, variant="lu") for Ak, bk, ck, bas, non in cols]
    with DeviceSimplex(pricing=rule) as s:
        r = s.run_batched(A, b, c, basic, nonbasic)
        assert s.launch_count() == 1
    for k in range(batch):
        assert r["term"][k] == 1, k
        assert r["iterations"][k] == want[k]["iterations"], k
        assert basic[k].tolist() == want[k]["basic"].tolist(), k
        assert nonbasic[k].tolist() == want[k]["nonbasic"].tolist(), k
        assert rel_close(r["objective"][k], want[k]["objective"]), k
        assert rel_close(r["xB"][k], want[k]["x"]), k


def test_batched_fast_kernel_hands_over_what_it_cannot_take():
    """An LP whose start basis is not made of unit columns is left to the generic kernel by the
    fast one (TERM_PENDING); the mixed batch still gives the oracle's result for every LP."""
    from simplex_b200 import DeviceSimplex
    m, n, batch = 32, 64, 8
    A, b, c, basic, nonbasic, cols = _batch_of(m, n, batch, 77)
    want = []
    for k, (Ak, bk, ck, bas, non) in enumerate(cols):
        if k % 2 == 1:
            # start from the optimal basis of this LP instead of the slack basis
            first = oracle.simplex(Ak, bk, ck, bas, non, rule=oracle.RULE_MOST_NEG, variant="lu")
            bas, non = first["basic"], first["nonbasic"]
            basic[k], nonbasic[k] = bas, non
        want.append(oracle.simplex(Ak, bk, ck, bas, non, rule=oracle.RULE_MOST_NEG, variant="lu"))
    with DeviceSimplex(pricing="most_neg") as s:
        r = s.run_batched(A, b, c, basic, nonbasic)
        assert s.launch_count() == 2  # fast kernel + generic kernel for the four pending LPs
    for k in range(batch):
        assert r["term"][k] == 1, k
        assert r["iterations"][k] == want[k]["iterations"], k
        assert basic[k].tolist() == want[k]["basic"].tolist(), k
        assert rel_close(r["objective"][k], want[k]["objective"]), k
        assert rel_close(r["xB"][k], want[k]["x"]), k


def test_batched_unbounded_and_iteration_limit():
    """Terminations of the batched kernels: an unbounded LP (no positive pivot entry) and the
    iteration cap."""
    from simplex_b200 import DeviceSimplex
    m, n, batch = 8, 8, 2
    N = n + m
    A = np.zeros((batch, N, m))
    for k in range(batch):
        A[k, :n, :] = -1.0 if k == 0 else 0.5     # LP 0: every structural column <= 0 -> unbounded
        A[k, n:, :] = np.eye(m)
    b = np.ones((batch, m))
    c = np.zeros((batch, N))
    c[:, :n] = -1.0
    for kernel_env in (None, "1"):
        basic = np.tile(np.arange(n, N, dtype=np.int32), (batch, 1))
        nonbasic = np.tile(np.arange(n, dtype=np.int32), (batch, 1))
        if kernel_env:
            os.environ["SB200_BATCHED_GENERIC"] = kernel_env
        try:
            with DeviceSimplex(pricing="most_neg") as s:
                r = s.run_batched(A, b, c, basic, nonbasic)
            assert r["term"].tolist() == [2, 1]
            with DeviceSimplex(pricing="most_neg", iter_limit=1) as s:
                basic2 = np.tile(np.arange(n, N, dtype=np.int32), (batch, 1))
                nonbasic2 = np.tile(np.arange(n, dtype=np.int32), (batch, 1))
                r = s.run_batched(A, b, c, basic2, nonbasic2)
            assert r["term"][1] == 3
        finally:
            os.environ.pop("SB200_BATCHED_GENERIC", None)


@pytest.mark.parametrize("engine,dual", [("persistent", "update"), ("streaming", "update"), ("persistent", "recompute"),
                                         ("staged", "recompute")])
def test_periodic_reinversion_keeps_the_pivot_sequence(golden_synth, engine, dual):
    """sb200_opts.reinvert_period: B^-1 and x_B are rebuilt from the resident A_B every 20 pivots;
    the run still walks the reference's pivot sequence and ends with a tiny primal residual."""
    from simplex_b200 import Basis, DeviceSimplex, synth
    rec = next(r for r in golden_synth if (r["m"], r["n"], r["n_eq"], r["rule"]) == (128, 256, 0, "dantzig"))
    A, b, c, basic, nonbasic = synth.tableau(128, 256, 0, rec["seed"])
    call = rec["calls"][0]
    with DeviceSimplex(pricing="most_neg", engine=engine, dual=dual, chunk_iters=10, reinvert_period=20,
                       trace_capacity=4096) as s:
        s.load(A, b, c)
        basis = Basis(basic, nonbasic)
        r = s.simplex(basis)
        trace, _ = s.trace()
        assert r["term"] == "optimal" and r["iterations"] == call["iterations"]
        assert trace.tolist() == call["pivots"]
        assert basis.basic.tolist() == call["basic_out"]
        assert s.reinversion_count() >= len(call["pivots"]) // 20 - 1
        assert s.primal_residual() < 1e-10
        assert rel_close(r["objective"], rec["objective"])


@pytest.mark.parametrize("kernel_env", [None, "SB200_BATCHED_GENERIC", "SB200_BATCHED_HYBRID"])
def test_batched_degenerate_shapes(kernel_env, monkeypatch):
    """Smallest shapes: one row / one column, and no non-basic column at all (N == m): optimal at once."""
    from simplex_b200 import DeviceSimplex
    if kernel_env:
        monkeypatch.setenv(kernel_env, "1")
    # m = 1, n = 1: min -x  s.t. 2x + s = 3  ->  x = 1.5, objective -1.5, one pivot
    A = np.array([[[2.0], [1.0]]])           # (batch=1, N=2, m=1)
    b = np.array([[3.0]])
    c = np.array([[-1.0, 0.0]])
    basic = np.array([[1]], dtype=np.int32)
    nonbasic = np.array([[0]], dtype=np.int32)
    with DeviceSimplex(pricing="most_neg") as s:
        r = s.run_batched(A, b, c, basic, nonbasic)
    assert r["term"].tolist() == [1] and r["iterations"].tolist() == [2]
    assert basic.tolist() == [[0]] and nonbasic.tolist() == [[1]]
    assert abs(r["xB"][0, 0] - 1.5) < 1e-15 and abs(r["objective"][0] + 1.5) < 1e-15
    # N == m: nothing to price
    A = np.tile(np.eye(3)[None], (2, 1, 1)) * 2.0
    b = np.array([[2.0, 4.0, 6.0], [1.0, 1.0, 1.0]])
    c = np.ones((2, 3))
    basic = np.tile(np.arange(3, dtype=np.int32), (2, 1))
    nonbasic = np.zeros((2, 0), dtype=np.int32)
    with DeviceSimplex(pricing="first") as s:
        r = s.run_batched(A, b, c, basic, nonbasic)
    assert r["term"].tolist() == [1, 1] and r["iterations"].tolist() == [1, 1]
    assert np.allclose(r["xB"], b / 2.0) and np.allclose(r["objective"], [6.0, 1.5])


def test_invariants_at_scale():
    """Size-independent properties at a size the LU oracle cannot reach in test time
    (m=1024, n=2048 — BASELINE config 2 shape): after P pivots B^-1 A_B = I, x_B = B^-1 b >= 0,
    and both engines walk the same pivot sequence."""
    from simplex_b200 import Basis, DeviceSimplex, synth
    m, n, P = 1024, 2048, 300
    A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 1234)
    traces = []
    for engine, dual in (("staged", "recompute"), ("persistent", "recompute"), ("persistent", "update"),
                         ("streaming", "update")):
        with DeviceSimplex(pricing="most_neg", engine=engine, dual=dual, iter_limit=P, trace_capacity=P) as s:
            s.load(A, b, c)
            basis = Basis(basic, nonbasic)
            r = s.simplex(basis)
            assert r["term"] == "iter_limit" and r["pivots"] == P
            Binv, x = s.Binv(), s.xB()
            AB = A[:, basis.basic]
            assert np.abs(Binv @ AB - np.eye(m)).max() < 1e-9
            assert np.allclose(x, Binv @ b, rtol=1e-9, atol=1e-10)
            assert x.min() > -1e-9
            assert sorted(basis.basic.tolist() + basis.non_basic.tolist()) == list(range(A.shape[1]))
            tr, _ = s.trace()
            traces.append(tr.tolist())
    assert traces[0] == traces[1] == traces[2] == traces[3]


def test_headline_size_invariants():
    """BASELINE configs[2] at full size (m=4096, n=16384, 640 MiB of A in HBM), 1500 pivots: the two
    dual modes of the persistent engine walk the same pivot sequence; the lists stay a permutation;
    x_B = B^-1 b >= 0 with a tiny primal residual; sampled rows of B^-1 A_B are rows of I; the
    objective never increases along the trace."""
    from simplex_b200 import Basis, DeviceSimplex, synth
    m, n, P = 4096, 16384, 1500
    A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 1234)
    traces, objs = [], []
    for dual in ("update", "recompute"):
        with DeviceSimplex(pricing="most_neg", engine="persistent", dual=dual, iter_limit=P, chunk_iters=250,
                           trace_capacity=P) as s:
            s.load(A, b, c)
            basis = Basis(basic, nonbasic)
            r = s.simplex(basis)
            assert r["term"] == "iter_limit" and r["pivots"] == P
            tr, _ = s.trace()
            traces.append(tr.tolist())
            objs.append(r["objective"])
            x = s.xB()
            assert x.min() > -1e-9
            assert s.primal_residual() < 1e-9
            assert sorted(basis.basic.tolist() + basis.non_basic.tolist()) == list(range(A.shape[1]))
            if dual == "update":
                Binv = s.Binv()
                rows = np.random.default_rng(0).choice(m, size=48, replace=False)
                AB = A[:, basis.basic]
                got = Binv[rows] @ AB
                want = np.zeros_like(got)
                want[np.arange(rows.size), rows] = 1.0
                assert np.abs(got - want).max() < 1e-9
                # x_B is carried by the recurrence x -= theta * d: compare on the scale of b
                assert np.abs(x - Binv @ b).max() < 1e-11 * np.abs(b).max()
    assert traces[0] == traces[1]
    assert rel_close(objs[0], objs[1])
    assert objs[0] < 0.0


def test_batched_kernels_agree_at_scale(monkeypatch):
    """8192 LPs of the configs[4] shape: the register-resident kernel and the generic
    shared-memory kernel end every LP after the same number of iterations on the same basis."""
    from simplex_b200 import DeviceSimplex, synth
    m, n, count = 64, 128, 8192
    A, b, c, basic, nonbasic = synth.batch(count, m, n, 99)
    out = {}
    for kernel in ("fast", "generic"):
        if kernel == "generic":
            monkeypatch.setenv("SB200_BATCHED_GENERIC", "1")
        bas, non = basic.copy(), nonbasic.copy()
        with DeviceSimplex(pricing="most_neg") as s:
            r = s.run_batched(A, b, c, bas, non)
        out[kernel] = (r, bas, non)
    rf, bf, nf = out["fast"]
    rg, bg, ng = out["generic"]
    assert (rf["term"] == 1).all() and (rg["term"] == 1).all()
    assert np.array_equal(rf["iterations"], rg["iterations"])
    assert np.array_equal(bf, bg) and np.array_equal(nf, ng)
    assert rel_close(rf["objective"], rg["objective"])
    assert rel_close(rf["xB"], rg["xB"])
    assert 40 < rf["iterations"].mean() < 80


def test_sharded_engine_single_rank_matches_reference(golden_synth):
    """The sharded kernel (peer-window exchanges, chained dual solve) with world = 1: every
    all-'<=' synthetic golden case walks the reference's pivot sequence."""
    from simplex_b200 import Basis, synth
    from simplex_b200.dist import ShardedSimplex
    n_checked = 0
    for rec in golden_synth:
        if rec["n_eq"] != 0:
            continue
        m, n = rec["m"], rec["n"]
        A, b, c, basic, nonbasic = synth.tableau(m, n, 0, rec["seed"])
        with ShardedSimplex(0, 1, pricing=rec["rule"], device=0, trace_capacity=1 << 16, chunk_iters=32,
                            iter_limit=10 ** 6) as s:
            s.load_shard(m, A.shape[1], A, b, c)
            basis = Basis(basic, nonbasic)
            r = s.simplex(basis)
            tr, _ = s.trace()
            x = s.xB()
        call = rec["calls"][-1]
        assert r["term"] == "optimal" and r["iterations"] == call["iterations"], (m, n, rec["rule"])
        assert tr.tolist() == call["pivots"] and basis.basic.tolist() == call["basic_out"]
        assert rel_close(r["objective"], rec["objective"])
        xs = np.zeros(n)
        for i, col in enumerate(basis.basic):
            if col < n:
                xs[col] = x[i]
        assert rel_close(xs, rec["x_struct"])
        n_checked += 1
    assert n_checked >= 8


def test_sharded_engine_two_gpus_same_pivot_sequence():
    """One LP over 2 GPUs (cyclic column shards, row-sharded B^-1, NVLink exchanges): pivot trace,
    final lists and x_B bits identical to the single-GPU run (G-independence)."""
    import json
    import os
    import subprocess
    import sys

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rows, cols, iters, chunk in ((256, 1024, 5000, 64), (1024, 4096, 600, 128)):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                            "--master-addr", "127.0.0.1", "--master-port", "29517",
                            os.path.join(root, "scripts", "sharded_check.py"), "--rows", str(rows), "--cols", str(cols),
                            "--iters", str(iters), "--chunk", str(chunk)], capture_output=True, text=True, timeout=280)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        assert out["same_trace"] and out["same_lists"] and out["same_x_bits"], out


def _engine_run(engine, A, b, c, basic, nonbasic, rule, **kw):
    from simplex_b200 import Basis, DeviceSimplex
    with DeviceSimplex(pricing=rule, engine=engine, dual="update", trace_capacity=1 << 16, **kw) as s:
        s.load(A, b, c)
        basis = Basis(basic, nonbasic)
        r = s.simplex(basis)
        tr, _ = s.trace()
        return {"engine": s.engine_used(), "term": r["term"], "iterations": r["iterations"], "pivots": r["pivots"],
                "objective": r["objective"], "trace": tr.tolist(), "basic": basis.basic.tolist(),
                "nonbasic": basis.non_basic.tolist(), "x": s.xB(), "y": s.y(), "Binv": s.Binv()}


@pytest.mark.parametrize("rule", ["most_neg", "first"])
@pytest.mark.parametrize("shape", [(5, 7, 0), (64, 128, 0), (96, 200, 24), (150, 90, 0), (300, 640, 0), (513, 700, 100),
                                   (48, 6000, 8)])
def test_resident_engine_is_bit_identical_to_streaming(shape, rule):
    """The shared-memory-resident kernel (tableau spread over the SMs' shared memory, two slot
    exchanges per pivot) does the arithmetic of the HBM-streaming fused kernel operation by operation:
    same pivot trace, same lists, bit-identical x_B, y and B^-1 -- for a launch per 7 pivots (state
    written back and re-read at every launch boundary) and for one long launch. Shapes include fewer
    rows than CTAs, ragged leading dimensions, Phase-I tableaux with artificial unit columns and a wide LP
    (40 dense columns and 41 positions per CTA: several trips of the pricing loops, slices on two warps)."""
    from simplex_b200 import synth
    m, n, n_eq = shape
    A, b, c, basic, nonbasic = synth.tableau(m, n, n_eq, 77, phase1=(n_eq > 0))
    for chunk in (7, 4096):   # y is refreshed by a full dual solve at every launch: compare like with like
        want = _engine_run("streaming", A, b, c, basic, nonbasic, rule, chunk_iters=chunk)
        assert want["engine"] == "fused" and want["term"] == "optimal"
        got = _engine_run("persistent", A, b, c, basic, nonbasic, rule, chunk_iters=chunk)
        assert got["engine"] == "resident", chunk
        for key in ("term", "iterations", "pivots", "trace", "basic", "nonbasic"):
            assert got[key] == want[key], (key, chunk)
        for key in ("x", "y", "Binv"):
            assert np.array_equal(got[key], want[key]), (key, chunk)
        assert got["objective"] == want["objective"]


def test_resident_engine_sequence_numbers_wrap(monkeypatch):
    """The mail words carry a 32-bit pivot sequence number that never repeats within a context; before it
    runs out the library wipes the mail and restarts the count. Started 20 pivots before that point, with a
    launch every 7 pivots, the run crosses it and still walks the streaming engine's pivot sequence."""
    from simplex_b200 import synth
    A, b, c, basic, nonbasic = synth.tableau(96, 200, 0, 11)
    want = _engine_run("streaming", A, b, c, basic, nonbasic, "most_neg", chunk_iters=7)
    monkeypatch.setenv("SB200_RES_SEQ0", str(0xF0000000 - 20))
    got = _engine_run("persistent", A, b, c, basic, nonbasic, "most_neg", chunk_iters=7)
    assert got["engine"] == "resident" and want["pivots"] > 40
    for key in ("term", "iterations", "pivots", "trace", "basic", "nonbasic"):
        assert got[key] == want[key], key
    for key in ("x", "y", "Binv"):
        assert np.array_equal(got[key], want[key]), key


def test_resident_engine_terminations():
    """Iteration limit in the middle of a launch, an unbounded LP (no ratio candidate in any CTA) and a
    start basis that is already optimal, all through the resident kernel."""
    from simplex_b200 import Basis, DeviceSimplex, synth
    A, b, c, basic, nonbasic = synth.tableau(64, 128, 0, 5)
    with DeviceSimplex(pricing="most_neg", engine="persistent", dual="update", iter_limit=7, chunk_iters=4) as s:
        s.load(A, b, c)
        r = s.simplex(Basis(basic, nonbasic))
        assert (s.engine_used(), r["term"], r["iterations"]) == ("resident", "iter_limit", 7)
    # min -x0 s.t. x0 - x1 + s = 1: the ray x0 = x1 -> inf is feasible
    A2 = np.array([[1.0, -1.0, 1.0, 0.0], [-1.0, -2.0, 0.0, 1.0]])
    with DeviceSimplex(pricing="most_neg", engine="persistent", dual="update") as s:
        s.load(A2, np.array([1.0, 1.0]), np.array([-1.0, 0.0, 0.0, 0.0]))
        r = s.simplex(Basis(np.array([2, 3]), np.array([0, 1])))
        assert (s.engine_used(), r["term"]) == ("resident", "unbounded")
    with DeviceSimplex(pricing="first", engine="persistent", dual="update") as s:
        s.load(A2, np.array([1.0, 1.0]), np.array([1.0, 1.0, 0.0, 0.0]))
        r = s.simplex(Basis(np.array([2, 3]), np.array([0, 1])))
        assert (s.engine_used(), r["term"], r["iterations"], r["pivots"]) == ("resident", "optimal", 1, 0)


def test_engine_choice_follows_the_tableau():
    """sb200_engine_used: a tableau that fits the chip's shared memory -> resident, with or without finite bounds
    (bounded variants of both kernels), also when only the rows of B^-1 fit and A lives in L2; dual recompute ->
    persistent; rows of B^-1 larger than the chip's shared memory -> fused; SB200_ENGINE_STREAMING -> fused."""
    from simplex_b200 import Basis, DeviceSimplex, synth
    A, b, c, basic, nonbasic = synth.tableau(64, 128, 0, 5)
    ub = np.full(A.shape[1], 50.0)
    for kw, ubx, want in ((dict(engine="persistent", dual="update"), None, "resident"),
                          (dict(engine="streaming", dual="update"), None, "fused"),
                          (dict(engine="persistent", dual="recompute"), None, "persistent"),
                          (dict(engine="persistent", dual="update"), ub, "resident"),
                          (dict(engine="streaming", dual="update"), ub, "fused"),
                          (dict(engine="persistent", dual="recompute"), ub, "persistent"),
                          (dict(engine="staged", dual="update"), None, "staged")):
        with DeviceSimplex(pricing="most_neg", iter_limit=5, **kw) as s:
            s.load(A, b, c, ubx)
            s.simplex(Basis(basic, nonbasic))
            assert s.engine_used() == want, kw
    # 28 rows+columns of 12 KB per SM > 227 KB, the 11 rows alone fit and A lives in L2: resident with the columns
    # priced out of L2 (k_resident_stream); at m=2048 the rows alone (14 x 16 KB + 3 vectors) no longer fit: fused
    for shape, want in (((1536, 3072), "resident"), ((2048, 4096), "fused")):
        A, b, c, basic, nonbasic = synth.tableau(shape[0], shape[1], 0, 5)
        with DeviceSimplex(pricing="most_neg", engine="persistent", dual="update", iter_limit=5) as s:
            s.load(A, b, c)
            s.simplex(Basis(basic, nonbasic))
            assert s.engine_used() == want, shape


def _bounded_synth(m, n, seed, tight):
    """G(m, n, 0, seed) with finite upper bounds on the structural columns (a share `tight` of them small
    enough to be hit) and on a few slacks: from the all-slack start every column sits at its lower bound, so
    the start is feasible whatever the bounds are (reference src/Simplex.cpp:116-179 is the path under test)."""
    from simplex_b200 import synth
    from helpers import DBL_MAX
    A, b, c, basic, nonbasic = synth.tableau(m, n, 0, seed)
    rng = np.random.default_rng(seed)
    ub = np.full(A.shape[1], DBL_MAX)
    ub[:n] = np.where(rng.random(n) < tight, rng.uniform(0.01, 0.3, n), rng.uniform(1.0, 50.0, n))
    big = rng.choice(m, size=max(1, m // 8), replace=False)
    ub[n + big] = b[big] * rng.uniform(1.5, 4.0, big.size)   # basic at the start: above their values
    return A, b, c, ub, basic, nonbasic


@pytest.mark.parametrize("rule", ["most_neg", "first"])
@pytest.mark.parametrize("shape", [(16, 32, 0.5), (64, 128, 0.3), (96, 200, 0.6), (150, 90, 0.4), (256, 512, 0.3)])
def test_bounded_fused_engine_matches_oracle(shape, rule):
    """Finite upper bounds on the fused engine: t2 candidates, flips of the entering variable (iterations
    without a pivot) and substitution of a leaving variable at its bound, against the LU oracle (pinned on the
    reference's bounded fixtures): same iteration count, pivot trace, flip flags, final lists; x_B, b and the
    objective within 1e-9; and bit for bit the staged and the 4-phase engines, which rewrite A in place."""
    m, n, tight = shape
    A, b, c, ub, basic, nonbasic = _bounded_synth(m, n, 31 + m, tight)
    want = oracle.simplex(A, b, c, basic, nonbasic, ub=ub, rule=oracle.RULES[rule], variant="lu")
    assert want["term"] == "optimal" and want["flips"] > 0
    runs = {}
    for engine, dual in (("streaming", "update"), ("persistent", "update"), ("persistent", "recompute"), ("staged", "update")):
        for chunk in ((64, 5) if engine == "streaming" else (64,)):   # a launch every 5 iterations: epilogue + re-entry
            got = _run(A, b, c, basic, nonbasic, ub=ub, pricing=rule, engine=engine, dual=dual, chunk_iters=chunk)
            assert got["engine_used"] == ("staged" if engine == "staged" else "persistent" if dual == "recompute" else
                                          "fused" if engine == "streaming" else "resident")
            assert got["term"] == "optimal"
            assert got["iterations"] == want["iterations"] and got["flips"] == want["flips"], (engine, dual, chunk)
            assert got["trace"].tolist() == want["trace"].tolist(), (engine, dual, chunk)
            assert got["flip"].tolist() == want["flip"].tolist()
            assert got["basic"].tolist() == want["basic"].tolist() and got["nonbasic"].tolist() == want["nonbasic"].tolist()
            assert rel_close(got["x"], want["x"]) and rel_close(got["objective"], want["objective"])
            runs[(engine, dual, chunk)] = got
    base = runs[("staged", "update", 64)]
    for key, got in runs.items():
        assert rel_close(got["x"], base["x"], tol=1e-12), key
    # one launch or a launch every 5 iterations (columns rewritten at every exit, flags re-read at every entry)
    assert np.array_equal(runs[("streaming", "update", 64)]["x"], runs[("streaming", "update", 5)]["x"])


def test_bounded_fused_engine_leaves_the_tableau_as_the_reference_does():
    """After a bounded run the device holds what the reference's loop leaves in matrix_A / vector_b /
    vector_c (flipped columns negated, b shifted, src/LinearProgram.cpp:154-169): b is compared with the oracle's,
    and a second run from the final basis on the same context (A re-read, unit-column table still valid) is
    optimal at once."""
    from simplex_b200 import Basis
    A, b, c, ub, basic, nonbasic = _bounded_synth(96, 200, 5, 0.6)
    want = oracle.simplex(A, b, c, basic, nonbasic, ub=ub, rule=oracle.RULE_MOST_NEG, variant="lu")
    with _device_solver(pricing="most_neg", engine="streaming", dual="update", chunk_iters=7) as s:
        s.load(A, b, c, ub)
        basis = Basis(basic, nonbasic)
        r = s.simplex(basis)
        assert s.engine_used() == "fused" and r["flips"] == want["flips"] > 0
        assert rel_close(s.b(), want["b"])
        x1 = s.xB()
        r2 = s.simplex(basis)          # the flags restart at zero (Simplex.cpp:271), A stays as it was left
        assert (r2["term"], r2["iterations"], r2["pivots"], r2["flips"]) == ("optimal", 1, 0, 0)
        assert rel_close(s.xB(), x1)


def test_bounded_headline_shape_invariants():
    """m=1024, n=2048 with finite bounds, 400 iterations: the fused bounded kernel and the 4-phase kernel walk
    the same pivot sequence with the same flips; B^-1 A_B = I on the REWRITTEN columns; lists stay a permutation."""
    from simplex_b200 import Basis, DeviceSimplex
    A, b, c, ub, basic, nonbasic = _bounded_synth(1024, 2048, 1234, 0.4)
    out = []
    for engine, dual in (("streaming", "update"), ("persistent", "recompute")):
        with DeviceSimplex(pricing="most_neg", engine=engine, dual=dual, iter_limit=400, trace_capacity=400,
                           chunk_iters=64) as s:
            s.load(A, b, c, ub)
            basis = Basis(basic, nonbasic)
            r = s.simplex(basis)
            tr, _ = s.trace()
            out.append((r["pivots"], r["flips"], tr.tolist(), s.flip_flags().tolist(), basis.basic.tolist()))
            assert r["term"] == "iter_limit" and r["flips"] > 0 and r["pivots"] + r["flips"] == 400
            assert s.inverse_residual(16, 3) < 1e-9
            assert s.primal_residual() < 1e-9
            assert sorted(basis.basic.tolist() + basis.non_basic.tolist()) == list(range(A.shape[1]))
    assert out[0] == out[1]


def test_hygiene_refresh_does_not_change_the_pivot_sequence(golden_synth):
    """sb200_opts.hygiene_*: with a tolerance nothing can meet, B^-1 and x_B are rebuilt from A_B (Gauss-Jordan)
    behind every launch and twice more before the ending is accepted; the run still walks the reference's pivot
    sequence, ends after the reference's iteration count (the re-examined last iteration is not counted twice)
    and on its optimum. With the default tolerance the checks run and trigger nothing."""
    from simplex_b200 import Basis, synth
    n_checked = 0
    for rec in golden_synth:
        if rec["n_eq"] != 0 or rec["m"] > 128:
            continue
        A, b, c, basic, nonbasic = synth.tableau(rec["m"], rec["n"], 0, rec["seed"])
        call = rec["calls"][-1]
        for engine in ("persistent", "streaming", "staged"):
            for tol, want_refresh in ((1e-300, True), (None, False)):
                with _device_solver(pricing=rec["rule"], engine=engine, dual="update", chunk_iters=16, trace_capacity=1 << 14,
                                    hygiene_tol=tol, hygiene_period=1) as s:
                    s.load(A, b, c)
                    basis = Basis(basic, nonbasic)
                    r = s.simplex(basis)
                    tr, _ = s.trace()
                    st = s.hygiene_stats()
                assert r["term"] == "optimal" and r["iterations"] == call["iterations"], (engine, tol)
                assert tr.tolist() == call["pivots"] and basis.basic.tolist() == call["basic_out"], (engine, tol)
                assert rel_close(r["objective"], rec["objective"])
                assert st["checks"] >= 1 and (st["refreshes"] >= 2) == want_refresh, (engine, tol, st)
                if not want_refresh:
                    assert st["refreshes"] == 0 and st["worst"] < 1e-10
        n_checked += 1
    assert n_checked >= 4


def test_sharded_and_plain_state_do_not_mix():
    """A context that holds a column shard refuses the single-GPU entry points (its buffers are shard-sized),
    and a plain load after it starts from fresh allocations."""
    from simplex_b200 import Basis, synth
    from simplex_b200._abi import Sb200Error
    from simplex_b200.dist import ShardedSimplex
    A, b, c, basic, nonbasic = synth.tableau(64, 128, 0, 3)
    with ShardedSimplex(0, 1, pricing="most_neg", device=0, chunk_iters=32) as sh:
        sh.load_shard(64, A.shape[1], A, b, c)
        L, ctx = sh._L, sh._ctx
        import ctypes as C
        bas = np.ascontiguousarray(basic, dtype=np.int64)
        non = np.ascontiguousarray(nonbasic, dtype=np.int64)
        lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        assert L.sb200_prepare(ctx, lp(bas), bas.size, lp(non), non.size) == 4      # SB200_ERR_STATE
        assert L.sb200_set_costs(ctx, c.size, c.ctypes.data_as(C.POINTER(C.c_double))) == 4
        r = sh.simplex(Basis(basic, nonbasic))
        assert r["term"] == "optimal"
        out = np.zeros(64 * 64)
        assert L.sb200_get_Binv(ctx, out.ctypes.data_as(C.POINTER(C.c_double))) == 4
        # the same context, now as a plain single-GPU one: same shape, nothing of the shard is reused
        Af = np.asfortranarray(A)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        assert L.sb200_load(ctx, 64, A.shape[1], dp(Af), 64, dp(b), dp(c), None) == 0
        from simplex_b200 import _abi
        res = _abi.Result()
        assert L.sb200_run(ctx, lp(bas), bas.size, lp(non), non.size, C.byref(res)) == 0
        assert res.term == 1 and res.iterations == r["iterations"]


@pytest.mark.parametrize("engine,dual", ENGINE_DUAL)
def test_repeated_runs_are_bit_identical(engine, dual):
    """Ten runs of the same LP on one context and on fresh contexts: same trace and the same bits in x_B, y and
    B^-1 every time. The engines exchange candidates and vectors through self-validating words and late list
    swaps without fences; a race there would show up as a run that differs from the others."""
    from simplex_b200 import Basis, synth
    A, b, c, basic, nonbasic = synth.tableau(200, 420, 0, 21)
    first = None
    for fresh in (False, True):
        solver = None
        for rep in range(5):
            if fresh or solver is None:
                if solver is not None:
                    solver.close()
                solver = _device_solver(pricing="most_neg", engine=engine, dual=dual, chunk_iters=13, trace_capacity=4096)
                solver.load(A, b, c)
            basis = Basis(basic, nonbasic)
            r = solver.simplex(basis)
            tr, _ = solver.trace()
            got = (r["iterations"], tr.tobytes(), solver.xB().tobytes(), solver.y().tobytes(), solver.Binv().tobytes())
            if first is None:
                first = got
            assert got == first, (fresh, rep)
        solver.close()


def test_sharded_exchange_timeout_returns_instead_of_hanging(monkeypatch):
    """A rank whose peer never shows up gives up after the exchange timeout and returns an error: every spin of
    the kernel (grid barrier, mail poll, vector decode) honours the abort word, so no CTA is left waiting for
    one that has gone (round-1 advice: the timeout path was not grid-uniform)."""
    import ctypes as C
    import time
    from simplex_b200 import Basis, _abi, synth
    from simplex_b200.dist import LocalShardedSimplex
    from simplex_b200.solver import _dp, _lp
    monkeypatch.setenv("SB200_XCHG_TIMEOUT_MS", "300")
    A, b, c, basic, nonbasic = synth.tableau(64, 128, 0, 9)
    with LocalShardedSimplex(2, pricing="most_neg", chunk_iters=16) as s:
        s.load(A, b, c)
        basis = Basis(basic, nonbasic)
        diag = np.zeros(64)
        for ctx in s._ctxs:
            part = np.zeros(64)
            assert s._L.sb200_shard_prepare(ctx, _lp(basis.basic), 64, _lp(basis.non_basic), basis.non_basic.size, _dp(part)) == 0
            diag += part
        for ctx in s._ctxs:
            assert s._L.sb200_shard_finish_prepare(ctx, _dp(diag)) == 0
        res = _abi.Result()
        t0 = time.time()
        rc = s._L.sb200_shard_run(s._ctxs[0], _lp(basis.basic), _lp(basis.non_basic), C.byref(res))   # rank 1 never runs
        dt = time.time() - t0
        assert rc != 0 and b"did not answer" in s._L.sb200_last_error(s._ctxs[0]), (rc, s._L.sb200_last_error(s._ctxs[0]))
        assert dt < 5.0, dt
    # and the pair works when both ranks run
    with LocalShardedSimplex(2, pricing="most_neg", chunk_iters=16) as s:
        s.load(A, b, c)
        r = s.simplex(Basis(basic, nonbasic))
        assert r["term"] == "optimal"


def test_fuzz_small_shapes_against_oracle():
    """Seeded random shapes (3 <= m <= 140, ragged leading dimensions, fewer rows than CTAs, few or many columns),
    with and without equality rows (Phase-I tableaux: artificial unit columns) and with and without finite bounds
    (also on slack columns, so that unit columns get substituted and the unit-column table has to follow), both
    pricing rules, on the fused kernel with a launch every 11 iterations: every non-degenerate run walks the LU
    oracle's pivot sequence to the same flags, lists, x_B and objective."""
    from helpers import DBL_MAX
    from simplex_b200 import synth
    rng = np.random.default_rng(2025)
    n_runs = n_bounded = n_unit_flips = 0
    for case in range(60):
        m = int(rng.integers(3, 141))
        n = int(rng.integers(2, 221))
        n_eq = int(rng.integers(0, max(1, m // 3))) if case % 3 == 0 else 0
        seed = 1000 + case
        A, b, c, basic, nonbasic = synth.tableau(m, n, n_eq, seed, phase1=(n_eq > 0))
        N = A.shape[1]
        ub = None
        if case % 2 == 0:
            ub = np.full(N, DBL_MAX)
            ub[:n] = np.where(rng.random(n) < 0.5, rng.uniform(0.01, 0.4, n), rng.uniform(1.0, 40.0, n))
            ns = min(N - n, m)
            pick = rng.random(ns) < 0.5
            ub[n:n + ns] = np.where(pick, np.maximum(b[:ns], 0.05) * rng.uniform(1.2, 3.0, ns), DBL_MAX)
        rule = "most_neg" if case % 4 < 2 else "first"
        want = oracle.simplex(A, b, c, basic, nonbasic, ub=ub, rule=oracle.RULES[rule], variant="lu")
        got = _run(A, b, c, basic, nonbasic, ub=ub, pricing=rule, engine="streaming", dual="update", chunk_iters=11)
        assert got["engine_used"] == "fused"
        assert got["term"] == want["term"], (case, m, n, n_eq, rule)
        if want["x"].min() < 1e-9 and want["term"] == "optimal" and n_eq > 0:
            continue      # degenerate Phase-I vertex: ties are decided by rounding noise
        assert got["iterations"] == want["iterations"] and got["flips"] == want["flips"], (case, m, n, n_eq, rule)
        assert got["trace"].tolist() == want["trace"].tolist(), (case, m, n, n_eq, rule)
        assert got["flip"].tolist() == want["flip"].tolist(), case
        assert got["basic"].tolist() == want["basic"].tolist() and got["nonbasic"].tolist() == want["nonbasic"].tolist(), case
        if want["term"] == "optimal":
            assert rel_close(got["x"], want["x"]) and rel_close(got["objective"], want["objective"]), case
        n_runs += 1
        n_bounded += ub is not None
        n_unit_flips += int(want["flip"][n:].sum()) if ub is not None else 0
    assert n_runs >= 40 and n_bounded >= 20, (n_runs, n_bounded)
    print("fuzz: %d runs, %d bounded, %d substituted unit columns at the end" % (n_runs, n_bounded, n_unit_flips))


@pytest.mark.parametrize("rule", ["most_neg", "first"])
def test_bounded_unit_columns_are_substituted_like_dense_ones(rule):
    """Structural columns with a single nonzero, attractive costs and small upper bounds beside a dense LP: they
    enter, hit their own bounds (bound flips of UNIT columns, which the fused kernel prices from its unit-column
    table as -(c_j - v y_r) while the flag is up) and the table entry is rewritten at the exit of every launch (a
    launch every 3 iterations). Pivot sequence, flips, flags and x_B against the LU oracle; the flags of unit columns
    must be among the ones set."""
    from helpers import DBL_MAX
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
        got = _run(A, b, c, basic, nonbasic, ub=ub, pricing=rule, engine="streaming", dual="update", chunk_iters=chunk)
        assert got["engine_used"] == "fused"
        assert (got["term"], got["iterations"], got["flips"]) == (want["term"], want["iterations"], want["flips"]), chunk
        assert got["trace"].tolist() == want["trace"].tolist(), chunk
        assert got["flip"].tolist() == want["flip"].tolist(), chunk
        assert got["basic"].tolist() == want["basic"].tolist() and got["nonbasic"].tolist() == want["nonbasic"].tolist()
        assert rel_close(got["x"], want["x"]) and rel_close(got["objective"], want["objective"]), chunk


@pytest.mark.parametrize("rule", ["most_neg", "first"])
@pytest.mark.parametrize("m,n", [(16, 32), (40, 72), (64, 128)])
def test_batched_bounded_matches_oracle(m, n, rule):
    """Batched mode with finite upper bounds (sb200_run_batched_bounded): 24 independent LPs, one CTA each, the
    bounded ratio test and both kinds of substitutions on the shared-memory tableau, against the LU oracle LP by LP:
    iteration counts, flip counts and flags, final lists, x_B and objective. Host path (chunked pipeline, pageable
    and pinned inputs give the same results)."""
    from helpers import DBL_MAX
    from simplex_b200 import DeviceSimplex, synth
    batch = 24
    A, b, c, basic, nonbasic = synth.batch(batch, m, n, 4321)
    N = m + n
    rng = np.random.default_rng(99)
    ub = np.full((batch, N), DBL_MAX)
    ub[:, :n] = np.where(rng.random((batch, n)) < 0.5, rng.uniform(0.01, 0.4, (batch, n)), rng.uniform(1.0, 40.0, (batch, n)))
    orule = oracle.RULES[rule]
    want = []
    for k in range(batch):
        Ak = np.asfortranarray(A[k].T)
        want.append(oracle.simplex(Ak, b[k], c[k], basic[k].astype(np.int64), nonbasic[k].astype(np.int64), ub=ub[k],
                                   rule=orule, variant="lu"))
    assert sum(w["flips"] for w in want) > 0
    bas, non = basic.copy(), nonbasic.copy()
    with DeviceSimplex(pricing=rule) as s:
        r = s.run_batched(A, b, c, bas, non, ub=ub)
    for k, w in enumerate(want):
        assert r["term"][k] == 1 and w["term"] == "optimal", k
        assert r["iterations"][k] == w["iterations"] and r["flips"][k] == w["flips"], (k, r["iterations"][k], w["iterations"])
        assert r["flip"][k].tolist() == w["flip"].tolist(), k
        assert bas[k].tolist() == w["basic"].tolist() and non[k].tolist() == w["nonbasic"].tolist(), k
        assert rel_close(r["xB"][k], w["x"]) and rel_close(r["objective"][k], w["objective"]), k
    # the unbounded entry point is untouched by the bounded code path
    bas2, non2 = basic.copy(), nonbasic.copy()
    with DeviceSimplex(pricing=rule) as s:
        r2 = s.run_batched(A, b, c, bas2, non2)
    assert (r2["term"] == 1).all()
