"""The oracle (oracle/simplex_oracle.c) pinned against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py). CPU only."""
import numpy as np
import pytest

from helpers import DEGENERATE_CALLS, call_inputs, has_bounds, rel_close
from oracle import oracle


def _check_call(name, k, call, res, exact=True):
    want_term = "unbounded" if call["unbounded"] else "optimal"
    assert res["term"] == want_term, (name, k)
    if exact:
        assert res["iterations"] == call["iterations"], (name, k)
        assert res["trace"].tolist() == call["pivots"], (name, k)
        assert res["basic"].tolist() == call["basic_out"], (name, k)
        assert res["nonbasic"].tolist() == call["nonbasic_out"], (name, k)


@pytest.mark.parametrize("variant", ["lu", "pfi"])
def test_fixture_calls_match_reference(golden_fixtures, variant):
    """Every recorded Simplex::simplex() call of the 15 fixture runs: same status, iteration
    count, pivot trace and final lists (stock first-eligible rule)."""
    n_calls = 0
    for name, rec in golden_fixtures.items():
        bounded = has_bounds(rec)
        for k, call in enumerate(rec["calls"]):
            A, b, c, ub, basic, nonbasic = call_inputs(call)
            res = oracle.simplex(A, b, c, basic, nonbasic, ub=ub if bounded else None, rule=oracle.RULE_FIRST,
                                 variant=variant)
            exact = not (variant == "pfi" and (name, k) in DEGENERATE_CALLS)
            _check_call(name, k, call, res, exact=exact)
            n_calls += 1
    assert n_calls >= 20


def test_fixture_objectives(golden_fixtures):
    """Known optima written in the reference's fixture files."""
    want = {"sched": 187, "sched_1": 260, "sched_2": 137, "test": 2, "test_1": -20, "test_2": 4.5, "test_bnds_1": 36,
            "klee-minty": 10000, "klee-minty-n10": 1048576, "gen_kleeminty_n10_rhs5": 9765625}
    for name, v in want.items():
        assert abs(golden_fixtures[name]["objective"] - v) < 1e-6, name


def test_synth_generator_matches_reference(golden_synth):
    """oracle_synth_tableau (mt19937_64 restated + tableau + scaling) is bit-identical to the
    tableau the reference builds through its own model API for the same G(m,n,n_eq,seed)."""
    seen = 0
    for rec in golden_synth:
        for k, call in enumerate(rec["calls"]):
            if "A" not in call:
                continue
            phase1 = (rec["n_eq"] > 0 and k == 0)
            A, b, c, basic, nonbasic = oracle.synth_tableau(rec["m"], rec["n"], rec["n_eq"], rec["seed"], phase1)
            Ar, br, cr, _, bas_r, non_r = call_inputs(call)
            assert A.shape == Ar.shape
            assert np.array_equal(A, Ar) and np.array_equal(b, br) and np.array_equal(c, cr)
            if basic is not None:
                assert basic.tolist() == bas_r.tolist() and nonbasic.tolist() == non_r.tolist()
            seen += 1
    assert seen >= 6


def _run_two_phase(rec, variant):
    m, n, n_eq, seed = rec["m"], rec["n"], rec["n_eq"], rec["seed"]
    rule = oracle.RULES[rec["rule"]]
    out = []
    if n_eq > 0:
        A, b, c, basic, nonbasic = oracle.synth_tableau(m, n, n_eq, seed, True)
        r1 = oracle.simplex(A, b, c, basic, nonbasic, rule=rule, variant=variant)
        out.append(r1)
        basic, nonbasic = oracle.phase2_lists_from_phase1(r1["basic"], r1["nonbasic"], n + (m - n_eq))
        A, b, c, _, _ = oracle.synth_tableau(m, n, n_eq, seed, False)
    else:
        A, b, c, basic, nonbasic = oracle.synth_tableau(m, n, n_eq, seed, False)
    out.append(oracle.simplex(A, b, c, basic, nonbasic, rule=rule, variant=variant))
    return out


@pytest.mark.parametrize("variant,max_m", [("lu", 128), ("pfi", 512)])
def test_synth_traces_match_reference(golden_synth, variant, max_m):
    """Synthetic table (SURVEY.md App. A.3 and more): pivot traces, iteration counts, final bases
    identical to the reference; objective and structural x within 1e-9 relative."""
    n = 0
    for rec in golden_synth:
        if rec["m"] > max_m:
            continue
        res = _run_two_phase(rec, variant)
        assert len(res) == len(rec["calls"])
        for k, (r, call) in enumerate(zip(res, rec["calls"])):
            _check_call((rec["m"], rec["n"], rec["n_eq"], rec["seed"], rec["rule"]), k, call, r)
        last = res[-1]
        assert rel_close(last["objective"], rec["objective"])
        x = np.zeros(rec["n"])
        for i, col in enumerate(last["basic"]):
            if col < rec["n"]:
                x[col] = last["x"][i]
        assert rel_close(x, rec["x_struct"])
        n += 1
    assert n >= 8


def test_rescale_and_flip_primitives():
    rng = np.random.default_rng(0)
    m, N = 5, 9
    A = np.asfortranarray(rng.uniform(-3, 3, (m, N)))
    b = rng.uniform(0, 5, m)
    c = rng.uniform(-1, 1, N)
    A0, b0, c0 = A.copy(order="F"), b.copy(), c.copy()
    oracle.lib().oracle_rescale(m, N, oracle._dp(A), oracle._dp(b), oracle._dp(c))
    rmax = np.abs(A0).max(axis=1)
    scale = np.where(rmax > 1.0, rmax, 1.0)      # LinearProgram.cpp:179: rows only when max > 1
    A1 = A0 / scale[:, None]
    cmax = np.abs(A1).max(axis=0)                # LinearProgram.cpp:186-194: every column
    assert np.allclose(A, A1 / cmax[None, :], rtol=0, atol=0)
    assert np.array_equal(b, b0 / scale) and np.array_equal(c, c0 / cmax)


# ------------------------------------------------------------------ BASELINE sizes (tests/golden/large.json.gz)
def _replay(m, n, pivots):
    basic = [n + i for i in range(m)]
    nonbasic = list(range(n))
    for lc, row, ec, pos in pivots:
        assert basic[row] == lc and nonbasic[pos] == ec
        basic[row], nonbasic[pos] = ec, lc
    return basic, nonbasic


def test_large_golden_is_consistent(golden_large):
    """Every recorded trace replays (Basis.cpp:48-53) from its start lists onto its recorded end lists."""
    assert {"config1", "headline_prefix", "sharded_check"} <= set(golden_large)
    for name, g in golden_large.items():
        if g["mode"] == "synth":
            for call in g["calls"]:
                basic, nonbasic = list(call["basic_in"]), list(call["nonbasic_in"])
                for lc, row, ec, pos in call["pivots"]:
                    assert basic[row] == lc and nonbasic[pos] == ec, name
                    basic[row], nonbasic[pos] = ec, lc
                assert basic == call["basic_out"] and nonbasic == call["nonbasic_out"], name
        else:
            basic, _ = _replay(g["m"], g["n"], g["pivots"])
            assert basic == g["basic_after"], name


def test_config1_full_run_matches_reference(golden_large):
    """BASELINE configs[1], G(1024, 2048, 256, 1234), two-phase, most-negative pricing: the restatement's
    explicit-inverse variant walks the reference's 727 + 6059 pivots and ends on the reference's optimum
    (reference: Simplex::solve through the model API, /root/reference/src/Simplex.cpp:22-45, 255-353)."""
    g = golden_large["config1"]
    res = _run_two_phase(g, "pfi")
    assert g["status"] == "optimal" and len(res) == len(g["calls"]) == 2
    for k, (r, call) in enumerate(zip(res, g["calls"])):
        _check_call("config1", k, call, r)
    assert rel_close(res[-1]["objective"], g["objective"])
    x = np.zeros(g["n"])
    for i, col in enumerate(res[-1]["basic"]):
        if col < g["n"]:
            x[col] = res[-1]["x"][i]
    assert rel_close(x, g["x_struct"])


@pytest.mark.parametrize("name,pivots", [("headline_prefix", 96), ("sharded_check", 600), ("mid_prefix", 300),
                                          ("first_eligible", 600)])
def test_large_prefixes_match_reference(golden_large, name, pivots):
    """The first pivots of the reference at the headline shape (m=4096, n=16384), the wide shape of the
    multi-rank checks (m=1024, n=8192), a mid-size LP and the stock first-eligible rule at m=1024: the
    restatement (explicit inverse) takes the same decisions. (The GPU tests compare the whole recorded
    prefixes; here the length is bounded by what the CPU suite can afford.)"""
    if name not in golden_large:
        pytest.skip("%s not generated yet" % name)
    g = golden_large[name]
    P = min(pivots, len(g["pivots"]))
    A, b, c, basic, nonbasic = oracle.synth_tableau(g["m"], g["n"], 0, g["seed"], False)
    r = oracle.simplex(A, b, c, basic, nonbasic, rule=oracle.RULES[g["rule"]], variant="pfi", iter_limit=P)
    assert r["trace"].tolist() == g["pivots"][:P], name


def _reference_driver(rule):
    import os
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                     "ref_driver_" + ("dantzig" if rule == "dantzig" else "stock"))
    return p if os.path.exists(p) else None


@pytest.mark.parametrize("rule", ["dantzig", "stock"])
def test_oracle_against_reference_binaries_live(rule):
    """Beyond the committed golden files: the reference itself (oracle/_ref/ref_driver_*, its own sources
    compiled where they lie, see oracle/Makefile) run NOW on 20 random two-phase synthetic LPs per rule, shapes
    and seeds the golden files do not hold, against the C oracle on the same generator arguments: same status,
    iteration counts and complete pivot traces of every Simplex::simplex() call, objective within 1e-9.
    Runs the prebuilt binary only (nothing under /root/reference is read at run time)."""
    import json
    import subprocess
    import tempfile
    exe = _reference_driver(rule)
    if exe is None:
        pytest.skip("oracle/_ref binaries not built (make -C oracle ref needs the reference tree)")
    rng = np.random.default_rng(20261020 if rule == "dantzig" else 20261021)
    checked = 0
    for _ in range(20):
        m = int(rng.integers(6, 80))
        n = int(rng.integers(m // 2 + 2, 3 * m))
        n_eq = int(rng.integers(0, max(1, m // 3)))
        seed = int(rng.integers(1, 1 << 30))
        with tempfile.TemporaryDirectory() as td:
            out = td + "/o.json"
            subprocess.run([exe, "synth", str(m), str(n), str(n_eq), str(seed), "--out", out], cwd=td,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False, timeout=120)
            with open(out) as f:
                ref = json.load(f)
        rec = {"m": m, "n": n, "n_eq": n_eq, "seed": seed, "rule": rule}
        if ref["status"] != "optimal" or len(ref["calls"]) != (2 if n_eq > 0 else 1):
            continue  # (infeasible / stopped in Phase I: the two-phase replay below does not apply)
        for variant in ("lu", "pfi"):  # the reference's own arithmetic, and the product form the GPU keeps
            res = _run_two_phase(rec, variant)
            assert len(res) == len(ref["calls"]), rec
            for k, (r, call) in enumerate(zip(res, ref["calls"])):
                assert r["term"] == ("unbounded" if call["unbounded"] else "optimal"), (rec, variant, k)
                assert r["iterations"] == call["iterations"], (rec, variant, k)
                assert r["trace"].tolist() == call["pivots"], (rec, variant, k)
                assert r["basic"].tolist() == call["last_basis"], (rec, variant, k)
            assert rel_close(res[-1]["objective"], ref["objective"]), (rec, variant)
        checked += 1
    assert checked >= 12, checked


@pytest.mark.parametrize("rule", ["dantzig", "stock"])
def test_oracle_against_reference_binaries_live_bounded(rule):
    """The same live check for LPs WITH finite upper bounds (bounded ratio test, both kinds of bound flips,
    substituted leaving variables: reference Simplex.cpp:116-179): 12 random bounded LP texts per rule, some
    with equality rows and shifted lower bounds, through the reference's own parser, presolve and
    Simplex::solve(); the oracle replays every Simplex::simplex() call from the tableau the reference dumped."""
    import json
    import os
    import subprocess
    import sys
    import tempfile
    exe = _reference_driver(rule)
    if exe is None:
        pytest.skip("oracle/_ref binaries not built (make -C oracle ref needs the reference tree)")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_bounded_golden import bounded_lp_text
    from make_golden import complete_lists
    rng = np.random.default_rng(777 if rule == "dantzig" else 778)
    calls = flips = 0
    for _ in range(12):
        m = int(rng.integers(8, 48))
        n = int(rng.integers(m, 2 * m + 8))
        n_eq = int(rng.integers(0, 2)) * int(rng.integers(1, max(2, m // 4)))
        text = bounded_lp_text(int(rng.integers(1, 1 << 30)), m, n, n_eq, tight=float(rng.uniform(0.2, 0.7)),
                               lower=float(rng.choice([0.0, 0.15])))
        with tempfile.TemporaryDirectory() as td:
            with open(td + "/p.lp", "w") as f:
                f.write(text)
            subprocess.run([exe, "lp", td + "/p.lp", "--out", td + "/o.json", "--dump-tableau"], cwd=td,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False, timeout=120)
            with open(td + "/o.json") as f:
                ref = json.load(f)
        for k, call in enumerate(ref["calls"]):
            complete_lists(call)
            A, b, c, ub, basic, nonbasic = call_inputs(call)
            res = oracle.simplex(A, b, c, basic, nonbasic, ub=ub, rule=oracle.RULES[rule], variant="lu")
            _check_call((m, n, n_eq, rule), k, call, res)
            flips += call["iterations"] - len(call["pivots"]) - 1  # iterations that only flipped the entering variable
            calls += 1
    assert calls >= 12 and flips >= 5, (calls, flips)
