"""The reference-facing flow on the GPU: .lp text -> reader -> presolve -> Phase I / Phase II
(device loop through the C ABI) -> solution read-out -> sol_test.xml, through the host library
and through the `simplex <file.lp>` command line, against what the unmodified reference
produced for the same inputs (tests/golden/*.json.gz). Run with -m gpu on the B200 box."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import DEGENERATE_CALLS, rel_close
from simplex_b200 import lp as hostlp

pytestmark = pytest.mark.gpu


def parse_xml(text):
    head = re.search(r'<header objectiveValue="([^"]*)"/>', text)
    rows = re.findall(r'<variable name="([^"]*)" value="([^"]*)"/>', text)
    return float(head.group(1)), [(n, float(v)) for n, v in rows]


@pytest.mark.parametrize("engine,dual", [("persistent", "update"), ("persistent", "recompute"), ("staged", "recompute")])
def test_fixtures_through_host_library(golden_fixtures, engine, dual):
    """Simplex::solve() on every fixture: status, per-phase iteration and pivot counts, objective
    and every variable within 1e-9 relative of the reference's full-precision read-out."""
    for name, rec in golden_fixtures.items():
        with hostlp.LpSession(engine=engine, dual=dual) as s:
            s.read(rec["lp_text"])
            s.solve()
            calls = s.calls()
            assert len(calls) == len(rec["calls"]), name
            for k, (got, want) in enumerate(zip(calls, rec["calls"])):
                assert (got["M"], got["N"]) == (want["M"], want["N"]), (name, k)
                assert got["term"] == (2 if want["unbounded"] else 1), (name, k)
                if (name, k) not in DEGENERATE_CALLS and not any((name, j) in DEGENERATE_CALLS for j in range(k)):
                    assert got["iterations"] == want["iterations"], (name, k, got, want["iterations"])
                    assert got["pivots"] == len(want["pivots"]), (name, k)
            assert s.status() == rec["status"], name
            if rec["status"] == "optimal":
                assert rel_close(s.objective(), rec["objective"]), (name, s.objective(), rec["objective"])
                sol = s.solution()
                assert sorted(sol) == sorted(rec["solution"]), name
                if not any((name, j) in DEGENERATE_CALLS for j in range(len(calls))):
                    for k, v in rec["solution"].items():
                        assert rel_close(sol[k], v), (name, k, sol[k], v)


def test_phase_two_reuses_the_resident_tableau(golden_fixtures):
    """Two-phase LPs keep A and b in HBM across the phases (only c is replaced) unless a bound
    flip of Phase I rewrote them on the device."""
    reused = 0
    for name, rec in golden_fixtures.items():
        if len(rec["calls"]) != 2:
            continue
        with hostlp.LpSession() as s:
            s.read(rec["lp_text"])
            s.solve()
            calls = s.calls()
            assert calls[0]["resident"] == 0
            assert calls[1]["resident"] == (1 if calls[0]["bound_flips"] == 0 else 0), name
            reused += calls[1]["resident"]
            assert rel_close(s.objective(), rec["objective"]), name
    assert reused >= 2


def test_command_line_matches_reference(golden_cli, tmp_path):
    """`simplex file.lp`: same exit status, same plain stdout lines, sol_test.xml present/absent as
    for the reference, same variables in the same order, values equal to the printed 6 decimals."""
    assert os.path.exists(hostlp.CLI_PATH), "simplex_b200/simplex not built"
    env = dict(os.environ, SIMPLEX_LOG="off")
    exact = 0
    for name, rec in golden_cli.items():
        wd = tmp_path / re.sub(r"\W", "_", name)
        wd.mkdir()
        (wd / "in.lp").write_text(rec["lp_text"])
        p = subprocess.run([hostlp.CLI_PATH, "in.lp"], cwd=wd, capture_output=True, text=True, env=env, timeout=300)
        xml_file = wd / "sol_test.xml"
        if rec.get("degenerate"):
            # tied ratios: the updated inverse may legitimately walk another path;
            # only an optimum reached by both is compared, and only its objective
            if rec["xml"] is not None and xml_file.exists() and p.returncode == 0 and not rec["unbounded"]:
                got_obj, _ = parse_xml(xml_file.read_text())
                want_obj, _ = parse_xml(rec["xml"])
                # (quirk App. B 11 can turn DBL_MAX bounds into 1e307-sized "optima": not comparable)
                if "LP IS UNBOUNDED" not in p.stdout and abs(want_obj) < 1e100:
                    assert abs(got_obj - want_obj) <= 1.5e-6 * max(1.0, abs(want_obj)), (name, got_obj, want_obj)
            continue
        assert p.returncode == rec["returncode"], (name, p.returncode, p.stdout[-400:], p.stderr[-400:])
        assert xml_file.exists() == (rec["xml"] is not None), name
        if rec["returncode"] == 0:
            plain = [ln for ln in p.stdout.splitlines() if ln and not ln.startswith("[")]
            assert plain == rec["stdout_plain"], (name, plain)
        if rec["xml"] is None:
            continue
        got_obj, got_rows = parse_xml(xml_file.read_text())
        want_obj, want_rows = parse_xml(rec["xml"])
        assert [n for n, _ in got_rows] == [n for n, _ in want_rows], name
        degenerate = name.startswith("fixture:") and any(d[0] == name[8:] for d in DEGENERATE_CALLS)
        if np.isfinite(want_obj):
            assert abs(got_obj - want_obj) <= 1.5e-6 * max(1.0, abs(want_obj)), (name, got_obj, want_obj)
        if not degenerate:
            for (n, g), (_, w) in zip(got_rows, want_rows):
                if np.isfinite(w):
                    assert abs(g - w) <= 1.5e-6 * max(1.0, abs(w)), (name, n, g, w)
        exact += xml_file.read_text() == rec["xml"]
    assert exact >= len([r for r in golden_cli.values() if r["xml"] is not None]) - 4


def test_dialect_and_random_texts_through_host_library(golden_parser_cases):
    """Every hand-written and seeded random .lp text the reference solved (bounded variables,
    equalities, '>=' rows: Phase I, bounded ratio test, flip reversal): same status, per-call
    iteration and pivot counts, objective and solution within 1e-9 — unless the case is flagged
    degenerate (tied ratios), where only a common optimum's objective is compared."""
    n_strict = n_flips = 0
    for name, rec in golden_parser_cases.items():
        if not rec["calls"]:
            continue
        with hostlp.LpSession() as s:
            s.read(rec["lp_text"])
            rc = 0
            try:
                s.solve()
            except hostlp.LpError as e:
                rc = e.rc
            calls = s.calls()
            if rec["degenerate"]:
                if rec["status"] == "optimal" and rc == 0 and s.status() == "optimal" and abs(rec["objective"]) < 1e100:
                    assert rel_close(s.objective(), rec["objective"], 1e-7), name
                continue
            if rec["status"] == "exception_cstr":
                assert rc == hostlp.READ_FATAL, (name, rc)      # artificial variable left in the basis
                continue
            assert rc == 0, (name, s.error())
            assert s.status() == rec["status"], (name, s.status())
            assert [(c["M"], c["N"], c["iterations"], c["pivots"]) for c in calls] == \
                   [(c["M"], c["N"], c["iterations"], c["pivots"]) for c in rec["call_stats"]], name
            n_flips += sum(c["bound_flips"] for c in calls)
            if rec["status"] == "optimal":
                want = rec["objective"]
                if np.isfinite(want):
                    assert rel_close(s.objective(), want), (name, s.objective(), want)
                else:   # a zero column scaled by 0/0 (reference src/LinearProgram.cpp:192): NaN in, NaN out
                    assert np.isnan(s.objective()) == np.isnan(want), (name, s.objective(), want)
                sol = s.solution()
                assert sorted(sol) == sorted(rec["solution"]), name
                for k, v in rec["solution"].items():
                    if np.isfinite(v):
                        assert rel_close(sol[k], v), (name, k, sol[k], v)
            n_strict += 1
    assert n_strict >= 35
    assert n_flips >= 10      # the bounded path (K4-SUB) is really exercised


def test_command_line_argument_checks(tmp_path):
    """reference src/main.cpp:20-41: the three argument errors, exit status 1."""
    for args, msg in (([], "You must provide a file name."), (["model.txt"], "Unrecognized file extension (must be .lp)."),
                      (["missing.lp"], "Invalid input file name.")):
        p = subprocess.run([hostlp.CLI_PATH] + args, cwd=tmp_path, capture_output=True, text=True)
        assert p.returncode == 1 and p.stdout.strip() == msg


@pytest.mark.parametrize("rule", ["stock", "dantzig"])
def test_synthetic_two_phase_through_model_api(golden_synth, rule):
    """G(m,n,n_eq,seed) entered through the model API and solved by Simplex::solve(): iteration
    and pivot counts of both phases, objective and the structural x against the reference."""
    seen = 0
    for rec in golden_synth:
        if rec["rule"] != rule or rec["m"] > 256:
            continue
        with hostlp.LpSession(pricing=rule) as s:
            s.build_synth(rec["m"], rec["n"], rec["n_eq"], rec["seed"])
            s.solve()
            calls = s.calls()
            assert [c["iterations"] for c in calls] == [c["iterations"] for c in rec["calls"]], (rec["m"], rec["n_eq"])
            assert [c["pivots"] for c in calls] == [len(c["pivots"]) for c in rec["calls"]]
            assert s.status() == "optimal"
            assert rel_close(s.objective(), rec["objective"])
            sol = s.solution()
            x = np.array([sol["x%07d" % j] for j in range(rec["n"])])
            assert rel_close(x, np.array(rec["x_struct"]))
            if rec["n_eq"] > 0:
                assert calls[1]["resident"] == 1
        seen += 1
    assert seen >= 5


@pytest.mark.parametrize("rule", ["stock", "dantzig"])
def test_dense_fast_path_matches_reference(golden_synth, rule):
    """solve_dense (host/dense_solver.cpp: the two-phase flow on plain arrays, no string-keyed model)
    on G(m,n,n_eq,seed): per-phase iteration and pivot counts, objective and x of the reference."""
    from oracle import oracle
    seen = 0
    for rec in golden_synth:
        if rec["rule"] != rule:
            continue
        m, n, n_eq = rec["m"], rec["n"], rec["n_eq"]
        c, a, rhs = oracle.synth(m, n, n_eq, rec["seed"])        # (c, a row-major, rhs): generator only
        types = ["="] * n_eq + ["<"] * (m - n_eq)
        got = hostlp.solve_dense(a, types, rhs, c, pricing=rule)
        assert got["status"] == "optimal"
        assert [k["iterations"] for k in got["calls"]] == [k["iterations"] for k in rec["calls"]], (m, n, n_eq)
        assert [k["pivots"] for k in got["calls"]] == [len(k["pivots"]) for k in rec["calls"]]
        assert rel_close(got["objective"], rec["objective"])
        assert rel_close(got["x_scaled"], np.array(rec["x_struct"]))
        assert np.all(a @ got["x"] <= rhs * (1 + 1e-9) + 1e-9)          # the unscaled x is feasible
        assert rel_close(float(c @ got["x"]), rec["objective"], 1e-8)
        if n_eq > 0:
            assert got["calls"][1]["resident"] == 1
        seen += 1
    assert seen >= 7


def test_dense_fast_path_with_bounds_matches_text_flow():
    """Bounded variables and '>=' rows through solve_dense == the same LP written as .lp text and
    solved through reader + presolve + Simplex::solve()."""
    rng = np.random.default_rng(5)
    for trial in range(6):
        m, n = 12, 18
        a = np.round(rng.uniform(0.1, 4.0, size=(m, n)), 2)
        x0 = rng.uniform(0.0, 3.0, size=n)
        types = [rng.choice(["<", ">", "="], p=[0.6, 0.25, 0.15]) for _ in range(m)]
        slack = rng.uniform(0.5, 3.0, size=m)
        rhs = np.round(a @ x0 + np.where(np.array(types) == "<", slack, np.where(np.array(types) == ">", -slack, 0.0)), 6)
        c = -np.round(rng.uniform(0.5, 2.0, size=n), 2)
        ub = np.round(x0 + rng.uniform(0.5, 2.0, size=n), 2)
        lines = ["Minimize", " obj: " + " ".join("%+.17g x%02d" % (c[j], j) for j in range(n)), "Subject To"]
        for i in range(m):
            rel = {"<": "<=", ">": ">=", "=": "="}[types[i]]
            lines.append(" r%02d: %s %s %.17g" % (i, " ".join("%+.17g x%02d" % (a[i, j], j) for j in range(n)), rel, rhs[i]))
        lines.append("Bounds")
        lines += [" x%02d <= %.17g" % (j, ub[j]) for j in range(n)]
        lines.append("End")
        with hostlp.LpSession() as s:
            rc = s.read("\n".join(lines) + "\n", check=False)
            assert rc == 0, s.error()
            try:
                s.solve()
                text_status, text_obj, text_calls, text_sol = s.status(), s.objective(), s.calls(), s.solution()
            except hostlp.LpError:
                text_status = "fatal"
        try:
            got = hostlp.solve_dense(a, types, rhs, c, ub=ub)
        except hostlp.LpError:
            assert text_status == "fatal"
            continue
        assert got["status"] == text_status, trial
        assert [(k["iterations"], k["pivots"], k["bound_flips"]) for k in got["calls"]] == \
               [(k["iterations"], k["pivots"], k["bound_flips"]) for k in text_calls], trial
        if text_status == "optimal":
            assert rel_close(got["objective"], text_obj)
            assert rel_close(got["x_scaled"], np.array([text_sol["x%02d" % j] for j in range(n)]))


REF_GPU = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "simplex_ref_gpu")


def test_reference_cli_with_the_gpu_seam_dropped_in(golden_cli, tmp_path):
    """The drop-in, dropped in: the REFERENCE's own command line (its parser, presolve, InitialBasis, solve,
    solution_found, XML writer, main — compiled from /root/reference/src in the build container) with only the
    body of Simplex::simplex replaced by the binding of include/binding/Simplex_seam_b200.cpp (oracle/Makefile,
    target ref_gpu). On every golden .lp text it exits like the stock binary and writes the stock binary's
    sol_test.xml (degenerate Phase-I fixtures: objective only)."""
    assert os.path.exists(REF_GPU), "oracle/_ref/simplex_ref_gpu missing: make -C oracle ref_gpu in the build container"
    n_exact = n_xml = 0
    for name, rec in golden_cli.items():
        wd = tmp_path / re.sub(r"\W", "_", name)
        wd.mkdir()
        (wd / "in.lp").write_text(rec["lp_text"])
        p = subprocess.run([REF_GPU, "in.lp"], cwd=wd, capture_output=True, text=True, timeout=300)
        xml_file = wd / "sol_test.xml"
        if rec.get("degenerate"):
            if rec["xml"] is not None and xml_file.exists() and p.returncode == 0 and not rec["unbounded"]:
                got_obj, _ = parse_xml(xml_file.read_text())
                want_obj, _ = parse_xml(rec["xml"])
                if "LP IS UNBOUNDED" not in p.stdout and abs(want_obj) < 1e100:
                    assert abs(got_obj - want_obj) <= 1.5e-6 * max(1.0, abs(want_obj)), (name, got_obj, want_obj)
            continue
        assert p.returncode == rec["returncode"], (name, p.returncode, p.stdout[-400:], p.stderr[-400:])
        assert xml_file.exists() == (rec["xml"] is not None), name
        if rec["xml"] is None:
            continue
        n_xml += 1
        got_obj, got_rows = parse_xml(xml_file.read_text())
        want_obj, want_rows = parse_xml(rec["xml"])
        assert [n for n, _ in got_rows] == [n for n, _ in want_rows], name
        degenerate = name.startswith("fixture:") and any(d[0] == name[8:] for d in DEGENERATE_CALLS)
        if np.isfinite(want_obj):
            assert abs(got_obj - want_obj) <= 1.5e-6 * max(1.0, abs(want_obj)), (name, got_obj, want_obj)
        if not degenerate:
            for (n, g), (_, w) in zip(got_rows, want_rows):
                if np.isfinite(w):
                    assert abs(g - w) <= 1.5e-6 * max(1.0, abs(w)), (name, n, g, w)
        n_exact += xml_file.read_text() == rec["xml"]
    assert n_xml >= 40 and n_exact >= n_xml - 4, (n_exact, n_xml)


@pytest.mark.parametrize("engine,dual", [("persistent", "update"), ("streaming", "update"), ("persistent", "recompute"),
                                         ("staged", "update")])
def test_bounded_mid_size_lps_match_the_reference(golden_bounded, engine, dual):
    """Dense LPs with finite upper bounds, 60 ... 150 rows (one of them two-phase), through the whole flow (.lp text ->
    reader -> presolve -> Phase I / II on the GPU) for both pricing rules, against the reference's own runs: per-call
    iteration and pivot counts (bound-flip iterations included), status, objective and every variable within 1e-9.
    ("persistent", "update") is the bounded variant of the shared-memory-resident kernel at these sizes, ("streaming",
    "update") the fused kernel's: t2 candidates, both kinds of flips and the substitution of leaving variables spread
    over all CTAs."""
    n_runs = 0
    for name, case in golden_bounded.items():
        for rule, rec in case["runs"].items():
            with hostlp.LpSession(pricing=rule, engine=engine, dual=dual, chunk_iters=64) as s:
                s.read(case["lp_text"])
                s.solve()
                calls = s.calls()
                assert s.status() == rec["status"] == "optimal", (name, rule)
                assert [(c["iterations"], c["pivots"]) for c in calls] == \
                       [(c["iterations"], len(c["pivots"])) for c in rec["calls"]], (name, rule)
                assert sum(c["bound_flips"] for c in calls) > 0, (name, rule)
                assert rel_close(s.objective(), rec["objective"]), (name, rule, s.objective(), rec["objective"])
                sol = s.solution()
                assert sorted(sol) == sorted(rec["solution"]), (name, rule)
                for k, v in rec["solution"].items():
                    assert rel_close(sol[k], v), (name, rule, k, sol[k], v)
            n_runs += 1
    assert n_runs >= 7
