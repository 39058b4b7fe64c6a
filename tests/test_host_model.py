"""Host side (simplex_b200/host: text model, .lp reader, presolve, phase control) against the
golden vectors recorded from the unmodified reference (tests/golden/make_golden.py,
make_parser_cases.py). CPU only: the host halves of Simplex::solve() are driven one step at a
time through the C API of libsimplex_b200_host.so; where the flow needs the result of a device
loop (final lists, x_B, flip flags) the test feeds the reference's recorded lists and the
oracle's vectors. Bit-exact for everything that reaches the device (A, b, c, ub, index lists).
"""
import os
import subprocess

import numpy as np
import pytest

from helpers import DBL_MAX, call_inputs, has_bounds, rel_close
from oracle import oracle
from simplex_b200 import lp as hostlp


def same_bits(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def check_seam_inputs(session, call, bounded, tag):
    """The tableau and lists the host would upload == what the reference's loop saw at entry."""
    A, b, c, ub = session.tableau()
    Ar, br, cr, ubr, bas_r, non_r = call_inputs(call)
    assert A.shape == Ar.shape, tag
    assert same_bits(A, Ar), tag
    assert same_bits(b, br), tag
    assert same_bits(c, cr), tag
    assert session.dims()["bounded"] == bounded, tag
    assert same_bits(ub, ubr), tag
    basic, non_basic = session.basis()
    assert basic.tolist() == bas_r.tolist(), tag
    assert non_basic.tolist() == non_r.tolist(), tag
    return Ar, br, cr, ubr


def finish_call_like_the_device(session, call, inputs, bounded):
    """Stand-in for the device loop in a CPU test: final lists from the reference's record,
    x_B and flip flags from the oracle run on the same tableau (checked to end on those lists)."""
    Ar, br, cr, ubr = inputs
    res = oracle.simplex(Ar, br, cr, np.array(call["basic_in"]), np.array(call["nonbasic_in"]),
                         ub=ubr if bounded else None, rule=oracle.RULE_FIRST, variant="lu")
    assert res["basic"].tolist() == call["basic_out"] and res["nonbasic"].tolist() == call["nonbasic_out"]
    session.set_basis(call["basic_out"], call["nonbasic_out"])
    if call["optimal"]:
        session.absorb_optimum(res["x"], res["flip"] if bounded else None)
    return res


def test_two_phase_flow_matches_reference_on_every_fixture(golden_fixtures):
    """All 15 recorded runs: names, ids, row order, shifts, every tableau handed to the seam,
    crash basis, artificial removal, and the solution read-out."""
    for name, rec in golden_fixtures.items():
        bounded = has_bounds(rec)
        calls = rec["calls"]
        with hostlp.LpSession() as s:
            s.read(rec["lp_text"])
            assert s.sense() == rec["sense"], name
            n_art = s.open_phase_one()
            assert (n_art > 0) == (len(calls) == 2), name
            if n_art > 0:
                inputs = check_seam_inputs(s, calls[0], bounded, (name, "phase I"))
                finish_call_like_the_device(s, calls[0], inputs, bounded)
                s.close_phase_one()
            s.initialize_tableau()
            last = calls[-1]
            inputs = check_seam_inputs(s, last, bounded, (name, "phase II"))
            ids, rows = s.names()
            assert [ids[k] for k in sorted(ids)] == rec["var_names"], name
            assert sorted(ids) == list(range(len(ids))), name
            assert rows == rec["row_labels"], name
            assert {str(k): v for k, v in s.shifts().items()} == rec["var_shifts"], name
            finish_call_like_the_device(s, last, inputs, bounded)
            if rec["status"] == "optimal":
                assert s.status() == "optimal"
                assert rel_close(s.objective(), rec["objective"]), (name, s.objective(), rec["objective"])
                sol = s.solution()
                assert sorted(sol) == sorted(rec["solution"]), name
                for k, v in rec["solution"].items():
                    assert rel_close(sol[k], v), (name, k, sol[k], v)


def test_xml_writer_format(golden_fixtures, tmp_path):
    """write() gives the reference's sol_test.xml text (fixed, 6 decimals, name order)."""
    rec = golden_fixtures["test_bnds_1"]
    with hostlp.LpSession() as s:
        s.read(rec["lp_text"])
        assert s.open_phase_one() == 0
        s.initialize_tableau()
        call = rec["calls"][0]
        inputs = call_inputs(call)[:4]
        finish_call_like_the_device(s, call, inputs, True)
        out = tmp_path / "sol_test.xml"
        s.write(str(out))
    assert out.read_text() == (
        '<?xml version = "1.0" encoding="UTF-8" standalone="yes"?>\n<TestSolution>\n<header objectiveValue="36.000000"/>\n'
        '<variable name="__SLACK0" value="0.000000"/>\n<variable name="__SLACK1" value="3.000000"/>\n'
        '<variable name="x1" value="6.000000"/>\n<variable name="x2" value="8.000000"/>\n</TestSolution>\n')


def test_reader_dialect_cases(golden_parser_cases):
    """Hand-written .lp texts that exercise the dialect (comments, squeezed white space, split
    rules, unlabelled rows, bounds forms, presolve exits): same outcome class as the reference and,
    when the reference reaches the loop, bit-identical first tableau, ids, row order and shifts."""
    seen = 0
    for name, rec in golden_parser_cases.items():
        with hostlp.LpSession() as s:
            rc = s.read(rec["lp_text"], check=False)
            status = rec["status"]
            if status == "exception_string":
                assert rc == hostlp.READ_EXCEPTION, (name, rc, s.error())
                assert s.error() == "Exception: " + rec["exception"], name
                continue
            if status == "exception_cstr" and not rec["calls"]:
                assert rc == hostlp.READ_FATAL, (name, rc, s.error())
                continue
            if status == "exit_called":
                assert rc == hostlp.READ_PRESOLVE_INFEASIBLE, (name, rc, s.error())
                continue
            assert rc == hostlp.READ_OK, (name, rc, s.error())
            bounded = rec["bounded"]
            n_art = s.open_phase_one()
            if n_art == 0:
                s.initialize_tableau()
            check_seam_inputs(s, rec["calls"][0], bounded, name)
            seen += 1
    assert seen >= 10


def test_synthetic_models_through_the_model_api(golden_synth):
    """G(m,n,n_eq,seed) entered through add_variable/add_constraint: the tableau and crash basis of
    the first seam call are bit-identical to the reference's (which built the same LP through its
    own model API); for the larger records (no tableau dump) the crash lists are compared."""
    seen = set()
    for rec in golden_synth:
        key = (rec["m"], rec["n"], rec["n_eq"], rec["seed"])
        if key in seen or rec["m"] > 256:
            continue
        seen.add(key)
        with hostlp.LpSession() as s:
            s.build_synth(*key)
            n_art = s.open_phase_one()
            assert n_art == rec["n_eq"]
            if n_art == 0:
                s.initialize_tableau()
            call = rec["calls"][0]
            if "A" in call:
                check_seam_inputs(s, call, False, key)
            else:
                basic, non_basic = s.basis()
                assert basic.tolist() == call["basic_in"] and non_basic.tolist() == call["nonbasic_in"], key
                d = s.dims()
                assert (d["rows"], d["cols"]) == (call["M"], call["N"])
    assert len(seen) >= 8


def test_reference_unit_tests_constraint_and_model():
    """The reference's own two gtest cases (tests/UnitTests.cpp:12-56) restated against the host
    library through a tiny C++ program: same calls, same expectations."""
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(here, "..", "simplex_b200", "host_unit_tests")
    if not os.path.exists(exe):
        pytest.skip("host_unit_tests not built (simplex_b200/host/build.sh)")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL PASSED" in out.stdout


def test_no_cpu_loop_in_host_library():
    """Without a GPU the whole flow must fail loudly at the seam, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with hostlp.LpSession() as s:
        s.read("Maximize\n obj: x + y\nSubject To\n c1: x + y <= 4\nEnd\n")
        with pytest.raises(hostlp.LpError) as e:
            s.solve()
        assert "GPU solver unavailable" in str(e.value)
