#!/usr/bin/env python3
"""Regenerates tests/golden/*.json from the UNMODIFIED reference (test infrastructure).

Runs here, in the build container, where /root/reference exists; the GPU box only sees the
committed JSON. The reference sources are compiled where they lie by `make -C oracle ref`
(against the Eigen stand-in oracle/eigen_shim, see oracle/Makefile) into oracle/_ref/:

    ref_driver_stock    reference as shipped (first-eligible pricing, Simplex.cpp:249)
    ref_driver_dantzig  the one-line switch the reference keeps commented (Simplex.cpp:250)

Outputs
    fixtures.json.gz  the 14 .lp files of /root/reference/data/test_set_1 plus the Klee-Minty n=10
                      cube written by /root/reference/data/gen_kleeminty.py: LP text (the input
                      vector), status, objective and solution at full precision, and for every
                      Simplex::simplex() call the tableau at entry, the index lists in/out, the
                      pivot trace and the iteration count
    synth.json.gz     the synthetic table of SURVEY.md App. A.3 (+ a few more seeds): traces,
                      objectives, structural x, and small entry tableaux that pin the generator
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SB200_REFERENCE", "/root/reference")
DRV = {"stock": os.path.join(ROOT, "oracle/_ref/ref_driver_stock"),
       "dantzig": os.path.join(ROOT, "oracle/_ref/ref_driver_dantzig")}


def run_driver(kind, args, dump=True, via_solve=False):
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out.json")
        cmd = [DRV[kind]] + [str(a) for a in args] + ["--out", out]
        if dump:
            cmd.append("--dump-tableau")
        if via_solve:
            cmd.append("--via-solve")
        subprocess.run(cmd, cwd=td, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)
        with open(out) as f:
            return json.load(f)


def complete_lists(call):
    """Phase-I calls happen inside InitialBasis::get_basis, so the driver cannot see the Basis
    object: rebuild it. basic_in is the first logged basis; nonbasic_in is the ascending
    complement (InitialBasis.cpp:9-18); the out lists follow by replaying the logged swaps
    (Basis.cpp:48-53)."""
    if "basic_in" in call:
        return
    basic = list(call["first_basis"])
    inb = set(basic)
    nonbasic = [j for j in range(call["N"]) if j not in inb]
    call["basic_in"] = list(basic)
    call["nonbasic_in"] = list(nonbasic)
    for lc, row, ec, pos in call["pivots"]:
        assert basic[row] == lc and nonbasic[pos] == ec
        basic[row] = ec
        nonbasic[pos] = lc
    assert basic == call["last_basis"]
    call["basic_out"] = basic
    call["nonbasic_out"] = nonbasic


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not found: golden vectors can only be regenerated in the build container")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL)

    # ---------------------------------------------------------------- fixtures
    fixtures = {}
    d = os.path.join(REF, "data/test_set_1")
    files = sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".lp"))
    with tempfile.TemporaryDirectory() as td:
        subprocess.run([sys.executable, os.path.join(REF, "data/gen_kleeminty.py"), "10"], cwd=td, check=True)
        km = os.path.join(td, "klee-minty-n10.lp")
        cases = [(os.path.basename(f)[:-3], f) for f in files] + [("gen_kleeminty_n10_rhs5", km)]
        for name, path in cases:
            rec = run_driver("stock", ["lp", path])
            chk = run_driver("stock", ["lp", path], dump=False, via_solve=True)
            # the replicated body of Simplex::solve() must behave exactly like solve() itself
            assert rec["status"] == chk["status"], name
            assert rec.get("objective") == chk.get("objective"), name
            assert [c["pivots"] for c in rec["calls"]] == [c["pivots"] for c in chk["calls"]], name
            for c in rec["calls"]:
                complete_lists(c)
            rec["lp_text"] = open(path).read()
            fixtures[name] = rec
            print(name, rec["status"], [(c["iterations"], len(c["pivots"])) for c in rec["calls"]], rec.get("objective"))
    with gzip.open(os.path.join(HERE, "fixtures.json.gz"), "wt") as f:
        json.dump(fixtures, f, separators=(",", ":"))

    # ---------------------------------------------------------------- synthetic G(m,n,n_eq,seed)
    table = [  # (m, n, n_eq, seed, rules, dump tableau?)
        (8, 16, 0, 1234, ("stock", "dantzig"), True),
        (16, 32, 4, 1234, ("stock", "dantzig"), True),
        (32, 64, 0, 1234, ("stock", "dantzig"), True),
        (64, 128, 0, 1234, ("stock", "dantzig"), False),
        (64, 128, 0, 1235, ("dantzig",), False),
        (64, 128, 0, 1236, ("dantzig",), False),
        (64, 128, 16, 1234, ("stock", "dantzig"), False),
        (128, 256, 0, 1234, ("stock", "dantzig"), False),
        (128, 256, 32, 1234, ("stock", "dantzig"), False),
        (200, 300, 50, 7, ("dantzig",), False),
        (256, 1024, 0, 1234, ("dantzig",), False),
        (256, 512, 64, 1234, ("dantzig",), False),
        (512, 1024, 0, 1234, ("dantzig",), False),
    ]
    synth = []
    for m, n, n_eq, seed, rules, dump in table:
        for rule in rules:
            rec = run_driver(rule, ["synth", m, n, n_eq, seed], dump=dump)
            for c in rec["calls"]:
                complete_lists(c)
                c.pop("first_basis", None)
            sol = rec.pop("solution")
            rec["x_struct"] = [sol["x%07d" % j] for j in range(n)]
            for k in ("var_names", "row_labels", "var_shifts"):
                rec.pop(k, None)
            rec.update({"m": m, "n": n, "n_eq": n_eq, "seed": seed, "rule": rule})
            synth.append(rec)
            print(m, n, n_eq, seed, rule, rec["status"], [(c["iterations"], len(c["pivots"])) for c in rec["calls"]],
                  "%.9f" % rec["objective"])
    with gzip.open(os.path.join(HERE, "synth.json.gz"), "wt") as f:
        json.dump(synth, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
