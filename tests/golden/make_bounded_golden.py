#!/usr/bin/env python3
"""Golden vectors of the UNMODIFIED reference on mid-size dense LPs WITH finite upper bounds (test infrastructure).

The reference's own bounded fixtures have at most 10 rows; these are big enough for the bounded ratio test, both
kinds of bound flips and the substitution of leaving variables to be spread over many CTAs of the fused kernel
(reference src/Simplex.cpp:116-179, src/LinearProgram.cpp:154-169). Runs only in the build container:

    python tests/golden/make_bounded_golden.py        -> tests/golden/bounded.json.gz

Each case is a CPLEX-LP text written by `bounded_lp_text` below (continuous random data: no ties), solved by
oracle/_ref/ref_driver_{stock,dantzig} `lp FILE` (the reference's parser, presolve, Simplex::solve). Kept: the
text, status, objective and solution at full precision, and per Simplex::simplex() call the iteration count, the
pivot trace and the index lists (no tableau: the GPU tests rebuild it from the text through the host library,
whose tableaux are pinned bit for bit against the reference's elsewhere)."""
import gzip
import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import complete_lists, run_driver  # noqa: E402


def bounded_lp_text(seed, m, n, n_eq=0, tight=0.4, lower=0.0):
    """min c.x, a_i.x <= rhs_i (the first n_eq rows '='), 0 <= x_j <= u_j: a share `tight` of the bounds small
    enough to be reached, a few variables with a positive lower bound (presolve shifts them)."""
    rng = random.Random(seed)
    x0 = [rng.random() for _ in range(n)]
    out = ["\\ dense bounded LP, seed %d" % seed, "Minimize", " obj: " + " ".join(
        "- %.10g x%04d" % (0.5 + rng.random(), j) for j in range(n)), "Subject To"]
    for i in range(m):
        a = [0.05 + 0.95 * rng.random() for _ in range(n)]
        lhs = sum(v * x for v, x in zip(a, x0))
        if i < n_eq:
            rel, rhs = "=", lhs
        else:
            rel, rhs = "<=", lhs * (1.0 + 0.5 * rng.random())
        out.append(" R%04d: " % i + " + ".join("%.10g x%04d" % (v, j) for j, v in enumerate(a)) + " %s %.10g" % (rel, rhs))
    out.append("Bounds")
    for j in range(n):
        u = x0[j] + (rng.uniform(0.01, 0.3) if rng.random() < tight else rng.uniform(1.0, 30.0))
        if rng.random() < lower:
            out.append(" %.10g <= x%04d <= %.10g" % (0.5 * x0[j], j, u))
        else:
            out.append(" x%04d <= %.10g" % (j, u))
    out.append("End")
    return "\n".join(out) + "\n"


CASES = [  # name, seed, m, n, n_eq, tight, lower, rules
    ("b60x100", 11, 60, 100, 0, 0.4, 0.0, ("stock", "dantzig")),
    ("b120x200", 12, 120, 200, 0, 0.5, 0.0, ("stock", "dantzig")),
    ("b150x400", 13, 150, 400, 0, 0.3, 0.1, ("dantzig",)),
    ("b96x160_eq24", 14, 96, 160, 24, 0.4, 0.0, ("stock", "dantzig")),
]


def main():
    gold = {}
    with tempfile.TemporaryDirectory() as td:
        for name, seed, m, n, n_eq, tight, lower, rules in CASES:
            text = bounded_lp_text(seed, m, n, n_eq, tight, lower)
            path = os.path.join(td, name + ".lp")
            with open(path, "w") as f:
                f.write(text)
            gold[name] = {"lp_text": text, "m": m, "n": n, "n_eq": n_eq, "runs": {}}
            for rule in rules:
                rec = run_driver(rule, ["lp", path], dump=False)
                for c in rec["calls"]:
                    complete_lists(c)
                    c.pop("first_basis", None)
                    c.pop("last_basis", None)
                gold[name]["runs"][rule] = {"status": rec["status"], "objective": rec.get("objective"),
                                            "solution": rec.get("solution"), "calls": rec["calls"]}
                print(name, rule, rec["status"], [(c["iterations"], len(c["pivots"])) for c in rec["calls"]], rec.get("objective"))
    path = os.path.join(HERE, "bounded.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(gold, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
