#!/usr/bin/env python3
"""Regenerates tests/golden/parser_cases.json.gz (test infrastructure; build container only).

Hand-written .lp texts that exercise the dialect of the reference's reader and presolve
(reference src/Parser.cpp, src/Presolve.cpp) are run through the UNMODIFIED reference
(oracle/_ref/ref_driver_stock, see oracle/Makefile). For each text the outcome class (solved /
"Exception: ..." / abort / exit(0) in presolve), ids, row order, shifts and the tableau at the
entry of the first Simplex::simplex() call are recorded. tests/test_host_model.py replays the
texts through this repo's host library and compares.
"""
import gzip
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRV = os.path.join(ROOT, "oracle/_ref/ref_driver_stock")

CASES = {
    # three comment markers; a trailing comment swallows its newline and glues two lines together
    "comments": """\\ backslash comment
# hash comment
Maximize // trailing comment glues the next line on
 obj: 3 x + 2 y
Subject To
 c1: x + y <= 4
 c2: x + 3 y <= 6 # glued to c3's line, which therefore disappears into this row's rhs text
 c3: x <= 3
End
""",
    "case_and_spacing": """MAXIMIZE
 Obj : 3 x+2 y
   - z
SUBJECT   TO
 Row1 : x + y
      + z <= 10
 Row2 : x - y >= -2
 Row3 : x + 3 y + z = 9
BOUNDS
 z <= 4
END
""",
    "label_order": """Minimize
 cost: 2 a + 3 b + 4 c
Subject To
 c10: a + b >= 2
 c2: b + c >= 3
 c1: a + c >= 4
 C3: a + b + c <= 20
End
""",
    "unlabelled_rows_collide": """Maximize
 x + y + w
Subject To
 x + y <= 4
 x + w <= 5
 named: y + w <= 3
End
""",
    "exponent_sign_splits_term": """Minimize
 obj: 1e-1 x + 2 y
Subject To
 r1: x + y >= 1
 r2: 2.5e+1 x + y <= 30
End
""",
    "duplicate_terms_and_tiny": """Minimize
 obj: x + y
Subject To
 r1: x + x + y >= 3
 r2: 0.00000001 x + y <= 8
 r3: x - y <= 2
End
""",
    "rhs_forms": """Maximize
 obj: x + 2 y
Subject To
 r1: x + y <= +7
 r2: x - y => -3
 r3: x =< 5
 r4: y <= .5
End
""",
    "equalities_negative_rhs": """Minimize
 obj: x + y + z
Subject To
 e1: x - y = -1
 e2: -x - z <= -2
 e3: y + z >= 1
End
""",
    "bounds_forms": """Maximize
 obj: x + y + z + w
Subject To
 r1: x + y + z + w <= 100
 r2: x - y >= -50
Bounds
 1 <= x <= 10
 y <= 20
 2 <= z
 -1 <= w <= 3
End
""",
    "bounds_inf": """Maximize
 obj: x + y
Subject To
 r1: x + y <= 10
Bounds
 x <= +inf
 0 <= y <= 4
End
""",
    "bounds_unknown_variable": """Maximize
 obj: x + y
Subject To
 r1: x + y <= 10
Bounds
 q <= 4
End
""",
    "bounds_ge_is_not_understood": """Maximize
 obj: x + y
Subject To
 r1: x + y <= 10
Bounds
 x >= 2
End
""",
    "objective_repeats_variable": """Maximize
 obj: x + y + x
Subject To
 r1: x + y <= 10
End
""",
    "objective_constant_token": """Maximize
 obj: x + 3
Subject To
 r1: x <= 10
End
""",
    "objective_only_variable_zero_column": """Maximize
 obj: x + y
Subject To
 r1: x <= 10
End
""",
    "presolve_redundant_row": """Maximize
 obj: x + y
Subject To
 r1: x + y <= 100
 r2: x - y <= 3
Bounds
 x <= 4
 y <= 5
End
""",
    "presolve_infeasible_row": """Maximize
 obj: x + y
Subject To
 r1: x + y >= 100
Bounds
 x <= 4
 y <= 5
End
""",
    "presolve_lb_above_ub": """Maximize
 obj: x
Subject To
 r1: x <= 10
Bounds
 5 <= x <= 3
End
""",
    "presolve_fixed_variable": """Maximize
 obj: x + y
Subject To
 r1: x + y <= 10
Bounds
 3 <= x <= 3
End
""",
    "short_first_line": """mx
 obj: x
Subject To
 r1: x <= 1
End
""",
    "no_end_keyword": """Minimize
 obj: x + y
Subject To
 r1: x + y >= 2
 r2: x - y <= 1""",
    "minimize_with_shift": """Minimize
 obj: 2 x + 3 y
Subject To
 r1: x + y >= 6
 r2: x - y <= 2
Bounds
 1 <= x <= 8
 2 <= y <= 9
End
""",
}


sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from lp_text_gen import random_lp_text  # noqa: E402
from oracle import oracle  # noqa: E402

# seeded random texts inside the dialect (bounded variables, equalities, '>=' rows, split rows): they
# exercise Phase I, the bounded ratio test and the flip reversal far more than the 14 fixtures do
for _seed in range(100, 124):
    CASES["random_%d" % _seed] = random_lp_text(_seed)
for _seed in range(200, 206):
    CASES["random_big_%d" % _seed] = random_lp_text(_seed, m=24, n=36)


def main():
    if not os.path.exists(DRV):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL)
    out = {}
    for name, text in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, name + ".lp")
            with open(path, "w") as f:
                f.write(text)
            res = os.path.join(td, "out.json")
            try:
                p = subprocess.run([DRV, "lp", path, "--out", res, "--dump-tableau"], cwd=td, stdout=subprocess.DEVNULL,
                                   stderr=subprocess.PIPE, timeout=20)
                rc = p.returncode
            except subprocess.TimeoutExpired:
                out[name] = {"lp_text": text, "status": "hang", "calls": []}
                print(name, "HANG")
                continue
            if os.path.exists(res):
                raw = open(res).read()
                # the driver prints non-finite doubles the printf way
                raw = re.sub(r"(?<![\w\"])-?nan(?![\w\"])", "NaN", raw)
                raw = re.sub(r"(?<![\w\"])(-?)inf(?![\w\"])", r"\1Infinity", raw)
                rec = json.loads(raw)
            else:
                # uncaught std::invalid_argument / std::out_of_range: the reference process aborted
                rec = {"status": "exception_cstr", "calls": [], "abort_rc": rc}
            keep = {"lp_text": text, "status": rec["status"], "exception": rec.get("exception"),
                    "var_names": rec.get("var_names"), "row_labels": rec.get("row_labels"),
                    "var_shifts": rec.get("var_shifts"), "sense": rec.get("sense"), "objective": rec.get("objective"),
                    "calls": []}
            for c in rec.get("calls", [])[:1]:
                basic = list(c.get("basic_in") or c["first_basis"])
                inb = set(basic)
                keep["calls"].append({"M": c["M"], "N": c["N"], "A": c["A"], "b": c["b"], "c": c["c"], "ub": c["ub"],
                                      "basic_in": basic,
                                      "nonbasic_in": c.get("nonbasic_in") or [j for j in range(c["N"]) if j not in inb]})
            # m_has_UBS (LinearProgram.cpp:151) is set by the first finite upper bound the reader stores
            keep["bounded"] = any(u < sys.float_info.max for c in keep["calls"] for u in c["ub"])
            # Degenerate vertices: the reference's fresh LU and an updated inverse round tied ratios
            # differently, so the two may walk different, equally valid paths. A case
            # is flagged when the oracle's LU and product-form variants disagree on any recorded call;
            # for those only the objective is pinned (north_star: "on non-degenerate instances").
            keep["degenerate"] = False
            for c in rec.get("calls", []):
                if "A" not in c:
                    continue
                basic = list(c.get("basic_in") or c["first_basis"])
                inb = set(basic)
                nonb = c.get("nonbasic_in") or [j for j in range(c["N"]) if j not in inb]
                A = np.array(c["A"], dtype=np.float64).reshape((c["N"], c["M"])).T.copy(order="F")
                ub = np.array(c["ub"]) if keep["bounded"] else None
                runs = [oracle.simplex(A, np.array(c["b"]), np.array(c["c"]), np.array(basic), np.array(nonb), ub=ub,
                                       rule=oracle.RULE_FIRST, variant=v, iter_limit=100000) for v in ("lu", "pfi")]
                if (runs[0]["trace"].tolist() != runs[1]["trace"].tolist() or runs[0]["term"] != runs[1]["term"]
                        or runs[0]["trace"].tolist() != [list(p) for p in c["pivots"]]):
                    keep["degenerate"] = True
            keep["n_calls"] = len(rec.get("calls", []))
            keep["solution"] = rec.get("solution")
            keep["call_stats"] = [{"M": c["M"], "N": c["N"], "iterations": c["iterations"], "pivots": len(c["pivots"]),
                                   "optimal": c["optimal"], "unbounded": c["unbounded"]} for c in rec.get("calls", [])]
            out[name] = keep
            print(name, keep["status"], keep["exception"], [(c["M"], c["N"]) for c in keep["calls"]], "degenerate" if keep["degenerate"] else "")
    with gzip.open(os.path.join(HERE, "parser_cases.json.gz"), "wt") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
