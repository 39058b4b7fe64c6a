"""Shared-memory-resident engine against the fused engine on G(m, n, n_eq, seed):
    scripts/resident_check.py [m n n_eq pivots chunk] ...
For every shape: both engines from the same start (Phase-I tableau when n_eq > 0), same pivot trace,
bit-identical x_B / y / B^-1, time per pivot and the per-phase breakdown of CTA 0
(price, exchange 1, fused pass, exchange 2, tail)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from simplex_b200 import Basis, DeviceSimplex, synth


def run(m, n, n_eq, P, chunk, resident, reps=3):
    if resident:
        os.environ.pop("SB200_NO_RESIDENT", None)
    else:
        os.environ["SB200_NO_RESIDENT"] = "1"
    A, b, c, basic, nonbasic = synth.tableau(m, n, n_eq, 1234, phase1=(n_eq > 0))
    with DeviceSimplex(pricing="most_neg", engine="persistent", dual="update", iter_limit=P, chunk_iters=chunk,
                       trace_capacity=max(P, 1)) as s:
        s.load(A, b, c)
        best = None
        for _ in range(reps):
            bs = Basis(basic, nonbasic)
            r = s.simplex(bs)
            best = r["loop_seconds"] if best is None else min(best, r["loop_seconds"])
        pc = s.phase_cycles()
        it = max(int(pc[7]), 1)
        tr, _ = s.trace()
        out = {"engine": s.engine_used(), "term": r["term"], "pivots": r["pivots"], "iterations": r["iterations"],
               "us_per_pivot": 1e6 * best / max(r["pivots"], 1), "objective": r["objective"],
               "phases_us": [round(float(pc[i]) / it / 1965.0, 2) for i in range(5)]}
        return out, tr, s.xB(), s.y(), s.Binv(), bs


def main():
    args = [int(v) for v in sys.argv[1:]]
    shapes = [tuple(args[i:i + 5]) for i in range(0, len(args), 5)] or [(64, 128, 0, 40, 16), (96, 200, 24, 10 ** 6, 64),
                                                                        (1024, 2048, 0, 600, 256),
                                                                        (1024, 2048, 256, 10 ** 6, 256)]
    ok = True
    for m, n, n_eq, P, chunk in shapes:
        ro, tro, xo, yo, Bo, bso = run(m, n, n_eq, P, chunk, True)
        rf, trf, xf, yf, Bf, bsf = run(m, n, n_eq, P, chunk, False)
        same = {"trace": bool(np.array_equal(tro, trf)), "x_bits": bool(np.array_equal(xo, xf)),
                "y_bits": bool(np.array_equal(yo, yf)), "Binv_bits": bool(np.array_equal(Bo, Bf)),
                "lists": bool(np.array_equal(bso.basic, bsf.basic) and np.array_equal(bso.non_basic, bsf.non_basic)),
                "counts": (ro["pivots"], ro["iterations"], ro["term"]) == (rf["pivots"], rf["iterations"], rf["term"])}
        ok = ok and all(same.values()) and ro["engine"] == "resident" and rf["engine"] == "fused"
        print(json.dumps({"shape": [m, n, n_eq, P, chunk], "resident": ro, "fused": rf, "same": same}), flush=True)
    print("RESIDENT_CHECK", "OK" if ok else "MISMATCH", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
