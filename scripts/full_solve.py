"""One LP G(m, n, 0, seed) solved to the optimum on one GPU (fused engine): pivots, time, objective, hygiene.
usage: full_solve.py M N [seed]"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplex_b200 import Basis, DeviceSimplex, algorithmic_bytes_per_pivot, synth

m, n = int(sys.argv[1]), int(sys.argv[2])
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1234
A, b, c, basic, nonbasic = synth.tableau(m, n, 0, seed)
with DeviceSimplex(pricing="most_neg", engine="persistent", dual="update", chunk_iters=256) as s:
    s.load(A, b, c)
    basis = Basis(basic, nonbasic)
    t0 = time.time()
    r = s.simplex(basis)
    wall = time.time() - t0
    out = {"m": m, "n": n, "seed": seed, "engine": s.engine_used(), "term": r["term"], "pivots": r["pivots"],
           "iterations": r["iterations"], "objective": r["objective"], "loop_seconds": r["loop_seconds"], "wall_seconds": wall,
           "us_per_pivot": 1e6 * r["loop_seconds"] / max(r["pivots"], 1),
           "algorithmic_TBs": algorithmic_bytes_per_pivot(m, n) * r["pivots"] / r["loop_seconds"] / 1e12,
           "hygiene": s.hygiene_stats(), "primal_residual": s.primal_residual(), "inverse_residual_32_rows": s.inverse_residual(32, 3),
           "min_xB": float(s.xB().min())}
print(json.dumps(out), flush=True)
