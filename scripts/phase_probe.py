"""Per-phase time of the fused persistent engine on G(m, n, 0, seed): scripts/phase_probe.py m n pivots"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from simplex_b200 import Basis, DeviceSimplex, synth

m, n, P = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 1234)
for dual in ("update", "recompute"):
    s = DeviceSimplex(pricing="most_neg", engine="persistent", dual=dual, iter_limit=P, chunk_iters=256)
    s.load(A, b, c)
    best = None
    for rep in range(3):
        r = s.simplex(Basis(basic, nonbasic))
        best = r["loop_seconds"] if best is None else min(best, r["loop_seconds"])
    pc = s.phase_cycles()
    it = max(int(pc[7]), 1)
    print(dual, "us/pivot %.2f" % (1e6 * best / r["pivots"]), "phases", [round(float(pc[i]) / it / 1965.0, 2) for i in range(7)], flush=True)
    s.close()
