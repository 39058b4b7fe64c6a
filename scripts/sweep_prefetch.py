"""Sweep of the L2-prefetch knobs of the fused engine on the headline LP (one load, many runs)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from simplex_b200 import Basis, DeviceSimplex, synth

m, n, P = 4096, 16384, 1024
A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 1234)
s = DeviceSimplex(pricing="most_neg", engine="persistent", dual="update", iter_limit=P, chunk_iters=256, trace_capacity=P)
s.load(A, b, c)
ref = None
combos = [(0, 0), (8, 0), (12, 0), (16, 0), (24, 0), (32, 0), (0, 8), (0, 16), (0, 28), (12, 8), (16, 16), (32, 28)]
if len(sys.argv) > 1:
    combos = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for cols, rows in combos:
    os.environ["SB200_PREFETCH_COLS"] = str(cols)
    os.environ["SB200_PREFETCH_ROWS"] = str(rows)
    best = None
    for rep in range(3):
        bs = Basis(basic, nonbasic)
        r = s.simplex(bs)
        best = r["loop_seconds"] if best is None else min(best, r["loop_seconds"])
    tr, _ = s.trace()
    if ref is None:
        ref = tr.copy()
    pc = s.phase_cycles()
    it = max(int(pc[7]), 1)
    ph = [round(float(pc[i]) / it / 1965.0, 1) for i in range(5)]
    print("cols %2d rows %2d: %.2f us/pivot  %.0f pivots/s  same_trace %s phases(price,bar1,fused,bar2,tail) %s" % (
        cols, rows, 1e6 * best / r["pivots"], r["pivots"] / best, bool(np.array_equal(tr, ref)), ph), flush=True)
s.close()
