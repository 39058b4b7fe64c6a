"""Sweep of SB200_L2_KEEP_ROWS (rows of B^-1 per CTA read and written with the L2 evict-last policy) together with
the row prefetch, on the headline LP (one load, many runs). usage: sweep_keep_rows.py [keep,prefetch_rows,prefetch_cols ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from simplex_b200 import Basis, DeviceSimplex, synth

m, n, P = 4096, 16384, 1024
A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 1234)
s = DeviceSimplex(pricing="most_neg", engine="persistent", dual="update", iter_limit=P, chunk_iters=256, trace_capacity=P)
s.load(A, b, c)
ref = None
combos = [(0, 8, 8), (4, 8, 8), (8, 8, 8), (12, 8, 8), (16, 8, 8), (20, 8, 8), (24, 4, 8), (28, 0, 8), (12, 0, 8), (16, 0, 8), (16, 8, 16)]
if len(sys.argv) > 1:
    combos = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for combo in combos:
    keep, rows, cols = combo[:3]
    mode = combo[3] if len(combo) > 3 else 0
    os.environ["SB200_L2_KEEP_MODE"] = str(mode)
    os.environ["SB200_L2_KEEP_ROWS"] = str(keep)
    os.environ["SB200_PREFETCH_ROWS"] = str(rows)
    os.environ["SB200_PREFETCH_COLS"] = str(cols)
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
    us = 1e6 * best / r["pivots"]
    cyc = [float(pc[i]) / it for i in range(5)]
    ph = [round(v * us / sum(cyc), 1) for v in cyc]
    print("mode %d keep %2d prefetch rows %2d cols %2d: %.2f us/pivot  %.0f pivots/s  same_trace %s phases(price,bar1,fused,bar2,tail) %s" % (
        mode, keep, rows, cols, us, r["pivots"] / best, bool(np.array_equal(tr, ref)), ph), flush=True)
s.close()
