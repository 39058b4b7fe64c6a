#!/usr/bin/env python3
"""Developer probe (not a bench): pivots/s of the engines on the BASELINE shapes, bounded by an
iteration limit so that one call stays short."""
import argparse
import time

import numpy as np

from simplex_b200 import Basis, DeviceSimplex, algorithmic_bytes_per_pivot, synth

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=4096)
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--iters", type=int, default=600)
ap.add_argument("--chunk", type=int, default=64)
ap.add_argument("--combos", default="staged:recompute,staged:update,persistent:recompute,persistent:update")
a = ap.parse_args()

t = time.time()
A, b, c, basic, nonbasic = synth.tableau(a.m, a.n, 0, 1234)
print("generated %dx%d in %.1fs" % (A.shape[0], A.shape[1], time.time() - t), flush=True)
bytes_pp = algorithmic_bytes_per_pivot(a.m, a.n)
ref_trace = None
for combo in a.combos.split(","):
    engine, dual = combo.split(":")
    with DeviceSimplex(pricing="most_neg", engine=engine, dual=dual, iter_limit=a.iters, chunk_iters=a.chunk,
                       trace_capacity=a.iters) as s:
        t = time.time()
        s.load(A, b, c)
        tl = time.time() - t
        best = None
        for rep in range(3):
            r = s.simplex(Basis(basic, nonbasic))
            if best is None or r["loop_seconds"] < best["loop_seconds"]:
                best = r
        tr, _ = s.trace()
        if engine == "persistent":
            pc = s.phase_cycles()
            it = max(int(pc[7]), 1)
            names = ["price", "bar1", "ftran", "bar2", "update", "bar3", "dual"]
            print("   phases us/iter @1.965GHz: " + "  ".join("%s=%.1f" % (nm, pc[i] / it / 1965.0)
                                                               for i, nm in enumerate(names)), flush=True)
        if ref_trace is None:
            ref_trace = tr
        same = np.array_equal(tr, ref_trace)
        us = best["loop_seconds"] / max(best["pivots"], 1) * 1e6
        print("%-10s %-9s term=%s pivots=%d  %.1f us/pivot  %.0f pivots/s  %.2f TB/s algorithmic  obj=%.9f  "
              "load=%.2fs same_trace=%s" % (engine, dual, best["term"], best["pivots"], us, 1e6 / us,
                                            bytes_pp / us / 1e6, best["objective"], tl, same), flush=True)
