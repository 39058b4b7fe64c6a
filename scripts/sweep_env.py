"""Headline LP (one load, many runs) under different environment settings of the library's tuning knobs.
usage: sweep_env.py "SB200_SPLIT_TAIL=0" "SB200_SPLIT_TAIL=32 SB200_L2_KEEP_ROWS=16" ...   [--m M --n N --pivots P]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from simplex_b200 import Basis, DeviceSimplex, synth

args = [a for a in sys.argv[1:] if not a.startswith("--")]
opt = {a.split("=")[0][2:]: int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--")}
m, n, P = opt.get("m", 4096), opt.get("n", 16384), opt.get("pivots", 1024)
A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 1234)
s = DeviceSimplex(pricing="most_neg", engine="persistent", dual="update", iter_limit=P, chunk_iters=256, trace_capacity=P)
s.load(A, b, c)
ref = None
touched = set()
for setting in args or [""]:
    for k in touched:
        os.environ.pop(k, None)
    for kv in setting.split():
        k, v = kv.split("=")
        os.environ[k] = v
        touched.add(k)
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
    print("%-60s %.2f us/pivot  %.0f pivots/s  same_trace %s  %s  phases(price,bar1,fused,bar2,tail) %s" % (
        setting or "(defaults)", us, r["pivots"] / best, bool(np.array_equal(tr, ref)), s.engine_used(), ph), flush=True)
s.close()
