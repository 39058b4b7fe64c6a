#!/usr/bin/env python3
"""us per pivot of the default engine choice against the fused kernel at shapes around the eligibility limits of
k_resident_stream (rows of B^-1 on chip, columns priced out of L2). usage: stream_cols_sweep.py [m,n ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplex_b200 import Basis, DeviceSimplex, synth  # noqa: E402

shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(1024, 4096), (1200, 2400), (1776, 2048),
                                                                        (512, 8192), (256, 20000), (64, 60000)]
for m, n in shapes:
    A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 5)
    row = {}
    for label, engine, env in (("default", "persistent", "2"), ("fused", "streaming", "0")):
        os.environ["SB200_RES_STREAM_COLS"] = env
        with DeviceSimplex(pricing="most_neg", engine=engine, dual="update", iter_limit=800, chunk_iters=400) as s:
            for _ in range(2):
                s.load(A, b, c)
                r = s.simplex(Basis(basic, nonbasic))
            row[label] = (s.engine_used(), 1e6 * r["loop_seconds"] / max(1, r["iterations"]), r["iterations"], r["objective"])
    print("m=%d n=%d A=%.0f MB: stream-forced %s %.2f us/iter | fused %.2f us/iter | same objective %s"
          % (m, n, A.nbytes / 2 ** 20, row["default"][0], row["default"][1], row["fused"][1],
             row["default"][3] == row["fused"][3]), flush=True)
