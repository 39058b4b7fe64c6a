#!/usr/bin/env python3
"""Developer probe: walks one tiny LP through the step API, then both engines, printing as it goes."""
import faulthandler
import sys

import numpy as np

faulthandler.dump_traceback_later(50, exit=True)
from simplex_b200 import Basis, DeviceSimplex, synth  # noqa: E402

m, n = 8, 16
A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 1234)
print("tableau ok", flush=True)
with DeviceSimplex(pricing="most_neg", engine="staged", dual="recompute", iter_limit=50) as s:
    s.load(A, b, c)
    print("load ok", flush=True)
    s.prepare(Basis(basic, nonbasic))
    print("prepare ok", s.xB()[:4], flush=True)
    for it in range(3):
        s.step_dual()
        pos, val = s.step_price()
        print("price", pos, val, flush=True)
        if pos < 0:
            break
        s.step_ftran(pos)
        print("ftran d", s.d()[:4], flush=True)
        row, theta = s.step_ratio(pos)
        print("ratio", row, theta, flush=True)
        s.step_update()
        print("update ok", flush=True)
for engine in sys.argv[1:] or ["staged", "persistent"]:
    for dual in ("recompute", "update"):
        with DeviceSimplex(pricing="most_neg", engine=engine, dual=dual, iter_limit=50, chunk_iters=8,
                           trace_capacity=64) as s:
            s.load(A, b, c)
            r = s.simplex(Basis(basic, nonbasic))
            print(engine, dual, r, flush=True)
