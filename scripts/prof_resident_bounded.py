#!/usr/bin/env python3
"""One 256-iteration launch of k_resident_bounded on a bounded LP at the two-phase shape (m=1024, n=2048), after a
warm-up run; used under `ncu -k regex:k_resident_bounded` (profiles/r2_ncu_full_k_resident_bounded.txt)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplex_b200 import Basis, DeviceSimplex, synth  # noqa: E402

m, n, its = 1024, 2048, 256
A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 1234)
rng = np.random.default_rng(1234)
ub = np.full(A.shape[1], np.finfo(np.float64).max)
ub[:n] = np.where(rng.random(n) < 0.4, rng.uniform(0.01, 0.3, n), rng.uniform(1.0, 50.0, n))
with DeviceSimplex(pricing="most_neg", dual="update", iter_limit=its, chunk_iters=its, hygiene_period=0) as s:
    for _ in range(2):
        s.load(A, b, c, ub)
        r = s.simplex(Basis(basic, nonbasic))
    assert s.engine_used() == "resident", s.engine_used()
    print("k_resident_bounded: %d iterations (%d pivots, %d flips), %.2f us per iteration"
          % (r["iterations"], r["pivots"], r["flips"], 1e6 * r["loop_seconds"] / r["iterations"]))
