#!/usr/bin/env python3
"""One small run of one loop kernel, for compute-sanitizer (scripts/sanitize.sh). No torch: numpy + ctypes only.
usage: sanitize_case.py staged|persistent|fused|fused_bounded|resident|resident_bounded|sharded1|sharded2|batched_fast|batched_generic"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplex_b200 import Basis, DeviceSimplex, synth  # noqa: E402

case = sys.argv[1]
m, n = 40, 72
A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 7)
ub = None
kw = {"pricing": "most_neg", "chunk_iters": 9, "trace_capacity": 256, "hygiene_period": 2}
want = {"staged": "staged", "persistent": "persistent", "fused": "fused", "fused_bounded": "fused", "resident": "resident",
        "resident_bounded": "resident"}
if case in want:
    if case == "staged":
        kw.update(engine="staged", dual="update")
    elif case == "persistent":
        kw.update(engine="persistent", dual="recompute")
    elif case in ("resident", "resident_bounded"):
        kw.update(engine="persistent", dual="update")
    else:
        kw.update(engine="streaming", dual="update")
    if case in ("fused_bounded", "resident_bounded"):
        rng = np.random.default_rng(3)
        ub = np.full(A.shape[1], np.finfo(np.float64).max)
        ub[:n] = np.where(rng.random(n) < 0.5, rng.uniform(0.01, 0.3, n), 20.0)
    with DeviceSimplex(**kw) as s:
        s.load(A, b, c, ub)
        basis = Basis(basic, nonbasic)
        r = s.simplex(basis)
        assert s.engine_used() == want[case], s.engine_used()
        assert r["term"] == "optimal", r
        print(case, r["iterations"], r["pivots"], r["flips"], r["objective"], s.hygiene_stats())
elif case in ("sharded1", "sharded2"):
    from simplex_b200.dist import LocalShardedSimplex
    world = 1 if case == "sharded1" else 2
    with LocalShardedSimplex(world, pricing="most_neg", chunk_iters=9, trace_capacity=256) as s:
        s.load(A, b, c)
        basis = Basis(basic, nonbasic)
        r = s.simplex(basis)
        assert r["term"] == "optimal", r
        print(case, r["iterations"], r["pivots"], r["objective"])
else:
    if case == "batched_generic":
        os.environ["SB200_BATCHED_GENERIC"] = "1"
    Ab, bb, cb, bas, non = synth.batch(12, 24, 40, 900)
    with DeviceSimplex(pricing="most_neg") as s:
        r = s.run_batched(Ab, bb, cb, bas, non)
        assert (r["term"] == 1).all(), r["term"]
        print(case, r["iterations"].tolist())
