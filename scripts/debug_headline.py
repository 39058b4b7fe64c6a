import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from simplex_b200 import Basis, DeviceSimplex, synth
m, n = 4096, 16384
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 1234)
for pf in ("0", None):
    for dual in ("update", "recompute"):
        if pf is not None:
            os.environ["SB200_PREFETCH_COLS"] = pf; os.environ["SB200_PREFETCH_ROWS"] = pf
        else:
            os.environ.pop("SB200_PREFETCH_COLS", None); os.environ.pop("SB200_PREFETCH_ROWS", None)
        with DeviceSimplex(pricing="most_neg", engine="persistent", dual=dual, iter_limit=P, chunk_iters=250, trace_capacity=P) as s:
            s.load(A, b, c)
            basis = Basis(basic, nonbasic)
            r = s.simplex(basis)
            x = s.xB(); Binv = s.Binv(); res = s.primal_residual()
            AB = A[:, basis.basic]
            E = Binv @ AB - np.eye(m)
            rowerr = np.abs(E).max(axis=1)
            bad = np.nonzero(rowerr > 1e-9)[0]
            xb = Binv @ b
            dx = np.abs(x - xb)
            print("prefetch", pf, "dual", dual, "pivots", r["pivots"], "residual", res, "max|BinvAB-I|", rowerr.max(), "bad rows", bad[:10], len(bad),
                  "max|x-Binv b|", dx.max(), "argmax", int(dx.argmax()), "rel", (dx / np.maximum(1, np.abs(x))).max(), flush=True)
