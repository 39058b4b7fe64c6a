import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplex_b200 import Basis, synth
from simplex_b200.dist import ShardedSimplex
from simplex_b200.solver import _dp
m, n = 256, 1024
A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 1234)
N = A.shape[1]
for lim in (1, 2, 3, 4, 6, 10, 20):
    s = ShardedSimplex(0, 1, pricing="most_neg", device=0, iter_limit=lim, chunk_iters=512, trace_capacity=64)
    s.load_shard(m, N, A, b, c); b2 = Basis(basic, nonbasic); s.simplex(b2)
    B2 = np.zeros((m, m)); s._check(s._L.sb200_get_Binv(s._ctx, _dp(B2)))
    y2 = np.zeros(m); s._check(s._L.sb200_get_y(s._ctx, _dp(y2)))
    tr = s.trace()[0]; s.close()
    Bt = np.linalg.inv(A[:, b2.basic]); yt = c[b2.basic] @ Bt
    bad = np.nonzero(np.abs(B2 - Bt).max(axis=1) > 1e-9)[0]
    print(lim, "bad rows", bad[:6], "y err", np.abs(y2 - yt).max(), "pivot rows", tr[:, 1].tolist()[-6:], flush=True)
