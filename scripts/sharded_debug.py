import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplex_b200 import Basis, DeviceSimplex, synth
from simplex_b200.dist import ShardedSimplex
import ctypes as C
from simplex_b200.solver import _dp
m, n = 256, 1024
A, b, c, basic, nonbasic = synth.tableau(m, n, 0, 1234)
N = A.shape[1]
for lim, chunk in ((64, 64), (65, 64), (65, 512), (130, 64)):
    with DeviceSimplex(pricing="most_neg", engine="persistent", dual="update", iter_limit=lim, chunk_iters=chunk, trace_capacity=512) as one:
        one.load(A, b, c); b1 = Basis(basic, nonbasic); r1 = one.simplex(b1)
        y1, B1, x1, t1 = one.y(), one.Binv(), one.xB(), one.trace()[0]
    s = ShardedSimplex(0, 1, pricing="most_neg", device=0, iter_limit=lim, chunk_iters=chunk, trace_capacity=512)
    s.load_shard(m, N, A, b, c); b2 = Basis(basic, nonbasic); r2 = s.simplex(b2)
    y2 = np.zeros(m); s._check(s._L.sb200_get_y(s._ctx, _dp(y2)))
    B2 = np.zeros((m, m)); s._check(s._L.sb200_get_Binv(s._ctx, _dp(B2)))
    x2 = s.xB(); t2 = s.trace()[0]
    s.close()
    print(lim, chunk, "piv", r1["pivots"], r2["pivots"], "trace", np.array_equal(t1, t2), "y", np.abs(y1-y2).max(), "B", np.abs(B1-B2).max(),
          "x", np.abs(x1-x2).max(), "lists", np.array_equal(b1.basic, b2.basic), np.array_equal(b1.non_basic, b2.non_basic), flush=True)
    if np.abs(x1-x2).max() > 0:
        idx = np.nonzero(x1 != x2)[0]; print("   x diff rows", idx[:8], x1[idx[:4]], x2[idx[:4]], "last pivot", t1[-1], t2[-1])
    if np.abs(B1-B2).max() > 0:
        rows = np.nonzero((B1 != B2).any(axis=1))[0]; print("   B diff rows", rows[:10], len(rows))
print("---- invariants after 64 iterations, chunk 64")
lim, chunk = 64, 64
with DeviceSimplex(pricing="most_neg", engine="persistent", dual="update", iter_limit=lim, chunk_iters=chunk) as one:
    one.load(A, b, c); b1 = Basis(basic, nonbasic); one.simplex(b1); B1 = one.Binv(); y1 = one.y()
s = ShardedSimplex(0, 1, pricing="most_neg", device=0, iter_limit=lim, chunk_iters=chunk)
s.load_shard(m, N, A, b, c); b2 = Basis(basic, nonbasic); s.simplex(b2)
B2 = np.zeros((m, m)); s._check(s._L.sb200_get_Binv(s._ctx, _dp(B2))); y2 = np.zeros(m); s._check(s._L.sb200_get_y(s._ctx, _dp(y2))); s.close()
AB = A[:, b1.basic]
Bt = np.linalg.inv(AB); yt = c[b1.basic] @ Bt
print("single  |B-inv|", np.abs(B1 - Bt).max(), "row0", np.abs(B1[0] - Bt[0]).max(), "|y-yt|", np.abs(y1 - yt).max())
print("sharded |B-inv|", np.abs(B2 - Bt).max(), "row0", np.abs(B2[0] - Bt[0]).max(), "|y-yt|", np.abs(y2 - yt).max())
print("basic[0]", b1.basic[0], "cols differing in row0:", np.nonzero(B1[0] != B2[0])[0][:10])
