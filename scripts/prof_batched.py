"""Runs the batched kernel once on `count` LPs of the config-5 shape (for ncu captures)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simplex_b200 import DeviceSimplex, synth

count = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m, n = 64, 128
A, b, c, basic, nonbasic = synth.batch(count, m, n, 1234)
dA, db, dc = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(c).cuda()
s = DeviceSimplex(pricing="most_neg", stream=torch.cuda.current_stream().cuda_stream, iter_limit=100000)
for _ in range(reps):
    r = s.run_batched_device(count, m, m + n, dA, db, dc, torch.from_numpy(basic).cuda(), torch.from_numpy(nonbasic).cuda())
    its = r["iterations"].cpu().numpy()
    print("seconds %.6f lps/s %.0f pivots %d us/pivot/cta(2 per SM) %.3f" % (
        r["seconds"], count / r["seconds"], its.sum() - count, r["seconds"] * 296 / (its.sum() - count) * 1e6))
