#!/usr/bin/env python3
"""Sharded engine check/probe, one process per GPU:
    torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        scripts/sharded_check.py --m 1024 --n 8192 --iters 500
Every rank builds its own column block of G(m,n,0,seed), the ranks solve the LP together, rank 0
also solves it alone (single-GPU fused engine) and compares pivot traces, x_B and objective."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplex_b200 import Basis, DeviceSimplex, algorithmic_bytes_per_pivot, synth  # noqa: E402
from simplex_b200.dist import ShardedSimplex, shard_plan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", dest="m", type=int, default=1024)
ap.add_argument("--cols", dest="n", type=int, default=8192)
ap.add_argument("--seed", type=int, default=1234)
ap.add_argument("--iters", type=int, default=500)
ap.add_argument("--chunk", type=int, default=128)
ap.add_argument("--pricing", default="most_neg")
ap.add_argument("--no-compare", action="store_true")
ap.add_argument("--reps", type=int, default=2, help="runs of the solve (the fastest is reported); 1 for a long full solve")
ap.add_argument("--sweep", default="", help="comma list of ENV=VALUE[+ENV=VALUE] settings to time one after the other")
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")          # host plumbing only; the data path is NVLink peer memory

m, n = a.m, a.n
N = synth.num_cols(m, n, 0, False)
plan = shard_plan(m, N, world)
me = plan[rank]
t0 = time.time()
A_cols, b, c, basic, nonbasic = synth.tableau(m, n, 0, a.seed, col0=me["col0"], ncols=me["ncols"],
                                              col_stride=me["col_stride"])
tgen = time.time() - t0

s = ShardedSimplex(rank, world, pricing=a.pricing, device=local, iter_limit=a.iters, chunk_iters=a.chunk,
                   trace_capacity=a.iters)
s.load_shard(m, N, A_cols, b, c)
best = None
for rep in range(max(1, a.reps)):
    basis = Basis(basic, nonbasic)
    r = s.simplex(basis)
    if best is None or r["loop_seconds"] < best["loop_seconds"]:
        best = r
trace, _ = s.trace()
x = s.xB()
residual = s.primal_residual(b)          # max |b - A_B x_B| over the ranks: drift of the carried B^-1 / x_B
pc = s.phase_cycles()
out = {"rank": rank, "world": world, "term": best["term"], "pivots": best["pivots"], "objective": best["objective"],
       "us_per_pivot": 1e6 * best["loop_seconds"] / max(best["pivots"], 1), "gen_s": round(tgen, 2),
       "primal_residual_max_abs": residual, "b_max_abs": float(np.abs(b).max()), "min_xB": float(x.min())}
PHASES = ["price", "xchg1", "fused", "xchg2", "tail", "wait_cta_price", "wait_cta_fused"]


def phases_of(pc, us):
    # SM-cycle counters of CTA 0, scaled so that they add up to the event-timed pivot (the SM clock under load is not nominal)
    it = max(int(pc[7]), 1)
    cyc = [float(pc[i]) / it for i in range(7)]
    return {k: round(cyc[i] * us / max(sum(cyc), 1e-9), 1) for i, k in enumerate(PHASES)}, round(sum(cyc) / us, 0)


out["phases_us"], out["sm_mhz_implied"] = phases_of(pc, out["us_per_pivot"])
sweep = []
for setting in [x for x in a.sweep.split(",") if x]:
    for kv in setting.split("+"):
        k, v = kv.split("=")
        os.environ[k] = v
    best2 = None
    for rep in range(2):
        r2 = s.simplex(Basis(basic, nonbasic))
        if best2 is None or r2["loop_seconds"] < best2["loop_seconds"]:
            best2 = r2
    us2 = 1e6 * best2["loop_seconds"] / max(best2["pivots"], 1)
    ph2, _ = phases_of(s.phase_cycles(), us2)
    sweep.append({"setting": setting, "us_per_pivot": round(us2, 1), "phases_us": ph2, "objective": best2["objective"]})
if sweep:
    out["sweep"] = sweep
s.close()

ok = True
if rank == 0 and not a.no_compare:
    A, b1, c1, bas1, non1 = synth.tableau(m, n, 0, a.seed)
    with DeviceSimplex(pricing=a.pricing, engine="persistent", dual="update", device=local, iter_limit=a.iters,
                       chunk_iters=a.chunk, trace_capacity=a.iters) as one:
        one.load(A, b1, c1)
        b1s = Basis(bas1, non1)
        r1 = one.simplex(b1s)
        t1, _ = one.trace()
        x1 = one.xB()
    same_trace = np.array_equal(t1, trace)
    same_x = np.array_equal(x1, x)
    nmin = min(len(t1), len(trace))
    dif = np.nonzero((t1[:nmin] != trace[:nmin]).any(axis=1))[0]
    out["first_trace_diff"] = int(dif[0]) if dif.size else -1
    if dif.size:
        out["trace_at_diff"] = [t1[dif[0]].tolist(), trace[dif[0]].tolist()]
    out["max_abs_dx"] = float(np.abs(x1 - x).max()) if x1.shape == x.shape else None
    out["n_dx"] = int((x1 != x).sum()) if x1.shape == x.shape else None
    out.update({"single_gpu_us_per_pivot": 1e6 * r1["loop_seconds"] / max(r1["pivots"], 1), "same_trace": bool(same_trace),
                "same_x_bits": bool(same_x), "same_lists": bool(np.array_equal(b1s.basic, basis.basic)),
                "obj_diff": abs(r1["objective"] - best["objective"])})
    ok = same_trace and out["same_lists"] and out["obj_diff"] <= 1e-9 * max(1.0, abs(r1["objective"]))
if rank == 0:
    bytes_pp = algorithmic_bytes_per_pivot(m, n)
    out["algorithmic_TBs"] = bytes_pp / out["us_per_pivot"] / 1e6
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
