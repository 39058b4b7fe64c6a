"""Host-side logic of the multi-GPU path, on CPU: world_size-2 gloo processes for the plumbing
(handle exchange, diagonal sum, row gather) and the partition / arg-min combine rules."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_plan_covers_everything_once():
    from simplex_b200.dist import owner_of_column, owner_of_row, shard_plan
    for m, N, world in [(8192, 270336, 8), (4096, 20480, 2), (1000, 3001, 4), (64, 192, 8), (33, 70, 3), (5, 9, 1)]:
        plan = shard_plan(m, N, world)
        assert sum(p["ncols"] for p in plan) == N and sum(p["nrows"] for p in plan) == m
        seen = []
        for p in plan:
            cols = list(range(p["col0"], N, p["col_stride"]))
            assert len(cols) == p["ncols"]            # cyclic: {rank + k*world}
            seen += cols
            assert p["rows_per_rank"] % 32 == 0      # dual-solve blocks never straddle ranks
            for j in cols[:2] + cols[-2:]:
                assert owner_of_column(j, N, world) == p["rank"]
        assert sorted(seen) == list(range(N))
        for p in plan:
            for i in (p["row0"], p["row0"] + p["nrows"] - 1):
                if p["nrows"]:
                    assert owner_of_row(i, m, world) == p["rank"]


def test_combine_candidates_is_the_reference_tie_break():
    from simplex_b200.dist import combine_candidates
    # pricing: most negative value, ties to the LOWEST position whatever the rank order
    assert combine_candidates([(-2.0, 0, 900), (-2.0, 0, 17), (-1.0, 0, 3)]) == (-2.0, 0, 17)
    assert combine_candidates([(0.0, 0, -1), (-0.5, 0, 40)]) == (-0.5, 0, 40)
    assert combine_candidates([(0.0, 0, -1), (0.0, 0, -1)])[2] == -1
    # ratio: a t1 row (cls 0) beats a t2 row (cls 1) of equal ratio even with a higher index
    assert combine_candidates([(0.25, 1, 2), (0.25, 0, 9)]) == (0.25, 0, 9)
    # associativity/commutativity: any grouping over ranks gives the same winner
    rng = np.random.default_rng(1)
    c = [(float(rng.integers(0, 4)), int(rng.integers(0, 2)), int(rng.integers(0, 50))) for _ in range(64)]
    whole = combine_candidates(c)
    halves = combine_candidates([combine_candidates(c[:20]), combine_candidates(c[20:])])
    assert whole == halves == combine_candidates(list(reversed(c)))


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from simplex_b200.dist import exchange_handles, gather_rows, shard_plan, sum_over_ranks, worst_rc
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 64-byte window handles travel rank-ordered
    mine = bytes([rank * 16 + (k %% 16) for k in range(64)])
    allh = exchange_handles(mine)
    assert len(allh) == 64 * world
    for r in range(world):
        assert allh[64 * r:64 * (r + 1)] == bytes([r * 16 + (k %% 16) for k in range(64)])
    # crash-basis diagonal: every entry is owned by exactly one rank
    m = 100
    diag = np.zeros(m)
    diag[rank::world] = 1.0 + np.arange(m)[rank::world]
    tot = sum_over_ranks(diag)
    assert np.array_equal(tot, 1.0 + np.arange(m))
    # row-sharded x_B comes back in row order
    plan = shard_plan(m, 300, world)
    me = plan[rank]
    local = np.arange(me["row0"], me["row0"] + me["nrows"], dtype=np.float64)
    full = gather_rows(local, plan)
    assert np.array_equal(full, np.arange(m, dtype=np.float64))
    # a return code that is bad on ONE rank is bad on every rank (nobody is left waiting in the next collective)
    assert worst_rc(0) == 0
    assert worst_rc(7 if rank == 1 else 0) == 7
    dist.barrier()
    dist.destroy_process_group()
    print("ok", rank)
""")


def test_gloo_world2_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env,
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
