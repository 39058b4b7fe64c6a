"""Parity of the CUDA path with the UNMODIFIED reference at the sizes BASELINE.json quotes
(tests/golden/large.json.gz, written by tests/golden/make_golden_large.py from runs of
oracle/_ref/ref_driver_*: /root/reference/src/Simplex.cpp:255-353 compiled where it lies).

north_star: "identical pivot sequences and objectives to the reference within 1e-9" on
configs[1] (m=1024, n=2048, n_eq=256, two-phase) and configs[2] (m=4096, n=16384, the headline).
Everything here needs a GPU and goes through the C ABI."""
import numpy as np
import pytest

from helpers import rel_close
from test_gpu_parity import ENGINE_DUAL, _device_solver, _two_phase_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("engine,dual", ENGINE_DUAL)
def test_config1_full_run_matches_reference(golden_large, engine, dual):
    """BASELINE configs[1], G(1024, 2048, 256, 1234), most-negative pricing: Phase I (727 pivots) and Phase II
    (6059 pivots) walk the reference's pivot sequence pivot by pivot on every engine (staged graph, 4-phase
    persistent kernel, shared-memory-resident kernel, HBM-streaming fused kernel); final lists identical,
    objective and every structural x within 1e-9 relative."""
    g = golden_large["config1"]
    res = _two_phase_gpu(g, engine, dual, resident=True)
    assert len(res) == len(g["calls"]) == 2
    for k, (r, call) in enumerate(zip(res, g["calls"])):
        assert r["term"] == "optimal", k
        assert r["iterations"] == call["iterations"], k
        assert r["trace"].tolist() == call["pivots"], k
        assert r["basic"].tolist() == call["basic_out"], k
        assert r["nonbasic"].tolist() == call["nonbasic_out"], k
    last = res[-1]
    assert rel_close(last["objective"], g["objective"])
    x = np.zeros(g["n"])
    for i, col in enumerate(last["basic"]):
        if col < g["n"]:
            x[col] = last["x"][i]
    assert rel_close(x, g["x_struct"])


def test_config1_host_flow_matches_reference(golden_large):
    """The same LP through the host C++ flow (simplex::solve_dense: standard form, scaling, crash basis,
    Phase I, hand-off with A resident, Phase II, read-out): the reference's iteration counts, objective and x."""
    from simplex_b200 import lp as hostlp
    from simplex_b200 import synth
    g = golden_large["config1"]
    cr, ar, rhsr, types = synth.raw(g["m"], g["n"], g["n_eq"], g["seed"])
    hd = hostlp.solve_dense(ar, types, rhsr, cr, pricing="most_neg")
    assert hd["status"] == "optimal"
    assert [k["iterations"] for k in hd["calls"]] == [c["iterations"] for c in g["calls"]]
    assert [k["pivots"] for k in hd["calls"]] == [len(c["pivots"]) for c in g["calls"]]
    assert rel_close(hd["objective"], g["objective"])
    assert rel_close(hd["x_scaled"], g["x_struct"])


def _prefix_run(g, P, **kw):
    from simplex_b200 import Basis, synth
    A, b, c, basic, nonbasic = synth.tableau(g["m"], g["n"], 0, g["seed"])
    with _device_solver(pricing=g["rule"], iter_limit=P, trace_capacity=max(P, 1), **kw) as s:
        s.load(A, b, c)
        basis = Basis(basic, nonbasic)
        r = s.simplex(basis)
        tr, _ = s.trace()
        return r, tr.tolist(), basis.basic.tolist(), s.engine_used()


@pytest.mark.parametrize("engine,dual,kernel", [("persistent", "update", "fused"), ("persistent", "recompute", "persistent"),
                                                ("staged", "recompute", "staged")])
def test_headline_prefix_matches_reference(golden_large, engine, dual, kernel):
    """BASELINE configs[2], the headline roofline case G(4096, 16384, 0, 1234): the first recorded pivots of the
    reference (ref_driver seam mode: Simplex::simplex on the dense tableau, a fresh LU every iteration) are the
    pivots the GPU engines take, entry by entry (leaving column, row, entering column, position)."""
    g = golden_large["headline_prefix"]
    P = len(g["pivots"])
    assert P >= 64
    r, trace, basic, used = _prefix_run(g, P, engine=engine, dual=dual, chunk_iters=256)
    assert used == kernel
    assert r["pivots"] == P and r["term"] == "iter_limit"
    assert trace == g["pivots"]
    assert basic == g["basic_after"]


@pytest.mark.parametrize("name", ["mid_prefix", "first_eligible", "sharded_check"])
@pytest.mark.parametrize("engine,dual", ENGINE_DUAL)
def test_large_prefixes_match_reference(golden_large, name, engine, dual):
    """A mid-size LP (m=2048, n=4096: tableau in L2, not in shared memory), the stock first-eligible rule at
    m=1024 and the wide shape of the multi-rank checks (m=1024, n=8192, complete run): reference traces."""
    if name not in golden_large:
        pytest.skip("%s not generated yet" % name)
    g = golden_large[name]
    P = len(g["pivots"])
    r, trace, basic, _ = _prefix_run(g, P if not g["complete"] else 10 ** 9, engine=engine, dual=dual)
    if g["complete"]:
        assert r["term"] == "optimal" and r["iterations"] == g["iterations"]
    assert trace == g["pivots"]
    assert basic == g["basic_after"]


def test_wide_shape_prefix_matches_reference(golden_large):
    """BASELINE configs[3] shape, G(8192, 262144, 0, 1234) (17.7 GB of A): the reference's first iterations
    (35 GB of host memory and about a minute each on the CPU) against the fused engine on ONE GPU. The sharded
    engine is checked against the same vectors by bench.py --gpus N (the test box has one GPU)."""
    if "wide_prefix" not in golden_large:
        pytest.skip("wide_prefix not generated yet")
    import torch
    if torch.cuda.mem_get_info()[1] < 40 * (1 << 30):
        pytest.skip("needs 40 GB of device memory")
    g = golden_large["wide_prefix"]
    P = len(g["pivots"])
    r, trace, basic, used = _prefix_run(g, P, engine="persistent", dual="update", chunk_iters=P)
    assert used == "fused" and r["pivots"] == P
    assert trace == g["pivots"]
    assert basic == g["basic_after"]


def test_sharded_kernel_single_rank_on_reference_trace(golden_large):
    """The sharded kernel (world = 1: every exchange goes through its own window) on the complete reference
    trace of G(1024, 8192, 0, 1234): 5598 pivots to the optimum."""
    from simplex_b200 import Basis, synth
    from simplex_b200.dist import ShardedSimplex
    g = golden_large["sharded_check"]
    A, b, c, basic, nonbasic = synth.tableau(g["m"], g["n"], 0, g["seed"])
    with ShardedSimplex(0, 1, pricing=g["rule"], device=0, trace_capacity=1 << 16, chunk_iters=512,
                        iter_limit=10 ** 6) as s:
        s.load_shard(g["m"], A.shape[1], A, b, c)
        basis = Basis(basic, nonbasic)
        r = s.simplex(basis)
        tr, _ = s.trace()
    assert r["term"] == "optimal" and r["iterations"] == g["iterations"]
    assert tr.tolist() == g["pivots"] and basis.basic.tolist() == g["basic_after"]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_ranks_sharing_one_gpu_walk_the_reference_trace(golden_large, world):
    """G-independence where the driver can see it: `world` ranks of the sharded engine in one process, sharing
    ONE GPU (each rank's cooperative grid on its share of the SMs, windows wired directly, every per-pivot
    exchange through them exactly as over NVLink). Cyclic column shards, row-sharded B^-1: the ranks walk the
    reference's pivot sequence on G(1024, 8192, 0, 1234), and x_B has the bits the single-GPU fused engine
    computes (every dot product is evaluated by one fixed summation order, ties break on indices)."""
    from simplex_b200 import Basis, DeviceSimplex, synth
    from simplex_b200.dist import LocalShardedSimplex
    g = golden_large["sharded_check"]
    P = 1000
    A, b, c, basic, nonbasic = synth.tableau(g["m"], g["n"], 0, g["seed"])
    with LocalShardedSimplex(world, pricing=g["rule"], iter_limit=P, trace_capacity=P, chunk_iters=200) as s:
        s.load(A, b, c)
        basis = Basis(basic, nonbasic)
        r = s.simplex(basis)
        tr, _ = s.trace()
        x = s.xB()
    assert r["term"] == "iter_limit" and r["pivots"] == P
    assert tr.tolist() == g["pivots"][:P]
    with DeviceSimplex(pricing=g["rule"], engine="streaming", dual="update", iter_limit=P, chunk_iters=200) as one:
        one.load(A, b, c)
        b1 = Basis(basic, nonbasic)
        r1 = one.simplex(b1)
        x1 = one.xB()
    assert basis.basic.tolist() == b1.basic.tolist() and basis.non_basic.tolist() == b1.non_basic.tolist()
    assert np.array_equal(x, x1)
    assert rel_close(r["objective"], r1["objective"])


def test_sharded_ranks_sharing_one_gpu_run_to_the_optimum(golden_synth):
    """Two and three ranks on one GPU, complete runs (optimal ending decided by every rank at the same iteration),
    both pricing rules, ragged shapes (rows not a multiple of the row-block size, more ranks than some shards
    have columns of a group)."""
    from simplex_b200 import Basis, synth
    from simplex_b200.dist import LocalShardedSimplex
    n_checked = 0
    for rec in golden_synth:
        if rec["n_eq"] != 0 or rec["m"] > 256:
            continue
        A, b, c, basic, nonbasic = synth.tableau(rec["m"], rec["n"], 0, rec["seed"])
        call = rec["calls"][-1]
        for world in (2, 3):
            with LocalShardedSimplex(world, pricing=rec["rule"], trace_capacity=1 << 14, chunk_iters=37) as s:
                s.load(A, b, c)
                basis = Basis(basic, nonbasic)
                r = s.simplex(basis)
                tr, _ = s.trace()
            assert r["term"] == "optimal" and r["iterations"] == call["iterations"], (rec["m"], rec["rule"], world)
            assert tr.tolist() == call["pivots"] and basis.basic.tolist() == call["basic_out"]
            assert rel_close(r["objective"], rec["objective"])
        n_checked += 1
    assert n_checked >= 6


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_two_phase_matches_reference(golden_large, world):
    """BASELINE configs[1] on the SHARDED engine (ranks sharing one GPU): Phase I from the slack/artificial crash
    basis, hand-off as the reference does it (artificial ids deleted from the lists, src/InitialBasis.cpp:85-121;
    costs replaced, trailing columns dropped: sb200_shard_set_costs), Phase II from the carried-over row-sharded
    B^-1 (sb200_shard_continue). Both calls walk the reference's pivot sequence; objective, x and the primal
    residual of the final basis as on one GPU."""
    from simplex_b200 import Basis, synth
    from simplex_b200.dist import LocalShardedSimplex
    g = golden_large["config1"]
    m, n, n_eq, seed = g["m"], g["n"], g["n_eq"], g["seed"]
    A1, b1, c1, basic, nonbasic = synth.tableau(m, n, n_eq, seed, phase1=True)
    _, _, c2, _, _ = synth.tableau(m, n, n_eq, seed, phase1=False)
    with LocalShardedSimplex(world, pricing=g["rule"], trace_capacity=1 << 14, chunk_iters=256) as s:
        s.load(A1, b1, c1)
        basis = Basis(basic, nonbasic)
        r1 = s.simplex(basis)
        t1, _ = s.trace()
        assert r1["term"] == "optimal" and r1["iterations"] == g["calls"][0]["iterations"]
        assert t1.tolist() == g["calls"][0]["pivots"]
        assert basis.basic.tolist() == g["calls"][0]["basic_out"] and basis.non_basic.tolist() == g["calls"][0]["nonbasic_out"]
        bas2, non2 = synth.phase2_lists(basis.basic, basis.non_basic, n + (m - n_eq))
        assert bas2.tolist() == g["calls"][1]["basic_in"] and non2.tolist() == g["calls"][1]["nonbasic_in"]
        s.set_costs(c2)
        basis2 = Basis(bas2, non2)
        r2 = s.simplex(basis2, carry_over=True)
        t2, _ = s.trace()
        x = s.xB()
        res = s.primal_residual(b1)
    assert r2["term"] == "optimal" and r2["iterations"] == g["calls"][1]["iterations"]
    assert t2.tolist() == g["calls"][1]["pivots"]
    assert basis2.basic.tolist() == g["calls"][1]["basic_out"] and basis2.non_basic.tolist() == g["calls"][1]["nonbasic_out"]
    assert rel_close(r2["objective"], g["objective"])
    xs = np.zeros(n)
    for i, col in enumerate(basis2.basic):
        if col < n:
            xs[col] = x[i]
    assert rel_close(xs, g["x_struct"])
    assert res < 1e-9 * max(1.0, np.abs(b1).max())


@pytest.mark.parametrize("carry", [True, False])
def test_phase_two_carries_the_inverse_over(golden_large, carry, monkeypatch):
    """configs[1]: Phase II starts from the basis Phase I ended on, with A resident (sb200_set_costs). By default the
    device keeps the B^-1 and x_B that Phase I left (sb200_inverse_carried); with SB200_NO_CARRY=1 the basis is
    inverted afresh (Gauss-Jordan), as the reference factorises it afresh. Both walk the reference's Phase-II trace."""
    from simplex_b200 import Basis, synth
    if not carry:
        monkeypatch.setenv("SB200_NO_CARRY", "1")
    g = golden_large["config1"]
    m, n, n_eq, seed = g["m"], g["n"], g["n_eq"], g["seed"]
    A1, b1, c1, basic, nonbasic = synth.tableau(m, n, n_eq, seed, phase1=True)
    _, _, c2, _, _ = synth.tableau(m, n, n_eq, seed, phase1=False)
    with _device_solver(pricing=g["rule"], engine="persistent", dual="update", trace_capacity=1 << 14) as s:
        s.load(A1, b1, c1)
        basis = Basis(basic, nonbasic)
        s.simplex(basis)
        assert not s.inverse_carried()
        bas2, non2 = synth.phase2_lists(basis.basic, basis.non_basic, n + (m - n_eq))
        s.set_costs(c2)
        basis2 = Basis(bas2, non2)
        r2 = s.simplex(basis2)
        assert s.inverse_carried() == carry
        tr, _ = s.trace()
        assert r2["term"] == "optimal" and r2["iterations"] == g["calls"][1]["iterations"]
        assert tr.tolist() == g["calls"][1]["pivots"]
        assert rel_close(r2["objective"], g["objective"])
        assert s.primal_residual() < 1e-9 and s.inverse_residual(16, 5) < 1e-9


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_general_start_basis_and_reinversion(world):
    """Any start basis on the sharded engine: the basis a single-GPU run has reached after 150 pivots (no unit
    basis) is handed to `world` ranks, every rank inverts the gathered basis matrix with the single-GPU Gauss-Jordan
    and keeps its rows; the sharded run then continues exactly as the single-GPU engine continues from the same
    general basis (same trace, bit-identical x_B). A re-inversion between two sharded runs rebuilds B^-1 and x_B of
    the current basis (tiny primal residual) and the solve goes on to the same optimum."""
    from simplex_b200 import Basis, DeviceSimplex, synth
    from simplex_b200.dist import LocalShardedSimplex
    A, b, c, basic, nonbasic = synth.tableau(300, 640, 0, 77)
    with DeviceSimplex(pricing="most_neg", engine="streaming", dual="update", iter_limit=150) as s:
        s.load(A, b, c)
        mid = Basis(basic, nonbasic)
        assert s.simplex(mid)["pivots"] == 150
    with DeviceSimplex(pricing="most_neg", engine="streaming", dual="update", trace_capacity=1 << 14) as s:
        s.load(A, b, c)
        b1 = mid.copy()
        r1 = s.simplex(b1)                       # from the general basis: Gauss-Jordan on the device
        t1, _ = s.trace()
        x1 = s.xB()
    assert r1["term"] == "optimal" and r1["pivots"] > 50
    with LocalShardedSimplex(world, pricing="most_neg", trace_capacity=1 << 14, chunk_iters=64, iter_limit=10 ** 6) as sh:
        sh.load(A, b, c)
        b2 = mid.copy()
        r2 = sh.simplex(b2, general=True)
        t2, _ = sh.trace()
        x2 = sh.xB()
        assert (r2["term"], r2["iterations"]) == (r1["term"], r1["iterations"])
        assert t2.tolist() == t1.tolist() and b2.basic.tolist() == b1.basic.tolist()
        assert np.array_equal(x2, x1)
        sh.reinvert()                            # B^-1, x_B of the final basis from scratch
        assert sh.primal_residual(b) < 1e-10
        assert rel_close(sh.xB(), x1, tol=1e-10)
        r3 = sh.simplex(b2, carry_over=True)     # still optimal: one iteration, no pivot
        assert (r3["term"], r3["iterations"], r3["pivots"]) == ("optimal", 1, 0)


@pytest.mark.parametrize("world", [1, 3])
def test_sharded_periodic_reinversion_and_hygiene(golden_synth, world):
    """Re-inversion on the sharded engine (the run stops every P iterations, every rank rebuilds B^-1 and x_B of the
    current basis from the gathered basis matrix, the run resumes with the same counters and exchange sequence
    numbers): the reference's pivot sequence, iteration count and optimum, with re-inversions every 25 iterations; and
    the residual-triggered variant (a check every 40 iterations, tolerance met: no re-inversion; tolerance nothing can
    meet: one per check)."""
    from simplex_b200 import Basis, synth
    from simplex_b200.dist import LocalShardedSimplex
    rec = [r for r in golden_synth if (r["m"], r["n"], r["n_eq"], r["rule"]) == (128, 256, 0, "dantzig")][0]
    A, b, c, basic, nonbasic = synth.tableau(128, 256, 0, rec["seed"])
    call = rec["calls"][-1]
    for kw, want_refresh in ((dict(reinvert_period=25), True), (dict(hygiene_tol=1e-9, hygiene_period=40, b=b), False),
                             (dict(hygiene_tol=1e-300, hygiene_period=40, b=b), True)):
        with LocalShardedSimplex(world, pricing="most_neg", trace_capacity=1 << 14, chunk_iters=16, iter_limit=10 ** 6) as s:
            s.load(A, b, c)
            basis = Basis(basic, nonbasic)
            r = s.simplex(basis, **kw)
            tr, _ = s.trace()
            assert r["term"] == "optimal" and r["iterations"] == call["iterations"], kw
            assert tr.tolist() == call["pivots"] and basis.basic.tolist() == call["basic_out"], kw
            assert rel_close(r["objective"], rec["objective"])
            assert (s.hygiene["refreshes"] > 0) == want_refresh, (kw, s.hygiene)
            if "hygiene_tol" in kw:
                assert s.hygiene["checks"] >= 1 and (want_refresh or s.hygiene["worst"] < 1e-10)
            assert s.primal_residual(b) < 1e-9


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("rule", ["most_neg", "first"])
def test_sharded_bounded_matches_oracle(world, rule):
    """Finite upper bounds on the sharded engine (sb200_shard_set_bounds; ranks sharing one GPU): t2 candidates, flips of
    the entering variable, substitution of leaving variables (their column pushed by its owner so that every rank's
    replica of b follows), sign flags until the exit of a launch (a launch every 7 iterations), against the LU oracle:
    iterations, flips, pivot trace, flags, lists, b, x_B, objective."""
    from simplex_b200 import Basis
    from simplex_b200.dist import LocalShardedSimplex
    from test_gpu_parity import _bounded_synth
    from oracle import oracle
    for m, n, tight in ((64, 128, 0.3), (150, 90, 0.4), (96, 200, 0.6)):
        A, b, c, ub, basic, nonbasic = _bounded_synth(m, n, 31 + m, tight)
        want = oracle.simplex(A, b, c, basic, nonbasic, ub=ub, rule=oracle.RULES[rule], variant="lu")
        assert want["term"] == "optimal" and want["flips"] > 0
        with LocalShardedSimplex(world, pricing=rule, trace_capacity=1 << 14, chunk_iters=7, iter_limit=10 ** 6) as s:
            s.load(A, b, c, ub=ub)
            basis = Basis(basic, nonbasic)
            r = s.simplex(basis)
            tr, _ = s.trace()
            x = s.xB()
            assert (r["term"], r["iterations"], r["flips"]) == (want["term"], want["iterations"], want["flips"]), (m, n)
            assert tr.tolist() == want["trace"].tolist(), (m, n)
            assert basis.basic.tolist() == want["basic"].tolist() and basis.non_basic.tolist() == want["nonbasic"].tolist()
            for rank in range(world):
                assert s.flip_flags(rank).tolist() == want["flip"].tolist(), (m, n, rank)
                assert rel_close(s.b(rank), want["b"]), (m, n, rank)
            assert rel_close(x, want["x"]) and rel_close(r["objective"], want["objective"]), (m, n)
            assert s.primal_residual(want["b"]) < 1e-9
