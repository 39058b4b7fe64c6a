#!/usr/bin/env python3
"""Benchmark of the hot path named by BASELINE.json: simplex pivots/sec (fp64) against the HBM
roofline, with the reference's own CPU implementation timed beside it.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path

Workload at N=1 (config.workload): BASELINE.json configs[2], the headline roofline case —
G(m=4096, n=16384, n_eq=0, seed=1234), all '<=' rows, most-negative ("Dantzig") pricing, dense
B^-1 (128 MiB) and A (640 MiB) resident in HBM. The inputs are 5x larger than the 126 MB L2, so
no L2 flush is needed between iterations (config.l2: "inputs_larger_than_l2").

A "step" is one pass of the hot path over the LP: sb200_run from the crash basis for
`pivots_per_step` pivots (device-resident loop, one persistent kernel launch per chunk).
  value  pivots/s with A, b, c already in HBM (CUDA events on the launching stream)
  e2e    the same pivots/s through the C-ABI with HOST buffers: every step copies the tableau
         from pinned host memory (sb200_load), runs, and reads x_B + the index lists back
After the timed steps one full solve to optimality gives time_to_optimal_s (not part of `value`).

With --gpus N > 1 (torchrun, one rank per GPU) every rank solves its own LP of the same shape
(seed + rank): independent units, no data-path collective, "scaling": "weak". After that the
ranks solve ONE wide LP together (BASELINE configs[3], m=8192, n=262144, column-sharded A,
row-sharded B^-1, in-kernel NVLink exchanges); its numbers go under "wide_lp" (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", type=int, default=4096)
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--pivots-per-step", type=int, default=2048)
    ap.add_argument("--engine", default="persistent")
    ap.add_argument("--dual", default="update")
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--no-full-solve", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config2 / batched sections")
    ap.add_argument("--batched-lps", type=int, default=65536, help="total LPs of the batched section (split over GPUs)")
    ap.add_argument("--no-wide", action="store_true", help="N>1: skip the sharded wide-LP measurement")
    ap.add_argument("--wide-m", type=int, default=8192)
    ap.add_argument("--wide-n", type=int, default=262144)
    ap.add_argument("--wide-iters", type=int, default=256)
    ap.add_argument("--cpu-iters", type=int, default=12, help="reference iterations timed for cpu_baseline")
    return ap.parse_args()


def workload_string(a):
    """config.workload of BOTH arms (the driver compares the strings)."""
    return ("BASELINE configs[2]: G(m=%d,n=%d,n_eq=0,seed=%d) dense feasible LP, N=%d columns, most-negative (Dantzig) "
            "pricing, all-slack start" % (a.m, a.n, a.seed, a.m + a.n))


def golden_prefix(a):
    """The reference's own first pivots at the headline shape (tests/golden/large.json.gz, written by
    tests/golden/make_golden_large.py from oracle/_ref/ref_driver_dantzig in the build container); None when the
    run is not the golden case."""
    import gzip
    p = os.path.join(ROOT, "tests", "golden", "large.json.gz")
    if not os.path.exists(p):
        return None
    with gzip.open(p, "rt") as f:
        g = json.load(f).get("headline_prefix")
    if not g or (g["m"], g["n"], g["seed"]) != (a.m, a.n, a.seed):
        return None
    return g["pivots"]


def prefix_parity(trace, gold):
    """{"compared": K, "identical": bool, "first_difference": i or None} of a pivot trace against the golden one."""
    if gold is None:
        return None
    k = min(len(trace), len(gold))
    first = next((i for i in range(k) if list(trace[i]) != list(gold[i])), None)
    return {"golden": "tests/golden/large.json.gz:headline_prefix (unmodified reference, Simplex.cpp:255-353)",
            "compared": k, "identical": first is None and k > 0, "first_difference": first}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU baseline
def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "ref_driver_dantzig")
    return p if os.path.exists(p) else None


def time_reference(m, n, seed, iters, budget_s=240.0):
    """Times the UNMODIFIED reference loop (Simplex::simplex, built from /root/reference/src
    against the Eigen stand-in, most-negative switch applied) on the host cores of this box:
    `iters` iterations of the same LP, dense fast path into its matrix_A (seam mode of
    oracle/ref_driver.cpp). Returns (pivots_per_s, seconds_per_iteration list, kind, note)."""
    exe = reference_binary()
    if exe is None:
        return None
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "t.json")
        t0 = time.time()
        try:
            subprocess.run([exe, "seam", str(m), str(n), "0", str(seed), str(iters), "--out", out], cwd=td,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=budget_s)
        except subprocess.TimeoutExpired:
            pass
        if not os.path.exists(out):
            return {"error": "reference run exceeded %.0fs" % budget_s, "wall": time.time() - t0}
        with open(out) as f:
            rec = json.load(f)
    ts = rec.get("iter_timestamps", [])
    if len(ts) < 2:
        return {"error": "reference finished in fewer than 2 iterations"}
    dts = [b - a for a, b in zip(ts[:-1], ts[1:])]
    calls = rec.get("calls") or [{}]
    return {"pivots_per_s": len(dts) / sum(dts), "sec_per_iter": dts, "pivots": calls[0].get("pivots", [])}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count()


def cpu_baseline_block(a, iters):
    r = time_reference(a.m, a.n, a.seed, iters)
    base = {"unit": "pivots/s", "cores": 1, "kind": "reference",
            "host_cores_available": host_cores(),
            "note": "unmodified reference sources (single-threaded by design) + Eigen stand-in with a blocked "
                    "right-looking LU whose trailing update is a packed register-blocked GEMM (genuine Eigen is not "
                    "vendored/installed); one iteration = gather + dense LU of the %dx%d basis + 3 solves + "
                    "pricing" % (a.m, a.m)}
    if r is None:
        base.update({"value": None, "sample": "oracle/_ref binaries missing (make -C oracle ref)"})
    elif "error" in r:
        base.update({"value": None, "sample": r["error"]})
    else:
        base.update({"value": r["pivots_per_s"],
                     "sample": "first %d iterations of the same LP (G(%d,%d,0,%d), most-negative pricing) through "
                               "Simplex::simplex()" % (len(r["sec_per_iter"]), a.m, a.n, a.seed),
                     "sec_per_iter": [round(x, 4) for x in r["sec_per_iter"]]})
    return base


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    iters = max(1, a.steps)
    warm = max(0, a.warmup)
    # one reference iteration at the headline size takes seconds: bound the whole run
    r = time_reference(a.m, a.n, a.seed, iters + warm, budget_s=400.0)
    line = {"impl": "reference", "metric": "simplex_pivots_per_sec", "unit": "pivots/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(a),
                       "step": "one iteration of the reference loop (gather, fresh LU of the basis, 3 solves, pricing)"}}
    cb = {"unit": "pivots/s", "cores": 1, "kind": "reference", "host_cores_available": host_cores()}
    if r is None or "error" in r:
        why = "oracle/_ref binaries missing" if r is None else r["error"]
        line.update({"value": None, "ms_per_step": None, "unavailable": why})
        cb.update({"value": None, "sample": why})
    else:
        dts = r["sec_per_iter"][warm:] or r["sec_per_iter"]
        v = len(dts) / sum(dts)
        line.update({"value": v, "ms_per_step": 1e3 * sum(dts) / len(dts)})
        cb.update({"value": v, "sample": "%d timed iterations after %d warm-up iterations of Simplex::simplex() on the "
                                         "same LP; unmodified reference sources + Eigen stand-in (blocked LU), "
                                         "single-threaded as the reference is" % (len(dts), warm)})
    line["cpu_baseline"] = cb
    line["e2e"] = {"value": line["value"], "unit": "pivots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if r is not None and "error" not in r:
        # the pivots this very run took, against the committed golden prefix the GPU arm is checked against
        line["parity_prefix"] = prefix_parity(r.get("pivots", []), golden_prefix(a))
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha():
    import hashlib
    h = hashlib.sha256()
    for f in ("sb200_fused.cuh", "sb200_persistent.cuh", "sb200_state.cuh", "sb200_phases.cuh"):
        with open(os.path.join(ROOT, "simplex_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def traffic_per_launch(chunk):
    """(dram bytes per k_persistent_fused launch, note). From the committed `ncu --set full` capture
    (profiles/traffic.json, written by scripts/capture_traffic.sh), scaled to this run's pivots per launch. The
    capture carries the hash of the kernel sources it was taken from: a capture of OTHER sources is not reported
    (traffic: null) instead of going stale silently."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, "no capture committed"
    with open(p) as f:
        t = json.load(f)
    if t.get("kernel_source_sha256") != kernel_source_sha():
        return None, "profiles/traffic.json was captured from other kernel sources (re-run scripts/capture_traffic.sh)"
    return (t.get("dram_bytes_per_pivot", 0) * chunk or None), t.get("source", "profiles/traffic.json")


def bench_config2(a, stream, local):
    """BASELINE configs[1]: G(m=1024, n=2048, n_eq=256, seed), two-phase, most-negative pricing.
    Phase I (N=3072 with the 256 artificial columns) and Phase II (N=2816, the same A kept resident:
    sb200_set_costs) through the C ABI; the 31 MB working set lives in L2, so this shape is bound by
    grid-barrier latency, not by HBM. Reports time-to-optimal and pivots/s over both loops."""
    import numpy as np
    import torch

    from simplex_b200 import Basis, DeviceSimplex, algorithmic_bytes_per_pivot, synth
    m, n, n_eq = 1024, 2048, 256
    A1, b1, c1, basic1, nonbasic1 = synth.tableau(m, n, n_eq, a.seed, phase1=True)
    _, _, c2, _, _ = synth.tableau(m, n, n_eq, a.seed, phase1=False)
    n2 = synth.num_cols(m, n, n_eq, False)
    best = None
    for rep in range(3):
        s = DeviceSimplex(pricing="most_neg", engine=a.engine, dual=a.dual, device=local, chunk_iters=a.chunk,
                          stream=stream.cuda_stream, trace_capacity=8192)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        s.load(A1, b1, c1)
        bs = Basis(basic1, nonbasic1)
        r1 = s.simplex(bs)
        if rep == 0:
            tr1, _ = s.trace()
        bas2, non2 = synth.phase2_lists(bs.basic, bs.non_basic, n2)
        s.set_costs(c2)
        bs2 = Basis(bas2, non2)
        r2 = s.simplex(bs2)
        x = s.xB()
        e1.record(stream)
        torch.cuda.synchronize()
        if rep == 0:
            tr2, _ = s.trace()
        used = s.engine_used()
        s.close()
        tot = e0.elapsed_time(e1) * 1e-3
        loop = r1["loop_seconds"] + r2["loop_seconds"]
        rec = {"time_to_optimal_s": tot, "loop_s": loop, "phase1": r1, "phase2": r2, "min_xB": float(x.min()),
               "engine_used": used}
        if best is None or rec["loop_s"] < best["loop_s"]:
            best = rec
    # the same LP through the host C++ flow (simplex_b200/host/dense_solver.cpp: standard form, scaling,
    # crash basis, Phase I, hand-off, Phase II, read-out; wall clock of the whole call)
    from simplex_b200 import lp as hostlp
    cr, ar, rhsr, types = synth.raw(m, n, n_eq, a.seed)
    hostlp.solve_dense(ar, types, rhsr, cr, pricing="most_neg", device=local, engine=a.engine, dual=a.dual, chunk_iters=a.chunk)
    t0 = time.time()
    hd = hostlp.solve_dense(ar, types, rhsr, cr, pricing="most_neg", device=local, engine=a.engine, dual=a.dual,
                            chunk_iters=a.chunk)
    t_host = time.time() - t0
    r1, r2 = best["phase1"], best["phase2"]
    piv = r1["pivots"] + r2["pivots"]
    # both phases against the reference's own complete run of this LP (tests/golden/large.json.gz:config1)
    parity = None
    try:
        import gzip
        with gzip.open(os.path.join(ROOT, "tests", "golden", "large.json.gz"), "rt") as f:
            g = json.load(f)["config1"]
        if (g["m"], g["n"], g["n_eq"], g["seed"]) == (m, n, n_eq, a.seed):
            p1, p2 = prefix_parity(tr1.tolist(), g["calls"][0]["pivots"]), prefix_parity(tr2.tolist(), g["calls"][1]["pivots"])
            parity = {"golden": "tests/golden/large.json.gz:config1 (unmodified reference through Simplex::solve())",
                      "phase1": [p1["compared"], p1["identical"]], "phase2": [p2["compared"], p2["identical"]],
                      "reference_iterations": [c["iterations"] for c in g["calls"]],
                      "reference_objective": g["objective"]}
    except (OSError, KeyError):
        parity = None
    bytes2 = algorithmic_bytes_per_pivot(m, n2 - m)
    return {"workload": "BASELINE configs[1]: G(m=1024,n=2048,n_eq=256,seed=%d) two-phase, most-negative pricing, "
                        "A resident across the phases (sb200_set_costs)" % a.seed,
            "status": r2["term"], "phase1_iterations": r1["iterations"], "phase2_iterations": r2["iterations"],
            "pivots": piv, "objective": r2["objective"], "phase1_objective": r1["objective"],
            "time_to_optimal_s": best["time_to_optimal_s"], "device_loop_s": best["loop_s"],
            "pivots_per_s": piv / best["loop_s"], "us_per_pivot": 1e6 * best["loop_s"] / max(piv, 1),
            "engine_used": best["engine_used"], "parity": parity,
            "bound": ("latency (tableau resident in the SMs' shared memory; two fence-free mail exchanges through L2 per pivot)"
                      if best["engine_used"] == "resident" else
                      "latency (working set 31 MB < 126 MB L2; 2 grid barriers per pivot)"),
            "algorithmic_GBs_phase2": bytes2 * r2["pivots"] / max(r2["loop_seconds"], 1e-12) / 1e9,
            "min_xB": best["min_xB"],
            "host_cpp_flow": {"entry": "simplex::solve_dense (libsimplex_b200_host.so)", "status": hd["status"],
                              "iterations": [k["iterations"] for k in hd["calls"]], "objective": hd["objective"],
                              "time_to_optimal_s_wall": t_host,
                              "device_loop_s": sum(k["loop_us"] for k in hd["calls"]) * 1e-6,
                              "phase2_reused_resident_tableau": bool(hd["calls"][-1]["resident"])},
            "cpu_baseline": config2_cpu_baseline(a, m, n) if not a.no_cpu_baseline else None}


def config2_cpu_baseline(a, m, n):
    """Reference loop at the config-2 shape on one host core: the seam mode of oracle/ref_driver.cpp
    (all-'<=' variant of the same G(m,n): same M, N_nb = n, same per-iteration work as Phase II)."""
    r = time_reference(m, n, a.seed, 12, budget_s=120.0)
    if r is None or "error" in r:
        return {"value": None, "sample": "reference binary missing" if r is None else r["error"]}
    return {"value": r["pivots_per_s"], "unit": "pivots/s", "cores": 1, "kind": "reference",
            "sample": "first %d iterations of G(%d,%d,0,%d) through Simplex::simplex() (unmodified reference sources + Eigen "
                      "stand-in)" % (len(r["sec_per_iter"]), m, n, a.seed)}


def batched_cpu_baseline(a, m, n, sample, gpu_iterations=None, gpu_objective=None):
    """The oracle's C restatement of the reference loop (LU per iteration, oracle/simplex_oracle.c) on
    `sample` LPs of the batch, one host core, sequentially — what the reference would do with a batch. The same
    runs also CHECK the GPU's results for those LPs (iteration counts equal, objectives within 1e-9)."""
    from oracle import oracle

    from simplex_b200 import synth
    t_tot, piv, same_its, same_obj = 0.0, 0, 0, 0
    for k in range(sample):
        A, b, c, bas, non = synth.tableau(m, n, 0, a.seed + k)
        t0 = time.time()
        r = oracle.simplex(A, b, c, bas, non, rule=oracle.RULE_MOST_NEG, variant="lu")
        t_tot += time.time() - t0
        piv += r["pivots"]
        if gpu_iterations is not None:
            same_its += int(gpu_iterations[k]) == int(r["iterations"])
            same_obj += abs(float(gpu_objective[k]) - r["objective"]) <= 1e-9 * max(1.0, abs(r["objective"]))
    out = {"value": sample / t_tot, "unit": "LPs/s", "pivots_per_s": piv / t_tot, "cores": 1, "kind": "port",
           "sample": "%d LPs G(%d,%d,0,seed+k) solved one after the other by the oracle's restatement of the reference "
                     "loop (fresh LU every iteration)" % (sample, m, n)}
    if gpu_iterations is not None:
        out["gpu_matches_oracle"] = {"lps": sample, "same_iteration_count": same_its, "objective_within_1e-9": same_obj}
    return out


def bench_batched(a, stream, local, count, world, rank):
    """BASELINE configs[4]: independent small LPs G(64, 128, 0, seed+k), one CTA per LP, all state
    in shared memory. `count` LPs on this GPU (65536 split over the GPUs). HBM sees the inputs once."""
    import numpy as np
    import torch

    from simplex_b200 import DeviceSimplex, synth
    m, n = 64, 128
    N = m + n
    # inputs in PINNED host memory (the e2e leg copies them from there, chunk by chunk, beside the kernels)
    A_pin = torch.empty((count, N, m), dtype=torch.float64, pin_memory=True)
    A, b, c, basic, nonbasic = synth.batch(count, m, n, a.seed + rank * count, out=A_pin.numpy())
    b_pin, c_pin = torch.from_numpy(b).pin_memory(), torch.from_numpy(c).pin_memory()
    dA = A_pin.cuda()
    db = torch.from_numpy(b).cuda()
    dc = torch.from_numpy(c).cuda()
    s = DeviceSimplex(pricing="most_neg", device=local, stream=stream.cuda_stream, iter_limit=100000)
    best = None
    for rep in range(3):
        dbas = torch.from_numpy(basic).cuda()
        dnon = torch.from_numpy(nonbasic).cuda()
        r = s.run_batched_device(count, m, N, dA, db, dc, dbas, dnon)
        if best is None or r["seconds"] < best["seconds"]:
            best = r
    its = best["iterations"].cpu().numpy()
    term = best["term"].cpu().numpy()
    del dA
    # end to end from host buffers (H2D of every LP + D2H of the results inside the call): chunks of 4096 LPs through
    # two device buffers, copy of chunk k+1 beside the kernel of chunk k; first call allocates the persistent buffers
    s.run_batched(A, b_pin.numpy(), c_pin.numpy(), basic.copy(), nonbasic.copy())
    t_host = None
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.time()
        r_host = s.run_batched(A, b_pin.numpy(), c_pin.numpy(), basic.copy(), nonbasic.copy())
        dt = time.time() - t0
        t_host = dt if t_host is None else min(t_host, dt)
    # the same from pageable buffers (the library stages them through its own pinned buffers with a few copy threads)
    A_pageable = np.array(A[:min(count, 16384)])
    npg = A_pageable.shape[0]
    t0 = time.time()
    s.run_batched(A_pageable, b[:npg].copy(), c[:npg].copy(), basic[:npg].copy(), nonbasic[:npg].copy())
    t_pageable = time.time() - t0
    del A_pageable
    s.close()
    pivots = int(its.sum() - count)   # the last iteration of every LP only detects optimality
    hbm_bytes = count * 8.0 * (N * m + m + N) + count * (8.0 * m + 4.0 * N + 16)
    return {"workload": "BASELINE configs[4]: %d independent LPs G(m=64,n=128,n_eq=0) on this GPU (65536 over %d GPUs), "
                        "one CTA per LP, most-negative pricing" % (count, world),
            "lps": count, "all_optimal": bool((term == 1).all()), "pivots": pivots,
            "pivots_per_lp_mean": float(its.mean() - 1.0), "seconds": best["seconds"],
            "lps_per_s": count / best["seconds"], "pivots_per_s": pivots / best["seconds"],
            "hbm_GBs": hbm_bytes / best["seconds"] / 1e9,
            "onchip_GBs": pivots * 8.0 * (m * n + 4 * m * m) / best["seconds"] / 1e9,
            "bound": "shared-memory bandwidth / fp64 issue (state on chip; HBM reads the inputs once)",
            "e2e_lps_per_s": count / t_host, "e2e_seconds": t_host,
            "e2e_h2d_bytes": int(count * (8 * (N * m + m + N) + 4 * N)), "e2e_h2d_GBs": count * (8.0 * (N * m + m + N) + 4 * N) / t_host / 1e9,
            "e2e_note": "inputs in pinned host memory, 4096-LP chunks through two device buffers, H2D of chunk k+1 beside the "
                        "kernel of chunk k, results read back per chunk; bound by the host link, not by the kernel",
            "e2e_pageable_lps_per_s": npg / t_pageable, "e2e_pageable_lps": npg,
            "e2e_matches_resident": bool(np.array_equal(r_host["iterations"], its)),
            "cpu_baseline": batched_cpu_baseline(a, m, n, 256, its, best["objective"].cpu().numpy())
            if (rank == 0 and not a.no_cpu_baseline) else None}


def bench_bounded(a, stream, local, A_host, b, c, basic, nonbasic, n):
    """The headline LP with FINITE upper bounds (reference: has_non_trivial_upper_bounds(), Simplex.cpp:108, bounded
    ratio test :116-179): (1) bounds nobody reaches (the bounded code path, no flip), (2) bounds tight enough for flips.
    Both run the fused engine's bounded variant; the 4-phase kernel is timed beside it. (3) A bounded LP at the
    two-phase shape (m=1024, n=2048) on the resident kernel's bounded variant, the fused one beside it."""
    import numpy as np

    from simplex_b200 import Basis, DeviceSimplex
    DBL_MAX = float(np.finfo(np.float64).max)
    out = {}
    its = 1024
    rng = np.random.default_rng(a.seed)
    for name, ubs in (("loose", np.full(n, 1e6)), ("tight", np.where(rng.random(n) < 0.3, rng.uniform(0.01, 0.3, n), 50.0))):
        ub = np.full(A_host.shape[1], DBL_MAX)
        ub[:n] = ubs
        rec = {}
        for dual in ("update", "recompute"):
            with DeviceSimplex(pricing="most_neg", engine=a.engine, dual=dual, device=local, iter_limit=its,
                               chunk_iters=a.chunk, stream=stream.cuda_stream) as s:
                s.load(A_host, b, c, ub)
                r = s.simplex(Basis(basic, nonbasic))      # warm-up (tuning-free engine: same run twice)
                s.load(A_host, b, c, ub)
                r = s.simplex(Basis(basic, nonbasic))
                rec[s.engine_used()] = {"us_per_iteration": 1e6 * r["loop_seconds"] / max(r["iterations"], 1),
                                        "iterations": r["iterations"], "pivots": r["pivots"], "flips": r["flips"],
                                        "objective": r["objective"]}
        rec["same_objective"] = bool(rec["fused"]["objective"] == rec["persistent"]["objective"]) if \
            ("fused" in rec and "persistent" in rec) else None
        out[name] = rec
    # the same at the shape configs[1] is quoted on (m=1024, n=2048): small enough for the tableau to live in the SMs'
    # shared memory, so the default engine is the resident kernel's bounded variant; the fused one is timed beside it
    from simplex_b200 import synth
    ms, ns = 1024, 2048
    As, bs, cs, bas, non = synth.tableau(ms, ns, 0, a.seed)
    ub = np.full(As.shape[1], DBL_MAX)
    ub[:ns] = np.where(rng.random(ns) < 0.4, rng.uniform(0.01, 0.3, ns), rng.uniform(1.0, 50.0, ns))
    rec = {}
    for engine in ("persistent", "streaming"):
        with DeviceSimplex(pricing="most_neg", engine=engine, dual="update", device=local, iter_limit=its,
                           chunk_iters=a.chunk, stream=stream.cuda_stream) as s:
            for _ in range(2):                             # (the first run warms up)
                s.load(As, bs, cs, ub)
                r = s.simplex(Basis(bas, non))
            rec[s.engine_used()] = {"us_per_iteration": 1e6 * r["loop_seconds"] / max(r["iterations"], 1),
                                    "iterations": r["iterations"], "pivots": r["pivots"], "flips": r["flips"],
                                    "objective": r["objective"]}
    rec["same_objective"] = bool(rec["resident"]["objective"] == rec["fused"]["objective"]) if \
        ("resident" in rec and "fused" in rec) else None
    out["m1024_n2048_tight"] = rec
    return out


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    from simplex_b200 import Basis, DeviceSimplex, algorithmic_bytes_per_pivot, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: simplex_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    m, n, P = a.m, a.n, a.pivots_per_step
    seed = a.seed + rank
    N = synth.num_cols(m, n, 0, False)
    # pinned host tableau (column-major, as the reference's matrix_A)
    A_pin = torch.empty((N, m), dtype=torch.float64, pin_memory=True)
    A_host = A_pin.numpy().T  # (m, N) Fortran-ordered view of the pinned buffer
    _, b, c, basic, nonbasic = synth.tableau(m, n, 0, seed, out=A_host)
    b_pin = torch.from_numpy(b).pin_memory()
    c_pin = torch.from_numpy(c).pin_memory()

    stream = torch.cuda.current_stream()
    solver = DeviceSimplex(pricing="most_neg", engine=a.engine, dual=a.dual, device=local, iter_limit=P,
                           chunk_iters=a.chunk, stream=stream.cuda_stream, trace_capacity=P)
    solver.load(A_host, b_pin.numpy(), c_pin.numpy())

    def step_resident():
        return solver.simplex(Basis(basic, nonbasic))

    def step_e2e():
        solver.load(A_host, b_pin.numpy(), c_pin.numpy())       # H2D of the whole tableau
        bs = Basis(basic, nonbasic)
        r = solver.simplex(bs)                                   # lists come back in bs (D2H)
        x = solver.xB()                                          # D2H of the result vector
        return r, x

    # ---- resident: value
    for _ in range(a.warmup):
        step_resident()
    launches0 = solver.launch_count()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    pivots = 0
    loop_s = 0.0
    for _ in range(a.steps):
        r = step_resident()
        pivots += r["pivots"]
        loop_s += r["loop_seconds"]
    e1.record(stream)
    barrier()
    clk = clocks.stop()
    launches = solver.launch_count() - launches0
    t_res = e0.elapsed_time(e1) * 1e-3
    last_obj = r["objective"]
    # the pivots of the last timed step against the reference's own first pivots on this LP (rank 0: seed = a.seed)
    parity = None
    if rank == 0:
        step_trace, _ = solver.trace()
        parity = prefix_parity(step_trace.tolist(), golden_prefix(a))

    # ---- end to end through the C ABI with host buffers
    for _ in range(min(a.warmup, 2)):
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    piv_e2e = 0
    for _ in range(a.steps):
        r2, _x = step_e2e()
        piv_e2e += r2["pivots"]
    f1.record(stream)
    barrier()
    t_e2e = f0.elapsed_time(f1) * 1e-3

    # ---- max over ranks, totals
    if world > 1:
        tt = torch.tensor([t_res, t_e2e, loop_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e, loop_s = [float(v) for v in tt.tolist()]
        pp = torch.tensor([pivots, piv_e2e, launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(pp, op=dist.ReduceOp.SUM)
        pivots, piv_e2e, launches = [int(v) for v in pp.tolist()]

    # ---- one full solve to optimality (time-to-optimal), rank 0 only reports it
    full = None
    if not a.no_full_solve:
        solver.close()
        solver = DeviceSimplex(pricing="most_neg", engine=a.engine, dual=a.dual, device=local, chunk_iters=a.chunk,
                               stream=stream.cuda_stream)
        solver.load(A_host, b_pin.numpy(), c_pin.numpy())
        bs = Basis(basic, nonbasic)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        rf = solver.simplex(bs)
        g1.record(stream)
        torch.cuda.synchronize()
        x = solver.xB()
        # drift of the carried B^-1 / x_B at the end of the whole solve (the reference re-factorises every iteration)
        full = {"term": rf["term"], "pivots": rf["pivots"], "iterations": rf["iterations"],
                "time_to_optimal_s": g0.elapsed_time(g1) * 1e-3, "objective": rf["objective"],
                "pivots_per_s": rf["pivots"] / max(rf["loop_seconds"], 1e-12),
                "min_xB": float(x.min()), "hygiene": solver.hygiene_stats(),
                "primal_residual_max_abs": solver.primal_residual(),
                "inverse_residual_64_rows": solver.inverse_residual(64, 17),
                "reinversions": solver.reinversion_count()}
    solver.close()

    # ---- the other single-GPU configurations of BASELINE.json, short runs, reported as extra keys
    extras = {}
    if not a.no_extras:
        # a failure in a secondary section must never cost the headline line (nor desynchronise the ranks:
        # the collectives below run whatever happened locally)
        cfg2 = None
        bnd = None
        if rank == 0:
            try:
                cfg2 = bench_config2(a, stream, local)
            except Exception as exc:  # noqa: BLE001
                cfg2 = {"error": "%s: %s" % (type(exc).__name__, exc)}
            try:
                bnd = bench_bounded(a, stream, local, A_host, b, c, basic, nonbasic, n)
            except Exception as exc:  # noqa: BLE001
                bnd = {"error": "%s: %s" % (type(exc).__name__, exc)}
        per_gpu = max(1, a.batched_lps // world)
        try:
            bat = bench_batched(a, stream, local, per_gpu, world, rank)
        except Exception as exc:  # noqa: BLE001
            bat = {"error": "%s: %s" % (type(exc).__name__, exc), "seconds": 0.0, "pivots": 0, "lps": 0}
        if world > 1:
            tt = torch.tensor([bat["seconds"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            pp = torch.tensor([bat["pivots"], bat["lps"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(pp, op=dist.ReduceOp.SUM)
            if float(tt.item()) > 0.0:
                bat["all_gpus"] = {"lps": int(pp[1].item()), "pivots": int(pp[0].item()),
                                   "seconds_max_over_ranks": float(tt.item()),
                                   "lps_per_s": float(pp[1].item() / tt.item()),
                                   "pivots_per_s": float(pp[0].item() / tt.item()),
                                   "scaling": "strong (%d LPs split over the GPUs, no communication)" % a.batched_lps}
        extras = {"config2_two_phase": cfg2, "bounded_headline": bnd, "batched_small_lps": bat}

    # ---- N > 1 only: ONE wide LP sharded over all GPUs (BASELINE configs[3]: m=8192, n=262144,
    # A column-sharded, B^-1 row-sharded, NVLink arg-min exchanges and column/row broadcasts done
    # by the kernels over peer memory). Reported under "wide_lp"; not part of `value`.
    wide = None
    wide_short = None
    if world > 1 and not a.no_wide:
        try:
            import gzip
            from simplex_b200.dist import ShardedSimplex, shard_plan
            del A_pin, A_host
            with gzip.open(os.path.join(ROOT, "tests", "golden", "large.json.gz"), "rt") as f:
                gold = json.load(f)

            def sharded_run(wm, wn, wseed, wit, reps, chunk):
                wN = synth.num_cols(wm, wn, 0, False)
                me = shard_plan(wm, wN, world)[rank]
                Aw, bw, cw, basw, nonw = synth.tableau(wm, wn, 0, wseed, col0=me["col0"], ncols=me["ncols"],
                                                       col_stride=me["col_stride"])
                sh = ShardedSimplex(rank, world, pricing="most_neg", device=local, iter_limit=wit, chunk_iters=chunk,
                                    stream=stream.cuda_stream, trace_capacity=wit)
                sh.load_shard(wm, wN, Aw, bw, cw)
                best = None
                for _ in range(reps):            # first pass warms up, the last is reported
                    best = sh.simplex(Basis(basw, nonw))
                tt = torch.tensor([best["loop_seconds"]], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                tr, _ = sh.trace()
                pc = sh.phase_cycles()
                best["primal_residual_max_abs"] = sh.primal_residual(bw)
                sh.close()
                return best, float(tt.item()), tr.tolist(), pc

            # (1) the reference's own trace on a wide LP the CPU can finish: G(1024, 8192, 0, 1234), first 1000 pivots
            gs = gold.get("sharded_check")
            small = None
            if gs is not None:
                rs, _, trs, _ = sharded_run(gs["m"], gs["n"], gs["seed"], 1000, 1, 250)
                small = prefix_parity(trs, gs["pivots"])
                small["golden"] = "tests/golden/large.json.gz:sharded_check (unmodified reference, complete run)"
            # (2) BASELINE configs[3]
            wm, wn, wit = a.wide_m, a.wide_n, a.wide_iters
            best, wsec, trw, pcw = sharded_run(wm, wn, a.seed, wit, 2, 128)
            gw = gold.get("wide_prefix")
            wpar = None
            if gw is not None and (gw["m"], gw["n"], gw["seed"]) == (wm, wn, a.seed):
                wpar = prefix_parity(trw, gw["pivots"])
                wpar["golden"] = "tests/golden/large.json.gz:wide_prefix (unmodified reference, first iterations)"
            itw = max(int(pcw[7]), 1)
            us_pp = 1e6 * wsec / max(best["pivots"], 1)
            # phase shares of rank 0 (SM-cycle counters of CTA 0), scaled to the event-timed pivot: the SM clock
            # under this load is not the nominal one
            cyc = [float(pcw[i]) / itw for i in range(7)]
            scale = us_pp / max(sum(cyc), 1e-9)
            wphases = {k: round(cyc[i] * scale, 1)
                       for i, k in enumerate(["price", "xchg1_argmin", "fused_update_ftran", "xchg2_argmin", "pivot_row_tail",
                                              "wait_slowest_cta_after_price", "wait_slowest_cta_after_fused"])}
            wbytes = algorithmic_bytes_per_pivot(wm, wn)
            peak_w, _ = peaks()
            gbs = wbytes * best["pivots"] / wsec / 1e9
            wide = {"workload": "BASELINE configs[3]: G(m=%d,n=%d,n_eq=0,seed=%d), one LP over %d GPUs, cyclic column "
                                "shards of A, row shards of B^-1" % (wm, wn, a.seed, world),
                    "pivots": best["pivots"], "us_per_pivot": us_pp,
                    "pivots_per_s": best["pivots"] / wsec, "scaling": "strong",
                    "algorithmic_GBs_all_gpus": gbs,
                    "frac_of_measured_hbm_x_gpus": gbs / (peak_w * world),
                    "frac_of_8TBs_nominal_x_gpus": gbs / (8000.0 * world),
                    "hbm_roofline_us_per_pivot_at_8TBs": wbytes / world / 8e12 * 1e6,
                    "exchange": "in-kernel NVLink peer-window stores: 2 arg-min exchanges per pivot; every rank pushes its "
                                "candidate column / candidate row beside its candidate (self-validating words, no fence)",
                    "phases_us_rank0": wphases,
                    "nvlink_bytes_per_pivot": int(2 * 16 * wm * world * (world - 1) + 2 * 48 * world * world),
                    "nvlink_note": "latency-bound: < 8 MB per pivot over the whole box; xchg1/xchg2/tail include the skew "
                                   "between ranks and the local grid barriers",
                    "objective": best["objective"], "primal_residual_max_abs": best["primal_residual_max_abs"],
                    "parity_prefix": wpar, "reference_trace_check": small}
            wide_short = {"gpus": world, "us_per_pivot": round(us_pp, 1), "frac_8TBs": round(gbs / (8000.0 * world), 4),
                          "frac_hbm_measured": round(gbs / (peak_w * world), 4),
                          "xchg_us": round(wphases["xchg1_argmin"] + wphases["xchg2_argmin"] + wphases["pivot_row_tail"], 1),
                          "local_barrier_us": round(wphases["wait_slowest_cta_after_price"] + wphases["wait_slowest_cta_after_fused"], 1),
                          "price_us": wphases["price"], "fused_us": wphases["fused_update_ftran"],
                          "wide_prefix_identical": None if wpar is None else [wpar["compared"], wpar["identical"]],
                          "ref_trace_1024x8192_identical": None if small is None else [small["compared"], small["identical"]],
                          "objective": best["objective"]}
        except Exception as exc:  # noqa: BLE001 - the headline line must still be printed
            wide = {"error": "%s: %s" % (type(exc).__name__, exc)}
            wide_short = {"error": wide["error"][:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    bytes_pp = algorithmic_bytes_per_pivot(m, n)  # N_nb = n for the all-'<=' LP
    peak, peak_src = peaks()
    value = pivots / t_res
    # the dominant kernel is k_persistent: one launch runs `chunk` pivots; its average duration is the
    # library's CUDA-event loop time (same stream) divided by the number of launches
    n_launch_loop = max(1, a.steps * world * -(-P // a.chunk))
    per_world_loop = loop_s  # max over ranks of the summed loop time
    achieved = (bytes_pp * (pivots / world)) / per_world_loop / 1e9
    line = {
        "metric": "simplex_pivots_per_sec", "value": value, "unit": "pivots/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t_res / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(a),
                   "step": "sb200_run from the crash basis for pivots_per_step pivots, A and B^-1 resident in HBM",
                   "pivots_per_step": P, "engine": a.engine, "dual": a.dual, "pivots_per_launch": a.chunk,
                   "l2": "inputs_larger_than_l2 (A 640 MiB + B^-1 128 MiB vs 126 MB L2)",
                   "parallelism": "1 LP per GPU" if world > 1 else "1 GPU"},
        "gpu_launches": launches,
        "clocks": clk,
        "e2e": {"value": piv_e2e / t_e2e, "unit": "pivots/s",
                "h2d_bytes_per_step": int(8 * (m * N + m + N)), "d2h_bytes_per_step": int(8 * m + 8 * N + 64),
                "ms_per_step": 1e3 * t_e2e / a.steps},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "frac_of_8TBs_nominal": achieved / 8000.0,
                     "kernel": ("k_persistent_fused" if (a.engine == "persistent" and a.dual == "update")
                                else "k_persistent" if a.engine == "persistent" else "staged K1-K5 graph"),
                     "note": "achieved = ALGORITHMIC bytes (SURVEY 8d: 8(m*N_nb+4m^2), dual solve + pricing + FTRAN + "
                             "rank-1 read/write counted separately) / measured time; the fused kernel moves fewer DRAM "
                             "bytes than that (update k and FTRAN k+1 share one pass over B^-1, duals by recurrence), "
                             "see `traffic`, which is why frac can exceed 1",
                     "algorithmic_bytes_per_pivot": bytes_pp,
                     "pivots_per_launch": a.chunk, "launches_timed": n_launch_loop,
                     "avg_launch_ms": 1e3 * per_world_loop / (n_launch_loop / world),
                     "traffic": traffic_per_launch(a.chunk)[0], "traffic_note": traffic_per_launch(a.chunk)[1]},
        "objective_after_step": last_obj,
        "parity_prefix": parity,
    }
    tpl = line["roofline"]["traffic"]
    if tpl:
        # real DRAM bytes (ncu capture in profiles/) against the measured launch time: the kernel-quality
        # number next to the algorithmic one
        dram_gbs = tpl / (line["roofline"]["avg_launch_ms"] * 1e-3) / 1e9
        line["roofline"].update({"dram_GBs_from_ncu_traffic": dram_gbs, "dram_frac_of_measured_peak": dram_gbs / peak,
                                 "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + "
                                                   "dram__bytes_write.sum per launch, scaled to pivots_per_launch; "
                                                   "stamped with the hash of the kernel sources)"})
    if full is not None:
        line["full_solve"] = full
    if wide is not None:
        line["wide_lp"] = wide
    for k, v in extras.items():
        if v is not None:
            line[k] = v
    if not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_block(a, a.cpu_iters)
    if wide_short is not None:
        line["wide"] = wide_short   # configs[3] in a few flat numbers, LAST on the line (the driver keeps the tail)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
