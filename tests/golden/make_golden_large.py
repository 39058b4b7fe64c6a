#!/usr/bin/env python3
"""Golden vectors of the UNMODIFIED reference at the sizes BASELINE.json quotes (test infrastructure).

tests/golden/make_golden.py stops at m=512 (seconds per case). The cases below take minutes to
half an hour of one host core each — the reference re-factorises the m x m basis every iteration
(/root/reference/src/Simplex.cpp:305) — so they live in their own script and their own file,
tests/golden/large.json.gz. Runs only in the build container (needs /root/reference); the GPU box
sees the committed JSON.

    python tests/golden/make_golden_large.py [--work DIR] [--jobs J] [--only NAME ...]

Every case is one run of oracle/_ref/ref_driver_{dantzig,stock} (reference sources compiled where
they lie, see oracle/Makefile); raw driver outputs are kept in --work (default /tmp/gold) and reused
when present, so an interrupted generation can be resumed.

Cases (name: driver arguments -> what is kept)
    config1            synth 1024 2048 256 1234   BASELINE configs[1], two-phase, most-negative pricing, through
                                                  the reference's model API + Simplex::solve(): both calls of
                                                  Simplex::simplex() with index lists in/out, the complete pivot
                                                  traces (727 + 6059 pivots), objective and structural x
    headline_prefix    seam 4096 16384 0 1234 K   BASELINE configs[2] (the headline roofline case): the first K
                                                  iterations of Simplex::simplex() on the dense tableau (seam mode
                                                  of oracle/ref_driver.cpp; ~1 s per iteration)
    wide_prefix        seam 8192 262144 0 1234 K  BASELINE configs[3] shape (35 GB of host memory, ~13 s/iteration)
    sharded_check      seam 1024 8192 0 1234 -    the shape scripts/sharded_check.py and the multi-rank tests use:
                                                  complete trace to the optimum
    mid_prefix         seam 2048 4096 0 1234 K    a mid-size LP (tableau in L2, not in shared memory)
    first_eligible     seam 1024 2048 0 1234 K    the reference's stock rule (first eligible position) at m=1024
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SB200_REFERENCE", "/root/reference")
DRV = {"stock": os.path.join(ROOT, "oracle/_ref/ref_driver_stock"),
       "dantzig": os.path.join(ROOT, "oracle/_ref/ref_driver_dantzig")}

# name, driver, mode, m, n, n_eq, seed, K (seam: iterations to run, 0 = to the end)
CASES = [
    ("config1", "dantzig", "synth", 1024, 2048, 256, 1234, 0),
    ("headline_prefix", "dantzig", "seam", 4096, 16384, 0, 1234, 2048),
    ("wide_prefix", "dantzig", "seam", 8192, 262144, 0, 1234, 12),
    ("sharded_check", "dantzig", "seam", 1024, 8192, 0, 1234, 0),
    ("mid_prefix", "dantzig", "seam", 2048, 4096, 0, 1234, 1500),
    ("first_eligible", "stock", "seam", 1024, 2048, 0, 1234, 3000),
]


def run_case(work, case):
    name, drv, mode, m, n, n_eq, seed, K = case
    out = os.path.join(work, name + ".json")
    if not os.path.exists(out):
        cmd = [DRV[drv], mode, str(m), str(n), str(n_eq), str(seed)]
        if mode == "seam":
            cmd.append(str(K if K > 0 else 1000000000))
        cmd += ["--out", out + ".part"]
        subprocess.run(cmd, cwd=work, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)
        os.replace(out + ".part", out)
    with open(out) as f:
        return json.load(f)


def replay(m, n, pivots):
    """Index lists after the recorded swaps (Basis.cpp:48-53) from the all-slack start of the seam mode:
    basic = slack of every row in row order, non-basic = the structural columns ascending."""
    basic = [n + i for i in range(m)]
    nonbasic = list(range(n))
    for lc, row, ec, pos in pivots:
        assert basic[row] == lc and nonbasic[pos] == ec, "trace does not replay"
        basic[row] = ec
        nonbasic[pos] = lc
    return basic, nonbasic


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--work", default="/tmp/gold")
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--only", nargs="*")
    a = ap.parse_args()
    if not os.path.isdir(REF):
        sys.exit("reference tree not found: golden vectors can only be regenerated in the build container")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL)
    os.makedirs(a.work, exist_ok=True)
    cases = [c for c in CASES if not a.only or c[0] in a.only]
    with ThreadPoolExecutor(max_workers=a.jobs) as ex:
        raw = list(ex.map(lambda c: run_case(a.work, c), cases))

    path = os.path.join(HERE, "large.json.gz")
    gold = {}
    if a.only and os.path.exists(path):
        with gzip.open(path, "rt") as f:
            gold = json.load(f)
    sys.path.insert(0, HERE)
    from make_golden import complete_lists
    for (name, drv, mode, m, n, n_eq, seed, K), rec in zip(cases, raw):
        g = {"m": m, "n": n, "n_eq": n_eq, "seed": seed, "rule": drv, "mode": mode, "driver_status": rec["status"]}
        if mode == "synth":
            for c in rec["calls"]:
                complete_lists(c)
            g["status"] = rec["status"]
            g["objective"] = rec["objective"]
            g["x_struct"] = [rec["solution"]["x%07d" % j] for j in range(n)]
            g["calls"] = [{k: c[k] for k in ("M", "N", "iterations", "optimal", "unbounded", "pivots", "basic_in",
                                             "nonbasic_in", "basic_out", "nonbasic_out")} for c in rec["calls"]]
        else:
            c = rec["calls"][0]
            # a timing stop cuts the run inside iteration K+1, after its entering/leaving decision was
            # logged: every recorded pivot is a completed decision of the reference
            g["complete"] = rec["status"] != "timing_stop"
            g["optimal"] = bool(c["optimal"])
            g["iterations"] = c["iterations"]
            g["pivots"] = c["pivots"]
            bas, non = replay(m, n, c["pivots"])
            g["basic_after"] = bas
            ts = rec.get("iter_timestamps", [])
            if len(ts) > 1:
                g["ref_seconds_per_iteration"] = (ts[-1] - ts[0]) / (len(ts) - 1)
        gold[name] = g
        print(name, g.get("status", g.get("complete")), g.get("iterations", [c["iterations"] for c in g.get("calls", [])]),
              len(g.get("pivots", [])) or [len(c["pivots"]) for c in g["calls"]])
    with gzip.open(path, "wt") as f:
        json.dump(gold, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
