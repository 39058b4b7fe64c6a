"""Differential fuzz of the host side against the reference itself (CPU only): seeded random .lp
texts are read by this repo's reader + presolve + tableau builder and by the UNMODIFIED reference
(oracle/_ref/ref_driver_stock, compiled from /root/reference/src by oracle/Makefile). Runs where that
binary exists (the build container; it also travels to the GPU box) and is skipped elsewhere — the
committed golden vectors (tests/test_host_model.py) cover the same ground without it."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import call_inputs
from lp_text_gen import random_lp_text
from simplex_b200 import lp as hostlp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "oracle", "_ref", "ref_driver_stock")


def run_reference(text, td):
    path = os.path.join(td, "in.lp")
    with open(path, "w") as f:
        f.write(text)
    out = os.path.join(td, "out.json")
    if os.path.exists(out):
        os.remove(out)
    subprocess.run([DRV, "lp", path, "--out", out, "--dump-tableau"], cwd=td, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL, timeout=30)
    if not os.path.exists(out):
        return None
    raw = open(out).read()
    raw = re.sub(r'(?<![\w"])-?nan(?![\w"])', "NaN", raw)
    raw = re.sub(r'(?<![\w"])(-?)inf(?![\w"])', r"\1Infinity", raw)
    return json.loads(raw)


@pytest.mark.skipif(not os.path.exists(DRV), reason="reference binary not built (make -C oracle ref)")
def test_random_texts_give_the_reference_tableau(tmp_path):
    checked = 0
    for seed in range(1000, 1160):
        text = random_lp_text(seed, m=None if seed % 5 else 14, n=None if seed % 5 else 20)
        rec = run_reference(text, str(tmp_path))
        with hostlp.LpSession() as s:
            rc = s.read(text, check=False)
            if rec is None:                      # the reference aborted while reading
                assert rc == hostlp.READ_FATAL, (seed, rc)
                continue
            if rec["status"] == "exception_string":
                assert rc == hostlp.READ_EXCEPTION and s.error() == "Exception: " + rec["exception"], seed
                continue
            if rec["status"] == "exit_called":
                assert rc == hostlp.READ_PRESOLVE_INFEASIBLE, seed
                continue
            if not rec["calls"]:
                assert rc == hostlp.READ_FATAL, (seed, rec["status"])
                continue
            assert rc == hostlp.READ_OK, (seed, s.error())
            call = rec["calls"][0]
            n_art = s.open_phase_one()
            if n_art == 0:
                s.initialize_tableau()
            A, b, c, ub = s.tableau()
            first = list(call.get("basic_in") or call["first_basis"])
            call.setdefault("basic_in", first)
            call.setdefault("nonbasic_in", [j for j in range(call["N"]) if j not in set(first)])
            Ar, br, cr, ubr, _, non_r = call_inputs(call)
            assert A.shape == Ar.shape, seed
            for got, want in ((A, Ar), (b, br), (c, cr), (ub, ubr)):
                assert np.array_equal(got, want, equal_nan=True), seed
            basic, non_basic = s.basis()
            assert basic.tolist() == first and non_basic.tolist() == non_r.tolist(), seed
            checked += 1
    assert checked >= 120
