#!/usr/bin/env python3
"""SASS evidence of the built library (runs without a GPU): per kernel, how many instructions of the kinds the
design relies on, and the first occurrences verbatim. usage: python scripts/sass_excerpt.py > profiles/<tag>_sass_excerpt.txt

    UBLKCP   cp.async.bulk global -> shared (the TMA engine's 1-D bulk copy: staged vectors y, a_q, pivot row)
    UBLKPF   cp.async.bulk.prefetch.L2 (columns / rows pulled into L2 behind the grid barriers)
    SYNCS    mbarrier arrive / try_wait (completion of the bulk copies)
    REDUX    redux.sync arg-min reductions (CREDUX.MIN on sm_100a)
    LDG.E.EF / .128   ld.global.cs.v2.f64: 128-bit streaming loads of the pricing sweep
    LDG/STG ... with a cache-policy register (UR/R desc): ld/st.global.L2::cache_hint of the fused pass (evict-last rows)
    DFMA / DADD / MUFU.RCP64H: fp64 arithmetic (no tensor-core instruction anywhere: no HMMA/UTCMMA/...)
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "simplex_b200", "libsimplex_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
head = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
print("# cuobjdump -sass simplex_b200/libsimplex_b200.so   (%s)" % ", ".join(l.strip() for l in head.splitlines() if "sm_" in l))
kinds = [("UBLKCP", r"\bUBLKCP"), ("UBLKPF", r"\bUBLKPF"), ("SYNCS", r"\bSYNCS"), ("REDUX", r"\bC?REDUX"),
         ("LDG.E.EF.128", r"\bLDG\.E\.EF\.128"), ("LDG.128", r"\bLDG\.E[.\w]*\.128"), ("STG.128", r"\bSTG\.E[.\w]*\.128"),
         ("LDS.128", r"\bLDS\.128"), ("DFMA", r"\bDFMA"), ("MUFU.RCP64H", r"\bMUFU\.RCP64H"), ("ATOM/RED", r"\b(ATOMG|REDG|ATOM|RED)\b"),
         ("tensor-core", r"\b(HMMA|IMMA|DMMA|UTCMMA|UTCHMMA|QGMMA|HGMMA)")]
cur, body = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        body[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        body[cur].append(line.strip())
demangle = subprocess.run(["c++filt"] + list(body), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(body, demangle))
print("\n%-58s %6s " % ("kernel", "instr") + " ".join("%12s" % k for k, _ in kinds))
for fn, lines in body.items():
    nm = re.sub(r"\(.*", "", names.get(fn, fn))
    print("%-58s %6d " % (nm[:58], len(lines)) + " ".join("%12d" % sum(1 for l in lines if re.search(rx, l)) for _, rx in kinds))
print("\n# first occurrences, verbatim")
for fn in body:
    nm = re.sub(r"\(.*", "", names.get(fn, fn))
    if not any(k in nm for k in ("k_persistent_fused", "k_sharded_fused", "k_resident", "k_batched_fast<true, true>")):
        continue
    print("\n## %s" % nm)
    for label, rx in (("UBLKCP", r"\bUBLKCP"), ("UBLKPF", r"\bUBLKPF"), ("SYNCS", r"\bSYNCS"), ("REDUX", r"\bC?REDUX"),
                      ("LDG.E.EF.128", r"\bLDG\.E\.EF\.128"), ("STG.128", r"\bSTG\.E[.\w]*\.128"), ("DFMA", r"\bDFMA")):
        hits = [l for l in body[fn] if re.search(rx, l)]
        for l in hits[:2]:
            print("   %-13s %s" % (label, re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l)))
