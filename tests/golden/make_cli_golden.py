#!/usr/bin/env python3
"""Regenerates tests/golden/cli_outputs.json.gz (test infrastructure; build container only).

Runs the reference's stock command line (oracle/_ref/simplex_ref = /root/reference/src/*.cpp
compiled unmodified, see oracle/Makefile) on every .lp text of fixtures.json.gz and
parser_cases.json.gz and records what a user sees: exit status, whether sol_test.xml was
written and its text, and the stdout lines that are not log lines."""
import gzip
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CLI = os.path.join(ROOT, "oracle/_ref/simplex_ref")


def load(name):
    with gzip.open(os.path.join(HERE, name), "rt") as f:
        return json.load(f)


def main():
    texts, degenerate = {}, {}
    for k, v in load("fixtures.json.gz").items():
        texts["fixture:" + k] = v["lp_text"]
    for k, v in load("parser_cases.json.gz").items():
        texts["case:" + k] = v["lp_text"]
        degenerate["case:" + k] = bool(v.get("degenerate"))
    out = {}
    for name, text in texts.items():
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "in.lp")
            with open(path, "w") as f:
                f.write(text)
            p = subprocess.run([CLI, path], cwd=td, capture_output=True, text=True, timeout=60)
            xml_path = os.path.join(td, "sol_test.xml")
            xml = open(xml_path).read() if os.path.exists(xml_path) else None
            plain = [ln for ln in p.stdout.splitlines() if ln and not ln.startswith("[") and not ln.startswith("Value =")]
            out[name] = {"lp_text": text, "returncode": p.returncode, "xml": xml, "stdout_plain": plain,
                         "unbounded": "LP IS UNBOUNDED." in p.stdout, "degenerate": degenerate.get(name, False)}
            print(name, p.returncode, plain, None if xml is None else len(xml))
    with gzip.open(os.path.join(HERE, "cli_outputs.json.gz"), "wt") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
