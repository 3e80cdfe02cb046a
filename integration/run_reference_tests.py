"""Runs the reference's own unit tests (src/unitTests/*.cxx, compiled unmodified by integration/Makefile) one by one
on (a) the CUDA build — reference library + integration/blasLapackWrapper_xb200.cpp + libxb200.so, no OpenBLAS — and
(b) the control build on the reference's CPU wrapper, and prints the pass/fail table.

    make -C oracle && make -C integration          # in the container that has /root/reference
    python integration/run_reference_tests.py [--only-xb200] [--json out.json]

The list of tests is committed (integration/reference_unittests.txt) because the reference tree does not travel to the
GPU box; regenerate it with --list when the reference is present.
"""
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIST = os.path.join(HERE, "reference_unittests.txt")
FILES = ["fullTensor_product", "fullTensor_factorisations", "ttRounding", "ttCreation", "ttArithmetic", "ttOther", "als", "tensorNetwork"]


def make_list(ref="/root/reference"):
    names = []
    for f in FILES:
        src = open(os.path.join(ref, "src", "unitTests", f + ".cxx")).read()
        names += ["%s:%s" % m for m in re.findall(r'UnitTest\s+\w+\(\s*"([^"]+)"\s*,\s*"([^"]+)"', src)]
    with open(LIST, "w") as fh:
        fh.write("\n".join(names) + "\n")
    return names


def run(exe, name, timeout=300):
    try:
        p = subprocess.run([exe, name], capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return "timeout"
    out = p.stdout + p.stderr
    if p.returncode == 0 and "passed!" in out and "FAILED" not in out:
        return "pass"
    if "SuiteSparse" in out or "sparse" in out.lower() and "not available" in out:
        return "needs-sparse"
    return "FAIL"


def main():
    if "--list" in sys.argv:
        print(len(make_list()), "tests listed")
        return
    names = [l.strip() for l in open(LIST) if l.strip()]
    exes = {"xb200": os.path.join(HERE, "_build", "XerusTest_xb200")}
    if "--only-xb200" not in sys.argv:
        exes["reference_cpu"] = os.path.join(HERE, "_build", "XerusTest_ref")
    table = {}
    for n in names:
        table[n] = {k: run(e, n) for k, e in exes.items()}
        print("%-45s %s" % (n, "  ".join("%s=%s" % kv for kv in table[n].items())), flush=True)
    summary = {k: {s: sum(1 for r in table.values() if r[k] == s) for s in ["pass", "FAIL", "needs-sparse", "timeout"]} for k in exes}
    print(json.dumps(summary))
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as fh:
            json.dump({"summary": summary, "tests": table}, fh, indent=1)


if __name__ == "__main__":
    main()
