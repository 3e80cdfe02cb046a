"""TEST INFRASTRUCTURE ONLY.

Golden fixtures of the *unmodified* reference at the reduced sizes SURVEY.md 8d names for the configs the reference
cannot run at full size:

    make -C oracle
    python oracle/make_golden_sizes.py          # ~5 minutes of CPU, several GB of RAM for the r = 32 DMRG case

  c2_r8 / c2_r20   ALS_SPD(Laplace d=16 n=10, x0 = random rank r, b = ones, 2 half-sweeps)      (BASELINE configs[1], reduced rank)
  c4_r8/16/32      DMRG_SPD(Laplace d=10 n=4, x0 = random rank r, b = ones, 1 half-sweep)       (BASELINE configs[3], reduced bond;
                                                                                                 the reference's two-site driver
                                                                                                 only survives one half-sweep)
Inputs (A, b, x0), the reference's result x, its ranks and the energy it returns go to tests/golden/xerus_ref_sizes_v1.npz.
The full-size configs 3 and 5 are checked against `oracle/_ref/ref_bench ... dump` run at test time (the binary travels).
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from make_golden import read_container  # noqa: E402

CASES = [("c2_r8", ["als", "16", "10", "8", "2", "1"]), ("c2_r20", ["als", "16", "10", "20", "2", "1"]),
         ("c4_r8", ["dmrg", "10", "4", "8", "1", "1"]), ("c4_r16", ["dmrg", "10", "4", "16", "1", "1"]),
         ("c4_r32", ["dmrg", "10", "4", "32", "1", "1"])]


def main():
    exe = os.path.join(HERE, "_ref", "ref_bench")
    if not os.path.exists(exe):
        sys.exit("build it first: make -C oracle")
    rec = {}
    only = sys.argv[1:]
    dst = os.path.join(ROOT, "tests", "golden", "xerus_ref_sizes_v1.npz")
    if only and os.path.exists(dst):
        rec = dict(np.load(dst))
    with tempfile.TemporaryDirectory() as td:
        for tag, argv in CASES:
            if only and tag not in only:
                continue
            dump = os.path.join(td, tag + ".bin")
            out = subprocess.run([exe] + argv + [dump], check=True, capture_output=True, text=True).stdout
            print(tag, out.strip())
            for k, v in read_container(dump).items():
                if k.startswith("A.") or k.startswith("b."):
                    continue                      # the Laplace-like operator and b = ones are rebuilt by the tests
                rec[tag + "." + k] = v
    np.savez_compressed(dst, **rec)
    print("wrote %d records (%.1f KB raw) to %s" % (len(rec), sum(v.nbytes for v in rec.values()) / 1024, dst))


if __name__ == "__main__":
    main()
