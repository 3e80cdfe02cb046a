"""TEST INFRASTRUCTURE ONLY.

Regenerates the committed golden fixtures from the *unmodified* reference:

    make -C oracle            # builds oracle/_ref/ref_golden from /root/reference (needs the reference tree)
    python oracle/make_golden.py

`oracle/_ref/ref_golden` dumps inputs + reference outputs (seed 0xBAADF00D) in a tiny record container; this
script converts the dump to `tests/golden/xerus_ref_v1.npz`.  The fixtures travel to the GPU box, the reference
tree does not.
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def read_container(path):
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"XBGOLD01", "bad magic"
        while True:
            hdr = f.read(8)
            if not hdr:
                break
            (ln,) = struct.unpack("<Q", hdr)
            name = f.read(ln).decode()
            (nd,) = struct.unpack("<Q", f.read(8))
            dims = struct.unpack("<%dQ" % nd, f.read(8 * nd)) if nd else ()
            n = int(np.prod(dims)) if nd else 1
            data = np.frombuffer(f.read(8 * n), dtype="<f8").copy()
            out[name] = data.reshape(dims) if nd else data.reshape(())
    return out


def main():
    exe = os.path.join(HERE, "_ref", "ref_golden")
    if not os.path.exists(exe):
        sys.exit("build it first: make -C oracle")
    dst = os.path.join(ROOT, "tests", "golden")
    os.makedirs(dst, exist_ok=True)
    # `python oracle/make_golden.py v2` regenerates only the second file (soft_threshold, operator TT-SVD, ASD)
    for name, extra in [("xerus_ref_v1.npz", []), ("xerus_ref_v2.npz", ["v2"])]:
        if sys.argv[1:] and sys.argv[1:] != extra:
            continue
        with tempfile.TemporaryDirectory() as td:
            dump = os.path.join(td, "golden.bin")
            subprocess.check_call([exe, dump] + extra)
            rec = read_container(dump)
        np.savez_compressed(os.path.join(dst, name), **rec)
        print("wrote %d records (%.1f KB raw) to tests/golden/%s" % (len(rec), sum(v.nbytes for v in rec.values()) / 1024, name))


if __name__ == "__main__":
    main()
