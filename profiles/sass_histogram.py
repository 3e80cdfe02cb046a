"""SASS opcode histogram of xerus_b200/libxb200.so (runs here, no GPU):  python profiles/sass_histogram.py > profiles/r2_sass_histogram.txt
What proves what (B200_PROFILING.md): DMMA = the FP64 tensor pipe (mma.sync.m8n8k4.f64; FP64 has no tcgen05 kind, so no UTC*MMA / LDTM
is expected), LDGSTS = cp.async staging of the GEMM, UBLKCP / SYNCS / STAS / UCGABAR = DSMEM bulk copies, mbarrier transactions,
st.async and cluster barriers of the cluster kernels (QR panels, Jacobi hand-over, CG), UTMALDG / UTMASTG = TMA tensor copies."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "xerus_b200/libxb200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
per_kernel = collections.defaultdict(collections.Counter)
total = collections.Counter()
kernel = None
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        kernel = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kernel:
        per_kernel[kernel][m.group(1)] += 1
        total[m.group(1)] += 1
print("# %s: %d kernels, %d SASS instructions" % (lib, len(per_kernel), sum(total.values())))
print("# opcode totals")
for op, n in total.most_common():
    print("%8d  %s" % (n, op))
marks = ["DMMA", "LDGSTS", "UBLKCP", "SYNCS", "STAS", "UCGABAR", "UTMALDG", "UTMASTG", "UTCHMMA", "LDTM"]
print("# kernels using the marker opcodes (%s)" % ", ".join(marks))
for k in sorted(per_kernel):
    hits = ["%s=%d" % (m, per_kernel[k][m]) for m in marks if per_kernel[k][m]]
    if hits:
        print("%-70s %s" % (k[:70], " ".join(hits)))
