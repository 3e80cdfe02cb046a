"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv   > profiles/r1_launches_c3.txt
    python profiles/summarize.py kernel   gpurun_out/prof_jacobi_r1.ncu-rep > profiles/r1_jacobi_persistent.txt
    python profiles/summarize.py source   gpurun_out/prof_jacobi_split_r1b.ncu-rep [min_share] > profiles/r1_jacobi_split_source.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1e-3)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)")
    print("# %d launches, %.3f ms of kernel time" % (sum(v[0] for v in agg.values()), tot / 1e3))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s n=%6d total=%9.3f ms avg=%9.2f us share=%5.1f%%" % (k[:58], v[0], v[1] / 1e3, v[1] / v[0], 100 * v[1] / tot))


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("# kernel:", d.get("Kernel Name"), " grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
        for k in KEYS:
            if k in d:
                print("%-85s %s %s" % (k, d[k], units[hdr.index(k)]))


def source(path, min_share=0.004):
    """Warp-stall samples per SASS instruction (ncu --set full --import-source on): totals per stall reason and every
    instruction that holds at least min_share of the samples, in program order."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print("# kernel:", rows[0][1][:150])
    hdr, data = rows[1], rows[2:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp]) for r in data)
    agg = collections.Counter()
    for r in data:
        for i, h in stalls:
            agg[h] += int(r[i])
    print("# %d SASS instructions, %d warp-stall samples" % (len(data), tot))
    print("# samples by stall reason:", ", ".join("%s %.1f%%" % (h, 100.0 * v / max(tot, 1)) for h, v in agg.most_common(8)))
    print("# idx  instruction                                                     samples  share  executed  top stall reasons")
    for k, r in enumerate(data):
        n = int(r[isamp])
        if n >= tot * float(min_share):
            top = sorted(((int(r[i]), h) for i, h in stalls), reverse=True)[:2]
            print("%5d  %-64s %7d %5.1f%% %9s  %s" % (k, r[isrc].strip()[:64], n, 100.0 * n / tot, r[iex],
                                                       ", ".join("%s %d" % (h, v) for v, h in top if v)))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "source": source}[sys.argv[1]](*sys.argv[2:])
