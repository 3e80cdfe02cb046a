"""python profiles/summarize_dram.py <ncu csv with dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum> "<header line>" """
import collections, csv, re, sys
rows = list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")]))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    if name.startswith("void at::"):
        continue                                  # torch's L2 flush fill between timed iterations is not part of the round
    v = float(r["Metric Value"].replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1)
    a = agg[name]
    if r["Metric Name"] == "dram__bytes_read.sum": a[1] += v * mult
    elif r["Metric Name"] == "dram__bytes_write.sum": a[2] += v * mult
    elif r["Metric Name"] == "gpu__time_duration.sum": a[3] += v * mult; a[0] += 1
tr, tw, tt, n = (sum(a[i] for a in agg.values()) for i in (1, 2, 3, 0))
print("# " + sys.argv[2])
print("# %d kernel launches of xb200: DRAM read %.1f MB + written %.1f MB = %.1f MB; kernel time under ncu %.1f ms" % (n, tr / 1e6, tw / 1e6, (tr + tw) / 1e6, tt / 1e3))
print("# algorithmic minimum of one round (SURVEY 8d: every core read and written once per sweep, two sweeps): 72.7 MB; cores in 18.2 MB, cores out 5.1 MB")
for name, a in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
    print("%-52s n=%5d read=%9.2f MB write=%9.2f MB time=%8.2f ms" % (name[:52], a[0], a[1] / 1e6, a[2] / 1e6, a[3] / 1e3))
