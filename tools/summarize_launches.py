"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step).
usage: python tools/summarize_launches.py <csv> [last_n_launches]"""
import collections, csv, re, sys
path = sys.argv[1]
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith("==")]
durs = [(r["Kernel Name"], float(r["Metric Value"]), r.get("Grid Size", ""), r.get("Block Size", "")) for r in csv.DictReader(lines)
        if r["Metric Name"] == "gpu__time_duration.sum"]
if last:
    durs = durs[-last:]
tot = sum(d for _, d, _, _ in durs)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, d, _, _ in durs:
    k = re.sub(r"\(.*", "", k)
    agg[k][0] += 1; agg[k][1] += d
print("launches %d, sum of kernel durations %.3f ms (cold-cache, serialised under ncu: compare SHARES)" % (len(durs), tot / 1e6))
for k, (n, d) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-58s n=%4d  %8.3f ms  %5.1f%%  avg %7.1f us" % (k[:58], n, d / 1e6, 100 * d / tot, d / n / 1e3))
if "--top" in sys.argv:
    print("--- 25 slowest launches")
    for k, d, g, b in sorted(durs, key=lambda x: -x[1])[:25]:
        print("%8.1f us  grid %-14s %s" % (d / 1e3, g, re.sub(r"\(.*", "", k)[:60]))
