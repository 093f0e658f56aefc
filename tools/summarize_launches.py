"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print("total %.3f ms over %d launches (ncu per-launch times are cold-cache and serialised: compare shares)" % (tot / 1e3, sum(a[0] for a in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%9.3f ms %5.1f%% n=%4d  %s" % (t / 1e3, 100 * t / tot, n, k))
