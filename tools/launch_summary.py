#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
    agg[name].append(v)
tot = sum(sum(v) for v in agg.values())
print("%-62s %5s %10s %12s %7s" % ("kernel", "n", "mean us", "total us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-62s %5d %10.2f %12.1f %6.1f%%" % (k[:62], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
