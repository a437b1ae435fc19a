#!/usr/bin/env python
"""Executed warp-instructions per CUDA source line of an `ncu --set full --import-source on` capture, summed by line ranges.
usage: tools/ncu_sections.py <report.ncu-rep> [file:lo-hi=name ...]   (lines outside every range are listed per file)"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
ranges = []
for a in sys.argv[2:]:
    spec, name = a.split("=")
    f, lh = spec.split(":")
    lo, hi = lh.split("-")
    ranges.append((f, int(lo), int(hi), name))
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(out.splitlines()))
fname, hdr = None, None
acc, per_line = collections.Counter(), []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname, hdr = r[1].split("/")[-1], None
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    # a CUDA source line carries the sum over its SASS instructions; the SASS rows under it have an empty line number
    if hdr and len(r) >= len(hdr) - 1 and r[0].isdigit() and "Instructions Executed" in hdr:
        n = r[hdr.index("Instructions Executed")]
        n = int(n) if n.isdigit() else 0
        if not n:
            continue
        line = int(r[0])
        key = next((nm for (f, lo, hi, nm) in ranges if f == fname and lo <= line <= hi), fname)
        acc[key] += n
        per_line.append((n, fname, line, r[1].strip()[:100], key == fname))
tot = sum(acc.values())
print("total warp-instructions", tot)
for k, v in acc.most_common():
    print("%-28s %12d %5.1f%%" % (k, v, 100.0 * v / tot))
print("-- top lines")
for n, f, l, s, un in sorted(per_line, reverse=True)[:25]:
    print("%9d %5.1f%% %s:%d  %s" % (n, 100.0 * n / tot, f, l, s))
print("-- top lines outside every named range")
for n, f, l, s, un in [x for x in sorted(per_line, reverse=True) if x[4]][:25]:
    print("%9d %5.1f%% %s:%d  %s" % (n, 100.0 * n / tot, f, l, s))
