#!/usr/bin/env python
"""Join an ncu SASS source page with nvdisasm line info: per-source-line executed-instruction counts.
usage: tools/ncu_lines.py <report.ncu-rep> <object.o> <kernel-substring> [per_unit_divisor]"""
import csv, io, os, re, subprocess, sys, collections
rep, obj, kern = sys.argv[1:4]
div = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
sass = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + os.environ.get("NCU_K", kern)], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(io.StringIO(sass)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
inst = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[ix["Address"]].startswith("0x"):
        break
    inst.append((int(r[ix["Address"]], 16), r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)))
base = inst[0][0]
dis = subprocess.check_output(["nvdisasm", "-g", "-c", obj]).decode().splitlines()
# walk functions; pick the one whose name contains kern
line_of = {}
cur_fn, cur_line, active = None, None, False
for l in dis:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        active = kern in m.group(1); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m and active:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and active:
        line_of[int(m.group(1), 16)] = cur_line
agg = collections.Counter(); samp = collections.Counter(); tot = 0
for a, s, n, sm in inst:
    ln = line_of.get(a - base)
    key = ("%s:%d" % (ln[0], ln[1])) if ln else "?"
    agg[key] += n; samp[key] += sm; tot += n
src = {}
print("total instructions %d (%.1f per unit)" % (tot, tot / div))
stot = sum(samp.values()) or 1
order = samp.most_common(40) if len(sys.argv) > 5 and sys.argv[5] == "samples" else agg.most_common(40)
for k, _n in order:
    n = agg[k]
    f, _, ln = k.partition(":")
    text = ""
    try:
        if f not in src:
            import glob
            cand = glob.glob("/root/repo/**/" + f, recursive=True)
            src[f] = open(cand[0]).read().splitlines() if cand else []
        text = src[f][int(ln) - 1].strip()[:90]
    except Exception:
        pass
    print("%9.1f %5.1f%% samp %5.1f%% | %-16s %s" % (n / div, 100.0 * n / tot, 100.0 * samp[k] / stot, k, text))
