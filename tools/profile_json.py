#!/usr/bin/env python
"""Turns the ncu captures of tools/r2_profiles.sh into the two small JSON files bench.py quotes (roofline.traffic, lk.ncu).
Each file carries a hash of the kernel sources it was captured from; bench.py ignores it when the sources have changed since.
usage: tools/profile_json.py <pyr_4frames.ncu-rep> <lk_w21.ncu-rep> [nominal solves of the LK launch = 32000]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def rows(rep):
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    return [dict(zip(hdr, x)) for x in r[2:]], dict(zip(hdr, units))


def num(d, k):
    return float(d[k].replace(",", ""))


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


pyr_rep, lk_rep = sys.argv[1], sys.argv[2]
solves = int(sys.argv[3]) if len(sys.argv) > 3 else 32000
pr, pu = rows(pyr_rep)
names = [r["Kernel Name"][:60] for r in pr]
rd = [to_bytes(num(r, "dram__bytes_read.sum"), pu["dram__bytes_read.sum"]) for r in pr]
wr = [to_bytes(num(r, "dram__bytes_write.sum"), pu["dram__bytes_write.sum"]) for r in pr]
json.dump({"source": "ncu --set full --clock-control none of the c2 pyramid launch group (4 frames 768x576, 4 levels; tools/r2_profiles.sh, "
                     "summary in profiles/r02_ncu_pyr_4.txt): " + " + ".join(names),
           "sources_sha256": bench.sources_hash(bench.PYR_SOURCES),
           "per_launch_dram_bytes_read": rd, "per_launch_dram_bytes_write": wr, "traffic_bytes_per_group": sum(rd) + sum(wr),
           "note": "dram__bytes_write is ~0 inside the launch: the 10.5 MB of outputs stay in the 126 MB L2 and are written back later; "
                   "the reads are the 1.77 MB of input frames"},
          open(os.path.join(ROOT, "profiles", "r02_pyr_traffic.json"), "w"), indent=1)
lr, lu = rows(lk_rep)
k = [r for r in lr if "lk_w21" in r["Kernel Name"]][0]
stall = {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: round(num(k, h), 2)
         for h in k if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and k[h]}
top = dict(sorted(stall.items(), key=lambda kv: -kv[1])[:7])
inst = num(k, "smsp__inst_executed.sum")
json.dump({"source": "ncu --set full --clock-control none of one lk_w21_kernel launch of the c2 batch (tools/quick_time.py; summary in "
                     "profiles/r02_ncu_lk_w21.txt)",
           "sources_sha256": bench.sources_hash(bench.LK_SOURCES),
           "warp_instructions_per_launch": inst, "nominal_solves_per_launch": solves, "warp_instructions_per_solve": inst / solves,
           "issue_slots_active_pct_of_active_cycles": round(num(k, "smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
           "sm_active_pct_of_elapsed": round(100.0 * num(k, "smsp__cycles_active.avg") / num(k, "sm__cycles_elapsed.avg"), 1),
           "registers": int(num(k, "launch__registers_per_thread")),
           "shared_bank_conflicts_per_launch": num(k, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
           "dram_bytes_read_per_launch": to_bytes(num(k, "dram__bytes_read.sum"), lu["dram__bytes_read.sum"]),
           "pipe_util_pct": {p: round(num(k, "sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active" % p), 1) for p in ("alu", "fma", "lsu", "xu", "adu")},
           "stalls_per_issued_instruction": top},
          open(os.path.join(ROOT, "profiles", "r02_lk_issue.json"), "w"), indent=1)
print("wrote profiles/r02_pyr_traffic.json, profiles/r02_lk_issue.json")
