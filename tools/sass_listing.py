#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel mnemonic histogram (TMA = UTMALDG, mbarrier = SYNCS, warp reductions = REDUX,
packed dot products = IDP) and the full listing of the named kernels.  usage: tools/sass_listing.py  (after the build)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mcmtt_opticalflow_b200", "csrc")
WANT = {"lk.o": ["lk_w21_kernelILb0", "lk_generic_kernel"], "pyramid_fused.o": ["pyr_fused_kernelILi4ELi8ELi512", "pyr_fused_kernelILi3ELi16ELi256"],
        "pyramid.o": ["pyr_scharr_kernelILb1"], "detect.o": ["detect_select_kernel", "min_eig_kernel"], "box.o": ["box_update_kernel"]}
FULL = {"lk_w21_kernelILb0", "pyr_fused_kernelILi4ELi8ELi512", "pyr_scharr_kernelILb1"}
KEY = ["UTMALDG", "SYNCS", "REDUX", "IDP", "LDS", "STS", "LDG", "STG", "IMAD", "SHF", "PRMT", "ATOM", "BAR", "F2I", "I2F", "FMUL", "FADD", "FFMA", "DADD", "DMUL"]
for obj, names in WANT.items():
    txt = subprocess.check_output(["cuobjdump", "-sass", os.path.join(CSRC, obj)]).decode()
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    for f in funcs:
        fname = f.split("\n", 1)[0].strip()
        hit = [n for n in names if n in fname]
        if not hit:
            continue
        ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
        hist = collections.Counter(i.split(".")[0] for i in ins)
        demangled = subprocess.run(["c++filt", fname], capture_output=True, text=True).stdout.strip() or fname
        out = os.path.join(ROOT, "profiles", "sass_%s.txt" % re.sub(r"[^A-Za-z0-9]+", "_", hit[0]).strip("_"))
        with open(out, "w") as o:
            o.write("# %s\n# cuobjdump -sass mcmtt_opticalflow_b200/csrc/%s (sm_100a), %d instructions\n" % (demangled[:160], obj, len(ins)))
            o.write("# mnemonics of interest: " + ", ".join("%s %d" % (k, sum(v for m, v in hist.items() if m.startswith(k))) for k in KEY) + "\n")
            o.write("# all mnemonics: " + ", ".join("%s %d" % kv for kv in hist.most_common()) + "\n\n")
            if hit[0] in FULL:          # full listing (addresses + instructions, encodings stripped)
                for line in f.splitlines():
                    m = re.match(r"\s*(/\*[0-9a-f]{4}\*/\s+.*?;)", line)
                    if m:
                        o.write(m.group(1) + "\n")
                    elif not re.match(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", line) and line.strip() and "/*" not in line[:12]:
                        o.write(line.rstrip() + "\n")
        print(out, len(ins), {k: sum(v for m, v in hist.items() if m.startswith(k)) for k in ("UTMALDG", "SYNCS", "REDUX", "IDP")})
