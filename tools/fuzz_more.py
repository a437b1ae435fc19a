#!/usr/bin/env python
"""Extended run of the stream-ordering fuzz test (tests/test_gpu_pipeline.py): many seeds, longer command sequences.
usage: tools/fuzz_more.py [first_seed] [n_seeds] [n_cmds]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_pipeline as T


class Env:
    def setenv(self, k, v):
        os.environ[k] = v


first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cmds = int(sys.argv[3]) if len(sys.argv) > 3 else 150
bad = 0
for seed in range(first, first + n):
    for mode in ("auto", "level", "hybrid", "walk"):
        a, pa = T._fuzz_run(Env(), seed, False, mode, n_cmds=cmds)
        b, pb = T._fuzz_run(Env(), seed, True, mode, n_cmds=cmds)
        ok = len(a) == len(b) and all(np.array_equal(x, y) for u, v in zip(a, b) for x, y in zip(u, v))
        ok = ok and all(np.array_equal(ia, ib) and np.array_equal(da, db) for p, q in zip(pa, pb) for (ia, da), (ib, db) in zip(p, q))
        bad += 0 if ok else 1
        if not ok:
            print("MISMATCH seed", seed, "mode", mode, flush=True)
print("fuzz: %d seeds x 4 builders x %d commands, %d mismatches" % (n, cmds, bad))
