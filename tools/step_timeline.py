#!/usr/bin/env python
"""Where a bench step's time goes on the context stream: CUDA events between the C-ABI calls of GpuRig.step (bring-up aid).
usage: tools/step_timeline.py [workload] [host: 0|1]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
host = len(sys.argv) > 2 and sys.argv[2] == "1"
rig = bench.GpuRig(wl, 0, 0, pipeline=os.environ.get("QT_PIPELINE", "1") == "1")
rig.prime(host)
klt, ctx, t = rig.klt, rig.ctx, torch
marks = []

def mark(name, ev_list):
    e = t.cuda.Event(enable_timing=True); e.record(rig.stream); ev_list.append((name, e))

orig = {}
def wrap(obj, attr, name):
    f = getattr(obj, attr)
    def g(*a, **k):
        r = f(*a, **k); mark(name, cur[0]); return r
    setattr(obj, attr, g)

cur = [None]
wrap(rig, "_detect", "detect(call)")
wrap(rig, "_push", "push(call)")
wrap(ctx.lib, "psnw_klt_track_batch", "track") if False else None
_tb = ctx.lib.psnw_klt_track_batch
def tb(*a):
    r = _tb(*a); mark("track", cur[0]); return r
class LibProxy:
    def __init__(self, lib): self._lib = lib
    def __getattr__(self, n): return tb if n == "psnw_klt_track_batch" else getattr(self._lib, n)
ctx.lib = LibProxy(ctx.lib)
_bu = ctx.box_update_raw
def bu(*a):
    r = _bu(*a); mark("box_update", cur[0]); return r
ctx.box_update_raw = bu
_jn = ctx.join
def jn():
    r = _jn(); mark("join", cur[0]); return r
ctx.join = jn

acc = {}
all_evs = []
with t.cuda.stream(rig.stream):
    for i in range(60):
        rig.flush.zero_()
        evs = []; cur[0] = evs
        mark("start", evs)
        rig.step(host)
        mark("end", evs)
        if host: ctx.sync()
        if os.environ.get('QT_SYNC', '1') == '1': rig.stream.synchronize()
        all_evs.append(evs)

rig.stream.synchronize()
for evs in all_evs[10:]:
    for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
        acc.setdefault(n1, []).append(e0.elapsed_time(e1) * 1e3)
print(json.dumps({k: round(float(np.mean(v)), 1) for k, v in acc.items()}), "total_us", round(sum(float(np.mean(v)) for v in acc.values()), 1))
