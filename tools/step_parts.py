#!/usr/bin/env python
"""Which part of a bench step is on the critical path: the same step with parts left out (bring-up aid).
usage: tools/step_parts.py [workload]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
rig = bench.GpuRig(wl, 0, 0)
rig.prime(False)
for _ in range(5):
    rig.step(False)
klt, ctx, t = rig.klt, rig.ctx, torch
seeds = rig.d_seeds[rig.which]
jobs = rig._jobs(seeds.data_ptr(), rig.d_out.data_ptr(), rig.d_status.data_ptr(), rig.d_err.data_ptr(), rig.d_fb.data_ptr(), 1, 0)

def track():
    ctx._ck(ctx.lib.psnw_klt_track_batch(ctx._h, jobs, rig.ncam, klt.MEM_DEVICE))
def box():
    ctx.box_update_raw(rig.d_boxes_in.data_ptr(), rig.ntargets, seeds.data_ptr(), rig.d_out.data_ptr(), rig.d_status.data_ptr(),
                       rig.d_off.data_ptr(), rig.d_boxes_out.data_ptr(), rig.d_mask.data_ptr(), rig.d_counts.data_ptr(), klt.MEM_DEVICE)
def detect():
    rig._detect(bench.pingpong(rig.step_idx), rig.d_seeds[rig.which ^ 1], rig.d_counts_det, klt.MEM_DEVICE | klt.MEM_OVERLAP, slot=1)
def rebuild():
    ctx.rebuild_pyramids(rig.cams)

variants = {
    "track": [track],
    "track+box": [track, box],
    "detect|track+box+join": [detect, track, box, ctx.join],
    "box": [box],
    "track+box+rebuild(main)": [track, box, rebuild],
}
res = {}
with t.cuda.stream(rig.stream):
    for name, seq in variants.items():
        ev = [(t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)) for _ in range(30)]
        for i in range(33):
            rig.flush.zero_()
            if i >= 3: ev[i - 3][0].record(rig.stream)
            for f in seq: f()
            if i >= 3: ev[i - 3][1].record(rig.stream)
        rig.stream.synchronize()
        res[name] = round(float(np.mean([a.elapsed_time(b) for a, b in ev])) * 1e3, 1)
print(json.dumps(res))
