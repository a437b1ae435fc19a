#!/usr/bin/env python
"""Frames/s of the drop-in C++ CPSNWhere_Tracker2D mirror (reference call pattern: per-target windows (w,w)/(w,h),
default 4 levels, 3-step backward chain + forward tracking + matching) on one 768x576 camera with 40 detections,
next to the same pattern on the CPU (tools/cpu_variants.py).  Wall clock per Run(), host logic included."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmtt_opticalflow_b200 import klt, synth
from integration import tracker2d

w, h, ntgt, T = 768, 576, 40, 24
cam = synth.Camera(w, h, ntgt, cam=0)
frames = [cam.frame(t) for t in range(T)]
boxes = [cam.boxes(t) for t in range(T)]
with klt.KLTContext(w, h, levels=4, fb_thresh=-1.0, max_points=1 << 16, max_targets=1 << 10,
                    apron=int(os.environ.get("PSNW_APRON", "64"))) as ctx:        # per-target windows up to ~55 x 130: see psnw_params.apron
    trk = tracker2d.Tracker2D(ctx)
    ts, nobj = [], []
    for t in range(T):
        t0 = time.perf_counter()
        ids, ob, nf = trk.run(boxes[t], frames[t], t)
        ts.append(time.perf_counter() - t0)
        nobj.append(len(ids))
    trk.close()
ts = np.array(ts[6:])
print(json.dumps({"workload": "reference call pattern, 1 camera 768x576, %d detections/frame" % ntgt, "frames": int(len(ts)),
                  "ms_per_frame_median": round(1e3 * float(np.median(ts)), 3), "frames_per_s": round(1.0 / float(np.median(ts)), 1),
                  "objects_per_frame": int(np.median(nobj))}))
