#!/usr/bin/env python
"""Every kernel once on small shapes (for compute-sanitizer runs): push (gray + BGR), detect (sorted and fallback
paths), track (21x21 and generic windows, border points), box_update, overlapped detect."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmtt_opticalflow_b200 import klt, synth

w, h = 200, 150
cam = synth.Camera(w, h, 4, cam=3)
f0, f1 = cam.frame(0), cam.frame(1)
with klt.KLTContext(w, h, n_cameras=2, levels=3, fb_thresh=1.0) as c:
    c.push_frames([f0, f0]); c.push_frames([f1, f1])
    c.push_frames_bgr([np.stack([f1, f1, f1], -1)], cams=[1])
    boxes = np.array([[10, 10, 40, 96], [100, 20, 50, 110], [-5, -5, 30, 60], [0, 0, w, h]], np.float64)
    pts, cnt = c.detect(boxes, 50, 0.01, 3.0)
    pts2, cnt2 = c.detect([[0, 0, w, h]], 1500, 0.0001, 0.0)          # > kMaxAccepted -> global-list path
    rng = np.random.default_rng(0)
    p = np.concatenate([pts[0, :cnt[0]], rng.uniform(-30, 230, size=(200, 2)).astype(np.float32)])
    o, st, er, fb = c.track(p)
    tgt = (np.arange(len(p)) % 3).astype(np.int32)
    o2, st2, er2, fb2 = c.track(p, target_of_pt=tgt, win_wh=[[21, 21], [9, 15], [40, 100]])
    off = np.array([0, 60, 120, len(p)], np.int32)
    b, m, n = c.box_update(boxes[:3], p, o, off, status=st)
    print("sanity ok", cnt.tolist(), int(cnt2[0]), int(st.sum()), int(st2.sum()), n.tolist())
