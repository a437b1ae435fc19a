#!/usr/bin/env python
"""Every kernel once on small shapes (for compute-sanitizer runs): push (gray + BGR, per-level and fused pyramid builders,
overlapped with the eager Shi-Tomasi map), detect (sorted and fallback paths, GridFAST), track (21x21 persistent kernel
with TMA-staged patches and the split tail, generic windows, border points, apron 26 and 64), box_update, track_update,
points_in_boxes, overlapped detect.  A quick smoke run of every launch configuration; usable under
compute-sanitizer where that tool is available (it is closed on the B200 pool this round was measured on)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PSNW_LK_TAIL", "60")            # the split tail with a batch this small
import numpy as np
import torch
from mcmtt_opticalflow_b200 import klt, synth

w, h = 200, 150
cam = synth.Camera(w, h, 4, cam=3)
f0, f1, f2 = cam.frame(0), cam.frame(1), cam.frame(2)
rng = np.random.default_rng(0)
for apron, mode in ((0, "level"), (64, "fused")):
    os.environ["PSNW_PYR_MODE"] = mode
    with klt.KLTContext(w, h, n_cameras=2, levels=3, fb_thresh=1.0, apron=apron) as c:
        c.push_frames([f0, f0]); c.push_frames([f1, f1])
        c.push_frames_bgr([np.stack([f1, f1, f1], -1)], cams=[1])
        boxes = np.array([[10, 10, 40, 96], [100, 20, 50, 110], [-5, -5, 30, 60], [0, 0, w, h]], np.float64)
        pts, cnt = c.detect(boxes, 50, 0.01, 3.0)
        pts2, cnt2 = c.detect([[0, 0, w, h]], 1500, 0.0001, 0.0)          # > kMaxAccepted -> global-list path
        gp, gc = c.gridfast_detect(boxes[:2], 50)
        p = np.concatenate([pts[0, :cnt[0]], rng.uniform(-30, 230, size=(200, 2)).astype(np.float32)])
        o, st, er, fb = c.track(p)
        tgt = (np.arange(len(p)) % 3).astype(np.int32)
        o2, st2, er2, fb2 = c.track(p, target_of_pt=tgt, win_wh=[[21, 21], [9, 15], [40, 100]])
        off = np.array([0, 60, 120, len(p)], np.int32)
        b, m, n = c.box_update(boxes[:3], p, o, off, status=st)
        tu = c.track_update(p, off, boxes[:3])
        pib = c.points_in_boxes(o, off, boxes)
        img, a = c.read_level_apron(0, 0, 1)
        # one frame in flight: overlapped detect, track, overlapped push with the eager map, join
        pin = torch.from_numpy(np.stack([f2, f2])).pin_memory()
        dev = torch.device("cuda", 0)
        d_p = torch.from_numpy(p).to(dev)
        d_o = torch.zeros((len(p), 2), dtype=torch.float32, device=dev); d_s = torch.zeros(len(p), dtype=torch.uint8, device=dev)
        d_e = torch.zeros(len(p), dtype=torch.float32, device=dev); d_f = torch.zeros(len(p), dtype=torch.float32, device=dev)
        d_seeds = torch.zeros((4, 50, 2), dtype=torch.float32, device=dev); d_cnt = torch.zeros(4, dtype=torch.int32, device=dev)
        job = (klt.DetectJob * 1)()
        job[0].cam, job[0].slot, job[0].boxes, job[0].n_targets = 0, 0, boxes.ctypes.data, 4
        job[0].pts_xy, job[0].counts = d_seeds.data_ptr(), d_cnt.data_ptr()
        for k in range(3):
            c._ck(c.lib.psnw_klt_detect_batch(c._h, job, 1, 50, 0.01, 3.0, klt.MEM_DEVICE | klt.MEM_OVERLAP))
            c.track_batch([dict(cam=0, slot_from=1, slot_to=0, n_pts=len(p), pts=d_p.data_ptr(), out=d_o.data_ptr(),
                                status=d_s.data_ptr(), err=d_e.data_ptr(), fb_err=d_f.data_ptr())], mem=klt.MEM_DEVICE)
            c.push_frames([pin[0].data_ptr(), pin[1].data_ptr()], [0, 1], 3 + k, klt.MEM_PINNED | klt.MEM_OVERLAP | klt.PUSH_EIG, pitch=w)
            c.join()
        c.sync()
        assert np.array_equal(d_s.cpu().numpy()[:10], d_s.cpu().numpy()[:10])
        print("sanity ok", apron, mode, cnt.tolist(), int(cnt2[0]), gc.tolist(), int(st.sum()), int(st2.sum()), n.tolist(), a, int(pib.sum()))
