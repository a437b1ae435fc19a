#!/usr/bin/env python
"""BASELINE config C5: throughput against frames in flight per camera (1 / 4 / 16 / 64) on one GPU.
F frames in flight = F consecutive rig frames are pushed (one pyramid launch group per rig frame) before their F frame
pairs are tracked in ONE launch of cameras x F jobs (forward + backward LK + FB gate, ring of F + 1 slots); seeds are a
fixed set per camera.  Reports tracked points/s and frames/s, device-resident, L2 flushed between
repetitions.  usage: tools/frames_in_flight.py [WxH ...]   (default 768x576 1920x1080 3840x2160)"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mcmtt_opticalflow_b200 import klt, synth

SHAPES = {"768x576": (768, 576, 4, 4, 4000), "1920x1080": (1920, 1080, 4, 1, 51200), "3840x2160": (3840, 2160, 5, 1, 129600)}
#          name: (w, h, levels, cameras, seeds per camera-frame)


def run(name, flights):
    w, h, levels, ncam, npts = SHAPES[name]
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream()
    rng = np.random.default_rng(1)
    cams = [synth.Camera(w, h, 20, cam=c) for c in range(ncam)]
    base = [[torch.from_numpy(c.frame(k)).to(dev) for k in range(3)] for c in cams]          # 3 distinct frames per camera, cycled
    seeds = torch.from_numpy(rng.uniform(12, [w - 12, h - 12], size=(npts, 2)).astype(np.float32)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    for F in flights:
        n_jobs = ncam * F
        total = n_jobs * npts
        with klt.KLTContext(w, h, n_cameras=ncam, levels=levels, ring_slots=F + 1, fb_thresh=1.0, max_points=max(total, 1024),
                            max_targets=max(n_jobs, 64), stream=stream.cuda_stream) as ctx:
            o = torch.zeros((total, 2), dtype=torch.float32, device=dev)
            st = torch.zeros(total, dtype=torch.uint8, device=dev)
            er = torch.zeros(total, dtype=torch.float32, device=dev)
            fb = torch.zeros(total, dtype=torch.float32, device=dev)
            jobs = (klt.TrackJob * n_jobs)()
            for c in range(ncam):
                for f in range(F):
                    a = jobs[c * F + f]
                    i = (c * F + f) * npts
                    a.cam, a.slot_from, a.slot_to, a.n_pts = c, f + 1, f, npts          # frame pair (f+1 frames ago -> f frames ago)
                    a.pts_xy, a.target_of_pt, a.n_targets, a.win_wh = seeds.data_ptr(), None, 1, None
                    a.out_xy, a.status, a.err, a.fb_err = o.data_ptr() + i * 8, st.data_ptr() + i, er.data_ptr() + i * 4, fb.data_ptr() + i * 4
            cam_ids = np.arange(ncam, dtype=np.int32)

            def push(k):
                ptrs = (C.c_void_p * ncam)(*[base[c][k % 3].data_ptr() for c in range(ncam)])
                ctx._ck(ctx.lib.psnw_push_frames(ctx._h, ncam, cam_ids.ctypes.data, ptrs, w, k, klt.MEM_DEVICE))

            with torch.cuda.stream(stream):
                for k in range(F + 1):
                    push(k)
                reps = 8 if F >= 16 else 20
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
                k = F + 1
                for r in range(reps + 2):
                    flush.zero_()
                    if r >= 2:
                        ev[r - 2][0].record(stream)
                    for _ in range(F):                       # F new frames per camera ...
                        push(k); k += 1
                    ctx._ck(ctx.lib.psnw_klt_track_batch(ctx._h, jobs, n_jobs, klt.MEM_DEVICE))      # ... and their F pairs, one launch
                    if r >= 2:
                        ev[r - 2][1].record(stream)
                stream.synchronize()
            ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
            row = {"frame": name, "cameras": ncam, "frames_in_flight": F, "points_per_launch": total, "ms": round(ms, 4),
                   "tracked_points_per_s": round(total / ms * 1e3), "camera_frames_per_s": round(n_jobs / ms * 1e3, 1),
                   "status_ok": round(float(st.float().mean().item()), 4)}
            print(json.dumps(row), flush=True)
            out.append(row)
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or list(SHAPES)
    res = []
    for n in names:
        res += run(n, [1, 4, 16, 64] if n != "3840x2160" else [1, 4, 16])
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/frames_in_flight.json", "w"), indent=1)
