#!/usr/bin/env python
"""CPU baselines of SURVEY.md section 8(d) on this host (TEST INFRASTRUCTURE: uses cv2 and the oracle):
  (i)  batched: one calcOpticalFlowPyrLK per direction over all points of a camera-frame, raw images (the cv2 Python
       binding rejects prebuilt pyramids), 1 thread and all threads; buildOpticalFlowPyramid timed separately;
  (ii) reference call pattern (psn_where/PSNWhere_Tracker2D.cpp:776-782, :871-877): per-target calls on raw images,
       window (int)w x (int)w backward over 3 frames and (int)w x (int)h forward, default maxLevel 3, 3*D + T calls.
Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
from mcmtt_opticalflow_b200 import synth
from oracle import cv2_ref, psnw_oracle as po


def median_time(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return float(np.median(ts))


def main():
    w, h, ntgt, per = 768, 576, 40, 100
    cam = synth.Camera(w, h, ntgt, cam=0)
    frames = [cam.frame(t) for t in range(4)]
    boxes = cam.boxes(3)
    eig = cv2.cornerMinEigenVal(frames[3], 3, ksize=3)
    seeds = [po.detect(frames[3], b, per, 0.01, 3.0, eig=eig) for b in boxes]
    allpts = np.concatenate(seeds).astype(np.float32)
    out = {"host_cores": os.cpu_count(), "cv2": cv2.__version__, "shape": [w, h], "targets": ntgt, "points": int(len(allpts))}
    for threads in (1, os.cpu_count()):
        cv2.setNumThreads(threads)
        key = "t%d" % threads
        r = {}
        r["pyramid_scharr_ms"] = 1e3 * median_time(lambda: cv2.buildOpticalFlowPyramid(frames[3], (21, 21), 3, withDerivatives=True))
        t_fb = median_time(lambda: cv2_ref.track_fb(frames[2], frames[3], allpts, (21, 21), 3))
        r["batched_fb_ms"] = 1e3 * t_fb
        r["batched_fb_points_per_s"] = len(allpts) / t_fb

        def reference_pattern():
            n_tracks = 0
            for b, s in zip(boxes, seeds):
                if len(s) < 4:
                    continue
                cur, win = s, (max(int(b[2]), 3), max(int(b[2]), 3))
                for k in range(1, 4):                                   # backward chain, :763-811
                    o, st, er = cv2_ref.lk(frames[4 - k], frames[3 - k], cur, win, 3)
                    nb, inl = po.local_search_klt(b, cur, o)
                    n_tracks += len(cur)
                    if len(inl) < 4:
                        break
                    cur = o[inl]
                o, st, er = cv2_ref.lk(frames[2], frames[3], s, (max(int(b[2]), 3), max(int(b[3]), 3)), 3)   # forward, :871-877
                po.local_search_klt(b, s, o, st)
                n_tracks += len(s)
            return n_tracks
        n_tracks = reference_pattern()
        t_ref = median_time(reference_pattern, reps=3, warm=1)
        r["reference_pattern_ms"] = 1e3 * t_ref
        r["reference_pattern_point_tracks"] = n_tracks
        r["reference_pattern_point_tracks_per_s"] = n_tracks / t_ref
        r["reference_pattern_frames_per_s"] = 1.0 / t_ref
        out[key] = r
    print(json.dumps(out))


if __name__ == "__main__":
    main()
