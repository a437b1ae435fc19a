#!/usr/bin/env python
"""Pyramid + Scharr (K1) bandwidth sweep over frame sizes / frames per launch (SURVEY.md C5 shapes)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcmtt_opticalflow_b200 import klt

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
stream = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []
ONLY = os.environ.get("PYR_SWEEP_ONLY")        # e.g. "768,576,4,64": a single configuration (profiling aid)
CFGS = [tuple(int(v) for v in c.split(",")) for c in ONLY.split(";")] if ONLY else [(768, 576, 4, 1), (768, 576, 4, 4), (768, 576, 4, 16), (768, 576, 4, 64), (1920, 1080, 4, 1),
                             (1920, 1080, 4, 8), (1920, 1080, 4, 32), (3840, 2160, 5, 1), (3840, 2160, 5, 4), (3840, 2160, 5, 16)]
for (w, h, levels, ncam) in CFGS:
    ctx = klt.KLTContext(w, h, n_cameras=ncam, levels=levels, ring_slots=1, stream=stream.cuda_stream, max_points=1024, max_targets=64)
    rng = np.random.default_rng(0)
    frames = [rng.integers(0, 256, size=(h, w), dtype=np.uint8) for _ in range(min(ncam, 4))]
    ctx.push_frames([frames[i % len(frames)] for i in range(ncam)])
    cams = np.arange(ncam, dtype=np.int32)
    with torch.cuda.stream(stream):
        for _ in range(3):
            ctx.rebuild_pyramids(cams)
        reps = 20
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i in range(reps):
            flush.zero_()
            ev[i][0].record(stream)
            ctx.rebuild_pyramids(cams)
            ev[i][1].record(stream)
        stream.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    alg = ctx.push_frame_bytes * ncam
    r = {"w": w, "h": h, "levels": levels, "frames": ncam, "MB": round(alg / 1e6, 2), "ms_mean": round(float(ms.mean()), 4),
         "ms_min": round(float(ms.min()), 4), "GBs_mean": round(alg / ms.mean() / 1e6, 1), "frac": round(alg / ms.mean() / 1e6 / PEAK, 3)}
    print(json.dumps(r), flush=True); out.append(r)
    ctx.close()
json.dump(out, open(os.environ.get("PYR_SWEEP_OUT", "gpurun_out/pyr_sweep.json"), "w"), indent=1)
