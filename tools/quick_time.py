#!/usr/bin/env python
"""Quick device timing of the LK launch and the whole step for the library named by PSNW_KLT_LIB (bring-up aid)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rig = bench.GpuRig(os.environ.get("QT_WORKLOAD", "c2"), 0, 0, pipeline=os.environ.get("QT_PIPELINE", "1") == "1")
rig.prime(False)
lk, frac = rig.time_track(reps=20)
sec, launches = rig.timed(50, 5, host=False)
print(json.dumps({"lib": os.path.basename(os.environ.get("PSNW_KLT_LIB", "default")), "lk_ms": round(lk * 1e3, 4),
                  "step_ms": round(sec / 50 * 1e3, 4), "status_ok": round(frac, 4)}))
