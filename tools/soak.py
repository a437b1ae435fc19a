#!/usr/bin/env python
"""Race hunt: the bench rig plays its frames ping-pong, so the pipelined step sequence (one frame in flight, overlapped
detection, eager Shi-Tomasi map, persistent TMA-staged LK) is periodic.  Runs N steps and checks that every step's
result blob (points, boxes, counts, status, mask) and next seeds are bit-identical to those one period earlier.
usage: tools/soak.py [steps] [workload] [host 0|1]"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
wl = sys.argv[2] if len(sys.argv) > 2 else "c2"
host = len(sys.argv) > 3 and sys.argv[3] == "1"
rig = bench.GpuRig(wl, 0, 0)
rig.prime(host)
period = 2 * bench.CYCLE - 2
sums, bad, snaps = [], 0, []
import torch
free_run = os.environ.get("SOAK_FREE_RUN", "1") == "1"     # no host sync between steps: the streams overlap across steps too
with torch.cuda.stream(rig.stream):
    for k in range(steps):
        rig.step(host)
        if free_run and not host:
            snaps.append((rig.d_blob.clone(), rig.d_seeds[rig.which].clone()))      # stream-ordered snapshots, compared at the end
            continue
        rig.ctx.sync()
        rig.stream.synchronize()
        snaps.append((rig.d_blob.cpu(), rig.d_seeds[rig.which].cpu()))
    rig.ctx.sync()
    rig.stream.synchronize()
for k, (b, sd) in enumerate(snaps):
    h = hashlib.sha1(b.cpu().numpy().tobytes())
    h.update(sd.cpu().numpy().tobytes())
    sums.append(h.hexdigest())
    if k >= 2 * period and sums[k] != sums[k - period]:
        bad += 1
        print("step %d differs from step %d" % (k, k - period))
print("soak %s host=%d: %d steps, period %d, %d distinct results, %d mismatches" % (wl, host, steps, period, len(set(sums[2 * period:])), bad))
sys.exit(1 if bad else 0)
