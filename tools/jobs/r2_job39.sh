#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
python - <<'P'
import os, sys, subprocess
for thr in ("512", "256"):
  for var in ("0", "1"):
    env = dict(os.environ, PSNW_FUSED_THREADS=thr, PSNW_FUSED_VARIANT=var)
    out = subprocess.run([sys.executable, "-c", """
import bench
rig = bench.GpuRig('c2', 0, 0); rig.prime(False)
for cfg in [(768,576,4,4),(768,576,4,1),(768,576,4,16)]:
    m = bench.pyramid_modes(rig, *cfg, isolated_reps=8, stream_launches=48)
    print(cfg[3], 'iso %.4f stream %.4f frac %.3f' % (m['isolated_ms'], m['stream_ms'], m['algorithmic_bytes']/m['stream_ms']/1e6/6537.3), end=' | ')
"""], env=env, capture_output=True, text=True)
    print("threads", thr, "variant", var, out.stdout.strip(), out.stderr[-300:] if out.returncode else "")
P
