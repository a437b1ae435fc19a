#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
python - <<'P'
import os, sys, subprocess
for mode in ("auto", "fused", "hybrid", "level"):
    env = dict(os.environ, PSNW_PYR_MODE=mode)
    out = subprocess.run([sys.executable, "-c", """
import bench
rig = bench.GpuRig('c2', 0, 0); rig.prime(False)
for cfg in [(768,576,4,4),(768,576,4,1),(768,576,4,16),(768,576,4,64),(1920,1080,4,4)]:
    m = bench.pyramid_modes(rig, *cfg, isolated_reps=6, stream_launches=48 if cfg[3] < 16 else 16)
    print(cfg[3], 'iso %.4f stream %.4f frac %.3f' % (m['isolated_ms'], m['stream_ms'], m['algorithmic_bytes']/m['stream_ms']/1e6/6537.3), end=' | ')
"""], env=env, capture_output=True, text=True)
    print(mode, out.stdout.strip(), out.stderr[-300:] if out.returncode else "")
P
