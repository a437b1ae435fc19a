#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
python - <<'P'
import os, sys, json
sys.path.insert(0, os.getcwd())
import subprocess
for pad in (0, 20, 32, 48, 64):
    env = dict(os.environ, PSNW_FUSED_PAD_KB=str(pad))
    out = subprocess.run([sys.executable, "-c", """
import bench, torch, json
rig = bench.GpuRig('c2', 0, 0); rig.prime(False)
for cfg in [(768,576,4,64),(768,576,4,16)]:
    m = bench.pyramid_modes(rig, *cfg, isolated_reps=8, stream_launches=16)
    print(cfg, 'iso %.4f stream %.4f' % (m['isolated_ms'], m['stream_ms']), 'frac_stream %.3f' % (m['algorithmic_bytes']/m['stream_ms']/1e6/6537.3))
"""], env=env, capture_output=True, text=True)
    print("pad", pad, out.stdout.strip().replace("\n", " | "), out.stderr[-200:] if out.returncode else "")
P
