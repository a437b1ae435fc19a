#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
python - <<'P'
import os, sys, subprocess
for mode, thr, var in (("hybrid","",""), ("fused","",""), ("fused","256","1"), ("fused","256","2"), ("fused","512","2")):
    env = dict(os.environ, PSNW_PYR_MODE=mode)
    if thr: env["PSNW_FUSED_THREADS"] = thr
    if var: env["PSNW_FUSED_VARIANT"] = var
    out = subprocess.run([sys.executable, "-c", """
import bench
rig = bench.GpuRig('c2', 0, 0); rig.prime(False)
for cfg in [(768,576,4,64),(768,576,4,16)]:
    m = bench.pyramid_modes(rig, *cfg, isolated_reps=6, stream_launches=16)
    print(cfg[3], 'iso %.4f stream %.4f frac %.3f' % (m['isolated_ms'], m['stream_ms'], m['algorithmic_bytes']/m['stream_ms']/1e6/6537.3), end=' | ')
"""], env=env, capture_output=True, text=True)
    print(mode, thr, var, out.stdout.strip(), out.stderr[-300:] if out.returncode else "")
P
