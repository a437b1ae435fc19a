#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S="768,576,4,4;768,576,4,16;768,576,4,64;1920,1080,4,8"
for v in "" build/variants/fminb5.so build/variants/fminb6.so; do
  for t in 0 1 2; do
  echo "lib=$v variant=$t"; PSNW_FUSED_VARIANT=$t PSNW_KLT_LIB=${v:+$PWD/$v} PYR_SWEEP_ONLY="$S" timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200
  done
done
