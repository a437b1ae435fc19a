#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
S="768,576,4,4;768,576,4,16;768,576,4,64;1920,1080,4,8"
for v in a_base b_carve4 c_carve5 c_carve6 c_carve8; do
  echo "lib=$v"; PSNW_KLT_LIB=$PWD/build/variants/$v.so PYR_SWEEP_ONLY="$S" timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200
done
