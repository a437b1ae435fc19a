#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
S="768,576,4,16;768,576,4,32;768,576,4,64;768,576,4,128;1920,1080,4,8"
for m in level hybrid; do echo "mode=$m"; PSNW_PYR_MODE=$m PYR_SWEEP_ONLY="$S" timeout 300 python tools/pyr_sweep.py 2>&1 | cut -c1-60,88-200; done
