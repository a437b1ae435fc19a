#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false --cudart shared -I include -o /tmp/pyr_timeline tools/pyr_timeline.cu -lcuda || exit 1
for args in "4 0 0"; do echo "== $args"; /tmp/pyr_timeline $args | tail -24; done
