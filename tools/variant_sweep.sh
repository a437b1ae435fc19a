#!/bin/bash
# times the LK launch and the c2 step for every library under build/variants (bring-up aid: build variants of lk.cu with
# -DPSNW_LK_WARPS=.. -DPSNW_LK_MINB=.. and link them like csrc/Makefile does)
for f in build/variants/*.so; do
  PSNW_KLT_LIB=$PWD/$f timeout 300 python tools/quick_time.py 2>&1 | tail -1
done
