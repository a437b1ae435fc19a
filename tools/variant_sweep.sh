#!/bin/bash
# times the LK launch and the c2 step for every library under build/variants (bring-up aid)
for f in build/variants/*.so; do
  PSNW_KLT_LIB=$PWD/$f timeout 300 python tools/quick_time.py 2>&1 | tail -1
done
