#!/bin/bash
# times the LK launch and the c2 step for library / persistent-grid variants (bring-up aid)
for cfg in "lib_w4_b5.so 4" "lib_w4_b5.so 5" "lib_w4_b6.so 4" "lib_w4_b6.so 5" "lib_w4_b6.so 6"; do
  set -- $cfg
  echo -n "ctas_per_sm=$2 "; PSNW_LK_CTAS_PER_SM=$2 PSNW_KLT_LIB=$PWD/build/variants/$1 timeout 300 python tools/quick_time.py 2>&1 | tail -1
done
