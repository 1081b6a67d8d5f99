#!/bin/bash
# times every tuning build under build/variants/ (tools/variants.sh) on the same workload: 4 waves of cells, 3 whole days
for f in build/variants/lib_*.so; do
  echo "== $f"
  SUEWS_B200_LIB=$PWD/$f timeout 300 python tools/prof_run.py 606208 288 default 0 2>&1 | grep -v "^ "
done
