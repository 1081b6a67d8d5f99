#!/bin/bash
# ncu --set full captures of the other kernel instances (policy P0): snow on (<2,1,0>), the BaseRun set (snow + NetRadiationMethod
# 13, same instance), AnOHM and ESTM (the general-storage instance <2,0,-1>) on a winter day, each only after the same command
# exited 0 without ncu; only the counters are kept.   usage (GPU box): tools/run_captures_schemes.sh TAG
tag=${1:-r2}
for sch in snow baserun anohm estm; do
  timeout 300 python tools/prof_run.py 151552 48 $sch 120 > gpurun_out/cap_${tag}_$sch.plain.log 2>&1 || { echo "plain run $sch failed"; tail -3 gpurun_out/cap_${tag}_$sch.plain.log; continue; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:suews_run_kernel -s 2 -c 1 -f -o gpurun_out/prof_${tag}_$sch \
      python tools/prof_run.py 151552 48 $sch 120 > gpurun_out/cap_${tag}_$sch.ncu.log 2>&1
  tail -2 gpurun_out/cap_${tag}_$sch.plain.log
  python tools/ncu_counters.py gpurun_out/prof_${tag}_$sch.ncu-rep 151552 48 --json gpurun_out/counters_${tag}_$sch.json > /dev/null 2>&1
  rm -f gpurun_out/prof_${tag}_$sch.ncu-rep
done
