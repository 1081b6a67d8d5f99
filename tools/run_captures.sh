#!/bin/bash
# ncu --set full captures of the P0 kernel (default scheme set) at four points of the year: winter night, winter day, summer
# night, summer day (48 five-minute steps each, one wave of cells), each only after the same command exited 0 without ncu.
# usage (on the GPU box): tools/run_captures.sh TAG   -> gpurun_out/prof_TAG_{wn,wd,sn,sd}.ncu-rep + plain logs
tag=${1:-r2}
declare -A ROW=( [wn]=0 [wd]=120 [sn]=$((190*288)) [sd]=$((190*288+132)) )
for k in wn wd sn sd; do
  r=${ROW[$k]}
  timeout 300 python tools/prof_run.py 151552 48 default $r > gpurun_out/cap_${tag}_$k.plain.log 2>&1 || { echo "plain run $k failed"; continue; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:suews_run_kernel -s 2 -c 1 -f -o gpurun_out/prof_${tag}_$k \
      python tools/prof_run.py 151552 48 default $r > gpurun_out/cap_${tag}_$k.ncu.log 2>&1
  tail -2 gpurun_out/cap_${tag}_$k.plain.log
  # the reports are ~20 MB each and gpurun brings back 64 MB: keep the counters of all four, the report of the winter day only
  python tools/ncu_counters.py gpurun_out/prof_${tag}_$k.ncu-rep 151552 48 --json gpurun_out/counters_${tag}_$k.json > /dev/null 2>&1
  [ $k = wd ] || rm -f gpurun_out/prof_${tag}_$k.ncu-rep
done
