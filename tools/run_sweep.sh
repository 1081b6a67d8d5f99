#!/bin/bash
# BASELINE configs[3]: physics-scheme sweep at 1M cells (one bench line per scheme), plus output policy P1
mkdir -p gpurun_out/sweep
for s in snow baserun estm stab2 stab4 netrad13; do
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --scheme $s > gpurun_out/sweep/bench_$s.json 2> gpurun_out/sweep/bench_$s.err
  echo "$s rc=$? $(head -c 160 gpurun_out/sweep/bench_$s.json)"
done
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sums hourly > gpurun_out/sweep/bench_default_P1.json 2> gpurun_out/sweep/bench_default_P1.err
echo "P1 rc=$? $(head -c 160 gpurun_out/sweep/bench_default_P1.json)"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sums all82 > gpurun_out/sweep/bench_default_P0plus.json 2> gpurun_out/sweep/bench_default_P0plus.err
echo "P0+ rc=$? $(head -c 160 gpurun_out/sweep/bench_default_P0plus.json)"
