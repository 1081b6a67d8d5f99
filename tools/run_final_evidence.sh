#!/bin/bash
# round-1 final evidence for the final build
set -x
python tools/prof_run.py 151552 48 > gpurun_out/prof_plain_r1l.txt 2>&1; tail -4 gpurun_out/prof_plain_r1l.txt
ncu --set full --clock-control none --import-source on -k regex:suews_run_kernel -s 2 -c 1 -f -o gpurun_out/prof_r1l python tools/prof_run.py 151552 48 > gpurun_out/ncu_r1l.log 2>&1; tail -2 gpurun_out/ncu_r1l.log
python tools/prof_run.py 151552 48 snow > gpurun_out/prof_plain_r1l_snow.txt 2>&1; tail -3 gpurun_out/prof_plain_r1l_snow.txt
ncu --set full --clock-control none --import-source on -k regex:suews_run_kernel -s 2 -c 1 -f -o gpurun_out/prof_r1l_snow python tools/prof_run.py 151552 48 snow > gpurun_out/ncu_r1l_snow.log 2>&1; tail -2 gpurun_out/ncu_r1l_snow.log
python bench.py --steps 2 --warmup 3 --cells-per-gpu 151552 --no-cpu-baseline --no-flux-only > gpurun_out/bench_small_r1l.json 2> gpurun_out/bench_small_r1l.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --steps 2 --warmup 3 --cells-per-gpu 151552 --no-cpu-baseline --no-flux-only > gpurun_out/ncu_launch_r1l.log 2>&1
python bench.py > gpurun_out/bench_r1_final3.json 2> gpurun_out/bench_r1_final3.err; cat gpurun_out/bench_r1_final3.json
bash tools/run_sweep.sh 2>&1 | tee gpurun_out/sweep_r1l.txt
