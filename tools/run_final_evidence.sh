#!/bin/bash
# Round-2 final evidence for the final build (on the GPU box): GPU tests, smoke, the driver-style bench line, four ncu captures
# of the P0 kernel across the year (each only after the same command exited 0 without ncu), the launch list of a bench run.
# usage: tools/run_final_evidence.sh TAG  -> gpurun_out/TAG_*
tag=${1:-r2final}
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${tag}_gputest.txt; cat gpurun_out/${tag}_gputest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["all_stages"]["value"], d["roofline_fp64"]["frac"])
print({k: "%.3e" % v["value"] for k, v in d["schemes"].items()})
PY
bash tools/run_captures.sh ${tag}
python tools/prof_run.py 606208 288 default 0 2>&1 | tail -3 > gpurun_out/${tag}_prof_default_jan.txt
python tools/prof_run.py 606208 288 anohm 0 2>&1 | tail -3 > gpurun_out/${tag}_prof_anohm_jan.txt
python tools/prof_run.py 606208 288 estm 0 2>&1 | tail -3 > gpurun_out/${tag}_prof_estm_jan.txt
python bench.py --steps 2 --warmup 3 --cells-per-gpu 151552 --no-cpu-baseline --scheme-steps 0 > gpurun_out/${tag}_bench_small.json 2> gpurun_out/${tag}_bench_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_151552cells.csv \
    python bench.py --steps 2 --warmup 3 --cells-per-gpu 151552 --no-cpu-baseline --scheme-steps 0 > gpurun_out/${tag}_ncu_launches.log 2>&1
tail -2 gpurun_out/${tag}_prof_*_jan.txt
