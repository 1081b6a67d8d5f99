#!/usr/bin/env python3
"""Summarise one `ncu --set full` capture of suews_run_kernel into profiles/kernel_counters.json (read by
bench.py for the `roofline.traffic` and `fp64` fields) and print a short table.

usage: tools/ncu_counters.py report.ncu-rep ncell nsteps [--write]
The capture is of tools/prof_run.py (one launch of `ncell` cells x `nsteps` timesteps)."""
import csv
import json
import os
import subprocess
import sys

rep, ncell, nsteps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, val = rows[0], rows[2]
m = {}
for h, v in zip(hdr, val):
    try:
        m[h] = float(v.replace(",", ""))
    except ValueError:
        pass
units = dict(zip(hdr, rows[1]))


def scaled(name):
    u, v = units[name], m[name]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)


cs = float(ncell) * nsteps
cycles = m["sm__cycles_elapsed.avg"]
flop_per_cycle = (2.0 * m["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"] +
                  m["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"] +
                  m["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"])
flops = flop_per_cycle * cycles
dram = scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum")
t = scaled("gpu__time_duration.sum")
res = {
    "source": "ncu --set full of `tools/prof_run.py %d %d` (%s); per launch of %d cells x %d steps" % (
        ncell, nsteps, os.path.basename(rep), ncell, nsteps),
    "kernel_ms_under_ncu": 1e3 * t,
    "fp64_flops_per_cell_step": flops / cs,
    "dram_bytes_per_cell_step": dram / cs,
    "dram_bytes_per_launch_of_profile": dram,
    "warp_instructions_per_cell_step": m["smsp__inst_executed.sum"] / (cs / 32.0),
    "fp64_pipe_pct_of_peak_active": m["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"],
    "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
    "warps_active_pct_of_peak": m["sm__warps_active.avg.pct_of_peak_sustained_active"],
    "threads_per_warp_instruction": m["smsp__thread_inst_executed_per_inst_executed.ratio"],
    "branch_targets_uniform_pct": m.get("smsp__sass_average_branch_targets_threads_uniform.pct"),
    "l1_hit_pct": m["l1tex__t_sector_hit_rate.pct"],
    "l2_hit_pct": m["lts__t_sector_hit_rate.pct"],
    "registers_per_thread": m["launch__registers_per_thread"],
    "stall_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: v for k, v in m.items()
                        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")},
}
for k, v in res.items():
    print(k, "=", v)
if "--json" in sys.argv:   # one capture's counters to a file of its own (tools/run_captures.sh keeps these, not the reports)
    with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
        json.dump(res, f, indent=1)
if "--write" in sys.argv:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "profiles", "kernel_counters.json"), "w") as f:
        json.dump(res, f, indent=1)
