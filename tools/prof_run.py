#!/usr/bin/env python3
"""Small fixed workload for ncu: one wave of cells, a few dozen timesteps, 1 warm-up + 2 launches.
usage: python tools/prof_run.py [ncell] [nsteps] [scheme] [row0]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from suews_b200 import OUT_SUM7, OUT_SUM82, SuewsBatch, synth
from suews_b200.batch import OUT_ALL_STAGES
OUT = OUT_SUM82 if os.environ.get('SU_SUMS', '7') == '82' else OUT_SUM7   # output policy of the profiled launches
if os.environ.get('SU_ALL_STAGES', '0') == '1':   # P0 with the diagnostics its sums cannot observe (T2/q2/U10/RH2) kept
    OUT |= OUT_ALL_STAGES

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 48
scheme = sys.argv[3] if len(sys.argv) > 3 else "default"
row0 = int(sys.argv[4]) if len(sys.argv) > 4 else 120   # start at 10:00 so day-time physics is in the sample
base, met = bench.load_inputs(max(2, (row0 + 3 * nsteps + 287) // 288))
b = SuewsBatch(ncell, 1)
ts5 = synth.estm_ts_forcing(met[: row0 + 3 * nsteps]) if synth.SCHEMES[scheme]["storageheatmethod"] in (4, 14) else None
synth.build_batch(b, base, scheme=scheme, ts_row0=None if ts5 is None else ts5[0])
b.upload()
b.set_forcing(met[: row0 + 3 * nsteps])
if ts5 is not None:
    b.set_estm_forcing(ts5)
if row0 > 0:
    b.advance(0, row0, OUT)
import ctypes as C
from suews_b200._lib import lib
PH = ["P0 geometry/daily", "P1 moist air+hoisted", "L1 rad/OHM/melt", "L2 stab/resist", "L3 QE/QH/Tsfc", "F1 water balance",
      "F2 RSL profile", "F3 commit/output"]
clk = (C.c_ulonglong * 16)()
has_clk = hasattr(lib(), "suews_b200_debug_phase_clocks")
for i in range(3):
    if has_clk:
        lib().suews_b200_debug_phase_clocks(clk, 1)
    b.advance(row0 + i * nsteps, nsteps, OUT)
    print("launch %d: %.2f ms -> %.3e cell-steps/s" % (i, b.last_ms(), ncell * nsteps / (b.last_ms() * 1e-3)))
    if has_clk:
        lib().suews_b200_debug_phase_clocks(clk, 0)
        tot = float(sum(clk[:8])) or 1.0
        tot += float(sum(clk[10:16]))
        print("   block-steps %d, Tsurf trips/step %.2f; share of block cycles: " % (clk[9], clk[8] / max(clk[9], 1)) +
              ", ".join("%s %.1f%%" % (PH[k], 100.0 * clk[k] / tot) for k in range(8)))
        SUB = ["P1a cal_AtmMoist (1st Lat_vap search)", "P1b soil moisture, water use, QF, snow albedo", "P1c drainage", "P1d ReDistributeWater",
               "P1e SurfaceResistance", "F1a cal_QE_full (state update)"]
        print("   sub-stages (own barrier; taken out of P1 / F1 above): " + ", ".join("%s %.1f%%" % (SUB[k], 100.0 * clk[10 + k] / tot) for k in range(6)))
fatal, warn, iters = b.diag()
print("fatal cells", int((fatal != 0).sum()), "iters/step", iters.sum() / (ncell * (row0 + 3.0 * nsteps)))
