#!/usr/bin/env python3
"""Algorithmic FP64 work per cell-timestep from the oracle's op counter.

oracle/libsuews_oracle_count.so (`make -C oracle count`) is the oracle compiled with a counting `double`: it executes the
restated reference algorithm unchanged and counts every FP64 add/subtract, multiply, divide, sqrt and libm call.  Two
counts per run: "literal" = everything SUEWS_cal_Main does as written (all stages on every trip of the surface-temperature
loop), "hoisted" = the minimal algorithm of SURVEY 8(d): trip-invariant stages once per step, the time-only part of the
sun position left to the per-time-zone pre-pass, the state update once.  The diagnostics no flux and no state depends
on (RSL profile, SOLWEIG) are counted apart.  Weighted flops: add/mul/div/sqrt 1, exp 20, log 25, pow 60, sin/cos/tan 30,
atan/asin/acos/atan2 35, tanh 30.

usage: tools/oracle_opcount.py [scheme] [ncells] [day,day,...]   -> prints a table and writes profiles/oracle_opcount_<scheme>.json
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import oracle_lib as ol  # noqa: E402
from suews_b200 import synth  # noqa: E402

NAMES = ["add", "mul", "div", "sqrt", "exp", "log", "pow", "sincos", "atan", "tanh"]
WEIGHT = [1, 1, 1, 1, 20, 25, 60, 30, 35, 30]


def main():
    scheme = sys.argv[1] if len(sys.argv) > 1 else "default"
    ncell = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    days = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 73, 146, 219, 292]
    import subprocess
    subprocess.check_call(["make", "-C", ol.ORACLE_DIR, "-s", "count"])
    ol.LIB_PATH = os.path.join(ol.ORACLE_DIR, "libsuews_oracle_count.so")
    ol._lib = None
    L = ol.lib()
    L.so_opcount_start.argtypes = [C.c_int]
    L.so_opcount_stop.argtypes = [C.POINTER(C.c_ulonglong)]
    base, met_year = bench.load_inputs(366)
    d = synth.base_with_scheme(base, scheme)
    site = synth.cell_site_arrays(0, ncell, d)
    res = {}
    for mode, label in ((1, "hoisted"), (0, "literal")):
        tot = np.zeros(32)
        steps = 0
        trips = 0
        for day in days:
            met = met_year[day * 288:(day + 1) * 288]
            for c in range(ncell):
                dd = dict(d)
                for k in ("sfr", "lat", "lng", "bldgH", "EveTreeH", "DecTreeH", "Z"):
                    dd[k.lower()] = np.atleast_1d(site[k][c])
                m = met.copy()
                m[:, 11] += site["dTemp_C"][c]
                m[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), m[:, 10] + site["dRH"][c]))
                cell = ol.Cell(dd)
                buf = (C.c_ulonglong * 32)()
                L.so_opcount_start(mode)
                cell.run(m, want=())
                L.so_opcount_stop(buf)
                tot += np.array(list(buf), dtype=float)
                steps += 288
        main_c, diag_c = tot[:10] / steps, tot[16:26] / steps
        res[label] = {
            "ops_per_cell_step": dict(zip(NAMES, [float(x) for x in main_c])),
            "weighted_flops_per_cell_step": float(np.dot(main_c, WEIGHT)),
            "diagnostics_ops_per_cell_step": dict(zip(NAMES, [float(x) for x in diag_c])),
            "diagnostics_weighted_flops_per_cell_step": float(np.dot(diag_c, WEIGHT)),
        }
        res[label]["cal_Stab"] = {"calls_per_cell_step": float(tot[12] / steps), "iterations_per_call": float(tot[13] / max(tot[12], 1)),
                                  "share_of_calls_at_the_330_iteration_cap": float(tot[14] / max(tot[12], 1))}
        print("%-8s flux/state path: %8.0f weighted flops per cell-step (%s)" % (
            label, res[label]["weighted_flops_per_cell_step"], ", ".join("%s %.0f" % (n, v) for n, v in zip(NAMES, main_c))))
        print("%-8s diagnostics    : %8.0f weighted flops per cell-step (%s)" % (
            label, res[label]["diagnostics_weighted_flops_per_cell_step"], ", ".join("%s %.0f" % (n, v) for n, v in zip(NAMES, diag_c))))
    out = {"scheme": scheme, "cells": ncell, "days_of_year": [x + 1 for x in days], "weights": dict(zip(NAMES, WEIGHT)),
           "source": "oracle/libsuews_oracle_count.so (tools/oracle_opcount.py): synthetic cells of suews_b200/synth.py, one day each per listed day of 2004",
           **res}
    path = os.path.join(ROOT, "profiles", "oracle_opcount_%s.json" % scheme)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
