#!/usr/bin/env python3
"""FP64 operations the KERNEL's algorithm performs per cell-timestep, counted on the host build of the kernel's step.

tests/host_step.cpp compiled with the oracle's counting `double` (oracle/so_count.h) executes suews_b200/csrc/dev_step.cuh --
the reorganised step as the GPU runs it, libm in place of the CUDA math wrappers -- and counts every add/subtract, multiply,
divide, sqrt and libm call: the count next to tools/oracle_opcount.py's "hoisted" estimate, which is derived from the oracle
(the as-written algorithm with the trip-invariant stages counted once).  Same sample (synthetic cells, days spread over
2004), same weights.  Two policies: P0 with the flux-only switch (what the bench launches) and FULL (every output line).

usage: tools/host_step_opcount.py [scheme] [ncells] [day,day,...]   -> prints a table and writes profiles/host_step_opcount_<scheme>.json
"""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import test_step_host as T  # noqa: E402
from suews_b200 import synth  # noqa: E402

NAMES = ["add", "mul", "div", "sqrt", "exp", "log", "pow", "sincos", "atan", "tanh"]
WEIGHT = [1, 1, 1, 1, 20, 25, 60, 30, 35, 30]


def main():
    scheme = sys.argv[1] if len(sys.argv) > 1 else "default"
    ncell = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    days = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 73, 146, 219, 292]
    so = os.path.join(tempfile.mkdtemp(prefix="host_step_count_"), "host_step_count.so")
    subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-fpermissive", "-w",
                           "-include", "vector", "-include", os.path.join(ROOT, "oracle", "so_count.h"), "-o", so,
                           os.path.join(ROOT, "tests", "host_step.cpp")])
    H = C.CDLL(so)
    H.host_step_run.argtypes = [T.ip, T.dp, T.dp, T.dp, T.dp, T.dp, C.c_int, T.dp, T.dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(T.dp), C.POINTER(C.c_ulonglong), C.POINTER(C.c_longlong), C.c_int, C.c_int, T.dp]
    buf = (C.c_ulonglong * 32).in_dll(H, "so_opcount_buf")
    on = C.c_int.in_dll(H, "so_opcount_on")
    base, met_year = bench.load_inputs(366)
    d = synth.base_with_scheme(base, scheme)
    site = synth.cell_site_arrays(0, ncell, d)
    record = d["storageheatmethod"][0] == 3
    special = 0 if record else (1 if d["snowuse"][0] == 0 else 2)
    res = {}
    for label, kw in (("P0_flux_only", dict(kind=2, flux_only=True, instance=special)), ("FULL", dict(kind=0, instance=0))):
        tot, steps = np.zeros(10), 0
        for day in days:
            met = met_year[day * 288:(day + 1) * 288]
            for c in range(ncell):
                dd = dict(d)
                for k, v in site.items():
                    dd[k.lower()] = list(np.atleast_1d(v[c]))
                for i in range(32):
                    buf[i] = 0
                on.value = 1
                acc = np.zeros(7)
                rc, _b, _s, _w = T.run_host(H, dd, met, ("SUEWS",) if kw["kind"] == 0 else (), record=record, acc=acc if kw["kind"] else None, **kw)
                on.value = 0
                assert rc == 0
                tot += np.array([buf[i] for i in range(10)], dtype=float)
                steps += 288
        per = tot / steps
        res[label] = {"ops_per_cell_step": dict(zip(NAMES, [float(x) for x in per])), "weighted_flops_per_cell_step": float(np.dot(per, WEIGHT))}
        print("%-13s %8.0f weighted flops per cell-step (%s)" % (label, res[label]["weighted_flops_per_cell_step"],
                                                                 ", ".join("%s %.0f" % (n, v) for n, v in zip(NAMES, per))))
    out = {"scheme": scheme, "cells": ncell, "days_of_year": [x + 1 for x in days], "weights": dict(zip(NAMES, WEIGHT)),
           "source": "tests/host_step.cpp with oracle/so_count.h (tools/host_step_opcount.py): the kernel's own step, synthetic cells of "
                     "suews_b200/synth.py, one day each per listed day of 2004", **res}
    path = os.path.join(ROOT, "profiles", "host_step_opcount_%s.json" % scheme)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
