#!/usr/bin/env python3
"""First step at which one synthetic cell differs between the CUDA library and the oracle (debugging aid).
usage: python tools/diff_cell.py scheme day0 cell [ndays] [tol]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures as fx
import oracle_lib as ol
from suews_b200 import OUT_FULL, SuewsBatch, synth
from test_gpu_parity import rel_err

scheme, day0, cellno = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ndays = int(sys.argv[4]) if len(sys.argv) > 4 else 3
tol = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-9
base = fx.state_init(2005)
met = fx.forcing(2004)[day0 * 288:(day0 + ndays) * 288]
if os.environ.get("SU_LDOWN"):      # a physical incoming long-wave series (the test forcing's column is not one)
    met = met.copy()
    met[:, 16] = 0.85 * 5.67e-8 * (met[:, 11] + 273.15) ** 4
b = SuewsBatch(1, 1)
_m, site = synth.build_batch(b, base, cell0=17 * day0 + cellno, scheme=scheme)
over = {}
for kv in os.environ.get("SU_OVER", "").split(","):      # e.g. SU_OVER=netradiationmethod=11,gsmodel=1
    if "=" in kv:
        k, v = kv.split("=")
        over[k] = [float(v)]
        b.set_class_param(0, k, np.asarray([float(v)]))
b.upload(); b.set_forcing(met); b.advance(0, met.shape[0], OUT_FULL)
d0 = synth.base_with_scheme(base, scheme)
d0.update(over)
dd = dict(d0)
for k in ("sfr", "lat", "lng", "bldgH", "EveTreeH", "DecTreeH", "Z"):
    dd[k.lower()] = np.atleast_1d(site[k][0])
m = met.copy()
m[:, 11] += site["dTemp_C"][0]
m[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), m[:, 10] + site["dRH"][0]))
ref = ol.Cell(dd).run(m, want=("SUEWS", "snow", "RSL", "DailyState"))
names = {g: [v["name"] for v in fx.varlist(g)] for g in ("SUEWS", "snow", "RSL", "DailyState")}
for g in ("SUEWS", "snow"):
    out = b.block(g, 0, met.shape[0])
    e = rel_err(out[:, 5:], ref[g][:, 5:], g)
    bad = np.argwhere(e > tol)
    if not bad.size:
        print(g, "no difference above", tol); continue
    i = bad[:, 0].min()
    cols = sorted(set(bad[bad[:, 0] == i][:, 1]))
    print("%s: first difference at step %d (%02d:%02d of day %d); columns %s" % (
        g, i, int(ref[g][i, 2]), int(ref[g][i, 3]), int(ref[g][i, 1]), [(int(c), names[g][c] if c < len(names[g]) else "") for c in cols][:24]))
    for c in cols[:8]:
        print("     col %3d %-12s gpu %.12g  ref %.12g" % (c, names[g][c] if c < len(names[g]) else "", out[i, 5 + c], ref[g][i, 5 + c]))
    if i > 0:
        print("  forcing row: T %.3f RH %.2f U %.2f P %.3f kdn %.1f" % (m[i, 11], m[i, 10], m[i, 9], m[i, 13], m[i, 14]))
if os.environ.get("SU_DUMP"):
    g, cols, rows = os.environ["SU_DUMP"].split(":")
    cols = [int(c) for c in cols.split(",")]
    r0, r1 = [int(x) for x in rows.split("-")]
    out = b.block(g, 0, met.shape[0])
    for i in range(r0, r1 + 1):
        print(i, " ".join("c%d gpu %.10g ref %.10g |" % (c, out[i, 5 + c], ref[g][i, 5 + c]) for c in cols))
