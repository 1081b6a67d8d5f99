#!/usr/bin/env python3
"""Output policy P2: every output block of every step kept on the device, the reference's
SUEWS_cal_multitsteps semantics, for ncell cells x nsteps steps.  usage: python tools/p2_run.py [ncell] [nsteps] [scheme]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from suews_b200 import OUT_FULL, SuewsBatch, synth

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 10240
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 288
scheme = sys.argv[3] if len(sys.argv) > 3 else "default"
base, met = bench.load_inputs(max(2, (4 * nsteps + 287) // 288))
b = SuewsBatch(ncell, 1)
synth.build_batch(b, base, scheme=scheme)
b.upload()
b.set_forcing(met[: 4 * nsteps])
for i in range(4):
    b.advance(i * nsteps, nsteps, OUT_FULL)
    gb = ncell * nsteps * (87 + 103 * (scheme != "default") + 138 + 50 + 31) * 8 / 1e9
    print("launch %d: %.2f ms -> %.3e cell-steps/s, %.2f GB of blocks written (%.0f GB/s)" % (
        i, b.last_ms(), ncell * nsteps / (b.last_ms() * 1e-3), gb, gb / (b.last_ms() * 1e-3)))
blk = b.block("SUEWS", ncell - 1, nsteps)
print("last cell, last step QN..QE:", blk[-1, 10:15])
