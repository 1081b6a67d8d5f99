#!/usr/bin/env python3
"""Cross-check two builds of libsuews_b200.so bit for bit (e.g. default -O3 vs `-Xptxas -O1`): arithmetic is
contraction-free and the math functions are the same, so any difference is a code-generation problem.
usage (GPU box): python tools/build_consistency.py libA.so libB.so [ncell] [ndays]"""
import os, shutil, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import bench
from suews_b200 import OUT_SUM82, SuewsBatch, synth
from suews_b200.batch import all_fields
ncell, ndays, scheme, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
base, met = bench.load_inputs(ndays)
b = SuewsBatch(ncell, 1)
synth.build_batch(b, base, scheme=scheme)
b.upload(); b.set_forcing(met)
for d in range(ndays):
    b.advance(d * 288, 288, OUT_SUM82)
res = {"sums": b.sums()}
for nm, kind, n in all_fields():
    if kind == 3:
        res["st_" + nm] = b.get_state(nm)
np.savez(out, **res)
''' % (ROOT, ROOT)

def run(lib, tag, ncell, ndays):
    target = os.path.join(ROOT, "suews_b200", "libsuews_b200.so")
    if os.path.abspath(lib) != target:
        shutil.copy(lib, target + ".tmp"); os.replace(target + ".tmp", target)
    outs = {}
    for scheme in ("default", "snow", "baserun"):
        out = "/tmp/consistency_%s_%s.npz" % (tag, scheme)
        subprocess.check_call([sys.executable, "-c", CHILD, str(ncell), str(ndays), scheme, out])
        outs[scheme] = dict(np.load(out))
    return outs

def main():
    a, b = sys.argv[1], sys.argv[2]
    ncell = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    ndays = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    keep = os.path.join(ROOT, "suews_b200", "libsuews_b200.so")
    shutil.copy(keep, "/tmp/keep.so")
    ra = run(a, "a", ncell, ndays)
    rb = run(b, "b", ncell, ndays)
    shutil.copy("/tmp/keep.so", keep)
    bad = 0
    for scheme in ra:
        for k in ra[scheme]:
            x, y = ra[scheme][k], rb[scheme][k]
            same = (x == y) | (np.isnan(x) & np.isnan(y))
            if not same.all():
                bad += 1
                idx = np.argwhere(~same)
                print("DIFF scheme=%s field=%s: %d of %d values, first at %s: %r vs %r" % (
                    scheme, k, (~same).sum(), same.size, tuple(idx[0]), x[tuple(idx[0])], y[tuple(idx[0])]))
    print("build consistency: %s (%d cells x %d days x 3 scheme sets)" % ("IDENTICAL" if bad == 0 else "%d fields differ" % bad, ncell, ndays))
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
