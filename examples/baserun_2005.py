#!/usr/bin/env python3
"""The reference's BaseRun case (Test/BaseRun/2020b, 2005 leg: 1440 five-minute steps of one grid) through the CUDA
library, written out in the reference's own file formats:

    test1_2005_SUEWS_60.txt                     hourly SUEWS output (SUEWS_Output_Agg + SUEWS_Write_txt), aggregated in the kernel
    InitialConditionstest1_2005_EndofRun.nml    the end-of-run checkpoint (NextInitial)

and compared with the files the reference wrote (tests/golden/).  Needs a CUDA device.

    python examples/baserun_2005.py [outdir]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fixtures as fx  # noqa: E402  (the golden inputs: argument dump + hourly forcing)
from suews_b200 import OUT_AGG, SuewsBatch, _lib, checkpoint  # noqa: E402
from suews_b200.forcing import read_met  # noqa: E402


def write_txt(block, group, outlevel=2):
    f = _lib.lib().suews_write_txt
    f.restype = C.c_longlong
    f.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_longlong]
    b = np.asfortranarray(block, dtype=np.float64)
    args = (b.ctypes.data_as(C.POINTER(C.c_double)), b.shape[0], b.shape[1], group.encode(), outlevel, 1)
    n = f(*args, None, 0)
    buf = C.create_string_buffer(n + 1)
    f(*args, buf, n + 1)
    return buf.value.decode()


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    b = SuewsBatch(1, 1)
    b.load_reference_args(fx.state_init(2005))
    b.upload()
    # hourly forcing file -> 5-min MetForcingBlock on the device (MetRead + DisaggregateMet)
    nsteps = b.set_forcing_orig(read_met(os.path.join(fx.GOLDEN, "forcing_2005_60min.txt")), resolution_in=3600)
    b.advance(0, nsteps, OUT_AGG)           # one launch; hourly aggregation inside the kernel
    txt = write_txt(b.agg_block(0, nsteps // 12), "SUEWS")
    nml = checkpoint.write_text(checkpoint.state_vector(lambda n: b.get_state(n)[0]), snow_use=1)
    open(os.path.join(outdir, "test1_2005_SUEWS_60.txt"), "w").write(txt)
    open(os.path.join(outdir, "InitialConditionstest1_2005_EndofRun.nml"), "w").write(nml)
    b.close()

    gold = open(os.path.join(fx.GOLDEN, "ref_2005_SUEWS_60.txt")).read()
    fields = sum(len(ln.split()) for ln in gold.splitlines()[1:])
    bad = sum(x != y for a, c in zip(txt.splitlines(), gold.splitlines()) for x, y in zip(a.split(), c.split()))
    ref_nml = checkpoint.read(os.path.join(fx.GOLDEN, "ref_InitialConditions_2005_EndofRun.nml"))
    ours = checkpoint.read_text(nml)
    worst = max(abs(ours[k] - r) / max(abs(r), 1e-3) for k, r in ref_nml.items())
    print("SUEWS_60.txt: %d of %d printed fields differ from the reference's file (%s)" % (
        bad, fields, "byte-identical" if txt == gold else "not byte-identical"))
    print("EndofRun.nml: largest relative difference of the 54 state values %.2e" % worst)
    return bad, fields, worst


def year_2004(outdir):
    """The 2004 leg of the same golden run: 105 408 steps, then the checkpoint the next year starts from."""
    from suews_b200 import OUT_SUM7
    os.makedirs(outdir, exist_ok=True)
    b = SuewsBatch(1, 1)
    b.load_reference_args(fx.state_init(2004))
    b.upload()
    nsteps = b.set_forcing_orig(read_met(os.path.join(fx.GOLDEN, "forcing_2004_60min.txt")), resolution_in=3600)
    for day0 in range(0, 366, 61):
        b.advance(day0 * 288, 61 * 288, OUT_SUM7)
    nml = checkpoint.write_text(checkpoint.state_vector(lambda n: b.get_state(n)[0]), snow_use=1)
    open(os.path.join(outdir, "InitialConditionstest1_2005.nml"), "w").write(nml)
    b.close()
    ref_text = open(os.path.join(fx.GOLDEN, "ref_InitialConditions_2005.nml")).read()
    ref_nml, ours = checkpoint.read_text(ref_text), checkpoint.read_text(nml)
    diffs = {k: abs(ours[k] - r) / max(abs(r), 1e-3) for k, r in ref_nml.items()}
    worst = max(diffs, key=diffs.get)
    print("2004 (%d steps) -> InitialConditions_2005.nml: %d of 54 values bit-identical to the reference's, largest relative "
          "difference %.2e (%s)%s" % (nsteps, sum(v == 0 for v in diffs.values()), diffs[worst], worst,
                                      ", file byte-identical" if nml == ref_text else ""))
    return diffs


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "scratch", "baserun_2005_out")
    main(out)
    if "--year-2004" in sys.argv:
        year_2004(out)
