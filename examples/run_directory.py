#!/usr/bin/env python3
"""A reference run directory through the CUDA library, end to end: RunControl.nml + Input/SUEWS_*.txt (the table join),
the user's InitialConditions namelist (InitialState), CBLInput.nml + its initial-data file (BLUEWS), the hourly forcing file
(MetRead + DisaggregateMet on the device) -> one year of five-minute steps on the GPU -> the reference's output files:

    <code>1_<year>_SUEWS_60.txt, _BL_60.txt   hourly output (SUEWS_Output_Agg + SUEWS_Write_txt)
    <code>1_<year>_DailyState.txt             one row per day
    InitialConditions<code>1_<year+1>.nml     the checkpoint the next year starts from (NextInitial)

With no arguments it runs the 2004 leg of the reference's own golden run (Test/BaseRun/2020b, kept under tests/golden/) and
compares what it wrote with the files the reference wrote.  Needs a CUDA device.

Every grid SUEWS_SiteSelect.txt lists for the year runs in the same batch (tables.grid_arguments), one set of files per grid.

    python examples/run_directory.py [run_dir year forcing_file [outdir]]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))

from baserun_2005 import write_txt  # noqa: E402
from suews_b200 import OUT_FULL, SuewsBatch, checkpoint, tables  # noqa: E402
from suews_b200.forcing import read_met  # noqa: E402


def hourly(block, group, nper=12):
    """SUEWS_Output_Agg on a per-step block (mean / sum / last per column, additions in row order from +0)"""
    import json
    vl = [v for v in json.load(open(os.path.join(ROOT, "suews_b200", "varlist.json"))) if v["group"] == group]
    n = block.shape[0] // nper
    out = np.zeros((n, block.shape[1]))
    out[:, :5] = block[nper - 1::nper, :5][:n]
    for c in range(block.shape[1] - 5):
        col = block[:n * nper, 5 + c].reshape(n, nper)
        s = np.cumsum(np.concatenate([np.zeros((n, 1)), col], axis=1), axis=1)[:, -1]
        a = vl[c]["aggreg"]
        out[:, 5 + c] = s / nper if a == "A" else (s if a == "S" else col[:, -1])
    return out


def run(run_dir, year, forcing_file, outdir, grids=None, days_per_launch=61):
    """Every grid SUEWS_SiteSelect.txt lists for `year` (or `grids`) as one batch -- the grid loop of the reference's main
    program -- under the one forcing file (MultipleMetFiles = 0); one set of output files per grid."""
    os.makedirs(outdir, exist_ok=True)
    ga = tables.grid_arguments(run_dir, year, grids)
    rc, ng = ga["runcontrol"], len(ga["grids"])
    if int(rc.get("multiplemetfiles", 0)) == 1 and ng > 1:
        raise NotImplementedError("MultipleMetFiles = 1: run the grids one by one, each with its own forcing file")
    cbl_on = ga["cbl_initial"] is not None
    b = tables.load_batch(SuewsBatch(ng, len(ga["classes"])), ga)
    b.upload()
    if cbl_on:
        b.set_cbl_initial(ga["cbl_initial"])
    tstep, res_in = int(rc.get("tstep", 300)), int(rc.get("resolutionfilesin", 3600))
    site = tuple(float(ga["cells"][k][0, 0]) for k in ("lat", "lng", "timezone", "alt"))
    if int(rc.get("kdownzen", 0)) == 1 and any(np.ptp(ga["cells"][k][:, 0]) != 0 for k in ("lat", "lng", "timezone", "alt")):
        raise NotImplementedError("KdownZen = 1 with grids at different sites: one batch per site")
    nsteps = b.set_forcing_orig(read_met(forcing_file), resolution_in=res_in, disagg_method=int(rc.get("disaggmethod", 1)),
                                rain_disagg_method=int(rc.get("raindisaggmethod", 100)), kdown_zen=int(rc.get("kdownzen", 0)),
                                site=site)
    per_day = 86400 // tstep
    groups = ("SUEWS", "DailyState") + (("BL",) if cbl_on else ())
    blocks = [{g: [] for g in groups} for _ in range(ng)]
    for s0 in range(0, nsteps, days_per_launch * per_day):
        n = min(days_per_launch * per_day, nsteps - s0)
        b.advance(s0, n, OUT_FULL)
        for i in range(ng):
            for g in groups:
                blocks[i][g].append(b.block(g, i, n))
    nper = int(rc.get("resolutionfilesout", 3600)) // tstep
    files = {}
    for i, grid in enumerate(ga["grids"]):
        blk = {g: np.vstack(v) for g, v in blocks[i].items()}
        code = "%s%d_%d" % (rc.get("filecode", "run"), grid, year)
        files[code + "_SUEWS_60.txt"] = write_txt(hourly(blk["SUEWS"], "SUEWS", nper), "SUEWS", int(rc.get("writeoutoption", 0)))
        if cbl_on:
            files[code + "_BL_60.txt"] = write_txt(hourly(blk["BL"], "BL", nper), "BL", 2)
        ds = blk["DailyState"]
        files[code + "_DailyState.txt"] = write_txt(ds[ds[:, 5] != -999.0], "DailyState", 2)
        files["InitialConditions%s%d_%d.nml" % (rc.get("filecode", "run"), grid, year + 1)] = checkpoint.write_text(
            checkpoint.state_vector(lambda n, i=i: b.get_state(n)[i]), snow_use=int(rc.get("snowuse", 0)))
    b.close()
    for name, text in files.items():
        open(os.path.join(outdir, name), "w").write(text)
    return files, nsteps


def golden_2004(outdir):
    """The 2004 leg of Test/BaseRun/2020b from its own input directory, against the files the reference wrote"""
    G = os.path.join(ROOT, "tests", "golden")
    files, nsteps = run(os.path.join(G, "baserun_2020b"), 2004, os.path.join(G, "forcing_2004_60min.txt"), outdir)
    report = {}
    for ours, ref in (("test1_2004_BL_60.txt", "ref_2004_BL_60.txt"), ("test1_2004_DailyState.txt", "ref_2004_DailyState.txt")):
        ref_text = open(os.path.join(G, ref)).read()
        a, r = files[ours].splitlines(), ref_text.splitlines()
        nf = sum(len(x.split()) for x in r[1:])
        bad = sum(x != y for la, lr in zip(a[1:], r[1:]) for x, y in zip(la.split(), lr.split())) + abs(len(a) - len(r))
        report[ours] = (bad, nf, files[ours] == ref_text)
        print("%s: %d of %d printed fields differ from the reference's file%s" % (ours, bad, nf, " (byte-identical)" if report[ours][2] else ""))
    ref_nml = checkpoint.read(os.path.join(G, "ref_InitialConditions_2005.nml"))
    mine = checkpoint.read_text(files["InitialConditionstest1_2005.nml"])
    diffs = {k: abs(mine[k] - v) / max(abs(v), 1e-3) for k, v in ref_nml.items()}
    report["checkpoint"] = (sum(v == 0 for v in diffs.values()), max(diffs.values()))
    print("InitialConditionstest1_2005.nml after %d steps: %d of 54 values bit-identical to the reference's, largest relative difference %.1e"
          % (nsteps, report["checkpoint"][0], report["checkpoint"][1]))
    return report


if __name__ == "__main__":
    if len(sys.argv) >= 4:
        run(sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "out")
    else:
        golden_2004(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "scratch", "run_directory_out"))
