"""Shared golden-fixture helpers (tests only)."""
import json
import os

import numpy as np

from suews_b200.forcing import disaggregate_met, read_met

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Arguments that the 2020b *_state_init.txt dump does not carry (it predates them); values from
# Test/BaseRun/2020b/RunControl.nml (BaseTMethod=1) and SUEWS_Irrigation.txt / SUEWS_SiteSelect.txt
# (no irrigation: H_maintain=-999, IrrFr_*=0) of the same BaseRun.
DUMP_DEFAULTS = dict(basetmethod=[1], h_maintain=[-999.0], irrfracpaved=[0.0], irrfracbldgs=[0.0],
                     irrfracbsoil=[0.0], irrfracwater=[0.0])


def state_init(year=2005):
    d = json.load(open(os.path.join(GOLDEN, "state_init_%d.json" % year)))
    d.update(DUMP_DEFAULTS)
    return d


def forcing(year=2005):
    """5-min MetForcingBlock rebuilt in double from the hourly file, as the binary does in memory."""
    return disaggregate_met(read_met(os.path.join(GOLDEN, "forcing_%d_60min.txt" % year)))


def load_ref_table(name):
    lines = open(os.path.join(GOLDEN, name)).readlines()
    hdr = lines[0].split()
    arr = np.array([[float(t) if "*" not in t else np.nan for t in ln.split()] for ln in lines[1:]])
    return hdr, arr


def varlist(group):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    vl = json.load(open(os.path.join(root, "suews_b200", "varlist.json")))
    return [v for v in vl if v["group"] == group]


def aggregate_hourly(block, group, nper=12):
    """SUEWS_Output_Agg (suews_ctrl_output.f95:913-957): A = mean, S = sum, L = last over each hour."""
    vl = varlist(group)
    n = block.shape[0] // nper
    out = np.zeros((n, block.shape[1] - 5))
    for c in range(block.shape[1] - 5):
        col = block[: n * nper, 5 + c].reshape(n, nper)
        a = vl[c]["aggreg"] if c < len(vl) else "A"
        out[:, c] = col.mean(axis=1) if a == "A" else (col.sum(axis=1) if a == "S" else col[:, -1])
    return out


def text_tolerance(group, ncol):
    """half a unit of the last printed decimal (f10.4 / f10.6 / f14.6, suews_ctrl_output.f95:39-46)"""
    vl = varlist(group)
    tol = np.full(ncol, 0.6e-4)
    for c in range(min(ncol, len(vl))):
        if vl[c]["fmt"].startswith(("f10.6", "f14.6")):
            tol[c] = 0.6e-6
    return tol


# ESTM example parameter set / synthetic surface-temperature forcing: shared with the bench's scheme sweep
from suews_b200.synth import C2K, estm_example, estm_ts_forcing  # noqa: E402,F401


def cbl_inputs():
    """BLUEWS inputs of the golden run (CBLInput.nml + CBLinputfiles/CBL_initial_data.txt): class fields and IniCBLdata"""
    from suews_b200 import tables
    params, ini = tables.read_cbl(os.path.join(GOLDEN, "baserun_2020b", "Input"))
    params["cbluse"] = [1]
    return params, ini
