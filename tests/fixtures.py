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


C2K = float(np.float32(273.15))   # REAL(KIND(1d0)), PARAMETER :: C2K = 273.15 (a REAL(4) literal)


def estm_example(surface_area, **over):
    """An ESTM (StorageHeatMethod 4) parameter set as the reference's SUEWS_Translate / ESTM_initials leave it in
    MODULE ESTM_data for the BaseRun site: building class 801 and paved class 807 of
    Test/BaseRun/2020b/Input/SUEWS_ESTMCoefficients.txt, ESTMinput.nml of the same directory, the internal view
    factors and floor temperature hard-coded in ESTM_translate (suews_phys_estm.f95:673-692).  AreaWall is -999 in
    that SiteSelect row (ESTM is off in the BaseRun); 0.4 x SurfaceArea is used here."""
    def pad(v):
        return list(v) + [-999.0] * (5 - len(v))
    d = dict(
        evolveTibld=[0], TsurfChoice=[0], ibldCHmod=[0], Nibld=[3], Nwall=[3], Nroof=[3], Nground=[4],
        LBC_soil=[8.5], THEAT_ON=[18.0 + C2K], THEAT_OFF=[22.0 + C2K], THEAT_fix=[19.0 + C2K],
        ivf_iw=[float(np.float32(0.1))], ivf_ir=[0.0], ivf_ii=[float(np.float32(0.9))], ivf_if=[0.0],
        ivf_ww=[float(np.float32(0.05))], ivf_wr=[0.0], ivf_wi=[float(np.float32(0.95))], ivf_wf=[0.0],
        ivf_rw=[float(np.float32(0.05))], ivf_ri=[float(np.float32(0.95))], ivf_rf=[0.0],
        ivf_fw=[float(np.float32(0.05))], ivf_fr=[0.0], ivf_fi=[float(np.float32(0.95))],
        zibld=pad([0.035] * 3), kibld=pad([0.93] * 3), ribld=pad([1500000.0] * 3),
        zwall=pad([0.087, 0.0497, 0.0497]), kwall=pad([0.760645, 0.93, 0.93]), rwall=pad([1046908.363, 1500000.0, 1500000.0]),
        zroof=pad([0.05, 0.15, 0.12]), kroof=pad([1.142, 0.93, 0.05]), rroof=pad([1606620.0, 1500000.0, 290000.0]),
        zground=pad([0.05, 0.10, 0.5, 0.7]), kground=pad([1.16, 1.16, 0.63, 0.63]),
        rground=pad([2285607.0, 2285607.0, 1200000.0, 1200000.0]),
        nroom=[10.0], alb_ibld=[0.5], em_ibld=[1.0], CH_iroof=[0.0015], CH_iwall=[0.0015], CH_ibld=[0.0015],
        AreaWall=[0.4 * float(surface_area)], Tfloor=[float(np.float32(20.0)) + C2K],
    )
    for k, v in over.items():
        d[k] = list(np.atleast_1d(v))
    return d


def estm_ts_forcing(met, tiair=20.0, offsets=(0.0, 1.0, 2.0, 0.5, -0.5, 0.3, 0.8, 0.1)):
    """Synthetic ESTM surface-temperature forcing [len_sim, 13] (iy id it imin Tiair Tsurf Troof Troad Twall
    Twall_n Twall_e Twall_s Twall_w): the air temperature of the met block plus a fixed offset per column
    ("synthetic Ts forcing = air temperature")."""
    met = np.asarray(met)
    ts = np.zeros((met.shape[0], 13))
    ts[:, :4] = met[:, :4]
    ts[:, 4] = tiair
    for j, off in enumerate(offsets):
        ts[:, 5 + j] = met[:, 11] + off
    return ts
