"""ESTM (StorageHeatMethod 4/14) host-side helpers.

The reference initialises ESTM's element temperatures on the first call of ESTM_translate (ESTMStart == 1,
suews_phys_estm.f95:665-716 and :941-960) from the first row of the surface-temperature forcing; that is the step
before the hot path, done once per run, so it lives on the host.  Arithmetic in double, in the reference's order.
"""
import numpy as np

C2K = float(np.float32(273.15))          # suews_ctrl_const.f95:2086 (a REAL(4) literal)
SBCONST = float(np.float32(5.67051e-8))  # suews_ctrl_const.f95:2087
EM_WALL = float(np.float32(0.9))         # suews_phys_estm.f95:169

PAV, BLDG, CONIF, DECID, GRASS, BSOIL, WATER = range(7)


def _p4(x):
    a = x * x
    return a * a


def initial_state(params, sfr, emis, ts_row):
    """Element temperatures and longwave terms as ESTM_translate leaves them on its first call.

    params: dict with Nground, Nwall, Nroof, Nibld, LBC_soil; sfr, emis: the 7 surface fractions / emissivities;
    ts_row: first row [13] of the ESTM forcing (iy id it imin Tiair Tsurf Troof Troad Twall ...).
    Returns {state name: values} for SuewsBatch.set_state."""
    g = lambda k: float(np.atleast_1d(params[k])[0])   # noqa: E731
    ng, nw, nr, ni = (int(g(k)) for k in ("Nground", "Nwall", "Nroof", "Nibld"))
    lbc = g("LBC_soil")
    ts = [float(v) for v in ts_row]
    tiair, troof, troad, twall = ts[4], ts[6], ts[7], ts[8]
    sfr = [float(v) for v in sfr]
    emis = [float(v) for v in emis]
    st = {k: [0.0] * 5 for k in ("Tground", "Twall", "Troof", "Tibld")}
    for i in range(1, ng + 1):
        st["Tground"][i - 1] = (lbc - troad) * (i - 1) / (ng - 1) + troad + C2K
    for i in range(1, nw + 1):
        st["Twall"][i - 1] = (tiair - twall) * (i - 1) / (nw - 1) + twall + C2K
    for i in range(1, nr + 1):
        st["Troof"][i - 1] = (tiair - troof) * (i - 1) / (nr - 1) + troof + C2K
    for i in range(ni):
        st["Tibld"][i] = tiair + C2K
    tw4 = [0.0] * 20
    for w in range(4):
        for i in range(nw):
            tw4[i + 5 * w] = st["Twall"][i]
    st["Tw_4"] = tw4
    st["Tievolve"] = [float(np.float32(20.0)) + C2K]
    st["T0_ground"] = [st["Tground"][0]]
    st["T0_wall"] = [st["Twall"][0]]
    st["T0_roof"] = [st["Troof"][0]]
    st["T0_ibld"] = [st["Tibld"][0]]
    st["TN_roof"] = [st["Troof"][nr - 1]]
    st["TN_wall"] = [st["Twall"][nw - 1]]
    fground = sfr[PAV] + sfr[CONIF] + sfr[DECID] + sfr[GRASS] + sfr[BSOIL] + sfr[WATER]
    if fground != 0:
        em_ground = (emis[CONIF] * sfr[CONIF] + emis[DECID] * sfr[DECID] + emis[GRASS] * sfr[GRASS] + emis[PAV] * sfr[PAV] +
                     emis[BSOIL] * sfr[BSOIL] + emis[WATER] * sfr[WATER]) / fground
    else:
        em_ground = -999.0
    st["lup_ground"] = [SBCONST * em_ground * _p4(st["T0_ground"][0])]
    st["lup_wall"] = [SBCONST * EM_WALL * _p4(st["T0_wall"][0])]
    st["lup_roof"] = [SBCONST * emis[BLDG] * _p4(st["T0_roof"][0])]
    for k in ("Tair2", "HVAC", "Tair2Set", "NewtonConverged"):
        st[k] = [0.0]
    return st
