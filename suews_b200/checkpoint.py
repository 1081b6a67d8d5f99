"""InitialConditions*.nml: the reference's checkpoint between years / at the end of a run.

Writer  = NextInitial (SUEWS-SourceCode/suews_ctrl_init.f95:2022-2149): list-directed WRITE of the state a following
          run starts from.  The text is produced by the library (suews_write_initial_conditions, include/suews_b200.h)
          in gfortran's list-directed layout, so a file written here is byte-identical to the reference's when the
          state is.
Reader  = the namelist those files are read back with (InitialConditions, suews_ctrl_init.f95:1443-1560): here a plain
          name -> value parser plus the mapping of the 1:1 entries onto SUEWS_cal_Main state arguments.
Fold    = initial_state(): what InitialState (:1443-1999) + SUEWS_Translate(ir = 0) (suews_ctrl_translate.f95:886-990) make
          of the namelist for a run's first timestep -- defaults, the LeavesOutInitially / SnowInitially shortcuts and the
          HDD / GDD / LAI arrays -- so that a batch can start from the file a SUEWS user writes.
"""
import ctypes as C
import re

import numpy as np

from . import _lib

SURF = ("Paved", "Bldgs", "EveTr", "DecTr", "Grass", "BSoil", "Water")


def state_vector(get):
    """The 54 values NextInitial writes, in its order, from a getter name -> 1-D array (SuewsBatch.get_state(name)[cell]
    or an oracle cell's get_state)."""
    hdd, gdd, lai = get("HDD_id"), get("GDD_id"), get("LAI_id")
    v = [float(int(hdd[11])), hdd[8], gdd[0], gdd[1], lai[0], lai[1], lai[2], get("albEveTr_id")[0], get("albDecTr_id")[0],
         get("albGrass_id")[0], get("DecidCap_id")[0], get("porosity_id")[0]]
    v += list(get("soilstore_id")[:6]) + list(get("state_id")[:7])
    for nm in ("SnowWater", "SnowPack", "SnowFrac", "SnowDens"):
        v += list(get(nm)[:7])
    v.append(get("SnowAlb")[0])
    return np.array(v, dtype=np.float64)


def write_text(values54, snow_use=1):
    """The namelist text for the 54 values of state_vector."""
    L = _lib.lib()
    f = L.suews_write_initial_conditions
    f.restype = C.c_longlong
    f.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_char_p, C.c_longlong]
    a = np.ascontiguousarray(values54, dtype=np.float64)
    assert a.size == 54
    buf = C.create_string_buffer(8192)
    n = f(a.ctypes.data_as(C.POINTER(C.c_double)), int(snow_use), buf, 8192)
    assert 0 < n < 8192
    return buf.value.decode()


def read(path):
    """name -> float for every `name = value` entry of an InitialConditions namelist."""
    return read_text(open(path).read())


def read_text(text):
    out = {}
    for ln in text.splitlines():
        m = re.match(r"\s*(\w+)\s*=\s*([-+0-9.EeDd]+)", ln)
        if m:
            out[m.group(1)] = float(m.group(2).replace("D", "E").replace("d", "e"))
    return out


def to_state(nml):
    """The namelist entries that are SUEWS_cal_Main state arguments as they stand (the others -- LAI, GDD, albedos,
    Temp_C0, DaysSinceRain -- are folded into HDD_id / GDD_id / LAI_id by InitialState, suews_ctrl_init.f95:1620-1850,
    which is outside the hot path)."""
    st = {
        "soilstore_id": [nml.get("SoilStore%sState" % s, 0.0) for s in SURF[:6]] + [0.0],
        "state_id": [nml.get("%sState" % s, 0.0) for s in SURF],
        "SnowWater": [nml.get("SnowWater%sState" % s, 0.0) for s in SURF],
        "SnowPack": [nml.get("SnowPack%s" % s, 0.0) for s in SURF],
        "SnowFrac": [nml.get("SnowFrac%s" % s, 0.0) for s in SURF],
        "SnowDens": [nml.get("SnowDens%s" % s, 0.0) for s in SURF],
        "SnowAlb": [nml.get("SnowAlb0", 0.0)],
        "albEveTr_id": [nml["AlbEveTr0"]], "albDecTr_id": [nml["AlbDecTr0"]], "albGrass_id": [nml["AlbGrass0"]],
        "DecidCap_id": [nml["DecidCap0"]], "porosity_id": [nml["Porosity0"]],
        "LAI_id": [nml["LAIinitialEveTr"], nml["LAIinitialDecTr"], nml["LAIinitialGrass"]],
    }
    return st


NAN = -999.0


def initial_state(nml, params, sfr=None, water_depth=None):
    """The SUEWS_cal_Main state arguments of the FIRST timestep of a run started from an InitialConditions namelist.

    InitialState (suews_ctrl_init.f95:1443-1999) fills in what the user left out, SUEWS_Translate(ir = 0)
    (suews_ctrl_translate.f95:886-990) turns the result into the model's variables:

    * LeavesOutInitially = 1 / 0: summer / winter values from the vegetation tables (:1640-1665) -- LAIMax / LAIMin,
      AlbMax / AlbMin, CapMax_dec / CapMin_dec, PorMin_dec / PorMax_dec, GDD_1_0 or GDD_2_0 = the vegetation-fraction
      weighted GDDFull / SDDFull (NormalizeVegChar, :2003-2025); -999 / absent: all of them must be given;
    * surface states default to 0, the water body to WaterDepth (:1672-1680); soil stores must be given (:1683-1686);
    * snowUse = 0 or SnowInitially = 0 zeroes the snow entries (:1690-1719);
    * GDD_id(:) = GDD_1_0, SDD_id(:) = GDD_2_0, Tmin_id = 90, Tmax_id = -90, lenDay_id = 0, HDD_id = 0, WUDay_id = 0
      (translate :915-960; DaysSinceRain and Temp_C0 are read but no longer folded into HDD_id in this version);
    * qn1_av = dqndt = 0 and Tair_av = 273.15 (REAL(4)) at the start of a year (suews_program.f95:315-317), the clock at 0.

    params: the class / site values by reference argument name (any case): LAIMax, LAIMin, AlbMax_*, AlbMin_*, CapMax_dec,
    CapMin_dec, PorMax_dec, PorMin_dec, GDDFull, SDDFull, snowUse, tstep; sfr (7) for the vegetation weighting (default:
    params['sfr']); water_depth: WaterDepth of SUEWS_Water.txt (not a SUEWS_cal_Main argument).
    Returns {state argument name: list of values}; raises ValueError where the reference raises ErrorHint 36."""
    low = {k.lower(): v for k, v in nml.items()}
    P = {k.lower(): np.atleast_1d(np.asarray(v, dtype=np.float64)).ravel() for k, v in params.items()}
    sfr = np.asarray(P["sfr"] if sfr is None else sfr, dtype=np.float64).ravel()

    def g(name):
        return low.get(name.lower(), NAN)

    def norm_veg(col):   # NormalizeVegChar: weighted by the three vegetation fractions (EveTr, DecTr, Grass)
        fv = sfr[2:5]
        if fv.sum() == 0.0:
            return 0.0
        return float((col[0] * fv[0] + col[1] * fv[1] + col[2] * fv[2]) / (fv[0] + fv[1] + fv[2]))

    leaves = int(g("LeavesOutInitially"))
    veg = {k: g(k) for k in ("GDD_1_0", "GDD_2_0", "LAIinitialEveTr", "LAIinitialDecTr", "LAIinitialGrass", "AlbEveTr0",
                             "AlbDecTr0", "AlbGrass0", "DecidCap0", "Porosity0")}
    if leaves == int(NAN):
        if any(v == NAN for v in veg.values()):
            raise ValueError("ErrorHint 36: LeavesOutInitially is not set, so GDD_1_0, GDD_2_0, the LAI and albedo of the three "
                             "vegetated surfaces, DecidCap0 and Porosity0 must all be given")
    elif leaves in (0, 1):
        hi = leaves == 1
        veg["GDD_1_0"] = norm_veg(P["gddfull"]) if hi else 0.0
        veg["GDD_2_0"] = 0.0 if hi else norm_veg(P["sddfull"])
        lai = P["laimax"] if hi else P["laimin"]
        veg["LAIinitialEveTr"], veg["LAIinitialDecTr"], veg["LAIinitialGrass"] = float(lai[0]), float(lai[1]), float(lai[2])
        sfx = "albmax_" if hi else "albmin_"
        veg["AlbEveTr0"], veg["AlbDecTr0"], veg["AlbGrass0"] = (float(P[sfx + "evetr"][0]), float(P[sfx + "dectr"][0]),
                                                                float(P[sfx + "grass"][0]))
        veg["DecidCap0"] = float(P["capmax_dec" if hi else "capmin_dec"][0])
        veg["Porosity0"] = float(P["pormin_dec" if hi else "pormax_dec"][0])
    else:
        raise ValueError("ErrorHint 36: LeavesOutInitially must be 0, 1, or -999")

    state = [g("%sState" % s) for s in SURF]
    state = [0.0 if v == NAN else v for v in state[:6]] + [state[6]]
    if state[6] == NAN:
        if water_depth is None:
            raise ValueError("WaterState is not in the namelist: pass water_depth (WaterDepth of SUEWS_Water.txt)")
        state[6] = float(water_depth)
    soil = [g("SoilStore%sState" % s) for s in SURF[:6]]
    if any(v == NAN for v in soil):
        raise ValueError("ErrorHint 36: initial soil moisture must be provided for all surface types except water")

    snow_use = int(P["snowuse"][0])
    snow_init = int(g("SnowInitially"))
    snow = {k: [g(fmt % s) for s in SURF] for k, fmt in (("SnowWater", "SnowWater%sState"), ("SnowPack", "SnowPack%s"),
                                                          ("SnowFrac", "SnowFrac%s"), ("SnowDens", "SnowDens%s"))}
    snow_alb = g("SnowAlb0")
    if snow_use == 0 or snow_init == 0:
        snow = {k: [0.0] * 7 for k in snow}
        snow_alb = 0.0
    elif snow_init == int(NAN):
        if any(v == NAN for vals in snow.values() for v in vals) or snow_alb == NAN:
            raise ValueError("ErrorHint 36: SnowWater, SnowPack, SnowFrac, SnowDens of all 7 surfaces and SnowAlb0 must be given")
    else:
        raise ValueError("ErrorHint 36: SnowInitially must be 0 or -999")

    tstep = int(P["tstep"][0]) if "tstep" in P else 300
    return {
        "state_id": state, "soilstore_id": soil + [0.0],
        "SnowWater": snow["SnowWater"], "SnowPack": snow["SnowPack"], "SnowFrac": snow["SnowFrac"], "SnowDens": snow["SnowDens"],
        "SnowAlb": [snow_alb], "SnowfallCum": [0.0],
        # (InitialState sets the module's IceFrac to 0.2, but the first step reads IceFrac_grids, still zero)
        "IceFrac": [0.0] * 7,
        "LAI_id": [veg["LAIinitialEveTr"], veg["LAIinitialDecTr"], veg["LAIinitialGrass"]],
        "GDD_id": [veg["GDD_1_0"]] * 3, "SDD_id": [veg["GDD_2_0"]] * 3,
        "Tmin_id": [90.0], "Tmax_id": [-90.0], "lenDay_id": [0.0],
        "albEveTr_id": [veg["AlbEveTr0"]], "albDecTr_id": [veg["AlbDecTr0"]], "albGrass_id": [veg["AlbGrass0"]],
        "DecidCap_id": [veg["DecidCap0"]], "porosity_id": [veg["Porosity0"]],
        "HDD_id": [0.0] * 12, "WUDay_id": [0.0] * 9,
        "qn1_av": [0.0], "dqndt": [0.0], "qn1_s_av": [0.0], "dqnsdt": [0.0], "Tair_av": [float(np.float32(273.15))],
        "dt_since_start": [0], "tstep_prev": [tstep],
    }
