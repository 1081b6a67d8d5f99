"""InitialConditions*.nml: the reference's checkpoint between years / at the end of a run.

Writer  = NextInitial (SUEWS-SourceCode/suews_ctrl_init.f95:2022-2149): list-directed WRITE of the state a following
          run starts from.  The text is produced by the library (suews_write_initial_conditions, include/suews_b200.h)
          in gfortran's list-directed layout, so a file written here is byte-identical to the reference's when the
          state is.
Reader  = the namelist those files are read back with (InitialConditions, suews_ctrl_init.f95:1443-1560): here a plain
          name -> value parser plus the mapping of the 1:1 entries onto SUEWS_cal_Main state arguments.
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
