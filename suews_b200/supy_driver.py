"""Host-side mirror of the reference's f2py module ``supy_driver.suews_driver``
(supy-driver/setup.py:29-50 wraps suews_ctrl_driver.f95; supy calls it with keyword arguments).

    from suews_b200.supy_driver import suews_driver as sd
    blocks = sd.suews_cal_multitsteps(metforcingblock=..., len_sim=..., alb=..., ...)

Same names (f2py lower-cases them; any case is accepted here), same argument meaning, same return
convention: INTENT(OUT) arrays come back as a tuple in declaration order, INTENT(INOUT) arguments are
numpy arrays updated in place, a fatal SUEWS condition raises (the reference STOPs; nonstopf2py.py:54-60
turns that into an exception).  All computation happens in libsuews_b200.so on the GPU through
``suews_cal_multitsteps_c`` / ``suews_cal_main_c`` (include/suews_b200.h); without the library or a CUDA
device these calls raise -- there is no CPU path.
"""
import ctypes as C

import numpy as np

from ._lib import SuewsError, last_error, lib

_TABLES = {}


def arg_table(which):
    """[(name, is_int, intent, elements, per_row)] of SUEWS_cal_Main (0) / SUEWS_cal_multitsteps (1), in the
    reference's positional order (generated table suews_b200/csrc/suews_refargs.inc)."""
    if which not in _TABLES:
        L = lib()
        f = L.suews_refapi_arginfo
        f.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_char_p)] + [C.POINTER(C.c_int)] * 4
        f.restype = C.c_int
        rows, i = [], 0
        while True:
            nm = C.c_char_p()
            a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            if f(which, i, C.byref(nm), C.byref(a), C.byref(b), C.byref(c), C.byref(d)) != 0:
                break
            rows.append((nm.value.decode(), bool(a.value), b.value, c.value, bool(d.value)))
            i += 1
        _TABLES[which] = rows
    return _TABLES[which]


class _SuewsDriver:
    """Stands where ``supy_driver.suews_driver`` stands."""

    def _call(self, cname, which, kw, len_sim=1):
        L = lib()
        low = {k.lower(): v for k, v in kw.items()}
        keep, ptrs, results = [], [], []
        for nm, is_int, intent, n, per_row in arg_table(which):
            key = nm.lower()
            if intent == 2:
                a = np.full((len_sim, n) if per_row else (n,), -999.0, dtype=np.float64, order="F")
                results.append(a)
            elif n == 0 and key not in low:      # assumed-shape AnOHM/ESTM inputs the path never reads
                ptrs.append(None)
                continue
            else:
                if key not in low:
                    raise TypeError("%s() missing required argument '%s'" % (cname[:-2], key))
                v = low[key]
                dt = np.int32 if is_int else np.float64
                if isinstance(v, np.ndarray) and v.dtype == dt and (v.flags.f_contiguous or v.ndim <= 1):
                    a = v          # by reference: INTENT(INOUT) updates land in the caller's array
                else:
                    a = np.asfortranarray(np.asarray(v, dtype=dt))
                want = n * (len_sim if per_row else 1)
                if n > 0 and a.size != want:
                    raise ValueError("%s: argument '%s' has %d elements, the reference declares %d" % (cname[:-2], key, a.size, want))
            keep.append(a)
            ptrs.append(a.ctypes.data)
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        f = getattr(L, cname)
        f.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        f.restype = C.c_int
        rc = f(arr, len(ptrs))
        if rc != 0:
            code, msg = last_error()
            raise SuewsError(rc, msg)
        return tuple(results)

    def _positional(self, which, args, kw, optional=()):
        """f2py call convention: positional arguments are the Fortran dummy arguments in declaration order without the
        INTENT(OUT) ones (driver.f95:45-91 / :2814-2860); arguments f2py makes optional move behind them."""
        if not args:
            return kw
        names = [nm.lower() for nm, _i, intent, _n, _r in arg_table(which) if intent != 2 and nm.lower() not in optional]
        names += list(optional)
        if len(args) > len(names):
            raise TypeError("takes at most %d positional arguments (%d given)" % (len(names), len(args)))
        out = dict(kw)
        for nm, v in zip(names, args):
            if nm in {k.lower() for k in kw}:
                raise TypeError("got multiple values for argument '%s'" % nm)
            out[nm] = v
        return out

    def suews_cal_multitsteps(self, *args, **kw):
        """SUEWS_cal_multitsteps (driver.f95:2814): returns (dataoutblocksuews, dataoutblocksnow, dataoutblockestm,
        dataoutblockrsl, dataoutblocksol, dailystateblock)."""
        low = {k.lower(): v for k, v in self._positional(1, args, kw, optional=("len_sim",)).items()}
        met = np.asfortranarray(np.asarray(low["metforcingblock"], dtype=np.float64))
        low["metforcingblock"] = met
        low.setdefault("len_sim", met.shape[0])   # f2py makes len_sim optional (= shape(metforcingblock, 0))
        return self._call("suews_cal_multitsteps_c", 1, low, len_sim=int(low["len_sim"]))

    def suews_cal_main(self, *args, **kw):
        """SUEWS_cal_Main (driver.f95:45): returns (datetimeline, dataoutlinesuews, dataoutlinesnow, dataoutlineestm,
        dataoutlinersl, dataoutlinesolweig, dailystateline).  Positional (f2py order) or keyword arguments."""
        return self._call("suews_cal_main_c", 0, self._positional(0, args, kw))

    def output_size(self):
        """output_size (driver.f95:2802): SIZE(varListAll) = 500."""
        n = C.c_int()
        lib().suews_output_size(C.byref(n))
        return n.value

    def output_name_n(self, i):
        """output_name_n (driver.f95:2778): (name, group, aggreg, outlevel); strings are 15-byte blank-padded
        bytes, as f2py returns CHARACTER(len=15)."""
        L = lib()
        L.suews_output_name_n.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
        a, b, c = (C.create_string_buffer(16) for _ in range(3))
        lvl = C.c_int()
        L.suews_output_name_n(int(i), a, b, c, C.byref(lvl))
        return tuple(x.value.ljust(15) for x in (a, b, c)) + (lvl.value,)

    def suews_cal_sunposition(self, year, idectime, utc, locationlatitude, locationlongitude, locationaltitude):
        """SUEWS_cal_sunposition (driver.f95:3508): returns (sunazimuth, sunzenith)."""
        az, zen = C.c_double(), C.c_double()
        rc = lib().suews_cal_sunposition_c(float(year), float(idectime), float(utc), float(locationlatitude),
                                           float(locationlongitude), float(locationaltitude), C.byref(az), C.byref(zen))
        if rc != 0:
            code, msg = last_error()
            raise SuewsError(rc, msg)
        return az.value, zen.value

    # ---- the small pure functions f2py also exposes on suews_driver (scalar arithmetic, host side)
    @staticmethod
    def set_nan(x):
        """set_nan (driver.f95:2737): values beyond +-30000 become -999."""
        a = np.asarray(x, dtype=np.float64)
        r = np.where(np.abs(a) > 30000.0, -999.0, a)
        return float(r) if r.ndim == 0 else r

    @staticmethod
    def square(x):
        """square (driver.f95:2753): the f2py conversion test function."""
        return float(x) * float(x)

    @staticmethod
    def square_real(x):
        """square_real (driver.f95:2765): the REAL(4) twin of square."""
        return float(np.float32(x) * np.float32(x))

    @staticmethod
    def cal_tair_av(tair_av_prev, dt_since_start, tstep, temp_c):
        """cal_tair_av (driver.f95:3530): running 24 h mean air temperature [K]; 273.15 and tstep*1. are REAL(4)."""
        len_day_s = 24.0 * 3600.0
        len_cal_s = len_day_s if dt_since_start > len_day_s else float(int(dt_since_start) + int(tstep))
        temp_k = float(temp_c) + float(np.float32(273.15))
        return float(tair_av_prev) * (len_cal_s - float(np.float32(tstep) * np.float32(1.0))) / len_cal_s + temp_k * int(tstep) / len_cal_s

    @staticmethod
    def cal_tsfc(qh, avdens, avcp, ra, temp_c):
        """cal_tsfc (driver.f95:3558): surface temperature from the sensible heat flux [degC]."""
        return float(qh) / (float(avdens) * float(avcp)) * float(ra) + float(temp_c)

    def reset_hidden_state(self):
        """A fresh reference process starts with the SAVE'd Lat_vap switches and NARP alb1 cleared (and an empty
        MODULE ESTM_data)."""
        lib().suews_refapi_reset_hidden()

    # MODULE ESTM_data (suews_phys_estm.f95:1-176).  SUEWS_cal_Main reads ESTM's parameters and element temperatures
    # from that module, not from its arguments; the reference's hosts fill it (ESTM_initials, SUEWS_Translate,
    # ESTM_translate) before calling.  f2py would expose the module as an attribute object; these two calls stand there.
    def estm_data_set(self, **kw):
        L = lib()
        L.suews_estm_data_set.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int]
        for name, v in kw.items():
            a = np.ascontiguousarray(np.atleast_1d(np.asarray(v, dtype=np.float64)).ravel())
            rc = L.suews_estm_data_set(name.encode(), a.ctypes.data_as(C.POINTER(C.c_double)), a.size)
            if rc != 0:
                raise SuewsError(rc, "estm_data_set: '%s' (%d values) is not an ESTM_data variable of that size" % (name, a.size))

    def anohm_set_met_grid(self, met=None):
        """MetForcingData_grid for calls with StorageHeatMethod 3 (AnOHM): the grid's forcing block [nrows, 24], which the
        reference passes as an assumed-shape argument (suews_ctrl_driver.f95:277); None forgets it."""
        L = lib()
        L.suews_anohm_set_met_grid.argtypes = [C.POINTER(C.c_double), C.c_int]
        if met is None:
            L.suews_anohm_set_met_grid(None, 0)
            return
        m = np.asfortranarray(np.asarray(met, dtype=np.float64))
        if m.ndim != 2 or m.shape[1] != 24:
            raise ValueError("MetForcingData_grid is [nrows, 24]")
        L.suews_anohm_set_met_grid(m.ctypes.data_as(C.POINTER(C.c_double)), m.shape[0])

    def estm_data_get(self, name, n=1):
        L = lib()
        L.suews_estm_data_get.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int]
        a = np.zeros(int(n))
        rc = L.suews_estm_data_get(name.encode(), a.ctypes.data_as(C.POINTER(C.c_double)), a.size)
        if rc != 0:
            raise SuewsError(rc, "estm_data_get: '%s' (%d values) has not been set" % (name, a.size))
        return a


suews_driver = _SuewsDriver()


class _FortranModule:
    """Stands where f2py puts a Fortran MODULE of the reference (supy_driver/__init__.py:13-27 re-exports twelve of them
    beside suews_driver).  supy itself calls only suews_driver; the stage routines of these modules run inside the kernel
    and have no host entry point of their own -- asking for one says so instead of silently computing on the CPU."""

    def __init__(self, name, source, routines=None):
        self.__name__ = name
        self._source = source
        self._routines = dict(routines or {})

    def __getattr__(self, item):
        r = self.__dict__.get("_routines", {})
        if item.lower() in r:
            return r[item.lower()]
        raise NotImplementedError(
            "%s.%s: the routines of %s run inside the CUDA timestep kernel (libsuews_b200.so); only suews_driver's entry "
            "points are callable from the host%s" % (self.__name__, item, self._source,
                                                     (" and %s" % ", ".join(sorted(r))) if r else ""))

    def __dir__(self):
        return sorted(self._routines)


# NARP_cal_SunPosition (narp.f95:460) is the one stage routine with its own C entry point
narp_module = _FortranModule("narp_module", "suews_phys_narp.f95",
                             {"narp_cal_sunposition": suews_driver.suews_cal_sunposition})
atmmoiststab_module = _FortranModule("atmmoiststab_module", "suews_phys_atmmoiststab.f95")
resist_module = _FortranModule("resist_module", "suews_phys_resist.f95")
evap_module = _FortranModule("evap_module", "suews_phys_evap.f95")
snow_module = _FortranModule("snow_module", "suews_phys_snow.f95")
dailystate_module = _FortranModule("dailystate_module", "suews_phys_dailystate.f95")
lumps_module = _FortranModule("lumps_module", "suews_phys_lumps.f95")
anemsn_module = _FortranModule("anemsn_module", "suews_phys_anemsn.f95")
rsl_module = _FortranModule("rsl_module", "suews_phys_rslprof.f95")
co2_module = _FortranModule("co2_module", "suews_phys_biogenco2.f95")
waterdist_module = _FortranModule("waterdist_module", "suews_phys_waterdist.f95")
meteo = _FortranModule("meteo", "suews_util_meteo.f95")
__version__ = "2020b8+b200"
