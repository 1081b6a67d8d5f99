"""ctypes access to the CPU oracle (oracle/libsuews_oracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libsuews_oracle.so")

NCOL = {"SUEWS": 87, "snow": 103, "ESTM": 32, "RSL": 138, "SOLWEIG": 31, "DailyState": 50}

_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.so_params_new.restype = C.c_void_p
        L.so_state_new.restype = C.c_void_p
        L.so_free.argtypes = [C.c_void_p]
        L.so_state_copy.argtypes = [C.c_void_p, C.c_void_p]
        for fn in (L.so_params_set, L.so_state_set, L.so_params_get, L.so_state_get):
            fn.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
            fn.restype = C.c_int
        L.so_params_field.restype = C.c_char_p
        L.so_params_field.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.so_state_field.restype = C.c_char_p
        L.so_state_field.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        dp = C.POINTER(C.c_double)
        L.so_run.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_int, dp, dp, dp, dp, dp, dp, C.c_char_p, C.c_int,
                             C.POINTER(C.c_long), C.POINTER(C.c_int)]
        L.so_run.restype = C.c_int
        L.so_run_cells.argtypes = [C.c_int, C.c_void_p, C.c_void_p, dp, C.c_long, C.c_int, dp]
        L.so_run_cells.restype = C.c_int
        L.so_sizeof_params.restype = C.c_size_t
        L.so_sizeof_state.restype = C.c_size_t
        L.so_NARP_cal_SunPosition.argtypes = [C.c_double] * 6 + [dp, dp]
        L.so_set_ts5_block.argtypes = [dp]
        L.so_estm_start.argtypes = [C.c_void_p, C.c_void_p, dp]
        _lib = L
    return _lib


def _fields(kind):
    L = lib()
    n = L.so_params_nfields() if kind == "params" else L.so_state_nfields()
    get = L.so_params_field if kind == "params" else L.so_state_field
    out = []
    for i in range(n):
        is_int, cnt = C.c_int(), C.c_int()
        nm = get(i, C.byref(is_int), C.byref(cnt)).decode()
        out.append((nm, bool(is_int.value), cnt.value))
    return out


def param_fields():
    return _fields("params")


def state_fields():
    return _fields("state")


class Cell:
    """One grid cell: SoParams + SoState filled by (case-insensitive) reference argument names."""

    def __init__(self, values=None):
        L = lib()
        self.p = C.c_void_p(L.so_params_new())
        self.s = C.c_void_p(L.so_state_new())
        if values is not None:
            self.load(values)

    def __del__(self):
        try:
            lib().so_free(self.p)
            lib().so_free(self.s)
        except Exception:
            pass

    def load(self, values):
        """values: dict name(lower) -> sequence.  Unknown keys are ignored, missing keys left 0."""
        L = lib()
        low = {k.lower(): v for k, v in values.items()}
        missing = []
        for kind, fields, setter, base in (("p", param_fields(), L.so_params_set, self.p),
                                           ("s", state_fields(), L.so_state_set, self.s)):
            for nm, _is_int, cnt in fields:
                if nm.lower() not in low:
                    missing.append(nm)
                    continue
                arr = np.ascontiguousarray(np.atleast_1d(np.asarray(low[nm.lower()], dtype=np.float64)).ravel())
                assert arr.size == cnt, (nm, arr.size, cnt)
                rc = setter(base, nm.encode(), arr.ctypes.data_as(C.POINTER(C.c_double)), cnt)
                assert rc == 0, (nm, rc)
        return missing

    def set_param(self, name, vals):
        arr = np.ascontiguousarray(np.atleast_1d(np.asarray(vals, dtype=np.float64)).ravel())
        rc = lib().so_params_set(self.p, name.encode(), arr.ctypes.data_as(C.POINTER(C.c_double)), arr.size)
        assert rc == 0, (name, rc)

    def set_state(self, name, vals):
        arr = np.ascontiguousarray(np.atleast_1d(np.asarray(vals, dtype=np.float64)).ravel())
        rc = lib().so_state_set(self.s, name.encode(), arr.ctypes.data_as(C.POINTER(C.c_double)), arr.size)
        assert rc == 0, (name, rc)

    def get_state(self, name):
        for nm, _i, cnt in state_fields():
            if nm.lower() == name.lower():
                arr = np.zeros(cnt)
                rc = lib().so_state_get(self.s, nm.encode(), arr.ctypes.data_as(C.POINTER(C.c_double)), cnt)
                assert rc == 0
                return arr
        raise KeyError(name)

    def state_dict(self):
        return {nm: self.get_state(nm) for nm, _i, _c in state_fields()}

    def estm_start(self, ts_row):
        """First-call initialisation of the ESTM element temperatures (ESTM_translate, ESTMStart == 1) from the
        first row [13] of the surface-temperature forcing."""
        row = np.ascontiguousarray(np.asarray(ts_row, dtype=np.float64))
        assert row.size == 13
        lib().so_estm_start(self.p, self.s, row.ctypes.data_as(C.POINTER(C.c_double)))

    def run(self, met, want=("SUEWS", "snow", "RSL", "DailyState"), ts5=None):
        """met: [len_sim, 24] array in the reference's MetForcingBlock column order; ts5: [len_sim, 13] ESTM
        surface-temperature forcing (only read when StorageHeatMethod is 4/14).
        Returns dict of output blocks [len_sim, ncol] (datetime columns included)."""
        met = np.asfortranarray(np.asarray(met, dtype=np.float64))
        n = met.shape[0]
        if ts5 is not None:
            ts5 = np.asfortranarray(np.asarray(ts5, dtype=np.float64))
            assert ts5.shape == (n, 13)
        lib().so_set_ts5_block(ts5.ctypes.data_as(C.POINTER(C.c_double)) if ts5 is not None else None)
        blocks = {}
        ptrs = []
        for key in ("SUEWS", "snow", "ESTM", "RSL", "SOLWEIG", "DailyState"):
            if key in want:
                b = np.full((n, NCOL[key]), -999.0, order="F")
                blocks[key] = b
                ptrs.append(b.ctypes.data_as(C.POINTER(C.c_double)))
            else:
                ptrs.append(None)
        msg = C.create_string_buffer(200)
        warn = (C.c_long * 128)()
        nit = C.c_int()
        rc = lib().so_run(self.p, self.s, met.ctypes.data_as(C.POINTER(C.c_double)), n, *ptrs, msg, 200, warn,
                          C.byref(nit))
        if rc != 0:
            raise RuntimeError("SUEWS fatal error %d: %s" % (rc, msg.value.decode()))
        blocks["_warn"] = np.array(warn[:])
        self.n_iter_last = nit.value      # Tsurf-loop trips of the last step
        return blocks


def sun_position(year, dectime, utc, lat, lng, alt):
    az, zen = C.c_double(), C.c_double()
    lib().so_NARP_cal_SunPosition(year, dectime, utc, lat, lng, alt, C.byref(az), C.byref(zen))
    return az.value, zen.value
