"""SuewsBatch: many grid cells resident on one B200, advanced many timesteps per launch.

Host-side mirror of the batched C ABI (include/suews_b200.h).  Parameters are given with the
reference's own argument names (SUEWS_cal_Main, suews_ctrl_driver.f95:45-91): class-level values
per class, site values and state per cell.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, dp, ip

NCOL = {"SUEWS": 87, "snow": 103, "RSL": 138, "DailyState": 50, "SOLWEIG": 31, "ESTM": 32, "BL": 22}
OUT_SUM82 = 1
OUT_FULL = 2
OUT_SUM7 = 4   # only the seven parity fluxes (output policy P0)
OUT_FLUX_ONLY = 16   # (first release; OUT_SUM7 alone now skips the diagnostics its seven sums cannot observe)
OUT_ALL_STAGES = 32  # with OUT_SUM7: run every stage of SUEWS_cal_Main anyway (T2/q2/U10/RH2 included)
OUT_AGG = 8    # SUEWS block aggregated to the output frequency inside the kernel (output policy P1)

KIND_CLASS_INT, KIND_CLASS_DBL, KIND_CELL, KIND_STATE, KIND_ESTM_STATE = 0, 1, 2, 3, 4


def field_info(name):
    kind = C.c_int()
    n = _lib.lib().suews_b200_field_info(name.encode(), C.byref(kind))
    if n < 0:
        return None
    return kind.value, n


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


class SuewsBatch:
    def __init__(self, ncell, nclass=1, device=0):
        self.L = _lib.lib()
        self.ncell, self.nclass, self.device = int(ncell), int(nclass), int(device)
        self.h = C.c_void_p()
        check(self.L.suews_b200_create(self.ncell, self.nclass, self.device, C.byref(self.h)))
        self.nsteps = 0
        self._class_cache = {}

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.suews_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters
    def _class_value(self, name):
        return self._class_cache[name.lower()]

    _PREFIX = "suews_b200_"

    def _fn(self, name):
        # (the calls SuewsMultiBatch shares with this class resolve to its suews_b200_multi_* twins there)
        return getattr(self.L, self._PREFIX + name)

    def set_class_param(self, cls, name, vals):
        a = _arr(np.atleast_1d(vals)).ravel()
        if int(cls) == 0 and a.size == 1:
            self._class_cache[name.lower()] = float(a[0])
        check(self.L.suews_b200_set_class_param(self.h, int(cls), name.encode(), a.ctypes.data_as(dp), a.size))

    def set_cell_param(self, name, vals):
        """vals: scalar / [n] (broadcast to all cells) or [ncell, n]."""
        kind, n = field_info(name)
        a = _arr(vals)
        if a.ndim <= 1 and a.size == n:
            a = np.broadcast_to(a.reshape(1, n), (self.ncell, n))
        a = _arr(a).reshape(self.ncell, n)
        check(self.L.suews_b200_set_cell_param(self.h, name.encode(), a.ctypes.data_as(dp), n))

    def set_state(self, name, vals):
        kind, n = field_info(name)
        a = _arr(vals)
        if a.ndim <= 1 and a.size == n:
            a = np.broadcast_to(a.reshape(1, n), (self.ncell, n))
        a = _arr(a).reshape(self.ncell, n)
        check(self.L.suews_b200_set_state(self.h, name.encode(), a.ctypes.data_as(dp), n))

    def get_state(self, name):
        kind, n = field_info(name)
        a = np.zeros((self.ncell, n))
        check(self.L.suews_b200_get_state(self.h, name.encode(), a.ctypes.data_as(dp), n))
        return a

    def set_cell_class(self, ids):
        a = np.ascontiguousarray(np.asarray(ids, dtype=np.int32))
        assert a.size == self.ncell
        check(self.L.suews_b200_set_cell_class(self.h, a.ctypes.data_as(ip)))

    def set_clock(self, dt_since_start, tstep_prev):
        check(self.L.suews_b200_set_clock(self.h, int(dt_since_start), int(tstep_prev)))

    def get_clock(self):
        a, b = C.c_int(), C.c_int()
        check(self.L.suews_b200_get_clock(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def load_reference_args(self, values, cls=0, cells=True):
        """Fill class `cls` (and, if cells, every cell's site values + state) from a dict keyed by the
        reference's argument names (e.g. a parsed *_state_init.txt dump).  Returns names not found."""
        low = {k.lower(): v for k, v in values.items()}
        missing = []
        for name, kind, n in all_fields():
            key = name.lower()
            if key not in low:
                # optional: synthetic forcing offsets, hidden SAVE state, and MODULE ESTM_data (not SUEWS_cal_Main arguments)
                if name not in ("dTemp_C", "dRH", "alb1_hidden", "lv_switch", "rsl_cache") and name not in estm_names():
                    missing.append(name)
                continue
            v = np.atleast_1d(np.asarray(low[key], dtype=np.float64)).ravel()
            assert v.size == n, (name, v.size, n)
            if kind in (KIND_CLASS_INT, KIND_CLASS_DBL):
                self.set_class_param(cls, name, v)
            elif cells and kind == KIND_CELL:
                self.set_cell_param(name, v)
            elif cells and kind in (KIND_STATE, KIND_ESTM_STATE):
                self.set_state(name, v)
        if "dt_since_start" in low:
            self.set_clock(int(low["dt_since_start"][0]), int(low.get("tstep_prev", low.get("tstep"))[0]))
        return missing

    # ---- run
    def upload(self):
        check(self.L.suews_b200_upload(self.h))

    def resort(self):
        """sort the thread slots again on the current state (snow activity, canopy height); the run carries on"""
        check(self.L.suews_b200_resort(self.h))

    def set_forcing(self, met):
        """met: [nsteps, 24] MetForcingBlock (driver.f95:3177-3201), pressure in hPa."""
        m = np.asfortranarray(np.asarray(met, dtype=np.float64))
        assert m.shape[1] == 24
        check(self.L.suews_b200_set_forcing(self.h, m.ctypes.data_as(dp), m.shape[0]))
        self.nsteps = m.shape[0]

    def set_forcing_orig(self, met_orig, resolution_in=3600, disagg_method=1, rain_disagg_method=100, raw_units=False,
                         smd_method=0, kdown_zen=0, site=None, **random_rain):
        """Forcing at file resolution, disaggregated to the model timestep on the device (DisaggregateMet,
        suews_ctrl_input.f95:1380-1640).  met_orig: [nrows, 24]; raw_units: the file's own units (pressure in kPa),
        i.e. a block that has not been through forcing.read_met.  kdown_zen = 1 (RunControl's KdownZen, :1568-1600) needs
        site = (lat, lng, timezone, alt): the reference re-weights K-down per grid with the grid's own sun, a batch
        shares one forcing block, so the site is the forcing's.  Returns the number of model timesteps."""
        if int(kdown_zen) not in (0, 1):
            raise ValueError("KdownZen is 0 or 1")
        if int(rain_disagg_method) in (101, 102):
            # random sub-intervals (RandomSamples, suews_ctrl_input.f95:2147): drawn on the host -- rng=, rain_among_n= or
            # mult_rain_among_n= / mult_rain_among_n_upper_i= as forcing.disaggregate_met takes them -- and the block staged whole
            from . import forcing
            if raw_units:
                raise ValueError("RainDisaggMethod 101/102: pass the block through forcing.read_met first")
            if int(kdown_zen) == 1:
                raise NotImplementedError("RainDisaggMethod 101/102 together with KdownZen = 1: disaggregate with forcing.disaggregate_met and set_forcing")
            block = forcing.disaggregate_met(np.asarray(met_orig, dtype=np.float64), tstep=self.tstep(), resolution_in=int(resolution_in),
                                             disagg_method=int(disagg_method), rain_disagg_method=int(rain_disagg_method), **random_rain)
            self.set_forcing(block)
            return self.nsteps
        if int(kdown_zen) == 1 and site is None:
            raise ValueError("KdownZen = 1 needs site=(lat, lng, timezone, alt)")
        m = np.asfortranarray(np.asarray(met_orig, dtype=np.float64))
        assert m.shape[1] == 24
        check(self._fn("set_forcing_orig")(self.h, m.ctypes.data_as(dp), m.shape[0], int(resolution_in), int(disagg_method),
                                                 int(rain_disagg_method), int(bool(raw_units)), int(smd_method)))
        self.nsteps = m.shape[0] * (int(resolution_in) // self.tstep())
        if int(kdown_zen) == 1:
            check(self._fn("apply_kdown_zen")(self.h, *[float(v) for v in site]))
        return self.nsteps

    def tstep(self):
        # class 0 carries the timestep all classes share (checked at upload)
        return int(self._class_value("tstep"))

    def forcing_block(self):
        """The MetForcingBlock [nsteps, 24] the device built in set_forcing_orig."""
        b = np.zeros((self.nsteps, 24), order="F")
        check(self.L.suews_b200_get_forcing_block(self.h, b.ctypes.data_as(dp), self.nsteps))
        return b

    def set_estm_forcing(self, ts5):
        """ts5: [nsteps, 13] ESTM surface-temperature forcing (iy id it imin Tiair Tsurf Troof Troad Twall Twall_n
        Twall_e Twall_s Twall_w; the reference's ESTMForcingData, suews_ctrl_translate.f95:1264), same rows as the
        forcing block set just before.  Only read when StorageHeatMethod is 4/14."""
        t = np.asfortranarray(np.asarray(ts5, dtype=np.float64))
        assert t.shape == (self.nsteps, 13)
        check(self._fn("set_estm_forcing")(self.h, t.ctypes.data_as(dp), t.shape[0]))

    def set_cbl_initial(self, ini):
        """BLUEWS: the rows of CBL_initial_data.txt, [nline, 8] (DOY zi0 gamt gamq theta+ q+ theta q), shared by the batch
        (IniCBLdata, suews_phys_bluews.f95:319-341)."""
        t = np.ascontiguousarray(np.asarray(ini, dtype=np.float64))
        assert t.ndim == 2 and t.shape[1] == 8
        check(self._fn("set_cbl_initial")(self.h, t.ctypes.data_as(dp), t.shape[0]))

    def set_anohm_forcing(self, met):
        """AnOHM's MetForcingData_grid ([nrows, 24], the MetForcingBlock columns) when it is not the staged forcing block
        itself (suews_ctrl_translate.f95:1216); holds until the next set_forcing."""
        m = np.asfortranarray(np.asarray(met, dtype=np.float64))
        assert m.ndim == 2 and m.shape[1] == 24
        check(self._fn("set_anohm_forcing")(self.h, m.ctypes.data_as(dp), m.shape[0]))

    def set_estm_forcing_orig(self, ts_orig, resolution_in=3600, disagg_method=1):
        """ESTM surface temperatures at file resolution [nrows, 13], disaggregated on the device (DisaggregateESTM,
        suews_ctrl_input.f95:1650-1849).  Returns the [nsteps, 13] block the device built."""
        t = np.asfortranarray(np.asarray(ts_orig, dtype=np.float64))
        assert t.shape[1] == 13
        check(self._fn("set_estm_forcing_orig")(self.h, t.ctypes.data_as(dp), t.shape[0], int(resolution_in), int(disagg_method)))
        b = np.zeros((self.nsteps, 13), order="F")
        check(self.L.suews_b200_get_estm_forcing_block(self.h, b.ctypes.data_as(dp), self.nsteps))
        return b

    def advance(self, step0, nsteps, out_flags=OUT_SUM82, sync=True):
        check(self.L.suews_b200_advance(self.h, int(step0), int(nsteps), int(out_flags)))
        if sync:
            self.sync()

    def sync(self):
        check(self.L.suews_b200_sync(self.h))

    def set_output_freq(self, seconds):
        check(self.L.suews_b200_set_output_freq(self.h, int(seconds)))

    def agg_block(self, cell, nrows):
        """One cell's SUEWS block aggregated as SUEWS_Output_Agg does (suews_ctrl_output.f95:912-957), [nrows, 87]."""
        b = np.zeros((nrows, 87), order="F")
        check(self.L.suews_b200_get_agg_block(self.h, int(cell), b.ctypes.data_as(dp), int(nrows)))
        return b

    def agg_daily_block(self, cell):
        """One cell's DailyState rows (one per day that ended inside the last OUT_AGG launch), [ndays, 50]."""
        n = C.c_int()
        check(self.L.suews_b200_agg_daily_rows(self.h, C.byref(n)))
        b = np.zeros((n.value, 50), order="F")
        check(self.L.suews_b200_get_agg_daily_block(self.h, int(cell), b.ctypes.data_as(dp), n.value))
        return b

    def agg_column(self, row, col):
        a = np.zeros(self.ncell)
        check(self.L.suews_b200_get_agg_column(self.h, int(row), int(col), a.ctypes.data_as(dp)))
        return a

    def last_ms(self):
        ms = C.c_float()
        check(self.L.suews_b200_last_advance_ms(self.h, C.byref(ms)))
        return ms.value

    def sums(self):
        a = np.zeros((self.ncell, 82))
        check(self.L.suews_b200_get_sums(self.h, a.ctypes.data_as(dp)))
        return a

    def sums_cols(self, cols):
        c = np.ascontiguousarray(np.asarray(cols, dtype=np.int32))
        a = np.zeros((self.ncell, c.size))
        check(self.L.suews_b200_get_sums_cols(self.h, c.ctypes.data_as(ip), c.size, a.ctypes.data_as(dp)))
        return a

    def device_sync(self):
        check(self.L.suews_b200_device_sync(self.h))

    def reset_sums(self):
        check(self.L.suews_b200_reset_sums(self.h))

    def block(self, group, cell, nsteps):
        b = np.zeros((nsteps, NCOL[group]), order="F")
        check(self._fn("get_block")(self.h, group.encode(), int(cell), b.ctypes.data_as(dp), nsteps, NCOL[group]))
        return b

    def diag(self):
        fatal = np.zeros(self.ncell, dtype=np.int32)
        warn = np.zeros((self.ncell, 2), dtype=np.uint64)
        iters = np.zeros(self.ncell, dtype=np.int64)
        check(self.L.suews_b200_get_diag(self.h, fatal.ctypes.data_as(ip), warn.ctypes.data_as(C.POINTER(C.c_ulonglong)),
                                         iters.ctypes.data_as(C.POINTER(C.c_longlong))))
        return fatal, warn, iters


class SuewsMultiBatch(SuewsBatch):
    """The cells of one run spread over several GPUs of one box behind one handle (suews_b200_multi_*,
    include/suews_b200.h): contiguous blocks of the cell index, one per device, one host thread, no exchange while the
    steps advance; gather_sums() is the final gather (device -> first device over NVLink -> host).  Setters take the arrays
    of the whole run, like SuewsBatch."""

    def __init__(self, ncell, nclass=1, devices=None, ndev=None):
        self.L = _lib.lib()
        if devices is None:
            devices = list(range(int(ndev) if ndev else max(1, self.L.suews_b200_device_count())))
        self.devices = [int(d) for d in devices]
        self.ncell, self.nclass = int(ncell), int(nclass)
        self.h = C.c_void_p()
        dv = (C.c_int * len(self.devices))(*self.devices)
        check(self.L.suews_b200_multi_create(self.ncell, self.nclass, len(self.devices), dv, C.byref(self.h)))
        self.nsteps = 0
        self._class_cache = {}

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.suews_b200_multi_destroy(self.h)
            self.h = C.c_void_p()

    def parts(self):
        """[(device, first cell, cell count)] of the partition"""
        out = []
        for d in range(len(self.devices)):
            c0, n = C.c_int(), C.c_int()
            check(self.L.suews_b200_multi_part(self.h, d, None, C.byref(c0), C.byref(n)))
            out.append((self.devices[d], c0.value, n.value))
        return out

    def set_class_param(self, cls, name, vals):
        a = _arr(np.atleast_1d(vals)).ravel()
        if int(cls) == 0 and a.size == 1:
            self._class_cache[name.lower()] = float(a[0])
        check(self.L.suews_b200_multi_set_class_param(self.h, int(cls), name.encode(), a.ctypes.data_as(dp), a.size))

    def _per_cell(self, name, vals):
        kind, n = field_info(name)
        a = _arr(vals)
        if a.ndim <= 1 and a.size == n:
            a = np.broadcast_to(a.reshape(1, n), (self.ncell, n))
        return _arr(a).reshape(self.ncell, n), n

    def set_cell_param(self, name, vals):
        a, n = self._per_cell(name, vals)
        check(self.L.suews_b200_multi_set_cell_param(self.h, name.encode(), a.ctypes.data_as(dp), n))

    def set_state(self, name, vals):
        a, n = self._per_cell(name, vals)
        check(self.L.suews_b200_multi_set_state(self.h, name.encode(), a.ctypes.data_as(dp), n))

    def get_state(self, name):
        kind, n = field_info(name)
        a = np.zeros((self.ncell, n))
        check(self.L.suews_b200_multi_get_state(self.h, name.encode(), a.ctypes.data_as(dp), n))
        return a

    def set_cell_class(self, ids):
        a = np.ascontiguousarray(np.asarray(ids, dtype=np.int32))
        assert a.size == self.ncell
        check(self.L.suews_b200_multi_set_cell_class(self.h, a.ctypes.data_as(ip)))

    def set_clock(self, dt_since_start, tstep_prev):
        check(self.L.suews_b200_multi_set_clock(self.h, int(dt_since_start), int(tstep_prev)))

    def upload(self):
        check(self.L.suews_b200_multi_upload(self.h))

    def set_forcing(self, met):
        m = np.asfortranarray(np.asarray(met, dtype=np.float64))
        assert m.shape[1] == 24
        check(self.L.suews_b200_multi_set_forcing(self.h, m.ctypes.data_as(dp), m.shape[0]))
        self.nsteps = m.shape[0]

    # the blocks a batch shares (set_forcing_orig, set_estm_forcing, set_cbl_initial, set_anohm_forcing) and one cell's
    # per-step output (block) are SuewsBatch's methods on the multi-device entry points
    _PREFIX = "suews_b200_multi_"

    def _part_handle(self, d=0):
        h = C.c_void_p()
        check(self.L.suews_b200_multi_part(self.h, int(d), C.byref(h), None, None))
        return h

    def forcing_block(self):
        """The MetForcingBlock [nsteps, 24] the devices built in set_forcing_orig (every device holds the same rows)."""
        b = np.zeros((self.nsteps, 24), order="F")
        check(self.L.suews_b200_get_forcing_block(self._part_handle(0), b.ctypes.data_as(dp), self.nsteps))
        return b

    def set_estm_forcing_orig(self, ts_orig, resolution_in=3600, disagg_method=1):
        t = np.asfortranarray(np.asarray(ts_orig, dtype=np.float64))
        assert t.shape[1] == 13
        check(self.L.suews_b200_multi_set_estm_forcing_orig(self.h, t.ctypes.data_as(dp), t.shape[0], int(resolution_in), int(disagg_method)))
        b = np.zeros((self.nsteps, 13), order="F")
        check(self.L.suews_b200_get_estm_forcing_block(self._part_handle(0), b.ctypes.data_as(dp), self.nsteps))
        return b

    def advance(self, step0, nsteps, out_flags=OUT_SUM7, sync=True):
        check(self.L.suews_b200_multi_advance(self.h, int(step0), int(nsteps), int(out_flags)))
        if sync:
            self.sync()

    def sync(self):
        check(self.L.suews_b200_multi_sync(self.h))

    def last_ms(self):
        ms = C.c_float()
        check(self.L.suews_b200_multi_last_advance_ms(self.h, C.byref(ms)))
        return ms.value

    def gather_sums(self, cols):
        """the final gather: [ncell, len(cols)] on the host; returns (array, milliseconds the gather took)"""
        c = np.ascontiguousarray(np.asarray(cols, dtype=np.int32))
        a = np.zeros((self.ncell, c.size))
        ms = C.c_float()
        check(self.L.suews_b200_multi_gather_sums(self.h, c.ctypes.data_as(ip), c.size, a.ctypes.data_as(dp), C.byref(ms)))
        return a, ms.value

    def sums_cols(self, cols):
        return self.gather_sums(cols)[0]

    def diag(self):
        fatal = np.zeros(self.ncell, dtype=np.int32)
        warn = np.zeros((self.ncell, 2), dtype=np.uint64)
        iters = np.zeros(self.ncell, dtype=np.int64)
        check(self.L.suews_b200_multi_get_diag(self.h, fatal.ctypes.data_as(ip), warn.ctypes.data_as(C.POINTER(C.c_ulonglong)),
                                               iters.ctypes.data_as(C.POINTER(C.c_longlong))))
        return fatal, warn, iters

    def device_sync(self):
        self.sync()


_FIELDS = None


def all_fields():
    """[(name, kind, n)] for every parameter/state field the library knows (from suews_fields.h)."""
    global _FIELDS
    if _FIELDS is None:
        import os
        import re
        hdr = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "suews_fields.h")).read()
        out = []
        # (cell parameters and state are listed as COLD fields then HOT fields: the order of the handle's SoA arrays)
        for macro, kind in (("SUEWS_CLASS_INT_FIELDS", 0), ("SUEWS_CLASS_DBL_FIELDS", 1), ("SUEWS_CELL_COLD_FIELDS", 2),
                            ("SUEWS_CELL_HOT_FIELDS", 2), ("SUEWS_STATE_COLD_FIELDS", 3), ("SUEWS_STATE_HOT_FIELDS", 3),
                            ("SUEWS_ESTM_STATE_FIELDS", 4)):
            m = re.search(r"#define %s\(X\)((?:[^\n]*\\\n)*[^\n]*)\n" % macro, hdr)   # the macro body with its continuation lines
            for nm, n in re.findall(r"X\((\w+),\s*(\d+)\)", m.group(1)):
                out.append((nm, kind, int(n)))
        _FIELDS = out
    return _FIELDS


_ESTM_NAMES = None


def estm_names():
    """Names of the fields that mirror MODULE ESTM_data (suews_phys_estm.f95:1-176): ESTM parameters and state."""
    global _ESTM_NAMES
    if _ESTM_NAMES is None:
        names = {nm for nm, kind, _n in all_fields() if kind == KIND_ESTM_STATE}
        names |= {"evolveTibld", "TsurfChoice", "ibldCHmod", "Nibld", "Nwall", "Nroof", "Nground", "LBC_soil", "THEAT_ON",
                  "THEAT_OFF", "THEAT_fix", "nroom", "alb_ibld", "em_ibld", "CH_iroof", "CH_iwall", "CH_ibld", "Tfloor",
                  "AreaWall"}
        names |= {"ivf_" + a + b for a in "iwrf" for b in "iwrf"}
        names |= {p + e for p in "zkr" for e in ("ibld", "wall", "roof", "ground")}
        # CBLInput.nml / RunControl's CBLuse (BLUEWS runs beside SUEWS_cal_Main, not inside its argument list)
        names |= {"CBLuse", "EntrainmentType", "QH_choice", "isubs", "CO2_included", "InitialData_use", "sondeflag", "wsb"}
        _ESTM_NAMES = names
    return _ESTM_NAMES
