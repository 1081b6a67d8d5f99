"""The reference-shaped entry points (SUEWS_cal_multitsteps / SUEWS_cal_Main argument lists) on the GPU
vs the CPU oracle.  These read like a supy_driver test: keyword arguments in, output blocks out,
INTENT(INOUT) state arrays updated in place."""
import ctypes as C

import numpy as np
import pytest

import fixtures as fx
import oracle_lib as ol
from test_gpu_parity import TOL_DAY1, rel_err

pytestmark = pytest.mark.gpu

GROUP_OF_BLOCK = {0: "SUEWS", 1: "snow", 3: "RSL", 4: "SOLWEIG", 5: "DailyState"}


def _kwargs(d, which):
    """numpy arguments (int32 for INTEGER dummies, float64 otherwise) from a state_init dump"""
    from suews_b200.supy_driver import arg_table
    low = {k.lower(): v for k, v in d.items()}
    kw = {}
    for nm, is_int, intent, n, per_row in arg_table(which):
        key = nm.lower()
        if intent == 2 or key not in low or per_row:
            continue
        kw[key] = np.array(low[key], dtype=np.int32 if is_int else np.float64).ravel()
    return kw


def _check_blocks(outs, ref, nrows, tol=TOL_DAY1):
    for k, group in GROUP_OF_BLOCK.items():
        out, r = outs[k], ref[group][:nrows]
        assert out.shape == r.shape, group
        assert np.array_equal(out[:, :5], r[:, :5]), group
        assert np.array_equal(out == -999.0, r == -999.0), group
        e = rel_err(out[:, 5:], r[:, 5:], group)
        assert e.max() <= tol, (group, float(e.max()))


def test_cal_multitsteps_matches_oracle():
    from suews_b200.supy_driver import suews_driver as sd
    sd.reset_hidden_state()
    d = fx.state_init(2005)
    met = fx.forcing(2005)[:288]
    cell = ol.Cell(d)
    ref = cell.run(met, want=("SUEWS", "snow", "RSL", "DailyState", "SOLWEIG"))
    kw = _kwargs(d, 1)
    dt0 = int(kw["dt_since_start"][0])
    outs = sd.suews_cal_multitsteps(metforcingblock=met, len_sim=288, **kw)
    assert len(outs) == 6
    assert [o.shape[1] for o in outs] == [87, 103, 32, 138, 31, 50]
    _check_blocks(outs, ref, 288)
    # ESTM block: not written unless StorageHeatMethod = 4 (driver.f95:2503): date columns + -999
    assert np.array_equal(outs[2][:, :5], ref["SUEWS"][:, :5]) and np.all(outs[2][:, 5:] == -999.0)
    # INTENT(INOUT) arguments were advanced in place
    assert int(kw["dt_since_start"][0]) == dt0 + 288 * 300
    st = cell.state_dict()
    for nm, v in st.items():
        if nm.lower() in kw and nm not in ("dt_since_start", "tstep_prev"):
            assert np.allclose(kw[nm.lower()], v, rtol=1e-9, atol=1e-12), nm

    # a second call continues from the updated arguments (supy's chunked runs)
    met2 = fx.forcing(2005)[288:288 + 36]
    ref2 = cell.run(met2, want=("SUEWS", "snow", "RSL", "DailyState", "SOLWEIG"))
    outs2 = sd.suews_cal_multitsteps(metforcingblock=met2, **kw)
    _check_blocks(outs2, ref2, 36, tol=1e-9)


def test_cal_main_step_by_step():
    from suews_b200.supy_driver import suews_driver as sd
    sd.reset_hidden_state()
    d = fx.state_init(2005)
    nstep = 14
    met = fx.forcing(2005)[:nstep]
    ref = ol.Cell(d).run(met, want=("SUEWS", "snow", "RSL", "DailyState", "SOLWEIG"))
    kw = _kwargs(d, 0)
    cols = {"iy": 0, "id": 1, "it": 2, "imin": 3, "qn1_obs": 4, "qs_obs": 7, "qf_obs": 8, "avu1": 9, "avrh": 10, "temp_c": 11,
            "press_hpa": 12, "precip": 13, "avkdn": 14, "snowfrac_obs": 15, "ldown_obs": 16, "fcld_obs": 17, "wu_m3": 18,
            "xsmd": 19, "lai_obs": 20}
    tstep = int(kw["tstep"][0])
    for i in range(nstep):
        for k, c in cols.items():
            kw[k] = np.array([met[i, c]], dtype=kw[k].dtype)
        kw["isec"] = np.array([0], dtype=np.int32)
        line = sd.suews_cal_main(**kw)
        assert [x.size for x in line] == [5, 82, 98, 27, 133, 26, 45]
        assert np.array_equal(line[0], ref["SUEWS"][i, :5])
        for k, group in ((1, "SUEWS"), (4, "RSL"), (5, "SOLWEIG"), (6, "DailyState")):
            r = ref[group][i:i + 1, 5:]
            e = rel_err(line[k][None, :], r, group)
            # single rows have no column history for the scale: use the day-1 tolerance on max(|ref|, 1)
            e = np.minimum(e, np.abs(line[k][None, :] - r) / np.maximum(np.abs(r), 1.0))
            assert np.nanmax(e) <= TOL_DAY1, (i, group, float(np.nanmax(e)))
        # the caller advances the clock between SUEWS_cal_Main calls (driver.f95:3482, SuMin :590)
        kw["dt_since_start"] = kw["dt_since_start"] + tstep


def test_gfortran_positional_symbol_equals_pointer_table_call():
    """suews_cal_multitsteps_ (200 by-reference parameters) and suews_cal_multitsteps_c are the same call"""
    from suews_b200._lib import last_error, lib
    from suews_b200.supy_driver import arg_table
    from suews_b200.supy_driver import suews_driver as sd
    d = fx.state_init(2005)
    met = np.asfortranarray(fx.forcing(2005)[:24])
    sd.reset_hidden_state()
    kw = _kwargs(d, 1)
    ref = sd.suews_cal_multitsteps(metforcingblock=met, **kw)

    sd.reset_hidden_state()
    kw = _kwargs(d, 1)
    kw["metforcingblock"] = met
    kw["len_sim"] = np.array([24], dtype=np.int32)
    keep, ptrs, outs = [], [], []
    for nm, is_int, intent, n, per_row in arg_table(1):
        if intent == 2:
            a = np.full((24, n), -999.0, order="F")
            outs.append(a)
        elif n == 0:
            ptrs.append(C.c_void_p(None))
            continue
        else:
            a = kw[nm.lower()]
        keep.append(a)
        ptrs.append(C.c_void_p(a.ctypes.data))
    f = lib().suews_cal_multitsteps_
    f.restype = None
    f.argtypes = [C.c_void_p] * len(ptrs)
    assert len(ptrs) == 200
    f(*ptrs)
    # (a non-zero return code inside the shim is the reference's STOP: the process would have ended, suews_refapi.cpp fortran_stop)
    for a, b in zip(outs, ref):
        assert np.array_equal(a, b)


def test_fatal_condition_raises_like_the_reference():
    from suews_b200 import SuewsError
    from suews_b200.supy_driver import suews_driver as sd
    d = fx.state_init(2005)
    kw = _kwargs(d, 1)
    kw["storageheatmethod"] = np.array([2], dtype=np.int32)      # no such storage-heat scheme
    with pytest.raises(SuewsError) as ei:
        sd.suews_cal_multitsteps(metforcingblock=fx.forcing(2005)[:3], **kw)
    assert ei.value.code == 1003
    kw["storageheatmethod"] = np.array([4], dtype=np.int32)      # ESTM without MODULE ESTM_data / Ts5mindata_ir
    with pytest.raises(SuewsError) as ei:
        sd.suews_cal_multitsteps(metforcingblock=fx.forcing(2005)[:3], **kw)
    assert ei.value.code == 1005


def test_cal_multitsteps_with_estm():
    """StorageHeatMethod 4 through the reference-shaped call: ESTM's parameters and element temperatures are the
    variables of MODULE ESTM_data (set by the host, as ESTM_initials / SUEWS_Translate / ESTM_translate do in the
    reference), Ts5mindata_ir is the argument -- one row for the whole block (driver.f95:3044)."""
    from suews_b200 import estm
    from suews_b200.supy_driver import suews_driver as sd
    sd.reset_hidden_state()
    d = dict(fx.state_init(2005))
    d["storageheatmethod"] = [4]
    ep = fx.estm_example(d["surfacearea"][0], evolveTibld=1, ibldCHmod=1)
    n = 48
    met = fx.forcing(2005)[:n]
    row = fx.estm_ts_forcing(met)[0]
    ts = np.tile(row, (n, 1))
    dd = dict(d)
    dd.update({k.lower(): v for k, v in ep.items()})
    cell = ol.Cell(dd)
    cell.estm_start(row)
    ref = cell.run(met, want=("SUEWS", "snow", "RSL", "DailyState", "SOLWEIG", "ESTM"), ts5=ts)
    sd.estm_data_set(**ep)
    sd.estm_data_set(**estm.initial_state(ep, d["sfr"], d["emis"], row))
    kw = _kwargs(d, 1)
    outs = sd.suews_cal_multitsteps(metforcingblock=met, len_sim=n, ts5mindata_ir=row.copy(), **kw)
    _check_blocks(outs, ref, n)
    oe, re_ = outs[2], ref["ESTM"]
    assert np.array_equal(oe[:, :5], re_[:, :5]) and np.array_equal(oe == -999.0, re_ == -999.0)
    assert np.abs(oe[:, 5:] - re_[:, 5:]).max() <= 1e-9
    # the module's element temperatures were advanced and can be read back
    np.testing.assert_allclose(sd.estm_data_get("Twall", 5), cell.get_state("Twall"), rtol=1e-11)
    np.testing.assert_allclose(sd.estm_data_get("Tievolve", 1), cell.get_state("Tievolve"), rtol=1e-11)
    sd.reset_hidden_state()


def test_positional_call_in_f2py_order_equals_keyword_call():
    """f2py lets the reference's suews_cal_multitsteps be called positionally, the Fortran dummies in declaration order
    without the INTENT(OUT) ones and the optional len_sim behind them (driver.f95:2814-2860); supy_driver here too."""
    from suews_b200.supy_driver import arg_table, suews_driver as sd
    d = fx.state_init(2005)
    met = np.asfortranarray(fx.forcing(2005)[:24])
    outs = []
    for positional in (False, True):
        sd.reset_hidden_state()
        kw = _kwargs(d, 1)
        kw["metforcingblock"] = met
        if positional:
            names = [nm.lower() for nm, _i, intent, _n, _r in arg_table(1) if intent != 2 and nm.lower() != "len_sim"]
            args = []
            for nm in names:
                if nm not in kw:
                    break               # trailing assumed-shape arguments the path never reads
                args.append(kw.pop(nm))
            outs.append(sd.suews_cal_multitsteps(*args, **kw))
        else:
            outs.append(sd.suews_cal_multitsteps(**kw))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_dropin_package_serves_the_reference_import():
    """`import supy_driver as sd; sd.suews_driver.output_name_n(3)` -- the reference's own package test
    (supy-driver/supy_driver/test/test-supy_driver.py:12) -- against the drop-in package."""
    import importlib
    import os
    import sys
    here = os.path.join(os.path.dirname(fx.GOLDEN), "..", "suews_b200", "dropin")
    sys.path.insert(0, os.path.abspath(here))
    try:
        sd = importlib.import_module("supy_driver")
        s = sd.suews_driver.output_name_n(3)[0]
        assert isinstance(s, bytes) and len(s) == 15
        assert sd.__version__.startswith("2020b")
        for nm in ("narp_module", "atmmoiststab_module", "resist_module", "evap_module", "snow_module", "dailystate_module",
                   "lumps_module", "anemsn_module", "rsl_module", "co2_module", "waterdist_module", "meteo"):
            assert hasattr(sd, nm)
        az, zen = sd.narp_module.narp_cal_sunposition(2005.0, 1.5, 2.0, 57.77, 11.87, 10.0)
        assert 0 <= az <= 360 and 0 <= zen <= 180
    finally:
        sys.path.pop(0)
        sys.modules.pop("supy_driver", None)


def _anohm_dump():
    from suews_b200 import synth
    d = fx.state_init(2005)
    d.update({k: list(v) for k, v in synth.ANOHM_EXAMPLE.items()})
    d["storageheatmethod"] = [3]
    return d


def test_anohm_through_the_reference_shaped_calls():
    """StorageHeatMethod 3 through suews_cal_multitsteps (MetForcingData_grid = its own MetForcingBlock; a chunked second
    call continues with the process-wide SAVE'd coefficients) and step by step through suews_cal_main with the grid's
    block handed over by anohm_set_met_grid; without it the single-step call refuses (1003)."""
    from suews_b200 import SuewsError
    from suews_b200.supy_driver import suews_driver as sd
    sd.reset_hidden_state()
    sd.anohm_set_met_grid(None)
    d = _anohm_dump()
    n = 288 + 30
    met = fx.forcing(2005)[:n]
    cell = ol.Cell(d)
    ref = cell.run(met, want=("SUEWS", "snow", "RSL", "DailyState", "SOLWEIG"))
    kw = _kwargs(d, 1)
    outs = sd.suews_cal_multitsteps(metforcingblock=met, len_sim=n, **kw)
    from test_gpu_parity import TOL_ANOHM      # the fit amplifies libm's last-ulp differences: see there
    _check_blocks(outs, ref, n, tol=TOL_ANOHM)
    assert np.allclose(sd.estm_data_get("anohm_coeff_grid_day", 21), cell.get_state("anohm_coeff_grid_day"), rtol=TOL_ANOHM, atol=0)
    assert sd.estm_data_get("anohm_coef_id_save", 1)[0] == cell.get_state("anohm_coef_id_save")[0]

    # single steps: the first 20 steps of the same run, the whole block as MetForcingData_grid
    sd.reset_hidden_state()
    kw = _kwargs(d, 0)
    cols = {"iy": 0, "id": 1, "it": 2, "imin": 3, "qn1_obs": 4, "qs_obs": 7, "qf_obs": 8, "avu1": 9, "avrh": 10, "temp_c": 11,
            "press_hpa": 12, "precip": 13, "avkdn": 14, "snowfrac_obs": 15, "ldown_obs": 16, "fcld_obs": 17, "wu_m3": 18,
            "xsmd": 19, "lai_obs": 20}
    tstep = int(kw["tstep"][0])

    def step(i):
        for k, c in cols.items():
            kw[k] = np.array([met[i, c]], dtype=kw[k].dtype)
        kw["isec"] = np.array([0], dtype=np.int32)
        return sd.suews_cal_main(**kw)
    with pytest.raises(SuewsError) as e:
        step(0)
    assert e.value.code == 1003
    sd.anohm_set_met_grid(met)
    kw = _kwargs(d, 0)
    for i in range(20):
        line = step(i)
        r = ref["SUEWS"][i:i + 1, 5:]
        err = rel_err(line[1][None, :], r, "SUEWS")
        assert err.max() <= TOL_ANOHM, (i, float(err.max()))
        kw["dt_since_start"] = kw["dt_since_start"] + tstep
    sd.anohm_set_met_grid(None)
    sd.reset_hidden_state()
