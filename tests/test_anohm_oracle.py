"""AnOHM (StorageHeatMethod 3): what pins the oracle's restatement, and the device code against the oracle, without a GPU.

The reference repository holds no output of any AnOHM run, so oracle/so_anohm.c is "parity unpinned" against the
reference itself (its header says so).  What can be pinned here:
  * the two MINPACK drivers it stands on (oracle/so_minpack.c, restated from the reference's suews_util_minpack.f95)
    against an independent implementation of the same published algorithm: scipy.optimize.leastsq / fsolve, which wrap
    MINPACK's lmdif / hybrd -- iterates identical to the last bit, same number of function evaluations;
  * the properties the scheme must have (a fitted sinusoid is recovered, the Bowen-ratio solve lands on the closed form
    its residual reduces to, the coefficients are formed once a day and reloaded, QS is OHM's formula with dqndt_prev);
  * the device-side code (suews_b200/csrc/dev_anohm.cuh: lmdif for n = 3, hybrd reduced to scalars for n = 1, the day's
    rows read through the library's hour table), compiled for the host by tests/host_anohm.cpp, bit for bit against
    the oracle's general-n restatement over a year of days.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fixtures as fx  # noqa: E402
import oracle_lib as ol  # noqa: E402

dp = C.POINTER(C.c_double)
PI4 = float(np.float32(3.14159274))   # PI = ATAN(1.0)*4 (suews_phys_anohm.f95:1015)
ALB = np.array([0.12, 0.15, 0.12, 0.18, 0.21, 0.21, 0.1])
EMIS = np.array([0.95, 0.91, 0.98, 0.98, 0.93, 0.93, 0.95])
CP = np.array([2e6, 2e6, 1e6, 1e6, 1e6, 2e6, 4.1e6])
KK = np.full(7, 1.2)
CH = np.full(7, 4.0)


def P(a):
    return a.ctypes.data_as(dp)


def _minpack():
    L = ol.lib()
    LM = C.CFUNCTYPE(None, C.c_int, C.c_int, dp, dp, C.c_void_p)
    HY = C.CFUNCTYPE(None, C.c_int, dp, dp, C.c_void_p)
    L.so_lmdif1.argtypes = [LM, C.c_void_p, C.c_int, C.c_int, dp, dp, C.c_double]
    L.so_lmdif1.restype = C.c_int
    L.so_hybrd1.argtypes = [HY, C.c_void_p, C.c_int, dp, dp, C.c_double]
    L.so_hybrd1.restype = C.c_int
    return L, LM, HY


def test_lmdif1_is_scipys_minpack_bit_for_bit():
    """so_lmdif1 (ftol = xtol = tol, gtol = 0, maxfev = 200(n+1), factor 100, mode 1) against scipy's lmdif."""
    from scipy.optimize import leastsq
    L, LM, _ = _minpack()
    rng = np.random.default_rng(1)
    for trial in range(120):
        m = int(rng.integers(4, 16))
        t = np.sort(rng.choice(np.arange(24.0), m, replace=False))
        mean, amp, tp = rng.uniform(-5, 300), rng.uniform(1, 300), rng.uniform(3, 9)
        y = mean + amp * np.sin(2 * PI4 / 24 * (t - tp)) + rng.normal(0, 0.05 * amp, m)

        def res(x):
            return x[0] + x[1] * np.sin(2 * PI4 / 24 * (t - x[2])) - y
        x0 = np.array([y.mean(), (y.max() - y.min()) / 2, 12.0])
        xs, _cov, info, _msg, ier = leastsq(res, x0.copy(), full_output=True, ftol=1e-5, xtol=1e-5, gtol=0.0, maxfev=800,
                                            factor=100)
        calls = [0]

        def cb(mm, nn, xp, fp, _ctx):
            calls[0] += 1
            r = res(np.array([xp[0], xp[1], xp[2]]))
            for i in range(mm):
                fp[i] = r[i]
        x = x0.copy()
        fv = np.zeros(m)
        rc = L.so_lmdif1(LM(cb), None, m, 3, P(x), P(fv), 1e-5)
        assert np.array_equal(x, xs), (trial, x, xs)
        assert calls[0] == info["nfev"] and rc == ier, (trial, calls[0], info["nfev"], rc, ier)


def test_hybrd1_is_scipys_minpack_bit_for_bit():
    """so_hybrd1 (mode 2 with unit scaling, ml = mu = n - 1, maxfev = 200(n+1), factor 100) against scipy's hybrd, for the
    n = 1 residuals AnOHM has (x - const is the usual one: the MIN(Ts, -40.) clamp removes the dependence on x) and for
    n = 2, 3 systems, which exercise qform / dogleg / r1updt / r1mpyq beyond their scalar forms."""
    from scipy.optimize import fsolve
    L, _, HY = _minpack()
    rng = np.random.default_rng(2)
    tol = float(np.float32(1e-20))
    for trial in range(90):
        a, b, c = rng.uniform(0.5, 3), rng.uniform(-5, 5), rng.uniform(0.1, 2)
        kind = trial % 3

        def g(x):
            x = float(x)
            return x - b if kind == 0 else (a * x + b + c * np.sin(x) if kind == 1 else np.exp(0.3 * x) - a)
        xs, info, _ier, _msg = fsolve(lambda v: [g(v[0])], [10.0], full_output=True, xtol=tol, maxfev=400, factor=100, diag=[1.0])
        calls = [0]

        def cb(_n, xp, fp, _ctx):
            calls[0] += 1
            fp[0] = g(xp[0])
        x = np.array([10.0])
        fv = np.zeros(1)
        L.so_hybrd1(HY(cb), None, 1, P(x), P(fv), tol)
        assert x[0] == xs[0], (trial, kind, x[0], xs[0])
        assert calls[0] + 2 == info["nfev"]      # fsolve spends two calls of its own on shape checks
    for trial in range(40):
        n = 2 + trial % 2
        A = rng.normal(size=(n, n)) + 3 * np.eye(n)
        bb = rng.normal(size=n)

        def gv(x):
            x = np.asarray(x)
            return A @ x + 0.3 * np.sin(x) - bb
        xs, info, _ier, _msg = fsolve(gv, np.ones(n), full_output=True, xtol=1e-10, maxfev=200 * (n + 1), factor=100, diag=np.ones(n))

        def cbv(nn, xp, fp, _ctx):
            r = gv([xp[i] for i in range(nn)])
            for i in range(nn):
                fp[i] = r[i]
        x = np.ones(n)
        fv = np.zeros(n)
        L.so_hybrd1(HY(cbv), None, n, P(x), P(fv), 1e-10)
        assert np.array_equal(x, xs), (trial, x, xs)


def test_shapefit_recovers_a_sinusoid_and_folds_negative_amplitudes():
    L = ol.lib()
    L.so_AnOHM_ShapeFit.argtypes = [C.c_int, dp, dp, dp, dp, dp]
    t = np.arange(6.0, 19.0)
    for mean, amp, tpeak in ((300.0, 250.0, 12.5), (15.0, 4.0, 15.0), (280.0, 6.0, 9.0)):
        y = mean + amp * np.sin(2 * PI4 / 24 * (t - (tpeak - 6)))
        a, m, tp = C.c_double((y.max() - y.min()) / 2), C.c_double(y.mean()), C.c_double(12.0)
        L.so_AnOHM_ShapeFit(len(t), P(t), P(y), C.byref(a), C.byref(m), C.byref(tp))
        assert abs(a.value - amp) < 1e-4 * amp and abs(m.value - mean) < 1e-4 * abs(mean) and abs(tp.value - tpeak) < 1e-3
        # a start on the wrong side of the wave converges to -amp, half a period away: folded back (anohm.f95:989-995)
        a, m, tp = C.c_double(-(y.max() - y.min()) / 2), C.c_double(y.mean()), C.c_double(0.0)
        L.so_AnOHM_ShapeFit(len(t), P(t), P(y), C.byref(a), C.byref(m), C.byref(tp))
        assert a.value > 0 and abs(a.value - amp) < 1e-3 * amp
        assert min(abs(tp.value - tpeak), 24 - abs(tp.value - tpeak)) < 1e-2


def _anohm_cell(**over):
    d = fx.state_init(2005)
    d.update(cpanohm=list(CP), kkanohm=list(KK), chanohm=list(CH), storageheatmethod=[3])
    d.update(over)
    return d


def test_coefficients_are_formed_once_a_day_and_reloaded():
    d = _anohm_cell()
    met = fx.forcing(2005)[:288 * 3]
    cell = ol.Cell(d)
    seen = []
    step = 36
    for i in range(0, met.shape[0], step):
        # MetForcingData_grid is the whole block in every call, as in the reference binary
        mf = np.asfortranarray(met)
        ol.lib().so_set_met_grid(mf.ctypes.data_as(dp), mf.shape[0])
        out = cell.run(met[i:i + step], want=("SUEWS",))
        seen.append((int(met[i + step - 1, 1]), cell.get_state("anohm_coeff_grid_day").copy(), out["SUEWS"][-1, 5:], cell.get_state("dqndt")[0]))
    ol.lib().so_set_met_grid(None, 0)
    days = sorted({s[0] for s in seen})
    assert len(days) >= 3
    for day in days:
        coefs = [s[1] for s in seen if s[0] == day]
        assert all(np.array_equal(coefs[0], c) for c in coefs)          # reloaded, not refitted, within a day
    c1 = [s[1] for s in seen if s[0] == days[0]][0]
    c2 = [s[1] for s in seen if s[0] == days[1]][0]
    assert not np.array_equal(c1, c2) and np.all(np.isfinite(c1)) and np.all(np.isfinite(c2))
    # the block ends on 00:00 of a fourth day, which has one on-the-hour row in it: not enough data (< 6 daylight rows), so
    # the step falls back on the last good day (id_save, anohm.f95:96-110) and reloads that day's coefficients
    assert len(days) == 4 and cell.get_state("anohm_id_save")[0] == days[-2] and cell.get_state("anohm_coef_id_save")[0] == days[-2]
    assert np.array_equal(seen[-1][1], [s[1] for s in seen if s[0] == days[-2]][0])


def test_qs_is_ohms_formula_with_dqndt_prev():
    """QS = a1 Qn + a2 dQn/dt + a3 with the area-weighted day coefficients and the rate of change of the PREVIOUS step
    (anohm.f95:139 passes dqndt_prev where ohm.f95:140 passes dqndt_next)."""
    d = _anohm_cell()
    met = fx.forcing(2005)[:200]
    cell = ol.Cell(d)
    mf = np.asfortranarray(met)
    ol.lib().so_set_met_grid(mf.ctypes.data_as(dp), mf.shape[0])
    cell.run(met[:150], want=("SUEWS",))
    for i in range(150, 160):
        dq_prev = cell.get_state("dqndt")[0]
        out = cell.run(met[i:i + 1], want=("SUEWS",))["SUEWS"][0, 5:]
        co = cell.get_state("anohm_coeff_grid_day").reshape(3, 7)
        a = co @ np.asarray(d["sfr"])
        qn_use = out[5] + (out[6] if d["ohmincqf"][0] == 1 else 0.0)
        assert abs(out[7] - (qn_use * a[0] + dq_prev * a[1] + a[2])) <= 1e-11 * max(1.0, abs(out[7]))
        assert dq_prev != cell.get_state("dqndt")[0]
    ol.lib().so_set_met_grid(None, 0)


def test_bowen_ratio_is_the_closed_form_its_residual_reduces_to():
    """MIN(Ts - 273.15, -40.) pins the surface temperature at -40 degC whenever the modelled one is warmer, so
    Bo = SUM(cp (-40 - Ta)) / SUM(xSM Lv (qs(-40) - qa)) and the land coefficients follow from it (anohm.f95:1337-1355)."""
    L = ol.lib()
    L.so_anohm_coefs.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_double, dp, dp, dp, dp, dp, dp, dp]
    L.so_AnOHM_coef_land_cal.argtypes = [C.c_double] * 14 + [dp] * 6
    met = np.asfortranarray(fx.forcing(2004)[288 * 180:288 * 181])        # a summer day
    day = int(met[0, 1])
    moist = np.array([0.5, 0.2, 1.4, 0.9, 0.7, 0.3, 1.0])
    qf = 25.0
    out = np.zeros(21)
    L.so_anohm_coefs(P(met), met.shape[0], day, 2, qf, P(moist), P(ALB), P(EMIS), P(CP), P(KK), P(CH), P(out))
    rows = met[(met[:, 3] == 0) & (met[:, 14] > 5)]
    Ta, RH, pres = rows[:, 11], rows[:, 10], rows[:, 12]
    f32 = lambda v: float(np.float32(v))  # noqa: E731
    esat = lambda T: f32(6.106) * np.exp(f32(17.27) * T / (f32(237.3) + T))  # noqa: E731
    w = f32(0.0180153) / f32(0.028965)
    qs = w * esat(-40.0) / pres
    qa = w * esat(Ta) * RH / 100 / pres
    for is_ in range(6):
        xsm = min(moist[is_], 1.0)
        bo = np.sum(f32(1006.38) * (-40.0 - Ta)) / np.sum(xsm * f32(2264.705E3) * (qs - qa))
        # the fitted forcing scales are not exposed; recover them through a second surface with the same Bo instead:
        # the coefficient routine is a pure function of (forcing scales, surface properties, Bo), so two surfaces with
        # equal properties and equal wetness must get equal coefficients
        moist2 = moist.copy()
        j = (is_ + 1) % 6
        moist2[j] = moist[is_]
        alb2, emis2, cp2 = ALB.copy(), EMIS.copy(), CP.copy()
        alb2[j], emis2[j], cp2[j] = ALB[is_], EMIS[is_], CP[is_]
        out2 = np.zeros(21)
        L.so_anohm_coefs(P(met), met.shape[0], day, 2, qf, P(moist2), P(alb2), P(emis2), P(cp2), P(KK), P(CH), P(out2))
        assert np.array_equal(out2[[j, 7 + j, 14 + j]], out[[is_, 7 + is_, 14 + is_]])
        assert np.isfinite(bo)
    assert np.all(np.isfinite(out))


def _host_lib(tmp_path_factory, *defs):
    src = os.path.join(ol.ROOT, "tests", "host_anohm.cpp")
    out = str(tmp_path_factory.mktemp("host_anohm") / "host_anohm.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", *defs, "-o", out, src])
    return C.CDLL(out)


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    return _host_lib(tmp_path_factory)


@pytest.mark.parametrize("dT,dRH", [(0.0, 0.0), (1.7, -4.0), (-1.2, 3.5)])
def test_device_code_on_the_host_equals_the_oracle_bit_for_bit(host_lib, dT, dRH):
    """dev_anohm.cuh (the kernel's source, host build) against so_anohm.c / so_minpack.c: every eighth day of 2004, random
    wetness (with a dry surface: the xSM = 0 arm), random QF, EmissionsMethod 0 (QF from the forcing file) and > 0, with
    and without the per-cell forcing offsets."""
    L = ol.lib()
    L.so_anohm_coefs.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_double, dp, dp, dp, dp, dp, dp, dp]
    host_lib.an_host_day_coefs.argtypes = [dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double] + [dp] * 7
    met = fx.forcing(2004)
    hrs = np.ascontiguousarray(met[met[:, 3] == 0][:, [1, 2, 14, 11, 10, 12, 9, 8]])
    mo = met.copy()                                   # what the step sees of this cell's forcing
    mo[:, 11] += dT
    if dRH != 0:
        mo[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), mo[:, 10] + dRH))
    mo = np.asfortranarray(mo)
    rng = np.random.default_rng(7)
    for day in range(1, 367, 8):
        moist = rng.uniform(0, 1.2, 7)
        moist[rng.integers(0, 7)] = 0.0
        qf = rng.uniform(5, 60)
        em = int(rng.integers(0, 3))
        a, b = np.zeros(21), np.zeros(21)
        L.so_anohm_coefs(P(mo), mo.shape[0], day, em, qf, P(moist), P(ALB), P(EMIS), P(CP), P(KK), P(CH), P(a))
        host_lib.an_host_day_coefs(P(hrs), hrs.shape[0], day, dT, dRH, em, qf, P(moist), P(ALB), P(EMIS), P(CP), P(KK), P(CH), P(b))
        assert np.all(np.isfinite(a)), day
        assert np.array_equal(a, b), (day, a, b)


def test_the_fit_amplifies_last_ulp_differences_of_libm(host_lib, tmp_path_factory):
    """Why the GPU parity tests of this scheme cannot hold 1e-10: the sinusoid fit stops at a relative tolerance of 1e-5
    (anohm.f95:972) with a forward-difference Jacobian of step 1.5e-8 |x| (fdjac2), so a last-ulp difference in sin / cos /
    exp -- CUDA's libm against glibc's, or one glibc against another under the reference itself -- moves the iterates by
    ~1e-8 and can move the stopping point.  Measured here with the same source and libm results shifted by one ulp:
    the day's coefficients move by ~1e-7 (median) and up to a few 1e-4 on single days.  The GPU tests therefore bound this
    scheme's outputs at 1e-5 per step and say so; the algorithm itself is pinned by the bit-for-bit test above."""
    ulp = _host_lib(tmp_path_factory, "-DSU_ANOHM_ULP")
    for lib in (host_lib, ulp):
        lib.an_host_day_coefs.argtypes = [dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double] + [dp] * 7
    met = fx.forcing(2004)
    hrs = np.ascontiguousarray(met[met[:, 3] == 0][:, [1, 2, 14, 11, 10, 12, 9, 8]])
    rng = np.random.default_rng(0)
    moved = []
    for day in range(1, 367, 3):
        moist = rng.uniform(0.05, 1.2, 7)
        qf = rng.uniform(5, 60)
        a, b = np.zeros(21), np.zeros(21)
        for lib, o in ((host_lib, a), (ulp, b)):
            lib.an_host_day_coefs(P(hrs), hrs.shape[0], day, 0.0, 0.0, 2, qf, P(moist), P(ALB), P(EMIS), P(CP), P(KK), P(CH), P(o))
        moved.append(float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-3))))
    moved = np.array(moved)
    print("AnOHM coefficients under one-ulp libm shifts: median %.1e, 90%% %.1e, max %.1e" % (np.median(moved), np.percentile(moved, 90), moved.max()))
    assert np.median(moved) > 1e-10          # far above the 1e-10 of the other schemes: the fit is the amplifier
    assert moved.max() < 1e-2                # ... and still a small change of the coefficients
