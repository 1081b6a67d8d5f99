"""ESTM (StorageHeatMethod 4/14) -- CPU checks of the oracle restatement and of the host-side initial state.

PARITY UNPINNED: the reference repository has ESTM inputs but no ESTM output of any run, and the
reference cannot be compiled here.  These tests hold the restatement to the properties the scheme must have
(suews_phys_estm.f95: heatcond1d :480-510, NewtonPolynomial :206-250, ESTM :970-1605)."""
import ctypes as C

import numpy as np

import fixtures as fx
import oracle_lib as ol

dp = C.POINTER(C.c_double)


def _ptr(a):
    return a.ctypes.data_as(dp)


def test_heatcond1d_uniform_slab_is_steady_and_fluxes_balance():
    L = ol.lib()
    L.so_heatcond1d.argtypes = [dp, dp, dp, C.c_double, dp, dp, C.c_int, dp, C.POINTER(C.c_int)]
    dx = np.array([0.05, 0.15, 0.12, 0, 0]); k = np.array([1.142, 0.93, 0.05, 0, 0]); r = np.array([1.6e6, 1.5e6, 2.9e5, 0, 0])
    bct = (C.c_int * 2)(0, 0)
    # uniform temperature equal to both boundary values: nothing moves, no storage flux
    T = np.full(5, 283.15); qs = np.zeros(1); bc = np.array([283.15, 283.15])
    L.so_heatcond1d(_ptr(T), _ptr(qs), _ptr(dx), 300.0, _ptr(k), _ptr(r), 3, _ptr(bc), bct)
    assert qs[0] == 0.0 and np.all(T[:3] == 283.15)
    # a warm outer face: heat enters (Qs > 0), and the energy gained by the layers equals Qs * dt (explicit scheme,
    # interface temperatures weighted by conductance: the interior fluxes cancel exactly in exact arithmetic)
    T = np.array([283.15, 284.0, 285.0, 0, 0]); T0 = T.copy(); bc = np.array([290.0, 286.0])
    L.so_heatcond1d(_ptr(T), _ptr(qs), _ptr(dx), 300.0, _ptr(k), _ptr(r), 3, _ptr(bc), bct)
    gained = float(np.sum((T[:3] - T0[:3]) * r[:3] * dx[:3])) / 300.0
    assert qs[0] > 0
    assert abs(gained - qs[0]) <= 1e-9 * abs(qs[0])


def test_newton_polynomial_solves_the_surface_balance_and_keeps_its_saved_flag():
    L = ol.lib()
    L.so_NewtonPolynomial.restype = C.c_double
    L.so_NewtonPolynomial.argtypes = [C.c_double, dp, C.c_double, C.c_int, dp]
    sb = float(np.float32(5.67051e-8))
    P = np.array([0.1 * sb, 0.0, 0.0, 53.1, -53.1 * 290.0 - 400.0])
    flag = np.zeros(1)
    x = L.so_NewtonPolynomial(293.0, _ptr(P), float(np.float32(1e-4)), 100, _ptr(flag))
    assert flag[0] == 1.0
    assert abs(P[0] * x ** 4 + P[3] * x + P[4]) < 1e-6
    # maxiter = 1: no second pass, so no convergence test -- but `converged` is SAVE'd .TRUE. from the call before,
    # and the un-converged iterate is returned instead of x0 (estm.f95:222, :245-248)
    x1 = L.so_NewtonPolynomial(250.0, _ptr(P), float(np.float32(1e-4)), 1, _ptr(flag))
    assert x1 != 250.0
    flag[0] = 0.0
    assert L.so_NewtonPolynomial(250.0, _ptr(P), float(np.float32(1e-4)), 1, _ptr(flag)) == 250.0


def _estm_cell(**over):
    d = dict(fx.state_init(2005))
    d["storageheatmethod"] = [4]
    ep = fx.estm_example(d["surfacearea"][0], **over)
    d.update({k.lower(): v for k, v in ep.items()})
    return d, ep


def test_host_initial_state_is_the_oracle_s():
    from suews_b200 import estm
    d, ep = _estm_cell()
    met = fx.forcing(2005)[:1]
    ts = fx.estm_ts_forcing(met)
    cell = ol.Cell(d)
    cell.estm_start(ts[0])
    st = estm.initial_state(ep, d["sfr"], d["emis"], ts[0])
    for name, v in st.items():
        assert np.array_equal(np.asarray(v, dtype=float), cell.get_state(name)), name


def test_estm_run_is_consistent():
    d, _ep = _estm_cell()
    met = fx.forcing(2005)[:288]
    ts = fx.estm_ts_forcing(met)
    cell = ol.Cell(d)
    cell.estm_start(ts[0])
    out = cell.run(met, want=("SUEWS", "ESTM"), ts5=ts)
    su, es = out["SUEWS"], out["ESTM"]
    assert np.array_equal(su[:, :5], es[:, :5])
    assert np.array_equal(su[:, 5 + 7], es[:, 5 + 0])                    # QS of the SUEWS line is ESTM's QS
    parts = es[:, 5 + 2] + es[:, 5 + 3] + es[:, 5 + 4] + es[:, 5 + 5]    # wall + roof + ground + internal
    np.testing.assert_allclose(parts, es[:, 5], rtol=1e-12, atol=1e-9)
    # layer counts of the example: 3 wall / roof / internal layers, 4 ground layers; the rest is -999
    assert np.all(es[:, 5 + 6 + 3:5 + 6 + 5] == -999) and np.all(es[:, 5 + 16 + 4] == -999)
    assert np.all(es[:, 5 + 16:5 + 16 + 4] > 250) and np.all(es[:, 5 + 16:5 + 16 + 4] < 300)
    assert np.all(es[:, 5 + 26] == 20.0 + fx.C2K)                        # evolveTibld = 0: measured indoor temperature
    # one ESTM call per trip of the surface-temperature loop (driver.f95:642, :776): at least one per step
    assert cell.get_state("Tair2Set")[0] >= 288
    # forced surface temperature: the outer wall / roof / ground faces follow the forcing
    assert cell.get_state("T0_roof")[0] == ts[-1, 5] + fx.C2K


def test_estm_without_forcing_rows_is_a_no_data_run():
    # the oracle feeds -999 when no block is given; the run still completes (temperatures are then nonsense) --
    # what matters is that the product rejects that case (tests/test_gpu_parity.py::test_estm_without_inputs_fails_loudly)
    d, _ep = _estm_cell()
    cell = ol.Cell(d)
    cell.estm_start(fx.estm_ts_forcing(fx.forcing(2005)[:1])[0])
    out = cell.run(fx.forcing(2005)[:3], want=("ESTM",))
    assert out["ESTM"].shape == (3, 32)


def test_estm_forcing_disaggregation_matches_the_reference_5min_file():
    """DisaggregateESTM (suews_ctrl_input.f95:1650-1849): the reference wrote its 5-min series (f9.4) next to the hourly
    input in Release/InputTables/2020a; the numpy restatement reproduces every printed digit and every time stamp"""
    from suews_b200.forcing import disaggregate_estm, read_estm_ts
    fast = disaggregate_estm(read_estm_ts(fx.GOLDEN + "/estm_ts_2005_60min.txt"))
    ref = read_estm_ts(fx.GOLDEN + "/ref_estm_ts_2005_5min.txt")
    assert fast.shape == ref.shape == (1440, 13)
    assert np.array_equal(fast[:, :4], ref[:, :4])
    assert np.abs(fast[:, 4:] - ref[:, 4:]).max() <= 0.5e-4 + 1e-12
