"""The device-side timestep compiled for the host, against the oracle BIT FOR BIT.

tests/host_step.cpp includes suews_b200/csrc/dev_step.cuh -- the source of the CUDA kernel's cell-step with every stage header it
pulls in -- defines the device qualifiers away, maps the CUDA math wrappers to libm (the functions the oracle calls) and drives
one cell the way suews_run_kernel drives a thread under output policy FULL.  Hoisting out of the surface-temperature loop, the
deferred state update, the wet-bulb memo and its float screen, cycle detection, the in-place stores, the per-instance
specialisations: each of them claims to leave the reference's arithmetic untouched.  Here the claim is checked on the CPU,
without tolerance: every output column of every step and the state at the end equal the oracle's to the last bit, for the
BaseRun over a whole year, the scheme sweep of the GPU parity tests, ESTM, AnOHM and BLUEWS.  The single documented
exception is the psi-hat integration of the roughness-sublayer diagnostics (dev_rsl.cuh adds the same terms in another order,
<= n 2^-53 relative): T2, Q2, U10, RH2 and the RSL profile block are compared at 1e-10 (they agree to ~1e-15 nearly everywhere).
What remains between this build and the GPU is the math library alone (and the fast paths of dev_common.cuh / dev_math.cuh,
which the device selftests check): that is the 1e-10 of the GPU parity tests."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fixtures as fx  # noqa: E402
import oracle_lib as ol  # noqa: E402
from suews_b200.batch import all_fields  # noqa: E402

dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
NCOL = {"SUEWS": 82, "snow": 98, "RSL": 133, "DailyState": 45, "SOLWEIG": 26, "ESTM": 27, "BL": 17}
GROUPS = ["SUEWS", "snow", "RSL", "DailyState", "SOLWEIG", "ESTM", "BL"]
RSL_COLS = [5 + 78, 5 + 79, 5 + 80, 5 + 81]          # T2 Q2 U10 RH2 of the SUEWS line: formed from the RSL profiles
PARITY_COLS = [5, 6, 7, 8, 9, 16, 71]                # QN QF QS QH QE RO SWE in the 82-column line
INTERNAL = ("alb1_hidden", "lv_switch", "rsl_cache")  # the library's own memo fields: no counterpart in the oracle's state


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_step") / "host_step.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", out,
                           os.path.join(ol.ROOT, "tests", "host_step.cpp")])
    H = C.CDLL(out)
    H.host_step_run.argtypes = [ip, dp, dp, dp, dp, dp, C.c_int, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(dp),
                                C.POINTER(C.c_ulonglong), C.POINTER(C.c_longlong), C.c_int, C.c_int, dp]
    return H


def _flat(d, kinds):
    low = {k.lower(): v for k, v in d.items()}
    out = []
    for nm, kind, n in all_fields():
        if kind not in kinds:
            continue
        if nm.lower() in low:
            v = np.atleast_1d(np.asarray(low[nm.lower()], dtype=np.float64)).ravel()
            assert v.size == n, (nm, v.size, n)
        else:
            v = np.full(n, -999.0 if nm == "anohm_coeff_grid_day" else 0.0)   # as suews_b200_create leaves them
        out.append(v)
    return np.concatenate(out)


def run_host(H, d, met, groups, instance=0, record=False, ts5=None, cbl_ini=None, kind=0, flux_only=False, acc=None):
    """-> (fatal code, {group: block [n, 5 + ncol]}, {state name: values}, warning masks)
    kind 1 / 2: no blocks; the 82 / 7 running sums are added to `acc` (output policies P0+ / P0 of the library)"""
    ci = np.ascontiguousarray(_flat(d, (0,)).astype(np.int32))
    cd, ce, st = (np.ascontiguousarray(_flat(d, (k,))) for k in (1, 2, 3))
    es = np.ascontiguousarray(_flat(d, (4,))) if record else None
    n = met.shape[0]
    m = np.asfortranarray(met)
    t = np.asfortranarray(ts5) if ts5 is not None else None
    ini = np.ascontiguousarray(cbl_ini) if cbl_ini is not None else None
    blks = {g: np.full((n, 5 + NCOL[g]), -999.0, order="F") for g in groups}
    arr = (dp * 7)(*[blks[g].ctypes.data_as(dp) if g in blks else None for g in GROUPS])
    warn = (C.c_ulonglong * 2)()
    iters = C.c_longlong()
    rc = H.host_step_run(ci.ctypes.data_as(ip), cd.ctypes.data_as(dp), ce.ctypes.data_as(dp), st.ctypes.data_as(dp),
                         es.ctypes.data_as(dp) if record else None, m.ctypes.data_as(dp), n,
                         t.ctypes.data_as(dp) if t is not None else None, ini.ctypes.data_as(dp) if ini is not None else None,
                         ini.shape[0] if ini is not None else 0, int(d["dt_since_start"][0]),
                         int(d.get("tstep_prev", d["tstep"])[0]), instance, arr, warn, C.byref(iters), int(kind), int(bool(flux_only)),
                         acc.ctypes.data_as(dp) if acc is not None else None)
    state, k3, k4 = {}, 0, 0
    for nm, kind, cnt in all_fields():
        if kind == 3:
            state[nm] = st[k3:k3 + cnt]
            k3 += cnt
        elif kind == 4 and record:
            state[nm] = es[k4:k4 + cnt]
            k4 += cnt
    return rc, blks, state, (warn[0], warn[1])


def same_bits(a, b):
    return (a == b) | (np.isnan(a) & np.isnan(b))


FC_COLS = [5 + 58, 5 + 59, 5 + 60]                   # Fc FcPhoto FcRespi
# the re-ordered psi-hat sums: <= n 2^-53 of the sums, more once a small difference of them is taken (U10 near calm, T2
# near 0 degC); 2e-12 is the largest seen in 4 000 random cases
LOOSE_TOL = 1e-10


def check_blocks(blks, ref, where="", t2_feeds_co2=False):
    """t2_feeds_co2: with gsModel 3 / 4 the biogenic CO2 fluxes are evaluated at the RSL's 2 m temperature (driver.f95:1224-1262),
    so they inherit its last-ulp difference."""
    for g, a in blks.items():
        b = ref[g]
        ok = same_bits(a, b)
        loose = (RSL_COLS + (FC_COLS if t2_feeds_co2 else [])) if g == "SUEWS" else (list(range(5, a.shape[1])) if g == "RSL" else [])
        strict = np.ones(a.shape[1], dtype=bool)
        strict[loose] = False
        bad = np.argwhere(~ok[:, strict])
        assert bad.size == 0, (where, g, "first differing (row, strict column):", bad[:5].tolist())
        if loose:
            al, bl = a[:, loose], b[:, loose]
            assert np.array_equal(np.isnan(al), np.isnan(bl)), (where, g, "NaN pattern")
            fin = ~np.isnan(bl)
            if fin.any():
                with np.errstate(invalid="ignore"):               # (equal infinities are equal)
                    e = np.where(al[fin] == bl[fin], 0.0, np.abs(al[fin] - bl[fin]) / np.maximum(np.abs(bl[fin]), 1e-3))
                assert float(e.max()) <= LOOSE_TOL, (where, g, float(e.max()))


def check_warnings(masks, ref, where=""):
    """the warnings raised on the way (ErrorHint codes that let the run continue) are the same set"""
    raised = {c for c in range(128) if ref["_warn"][c] > 0}
    mask = {c for c in range(128) if (masks[c >> 6] >> (c & 63)) & 1}
    assert mask == raised, (where, sorted(mask), sorted(raised))


def check_state(state, cell, names=None, where=""):
    n = 0
    for nm, v in state.items():
        if nm in INTERNAL or (names is not None and nm not in names):
            continue
        try:
            r = cell.get_state(nm)
        except KeyError:
            continue
        assert np.all(same_bits(v, r)), (where, nm, v, r)
        n += 1
    return n


def test_baserun_year_2004_is_bit_identical(host):
    """All 105 408 steps of 2004 (snow on the ground in January, melt, summer, rain, every daily update) from the
    reference's dump: the 82-column line, the snow line and the DailyState line of every step, and the state at the end."""
    d = fx.state_init(2004)
    met = fx.forcing(2004)
    groups = ("SUEWS", "snow", "DailyState")
    rc, blks, state, _w = run_host(host, d, met, groups)
    assert rc == 0
    cell = ol.Cell(d)
    ref = cell.run(met, want=groups)
    check_blocks(blks, ref, "2004")
    assert check_state(state, cell, where="end of 2004") >= 25
    check_warnings(_w, ref, "2004")
    differ = int((~same_bits(blks["SUEWS"][:, RSL_COLS], ref["SUEWS"][:, RSL_COLS])).sum())
    assert differ < 0.5 * 4 * met.shape[0]          # even the re-ordered psi-hat sums mostly land on the same double


@pytest.mark.parametrize("instance", [0, 2], ids=["generic", "snow-on-OHM"])
def test_baserun_2005_every_output_group(host, instance):
    """The five golden days of 2005, all five output groups (SOLWEIG and the RSL profiles included), through the generic
    instance and through the snow-on specialisation the BaseRun's bucket launches."""
    d = fx.state_init(2005)
    met = fx.forcing(2005)[:1440]
    groups = ("SUEWS", "snow", "RSL", "DailyState", "SOLWEIG")
    rc, blks, state, _w = run_host(host, d, met, groups, instance=instance)
    assert rc == 0
    cell = ol.Cell(d)
    check_blocks(blks, cell.run(met, want=groups), "2005")
    check_state(state, cell, where="2005 day 5")


BIOCO2 = dict(theta_bioco2=[0.96] * 3, alpha_enh_bioco2=[0.016] * 3, beta_enh_bioco2=[33.454] * 3)
SWEEP = [
    dict(snowuse=0, netradiationmethod=3, emissionsmethod=2),
    dict(netradiationmethod=1), dict(netradiationmethod=2), dict(netradiationmethod=3), dict(netradiationmethod=11),
    dict(netradiationmethod=100), dict(netradiationmethod=300), dict(netradiationmethod=0),
    dict(netradiationmethod=0, snowuse=0), dict(netradiationmethod=12), dict(netradiationmethod=200),
    dict(stabilitymethod=2), dict(stabilitymethod=4), dict(stabilitymethod=1), dict(stabilitymethod=5), dict(stabilitymethod=6),
    dict(stabilitymethod=7),
    dict(roughlenheatmethod=1), dict(roughlenheatmethod=3), dict(roughlenheatmethod=4), dict(roughlenheatmethod=5),
    dict(roughlenmommethod=1), dict(roughlenmommethod=3),
    dict(emissionsmethod=0), dict(emissionsmethod=2), dict(emissionsmethod=3), dict(emissionsmethod=4),
    dict(emissionsmethod=5), dict(emissionsmethod=6),
    dict(storageheatmethod=0), dict(ohmincqf=1), dict(evapmethod=1), dict(aerodynamicresistancemethod=1),
    dict(aerodynamicresistancemethod=3), dict(gsmodel=1), dict(smdmethod=1), dict(waterusemethod=1),
    dict(basetmethod=2), dict(laicalcyes=0),
    dict(emissionsmethod=11), dict(emissionsmethod=14), dict(emissionsmethod=22, **BIOCO2), dict(emissionsmethod=33, **BIOCO2),
    dict(emissionsmethod=36, theta_bioco2=[0.9, 0.96, 0.8], alpha_enh_bioco2=[0.016] * 3, beta_enh_bioco2=[33.454] * 3),
    dict(emissionsmethod=41), dict(emissionsmethod=45, gsmodel=4), dict(emissionsmethod=12, gsmodel=3),
]


def _with(d, over):
    for k, v in over.items():
        d[k] = list(v) if isinstance(v, (list, tuple)) else [v]
    if over.get("snowuse", 1) == 0:
        for k in ("snowpack", "snowfrac", "snowwater"):
            d[k] = [0.0] * 7
    return d


@pytest.mark.parametrize("year,days", [(2005, 2), (2004, 3)], ids=["mild-start", "winter-start"])
def test_scheme_sweep_is_bit_identical(host, year, days):
    """Every scheme option of the GPU parity sweep (the reference's BTS_config.nml physics sweep and the stability /
    emission / radiation methods it leaves out), from the mild 2005 start and from the 2004 start with snow on the ground."""
    met0 = fx.forcing(year)[:288 * days]
    for over in SWEEP:
        d = _with(fx.state_init(year), over)
        met = met0.copy()
        if over.get("smdmethod", 0) > 0:
            met[:, 19] = 30.0
        groups = ("SUEWS", "snow") if d["snowuse"][0] == 1 else ("SUEWS",)
        inst = [0] + ([1 if d["snowuse"][0] == 0 else 2] if d["storageheatmethod"][0] in (0, 1) else [])
        cell = ol.Cell(d)
        ref = cell.run(met, want=groups)
        for instance in inst:
            rc, blks, state, _w = run_host(host, d, met, groups, instance=instance)
            assert rc == 0, over
            check_blocks(blks, ref, (year, over, instance), t2_feeds_co2=over.get("gsmodel", 2) in (3, 4))
            check_warnings(_w, ref, (year, over, instance))
            check_state(state, cell, where=(year, over, instance))


def test_perturbed_cells_are_bit_identical(host):
    """The synthetic cells of the bench (perturbed fractions, site, morphology, forcing offsets) for the bench's scheme sets
    that need no extra input (default, snow, BaseRun, NetRadiationMethod 13, stability 2 / 4, AnOHM): one day each against a
    per-cell oracle run."""
    from suews_b200 import synth
    met = fx.forcing(2005)[:288]
    for scheme in ("default", "snow", "baserun", "netrad13", "netrad300", "stab2", "stab4", "anohm"):
        base = synth.base_with_scheme(fx.state_init(2005), scheme)
        arrs = synth.cell_site_arrays(0, 24, base, offsets=True)
        for c in range(24):
            d = dict(base)
            for k, v in arrs.items():
                d[k.lower()] = list(np.atleast_1d(v[c]))
            # the cell's forcing offsets are applied by the step itself; the oracle gets the shifted series
            m = met.copy()
            m[:, 11] += d["dtemp_c"][0]
            m[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), m[:, 10] + d["drh"][0]))
            cell = ol.Cell({k: v for k, v in d.items() if k not in ("dtemp_c", "drh")})
            ref = cell.run(m, want=("SUEWS",))
            rc, blks, state, _w = run_host(host, d, met, ("SUEWS",), record=(scheme == "anohm"))
            assert rc == 0
            check_blocks(blks, ref, (scheme, c))
            check_state(state, cell, where=(scheme, c))


ESTM_VARIANTS = [dict(), dict(evolveTibld=1, ibldCHmod=1), dict(evolveTibld=2, ibldCHmod=2), dict(TsurfChoice=1),
                 dict(TsurfChoice=2, evolveTibld=1), dict(storageheatmethod=14)]


@pytest.mark.parametrize("over", ESTM_VARIANTS, ids=lambda o: "-".join("%s%s" % kv for kv in o.items()) or "baserun")
def test_estm_is_bit_identical(host, over):
    """StorageHeatMethod 4 / 14 over two days: the 82-column line, the ESTM line and the element temperatures ESTM carries."""
    from suews_b200 import estm
    d = dict(fx.state_init(2005))
    over = dict(over)
    d["storageheatmethod"] = [over.pop("storageheatmethod", 4)]
    ep = fx.estm_example(d["surfacearea"][0], **over)
    d.update({k.lower(): v for k, v in ep.items()})
    met = fx.forcing(2005)[:576]
    ts = fx.estm_ts_forcing(met)
    cell = ol.Cell(d)
    cell.estm_start(ts[0])
    ref = cell.run(met, want=("SUEWS", "ESTM"), ts5=ts)
    dh = dict(d)
    dh.update({k.lower(): list(np.atleast_1d(v)) for k, v in estm.initial_state(ep, d["sfr"], d["emis"], ts[0]).items()})
    groups = ("SUEWS", "ESTM") if d["storageheatmethod"][0] == 4 else ("SUEWS",)
    rc, blks, state, _w = run_host(host, dh, met, groups, record=True, ts5=ts)
    assert rc == 0
    check_blocks(blks, ref, over)
    names = ("Tievolve", "Tground", "Twall", "Troof", "Tibld", "Tw_4", "T0_ibld", "TN_wall", "TN_roof", "lup_ground", "lup_wall",
             "lup_roof", "Tair2", "HVAC", "Tair2Set", "NewtonConverged", "state_id", "soilstore_id", "qn1_av", "dqndt")
    assert check_state(state, cell, names=names, where=over) == len(names)


@pytest.mark.parametrize("over", [{}, {"snowuse": 0, "netradiationmethod": 3, "emissionsmethod": 2},
                                  {"snowuse": 0, "netradiationmethod": 3, "emissionsmethod": 0, "ohmincqf": 1}],
                         ids=["baserun", "default-set", "incqf"])
def test_anohm_is_bit_identical(host, over):
    """StorageHeatMethod 3 over 2 days + 20 steps (three Levenberg-Marquardt / hybrid fits per surface): with the same libm
    on both sides even the fitted coefficients agree to the last bit -- the 1e-5 of the GPU tests is the math library's."""
    from suews_b200 import synth
    d = fx.state_init(2005)
    d.update({k: list(v) for k, v in synth.ANOHM_EXAMPLE.items()})
    d["storageheatmethod"] = [3]
    _with(d, over)
    met = fx.forcing(2005)[:288 * 2 + 20]
    cell = ol.Cell(d)
    ref = cell.run(met, want=("SUEWS", "DailyState"))
    rc, blks, state, _w = run_host(host, d, met, ("SUEWS", "DailyState"), record=True)
    assert rc == 0
    check_blocks(blks, ref, over)
    names = ("anohm_id_save", "anohm_coef_id_save", "anohm_coef_grid_save", "anohm_coeff_grid_day", "qn1_av", "dqndt", "state_id",
             "soilstore_id")
    assert check_state(state, cell, names=names, where=over) == len(names)
    assert (state["anohm_coeff_grid_day"] != -999.0).all()


@pytest.mark.parametrize("mode,et,isubs", [(2, 4, 0), (1, 1, 1), (2, 3, 0)])
def test_bluews_in_the_step_is_bit_identical(host, mode, et, isubs):
    """BLUEWS as the step's last stage (psih recomputed from the step's own z/L for InitialData_use 1): the BL line of three
    summer days and the slab variables."""
    L = ol.lib()
    L.so_set_cbl_initial.argtypes = [dp, C.c_int]
    L.so_set_bl_block.argtypes = [dp]
    params, ini = fx.cbl_inputs()
    ini = np.ascontiguousarray(ini)
    d = fx.state_init(2004)
    d.update(params)
    d.update(initialdata_use=[mode], entrainmenttype=[et], isubs=[isubs])
    met = fx.forcing(2004)[288 * 150:288 * 153]
    n = met.shape[0]
    bl = np.full((n, 22), -999.0, order="F")
    L.so_set_cbl_initial(ini.ctypes.data_as(dp), ini.shape[0])
    L.so_set_bl_block(bl.ctypes.data_as(dp))
    cell = ol.Cell(d)
    try:
        ref = cell.run(met, want=("SUEWS",))
    finally:
        L.so_set_bl_block(None)
        L.so_set_cbl_initial(None, 0)
    ref["BL"] = bl
    rc, blks, state, _w = run_host(host, d, met, ("SUEWS", "BL"), record=True, cbl_ini=ini)
    assert rc == 0
    check_blocks(blks, ref, (mode, et, isubs))
    names = ("cbl_blh_m", "cbl_tm_K", "cbl_qm_kgkg", "cbl_tp_K", "cbl_qp_kgkg", "cbl_gamt_Km", "cbl_gamq_kgkgm")
    assert check_state(state, cell, names=names) == len(names)
    assert bl[:, 5].max() > 400.0


@pytest.mark.parametrize("instance", [0, 2], ids=["generic", "snow-on-OHM"])
def test_accumulating_policies_over_the_year(host, instance):
    """What the bench launches: output policy P0 (seven flux sums per cell, no output line formed, the diagnostics no flux
    depends on skipped) and the 82-column running sums, over all of 2004.  The sums equal the oracle's lines added up in step
    order, bit for bit, and the state at the end is the state of the run that wrote every line."""
    d = fx.state_init(2004)
    met = fx.forcing(2004)
    cell = ol.Cell(d)
    ref = cell.run(met, want=("SUEWS",))["SUEWS"][:, 5:]
    seq = np.cumsum(ref, axis=0)[-1]                  # (cumsum adds strictly in order, as the accumulators do)
    acc7 = np.zeros(7)
    rc, _b, state, _w = run_host(host, d, met, (), instance=instance, kind=2, flux_only=True, acc=acc7)
    assert rc == 0
    assert np.array_equal(acc7, seq[PARITY_COLS]), (acc7, seq[PARITY_COLS])
    check_state(state, cell, where="P0, flux only")
    acc82 = np.zeros(82)
    rc, _b, state, _w = run_host(host, d, met, (), instance=instance, kind=1, acc=acc82)
    assert rc == 0
    strict = [c for c in range(82) if c + 5 not in RSL_COLS]
    assert np.all(same_bits(acc82[strict], seq[strict])), np.argwhere(~same_bits(acc82[strict], seq[strict])).ravel()
    loose = [c - 5 for c in RSL_COLS]
    assert np.allclose(acc82[loose], seq[loose], rtol=LOOSE_TOL, atol=0, equal_nan=True)
    check_state(state, cell, where="P0+")


def test_policy_p0_in_two_launches_equals_one(host):
    """A run cut in two launches (state, clock and sums handed over as the library does between advance calls) is the run."""
    d = fx.state_init(2005)
    met = fx.forcing(2005)[:1440]
    one = np.zeros(7)
    rc, _b, st1, _w = run_host(host, d, met, (), kind=2, flux_only=True, acc=one)
    assert rc == 0
    two = np.zeros(7)
    rc, _b, mid, _w = run_host(host, d, met[:500], (), kind=2, flux_only=True, acc=two)
    assert rc == 0
    d2 = dict(d)
    d2.update({k.lower(): list(v) for k, v in mid.items()})
    d2["dt_since_start"] = [d["dt_since_start"][0] + 500 * d["tstep"][0]]
    rc, _b, st2, _w = run_host(host, d2, met[500:], (), kind=2, flux_only=True, acc=two)
    assert rc == 0
    assert np.array_equal(one, two)
    for k in st1:
        assert np.all(same_bits(st1[k], st2[k])), k


@pytest.mark.parametrize("over,met_edit,own", [
    # no such storage-heat scheme: the reference runs on with an unset QS; the library refuses (1003) instead of inventing one
    (dict(storageheatmethod=2), None, 1003),
    (dict(netradiationmethod=7), None, None),                 # RadMethod's bare STOP (narp.f95:95-99)
    (dict(storageheatmethod=4), None, 1005),                  # ESTM without its surface-temperature forcing: the library's own check
    (dict(), ("Press_hPa", 12, 8000.0), None),                # forcing out of the accepted range
    (dict(), ("avU1", 9, -1.0), None),
    (dict(smdmethod=1), ("xsmd", 19, -999.0), None),          # observed soil moisture asked for, none given
], ids=["storage2", "netrad7", "estm-no-forcing", "pressure", "wind", "xsmd"])
def test_fatal_errors_are_the_oracles(host, over, met_edit, own):
    """Where the reference STOPs (ErrorHint's fatal codes and the driver's own checks) the step stops the cell with the code
    the oracle reports -- or, if the reference carries on, so do both."""
    d = _with(fx.state_init(2005), over)
    met = fx.forcing(2005)[:24].copy()
    if met_edit:
        met[6:, met_edit[1]] = met_edit[2]
    try:
        ol.Cell(d).run(met, want=("SUEWS",))
        code = 0
    except RuntimeError as e:
        code = int(str(e).split("error")[1].split(":")[0])
    rc, _b, _s, _w = run_host(host, d, met, ("SUEWS",), record=(over.get("storageheatmethod") == 4))
    assert rc == (own if own is not None else code), (rc, code)


def test_kernel_source_writes_the_reference_files(host):
    """No oracle in between: the host build of the kernel's step, the hourly aggregation and the library's text writer
    against the files the reference wrote -- the SUEWS and snow files of 2005 byte for byte, every printed field of the RSL
    and SOLWEIG files (bar SOLWEIG's unassigned night-time I0), the DailyState files of 2005 and of all 366 days of 2004."""
    from test_writer import _agg, _write_txt
    d = fx.state_init(2005)
    groups = ("SUEWS", "snow", "RSL", "DailyState", "SOLWEIG")
    rc, blks, _s, _w = run_host(host, d, fx.forcing(2005), groups)
    assert rc == 0
    gold = lambda f: open(os.path.join(fx.GOLDEN, f)).read()      # noqa: E731
    for g in ("SUEWS", "snow"):
        assert _write_txt(_agg(blks[g], g), g, outlevel=2) == gold("ref_2005_%s_60.txt" % g), g
    assert _write_txt(_agg(blks["RSL"], "RSL"), "RSL", outlevel=2).splitlines()[1:] == gold("ref_2005_RSL_60.txt").splitlines()[1:]
    ours, ref = _write_txt(_agg(blks["SOLWEIG"], "SOLWEIG"), "SOLWEIG", outlevel=2).splitlines(), gold("ref_2005_SOLWEIG_60.txt").splitlines()
    i0 = ref[0].split().index("I0")
    assert ours[0] == ref[0] and len(ours) == len(ref)
    assert all(x == y for a, b in zip(ours[1:], ref[1:]) for c, (x, y) in enumerate(zip(a.split(), b.split())) if c != i0)
    rows = blks["DailyState"]
    assert _write_txt(rows[rows[:, 5] != -999], "DailyState", outlevel=2) == gold("ref_2005_DailyState.txt")
    rc, blks, _s, _w = run_host(host, fx.state_init(2004), fx.forcing(2004), ("DailyState",))
    assert rc == 0
    rows = blks["DailyState"]
    assert _write_txt(rows[rows[:, 5] != -999], "DailyState", outlevel=2) == gold("ref_2004_DailyState.txt")


@pytest.mark.parametrize("year,fname", [(2005, "ref_InitialConditions_2005_EndofRun.nml"), (2004, "ref_InitialConditions_2005.nml")])
def test_kernel_source_ends_on_the_reference_checkpoint(host, year, fname):
    """The InitialConditions namelist the reference wrote at the end of each leg of its golden run holds the exact doubles of
    54 state variables (17 significant digits): the host build of the kernel's step, run under the bench's policy P0, ends
    on every one of them bit for bit, and the library's writer turns that state into the reference's file byte for byte."""
    from suews_b200 import checkpoint
    acc = np.zeros(7)
    rc, _b, state, _w = run_host(host, fx.state_init(year), fx.forcing(year), (), kind=2, flux_only=True, acc=acc,
                                 instance=2)
    assert rc == 0
    ref = checkpoint.read(os.path.join(fx.GOLDEN, fname))
    text = checkpoint.write_text(checkpoint.state_vector(lambda n: state[next(k for k in state if k.lower() == n.lower())]), snow_use=1)
    ours = checkpoint.read_text(text)
    assert ours.keys() == ref.keys() and len(ref) == 54
    bad = [(k, ours[k], r) for k, r in ref.items() if ours[k] != r]
    assert not bad, bad
    assert text == open(os.path.join(fx.GOLDEN, fname)).read()


FUZZ_CHOICES = dict(netradiationmethod=[0, 1, 2, 3, 11, 12, 13, 100, 200, 300], stabilitymethod=[1, 2, 3, 4, 5, 6, 7],
                    emissionsmethod=[0, 1, 2, 3, 4, 5, 6, 11, 14, 41], snowuse=[0, 1], roughlenheatmethod=[1, 2, 3, 4, 5],
                    roughlenmommethod=[1, 2, 3], smdmethod=[0, 0, 1], waterusemethod=[0, 1], evapmethod=[1, 2], gsmodel=[1, 2],
                    laicalcyes=[0, 1], storageheatmethod=[0, 1, 1], ohmincqf=[0, 1], basetmethod=[1, 2],
                    aerodynamicresistancemethod=[1, 2, 2, 3])


# class parameters the generator leaves alone: the timestep and the tables whose rows must keep their sums
FUZZ_FIXED = {"tstep", "ohm_coef", "waterdist", "ahprof_24hr", "humactivity_24hr", "popprof_24hr", "traffprof_24hr", "wuprofa_24hr",
              "wuprofm_24hr", "snowprof_24hr"}


EXTRA_GROUPS = ("RSL", "DailyState", "SOLWEIG")


def fuzz_trial(host, rng, snaps, met_all):
    """One random case: a state of the year, a random combination of scheme switches, scaled class parameters, a random cell
    (fractions with absent surfaces, heights, site) and shifted forcing; -> None, or a description of what differs from the oracle."""
    step, d0 = snaps[rng.integers(len(snaps))]
    d = dict(d0)
    for k in INTERNAL:
        d.pop(k, None)                                         # the library's memos restart, like the oracle's SAVE'd ones
    over = {k: int(rng.choice(v)) for k, v in FUZZ_CHOICES.items() if rng.random() < 0.5}
    _with(d, over)
    if rng.random() < 0.5:                                     # six class parameters (albedos, capacities, conductances, ...) scaled
        names = [nm.lower() for nm, kind, _n in all_fields() if kind == 1 and nm.lower() in d and nm.lower() not in FUZZ_FIXED]
        for nm in rng.choice(names, size=6, replace=False):
            v = np.asarray(d[nm], dtype=np.float64)
            d[nm] = list(v * rng.uniform(0.6, 1.4, size=v.shape))
    if rng.random() < 0.8:
        u = rng.random(7) * (rng.random(7) < 0.8)
        u[1] = max(u[1], 0.02)
        sfr = u / u.sum()
        sfr[6] = max(0.0, 1 - sfr[:6].sum())
        d["sfr"] = list(sfr / sfr.sum())
    if rng.random() < 0.5:
        d["bldgh"], d["evetreeh"], d["dectreeh"] = [float(rng.uniform(3, 40))], [float(rng.uniform(2, 25))], [float(rng.uniform(2, 25))]
        # (above 10 m: with no profile level over 10 m the reference's interp_z indexes element 0 of a 1-based array,
        # rslprof.f95:331-342 -- U10 is then whatever lies in front of the array; the library keeps the index in range)
        d["z"] = [float(max(12.0, max(d["bldgh"][0], d["evetreeh"][0], d["dectreeh"][0]) * rng.uniform(1.5, 4)))]
    if rng.random() < 0.5:
        d["lat"], d["lng"] = [float(rng.uniform(-60, 70))], [float(rng.uniform(-170, 170))]
    off = int(rng.integers(0, 288)) if rng.random() < 0.3 else 0
    met = met_all[step + off:step + off + 288 * int(rng.integers(1, 3))].copy()
    d["dt_since_start"] = [d["dt_since_start"][0] + off * 300]
    if over.get("smdmethod", 0) > 0:
        met[:, 19] = float(rng.uniform(5, 80))
    if rng.random() < 0.4:
        met[:, 11] += float(rng.uniform(-8, 8))
    if rng.random() < 0.3:
        met[:, 13] *= float(rng.uniform(0, 5))
    if rng.random() < 0.3:
        met[:, 9] = np.maximum(0.05, met[:, 9] * float(rng.uniform(0.1, 3)))
    groups = (("SUEWS", "snow") if d["snowuse"][0] == 1 else ("SUEWS",)) + EXTRA_GROUPS
    # one case in three keeps a storage record: AnOHM, ESTM (random variant) or BLUEWS on top of the storage scheme drawn
    record, ts, ini, dh = False, None, None, d
    kind = rng.random()
    if kind < 0.1:
        from suews_b200 import synth
        d.update({k: list(v) for k, v in synth.ANOHM_EXAMPLE.items()})
        d["storageheatmethod"], record = [3], True
    elif kind < 0.2:
        from suews_b200 import estm
        d["storageheatmethod"] = [int(rng.choice([4, 14]))]
        ep = fx.estm_example(d["surfacearea"][0], **ESTM_VARIANTS[rng.integers(5)])
        d.update({k.lower(): v for k, v in ep.items()})
        ts, record = fx.estm_ts_forcing(met), True
        dh = dict(d)
        dh.update({k.lower(): list(np.atleast_1d(v)) for k, v in estm.initial_state(ep, d["sfr"], d["emis"], ts[0]).items()})
        if d["storageheatmethod"][0] == 4:
            groups = groups + ("ESTM",)
    elif kind < 0.33:
        params, ini = fx.cbl_inputs()
        d.update(params)
        d.update(initialdata_use=[int(rng.choice([1, 2]))], entrainmenttype=[int(rng.integers(1, 5))], isubs=[int(rng.integers(0, 2))],
                 qh_choice=[int(rng.integers(1, 3))])
        ini, record = np.ascontiguousarray(ini), True
    L = ol.lib()
    bl = np.full((met.shape[0], 22), -999.0, order="F")
    try:
        if ini is not None:
            L.so_set_cbl_initial.argtypes = [dp, C.c_int]
            L.so_set_bl_block.argtypes = [dp]
            L.so_set_cbl_initial(ini.ctypes.data_as(dp), ini.shape[0])
            L.so_set_bl_block(bl.ctypes.data_as(dp))
        cell = ol.Cell(d)
        if ts is not None:
            cell.estm_start(ts[0])
        ref = cell.run(met, want=groups, ts5=ts)
        code = 0
    except RuntimeError as e:
        code = int(str(e).split("error")[1].split(":")[0])
    finally:
        if ini is not None:
            L.so_set_bl_block(None)
            L.so_set_cbl_initial(None, 0)
    if ini is not None and code == 0:
        ref["BL"] = bl
        groups = groups + ("BL",)
    # the generic instance or the specialisation this class would be launched with
    special = (1 if d["snowuse"][0] == 0 else 2) if (d["storageheatmethod"][0] in (0, 1) and not record) else 0
    instance = special if rng.random() < 0.5 else 0
    rc, blks, state, warn = run_host(host, dh, met, groups, instance=instance, record=record, ts5=ts, cbl_ini=ini)
    d = dh
    if rc != code:
        # the library's own refusals where the reference runs on with an unset value: an unknown storage scheme (QS unset),
        # BLUEWS' diagnosed start without the psih AerodynamicResistanceMethod 2 leaves behind
        own = (rc == 1003 and code == 0) or (rc == 1007 and code == 0 and ini is not None and d["initialdata_use"][0] == 1
                                             and d["aerodynamicresistancemethod"][0] != 2)
        return None if own else ("fatal code", over, rc, code)
    if code:
        return None
    try:
        check_blocks(blks, ref, over)
        check_state(state, cell, where=over)
        check_warnings(warn, ref, over)
        if rng.random() < 0.3:                                   # the same case under policy P0, flux only
            acc = np.zeros(7)
            rc, _b, state0, _w = run_host(host, d, met, (), instance=instance, kind=2, flux_only=True, acc=acc, record=record,
                                          ts5=ts, cbl_ini=ini)
            seq = np.cumsum(ref["SUEWS"][:, 5:], axis=0)[-1][PARITY_COLS]
            assert rc == 0 and np.all(same_bits(acc, seq)), ("P0 sums", over, acc, seq)
            check_state(state0, cell, where=("P0", over))
    except AssertionError as e:
        return str(e)[:600]
    return None


def year_snapshots(host):
    """the BaseRun state at the start of 2004 and after every 30 days of it (the host build's, which is the oracle's)"""
    met_all = fx.forcing(2004)
    base = fx.state_init(2004)
    snaps, d, step = [(0, dict(base))], dict(base), 0
    for _k in range(11):
        rc, _b, st, _w = run_host(host, d, met_all[step:step + 288 * 30], ())
        assert rc == 0
        step += 288 * 30
        d = dict(d)
        d.update({k.lower(): list(v) for k, v in st.items()})
        d["dt_since_start"] = [base["dt_since_start"][0] + step * 300]
        snaps.append((step, dict(d)))
    return snaps, met_all


def test_random_scheme_cell_and_weather_combinations(host):
    """150 random cases (tools/fuzz_step_host.py runs the same generator for as long as one likes: 10 000 cases without a
    difference when this was written): states from every month of 2004, scheme switches combined at random -- 10 radiation,
    7 stability, 10 emission methods, snow on / off, ... --, cells with absent surfaces, other heights and sites, warmer /
    colder / wetter / windier forcing, one in three with AnOHM, ESTM or BLUEWS, the generic or the specialised instance, some again
    under policy P0.  Every output value, the end state, the warnings and the fatal codes are the oracle's."""
    snaps, met_all = year_snapshots(host)
    rng = np.random.default_rng(20261020)
    bad = [r for r in (fuzz_trial(host, rng, snaps, met_all) for _ in range(150)) if r]
    assert not bad, bad[:3]


def test_estm_from_the_run_directory_with_the_reference_s_own_forcing(host, tmp_path):
    """StorageHeatMethod 4 end to end on the host: the golden run directory with ESTM switched on (table join, ESTM class
    weighting, ESTMinput.nml), the reference's own surface-temperature file of 2005 (its 5-min disaggregation as the reference
    wrote it), the 2005 checkpoint as InitialConditions -- five days through the host build of the step against the oracle."""
    import shutil
    from suews_b200 import tables
    src = os.path.join(fx.GOLDEN, "baserun_2020b")
    run = tmp_path / "run"
    shutil.copytree(src, run)
    rcp = run / "RunControl.nml"
    rcp.write_text(rcp.read_text().replace("StorageHeatMethod=1", "StorageHeatMethod=4"))
    ts5 = np.loadtxt(os.path.join(fx.GOLDEN, "ref_estm_ts_2005_5min.txt"), skiprows=1)[:1440]
    met = fx.forcing(2005)
    assert ts5.shape == (1440, 13) and np.array_equal(ts5[:, :4], met[:, :4])
    ic = os.path.join(fx.GOLDEN, "ref_InitialConditions_2005.nml")
    a0, _aux, _rc = tables.run_arguments(str(run), 1, 2005, initial_conditions=ic)
    a1, _aux, _rc = tables.run_arguments(str(run), 1, 2005, initial_conditions=ic, estm_ts_row0=ts5[0])
    assert a0["storageheatmethod"] == [4] and "tw_4" in a1 and "tw_4" not in a0
    cell = ol.Cell(a0)
    cell.estm_start(ts5[0])
    ref = cell.run(met, want=("SUEWS", "ESTM"), ts5=ts5)
    rc, blks, state, _w = run_host(host, a1, met, ("SUEWS", "ESTM"), record=True, ts5=ts5)
    assert rc == 0
    check_blocks(blks, ref, "ESTM from the run directory")
    names = ("Tievolve", "Tground", "Twall", "Troof", "Tibld", "Tw_4", "T0_ibld", "TN_wall", "TN_roof", "Tair2", "state_id", "soilstore_id")
    assert check_state(state, cell, names=names) == len(names)
    qs = blks["SUEWS"][:, 5 + 7]
    assert np.isfinite(qs).all() and qs.max() - qs.min() > 10.0          # a storage heat flux with a diurnal course


def test_anohm_from_the_run_directory(host, tmp_path):
    """StorageHeatMethod 3 with the AnOHM columns of the run directory's own surface tables (SUEWS_NonVeg / Veg / Water)."""
    import shutil
    from suews_b200 import tables
    run = tmp_path / "run"
    shutil.copytree(os.path.join(fx.GOLDEN, "baserun_2020b"), run)
    rcp = run / "RunControl.nml"
    rcp.write_text(rcp.read_text().replace("StorageHeatMethod=1", "StorageHeatMethod=3"))
    a, _aux, _rc = tables.run_arguments(str(run), 1, 2004)
    assert a["storageheatmethod"] == [3] and min(a["cpanohm"]) > 0 and min(a["kkanohm"]) > 0 and min(a["chanohm"]) > 0
    met = fx.forcing(2004)[288 * 170:288 * 172 + 30]                # midsummer: long days, a fit on every surface
    a["dt_since_start"] = [288 * 170 * 300]
    cell = ol.Cell(a)
    ref = cell.run(met, want=("SUEWS", "DailyState"))
    rc, blks, state, _w = run_host(host, a, met, ("SUEWS", "DailyState"), record=True)
    assert rc == 0
    check_blocks(blks, ref, "AnOHM from the run directory")
    assert check_state(state, cell, names=("anohm_coeff_grid_day", "anohm_id_save", "qn1_av", "dqndt")) == 4
    assert (state["anohm_coeff_grid_day"] != -999.0).all()
