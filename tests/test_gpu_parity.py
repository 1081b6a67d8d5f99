"""CUDA path vs the CPU oracle through the C ABI (pytest -m gpu; runs on the B200 box).

Tolerances (BASELINE.json north_star): integer/flag outputs bit-exact; floating-point outputs within
1e-10 relative per output step over the first simulated day, 1e-6 relative on flux sums.

"Relative" needs a scale for values that cross zero.  It is max(|ref|, S_col) with S_col the size of the
quantities the column is computed from: the column's own maximum over the run for fluxes and stores;
273.15 for temperatures reported in deg C (they are Kelvin values minus 273.16: Tsurf, Ts, T2, the RSL
T profile, Tsnow); and rho*cp*273/RA ~ 1e4 W m-2 for QHresis = rho*cp*(Tsurf - Tair)/RA, a difference
of two ~273 K numbers.  The only source of CUDA-vs-oracle differences is the last-ulp behaviour of
exp/log/pow/sin/cos in the two math libraries (the arithmetic is contraction-free on both sides).
"""
import os

import numpy as np
import pytest

import fixtures as fx
import oracle_lib as ol

pytestmark = pytest.mark.gpu

TOL_DAY1 = 1e-10
TOL_SUM = 1e-6
PARITY_COLS = {"QN": 5, "QF": 6, "QS": 7, "QH": 8, "QE": 9, "RO": 16, "SWE": 71}


TEMP_COLS = {"SUEWS": [4, 77, 78], "RSL": list(range(60, 90)), "snow": list(range(90, 97)), "DailyState": [12, 13],
             "SOLWEIG": [17, 24, 25]}
BIG_COLS = {"SUEWS": {12: 1.0e4}}


def rel_err(out, ref, group="SUEWS"):
    """out, ref: [nstep, ncol] without the 5 datetime columns"""
    colmax = np.nanmax(np.abs(np.where(np.isfinite(ref), ref, 0.0)), axis=0)
    floor = colmax + 1e-300
    for c in TEMP_COLS.get(group, []):
        if c < floor.size:
            floor[c] = max(floor[c], 273.15)
    for c, v in BIG_COLS.get(group, {}).items():
        floor[c] = max(floor[c], v)
    scale = np.maximum(np.abs(ref), floor[None, :])
    same_nonfinite = (~np.isfinite(ref)) & (~np.isfinite(out)) & ((np.isnan(ref) == np.isnan(out)))
    e = np.abs(out - ref) / scale
    e[same_nonfinite] = 0.0
    return e


# The second measure: the plain per-value relative error |out - ref| / max(|ref|, floor) with an explicit absolute floor.
# The floor says below which magnitude a value is compared absolutely (at floor * tol): 1e-3 in the column's unit for
# fluxes, stores and everything else; for the columns that are differences of two ~273 K / ~1e4 W m-2 numbers (deg C
# temperatures, QHresis) the last-ulp noise of the operands is ~5e-14 / ~2e-12 absolute whatever the value, so their
# floor is 1 (deg C, W m-2).
VALUE_FLOOR = 1e-3
VALUE_FLOOR_CANCEL = 1.0


def rel_err_value(out, ref, group="SUEWS"):
    """out, ref: [nstep, ncol] without the 5 datetime columns -> per-value relative error with the absolute floors above"""
    floor = np.full(ref.shape[1], VALUE_FLOOR)
    for c in TEMP_COLS.get(group, []) + list(BIG_COLS.get(group, {}).keys()):
        if c < floor.size:
            floor[c] = VALUE_FLOOR_CANCEL
    scale = np.maximum(np.abs(ref), floor[None, :])
    same_nonfinite = (~np.isfinite(ref)) & (~np.isfinite(out)) & ((np.isnan(ref) == np.isnan(out)))
    with np.errstate(invalid="ignore"):
        e = np.abs(out - ref) / scale
    e[same_nonfinite] = 0.0
    return e


@pytest.fixture(scope="module")
def gpu_and_ref_2005():
    from suews_b200 import OUT_FULL, OUT_SUM82, SuewsBatch
    d = fx.state_init(2005)
    met = fx.forcing(2005)
    cell = ol.Cell(d)
    ref = cell.run(met, want=("SUEWS", "snow", "RSL", "DailyState", "SOLWEIG"))
    b = SuewsBatch(1, 1)
    assert b.load_reference_args(d) == []
    b.upload()
    b.set_forcing(met)
    b.advance(0, met.shape[0], OUT_FULL | OUT_SUM82)
    yield b, ref, cell, met
    b.close()


@pytest.mark.parametrize("group", ["SUEWS", "snow", "RSL", "DailyState", "SOLWEIG"])
def test_blocks_match_oracle(gpu_and_ref_2005, group):
    b, ref, _cell, met = gpu_and_ref_2005
    out = b.block(group, 0, met.shape[0])
    r = ref[group]
    assert np.array_equal(out[:, :5], r[:, :5])                 # datetime columns: exact
    assert np.array_equal(out == -999.0, r == -999.0)           # sentinels: exact
    e = rel_err(out[:, 5:], r[:, 5:], group)
    assert e[:288].max() <= TOL_DAY1, (group, float(e[:288].max()), np.unravel_index(e[:288].argmax(), e[:288].shape))
    assert e.max() <= 1e-8, (group, float(e.max()))
    ev = rel_err_value(out[:, 5:], r[:, 5:], group)      # per value, not per column scale
    print("%s: worst error relative to the column scale %.3e, relative to the value itself %.3e (day 1)" % (group, e[:288].max(), ev[:288].max()))
    assert ev[:288].max() <= TOL_DAY1, (group, float(ev[:288].max()), np.unravel_index(ev[:288].argmax(), ev[:288].shape))
    if group == "RSL":
        assert np.array_equal(out[:, -1], r[:, -1])             # flag_RSL
    if group == "DailyState":
        assert np.array_equal(out[:, 10], r[:, 10])             # DaysSR


def test_flux_sums_and_state(gpu_and_ref_2005):
    b, ref, cell, met = gpu_and_ref_2005
    sums = b.sums()[0]
    rs = ref["SUEWS"][:, 5:].sum(axis=0)
    for nm, c in PARITY_COLS.items():
        assert abs(sums[c] - rs[c]) <= TOL_SUM * max(abs(rs[c]), 1e-30), (nm, sums[c], rs[c])
    sd = cell.state_dict()
    for nm, v in sd.items():
        if nm in ("dt_since_start", "tstep_prev", "lv_switch1", "lv_switch2"):
            continue
        got = b.get_state(nm)[0]
        assert np.allclose(got, v, rtol=1e-9, atol=1e-12), nm
    assert b.get_clock() == (int(sd["dt_since_start"][0]), 300)
    sw = int(b.get_state("lv_switch")[0, 0])
    assert sw == int(sd["lv_switch1"][0]) + 2 * int(sd["lv_switch2"][0])


def test_state_after_the_2005_leg_against_the_reference_checkpoint(gpu_and_ref_2005):
    """the reference's own end-of-run InitialConditions namelist (17 significant digits) after the same 1440 steps"""
    from suews_b200 import checkpoint
    b = gpu_and_ref_2005[0]
    ref_nml = checkpoint.read(fx.GOLDEN + "/ref_InitialConditions_2005_EndofRun.nml")
    ours = checkpoint.read_text(checkpoint.write_text(checkpoint.state_vector(lambda n: b.get_state(n)[0]), snow_use=1))
    assert ours.keys() == ref_nml.keys()
    worst = max(abs(ours[k] - r) / max(abs(r), 1e-3) for k, r in ref_nml.items())
    assert worst <= 1e-12, worst        # measured: 0 (all 54 values bit-identical)


def test_parity_sums_policy_equals_full_sums(gpu_and_ref_2005):
    """output policy P0 (seven parity fluxes) accumulates exactly what the 82-column policy holds in those columns"""
    from suews_b200 import OUT_SUM7, SuewsBatch
    b1, _ref, _cell, met = gpu_and_ref_2005
    b = SuewsBatch(1, 1)
    b.load_reference_args(fx.state_init(2005))
    b.upload()
    b.set_forcing(met)
    b.advance(0, met.shape[0], OUT_SUM7)
    cols = list(PARITY_COLS.values())
    assert np.array_equal(b.sums()[:, cols], b1.sums()[:, cols])
    other = [c for c in range(82) if c not in cols]
    assert np.all(b.sums()[:, other] == 0.0)
    b.close()


def test_chunked_advance_equals_single_launch(gpu_and_ref_2005):
    """state carried across launches == one long launch (the persistent-state contract)"""
    from suews_b200 import OUT_SUM82, SuewsBatch
    b1, _ref, _cell, met = gpu_and_ref_2005
    d = fx.state_init(2005)
    b = SuewsBatch(1, 1)
    b.load_reference_args(d)
    b.upload()
    b.set_forcing(met)
    for s0 in range(0, 1440, 97):
        b.advance(s0, min(97, 1440 - s0), OUT_SUM82)
    for nm in ("soilstore_id", "state_id", "SnowPack", "HDD_id", "qn1_av", "Tair_av", "LAI_id"):
        assert np.array_equal(b.get_state(nm), b1.get_state(nm)), nm
    assert np.array_equal(b.sums(), b1.sums())
    b.close()


BIOCO2 = dict(theta_bioco2=[0.96] * 3, alpha_enh_bioco2=[0.016] * 3, beta_enh_bioco2=[33.454] * 3)


@pytest.mark.parametrize("scheme_over", [
    dict(snowuse=0, netradiationmethod=3, emissionsmethod=2),                 # user-facing default set
    dict(netradiationmethod=1), dict(netradiationmethod=2), dict(netradiationmethod=3), dict(netradiationmethod=11),
    dict(netradiationmethod=100), dict(netradiationmethod=300), dict(netradiationmethod=0),
    dict(netradiationmethod=0, snowuse=0), dict(netradiationmethod=12), dict(netradiationmethod=200),
    dict(stabilitymethod=2), dict(stabilitymethod=4),
    # the stability functions BTS_config.nml does not list (atmmoiststab.f95:252-497 has them all)
    dict(stabilitymethod=1), dict(stabilitymethod=5), dict(stabilitymethod=6), dict(stabilitymethod=7),
    dict(roughlenheatmethod=1), dict(roughlenheatmethod=3), dict(roughlenheatmethod=4), dict(roughlenheatmethod=5),
    dict(roughlenmommethod=1), dict(roughlenmommethod=3),
    dict(emissionsmethod=0), dict(emissionsmethod=2), dict(emissionsmethod=3), dict(emissionsmethod=4),
    dict(emissionsmethod=5), dict(emissionsmethod=6),
    dict(storageheatmethod=0), dict(ohmincqf=1), dict(evapmethod=1), dict(aerodynamicresistancemethod=1),
    dict(aerodynamicresistancemethod=3), dict(gsmodel=1), dict(smdmethod=1), dict(waterusemethod=1),
    dict(basetmethod=2), dict(laicalcyes=0),
    # biogenic CO2 families (biogenco2.f95): rectangular / non-rectangular / bulk hyperbola, conductance based
    dict(emissionsmethod=11), dict(emissionsmethod=14), dict(emissionsmethod=22, **BIOCO2), dict(emissionsmethod=33, **BIOCO2),
    dict(emissionsmethod=36, theta_bioco2=[0.9, 0.96, 0.8], alpha_enh_bioco2=[0.016] * 3, beta_enh_bioco2=[33.454] * 3),
    dict(emissionsmethod=41), dict(emissionsmethod=45, gsmodel=4), dict(emissionsmethod=12, gsmodel=3),
])
def test_scheme_options_match_oracle(scheme_over):
    """Every option of the reference's physics sweep (Test/code/BTS_config.nml &physics_test) that is
    inside this library: CUDA == oracle on the first day."""
    from suews_b200 import OUT_FULL, SuewsBatch
    d = fx.state_init(2005)
    for k, v in scheme_over.items():
        d[k] = list(v) if isinstance(v, (list, tuple)) else [v]
    if scheme_over.get("snowuse", 1) == 0:
        for k in ("snowpack", "snowfrac", "snowwater"):
            d[k] = [0.0] * 7
    met = fx.forcing(2005)[:288].copy()
    if scheme_over.get("smdmethod", 0) > 0:
        met[:, 19] = 30.0
    ref = ol.Cell(d).run(met, want=("SUEWS",))["SUEWS"]
    b = SuewsBatch(1, 1)
    b.load_reference_args(d)
    b.upload()
    b.set_forcing(met)
    b.advance(0, 288, OUT_FULL)
    out = b.block("SUEWS", 0, 288)
    b.close()
    assert np.array_equal(out[:, :5], ref[:, :5])
    e = rel_err(out[:, 5:], ref[:, 5:])
    assert e.max() <= TOL_DAY1, (scheme_over, float(e.max()), np.unravel_index(e.argmax(), e.shape))


@pytest.mark.parametrize("scheme_over", [
    dict(), dict(snowuse=0), dict(netradiationmethod=3), dict(netradiationmethod=1), dict(netradiationmethod=11),
    dict(netradiationmethod=100), dict(netradiationmethod=0), dict(stabilitymethod=2), dict(stabilitymethod=4),
    dict(roughlenheatmethod=1), dict(roughlenmommethod=1), dict(emissionsmethod=2), dict(evapmethod=1),
    dict(storageheatmethod=0), dict(smdmethod=1), dict(waterusemethod=1), dict(aerodynamicresistancemethod=1),
])
def test_scheme_options_winter_start(scheme_over):
    """The same option sweep from the 2004 dump: sub-zero air, snow on the ground, freezing surface water and
    rain in the first days -- the snow/freeze branches the mild 2005 start never enters.  Three days."""
    from suews_b200 import OUT_FULL, SuewsBatch
    d = fx.state_init(2004)
    for k, v in scheme_over.items():
        d[k] = [v]
    if scheme_over.get("snowuse", 1) == 0:
        for k in ("snowpack", "snowfrac", "snowwater"):
            d[k] = [0.0] * 7
    n = 3 * 288
    met = fx.forcing(2004)[:n].copy()
    if scheme_over.get("smdmethod", 0) > 0:
        met[:, 19] = 30.0
    ref = ol.Cell(d).run(met, want=("SUEWS", "snow"))
    b = SuewsBatch(1, 1)
    b.load_reference_args(d)
    b.upload()
    b.set_forcing(met)
    b.advance(0, n, OUT_FULL)
    groups = ("SUEWS", "snow") if d["snowuse"][0] == 1 else ("SUEWS",)
    for g in groups:
        out = b.block(g, 0, n)
        assert np.array_equal(out[:, :5], ref[g][:, :5])
        e = rel_err(out[:, 5:], ref[g][:, 5:], g)
        assert e[:288].max() <= TOL_DAY1, (scheme_over, g, float(e[:288].max()), np.unravel_index(e[:288].argmax(), e[:288].shape))
        # beyond day 1 a last-ulp libm difference can flip the Tsurf loop's |dT| > 0.1 test for one step
        # (one trip more or less); the north_star bounds those days through the flux sums
        assert e.max() <= 1e-6, (scheme_over, g, float(e.max()), np.unravel_index(e.argmax(), e.shape))
        if g == "SUEWS":
            so, sr = out[:, 5:].sum(axis=0), ref[g][:, 5:].sum(axis=0)
            for nm, c in PARITY_COLS.items():
                assert abs(so[c] - sr[c]) <= TOL_SUM * max(abs(sr[c]), 1e-30), (scheme_over, nm)
    b.close()


def test_many_cells_match_per_cell_oracle():
    """Synthetic multi-cell batch (perturbed fractions, location, morphology, forcing offsets): every
    cell equals its own single-cell oracle run -- the reference's multi-grid == single-grid property
    (Test/code/Test_SUEWS.py:384) at 1e-10 instead of 4 decimals."""
    from suews_b200 import OUT_SUM82, SuewsBatch, synth
    base = fx.state_init(2005)
    met = fx.forcing(2005)[:288]
    n = 96
    b = SuewsBatch(n, 1)
    _missing, site = synth.build_batch(b, base, cell0=1000, scheme="baserun")
    b.upload()
    b.set_forcing(met)
    b.advance(0, 288, OUT_SUM82)
    sums = b.sums()
    fatal, _warn, _it = b.diag()
    assert (fatal == 0).all()
    d0 = synth.base_with_scheme(base, "baserun")
    worst = 0.0
    for c in range(0, n, 5):
        dd = dict(d0)
        for k in ("sfr", "lat", "lng", "bldgH", "EveTreeH", "DecTreeH", "Z"):
            dd[k.lower()] = np.atleast_1d(site[k][c])
        m = met.copy()
        m[:, 11] += site["dTemp_C"][c]
        m[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), m[:, 10] + site["dRH"][c]))
        ref = ol.Cell(dd).run(m, want=("SUEWS",))["SUEWS"][:, 5:].sum(axis=0)
        for nm, col in PARITY_COLS.items():
            err = abs(sums[c, col] - ref[col]) / max(abs(ref[col]), 1e-30)
            worst = max(worst, err)
            assert err <= TOL_SUM, (c, nm, sums[c, col], ref[col])
    b.close()
    print("worst flux-sum rel err over sampled cells: %.3e" % worst)


def test_non_converging_tsurf_loop_is_exact():
    """Stable winter nights put the Tsurf loop (driver.f95:642) into an exact 2-cycle and the reference runs all
    99 trips; the kernel detects the cycle and jumps to the state trip 99 would leave.  Results must be the
    99-trip results: compare full output blocks on steps where the oracle really used 99 trips."""
    from suews_b200 import OUT_FULL, SuewsBatch, synth
    base = fx.state_init(2005)
    met = fx.forcing(2004)[:60]
    n = 8
    b = SuewsBatch(n, 1)
    _missing, site = synth.build_batch(b, base, cell0=0, scheme="default")
    b.upload()
    b.set_forcing(met)
    b.advance(0, 60, OUT_FULL)
    _fatal, _warn, iters = b.diag()
    d0 = synth.base_with_scheme(base, "default")
    n99 = 0
    for c in (0, 1, 5):
        dd = dict(d0)
        for k in ("sfr", "lat", "lng", "bldgH", "EveTreeH", "DecTreeH", "Z"):
            dd[k.lower()] = np.atleast_1d(site[k][c])
        m = met.copy()
        m[:, 11] += site["dTemp_C"][c]
        m[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), m[:, 10] + site["dRH"][c]))
        cell = ol.Cell(dd)
        trips = 0
        ref = {g: np.zeros((60, ol.NCOL[g]), order="F") for g in ("SUEWS", "RSL")}
        for i in range(60):     # one step at a time to read the oracle's trip count of every step
            r = cell.run(m[i:i + 1], want=("SUEWS", "RSL"))
            nit = cell.n_iter_last
            trips += nit
            n99 += int(nit == 99)
            for g in ref:
                ref[g][i] = r[g][0]
        assert int(iters[c]) == trips, (c, int(iters[c]), trips)      # logical trip count == the reference's
        for g in ref:
            out = b.block(g, c, 60)
            assert np.array_equal(out == -999.0, ref[g] == -999.0)
            e = rel_err(out[:, 5:], ref[g][:, 5:], g)
            assert e.max() <= TOL_DAY1, (c, g, float(e.max()), np.unravel_index(e.argmax(), e.shape))
    assert n99 >= 20, n99      # the case under test really occurs in this window
    b.close()


def test_annual_flux_sums_match_oracle():
    """north_star: 1e-6 relative on annual flux sums (QN, QF, QS, QH, QE, runoff, SWE).  Full leap year 2004
    (105 408 five-minute steps) from the 2004 state_init dump, snow on; plus the end-of-year state."""
    from suews_b200 import OUT_SUM7, SuewsBatch
    d = fx.state_init(2004)
    met = fx.forcing(2004)
    assert met.shape[0] == 366 * 288
    cell = ol.Cell(d)
    ref = cell.run(met, want=("SUEWS",))["SUEWS"][:, 5:]
    b = SuewsBatch(1, 1)
    assert b.load_reference_args(d) == []
    b.upload()
    b.set_forcing(met)
    for day0 in range(0, 366, 61):                      # six launches of 61 days: state carried on the device
        b.advance(day0 * 288, 61 * 288, OUT_SUM7)
    sums = b.sums()[0]
    rs = ref.sum(axis=0)
    for nm, c in PARITY_COLS.items():
        assert abs(sums[c] - rs[c]) <= TOL_SUM * max(abs(rs[c]), 1e-30), (nm, sums[c], rs[c])
    fatal, _w, iters = b.diag()
    assert fatal[0] == 0
    for nm, v in cell.state_dict().items():
        if nm in ("dt_since_start", "tstep_prev", "lv_switch1", "lv_switch2"):
            continue
        assert np.allclose(b.get_state(nm)[0], v, rtol=1e-7, atol=1e-9), nm
    # ... and against the REFERENCE itself: the InitialConditions namelist it wrote at the end of its own 2004 leg holds
    # the exact doubles of 54 state variables (tests/test_oracle_golden.py: the oracle equals them bit for bit)
    from suews_b200 import checkpoint
    ref_nml = checkpoint.read(fx.GOLDEN + "/ref_InitialConditions_2005.nml")
    ours = checkpoint.read_text(checkpoint.write_text(checkpoint.state_vector(lambda n: b.get_state(n)[0]), snow_use=1))
    assert ours.keys() == ref_nml.keys()
    for k, r in ref_nml.items():   # measured: 53 of the 54 values bit-identical, the other one ulp off (1.4e-16)
        assert abs(ours[k] - r) <= 1e-10 * max(abs(r), 1e-3), (k, ours[k], r)
    b.close()


def test_replicated_cells_equal_single_cell():
    """BASELINE configs[1]: the BaseRun cell replicated to 10 000 grids over the whole 2005 leg of the golden run -- the
    reference's multi-grid == single-grid property (Test/code/Test_SUEWS.py:384), here bit for bit, and every grid ends
    on the checkpoint the reference wrote after the same 1440 steps"""
    from suews_b200 import OUT_SUM82, SuewsBatch, checkpoint
    d = fx.state_init(2005)
    met = fx.forcing(2005)
    n = 10000
    b = SuewsBatch(n, 1)
    b.load_reference_args(d)
    b.upload()
    b.set_forcing(met)
    b.advance(0, met.shape[0], OUT_SUM82)
    sums = b.sums()
    assert np.array_equal(sums, np.broadcast_to(sums[0], sums.shape))
    for nm in ("state_id", "soilstore_id", "SnowPack", "SnowDens", "HDD_id"):
        st = b.get_state(nm)
        assert np.array_equal(st, np.broadcast_to(st[0], st.shape)), nm
    b1 = SuewsBatch(1, 1)
    b1.load_reference_args(d)
    b1.upload()
    b1.set_forcing(met)
    b1.advance(0, met.shape[0], OUT_SUM82)
    assert np.array_equal(b1.sums()[0], sums[0])
    ref_nml = checkpoint.read(fx.GOLDEN + "/ref_InitialConditions_2005_EndofRun.nml")
    states = {nm: b.get_state(nm) for nm in ("HDD_id", "GDD_id", "LAI_id", "albEveTr_id", "albDecTr_id", "albGrass_id", "DecidCap_id",
                                             "porosity_id", "soilstore_id", "state_id", "SnowWater", "SnowPack", "SnowFrac",
                                             "SnowDens", "SnowAlb")}
    for cell in (0, 4999, 9999):
        ours = checkpoint.read_text(checkpoint.write_text(checkpoint.state_vector(lambda k: states[k][cell]), snow_use=1))
        assert max(abs(ours[k] - r) / max(abs(r), 1e-3) for k, r in ref_nml.items()) <= 1e-9
    b.close()
    b1.close()


@pytest.mark.parametrize("scheme", ["default", "snow", "baserun"])
def test_differential_seasons(scheme):
    """Synthetic cells started in winter, spring and late summer (melting snow, freezing rain, temperatures crossing
    0 degC: the branches a January start never reaches), two days each: flux sums and end state equal the per-cell
    oracle runs far inside the tolerance.  tools/diff_fuzz.py is the long version of this test."""
    from suews_b200 import OUT_SUM82, SuewsBatch, synth
    base = fx.state_init(2005)
    allmet = fx.forcing(2004)
    n = 16
    for day0 in (59, 121, 288):
        met = allmet[day0 * 288:(day0 + 2) * 288]
        b = SuewsBatch(n, 1)
        _m, site = synth.build_batch(b, base, cell0=17 * day0, scheme=scheme)
        b.upload()
        b.set_forcing(met)
        b.advance(0, met.shape[0], OUT_SUM82)
        sums = b.sums()
        d0 = synth.base_with_scheme(base, scheme)
        for c in range(0, n, 2):
            dd = dict(d0)
            for k in ("sfr", "lat", "lng", "bldgH", "EveTreeH", "DecTreeH", "Z"):
                dd[k.lower()] = np.atleast_1d(site[k][c])
            m = met.copy()
            m[:, 11] += site["dTemp_C"][c]
            m[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), m[:, 10] + site["dRH"][c]))
            cell = ol.Cell(dd)
            rs = cell.run(m, want=("SUEWS",))["SUEWS"][:, 5:].sum(axis=0)
            scale = np.abs(rs[list(PARITY_COLS.values())]).max()
            for nm, col in PARITY_COLS.items():
                assert abs(sums[c, col] - rs[col]) <= 1e-9 * max(abs(rs[col]), 1e-6 * scale), (scheme, day0, c, nm)
            for nm in ("state_id", "soilstore_id", "SnowPack", "SnowWater", "SnowFrac", "qn1_av"):
                assert np.allclose(b.get_state(nm)[c], cell.get_state(nm), rtol=1e-8, atol=1e-10), (scheme, day0, c, nm)
        b.close()


def test_full_size_batch_properties():
    """BASELINE configs[2] size (1 048 576 cells on one GPU) through size-independent properties: no fatal cell,
    energy balance closed by construction (QH is the residual), water balance closed to the model's own
    accuracy, and any block of cells equal, bit for bit, to the same cells run as a small batch."""
    import bench
    from suews_b200 import OUT_SUM82, SuewsBatch, synth
    base, met = bench.load_inputs(1)
    n = 1 << 20
    b = SuewsBatch(n, 1)
    synth.build_batch(b, base, cell0=0, scheme="default")
    b.upload()
    b.set_forcing(met[:144])
    b.advance(0, 96, OUT_SUM82)          # night, including the non-converging stable hours
    b.reset_sums()
    b.advance(96, 48, OUT_SUM82)         # 08:00-12:00
    fatal, _w, _it = b.diag()
    assert (fatal == 0).all()
    s = b.sums()
    assert np.isfinite(s).all()
    energy = s[:, 5] + s[:, 6] - (s[:, 7] + s[:, 8] + s[:, 9])            # QN + QF - QS - QH - QE (snow off: QM = 0)
    assert np.abs(energy).max() <= 1e-9 * np.abs(s[:, 5]).max()
    water = s[:, 13] + s[:, 14] - (s[:, 15] + s[:, 16] + s[:, 17])         # Rain + Irr - Evap - RO - TotCh  [mm]
    assert np.abs(water).max() <= 1e-3
    c0 = 777000
    small = SuewsBatch(64, 1)
    synth.build_batch(small, base, cell0=c0, scheme="default")
    small.upload()
    small.set_forcing(met[:144])
    small.advance(0, 96, OUT_SUM82)
    small.reset_sums()
    small.advance(96, 48, OUT_SUM82)
    assert np.array_equal(small.sums(), s[c0:c0 + 64])
    assert np.array_equal(small.get_state("soilstore_id"), b.get_state("soilstore_id")[c0:c0 + 64])
    small.close()
    b.close()


def test_library_exp_log_equal_cuda_libm_bit_for_bit():
    """dev_math.cuh restates exp/log with constant-bank coefficients; same operations, so the same bits"""
    import ctypes as C
    from suews_b200._lib import lib
    L = lib()
    L.suews_b200_selftest_math.argtypes = [C.c_longlong, C.c_ulonglong, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    be, bl = C.c_longlong(-1), C.c_longlong(-1)
    assert L.suews_b200_selftest_math(200_000_000, 20261019, C.byref(be), C.byref(bl)) == 0
    assert (be.value, bl.value) == (0, 0)


def test_wet_bulb_search_levels_equal_the_plain_loop():
    """dev_latvap.cuh: the screened wet-bulb search of the build (single-precision screen, then double; with
    -DSU_LATVAP_MODE=2 also the root bound) against the plain loop of the reference on 2x10^8 random arguments -- whole temperature range, around 0 degC, low pressure, near saturation, every
    entry state of the SAVE'd switches: same Lv, same exit switches, same warnings, bit for bit."""
    import ctypes as C
    from suews_b200._lib import lib
    L = lib()
    L.suews_b200_selftest_latvap.argtypes = [C.c_longlong, C.c_ulonglong, C.POINTER(C.c_longlong)]
    for seed in (20261020, 7):
        bad = C.c_longlong(-1)
        assert L.suews_b200_selftest_latvap(100_000_000, seed, C.byref(bad)) == 0
        assert bad.value == 0, bad.value


def test_sun_position_entry_point():
    import ctypes as C
    from suews_b200 import _lib
    az, zen = C.c_double(), C.c_double()
    assert _lib.lib().suews_cal_sunposition_c(2005.0, 0.5416666 - 150, 2.0, 57.77, 11.87, 15.0, C.byref(az), C.byref(zen)) == 0
    raz, rzen = ol.sun_position(2005.0, 0.5416666 - 150, 2.0, 57.77, 11.87, 15.0)
    assert abs(az.value - raz) < 1e-9 and abs(zen.value - rzen) < 1e-9


def test_fatal_error_is_reported_not_aborted():
    """RunControl errors surface as SuewsError (the reference STOPs, supy_driver raises RuntimeError)"""
    from suews_b200 import OUT_SUM82, SuewsBatch, SuewsError
    for key, val, code in (("storageheatmethod", 2, 1003),      # no such storage-heat scheme
                           ("netradiationmethod", 7, 1001)):    # RadMethod's bare STOP (narp.f95:95-99)
        d = fx.state_init(2005)
        d[key] = [val]
        b = SuewsBatch(1, 1)
        b.load_reference_args(d)
        b.upload()
        b.set_forcing(fx.forcing(2005)[:12])
        with pytest.raises(SuewsError) as ei:
            b.advance(0, 12, OUT_SUM82)
        assert ei.value.code == code
        b.close()


# ---- ESTM (StorageHeatMethod 4/14): parity unpinned -- no ESTM output of the reference exists; CUDA == oracle
ESTM_VARIANTS = [
    dict(),                                      # ESTMinput.nml of the BaseRun: forced Tsurf, measured indoor temperature
    dict(evolveTibld=1, ibldCHmod=1),            # HVAC hysteresis, ASHRAE convection
    dict(evolveTibld=2, ibldCHmod=2),            # diagnosed indoor temperature, Awbi convection
    dict(TsurfChoice=1),
    dict(TsurfChoice=2, evolveTibld=1),          # four walls
    dict(storageheatmethod=14),                  # same physics, no ESTM output block (driver.f95:2503)
]


def _estm_case(over):
    d = dict(fx.state_init(2005))
    over = dict(over)
    d["storageheatmethod"] = [over.pop("storageheatmethod", 4)]
    ep = fx.estm_example(d["surfacearea"][0], **over)
    d.update({k.lower(): v for k, v in ep.items()})
    return d, ep


@pytest.mark.gpu
@pytest.mark.parametrize("over", ESTM_VARIANTS, ids=lambda o: "-".join("%s%s" % kv for kv in o.items()) or "baserun")
def test_estm_first_day_matches_oracle(over):
    from suews_b200 import OUT_FULL, SuewsBatch, estm
    d, ep = _estm_case(over)
    n = 288
    met = fx.forcing(2005)[:n]
    ts = fx.estm_ts_forcing(met)
    cell = ol.Cell(d)
    cell.estm_start(ts[0])
    ref = cell.run(met, want=("SUEWS", "ESTM"), ts5=ts)
    b = SuewsBatch(1, 1)
    b.load_reference_args(d)
    st0 = estm.initial_state(ep, d["sfr"], d["emis"], ts[0])
    for k, v in st0.items():
        b.set_state(k, v)
    b.upload()
    b.set_forcing(met)
    b.set_estm_forcing(ts)
    b.advance(0, n, OUT_FULL)
    out = b.block("SUEWS", 0, n)
    assert np.array_equal(out[:, :5], ref["SUEWS"][:, :5])
    assert float(rel_err(out[:, 5:], ref["SUEWS"][:, 5:], "SUEWS").max()) <= 1e-10
    if d["storageheatmethod"][0] == 4:
        oe, re_ = b.block("ESTM", 0, n), ref["ESTM"]
        assert np.array_equal(oe[:, :5], re_[:, :5])
        assert np.array_equal(oe == -999.0, re_ == -999.0)
        scale = np.maximum(np.abs(re_[:, 5:]), np.where(np.abs(re_[:, 5:]).max(axis=0) > 200, 273.15, 1.0))
        assert float((np.abs(oe[:, 5:] - re_[:, 5:]) / scale).max()) <= 1e-10
    # the element temperatures ESTM carries between calls, and its SAVE'd switches
    for name in ("Tievolve", "Tground", "Twall", "Troof", "Tibld", "Tw_4", "T0_ibld", "TN_wall", "TN_roof", "lup_ground",
                 "lup_wall", "lup_roof", "Tair2"):
        np.testing.assert_allclose(b.get_state(name)[0], cell.get_state(name), rtol=1e-11, atol=0, err_msg=name)
    for name in ("HVAC", "Tair2Set", "NewtonConverged"):
        assert np.array_equal(b.get_state(name)[0], cell.get_state(name)), name
    b.close()


@pytest.mark.gpu
def test_estm_without_inputs_fails_loudly():
    from suews_b200 import OUT_FULL, SuewsBatch, SuewsError
    d, _ep = _estm_case({})
    b = SuewsBatch(1, 1)
    b.load_reference_args(d)
    b.upload()
    b.set_forcing(fx.forcing(2005)[:2])
    with pytest.raises(SuewsError) as e:
        b.advance(0, 2, OUT_FULL)       # no set_estm_forcing
    assert e.value.code == 1005
    b.close()


# ---- forcing preparation on the device: DisaggregateMet as a kernel, bit-identical to the numpy
# restatement whose output reproduces the reference goldens (tests/test_oracle_golden.py)
@pytest.mark.gpu
@pytest.mark.parametrize("year,method", [(2005, 1), (2004, 1), (2005, 2), (2005, 3)])
def test_device_disaggregation_is_bit_identical(year, method):
    import os
    from suews_b200 import SuewsBatch
    from suews_b200.forcing import disaggregate_met, read_met
    path = os.path.join(fx.GOLDEN, "forcing_%d_60min.txt" % year)
    host = disaggregate_met(read_met(path), disagg_method=method)
    # the file's own values (pressure in kPa): the device applies MetRead's conversions
    rows = [ln.split()[:24] for ln in open(path).readlines()[1:] if len(ln.split()) >= 24 and not ln.startswith("-9")]
    raw = np.array(rows, dtype=np.float64)  # exactly what the file holds
    b = SuewsBatch(1, 1)
    b.load_reference_args(fx.state_init(2005))
    b.upload()
    n = b.set_forcing_orig(raw, resolution_in=3600, disagg_method=method, raw_units=True, smd_method=0)
    assert n == host.shape[0]
    dev = b.forcing_block()
    cols = [c for c in range(24) if c not in (23,)]       # wind direction: left unset by the reference (-999 both sides)
    assert np.array_equal(dev[:, cols], host[:, cols])
    assert np.all(dev[:, 23] == -999.0)
    b.close()


@pytest.mark.gpu
def test_run_from_device_disaggregated_forcing_equals_host_path():
    import os
    from suews_b200 import OUT_SUM82, SuewsBatch
    from suews_b200.forcing import read_met
    d = fx.state_init(2005)
    met = fx.forcing(2005)
    sums = []
    for dev_path in (False, True):
        b = SuewsBatch(1, 1)
        b.load_reference_args(d)
        b.upload()
        if dev_path:
            b.set_forcing_orig(read_met(os.path.join(fx.GOLDEN, "forcing_2005_60min.txt")), raw_units=False)
        else:
            b.set_forcing(met)
        b.advance(0, 576, OUT_SUM82)
        sums.append(b.sums().copy())
        b.close()
    assert np.array_equal(sums[0], sums[1])


# ---- output policy P1: SUEWS_Output_Agg (suews_ctrl_output.f95:912-957) inside the kernel
def _agg_sequential(block, nper):
    """mean / sum / last per column with the additions in row order, as the reference's SUM over a column slice"""
    vl = fx.varlist("SUEWS")
    n = block.shape[0] // nper
    out = np.zeros((n, block.shape[1]))
    out[:, :5] = block[nper - 1::nper, :5][:n]
    for c in range(block.shape[1] - 5):
        col = block[: n * nper, 5 + c].reshape(n, nper)
        a = vl[c]["aggreg"]
        s = np.cumsum(col, axis=1)[:, -1]
        out[:, 5 + c] = s / nper if a == "A" else (s if a == "S" else col[:, -1])
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("freq", [3600, 1800])
def test_in_kernel_aggregation_equals_aggregating_the_full_block(gpu_and_ref_2005, freq):
    from suews_b200 import OUT_AGG, SuewsBatch
    bfull, ref, _cell, met = gpu_and_ref_2005
    n = met.shape[0]
    nper = freq // 300
    full = bfull.block("SUEWS", 0, n)
    b = SuewsBatch(1, 1)
    b.load_reference_args(fx.state_init(2005))
    b.upload()
    b.set_forcing(met)
    b.set_output_freq(freq)
    b.advance(0, n, OUT_AGG)
    got = b.agg_block(0, n // nper)
    want = _agg_sequential(full, nper)
    assert np.array_equal(got, want)                      # same kernel arithmetic, same order of additions: bit for bit
    oracle = _agg_sequential(ref["SUEWS"], nper)
    assert np.array_equal(got[:, :5], oracle[:, :5])
    assert float(rel_err(got[:288 // nper, 5:], oracle[:288 // nper, 5:], "SUEWS").max()) <= TOL_DAY1
    assert np.array_equal(b.agg_column(3, 5 + 5), got[3:4, 5 + 5])
    # the DailyState rows of the same launch: the rows of the full block that carry data, bit for bit
    fd = bfull.block("DailyState", 0, n)
    assert np.array_equal(b.agg_daily_block(0), fd[fd[:, 5] != -999])
    b.close()


@pytest.mark.gpu
def test_hourly_aggregates_match_the_reference_golden_file():
    """the reference's own hourly output (Test/BaseRun/2020b, 120 h x 82 columns, 4-6 decimals)"""
    from suews_b200 import OUT_AGG, SuewsBatch
    met = fx.forcing(2005)
    b = SuewsBatch(1, 1)
    b.load_reference_args(fx.state_init(2005))
    b.upload()
    b.set_forcing(met)
    b.advance(0, met.shape[0], OUT_AGG)
    got = b.agg_block(0, met.shape[0] // 12)
    hdr, gold = fx.load_ref_table("ref_2005_SUEWS_60.txt")
    assert got.shape == gold.shape == (120, 87)
    assert np.array_equal(got[:, :4], gold[:, :4])
    tol = fx.text_tolerance("SUEWS", 82)
    for c in range(82):
        r = gold[:, 5 + c]
        ok = np.isnan(r) | (np.abs(got[:, 5 + c] - r) <= tol[c] + 1e-9 * np.abs(np.nan_to_num(r)))
        assert ok.all(), (hdr[5 + c], int(np.argmin(ok)), got[~ok, 5 + c][:3], r[~ok][:3])
    # ... and its DailyState file, through the library's writer: byte for byte
    import ctypes as C
    from suews_b200 import _lib
    f = _lib.lib().suews_write_txt
    f.restype = C.c_longlong
    f.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_longlong]
    d = np.asfortranarray(b.agg_daily_block(0))
    buf = C.create_string_buffer(1 << 16)
    f(d.ctypes.data_as(C.POINTER(C.c_double)), d.shape[0], d.shape[1], b"DailyState", 2, 1, buf, 1 << 16)
    assert buf.value.decode() == open(fx.GOLDEN + "/ref_2005_DailyState.txt").read()
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("method", [1, 2])
def test_device_estm_forcing_disaggregation(method):
    """DisaggregateESTM on the device: bit-identical to the numpy restatement (itself pinned to the reference's 5-min
    file), and an ESTM run driven by it equals the run driven by the host-disaggregated block"""
    import os
    from suews_b200 import OUT_SUM82, SuewsBatch, estm
    from suews_b200.forcing import disaggregate_estm, read_estm_ts, read_met
    d, ep = _estm_case({})
    ts_h = read_estm_ts(os.path.join(fx.GOLDEN, "estm_ts_2005_60min.txt"))
    host = disaggregate_estm(ts_h, disagg_method=method)
    sums = []
    for dev_path in (True, False):
        b = SuewsBatch(1, 1)
        b.load_reference_args(d)
        for k, v in estm.initial_state(ep, d["sfr"], d["emis"], host[0]).items():
            b.set_state(k, v)
        b.upload()
        b.set_forcing_orig(read_met(os.path.join(fx.GOLDEN, "forcing_2005_60min.txt")))
        if dev_path:
            dev = b.set_estm_forcing_orig(ts_h, disagg_method=method)
            assert np.array_equal(dev, host)
        else:
            b.set_estm_forcing(host)
        b.advance(0, 288, OUT_SUM82)
        sums.append(b.sums().copy())
        b.close()
    assert np.array_equal(sums[0], sums[1])


@pytest.mark.gpu
def test_flux_only_policy_changes_no_flux_and_no_state():
    """Policy P0 (SUEWS_OUT_SUM7 alone) skips the roughness-sublayer diagnostics, which its seven sums cannot observe:
    the flux sums and the evolving state are bit-identical to those of a launch that runs every stage
    (SUEWS_OUT_ALL_STAGES); only the library's own z0_RSL memo differs"""
    from suews_b200 import OUT_ALL_STAGES, OUT_SUM7, SuewsBatch, synth
    base = fx.state_init(2005)
    met = fx.forcing(2005)[:576]
    res = []
    for flags in (OUT_SUM7 | OUT_ALL_STAGES, OUT_SUM7):
        b = SuewsBatch(256, 1)
        synth.build_batch(b, base, cell0=0, scheme="default")
        b.upload()
        b.set_forcing(met)
        b.advance(0, 576, flags)
        st = {nm: b.get_state(nm).copy() for nm, kind, _n in __import__("suews_b200.batch", fromlist=["all_fields"]).all_fields()
              if kind == 3 and nm != "rsl_cache"}
        res.append((b.sums()[:, list(PARITY_COLS.values())].copy(), st, b.last_ms()))
        b.close()
    assert np.array_equal(res[0][0], res[1][0])
    for nm in res[0][1]:
        assert np.array_equal(res[0][1][nm], res[1][1][nm]), nm


@pytest.mark.gpu
def test_mixed_scheme_classes_in_one_batch():
    """Several parameter classes in one batch (the code-keyed tables of the reference: cells of different scheme sets and
    surface properties side by side).  upload() groups the threads by class; every cell must come out exactly as in a
    single-class batch of its own scheme."""
    from suews_b200 import OUT_SUM82, SuewsBatch, synth
    from suews_b200.batch import all_fields, KIND_CELL, KIND_STATE
    base = fx.state_init(2005)
    met = fx.forcing(2005)[:288]
    schemes = ["default", "baserun", "stab2"]
    ncell = 96
    cls_of = np.arange(ncell) % 3
    site = synth.cell_site_arrays(0, ncell, {k.lower(): v for k, v in base.items()})
    dicts = [synth.base_with_scheme(base, s) for s in schemes]
    # reference: one single-class batch per scheme over all 96 cells
    want_sums, want_state = [], []
    for d in dicts:
        b = SuewsBatch(ncell, 1)
        b.load_reference_args(d)
        for k, v in site.items():
            b.set_cell_param(k, v if v.ndim == 2 else v.reshape(-1, 1))
        b.upload()
        b.set_forcing(met)
        b.advance(0, 288, OUT_SUM82)
        want_sums.append(b.sums().copy())
        want_state.append({nm: b.get_state(nm).copy() for nm in ("state_id", "soilstore_id", "SnowPack", "alb", "HDD_id")})
        b.close()
    # the mixed batch
    b = SuewsBatch(ncell, 3)
    for k, d in enumerate(dicts):
        b.load_reference_args(d, cls=k, cells=(k == 0))
    for k, v in site.items():
        b.set_cell_param(k, v if v.ndim == 2 else v.reshape(-1, 1))
    for nm, kind, n in all_fields():           # per-cell state follows the cell's class (snow-free start for snow-off sets)
        if kind == KIND_STATE and nm.lower() in dicts[0]:
            vals = np.stack([np.asarray(dicts[c][nm.lower()], dtype=float).ravel() for c in cls_of])
            b.set_state(nm, vals)
    b.set_cell_class(cls_of)
    b.upload()
    b.set_forcing(met)
    b.advance(0, 288, OUT_SUM82)
    sums = b.sums()
    fatal, _w, _it = b.diag()
    assert not fatal.any()
    for c in range(ncell):   # (the 2 m humidity diagnostics of some synthetic cells are NaN in the oracle too: equal_nan)
        assert np.array_equal(sums[c], want_sums[cls_of[c]][c], equal_nan=True), c
    for nm in want_state[0]:
        st = b.get_state(nm)
        for c in range(ncell):
            assert np.array_equal(st[c], want_state[cls_of[c]][nm][c], equal_nan=True), (nm, c)
    b.close()


@pytest.mark.gpu
def test_baserun_output_files_from_the_gpu_path(tmp_path):
    """examples/baserun_2005.py: hourly forcing file in, the reference's own output file and checkpoint out.  The CUDA
    path differs from the reference by libm ulps, so a printed digit can only differ for a value sitting on a rounding
    boundary of its last decimal"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("baserun_2005", os.path.join(os.path.dirname(fx.GOLDEN), "..", "examples", "baserun_2005.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    bad, fields, worst = mod.main(str(tmp_path))
    assert fields == 120 * 87
    assert bad <= 3, bad                # measured: 0 -- the file is byte-identical to the reference's
    assert worst <= 1e-12               # measured: 0


@pytest.mark.gpu
def test_multi_device_handle_equals_one_batch():
    """suews_b200_multi_*: the cells of a run in contiguous blocks on several devices (here two blocks -- on two GPUs when
    the box has them, else both on device 0), advanced side by side, sums gathered to the first device and to the host.
    Cells are independent, so every cell must come out bit for bit as in one single-device batch."""
    from suews_b200 import OUT_SUM7, SuewsBatch, SuewsMultiBatch, synth, _lib
    base = fx.state_init(2005)
    met = fx.forcing(2005)[:288]
    ncell = 2500                                   # not a multiple of anything: blocks of 834 / 833 / 833 cells
    ngpu = _lib.lib().suews_b200_device_count()
    devices = [0, 1 % ngpu, 2 % ngpu]
    cols = list(PARITY_COLS.values())
    one = SuewsBatch(ncell, 1)
    synth.build_batch(one, base, scheme="baserun")
    one.upload()
    one.set_forcing(met)
    one.advance(0, 288, OUT_SUM7)
    multi = SuewsMultiBatch(ncell, 1, devices=devices)
    synth.build_batch(multi, base, scheme="baserun")
    assert [p[1:] for p in multi.parts()] == [(0, 834), (834, 833), (1667, 833)]
    multi.upload()
    multi.set_forcing(met)
    multi.advance(0, 144, OUT_SUM7, sync=False)    # two launches per device, asynchronous
    multi.advance(144, 144, OUT_SUM7)
    got, ms = multi.gather_sums(cols)
    assert ms > 0
    assert np.array_equal(got, one.sums_cols(cols))
    for nm in ("state_id", "soilstore_id", "SnowPack", "HDD_id", "qn1_av"):
        assert np.array_equal(multi.get_state(nm), one.get_state(nm)), nm
    f1, _w1, it1 = one.diag()
    f2, _w2, it2 = multi.diag()
    assert np.array_equal(f1, f2) and np.array_equal(it1, it2)
    one.close()
    multi.close()


@pytest.mark.gpu
def test_multi_device_handle_shares_the_forcing_and_cbl_blocks(tmp_path):
    """The rest of the multi-device surface: a three-grid run directory (two classes, BLUEWS on) loaded into a
    SuewsMultiBatch, forcing at file resolution disaggregated on every device with KdownZen = 1, the CBL initial data and an
    explicit AnOHM forcing grid set through the multi handle -- per-step blocks of every cell, from whichever device holds
    it, bit for bit those of one single-device batch."""
    from suews_b200 import OUT_FULL, SuewsBatch, SuewsMultiBatch, tables, _lib
    from suews_b200.forcing import read_met
    from test_tables import _three_grid_run
    ga = tables.grid_arguments(_three_grid_run(tmp_path), 2004)
    hourly = read_met(os.path.join(fx.GOLDEN, "forcing_2004_60min.txt"))[:72]
    site = tuple(float(ga["cells"][k][0, 0]) for k in ("lat", "lng", "timezone", "alt"))
    ngpu = _lib.lib().suews_b200_device_count()
    out = []
    for make in (lambda: SuewsBatch(3, 2), lambda: SuewsMultiBatch(3, 2, devices=[0, 1 % ngpu])):
        b = tables.load_batch(make(), ga)
        b.upload()
        b.set_cbl_initial(ga["cbl_initial"])
        n = b.set_forcing_orig(hourly, resolution_in=3600, kdown_zen=1, site=site)
        assert n == 72 * 12
        b.set_anohm_forcing(b.forcing_block())        # (accepted; no AnOHM class in this run)
        b.advance(0, n, OUT_FULL)
        out.append(dict(met=b.forcing_block(), su=[b.block("SUEWS", i, n) for i in range(3)], bl=[b.block("BL", i, n) for i in range(3)],
                        st={k: b.get_state(k) for k in ("cbl_blh_m", "cbl_tm_K", "soilstore_id", "state_id")}))
        b.close()
    one, multi = out
    assert np.array_equal(one["met"], multi["met"])
    for i in range(3):
        assert np.array_equal(one["su"][i], multi["su"][i]) and np.array_equal(one["bl"][i], multi["bl"][i]), i
    assert float(np.abs(one["bl"][0][:, 5]).max()) > 200.0            # the slab was running
    for k, v in one["st"].items():
        assert np.array_equal(v, multi["st"][k]), k


def _oracle_cells_worker(args):
    """(oracle, in a pool worker) flux sums of day 1 for synthetic cells [c0, c1) of the given scheme"""
    c0, c1, scheme = args
    from suews_b200 import synth
    base = fx.state_init(2005)
    met = fx.forcing(2005)[:288]
    d0 = synth.base_with_scheme(base, scheme)
    site = synth.cell_site_arrays(c0, c1 - c0, {k.lower(): v for k, v in base.items()})
    out = np.zeros((c1 - c0, 82))
    for c in range(c1 - c0):
        dd = dict(d0)
        for k in ("sfr", "lat", "lng", "bldgH", "EveTreeH", "DecTreeH", "Z"):
            dd[k.lower()] = np.atleast_1d(site[k][c])
        m = met.copy()
        m[:, 11] += site["dTemp_C"][c]
        m[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), m[:, 10] + site["dRH"][c]))
        out[c] = ol.Cell(dd).run(m, want=("SUEWS",))["SUEWS"][:, 5:].sum(axis=0)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["baserun", "default"])
def test_ten_thousand_perturbed_cells_match_per_cell_oracle(scheme):
    """BASELINE configs[1] scale: 10 000 cells, every one with its own surface fractions, location, morphology and forcing
    offsets (not replicas), one launch of the first day; EVERY cell is compared with its own single-cell oracle run (one
    oracle process per host core) on the day sums of all 82 columns.  The reference's multi-grid == single-grid property
    (Test/code/Test_SUEWS.py:384) at the north_star tolerance instead of 4 decimals."""
    import multiprocessing as mp
    import os
    from suews_b200 import OUT_SUM82, SuewsBatch, synth
    n = 10000
    base = fx.state_init(2005)
    met = fx.forcing(2005)[:288]
    b = SuewsBatch(n, 1)
    synth.build_batch(b, base, cell0=0, scheme=scheme)
    b.upload()
    b.set_forcing(met)
    b.advance(0, 288, OUT_SUM82)
    sums = b.sums()
    fatal, _warn, _it = b.diag()
    b.close()
    assert (fatal == 0).all()
    ol.build()
    ncore = max(1, min(32, os.cpu_count() or 1))
    step = (n + 4 * ncore - 1) // (4 * ncore)
    jobs = [(c0, min(n, c0 + step), scheme) for c0 in range(0, n, step)]
    with mp.get_context("fork").Pool(ncore) as pool:
        ref = np.concatenate(pool.map(_oracle_cells_worker, jobs), axis=0)
    assert ref.shape == sums.shape
    # columns that are NaN in the reference itself (q2 / RH2 on steps with QE = 0, rslprof.f95:285) must be NaN here too
    assert np.array_equal(np.isnan(sums), np.isnan(ref))
    nan_cells = int(np.isnan(ref).any(axis=1).sum())
    ok = ~np.isnan(ref)
    colscale = np.nanmax(np.abs(ref), axis=0, keepdims=True) + 1e-300
    with np.errstate(invalid="ignore", divide="ignore"):
        e_val = np.where(ok, np.abs(sums - ref) / np.maximum(np.abs(ref), 1e-3 * 288), 0.0)     # per value, floor 1e-3 per step
        e_col = np.where(ok, np.abs(sums - ref) / np.maximum(np.abs(ref), colscale), 0.0)       # against the column's largest sum
    cols = list(PARITY_COLS.values())
    print("%s: %d cells, %d with NaN q2/RH2 sums as in the reference; worst day-sum error: parity fluxes %.2e (per value), all "
          "columns %.2e (per value) / %.2e (column scale)" % (scheme, n, nan_cells, e_val[:, cols].max(), e_val.max(), e_col.max()))
    # A last-ulp libm difference can flip one of the model's threshold tests (|dTs| > 0.1 of the surface-temperature loop,
    # state <= 0.001, ...) for a single step of a single cell: ~1e-6 on that step, ~1e-8 on the day sum.  Among 2.9 million
    # cell-steps that may happen to a few cells, so the bound is statistical: every cell inside the flux-sum tolerance of
    # north_star (1e-6), and all but at most one in a thousand inside ten times the per-step tolerance.
    per_cell = np.maximum(e_val[:, cols].max(axis=1), e_col.max(axis=1))
    flipped = int((per_cell > TOL_DAY1 * 10).sum())
    print("   cells beyond 1e-9: %d of %d (worst %.2e)" % (flipped, n, per_cell.max()))
    assert per_cell.max() <= TOL_SUM
    assert flipped <= n // 1000


@pytest.mark.gpu
def test_resort_between_launches_changes_nothing():
    """suews_b200_resort: the thread slots are sorted again on the current state (snow activity, canopy height) in the
    middle of a run; sums, state, diagnostics and the clock must be those of the uninterrupted run, bit for bit."""
    from suews_b200 import OUT_SUM7, SuewsBatch, synth
    base = fx.state_init(2004)             # winter start: snow on the ground, melting unevenly over the perturbed cells
    met = fx.forcing(2004)[:3 * 288]
    n = 3000
    runs = []
    for resort in (False, True):
        b = SuewsBatch(n, 1)
        synth.build_batch(b, base, scheme="baserun")
        # half of the cells start bare: the snow-activity key has something to separate
        b.upload()
        bare = np.arange(n) % 2 == 1
        for nm in ("SnowPack", "SnowFrac", "SnowWater"):
            v = np.tile(np.asarray(base[nm.lower()], dtype=float), (n, 1))
            v[bare] = 0.0
            b.set_state(nm, v)
        b.upload()
        b.set_forcing(met)
        b.advance(0, 288, OUT_SUM7)
        if resort:
            b.resort()
        b.advance(288, 288, OUT_SUM7)
        if resort:
            b.resort()
        b.advance(576, 288, OUT_SUM7)
        runs.append((b.sums_cols(list(PARITY_COLS.values())), {nm: b.get_state(nm) for nm in ("SnowPack", "state_id", "soilstore_id", "HDD_id", "SnowDens")},
                     b.diag(), b.get_clock()))
        b.close()
    assert np.array_equal(runs[0][0], runs[1][0])
    for nm in runs[0][1]:
        assert np.array_equal(runs[0][1][nm], runs[1][1][nm]), nm
    assert np.array_equal(runs[0][2][0], runs[1][2][0]) and np.array_equal(runs[0][2][2], runs[1][2][2])
    assert runs[0][3] == runs[1][3]


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["default", "snow", "estm", "netrad13"])
def test_specialised_p0_instances_equal_the_generic_kernel(scheme):
    """Batches large enough for the 1024-thread geometry run policy P0 through the kernel instance compiled for their
    bucket (snow off / snow on / ESTM per class; suews_kernel.cuh): sums, state and ESTM element temperatures must be
    bit-identical to the accumulating kernel that runs every stage (SUEWS_OUT_ALL_STAGES, the generic instance)."""
    from suews_b200 import OUT_ALL_STAGES, OUT_SUM7, SuewsBatch, synth
    from suews_b200.batch import all_fields
    base = fx.state_init(2004 if scheme == "snow" else 2005)
    met = fx.forcing(2004 if scheme == "snow" else 2005)[:288]
    n = 60000                                    # above the small-geometry threshold (1.5 x 256 cells per SM)
    use_estm = scheme == "estm"
    ts5 = synth.estm_ts_forcing(met) if use_estm else None
    res = []
    for flags in (OUT_SUM7 | OUT_ALL_STAGES, OUT_SUM7):
        b = SuewsBatch(n, 1)
        synth.build_batch(b, base, cell0=0, scheme=scheme, ts_row0=ts5[0] if use_estm else None)
        b.upload()
        b.set_forcing(met)
        if use_estm:
            b.set_estm_forcing(ts5)
        b.advance(0, 144, flags)
        b.advance(144, 144, flags)
        st = {nm: b.get_state(nm).copy() for nm, kind, _n in all_fields()
              if (kind == 3 and nm != "rsl_cache") or (kind == 4 and use_estm)}
        fatal, _w, iters = b.diag()
        res.append((b.sums_cols(list(PARITY_COLS.values())), st, fatal.copy(), iters.copy()))
        b.close()
    assert not res[0][2].any() and not res[1][2].any()
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][3], res[1][3])
    for nm in res[0][1]:
        assert np.array_equal(res[0][1][nm], res[1][1][nm], equal_nan=True), nm


# ---- AnOHM (StorageHeatMethod 3): the once-a-day fit runs inside the kernel; its memory is per-cell state.
# Tolerance: the fit (Levenberg-Marquardt to a relative 1e-5 with a forward-difference Jacobian) amplifies last-ulp
# differences of sin / cos / exp between CUDA's libm and glibc by ~1e8: tests/test_anohm_oracle.py measures the
# coefficients moving by 1e-7 (median) to a few 1e-4 under one-ulp shifts, with the same source on the host bit-identical to
# the oracle.  So this scheme is bounded at 1e-5 per step / on flux sums instead of 1e-10 / 1e-6 (measured on the B200:
# 2e-8 on the full output of the first two days).
TOL_ANOHM = 1e-5
ANOHM_VARIANTS = [{}, {"snowuse": 0, "netradiationmethod": 3, "emissionsmethod": 2},
                  {"snowuse": 0, "netradiationmethod": 3, "emissionsmethod": 0, "ohmincqf": 1}]


def _anohm_case(over):
    from suews_b200 import synth
    d = fx.state_init(2005)
    d.update({k: list(v) for k, v in synth.ANOHM_EXAMPLE.items()})
    d["storageheatmethod"] = [3]
    for k, v in over.items():
        d[k] = [v]
    if d["snowuse"][0] == 0:
        for k in ("snowpack", "snowfrac", "snowwater"):
            d[k] = [0.0] * 7
    return d


@pytest.mark.gpu
@pytest.mark.parametrize("over", ANOHM_VARIANTS, ids=lambda o: "-".join("%s%s" % kv for kv in o.items()) or "baserun")
def test_anohm_two_days_match_oracle(over):
    """Full output of 2 days + 20 steps (three fits: the first step, and the first step of each new day), the DailyState
    row's a1/a2/a3 and AnOHM's memory, against the oracle whose MetForcingData_grid is the same block."""
    from suews_b200 import OUT_FULL, SuewsBatch
    d = _anohm_case(over)
    n = 288 * 2 + 20
    met = fx.forcing(2005)[:n]
    cell = ol.Cell(d)
    ref = cell.run(met, want=("SUEWS", "DailyState"))
    b = SuewsBatch(1, 1)
    b.load_reference_args(d)
    b.upload()
    b.set_forcing(met)
    b.advance(0, n, OUT_FULL)
    out = b.block("SUEWS", 0, n)
    assert np.array_equal(out[:, :5], ref["SUEWS"][:, :5])
    e = rel_err(out[:, 5:], ref["SUEWS"][:, 5:], "SUEWS")
    print("AnOHM %s: worst rel err %.3e" % (over, float(e.max())))
    assert float(e.max()) <= TOL_ANOHM, (float(e.max()), np.unravel_index(e.argmax(), e.shape))
    od, rd = b.block("DailyState", 0, n), ref["DailyState"]
    assert np.array_equal(od == -999.0, rd == -999.0)
    np.testing.assert_allclose(od[:, -3:], rd[:, -3:], rtol=TOL_ANOHM, atol=0)
    for name in ("anohm_id_save", "anohm_coef_id_save", "anohm_coef_grid_save"):
        assert np.array_equal(b.get_state(name)[0], cell.get_state(name)), name
    np.testing.assert_allclose(b.get_state("anohm_coeff_grid_day")[0], cell.get_state("anohm_coeff_grid_day"), rtol=TOL_ANOHM, atol=0)
    for name in ("qn1_av", "dqndt", "state_id", "soilstore_id"):
        np.testing.assert_allclose(b.get_state(name)[0], cell.get_state(name), rtol=TOL_ANOHM, atol=1e-9, err_msg=name)
    b.close()


@pytest.mark.gpu
def test_anohm_many_cells_policy_p0_and_launch_boundaries():
    """Perturbed cells (fractions, location, morphology, forcing offsets) under policy P0, advanced in three launches that
    cut a day in two (the fit's memory must carry over), each against its own oracle run; the forcing block is staged
    once, so every launch sees the same MetForcingData_grid."""
    from suews_b200 import OUT_SUM7, SuewsBatch, synth
    base = fx.state_init(2005)
    n = 288 * 2
    met = fx.forcing(2005)[:n]
    nc = 40
    b = SuewsBatch(nc, 1)
    _missing, site = synth.build_batch(b, base, cell0=500, scheme="anohm")
    b.upload()
    b.set_forcing(met)
    for s0, s1 in ((0, 100), (100, 400), (400, n)):
        b.advance(s0, s1 - s0, OUT_SUM7)
    sums = b.sums()
    fatal, _warn, _it = b.diag()
    assert (fatal == 0).all()
    d0 = synth.base_with_scheme(base, "anohm")
    worst = 0.0
    for c in range(0, nc, 4):
        dd = dict(d0)
        for k in ("sfr", "lat", "lng", "bldgH", "EveTreeH", "DecTreeH", "Z"):
            dd[k.lower()] = np.atleast_1d(site[k][c])
        m = met.copy()
        m[:, 11] += site["dTemp_C"][c]
        m[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), m[:, 10] + site["dRH"][c]))
        cell = ol.Cell(dd)
        ref = cell.run(m, want=("SUEWS",))["SUEWS"][:, 5:].sum(axis=0)
        for nm, col in PARITY_COLS.items():
            err = abs(sums[c, col] - ref[col]) / max(abs(ref[col]), 1e-30)
            worst = max(worst, err)
            assert err <= TOL_ANOHM, (c, nm, sums[c, col], ref[col])
        np.testing.assert_allclose(b.get_state("anohm_coeff_grid_day")[c], cell.get_state("anohm_coeff_grid_day"), rtol=10 * TOL_ANOHM, atol=0)
    print("AnOHM cells: worst flux-sum rel err %.3e" % worst)
    b.close()


@pytest.mark.gpu
def test_anohm_explicit_met_grid_and_too_many_rows():
    """set_anohm_forcing: a forcing staged in pieces with the whole block handed over as MetForcingData_grid gives the
    whole-block run; a block with more than 24 on-the-hour rows of one day of year (two years) is refused."""
    from suews_b200 import OUT_FULL, SuewsBatch, SuewsError
    d = _anohm_case(ANOHM_VARIANTS[1])
    n = 288 + 50
    met = fx.forcing(2005)[:n]
    b = SuewsBatch(1, 1)
    b.load_reference_args(d)
    b.upload()
    b.set_forcing(met)
    b.advance(0, n, OUT_FULL)
    whole = b.block("SUEWS", 0, n).copy()
    b.close()
    b = SuewsBatch(1, 1)
    b.load_reference_args(d)
    b.upload()
    parts = []
    for s0 in range(0, n, 97):
        s1 = min(n, s0 + 97)
        b.set_forcing(met[s0:s1])
        b.set_anohm_forcing(met)
        b.advance(0, s1 - s0, OUT_FULL)
        parts.append(b.block("SUEWS", 0, s1 - s0).copy())
    assert np.array_equal(np.vstack(parts), whole)
    two_years = np.vstack([met[:288], met[:288]])
    with pytest.raises(SuewsError):
        b.set_forcing(two_years)
    b.close()


# ---- BLUEWS (suews_phys_bluews.f95): the boundary-layer slab model as the last stage of the step
@pytest.mark.gpu
def test_bluews_matches_oracle_and_the_reference_bl_file():
    """The BL output block of the 2005 golden run (1440 steps, nights and days) against the oracle's line step by step, its
    hourly means against the reference's test1_2005_BL_60.txt to the printed digit, and the slab variables at the end."""
    import ctypes as C
    from suews_b200 import OUT_FULL, SuewsBatch
    params, ini = fx.cbl_inputs()
    d = fx.state_init(2005)
    d.update(params)
    met = fx.forcing(2005)
    n = met.shape[0]
    dp = C.POINTER(C.c_double)
    L = ol.lib()
    L.so_set_cbl_initial.argtypes = [dp, C.c_int]
    L.so_set_bl_block.argtypes = [dp]
    ini = np.ascontiguousarray(ini)
    bl = np.full((n, 22), -999.0, order="F")
    L.so_set_cbl_initial(ini.ctypes.data_as(dp), ini.shape[0])
    L.so_set_bl_block(bl.ctypes.data_as(dp))
    cell = ol.Cell(d)
    try:
        ref = cell.run(met, want=("SUEWS",))
    finally:
        L.so_set_bl_block(None)
        L.so_set_cbl_initial(None, 0)
    b = SuewsBatch(1, 1)
    b.load_reference_args(d)
    b.upload()
    b.set_cbl_initial(ini)
    b.set_forcing(met)
    b.advance(0, n, OUT_FULL)
    out = b.block("BL", 0, n)
    assert np.array_equal(out[:, :5], bl[:, :5])
    scale = np.maximum(np.abs(bl[:, 5:]), np.abs(bl[:, 5:]).max(axis=0) * 1e-3)
    e = np.abs(out[:, 5:] - bl[:, 5:]) / scale
    assert float(e.max()) <= TOL_DAY1, (float(e.max()), np.unravel_index(e.argmax(), e.shape))
    su = b.block("SUEWS", 0, n)                                  # the surface scheme is untouched by it
    assert float(rel_err(su[:, 5:], ref["SUEWS"][:, 5:], "SUEWS").max()) <= TOL_DAY1
    _hdr, gold = fx.load_ref_table("ref_2005_BL_60.txt")
    agg = out[:gold.shape[0] * 12, 5:].reshape(gold.shape[0], 12, 17).mean(axis=1)
    assert np.all(np.abs(agg - gold[:, 5:]) <= fx.text_tolerance("BL", 17))
    for name in ("cbl_blh_m", "cbl_tm_K", "cbl_qm_kgkg", "cbl_tp_K", "cbl_qp_kgkg", "cbl_gamt_Km", "cbl_gamq_kgkgm"):
        np.testing.assert_allclose(b.get_state(name)[0], cell.get_state(name), rtol=1e-10, atol=0, err_msg=name)
    b.close()


@pytest.mark.gpu
def test_bluews_entrainment_schemes_and_policy_p0():
    """EntrainmentType 1-4, subsidence on, QH_choice 2 (LUMPS fluxes): slab state after a day under policy P0 (no output
    line) equals the oracle's; without the initial-data table the run stops with 1007."""
    import ctypes as C
    from suews_b200 import OUT_SUM7, SuewsBatch, SuewsError
    params, ini = fx.cbl_inputs()
    ini = np.ascontiguousarray(ini)
    met = fx.forcing(2004)[288 * 150:288 * 151 + 100]         # a summer day: the slab grows
    dp = C.POINTER(C.c_double)
    L = ol.lib()
    L.so_set_cbl_initial.argtypes = [dp, C.c_int]
    for et, isubs, qhc in ((1, 0, 1), (2, 0, 1), (3, 1, 1), (4, 1, 2)):
        d = fx.state_init(2004)
        d.update(params)
        d.update(entrainmenttype=[et], isubs=[isubs], qh_choice=[qhc], snowuse=[0], netradiationmethod=[3])
        for k in ("snowpack", "snowfrac", "snowwater"):
            d[k] = [0.0] * 7
        L.so_set_cbl_initial(ini.ctypes.data_as(dp), ini.shape[0])
        cell = ol.Cell(d)
        try:
            cell.run(met, want=("SUEWS",))
        finally:
            L.so_set_cbl_initial(None, 0)
        b = SuewsBatch(1, 1)
        b.load_reference_args(d)
        b.upload()
        b.set_cbl_initial(ini)
        b.set_forcing(met)
        b.advance(0, met.shape[0], OUT_SUM7)
        for name in ("cbl_blh_m", "cbl_tm_K", "cbl_qm_kgkg", "cbl_tp_K", "cbl_qp_kgkg"):
            np.testing.assert_allclose(b.get_state(name)[0], cell.get_state(name), rtol=1e-9, atol=0, err_msg="%s et=%d" % (name, et))
        assert b.get_state("cbl_blh_m")[0, 0] > 250.0
        b.close()
    b = SuewsBatch(1, 1)
    b.load_reference_args(d)
    b.upload()
    b.set_forcing(met[:10])
    with pytest.raises(SuewsError) as e:
        b.advance(0, 10, OUT_SUM7)
    assert e.value.code == 1007
    b.close()


@pytest.mark.gpu
def test_bluews_start_values_diagnosed_from_the_forcing():
    """InitialData_use = 1: theta and q of the slab's start diagnosed from the night's line through psih (recomputed on the
    device from the step's last z/L; the oracle keeps the value AerodynamicResistance left behind) on a day the file lists,
    the night's 200 m values on a day it does not -- line by line and in the state, with the output line (policy FULL) and
    without it (P0)."""
    import ctypes as C
    from suews_b200 import OUT_FULL, OUT_SUM7, SuewsBatch, SuewsError
    dp = C.POINTER(C.c_double)
    L = ol.lib()
    L.so_set_cbl_initial.argtypes = [dp, C.c_int]
    L.so_set_bl_block.argtypes = [dp]
    params, ini = fx.cbl_inputs()
    ini = np.ascontiguousarray(ini[:1])                       # day 1 only: day 2 finds no row of its own
    d = fx.state_init(2005)
    d.update(params)
    d["initialdata_use"] = [1]
    met = fx.forcing(2005)[:288 + 150]                        # day 1, and day 2 until after noon: the slab grows from the 200 m start
    names = ("cbl_blh_m", "cbl_tm_K", "cbl_qm_kgkg", "cbl_tp_K", "cbl_qp_kgkg", "cbl_gamt_Km", "cbl_gamq_kgkgm")
    for nsteps, policy in ((36, OUT_SUM7), (met.shape[0], OUT_FULL)):
        bl = np.full((nsteps, 22), -999.0, order="F")
        L.so_set_cbl_initial(ini.ctypes.data_as(dp), ini.shape[0])
        L.so_set_bl_block(bl.ctypes.data_as(dp))
        cell = ol.Cell(d)
        try:
            cell.run(met[:nsteps], want=("SUEWS",))
        finally:
            L.so_set_bl_block(None)
            L.so_set_cbl_initial(None, 0)
        b = SuewsBatch(1, 1)
        b.load_reference_args(d)
        b.upload()
        b.set_cbl_initial(ini)
        b.set_forcing(met[:nsteps])
        b.advance(0, nsteps, policy)
        for name in names:
            np.testing.assert_allclose(b.get_state(name)[0], cell.get_state(name), rtol=1e-9, atol=0, err_msg="%s after %d steps" % (name, nsteps))
        if policy == OUT_FULL:
            out = b.block("BL", 0, nsteps)
            scale = np.maximum(np.abs(bl[:, 5:]), np.abs(bl[:, 5:]).max(axis=0) * 1e-3)
            e = np.abs(out[:, 5:] - bl[:, 5:]) / scale
            assert np.array_equal(out[:, :5], bl[:, :5]) and float(e.max()) <= 1e-9, float(e.max())
            assert out[:, 5].max() > 250.0                                      # it grew
        else:
            assert b.get_state("cbl_tm_K")[0, 0] != ini[0, 6] and b.get_state("cbl_blh_m")[0, 0] == ini[0, 1]
        b.close()
    d["initialdata_use"] = [0]                                # no behaviour in the reference (a table it never allocated): refused
    b = SuewsBatch(1, 1)
    b.load_reference_args(d)
    b.upload()
    b.set_cbl_initial(ini)
    b.set_forcing(met[:4])
    with pytest.raises(SuewsError) as err:
        b.advance(0, 4, OUT_SUM7)
    assert err.value.code == 1007
    b.close()


@pytest.mark.gpu
def test_kdown_zen_on_the_device_equals_the_numpy_restatement():
    """KdownZen = 1 (suews_ctrl_input.f95:1568-1600) applied by the device to the block it disaggregated, against the numpy
    restatement fed with the oracle's sun position: same mask, same day sums, same bits."""
    import os
    from suews_b200 import SuewsBatch
    from suews_b200.forcing import disaggregate_met, read_met
    met = read_met(os.path.join(fx.GOLDEN, "forcing_2005_60min.txt"))
    lat, lng, tz, alt = 51.5, -0.1, 0.0, 20.0
    host = disaggregate_met(met, kdown_zen=1, zenith_fn=lambda y, t: ol.sun_position(y, t, tz, lat, lng, alt)[1])
    b = SuewsBatch(1, 1)
    b.load_reference_args(fx.state_init(2005))
    b.upload()
    n = b.set_forcing_orig(met, resolution_in=3600, disagg_method=1, kdown_zen=1, site=(lat, lng, tz, alt))
    assert n == host.shape[0]
    dev = b.forcing_block()
    cols = [c for c in range(24) if c != 23]
    # the zenith at the horizon crossing can differ in its last bits between CUDA's and glibc's trigonometry: a step whose
    # zenith sits within 1e-9 deg of 90 may be masked on one side only -- none in this file
    assert np.array_equal(dev[:, cols], host[:, cols])
    assert (dev[:, 14] == 0).sum() > (disaggregate_met(met)[:, 14] == 0).sum()
    with pytest.raises(ValueError):
        b.set_forcing_orig(met, kdown_zen=1)
    b.close()


@pytest.mark.gpu
def test_grid_loop_of_a_run_directory_on_the_device(tmp_path):
    """tables.grid_arguments + load_batch: a run directory with three grids (two classes) as one batch, BLUEWS on -- every
    cell's first day equals the oracle started from that grid's own one-grid join."""
    import ctypes as C
    import os
    from suews_b200 import OUT_FULL, SuewsBatch, tables
    from test_tables import _three_grid_run
    run = _three_grid_run(tmp_path)
    ga = tables.grid_arguments(run, 2004)
    met = fx.forcing(2004)[:288]
    b = tables.load_batch(SuewsBatch(len(ga["grids"]), len(ga["classes"])), ga)
    b.upload()
    ini = np.ascontiguousarray(ga["cbl_initial"])
    b.set_cbl_initial(ini)
    b.set_forcing(met)
    b.advance(0, met.shape[0], OUT_FULL)
    dp = C.POINTER(C.c_double)
    L = ol.lib()
    L.so_set_cbl_initial.argtypes = [dp, C.c_int]
    cbl_params, _ini = tables.read_cbl(os.path.join(run, "Input"))
    blocks = []
    for i, g in enumerate(ga["grids"]):
        d, _aux, rc = tables.run_arguments(run, g, 2004)
        d.update(cbl_params)
        d["cbluse"] = [int(rc["cbluse"])]
        L.so_set_cbl_initial(ini.ctypes.data_as(dp), ini.shape[0])
        cell = ol.Cell(d)
        try:
            ref = cell.run(met, want=("SUEWS",))
        finally:
            L.so_set_cbl_initial(None, 0)
        out = b.block("SUEWS", i, met.shape[0])
        blocks.append(out)
        assert np.array_equal(out[:, :5], ref["SUEWS"][:, :5])
        assert float(rel_err(out[:, 5:], ref["SUEWS"][:, 5:], "SUEWS").max()) <= TOL_DAY1, g
        for name in ("cbl_blh_m", "cbl_tm_K", "cbl_qm_kgkg", "soilstore_id", "state_id"):
            np.testing.assert_allclose(b.get_state(name)[i], cell.get_state(name), rtol=1e-10, atol=1e-300, err_msg="%s grid %d" % (name, g))
    assert not np.array_equal(blocks[0][:, 5:], blocks[1][:, 5:])      # other fractions, other fluxes
    b.close()


@pytest.mark.gpu
def test_reference_run_directory_to_reference_output_files(tmp_path):
    """examples/run_directory.py: the 2004 leg of the reference's golden run from its own input directory (table join,
    InitialState, CBLInput, hourly forcing disaggregated on the device), 105 408 steps on the GPU, written in the
    reference's file formats -- against the files the reference wrote: the boundary-layer file (8784 h x 17 columns), the
    DailyState file (366 days x 45) and the checkpoint the 2005 leg starts from."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("run_directory", os.path.join(ol.ROOT, "examples", "run_directory.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rep = mod.golden_2004(str(tmp_path))
    bad, nf, _same = rep["test1_2004_BL_60.txt"]
    assert nf == 8784 * 22 and bad <= 1e-4 * nf, (bad, nf)         # last-decimal roundings of values on a rounding boundary only
    bad, nf, _same = rep["test1_2004_DailyState.txt"]
    assert nf == 366 * 50 and bad <= 1e-3 * nf, (bad, nf)
    nbit, worst = rep["checkpoint"]
    assert nbit >= 50 and worst <= 1e-10, (nbit, worst)
