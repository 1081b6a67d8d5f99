"""Pin the CPU oracle against the reference's own golden run.

Test/BaseRun/2020b: the complete SUEWS_cal_Main argument dump at the first step of 2005
(test1_2005_state_init.txt) is replayed for 1440 five-minute steps; the hourly aggregates must
reproduce the reference's text outputs to the printed precision: SUEWS (82 cols), snow (98),
RSL (133), DailyState (45), SOLWEIG (26).  The 2004 leg (105 408 steps) is pinned by the year-long
DailyState (366 x 45) and SOLWEIG (8784 x 26) goldens, the full-year outputs that survive in the reference
checkout.
"""
import numpy as np
import pytest

import fixtures as fx
import oracle_lib as ol


@pytest.fixture(scope="module")
def run2005():
    cell = ol.Cell(fx.state_init(2005))
    return cell.run(fx.forcing(2005), want=("SUEWS", "snow", "RSL", "DailyState", "SOLWEIG"))


def test_forcing_disaggregation_matches_dump():
    """first disaggregated row == the forcing scalars in the argument dump, bit for bit"""
    d = fx.state_init(2005)
    m = fx.forcing(2005)
    assert m.shape == (1440, 24)
    cols = dict(qn1_obs=4, qs_obs=7, qf_obs=8, avu1=9, avrh=10, temp_c=11, press_hpa=12, precip=13, avkdn=14,
                snowfrac_obs=15, ldown_obs=16, fcld_obs=17, wu_m3=18, xsmd=19, lai_obs=20)
    for k, c in cols.items():
        assert m[0, c] == d[k][0], k
    assert [int(x) for x in m[0, :4]] == [d["iy"][0], d["id"][0], d["it"][0], d["imin"][0]]
    # and the reference's own 4-6 decimal text dump of the disaggregated series
    ref = np.loadtxt(fx.GOLDEN + "/forcing_2005_5min.txt", skiprows=1)
    ref[:, 12] *= 10
    assert np.all(np.abs(m - ref) <= 5.1e-5)


@pytest.mark.parametrize("group,fname", [("SUEWS", "ref_2005_SUEWS_60.txt"), ("snow", "ref_2005_snow_60.txt"),
                                         ("RSL", "ref_2005_RSL_60.txt")])
def test_oracle_matches_reference_hourly(run2005, group, fname):
    hdr, ref = fx.load_ref_table(fname)
    ours = fx.aggregate_hourly(run2005[group], group)
    assert ours.shape == (120, ref.shape[1] - 5)
    tol = fx.text_tolerance(group, ours.shape[1])
    for c in range(ours.shape[1]):
        r = ref[:, 5 + c]
        ok = np.isnan(r) | (np.abs(ours[:, c] - r) <= tol[c] + 1e-9 * np.abs(np.nan_to_num(r)))
        assert ok.all(), (group, hdr[5 + c], int(np.argmin(ok)), ours[~ok, c][:3], r[~ok][:3])
    # integer/flag outputs are exact: datetime columns
    assert np.array_equal(run2005[group][11::12, :4], ref[:, :4])


def _check_daily(block, fname):
    hdr, ref = fx.load_ref_table(fname)
    rows = block[block[:, 5] != -999]
    assert rows.shape == ref.shape
    assert np.array_equal(rows[:, :2], ref[:, :2])
    for c in range(5, 50):
        r = ref[:, c]
        ok = np.isnan(r) | (np.abs(rows[:, c] - r) <= 0.6e-4 + 1e-9 * np.abs(np.nan_to_num(r)))
        assert ok.all(), (hdr[c], int(np.argmin(ok)), rows[~ok, c][:3], r[~ok][:3])
    assert np.array_equal(rows[:, 10], ref[:, 10])  # DaysSR: integer-valued, exact


def test_oracle_dailystate_2005(run2005):
    _check_daily(run2005["DailyState"], "ref_2005_DailyState.txt")


@pytest.fixture(scope="module")
def run2004():
    cell = ol.Cell(fx.state_init(2004))
    return cell.run(fx.forcing(2004), want=("DailyState", "SOLWEIG"))


@pytest.mark.slow
def test_oracle_dailystate_full_year_2004(run2004):
    _check_daily(run2004["DailyState"], "ref_2004_DailyState.txt")


def _check_solweig(block, fname, nhours):
    """SOLWEIG point diagnostics (26 columns, all hourly means).  I0 is only assigned in daylight by the
    reference (solweig.f95:270); at night its output is stack residue, so hours with any night step are masked."""
    hdr, ref = fx.load_ref_table(fname)
    ours = fx.aggregate_hourly(block, "SOLWEIG")
    assert ours.shape == (nhours, 26) and ref.shape == (nhours, 31)
    assert np.array_equal(block[11::12, :4], ref[:, :4])
    night = (block[:, 5 + 1].reshape(nhours, 12) <= 0).any(axis=1)          # altitude <= 0 at some step of the hour
    for c in range(26):
        r = ref[:, 5 + c]
        ok = np.isnan(r) | (np.abs(ours[:, c] - r) <= 0.6e-6 + 1e-9 * np.abs(np.nan_to_num(r)))
        if hdr[5 + c] == "I0":
            ok = ok | night
        assert ok.all(), (hdr[5 + c], int(np.argmin(ok)), ours[~ok, c][:3], r[~ok][:3])


def test_oracle_solweig_2005(run2005):
    _check_solweig(run2005["SOLWEIG"], "ref_2005_SOLWEIG_60.txt", 120)


@pytest.mark.slow
def test_oracle_solweig_full_year_2004(run2004):
    _check_solweig(run2004["SOLWEIG"], "ref_2004_SOLWEIG_60.txt", 8784)


def test_sun_position_golden_quirk():
    """The reference evaluates the sun 150 'days' early (dectime - tstep/2 with integer seconds,
    driver.f95:689): Zenith 40.55 deg at 13:00 on 1 Jan at 57.77N is in the golden file."""
    hdr, ref = fx.load_ref_table("ref_2005_SUEWS_60.txt")
    zen_col = hdr.index("Zenith")
    row = ref[12]  # 13:00 stamp
    assert int(row[2]) == 13
    dectime = (1 - 1) + 13 / 24.0
    az, zen = ol.sun_position(2005.0, dectime - 150, 2.0, 57.77, 11.87, 15.0)
    assert abs(zen - row[zen_col]) < 0.6e-4


# ---- the sharpest pin in the reference repository: the InitialConditions namelists NextInitial wrote at the end of
# the 2004 leg (105 408 steps) and of the 2005 leg (1 440 steps) of the golden run, list-directed = 17 significant
# digits = the exact doubles of 54 state variables (surface and soil stores, snow pack / fraction / density / water,
# albedos, LAI, porosity, the 5-day mean temperature).  The oracle reproduces every one of them BIT FOR BIT.
def _nml_pairs(cell, nml):
    from suews_b200 import checkpoint
    v = checkpoint.state_vector(cell.get_state)
    names = (["DaysSinceRain", "Temp_C0", "GDD_1_0", "GDD_2_0", "LAIinitialEveTr", "LAIinitialDecTr", "LAIinitialGrass",
              "AlbEveTr0", "AlbDecTr0", "AlbGrass0", "DecidCap0", "Porosity0"] +
             ["SoilStore%sState" % s for s in checkpoint.SURF[:6]] + ["%sState" % s for s in checkpoint.SURF] +
             ["SnowWater%sState" % s for s in checkpoint.SURF] + ["SnowPack%s" % s for s in checkpoint.SURF] +
             ["SnowFrac%s" % s for s in checkpoint.SURF] + ["SnowDens%s" % s for s in checkpoint.SURF] + ["SnowAlb0"])
    assert len(names) == v.size == 54
    return names, v, np.array([nml[k] for k in names])


@pytest.fixture(scope="module")
def cell_after_2004():
    cell = ol.Cell(fx.state_init(2004))
    cell.run(fx.forcing(2004), want=())
    return cell


def test_state_after_2005_leg_is_the_reference_s_bit_for_bit():
    from suews_b200 import checkpoint
    cell = ol.Cell(fx.state_init(2005))
    cell.run(fx.forcing(2005), want=())
    names, ours, ref = _nml_pairs(cell, checkpoint.read(fx.GOLDEN + "/ref_InitialConditions_2005_EndofRun.nml"))
    bad = [(n, a, b) for n, a, b in zip(names, ours, ref) if a != b]
    assert not bad, bad


def test_state_after_full_year_2004_is_the_reference_s_bit_for_bit(cell_after_2004):
    from suews_b200 import checkpoint
    names, ours, ref = _nml_pairs(cell_after_2004, checkpoint.read(fx.GOLDEN + "/ref_InitialConditions_2005.nml"))
    bad = [(n, a, b) for n, a, b in zip(names, ours, ref) if a != b]
    assert not bad, bad


@pytest.mark.parametrize("fname,year", [("ref_InitialConditions_2005_EndofRun.nml", 2005), ("ref_InitialConditions_2005.nml", 2004)])
def test_checkpoint_writer_reproduces_the_reference_file(fname, year, cell_after_2004):
    """NextInitial's list-directed text (suews_ctrl_init.f95:2022-2149) from the library's writer: byte for byte"""
    from suews_b200 import checkpoint
    if year == 2004:
        cell = cell_after_2004
    else:
        cell = ol.Cell(fx.state_init(2005))
        cell.run(fx.forcing(2005), want=())
    text = checkpoint.write_text(checkpoint.state_vector(cell.get_state), snow_use=1)
    assert text == open(fx.GOLDEN + "/" + fname).read()


def test_rsl_humidity_diagnostics_are_nan_exactly_where_the_reference_divides_zero_by_zero():
    """RSLProfile forms the in-canopy humidity profile as (q_h*EXP(..) - q_h)/qStar_RSL (suews_phys_rslprof.f95:285) with
    q_h = Scc*qStar_RSL/(beta*f) (:281) and qStar_RSL = -(qe/lv*avdens)/UStar (:266): when QE is exactly zero (saturated
    air over frozen, snow-covered surfaces) that is 0/0, and the reference's q profile, q2 and RH2 are NaN for that step
    (T2 and U10 are not).  The restatement keeps it; this pins where it happens: on those steps and on no other."""
    from suews_b200 import synth
    base = fx.state_init(2005)
    met = fx.forcing(2005)[:288].copy()
    d = synth.base_with_scheme(base, "baserun")
    site = synth.cell_site_arrays(12, 1, {k.lower(): v for k, v in base.items()})   # a synthetic cell that saturates (RH + dRH >= 100)
    for k in ("sfr", "lat", "lng", "bldgH", "EveTreeH", "DecTreeH", "Z"):
        d[k.lower()] = np.atleast_1d(site[k][0])
    met[:, 11] += site["dTemp_C"][0]
    met[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), met[:, 10] + site["dRH"][0]))
    out = ol.Cell(d).run(met, want=("SUEWS",))["SUEWS"][:, 5:]
    q2_nan, rh2_nan = np.isnan(out[:, 79]), np.isnan(out[:, 81])
    assert q2_nan.sum() > 0                                     # the case exists in this cell's first day
    assert np.array_equal(q2_nan, rh2_nan)
    assert np.array_equal(q2_nan, out[:, 9] == 0.0)             # exactly the steps with QE == 0
    assert not np.isnan(out[:, 78]).any() and not np.isnan(out[:, 80]).any()   # T2, U10 stay finite
    other = [c for c in range(82) if c not in (79, 81)]
    assert not np.isnan(out[:, other]).any()


def test_op_counter_build_computes_the_same_numbers():
    """oracle/libsuews_oracle_count.so (`make count`: the same sources as C++ with a counting `double`) must be
    the same arithmetic: identical output blocks, non-zero counts, and the hoisted count below the as-written one."""
    import ctypes as C
    import os
    import subprocess
    d = fx.state_init(2005)
    met = fx.forcing(2005)[:48]
    plain = ol.Cell(d).run(met, want=("SUEWS", "RSL"))
    subprocess.check_call(["make", "-C", ol.ORACLE_DIR, "-s", "count"])
    keep = (ol.LIB_PATH, ol._lib)
    try:
        ol.LIB_PATH = os.path.join(ol.ORACLE_DIR, "libsuews_oracle_count.so")
        ol._lib = None
        L = ol.lib()
        L.so_opcount_start.argtypes = [C.c_int]
        L.so_opcount_stop.argtypes = [C.POINTER(C.c_ulonglong)]
        counts = {}
        for mode in (1, 0):
            buf = (C.c_ulonglong * 32)()
            L.so_opcount_start(mode)
            got = ol.Cell(d).run(met, want=("SUEWS", "RSL"))
            L.so_opcount_stop(buf)
            counts[mode] = np.array(list(buf), dtype=float)
            for k in plain:
                assert np.array_equal(plain[k], got[k], equal_nan=True), k
        assert counts[1][:4].sum() > 0 and counts[0][:4].sum() > counts[1][:4].sum()      # add/mul/div/sqrt of the flux path
        assert np.array_equal(counts[0][16:], counts[1][16:])                             # diagnostics are counted alike
    finally:
        ol.LIB_PATH, ol._lib = keep


@pytest.mark.parametrize("year", [2005, 2004])
def test_bluews_reproduces_the_reference_bl_files(year):
    """BLUEWS (suews_phys_bluews.f95) run after every step from the oracle's QH / QE / UStar and moist-air values: the hourly
    means of its line are the reference's test1_<year>_BL_60.txt to the printed digit -- 120 h of 2005 and all 8784 h of
    2004 (nocturnal resets from CBL_initial_data.txt, daytime Runge-Kutta growth with EntrainmentType 4)."""
    import ctypes as C
    dp = C.POINTER(C.c_double)
    L = ol.lib()
    L.so_set_cbl_initial.argtypes = [dp, C.c_int]
    L.so_set_bl_block.argtypes = [dp]
    params, ini = fx.cbl_inputs()
    ini = np.ascontiguousarray(ini)
    d = fx.state_init(year)
    d.update(params)
    met = fx.forcing(year)
    bl = np.full((met.shape[0], 22), -999.0, order="F")
    L.so_set_cbl_initial(ini.ctypes.data_as(dp), ini.shape[0])
    L.so_set_bl_block(bl.ctypes.data_as(dp))
    try:
        ol.Cell(d).run(met, want=("SUEWS",))
    finally:
        L.so_set_bl_block(None)
        L.so_set_cbl_initial(None, 0)
    _hdr, ref = fx.load_ref_table("ref_%d_BL_60.txt" % year)
    nh = ref.shape[0]
    assert nh == (120 if year == 2005 else 8784)
    assert np.array_equal(bl[11:nh * 12:12, :4], ref[:, :4])            # the hour's last step carries its date columns
    agg = bl[:nh * 12, 5:].reshape(nh, 12, 17).mean(axis=1)
    tol = fx.text_tolerance("BL", 17)
    assert np.all(np.abs(agg - ref[:, 5:]) <= tol), np.abs(agg - ref[:, 5:]).max(axis=0)
    assert agg[:, 0].max() > 1000 and agg[:, 0].min() == 200.0            # the slab really grows by day and resets at night


def test_bluews_start_values_diagnosed_from_the_forcing():
    """InitialData_use = 1 (suews_phys_bluews.f95:570-581): on a day the initial-data file lists, height and gradients come
    from the file and theta, q from the night's line through psih (the value AerodynamicResistance left behind); on any
    other day the night's 200 m values stay.  Checked against the formulas written out here from the oracle's own output
    columns (UStar, Lob, zdm, QH, QE and the BL line)."""
    import ctypes as C
    dp = C.POINTER(C.c_double)
    L = ol.lib()
    L.so_set_cbl_initial.argtypes = [dp, C.c_int]
    L.so_set_bl_block.argtypes = [dp]
    L.so_stab_psi_heat.restype = C.c_double
    L.so_stab_psi_heat.argtypes = [C.c_int, C.c_double]
    params, ini = fx.cbl_inputs()
    ini = np.ascontiguousarray(ini[:1])                       # day 1 only: day 2 finds no row of its own
    d = fx.state_init(2005)
    d.update(params)
    d["initialdata_use"] = [1]
    met = fx.forcing(2005)[:288 + 36]
    n = met.shape[0]

    def run(nsteps):
        bl = np.full((nsteps, 22), -999.0, order="F")
        L.so_set_cbl_initial(ini.ctypes.data_as(dp), ini.shape[0])
        L.so_set_bl_block(bl.ctypes.data_as(dp))
        cell = ol.Cell(d)
        try:
            out = cell.run(met[:nsteps], want=("SUEWS",))["SUEWS"]
        finally:
            L.so_set_bl_block(None)
            L.so_set_cbl_initial(None, 0)
        return cell, out, bl
    for nsteps in (36, n):                                    # 03:00 of day 1 (a night step), 03:00 of day 2
        cell, su, bl = run(nsteps)
        assert met[nsteps - 1, 14] < 5                        # Kdown: night
        get = lambda k: float(cell.get_state(k)[0])           # noqa: E731
        row, line, s = ini[0], bl[-1], su[-1]
        assert line[5] == 200.0 and line[6] == row[6] and line[7] == row[7]
        assert get("cbl_gamq_kgkgm") == row[3] / float(np.float32(1000.0))
        if nsteps == 36:
            k, c2k = float(np.float32(0.4)), float(np.float32(273.16))
            Temp_C, avrh, qh, qe, press, ustar, avdens, avcp = line[10], line[11], line[12], line[13], line[14], line[16], line[17], line[19]
            assert qh == s[5 + 8] and qe == s[5 + 9] and ustar == s[5 + 54]
            psih = L.so_stab_psi_heat(3, (d["z"][0] - s[5 + 53]) / s[5 + 55])
            gas = float(np.float32(8.31451) / np.float32(0.028965))
            tm_zm = (Temp_C + c2k) * (1000 / press) ** (gas / avcp)
            tm = tm_zm - psih * qh / (k * ustar * avcp * avdens)
            assert get("cbl_blh_m") == row[1] and get("cbl_gamt_Km") == row[2]
            assert abs(get("cbl_tm_K") - tm) <= 1e-12 * tm and get("cbl_tp_K") == get("cbl_tm_K")
            assert abs(tm - row[6]) < 5.0 and tm != row[6]     # a correction of the file's theta, not the file's theta
            assert get("cbl_qp_kgkg") == get("cbl_qm_kgkg") and 0 < get("cbl_qm_kgkg") < 0.03
        else:
            assert get("cbl_blh_m") == 200.0 and get("cbl_tm_K") == row[6] and get("cbl_tp_K") == row[4]
            assert get("cbl_qp_kgkg") <= row[5] / 1000 and get("cbl_qm_kgkg") <= row[7] / 1000
    # InitialData_use = 0 has no behaviour in the reference (it reads a table it never allocated): refused
    d["initialdata_use"] = [0]
    cell = ol.Cell(d)
    L.so_set_cbl_initial(ini.ctypes.data_as(dp), ini.shape[0])
    try:
        with pytest.raises(Exception):
            cell.run(met[:2], want=("SUEWS",))
    finally:
        L.so_set_cbl_initial(None, 0)
