"""The table join (suews_b200/tables.py: InitializeSurfaceCharacteristics + SUEWS_Translate) against the reference's own
argument dumps of the same run: Test/BaseRun/2020b's input tables -> every table-derived argument of
test1_{2004,2005}_state_init.txt, exactly."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fixtures as fx  # noqa: E402

from suews_b200 import tables  # noqa: E402

RUN = os.path.join(fx.GOLDEN, "baserun_2020b")

# arguments of the dump that do not come from the tables: clock, forcing row, evolving state (InitialState / the previous
# year's end), and the arguments the 2020b dump predates
NOT_TABLE = {"iy", "id", "it", "imin", "isec", "dt_since_start", "tstep_prev", "avkdn", "avrh", "avu1", "fcld_obs", "lai_obs",
             "ldown_obs", "precip", "press_hpa", "qf_obs", "qh_obs", "qn1_obs", "qs_obs", "snowfrac_obs", "temp_c", "wu_m3", "xsmd",
             "qn1_av", "dqndt", "qn1_s_av", "dqnsdt", "snowfallcum", "snowalb", "icefrac", "snowwater", "snowdens", "snowfrac",
             "snowpack", "soilstore_id", "state_id", "gdd_id", "sdd_id", "lai_id", "tmin_id", "tmax_id", "lenday_id",
             "decidcap_id", "albdectr_id", "albevetr_id", "albgrass_id", "porosity_id", "hdd_id", "wuday_id", "tair_av"}


@pytest.mark.parametrize("year", [2004, 2005])
def test_table_join_lands_on_the_reference_argument_dump(year):
    rc = tables.read_runcontrol(os.path.join(RUN, "RunControl.nml"))
    tabs = tables.load_tables(os.path.join(RUN, "Input"))
    a = tables.site_arguments(tabs, rc, grid=1, year=year)
    aux = a.pop("_aux")
    dump = fx.state_init(year)
    dump = {k: v for k, v in dump.items() if k not in fx.DUMP_DEFAULTS}
    checked = 0
    for k, want in dump.items():
        if k in NOT_TABLE:
            continue
        assert k in a, "argument %s of the dump is not produced by the join" % k
        got = np.asarray(a[k], dtype=np.float64)
        want = np.asarray(want, dtype=np.float64)
        assert got.shape == want.shape, (k, got.shape, want.shape)
        if k == "alb":              # the vegetation albedos of the first step come from the daily state (translate.f95:1147)
            got, want = got[[0, 1, 5, 6]], want[[0, 1, 5, 6]]
        if k == "storedrainprm":    # StoreDrainPrm(6, DecTr) is the deciduous capacity of the day
            keep = np.ones(42, bool)
            keep[5 + 6 * 3] = False
            got, want = got[keep], want[keep]
        assert np.array_equal(got, want), (k, got, want)
        checked += 1
    assert checked >= 140, checked
    # the join produces nothing the argument list does not have (bar the arguments the 2020b dump predates)
    extra = set(a) - set(dump) - set(fx.DUMP_DEFAULTS) - {"irrfracpaved", "irrfracbldgs", "irrfracbsoil", "irrfracwater", "h_maintain",
                                                          "basetmethod"}
    assert not extra, extra
    assert aux["waterdepth"] > 0 and aux["grid"] == 1 and aux["year"] == year


def test_run_arguments_reproduce_the_first_step_of_2004():
    """run directory -> (table join + InitialState) -> the complete first-step argument set of the reference's 2004 leg."""
    a, aux, rc = tables.run_arguments(RUN, grid=1, year=2004,
                                      initial_conditions=os.path.join(RUN, "Input", "InitialConditionstest_2004.nml"))
    dump = fx.state_init(2004)
    skip = {"iy", "id", "it", "imin", "isec", "avkdn", "avrh", "avu1", "fcld_obs", "lai_obs", "ldown_obs", "precip", "press_hpa",
            "qf_obs", "qh_obs", "qn1_obs", "qs_obs", "snowfrac_obs", "temp_c", "wu_m3", "xsmd"}
    n = 0
    for k, want in dump.items():
        if k in skip or k in fx.DUMP_DEFAULTS:
            continue
        assert k in a, k
        assert np.array_equal(np.asarray(a[k], dtype=np.float64), np.asarray(want, dtype=np.float64)), (k, a[k], want)
        n += 1
    assert n >= 180, n
    assert rc["tstep"] == 300


def test_header_and_code_errors_are_the_references():
    tabs = tables.load_tables(os.path.join(RUN, "Input"))
    rc = tables.read_runcontrol(os.path.join(RUN, "RunControl.nml"))
    with pytest.raises(tables.TableError):
        tables.site_arguments(tabs, rc, grid=1, year=1999)               # no such SiteSelect row
    bad = tables.Table(os.path.join(RUN, "Input", "SUEWS_Soil.txt"), tables.HEADERS["SUEWS_Soil.txt"])
    bad.rows = bad.rows[:0]
    broken = dict(tabs, Soil=bad)
    with pytest.raises(tables.TableError) as e:
        tables.site_arguments(broken, rc, grid=1, year=2005)             # a code that matches no row: ErrorHint 57
    assert "57" in str(e.value)
    with pytest.raises(tables.TableError) as e:
        tables.Table(os.path.join(RUN, "Input", "SUEWS_Soil.txt"), tables.HEADERS["SUEWS_Conductance.txt"])   # ErrorHint 56
    assert "56" in str(e.value)


def test_estm_translation_of_the_baserun_tables():
    """ESTM facet parameters from SUEWS_ESTMCoefficients.txt + the SiteSelect row (SUEWS_Translate :393-668): roof, wall and
    internal elements are building class 801 (what suews_b200.synth.estm_example was written out from by hand), the ground
    facet is the area-weighted mean over the non-building surfaces, the layer counts stop at the first non-positive
    thickness.  (No reference output pins this: the BaseRun has ESTM off, so its dump does not carry these variables.)"""
    from suews_b200 import synth
    rc = tables.read_runcontrol(os.path.join(RUN, "RunControl.nml"))
    tabs = tables.load_tables(os.path.join(RUN, "Input"), estm=True)
    a = tables.site_arguments(tabs, rc, grid=1, year=2005)
    aux = a.pop("_aux")
    e = tables.estm_arguments(tabs, a, aux, tables.read_runcontrol(os.path.join(RUN, "Input", "ESTMinput.nml")))
    ex = synth.estm_example(a["surfacearea"][0])
    for k, want in ex.items():
        if k in ("AreaWall", "zground", "kground", "rground", "Nground"):
            continue
        got = np.asarray(e[k], dtype=np.float64)
        want = np.asarray(want, dtype=np.float64)
        n = int(e["N" + k[1:]][0]) if k[0] in "zkr" and k[1:] in ("ibld", "wall", "roof") else want.size
        assert np.allclose(got[:n], want[:n], rtol=1e-12, atol=0), (k, got, want)
    E = tabs["ESTMCoefficients"]
    sfr = np.array(a["sfr"])
    fg = sfr.sum() - sfr[1]
    codes = [807.0, None, 810.0, 811.0, 812.0, 809.0, 813.0]      # paved: class 807 with fraction 1; the others by ESTMCode
    z1 = sum(E.get(E.lookup(c, "t"), "Surf_thick1") * sfr[i] / fg for i, c in enumerate(codes) if c is not None)
    assert abs(e["zground"][0] - z1) < 1e-15 and e["Nground"] == [4] and e["Nroof"] == [3]
    assert e["AreaWall"] == [-999.0]        # not given in this SiteSelect row: ESTM is off in the BaseRun


def test_run_arguments_with_estm_switched_on(tmp_path):
    """The same run directory with StorageHeatMethod = 4 in RunControl.nml: run_arguments adds MODULE ESTM_data's parameters
    and, given the first surface-temperature row, the start temperatures -- every ESTM field of the library is covered."""
    import shutil
    from suews_b200 import batch
    run = tmp_path / "run"
    shutil.copytree(RUN, run)
    rcp = run / "RunControl.nml"
    rcp.write_text(rcp.read_text().replace("StorageHeatMethod=1", "StorageHeatMethod=4"))
    ts_row0 = [2005, 1, 0, 5, 20.0, 1.0, 2.0, 1.5, 0.5, 0.3, 0.8, 0.1, 0.2]
    a, aux, rc = tables.run_arguments(str(run), grid=1, year=2004, estm_ts_row0=ts_row0,
                                      initial_conditions=os.path.join(RUN, "Input", "InitialConditionstest_2004.nml"))
    assert a["storageheatmethod"] == [4] and a["nroof"] == [3] and a["nground"] == [4]
    names = {nm.lower() for nm in batch.estm_names()} - {"cbluse", "entrainmenttype", "qh_choice", "isubs", "co2_included",
                                                          "initialdata_use", "sondeflag", "wsb"}
    names = {n for n in names if not n.startswith(("anohm_", "cbl_"))} & {nm.lower() for nm, _k, _n in batch.all_fields()}
    missing = sorted(n for n in names if n not in a)
    assert not missing, missing
    assert len(a["tw_4"]) == 20 and a["tievolve"][0] == 20.0 + float(np.float32(273.15))


def _three_grid_run(tmp_path):
    """the golden run directory with three grids in 2004: grid 2 differs in site values (fractions, building height),
    grid 3 in a value the grids of a class share (the daylight-saving start day)"""
    import shutil
    run = tmp_path / "run3"
    shutil.copytree(RUN, run)
    p = run / "Input" / "SUEWS_SiteSelect.txt"
    lines = p.read_text(encoding="latin-1").splitlines()
    names = lines[1].split()
    i04 = next(i for i, ln in enumerate(lines) if ln.split()[:2] == ["1", "2004"])

    def variant(grid, **kw):
        tok = lines[i04].split("!")[0].split()
        tok[0] = str(grid)
        for k, v in kw.items():
            tok[names.index(k)] = str(v)
        return "   ".join(tok)
    lines[i04 + 1:i04 + 1] = [variant(2, Fr_Paved=0.1, Fr_Grass=0.2, H_Bldgs=22.0), variant(3, StartDLS=90)]
    p.write_text("\n".join(lines) + "\n", encoding="latin-1")
    return str(run)


def test_grid_loop_of_a_run_directory_as_classes_and_cells(tmp_path):
    """grid_arguments: every SiteSelect row of the year is a cell, rows with equal class-kind values share a class; each
    grid's values are those of the one-grid join."""
    run = _three_grid_run(tmp_path)
    ga = tables.grid_arguments(run, 2004)
    assert ga["grids"] == [1, 2, 3] and list(ga["class_of_grid"]) == [0, 0, 1] and len(ga["classes"]) == 2
    assert ga["clock"] == (0, 300) and ga["cbl_initial"].shape[1] == 8 and ga["classes"][0]["CBLuse"][0] == 1
    assert ga["classes"][0]["startDLS"][0] == 85 and ga["classes"][1]["startDLS"][0] == 90
    differ = [k for k in ga["classes"][0] if not np.array_equal(ga["classes"][0][k], ga["classes"][1][k])]
    assert differ == ["startDLS"]
    assert list(ga["cells"]["sfr"][:, 0]) == [0.2, 0.1, 0.2] and list(ga["cells"]["bldgH"][:, 0]) == [15.0, 22.0, 15.0]
    for i, g in enumerate(ga["grids"]):
        a, _aux, _rc = tables.run_arguments(run, g, 2004)
        for d in (ga["cells"], ga["state"]):
            for nm, v in d.items():
                assert np.array_equal(v[i], np.asarray(a[nm.lower()], dtype=np.float64).ravel()), (g, nm)
    # (state that depends on the fractions follows its grid: a surface that is absent has no store)
    assert ga["state"]["soilstore_id"].shape == (3, 7)
    with pytest.raises(tables.TableError):
        tables.grid_arguments(run, 1999)

    class Recorder:          # load_batch's calls, without a device
        def __init__(self):
            self.calls = []

        def __getattr__(self, name):
            return lambda *a: self.calls.append((name, a))
    r = tables.load_batch(Recorder(), ga)
    kinds = [c[0] for c in r.calls]
    assert kinds.count("set_class_param") == 2 * len(ga["classes"][0]) and kinds.count("set_cell_class") == 1
    assert kinds.count("set_cell_param") == len(ga["cells"]) and kinds.count("set_state") == len(ga["state"]) and kinds[-1] == "set_clock"


def test_grid_loop_picks_the_initial_conditions_file_as_the_reference_does(tmp_path):
    """MultipleInitFiles = 1: InitialConditions<FileCode><grid>_<year>.nml per grid (suews_ctrl_init.f95, InitialState's file
    name); 0: the shared InitialConditions<FileCode>_<year>.nml; a function grid -> path overrides both."""
    import shutil
    run = _three_grid_run(tmp_path)
    inp = os.path.join(run, "Input")
    shared = os.path.join(inp, "InitialConditionstest_2004.nml")
    text = open(shared).read()
    assert "SoilstorePavedState" in text
    for g, v in ((1, 100.0), (2, 110.0), (3, 120.0)):
        lines = [("SoilstorePavedState=%g" % v) if ln.strip().lower().startswith("soilstorepavedstate") else ln for ln in text.splitlines()]
        open(os.path.join(inp, "InitialConditionstest%d_2004.nml" % g), "w").write("\n".join(lines) + "\n")
    rcp = os.path.join(run, "RunControl.nml")
    rc_text = open(rcp).read()
    assert "MultipleInitFiles=0" in rc_text.replace(" ", "")
    ga0 = tables.grid_arguments(run, 2004)
    assert len(set(ga0["state"]["soilstore_id"][:, 0])) == 1
    open(rcp, "w").write(rc_text.replace("MultipleInitFiles=0", "MultipleInitFiles=1").replace("MultipleInitFiles = 0", "MultipleInitFiles=1"))
    ga1 = tables.grid_arguments(run, 2004)
    assert list(ga1["state"]["soilstore_id"][:, 0]) == [100.0, 110.0, 120.0]
    ga2 = tables.grid_arguments(run, 2004, initial_conditions=lambda g: os.path.join(inp, "InitialConditionstest%d_2004.nml" % (4 - g)))
    assert list(ga2["state"]["soilstore_id"][:, 0]) == [120.0, 110.0, 100.0]
    ga3 = tables.grid_arguments(run, 2004, grids=[3, 1])
    assert ga3["grids"] == [3, 1] and list(ga3["class_of_grid"]) == [0, 1] and list(ga3["state"]["soilstore_id"][:, 0]) == [120.0, 100.0]
    shutil.rmtree(run)
