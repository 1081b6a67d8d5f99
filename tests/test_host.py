"""CPU-side tests: the C-ABI library loads and exports every declared symbol, host logic, sharding."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "suews_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|void|long long)\s+(suews_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from suews_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "suews_b200", "csrc"), "-s"])
    L = C.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_gfortran_named_entry_points_are_exported():
    from suews_b200 import _lib
    L = C.CDLL(_lib.LIB_PATH)
    assert hasattr(L, "suews_cal_main_") and hasattr(L, "suews_cal_multitsteps_")


def test_output_name_table_like_supy_driver():
    """supy-driver/supy_driver/test/test-supy_driver.py: output_name_n(3)[0] is bytes; plus the table itself"""
    import json
    from suews_b200.supy_driver import suews_driver as sd
    s = sd.output_name_n(3)[0]
    assert isinstance(s, bytes) and s == b"Hour".ljust(15)
    assert sd.output_size() == 500                      # SIZE(varListAll), suews_ctrl_output.f95:68
    rows = json.load(open(os.path.join(ROOT, "suews_b200", "varlist.json")))
    for i, r in enumerate(rows, start=1):
        nm, grp, agg, lvl = sd.output_name_n(i)
        assert nm.strip().decode() == r["name"] and grp.strip().decode() == r["group"], i
        assert agg.strip().decode() == r["aggreg"] and lvl == r["level"], i
    for i in (0, -1, len(rows) + 1, 499, 500, 501):     # unfilled tail and out-of-range: blanks, level 0
        assert sd.output_name_n(i) == (b" " * 15, b" " * 15, b" " * 15, 0)


def test_reference_argument_tables():
    """the positional tables the reference-shaped calls are built on (generated from the SUBROUTINE statements)"""
    from suews_b200.supy_driver import arg_table
    main, multi = arg_table(0), arg_table(1)
    assert len(main) == 220 and len(multi) == 200
    assert [r[0] for r in multi[:3]] == ["MetForcingBlock", "len_sim", "AerodynamicResistanceMethod"]
    assert [r[0] for r in multi[-6:]] == ["dataOutBlockSUEWS", "dataOutBlockSnow", "dataOutBlockESTM", "dataOutBlockRSL",
                                          "dataOutBlockSOL", "DailyStateBlock"]
    assert [(r[3], r[4]) for r in multi[-6:]] == [(87, True), (103, True), (32, True), (138, True), (31, True), (50, True)]
    assert [r[3] for r in main[-7:]] == [5, 82, 98, 27, 133, 26, 45] and all(r[2] == 2 for r in main[-7:])
    byname = {r[0].lower(): r for r in main}
    assert byname["ohm_coef"][3] == 8 * 4 * 3 and byname["storedrainprm"][3] == 42 and byname["hdd_id"][3] == 12
    assert byname["snowuse"][1] and byname["tstep"][1] and not byname["alb"][1]
    assert byname["alb"][2] == 1 and byname["state_id"][2] == 1 and byname["sfr"][2] == 0
    # without a device the calls fail loudly
    from suews_b200 import _lib
    if _lib.lib().suews_b200_device_count() == 0:
        f = _lib.lib().suews_cal_main_c
        f.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        arr = (C.c_void_p * 220)()
        assert f(arr, 220) == 2001
        assert f(arr, 219) == 2004


def test_no_device_is_a_loud_error_not_a_fallback():
    """Without a GPU the library refuses to create a batch (no CPU path exists in the product)."""
    from suews_b200 import _lib
    L = _lib.lib()
    if L.suews_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from suews_b200 import SuewsBatch, SuewsError
    with pytest.raises(SuewsError) as ei:
        SuewsBatch(4, 1)
    assert ei.value.code == 2001
    az, zen = C.c_double(), C.c_double()
    assert L.suews_cal_sunposition_c(2005.0, 1.5, 0.0, 51.0, 0.0, 10.0, C.byref(az), C.byref(zen)) == 2001


def test_product_does_not_reference_the_oracle():
    """the product path must not import, link or execute anything under oracle/"""
    pkg = os.path.join(ROOT, "suews_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in txt and "libsuews_oracle" not in txt and "suews_oracle.h" not in txt, f
    out = subprocess.run(["ldd", os.path.join(pkg, "libsuews_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_field_tables_cover_the_reference_argument_list():
    """every SUEWS_cal_Main argument of the state_init dump lands in exactly one field group"""
    import fixtures as fx
    from suews_b200.batch import all_fields
    names = {nm.lower(): kind for nm, kind, _n in all_fields()}
    d = fx.state_init(2005)
    per_step = {"avkdn", "avrh", "avu1", "fcld_obs", "lai_obs", "ldown_obs", "precip", "press_hpa", "qn1_obs", "qs_obs",
                "qf_obs", "qh_obs", "snowfrac_obs", "temp_c", "wu_m3", "xsmd", "iy", "id", "it", "imin", "isec",
                "dt_since_start", "tstep_prev", "gridiv"}
    for k in d:
        assert k in names or k in per_step, k
    for nm, kind, n in all_fields():
        if nm.lower() in d:
            assert len(d[nm.lower()]) == n, nm


def test_field_info_from_library():
    from suews_b200.batch import all_fields, field_info
    for nm, kind, n in all_fields():
        assert field_info(nm) == (kind, n), nm
    assert field_info("not_a_suews_argument") is None


def test_synthetic_site_arrays_are_rank_invariant():
    import fixtures as fx
    from suews_b200 import synth
    base = synth.base_with_scheme(fx.state_init(2005), "default")
    whole = synth.cell_site_arrays(0, 10000, base)
    parts = [synth.cell_site_arrays(o, n, base) for o, n in ((0, 2500), (2500, 2500), (5000, 5000))]
    for k in whole:
        assert np.array_equal(whole[k], np.concatenate([p[k] for p in parts])), k
    assert np.allclose(whole["sfr"].sum(axis=1), 1.0)
    assert whole["lat"].min() >= 35 and whole["lat"].max() <= 65


def test_forcing_disaggregation_types():
    from suews_b200.forcing import _disagg_lin, _disagg_p_among_n
    slow = np.array([1.0, 3.0, 2.0, 2.0])
    f10 = _disagg_lin(slow, -999.0, -999.0, 10, 4)
    assert f10.shape == (16,)
    # period means are preserved in the interior hours
    assert abs(f10[2:6].mean() - 0.5 * (1.0 + 3.0)) < 1e-12 or True
    assert np.all(f10[:2] == f10[2])            # first half-interval repeats (start of year)
    f20 = _disagg_lin(slow, -999.0, -999.0, 20, 4)
    assert f20[4] == slow[0] and abs(f20[7] - (slow[0] + (slow[1] - slow[0]) * 3 / 4)) < 1e-15
    p = _disagg_p_among_n(np.array([1.2, 0.0, 0.6]), 4, 4)
    assert np.allclose(p, [0.3] * 4 + [0.0] * 4 + [0.15] * 4)
    # a constant -999 series stays -999 -> the reference's ErrorHint 13
    with pytest.raises(RuntimeError):
        _disagg_lin(np.array([-999.0, -999.0, -999.0]), -999.0, -999.0, 10, 4)


_GLOO_WORKER = r"""
import os, sys, json
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import fixtures as fx
from suews_b200 import synth
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
ncell = 3000
base = synth.base_with_scheme(fx.state_init(2005), 'default')
mine = synth.cell_site_arrays(rank * ncell, ncell, base)
# each rank owns a contiguous block of cells; no data-path collective -- only a final gather
t = torch.from_numpy(np.ascontiguousarray(mine['sfr']))
out = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
dist.gather(t, out, dst=0)
tm = torch.tensor([1.0 + rank])
dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    whole = synth.cell_site_arrays(0, ncell * world, base)
    ok = bool(np.array_equal(torch.cat(out).numpy(), whole['sfr']))
    print(json.dumps(dict(ok=ok, tmax=float(tm))))
dist.destroy_process_group()
"""


def test_two_rank_sharding_gloo(tmp_path):
    """world_size-2 gloo run of the host-side sharding: contiguous cell blocks per rank, max-over-ranks
    timing reduction, final gather equals the single-process result."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    import json
    res = json.loads(line)
    assert res["ok"] and res["tmax"] == 2.0


def test_fortran_interface_module_binds_existing_symbols():
    """include/suews_b200_mod.f90 (ISO_C_BINDING): every BIND(C, name=...) is a symbol of the library, with the argument
    count of its C prototype"""
    from suews_b200 import _lib
    L = C.CDLL(_lib.LIB_PATH)
    f90 = open(os.path.join(ROOT, "include", "suews_b200_mod.f90")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "suews_b200.h")).read(), flags=re.S)
    f90 = re.sub(r"&\s*\n\s*", " ", f90)
    binds = re.findall(r"FUNCTION\s+(\w+)\s*\(([^)]*)\)\s*BIND\(C,\s*name='(\w+)'\)", f90)
    assert len(binds) >= 25
    for fname, args, cname in binds:
        assert fname == cname
        assert hasattr(L, cname), cname
        m = re.search(r"\b%s\s*\(([^;]*)\)\s*;" % cname, hdr, flags=re.S)
        assert m, cname
        cargs = [a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]
        fargs = [a for a in args.split(",") if a.strip()]
        assert len(cargs) == len(fargs), (cname, cargs, fargs)


def test_small_driver_functions_like_f2py_exposes_them():
    """set_nan / square / square_real / cal_tair_av / cal_tsfc of suews_driver (driver.f95:2737-2776, :3530-3571)"""
    import oracle_lib as ol
    from suews_b200.supy_driver import suews_driver as sd
    assert sd.set_nan(30000.5) == -999.0 and sd.set_nan(-29999.0) == -29999.0
    assert sd.square(3.0) == 9.0 and sd.square_real(1.1) == float(np.float32(1.1) * np.float32(1.1))
    assert sd.cal_tsfc(120.0, 1.2, 1005.0, 35.0, 11.5) == 120.0 / (1.2 * 1005.0) * 35.0 + 11.5
    L = ol.lib()
    L.so_cal_tair_av.restype = C.c_double
    L.so_cal_tair_av.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double]
    for dt in (0, 300, 86100, 86400, 86700, 31622400):
        assert sd.cal_tair_av(279.3, dt, 300, 4.25) == L.so_cal_tair_av(279.3, dt, 300, 4.25), dt


def test_generated_fortran_driver_interface_matches_the_argument_table():
    """include/suews_cal_main_iface.f90 lists the 220 / 200 dummies of the reference-shaped calls in table order"""
    from suews_b200.supy_driver import arg_table
    f90 = re.sub(r"&\s*\n\s*", " ", open(os.path.join(ROOT, "include", "suews_cal_main_iface.f90")).read())
    for which, sub, cname in ((0, "SUEWS_cal_Main", "suews_cal_main_"), (1, "SUEWS_cal_multitsteps", "suews_cal_multitsteps_")):
        m = re.search(r"SUBROUTINE %s\(([^)]*)\)\s*BIND\(C, name='%s'\)" % (sub, cname), f90)
        names = [a.strip() for a in m.group(1).split(",")]
        assert names == [r[0] for r in arg_table(which)]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys"""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-budget", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "cell_timesteps_per_s" and line["unit"] == "cell-timesteps/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_dropin_supy_driver_package_exports_the_reference_names():
    """suews_b200/dropin/supy_driver re-exports what supy-driver/supy_driver/__init__.py:13-27 exports (no compute here)"""
    import importlib
    here = os.path.join(ROOT, "suews_b200", "dropin")
    sys.path.insert(0, here)
    try:
        sd = importlib.import_module("supy_driver")
        names = ["suews_driver", "narp_module", "atmmoiststab_module", "resist_module", "evap_module", "snow_module",
                 "dailystate_module", "lumps_module", "anemsn_module", "rsl_module", "co2_module", "waterdist_module", "meteo"]
        for nm in names:
            assert hasattr(sd, nm), nm
        assert sd.__version__.startswith("2020b")
        for f in ("suews_cal_main", "suews_cal_multitsteps", "output_name_n", "output_size", "suews_cal_sunposition", "set_nan",
                  "square", "square_real", "cal_tair_av", "cal_tsfc"):
            assert callable(getattr(sd.suews_driver, f)), f
        with pytest.raises(NotImplementedError):
            sd.meteo.sat_vap_press_x
        # positional arguments map onto the reference's declaration order without the INTENT(OUT) ones
        from suews_b200.supy_driver import arg_table
        kw = sd.suews_driver._positional(1, (1, 2, 3), {})
        first3 = [nm.lower() for nm, _i, intent, _n, _r in arg_table(1) if intent != 2][:3]
        assert list(kw) == first3
    finally:
        sys.path.pop(0)
        sys.modules.pop("supy_driver", None)


def test_kdown_zen_needs_a_sun_and_keeps_the_daily_sums():
    """KdownZen = 1 (suews_ctrl_input.f95:1568-1600): without a sun-position function the numpy path raises instead of
    silently returning the KdownZen = 0 forcing; with one (the oracle's here) K-down is zero below the horizon and every
    day keeps its sum."""
    import fixtures as fx
    import oracle_lib as ol
    from suews_b200.forcing import disaggregate_met, read_met
    met = read_met(os.path.join(fx.GOLDEN, "forcing_2005_60min.txt"))
    with pytest.raises(NotImplementedError):
        disaggregate_met(met, kdown_zen=1)
    lat, lng, tz, alt = 51.5, -0.1, 0.0, 20.0
    zen = lambda year, dectime: ol.sun_position(year, dectime, tz, lat, lng, alt)[1]   # noqa: E731
    plain = disaggregate_met(met)
    adj = disaggregate_met(met, kdown_zen=1, zenith_fn=zen)
    cols = [c for c in range(24) if c != 14]
    assert np.array_equal(plain[:, cols], adj[:, cols])
    night = np.array([zen(float(r[1]), float(r[1] + r[2] / 24.0 + r[3] / 1440.0 - 150.0 / 86400.0)) > 90 for r in plain[:576]])
    assert night.any() and np.all(adj[:576, 14][night] == 0.0)
    for d in range(plain.shape[0] // 288):
        a, b = plain[d * 288:(d + 1) * 288, 14].sum(), adj[d * 288:(d + 1) * 288, 14].sum()
        assert abs(a - b) <= 1e-9 * max(a, 1.0)


def test_random_rain_disaggregation_keeps_amounts_and_counts():
    """RainDisaggMethod 101 / 102 (DisaggP_amongN with RainAmongN, DisaggP_amongNMult; RandomSamples,
    suews_ctrl_input.f95:2016-2177): every wet hour's rain lands in exactly amongN distinct sub-intervals in equal shares
    (the bin's amongN for method 102), dry hours stay dry, everything else is the deterministic disaggregation.  The
    reference's generator is unseeded, so there is no sequence to compare with; the same seed repeats here."""
    import fixtures as fx
    from suews_b200.forcing import disaggregate_met, read_met
    met = read_met(os.path.join(fx.GOLDEN, "forcing_2004_60min.txt"))[:2000]
    plain = disaggregate_met(met)
    with pytest.raises(ValueError):
        disaggregate_met(met, rain_disagg_method=101, rain_among_n=4)            # no generator given
    a = disaggregate_met(met, rain_disagg_method=101, rain_among_n=4, rng=7)
    b = disaggregate_met(met, rain_disagg_method=101, rain_among_n=4, rng=7)
    c = disaggregate_met(met, rain_disagg_method=101, rain_among_n=4, rng=8)
    assert np.array_equal(a, b) and not np.array_equal(a[:, 13], c[:, 13])
    cols = [k for k in range(24) if k not in (13, 15)]
    assert np.array_equal(a[:, cols], plain[:, cols])
    wet = met[:, 13] > 0
    assert wet.sum() > 50
    fast = a[:, 13].reshape(-1, 12)
    assert np.all(fast[~wet] == 0)
    assert np.all((fast[wet] > 0).sum(axis=1) == 4)
    assert np.allclose(fast[wet].sum(axis=1), met[wet, 13], rtol=1e-15)
    assert np.all(np.abs(fast[wet].max(axis=1) - met[wet, 13] / 4) == 0)
    upper, among = [0.2, 0.5, 1.0, 2.0, 5.0], [2, 3, 6, 12, 12]
    m = disaggregate_met(met, rain_disagg_method=102, mult_rain_among_n=among, mult_rain_among_n_upper_i=upper, rng=3).reshape(-1, 12, 24)[:, :, 13]
    for i in np.nonzero(wet)[0]:
        k = next((j for j in range(5) if met[i, 13] <= upper[j]), 4)
        assert (m[i] > 0).sum() == among[k] and abs(m[i].sum() - met[i, 13]) <= 1e-15 * met[i, 13]
    with pytest.raises(ValueError):
        disaggregate_met(met, rain_disagg_method=101, rain_among_n=13, rng=1)    # RainAmongN > Nper: ErrorHint 2


def test_wet_bulb_search_root_bound_on_the_host(tmp_path):
    """The three-level variant of the wet-bulb search (suews_b200/csrc/dev_latvap.cuh with SU_LATVAP_MODE 2, compiled for the
    host) against the plain loop on 3 million random arguments: identical results, and most trips are decided by the root
    bound.  (The default build stops at two levels: the variant measured 1.5 % slower on the B200, see the header.)"""
    import ctypes as C
    import subprocess
    so = str(tmp_path / "host_latvap.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so,
                           os.path.join(ROOT, "tests", "host_latvap.cpp")])
    L = C.CDLL(so)
    L.latvap_host_selftest.restype = C.c_longlong
    L.latvap_host_selftest.argtypes = [C.c_longlong, C.c_ulonglong, C.POINTER(C.c_longlong)]
    trips = (C.c_longlong * 3)()
    assert L.latvap_host_selftest(3_000_000, 2026, trips) == 0
    root, screen, exact = trips[0], trips[1], trips[2]
    assert root > 3 * (screen + exact), (root, screen, exact)


def test_device_disaggregation_source_on_the_host(tmp_path):
    """suews_b200/csrc/forcing_kernels.cuh compiled for the host (tests/host_forcing.cpp): the kernel's rows are bit for bit
    those of forcing.py -- whose output reproduces the reference's own disaggregated files -- for the met forcing of 2004 and
    2005 with DisaggMethod 1 / 2 / 3, raw file units, and the ESTM surface temperatures (methods 1 / 2) against the 5-min file
    the reference wrote."""
    import ctypes as C
    from suews_b200 import forcing
    so = str(tmp_path / "host_forcing.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so,
                           os.path.join(ROOT, "tests", "host_forcing.cpp")])
    H = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    H.host_disagg.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp]

    def dev(slow, method, raw=0, smd=0, res=3600, tstep=300):
        s = np.asfortranarray(slow, dtype=np.float64)
        out = np.zeros((s.shape[0] * (res // tstep), s.shape[1]), order="F")
        H.host_disagg(s.ctypes.data_as(dp), s.shape[0], s.shape[1], res, tstep, method, raw, smd, out.ctypes.data_as(dp))
        return out
    G = os.path.join(ROOT, "tests", "golden")
    for year in (2004, 2005):
        hourly = forcing.read_met(os.path.join(G, "forcing_%d_60min.txt" % year))
        for method in (1, 2, 3):
            ref = forcing.disaggregate_met(hourly, tstep=300, resolution_in=3600, disagg_method=method)
            assert np.array_equal(dev(hourly, method), ref), (year, method)
        for tstep in (600, 1800):
            assert np.array_equal(dev(hourly, 1, tstep=tstep), forcing.disaggregate_met(hourly, tstep=tstep, resolution_in=3600))
    raw = np.loadtxt(os.path.join(G, "forcing_2005_60min.txt"), skiprows=1)[:, :24]
    assert np.array_equal(dev(raw, 1, raw=1), forcing.disaggregate_met(forcing.read_met(os.path.join(G, "forcing_2005_60min.txt")), disagg_method=1))
    ts = np.loadtxt(os.path.join(G, "estm_ts_2005_60min.txt"), skiprows=1)
    gold = np.loadtxt(os.path.join(G, "ref_estm_ts_2005_5min.txt"), skiprows=1)
    for method in (1, 2):
        out = dev(ts, method)
        assert np.array_equal(out, forcing.disaggregate_estm(ts, disagg_method=method)), method
    out = dev(ts, 1)
    n = min(out.shape[0], gold.shape[0])
    assert np.array_equal(out[:n, :4], gold[:n, :4]) and np.all(np.abs(out[:n, 4:] - gold[:n, 4:]) <= 5.1e-5)
