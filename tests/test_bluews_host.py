"""The device-side BLUEWS stage (suews_b200/csrc/dev_bluews.cuh) compiled for the host (tests/host_bluews.cpp) and replayed on
the inputs the oracle's own run handed to its BLUEWS -- every one of them is on the BL output line; z/L of the step comes from
the zdm and Lob columns -- against the oracle's lines and slab state, bit for bit: what the kernel compiles is the algorithm
the golden BL files pin, checked without a GPU."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fixtures as fx  # noqa: E402
import oracle_lib as ol  # noqa: E402

dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
PSI = C.CFUNCTYPE(C.c_double, C.c_int, C.c_double)
NAMES = ("cbl_blh_m", "cbl_tm_K", "cbl_qm_kgkg", "cbl_tp_K", "cbl_qp_kgkg", "cbl_cm", "cbl_gamt_Km", "cbl_gamq_kgkgm")


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_bluews") / "host_bluews.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", out,
                           os.path.join(ol.ROOT, "tests", "host_bluews.cpp")])
    H = C.CDLL(out)
    H.bluews_host_run.argtypes = [C.c_int, ip, dp, ip, C.c_double, dp, C.c_int, PSI, dp, dp, C.POINTER(C.c_ulonglong)]
    return H


# (InitialData_use, EntrainmentType, isubs, rows of the initial-data table kept)
@pytest.mark.parametrize("mode,et,isubs,rows", [(2, 4, 0, None), (2, 1, 1, None), (2, 2, 0, None), (1, 3, 0, None), (1, 4, 1, slice(0, 151))])
def test_device_bluews_source_equals_the_oracle(host, mode, et, isubs, rows):
    """Three summer days of 2004 (days 151-153; with rows 1-151 only, InitialData_use = 1 diagnoses day 151 and leaves the
    night's 200 m values on the other two): the 17 values of every step's line and the eight slab variables at the end."""
    L = ol.lib()
    L.so_set_cbl_initial.argtypes = [dp, C.c_int]
    L.so_set_bl_block.argtypes = [dp]
    L.so_stab_psi_heat.restype = C.c_double
    L.so_stab_psi_heat.argtypes = [C.c_int, C.c_double]
    psih = PSI(lambda sm, zl: L.so_stab_psi_heat(sm, zl))
    params, ini = fx.cbl_inputs()
    ini = np.ascontiguousarray(ini if rows is None else ini[rows])
    d = fx.state_init(2004)
    d.update(params)
    d.update(initialdata_use=[mode], entrainmenttype=[et], isubs=[isubs])
    nsteps = 288 * 3
    met = fx.forcing(2004)[288 * 150:288 * 150 + nsteps]
    bl = np.full((nsteps, 22), -999.0, order="F")
    cell = ol.Cell(d)
    start = np.array([float(cell.get_state(k)[0]) for k in NAMES])
    L.so_set_cbl_initial(ini.ctypes.data_as(dp), ini.shape[0])
    L.so_set_bl_block(bl.ctypes.data_as(dp))
    try:
        su = cell.run(met, want=("SUEWS",))["SUEWS"]
    finally:
        L.so_set_bl_block(None)
        L.so_set_cbl_initial(None, 0)
    end = np.array([float(cell.get_state(k)[0]) for k in NAMES])
    clock = np.ascontiguousarray(bl[:, :4].astype(np.int32))
    zl = (d["z"][0] - su[:, 5 + 53]) / su[:, 5 + 55]                      # (z - zdm) / Lob: ZZD / L_mod of the step's last trip
    # avkdn Press_hPa avU1 qh_use qe_use UStar avdens lv_J_kg avcp z/L
    drive = np.ascontiguousarray(np.column_stack([met[:, 14], bl[:, 14], bl[:, 15], bl[:, 12], bl[:, 13], bl[:, 16], bl[:, 17],
                                                  bl[:, 18], bl[:, 19], zl]))
    cfg = np.array([et, int(d["qh_choice"][0]), isubs, int(d["co2_included"][0]), mode, 0, int(d["tstep"][0]),
                    int(d["stabilitymethod"][0]), 2], dtype=np.int32)
    for with_lines in (True, False):                                      # policy FULL, and P0 (no line: the night's Temp_C / rh only when needed)
        st = start.copy()
        lines = np.zeros((nsteps, 17))
        warn = (C.c_ulonglong * 2)()
        rc = host.bluews_host_run(nsteps, clock.ctypes.data_as(ip), drive.ctypes.data_as(dp), cfg.ctypes.data_as(ip), float(d["wsb"][0]),
                                  ini.ctypes.data_as(dp), ini.shape[0], psih, st.ctypes.data_as(dp),
                                  lines.ctypes.data_as(dp) if with_lines else None, warn)
        assert rc == 0
        assert np.array_equal(st, end), (with_lines, st - end)
        if with_lines:
            assert np.array_equal(lines, bl[:, 5:]), np.abs(lines - bl[:, 5:]).max(axis=0)
    assert bl[:, 5].max() > 400.0 and bl[:, 5].min() <= 200.0              # the slab grew by day and was reset at night (subsidence can shrink it further)
    if mode == 1:
        day151_start = bl[np.nonzero(met[:, 14] >= 5)[0][0], 6]           # theta of the first daytime line
        assert abs(day151_start - ini[min(150, ini.shape[0] - 1), 6]) > 1e-6   # diagnosed, not the file's theta


def test_device_bluews_refuses_what_the_reference_cannot_run(host):
    psih = PSI(lambda sm, zl: 0.0)
    _params, ini = fx.cbl_inputs()
    ini = np.ascontiguousarray(ini)
    clock = np.array([[2004, 1, 0, 5]], dtype=np.int32)
    drive = np.array([[0.0, 1000.0, 3.0, 10.0, 5.0, 0.3, 1.2, 2.5e6, 1005.0, 0.1]])
    for cfg9, nline in (([4, 1, 0, 0, 0, 0, 300, 3, 2], ini.shape[0]),      # InitialData_use 0
                        ([4, 1, 0, 0, 2, 1, 300, 3, 2], ini.shape[0]),      # sounding files
                        ([4, 1, 0, 0, 1, 0, 300, 3, 1], ini.shape[0]),      # psih does not exist with AerodynamicResistanceMethod 1
                        ([4, 1, 0, 0, 2, 0, 300, 3, 2], 0)):                # no table
        cfg = np.array(cfg9, dtype=np.int32)
        st = np.zeros(8)
        rc = host.bluews_host_run(1, clock.ctypes.data_as(ip), drive.ctypes.data_as(dp), cfg.ctypes.data_as(ip), 0.0,
                                  ini.ctypes.data_as(dp), nline, psih, st.ctypes.data_as(dp), None, None)
        assert rc == 1007, cfg9
