#!/usr/bin/env python3
"""Build tests/golden/ from the reference's own BaseRun fixtures.

Run in the authoring container only (needs /root/reference); outputs are small
and committed.  Nothing here is reference *code*: only its test data.

  state_init_{2004,2005}.json   the complete SUEWS_cal_Main argument dump at the first
                                step (Test/BaseRun/2020b/Output/test1_*_state_init.txt,
                                written by suews_ctrl_translate.f95:1280-1500)
  forcing_2005_5min.txt         Test/BaseRun/2019a/Input/test_2005_data_5.txt (1440 rows)
  forcing_2005_60min.txt        Test/BaseRun/2020b/Input/test_2005_data_60.txt (119 rows)
  forcing_2004_60min.txt        Test/BaseRun/2020b/Input/test_2004_data_60.txt.long (8784 rows)
  ref_2005_{SUEWS,snow,RSL,SOLWEIG,BL}_60.txt, ref_2005_DailyState.txt, ref_2004_DailyState.txt, ref_2004_{SOLWEIG,BL}_60.txt
                                the reference's golden outputs (hourly, 4-6 decimals)
  estm_ts_2005_60min.txt, ref_estm_ts_2005_5min.txt
                                ESTM surface-temperature forcing: the hourly input and the 5-min series the reference wrote
                                from it (DisaggregateESTM with KeepTstepFilesIn = 1, f9.4; Release/InputTables/2020a/)
  ref_InitialConditions_2004_input.nml (the user-written file the golden run starts from), ref_InitialConditions_2005.nml,
  ref_InitialConditions_2005_EndofRun.nml
                                the state the reference wrote at the end of its 2004 leg (105 408 steps) and of its 2005
                                leg (1 440 steps): NextInitial, suews_ctrl_init.f95:2022-2149, list-directed, 17
                                significant digits -- i.e. the exact doubles (Test/BaseRun/2020b/Input/)
  baserun_2020b/RunControl.nml, baserun_2020b/Input/SUEWS_*.txt, ESTMinput.nml, InitialConditionstest_2004.nml
                                the run directory of the same golden run: the code-keyed input tables the state_init
                                dumps were translated from
"""
import json, os, re, shutil

REF = "/root/reference/Test/BaseRun"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def parse_state_init(path):
    d = {}
    for ln in open(path):
        ln = ln.strip()
        if not ln or ln.startswith("&") or ln == "/":
            continue
        m = re.match(r"^(\w+)\s*=\s*(.*)$", ln)
        if not m:
            continue
        key = m.group(1).lower()
        vals = []
        for tok in m.group(2).replace(",", " ").split():
            mm = re.match(r"^(\d+)\*(.*)$", tok)  # Fortran repeat count r*c
            if mm:
                vals += [mm.group(2)] * int(mm.group(1))
            else:
                vals.append(tok)
        out = []
        for v in vals:
            if re.fullmatch(r"[-+]?\d+", v):
                out.append(int(v))
            else:
                out.append(float(v.replace("D", "E").replace("d", "e")))
        d[key] = out
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    for yr in (2004, 2005):
        d = parse_state_init(f"{REF}/2020b/Output/test1_{yr}_state_init.txt")
        with open(os.path.join(OUT, f"state_init_{yr}.json"), "w") as fh:
            json.dump(d, fh, indent=0, sort_keys=True)
        print(yr, len(d), "entries")
    cp = [
        ("2019a/Input/test_2005_data_5.txt", "forcing_2005_5min.txt"),
        ("2020b/Input/test_2005_data_60.txt", "forcing_2005_60min.txt"),
        ("2020b/Input/test_2004_data_60.txt.long", "forcing_2004_60min.txt"),
        ("2020b/Output/test1_2005_SUEWS_60.txt", "ref_2005_SUEWS_60.txt"),
        ("2020b/Output/test1_2005_snow_60.txt", "ref_2005_snow_60.txt"),
        ("2020b/Output/test1_2005_RSL_60.txt", "ref_2005_RSL_60.txt"),
        ("2020b/Output/test1_2005_DailyState.txt", "ref_2005_DailyState.txt"),
        ("2020b/Output/test1_2004_DailyState.txt", "ref_2004_DailyState.txt"),
        ("2020b/Output/test1_2005_SOLWEIG_60.txt", "ref_2005_SOLWEIG_60.txt"),
        ("2020b/Output/test1_2004_SOLWEIG_60.txt", "ref_2004_SOLWEIG_60.txt"),
        ("2020b/Output/test1_2005_BL_60.txt", "ref_2005_BL_60.txt"),
        ("2020b/Output/test1_2004_BL_60.txt", "ref_2004_BL_60.txt"),
        ("../../Release/InputTables/2020a/test1_2005_ESTM_Ts_data_60.txt", "estm_ts_2005_60min.txt"),
        ("../../Release/InputTables/2020a/test1_2005_ESTM_Ts_data_5.txt", "ref_estm_ts_2005_5min.txt"),
        ("2020b/Input/InitialConditionstest_2004.nml", "ref_InitialConditions_2004_input.nml"),
        ("2020b/Input/InitialConditionstest1_2005.nml", "ref_InitialConditions_2005.nml"),
        ("2020b/Input/InitialConditionstest1_2005_EndofRun.nml", "ref_InitialConditions_2005_EndofRun.nml"),
    ]
    for src, dst in cp:
        shutil.copyfile(f"{REF}/{src}", os.path.join(OUT, dst))
        os.chmod(os.path.join(OUT, dst), 0o644)
    # the input tables and RunControl.nml of the same run: what suews_b200/tables.py joins (tests/test_tables.py)
    run = os.path.join(OUT, "baserun_2020b")
    os.makedirs(os.path.join(run, "Input"), exist_ok=True)
    shutil.copyfile(f"{REF}/2020b/RunControl.nml", os.path.join(run, "RunControl.nml"))
    for fn in sorted(os.listdir(f"{REF}/2020b/Input")):
        if fn.startswith("SUEWS_") and fn.endswith(".txt") or fn in ("ESTMinput.nml", "InitialConditionstest_2004.nml", "CBLInput.nml"):
            shutil.copyfile(f"{REF}/2020b/Input/{fn}", os.path.join(run, "Input", fn))
            os.chmod(os.path.join(run, "Input", fn), 0o644)
    os.chmod(os.path.join(run, "RunControl.nml"), 0o644)
    os.makedirs(os.path.join(run, "Input", "CBLinputfiles"), exist_ok=True)   # BLUEWS: the slab's initial values per day of year
    shutil.copyfile(f"{REF}/2020b/Input/CBLinputfiles/CBL_initial_data.txt", os.path.join(run, "Input", "CBLinputfiles", "CBL_initial_data.txt"))
    os.chmod(os.path.join(run, "Input", "CBLinputfiles", "CBL_initial_data.txt"), 0o644)
    print("golden files in", OUT)


if __name__ == "__main__":
    main()
