"""The reference's input tables -> the arguments of SUEWS_cal_Main (the table join).

A SUEWS run directory describes every grid by one row of SUEWS_SiteSelect.txt whose code columns point into the
code-keyed tables (SUEWS_NonVeg / Veg / Water / Snow / Soil / Conductance / OHMCoefficients / AnthropogenicEmission /
Irrigation / Profiles / WithinGridWaterDist / BiogenCO2 / ESTMCoefficients .txt).  The reference joins them in
InitializeSurfaceCharacteristics (suews_ctrl_init.f95:454-1432, the CodeMatch* lookups of suews_ctrl_input.f95:832-1321)
into one long SurfaceChar row per grid, and SUEWS_Translate (suews_ctrl_translate.f95:38-1260) unpacks that row into the
variables SUEWS_cal_Main receives.  This module does both steps for the library's hosts: `load_tables` reads a run
directory, `site_arguments` returns {argument name: values} for one (grid, year) -- ready for
SuewsBatch.load_reference_args, and the `params` of checkpoint.initial_state.

Pinned on the reference's own dump: for Test/BaseRun/2020b it lands on every table-derived argument of the
`*_state_init.txt` dumps the reference wrote for 2004 and 2005 (tests/test_tables.py; fixtures are byte copies of that
run's Input/ directory).

Kept as the reference has them: the table columns are taken by position (InputHeaderCheck makes the header names and
positions agree, suews_ctrl_input.f95:832; here a wrong header is an error as well), beta_enh_bioCO2 is filled from the
alpha_enh column (suews_ctrl_init.f95: `c_beta_enh_bioCO2 = Biogen_Coeff(iv5, cB_alpha_enh)`), the water surface has no
soil store, snow-patch limit or deep-soil row (zeros), the ToRunoff / ToSoilStore columns of the within-grid water
distribution share row 8 of WaterDist (whichever is non-zero), PopDensDaytime(2) is derived with the weekend fraction
FrPDDwe, and a missing day / night population density is replaced by the other one.
"""
import os
import re

import numpy as np

SURF = ("Paved", "Bldgs", "EveTr", "DecTr", "Grass", "BSoil", "Water")
NAN = -999.0

HEADERS = {
    "SUEWS_NonVeg.txt": "Code AlbedoMin AlbedoMax Emissivity StorageMin StorageMax WetThreshold StateLimit DrainageEq "
                        "DrainageCoef1 DrainageCoef2 SoilTypeCode SnowLimPatch SnowLimRemove OHMCode_SummerWet "
                        "OHMCode_SummerDry OHMCode_WinterWet OHMCode_WinterDry OHMThresh_SW OHMThresh_WD ESTMCode AnOHM_Cp "
                        "AnOHM_Kk AnOHM_Ch",
    "SUEWS_Veg.txt": "Code AlbedoMin AlbedoMax Emissivity StorageMin StorageMax WetThreshold StateLimit DrainageEq "
                     "DrainageCoef1 DrainageCoef2 SoilTypeCode SnowLimPatch BaseT BaseTe GDDFull SDDFull LAIMin LAIMax "
                     "PorosityMin PorosityMax MaxConductance LAIEq LeafGrowthPower1 LeafGrowthPower2 LeafOffPower1 "
                     "LeafOffPower2 OHMCode_SummerWet OHMCode_SummerDry OHMCode_WinterWet OHMCode_WinterDry OHMThresh_SW "
                     "OHMThresh_WD ESTMCode AnOHM_Cp AnOHM_Kk AnOHM_Ch BiogenCO2Code",
    "SUEWS_Water.txt": "Code AlbedoMin AlbedoMax Emissivity StorageMin StorageMax WetThreshold StateLimit WaterDepth "
                       "DrainageEq DrainageCoef1 DrainageCoef2 OHMCode_SummerWet OHMCode_SummerDry OHMCode_WinterWet "
                       "OHMCode_WinterDry OHMThresh_SW OHMThresh_WD ESTMCode AnOHM_Cp AnOHM_Kk AnOHM_Ch",
    "SUEWS_Snow.txt": "Code RadMeltFactor TempMeltFactor AlbedoMin AlbedoMax Emissivity tau_a tau_f PrecipLimAlb SnowDensMin "
                      "SnowDensMax tau_r CRWMin CRWMax PrecipLimSnow OHMCode_SummerWet OHMCode_SummerDry OHMCode_WinterWet "
                      "OHMCode_WinterDry OHMThresh_SW OHMThresh_WD ESTMCode AnOHM_Cp AnOHM_Kk AnOHM_Ch",
    "SUEWS_Soil.txt": "Code SoilDepth SoilStoreCap SatHydraulicCond SoilDensity InfiltrationRate OBS_SMDepth OBS_SMCap "
                      "OBS_SoilNotRocks",
    "SUEWS_Conductance.txt": "Code G1 G2 G3 G4 G5 G6 TH TL S1 S2 Kmax gsModel",
    "SUEWS_OHMCoefficients.txt": "Code a1 a2 a3",
    "SUEWS_AnthropogenicEmission.txt": "Code BaseT_HC QF_A_WD QF_B_WD QF_C_WD QF_A_WE QF_B_WE QF_C_WE AHMin_WD AHMin_WE "
                                       "AHSlope_Heating_WD AHSlope_Heating_WE AHSlope_Cooling_WD AHSlope_Cooling_WE "
                                       "TCritic_Heating_WD TCritic_Heating_WE TCritic_Cooling_WD TCritic_Cooling_WE "
                                       "EnergyUseProfWD EnergyUseProfWE ActivityProfWD ActivityProfWE TraffProfWD TraffProfWE "
                                       "PopProfWD PopProfWE MinQFMetab MaxQFMetab MinFCMetab MaxFCMetab FrPDDwe "
                                       "FrFossilFuel_Heat FrFossilFuel_NonHeat EF_umolCO2perJ EnEF_v_Jkm FcEF_v_kgkmWD "
                                       "FcEF_v_kgkmWE CO2PointSource TrafficUnits",
    "SUEWS_Irrigation.txt": "Code Ie_start Ie_end InternalWaterUse Faut H_maintain Ie_a1 Ie_a2 Ie_a3 Ie_m1 Ie_m2 Ie_m3 "
                            "DayWat(1) DayWat(2) DayWat(3) DayWat(4) DayWat(5) DayWat(6) DayWat(7) DayWatPer(1) DayWatPer(2) "
                            "DayWatPer(3) DayWatPer(4) DayWatPer(5) DayWatPer(6) DayWatPer(7)",
    "SUEWS_Profiles.txt": "Code " + " ".join(str(h) for h in range(24)),
    "SUEWS_WithinGridWaterDist.txt": "Code ToPaved ToBldgs ToEveTr ToDecTr ToGrass ToBSoil ToWater ToRunoff ToSoilStore",
    "SUEWS_BiogenCO2.txt": "Code alpha beta theta alpha_enh beta_enh resp_a resp_b min_respi",
    "SUEWS_SiteSelect.txt": "Grid Year StartDLS EndDLS lat lng Timezone SurfaceArea Alt z id ih imin Fr_Paved Fr_Bldgs Fr_EveTr "
                            "Fr_DecTr Fr_Grass Fr_Bsoil Fr_Water IrrFr_Paved IrrFr_Bldgs IrrFr_EveTr IrrFr_DecTr IrrFr_Grass "
                            "IrrFr_BSoil IrrFr_Water H_Bldgs H_EveTr H_DecTr z0 zd FAI_Bldgs FAI_EveTr FAI_DecTr PopDensDay "
                            "PopDensNight TrafficRate_WD TrafficRate_WE QF0_BEU_WD QF0_BEU_WE Code_Paved Code_Bldgs Code_EveTr "
                            "Code_DecTr Code_Grass Code_Bsoil Code_Water LUMPS_DrRate LUMPS_Cover LUMPS_MaxRes NARP_Trans "
                            "CondCode SnowCode SnowClearingProfWD SnowClearingProfWE AnthropogenicCode IrrigationCode "
                            "WaterUseProfManuWD WaterUseProfManuWE WaterUseProfAutoWD WaterUseProfAutoWE FlowChange "
                            "RunoffToWater PipeCapacity GridConnection1of8 Fraction1of8 GridConnection2of8 Fraction2of8 "
                            "GridConnection3of8 Fraction3of8 GridConnection4of8 Fraction4of8 GridConnection5of8 Fraction5of8 "
                            "GridConnection6of8 Fraction6of8 GridConnection7of8 Fraction7of8 GridConnection8of8 Fraction8of8 "
                            "WithinGridPavedCode WithinGridBldgsCode WithinGridEveTrCode WithinGridDecTrCode "
                            "WithinGridGrassCode WithinGridUnmanBSoilCode WithinGridWaterCode AreaWall Fr_ESTMClass_Paved1 "
                            "Fr_ESTMClass_Paved2 Fr_ESTMClass_Paved3 Code_ESTMClass_Paved1 Code_ESTMClass_Paved2 "
                            "Code_ESTMClass_Paved3 Fr_ESTMClass_Bldgs1 Fr_ESTMClass_Bldgs2 Fr_ESTMClass_Bldgs3 "
                            "Fr_ESTMClass_Bldgs4 Fr_ESTMClass_Bldgs5 Code_ESTMClass_Bldgs1 Code_ESTMClass_Bldgs2 "
                            "Code_ESTMClass_Bldgs3 Code_ESTMClass_Bldgs4 Code_ESTMClass_Bldgs5",
}


class TableError(ValueError):
    """ErrorHint 56 / 57 / 58 of the reference: header mismatch, code not found, code not unique."""


class Table:
    """One SUEWS_*.txt table: row 1 column numbers, row 2 names, data rows, a line starting with -9 ends the table
    (ReadCoeff, suews_ctrl_input.f95:1130-1321).  `!` starts a comment."""

    def __init__(self, path, expect=None):
        self.name = os.path.basename(path)
        lines = open(path, encoding="latin-1").read().splitlines()   # (comments hold Latin-1 names)
        names = lines[1].split("!")[0].split()
        if expect is not None:
            want = expect.split()
            if [n.lower() for n in names[:len(want)]] != [w.lower() for w in want]:
                bad = next((a, b) for a, b in zip(names + [None] * len(want), want) if a is None or a.lower() != b.lower())
                raise TableError("ErrorHint 56: names or order of the columns of %s do not match the model code (found %r, expected %r)"
                                 % (self.name, bad[0], bad[1]))
            names = want
        self.names = names
        self.col = {n.lower(): i for i, n in enumerate(names)}
        rows = []
        for ln in lines[2:]:
            tok = ln.split("!")[0].split()
            if not tok:
                continue
            if tok[0] in ("-9", "-9.0"):
                break
            rows.append([float(t) for t in tok[:len(names)]])
        self.rows = np.array(rows, dtype=np.float64).reshape(len(rows), len(names))

    def lookup(self, code, what):
        """CodeMatch*: the row whose Code equals `code` (exactly one, else ErrorHint 57)."""
        hit = np.nonzero(self.rows[:, 0] == code)[0]
        if hit.size != 1:
            raise TableError("ErrorHint 57: %s code %g %s in %s" % (what, code, "not found" if hit.size == 0 else "is not unique", self.name))
        return self.rows[hit[0]]

    def get(self, row, name):
        return float(row[self.col[name.lower()]])


def read_runcontrol(path):
    """RunControl.nml -> {lower-case name: int / float / str}"""
    out = {}
    for ln in open(path).read().splitlines():
        ln = ln.split("!")[0].strip()
        m = re.match(r"(\w+)\s*=\s*(.+)$", ln)
        if not m:
            continue
        v = m.group(2).strip().rstrip(",")
        if v[:1] in "'\"":
            out[m.group(1).lower()] = v.strip("'\"")
        else:
            try:
                out[m.group(1).lower()] = int(v)
            except ValueError:
                out[m.group(1).lower()] = float(v.lower().replace("d", "e"))
    return out


def load_tables(input_dir, estm=False):
    """Every table of a run directory's Input/ (SUEWS_ESTMCoefficients.txt only when asked: 52 columns, read by name)."""
    t = {}
    for fn, hdr in HEADERS.items():
        t[fn[6:-4]] = Table(os.path.join(input_dir, fn), hdr)
    if estm:
        t["ESTMCoefficients"] = Table(os.path.join(input_dir, "SUEWS_ESTMCoefficients.txt"))
    return t


# defaults of the switches that are not in RunControl.nml (suews_ctrl_init.f95:111-112, :297-298; suews_ctrl_const.f95:1377,
# :1407) and of the namelist entries the reference gives a default (suews_ctrl_const.f95)
_RUN_DEFAULTS = dict(diagqn=0, diagqs=0, laicalcyes=1, evapmethod=2, veg_type=1, aerodynamicresistancemethod=2, diagnose=0,
                     basetmethod=1, snowuse=0, ohmincqf=0, smdmethod=0, waterusemethod=0, emissionsmethod=2,
                     netradiationmethod=3, storageheatmethod=1, stabilitymethod=3, roughlenheatmethod=2, roughlenmommethod=2,
                     tstep=300)


def site_arguments(tables, runcontrol, grid, year):
    """SUEWS_cal_Main's table-derived arguments for one (grid, year) row of SUEWS_SiteSelect.txt, lower-case names.
    Also returned under '_aux': values the reference keeps beside the argument list (WaterDepth for InitialState, the
    ESTM codes and class fractions, the profile codes)."""
    rc = dict(_RUN_DEFAULTS)
    rc.update({k.lower(): v for k, v in runcontrol.items()})
    ss = tables["SiteSelect"]
    hit = np.nonzero((ss.rows[:, 0] == grid) & (ss.rows[:, 1] == year))[0]
    if hit.size != 1:
        raise TableError("SUEWS_SiteSelect.txt: grid %s, year %s %s" % (grid, year, "not found" if hit.size == 0 else "is not unique"))
    s = ss.rows[hit[0]]
    S = lambda name: ss.get(s, name)   # noqa: E731

    nonveg, veg, water, snow, soil = tables["NonVeg"], tables["Veg"], tables["Water"], tables["Snow"], tables["Soil"]
    ohm, dist, prof, biogen = tables["OHMCoefficients"], tables["WithinGridWaterDist"], tables["Profiles"], tables["BiogenCO2"]
    rows, tabs = [], []
    for nm, tab in zip(("Paved", "Bldgs", "EveTr", "DecTr", "Grass", "Bsoil", "Water"),
                       (nonveg, nonveg, veg, veg, veg, nonveg, water)):
        rows.append(tab.lookup(S("Code_" + nm), "surface " + nm))
        tabs.append(tab)
    snow_row = snow.lookup(S("SnowCode"), "snow")

    def per_surf(name, n=7, missing=0.0):
        return [tabs[i].get(rows[i], name) if name.lower() in tabs[i].col else missing for i in range(n)]

    a = {}
    a["lat"], a["lng"], a["timezone"], a["alt"], a["z"] = [S("lat")], [S("lng")], [S("Timezone")], [S("Alt")], [S("z")]
    a["surfacearea"] = [S("SurfaceArea") * 10000]
    a["sfr"] = [S("Fr_" + n) for n in ("Paved", "Bldgs", "EveTr", "DecTr", "Grass", "Bsoil", "Water")]
    if not 0.999 <= sum(a["sfr"]) <= 1.001:
        raise TableError("ErrorHint 10: surface fractions (Fr_) should add up to 1 (%g)" % sum(a["sfr"]))
    for n in SURF:
        a["irrfrac" + n.lower()] = [S("IrrFr_" + n)]
    a["bldgh"], a["evetreeh"], a["dectreeh"] = [S("H_Bldgs")], [S("H_EveTr")], [S("H_DecTr")]
    a["faibldg"], a["faievetree"], a["faidectree"] = [S("FAI_Bldgs")], [S("FAI_EveTr")], [S("FAI_DecTr")]
    a["z0m_in"], a["zdm_in"] = [S("z0")], [S("zd")]
    a["startdls"], a["enddls"] = [int(S("StartDLS"))], [int(S("EndDLS"))]
    a["drainrt"], a["raincover"], a["rainmaxres"] = [S("LUMPS_DrRate")], [S("LUMPS_Cover")], [S("LUMPS_MaxRes")]
    a["narp_trans_site"] = [S("NARP_Trans")]
    a["flowchange"], a["runofftowater"], a["pipecapacity"] = [S("FlowChange")], [S("RunoffToWater")], [S("PipeCapacity")]
    a["trafficrate"] = [S("TrafficRate_WD"), S("TrafficRate_WE")]
    a["qf0_beu"] = [S("QF0_BEU_WD"), S("QF0_BEU_WE")]

    # surface tables
    a["alb"] = per_surf("AlbedoMax")
    albmin = per_surf("AlbedoMin")
    for i, n in ((2, "evetr"), (3, "dectr"), (4, "grass")):
        a["albmin_" + n], a["albmax_" + n] = [albmin[i]], [a["alb"][i]]
    a["emis"] = per_surf("Emissivity")
    sdp = np.zeros((6, 7))
    sdp[0] = per_surf("StorageMin")
    sdp[4] = per_surf("StorageMax")
    sdp[5] = sdp[0]
    sdp[1], sdp[2], sdp[3] = per_surf("DrainageEq"), per_surf("DrainageCoef1"), per_surf("DrainageCoef2")
    a["storedrainprm"] = list(sdp.T.ravel())              # (6, nsurf) column-major
    a["capmin_dec"], a["capmax_dec"] = [sdp[0, 3]], [sdp[4, 3]]
    a["pormin_dec"], a["pormax_dec"] = [veg.get(rows[3], "PorosityMin")], [veg.get(rows[3], "PorosityMax")]
    a["wetthresh"], a["statelimit"] = per_surf("WetThreshold"), per_surf("StateLimit")
    a["snowpacklimit"] = per_surf("SnowLimPatch", 6) + [0.0]
    a["snowlimpaved"], a["snowlimbldg"] = [nonveg.get(rows[0], "SnowLimRemove")], [nonveg.get(rows[1], "SnowLimRemove")]
    soil_rows = [soil.lookup(tabs[i].get(rows[i], "SoilTypeCode"), "soil (%s)" % SURF[i]) for i in range(6)]
    a["soildepth"] = [soil.get(r, "SoilDepth") for r in soil_rows] + [0.0]
    a["soilstorecap"] = [soil.get(r, "SoilStoreCap") for r in soil_rows] + [0.0]
    a["sathydraulicconduct"] = [soil.get(r, "SatHydraulicCond") for r in soil_rows] + [0.0]
    vrows = rows[2:5]
    for arg, colname in (("baset", "BaseT"), ("basete", "BaseTe"), ("gddfull", "GDDFull"), ("sddfull", "SDDFull"),
                         ("laimin", "LAIMin"), ("laimax", "LAIMax"), ("maxconductance", "MaxConductance")):
        a[arg] = [veg.get(r, colname) for r in vrows]
    a["laitype"] = [int(veg.get(r, "LAIEq")) for r in vrows]
    lp = np.array([[veg.get(r, c) for r in vrows] for c in ("LeafGrowthPower1", "LeafGrowthPower2", "LeafOffPower1", "LeafOffPower2")])
    a["laipower"] = list(lp.T.ravel())                     # (4, nvegsurf) column-major
    brows = [biogen.lookup(veg.get(r, "BiogenCO2Code"), "biogenic CO2") for r in vrows]
    for arg, colname in (("alpha_bioco2", "alpha"), ("beta_bioco2", "beta"), ("theta_bioco2", "theta"),
                         ("alpha_enh_bioco2", "alpha_enh"), ("beta_enh_bioco2", "alpha_enh"),   # (sic, see the module header)
                         ("resp_a", "resp_a"), ("resp_b", "resp_b"), ("min_res_bioco2", "min_respi")):
        a[arg] = [biogen.get(r, colname) for r in brows]
    if int(rc["storageheatmethod"]) == 3:                  # translated for AnOHM runs only (translate.f95:671-675)
        a["cpanohm"], a["kkanohm"], a["chanohm"] = per_surf("AnOHM_Cp"), per_surf("AnOHM_Kk"), per_surf("AnOHM_Ch")
    else:
        a["cpanohm"], a["kkanohm"], a["chanohm"] = [0.0] * 7, [0.0] * 7, [0.0] * 7

    # OHM coefficients (8 surfaces incl. snow, 4 season/wetness classes, a1-a3) and thresholds
    oc = np.zeros((8, 4, 3))
    for i in range(8):
        tab, row = (tabs[i], rows[i]) if i < 7 else (snow, snow_row)
        for j, cls in enumerate(("SummerWet", "SummerDry", "WinterWet", "WinterDry")):
            orow = ohm.lookup(tab.get(row, "OHMCode_" + cls), "OHM (%s, %s)" % (SURF[i] if i < 7 else "snow", cls))
            oc[i, j] = [ohm.get(orow, "a1"), ohm.get(orow, "a2"), ohm.get(orow, "a3")]
    a["ohm_coef"] = list(oc.ravel(order="F"))              # (nsurf+1, 4, 3) column-major
    a["ohm_threshsw"] = per_surf("OHMThresh_SW") + [snow.get(snow_row, "OHMThresh_SW")]
    a["ohm_threshwd"] = per_surf("OHMThresh_WD") + [snow.get(snow_row, "OHMThresh_WD")]

    # snow
    for arg, colname in (("radmeltfact", "RadMeltFactor"), ("tempmeltfact", "TempMeltFactor"), ("snowalbmin", "AlbedoMin"),
                         ("snowalbmax", "AlbedoMax"), ("tau_a", "tau_a"), ("tau_f", "tau_f"), ("preciplimitalb", "PrecipLimAlb"),
                         ("snowdensmin", "SnowDensMin"), ("snowdensmax", "SnowDensMax"), ("tau_r", "tau_r"), ("crwmin", "CRWMin"),
                         ("crwmax", "CRWMax"), ("preciplimit", "PrecipLimSnow")):
        a[arg] = [snow.get(snow_row, colname)]
    if int(rc["snowuse"]) == 1:
        a["narp_emis_snow"] = [snow.get(snow_row, "Emissivity")]

    # conductance
    cond = tables["Conductance"]
    crow = cond.lookup(S("CondCode"), "conductance")
    for n in ("G1", "G2", "G3", "G4", "G5", "G6", "TH", "TL", "S1", "S2", "Kmax"):
        a[n.lower()] = [cond.get(crow, n)]
    a["gsmodel"] = [int(cond.get(crow, "gsModel"))]

    # anthropogenic emissions
    an = tables["AnthropogenicEmission"]
    arow = an.lookup(S("AnthropogenicCode"), "anthropogenic")
    A = lambda n: an.get(arow, n)   # noqa: E731
    a["baset_hc"] = [A("BaseT_HC")]
    for arg, stem in (("qf_a", "QF_A"), ("qf_b", "QF_B"), ("qf_c", "QF_C"), ("ah_min", "AHMin"), ("ah_slope_heating", "AHSlope_Heating"),
                      ("ah_slope_cooling", "AHSlope_Cooling"), ("baset_heating", "TCritic_Heating"), ("baset_cooling", "TCritic_Cooling")):
        a[arg] = [A(stem + "_WD"), A(stem + "_WE")]
    for n in ("MinQFMetab", "MaxQFMetab", "MinFCMetab", "MaxFCMetab", "FrFossilFuel_Heat", "FrFossilFuel_NonHeat", "EF_umolCO2perJ",
              "EnEF_v_Jkm", "CO2PointSource", "TrafficUnits"):
        a[n.lower()] = [A(n)]
    a["fcef_v_kgkm"] = [A("FcEF_v_kgkmWD"), A("FcEF_v_kgkmWE")]
    pd_day, pd_night = S("PopDensDay"), S("PopDensNight")
    if pd_day >= 0 and pd_night < 0:
        pd_night = pd_day
    if pd_day < 0 and pd_night >= 0:
        pd_day = pd_night
    a["popdensdaytime"] = [pd_day, pd_night + (pd_day - pd_night) * A("FrPDDwe")]
    a["popdensnighttime"] = [pd_night]

    # irrigation
    irr = tables["Irrigation"]
    irow = irr.lookup(S("IrrigationCode"), "irrigation")
    a["ie_start"], a["ie_end"] = [int(irr.get(irow, "Ie_start"))], [int(irr.get(irow, "Ie_end"))]
    a["internalwateruse_h"], a["faut"], a["h_maintain"] = [irr.get(irow, "InternalWaterUse")], [irr.get(irow, "Faut")], [irr.get(irow, "H_maintain")]
    a["ie_a"] = [irr.get(irow, "Ie_a%d" % k) for k in (1, 2, 3)]
    a["ie_m"] = [irr.get(irow, "Ie_m%d" % k) for k in (1, 2, 3)]
    a["daywat"] = [irr.get(irow, "DayWat(%d)" % k) for k in range(1, 8)]
    a["daywatper"] = [irr.get(irow, "DayWatPer(%d)" % k) for k in range(1, 8)]

    # 24-hour profiles, (0:23, 2): weekday column then weekend column
    def prof24(code_wd, code_we, what):
        return list(prof.lookup(code_wd, what + " profile (WD)")[1:25]) + list(prof.lookup(code_we, what + " profile (WE)")[1:25])
    a["ahprof_24hr"] = prof24(A("EnergyUseProfWD"), A("EnergyUseProfWE"), "energy use")
    a["humactivity_24hr"] = prof24(A("ActivityProfWD"), A("ActivityProfWE"), "human activity")
    a["traffprof_24hr"] = prof24(A("TraffProfWD"), A("TraffProfWE"), "traffic")
    a["popprof_24hr"] = prof24(A("PopProfWD"), A("PopProfWE"), "population")
    a["wuprofm_24hr"] = prof24(S("WaterUseProfManuWD"), S("WaterUseProfManuWE"), "manual irrigation")
    a["wuprofa_24hr"] = prof24(S("WaterUseProfAutoWD"), S("WaterUseProfAutoWE"), "automatic irrigation")
    a["snowprof_24hr"] = prof24(S("SnowClearingProfWD"), S("SnowClearingProfWE"), "snow clearing")

    # within-grid water distribution: WaterDist(nsurf+1, nsurf-1), receiver x donor
    wd = np.zeros((8, 6))
    codes = ("WithinGridPavedCode", "WithinGridBldgsCode", "WithinGridEveTrCode", "WithinGridDecTrCode", "WithinGridGrassCode",
             "WithinGridUnmanBSoilCode")
    for j, cname in enumerate(codes):                       # donor surface j
        drow = dist.lookup(S(cname), "within-grid water distribution (%s)" % SURF[j])
        for i, to in enumerate(("ToPaved", "ToBldgs", "ToEveTr", "ToDecTr", "ToGrass", "ToBSoil", "ToWater")):
            wd[i, j] = dist.get(drow, to)
        wd[7, j] = dist.get(drow, "ToRunoff") if dist.get(drow, "ToRunoff") != 0 else dist.get(drow, "ToSoilStore")
    a["waterdist"] = list(wd.ravel(order="F"))

    # run switches
    for k in ("aerodynamicresistancemethod", "basetmethod", "diagnose", "diagqn", "diagqs", "emissionsmethod", "evapmethod",
              "laicalcyes", "netradiationmethod", "ohmincqf", "roughlenheatmethod", "roughlenmommethod", "smdmethod", "snowuse",
              "stabilitymethod", "storageheatmethod", "tstep", "veg_type", "waterusemethod"):
        a[k] = [int(rc[k])]
    a["gridiv"] = [1]
    a["_aux"] = dict(waterdepth=water.get(rows[6], "WaterDepth"), grid=int(S("Grid")), year=int(S("Year")),
                     areawall=S("AreaWall"),
                     estmcode=per_surf("ESTMCode") + [snow.get(snow_row, "ESTMCode")],
                     fr_estmclass_paved=[S("Fr_ESTMClass_Paved%d" % k) for k in (1, 2, 3)],
                     code_estmclass_paved=[S("Code_ESTMClass_Paved%d" % k) for k in (1, 2, 3)],
                     fr_estmclass_bldgs=[S("Fr_ESTMClass_Bldgs%d" % k) for k in range(1, 6)],
                     code_estmclass_bldgs=[S("Code_ESTMClass_Bldgs%d" % k) for k in range(1, 6)])
    return a


def estm_arguments(tables, args, aux, estm_nml):
    """ESTM's parameters as SUEWS_Translate (suews_ctrl_translate.f95:393-668) and ESTM_initials (suews_phys_estm.f95:589-634)
    leave them in MODULE ESTM_data, by the library's field names: layer thickness / conductivity / heat capacity of the roof,
    the wall, the internal elements and the ground facet (area-weighted over the non-building surfaces), layer counts, the
    indoor properties, AreaWall, the heating thresholds in K and the constants ESTM_translate hard-codes (:673-692).
    tables needs 'ESTMCoefficients' (load_tables(..., estm=True)); args / aux are site_arguments' results; estm_nml is
    ESTMinput.nml as read_runcontrol returns it.  An ESTMCode of 0 for the paved / building surface means the three / five
    ESTM classes of the SiteSelect row, weighted by their fractions."""
    E = tables["ESTMCoefficients"]
    sfr = np.asarray(args["sfr"], dtype=np.float64)

    def trio(row, stem):
        return (np.array([E.get(row, "%s_thick%d" % (stem, k)) for k in range(1, 6)]),
                np.array([E.get(row, "%s_k%d" % (stem, k)) for k in range(1, 6)]),
                np.array([E.get(row, "%s_rhoCp%d" % (stem, k)) for k in range(1, 6)]))

    def classes(codes, fracs, what):
        fr = np.asarray(fracs, dtype=np.float64)
        if not 0.999 <= fr.sum() <= 1.001:
            if (sfr[0] if what == "Paved" else sfr[1]) > 0:
                raise TableError("ErrorHint 10: surface fractions (Fr_ESTMClass_%s) should sum to 1 (%g)" % (what, fr.sum()))
            fr = np.array([1.0] + [0.0] * (len(fr) - 1))      # no such surface: the first class stands in (ErrorHint 67)
        return [E.lookup(c, "ESTM class (%s)" % what) for c in codes], fr

    def weighted(rows, fr, getter):
        acc = getter(rows[0]) * fr[0]
        for r, f in zip(rows[1:], fr[1:]):
            acc = acc + getter(r) * f
        return acc

    z = np.zeros((5, 8)); k = np.zeros((5, 8)); r = np.zeros((5, 8))
    codes = aux["estmcode"]
    out = {}
    # paved
    if codes[0] == 0:
        rows, fr = classes(aux["code_estmclass_paved"], aux["fr_estmclass_paved"], "Paved")
        for j, arr in enumerate((z, k, r)):
            arr[:, 0] = weighted(rows, fr, lambda row, j=j: trio(row, "Surf")[j])
    else:
        z[:, 0], k[:, 0], r[:, 0] = trio(E.lookup(codes[0], "ESTM (Paved)"), "Surf")
    # buildings: roof = the surface columns, plus wall and internal elements and the indoor scalars
    scal = (("nroom", "nroom"), ("alb_ibld", "Internal_albedo"), ("em_ibld", "Internal_emissivity"), ("CH_iwall", "Internal_CHwall"),
            ("CH_iroof", "Internal_CHroof"), ("CH_ibld", "Internal_CHbld"))
    if codes[1] == 0:
        rows, fr = classes(aux["code_estmclass_bldgs"], aux["fr_estmclass_bldgs"], "Bldgs")
        for j, arr in enumerate((z, k, r)):
            arr[:, 1] = weighted(rows, fr, lambda row, j=j: trio(row, "Surf")[j])
        wall = [weighted(rows, fr, lambda row, j=j: trio(row, "Wall")[j]) for j in range(3)]
        ibld = [weighted(rows, fr, lambda row, j=j: trio(row, "Internal")[j]) for j in range(3)]
        for nm, col in scal:
            out[nm] = [float(weighted(rows, fr, lambda row, col=col: np.float64(E.get(row, col))))]
    else:
        row = E.lookup(codes[1], "ESTM (Bldgs)")
        z[:, 1], k[:, 1], r[:, 1] = trio(row, "Surf")
        wall, ibld = trio(row, "Wall"), trio(row, "Internal")
        for nm, col in scal:
            out[nm] = [E.get(row, col)]
    for iv in range(2, 8):                                  # EveTr .. Water, snow
        z[:, iv], k[:, iv], r[:, iv] = trio(E.lookup(codes[iv], "ESTM (%s)" % (SURF[iv] if iv < 7 else "snow")), "Surf")
    fground = sfr[0] + sfr[2] + sfr[3] + sfr[4] + sfr[5] + sfr[6]
    zg = np.zeros(5); kg = np.zeros(5); rg = np.zeros(5)
    for iv in range(7):
        if iv != 1 and fground != 0:
            zg = zg + z[:, iv] * sfr[iv] / fground
            kg = kg + k[:, iv] * sfr[iv] / fground
            rg = rg + r[:, iv] * sfr[iv] / fground
        elif fground == 0:
            zg, kg, rg = zg + F32(0.01), kg + F32(0.01), rg + F32(0.01)

    def count(zz):
        for i in range(5):
            if zz[i] <= 0:
                return i
        return 5
    out.update(zground=list(zg), kground=list(kg), rground=list(rg), zroof=list(z[:, 1]), kroof=list(k[:, 1]), rroof=list(r[:, 1]),
               zwall=list(wall[0]), kwall=list(wall[1]), rwall=list(wall[2]), zibld=list(ibld[0]), kibld=list(ibld[1]), ribld=list(ibld[2]),
               Nground=[count(zg)], Nroof=[count(z[:, 1])], Nwall=[count(wall[0])], Nibld=[count(ibld[0])],
               AreaWall=[aux["areawall"]])
    c2k = F32(273.15)
    nml = {kk.lower(): v for kk, v in estm_nml.items()}
    out.update(TsurfChoice=[int(nml["tsurfchoice"])], evolveTibld=[int(nml["evolvetibld"])], ibldCHmod=[int(nml["ibldchmod"])],
               LBC_soil=[float(nml["lbc_soil"])], THEAT_ON=[float(nml["theat_on"]) + c2k], THEAT_OFF=[float(nml["theat_off"]) + c2k],
               THEAT_fix=[float(nml["theat_fix"]) + c2k])
    # internal view factors and the floor temperature of ESTM_translate (suews_phys_estm.f95:673-692)
    out.update(ivf_iw=[F32(0.1)], ivf_ir=[0.0], ivf_ii=[F32(0.9)], ivf_if=[0.0], ivf_ww=[F32(0.05)], ivf_wr=[0.0], ivf_wi=[F32(0.95)],
               ivf_wf=[0.0], ivf_rw=[F32(0.05)], ivf_ri=[F32(0.95)], ivf_rf=[0.0], ivf_fw=[F32(0.05)], ivf_fr=[0.0], ivf_fi=[F32(0.95)],
               Tfloor=[F32(20.0) + c2k])
    return out


def F32(x):
    return float(np.float32(x))


def read_cbl(input_dir):
    """BLUEWS inputs of a run directory (CBL_ReadInputData, suews_phys_bluews.f95:300-357): the switches of CBLInput.nml as
    class fields, and IniCBLdata -- the rows of its InitialDataFileName up to a -9 line, [nline, 8]."""
    nml = read_runcontrol(os.path.join(input_dir, "CBLInput.nml"))
    params = {k: [nml[k]] for k in ("entrainmenttype", "qh_choice", "isubs", "co2_included", "initialdata_use", "sondeflag", "wsb")
              if k in nml}
    ini = None
    if int(nml.get("initialdata_use", 0)) in (1, 2):
        rows = []
        for ln in open(os.path.join(input_dir, nml["initialdatafilename"])).read().splitlines()[1:]:
            tok = ln.split()
            if not tok or float(tok[0]) == -9:
                break
            rows.append([float(t) for t in tok[:8]])
        ini = np.array(rows, dtype=np.float64)
    return params, ini


def run_arguments(run_dir, grid, year, initial_conditions=None, estm_ts_row0=None):
    """Everything SuewsBatch.load_reference_args needs for one grid of a reference run directory: the table join above plus
    the first step's state from an InitialConditions namelist (checkpoint.initial_state; default
    Input/InitialConditions<FileCode>_<year>.nml).  With StorageHeatMethod 4 / 14 in RunControl.nml the ESTM parameters
    (estm_arguments, ESTMinput.nml) are included, and -- given the first row [13] of the surface-temperature forcing -- the
    element temperatures ESTM_translate starts from (estm.initial_state)."""
    from . import checkpoint
    rc = read_runcontrol(os.path.join(run_dir, "RunControl.nml"))
    input_dir = os.path.join(run_dir, rc.get("fileinputpath", "./Input/"))
    with_estm = int(rc.get("storageheatmethod", 1)) in (4, 14)
    tabs = load_tables(input_dir, estm=with_estm)
    a = site_arguments(tabs, rc, grid, year)
    aux = a.pop("_aux")
    if with_estm:
        e = estm_arguments(tabs, a, aux, read_runcontrol(os.path.join(input_dir, "ESTMinput.nml")))
        a.update({k.lower(): v for k, v in e.items()})
        if estm_ts_row0 is not None:
            from . import estm
            a.update({k.lower(): list(np.atleast_1d(v)) for k, v in estm.initial_state(e, a["sfr"], a["emis"], estm_ts_row0).items()})
    if initial_conditions is None:
        initial_conditions = os.path.join(input_dir, "InitialConditions%s_%d.nml" % (rc.get("filecode", ""), year))
    st = checkpoint.initial_state(checkpoint.read(initial_conditions), a, water_depth=aux["waterdepth"])
    a.update({k.lower(): list(np.atleast_1d(v)) for k, v in st.items()})
    # (alb keeps the AlbMax defaults of SUEWS_Translate on the first step: SUEWS_cal_Qn puts the day's vegetation albedos
    # into alb_next during the step, driver.f95:1409-1413)
    return a, aux, rc


def grid_arguments(run_dir, year, grids=None, initial_conditions=None):
    """The grid loop of the reference's main program (suews_program.f95: one SUEWS_Translate + SUEWS_cal_Main per SiteSelect
    row of the year) as the batch layout: every grid of `year` in SiteSelect -> one cell; grids whose class-kind values (the
    scheme switches and the code-keyed tables they share) are equal -> one class.

    Returns {"grids": [ids], "class_of_grid": int32[ngrid], "classes": [{name: values}], "cells": {name: [ngrid, n]},
             "state": {name: [ngrid, n]}, "clock": (dt_since_start, tstep_prev), "runcontrol": rc, "aux": [aux per grid],
             "cbl_initial": the rows of the CBL initial-data file (CBLuse >= 1; for set_cbl_initial after upload) or None}
    with names as suews_fields.h spells them.  `initial_conditions`: a path used for every grid, a function grid -> path, or
    None for the reference's own rule (MultipleInitFiles = 1: InitialConditions<FileCode><grid>_<year>.nml, else the shared
    InitialConditions<FileCode>_<year>.nml)."""
    from .batch import KIND_CELL, KIND_CLASS_DBL, KIND_CLASS_INT, all_fields
    rc = read_runcontrol(os.path.join(run_dir, "RunControl.nml"))
    input_dir = os.path.join(run_dir, rc.get("fileinputpath", "./Input/"))
    if grids is None:
        ss = Table(os.path.join(input_dir, "SUEWS_SiteSelect.txt"))
        grids = [int(r[0]) for r in ss.rows if int(r[1]) == int(year)]   # (columns Grid, Year)
        if not grids:
            raise TableError("SUEWS_SiteSelect.txt has no row for year %d" % year)
    if len(set(grids)) != len(grids):
        raise TableError("SUEWS_SiteSelect.txt lists a grid twice for year %d" % year)

    def ic_path(g):
        if callable(initial_conditions):
            return initial_conditions(g)
        if initial_conditions is not None:
            return initial_conditions
        code = rc.get("filecode", "")
        return os.path.join(input_dir, "InitialConditions%s%s_%d.nml" % (code, g if int(rc.get("multipleinitfiles", 0)) == 1 else "", year))

    fields = all_fields()
    cbl_on = int(rc.get("cbluse", 0)) >= 1
    cbl_params, cbl_ini = read_cbl(input_dir) if cbl_on else ({}, None)
    keys, classes, class_of_grid, auxs = {}, [], [], []
    cells = {nm: [] for nm, kind, _n in fields if kind == KIND_CELL}
    state = {nm: [] for nm, kind, _n in fields if kind not in (KIND_CLASS_INT, KIND_CLASS_DBL, KIND_CELL)}
    clock = None
    for g in grids:
        a, aux, _rc = run_arguments(run_dir, g, year, initial_conditions=ic_path(g))
        if cbl_on:
            a.update({k.lower(): v for k, v in cbl_params.items()})
            a["cbluse"] = [int(rc["cbluse"])]
        auxs.append(aux)
        cls = {}
        for nm, kind, n in fields:
            if nm.lower() not in a:
                continue
            v = np.atleast_1d(np.asarray(a[nm.lower()], dtype=np.float64)).ravel()
            if v.size != n:
                raise TableError("%s: %d values from the tables, the library holds %d" % (nm, v.size, n))
            if kind in (KIND_CLASS_INT, KIND_CLASS_DBL):
                cls[nm] = v
            elif kind == KIND_CELL:
                cells[nm].append(v)
            else:
                state[nm].append(v)
        key = tuple((nm, v.tobytes()) for nm, v in sorted(cls.items()))
        if key not in keys:
            keys[key] = len(classes)
            classes.append(cls)
        class_of_grid.append(keys[key])
        c = (int(a["dt_since_start"][0]), int(a.get("tstep_prev", a["tstep"])[0]))
        if clock is not None and c != clock:
            raise TableError("grids of one run start from different clocks")
        clock = c
    ng = len(grids)
    for d in (cells, state):
        for nm in [k for k, v in d.items() if len(v) not in (0, ng)]:
            raise TableError("%s is set for some grids only" % nm)
    return {"grids": list(grids), "class_of_grid": np.asarray(class_of_grid, dtype=np.int32), "classes": classes,
            "cells": {k: np.vstack(v) for k, v in cells.items() if v}, "state": {k: np.vstack(v) for k, v in state.items() if v},
            "clock": clock, "runcontrol": rc, "aux": auxs, "cbl_initial": cbl_ini}


def load_batch(batch, ga):
    """grid_arguments' result into a SuewsBatch / SuewsMultiBatch made for len(ga["grids"]) cells and len(ga["classes"])
    classes (cell i = grid ga["grids"][i]); the caller uploads."""
    for c, cls in enumerate(ga["classes"]):
        for nm, v in cls.items():
            batch.set_class_param(c, nm, v)
    batch.set_cell_class(ga["class_of_grid"])
    for nm, v in ga["cells"].items():
        batch.set_cell_param(nm, v)
    for nm, v in ga["state"].items():
        batch.set_state(nm, v)
    batch.set_clock(*ga["clock"])
    return batch
