// suews_fields.h -- parameter / state layout of the B200 SUEWS library.
//
// The reference passes ~220 arguments per cell-step (SUEWS_cal_Main,
// SUEWS-SourceCode/suews_ctrl_driver.f95:45-91).  In its own input model they fall in two groups:
// per-grid site columns (SUEWS_SiteSelect.txt) and code-keyed class tables shared by many grids
// (suews_ctrl_input.f95:832-1321).  The library keeps that split in HBM:
//
//   ClassParams  one record per distinct class, AoS; every thread of a warp reads the same address
//                (cells are sorted by class), so a class read is one broadcast transaction.
//   cell params  SoA [field][cell]  -> coalesced.
//   cell state   SoA [field][cell]  -> coalesced, loaded once per launch, kept on-chip for the chunk.
//
// Names are the reference's argument names, so the C-ABI can be driven by name.
#pragma once

#define SUEWS_NSURF 7
#define SUEWS_NVEG 3

// ---- class-level integers: X(name, count)
#define SUEWS_CLASS_INT_FIELDS(X) \
   X(AerodynamicResistanceMethod, 1) X(BaseTMethod, 1) X(Diagnose, 1) X(DiagQN, 1) X(DiagQS, 1) \
   X(startDLS, 1) X(endDLS, 1) X(EmissionsMethod, 1) X(gsModel, 1) X(Ie_end, 1) X(Ie_start, 1) \
   X(EvapMethod, 1) X(LAICalcYes, 1) X(NetRadiationMethod, 1) X(OHMIncQF, 1) X(RoughLenHeatMethod, 1) \
   X(RoughLenMomMethod, 1) X(SMDMethod, 1) X(snowUse, 1) X(StabilityMethod, 1) X(StorageHeatMethod, 1) \
   X(tstep, 1) X(veg_type, 1) X(WaterUseMethod, 1) X(LAIType, 3) \
   /* MODULE ESTM_data (suews_phys_estm.f95:1-176): ESTMinput.nml switches and layer counts */ \
   X(evolveTibld, 1) X(TsurfChoice, 1) X(ibldCHmod, 1) X(Nibld, 1) X(Nwall, 1) X(Nroof, 1) X(Nground, 1) \
   /* BLUEWS (suews_phys_bluews.f95): RunControl's CBLuse and CBLInput.nml */ \
   X(CBLuse, 1) X(EntrainmentType, 1) X(QH_choice, 1) X(isubs, 1) X(CO2_included, 1) X(InitialData_use, 1) X(sondeflag, 1)

// ---- class-level doubles
#define SUEWS_CLASS_DBL_FIELDS(X) \
   X(AlbMax_DecTr, 1) X(AlbMax_EveTr, 1) X(AlbMax_Grass, 1) X(AlbMin_DecTr, 1) X(AlbMin_EveTr, 1) \
   X(AlbMin_Grass, 1) X(BaseT_HC, 1) X(CapMax_dec, 1) X(CapMin_dec, 1) X(CO2PointSource, 1) X(CRWmax, 1) \
   X(CRWmin, 1) X(EF_umolCO2perJ, 1) X(EnEF_v_Jkm, 1) X(Faut, 1) X(FrFossilFuel_Heat, 1) \
   X(FrFossilFuel_NonHeat, 1) X(G1, 1) X(G2, 1) X(G3, 1) X(G4, 1) X(G5, 1) X(G6, 1) X(H_maintain, 1) \
   X(InternalWaterUse_h, 1) X(Kmax, 1) X(MaxFCMetab, 1) X(MaxQFMetab, 1) X(MinFCMetab, 1) X(MinQFMetab, 1) \
   X(NARP_EMIS_SNOW, 1) X(PorMax_dec, 1) X(PorMin_dec, 1) X(PrecipLimit, 1) X(PrecipLimitAlb, 1) \
   X(RadMeltFact, 1) X(S1, 1) X(S2, 1) X(SnowAlbMax, 1) X(SnowAlbMin, 1) X(SnowDensMax, 1) X(SnowDensMin, 1) \
   X(SnowLimBldg, 1) X(SnowLimPaved, 1) X(tau_a, 1) X(tau_f, 1) X(tau_r, 1) X(TempMeltFact, 1) X(TH, 1) \
   X(TL, 1) X(TrafficUnits, 1) \
   X(AH_MIN, 2) X(AH_SLOPE_Cooling, 2) X(AH_SLOPE_Heating, 2) X(FcEF_v_kgkm, 2) X(Qf_A, 2) X(Qf_B, 2) \
   X(Qf_C, 2) X(BaseT_Cooling, 2) X(BaseT_Heating, 2) X(Ie_a, 3) X(Ie_m, 3) X(MaxConductance, 3) \
   X(DayWat, 7) X(DayWatPer, 7) X(OHM_threshSW, 8) X(OHM_threshWD, 8) X(chAnOHM, 7) X(cpAnOHM, 7) X(emis, 7) \
   X(kkAnOHM, 7) X(SatHydraulicConduct, 7) X(SnowPackLimit, 7) X(SoilDepth, 7) X(SoilStoreCap, 7) \
   X(StateLimit, 7) X(WetThresh, 7) X(alpha_bioCO2, 3) X(alpha_enh_bioCO2, 3) X(BaseT, 3) X(BaseTe, 3) \
   X(beta_bioCO2, 3) X(beta_enh_bioCO2, 3) X(GDDFull, 3) X(LAIMax, 3) X(LAIMin, 3) X(min_res_bioCO2, 3) \
   X(resp_a, 3) X(resp_b, 3) X(SDDFull, 3) X(theta_bioCO2, 3) X(SnowProf_24hr, 48) X(LAIPower, 12) \
   X(OHM_coef, 96) X(WaterDist, 48) X(AHProf_24hr, 48) X(HumActivity_24hr, 48) X(PopProf_24hr, 48) \
   X(TraffProf_24hr, 48) X(WUProfA_24hr, 48) X(WUProfM_24hr, 48) \
   /* MODULE ESTM_data as SUEWS_Translate / ESTM_initials leave it (THEAT_* in K, layers of the four elements) */ \
   X(LBC_soil, 1) X(THEAT_ON, 1) X(THEAT_OFF, 1) X(THEAT_fix, 1) X(ivf_iw, 1) X(ivf_ir, 1) X(ivf_ii, 1) \
   X(ivf_if, 1) X(ivf_ww, 1) X(ivf_wr, 1) X(ivf_wi, 1) X(ivf_wf, 1) X(ivf_rw, 1) X(ivf_ri, 1) X(ivf_rf, 1) \
   X(ivf_fw, 1) X(ivf_fr, 1) X(ivf_fi, 1) X(zibld, 5) X(kibld, 5) X(ribld, 5) X(zwall, 5) X(kwall, 5) X(rwall, 5) \
   X(zroof, 5) X(kroof, 5) X(rroof, 5) X(zground, 5) X(kground, 5) X(rground, 5) X(nroom, 1) X(alb_ibld, 1) \
   X(em_ibld, 1) X(CH_iroof, 1) X(CH_iwall, 1) X(CH_ibld, 1) X(Tfloor, 1) \
   X(wsb, 1) /* BLUEWS: subsidence velocity (CBLInput.nml) */

// ---- per-cell doubles (the SUEWS_SiteSelect.txt site columns)
// HOT fields are the small per-surface arrays every stage of a step indexes with a run-time surface number: on the device
// they live in shared memory (a per-thread record of SUEWS_N_HOT_DBL doubles, suews_kernel.cuh), the COLD ones in the
// thread's own frame.  The lists are COLD first, HOT last, and the SoA arrays of the handle follow that order.
#define SUEWS_CELL_HOT_FIELDS(X) X(sfr, 7)
#define SUEWS_CELL_COLD_FIELDS(X) \
   X(lat, 1) X(lng, 1) X(timezone, 1) X(SurfaceArea, 1) X(alt, 1) X(Z, 1) X(IrrFracPaved, 1) \
   X(IrrFracBldgs, 1) X(IrrFracEveTr, 1) X(IrrFracDecTr, 1) X(IrrFracGrass, 1) X(IrrFracBSoil, 1) \
   X(IrrFracWater, 1) X(bldgH, 1) X(EveTreeH, 1) X(DecTreeH, 1) X(z0m_in, 1) X(zdm_in, 1) X(FAIBldg, 1) \
   X(FAIEveTree, 1) X(FAIDecTree, 1) X(PopDensDaytime, 2) X(PopDensNighttime, 1) X(TrafficRate, 2) \
   X(QF0_BEU, 2) X(DRAINRT, 1) X(RAINCOVER, 1) X(RainMaxRes, 1) X(NARP_TRANS_SITE, 1) X(FlowChange, 1) \
   X(RunoffToWater, 1) X(PipeCapacity, 1) \
   X(dTemp_C, 1) X(dRH, 1) /* optional additive forcing offsets (synthetic configs) */ \
   X(AreaWall, 1)          /* ESTM: fwall = AreaWall / SurfaceArea (suews_ctrl_translate.f95:393-394) */
#define SUEWS_CELL_DBL_FIELDS(X) SUEWS_CELL_COLD_FIELDS(X) SUEWS_CELL_HOT_FIELDS(X)

// ---- evolving per-cell state (the 29 INTENT(INOUT) arguments, driver.f95:290-333, + hidden SAVEs)
#define SUEWS_STATE_HOT_FIELDS(X) X(state_id, 7) X(soilstore_id, 7)
#define SUEWS_STATE_COLD_FIELDS(X) \
   X(qn1_av, 1) X(dqndt, 1) X(qn1_s_av, 1) X(dqnsdt, 1) X(SnowfallCum, 1) X(SnowAlb, 1) X(IceFrac, 7) \
   X(SnowWater, 7) X(SnowDens, 7) X(SnowFrac, 7) X(SnowPack, 7) \
   X(StoreDrainPrm, 42) X(alb, 7) X(GDD_id, 3) X(SDD_id, 3) X(LAI_id, 3) X(Tmin_id, 1) X(Tmax_id, 1) \
   X(lenDay_id, 1) X(DecidCap_id, 1) X(albDecTr_id, 1) X(albEveTr_id, 1) X(albGrass_id, 1) X(porosity_id, 1) \
   X(HDD_id, 12) X(WUDay_id, 9) X(Tair_av, 1) X(alb1_hidden, 1) X(lv_switch, 1) \
   X(rsl_cache, 10) /* library-internal: last z0_RSL per stability sign (pure-function memo, dev_rsl.cuh) */
#define SUEWS_STATE_DBL_FIELDS(X) SUEWS_STATE_COLD_FIELDS(X) SUEWS_STATE_HOT_FIELDS(X)

// ---- ESTM state: what MODULE ESTM_data carries from call to call (the *_grids copies, estm.f95:1587-1603, and
// the SAVE'd HVAC, Tair2Set and NewtonPolynomial's `converged`).  A separate SoA [field][cell], resident only when
// a class runs StorageHeatMethod 4/14.  Tw_4 is (layer, wall) column-major with leading dimension 5.
// The same record carries what AnOHM (StorageHeatMethod 3) keeps in SAVE'd locals (suews_phys_anohm.f95:79, :215-216): the last
// day of year with enough daylight rows, the day and grid of the cached coefficients, and coeff_grid_day(7,3) (initially
// -999) -- the record is "the storage-heat scheme's memory outside the SUEWS_cal_Main argument list" and is resident when
// a class runs StorageHeatMethod 3, 4 or 14.
#define SUEWS_ESTM_STATE_FIELDS(X) \
   X(Tair2, 1) X(Tievolve, 1) X(lup_ground, 1) X(lup_wall, 1) X(lup_roof, 1) X(T0_ibld, 1) X(T0_ground, 1) \
   X(T0_wall, 1) X(T0_roof, 1) X(TN_wall, 1) X(TN_roof, 1) X(Tground, 5) X(Twall, 5) X(Troof, 5) X(Tibld, 5) \
   X(Tw_4, 20) X(HVAC, 1) X(Tair2Set, 1) X(NewtonConverged, 1) \
   X(anohm_id_save, 1) X(anohm_coef_id_save, 1) X(anohm_coef_grid_save, 1) X(anohm_coeff_grid_day, 21) \
   /* BLUEWS: MODULE cbl_MODULE between calls (suews_phys_bluews.f95:24-47) */ \
   X(cbl_blh_m, 1) X(cbl_tm_K, 1) X(cbl_qm_kgkg, 1) X(cbl_tp_K, 1) X(cbl_qp_kgkg, 1) X(cbl_cm, 1) X(cbl_gamt_Km, 1) X(cbl_gamq_kgkgm, 1)

#define SUEWS_COUNT_X(name, n) +(n)
enum {
   SUEWS_N_ESTM_DBL = 0 SUEWS_ESTM_STATE_FIELDS(SUEWS_COUNT_X),
   SUEWS_N_CLASS_INT = 0 SUEWS_CLASS_INT_FIELDS(SUEWS_COUNT_X),
   SUEWS_N_CLASS_DBL = 0 SUEWS_CLASS_DBL_FIELDS(SUEWS_COUNT_X),
   SUEWS_N_CELL_DBL = 0 SUEWS_CELL_DBL_FIELDS(SUEWS_COUNT_X),
   SUEWS_N_STATE_DBL = 0 SUEWS_STATE_DBL_FIELDS(SUEWS_COUNT_X),
   SUEWS_N_CELL_HOT = 0 SUEWS_CELL_HOT_FIELDS(SUEWS_COUNT_X),
   SUEWS_N_STATE_HOT = 0 SUEWS_STATE_HOT_FIELDS(SUEWS_COUNT_X),
   SUEWS_N_CELL_COLD = SUEWS_N_CELL_DBL - SUEWS_N_CELL_HOT,
   SUEWS_N_STATE_COLD = SUEWS_N_STATE_DBL - SUEWS_N_STATE_HOT,
   SUEWS_N_HOT_DBL = SUEWS_N_CELL_HOT + SUEWS_N_STATE_HOT   /* 21: an odd stride in doubles keeps the per-thread records bank-conflict free */
};

// Output widths without the 5 datetime columns (suews_ctrl_const.f95:49-55)
enum {
   SUEWS_NCOL_SUEWS = 82, SUEWS_NCOL_SNOW = 98, SUEWS_NCOL_ESTM = 27, SUEWS_NCOL_RSL = 133,
   SUEWS_NCOL_SOLWEIG = 26, SUEWS_NCOL_DAILY = 45, SUEWS_NCOL_MET = 24, SUEWS_NCOL_BL = 17
};

// ---- structs generated from the lists
// A field of n values: n == 1 reads like a scalar, n > 1 like an array.  Plain data (memcpy-able).
#if defined(__CUDACC__)
#define SUEWS_HD __host__ __device__ __forceinline__
#else
#define SUEWS_HD inline
#endif
template <typename T, int N> struct SuewsFld {
   T v[N];
   SUEWS_HD T &operator[](int i) { return v[i]; }
   SUEWS_HD const T &operator[](int i) const { return v[i]; }
};
template <typename T> struct SuewsFld<T, 1> {
   T v[1];
   SUEWS_HD operator T() const { return v[0]; }
   SUEWS_HD SuewsFld &operator=(T x) { v[0] = x; return *this; }
};
// a HOT field: same spelling at the use sites (name[i], name.v), storage elsewhere (shared memory)
template <typename T, int N> struct SuewsRef {
   T *v;
   SUEWS_HD T &operator[](int i) { return v[i]; }
   SUEWS_HD const T &operator[](int i) const { return v[i]; }
};
#define SUEWS_DECL_INT(name, n) SuewsFld<int, n> name;
#define SUEWS_DECL_DBL(name, n) SuewsFld<double, n> name;
#define SUEWS_DECL_REF(name, n) SuewsRef<double, n> name;

struct SuewsClassParams {
   SUEWS_CLASS_INT_FIELDS(SUEWS_DECL_INT)
   int pad_;
   SUEWS_CLASS_DBL_FIELDS(SUEWS_DECL_DBL)
   // derived at upload (bit-identical to what the reference recomputes every call):
   double prof_sum[6][2];   // sequential SUM of AH, HumAct, Pop, Traff, WUProfA, WUProfM (weekday, weekend)
   double gs_pow_TH_G5;     // (TH - G5)**TC, TC = (TH - G5)/(G5 - TL): the class constant inside SurfaceResistance's TC2 (resist.f95:269-270)
};

// device-side records of one cell: the COLD fields by value (the first SUEWS_N_*_COLD doubles of the struct, in list
// order), the HOT ones as references into the thread's shared-memory record
struct SuewsCellParams {
   SUEWS_CELL_COLD_FIELDS(SUEWS_DECL_DBL)
   SUEWS_CELL_HOT_FIELDS(SUEWS_DECL_REF)
};

struct SuewsCellState {
   SUEWS_STATE_COLD_FIELDS(SUEWS_DECL_DBL)
   SUEWS_STATE_HOT_FIELDS(SUEWS_DECL_REF)
};

struct SuewsEstmState {
   SUEWS_ESTM_STATE_FIELDS(SUEWS_DECL_DBL)
};
enum { SUEWS_N_TS5 = 9 };   // Tiair Tsurf Troof Troad Twall Twall_n Twall_e Twall_s Twall_w (Ts5mindata_ir(5:13))

// One record per timestep, shared by all cells: clock + forcing (driver.f95:3177-3201).
// Staged into shared memory in chunks by bulk async copy.  160 bytes.
struct SuewsStepRec {
   int iy, id, it, imin;           // clock
   int dow, month;                 // day of week 1=Sun..7=Sat, month (SUEWS_cal_weekday, util_time.f95:301)
   int pad0_, pad1_;               // keep the record a multiple of 16 bytes for cp.async.bulk
   double dectime;                 // SUEWS_cal_dectime
   double avkdn, avRh, avU1, fcld_obs, LAI_obs, ldown_obs, Precip, Press_hPa;
   double qn1_obs, qs_obs, qf_obs, snowFrac_obs, Temp_C, wu_m3, xsmd;
};

// Time-only part of the NREL SPA (steps 1-10, narp.f95:460-548), once per step per time zone.
struct SuewsEphem {
   double app_stime;   // apparent sidereal time at Greenwich [deg]
   double sun_ra;      // geocentric right ascension [deg]
   double sun_dec;     // geocentric declination [deg]
   double eh_rad;      // earth heliocentric radius [AU]
};
