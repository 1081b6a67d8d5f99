/*
 * suews_oracle.h -- CPU ORACLE for the SUEWS per-grid-cell timestep physics.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke()
 * entry and bench.py's cpu_baseline/reference leg may build, link or call
 * anything under oracle/.  The product path (suews_b200/, libsuews_b200.so)
 * never links against it.
 *
 * It is a plain-C restatement of the reference algorithm
 * (SUEWS-SourceCode/suews_ctrl_driver.f95 SUEWS_cal_Main :45-1045 and the
 * physics modules it calls), literal-by-literal: every un-suffixed Fortran
 * literal is a REAL(4) constant that is widened to double at use, so it is
 * spelled F(c) == (double)(c##f) here; integer sub-expressions keep their
 * truncating semantics; evaluation order follows Fortran's left-to-right
 * rule; x**n with integer n is a multiplication chain, x**r is pow().
 *
 * Parity pin: the reference cannot be compiled in this image (no Fortran
 * compiler of any kind).  The oracle is pinned against the reference's own
 * golden run Test/BaseRun/2020b (hourly text outputs, 4-6 decimals) replayed
 * from its full-precision argument dump -- see tests/test_oracle_golden.py.
 */
#ifndef SUEWS_ORACLE_H
#define SUEWS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define F(c) ((double)(c##f)) /* REAL(4) literal widened to REAL(8) */

/* Op counter ("the oracle carries an op counter").  `make count` builds libsuews_oracle_count.so from these
 * same sources as C++ with `double` replaced by a counting type (so_count.h): every FP64 add / multiply / divide / sqrt
 * and every libm call the restated algorithm executes is counted, none of the arithmetic changes.  The regions below
 * only steer where the counts go; in the normal C build they are empty. */
#ifdef SO_OPCOUNT
extern int so_opcount_on, so_opcount_hoisted;
extern unsigned long long *so_opcount_cur, so_opcount_main[16], so_opcount_diag[16];
#define SO_CNT_ONCE_BEGIN(cond) { int oc_save_ = so_opcount_on; if (so_opcount_hoisted && (cond)) so_opcount_on = 0
#define SO_CNT_END so_opcount_on = oc_save_; }
#define SO_CNT_DIAG_BEGIN() (so_opcount_cur = so_opcount_diag)
#define SO_CNT_DIAG_END() (so_opcount_cur = so_opcount_main)
#else
#define SO_CNT_ONCE_BEGIN(cond) {
#define SO_CNT_END }
#define SO_CNT_DIAG_BEGIN() ((void)0)
#define SO_CNT_DIAG_END() ((void)0)
#endif

enum { NSURF = 7, NVEG = 3 };
/* 0-based surface indices (suews_ctrl_const.f95:188-199 are 1-based) */
enum { PAV = 0, BLDG = 1, CONIF = 2, DECID = 3, GRASS = 4, BSOIL = 5, WATER = 6 };
enum { IVCONIF = 0, IVDECID = 1, IVGRASS = 2 };

/* output widths without the 5 datetime columns (suews_ctrl_const.f95:49-55) */
enum {
   NCOL_SUEWS = 82, NCOL_SNOW = 98, NCOL_ESTM = 27, NCOL_RSL = 133,
   NCOL_SOLWEIG = 26, NCOL_DAILY = 45
};

/* ESTM (StorageHeatMethod 4/14) does not receive its parameters and state through the SUEWS_cal_Main argument
 * list: they are variables of MODULE ESTM_data (suews_phys_estm.f95:1-176), filled by ESTM_initials (:589-634,
 * namelist ESTMinput.nml), by SUEWS_Translate (suews_ctrl_translate.f95:393-394, :470-670: facet layers) and by
 * ESTM_translate (:637-967).  Here the facet-level values after SUEWS_Translate are parameters (SoEstmParams) */
typedef struct {
   int evolveTibld, TsurfChoice, ibldCHmod;   /* ESTMinput.nml */
   int Nibld, Nwall, Nroof, Nground;          /* layers per element (suews_ctrl_translate.f95:641-670) */
   double LBC_soil;                            /* deg C */
   double THEAT_ON, THEAT_OFF, THEAT_fix;      /* K (converted by ESTM_initials, estm.f95:617-619) */
   double ivf_iw, ivf_ir, ivf_ii, ivf_if, ivf_ww, ivf_wr, ivf_wi, ivf_wf, ivf_rw, ivf_ri, ivf_rf, ivf_fw, ivf_fr, ivf_fi;
   double zibld[5], kibld[5], ribld[5], zwall[5], kwall[5], rwall[5], zroof[5], kroof[5], rroof[5];
   double zground[5], kground[5], rground[5];
   double nroom, alb_ibld, em_ibld, CH_iroof, CH_iwall, CH_ibld;
   double AreaWall;                            /* fwall = AreaWall/SurfaceArea (suews_ctrl_translate.f95:393-394) */
   double Tfloor;                              /* K (estm.f95:673-674: 20 degC) */
} SoEstmParams;

/* and the variables ESTM carries from call to call are state (the *_grids copies, estm.f95:1587-1603, and the
 * SAVE'd HVAC / Tair2Set / NewtonPolynomial's `converged`). */
typedef struct {
   double Tair2, Tievolve, lup_ground, lup_wall, lup_roof, T0_ibld, T0_ground, T0_wall, T0_roof, TN_wall, TN_roof;
   double Tground[5], Twall[5], Troof[5], Tibld[5];
   double Tw_4[20];            /* (Nwall,4) column-major with leading dimension 5 here */
   double HVAC;                /* LOGICAL, SAVE (estm.f95:165) */
   double Tair2Set;            /* INTEGER :: Tair2Set = 0, SAVE (estm.f95:1162) */
   double NewtonConverged;     /* LOGICAL :: converged = .FALSE., SAVE (estm.f95:222): once true, true for ever */
} SoEstmState;

/* AnOHM (StorageHeatMethod 3) keeps its memory in SAVE'd locals: AnOHM's id_save (suews_phys_anohm.f95:79, the last
 * day whose forcing had enough daylight rows) and AnOHM_coef's id_save / grid_save / coeff_grid_day(7,3) (:215-216,
 * the day's coefficients, formed on the first call of a day and reloaded afterwards).  gfortran gives uninitialised
 * SAVE integers static storage, i.e. 0; coeff_grid_day is initialised to -999. */
typedef struct {
   double id_save;                        /* AnOHM */
   double coef_id_save, coef_grid_save;   /* AnOHM_coef */
   double coeff_grid_day[21];             /* (7,3) column-major: [is + 7*k] */
} SoAnohmState;

/* BLUEWS (suews_phys_bluews.f95): CBLInput.nml, and the slab's variables in MODULE cbl_MODULE between calls */
typedef struct {
   int CBLuse, EntrainmentType, QH_choice, isubs, CO2_included, InitialData_use, sondeflag;
   double wsb;
} SoCblParams;
typedef struct {
   double blh_m, tm_K, qm_kgkg, tp_K, qp_kgkg, cm, gamt_Km, gamq_kgkgm;
} SoCblState;

/* Static per-cell configuration: the INTENT(IN) arguments of SUEWS_cal_Main
 * that do not change from step to step (suews_ctrl_driver.f95:97-285).
 * Arrays keep Fortran column-major storage; see the index macros below. */
typedef struct {
   int AerodynamicResistanceMethod, BaseTMethod, Diagnose, DiagQN, DiagQS;
   int startDLS, endDLS, EmissionsMethod, Gridiv, gsModel, Ie_end, Ie_start;
   int EvapMethod, LAICalcYes, NetRadiationMethod, OHMIncQF;
   int RoughLenHeatMethod, RoughLenMomMethod, SMDMethod, snowUse;
   int StabilityMethod, StorageHeatMethod, tstep, veg_type, WaterUseMethod;
   int LAIType[NVEG];

   double AlbMax_DecTr, AlbMax_EveTr, AlbMax_Grass, AlbMin_DecTr, AlbMin_EveTr, AlbMin_Grass;
   double alt, BaseT_HC, bldgH, CapMax_dec, CapMin_dec, CO2PointSource, CRWmax, CRWmin;
   double DecTreeH, DRAINRT, EF_umolCO2perJ, EnEF_v_Jkm, EveTreeH, FAIBldg, FAIDecTree, FAIEveTree;
   double Faut, FlowChange, FrFossilFuel_Heat, FrFossilFuel_NonHeat;
   double G1, G2, G3, G4, G5, G6, H_maintain, InternalWaterUse_h;
   double IrrFracPaved, IrrFracBldgs, IrrFracEveTr, IrrFracDecTr, IrrFracGrass, IrrFracBSoil, IrrFracWater;
   double Kmax, lat, lng, MaxFCMetab, MaxQFMetab, MinFCMetab, MinQFMetab;
   double NARP_EMIS_SNOW, NARP_TRANS_SITE, PipeCapacity, PopDensNighttime, PorMax_dec, PorMin_dec;
   double PrecipLimit, PrecipLimitAlb, RadMeltFact, RAINCOVER, RainMaxRes, RunoffToWater, S1, S2;
   double SnowAlbMax, SnowAlbMin, SnowDensMax, SnowDensMin, SnowLimBldg, SnowLimPaved, SurfaceArea;
   double tau_a, tau_f, tau_r, TempMeltFact, TH, timezone, TL, TrafficUnits, Z, z0m_in, zdm_in;

   double AH_MIN[2], AH_SLOPE_Cooling[2], AH_SLOPE_Heating[2], FcEF_v_kgkm[2], QF0_BEU[2];
   double Qf_A[2], Qf_B[2], Qf_C[2], PopDensDaytime[2], BaseT_Cooling[2], BaseT_Heating[2], TrafficRate[2];
   double Ie_a[3], Ie_m[3], MaxConductance[3], DayWat[7], DayWatPer[7];
   double OHM_threshSW[NSURF + 1], OHM_threshWD[NSURF + 1];
   double chAnOHM[NSURF], cpAnOHM[NSURF], emis[NSURF], kkAnOHM[NSURF], SatHydraulicConduct[NSURF];
   double sfr[NSURF], SnowPackLimit[NSURF], SoilDepth[NSURF], SoilStoreCap[NSURF], StateLimit[NSURF], WetThresh[NSURF];
   double alpha_bioCO2[NVEG], alpha_enh_bioCO2[NVEG], BaseT[NVEG], BaseTe[NVEG], beta_bioCO2[NVEG], beta_enh_bioCO2[NVEG];
   double GDDFull[NVEG], LAIMax[NVEG], LAIMin[NVEG], min_res_bioCO2[NVEG], resp_a[NVEG], resp_b[NVEG], SDDFull[NVEG], theta_bioCO2[NVEG];
   double SnowProf_24hr[48];     /* (0:23,2)  -> [hr + 24*col] */
   double LAIPower[12];          /* (4,3)     -> [i + 4*iv]    */
   double OHM_coef[(NSURF + 1) * 4 * 3]; /* (8,4,3) -> [is + 8*(j + 4*k)] */
   double WaterDist[(NSURF + 1) * (NSURF - 1)]; /* (8,6) -> [i + 8*j] */
   double AHProf_24hr[48], HumActivity_24hr[48], PopProf_24hr[48], TraffProf_24hr[48];
   double WUProfA_24hr[48], WUProfM_24hr[48];
   SoEstmParams estm;
   SoCblParams cbl;
} SoParams;

#define PROF24(p, hr, col) ((p)[(hr) + 24 * (col)])       /* col 0 = weekday, 1 = weekend */
#define LAIPOW(p, i, iv) ((p)[(i) + 4 * (iv)])             /* i 0..3, iv 0..2 */
#define OHMCOEF(p, is, j, k) ((p)[(is) + 8 * ((j) + 4 * (k))]) /* is 0..7, j 0..3, k 0..2 */
#define WDIST(p, i, j) ((p)[(i) + 8 * (j)])                /* i 0..7 (receiver), j 0..5 (donor) */
#define SDP(p, i, is) ((p)[(i) + 6 * (is)])                /* StoreDrainPrm(6,nsurf): i 0..5 */

/* Per-step forcing + clock (suews_ctrl_driver.f95:3177-3201) */
typedef struct {
   int iy, id, it, imin, isec;
   double avkdn, avRh, avU1, fcld_obs, LAI_obs, ldown_obs, Precip, Press_hPa;
   double qn1_obs, qs_obs, qf_obs, snowFrac_obs, Temp_C, wu_m3, xsmd;
   double Ts5mindata_ir[13];   /* ESTM surface-temperature forcing row (driver.f95:276); unused unless ESTM runs */
} SoForcing;

/* Evolving per-cell state: the 29 INTENT(INOUT) arguments
 * (suews_ctrl_driver.f95:290-333) + caller-side clock + hidden Lat_vap SAVEs
 * (suews_util_meteo.f95:290). */
typedef struct {
   double qn1_av, dqndt, qn1_s_av, dqnsdt;
   double SnowfallCum, SnowAlb;
   double IceFrac[NSURF], SnowWater[NSURF], SnowDens[NSURF], SnowFrac[NSURF], SnowPack[NSURF];
   double soilstore_id[NSURF], state_id[NSURF];
   double StoreDrainPrm[6 * NSURF];
   double alb[NSURF];
   double GDD_id[NVEG], SDD_id[NVEG], LAI_id[NVEG];
   double Tmin_id, Tmax_id, lenDay_id;
   double DecidCap_id, albDecTr_id, albEveTr_id, albGrass_id, porosity_id;
   double HDD_id[12];
   double WUDay_id[9];
   double Tair_av;
   int dt_since_start, tstep_prev;
   int lv_switch1, lv_switch2; /* Lat_vap implicit SAVE */
   /* cal_Main's local alb1 is only assigned when some surface carries snow (narp.f95:366-370); its
    * stack slot keeps the previous call's value otherwise (golden snow file: 0 on step 1, 0.18 after) */
   double alb1_hidden;
   SoEstmState estm;
   SoAnohmState anohm;
   SoCblState cbl;
} SoState;

typedef struct {
   double datetimeLine[5];
   double SUEWS[NCOL_SUEWS];
   double Snow[NCOL_SNOW];
   double ESTM[NCOL_ESTM];
   double RSL[NCOL_RSL];
   double SOLWEIG[NCOL_SOLWEIG];
   double DailyState[NCOL_DAILY];
} SoOutLines;

/* Error/warning sink replacing ErrorHint (suews_ctrl_error.f95:1). */
typedef struct {
   int fatal_code;       /* first fatal code seen, 0 if none */
   char fatal_msg[160];
   double fatal_v1, fatal_v2;
   int fatal_iv;
   long warn_count[128]; /* per warning code */
   /* hidden Lat_vap SAVE state lives in SoState; a pointer to the active one */
   int *lv_switch1, *lv_switch2;
   int n_iter_last; /* diagnostics: Tsurf iterations used by the last step */
   /* what the last SUEWS_cal_Main left in MODULE moist (driver.f95: use moist, only: avcp, avdens, lv_J_kg) and its
    * friction velocity: the inputs of BLUEWS (suews_ctrl_calculations.f95:194-218) */
   double last_avcp, last_avdens, last_lv_J_kg;
   double last_psih;   /* MODULE sues_data's psih: AerodynamicResistance leaves it there (resist.f95:40,78), BLUEWS reads it */
} SoCtx;

void so_ctx_init(SoCtx *cx);
void so_error_hint(SoCtx *cx, int code, const char *msg, double v1, double v2, int iv);

/* ---- suews_util_time.f95 ---- */
void so_day2month(int b, int *mb, int *md, int *seas, int year, double latitude);
void so_day_of_week(int date, int month, int year, int *dow);
void so_dectime_to_timevec(double dectime, int *hours, int *mins, double *secs);
void so_cal_dectime(int id, int it, int imin, int isec, double *dectime);
void so_cal_tstep(int tstep, int *nsh, double *nsh_real, double *tstep_real);
void so_cal_weekday(int iy, int id, double lat, int dayofWeek_id[3]);
void so_cal_DLS(int id, int startDLS, int endDLS, int *DLS);

/* ---- suews_util_meteo.f95 ---- */
double so_sat_vap_press(double TK, double P);
double so_sat_vap_press_x(SoCtx *cx, double *Temp_C, double Press_hPa, int from, double dectime);
double so_sat_vap_pressIce(SoCtx *cx, double *Temp_C, double Press_hPa, int from, double dectime);
double so_spec_hum_def(double vpd_hPa, double press_hPa);
double so_spec_heat_beer(SoCtx *cx, double Temp_C, double rh, double rho_v, double rho_d);
double so_Lat_vap(SoCtx *cx, double Temp_C, double Ea_hPa, double Press_hPa, double cp, double dectime);
double so_Lat_vapSublim(double Temp_C, double Ea_hPa, double Press_hPa, double cp);
double so_psyc_const(SoCtx *cx, double cp, double Press_hPa, double lv_J_kg);
double so_slope_svp(double temp_C);
double so_slopeIce_svp(double temp_C);
double so_qsatf(SoCtx *cx, double T, double PMB);
double so_RH2qa(double RH_dec, double pres_hPa, double Ta_degC);
double so_qa2RH(double qa_gkg, double pres_hPa, double Ta_degC);

/* ---- suews_phys_atmmoiststab.f95 ---- */
void so_cal_AtmMoist(SoCtx *cx, double *Temp_C, double Press_hPa, double avRh, double dectime,
                     double *lv_J_kg, double *lvS_J_kg, double *es_hPa, double *Ea_hPa, double *VPd_hpa,
                     double *VPD_Pa, double *dq, double *dens_dry, double *avcp, double *air_dens);
void so_cal_Stab(SoCtx *cx, int StabilityMethod, double zzd, double z0m, double zdm, double avU1, double Temp_C,
                 double QH_init, double avdens, double avcp, double *L_MOD, double *TStar, double *UStar, double *zL);
double so_stab_psi_mom(int StabilityMethod, double *zL);
double so_stab_phi_mom(int StabilityMethod, double zL);
double so_stab_psi_heat(int StabilityMethod, double zL);
double so_stab_phi_heat(int StabilityMethod, double zL);

/* ---- suews_phys_resist.f95 ---- */
void so_AerodynamicResistance(SoCtx *cx, double ZZD, double z0m, double AVU1, double L_mod, double UStar,
                              double VegFraction, int AerodynamicResistanceMethod, int StabilityMethod,
                              int RoughLenHeatMethod, double *RA);
void so_SurfaceResistance(SoCtx *cx, int id, int it, int SMDMethod, const double SnowFrac[NSURF],
                          const double sfr[NSURF], double avkdn, double Temp_C, double dq, double xsmd, double vsmd,
                          const double MaxConductance[3], const double LAIMax[3], const double LAI_id[3],
                          int gsModel, double Kmax, double G1, double G2, double G3, double G4, double G5, double G6,
                          double TH, double TL, double S1, double S2, double *gfunc, double *gsc, double *ResistSurf);
void so_BoundaryLayerResistance(double zzd, double z0m, double avU1, double *UStar, double *rb);
void so_cal_RoughnessParameters(SoCtx *cx, int RoughLenMomMethod, const double sfr[NSURF], double bldgH,
                                double EveTreeH, double DecTreeH, double porosity_id, double FAIBldg,
                                double FAIEveTree, double FAIDecTree, double z0m_in, double zdm_in, double Z,
                                double *FAI, double *Zh, double *z0m, double *zdm, double *ZZD);
double so_cal_z0V(int RoughLenHeatMethod, double z0m, double VegFraction, double UStar);

/* ---- suews_phys_narp.f95 ---- */
void so_RadMethod(SoCtx *cx, int NetRadiationMethod, int snowUse, int *NetRadiationMethod_use, int *AlbedoChoice,
                  int *ldown_option);
void so_NARP_cal_SunPosition(double year, double idectime, double UTC, double locationlatitude,
                             double locationlongitude, double locationaltitude, double *sunazimuth,
                             double *sunzenith);
void so_NARP(SoCtx *cx, const double sfr[NSURF], const double SnowFrac[NSURF], const double alb[NSURF],
             const double emis[NSURF], const double IceFrac[NSURF], double NARP_TRANS_SITE, double NARP_EMIS_SNOW,
             double DTIME, double ZENITH_deg, double tsurf_0, double kdown, double Temp_C, double RH, double Press_hPa,
             double qn1_obs, double SnowAlb, int AlbedoChoice, int ldown_option, int NetRadiationMethod_use,
             int DiagQN, double *qstarall, double *qstar_sf, double *qstar_s, double *kclear, double *kupall,
             double *ldown, double *lupall, double *fcld, double *tsurfall, double qn1_ind_snow[NSURF],
             double kup_ind_snow[NSURF], double Tsurf_ind_snow[NSURF], double Tsurf_ind[NSURF], double *albedo_snowfree,
             double *albedo_snow);

/* ---- suews_phys_ohm.f95 ---- */
void so_OHM(SoCtx *cx, double qn1, double qn1_av_prev, double dqndt_prev, double *qn1_av_next, double *dqndt_next,
            double qn1_S, double qn1_s_av_prev, double dqnsdt_prev, double *qn1_s_av_next, double *dqnsdt_next,
            int tstep, int dt_since_start, const double sfr[NSURF], double Tair_mav_5d, const double *OHM_coef,
            const double OHM_threshSW[NSURF + 1], const double OHM_threshWD[NSURF + 1],
            const double soilstore_id[NSURF], const double SoilStoreCap[NSURF], const double state_id[NSURF],
            int SnowUse, const double SnowFrac[NSURF], int DiagQS, double *a1, double *a2, double *a3, double *qs,
            double deltaQi[NSURF]);

void so_OHM_dqndt_cal_X(int dt, int dt_since_start, double qn1_av_prev, double qn1, double dqndt_prev,
                        double *qn1_av_next, double *dqndt_next);
void so_OHM_QS_cal(double qn1, double dqndt, double a1, double a2, double a3, double *qs);

/* ---- suews_util_minpack.f95 (the two drivers AnOHM calls) ---- */
enum { SO_MP_MAXN = 4, SO_MP_MAXM = 96 };
typedef void (*so_lm_fcn)(int m, int n, const double *x, double *fvec, void *ctx);
typedef void (*so_hy_fcn)(int n, const double *x, double *fvec, void *ctx);
int so_lmdif1(so_lm_fcn fcn, void *ctx, int m, int n, double *x, double *fvec, double tol);
int so_hybrd1(so_hy_fcn fcn, void *ctx, int n, double *x, double *fvec, double tol);

/* ---- suews_phys_anohm.f95 ---- */
/* MetForcingData_grid of the cell being stepped: [nrows x 24] column-major (the host's whole forcing block,
 * suews_ctrl_translate.f95:1216), or NULL.  so_cal_multitsteps / so_run_cells point it at their own block unless the
 * caller set one. */
extern const double *so_met_grid;
extern int so_met_grid_nrows;
void so_AnOHM_coef_land_cal(double ASd, double mSd, double ATa, double mTa, double tau, double mWS, double mWF,
                            double mAH, double xalb, double xemis, double xcp, double xk, double xch, double xBo,
                            double *xa1, double *xa2, double *xa3, double *ATs, double *mTs, double *gamma);
void so_AnOHM_coef_water_cal(double ASd, double mSd, double ATa, double mTa, double tau, double mWS, double mWF,
                             double mAH, double xalb, double xemis, double xcp, double xk, double xch, double xBo,
                             double xeta, double xmu, double *xa1, double *xa2, double *xa3, double *ATs, double *mTs,
                             double *gamma);
void so_AnOHM_ShapeFit(int m, const double *tHr, const double *obs, double *amp, double *mean, double *tpeak);
void so_AnOHM(SoCtx *cx, SoAnohmState *an, int tstep, int dt_since_start, double qn1, double qn1_av_prev,
              double dqndt_prev, double qf, const double *MetForcingData_grid, int nrows,
              const double moist_surf[NSURF], const double alb[NSURF], const double emis[NSURF],
              const double cpAnOHM[NSURF], const double kkAnOHM[NSURF], const double chAnOHM[NSURF],
              const double sfr[NSURF], int EmissionsMethod, int id, int Gridiv, double *qn1_av_next,
              double *dqndt_next, double *a1, double *a2, double *a3, double *qs, double deltaQi[NSURF]);

/* ---- suews_phys_bluews.f95 ---- */
extern const double *so_cbl_ini;   /* IniCBLdata [nline x 8] row-major, or NULL */
extern int so_cbl_ini_nline;
extern double *so_blk_BL;          /* optional [len_sim x 22] column-major block filled by so_cal_multitsteps */
void so_CBL(SoCtx *cx, const SoCblParams *p, SoCblState *s, const SoForcing *f, double dectime, int tstep, double qh_use,
            double qe_use, double UStar, double avdens, double lv_J_kg, double avcp, double line22[22]);

/* ---- suews_phys_lumps.f95 ---- */
void so_LUMPS_cal_QHQE(SoCtx *cx, int veg_type, int snowUse, double qn1, double qf, double qs, double Qm,
                       double Temp_C, double Veg_Fr, double avcp, double Press_hPa, double lv_J_kg,
                       double tstep_real, double DRAINRT, double nsh_real, double Precip, double RainMaxRes,
                       double RAINCOVER, const double sfr[NSURF], const double LAI_id_prev[NVEG],
                       const double LAImax[NVEG], const double LAImin[NVEG], double *QH_LUMPS, double *QE_LUMPS,
                       double *psyc_hPa, double *s_hPa, double *sIce_hpa, double *TempVeg, double *VegPhenLumps);

/* ---- suews_phys_evap.f95 ---- */
void so_cal_evap(int EvapMethod, double state_is, double WetThresh_is, double capStore, double vpd_hPa, double avdens,
                 double avcp, double qn_e, double s_hPa, double psyc_hPa, double ResistSurf, double RA, double rb,
                 double tlv, double *rss, double *ev, double *qe);

/* ---- suews_ctrl_input.f95:743-826 (profile lookups used by the physics) ---- */
double so_get_Prof_SpecTime_inst(int Hour, int Min, int Sec, const double Prof_24h[24]);
double so_get_Prof_SpecTime_mean(int Hour, int Min, int Sec, const double Prof_24h[24]);
double so_get_Prof_SpecTime_sum(int Hour, int Min, int Sec, const double Prof_24h[24], int dt);

/* ---- suews_phys_anemsn.f95 / biogenco2 ---- */
void so_AnthropogenicEmissions(SoCtx *cx, double CO2PointSource, int EmissionsMethod, int it, int imin, int DLS,
                               const int DayofWeek_id[3], double EF_umolCO2perJ, const double FcEF_v_kgkm[2],
                               double EnEF_v_Jkm, double TrafficUnits, double FrFossilFuel_Heat,
                               double FrFossilFuel_NonHeat, double MinFCMetab, double MaxFCMetab, double MinQFMetab,
                               double MaxQFMetab, const double PopDensDaytime[2], double PopDensNighttime,
                               double Temp_C, const double HDD_id[12], const double Qf_A[2], const double Qf_B[2],
                               const double Qf_C[2], const double AH_MIN[2], const double AH_SLOPE_Heating[2],
                               const double AH_SLOPE_Cooling[2], const double BaseT_Heating[2],
                               const double BaseT_Cooling[2], const double TrafficRate[2], const double QF0_BEU[2],
                               double *QF_SAHP, double *Fc_anthro, double *Fc_metab, double *Fc_traff,
                               double *Fc_build, double *Fc_point, const double AHProf_24hr[48],
                               const double HumActivity_24hr[48], const double TraffProf_24hr[48],
                               const double PopProf_24hr[48], double SurfaceArea);
void so_CO2_biogen(SoCtx *cx, const double alpha_bioCO2[NVEG], const double alpha_enh_bioCO2[NVEG], double avkdn,
                   const double beta_bioCO2[NVEG], const double beta_enh_bioCO2[NVEG], double dectime,
                   int EmissionsMethod, double gfunc, double gfunc2, int gsmodel, int id, int it,
                   const double LAI_id[NVEG], const double LAIMin[NVEG], const double LAIMax[NVEG],
                   const double min_res_bioCO2[NVEG], const double resp_a[NVEG], const double resp_b[NVEG],
                   const double sfr[NSURF], const double SnowFrac[NSURF], double t2, double Temp_C,
                   const double theta_bioCO2[NVEG], double *Fc_biogen, double *Fc_photo, double *Fc_respi);

/* ---- suews_phys_waterdist.f95 ---- */
void so_drainage(SoCtx *cx, int is, double state_is, double StorCap, double DrainEq, double DrainCoef1,
                 double DrainCoef2, double nsh_real, double *drain_is);
void so_cal_water_storage(SoCtx *cx, int is, const double sfr[NSURF], double PipeCapacity, double RunoffToWater,
                          double pin, const double WU_nsurf[NSURF], const double drain[NSURF],
                          const double AddWater[NSURF], double addImpervious, double nsh_real,
                          const double stateOld[NSURF], const double AddWaterRunoff[NSURF], double PervFraction,
                          double addVeg, const double SoilStoreCap[NSURF], double addWaterBody, double FlowChange,
                          const double StateLimit[NSURF], double *runoffAGimpervious, double *surplusWaterBody,
                          double *runoffAGveg, double *runoffPipes, double *ev, double soilstore_id[NSURF],
                          double SurplusEvap[NSURF], double *runoffWaterBody, double *p_mm, double chang[NSURF],
                          double runoff[NSURF], double state_id[NSURF]);
void so_updateFlood(int is, const double runoff[NSURF], const double sfr[NSURF], double PipeCapacity,
                    double RunoffToWater, double *runoffAGimpervious, double *surplusWaterBody, double *runoffAGveg,
                    double *runoffPipes);
void so_ReDistributeWater(int snowUse, const double *WaterDist, const double sfr[NSURF], const double Drain[NSURF],
                          double AddWaterRunoff[NSURF], double AddWater[NSURF]);
void so_update_SoilMoist(double NonWaterFraction, const double SoilStoreCap[NSURF], const double sfr[NSURF],
                         const double soilstore_id[NSURF], double *SoilMoistCap, double *SoilState, double *vsmd,
                         double *smd);
void so_cal_SoilState(SoCtx *cx, int SMDMethod, double xsmd, double NonWaterFraction, double SoilMoistCap,
                      const double SoilStoreCap[NSURF], double surf_chang_per_tstep, const double soilstore_id[NSURF],
                      const double soilstoreOld[NSURF], const double sfr[NSURF], double *smd, double smd_nsurf[NSURF],
                      double *tot_chang_per_tstep, double *SoilState);
void so_cal_HorizontalSoilWater(const double sfr[NSURF], const double SoilStoreCap[NSURF],
                                const double SoilDepth[NSURF], const double SatHydraulicConduct[NSURF],
                                double SurfaceArea, double NonWaterFraction, double tstep_real,
                                double soilstore_id[NSURF], double runoffSoil[NSURF], double *runoffSoil_per_tstep);
void so_cal_WaterUse(SoCtx *cx, double nsh_real, double wu_m3, double SurfaceArea, const double sfr[NSURF],
                     double IrrFracPaved, double IrrFracBldgs, double IrrFracEveTr, double IrrFracDecTr,
                     double IrrFracGrass, double IrrFracBSoil, double IrrFracWater, const int DayofWeek_id[3],
                     const double WUProfA_24hr[48], const double WUProfM_24hr[48], double InternalWaterUse_h,
                     const double HDD_id[12], const double WUDay_id[9], int WaterUseMethod, int NSH, int it,
                     int imin, int DLS, double wu_nsurf[NSURF], double *wu_int, double *wu_ext);

/* ---- suews_phys_snow.f95 ---- */
double so_update_snow_albedo(int tstep, const double SnowPack_prev[NSURF], double SnowAlb_prev, double Temp_C,
                             double tau_a, double tau_f, double SnowAlbMax, double SnowAlbMin);
void so_update_snow_dens(int tstep, const double SnowFrac[NSURF], const double SnowDens_prev[NSURF], double tau_r,
                         double SnowDensMax, double SnowDensMin, double SnowDens_next[NSURF]);
void so_Snow_cal_MeltHeat(SoCtx *cx, int snowUse, int tstep, double tau_r, double SnowDensMax, double lvS_J_kg,
                          double lv_J_kg, double tstep_real, double RadMeltFact, double TempMeltFact,
                          double SnowAlbMax, double SnowDensMin, double Temp_C, double Precip, double PrecipLimit,
                          double PrecipLimitAlb, double nsh_real, const double sfr[NSURF],
                          const double Tsurf_ind[NSURF], const double Tsurf_ind_snow[NSURF],
                          const double state_id[NSURF], const double qn1_ind_snow[NSURF],
                          const double kup_ind_snow[NSURF], const double SnowWater[NSURF],
                          const double deltaQi[NSURF], double alb1, const double SnowPack_in[NSURF],
                          const double SnowFrac_in[NSURF], double SnowAlb_in, const double SnowDens_in[NSURF],
                          double SnowfallCum_in, double SnowPack_out[NSURF], double SnowFrac_out[NSURF],
                          double *SnowAlb_out, double SnowDens_out[NSURF], double *SnowfallCum_out, double *mwh,
                          double *Qm, double *QmFreez, double *QmRain, double *veg_fr, int snowCalcSwitch[NSURF],
                          double Qm_melt[NSURF], double Qm_freezState[NSURF], double Qm_rain[NSURF],
                          double FreezMelt[NSURF], double FreezState[NSURF], double FreezStateVol[NSURF],
                          double rainOnSnow[NSURF], double SnowDepth[NSURF], double mw_ind[NSURF],
                          double dataOutLineSnow[NCOL_SNOW]);
void so_SnowCalc(SoCtx *cx, int tstep, int imin, int it, double dectime, int is, int EvapMethod, double CRWmin,
                 double CRWmax, double nsh_real, double lvS_J_kg, double avdens, double avRh, double Press_hPa,
                 double Temp_C, double RAsnow, double psyc_hPa, double avcp, double sIce_hPa, double PervFraction,
                 double vegfraction, double addimpervious, double vpd_hPa, double qn_e, double s_hPa,
                 double ResistSurf, double RA, double rb, double tlv, double snowdensmin,
                 const double SnowProf_24hr[48], double precip, double PipeCapacity, double RunoffToWater,
                 double addVeg, double SnowLimPaved, double SnowLimBldg, double FlowChange,
                 const double drain[NSURF], const double WetThresh[NSURF], const double stateOld[NSURF],
                 const double mw_ind[NSURF], const double SoilStoreCap[NSURF], const double rainonsnow[NSURF],
                 const double freezmelt[NSURF], const double freezstate[NSURF], const double freezstatevol[NSURF],
                 const double Qm_Melt[NSURF], const double Qm_rain[NSURF], const double Tsurf_ind[NSURF],
                 const double sfr[NSURF], const int DayofWeek_id[3], const double *StoreDrainPrm,
                 const double SnowPackLimit[NSURF], const double AddWater[NSURF],
                 const double addwaterrunoff[NSURF], double soilstore_id[NSURF], double SnowPack[NSURF],
                 double SurplusEvap[NSURF], double SnowFrac[NSURF], double SnowWater[NSURF], double iceFrac[NSURF],
                 double SnowDens[NSURF], double *runoffAGimpervious, double *runoffAGveg, double *surplusWaterBody,
                 double rss_nsurf[NSURF], double runoffSnow[NSURF], double runoff[NSURF], double runoffSoil[NSURF],
                 double chang[NSURF], double changSnow[NSURF], double SnowToSurf[NSURF], double state_id[NSURF],
                 double ev_snow[NSURF],
                 double SnowDepth[NSURF], double SnowRemoval[2], double *swe, double *ev, double *chSnow_tot,
                 double *ev_tot, double *qe_tot, double *runoff_tot, double *surf_chang_tot, double *runoffPipes,
                 double *mwstore, double *runoffwaterbody);

/* ---- suews_phys_dailystate.f95 ---- */
void so_cal_DailyState(SoCtx *cx, const SoParams *p, const SoForcing *f, int tstep_prev, int dt_since_start,
                       const int DayofWeek_id[3], double Tmin_id_prev, double Tmax_id_prev, double lenDay_id_prev,
                       double nsh_real, double DecidCap_id_prev, const double *StoreDrainPrm_prev,
                       const double LAI_id_prev[NVEG], const double GDD_id_prev[NVEG], const double SDD_id_prev[NVEG],
                       double albDecTr_id_prev, double albEveTr_id_prev, double albGrass_id_prev,
                       double porosity_id_prev, const double HDD_id_prev[12], const double state_id[NSURF],
                       const double soilstore_id[NSURF], double HDD_id_next[12], double *Tmin_id_next,
                       double *Tmax_id_next, double *lenDay_id_next, double *albDecTr_id_next,
                       double *albEveTr_id_next, double *albGrass_id_next, double *porosity_id_next,
                       double *DecidCap_id_next, double *StoreDrainPrm_next, double LAI_id_next[NVEG],
                       double GDD_id_next[NVEG], double SDD_id_next[NVEG], double *deltaLAI, double WUDay_id[9]);
void so_update_DailyStateLine(int it, int imin, double nsh_real, const double GDD_id[NVEG], const double HDD_id[12],
                              const double LAI_id[NVEG], const double SDD_id[NVEG], double Tmin_id, double Tmax_id,
                              double lenday_id, double DecidCap_id, double albDecTr_id, double albEveTr_id,
                              double albGrass_id, double porosity_id, const double WUDay_id[9], double deltaLAI,
                              double VegPhenLumps, double SnowAlb, const double SnowDens[NSURF], double a1,
                              double a2, double a3, double DailyStateLine[NCOL_DAILY]);

/* ---- suews_phys_rslprof.f95 ---- */
void so_RSLProfile(SoCtx *cx, double Zh, double z0m, double zdm, double L_MOD, const double sfr[NSURF], double FAI,
                   int StabilityMethod, double avcp, double lv_J_kg, double avdens, double avU1, double Temp_C,
                   double avRH, double Press_hPa, double zMeas, double qh, double qe, double *T2_C, double *q2_gkg,
                   double *U10_ms, double *RH2, double dataoutLineRSL[NCOL_RSL]);

/* ---- suews_ctrl_driver.f95 ---- */
double so_cal_tair_av(double tair_av_prev, int dt_since_start, int tstep, double temp_c);
double so_cal_tsfc(double qh, double avdens, double avcp, double RA, double temp_c);
double so_set_nan(double x);
/* suews_phys_solweig.f95 (point version) */
void so_SOLWEIG_cal_main(int id, int it, double dectime, double lamdaP, double lamdaF, double avkdn, double ldown,
                         double Temp_C, double avrh, double Press_hPa, double Tg, double lat, double zenith_deg,
                         double azimuth, double scale, double alb_ground, double alb_bldg, double emis_ground,
                         double emis_wall, double heightgravity, double out[NCOL_SOLWEIG]);

/* ---- suews_phys_estm.f95 ---- */
double so_heatcapacity_air(double TK, double RH, double P);
double so_NewtonPolynomial(double x0, const double Pcoeff[5], double conv, int maxiter, double *converged_save);
void so_heatcond1d(double *T, double *Qs, const double *dx, double dt, const double *k, const double *rhocp, int n,
                   const double bc[2], const int bctype[2]);
void so_ESTM_translate_start(const SoParams *p, const double alb[NSURF], const double Ts5mindata_row1[13],
                             SoEstmState *e);
void so_ESTM(SoCtx *cx, const SoParams *p, SoEstmState *e, int tstep, double avkdn, double avu1, double temp_c,
             double zenith_deg, double avrh, double press_hpa, double ldown, double bldgh,
             const double Ts5mindata_ir[13], double Tair_av, double dataOutLineESTM[NCOL_ESTM], double *QS);
/* optional ESTM forcing block for so_cal_multitsteps / so_run_cells: [len_sim x 13] column-major, or NULL */
extern const double *so_ts5_block;

/* One cell, one timestep == SUEWS_cal_Main.  Mutates *st (including
 * nothing of dt_since_start: the caller advances it, as in the reference). */
void so_cal_main(SoCtx *cx, const SoParams *p, const SoForcing *f, SoState *st, SoOutLines *out);

/* One cell, len_sim steps == SUEWS_cal_multitsteps (suews_ctrl_driver.f95:2814).
 * met: column-major [len_sim x 24].  Output blocks are column-major
 * [len_sim x (ncol+5)], may be NULL to skip.  Advances st->dt_since_start. */
void so_cal_multitsteps(SoCtx *cx, const SoParams *p, SoState *st, const double *met, int len_sim,
                        double *blkSUEWS, double *blkSnow, double *blkESTM, double *blkRSL, double *blkSOL,
                        double *blkDaily);

#ifdef __cplusplus
}
#endif
#endif
