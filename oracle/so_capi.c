/* ORACLE (test infrastructure only -- see suews_oracle.h).
 * Name-keyed C API so tests can fill SoParams/SoState from the reference's own
 * argument names (as they appear in suews_ctrl_driver.f95:45-91 and in the
 * *_state_init.txt dumps written by suews_ctrl_translate.f95:1280-1500). */
#include "suews_oracle.h"
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#ifdef __cplusplus   /* (the counting build compiles these sources as C++, so_count.h) */
extern "C" {
#endif

typedef struct {
   const char *name;
   int is_int;
   size_t off;
   int n;
} SoField;

#define PD(f) {#f, 0, offsetof(SoParams, f), (int)(sizeof(((SoParams *)0)->f) / sizeof(double))}
#define PI(f) {#f, 1, offsetof(SoParams, f), (int)(sizeof(((SoParams *)0)->f) / sizeof(int))}
#define PDE(f) {#f, 0, offsetof(SoParams, estm.f), (int)(sizeof(((SoParams *)0)->estm.f) / sizeof(double))}
#define PIE(f) {#f, 1, offsetof(SoParams, estm.f), (int)(sizeof(((SoParams *)0)->estm.f) / sizeof(int))}
static const SoField PARAM_FIELDS[] = {
   PI(AerodynamicResistanceMethod), PI(BaseTMethod), PI(Diagnose), PI(DiagQN), PI(DiagQS), PI(startDLS), PI(endDLS),
   PI(EmissionsMethod), PI(Gridiv), PI(gsModel), PI(Ie_end), PI(Ie_start), PI(EvapMethod), PI(LAICalcYes),
   PI(NetRadiationMethod), PI(OHMIncQF), PI(RoughLenHeatMethod), PI(RoughLenMomMethod), PI(SMDMethod), PI(snowUse),
   PI(StabilityMethod), PI(StorageHeatMethod), PI(tstep), PI(veg_type), PI(WaterUseMethod), PI(LAIType),
   PD(AlbMax_DecTr), PD(AlbMax_EveTr), PD(AlbMax_Grass), PD(AlbMin_DecTr), PD(AlbMin_EveTr), PD(AlbMin_Grass),
   PD(alt), PD(BaseT_HC), PD(bldgH), PD(CapMax_dec), PD(CapMin_dec), PD(CO2PointSource), PD(CRWmax), PD(CRWmin),
   PD(DecTreeH), PD(DRAINRT), PD(EF_umolCO2perJ), PD(EnEF_v_Jkm), PD(EveTreeH), PD(FAIBldg), PD(FAIDecTree),
   PD(FAIEveTree), PD(Faut), PD(FlowChange), PD(FrFossilFuel_Heat), PD(FrFossilFuel_NonHeat), PD(G1), PD(G2), PD(G3),
   PD(G4), PD(G5), PD(G6), PD(H_maintain), PD(InternalWaterUse_h), PD(IrrFracPaved), PD(IrrFracBldgs),
   PD(IrrFracEveTr), PD(IrrFracDecTr), PD(IrrFracGrass), PD(IrrFracBSoil), PD(IrrFracWater), PD(Kmax), PD(lat),
   PD(lng), PD(MaxFCMetab), PD(MaxQFMetab), PD(MinFCMetab), PD(MinQFMetab), PD(NARP_EMIS_SNOW), PD(NARP_TRANS_SITE),
   PD(PipeCapacity), PD(PopDensNighttime), PD(PorMax_dec), PD(PorMin_dec), PD(PrecipLimit), PD(PrecipLimitAlb),
   PD(RadMeltFact), PD(RAINCOVER), PD(RainMaxRes), PD(RunoffToWater), PD(S1), PD(S2), PD(SnowAlbMax), PD(SnowAlbMin),
   PD(SnowDensMax), PD(SnowDensMin), PD(SnowLimBldg), PD(SnowLimPaved), PD(SurfaceArea), PD(tau_a), PD(tau_f),
   PD(tau_r), PD(TempMeltFact), PD(TH), PD(timezone), PD(TL), PD(TrafficUnits), PD(Z), PD(z0m_in), PD(zdm_in),
   PD(AH_MIN), PD(AH_SLOPE_Cooling), PD(AH_SLOPE_Heating), PD(FcEF_v_kgkm), PD(QF0_BEU), PD(Qf_A), PD(Qf_B), PD(Qf_C),
   PD(PopDensDaytime), PD(BaseT_Cooling), PD(BaseT_Heating), PD(TrafficRate), PD(Ie_a), PD(Ie_m), PD(MaxConductance),
   PD(DayWat), PD(DayWatPer), PD(OHM_threshSW), PD(OHM_threshWD), PD(chAnOHM), PD(cpAnOHM), PD(emis), PD(kkAnOHM),
   PD(SatHydraulicConduct), PD(sfr), PD(SnowPackLimit), PD(SoilDepth), PD(SoilStoreCap), PD(StateLimit),
   PD(WetThresh), PD(alpha_bioCO2), PD(alpha_enh_bioCO2), PD(BaseT), PD(BaseTe), PD(beta_bioCO2),
   PD(beta_enh_bioCO2), PD(GDDFull), PD(LAIMax), PD(LAIMin), PD(min_res_bioCO2), PD(resp_a), PD(resp_b), PD(SDDFull),
   PD(theta_bioCO2), PD(SnowProf_24hr), PD(LAIPower), PD(OHM_coef), PD(WaterDist), PD(AHProf_24hr),
   PD(HumActivity_24hr), PD(PopProf_24hr), PD(TraffProf_24hr), PD(WUProfA_24hr), PD(WUProfM_24hr),
   /* MODULE ESTM_data (suews_phys_estm.f95:1-176) */
   PIE(evolveTibld), PIE(TsurfChoice), PIE(ibldCHmod), PIE(Nibld), PIE(Nwall), PIE(Nroof), PIE(Nground),
   PDE(LBC_soil), PDE(THEAT_ON), PDE(THEAT_OFF), PDE(THEAT_fix), PDE(ivf_iw), PDE(ivf_ir), PDE(ivf_ii), PDE(ivf_if),
   PDE(ivf_ww), PDE(ivf_wr), PDE(ivf_wi), PDE(ivf_wf), PDE(ivf_rw), PDE(ivf_ri), PDE(ivf_rf), PDE(ivf_fw), PDE(ivf_fr),
   PDE(ivf_fi), PDE(zibld), PDE(kibld), PDE(ribld), PDE(zwall), PDE(kwall), PDE(rwall), PDE(zroof), PDE(kroof),
   PDE(rroof), PDE(zground), PDE(kground), PDE(rground), PDE(nroom), PDE(alb_ibld), PDE(em_ibld), PDE(CH_iroof),
   PDE(CH_iwall), PDE(CH_ibld), PDE(AreaWall), PDE(Tfloor),
   /* CBLInput.nml (suews_phys_bluews.f95:300-357) and RunControl's CBLuse */
   {"CBLuse", 1, offsetof(SoParams, cbl.CBLuse), 1}, {"EntrainmentType", 1, offsetof(SoParams, cbl.EntrainmentType), 1},
   {"QH_choice", 1, offsetof(SoParams, cbl.QH_choice), 1}, {"isubs", 1, offsetof(SoParams, cbl.isubs), 1},
   {"CO2_included", 1, offsetof(SoParams, cbl.CO2_included), 1}, {"InitialData_use", 1, offsetof(SoParams, cbl.InitialData_use), 1},
   {"sondeflag", 1, offsetof(SoParams, cbl.sondeflag), 1}, {"wsb", 0, offsetof(SoParams, cbl.wsb), 1},
};
#undef PD
#undef PI

#define SD(f) {#f, 0, offsetof(SoState, f), (int)(sizeof(((SoState *)0)->f) / sizeof(double))}
#define SI(f) {#f, 1, offsetof(SoState, f), (int)(sizeof(((SoState *)0)->f) / sizeof(int))}
#define SDE(f) {#f, 0, offsetof(SoState, estm.f), (int)(sizeof(((SoState *)0)->estm.f) / sizeof(double))}
static const SoField STATE_FIELDS[] = {
   SD(qn1_av), SD(dqndt), SD(qn1_s_av), SD(dqnsdt), SD(SnowfallCum), SD(SnowAlb), SD(IceFrac), SD(SnowWater),
   SD(SnowDens), SD(SnowFrac), SD(SnowPack), SD(soilstore_id), SD(state_id), SD(StoreDrainPrm), SD(alb), SD(GDD_id),
   SD(SDD_id), SD(LAI_id), SD(Tmin_id), SD(Tmax_id), SD(lenDay_id), SD(DecidCap_id), SD(albDecTr_id),
   SD(albEveTr_id), SD(albGrass_id), SD(porosity_id), SD(HDD_id), SD(WUDay_id), SD(Tair_av), SI(dt_since_start),
   SI(tstep_prev), SI(lv_switch1), SI(lv_switch2), SD(alb1_hidden),
   SDE(Tair2), SDE(Tievolve), SDE(lup_ground), SDE(lup_wall), SDE(lup_roof), SDE(T0_ibld), SDE(T0_ground), SDE(T0_wall),
   SDE(T0_roof), SDE(TN_wall), SDE(TN_roof), SDE(Tground), SDE(Twall), SDE(Troof), SDE(Tibld), SDE(Tw_4), SDE(HVAC),
   SDE(Tair2Set), SDE(NewtonConverged),
   /* AnOHM's SAVE'd locals (suews_phys_anohm.f95:79, :215-216) */
   {"anohm_id_save", 0, offsetof(SoState, anohm.id_save), 1}, {"anohm_coef_id_save", 0, offsetof(SoState, anohm.coef_id_save), 1},
   {"anohm_coef_grid_save", 0, offsetof(SoState, anohm.coef_grid_save), 1},
   {"anohm_coeff_grid_day", 0, offsetof(SoState, anohm.coeff_grid_day), 21},
   /* MODULE cbl_MODULE between calls (suews_phys_bluews.f95:24-47) */
   {"cbl_blh_m", 0, offsetof(SoState, cbl.blh_m), 1}, {"cbl_tm_K", 0, offsetof(SoState, cbl.tm_K), 1},
   {"cbl_qm_kgkg", 0, offsetof(SoState, cbl.qm_kgkg), 1}, {"cbl_tp_K", 0, offsetof(SoState, cbl.tp_K), 1},
   {"cbl_qp_kgkg", 0, offsetof(SoState, cbl.qp_kgkg), 1}, {"cbl_cm", 0, offsetof(SoState, cbl.cm), 1},
   {"cbl_gamt_Km", 0, offsetof(SoState, cbl.gamt_Km), 1}, {"cbl_gamq_kgkgm", 0, offsetof(SoState, cbl.gamq_kgkgm), 1},
};
#undef SD
#undef SI

static const SoField *find(const SoField *tab, int ntab, const char *name)
{
   int i;
   for (i = 0; i < ntab; i++)
      if (strcasecmp(tab[i].name, name) == 0) return &tab[i];
   return NULL;
}

static int set_field(void *base, const SoField *fd, const double *vals, int n)
{
   int i;
   if (!fd) return -1;
   if (n != fd->n) return -2;
   if (fd->is_int) {
      int *d = (int *)((char *)base + fd->off);
      for (i = 0; i < n; i++) d[i] = (int)vals[i];
   } else {
      double *d = (double *)((char *)base + fd->off);
      for (i = 0; i < n; i++) d[i] = vals[i];
   }
   return 0;
}

static int get_field(const void *base, const SoField *fd, double *vals, int n)
{
   int i;
   if (!fd) return -1;
   if (n != fd->n) return -2;
   if (fd->is_int) {
      const int *d = (const int *)((const char *)base + fd->off);
      for (i = 0; i < n; i++) vals[i] = d[i];
   } else {
      const double *d = (const double *)((const char *)base + fd->off);
      for (i = 0; i < n; i++) vals[i] = d[i];
   }
   return 0;
}

#define NPF ((int)(sizeof(PARAM_FIELDS) / sizeof(PARAM_FIELDS[0])))
#define NSF ((int)(sizeof(STATE_FIELDS) / sizeof(STATE_FIELDS[0])))

SoParams *so_params_new(void) { return (SoParams *)calloc(1, sizeof(SoParams)); }
SoState *so_state_new(void)
{
   SoState *s = (SoState *)calloc(1, sizeof(SoState));
   int k;
   if (s)
      for (k = 0; k < 21; k++) s->anohm.coeff_grid_day[k] = -999.;   /* coeff_grid_day(7,3) = -999. (anohm.f95:216) */
   return s;
}
void so_free(void *p) { free(p); }
void so_state_copy(SoState *dst, const SoState *src) { *dst = *src; }

int so_params_nfields(void) { return NPF; }
int so_state_nfields(void) { return NSF; }
const char *so_params_field(int i, int *is_int, int *n)
{
   if (i < 0 || i >= NPF) return NULL;
   *is_int = PARAM_FIELDS[i].is_int;
   *n = PARAM_FIELDS[i].n;
   return PARAM_FIELDS[i].name;
}
const char *so_state_field(int i, int *is_int, int *n)
{
   if (i < 0 || i >= NSF) return NULL;
   *is_int = STATE_FIELDS[i].is_int;
   *n = STATE_FIELDS[i].n;
   return STATE_FIELDS[i].name;
}
int so_params_set(SoParams *p, const char *name, const double *vals, int n)
{
   return set_field(p, find(PARAM_FIELDS, NPF, name), vals, n);
}
int so_params_get(const SoParams *p, const char *name, double *vals, int n)
{
   return get_field(p, find(PARAM_FIELDS, NPF, name), vals, n);
}
int so_state_set(SoState *s, const char *name, const double *vals, int n)
{
   return set_field(s, find(STATE_FIELDS, NSF, name), vals, n);
}
int so_state_get(const SoState *s, const char *name, double *vals, int n)
{
   return get_field(s, find(STATE_FIELDS, NSF, name), vals, n);
}

/* Run len_sim steps for one cell.  Returns the first fatal ErrorHint code (0 = ok). */
int so_run(const SoParams *p, SoState *st, const double *met, int len_sim, double *blkSUEWS, double *blkSnow,
           double *blkESTM, double *blkRSL, double *blkSOL, double *blkDaily, char *msg, int msglen,
           long *warn_count128, int *n_iter_last)
{
   SoCtx cx;
   int i;
   so_ctx_init(&cx);
   so_cal_multitsteps(&cx, p, st, met, len_sim, blkSUEWS, blkSnow, blkESTM, blkRSL, blkSOL, blkDaily);
   if (msg && msglen > 0) {
      strncpy(msg, cx.fatal_msg, (size_t)msglen - 1);
      msg[msglen - 1] = 0;
   }
   if (warn_count128)
      for (i = 0; i < 128; i++) warn_count128[i] = cx.warn_count[i];
   if (n_iter_last) *n_iter_last = cx.n_iter_last;
   return cx.fatal_code;
}

/* ESTM forcing block for the next so_run / so_run_cells: [len_sim x 13] column-major (iy id it imin Tiair Tsurf
 * Troof Troad Twall Twall_n Twall_e Twall_s Twall_w), or NULL.  The pointer is borrowed. */
void so_set_ts5_block(const double *ts5) { so_ts5_block = ts5; }

/* BLUEWS: the initial-data table (IniCBLdata, [nline x 8] row-major; borrowed) and an output block for so_run */
void so_set_cbl_initial(const double *ini, int nline) { so_cbl_ini = ini; so_cbl_ini_nline = ini ? nline : 0; }
void so_set_bl_block(double *blk) { so_blk_BL = blk; }

/* MetForcingData_grid for AnOHM: [nrows x 24] column-major, or NULL (then so_run uses its own block).  Borrowed. */
void so_set_met_grid(const double *met, int nrows) { so_met_grid = met; so_met_grid_nrows = met ? nrows : 0; }

/* The day's AnOHM coefficients of the seven surfaces as a fresh process would form them on its first call of day `id`
 * (tests of the device-side fit): coeff21 = coeff_grid_day(7,3) column-major. */
void so_anohm_coefs(const double *met, int nrows, int id, int EmissionsMethod, double qf, const double *moist,
                    const double *alb, const double *emis, const double *cp, const double *kk, const double *ch,
                    double *coeff21)
{
   SoCtx cx;
   SoAnohmState an;
   double sfr[NSURF] = {0, 0, 0, 0, 0, 0, 0}, dq[NSURF], qa, dqn, a1, a2, a3, qs;
   int k;
   so_ctx_init(&cx);
   an.id_save = 0; an.coef_id_save = 0; an.coef_grid_save = 0;
   for (k = 0; k < 21; k++) an.coeff_grid_day[k] = -999.;
   so_AnOHM(&cx, &an, 300, 0, 0.0, 0.0, 0.0, qf, met, nrows, moist, alb, emis, cp, kk, ch, sfr, EmissionsMethod, id, 1,
            &qa, &dqn, &a1, &a2, &a3, &qs, dq);
   for (k = 0; k < 21; k++) coeff21[k] = an.coeff_grid_day[k];
}

/* First-call initialisation of the ESTM element temperatures (ESTM_translate, ESTMStart == 1) */
void so_estm_start(const SoParams *p, SoState *st, const double *ts_row1_13)
{
   so_ESTM_translate_start(p, st->alb, ts_row1_13, &st->estm);
}

/* Many independent cells, each with its own params/state; shared or per-cell forcing.
 * P0-style reduction for the CPU baseline: per-cell sums of the SUEWS line over the run.
 * met_stride = 0 -> all cells share one forcing block. */
int so_run_cells(int ncell, const SoParams *params, SoState *states, const double *met, long met_stride, int len_sim,
                 double *sum_suews /* [ncell x NCOL_SUEWS] or NULL */)
{
   int c, ir, k, rc = 0;
   for (c = 0; c < ncell; c++) {
      SoCtx cx;
      SoForcing f;
      SoOutLines out;
      const double *m = met + (size_t)c * (size_t)met_stride;
      so_ctx_init(&cx);
      so_met_grid = m; so_met_grid_nrows = len_sim;
      if (sum_suews)
         for (k = 0; k < NCOL_SUEWS; k++) sum_suews[(size_t)c * NCOL_SUEWS + k] = 0;
      for (ir = 0; ir < len_sim; ir++) {
#define MET(col) m[(size_t)ir + (size_t)len_sim * (col)]
         f.iy = (int)MET(0); f.id = (int)MET(1); f.it = (int)MET(2); f.imin = (int)MET(3); f.isec = 0;
         f.qn1_obs = MET(4); f.qs_obs = MET(7); f.qf_obs = MET(8); f.avU1 = MET(9); f.avRh = MET(10);
         f.Temp_C = MET(11); f.Press_hPa = MET(12); f.Precip = MET(13); f.avkdn = MET(14); f.snowFrac_obs = MET(15);
         f.ldown_obs = MET(16); f.fcld_obs = MET(17); f.wu_m3 = MET(18); f.xsmd = MET(19); f.LAI_obs = MET(20);
         for (k = 0; k < 13; k++) f.Ts5mindata_ir[k] = so_ts5_block ? so_ts5_block[(size_t)ir + (size_t)len_sim * k] : -999;
#undef MET
         so_cal_main(&cx, &params[c], &f, &states[c], &out);
         states[c].dt_since_start += params[c].tstep;
         if (sum_suews)
            for (k = 0; k < NCOL_SUEWS; k++) sum_suews[(size_t)c * NCOL_SUEWS + k] += out.SUEWS[k];
         if (cx.fatal_code) { rc = cx.fatal_code; break; }
      }
   }
   so_met_grid = 0; so_met_grid_nrows = 0;
   return rc;
}

size_t so_sizeof_params(void) { return sizeof(SoParams); }
size_t so_sizeof_state(void) { return sizeof(SoState); }

#ifdef SO_OPCOUNT
/* ---- op counter of the counting build (so_count.h, `make count`) ---- */
int so_opcount_on = 0, so_opcount_hoisted = 1;
unsigned long long so_opcount_main[16], so_opcount_diag[16];
unsigned long long *so_opcount_cur = so_opcount_main;
/* mode: 0 = every operation the restated reference algorithm executes, 1 = the hoisted minimal algorithm */
void so_opcount_start(int hoisted)
{
   int k;
   for (k = 0; k < 16; k++) { so_opcount_main[k] = 0; so_opcount_diag[k] = 0; }
   so_opcount_cur = so_opcount_main;
   so_opcount_hoisted = hoisted;
   so_opcount_on = 1;
}
/* out32: 16 counts of the flux/state path, then 16 of the diagnostics (RSL profile, SOLWEIG) */
void so_opcount_stop(unsigned long long *out32)
{
   int k;
   so_opcount_on = 0;
   for (k = 0; k < 16; k++) { out32[k] = so_opcount_main[k]; out32[16 + k] = so_opcount_diag[k]; }
}
#endif

#ifdef __cplusplus
}
#endif
