/* ORACLE (test infrastructure only -- see suews_oracle.h).
 * Restates SUEWS-SourceCode/suews_ctrl_driver.f95: SUEWS_cal_Main :45-1045,
 * the stage wrappers :1049-2270, SUEWS_update_outputLine :2273-2457,
 * SUEWS_update_output :2461-2509, SUEWS_cal_multitsteps :2814-3505,
 * cal_tair_av :3530, cal_tsfc :3558, set_nan :2737.
 *
 * Not restated (fatal code 1003 when selected): StorageHeatMethod 3 (AnOHM)
 * -- needs module state outside the cal_Main argument
 * list.  SOLWEIG_cal_main (diagnostic only) is not restated: its output line
 * is -999-filled. */
#include "suews_oracle.h"
extern double *so_debug_trace;
#include <math.h>
#include <string.h>

#define FMAX(a, b) ((a) > (b) ? (a) : (b))

/* driver.f95:2737-2750 */
double so_set_nan(double x)
{
   const double pNAN = 30000, NAN_ = -999;
   return (fabs(x) > pNAN) ? NAN_ : x;
}

/* driver.f95:3530-3556: 273.15 and tstep*1. are REAL(4) */
double so_cal_tair_av(double tair_av_prev, int dt_since_start, int tstep, double temp_c)
{
   const double len_day_s = 24 * 3600;
   double len_cal_s, temp_k;
   if (dt_since_start > len_day_s) {
      len_cal_s = len_day_s;
   } else {
      len_cal_s = dt_since_start + tstep;
   }
   temp_k = temp_c + F(273.15);
   return tair_av_prev * (len_cal_s - (double)((float)tstep * 1.f)) / len_cal_s + temp_k * tstep / len_cal_s;
}

/* driver.f95:3558-3571 */
double so_cal_tsfc(double qh, double avdens, double avcp, double RA, double temp_c)
{
   return qh / (avdens * avcp) * RA + temp_c;
}

/* driver.f95:2582-2598 */
static void cal_surf(const double sfr[NSURF], double *VegFraction, double *ImpervFraction, double *PervFraction,
                     double *NonWaterFraction)
{
   *VegFraction = sfr[CONIF] + sfr[DECID] + sfr[GRASS];
   *ImpervFraction = sfr[PAV] + sfr[BLDG];
   *PervFraction = 1 - *ImpervFraction;
   *NonWaterFraction = 1 - sfr[WATER];
}

/* driver.f95:1049-1154 */
static void cal_AnthropogenicEmission(SoCtx *cx, const SoParams *p, const int dayofWeek_id[3], int DLS,
                                      const double HDD_id[12], int imin, int it, double QF_obs, double Temp_C,
                                      double *QF, double *QF_SAHP, double *Fc_anthro, double *Fc_build,
                                      double *Fc_metab, double *Fc_point, double *Fc_traff)
{
   int EmissionsMethod = p->EmissionsMethod;
   if (EmissionsMethod == 0) {
      *QF = QF_obs;
   } else if ((EmissionsMethod > 0 && EmissionsMethod <= 6) || EmissionsMethod >= 11) {
      so_AnthropogenicEmissions(cx, p->CO2PointSource, EmissionsMethod, it, imin, DLS, dayofWeek_id,
                                p->EF_umolCO2perJ, p->FcEF_v_kgkm, p->EnEF_v_Jkm, p->TrafficUnits,
                                p->FrFossilFuel_Heat, p->FrFossilFuel_NonHeat, p->MinFCMetab, p->MaxFCMetab,
                                p->MinQFMetab, p->MaxQFMetab, p->PopDensDaytime, p->PopDensNighttime, Temp_C, HDD_id,
                                p->Qf_A, p->Qf_B, p->Qf_C, p->AH_MIN, p->AH_SLOPE_Heating, p->AH_SLOPE_Cooling,
                                p->BaseT_Heating, p->BaseT_Cooling, p->TrafficRate, p->QF0_BEU, QF_SAHP, Fc_anthro,
                                Fc_metab, Fc_traff, Fc_build, Fc_point, p->AHProf_24hr, p->HumActivity_24hr,
                                p->TraffProf_24hr, p->PopProf_24hr, p->SurfaceArea);
   } else {
      so_error_hint(cx, 73, "RunControl.nml:EmissionsMethod unusable", -999, -999, EmissionsMethod);
   }
   if (EmissionsMethod >= 1) *QF = *QF_SAHP;
   if (EmissionsMethod >= 0 && EmissionsMethod <= 6) {
      *Fc_anthro = 0; *Fc_metab = 0; *Fc_traff = 0; *Fc_build = 0; *Fc_point = 0;
   }
}

/* driver.f95:1288-1455 */
static void cal_Qn(SoCtx *cx, const SoParams *p, const SoForcing *f, const double SnowPack_prev[NSURF],
                   double dectime, double ZENITH_deg, double Tsurf_0, double Temp_C, double ea_hPa,
                   double SnowAlb_prev, const double snowFrac_prev[NSURF], const double IceFrac[NSURF],
                   const double alb_prev[NSURF], double albDecTr_id, double albEveTr_id, double albGrass_id,
                   double alb_next[NSURF], double *ldown, double *fcld, double *qn1, double *qn1_snowfree,
                   double *qn1_S, double *kclear, double *kup, double *lup, double *tsurf,
                   double qn1_ind_snow[NSURF], double kup_ind_snow[NSURF], double Tsurf_ind_snow[NSURF],
                   double Tsurf_ind[NSURF], double *alb1, double snowFrac_next[NSURF], double *SnowAlb_next)
{
   const double NAN_ = -999;
   double alb[NSURF], SnowFrac[NSURF], SnowAlb, alb0;
   int NetRadiationMethod_use, AlbedoChoice, ldown_option, is;

   for (is = 0; is < NSURF; is++) alb[is] = alb_prev[is];
   SnowAlb = so_update_snow_albedo(p->tstep, SnowPack_prev, SnowAlb_prev, Temp_C, p->tau_a, p->tau_f, p->SnowAlbMax,
                                   p->SnowAlbMin);
   so_RadMethod(cx, p->NetRadiationMethod, p->snowUse, &NetRadiationMethod_use, &AlbedoChoice, &ldown_option);

   for (is = 0; is < NSURF; is++) SnowFrac[is] = snowFrac_prev[is];
   if (NetRadiationMethod_use > 0) {
      if (p->snowUse == 0)
         for (is = 0; is < NSURF; is++) SnowFrac[is] = 0;
      if (ldown_option == 1) {
         *ldown = f->ldown_obs;
      } else {
         *ldown = -9;
      }
      if (ldown_option == 2) *fcld = f->fcld_obs;

      alb[DECID] = albDecTr_id;
      alb[CONIF] = albEveTr_id;
      alb[GRASS] = albGrass_id;

      so_NARP(cx, p->sfr, SnowFrac, alb, p->emis, IceFrac, p->NARP_TRANS_SITE, p->NARP_EMIS_SNOW, dectime, ZENITH_deg,
              Tsurf_0, f->avkdn, Temp_C, f->avRh, ea_hPa, f->qn1_obs, SnowAlb, AlbedoChoice, ldown_option,
              NetRadiationMethod_use, p->DiagQN, qn1, qn1_snowfree, qn1_S, kclear, kup, ldown, lup, fcld, tsurf,
              qn1_ind_snow, kup_ind_snow, Tsurf_ind_snow, Tsurf_ind, &alb0, alb1);
   } else {
      for (is = 0; is < NSURF; is++) SnowFrac[is] = f->snowFrac_obs;
      *qn1 = f->qn1_obs;
      *qn1_snowfree = f->qn1_obs;
      *qn1_S = f->qn1_obs;
      *ldown = NAN_;
      *lup = NAN_;
      *kup = NAN_;
      *tsurf = NAN_;
      for (is = 0; is < NSURF; is++) Tsurf_ind[is] = NAN_;
      *fcld = NAN_;
   }
   for (is = 0; is < NSURF; is++) snowFrac_next[is] = SnowFrac[is];

   if (ldown_option == 1) *fcld = NAN_;

   for (is = 0; is < NSURF; is++) alb_next[is] = alb[is];
   *SnowAlb_next = SnowAlb;
}

/* driver.f95:1457-1612 */
static void cal_Qs(SoCtx *cx, const SoParams *p, const SoForcing *f, int dt_since_start, SoEstmState *estm,
                   SoAnohmState *anohm, const double alb_next[NSURF], const double *StoreDrainPrm_next,
                   double zenith_deg, double ldown, double Tair_av,
                   const double soilstore_id[NSURF], const double state_id[NSURF], const double SnowFrac[NSURF],
                   const double HDD_id[12], double qf, double qn1, double qn1_av_prev, double dqndt_prev,
                   double qn1_s_av_prev, double dqnsdt_prev, double qn1_S, double dataOutLineESTM[NCOL_ESTM],
                   double *qs, double *qn1_av_next, double *dqndt_next, double *qn1_s_av_next, double *dqnsdt_next,
                   double deltaQi[NSURF], double *a1, double *a2, double *a3)
{
   double qn1_use = 0, Tair_mav_5d;
   int i;
   for (i = 0; i < NCOL_ESTM; i++) dataOutLineESTM[i] = -999;
   *qs = -999; *a1 = -999; *a2 = -999; *a3 = -999;

   if (p->OHMIncQF == 1) {
      qn1_use = qf + qn1;
   } else if (p->OHMIncQF == 0) {
      qn1_use = qn1;
   }

   if (p->StorageHeatMethod == 0) {
      *qs = f->qs_obs;
   } else if (p->StorageHeatMethod == 1) {
      Tair_mav_5d = HDD_id[9];
      so_OHM(cx, qn1_use, qn1_av_prev, dqndt_prev, qn1_av_next, dqndt_next, qn1_S, qn1_s_av_prev, dqnsdt_prev,
             qn1_s_av_next, dqnsdt_next, p->tstep, dt_since_start, p->sfr, Tair_mav_5d, p->OHM_coef, p->OHM_threshSW,
             p->OHM_threshWD, soilstore_id, p->SoilStoreCap, state_id, p->snowUse, SnowFrac, p->DiagQS, a1, a2, a3, qs,
             deltaQi);
   } else if (p->StorageHeatMethod == 3) {
      /* driver.f95:1577-1594; MetForcingData_grid is the host's whole forcing block of the grid */
      double moist_surf[NSURF];
      int is;
      for (is = 0; is < NSURF; is++) moist_surf[is] = state_id[is] / SDP(StoreDrainPrm_next, 5, is);
      if (!so_met_grid) {
         so_error_hint(cx, 1006, "oracle: StorageHeatMethod=3 (AnOHM) needs MetForcingData_grid", -999, -999, 3);
      } else {
         so_AnOHM(cx, anohm, p->tstep, dt_since_start, qn1_use, qn1_av_prev, dqndt_prev, qf, so_met_grid,
                  so_met_grid_nrows, moist_surf, alb_next, p->emis, p->cpAnOHM, p->kkAnOHM, p->chAnOHM, p->sfr,
                  p->EmissionsMethod, f->id, p->Gridiv, qn1_av_next, dqndt_next, a1, a2, a3, qs, deltaQi);
      }
   } else if (p->StorageHeatMethod == 4 || p->StorageHeatMethod == 14) {
      so_ESTM(cx, p, estm, p->tstep, f->avkdn, f->avU1, f->Temp_C, zenith_deg, f->avRh, f->Press_hPa, ldown, p->bldgH,
              f->Ts5mindata_ir, Tair_av, dataOutLineESTM, qs);
   }
}

/* driver.f95:1614-1694 */
static void cal_Water(SoCtx *cx, const SoParams *p, double NonWaterFraction, double addPipes, double addImpervious,
                      double addVeg, double addWaterBody, const double state_id[NSURF],
                      const double soilstore_id[NSURF], const double *StoreDrainPrm, double nsh_real,
                      double *drain_per_tstep, double drain[NSURF], double frac_water2runoff[NSURF],
                      double *AdditionalWater, double *runoffPipes, double *runoff_per_interval,
                      double AddWater[NSURF], double stateOld[NSURF], double soilstoreOld[NSURF])
{
   int is;
   double s;
   for (is = 0; is < NSURF; is++) { stateOld[is] = state_id[is]; soilstoreOld[is] = soilstore_id[is]; }
   *AdditionalWater = addPipes + addImpervious + addVeg + addWaterBody;
   *runoffPipes = addPipes;
   *runoff_per_interval = addPipes;

   if (NonWaterFraction != 0) {
      for (is = 0; is < NSURF - 1; is++) {
         so_drainage(cx, is, state_id[is], SDP(StoreDrainPrm, 5, is), SDP(StoreDrainPrm, 1, is),
                     SDP(StoreDrainPrm, 2, is), SDP(StoreDrainPrm, 3, is), nsh_real, &drain[is]);
      }
      s = 0;
      for (is = 0; is < NSURF - 1; is++) s += drain[is] * p->sfr[is];
      *drain_per_tstep = s / NonWaterFraction;
   } else {
      for (is = 0; is < NSURF - 1; is++) drain[is] = 0;
      *drain_per_tstep = 0;
   }
   drain[WATER] = 0;

   so_ReDistributeWater(p->snowUse, p->WaterDist, p->sfr, drain, frac_water2runoff, AddWater);
}

/* driver.f95:2134-2270 */
static void cal_Resistance(SoCtx *cx, const SoParams *p, const SoForcing *f, double avdens, double avcp,
                           double QH_init, double zzd, double z0m, double zdm, double Temp_C, double VegFraction,
                           double dq, double vsmd, const double LAI_id[NVEG], const double SnowFrac[NSURF],
                           double *UStar, double *TStar, double *L_mod, double *zL, double *gsc, double *ResistSurf,
                           double *RA, double *RAsnow, double *rb)
{
   double gfunc;
   so_cal_Stab(cx, p->StabilityMethod, zzd, z0m, zdm, f->avU1, Temp_C, QH_init, avdens, avcp, L_mod, TStar, UStar,
               zL);
   so_AerodynamicResistance(cx, zzd, z0m, f->avU1, *L_mod, *UStar, VegFraction, p->AerodynamicResistanceMethod,
                            p->StabilityMethod, p->RoughLenHeatMethod, RA);
   if (p->snowUse == 1) {
      so_AerodynamicResistance(cx, zzd, z0m, f->avU1, *L_mod, *UStar, VegFraction, p->AerodynamicResistanceMethod,
                               p->StabilityMethod, 3, RAsnow);
   }
   so_SurfaceResistance(cx, f->id, f->it, p->SMDMethod, SnowFrac, p->sfr, f->avkdn, Temp_C, dq, f->xsmd, vsmd,
                        p->MaxConductance, p->LAIMax, LAI_id, p->gsModel, p->Kmax, p->G1, p->G2, p->G3, p->G4, p->G5,
                        p->G6, p->TH, p->TL, p->S1, p->S2, &gfunc, gsc, ResistSurf);
   so_BoundaryLayerResistance(zzd, z0m, f->avU1, UStar, rb);
}

/* driver.f95:1744-2082 */
static void cal_QE(SoCtx *cx, const SoParams *p, const SoForcing *f, const int snowCalcSwitch[NSURF],
                   const int dayofWeek_id[3], double dectime, double avdens, double avcp, double lv_J_kg,
                   double lvS_J_kg, double Temp_C, double RAsnow, double psyc_hPa, double sIce_hPa,
                   double PervFraction, double vegfraction, double addimpervious, double qn1_snowfree, double qf,
                   double qs, double vpd_hPa, double s_hPa, double ResistSurf, double RA, double rb,
                   double NonWaterFraction, const double WU_nsurf[NSURF], double addVeg, double addWaterBody,
                   const double drain[NSURF], const double stateOld[NSURF], const double mw_ind[NSURF],
                   const double rainonsnow[NSURF], const double freezmelt[NSURF], const double freezstate[NSURF],
                   const double freezstatevol[NSURF], const double Qm_Melt[NSURF], const double Qm_rain[NSURF],
                   const double Tsurf_ind[NSURF], const double AddWater[NSURF], const double addwaterrunoff[NSURF],
                   const double *StoreDrainPrm, const double SnowPack_in[NSURF], const double SnowFrac_in[NSURF],
                   const double SnowWater_in[NSURF], const double iceFrac_in[NSURF], const double SnowDens_in[NSURF],
                   double runoff_per_interval_in, const double state_id_in[NSURF],
                   const double soilstore_id_in[NSURF], double state_id_out[NSURF], double soilstore_id_out[NSURF],
                   double SnowPack_out[NSURF], double SnowFrac_out[NSURF], double SnowWater_out[NSURF],
                   double iceFrac_out[NSURF], double SnowDens_out[NSURF], double runoffSoil[NSURF],
                   double SnowRemoval[2], double *state_per_tstep, double *NWstate_per_tstep, double *qe, double *swe,
                   double *chSnow_per_interval, double *ev_per_tstep, double *runoff_per_tstep,
                   double *surf_chang_per_tstep, double *runoffPipes, double *mwstore, double *runoffwaterbody,
                   double *runoffAGveg, double *runoffAGimpervious)
{
   double runoffSnow[NSURF], runoff[NSURF], chang[NSURF], ChangSnow[NSURF], SnowToSurf[NSURF], state_id[NSURF],
      ev_snow[NSURF], soilstore_id[NSURF], SnowDepth[NSURF], SnowPack[NSURF], SnowFrac[NSURF], SnowWater[NSURF],
      iceFrac[NSURF], SnowDens[NSURF], rss_nsurf[NSURF], SurplusEvap[NSURF], evap[NSURF];
   double p_mm, qe_surf, ev, qe_per_tstep, surplusWaterBody, pin, qn_e, tlv, nsh_real, tstep_real, ev_tot, qe_tot,
      surf_chang_tot, runoff_tot, chSnow_tot, runoff_per_interval, capStore_is;
   int is;

   runoff_per_interval = runoff_per_interval_in;
   for (is = 0; is < NSURF; is++) {
      state_id[is] = state_id_in[is];
      soilstore_id[is] = soilstore_id_in[is];
      SnowPack[is] = SnowPack_in[is];
      SnowFrac[is] = SnowFrac_in[is];
      SnowWater[is] = SnowWater_in[is];
      iceFrac[is] = iceFrac_in[is];
      SnowDens[is] = SnowDens_in[is];
   }

   tstep_real = p->tstep * 1.0;
   nsh_real = 3600 / tstep_real;
   tlv = lv_J_kg / tstep_real;
   pin = FMAX(F(0.), f->Precip);

   qe_surf = 0; ev = 0; qe_tot = 0; ev_tot = 0; *swe = 0;
   *ev_per_tstep = 0; qe_per_tstep = 0; *surf_chang_per_tstep = 0; *runoff_per_tstep = 0; *state_per_tstep = 0;
   *NWstate_per_tstep = 0; *qe = 0; *runoffwaterbody = 0; *chSnow_per_interval = 0; *mwstore = 0; *runoffAGveg = 0;
   *runoffAGimpervious = 0; surplusWaterBody = 0;
   for (is = 0; is < NSURF; is++) {
      ev_snow[is] = 0; runoffSoil[is] = 0; runoff[is] = 0; chang[is] = 0; SurplusEvap[is] = 0; rss_nsurf[is] = 0;
      runoffSnow[is] = 0; ChangSnow[is] = 0; SnowToSurf[is] = 0; SnowDepth[is] = 0; evap[is] = 0;
   }
   SnowRemoval[0] = 0; SnowRemoval[1] = 0;
   surf_chang_tot = 0; runoff_tot = 0; chSnow_tot = 0;

   qn_e = qn1_snowfree + qf - qs;

   for (is = 0; is < NSURF; is++) {
      if (p->snowUse == 1 && snowCalcSwitch[is] == 1) {
         if (p->sfr[is] != 0) {
            so_SnowCalc(cx, p->tstep, f->imin, f->it, dectime, is, p->EvapMethod, p->CRWmin, p->CRWmax, nsh_real,
                        lvS_J_kg, avdens, f->avRh, f->Press_hPa, Temp_C, RAsnow, psyc_hPa, avcp, sIce_hPa,
                        PervFraction, vegfraction, addimpervious, vpd_hPa, qn_e, s_hPa, ResistSurf, RA, rb, tlv,
                        p->SnowDensMin, p->SnowProf_24hr, f->Precip, p->PipeCapacity, p->RunoffToWater, addVeg,
                        p->SnowLimPaved, p->SnowLimBldg, p->FlowChange, drain, p->WetThresh, stateOld, mw_ind,
                        p->SoilStoreCap, rainonsnow, freezmelt, freezstate, freezstatevol, Qm_Melt, Qm_rain,
                        Tsurf_ind, p->sfr, dayofWeek_id, StoreDrainPrm, p->SnowPackLimit, AddWater, addwaterrunoff,
                        soilstore_id, SnowPack, SurplusEvap, SnowFrac, SnowWater, iceFrac, SnowDens,
                        runoffAGimpervious, runoffAGveg, &surplusWaterBody, rss_nsurf, runoffSnow, runoff, runoffSoil,
                        chang, ChangSnow, SnowToSurf, state_id, ev_snow, SnowDepth, SnowRemoval, swe, &ev,
                        &chSnow_tot, &ev_tot, &qe_tot, &runoff_tot, &surf_chang_tot, runoffPipes, mwstore,
                        runoffwaterbody);

            *ev_per_tstep = *ev_per_tstep + ev_tot;
            qe_per_tstep = qe_per_tstep + qe_tot;
            *runoff_per_tstep = *runoff_per_tstep + runoff_tot;
            *surf_chang_per_tstep = *surf_chang_per_tstep + surf_chang_tot;
            *chSnow_per_interval = *chSnow_per_interval + chSnow_tot;
         } else {
            SnowFrac[is] = 0;
            SnowDens[is] = 0;
            SnowPack[is] = 0;
         }
         evap[is] = ev_tot;
      } else {
         capStore_is = SDP(StoreDrainPrm, 5, is);
         so_cal_evap(p->EvapMethod, state_id[is], p->WetThresh[is], capStore_is, vpd_hPa, avdens, avcp, qn_e, s_hPa,
                     psyc_hPa, ResistSurf, RA, rb, tlv, &rss_nsurf[is], &ev, &qe_surf);

         so_cal_water_storage(cx, is, p->sfr, p->PipeCapacity, p->RunoffToWater, pin, WU_nsurf, drain, AddWater,
                              addimpervious, nsh_real, stateOld, addwaterrunoff, PervFraction, addVeg, p->SoilStoreCap,
                              addWaterBody, p->FlowChange, p->StateLimit, runoffAGimpervious, &surplusWaterBody,
                              runoffAGveg, runoffPipes, &ev, soilstore_id, SurplusEvap, runoffwaterbody, &p_mm, chang,
                              runoff, state_id);

         evap[is] = ev;
         *ev_per_tstep = *ev_per_tstep + evap[is] * p->sfr[is];
         qe_per_tstep = qe_per_tstep + qe_surf * p->sfr[is];
         *surf_chang_per_tstep = *surf_chang_per_tstep + (state_id[is] - stateOld[is]) * p->sfr[is];
         *runoff_per_tstep = *runoff_per_tstep + runoff[is] * p->sfr[is];
         *state_per_tstep = *state_per_tstep + state_id[is] * p->sfr[is];
         runoff_per_interval = runoff_per_interval + (runoff[is] * p->sfr[is]);
         if (NonWaterFraction != 0 && is != WATER) {
            *NWstate_per_tstep = *NWstate_per_tstep + (state_id[is] * p->sfr[is] / NonWaterFraction);
         }
         ChangSnow[is] = 0;
         runoffSnow[is] = 0;
      }
   }

   *qe = qe_per_tstep;

   for (is = 0; is < NSURF; is++) {
      state_id_out[is] = state_id[is];
      soilstore_id_out[is] = soilstore_id[is];
      SnowPack_out[is] = SnowPack[is];
      SnowFrac_out[is] = SnowFrac[is];
      SnowWater_out[is] = SnowWater[is];
      iceFrac_out[is] = iceFrac[is];
      SnowDens_out[is] = SnowDens[is];
   }
}

/* driver.f95:2084-2131 (QHMethod fixed to 1 by the caller, driver.f95:879) */
static void cal_QH(double qn1, double qf, double QmRain, double qeOut, double qs, double QmFreez, double qm,
                   double avdens, double avcp, double tsurf, double Temp_C, double RA, double *qh,
                   double *qh_residual, double *qh_resist)
{
   *qh_residual = (qn1 + qf + QmRain) - (qeOut + qs + qm + QmFreez);
   if (RA != 0) {
      *qh_resist = avdens * avcp * (tsurf - Temp_C) / RA;
   } else {
      *qh_resist = -999;
   }
   *qh = *qh_residual;
}

/* driver.f95:1156-1286 */
static void cal_BiogenCO2(SoCtx *cx, const SoParams *p, const SoForcing *f, double dectime, double Fc_anthro,
                          double gfunc, const double LAI_id[NVEG], const double SnowFrac[NSURF], double t2_C,
                          double Temp_C, double vsmd, double *Fc, double *Fc_biogen, double *Fc_photo,
                          double *Fc_respi)
{
   double gfunc2 = 0, dq, t2 = 0, d1, d2, d3, d4, d5, d6, d7, d8, d9, d10, d11;
   int EmissionsMethod = p->EmissionsMethod, gsmodel = p->gsModel;

   if (EmissionsMethod >= 11) {
      if (gsmodel == 3 || gsmodel == 4) {
         if (fabs(Temp_C - t2_C) > 5) t2 = Temp_C; else t2 = t2_C;
         so_cal_AtmMoist(cx, &t2, f->Press_hPa, f->avRh, dectime, &d1, &d2, &d3, &d4, &d5, &d6, &dq, &d7, &d8, &d9);
         so_SurfaceResistance(cx, f->id, f->it, p->SMDMethod, SnowFrac, p->sfr, f->avkdn, t2, dq, f->xsmd, vsmd,
                              p->MaxConductance, p->LAIMax, LAI_id, gsmodel, p->Kmax, p->G1, p->G2, p->G3, p->G4,
                              p->G5, p->G6, p->TH, p->TL, p->S1, p->S2, &gfunc2, &d10, &d11);
      }
      so_CO2_biogen(cx, p->alpha_bioCO2, p->alpha_enh_bioCO2, f->avkdn, p->beta_bioCO2, p->beta_enh_bioCO2, dectime,
                    EmissionsMethod, gfunc, gfunc2, gsmodel, f->id, f->it, LAI_id, p->LAIMin, p->LAIMax,
                    p->min_res_bioCO2, p->resp_a, p->resp_b, p->sfr, SnowFrac, t2, Temp_C, p->theta_bioCO2, Fc_biogen,
                    Fc_photo, Fc_respi);
   }
   if (EmissionsMethod >= 0 && EmissionsMethod <= 6) {
      *Fc_biogen = 0; *Fc_photo = 0; *Fc_respi = 0;
   }
   *Fc = Fc_anthro + *Fc_biogen;
}

/* ===== SUEWS_cal_Main driver.f95:45-1045 ===== */
void so_cal_main(SoCtx *cx, const SoParams *p, const SoForcing *f, SoState *st, SoOutLines *out)
{
   /* locals (driver.f95:348-568) */
   double a1 = 0, a2 = 0, a3 = 0, AdditionalWater = 0, U10_ms, azimuth = 0, chSnow_per_interval = 0, dens_dry;
   double deltaLAI = 0, drain_per_tstep = 0, Ea_hPa = 0, QE_LUMPS = 0, es_hPa, ev_per_tstep = 0;
   double wu_ext = 0, Fc, Fc_anthro = 0, Fc_biogen = 0, Fc_build = 0, fcld = 0, Fc_metab = 0,
          Fc_photo = 0, Fc_point = 0, Fc_respi = 0, Fc_traff = 0;
   double gfunc = 0, gsc, QH_LUMPS = 0, wu_int = 0, kclear = 0, kup = 0, ldown = 0, lup = 0, L_mod = 0, mwh = 0,
          mwstore = 0, NWstate_per_tstep = 0, FAI = 0, zL = 0, q2_gkg, qe = 0, qf = 0, QF_SAHP = 0, qh = 0,
          qh_residual, qh_resist = 0, Qm = 0, QmFreez = 0, QmRain = 0, qn1 = 0, qn1_S = 0, qn1_snowfree = 0, qs = 0,
          RA = 0, ResistSurf = 0, RH2, runoffAGveg = 0, runoffAGimpervious = 0, runoff_per_tstep = 0, runoffPipes = 0,
          runoffSoil_per_tstep = 0, runoffwaterbody = 0, smd = 0, SoilState, state_per_tstep = 0,
          surf_chang_per_tstep = 0, swe = 0, t2_C, TSfc_C = 0, TempVeg, tot_chang_per_tstep = 0, TStar, tsurf = 0,
          UStar = 0, VPD_Pa, z0m = 0, zdm = 0, ZENITH_deg = 0, zH = 0;
   double SnowRemoval[2] = {0, 0}, wu_nsurf[NSURF], FreezMelt[NSURF], kup_ind_snow[NSURF], mw_ind[NSURF],
          Qm_freezState[NSURF], Qm_melt[NSURF], Qm_rain[NSURF], qn1_ind_snow[NSURF], rainOnSnow[NSURF],
          runoffSoil[NSURF], smd_nsurf[NSURF], snowDepth[NSURF], Tsurf_ind_snow[NSURF];
   int snowCalcSwitch[NSURF], dayofWeek_id[3], DLS;
   double dq, lvS_J_kg, psyc_hPa = 0, RAsnow = 0, rb = 0, runoff_per_interval, s_hPa = 0,
          sIce_hpa = 0, SoilMoistCap, veg_fr, VegPhenLumps = 0, VPd_hpa, vsmd, ZZD = 0;
   double avcp, avdens, lv_J_kg; /* module moist globals in the reference (driver.f95:38) */
   double deltaQi[NSURF], drain[NSURF], FreezState[NSURF], FreezStateVol[NSURF], soilstore_updated[NSURF],
      state_id_updated[NSURF], tsurf_ind[NSURF];
   const double addImpervious = 0, addPipes = 0, addVeg = 0, addWaterBody = 0;
   double AddWater[NSURF], frac_water2runoff[NSURF];
   int nsh;
   double nsh_real, tstep_real, dectime;
   double VegFraction, ImpervFraction, PervFraction, NonWaterFraction;
   double alb1 = st->alb1_hidden;
   SoState pv, nx; /* *_prev / *_next copies (driver.f95:572-631) */
   int flag_converge, i_iter, is, i, k;
   double Ts_iter, QH_Init;
   double Temp_C = f->Temp_C; /* local copy: sat_vap_press_x may rewrite it (meteo:161-168) */
   SoForcing fl = *f;          /* forcing view whose Temp_C follows the rewrite */
   const double NAN_ = -999;

   cx->lv_switch1 = &st->lv_switch1;
   cx->lv_switch2 = &st->lv_switch2;

   for (is = 0; is < NSURF; is++) {
      wu_nsurf[is] = 0; FreezMelt[is] = 0; kup_ind_snow[is] = 0; mw_ind[is] = 0;
      Qm_freezState[is] = 0; Qm_melt[is] = 0; Qm_rain[is] = 0; qn1_ind_snow[is] = 0;
      rainOnSnow[is] = 0; runoffSoil[is] = 0; smd_nsurf[is] = 0; snowDepth[is] = 0;
      Tsurf_ind_snow[is] = 0; deltaQi[is] = 0; drain[is] = 0;
      FreezState[is] = 0; FreezStateVol[is] = 0; tsurf_ind[is] = 0; AddWater[is] = 0; frac_water2runoff[is] = 0;
      snowCalcSwitch[is] = 0;
   }

   pv = *st;
   nx = *st;

   flag_converge = 0;
   Ts_iter = Temp_C;
   i_iter = 1;
   while ((!flag_converge) && i_iter < 100) {
      /* SO_CNT_*: op-counter regions of the counting build (oracle/so_count.h; no-ops in the normal build).  In "hoisted"
       * mode the stages whose inputs do not change from trip to trip are counted on the first trip only: that is the
       * minimal algorithm SURVEY 8(d) defines the algorithmic flops on; the arithmetic itself is untouched. */
      SO_CNT_ONCE_BEGIN(i_iter > 1);
      so_cal_dectime(f->id, f->it, f->imin, f->isec, &dectime);
      so_cal_tstep(p->tstep, &nsh, &nsh_real, &tstep_real);
      cal_surf(p->sfr, &VegFraction, &ImpervFraction, &PervFraction, &NonWaterFraction);
      so_cal_weekday(f->iy, f->id, p->lat, dayofWeek_id);
      so_cal_DLS(f->id, p->startDLS, p->endDLS, &DLS);

      nx.Tair_av = so_cal_tair_av(pv.Tair_av, st->dt_since_start, p->tstep, Temp_C);

      so_cal_RoughnessParameters(cx, p->RoughLenMomMethod, p->sfr, p->bldgH, p->EveTreeH, p->DecTreeH, pv.porosity_id,
                                 p->FAIBldg, p->FAIEveTree, p->FAIDecTree, p->z0m_in, p->zdm_in, p->Z, &FAI, &zH, &z0m,
                                 &zdm, &ZZD);

      /* dectime [days] minus the INTEGER tstep/2 [seconds]: as the reference (driver.f95:689) */
      so_NARP_cal_SunPosition((double)f->iy, dectime - p->tstep / 2, p->timezone, p->lat, p->lng, p->alt, &azimuth,
                              &ZENITH_deg);

      fl.Temp_C = Temp_C;
      so_cal_DailyState(cx, p, &fl, st->tstep_prev, st->dt_since_start, dayofWeek_id, pv.Tmin_id, pv.Tmax_id,
                        pv.lenDay_id, nsh_real, pv.DecidCap_id, pv.StoreDrainPrm, pv.LAI_id, pv.GDD_id, pv.SDD_id,
                        pv.albDecTr_id, pv.albEveTr_id, pv.albGrass_id, pv.porosity_id, pv.HDD_id, pv.state_id,
                        pv.soilstore_id, nx.HDD_id, &nx.Tmin_id, &nx.Tmax_id, &nx.lenDay_id, &nx.albDecTr_id,
                        &nx.albEveTr_id, &nx.albGrass_id, &nx.porosity_id, &nx.DecidCap_id, nx.StoreDrainPrm,
                        nx.LAI_id, nx.GDD_id, nx.SDD_id, &deltaLAI, nx.WUDay_id);
      SO_CNT_END;

      so_cal_AtmMoist(cx, &Temp_C, f->Press_hPa, f->avRh, dectime, &lv_J_kg, &lvS_J_kg, &es_hPa, &Ea_hPa, &VPd_hpa,
                      &VPD_Pa, &dq, &dens_dry, &avcp, &avdens);
      fl.Temp_C = Temp_C;

      SO_CNT_ONCE_BEGIN(i_iter > 1);
      so_update_SoilMoist(NonWaterFraction, p->SoilStoreCap, p->sfr, pv.soilstore_id, &SoilMoistCap, &SoilState, &vsmd,
                          &smd);

      so_cal_WaterUse(cx, nsh_real, f->wu_m3, p->SurfaceArea, p->sfr, p->IrrFracPaved, p->IrrFracBldgs,
                      p->IrrFracEveTr, p->IrrFracDecTr, p->IrrFracGrass, p->IrrFracBSoil, p->IrrFracWater, dayofWeek_id,
                      p->WUProfA_24hr, p->WUProfM_24hr, p->InternalWaterUse_h, nx.HDD_id, nx.WUDay_id,
                      p->WaterUseMethod, nsh, f->it, f->imin, DLS, wu_nsurf, &wu_int, &wu_ext);

      cal_AnthropogenicEmission(cx, p, dayofWeek_id, DLS, nx.HDD_id, f->imin, f->it, f->qf_obs, Temp_C, &qf, &QF_SAHP,
                                &Fc_anthro, &Fc_build, &Fc_metab, &Fc_point, &Fc_traff);
      SO_CNT_END;

      /* radiation, storage heat and melt heat repeat per trip only when the surface temperature feeds back into NARP
       * (NetRadiationMethod 11-13) or ESTM advances its element temperatures (StorageHeatMethod 4/14) */
      SO_CNT_ONCE_BEGIN(i_iter > 1 && !((p->NetRadiationMethod >= 10 && p->NetRadiationMethod < 100) || p->StorageHeatMethod == 4 || p->StorageHeatMethod == 14));
      cal_Qn(cx, p, &fl, pv.SnowPack, dectime, ZENITH_deg, Ts_iter, Temp_C, Ea_hPa, pv.SnowAlb, pv.SnowFrac,
             pv.IceFrac, pv.alb, nx.albDecTr_id, nx.albEveTr_id, nx.albGrass_id, nx.alb, &ldown, &fcld, &qn1,
             &qn1_snowfree, &qn1_S, &kclear, &kup, &lup, &tsurf, qn1_ind_snow, kup_ind_snow, Tsurf_ind_snow, tsurf_ind,
             &alb1, nx.SnowFrac, &nx.SnowAlb);

      cal_Qs(cx, p, &fl, st->dt_since_start, &st->estm, &st->anohm, nx.alb, nx.StoreDrainPrm, ZENITH_deg, ldown, nx.Tair_av, pv.soilstore_id, pv.state_id, nx.SnowFrac, nx.HDD_id, qf, qn1, pv.qn1_av,
             pv.dqndt, pv.qn1_s_av, pv.dqnsdt, qn1_S, out->ESTM, &qs, &nx.qn1_av, &nx.dqndt, &nx.qn1_s_av, &nx.dqnsdt,
             deltaQi, &a1, &a2, &a3);

      so_Snow_cal_MeltHeat(cx, p->snowUse, p->tstep, p->tau_r, p->SnowDensMax, lvS_J_kg, lv_J_kg, tstep_real,
                           p->RadMeltFact, p->TempMeltFact, p->SnowAlbMax, p->SnowDensMin, Temp_C, f->Precip,
                           p->PrecipLimit, p->PrecipLimitAlb, nsh_real, p->sfr, tsurf_ind, Tsurf_ind_snow, pv.state_id,
                           qn1_ind_snow, kup_ind_snow, pv.SnowWater, deltaQi, alb1, pv.SnowPack, nx.SnowFrac,
                           nx.SnowAlb, pv.SnowDens, pv.SnowfallCum, nx.SnowPack, nx.SnowFrac, &nx.SnowAlb, nx.SnowDens,
                           &nx.SnowfallCum, &mwh, &Qm, &QmFreez, &QmRain, &veg_fr, snowCalcSwitch, Qm_melt,
                           Qm_freezState, Qm_rain, FreezMelt, FreezState, FreezStateVol, rainOnSnow, snowDepth, mw_ind,
                           out->Snow);
      SO_CNT_END;

      if (i_iter == 1) {
         so_LUMPS_cal_QHQE(cx, p->veg_type, p->snowUse, qn1, qf, qs, Qm, Temp_C, veg_fr, avcp, f->Press_hPa, lv_J_kg,
                           tstep_real, p->DRAINRT, nsh_real, f->Precip, p->RainMaxRes, p->RAINCOVER, p->sfr, nx.LAI_id,
                           p->LAIMax, p->LAIMin, &QH_LUMPS, &QE_LUMPS, &psyc_hPa, &s_hPa, &sIce_hpa, &TempVeg,
                           &VegPhenLumps);
         QH_Init = QH_LUMPS;
      } else {
         QH_Init = qh;
      }

      SO_CNT_ONCE_BEGIN(i_iter > 1);
      cal_Water(cx, p, NonWaterFraction, addPipes, addImpervious, addVeg, addWaterBody, pv.state_id, pv.soilstore_id,
                nx.StoreDrainPrm, nsh_real, &drain_per_tstep, drain, frac_water2runoff, &AdditionalWater, &runoffPipes,
                &runoff_per_interval, AddWater, state_id_updated, soilstore_updated);
      SO_CNT_END;

      cal_Resistance(cx, p, &fl, avdens, avcp, QH_Init, ZZD, z0m, zdm, Temp_C, VegFraction, dq, vsmd, nx.LAI_id,
                     nx.SnowFrac, &UStar, &TStar, &L_mod, &zL, &gsc, &ResistSurf, &RA, &RAsnow, &rb);

      cal_QE(cx, p, &fl, snowCalcSwitch, dayofWeek_id, dectime, avdens, avcp, lv_J_kg, lvS_J_kg, Temp_C, RAsnow,
             psyc_hPa, sIce_hpa, PervFraction, VegFraction, addImpervious, qn1_snowfree, qf, qs, VPd_hpa, s_hPa,
             ResistSurf, RA, rb, NonWaterFraction, wu_nsurf, addVeg, addWaterBody, drain, state_id_updated, mw_ind,
             rainOnSnow, FreezMelt, FreezState, FreezStateVol, Qm_melt, Qm_rain, tsurf_ind, AddWater, frac_water2runoff,
             nx.StoreDrainPrm, nx.SnowPack, nx.SnowFrac, pv.SnowWater, pv.IceFrac, nx.SnowDens, runoff_per_interval,
             pv.state_id, pv.soilstore_id, nx.state_id, nx.soilstore_id, nx.SnowPack, nx.SnowFrac, nx.SnowWater,
             nx.IceFrac, nx.SnowDens, runoffSoil, SnowRemoval, &state_per_tstep, &NWstate_per_tstep, &qe, &swe,
             &chSnow_per_interval, &ev_per_tstep, &runoff_per_tstep, &surf_chang_per_tstep, &runoffPipes, &mwstore,
             &runoffwaterbody, &runoffAGveg, &runoffAGimpervious);

      cal_QH(qn1, qf, QmRain, qe, qs, QmFreez, Qm, avdens, avcp, tsurf, Temp_C, RA, &qh, &qh_residual, &qh_resist);

      /* (every trip restarts from the *_prev state: only the last trip's soil-water exchange survives) */
      SO_CNT_ONCE_BEGIN(i_iter > 1);
      so_cal_HorizontalSoilWater(p->sfr, p->SoilStoreCap, p->SoilDepth, p->SatHydraulicConduct, p->SurfaceArea,
                                 NonWaterFraction, tstep_real, nx.soilstore_id, runoffSoil, &runoffSoil_per_tstep);

      so_cal_SoilState(cx, p->SMDMethod, f->xsmd, NonWaterFraction, SoilMoistCap, p->SoilStoreCap,
                       surf_chang_per_tstep, nx.soilstore_id, soilstore_updated, p->sfr, &smd, smd_nsurf,
                       &tot_chang_per_tstep, &SoilState);
      SO_CNT_END;

      TSfc_C = so_cal_tsfc(qh, avdens, avcp, RA, Temp_C);

      if (so_debug_trace && i_iter < 100) { /* test hook: the fixed-point sequence of the Tsurf loop */
         so_debug_trace[3 * (i_iter - 1)] = qh;
         so_debug_trace[3 * (i_iter - 1) + 1] = L_mod;
         so_debug_trace[3 * (i_iter - 1) + 2] = TSfc_C;
      }
      if (fabs(Ts_iter - TSfc_C) > F(0.1)) flag_converge = 0; else flag_converge = 1;
      Ts_iter = TSfc_C;
      i_iter = i_iter + 1;
   }
   cx->n_iter_last = i_iter - 1;
   cx->last_avcp = avcp; cx->last_avdens = avdens; cx->last_lv_J_kg = lv_J_kg;   /* MODULE moist after the call */

   /* diagnostics no flux sum and no state depends on (T2, q2, U10, RH2 and the SOLWEIG point values): counted apart */
   SO_CNT_DIAG_BEGIN();
   so_RSLProfile(cx, zH, z0m, zdm, L_mod, p->sfr, FAI, p->StabilityMethod, avcp, lv_J_kg, avdens, f->avU1, Temp_C,
                 f->avRh, f->Press_hPa, p->Z, qh, qe, &t2_C, &q2_gkg, &U10_ms, &RH2, out->RSL);
   SO_CNT_DIAG_END();

   /* BiogenCO2 still sees the un-committed SnowFrac argument (driver.f95:957) */
   cal_BiogenCO2(cx, p, &fl, dectime, Fc_anthro, gfunc, nx.LAI_id, st->SnowFrac, t2_C, Temp_C, vsmd, &Fc, &Fc_biogen,
                 &Fc_photo, &Fc_respi);

   /* commit (driver.f95:965-994); caller-side clock and hidden SAVEs are not part of *_next */
   {
      int dt = st->dt_since_start, tp = st->tstep_prev, s1 = st->lv_switch1, s2 = st->lv_switch2;
      nx.estm = st->estm;   /* MODULE ESTM_data variables: advanced in place by every trip */
      nx.anohm = st->anohm; /* AnOHM's SAVE'd locals likewise */
      *st = nx;
      st->dt_since_start = dt;
      st->tstep_prev = tp;
      st->lv_switch1 = s1;
      st->lv_switch2 = s2;
   }
   st->alb1_hidden = alb1;
   /* SOLWEIG point diagnostics (driver.f95:997-1002): the state has been committed, so alb is alb_next */
   if (p->sfr[BLDG] > 0) {
SO_CNT_DIAG_BEGIN();
            so_SOLWEIG_cal_main(f->id, f->it, dectime, 0.8, FAI, f->avkdn, ldown, Temp_C, f->avRh, f->Press_hPa, TSfc_C, p->lat,
                          ZENITH_deg, azimuth, 1.0, st->alb[0], st->alb[1], p->emis[0], p->emis[1], p->bldgH, out->SOLWEIG);
      SO_CNT_DIAG_END();
   } else {
      for (i = 0; i < NCOL_SOLWEIG; i++) out->SOLWEIG[i] = NAN_;
   }

   /* SUEWS_update_outputLine driver.f95:2273-2457 */
   {
      double state_x[NSURF], smd_nsurf_x[NSURF], LAI_wt, bulkalbedo, RH2_pct, wu_DecTr, wu_EveTr, wu_Grass;
      double *o = out->SUEWS;
      for (is = 0; is < NSURF; is++) {
         state_x[is] = (p->sfr[is] < F(0.00001)) ? NAN_ : st->state_id[is];
         smd_nsurf_x[is] = (p->sfr[is] < F(0.00001)) ? NAN_ : smd_nsurf[is];
      }
      LAI_wt = st->LAI_id[0] * p->sfr[2] + st->LAI_id[1] * p->sfr[3] + st->LAI_id[2] * p->sfr[4];
      bulkalbedo = 0;
      for (is = 0; is < NSURF; is++) bulkalbedo += st->alb[is] * p->sfr[is];
      RH2_pct = RH2 * F(100.0);
      wu_DecTr = wu_nsurf[2];
      wu_EveTr = wu_nsurf[3];
      wu_Grass = wu_nsurf[4];

      out->datetimeLine[0] = (double)f->iy;
      out->datetimeLine[1] = (double)f->id;
      out->datetimeLine[2] = (double)f->it;
      out->datetimeLine[3] = (double)f->imin;
      out->datetimeLine[4] = dectime;

      k = 0;
      o[k++] = f->avkdn; o[k++] = kup; o[k++] = ldown; o[k++] = lup; o[k++] = tsurf;
      o[k++] = qn1; o[k++] = qf; o[k++] = qs; o[k++] = qh; o[k++] = qe;
      o[k++] = QH_LUMPS; o[k++] = QE_LUMPS; o[k++] = qh_resist;
      o[k++] = f->Precip; o[k++] = wu_ext; o[k++] = ev_per_tstep; o[k++] = runoff_per_tstep;
      o[k++] = tot_chang_per_tstep;
      o[k++] = surf_chang_per_tstep; o[k++] = state_per_tstep; o[k++] = NWstate_per_tstep; o[k++] = drain_per_tstep;
      o[k++] = smd;
      o[k++] = p->FlowChange / nsh_real; o[k++] = AdditionalWater;
      o[k++] = runoffSoil_per_tstep; o[k++] = runoffPipes; o[k++] = runoffAGimpervious; o[k++] = runoffAGveg;
      o[k++] = runoffwaterbody;
      o[k++] = wu_int; o[k++] = wu_EveTr; o[k++] = wu_DecTr; o[k++] = wu_Grass;
      for (is = 0; is < NSURF - 1; is++) o[k++] = smd_nsurf_x[is];
      for (is = 0; is < NSURF; is++) o[k++] = state_x[is];
      o[k++] = ZENITH_deg; o[k++] = azimuth; o[k++] = bulkalbedo; o[k++] = fcld;
      o[k++] = LAI_wt; o[k++] = z0m; o[k++] = zdm;
      o[k++] = UStar; o[k++] = L_mod; o[k++] = RA; o[k++] = ResistSurf;
      o[k++] = Fc;
      o[k++] = Fc_photo; o[k++] = Fc_respi; o[k++] = Fc_metab; o[k++] = Fc_traff; o[k++] = Fc_build;
      o[k++] = Fc_point;
      o[k++] = qn1_snowfree; o[k++] = qn1_S; o[k++] = st->SnowAlb;
      o[k++] = Qm; o[k++] = QmFreez; o[k++] = QmRain; o[k++] = swe; o[k++] = mwh; o[k++] = mwstore;
      o[k++] = chSnow_per_interval;
      o[k++] = SnowRemoval[0]; o[k++] = SnowRemoval[1];
      o[k++] = TSfc_C; o[k++] = t2_C; o[k++] = q2_gkg; o[k++] = U10_ms; o[k++] = RH2_pct;
   }

   so_update_DailyStateLine(f->it, f->imin, nsh_real, st->GDD_id, st->HDD_id, st->LAI_id, st->SDD_id, st->Tmin_id,
                            st->Tmax_id, st->lenDay_id, st->DecidCap_id, st->albDecTr_id, st->albEveTr_id,
                            st->albGrass_id, st->porosity_id, st->WUDay_id, deltaLAI, VegPhenLumps, st->SnowAlb,
                            st->SnowDens, a1, a2, a3, out->DailyState);
   (void)TStar; (void)gsc; (void)zL; (void)SoilState; (void)TempVeg; (void)VPD_Pa; (void)dens_dry; (void)es_hPa;
   (void)kclear; (void)qh_residual; (void)ImpervFraction;
}

/* test hook (tests only): when non-NULL, [99][3] = (qh, L_mod, TSfc_C) of every trip of the last step */
double *so_debug_trace = 0;

/* ===== SUEWS_cal_multitsteps driver.f95:2814-3505 ===== */
static const int NCOLS_BLK[6] = {NCOL_SUEWS + 5, NCOL_SNOW + 5, NCOL_ESTM + 5, NCOL_RSL + 5, NCOL_SOLWEIG + 5,
                                 NCOL_DAILY + 5};

void so_cal_multitsteps(SoCtx *cx, const SoParams *p, SoState *st, const double *met, int len_sim, double *blkSUEWS,
                        double *blkSnow, double *blkESTM, double *blkRSL, double *blkSOL, double *blkDaily)
{
   SoForcing f;
   SoOutLines out;
   int ir, c;
#define MET(col) met[(size_t)ir + (size_t)len_sim * (col)]
#define PUT(blk, col, v) (blk)[(size_t)ir + (size_t)len_sim * (col)] = (v)
   const int own_grid = (so_met_grid == 0);
   (void)NCOLS_BLK;
   if (own_grid) { so_met_grid = met; so_met_grid_nrows = len_sim; }
   for (ir = 0; ir < len_sim; ir++) {
      f.iy = (int)MET(0);
      f.id = (int)MET(1);
      f.it = (int)MET(2);
      f.imin = (int)MET(3);
      f.isec = 0;
      f.qn1_obs = MET(4);
      f.qs_obs = MET(7);
      f.qf_obs = MET(8);
      f.avU1 = MET(9);
      f.avRh = MET(10);
      f.Temp_C = MET(11);
      f.Press_hPa = MET(12);
      f.Precip = MET(13);
      f.avkdn = MET(14);
      f.snowFrac_obs = MET(15);
      f.ldown_obs = MET(16);
      f.fcld_obs = MET(17);
      f.wu_m3 = MET(18);
      f.xsmd = MET(19);
      f.LAI_obs = MET(20);
      for (c = 0; c < 13; c++) f.Ts5mindata_ir[c] = so_ts5_block ? so_ts5_block[(size_t)ir + (size_t)len_sim * c] : -999;

      so_cal_main(cx, p, &f, st, &out);
      st->dt_since_start = st->dt_since_start + p->tstep;

      /* BLUEWS after the step (suews_ctrl_calculations.f95:194-218): QH / QE of SUEWS (columns 9, 10) or of LUMPS (11, 12),
       * UStar = column 55 */
      if (p->cbl.CBLuse >= 1 && !cx->fatal_code) {
         double line[22], qh_use = out.SUEWS[8], qe_use = out.SUEWS[9];
         if (p->cbl.QH_choice == 2) { qh_use = out.SUEWS[10]; qe_use = out.SUEWS[11]; }
         so_CBL(cx, &p->cbl, &st->cbl, &f, out.datetimeLine[4], p->tstep, qh_use, qe_use, out.SUEWS[54], cx->last_avdens,
                cx->last_lv_J_kg, cx->last_avcp, line);
         if (so_blk_BL)
            for (c = 0; c < 22; c++) so_blk_BL[(size_t)ir + (size_t)len_sim * c] = line[c];
      }

      if (blkDaily) {
         for (c = 0; c < 5; c++) PUT(blkDaily, c, out.datetimeLine[c]);
         for (c = 0; c < NCOL_DAILY; c++) PUT(blkDaily, 5 + c, out.DailyState[c]);
      }
      /* SUEWS_update_output driver.f95:2461-2509 */
      if (blkSUEWS) {
         for (c = 0; c < 5; c++) PUT(blkSUEWS, c, out.datetimeLine[c]);
         for (c = 0; c < NCOL_SUEWS; c++) PUT(blkSUEWS, 5 + c, out.SUEWS[c]);
      }
      if (blkRSL) {
         for (c = 0; c < 5; c++) PUT(blkRSL, c, out.datetimeLine[c]);
         for (c = 0; c < NCOL_RSL; c++) PUT(blkRSL, 5 + c, out.RSL[c]);
      }
      if (blkSOL) {
         for (c = 0; c < 5; c++) PUT(blkSOL, c, out.datetimeLine[c]);
         for (c = 0; c < NCOL_SOLWEIG; c++) PUT(blkSOL, 5 + c, so_set_nan(out.SOLWEIG[c]));
      }
      if (blkSnow && p->snowUse == 1) {
         for (c = 0; c < 5; c++) PUT(blkSnow, c, out.datetimeLine[c]);
         for (c = 0; c < NCOL_SNOW; c++) PUT(blkSnow, 5 + c, so_set_nan(out.Snow[c]));
      }
      if (blkESTM && p->StorageHeatMethod == 4) {
         for (c = 0; c < 5; c++) PUT(blkESTM, c, out.datetimeLine[c]);
         for (c = 0; c < NCOL_ESTM; c++) PUT(blkESTM, 5 + c, so_set_nan(out.ESTM[c]));
      }
      if (cx->fatal_code) break;
   }
   if (own_grid) { so_met_grid = 0; so_met_grid_nrows = 0; }
#undef MET
#undef PUT
}
