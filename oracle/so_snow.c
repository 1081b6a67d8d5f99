/* ORACLE (test infrastructure only -- see suews_oracle.h).
 * Restates SUEWS-SourceCode/suews_phys_snow.f95. */
#include "suews_oracle.h"
#include <math.h>

#define FMAX(a, b) ((a) > (b) ? (a) : (b))
static const double waterDens = F(999.8395); /* snow.f95:54,505 */

/* snow.f95:1411-1452 */
double so_update_snow_albedo(int tstep, const double SnowPack_prev[NSURF], double SnowAlb_prev, double Temp_C,
                             double tau_a, double tau_f, double SnowAlbMax, double SnowAlbMin)
{
   const double tau_1 = 24 * 60 * 60;
   double alb_change, SnowAlb_next, s = 0;
   int is;
   for (is = 0; is < NSURF; is++) s += SnowPack_prev[is];
   if (s > 0) {
      if (Temp_C < 0) {
         alb_change = tau_a * (tstep) / tau_1;
         SnowAlb_next = SnowAlb_prev - alb_change;
      } else {
         alb_change = exp(-tau_f * (tstep) / tau_1);
         SnowAlb_next = (SnowAlb_prev - SnowAlbMin) * alb_change + SnowAlbMin;
      }
      if (SnowAlb_next < SnowAlbMin) SnowAlb_next = SnowAlbMin;
      if (SnowAlb_next > SnowAlbMax) SnowAlb_next = SnowAlbMax;
   } else {
      SnowAlb_next = 0;
   }
   return SnowAlb_next;
}

/* snow.f95:1454-1489 */
void so_update_snow_dens(int tstep, const double SnowPack_prev[NSURF], const double SnowDens_prev[NSURF],
                         double tau_r, double SnowDensMax, double SnowDensMin, double SnowDens_next[NSURF])
{
   const double tau_1 = 24 * 60 * 60;
   double dens_change;
   int is;
   for (is = 0; is < NSURF; is++) {
      if (SnowPack_prev[is] > 0) {
         dens_change = exp(-tau_r * (tstep) / tau_1);
         SnowDens_next[is] = (SnowDens_prev[is] - SnowDensMax) * dens_change + SnowDensMax;
         if (SnowDens_next[is] > SnowDensMax) SnowDens_next[is] = SnowDensMax;
      } else {
         SnowDens_next[is] = SnowDensMin;
      }
   }
}

/* snow.f95:195-473 */
static void MeltHeat(double lvS_J_kg, double lv_J_kg, double tstep_real, double RadMeltFact, double TempMeltFact,
                     double SnowAlbMax, double SnowDensMin, double Temp_C, double Precip, double PrecipLimit,
                     double PrecipLimitAlb, double nsh_real, const double sfr[NSURF], const double Tsurf_ind[NSURF],
                     const double state_id[NSURF], const double qn1_ind_snow[NSURF], const double SnowWater[NSURF],
                     const double deltaQi[NSURF], double SnowPack[NSURF], double SnowFrac[NSURF], double *SnowAlb,
                     double SnowDens[NSURF], double *SnowfallCum, double *mwh, double *fwh, double *Qm,
                     double *QmFreez, double *QmRain, int snowCalcSwitch[NSURF], double Qm_melt[NSURF],
                     double Qm_freezState[NSURF], double Qm_rain[NSURF], double FreezMelt[NSURF],
                     double FreezState[NSURF], double FreezStateVol[NSURF], double rainOnSnow[NSURF],
                     double SnowDepth[NSURF], double mw_ind[NSURF])
{
   const double cw = 4190;
   double AdjMeltFact, Watfreeze, s;
   int is;

   *mwh = 0; *fwh = 0; *Qm = 0; *QmFreez = 0; *QmRain = 0;
   for (is = 0; is < NSURF; is++) {
      snowCalcSwitch[is] = 0; Qm_melt[is] = 0; Qm_freezState[is] = 0; Qm_rain[is] = 0; FreezMelt[is] = 0;
      FreezState[is] = 0; FreezStateVol[is] = 0; rainOnSnow[is] = 0; SnowDepth[is] = 0; mw_ind[is] = 0;
   }

   for (is = 0; is < NSURF; is++) {
      if (sfr[is] != 0) {
         if (SnowPack[is] > 0) {
            SnowDepth[is] = (SnowPack[is] / 1000) * waterDens / SnowDens[is];
            if (Temp_C >= 0) {
               if (qn1_ind_snow[is] < 0) {
                  mw_ind[is] = TempMeltFact * Temp_C;
               } else {
                  mw_ind[is] = RadMeltFact * (qn1_ind_snow[is]);
               }
            } else {
               AdjMeltFact = 1;
               mw_ind[is] = TempMeltFact * Temp_C * AdjMeltFact;
            }
            mw_ind[is] = mw_ind[is] / nsh_real;
            if (mw_ind[is] > SnowPack[is]) mw_ind[is] = SnowPack[is];
            Qm_melt[is] = waterDens * ((mw_ind[is] / tstep_real) / 1000) * (lvS_J_kg - lv_J_kg);
            if (mw_ind[is] < 0) {
               FreezMelt[is] = -mw_ind[is];
               mw_ind[is] = 0;
               if (FreezMelt[is] > SnowWater[is]) FreezMelt[is] = SnowWater[is];
               Qm_melt[is] = waterDens * ((-FreezMelt[is] / tstep_real) / 1000) * (lvS_J_kg - lv_J_kg);
            }
            if (Temp_C >= PrecipLimit && Precip > 0) {
               Qm_rain[is] = waterDens * cw * (Temp_C - PrecipLimit) * (Precip * F(0.001) / tstep_real);
               if (Qm_rain[is] < 0) {
                  Qm_rain[is] = 0;
               } else {
                  rainOnSnow[is] = Precip;
               }
            }
         }

         if (Tsurf_ind[is] < 0 && state_id[is] > 0) {
            snowCalcSwitch[is] = 1;
            if (is != WATER) {
               FreezState[is] = -TempMeltFact * Tsurf_ind[is] / nsh_real;
               if (FreezState[is] > state_id[is]) FreezState[is] = state_id[is];
               if (SnowPack[is] == 0 || SnowFrac[is] == 0) {
                  FreezStateVol[is] = FreezState[is];
               } else {
                  FreezStateVol[is] = FreezState[is] * (1 - SnowFrac[is]) / SnowFrac[is];
               }
               if (FreezStateVol[is] < F(0.00000000001) && FreezState[is] < state_id[is]) {
                  FreezState[is] = 0;
                  FreezStateVol[is] = 0;
               }
               Qm_freezState[is] = -waterDens * (FreezState[is] / tstep_real / 1000) * (lvS_J_kg - lv_J_kg);
            } else {
               Watfreeze = 100 * (0 - Temp_C) / (waterDens * (lvS_J_kg - lv_J_kg));
               FreezState[is] = Watfreeze;
               Qm_freezState[is] = -waterDens * (Watfreeze / tstep_real / 1000) * (lvS_J_kg - lv_J_kg);
            }
         }

         if (is != WATER) {
            if (SnowPack[is] > 0 || (Precip > 0 && Tsurf_ind[is] < 0)) snowCalcSwitch[is] = 1;
         } else {
            if (SnowPack[WATER] > 0 || FreezState[WATER] > 0) snowCalcSwitch[WATER] = 1;
         }

         if (Precip > 0 && Tsurf_ind[is] < 0 && SnowPack[is] > 0) {
            SnowDens[is] = SnowDens[is] * SnowPack[is] / (SnowPack[is] + Precip) +
                           SnowDensMin * Precip / (SnowPack[is] + Precip);
         }

         *mwh = *mwh + mw_ind[is] * sfr[is] * SnowFrac[is];
         *fwh = *fwh + FreezMelt[is] * sfr[is] * SnowFrac[is];
         *Qm = *Qm + Qm_melt[is] * sfr[is] * SnowFrac[is];
         *QmRain = *QmRain + Qm_rain[is] * sfr[is] * SnowFrac[is];
         *QmFreez =
            *QmFreez + deltaQi[is] * sfr[is] * SnowFrac[is] + Qm_freezState[is] * sfr[is] * (1 - SnowFrac[is]);
      }
   }

   s = 0;
   for (is = 0; is < NSURF; is++) s += SnowPack[is];
   if (Precip > 0 && s > 0 && Temp_C < 0) {
      *SnowfallCum = *SnowfallCum + Precip;
      if (*SnowfallCum > PrecipLimitAlb) {
         *SnowAlb = SnowAlbMax;
         *SnowfallCum = 0;
      }
   } else {
      *SnowfallCum = 0;
   }
}

/* snow.f95:33-193 */
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
                          double dataOutLineSnow[NCOL_SNOW])
{
   double SnowAlb, SnowfallCum, SnowPack[NSURF], SnowFrac[NSURF], SnowDens[NSURF], fwh;
   int is, k;
   (void)cx;

   SnowAlb = SnowAlb_in;
   SnowfallCum = SnowfallCum_in;
   for (is = 0; is < NSURF; is++) {
      SnowPack[is] = SnowPack_in[is];
      SnowFrac[is] = SnowFrac_in[is];
      SnowDens[is] = SnowDens_in[is];
   }

   if (snowUse == 1) {
      /* reference passes SnowFrac_in where SnowPack_prev is expected (snow.f95:137-139) */
      so_update_snow_dens(tstep, SnowFrac_in, SnowDens_in, tau_r, SnowDensMax, SnowDensMin, SnowDens);
      MeltHeat(lvS_J_kg, lv_J_kg, tstep_real, RadMeltFact, TempMeltFact, SnowAlbMax, SnowDensMin, Temp_C, Precip,
               PrecipLimit, PrecipLimitAlb, nsh_real, sfr, Tsurf_ind, state_id, qn1_ind_snow, SnowWater, deltaQi,
               SnowPack, SnowFrac, &SnowAlb, SnowDens, &SnowfallCum, mwh, &fwh, Qm, QmFreez, QmRain, snowCalcSwitch,
               Qm_melt, Qm_freezState, Qm_rain, FreezMelt, FreezState, FreezStateVol, rainOnSnow, SnowDepth, mw_ind);
   } else {
      *mwh = 0; fwh = 0; *Qm = 0; *QmFreez = 0; *QmRain = 0;
      SnowfallCum = 0;
      for (is = 0; is < NSURF; is++) {
         snowCalcSwitch[is] = 0; Qm_melt[is] = 0; Qm_freezState[is] = 0; Qm_rain[is] = 0; FreezMelt[is] = 0;
         FreezState[is] = 0; FreezStateVol[is] = 0; rainOnSnow[is] = 0; SnowDepth[is] = 0; mw_ind[is] = 0;
         SnowFrac[is] = 0;
      }
   }

   /* veg_fr_snow snow.f95:1299-1315 */
   *veg_fr = 0;
   for (is = 2; is < 7; is++) *veg_fr += sfr[is] * (1 - SnowFrac[is]);

   *SnowAlb_out = SnowAlb;
   *SnowfallCum_out = SnowfallCum;
   for (is = 0; is < NSURF; is++) {
      SnowPack_out[is] = SnowPack[is];
      SnowFrac_out[is] = SnowFrac[is];
      SnowDens_out[is] = SnowDens[is];
   }

   /* snow.f95:183-190 */
   k = 0;
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = SnowPack_out[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = mw_ind[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = Qm_melt[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = Qm_rain[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = Qm_freezState[is];
   for (is = 0; is < NSURF - 1; is++) dataOutLineSnow[k++] = SnowFrac_out[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = rainOnSnow[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = qn1_ind_snow[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = kup_ind_snow[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = FreezMelt[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = SnowWater[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = SnowDens_out[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = SnowDepth[is];
   for (is = 0; is < NSURF; is++) dataOutLineSnow[k++] = Tsurf_ind_snow[is];
   dataOutLineSnow[k++] = alb1;
}

/* snow.f95:1154-1207 */
static void Evap_SUEWS_Snow(SoCtx *cx, double Qm, double QP, double lvS_J_kg, double avdens, double avRh,
                            double Press_hPa, double *Temp_C, double RAsnow, double psyc_hPa, int tstep, double avcp,
                            double sIce_hPa, double dectime, double *ev_snow, double *tlv_sub)
{
   double e_snow, sae_snow, qe_snow, vdrcIce, esIce_hPa, EaIce_hPa, tstep_real;
   const int from = 1;
   tstep_real = (double)tstep;
   sae_snow = sIce_hPa * (QP - Qm);
   esIce_hPa = so_sat_vap_pressIce(cx, Temp_C, Press_hPa, from, dectime);
   EaIce_hPa = avRh / 100 * esIce_hPa;
   vdrcIce = (esIce_hPa - EaIce_hPa) * avdens * avcp;
   *tlv_sub = lvS_J_kg / tstep_real;
   e_snow = sae_snow + vdrcIce / RAsnow;
   qe_snow = e_snow / (sIce_hPa + psyc_hPa);
   *ev_snow = qe_snow / (*tlv_sub);
}

/* snow.f95:1210-1240 */
static void snowRem(int is, const double SnowFrac[NSURF], const double sfr[NSURF], double SnowPack[NSURF],
                    double SnowRemoval[2], double SnowLimPaved, double SnowLimBldg)
{
   if (is == PAV) {
      if (SnowPack[PAV] > SnowLimPaved) {
         SnowRemoval[PAV] = (SnowPack[PAV] - SnowLimPaved) * sfr[PAV] * SnowFrac[PAV];
         SnowPack[PAV] = SnowLimPaved;
      }
   }
   if (is == BLDG) {
      if (SnowPack[BLDG] > SnowLimBldg) {
         SnowRemoval[1] = (SnowPack[BLDG] - SnowLimBldg) * sfr[BLDG] * SnowFrac[BLDG];
         SnowPack[BLDG] = SnowLimBldg;
      }
   }
}

/* snow.f95:1243-1297 */
static double SnowDepletionCurve(int is, double swe, double sweD)
{
   double asc = 1, r;
   if (is == PAV) {
      if (swe <= sweD) {
         r = swe / sweD;
         asc = r * r;
      } else {
         asc = 1;
      }
   } else if (is == BLDG) {
      if (swe <= sweD) {
         if ((swe / sweD) < F(0.9)) {
            asc = (swe / sweD) * F(0.5);
         } else {
            r = swe / sweD;
            r = r * r; r = r * r; asc = r * r; /* **8 */
         }
      } else {
         asc = 1;
      }
   } else if (is == WATER) {
      if (swe > 0) asc = 1;
   } else {
      if (swe <= sweD) {
         /* (1/3.1416) is a REAL(4) quotient */
         asc = 1 - pow((double)(1 / 3.1416f) * acos(2 * (swe / sweD) - 1), F(1.7));
      } else {
         asc = 1;
      }
   }
   return asc;
}

/* snow.f95:477-1148.  `is` is 0-based here.  Temp_C is by reference because
 * sat_vap_pressIce may rewrite it (meteo:208-214). */
void so_SnowCalc(SoCtx *cx, int tstep, int imin, int it, double dectime, int is, int EvapMethod, double CRWmin,
                 double CRWmax, double nsh_real, double lvS_J_kg, double avdens, double avRh, double Press_hPa,
                 double Temp_C_in, double RAsnow, double psyc_hPa, double avcp, double sIce_hPa, double PervFraction,
                 double vegfraction, double addimpervious, double vpd_hPa, double qn_e, double s_hPa,
                 double ResistSurf, double RA, double rb, double tlv, double snowdensmin,
                 const double SnowProf_24hr[48], double Precip, double PipeCapacity, double RunoffToWater,
                 double addVeg, double SnowLimPaved, double SnowLimBldg, double FlowChange,
                 const double drain[NSURF], const double WetThresh[NSURF], const double stateOld[NSURF],
                 const double mw_ind[NSURF], const double SoilStoreCap[NSURF], const double rainOnSnow[NSURF],
                 const double freezMelt[NSURF], const double FreezState[NSURF], const double FreezStateVol[NSURF],
                 const double Qm_Melt[NSURF], const double Qm_rain[NSURF], const double Tsurf_ind[NSURF],
                 const double sfr[NSURF], const int DayofWeek_id[3], const double *StoreDrainPrm,
                 const double SnowPackLimit[NSURF], const double AddWater[NSURF],
                 const double AddWaterRunoff[NSURF], double soilstore_id[NSURF], double SnowPack[NSURF],
                 double SurplusEvap[NSURF], double SnowFrac[NSURF], double SnowWater[NSURF], double iceFrac[NSURF],
                 double SnowDens[NSURF], double *runoffAGimpervious, double *runoffAGveg, double *surplusWaterBody,
                 double rss_nsurf[NSURF], double runoffSnow[NSURF], double runoff[NSURF], double runoffSoil[NSURF],
                 double chang[NSURF], double changSnow[NSURF], double SnowToSurf[NSURF], double state_id[NSURF],
                 double ev_snow[NSURF], double SnowDepth[NSURF], double SnowRemoval[2], double *swe, double *ev,
                 double *chSnow_tot, double *ev_tot, double *qe_tot, double *runoff_tot, double *surf_chang_tot,
                 double *runoffPipes, double *mwstore, double *runoffwaterbody)
{
   const double IPThreshold_mmhr = 10;
   double qe = 0, MeltExcess, snowTotInit, EvPart, runoffTest, snowFracFresh1, snowFracFresh2, snowFracOld;
   double WaterHoldCapFrac, FWC, tlv_sub, capStore_is, Temp_C = Temp_C_in;
   int iu, k;
   (void)snowdensmin;

   *ev_tot = 0; *qe_tot = 0; *runoff_tot = 0; *surf_chang_tot = 0; *chSnow_tot = 0;
   for (k = 0; k < NSURF; k++) { rss_nsurf[k] = 0; SnowDepth[k] = 0; ev_snow[k] = 0; }
   tlv_sub = 0;

   iu = 1;
   if (DayofWeek_id[0] == 1 || DayofWeek_id[0] == 7) iu = 2;

   runoffSnow[is] = 0;
   runoff[is] = 0;
   runoffSoil[is] = 0;
   chang[is] = 0;
   changSnow[is] = 0;
   runoffTest = 0;
   SnowToSurf[is] = 0;
   EvPart = 0;
   *ev = 0;
   snowFracFresh1 = 0;
   snowFracFresh2 = 0;
   snowFracOld = 0;

   snowTotInit = SnowPack[is] + SnowWater[is];

   if (SnowDens[is] >= 200) {
      WaterHoldCapFrac = CRWmin;
   } else {
      WaterHoldCapFrac = CRWmin + (CRWmax - CRWmin) * (200 - SnowDens[is]) / 200;
   }

   capStore_is = SDP(StoreDrainPrm, 5, is);
   if (SnowFrac[is] < 1)
      so_cal_evap(EvapMethod, state_id[is], WetThresh[is], capStore_is, vpd_hPa, avdens, avcp, qn_e, s_hPa, psyc_hPa,
                  ResistSurf, RA, rb, tlv, &rss_nsurf[is], ev, &qe);

   if (SnowFrac[is] > 0) {
      Evap_SUEWS_Snow(cx, Qm_Melt[is], Qm_rain[is], lvS_J_kg, avdens, avRh, Press_hPa, &Temp_C, RAsnow, psyc_hPa,
                      tstep, avcp, sIce_hPa, dectime, &ev_snow[is], &tlv_sub);
   }

   if (is > 1) {
      if (PervFraction != 0) {
         EvPart = (SurplusEvap[PAV] * sfr[PAV] + SurplusEvap[BLDG] * sfr[BLDG]) / PervFraction;
      }
   }

   if (is == WATER && sfr[WATER] > 0) goto L606;

   if (SnowPack[is] > 0 && SnowFrac[is] == 1) {
      ev_snow[is] = ev_snow[is] + EvPart;
      changSnow[is] = (Precip + freezMelt[is]) - (mw_ind[is] + ev_snow[is]);
      if (rainOnSnow[is] > 0) {
         changSnow[is] = changSnow[is] - Precip;
         SnowWater[is] = SnowWater[is] + rainOnSnow[is];
      }
      SnowPack[is] = SnowPack[is] + changSnow[is];

      if (SnowPack[is] > 0) {
         SnowWater[is] = SnowWater[is] + mw_ind[is] - freezMelt[is];
         FWC = WaterHoldCapFrac * SnowPack[is];
         if (SnowWater[is] >= FWC) {
            MeltExcess = SnowWater[is] - FWC;
            SnowWater[is] = FWC;
            runoffSnow[is] = runoffSnow[is] + MeltExcess;
         }
         if (PROF24(SnowProf_24hr, it, iu - 1) == 1 && is < 2 && (imin == (nsh_real - 1) / nsh_real * 60))
            snowRem(is, SnowFrac, sfr, SnowPack, SnowRemoval, SnowLimPaved, SnowLimBldg);
      } else if (SnowPack[is] < 0) {
         SnowWater[is] = SnowWater[is] - freezMelt[is] + mw_ind[is] + SnowPack[is];
         SnowPack[is] = F(0.0);
         snowFracOld = 1;
         SnowFrac[is] = 0;
         SnowDens[is] = 0;
         if (SnowWater[is] < 0) {
            ev_snow[is] = ev_snow[is] + SnowWater[is];
            if (ev_snow[is] < 0) ev_snow[is] = 0;
            changSnow[is] = changSnow[is] + SnowWater[is];
            SnowWater[is] = 0;
         } else {
            chang[is] = SnowWater[is];
            state_id[is] = state_id[is] + chang[is];
            SnowWater[is] = 0;
         }
      }
   } else if (SnowFrac[is] < 1) {
      if (SnowPack[is] > 0) {
         ev_snow[is] = ev_snow[is] + EvPart;
         changSnow[is] = (Precip + freezMelt[is] + FreezStateVol[is]) - (mw_ind[is] + ev_snow[is]);
         if (rainOnSnow[is] > 0) {
            changSnow[is] = changSnow[is] - Precip;
            SnowWater[is] = SnowWater[is] + rainOnSnow[is];
         }
         SnowPack[is] = SnowPack[is] + changSnow[is];

         if (Precip > 0 && FreezState[is] == state_id[is]) {
            snowFracFresh1 = 1;
         } else if (Precip == 0 && FreezState[is] > 0 && FreezState[is] == state_id[is]) {
            snowFracFresh1 = 1;
         } else if (FreezState[is] > 0 && FreezState[is] < state_id[is]) {
            snowFracFresh1 = F(0.95);
         }
      } else if (SnowPack[is] == 0 && Tsurf_ind[is] < 0) {
         if ((Precip > 0 && state_id[is] == 0) || (Precip == 0 && FreezState[is] == state_id[is]) ||
             (Precip > 0 && FreezState[is] == state_id[is])) {
            changSnow[is] = Precip + FreezStateVol[is];
            SnowPack[is] = SnowPack[is] + changSnow[is];
            snowFracFresh1 = 1;
            iceFrac[is] = FreezState[is] / (FreezState[is] + Precip);
            SnowDens[is] = snowdensmin;
         }
         if (FreezState[is] > 0 && FreezState[is] < state_id[is]) {
            changSnow[is] = Precip + FreezStateVol[is];
            SnowPack[is] = SnowPack[is] + changSnow[is];
            snowFracFresh2 = F(0.95);
            iceFrac[is] = 1;
            SnowDens[is] = snowdensmin;
         }
      }

      if (SnowPack[is] > 0) {
         SnowWater[is] = SnowWater[is] + mw_ind[is] - freezMelt[is];
         FWC = WaterHoldCapFrac * SnowPack[is];
         if (SnowWater[is] >= FWC) {
            MeltExcess = SnowWater[is] - FWC;
            SnowWater[is] = FWC;
            if ((SnowFrac[is] > F(0.9) && is != BLDG) || (is == BLDG)) {
               runoffSnow[is] = runoffSnow[is] + MeltExcess;
            } else {
               SnowToSurf[is] = SnowToSurf[is] + MeltExcess * SnowFrac[is] / (1 - SnowFrac[is]);
            }
         }
         if (PROF24(SnowProf_24hr, it, iu - 1) == 1 && is < 2 && (imin == (nsh_real - 1) / nsh_real * 60))
            snowRem(is, SnowFrac, sfr, SnowPack, SnowRemoval, SnowLimPaved, SnowLimBldg);
      } else if (SnowPack[is] < 0) {
         SnowWater[is] = SnowWater[is] - freezMelt[is] + mw_ind[is] + SnowPack[is];
         SnowPack[is] = F(0.0);
         snowFracFresh1 = 0;
         snowFracFresh2 = 0;
         SnowDens[is] = 0;
         if (SnowWater[is] < 0) {
            ev_snow[is] = ev_snow[is] + SnowWater[is];
            if (ev_snow[is] < 0) ev_snow[is] = 0;
            changSnow[is] = changSnow[is] + SnowWater[is];
            SnowWater[is] = 0;
         } else {
            SnowToSurf[is] = SnowToSurf[is] + SnowWater[is] * SnowFrac[is] / (1 - SnowFrac[is]);
            SnowWater[is] = 0;
         }
      }

      if ((is == PAV || is == BLDG) && SnowFrac[is] < 1) {
         if (Precip > IPThreshold_mmhr / nsh_real) {
            runoff[is] = runoff[is] + (Precip + SnowToSurf[is] + AddWater[is] - IPThreshold_mmhr / nsh_real);
            chang[is] = IPThreshold_mmhr / nsh_real - (drain[is] + *ev + FreezState[is]);
         } else {
            chang[is] = Precip + SnowToSurf[is] + AddWater[is] - (drain[is] + *ev + FreezState[is]);
         }
         state_id[is] = state_id[is] + chang[is];
         if (is == PAV && sfr[PAV] > 0) state_id[is] = state_id[is] + (addimpervious) / sfr[PAV];
         runoff[is] = runoff[is] + drain[is] * AddWaterRunoff[is];
         if (state_id[is] < F(0.0)) {
            SurplusEvap[is] = fabs(state_id[is]);
            *ev = *ev - SurplusEvap[is];
            state_id[is] = F(0.0);
         }
      } else if (is >= 2 && SnowFrac[is] < 1) {
         *ev = *ev + EvPart;
         if (vegfraction > 0) {
            if (Precip + addVeg * (sfr[is] / vegfraction) > (IPThreshold_mmhr / nsh_real)) {
               runoff[is] = runoff[is] + (Precip + addVeg * (sfr[is] / vegfraction) + SnowToSurf[is] + AddWater[is] -
                                          (IPThreshold_mmhr / nsh_real));
               chang[is] = (IPThreshold_mmhr / nsh_real) - (drain[is] + *ev + FreezState[is]);
            } else {
               chang[is] = Precip + addVeg * (sfr[is] / vegfraction) + SnowToSurf[is] + AddWater[is] -
                           (drain[is] + *ev + FreezState[is]);
            }
         } else {
            chang[is] = Precip + SnowToSurf[is] + AddWater[is] - (drain[is] + *ev + FreezState[is]);
         }
         state_id[is] = state_id[is] + chang[is];

         if (Temp_C > 0) {
            soilstore_id[is] = soilstore_id[is] + drain[is] * AddWaterRunoff[is] * (1 - SnowFrac[is]);
         } else {
            runoff[is] = runoff[is] + drain[is] * AddWaterRunoff[is];
         }

         if (state_id[is] < F(0.0)) {
            if ((soilstore_id[is] + state_id[is]) >= 0 && Temp_C > 0) {
               soilstore_id[is] = soilstore_id[is] + state_id[is] * (1 - SnowFrac[is]);
               state_id[is] = F(0.0);
            } else {
               chang[is] = chang[is] + state_id[is];
               *ev = *ev + state_id[is];
               state_id[is] = F(0.0);
            }
         }

         if (soilstore_id[is] > SoilStoreCap[is]) {
            runoffTest = runoffTest + (soilstore_id[is] - SoilStoreCap[is]);
            soilstore_id[is] = SoilStoreCap[is];
         } else if (soilstore_id[is] < 0) {
            soilstore_id[is] = 0;
         }
      }
   }

   if (snowFracFresh2 > 0) {
      *surf_chang_tot =
         (state_id[is] - stateOld[is]) * sfr[is] * (1 - SnowFrac[is]) - Precip * sfr[is] * (1 - snowFracFresh2);
      *chSnow_tot = ((SnowPack[is] + SnowWater[is]) - snowTotInit) * sfr[is] * (1 - SnowFrac[is]) -
                    Precip * sfr[is] * snowFracFresh2;
   } else {
      *surf_chang_tot = (state_id[is] - stateOld[is]) * sfr[is] * (1 - SnowFrac[is]);
      *chSnow_tot = ((SnowPack[is] + SnowWater[is]) - snowTotInit) * sfr[is] * FMAX(SnowFrac[is], snowFracOld);
   }

   if (is == BLDG || is == PAV) {
      *ev_tot = *ev * sfr[is] * (1 - SnowFrac[is]) + ev_snow[is] * sfr[is] * FMAX(SnowFrac[is], snowFracOld);
      *qe_tot = ev_snow[is] * tlv_sub * sfr[is] * SnowFrac[is] + *ev * tlv * sfr[is] * (1 - SnowFrac[is]);
   } else {
      *ev_tot = *ev * sfr[is] * (1 - SnowFrac[is]) + ev_snow[is] * sfr[is] * FMAX(SnowFrac[is], snowFracOld);
      *qe_tot = ev_snow[is] * tlv_sub * sfr[is] * FMAX(SnowFrac[is], snowFracOld) +
                *ev * tlv * sfr[is] * (1 - SnowFrac[is]);
   }

   *runoffPipes = *runoffPipes + runoffSnow[is] * sfr[is] * FMAX(SnowFrac[is], snowFracOld) +
                  runoff[is] * sfr[is] * (1 - SnowFrac[is]) + runoffTest * sfr[is];
   so_updateFlood(is, runoff, sfr, PipeCapacity, RunoffToWater, runoffAGimpervious, surplusWaterBody, runoffAGveg,
                  runoffPipes);

   *runoff_tot = runoffSnow[is] * sfr[is] * FMAX(SnowFrac[is], snowFracOld) +
                 runoff[is] * sfr[is] * (1 - SnowFrac[is]) + runoffTest * sfr[is];

   if (SnowDens[is] != 0) {
      SnowDepth[is] = SnowPack[is] * waterDens / SnowDens[is];
   }

   *swe = *swe + SnowPack[is] * sfr[is] * FMAX(SnowFrac[is], snowFracOld);
   *mwstore = *mwstore + SnowWater[is] * sfr[is] * FMAX(SnowFrac[is], snowFracOld);

   if (snowFracFresh1 > 0) SnowFrac[is] = snowFracFresh1;
   if (snowFracFresh2 > 0) SnowFrac[is] = snowFracFresh2;

   /* SnowFractionChoice == 2 */
   if (SnowPack[is] > 0 && mw_ind[is] > 0) {
      SnowFrac[is] = SnowDepletionCurve(is, SnowPack[is], SnowPackLimit[is]);
      if (SnowFrac[is] < F(0.001)) SnowFrac[is] = F(0.001);
   } else if (SnowPack[is] == 0) {
      SnowFrac[is] = 0;
   }
   return;

L606:
   changSnow[WATER] = (Precip + freezMelt[WATER] + FreezState[WATER]) - (mw_ind[WATER] + ev_snow[WATER]);
   SnowPack[WATER] = SnowPack[WATER] + changSnow[WATER];
   state_id[WATER] = state_id[WATER] + FlowChange - FreezState[WATER];

   if (SnowPack[WATER] > 0) {
      SnowWater[WATER] = SnowWater[WATER] + mw_ind[WATER] - freezMelt[WATER];
      FWC = WaterHoldCapFrac * SnowPack[WATER];
      if (SnowWater[WATER] >= FWC && Temp_C >= 0) {
         state_id[WATER] = state_id[WATER] + (SnowWater[WATER] - FWC);
         SnowWater[WATER] = FWC;
      }
   } else if (SnowPack[is] < 0) {
      SnowWater[WATER] = SnowWater[WATER] - freezMelt[WATER] + mw_ind[WATER];
      state_id[WATER] = state_id[WATER] + SnowWater[WATER] + SnowPack[WATER];
      SnowPack[WATER] = 0;
      if (state_id[WATER] < 0) ev_snow[WATER] = ev_snow[WATER] + state_id[WATER];
   }

   if (state_id[WATER] > SDP(StoreDrainPrm, 4, WATER)) {
      runoff[WATER] = runoff[WATER] + (state_id[WATER] - SDP(StoreDrainPrm, 4, WATER));
      state_id[WATER] = SDP(StoreDrainPrm, 4, WATER);
      *runoffwaterbody = *runoffwaterbody + runoff[WATER] * sfr[WATER];
   } else {
      state_id[WATER] = state_id[WATER] + *surplusWaterBody;
      if (state_id[WATER] > SDP(StoreDrainPrm, 4, WATER)) {
         *runoffwaterbody = *runoffwaterbody + (state_id[WATER] - SDP(StoreDrainPrm, 4, WATER)) * sfr[WATER];
         state_id[WATER] = SDP(StoreDrainPrm, 4, WATER);
      }
   }

   *chSnow_tot = ((SnowPack[WATER] + SnowWater[WATER]) - snowTotInit) * sfr[WATER];
   *surf_chang_tot = (state_id[WATER] - stateOld[WATER]) * sfr[WATER];
   *ev_tot = *ev * sfr[WATER] + ev_snow[WATER] * sfr[WATER];
   *qe_tot = ev_snow[WATER] * tlv_sub * sfr[WATER] + *ev * tlv * sfr[WATER];
   *runoff_tot = runoff[is];

   if (SnowPack[WATER] > 0) {
      SnowFrac[WATER] = 1;
   } else {
      SnowFrac[WATER] = 0;
   }
}
