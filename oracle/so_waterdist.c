/* ORACLE (test infrastructure only -- see suews_oracle.h).
 * Restates SUEWS-SourceCode/suews_phys_waterdist.f95. */
#include "suews_oracle.h"
#include <math.h>

/* waterdist.f95:20-89 */
void so_drainage(SoCtx *cx, int is, double state_is, double StorCap, double DrainEq, double DrainCoef1,
                 double DrainCoef2, double nsh_real, double *drain_is)
{
   if (state_is < F(0.000000001)) {
      *drain_is = F(0.0);
   } else {
      if ((int)DrainEq == 1) {
         if (state_is < StorCap) {
            *drain_is = 0;
         } else {
            *drain_is = (DrainCoef1 * pow(state_is - StorCap, DrainCoef2)) / nsh_real;
         }
      } else if ((int)DrainEq == 2) {
         *drain_is = (DrainCoef1 * (exp(DrainCoef2 * state_is) - 1)) / nsh_real;
      } else if ((int)DrainEq == 3) {
         *drain_is = (DrainCoef1 * (pow(state_is, DrainCoef2))) / nsh_real;
      }
      if (*drain_is > state_is) {
         so_error_hint(cx, 61, "SUEWS_drain: drain_is > state_is for surface is ", *drain_is, state_is, is + 1);
         *drain_is = state_is;
      } else if (*drain_is < F(0.0001)) {
         *drain_is = 0;
      }
   }
}

/* waterdist.f95:405-451 */
void so_updateFlood(int is, const double runoff[NSURF], const double sfr[NSURF], double PipeCapacity,
                    double RunoffToWater, double *runoffAGimpervious, double *surplusWaterBody, double *runoffAGveg,
                    double *runoffPipes)
{
   *runoffPipes = *runoffPipes + (runoff[is] * sfr[is]);
   if (*runoffPipes > PipeCapacity) {
      if (is == PAV || is == BLDG) {
         if (sfr[WATER] > F(0.0000001)) {
            *runoffAGimpervious = *runoffAGimpervious + (*runoffPipes - PipeCapacity) * (1 - RunoffToWater);
            *surplusWaterBody = *surplusWaterBody + (*runoffPipes - PipeCapacity) * RunoffToWater;
         } else {
            *runoffAGimpervious = *runoffAGimpervious + (*runoffPipes - PipeCapacity);
         }
      } else if (is >= CONIF && is <= BSOIL) {
         if (sfr[WATER] > F(0.0000001)) {
            *runoffAGveg = *runoffAGveg + (*runoffPipes - PipeCapacity) * (1 - RunoffToWater);
            *surplusWaterBody = *surplusWaterBody + (*runoffPipes - PipeCapacity) * RunoffToWater;
         } else {
            *runoffAGveg = *runoffAGveg + (*runoffPipes - PipeCapacity);
         }
      }
      *runoffPipes = PipeCapacity;
   }
}

/* waterdist.f95:92-402 */
void so_cal_water_storage(SoCtx *cx, int is, const double sfr[NSURF], double PipeCapacity, double RunoffToWater,
                          double pin, const double WU_nsurf[NSURF], const double drain[NSURF],
                          const double AddWater[NSURF], double addImpervious, double nsh_real,
                          const double stateOld[NSURF], const double AddWaterRunoff[NSURF], double PervFraction,
                          double addVeg, const double SoilStoreCap[NSURF], double addWaterBody, double FlowChange,
                          const double StateLimit[NSURF], double *runoffAGimpervious, double *surplusWaterBody,
                          double *runoffAGveg, double *runoffPipes, double *ev, double soilstore_id[NSURF],
                          double SurplusEvap[NSURF], double *runoffWaterBody, double *p_mm, double chang[NSURF],
                          double runoff[NSURF], double state_id[NSURF])
{
   const double NotUsed = F(-55.5), IPThreshold_mmhr = 10;
   double EvPart;

   EvPart = 0;
   runoff[is] = 0;
   *p_mm = pin + WU_nsurf[is];
   *p_mm = *p_mm + AddWater[is];

   if (is == PAV || is == BLDG) {
      if (is == PAV) {
         if (sfr[PAV] != 0) *p_mm = *p_mm + addImpervious / sfr[PAV];
      }
      chang[is] = *p_mm - (drain[is] + *ev);
      if (*p_mm > IPThreshold_mmhr / nsh_real) {
         runoff[is] = runoff[is] + (*p_mm - IPThreshold_mmhr / nsh_real);
         chang[is] = IPThreshold_mmhr / nsh_real - (drain[is] + *ev);
      }
      state_id[is] = stateOld[is] + chang[is];
      if (state_id[is] < F(0.0)) {
         SurplusEvap[is] = fabs(state_id[is]);
         *ev = *ev - SurplusEvap[is];
         state_id[is] = F(0.0);
      }
      chang[is] = state_id[is] - stateOld[is];
      runoff[is] = runoff[is] + drain[is] * AddWaterRunoff[is];
   } else if (is >= CONIF && is <= BSOIL) {
      if (PervFraction != 0)
         EvPart = SurplusEvap[PAV] * (sfr[PAV] / PervFraction) + SurplusEvap[BLDG] * (sfr[BLDG] / PervFraction);
      else
         EvPart = 0.0;
      *ev = *ev + EvPart;
      if (is == GRASS || is == BSOIL) {
         if ((sfr[GRASS] + sfr[BSOIL]) != 0) *p_mm = *p_mm + addVeg / (sfr[GRASS] + sfr[BSOIL]);
      }
      chang[is] = *p_mm - (drain[is] + *ev);
      if (*p_mm > IPThreshold_mmhr / nsh_real) {
         runoff[is] = runoff[is] + (*p_mm - IPThreshold_mmhr / nsh_real);
         chang[is] = IPThreshold_mmhr / nsh_real - (drain[is] + *ev);
      }
      state_id[is] = stateOld[is] + chang[is];
      if (state_id[is] < F(0.0)) {
         if ((soilstore_id[is] + state_id[is]) >= 0) {
            soilstore_id[is] = soilstore_id[is] + state_id[is];
            state_id[is] = F(0.0);
         } else {
            *ev = *ev - fabs(state_id[is]);
            state_id[is] = F(0.0);
         }
      }
      chang[is] = state_id[is] - stateOld[is];
      soilstore_id[is] = soilstore_id[is] + drain[is] * AddWaterRunoff[is];
      if (soilstore_id[is] > SoilStoreCap[is]) {
         runoff[is] = runoff[is] + (soilstore_id[is] - SoilStoreCap[is]);
         soilstore_id[is] = SoilStoreCap[is];
      } else if (soilstore_id[is] < 0) {
         so_error_hint(cx, 62, "SUEWS_store: soilstore_id(is) < 0 ", soilstore_id[is], NotUsed, is + 1);
      }
   } else if (is == WATER) {
      if (sfr[WATER] != 0) {
         *p_mm = *p_mm + addWaterBody / sfr[WATER];
         chang[is] = *p_mm + FlowChange / nsh_real - *ev;
         state_id[is] = stateOld[is] + chang[is];
         if (state_id[is] < F(0.0)) {
            *ev = *ev - fabs(state_id[is]);
            state_id[is] = F(0.0);
         }
         chang[is] = state_id[is] - stateOld[is];
         if (state_id[WATER] > StateLimit[WATER]) {
            runoff[WATER] = runoff[WATER] + (state_id[WATER] - StateLimit[WATER]);
            state_id[WATER] = StateLimit[WATER];
            *runoffWaterBody = *runoffWaterBody + runoff[WATER] * sfr[WATER];
         } else {
            state_id[WATER] = state_id[WATER] + *surplusWaterBody;
            if (state_id[WATER] > StateLimit[WATER]) {
               *runoffWaterBody = *runoffWaterBody + (state_id[WATER] - StateLimit[WATER]) * sfr[WATER];
               state_id[WATER] = StateLimit[WATER];
            }
         }
         chang[is] = state_id[is] - stateOld[is];
      }
   }

   if (is < WATER) {
      so_updateFlood(is, runoff, sfr, PipeCapacity, RunoffToWater, runoffAGimpervious, surplusWaterBody, runoffAGveg,
                     runoffPipes);
   }
}

/* waterdist.f95:454-504 */
void so_ReDistributeWater(int snowUse, const double *WaterDist, const double sfr[NSURF], const double Drain[NSURF],
                          double AddWaterRunoff[NSURF], double AddWater[NSURF])
{
   int ii, jj;
   for (ii = 0; ii < NSURF - 1; ii++) AddWaterRunoff[ii] = WDIST(WaterDist, 7, ii);
   AddWaterRunoff[WATER] = 0;
   for (ii = 0; ii < NSURF; ii++) AddWater[ii] = 0;
   for (ii = 0; ii < NSURF; ii++) {
      for (jj = 0; jj < NSURF - 1; jj++) {
         if (sfr[ii] != 0) {
            if (snowUse == 0) {
               AddWater[ii] = AddWater[ii] + (Drain[jj] * sfr[jj] / sfr[ii]) * WDIST(WaterDist, ii, jj);
            } else {
               AddWaterRunoff[jj] = AddWaterRunoff[jj] + WDIST(WaterDist, ii, jj);
            }
         } else {
            AddWaterRunoff[jj] = AddWaterRunoff[jj] + WDIST(WaterDist, ii, jj);
         }
      }
   }
}

/* waterdist.f95:507-550 */
void so_update_SoilMoist(double NonWaterFraction, const double SoilStoreCap[NSURF], const double sfr[NSURF],
                         const double soilstore_id[NSURF], double *SoilMoistCap, double *SoilState, double *vsmd,
                         double *smd)
{
   int is;
   *SoilMoistCap = 0;
   *SoilState = 0;
   if (NonWaterFraction != 0) {
      for (is = 0; is < NSURF - 1; is++) {
         *SoilMoistCap = *SoilMoistCap + (SoilStoreCap[is] * sfr[is] / NonWaterFraction);
         *SoilState = *SoilState + (soilstore_id[is] * sfr[is] / NonWaterFraction);
      }
   }
   *smd = *SoilMoistCap - *SoilState;
   *vsmd = 0;
   for (is = CONIF; is <= GRASS; is++) {
      if (sfr[CONIF] + sfr[DECID] + sfr[GRASS] == 0) {
         *vsmd = 0;
      } else {
         *vsmd = *vsmd + (SoilStoreCap[is] - soilstore_id[is]) * sfr[is] / (sfr[CONIF] + sfr[DECID] + sfr[GRASS]);
      }
   }
}

/* waterdist.f95:553-619 */
void so_cal_SoilState(SoCtx *cx, int SMDMethod, double xsmd, double NonWaterFraction, double SoilMoistCap,
                      const double SoilStoreCap[NSURF], double surf_chang_per_tstep, const double soilstore_id[NSURF],
                      const double soilstoreOld[NSURF], const double sfr[NSURF], double *smd, double smd_nsurf[NSURF],
                      double *tot_chang_per_tstep, double *SoilState)
{
   const double NotUsed = -999, NAN_ = -999;
   int is;
   *SoilState = 0;
   if (NonWaterFraction != 0) {
      for (is = 0; is < NSURF - 1; is++) {
         *SoilState = *SoilState + (soilstore_id[is] * sfr[is] / NonWaterFraction);
         if (*SoilState < 0) {
            so_error_hint(cx, 62, "SUEWS_Calculations: total SoilState < 0 (just added surface is) ", *SoilState,
                          NotUsed, is + 1);
         } else if (*SoilState > SoilMoistCap) {
            so_error_hint(cx, 62, "SUEWS_Calculations: total SoilState > capacity (just added surface is) ", *SoilState,
                          NotUsed, is + 1);
         }
      }
   }
   *smd = SoilMoistCap - *SoilState;
   for (is = 0; is < NSURF; is++) smd_nsurf[is] = SoilStoreCap[is] - soilstore_id[is];
   *tot_chang_per_tstep = surf_chang_per_tstep;
   for (is = 0; is < NSURF - 1; is++) {
      *tot_chang_per_tstep = *tot_chang_per_tstep + ((soilstore_id[is] - soilstoreOld[is]) * sfr[is]);
   }
   if (SMDMethod > 0) {
      for (is = 0; is < NSURF; is++) smd_nsurf[is] = NAN_;
      *smd = xsmd;
   }
}

/* waterdist.f95:622-855 (van Genuchten 1980) */
static void vg_store(double soilstore, double SoilStoreCap, double SoilDepth, double Ksat, double *MatPot, double *Km)
{
   const double alphavG = F(0.0005);
   double SoilMoistCap_Vol, SoilMoist_vol, B_r, D, t;
   SoilMoistCap_Vol = SoilStoreCap / SoilDepth;
   SoilMoist_vol = soilstore / SoilDepth;
   B_r = F(0.1);
   if (B_r >= SoilMoist_vol) {
      *MatPot = 100000;
      *Km = 0;
   } else {
      D = (SoilMoist_vol - B_r) / (SoilMoistCap_Vol - B_r);
      if (D > F(0.99999)) D = D - F(0.0001);
      if (D < F(0.00000005)) D = D + F(0.0000001);
      *MatPot = sqrt(1 / (D * D) - 1) / alphavG;
      t = (1 - pow(1 - D * D, F(0.5)));
      *Km = Ksat * sqrt(D) * (t * t);
      if (*MatPot > 100000) {
         *MatPot = 100000;
         *Km = 0;
      }
   }
}

void so_cal_HorizontalSoilWater(const double sfr[NSURF], const double SoilStoreCap[NSURF],
                                const double SoilDepth[NSURF], const double SatHydraulicConduct[NSURF],
                                double SurfaceArea, double NonWaterFraction, double tstep_real,
                                double soilstore_id[NSURF], double runoffSoil[NSURF], double *runoffSoil_per_tstep)
{
   const double NUnits = 1;
   double MatPot1, Km1, MatPot2, Km2, Distance, KmWeight, dI, dI_dt;
   int jj, is;

   *runoffSoil_per_tstep = 0;
   for (is = 0; is < NSURF - 1; is++) {
      if (sfr[is] != 0 && SoilStoreCap[is] > 0) {
         for (jj = is + 1; jj < NSURF - 1; jj++) {
            if (sfr[jj] != 0 && SoilStoreCap[jj] > 0) {
               vg_store(soilstore_id[is], SoilStoreCap[is], SoilDepth[is], SatHydraulicConduct[is], &MatPot1, &Km1);
               vg_store(soilstore_id[jj], SoilStoreCap[jj], SoilDepth[jj], SatHydraulicConduct[jj], &MatPot2, &Km2);

               Distance = (sqrt(sfr[is] * SurfaceArea / NUnits) + sqrt(sfr[jj] * SurfaceArea / NUnits)) / 2;
               KmWeight = (sfr[is] * Km1 + sfr[jj] * Km2) / (sfr[is] + sfr[jj]);
               dI_dt = -(KmWeight) * (-MatPot1 + MatPot2) / (Distance * 1000);
               dI = dI_dt * tstep_real;

               if ((soilstore_id[jj] >= dI * sfr[is] / sfr[jj]) && ((soilstore_id[is] + dI) >= 0)) {
                  soilstore_id[is] = soilstore_id[is] + dI;
                  soilstore_id[jj] = soilstore_id[jj] - dI * sfr[is] / sfr[jj];
               } else if ((soilstore_id[is] + dI) < 0) {
                  soilstore_id[jj] = soilstore_id[jj] + soilstore_id[is] * sfr[is] / sfr[jj];
                  soilstore_id[is] = 0;
               } else {
                  soilstore_id[is] = soilstore_id[is] + soilstore_id[jj] * sfr[jj] / sfr[is];
                  soilstore_id[jj] = 0;
               }

               if (soilstore_id[is] > SoilStoreCap[is]) {
                  runoffSoil[is] = runoffSoil[is] + (soilstore_id[is] - SoilStoreCap[is]);
                  soilstore_id[is] = SoilStoreCap[is];
               }
               if (soilstore_id[jj] > SoilStoreCap[jj]) {
                  runoffSoil[jj] = runoffSoil[jj] + (soilstore_id[jj] - SoilStoreCap[jj]);
                  soilstore_id[jj] = SoilStoreCap[jj];
               }
            }
         }
         *runoffSoil_per_tstep = *runoffSoil_per_tstep + (runoffSoil[is] * sfr[is] / NonWaterFraction);
      }
   }
}

/* waterdist.f95:858-1124 */
void so_cal_WaterUse(SoCtx *cx, double nsh_real, double wu_m3, double SurfaceArea, const double sfr[NSURF],
                     double IrrFracPaved, double IrrFracBldgs, double IrrFracEveTr, double IrrFracDecTr,
                     double IrrFracGrass, double IrrFracBSoil, double IrrFracWater, const int DayofWeek_id[3],
                     const double WUProfA_24hr[48], const double WUProfM_24hr[48], double InternalWaterUse_h,
                     const double HDD_id[12], const double WUDay_id[9], int WaterUseMethod, int NSH, int it,
                     int imin, int DLS, double wu_nsurf[NSURF], double *wu_int, double *wu_ext)
{
   const double NAN_ = F(-999.);
   double WUDay_A_id[NSURF], WUDay_M_id[NSURF], IrrFrac[NSURF], WUArea[NSURF];
   double WUAreaTotal_m2, InternalWaterUse, flag_WuM, wu, OverUse, rain_cum_daily, WUProfA_tstep, WUProfM_tstep;
   int ih, iu, tstep, is;
   (void)cx;

   OverUse = 0;
   wu = 0;
   tstep = (int)(3600 / NSH);
   rain_cum_daily = HDD_id[10];

   IrrFrac[0] = IrrFracPaved; IrrFrac[1] = IrrFracBldgs; IrrFrac[2] = IrrFracEveTr; IrrFrac[3] = IrrFracDecTr;
   IrrFrac[4] = IrrFracGrass; IrrFrac[5] = IrrFracBSoil; IrrFrac[6] = IrrFracWater;

   if (WaterUseMethod == 1) {
      WUAreaTotal_m2 = 0;
      for (is = 0; is < NSURF; is++) {
         WUArea[is] = IrrFrac[is] * sfr[is] * SurfaceArea;
         WUAreaTotal_m2 += WUArea[is];
      }
      for (is = 0; is < NSURF; is++) wu_nsurf[is] = 0;
      if (wu_m3 == NAN_ || wu_m3 == 0) {
         wu = 0;
      } else {
         if (WUAreaTotal_m2 > 0) {
            wu = (wu_m3 / WUAreaTotal_m2 * 1000);
            for (is = 0; is < NSURF; is++) wu_nsurf[is] = wu * IrrFrac[is];
            wu = (wu_m3 / SurfaceArea * 1000);
         }
      }
   } else if (WaterUseMethod == 0) {
      ih = it - DLS;
      if (ih < 0) ih = 23;
      iu = 1;
      if (DayofWeek_id[0] == 1 || DayofWeek_id[0] == 7) iu = 2;

      for (is = 0; is < NSURF; is++) { WUDay_A_id[is] = 0; WUDay_M_id[is] = 0; }
      WUDay_A_id[CONIF] = WUDay_id[1];
      WUDay_A_id[DECID] = WUDay_id[4];
      WUDay_A_id[GRASS] = WUDay_id[7];
      WUDay_M_id[CONIF] = WUDay_id[2];
      WUDay_M_id[DECID] = WUDay_id[5];
      WUDay_M_id[GRASS] = WUDay_id[8];

      WUProfA_tstep = so_get_Prof_SpecTime_sum(ih, imin, 0, &WUProfA_24hr[24 * (iu - 1)], tstep);
      flag_WuM = 1;
      if (rain_cum_daily > 2) flag_WuM = 0;
      WUProfM_tstep = so_get_Prof_SpecTime_sum(ih, imin, 0, &WUProfM_24hr[24 * (iu - 1)], tstep);

      for (is = 0; is < NSURF; is++)
         wu_nsurf[is] = WUProfA_tstep * WUDay_A_id[is] + WUProfM_tstep * WUDay_M_id[is] * flag_WuM;
      for (is = 0; is < NSURF; is++) wu_nsurf[is] = wu_nsurf[is] * IrrFrac[is];
      wu = 0;
      for (is = 0; is < NSURF; is++) wu += wu_nsurf[is] * sfr[is];
   }

   InternalWaterUse = InternalWaterUse_h / nsh_real;
   *wu_ext = wu - (InternalWaterUse + OverUse);
   if (*wu_ext < 0) {
      OverUse = fabs(*wu_ext);
      *wu_ext = 0;
   } else {
      OverUse = 0;
   }
   *wu_int = wu - *wu_ext;

   if (*wu_ext != 0 && wu != 0) {
      for (is = 0; is < NSURF; is++) wu_nsurf[is] = wu_nsurf[is] * (*wu_ext) / wu;
   }
}
