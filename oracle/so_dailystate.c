/* ORACLE (test infrastructure only -- see suews_oracle.h).
 * Restates SUEWS-SourceCode/suews_phys_dailystate.f95. */
#include "suews_oracle.h"
#include <math.h>

#define FMAX(a, b) ((a) > (b) ? (a) : (b))
#define FMIN(a, b) ((a) < (b) ? (a) : (b))

/* dailystate.f95:994-1015: dummy is DIMENSION(6) of the 12-vector */
static void update_DailyState_Start(int it, int imin, double HDD_id[12])
{
   double HDD_id_mav = HDD_id[3], HDD_id_daysSR = HDD_id[5];
   int i;
   if (it == 0 && imin == 0) {
      for (i = 0; i < 6; i++) HDD_id[i] = 0;
      HDD_id[3] = HDD_id_mav;
      HDD_id[5] = HDD_id_daysSR;
   }
}

/* dailystate.f95:496-575 */
static void update_DailyState_Day(SoCtx *cx, int BaseTMethod, const int DayofWeek_id[3], double avkdn, double Temp_C,
                                  double Precip, double BaseT_HC, const double BaseT_Heating[2],
                                  const double BaseT_Cooling[2], double nsh_real, double *Tmin_id, double *Tmax_id,
                                  double *lenDay_id, double HDD_id[12])
{
   int iu = 1;
   double dT_heating, dT_cooling, BaseT_Heating_use = 0, BaseT_Cooling_use = 0;
   if (DayofWeek_id[0] == 1 || DayofWeek_id[0] == 7) iu = 2;
   switch (BaseTMethod) {
   case 1:
      BaseT_Heating_use = BaseT_HC;
      BaseT_Cooling_use = BaseT_HC;
      break;
   case 2:
      BaseT_Heating_use = BaseT_Heating[iu - 1];
      BaseT_Cooling_use = BaseT_Cooling[iu - 1];
      break;
   default:
      so_error_hint(cx, 75, "RunControl.nml", -999, -999, -999);
   }
   *Tmin_id = FMIN(Temp_C, *Tmin_id);
   *Tmax_id = FMAX(Temp_C, *Tmax_id);
   if (avkdn > 10) *lenDay_id = *lenDay_id + 1 / nsh_real;
   dT_heating = BaseT_Heating_use - Temp_C;
   dT_cooling = Temp_C - BaseT_Cooling_use;
   HDD_id[0] = HDD_id[0] + (dT_heating >= 0 ? dT_heating : 0.0);
   HDD_id[1] = HDD_id[1] + (dT_cooling >= 0 ? dT_cooling : 0.0);
   HDD_id[2] = HDD_id[2] + Temp_C;
   HDD_id[4] = HDD_id[4] + Precip;
}

/* dailystate.f95:957-992 */
static void update_HDD(int dt_since_start, int it, int imin, int tstep, double HDD_id[12])
{
   int days_prev, i, fl;
   double tstepcount;
   /* integer quotient, then *1. (REAL(4)) */
   tstepcount = (double)((float)((it * 60 + imin) * 60 / tstep) * 1.f);
   HDD_id[0] = HDD_id[0] / tstepcount;
   HDD_id[1] = HDD_id[1] / tstepcount;
   HDD_id[2] = HDD_id[2] / tstepcount;
   fl = (int)floorf((float)(dt_since_start / (24 * 60 * 60)) * 1.f);
   days_prev = (4 < fl) ? 4 : fl;
   HDD_id[3] = (HDD_id[3] * days_prev + HDD_id[2]) / (days_prev + 1);
   if (HDD_id[4] > 0) {
      HDD_id[5] = 0;
   } else {
      HDD_id[5] = HDD_id[5] + 1;
   }
   for (i = 0; i < 6; i++) HDD_id[6 + i] = HDD_id[i];
}

/* dailystate.f95:854-954 */
static void update_WaterUse(int id, int WaterUseMethod, const int DayofWeek_id[3], double lat, double FrIrriAuto,
                            const double HDD_id[12], const double state_id[NSURF], const double soilstore_id[NSURF],
                            const double SoilStoreCap[NSURF], double H_maintain, const double Ie_a[3],
                            const double Ie_m[3], int Ie_start, int Ie_end, const double DayWatPer[7],
                            const double DayWat[7], double WUDay_id[9])
{
   double h_need[3], store_total[3], WUDay_P[3], WUDay_A[3], WUDay_M[3], WUDay_total[3], temp_avg, days_since_rain,
      a, m;
   int wd, calc, i;
   temp_avg = HDD_id[8];
   days_since_rain = HDD_id[11];
   for (i = 0; i < 9; i++) WUDay_id[i] = 0;
   if (WaterUseMethod == 0) {
      wd = DayofWeek_id[0];
      if (DayWat[wd - 1] == F(1.0)) {
         calc = 0;
         if (lat >= 0) {
            if (id >= Ie_start - 1 && id <= Ie_end + 1) calc = 1;
         } else {
            calc = 1;
            if (id >= Ie_end && id <= Ie_start) calc = 0;
         }
         if (calc == 1) {
            a = FrIrriAuto * (Ie_a[0] + Ie_a[1] * temp_avg + Ie_a[2] * days_since_rain) * DayWatPer[wd - 1];
            m = (1 - FrIrriAuto) * (Ie_m[0] + Ie_m[1] * temp_avg + Ie_m[2] * days_since_rain) * DayWatPer[wd - 1];
            for (i = 0; i < 3; i++) {
               h_need[i] = SoilStoreCap[2 + i] + H_maintain;
               store_total[i] = state_id[2 + i] + soilstore_id[2 + i];
               WUDay_P[i] = h_need[i] - store_total[i];
               WUDay_P[i] = WUDay_P[i] > 0 ? WUDay_P[i] : 0.0;
               WUDay_A[i] = a;
               WUDay_A[i] = WUDay_A[i] > 0 ? WUDay_A[i] : 0.0;
               WUDay_A[i] = WUDay_A[i] + WUDay_P[i];
               WUDay_M[i] = m;
               WUDay_M[i] = WUDay_M[i] > 0 ? WUDay_M[i] : 0.0;
               WUDay_total[i] = WUDay_P[i] + WUDay_A[i] + WUDay_M[i];
               WUDay_id[i * 3 + 0] = WUDay_total[i];
               WUDay_id[i * 3 + 1] = WUDay_A[i];
               WUDay_id[i * 3 + 2] = WUDay_M[i];
            }
         } else {
            for (i = 0; i < 9; i++) WUDay_id[i] = 0;
         }
      }
   }
}

/* dailystate.f95:684-851 */
static void update_GDDLAI(int id, int LAICalcYes, double lat, double LAI_obs, double Tmin_id_prev, double Tmax_id_prev,
                          double lenDay_id_prev, const double BaseT[3], const double BaseTe[3],
                          const double GDDFull[3], const double SDDFull[3], const double LAIMin[3],
                          const double LAIMax[3], const double *LAIPower, const int LAIType[3],
                          const double LAI_id_prev[3], double GDD_id[3], double SDD_id[3], double LAI_id_next[3])
{
   double no, yes, indHelp, GDD_id_prev[3], SDD_id_prev[3];
   const int critDays = 50;
   int iv;
   for (iv = 0; iv < 3; iv++) { GDD_id_prev[iv] = GDD_id[iv]; SDD_id_prev[iv] = SDD_id[iv]; }

   for (iv = 0; iv < NVEG; iv++) {
      yes = ((Tmin_id_prev + Tmax_id_prev) / 2 - BaseT[iv]);
      no = ((Tmin_id_prev + Tmax_id_prev) / 2 - BaseTe[iv]);
      indHelp = 0;
      if (yes < 0) {
         indHelp = yes;
         yes = 0;
      }
      if (no > 0) no = 0;

      GDD_id[iv] = GDD_id_prev[iv] + yes;
      SDD_id[iv] = SDD_id_prev[iv] + no;

      if (SDD_id[iv] <= SDDFull[iv] && indHelp < 0) GDD_id[iv] = 0;

      if (GDD_id[iv] >= GDDFull[iv]) {
         GDD_id[iv] = GDDFull[iv];
         if (SDD_id[iv] < -critDays) GDD_id[iv] = 0;
      }
      if (SDD_id[iv] <= SDDFull[iv]) {
         SDD_id[iv] = SDDFull[iv];
         if (GDD_id[iv] > critDays) SDD_id[iv] = 0;
      }

      if (SDD_id[iv] < -critDays && SDD_id[iv] > SDDFull[iv]) GDD_id[iv] = 0;
      if (GDD_id[iv] > critDays && GDD_id[iv] < GDDFull[iv]) SDD_id[iv] = 0;

      if (lat >= 0) {
         if (id == 140 && SDD_id[iv] != 0) SDD_id[iv] = 0;
         if (GDD_id[iv] > critDays && id < 170) SDD_id[iv] = 0;
         if (SDD_id[iv] < -critDays && id > 170) GDD_id[iv] = 0;

         if (LAIType[iv] < F(0.5)) {
            if (GDD_id[iv] > 0 && GDD_id[iv] < GDDFull[iv]) {
               LAI_id_next[iv] = (pow(LAI_id_prev[iv], LAIPOW(LAIPower, 0, iv)) * GDD_id[iv] * LAIPOW(LAIPower, 1, iv)) +
                                 LAI_id_prev[iv];
            } else if (SDD_id[iv] < 0 && SDD_id[iv] > SDDFull[iv]) {
               LAI_id_next[iv] = (pow(LAI_id_prev[iv], LAIPOW(LAIPower, 2, iv)) * SDD_id[iv] * LAIPOW(LAIPower, 3, iv)) +
                                 LAI_id_prev[iv];
            } else {
               LAI_id_next[iv] = LAI_id_prev[iv];
            }
         } else if (LAIType[iv] >= F(0.5)) {
            if (GDD_id[iv] > 0 && GDD_id[iv] < GDDFull[iv]) {
               LAI_id_next[iv] = (pow(LAI_id_prev[iv], LAIPOW(LAIPower, 0, iv)) * GDD_id[iv] * LAIPOW(LAIPower, 1, iv)) +
                                 LAI_id_prev[iv];
            } else if (lenDay_id_prev <= 12 && SDD_id[iv] > SDDFull[iv]) {
               LAI_id_next[iv] =
                  (LAI_id_prev[iv] * LAIPOW(LAIPower, 2, iv) * (1 - SDD_id[iv]) * LAIPOW(LAIPower, 3, iv)) +
                  LAI_id_prev[iv];
            } else {
               LAI_id_next[iv] = LAI_id_prev[iv];
            }
         }
      } else if (lat < 0) {
         if (id == 300 && SDD_id[iv] != 0) SDD_id[iv] = 0;
         if (GDD_id[iv] > critDays && id > 250) SDD_id[iv] = 0;
         if (SDD_id[iv] < -critDays && id < 250) GDD_id[iv] = 0;

         if (LAIType[iv] < F(0.5)) {
            if (GDD_id[iv] > 0 && GDD_id[iv] < GDDFull[iv]) {
               LAI_id_next[iv] = (pow(LAI_id_prev[iv], LAIPOW(LAIPower, 0, iv)) * GDD_id[iv] * LAIPOW(LAIPower, 1, iv)) +
                                 LAI_id_prev[iv];
            } else if (SDD_id[iv] < 0 && SDD_id[iv] > SDDFull[iv]) {
               LAI_id_next[iv] = (pow(LAI_id_prev[iv], LAIPOW(LAIPower, 2, iv)) * SDD_id[iv] * LAIPOW(LAIPower, 3, iv)) +
                                 LAI_id_prev[iv];
            } else {
               LAI_id_next[iv] = LAI_id_prev[iv];
            }
         } else {
            if (GDD_id[iv] > 0 && GDD_id[iv] < GDDFull[iv]) {
               LAI_id_next[iv] = (pow(LAI_id_prev[iv], LAIPOW(LAIPower, 0, iv)) * GDD_id[iv] * LAIPOW(LAIPower, 1, iv)) +
                                 LAI_id_prev[iv];
            } else if (SDD_id[iv] < 0 && SDD_id[iv] > SDDFull[iv]) {
               LAI_id_next[iv] =
                  (LAI_id_prev[iv] * LAIPOW(LAIPower, 2, iv) * (1 - SDD_id[iv]) * LAIPOW(LAIPower, 3, iv)) +
                  LAI_id_prev[iv];
            } else {
               LAI_id_next[iv] = LAI_id_prev[iv];
            }
         }
      }

      if (LAI_id_next[iv] > LAIMax[iv]) {
         LAI_id_next[iv] = LAIMax[iv];
      } else if (LAI_id_next[iv] < LAIMin[iv]) {
         LAI_id_next[iv] = LAIMin[iv];
      }
   }

   if (LAICalcYes == 0) {
      for (iv = 0; iv < 3; iv++) LAI_id_next[iv] = LAI_obs;
   }
}

/* dailystate.f95:578-681 */
static void update_Veg(const double LAImax[3], const double LAIMin[3], double AlbMax_DecTr, double AlbMax_EveTr,
                       double AlbMax_Grass, double AlbMin_DecTr, double AlbMin_EveTr, double AlbMin_Grass,
                       double CapMax_dec, double CapMin_dec, double PorMax_dec, double PorMin_dec,
                       const double LAI_id[3], const double LAI_id_prev[3], double *DecidCap_id, double *albDecTr_id,
                       double *albEveTr_id, double *albGrass_id, double *porosity_id, double *StoreDrainPrm,
                       double *deltaLAI)
{
   double albChangeDecTr = 0, albChangeEveTr = 0, albChangeGrass = 0, CapChange = 0, deltaLAIEveTr = 0,
          deltaLAIGrass = 0, porChange = 0;
   int iv;
   *deltaLAI = 0;

   iv = IVDECID;
   if ((LAI_id[iv] - LAI_id_prev[iv]) != 0) {
      *deltaLAI = (LAI_id[iv] - LAI_id_prev[iv]) / (LAImax[iv] - LAIMin[iv]);
      albChangeDecTr = (AlbMax_DecTr - AlbMin_DecTr) * (*deltaLAI);
      CapChange = (CapMin_dec - CapMax_dec) * (*deltaLAI);
      porChange = (PorMin_dec - PorMax_dec) * (*deltaLAI);
   }
   iv = IVCONIF;
   if ((LAI_id[iv] - LAI_id_prev[iv]) != 0) {
      deltaLAIEveTr = (LAI_id[iv] - LAI_id_prev[iv]) / (LAImax[iv] - LAIMin[iv]);
      albChangeEveTr = (AlbMax_EveTr - AlbMin_EveTr) * deltaLAIEveTr;
   }
   iv = IVGRASS;
   if ((LAI_id[iv] - LAI_id_prev[iv]) != 0) {
      deltaLAIGrass = (LAI_id[iv] - LAI_id_prev[iv]) / (LAImax[iv] - LAIMin[iv]);
      albChangeGrass = (AlbMax_Grass - AlbMin_Grass) * deltaLAIGrass;
   }

   *DecidCap_id = *DecidCap_id - CapChange;
   SDP(StoreDrainPrm, 5, DECID) = *DecidCap_id;
   *porosity_id = *porosity_id + porChange;
   *albDecTr_id = *albDecTr_id + albChangeDecTr;
   *albEveTr_id = *albEveTr_id + albChangeEveTr;
   *albGrass_id = *albGrass_id + albChangeGrass;
}

/* dailystate.f95:57-352.  first/last step of day come from datetime%getDay
 * comparisons (dailystate.f95:266-272) == crossing midnight. */
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
                       double GDD_id_next[NVEG], double SDD_id_next[NVEG], double *deltaLAI, double WUDay_id[9])
{
   double GDD_id[3], SDD_id[3], LAI_id[3], LAI_id_in[3], HDD_id[12], StoreDrainPrm[6 * NSURF];
   double Tmin_id, Tmax_id, lenDay_id, DecidCap_id, albDecTr_id, albEveTr_id, albGrass_id, porosity_id;
   int first_tstep_Q, last_tstep_Q, i, sod;

   for (i = 0; i < 3; i++) { LAI_id[i] = LAI_id_prev[i]; GDD_id[i] = GDD_id_prev[i]; SDD_id[i] = SDD_id_prev[i]; }
   Tmin_id = Tmin_id_prev;
   Tmax_id = Tmax_id_prev;
   lenDay_id = lenDay_id_prev;
   for (i = 0; i < 6 * NSURF; i++) StoreDrainPrm[i] = StoreDrainPrm_prev[i];
   DecidCap_id = DecidCap_id_prev;
   albDecTr_id = albDecTr_id_prev;
   albEveTr_id = albEveTr_id_prev;
   albGrass_id = albGrass_id_prev;
   porosity_id = porosity_id_prev;
   for (i = 0; i < 12; i++) HDD_id[i] = HDD_id_prev[i];

   sod = f->it * 3600 + f->imin * 60 + f->isec; /* seconds of day of time_now */
   first_tstep_Q = (sod - tstep_prev) < 0;      /* time_prev falls on the previous calendar day */
   last_tstep_Q = (sod + p->tstep) >= 86400;    /* time_next falls on the next calendar day */

   if (first_tstep_Q) {
      update_DailyState_Start(f->it, f->imin, HDD_id);
      Tmin_id = f->Temp_C;
      Tmax_id = f->Temp_C;
      lenDay_id = 0;
   }

   update_DailyState_Day(cx, p->BaseTMethod, DayofWeek_id, f->avkdn, f->Temp_C, f->Precip, p->BaseT_HC,
                         p->BaseT_Heating, p->BaseT_Cooling, nsh_real, &Tmin_id, &Tmax_id, &lenDay_id, HDD_id);

   if (last_tstep_Q) {
      /* update_DailyState_End dailystate.f95:355-493 */
      update_HDD(dt_since_start, f->it, f->imin, p->tstep, HDD_id);
      update_WaterUse(f->id, p->WaterUseMethod, DayofWeek_id, p->lat, p->Faut, HDD_id, state_id, soilstore_id,
                      p->SoilStoreCap, p->H_maintain, p->Ie_a, p->Ie_m, p->Ie_start, p->Ie_end, p->DayWatPer,
                      p->DayWat, WUDay_id);
      for (i = 0; i < 3; i++) LAI_id_in[i] = LAI_id[i];
      update_GDDLAI(f->id, p->LAICalcYes, p->lat, f->LAI_obs, Tmin_id, Tmax_id, lenDay_id, p->BaseT, p->BaseTe,
                    p->GDDFull, p->SDDFull, p->LAIMin, p->LAIMax, p->LAIPower, p->LAIType, LAI_id_in, GDD_id, SDD_id,
                    LAI_id);
      update_Veg(p->LAIMax, p->LAIMin, p->AlbMax_DecTr, p->AlbMax_EveTr, p->AlbMax_Grass, p->AlbMin_DecTr,
                 p->AlbMin_EveTr, p->AlbMin_Grass, p->CapMax_dec, p->CapMin_dec, p->PorMax_dec, p->PorMin_dec, LAI_id,
                 LAI_id_in, &DecidCap_id, &albDecTr_id, &albEveTr_id, &albGrass_id, &porosity_id, StoreDrainPrm,
                 deltaLAI);
   }

   for (i = 0; i < 3; i++) { LAI_id_next[i] = LAI_id[i]; GDD_id_next[i] = GDD_id[i]; SDD_id_next[i] = SDD_id[i]; }
   *Tmin_id_next = Tmin_id;
   *Tmax_id_next = Tmax_id;
   *lenDay_id_next = lenDay_id;
   for (i = 0; i < 6 * NSURF; i++) StoreDrainPrm_next[i] = StoreDrainPrm[i];
   *DecidCap_id_next = DecidCap_id;
   *albDecTr_id_next = albDecTr_id;
   *albEveTr_id_next = albEveTr_id;
   *albGrass_id_next = albGrass_id;
   *porosity_id_next = porosity_id;
   for (i = 0; i < 12; i++) HDD_id_next[i] = HDD_id[i];
}

/* dailystate.f95:1044-1113.  HDD_id is DIMENSION(6) in the packer. */
void so_update_DailyStateLine(int it, int imin, double nsh_real, const double GDD_id[NVEG], const double HDD_id[12],
                              const double LAI_id[NVEG], const double SDD_id[NVEG], double Tmin_id, double Tmax_id,
                              double lenday_id, double DecidCap_id, double albDecTr_id, double albEveTr_id,
                              double albGrass_id, double porosity_id, const double WUDay_id[9], double deltaLAI,
                              double VegPhenLumps, double SnowAlb, const double SnowDens[NSURF], double a1,
                              double a2, double a3, double DailyStateLine[NCOL_DAILY])
{
   int i, k = 0;
   for (i = 0; i < NCOL_DAILY; i++) DailyStateLine[i] = -999;
   if (it == 23 && imin == (nsh_real - 1) / nsh_real * 60) {
      for (i = 0; i < 6; i++) DailyStateLine[k++] = HDD_id[i];
      for (i = 0; i < 3; i++) DailyStateLine[k++] = GDD_id[i];
      for (i = 0; i < 3; i++) DailyStateLine[k++] = SDD_id[i];
      DailyStateLine[k++] = Tmin_id;
      DailyStateLine[k++] = Tmax_id;
      DailyStateLine[k++] = lenday_id;
      for (i = 0; i < 3; i++) DailyStateLine[k++] = LAI_id[i];
      DailyStateLine[k++] = DecidCap_id;
      DailyStateLine[k++] = porosity_id;
      DailyStateLine[k++] = albEveTr_id;
      DailyStateLine[k++] = albDecTr_id;
      DailyStateLine[k++] = albGrass_id;
      for (i = 0; i < 9; i++) DailyStateLine[k++] = WUDay_id[i];
      DailyStateLine[k++] = deltaLAI;
      DailyStateLine[k++] = VegPhenLumps;
      DailyStateLine[k++] = SnowAlb;
      for (i = 0; i < NSURF; i++) DailyStateLine[k++] = SnowDens[i];
      DailyStateLine[k++] = a1;
      DailyStateLine[k++] = a2;
      DailyStateLine[k++] = a3;
   }
}
