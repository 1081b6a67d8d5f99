// dev_daily.cuh -- once-a-day phenology / degree-day / irrigation state.
// Device restatement of suews_phys_dailystate.f95.
#pragma once
#include "dev_common.cuh"

// The evolving daily state touched by SUEWS_cal_DailyState (dailystate.f95:57-352)
struct SuDaily {
   double HDD_id[12], GDD_id[3], SDD_id[3], LAI_id[3], WUDay_id[9];
   double Tmin_id, Tmax_id, lenDay_id, DecidCap_id, albDecTr_id, albEveTr_id, albGrass_id, porosity_id;
   double DecidStoreCap;   // StoreDrainPrm(6, DecidSurf): the only StoreDrainPrm element that evolves (dailystate.f95:671)
   double deltaLAI;
};

// update_DailyState_End, dailystate.f95:355-493 (update_HDD :957, update_WaterUse :854, update_GDDLAI :684,
// update_Veg :578).  Runs on the last step of a day only (1 step in 288).
SU_DEVNI void update_DailyState_End(const SuewsClassParams &P, double lat, int id, int it, int imin, int dow,
                                    int dt_since_start, double LAI_obs, const double *state_id,
                                    const double *soilstore_id, SuDaily &d)
{
   // ---- update_HDD
   {
      double tstepcount = (double)((float)((it * 60 + imin) * 60 / P.tstep) * 1.f);   // integer quotient * 1. (REAL(4))
      d.HDD_id[0] = d.HDD_id[0] / tstepcount;
      d.HDD_id[1] = d.HDD_id[1] / tstepcount;
      d.HDD_id[2] = d.HDD_id[2] / tstepcount;
      int fl = (int)floorf((float)(dt_since_start / (24 * 60 * 60)) * 1.f);
      int days_prev = (4 < fl) ? 4 : fl;
      d.HDD_id[3] = (d.HDD_id[3] * days_prev + d.HDD_id[2]) / (days_prev + 1);
      if (d.HDD_id[4] > 0) d.HDD_id[5] = 0;
      else d.HDD_id[5] = d.HDD_id[5] + 1;
      for (int i = 0; i < 6; i++) d.HDD_id[6 + i] = d.HDD_id[i];
   }
   // ---- update_WaterUse
   {
      double temp_avg = d.HDD_id[8], days_since_rain = d.HDD_id[11];
      for (int i = 0; i < 9; i++) d.WUDay_id[i] = 0;
      if (P.WaterUseMethod == 0) {
         int wd = dow;
         if (P.DayWat[wd - 1] == F(1.0)) {
            int calc = 0;
            if (lat >= 0) {
               if (id >= P.Ie_start - 1 && id <= P.Ie_end + 1) calc = 1;
            } else {
               calc = 1;
               if (id >= P.Ie_end && id <= P.Ie_start) calc = 0;
            }
            if (calc == 1) {
               double a = P.Faut * (P.Ie_a[0] + P.Ie_a[1] * temp_avg + P.Ie_a[2] * days_since_rain) * P.DayWatPer[wd - 1];
               double m = (1 - P.Faut) * (P.Ie_m[0] + P.Ie_m[1] * temp_avg + P.Ie_m[2] * days_since_rain) * P.DayWatPer[wd - 1];
               a = a > 0 ? a : 0.0;
               m = m > 0 ? m : 0.0;
               for (int i = 0; i < 3; i++) {
                  double pnd = (P.SoilStoreCap[2 + i] + P.H_maintain) - (state_id[2 + i] + soilstore_id[2 + i]);
                  pnd = pnd > 0 ? pnd : 0.0;
                  double A = a + pnd;
                  d.WUDay_id[i * 3 + 0] = pnd + A + m;
                  d.WUDay_id[i * 3 + 1] = A;
                  d.WUDay_id[i * 3 + 2] = m;
               }
            }
         }
      }
   }
   // ---- update_GDDLAI
   double LAI_in[3];
   for (int iv = 0; iv < 3; iv++) LAI_in[iv] = d.LAI_id[iv];
   {
      const int critDays = 50;
      for (int iv = 0; iv < NVEG; iv++) {
         double GDD = d.GDD_id[iv], SDD = d.SDD_id[iv];
         double yes = ((d.Tmin_id + d.Tmax_id) / 2 - P.BaseT[iv]);
         double no = ((d.Tmin_id + d.Tmax_id) / 2 - P.BaseTe[iv]);
         double indHelp = 0;
         if (yes < 0) { indHelp = yes; yes = 0; }
         if (no > 0) no = 0;
         GDD = GDD + yes;
         SDD = SDD + no;
         const double GDDFull = P.GDDFull[iv], SDDFull = P.SDDFull[iv];
         if (SDD <= SDDFull && indHelp < 0) GDD = 0;
         if (GDD >= GDDFull) {
            GDD = GDDFull;
            if (SDD < -critDays) GDD = 0;
         }
         if (SDD <= SDDFull) {
            SDD = SDDFull;
            if (GDD > critDays) SDD = 0;
         }
         if (SDD < -critDays && SDD > SDDFull) GDD = 0;
         if (GDD > critDays && GDD < GDDFull) SDD = 0;
         const double Lp = LAI_in[iv];
         const double p1 = SU_LAIPOW(P.LAIPower.v, 0, iv), p2 = SU_LAIPOW(P.LAIPower.v, 1, iv),
                      p3 = SU_LAIPOW(P.LAIPower.v, 2, iv), p4 = SU_LAIPOW(P.LAIPower.v, 3, iv);
         double Ln = Lp;
         const bool typ0 = (P.LAIType[iv] < F(0.5));
         if (lat >= 0) {
            if (id == 140 && SDD != 0) SDD = 0;
            if (GDD > critDays && id < 170) SDD = 0;
            if (SDD < -critDays && id > 170) GDD = 0;
            if (GDD > 0 && GDD < GDDFull) {
               Ln = (pow(Lp, p1) * GDD * p2) + Lp;
            } else if (typ0) {
               if (SDD < 0 && SDD > SDDFull) Ln = (pow(Lp, p3) * SDD * p4) + Lp;
            } else {
               if (d.lenDay_id <= 12 && SDD > SDDFull) Ln = (Lp * p3 * (1 - SDD) * p4) + Lp;
            }
         } else {
            if (id == 300 && SDD != 0) SDD = 0;
            if (GDD > critDays && id > 250) SDD = 0;
            if (SDD < -critDays && id < 250) GDD = 0;
            if (GDD > 0 && GDD < GDDFull) {
               Ln = (pow(Lp, p1) * GDD * p2) + Lp;
            } else if (SDD < 0 && SDD > SDDFull) {
               Ln = typ0 ? (pow(Lp, p3) * SDD * p4) + Lp : (Lp * p3 * (1 - SDD) * p4) + Lp;
            }
         }
         if (Ln > P.LAIMax[iv]) Ln = P.LAIMax[iv];
         else if (Ln < P.LAIMin[iv]) Ln = P.LAIMin[iv];
         d.GDD_id[iv] = GDD;
         d.SDD_id[iv] = SDD;
         d.LAI_id[iv] = Ln;
      }
      if (P.LAICalcYes == 0)
         for (int iv = 0; iv < 3; iv++) d.LAI_id[iv] = LAI_obs;
   }
   // ---- update_Veg
   {
      double albChangeDecTr = 0, albChangeEveTr = 0, albChangeGrass = 0, CapChange = 0, porChange = 0;
      d.deltaLAI = 0;
      int iv = 1;
      if ((d.LAI_id[iv] - LAI_in[iv]) != 0) {
         d.deltaLAI = (d.LAI_id[iv] - LAI_in[iv]) / (P.LAIMax[iv] - P.LAIMin[iv]);
         albChangeDecTr = (P.AlbMax_DecTr - P.AlbMin_DecTr) * d.deltaLAI;
         CapChange = (P.CapMin_dec - P.CapMax_dec) * d.deltaLAI;
         porChange = (P.PorMin_dec - P.PorMax_dec) * d.deltaLAI;
      }
      iv = 0;
      if ((d.LAI_id[iv] - LAI_in[iv]) != 0) {
         double dl = (d.LAI_id[iv] - LAI_in[iv]) / (P.LAIMax[iv] - P.LAIMin[iv]);
         albChangeEveTr = (P.AlbMax_EveTr - P.AlbMin_EveTr) * dl;
      }
      iv = 2;
      if ((d.LAI_id[iv] - LAI_in[iv]) != 0) {
         double dl = (d.LAI_id[iv] - LAI_in[iv]) / (P.LAIMax[iv] - P.LAIMin[iv]);
         albChangeGrass = (P.AlbMax_Grass - P.AlbMin_Grass) * dl;
      }
      d.DecidCap_id = d.DecidCap_id - CapChange;
      d.DecidStoreCap = d.DecidCap_id;
      d.porosity_id = d.porosity_id + porChange;
      d.albDecTr_id = d.albDecTr_id + albChangeDecTr;
      d.albEveTr_id = d.albEveTr_id + albChangeEveTr;
      d.albGrass_id = d.albGrass_id + albChangeGrass;
   }
}

// SUEWS_cal_DailyState, dailystate.f95:57-352.  `d` enters holding the *_prev values and leaves
// holding *_next.  first/last step of a day are datetime%getDay comparisons in the reference
// (dailystate.f95:266-272) == crossing midnight.
SU_DEV void cal_DailyState(SuDiag &dg, const SuewsClassParams &P, double lat, int id, int it, int imin, int isec,
                           int dow, int tstep_prev, int dt_since_start, double nsh_real, double avkdn, double Temp_C,
                           double Precip, double LAI_obs, const double *state_id, const double *soilstore_id,
                           SuDaily &d)
{
   const int sod = it * 3600 + imin * 60 + isec;
   const bool first_tstep_Q = (sod - tstep_prev) < 0;
   const bool last_tstep_Q = (sod + P.tstep) >= 86400;

   if (first_tstep_Q) {
      // update_DailyState_Start, dailystate.f95:994-1015 (dummy is DIMENSION(6) of the 12-vector)
      if (it == 0 && imin == 0) {
         double mav = d.HDD_id[3], dsr = d.HDD_id[5];
         for (int i = 0; i < 6; i++) d.HDD_id[i] = 0;
         d.HDD_id[3] = mav;
         d.HDD_id[5] = dsr;
      }
      d.Tmin_id = Temp_C;
      d.Tmax_id = Temp_C;
      d.lenDay_id = 0;
   }
   // update_DailyState_Day, dailystate.f95:496-575
   {
      const int u = (dow == 1 || dow == 7) ? 1 : 0;
      double BaseT_Heating_use = 0, BaseT_Cooling_use = 0;
      if (P.BaseTMethod == 1) {
         BaseT_Heating_use = P.BaseT_HC;
         BaseT_Cooling_use = P.BaseT_HC;
      } else if (P.BaseTMethod == 2) {
         BaseT_Heating_use = P.BaseT_Heating[u];
         BaseT_Cooling_use = P.BaseT_Cooling[u];
      } else {
         su_fatal(dg, 75);
      }
      d.Tmin_id = SU_MIN(Temp_C, d.Tmin_id);
      d.Tmax_id = SU_MAX(Temp_C, d.Tmax_id);
      if (avkdn > 10) d.lenDay_id = d.lenDay_id + 1 / nsh_real;
      double dT_heating = BaseT_Heating_use - Temp_C;
      double dT_cooling = Temp_C - BaseT_Cooling_use;
      d.HDD_id[0] = d.HDD_id[0] + (dT_heating >= 0 ? dT_heating : 0.0);
      d.HDD_id[1] = d.HDD_id[1] + (dT_cooling >= 0 ? dT_cooling : 0.0);
      d.HDD_id[2] = d.HDD_id[2] + Temp_C;
      d.HDD_id[4] = d.HDD_id[4] + Precip;
   }
   if (last_tstep_Q) update_DailyState_End(P, lat, id, it, imin, dow, dt_since_start, LAI_obs, state_id, soilstore_id, d);
}
