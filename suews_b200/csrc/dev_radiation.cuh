// dev_radiation.cuh -- NARP net all-wave radiation and the NREL solar position algorithm.
// Device restatement of suews_phys_narp.f95.  The SPA is split the B200 way:
//   spa_time_part()      steps 1-10 (narp.f95:488-541): depend on (time, UTC offset) only -> one thread
//                        per timestep in a tiny pre-kernel, ~390 trig calls done once instead of once
//                        per cell per Tsurf iteration as the reference does (driver.f95:687 inside :642);
//   spa_location_part()  steps 11-14 (narp.f95:542-564): per cell.
// Same arithmetic, same order, so zenith/azimuth are what the reference computes.
#pragma once
#include "dev_common.cuh"
#include "spa_tables.cuh"

#define SU_PI 3.141592653589793

// narp.f95:1219-1236
SU_INL double set_to_range(double var)
{
   double vari = var - F(360.0) * (int)floor(var / F(360.0));
   if (vari < F(0.0)) vari = vari + F(360.0);
   return vari;
}

SU_DEV double spa_sum(const double *A, const double *B, const double *C, int n, double JME)
{
   double s = 0;
   for (int i = 0; i < n; i++) s += A[i] * cos(B[i] + (C[i] * JME));
   return s;
}

// util_time.f95:23-96 (month/day only; negative day-of-year falls into January as in the reference)
SU_DEV void day2month(int b, int year, int &mb, int &md)
{
   int k = ((year % 4 == 0 && year % 100 != 0) || year % 400 == 0) ? 1 : 0;   // year > 0 here
   if (b <= 31) { mb = 1; md = b; }
   else if (b <= 59 + k) { mb = 2; md = b - 31; }
   else if (b <= 90 + k) { mb = 3; md = b - (59 + k); }
   else if (b <= 120 + k) { mb = 4; md = b - (90 + k); }
   else if (b <= 151 + k) { mb = 5; md = b - (120 + k); }
   else if (b <= 181 + k) { mb = 6; md = b - (151 + k); }
   else if (b <= 212 + k) { mb = 7; md = b - (181 + k); }
   else if (b <= 243 + k) { mb = 8; md = b - (212 + k); }
   else if (b <= 273 + k) { mb = 9; md = b - (243 + k); }
   else if (b <= 304 + k) { mb = 10; md = b - (273 + k); }
   else if (b <= 334 + k) { mb = 11; md = b - (304 + k); }
   else { mb = 12; md = b - (334 + k); }
}

// util_time.f95:173-199
SU_INL int day_of_week(int date, int month, int year)
{
   int yr = year, mn = month;
   if (!(mn > 2)) { mn = mn + 12; yr = yr - 1; }
   int n1 = (26 * (mn + 1)) / 10;
   int n2 = (125 * yr) / 100;
   int day = (date + n1 + n2 - (yr / 100) + (yr / 400) - 1);
   return day % 7 + 1;
}

// NARP_cal_SunPosition steps 1-10, narp.f95:460-541 + helpers :568-1100.
// idectime is what the reference passes: dectime - tstep/2 with the INTEGER tstep/2 (driver.f95:689).
static __device__ void spa_time_part(double year, double idectime, double UTC, SuewsEphem &e)
{
   // dectime_to_timevec util_time.f95:202
   int doy = (int)floor(idectime);
   double DH = idectime - doy;
   int hour = (int)(24 * DH);
   double DM = 24 * DH - hour;
   int min = (int)(60 * DM);
   double DS = 60 * DM - min;
   double sec = (int)(60 * DS);
   int month, day;
   day2month(doy, (int)year, month, day);

   // julian_calculation narp.f95:568-629
   double Y, M, A, B;
   if (month == 1 || month == 2) { Y = year - F(1.); M = month + 12; } else { Y = year; M = month; }
   double ut_time = (((double)(float)hour - UTC) / F(24.)) + (double)((float)min / (60.f * 24.f)) +
                    (sec / (double)(60.f * 60.f * 24.f));
   double D = day + ut_time;
   if (year == F(1582.)) {
      if (month == 10) {
         if (day <= 4) B = 0;
         else if (day >= 15) { A = floor(Y / 100); B = 2 - A + floor(A / 4); }
         else B = 0;
      } else if (month < 10) B = 0;
      else { A = floor(Y / 100); B = 2 - A + floor(A / 4); }
   } else if (year < F(1582.)) {
      B = 0;
   } else {
      A = floor(Y / 100);
      B = 2 - A + floor(A / 4);
   }
   double julianday = ((int)floor(F(365.25) * (Y + F(4716.))) + (int)floor(F(30.6001) * (M + 1))) + D + B - F(1524.5);
   double jde = julianday + (F(0.) / 86400);
   double jc = (julianday - F(2451545.)) / F(36525.);
   double jce = (jde - F(2451545.)) / F(36525.);
   double jme = jce / F(10.);

   // earth_heliocentric_position_calculation narp.f95:631-828
   double L0 = spa_sum(c_spa_A0, c_spa_B0, c_spa_C0, 64, jme);
   double L1 = spa_sum(c_spa_A1, c_spa_B1, c_spa_C1, 34, jme);
   double L2 = spa_sum(c_spa_A2, c_spa_B2, c_spa_C2, 20, jme);
   double L3 = spa_sum(c_spa_A3, c_spa_B3, c_spa_C3, 7, jme);
   double L4 = spa_sum(c_spa_A4, c_spa_B4, c_spa_C4, 3, jme);
   double L5 = F(1.) * cos(F(3.1400) + (F(0.) * jme));
   double eh_lon = (L0 + (L1 * jme) + (L2 * su_p2(jme)) + (L3 * su_p3(jme)) + (L4 * su_p4(jme)) + (L5 * su_p5(jme))) / F(1e8);
   eh_lon = eh_lon * F(180.) / SU_PI;
   eh_lon = set_to_range(eh_lon);
   L0 = spa_sum(c_spa_A0i, c_spa_B0i, c_spa_C0i, 5, jme);
   L1 = spa_sum(c_spa_A1i, c_spa_B1i, c_spa_C1i, 2, jme);
   double eh_lat = (L0 + (L1 * jme)) / F(1e8);
   eh_lat = eh_lat * 180 / SU_PI;
   eh_lat = set_to_range(eh_lat);
   L0 = spa_sum(c_spa_A0j, c_spa_B0j, c_spa_C0j, 40, jme);
   L1 = spa_sum(c_spa_A1j, c_spa_B1j, c_spa_C1j, 10, jme);
   L2 = spa_sum(c_spa_A2j, c_spa_B2j, c_spa_C2j, 6, jme);
   L3 = spa_sum(c_spa_A3j, c_spa_B3j, c_spa_C3j, 2, jme);
   L4 = 4 * cos(F(2.56) + (F(6283.08000) * jme));
   double eh_rad = (L0 + (L1 * jme) + (L2 * su_p2(jme)) + (L3 * su_p3(jme)) + (L4 * su_p4(jme))) / F(1e8);

   // sun_geocentric_position_calculation narp.f95:830-840
   double sg_lon = set_to_range(eh_lon + F(180.0));
   double sg_lat = set_to_range(-eh_lat);

   // nutation_calculation narp.f95:842-961
   double X[5];
   {
      const double p0[4] = {(double)(1 / 189474.f), F(-0.0019142), F(445267.11148), F(297.85036)};
      const double p1[4] = {(double)(-(1 / 300000.f)), F(-0.0001603), F(35999.05034), F(357.52772)};
      const double p2[4] = {(double)(1 / 56250.f), F(0.0086972), F(477198.867398), F(134.96298)};
      const double p3[4] = {(double)(1 / 327270.f), F(-0.0036825), F(483202.017538), F(93.27191)};
      const double p4[4] = {(double)(1 / 450000.f), F(0.0020708), F(-1934.136261), F(125.04452)};
      double j3 = su_p3(jce), j2 = su_p2(jce);
      X[0] = p0[0] * j3 + p0[1] * j2 + p0[2] * jce + p0[3];
      X[1] = p1[0] * j3 + p1[1] * j2 + p1[2] * jce + p1[3];
      X[2] = p2[0] * j3 + p2[1] * j2 + p2[2] * jce + p2[3];
      X[3] = p3[0] * j3 + p3[1] * j2 + p3[2] * jce + p3[3];
      X[4] = p4[0] * j3 + p4[1] * j2 + p4[2] * jce + p4[3];
   }
   double sl = 0, so = 0;
   for (int i = 0; i < 63; i++) {
      const signed char *Yt = &c_spa_Y_terms1[5 * i];
      const double *N = &c_spa_nutation_terms1[4 * i];
      double arg = ((Yt[0] * X[0]) + (Yt[1] * X[1]) + (Yt[2] * X[2]) + (Yt[3] * X[3]) + (Yt[4] * X[4])) * SU_PI / 180;
      sl += ((N[0] + (N[1] * jce))) * sin(arg);
      so += ((N[2] + (N[3] * jce))) * cos(arg);
   }
   double nut_lon = sl / F(36000000.0);
   double nut_obl = so / F(36000000.0);

   // corr_obliquity_calculation narp.f95:963-983 (U**10 .. U**2 as powi chains)
   double U = jme / 10;
   double u2 = U * U, u3 = U * u2, u4 = u2 * u2, u5 = u2 * u3, u6 = u3 * u3, u7 = u3 * u4, u8 = u4 * u4, u9 = u3 * u6,
          u10 = u5 * u5;
   const double *p = c_spa_p;
   double mean_obl = p[0] * u10 + p[1] * u9 + p[2] * u8 + p[3] * u7 + p[4] * u6 + p[5] * u5 + p[6] * u4 + p[7] * u3 +
                     p[8] * u2 + p[9] * U + p[10];
   double corr_obl = (mean_obl / 3600) + nut_obl;

   double aberr = F(-20.4898) / (3600 * eh_rad);        // narp.f95:985-993
   double app_sun_lon = sg_lon + nut_lon + aberr;        // narp.f95:995-1006
   // apparent_stime_at_greenwich_calculation narp.f95:1008-1040 (true double literals; pi with 15 digits)
   double mean_stime = 280.46061837 + (360.98564736629 * (julianday - 2451545.0)) + (0.000387933 * su_p2(jc)) -
                       (su_p3(jc) / 38710000.0);
   mean_stime = set_to_range(mean_stime);
   e.app_stime = mean_stime + (nut_lon * cos(corr_obl * 3.14159265358979 / 180));

   // sun_rigth_ascension_calculation narp.f95:1042-1064
   double num = (sin(app_sun_lon * SU_PI / F(180.0)) * cos(corr_obl * SU_PI / F(180.0))) -
                (tan(sg_lat * SU_PI / F(180.0)) * sin(corr_obl * SU_PI / F(180.0)));
   double den = cos(app_sun_lon * SU_PI / F(180.0));
   e.sun_ra = set_to_range(atan2(num, den) * F(180.0) / SU_PI);
   // sun_geocentric_declination_calculation narp.f95:1066-1084
   double argument = (sin(sg_lat * SU_PI / F(180.0)) * cos(corr_obl * SU_PI / F(180.0))) +
                     (cos(sg_lat * SU_PI / F(180.0)) * sin(corr_obl * SU_PI / 180) * sin(app_sun_lon * SU_PI / F(180.0)));
   e.sun_dec = asin(argument) * F(180.0) / SU_PI;
   e.eh_rad = eh_rad;
}

// NARP_cal_SunPosition steps 11-14, narp.f95:1086-1217.  COS(altitude*pi/180) in x is the reference's.
// The observer terms of the parallax step (narp.f95:1105-1119) and sin/cos of the latitude depend on the site only:
// SuSpaSite holds them, evaluated once per launch with the reference's expressions (6 sin/cos + tan + atan per step saved).
struct SuSpaSite {
   double x, y, sin_lat, cos_lat;
};
SU_DEV void spa_site_part(double lat, double alt, SuSpaSite &s)
{
   double u = atan(F(0.99664719) * tan(lat * SU_PI / 180));
   s.x = cos(u) + ((alt / 6378140) * cos(alt * SU_PI / 180));
   s.y = (0.99664719 * sin(u)) + ((alt / 6378140) * sin(lat * SU_PI / 180));
   s.sin_lat = sin(lat * SU_PI / F(180.0));
   s.cos_lat = cos(lat * SU_PI / F(180.0));
}
SU_BIG3 void spa_location_part(const SuewsEphem &e, const SuSpaSite &s, double lng, double &sunazimuth, double &sunzenith)
{
   double obs_hour = set_to_range(e.app_stime + lng - e.sun_ra);
   double eq_par = F(8.794) / (3600 * e.eh_rad);
   const double x = s.x, y = s.y;
   double sin_par = sin(eq_par * SU_PI / F(180.0));
   double sin_h = sin(obs_hour * SU_PI / F(180.0)), cos_h = cos(obs_hour * SU_PI / F(180.0));
   double sin_d = sin(e.sun_dec * SU_PI / F(180.0)), cos_d = cos(e.sun_dec * SU_PI / F(180.0));
   double num = -x * sin_par * sin_h;
   double den = cos_d - (x * sin_par * cos_h);
   double ra_par = atan2(num, den);
   double topo_ra_par = ra_par * F(180.0) / SU_PI;
   num = (sin_d - (y * sin_par)) * cos(ra_par);
   den = cos_d - (y * sin_par) * cos_h;
   double topo_dec = atan2(num, den) * F(180.0) / SU_PI;
   double topo_hour = obs_hour - topo_ra_par;

   const double sin_lat = s.sin_lat, cos_lat = s.cos_lat;
   double sin_th = sin(topo_hour * SU_PI / F(180.0)), cos_th = cos(topo_hour * SU_PI / F(180.0));
   double argument = (sin_lat * sin(topo_dec * SU_PI / F(180.0))) + (cos_lat * cos(topo_dec * SU_PI / F(180.0)) * cos_th);
   double corr_elev = asin(argument) * F(180.0) / SU_PI;
   argument = corr_elev + (F(10.3) / (corr_elev + F(5.11)));
   double refr = F(1.02) / (60 * tan(argument * SU_PI / F(180.0)));
   sunzenith = F(90.0) - (corr_elev + refr);
   num = sin_th;
   den = (cos_th * sin_lat) - (tan(topo_dec * SU_PI / F(180.0)) * cos_lat);
   sunazimuth = set_to_range((atan2(num, den) * F(180.0) / SU_PI) + F(180.0));
}
SU_INL void spa_location_part(const SuewsEphem &e, double lat, double lng, double alt, double &sunazimuth, double &sunzenith)
{
   SuSpaSite s;
   spa_site_part(lat, alt, s);
   spa_location_part(e, s, lng, sunazimuth, sunzenith);
}

// ------------------------------------------------------------------ NARP
// narp.f95:44-102
SU_INL void RadMethod(SuDiag &dg, int NetRadiationMethod, int snowUse, int &nrm_use, int &AlbedoChoice, int &ldown_option)
{
   AlbedoChoice = 0;
   ldown_option = 0;
   nrm_use = 0;
   if (NetRadiationMethod == 0) {
      if (snowUse == 1) { nrm_use = 3000; ldown_option = 3; }
   } else if (NetRadiationMethod > 0) {
      AlbedoChoice = -9;
      if (NetRadiationMethod < 100) {
         AlbedoChoice = 0;
         int m = NetRadiationMethod % 10;
         if (m >= 1 && m <= 3) ldown_option = m;
         nrm_use = NetRadiationMethod;
      } else if (NetRadiationMethod < 1000) {
         AlbedoChoice = 1;
         if (NetRadiationMethod == 100) ldown_option = 1;
         if (NetRadiationMethod == 200) ldown_option = 2;
         if (NetRadiationMethod == 300) ldown_option = 3;
         nrm_use = NetRadiationMethod / 100;
      }
      if (NetRadiationMethod % 10 > 3 || AlbedoChoice == -9) su_fatal(dg, 1001);   // bare STOP, narp.f95:95-99
   }
}

struct SuRad {
   double qn1, qn1_snowfree, qn1_S, kup, ldown, lup, fcld, tsurf, alb1;
   double qn1_ind_snow[NSURF], kup_ind_snow[NSURF], Tsurf_ind_snow[NSURF], Tsurf_ind[NSURF];
};

// narp.f95:105-457.  ea_hPa is what SUEWS_cal_Qn hands to NARP's Press_hPa slot (driver.f95:1437).
// PRATA_EMIS / WC_fraction do not depend on the surface and are evaluated once, not 7 times.
SU_BIG3 void NARP(const double *sfr, const double *SnowFrac, const double *alb, const double *emis,
                   const double *IceFrac, double NARP_EMIS_SNOW, double ZENITH_deg, double tsurf_0, double kdown,
                   double Temp_C, double RH, double ea_hPa, double qn1_obs, double SnowAlb, int AlbedoChoice,
                   int ldown_option, int nrm_use, bool store_ind, SuRad &r)
{   // store_ind: keep the per-surface arrays (Tsurf_ind and the snow-part terms) -- read by the snow routines and the snow
    // output line only, so snow-off launches without that line do not write them
   const double SIGMA_SB = F(5.67E-8);
   const double acos0 = (double)1.57079637f;   // ACOS(0.0) in REAL(4)
   const double tsurf_0_K = tsurf_0 + F(273.16);
   const double Temp_K = Temp_C + F(273.16);
   const double SIGMATK4 = SIGMA_SB * su_p4(Temp_K);
   const double ZENITH = ZENITH_deg * F(0.017453292);
   const bool zen_alb = (AlbedoChoice == 1 && 180 * ZENITH / acos0 < 90);
   const double dalb = zen_alb ? F(0.5e-16) * su_p8(180 * ZENITH / acos0) : 0.0;

   // PRATA_EMIS narp.f95:1253
   double W = F(46.5) * (ea_hPa / Temp_K);
   double EMIS_A = F(1.) - (F(1.) + W) * exp(-sqrt(F(1.2) + F(3.) * W));
   if (ldown_option == 3) {   // WC_fraction narp.f95:1284
      double Bc = F(0.00019) * Temp_C + F(0.015);
      double FWC = F(0.185) * (exp(Bc * RH) - 1);
      if (FWC > F(1.)) FWC = F(1.);
      if (FWC < F(0.)) FWC = F(0.);
      r.fcld = FWC;
   }
   if (ldown_option == 3 || ldown_option == 2) EMIS_A = EMIS_A + (F(1.) - EMIS_A) * r.fcld;   // EMIS_CLOUD
   if (ldown_option > 1) r.ldown = EMIS_A * SIGMATK4;
   const double LDOWN = r.ldown;

   double SF_all = 0;
   for (int is = 0; is < NSURF; is++)
      if (sfr[is] != 0) SF_all = SF_all + sfr[is] * (1 - SnowFrac[is]);

   double qn1_cum = 0, kup_cum = 0, lup_cum = 0, tsurf_cum = 0, qsf = 0, qs = 0;
   const double TS_SNOW_NARP =
      (nrm_use < 10) ? su_root4_cr((NARP_EMIS_SNOW * SIGMATK4) / (NARP_EMIS_SNOW * SIGMA_SB)) : 0.0;   // **0.25, correctly rounded
   for (int is = 0; is < NSURF; is++) {
      double ALB0 = zen_alb ? alb[is] + dalb : alb[is];
      double EMIS0 = emis[is];
      double LUPCORR = (kdown > 0) ? (1 - ALB0) * (F(0.08) * kdown) : F(0.);
      double KUP = ALB0 * kdown;
      double TSURF, LUP;
      if (nrm_use < 10) {
         TSURF = pow((EMIS0 * SIGMATK4 + LUPCORR) / (EMIS0 * SIGMA_SB), F(0.25));
         LUP = EMIS0 * SIGMATK4 + LUPCORR + (1 - EMIS0) * LDOWN;
      } else {
         TSURF = tsurf_0_K;
         LUP = EMIS0 * SIGMA_SB * su_p4(TSURF) + (1 - EMIS0) * LDOWN;
      }
      double QSTAR = kdown - KUP + LDOWN - LUP;
      TSURF = TSURF - F(273.16);

      double KUP_SNOW = 0, LUP_SNOW = 0, TSURF_SNOW = 0, QSTAR_SNOW = 0;
      if (SnowFrac[is] > 0) {
         double ALB1 = zen_alb ? SnowAlb + dalb : SnowAlb;
         r.alb1 = ALB1;
         KUP_SNOW = (ALB1 * (SnowFrac[is] - SnowFrac[is] * IceFrac[is]) + ALB0 * SnowFrac[is] * IceFrac[is]) * kdown;
         TSURF_SNOW = (nrm_use < 10) ? TS_SNOW_NARP : tsurf_0_K;
         LUP_SNOW = NARP_EMIS_SNOW * SIGMA_SB * su_p4(TSURF_SNOW) + (1 - NARP_EMIS_SNOW) * LDOWN;
         QSTAR_SNOW = kdown - KUP_SNOW + LDOWN - LUP_SNOW;
         TSURF_SNOW = TSURF_SNOW - F(273.16);
      }
      if (store_ind) {
         r.qn1_ind_snow[is] = QSTAR_SNOW;
         r.kup_ind_snow[is] = KUP_SNOW;
         r.Tsurf_ind_snow[is] = TSURF_SNOW;
      }

      if (SF_all != 0) qsf = qsf + QSTAR * sfr[is] * (1 - SnowFrac[is]) / SF_all;
      else qsf = qsf + QSTAR * sfr[is] * (1 - SnowFrac[is]);
      if ((1 - SF_all) != 0) qs = qs + su_div0(QSTAR_SNOW * sfr[is] * SnowFrac[is], 1 - SF_all);
      else qs = qs + QSTAR_SNOW * sfr[is] * SnowFrac[is];

      double qn1_is = QSTAR * (1 - SnowFrac[is]) + QSTAR_SNOW * SnowFrac[is];
      double kup_is = KUP * (1 - SnowFrac[is]) + KUP_SNOW * SnowFrac[is];
      double lup_is = LUP * (1 - SnowFrac[is]) + LUP_SNOW * SnowFrac[is];
      double tsurf_is = TSURF * (1 - SnowFrac[is]) + TSURF_SNOW * SnowFrac[is];
      qn1_cum = qn1_cum + (qn1_is * sfr[is]);
      kup_cum = kup_cum + (kup_is * sfr[is]);
      lup_cum = lup_cum + (lup_is * sfr[is]);
      tsurf_cum = tsurf_cum + (tsurf_is * sfr[is]);
      if (store_ind) r.Tsurf_ind[is] = tsurf_is;
   }
   r.qn1_snowfree = qsf;
   r.qn1_S = qs;
   r.qn1 = (nrm_use != 3000) ? qn1_cum : qn1_obs;
   r.kup = kup_cum;
   r.lup = lup_cum;
   r.tsurf = tsurf_cum;
}
