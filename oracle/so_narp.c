/* ORACLE (test infrastructure only -- see suews_oracle.h).
 * Restates SUEWS-SourceCode/suews_phys_narp.f95: RadMethod :44, NARP :105,
 * NARP_cal_SunPosition :460 and its helpers :568-1236, emissivity helpers
 * :1239-1297.  ldown_option 4/5 branches are unreachable through RadMethod
 * (it only yields 0-3) and are not restated. */
#include "suews_oracle.h"
#include "so_spa_tables.h"
#include <math.h>

/* GCC expands x**n (constant integer n, -O3, no -ffast-math) with its optimal
 * powi table: x3 = x*x2, x4 = x2*x2, x5 = x2*x3, ... (tree-ssa-math-opts). */
static double p2(double x) { return x * x; }
static double p3(double x) { return x * (x * x); }
static double p4(double x) { double a = x * x; return a * a; }
static double p5(double x) { double a = x * x; return a * (x * a); }
static double p6(double x) { double a = x * (x * x); return a * a; }
static double p7(double x) { double a = x * x; return (x * a) * (a * a); }
static double p8(double x) { double a = x * x; a = a * a; return a * a; }
static double p9(double x) { double c = x * (x * x); return c * (c * c); }
static double p10(double x) { double a = x * x; double e = a * (x * a); return e * e; }

/* narp.f95:44-102 */
void so_RadMethod(SoCtx *cx, int NetRadiationMethod, int snowUse, int *NetRadiationMethod_use, int *AlbedoChoice,
                  int *ldown_option)
{
   *AlbedoChoice = 0;
   *ldown_option = 0;
   if (NetRadiationMethod == 0) {
      *NetRadiationMethod_use = 0;
      if (snowUse == 1) {
         *NetRadiationMethod_use = 3000;
         *ldown_option = 3;
      }
   } else if (NetRadiationMethod > 0) {
      *AlbedoChoice = -9;
      if (NetRadiationMethod < 100) {
         *AlbedoChoice = 0;
         *NetRadiationMethod_use = NetRadiationMethod % 10;
         if (*NetRadiationMethod_use == 1) *ldown_option = 1;
         if (*NetRadiationMethod_use == 2) *ldown_option = 2;
         if (*NetRadiationMethod_use == 3) *ldown_option = 3;
         *NetRadiationMethod_use = NetRadiationMethod;
      } else if (NetRadiationMethod >= 100 && NetRadiationMethod < 1000) {
         *AlbedoChoice = 1;
         if (NetRadiationMethod == 100) *ldown_option = 1;
         if (NetRadiationMethod == 200) *ldown_option = 2;
         if (NetRadiationMethod == 300) *ldown_option = 3;
         *NetRadiationMethod_use = NetRadiationMethod / 100;
      }
      if (NetRadiationMethod % 10 > 3 || *AlbedoChoice == -9) {
         /* reference: bare STOP (narp.f95:95-99); reported as fatal code 1001 */
         so_error_hint(cx, 1001, "NetRadiationMethod: Value not usable", -55, -55, NetRadiationMethod);
      }
   }
}

/* narp.f95:1253-1260 */
static double prata_emis(double Temp_K, double EA_hPa)
{
   double W = F(46.5) * (EA_hPa / Temp_K);
   return F(1.) - (F(1.) + W) * exp(-sqrt(F(1.2) + F(3.) * W));
}

/* narp.f95:1262-1269 */
static double emis_cloud(double EMIS_A, double FCLD) { return EMIS_A + (F(1.) - EMIS_A) * FCLD; }

/* narp.f95:1284-1304 */
static double wc_fraction(double RH, double Temp)
{
   double A, B, FWC;
   A = F(0.185);
   B = F(0.00019) * Temp + F(0.015);
   FWC = A * (exp(B * RH) - 1);
   if (FWC > F(1.)) FWC = F(1.);
   if (FWC < F(0.)) FWC = F(0.);
   return FWC;
}

/* narp.f95:105-457.  Press_hPa is what the caller passes there: SUEWS_cal_Qn
 * hands in ea_hPa (driver.f95:1437). */
void so_NARP(SoCtx *cx, const double sfr[NSURF], const double SnowFrac[NSURF], const double alb[NSURF],
             const double emis[NSURF], const double IceFrac[NSURF], double NARP_TRANS_SITE, double NARP_EMIS_SNOW,
             double DTIME, double ZENITH_deg, double tsurf_0, double kdown, double Temp_C, double RH, double Press_hPa,
             double qn1_obs, double SnowAlb, int AlbedoChoice, int ldown_option, int NetRadiationMethod_use,
             int DiagQN, double *qstarall, double *qstar_sf, double *qstar_s, double *kclear, double *kupall,
             double *ldown, double *lupall, double *fcld, double *tsurfall, double qn1_ind_snow[NSURF],
             double kup_ind_snow[NSURF], double Tsurf_ind_snow[NSURF], double Tsurf_ind[NSURF], double *albedo_snowfree,
             double *albedo_snow)
{
   const double DEG2RAD = F(0.017453292), SIGMA_SB = F(5.67E-8);
   const double acos0 = (double)1.57079637f; /* ACOS(0.0) in REAL(4) */
   double tsurf_0_K, Temp_K, ZENITH, QSTAR, QSTAR_SNOW, KUP_SNOW, LUP_SNOW, TSURF_SNOW, KUP, LUP, TSURF;
   double EMIS0, EMIS_A, LUPCORR, SIGMATK4, KDOWN_HR;
   double qn1_cum, kup_cum, lup_cum, tsurf_cum, qn1_is, kup_is, lup_is, tsurf_is, SF_all;
   double ALB0 = 0, ALB1 = *albedo_snow; /* alb1 keeps its incoming value when no surface carries snow */
   int is;
   (void)cx; (void)NARP_TRANS_SITE; (void)DTIME; (void)DiagQN; (void)kclear;

   tsurf_0_K = tsurf_0 + F(273.16);
   Temp_K = Temp_C + F(273.16);
   SIGMATK4 = SIGMA_SB * p4(Temp_K);
   ZENITH = ZENITH_deg * DEG2RAD;

   qn1_cum = 0; kup_cum = 0; lup_cum = 0; tsurf_cum = 0;
   *qstar_sf = 0; *qstar_s = 0;
   for (is = 0; is < NSURF; is++) { qn1_ind_snow[is] = 0; kup_ind_snow[is] = 0; Tsurf_ind_snow[is] = 0; }

   SF_all = 0;
   for (is = 0; is < NSURF; is++)
      if (sfr[is] != 0) SF_all = SF_all + sfr[is] * (1 - SnowFrac[is]);

   for (is = 0; is < NSURF; is++) {
      EMIS_A = prata_emis(Temp_K, Press_hPa);

      if (AlbedoChoice == 1 && 180 * ZENITH / acos0 < 90) {
         ALB0 = alb[is] + F(0.5e-16) * p8(180 * ZENITH / acos0);
      } else {
         ALB0 = alb[is];
      }
      EMIS0 = emis[is];

      if (ldown_option == 3) {
         *fcld = wc_fraction(RH, Temp_C);
         EMIS_A = emis_cloud(EMIS_A, *fcld);
      } else if (ldown_option == 2) {
         EMIS_A = emis_cloud(EMIS_A, *fcld);
      }
      if (ldown_option > 1) {
         *ldown = EMIS_A * SIGMATK4;
      }

      KDOWN_HR = kdown;
      if (KDOWN_HR > 0) {
         LUPCORR = (1 - ALB0) * (F(0.08) * KDOWN_HR);
      } else {
         LUPCORR = F(0.);
      }

      KUP = ALB0 * kdown;

      if (NetRadiationMethod_use < 10) {
         TSURF = pow((EMIS0 * SIGMATK4 + LUPCORR) / (EMIS0 * SIGMA_SB), F(0.25));
         LUP = EMIS0 * SIGMATK4 + LUPCORR + (1 - EMIS0) * (*ldown);
      } else {
         TSURF = tsurf_0_K;
         LUP = EMIS0 * SIGMA_SB * p4(TSURF) + (1 - EMIS0) * (*ldown);
      }

      QSTAR = kdown - KUP + *ldown - LUP;
      TSURF = TSURF - F(273.16);

      if (SnowFrac[is] > 0) {
         if (AlbedoChoice == 1 && 180 * ZENITH / acos0 < 90) {
            ALB1 = SnowAlb + F(0.5e-16) * p8(180 * ZENITH / acos0);
         } else {
            ALB1 = SnowAlb;
         }
         KUP_SNOW = (ALB1 * (SnowFrac[is] - SnowFrac[is] * IceFrac[is]) + ALB0 * SnowFrac[is] * IceFrac[is]) * kdown;
         if (NetRadiationMethod_use < 10) {
            TSURF_SNOW = pow((NARP_EMIS_SNOW * SIGMATK4) / (NARP_EMIS_SNOW * SIGMA_SB), F(0.25));
            LUP_SNOW = NARP_EMIS_SNOW * SIGMA_SB * p4(TSURF_SNOW) + (1 - NARP_EMIS_SNOW) * (*ldown);
         } else {
            TSURF_SNOW = tsurf_0_K;
            LUP_SNOW = NARP_EMIS_SNOW * SIGMA_SB * p4(TSURF_SNOW) + (1 - NARP_EMIS_SNOW) * (*ldown);
         }
         QSTAR_SNOW = kdown - KUP_SNOW + *ldown - LUP_SNOW;
         TSURF_SNOW = TSURF_SNOW - F(273.16);
      } else {
         KUP_SNOW = 0;
         LUP_SNOW = 0;
         TSURF_SNOW = 0;
         QSTAR_SNOW = 0;
      }

      qn1_ind_snow[is] = QSTAR_SNOW;
      kup_ind_snow[is] = KUP_SNOW;
      Tsurf_ind_snow[is] = TSURF_SNOW;

      if (SF_all != 0) {
         *qstar_sf = *qstar_sf + QSTAR * sfr[is] * (1 - SnowFrac[is]) / SF_all;
      } else {
         *qstar_sf = *qstar_sf + QSTAR * sfr[is] * (1 - SnowFrac[is]);
      }
      if ((1 - SF_all) != 0) {
         *qstar_s = *qstar_s + QSTAR_SNOW * sfr[is] * SnowFrac[is] / (1 - SF_all);
      } else {
         *qstar_s = *qstar_s + QSTAR_SNOW * sfr[is] * SnowFrac[is];
      }

      qn1_is = QSTAR * (1 - SnowFrac[is]) + QSTAR_SNOW * SnowFrac[is];
      kup_is = KUP * (1 - SnowFrac[is]) + KUP_SNOW * SnowFrac[is];
      lup_is = LUP * (1 - SnowFrac[is]) + LUP_SNOW * SnowFrac[is];
      tsurf_is = TSURF * (1 - SnowFrac[is]) + TSURF_SNOW * SnowFrac[is];

      qn1_cum = qn1_cum + (qn1_is * sfr[is]);
      kup_cum = kup_cum + (kup_is * sfr[is]);
      lup_cum = lup_cum + (lup_is * sfr[is]);
      tsurf_cum = tsurf_cum + (tsurf_is * sfr[is]);

      Tsurf_ind[is] = tsurf_is;
   }

   if (NetRadiationMethod_use != 3000) {
      *qstarall = qn1_cum;
   } else {
      *qstarall = qn1_obs;
   }
   *kupall = kup_cum;
   *lupall = lup_cum;
   *tsurfall = tsurf_cum;
   *albedo_snowfree = ALB0;
   *albedo_snow = ALB1;
}

/* ---------------- NREL SPA (narp.f95:460-1236) ---------------- */

/* narp.f95:1219-1236 */
static double set_to_range(double var)
{
   double max_interval = F(360.0), min_interval = F(0.0), vari;
   vari = var - max_interval * (int)floor(var / max_interval);
   if (vari < min_interval) vari = vari + max_interval;
   return vari;
}

/* narp.f95:568-629 */
static void julian_calculation(double year, int month, int day, int hour, int min, double sec, double UTC,
                               double *juliancentury, double *julianday, double *julianephemeris_century,
                               double *julianephemeris_day, double *julianephemeris_millenium)
{
   double A, B, D, delta_t, M, ut_time, Y;
   if (month == 1 || month == 2) {
      Y = year - F(1.);
      M = month + 12;
   } else {
      Y = year;
      M = month;
   }
   /* float(min)/(60.*24.) is a REAL(4) quotient; float(hour) is REAL(4) widened for "- UTC" */
   ut_time = (((double)(float)hour - UTC) / F(24.)) + (double)((float)min / (60.f * 24.f)) +
             (sec / (double)(60.f * 60.f * 24.f));
   D = day + ut_time;
   if (year == F(1582.)) {
      if (month == 10) {
         if (day <= 4) {
            B = 0;
         } else if (day >= 15) {
            A = floor(Y / 100);
            B = 2 - A + floor(A / 4);
         } else {
            B = 0;
         }
      } else if (month < 10) {
         B = 0;
      } else {
         A = floor(Y / 100);
         B = 2 - A + floor(A / 4);
      }
   } else if (year < F(1582.)) {
      B = 0;
   } else {
      A = floor(Y / 100);
      B = 2 - A + floor(A / 4);
   }
   /* FLOOR() yields default INTEGERs: the first sum is an integer add */
   *julianday = ((int)floor(F(365.25) * (Y + F(4716.))) + (int)floor(F(30.6001) * (M + 1))) + D + B - F(1524.5);
   delta_t = F(0.);
   *julianephemeris_day = *julianday + (delta_t / 86400);
   *juliancentury = (*julianday - F(2451545.)) / F(36525.);
   *julianephemeris_century = (*julianephemeris_day - F(2451545.)) / F(36525.);
   *julianephemeris_millenium = *julianephemeris_century / F(10.);
}

static double sum_acos_terms(const double *A, const double *B, const double *C, int n, double JME)
{
   double s = 0;
   int i;
   for (i = 0; i < n; i++) s += A[i] * cos(B[i] + (C[i] * JME));
   return s;
}

/* narp.f95:631-828 */
static void earth_heliocentric_position_calculation(double JME, double *lat, double *lon, double *rad)
{
   const double pi = 3.141592653589793;
   double L0, L1, L2, L3, L4, L5;
   const double A5 = F(1.), B5 = F(3.1400), C5 = F(0.);
   const double A4j = 4, B4j = F(2.56), C4j = F(6283.08000);

   L0 = sum_acos_terms(SPA_A0, SPA_B0, SPA_C0, 64, JME);
   L1 = sum_acos_terms(SPA_A1, SPA_B1, SPA_C1, 34, JME);
   L2 = sum_acos_terms(SPA_A2, SPA_B2, SPA_C2, 20, JME);
   L3 = sum_acos_terms(SPA_A3, SPA_B3, SPA_C3, 7, JME);
   L4 = sum_acos_terms(SPA_A4, SPA_B4, SPA_C4, 3, JME);
   L5 = A5 * cos(B5 + (C5 * JME));
   *lon = (L0 + (L1 * JME) + (L2 * p2(JME)) + (L3 * p3(JME)) + (L4 * p4(JME)) + (L5 * p5(JME))) / F(1e8);
   *lon = *lon * F(180.) / pi;
   *lon = set_to_range(*lon);

   L0 = sum_acos_terms(SPA_A0i, SPA_B0i, SPA_C0i, 5, JME);
   L1 = sum_acos_terms(SPA_A1i, SPA_B1i, SPA_C1i, 2, JME);
   *lat = (L0 + (L1 * JME)) / F(1e8);
   *lat = *lat * 180 / pi;
   *lat = set_to_range(*lat);

   L0 = sum_acos_terms(SPA_A0j, SPA_B0j, SPA_C0j, 40, JME);
   L1 = sum_acos_terms(SPA_A1j, SPA_B1j, SPA_C1j, 10, JME);
   L2 = sum_acos_terms(SPA_A2j, SPA_B2j, SPA_C2j, 6, JME);
   L3 = sum_acos_terms(SPA_A3j, SPA_B3j, SPA_C3j, 2, JME);
   L4 = A4j * cos(B4j + (C4j * JME));
   *rad = (L0 + (L1 * JME) + (L2 * p2(JME)) + (L3 * p3(JME)) + (L4 * p4(JME))) / F(1e8);
}

/* narp.f95:842-961 */
static void nutation_calculation(double JCE, double *nutationlongitude, double *nutationobliquity)
{
   const double pi = 3.141592653589793;
   double X[5], arg, sl = 0, so = 0;
   int i;
   /* (1/189474.) etc. are REAL(4) quotients */
   const double p0[4] = {(double)(1 / 189474.f), F(-0.0019142), F(445267.11148), F(297.85036)};
   const double p1[4] = {(double)(-(1 / 300000.f)), F(-0.0001603), F(35999.05034), F(357.52772)};
   const double pp2[4] = {(double)(1 / 56250.f), F(0.0086972), F(477198.867398), F(134.96298)};
   const double pp3[4] = {(double)(1 / 327270.f), F(-0.0036825), F(483202.017538), F(93.27191)};
   const double pp4[4] = {(double)(1 / 450000.f), F(0.0020708), F(-1934.136261), F(125.04452)};
   X[0] = p0[0] * p3(JCE) + p0[1] * p2(JCE) + p0[2] * JCE + p0[3];
   X[1] = p1[0] * p3(JCE) + p1[1] * p2(JCE) + p1[2] * JCE + p1[3];
   X[2] = pp2[0] * p3(JCE) + pp2[1] * p2(JCE) + pp2[2] * JCE + pp2[3];
   X[3] = pp3[0] * p3(JCE) + pp3[1] * p2(JCE) + pp3[2] * JCE + pp3[3];
   X[4] = pp4[0] * p3(JCE) + pp4[1] * p2(JCE) + pp4[2] * JCE + pp4[3];
   for (i = 0; i < 63; i++) {
      const int *Y = &SPA_Y_terms1[5 * i];
      const double *N = &SPA_nutation_terms1[4 * i];
      arg = ((Y[0] * X[0]) + (Y[1] * X[1]) + (Y[2] * X[2]) + (Y[3] * X[3]) + (Y[4] * X[4])) * pi / 180;
      sl += ((N[0] + (N[1] * JCE))) * sin(arg);
      so += ((N[2] + (N[3] * JCE))) * cos(arg);
   }
   *nutationlongitude = sl / F(36000000.0);
   *nutationobliquity = so / F(36000000.0);
}

/* narp.f95:963-983 */
static double corr_obliquity_calculation(double JME, double nutationobliquity)
{
   const double *p = SPA_p;
   double U = JME / 10, mean_obliquity;
   mean_obliquity = p[0] * p10(U) + p[1] * p9(U) + p[2] * p8(U) + p[3] * p7(U) + p[4] * p6(U) + p[5] * p5(U) +
                    p[6] * p4(U) + p[7] * p3(U) + p[8] * p2(U) + p[9] * U + p[10];
   return (mean_obliquity / 3600) + nutationobliquity;
}

/* narp.f95:460-566 */
void so_NARP_cal_SunPosition(double year, double idectime, double UTC, double locationlatitude,
                             double locationlongitude, double locationaltitude, double *sunazimuth,
                             double *sunzenith)
{
   const double pi = 3.141592653589793;
   const double pi14 = 3.14159265358979; /* narp.f95:1024 carries fewer digits */
   double sec, juliancentury, julianday, jce, jde, jme;
   double eh_lat, eh_lon, eh_rad, sg_lat, sg_lon, nut_lon, nut_obl, corr_obl, aberr, app_sun_lon, app_stime;
   double mean_stime, sun_ra, sun_dec, obs_hour, topo_ra_par, topo_dec, topo_hour;
   double num, den, argument, eq_par, u, x, y, ra_par, corr_elev, refr, app_elev;
   int month, day, hour, min, seas, dayofyear, year_int;

   /* steps 1-10 (narp.f95:460-548) depend on the time and the time zone only: the minimal algorithm forms them once per
    * timestep for all the cells of a time zone, so the hoisted count leaves them out of the per-cell figure */
   SO_CNT_ONCE_BEGIN(1);
   so_dectime_to_timevec(idectime, &hour, &min, &sec);
   dayofyear = (int)floor(idectime);
   year_int = (int)year;
   so_day2month(dayofyear, &month, &day, &seas, year_int, locationlatitude);

   julian_calculation(year, month, day, hour, min, sec, UTC, &juliancentury, &julianday, &jce, &jde, &jme);
   earth_heliocentric_position_calculation(jme, &eh_lat, &eh_lon, &eh_rad);

   /* sun_geocentric_position_calculation narp.f95:830-840 */
   sg_lon = eh_lon + F(180.0);
   sg_lon = set_to_range(sg_lon);
   sg_lat = -eh_lat;
   sg_lat = set_to_range(sg_lat);

   nutation_calculation(jce, &nut_lon, &nut_obl);
   corr_obl = corr_obliquity_calculation(jme, nut_obl);

   /* narp.f95:985-993 */
   aberr = F(-20.4898) / (3600 * eh_rad);
   /* narp.f95:995-1006 */
   app_sun_lon = sg_lon + nut_lon + aberr;
   /* narp.f95:1008-1040: genuine double literals */
   mean_stime = 280.46061837 + (360.98564736629 * (julianday - 2451545.0)) + (0.000387933 * p2(juliancentury)) -
                (p3(juliancentury) / 38710000.0);
   mean_stime = set_to_range(mean_stime);
   app_stime = mean_stime + (nut_lon * cos(corr_obl * pi14 / 180));

   /* sun_rigth_ascension_calculation narp.f95:1042-1064 */
   num = (sin(app_sun_lon * pi / F(180.0)) * cos(corr_obl * pi / F(180.0))) -
         (tan(sg_lat * pi / F(180.0)) * sin(corr_obl * pi / F(180.0)));
   den = cos(app_sun_lon * pi / F(180.0));
   sun_ra = atan2(num, den) * F(180.0) / pi;
   sun_ra = set_to_range(sun_ra);

   /* sun_geocentric_declination_calculation narp.f95:1066-1084 */
   argument = (sin(sg_lat * pi / F(180.0)) * cos(corr_obl * pi / F(180.0))) +
              (cos(sg_lat * pi / F(180.0)) * sin(corr_obl * pi / 180) * sin(app_sun_lon * pi / F(180.0)));
   sun_dec = asin(argument) * F(180.0) / pi;

   SO_CNT_END;
   /* observer_local_hour_calculation narp.f95:1086-1100 */
   obs_hour = app_stime + locationlongitude - sun_ra;
   obs_hour = set_to_range(obs_hour);

   /* topocentric_sun_position_calculate narp.f95:1102-1157 (altitude, not latitude, in COS: as the reference) */
   eq_par = F(8.794) / (3600 * eh_rad);
   u = atan(F(0.99664719) * tan(locationlatitude * pi / 180));
   x = cos(u) + ((locationaltitude / 6378140) * cos(locationaltitude * pi / 180));
   y = (0.99664719 * sin(u)) + ((locationaltitude / 6378140) * sin(locationlatitude * pi / 180));
   num = -x * sin(eq_par * pi / F(180.0)) * sin(obs_hour * pi / F(180.0));
   den = cos(sun_dec * pi / F(180.0)) - (x * sin(eq_par * pi / F(180.0)) * cos(obs_hour * pi / F(180.0)));
   ra_par = atan2(num, den);
   topo_ra_par = ra_par * F(180.0) / pi;
   num = (sin(sun_dec * pi / F(180.0)) - (y * sin(eq_par * pi / F(180.0)))) * cos(ra_par);
   den = cos(sun_dec * pi / F(180.0)) - (y * sin(eq_par * pi / F(180.0))) * cos(obs_hour * pi / F(180.0));
   topo_dec = atan2(num, den) * F(180.0) / pi;

   /* topocentric_local_hour_calculate narp.f95:1159-1171 */
   topo_hour = obs_hour - topo_ra_par;

   /* sun_topocentric_zenith_angle_calculate narp.f95:1173-1217 */
   argument = (sin(locationlatitude * pi / F(180.0)) * sin(topo_dec * pi / F(180.0))) +
              (cos(locationlatitude * pi / F(180.0)) * cos(topo_dec * pi / F(180.0)) * cos(topo_hour * pi / F(180.0)));
   corr_elev = asin(argument) * F(180.0) / pi;
   argument = corr_elev + (F(10.3) / (corr_elev + F(5.11)));
   refr = F(1.02) / (60 * tan(argument * pi / F(180.0)));
   app_elev = corr_elev + refr;
   *sunzenith = F(90.0) - app_elev;
   num = sin(topo_hour * pi / F(180.0));
   den = (cos(topo_hour * pi / F(180.0)) * sin(locationlatitude * pi / F(180.0))) -
         (tan(topo_dec * pi / F(180.0)) * cos(locationlatitude * pi / F(180.0)));
   *sunazimuth = (atan2(num, den) * F(180.0) / pi) + F(180.0);
   *sunazimuth = set_to_range(*sunazimuth);
}
