/* ORACLE (test infrastructure only -- see suews_oracle.h).
 * Restates SUEWS-SourceCode/suews_util_time.f95, suews_util_meteo.f95,
 * suews_ctrl_error.f95 (classification only) and the three profile lookups in
 * suews_ctrl_input.f95:743-826. */
#include "suews_oracle.h"
#include <math.h>
#include <string.h>
#include <stdio.h>

/* ---------------- error sink (suews_ctrl_error.f95:1-415) ---------------- */
void so_ctx_init(SoCtx *cx) { memset(cx, 0, sizeof(*cx)); }

/* returnTrue (= warning, run continues) codes: suews_ctrl_error.f95:66-300 */
static int so_is_warning(int code)
{
   switch (code) {
   case 4: case 6: case 7: case 15: case 19: case 22: case 28: case 29: case 31: case 34: case 37: case 38:
   case 42: case 43: case 44: case 45: case 46: case 53: case 54: case 55: case 61: case 65: case 67:
   case 68: case 69:
      return 1;
   default:
      return 0;
   }
}

void so_error_hint(SoCtx *cx, int code, const char *msg, double v1, double v2, int iv)
{
   if (!cx) return;
   if (so_is_warning(code)) {
      if (code >= 0 && code < 128) cx->warn_count[code]++;
      return;
   }
   if (cx->fatal_code == 0) {
      cx->fatal_code = code;
      snprintf(cx->fatal_msg, sizeof(cx->fatal_msg), "%s", msg);
      cx->fatal_v1 = v1;
      cx->fatal_v2 = v2;
      cx->fatal_iv = iv;
   }
}

/* ---------------- suews_util_time.f95 ---------------- */

static int f_modulo(int a, int p) /* Fortran MODULO: result has sign of p */
{
   int r = a % p;
   if (r != 0 && ((r < 0) != (p < 0))) r += p;
   return r;
}

/* suews_util_time.f95:23-96 */
void so_day2month(int b, int *mb, int *md, int *seas, int year, double latitude)
{
   int k;
   *mb = 1;
   if ((f_modulo(year, 4) == 0 && f_modulo(year, 100) != 0) || f_modulo(year, 400) == 0)
      k = 1;
   else
      k = 0;
   if (b <= 31) { *mb = 1; *md = b; }
   else if (b > 31 && b <= 59 + k) { *mb = 2; *md = b - 31; }
   else if (b > 59 + k && b <= 90 + k) { *mb = 3; *md = b - (59 + k); }
   else if (b > 90 + k && b <= 120 + k) { *mb = 4; *md = b - (90 + k); }
   else if (b > 120 + k && b <= 151 + k) { *mb = 5; *md = b - (120 + k); }
   else if (b > 151 + k && b <= 181 + k) { *mb = 6; *md = b - (151 + k); }
   else if (b > 181 + k && b <= 212 + k) { *mb = 7; *md = b - (181 + k); }
   else if (b > 212 + k && b <= 243 + k) { *mb = 8; *md = b - (212 + k); }
   else if (b > 243 + k && b <= 273 + k) { *mb = 9; *md = b - (243 + k); }
   else if (b > 273 + k && b <= 304 + k) { *mb = 10; *md = b - (273 + k); }
   else if (b > 304 + k && b <= 334 + k) { *mb = 11; *md = b - (304 + k); }
   else if (b > 334 + k) { *mb = 12; *md = b - (334 + k); }

   if (latitude > 0) {
      *seas = (*mb > 3 && *mb < 10) ? 1 : 2;
   } else {
      *seas = (*mb < 4 || *mb > 9) ? 1 : 2;
   }
}

/* suews_util_time.f95:173-199 (integer division truncates toward zero, as C) */
void so_day_of_week(int date, int month, int year, int *dow)
{
   int yr = year, mn = month, n1, n2, day;
   if (!(mn > 2)) {
      mn = mn + 12;
      yr = yr - 1;
   }
   n1 = (26 * (mn + 1)) / 10;
   n2 = (125 * yr) / 100;
   day = (date + n1 + n2 - (yr / 100) + (yr / 400) - 1);
   *dow = day % 7 + 1; /* Fortran MOD: sign of dividend, same as C % */
}

/* suews_util_time.f95:202-223 */
void so_dectime_to_timevec(double dectime, int *hours, int *mins, double *secs)
{
   int doy = (int)floor(dectime);
   double DH = dectime - doy;
   double DM, DS;
   *hours = (int)(24 * DH);
   DM = 24 * DH - *hours;
   *mins = (int)(60 * DM);
   DS = 60 * DM - *mins;
   *secs = (int)(60 * DS);
}

/* suews_util_time.f95:270-283 */
void so_cal_dectime(int id, int it, int imin, int isec, double *dectime)
{
   *dectime = (double)(id - 1) + (double)it / 24 + (double)imin / (60 * 24) + (double)isec / (60 * 60 * 24);
}

/* suews_util_time.f95:286-298: nsh*1.0 and tstep*1.0 are REAL(4) products */
void so_cal_tstep(int tstep, int *nsh, double *nsh_real, double *tstep_real)
{
   *nsh = 3600 / tstep;
   *nsh_real = (double)((float)(*nsh) * 1.0f);
   *tstep_real = (double)((float)tstep * 1.0f);
}

/* suews_util_time.f95:301-323 */
void so_cal_weekday(int iy, int id, double lat, int dayofWeek_id[3])
{
   int wd, mb, date, seas;
   so_day2month(id, &mb, &date, &seas, iy, lat);
   so_day_of_week(date, mb, iy, &wd);
   dayofWeek_id[0] = wd;
   dayofWeek_id[1] = mb;
   dayofWeek_id[2] = seas;
}

/* suews_util_time.f95:326-337 */
void so_cal_DLS(int id, int startDLS, int endDLS, int *DLS)
{
   *DLS = 0;
   if (id > startDLS && id < endDLS) *DLS = 1;
}

/* ---------------- suews_util_meteo.f95 ---------------- */

/* suews_util_meteo.f95:25-39 (Buck 1981; used by RH2qa/qa2RH) */
double so_sat_vap_press(double TK, double P)
{
   double TC, e, f;
   TC = TK - F(273.15);
   if (TC == 0) TC = F(0.001);
   e = F(6.1121) * exp(((F(18.729) - TC / F(227.3)) * TC) / (TC + F(257.87)));
   f = F(1.00072) + P * (F(3.2E-6) + F(5.9E-10) * (TC * TC));
   return e * f;
}

/* suews_util_meteo.f95:148-192.  NB writes back to its Temp_C argument. */
double so_sat_vap_press_x(SoCtx *cx, double *Temp_C, double Press_hPa, int from, double dectime)
{
   double e_mb, f, Press_kPa, es_hPa;
   es_hPa = 1000;
   if (fabs(*Temp_C) < F(0.001000)) {
      if (from == 1) {
         int iv = (int)Press_hPa;
         so_error_hint(cx, 29, "Function sat_vap_press: temp_C, dectime,press_Hpa = ", *Temp_C, dectime, iv);
      }
      *Temp_C = F(0.001000);
   }
   Press_kPa = Press_hPa / 10;
   if (*Temp_C < 50 && *Temp_C > -40) {
      double T = *Temp_C;
      if (T >= F(0.001000)) {
         e_mb = F(6.1121) * exp(((F(18.678) - T / F(234.5)) * T) / (T + F(257.14)));
         f = F(1.00072) + Press_kPa * (F(3.2E-6) + F(5.9E-10) * (T * T));
         es_hPa = e_mb * f;
      } else if (T <= -F(0.001000)) {
         e_mb = F(6.1115) * exp(((F(23.036) - T / F(333.7)) * T) / (T + F(279.82)));
         f = F(1.00022) + Press_kPa * (F(3.83E-6) + F(6.4E-10) * (T * T));
         es_hPa = e_mb * f;
      }
   } else {
      so_error_hint(cx, 28, "FUNCTION sat_vap_press: [Temperature is out of range], Temp_C,dectime", *Temp_C, dectime,
                    -55);
   }
   return es_hPa;
}

/* suews_util_meteo.f95:195-228 */
double so_sat_vap_pressIce(SoCtx *cx, double *Temp_C, double Press_hPa, int from, double dectime)
{
   double e_mb, f, Press_kPa, es_hPa;
   es_hPa = 10;
   if (fabs(*Temp_C) < F(0.001000)) {
      if (from == 1) {
         int iv = (int)Press_hPa;
         so_error_hint(cx, 29, "Function sat_vap_press: temp_C, dectime,press_Hpa = ", *Temp_C, dectime, iv);
      }
      *Temp_C = F(0.001000);
   }
   Press_kPa = Press_hPa / 10;
   if (*Temp_C < 50 && *Temp_C > -40) {
      double T = *Temp_C;
      e_mb = F(6.1115) * exp(((F(23.036) - T / F(333.7)) * T) / (T + F(279.82)));
      f = F(1.00022) + Press_kPa * (F(3.83E-6) + F(6.4E-10) * (T * T));
      es_hPa = e_mb * f;
   } else {
      so_error_hint(cx, 28, "FUNCTION sat_vap_press: [Temperature is out of range], Temp_C,dectime", *Temp_C, dectime,
                    -55);
   }
   return es_hPa;
}

/* suews_util_meteo.f95:235-241 */
double so_spec_hum_def(double vpd_hPa, double press_hPa)
{
   const double epsil_gkg = F(621.97);
   return epsil_gkg * vpd_hPa / press_hPa;
}

/* suews_util_meteo.f95:244-271 */
double so_spec_heat_beer(SoCtx *cx, double Temp_C, double rh, double rho_v, double rho_d)
{
   double cpd, cpm, cp, t;
   t = Temp_C + F(23.16);
   cpd = F(1005.0) + (t * t) / F(3364.0);
   t = Temp_C / F(100.);
   cpm = 1859 + F(0.13) * rh + (F(19.3) + F(0.569) * rh) * (Temp_C / F(100.)) + (F(10.) + F(0.5) * rh) * (t * t);
   if (fabs(rho_d) < F(0.000100) || fabs(rho_v) < F(0.000100) || fabs(rho_d + rho_v) < F(0.000100)) {
      so_error_hint(cx, 42, "spec-heat_beer", rho_v, rho_d, (int)Temp_C);
   }
   cp = cpd * (rho_d / (rho_d + rho_v)) + cpm * (rho_v / (rho_d + rho_v));
   return cp;
}

/* suews_util_meteo.f95:409-427; epsil from module gas (suews_ctrl_const.f95:1292) */
double so_psyc_const(SoCtx *cx, double cp, double Press_hPa, double lv_J_kg)
{
   const double epsil = F(0.62197);
   if (cp * Press_hPa < 900 || lv_J_kg < 10000) {
      so_error_hint(cx, 19, "in psychrometric constant calculation:  cp [J kg-1 K-1], p [hPa], Lv [J kg-1]", cp,
                    Press_hPa, (int)lv_J_kg);
   }
   return (cp * Press_hPa) / (epsil * lv_J_kg);
}

/* suews_util_meteo.f95:277-346.  switch1/switch2 are implicit-SAVE locals in
 * the reference; they live in the per-cell state here (cx->lv_switch*). */
double so_Lat_vap(SoCtx *cx, double Temp_C, double Ea_hPa, double Press_hPa, double cp, double dectime)
{
   double lv_J_kg, ea_fix, tw, incr, es_tw, psyc, ea_est;
   int ii;
   const int from = 2;
   const double notUsed = F(-55.55);
   int *switch1 = cx->lv_switch1, *switch2 = cx->lv_switch2;

   ea_fix = Ea_hPa;
   lv_J_kg = (F(2500.25) - F(2.365) * Temp_C) * 1000;
   tw = Temp_C / F(2.);
   incr = F(3.);
   for (ii = 1; ii <= 100; ii++) {
      if (Press_hPa < 900) so_error_hint(cx, 45, "function Lat_vap", Press_hPa, notUsed, ii);
      es_tw = so_sat_vap_press_x(cx, &tw, Press_hPa, from, dectime);
      if (Press_hPa < 900) so_error_hint(cx, 45, "function Lat_vap - 2", Press_hPa, notUsed, ii);
      psyc = so_psyc_const(cx, cp, Press_hPa, lv_J_kg);
      if (Press_hPa < 900) so_error_hint(cx, 45, "function Lat _vap -31", Press_hPa, notUsed, ii);
      ea_est = es_tw - psyc * (Temp_C - tw);
      lv_J_kg = (F(2500.25) - F(2.365) * tw) * F(1e3);
      if (*switch1 && *switch2) {
         incr = incr / F(10.);
         *switch1 = 0;
         *switch2 = 0;
      }
      if (fabs(ea_est - ea_fix) < F(0.001000)) {
         return lv_J_kg;
      } else if (ea_est > ea_fix) {
         tw = tw - incr;
         *switch1 = 1;
      } else if (ea_est < ea_fix) {
         tw = tw + incr;
         *switch2 = 1;
      }
   }
   return lv_J_kg;
}

/* suews_util_meteo.f95:348-401 (search loop is commented out in the reference) */
double so_Lat_vapSublim(double Temp_C, double Ea_hPa, double Press_hPa, double cp)
{
   (void)Ea_hPa; (void)Press_hPa; (void)cp;
   return (F(2834.1) - F(0.29) * Temp_C) * F(1e3);
}

/* suews_util_meteo.f95:444-468: true double-precision (D) literals */
double so_slope_svp(double t)
{
   const double B1 = 4.438099984e-1, B2 = 2.857002636e-2, B3 = 7.938054040e-4, B4 = 1.215215065e-5,
                B5 = 1.036561403e-7, B6 = 3.532421810e-10, B7 = -7.090244804e-13;
   return B1 + t * (B2 + t * (B3 + t * (B4 + t * (B5 + t * (B6 + B7 * t)))));
}

/* suews_util_meteo.f95:471-492 */
double so_slopeIce_svp(double t)
{
   const double B1 = 5.030305237e-1, B2 = 3.773255020e-2, B3 = 1.267995369e-3, B4 = 2.477563108e-5,
                B5 = 3.005693132e-7, B6 = 2.158542548e-9, B7 = 7.131097725e-12;
   return B1 + t * (B2 + t * (B3 + t * (B4 + t * (B5 + t * (B6 + B7 * t)))));
}

/* suews_util_meteo.f95:495-516 */
double so_qsatf(SoCtx *cx, double T, double PMB)
{
   const double A = F(6.106), B = F(17.27), C = F(237.3), molar = F(0.028965), molar_wat_vap = F(0.0180153);
   double ES;
   if (T > 55) so_error_hint(cx, 34, "Function qsatf", T, 0.0, -55);
   ES = A * exp(B * T / (C + T));
   return (molar_wat_vap / molar) * ES / PMB;
}

/* suews_util_meteo.f95:518-534 */
double so_RH2qa(double RH_dec, double pres_hPa, double Ta_degC)
{
   double es = so_sat_vap_press(Ta_degC + F(273.15), pres_hPa);
   double ea = es * RH_dec;
   return F(0.622) * ea / (pres_hPa - F(0.378) * ea) * 1000;
}

/* suews_util_meteo.f95:536-555 */
double so_qa2RH(double qa_gkg, double pres_hPa, double Ta_degC)
{
   double qa_kgkg = qa_gkg / 1000;
   double es = so_sat_vap_press(Ta_degC + F(273.15), pres_hPa);
   double ea = 500 * pres_hPa * qa_kgkg / (311 + 189 * qa_kgkg);
   return ea / es;
}

/* ---------------- suews_ctrl_input.f95:743-826 ---------------- */

static double sum24(const double p[24])
{
   double s = 0;
   int i;
   for (i = 0; i < 24; i++) s += p[i]; /* sequential, as gfortran SUM without -ffast-math */
   return s;
}

/* suews_ctrl_input.f95:743-764 */
double so_get_Prof_SpecTime_inst(int Hour, int Min, int Sec, const double Prof_24h[24])
{
   int total_sec = Min * 60 + Sec, SecPerHour = 3600;
   int i = Hour, j = i + 1;
   double deltaProf;
   if (j == 24) j = 0;
   deltaProf = (Prof_24h[j] - Prof_24h[i]) / SecPerHour;
   return Prof_24h[Hour] + deltaProf * total_sec;
}

/* suews_ctrl_input.f95:769-796 */
double so_get_Prof_SpecTime_mean(int Hour, int Min, int Sec, const double Prof_24h[24])
{
   int total_sec = Min * 60 + Sec, SecPerHour = 3600;
   int i = Hour, j = i + 1;
   double s = sum24(Prof_24h), pi_, pj_, deltaProf;
   if (j == 24) j = 0;
   if (s != 0) {
      double m = s / 24;
      pi_ = Prof_24h[i] / m;
      pj_ = Prof_24h[j] / m;
   } else {
      pi_ = 0.0;
      pj_ = 0.0;
   }
   deltaProf = (pj_ - pi_) / SecPerHour;
   return pi_ + deltaProf * total_sec;
}

/* suews_ctrl_input.f95:801-825 */
double so_get_Prof_SpecTime_sum(int Hour, int Min, int Sec, const double Prof_24h[24], int dt)
{
   int total_sec = Min * 60 + Sec, SecPerHour = 3600;
   int i = Hour, j = i + 1;
   double s = sum24(Prof_24h), pi_, pj_, deltaProf, v;
   if (j == 24) j = 0;
   if (s != 0) {
      pi_ = Prof_24h[i] / s;
      pj_ = Prof_24h[j] / s;
   } else {
      pi_ = 0.0;
      pj_ = 0.0;
   }
   deltaProf = (pj_ - pi_) / SecPerHour;
   v = pi_ + deltaProf * total_sec;
   v = v * dt / SecPerHour;
   return v;
}
