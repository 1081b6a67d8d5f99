// dev_estm.cuh -- ESTM, the Element Surface Temperature Method for the storage heat flux
// (StorageHeatMethod 4/14): SUEWS-SourceCode/suews_phys_estm.f95 -- NewtonPolynomial :206-250,
// heatcond1d :480-510, the every-call half of ESTM_translate :718-939 and ESTM :970-1605 -- with
// HEATCAPACITY_AIR and its helpers (suews_util_meteo.f95:25-39, :62-93, :103-113).
//
// The reference keeps ESTM's parameters and element temperatures in MODULE ESTM_data, outside the
// SUEWS_cal_Main argument list.  Here the parameters are class fields (suews_fields.h), the per-cell wall area
// is a cell field, and the variables carried from call to call are a per-cell state record (SuewsEstmState).
//
// Kept from the reference on purpose:
//  * SUEWS_cal_Qs is inside the surface-temperature DO WHILE (driver.f95:642, :776) and ESTM stores its element
//    temperatures at the end of every call (:1587-1603), so the conduction scheme advances once per TRIP of that
//    loop.  The caller therefore runs this once per trip and never hoists it.
//  * radforce / groundradforce are hard-coded .FALSE. (:1191-1192): the external radiation balance (Rs_*,
//    Rl_ground/wall/roof, kup_estm, alb_avg, RN_*, EM_EQUIV, QHestm, QFBld) and the view factors of
//    ESTM_translate reach neither the state nor an output, and are not formed.
//  * NewtonPolynomial's `converged`, HVAC and Tair2Set are SAVE'd; Tstep/3600.0 and 4.*Tstep/3600.0 are REAL(4).
#pragma once
#include "dev_atmos.cuh"

#define SU_C2K F(273.15)          // suews_ctrl_const.f95:2086
#define SU_SBCONST F(5.67051e-8)  // suews_ctrl_const.f95:2087

// suews_util_meteo.f95:84-93
SU_DEV double heatcapacity_air(double TK, double RH, double P)
{
   const double R_DRY_MASS = F(8.31451) / F(0.028965);   // meteo.f95:14-15
   const double EPSIL = F(0.62197);
   const double RHO_D = P * F(100.) / (R_DRY_MASS * TK);
   const double EA = sat_vap_press(TK, P) * RH / F(100.);
   const double RHO_V = (EA * F(100.) * EPSIL) / (R_DRY_MASS * TK);
   const double TC = TK - F(273.15);
   const double CPD = F(1005.) + ((TC + F(23.15)) * (TC + F(23.15))) / F(3364.);
   const double TC_100 = (TK - F(273.15)) / F(100.);
   const double CPV = F(1859.) + F(0.13) * RH + (F(19.3) + F(0.569) * RH) * TC_100 + (F(10.) + F(0.5) * RH) * (TC_100 * TC_100);
   return RHO_D * CPD + RHO_V * CPV;
}

// estm.f95:206-250 for the quartic a x^4 + 0 x^3 + 0 x^2 + c x + d every call site passes.  x**k with a run-time
// integer k is libgcc's square-and-multiply (x^4 = (x^2)^2, x^3 = x * x^2); the zero coefficients add exact zeros.
SU_DEV double NewtonPolynomial4(double x0, double a, double c, double d, double &converged_save)
{
   const double conv = F(0.0001);
   double e = (double)3.40282347e+38f;   // HUGE(1.)
   double x = x0;
#pragma unroll 1
   for (int i = 1; i <= 100; i++) {
      if (fabs(e) < conv) {
         converged_save = 1;
         break;
      }
      const double x2 = x * x;
      double f = 0.0 + a * (x2 * x2);
      f = f + 0.0 * (x * x2);
      f = f + 0.0 * x2;
      f = f + c * x;
      double fp = 0.0 + a * 4.0 * (x * x2);
      fp = fp + 0.0 * 3.0 * x2;
      fp = fp + 0.0 * 2.0 * x;
      fp = fp + c * 1.0 * 1.0;
      f = f + d;
      const double xprev = x;
      if (fp == 0) fp = (double)1.17549435e-38f;   // TINY(1.)
      x = xprev - f / fp;
      e = x - xprev;
   }
   if (converged_save == 0) x = x0;
   return x;
}

// estm.f95:480-510 (bctype is .FALSE. at every call site)
SU_DEV double heatcond1d(double *T, const double *dx, double dt, const double *k, const double *rhocp, int n, double bc1,
                         double bc2)
{
   double w[6], a[5], T1[5];
   for (int i = 1; i <= n; i++) w[i] = T[i - 1];
   w[0] = bc1;
   w[n] = bc2;
   for (int i = 0; i < n; i++) a[i] = k[i] / dx[i];
   for (int i = 1; i <= n - 1; i++) w[i] = (T[i] * a[i] + T[i - 1] * a[i - 1]) / (a[i - 1] + a[i]);
   for (int i = 1; i <= n; i++)
      T1[i - 1] = (dt / rhocp[i - 1]) * (w[i - 1] - 2 * T[i - 1] + w[i]) * 2 * a[i - 1] / dx[i - 1] + T[i - 1];
   const double Qs = (w[0] - T[0]) * 2 * a[0] + (w[n] - T[n - 1]) * 2 * a[n - 1];
   for (int i = 0; i < n; i++) T[i] = T1[i];
   return Qs;
}

// estm.f95:970-1605.  Ts: Ts5mindata_ir(5:13).  o: dataOutLineESTM (27) or nullptr.  Returns QS.
SU_COLD double ESTM(const SuewsClassParams &P, const SuewsCellParams &C, SuewsEstmState &E, double avu1, double temp_c,
                   double avrh, double press_hpa, const double *Ts, double Tair_av, double *o)
{
   const double *sfr = C.sfr.v;
   const int tstep = P.tstep;
   const double bldgh = C.bldgH;
   const double em_wall = F(0.9);   // estm.f95:169
   // ---- ESTM_translate, every-call half, and suews_ctrl_translate.f95:393-394, :610-614
   const double fwall = C.AreaWall / C.SurfaceArea;
   const double froof = sfr[BLDG];
   const double fground = sfr[PAV] + sfr[CONIF] + sfr[DECID] + sfr[GRASS] + sfr[BSOIL] + sfr[WATER];
   const double ZREF = F(2.0) * bldgh;
   const double em_roof = P.emis[BLDG];
   const double em = P.em_ibld;   // em_r = em_w = em_i = em_f = em_ibld (estm.f95:899)
   double em_ground = -999;
   if (fground != 0)
      em_ground = (P.emis[CONIF] * sfr[CONIF] + P.emis[DECID] * sfr[DECID] + P.emis[GRASS] * sfr[GRASS] +
                   P.emis[PAV] * sfr[PAV] + P.emis[BSOIL] * sfr[BSOIL] + P.emis[WATER] * sfr[WATER]) / fground;
   double fibld;
   if (P.nroom == 0) fibld = (floor(bldgh / F(3.1) - F(0.5)) - 1) * froof;
   else fibld = (F(2.) - F(2.) / P.nroom) * fwall + (floor(bldgh / F(3.1) - F(0.5)) - 1) * froof;
   if (fibld == 0) fibld = F(0.00001);
   double finternal = froof + fibld + fwall;
   if (finternal == 0) finternal = F(0.00001);
   double fair = ZREF - bldgh * froof;
   if (fair == 0) fair = F(0.00001);

   double &Tievolve = E.Tievolve.v[0], &T0_ibld = E.T0_ibld.v[0], &TN_wall = E.TN_wall.v[0], &TN_roof = E.TN_roof.v[0];
   double &HVAC = E.HVAC.v[0], &conv_save = E.NewtonConverged.v[0];

   // ---- ESTM
   const double bldgHX = (bldgh > F(0.001)) ? bldgh : F(0.001);
   const double Tair1 = temp_c + SU_C2K;
   E.Tair2Set = E.Tair2Set + 1;   // Gridiv == 1
   if (E.Tair2Set == 1) E.Tair2 = temp_c + SU_C2K;
   (void)avu1;   // WS only enters the radforce branches

   const double Tinternal = Ts[0], Tsurf_all = Ts[1], Troof_in = Ts[2], Troad = Ts[3], Twall_all = Ts[4];
   const double Tw_n = Ts[5], Tw_e = Ts[6], Tw_s = Ts[7], Tw_w = Ts[8];

   const double SHC_air = heatcapacity_air(Tair1, avrh, press_hpa);
   bool diagnoseTi = false;
   if (P.evolveTibld == 0) {
      HVAC = 0;
   } else if (P.evolveTibld == 1) {
      diagnoseTi = true;
      if (Tievolve > P.THEAT_OFF) HVAC = 0;
      else if (Tievolve < P.THEAT_ON) HVAC = 1;
   } else if (P.evolveTibld == 2) {
      diagnoseTi = true;
   }
   double AIREXHR;
   if (Tair_av > F(20.) + SU_C2K && Tievolve > F(25.) + SU_C2K && Tair1 < Tievolve && !(HVAC != 0)) AIREXHR = F(2.0);
   else if (Tair_av < F(17.) + SU_C2K || HVAC != 0) AIREXHR = F(0.5);
   else AIREXHR = F(1.0);
   const double AIREXDT = AIREXHR * (double)((float)tstep / 3600.0f);
   double shc_airbld = heatcapacity_air(Tievolve, avrh, press_hpa);
   if (!(shc_airbld > F(0.00001))) shc_airbld = F(0.00001);

   double CH_ibld = P.CH_ibld, CH_iwall = P.CH_iwall, CH_iroof = P.CH_iroof;
   if (P.ibldCHmod == 1) {   // ASHRAE 2001
      CH_ibld = F(1.31) * pow(fabs(T0_ibld - Tievolve), F(0.25)) / shc_airbld;
      CH_iwall = F(1.31) * pow(fabs(TN_wall - Tievolve), F(0.25)) / shc_airbld;
      CH_iroof = F(1.52) * pow(fabs(TN_roof - Tievolve), F(0.25)) / shc_airbld;
      if (fabs(TN_roof - Tievolve) > 0) CH_iroof = CH_iroof * F(0.39);
   } else if (P.ibldCHmod == 2) {   // Awbi 1998
      CH_ibld = F(1.823) * pow(fabs(T0_ibld - Tievolve), F(0.293)) / shc_airbld;
      CH_iwall = F(1.823) * pow(fabs(TN_wall - Tievolve), F(0.293)) / shc_airbld;
      CH_iroof = F(2.175) * pow(fabs(TN_roof - Tievolve), F(0.308)) / shc_airbld;
      if (fabs(TN_roof - Tievolve) > 0) CH_iroof = F(0.704) * pow(fabs(TN_roof - Tievolve), F(0.133)) / shc_airbld;
   }

   const double Tairmix = (Tievolve + Tair1 * AIREXDT) / (F(1.0) + AIREXDT);
   Tievolve = Tairmix + tstep / bldgHX / finternal *
                           (CH_ibld * fibld * (T0_ibld - Tievolve) + CH_iroof * froof * (TN_roof - Tievolve) +
                            CH_iwall * fwall * (TN_wall - Tievolve));
   if (!diagnoseTi) Tievolve = Tinternal + SU_C2K;
   if (HVAC != 0) {
      const float lim4 = 4.f * (float)tstep / 3600.0f;
      const double dT = P.THEAT_fix - Tievolve;
      const double Tadd = (copysign(1.0, dT) + P.THEAT_fix - Tievolve) * (double)((lim4 < 0.9f) ? lim4 : 0.9f);
      Tievolve = Tievolve + Tadd;
   }

   // internal longwave exchange (estm.f95:1404-1417); Rs_i* = 0
   const double tw4 = su_p4(TN_wall), tr4 = su_p4(TN_roof), ti4 = su_p4(T0_ibld), tf4 = su_p4(P.Tfloor);
   const double Rl_ibld = SU_SBCONST * (P.ivf_iw * em * tw4 + P.ivf_ir * em * tr4 + P.ivf_if * em * tf4);
   const double Rl_iwall = SU_SBCONST * (P.ivf_wi * em * ti4 + P.ivf_wr * em * tr4 + P.ivf_wf * em * tf4);
   const double Rl_iroof = SU_SBCONST * (P.ivf_ri * em * ti4 + P.ivf_rw * em * tw4 + P.ivf_rf * em * tf4);
   const double dt = (double)tstep;
   const int Nibld = P.Nibld, Nwall = P.Nwall, Nroof = P.Nroof, Nground = P.Nground;

   // ========> internal mass
   double kdz = 2 * P.kibld[0] / P.zibld[0];
   T0_ibld = NewtonPolynomial4(T0_ibld, P.em_ibld * SU_SBCONST * (1 - P.ivf_ii * P.em_ibld), kdz + shc_airbld * CH_ibld,
                               -kdz * E.Tibld[0] - shc_airbld * CH_ibld * Tievolve - 0.0 - Rl_ibld, conv_save);
   double Qsibld = heatcond1d(E.Tibld.v, P.zibld.v, dt, P.kibld.v, P.ribld.v, Nibld, T0_ibld, T0_ibld);

   // ========> walls
   kdz = 2 * P.kwall[Nwall - 1] / P.zwall[Nwall - 1];
   TN_wall = NewtonPolynomial4(TN_wall, P.em_ibld * SU_SBCONST * (1 - P.ivf_ww * P.em_ibld), kdz + shc_airbld * CH_iwall,
                               -kdz * E.Twall[Nwall - 1] - shc_airbld * CH_iwall * Tievolve - 0.0 - Rl_iwall, conv_save);
   double Qswall = 0;
   if (P.TsurfChoice < 2) {
      double bc1 = T0_ibld;   // bc(1) is left from the internal element when TsurfChoice < 0
      if (P.TsurfChoice == 0) { bc1 = Tsurf_all + SU_C2K; E.T0_wall = bc1; }
      else if (P.TsurfChoice == 1) { bc1 = Twall_all + SU_C2K; E.T0_wall = bc1; }
      Qswall = heatcond1d(E.Twall.v, P.zwall.v, dt, P.kwall.v, P.rwall.v, Nwall, bc1, TN_wall);
   } else if (P.TsurfChoice == 2) {   // four walls, each driven by the SUM of the four wall temperatures (as written)
      double t0 = 0, qsum = 0;
#pragma unroll 1
      for (int w = 0; w < 4; w++) {
         const double bc1 = Tw_n + Tw_e + Tw_s + Tw_w + SU_C2K;
         t0 = t0 + bc1;
         qsum = qsum + heatcond1d(&E.Tw_4.v[5 * w], P.zwall.v, dt, P.kwall.v, P.rwall.v, Nwall, bc1, TN_wall);
      }
      E.T0_wall = t0 / F(4.);
      Qswall = qsum / F(4.);
      for (int i = 0; i < Nwall; i++) {
         double s = 0;
         for (int w = 0; w < 4; w++) s = s + E.Tw_4[i + 5 * w];
         E.Twall[i] = s / F(4.);
      }
   }

   // ========> roof
   kdz = 2 * P.kroof[Nroof - 1] / P.zroof[Nroof - 1];
   TN_roof = NewtonPolynomial4(TN_roof, P.em_ibld * SU_SBCONST, kdz + shc_airbld * CH_iroof,
                               -kdz * E.Troof[Nroof - 1] - shc_airbld * CH_iroof * Tievolve - 0.0 - Rl_iroof, conv_save);
   {
      const double bc1 = ((P.TsurfChoice == 0) ? Tsurf_all : Troof_in) + SU_C2K;
      E.T0_roof = bc1;
      // fall through to the conduction sweep
   }
   double Qsroof = heatcond1d(E.Troof.v, P.zroof.v, dt, P.kroof.v, P.rroof.v, Nroof, E.T0_roof, TN_roof);

   // ========> ground
   E.T0_ground = ((P.TsurfChoice == 0) ? Tsurf_all : Troad) + SU_C2K;
   double Qsground = -999;
   if (fground != 0)
      Qsground = heatcond1d(E.Tground.v, P.zground.v, dt, P.kground.v, P.rground.v, Nground, E.T0_ground, P.LBC_soil + SU_C2K);

   const double Qsair = fair * SHC_air * (Tair1 - E.Tair2) / tstep;
   Qsibld = Qsibld * fibld;
   Qswall = Qswall * fwall;
   Qsroof = Qsroof * froof;
   Qsground = Qsground * fground;
   const double QS = Qsibld + Qswall + Qsroof + Qsground;

   E.lup_ground = SU_SBCONST * em_ground * su_p4(E.T0_ground);
   E.lup_wall = SU_SBCONST * em_wall * su_p4(E.T0_wall);
   E.lup_roof = SU_SBCONST * em_roof * su_p4(E.T0_roof);

   if (o) {
      o[0] = QS; o[1] = Qsair; o[2] = Qswall; o[3] = Qsroof; o[4] = Qsground; o[5] = Qsibld;
      for (int i = 0; i < 5; i++) {
         o[6 + i] = (i < Nwall) ? (double)E.Twall[i] : -999.0;
         o[11 + i] = (i < Nroof) ? (double)E.Troof[i] : -999.0;
         o[16 + i] = (i < Nground) ? (double)E.Tground[i] : -999.0;
         o[21 + i] = (i < Nibld) ? (double)E.Tibld[i] : -999.0;
      }
      o[26] = Tievolve;
      for (int i = 0; i < SUEWS_NCOL_ESTM; i++)   // the module's own set_nan, pNAN = 9999 (estm.f95:1582, :1608-1621)
         if (fabs(o[i]) > 9999) o[i] = -999;
   }
   E.Tair2 = Tair1;
   return QS;
}
