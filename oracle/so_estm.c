/* ORACLE (test infrastructure only -- see suews_oracle.h).
 *
 * ESTM, the Element Surface Temperature Method for the storage heat flux (StorageHeatMethod 4/14):
 * SUEWS-SourceCode/suews_phys_estm.f95 -- NewtonPolynomial :206-250, heatcond1d :480-510, the first-call half
 * of ESTM_translate :665-716 / :941-960, its every-call half :718-939, and ESTM itself :970-1605 -- with
 * HEATCAPACITY_AIR and the helpers it calls from suews_util_meteo.f95:25-39, :62-93, :103-113.
 *
 * PARITY UNPINNED: the reference repository holds ESTM inputs (ESTMinput.nml, SUEWS_ESTMCoefficients.txt,
 * *_ESTM_Ts_data_*.txt) but no ESTM output of any run, and the reference cannot be compiled here.  This file is
 * a restatement checked by reading only; the CUDA path is tested against it.
 *
 * Reference behaviour kept on purpose:
 *  - SUEWS_cal_Qs sits inside the surface-temperature DO WHILE of SUEWS_cal_Main (driver.f95:642, :776), and
 *    ESTM saves its element temperatures at the end of every call (:1587-1603): the conduction scheme is
 *    stepped once per TRIP of that loop, not once per timestep.
 *  - radforce / groundradforce are hard-coded .FALSE. (:1191-1192), so the external radiation terms
 *    (Rs_*, Rl_ground/wall/roof, kup_estm, alb_avg, RN_*, EM_EQUIV, QHestm, QFBld) reach neither the state nor
 *    an output; they are not formed here.  The sky/wall view factors of ESTM_translate feed only those terms.
 *  - `converged` of NewtonPolynomial and HVAC / Tair2Set of ESTM are SAVE'd.
 *  - Tstep/3600.0 and 4.*Tstep/3600.0 are REAL(4) expressions.
 */
#include "suews_oracle.h"
#include <math.h>

const double *so_ts5_block = 0;

#define C2K F(273.15)        /* suews_ctrl_const.f95:2086 */
#define SBCONST F(5.67051e-8) /* suews_ctrl_const.f95:2087 */

/* suews_util_meteo.f95:84-93 with DENSITY_DRYAIR :110, DENSITY_VAPOR :103, SPEC_HEAT_DRYAIR :62, SPEC_HEAT_VAPOR :68 */
double so_heatcapacity_air(double TK, double RH, double P)
{
   const double R_DRY_MASS = F(8.31451) / F(0.028965); /* meteo.f95:14-15 */
   const double EPSIL = F(0.62197);
   double RHO_D, RHO_V, CPD, CPV, EA, TC, TC_100;
   RHO_D = P * F(100.) / (R_DRY_MASS * TK);
   EA = so_sat_vap_press(TK, P) * RH / F(100.);
   RHO_V = (EA * F(100.) * EPSIL) / (R_DRY_MASS * TK);
   TC = TK - F(273.15);
   CPD = F(1005.) + ((TC + F(23.15)) * (TC + F(23.15))) / F(3364.);
   TC_100 = (TK - F(273.15)) / F(100.);
   CPV = F(1859.) + F(0.13) * RH + (F(19.3) + F(0.569) * RH) * TC_100 + (F(10.) + F(0.5) * RH) * (TC_100 * TC_100);
   return RHO_D * CPD + RHO_V * CPV;
}

/* x**m for a run-time integer m >= 0: gfortran calls libgcc's __powidf2 (square-and-multiply, LSB first) */
static double powi(double x, int m)
{
   unsigned n = (unsigned)m;
   double y = (n & 1u) ? x : 1.0;
   while (n >>= 1) {
      x = x * x;
      if (n & 1u) y = y * x;
   }
   return y;
}

/* estm.f95:206-250; n = SIZE(Pcoeff) = 5 at every call site.  *converged_save is the SAVE'd `converged`. */
double so_NewtonPolynomial(double x0, const double Pcoeff[5], double conv, int maxiter, double *converged_save)
{
   const int n = 5;
   double e = (double)3.40282347e+38f; /* HUGE(1.) */
   double x = x0, xprev, f, fp;
   int i, j;
   for (i = 1; i <= maxiter; i++) {
      if (fabs(e) < conv) {
         *converged_save = 1;
         break;
      }
      f = 0;
      fp = 0;
      for (j = 1; j <= n - 1; j++) {
         f = f + Pcoeff[j - 1] * powi(x, n - j);
         fp = fp + Pcoeff[j - 1] * (double)(n - j) * powi(x, n - j - 1);
      }
      f = f + Pcoeff[n - 1];
      xprev = x;
      if (fp == 0) fp = (double)1.17549435e-38f; /* TINY(1.) */
      x = xprev - f / fp;
      e = x - xprev;
   }
   if (*converged_save == 0) x = x0;
   return x;
}

/* estm.f95:480-510 */
void so_heatcond1d(double *T, double *Qs, const double *dx, double dt, const double *k, const double *rhocp, int n,
                   const double bc[2], const int bctype[2])
{
   double w[6], a[5], T1[5];
   int i;
   for (i = 1; i <= n; i++) w[i] = T[i - 1];
   w[0] = bc[0];
   w[n] = bc[1];
   if (bctype[0]) w[0] = bc[0] * F(0.5) * dx[0] / k[0] + w[1];
   if (bctype[1]) w[n] = bc[1] * F(0.5) * dx[n - 1] / k[n - 1] + w[n];
   for (i = 0; i < n; i++) a[i] = k[i] / dx[i];
   for (i = 1; i <= n - 1; i++) w[i] = (T[i] * a[i] + T[i - 1] * a[i - 1]) / (a[i - 1] + a[i]);
   for (i = 1; i <= n; i++)
      T1[i - 1] = (dt / rhocp[i - 1]) * (w[i - 1] - 2 * T[i - 1] + w[i]) * 2 * a[i - 1] / dx[i - 1] + T[i - 1];
   *Qs = (w[0] - T[0]) * 2 * a[0] + (w[n] - T[n - 1]) * 2 * a[n - 1];
   for (i = 0; i < n; i++) T[i] = T1[i];
}

/* The ESTMStart == 1 half of ESTM_translate (estm.f95:665-716, :941-960): element temperatures interpolated
 * between the first row of the surface-temperature forcing and the indoor / deep-soil temperatures.
 * alb: surface albedos at that moment (em_ground needs only emis). */
void so_ESTM_translate_start(const SoParams *p, const double alb[NSURF], const double Ts[13], SoEstmState *e)
{
   const SoEstmParams *q = &p->estm;
   const double *sfr = p->sfr;
   const double em_wall = F(0.9); /* estm.f95:169 */
   double fground, em_ground, em_roof;
   int i, w;
   /* cTs_Tiair = 5, cTs_Troof = 7, cTs_Troad = 8, cTs_Twall = 9 (1-based) */
   const double Tiair = Ts[4], Troof_ = Ts[6], Troad = Ts[7], Twall_ = Ts[8];
   (void)alb;
   e->Tievolve = F(20.0) + C2K;
   for (i = 1; i <= q->Nground; i++)
      e->Tground[i - 1] = (q->LBC_soil - Troad) * (i - 1) / (q->Nground - 1) + Troad + C2K;
   for (i = 1; i <= q->Nwall; i++) e->Twall[i - 1] = (Tiair - Twall_) * (i - 1) / (q->Nwall - 1) + Twall_ + C2K;
   for (i = 1; i <= q->Nroof; i++) e->Troof[i - 1] = (Tiair - Troof_) * (i - 1) / (q->Nroof - 1) + Troof_ + C2K;
   for (i = 0; i < q->Nibld; i++) e->Tibld[i] = Tiair + C2K;
   for (w = 0; w < 4; w++)
      for (i = 0; i < q->Nwall; i++) e->Tw_4[i + 5 * w] = e->Twall[i];
   e->T0_ground = e->Tground[0];
   e->T0_wall = e->Twall[0];
   e->T0_roof = e->Troof[0];
   e->T0_ibld = e->Tibld[0];
   e->TN_roof = e->Troof[q->Nroof - 1];
   e->TN_wall = e->Twall[q->Nwall - 1];
   /* suews_ctrl_translate.f95:610-612 and estm.f95:735-745 */
   fground = sfr[PAV] + sfr[CONIF] + sfr[DECID] + sfr[GRASS] + sfr[BSOIL] + sfr[WATER];
   em_roof = p->emis[BLDG];
   if (fground != 0)
      em_ground = (p->emis[CONIF] * sfr[CONIF] + p->emis[DECID] * sfr[DECID] + p->emis[GRASS] * sfr[GRASS] +
                   p->emis[PAV] * sfr[PAV] + p->emis[BSOIL] * sfr[BSOIL] + p->emis[WATER] * sfr[WATER]) /
                  fground;
   else
      em_ground = -999;
   e->lup_ground = SBCONST * em_ground * ((e->T0_ground * e->T0_ground) * (e->T0_ground * e->T0_ground));
   e->lup_wall = SBCONST * em_wall * ((e->T0_wall * e->T0_wall) * (e->T0_wall * e->T0_wall));
   e->lup_roof = SBCONST * em_roof * ((e->T0_roof * e->T0_roof) * (e->T0_roof * e->T0_roof));
   e->Tair2 = 0;
   e->HVAC = 0;
   e->Tair2Set = 0;
   e->NewtonConverged = 0;
}

static double p4(double x)
{
   double a = x * x;
   return a * a;
}

/* estm.f95:970-1605 */
void so_ESTM(SoCtx *cx, const SoParams *p, SoEstmState *e, int tstep, double avkdn, double avu1, double temp_c,
             double zenith_deg, double avrh, double press_hpa, double ldown, double bldgh, const double Ts[13],
             double Tair_av, double o[NCOL_ESTM], double *QS)
{
   const SoEstmParams *q = &p->estm;
   const double *sfr = p->sfr;
   const double conv = F(0.0001);
   const int maxiter = 100;
   const double em_wall = F(0.9);
   const double WSmin = F(0.1);
   const double NAN_ = -999;
   /* ---- ESTM_translate, every-call half (estm.f95:718-939) and suews_ctrl_translate.f95:393-394, :610-614 */
   const double fwall = q->AreaWall / p->SurfaceArea;
   const double froof = sfr[BLDG];
   const double fground = sfr[PAV] + sfr[CONIF] + sfr[DECID] + sfr[GRASS] + sfr[BSOIL] + sfr[WATER];
   const double ZREF = F(2.0) * bldgh;
   const double em_roof = p->emis[BLDG];
   const double em_r = q->em_ibld, em_w = q->em_ibld, em_i = q->em_ibld, em_f = q->em_ibld;
   double em_ground, fibld, finternal, fair;
   double bldgHX, WS, Tair1, Tinternal, Tsurf_all, Troof_in, Troad, Twall_all, Tw_n, Tw_e, Tw_s, Tw_w;
   double SHC_air, shc_airbld, AIREXHR, AIREXDT, Tairmix, Tadd, kdz;
   double CH_ibld = q->CH_ibld, CH_iwall = q->CH_iwall, CH_iroof = q->CH_iroof;
   double Rs_ibld, Rl_ibld, Rs_iwall, Rl_iwall, Rs_iroof, Rl_iroof;
   double Pcoeff[5], bc[2], Qsibld = 0, Qswall = 0, Qsroof = 0, Qsground = 0, Qsair, Qs_4[4];
   double Twallout[5], Troofout[5], Tgroundout[5], Tibldout[5];
   int bctype[2] = {0, 0}, diagnoseTi = 0, i, w;
   (void)cx; (void)avkdn; (void)zenith_deg; (void)ldown;

   if (fground != 0)
      em_ground = (p->emis[CONIF] * sfr[CONIF] + p->emis[DECID] * sfr[DECID] + p->emis[GRASS] * sfr[GRASS] +
                   p->emis[PAV] * sfr[PAV] + p->emis[BSOIL] * sfr[BSOIL] + p->emis[WATER] * sfr[WATER]) /
                  fground;
   else
      em_ground = NAN_;
   if (q->nroom == 0)
      fibld = (floor(bldgh / F(3.1) - F(0.5)) - 1) * froof;
   else
      fibld = (F(2.) - F(2.) / q->nroom) * fwall + (floor(bldgh / F(3.1) - F(0.5)) - 1) * froof;
   if (fibld == 0) fibld = F(0.00001);
   finternal = froof + fibld + fwall;
   if (finternal == 0) finternal = F(0.00001);
   fair = ZREF - bldgh * froof;
   if (fair == 0) fair = F(0.00001);

   /* ---- ESTM */
   bldgHX = (bldgh > F(0.001)) ? bldgh : F(0.001);
   WS = avu1;
   if (WS < WSmin) WS = WSmin;
   Tair1 = temp_c + C2K;
   e->Tair2Set = e->Tair2Set + 1; /* Gridiv == 1 */
   if (e->Tair2Set == 1) e->Tair2 = temp_c + C2K;
   /* else: the *_grids copies are the state itself */

   Tinternal = Ts[4]; Tsurf_all = Ts[5]; Troof_in = Ts[6]; Troad = Ts[7]; Twall_all = Ts[8];
   Tw_n = Ts[9]; Tw_e = Ts[10]; Tw_s = Ts[11]; Tw_w = Ts[12];

   SHC_air = so_heatcapacity_air(Tair1, avrh, press_hpa);

   switch (q->evolveTibld) {
   case 0:
      diagnoseTi = 0;
      e->HVAC = 0;
      break;
   case 1:
      diagnoseTi = 1;
      if (e->Tievolve > q->THEAT_OFF) e->HVAC = 0;
      else if (e->Tievolve < q->THEAT_ON) e->HVAC = 1;
      break;
   case 2:
      diagnoseTi = 1;
      break;
   default:   /* diagnoseTi is SAVE'd .FALSE. and no CASE assigns it */
      break;
   }

   if (Tair_av > F(20.) + C2K && e->Tievolve > F(25.) + C2K && Tair1 < e->Tievolve && !(e->HVAC != 0))
      AIREXHR = F(2.0);
   else if (Tair_av < F(17.) + C2K || e->HVAC != 0)
      AIREXHR = F(0.5);
   else
      AIREXHR = F(1.0);
   AIREXDT = AIREXHR * (double)((float)tstep / 3600.0f);
   shc_airbld = so_heatcapacity_air(e->Tievolve, avrh, press_hpa);
   if (!(shc_airbld > F(0.00001))) shc_airbld = F(0.00001);   /* MAX(..., 0.00001) */

   if (q->ibldCHmod == 1) {
      CH_ibld = F(1.31) * pow(fabs(e->T0_ibld - e->Tievolve), F(0.25)) / shc_airbld;
      CH_iwall = F(1.31) * pow(fabs(e->TN_wall - e->Tievolve), F(0.25)) / shc_airbld;
      CH_iroof = F(1.52) * pow(fabs(e->TN_roof - e->Tievolve), F(0.25)) / shc_airbld;
      if (fabs(e->TN_roof - e->Tievolve) > 0) CH_iroof = CH_iroof * F(0.39);
   } else if (q->ibldCHmod == 2) {
      CH_ibld = F(1.823) * pow(fabs(e->T0_ibld - e->Tievolve), F(0.293)) / shc_airbld;
      CH_iwall = F(1.823) * pow(fabs(e->TN_wall - e->Tievolve), F(0.293)) / shc_airbld;
      CH_iroof = F(2.175) * pow(fabs(e->TN_roof - e->Tievolve), F(0.308)) / shc_airbld;
      if (fabs(e->TN_roof - e->Tievolve) > 0)
         CH_iroof = F(0.704) * pow(fabs(e->TN_roof - e->Tievolve), F(0.133)) / shc_airbld;
   }

   Tairmix = (e->Tievolve + Tair1 * AIREXDT) / (F(1.0) + AIREXDT);
   e->Tievolve = Tairmix + tstep / bldgHX / finternal *
                              (CH_ibld * fibld * (e->T0_ibld - e->Tievolve) + CH_iroof * froof * (e->TN_roof - e->Tievolve) +
                               CH_iwall * fwall * (e->TN_wall - e->Tievolve));
   if (!diagnoseTi) e->Tievolve = Tinternal + C2K;
   if (e->HVAC != 0) {
      const float lim4 = 4.f * (float)tstep / 3600.0f;
      const double d = q->THEAT_fix - e->Tievolve;
      Tadd = (copysign(1.0, d) + q->THEAT_fix - e->Tievolve) * (double)((lim4 < 0.9f) ? lim4 : 0.9f);
      e->Tievolve = e->Tievolve + Tadd;
   }

   /* internal longwave exchange (estm.f95:1404-1417) */
   Rs_ibld = 0;
   Rl_ibld = SBCONST * (q->ivf_iw * em_w * p4(e->TN_wall) + q->ivf_ir * em_r * p4(e->TN_roof) + q->ivf_if * em_f * p4(q->Tfloor));
   Rs_iwall = 0;
   Rl_iwall = SBCONST * (q->ivf_wi * em_i * p4(e->T0_ibld) + q->ivf_wr * em_r * p4(e->TN_roof) + q->ivf_wf * em_f * p4(q->Tfloor));
   Rs_iroof = 0;
   Rl_iroof = SBCONST * (q->ivf_ri * em_i * p4(e->T0_ibld) + q->ivf_rw * em_w * p4(e->TN_wall) + q->ivf_rf * em_f * p4(q->Tfloor));

   /* ========> INTERNAL */
   kdz = 2 * q->kibld[0] / q->zibld[0];
   Pcoeff[0] = q->em_ibld * SBCONST * (1 - q->ivf_ii * q->em_ibld);
   Pcoeff[1] = 0; Pcoeff[2] = 0;
   Pcoeff[3] = kdz + shc_airbld * CH_ibld;
   Pcoeff[4] = -kdz * e->Tibld[0] - shc_airbld * CH_ibld * e->Tievolve - Rs_ibld - Rl_ibld;
   e->T0_ibld = so_NewtonPolynomial(e->T0_ibld, Pcoeff, conv, maxiter, &e->NewtonConverged);
   bc[0] = e->T0_ibld;
   bc[1] = bc[0];
   so_heatcond1d(e->Tibld, &Qsibld, q->zibld, (double)tstep, q->kibld, q->ribld, q->Nibld, bc, bctype);

   /* ========> WALLS */
   kdz = 2 * q->kwall[q->Nwall - 1] / q->zwall[q->Nwall - 1];
   Pcoeff[0] = q->em_ibld * SBCONST * (1 - q->ivf_ww * q->em_ibld);
   Pcoeff[3] = kdz + shc_airbld * CH_iwall;
   Pcoeff[4] = -kdz * e->Twall[q->Nwall - 1] - shc_airbld * CH_iwall * e->Tievolve - Rs_iwall - Rl_iwall;
   e->TN_wall = so_NewtonPolynomial(e->TN_wall, Pcoeff, conv, maxiter, &e->NewtonConverged);
   bc[1] = e->TN_wall;
   if (q->TsurfChoice < 2) {
      if (q->TsurfChoice == 0) {
         bc[0] = Tsurf_all + C2K;
         e->T0_wall = bc[0];
      } else if (q->TsurfChoice == 1) {
         bc[0] = Twall_all + C2K;
         e->T0_wall = bc[0];
      }
      so_heatcond1d(e->Twall, &Qswall, q->zwall, (double)tstep, q->kwall, q->rwall, q->Nwall, bc, bctype);
   } else if (q->TsurfChoice == 2) {
      double s;
      e->T0_wall = 0;
      for (w = 0; w < 4; w++) {
         bc[0] = Tw_n + Tw_e + Tw_s + Tw_w + C2K;
         e->T0_wall = e->T0_wall + bc[0];
         so_heatcond1d(&e->Tw_4[5 * w], &Qs_4[w], q->zwall, (double)tstep, q->kwall, q->rwall, q->Nwall, bc, bctype);
      }
      e->T0_wall = e->T0_wall / F(4.);
      s = 0;
      for (w = 0; w < 4; w++) s = s + Qs_4[w];
      Qswall = s / F(4.);
      for (i = 0; i < q->Nwall; i++) {
         s = 0;
         for (w = 0; w < 4; w++) s = s + e->Tw_4[i + 5 * w];
         e->Twall[i] = s / F(4.);
      }
   }

   /* ========> ROOF */
   kdz = 2 * q->kroof[q->Nroof - 1] / q->zroof[q->Nroof - 1];
   Pcoeff[0] = q->em_ibld * SBCONST;
   Pcoeff[3] = kdz + shc_airbld * CH_iroof;
   Pcoeff[4] = -kdz * e->Troof[q->Nroof - 1] - shc_airbld * CH_iroof * e->Tievolve - Rs_iroof - Rl_iroof;
   e->TN_roof = so_NewtonPolynomial(e->TN_roof, Pcoeff, conv, maxiter, &e->NewtonConverged);
   bc[1] = e->TN_roof;
   if (q->TsurfChoice == 0) {
      bc[0] = Tsurf_all + C2K;
      e->T0_roof = bc[0];
   } else {
      bc[0] = Troof_in + C2K;
      e->T0_roof = bc[0];
   }
   so_heatcond1d(e->Troof, &Qsroof, q->zroof, (double)tstep, q->kroof, q->rroof, q->Nroof, bc, bctype);

   /* ========> GROUND */
   if (q->TsurfChoice == 0) {
      bc[0] = Tsurf_all + C2K;
      e->T0_ground = bc[0];
   } else {
      bc[0] = Troad + C2K;
      e->T0_ground = bc[0];
   }
   bc[1] = q->LBC_soil + C2K;
   if (fground != 0)
      so_heatcond1d(e->Tground, &Qsground, q->zground, (double)tstep, q->kground, q->rground, q->Nground, bc, bctype);
   else
      Qsground = NAN_;

   Qsair = fair * SHC_air * (Tair1 - e->Tair2) / tstep;
   Qsibld = Qsibld * fibld;
   Qswall = Qswall * fwall;
   Qsroof = Qsroof * froof;
   Qsground = Qsground * fground;
   *QS = Qsibld + Qswall + Qsroof + Qsground;

   e->lup_ground = SBCONST * em_ground * p4(e->T0_ground);
   e->lup_wall = SBCONST * em_wall * p4(e->T0_wall);
   e->lup_roof = SBCONST * em_roof * p4(e->T0_roof);

   for (i = 0; i < 5; i++) {
      Twallout[i] = (i < q->Nwall) ? e->Twall[i] : -999.0;
      Troofout[i] = (i < q->Nroof) ? e->Troof[i] : -999.0;
      Tgroundout[i] = (i < q->Nground) ? e->Tground[i] : -999.0;
      Tibldout[i] = (i < q->Nibld) ? e->Tibld[i] : -999.0;
   }
   o[0] = *QS; o[1] = Qsair; o[2] = Qswall; o[3] = Qsroof; o[4] = Qsground; o[5] = Qsibld;
   for (i = 0; i < 5; i++) {
      o[6 + i] = Twallout[i];
      o[11 + i] = Troofout[i];
      o[16 + i] = Tgroundout[i];
      o[21 + i] = Tibldout[i];
   }
   o[26] = e->Tievolve;
   for (i = 0; i < NCOL_ESTM; i++)   /* estm.f95:1582 with the module's own set_nan (:1608-1621, pNAN = 9999) */
      if (fabs(o[i]) > 9999) o[i] = -999;

   e->Tair2 = Tair1;
}
