// dev_anohm.cuh -- AnOHM, the Analytical Objective Hysteresis Model (StorageHeatMethod 3), on the device.
//
// Reference: SUEWS-SourceCode/suews_phys_anohm.f95 (AnOHM :28-143, AnOHM_coef :149-283, AnOHM_xTs :290-360,
// AnOHM_coef_land_cal :363-515, AnOHM_coef_water_cal :520-663, AnOHM_Fc / FcLoad / FcCal :667-936, AnOHM_ShapeFit
// :940-998, fSin :1005-1032, AnOHM_Bo_cal :1037-1166, fcnBo :1170-1395, esat_fn / qsat_fn / qa_fn :1400-1449), called from
// SUEWS_cal_Qs (suews_ctrl_driver.f95:1577-1594), with the two MINPACK drivers it uses (suews_util_minpack.f95: lmdif1
// :3477 -> lmdif :2942, fdjac2 :740, qrfac :4743, lmpar :3617, qrsolv :4937; hybrd1 :1413 -> hybrd :862, fdjac1 :560, qform
// :4642, dogleg :171, r1updt :5266, r1mpyq :5155).
//
// How it sits in the kernel.  The scheme is a once-a-day fit: on the first step of a day (the first trip of its
// surface-temperature loop) the day's on-the-hour forcing rows give sinusoid fits of Sd and Ta (Levenberg-Marquardt, three
// unknowns, <= 24 residuals), a Bowen ratio per surface (Powell hybrid, one unknown) and from those the seven surfaces'
// a1, a2, a3; every other step of the day reloads them (the reference keeps them in SAVE'd locals, here they are a
// per-cell state record, SUEWS_ANOHM_STATE_FIELDS) and QS = a1 Qn + a2 dQn/dt + a3 as in OHM.  So the fit is a cold,
// out-of-line function: 1 step in 288 enters it, and its work arrays live in its own frame.
//   * MetForcingData_grid -- the host's whole forcing block of the grid (suews_ctrl_translate.f95:1216) -- is the staged
//     forcing block.  The host side (suews_b200.cu, build_anohm_hours) keeps its on-the-hour rows grouped by day of year,
//     in block order, so a cell reads the <= 24 rows of its day directly; the cell's forcing offsets (dTemp_C, dRH) are
//     applied to them exactly as the step applies them to its own row.
//   * lmdif is written for its one use (n = 3 unknowns, pivoting kept); hybrd for n = 1, where QR factorisation, the
//     dogleg step and the rank-one update reduce to scalars -- every surviving operation is kept in the reference's order,
//     so the iterates are the reference's bit for bit wherever libm agrees (checked against scipy's MINPACK and the
//     oracle's general-n restatement on the host, tests/test_anohm_oracle.py).
//   * the reference's quirks are kept: PI is the REAL(4) pi, MIN(Ts - 273.15, -40.) in the Bowen residual, QS formed with
//     dqndt_prev (anohm.f95:139), hybrd1's tolerance 1E-20, the cache marked valid by the water surface only.
// Each cell behaves as a single-grid run of the reference (Gridiv constant): with several grids the reference's
// cache (one grid_save for the whole process) misses on every call and the fit is redone every timestep with that
// step's qf and wetness; a batch of cells has no such coupling between cells.
//
// SU_ANOHM_HOST: the same code compiled for the host by the CPU test (tests/host_anohm.cpp), which compares it with the
// oracle without a GPU.
#pragma once
#ifdef SU_ANOHM_HOST
#include <math.h>
#define AN_FN static inline
#define AN_COLD static
#ifndef F
#define F(c) ((double)(c##f))
#endif
#else
#define AN_FN SU_INL
#define AN_COLD SU_COLD
#endif

#define AN_MAXM 24   // on-the-hour rows of one day of year in the forcing block
#define AN_EPSMCH 2.220446049250313e-16    // epsilon(1d0)
#define AN_DWARF 2.2250738585072014e-308   // tiny(1d0)
#define AN_PI F(3.14159274)                        // PI = ATAN(1.0)*4, a REAL(4) value (anohm.f95:329)
#define AN_OMEGA ((2 * AN_PI) / (24 * 60 * 60))

// One on-the-hour row of MetForcingData_grid: columns 2, 3, 15, 12, 11, 13, 10, 9 (anohm.f95:752-753)
struct SuewsHourRec {
   double id, tHr, Sd, Ta, RH, pres, WS, qf_obs;
};
// The staged forcing block as AnOHM reads it: hours[] holds the rows with imin == 0 grouped by day of year (block order
// inside a day); day d occupies [day_first[d], day_first[d] + day_count[d]); day_nsun[d] counts its rows with Sd > 0.
struct SuAnohmIn {
   const SuewsHourRec *hours;
   const int *day_first, *day_count, *day_nsun;   // [367], index = day of year (0 unused)
   // (the record also carries BLUEWS's shared table: IniCBLdata [cbl_nline][8] row-major, dev_bluews.cuh)
   const double *cbl_ini;
   int cbl_nline;
};

struct AnDay {
   int n;         // rows of the day
   int lenDay;    // of which Sd > 5
   double tHr[AN_MAXM], Sd[AN_MAXM], Ta[AN_MAXM], RH[AN_MAXM], pres[AN_MAXM];   // the daytime rows (Sd > 5), packed
   double sumWS, sumAH;   // SUM over the daytime rows, in row order
};
struct AnFc {
   double ASd, mSd, tSd, ATa, mTa, tTa, tau, mWS, mWF, mAH;
};

AN_FN double an_enorm3(const double *x)
{
   double s = 0;
   for (int i = 0; i < 3; i++) s = s + x[i] * x[i];
   return sqrt(s);
}
AN_FN double an_enorm(int n, const double *x)
{
   double s = 0;
   for (int i = 0; i < n; i++) s = s + x[i] * x[i];
   return sqrt(s);
}

// fSin (anohm.f95:1005-1032)
AN_FN void an_fsin(int m, const double *x, const double *xdat, const double *ydat, double *fvec)
{
   for (int i = 0; i < m; i++) fvec[i] = x[0] + x[1] * sin(2 * AN_PI / 24 * (xdat[i] - x[2])) - ydat[i];
}

// qrsolv for n = 3 (minpack.f95:4937-5154); r is [3][3] column-major: r[j][i] = r(i, j)
AN_FN void an_qrsolv3(double (*r)[3], const int *ipvt, const double *diag, const double *qtb, double *x, double *sdiag)
{
   double wa[3];
   for (int j = 0; j < 3; j++) {
      for (int i = j; i < 3; i++) r[j][i] = r[i][j];
      x[j] = r[j][j];
   }
   for (int j = 0; j < 3; j++) wa[j] = qtb[j];
   for (int j = 0; j < 3; j++) {
      const int l = ipvt[j];
      if (diag[l] != 0.0) {
         for (int k = j; k < 3; k++) sdiag[k] = 0.0;
         sdiag[j] = diag[l];
         double qtbpj = 0.0;
         for (int k = j; k < 3; k++) {
            if (sdiag[k] != 0.0) {
               double c, s;
               if (fabs(r[k][k]) < fabs(sdiag[k])) {
                  const double cotan = r[k][k] / sdiag[k];
                  s = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                  c = s * cotan;
               } else {
                  const double t = sdiag[k] / r[k][k];
                  c = 0.5 / sqrt(0.25 + 0.25 * (t * t));
                  s = c * t;
               }
               r[k][k] = c * r[k][k] + s * sdiag[k];
               double temp = c * wa[k] + s * qtbpj;
               qtbpj = -s * wa[k] + c * qtbpj;
               wa[k] = temp;
               for (int i = k + 1; i < 3; i++) {
                  temp = c * r[k][i] + s * sdiag[i];
                  sdiag[i] = -s * r[k][i] + c * sdiag[i];
                  r[k][i] = temp;
               }
            }
         }
      }
      sdiag[j] = r[j][j];
      r[j][j] = x[j];
   }
   int nsing = 3;
   for (int j = 0; j < 3; j++) {
      if (sdiag[j] == 0.0 && nsing == 3) nsing = j;
      if (nsing < 3) wa[j] = 0.0;
   }
   for (int j = nsing - 1; j >= 0; j--) {
      double sum2 = 0;
      for (int i = j + 1; i < nsing; i++) sum2 = sum2 + wa[i] * r[j][i];
      wa[j] = (wa[j] - sum2) / sdiag[j];
   }
   for (int j = 0; j < 3; j++) x[ipvt[j]] = wa[j];
}

// lmpar for n = 3 (minpack.f95:3617-3920)
AN_FN void an_lmpar3(double (*r)[3], const int *ipvt, const double *diag, const double *qtb, double delta, double &par,
                     double *x, double *sdiag)
{
   double wa1[3], wa2[3];
   int nsing = 3;
   for (int j = 0; j < 3; j++) {
      wa1[j] = qtb[j];
      if (r[j][j] == 0.0 && nsing == 3) nsing = j;
      if (nsing < 3) wa1[j] = 0.0;
   }
   for (int k = 1; k <= nsing; k++) {
      const int j = nsing - k;
      wa1[j] = wa1[j] / r[j][j];
      const double temp = wa1[j];
      for (int i = 0; i < j; i++) wa1[i] = wa1[i] - r[j][i] * temp;
   }
   for (int j = 0; j < 3; j++) x[ipvt[j]] = wa1[j];
   int iter = 0;
   for (int j = 0; j < 3; j++) wa2[j] = diag[j] * x[j];
   double dxnorm = an_enorm3(wa2);
   double fp = dxnorm - delta;
   if (fp <= 0.1 * delta) {
      par = 0.0;
      return;
   }
   double parl = 0.0;
   if (3 <= nsing) {
      for (int j = 0; j < 3; j++) {
         const int l = ipvt[j];
         wa1[j] = diag[l] * (wa2[l] / dxnorm);
      }
      for (int j = 0; j < 3; j++) {
         double sum2 = 0;
         for (int i = 0; i < j; i++) sum2 = sum2 + wa1[i] * r[j][i];
         wa1[j] = (wa1[j] - sum2) / r[j][j];
      }
      const double temp = an_enorm3(wa1);
      parl = ((fp / delta) / temp) / temp;
   }
   for (int j = 0; j < 3; j++) {
      double sum2 = 0;
      for (int i = 0; i <= j; i++) sum2 = sum2 + qtb[i] * r[j][i];
      wa1[j] = sum2 / diag[ipvt[j]];
   }
   const double gnorm = an_enorm3(wa1);
   double paru = gnorm / delta;
   if (paru == 0.0) paru = AN_DWARF / ((delta < 0.1) ? delta : 0.1);
   par = (par > parl) ? par : parl;
   par = (par < paru) ? par : paru;
   if (par == 0.0) par = gnorm / dxnorm;
   for (;;) {
      iter = iter + 1;
      if (par == 0.0) par = (AN_DWARF > 0.001 * paru) ? AN_DWARF : 0.001 * paru;
      const double sq = sqrt(par);
      for (int j = 0; j < 3; j++) wa1[j] = sq * diag[j];
      an_qrsolv3(r, ipvt, wa1, qtb, x, sdiag);
      for (int j = 0; j < 3; j++) wa2[j] = diag[j] * x[j];
      dxnorm = an_enorm3(wa2);
      double temp = fp;
      fp = dxnorm - delta;
      if (fabs(fp) <= 0.1 * delta) break;
      if (parl == 0.0 && fp <= temp && temp < 0.0) break;
      if (iter == 10) break;
      for (int j = 0; j < 3; j++) {
         const int l = ipvt[j];
         wa1[j] = diag[l] * (wa2[l] / dxnorm);
      }
      for (int j = 0; j < 3; j++) {
         wa1[j] = wa1[j] / sdiag[j];
         temp = wa1[j];
         for (int i = j + 1; i < 3; i++) wa1[i] = wa1[i] - r[j][i] * temp;
      }
      temp = an_enorm3(wa1);
      const double parc = ((fp / delta) / temp) / temp;
      if (0.0 < fp) parl = (parl > par) ? parl : par;
      else if (fp < 0.0) paru = (paru < par) ? paru : par;
      par = (parl > par + parc) ? parl : par + parc;
   }
}

// lmdif1 (minpack.f95:3477-3616) for the sinusoid fit: n = 3, m <= AN_MAXM, tol = 1e-5.  fjac[j][i] = fjac(i, j); after
// qrfac its upper 3 x 3 corner is what lmpar / qrsolv work on (they get it as a separate [3][3] array: the rows below the
// third are the Householder vectors, which lmpar never touches).
AN_COLD void an_lmdif1_sin(int m, const double *xdat, const double *ydat, double *x)
{
   const double tol = 0.00001;
   const double factor = 100.0, ftol = tol, xtol = tol;
   const int maxfev = 200 * (3 + 1);
   double fjac[3][AN_MAXM], fvec[AN_MAXM], wa4[AN_MAXM];
   double r3[3][3];
   double diag[3], qtf[3], wa1[3], wa2[3], wa3[3], rdiag[3], acnorm[3], wap[3];
   int ipvt[3];
   if (m < 3) return;
   an_fsin(m, x, xdat, ydat, fvec);
   int nfev = 1;
   double fnorm = an_enorm(m, fvec);
   double par = 0.0, delta = 0, xnorm = 0, temp = 0;
   int iter = 1;
   for (;;) {
      // fdjac2 (minpack.f95:740-860), epsfcn = 0
      const double eps = sqrt(AN_EPSMCH);
      for (int j = 0; j < 3; j++) {
         temp = x[j];
         double h = eps * fabs(temp);
         if (h == 0.0) h = eps;
         x[j] = temp + h;
         an_fsin(m, x, xdat, ydat, wa4);
         x[j] = temp;
         for (int i = 0; i < m; i++) fjac[j][i] = (wa4[i] - fvec[i]) / h;
      }
      nfev = nfev + 3;
      // qrfac with pivoting (minpack.f95:4743-4936)
      for (int j = 0; j < 3; j++) {
         acnorm[j] = an_enorm(m, fjac[j]);
         rdiag[j] = acnorm[j];
         wap[j] = acnorm[j];
         ipvt[j] = j;
      }
      for (int j = 0; j < 3; j++) {
         int kmax = j;
         for (int k = j; k < 3; k++)
            if (rdiag[kmax] < rdiag[k]) kmax = k;
         if (kmax != j) {
            for (int i = 0; i < m; i++) {
               const double t = fjac[j][i];
               fjac[j][i] = fjac[kmax][i];
               fjac[kmax][i] = t;
            }
            rdiag[kmax] = rdiag[j];
            wap[kmax] = wap[j];
            const int it = ipvt[j];
            ipvt[j] = ipvt[kmax];
            ipvt[kmax] = it;
         }
         double ajnorm = an_enorm(m - j, &fjac[j][j]);
         if (ajnorm != 0.0) {
            if (fjac[j][j] < 0.0) ajnorm = -ajnorm;
            for (int i = j; i < m; i++) fjac[j][i] = fjac[j][i] / ajnorm;
            fjac[j][j] = fjac[j][j] + 1.0;
            for (int k = j + 1; k < 3; k++) {
               double s = 0;
               for (int i = j; i < m; i++) s = s + fjac[j][i] * fjac[k][i];
               temp = s / fjac[j][j];
               for (int i = j; i < m; i++) fjac[k][i] = fjac[k][i] - temp * fjac[j][i];
               if (rdiag[k] != 0.0) {
                  temp = fjac[k][j] / rdiag[k];
                  const double t2 = 1.0 - temp * temp;
                  rdiag[k] = rdiag[k] * sqrt((t2 > 0.0) ? t2 : 0.0);
                  const double q = rdiag[k] / wap[k];
                  if (0.05 * (q * q) <= AN_EPSMCH) {
                     rdiag[k] = an_enorm(m - j - 1, &fjac[k][j + 1]);
                     wap[k] = rdiag[k];
                  }
               }
            }
         }
         rdiag[j] = -ajnorm;
      }
      // (wa1 = rdiag, wa2 = acnorm in lmdif's naming)
      if (iter == 1) {
         for (int j = 0; j < 3; j++) {
            diag[j] = acnorm[j];
            if (acnorm[j] == 0.0) diag[j] = 1.0;
         }
         for (int j = 0; j < 3; j++) wa3[j] = diag[j] * x[j];
         xnorm = an_enorm3(wa3);
         delta = factor * xnorm;
         if (delta == 0.0) delta = factor;
      }
      for (int i = 0; i < m; i++) wa4[i] = fvec[i];
      for (int j = 0; j < 3; j++) {
         if (fjac[j][j] != 0.0) {
            double sum2 = 0;
            for (int i = j; i < m; i++) sum2 = sum2 + wa4[i] * fjac[j][i];
            temp = -sum2 / fjac[j][j];
            for (int i = j; i < m; i++) wa4[i] = wa4[i] + fjac[j][i] * temp;
         }
         fjac[j][j] = rdiag[j];
         qtf[j] = wa4[j];
      }
      double gnorm = 0.0;
      if (fnorm != 0.0) {
         for (int j = 0; j < 3; j++) {
            const int l = ipvt[j];
            if (acnorm[l] != 0.0) {
               double sum2 = 0.0;
               for (int i = 0; i <= j; i++) sum2 = sum2 + fjac[j][i] * (qtf[i] / fnorm);
               const double g = fabs(sum2 / acnorm[l]);
               gnorm = (gnorm > g) ? gnorm : g;
            }
         }
      }
      if (gnorm <= 0.0) return;   // gtol = 0: info 4
      for (int j = 0; j < 3; j++) diag[j] = (diag[j] > acnorm[j]) ? diag[j] : acnorm[j];
      for (;;) {
         // lmpar / qrsolv overwrite the strict lower triangle of the 3 x 3 corner and restore its diagonal; the upper
         // triangle (R) is what the prediction below reads, and they leave it alone
         for (int j = 0; j < 3; j++)
            for (int i = 0; i < 3; i++) r3[j][i] = fjac[j][i];
         an_lmpar3(r3, ipvt, diag, qtf, delta, par, wa1, wa2);
         for (int j = 0; j < 3; j++)
            for (int i = 0; i < 3; i++) fjac[j][i] = r3[j][i];
         for (int j = 0; j < 3; j++) wa1[j] = -wa1[j];
         for (int j = 0; j < 3; j++) wa2[j] = x[j] + wa1[j];
         for (int j = 0; j < 3; j++) wa3[j] = diag[j] * wa1[j];
         const double pnorm = an_enorm3(wa3);
         if (iter == 1) delta = (delta < pnorm) ? delta : pnorm;
         an_fsin(m, wa2, xdat, ydat, wa4);
         nfev = nfev + 1;
         const double fnorm1 = an_enorm(m, wa4);
         double actred;
         if (0.1 * fnorm1 < fnorm) {
            const double q = fnorm1 / fnorm;
            actred = 1.0 - q * q;
         } else {
            actred = -1.0;
         }
         for (int j = 0; j < 3; j++) {
            wa3[j] = 0.0;
            temp = wa1[ipvt[j]];
            for (int i = 0; i <= j; i++) wa3[i] = wa3[i] + fjac[j][i] * temp;
         }
         const double temp1 = an_enorm3(wa3) / fnorm;
         const double temp2 = (sqrt(par) * pnorm) / fnorm;
         const double prered = temp1 * temp1 + (temp2 * temp2) / 0.5;
         const double dirder = -(temp1 * temp1 + temp2 * temp2);
         double ratio = 0.0;
         if (prered != 0.0) ratio = actred / prered;
         if (ratio <= 0.25) {
            if (actred >= 0.0) temp = 0.5;
            if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
            if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
            delta = temp * ((delta < pnorm / 0.1) ? delta : pnorm / 0.1);
            par = par / temp;
         } else {
            if (par == 0.0 || ratio >= 0.75) {
               delta = 2.0 * pnorm;
               par = 0.5 * par;
            }
         }
         if (0.0001 <= ratio) {
            for (int j = 0; j < 3; j++) x[j] = wa2[j];
            for (int j = 0; j < 3; j++) wa2[j] = diag[j] * x[j];
            for (int i = 0; i < m; i++) fvec[i] = wa4[i];
            xnorm = an_enorm3(wa2);
            fnorm = fnorm1;
            iter = iter + 1;
         }
         int info = 0;
         if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) info = 1;
         if (delta <= xtol * xnorm) info = 2;
         if (info != 0) return;
         if (maxfev <= nfev) info = 5;
         if (fabs(actred) <= AN_EPSMCH && prered <= AN_EPSMCH && 0.5 * ratio <= 1.0) info = 6;
         if (delta <= AN_EPSMCH * xnorm) info = 7;
         if (gnorm <= AN_EPSMCH) info = 8;
         if (info != 0) return;
         if (!(ratio < 0.0001)) break;
      }
   }
}

// AnOHM_ShapeFit (anohm.f95:940-998)
AN_FN void an_shapefit(int m, const double *tHr, const double *obs, double &amp, double &mean, double &tpeak)
{
   double x[3] = {mean, amp, tpeak};
   an_lmdif1_sin(m, tHr, obs, x);
   mean = x[0];
   amp = x[1];
   tpeak = x[2] + 6;
   if (amp < 0) {
      amp = fabs(amp);
      tpeak = x[2] - 12 + 6 + 24;
      tpeak = fmod(tpeak, F(24.));
   }
   tpeak = fmod(tpeak, F(24.));
}

// AnOHM_coef_land_cal (anohm.f95:363-515)
AN_COLD void an_coef_land(const AnFc &fc, double xalb, double xemis, double xcp, double xk, double xch, double xBo, double &xa1,
                          double &xa2, double &xa3, double &ATs, double &mTs, double &gamma)
{
   const double SIGMA = F(5.67e-8), OMEGA = AN_OMEGA;
   const double ASd = fc.ASd, mSd = fc.mSd, ATa = fc.ATa, mTa = fc.mTa, tau = fc.tau, mWF = fc.mWF;
   const double xchWS = xch * fc.mWS;
   double beta;
   if (fabs(xBo) < F(0.001)) beta = (double)(1.f / 0.001f);
   else beta = 1 / xBo;
   const double mTa3 = mTa * mTa * mTa;
   const double f = ((1 + beta) * xchWS + 4 * SIGMA * xemis * mTa3);
   const double fL = 4 * SIGMA * xemis * mTa3;
   const double fT = (1 + beta) * xchWS;
   const double lambda = xk / xcp;
   const double mWF2 = mWF * mWF;
   const double delta = sqrt(F(.5) * (mWF2 + sqrt(mWF2 * mWF2 + 16 * (lambda * lambda) * (OMEGA * OMEGA))));
   const double m = (2 * lambda) / (delta + mWF);
   const double n = delta / OMEGA;
   mTs = (mSd * (1 - xalb) / f) + mTa;
   const double ms = 1 + xk / (f * m);
   const double ns = xk / (f * n);
   const double st = sin(tau), ct = cos(tau);
   double xx1 = f * st * ATa;
   double xx2 = (1 - xalb) * ASd + f * ct * ATa;
   gamma = atan(ns / ms) + atan(xx1 / xx2);
   const double sg = sin(gamma), cg = cos(gamma);
   ATs = -(st * ATa) / (ns * cg - ms * sg);
   xx1 = (ns * cg + sg - ms * sg) * st * ATa * fL;
   xx2 = (xalb - 1) * (ns * cg - ms * sg) * ASd;
   double xx3 = (-ms * ct * st + cg * (ns * ct + st)) * ATa * fL;
   xx2 = xx2 - xx3;
   const double phi = atan(xx1 / xx2);
   xx3 = (ns * cg - ms * sg);
   xx1 = (1 + sg / xx3) * st * ATa * fL;
   xx2 = (xalb - 1) * ASd - (ct + cg * st / xx3) * ATa * fL;
   const double cphi = sqrt(xx1 * xx1 + xx2 * xx2);
   xx1 = m * cg - n * sg;
   xx2 = m * sg + n * cg;
   const double eta = atan(xx1 / xx2);
   xx1 = (xk * xk) * (m * m + n * n) * (ATs * ATs);
   xx2 = (m * m) * (n * n);
   const double ceta = sqrt(xx1 / xx2);
   const double xlag = eta - phi;
   xa1 = (ceta * cos(xlag)) / cphi;
   xa2 = (ceta * sin(xlag)) / (OMEGA * cphi);
   xa2 = xa2 / 3600;
   xa3 = -xa1 * (fT / f) * (mSd * (1 - xalb)) - fc.mAH;
}

// AnOHM_coef_water_cal (anohm.f95:520-663), xeta = 0.3, xmu = 0.2 at both call sites
AN_COLD void an_coef_water(const AnFc &fc, double xalb, double xemis, double xcp, double xk, double xch, double xBo, double &xa1,
                           double &xa2, double &xa3, double &ATs, double &mTs, double &gamma)
{
   const double SIGMA = F(5.67e-8), PI = AN_PI, OMEGA = AN_OMEGA, xeta = F(0.3), xmu = F(0.2);
   const double ASd = fc.ASd, mSd = fc.mSd, ATa = fc.ATa, mTa = fc.mTa, tau = fc.tau;
   const double xm = xk * (xmu * xmu);
   const double xn = xcp * OMEGA;
   const double phi = atan(xn / xm);
   const double kappa = sqrt(xcp * OMEGA / (2 * xk));
   const double xchWS = xch * fc.mWS;
   const double beta = 1 / xBo;
   const double mTa3 = mTa * mTa * mTa;
   const double f = ((1 + beta) * xchWS + 4 * SIGMA * xemis * mTa3);
   const double fL = 4 * SIGMA * xemis * mTa3;
   const double fT = (1 + beta) * xchWS;
   const double calb = 1 - xalb;
   const double lambda = sqrt(xm * xm + xn * xn);
   const double xkk = xk * kappa;
   const double dtau = atan(xkk / (f + xkk));
   const double dpsi = atan((xk - xmu) / kappa);
   const double dphi = atan(kappa * (f + xk * xmu) / (f * (kappa - xmu) + xkk * (2 * kappa - xmu)));
   const double cdtau = sqrt(xkk * xkk + (f + xkk) * (f + xkk));
   const double cdpsi = sqrt((xk - xmu) * (xk - xmu) + kappa * kappa);
   const double cdphi = cdtau * cdpsi;
   mTs = (mSd * (1 - xalb + xeta) / f) + mTa;
   double xx1 = (xk * xeta * xmu * calb * ASd * cdpsi);
   xx1 = xx1 * xx1;
   double xx2 = 2 * lambda * sqrt(xx1) * (calb * ASd * sin(phi - dpsi) + f * ATa * sin(tau + phi - dpsi));
   const double u = (calb * ASd + cos(tau) * f * ATa), v = (sin(tau) * f * ATa);
   double xx3 = (lambda * lambda) * (u * u + v * v);
   ATs = 1 / (cdtau * lambda) * sqrt(xx1 + xx2 + xx3);
   xx1 = (xkk * calb * ASd + cdtau * f * ATa * sin(tau + dtau)) * lambda + (xk * xeta * xmu * calb * ASd * cdphi * sin(phi + dphi));
   xx2 = ((f + xkk) * calb * ASd - cdtau * f * ATa * cos(tau + dtau)) * lambda - xk * xeta * xmu * calb * ASd * cdphi * cos(phi + dphi);
   const double delta = atan(xx1 / xx2);
   gamma = delta;
   xx1 = fL * (ATs * sin(delta) - ATa * sin(tau));
   xx2 = calb * ASd - fL * (ATs * cos(delta) - ATa * cos(tau));
   const double theta = atan(xx1 / xx2);
   const double ctheta = sqrt(xx1 * xx1 + xx2 * xx2);
   const double rt2 = (double)1.41421354f;   // SQRT(2.) in REAL(4)
   const double xxT = rt2 * kappa * lambda * ATs;
   const double xxkappa = cdpsi * xeta * xmu * ASd;
   const double xxdltphi = cos(delta) * sin(dpsi) * cos(phi) - sin(delta) * cos(dpsi) * sin(phi);
   xx1 = xxT * sin(PI / 4 - delta) + xxkappa * sin(phi + dpsi);
   xx2 = xxT * sin(PI / 4 + delta) - xxkappa * sin(phi - dpsi);
   const double zeta = atan(xx1 / xx2);
   xx1 = 2 * rt2 * xxkappa * xxT * xxdltphi;
   xx2 = (1 - cos(2 * dpsi) * cos(2 * phi)) * (xxkappa * xxkappa);
   xx3 = xxT * xxT;
   const double czeta = xk / lambda * sqrt(xx1 + xx2 + xx3);
   const double xlag = zeta - theta;
   xa1 = (czeta * cos(xlag)) / ctheta;
   xa2 = (czeta * sin(xlag)) / (OMEGA * ctheta);
   xa2 = xa2 / 3600;
   xa3 = mSd * (xalb - 1) * (xeta + (fT - fL * xeta) / f * xa1);
}

AN_FN double an_esat(double Ta) { return F(6.106) * exp(F(17.27) * Ta / (F(237.3) + Ta)); }

// fcnBo (anohm.f95:1170-1395): x - Bo(x) over the daytime rows
AN_COLD double an_fcnBo(double x, int sfc_typ, const AnDay &d, const AnFc &fc, double xalb, double xemis, double xcp, double xk,
                        double xch, double xSM_in)
{
   const double C2K = F(273.15), cp_air = F(1006.38), Lv_air = F(2264.705E3), OMEGA = AN_OMEGA;
   const double wratio = F(0.0180153) / F(0.028965);
   double xBo = x;
   const double xSM = (xSM_in < F(1.0)) ? xSM_in : F(1.0);
   if (xSM == 0) {
      xBo = 1000;
   } else {
      // AnOHM_xTs: the coefficients depend on xBo, not on the hour
      double xa1, xa2, xa3, ATs, mTs, gamma;
      if (sfc_typ <= 6) an_coef_land(fc, xalb, xemis, xcp, xk, xch, xBo, xa1, xa2, xa3, ATs, mTs, gamma);
      else an_coef_water(fc, xalb, xemis, xcp, xk, xch, xBo, xa1, xa2, xa3, ATs, mTs, gamma);
      double sQH = 0, sQE = 0;
      for (int i = 0; i < d.lenDay; i++) {
         double Ts = ATs * sin(OMEGA * (d.tHr[i] - fc.tSd + 6) * 3600 - gamma) + mTs;
         Ts = Ts - C2K;
         Ts = (Ts < F(-40.)) ? Ts : F(-40.);
         const double qs = wratio * an_esat(Ts) / d.pres[i];
         const double ea = an_esat(d.Ta[i]) * d.RH[i] / 100;
         const double qa = wratio * ea / d.pres[i];
         sQH = sQH + cp_air * (Ts - d.Ta[i]);
         sQE = sQE + xSM * Lv_air * (qs - qa);
      }
      xBo = sQH / sQE;
   }
   return x - xBo;
}

// AnOHM_Bo_cal (anohm.f95:1037-1166): hybrd1 with one unknown (minpack.f95:1413-1557, :862-1411).  With n = 1 the QR
// factor, the dogleg step and the rank-one update are scalars: fjac is the Householder vector / Q, r the triangular factor.
AN_COLD double an_Bo_cal(int sfc_typ, const AnDay &d, const AnFc &fc, double xalb, double xemis, double xcp, double xk, double xch,
                         double xSM)
{
#define AN_F(xx) an_fcnBo((xx), sfc_typ, d, fc, xalb, xemis, xcp, xk, xch, xSM)
#define AN_NRM(v) sqrt((v) * (v))   // enorm of one element
   const double xtol = (double)1E-20f, factor = 100.0, epsmch = AN_EPSMCH;
   const int maxfev = 200 * (1 + 1);
   double x = 10;
   double fvec = AN_F(x);
   int nfev = 1;
   double fnorm = AN_NRM(fvec);
   int iter = 1, ncsuc = 0, ncfail = 0, nslow1 = 0, nslow2 = 0;
   double delta = 0, xnorm = 0;
   for (;;) {
      bool jeval = true;
      // fdjac1
      const double eps = sqrt(epsmch);
      double temp = x;
      double h = eps * fabs(temp);
      if (h == 0.0) h = eps;
      double fjac = (AN_F(temp + h) - fvec) / h;
      nfev = nfev + 1;
      // qrfac
      const double acnorm = AN_NRM(fjac);
      double ajnorm = AN_NRM(fjac);
      if (ajnorm != 0.0) {
         if (fjac < 0.0) ajnorm = -ajnorm;
         fjac = fjac / ajnorm;
         fjac = fjac + 1.0;
      }
      const double rdiag = -ajnorm;
      (void)acnorm;   // mode = 2: diag stays 1
      if (iter == 1) {
         xnorm = AN_NRM(x);
         delta = factor * xnorm;
         if (delta == 0.0) delta = factor;
      }
      double qtf = fvec;
      if (fjac != 0.0) {
         temp = -(0.0 + qtf * fjac) / fjac;
         qtf = qtf + fjac * temp;
      }
      double r = rdiag;
      // qform
      {
         const double wa = fjac;
         fjac = 1.0;
         if (wa != 0.0) {
            temp = (0.0 + wa * fjac) / wa;
            fjac = fjac - temp * wa;
         }
      }
      for (;;) {
         // dogleg
         temp = r;
         if (temp == 0.0) temp = epsmch;   // (max(0, |r|) = 0, so epsmch)
         double xd = qtf / temp;
         const double qnorm = AN_NRM(xd);
         if (!(qnorm <= delta)) {
            double wa1d = 0.0 + r * qtf;
            const double gnorm = AN_NRM(wa1d);
            double sgnorm = 0.0, alpha = delta / qnorm;
            if (gnorm != 0.0) {
               wa1d = (wa1d / gnorm);
               const double wa2d = 0.0 + r * wa1d;
               temp = AN_NRM(wa2d);
               sgnorm = (gnorm / temp) / temp;
               alpha = 0.0;
               if (sgnorm < delta) {
                  const double bnorm = AN_NRM(qtf);
                  const double dq = delta / qnorm, sd = sgnorm / delta;
                  temp = (bnorm / gnorm) * (bnorm / qnorm) * (sgnorm / delta);
                  const double t1 = temp - dq;
                  temp = temp - dq * (sd * sd) + sqrt(t1 * t1 + (1.0 - dq * dq) * (1.0 - sd * sd));
                  alpha = (dq * (1.0 - sd * sd)) / temp;
               }
            }
            temp = (1.0 - alpha) * ((sgnorm < delta) ? sgnorm : delta);
            xd = temp * wa1d + alpha * xd;
         }
         double wa1 = -xd;
         double wa2 = x + wa1;
         double wa3 = wa1;
         const double pnorm = AN_NRM(wa3);
         if (iter == 1) delta = (delta < pnorm) ? delta : pnorm;
         const double wa4 = AN_F(wa2);
         nfev = nfev + 1;
         const double fnorm1 = AN_NRM(wa4);
         double actred = -1.0;
         if (fnorm1 < fnorm) {
            const double q = fnorm1 / fnorm;
            actred = 1.0 - q * q;
         }
         wa3 = qtf + (0.0 + r * wa1);
         temp = AN_NRM(wa3);
         double prered = 0.0;
         if (temp < fnorm) {
            const double q = temp / fnorm;
            prered = 1.0 - q * q;
         }
         double ratio = 0.0;
         if (0.0 < prered) ratio = actred / prered;
         if (ratio < 0.1) {
            ncsuc = 0;
            ncfail = ncfail + 1;
            delta = 0.5 * delta;
         } else {
            ncfail = 0;
            ncsuc = ncsuc + 1;
            if (0.5 <= ratio || 1 < ncsuc) delta = (delta > pnorm / 0.5) ? delta : pnorm / 0.5;
            if (fabs(ratio - 1.0) <= 0.1) delta = pnorm / 0.5;
         }
         if (0.0001 <= ratio) {
            x = wa2;
            fvec = wa4;
            xnorm = AN_NRM(x);
            fnorm = fnorm1;
            iter = iter + 1;
         }
         nslow1 = nslow1 + 1;
         if (0.001 <= actred) nslow1 = 0;
         if (jeval) nslow2 = nslow2 + 1;
         if (0.1 <= actred) nslow2 = 0;
         if (delta <= xtol * xnorm || fnorm == 0.0) return x;
         int info = 0;
         if (maxfev <= nfev) info = 2;
         const double d01 = 0.1 * delta;
         if (0.1 * ((d01 > pnorm) ? d01 : pnorm) <= epsmch * xnorm) info = 3;
         if (nslow2 == 5) info = 4;
         if (nslow1 == 10) info = 5;
         if (info != 0) return x;
         if (ncfail == 2) break;
         const double sum2 = 0.0 + wa4 * fjac;
         wa2 = (sum2 - wa3) / pnorm;
         wa1 = wa1 / pnorm;
         if (0.0001 <= ratio) qtf = sum2;
         r = r + wa2 * wa1;   // r1updt; r1mpyq is empty for n = 1
         jeval = false;
      }
   }
#undef AN_F
#undef AN_NRM
}

// AnOHM_FcLoad + AnOHM_FcCal (anohm.f95:717-936): the day's rows, the cell's forcing offsets applied, then the fits
AN_COLD void an_day_forcing(const SuAnohmIn &in, int xid, double dTemp_C, double dRH, int EmissionsMethod, double qf, AnDay &d,
                            AnFc &fc)
{
   d.n = 0;
   d.lenDay = 0;
   d.sumWS = 0;
   d.sumAH = 0;
   double maxSd = -1.7976931348623157e308, minSd = 1.7976931348623157e308, maxTa = maxSd, minTa = minSd, sumSd = 0, sumTa = 0;
   if (xid >= 1 && xid <= 366) {
      const int first = in.day_first[xid];
      int cnt = in.day_count[xid];
      if (cnt > AN_MAXM) cnt = AN_MAXM;   // (the host refuses forcing blocks with more rows per day of year)
      d.n = cnt;
      for (int k = 0; k < cnt; k++) {
         const SuewsHourRec &h = in.hours[first + k];
         if (h.Sd > 5) {
            const int j = d.lenDay++;
            double RH = h.RH + dRH;
            if (dRH != 0) RH = (RH > F(0.01)) ? ((RH < 100.0) ? RH : 100.0) : F(0.01);
            const double Ta = h.Ta + dTemp_C;
            d.tHr[j] = h.tHr; d.Sd[j] = h.Sd; d.Ta[j] = Ta; d.RH[j] = RH; d.pres[j] = h.pres;
            if (h.Sd > maxSd) maxSd = h.Sd;
            if (h.Sd < minSd) minSd = h.Sd;
            if (Ta > maxTa) maxTa = Ta;
            if (Ta < minTa) minTa = Ta;
            sumSd = sumSd + h.Sd;
            sumTa = sumTa + Ta;
            d.sumWS = d.sumWS + h.WS;
            d.sumAH = d.sumAH + ((EmissionsMethod == 0) ? h.qf_obs : qf);
         }
      }
   }
   const int lenDay = d.lenDay;
   fc.ASd = (maxSd - minSd) / 2;
   fc.mSd = sumSd / lenDay;
   fc.tSd = 12;
   an_shapefit(lenDay, d.tHr, d.Sd, fc.ASd, fc.mSd, fc.tSd);
   if (fc.mSd + fc.ASd < maxSd) fc.mSd = maxSd - fc.ASd;
   fc.ATa = (maxTa - minTa) / 2;
   fc.mTa = sumTa / lenDay;
   fc.tTa = 12;
   an_shapefit(lenDay, d.tHr, d.Ta, fc.ATa, fc.mTa, fc.tTa);
   if (fc.mTa < 60) fc.mTa = fc.mTa + F(273.15);
   fc.tau = (fc.tTa - fc.tSd) / 24 * 2 * AN_PI;
   fc.mWS = d.sumWS / lenDay;
   fc.mWF = 0.0 / lenDay;   // WF = 0 "for the moment" (anohm.f95:797)
   fc.mAH = d.sumAH / lenDay;
}

// The day's coefficients of the seven surfaces (AnOHM_coef's ELSE branch, anohm.f95:229-281, for is = 1..7)
AN_COLD void an_day_coefs(const SuAnohmIn &in, int xid, double dTemp_C, double dRH, int EmissionsMethod, double qf,
                          const double *moist, const double *alb, const double *emis, const double *cp, const double *kk,
                          const double *ch, double *coeff21)
{
   AnDay d;
   AnFc fc;
   an_day_forcing(in, xid, dTemp_C, dRH, EmissionsMethod, qf, d, fc);
   for (int is = 0; is < 7; is++) {
      const double xBo = an_Bo_cal(is + 1, d, fc, alb[is], emis[is], cp[is], kk[is], ch[is], moist[is]);
      double xa1, xa2, xa3, ATs, mTs, gamma;
      if (is < 6) an_coef_land(fc, alb[is], emis[is], cp[is], kk[is], ch[is], xBo, xa1, xa2, xa3, ATs, mTs, gamma);
      else an_coef_water(fc, alb[is], emis[is], cp[is], kk[is], ch[is], xBo, xa1, xa2, xa3, ATs, mTs, gamma);
      coeff21[is] = xa1;
      coeff21[7 + is] = xa2;
      coeff21[14 + is] = xa3;
   }
}

#ifndef SU_ANOHM_HOST
// AnOHM (anohm.f95:28-143) for one cell-step.  ES carries the SAVE'd locals; the class record the AnOHM columns of the
// surface tables; alb is alb_next, state_id / StoreDrainPrm(6,:) give the wetness moist_surf (driver.f95:1585).
SU_COLD void AnOHM(SuDiag &dg, const SuewsClassParams &P, const SuewsCellParams &C, SuewsEstmState &ES, const SuAnohmIn &in, int id,
                   const double *sfr, const double *state_id, const double *StoreDrainPrm, double DecidStoreCap, const double *alb,
                   double qn1, double qf, double qn1_av_prev, double dqndt_prev, int dt_since_start, double &qn1_av_next,
                   double &dqndt_next, double &a1, double &a2, double &a3, double &qs, double *deltaQi)
{
   for (int is = 0; is < NSURF; is++) deltaQi[is] = 0;
   // enough data for the day: at least six on-the-hour rows with Sd > 0 (anohm.f95:96-110)
   int xid;
   if (id >= 1 && id <= 366 && in.day_nsun[id] >= 6) {
      xid = id;
      ES.anohm_id_save = (double)id;
   } else {
      xid = (int)ES.anohm_id_save;
   }
   const int Gridiv = 1;   // a cell is its own single-grid run (see the header)
   if (!(xid == (int)ES.anohm_coef_id_save && Gridiv == (int)ES.anohm_coef_grid_save)) {
      double moist[NSURF];
      for (int is = 0; is < NSURF; is++)
         moist[is] = state_id[is] / ((is == DECID) ? DecidStoreCap : SU_SDP(StoreDrainPrm, 5, is));
      an_day_coefs(in, xid, C.dTemp_C, C.dRH, P.EmissionsMethod, qf, moist, alb, P.emis.v, P.cpAnOHM.v, P.kkAnOHM.v, P.chAnOHM.v,
                   ES.anohm_coeff_grid_day.v);
      ES.anohm_coef_id_save = (double)xid;
      ES.anohm_coef_grid_save = (double)Gridiv;
   }
   a1 = 0; a2 = 0; a3 = 0;
   for (int is = 0; is < NSURF; is++) a1 = a1 + ES.anohm_coeff_grid_day[is] * sfr[is];
   for (int is = 0; is < NSURF; is++) a2 = a2 + ES.anohm_coeff_grid_day[7 + is] * sfr[is];
   for (int is = 0; is < NSURF; is++) a3 = a3 + ES.anohm_coeff_grid_day[14 + is] * sfr[is];
   qs = -999;
   if (qn1 > -999) {
      OHM_dqndt_cal_X(P.tstep, dt_since_start, qn1_av_prev, qn1, dqndt_prev, qn1_av_next, dqndt_next);
      qs = qn1 * a1 + dqndt_prev * a2 + a3;   // dqndt_prev, as the reference has it (anohm.f95:139)
   } else {
      su_fatal(dg, 21);
   }
}
#endif
