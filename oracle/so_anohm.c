/* ORACLE (test infrastructure only -- see suews_oracle.h).
 *
 * AnOHM, the Analytical Objective Hysteresis Model (StorageHeatMethod 3): SUEWS-SourceCode/suews_phys_anohm.f95,
 * called from SUEWS_cal_Qs (suews_ctrl_driver.f95:1577-1594).  Once a day the OHM coefficients a1, a2, a3 of every
 * surface are derived from the day's on-the-hour forcing rows: sinusoids are fitted to the daytime Sd and Ta
 * (AnOHM_ShapeFit -> MINPACK lmdif1), a Bowen ratio is solved for (AnOHM_Bo_cal -> hybrd1) and the analytical
 * solution of the surface heat conduction problem gives the coefficients (AnOHM_coef_land_cal / _water_cal); the rest of
 * the day reloads them from SAVE'd locals.  QS is then OHM's QS = a1 Qn + a2 dQn/dt + a3.
 *
 * PARITY UNPINNED: the reference repository holds no output of any AnOHM run (no test, no golden file selects
 * StorageHeatMethod 3), and the reference cannot be compiled here.  This file is a restatement checked by reading and
 * by the properties tests/test_anohm_oracle.py lists; the CUDA path is tested against it.
 *
 * Kept as the reference has them: REAL(4) literals widened at use (PI = ATAN(1.0)*4 is the REAL(4) pi), the Bowen-ratio
 * residual's MIN(Ts - 273.15, -40.) (the modelled surface temperature never enters above -40 degC), QS formed with
 * dqndt_prev where OHM uses dqndt_next (anohm.f95:139 against ohm.f95:140), hybrd1's tolerance 1E-20 (below the
 * machine precision, so the solve ends on one of MINPACK's "no progress" exits), the 0.3 / 0.2 placeholders of the water
 * surface, and the cache that is marked valid only after the water surface (the 7th) has been done.
 */
#include "suews_oracle.h"
#include <math.h>
#include <stddef.h>
/* epsilon(), tiny(), huge() of REAL(8) */
#define SO_EPSMCH 2.220446049250313e-16
#define SO_DWARF 2.2250738585072014e-308
#define SO_GIANT 1.7976931348623157e308

const double *so_met_grid = 0;
int so_met_grid_nrows = 0;

#define AN_PI F(3.14159274)                         /* REAL(KIND(1d0)), PARAMETER :: PI = ATAN(1.0)*4 */
#define AN_OMEGA ((2 * AN_PI) / (24 * 60 * 60))     /* 2*Pi/(24*60*60) */
#define AN_SIGMA F(5.67e-8)
#define AN_C2K F(273.15)

/* anohm.f95:363-515 */
void so_AnOHM_coef_land_cal(double ASd, double mSd, double ATa, double mTa, double tau, double mWS, double mWF,
                            double mAH, double xalb, double xemis, double xcp, double xk, double xch, double xBo,
                            double *xa1, double *xa2, double *xa3, double *ATs, double *mTs, double *gamma)
{
   const double SIGMA = AN_SIGMA, OMEGA = AN_OMEGA;
   double beta, f, fL, fT, lambda, delta, m, n, ms, ns, ceta, cphi, eta, phi, xlag, xx1, xx2, xx3, xchWS;
   double st, ct, sg, cg;
   xchWS = xch * mWS;
   if (fabs(xBo) < F(0.001)) {
      beta = (double)(1.f / 0.001f);   /* 1/0.001: a REAL(4) quotient */
   } else {
      beta = 1 / xBo;
   }
   f = ((1 + beta) * xchWS + 4 * SIGMA * xemis * (mTa * mTa * mTa));
   fL = 4 * SIGMA * xemis * (mTa * mTa * mTa);
   fT = (1 + beta) * xchWS;
   lambda = xk / xcp;
   delta = sqrt(F(.5) * (mWF * mWF + sqrt((mWF * mWF) * (mWF * mWF) + 16 * (lambda * lambda) * (OMEGA * OMEGA))));
   m = (2 * lambda) / (delta + mWF);
   n = delta / OMEGA;

   *mTs = (mSd * (1 - xalb) / f) + mTa;
   ms = 1 + xk / (f * m);
   ns = xk / (f * n);
   st = sin(tau); ct = cos(tau);
   xx1 = f * st * ATa;
   xx2 = (1 - xalb) * ASd + f * ct * ATa;
   *gamma = atan(ns / ms) + atan(xx1 / xx2);
   sg = sin(*gamma); cg = cos(*gamma);
   *ATs = -(st * ATa) / (ns * cg - ms * sg);

   xx1 = (ns * cg + sg - ms * sg) * st * ATa * fL;
   xx2 = (xalb - 1) * (ns * cg - ms * sg) * ASd;
   xx3 = (-ms * ct * st + cg * (ns * ct + st)) * ATa * fL;
   xx2 = xx2 - xx3;
   phi = atan(xx1 / xx2);
   xx3 = (ns * cg - ms * sg);
   xx1 = (1 + sg / xx3) * st * ATa * fL;
   xx2 = (xalb - 1) * ASd - (ct + cg * st / xx3) * ATa * fL;
   cphi = sqrt(xx1 * xx1 + xx2 * xx2);

   xx1 = m * cg - n * sg;
   xx2 = m * sg + n * cg;
   eta = atan(xx1 / xx2);
   xx1 = (xk * xk) * (m * m + n * n) * (*ATs * *ATs);
   xx2 = (m * m) * (n * n);
   ceta = sqrt(xx1 / xx2);

   xlag = eta - phi;
   *xa1 = (ceta * cos(xlag)) / cphi;
   *xa2 = (ceta * sin(xlag)) / (OMEGA * cphi);
   *xa2 = *xa2 / 3600;
   *xa3 = -*xa1 * (fT / f) * (mSd * (1 - xalb)) - mAH;
}

/* anohm.f95:520-663 ("this SUBROUTINE hasn't been well tested", says the reference) */
void so_AnOHM_coef_water_cal(double ASd, double mSd, double ATa, double mTa, double tau, double mWS, double mWF,
                             double mAH, double xalb, double xemis, double xcp, double xk, double xch, double xBo,
                             double xeta, double xmu, double *xa1, double *xa2, double *xa3, double *ATs, double *mTs,
                             double *gamma)
{
   const double SIGMA = AN_SIGMA, PI = AN_PI, OMEGA = AN_OMEGA;
   double beta, f, fL, fT, lambda, calb, delta, xm, xn, phi, czeta, ctheta, zeta, theta, xlag, xx1, xx2, xx3, kappa;
   double dtau, dpsi, dphi, cdtau, cdpsi, cdphi, xxT, xxkappa, xxdltphi, xchWS;
   (void)mAH; (void)mWF;   /* dummy = mah + mwf */
   xm = xk * (xmu * xmu);
   xn = xcp * OMEGA;
   phi = atan(xn / xm);
   kappa = sqrt(xcp * OMEGA / (2 * xk));

   xchWS = xch * mWS;
   beta = 1 / xBo;
   f = ((1 + beta) * xchWS + 4 * SIGMA * xemis * (mTa * mTa * mTa));
   fL = 4 * SIGMA * xemis * (mTa * mTa * mTa);
   fT = (1 + beta) * xchWS;
   calb = 1 - xalb;
   lambda = sqrt(xm * xm + xn * xn);

   dtau = atan(xk * kappa / (f + xk * kappa));
   dpsi = atan((xk - xmu) / kappa);
   dphi = atan(kappa * (f + xk * xmu) / (f * (kappa - xmu) + xk * kappa * (2 * kappa - xmu)));
   cdtau = sqrt((xk * kappa) * (xk * kappa) + (f + xk * kappa) * (f + xk * kappa));
   cdpsi = sqrt((xk - xmu) * (xk - xmu) + kappa * kappa);
   cdphi = cdtau * cdpsi;

   *mTs = (mSd * (1 - xalb + xeta) / f) + mTa;
   xx1 = (xk * xeta * xmu * calb * ASd * cdpsi);
   xx1 = xx1 * xx1;
   xx2 = 2 * lambda * sqrt(xx1) * (calb * ASd * sin(phi - dpsi) + f * ATa * sin(tau + phi - dpsi));
   {
      double u = (calb * ASd + cos(tau) * f * ATa), v = (sin(tau) * f * ATa);
      xx3 = (lambda * lambda) * (u * u + v * v);
   }
   *ATs = 1 / (cdtau * lambda) * sqrt(xx1 + xx2 + xx3);
   xx1 = (xk * kappa * calb * ASd + cdtau * f * ATa * sin(tau + dtau)) * lambda +
         (xk * xeta * xmu * calb * ASd * cdphi * sin(phi + dphi));
   xx2 = ((f + xk * kappa) * calb * ASd - cdtau * f * ATa * cos(tau + dtau)) * lambda -
         xk * xeta * xmu * calb * ASd * cdphi * cos(phi + dphi);
   delta = atan(xx1 / xx2);
   *gamma = delta;

   xx1 = fL * (*ATs * sin(delta) - ATa * sin(tau));
   xx2 = calb * ASd - fL * (*ATs * cos(delta) - ATa * cos(tau));
   theta = atan(xx1 / xx2);
   ctheta = sqrt(xx1 * xx1 + xx2 * xx2);

   xxT = (double)sqrtf(2.f) * kappa * lambda * *ATs;   /* SQRT(2.) is REAL(4) */
   xxkappa = cdpsi * xeta * xmu * ASd;
   xxdltphi = cos(delta) * sin(dpsi) * cos(phi) - sin(delta) * cos(dpsi) * sin(phi);
   xx1 = xxT * sin(PI / 4 - delta) + xxkappa * sin(phi + dpsi);
   xx2 = xxT * sin(PI / 4 + delta) - xxkappa * sin(phi - dpsi);
   zeta = atan(xx1 / xx2);
   xx1 = 2 * (double)sqrtf(2.f) * xxkappa * xxT * xxdltphi;
   xx2 = (1 - cos(2 * dpsi) * cos(2 * phi)) * (xxkappa * xxkappa);
   xx3 = xxT * xxT;
   czeta = xk / lambda * sqrt(xx1 + xx2 + xx3);

   xlag = zeta - theta;
   *xa1 = (czeta * cos(xlag)) / ctheta;
   *xa2 = (czeta * sin(xlag)) / (OMEGA * ctheta);
   *xa2 = *xa2 / 3600;
   *xa3 = mSd * (xalb - 1) * (xeta + (fT - fL * xeta) / f * *xa1);
}

/* anohm.f95:290-360 */
static double AnOHM_xTs(int sfc_typ, double ASd, double mSd, double ATa, double mTa, double tau, double mWS, double mWF,
                        double mAH, double xalb, double xemis, double xcp, double xk, double xch, double xBo,
                        double tSd, double xTHr)
{
   const double OMEGA = AN_OMEGA;
   double xa1, xa2, xa3, ATs, mTs, gamma;
   if (sfc_typ >= 1 && sfc_typ <= 6) {
      so_AnOHM_coef_land_cal(ASd, mSd, ATa, mTa, tau, mWS, mWF, mAH, xalb, xemis, xcp, xk, xch, xBo, &xa1, &xa2, &xa3,
                             &ATs, &mTs, &gamma);
   } else {
      so_AnOHM_coef_water_cal(ASd, mSd, ATa, mTa, tau, mWS, mWF, mAH, xalb, xemis, xcp, xk, xch, xBo, F(0.3), F(0.2),
                              &xa1, &xa2, &xa3, &ATs, &mTs, &gamma);
   }
   return ATs * sin(OMEGA * (xTHr - tSd + 6) * 3600 - gamma) + mTs;
}

/* anohm.f95:1400-1449 */
static double esat_fn(double Ta) { return F(6.106) * exp(F(17.27) * Ta / (F(237.3) + Ta)); }
static double qsat_fn(double Ta, double pres)
{
   double es = esat_fn(Ta);
   return (F(0.0180153) / F(0.028965)) * es / pres;
}
static double qa_fn(double Ta, double RH, double pres)
{
   double es = esat_fn(Ta);
   double ea = es * RH / 100;
   return (F(0.0180153) / F(0.028965)) * ea / pres;
}

/* anohm.f95:1005-1032 */
typedef struct { const double *xdat, *ydat; } FSinCtx;
static void fSin(int m, int n, const double *x, double *fvec, void *vc)
{
   const FSinCtx *c = (const FSinCtx *)vc;
   const double PI = AN_PI;
   int i;
   (void)n;
   for (i = 0; i < m; i++) fvec[i] = x[0] + x[1] * sin(2 * PI / 24 * (c->xdat[i] - x[2])) - c->ydat[i];
}

/* anohm.f95:940-998 */
void so_AnOHM_ShapeFit(int m, const double *tHr, const double *obs, double *amp, double *mean, double *tpeak)
{
   FSinCtx c;
   double x[3], fvec[SO_MP_MAXM];
   c.xdat = tHr; c.ydat = obs;
   x[0] = *mean; x[1] = *amp; x[2] = *tpeak;
   if (m <= SO_MP_MAXM) so_lmdif1(fSin, &c, m, 3, x, fvec, 0.00001);
   *mean = x[0];
   *amp = x[1];
   *tpeak = x[2] + 6;
   if (*amp < 0) {
      *amp = fabs(*amp);
      *tpeak = x[2] - 12 + 6 + 24;
      *tpeak = fmod(*tpeak, F(24.));
   }
   *tpeak = fmod(*tpeak, F(24.));
}

/* the day's on-the-hour rows (AnOHM_FcLoad, anohm.f95:717-805) */
typedef struct {
   int n;
   double tHr[SO_MP_MAXM], Sd[SO_MP_MAXM], Ta[SO_MP_MAXM], RH[SO_MP_MAXM], pres[SO_MP_MAXM], WS[SO_MP_MAXM],
      WF[SO_MP_MAXM], AH[SO_MP_MAXM];
} AnDay;

static void AnOHM_FcLoad(int xid, const double *met, int nrows, int EmissionsMethod, double qf, AnDay *d)
{
   int ir;
#define MET(col) met[(size_t)ir + (size_t)nrows * ((col) - 1)]
   d->n = 0;
   for (ir = 0; ir < nrows; ir++) {
      if (MET(2) == xid && MET(4) == 0) {
         int k = d->n;
         if (k >= SO_MP_MAXM) break;
         d->tHr[k] = MET(3);
         d->Sd[k] = MET(15);
         d->Ta[k] = MET(12);
         d->RH[k] = MET(11);
         d->pres[k] = MET(13);
         d->WS[k] = MET(10);
         d->WF[k] = 0;
         d->AH[k] = (EmissionsMethod == 0) ? MET(9) : qf;
         d->n = k + 1;
      }
   }
#undef MET
}

static double maxval(int n, const double *x)
{
   double v = -SO_GIANT;
   int i;
   for (i = 0; i < n; i++) if (x[i] > v) v = x[i];
   return v;
}
static double minval(int n, const double *x)
{
   double v = SO_GIANT;
   int i;
   for (i = 0; i < n; i++) if (x[i] < v) v = x[i];
   return v;
}
static double sumval(int n, const double *x)
{
   double s = 0;
   int i;
   for (i = 0; i < n; i++) s = s + x[i];
   return s;
}

typedef struct { double ASd, mSd, tSd, ATa, mTa, tTa, tau, mWS, mWF, mAH; } AnFc;

/* anohm.f95:808-936 */
static void AnOHM_FcCal(const AnDay *d, AnFc *o)
{
   const double PI = AN_PI;
   double tHrDay[SO_MP_MAXM], selX[SO_MP_MAXM];
   int i, lenDay = 0;
   for (i = 0; i < d->n; i++) if (d->Sd[i] > 5) tHrDay[lenDay++] = d->tHr[i];
#define PACK(src) do { int k_ = 0; for (i = 0; i < d->n; i++) if (d->Sd[i] > 5) selX[k_++] = d->src[i]; } while (0)
   PACK(Sd);
   o->ASd = (maxval(lenDay, selX) - minval(lenDay, selX)) / 2;
   o->mSd = sumval(lenDay, selX) / lenDay;
   o->tSd = 12;
   so_AnOHM_ShapeFit(lenDay, tHrDay, selX, &o->ASd, &o->mSd, &o->tSd);
   if (o->mSd + o->ASd < maxval(lenDay, selX)) o->mSd = maxval(lenDay, selX) - o->ASd;

   PACK(Ta);
   o->ATa = (maxval(lenDay, selX) - minval(lenDay, selX)) / 2;
   o->mTa = sumval(lenDay, selX) / lenDay;
   o->tTa = 12;
   so_AnOHM_ShapeFit(lenDay, tHrDay, selX, &o->ATa, &o->mTa, &o->tTa);
   if (o->mTa < 60) o->mTa = o->mTa + AN_C2K;

   o->tau = (o->tTa - o->tSd) / 24 * 2 * PI;

   PACK(WS);
   o->mWS = sumval(lenDay, selX) / lenDay;
   PACK(WF);
   o->mWF = sumval(lenDay, selX) / lenDay;
   PACK(AH);
   o->mAH = sumval(lenDay, selX) / lenDay;
#undef PACK
}

/* anohm.f95:1170-1395: the residual x - Bo(x) whose root hybrd1 looks for */
typedef struct {
   int sfc_typ, lenDay;
   double ASd, mSd, ATa, mTa, tau, mWS, mWF, mAH, xalb, xemis, xcp, xk, xch, xSM, tSd;
   double Sd[SO_MP_MAXM], Ta[SO_MP_MAXM], RH[SO_MP_MAXM], pres[SO_MP_MAXM], tHr[SO_MP_MAXM];
} BoCtx;

static void fcnBo(int n, const double *x, double *fvec, void *vc)
{
   const BoCtx *c = (const BoCtx *)vc;
   const double C2K = AN_C2K, cp_air = F(1006.38), Lv_air = F(2264.705E3);
   double xBo = x[0], xSM, sQH = 0, sQE = 0;
   int i;
   (void)n;
   xSM = (c->xSM < F(1.0)) ? c->xSM : F(1.0);   /* MIN(prms(14), 1.0) */
   if (xSM == 0) {
      xBo = 1000;
   } else {
      for (i = 0; i < c->lenDay; i++) {
         double Ts = AnOHM_xTs(c->sfc_typ, c->ASd, c->mSd, c->ATa, c->mTa, c->tau, c->mWS, c->mWF, c->mAH, c->xalb,
                               c->xemis, c->xcp, c->xk, c->xch, xBo, c->tSd, c->tHr[i]);
         double qs, qa;
         Ts = Ts - C2K;
         Ts = (Ts < F(-40.)) ? Ts : F(-40.);   /* MIN(Ts(i) - C2K, -40.) */
         qs = qsat_fn(Ts, c->pres[i]);
         qa = qa_fn(c->Ta[i], c->RH[i], c->pres[i]);
         sQH = sQH + cp_air * (Ts - c->Ta[i]);
         sQE = sQE + xSM * Lv_air * (qs - qa);
      }
      xBo = sQH / sQE;
   }
   fvec[0] = x[0] - xBo;
}

/* anohm.f95:1037-1166 */
static double AnOHM_Bo_cal(int sfc_typ, const AnDay *d, const AnFc *fc, double xalb, double xemis, double xcp, double xk,
                           double xch, double xSM)
{
   BoCtx c;
   double x[1], fvec[1];
   int i;
   c.sfc_typ = sfc_typ;
   c.ASd = fc->ASd; c.mSd = fc->mSd; c.ATa = fc->ATa; c.mTa = fc->mTa; c.tau = fc->tau; c.mWS = fc->mWS; c.mWF = fc->mWF;
   c.mAH = fc->mAH; c.xalb = xalb; c.xemis = xemis; c.xcp = xcp; c.xk = xk; c.xch = xch; c.xSM = xSM; c.tSd = fc->tSd;
   c.lenDay = 0;
   for (i = 0; i < d->n; i++) {
      if (d->Sd[i] > 5) {
         int k = c.lenDay++;
         c.Sd[k] = d->Sd[i]; c.Ta[k] = d->Ta[i]; c.RH[k] = d->RH[i]; c.pres[k] = d->pres[i]; c.tHr[k] = d->tHr[i];
      }
   }
   x[0] = 10;
   so_hybrd1(fcnBo, &c, 1, x, fvec, (double)1E-20f);   /* REAL(kind=8) :: tol = 1E-20 */
   return x[0];
}

/* anohm.f95:149-283 */
static void AnOHM_coef(SoAnohmState *an, int sfc_typ /* 1..7 */, int xid, int xgrid, const double *met, int nrows,
                       const double moist[NSURF], int EmissionsMethod, double qf, const double alb[NSURF],
                       const double emis[NSURF], const double cpAnOHM[NSURF], const double kkAnOHM[NSURF],
                       const double chAnOHM[NSURF], double *xa1, double *xa2, double *xa3)
{
   if (xid == (int)an->coef_id_save && xgrid == (int)an->coef_grid_save) {
      *xa1 = an->coeff_grid_day[(sfc_typ - 1) + 7 * 0];
      *xa2 = an->coeff_grid_day[(sfc_typ - 1) + 7 * 1];
      *xa3 = an->coeff_grid_day[(sfc_typ - 1) + 7 * 2];
   } else {
      AnDay d;
      AnFc fc;
      double xalb, xemis, xcp, xk, xch, xmoist, xBo, ATs, mTs, gamma;
      AnOHM_FcLoad(xid, met, nrows, EmissionsMethod, qf, &d);   /* AnOHM_Fc = FcLoad + FcCal, then FcLoad again */
      AnOHM_FcCal(&d, &fc);
      xalb = alb[sfc_typ - 1];
      xemis = emis[sfc_typ - 1];
      xcp = cpAnOHM[sfc_typ - 1];
      xk = kkAnOHM[sfc_typ - 1];
      xch = chAnOHM[sfc_typ - 1];
      xmoist = moist[sfc_typ - 1];
      xBo = AnOHM_Bo_cal(sfc_typ, &d, &fc, xalb, xemis, xcp, xk, xch, xmoist);
      if (sfc_typ >= 1 && sfc_typ <= 6) {
         so_AnOHM_coef_land_cal(fc.ASd, fc.mSd, fc.ATa, fc.mTa, fc.tau, fc.mWS, fc.mWF, fc.mAH, xalb, xemis, xcp, xk, xch,
                                xBo, xa1, xa2, xa3, &ATs, &mTs, &gamma);
      } else {
         so_AnOHM_coef_water_cal(fc.ASd, fc.mSd, fc.ATa, fc.mTa, fc.tau, fc.mWS, fc.mWF, fc.mAH, xalb, xemis, xcp, xk,
                                 xch, xBo, F(0.3), F(0.2), xa1, xa2, xa3, &ATs, &mTs, &gamma);
         an->coef_id_save = xid;
         an->coef_grid_save = xgrid;
      }
      an->coeff_grid_day[(sfc_typ - 1) + 7 * 0] = *xa1;
      an->coeff_grid_day[(sfc_typ - 1) + 7 * 1] = *xa2;
      an->coeff_grid_day[(sfc_typ - 1) + 7 * 2] = *xa3;
   }
}

/* anohm.f95:28-143 */
void so_AnOHM(SoCtx *cx, SoAnohmState *an, int tstep, int dt_since_start, double qn1, double qn1_av_prev,
              double dqndt_prev, double qf, const double *met, int nrows, const double moist_surf[NSURF],
              const double alb[NSURF], const double emis[NSURF], const double cpAnOHM[NSURF],
              const double kkAnOHM[NSURF], const double chAnOHM[NSURF], const double sfr[NSURF], int EmissionsMethod,
              int id, int Gridiv, double *qn1_av_next, double *dqndt_next, double *a1, double *a2, double *a3,
              double *qs, double deltaQi[NSURF])
{
   double xa1[NSURF], xa2[NSURF], xa3[NSURF];
   int is, ir, xid, cnt = 0;
   for (is = 0; is < NSURF; is++) { xa1[is] = F(0.1); xa2[is] = F(0.2); xa3[is] = 10; deltaQi[is] = 0; }
   *qs = -999;
   /* enough data for the day: at least six on-the-hour rows of day `id` with Sd > 0 */
   for (ir = 0; ir < nrows; ir++)
      if (met[(size_t)ir + (size_t)nrows * 1] == id && met[(size_t)ir + (size_t)nrows * 3] == 0 &&
          met[(size_t)ir + (size_t)nrows * 14] > 0)
         cnt++;
   if (cnt >= 6) {
      xid = id;
      an->id_save = id;
   } else {
      xid = (int)an->id_save;
   }
   for (is = 0; is < NSURF; is++)
      AnOHM_coef(an, is + 1, xid, Gridiv, met, nrows, moist_surf, EmissionsMethod, qf, alb, emis, cpAnOHM, kkAnOHM,
                 chAnOHM, &xa1[is], &xa2[is], &xa3[is]);
   *a1 = 0; *a2 = 0; *a3 = 0;
   for (is = 0; is < NSURF; is++) *a1 = *a1 + xa1[is] * sfr[is];
   for (is = 0; is < NSURF; is++) *a2 = *a2 + xa2[is] * sfr[is];
   for (is = 0; is < NSURF; is++) *a3 = *a3 + xa3[is] * sfr[is];
   *qs = -999;
   if (qn1 > -999) {
      so_OHM_dqndt_cal_X(tstep, dt_since_start, qn1_av_prev, qn1, dqndt_prev, qn1_av_next, dqndt_next);
      so_OHM_QS_cal(qn1, dqndt_prev, *a1, *a2, *a3, qs);   /* dqndt_prev, as the reference has it (anohm.f95:139) */
   } else {
      so_error_hint(cx, 21, "SUEWS_AnOHM.f95: bad value for qn found during qs calculation.", qn1, -55.5, -55);
   }
}
