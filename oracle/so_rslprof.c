/* ORACLE (test infrastructure only -- see suews_oracle.h).
 * Restates SUEWS-SourceCode/suews_phys_rslprof.f95 literally, including the
 * recursive cal_psim_hat (:366) and the interp_z quirk (:331,:342: the branch
 * tests an uninitialised `idx`, in practice 0 -> always interpolates). */
#include "suews_oracle.h"
#include <math.h>

enum { NZ = 30 };
#define FMAX(a, b) ((a) > (b) ? (a) : (b))
#define FMIN(a, b) ((a) < (b) ? (a) : (b))
static const double kappa = F(0.40);

static double phi_mom(int sm, double zl) { return so_stab_phi_mom(sm, zl); }
static double psi_mom(int sm, double zl) { return so_stab_psi_mom(sm, &zl); }

/* rslprof.f95:356-363: 2.*beta**3*Lc */
static double cal_elm_RSL(double beta, double Lc) { return F(2.) * (beta * (beta * beta)) * Lc; }

/* rslprof.f95:718-727: beta**2. folds to beta*beta */
static double cal_zd_RSL(double zh_RSL, double beta, double Lc) { return zh_RSL - (beta * beta) * Lc; }

/* rslprof.f95:455-505 */
static void cal_cm(int sm, double zh_RSL, double zd_RSL, double Lc, double beta, double L_MOD, double *c2, double *cm,
                   double *phi_hatmZh, double *phim_zh)
{
   const double dz = F(0.1);
   double phim_zhdz, dphi;
   *phim_zh = phi_mom(sm, (zh_RSL - zd_RSL) / L_MOD);
   phim_zhdz = phi_mom(sm, (zh_RSL - zd_RSL + dz) / L_MOD);
   dphi = (phim_zhdz - *phim_zh) / dz;
   if (*phim_zh != F(0.)) {
      *phi_hatmZh = kappa / (F(2.) * beta * (*phim_zh));
   } else {
      *phi_hatmZh = F(1.);
   }
   if (*phi_hatmZh >= F(1.)) {
      *c2 = F(0.5);
      *phi_hatmZh = F(1.);
   } else {
      *c2 = (kappa * (F(3.) - (F(2.) * (beta * beta) * Lc / (*phim_zh) * dphi))) / (F(2.) * beta * (*phim_zh) - kappa);
   }
   *c2 = F(0.5);
   *cm = (F(1.) - *phi_hatmZh) * exp(*c2 / F(2.));
}

/* rslprof.f95:508-557 */
static void cal_ch(int sm, double zh_RSL, double zd_RSL, double Lc, double beta, double L_MOD, double Scc, double f,
                   double *c2h, double *ch)
{
   double phih_zh, phih_zhdz, dphih, phi_hathZh;
   phih_zh = so_stab_phi_heat(sm, (zh_RSL - zd_RSL) / L_MOD);
   phih_zhdz = so_stab_phi_heat(sm, (zh_RSL - zd_RSL + F(1.)) / L_MOD);
   dphih = phih_zhdz - phih_zh;
   if (phih_zh != F(0.)) {
      phi_hathZh = kappa * Scc / (F(2.) * beta * phih_zh);
   } else {
      phi_hathZh = F(1.);
   }
   if (phi_hathZh >= F(1.)) {
      *c2h = F(0.5);
      phi_hathZh = F(1.);
   } else {
      *c2h = (kappa * Scc * (F(2.) + f - (dphih * F(2.) * (beta * beta) * Lc / phih_zh))) /
             (F(2.) * beta * phih_zh - kappa * Scc);
   }
   *c2h = F(0.5);
   *ch = (F(1.) - phi_hathZh) * exp(*c2h / F(2.));
}

/* rslprof.f95:426-452 */
static double cal_phim_hat(int sm, double z, double zh_RSL, double L_MOD, double beta, double lc)
{
   double elm, zd_RSL, c2, cm, phi_hatmZh, phim_zh;
   elm = cal_elm_RSL(beta, lc);
   zd_RSL = cal_zd_RSL(zh_RSL, beta, lc);
   cal_cm(sm, zh_RSL, zd_RSL, lc, beta, L_MOD, &c2, &cm, &phi_hatmZh, &phim_zh);
   return 1 - cm * exp(F(-1.) * c2 * beta * (z - zd_RSL) / elm);
}

/* rslprof.f95:366-424 (recursive, z -> 1.01 z until z > 100) */
static double cal_psim_hat(int sm, double z, double zh_RSL, double L_MOD, double beta, double Lc)
{
   const double dz = F(0.1);
   double zp, zd_RSL, phim_z, phim_zp, phim_hat_zp, phim_hat_z, psim_hat_zp, psim_hat_z;
   if (z > 100) return F(0.);
   zp = F(1.01) * z;
   zd_RSL = cal_zd_RSL(zh_RSL, beta, Lc);
   phim_z = phi_mom(sm, (z - zd_RSL) / L_MOD);
   phim_zp = phi_mom(sm, (zp - zd_RSL) / L_MOD);
   psim_hat_zp = cal_psim_hat(sm, zp, zh_RSL, L_MOD, beta, Lc);
   phim_hat_zp = cal_phim_hat(sm, zp, zh_RSL, L_MOD, beta, Lc);
   phim_hat_z = cal_phim_hat(sm, z, zh_RSL, L_MOD, beta, Lc);
   psim_hat_z = psim_hat_zp + dz / F(2.) * phim_zp * (1 - phim_hat_zp) / (zp - zd_RSL);
   psim_hat_z = psim_hat_z + dz / F(2.) * phim_z * (1 - phim_hat_z) / (z - zd_RSL);
   return psim_hat_z;
}

/* rslprof.f95:730-771 */
static double cal_z0_RSL(int sm, double zH_RSL, double zd_RSL, double beta, double L_MOD_RSL, double Lc)
{
   double z0_RSL, psimZh, psimz0, z0_RSL_x, psihatm_Zh, err;
   int it;
   psimZh = psi_mom(sm, (zH_RSL - zd_RSL) / L_MOD_RSL);
   psihatm_Zh = cal_psim_hat(sm, zH_RSL, zH_RSL, L_MOD_RSL, beta, Lc);
   z0_RSL = F(0.1) * zH_RSL;
   err = F(10.);
   it = 1;
   while ((err > F(0.001)) && (it < 10)) {
      z0_RSL_x = z0_RSL;
      psimz0 = psi_mom(sm, z0_RSL_x / L_MOD_RSL);
      z0_RSL = (zH_RSL - zd_RSL) * exp(F(-1.) * kappa / beta) * exp(F(-1.) * psimZh + psimz0) * exp(psihatm_Zh);
      err = fabs(z0_RSL_x - z0_RSL);
      it = it + 1;
   }
   return (z0_RSL > 1e-2) ? z0_RSL : 1e-2;
}

/* rslprof.f95:921-952 */
static double cal_beta_lc(int sm, double beta0, double lc_over_l)
{
   double beta_x = 0, phim, err, beta_x0;
   int it = 1;
   phim = 1;
   err = 1;
   while ((err > F(0.001)) && (it < 20)) {
      beta_x0 = beta0 / phim;
      phim = phi_mom(sm, (beta_x0 * beta_x0) * lc_over_l);
      beta_x = beta0 / phim;
      err = fabs(beta_x - beta_x0);
      it = it + 1;
   }
   return beta_x;
}

/* rslprof.f95:870-919 */
static double cal_beta_RSL(int sm, double PAI, double sfr_tr, double lc_over_L)
{
   const double kappa4 = F(0.4), a1 = F(4.), a2 = F(-0.1), a3 = F(1.5);
   double beta, betaHF, betaNL, betaN2;
   if (PAI > 0) {
      betaN2 = F(0.30) * sfr_tr / PAI + (PAI - sfr_tr) / PAI * F(0.4);
   } else {
      betaN2 = F(0.35);
   }
   betaHF = cal_beta_lc(sm, betaN2, lc_over_L);
   betaNL = cal_beta_lc(sm, kappa4 / F(2.), lc_over_L);
   if (lc_over_L > a2) {
      beta = betaHF;
   } else {
      beta = betaNL + ((betaHF - betaNL) / (F(1.) + a1 * pow(fabs(lc_over_L - a2), a3)));
   }
   if (beta > F(0.5)) beta = F(0.5);
   return beta;
}

/* rslprof.f95:774-868 */
static void RSL_cal_prms(int sm, double zh, double L_MOD, const double sfr[NSURF], double FAI, double *zH_RSL,
                         double *L_MOD_RSL, double *Lc, double *beta, double *zd_RSL, double *z0_RSL, double *elm,
                         double *Scc, double *f, double *PAI)
{
   const double r = F(0.1), Zh_min = F(0.4);
   double sfr_tr, lc_over_L;
   *zH_RSL = FMAX(zh, Zh_min);
   *PAI = sfr[BLDG] + sfr[CONIF] + sfr[DECID];
   sfr_tr = sfr[CONIF] + sfr[DECID];
   *Lc = (F(1.) - *PAI) / FAI * (*zH_RSL);
   *Lc = (*Lc > F(0.5) * (*zH_RSL)) ? *Lc : F(0.5) * (*zH_RSL);
   lc_over_L = *Lc / L_MOD;
   if (lc_over_L > 0) {
      lc_over_L = FMIN(F(2.), lc_over_L);
   } else {
      lc_over_L = FMAX(F(-2.), lc_over_L);
   }
   *L_MOD_RSL = *Lc / lc_over_L;
   *beta = cal_beta_RSL(sm, *PAI, sfr_tr, lc_over_L);
   *Scc = F(0.5) + F(0.3) * tanh(F(2.) * lc_over_L);
   *f = F(0.5) * (pow(F(1.) + F(4.) * r * (*Scc), F(0.5))) - F(0.5);
   *zd_RSL = cal_zd_RSL(*zH_RSL, *beta, *Lc);
   *elm = cal_elm_RSL(*beta, *Lc);
   *z0_RSL = cal_z0_RSL(sm, *zH_RSL, *zd_RSL, *beta, *L_MOD_RSL, *Lc);
}

/* rslprof.f95:315-353 */
static double interp_z(double z_x, const double z[NZ], const double v[NZ])
{
   double dif[NZ], dz, slope, best;
   int i, idx_low = 0, idx_high = 0; /* 1-based, 0 = none (MAXLOC/MINLOC with empty mask) */
   for (i = 0; i < NZ; i++) dif[i] = z[i] - z_x;
   best = 0;
   for (i = 0; i < NZ; i++)
      if (dif[i] < F(0.) && (idx_low == 0 || dif[i] > best)) { idx_low = i + 1; best = dif[i]; }
   for (i = 0; i < NZ; i++)
      if (dif[i] > F(0.) && (idx_high == 0 || dif[i] < best)) { idx_high = i + 1; best = dif[i]; }
   /* `if (idx > 0)` on the uninitialised idx: interpolation branch always taken */
   dz = z[idx_high - 1] - z[idx_low - 1];
   slope = (v[idx_high - 1] - v[idx_low - 1]) / dz;
   return v[idx_low - 1] + (z_x - z[idx_low - 1]) * slope;
}

/* rslprof.f95:13-313 */
void so_RSLProfile(SoCtx *cx, double Zh, double z0m, double zdm, double L_MOD, const double sfr[NSURF], double FAI,
                   int StabilityMethod, double avcp, double lv_J_kg, double avdens, double avU1, double Temp_C,
                   double avRH, double Press_hPa, double zMeas, double qh, double qe, double *T2_C, double *q2_gkg,
                   double *U10_ms, double *RH2, double dataoutLineRSL[NCOL_RSL])
{
   const int sm = StabilityMethod;
   double psihatm_z[NZ], psihath_z[NZ], dif[NZ], zarray[NZ], U[NZ], T[NZ], q[NZ];
   double z0_RSL, zd_RSL, Lc, Scc, psimz, psimz0, psimza, phimzp, phimz, phihzp, phihz, psihz, psihza, beta, elm, f,
      cm, c2, ch, c2h, t_h, q_h, TStar_RSL, UStar_RSL, PAI, L_MOD_RSL, zH_RSL, dz, phi_hatmZh, phim_zh, qa_gkg,
      qStar_RSL, mn;
   int i, z, idx_can, idx_za, nz_can, flag_RSL, k;
   (void)cx;

   RSL_cal_prms(sm, Zh, L_MOD, sfr, FAI, &zH_RSL, &L_MOD_RSL, &Lc, &beta, &zd_RSL, &z0_RSL, &elm, &Scc, &f, &PAI);

   if (zH_RSL <= 2) {
      nz_can = 3;
   } else if (zH_RSL <= 10) {
      nz_can = 10;
   } else {
      nz_can = 15;
   }
   dz = zH_RSL / nz_can;
   for (i = 1; i <= nz_can; i++) zarray[i - 1] = dz * i;
   if (dz > 2) zarray[0] = F(1.999);
   zarray[nz_can - 1] = zH_RSL;

   dz = (zMeas - zH_RSL) / (NZ - nz_can);
   for (i = nz_can + 1; i <= NZ; i++) zarray[i - 1] = zH_RSL + (i - nz_can) * dz;

   for (z = 0; z < NZ; z++) dif[z] = fabs(zarray[z] - zH_RSL);
   idx_can = 1;
   mn = dif[0];
   for (z = 1; z < NZ; z++)
      if (dif[z] < mn) { mn = dif[z]; idx_can = z + 1; }
   zarray[idx_can - 1] = zH_RSL;

   idx_za = NZ;
   zarray[idx_za - 1] = zMeas;

   flag_RSL = ((F(1.) - PAI) / FAI <= 18) && ((F(1.) - PAI) / FAI > F(.021)) && (zH_RSL > 6);

   if (flag_RSL) {
      cal_ch(sm, zH_RSL, zd_RSL, Lc, beta, L_MOD_RSL, Scc, f, &c2h, &ch);
      cal_cm(sm, zH_RSL, zd_RSL, Lc, beta, L_MOD_RSL, &c2, &cm, &phi_hatmZh, &phim_zh);
      for (z = 0; z < NZ; z++) { psihatm_z[z] = F(0.) * zarray[z]; psihath_z[z] = F(0.) * zarray[z]; }
      for (z = NZ - 1; z >= idx_can; z--) { /* 1-based z = nz-1 .. idx_can */
         double zz = zarray[z - 1], zzp = zarray[z];
         phimz = phi_mom(sm, (zz - zd_RSL) / L_MOD_RSL);
         phimzp = phi_mom(sm, (zzp - zd_RSL) / L_MOD_RSL);
         phihz = so_stab_phi_heat(sm, (zz - zd_RSL) / L_MOD_RSL);
         phihzp = so_stab_phi_heat(sm, (zzp - zd_RSL) / L_MOD_RSL);
         psihatm_z[z - 1] =
            psihatm_z[z] + dz / F(2.) * phimzp * (cm * exp(F(-1.) * c2 * beta * (zzp - zd_RSL) / elm)) / (zzp - zd_RSL);
         psihatm_z[z - 1] =
            psihatm_z[z - 1] + dz / F(2.) * phimz * (cm * exp(F(-1.) * c2 * beta * (zz - zd_RSL) / elm)) / (zz - zd_RSL);
         psihath_z[z - 1] =
            psihath_z[z] + dz / F(2.) * phihzp * (ch * exp(F(-1.) * c2h * beta * (zzp - zd_RSL) / elm)) / (zzp - zd_RSL);
         psihath_z[z - 1] =
            psihath_z[z - 1] + dz / F(2.) * phihz * (ch * exp(F(-1.) * c2h * beta * (zz - zd_RSL) / elm)) / (zz - zd_RSL);
      }
   } else {
      for (z = 0; z < NZ; z++) { psihatm_z[z] = 0; psihath_z[z] = 0; }
      L_MOD_RSL = L_MOD;
      z0_RSL = z0m;
      zd_RSL = zdm;
      if (L_MOD_RSL > 0) {
         Lc = FMIN(F(2.2) * kappa / beta * L_MOD_RSL, Lc);
      } else {
         Lc = FMAX(-(2 / (beta * beta) * L_MOD_RSL), Lc);
      }
      elm = cal_elm_RSL(beta, Lc);
   }

   psimz0 = psi_mom(sm, z0_RSL / L_MOD_RSL);
   psimza = psi_mom(sm, (zMeas - zd_RSL) / L_MOD_RSL);
   psihza = so_stab_psi_heat(sm, (zMeas - zd_RSL) / L_MOD_RSL);
   UStar_RSL = avU1 * kappa / (log((zMeas - zd_RSL) / z0_RSL) - psimza + psimz0 + psihatm_z[NZ - 1]);
   UStar_RSL = (UStar_RSL > 0.15) ? UStar_RSL : 0.15;
   TStar_RSL = F(-1.) * (qh / (avcp * avdens)) / UStar_RSL;
   qStar_RSL = F(-1.) * (qe / lv_J_kg * avdens) / UStar_RSL;
   qa_gkg = so_RH2qa(avRH / 100, Press_hPa, Temp_C);

   for (z = idx_can; z <= NZ; z++) {
      double zz = zarray[z - 1];
      psimz = psi_mom(sm, (zz - zd_RSL) / L_MOD_RSL);
      psihz = so_stab_psi_heat(sm, (zz - zd_RSL) / L_MOD_RSL);
      U[z - 1] = (log((zz - zd_RSL) / z0_RSL) - psimz + psimz0 + psihatm_z[z - 1]) / kappa;
      T[z - 1] =
         (log((zz - zd_RSL) / (zMeas - zd_RSL)) - psihz + psihza + psihath_z[z - 1] - psihath_z[idx_za - 1]) / kappa;
      q[z - 1] =
         (log((zz - zd_RSL) / (zMeas - zd_RSL)) - psihz + psihza + psihath_z[z - 1] - psihath_z[idx_za - 1]) / kappa;
   }
   if (idx_can > 1) {
      t_h = Scc * TStar_RSL / (beta * f);
      q_h = Scc * qStar_RSL / (beta * f);
      for (z = 1; z <= idx_can - 1; z++) {
         double zz = zarray[z - 1];
         U[z - 1] = U[idx_can - 1] * exp(beta * (zz - zH_RSL) / elm);
         T[z - 1] = T[idx_can - 1] + (t_h * exp(beta * f * (zz - zH_RSL) / elm) - t_h) / TStar_RSL;
         q[z - 1] = q[idx_can - 1] + (q_h * exp(beta * f * (zz - zH_RSL) / elm) - q_h) / qStar_RSL;
      }
   }

   for (z = 0; z < NZ; z++) {
      U[z] = U[z] * UStar_RSL;
      T[z] = T[z] * TStar_RSL + Temp_C;
      q[z] = (q[z] * qStar_RSL + qa_gkg / F(1000.)) * F(1000.);
   }

   k = 0;
   for (z = 0; z < NZ; z++) dataoutLineRSL[k++] = zarray[z];
   for (z = 0; z < NZ; z++) dataoutLineRSL[k++] = U[z];
   for (z = 0; z < NZ; z++) dataoutLineRSL[k++] = T[z];
   for (z = 0; z < NZ; z++) dataoutLineRSL[k++] = q[z];
   dataoutLineRSL[k++] = L_MOD_RSL;
   dataoutLineRSL[k++] = zH_RSL;
   dataoutLineRSL[k++] = Lc;
   dataoutLineRSL[k++] = beta;
   dataoutLineRSL[k++] = zd_RSL;
   dataoutLineRSL[k++] = z0_RSL;
   dataoutLineRSL[k++] = elm;
   dataoutLineRSL[k++] = Scc;
   dataoutLineRSL[k++] = f;
   dataoutLineRSL[k++] = UStar_RSL;
   dataoutLineRSL[k++] = FAI;
   dataoutLineRSL[k++] = PAI;
   dataoutLineRSL[k++] = flag_RSL ? 1.0 : 0.0;

   *T2_C = interp_z(2.0, zarray, T);
   *q2_gkg = interp_z(2.0, zarray, q);
   *U10_ms = interp_z(10.0, zarray, U);
   *RH2 = so_qa2RH(*q2_gkg, Press_hPa, *T2_C);
}
