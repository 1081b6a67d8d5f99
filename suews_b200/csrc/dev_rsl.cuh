// dev_rsl.cuh -- roughness-sublayer profile diagnostics (T2, q2, U10, RH2 and the 30-level profile).
// Device restatement of suews_phys_rslprof.f95.  Differences in organisation, not in arithmetic:
//  * cal_psim_hat (rslprof.f95:366) is a recursion z -> 1.01 z up to 100 m; here it is one upward loop
//    with the level-invariant cal_cm / zd / elm evaluated once instead of twice per level and phim(zp)
//    reused as the next level's phim(z) (the additions run bottom-up: a few ulp, see cal_psim_hat_zh);
//  * the whole z0_RSL search is skipped when the RSL correction is off (flag_RSL false): the
//    reference computes it and then overwrites z0_RSL with z0m (rslprof.f95:220-223).
#pragma once
#include "dev_atmos.cuh"

#define SU_NZ 30
#define SU_KAPPA F(0.40)

SU_INL double cal_elm_RSL(double beta, double Lc) { return F(2.) * (beta * (beta * beta)) * Lc; }   // rslprof.f95:356
SU_INL double cal_zd_RSL(double zh, double beta, double Lc) { return zh - (beta * beta) * Lc; }      // rslprof.f95:718

// rslprof.f95:455-505 (c2 is forced to 0.5 by the reference)
SU_DEV void cal_cm(int sm, double zh, double zd, double beta, double L_MOD, double &c2, double &cm)
{
   double phim_zh = stab_phi_mom(sm, (zh - zd) / L_MOD);
   double phi_hat = (phim_zh != F(0.)) ? SU_KAPPA / (F(2.) * beta * phim_zh) : F(1.);
   if (phi_hat >= F(1.)) phi_hat = F(1.);
   c2 = F(0.5);
   cm = (F(1.) - phi_hat) * exp(c2 / F(2.));
}

// rslprof.f95:508-557
SU_DEV void cal_ch(int sm, double zh, double zd, double beta, double L_MOD, double Scc, double &c2h, double &ch)
{
   double phih_zh = stab_phi_heat(sm, (zh - zd) / L_MOD);
   double phi_hat = (phih_zh != F(0.)) ? SU_KAPPA * Scc / (F(2.) * beta * phih_zh) : F(1.);
   if (phi_hat >= F(1.)) phi_hat = F(1.);
   c2h = F(0.5);
   ch = (F(1.) - phi_hat) * exp(c2h / F(2.));
}

// cal_psim_hat(z = zh), rslprof.f95:366-424, as a loop.
SU_BIG5 double cal_psim_hat_zh(int sm, double zh, double L_MOD, double beta, double Lc)
{
   const double dzh = F(0.1) / F(2.);
   const double zd = cal_zd_RSL(zh, beta, Lc);
   const double elm = cal_elm_RSL(beta, Lc);
   double c2, cm;
   cal_cm(sm, zh, zd, beta, L_MOD, c2, cm);
   // the two per-level divisions by loop constants become multiplications (<= 1 ulp on the arguments)
   const double invL = 1.0 / L_MOD;
   const double nae = (F(-1.) * c2 * beta) / elm;
   // Levels z_0 = zh, z_{k+1} = 1.01 z_k up to the first z_n > 100, where the recursion returns 0;
   // psi(z_k) = psi(z_{k+1}) + term(z_{k+1}) + term(z_k) with
   //    term(z) = dz/2 * phim(z) * (1 - phim_hat(z)) / (z - zd),  phim_hat(z) = 1 - cm*exp(-c2*beta*(z - zd)/elm).
   // The recursion adds the terms from the top level down.  Walking down needs the whole ladder of z_k in
   // (local) memory -- 2 x ~200 DRAM-backed accesses per call, 6 % of the kernel's local traffic -- so the terms
   // are accumulated upwards instead, each level formed from the previous one in registers.  Same terms, other
   // order of the additions: the sum differs from the recursion's by a few ulp (<= n * 2^-53 relative), far
   // inside the 1e-10 parity tolerance (numerics).
   double psi = F(0.);
   double z = zh;
   // term(z) with the division by (z - zd) as a multiplication by a Newton reciprocal (<= 1 ulp)
#define SU_PSIHAT_TERM(zz_) (dzh * stab_phi_mom(sm, ((zz_) - zd) * invL) * (1 - (1 - cm * su_exp_fast(nae * ((zz_) - zd)))) * su_rcp_fast((zz_) - zd))
   double term_z = SU_PSIHAT_TERM(z);
   // two levels per pass: their evaluations are independent, which gives the FP64 pipe two chains to overlap
   while (!(z > 100)) {
      const double z1 = F(1.01) * z;
      const double z2 = F(1.01) * z1;
      const double t1 = SU_PSIHAT_TERM(z1);
      const double t2 = SU_PSIHAT_TERM(z2);
      psi = psi + t1;
      psi = psi + term_z;
      if (z1 > 100) { z = z1; break; }
      psi = psi + t2;
      psi = psi + t1;
      term_z = t2;
      z = z2;
   }
#undef SU_PSIHAT_TERM
   return psi;
}

// rslprof.f95:921-952
SU_DEV double cal_beta_lc(int sm, double beta0, double lc_over_l)
{
   double beta_x = 0, phim = 1, err = 1;
   int it = 1;
   while ((err > F(0.001)) && (it < 20)) {
      double beta_x0 = beta0 / phim;
      phim = stab_phi_mom(sm, (beta_x0 * beta_x0) * lc_over_l);
      beta_x = beta0 / phim;
      err = fabs(beta_x - beta_x0);
      it = it + 1;
   }
   return beta_x;
}

struct SuRSL {
   double T2_C, q2_gkg, U10_ms, RH2;
};

// interp_z, rslprof.f95:315-353 (the reference's `if (idx > 0)` tests an uninitialised variable that is
// 0 in practice: the interpolation branch is always taken)
SU_INL void bracket_z(double z_x, const double *z, int &idx_low, int &idx_high)
{
   idx_low = -1;
   idx_high = -1;
   double bl = 0, bh = 0;
   for (int i = 0; i < SU_NZ; i++) {
      double dif = z[i] - z_x;
      if (dif < F(0.) && (idx_low < 0 || dif > bl)) { idx_low = i; bl = dif; }
      if (dif > F(0.) && (idx_high < 0 || dif < bh)) { idx_high = i; bh = dif; }
   }
   if (idx_low < 0) idx_low = 0;     // reference would index element 0 (out of bounds); keep in range
   if (idx_high < 0) idx_high = 0;
}
SU_INL double interp_2(double z_x, double z_low, double z_high, double v_low, double v_high)
{
   double dz = z_high - z_low;
   double slope = (v_high - v_low) / dz;
   return v_low + (z_x - z_low) * slope;
}
SU_DEV double interp_z(double z_x, const double *z, const double *v)
{
   int idx_low, idx_high;
   bracket_z(z_x, z, idx_low, idx_high);
   return interp_2(z_x, z[idx_low], z[idx_high], v[idx_low], v[idx_high]);
}

// the vertical grid of RSLProfile (rslprof.f95:153-191), element by element
struct SuRslGrid {
   double zH, zMeas, dzc, dza;
   int nz_can, idx_can;
   // zarray(i0 + 1) before the idx_can / zMeas overrides
   SU_INL double pre(int i0) const
   {
      const int i = i0 + 1;
      if (i <= nz_can) {
         if (i == nz_can) return zH;
         if (i == 1 && dzc > 2) return F(1.999);
         return dzc * i;
      }
      return zH + (i - nz_can) * dza;
   }
   SU_INL double at(int i0) const
   {
      if (i0 == SU_NZ - 1) return zMeas;
      if (i0 == idx_can - 1) return zH;
      return pre(i0);
   }
};
SU_INL void bracket_grid(double z_x, const SuRslGrid &g, int &idx_low, int &idx_high)
{
   idx_low = -1;
   idx_high = -1;
   double bl = 0, bh = 0;
   for (int i = 0; i < SU_NZ; i++) {
      double dif = g.at(i) - z_x;
      if (dif < F(0.) && (idx_low < 0 || dif > bl)) { idx_low = i; bl = dif; }
      if (dif > F(0.) && (idx_high < 0 || dif < bh)) { idx_high = i; bh = dif; }
   }
   if (idx_low < 0) idx_low = 0;
   if (idx_high < 0) idx_high = 0;
}

// RSLProfile, rslprof.f95:13-313.  line133 may be NULL when the 133-column profile is not wanted.
SU_BIG4 void RSLProfile(int sm, double Zh, double z0m, double zdm, double L_MOD, const double *sfr, double FAI,
                         double avcp, double lv_J_kg, double avdens, double avU1, double Temp_C, double avRH,
                         double Press_hPa, double zMeas, double qh, double qe, double *cache10, SuRSL &o, double *line133)
{
   // ---- RSL_cal_prms rslprof.f95:774-868
   const double zH_RSL = SU_MAX(Zh, F(0.4));
   const double PAI = sfr[BLDG] + sfr[CONIF] + sfr[DECID];
   const double sfr_tr = sfr[CONIF] + sfr[DECID];
   double Lc = (F(1.) - PAI) / FAI * zH_RSL;
   Lc = (Lc > F(0.5) * zH_RSL) ? Lc : F(0.5) * zH_RSL;
   double lc_over_L = Lc / L_MOD;
   if (lc_over_L > 0) lc_over_L = SU_MIN(F(2.), lc_over_L);
   else lc_over_L = SU_MAX(F(-2.), lc_over_L);
   double L_MOD_RSL = Lc / lc_over_L;
   double beta;
   {   // cal_beta_RSL rslprof.f95:870-919
      double betaN2 = (PAI > 0) ? F(0.30) * sfr_tr / PAI + (PAI - sfr_tr) / PAI * F(0.4) : F(0.35);
      double betaHF = cal_beta_lc(sm, betaN2, lc_over_L);
      if (lc_over_L > F(-0.1)) {
         beta = betaHF;
      } else {
         double betaNL = cal_beta_lc(sm, F(0.4) / F(2.), lc_over_L);
         beta = betaNL + ((betaHF - betaNL) / (F(1.) + F(4.) * pow(fabs(lc_over_L - F(-0.1)), F(1.5))));
      }
      if (beta > F(0.5)) beta = F(0.5);
   }
   const double Scc = F(0.5) + F(0.3) * tanh(F(2.) * lc_over_L);
   const double f = F(0.5) * (pow(F(1.) + F(4.) * F(0.1) * Scc, F(0.5))) - F(0.5);
   double zd_RSL = cal_zd_RSL(zH_RSL, beta, Lc);
   double elm = cal_elm_RSL(beta, Lc);
   const bool flag_RSL = ((F(1.) - PAI) / FAI <= 18) && ((F(1.) - PAI) / FAI > F(.021)) && (zH_RSL > 6);
   double z0_RSL;

   // ---- vertical grid rslprof.f95:153-191.  The 30 heights follow from (zH_RSL, zMeas); they are generated on
   // demand (SuRslGrid::at returns exactly what the reference's zarray holds) and only stored for the profile line.
   SuRslGrid grid;
   grid.zH = zH_RSL; grid.zMeas = zMeas;
   grid.nz_can = (zH_RSL <= 2) ? 3 : ((zH_RSL <= 10) ? 10 : 15);
   grid.dzc = zH_RSL / grid.nz_can;
   grid.dza = (zMeas - zH_RSL) / (SU_NZ - grid.nz_can);
   grid.idx_can = 1;
   {
      double mn = fabs(grid.pre(0) - zH_RSL);
      for (int z = 1; z < SU_NZ; z++) {
         double dif = fabs(grid.pre(z) - zH_RSL);
         if (dif < mn) { mn = dif; grid.idx_can = z + 1; }
      }
   }
   const int nz_can = grid.nz_can, idx_can = grid.idx_can;
   const double dz = grid.dza;

   // The 30-level profile is only an output (dataoutLineRSL); the step itself consumes T2, q2 and U10, which
   // interp_z takes from the two levels bracketing 2 m and 10 m.  Without the profile line only those levels
   // (and the canopy-top level the in-canopy ones hang on) are evaluated -- the same expressions, so the same bits.
   const bool full = (line133 != nullptr);
   int need[5];
   double need_pm[5] = {0, 0, 0, 0, 0}, need_ph[5] = {0, 0, 0, 0, 0};
   bracket_grid(2.0, grid, need[0], need[1]);
   bracket_grid(10.0, grid, need[2], need[3]);
   need[4] = idx_can - 1;
   double psihatm_z[SU_NZ], psihath_z[SU_NZ], zarray[SU_NZ];
   if (full)
      for (int z = 0; z < SU_NZ; z++) { psihatm_z[z] = 0; psihath_z[z] = 0; zarray[z] = grid.at(z); }

   if (flag_RSL) {
      // cal_z0_RSL rslprof.f95:730-771.  z0_RSL is a pure function of (zH, L_MOD_RSL, beta, Lc); when
      // Lc/L saturates at +-2 (rslprof.f95:834-838) those repeat from step to step, so the last result per
      // stability sign is kept with its exact inputs as the key.
      {
         double *ck = cache10 + ((lc_over_L > 0) ? 0 : 5);
         if (ck[0] == zH_RSL && ck[1] == L_MOD_RSL && ck[2] == beta && ck[3] == Lc) {
            z0_RSL = ck[4];
         } else {
            double psimZh = stab_psi_mom_v(sm, (zH_RSL - zd_RSL) / L_MOD_RSL);
            double psihatm_Zh = cal_psim_hat_zh(sm, zH_RSL, L_MOD_RSL, beta, Lc);
            z0_RSL = F(0.1) * zH_RSL;
            double err = F(10.);
            int it = 1;
            const double e1 = exp(F(-1.) * SU_KAPPA / beta), e3 = exp(psihatm_Zh);
            while ((err > F(0.001)) && (it < 10)) {
               double z0x = z0_RSL;
               double psimz0 = stab_psi_mom_v(sm, z0x / L_MOD_RSL);
               z0_RSL = (zH_RSL - zd_RSL) * e1 * exp(F(-1.) * psimZh + psimz0) * e3;
               err = fabs(z0x - z0_RSL);
               it = it + 1;
            }
            z0_RSL = (z0_RSL > 1e-2) ? z0_RSL : 1e-2;
            ck[0] = zH_RSL; ck[1] = L_MOD_RSL; ck[2] = beta; ck[3] = Lc; ck[4] = z0_RSL;
         }
      }
      double c2h, ch, c2, cm;
      cal_ch(sm, zH_RSL, zd_RSL, beta, L_MOD_RSL, Scc, c2h, ch);
      cal_cm(sm, zH_RSL, zd_RSL, beta, L_MOD_RSL, c2, cm);
      // Taylor integration downward from the forcing height (rslprof.f95:198-214); the z+1 terms of one
      // level are the z terms of the level above.  c2 == c2h == 0.5 (rslprof.f95:497,549), so the momentum and
      // heat terms share one EXP per level.
      double zzp = grid.at(SU_NZ - 1);
      double ex = su_exp_fast(F(-1.) * c2 * beta * (zzp - zd_RSL) / elm);
      double tmp = dz / F(2.) * stab_phi_mom(sm, (zzp - zd_RSL) / L_MOD_RSL) * (cm * ex) / (zzp - zd_RSL);
      double thp = dz / F(2.) * stab_phi_heat(sm, (zzp - zd_RSL) / L_MOD_RSL) * (ch * ex) / (zzp - zd_RSL);
      double pm_run = 0, ph_run = 0;   // psihat at the level above
      for (int z = SU_NZ - 1; z >= idx_can; z--) {
         double zz = grid.at(z - 1);
         ex = su_exp_fast(F(-1.) * c2 * beta * (zz - zd_RSL) / elm);
         double tm = dz / F(2.) * stab_phi_mom(sm, (zz - zd_RSL) / L_MOD_RSL) * (cm * ex) / (zz - zd_RSL);
         double th = dz / F(2.) * stab_phi_heat(sm, (zz - zd_RSL) / L_MOD_RSL) * (ch * ex) / (zz - zd_RSL);
         pm_run = pm_run + tmp;
         pm_run = pm_run + tm;
         ph_run = ph_run + thp;
         ph_run = ph_run + th;
         if (full) { psihatm_z[z - 1] = pm_run; psihath_z[z - 1] = ph_run; }
         else
            for (int j = 0; j < 5; j++)
               if (need[j] == z - 1) { need_pm[j] = pm_run; need_ph[j] = ph_run; }
         tmp = tm;
         thp = th;
      }
   } else {
      L_MOD_RSL = L_MOD;
      z0_RSL = z0m;
      zd_RSL = zdm;
      if (L_MOD_RSL > 0) Lc = SU_MIN(F(2.2) * SU_KAPPA / beta * L_MOD_RSL, Lc);
      else Lc = SU_MAX(-(2 / (beta * beta) * L_MOD_RSL), Lc);
      elm = cal_elm_RSL(beta, Lc);
   }

   const double psimz0 = stab_psi_mom_v(sm, z0_RSL / L_MOD_RSL);
   const double psimza = stab_psi_mom_v(sm, (zMeas - zd_RSL) / L_MOD_RSL);
   const double psihza = stab_psi_heat(sm, (zMeas - zd_RSL) / L_MOD_RSL);
   double UStar_RSL = avU1 * SU_KAPPA / (log((zMeas - zd_RSL) / z0_RSL) - psimza + psimz0 + F(0.));
   UStar_RSL = (UStar_RSL > 0.15) ? UStar_RSL : 0.15;
   const double TStar_RSL = F(-1.) * (qh / (avcp * avdens)) / UStar_RSL;
   const double qStar_RSL = F(-1.) * (qe / lv_J_kg * avdens) / UStar_RSL;
   const double qa_gkg = RH2qa(avRH / 100, Press_hPa, Temp_C);

   if (!full) {
      // dimensionless U, T (= q above the canopy) at the needed levels
      double Ud[5], Td[5], qd[5];
      for (int j = 4; j >= 0; j--) {
         const int zi = need[j];
         const double zz = grid.at(zi);
         if (zi >= idx_can - 1) {
            double psimz = stab_psi_mom_v(sm, (zz - zd_RSL) / L_MOD_RSL);
            double psihz = stab_psi_heat(sm, (zz - zd_RSL) / L_MOD_RSL);
            Ud[j] = (log((zz - zd_RSL) / z0_RSL) - psimz + psimz0 + need_pm[j]) / SU_KAPPA;
            Td[j] = (log((zz - zd_RSL) / (zMeas - zd_RSL)) - psihz + psihza + need_ph[j] - F(0.)) / SU_KAPPA;
            qd[j] = Td[j];
         } else {   // in-canopy level: hangs on the canopy-top level (j = 4, evaluated first)
            double t_h = Scc * TStar_RSL / (beta * f);
            double q_h = Scc * qStar_RSL / (beta * f);
            Ud[j] = Ud[4] * exp(beta * (zz - zH_RSL) / elm);
            double e = exp(beta * f * (zz - zH_RSL) / elm);
            Td[j] = Td[4] + (t_h * e - t_h) / TStar_RSL;
            qd[j] = qd[4] + (q_h * e - q_h) / qStar_RSL;
         }
      }
      for (int j = 0; j < 4; j++) {
         Ud[j] = Ud[j] * UStar_RSL;
         Td[j] = Td[j] * TStar_RSL + Temp_C;
         qd[j] = (qd[j] * qStar_RSL + qa_gkg / F(1000.)) * F(1000.);
      }
      o.T2_C = interp_2(2.0, grid.at(need[0]), grid.at(need[1]), Td[0], Td[1]);
      o.q2_gkg = interp_2(2.0, grid.at(need[0]), grid.at(need[1]), qd[0], qd[1]);
      o.U10_ms = interp_2(10.0, grid.at(need[2]), grid.at(need[3]), Ud[2], Ud[3]);
      o.RH2 = qa2RH(o.q2_gkg, Press_hPa, o.T2_C);
      return;
   }
   double U[SU_NZ], T[SU_NZ];   // T and q share the same dimensionless profile (rslprof.f95:250-251)
   for (int z = idx_can; z <= SU_NZ; z++) {
      double zz = zarray[z - 1];
      double psimz = stab_psi_mom_v(sm, (zz - zd_RSL) / L_MOD_RSL);
      double psihz = stab_psi_heat(sm, (zz - zd_RSL) / L_MOD_RSL);
      U[z - 1] = (log((zz - zd_RSL) / z0_RSL) - psimz + psimz0 + psihatm_z[z - 1]) / SU_KAPPA;
      T[z - 1] = (log((zz - zd_RSL) / (zMeas - zd_RSL)) - psihz + psihza + psihath_z[z - 1] - psihath_z[SU_NZ - 1]) / SU_KAPPA;
   }
   double q[SU_NZ];
   for (int z = idx_can; z <= SU_NZ; z++) q[z - 1] = T[z - 1];
   if (idx_can > 1) {
      double t_h = Scc * TStar_RSL / (beta * f);
      double q_h = Scc * qStar_RSL / (beta * f);
      for (int z = 1; z <= idx_can - 1; z++) {
         double zz = zarray[z - 1];
         U[z - 1] = U[idx_can - 1] * exp(beta * (zz - zH_RSL) / elm);
         double e = exp(beta * f * (zz - zH_RSL) / elm);
         T[z - 1] = T[idx_can - 1] + (t_h * e - t_h) / TStar_RSL;
         q[z - 1] = q[idx_can - 1] + (q_h * e - q_h) / qStar_RSL;
      }
   }
   for (int z = 0; z < SU_NZ; z++) {
      U[z] = U[z] * UStar_RSL;
      T[z] = T[z] * TStar_RSL + Temp_C;
      q[z] = (q[z] * qStar_RSL + qa_gkg / F(1000.)) * F(1000.);
   }
   o.T2_C = interp_z(2.0, zarray, T);
   o.q2_gkg = interp_z(2.0, zarray, q);
   o.U10_ms = interp_z(10.0, zarray, U);
   o.RH2 = qa2RH(o.q2_gkg, Press_hPa, o.T2_C);

   if (line133) {
      int k = 0;
      for (int z = 0; z < SU_NZ; z++) line133[k++] = zarray[z];
      for (int z = 0; z < SU_NZ; z++) line133[k++] = U[z];
      for (int z = 0; z < SU_NZ; z++) line133[k++] = T[z];
      for (int z = 0; z < SU_NZ; z++) line133[k++] = q[z];
      line133[k++] = L_MOD_RSL; line133[k++] = zH_RSL; line133[k++] = Lc; line133[k++] = beta; line133[k++] = zd_RSL;
      line133[k++] = z0_RSL; line133[k++] = elm; line133[k++] = Scc; line133[k++] = f; line133[k++] = UStar_RSL;
      line133[k++] = FAI; line133[k++] = PAI; line133[k++] = flag_RSL ? 1.0 : 0.0;
   }
}
