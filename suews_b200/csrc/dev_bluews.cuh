// dev_bluews.cuh -- BLUEWS, the convective boundary-layer slab model, on the device.
//
// Reference: SUEWS-SourceCode/suews_phys_bluews.f95 (CBL :69-296, NBL :482-622, RKUTTA :631-704, diff :708-849), called by the
// reference's grid loop right after SUEWS_cal_Main (suews_ctrl_calculations.f95:194-218) with that step's QH / QE (SUEWS's,
// or the LUMPS pair for QH_choice 2), its friction velocity, and the air density / heat capacity / latent heat the step
// left in MODULE moist.  It is a per-cell sequential model fed by the step's results, so here it is the last stage of the
// step (phase F3): at night (Kdown < 5) the slab is reset from the day's row of the initial-data table and a 200 m line is
// written; by day height, theta, q and their values above the inversion advance by one Runge-Kutta step over the
// timestep.  Nothing of it feeds back into the surface scheme (SUEWS_Translate re-reads the forcing every step).
// Its variables (MODULE cbl_MODULE) are per-cell state in the storage record (cbl_* of SUEWS_ESTM_STATE_FIELDS), the
// CBLInput.nml switches are class fields, IniCBLdata is a table shared by the batch (suews_b200_set_cbl_initial).
// Built: InitialData_use = 2 (all start values from the table) and 1 (height and gradients from the table on the days it
// lists, theta and q diagnosed from the night's line through psih -- the value AerodynamicResistance left in MODULE sues_data,
// resist.f95:40,78, recomputed here from the step's last z/L), any EntrainmentType, isubs.  InitialData_use = 0 cannot run in
// the reference (NBL reads IniCBLdata, :523-536, allocated for 1 and 2 only, :325) and sounding files (sondeflag = 1) are not
// built: error 1007.  The reference's REAL(4) clock arithmetic and exponents are kept.
#pragma once

#define SU_CBL_C2K F(273.16)
#define SU_CBL_GAS_CT_DRY ((double)(8.31451f / 0.028965f))
#define SU_CBL_GRAV F(9.80665)

struct SuCblFlux {
   double fhbl_Kms, febl_kgkgms, fcbl, tm_K, gamt_Km, gamq_kgkgm, ustar;
};

// diff (bluews.f95:708-849)
SU_COLD void cbl_diff(const SuewsClassParams &P, const SuCblFlux &c, const double *y1, double *dyds)
{
   const double h1 = y1[0];
   const double delt_K = y1[4] - y1[1], delq_kgkg = y1[5] - y1[2], delc = 0 - y1[3];
   const double ftv_Kms = c.fhbl_Kms + F(0.61) * c.tm_K * c.febl_kgkgms;
   const double gamtv_Km = c.gamt_Km + F(0.61) * c.tm_K * c.gamq_kgkgm;
   const double deltv_K = delt_K + F(0.61) * c.tm_K * delq_kgkg;
   const double ftva_Kms = (ftv_Kms > 0.0) ? ftv_Kms : 0.0;
   const double ws = su_pow_general(h1 * ftva_Kms * SU_CBL_GRAV / c.tm_K, F(0.3333333333));
   double dhds = 0;
   const int et = P.EntrainmentType;
   if (et == 2 || ((et == 1 || et == 3 || et == 4) && deltv_K <= F(0.01))) {
      dhds = ftva_Kms / (h1 * gamtv_Km);
   } else if (et == 1) {
      const double delb = SU_CBL_GRAV * deltv_K / c.tm_K;
      dhds = (F(0.2) * (ws * ws * ws) + F(5.0) * (c.ustar * c.ustar * c.ustar)) / (h1 * delb);
   } else if (et == 4) {
      dhds = F(0.2) * ftva_Kms / deltv_K;
      if (P.isubs == 1) dhds = dhds + P.wsb;
   } else if (et == 3) {
      const double cu = F(1.33) * c.ustar;
      const double qs3 = ws * ws * ws + cu * cu * cu;
      const double qs2 = su_pow_general(qs3, F(0.6666666667));
      const double delb = SU_CBL_GRAV * deltv_K / c.tm_K;
      dhds = (F(0.18) * qs3) / (F(0.80) * qs2 + h1 * delb);
   }
   const double dh = (P.isubs == 1) ? (dhds - P.wsb) : dhds;
   dyds[0] = dhds;
   dyds[1] = c.fhbl_Kms / h1 + delt_K * dh / h1;
   dyds[2] = c.febl_kgkgms / h1 + delq_kgkg * dh / h1;
   dyds[3] = c.fcbl / h1 + delc * dh / h1;
   dyds[4] = c.gamt_Km * dh;
   dyds[5] = c.gamq_kgkgm * dh;
}

// qsatf (suews_util_meteo.f95:495-516); its T > 55 warning (ErrorHint 34) included
SU_INL double cbl_qsatf(SuDiag &dg, double T, double PMB)
{
   if (T > 55) su_warn(dg, 34);
   const double ES = F(6.106) * exp(F(17.27) * T / (F(237.3) + T));
   return (F(0.0180153) / F(0.028965)) * ES / PMB;
}

// CBL / NBL for one cell-step.  bl17: z theta q theta+ q+ Temp_C rh QH_use QE_use Press_hPa avu1 UStar avdens lv_J_kg avcp gamt
// gamq (the BL output line without its five date columns), or nullptr
SU_COLD void BLUEWS(SuDiag &dg, const SuewsClassParams &P, SuewsEstmState &S, const double *ini, int nline, const SuewsStepRec &R,
                    double qh_use, double qe_use, double UStar, double avdens, double lv_J_kg, double avcp, double zl_meas, double *bl17)
{
   const bool diagnosed = (P.InitialData_use == 1);
   if ((P.InitialData_use != 2 && !diagnosed) || (diagnosed && P.AerodynamicResistanceMethod != 2) || P.sondeflag != 0 || !ini ||
       nline <= 0) {
      su_fatal(dg, 1007);
      return;
   }
   if (qh_use < -900 || qe_use < -900) su_warn(dg, 22);
   const double Press_hPa = R.Press_hPa;
   double qm_gkg, qp_gkg_out, gamq_out;
   if (R.avkdn < 5) {
      // NBL: the day's row is the last one whose day of year is <= id (bluews.f95:527-532)
      int n = 0;
      for (int i = 0; i < nline; i++)
         if ((int)ini[(size_t)i * 8] <= R.id) n++;
      if (n == 0) {
         su_fatal(dg, 1007);
         return;
      }
      const double *row = ini + (size_t)(n - 1) * 8;
      qm_gkg = row[7];
      qp_gkg_out = row[5];
      gamq_out = S.cbl_gamq_kgkgm;   // the line carries the previous call's gamq (bluews.f95:567 before :609)
      S.cbl_blh_m = 200;
      S.cbl_tm_K = row[6];
      S.cbl_tp_K = row[4];
      S.cbl_gamt_Km = row[2];
   } else {
      // CBL: one Runge-Kutta step over the timestep (bluews.f95:150-205, :631-704)
      SuCblFlux c;
      const double hr = (double)((float)R.it + (float)R.imin / 60.f);
      const double secs0 = hr * F(3600.);
      const double secs1 = secs0 + (double)(float)P.tstep;
      c.fhbl_Kms = qh_use / (avdens * avcp);
      c.febl_kgkgms = qe_use / (avdens * lv_J_kg);
      c.fcbl = 0;
      if (P.CO2_included != 1) S.cbl_cm = -999;
      c.tm_K = S.cbl_tm_K; c.gamt_Km = S.cbl_gamt_Km; c.gamq_kgkgm = S.cbl_gamq_kgkgm; c.ustar = UStar;
      double y[6] = {S.cbl_blh_m, S.cbl_tm_K, S.cbl_qm_kgkg, S.cbl_cm, S.cbl_tp_K, S.cbl_qp_kgkg};
      const double coef[4] = {1.0, 0.5, 0.5, 1.0};
      double rk[6][5], arg[6], dydx[6];
      const double step = (secs1 - secs0) / 1;
      for (int nj = 0; nj < 6; nj++) rk[nj][0] = 0;
      for (int n = 0; n < 4; n++) {
         for (int nj = 0; nj < 6; nj++) arg[nj] = y[nj] + coef[n] * rk[nj][n];
         cbl_diff(P, c, arg, dydx);
         for (int nj = 0; nj < 6; nj++) rk[nj][n + 1] = step * dydx[nj];
      }
      for (int nj = 0; nj < 6; nj++)
         for (int n = 0; n < 4; n++) y[nj] = y[nj] + rk[nj][n + 1] / (6 * coef[n]);
      S.cbl_blh_m = y[0]; S.cbl_tm_K = y[1]; S.cbl_qm_kgkg = y[2]; S.cbl_cm = y[3]; S.cbl_tp_K = y[4]; S.cbl_qp_kgkg = y[5];
      qm_gkg = S.cbl_qm_kgkg * 1000;
      qp_gkg_out = S.cbl_qp_kgkg * 1000;
      gamq_out = S.cbl_gamq_kgkgm;
   }
   double Temp_C = 0, avrh = 0;
   if (bl17 || (diagnosed && R.avkdn < 5)) {
      Temp_C = S.cbl_tm_K / su_pow_general(1000 / Press_hPa, SU_CBL_GAS_CT_DRY / avcp) - SU_CBL_C2K;
      const double es_hPa = sat_vap_press_x(dg, Temp_C, Press_hPa, 1);   // (rewrites Temp_C when |T| < 0.001, as in the reference)
      avrh = 100 * ((qm_gkg * Press_hPa / (622 + qm_gkg)) / es_hPa);
      if (avrh > 100) {
         su_warn(dg, 34);
         avrh = 100;
      }
   }
   if (bl17) {
      bl17[0] = S.cbl_blh_m; bl17[1] = S.cbl_tm_K; bl17[2] = qm_gkg; bl17[3] = S.cbl_tp_K; bl17[4] = qp_gkg_out;
      bl17[5] = Temp_C; bl17[6] = avrh; bl17[7] = qh_use; bl17[8] = qe_use; bl17[9] = Press_hPa; bl17[10] = R.avU1;
      bl17[11] = UStar; bl17[12] = avdens; bl17[13] = lv_J_kg; bl17[14] = avcp; bl17[15] = S.cbl_gamt_Km; bl17[16] = gamq_out;
   }
   if (R.avkdn < 5) {
      // the slab's start values of the day (bluews.f95:562-594, :596-621)
      int n = 0;
      for (int i = 0; i < nline; i++)
         if ((int)ini[(size_t)i * 8] <= R.id) n++;
      const double *row = ini + (size_t)(n - 1) * 8;
      double qp_gkg = row[5];
      qm_gkg = row[7];
      if (!diagnosed) {
         S.cbl_blh_m = row[1];
         S.cbl_gamt_Km = row[2];
         S.cbl_tp_K = row[4];
         S.cbl_tm_K = row[6];
      } else if (row[0] == (double)R.id) {
         // InitialData_use = 1, on a day the table lists (otherwise the night's 200 m line stays)
         const double k = F(0.4);
         const double psih = stab_psi_heat(P.StabilityMethod, zl_meas);
         S.cbl_blh_m = row[1];
         S.cbl_gamt_Km = row[2];
         const double tm_K_zm = (Temp_C + SU_CBL_C2K) * su_pow_general(1000 / Press_hPa, SU_CBL_GAS_CT_DRY / avcp);
         S.cbl_tm_K = tm_K_zm - psih * qh_use / (k * UStar * avcp * avdens);
         const double es_hPa = sat_vap_press_x(dg, Temp_C, Press_hPa, 1);
         const double qm_gkg_zm = 622 * avrh / (100 * Press_hPa / es_hPa - avrh);
         const double lv = (F(2500.25) - F(2.365) * Temp_C) * 1000;
         qm_gkg = qm_gkg_zm - psih * qe_use / (k * UStar * avdens * lv);
         S.cbl_tp_K = S.cbl_tm_K;
         qp_gkg = qm_gkg;
      }
      S.cbl_gamq_kgkgm = row[3] / F(1000.);
      S.cbl_qp_kgkg = qp_gkg / 1000;
      S.cbl_qm_kgkg = qm_gkg / 1000;
      double qs = cbl_qsatf(dg, S.cbl_tp_K - SU_CBL_C2K, Press_hPa);
      if (S.cbl_qp_kgkg > qs || S.cbl_qp_kgkg < 0) S.cbl_qp_kgkg = qs;
      qs = cbl_qsatf(dg, S.cbl_tm_K - SU_CBL_C2K, Press_hPa);
      if (S.cbl_qm_kgkg > qs || S.cbl_qm_kgkg < 0) S.cbl_qm_kgkg = qs;
   }
}
