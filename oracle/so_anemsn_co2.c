/* ORACLE (test infrastructure only -- see suews_oracle.h).
 * Restates SUEWS-SourceCode/suews_phys_anemsn.f95 (:29-338) and
 * suews_phys_biogenco2.f95 (:33-231). */
#include "suews_oracle.h"
#include <math.h>

#define FMAX(a, b) ((a) > (b) ? (a) : (b))

static int em_in(int m, int lo, int hi) /* any decade 0..4: m in [lo+10k, hi+10k] */
{
   int k;
   for (k = 0; k <= 40; k += 10)
      if (m >= lo + k && m <= hi + k) return 1;
   return 0;
}
static int em_is(int m, int a, int b) /* m == a+10k or b+10k */
{
   int k;
   for (k = 0; k <= 40; k += 10)
      if (m == a + k || m == b + k) return 1;
   return 0;
}

void so_AnthropogenicEmissions(SoCtx *cx, double CO2PointSource, int EmissionsMethod, int it, int imin, int DLS,
                               const int DayofWeek_id[3], double EF_umolCO2perJ, const double FcEF_v_kgkm[2],
                               double EnEF_v_Jkm, double TrafficUnits, double FrFossilFuel_Heat,
                               double FrFossilFuel_NonHeat, double MinFCMetab, double MaxFCMetab, double MinQFMetab,
                               double MaxQFMetab, const double PopDensDaytime[2], double PopDensNighttime,
                               double Temp_C, const double HDD_id[12], const double Qf_A[2], const double Qf_B[2],
                               const double Qf_C[2], const double AH_MIN[2], const double AH_SLOPE_Heating[2],
                               const double AH_SLOPE_Cooling[2], const double BaseT_Heating[2],
                               const double BaseT_Cooling[2], const double TrafficRate[2], const double QF0_BEU[2],
                               double *QF_SAHP, double *Fc_anthro, double *Fc_metab, double *Fc_traff,
                               double *Fc_build, double *Fc_point, const double AHProf_24hr[48],
                               const double HumActivity_24hr[48], const double TraffProf_24hr[48],
                               const double PopProf_24hr[48], double SurfaceArea)
{
   int iu, ih, u;
   double HDD_daily, CDD_daily, Tair_avg_daily, NumCapita[2] = {0, 0}, DP_x_RhoPop, DP_x_RhoPop_traff;
   double QF_build, QF_metab, QF_traff, QF_SAHP_base = 0, QF_SAHP_heating, QF_SAHP_cooling;
   double PopDorNorT, ActDorNorT, TraffDorNorT, AHDorNorT;

   QF_traff = 0;
   QF_SAHP_heating = 0;
   QF_SAHP_cooling = 0;

   HDD_daily = HDD_id[6];
   CDD_daily = HDD_id[7];
   Tair_avg_daily = HDD_id[9];

   if (PopDensDaytime[0] >= 0 && PopDensNighttime >= 0) NumCapita[0] = (PopDensDaytime[0] + PopDensNighttime) / 2;
   if (PopDensDaytime[1] >= 0 && PopDensNighttime >= 0) NumCapita[1] = (PopDensDaytime[1] + PopDensNighttime) / 2;

   ih = it - DLS;
   if (ih < 0) ih = 23;
   iu = 1;
   if (DayofWeek_id[0] == 1 || DayofWeek_id[0] == 7) iu = 2;
   u = iu - 1;

   PopDorNorT = so_get_Prof_SpecTime_inst(ih, imin, 0, &PopProf_24hr[24 * u]);
   ActDorNorT = so_get_Prof_SpecTime_inst(ih, imin, 0, &HumActivity_24hr[24 * u]);
   TraffDorNorT = so_get_Prof_SpecTime_mean(ih, imin, 0, &TraffProf_24hr[24 * u]);
   AHDorNorT = so_get_Prof_SpecTime_mean(ih, imin, 0, &AHProf_24hr[24 * u]);

   DP_x_RhoPop = AHDorNorT * NumCapita[u];

   QF_metab = (PopDensNighttime * MinQFMetab * ((2 - ActDorNorT) + (2 - PopDorNorT)) / 2 +
               PopDensDaytime[u] * MaxQFMetab * ((ActDorNorT - 1) + (PopDorNorT - 1)) / 2) /
              10000;
   *Fc_metab = (PopDensNighttime * MinFCMetab * ((2 - ActDorNorT) + (2 - PopDorNorT)) / 2 +
                PopDensDaytime[u] * MaxFCMetab * ((ActDorNorT - 1) + (PopDorNorT - 1)) / 2) /
               10000;

   if (em_is(EmissionsMethod, 1, 4)) {
      if (Temp_C < BaseT_Heating[u]) {
         QF_SAHP_heating = (AH_SLOPE_Heating[u] * (BaseT_Heating[u] - Temp_C)) * AHDorNorT;
      } else {
         QF_SAHP_heating = 0;
      }
      QF_SAHP_base = AH_MIN[u] * AHDorNorT;
      QF_SAHP_cooling = 0;
   } else if (em_is(EmissionsMethod, 2, 5)) {
      QF_SAHP_base = (Qf_A[u]) * DP_x_RhoPop;
      QF_SAHP_heating = (Qf_C[u] * HDD_daily) * DP_x_RhoPop;
      QF_SAHP_cooling = (Qf_B[u] * CDD_daily) * DP_x_RhoPop;
   } else if (em_is(EmissionsMethod, 3, 6)) {
      QF_SAHP_base = AH_MIN[u] * AHDorNorT;
      if (Tair_avg_daily < BaseT_Heating[u]) {
         QF_SAHP_heating = (AH_SLOPE_Heating[u] * (BaseT_Heating[u] - Tair_avg_daily)) * AHDorNorT;
         QF_SAHP_cooling = 0;
      } else if (Tair_avg_daily > BaseT_Cooling[u]) {
         QF_SAHP_heating = 0;
         QF_SAHP_cooling = (AH_SLOPE_Cooling[u] * (Tair_avg_daily - BaseT_Cooling[u])) * AHDorNorT;
      } else {
         QF_SAHP_heating = 0;
         QF_SAHP_cooling = 0;
      }
   }

   *QF_SAHP = QF_SAHP_base + QF_SAHP_heating + QF_SAHP_cooling;

   if (em_in(EmissionsMethod, 1, 3)) {
      if ((QF_SAHP_base - QF_metab) > 0) {
         QF_build = QF_SAHP_base * QF0_BEU[u] + QF_SAHP_heating + QF_SAHP_cooling;
      } else {
         so_error_hint(cx, 69, "QF metab exceeds base QF.", QF_metab, QF_SAHP_base, -55);
         QF_build = QF_SAHP_heating + QF_SAHP_cooling + (QF_SAHP_base - QF_metab);
      }
      (void)QF_build;
      *Fc_build = QF_SAHP_heating * FrFossilFuel_Heat * EF_umolCO2perJ;
      if ((QF_SAHP_base - QF_metab) > 0) {
         *Fc_build = *Fc_build + QF_SAHP_base * QF0_BEU[u] * FrFossilFuel_NonHeat * EF_umolCO2perJ;
      }
      QF_traff = QF_SAHP_base * (F(1.0) - QF0_BEU[u]) - QF_metab;
      *Fc_traff = QF_traff / EnEF_v_Jkm * FcEF_v_kgkm[u] * F(1e3) * F(1e6) / 44;
      if (CO2PointSource > 0) {
         *Fc_point = CO2PointSource * F(1e3) * F(1e6) / (12 * 60 * 60 * 24 * SurfaceArea);
      } else {
         *Fc_point = 0;
      }
      *Fc_anthro = *Fc_metab + *Fc_traff + *Fc_build + *Fc_point;
   } else if (em_in(EmissionsMethod, 4, 6)) {
      if (TrafficUnits == 1) {
         QF_traff = TrafficRate[u] / (60 * 60 * 24) * EnEF_v_Jkm * TraffDorNorT;
         *Fc_traff = TrafficRate[u] / (60 * 60 * 24) * FcEF_v_kgkm[u] * F(1e3) * F(1e6) / 44 * TraffDorNorT;
      } else if (TrafficUnits == 2) {
         DP_x_RhoPop_traff = TraffDorNorT * NumCapita[u] / 10000;
         QF_traff = TrafficRate[u] / (60 * 60 * 24) * EnEF_v_Jkm * DP_x_RhoPop_traff;
         *Fc_traff = TrafficRate[u] / (60 * 60 * 24) * FcEF_v_kgkm[u] * F(1e3) * F(1e6) / 44 * DP_x_RhoPop_traff;
      } else {
         so_error_hint(cx, 75, "Check TrafficUnits", TrafficUnits, -9990.0, -55);
      }
      QF_build = ((QF_SAHP_base * QF0_BEU[u] + QF_SAHP_heating + QF_SAHP_cooling) / DP_x_RhoPop) *
                 (PopDensNighttime * (2 - PopDorNorT) + PopDensDaytime[u] * (PopDorNorT - 1));
      *Fc_build = QF_SAHP_heating * FrFossilFuel_Heat * EF_umolCO2perJ;
      *Fc_build = *Fc_build + QF_SAHP_base * QF0_BEU[u] * FrFossilFuel_NonHeat * EF_umolCO2perJ;
      if (CO2PointSource > 0) {
         *Fc_point = CO2PointSource * F(1e3) * F(1e6) / (12 * 60 * 60 * 24 * SurfaceArea);
      } else {
         *Fc_point = 0;
      }
      QF_SAHP_base = QF_SAHP_base + QF_traff + QF_metab;
      *QF_SAHP = QF_metab + QF_traff + QF_build;
      *Fc_anthro = *Fc_metab + *Fc_traff + *Fc_build + *Fc_point;
   }
}

void so_CO2_biogen(SoCtx *cx, const double alpha_bioCO2[NVEG], const double alpha_enh_bioCO2[NVEG], double avkdn,
                   const double beta_bioCO2[NVEG], const double beta_enh_bioCO2[NVEG], double dectime,
                   int EmissionsMethod, double gfunc, double gfunc2, int gsmodel, int id, int it,
                   const double LAI_id[NVEG], const double LAIMin[NVEG], const double LAIMax[NVEG],
                   const double min_res_bioCO2[NVEG], const double resp_a[NVEG], const double resp_b[NVEG],
                   const double sfr[NSURF], const double SnowFrac[NSURF], double t2, double Temp_C,
                   const double theta_bioCO2[NVEG], double *Fc_biogen, double *Fc_photo, double *Fc_respi)
{
   const double JtoumolPAR = F(4.6), KdntoPAR = F(0.46);
   double PAR, Pho = 0, Res, Res_surf, VegFracSum, active_veg_fr[3], active_veg_fr0[3], a2, b2, t2v, s4, q;
   int iv;
   (void)cx; (void)dectime; (void)id; (void)it; (void)LAIMin;

   PAR = JtoumolPAR * KdntoPAR * avkdn;
   VegFracSum = sfr[CONIF] + sfr[DECID] + sfr[GRASS];
   for (iv = 0; iv < 3; iv++) {
      active_veg_fr[iv] = (sfr[iv + 2] * (1 - SnowFrac[iv + 2])) * (LAI_id[iv] / LAIMax[iv]);
      active_veg_fr0[iv] = (sfr[iv + 2] * (1 - SnowFrac[iv + 2])) * LAI_id[iv];
   }

   if (EmissionsMethod >= 11 && EmissionsMethod <= 16) {
      *Fc_photo = 0;
      for (iv = 0; iv < 3; iv++) {
         q = -beta_bioCO2[iv] * alpha_bioCO2[iv] * PAR / (alpha_bioCO2[iv] * PAR + beta_bioCO2[iv]);
         *Fc_photo = *Fc_photo + q * active_veg_fr[iv];
      }
   } else if (EmissionsMethod >= 21 && EmissionsMethod <= 26) {
      Pho = 0;
      for (iv = 0; iv < 3; iv++) {
         q = alpha_bioCO2[iv] * PAR + beta_bioCO2[iv];
         q = -(1 / (2 * theta_bioCO2[iv]) *
               (alpha_bioCO2[iv] * PAR + beta_bioCO2[iv] -
                sqrt(q * q - 4 * alpha_bioCO2[iv] * beta_bioCO2[iv] * theta_bioCO2[iv] * PAR)));
         Pho = Pho + q * active_veg_fr[iv];
      }
      *Fc_photo = Pho;
   } else if (EmissionsMethod >= 31 && EmissionsMethod <= 36) {
      s4 = (sfr[CONIF] + sfr[DECID] + sfr[GRASS] + sfr[BSOIL]);
      if (alpha_bioCO2[0] == alpha_bioCO2[1] && alpha_bioCO2[0] == alpha_bioCO2[2] &&
          beta_bioCO2[0] == beta_bioCO2[1] && beta_bioCO2[0] == beta_bioCO2[2] &&
          theta_bioCO2[0] == theta_bioCO2[1] && theta_bioCO2[0] == theta_bioCO2[2]) {
         a2 = alpha_bioCO2[0] + alpha_enh_bioCO2[0] * s4;
         b2 = -beta_bioCO2[0] + beta_enh_bioCO2[0] * s4;
         q = a2 * PAR + b2;
         Pho = -(1 / (2 * theta_bioCO2[0]) * (a2 * PAR + b2 - sqrt(q * q - 4 * a2 * b2 * theta_bioCO2[0] * PAR)));
      } else {
         a2 = (alpha_bioCO2[0] * sfr[CONIF] / VegFracSum + alpha_bioCO2[1] * sfr[DECID] / VegFracSum +
               alpha_bioCO2[2] * sfr[GRASS] / VegFracSum) /
              (alpha_bioCO2[0] + alpha_bioCO2[1] + alpha_bioCO2[2]);
         b2 = (beta_bioCO2[0] * sfr[CONIF] / VegFracSum + beta_bioCO2[1] * sfr[DECID] / VegFracSum +
               beta_bioCO2[2] * sfr[GRASS] / VegFracSum) /
              (beta_bioCO2[0] + beta_bioCO2[1] + beta_bioCO2[2]);
         t2v = (theta_bioCO2[0] * sfr[CONIF] / VegFracSum + theta_bioCO2[1] * sfr[DECID] / VegFracSum +
                theta_bioCO2[2] * sfr[GRASS] / VegFracSum) /
               (theta_bioCO2[0] + theta_bioCO2[1] + theta_bioCO2[2]);
         a2 = a2 + alpha_enh_bioCO2[0] * s4;
         b2 = -b2 + beta_enh_bioCO2[0] * s4;
         q = a2 * PAR + b2;
         Pho = -(1 / (2 * t2v) * (a2 * PAR + b2 - sqrt(q * q - 4 * a2 * b2 * t2v * PAR)));
      }
      *Fc_photo = Pho * active_veg_fr[0] + Pho * active_veg_fr[1] + Pho * active_veg_fr[2];
   } else if (EmissionsMethod >= 41 && EmissionsMethod <= 46) {
      *Fc_photo = 0;
      for (iv = 0; iv < 3; iv++) *Fc_photo = *Fc_photo + active_veg_fr0[iv] * beta_bioCO2[iv];
      if (gsmodel == 1 || gsmodel == 2) {
         *Fc_photo = -(*Fc_photo) * gfunc;
      } else if (gsmodel == 3 || gsmodel == 4) {
         *Fc_photo = -(*Fc_photo) * gfunc2;
      }
   }

   Res = F(0.0);
   Res_surf = F(0.0);
   if (VegFracSum > F(0.01)) {
      for (iv = 0; iv < 3; iv++) {
         if (sfr[2 + iv] > F(0.005)) {
            if (gsmodel == 1 || gsmodel == 2) {
               Res_surf = FMAX(min_res_bioCO2[iv], resp_a[iv] * exp(resp_b[iv] * Temp_C));
            } else if (gsmodel == 3 || gsmodel == 4) {
               Res_surf = FMAX(min_res_bioCO2[iv], resp_a[iv] * exp(resp_b[iv] * t2));
            }
            Res = Res + Res_surf * sfr[2 + iv] / VegFracSum;
         }
      }
   }
   *Fc_respi = Res * (sfr[CONIF] + sfr[DECID] + sfr[GRASS] + sfr[BSOIL]);
   *Fc_biogen = *Fc_photo + *Fc_respi;
}
