// dev_atmos.cuh -- moist-air thermodynamics, Obukhov-length iteration and resistances.
// Device restatement of suews_util_meteo.f95, suews_phys_atmmoiststab.f95, suews_phys_resist.f95.
#pragma once
#include "dev_common.cuh"

// ------------------------------------------------------------------ meteo
// suews_util_meteo.f95:25-39 (Buck 1981)
SU_DEV double sat_vap_press(double TK, double P)
{
   double TC = TK - F(273.15);
   if (TC == 0) TC = F(0.001);
   double e = F(6.1121) * exp(((F(18.729) - TC / F(227.3)) * TC) / (TC + F(257.87)));
   double f = F(1.00072) + P * (F(3.2E-6) + F(5.9E-10) * (TC * TC));
   return e * f;
}

#include "dev_latvap.cuh"   // sat_vap_press_x, psyc_const and Lat_vap (suews_util_meteo.f95:148-192, :277-346, :409-427)

// suews_util_meteo.f95:195-228
SU_COLD double sat_vap_pressIce(SuDiag &dg, double &T, double Press_hPa, int from)
{
   double es_hPa = 10;
   if (fabs(T) < F(0.001000)) {
      if (from == 1) su_warn(dg, 29);
      T = F(0.001000);
   }
   double Press_kPa = Press_hPa / 10;
   if (T < 50 && T > -40) {
      double e_mb = F(6.1115) * exp(((F(23.036) - T / F(333.7)) * T) / (T + F(279.82)));
      double f = F(1.00022) + Press_kPa * (F(3.83E-6) + F(6.4E-10) * (T * T));
      es_hPa = e_mb * f;
   } else {
      su_warn(dg, 28);
   }
   return es_hPa;
}

// suews_util_meteo.f95:444-492 (genuine double-precision literals)
SU_INL double slope_svp(double t)
{
   return 4.438099984e-1 +
          t * (2.857002636e-2 +
               t * (7.938054040e-4 + t * (1.215215065e-5 + t * (1.036561403e-7 + t * (3.532421810e-10 + -7.090244804e-13 * t)))));
}
SU_INL double slopeIce_svp(double t)
{
   return 5.030305237e-1 +
          t * (3.773255020e-2 +
               t * (1.267995369e-3 + t * (2.477563108e-5 + t * (3.005693132e-7 + t * (2.158542548e-9 + 7.131097725e-12 * t)))));
}

// suews_util_meteo.f95:518-555
SU_DEV double RH2qa(double RH_dec, double pres_hPa, double Ta_degC)
{
   double es = sat_vap_press(Ta_degC + F(273.15), pres_hPa);
   double ea = es * RH_dec;
   return F(0.622) * ea / (pres_hPa - F(0.378) * ea) * 1000;
}
SU_DEV double qa2RH(double qa_gkg, double pres_hPa, double Ta_degC)
{
   double qa_kgkg = qa_gkg / 1000;
   double es = sat_vap_press(Ta_degC + F(273.15), pres_hPa);
   double ea = 500 * pres_hPa * qa_kgkg / (311 + 189 * qa_kgkg);
   return ea / es;
}

struct SuAtm {
   double lv_J_kg, lvS_J_kg, es_hPa, Ea_hPa, VPd_hpa, dq, avcp, avdens;
};

// suews_phys_atmmoiststab.f95:9-100
SU_DEV void cal_AtmMoist(SuDiag &dg, double &Temp_C, double Press_hPa, double avRh, int &lv_sw, SuAtm &a)
{
   const double gas_ct_dry = (double)(8.31451f / 0.028965f);   // REAL(4) quotient (atmmoiststab.f95:45)
   const double gas_ct_wv = (double)(8.31451f / 0.0180153f);
   a.es_hPa = sat_vap_press_x(dg, Temp_C, Press_hPa, 1);
   a.Ea_hPa = avRh / 100 * a.es_hPa;
   a.VPd_hpa = a.es_hPa - a.Ea_hPa;
   a.dq = su_div0(F(621.97) * a.VPd_hpa, Press_hPa);   // spec_hum_def, meteo.f95:235
   double vap_dens = (a.Ea_hPa * 100 / ((Temp_C + F(273.16)) * gas_ct_wv));
   double dens_dry = ((Press_hPa - a.Ea_hPa) * 100) / (gas_ct_dry * (F(273.16) + Temp_C));
   a.avdens = (Press_hPa * 100) / (gas_ct_dry * (Temp_C + F(273.16)));
   {   // spec_heat_beer, meteo.f95:244-271
      double t = Temp_C + F(23.16);
      double cpd = F(1005.0) + (t * t) / F(3364.0);
      double t100 = Temp_C / F(100.);
      double cpm = 1859 + F(0.13) * avRh + (F(19.3) + F(0.569) * avRh) * (Temp_C / F(100.)) +
                   (F(10.) + F(0.5) * avRh) * (t100 * t100);
      if (fabs(dens_dry) < F(0.000100) || fabs(vap_dens) < F(0.000100) || fabs(dens_dry + vap_dens) < F(0.000100))
         su_warn(dg, 42);
      a.avcp = cpd * (dens_dry / (dens_dry + vap_dens)) + cpm * (vap_dens / (dens_dry + vap_dens));
   }
   a.lv_J_kg = Lat_vap(dg, Temp_C, a.Ea_hPa, Press_hPa, a.avcp, lv_sw);
   if (Temp_C < F(0.000)) {
      a.lvS_J_kg = (F(2834.1) - F(0.29) * Temp_C) * F(1e3);   // Lat_vapSublim, meteo.f95:348
   } else {
      a.lvS_J_kg = 2834000;
   }
   if (Press_hPa < 900) su_warn(dg, 46);
}

// ------------------------------------------------------------------ stability functions
#define SU_NEUT_LIMIT F(1.E-7)
#define SU_PIOVER2 ((double)(3.14159274f / 2.f))   // ACOS(-1.)/2. in REAL(4)

// atmmoiststab.f95:252-331.  SM is a template parameter where the caller knows the scheme.
SU_HOT double stab_psi_mom(int sm, double zl)   // by value: the reference's in-place clip of zL (sm = 2) never outlives its callers' own clips
{
   double psim = 0, x, x2;
   if (fabs(zl) < SU_NEUT_LIMIT) {
      psim = 0;
   } else if (zl < -SU_NEUT_LIMIT) {
      if (sm == 3) {
         psim = (double)(0.6f * 2.f) * log((1 + pow(1 - 16 * zl, F(0.5))) / 2);
      } else if (sm == 1) {
         psim = pow(F(1.) - F(16.) * zl, F(0.25)) - 1;
      } else if (sm == 2 || sm == 4 || sm == 7 || sm == 5 || sm == 6) {
         if (sm == 2) x = pow(F(1.) - (F(15.2) * zl), F(0.25));
         else if (sm == 4) x = pow(1 - F(19.3) * zl, F(0.25));
         else if (sm == 7) x = pow(1 - (F(28.) * zl), F(0.25));
         else if (sm == 5) x = (zl >= F(-0.16)) ? 1 + F(1.38) * zl : (double)(0.42f * -1.f) * pow(zl, F(0.333));
         else x = (zl >= F(0.06)) ? 1.0 : pow((-1) * zl / F(0.06), F(0.25));
         x2 = log((1 + (x * x)) / F(2.));
         psim = (F(2.) * log((1 + x) / F(2.))) + x2 - (F(2.) * atan(x)) + SU_PIOVER2;
      }
   } else if (zl > SU_NEUT_LIMIT) {
      if (sm == 3) {
         psim = (-6) * log(1 + zl);
      } else if (sm == 1) {
         psim = (F(-4.8)) * zl;
      } else if (sm == 2) {
         if (zl > F(1000.)) zl = F(1000.);
         psim = (F(-17.) * (F(1.) - exp(F(-0.29) * zl)));
      } else if (sm == 4) {
         psim = (-6) * zl;
      }
   }
   return psim;
}
SU_INL double stab_psi_mom_v(int sm, double zl) { return stab_psi_mom(sm, zl); }

// atmmoiststab.f95:334-394
SU_INL double stab_phi_mom(int sm, double zl)
{
   double phim = 0;
   if (fabs(zl) < SU_NEUT_LIMIT) {
      phim = 0;
   } else if (zl < -SU_NEUT_LIMIT) {
      if (sm == 1 || sm == 3) phim = pow(F(1.) - F(16.) * zl, F(-0.25));
      else if (sm == 2) phim = pow(F(1.) - (F(15.2) * zl), F(-0.25));
      else if (sm == 4) phim = pow(F(1.) - F(19.) * zl, F(-0.25));
      else if (sm == 7) phim = pow(F(1.) - (F(28.) * zl), F(-0.25));
      else if (sm == 5) phim = (zl >= F(-0.16)) ? 1 + F(1.38) * zl : (double)(0.42f * -1.f) * zl * (F(-0.333));
      else if (sm == 6) phim = (zl >= F(0.06)) ? 1.0 : pow((-1) * zl / F(0.06), F(-0.25));
   } else if (zl > SU_NEUT_LIMIT) {
      if (sm == 1 || sm == 2) phim = F(1.) + (F(4.8)) * zl;
      else if (sm == 4) phim = 1 + 6 * zl;
      else if (sm == 3) phim = F(1.) + F(6.) * zl / (F(1.) + zl);
   }
   return phim;
}

// atmmoiststab.f95:397-452
SU_DEV double stab_psi_heat(int sm, double zl)
{
   double psih = 0, x = 0;
   if (fabs(zl) < SU_NEUT_LIMIT) {
      psih = 0;
   } else if (zl < -SU_NEUT_LIMIT) {
      if (sm == 3) {
         psih = (2) * log((1 + pow(1 - 16 * zl, F(0.5))) / 2);
      } else {
         if (sm == 4) x = F(0.95) * pow(F(1.) - F(11.6) * zl, F(0.5));
         else if (sm == 7) x = pow(1 - (F(28.) * zl), F(0.25));
         else if (sm == 2) x = F(0.95) * pow(F(1.) - F(15.2) * zl, F(0.5));
         psih = 2 * log((1 + x) / 2);
      }
   } else if (zl > SU_NEUT_LIMIT) {
      if (zl <= 1) {
         psih = (sm == 4) ? (F(-7.8)) * zl : (F(-4.5)) * zl;
      } else {
         psih = (sm == 4) ? (F(-7.8)) * (1 + log(zl)) : (F(-4.5)) * (1 + log(zl));
      }
   }
   return psih;
}

// atmmoiststab.f95:455-497
SU_INL double stab_phi_heat(int sm, double zl)
{
   double phih = 0;
   if (fabs(zl) < SU_NEUT_LIMIT) {
      phih = 0;
   } else if (zl < -SU_NEUT_LIMIT) {
      if (sm == 3) phih = pow(F(1.) - F(16.) * zl, F(-0.5));
      else if (sm == 4) phih = F(0.95) * pow(F(1.) - F(11.6) * zl, F(-0.5));
      else if (sm == 7) phih = pow(1 - (F(28.) * zl), F(-0.25));
      else if (sm == 2) phih = F(0.95) * pow(F(1.) - F(15.2) * zl, F(-0.5));
   } else if (zl > SU_NEUT_LIMIT) {
      phih = (sm == 4) ? F(1.) + F(7.8) * zl : F(1.) + F(4.5) * zl;
   }
   return phih;
}

// atmmoiststab.f95:104-248: fixed-point iteration for the Obukhov length (<= 329 trips).
// log(zzd/z0m) is step-invariant: evaluated once per step by the caller (lg).
SU_BIG1 void cal_Stab(SuDiag &dg, int sm, double zzd, double z0m, double lg, double avU1, double Temp_C, double QH_init,
                       double avdens, double avcp, double &L_MOD_out, double &TStar_out, double &UStar_out, double &zL_out)
{
   // the iteration runs on register copies; the by-reference results (stack slots of the caller) are written once
   const double k = F(0.4), grav = F(9.80665);
   double H_init = QH_init / (avdens * avcp);
   double G_T_K = (grav / (Temp_C + F(273.16))) * k;
   double KUZ = k * avU1;
   if (zzd < 0) su_fatal(dg, 32);
   // lg = log(zzd / z0m): the same for every trip of the step, evaluated once by the caller
#ifdef SU_NO_LG_HOIST
   lg = log(zzd / z0m);
#endif
   double UStar = KUZ / lg;
   double h = (fabs(H_init) < F(0.001)) ? F(0.001) : H_init;
   double TStar = (-h / UStar);
   double L_MOD = (UStar * UStar) / (G_T_K * TStar);
   double zL = zL_out;
   if (lg < F(0.001000)) su_fatal(dg, 17);
   int i = 1;
   double LOLD = F(-999.), z0L, psim, psimz0;
   while ((fabs(LOLD - L_MOD) > F(0.01)) && (i < 330)) {
      LOLD = L_MOD;
      zL = zzd / L_MOD;
      z0L = z0m / L_MOD;
      psim = stab_psi_mom(sm, zL);
      psimz0 = stab_psi_mom(sm, z0L);
      UStar = KUZ / (lg - psim + psimz0);
      TStar = (-h / UStar);
      L_MOD = (UStar * UStar) / (G_T_K * TStar);
      i = i + 1;
   }
   zL = SU_MIN(F(5.), SU_MAX(F(-5.), zL));
   L_MOD = zzd / zL;
   if (fabs(L_MOD) > F(3e4)) L_MOD = L_MOD / fabs(L_MOD) * F(3e4);
   z0L = z0m / L_MOD;
   psim = stab_psi_mom(sm, zL);
   psimz0 = stab_psi_mom(sm, z0L);
   UStar = KUZ / (lg - psim + psimz0);
   UStar = SU_MAX(F(0.15), UStar);
   TStar = (-h / UStar);
   L_MOD_out = L_MOD; TStar_out = TStar; UStar_out = UStar; zL_out = zL;
}

// ------------------------------------------------------------------ resistances
// resist.f95:514-551
// veg029 = VegFraction**0.29 (a site constant, evaluated once per launch)
SU_DEV double cal_z0V(int RoughLenHeatMethod, double z0m, double VegFraction, double veg029, double UStar)
{
   const double muu = F(1.46e-5);
   double z0V = F(0.01);
   if (RoughLenHeatMethod == 1) {
      z0V = z0m / 10;
   } else if (RoughLenHeatMethod == 2 || (RoughLenHeatMethod == 5 && !(VegFraction > F(.999)))) {
      z0V = z0m * exp(2 - (F(1.2) - F(0.9) * veg029) * pow(UStar * z0m / muu, F(0.25)));
   } else if (RoughLenHeatMethod == 3) {
      z0V = z0m * 0x1.1b4866p-29;   // EXP(-20.) folded in REAL(4) by the reference compiler
   } else if (RoughLenHeatMethod == 4) {
      z0V = z0m * exp(2 - F(1.29) * pow(UStar * z0m / muu, F(0.25)));
   } else if (RoughLenHeatMethod == 5) {
      z0V = z0m / 10;
   }
   return z0V;
}

// resist.f95:5-104.  psim/psih for method 2 are passed in when already known (same zL for RA and RAsnow).
SU_DEV double AerodynamicResistance(SuDiag &dg, double ZZD, double z0m, double lg_zz0, double AVU1, double L_mod, double UStar,
                                    double VegFraction, double veg029, int ARMethod, int sm, int RoughLenHeatMethod)
{
   const double k2 = F(0.16);
   double RA = 0;
   if (ARMethod == 1) {
      double lg = lg_zz0;   // log(ZZD / z0m)
      RA = (lg * lg) / (k2 * AVU1);
   } else if (ARMethod == 2) {
      double psim = stab_psi_mom_v(sm, ZZD / L_mod);
      double psih = stab_psi_heat(sm, ZZD / L_mod);
#ifdef SU_NO_SITE_HOIST
      veg029 = pow(VegFraction, F(0.29));
#endif
#ifdef SU_NO_LG_HOIST
      lg_zz0 = log(ZZD / z0m);
#endif
      double z0V = cal_z0V(RoughLenHeatMethod, z0m, VegFraction, veg029, UStar);
      if (ZZD / L_mod == 0 || UStar == 0) {
         RA = (lg_zz0 * log(ZZD / z0V)) / (k2 * AVU1);
      } else {
         RA = ((lg_zz0 - psim) * (log(ZZD / z0V) - psih)) / (k2 * AVU1);
      }
   } else if (ARMethod == 3) {
      double lg = lg_zz0;
      RA = (F(4.72) * (lg * lg)) / (1 + F(0.54) * AVU1);
   }
   if (RA > 200) {
      su_warn(dg, 7);
      RA = 200;
   } else if (RA < 10) {
      su_warn(dg, 7);
      RA = 10;
   }
   return RA;
}

// resist.f95:108-368 (gsModel 1/3 = Jarvi 2011, 2/4 = Ward 2016)
// (SnowFrac == NULL: no snow anywhere -- the snow-off instances do not build an array of zeros for it)
SU_DEV void SurfaceResistance(SuDiag &dg, const SuewsClassParams &P, const double *sfr, const double *SnowFrac,
                              double avkdn, double Temp_C, double dq, double xsmd, double vsmd,
                              const double *LAI_id, double &gfunc, double &gsc, double &ResistSurf)
{
   const int gsModel = P.gsModel;
   const double G1 = P.G1, G2 = P.G2, G3 = P.G3, G4 = P.G4, G5 = P.G5, G6 = P.G6, TH = P.TH, TL = P.TL, S1 = P.S1,
                S2 = P.S2, Kmax = P.Kmax;
   double gdq = F(0.5), gtemp = F(0.5), gs = F(0.5), gq = F(0.5);
   gsc = F(0.1);
   if (gsModel >= 1 && gsModel <= 4) {
      const bool jarvi = (gsModel == 1 || gsModel == 3);
      double QNM = Kmax / (Kmax + G2);
      if (jarvi) {
         gq = (avkdn / (G2 + avkdn)) / QNM;
         gdq = (dq < G4) ? 1 - G3 * dq : 1 - G3 * G4;
      } else {
         gq = su_div0(su_div0(avkdn, avkdn + G2), QNM);
         if (avkdn >= Kmax) gq = 1;
         gdq = G3 + (1 - G3) * pow(G4, dq);
      }
      double TC = (TH - G5) / (G5 - TL);
      double TC2 = (G5 - TL) * P.gs_pow_TH_G5;   // pow(TH - G5, TC): a class constant, formed at upload
      if (Temp_C <= TL) {
         gtemp = (TL + F(0.1) - TL) * pow(TH - (TL + F(0.1)), TC) / TC2;
         double mn = SnowFrac ? SnowFrac[0] : 0.0;
         if (SnowFrac)
            for (int i = 1; i < 6; i++) mn = SU_MIN(mn, SnowFrac[i]);
         if (mn != 1) su_warn(dg, 29);
      } else if (Temp_C >= TH) {
         gtemp = ((TH - F(0.1)) - TL) * pow(TH - (TH - F(0.1)), TC) / TC2;
         su_warn(dg, 29);
      } else {
         gtemp = (Temp_C - TL) * pow(TH - Temp_C, TC) / TC2;
      }
      double sdp = S1 / G6 + S2;
      double smd_use = (P.SMDMethod > 0) ? xsmd : vsmd;
      if (jarvi) gs = 1 - exp(G6 * (smd_use - sdp));
      else gs = (1 - exp(G6 * (smd_use - sdp))) / (1 - exp(G6 * (-sdp)));
      if (!(P.SMDMethod > 0)) {
         if (sfr[CONIF] + sfr[DECID] + sfr[GRASS] == 0 || sfr[WATER] == 1) gs = 0;
      }
      if (gs < 0) {
         su_warn(dg, 65);
         gs = F(0.0001);
      }
      double gl = 0;
      for (int iv = 0; iv < 3; iv++)
         gl = gl + (sfr[iv + 2] * (1 - (SnowFrac ? SnowFrac[iv + 2] : 0.0))) * LAI_id[iv] / P.LAIMax[iv] * P.MaxConductance[iv];
      if (avkdn <= 0) gsc = F(0.1);
      else gsc = (G1 * gq * gdq * gtemp * gs * gl);
      if (gsc <= 0) {
         su_warn(dg, 65);
         gsc = F(0.1);
      }
   } else {
      su_fatal(dg, 71);
   }
   ResistSurf = 1 / (gsc / 1000);
   gfunc = gdq * gtemp * gs * gq;
}

// resist.f95:372-397
SU_DEV double BoundaryLayerResistance(double zzd, double z0m, double avU1, double &UStar)
{
   if (UStar < F(0.01)) UStar = avU1 / log(zzd / z0m) * F(0.4);
   return (F(1.1) / UStar) + (F(5.6) * pow(UStar, F(0.333333)));
}

// resist.f95:400-512
SU_DEV void cal_RoughnessParameters(SuDiag &dg, int RoughLenMomMethod, const SuewsCellParams &C, double porosity_id,
                                    double &FAI, double &Zh, double &z0m, double &zdm, double &ZZD)
{
   const double *sfr = C.sfr.v;
   double areaZh = (sfr[BLDG] + sfr[CONIF] + sfr[DECID]);
   if (areaZh != 0) {
      Zh = (C.bldgH * sfr[BLDG] + C.EveTreeH * sfr[CONIF] + (C.DecTreeH * (1 - porosity_id)) * sfr[DECID]) / areaZh;
      FAI = C.FAIBldg * sfr[BLDG] + C.FAIEveTree * sfr[CONIF] + (C.FAIDecTree * (1 - porosity_id)) * sfr[DECID];
      FAI = SU_MAX(FAI, F(1e-5));
   } else {
      Zh = 0;
      FAI = F(1e-5);
   }
   z0m = 0;
   zdm = 0;
   if (Zh != 0) {
      if (RoughLenMomMethod == 2) {
         z0m = F(0.1) * Zh;
         zdm = F(0.7) * Zh;
      } else if (RoughLenMomMethod == 3) {
         const double c = (double)(0.5f * 1.0f * 1.2f / (0.4f * 0.4f));   // folded in REAL(4)
         zdm = (1 + pow(F(4.43), -sfr[BLDG]) * (sfr[BLDG] - 1)) * Zh;
         z0m = ((1 - zdm / Zh) * exp(-pow(c * (1 - zdm / Zh) * FAI, F(-0.5)))) * Zh;
      }
   } else {
      if (areaZh != 0) su_warn(dg, 15);
      if (areaZh != 1) {
         z0m = (F(0.003) * sfr[PAV] + F(0.02) * sfr[GRASS] + F(0.002) * sfr[BSOIL] + F(0.0005) * sfr[WATER]) /
               (1 - areaZh);
         zdm = 0;
         su_warn(dg, 15);
      } else {
         z0m = 1;
         zdm = 7;
         su_warn(dg, 15);
      }
   }
   if (RoughLenMomMethod == 1) {
      z0m = C.z0m_in;
      zdm = C.zdm_in;
   }
   ZZD = C.Z - zdm;
   if (z0m < 0 || zdm < 0 || ZZD < 0) su_fatal(dg, 14);
}
