/* ORACLE (test infrastructure only -- see suews_oracle.h).
 * Restates SUEWS-SourceCode/suews_phys_atmmoiststab.f95 and
 * suews_phys_resist.f95. */
#include "suews_oracle.h"
#include <math.h>

static const double neut_limit = F(1.E-7); /* atmmoiststab.f95:3 */
#define FMAX(a, b) ((a) > (b) ? (a) : (b))
#define FMIN(a, b) ((a) < (b) ? (a) : (b))

/* atmmoiststab.f95:9-100 */
void so_cal_AtmMoist(SoCtx *cx, double *Temp_C, double Press_hPa, double avRh, double dectime, double *lv_J_kg,
                     double *lvS_J_kg, double *es_hPa, double *Ea_hPa, double *VPd_hpa, double *VPD_Pa, double *dq,
                     double *dens_dry, double *avcp, double *air_dens)
{
   /* 8.31451/0.028965 etc. are REAL(4) divisions (atmmoiststab.f95:45-46) */
   const double gas_ct_dry = (double)(8.31451f / 0.028965f);
   const double gas_ct_wv = (double)(8.31451f / 0.0180153f);
   const int from = 1;
   double vap_dens;

   *es_hPa = so_sat_vap_press_x(cx, Temp_C, Press_hPa, from, dectime); /* may rewrite *Temp_C (meteo:161-168) */
   *Ea_hPa = avRh / 100 * (*es_hPa);
   *VPd_hpa = *es_hPa - *Ea_hPa;
   *VPD_Pa = (*es_hPa * 100) - (*Ea_hPa * 100);
   *dq = so_spec_hum_def(*VPd_hpa, Press_hPa);
   vap_dens = (*Ea_hPa * 100 / ((*Temp_C + F(273.16)) * gas_ct_wv));
   *dens_dry = ((Press_hPa - *Ea_hPa) * 100) / (gas_ct_dry * (F(273.16) + *Temp_C));
   *air_dens = (Press_hPa * 100) / (gas_ct_dry * (*Temp_C + F(273.16)));
   *avcp = so_spec_heat_beer(cx, *Temp_C, avRh, vap_dens, *dens_dry);
   *lv_J_kg = so_Lat_vap(cx, *Temp_C, *Ea_hPa, Press_hPa, *avcp, dectime);
   if (*Temp_C < F(0.000)) {
      *lvS_J_kg = so_Lat_vapSublim(*Temp_C, *Ea_hPa, Press_hPa, *avcp);
   } else {
      *lvS_J_kg = 2834000;
   }
   if (Press_hPa < 900) so_error_hint(cx, 46, "Function LUMPS_cal_AtmMoist", Press_hPa, -55.55, -55);
}

/* atmmoiststab.f95:252-331.  zL is by-reference: method 2 clamps it to 1000. */
double so_stab_psi_mom(int StabilityMethod, double *zl_p)
{
   double piover2, psim, x, x2, zl = *zl_p;
   psim = 0;
   piover2 = (double)(3.14159274f / 2.f); /* ACOS(-1.)/2. evaluated in REAL(4) */
   if (fabs(zl) < neut_limit) {
      psim = 0;
   } else if (zl < -neut_limit) { /* unstable */
      if (StabilityMethod == 1) {
         psim = pow(F(1.) - F(16.) * zl, F(0.25)) - 1;
      } else if (StabilityMethod == 2) {
         x = pow(F(1.) - (F(15.2) * zl), F(0.25));
         x2 = log((1 + (x * x)) / F(2.));
         psim = (F(2.) * log((1 + x) / F(2.))) + x2 - (F(2.) * atan(x)) + piover2;
      } else if (StabilityMethod == 3) {
         psim = (double)(0.6f * 2.f) * log((1 + pow(1 - 16 * zl, F(0.5))) / 2);
      } else if (StabilityMethod == 4) {
         x = pow(1 - F(19.3) * zl, F(0.25));
         x2 = log((1 + (x * x)) / F(2.));
         psim = (F(2.) * log((1 + x) / F(2.))) + x2 - (F(2.) * atan(x)) + piover2;
      } else if (StabilityMethod == 7) {
         x = pow(1 - (F(28.) * zl), F(0.25));
         x2 = log((1 + x * x) / F(2.));
         psim = (F(2.) * log((1 + x) / F(2.))) + x2 - (F(2.) * atan(x)) + piover2;
      } else if (StabilityMethod == 5) {
         if (zl >= F(-0.16)) {
            x = 1 + F(1.38) * zl;
         } else {
            x = (double)(0.42f * (float)(-1)) * pow(zl, F(0.333));
         }
         x2 = log((1 + (x * x)) / F(2.));
         psim = (F(2.) * log((1 + x) / F(2.))) + x2 - (F(2.) * atan(x)) + piover2;
      } else if (StabilityMethod == 6) {
         if (zl >= F(0.06)) {
            x = 1;
         } else {
            x = pow((-1) * zl / F(0.06), F(0.25));
         }
         x2 = log((1 + (x * x)) / F(2.));
         psim = (F(2.) * log((1 + x) / F(2.))) + x2 - (F(2.) * atan(x)) + piover2;
      }
   } else if (zl > neut_limit) { /* stable */
      if (StabilityMethod == 1) {
         psim = (F(-4.8)) * zl;
      } else if (StabilityMethod == 2) {
         if (zl > F(1000.)) {
            zl = F(1000.);
            *zl_p = zl;
         }
         psim = (F(-17.) * (F(1.) - exp(F(-0.29) * zl)));
      } else if (StabilityMethod == 4) {
         psim = (-6) * zl;
      } else if (StabilityMethod == 3) {
         psim = (-6) * log(1 + zl);
      }
   }
   return psim;
}

/* atmmoiststab.f95:334-394 */
double so_stab_phi_mom(int StabilityMethod, double zl)
{
   double phim = 0;
   if (fabs(zl) < neut_limit) {
      phim = 0;
   } else if (zl < -neut_limit) {
      if (StabilityMethod == 1) {
         phim = pow(F(1.) - F(16.) * zl, F(-0.25));
      } else if (StabilityMethod == 2) {
         phim = pow(F(1.) - (F(15.2) * zl), F(-0.25));
      } else if (StabilityMethod == 3) {
         phim = pow(F(1.) - F(16.) * zl, F(-0.25));
      } else if (StabilityMethod == 4) {
         phim = pow(F(1.) - F(19.) * zl, F(-0.25));
      } else if (StabilityMethod == 7) {
         phim = pow(F(1.) - (F(28.) * zl), F(-0.25));
      } else if (StabilityMethod == 5) {
         if (zl >= F(-0.16)) {
            phim = 1 + F(1.38) * zl;
         } else {
            phim = (double)(0.42f * (float)(-1)) * zl * (F(-0.333));
         }
      } else if (StabilityMethod == 6) {
         if (zl >= F(0.06)) {
            phim = 1;
         } else {
            phim = pow((-1) * zl / F(0.06), F(-0.25));
         }
      }
   } else if (zl > neut_limit) {
      if (StabilityMethod == 1) {
         phim = F(1.) + (F(4.8)) * zl;
      } else if (StabilityMethod == 2) {
         phim = F(1.) + (F(4.8)) * zl;
      } else if (StabilityMethod == 4) {
         phim = 1 + 6 * zl;
      } else if (StabilityMethod == 3) {
         phim = F(1.) + F(6.) * zl / (F(1.) + zl);
      }
   }
   return phim;
}

/* atmmoiststab.f95:397-452 */
double so_stab_psi_heat(int StabilityMethod, double zl)
{
   double psih = 0, x = 0;
   if (fabs(zl) < neut_limit) {
      psih = 0;
   } else if (zl < -neut_limit) {
      if (StabilityMethod == 3) {
         psih = (2) * log((1 + pow(1 - 16 * zl, F(0.5))) / 2);
      } else {
         if (StabilityMethod == 4) {
            x = F(0.95) * pow(F(1.) - F(11.6) * zl, F(0.5));
         } else if (StabilityMethod == 7) {
            x = pow(1 - (F(28.) * zl), F(0.25));
         } else if (StabilityMethod == 2) {
            x = F(0.95) * pow(F(1.) - F(15.2) * zl, F(0.5));
         }
         psih = 2 * log((1 + x) / 2);
      }
   } else if (zl > neut_limit) {
      if (zl <= 1) {
         if (StabilityMethod == 4) {
            psih = (F(-7.8)) * zl;
         } else {
            psih = (F(-4.5)) * zl;
         }
      } else {
         if (StabilityMethod == 4) {
            psih = (F(-7.8)) * (1 + log(zl));
         } else {
            psih = (F(-4.5)) * (1 + log(zl));
         }
      }
   }
   return psih;
}

/* atmmoiststab.f95:455-497 */
double so_stab_phi_heat(int StabilityMethod, double zl)
{
   double phih = 0;
   if (fabs(zl) < neut_limit) {
      phih = 0;
   } else if (zl < -neut_limit) {
      if (StabilityMethod == 3) {
         phih = pow(F(1.) - F(16.) * zl, F(-0.5));
      } else {
         if (StabilityMethod == 4) {
            phih = F(0.95) * pow(F(1.) - F(11.6) * zl, F(-0.5));
         } else if (StabilityMethod == 7) {
            phih = pow(1 - (F(28.) * zl), F(-0.25));
         } else if (StabilityMethod == 2) {
            phih = F(0.95) * pow(F(1.) - F(15.2) * zl, F(-0.5));
         }
      }
   } else if (zl > neut_limit) {
      if (StabilityMethod == 4) {
         phih = F(1.) + F(7.8) * zl;
      } else {
         phih = F(1.) + F(4.5) * zl;
      }
   }
   return phih;
}

/* atmmoiststab.f95:104-248 */
void so_cal_Stab(SoCtx *cx, int StabilityMethod, double zzd, double z0m, double zdm, double avU1, double Temp_C,
                 double QH_init, double avdens, double avcp, double *L_MOD, double *TStar, double *UStar, double *zL)
{
   const double k = F(0.4), grav = F(9.80665);
   double G_T_K, KUZ, LOLD, psim, z0L, psimz0, H_init, h;
   int i;

   H_init = QH_init / (avdens * avcp);
   G_T_K = (grav / (Temp_C + F(273.16))) * k;
   KUZ = k * avU1;
   if (zzd < 0)
      so_error_hint(cx, 32, "Windspeed Ht too low relative to zdm [Stability calc]- values [z-zdm, zdm]", zzd, zdm,
                    -55);
   *UStar = KUZ / log(zzd / z0m);
   if (fabs(H_init) < F(0.001)) {
      h = F(0.001);
   } else {
      h = H_init;
   }
   *TStar = (-h / *UStar);
   *L_MOD = ((*UStar) * (*UStar)) / (G_T_K * (*TStar));
   if (log(zzd / z0m) < F(0.001000)) {
      so_error_hint(cx, 17, "In stability subroutine, (z-zd) < z0.", zzd, z0m, -55);
   }
   i = 1;
   LOLD = F(-999.);
   while ((fabs(LOLD - *L_MOD) > F(0.01)) && (i < 330)) {
      LOLD = *L_MOD;
      *zL = zzd / *L_MOD;
      z0L = z0m / *L_MOD;
      psim = so_stab_psi_mom(StabilityMethod, zL);
      psimz0 = so_stab_psi_mom(StabilityMethod, &z0L);
      *UStar = KUZ / (log(zzd / z0m) - psim + psimz0);
      *TStar = (-h / *UStar);
      *L_MOD = ((*UStar) * (*UStar)) / (G_T_K * (*TStar));
      i = i + 1;
   }
#ifdef SO_OPCOUNT   /* statistics of the counting build: calls, iterations, calls that ran to the cap of 330 */
   if (so_opcount_on) { so_opcount_cur[12] += 1; so_opcount_cur[13] += (unsigned long long)(i - 1); so_opcount_cur[14] += (i >= 330); }
#endif
   *zL = FMIN(F(5.), FMAX(F(-5.), *zL));
   *L_MOD = zzd / *zL;
   if (fabs(*L_MOD) > F(3e4)) *L_MOD = *L_MOD / fabs(*L_MOD) * F(3e4);
   z0L = z0m / *L_MOD;
   psim = so_stab_psi_mom(StabilityMethod, zL);
   psimz0 = so_stab_psi_mom(StabilityMethod, &z0L);
   *UStar = KUZ / (log(zzd / z0m) - psim + psimz0);
   *UStar = FMAX(F(0.15), *UStar);
   *TStar = (-h / *UStar);
}

/* ---------------- suews_phys_resist.f95 ---------------- */

/* resist.f95:514-551 */
double so_cal_z0V(int RoughLenHeatMethod, double z0m, double VegFraction, double UStar)
{
   const double muu = F(1.46e-5);
   double z0V = F(0.01);
   if (RoughLenHeatMethod == 1) {
      z0V = z0m / 10;
   } else if (RoughLenHeatMethod == 2) {
      z0V = z0m * exp(2 - (F(1.2) - F(0.9) * pow(VegFraction, F(0.29))) * pow(UStar * z0m / muu, F(0.25)));
   } else if (RoughLenHeatMethod == 3) {
      z0V = z0m * 0x1.1b4866p-29; /* EXP(-20.) folded in REAL(4) */
   } else if (RoughLenHeatMethod == 4) {
      z0V = z0m * exp(2 - F(1.29) * pow(UStar * z0m / muu, F(0.25)));
   } else if (RoughLenHeatMethod == 5) {
      if (VegFraction > F(.999)) {
         z0V = z0m / 10;
      } else {
         z0V = z0m * exp(2 - (F(1.2) - F(0.9) * pow(VegFraction, F(0.29))) * pow(UStar * z0m / muu, F(0.25)));
      }
   }
   return z0V;
}

/* resist.f95:5-104 (module global psih becomes a local) */
void so_AerodynamicResistance(SoCtx *cx, double ZZD, double z0m, double AVU1, double L_mod, double UStar,
                              double VegFraction, int AerodynamicResistanceMethod, int StabilityMethod,
                              int RoughLenHeatMethod, double *RA)
{
   const double notUsed = F(-55.5), k2 = F(0.16);
   double psim, psih, z0V, lg, t;

   if (AerodynamicResistanceMethod == 1) {
      lg = log(ZZD / z0m);
      *RA = (lg * lg) / (k2 * AVU1);
   } else if (AerodynamicResistanceMethod == 2) {
      t = ZZD / L_mod;
      psim = so_stab_psi_mom(StabilityMethod, &t);
      psih = so_stab_psi_heat(StabilityMethod, ZZD / L_mod);
      cx->last_psih = psih;   /* (a module variable in the reference: USE sues_data, ONLY: psih) */
      z0V = so_cal_z0V(RoughLenHeatMethod, z0m, VegFraction, UStar);
      if (ZZD / L_mod == 0 || UStar == 0) {
         *RA = (log(ZZD / z0m) * log(ZZD / z0V)) / (k2 * AVU1);
      } else {
         *RA = ((log(ZZD / z0m) - psim) * (log(ZZD / z0V) - psih)) / (k2 * AVU1);
      }
   } else if (AerodynamicResistanceMethod == 3) {
      lg = log(ZZD / z0m);
      *RA = (F(4.72) * (lg * lg)) / (1 + F(0.54) * AVU1);
   }

   if (*RA > 200) {
      so_error_hint(cx, 7, "In AerodynamicResistance.f95, calculated RA > 200 s m-1; RA set to 200 s m-1", *RA, notUsed,
                    -55);
      *RA = 200;
   } else if (*RA < 10) {
      so_error_hint(cx, 7, "In AerodynamicResistance.f95, calculated RA < 10 s m-1; RA set to 10 s m-1", *RA, notUsed,
                    -55);
      *RA = 10;
   }
}

/* resist.f95:108-368 */
void so_SurfaceResistance(SoCtx *cx, int id, int it, int SMDMethod, const double SnowFrac[NSURF],
                          const double sfr[NSURF], double avkdn, double Temp_C, double dq, double xsmd, double vsmd,
                          const double MaxConductance[3], const double LAIMax[3], const double LAI_id[3],
                          int gsModel, double Kmax, double G1, double G2, double G3, double G4, double G5, double G6,
                          double TH, double TL, double S1, double S2, double *gfunc, double *gsc, double *ResistSurf)
{
   double gl, QNM, gq, gdq, TC, TC2, gtemp, sdp, gs, id_real, mn;
   int iv, i;

   gdq = F(0.5);
   gtemp = F(0.5);
   gs = F(0.5);
   gq = F(0.5);
   id_real = (double)(float)id;

   if (gsModel == 1 || gsModel == 3) {
      QNM = Kmax / (Kmax + G2);
      gq = (avkdn / (G2 + avkdn)) / QNM;
      if (dq < G4) {
         gdq = 1 - G3 * dq;
      } else {
         gdq = 1 - G3 * G4;
      }
      TC = (TH - G5) / (G5 - TL);
      TC2 = (G5 - TL) * pow(TH - G5, TC);
      if (Temp_C <= TL) {
         gtemp = (TL + F(0.1) - TL) * pow(TH - (TL + F(0.1)), TC) / TC2;
         mn = SnowFrac[0];
         for (i = 1; i < 6; i++) mn = FMIN(mn, SnowFrac[i]);
         if (mn != 1)
            so_error_hint(cx, 29, "subroutine SurfaceResistance.f95: T changed to fit limits TL=0.1,Temp_c,id,it",
                          Temp_C, id_real, it);
      } else if (Temp_C >= TH) {
         gtemp = ((TH - F(0.1)) - TL) * pow(TH - (TH - F(0.1)), TC) / TC2;
         so_error_hint(cx, 29, "subroutine SurfaceResistance.f95: T changed to fit limits TH=39.9,Temp_c,id,it",
                       Temp_C, id_real, it);
      } else {
         gtemp = (Temp_C - TL) * pow(TH - Temp_C, TC) / TC2;
      }
      sdp = S1 / G6 + S2;
      if (SMDMethod > 0) {
         gs = 1 - exp(G6 * (xsmd - sdp));
      } else {
         gs = 1 - exp(G6 * (vsmd - sdp));
         if (sfr[CONIF] + sfr[DECID] + sfr[GRASS] == 0 || sfr[WATER] == 1) gs = 0;
      }
      if (gs < 0) {
         so_error_hint(cx, 65,
                       "subroutine SurfaceResistance.f95 (gsModel=1): g(smd) < 0 calculated, setting to 0.0001", gs,
                       id_real, it);
         gs = F(0.0001);
      }
      gl = 0;
      for (iv = 0; iv < 3; iv++) {
         gl = gl + (sfr[iv + 2] * (1 - SnowFrac[iv + 2])) * LAI_id[iv] / LAIMax[iv] * MaxConductance[iv];
      }
      if (avkdn <= 0) {
         *gsc = F(0.1);
      } else {
         *gsc = (G1 * gq * gdq * gtemp * gs * gl);
      }
      if (*gsc <= 0) {
         so_error_hint(cx, 65, "subroutine SurfaceResistance.f95 (gsModel=1): gs <= 0, setting to 0.1 mm s-1", *gsc,
                       id_real, it);
         *gsc = F(0.1);
      }
   } else if (gsModel == 2 || gsModel == 4) {
      QNM = Kmax / (Kmax + G2);
      gq = (avkdn / (avkdn + G2)) / QNM;
      if (avkdn >= Kmax) {
         gq = 1;
      }
      gdq = G3 + (1 - G3) * pow(G4, dq);
      TC = (TH - G5) / (G5 - TL);
      TC2 = (G5 - TL) * pow(TH - G5, TC);
      if (Temp_C <= TL) {
         gtemp = (TL + F(0.1) - TL) * pow(TH - (TL + F(0.1)), TC) / TC2;
         mn = SnowFrac[0];
         for (i = 1; i < 6; i++) mn = FMIN(mn, SnowFrac[i]);
         if (mn != 1)
            so_error_hint(cx, 29, "subroutine SurfaceResistance.f95: T changed to fit limits TL+0.1,Temp_C,id,it",
                          Temp_C, id_real, it);
      } else if (Temp_C >= TH) {
         gtemp = ((TH - F(0.1)) - TL) * pow(TH - (TH - F(0.1)), TC) / TC2;
         so_error_hint(cx, 29, "subroutine SurfaceResistance.f95: T changed to fit limits TH-0.1,Temp_C,id,it",
                       Temp_C, id_real, it);
      } else {
         gtemp = (Temp_C - TL) * pow(TH - Temp_C, TC) / TC2;
      }
      sdp = S1 / G6 + S2;
      if (SMDMethod > 0) {
         gs = (1 - exp(G6 * (xsmd - sdp))) / (1 - exp(G6 * (-sdp)));
      } else {
         gs = (1 - exp(G6 * (vsmd - sdp))) / (1 - exp(G6 * (-sdp)));
         if (sfr[CONIF] + sfr[DECID] + sfr[GRASS] == 0 || sfr[WATER] == 1) gs = 0;
      }
      if (gs < 0) {
         so_error_hint(cx, 65, "subroutine SurfaceResistance.f95 (gsModel=2): gs < 0 calculated, setting to 0.0001",
                       gs, id_real, it);
         gs = F(0.0001);
      }
      gl = 0;
      for (iv = 0; iv < 3; iv++) {
         gl = gl + (sfr[iv + 2] * (1 - SnowFrac[iv + 2])) * LAI_id[iv] / LAIMax[iv] * MaxConductance[iv];
      }
      if (avkdn <= 0) {
         *gsc = F(0.1);
      } else {
         *gsc = (G1 * gq * gdq * gtemp * gs * gl);
      }
      if (*gsc <= 0) {
         so_error_hint(cx, 65, "subroutine SurfaceResistance.f95 (gsModel=2): gsc <= 0, setting to 0.1 mm s-1", *gsc,
                       id_real, it);
         *gsc = F(0.1);
      }
   } else if (gsModel < 1 || gsModel > 4) {
      so_error_hint(cx, 71, "Value of gsModel not recognised.", -55, -55, gsModel);
   }

   *ResistSurf = 1 / (*gsc / 1000);
   *gfunc = gdq * gtemp * gs * gq;
}

/* resist.f95:372-397 */
void so_BoundaryLayerResistance(double zzd, double z0m, double avU1, double *UStar, double *rb)
{
   const double k = F(0.4);
   if (*UStar < F(0.01)) {
      *UStar = avU1 / log(zzd / z0m) * k;
   }
   *rb = (F(1.1) / *UStar) + (F(5.6) * pow(*UStar, F(0.333333)));
}

/* resist.f95:400-512 */
void so_cal_RoughnessParameters(SoCtx *cx, int RoughLenMomMethod, const double sfr[NSURF], double bldgH,
                                double EveTreeH, double DecTreeH, double porosity_id, double FAIBldg,
                                double FAIEveTree, double FAIDecTree, double z0m_in, double zdm_in, double Z,
                                double *FAI, double *Zh, double *z0m, double *zdm, double *ZZD)
{
   const double notUsed = F(-55.5);
   double areaZh, z0m4Paved, z0m4Grass, z0m4BSoil, z0m4Water;

   areaZh = (sfr[BLDG] + sfr[CONIF] + sfr[DECID]);
   z0m4Paved = F(0.003);
   z0m4Grass = F(0.02);
   z0m4BSoil = F(0.002);
   z0m4Water = F(0.0005);

   if (areaZh != 0) {
      /* DOT_PRODUCT: sequential accumulation from 0 */
      *Zh = (bldgH * sfr[BLDG] + EveTreeH * sfr[CONIF] + (DecTreeH * (1 - porosity_id)) * sfr[DECID]) / areaZh;
      *FAI = FAIBldg * sfr[BLDG] + FAIEveTree * sfr[CONIF] + (FAIDecTree * (1 - porosity_id)) * sfr[DECID];
      *FAI = FMAX(*FAI, F(1e-5));
   } else {
      *Zh = 0;
      *FAI = F(1e-5);
   }

   if (*Zh != 0) {
      if (RoughLenMomMethod == 2) {
         *z0m = F(0.1) * (*Zh);
         *zdm = F(0.7) * (*Zh);
      } else if (RoughLenMomMethod == 3) {
         /* 0.5*1.0*1.2/0.4**2 is folded in REAL(4) */
         const double c = (double)(0.5f * 1.0f * 1.2f / (0.4f * 0.4f));
         *zdm = (1 + pow(F(4.43), -sfr[BLDG]) * (sfr[BLDG] - 1)) * (*Zh);
         *z0m = ((1 - *zdm / *Zh) * exp(-pow(c * (1 - *zdm / *Zh) * (*FAI), F(-0.5)))) * (*Zh);
      }
   } else if (*Zh == 0) {
      if (areaZh != 0)
         so_error_hint(cx, 15, "In SUEWS_RoughnessParameters.f95, zh = 0 m but areaZh > 0", *Zh, areaZh, -55);
      if (areaZh != 1) {
         *z0m = (z0m4Paved * sfr[PAV] + z0m4Grass * sfr[GRASS] + z0m4BSoil * sfr[BSOIL] + z0m4Water * sfr[WATER]) /
                (1 - areaZh);
         *zdm = 0;
         so_error_hint(cx, 15, "Setting z0m and zdm using default values", *z0m, *zdm, -55);
      } else if (areaZh == 1) {
         *z0m = 1;
         *zdm = 7;
         so_error_hint(cx, 15, "Assuming mean height = 10 m, Setting z0m and zdm to default value", *z0m, *zdm, -55);
      }
   }

   if (RoughLenMomMethod == 1) {
      *z0m = z0m_in;
      *zdm = zdm_in;
   }

   *ZZD = Z - *zdm;

   if (*z0m < 0) so_error_hint(cx, 14, "In SUEWS_cal_RoughnessParameters, z0 < 0 m.", *z0m, notUsed, -55);
   if (*zdm < 0) so_error_hint(cx, 14, "In SUEWS_cal_RoughnessParameters, zd < 0 m.", *zdm, notUsed, -55);
   if (*ZZD < 0) so_error_hint(cx, 14, "In SUEWS_cal_RoughnessParameters, (z-zd) < 0 m.", *ZZD, notUsed, -55);
}
