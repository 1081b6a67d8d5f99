// dev_step.cuh -- one grid cell, one timestep: the device form of SUEWS_cal_Main
// (SUEWS-SourceCode/suews_ctrl_driver.f95:45-1045) and of its stage wrappers (:1049-2270).
//
// The reference re-runs everything inside its Tsurf-convergence DO WHILE (driver.f95:642-936), including
// work that cannot change between trips (clock decode, roughness, the whole solar position, DailyState,
// water use, QF, drainage, surface resistance).  Here each of those is evaluated once per step; what
// depends on the iterated surface temperature stays in the loop.  One reference side effect couples the
// two groups: sat_vap_press_x rewrites Temp_C to 0.001 when |Temp_C| < 0.001 (meteo.f95:161-168), so
// trips >= 2 see a different Temp_C than trip 1 for the stages that run before cal_AtmMoist.  That case is
// detected and the affected stages are redone once, which reproduces the reference exactly.
#pragma once
#include "dev_atmos.cuh"
#include "dev_radiation.cuh"
#include "dev_surface.cuh"
#include "dev_snow.cuh"
#include "dev_daily.cuh"
#include "dev_rsl.cuh"

// Optional per-step output lines (reference layouts).  NULL = not wanted.
struct SuLines {
   double *datetime5, *suews82, *snow98, *rsl133, *daily45;
};

// Step-invariant hoisted quantities that survive across the Tsurf loop
struct SuHoist {
   double Tair_av_next;
   SuDaily d;
   double wu_nsurf[NSURF], wu_int, wu_ext;
   SuAnthro an;
   double SoilMoistCap, vsmd;
   double drain[NSURF], drain_per_tstep, AddWater[NSURF], AddWaterRunoff[NSURF];
   double gsc, ResistSurf, gfunc;
   double SnowAlb_q;   // snow albedo after ageing (update_snow_albedo)
};

SU_DEV double cal_tair_av(double tair_av_prev, int dt_since_start, int tstep, double temp_c)
{   // driver.f95:3530-3556
   const double len_day_s = 24 * 3600;
   double len_cal_s = (dt_since_start > len_day_s) ? len_day_s : (double)(dt_since_start + tstep);
   double temp_k = temp_c + F(273.15);
   return tair_av_prev * (len_cal_s - (double)((float)tstep * 1.f)) / len_cal_s + temp_k * tstep / len_cal_s;
}

// Advances S by one step.  R: clock + forcing of this step; E: solar ephemeris of this step.
// Returns the number of Tsurf iterations used.
__device__ int suews_step(SuDiag &dg, const SuewsClassParams &P, const SuewsCellParams &C, const SuewsStepRec &R,
                          const SuewsEphem &E, int dt_since_start, int tstep_prev, SuewsCellState &S, double *acc82,
                          const SuLines &L)
{
   const double *sfr = C.sfr.v;
   const int sm = P.StabilityMethod;
   const int snowUse = P.snowUse;
   // ---------------- step invariants (driver.f95:645-691)
   const double dectime = R.dectime;
   const int nsh = 3600 / P.tstep;
   const double nsh_real = (double)((float)nsh * 1.0f);        // util_time.f95:295-297
   const double tstep_real = (double)((float)P.tstep * 1.0f);
   const double VegFraction = sfr[CONIF] + sfr[DECID] + sfr[GRASS];
   const double ImpervFraction = sfr[PAV] + sfr[BLDG];
   const double PervFraction = 1 - ImpervFraction;
   const double NonWaterFraction = 1 - sfr[WATER];
   const int dow = R.dow;
   const int DLS = (R.id > P.startDLS && R.id < P.endDLS) ? 1 : 0;
   double FAI, zH, z0m, zdm, ZZD;
   cal_RoughnessParameters(dg, P.RoughLenMomMethod, C, S.porosity_id, FAI, zH, z0m, zdm, ZZD);
   double azimuth, zenith_deg;
   spa_location_part(E, C.lat, C.lng, C.alt, azimuth, zenith_deg);
   int nrm_use, AlbedoChoice, ldown_option;
   RadMethod(dg, P.NetRadiationMethod, snowUse, nrm_use, AlbedoChoice, ldown_option);

   // forcing (with the optional per-cell offsets of the synthetic configurations)
   double Temp_C = R.Temp_C + C.dTemp_C;
   double avRh = R.avRh + C.dRH;
   if (C.dRH != 0) avRh = SU_MIN(100.0, SU_MAX(F(0.01), avRh));
   const double avkdn = R.avkdn, avU1 = R.avU1, Precip = R.Precip, Press_hPa = R.Press_hPa;

   int lv_sw = (int)S.lv_switch;
   double alb1 = S.alb1_hidden;

   // *_next values produced inside the loop
   double qn1_av_next = S.qn1_av, dqndt_next = S.dqndt, qn1_s_av_next = S.qn1_s_av, dqnsdt_next = S.dqnsdt;
   double SnowfallCum_next = S.SnowfallCum, SnowAlb_next = S.SnowAlb;
   double alb_next[NSURF];
   SuHoist H;
   SuAtm A;
   SuRad Rd;
   SuSnowHeat Hs;
   SuLumps Lm;
   SuQEWork W;
   Lm.sIce_hpa = 0; Lm.QH_LUMPS = 0; Lm.QE_LUMPS = 0; Lm.VegPhenLumps = 0; Lm.psyc_hPa = 0; Lm.s_hPa = 0;
   Rd.fcld = 0; Rd.ldown = 0; Rd.alb1 = alb1;
   double deltaQi[NSURF], smd_nsurf[NSURF], soilstoreOld[NSURF];
   double SnowFrac_q[NSURF], SnowPack_m[NSURF], SnowDens_m[NSURF], SnowWater_in[NSURF];
   for (int is = 0; is < NSURF; is++) { deltaQi[is] = 0; smd_nsurf[is] = 0; SnowWater_in[is] = S.SnowWater[is]; }
   double a1 = 0, a2 = 0, a3 = 0, qs = 0, qn1 = 0, qf = 0, qh = 0, qe = 0, qh_resist = 0, RA = 0, RAsnow = 0, rb = 0;
   double UStar = 0, TStar = 0, L_mod = 0, zL = 0, TSfc_C = 0, smd = 0, tot_chang_per_tstep = 0;
   double runoffSoil_per_tstep = 0, state_per_tstep = 0, NWstate_per_tstep = 0, ev_per_tstep = 0, runoff_per_tstep = 0,
          surf_chang_per_tstep = 0, chSnow_per_interval = 0;
   const bool temp_will_rewrite = fabs(Temp_C) < F(0.001000);

   bool converged = false;
   double Ts_iter = Temp_C;
   int i_iter = 1;
   while (!converged && i_iter < 100) {
      if (i_iter == 1 || (i_iter == 2 && temp_will_rewrite)) {
         // ---- stages the reference runs before cal_AtmMoist (driver.f95:670-716)
         H.Tair_av_next = cal_tair_av(S.Tair_av, dt_since_start, P.tstep, Temp_C);
         for (int i = 0; i < 12; i++) H.d.HDD_id[i] = S.HDD_id[i];
         for (int i = 0; i < 3; i++) { H.d.GDD_id[i] = S.GDD_id[i]; H.d.SDD_id[i] = S.SDD_id[i]; H.d.LAI_id[i] = S.LAI_id[i]; }
         for (int i = 0; i < 9; i++) H.d.WUDay_id[i] = S.WUDay_id[i];
         H.d.Tmin_id = S.Tmin_id; H.d.Tmax_id = S.Tmax_id; H.d.lenDay_id = S.lenDay_id;
         H.d.DecidCap_id = S.DecidCap_id; H.d.albDecTr_id = S.albDecTr_id; H.d.albEveTr_id = S.albEveTr_id;
         H.d.albGrass_id = S.albGrass_id; H.d.porosity_id = S.porosity_id;
         H.d.DecidStoreCap = SU_SDP(S.StoreDrainPrm.v, 5, DECID);
         H.d.deltaLAI = 0;
         cal_DailyState(dg, P, C.lat, R.id, R.it, R.imin, 0, dow, tstep_prev, dt_since_start, nsh_real, avkdn, Temp_C,
                        Precip, R.LAI_obs, S.state_id.v, S.soilstore_id.v, H.d);
      }

      cal_AtmMoist(dg, Temp_C, Press_hPa, avRh, lv_sw, A);   // may rewrite Temp_C on trip 1

      if (i_iter == 1 || (i_iter == 2 && temp_will_rewrite)) {
         // ---- Tsurf-independent stages that follow cal_AtmMoist (driver.f95:727-755, :827-834, part of :838)
         double SoilState;
         update_SoilMoist(NonWaterFraction, P.SoilStoreCap.v, sfr, S.soilstore_id.v, H.SoilMoistCap, SoilState, H.vsmd, smd);
         cal_WaterUse(P, C, nsh_real, R.wu_m3, dow, H.d.HDD_id, H.d.WUDay_id, nsh, R.it, R.imin, DLS, H.wu_nsurf,
                      H.wu_int, H.wu_ext);
         cal_AnthropogenicEmission(dg, P, C, dow, DLS, H.d.HDD_id, R.it, R.imin, R.qf_obs, Temp_C, H.an);
         qf = H.an.qf;
         H.SnowAlb_q = update_snow_albedo(P.tstep, S.SnowPack.v, S.SnowAlb, Temp_C, P.tau_a, P.tau_f, P.SnowAlbMax, P.SnowAlbMin);
         // SUEWS_cal_Water driver.f95:1614-1694 (StoreDrainPrm_next differs from _prev in (6,Decid) only)
         if (NonWaterFraction != 0) {
            double s = 0;
            for (int is = 0; is < NSURF - 1; is++) {
               double cap = (is == DECID) ? H.d.DecidStoreCap : SU_SDP(S.StoreDrainPrm.v, 5, is);
               H.drain[is] = drainage(dg, S.state_id[is], cap, SU_SDP(S.StoreDrainPrm.v, 1, is),
                                      SU_SDP(S.StoreDrainPrm.v, 2, is), SU_SDP(S.StoreDrainPrm.v, 3, is), nsh_real);
               s += H.drain[is] * sfr[is];
            }
            H.drain_per_tstep = s / NonWaterFraction;
         } else {
            for (int is = 0; is < NSURF - 1; is++) H.drain[is] = 0;
            H.drain_per_tstep = 0;
         }
         H.drain[WATER] = 0;
         ReDistributeWater(snowUse, P.WaterDist.v, sfr, H.drain, H.AddWaterRunoff, H.AddWater);
      }

      // ---------------- SUEWS_cal_Qn driver.f95:1288-1455
      for (int is = 0; is < NSURF; is++) { alb_next[is] = S.alb[is]; SnowFrac_q[is] = S.SnowFrac[is]; }
      double tsurf, kup, lup, ldown, fcld, qn1_snowfree, qn1_S;
      if (nrm_use > 0) {
         if (snowUse == 0)
            for (int is = 0; is < NSURF; is++) SnowFrac_q[is] = 0;
         Rd.ldown = (ldown_option == 1) ? R.ldown_obs : -9.0;
         if (ldown_option == 2) Rd.fcld = R.fcld_obs;
         alb_next[DECID] = H.d.albDecTr_id;
         alb_next[CONIF] = H.d.albEveTr_id;
         alb_next[GRASS] = H.d.albGrass_id;
         NARP(sfr, SnowFrac_q, alb_next, P.emis.v, S.IceFrac.v, P.NARP_EMIS_SNOW, zenith_deg, Ts_iter, avkdn, Temp_C, avRh,
              A.Ea_hPa, R.qn1_obs, H.SnowAlb_q, AlbedoChoice, ldown_option, nrm_use, Rd);
         qn1 = Rd.qn1; qn1_snowfree = Rd.qn1_snowfree; qn1_S = Rd.qn1_S; kup = Rd.kup; lup = Rd.lup; ldown = Rd.ldown;
         tsurf = Rd.tsurf; fcld = Rd.fcld;
      } else {
         for (int is = 0; is < NSURF; is++) { SnowFrac_q[is] = R.snowFrac_obs; Rd.Tsurf_ind[is] = -999; }
         qn1 = R.qn1_obs; qn1_snowfree = R.qn1_obs; qn1_S = R.qn1_obs;
         ldown = -999; lup = -999; kup = -999; tsurf = -999; fcld = -999;
      }
      if (ldown_option == 1) fcld = -999;
      SnowAlb_next = H.SnowAlb_q;

      // ---------------- SUEWS_cal_Qs driver.f95:1457-1612
      {
         qs = -999; a1 = -999; a2 = -999; a3 = -999;
         double qn1_use = (P.OHMIncQF == 1) ? qf + qn1 : qn1;
         if (P.StorageHeatMethod == 0) {
            qs = R.qs_obs;
         } else if (P.StorageHeatMethod == 1) {
            OHM(dg, P, sfr, qn1_use, S.qn1_av, S.dqndt, qn1_av_next, dqndt_next, qn1_S, S.qn1_s_av, S.dqnsdt, qn1_s_av_next,
                dqnsdt_next, dt_since_start, H.d.HDD_id[9], S.soilstore_id.v, S.state_id.v, SnowFrac_q, a1, a2, a3, qs,
                deltaQi);
         } else {
            su_fatal(dg, 1003);   // AnOHM / ESTM need state outside the cal_Main argument list
         }
      }

      // ---------------- Snow_cal_MeltHeat driver.f95:793-805 (SnowFrac / SnowAlb alias in and out)
      Snow_cal_MeltHeat(P, sfr, A.lvS_J_kg, A.lv_J_kg, tstep_real, Temp_C, Precip, nsh_real, Rd.Tsurf_ind, S.state_id.v,
                        Rd.qn1_ind_snow, S.SnowWater.v, deltaQi, S.SnowPack.v, SnowFrac_q, SnowAlb_next, S.SnowDens.v,
                        S.SnowfallCum, SnowPack_m, SnowFrac_q, SnowAlb_next, SnowDens_m, SnowfallCum_next, Hs);

      // ---------------- LUMPS first guess on trip 1 (driver.f95:809-824)
      double QH_Init;
      if (i_iter == 1) {
         LUMPS_cal_QHQE(dg, P, sfr, qn1, qf, qs, Hs.Qm, Temp_C, Hs.veg_fr, A.avcp, Press_hPa, A.lv_J_kg, H.d.LAI_id, Lm);
         QH_Init = Lm.QH_LUMPS;
      } else {
         QH_Init = qh;
      }

      // ---------------- SUEWS_cal_Resistance driver.f95:2134-2270
      cal_Stab(dg, sm, ZZD, z0m, avU1, Temp_C, QH_Init, A.avdens, A.avcp, L_mod, TStar, UStar, zL);
      RA = AerodynamicResistance(dg, ZZD, z0m, avU1, L_mod, UStar, VegFraction, P.AerodynamicResistanceMethod, sm,
                                 P.RoughLenHeatMethod);
      if (snowUse == 1)
         RAsnow = AerodynamicResistance(dg, ZZD, z0m, avU1, L_mod, UStar, VegFraction, P.AerodynamicResistanceMethod, sm, 3);
      if (i_iter == 1 || (i_iter == 2 && temp_will_rewrite))
         SurfaceResistance(dg, P, sfr, SnowFrac_q, avkdn, Temp_C, A.dq, R.xsmd, H.vsmd, H.d.LAI_id, H.gfunc, H.gsc, H.ResistSurf);
      rb = BoundaryLayerResistance(ZZD, z0m, avU1, UStar);

      // ---------------- SUEWS_cal_QE driver.f95:1744-2082
      {
         for (int is = 0; is < NSURF; is++) {
            W.state_id[is] = S.state_id[is];
            W.soilstore_id[is] = S.soilstore_id[is];
            soilstoreOld[is] = S.soilstore_id[is];
            W.SnowPack[is] = SnowPack_m[is];
            W.SnowFrac[is] = SnowFrac_q[is];
            W.SnowWater[is] = S.SnowWater[is];
            W.iceFrac[is] = S.IceFrac[is];
            W.SnowDens[is] = SnowDens_m[is];
            W.SurplusEvap[is] = 0;
            W.runoffSoil[is] = 0;
         }
         W.SnowRemoval[0] = 0; W.SnowRemoval[1] = 0;
         W.fl.runoffAGimpervious = 0; W.fl.surplusWaterBody = 0; W.fl.runoffAGveg = 0; W.fl.runoffPipes = 0;
         W.fl.runoffWaterBody = 0;
         W.swe = 0; W.mwstore = 0;
         const double tstep_r = P.tstep * 1.0;
         const double nsh_r = 3600 / tstep_r;        // driver.f95:1930-1931: recomputed in double here
         const double tlv = A.lv_J_kg / tstep_r;
         const double pin = SU_MAX(F(0.), Precip);
         const double qn_e = qn1_snowfree + qf - qs;
         double qe_per_tstep = 0, ev = 0, qe_surf = 0;
         ev_per_tstep = 0; surf_chang_per_tstep = 0; runoff_per_tstep = 0; state_per_tstep = 0; NWstate_per_tstep = 0;
         chSnow_per_interval = 0;
         for (int is = 0; is < NSURF; is++) {
            if (snowUse == 1 && Hs.snowCalcSwitch[is] == 1) {
               if (sfr[is] != 0) {
                  double chSnow_tot, ev_tot, qe_tot, runoff_tot, surf_chang_tot;
                  SnowCalc(dg, P, C, is, R.imin, R.it, dow, nsh_r, A.lvS_J_kg, A.avdens, avRh, Press_hPa, Temp_C, RAsnow,
                           Lm.psyc_hPa, A.avcp, Lm.sIce_hpa, PervFraction, VegFraction, A.VPd_hpa, qn_e, Lm.s_hPa,
                           H.ResistSurf, RA, rb, tlv, Precip, H.drain, S.state_id.v, Hs, Rd.Tsurf_ind, S.StoreDrainPrm.v,
                           (is == DECID) ? H.d.DecidStoreCap : SU_SDP(S.StoreDrainPrm.v, 5, is), H.AddWater, H.AddWaterRunoff, W, chSnow_tot, ev_tot, qe_tot, runoff_tot, surf_chang_tot);
                  ev_per_tstep = ev_per_tstep + ev_tot;
                  qe_per_tstep = qe_per_tstep + qe_tot;
                  runoff_per_tstep = runoff_per_tstep + runoff_tot;
                  surf_chang_per_tstep = surf_chang_per_tstep + surf_chang_tot;
                  chSnow_per_interval = chSnow_per_interval + chSnow_tot;
               } else {
                  W.SnowFrac[is] = 0;
                  W.SnowDens[is] = 0;
                  W.SnowPack[is] = 0;
               }
            } else {
               double cap = (is == DECID) ? H.d.DecidStoreCap : SU_SDP(S.StoreDrainPrm.v, 5, is);
               cal_evap(P.EvapMethod, W.state_id[is], P.WetThresh[is], cap, A.VPd_hpa, A.avdens, A.avcp, qn_e, Lm.s_hPa,
                        Lm.psyc_hPa, H.ResistSurf, RA, rb, tlv, ev, qe_surf);
               double runoff_is, state_is = W.state_id[is];
               cal_water_storage(dg, is, P, C, pin, H.wu_nsurf[is], H.drain[is], H.AddWater[is], nsh_r, S.state_id[is],
                                 H.AddWaterRunoff[is], PervFraction, W.fl, ev, W.soilstore_id, W.SurplusEvap, runoff_is,
                                 state_is);
               W.state_id[is] = state_is;
               ev_per_tstep = ev_per_tstep + ev * sfr[is];
               qe_per_tstep = qe_per_tstep + qe_surf * sfr[is];
               surf_chang_per_tstep = surf_chang_per_tstep + (state_is - S.state_id[is]) * sfr[is];
               runoff_per_tstep = runoff_per_tstep + runoff_is * sfr[is];
               state_per_tstep = state_per_tstep + state_is * sfr[is];
               if (NonWaterFraction != 0 && is != WATER)
                  NWstate_per_tstep = NWstate_per_tstep + (state_is * sfr[is] / NonWaterFraction);
            }
         }
         qe = qe_per_tstep;
      }

      // ---------------- SUEWS_cal_QH driver.f95:2084-2131 (QHMethod = 1)
      qh = (qn1 + qf + Hs.QmRain) - (qe + qs + Hs.Qm + Hs.QmFreez);
      qh_resist = (RA != 0) ? A.avdens * A.avcp * (tsurf - Temp_C) / RA : -999.0;

      cal_HorizontalSoilWater(P, sfr, C.SurfaceArea, NonWaterFraction, tstep_real, W.soilstore_id, W.runoffSoil,
                              runoffSoil_per_tstep);
      cal_SoilState(dg, P.SMDMethod, R.xsmd, NonWaterFraction, H.SoilMoistCap, P.SoilStoreCap.v, surf_chang_per_tstep,
                    W.soilstore_id, soilstoreOld, sfr, smd, smd_nsurf, tot_chang_per_tstep);

      TSfc_C = qh / (A.avdens * A.avcp) * RA + Temp_C;   // cal_tsfc driver.f95:3558
      converged = !(fabs(Ts_iter - TSfc_C) > F(0.1));
      Ts_iter = TSfc_C;
      i_iter = i_iter + 1;
      (void)kup; (void)lup; (void)ldown; (void)fcld;
      if (dg.fatal) break;
   }

   // ---------------- diagnostics (driver.f95:943-958)
   SuRSL rs;
   RSLProfile(sm, zH, z0m, zdm, L_mod, sfr, FAI, A.avcp, A.lv_J_kg, A.avdens, avU1, Temp_C, avRh, Press_hPa, C.Z, qh, qe,
              rs, L.rsl133);
   // SUEWS_cal_BiogenCO2: Fc terms are zeroed for EmissionsMethod 0-6 (driver.f95:1276-1280); CO2_biogen
   // (EmissionsMethod >= 11) is not part of this kernel yet.
   if (P.EmissionsMethod >= 11) su_fatal(dg, 1004);
   const double Fc = H.an.Fc_anthro + 0.0, Fc_photo = 0, Fc_respi = 0;

   // ---------------- commit (driver.f95:965-994)
   S.qn1_av = qn1_av_next; S.dqndt = dqndt_next; S.qn1_s_av = qn1_s_av_next; S.dqnsdt = dqnsdt_next;
   S.SnowfallCum = SnowfallCum_next; S.SnowAlb = SnowAlb_next;
   for (int is = 0; is < NSURF; is++) {
      S.IceFrac[is] = W.iceFrac[is]; S.SnowWater[is] = W.SnowWater[is]; S.SnowDens[is] = W.SnowDens[is];
      S.SnowFrac[is] = W.SnowFrac[is]; S.SnowPack[is] = W.SnowPack[is]; S.soilstore_id[is] = W.soilstore_id[is];
      S.state_id[is] = W.state_id[is]; S.alb[is] = alb_next[is];
   }
   for (int i = 0; i < 3; i++) { S.GDD_id[i] = H.d.GDD_id[i]; S.SDD_id[i] = H.d.SDD_id[i]; S.LAI_id[i] = H.d.LAI_id[i]; }
   S.DecidCap_id = H.d.DecidCap_id; S.albDecTr_id = H.d.albDecTr_id; S.albEveTr_id = H.d.albEveTr_id;
   S.albGrass_id = H.d.albGrass_id; S.porosity_id = H.d.porosity_id;
   SU_SDP(S.StoreDrainPrm.v, 5, DECID) = H.d.DecidStoreCap;
   S.Tair_av = H.Tair_av_next; S.Tmin_id = H.d.Tmin_id; S.Tmax_id = H.d.Tmax_id; S.lenDay_id = H.d.lenDay_id;
   for (int i = 0; i < 12; i++) S.HDD_id[i] = H.d.HDD_id[i];
   for (int i = 0; i < 9; i++) S.WUDay_id[i] = H.d.WUDay_id[i];
   S.lv_switch = (double)lv_sw;
   S.alb1_hidden = Rd.alb1;

   // ---------------- SUEWS_update_outputLine driver.f95:2273-2457
   if (acc82 || L.suews82) {
      double o[SUEWS_NCOL_SUEWS];
      int k = 0;
      const double LAI_wt = S.LAI_id[0] * sfr[2] + S.LAI_id[1] * sfr[3] + S.LAI_id[2] * sfr[4];
      double bulkalbedo = 0;
      for (int is = 0; is < NSURF; is++) bulkalbedo += S.alb[is] * sfr[is];
      o[k++] = avkdn; o[k++] = Rd.kup; o[k++] = Rd.ldown; o[k++] = Rd.lup; o[k++] = Rd.tsurf;
      o[k++] = qn1; o[k++] = qf; o[k++] = qs; o[k++] = qh; o[k++] = qe;
      o[k++] = Lm.QH_LUMPS; o[k++] = Lm.QE_LUMPS; o[k++] = qh_resist;
      o[k++] = Precip; o[k++] = H.wu_ext; o[k++] = ev_per_tstep; o[k++] = runoff_per_tstep; o[k++] = tot_chang_per_tstep;
      o[k++] = surf_chang_per_tstep; o[k++] = state_per_tstep; o[k++] = NWstate_per_tstep; o[k++] = H.drain_per_tstep;
      o[k++] = smd;
      o[k++] = C.FlowChange / nsh_real; o[k++] = 0.0;   // AdditionalWater = addPipes + ... = 0 (driver.f95:1669)
      o[k++] = runoffSoil_per_tstep; o[k++] = W.fl.runoffPipes; o[k++] = W.fl.runoffAGimpervious; o[k++] = W.fl.runoffAGveg;
      o[k++] = W.fl.runoffWaterBody;
      o[k++] = H.wu_int; o[k++] = H.wu_nsurf[3]; o[k++] = H.wu_nsurf[2]; o[k++] = H.wu_nsurf[4];   // EveTr/DecTr swap, driver.f95:2421-2439
      for (int is = 0; is < NSURF - 1; is++) o[k++] = (sfr[is] < F(0.00001)) ? -999.0 : smd_nsurf[is];
      for (int is = 0; is < NSURF; is++) o[k++] = (sfr[is] < F(0.00001)) ? -999.0 : S.state_id[is];
      o[k++] = zenith_deg; o[k++] = azimuth; o[k++] = bulkalbedo;
      o[k++] = (nrm_use > 0) ? ((ldown_option == 1) ? -999.0 : Rd.fcld) : -999.0;
      o[k++] = LAI_wt; o[k++] = z0m; o[k++] = zdm;
      o[k++] = UStar; o[k++] = L_mod; o[k++] = RA; o[k++] = H.ResistSurf;
      o[k++] = Fc; o[k++] = Fc_photo; o[k++] = Fc_respi; o[k++] = H.an.Fc_metab; o[k++] = H.an.Fc_traff;
      o[k++] = H.an.Fc_build; o[k++] = H.an.Fc_point;
      o[k++] = (nrm_use > 0) ? Rd.qn1_snowfree : R.qn1_obs; o[k++] = (nrm_use > 0) ? Rd.qn1_S : R.qn1_obs; o[k++] = S.SnowAlb;
      o[k++] = Hs.Qm; o[k++] = Hs.QmFreez; o[k++] = Hs.QmRain; o[k++] = W.swe; o[k++] = Hs.mwh; o[k++] = W.mwstore;
      o[k++] = chSnow_per_interval; o[k++] = W.SnowRemoval[0]; o[k++] = W.SnowRemoval[1];
      o[k++] = TSfc_C; o[k++] = rs.T2_C; o[k++] = rs.q2_gkg; o[k++] = rs.U10_ms; o[k++] = rs.RH2 * F(100.0);
      if (nrm_use <= 0) { o[1] = -999; o[2] = -999; o[3] = -999; o[4] = -999; }
      if (acc82)
         for (int c = 0; c < SUEWS_NCOL_SUEWS; c++) acc82[c] += o[c];
      if (L.suews82)
         for (int c = 0; c < SUEWS_NCOL_SUEWS; c++) L.suews82[c] = o[c];
   }
   if (L.datetime5) {
      L.datetime5[0] = (double)R.iy; L.datetime5[1] = (double)R.id; L.datetime5[2] = (double)R.it;
      L.datetime5[3] = (double)R.imin; L.datetime5[4] = dectime;
   }
   if (L.snow98) {   // dataOutLineSnow, snow.f95:183-190 (values as they leave Snow_cal_MeltHeat)
      double *o = L.snow98;
      int k = 0;
      for (int is = 0; is < NSURF; is++) o[k++] = SnowPack_m[is];
      for (int is = 0; is < NSURF; is++) o[k++] = Hs.mw_ind[is];
      for (int is = 0; is < NSURF; is++) o[k++] = Hs.Qm_melt[is];
      for (int is = 0; is < NSURF; is++) o[k++] = Hs.Qm_rain[is];
      for (int is = 0; is < NSURF; is++) o[k++] = Hs.Qm_freezState[is];
      for (int is = 0; is < NSURF - 1; is++) o[k++] = SnowFrac_q[is];
      for (int is = 0; is < NSURF; is++) o[k++] = Hs.rainOnSnow[is];
      for (int is = 0; is < NSURF; is++) o[k++] = Rd.qn1_ind_snow[is];
      for (int is = 0; is < NSURF; is++) o[k++] = Rd.kup_ind_snow[is];
      for (int is = 0; is < NSURF; is++) o[k++] = Hs.FreezMelt[is];
      for (int is = 0; is < NSURF; is++) o[k++] = SnowWater_in[is];
      for (int is = 0; is < NSURF; is++) o[k++] = SnowDens_m[is];
      for (int is = 0; is < NSURF; is++) o[k++] = Hs.SnowDepth[is];
      for (int is = 0; is < NSURF; is++) o[k++] = Rd.Tsurf_ind_snow[is];
      o[k++] = Rd.alb1;
   }
   if (L.daily45) {   // update_DailyStateLine, dailystate.f95:1044-1113 (HDD_id is DIMENSION(6) in the packer)
      double *o = L.daily45;
      for (int i = 0; i < SUEWS_NCOL_DAILY; i++) o[i] = -999;
      if (R.it == 23 && R.imin == (nsh_real - 1) / nsh_real * 60) {
         int k = 0;
         for (int i = 0; i < 6; i++) o[k++] = S.HDD_id[i];
         for (int i = 0; i < 3; i++) o[k++] = S.GDD_id[i];
         for (int i = 0; i < 3; i++) o[k++] = S.SDD_id[i];
         o[k++] = S.Tmin_id; o[k++] = S.Tmax_id; o[k++] = S.lenDay_id;
         for (int i = 0; i < 3; i++) o[k++] = S.LAI_id[i];
         o[k++] = S.DecidCap_id; o[k++] = S.porosity_id; o[k++] = S.albEveTr_id; o[k++] = S.albDecTr_id; o[k++] = S.albGrass_id;
         for (int i = 0; i < 9; i++) o[k++] = S.WUDay_id[i];
         o[k++] = H.d.deltaLAI; o[k++] = Lm.VegPhenLumps; o[k++] = S.SnowAlb;
         for (int is = 0; is < NSURF; is++) o[k++] = S.SnowDens[is];
         o[k++] = a1; o[k++] = a2; o[k++] = a3;
      }
   }
   return i_iter - 1;
}
