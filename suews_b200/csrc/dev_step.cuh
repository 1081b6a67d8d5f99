// dev_step.cuh -- one grid cell, one timestep: the device form of SUEWS_cal_Main
// (SUEWS-SourceCode/suews_ctrl_driver.f95:45-1045) and of its stage wrappers (:1049-2270).
//
// What is organised differently from the reference (results are the same):
//
// 1. Hoisting.  The reference re-runs everything inside its Tsurf-convergence DO WHILE (driver.f95:642-936),
//    including work that cannot change between trips: clock decode, roughness, the whole solar position,
//    DailyState, water use, QF, drainage, surface resistance, and -- for NetRadiationMethod 1-3/100-300,
//    where Tsurf comes from NARP itself -- radiation, OHM and the snow melt heat.  Each is evaluated once.
// 2. Deferred state update.  Every trip restarts from the *_prev state, so the per-surface water balance,
//    the horizontal soil water flow and the soil-state diagnostics of all trips but the last are discarded
//    by the reference.  A trip only needs QE (7 x cal_evap when no surface is in the snow branch); the
//    state update runs once, after convergence, with the last trip's resistances.
// 3. Lat_vap memo.  Inside one step the inputs of the wet-bulb search (meteo.f95:277) are fixed; only its two
//    SAVE'd switches differ from trip to trip.  There are 4 switch states, so at most 4 searches per step.
// 4. One reference side effect couples trips: sat_vap_press_x rewrites Temp_C to 0.001 when |Temp_C| < 0.001
//    (meteo.f95:161-168), so trips >= 2 see another Temp_C than trip 1 in the stages ahead of cal_AtmMoist.
//    That case is detected and those stages are redone once.
// 5. Phase synchronisation.  A cell-step is ~25 k SASS instructions; warps wandering through it on their own
//    made instruction fetch the top stall (69 % of samples, profiles/r1_ncu_notes.md).  With SU_PHASE_SYNC
//    the warps of a block walk the phases together (block-uniform Tsurf loop, barrier between phases).
#pragma once
#include "dev_atmos.cuh"
#include "dev_radiation.cuh"
#include "dev_surface.cuh"
#include "dev_snow.cuh"
#include "dev_daily.cuh"
#include "dev_rsl.cuh"
#include "dev_solweig.cuh"
#include "dev_estm.cuh"
#include "dev_anohm.cuh"
#include "dev_bluews.cuh"

#ifndef SU_SKIP
#define SU_SKIP 0
#endif
// the three per-surface loops of the step are kept rolled (code size); SU_UNROLL_SURF lets the compiler unroll them
#ifdef SU_UNROLL_SURF
#define SU_UNROLL1
#else
#define SU_UNROLL1 _Pragma("unroll 1")
#endif

#if defined(SU_PHASE_SYNC) && defined(SU_WARP_LOOP)
// the surface-temperature loop is warp-uniform: a warp leaves it as soon as its own 32 cells have converged and waits
// at the barrier behind the loop; the phases inside the loop carry no block barrier (they are one ~40 KB code region)
#define SU_SYNC() __syncthreads()
#define SU_SYNC_OR(p) __any_sync(0xffffffffu, (p))
#define SU_SYNC_LOOP() ((void)0)
#elif defined(SU_PHASE_SYNC)
#define SU_SYNC() __syncthreads()
#define SU_SYNC_OR(p) __syncthreads_or(p)
#define SU_SYNC_LOOP() __syncthreads()
#else
#define SU_SYNC_LOOP() ((void)0)
#define SU_SYNC() ((void)0)
#define SU_SYNC_OR(p) (p)
#endif
// phase barrier k (0 P0|P1, 1 L1|L2, 2 L2|L3, 3 loop|F1, 4 F1|F2, 5 F2|F3); SU_NOSYNC_MASK drops some (tuning)
#ifndef SU_NOSYNC_MASK
#define SU_NOSYNC_MASK 0
#endif
#define SU_SYNCN(k) do { if (!((SU_NOSYNC_MASK >> (k)) & 1)) SU_SYNC(); } while (0)
#define SU_SYNCL(k) do { if (!((SU_NOSYNC_MASK >> (k)) & 1)) SU_SYNC_LOOP(); } while (0)

// Profiling build only (-DSU_PHASE_CLOCKS): SM cycles per phase, as seen by thread 0 of each block after the
// phase's barrier (so a phase's time is that of its slowest warp).  tools/prof_run.py prints them.
#ifdef SU_PHASE_CLOCKS
static __device__ unsigned long long su_phase_clk[16];
#define SU_CLK_DECL long long clk_ = clock64(); const bool clkb_ = true
#define SU_CLK(i)                                                              \
   do {                                                                        \
      if (threadIdx.x == 0) {                                                  \
         long long t_ = clock64();                                             \
         atomicAdd(&su_phase_clk[i], (unsigned long long)(t_ - clk_));         \
         clk_ = t_;                                                            \
      }                                                                        \
   } while (0)
#define SU_CNT(i, n)                                                           \
   do {                                                                        \
      if (threadIdx.x == 0) atomicAdd(&su_phase_clk[i], (unsigned long long)(n)); \
   } while (0)
// sub-stage clocks (slots 10-15) with their own barrier: only legal when every thread of the block is live, which the
// profiling runs arrange (tools/prof_run.py uses whole blocks); their time is taken out of the enclosing phase's slot
#define SU_CLKB(i) do { if (clkb_) { __syncthreads(); SU_CLK(i); } } while (0)
#define SU_CLK_PARAM , long long &clk_, bool clkb_
#define SU_CLK_PASS , clk_, true
#define SU_CLK_PASS_NOBAR , clk_, false
#define SU_CLK_PASS_INNER , clk_, clkb_
#else
#define SU_CLK_DECL
#define SU_CLK(i)
#define SU_CNT(i, n)
#define SU_CLKB(i)
#define SU_CLK_PARAM
#define SU_CLK_PASS
#define SU_CLK_PASS_NOBAR
#define SU_CLK_PASS_INNER
#endif

// Optional per-step output lines (reference layouts).  NULL = not wanted.
struct SuLines {
   double *datetime5, *suews82, *snow98, *rsl133, *daily45, *sol26, *estm27, *bl17;
};

// Tsurf-independent quantities of a step
struct SuHoist {
   double Tair_av_next;
   double wu_nsurf[NSURF], wu_int, wu_ext;
   SuAnthro an;
   double SoilMoistCap, vsmd, smd0;
   double drain[NSURF], drain_per_tstep, AddWater[NSURF], AddWaterRunoff[NSURF];
   double gsc, ResistSurf, gfunc;
   double SnowAlb_q;   // snow albedo after ageing (update_snow_albedo)
};

// Totals of SUEWS_cal_QE (driver.f95:1744-2082)
struct SuQETot {
   double qe, ev_per_tstep, runoff_per_tstep, surf_chang_per_tstep, state_per_tstep, NWstate_per_tstep,
      chSnow_per_interval;
};

SU_INL double cal_tair_av(double tair_av_prev, int dt_since_start, int tstep, double temp_c)
{   // driver.f95:3530-3556
   const double len_day_s = 24 * 3600;
   double len_cal_s = (dt_since_start > len_day_s) ? len_day_s : (double)(dt_since_start + tstep);
   double temp_k = temp_c + F(273.15);
   return tair_av_prev * (len_cal_s - (double)((float)tstep * 1.f)) / len_cal_s + temp_k * tstep / len_cal_s;
}

// The stages ahead of cal_AtmMoist (driver.f95:670-716): 24 h mean air temperature and DailyState.
// D is the cell's daily state, resident across the steps of a launch: it enters holding the values committed by the
// previous step and leaves holding this step's *_next values (cal_DailyState works in place).  Within a day only the
// running sums HDD_id(1:6), Tmin, Tmax and the day length move; everything else is rewritten on the last step of the
// day (dailystate.f95:355-493), so nothing has to be copied from step to step.
SU_DEV void stage_pre_atm(SuDiag &dg, const SuewsClassParams &P, const SuewsCellParams &C, const SuewsStepRec &R,
                          const SuewsCellState &S, int dt_since_start, int tstep_prev, double nsh_real, double Temp_C,
                          SuHoist &H, SuDaily &D)
{
   H.Tair_av_next = cal_tair_av(S.Tair_av, dt_since_start, P.tstep, Temp_C);
   D.deltaLAI = 0;
   cal_DailyState(dg, P, C.lat, R.id, R.it, R.imin, 0, R.dow, tstep_prev, dt_since_start, nsh_real, R.avkdn, Temp_C,
                  R.Precip, R.LAI_obs, S.state_id.v, S.soilstore_id.v, D);
}

// The water stages of the hoisted block -- irrigation (SUEWS_cal_WaterUse), drainage and its redistribution
// (SUEWS_cal_Water) -- read the *_prev state only and feed the state update alone (plus output columns).  With snow off
// nothing inside the surface-temperature loop reads them, so the step runs them right before the update (phase F1):
// their 28 per-surface values are then produced and consumed inside one phase instead of sitting in the frame across
// the whole loop.  (With snow on, a trip's evaporation surplus reads the drainage: they stay ahead of the loop.)
SU_DEV void stage_water_a(const SuewsClassParams &P, const SuewsCellParams &C, const SuewsStepRec &R, int nsh, double nsh_real,
                          int DLS, SuHoist &H, const SuDaily &D)
{
#if !(SU_SKIP & 2)
   cal_WaterUse(P, C, nsh_real, R.wu_m3, R.dow, D.HDD_id, D.WUDay_id, nsh, R.it, R.imin, DLS, H.wu_nsurf, H.wu_int,
                H.wu_ext);
#else
   for (int is = 0; is < NSURF; is++) H.wu_nsurf[is] = 0;
   H.wu_int = 0; H.wu_ext = 0;
#endif
}
SU_DEV void stage_water_b(SuDiag &dg, const SuewsClassParams &P, int snowUse, const SuewsCellParams &C, const SuewsCellState &S,
                          double nsh_real, double NonWaterFraction, SuHoist &H, const SuDaily &D SU_CLK_PARAM)
{
   const double *sfr = C.sfr.v;
   // SUEWS_cal_Water driver.f95:1614-1694 (StoreDrainPrm_next differs from _prev in (6,Decid) only)
   if (NonWaterFraction != 0 && !(SU_SKIP & 8)) {
      double s = 0;
SU_UNROLL1
      for (int is = 0; is < NSURF - 1; is++) {
         double cap = (is == DECID) ? D.DecidStoreCap : SU_SDP(S.StoreDrainPrm.v, 5, is);
         H.drain[is] = drainage(dg, S.state_id[is], cap, SU_SDP(S.StoreDrainPrm.v, 1, is), SU_SDP(S.StoreDrainPrm.v, 2, is),
                                SU_SDP(S.StoreDrainPrm.v, 3, is), nsh_real);
         s += H.drain[is] * sfr[is];
      }
      H.drain_per_tstep = s / NonWaterFraction;
   } else {
      for (int is = 0; is < NSURF - 1; is++) H.drain[is] = 0;
      H.drain_per_tstep = 0;
   }
   H.drain[WATER] = 0;
   SU_CLKB(12);
#if !(SU_SKIP & 16)
   ReDistributeWater(snowUse, P.WaterDist.v, sfr, H.drain, H.AddWaterRunoff, H.AddWater);
#else
   for (int is = 0; is < NSURF; is++) { H.AddWater[is] = 0; H.AddWaterRunoff[is] = 0.5; }
#endif
}

// Tsurf-independent stages that follow cal_AtmMoist (driver.f95:727-755, :827-834 and SurfaceResistance of :838)
SU_DEV void stage_post_atm(SuDiag &dg, const SuewsClassParams &P, int snowUse, const SuewsCellParams &C, const SuewsStepRec &R,
                           const SuewsCellState &S, const SuAtm &A, int DLS, int nsh, double nsh_real, double Temp_C,
                           double NonWaterFraction, const double *SnowFrac_res, SuHoist &H, const SuDaily &D, bool water_late SU_CLK_PARAM)
{
   const double *sfr = C.sfr.v;
   double SoilState;
   // (SU_SKIP: timing-only switches that drop one stage each -- the results are then wrong by construction; they price a
   // stage with its memory traffic included, tools/variants.sh)
#if !(SU_SKIP & 1)
   update_SoilMoist(NonWaterFraction, P.SoilStoreCap.v, sfr, S.soilstore_id.v, H.SoilMoistCap, SoilState, H.vsmd, H.smd0);
#else
   H.SoilMoistCap = 100; SoilState = 50; H.vsmd = 10; H.smd0 = 50;
#endif
   if (!water_late) stage_water_a(P, C, R, nsh, nsh_real, DLS, H, D);
#if !(SU_SKIP & 4)
   cal_AnthropogenicEmission(dg, P, C, R.dow, DLS, D.HDD_id, R.it, R.imin, R.qf_obs, Temp_C, H.an);
#else
   H.an.qf = 20; H.an.Fc_anthro = 0; H.an.Fc_metab = 0; H.an.Fc_traff = 0; H.an.Fc_build = 0; H.an.Fc_point = 0;
#endif
   H.SnowAlb_q = update_snow_albedo(P.tstep, S.SnowPack.v, S.SnowAlb, Temp_C, P.tau_a, P.tau_f, P.SnowAlbMax, P.SnowAlbMin);
   SU_CLKB(11);
   if (!water_late) stage_water_b(dg, P, snowUse, C, S, nsh_real, NonWaterFraction, H, D SU_CLK_PASS_INNER);
   SU_CLKB(13);
#if !(SU_SKIP & 32)
   SurfaceResistance(dg, P, sfr, SnowFrac_res, R.avkdn, Temp_C, A.dq, R.xsmd, H.vsmd, D.LAI_id, H.gfunc, H.gsc, H.ResistSurf);   // (SnowFrac_res: NULL in the snow-off instances)
#else
   H.gfunc = 1; H.gsc = 0.01; H.ResistSurf = 100;
#endif
   SU_CLKB(14);
}

// SUEWS_cal_QE driver.f95:1744-2082: per-surface evaporation + surface/soil/snow state update.
SU_DEV void cal_QE_full(SuDiag &dg, const SuewsClassParams &P, int snowUse, const SuewsCellParams &C, const SuewsStepRec &R,
                        SuewsCellState &S, const SuAtm &A, const SuLumps &Lm, const SuHoist &H, const SuSnowHeat &Hs,
                        const double *Tsurf_ind, const double *SnowPack_m, const double *SnowFrac_q,
                        const double *SnowDens_m, double Temp_C, double avRh, double RA, double RAsnow, double rb,
                        double qn_e, double PervFraction, double VegFraction, double NonWaterFraction, bool snow_arrays,
                        SuQEWork &W, SuQETot &T, const SuDaily &D)
{
   const double *sfr = C.sfr.v;
   W.state_id = S.state_id.v;          // updated in place (SuQEWork)
   W.soilstore_id = S.soilstore_id.v;
   for (int is = 0; is < NSURF; is++) {
      if (snow_arrays) {
         W.SnowPack[is] = SnowPack_m[is];
         W.SnowFrac[is] = SnowFrac_q[is];
         W.SnowWater[is] = S.SnowWater[is];
         W.iceFrac[is] = S.IceFrac[is];
         W.SnowDens[is] = SnowDens_m[is];
      }
      W.SurplusEvap[is] = 0;
      W.runoffSoil[is] = 0;
   }
   W.SnowRemoval[0] = 0; W.SnowRemoval[1] = 0;
   W.fl.runoffAGimpervious = 0; W.fl.surplusWaterBody = 0; W.fl.runoffAGveg = 0; W.fl.runoffPipes = 0;
   W.fl.runoffWaterBody = 0;
   W.swe = 0; W.mwstore = 0;
   const double tstep_r = P.tstep * 1.0;
   const double nsh_r = 3600 / tstep_r;   // driver.f95:1930-1931: recomputed in double here
   const double tlv = A.lv_J_kg / tstep_r;
   const double pin = SU_MAX(F(0.), R.Precip);
   double qe_per_tstep = 0, ev = 0, qe_surf = 0;
   SuEvapTrip et;
   evap_trip_init(et, P.EvapMethod, A.VPd_hpa, A.avdens, A.avcp, qn_e, Lm.s_hPa, Lm.psyc_hPa, H.ResistSurf, RA, rb, tlv);
   SuSnowIce ice;
   ice.have = false; ice.vdrc_RA = 0; ice.tlv_sub = 0;
   SuQEIn win;   // SnowCalc reads the entry values of its surface from the work arrays it commits to
   win.state_id = W.state_id; win.soilstore_id = W.soilstore_id; win.SnowPack = W.SnowPack; win.SnowFrac = W.SnowFrac;
   win.SnowWater = W.SnowWater; win.SnowDens = W.SnowDens;
   T.ev_per_tstep = 0; T.surf_chang_per_tstep = 0; T.runoff_per_tstep = 0; T.state_per_tstep = 0; T.NWstate_per_tstep = 0;
   T.chSnow_per_interval = 0;
SU_UNROLL1
   for (int is = 0; is < NSURF; is++) {
      const double cap = (is == DECID) ? D.DecidStoreCap : SU_SDP(S.StoreDrainPrm.v, 5, is);
      if (snowUse == 1 && Hs.snowCalcSwitch[is] == 1) {
         if (sfr[is] != 0) {
            double chSnow_tot, ev_tot, qe_tot, runoff_tot, surf_chang_tot;
            SnowCalc<true>(dg, P, C, is, R.imin, R.it, R.dow, nsh_r, A.lvS_J_kg, A.avdens, avRh, R.Press_hPa, Temp_C, RAsnow,
                           Lm.psyc_hPa, A.avcp, Lm.sIce_hpa, PervFraction, VegFraction, A.VPd_hpa, qn_e, Lm.s_hPa, H.ResistSurf,
                           RA, rb, tlv, R.Precip, H.drain, S.state_id.v, Hs, Tsurf_ind, S.StoreDrainPrm.v, cap, H.AddWater,
                           H.AddWaterRunoff, win, W.SurplusEvap, et, ice, W, chSnow_tot, ev_tot, qe_tot, runoff_tot, surf_chang_tot);
            T.ev_per_tstep = T.ev_per_tstep + ev_tot;
            qe_per_tstep = qe_per_tstep + qe_tot;
            T.runoff_per_tstep = T.runoff_per_tstep + runoff_tot;
            T.surf_chang_per_tstep = T.surf_chang_per_tstep + surf_chang_tot;
            T.chSnow_per_interval = T.chSnow_per_interval + chSnow_tot;
         } else {
            W.SnowFrac[is] = 0;
            W.SnowDens[is] = 0;
            W.SnowPack[is] = 0;
         }
      } else {
         const double state_old = S.state_id[is];
         evap_surface<true>(et, state_old, P.WetThresh[is], cap, ev, qe_surf);
         double runoff_is, state_is = state_old;
         cal_water_storage(dg, is, P, C, pin, H.wu_nsurf[is], H.drain[is], H.AddWater[is], nsh_r, state_old,
                           H.AddWaterRunoff[is], PervFraction, W.fl, ev, W.soilstore_id, W.SurplusEvap, runoff_is, state_is);
         W.state_id[is] = state_is;
         T.ev_per_tstep = T.ev_per_tstep + ev * sfr[is];
         qe_per_tstep = qe_per_tstep + qe_surf * sfr[is];
         T.surf_chang_per_tstep = T.surf_chang_per_tstep + (state_is - state_old) * sfr[is];
         T.runoff_per_tstep = T.runoff_per_tstep + runoff_is * sfr[is];
         T.state_per_tstep = T.state_per_tstep + state_is * sfr[is];
         if (NonWaterFraction != 0 && is != WATER)
            T.NWstate_per_tstep = T.NWstate_per_tstep + (state_is * sfr[is] / NonWaterFraction);
      }
   }
   T.qe = qe_per_tstep;
}

// QE of a trip when no surface takes the snow branch: the sum the full routine forms, without the state update
SU_DEV double cal_QE_light(const SuewsClassParams &P, const SuewsCellParams &C, const SuewsCellState &S, const SuAtm &A,
                           const SuLumps &Lm, const SuHoist &H, double RA, double rb, double qn_e, const SuDaily &D)
{
   const double tlv = A.lv_J_kg / (P.tstep * 1.0);
   double qe = 0, ev = 0, qe_surf = 0;
   SuEvapTrip et;
   evap_trip_init(et, P.EvapMethod, A.VPd_hpa, A.avdens, A.avcp, qn_e, Lm.s_hPa, Lm.psyc_hPa, H.ResistSurf, RA, rb, tlv);
   const bool need_cap = (P.EvapMethod == 1);   // the store capacity only enters the Rutter form (evap.f95:108-117)
SU_UNROLL1
   for (int is = 0; is < NSURF; is++) {
      const double cap = !need_cap ? 0.0 : ((is == DECID) ? D.DecidStoreCap : SU_SDP(S.StoreDrainPrm.v, 5, is));
      evap_surface<false>(et, S.state_id[is], P.WetThresh[is], cap, ev, qe_surf);
      qe = qe + qe_surf * C.sfr[is];
   }
   return qe;
}

// QE of a trip when some surface takes the snow branch (snowUse = 1): the sum cal_QE_full forms, by the same statements
// -- SnowCalc<false> for the surfaces in the snow branch, cal_evap for the others, and the evaporation surplus of the two
// impervious surfaces that the vegetated ones pick up (waterdist.f95:181-187 / snow.f95:585) -- without the state update.
SU_DEV double cal_QE_light_snow(SuDiag &dg, const SuewsClassParams &P, const SuewsCellParams &C, const SuewsStepRec &R,
                                const SuewsCellState &S, const SuAtm &A, const SuLumps &Lm, const SuHoist &H, const SuSnowHeat &Hs,
                                const double *Tsurf_ind, const double *SnowPack_m, const double *SnowFrac_q, const double *SnowDens_m,
                                double Temp_C, double avRh, double RA, double RAsnow, double rb, double qn_e, double PervFraction,
                                double VegFraction, SuQEWork &Wdummy, const SuDaily &D)
{
   const double *sfr = C.sfr.v;
   const double tstep_r = P.tstep * 1.0;
   const double nsh_r = 3600 / tstep_r;
   const double tlv = A.lv_J_kg / tstep_r;
   const double pin = SU_MAX(F(0.), R.Precip);
   double qe_per_tstep = 0, ev = 0, qe_surf = 0;
   double SurplusEvap[NSURF];
   for (int is = 0; is < NSURF; is++) SurplusEvap[is] = 0;
   SuEvapTrip et;
   evap_trip_init(et, P.EvapMethod, A.VPd_hpa, A.avdens, A.avcp, qn_e, Lm.s_hPa, Lm.psyc_hPa, H.ResistSurf, RA, rb, tlv);
   SuSnowIce ice;
   ice.have = false; ice.vdrc_RA = 0; ice.tlv_sub = 0;
   SuQEIn in;   // entry values of a trip: the *_prev state as Snow_cal_MeltHeat left it
   in.state_id = S.state_id.v; in.soilstore_id = S.soilstore_id.v; in.SnowPack = SnowPack_m; in.SnowFrac = SnowFrac_q;
   in.SnowWater = S.SnowWater.v; in.SnowDens = SnowDens_m;
SU_UNROLL1
   for (int is = 0; is < NSURF; is++) {
      const double cap = (is == DECID) ? D.DecidStoreCap : SU_SDP(S.StoreDrainPrm.v, 5, is);
      if (Hs.snowCalcSwitch[is] == 1) {
         if (sfr[is] != 0) {
            double chSnow_tot, ev_tot, qe_tot, runoff_tot, surf_chang_tot;
            SnowCalc<false>(dg, P, C, is, R.imin, R.it, R.dow, nsh_r, A.lvS_J_kg, A.avdens, avRh, R.Press_hPa, Temp_C, RAsnow,
                            Lm.psyc_hPa, A.avcp, Lm.sIce_hpa, PervFraction, VegFraction, A.VPd_hpa, qn_e, Lm.s_hPa, H.ResistSurf,
                            RA, rb, tlv, R.Precip, H.drain, S.state_id.v, Hs, Tsurf_ind, S.StoreDrainPrm.v, cap, H.AddWater,
                            H.AddWaterRunoff, in, SurplusEvap, et, ice, Wdummy, chSnow_tot, ev_tot, qe_tot, runoff_tot, surf_chang_tot);
            qe_per_tstep = qe_per_tstep + qe_tot;
         }
      } else {
         evap_surface<true>(et, S.state_id[is], P.WetThresh[is], cap, ev, qe_surf);
         if (is == PAV || is == BLDG) {   // cal_water_storage's impervious arm, as far as SurplusEvap (waterdist.f95:139-166)
            const double IPT = 10 / nsh_r;
            double p_mm = pin + H.wu_nsurf[is];
            p_mm = p_mm + H.AddWater[is];
            double chang = p_mm - (H.drain[is] + ev);
            if (p_mm > IPT) chang = IPT - (H.drain[is] + ev);
            const double state_is = S.state_id[is] + chang;
            if (state_is < F(0.0)) SurplusEvap[is] = fabs(state_is);
         }
         qe_per_tstep = qe_per_tstep + qe_surf * sfr[is];
      }
   }
   return qe_per_tstep;
}

SU_COLD void su_copy_daily(SuDaily &dst, const SuDaily &src) { dst = src; }

// Site constants of a launch (pure functions of the cell parameters, formed once by the kernel prologue)
struct SuSite {
   SuSpaSite spa;    // observer terms of the solar position
   double veg029;    // VegFraction**0.29 of cal_z0V (resist.f95:530)
};
SU_DEV void su_site_consts(const SuewsCellParams &C, SuSite &site)
{
   spa_site_part(C.lat, C.alt, site.spa);
   site.veg029 = pow(C.sfr[CONIF] + C.sfr[DECID] + C.sfr[GRASS], F(0.29));
}

// Advances S by one step.  R: clock + forcing of this step; E: solar ephemeris of this step.
// `live` is false for padding threads and for cells stopped by a fatal error; they only keep the barriers.
// Returns the number of Tsurf iterations used.
// K / SNOW / ESTM: the kernel variant (suews_kernel.cuh) -- output kind, and the bucket's snowUse / ESTM switches as
// compile-time constants (-1: read from the class record).
template <int K, int SNOWK, int ESTMK>
static __device__ int suews_step(bool live, SuDiag &dg, const SuewsClassParams &P, const SuewsCellParams &C, const SuSite &site,
                          const SuewsStepRec &R, const SuewsEphem &E, int dt_since_start, int tstep_prev,
                          SuewsCellState &S, SuDaily &D, double *acc82_in, double *acc7, const SuLines &L_in, SuewsEstmState *ES = nullptr,
                          const double *ts9 = nullptr, unsigned long long last_lo = 0, unsigned long long last_hi = 0,
                          bool flux_only = false, const SuAnohmIn *AN = nullptr)
{
   // what the variant rules out is ruled out for the compiler as well: no accumulator of the 82-column line and no
   // output line under policy P0, no per-step lines (bar the DailyState row of P1) when only accumulating
   double *const acc82 = (K == 2) ? nullptr : acc82_in;
   SuLines L = L_in;
   if (K != 0) { L.datetime5 = nullptr; L.suews82 = nullptr; L.snow98 = nullptr; L.rsl133 = nullptr; L.sol26 = nullptr; L.estm27 = nullptr; L.bl17 = nullptr; }
   if (K == 2) L.daily45 = nullptr;
   const double *sfr = C.sfr.v;
   const int sm = P.StabilityMethod;
   const int snowUse = (SNOWK < 0) ? (int)P.snowUse : SNOWK;
   // ---------------- step invariants (driver.f95:645-691)
   const int nsh = 3600 / P.tstep;
   const double nsh_real = (double)((float)nsh * 1.0f);   // util_time.f95:295-297
   const double tstep_real = (double)((float)P.tstep * 1.0f);
   const double VegFraction = sfr[CONIF] + sfr[DECID] + sfr[GRASS];
   const double PervFraction = 1 - (sfr[PAV] + sfr[BLDG]);
   const double NonWaterFraction = 1 - sfr[WATER];
   const int DLS = (R.id > P.startDLS && R.id < P.endDLS) ? 1 : 0;
   double FAI = 0, zH = 0, z0m = 0, zdm = 0, ZZD = 0, lg_zz0 = 0, azimuth = 0, zenith_deg = 0;
   int nrm_use = 0, AlbedoChoice = 0, ldown_option = 0;

   double Temp_C = R.Temp_C + C.dTemp_C;   // forcing with the optional per-cell offsets of the synthetic configs
   double avRh = R.avRh + C.dRH;
   if (C.dRH != 0) avRh = SU_MIN(100.0, SU_MAX(F(0.01), avRh));
   const double avkdn = R.avkdn, avU1 = R.avU1, Precip = R.Precip, Press_hPa = R.Press_hPa;
   const bool temp_will_rewrite = fabs(Temp_C) < F(0.001000);
   const double Temp_C_in = Temp_C;

   int lv_sw = (int)S.lv_switch;
   double qn1_av_next = S.qn1_av, dqndt_next = S.dqndt, qn1_s_av_next = S.qn1_s_av, dqnsdt_next = S.dqnsdt;
   double SnowfallCum_next = S.SnowfallCum, SnowAlb_next = S.SnowAlb;
   SuHoist H;
   SuDaily D_prev;   // only filled when Temp_C is about to be rewritten (rare)
   SuAtm A;
   SuRad Rd;
   SuSnowHeat Hs;
   SuLumps Lm;
   SuQEWork W;
   SuQETot QT;
   Lm.sIce_hpa = 0; Lm.QH_LUMPS = 0; Lm.QE_LUMPS = 0; Lm.VegPhenLumps = 0; Lm.psyc_hPa = 0; Lm.s_hPa = 0;
   Rd.fcld = 0; Rd.ldown = 0; Rd.alb1 = S.alb1_hidden; Rd.kup = 0; Rd.lup = 0; Rd.tsurf = 0;
   double deltaQi[NSURF], smd_nsurf[NSURF];
   double SnowFrac_q[NSURF], SnowPack_m[NSURF], SnowDens_m[NSURF], SnowWater_in[NSURF], SnowFrac_res[NSURF];
   double a1 = 0, a2 = 0, a3 = 0, qs = 0, qn1 = 0, qn1_snowfree = 0, qn1_S = 0, qf = 0, qh = 0, qe = 0, qh_resist = 0;
   double RA = 0, RAsnow = 0, rb = 0, UStar = 0, TStar = 0, L_mod = 0, zL = 0, TSfc_C = 0, smd = 0;
   double tot_chang_per_tstep = 0, runoffSoil_per_tstep = 0;
   // Lat_vap memo: entry switch state -> (exit state, lv)
   double lv_memo[4];
   int lv_exit[4] = {-1, -1, -1, -1};
   bool rad_invariant = false, rad_done = false, qe_full_done = false;
   // with snow off (and no snow output line wanted) the snow work arrays are neither formed nor committed
   const bool use_estm = (ESTMK != 0) && (P.StorageHeatMethod == 4 || P.StorageHeatMethod == 14);
   const bool want_snow_line = (L.snow98 != nullptr);
   const bool snow_arrays = (snowUse == 1) || want_snow_line;
#ifdef SU_WATER_EARLY
   const bool water_late = false;
#else
   const bool water_late = (SNOWK == 0);   // compile-time: the snow-off instances (see stage_water_a)
#endif
   double lv_at_melt = 0;

   SU_CLK_DECL;
   // ================= phase P0: geometry, sun, daily state =================
   if (live) {
      for (int is = 0; is < NSURF; is++) { deltaQi[is] = 0; smd_nsurf[is] = 0; SnowWater_in[is] = S.SnowWater[is]; }
      cal_RoughnessParameters(dg, P.RoughLenMomMethod, C, D.porosity_id, FAI, zH, z0m, zdm, ZZD);
      lg_zz0 = log(ZZD / z0m);   // cal_Stab and AerodynamicResistance of every trip (atmmoiststab.f95:223, resist.f95:60-84)
      RadMethod(dg, P.NetRadiationMethod, snowUse, nrm_use, AlbedoChoice, ldown_option);
      // Sun position of the cell.  The zenith angle enters the energy balance only through NARP's zenith-dependent albedo
      // (AlbedoChoice = 1, NetRadiationMethod 100-300; narp.f95:296-300) and ESTM; otherwise it is an output column.  Under
      // policy P0 -- no output line, seven flux sums -- it is formed only where one of those reads it.
      if (K != 2 || AlbedoChoice == 1 || use_estm) spa_location_part(E, site.spa, C.lng, azimuth, zenith_deg);
      // Tsurf enters NARP only through the iteration-based methods (narp.f95:337-346).  (ESTM, which advances its
      // element temperatures on every call of SUEWS_cal_Qs, is run per trip below whatever the radiation scheme.)
      rad_invariant = (nrm_use < 10);
      // snow fraction seen by SurfaceResistance = SnowFrac_next after Snow_cal_MeltHeat (driver.f95:848)
      for (int is = 0; is < NSURF; is++)
         if (SNOWK != 0) SnowFrac_res[is] = (snowUse == 1) ? ((nrm_use > 0) ? S.SnowFrac[is] : R.snowFrac_obs) : 0.0;
      // the reference runs the stages ahead of cal_AtmMoist again on trips >= 2 when sat_vap_press_x rewrites Temp_C
      // (|Temp_C| < 0.001, see phase L1): only then is a copy of the daily state kept to restart from
      if (temp_will_rewrite) su_copy_daily(D_prev, D);   // (out of line: a rare path whose 45 loads and stores must not be scheduled ahead of the test)
#if !(SU_SKIP & 512)
      stage_pre_atm(dg, P, C, R, S, dt_since_start, tstep_prev, nsh_real, Temp_C, H, D);
#else
      H.Tair_av_next = S.Tair_av;
#endif
   }
   SU_SYNCN(0);
   SU_CLK(0);
   // ================= phase P1: moist air + Tsurf-independent stages =================
   if (live) {
      int sw_in = lv_sw;
      cal_AtmMoist(dg, Temp_C, Press_hPa, avRh, lv_sw, A);   // may rewrite Temp_C
      lv_memo[sw_in] = A.lv_J_kg;
      lv_exit[sw_in] = lv_sw;
      SU_CLKB(10);
      stage_post_atm(dg, P, snowUse, C, R, S, A, DLS, nsh, nsh_real, Temp_C, NonWaterFraction, (SNOWK == 0) ? nullptr : SnowFrac_res, H, D, water_late SU_CLK_PASS);
      qf = H.an.qf;
   }

   bool done = !live || dg.fatal != 0;
   double Ts_iter = Temp_C_in;   // driver.f95:640: taken before cal_AtmMoist can rewrite Temp_C
   // state after the previous two trips, for the exact cycle detection in phase L3
   double cyc_qh[2] = {0, 0}, cyc_ts[2] = {0, 0}, cyc_alb[2] = {0, 0};
   int cyc_sw[2] = {-1, -1};
   int i_iter = 1;
   for (;;) {
      const bool run = !done;
      if (!SU_SYNC_OR(run)) break;
      SU_CLK(i_iter == 1 ? 1 : 4);
      SU_CNT(8, 1);
      // ================= phase L1: radiation, storage heat, melt heat =================
      if (run) {
         if (i_iter > 1) {
            if (i_iter == 2 && temp_will_rewrite) {   // trips >= 2 run the early stages with the rewritten Temp_C
               su_copy_daily(D, D_prev);
               stage_pre_atm(dg, P, C, R, S, dt_since_start, tstep_prev, nsh_real, Temp_C, H, D);
               stage_post_atm(dg, P, snowUse, C, R, S, A, DLS, nsh, nsh_real, Temp_C, NonWaterFraction, (SNOWK == 0) ? nullptr : SnowFrac_res, H, D, water_late SU_CLK_PASS_NOBAR);
               qf = H.an.qf;
               rad_done = false;
            }
            // cal_AtmMoist of this trip: only Lat_vap can differ (through its SAVE'd switches)
            int sw_in = lv_sw;
            if (lv_exit[sw_in] < 0) {
               lv_memo[sw_in] = Lat_vap(dg, Temp_C, A.Ea_hPa, Press_hPa, A.avcp, lv_sw);
               lv_exit[sw_in] = lv_sw;
            }
            A.lv_J_kg = lv_memo[sw_in];
            lv_sw = lv_exit[sw_in];
         }
         const bool redo_rad = !(rad_invariant && rad_done);
         if (redo_rad) {
            // ---------------- SUEWS_cal_Qn driver.f95:1288-1455
            // alb_next = alb_prev with the three vegetation albedos of the day (driver.f95:1409-1413): the same values on
            // every trip and nothing reads alb_prev afterwards, so the state array is updated in place
            if (nrm_use > 0 && snowUse == 0) {
               for (int is = 0; is < NSURF; is++) SnowFrac_q[is] = 0;
            } else {
               for (int is = 0; is < NSURF; is++) SnowFrac_q[is] = S.SnowFrac[is];
            }
            if (nrm_use > 0) {
               Rd.ldown = (ldown_option == 1) ? R.ldown_obs : -9.0;
               if (ldown_option == 2) Rd.fcld = R.fcld_obs;
               S.alb[DECID] = D.albDecTr_id;
               S.alb[CONIF] = D.albEveTr_id;
               S.alb[GRASS] = D.albGrass_id;
               NARP(sfr, SnowFrac_q, S.alb.v, P.emis.v, S.IceFrac.v, P.NARP_EMIS_SNOW, zenith_deg, Ts_iter, avkdn, Temp_C,
                    avRh, A.Ea_hPa, R.qn1_obs, H.SnowAlb_q, AlbedoChoice, ldown_option, nrm_use, snow_arrays, Rd);
               qn1 = Rd.qn1; qn1_snowfree = Rd.qn1_snowfree; qn1_S = Rd.qn1_S;
            } else {
               for (int is = 0; is < NSURF; is++) { SnowFrac_q[is] = R.snowFrac_obs; Rd.Tsurf_ind[is] = -999; }
               qn1 = R.qn1_obs; qn1_snowfree = R.qn1_obs; qn1_S = R.qn1_obs;
               Rd.tsurf = -999; Rd.kup = -999; Rd.lup = -999; Rd.ldown = -999; Rd.fcld = -999;
            }
            // ---------------- SUEWS_cal_Qs driver.f95:1457-1612
            qs = -999; a1 = -999; a2 = -999; a3 = -999;
            double qn1_use = (P.OHMIncQF == 1) ? qf + qn1 : qn1;
            if (P.StorageHeatMethod == 0) {
               qs = R.qs_obs;
            } else if (P.StorageHeatMethod == 1) {
               OHM(dg, P, snowUse, sfr, qn1_use, S.qn1_av, S.dqndt, qn1_av_next, dqndt_next, qn1_S, S.qn1_s_av, S.dqnsdt, qn1_s_av_next,
                   dqnsdt_next, dt_since_start, D.HDD_id[9], S.soilstore_id.v, S.state_id.v, SnowFrac_q, a1, a2, a3, qs,
                   deltaQi);
            } else if (ESTMK != 0 && P.StorageHeatMethod == 3) {
               // AnOHM (driver.f95:1577-1594); alb is alb_next, the stores are the step's entry values
               if (ES && AN && AN->hours) AnOHM(dg, P, C, *ES, *AN, R.id, sfr, S.state_id.v, S.StoreDrainPrm.v, D.DecidStoreCap, S.alb.v, qn1_use, qf,
                                   S.qn1_av, S.dqndt, dt_since_start, qn1_av_next, dqndt_next, a1, a2, a3, qs, deltaQi);
               else su_fatal(dg, 1003);   // no forcing block with on-the-hour rows behind this call
            } else if (!use_estm) {
               su_fatal(dg, 1003);
            }
         }
         if (use_estm) {
            // ESTM (driver.f95:1598-1606), once per trip: it reads none of cal_Qn's results (the surface temperatures are
            // forcing), so the radiation above stays hoisted when the scheme allows.  Parameters and element temperatures
            // live in MODULE ESTM_data in the reference, in the class record and the per-cell ESTM state here; Tair_av is
            // Tair_av_next (:784)
            if (ES && ts9) qs = ESTM(P, C, *ES, avU1, Temp_C, avRh, Press_hPa, ts9, H.Tair_av_next, L.estm27);
            else su_fatal(dg, 1005);   // ESTM selected but no ESTM state / surface-temperature forcing was supplied
         }
         if (redo_rad || A.lv_J_kg != lv_at_melt) {
            // ---------------- Snow_cal_MeltHeat driver.f95:793-805 (SnowFrac / SnowAlb alias in and out)
            if (!redo_rad)   // the aliased inputs must be the ones cal_Qn produced
               for (int is = 0; is < NSURF; is++)
                  SnowFrac_q[is] = (nrm_use > 0) ? ((snowUse == 0) ? 0.0 : (double)S.SnowFrac[is]) : R.snowFrac_obs;
            SnowAlb_next = H.SnowAlb_q;
            if (snowUse == 1 || want_snow_line) {
               Snow_cal_MeltHeat(P, sfr, A.lvS_J_kg, A.lv_J_kg, tstep_real, Temp_C, Precip, nsh_real, Rd.Tsurf_ind,
                                 S.state_id.v, Rd.qn1_ind_snow, S.SnowWater.v, deltaQi, S.SnowPack.v, SnowFrac_q, SnowAlb_next,
                                 S.SnowDens.v, S.SnowfallCum, SnowPack_m, SnowFrac_q, SnowAlb_next, SnowDens_m,
                                 SnowfallCum_next, Hs);
            } else {
               // snowUse = 0: what Snow_cal_MeltHeat reduces to (snow.f95:160-176, :1299-1315) -- no melt terms, snow
               // fraction 0, pack and density carried; none of the per-surface arrays is formed or read on this path
               Hs.mwh = 0; Hs.Qm = 0; Hs.QmFreez = 0; Hs.QmRain = 0;
               SnowfallCum_next = 0;
               double vf = 0;
               for (int is = 2; is < 7; is++) vf += sfr[is] * (1 - 0.0);
               Hs.veg_fr = vf;
            }
            lv_at_melt = A.lv_J_kg;
         }
         rad_done = true;
         if (i_iter == 1)   // LUMPS first guess (driver.f95:809-824)
            LUMPS_cal_QHQE(dg, P, sfr, qn1, qf, qs, Hs.Qm, Temp_C, Hs.veg_fr, A.avcp, Press_hPa, A.lv_J_kg, D.LAI_id, Lm);
      }
      SU_SYNCL(1);
      SU_CLK(2);
      // ================= phase L2: stability and resistances (driver.f95:2134-2270) =================
      if (run) {
         const double QH_Init = (i_iter == 1) ? Lm.QH_LUMPS : qh;
         cal_Stab(dg, sm, ZZD, z0m, lg_zz0, avU1, Temp_C, QH_Init, A.avdens, A.avcp, L_mod, TStar, UStar, zL);
         RA = AerodynamicResistance(dg, ZZD, z0m, lg_zz0, avU1, L_mod, UStar, VegFraction, site.veg029, P.AerodynamicResistanceMethod,
                                    sm, P.RoughLenHeatMethod);
         if (snowUse == 1)
            RAsnow = AerodynamicResistance(dg, ZZD, z0m, lg_zz0, avU1, L_mod, UStar, VegFraction, site.veg029,
                                           P.AerodynamicResistanceMethod, sm, 3);
         rb = BoundaryLayerResistance(ZZD, z0m, avU1, UStar);
      }
      SU_SYNCL(2);
      SU_CLK(3);
      // ================= phase L3: QE, QH, surface temperature, convergence =================
      if (run) {
         const double qn_e = qn1_snowfree + qf - qs;
         bool snow_branch = false;
         if (snowUse == 1)
            for (int is = 0; is < NSURF; is++) snow_branch = snow_branch || (Hs.snowCalcSwitch[is] == 1);
         if (snow_branch) {
            qe = cal_QE_light_snow(dg, P, C, R, S, A, Lm, H, Hs, Rd.Tsurf_ind, SnowPack_m, SnowFrac_q, SnowDens_m, Temp_C, avRh, RA,
                                   RAsnow, rb, qn_e, PervFraction, VegFraction, W, D);
            qe_full_done = false;
         } else {
            qe = cal_QE_light(P, C, S, A, Lm, H, RA, rb, qn_e, D);
            qe_full_done = false;
         }
         // SUEWS_cal_QH driver.f95:2084-2131 (QHMethod = 1) and cal_tsfc :3558
         qh = (qn1 + qf + Hs.QmRain) - (qe + qs + Hs.Qm + Hs.QmFreez);
         TSfc_C = qh / (A.avdens * A.avcp) * RA + Temp_C;
         const bool converged = !(fabs(Ts_iter - TSfc_C) > F(0.1));
         Ts_iter = TSfc_C;
         i_iter = i_iter + 1;
         if (converged || !(i_iter < 100) || dg.fatal != 0) done = true;
         if (!done) {
            // A trip is a pure function of what the previous trip left behind: (qh, Ts_iter, the Lat_vap switch
            // state, NARP's persistent alb1); everything else restarts from the *_prev state.  When the loop does
            // not converge it falls into an exact 2-cycle within 3-4 trips (stable nights: L_mod flips between
            // two values), and the reference then burns all 99 trips of DO WHILE (driver.f95:642).  Once the
            // state after trip n equals, bit for bit, the state after trip n-2, trips n+1.. repeat trips n-1,n..
            // exactly, so the state after trip 99 is the one after trip n (99-n even) or n+1 (99-n odd).
            const int n = i_iter - 1;
#ifndef SU_NO_CYCLE
            // (with ESTM a trip also depends on the element temperatures, which never repeat: no short cut)
            if (n >= 4 && !use_estm && qh == cyc_qh[0] && Ts_iter == cyc_ts[0] && lv_sw == cyc_sw[0] && Rd.alb1 == cyc_alb[0]) {
               if (((99 - n) & 1) == 0) { done = true; i_iter = 100; }
               else i_iter = 99;   // one more trip, which then ends the loop as the 99th
            }
#endif
            cyc_qh[0] = cyc_qh[1]; cyc_ts[0] = cyc_ts[1]; cyc_sw[0] = cyc_sw[1]; cyc_alb[0] = cyc_alb[1];
            cyc_qh[1] = qh; cyc_ts[1] = Ts_iter; cyc_sw[1] = lv_sw; cyc_alb[1] = Rd.alb1;
         }
      }
   }
   SU_SYNCN(3);
   SU_CLK(4);
   // ================= phase F1: state update of the last trip + soil water =================
   if (live && dg.fatal == 0) {
      if (water_late) {
         stage_water_a(P, C, R, nsh, nsh_real, DLS, H, D);
         stage_water_b(dg, P, snowUse, C, S, nsh_real, NonWaterFraction, H, D SU_CLK_PASS_NOBAR);
      }
      constexpr bool want_out = (K != 2);    // policy P0 forms no output line: SMD columns and TotCh are not needed
      double soilstoreOld[NSURF];
      if (want_out)
         for (int is = 0; is < NSURF; is++) soilstoreOld[is] = S.soilstore_id[is];
      {   // the state update of the step: once, with the converged trip's resistances (stores updated in place)
         const double qn_e = qn1_snowfree + qf - qs;
         cal_QE_full(dg, P, snowUse, C, R, S, A, Lm, H, Hs, Rd.Tsurf_ind, SnowPack_m, SnowFrac_q, SnowDens_m, Temp_C, avRh, RA, RAsnow, rb,
                     qn_e, PervFraction, VegFraction, NonWaterFraction, snow_arrays, W, QT, D);
      }
      SU_CLKB(15);
      qh_resist = (RA != 0) ? A.avdens * A.avcp * (Rd.tsurf - Temp_C) / RA : -999.0;
#if !(SU_SKIP & 128)
      cal_HorizontalSoilWater(P, sfr, C.SurfaceArea, NonWaterFraction, tstep_real, W.soilstore_id, W.runoffSoil,
                              runoffSoil_per_tstep);
#endif
      cal_SoilState<want_out>(dg, P.SMDMethod, R.xsmd, NonWaterFraction, H.SoilMoistCap, P.SoilStoreCap.v, QT.surf_chang_per_tstep,
                              W.soilstore_id, soilstoreOld, sfr, smd, smd_nsurf, tot_chang_per_tstep);
   }
   SU_SYNCN(4);
   SU_CLK(5);
   // ================= phase F2: roughness-sublayer diagnostics (driver.f95:943-949) =================
   SuRSL rs;
   rs.T2_C = 0; rs.q2_gkg = 0; rs.U10_ms = 0; rs.RH2 = 0;
   double Fc_biogen = 0, Fc_photo = 0, Fc_respi = 0;
   // flux_only (SUEWS_OUT_FLUX_ONLY, opt-in): T2/q2/U10/RH2 reach no flux sum and no state; the biogenic CO2 families that
   // read the modelled 2 m temperature (EmissionsMethod >= 11) keep the stage
   if (live && dg.fatal == 0 && !(flux_only && P.EmissionsMethod < 11 && !L.rsl133 && !L.suews82 && !acc82)) {
      RSLProfile(sm, zH, z0m, zdm, L_mod, sfr, FAI, A.avcp, A.lv_J_kg, A.avdens, avU1, Temp_C, avRh, Press_hPa, C.Z, qh, qe,
                 S.rsl_cache.v, rs, L.rsl133);
      // SUEWS_cal_BiogenCO2 driver.f95:1156-1286: zero for EmissionsMethod 0-6, CO2_biogen for >= 11.  It sees the
      // still un-committed SnowFrac (driver.f95:957) and, for gsModel 3/4, a second cal_AtmMoist / SurfaceResistance
      // at the modelled 2 m temperature (which also steps the Lat_vap switches, as in the reference).
      if (P.EmissionsMethod >= 11) {
         double gfunc2 = 0, t2 = 0;
         if (P.gsModel == 3 || P.gsModel == 4) {
            t2 = (fabs(Temp_C - rs.T2_C) > 5) ? Temp_C : rs.T2_C;
            SuAtm A2;
            cal_AtmMoist(dg, t2, Press_hPa, avRh, lv_sw, A2);
            double gsc2, rs2;
            SurfaceResistance(dg, P, sfr, S.SnowFrac.v, avkdn, t2, A2.dq, R.xsmd, H.vsmd, D.LAI_id, gfunc2, gsc2, rs2);
         }
         // SUEWS_cal_Main's own `gfunc` (driver.f95:374) is never assigned -- SUEWS_cal_Resistance keeps its gfunc
         // to itself (driver.f95:2199) -- so the conductance-based family with gsModel 1/2 multiplies by an
         // uninitialised local there; like the oracle, this takes it as 0
         Fc_biogen = CO2_biogen(P, sfr, S.SnowFrac.v, D.LAI_id, avkdn, 0.0, gfunc2, t2, Temp_C, Fc_photo, Fc_respi);
      }
   }
   SU_SYNCN(5);
   SU_CLK(6);
   SU_CNT(9, 1);
   // ================= phase F3: commit (driver.f95:965-994) and output lines =================
   if (!(live && dg.fatal == 0)) return i_iter - 1;

   S.qn1_av = qn1_av_next; S.dqndt = dqndt_next; S.qn1_s_av = qn1_s_av_next; S.dqnsdt = dqnsdt_next;
   S.SnowfallCum = SnowfallCum_next; S.SnowAlb = SnowAlb_next;
   for (int is = 0; is < NSURF; is++) {
      if (snow_arrays) {
         S.IceFrac[is] = W.iceFrac[is]; S.SnowWater[is] = W.SnowWater[is]; S.SnowDens[is] = W.SnowDens[is];
         S.SnowFrac[is] = W.SnowFrac[is]; S.SnowPack[is] = W.SnowPack[is];
      } else if (SNOWK != 0) {
         S.SnowFrac[is] = 0;   // snowUse = 0: ice fraction, snow water, density and pack are carried unchanged
      }                        // (SNOWK == 0: nothing reads SnowFrac during a snow-off launch; the kernel zeroes it once at the end)
   }
   // Daily state: D already holds this step's values (stage_pre_atm); on the last step of the day the one element of
   // StoreDrainPrm that evolves follows (dailystate.f95:671)
   S.Tair_av = H.Tair_av_next;
   if ((R.it * 3600 + R.imin * 60 + P.tstep) >= 86400) SU_SDP(S.StoreDrainPrm.v, 5, DECID) = D.DecidStoreCap;
   S.lv_switch = (double)lv_sw;
   S.alb1_hidden = Rd.alb1;
   // BLUEWS (suews_ctrl_calculations.f95:194-218): after the step, from its QH / QE (SUEWS's or LUMPS's), UStar and the
   // moist-air values of the last trip
   if (ESTMK != 0 && P.CBLuse >= 1) {
      if (ES && AN) BLUEWS(dg, P, *ES, AN->cbl_ini, AN->cbl_nline, R, (P.QH_choice == 2) ? Lm.QH_LUMPS : qh, (P.QH_choice == 2) ? Lm.QE_LUMPS : qe,
                           UStar, A.avdens, A.lv_J_kg, A.avcp, ZZD / L_mod, L.bl17);
      else su_fatal(dg, 1007);
   }

   // output policy P0: running sums of the seven parity fluxes only -- QN, QF, QS, QH, QE, RO, SWE,
   // the same values the 82-column line carries in columns 6-10, 17 and 72
   if (acc7) {
      acc7[0] += qn1; acc7[1] += qf; acc7[2] += qs; acc7[3] += qh; acc7[4] += qe;
      acc7[5] += QT.runoff_per_tstep; acc7[6] += W.swe;
   }
   // SUEWS_update_outputLine driver.f95:2273-2457
   if (acc82 || L.suews82) {
      int k = 0;
#define SU_OUT(v)                                  \
   do {                                            \
      const double v_ = (v);                       \
      /* columns aggregated as "last value" (bit set) are overwritten, the others summed */ \
      if (acc82) acc82[k] = (((k < 64 ? last_lo >> k : last_hi >> (k - 64)) & 1ull) ? 0.0 : acc82[k]) + v_; \
      if (L.suews82) L.suews82[k] = v_;            \
      k++;                                         \
   } while (0)
      const bool rad = nrm_use > 0;
      const double LAI_wt = D.LAI_id[0] * sfr[2] + D.LAI_id[1] * sfr[3] + D.LAI_id[2] * sfr[4];
      double bulkalbedo = 0;
      for (int is = 0; is < NSURF; is++) bulkalbedo += S.alb[is] * sfr[is];
      SU_OUT(avkdn); SU_OUT(rad ? Rd.kup : -999.0); SU_OUT(rad ? Rd.ldown : -999.0); SU_OUT(rad ? Rd.lup : -999.0);
      SU_OUT(rad ? Rd.tsurf : -999.0);
      SU_OUT(qn1); SU_OUT(qf); SU_OUT(qs); SU_OUT(qh); SU_OUT(qe);
      SU_OUT(Lm.QH_LUMPS); SU_OUT(Lm.QE_LUMPS); SU_OUT(qh_resist);
      SU_OUT(Precip); SU_OUT(H.wu_ext); SU_OUT(QT.ev_per_tstep); SU_OUT(QT.runoff_per_tstep); SU_OUT(tot_chang_per_tstep);
      SU_OUT(QT.surf_chang_per_tstep); SU_OUT(QT.state_per_tstep); SU_OUT(QT.NWstate_per_tstep); SU_OUT(H.drain_per_tstep);
      SU_OUT(smd);
      SU_OUT(C.FlowChange / nsh_real); SU_OUT(0.0);   // AdditionalWater = addPipes + ... = 0 (driver.f95:1669)
      SU_OUT(runoffSoil_per_tstep); SU_OUT(W.fl.runoffPipes); SU_OUT(W.fl.runoffAGimpervious); SU_OUT(W.fl.runoffAGveg);
      SU_OUT(W.fl.runoffWaterBody);
      SU_OUT(H.wu_int); SU_OUT(H.wu_nsurf[3]); SU_OUT(H.wu_nsurf[2]); SU_OUT(H.wu_nsurf[4]);   // EveTr/DecTr swap, driver.f95:2421-2439
      for (int is = 0; is < NSURF - 1; is++) SU_OUT((sfr[is] < F(0.00001)) ? -999.0 : smd_nsurf[is]);
      for (int is = 0; is < NSURF; is++) SU_OUT((sfr[is] < F(0.00001)) ? -999.0 : (double)S.state_id[is]);
      SU_OUT(zenith_deg); SU_OUT(azimuth); SU_OUT(bulkalbedo);
      SU_OUT(rad ? ((ldown_option == 1) ? -999.0 : Rd.fcld) : -999.0);
      SU_OUT(LAI_wt); SU_OUT(z0m); SU_OUT(zdm);
      SU_OUT(UStar); SU_OUT(L_mod); SU_OUT(RA); SU_OUT(H.ResistSurf);
      SU_OUT(H.an.Fc_anthro + Fc_biogen); SU_OUT(Fc_photo); SU_OUT(Fc_respi); SU_OUT(H.an.Fc_metab); SU_OUT(H.an.Fc_traff);
      SU_OUT(H.an.Fc_build); SU_OUT(H.an.Fc_point);
      SU_OUT(qn1_snowfree); SU_OUT(qn1_S); SU_OUT((double)S.SnowAlb);
      SU_OUT(Hs.Qm); SU_OUT(Hs.QmFreez); SU_OUT(Hs.QmRain); SU_OUT(W.swe); SU_OUT(Hs.mwh); SU_OUT(W.mwstore);
      SU_OUT(QT.chSnow_per_interval); SU_OUT(W.SnowRemoval[0]); SU_OUT(W.SnowRemoval[1]);
      SU_OUT(TSfc_C); SU_OUT(rs.T2_C); SU_OUT(rs.q2_gkg); SU_OUT(rs.U10_ms); SU_OUT(rs.RH2 * F(100.0));
#undef SU_OUT
   }
   if (L.datetime5) {
      L.datetime5[0] = (double)R.iy; L.datetime5[1] = (double)R.id; L.datetime5[2] = (double)R.it;
      L.datetime5[3] = (double)R.imin; L.datetime5[4] = R.dectime;
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
   if (L.sol26) {   // SOLWEIG point diagnostics, driver.f95:997-1002 (after the commit: alb is alb_next)
      if (sfr[BLDG] > 0) {
         SOLWEIG_cal_main(R.id, R.it, R.dectime, 0.8, FAI, avkdn, Rd.ldown, Temp_C, avRh, Press_hPa, TSfc_C, C.lat, zenith_deg,
                          azimuth, 1.0, S.alb[0], S.alb[1], P.emis[0], P.emis[1], C.bldgH, L.sol26);
      } else {
         for (int i = 0; i < SUEWS_NCOL_SOLWEIG; i++) L.sol26[i] = -999;
      }
   }
   if (L.daily45) {   // update_DailyStateLine, dailystate.f95:1044-1113 (HDD_id is DIMENSION(6) in the packer)
      double *o = L.daily45;
      for (int i = 0; i < SUEWS_NCOL_DAILY; i++) o[i] = -999;
      if (R.it == 23 && R.imin == (nsh_real - 1) / nsh_real * 60) {
         int k = 0;
         for (int i = 0; i < 6; i++) o[k++] = D.HDD_id[i];
         for (int i = 0; i < 3; i++) o[k++] = D.GDD_id[i];
         for (int i = 0; i < 3; i++) o[k++] = D.SDD_id[i];
         o[k++] = D.Tmin_id; o[k++] = D.Tmax_id; o[k++] = D.lenDay_id;
         for (int i = 0; i < 3; i++) o[k++] = D.LAI_id[i];
         o[k++] = D.DecidCap_id; o[k++] = D.porosity_id; o[k++] = D.albEveTr_id; o[k++] = D.albDecTr_id; o[k++] = D.albGrass_id;
         for (int i = 0; i < 9; i++) o[k++] = D.WUDay_id[i];
         o[k++] = D.deltaLAI; o[k++] = Lm.VegPhenLumps; o[k++] = S.SnowAlb;
         for (int is = 0; is < NSURF; is++) o[k++] = S.SnowDens[is];
         o[k++] = a1; o[k++] = a2; o[k++] = a3;
      }
   }
   SU_CLK(7);
   return i_iter - 1;
}

#ifndef SU_INLINE_ALL
#undef exp
#undef log
#undef pow
#undef sin
#undef cos
#undef tan
#undef atan
#undef atan2
#undef asin
#undef acos
#undef tanh
#endif
