// dev_snow.cuh -- snow pack energy and mass balance.  Device restatement of suews_phys_snow.f95.
#pragma once
#include "dev_surface.cuh"

#define SU_WATERDENS F(999.8395)   // snow.f95:54

// snow.f95:1411-1452
SU_DEV double update_snow_albedo(int tstep, const double *SnowPack_prev, double SnowAlb_prev, double Temp_C,
                                 double tau_a, double tau_f, double SnowAlbMax, double SnowAlbMin)
{
   const double tau_1 = 24 * 60 * 60;
   double s = 0;
   for (int is = 0; is < NSURF; is++) s += SnowPack_prev[is];
   double a = 0;
   if (s > 0) {
      if (Temp_C < 0) {
         a = SnowAlb_prev - tau_a * (tstep) / tau_1;
      } else {
         double ch = exp(-tau_f * (tstep) / tau_1);
         a = (SnowAlb_prev - SnowAlbMin) * ch + SnowAlbMin;
      }
      if (a < SnowAlbMin) a = SnowAlbMin;
      if (a > SnowAlbMax) a = SnowAlbMax;
   }
   return a;
}

// Everything Snow_cal_MeltHeat hands on to the rest of the step (snow.f95:33-193)
struct SuSnowHeat {
   double mwh, Qm, QmFreez, QmRain, veg_fr;
   int snowCalcSwitch[NSURF];
   double Qm_melt[NSURF], Qm_freezState[NSURF], Qm_rain[NSURF], FreezMelt[NSURF], FreezState[NSURF],
      FreezStateVol[NSURF], rainOnSnow[NSURF], SnowDepth[NSURF], mw_ind[NSURF];
};

// snow.f95:33-193 (Snow_cal_MeltHeat) + :195-473 (MeltHeat) + :1454-1489 (update_snow_dens).
// In/out arrays may alias exactly as in the reference call (driver.f95:793-805): inputs are read
// before any output is written.
SU_DEV void Snow_cal_MeltHeat(const SuewsClassParams &P, const double *sfr, double lvS_J_kg, double lv_J_kg,
                                double tstep_real, double Temp_C, double Precip, double nsh_real,
                                const double *Tsurf_ind, const double *state_id, const double *qn1_ind_snow,
                                const double *SnowWater, const double *deltaQi, const double *SnowPack_in,
                                const double *SnowFrac_in, double SnowAlb_in, const double *SnowDens_in,
                                double SnowfallCum_in, double *SnowPack_out, double *SnowFrac_out, double &SnowAlb_out,
                                double *SnowDens_out, double &SnowfallCum_out, SuSnowHeat &h)
{
   double SnowPack[NSURF], SnowFrac[NSURF], SnowDens[NSURF];
   double SnowAlb = SnowAlb_in, SnowfallCum = SnowfallCum_in;
   for (int is = 0; is < NSURF; is++) {
      SnowPack[is] = SnowPack_in[is];
      SnowFrac[is] = SnowFrac_in[is];
      SnowDens[is] = SnowDens_in[is];
   }
   h.mwh = 0; h.Qm = 0; h.QmFreez = 0; h.QmRain = 0;
   for (int is = 0; is < NSURF; is++) {
      h.snowCalcSwitch[is] = 0; h.Qm_melt[is] = 0; h.Qm_freezState[is] = 0; h.Qm_rain[is] = 0; h.FreezMelt[is] = 0;
      h.FreezState[is] = 0; h.FreezStateVol[is] = 0; h.rainOnSnow[is] = 0; h.SnowDepth[is] = 0; h.mw_ind[is] = 0;
   }

   if (P.snowUse == 1) {
      // update_snow_dens: the reference passes SnowFrac_in where SnowPack_prev is expected (snow.f95:137-139)
      const double tau_1 = 24 * 60 * 60;
      const double dens_change = exp(-P.tau_r * (P.tstep) / tau_1);
      for (int is = 0; is < NSURF; is++) {
         if (SnowFrac_in[is] > 0) {
            SnowDens[is] = (SnowDens_in[is] - P.SnowDensMax) * dens_change + P.SnowDensMax;
            if (SnowDens[is] > P.SnowDensMax) SnowDens[is] = P.SnowDensMax;
         } else {
            SnowDens[is] = P.SnowDensMin;
         }
      }
      // MeltHeat snow.f95:195-473
      const double cw = 4190;
      const double TempMeltFact = P.TempMeltFact;
      for (int is = 0; is < NSURF; is++) {
         if (sfr[is] != 0) {
            if (SnowPack[is] > 0) {
               h.SnowDepth[is] = (SnowPack[is] / 1000) * SU_WATERDENS / SnowDens[is];
               double mw;
               if (Temp_C >= 0) {
                  mw = (qn1_ind_snow[is] < 0) ? TempMeltFact * Temp_C : P.RadMeltFact * (qn1_ind_snow[is]);
               } else {
                  mw = TempMeltFact * Temp_C * 1.0;
               }
               mw = mw / nsh_real;
               if (mw > SnowPack[is]) mw = SnowPack[is];
               h.Qm_melt[is] = SU_WATERDENS * ((mw / tstep_real) / 1000) * (lvS_J_kg - lv_J_kg);
               if (mw < 0) {
                  h.FreezMelt[is] = -mw;
                  mw = 0;
                  if (h.FreezMelt[is] > SnowWater[is]) h.FreezMelt[is] = SnowWater[is];
                  h.Qm_melt[is] = SU_WATERDENS * su_div0(su_div0(-h.FreezMelt[is], tstep_real), 1000) * (lvS_J_kg - lv_J_kg);
               }
               h.mw_ind[is] = mw;
               if (Temp_C >= P.PrecipLimit && Precip > 0) {
                  h.Qm_rain[is] = SU_WATERDENS * cw * (Temp_C - P.PrecipLimit) * (Precip * F(0.001) / tstep_real);
                  if (h.Qm_rain[is] < 0) h.Qm_rain[is] = 0;
                  else h.rainOnSnow[is] = Precip;
               }
            }
            if (Tsurf_ind[is] < 0 && state_id[is] > 0) {
               h.snowCalcSwitch[is] = 1;
               if (is != WATER) {
                  double fs = -TempMeltFact * Tsurf_ind[is] / nsh_real;
                  if (fs > state_id[is]) fs = state_id[is];
                  double fsv;
                  if (SnowPack[is] == 0 || SnowFrac[is] == 0) fsv = fs;
                  else fsv = su_div0(fs * (1 - SnowFrac[is]), SnowFrac[is]);
                  if (fsv < F(0.00000000001) && fs < state_id[is]) { fs = 0; fsv = 0; }
                  h.FreezState[is] = fs;
                  h.FreezStateVol[is] = fsv;
                  h.Qm_freezState[is] = -SU_WATERDENS * (fs / tstep_real / 1000) * (lvS_J_kg - lv_J_kg);
               } else {
                  double Watfreeze = 100 * (0 - Temp_C) / (SU_WATERDENS * (lvS_J_kg - lv_J_kg));
                  h.FreezState[is] = Watfreeze;
                  h.Qm_freezState[is] = -SU_WATERDENS * (Watfreeze / tstep_real / 1000) * (lvS_J_kg - lv_J_kg);
               }
            }
            if (is != WATER) {
               if (SnowPack[is] > 0 || (Precip > 0 && Tsurf_ind[is] < 0)) h.snowCalcSwitch[is] = 1;
            } else {
               if (SnowPack[WATER] > 0 || h.FreezState[WATER] > 0) h.snowCalcSwitch[WATER] = 1;
            }
            if (Precip > 0 && Tsurf_ind[is] < 0 && SnowPack[is] > 0) {
               SnowDens[is] = SnowDens[is] * SnowPack[is] / (SnowPack[is] + Precip) +
                              P.SnowDensMin * Precip / (SnowPack[is] + Precip);
            }
            h.mwh = h.mwh + h.mw_ind[is] * sfr[is] * SnowFrac[is];
            h.Qm = h.Qm + h.Qm_melt[is] * sfr[is] * SnowFrac[is];
            h.QmRain = h.QmRain + h.Qm_rain[is] * sfr[is] * SnowFrac[is];
            h.QmFreez = h.QmFreez + deltaQi[is] * sfr[is] * SnowFrac[is] + h.Qm_freezState[is] * sfr[is] * (1 - SnowFrac[is]);
         }
      }
      double s = 0;
      for (int is = 0; is < NSURF; is++) s += SnowPack[is];
      if (Precip > 0 && s > 0 && Temp_C < 0) {
         SnowfallCum = SnowfallCum + Precip;
         if (SnowfallCum > P.PrecipLimitAlb) {
            SnowAlb = P.SnowAlbMax;
            SnowfallCum = 0;
         }
      } else {
         SnowfallCum = 0;
      }
   } else {
      SnowfallCum = 0;
      for (int is = 0; is < NSURF; is++) SnowFrac[is] = 0;
   }

   h.veg_fr = 0;   // veg_fr_snow snow.f95:1299-1315
   for (int is = 2; is < 7; is++) h.veg_fr += sfr[is] * (1 - SnowFrac[is]);

   SnowAlb_out = SnowAlb;
   SnowfallCum_out = SnowfallCum;
   for (int is = 0; is < NSURF; is++) {
      SnowPack_out[is] = SnowPack[is];
      SnowFrac_out[is] = SnowFrac[is];
      SnowDens_out[is] = SnowDens[is];
   }
}

// snow.f95:1243-1297
SU_DEV double SnowDepletionCurve(int is, double swe, double sweD)
{
   double asc = 1;
   if (is == PAV) {
      if (swe <= sweD) { double r = swe / sweD; asc = r * r; }
   } else if (is == BLDG) {
      if (swe <= sweD) {
         if ((swe / sweD) < F(0.9)) asc = (swe / sweD) * F(0.5);
         else asc = su_p8(swe / sweD);
      }
   } else if (is == WATER) {
      asc = 1;
   } else {
      if (swe <= sweD) asc = 1 - pow((double)(1 / 3.1416f) * acos(2 * (swe / sweD) - 1), F(1.7));
   }
   return asc;
}

// Working set shared by the 7 per-surface calls inside SUEWS_cal_QE (driver.f95:1973-2059)
struct SuQEWork {
   // The surface and soil stores are updated IN PLACE in the cell's state record: the committing evaluation runs once per
   // step, after the surface-temperature loop, every surface reads its old value before it writes the new one, and the
   // step's commit then has nothing to copy (the reference's *_next arrays, driver.f95:1973-2059)
   double *state_id, *soilstore_id;
   double SnowPack[NSURF], SnowFrac[NSURF], SnowWater[NSURF], iceFrac[NSURF],
      SnowDens[NSURF], SurplusEvap[NSURF], runoffSoil[NSURF], SnowDepth[NSURF];
   double SnowRemoval[2];
   SuFlood fl;
   double swe, mwstore;
};

// snow.f95:477-1148 for surface `is` (0-based).  Outputs the five *_tot contributions.
// The routine only ever touches element `is` of the per-surface work arrays (plus the evaporation surplus of the two
// impervious surfaces and the flood totals), so it works on register copies of those elements: `in` says where their
// entry values are read from, and with COMMIT they are written back to w at the end.  COMMIT = false is the evaluation a
// trip of the surface-temperature loop needs -- qe_tot (and SurplusEvap[is]) by the very same statements, nothing stored:
// the reference restarts every trip from the *_prev state, so only the last trip's update survives (dev_step.cuh).
struct SuQEIn {
   const double *state_id, *soilstore_id, *SnowPack, *SnowFrac, *SnowWater, *SnowDens;
};
// Terms of Evap_SUEWS_Snow (snow.f95:1154-1207) that are the same for the seven surfaces of a trip: formed by the first
// snow-covered surface that needs them (with the warnings the ice vapour pressure may raise), reused by the others.
struct SuSnowIce {
   bool have;
   double vdrc_RA;   // vdrcIce / RAsnow
   double tlv_sub;   // lvS_J_kg / tstep
};
// (inlined into its two callers -- the committing loop of cal_QE_full and the light loop of a trip: as a real function
// its 45 arguments and five by-reference results travelled through the stack, 10 % of the snow kernel's local traffic)
template <bool COMMIT>
SU_DEV void SnowCalc(SuDiag &dg, const SuewsClassParams &P, const SuewsCellParams &C, int is, int imin, int it, int dow,
                       double nsh_real, double lvS_J_kg, double avdens, double avRh, double Press_hPa, double Temp_C,
                       double RAsnow, double psyc_hPa, double avcp, double sIce_hPa, double PervFraction,
                       double vegfraction, double vpd_hPa, double qn_e, double s_hPa, double ResistSurf, double RA,
                       double rb, double tlv, double Precip, const double *drain, const double *stateOld,
                       const SuSnowHeat &h, const double *Tsurf_ind, const double *StoreDrainPrm, double capStore_is,
                       const double *AddWater,
                       const double *AddWaterRunoff, const SuQEIn &in, double *SurplusEvap, SuEvapTrip &et, SuSnowIce &ice, SuQEWork &w, double &chSnow_tot, double &ev_tot, double &qe_tot,
                       double &runoff_tot, double &surf_chang_tot)
{
   const double *sfr = C.sfr.v;
   const double IPT = 10 / nsh_real;
   double runoffSnow = 0, runoff = 0, chang = 0, changSnow = 0, runoffTest = 0, SnowToSurf = 0, EvPart = 0, ev = 0,
          qe = 0, ev_snow = 0, tlv_sub = 0;
   double snowFracFresh1 = 0, snowFracFresh2 = 0, snowFracOld = 0;
   double sp_ = in.SnowPack[is], sf_ = in.SnowFrac[is], sw_ = in.SnowWater[is], st_ = in.state_id[is], so_ = in.soilstore_id[is],
          sd_ = in.SnowDens[is], ic_ = 0;
   bool ic_set = false;
   SuFlood fl_local;   // (COMMIT = false: the flood totals are not formed)
   SuFlood &fl = COMMIT ? w.fl : fl_local;
   fl_local.runoffAGimpervious = 0; fl_local.surplusWaterBody = 0; fl_local.runoffAGveg = 0; fl_local.runoffPipes = 0; fl_local.runoffWaterBody = 0;
#define SU_SNOW_COMMIT()                                                                                         \
   do {                                                                                                          \
      if (COMMIT) {                                                                                              \
         w.SnowPack[is] = sp_; w.SnowFrac[is] = sf_; w.SnowWater[is] = sw_; w.state_id[is] = st_;               \
         w.soilstore_id[is] = so_; w.SnowDens[is] = sd_;                                                         \
         if (ic_set) w.iceFrac[is] = ic_;                                                                        \
      }                                                                                                          \
   } while (0)
   const double mw_ind = h.mw_ind[is], freezMelt = h.FreezMelt[is], FreezState = h.FreezState[is],
                FreezStateVol = h.FreezStateVol[is], rainOnSnow = h.rainOnSnow[is];
   const int u = (dow == 1 || dow == 7) ? 1 : 0;
   const bool rem_time = (is < 2 && SU_PROF24(P.SnowProf_24hr.v, it, u) == 1 && (imin == (nsh_real - 1) / nsh_real * 60));

   ev_tot = 0; qe_tot = 0; runoff_tot = 0; surf_chang_tot = 0; chSnow_tot = 0;
   if (COMMIT) {
      w.runoffSoil[is] = 0;
      for (int k = 0; k < NSURF; k++) w.SnowDepth[k] = 0;   // INTENT(OUT) array reset on every call
   }

   const double snowTotInit = sp_ + sw_;
   const double WaterHoldCapFrac =
      (sd_ >= 200) ? P.CRWmin : P.CRWmin + (P.CRWmax - P.CRWmin) * (200 - sd_) / 200;

   if (sf_ < 1) evap_surface<true>(et, st_, P.WetThresh[is], capStore_is, ev, qe);   // cal_evap with the trip's shared terms

   if (sf_ > 0) {   // Evap_SUEWS_Snow snow.f95:1154-1207
      if (!ice.have) {
         double T = Temp_C;
         double esIce = sat_vap_pressIce(dg, T, Press_hPa, 1);
         double EaIce = avRh / 100 * esIce;
         double vdrcIce = (esIce - EaIce) * avdens * avcp;
         ice.vdrc_RA = su_div0(vdrcIce, RAsnow);   // zero numerators are the rule here (su_div0)
         ice.tlv_sub = lvS_J_kg / (double)P.tstep;
         ice.have = true;
      }
      double sae_snow = sIce_hPa * (h.Qm_rain[is] - h.Qm_melt[is]);
      tlv_sub = ice.tlv_sub;
      double e_snow = sae_snow + ice.vdrc_RA;
      double qe_snow = su_div0(e_snow, sIce_hPa + psyc_hPa);
      ev_snow = su_div0(qe_snow, tlv_sub);
   }

   if (is > 1 && PervFraction != 0)
      EvPart = su_div0(SurplusEvap[PAV] * sfr[PAV] + SurplusEvap[BLDG] * sfr[BLDG], PervFraction);

   if (is == WATER && sfr[WATER] > 0) {
      // label 606, snow.f95:1083-1146
      const double lim = SU_SDP(StoreDrainPrm, 4, WATER);
      changSnow = (Precip + freezMelt + FreezState) - (mw_ind + ev_snow);
      sp_ = sp_ + changSnow;
      st_ = st_ + C.FlowChange - FreezState;
      if (sp_ > 0) {
         sw_ = sw_ + mw_ind - freezMelt;
         double FWC = WaterHoldCapFrac * sp_;
         if (sw_ >= FWC && Temp_C >= 0) {
            st_ = st_ + (sw_ - FWC);
            sw_ = FWC;
         }
      } else if (sp_ < 0) {
         sw_ = sw_ - freezMelt + mw_ind;
         st_ = st_ + sw_ + sp_;
         sp_ = 0;
         if (st_ < 0) ev_snow = ev_snow + st_;
      }
      if (st_ > lim) {
         runoff = runoff + (st_ - lim);
         st_ = lim;
         fl.runoffWaterBody = fl.runoffWaterBody + runoff * sfr[WATER];
      } else {
         st_ = st_ + fl.surplusWaterBody;
         if (st_ > lim) {
            fl.runoffWaterBody = fl.runoffWaterBody + (st_ - lim) * sfr[WATER];
            st_ = lim;
         }
      }
      chSnow_tot = ((sp_ + sw_) - snowTotInit) * sfr[WATER];
      surf_chang_tot = (st_ - stateOld[WATER]) * sfr[WATER];
      ev_tot = ev * sfr[WATER] + ev_snow * sfr[WATER];
      qe_tot = ev_snow * tlv_sub * sfr[WATER] + ev * tlv * sfr[WATER];
      runoff_tot = runoff;
      sf_ = (sp_ > 0) ? 1.0 : 0.0;
      SU_SNOW_COMMIT();
      return;
   }

   // which arm of the IF / ELSEIF at snow.f95:727 / :781 this call takes (decided on entry values)
   const bool armFull = (sp_ > 0 && sf_ == 1);
   const bool armPart = !armFull && (sf_ < 1);

   if (armFull) {
      ev_snow = ev_snow + EvPart;
      changSnow = (Precip + freezMelt) - (mw_ind + ev_snow);
      if (rainOnSnow > 0) {
         changSnow = changSnow - Precip;
         sw_ = sw_ + rainOnSnow;
      }
      sp_ = sp_ + changSnow;
      if (sp_ > 0) {
         sw_ = sw_ + mw_ind - freezMelt;
         double FWC = WaterHoldCapFrac * sp_;
         if (sw_ >= FWC) {
            runoffSnow = runoffSnow + (sw_ - FWC);
            sw_ = FWC;
         }
      } else if (sp_ < 0) {
         sw_ = sw_ - freezMelt + mw_ind + sp_;
         sp_ = F(0.0);
         snowFracOld = 1;
         sf_ = 0;
         sd_ = 0;
         if (sw_ < 0) {
            ev_snow = ev_snow + sw_;
            if (ev_snow < 0) ev_snow = 0;
            changSnow = changSnow + sw_;
            sw_ = 0;
         } else {
            chang = sw_;
            st_ = st_ + chang;
            sw_ = 0;
         }
      }
   } else if (armPart) {
      if (sp_ > 0) {
         ev_snow = ev_snow + EvPart;
         changSnow = (Precip + freezMelt + FreezStateVol) - (mw_ind + ev_snow);
         if (rainOnSnow > 0) {
            changSnow = changSnow - Precip;
            sw_ = sw_ + rainOnSnow;
         }
         sp_ = sp_ + changSnow;
         {   // snow.f95:810-820.  Written with named conditions: as an if / else-if / else-if chain on the array
             // element, nvcc 12.9 (sm_100a, -O3) dropped the third assignment (found by the 2004 annual parity run).
            const double st_now = st_;
            const bool c1 = (Precip > 0 && FreezState == st_now);
            const bool c2 = (Precip == 0 && FreezState > 0 && FreezState == st_now);
            const bool c3 = (FreezState > 0 && FreezState < st_now);
            if (c1 || c2) snowFracFresh1 = 1;
            else if (c3) snowFracFresh1 = F(0.95);
         }
      } else if (sp_ == 0 && Tsurf_ind[is] < 0) {
         if ((Precip > 0 && st_ == 0) || (Precip == 0 && FreezState == st_) ||
             (Precip > 0 && FreezState == st_)) {
            changSnow = Precip + FreezStateVol;
            sp_ = sp_ + changSnow;
            snowFracFresh1 = 1;
            ic_ = FreezState / (FreezState + Precip); ic_set = true;
            sd_ = P.SnowDensMin;
         }
         if (FreezState > 0 && FreezState < st_) {
            changSnow = Precip + FreezStateVol;
            sp_ = sp_ + changSnow;
            snowFracFresh2 = F(0.95);
            ic_ = 1; ic_set = true;
            sd_ = P.SnowDensMin;
         }
      }
      if (sp_ > 0) {
         sw_ = sw_ + mw_ind - freezMelt;
         double FWC = WaterHoldCapFrac * sp_;
         if (sw_ >= FWC) {
            double MeltExcess = sw_ - FWC;
            sw_ = FWC;
            if ((sf_ > F(0.9) && is != BLDG) || (is == BLDG)) runoffSnow = runoffSnow + MeltExcess;
            else SnowToSurf = SnowToSurf + MeltExcess * sf_ / (1 - sf_);
         }
      } else if (sp_ < 0) {
         sw_ = sw_ - freezMelt + mw_ind + sp_;
         sp_ = F(0.0);
         snowFracFresh1 = 0;
         snowFracFresh2 = 0;
         sd_ = 0;
         if (sw_ < 0) {
            ev_snow = ev_snow + sw_;
            if (ev_snow < 0) ev_snow = 0;
            changSnow = changSnow + sw_;
            sw_ = 0;
         } else {
            SnowToSurf = SnowToSurf + sw_ * sf_ / (1 - sf_);
            sw_ = 0;
         }
      }
   }
   // snowRem snow.f95:1210-1240 (paved / buildings, last step of an hour flagged in SnowProf_24hr)
   if (rem_time && (armFull || armPart) && sp_ > 0) {
      double lim = (is == PAV) ? P.SnowLimPaved : P.SnowLimBldg;
      if (sp_ > lim) {
         if (COMMIT) w.SnowRemoval[is] = (sp_ - lim) * sfr[is] * sf_;
         sp_ = lim;
      }
   }

   // liquid-water balance of the snow-free part (only inside the partial-cover arm, snow.f95:869-958;
   // addImpervious / addVeg are zero, driver.f95:478-486)
   if (armPart) {
      if (is == PAV || is == BLDG) {
         if (Precip > IPT) {
            runoff = runoff + (Precip + SnowToSurf + AddWater[is] - IPT);
            chang = IPT - (drain[is] + ev + FreezState);
         } else {
            chang = Precip + SnowToSurf + AddWater[is] - (drain[is] + ev + FreezState);
         }
         st_ = st_ + chang;
         runoff = runoff + drain[is] * AddWaterRunoff[is];
         if (st_ < F(0.0)) {
            SurplusEvap[is] = fabs(st_);
            ev = ev - SurplusEvap[is];
            st_ = F(0.0);
         }
      } else {
         ev = ev + EvPart;
         if (vegfraction > 0) {
            if (Precip > IPT) {
               runoff = runoff + (Precip + SnowToSurf + AddWater[is] - IPT);
               chang = IPT - (drain[is] + ev + FreezState);
            } else {
               chang = Precip + SnowToSurf + AddWater[is] - (drain[is] + ev + FreezState);
            }
         } else {
            chang = Precip + SnowToSurf + AddWater[is] - (drain[is] + ev + FreezState);
         }
         st_ = st_ + chang;
         if (Temp_C > 0) so_ = so_ + drain[is] * AddWaterRunoff[is] * (1 - sf_);
         else runoff = runoff + drain[is] * AddWaterRunoff[is];
         if (st_ < F(0.0)) {
            if ((so_ + st_) >= 0 && Temp_C > 0) {
               so_ = so_ + st_ * (1 - sf_);
               st_ = F(0.0);
            } else {
               chang = chang + st_;
               ev = ev + st_;
               st_ = F(0.0);
            }
         }
         if (so_ > P.SoilStoreCap[is]) {
            runoffTest = runoffTest + (so_ - P.SoilStoreCap[is]);
            so_ = P.SoilStoreCap[is];
         } else if (so_ < 0) {
            so_ = 0;
         }
      }
   }

   const double sfmax = SU_MAX(sf_, snowFracOld);
   if (snowFracFresh2 > 0) {
      surf_chang_tot = (st_ - stateOld[is]) * sfr[is] * (1 - sf_) - Precip * sfr[is] * (1 - snowFracFresh2);
      chSnow_tot = ((sp_ + sw_) - snowTotInit) * sfr[is] * (1 - sf_) - Precip * sfr[is] * snowFracFresh2;
   } else {
      surf_chang_tot = (st_ - stateOld[is]) * sfr[is] * (1 - sf_);
      chSnow_tot = ((sp_ + sw_) - snowTotInit) * sfr[is] * sfmax;
   }
   ev_tot = ev * sfr[is] * (1 - sf_) + ev_snow * sfr[is] * sfmax;
   if (is == BLDG || is == PAV) qe_tot = ev_snow * tlv_sub * sfr[is] * sf_ + ev * tlv * sfr[is] * (1 - sf_);
   else qe_tot = ev_snow * tlv_sub * sfr[is] * sfmax + ev * tlv * sfr[is] * (1 - sf_);

   if (COMMIT) {
      fl.runoffPipes = fl.runoffPipes + runoffSnow * sfr[is] * sfmax + runoff * sfr[is] * (1 - sf_) + runoffTest * sfr[is];
      updateFlood(is, runoff, sfr, C.PipeCapacity, C.RunoffToWater, fl);
   }
   runoff_tot = runoffSnow * sfr[is] * sfmax + runoff * sfr[is] * (1 - sf_) + runoffTest * sfr[is];

   if (COMMIT) {
      if (sd_ != 0) w.SnowDepth[is] = sp_ * SU_WATERDENS / sd_;
      w.swe = w.swe + sp_ * sfr[is] * sfmax;
      w.mwstore = w.mwstore + sw_ * sfr[is] * sfmax;
   }

   if (COMMIT) {   // (the surface's new snow fraction: state only)
      if (snowFracFresh1 > 0) sf_ = snowFracFresh1;
      if (snowFracFresh2 > 0) sf_ = snowFracFresh2;
      if (sp_ > 0 && mw_ind > 0) {
         sf_ = SnowDepletionCurve(is, sp_, P.SnowPackLimit[is]);
         if (sf_ < F(0.001)) sf_ = F(0.001);
      } else if (sp_ == 0) {
         sf_ = 0;
      }
   }
   SU_SNOW_COMMIT();
#undef SU_SNOW_COMMIT
}
