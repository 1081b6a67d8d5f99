// dev_surface.cuh -- storage heat (OHM), LUMPS first guess, Penman-Monteith evaporation, surface and
// soil water balance, anthropogenic heat.  Device restatement of suews_phys_ohm.f95,
// suews_phys_lumps.f95, suews_phys_evap.f95, suews_phys_waterdist.f95, suews_phys_anemsn.f95 and the
// profile lookups of suews_ctrl_input.f95:743-826.
#pragma once
#include "dev_atmos.cuh"

// ------------------------------------------------------------------ OHM (ohm.f95)
// ohm.f95:271-305
SU_INL void OHM_dqndt_cal_X(int dt, int dt_since_start, double qn1_av_prev, double qn1, double dqndt_prev,
                            double &qn1_av_next, double &dqndt_next)
{
   const double dt0_thresh = 3600, window_hr = 2;
   int dt0 = (dt_since_start < dt0_thresh) ? dt_since_start + dt : (int)dt0_thresh;
   double qn1_av_0 = qn1_av_prev - dqndt_prev * (window_hr - (double)((float)dt / 3600.f));   // dt/3600. is REAL(4)
   qn1_av_next = (qn1_av_prev * (dt0 - dt) + qn1 * dt) / (dt0);
   dqndt_next = (qn1_av_next - qn1_av_0) / window_hr;
}

// ohm.f95:12-199 with OHM_coef_cal :202-268 and OHM_QS_cal :347-356
SU_DEV void OHM(SuDiag &dg, const SuewsClassParams &P, int SnowUse, const double *sfr, double qn1, double qn1_av_prev,
                double dqndt_prev, double &qn1_av_next, double &dqndt_next, double qn1_S, double qn1_s_av_prev,
                double dqnsdt_prev, double &qn1_s_av_next, double &dqnsdt_next, int dt_since_start, double Tair_mav_5d,
                const double *soilstore_id, const double *state_id, const double *SnowFrac, double &a1, double &a2,
                double &a3, double &qs, double *deltaQi)
{
   a1 = 0; a2 = 0; a3 = 0;   // (SnowUse: the caller's value -- a compile-time constant in the specialised kernel instances)
   for (int is = 0; is < NSURF; is++) {
      double surfrac = sfr[is];
      int ii = (Tair_mav_5d >= P.OHM_threshSW[is]) ? 0 : 2;
      int i;
      if (state_id[is] > 0) {
         i = ii + 1;
      } else {
         i = ii + 2;
         if (is > BLDG && is != WATER) {
            if (soilstore_id[is] / P.SoilStoreCap[is] > P.OHM_threshWD[is]) i = ii + 1;
         }
      }
      if (SnowUse == 1 && is != BLDG && is != WATER) surfrac = surfrac * (1 - SnowFrac[is]);
      a1 = a1 + surfrac * SU_OHMCOEF(P.OHM_coef.v, is, i - 1, 0);
      a2 = a2 + surfrac * SU_OHMCOEF(P.OHM_coef.v, is, i - 1, 1);
      a3 = a3 + surfrac * SU_OHMCOEF(P.OHM_coef.v, is, i - 1, 2);
   }
   qs = -999;
   if (qn1 > -999) {
      OHM_dqndt_cal_X(P.tstep, dt_since_start, qn1_av_prev, qn1, dqndt_prev, qn1_av_next, dqndt_next);
      qs = qn1 * a1 + dqndt_next * a2 + a3;
   } else {
      su_fatal(dg, 21);
   }
   if (SnowUse == 1) {
      double dq0 = -999;
      if (qn1_S > -999) {
         OHM_dqndt_cal_X(P.tstep, dt_since_start, qn1_s_av_prev, qn1_S, dqnsdt_prev, qn1_s_av_next, dqnsdt_next);
         dq0 = qn1_S * SU_OHMCOEF(P.OHM_coef.v, NSURF, 2, 0) + dqnsdt_next * SU_OHMCOEF(P.OHM_coef.v, NSURF, 2, 1) +
               SU_OHMCOEF(P.OHM_coef.v, NSURF, 2, 2);
      } else {
         su_fatal(dg, 21);
      }
      for (int is = 0; is < NSURF; is++) deltaQi[is] = dq0;
   }
}

// ------------------------------------------------------------------ LUMPS (lumps.f95:5-174)
struct SuLumps {
   double QH_LUMPS, QE_LUMPS, psyc_hPa, s_hPa, sIce_hpa, VegPhenLumps;
};
SU_DEV void LUMPS_cal_QHQE(SuDiag &dg, const SuewsClassParams &P, const double *sfr, double qn1, double qf, double qs,
                           double Qm, double Temp_C, double Veg_Fr, double avcp, double Press_hPa, double lv_J_kg,
                           const double *LAI_id, SuLumps &o)
{
   o.VegPhenLumps = 0;
   o.s_hPa = slope_svp(Temp_C);
   o.psyc_hPa = psyc_const(dg, avcp, Press_hPa, lv_J_kg);
   double psyc_s = o.psyc_hPa / o.s_hPa;
   if (P.snowUse == 1) {
      o.sIce_hpa = (Temp_C <= 0) ? slopeIce_svp(Temp_C) : slope_svp(Temp_C);
      psyc_s = o.psyc_hPa / o.sIce_hpa;
   }
   double VegPhen = sfr[2] * LAI_id[0] + sfr[3] * LAI_id[1] + sfr[4] * LAI_id[2];
   double VegMax = sfr[2] * P.LAIMax[0] + sfr[3] * P.LAIMax[1] + sfr[4] * P.LAIMax[2];
   double Veg_Fr_temp;
   if (VegMax <= F(0.01000)) {
      Veg_Fr_temp = 0;
   } else {
      o.VegPhenLumps = (VegPhen) / (VegMax);
      Veg_Fr_temp = Veg_Fr * o.VegPhenLumps;
   }
   double alpha_sl = F(0.6), alpha_in = F(0.2), beta, alpha_qhqe;
   if (Veg_Fr_temp > F(0.9000)) {
      beta = (20 - 3) * Veg_Fr_temp + 3;
      alpha_qhqe = Veg_Fr_temp * F(0.8) + F(0.2);
   } else {
      beta = 3;
      if (P.veg_type == 1) { alpha_sl = F(0.686); alpha_in = F(0.189); }
      else if (P.veg_type == 2) { alpha_sl = F(0.610); alpha_in = F(0.222); }
      alpha_qhqe = Veg_Fr_temp * alpha_sl + alpha_in;
   }
   o.QH_LUMPS = ((1 - alpha_qhqe) + psyc_s) / (1 + psyc_s) * (qn1 + qf - qs - Qm) - beta;
   if (o.QH_LUMPS == -999) o.QH_LUMPS = qn1 * F(0.2);
   o.QE_LUMPS = (alpha_qhqe / (1 + psyc_s) * (qn1 + qf - qs - Qm)) + beta;
   // the LUMPS rain bucket (lumps.f95:160-171) only feeds locals and is dropped
}

// ------------------------------------------------------------------ evaporation (evap.f95:5-124)
SU_DEV void cal_evap(int EvapMethod, double state_is, double WetThresh_is, double capStore_is, double vpd_hPa,
                     double avdens, double avcp, double qn_e, double s_hPa, double psyc_hPa, double ResistSurf, double RA,
                     double rb, double tlv, double &ev, double &qe)
{
   double numPM = s_hPa * qn_e + su_div0(vpd_hPa * avdens * avcp, RA);
   if (state_is <= F(0.001)) {
      qe = numPM / (s_hPa + psyc_hPa * (1 + ResistSurf / RA));
      ev = qe / tlv;
   } else if (EvapMethod == 2) {
      double rbsg = rb * (s_hPa / psyc_hPa + 1);
      double rsrbsg = ResistSurf + rbsg;
      double W;
      if (state_is >= WetThresh_is || ResistSurf < 25) {
         W = 1;
      } else {
         double r = (ResistSurf / RA) * (RA - rb) / rsrbsg;
         W = (r - 1) / (r - (WetThresh_is / state_is));
      }
      double rss = (1 / ((W / rbsg) + ((1 - W) / rsrbsg))) - rbsg;
      qe = numPM / (s_hPa + psyc_hPa * (1 + rss / RA));
      ev = qe / tlv;
   } else if (EvapMethod == 1) {
      qe = numPM / (s_hPa + psyc_hPa);
      ev = qe / tlv;
      double x = (state_is > capStore_is) ? 1.0 : state_is / capStore_is;
      ev = ev * x;
      qe = ev * tlv;
   }
}

// cal_evap evaluated for the seven surfaces of one trip.  Everything but state, WetThresh and the store
// capacity is the same for all surfaces, so the shared sub-expressions (numPM, the dry-surface QE, rbsg,
// rsrbsg, r, the fully-wet QE) are formed once per trip -- by the same expressions, hence the same bits --
// instead of once per surface: 21 out-of-line calls and ~230 divisions per cell-step become ~20 divisions.
struct SuEvapTrip {
   double numPM, s_hPa, psyc_hPa, ResistSurf, RA, tlv;
   double qe_dry, qe_wet1, rbsg, rsrbsg, r, qe_m1;
   int EvapMethod;
   bool have_dry, have_m2, have_wet1;
};
SU_INL void evap_trip_init(SuEvapTrip &e, int EvapMethod, double vpd_hPa, double avdens, double avcp, double qn_e,
                           double s_hPa, double psyc_hPa, double ResistSurf, double RA, double rb, double tlv)
{
   e.numPM = s_hPa * qn_e + su_div0(vpd_hPa * avdens * avcp, RA);
   e.s_hPa = s_hPa; e.psyc_hPa = psyc_hPa; e.ResistSurf = ResistSurf; e.RA = RA; e.tlv = tlv;
   e.EvapMethod = EvapMethod;
   e.have_dry = false; e.have_m2 = false; e.have_wet1 = false;
   e.qe_dry = 0; e.qe_wet1 = 0; e.rbsg = rb; e.rsrbsg = 0; e.r = 0;   // rbsg holds rb until the method-2 terms are formed
   e.qe_m1 = (EvapMethod == 1) ? e.numPM / (s_hPa + psyc_hPa) : 0.0;
}
// qe of one surface (and ev when WANT_EV); like evap.f95, qe/ev are left untouched for an unknown EvapMethod
template <bool WANT_EV>
SU_INL void evap_surface(SuEvapTrip &e, double state_is, double WetThresh_is, double capStore_is, double &ev, double &qe)
{
   if (state_is <= F(0.001)) {
      if (!e.have_dry) {
         e.qe_dry = e.numPM / (e.s_hPa + e.psyc_hPa * (1 + e.ResistSurf / e.RA));
         e.have_dry = true;
      }
      qe = e.qe_dry;
      if (WANT_EV) ev = qe / e.tlv;
   } else if (e.EvapMethod == 2) {
      if (!e.have_m2) {
         const double rb = e.rbsg;
         e.rbsg = rb * (e.s_hPa / e.psyc_hPa + 1);
         e.rsrbsg = e.ResistSurf + e.rbsg;
         e.r = (e.ResistSurf / e.RA) * (e.RA - rb) / e.rsrbsg;
         e.have_m2 = true;
      }
      if (state_is >= WetThresh_is || e.ResistSurf < 25) {
         if (!e.have_wet1) {
            const double W = 1;
            double rss = (1 / ((W / e.rbsg) + su_div0(1 - W, e.rsrbsg))) - e.rbsg;
            e.qe_wet1 = e.numPM / (e.s_hPa + e.psyc_hPa * (1 + su_div0(rss, e.RA)));
            e.have_wet1 = true;
         }
         qe = e.qe_wet1;
      } else {
         double W = (e.r - 1) / (e.r - (WetThresh_is / state_is));
         double rss = (1 / ((W / e.rbsg) + ((1 - W) / e.rsrbsg))) - e.rbsg;
         qe = e.numPM / (e.s_hPa + e.psyc_hPa * (1 + rss / e.RA));
      }
      if (WANT_EV) ev = qe / e.tlv;
   } else if (e.EvapMethod == 1) {
      double ev1 = e.qe_m1 / e.tlv;
      double x = (state_is > capStore_is) ? 1.0 : state_is / capStore_is;
      ev1 = ev1 * x;
      qe = ev1 * e.tlv;
      ev = ev1;
   }
}

// ------------------------------------------------------------------ biogenic CO2 (biogenco2.f95:33-231)
// Photosynthesis by the four model families selected by the decade of EmissionsMethod (11-16 rectangular
// hyperbola, 21-26 non-rectangular hyperbola, 31-36 bulk non-rectangular, 41-46 conductance based) and
// temperature-driven respiration.  Returns Fc_biogen; Fc_photo / Fc_respi through the references.
SU_COLD double CO2_biogen(const SuewsClassParams &P, const double *sfr, const double *SnowFrac, const double *LAI_id,
                         double avkdn, double gfunc, double gfunc2, double t2, double Temp_C, double &Fc_photo,
                         double &Fc_respi)
{
   const int EM = P.EmissionsMethod, gsmodel = P.gsModel;
   const double *al = P.alpha_bioCO2.v, *be = P.beta_bioCO2.v, *th = P.theta_bioCO2.v;
   const double PAR = F(4.6) * F(0.46) * avkdn;
   const double VegFracSum = sfr[CONIF] + sfr[DECID] + sfr[GRASS];
   double avf[3], avf0[3];
   for (int iv = 0; iv < 3; iv++) {
      avf[iv] = (sfr[iv + 2] * (1 - SnowFrac[iv + 2])) * (LAI_id[iv] / P.LAIMax[iv]);
      avf0[iv] = (sfr[iv + 2] * (1 - SnowFrac[iv + 2])) * LAI_id[iv];
   }
   double photo = Fc_photo;   // left as it was for an unknown family, as the reference does
   if (EM >= 11 && EM <= 16) {
      photo = 0;
      for (int iv = 0; iv < 3; iv++) {
         double q = -be[iv] * al[iv] * PAR / (al[iv] * PAR + be[iv]);
         photo = photo + q * avf[iv];
      }
   } else if (EM >= 21 && EM <= 26) {
      double Pho = 0;
      for (int iv = 0; iv < 3; iv++) {
         double q = al[iv] * PAR + be[iv];
         q = -(1 / (2 * th[iv]) * (al[iv] * PAR + be[iv] - sqrt(q * q - 4 * al[iv] * be[iv] * th[iv] * PAR)));
         Pho = Pho + q * avf[iv];
      }
      photo = Pho;
   } else if (EM >= 31 && EM <= 36) {
      const double s4 = (sfr[CONIF] + sfr[DECID] + sfr[GRASS] + sfr[BSOIL]);
      double a2, b2, Pho, q;
      if (al[0] == al[1] && al[0] == al[2] && be[0] == be[1] && be[0] == be[2] && th[0] == th[1] && th[0] == th[2]) {
         a2 = al[0] + P.alpha_enh_bioCO2[0] * s4;
         b2 = -be[0] + P.beta_enh_bioCO2[0] * s4;
         q = a2 * PAR + b2;
         Pho = -(1 / (2 * th[0]) * (a2 * PAR + b2 - sqrt(q * q - 4 * a2 * b2 * th[0] * PAR)));
      } else {
         a2 = (al[0] * sfr[CONIF] / VegFracSum + al[1] * sfr[DECID] / VegFracSum + al[2] * sfr[GRASS] / VegFracSum) /
              (al[0] + al[1] + al[2]);
         b2 = (be[0] * sfr[CONIF] / VegFracSum + be[1] * sfr[DECID] / VegFracSum + be[2] * sfr[GRASS] / VegFracSum) /
              (be[0] + be[1] + be[2]);
         double t2v = (th[0] * sfr[CONIF] / VegFracSum + th[1] * sfr[DECID] / VegFracSum + th[2] * sfr[GRASS] / VegFracSum) /
                      (th[0] + th[1] + th[2]);
         a2 = a2 + P.alpha_enh_bioCO2[0] * s4;
         b2 = -b2 + P.beta_enh_bioCO2[0] * s4;
         q = a2 * PAR + b2;
         Pho = -(1 / (2 * t2v) * (a2 * PAR + b2 - sqrt(q * q - 4 * a2 * b2 * t2v * PAR)));
      }
      photo = Pho * avf[0] + Pho * avf[1] + Pho * avf[2];
   } else if (EM >= 41 && EM <= 46) {
      photo = 0;
      for (int iv = 0; iv < 3; iv++) photo = photo + avf0[iv] * be[iv];
      if (gsmodel == 1 || gsmodel == 2) photo = -photo * gfunc;
      else if (gsmodel == 3 || gsmodel == 4) photo = -photo * gfunc2;
   }
   double Res = F(0.0), Res_surf = F(0.0);
   if (VegFracSum > F(0.01)) {
      for (int iv = 0; iv < 3; iv++) {
         if (sfr[2 + iv] > F(0.005)) {
            if (gsmodel == 1 || gsmodel == 2) Res_surf = SU_MAX(P.min_res_bioCO2[iv], P.resp_a[iv] * exp(P.resp_b[iv] * Temp_C));
            else if (gsmodel == 3 || gsmodel == 4) Res_surf = SU_MAX(P.min_res_bioCO2[iv], P.resp_a[iv] * exp(P.resp_b[iv] * t2));
            Res = Res + Res_surf * sfr[2 + iv] / VegFracSum;
         }
      }
   }
   Fc_photo = photo;
   Fc_respi = Res * (sfr[CONIF] + sfr[DECID] + sfr[GRASS] + sfr[BSOIL]);
   return Fc_photo + Fc_respi;
}

// ------------------------------------------------------------------ water balance (waterdist.f95)
// waterdist.f95:20-89
SU_DEV double drainage(SuDiag &dg, double state_is, double StorCap, double DrainEq, double DrainCoef1, double DrainCoef2,
                       double nsh_real)
{
   double drain_is = 0;
   if (state_is < F(0.000000001)) {
      drain_is = F(0.0);
   } else {
      int eq = (int)DrainEq;
      if (eq == 1) {
         drain_is = (state_is < StorCap) ? 0.0 : (DrainCoef1 * pow(state_is - StorCap, DrainCoef2)) / nsh_real;
      } else if (eq == 2) {
         drain_is = (DrainCoef1 * (exp(DrainCoef2 * state_is) - 1)) / nsh_real;
      } else if (eq == 3) {
         drain_is = (DrainCoef1 * (pow(state_is, DrainCoef2))) / nsh_real;
      }
      if (drain_is > state_is) {
         su_warn(dg, 61);
         drain_is = state_is;
      } else if (drain_is < F(0.0001)) {
         drain_is = 0;
      }
   }
   return drain_is;
}

struct SuFlood {
   double runoffAGimpervious, surplusWaterBody, runoffAGveg, runoffPipes, runoffWaterBody;
};

// waterdist.f95:405-451
SU_INL void updateFlood(int is, double runoff_is, const double *sfr, double PipeCapacity, double RunoffToWater,
                        SuFlood &fl)
{
   fl.runoffPipes = fl.runoffPipes + (runoff_is * sfr[is]);
   if (fl.runoffPipes > PipeCapacity) {
      double ex = (fl.runoffPipes - PipeCapacity);
      if (is == PAV || is == BLDG) {
         if (sfr[WATER] > F(0.0000001)) {
            fl.runoffAGimpervious = fl.runoffAGimpervious + ex * (1 - RunoffToWater);
            fl.surplusWaterBody = fl.surplusWaterBody + ex * RunoffToWater;
         } else {
            fl.runoffAGimpervious = fl.runoffAGimpervious + ex;
         }
      } else if (is >= CONIF && is <= BSOIL) {
         if (sfr[WATER] > F(0.0000001)) {
            fl.runoffAGveg = fl.runoffAGveg + ex * (1 - RunoffToWater);
            fl.surplusWaterBody = fl.surplusWaterBody + ex * RunoffToWater;
         } else {
            fl.runoffAGveg = fl.runoffAGveg + ex;
         }
      }
      fl.runoffPipes = PipeCapacity;
   }
}

// waterdist.f95:92-402 for one surface.  addImpervious/addVeg/addWaterBody are hard-wired to zero by the
// reference (driver.f95:478-486: no grid-to-grid connection); p_mm + 0/sfr == p_mm, so the terms are dropped.
SU_DEV void cal_water_storage(SuDiag &dg, int is, const SuewsClassParams &P, const SuewsCellParams &C, double pin,
                              double WU_is, double drain_is, double AddWater_is, double nsh_real, double stateOld_is,
                              double AddWaterRunoff_is, double PervFraction, SuFlood &fl, double &ev,
                              double *soilstore_id, double *SurplusEvap, double &runoff_is, double &state_is)
{
   const double IPT = 10 / nsh_real;   // IPThreshold_mmhr/nsh_real
   const double *sfr = C.sfr.v;
   double chang;
   runoff_is = 0;
   double p_mm = pin + WU_is;
   p_mm = p_mm + AddWater_is;
   if (is == PAV || is == BLDG) {
      chang = p_mm - (drain_is + ev);
      if (p_mm > IPT) {
         runoff_is = runoff_is + (p_mm - IPT);
         chang = IPT - (drain_is + ev);
      }
      state_is = stateOld_is + chang;
      if (state_is < F(0.0)) {
         SurplusEvap[is] = fabs(state_is);
         ev = ev - SurplusEvap[is];
         state_is = F(0.0);
      }
      runoff_is = runoff_is + drain_is * AddWaterRunoff_is;
   } else if (is >= CONIF && is <= BSOIL) {
      double EvPart = (PervFraction != 0)
                         ? SurplusEvap[PAV] * (sfr[PAV] / PervFraction) + SurplusEvap[BLDG] * (sfr[BLDG] / PervFraction)
                         : 0.0;
      ev = ev + EvPart;
      chang = p_mm - (drain_is + ev);
      if (p_mm > IPT) {
         runoff_is = runoff_is + (p_mm - IPT);
         chang = IPT - (drain_is + ev);
      }
      state_is = stateOld_is + chang;
      if (state_is < F(0.0)) {
         if ((soilstore_id[is] + state_is) >= 0) {
            soilstore_id[is] = soilstore_id[is] + state_is;
            state_is = F(0.0);
         } else {
            ev = ev - fabs(state_is);
            state_is = F(0.0);
         }
      }
      soilstore_id[is] = soilstore_id[is] + drain_is * AddWaterRunoff_is;
      if (soilstore_id[is] > P.SoilStoreCap[is]) {
         runoff_is = runoff_is + (soilstore_id[is] - P.SoilStoreCap[is]);
         soilstore_id[is] = P.SoilStoreCap[is];
      } else if (soilstore_id[is] < 0) {
         su_fatal(dg, 62);
      }
   } else if (is == WATER) {
      if (sfr[WATER] != 0) {
         chang = p_mm + su_div0(C.FlowChange, nsh_real) - ev;
         state_is = stateOld_is + chang;
         if (state_is < F(0.0)) {
            ev = ev - fabs(state_is);
            state_is = F(0.0);
         }
         if (state_is > P.StateLimit[WATER]) {
            runoff_is = runoff_is + (state_is - P.StateLimit[WATER]);
            state_is = P.StateLimit[WATER];
            fl.runoffWaterBody = fl.runoffWaterBody + runoff_is * sfr[WATER];
         } else {
            state_is = state_is + fl.surplusWaterBody;
            if (state_is > P.StateLimit[WATER]) {
               fl.runoffWaterBody = fl.runoffWaterBody + (state_is - P.StateLimit[WATER]) * sfr[WATER];
               state_is = P.StateLimit[WATER];
            }
         }
      }
   }
   if (is < WATER) updateFlood(is, runoff_is, sfr, C.PipeCapacity, C.RunoffToWater, fl);
}

// waterdist.f95:454-504
SU_DEV void ReDistributeWater(int snowUse, const double *WaterDist, const double *sfr, const double *Drain,
                              double *AddWaterRunoff, double *AddWater)
{
   for (int ii = 0; ii < NSURF - 1; ii++) AddWaterRunoff[ii] = SU_WDIST(WaterDist, 7, ii);
   AddWaterRunoff[WATER] = 0;
   // same accumulation order as the reference's ii/jj loops; a surface that does not drain adds an exact zero
   // and is skipped, and the running sum of a receiving surface stays in a register
   // Drain[jj] * sfr[jj], the leading product of every term of column jj, once per column (in registers: the loop over jj
   // below is unrolled) instead of once per (ii, jj) pair from the frame
   double ds[NSURF - 1];
#pragma unroll
   for (int jj = 0; jj < NSURF - 1; jj++) ds[jj] = Drain[jj] * sfr[jj];
   for (int ii = 0; ii < NSURF; ii++) {
      if (sfr[ii] != 0 && snowUse == 0) {
         double a = 0;
         const double sfr_ii = sfr[ii];
#pragma unroll
         for (int jj = 0; jj < NSURF - 1; jj++) {
            // (most of the 7 x 6 WaterDist matrix is zero -- a surface hands its drainage to one or two others -- and a
            // zero fraction adds +0 to a sum of non-negative terms: the division is only made for the others; a column that
            // does not drain has ds = 0 (or -0 for a zero fraction, as Drain is never negative) and adds an exact zero too)
            const double wd = SU_WDIST(WaterDist, ii, jj);
            if (ds[jj] != 0 && wd != 0) a = a + (ds[jj] / sfr_ii) * wd;
         }
         AddWater[ii] = a;
      } else {
         AddWater[ii] = 0;
         for (int jj = 0; jj < NSURF - 1; jj++) AddWaterRunoff[jj] = AddWaterRunoff[jj] + SU_WDIST(WaterDist, ii, jj);
      }
   }
}

// waterdist.f95:507-550
SU_DEV void update_SoilMoist(double NonWaterFraction, const double *SoilStoreCap, const double *sfr,
                             const double *soilstore_id, double &SoilMoistCap, double &SoilState, double &vsmd,
                             double &smd)
{
   SoilMoistCap = 0;
   SoilState = 0;
   if (NonWaterFraction != 0) {
      for (int is = 0; is < NSURF - 1; is++) {
         SoilMoistCap = SoilMoistCap + (SoilStoreCap[is] * sfr[is] / NonWaterFraction);
         SoilState = SoilState + (soilstore_id[is] * sfr[is] / NonWaterFraction);
      }
   }
   smd = SoilMoistCap - SoilState;
   vsmd = 0;
   double vf = sfr[CONIF] + sfr[DECID] + sfr[GRASS];
   if (vf != 0)
      for (int is = CONIF; is <= GRASS; is++) vsmd = vsmd + su_div0((SoilStoreCap[is] - soilstore_id[is]) * sfr[is], vf);
}

// waterdist.f95:553-619
// (WANT_OUT = false: only the range check of the soil moisture state, for launches that form no output line)
template <bool WANT_OUT>
SU_DEV void cal_SoilState(SuDiag &dg, int SMDMethod, double xsmd, double NonWaterFraction, double SoilMoistCap,
                          const double *SoilStoreCap, double surf_chang_per_tstep, const double *soilstore_id,
                          const double *soilstoreOld, const double *sfr, double &smd, double *smd_nsurf,
                          double &tot_chang_per_tstep)
{
   double SoilState = 0;
   if (NonWaterFraction != 0) {
      for (int is = 0; is < NSURF - 1; is++) {
         SoilState = SoilState + (soilstore_id[is] * sfr[is] / NonWaterFraction);
         if (SoilState < 0 || SoilState > SoilMoistCap) su_fatal(dg, 62);
      }
   }
   if (!WANT_OUT) return;
   smd = SoilMoistCap - SoilState;
   for (int is = 0; is < NSURF; is++) smd_nsurf[is] = SoilStoreCap[is] - soilstore_id[is];
   tot_chang_per_tstep = surf_chang_per_tstep;
   for (int is = 0; is < NSURF - 1; is++)
      tot_chang_per_tstep = tot_chang_per_tstep + ((soilstore_id[is] - soilstoreOld[is]) * sfr[is]);
   if (SMDMethod > 0) {
      for (int is = 0; is < NSURF; is++) smd_nsurf[is] = -999;
      smd = xsmd;
   }
}

// van Genuchten (1980) matric potential and conductivity of one store, waterdist.f95:707-762
SU_HOT void vg_store(double soilstore, double SoilStoreCap, double SoilDepth, double Ksat, double &MatPot, double &Km)
{
   const double alphavG = F(0.0005);
   double capVol = SoilStoreCap / SoilDepth;
   double vol = soilstore / SoilDepth;
   const double B_r = F(0.1);
   if (B_r >= vol) {
      MatPot = 100000;
      Km = 0;
   } else {
      double D = (vol - B_r) / (capVol - B_r);
      if (D > F(0.99999)) D = D - F(0.0001);
      if (D < F(0.00000005)) D = D + F(0.0000001);
      MatPot = sqrt(1 / (D * D) - 1) / alphavG;
      double t = (1 - pow(1 - D * D, F(0.5)));
      Km = Ksat * sqrt(D) * (t * t);
      if (MatPot > 100000) {
         MatPot = 100000;
         Km = 0;
      }
   }
}

// waterdist.f95:622-855: horizontal flow between the 15 pairs of soil stores.  sqrt(sfr*SurfaceArea) is a
// per-cell constant, evaluated per pair by the reference and once per surface here.
SU_BIG2 void cal_HorizontalSoilWater(const SuewsClassParams &P, const double *sfr, double SurfaceArea,
                                      double NonWaterFraction, double tstep_real, double *soilstore_id,
                                      double *runoffSoil, double &runoffSoil_per_tstep)
{
   double rt[NSURF - 1];
   for (int is = 0; is < NSURF - 1; is++) rt[is] = sqrt(sfr[is] * SurfaceArea / 1.0);
   runoffSoil_per_tstep = 0;
   // The reference evaluates the van Genuchten pair of BOTH stores for each of the 15 pairs (30 evaluations); a store's
   // pair is a pure function of its content, which a pair only changes when water actually flows (equal matric
   // potentials -- e.g. two saturated stores -- give dI = 0).  Each store's last evaluation is kept with the content it
   // was made for and reused while the content is bit-identical.
   // The values of the outer store `is` live in registers for its whole row of pairs (it never comes back as a `jj`:
   // jj > is), only the inner store's are read from and written to the arrays: half the array traffic of the pair loop.
   double vg_at[NSURF - 1], vg_mp[NSURF - 1], vg_km[NSURF - 1];
   for (int is = 0; is < NSURF - 1; is++) vg_at[is] = -1.0;   // soil stores are never negative
   for (int is = 0; is < NSURF - 1; is++) {
      const double sfr_is = sfr[is], cap_is = P.SoilStoreCap[is];
      if (sfr_is != 0 && cap_is > 0) {
         double s_is = soilstore_id[is], at_is = vg_at[is], mp_is = vg_mp[is], km_is = vg_km[is], ro_is = runoffSoil[is];
         const double rt_is = rt[is], depth_is = P.SoilDepth[is], ksat_is = P.SatHydraulicConduct[is];
         for (int jj = is + 1; jj < NSURF - 1; jj++) {
            const double sfr_jj = sfr[jj], cap_jj = P.SoilStoreCap[jj];
            if (sfr_jj != 0 && cap_jj > 0) {
               double s_jj = soilstore_id[jj];
               if (at_is != s_is) {
                  vg_store(s_is, cap_is, depth_is, ksat_is, mp_is, km_is);
                  at_is = s_is;
               }
               if (vg_at[jj] != s_jj) {
                  vg_store(s_jj, cap_jj, P.SoilDepth[jj], P.SatHydraulicConduct[jj], vg_mp[jj], vg_km[jj]);
                  vg_at[jj] = s_jj;
               }
               const double MatPot1 = mp_is, Km1 = km_is, MatPot2 = vg_mp[jj], Km2 = vg_km[jj];
               // Equal matric potentials (two saturated stores, two dry ones): the gradient is an exact zero, dI = -0 and the
               // pair exchanges nothing -- both stores, already capped by the water balance or by an earlier pair, keep their
               // bits (x + -0 = x, x - -0 = x); with neither above its capacity the tests below do not fire either.
               if (MatPot1 == MatPot2 && !(s_is > cap_is) && !(s_jj > cap_jj)) continue;
               double Distance = (rt_is + rt[jj]) / 2;
               double KmWeight = (sfr_is * Km1 + sfr_jj * Km2) / (sfr_is + sfr_jj);
               double dI_dt = su_div0(-(KmWeight) * (-MatPot1 + MatPot2), Distance * 1000);
               double dI = dI_dt * tstep_real;
               const double dI_jj = su_div0(dI * sfr_is, sfr_jj);
               if ((s_jj >= dI_jj) && ((s_is + dI) >= 0)) {
                  s_is = s_is + dI;
                  s_jj = s_jj - dI_jj;
               } else if ((s_is + dI) < 0) {
                  s_jj = s_jj + s_is * sfr_is / sfr_jj;
                  s_is = 0;
               } else {
                  s_is = s_is + s_jj * sfr_jj / sfr_is;
                  s_jj = 0;
               }
               if (s_is > cap_is) {
                  ro_is = ro_is + (s_is - cap_is);
                  s_is = cap_is;
               }
               if (s_jj > cap_jj) {
                  runoffSoil[jj] = runoffSoil[jj] + (s_jj - cap_jj);
                  s_jj = cap_jj;
               }
               soilstore_id[jj] = s_jj;
            }
         }
         soilstore_id[is] = s_is;
         runoffSoil[is] = ro_is;
         runoffSoil_per_tstep = runoffSoil_per_tstep + su_div0(ro_is * sfr_is, NonWaterFraction);
      }
   }
}

// ------------------------------------------------------------------ diurnal profiles (ctrl_input.f95:743-826)
// The 24-term SUM the reference redoes on every call is a class constant (prof_sum), computed at
// upload with the same sequential order.
SU_INL double prof_inst(int Hour, int Min, const double *p24)
{
   int j = (Hour + 1 == 24) ? 0 : Hour + 1;
   double deltaProf = su_div0(p24[j] - p24[Hour], 3600.0);
   return p24[Hour] + deltaProf * (Min * 60);
}
SU_INL double prof_mean(int Hour, int Min, const double *p24, double sum24)
{
   int j = (Hour + 1 == 24) ? 0 : Hour + 1;
   double pi_ = 0.0, pj_ = 0.0;
   if (sum24 != 0) {
      double m = sum24 / 24;
      pi_ = p24[Hour] / m;
      pj_ = p24[j] / m;
   }
   double deltaProf = su_div0(pj_ - pi_, 3600.0);
   return pi_ + deltaProf * (Min * 60);
}
SU_INL double prof_sum(int Hour, int Min, const double *p24, double sum24, int dt)
{
   int j = (Hour + 1 == 24) ? 0 : Hour + 1;
   double pi_ = 0.0, pj_ = 0.0;
   if (sum24 != 0) {
      pi_ = p24[Hour] / sum24;
      pj_ = p24[j] / sum24;
   }
   double deltaProf = (pj_ - pi_) / 3600;
   double v = pi_ + deltaProf * (Min * 60);
   return v * dt / 3600;
}
enum { PS_AH = 0, PS_HUM = 1, PS_POP = 2, PS_TRAFF = 3, PS_WUA = 4, PS_WUM = 5 };

// waterdist.f95:858-1124
SU_DEV void cal_WaterUse(const SuewsClassParams &P, const SuewsCellParams &C, double nsh_real, double wu_m3, int dow,
                         const double *HDD_id, const double *WUDay_id, int NSH, int it, int imin, int DLS,
                         double *wu_nsurf, double &wu_int, double &wu_ext)
{
   const double *sfr = C.sfr.v;
   double IrrFrac[NSURF] = {C.IrrFracPaved, C.IrrFracBldgs, C.IrrFracEveTr, C.IrrFracDecTr,
                            C.IrrFracGrass, C.IrrFracBSoil, C.IrrFracWater};
   double OverUse = 0, wu = 0;
   for (int is = 0; is < NSURF; is++) wu_nsurf[is] = 0;
   if (P.WaterUseMethod == 1) {
      double tot = 0;
      for (int is = 0; is < NSURF; is++) tot += IrrFrac[is] * sfr[is] * C.SurfaceArea;
      if (wu_m3 == F(-999.) || wu_m3 == 0) {
         wu = 0;
      } else if (tot > 0) {
         wu = (wu_m3 / tot * 1000);
         for (int is = 0; is < NSURF; is++) wu_nsurf[is] = wu * IrrFrac[is];
         wu = (wu_m3 / C.SurfaceArea * 1000);
      }
   } else if (P.WaterUseMethod == 0) {
      int tstep = (int)(3600 / NSH);
      int ih = it - DLS;
      if (ih < 0) ih = 23;
      int u = (dow == 1 || dow == 7) ? 1 : 0;
      double A = prof_sum(ih, imin, &P.WUProfA_24hr.v[24 * u], P.prof_sum[PS_WUA][u], tstep);
      double flag_WuM = (HDD_id[10] > 2) ? 0.0 : 1.0;
      double M = prof_sum(ih, imin, &P.WUProfM_24hr.v[24 * u], P.prof_sum[PS_WUM][u], tstep);
      double WA[NSURF] = {0, 0, WUDay_id[1], WUDay_id[4], WUDay_id[7], 0, 0};
      double WM[NSURF] = {0, 0, WUDay_id[2], WUDay_id[5], WUDay_id[8], 0, 0};
      for (int is = 0; is < NSURF; is++) wu_nsurf[is] = A * WA[is] + M * WM[is] * flag_WuM;
      for (int is = 0; is < NSURF; is++) wu_nsurf[is] = wu_nsurf[is] * IrrFrac[is];
      for (int is = 0; is < NSURF; is++) wu += wu_nsurf[is] * sfr[is];
   }
   double InternalWaterUse = su_div0(P.InternalWaterUse_h, nsh_real);
   wu_ext = wu - (InternalWaterUse + OverUse);
   if (wu_ext < 0) wu_ext = 0;
   wu_int = wu - wu_ext;
   if (wu_ext != 0 && wu != 0)
      for (int is = 0; is < NSURF; is++) wu_nsurf[is] = wu_nsurf[is] * wu_ext / wu;
}

// ------------------------------------------------------------------ anthropogenic heat (anemsn.f95:29-338)
SU_INL bool em_is(int m, int a, int b)
{
   int r = m % 10, dcd = m / 10;
   return dcd <= 4 && (r == a || r == b);
}
SU_INL bool em_in(int m, int lo, int hi)
{
   int r = m % 10, dcd = m / 10;
   return dcd <= 4 && r >= lo && r <= hi;
}
struct SuAnthro {
   double qf, Fc_anthro, Fc_metab, Fc_traff, Fc_build, Fc_point;
};
SU_DEV void cal_AnthropogenicEmission(SuDiag &dg, const SuewsClassParams &P, const SuewsCellParams &C, int dow, int DLS,
                                      const double *HDD_id, int it, int imin, double QF_obs, double Temp_C, SuAnthro &o)
{
   const int EM = P.EmissionsMethod;
   o.Fc_anthro = 0; o.Fc_metab = 0; o.Fc_traff = 0; o.Fc_build = 0; o.Fc_point = 0;
   if (EM == 0) {
      o.qf = QF_obs;
      return;
   }
   if (!((EM > 0 && EM <= 6) || EM >= 11)) {
      su_fatal(dg, 73);
      return;
   }
   const int u = (dow == 1 || dow == 7) ? 1 : 0;
   double NumCapita[2] = {0, 0};
   if (C.PopDensDaytime[0] >= 0 && C.PopDensNighttime >= 0) NumCapita[0] = (C.PopDensDaytime[0] + C.PopDensNighttime) / 2;
   if (C.PopDensDaytime[1] >= 0 && C.PopDensNighttime >= 0) NumCapita[1] = (C.PopDensDaytime[1] + C.PopDensNighttime) / 2;
   int ih = it - DLS;
   if (ih < 0) ih = 23;
   double PopDorNorT = prof_inst(ih, imin, &P.PopProf_24hr.v[24 * u]);
   double ActDorNorT = prof_inst(ih, imin, &P.HumActivity_24hr.v[24 * u]);
   double TraffDorNorT = prof_mean(ih, imin, &P.TraffProf_24hr.v[24 * u], P.prof_sum[PS_TRAFF][u]);
   double AHDorNorT = prof_mean(ih, imin, &P.AHProf_24hr.v[24 * u], P.prof_sum[PS_AH][u]);
   double DP_x_RhoPop = AHDorNorT * NumCapita[u];
   const double PopN = C.PopDensNighttime, PopD = C.PopDensDaytime[u];
   double QF_metab = (PopN * P.MinQFMetab * ((2 - ActDorNorT) + (2 - PopDorNorT)) / 2 +
                      PopD * P.MaxQFMetab * ((ActDorNorT - 1) + (PopDorNorT - 1)) / 2) / 10000;
   double Fc_metab = (PopN * P.MinFCMetab * ((2 - ActDorNorT) + (2 - PopDorNorT)) / 2 +
                      PopD * P.MaxFCMetab * ((ActDorNorT - 1) + (PopDorNorT - 1)) / 2) / 10000;
   double base = 0, heat = 0, cool = 0, QF_traff = 0;
   if (em_is(EM, 1, 4)) {
      heat = (Temp_C < P.BaseT_Heating[u]) ? (P.AH_SLOPE_Heating[u] * (P.BaseT_Heating[u] - Temp_C)) * AHDorNorT : 0.0;
      base = P.AH_MIN[u] * AHDorNorT;
   } else if (em_is(EM, 2, 5)) {
      base = (P.Qf_A[u]) * DP_x_RhoPop;
      heat = (P.Qf_C[u] * HDD_id[6]) * DP_x_RhoPop;
      cool = (P.Qf_B[u] * HDD_id[7]) * DP_x_RhoPop;
   } else if (em_is(EM, 3, 6)) {
      const double Tavg = HDD_id[9];
      base = P.AH_MIN[u] * AHDorNorT;
      if (Tavg < P.BaseT_Heating[u]) heat = (P.AH_SLOPE_Heating[u] * (P.BaseT_Heating[u] - Tavg)) * AHDorNorT;
      else if (Tavg > P.BaseT_Cooling[u]) cool = (P.AH_SLOPE_Cooling[u] * (Tavg - P.BaseT_Cooling[u])) * AHDorNorT;
   }
   double QF_SAHP = base + heat + cool;
   double Fc_traff = 0, Fc_build = 0, Fc_point = 0;
   if (em_in(EM, 1, 3)) {
      if (!((base - QF_metab) > 0)) su_warn(dg, 69);
      Fc_build = heat * P.FrFossilFuel_Heat * P.EF_umolCO2perJ;
      if ((base - QF_metab) > 0) Fc_build = Fc_build + base * C.QF0_BEU[u] * P.FrFossilFuel_NonHeat * P.EF_umolCO2perJ;
      QF_traff = base * (F(1.0) - C.QF0_BEU[u]) - QF_metab;
      Fc_traff = su_div0(su_div0(QF_traff, P.EnEF_v_Jkm) * P.FcEF_v_kgkm[u] * F(1e3) * F(1e6), 44.0);
      Fc_point = (P.CO2PointSource > 0) ? P.CO2PointSource * F(1e3) * F(1e6) / (12 * 60 * 60 * 24 * C.SurfaceArea) : 0.0;
   } else if (em_in(EM, 4, 6)) {
      if (P.TrafficUnits == 1) {
         QF_traff = C.TrafficRate[u] / (60 * 60 * 24) * P.EnEF_v_Jkm * TraffDorNorT;
         Fc_traff = C.TrafficRate[u] / (60 * 60 * 24) * P.FcEF_v_kgkm[u] * F(1e3) * F(1e6) / 44 * TraffDorNorT;
      } else if (P.TrafficUnits == 2) {
         double dpt = TraffDorNorT * NumCapita[u] / 10000;
         QF_traff = C.TrafficRate[u] / (60 * 60 * 24) * P.EnEF_v_Jkm * dpt;
         Fc_traff = C.TrafficRate[u] / (60 * 60 * 24) * P.FcEF_v_kgkm[u] * F(1e3) * F(1e6) / 44 * dpt;
      } else {
         su_fatal(dg, 75);
      }
      double QF_build = ((base * C.QF0_BEU[u] + heat + cool) / DP_x_RhoPop) * (PopN * (2 - PopDorNorT) + PopD * (PopDorNorT - 1));
      Fc_build = heat * P.FrFossilFuel_Heat * P.EF_umolCO2perJ;
      Fc_build = Fc_build + base * C.QF0_BEU[u] * P.FrFossilFuel_NonHeat * P.EF_umolCO2perJ;
      Fc_point = (P.CO2PointSource > 0) ? P.CO2PointSource * F(1e3) * F(1e6) / (12 * 60 * 60 * 24 * C.SurfaceArea) : 0.0;
      QF_SAHP = QF_metab + QF_traff + QF_build;
   }
   o.qf = QF_SAHP;   // driver.f95:1143: qf = QF_SAHP for EmissionsMethod >= 1
   if (!(EM >= 0 && EM <= 6)) {   // driver.f95:1145-1151 zeroes the CO2 terms for methods 0-6
      o.Fc_metab = Fc_metab; o.Fc_traff = Fc_traff; o.Fc_build = Fc_build; o.Fc_point = Fc_point;
      o.Fc_anthro = Fc_metab + Fc_traff + Fc_build + Fc_point;
   }
}
