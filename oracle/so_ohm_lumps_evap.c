/* ORACLE (test infrastructure only -- see suews_oracle.h).
 * Restates SUEWS-SourceCode/suews_phys_ohm.f95, suews_phys_lumps.f95 and
 * suews_phys_evap.f95. */
#include "suews_oracle.h"
#include <math.h>

/* ohm.f95:202-268 */
static void OHM_coef_cal(const double sfr[NSURF], double Tair_mav_5d, const double *OHM_coef,
                         const double OHM_threshSW[NSURF + 1], const double OHM_threshWD[NSURF + 1],
                         const double soilstore_id[NSURF], const double SoilStoreCap[NSURF],
                         const double state_id[NSURF], int SnowUse, const double SnowFrac[NSURF], double *a1,
                         double *a2, double *a3)
{
   double surfrac;
   int i, ii, is; /* i, ii are the 1-based Fortran season/wetness column */
   *a1 = 0; *a2 = 0; *a3 = 0;
   for (is = 0; is < NSURF; is++) {
      surfrac = sfr[is];
      if (Tair_mav_5d >= OHM_threshSW[is]) ii = 0; else ii = 2;
      if (state_id[is] > 0) {
         i = ii + 1;
      } else {
         i = ii + 2;
         if (is > BLDG && is != WATER) {
            if (soilstore_id[is] / SoilStoreCap[is] > OHM_threshWD[is]) i = ii + 1;
         }
      }
      if (SnowUse == 1 && is != BLDG && is != WATER) surfrac = surfrac * (1 - SnowFrac[is]);
      *a1 = *a1 + surfrac * OHMCOEF(OHM_coef, is, i - 1, 0);
      *a2 = *a2 + surfrac * OHMCOEF(OHM_coef, is, i - 1, 1);
      *a3 = *a3 + surfrac * OHMCOEF(OHM_coef, is, i - 1, 2);
   }
}

/* ohm.f95:271-305 */
void so_OHM_dqndt_cal_X(int dt, int dt_since_start, double qn1_av_prev, double qn1, double dqndt_prev,
                            double *qn1_av_next, double *dqndt_next)
{
   const double dt0_thresh = 3600, window_hr = 2;
   int dt0;
   double qn1_av_0;
   if (dt_since_start < dt0_thresh) dt0 = dt_since_start + dt; else dt0 = (int)dt0_thresh;
   /* dt/3600. is a REAL(4) quotient */
   qn1_av_0 = qn1_av_prev - dqndt_prev * (window_hr - (double)((float)dt / 3600.f));
   *qn1_av_next = (qn1_av_prev * (dt0 - dt) + qn1 * dt) / (dt0);
   *dqndt_next = (*qn1_av_next - qn1_av_0) / window_hr;
}

/* ohm.f95:347-356 */
void so_OHM_QS_cal(double qn1, double dqndt, double a1, double a2, double a3, double *qs)
{
   *qs = qn1 * a1 + dqndt * a2 + a3;
}

/* ohm.f95:12-199 */
void so_OHM(SoCtx *cx, double qn1, double qn1_av_prev, double dqndt_prev, double *qn1_av_next, double *dqndt_next,
            double qn1_S, double qn1_s_av_prev, double dqnsdt_prev, double *qn1_s_av_next, double *dqnsdt_next,
            int tstep, int dt_since_start, const double sfr[NSURF], double Tair_mav_5d, const double *OHM_coef,
            const double OHM_threshSW[NSURF + 1], const double OHM_threshWD[NSURF + 1],
            const double soilstore_id[NSURF], const double SoilStoreCap[NSURF], const double state_id[NSURF],
            int SnowUse, const double SnowFrac[NSURF], int DiagQS, double *a1, double *a2, double *a3, double *qs,
            double deltaQi[NSURF])
{
   double deltaQi0;
   int is;
   (void)DiagQS;
   OHM_coef_cal(sfr, Tair_mav_5d, OHM_coef, OHM_threshSW, OHM_threshWD, soilstore_id, SoilStoreCap, state_id, SnowUse,
                SnowFrac, a1, a2, a3);
   *qs = -999;
   if (qn1 > -999) {
      so_OHM_dqndt_cal_X(tstep, dt_since_start, qn1_av_prev, qn1, dqndt_prev, qn1_av_next, dqndt_next);
      so_OHM_QS_cal(qn1, *dqndt_next, *a1, *a2, *a3, qs);
   } else {
      so_error_hint(cx, 21, "In SUEWS_OHM.f95: bad value for qn1 found during qs calculation.", qn1, -55.55, -55);
   }
   if (SnowUse == 1) {
      for (is = 0; is < NSURF; is++) deltaQi[is] = -999;
      if (qn1_S > -999) {
         so_OHM_dqndt_cal_X(tstep, dt_since_start, qn1_s_av_prev, qn1_S, dqnsdt_prev, qn1_s_av_next, dqnsdt_next);
         so_OHM_QS_cal(qn1_S, *dqnsdt_next, OHMCOEF(OHM_coef, NSURF, 2, 0), OHMCOEF(OHM_coef, NSURF, 2, 1),
                    OHMCOEF(OHM_coef, NSURF, 2, 2), &deltaQi0);
         for (is = 0; is < NSURF; is++) deltaQi[is] = deltaQi0;
      } else {
         so_error_hint(cx, 21, "In SUEWS_OHM.f95: bad value for qn1(snow) found during qs calculation.", qn1_S,
                       -55.55, -55);
      }
   }
}

/* lumps.f95:5-174 */
void so_LUMPS_cal_QHQE(SoCtx *cx, int veg_type, int snowUse, double qn1, double qf, double qs, double Qm,
                       double Temp_C, double Veg_Fr, double avcp, double Press_hPa, double lv_J_kg,
                       double tstep_real, double DRAINRT, double nsh_real, double Precip, double RainMaxRes,
                       double RAINCOVER, const double sfr[NSURF], const double LAI_id_prev[NVEG],
                       const double LAImax[NVEG], const double LAImin[NVEG], double *QH_LUMPS, double *QE_LUMPS,
                       double *psyc_hPa, double *s_hPa, double *sIce_hpa, double *Veg_Fr_temp, double *VegPhenLumps)
{
   const double NAN_ = -999;
   double sfrVeg[3], VegPhen, VegMax, VegMin, psyc_s, alpha_sl, alpha_in, beta, alpha_qhqe, RAINRES, RainBucket, tlv;
   int i;

   tlv = lv_J_kg / tstep_real;
   *VegPhenLumps = 0;
   RainBucket = F(0.);
   for (i = 0; i < 3; i++) sfrVeg[i] = sfr[IVCONIF + 2 + i];

   *s_hPa = so_slope_svp(Temp_C);
   *psyc_hPa = so_psyc_const(cx, avcp, Press_hPa, lv_J_kg);
   psyc_s = *psyc_hPa / *s_hPa;

   if (snowUse == 1) {
      if (Temp_C <= 0) *sIce_hpa = so_slopeIce_svp(Temp_C); else *sIce_hpa = so_slope_svp(Temp_C);
      psyc_s = *psyc_hPa / *sIce_hpa;
   }

   VegPhen = sfrVeg[0] * LAI_id_prev[0] + sfrVeg[1] * LAI_id_prev[1] + sfrVeg[2] * LAI_id_prev[2];
   VegMax = sfrVeg[0] * LAImax[0] + sfrVeg[1] * LAImax[1] + sfrVeg[2] * LAImax[2];
   VegMin = sfrVeg[0] * LAImin[0] + sfrVeg[1] * LAImin[1] + sfrVeg[2] * LAImin[2];
   (void)VegMin;

   if (VegMax <= F(0.01000)) {
      *Veg_Fr_temp = 0;
   } else {
      *VegPhenLumps = (VegPhen) / (VegMax);
      *Veg_Fr_temp = Veg_Fr * (*VegPhenLumps);
   }

   alpha_sl = F(0.6);
   alpha_in = F(0.2);

   if (*Veg_Fr_temp > F(0.9000)) {
      beta = (20 - 3) * (*Veg_Fr_temp) + 3;
      alpha_qhqe = (*Veg_Fr_temp) * F(0.8) + F(0.2);
   } else {
      beta = 3;
      if (veg_type == 1) {
         alpha_sl = F(0.686);
         alpha_in = F(0.189);
      } else if (veg_type == 2) {
         alpha_sl = F(0.610);
         alpha_in = F(0.222);
      }
      alpha_qhqe = (*Veg_Fr_temp) * alpha_sl + alpha_in;
   }

   *QH_LUMPS = ((1 - alpha_qhqe) + psyc_s) / (1 + psyc_s) * (qn1 + qf - qs - Qm) - beta;
   if (*QH_LUMPS == NAN_) *QH_LUMPS = qn1 * F(0.2);
   *QE_LUMPS = (alpha_qhqe / (1 + psyc_s) * (qn1 + qf - qs - Qm)) + beta;

   if (*QE_LUMPS > F(0.)) RainBucket = RainBucket - *QE_LUMPS / tlv;
   if (Temp_C > F(0.)) RainBucket = RainBucket - DRAINRT / nsh_real;
   if (RainBucket < F(0.)) RainBucket = F(0.);
   if (Precip > 0) RainBucket = (RainMaxRes < RainBucket + Precip) ? RainMaxRes : RainBucket + Precip;
   RAINRES = RainBucket;
   if (RAINRES > RAINCOVER) RAINRES = RAINCOVER;
   (void)RAINRES;
}

/* evap.f95:5-124 */
void so_cal_evap(int EvapMethod, double state_is, double WetThresh_is, double capStore_is, double vpd_hPa,
                 double avdens, double avcp, double qn_e, double s_hPa, double psyc_hPa, double ResistSurf, double RA,
                 double rb, double tlv, double *rss, double *ev, double *qe)
{
   double numPM, rbsg, rsrbsg, W, x, r;

   numPM = s_hPa * qn_e + vpd_hPa * avdens * avcp / RA;

   if (state_is <= F(0.001)) {
      *qe = numPM / (s_hPa + psyc_hPa * (1 + ResistSurf / RA));
      *ev = *qe / tlv;
   } else {
      if (EvapMethod == 2) {
         rbsg = rb * (s_hPa / psyc_hPa + 1);
         rsrbsg = ResistSurf + rbsg;
         if (state_is >= WetThresh_is || ResistSurf < 25) {
            W = 1;
         } else {
            r = (ResistSurf / RA) * (RA - rb) / rsrbsg;
            W = (r - 1) / (r - (WetThresh_is / state_is));
         }
         *rss = (1 / ((W / rbsg) + ((1 - W) / rsrbsg))) - rbsg;
         *qe = numPM / (s_hPa + psyc_hPa * (1 + *rss / RA));
         *ev = *qe / tlv;
      } else if (EvapMethod == 1) {
         *qe = numPM / (s_hPa + psyc_hPa);
         *ev = *qe / tlv;
         x = (state_is > capStore_is) ? 1.0 : state_is / capStore_is;
         *ev = *ev * x;
         *qe = *ev * tlv;
      }
   }
}
