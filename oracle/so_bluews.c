/* ORACLE (test infrastructure only -- see suews_oracle.h).
 *
 * BLUEWS, the convective boundary-layer slab model: SUEWS-SourceCode/suews_phys_bluews.f95 (CBL :69-296, CBL_initial
 * :361-480, NBL :482-622, RKUTTA :631-704, diff :708-849), called by the reference's grid loop after SUEWS_cal_Main
 * (suews_ctrl_calculations.f95:194-218) with that step's QH / QE (dataOutLineSUEWS(9:10), or the LUMPS pair for
 * QH_choice 2), friction velocity (dataOutLineSUEWS(55)) and the air density, heat capacity and latent heat of
 * vaporisation SUEWS_cal_Main left in MODULE moist (driver.f95: `use moist, only: avcp, avdens, lv_J_kg`).  At night
 * (Kdown < 5) NBL writes a 200 m line and re-initialises the slab from the day's row of the initial-data file; by day CBL
 * integrates height, theta, q (and their values above the inversion) over the timestep with one Runge-Kutta step.  The
 * temperature and humidity it derives replace nothing downstream: SUEWS_Translate re-reads the forcing every step.
 *
 * Pinned: the hourly means of this line reproduce the reference's test1_2004_BL_60.txt (8784 h) and test1_2005_BL_60.txt
 * (120 h) to the printed digit (tests/test_oracle_golden.py).
 *
 * Restated: InitialData_use = 2 (the BaseRun's choice: every quantity from the file) and 1 (height and gradients from the
 * file on the days it lists, theta and q diagnosed from the night's line through psih) with any EntrainmentType and isubs.
 * InitialData_use = 0 cannot run in the reference (NBL reads IniCBLdata, :523-536, which CBL_ReadInputData allocates for 1
 * and 2 only, :325) and sondeflag = 1 (sounding files) is not restated: both stop with 1007.  Only InitialData_use = 2 is
 * pinned (the BL golden files); 1 follows the source alone.
 * Kept as the reference has them: REAL(4) clock arithmetic (float(it) + float(imin)/60.), the REAL(4) exponents
 * 0.3333333333 and 0.6666666667, C2K = 273.16, cm = -999 integrated along when CO2 is off, gamq of the night line
 * being the previous call's value.
 */
#include "suews_oracle.h"
#include <math.h>
#include <stddef.h>

const double *so_cbl_ini = 0;   /* IniCBLdata [nline x 8] row-major: DOY zi0 gamt gamq theta+ q+ theta q */
int so_cbl_ini_nline = 0;
double *so_blk_BL = 0;          /* optional output block of so_cal_multitsteps: [len_sim x 22] column-major */

#define CBL_C2K F(273.16)
#define CBL_GAS_CT_DRY ((double)(8.31451f / 0.028965f))
#define CBL_GRAV F(9.80665)

typedef struct {
   const SoCblParams *p;
   double fhbl_Kms, febl_kgkgms, fcbl, tm_K, gamt_Km, gamq_kgkgm, ustar;
} DiffCtx;

/* bluews.f95:708-849 */
static void cbl_diff(const DiffCtx *c, const double y1[6], double dyds[6])
{
   const SoCblParams *p = c->p;
   double h1 = y1[0], t_K = y1[1], q_kgkg = y1[2], cc = y1[3], tp_K = y1[4], qp_kgkg = y1[5];
   double cp = 0, delt_K, delq_kgkg, delc, ftv_Kms, gamtv_Km, deltv_K, ftva_Kms, ws, dhds = 0, delb, qs2, qs3;
   delt_K = tp_K - t_K;
   delq_kgkg = qp_kgkg - q_kgkg;
   delc = cp - cc;
   ftv_Kms = c->fhbl_Kms + F(0.61) * c->tm_K * c->febl_kgkgms;
   gamtv_Km = c->gamt_Km + F(0.61) * c->tm_K * c->gamq_kgkgm;
   deltv_K = delt_K + F(0.61) * c->tm_K * delq_kgkg;
   ftva_Kms = (ftv_Kms > 0.0) ? ftv_Kms : 0.0;
   ws = pow(h1 * ftva_Kms * CBL_GRAV / c->tm_K, F(0.3333333333));
   if (p->EntrainmentType == 2) {
      dhds = ftva_Kms / (h1 * gamtv_Km);
   } else if (p->EntrainmentType == 1) {
      if (deltv_K <= F(0.01)) {
         dhds = ftva_Kms / (h1 * gamtv_Km);
      } else {
         delb = CBL_GRAV * deltv_K / c->tm_K;
         dhds = (F(0.2) * (ws * ws * ws) + F(5.0) * (c->ustar * c->ustar * c->ustar)) / (h1 * delb);
      }
   } else if (p->EntrainmentType == 4) {
      const double alpha3 = F(0.2);
      if (deltv_K <= F(0.01)) {
         dhds = ftva_Kms / (h1 * gamtv_Km);
      } else if (p->isubs == 1) {
         dhds = alpha3 * ftva_Kms / deltv_K + p->wsb;
      } else {
         dhds = alpha3 * ftva_Kms / deltv_K;
      }
   } else if (p->EntrainmentType == 3) {
      const double conn = F(1.33), conk = F(0.18), cont = F(0.80);
      double cu = conn * c->ustar;
      qs3 = ws * ws * ws + cu * cu * cu;
      qs2 = pow(qs3, F(0.6666666667));
      if (deltv_K <= F(0.01)) {
         dhds = ftva_Kms / (h1 * gamtv_Km);
      } else {
         delb = CBL_GRAV * deltv_K / c->tm_K;
         dhds = (conk * qs3) / (cont * qs2 + h1 * delb);
      }
   }
   if (p->isubs == 1) {
      dyds[1] = c->fhbl_Kms / h1 + delt_K * (dhds - p->wsb) / h1;
      dyds[2] = c->febl_kgkgms / h1 + delq_kgkg * (dhds - p->wsb) / h1;
      dyds[3] = c->fcbl / h1 + delc * (dhds - p->wsb) / h1;
      dyds[4] = c->gamt_Km * (dhds - p->wsb);
      dyds[5] = c->gamq_kgkgm * (dhds - p->wsb);
   } else {
      dyds[1] = c->fhbl_Kms / h1 + delt_K * (dhds) / h1;
      dyds[2] = c->febl_kgkgms / h1 + delq_kgkg * (dhds) / h1;
      dyds[3] = c->fcbl / h1 + delc * (dhds) / h1;
      dyds[4] = c->gamt_Km * (dhds);
      dyds[5] = c->gamq_kgkgm * (dhds);
   }
   dyds[0] = dhds;
}

/* bluews.f95:631-704 with NSTEPS = 1 */
static void cbl_rkutta(const DiffCtx *c, double xa, double xb, double y[6])
{
   const double coef[4] = {1.0, 0.5, 0.5, 1.0};
   double rk[6][5], arg[6], dydx[6], step = (xb - xa) / 1;
   int n, nj;
   for (nj = 0; nj < 6; nj++) rk[nj][0] = 0;
   for (n = 0; n < 4; n++) {
      for (nj = 0; nj < 6; nj++) arg[nj] = y[nj] + coef[n] * rk[nj][n];
      cbl_diff(c, arg, dydx);
      for (nj = 0; nj < 6; nj++) rk[nj][n + 1] = step * dydx[nj];
   }
   for (nj = 0; nj < 6; nj++)
      for (n = 0; n < 4; n++) y[nj] = y[nj] + rk[nj][n + 1] / (6 * coef[n]);
}

/* the day's row of the initial-data file: the last row whose day of year is <= id (bluews.f95:402-407) */
static const double *cbl_row(int id)
{
   int i, n = 0;
   for (i = 0; i < so_cbl_ini_nline; i++)
      if ((int)so_cbl_ini[(size_t)i * 8] <= id) n++;
   return n > 0 ? so_cbl_ini + (size_t)(n - 1) * 8 : 0;
}

/* CBL / NBL for one step.  line22: iy id it imin dectime z theta q theta+ q+ Temp_C rh QH_use QE_use Press_hPa avu1 UStar
 * avdens lv_J_kg avcp gamt gamq */
void so_CBL(SoCtx *cx, const SoCblParams *p, SoCblState *s, const SoForcing *f, double dectime, int tstep, double qh_use,
            double qe_use, double UStar, double avdens, double lv_J_kg, double avcp, double line22[22])
{
   const double *ini;
   double Temp_C, es_hPa, avrh, qm_gkg;
   int k;
   line22[0] = f->iy; line22[1] = f->id; line22[2] = f->it; line22[3] = f->imin; line22[4] = dectime;
   for (k = 5; k < 22; k++) line22[k] = -999;
   if ((p->InitialData_use != 2 && p->InitialData_use != 1) || p->sondeflag != 0 || !so_cbl_ini) {
      so_error_hint(cx, 1007, "oracle: BLUEWS restated for InitialData_use = 1 / 2, sondeflag = 0 only", -999, -999, p->InitialData_use);
      return;
   }
   if (qh_use < -900 || qe_use < -900) so_error_hint(cx, 22, "Unrealistic qh or qe_value for CBL.", qh_use, qe_use, p->QH_choice);
   if (f->avkdn < 5) {
      /* NBL (bluews.f95:482-622) */
      double qp_gkg;
      ini = cbl_row(f->id);
      if (!ini) { so_error_hint(cx, 1007, "oracle: no initial CBL data for this day", -999, -999, f->id); return; }
      s->tm_K = ini[6];
      qm_gkg = ini[7];
      s->blh_m = 200;
      s->gamt_Km = ini[2];
      s->tp_K = ini[4];
      qp_gkg = ini[5];
      Temp_C = s->tm_K / pow(1000 / f->Press_hPa, CBL_GAS_CT_DRY / avcp) - CBL_C2K;
      es_hPa = so_sat_vap_press_x(cx, &Temp_C, f->Press_hPa, 1, dectime);   /* (rewrites Temp_C when |T| < 0.001) */
      avrh = 100 * ((qm_gkg * f->Press_hPa / (622 + qm_gkg)) / es_hPa);
      if (avrh > 100) {
         so_error_hint(cx, 34, "subroutine CBL dectime, relative humidity", dectime, avrh, 100);
         avrh = 100;
      }
      line22[5] = s->blh_m; line22[6] = s->tm_K; line22[7] = qm_gkg; line22[8] = s->tp_K; line22[9] = qp_gkg;
      line22[10] = Temp_C; line22[11] = avrh; line22[12] = qh_use; line22[13] = qe_use; line22[14] = f->Press_hPa;
      line22[15] = f->avU1; line22[16] = UStar; line22[17] = avdens; line22[18] = lv_J_kg; line22[19] = avcp;
      line22[20] = s->gamt_Km; line22[21] = s->gamq_kgkgm;   /* gamq: still the previous call's value */
      if (p->InitialData_use == 2) {
         s->blh_m = ini[1];
         s->gamt_Km = ini[2];
         s->tp_K = ini[4];
         s->tm_K = ini[6];
         s->gamq_kgkgm = ini[3] / F(1000.);
         s->qp_kgkg = ini[5] / 1000;
         s->qm_kgkg = ini[7] / 1000;
      } else {
         /* InitialData_use == 1 (bluews.f95:570-581): only on a day the file lists; otherwise the night's 200 m values stay */
         double gamq_gkgm = ini[3];
         if (ini[0] == f->id) {
            const double k = F(0.4);
            double tm_K_zm, qm_gkg_zm, lv;
            s->blh_m = ini[1];
            s->gamt_Km = ini[2];
            tm_K_zm = (Temp_C + CBL_C2K) * pow(1000 / f->Press_hPa, CBL_GAS_CT_DRY / avcp);
            s->tm_K = tm_K_zm - cx->last_psih * qh_use / (k * UStar * avcp * avdens);
            es_hPa = so_sat_vap_press_x(cx, &Temp_C, f->Press_hPa, 1, dectime);
            qm_gkg_zm = 622 * avrh / (100 * f->Press_hPa / es_hPa - avrh);
            lv = (F(2500.25) - F(2.365) * Temp_C) * 1000;
            qm_gkg = qm_gkg_zm - cx->last_psih * qe_use / (k * UStar * avdens * lv);
            s->tp_K = s->tm_K;
            qp_gkg = qm_gkg;
         }
         s->gamq_kgkgm = gamq_gkgm / F(1000.);
         s->qp_kgkg = qp_gkg / 1000;
         s->qm_kgkg = qm_gkg / 1000;
      }
      {
         double tp_C = s->tp_K - CBL_C2K, tm_C = s->tm_K - CBL_C2K, qs;
         qs = so_qsatf(cx, tp_C, f->Press_hPa);
         if (s->qp_kgkg > qs || s->qp_kgkg < 0) s->qp_kgkg = qs;
         qs = so_qsatf(cx, tm_C, f->Press_hPa);
         if (s->qm_kgkg > qs || s->qm_kgkg < 0) s->qm_kgkg = qs;
      }
      return;
   }
   /* CBL (bluews.f95:69-296) */
   {
      DiffCtx c;
      double y[6], secs0, secs1, hr;
      hr = (double)((float)f->it + (float)f->imin / 60.f);   /* float(it) + float(imin)/60. */
      secs0 = hr * F(3600.);
      secs1 = secs0 + (double)(float)tstep;
      c.p = p;
      c.fhbl_Kms = qh_use / (avdens * avcp);
      c.febl_kgkgms = qe_use / (avdens * lv_J_kg);
      c.fcbl = 0;
      if (p->CO2_included != 1) s->cm = -999;
      c.tm_K = s->tm_K; c.gamt_Km = s->gamt_Km; c.gamq_kgkgm = s->gamq_kgkgm; c.ustar = UStar;
      y[0] = s->blh_m; y[1] = s->tm_K; y[2] = s->qm_kgkg; y[3] = s->cm; y[4] = s->tp_K; y[5] = s->qp_kgkg;
      cbl_rkutta(&c, secs0, secs1, y);
      s->blh_m = y[0]; s->tm_K = y[1]; s->qm_kgkg = y[2]; s->cm = y[3]; s->tp_K = y[4]; s->qp_kgkg = y[5];
      qm_gkg = s->qm_kgkg * 1000;
      Temp_C = s->tm_K / pow(1000 / f->Press_hPa, CBL_GAS_CT_DRY / avcp) - CBL_C2K;
      es_hPa = so_sat_vap_press_x(cx, &Temp_C, f->Press_hPa, 1, dectime);   /* (rewrites Temp_C when |T| < 0.001) */
      avrh = 100 * ((qm_gkg * f->Press_hPa / (622 + qm_gkg)) / es_hPa);
      if (avrh > 100) {
         so_error_hint(cx, 34, "subroutine CBL dectime, relative humidity", dectime, avrh, 100);
         avrh = 100;
      }
      line22[5] = s->blh_m; line22[6] = s->tm_K; line22[7] = s->qm_kgkg * 1000; line22[8] = s->tp_K;
      line22[9] = s->qp_kgkg * 1000; line22[10] = Temp_C; line22[11] = avrh; line22[12] = qh_use; line22[13] = qe_use;
      line22[14] = f->Press_hPa; line22[15] = f->avU1; line22[16] = UStar; line22[17] = avdens; line22[18] = lv_J_kg;
      line22[19] = avcp; line22[20] = s->gamt_Km; line22[21] = s->gamq_kgkgm;
   }
}
