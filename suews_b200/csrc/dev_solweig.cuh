// dev_solweig.cuh -- SOLWEIG point diagnostics (mean radiant temperature and the directional short/long-wave
// fluxes at a point), the last call of SUEWS_cal_Main for cells with buildings (suews_ctrl_driver.f95:997-1002).
// Device restatement of the point version in suews_phys_solweig.f95: SOLWEIG_cal_main :53-457,
// cal_ratio_height2width :459, hwToSVF_ground :477, clearnessindex_2013b :507, sun_distance :623,
// cylindric_wedge :637, diffusefraction :720, Kside_veg_v24 :1045, Kvikt_veg :1129, Lside_veg_v2 :1153,
// Lvikt_veg :1371, shadowingfunction_urban :1522, sunonsurface_veg :1860.
//
// The reference runs its raster routines on 1 x 1 "grids", where two of them degenerate: every shifted array
// section of shadowingfunction_urban is empty, so shadow = 1 whenever the sun is up; sunonsurface_veg copies its
// inputs only in the pass with index 0, so gvf1 = MIN((first + 1)/first, 1) and gvf2 = second/second for all 18
// azimuths.  They are evaluated as those results.  I0 is assigned in daylight only (the reference emits stack
// residue at night): -999 here.  The output is a pure function of the step's inputs and feeds nothing back, so it
// is only evaluated when the SOLWEIG output line is requested.
#pragma once
#include "dev_common.cuh"

#define PI_SOL F(3.141592653589793) // REAL(4) literal, like every "pi" in this file of the reference
#define DEG2RAD_SOL F(0.017453292)  // suews_ctrl_const.f95:387
#define SBC_SOL F(5.67051e-8)

SU_INL double sol_p6(double x) { double a = x * (x * x); return a * a; }

// the sixth-order view-factor polynomial shared by Kvikt_veg / Lvikt_veg
SU_INL double sol_vikt_poly(double s)
{
   return F(63.227) * sol_p6(s) - F(161.51) * su_p5(s) + F(156.91) * su_p4(s) - F(70.424) * su_p3(s) + F(16.773) * su_p2(s) -
          F(0.4863) * s;
}

// solweig.f95:637-716
SU_DEV double sol_cylindric_wedge(double zen, double svfalfa)
{
   double beta = zen, alfa = svfalfa;
   double xa = F(1.) - F(2.) / (tan(alfa) * tan(beta));
   double ha = F(2.) / (tan(alfa) * tan(beta));
   double ba = (F(1.) / tan(alfa));
   double hkil = F(2.) * ba * ha;
   double qa = 0.0, Za = 0.0, phi = 0.0, Ai = 0.0, ukil = 0.0, Ssurf;
   if (xa < 0) qa = tan(beta) / 2;
   if (xa < 0) {
      Za = pow(su_p2(ba) - su_p2(qa) / F(4.), F(0.5));
      phi = atan(Za / qa);
      Ai = (sin(phi) - phi * cos(phi)) / (1 - cos(phi));
      ukil = 2 * ba * xa * Ai;
   }
   Ssurf = hkil + ukil;
   return (2 * PI_SOL * ba - Ssurf) / (2 * PI_SOL * ba);
}

// solweig.f95:507-619
SU_DEV void sol_clearnessindex(double zen, int DOY, double Ta, double RH, double radG, double lat, double P_kPa,
                                 double &I0, double &CI, double &Kt, double &I0et, double &CIuncorr)
{
   double p, Itoa, D, m, Trpg, G[4] = {0, 0, 0, 0}, iG = 0, a2, b2, Td, u, Tw, Tar, corr, b;
   if (P_kPa == -999) p = 1013; else p = P_kPa * 10;
   Itoa = 1370;
   b = (double)(((2.f * 3.141592654f) * (float)DOY) / 365.f); // sun_distance :623-635: INTEGER jday, REAL(4) literals -> single-precision b
   D = sqrt(F(1.00011) + F(0.034221) * cos(b) + F(0.001280) * sin(b) + F(0.000719) * cos(2 * b) + F(0.000077) * sin(2 * b));
   m = F(35.) * cos(zen) * pow(F(1224.) * su_p2(cos(zen)) + F(1.), (double)(-1.f / 2.f));
   Trpg = F(1.021) - F(0.084) * pow(m * (F(0.000949) * p + F(0.051)), F(0.5));
#define SETG(a, b_, c, d) do { G[0] = F(a); G[1] = F(b_); G[2] = F(c); G[3] = F(d); } while (0)
   if (lat < 10) SETG(3.37, 2.85, 2.80, 2.64);
   else if (lat >= 10 && lat < 20) SETG(2.99, 3.02, 2.70, 2.93);
   else if (lat >= 20 && lat < 30) SETG(3.60, 3.00, 2.98, 2.93);
   else if (lat >= 30 && lat < 40) SETG(3.04, 3.11, 2.92, 2.94);
   else if (lat >= 40 && lat < 50) SETG(2.70, 2.95, 2.77, 2.71);
   else if (lat >= 50 && lat < 60) SETG(2.52, 3.07, 2.67, 2.93);
   else if (lat >= 60 && lat < 70) SETG(1.76, 2.69, 2.61, 2.61);
   else if (lat >= 70 && lat < 80) SETG(1.60, 1.67, 2.24, 2.63);
   else if (lat >= 80 && lat < 90) SETG(1.11, 1.44, 1.94, 2.02);
#undef SETG
   if (DOY > 335 || DOY <= 60) iG = G[0];
   else if (DOY > 60 && DOY <= 152) iG = G[1];
   else if (DOY > 152 && DOY <= 244) iG = G[2];
   else if (DOY > 244 && DOY <= 335) iG = G[3];
   a2 = F(17.27);
   b2 = F(237.7);
   Td = (b2 * (((a2 * Ta) / (b2 + Ta)) + log(RH))) / (a2 - (((a2 * Ta) / (b2 + Ta)) + log(RH)));
   Td = (Td * F(1.8)) + 32;
   u = exp(F(0.1133) - log(iG + 1) + F(0.0393) * Td);
   Tw = 1 - F(0.077) * pow(u * m, F(0.3));
   Tar = pow(F(0.935), m);
   I0 = Itoa * cos(zen) * Trpg * Tw * D * Tar;
   corr = F(0.1473) * log(90 - (zen / PI_SOL * 180)) + F(0.3454);
   CIuncorr = radG / I0;
   CI = CIuncorr + (1 - corr);
   I0et = Itoa * cos(zen) * D;
   Kt = radG / I0et;
}

// solweig.f95:720-775 (radI comes back as the beam, radD as the diffuse component)
SU_DEV void sol_diffusefraction(double radG, double altitude, double Kt, double Ta, double RH, double &radI, double &radD)
{
   double alfa = altitude * DEG2RAD_SOL, d = 0;
   if (Ta <= -99 || RH <= -99) {
      if (Kt <= F(0.3)) d = radG * (F(1.020) - F(0.248) * Kt);
      else if (Kt > F(0.3) && Kt < F(0.78)) d = radG * (F(1.45) - F(1.67) * Kt);
      else if (Kt >= F(0.78)) d = radG * F(0.147);
   } else {
      if (Kt <= F(0.3)) d = radG * (1 - F(0.232) * Kt + F(0.0239) * sin(alfa) - F(0.000682) * Ta + F(0.0195) * (RH / 100));
      else if (Kt > F(0.3) && Kt < F(0.78))
         d = radG * (F(1.329) - F(1.716) * Kt + F(0.267) * sin(alfa) - F(0.00357) * Ta + F(0.106) * (RH / 100));
      else if (Kt >= F(0.78)) d = radG * (F(0.426) * Kt - F(0.256) * sin(alfa) + F(0.00349) * Ta + F(0.0734) * (RH / 100));
   }
   d = (0.0 > d) ? 0.0 : d;
   d = (radG < d) ? radG : d;
   radD = d;
   radI = (radG - d) / (sin(alfa));
   if (altitude < 1 && radI > radG) radI = radG;
}

// one wall direction of Kside_veg_v24 (solweig.f95:1045-1127): all direction view factors are 1
SU_DEV double sol_kside_one(int lit, double radI, double radG, double radD, double shadow, double F_sh, double altitude,
                        double azi_dir, double psi, double albedo)
{
   const double vikttot = F(4.4897);
   double viktwall = (vikttot - sol_vikt_poly(1.0)) / vikttot;
   double svfvegbu = (1.0 + 1.0 - 1);
   double viktveg = (vikttot - sol_vikt_poly(svfvegbu)) / vikttot;
   double svfviktbuveg;
   viktveg = viktveg - viktwall;
   svfviktbuveg = (viktwall + (viktveg) * (1 - psi));
   if (lit)
      return radI * shadow * cos(altitude * (PI_SOL / 180)) * sin(azi_dir * (PI_SOL / 180)) + radD * (1 - svfviktbuveg) +
             radG * albedo * svfviktbuveg * (1 - F_sh);
   return radD * (1 - svfviktbuveg) + radG * albedo * svfviktbuveg * (1 - F_sh);
}

// one wall direction of Lside_veg_v2 (solweig.f95:1153-1369)
SU_DEV double sol_lside_one(int daytime, int lit, double azi_dir, double Ldown2d, double Lup2d, double Ta, double Tw,
                        double emis_wall, double F_sh, double Lsky_allsky, double svfalfaX)
{
   const double vikttot = F(4.4897);
   // Lvikt_veg :1371-1421 with isvf = isvfveg = isvfaveg = 1
   double viktonlywall = (vikttot - sol_vikt_poly(1.0)) / vikttot;
   double viktaveg = (vikttot - sol_vikt_poly(1.0)) / vikttot;
   double viktwall = viktonlywall - viktaveg;
   double svfvegbu = (1.0 + 1.0 - 1);
   double viktsky = sol_vikt_poly(svfvegbu) / vikttot;
   double viktrefl = (vikttot - sol_vikt_poly(svfvegbu)) / vikttot;
   double viktveg = (vikttot - sol_vikt_poly(svfvegbu)) / vikttot;
   double Lwallsun, Lwallsh, Lsky, Lveg, Lground, Lrefl;
   viktveg = viktveg - viktwall;
   if (daytime) {
      double alfaB = atan(svfalfaX);
      double betaB = atan(tan((svfalfaX)*F_sh));
      double betasun = ((alfaB - betaB) / 2) + betaB;
      if (lit) {
         Lwallsun = SBC_SOL * emis_wall * su_p4(Ta + F(273.15) + Tw * sin(azi_dir * (PI_SOL / 180))) * viktwall * (1 - F_sh) *
                    cos(betasun) * F(0.5);
         Lwallsh = SBC_SOL * emis_wall * su_p4(Ta + F(273.15)) * viktwall * F_sh * F(0.5);
      } else {
         Lwallsun = 0;
         Lwallsh = SBC_SOL * emis_wall * su_p4(Ta + F(273.15)) * viktwall * F(0.5);
      }
   } else {
      Lwallsun = 0;
      Lwallsh = SBC_SOL * emis_wall * su_p4(Ta + F(273.15)) * viktwall * F(0.5);
   }
   Lsky = ((1.0 + 1.0 - 1) * Lsky_allsky) * viktsky * F(0.5);
   Lveg = SBC_SOL * emis_wall * su_p4(Ta + F(273.15)) * viktveg * F(0.5);
   Lground = Lup2d * F(0.5);
   Lrefl = (Ldown2d + Lup2d) * (viktrefl) * (1 - emis_wall) * F(0.5);
   return Lsky + Lwallsun + Lwallsh + Lveg + Lground + Lrefl;
}

// SOLWEIG_cal_main, solweig.f95:53-457
SU_DEVNI void SOLWEIG_cal_main(int id, int it, double dectime, double lamdaP, double lamdaF, double avkdn, double ldown,
                         double Temp_C, double avrh, double Press_hPa, double Tg, double lat, double zenith_deg,
                         double azimuth, double scale, double alb_ground, double alb_bldg, double emis_ground,
                         double emis_wall, double heightgravity, double *out)
{
   const double absL = F(0.97), absK = F(0.7);
   const double Fside = F(0.22), Fup = F(0.06);
   const double psi = 0, t = 0, svfveg = 1, svfaveg = 1, buildings = 1, Knight = 0.0;
   double P = Press_hPa, Ta = Temp_C, RH = avrh, radG = avkdn, height = heightgravity;
   int DOY = id, first, second, j;
   double HW, svf, tmp, svfalfa, svf_blgd_veg, zen, altitude, ea, msteg, emis_sky;
   double CI, I0 = -999, Kt = 0, I0et, CIuncorr, radI = 0, radD = 0, shadow, gvf, F_sh, Tw;
   double Kdown2d, Kup2d, Keast, Knorth, Ksouth, Kwest, Lup2d, Ldown2d, Least, Lnorth, Lsouth, Lwest, Sstr, Tmrt;
   double CIlatenight = 1;
   (void)scale;

   {  // cal_ratio_height2width :459-475, hwToSVF_ground :477-490
      double lamdaW = 4 * lamdaF;
      HW = (lamdaW * lamdaP) / (2 * lamdaP * (1 - lamdaP));
      svf = F(0.5598) * exp(F(-0.2485) * HW) + F(0.4112) * exp(F(-0.02528) * HW);
   }
   tmp = 1 - (svf + svfveg - 1);
   if (tmp <= 0) tmp = F(0.000000001);
   svfalfa = asin(exp(log(tmp) / 2));
   first = (int)round(height);
   if (first == 0) first = 1;
   second = (int)round(height * 20);
   svf_blgd_veg = (svf - (1 - svfveg) * (1 - psi));
   zen = zenith_deg * DEG2RAD_SOL;
   altitude = 90 - zenith_deg;
   ea = F(6.107) * pow(10.0, (F(7.5) * Ta) / (F(237.3) + Ta)) * (RH / 100);
   msteg = F(46.5) * (ea / (Ta + F(273.15)));
   emis_sky = (1 - (1 + msteg) * exp(-(pow(F(1.2) + F(3.0) * msteg, F(0.5))))) - F(0.04);
   (void)emis_sky; // only used when SOLWEIG_ldown = 1

   if (altitude > 0) {
      double sos, gvf1, gvf2;
      sol_clearnessindex(zen, DOY, Ta, RH / 100, radG, lat, P, I0, CI, Kt, I0et, CIuncorr);
      if (CI > 1) CI = 1;
      CIlatenight = CI;
      sol_diffusefraction(radG, altitude, Kt, Ta, RH, radI, radD);
      shadow = 1.0; // shadowingfunction_urban on a 1x1 grid, see header
      // sunonsurface_veg on a 1x1 grid, see header; first / second are unchanged by scale = 1
      gvf1 = (first <= second - 1) ? ((0.0 + (double)(first + 1)) / first) : 0.0;
      if (gvf1 > 1) gvf1 = 1.;
      gvf2 = (second >= 1) ? ((double)second + 0.0) / second : 0.0;
      if (gvf2 > 1) gvf2 = 1.;
      sos = (gvf1 * F(0.5) + gvf2 * F(0.4)) / F(0.9);
      gvf = 0.0;
      for (j = 1; j <= 18; j++) gvf = gvf + sos;
      gvf = gvf / 18 + (buildings * (-1) + 1);
      F_sh = sol_cylindric_wedge(zen, svfalfa);
      Kdown2d = radI * shadow * sin(altitude * DEG2RAD_SOL) + radD * svf_blgd_veg +
                radG * alb_bldg * (1 - svf_blgd_veg) * (1 - F_sh);
      Kup2d = alb_ground * (radI * gvf * sin(altitude * DEG2RAD_SOL) + radD * svf_blgd_veg +
                            radG * alb_bldg * (1 - svf_blgd_veg) * (1 - F_sh));
      {  // Kside_veg_v24
         double aziE = azimuth + t, aziS = azimuth - 90 + t, aziW = azimuth - 180 + t, aziN = azimuth - 270 + t;
         Keast = sol_kside_one(azimuth > (360 - t) || azimuth <= (180 - t), radI, radG, radD, shadow, F_sh, altitude, aziE, psi, alb_ground);
         Ksouth = sol_kside_one(azimuth > (90 - t) && azimuth <= (270 - t), radI, radG, radD, shadow, F_sh, altitude, aziS, psi, alb_ground);
         Kwest = sol_kside_one(azimuth > (180 - t) && azimuth <= (360 - t), radI, radG, radD, shadow, F_sh, altitude, aziW, psi, alb_ground);
         Knorth = sol_kside_one(azimuth <= (90 - t) || azimuth > (270 - t), radI, radG, radD, shadow, F_sh, altitude, aziN, psi, alb_ground);
      }
      Tw = Tg;
      {
         double Tgmap0 = gvf * Tg + Ta, Tgmap1 = Tgmap0;
         double timeadd = (dectime + 1) - DOY - (double)((float)it / 24.f);
         double weight1 = exp(F(-33.27) * timeadd);
         Tgmap1 = Tgmap0 * (1 - weight1) + Tgmap1 * weight1;
         Lup2d = SBC_SOL * emis_ground * su_p4(Tgmap1 + F(273.15));
      }
   } else {
      if (dectime < (DOY + F(0.5)) && dectime > DOY && altitude < F(1.0)) CI = F(1.0);
      else CI = CIlatenight;
      Tw = F(0.0);
      radG = 0; radI = 0; radD = 0;
      Kdown2d = 0; Kwest = 0; Kup2d = 0; Keast = 0; Ksouth = 0; Knorth = 0;
      shadow = 0; gvf = 0;
      Lup2d = SBC_SOL * emis_ground * su_p4(Knight + Ta + Tg + F(273.15));
      I0 = -999; // unassigned in the reference at night
   }
   Ldown2d = (svf + svfveg - 1) * ldown + (2 - svfveg - svfaveg) * emis_wall * SBC_SOL * su_p4(Ta + F(273.15)) +
             (svfaveg - svf) * emis_wall * SBC_SOL * su_p4(Ta + F(273.15) + Tw) + (2 - svf - svfveg) * (1 - emis_wall) * ldown;
   {  // Lside_veg_v2
      const int day = altitude > 0;
      double oneminus = F(1.) - 1.0, svfalfaX, F2;
      double aziW = azimuth + t, aziN = azimuth - 90 + t, aziE = azimuth - 180 + t, aziS = azimuth - 270 + t;
      if (oneminus <= 0) oneminus = F(0.000000001);
      svfalfaX = asin(exp((log(oneminus)) / 2));
      F2 = sol_cylindric_wedge(zen, svfalfa);
      F2 = F(2.) * F2 - F(1.);
      Least = sol_lside_one(day, azimuth > (180 - t) && azimuth <= (360 - t), aziE, Ldown2d, Lup2d, Ta, Tw, emis_wall, F2, ldown, svfalfaX);
      Lsouth = sol_lside_one(day, azimuth <= (90 - t) || azimuth > (270 - t), aziS, Ldown2d, Lup2d, Ta, Tw, emis_wall, F2, ldown, svfalfaX);
      Lwest = sol_lside_one(day, azimuth > (360 - t) || azimuth <= (180 - t), aziW, Ldown2d, Lup2d, Ta, Tw, emis_wall, F2, ldown, svfalfaX);
      Lnorth = sol_lside_one(day, azimuth > (90 - t) && azimuth <= (270 - t), aziN, Ldown2d, Lup2d, Ta, Tw, emis_wall, F2, ldown, svfalfaX);
   }
   Sstr = absK * (Kdown2d * Fup + Kup2d * Fup + Knorth * Fside + Keast * Fside + Ksouth * Fside + Kwest * Fside) +
          absL * (Ldown2d * Fup + Lup2d * Fup + Lnorth * Fside + Least * Fside + Lsouth * Fside + Lwest * Fside);
   Tmrt = sqrt(sqrt((Sstr / (absL * SBC_SOL)))) - F(273.15);
   out[0] = azimuth; out[1] = altitude; out[2] = radG; out[3] = radI; out[4] = radD;
   out[5] = Kdown2d; out[6] = Kup2d; out[7] = Ksouth; out[8] = Kwest; out[9] = Knorth; out[10] = Keast;
   out[11] = Ldown2d; out[12] = Lup2d; out[13] = Lsouth; out[14] = Lwest; out[15] = Lnorth; out[16] = Least;
   out[17] = Tmrt; out[18] = I0; out[19] = CI; out[20] = gvf; out[21] = shadow; out[22] = svf; out[23] = svf_blgd_veg;
   out[24] = Ta; out[25] = Tg;
}
