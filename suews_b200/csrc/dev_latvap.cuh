// dev_latvap.cuh -- saturation vapour pressure, the psychrometric constant and Lat_vap, the wet-bulb search of
// suews_util_meteo.f95:148-192, :409-427, :277-346 (included by dev_atmos.cuh; also compiled for the host by the CPU test).
//
// The reference keeps switch1/switch2 in implicit-SAVE locals, so their exit values seed the next call; here they are
// per-cell state (sw = switch1 | switch2 << 1).
//
// The search takes ~21 trips on the bench forcing, each with a full sat_vap_press (EXP) and a division, and all it keeps
// of them is a three-way decision: |ea_est - ea_fix| < 0.001, ea_est above, ea_est below.  The result (tw, hence lv, and the
// two switches) depends on those decisions only, so a trip may take its decision any way that provably gives the
// reference's answer.  Three levels, cheapest first:
//
//  (1) ROOT: ea_est - ea_fix = g_k(tw) = es(tw) - psyc_k (T - tw) - ea, psyc_k = cp P / (eps lv_k), lv_k the latent heat the
//      previous trip left (from its tw; from T on the first trip).  G(t) = es(t) - psyc(lv(t)) (T - t) - ea is the same
//      function with lv taken at t itself.  On (-39, 49) G is increasing with G' >= m = 0.9 psyc_min (es' > 0; the water /
//      ice switch at 0.001 jumps upward; d/dt of the psyc term is psyc (1 - 2365 (T - t) / lv) >= 0.9 psyc for T - t < 90),
//      and |g_k(tw) - G(tw)| = |psyc_k - psyc(lv(tw))| |T - tw| <= psyc_max 0.001 |tw - t_k| |T - tw| =: D_k (lv >= 2.384e6
//      on the interval, so the relative change of 1/lv per kelvin is below 0.000992).  A few single-precision Newton steps
//      give a point r with a small residual Gf(r), whose true value is within e_f of it.  Then for tw > r:
//      g_k(tw) >= G(r) + m (tw - r) - D_k >= m (tw - r) - |Gf(r)| - e_f - D_k, which exceeds the 0.001 threshold -- the
//      decision is "above" -- as soon as m |tw - r| > |Gf(r)| + e_f + D_k + 0.001 (+ slack); symmetrically below.  No
//      assumption on how well Newton converged: a poor r only makes the test fail more often.
//  (2) SCREEN: g_k(tw) itself in single precision (fast intrinsics; error <= ~1.5e-6 of the terms' magnitudes), decided when
//      it clears the threshold by a margin several times that error.
//  (3) the reference expression in double: the last one to three trips, where tw is within a few thousandths of a kelvin
//      of the root.
// Same decisions, same tw sequence (the additions tw +- incr are the reference's), same bits: checked against the plain
// loop on 2x10^8 random arguments by suews_b200_selftest_latvap on the device and, for the logic of level (1), on the host
// by tests/test_host.py (tests/host_latvap.cpp compiles this file).  SU_LATVAP_MODE: 0 plain, 1 levels (2)+(3) (default), 2 all three.
#pragma once
// Measured on the B200 (606 208 cells x 288 steps, January days): levels (2)+(3) 4.74e8 cell-steps/s, all three 4.66e8 -- the
// root bound decides 25 of a search's 33 trips for a dozen instructions each, but its five Newton steps are a dependent
// chain of their own in front of the loop and the phase is latency-bound, not issue-bound; so the default is mode 1 and
// level (1) stays as a tuning variant (and as a checked statement of why the decisions are what they are).
#ifndef SU_LATVAP_MODE
#ifdef SU_LATVAP_PLAIN
#define SU_LATVAP_MODE 0
#else
#define SU_LATVAP_MODE 1
#endif
#endif

// suews_util_meteo.f95:148-192.  The reference rewrites its argument T to 0.001 when |T| < 0.001 (a side
// effect the callers see); by-reference scalars of out-of-line device functions live in local memory, so the
// rewrite is a separate inline step (su_fix_T) and the evaluation takes T by value.  Warnings 28/29 are
// raised by the callers through sat_vap_press_warn.
SU_INL double su_fix_T(double T) { return (fabs(T) < F(0.001000)) ? (double)F(0.001000) : T; }
SU_HOT double sat_vap_press_v(double T, double Press_hPa)   // T already through su_fix_T
{
   // water (T >= 0.001) and ice (T <= -0.001) branches of meteo.f95:171-186 differ only in seven constants;
   // selecting the constants keeps a warp whose cells straddle 0 degC on one instruction stream
   const bool w = (T >= F(0.001000));
   const double c0 = w ? F(6.1121) : F(6.1115), c1 = w ? F(18.678) : F(23.036), c2 = w ? F(234.5) : F(333.7),
                c3 = w ? F(257.14) : F(279.82), f0 = w ? F(1.00072) : F(1.00022), f1 = w ? F(3.2E-6) : F(3.83E-6),
                f2 = w ? F(5.9E-10) : F(6.4E-10);
   const double Press_kPa = Press_hPa / 10;
   const double e_mb = c0 * exp(((c1 - T / c2) * T) / (T + c3));
   const double f = f0 + Press_kPa * (f1 + f2 * (T * T));
   const bool in_range = (T < 50 && T > -40) && (w || T <= -F(0.001000));
   return in_range ? e_mb * f : 1000.0;
}
SU_INL double sat_vap_press_x(SuDiag &dg, double &T, double Press_hPa, int from)
{
   if (fabs(T) < F(0.001000)) {
      if (from == 1) su_warn(dg, 29);
      T = F(0.001000);
   }
   if (!(T < 50 && T > -40)) su_warn(dg, 28);
   return sat_vap_press_v(T, Press_hPa);
}

// suews_util_meteo.f95:409-427
SU_INL double psyc_const(SuDiag &dg, double cp, double Press_hPa, double lv_J_kg)
{
   const double epsil = F(0.62197);
   if (cp * Press_hPa < 900 || lv_J_kg < 10000) su_warn(dg, 19);
   return (cp * Press_hPa) / (epsil * lv_J_kg);
}

SU_INL float su_es_single(float T, float Press_kPa)
{
   const bool w = (T >= 0.001f);
   const float c0 = w ? 6.1121f : 6.1115f, c1 = w ? 18.678f : 23.036f, c2 = w ? 234.5f : 333.7f, c3 = w ? 257.14f : 279.82f,
               f0 = w ? 1.00072f : 1.00022f, f1 = w ? 3.2E-6f : 3.83E-6f, f2 = w ? 5.9E-10f : 6.4E-10f;
   // fast intrinsics (MUFU): __expf <= 6 ulp for arguments below 4, __fdividef <= 2 ulp
   return (c0 * __expf(__fdividef((c1 - __fdividef(T, c2)) * T, T + c3))) * (f0 + Press_kPa * (f1 + f2 * (T * T)));
}
// es and its temperature derivative in single precision (derivative: for the Newton step only, accuracy irrelevant)
SU_INL void su_es_single_d(float T, float Press_kPa, float &es, float &des)
{
   const bool w = (T >= 0.001f);
   const float c0 = w ? 6.1121f : 6.1115f, c1 = w ? 18.678f : 23.036f, c2 = w ? 234.5f : 333.7f, c3 = w ? 257.14f : 279.82f,
               f0 = w ? 1.00072f : 1.00022f, f1 = w ? 3.2E-6f : 3.83E-6f, f2 = w ? 5.9E-10f : 6.4E-10f;
   const float a = c1 - __fdividef(T, c2), d = T + c3, rd = __fdividef(1.0f, d);
   const float e = c0 * __expf(a * T * rd);
   const float f = f0 + Press_kPa * (f1 + f2 * (T * T));
   es = e * f;
   // u = a T / d, u' = ((c1 - 2T/c2) d - a T) / d^2
   const float du = ((a - __fdividef(T, c2)) * d - a * T) * rd * rd;
   des = es * du + e * Press_kPa * 2.0f * f2 * T;
}

template <int MODE>
SU_INL double Lat_vap_t(SuDiag &dg, double Temp_C, double Ea_hPa, double Press_hPa, double cp, int &sw)
{
   bool switch1 = (sw & 1) != 0, switch2 = (sw & 2) != 0;
   const double ea_fix = Ea_hPa;
   double lv_J_kg = (F(2500.25) - F(2.365) * Temp_C) * 1000;
   double tw = Temp_C / F(2.);
   double incr = F(3.);
   if (Press_hPa < 900) su_warn(dg, 45);
   const float Tf = (float)Temp_C, eaf = (float)Ea_hPa, pkf = (float)(Press_hPa / 10), cpPf = (float)(cp * Press_hPa);
   // ---- level (1): a point r near the root of G and the bound it supports
   float rf = 0.f, need0 = 0.f, m_lo = 0.f, psyc_hi = 0.f;
   bool have_r = false;
   double t_lv = Temp_C;   // the temperature lv_J_kg was formed from
   if (MODE == 2 && cpPf > 0.f && Temp_C < 49 && Temp_C > -39) {
      // psyc(t) = cpP / (0.62197 lv(t)), lv(t) = (2500.25 - 2.365 t) 1000: bounds over (-39, 49), rounded outward
      m_lo = 0.88f * __fdividef(cpPf, 0.62197f * 2.5926e6f);
      psyc_hi = 1.001f * __fdividef(cpPf, 0.62197f * 2.3843e6f);
      float t = 0.5f * Tf, G = 0.f, es_a = 0.f, pt_a = 0.f;
#pragma unroll 1
      for (int it = 0; it < 5; it++) {
         float des;
         su_es_single_d(t, pkf, es_a, des);
         const float psy = __fdividef(cpPf, 0.62197f * ((2500.25f - 2.365f * t) * 1000.0f));
         pt_a = psy * (Tf - t);
         G = (es_a - pt_a) - eaf;
         if (it == 4) break;
         float tn = t - __fdividef(G, des + psy);
         tn = fminf(fmaxf(tn, -38.5f), 48.5f);
         if (tn == t) break;
         t = tn;
      }
      // (the loop leaves G, es_a, pt_a evaluated at the final t)
      rf = t;
      // |G(r)| <= |Gf(r)| + e_f, e_f as in level (2): 0.001 + 8e-6 of the terms' magnitudes; + the threshold + slack
      need0 = fabsf(G) + (0.001f + 8e-6f * (fabsf(es_a) + fabsf(pt_a) + fabsf(eaf))) + 0.001f + 2e-4f;
      have_r = (G == G) && (rf > -38.6f) && (rf < 48.6f);
   }
   for (int ii = 1; ii <= 100; ii++) {
      int dec = 2;   // 0: within 0.001, 1: ea_est above, -1: below, 2: not decided yet
      if (MODE != 0 && tw < 49 && tw > -39 && !(fabs(tw) < F(0.001000))) {
         if (MODE == 2 && have_r && t_lv < 49 && t_lv > -39) {
            const float d = (float)(tw - (double)rf);
            const float Dk = psyc_hi * 0.001f * fabsf((float)(tw - t_lv)) * fabsf((float)(Temp_C - tw));
            // 1e-5 K covers the single-precision rounding of d at |tw| < 50
            if (m_lo * (fabsf(d) - 1e-5f) > (need0 + Dk) * 1.01f) dec = (d > 0.f) ? 1 : -1;
#ifdef SU_LATVAP_STATS
            if (dec != 2) su_latvap_stats[0]++;
#endif
         }
         if (dec == 2) {
            const float twf = (float)tw;
            const float es_a = su_es_single(twf, pkf);
            const float pt_a = __fdividef(cpPf, 0.62197f * (float)lv_J_kg) * (Tf - twf);
            const float d = (es_a - pt_a) - eaf;
            const float margin = 0.002f + 8e-6f * (fabsf(es_a) + fabsf(pt_a) + fabsf(eaf));
            if (d > margin) dec = 1;
            else if (d < -margin) dec = -1;
#ifdef SU_LATVAP_STATS
            if (dec != 2) su_latvap_stats[1]++;
#endif
         }
         if (dec != 2) {   // the side effects of the calls not made: warnings 28 (never: tw is in range here) and 19
            if (cp * Press_hPa < 900 || lv_J_kg < 10000) su_warn(dg, 19);
         }
      }
      if (dec == 2) {
#ifdef SU_LATVAP_STATS
         su_latvap_stats[2]++;
#endif
         double es_tw = sat_vap_press_x(dg, tw, Press_hPa, 2);
         double psyc = psyc_const(dg, cp, Press_hPa, lv_J_kg);
         double ea_est = es_tw - psyc * (Temp_C - tw);
         dec = (fabs(ea_est - ea_fix) < F(0.001000)) ? 0 : ((ea_est > ea_fix) ? 1 : ((ea_est < ea_fix) ? -1 : 3));
      }
      t_lv = tw;
      lv_J_kg = (F(2500.25) - F(2.365) * tw) * F(1e3);
      if (switch1 && switch2) {
         incr = incr / F(10.);
         switch1 = false;
         switch2 = false;
      }
      if (dec == 0) {
         break;
      } else if (dec == 1) {
         tw = tw - incr;
         switch1 = true;
      } else if (dec == -1) {
         tw = tw + incr;
         switch2 = true;
      }
   }
   sw = (switch1 ? 1 : 0) | (switch2 ? 2 : 0);
   return lv_J_kg;
}
// The search is a real function: its ~21-trip loop then has the register file to itself instead of spilling inside the
// loop next to the caller's live values (+1.3 %; SU_LATVAP_INLINE restores the inlined form).
#ifndef SU_LATVAP_INLINE
SU_DEVNI double Lat_vap(SuDiag &dg, double Temp_C, double Ea_hPa, double Press_hPa, double cp, int &sw)
#else
SU_INL double Lat_vap(SuDiag &dg, double Temp_C, double Ea_hPa, double Press_hPa, double cp, int &sw)
#endif
{
#if defined(SU_SKIP) && (SU_SKIP & 256)   // timing-only switch (dev_step.cuh)
   return (F(2500.25) - F(2.365) * Temp_C) * 1000;
#endif
   return Lat_vap_t<SU_LATVAP_MODE>(dg, Temp_C, Ea_hPa, Press_hPa, cp, sw);
}
