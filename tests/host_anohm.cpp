// Host build of the device-side AnOHM fit (suews_b200/csrc/dev_anohm.cuh) for the CPU tests: the same source the kernel
// compiles, with the device qualifiers defined away, so tests/test_anohm_oracle.py can compare it with the oracle without
// a GPU.  Test infrastructure; built by the test into a temporary directory.
#define SU_ANOHM_HOST 1
#ifdef SU_ANOHM_ULP
// -DSU_ANOHM_ULP: every sin / cos / atan / exp result moved by one ulp, up or down, for a pseudo-random two thirds of the
// arguments -- what another libm may return.  The test measures how far the fitted coefficients move (the tolerance of the
// GPU parity tests for this scheme rests on that number).
#include <math.h>
#include <stdint.h>
#include <string.h>
static inline double ulp_shift(double r, double x)
{
   uint64_t b;
   memcpy(&b, &x, 8);
   b ^= b >> 29; b *= 0x9E3779B97F4A7C15ull; b ^= b >> 32;
   const int k = (int)(b % 3);
   return k == 0 ? r : nextafter(r, k == 1 ? INFINITY : -INFINITY);
}
static inline double p_sin(double x) { return ulp_shift(sin(x), x); }
static inline double p_cos(double x) { return ulp_shift(cos(x), x); }
static inline double p_atan(double x) { return ulp_shift(atan(x), x); }
static inline double p_exp(double x) { return ulp_shift(exp(x), x); }
#define sin p_sin
#define cos p_cos
#define atan p_atan
#define exp p_exp
#endif
#include "../suews_b200/csrc/dev_anohm.cuh"
#include <vector>

extern "C" void an_host_day_coefs(const double *hours8, int nrows, int xid, double dTemp_C, double dRH, int EmissionsMethod,
                                  double qf, const double *moist, const double *alb, const double *emis, const double *cp,
                                  const double *kk, const double *ch, double *coeff21)
{
   // hours8: [nrows][8] rows (id, tHr, Sd, Ta, RH, pres, WS, qf_obs) in block order; grouped by day here as the library does
   std::vector<SuewsHourRec> hours;
   std::vector<int> first(367, 0), count(367, 0), nsun(367, 0);
   for (int d = 1; d <= 366; d++) {
      first[d] = (int)hours.size();
      for (int r = 0; r < nrows; r++) {
         const double *h = hours8 + (size_t)r * 8;
         if ((int)h[0] == d) {
            SuewsHourRec q = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]};
            hours.push_back(q);
            count[d]++;
            if (h[2] > 0) nsun[d]++;
         }
      }
   }
   SuAnohmIn in = {hours.data(), first.data(), count.data(), nsun.data(), nullptr, 0};
   an_day_coefs(in, xid, dTemp_C, dRH, EmissionsMethod, qf, moist, alb, emis, cp, kk, ch, coeff21);
}
