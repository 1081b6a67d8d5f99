// Host build of the wet-bulb search (suews_b200/csrc/dev_latvap.cuh) for the CPU tests: the three-level search against the
// plain loop on random arguments, with the arguments of the device selftest (suews_b200_selftest_latvap).  The fast device
// intrinsics are replaced by libm's single-precision functions (more accurate, so the margins hold a fortiori); the device
// selftest checks the same code with the real intrinsics.  Test infrastructure.
#include <math.h>
#include <stdint.h>
#define F(c) ((double)(c##f))
#define SU_INL static inline
#define SU_HOT static inline
#define SU_DEVNI static
#define __expf expf
#define __fdividef(a, b) ((a) / (b))
struct SuDiag { int fatal; unsigned long long warn[2]; };
static inline void su_warn(SuDiag &d, int code) { d.warn[(code >> 6) & 1] |= (1ull << (code & 63)); }
#define SU_LATVAP_MODE 2
#define SU_LATVAP_STATS 1
static long long su_latvap_stats[3];   // trips decided by the root bound / the single-precision screen / in double
#include "../suews_b200/csrc/dev_latvap.cuh"

// returns the number of mismatches between the plain loop and the three-level search over n random argument sets
extern "C" long long latvap_host_selftest(long long n, unsigned long long seed, long long *trips_out)
{
   unsigned long long st = seed * 0x9E3779B97F4A7C15ull + 1;
   auto next = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
   auto unif = [&]() { return (double)(next() >> 11) * 0x1p-53; };
   long long bad = 0;
   for (long long i = 0; i < n; i++) {
      const int kind = (int)(next() & 7);
      double T = (kind == 0) ? -60.0 + 130.0 * unif() : ((kind == 1) ? (unif() - 0.5) * 0.01 : -35.0 + 80.0 * unif());
      double P = (kind == 2) ? 500.0 + 700.0 * unif() : 950.0 + 100.0 * unif();
      double rh = (kind == 3) ? 100.0 * unif() * unif() * unif() : 100.0 * unif();
      if (kind == 4) rh = 100.0 - 1e-3 * unif();
      SuDiag d0 = {0, {0, 0}}, d1 = d0;
      double Tx = su_fix_T(T);
      const double es = sat_vap_press_v(Tx, P);
      const double ea = rh / 100 * es;
      const double cp = 1000.0 + 60.0 * unif();
      int s0 = (int)(next() & 3), s1 = s0;
      long long keep[3] = {su_latvap_stats[0], su_latvap_stats[1], su_latvap_stats[2]};
      const double a = Lat_vap_t<0>(d0, T, ea, P, cp, s0);
      for (int k = 0; k < 3; k++) su_latvap_stats[k] = keep[k];   // (the plain loop's trips are not part of the statistics)
      const double b = Lat_vap_t<2>(d1, T, ea, P, cp, s1);
      uint64_t ba, bb;
      __builtin_memcpy(&ba, &a, 8); __builtin_memcpy(&bb, &b, 8);
      if (ba != bb || s0 != s1 || d0.warn[0] != d1.warn[0] || d0.warn[1] != d1.warn[1]) bad++;
   }
   if (trips_out)
      for (int k = 0; k < 3; k++) trips_out[k] = su_latvap_stats[k];
   return bad;
}
