// dev_common.cuh -- shared device-side definitions for the SUEWS timestep kernels (sm_100a).
#pragma once
#include "suews_fields.h"
#include <math.h>

// The reference is compiled without -fdefault-real-8, so every un-suffixed literal is REAL(4) and
// is widened at use ("hard parts" 1).  F(c) is that value as a double constant.
#define F(c) ((double)(c##f))
#define SU_MAX(a, b) ((a) > (b) ? (a) : (b))
#define SU_MIN(a, b) ((a) < (b) ? (a) : (b))

// Code size is a first-order performance parameter here: one cell-step is ~45 k SASS instructions when
// everything is inlined (726 KB of text), far beyond the SM instruction caches, and the first profile
// showed 69 % of warp stalls on instruction fetch (profiles/r1_ncu_notes.md).  So physics routines are
// real functions (SU_DEV = noinline), the double-precision libm entry points are shared out-of-line
// wrappers, and only leaf helpers are inlined (SU_INL).
#define SU_INL __device__ __forceinline__
// (static: the headers are compiled into two translation units, suews_b200.cu and suews_kernel_small.cu)
#define SU_DEVNI static __device__ __noinline__
#if defined(SU_INLINE_ALL) || defined(SU_DEV_INLINE)   // SU_DEV_INLINE: physics routines inlined, libm wrappers still shared
#define SU_DEV __device__ __forceinline__
#else
#define SU_DEV static __device__ __noinline__
#endif

// routines the default scheme set never enters (ESTM, biogenic CO2, ice saturation pressure): out of line in every build,
// so they take no room in the instruction stream of the hot phases
#define SU_COLD static __device__ __noinline__
// the large routines that are called once or twice per step (tuning switch SU_BIG_INLINE)
#if defined(SU_BIG_INLINE) || defined(SU_INLINE_ALL)
#define SU_BIG __device__ __forceinline__
#else
#define SU_BIG static __device__ __noinline__
#endif
// per-routine switch: SU_BIG_MASK bit k-1 inlines group k (1 cal_Stab, 2 cal_HorizontalSoilWater, 3 NARP + sun position,
// 4 RSLProfile, 5 cal_psim_hat_zh)
#ifndef SU_BIG_MASK
#define SU_BIG_MASK 0
#endif
#if (SU_BIG_MASK & 1)
#define SU_BIG1 __device__ __forceinline__
#else
#define SU_BIG1 SU_BIG
#endif
#if (SU_BIG_MASK & 2)
#define SU_BIG2 __device__ __forceinline__
#else
#define SU_BIG2 SU_BIG
#endif
#if (SU_BIG_MASK & 4)
#define SU_BIG3 __device__ __forceinline__
#else
#define SU_BIG3 SU_BIG
#endif
#if (SU_BIG_MASK & 8)
#define SU_BIG4 __device__ __forceinline__
#else
#define SU_BIG4 SU_BIG
#endif
#if (SU_BIG_MASK & 16)
#define SU_BIG5 __device__ __forceinline__
#else
#define SU_BIG5 SU_BIG
#endif
#if defined(SU_HOT_INLINE) || defined(SU_INLINE_ALL)
#define SU_HOT SU_INL
#else
#define SU_HOT SU_DEV
#endif

#ifndef SU_INLINE_ALL
#include "dev_math.cuh"
// tuning switches (tools/variants.sh): SU_EXPLOG_INLINE inlines exp/log at their call sites, SU_HOT_INLINE does the
// same for the small physics helpers that are called dozens of times per step (SU_HOT)
#ifdef SU_EXPLOG_INLINE
#define SU_EXPLOG_ATTR SU_INL
#else
#define SU_EXPLOG_ATTR SU_DEVNI
#endif
#ifdef SU_LIBM_EXPLOG
SU_EXPLOG_ATTR double su_exp(double x) { return exp(x); }
SU_EXPLOG_ATTR double su_log(double x) { return log(x); }
#else
SU_EXPLOG_ATTR double su_exp(double x) { return su_exp_fast(x); }
SU_EXPLOG_ATTR double su_log(double x) { return su_log_fast(x); }
#endif
SU_DEVNI double su_pow_general(double x, double y) { return pow(x, y); }
// x**0.5, x**0.25, x**(-0.25), x**(-0.5), x**1.5 -- the exponents of the stability functions, NARP's
// Tsurf and van Genuchten -- go through DSQRT / rsqrt (<= 1.25 ulp) instead of the ~400-instruction general
// pow: the second profile had pow at 45 % of all executed instructions.  Literal exponents fold at compile
// time; run-time exponents (LAIPower, drainage coefficients, ...) take the same route when they match.
// x^(-1/4), the unstable-side phi of every stability scheme: single-precision seed (two MUFU ops) and two
// Newton steps y <- y (5 - x y^4) / 4 in double (quadratic: 2^-22 -> 1e-13 -> rounding), <= 2 ulp, a third of the
// instructions of rsqrt(sqrt(x)).  Outside the single-precision range the exact pair is used.
SU_INL double su_rqrt4(double x)
{
   if (!(x > 1e-30 && x < 1e30)) return rsqrt(sqrt(x));
   double y = (double)rsqrtf(sqrtf((float)x));
   const double qx = 0.25 * x;
   double y2 = y * y;
   y = y * __fma_rn(-qx, y2 * y2, 1.25);
   y2 = y * y;
   y = y * __fma_rn(-qx, y2 * y2, 1.25);
   return y;
}
// 1 / d for a normal d of ordinary magnitude: hardware seed + two Newton steps, <= 1 ulp (not correctly rounded;
// used only inside the psi-hat integrations of dev_rsl.cuh, see DESIGN.md "numerics")
SU_INL double su_rcp_fast(double d)
{
   double r;
   asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
   double e = __fma_rn(-d, r, 1.0);
   r = __fma_rn(r, e, r);
   e = __fma_rn(-d, r, 1.0);
   r = __fma_rn(r, e, r);
   return r;
}
// Correctly rounded x^(1/4) for a normal positive x: sqrt(sqrt(x)) (two roundings, ~0.75 ulp) corrected by one
// Newton step whose residual y^4 - x is formed in double-double.  Used where a caller raises the root back to the
// fourth power and subtracts the original (NARP's snow surface temperature): that difference is pure rounding
// residue, the reference branches on its sign, and glibc's pow -- which the oracle and the reference use -- is
// correctly rounded in all but ~4 % of arguments.
SU_INL double su_root4_cr(double x)
{
   double y = sqrt(sqrt(x));
   for (int it = 0; it < 2; it++) {
      const double h1 = y * y, l1 = __fma_rn(y, y, -h1);
      const double h2 = h1 * h1, l2 = __fma_rn(h1, h1, -h2) + 2.0 * h1 * l1;
      const double e = (h2 - x) + l2;              // y^4 - x
      const double yn = y - e / (4.0 * h1 * y);    // rounds to the double nearest the corrected root
      if (yn == y) break;
      y = yn;
   }
   return y;
}
SU_INL double su_pow(double x, double y)
{
   if (y == 0.5) return sqrt(x);
   if (y == 0.25) return sqrt(sqrt(x));
   if (y == -0.25) return su_rqrt4(x);
   if (y == -0.5) return rsqrt(x);
   if (y == 1.5) return x * sqrt(x);
   return su_pow_general(x, y);
}
SU_DEVNI double su_sin(double x) { return sin(x); }
SU_DEVNI double su_cos(double x) { return cos(x); }
SU_DEVNI double su_tan(double x) { return tan(x); }
SU_DEVNI double su_atan(double x) { return atan(x); }
SU_DEVNI double su_atan2(double y, double x) { return atan2(y, x); }
SU_DEVNI double su_asin(double x) { return asin(x); }
SU_DEVNI double su_acos(double x) { return acos(x); }
SU_DEVNI double su_tanh(double x) { return tanh(x); }
// from here on the physics headers call the shared wrappers (undone at the end of dev_step.cuh)
#define exp su_exp
#define log su_log
#define pow su_pow
#define sin su_sin
#define cos su_cos
#define tan su_tan
#define atan su_atan
#define atan2 su_atan2
#define asin su_asin
#define acos su_acos
#define tanh su_tanh
#endif

enum { PAV = 0, BLDG = 1, CONIF = 2, DECID = 3, GRASS = 4, BSOIL = 5, WATER = 6 };
enum { NSURF = SUEWS_NSURF, NVEG = SUEWS_NVEG };

// Per-cell diagnostics replacing ErrorHint (suews_ctrl_error.f95:1): the first fatal code sticks and
// stops the cell; warning codes (run continues in the reference) are remembered as a bit mask.
struct SuDiag {
   int fatal;                 // 0 = ok
   unsigned long long warn[2];   // bit c set <=> warning code c (< 128) was raised
};
SU_INL void su_warn(SuDiag &d, int code) { d.warn[(code >> 6) & 1] |= (1ull << (code & 63)); }
SU_INL void su_fatal(SuDiag &d, int code)
{
   if (d.fatal == 0) d.fatal = code;
}

// Fortran column-major accessors on the flat class arrays
#define SU_PROF24(p, hr, col) ((p)[(hr) + 24 * (col)])
#define SU_LAIPOW(p, i, iv) ((p)[(i) + 4 * (iv)])
#define SU_OHMCOEF(p, is, j, k) ((p)[(is) + 8 * ((j) + 4 * (k))])
#define SU_WDIST(p, i, j) ((p)[(i) + 8 * (j)])
#define SU_SDP(p, i, is) ((p)[(i) + 6 * (is)])

// x / y where x is often exactly zero (dry stores, no snow, night, RH = 100 %): CUDA's FP64 division sends a
// zero (or subnormal) numerator down its out-of-line slow path (~60 instructions and a stack frame), 56 times
// per cell-step in the first profiles.  0 / y is answered directly, with the sign IEEE gives it.
SU_INL double su_div0(double x, double y)
{
   if (x == 0.0 && y == y && y != 0.0)
      return __longlong_as_double((__double_as_longlong(x) ^ __double_as_longlong(y)) & (long long)0x8000000000000000ull);
   return x / y;
}

// x**n for constant integer n as gfortran -O3 expands it (optimal powi chains)
SU_INL double su_p2(double x) { return x * x; }
SU_INL double su_p3(double x) { return x * (x * x); }
SU_INL double su_p4(double x) { double a = x * x; return a * a; }
SU_INL double su_p5(double x) { double a = x * x; return a * (x * a); }
SU_INL double su_p8(double x) { double a = x * x; a = a * a; return a * a; }

// Per-step quantities that do not depend on the Tsurf iteration (hoisted out of the DO WHILE at
// driver.f95:642, which recomputes all of them on every trip).
struct SuStepInv {
   double dectime, nsh_real, tstep_real;
   int nsh, DLS, dow;
   double VegFraction, ImpervFraction, PervFraction, NonWaterFraction;
   double FAI, zH, z0m, zdm, ZZD;
   double azimuth, zenith_deg;
};
