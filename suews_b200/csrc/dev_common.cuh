// dev_common.cuh -- shared device-side definitions for the SUEWS timestep kernels (sm_100a).
#pragma once
#include "suews_fields.h"
#include <math.h>

// The reference is compiled without -fdefault-real-8, so every un-suffixed literal is REAL(4) and
// is widened at use ("hard parts" 1).  F(c) is that value as a double constant.
#define F(c) ((double)(c##f))
#define SU_MAX(a, b) ((a) > (b) ? (a) : (b))
#define SU_MIN(a, b) ((a) < (b) ? (a) : (b))

#define SU_DEV __device__ __forceinline__
#define SU_DEVNI __device__ __noinline__

enum { PAV = 0, BLDG = 1, CONIF = 2, DECID = 3, GRASS = 4, BSOIL = 5, WATER = 6 };
enum { NSURF = SUEWS_NSURF, NVEG = SUEWS_NVEG };

// Per-cell diagnostics replacing ErrorHint (suews_ctrl_error.f95:1): the first fatal code sticks and
// stops the cell; warning codes (run continues in the reference) are remembered as a bit mask.
struct SuDiag {
   int fatal;                 // 0 = ok
   unsigned long long warn[2];   // bit c set <=> warning code c (< 128) was raised
};
SU_DEV void su_warn(SuDiag &d, int code) { d.warn[(code >> 6) & 1] |= (1ull << (code & 63)); }
SU_DEV void su_fatal(SuDiag &d, int code)
{
   if (d.fatal == 0) d.fatal = code;
}

// Fortran column-major accessors on the flat class arrays
#define SU_PROF24(p, hr, col) ((p)[(hr) + 24 * (col)])
#define SU_LAIPOW(p, i, iv) ((p)[(i) + 4 * (iv)])
#define SU_OHMCOEF(p, is, j, k) ((p)[(is) + 8 * ((j) + 4 * (k))])
#define SU_WDIST(p, i, j) ((p)[(i) + 8 * (j)])
#define SU_SDP(p, i, is) ((p)[(i) + 6 * (is)])

// x**n for constant integer n as gfortran -O3 expands it (optimal powi chains)
SU_DEV double su_p2(double x) { return x * x; }
SU_DEV double su_p3(double x) { return x * (x * x); }
SU_DEV double su_p4(double x) { double a = x * x; return a * a; }
SU_DEV double su_p5(double x) { double a = x * x; return a * (x * a); }
SU_DEV double su_p8(double x) { double a = x * x; a = a * a; return a * a; }

// Per-step quantities that do not depend on the Tsurf iteration (hoisted out of the DO WHILE at
// driver.f95:642, which recomputes all of them on every trip).
struct SuStepInv {
   double dectime, nsh_real, tstep_real;
   int nsh, DLS, dow;
   double VegFraction, ImpervFraction, PervFraction, NonWaterFraction;
   double FAI, zH, z0m, zdm, ZZD;
   double azimuth, zenith_deg;
};
