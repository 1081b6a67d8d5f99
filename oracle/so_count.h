// so_count.h -- the oracle's op counter.  TEST INFRASTRUCTURE like the rest of oracle/.
// `make count` compiles the same C sources as C++ with this header force-included: `double` becomes a class that holds one
// double and counts every add/subtract, multiply, divide and libm call made on it.  Layout, arithmetic and results are
// those of the plain build (the class is 8 bytes, every operator forwards to the built-in one); only the counters move.
// Weights applied by the reader (tools/oracle_opcount.py), as SURVEY 8(d) lists them: add/mul/div/sqrt 1, exp 20, log 25,
// pow 60, sin/cos/tan 30, atan/asin/acos/atan2 35, tanh 30.
#ifndef SO_COUNT_H
#define SO_COUNT_H
#define SO_OPCOUNT 1
#include <stddef.h>
#include <cmath>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <float.h>
#include <stdint.h>
#include <cstring>
#include <cstdlib>
#include <cstdio>
extern "C" unsigned long long *so_opcount_cur;
extern "C" int so_opcount_on;
enum { OC_ADD = 0, OC_MUL, OC_DIV, OC_SQRT, OC_EXP, OC_LOG, OC_POW, OC_SINCOS, OC_ATAN, OC_TANH, OC_CMP, OC_FLOOR };
#define OC(k) do { if (so_opcount_on) so_opcount_cur[k]++; } while (0)
struct cdouble {
   double v;
   cdouble() = default;
   cdouble(double x) : v(x) {}
   cdouble(float x) : v(x) {}
   cdouble(int x) : v(x) {}
   cdouble(long x) : v((double)x) {}
   cdouble(long long x) : v((double)x) {}
   cdouble(unsigned x) : v(x) {}
   cdouble(unsigned long x) : v((double)x) {}
   explicit operator double() const { return v; }
   explicit operator float() const { return (float)v; }
   explicit operator int() const { return (int)v; }
   explicit operator long() const { return (long)v; }
   explicit operator bool() const { return v != 0; }
   cdouble &operator+=(cdouble o) { OC(OC_ADD); v += o.v; return *this; }
   cdouble &operator-=(cdouble o) { OC(OC_ADD); v -= o.v; return *this; }
   cdouble &operator*=(cdouble o) { OC(OC_MUL); v *= o.v; return *this; }
   cdouble &operator/=(cdouble o) { OC(OC_DIV); v /= o.v; return *this; }
};
#define OC_BIN(op, k) \
   inline cdouble operator op(cdouble a, cdouble b) { OC(k); return cdouble(a.v op b.v); } \
   inline cdouble operator op(cdouble a, double b) { OC(k); return cdouble(a.v op b); } \
   inline cdouble operator op(double a, cdouble b) { OC(k); return cdouble(a op b.v); } \
   inline cdouble operator op(cdouble a, int b) { OC(k); return cdouble(a.v op b); } \
   inline cdouble operator op(int a, cdouble b) { OC(k); return cdouble(a op b.v); } \
   inline cdouble operator op(cdouble a, float b) { OC(k); return cdouble(a.v op (double)b); } \
   inline cdouble operator op(float a, cdouble b) { OC(k); return cdouble((double)a op b.v); } \
   inline cdouble operator op(cdouble a, long b) { OC(k); return cdouble(a.v op (double)b); } \
   inline cdouble operator op(long a, cdouble b) { OC(k); return cdouble((double)a op b.v); }
OC_BIN(+, OC_ADD) OC_BIN(-, OC_ADD) OC_BIN(*, OC_MUL) OC_BIN(/, OC_DIV)
inline cdouble operator-(cdouble a) { return cdouble(-a.v); }
inline cdouble operator+(cdouble a) { return a; }
#define OC_CMP_OP(op) \
   inline bool operator op(cdouble a, cdouble b) { return a.v op b.v; } \
   inline bool operator op(cdouble a, double b) { return a.v op b; } \
   inline bool operator op(double a, cdouble b) { return a op b.v; } \
   inline bool operator op(cdouble a, int b) { return a.v op b; } \
   inline bool operator op(int a, cdouble b) { return a op b.v; } \
   inline bool operator op(cdouble a, float b) { return a.v op (double)b; } \
   inline bool operator op(float a, cdouble b) { return (double)a op b.v; }
OC_CMP_OP(<) OC_CMP_OP(>) OC_CMP_OP(<=) OC_CMP_OP(>=) OC_CMP_OP(==) OC_CMP_OP(!=)
inline bool operator!(cdouble a) { return a.v == 0; }
#define OC_F1(name, k) inline cdouble oc_##name(cdouble x) { OC(k); return cdouble(std::name(x.v)); }
OC_F1(sqrt, OC_SQRT) OC_F1(exp, OC_EXP) OC_F1(log, OC_LOG) OC_F1(sin, OC_SINCOS) OC_F1(cos, OC_SINCOS) OC_F1(tan, OC_SINCOS)
OC_F1(atan, OC_ATAN) OC_F1(asin, OC_ATAN) OC_F1(acos, OC_ATAN) OC_F1(tanh, OC_TANH)
inline cdouble oc_fabs(cdouble x) { return cdouble(std::fabs(x.v)); }
inline cdouble oc_floor(cdouble x) { return cdouble(std::floor(x.v)); }
inline cdouble oc_pow(cdouble x, cdouble y) { OC(OC_POW); return cdouble(std::pow(x.v, y.v)); }
inline cdouble oc_atan2(cdouble y, cdouble x) { OC(OC_ATAN); return cdouble(std::atan2(y.v, x.v)); }
inline cdouble oc_fmod(cdouble x, cdouble y) { OC(OC_DIV); return cdouble(std::fmod(x.v, y.v)); }
inline cdouble oc_fmax(cdouble x, cdouble y) { return cdouble(std::fmax(x.v, y.v)); }
inline cdouble oc_fmin(cdouble x, cdouble y) { return cdouble(std::fmin(x.v, y.v)); }
inline int oc_isnan(cdouble x) { return std::isnan(x.v); }
inline cdouble copysign(double a, cdouble b) { return cdouble(std::copysign(a, b.v)); }
inline cdouble copysign(cdouble a, cdouble b) { return cdouble(std::copysign(a.v, b.v)); }
inline cdouble round(cdouble x) { return cdouble(std::round(x.v)); }
#define sqrt oc_sqrt
#define exp oc_exp
#define log oc_log
#define sin oc_sin
#define cos oc_cos
#define tan oc_tan
#define atan oc_atan
#define asin oc_asin
#define acos oc_acos
#define tanh oc_tanh
#define fabs oc_fabs
#define floor oc_floor
#define pow oc_pow
#define atan2 oc_atan2
#define fmod oc_fmod
#define fmax oc_fmax
#define fmin oc_fmin
#define double cdouble
#endif
