// forcing_kernels.cuh -- met-forcing preparation on the device: the step immediately before the hot path.
//
//   suews_disagg_kernel       MetRead's unit conversion (suews_ctrl_input.f95:96-109) and DisaggregateMet (:1380-1640):
//                             DisaggregateDateTime :1853-1900, Disagg_Lin :1908-2011 (type 10 period averages,
//                             type 20 instantaneous), DisaggP_amongN :2016-2068 for RainDisaggMethod = 100
//   suews_pack_steps_kernel   the per-step unpacking of MetForcingBlock (suews_ctrl_driver.f95:3177-3201) with
//                             SUEWS_cal_dectime (util_time.f95:270) and SUEWS_cal_weekday (:301), once per step
//
// One thread per model-timestep row.  Reads the original-resolution block (12x smaller than the result for hourly
// input and 5-min steps), writes the [nfast x 24] MetForcingBlock column-major -- coalesced -- in HBM, where the
// timestep kernel's step records are then built without the block ever crossing PCIe.
// The arithmetic is the reference's, expression by expression (REAL(4) sub-expressions included), so the block is
// bit-identical to the array the Fortran binary holds in memory; suews_b200/forcing.py is the same algorithm in
// numpy and the reference goldens are reproduced from its output (tests/test_oracle_golden.py).
#pragma once
#include "suews_fields.h"

struct SuDisaggArgs {
   const double *slow;   // [n x ncol] column-major, as read from the file
   double *fast;         // [n*nper x ncol] column-major
   int n, nper, tstep;
   int ncol;             // 24: met forcing (DisaggregateMet); 13: ESTM surface temperatures (DisaggregateESTM, :1650-1849)
   int rain_col, snow_col, skip_col;   // 0-based special columns of the met block (13, 15, 23), -1 when there are none
   int raw_units;        // 1: column 13 (pressure) is in kPa -> hPa (ctrl_input.f95:109)
   int smd_method;       // 0: xsmd = -999 (ctrl_input.f95:96-102)
   unsigned all_missing; // bit c set: column c of the input is -999 throughout (ctrl_input.f95:1524, :1545)
   unsigned type20;      // bit c set: column c is "instantaneous" (DisaggMethod 2 / 3, ctrl_input.f95:1425-1436)
};

__host__ __device__ inline long long su_nint(double x)
{   // Fortran NINT: half away from zero
   return (x >= 0) ? (long long)floor(x + 0.5) : -(long long)floor(-x + 0.5);
}
__host__ __device__ inline int su_days_in_year(int y)
{
   return (((y % 4 == 0) && (y % 100 != 0)) || (y % 400 == 0)) ? 366 : 365;
}

__device__ __forceinline__ double su_slow_at(const SuDisaggArgs &a, int i, int c)
{
   double v = a.slow[(size_t)i + (size_t)a.n * c];
   if (a.ncol == SUEWS_NCOL_MET) {
      if (c == 12 && a.raw_units) v = v * 10.0;       // Pres_hPa = Pres_kPa*10.
      if (c == 19 && a.smd_method == 0) v = -999.0;   // no soil-moisture observations
   }
   return v;
}

__global__ void suews_disagg_kernel(SuDisaggArgs a)
{
   const int nper = a.nper, n = a.n;
   const long long nfast = (long long)n * nper;
   const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
   if (r >= nfast) return;
   const int i0 = (int)(r / nper);        // original row this sub-interval belongs to
   const int k = (int)(r % nper);         // 0 .. nper-1, the last one carries the original time stamp
   // ---- DisaggregateDateTime
   {
      const double iy0 = a.slow[(size_t)i0], id0 = a.slow[(size_t)i0 + (size_t)n], it0 = a.slow[(size_t)i0 + 2 * (size_t)n],
                   im0 = a.slow[(size_t)i0 + 3 * (size_t)n];
      const double dectime_orig = (id0 - 1) + it0 / 24.0 + im0 / (60.0 * 24.0);
      const float step = ((float)a.tstep / 60.0f) / (60.0f * 24.0f);
      const float off = step * (float)(nper - 1 - k);
      double td = dectime_orig - (double)off;
      td = (double)((float)su_nint(td * 60 * 60 * 24) / 86400.0f);
      int iy = (td < 0) ? (int)iy0 - 1 : (int)iy0;
      if (td < 0) td = td + su_days_in_year(iy);
      const int idd = (int)floor(td) + 1;
      const long long ihm = su_nint((td + 1 - (double)((float)idd / 1.0f)) * 60.0 * 24.0);
      long long ih = (ihm - ihm % 60) / 60;
      if (!(ih < 24)) ih = 0;
      const long long im = ihm % 60;
      a.fast[(size_t)r] = (double)iy;
      a.fast[(size_t)r + (size_t)nfast] = (double)idd;
      a.fast[(size_t)r + 2 * (size_t)nfast] = (double)ih;
      a.fast[(size_t)r + 3 * (size_t)nfast] = (double)im;
   }
   const int xn = (nper % 2 == 0) ? 2 : 1;
   const int half = nper / 2;
   for (int c = 4; c < a.ncol; c++) {
      double v = -999.0;
      const bool missing = (a.all_missing >> c) & 1u;
      if (c == a.rain_col) {   // rain: spread evenly over the nper sub-intervals (DisaggP_amongN with amongN = Nper)
         const double s = su_slow_at(a, i0, c);
         v = (s > 0) ? s / nper : 0.0;
      } else if (c == a.snow_col && missing) {
         v = 0.0;                // an all -999 snow column ends as 0 (ctrl_input.f95:1507-1508, :1611-1613)
      } else if (c == a.skip_col || missing) {
         v = -999.0;             // wind direction is not disaggregated; missing columns stay missing
         // (a snow column that is present is first spread like rain inside the ii == 14 branch and then overwritten
         //  by Disagg_Lin when the loop reaches ii == 16, ctrl_input.f95:1503-1563: only the second result survives)
      } else if (!((a.type20 >> c) & 1u)) {
         // type 10: the original value is the average of the period that ends at its time stamp
         long long q = r - half;
         int i, s;
         if (q < 0) { i = 0; s = 1; }                                                   // first half period: repeated
         else if (q >= (long long)nper * (n - 1)) { i = n - 2; s = nper; }              // last half period: repeated
         else { i = (int)(q / nper); s = (int)(q % nper) + 1; }
         const double s0 = su_slow_at(a, i, c), d = su_slow_at(a, i + 1, c) - s0;
         v = s0 - d / (xn * nper) + d / nper * s;
      } else {
         // type 20: the original value is instantaneous at its time stamp
         long long q = r - nper;
         int i, s;
         if (q < 0) { i = 0; s = 1; }
         else { i = (int)(q / nper); s = (int)(q % nper) + 1; }
         const double s0 = su_slow_at(a, i, c), d = su_slow_at(a, i + 1, c) - s0;
         v = (s0 + d / nper * 2 * (s - 1) + s0) / 2;
      }
      a.fast[(size_t)r + (size_t)nfast * c] = v;
   }
}
