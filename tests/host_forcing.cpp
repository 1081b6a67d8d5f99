// Host build of the device-side forcing disaggregation (suews_b200/csrc/forcing_kernels.cuh: MetRead's unit conversion,
// DisaggregateMet / DisaggregateESTM) for the CPU tests: the kernel's body run once per row with the thread index set by
// hand, the argument record filled as suews_b200_set_forcing_orig / set_estm_forcing_orig fill it.  Test infrastructure.
#include <math.h>
#include <stddef.h>
#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline
struct HostDim { unsigned x, y, z; };
static HostDim blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, threadIdx = {0, 0, 0};
#include "../suews_b200/csrc/forcing_kernels.cuh"

// slow: [n x ncol] column-major (ncol 24: met forcing; 13: ESTM surface temperatures); fast: [n * nper x ncol] column-major
extern "C" void host_disagg(const double *slow, int n, int ncol, int resolution_in_s, int tstep, int disagg_method, int raw_units,
                            int smd_method, double *fast)
{
   SuDisaggArgs a;
   a.slow = slow; a.fast = fast; a.n = n; a.nper = resolution_in_s / tstep; a.tstep = tstep; a.ncol = ncol;
   a.raw_units = raw_units; a.smd_method = smd_method;
   const bool met = (ncol == SUEWS_NCOL_MET);
   a.rain_col = met ? 13 : -1; a.snow_col = met ? 15 : -1; a.skip_col = met ? 23 : -1;
   a.all_missing = 0; a.type20 = 0;
   for (int c = 4; c < ncol; c++) {
      bool all = true;
      for (int i = 0; i < n && all; i++) all = (slow[(size_t)i + (size_t)n * c] == -999.0);
      if (met && c == 19 && smd_method == 0) all = true;
      if (all) a.all_missing |= (1u << c);
      if (met ? (disagg_method == 2 || (disagg_method == 3 && c >= 9 && c <= 12)) : (disagg_method == 2)) a.type20 |= (1u << c);
   }
   const long long nfast = (long long)n * a.nper;
   for (long long r = 0; r < nfast; r++) {
      blockIdx.x = (unsigned)r;
      suews_disagg_kernel(a);
   }
}
