// suews_b200.cu -- kernels and C ABI of libsuews_b200.so (sm_100a).
//
// Data layout in HBM:
//   classes   SuewsClassParams[nclass]            AoS, broadcast reads (cells are grouped by class)
//   cellp     double[SUEWS_N_CELL_DBL][stride]    SoA, coalesced
//   state     double[SUEWS_N_STATE_DBL][stride]   SoA, coalesced; loaded once per launch into the thread
//   steps     SuewsStepRec[nsteps]                clock + forcing, shared by all cells; staged into shared
//                                                 memory in chunks by cp.async.bulk (TMA) + mbarrier
//   ephem     SuewsEphem[ntz][nsteps]             time-only half of the solar position, one thread / step
//   sums      double[82][stride]                  per-cell running sums of the SUEWS output line
// One thread owns one cell and advances it through all requested timesteps of a launch.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <strings.h>
#include <algorithm>
#include <string>
#include <chrono>
#include <vector>

#include "suews_fields.h"
#include "dev_step.cuh"
#include "forcing_kernels.cuh"
#include "../../include/suews_b200.h"

#include "suews_kernel.cuh"

// clock decode of one forcing row (SUEWS_cal_dectime util_time.f95:270, SUEWS_cal_weekday :301)
__host__ __device__ static void decode_clock(SuewsStepRec &r)
{
   r.dectime = (double)(r.id - 1) + (double)r.it / 24 + (double)r.imin / (60 * 24) + (double)0 / (60 * 60 * 24);
   int year = r.iy, b = r.id, k, mb, md;
   k = ((year % 4 == 0 && year % 100 != 0) || year % 400 == 0) ? 1 : 0;   // MODULO(year, p) == 0 <=> year % p == 0
   if (b <= 31) { mb = 1; md = b; }
   else if (b <= 59 + k) { mb = 2; md = b - 31; }
   else if (b <= 90 + k) { mb = 3; md = b - (59 + k); }
   else if (b <= 120 + k) { mb = 4; md = b - (90 + k); }
   else if (b <= 151 + k) { mb = 5; md = b - (120 + k); }
   else if (b <= 181 + k) { mb = 6; md = b - (151 + k); }
   else if (b <= 212 + k) { mb = 7; md = b - (181 + k); }
   else if (b <= 243 + k) { mb = 8; md = b - (212 + k); }
   else if (b <= 273 + k) { mb = 9; md = b - (243 + k); }
   else if (b <= 304 + k) { mb = 10; md = b - (273 + k); }
   else if (b <= 334 + k) { mb = 11; md = b - (304 + k); }
   else { mb = 12; md = b - (334 + k); }
   int yr = year, mn = mb;
   if (!(mn > 2)) { mn += 12; yr -= 1; }
   int n1 = (26 * (mn + 1)) / 10, n2 = (125 * yr) / 100;
   int day = (md + n1 + n2 - (yr / 100) + (yr / 400) - 1);
   r.dow = day % 7 + 1;
   r.month = mb;
   r.pad0_ = 0;
   r.pad1_ = 0;
}


// MetForcingBlock row -> step record (driver.f95:3177-3201), one thread per row
__host__ __device__ static void pack_step(SuewsStepRec &r, const double *met, size_t n, size_t i)
{
#define MET(col) met[i + n * (col)]
   r.iy = (int)MET(0); r.id = (int)MET(1); r.it = (int)MET(2); r.imin = (int)MET(3);
   r.qn1_obs = MET(4); r.qs_obs = MET(7); r.qf_obs = MET(8); r.avU1 = MET(9); r.avRh = MET(10); r.Temp_C = MET(11);
   r.Press_hPa = MET(12); r.Precip = MET(13); r.avkdn = MET(14); r.snowFrac_obs = MET(15); r.ldown_obs = MET(16);
   r.fcld_obs = MET(17); r.wu_m3 = MET(18); r.xsmd = MET(19); r.LAI_obs = MET(20);
#undef MET
   decode_clock(r);
}
__global__ void suews_pack_steps_kernel(const double *met, int nsteps, SuewsStepRec *recs)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= nsteps) return;
   SuewsStepRec r;
   pack_step(r, met, (size_t)nsteps, (size_t)i);
   recs[i] = r;
}

// Time-only half of the solar position + nothing else: one thread per (time zone, step).
__global__ void suews_ephem_kernel(const SuewsStepRec *steps, int nsteps, const double *tz_list, int ntz, int tstep,
                                   SuewsEphem *ephem)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= nsteps * ntz) return;
   int s = i % nsteps, z = i / nsteps;
   // dectime [days] minus the INTEGER tstep/2 [seconds], exactly as the reference call (driver.f95:689)
   spa_time_part((double)steps[s].iy, steps[s].dectime - (double)(tstep / 2), tz_list[z], ephem[(size_t)z * nsteps + s]);
}

// Final gather, per device: selected columns of the slot-ordered running sums -> a compact cell-ordered block
// [ncols][n] (reads are contiguous per warp; the writes scatter through the slot -> cell map)
struct SuPackCols { int col[16]; int ncols; };
__global__ void suews_pack_sums_kernel(const double *sums, const int *cell_of_slot, long long stride, int n, SuPackCols pc,
                                       double *out)
{
   const int sl = blockIdx.x * blockDim.x + threadIdx.x;
   if (sl >= n) return;
   const int c = cell_of_slot[sl];
   for (int k = 0; k < pc.ncols; k++) out[(size_t)k * n + c] = sums[(size_t)pc.col[k] * stride + sl];
}

// KdownZen = 1 (DisaggregateMet, suews_ctrl_input.f95:1568-1600): the disaggregated K-down is set to zero wherever the sun is
// below the horizon at the middle of the step before (zenith > 90 deg at dectime - halftimestep) ...
__global__ void suews_kdownzen_mask_kernel(const double *fast, long long nfast, int tstep, double lat, double lng, double tz,
                                           double alt, double *adj)
{
   const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
   if (i >= nfast) return;
   const double id = fast[(size_t)i + (size_t)nfast], it = fast[(size_t)i + 2 * (size_t)nfast], im = fast[(size_t)i + 3 * (size_t)nfast];
   const double dectimeFast = id + it / 24.0 + im / (double)(60.0f * 24.0f);
   const double halftimestep = (double)(((float)tstep * 1.0f) / 2 / (24 * 3600));   // REAL(tstep_real)/2/(24*3600), ctrl_init.f95:328
   SuewsEphem e;
   spa_time_part(id, dectimeFast - halftimestep, tz, e);   // the reference passes the day of year where the year belongs (:1581)
   double az, zen;
   spa_location_part(e, lat, lng, alt, az, zen);
   const double k = fast[(size_t)i + 14 * (size_t)nfast];
   adj[i] = (zen > 90) ? 0.0 : k;
}
// ... and each block of nsd rows (a day's worth, counted from the first row) is rescaled to keep its sum; rows after the last
// whole block keep the masked values.  One thread per block: the sums are the reference's sequential SUMs.
__global__ void suews_kdownzen_renorm_kernel(double *kdown, const double *adj, long long nfast, int nsd)
{
   const long long d = blockIdx.x * (long long)blockDim.x + threadIdx.x;
   const long long ndays = nfast / nsd;
   if (d > ndays) return;
   if (d == ndays) {
      for (long long j = ndays * nsd; j < nfast; j++) kdown[j] = adj[j];
      return;
   }
   double s1 = 0, s2 = 0;
   for (int j = 0; j < nsd; j++) { s1 = s1 + kdown[d * nsd + j]; s2 = s2 + adj[d * nsd + j]; }
   for (int j = 0; j < nsd; j++) kdown[d * nsd + j] = adj[d * nsd + j] * s1 / s2;
}

__global__ void suews_fill_kernel(double *p, size_t n, double v)
{
   for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void suews_dfma_kernel(double *out, int iters, double m)
{
   double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
   const double b = 1e-9;
   for (int i = 0; i < iters; i++) {
      a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
      a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// Bit-for-bit check of dev_math.cuh against the CUDA math library on pseudo-random arguments.
__global__ void suews_math_selftest_kernel(long long n, unsigned long long seed, unsigned long long *bad)
{
   unsigned long long st = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(blockIdx.x * (long long)blockDim.x + threadIdx.x + 1);
   auto next = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
   unsigned long long be = 0, bl = 0;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const unsigned long long u = next(), v = next();
      const double f = (double)(u >> 11) * 0x1p-53;
      double xe, xl;
      switch ((int)(v & 7)) {
      case 0: xe = -760.0 + 1480.0 * f; break;                    // whole range, both tails
      case 1: xe = -40.0 + 80.0 * f; break;                       // where the physics lives
      case 2: xe = (f - 0.5) * 1e-3; break;
      case 3: xe = __longlong_as_double((long long)u); break;     // any bit pattern (NaN, inf, subnormal)
      default: xe = -20.0 * f; break;
      }
      switch ((int)((v >> 3) & 7)) {
      case 0: xl = __longlong_as_double((long long)(u & 0x7fffffffffffffffull)); break;   // any non-negative pattern
      case 1: xl = 0.5 + f; break;                                // around 1
      case 2: xl = __longlong_as_double((long long)u); break;     // negatives, NaN
      case 3: xl = f * 0x1p-1040; break;                          // subnormal
      default: xl = 1e-6 + 1e6 * f * f * f; break;
      }
      const double a = su_exp_fast(xe), b = exp(xe), c = su_log_fast(xl), d = log(xl);
      if (__double_as_longlong(a) != __double_as_longlong(b) && !(a != a && b != b)) be++;
      if (__double_as_longlong(c) != __double_as_longlong(d) && !(c != c && d != d)) bl++;
   }
   if (be) atomicAdd(&bad[0], be);
   if (bl) atomicAdd(&bad[1], bl);
}

// The screened wet-bulb search of dev_atmos.cuh (Lat_vap) against the plain loop, on pseudo-random arguments over the whole
// range the reference accepts and beyond it: latent heat (bit pattern), exit switch state and warning bits must agree.
__global__ void suews_latvap_selftest_kernel(long long n, unsigned long long seed, unsigned long long *bad)
{
   unsigned long long st = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(blockIdx.x * (long long)blockDim.x + threadIdx.x + 1);
   auto next = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
   auto unif = [&]() { return (double)(next() >> 11) * 0x1p-53; };
   unsigned long long nb = 0;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const int kind = (int)(next() & 7);
      double T = (kind == 0) ? -60.0 + 130.0 * unif() : ((kind == 1) ? (unif() - 0.5) * 0.01 : -35.0 + 80.0 * unif());
      double P = (kind == 2) ? 500.0 + 700.0 * unif() : 950.0 + 100.0 * unif();
      double rh = (kind == 3) ? 100.0 * unif() * unif() * unif() : 100.0 * unif();
      if (kind == 4) rh = 100.0 - 1e-3 * unif();
      SuDiag d0, d1;
      d0.fatal = 0; d0.warn[0] = 0; d0.warn[1] = 0; d1 = d0;
      double Tx = su_fix_T(T);
      const double es = sat_vap_press_v(Tx, P);
      const double ea = rh / 100 * es;
      const double cp = 1000.0 + 60.0 * unif();
      int s0 = (int)(next() & 3), s1 = s0;
      const double a = Lat_vap_t<true>(d0, T, ea, P, cp, s0);
      const double b = Lat_vap_t<false>(d1, T, ea, P, cp, s1);
      if (__double_as_longlong(a) != __double_as_longlong(b) || s0 != s1 || d0.warn[0] != d1.warn[0] || d0.warn[1] != d1.warn[1]) nb++;
   }
   if (nb) atomicAdd(bad, nb);
}

__global__ void suews_sunpos_kernel(double year, double idectime, double UTC, double lat, double lng, double alt, double *out)
{
   SuewsEphem e;
   spa_time_part(year, idectime, UTC, e);
   double az, zen;
   spa_location_part(e, lat, lng, alt, az, zen);
   out[0] = az;
   out[1] = zen;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct FieldDesc {
   const char *name;
   int kind;     // 0 class int, 1 class double, 2 cell double, 3 state double, 4 ESTM state double
   size_t off;   // element offset inside its group
   int n;
};

static std::vector<FieldDesc> build_fields()
{
   std::vector<FieldDesc> v;
   size_t o;
#define ADD(kind_) [&](const char *nm, int n) { v.push_back({nm, kind_, o, n}); o += (size_t)n; }
   {
      o = 0;
      auto add = ADD(0);
#define X(name, n) add(#name, n);
      SUEWS_CLASS_INT_FIELDS(X)
#undef X
   }
   {
      o = 0;
      auto add = ADD(1);
#define X(name, n) add(#name, n);
      SUEWS_CLASS_DBL_FIELDS(X)
#undef X
   }
   {
      o = 0;
      auto add = ADD(2);
#define X(name, n) add(#name, n);
      SUEWS_CELL_DBL_FIELDS(X)
#undef X
   }
   {
      o = 0;
      auto add = ADD(3);
#define X(name, n) add(#name, n);
      SUEWS_STATE_DBL_FIELDS(X)
#undef X
   }
   {
      o = 0;
      auto add = ADD(4);
#define X(name, n) add(#name, n);
      SUEWS_ESTM_STATE_FIELDS(X)
#undef X
   }
#undef ADD
   return v;
}
static const std::vector<FieldDesc> &fields()
{
   static std::vector<FieldDesc> f = build_fields();
   return f;
}
static const FieldDesc *find_field(const char *name)
{
   for (const auto &f : fields())
      if (strcasecmp(f.name, name) == 0) return &f;
   return nullptr;
}

static thread_local int tl_err_code = 0;
static thread_local char tl_err_msg[256] = "";
static int set_err(int code, const char *fmt, ...)
{
   tl_err_code = code;
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(tl_err_msg, sizeof(tl_err_msg), fmt, ap);
   va_end(ap);
   return code;
}
#define CUDA_OK(call)                                                                                       \
   do {                                                                                                     \
      cudaError_t e_ = (call);                                                                              \
      if (e_ != cudaSuccess) return set_err(SUEWS_E_NO_DEVICE, "CUDA: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
   } while (0)

static const char *fatal_text(int code)
{
   switch (code) {
   case 14: return "In SUEWS_cal_RoughnessParameters, z0 / zd / (z-zd) < 0 m.";
   case 17: return "In stability subroutine, (z-zd) < z0.";
   case 21: return "In SUEWS_OHM.f95: bad value for qn1 found during qs calculation.";
   case 32: return "Windspeed Ht too low relative to zdm [Stability calc]";
   case 62: return "SUEWS_store / SUEWS_Calculations: soil store out of range";
   case 71: return "Value of gsModel not recognised.";
   case 73: return "RunControl.nml:EmissionsMethod unusable";
   case 75: return "RunControl.nml: BaseTMethod / TrafficUnits not usable";
   case 1001: return "NetRadiationMethod: Value not usable";
   case 1007: return "BLUEWS (CBLuse >= 1) needs the initial-data table (suews_b200_set_cbl_initial) with a row for the day, InitialData_use = 1 or 2 and sondeflag = 0";
   case 1005: return "StorageHeatMethod 4/14 (ESTM) needs the ESTM state and a surface-temperature forcing block (suews_b200_set_estm_forcing)";
   case 1003: return "StorageHeatMethod: 0, 1, 3, 4 and 14 are available; AnOHM (3) needs the forcing block's on-the-hour rows (batched API: classes first, then the forcing; reference-shaped single steps: suews_anohm_set_met_grid)";
   case 1004: return "EmissionsMethod outside 0-6 / 11-46";
   default: return "Fatal error in SUEWS!";
   }
}

struct suews_b200_handle_s {
   int device = 0, ncell = 0, nclass = 0;
   long long stride = 0;
   std::vector<SuewsClassParams> classes;
   std::vector<double> cellp, state;   // SoA [field][stride]
   std::vector<double> estm;           // ESTM state SoA [SUEWS_N_ESTM_DBL][stride]; empty until ESTM is used
   std::vector<double> ts5;            // ESTM surface-temperature forcing [nsteps][SUEWS_N_TS5]
   std::vector<double> ts5_block;      // ... as disaggregated on the device, [nsteps x 13] column-major
   double *d_estm = nullptr, *d_ts5 = nullptr;
   // AnOHM (StorageHeatMethod 3): the on-the-hour rows of the staged forcing block grouped by day of year (dev_anohm.cuh)
   SuewsHourRec *d_an_hours = nullptr;
   int *d_an_day = nullptr;   // [3][367]: first row, row count, rows with Sd > 0 of each day of year
   double *d_met_orig = nullptr, *d_met_fast = nullptr;   // forcing at file resolution / disaggregated on the device
   double *d_agg = nullptr;                               // aggregated SUEWS block [rows][87][stride] (thread-slot order)
   double *d_agg_daily = nullptr;                         // DailyState rows of the same launch [rows][50][stride]
   size_t agg_daily_cap = 0;
   int agg_daily_rows = 0;
   std::vector<unsigned char> day_end_of_step;            // 1 where a forcing row is the last step of its day
   size_t agg_cap = 0;
   int agg_rows = 0, out_freq_s = 3600;
   std::vector<int> slot_of_cell, cell_of_slot;
   std::vector<double> xfer;   // staging for the cell-order <-> slot-order moves
   int ts5_nsteps = 0;
   std::vector<int> cls_of_cell, tz_of_cell;   // (cell order; the device copies d_cls / d_tz are in slot order)
   std::vector<double> tz_list;
   int dt_since_start = 0, tstep_prev = 0, tstep = 0;
   int nsteps = 0;
   bool uploaded = false;
   // an advance ran since the last upload: the device holds the only current copy of state, sums and diagnostics
   bool dev_dirty = false;
   bool keep_diag = false;   // keep_* hold sums / diagnostics pulled back before a parameter-only re-upload
   std::vector<double> keep_sums;
   std::vector<int> keep_fatal;
   std::vector<unsigned long long> keep_warn;
   std::vector<long long> keep_iters;
   // Buckets: thread slots are sorted by (bucket, class, canopy height); a bucket is a set of classes that agree on the
   // switches the kernel variants specialise on (bit 2 ESTM, bit 1 snowUse = 1, bit 0 radiation scheme that feeds the
   // surface temperature back, bit 3 neither snowUse 0 nor 1).  bucket_start[b] .. bucket_start[b + 1] are bucket b's slots.
   int bucket_start[17] = {0};
   cudaStream_t bstream[16] = {nullptr};   // side streams: the buckets of one advance run concurrently
   cudaEvent_t bev[16] = {nullptr};
   cudaEvent_t ev_fork = nullptr;
   bool roll_tstep_prev = true;   // batched API: the library's own clock; the reference-shaped calls keep tstep_prev constant
   unsigned carveout_done = 0;    // bit per kernel instance whose L1 carve-out was set on this handle's device
   // device
   SuewsClassParams *d_classes = nullptr;
   int *d_cls = nullptr, *d_tz = nullptr, *d_fatal = nullptr, *d_order = nullptr;
   double *d_cellp = nullptr, *d_state = nullptr, *d_sums = nullptr, *d_tzlist = nullptr;
   unsigned long long *d_warn = nullptr;
   long long *d_iters = nullptr;
   SuewsStepRec *d_steps = nullptr;
   SuewsEphem *d_ephem = nullptr;
   double *d_blk[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // SUEWS, snow, RSL, DailyState, SOLWEIG, ESTM, BL
   double *d_cbl_ini = nullptr;   // BLUEWS: IniCBLdata [cbl_nline][8] row-major
   int cbl_nline = 0;
   int blk_nsteps = 0;
   size_t blk_cap[7] = {0, 0, 0, 0, 0, 0, 0};
   cudaStream_t stream = nullptr;
   cudaEvent_t ev0 = nullptr, ev1 = nullptr;
   int first_fatal_code = 0, first_fatal_cell = -1;
   double *d_pack = nullptr;   // compact cell-ordered block of selected sums columns (get_sums_cols)
   size_t pack_cap = 0;
   void *pinned = nullptr;   // pinned host staging (forcing records in, selected sums out)
   size_t pinned_bytes = 0;
   void *pinned_in = nullptr;
   size_t pinned_in_bytes = 0;
   size_t steps_cap = 0;
};

template <typename T> static cudaError_t dev_alloc(T **p, size_t n)
{
   if (*p) { cudaFree(*p); *p = nullptr; }
   return cudaMalloc((void **)p, sizeof(T) * (n ? n : 1));
}

// Device copies of the per-cell SoA arrays (cell parameters, state, ESTM state) are stored in THREAD-SLOT order -- the
// cost-sorted order the kernel walks -- so that a warp's load of one field is one contiguous 256 B line.  The host
// keeps cell order; these two move a block of fields across, through h->slot_of_cell / the order map.
static void to_slot_order(const double *cell_major, double *slot_major, const std::vector<int> &order, size_t nfield, size_t stride)
{
   const size_t n = order.size();
   for (size_t j = 0; j < nfield; j++) {
      const double *src = cell_major + j * stride;
      double *dst = slot_major + j * stride;
      for (size_t sl = 0; sl < n; sl++) dst[sl] = src[(size_t)order[sl]];
      for (size_t sl = n; sl < stride; sl++) dst[sl] = 0.0;
   }
}
static void to_cell_order(const double *slot_major, double *cell_major, const std::vector<int> &order, size_t nfield, size_t stride)
{
   const size_t n = order.size();
   for (size_t j = 0; j < nfield; j++) {
      const double *src = slot_major + j * stride;
      double *dst = cell_major + j * stride;
      for (size_t sl = 0; sl < n; sl++) dst[(size_t)order[sl]] = src[sl];
   }
}

// The storage-scheme record of a fresh batch: zeros, AnOHM's coeff_grid_day = -999. (suews_phys_anohm.f95:216)
static void init_storage_record(std::vector<double> &v, size_t stride)
{
   v.assign((size_t)SUEWS_N_ESTM_DBL * stride, 0.0);
   const size_t o = offsetof(SuewsEstmState, anohm_coeff_grid_day) / sizeof(double);
   std::fill(v.begin() + (std::ptrdiff_t)(o * stride), v.begin() + (std::ptrdiff_t)((o + 21) * stride), -999.0);
}

// A setter called after an advance must not roll the run back: the host mirror of the state (and of the ESTM
// state) is refreshed from the device first, and the running sums and per-cell diagnostics are kept for the
// re-upload that follows (they restart from zero only on the first upload or through suews_b200_reset_sums).
static int pull_live(suews_b200_handle_s *h)
{
   if (!h->uploaded || !h->dev_dirty) return SUEWS_OK;
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaStreamSynchronize(h->stream));
   const size_t st = (size_t)h->stride;
   h->xfer.resize(std::max(h->xfer.size(), h->state.size()));
   CUDA_OK(cudaMemcpy(h->xfer.data(), h->d_state, sizeof(double) * h->state.size(), cudaMemcpyDeviceToHost));
   to_cell_order(h->xfer.data(), h->state.data(), h->cell_of_slot, SUEWS_N_STATE_DBL, st);
   if (h->d_estm && !h->estm.empty()) {
      h->xfer.resize(std::max(h->xfer.size(), h->estm.size()));
      CUDA_OK(cudaMemcpy(h->xfer.data(), h->d_estm, sizeof(double) * h->estm.size(), cudaMemcpyDeviceToHost));
      to_cell_order(h->xfer.data(), h->estm.data(), h->cell_of_slot, SUEWS_N_ESTM_DBL, st);
   }
   // sums and diagnostics: device slot order -> cell order (the next upload may sort the slots differently)
   h->keep_sums.resize((size_t)SUEWS_NCOL_SUEWS * st);
   h->keep_fatal.assign(st, 0);
   h->keep_warn.assign(2 * st, 0ull);
   h->keep_iters.assign(st, 0ll);
   {
      h->xfer.resize(std::max(h->xfer.size(), h->keep_sums.size()));
      CUDA_OK(cudaMemcpy(h->xfer.data(), h->d_sums, sizeof(double) * h->keep_sums.size(), cudaMemcpyDeviceToHost));
      to_cell_order(h->xfer.data(), h->keep_sums.data(), h->cell_of_slot, SUEWS_NCOL_SUEWS, st);
      std::vector<int> f(st);
      std::vector<unsigned long long> w(2 * st);
      std::vector<long long> it(st);
      CUDA_OK(cudaMemcpy(f.data(), h->d_fatal, sizeof(int) * st, cudaMemcpyDeviceToHost));
      CUDA_OK(cudaMemcpy(w.data(), h->d_warn, sizeof(unsigned long long) * 2 * st, cudaMemcpyDeviceToHost));
      CUDA_OK(cudaMemcpy(it.data(), h->d_iters, sizeof(long long) * st, cudaMemcpyDeviceToHost));
      for (int sl = 0; sl < h->ncell; sl++) {
         const size_t c = (size_t)h->cell_of_slot[(size_t)sl];
         h->keep_fatal[c] = f[(size_t)sl];
         h->keep_warn[c] = w[(size_t)sl];
         h->keep_warn[st + c] = w[st + (size_t)sl];
         h->keep_iters[c] = it[(size_t)sl];
      }
   }
   h->keep_diag = true;
   h->dev_dirty = false;
   return SUEWS_OK;
}

extern "C" {

int suews_last_error(int *code, char *buf, int len)
{
   if (code) *code = tl_err_code;
   if (buf && len > 0) {
      strncpy(buf, tl_err_msg, (size_t)len - 1);
      buf[len - 1] = 0;
   }
   return tl_err_code;
}

// (for the other translation units of the library)
extern "C" int suews_set_last_error(int code, const char *msg) { return set_err(code, "%s", msg); }

int suews_b200_device_count(void)
{
   int n = 0;
   if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
   return n;
}

int suews_b200_field_info(const char *name, int *kind)
{
   const FieldDesc *f = find_field(name);
   if (!f) return -1;
   if (kind) *kind = f->kind;
   return f->n;
}

int suews_b200_create(int ncell, int nclass, int device, suews_b200_handle *out)
{
   if (!out || ncell <= 0 || nclass <= 0) return set_err(SUEWS_E_BAD_ARG, "create: bad ncell/nclass");
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0)
      return set_err(SUEWS_E_NO_DEVICE, "libsuews_b200 needs a CUDA device (no CPU fallback): %s",
                     e != cudaSuccess ? cudaGetErrorString(e) : "no device found");
   if (device < 0 || device >= ndev) return set_err(SUEWS_E_BAD_ARG, "create: device %d of %d", device, ndev);
   CUDA_OK(cudaSetDevice(device));
   auto *h = new suews_b200_handle_s();
   h->device = device;
   h->ncell = ncell;
   h->nclass = nclass;
   h->stride = ((long long)ncell + 31) / 32 * 32;
   h->classes.assign((size_t)nclass, SuewsClassParams());
   memset(h->classes.data(), 0, sizeof(SuewsClassParams) * (size_t)nclass);
   h->cellp.assign((size_t)SUEWS_N_CELL_DBL * h->stride, 0.0);
   h->state.assign((size_t)SUEWS_N_STATE_DBL * h->stride, 0.0);
   h->cls_of_cell.assign((size_t)ncell, 0);
   h->tz_of_cell.assign((size_t)ncell, 0);
   if (cudaStreamCreate(&h->stream) != cudaSuccess || cudaEventCreate(&h->ev0) != cudaSuccess ||
       cudaEventCreate(&h->ev1) != cudaSuccess) {
      delete h;
      return set_err(SUEWS_E_NO_DEVICE, "CUDA stream/event creation failed");
   }
   *out = h;
   return SUEWS_OK;
}

static void free_dev(suews_b200_handle h)
{
   cudaSetDevice(h->device);
   cudaFree(h->d_classes); cudaFree(h->d_cls); cudaFree(h->d_tz); cudaFree(h->d_fatal); cudaFree(h->d_cellp);
   cudaFree(h->d_state); cudaFree(h->d_sums); cudaFree(h->d_tzlist); cudaFree(h->d_warn); cudaFree(h->d_iters);
   cudaFree(h->d_steps); cudaFree(h->d_ephem); cudaFree(h->d_order);
   for (int i = 0; i < 7; i++) cudaFree(h->d_blk[i]);
   cudaFree(h->d_cbl_ini);
   cudaFree(h->d_an_hours); cudaFree(h->d_an_day); cudaFree(h->d_pack); cudaFree(h->d_estm); cudaFree(h->d_ts5); cudaFree(h->d_met_orig); cudaFree(h->d_met_fast); cudaFree(h->d_agg); cudaFree(h->d_agg_daily);
}

int suews_b200_destroy(suews_b200_handle h)
{
   if (!h) return SUEWS_OK;
   free_dev(h);
   if (h->pinned) cudaFreeHost(h->pinned);
   if (h->pinned_in) cudaFreeHost(h->pinned_in);
   for (int k = 0; k < 16; k++) {
      if (h->bstream[k]) cudaStreamDestroy(h->bstream[k]);
      if (h->bev[k]) cudaEventDestroy(h->bev[k]);
   }
   if (h->ev_fork) cudaEventDestroy(h->ev_fork);
   if (h->stream) cudaStreamDestroy(h->stream);
   if (h->ev0) cudaEventDestroy(h->ev0);
   if (h->ev1) cudaEventDestroy(h->ev1);
   delete h;
   return SUEWS_OK;
}

int suews_b200_set_class_param(suews_b200_handle h, int cls, const char *name, const double *vals, int n)
{
   if (!h || cls < 0 || cls >= h->nclass) return set_err(SUEWS_E_BAD_ARG, "set_class_param: bad class %d", cls);
   const FieldDesc *f = find_field(name);
   if (!f || f->kind > 1) return set_err(SUEWS_E_BAD_NAME, "set_class_param: '%s' is not a class field", name);
   if (n != f->n) return set_err(SUEWS_E_BAD_SIZE, "set_class_param: '%s' has %d values, got %d", name, f->n, n);
   SuewsClassParams &P = h->classes[(size_t)cls];
   if (f->kind == 0) {
      int *base = reinterpret_cast<int *>(&P);
      for (int i = 0; i < n; i++) base[f->off + i] = (int)vals[i];
   } else {
      double *base = reinterpret_cast<double *>(&P.AlbMax_DecTr);
      for (int i = 0; i < n; i++) base[f->off + i] = vals[i];
   }
   { int rc_ = pull_live(h); if (rc_) return rc_; }
   h->uploaded = false;
   return SUEWS_OK;
}

static int set_soa(std::vector<double> &soa, long long stride, int ncell, const FieldDesc *f, const double *vals)
{
   for (int k = 0; k < f->n; k++) {
      double *dst = &soa[(f->off + (size_t)k) * (size_t)stride];
      for (int c = 0; c < ncell; c++) dst[c] = vals[(size_t)c * f->n + k];
   }
   return SUEWS_OK;
}

int suews_b200_set_cell_param(suews_b200_handle h, const char *name, const double *vals, int n)
{
   if (!h) return set_err(SUEWS_E_BAD_ARG, "null handle");
   const FieldDesc *f = find_field(name);
   if (!f || f->kind != 2) return set_err(SUEWS_E_BAD_NAME, "set_cell_param: '%s' is not a per-cell field", name);
   if (n != f->n) return set_err(SUEWS_E_BAD_SIZE, "set_cell_param: '%s' has %d values per cell, got %d", name, f->n, n);
   { int rc_ = pull_live(h); if (rc_) return rc_; }
   h->uploaded = false;
   return set_soa(h->cellp, h->stride, h->ncell, f, vals);
}

int suews_b200_set_cell_class(suews_b200_handle h, const int *class_of_cell)
{
   if (!h) return set_err(SUEWS_E_BAD_ARG, "null handle");
   for (int c = 0; c < h->ncell; c++) {
      if (class_of_cell[c] < 0 || class_of_cell[c] >= h->nclass)
         return set_err(SUEWS_E_BAD_ARG, "set_cell_class: cell %d has class %d of %d", c, class_of_cell[c], h->nclass);
      h->cls_of_cell[(size_t)c] = class_of_cell[c];
   }
   { int rc_ = pull_live(h); if (rc_) return rc_; }
   h->uploaded = false;
   return SUEWS_OK;
}

int suews_b200_set_state(suews_b200_handle h, const char *name, const double *vals, int n)
{
   if (!h) return set_err(SUEWS_E_BAD_ARG, "null handle");
   const FieldDesc *f = find_field(name);
   if (!f || (f->kind != 3 && f->kind != 4)) return set_err(SUEWS_E_BAD_NAME, "set_state: '%s' is not a state field", name);
   if (n != f->n) return set_err(SUEWS_E_BAD_SIZE, "set_state: '%s' has %d values per cell, got %d", name, f->n, n);
   { int rc_ = pull_live(h); if (rc_) return rc_; }
   h->uploaded = false;
   if (f->kind == 4) {
      if (h->estm.empty()) init_storage_record(h->estm, (size_t)h->stride);
      return set_soa(h->estm, h->stride, h->ncell, f, vals);
   }
   return set_soa(h->state, h->stride, h->ncell, f, vals);
}

int suews_b200_get_state(suews_b200_handle h, const char *name, double *vals, int n)
{
   if (!h) return set_err(SUEWS_E_BAD_ARG, "null handle");
   const FieldDesc *f = find_field(name);
   if (!f || (f->kind != 3 && f->kind != 4)) return set_err(SUEWS_E_BAD_NAME, "get_state: '%s' is not a state field", name);
   if (n != f->n) return set_err(SUEWS_E_BAD_SIZE, "get_state: '%s' has %d values per cell, got %d", name, f->n, n);
   if (f->kind == 4) {
      if (h->estm.empty()) init_storage_record(h->estm, (size_t)h->stride);
      if (h->uploaded && h->d_estm) {
         CUDA_OK(cudaSetDevice(h->device));
         CUDA_OK(cudaStreamSynchronize(h->stream));
         h->xfer.resize(std::max(h->xfer.size(), (size_t)f->n * (size_t)h->stride));
         CUDA_OK(cudaMemcpy(h->xfer.data(), h->d_estm + f->off * (size_t)h->stride,
                            sizeof(double) * (size_t)f->n * (size_t)h->stride, cudaMemcpyDeviceToHost));
         to_cell_order(h->xfer.data(), &h->estm[f->off * (size_t)h->stride], h->cell_of_slot, (size_t)f->n, (size_t)h->stride);
      }
      for (int k = 0; k < f->n; k++) {
         const double *src = &h->estm[(f->off + (size_t)k) * (size_t)h->stride];
         for (int c = 0; c < h->ncell; c++) vals[(size_t)c * f->n + k] = src[c];
      }
      return SUEWS_OK;
   }
   if (h->uploaded) {
      CUDA_OK(cudaSetDevice(h->device));
      CUDA_OK(cudaStreamSynchronize(h->stream));
      h->xfer.resize(std::max(h->xfer.size(), (size_t)f->n * (size_t)h->stride));
      CUDA_OK(cudaMemcpy(h->xfer.data(), h->d_state + f->off * (size_t)h->stride,
                         sizeof(double) * (size_t)f->n * (size_t)h->stride, cudaMemcpyDeviceToHost));
      to_cell_order(h->xfer.data(), &h->state[f->off * (size_t)h->stride], h->cell_of_slot, (size_t)f->n, (size_t)h->stride);
   }
   for (int k = 0; k < f->n; k++) {
      const double *src = &h->state[(f->off + (size_t)k) * (size_t)h->stride];
      for (int c = 0; c < h->ncell; c++) vals[(size_t)c * f->n + k] = src[c];
   }
   return SUEWS_OK;
}

// Forget the run on the device: the next upload starts the sums and diagnostics from zero and does not read the device
// state back (the reference-shaped calls reuse one resident handle and pass every argument anew).
int suews_b200_restart(suews_b200_handle h)
{
   if (!h) return set_err(SUEWS_E_BAD_ARG, "null handle");
   h->dev_dirty = false;
   h->keep_diag = false;
   h->uploaded = false;
   return SUEWS_OK;
}
// Re-sort the thread slots on the current state (the dynamic keys of upload(): snow activity, canopy height through the
// deciduous porosity).  The device state comes back to the host mirror, the slots are sorted again and everything is
// uploaded in the new order; running sums, diagnostics, the clock and the staged forcing carry on.  Worth calling every few
// weeks of simulated time in snow-on runs (melt and first snow change which cells are active).
int suews_b200_resort(suews_b200_handle h)
{
   if (!h || !h->uploaded) return set_err(SUEWS_E_BAD_ARG, "resort: nothing uploaded");
   h->dev_dirty = true;   // the device holds the current copy (also right after an upload: harmless round trip)
   int rc = pull_live(h);
   if (rc != SUEWS_OK) return rc;
   return suews_b200_upload(h);
}
// tstep_prev of SUEWS_cal_Main: rolls = 1 (default) lets the library's own clock set it to tstep after the first step of
// a run; rolls = 0 keeps the value of set_clock for every step, as SUEWS_cal_multitsteps does with its INTENT(IN) argument
int suews_b200_set_tstep_prev_rolls(suews_b200_handle h, int rolls)
{
   if (!h) return set_err(SUEWS_E_BAD_ARG, "null handle");
   h->roll_tstep_prev = rolls != 0;
   return SUEWS_OK;
}

int suews_b200_set_clock(suews_b200_handle h, int dt_since_start, int tstep_prev)
{
   if (!h) return set_err(SUEWS_E_BAD_ARG, "null handle");
   h->dt_since_start = dt_since_start;
   h->tstep_prev = tstep_prev;
   return SUEWS_OK;
}
int suews_b200_get_clock(suews_b200_handle h, int *dt_since_start, int *tstep_prev)
{
   if (!h) return set_err(SUEWS_E_BAD_ARG, "null handle");
   if (dt_since_start) *dt_since_start = h->dt_since_start;
   if (tstep_prev) *tstep_prev = h->tstep_prev;
   return SUEWS_OK;
}

int suews_b200_upload(suews_b200_handle h)
{
   if (!h) return set_err(SUEWS_E_BAD_ARG, "null handle");
   CUDA_OK(cudaSetDevice(h->device));
   // derived class constants: sequential 24-term sums of the diurnal profiles (ctrl_input.f95:784,816)
   h->tstep = h->classes[0].tstep;
   for (auto &P : h->classes) {
      const double *prof[6] = {P.AHProf_24hr.v, P.HumActivity_24hr.v, P.PopProf_24hr.v,
                               P.TraffProf_24hr.v, P.WUProfA_24hr.v, P.WUProfM_24hr.v};
      for (int k = 0; k < 6; k++)
         for (int u = 0; u < 2; u++) {
            double s = 0;
            for (int i = 0; i < 24; i++) s += prof[k][24 * u + i];
            P.prof_sum[k][u] = s;
         }
      {   // the same REAL(8) expression the reference evaluates on every call, with the host libm
         const double TH = P.TH, G5 = P.G5, TL = P.TL;
         const double TC = (TH - G5) / (G5 - TL);
         P.gs_pow_TH_G5 = pow(TH - G5, TC);
      }
      if (P.tstep != h->tstep || P.tstep <= 0)
         return set_err(SUEWS_E_BAD_ARG, "upload: all classes must share one positive tstep (got %d and %d)", h->tstep, (int)P.tstep);
   }
   // distinct time zones -> ephemeris index per cell
   const std::vector<double> tz_before = h->tz_list;
   h->tz_list.clear();
   const FieldDesc *ftz = find_field("timezone");
   for (int c = 0; c < h->ncell; c++) {
      double tz = h->cellp[ftz->off * (size_t)h->stride + c];
      int idx = -1;
      for (size_t k = 0; k < h->tz_list.size(); k++)
         if (h->tz_list[k] == tz) { idx = (int)k; break; }
      if (idx < 0) { idx = (int)h->tz_list.size(); h->tz_list.push_back(tz); }
      h->tz_of_cell[(size_t)c] = idx;
   }
   // Divergence control.  Thread slots are sorted by (bucket, class, canopy height): a bucket is the set of classes that
   // agree on snowUse, ESTM and on whether the radiation scheme feeds the surface temperature back -- each bucket is
   // launched on its own with the kernel instance compiled for those switches (suews_kernel.cuh), so no block ever mixes
   // scheme sets; inside a bucket warps are filled with cells of one class and similar canopy height.  The
   // roughness-sublayer search (cal_psim_hat, rslprof.f95:366) walks ln(100/zH)/ln(1.01) levels and is skipped for
   // zH <= 6 m, so zH is the per-cell cost driver.
   std::vector<int> order((size_t)h->ncell);
   std::vector<int> bucket_of_class((size_t)h->nclass);
   for (int k = 0; k < h->nclass; k++) {
      const SuewsClassParams &P = h->classes[(size_t)k];
      const int nrm = P.NetRadiationMethod;
      int b = 0;
      if (P.StorageHeatMethod == 4 || P.StorageHeatMethod == 14 || P.StorageHeatMethod == 3 || P.CBLuse >= 1) b |= 4;   // schemes with a record of their own
      if (P.snowUse == 1) b |= 2;
      else if (P.snowUse != 0) b |= 8;
      if (nrm >= 10 && nrm < 100) b |= 1;   // RadMethod: nrm_use >= 10 iterates on the surface temperature (narp.f95:337)
      bucket_of_class[(size_t)k] = b;
   }
   {
      const FieldDesc *fs = find_field("sfr"), *fb = find_field("bldgH"), *fe = find_field("EveTreeH"),
                      *fd = find_field("DecTreeH"), *fp = find_field("porosity_id"), *fsp = find_field("SnowPack");
      const size_t sd = (size_t)h->stride;
      std::vector<double> key((size_t)h->ncell);
      for (int c = 0; c < h->ncell; c++) {
         double s1 = h->cellp[(fs->off + 1) * sd + c], s2 = h->cellp[(fs->off + 2) * sd + c], s3 = h->cellp[(fs->off + 3) * sd + c];
         double a = s1 + s2 + s3;
         double zh = (a != 0) ? (h->cellp[fb->off * sd + c] * s1 + h->cellp[fe->off * sd + c] * s2 +
                                 h->cellp[fd->off * sd + c] * (1 - h->state[fp->off * sd + c]) * s3) / a : 0.0;
         const int k = h->cls_of_cell[(size_t)c];
         // snow activity (SUM(SnowPack) > 0): in a snow-on class the cells that carry a pack run the snow branch
         // of SUEWS_cal_QE and the melt-heat routine's inner arms, the bare ones do not; it moves with the seasons, so it
         // is a sort key of the current state -- suews_b200_resort() re-applies it between launches
         double active = 0;
         if (h->classes[(size_t)k].snowUse == 1) {
            double sp = 0;
            for (int is = 0; is < SUEWS_NSURF; is++) sp += h->state[(fsp->off + is) * sd + c];
            active = (sp > 0) ? 1.0 : 0.0;
         }
         // canopy heights are below 1e3 m: the activity flag sits above them, below the class number
         key[(size_t)c] = (((double)bucket_of_class[(size_t)k] * 1e6 + (double)k) * 2 + active) * 1e4 + std::min(zh, 9999.0);
         order[(size_t)c] = c;
      }
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key[(size_t)x] < key[(size_t)y]; });
      for (int b = 0; b <= 16; b++) h->bucket_start[b] = 0;
      for (int c = 0; c < h->ncell; c++) h->bucket_start[bucket_of_class[(size_t)h->cls_of_cell[(size_t)c]] + 1]++;
      for (int b = 0; b < 16; b++) h->bucket_start[b + 1] += h->bucket_start[b];
   }
   h->slot_of_cell.assign((size_t)h->ncell, 0);
   for (int sl = 0; sl < h->ncell; sl++) h->slot_of_cell[(size_t)order[(size_t)sl]] = sl;
   h->cell_of_slot = order;
   const size_t st = (size_t)h->stride;
   CUDA_OK(dev_alloc(&h->d_order, st));
   CUDA_OK(cudaMemcpy(h->d_order, order.data(), sizeof(int) * (size_t)h->ncell, cudaMemcpyHostToDevice));
   CUDA_OK(dev_alloc(&h->d_classes, (size_t)h->nclass));
   CUDA_OK(dev_alloc(&h->d_cls, st));
   CUDA_OK(dev_alloc(&h->d_tz, st));
   CUDA_OK(dev_alloc(&h->d_fatal, st));
   CUDA_OK(dev_alloc(&h->d_cellp, (size_t)SUEWS_N_CELL_DBL * st));
   CUDA_OK(dev_alloc(&h->d_state, (size_t)SUEWS_N_STATE_DBL * st));
   CUDA_OK(dev_alloc(&h->d_sums, (size_t)SUEWS_NCOL_SUEWS * st));
   CUDA_OK(dev_alloc(&h->d_warn, 2 * st));
   CUDA_OK(dev_alloc(&h->d_iters, st));
   CUDA_OK(dev_alloc(&h->d_tzlist, h->tz_list.size()));
   CUDA_OK(cudaMemcpy(h->d_classes, h->classes.data(), sizeof(SuewsClassParams) * (size_t)h->nclass, cudaMemcpyHostToDevice));
   {   // class and time-zone index per thread slot
      std::vector<int> tmp((size_t)h->ncell);
      for (int sl = 0; sl < h->ncell; sl++) tmp[(size_t)sl] = h->cls_of_cell[(size_t)order[(size_t)sl]];
      CUDA_OK(cudaMemcpy(h->d_cls, tmp.data(), sizeof(int) * (size_t)h->ncell, cudaMemcpyHostToDevice));
      for (int sl = 0; sl < h->ncell; sl++) tmp[(size_t)sl] = h->tz_of_cell[(size_t)order[(size_t)sl]];
      CUDA_OK(cudaMemcpy(h->d_tz, tmp.data(), sizeof(int) * (size_t)h->ncell, cudaMemcpyHostToDevice));
   }
   h->xfer.resize(std::max(h->cellp.size(), h->state.size()));
   to_slot_order(h->cellp.data(), h->xfer.data(), order, SUEWS_N_CELL_DBL, st);
   CUDA_OK(cudaMemcpy(h->d_cellp, h->xfer.data(), sizeof(double) * h->cellp.size(), cudaMemcpyHostToDevice));
   to_slot_order(h->state.data(), h->xfer.data(), order, SUEWS_N_STATE_DBL, st);
   CUDA_OK(cudaMemcpy(h->d_state, h->xfer.data(), sizeof(double) * h->state.size(), cudaMemcpyHostToDevice));
   CUDA_OK(cudaMemcpy(h->d_tzlist, h->tz_list.data(), sizeof(double) * h->tz_list.size(), cudaMemcpyHostToDevice));
   {   // ESTM state travels only when some class runs ESTM
      bool any_estm = false;
      for (const auto &P : h->classes) any_estm = any_estm || P.StorageHeatMethod == 4 || P.StorageHeatMethod == 14 || P.StorageHeatMethod == 3 || P.CBLuse >= 1;
      if (any_estm) {
         if (h->estm.empty()) init_storage_record(h->estm, (size_t)h->stride);
         CUDA_OK(dev_alloc(&h->d_estm, (size_t)SUEWS_N_ESTM_DBL * st));
         to_slot_order(h->estm.data(), h->xfer.data(), order, SUEWS_N_ESTM_DBL, st);
         CUDA_OK(cudaMemcpy(h->d_estm, h->xfer.data(), sizeof(double) * h->estm.size(), cudaMemcpyHostToDevice));
      } else if (h->d_estm) {
         cudaFree(h->d_estm);
         h->d_estm = nullptr;
      }
   }
   CUDA_OK(cudaMemset(h->d_sums, 0, sizeof(double) * (size_t)SUEWS_NCOL_SUEWS * st));
   CUDA_OK(cudaMemset(h->d_fatal, 0, sizeof(int) * st));
   CUDA_OK(cudaMemset(h->d_warn, 0, sizeof(unsigned long long) * 2 * st));
   CUDA_OK(cudaMemset(h->d_iters, 0, sizeof(long long) * st));
   if (h->keep_diag) {   // parameter-only re-upload in the middle of a run (pull_live): sums and diagnostics carry on
      h->xfer.resize(std::max(h->xfer.size(), h->keep_sums.size()));
      to_slot_order(h->keep_sums.data(), h->xfer.data(), order, SUEWS_NCOL_SUEWS, st);
      CUDA_OK(cudaMemcpy(h->d_sums, h->xfer.data(), sizeof(double) * h->keep_sums.size(), cudaMemcpyHostToDevice));
      std::vector<int> f(st, 0);
      std::vector<unsigned long long> w(2 * st, 0ull);
      std::vector<long long> it(st, 0ll);
      for (int sl = 0; sl < h->ncell; sl++) {
         const size_t c = (size_t)order[(size_t)sl];
         f[(size_t)sl] = h->keep_fatal[c];
         w[(size_t)sl] = h->keep_warn[c];
         w[st + (size_t)sl] = h->keep_warn[st + c];
         it[(size_t)sl] = h->keep_iters[c];
      }
      CUDA_OK(cudaMemcpy(h->d_fatal, f.data(), sizeof(int) * st, cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(h->d_warn, w.data(), sizeof(unsigned long long) * 2 * st, cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(h->d_iters, it.data(), sizeof(long long) * st, cudaMemcpyHostToDevice));
      h->keep_diag = false;
   }
   h->dev_dirty = false;
   h->uploaded = true;
   // the ephemeris table is laid out by distinct time zone: a forcing block staged before a re-upload stays usable only
   // while that list is unchanged
   if (h->tz_list != tz_before) { h->steps_cap = 0; h->nsteps = 0; }
   h->first_fatal_code = 0;
   h->first_fatal_cell = -1;
   return SUEWS_OK;
}

// AnOHM reads MetForcingData_grid, the host's whole forcing block of the grid (suews_ctrl_translate.f95:1216): its rows
// with imin == 0 of the day being fitted (AnOHM_FcLoad, suews_phys_anohm.f95:717-805).  Here: the staged block's
// on-the-hour rows grouped by day of year, block order kept inside a day, so that a cell reads its day's rows directly.
// Built only when a class runs StorageHeatMethod 3.  recs == nullptr: the step records are on the device only.
static int build_anohm_hours(suews_b200_handle h, const SuewsStepRec *recs, size_t n)
{
   bool any = false;
   for (const auto &P : h->classes) any = any || P.StorageHeatMethod == 3;
   if (!any) {
      if (h->d_an_hours) { cudaFree(h->d_an_hours); h->d_an_hours = nullptr; }
      if (h->d_an_day) { cudaFree(h->d_an_day); h->d_an_day = nullptr; }
      return SUEWS_OK;
   }
   std::vector<SuewsStepRec> back;
   if (!recs) {
      back.resize(n);
      CUDA_OK(cudaMemcpy(back.data(), h->d_steps, sizeof(SuewsStepRec) * n, cudaMemcpyDeviceToHost));
      recs = back.data();
   }
   std::vector<int> day(3 * 367, 0);
   int *first = day.data(), *count = first + 367, *nsun = count + 367;
   size_t nh = 0;
   for (size_t i = 0; i < n; i++)
      if (recs[i].imin == 0 && recs[i].id >= 1 && recs[i].id <= 366) {
         count[recs[i].id]++;
         if (recs[i].avkdn > 0) nsun[recs[i].id]++;
         nh++;
      }
   for (int d = 1; d <= 366; d++) {
      first[d] = first[d - 1] + count[d - 1];
      if (count[d] > AN_MAXM)
         return set_err(SUEWS_E_NOT_IMPL, "AnOHM: the forcing block holds %d on-the-hour rows of day of year %d (at most %d: stage one year at a time)", count[d], d, AN_MAXM);
   }
   std::vector<SuewsHourRec> hours(nh ? nh : 1);
   std::vector<int> fill(367, 0);
   for (size_t i = 0; i < n; i++)
      if (recs[i].imin == 0 && recs[i].id >= 1 && recs[i].id <= 366) {
         const SuewsStepRec &r = recs[i];
         SuewsHourRec &q = hours[(size_t)(first[r.id] + fill[r.id]++)];
         q.id = (double)r.id; q.tHr = (double)r.it; q.Sd = r.avkdn; q.Ta = r.Temp_C; q.RH = r.avRh; q.pres = r.Press_hPa;
         q.WS = r.avU1; q.qf_obs = r.qf_obs;
      }
   if (h->d_an_hours) { cudaFree(h->d_an_hours); h->d_an_hours = nullptr; }
   CUDA_OK(cudaMalloc((void **)&h->d_an_hours, sizeof(SuewsHourRec) * hours.size()));
   CUDA_OK(cudaMemcpy(h->d_an_hours, hours.data(), sizeof(SuewsHourRec) * hours.size(), cudaMemcpyHostToDevice));
   if (!h->d_an_day) CUDA_OK(cudaMalloc((void **)&h->d_an_day, sizeof(int) * day.size()));
   CUDA_OK(cudaMemcpy(h->d_an_day, day.data(), sizeof(int) * day.size(), cudaMemcpyHostToDevice));
   return SUEWS_OK;
}

int suews_b200_set_forcing(suews_b200_handle h, const double *met, int nsteps)
{
   if (!h || !met || nsteps <= 0) return set_err(SUEWS_E_BAD_ARG, "set_forcing: bad arguments");
   if (!h->uploaded) {
      int rc = suews_b200_upload(h);
      if (rc) return rc;
   }
   CUDA_OK(cudaSetDevice(h->device));
   const size_t n = (size_t)nsteps;
   if (h->pinned_in_bytes < sizeof(SuewsStepRec) * n) {
      if (h->pinned_in) cudaFreeHost(h->pinned_in);
      h->pinned_in = nullptr;
      CUDA_OK(cudaHostAlloc(&h->pinned_in, sizeof(SuewsStepRec) * n, cudaHostAllocDefault));
      h->pinned_in_bytes = sizeof(SuewsStepRec) * n;
   }
   SuewsStepRec *recs = (SuewsStepRec *)h->pinned_in;
   for (int i = 0; i < nsteps; i++) pack_step(recs[(size_t)i], met, n, (size_t)i);   // driver.f95:3177-3201
   if (h->steps_cap < n) {
      CUDA_OK(dev_alloc(&h->d_steps, n));
      CUDA_OK(dev_alloc(&h->d_ephem, n * h->tz_list.size()));
      h->steps_cap = n;
   }
   CUDA_OK(cudaMemcpyAsync(h->d_steps, recs, sizeof(SuewsStepRec) * n, cudaMemcpyHostToDevice, h->stream));
   int total = nsteps * (int)h->tz_list.size();
   suews_ephem_kernel<<<(total + 127) / 128, 128, 0, h->stream>>>(h->d_steps, nsteps, h->d_tzlist, (int)h->tz_list.size(),
                                                                 h->tstep, h->d_ephem);
   CUDA_OK(cudaGetLastError());
   CUDA_OK(cudaStreamSynchronize(h->stream));
   h->nsteps = nsteps;
   h->ts5_nsteps = 0;   // a surface-temperature block belongs to one forcing block
   if (h->d_met_fast) { cudaFree(h->d_met_fast); h->d_met_fast = nullptr; }
   return build_anohm_hours(h, recs, n);
}

// AnOHM's MetForcingData_grid given explicitly ([nrows x 24] column-major, the MetForcingBlock column layout) instead of
// the staged forcing block: for hosts that stage the forcing in pieces (the reference-shaped single-step call) while the
// reference's AnOHM sees the grid's whole block (suews_ctrl_translate.f95:1216).  Call it after set_forcing; the next
// set_forcing goes back to its own block.
int suews_b200_set_anohm_forcing(suews_b200_handle h, const double *met, int nrows)
{
   if (!h || !met || nrows <= 0 || !h->uploaded) return set_err(SUEWS_E_BAD_ARG, "set_anohm_forcing: bad arguments (upload and set the forcing first)");
   CUDA_OK(cudaSetDevice(h->device));
   std::vector<SuewsStepRec> recs((size_t)nrows);
   for (int i = 0; i < nrows; i++) pack_step(recs[(size_t)i], met, (size_t)nrows, (size_t)i);
   return build_anohm_hours(h, recs.data(), (size_t)nrows);
}

// BLUEWS: IniCBLdata (CBL_ReadInputData, suews_phys_bluews.f95:319-341), [nline x 8] row-major: day of year, zi0, gamt [K m-1],
// gamq [g kg-1 m-1], theta+ [K], q+ [g kg-1], theta [K], q [g kg-1]; shared by all cells of the batch.
int suews_b200_set_cbl_initial(suews_b200_handle h, const double *ini, int nline)
{
   if (!h || !ini || nline <= 0) return set_err(SUEWS_E_BAD_ARG, "set_cbl_initial: bad arguments");
   CUDA_OK(cudaSetDevice(h->device));
   if (h->d_cbl_ini) { cudaFree(h->d_cbl_ini); h->d_cbl_ini = nullptr; }
   CUDA_OK(cudaMalloc((void **)&h->d_cbl_ini, sizeof(double) * 8 * (size_t)nline));
   CUDA_OK(cudaMemcpy(h->d_cbl_ini, ini, sizeof(double) * 8 * (size_t)nline, cudaMemcpyHostToDevice));
   h->cbl_nline = nline;
   return SUEWS_OK;
}

// Forcing at its original resolution: disaggregated to the model timestep on the device.
int suews_b200_set_forcing_orig(suews_b200_handle h, const double *met_orig, int nrows, int resolution_in_s, int disagg_method,
                                int rain_disagg_method, int raw_units, int smd_method)
{
   if (!h || !met_orig || nrows < 2) return set_err(SUEWS_E_BAD_ARG, "set_forcing_orig: need at least two rows");
   if (!h->uploaded) {
      int rc = suews_b200_upload(h);
      if (rc) return rc;
   }
   const int tstep = h->tstep;
   if (resolution_in_s <= 0 || tstep <= 0 || resolution_in_s % tstep != 0)
      return set_err(SUEWS_E_BAD_ARG, "set_forcing_orig: ResolutionFilesIn %d is not a multiple of tstep %d", resolution_in_s, tstep);
   if (disagg_method < 1 || disagg_method > 3)
      return set_err(2, "DisaggMethod value should be 1, 2, or 3 (got %d)", disagg_method);   // ErrorHint 2, ctrl_input.f95:1438
   if (rain_disagg_method != 100)   // 101 / 102 draw unseeded random sub-intervals in the reference (RandomSamples, :2070)
      return set_err(SUEWS_E_NOT_IMPL, "set_forcing_orig: only RainDisaggMethod 100 is deterministic (got %d)", rain_disagg_method);
   CUDA_OK(cudaSetDevice(h->device));
   const int nper = resolution_in_s / tstep;
   const size_t n = (size_t)nrows, nfast = n * (size_t)nper;
   if (nfast > (size_t)0x7fffffff) return set_err(SUEWS_E_BAD_ARG, "set_forcing_orig: too many timesteps");
   SuDisaggArgs a;
   a.n = nrows; a.nper = nper; a.tstep = tstep; a.raw_units = raw_units; a.smd_method = smd_method;
   a.ncol = SUEWS_NCOL_MET; a.rain_col = 13; a.snow_col = 15; a.skip_col = 23;
   a.all_missing = 0; a.type20 = 0;
   for (int c = 4; c < SUEWS_NCOL_MET; c++) {
      bool all = true;
      for (size_t i = 0; i < n && all; i++) all = (met_orig[i + n * c] == -999.0);
      if (c == 19 && smd_method == 0) all = true;
      if (all) a.all_missing |= (1u << c);
      if (disagg_method == 2 || (disagg_method == 3 && c >= 9 && c <= 12)) a.type20 |= (1u << c);
   }
   CUDA_OK(dev_alloc(&h->d_met_orig, n * SUEWS_NCOL_MET));
   CUDA_OK(dev_alloc(&h->d_met_fast, nfast * SUEWS_NCOL_MET));
   a.slow = h->d_met_orig; a.fast = h->d_met_fast;
   if (h->steps_cap < nfast) {
      CUDA_OK(dev_alloc(&h->d_steps, nfast));
      CUDA_OK(dev_alloc(&h->d_ephem, nfast * h->tz_list.size()));
      h->steps_cap = nfast;
   }
   CUDA_OK(cudaEventRecord(h->ev0, h->stream));
   CUDA_OK(cudaMemcpyAsync(h->d_met_orig, met_orig, sizeof(double) * n * SUEWS_NCOL_MET, cudaMemcpyHostToDevice, h->stream));
   suews_disagg_kernel<<<(unsigned)((nfast + 255) / 256), 256, 0, h->stream>>>(a);
   suews_pack_steps_kernel<<<(unsigned)((nfast + 255) / 256), 256, 0, h->stream>>>(h->d_met_fast, (int)nfast, h->d_steps);
   const int total = (int)nfast * (int)h->tz_list.size();
   suews_ephem_kernel<<<(total + 127) / 128, 128, 0, h->stream>>>(h->d_steps, (int)nfast, h->d_tzlist, (int)h->tz_list.size(),
                                                                 h->tstep, h->d_ephem);
   CUDA_OK(cudaGetLastError());
   CUDA_OK(cudaEventRecord(h->ev1, h->stream));
   CUDA_OK(cudaStreamSynchronize(h->stream));
   h->nsteps = (int)nfast;
   h->ts5_nsteps = 0;
   return build_anohm_hours(h, nullptr, nfast);
}

// KdownZen = 1 on the block set_forcing_orig has just disaggregated: K-down zero below the horizon of (lat, lng, alt) in
// time zone tz and rescaled day by day (suews_ctrl_input.f95:1568-1600).  The reference does this per grid, with each grid's
// own coordinates; a batch shares one forcing block, so the site given here is the forcing's.
int suews_b200_apply_kdown_zen(suews_b200_handle h, double lat, double lng, double timezone, double alt)
{
   if (!h || !h->d_met_fast || h->nsteps <= 0) return set_err(SUEWS_E_BAD_ARG, "apply_kdown_zen: no device-disaggregated block (call set_forcing_orig first)");
   CUDA_OK(cudaSetDevice(h->device));
   const long long nfast = h->nsteps;
   const int tstep = h->tstep, nsd = 24 * (3600 / tstep);
   double *d_adj = nullptr;
   CUDA_OK(cudaMalloc((void **)&d_adj, sizeof(double) * (size_t)nfast));
   suews_kdownzen_mask_kernel<<<(unsigned)((nfast + 127) / 128), 128, 0, h->stream>>>(h->d_met_fast, nfast, tstep, lat, lng, timezone, alt, d_adj);
   const long long ndays = nfast / nsd;
   suews_kdownzen_renorm_kernel<<<(unsigned)((ndays + 1 + 63) / 64), 64, 0, h->stream>>>(h->d_met_fast + 14 * (size_t)nfast, d_adj, nfast, nsd);
   suews_pack_steps_kernel<<<(unsigned)((nfast + 255) / 256), 256, 0, h->stream>>>(h->d_met_fast, (int)nfast, h->d_steps);
   cudaError_t e = cudaGetLastError();
   if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
   cudaFree(d_adj);
   CUDA_OK(e);
   return build_anohm_hours(h, nullptr, (size_t)nfast);
}

// The disaggregated MetForcingBlock [nsteps x 24], column-major, as the device holds it
int suews_b200_get_forcing_block(suews_b200_handle h, double *met_fast, int nsteps)
{
   if (!h || !met_fast || !h->d_met_fast || nsteps != h->nsteps)
      return set_err(SUEWS_E_BAD_ARG, "get_forcing_block: no device-disaggregated block of %d rows", nsteps);
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaMemcpy(met_fast, h->d_met_fast, sizeof(double) * (size_t)nsteps * SUEWS_NCOL_MET, cudaMemcpyDeviceToHost));
   return SUEWS_OK;
}

// ESTM surface temperatures at file resolution -> model timestep on the device (DisaggregateESTM, ctrl_input.f95:1650-1849:
// the date/time and Disagg_Lin of the met block, DisaggMethodESTM 1 = period averages, 2 = instantaneous values)
int suews_b200_set_estm_forcing_orig(suews_b200_handle h, const double *ts_orig, int nrows, int resolution_in_s, int disagg_method)
{
   if (!h || !ts_orig || nrows < 2 || !h->uploaded) return set_err(SUEWS_E_BAD_ARG, "set_estm_forcing_orig: bad arguments");
   const int tstep = h->tstep;
   if (resolution_in_s <= 0 || resolution_in_s % tstep != 0)
      return set_err(SUEWS_E_BAD_ARG, "set_estm_forcing_orig: ResolutionFilesInESTM %d is not a multiple of tstep %d", resolution_in_s, tstep);
   if (disagg_method != 1 && disagg_method != 2)
      return set_err(2, "Problem in SUEWS_ESTMDisagg: DisaggMethodESTM value should be 1 or 2 (got %d)", disagg_method);
   const int nper = resolution_in_s / tstep;
   const size_t n = (size_t)nrows, nfast = n * (size_t)nper;
   if ((int)nfast != h->nsteps)
      return set_err(SUEWS_E_BAD_ARG, "set_estm_forcing_orig: %zu timesteps, but the forcing block has %d (set the forcing first)", nfast, h->nsteps);
   CUDA_OK(cudaSetDevice(h->device));
   SuDisaggArgs a;
   a.n = nrows; a.nper = nper; a.tstep = tstep; a.raw_units = 0; a.smd_method = 1;
   a.ncol = 13; a.rain_col = -1; a.snow_col = -1; a.skip_col = -1;
   a.all_missing = 0; a.type20 = 0;
   for (int c = 4; c < 13; c++) {
      bool all = true;
      for (size_t i = 0; i < n && all; i++) all = (ts_orig[i + n * c] == -999.0);
      if (all) a.all_missing |= (1u << c);
      if (disagg_method == 2) a.type20 |= (1u << c);
   }
   struct Scratch {   // freed on every path out of this function
      double *slow = nullptr, *fast = nullptr;
      ~Scratch() { cudaFree(slow); cudaFree(fast); }
   } sc;
   CUDA_OK(dev_alloc(&sc.slow, n * 13));
   CUDA_OK(dev_alloc(&sc.fast, nfast * 13));
   a.slow = sc.slow; a.fast = sc.fast;
   CUDA_OK(cudaMemcpyAsync(sc.slow, ts_orig, sizeof(double) * n * 13, cudaMemcpyHostToDevice, h->stream));
   suews_disagg_kernel<<<(unsigned)((nfast + 255) / 256), 256, 0, h->stream>>>(a);
   CUDA_OK(cudaGetLastError());
   h->ts5_block.resize(nfast * 13);
   CUDA_OK(cudaMemcpyAsync(h->ts5_block.data(), sc.fast, sizeof(double) * nfast * 13, cudaMemcpyDeviceToHost, h->stream));
   CUDA_OK(cudaStreamSynchronize(h->stream));
   // the [nsteps][9] layout the timestep kernel reads (a few hundred KB: repacked through the host copy kept for readback)
   return suews_b200_set_estm_forcing(h, h->ts5_block.data(), (int)nfast);
}

// the disaggregated surface-temperature block [nsteps x 13], column-major
int suews_b200_get_estm_forcing_block(suews_b200_handle h, double *ts_fast, int nsteps)
{
   if (!h || !ts_fast || h->ts5_block.size() != (size_t)nsteps * 13)
      return set_err(SUEWS_E_BAD_ARG, "get_estm_forcing_block: no device-disaggregated block of %d rows", nsteps);
   memcpy(ts_fast, h->ts5_block.data(), sizeof(double) * (size_t)nsteps * 13);
   return SUEWS_OK;
}

int suews_b200_set_estm_forcing(suews_b200_handle h, const double *ts5, int nsteps)
{
   if (!h || !ts5 || nsteps <= 0) return set_err(SUEWS_E_BAD_ARG, "set_estm_forcing: bad arguments");
   if (nsteps != h->nsteps)
      return set_err(SUEWS_E_BAD_ARG, "set_estm_forcing: %d rows, but the forcing block has %d (call set_forcing first)", nsteps, h->nsteps);
   CUDA_OK(cudaSetDevice(h->device));
   const size_t n = (size_t)nsteps;
   h->ts5.resize(n * SUEWS_N_TS5);
   for (size_t i = 0; i < n; i++)
      for (int c = 0; c < SUEWS_N_TS5; c++) h->ts5[i * SUEWS_N_TS5 + c] = ts5[i + n * (size_t)(4 + c)];   // Ts5mindata_ir(5:13)
   CUDA_OK(dev_alloc(&h->d_ts5, n * SUEWS_N_TS5));
   CUDA_OK(cudaMemcpy(h->d_ts5, h->ts5.data(), sizeof(double) * n * SUEWS_N_TS5, cudaMemcpyHostToDevice));
   h->ts5_nsteps = nsteps;
   return SUEWS_OK;
}

int suews_b200_advance(suews_b200_handle h, int step0, int nsteps, int out_flags)
{
   if (!h || !h->uploaded || !h->d_steps) return set_err(SUEWS_E_BAD_ARG, "advance: upload parameters and forcing first");
   if (step0 < 0 || nsteps <= 0 || step0 + nsteps > h->nsteps)
      return set_err(SUEWS_E_BAD_ARG, "advance: rows [%d,%d) outside the %d forcing rows", step0, step0 + nsteps, h->nsteps);
   CUDA_OK(cudaSetDevice(h->device));
   const bool full = (out_flags & SUEWS_OUT_FULL) != 0;
   if (full) {
      const int ncol[7] = {SUEWS_NCOL_SUEWS + 5, SUEWS_NCOL_SNOW + 5, SUEWS_NCOL_RSL + 5, SUEWS_NCOL_DAILY + 5, SUEWS_NCOL_SOLWEIG + 5,
                           SUEWS_NCOL_ESTM + 5, SUEWS_NCOL_BL + 5};
      bool any_cbl = false;
      for (const auto &P : h->classes) any_cbl = any_cbl || P.CBLuse >= 1;
      for (int i = 0; i < 7; i++) {
         if (i == 5 && !h->d_estm) continue;   // the ESTM block exists only when ESTM runs
         if (i == 6 && !any_cbl) continue;     // ... and the BL block only with BLUEWS
         size_t cnt = (size_t)h->ncell * (size_t)nsteps * (size_t)ncol[i];
         if (h->blk_cap[i] < cnt) {
            CUDA_OK(dev_alloc(&h->d_blk[i], cnt));
            h->blk_cap[i] = cnt;
         }
         // unwritten rows (snow block when snowUse = 0) are left unassigned by the reference; -999 here
         suews_fill_kernel<<<(unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 16), 256, 0, h->stream>>>(h->d_blk[i], cnt, -999.0);
      }
      CUDA_OK(cudaGetLastError());
      h->blk_nsteps = nsteps;
   }
   SuRunArgs a;
   a.agg = nullptr; a.agg_nlines = 1; a.agg_daily = nullptr; a.agg_daily_rows = 0;
   if (out_flags & SUEWS_OUT_AGG) {
      if (out_flags & (SUEWS_OUT_SUM82 | SUEWS_OUT_SUM7))
         return set_err(SUEWS_E_BAD_ARG, "advance: SUEWS_OUT_AGG cannot be combined with the running-sum policies");
      // nlinesOut = INT(nsh/(60.*60/outFreq_s)), suews_ctrl_output.f95:922
      const int nsh = 3600 / h->tstep;
      const int nlines = (int)((float)nsh / (60.f * 60 / (float)h->out_freq_s));
      if (nlines < 1) return set_err(SUEWS_E_BAD_ARG, "advance: output frequency %d s is finer than the timestep", h->out_freq_s);
      const int rows = nsteps / nlines;
      const size_t need = (size_t)(rows ? rows : 1) * (SUEWS_NCOL_SUEWS + 5) * (size_t)h->stride;
      if (h->agg_cap < need) {
         if (h->d_agg) { cudaFree(h->d_agg); h->d_agg = nullptr; }
         h->agg_cap = 0;
         if (cudaMalloc((void **)&h->d_agg, sizeof(double) * need) != cudaSuccess) {
            cudaGetLastError();
            return set_err(SUEWS_E_BAD_ARG, "advance: the aggregated block needs %.1f GB of device memory (%d rows x 87 x %d cells)",
                           (double)need * 8e-9, rows, h->ncell);
         }
         h->agg_cap = need;
      }
      {   // DailyState rows: one per day that ends inside [step0, step0 + nsteps)
         std::vector<SuewsStepRec> recs((size_t)nsteps);
         CUDA_OK(cudaMemcpy(recs.data(), h->d_steps + step0, sizeof(SuewsStepRec) * (size_t)nsteps, cudaMemcpyDeviceToHost));
         int nd = 0;
         for (const auto &r : recs) nd += (r.it * 3600 + r.imin * 60 + h->tstep) >= 86400;
         const size_t need_d = (size_t)(nd ? nd : 1) * (SUEWS_NCOL_DAILY + 5) * (size_t)h->stride;
         if (h->agg_daily_cap < need_d) {
            if (h->d_agg_daily) { cudaFree(h->d_agg_daily); h->d_agg_daily = nullptr; }
            h->agg_daily_cap = 0;
            CUDA_OK(cudaMalloc((void **)&h->d_agg_daily, sizeof(double) * need_d));
            h->agg_daily_cap = need_d;
         }
         CUDA_OK(cudaMemsetAsync(h->d_agg_daily, 0xFF, sizeof(double) * need_d, h->stream));
         h->agg_daily_rows = nd;
         a.agg_daily = h->d_agg_daily;
         a.agg_daily_rows = nd;
      }
      h->agg_rows = rows;
      // rows of cells that stop on a fatal error are never written: they read as NaN, not as a previous launch's values
      CUDA_OK(cudaMemsetAsync(h->d_agg, 0xFF, sizeof(double) * need, h->stream));
      a.agg = h->d_agg;
      a.agg_nlines = nlines;
   }
   a.classes = h->d_classes; a.class_of_slot = h->d_cls; a.tz_of_slot = h->d_tz; a.order = h->d_order; a.cellp = h->d_cellp; a.state = h->d_state;
   a.steps = h->d_steps + step0;
   a.ephem = h->d_ephem + step0;
   a.ephem_tz_stride = h->nsteps;
   a.sums = (out_flags & (SUEWS_OUT_SUM82 | SUEWS_OUT_SUM7)) ? h->d_sums : nullptr;
   a.sum7 = ((out_flags & SUEWS_OUT_SUM7) && !(out_flags & SUEWS_OUT_SUM82)) ? 1 : 0;
   // policy P0 observes seven flux sums only: stages nothing requested depends on are skipped (SUEWS_OUT_ALL_STAGES keeps them)
   a.flux_only = ((a.sum7 || !a.sums) && !full && !(out_flags & (SUEWS_OUT_AGG | SUEWS_OUT_ALL_STAGES))) ? 1 : 0;
   a.fatal = h->d_fatal; a.warn = h->d_warn; a.iters = h->d_iters;
   a.blkSUEWS = h->d_blk[0]; a.blkSnow = h->d_blk[1]; a.blkRSL = h->d_blk[2]; a.blkDaily = h->d_blk[3]; a.blkSOL = h->d_blk[4];
   a.blkESTM = h->d_estm ? h->d_blk[5] : nullptr;
   a.estm = h->d_estm;
   a.an_hours = h->d_an_hours;
   a.an_day = h->d_an_day;
   a.cbl_ini = h->d_cbl_ini;
   a.cbl_nline = h->cbl_nline;
   a.blkBL = (full && h->d_estm) ? h->d_blk[6] : nullptr;
   a.ts5 = (h->d_estm && h->ts5_nsteps == h->nsteps) ? h->d_ts5 + (size_t)step0 * SUEWS_N_TS5 : nullptr;
   a.ncell = h->ncell; a.stride = h->stride; a.nsteps = nsteps;
   a.dt_since_start = h->dt_since_start; a.tstep_prev = h->tstep_prev; a.tstep = h->tstep;
   a.roll_tstep_prev = h->roll_tstep_prev ? 1 : 0;
   int nsm = 148;
   cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
   // One launch per non-empty bucket, each with the kernel instance compiled for the bucket's switches and the requested
   // output kind (suews_kernel.cuh); the buckets of an advance run side by side on their own streams.
   const int kind = full ? SU_K_FULL : (((out_flags & (SUEWS_OUT_SUM82 | SUEWS_OUT_AGG)) || (out_flags & SUEWS_OUT_ALL_STAGES)) ? SU_K_ACC : SU_K_P0);
   int nb = 0;
   for (int b = 0; b < 16; b++) nb += (h->bucket_start[b + 1] > h->bucket_start[b]);
   CUDA_OK(cudaEventRecord(h->ev0, h->stream));
   if (nb > 1) {
      if (!h->ev_fork) CUDA_OK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
      CUDA_OK(cudaEventRecord(h->ev_fork, h->stream));
   }
   int side = 0;
   for (int b = 0; b < 16; b++) {
      const int n = h->bucket_start[b + 1] - h->bucket_start[b];
      if (n <= 0) continue;
      a.slot0 = h->bucket_start[b];
      a.nslot = n;
      const bool estm = (b & 4) != 0, snow1 = (b & 2) != 0, snow_rt = (b & 8) != 0;
      // launch geometry: 1024-cell blocks, one per SM, once they can occupy most of the chip; 256-cell blocks below that.
      // measured on B200 (cell-steps/s, large vs small geometry): 10 240 cells 0.87e7 / 1.7e7 (full output), 37 888 cells
      // 6.4e7 / 1.05e8, 75 776 cells 1.23e8 / 1.11e8 -- the crossover sits near 1.5 small blocks per SM
      bool small = (long long)n * 2 <= (long long)nsm * SU_BLOCK_SMALL * 3;
      if (const char *e = getenv("SUEWS_B200_GEOMETRY")) small = (strcmp(e, "small") == 0) ? true : ((strcmp(e, "large") == 0) ? false : small);
      cudaError_t (*launch)(int, cudaStream_t, const SuRunArgs &) = nullptr;
      if (kind == SU_K_FULL) launch = small ? SU_LAUNCHER_NAME(0, m1, m1, 256) : SU_LAUNCHER_NAME(0, m1, m1, 1024);
      else if (kind == SU_K_P0 && !small && estm && !snow1 && !snow_rt) launch = SU_LAUNCHER_NAME(2, 0, m1, 1024);
      else if (small || estm || snow_rt) launch = small ? SU_LAUNCHER_NAME(1, m1, m1, 256) : SU_LAUNCHER_NAME(1, m1, m1, 1024);
      else if (kind == SU_K_ACC) launch = snow1 ? SU_LAUNCHER_NAME(1, m1, m1, 1024) : SU_LAUNCHER_NAME(1, 0, 0, 1024);
      else launch = snow1 ? SU_LAUNCHER_NAME(2, 1, 0, 1024) : SU_LAUNCHER_NAME(2, 0, 0, 1024);
      cudaStream_t st = h->stream;
      if (nb > 1 && side > 0) {   // first bucket on the main stream, the others fork off it and join again below
         const int k = side - 1;
         if (!h->bstream[k]) {
            CUDA_OK(cudaStreamCreateWithFlags(&h->bstream[k], cudaStreamNonBlocking));
            CUDA_OK(cudaEventCreateWithFlags(&h->bev[k], cudaEventDisableTiming));
         }
         CUDA_OK(cudaStreamWaitEvent(h->bstream[k], h->ev_fork, 0));
         st = h->bstream[k];
      }
      CUDA_OK(launch(n, st, a));
      if (st != h->stream) {
         CUDA_OK(cudaEventRecord(h->bev[side - 1], st));
      }
      side++;
   }
   for (int k = 0; k + 1 < side; k++) CUDA_OK(cudaStreamWaitEvent(h->stream, h->bev[k], 0));
   CUDA_OK(cudaGetLastError());
   CUDA_OK(cudaEventRecord(h->ev1, h->stream));
   h->dt_since_start += nsteps * h->tstep;
   if (h->roll_tstep_prev) h->tstep_prev = h->tstep;
   h->dev_dirty = true;
   return SUEWS_OK;
}

int suews_b200_set_output_freq(suews_b200_handle h, int out_freq_s)
{
   if (!h || out_freq_s <= 0) return set_err(SUEWS_E_BAD_ARG, "set_output_freq: bad arguments");
   h->out_freq_s = out_freq_s;
   return SUEWS_OK;
}

// One cell's aggregated SUEWS block [nrows x 87], column-major, as SUEWS_Output_Agg returns it
int suews_b200_get_agg_block(suews_b200_handle h, int cell, double *block, int nrows)
{
   if (!h || !block || !h->d_agg || nrows != h->agg_rows || cell < 0 || cell >= h->ncell)
      return set_err(SUEWS_E_BAD_SIZE, "get_agg_block: the aggregated block is [%d x 87] per cell", h ? h->agg_rows : 0);
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaStreamSynchronize(h->stream));
   const int ncol = SUEWS_NCOL_SUEWS + 5;
   if (nrows == 0) return SUEWS_OK;   // fewer steps than one output period
   std::vector<double> tmp((size_t)nrows * ncol);   // [row][col]
   CUDA_OK(cudaMemcpy2D(tmp.data(), sizeof(double), h->d_agg + h->slot_of_cell[(size_t)cell], sizeof(double) * (size_t)h->stride,
                        sizeof(double), (size_t)nrows * ncol, cudaMemcpyDeviceToHost));
   for (int r = 0; r < nrows; r++)
      for (int c = 0; c < ncol; c++) block[(size_t)r + (size_t)nrows * c] = tmp[(size_t)r * ncol + c];
   return SUEWS_OK;
}

// One cell's DailyState rows of the last SUEWS_OUT_AGG launch [nrows x 50], column-major
int suews_b200_get_agg_daily_block(suews_b200_handle h, int cell, double *block, int nrows)
{
   if (!h || !block || !h->d_agg_daily || nrows != h->agg_daily_rows || cell < 0 || cell >= h->ncell)
      return set_err(SUEWS_E_BAD_SIZE, "get_agg_daily_block: the DailyState block is [%d x 50] per cell", h ? h->agg_daily_rows : 0);
   if (nrows == 0) return SUEWS_OK;
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaStreamSynchronize(h->stream));
   const int ncol = SUEWS_NCOL_DAILY + 5;
   std::vector<double> tmp((size_t)nrows * ncol);
   CUDA_OK(cudaMemcpy2D(tmp.data(), sizeof(double), h->d_agg_daily + h->slot_of_cell[(size_t)cell], sizeof(double) * (size_t)h->stride,
                        sizeof(double), (size_t)nrows * ncol, cudaMemcpyDeviceToHost));
   for (int r = 0; r < nrows; r++)
      for (int c = 0; c < ncol; c++) block[(size_t)r + (size_t)nrows * c] = tmp[(size_t)r * ncol + c];
   return SUEWS_OK;
}
int suews_b200_agg_daily_rows(suews_b200_handle h, int *nrows)
{
   if (!h || !nrows) return set_err(SUEWS_E_BAD_ARG, "null argument");
   *nrows = h->agg_daily_rows;
   return SUEWS_OK;
}

// One column of one aggregated row for every cell (a map), [ncell]
int suews_b200_get_agg_column(suews_b200_handle h, int row, int col, double *vals)
{
   if (!h || !vals || !h->d_agg || row < 0 || row >= h->agg_rows || col < 0 || col >= SUEWS_NCOL_SUEWS + 5)
      return set_err(SUEWS_E_BAD_SIZE, "get_agg_column: row/column outside the [%d x 87] aggregated block", h ? h->agg_rows : 0);
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaStreamSynchronize(h->stream));
   std::vector<double> tmp((size_t)h->ncell);
   CUDA_OK(cudaMemcpy(tmp.data(), h->d_agg + ((size_t)row * (SUEWS_NCOL_SUEWS + 5) + col) * (size_t)h->stride,
                      sizeof(double) * (size_t)h->ncell, cudaMemcpyDeviceToHost));
   for (int c = 0; c < h->ncell; c++) vals[c] = tmp[(size_t)h->slot_of_cell[(size_t)c]];
   return SUEWS_OK;
}

int suews_b200_sync(suews_b200_handle h)
{
   if (!h) return set_err(SUEWS_E_BAD_ARG, "null handle");
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaStreamSynchronize(h->stream));
   std::vector<int> fatal((size_t)h->ncell);   // slot order on the device; the lowest failing cell index is reported
   CUDA_OK(cudaMemcpy(fatal.data(), h->d_fatal, sizeof(int) * (size_t)h->ncell, cudaMemcpyDeviceToHost));
   int best = -1, code = 0;
   for (int sl = 0; sl < h->ncell; sl++)
      if (fatal[(size_t)sl] != 0 && (best < 0 || h->cell_of_slot[(size_t)sl] < best)) { best = h->cell_of_slot[(size_t)sl]; code = fatal[(size_t)sl]; }
   if (best >= 0) {
      h->first_fatal_code = code;
      h->first_fatal_cell = best;
      return set_err(code, "Fatal error in SUEWS! code %d in cell %d: %s", code, best, fatal_text(code));
   }
   return SUEWS_OK;
}

int suews_b200_last_advance_ms(suews_b200_handle h, float *ms)
{
   if (!h || !ms) return set_err(SUEWS_E_BAD_ARG, "null argument");
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaEventSynchronize(h->ev1));
   CUDA_OK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
   return SUEWS_OK;
}

int suews_b200_get_sums(suews_b200_handle h, double *sums)
{
   if (!h || !h->uploaded) return set_err(SUEWS_E_BAD_ARG, "get_sums: nothing uploaded");
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaStreamSynchronize(h->stream));
   std::vector<double> tmp((size_t)SUEWS_NCOL_SUEWS * (size_t)h->stride);
   CUDA_OK(cudaMemcpy(tmp.data(), h->d_sums, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
   for (int c = 0; c < h->ncell; c++) {
      const size_t sl = (size_t)h->slot_of_cell[(size_t)c];
      for (int k = 0; k < SUEWS_NCOL_SUEWS; k++) sums[(size_t)c * SUEWS_NCOL_SUEWS + k] = tmp[(size_t)k * h->stride + sl];
   }
   return SUEWS_OK;
}

// Selected columns of the running sums, [ncell][ncols] cell-major.  The slot -> cell permutation is done on the device
// (suews_pack_sums_kernel: contiguous reads, scattered writes into a compact cell-ordered block), one copy brings the block
// to pinned host memory, and the host only transposes [ncols][ncell] -> [ncell][ncols] with sequential reads -- the
// per-cell gather through slot_of_cell on the host was a third of the end-to-end overhead of a bench step.
int suews_b200_get_sums_cols(suews_b200_handle h, const int *cols, int ncols, double *out)
{
   if (!h || !h->uploaded || !cols || !out || ncols <= 0 || ncols > 16) return set_err(SUEWS_E_BAD_ARG, "get_sums_cols: 1..16 columns");
   CUDA_OK(cudaSetDevice(h->device));
   SuPackCols pc;
   pc.ncols = ncols;
   for (int k = 0; k < ncols; k++) {
      if (cols[k] < 0 || cols[k] >= SUEWS_NCOL_SUEWS) return set_err(SUEWS_E_BAD_ARG, "get_sums_cols: column %d", cols[k]);
      pc.col[k] = cols[k];
   }
   const size_t n = (size_t)h->ncell, total = (size_t)ncols * n;
   const size_t need = sizeof(double) * total;
   if (h->pinned_bytes < need) {
      if (h->pinned) cudaFreeHost(h->pinned);
      h->pinned = nullptr;
      h->pinned_bytes = 0;
      CUDA_OK(cudaHostAlloc(&h->pinned, need, cudaHostAllocDefault));
      h->pinned_bytes = need;
   }
   if (h->pack_cap < total) {
      if (h->d_pack) cudaFree(h->d_pack);
      h->d_pack = nullptr;
      h->pack_cap = 0;
      CUDA_OK(cudaMalloc((void **)&h->d_pack, need));
      h->pack_cap = total;
   }
   suews_pack_sums_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_sums, h->d_order, h->stride, (int)n, pc, h->d_pack);
   CUDA_OK(cudaGetLastError());
   double *pin = (double *)h->pinned;
   CUDA_OK(cudaMemcpyAsync(pin, h->d_pack, need, cudaMemcpyDeviceToHost, h->stream));
   CUDA_OK(cudaStreamSynchronize(h->stream));
   for (size_t c0 = 0; c0 < n; c0 += 4096) {   // blocked transpose: every source column is read sequentially
      const size_t c1 = std::min(n, c0 + 4096);
      for (int k = 0; k < ncols; k++) {
         const double *src = pin + (size_t)k * n;
         for (size_t c = c0; c < c1; c++) out[c * (size_t)ncols + k] = src[c];
      }
   }
   return SUEWS_OK;
}

int suews_b200_device_sync(suews_b200_handle h)
{
   if (!h) return set_err(SUEWS_E_BAD_ARG, "null handle");
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaDeviceSynchronize());
   return SUEWS_OK;
}

// Measured FP64 FMA throughput of this device [TFLOP/s]: the compute-roofline denominator
// (MEASURED_PEAKS.json has no FP64 figure).  8 independent DFMA chains per thread.
int suews_b200_fp64_peak(int device, double *tflops)
{
   CUDA_OK(cudaSetDevice(device));
   cudaDeviceProp prop;
   CUDA_OK(cudaGetDeviceProperties(&prop, device));
   const int iters = 1 << 14, threads = 256, blocks = prop.multiProcessorCount * 8;
   double *d = nullptr;
   CUDA_OK(cudaMalloc((void **)&d, sizeof(double) * (size_t)threads * blocks));
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0);
   cudaEventCreate(&e1);
   float best = 1e30f;
   for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      suews_dfma_kernel<<<blocks, threads>>>(d, iters, 1.0000001);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
   }
   cudaEventDestroy(e0);
   cudaEventDestroy(e1);
   cudaFree(d);
   CUDA_OK(cudaGetLastError());
   double flops = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks;
   *tflops = flops / (best * 1e-3) / 1e12;
   return SUEWS_OK;
}

int suews_b200_reset_sums(suews_b200_handle h)
{
   if (!h || !h->uploaded) return set_err(SUEWS_E_BAD_ARG, "reset_sums: nothing uploaded");
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaMemsetAsync(h->d_sums, 0, sizeof(double) * (size_t)SUEWS_NCOL_SUEWS * (size_t)h->stride, h->stream));
   return SUEWS_OK;
}

int suews_b200_get_block(suews_b200_handle h, const char *group, int cell, double *block, int nsteps, int ncol)
{
   if (!h || !group || !block) return set_err(SUEWS_E_BAD_ARG, "null argument");
   const char *names[7] = {"SUEWS", "snow", "RSL", "DailyState", "SOLWEIG", "ESTM", "BL"};
   const int ncols[7] = {SUEWS_NCOL_SUEWS + 5, SUEWS_NCOL_SNOW + 5, SUEWS_NCOL_RSL + 5, SUEWS_NCOL_DAILY + 5, SUEWS_NCOL_SOLWEIG + 5,
                         SUEWS_NCOL_ESTM + 5, SUEWS_NCOL_BL + 5};
   int g = -1;
   for (int i = 0; i < 7; i++)
      if (strcasecmp(group, names[i]) == 0) g = i;
   if (g < 0) return set_err(SUEWS_E_BAD_NAME, "get_block: unknown group '%s'", group);
   if (!h->d_blk[g] || nsteps != h->blk_nsteps || ncol != ncols[g] || cell < 0 || cell >= h->ncell)
      return set_err(SUEWS_E_BAD_SIZE, "get_block: %s block is [%d x %d] per cell", names[g], h->blk_nsteps, ncols[g]);
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaStreamSynchronize(h->stream));
   size_t cnt = (size_t)nsteps * (size_t)ncol;
   CUDA_OK(cudaMemcpy(block, h->d_blk[g] + (size_t)cell * cnt, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
   return SUEWS_OK;
}

int suews_b200_get_diag(suews_b200_handle h, int *fatal, unsigned long long *warn2, long long *iters)
{
   if (!h || !h->uploaded) return set_err(SUEWS_E_BAD_ARG, "get_diag: nothing uploaded");
   CUDA_OK(cudaSetDevice(h->device));
   CUDA_OK(cudaStreamSynchronize(h->stream));
   const size_t n = (size_t)h->ncell;   // device arrays are in slot order
   if (fatal) {
      std::vector<int> tmp(n);
      CUDA_OK(cudaMemcpy(tmp.data(), h->d_fatal, sizeof(int) * n, cudaMemcpyDeviceToHost));
      for (size_t sl = 0; sl < n; sl++) fatal[(size_t)h->cell_of_slot[sl]] = tmp[sl];
   }
   if (iters) {
      std::vector<long long> tmp(n);
      CUDA_OK(cudaMemcpy(tmp.data(), h->d_iters, sizeof(long long) * n, cudaMemcpyDeviceToHost));
      for (size_t sl = 0; sl < n; sl++) iters[(size_t)h->cell_of_slot[sl]] = tmp[sl];
   }
   if (warn2) {
      std::vector<unsigned long long> tmp(2 * (size_t)h->stride);
      CUDA_OK(cudaMemcpy(tmp.data(), h->d_warn, sizeof(unsigned long long) * tmp.size(), cudaMemcpyDeviceToHost));
      for (size_t sl = 0; sl < n; sl++) {
         const size_t c = (size_t)h->cell_of_slot[sl];
         warn2[2 * c] = tmp[sl];
         warn2[2 * c + 1] = tmp[(size_t)h->stride + sl];
      }
   }
   return SUEWS_OK;
}

#ifdef SU_PHASE_CLOCKS
// profiling builds only: read (and optionally clear) the per-phase cycle counters of dev_step.cuh
// (every kernel instance is its own translation unit with its own copy of the counters: each registers a reader)
static su_clk_reader_t su_clk_readers[32];
static int su_n_clk_readers = 0;
void su_register_clk_reader(su_clk_reader_t f)
{
   for (int i = 0; i < su_n_clk_readers; i++)
      if (su_clk_readers[i] == f) return;
   if (su_n_clk_readers < 32) su_clk_readers[su_n_clk_readers++] = f;
}
int suews_b200_debug_phase_clocks(unsigned long long *out16, int reset)
{
   CUDA_OK(cudaDeviceSynchronize());
   if (out16) for (int k = 0; k < 16; k++) out16[k] = 0;
   for (int i = 0; i < su_n_clk_readers; i++) {
      unsigned long long t[16] = {0};
      CUDA_OK(su_clk_readers[i](t, reset));
      if (out16) for (int k = 0; k < 16; k++) out16[k] += t[k];
   }
   return SUEWS_OK;
}
#endif

int suews_b200_device_ptrs(suews_b200_handle h, void **d_state, void **d_sums, long long *cell_stride)
{
   if (!h || !h->uploaded) return set_err(SUEWS_E_BAD_ARG, "device_ptrs: nothing uploaded");
   if (d_state) *d_state = h->d_state;
   if (d_sums) *d_sums = h->d_sums;
   if (cell_stride) *cell_stride = h->stride;
   return SUEWS_OK;
}

int suews_b200_selftest_math(long long n, unsigned long long seed, long long *bad_exp, long long *bad_log)
{
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_err(SUEWS_E_NO_DEVICE, "no CUDA device");
   unsigned long long *d = nullptr, h[2] = {0, 0};
   CUDA_OK(cudaMalloc(&d, sizeof(h)));
   CUDA_OK(cudaMemset(d, 0, sizeof(h)));
   suews_math_selftest_kernel<<<148 * 4, 256>>>(n, seed, d);
   CUDA_OK(cudaGetLastError());
   CUDA_OK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
   cudaFree(d);
   if (bad_exp) *bad_exp = (long long)h[0];
   if (bad_log) *bad_log = (long long)h[1];
   return SUEWS_OK;
}

int suews_b200_selftest_latvap(long long n, unsigned long long seed, long long *bad)
{
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_err(SUEWS_E_NO_DEVICE, "no CUDA device");
   unsigned long long *d = nullptr, h = 0;
   CUDA_OK(cudaMalloc(&d, sizeof(h)));
   CUDA_OK(cudaMemset(d, 0, sizeof(h)));
   suews_latvap_selftest_kernel<<<148 * 4, 256>>>(n, seed, d);
   CUDA_OK(cudaGetLastError());
   CUDA_OK(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
   cudaFree(d);
   if (bad) *bad = (long long)h;
   return SUEWS_OK;
}

int suews_cal_sunposition_c(double year, double idectime, double UTC, double lat, double lng, double alt, double *az,
                            double *zen)
{
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      return set_err(SUEWS_E_NO_DEVICE, "libsuews_b200 needs a CUDA device (no CPU fallback)");
   double *d = nullptr, hbuf[2];
   CUDA_OK(cudaMalloc((void **)&d, 2 * sizeof(double)));
   suews_sunpos_kernel<<<1, 1>>>(year, idectime, UTC, lat, lng, alt, d);
   cudaError_t e = cudaMemcpy(hbuf, d, sizeof(hbuf), cudaMemcpyDeviceToHost);
   cudaFree(d);
   CUDA_OK(e);
   if (az) *az = hbuf[0];
   if (zen) *zen = hbuf[1];
   return SUEWS_OK;
}

// ------------------------------------------------------------------------------------------------
// several GPUs of one box behind one handle (include/suews_b200.h): contiguous blocks of cells, one per device, one host
// thread, no exchange while the steps advance; the final gather goes device -> first device -> host
// ------------------------------------------------------------------------------------------------
struct suews_b200_multi_s {
   int ncell = 0, nclass = 0, ndev = 0;
   std::vector<suews_b200_handle> part;
   std::vector<int> cell0, n;          // block d covers cells [cell0[d], cell0[d] + n[d])
   std::vector<double *> d_pack;       // per device: compact cell-ordered sums block [ncols][n[d]]
   std::vector<size_t> pack_cap;
   double *d_gather = nullptr;         // on the first device: [ncols][ncell]
   size_t gather_cap = 0;
   double *pinned = nullptr;
   size_t pinned_cap = 0;
};

int suews_b200_multi_create(int ncell, int nclass, int ndev, const int *devices, suews_b200_multi *out)
{
   if (!out || ncell <= 0 || nclass <= 0 || ndev <= 0 || ndev > ncell) return set_err(SUEWS_E_BAD_ARG, "multi_create: bad arguments");
   suews_b200_multi_s *m = new suews_b200_multi_s;
   m->ncell = ncell; m->nclass = nclass; m->ndev = ndev;
   // equal contiguous blocks, the first ncell % ndev of them one cell longer
   int c0 = 0;
   for (int d = 0; d < ndev; d++) {
      const int n = ncell / ndev + (d < ncell % ndev ? 1 : 0);
      suews_b200_handle h = nullptr;
      int rc = suews_b200_create(n, nclass, devices ? devices[d] : d, &h);
      if (rc != SUEWS_OK) {
         for (auto p : m->part) suews_b200_destroy(p);
         delete m;
         return rc;
      }
      m->part.push_back(h);
      m->cell0.push_back(c0);
      m->n.push_back(n);
      c0 += n;
   }
   m->d_pack.assign((size_t)ndev, nullptr);
   m->pack_cap.assign((size_t)ndev, 0);
   // peer access towards the first device (the gather target); a refusal just means the copies are staged by the driver
   for (int d = 1; d < ndev; d++) {
      const int a = m->part[(size_t)d]->device, b = m->part[0]->device;
      if (a == b) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) {
         cudaSetDevice(a);
         if (cudaDeviceEnablePeerAccess(b, 0) != cudaSuccess) cudaGetLastError();
      }
   }
   *out = m;
   return SUEWS_OK;
}

int suews_b200_multi_destroy(suews_b200_multi m)
{
   if (!m) return SUEWS_OK;
   for (int d = 0; d < m->ndev; d++) {
      if (m->d_pack[(size_t)d]) { cudaSetDevice(m->part[(size_t)d]->device); cudaFree(m->d_pack[(size_t)d]); }
   }
   if (m->d_gather) { cudaSetDevice(m->part[0]->device); cudaFree(m->d_gather); }
   if (m->pinned) cudaFreeHost(m->pinned);
   for (auto p : m->part) suews_b200_destroy(p);
   delete m;
   return SUEWS_OK;
}

int suews_b200_multi_ndev(suews_b200_multi m, int *ndev)
{
   if (!m || !ndev) return set_err(SUEWS_E_BAD_ARG, "null argument");
   *ndev = m->ndev;
   return SUEWS_OK;
}

int suews_b200_multi_part(suews_b200_multi m, int d, suews_b200_handle *h, int *cell0, int *ncell)
{
   if (!m || d < 0 || d >= m->ndev) return set_err(SUEWS_E_BAD_ARG, "multi_part: block %d of %d", d, m ? m->ndev : 0);
   if (h) *h = m->part[(size_t)d];
   if (cell0) *cell0 = m->cell0[(size_t)d];
   if (ncell) *ncell = m->n[(size_t)d];
   return SUEWS_OK;
}

#define MULTI_EACH(call)                                   \
   do {                                                    \
      if (!m) return set_err(SUEWS_E_BAD_ARG, "null handle"); \
      for (int d = 0; d < m->ndev; d++) {                  \
         suews_b200_handle h = m->part[(size_t)d];         \
         const size_t c0 = (size_t)m->cell0[(size_t)d];    \
         (void)c0;                                         \
         int rc_ = (call);                                 \
         if (rc_ != SUEWS_OK) return rc_;                  \
      }                                                    \
      return SUEWS_OK;                                     \
   } while (0)

int suews_b200_multi_set_class_param(suews_b200_multi m, int cls, const char *name, const double *vals, int n)
{
   MULTI_EACH(suews_b200_set_class_param(h, cls, name, vals, n));
}
int suews_b200_multi_set_cell_param(suews_b200_multi m, const char *name, const double *vals, int n)
{
   MULTI_EACH(suews_b200_set_cell_param(h, name, vals + c0 * (size_t)n, n));
}
int suews_b200_multi_set_cell_class(suews_b200_multi m, const int *class_of_cell)
{
   MULTI_EACH(suews_b200_set_cell_class(h, class_of_cell + c0));
}
int suews_b200_multi_set_state(suews_b200_multi m, const char *name, const double *vals, int n)
{
   MULTI_EACH(suews_b200_set_state(h, name, vals + c0 * (size_t)n, n));
}
int suews_b200_multi_get_state(suews_b200_multi m, const char *name, double *vals, int n)
{
   MULTI_EACH(suews_b200_get_state(h, name, vals + c0 * (size_t)n, n));
}
int suews_b200_multi_set_clock(suews_b200_multi m, int dt_since_start, int tstep_prev)
{
   MULTI_EACH(suews_b200_set_clock(h, dt_since_start, tstep_prev));
}
int suews_b200_multi_upload(suews_b200_multi m) { MULTI_EACH(suews_b200_upload(h)); }
int suews_b200_multi_set_forcing(suews_b200_multi m, const double *met, int nsteps) { MULTI_EACH(suews_b200_set_forcing(h, met, nsteps)); }
// the forcing blocks are shared by the batch: every device gets its own copy of the same rows
int suews_b200_multi_set_forcing_orig(suews_b200_multi m, const double *met_orig, int nrows, int resolution_in_s, int disagg_method,
                                      int rain_disagg_method, int raw_units, int smd_method)
{
   MULTI_EACH(suews_b200_set_forcing_orig(h, met_orig, nrows, resolution_in_s, disagg_method, rain_disagg_method, raw_units, smd_method));
}
int suews_b200_multi_apply_kdown_zen(suews_b200_multi m, double lat, double lng, double timezone, double alt)
{
   MULTI_EACH(suews_b200_apply_kdown_zen(h, lat, lng, timezone, alt));
}
int suews_b200_multi_set_estm_forcing(suews_b200_multi m, const double *ts5, int nsteps) { MULTI_EACH(suews_b200_set_estm_forcing(h, ts5, nsteps)); }
int suews_b200_multi_set_estm_forcing_orig(suews_b200_multi m, const double *ts_orig, int nrows, int resolution_in_s, int disagg_method)
{
   MULTI_EACH(suews_b200_set_estm_forcing_orig(h, ts_orig, nrows, resolution_in_s, disagg_method));
}
int suews_b200_multi_set_anohm_forcing(suews_b200_multi m, const double *met, int nrows) { MULTI_EACH(suews_b200_set_anohm_forcing(h, met, nrows)); }
int suews_b200_multi_set_cbl_initial(suews_b200_multi m, const double *ini, int nline) { MULTI_EACH(suews_b200_set_cbl_initial(h, ini, nline)); }
// one cell's per-step block (policy FULL) from whichever device holds it; `cell` is the index of the whole run
int suews_b200_multi_get_block(suews_b200_multi m, const char *group, int cell, double *block, int nsteps, int ncol)
{
   if (!m || cell < 0 || cell >= m->ncell) return set_err(SUEWS_E_BAD_ARG, "multi_get_block: cell %d of %d", cell, m ? m->ncell : 0);
   for (int d = 0; d < m->ndev; d++)
      if (cell >= m->cell0[(size_t)d] && cell < m->cell0[(size_t)d] + m->n[(size_t)d])
         return suews_b200_get_block(m->part[(size_t)d], group, cell - m->cell0[(size_t)d], block, nsteps, ncol);
   return set_err(SUEWS_E_BAD_ARG, "multi_get_block: cell %d is in no block", cell);
}
// every launch is asynchronous on its device's stream: one host thread keeps all devices busy
int suews_b200_multi_advance(suews_b200_multi m, int step0, int nsteps, int out_flags) { MULTI_EACH(suews_b200_advance(h, step0, nsteps, out_flags)); }
int suews_b200_multi_sync(suews_b200_multi m)
{
   if (!m) return set_err(SUEWS_E_BAD_ARG, "null handle");
   int first = SUEWS_OK;
   for (int d = 0; d < m->ndev; d++) {   // wait for every device before reporting the first failure
      int rc = suews_b200_sync(m->part[(size_t)d]);
      if (rc != SUEWS_OK && first == SUEWS_OK) {
         first = rc;
         if (m->part[(size_t)d]->first_fatal_cell >= 0)   // the cell index of the whole run, not of the block
            set_err(rc, "Fatal error in SUEWS! code %d in cell %d: %s", rc, m->cell0[(size_t)d] + m->part[(size_t)d]->first_fatal_cell, fatal_text(rc));
      }
   }
   return first;
}
int suews_b200_multi_last_advance_ms(suews_b200_multi m, float *ms)
{
   if (!m || !ms) return set_err(SUEWS_E_BAD_ARG, "null argument");
   float worst = 0;
   for (int d = 0; d < m->ndev; d++) {
      float t = 0;
      int rc = suews_b200_last_advance_ms(m->part[(size_t)d], &t);
      if (rc != SUEWS_OK) return rc;
      worst = std::max(worst, t);
   }
   *ms = worst;
   return SUEWS_OK;
}
int suews_b200_multi_get_diag(suews_b200_multi m, int *fatal, unsigned long long *warn2, long long *iters)
{
   MULTI_EACH(suews_b200_get_diag(h, fatal ? fatal + c0 : nullptr, warn2 ? warn2 + 2 * c0 : nullptr, iters ? iters + c0 : nullptr));
}

int suews_b200_multi_gather_sums(suews_b200_multi m, const int *cols, int ncols, double *out, float *gather_ms)
{
   if (!m || !cols || !out || ncols <= 0 || ncols > 16) return set_err(SUEWS_E_BAD_ARG, "multi_gather_sums: 1..16 columns");
   SuPackCols pc;
   pc.ncols = ncols;
   for (int k = 0; k < ncols; k++) {
      if (cols[k] < 0 || cols[k] >= SUEWS_NCOL_SUEWS) return set_err(SUEWS_E_BAD_ARG, "multi_gather_sums: column %d", cols[k]);
      pc.col[k] = cols[k];
   }
   // the gather is timed on its own: wait for the launches that are still running first
   for (int d = 0; d < m->ndev; d++) {
      CUDA_OK(cudaSetDevice(m->part[(size_t)d]->device));
      CUDA_OK(cudaStreamSynchronize(m->part[(size_t)d]->stream));
   }
   const auto t0 = std::chrono::steady_clock::now();
   suews_b200_handle h0 = m->part[0];
   const size_t total = (size_t)ncols * (size_t)m->ncell;
   CUDA_OK(cudaSetDevice(h0->device));
   if (m->gather_cap < total) {
      if (m->d_gather) cudaFree(m->d_gather);
      m->d_gather = nullptr;
      CUDA_OK(cudaMalloc((void **)&m->d_gather, sizeof(double) * total));
      m->gather_cap = total;
   }
   if (m->pinned_cap < total) {
      if (m->pinned) cudaFreeHost(m->pinned);
      m->pinned = nullptr;
      CUDA_OK(cudaHostAlloc((void **)&m->pinned, sizeof(double) * total, cudaHostAllocDefault));
      m->pinned_cap = total;
   }
   // 1. every device packs its block in cell order, then sends it column by column to the first device.  Everything is
   //    ordered on the source device's launch stream (behind the last advance), so nothing here blocks the host.
   for (int d = 0; d < m->ndev; d++) {
      suews_b200_handle h = m->part[(size_t)d];
      if (!h->uploaded) return set_err(SUEWS_E_BAD_ARG, "multi_gather_sums: nothing uploaded");
      const size_t n = (size_t)m->n[(size_t)d], need = (size_t)ncols * n;
      CUDA_OK(cudaSetDevice(h->device));
      if (m->pack_cap[(size_t)d] < need) {
         if (m->d_pack[(size_t)d]) cudaFree(m->d_pack[(size_t)d]);
         m->d_pack[(size_t)d] = nullptr;
         CUDA_OK(cudaMalloc((void **)&m->d_pack[(size_t)d], sizeof(double) * need));
         m->pack_cap[(size_t)d] = need;
      }
      suews_pack_sums_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_sums, h->d_order, h->stride, (int)n, pc, m->d_pack[(size_t)d]);
      CUDA_OK(cudaGetLastError());
      for (int k = 0; k < ncols; k++)
         CUDA_OK(cudaMemcpyPeerAsync(m->d_gather + (size_t)k * m->ncell + (size_t)m->cell0[(size_t)d], h0->device,
                                     m->d_pack[(size_t)d] + (size_t)k * n, h->device, sizeof(double) * n, h->stream));
   }
   // 2. the first device waits for all blocks, then one copy to the host
   for (int d = 0; d < m->ndev; d++) {
      CUDA_OK(cudaSetDevice(m->part[(size_t)d]->device));
      CUDA_OK(cudaStreamSynchronize(m->part[(size_t)d]->stream));
   }
   CUDA_OK(cudaSetDevice(h0->device));
   CUDA_OK(cudaMemcpyAsync(m->pinned, m->d_gather, sizeof(double) * total, cudaMemcpyDeviceToHost, h0->stream));
   CUDA_OK(cudaStreamSynchronize(h0->stream));
   // [ncols][ncell] -> [ncell][ncols], the layout of suews_b200_get_sums_cols (blocked: sequential reads, cache-sized writes)
   for (size_t c0 = 0; c0 < (size_t)m->ncell; c0 += 4096) {
      const size_t c1 = std::min((size_t)m->ncell, c0 + 4096);
      for (int k = 0; k < ncols; k++) {
         const double *src = m->pinned + (size_t)k * m->ncell;
         for (size_t c = c0; c < c1; c++) out[c * (size_t)ncols + k] = src[c];
      }
   }
   if (gather_ms) *gather_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
   return SUEWS_OK;
}

}   // extern "C"
