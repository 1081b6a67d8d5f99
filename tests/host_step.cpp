// Host build of the whole device-side timestep (suews_b200/csrc/dev_step.cuh and everything it includes): the source the
// kernel compiles, with the device qualifiers defined away and the CUDA math wrappers replaced by libm (-> the same
// functions the oracle calls), driven for one cell the way suews_run_kernel drives a thread under output policy FULL.
// tests/test_step_host.py compares its output blocks and end state with the oracle's BIT FOR BIT: hoisting, the deferred
// state update, the wet-bulb memo, the surface-temperature loop with its cycle detection -- every reorganisation of the
// step claims to leave the reference's arithmetic untouched, and this is where that claim is checked without a GPU.  What
// differs on the device is then the math library alone (CUDA's exp/log/pow/sin/... and the few fast paths of
// dev_common.cuh / dev_math.cuh, which are checked on the device by the selftests and the GPU parity tests).
// Test infrastructure; built by the test into a temporary directory.  Never part of the product.
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __constant__
#define SU_INLINE_ALL 1   // no out-of-line math wrappers: exp / log / pow / sin ... are libm's
#ifdef SO_OPCOUNT
// -include oracle/so_count.h (tools/host_step_opcount.py): `double` is the oracle's counting class here, so the build counts the
// FP64 operations the kernel's algorithm performs -- same weights, same reader as the oracle's own count
extern "C" { unsigned long long so_opcount_buf[32]; unsigned long long *so_opcount_cur = so_opcount_buf; int so_opcount_on = 0; }
static inline double __fma_rn(double a, double b, double c) { return a * b + c; }
#else
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
#endif
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
static inline int __double2hiint(double x) { return (int)(__double_as_longlong(x) >> 32); }
static inline int __double2loint(double x) { return (int)(__double_as_longlong(x) & 0xffffffffll); }
static inline double __hiloint2double(int hi, int lo) { return __longlong_as_double(((long long)hi << 32) | (unsigned)lo); }
static inline float __int_as_float(int x) { float r; memcpy(&r, &x, 4); return r; }
#define __expf expf
#define __fdividef(a, b) ((a) / (b))
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
// the device's special-purpose roots and reciprocals, as the plain expressions they stand for
static inline double su_root4_cr(double x) { return pow(x, 0.25); }
static inline double su_exp_fast(double x) { return exp(x); }
static inline double su_rcp_fast(double d) { return 1.0 / d; }
static inline double su_pow_general(double x, double y) { return pow(x, y); }
#ifdef SO_OPCOUNT
// (int)field on a one-value field of the counting class would need two user conversions: give that one case its own
// specialisation (same layout, same operators, plus the integer conversion)
template <typename T, int N> struct SuewsFld;
template <> struct SuewsFld<double, 1> {
   double v[1];
   operator double() const { return v[0]; }
   explicit operator int() const { return (int)v[0]; }
   SuewsFld &operator=(double x) { v[0] = x; return *this; }
};
#endif
#include "../suews_b200/csrc/dev_step.cuh"

// clock decode of one forcing row: the same statements as decode_clock / pack_step of suews_b200.cu (SUEWS_cal_dectime
// util_time.f95:270, SUEWS_cal_weekday :301; MetForcingBlock columns driver.f95:3177-3201)
static void host_pack_step(SuewsStepRec &r, const double *met, size_t n, size_t i)
{
#define MET(col) met[i + n * (col)]
   r.iy = (int)MET(0); r.id = (int)MET(1); r.it = (int)MET(2); r.imin = (int)MET(3);
   r.qn1_obs = MET(4); r.qs_obs = MET(7); r.qf_obs = MET(8); r.avU1 = MET(9); r.avRh = MET(10); r.Temp_C = MET(11);
   r.Press_hPa = MET(12); r.Precip = MET(13); r.avkdn = MET(14); r.snowFrac_obs = MET(15); r.ldown_obs = MET(16);
   r.fcld_obs = MET(17); r.wu_m3 = MET(18); r.xsmd = MET(19); r.LAI_obs = MET(20);
#undef MET
   r.dectime = (double)(r.id - 1) + (double)r.it / 24 + (double)r.imin / (60 * 24) + (double)0 / (60 * 60 * 24);
   static const int cum[13] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334, 365};
   const int year = r.iy, b = r.id;
   const int k = ((year % 4 == 0 && year % 100 != 0) || year % 400 == 0) ? 1 : 0;
   int mb = 12;
   for (int m = 1; m <= 11; m++)
      if (b <= cum[m] + (m >= 2 ? k : 0)) { mb = m; break; }
   const int md = b - (cum[mb - 1] + (mb - 1 >= 2 ? k : 0));
   int yr = year, mn = mb;
   if (!(mn > 2)) { mn += 12; yr -= 1; }
   const int n1 = (26 * (mn + 1)) / 10, n2 = (125 * yr) / 100;
   const int day = (md + n1 + n2 - (yr / 100) + (yr / 400) - 1);
   r.dow = day % 7 + 1;
   r.month = mb;
   r.pad0_ = 0; r.pad1_ = 0;
}

// K: 0 every output line of every step (policy FULL), 1 the 82-column line accumulated into acc[82], 2 the seven parity sums
// into acc[7] (policy P0; flux_only: the diagnostics no flux sum depends on are skipped, as the bench's launches do)
template <int K, int SNOWK, int ESTMK>
static int run_cell(SuewsClassParams &P, SuewsCellParams &C, SuewsCellState &S, SuewsEstmState *ES, const SuAnohmIn *AN,
                    const double *met, int nsteps, const double *ts5, int dt_since_start, int tstep_prev, double **blk,
                    SuDiag &dg, long long *iters_out, double *acc = nullptr, bool flux_only = false)
{
   SuDaily D;
   for (int i = 0; i < 12; i++) D.HDD_id[i] = S.HDD_id[i];
   for (int i = 0; i < 3; i++) { D.GDD_id[i] = S.GDD_id[i]; D.SDD_id[i] = S.SDD_id[i]; D.LAI_id[i] = S.LAI_id[i]; }
   for (int i = 0; i < 9; i++) D.WUDay_id[i] = S.WUDay_id[i];
   D.Tmin_id = S.Tmin_id; D.Tmax_id = S.Tmax_id; D.lenDay_id = S.lenDay_id; D.DecidCap_id = S.DecidCap_id;
   D.albDecTr_id = S.albDecTr_id; D.albEveTr_id = S.albEveTr_id; D.albGrass_id = S.albGrass_id; D.porosity_id = S.porosity_id;
   D.DecidStoreCap = SU_SDP(S.StoreDrainPrm.v, 5, DECID);
   D.deltaLAI = 0;
   SuSite site;
   su_site_consts(C, site);
   long long iters = 0;
   const size_t n = (size_t)nsteps;
   static const int ncol[7] = {SUEWS_NCOL_SUEWS, SUEWS_NCOL_SNOW, SUEWS_NCOL_RSL, SUEWS_NCOL_DAILY, SUEWS_NCOL_SOLWEIG, SUEWS_NCOL_ESTM,
                               SUEWS_NCOL_BL};
   for (int gs = 0; gs < nsteps; gs++) {
      if (dg.fatal) break;
      SuewsStepRec R;
      host_pack_step(R, met, n, (size_t)gs);
      SuewsEphem E;
#ifdef SO_OPCOUNT
      const int counting = so_opcount_on;   // the time-only half of the sun position is the per-time-zone pre-pass, not the cell's work
      so_opcount_on = 0;
#endif
      spa_time_part((double)R.iy, R.dectime - (double)((int)P.tstep / 2), C.timezone, E);
#ifdef SO_OPCOUNT
      so_opcount_on = counting;
#endif
      double l_dt[5], l_su[SUEWS_NCOL_SUEWS], l_sn[SUEWS_NCOL_SNOW], l_rs[SUEWS_NCOL_RSL], l_da[SUEWS_NCOL_DAILY], l_so[SUEWS_NCOL_SOLWEIG],
         l_es[SUEWS_NCOL_ESTM], l_bl[SUEWS_NCOL_BL];
      SuLines L = {l_dt, l_su, l_sn, l_rs, l_da, l_so, l_es, (ES && blk[6]) ? l_bl : nullptr};
      double ts9[SUEWS_N_TS5];
      if (ts5)
         for (int c = 0; c < SUEWS_N_TS5; c++) ts9[c] = ts5[(size_t)gs + n * (size_t)(4 + c)];
      iters += suews_step<K, SNOWK, ESTMK>(true, dg, P, C, site, R, E, dt_since_start, (gs == 0) ? tstep_prev : (int)P.tstep, S, D,
                                           (K == 1) ? acc : nullptr, (K == 2) ? acc : nullptr, L, ES, ts5 ? ts9 : nullptr, 0, 0, flux_only, AN);
      if (K == 0 && !dg.fatal) {
         double *lines[7] = {l_su, l_sn, l_rs, l_da, l_so, l_es, l_bl};
         for (int g = 0; g < 7; g++) {
            if (!blk[g]) continue;
            for (int c = 0; c < 5; c++) blk[g][(size_t)gs + n * (size_t)c] = l_dt[c];
            for (int c = 0; c < ncol[g]; c++) {
               double v = lines[g][c];
               if ((g == 1 || g == 4 || g == 5) && fabs(v) > 30000) v = -999.0;   // set_nan (driver.f95:2737) on the snow, SOLWEIG and ESTM lines
               blk[g][(size_t)gs + n * (size_t)(5 + c)] = v;
            }
         }
      }
      dt_since_start += (int)P.tstep;
   }
   if (SNOWK == 0 && dg.fatal == 0 && nsteps > 0)
      for (int is = 0; is < NSURF; is++) S.SnowFrac[is] = 0;
   for (int i = 0; i < 12; i++) S.HDD_id[i] = D.HDD_id[i];
   for (int i = 0; i < 3; i++) { S.GDD_id[i] = D.GDD_id[i]; S.SDD_id[i] = D.SDD_id[i]; S.LAI_id[i] = D.LAI_id[i]; }
   for (int i = 0; i < 9; i++) S.WUDay_id[i] = D.WUDay_id[i];
   S.Tmin_id = D.Tmin_id; S.Tmax_id = D.Tmax_id; S.lenDay_id = D.lenDay_id; S.DecidCap_id = D.DecidCap_id;
   S.albDecTr_id = D.albDecTr_id; S.albEveTr_id = D.albEveTr_id; S.albGrass_id = D.albGrass_id; S.porosity_id = D.porosity_id;
   *iters_out = iters;
   return dg.fatal;
}

// class_int / class_dbl / cell / state / estm: the values of every field of that kind, flattened in the order of the lists of
// suews_fields.h (cell and state: COLD fields, then HOT fields).  state and estm are updated in place.  blk[7]: SUEWS snow
// RSL DailyState SOLWEIG ESTM BL, each [nsteps x (5 + ncol)] column-major, or NULL.  instance: 0 generic (switches read
// from the class record, <-1,-1>), 1 the snow-off OHM instance <0,0>, 2 the snow-on OHM instance <1,0>.
// kind: 0 output blocks, 1 the 82-column line summed into acc[82], 2 the seven parity sums (QN QF QS QH QE RO SWE) into acc[7];
// acc enters holding the sums so far (zero them before the first call).
// ts5: the ESTM surface-temperature forcing [nsteps x 13] column-major or NULL; cbl_ini: IniCBLdata [cbl_nline x 8] row-major or NULL.
#include <vector>
extern "C" int host_step_run(const int *class_int, const double *class_dbl, const double *cell, double *state, double *estm,
                             const double *met, int nsteps, const double *ts5, const double *cbl_ini, int cbl_nline,
                             int dt_since_start, int tstep_prev, int instance, double **blk, unsigned long long *warn2, long long *iters,
                             int kind, int flux_only, double *acc)
{
   static SuewsClassParams P;
   SuewsCellParams C;
   SuewsCellState S;
   SuewsEstmState ES;
   double hot[SUEWS_N_HOT_DBL];
   int k = 0;
#define X(name, n) for (int i = 0; i < n; i++) P.name.v[i] = class_int[k++];
   SUEWS_CLASS_INT_FIELDS(X)
#undef X
   k = 0;
#define X(name, n) for (int i = 0; i < n; i++) P.name.v[i] = class_dbl[k++];
   SUEWS_CLASS_DBL_FIELDS(X)
#undef X
   {   // the derived class constants of suews_b200_upload
      const double *prof[6] = {P.AHProf_24hr.v, P.HumActivity_24hr.v, P.PopProf_24hr.v, P.TraffProf_24hr.v, P.WUProfA_24hr.v, P.WUProfM_24hr.v};
      for (int q = 0; q < 6; q++)
         for (int u = 0; u < 2; u++) {
            double s = 0;
            for (int i = 0; i < 24; i++) s += prof[q][24 * u + i];
            P.prof_sum[q][u] = s;
         }
      const double TH = P.TH, G5 = P.G5, TL = P.TL;
      P.gs_pow_TH_G5 = pow(TH - G5, (TH - G5) / (G5 - TL));
   }
   k = 0;
   int h = 0;
#define X(name, n) for (int i = 0; i < n; i++) C.name.v[i] = cell[k++];
   SUEWS_CELL_COLD_FIELDS(X)
#undef X
#define X(name, n) C.name.v = hot + h; for (int i = 0; i < n; i++) hot[h++] = cell[k++];
   SUEWS_CELL_HOT_FIELDS(X)
#undef X
   k = 0;
#define X(name, n) for (int i = 0; i < n; i++) S.name.v[i] = state[k++];
   SUEWS_STATE_COLD_FIELDS(X)
#undef X
#define X(name, n) S.name.v = hot + h; for (int i = 0; i < n; i++) hot[h++] = state[k++];
   SUEWS_STATE_HOT_FIELDS(X)
#undef X
   if (estm) {
      k = 0;
#define X(name, n) for (int i = 0; i < n; i++) ES.name.v[i] = estm[k++];
      SUEWS_ESTM_STATE_FIELDS(X)
#undef X
   }
   // AnOHM's hour table as build_anohm_hours of suews_b200.cu lays it out: the block's on-the-hour rows grouped by day of year
   std::vector<int> day(3 * 367, 0);
   std::vector<SuewsHourRec> hours(1);
   SuAnohmIn an = {nullptr, nullptr, nullptr, nullptr, cbl_ini, cbl_nline};
   if (estm) {
      int *first = day.data(), *count = first + 367, *nsun = count + 367;
      std::vector<SuewsStepRec> recs((size_t)nsteps);
      for (int i = 0; i < nsteps; i++) host_pack_step(recs[(size_t)i], met, (size_t)nsteps, (size_t)i);
      size_t nh = 0;
      for (const auto &r : recs)
         if (r.imin == 0 && r.id >= 1 && r.id <= 366) { count[r.id]++; if (r.avkdn > 0) nsun[r.id]++; nh++; }
      for (int d = 1; d <= 366; d++) first[d] = first[d - 1] + count[d - 1];
      hours.resize(nh ? nh : 1);
      std::vector<int> fill(367, 0);
      for (const auto &r : recs)
         if (r.imin == 0 && r.id >= 1 && r.id <= 366) {
            SuewsHourRec &q = hours[(size_t)(first[r.id] + fill[r.id]++)];
            q.id = (double)r.id; q.tHr = (double)r.it; q.Sd = r.avkdn; q.Ta = r.Temp_C; q.RH = r.avRh; q.pres = r.Press_hPa;
            q.WS = r.avU1; q.qf_obs = r.qf_obs;
         }
      an.hours = hours.data(); an.day_first = first; an.day_count = count; an.day_nsun = nsun;
   }
   SuDiag dg = {0, {0, 0}};
   long long it = 0;
   int rc;
   SuewsEstmState *es = estm ? &ES : nullptr;
   const SuAnohmIn *anp = estm ? &an : nullptr;
#define RUN(K_, SN_, ES_, es_, an_) run_cell<K_, SN_, ES_>(P, C, S, es_, an_, met, nsteps, ts5, dt_since_start, tstep_prev, blk, dg, &it, acc, flux_only != 0)
   if (kind == 0) rc = (instance == 1) ? RUN(0, 0, 0, nullptr, nullptr) : ((instance == 2) ? RUN(0, 1, 0, nullptr, nullptr) : RUN(0, -1, -1, es, anp));
   else if (kind == 1) rc = (instance == 1) ? RUN(1, 0, 0, nullptr, nullptr) : ((instance == 2) ? RUN(1, 1, 0, nullptr, nullptr) : RUN(1, -1, -1, es, anp));
   else rc = (instance == 1) ? RUN(2, 0, 0, nullptr, nullptr) : ((instance == 2) ? RUN(2, 1, 0, nullptr, nullptr) : RUN(2, -1, -1, es, anp));
#undef RUN
   k = 0;
#define X(name, n) for (int i = 0; i < n; i++) state[k++] = S.name.v[i];
   SUEWS_STATE_COLD_FIELDS(X)
   SUEWS_STATE_HOT_FIELDS(X)
#undef X
   if (estm) {
      k = 0;
#define X(name, n) for (int i = 0; i < n; i++) estm[k++] = ES.name.v[i];
      SUEWS_ESTM_STATE_FIELDS(X)
#undef X
   }
   if (warn2) { warn2[0] = dg.warn[0]; warn2[1] = dg.warn[1]; }
   if (iters) *iters = it;
   return rc;
}
