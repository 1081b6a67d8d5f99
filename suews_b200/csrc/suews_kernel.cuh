// suews_kernel.cuh -- the timestep kernel of libsuews_b200.so: launch arguments, shared-memory staging helpers and
// suews_run_kernel<K, SNOW, ESTM, BLOCK, MINBLOCKS>.  Included by suews_b200.cu (host side) and by suews_kernel_inst.cu,
// which the Makefile compiles once per kernel instance (separate translation units: they compile side by side).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "suews_fields.h"
#include "dev_step.cuh"

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// One 1024-thread block per SM (64 registers): all 32 warps of an SM walk the same phase of the step
// together, which keeps instruction fetch off the critical path, and 32 warps hide the FP64 dependency
// and local-memory latencies better than 16 warps with 128 registers did (profiles/r1_ncu_notes.md:
// 512x1 1.01e8, 512x2 1.10e8, 768x2 1.23e8, 1024x1 1.29e8, 1024x2 1.14e8 cell-steps/s).
#ifndef SU_BLOCK
#define SU_BLOCK 1024
#endif
#ifndef SU_MINBLOCKS
#define SU_MINBLOCKS 1
#endif
#define SU_CHUNK 16   // timesteps staged per bulk copy (16 x 160 B = 2.5 KB per buffer)
// Shared memory of a block: the two forcing buffers (static) + one record of SUEWS_N_HOT_DBL doubles per thread (dynamic):
// the per-surface arrays every stage indexes with a run-time surface number -- surface fractions, surface and soil stores.
// In the thread's frame a run-time index makes them local-memory arrays whose every first touch of a step misses L1
// (1024 threads x 4 KB of frame); here they are one LDS away.  Record stride 21 doubles (odd): the 16 lanes of a 64-bit
// shared-memory transaction hit 16 distinct bank pairs.
// (SU_HOT_N = 21 all three arrays, 14 sfr + surface stores, 7 sfr only, 0 none; the others stay in the frame.  Measured
// on 606 208 cells x 288 steps, P0: 0 -> 3.63e8, 7 -> 3.82e8, 14 -> 3.93e8, 21 -> 3.84e8 cell-steps/s: beyond 14 the L1 that the
// shared memory takes away costs more than the third array gains)
#ifndef SU_HOT_N
#define SU_HOT_N 14
#endif
#define SU_HOT_STRIDE (SU_HOT_N | 1)
#define SU_HOT_BYTES(block) ((size_t)(block) * SU_HOT_STRIDE * sizeof(double))

struct SuRunArgs {
   const SuewsClassParams *classes;
   const int *class_of_slot;    // per-cell arrays of the device are all in thread-slot order (see upload)
   const int *tz_of_slot;
   const int *order;            // thread slot -> cell (only the cell-ordered full blocks of policy P2 need it)
   int slot0, nslot;            // the bucket of thread slots this launch covers
   int roll_tstep_prev;         // 1: tstep_prev follows the steps of the launch (batched API); 0: constant (reference-shaped call)
   const double *cellp;
   double *state;
   const SuewsStepRec *steps;   // already offset to step0
   const SuewsEphem *ephem;     // [ntz][nsteps_total], offset to step0 inside
   long long ephem_tz_stride;
   double *sums;                // [82][stride] (slot order) or NULL
   int sum7;                    // 1: accumulate only the seven parity columns (SUEWS_OUT_SUM7)
   int flux_only;               // 1: SUEWS_OUT_FLUX_ONLY (diagnostics no requested output depends on are skipped)
   int *fatal;                  // [stride] (slot order)
   unsigned long long *warn;    // [2][stride]
   long long *iters;            // [stride] (slot order)
   // full output (SUEWS_OUT_FULL): per cell column-major [nsteps x ncol]
   double *blkSUEWS, *blkSnow, *blkRSL, *blkDaily, *blkSOL, *blkESTM;
   // ESTM (StorageHeatMethod 4/14): per-cell element temperatures [SUEWS_N_ESTM_DBL][stride] and the shared
   // surface-temperature forcing [nsteps][SUEWS_N_TS5], offset to step0; NULL when no class runs ESTM
   double *estm;
   const double *ts5;
   // AnOHM (StorageHeatMethod 3): the forcing block's on-the-hour rows grouped by day of year and their index [3][367]
   // (dev_anohm.cuh); its per-cell memory is part of the `estm` record.  NULL when no class runs AnOHM
   const SuewsHourRec *an_hours;
   const int *an_day;
   // BLUEWS (CBLuse >= 1): IniCBLdata [cbl_nline][8] shared by the batch; its slab variables are part of the `estm` record
   // and its output line goes to blkBL (policy P2), [nsteps x 22] per cell
   const double *cbl_ini;
   int cbl_nline;
   double *blkBL;
   // output policy P1 (SUEWS_OUT_AGG): the SUEWS block aggregated over every agg_nlines steps as SUEWS_Output_Agg
   // does (suews_ctrl_output.f95:912-957), [row][87][stride] indexed by thread slot so that a warp writes 256 B lines
   double *agg;
   int agg_nlines;
   double *agg_daily;           // with agg: one DailyState row per day that ends inside the launch, [row][50][stride] (thread order)
   int agg_daily_rows;
   int ncell;
   long long stride;
   int nsteps;
   int dt_since_start, tstep_prev, tstep;
};

// QN, QF, QS, QH, QE, RO, SWE in the 82-column SUEWS line (0-based)
// aggregation method of the 82 SUEWS columns (varListAll, suews_ctrl_output.f95:80-166): 0 mean, 1 sum, 2 last
static __constant__ unsigned char su_agg82[SUEWS_NCOL_SUEWS] = {
#include "suews_agg82.inc"
};

__host__ __device__ __forceinline__ int su_parity_col(int i)
{
   return i < 5 ? 5 + i : (i == 5 ? 16 : 71);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
                "l"(src), "r"(bytes), "r"(smem_u32(bar))
                : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}


// A second launch geometry for batches that do not fill the chip with 1024-cell blocks: 256-thread blocks, two per SM
// (128 registers).  A 10 000-cell batch -- supy's scale -- occupies 40 SMs this way instead of 10.
#define SU_BLOCK_SMALL 256
#define SU_MINBLOCKS_SMALL 2

// Kernel variants.  A launch covers one BUCKET of cells (upload() sorts the thread slots by bucket, class and canopy
// height): cells whose classes agree on snowUse and on ESTM, so that those switches -- and the output policy -- are
// template parameters and the code of the schemes a bucket does not run is not in its instruction stream at all.
//   K     output kind: SU_K_FULL every output line of every step (reference SUEWS_cal_multitsteps semantics, policy P2),
//         SU_K_ACC the 82-column line accumulated (running sums P0+, or the in-kernel aggregation P1; also P0 for the
//         buckets without a specialised instance), SU_K_P0 the seven parity sums only (policy P0: no output line is formed)
//   SNOW  1 / 0: every / no class of the bucket has snowUse = 1; -1: decided per class at run time
//   ESTM  0: no class of the bucket keeps a record of its own (StorageHeatMethod 3 AnOHM, 4/14 ESTM, BLUEWS); -1: the record is
//         resident and the scheme is decided per class at run time
enum { SU_K_FULL = 0, SU_K_ACC = 1, SU_K_P0 = 2 };
#ifdef SU_PHASE_CLOCKS
typedef cudaError_t (*su_clk_reader_t)(unsigned long long *out16, int reset);
extern "C" void su_register_clk_reader(su_clk_reader_t f);
#endif
// the launchers are generated by suews_kernel_inst.cu, one translation unit per instance (they compile side by side)
#define SU_LAUNCHER_NAME_(k, s, e, b) suews_launch_k##k##_s##s##_e##e##_b##b
#define SU_LAUNCHER_NAME(k, s, e, b) SU_LAUNCHER_NAME_(k, s, e, b)
#define SU_DECL_LAUNCHER(k, s, e, b) cudaError_t SU_LAUNCHER_NAME(k, s, e, b)(int nslot, cudaStream_t stream, const struct SuRunArgs &a)
// (template arguments -1 are spelled m1 in the launcher names)
SU_DECL_LAUNCHER(0, m1, m1, 1024);
SU_DECL_LAUNCHER(0, m1, m1, 256);
SU_DECL_LAUNCHER(1, m1, m1, 1024);
SU_DECL_LAUNCHER(1, m1, m1, 256);
SU_DECL_LAUNCHER(1, 0, 0, 1024);
SU_DECL_LAUNCHER(2, 0, 0, 1024);
SU_DECL_LAUNCHER(2, 1, 0, 1024);
SU_DECL_LAUNCHER(2, 0, m1, 1024);   // policy P0, snow off, ESTM (StorageHeatMethod 4/14) per class at run time

#ifdef SU_CONST_CLASS   // tuning build: the (single) class record in the constant bank instead of behind a pointer
static __constant__ SuewsClassParams su_c_class;
#endif
template <int K, int SNOWK, int ESTMK, int BLOCK, int MINBLOCKS>
__global__ void __launch_bounds__(BLOCK, MINBLOCKS) suews_run_kernel(SuRunArgs a)
{
   constexpr bool FULL = (K == SU_K_FULL);
   __shared__ alignas(128) SuewsStepRec sm_steps[2][SU_CHUNK];
   __shared__ alignas(8) uint64_t sm_bar[2];
   extern __shared__ double sm_hot[];
   double hot_frame[SUEWS_N_HOT_DBL - SU_HOT_N + 1];   // (the arrays a tuning build keeps out of shared memory)
   double *const hot_sm = sm_hot + (size_t)threadIdx.x * SU_HOT_STRIDE;
#define SU_HOTP(j) (((j) < SU_HOT_N) ? (hot_sm + (j)) : (hot_frame + ((j) - SU_HOT_N)))

   const int slot = a.slot0 + blockIdx.x * blockDim.x + threadIdx.x;
   const bool active = slot < a.slot0 + a.nslot;
   const int nchunk = (a.nsteps + SU_CHUNK - 1) / SU_CHUNK;

   if (threadIdx.x == 0) {
      mbar_init(&sm_bar[0], 1);
      mbar_init(&sm_bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncthreads();
   if (threadIdx.x == 0 && nchunk > 0) {
      uint32_t n0 = (uint32_t)min(SU_CHUNK, a.nsteps) * (uint32_t)sizeof(SuewsStepRec);
      mbar_expect_tx(&sm_bar[0], n0);
      bulk_g2s(&sm_steps[0][0], a.steps, n0, &sm_bar[0]);
   }

   // ---- per-thread residents: cell parameters, state, accumulators.  Every per-cell array of the device is in
   // thread-slot order, so each of these loads and stores is one contiguous line per warp.
   SuewsCellParams C;
   SuewsCellState S;
   C.sfr.v = SU_HOTP(0);
   S.state_id.v = SU_HOTP(SUEWS_N_CELL_HOT);
   S.soilstore_id.v = SU_HOTP(SUEWS_N_CELL_HOT + 7);
   for (int j = 0; j < SUEWS_N_HOT_DBL; j++) *SU_HOTP(j) = 0.0;
   SuDiag dg;
   dg.fatal = 0; dg.warn[0] = 0; dg.warn[1] = 0;
   double acc[(K == SU_K_P0) ? 1 : SUEWS_NCOL_SUEWS];
   double acc7[7] = {0, 0, 0, 0, 0, 0, 0};
   long long iters = 0;
   int cls = 0, tz = 0;
   SuewsEstmState ES;
   const bool estm_on = (ESTMK != 0) && a.estm;
   SuAnohmIn an_in;
   an_in.hours = a.an_day ? a.an_hours : nullptr; an_in.day_first = a.an_day;
   an_in.day_count = a.an_day ? a.an_day + 367 : nullptr; an_in.day_nsun = a.an_day ? a.an_day + 734 : nullptr;
   an_in.cbl_ini = a.cbl_ini; an_in.cbl_nline = a.cbl_nline;
   if (active && estm_on) {
      double *ep = reinterpret_cast<double *>(&ES);
#pragma unroll 1
      for (int j = 0; j < SUEWS_N_ESTM_DBL; j++) ep[j] = a.estm[(size_t)j * a.stride + slot];
   }
   if (active) {
      cls = a.class_of_slot[slot];
      tz = a.tz_of_slot[slot];
      double *cp = reinterpret_cast<double *>(&C);   // COLD fields by value, then the HOT ones into the shared-memory record
#pragma unroll 1
      for (int j = 0; j < SUEWS_N_CELL_COLD; j++) cp[j] = a.cellp[(size_t)j * a.stride + slot];
      for (int j = 0; j < SUEWS_N_CELL_HOT; j++) *SU_HOTP(j) = a.cellp[(size_t)(SUEWS_N_CELL_COLD + j) * a.stride + slot];
      double *sp = reinterpret_cast<double *>(&S);
#pragma unroll 1
      for (int j = 0; j < SUEWS_N_STATE_COLD; j++) sp[j] = a.state[(size_t)j * a.stride + slot];
      for (int j = 0; j < SUEWS_N_STATE_HOT; j++) *SU_HOTP(SUEWS_N_CELL_HOT + j) = a.state[(size_t)(SUEWS_N_STATE_COLD + j) * a.stride + slot];
      dg.fatal = a.fatal[slot];
      dg.warn[0] = a.warn[slot];
      dg.warn[1] = a.warn[a.stride + slot];
      if (K != SU_K_P0 && a.agg) {
         for (int c = 0; c < SUEWS_NCOL_SUEWS; c++) acc[c] = 0.0;
      } else if (K != SU_K_P0 && a.sums && !a.sum7) {
#pragma unroll 1
         for (int c = 0; c < SUEWS_NCOL_SUEWS; c++) acc[c] = a.sums[(size_t)c * a.stride + slot];
      } else if (a.sums) {
         for (int c = 0; c < 7; c++) acc7[c] = a.sums[(size_t)su_parity_col(c) * a.stride + slot];
      }
   }
   // "last value" columns as a bit mask; only the aggregating policy overwrites instead of summing
   unsigned long long last_lo = 0, last_hi = 0;
   if (K != SU_K_P0 && a.agg)
      for (int c = 0; c < SUEWS_NCOL_SUEWS; c++)
         if (su_agg82[c] == 2) { if (c < 64) last_lo |= 1ull << c; else last_hi |= 1ull << (c - 64); }
   // the daily state lives in its own record for the length of the launch (dev_step.cuh, stage_pre_atm)
   SuDaily D;
   for (int i = 0; i < 12; i++) D.HDD_id[i] = S.HDD_id[i];
   for (int i = 0; i < 3; i++) { D.GDD_id[i] = S.GDD_id[i]; D.SDD_id[i] = S.SDD_id[i]; D.LAI_id[i] = S.LAI_id[i]; }
   for (int i = 0; i < 9; i++) D.WUDay_id[i] = S.WUDay_id[i];
   D.Tmin_id = S.Tmin_id; D.Tmax_id = S.Tmax_id; D.lenDay_id = S.lenDay_id; D.DecidCap_id = S.DecidCap_id;
   D.albDecTr_id = S.albDecTr_id; D.albEveTr_id = S.albEveTr_id; D.albGrass_id = S.albGrass_id; D.porosity_id = S.porosity_id;
   D.DecidStoreCap = SU_SDP(S.StoreDrainPrm.v, 5, DECID);
   D.deltaLAI = 0;
#ifdef SU_CONST_CLASS
   const SuewsClassParams &P = su_c_class;
#else
   const SuewsClassParams &P = a.classes[cls];
#endif
   SuSite site;
   site.spa.x = 0; site.spa.y = 0; site.spa.sin_lat = 0; site.spa.cos_lat = 1; site.veg029 = 0;
   if (active) su_site_consts(C, site);
   const SuewsEphem *eph = a.ephem + (size_t)tz * a.ephem_tz_stride;
   int dt_since_start = a.dt_since_start;
   int tstep_prev = a.tstep_prev;
   int n_day_rows = 0;
   const int cell = (FULL && active) ? a.order[slot] : 0;   // the reference-shaped blocks of policy P2 are in cell order

   for (int ch = 0; ch < nchunk; ch++) {
      const int buf = ch & 1;
      // prefetch the next chunk into the other buffer (its readers finished at the previous barrier)
      if (threadIdx.x == 0 && ch + 1 < nchunk) {
         int s1 = (ch + 1) * SU_CHUNK;
         uint32_t n1 = (uint32_t)min(SU_CHUNK, a.nsteps - s1) * (uint32_t)sizeof(SuewsStepRec);
         mbar_expect_tx(&sm_bar[buf ^ 1], n1);
         bulk_g2s(&sm_steps[buf ^ 1][0], a.steps + s1, n1, &sm_bar[buf ^ 1]);
      }
      mbar_wait(&sm_bar[buf], (uint32_t)((ch >> 1) & 1));
      const int s0 = ch * SU_CHUNK;
      const int ns = min(SU_CHUNK, a.nsteps - s0);
      for (int t = 0; t < ns; t++) {
         const int gs = s0 + t;
         const bool live = active && dg.fatal == 0;
         SuLines L;
         L.datetime5 = nullptr; L.suews82 = nullptr; L.snow98 = nullptr; L.rsl133 = nullptr; L.daily45 = nullptr; L.sol26 = nullptr;
         L.estm27 = nullptr; L.bl17 = nullptr;
         double l_dt[FULL ? 5 : 1], l_su[FULL ? SUEWS_NCOL_SUEWS : 1], l_sn[FULL ? SUEWS_NCOL_SNOW : 1], l_rs[FULL ? SUEWS_NCOL_RSL : 1],
            l_da[(K != SU_K_P0) ? SUEWS_NCOL_DAILY : 1], l_so[FULL ? SUEWS_NCOL_SOLWEIG : 1];
         double l_es[FULL ? SUEWS_NCOL_ESTM : 1], l_bl[(FULL && ESTMK != 0) ? SUEWS_NCOL_BL : 1];
         if (FULL && ESTMK != 0 && a.blkBL) L.bl17 = l_bl;
         if (FULL) { L.datetime5 = l_dt; L.suews82 = l_su; L.snow98 = l_sn; L.rsl133 = l_rs; L.daily45 = l_da; L.sol26 = l_so; L.estm27 = l_es; }
         // policy P1 also keeps the DailyState line, which only exists on the last step of a day (dailystate.f95:1094)
         const bool day_end = (K != SU_K_P0) && a.agg_daily &&
                              (sm_steps[buf][t].it * 3600 + sm_steps[buf][t].imin * 60 + a.tstep) >= 86400;
         if (day_end) L.daily45 = l_da;
         // every thread of the block enters the step (phase barriers inside); only live cells compute
         iters += suews_step<K, SNOWK, ESTMK>(live, dg, P, C, site, sm_steps[buf][t], eph[gs], dt_since_start,
                                            a.roll_tstep_prev ? tstep_prev : a.tstep_prev, S, D,
                                            (K != SU_K_P0 && (a.agg || (a.sums && !a.sum7))) ? acc : nullptr,
                                            ((K == SU_K_P0 || !a.agg) && a.sums && a.sum7) ? acc7 : nullptr, L,
                                            estm_on ? &ES : nullptr, a.ts5 ? a.ts5 + (size_t)gs * SUEWS_N_TS5 : nullptr, last_lo, last_hi,
                                            a.flux_only != 0, estm_on ? &an_in : nullptr);
         if (K != SU_K_P0) {
            if (day_end) {
               if (live && dg.fatal == 0 && n_day_rows < a.agg_daily_rows) {
                  const SuewsStepRec &R = sm_steps[buf][t];
                  double *o = a.agg_daily + (size_t)n_day_rows * (SUEWS_NCOL_DAILY + 5) * (size_t)a.stride + slot;
                  o[0] = (double)R.iy; o[(size_t)a.stride] = (double)R.id; o[2 * (size_t)a.stride] = (double)R.it;
                  o[3 * (size_t)a.stride] = (double)R.imin; o[4 * (size_t)a.stride] = R.dectime;
#pragma unroll 1
                  for (int c = 0; c < SUEWS_NCOL_DAILY; c++) o[(size_t)(5 + c) * a.stride] = l_da[c];
               }
               n_day_rows++;
            }
            if (a.agg && (gs + 1) % a.agg_nlines == 0) {   // SUEWS_Output_Agg: rows i-nlinesOut+1 .. i -> row i/nlinesOut
               if (live && dg.fatal == 0) {
                  const SuewsStepRec &R = sm_steps[buf][t];
                  double *o = a.agg + (size_t)((gs + 1) / a.agg_nlines - 1) * (SUEWS_NCOL_SUEWS + 5) * (size_t)a.stride + slot;
                  o[0] = (double)R.iy; o[(size_t)a.stride] = (double)R.id; o[2 * (size_t)a.stride] = (double)R.it;
                  o[3 * (size_t)a.stride] = (double)R.imin; o[4 * (size_t)a.stride] = R.dectime;
#pragma unroll 1
                  for (int c = 0; c < SUEWS_NCOL_SUEWS; c++)
                     o[(size_t)(5 + c) * a.stride] = (su_agg82[c] == 0) ? acc[c] / a.agg_nlines : acc[c];
               }
               for (int c = 0; c < SUEWS_NCOL_SUEWS; c++) acc[c] = 0.0;
            }
         }
         if (FULL && live && dg.fatal == 0) {
            const size_t n = (size_t)a.nsteps;
            double *b;
            b = a.blkSUEWS + (size_t)cell * n * (SUEWS_NCOL_SUEWS + 5);
            for (int c = 0; c < 5; c++) b[gs + n * c] = l_dt[c];
            for (int c = 0; c < SUEWS_NCOL_SUEWS; c++) b[gs + n * (5 + c)] = l_su[c];
            b = a.blkRSL + (size_t)cell * n * (SUEWS_NCOL_RSL + 5);
            for (int c = 0; c < 5; c++) b[gs + n * c] = l_dt[c];
            for (int c = 0; c < SUEWS_NCOL_RSL; c++) b[gs + n * (5 + c)] = l_rs[c];
            b = a.blkDaily + (size_t)cell * n * (SUEWS_NCOL_DAILY + 5);
            for (int c = 0; c < 5; c++) b[gs + n * c] = l_dt[c];
            for (int c = 0; c < SUEWS_NCOL_DAILY; c++) b[gs + n * (5 + c)] = l_da[c];
            b = a.blkSOL + (size_t)cell * n * (SUEWS_NCOL_SOLWEIG + 5);   // SUEWS_update_output with set_nan, driver.f95:2497
            for (int c = 0; c < 5; c++) b[gs + n * c] = l_dt[c];
            for (int c = 0; c < SUEWS_NCOL_SOLWEIG; c++) b[gs + n * (5 + c)] = (fabs(l_so[c]) > 30000) ? -999.0 : l_so[c];
            // the reference stores the ESTM line for StorageHeatMethod 4 only (driver.f95:2503-2505); the line itself is
            // also formed for 14 and is kept here for both (the reference-shaped block call filters on 4)
            if ((P.StorageHeatMethod == 4 || P.StorageHeatMethod == 14) && a.blkESTM) {
               b = a.blkESTM + (size_t)cell * n * (SUEWS_NCOL_ESTM + 5);
               for (int c = 0; c < 5; c++) b[gs + n * c] = l_dt[c];
               for (int c = 0; c < SUEWS_NCOL_ESTM; c++) b[gs + n * (5 + c)] = (fabs(l_es[c]) > 30000) ? -999.0 : l_es[c];
            }
            if (ESTMK != 0 && P.CBLuse >= 1 && a.blkBL) {   // dataOutBL (bluews.f95:130-133, :253-258)
               b = a.blkBL + (size_t)cell * n * (SUEWS_NCOL_BL + 5);
               for (int c = 0; c < 5; c++) b[gs + n * c] = l_dt[c];
               for (int c = 0; c < SUEWS_NCOL_BL; c++) b[gs + n * (5 + c)] = l_bl[c];
            }
            if (P.snowUse == 1) {   // SUEWS_update_output, driver.f95:2499-2501 (with set_nan :2737)
               b = a.blkSnow + (size_t)cell * n * (SUEWS_NCOL_SNOW + 5);
               for (int c = 0; c < 5; c++) b[gs + n * c] = l_dt[c];
               for (int c = 0; c < SUEWS_NCOL_SNOW; c++) b[gs + n * (5 + c)] = (fabs(l_sn[c]) > 30000) ? -999.0 : l_sn[c];
            }
         }
         dt_since_start += a.tstep;   // caller-side clock, driver.f95:3483
         tstep_prev = a.tstep;
      }
      __syncthreads();   // everyone is done with buffer `buf` before it is refilled two chunks later
   }

   if (active) {
      if (SNOWK == 0 && dg.fatal == 0 && a.nsteps > 0)   // SnowFrac_next of a snow-off step (dev_step.cuh, phase F3), once per launch
         for (int is = 0; is < NSURF; is++) S.SnowFrac[is] = 0;
      for (int i = 0; i < 12; i++) S.HDD_id[i] = D.HDD_id[i];
      for (int i = 0; i < 3; i++) { S.GDD_id[i] = D.GDD_id[i]; S.SDD_id[i] = D.SDD_id[i]; S.LAI_id[i] = D.LAI_id[i]; }
      for (int i = 0; i < 9; i++) S.WUDay_id[i] = D.WUDay_id[i];
      S.Tmin_id = D.Tmin_id; S.Tmax_id = D.Tmax_id; S.lenDay_id = D.lenDay_id; S.DecidCap_id = D.DecidCap_id;
      S.albDecTr_id = D.albDecTr_id; S.albEveTr_id = D.albEveTr_id; S.albGrass_id = D.albGrass_id; S.porosity_id = D.porosity_id;
      double *sp = reinterpret_cast<double *>(&S);
#pragma unroll 1
      for (int j = 0; j < SUEWS_N_STATE_COLD; j++) a.state[(size_t)j * a.stride + slot] = sp[j];
      for (int j = 0; j < SUEWS_N_STATE_HOT; j++) a.state[(size_t)(SUEWS_N_STATE_COLD + j) * a.stride + slot] = *SU_HOTP(SUEWS_N_CELL_HOT + j);
      if (estm_on) {
         double *ep = reinterpret_cast<double *>(&ES);
#pragma unroll 1
         for (int j = 0; j < SUEWS_N_ESTM_DBL; j++) a.estm[(size_t)j * a.stride + slot] = ep[j];
      }
      a.fatal[slot] = dg.fatal;
      a.warn[slot] = dg.warn[0];
      a.warn[a.stride + slot] = dg.warn[1];
      a.iters[slot] += iters;
      if (K != SU_K_P0 && a.agg) {
         // partial periods at the end of a block are dropped, as INT(irMax/nlinesOut) drops them
      } else if (K != SU_K_P0 && a.sums && !a.sum7) {
#pragma unroll 1
         for (int c = 0; c < SUEWS_NCOL_SUEWS; c++) a.sums[(size_t)c * a.stride + slot] = acc[c];
      } else if (a.sums) {
         for (int c = 0; c < 7; c++) a.sums[(size_t)su_parity_col(c) * a.stride + slot] = acc7[c];
      }
   }
}
