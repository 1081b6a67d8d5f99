/*
 * suews_b200.h -- C ABI of libsuews_b200.so: the B200-native SUEWS per-grid-cell timestep physics.
 *
 * Drop-in boundary for the reference's hot path:
 *
 *   suews_cal_multitsteps_c   replaces SUEWS_Driver::SUEWS_cal_multitsteps
 *                             (SUEWS-SourceCode/suews_ctrl_driver.f95:2814-3505), what
 *                             supy_driver.suews_driver.suews_cal_multitsteps binds through f2py
 *                             (supy-driver/setup.py:29-50);
 *   suews_cal_main_c          replaces SUEWS_Driver::SUEWS_cal_Main (driver.f95:45-1045), called by
 *                             SUEWS_Calculations (suews_ctrl_calculations.f95:127) and SuMin
 *                             (suews_ctrl_sumin.f95:543);
 *   suews_cal_sunposition_c   replaces SUEWS_cal_sunposition (driver.f95:3508);
 *   suews_output_name_n / suews_output_size   replace output_name_n / output_size (driver.f95:2778-2812);
 *   suews_last_error          replaces the ErrorHint -> problems.txt -> STOP convention
 *                             (suews_ctrl_error.f95:392-415; supy-driver/nonstopf2py.py:54-60).
 *
 * plus the batched device API the reference has no counterpart for (its hosts loop cells serially,
 * suews_program.f95:606-623): many cells resident in HBM, advanced many timesteps per launch.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; all reals are double, all integers
 * 32-bit int; arrays are Fortran (column-major) ordered exactly as the reference declares them
 * (driver.f95:226-343); every function returns 0 on success or a SUEWS error code (the ErrorHint
 * number of the first fatal condition, or a SUEWS_E_* value).  Nothing here aborts the process.
 * The library needs a CUDA device: there is no CPU fallback (SUEWS_E_NO_DEVICE).
 */
#ifndef SUEWS_B200_H
#define SUEWS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct suews_b200_handle_s *suews_b200_handle;

enum {
   SUEWS_OK = 0,
   SUEWS_E_NO_DEVICE = 2001,   /* no CUDA device / CUDA runtime failure (message in suews_last_error) */
   SUEWS_E_BAD_NAME = 2002,    /* unknown parameter/state name */
   SUEWS_E_BAD_SIZE = 2003,    /* value count does not match the field */
   SUEWS_E_BAD_ARG = 2004,
   SUEWS_E_NOT_IMPL = 1003,    /* scheme outside this library (AnOHM): see suews_last_error */
   SUEWS_E_ESTM_INPUT = 1005   /* StorageHeatMethod 4/14 without ESTM state / surface-temperature forcing */
};

/* where a named field lives (suews_b200_field_info) */
enum { SUEWS_KIND_CLASS_INT = 0, SUEWS_KIND_CLASS_DBL = 1, SUEWS_KIND_CELL_DBL = 2, SUEWS_KIND_STATE_DBL = 3,
       SUEWS_KIND_ESTM_STATE = 4 /* ESTM element temperatures: set/get through suews_b200_set_state / _get_state */ };

/* output policy bits for suews_b200_advance */
enum {
   SUEWS_OUT_SUM82 = 1,   /* accumulate per-cell sums of the 82 SUEWS columns over the advanced steps */
   SUEWS_OUT_FULL = 2,    /* keep every output block of every step (reference SUEWS_cal_multitsteps semantics) */
   SUEWS_OUT_SUM7 = 4,    /* accumulate only the seven parity fluxes QN, QF, QS, QH, QE, RO, SWE (columns 5-9, 16, 71
                             of the same [ncell][82] sums array; the other columns are left untouched) */
   SUEWS_OUT_FLUX_ONLY = 16, /* kept for callers of the first release: this is now what SUEWS_OUT_SUM7 alone does */
   SUEWS_OUT_ALL_STAGES = 32, /* A launch evaluates the stages its requested outputs (and the state) depend on.  Policy P0
                             (SUEWS_OUT_SUM7 alone) requests seven flux sums: the roughness-sublayer diagnostics T2 / q2 /
                             U10 / RH2 (RSLProfile, driver.f95:943-949) feed neither those nor the state and are skipped,
                             unless a class runs EmissionsMethod >= 11, whose biogenic CO2 reads the modelled 2 m
                             temperature.  This bit makes such a launch run every stage of SUEWS_cal_Main regardless. */
   SUEWS_OUT_AGG = 8      /* output policy P1: the SUEWS block aggregated to the output frequency (default hourly) inside
                             the kernel, as SUEWS_Output_Agg does after the run (suews_ctrl_output.f95:912-957: mean /
                             sum / last value per column as varListAll says); exclusive with the running sums */
};

/* ---- batched device API -------------------------------------------------------------------- */
int suews_b200_device_count(void);
int suews_b200_create(int ncell, int nclass, int device, suews_b200_handle *out);
int suews_b200_destroy(suews_b200_handle h);

/* Parameters by reference argument name (case-insensitive; driver.f95:45-91).  Integer-typed
 * arguments are passed as doubles holding integral values.  Class fields: one class at a time.
 * Cell fields and state: vals is [ncell][n] (cell-major).  suews_b200_field_info reports where a
 * name lives: kind 0 = class int, 1 = class double, 2 = per-cell double, 3 = state, 4 = ESTM state; returns n or -1. */
int suews_b200_field_info(const char *name, int *kind);
int suews_b200_set_class_param(suews_b200_handle h, int cls, const char *name, const double *vals, int n);
int suews_b200_set_cell_param(suews_b200_handle h, const char *name, const double *vals, int n);
int suews_b200_set_cell_class(suews_b200_handle h, const int *class_of_cell);
int suews_b200_set_state(suews_b200_handle h, const char *name, const double *vals, int n);
int suews_b200_get_state(suews_b200_handle h, const char *name, double *vals, int n);
int suews_b200_set_clock(suews_b200_handle h, int dt_since_start, int tstep_prev);
int suews_b200_get_clock(suews_b200_handle h, int *dt_since_start, int *tstep_prev);

/* Push parameters + state to HBM (SoA).  Call after the setters, before advancing. */
int suews_b200_upload(suews_b200_handle h);
/* Sort the thread slots again on the CURRENT state (snow activity, canopy height) -- divergence control for long runs:
 * cells of a snow-on class that carry a snow pack are grouped apart from the bare ones.  State, running
 * sums, diagnostics, clock and the staged forcing are kept; costs one device -> host -> device round trip of the state. */
int suews_b200_resort(suews_b200_handle h);

/* Forcing shared by all cells: MetForcingBlock [nsteps x 24], column-major, columns as in
 * driver.f95:3177-3201 (pressure already in hPa).  Copies host -> device and builds the per-step
 * clock/ephemeris records on the device. */
int suews_b200_set_forcing(suews_b200_handle h, const double *met, int nsteps);

/* The same from forcing at its ORIGINAL resolution (e.g. hourly): met_orig is the [nrows x 24] block as MetRead reads
 * it from the file (suews_ctrl_input.f95:13-135), column-major; the device applies MetRead's conversions when
 * raw_units = 1 (pressure kPa -> hPa :109; xsmd = -999 when smd_method = 0 :96-102) and DisaggregateMet (:1380-1640:
 * DisaggregateDateTime, Disagg_Lin with DisaggMethod 1/2/3, DisaggP_amongN with RainDisaggMethod 100) to the model
 * timestep of the uploaded classes, whole-file semantics (first and last half period repeated).  The block then holds
 * nrows * resolution_in_s / tstep rows, bit-identical to the reference's in-memory MetForcingData;
 * suews_b200_get_forcing_block reads it back ([nsteps x 24], column-major).  This is RunControl's KdownZen = 0; for
 * KdownZen = 1 call suews_b200_apply_kdown_zen (below) on the block this call built. */
int suews_b200_set_forcing_orig(suews_b200_handle h, const double *met_orig, int nrows, int resolution_in_s, int disagg_method,
                                int rain_disagg_method, int raw_units, int smd_method);
int suews_b200_get_forcing_block(suews_b200_handle h, double *met_fast, int nsteps);
/* KdownZen = 1 (RunControl.nml; DisaggregateMet, suews_ctrl_input.f95:1568-1600) on the block set_forcing_orig has just
 * built: K-down is zeroed wherever the sun is below the horizon of the given site at the middle of the step before, and
 * every day's worth of rows is rescaled to keep its sum.  The reference does it per grid; a batch shares one forcing
 * block, so the site is the forcing's own. */
int suews_b200_apply_kdown_zen(suews_b200_handle h, double lat, double lng, double timezone, double alt);

/* ESTM (StorageHeatMethod 4/14) only.  The reference keeps ESTM's parameters and element temperatures in MODULE
 * ESTM_data (suews_phys_estm.f95:1-176), outside the SUEWS_cal_Main argument list.  Here the parameters are class
 * fields under the module's variable names (evolveTibld, TsurfChoice, ibldCHmod, Nibld/Nwall/Nroof/Nground,
 * LBC_soil [degC], THEAT_ON/OFF/fix [K], ivf_*, z/k/r of ibld, wall, roof and ground (5 each), nroom, alb_ibld,
 * em_ibld, CH_iroof/iwall/ibld, Tfloor [K]), the wall area is the per-cell field AreaWall, and the variables ESTM
 * carries between calls are per-cell state of kind SUEWS_KIND_ESTM_STATE (Tair2, Tievolve, lup_ground/wall/roof,
 * T0_ibld/ground/wall/roof, TN_wall/roof, Tground/Twall/Troof/Tibld (5), Tw_4 (5 x 4), HVAC, Tair2Set,
 * NewtonConverged).  The surface-temperature forcing replaces ESTMForcingData / Ts5mindata_ir
 * (suews_ctrl_translate.f95:1264-1265): ts5 is [nsteps x 13] column-major (iy id it imin Tiair Tsurf Troof Troad
 * Twall Twall_n Twall_e Twall_s Twall_w), shared by all cells, the same rows as the forcing block set before it. */
int suews_b200_set_estm_forcing(suews_b200_handle h, const double *ts5, int nsteps);
/* ... or at file resolution ([nrows x 13], e.g. hourly): DisaggregateESTM (suews_ctrl_input.f95:1650-1849) on the device,
 * DisaggMethodESTM 1 (period averages) or 2 (instantaneous values); get_estm_forcing_block reads the result back. */
int suews_b200_set_estm_forcing_orig(suews_b200_handle h, const double *ts_orig, int nrows, int resolution_in_s, int disagg_method);
int suews_b200_get_estm_forcing_block(suews_b200_handle h, double *ts_fast, int nsteps);

/* AnOHM (StorageHeatMethod 3; suews_phys_anohm.f95, called at suews_ctrl_driver.f95:1586).  Parameters: the class fields
 * cpAnOHM, kkAnOHM, chAnOHM (7 each) next to alb / emis.  Its memory -- the SAVE'd locals of AnOHM and AnOHM_coef
 * (:79, :215-216) -- is per-cell state of kind SUEWS_KIND_ESTM_STATE: anohm_id_save, anohm_coef_id_save,
 * anohm_coef_grid_save, anohm_coeff_grid_day (7 x 3, initially -999).  MetForcingData_grid, the grid's whole forcing block
 * from which the once-a-day fit takes the day's on-the-hour rows (suews_ctrl_translate.f95:1216), is the forcing block
 * staged by set_forcing / set_forcing_orig (at most 24 on-the-hour rows per day of year, i.e. up to one year per block);
 * set_anohm_forcing supplies a different one ([nrows x 24] column-major) until the next set_forcing.  Every cell behaves
 * as a single-grid run of the reference. */
int suews_b200_set_anohm_forcing(suews_b200_handle h, const double *met, int nrows);

/* BLUEWS, the convective boundary-layer slab model (suews_phys_bluews.f95; the reference's grid loop runs it after every
 * SUEWS_cal_Main, suews_ctrl_calculations.f95:194-218).  Class fields: CBLuse (RunControl.nml) and CBLInput.nml's
 * EntrainmentType, QH_choice (1 SUEWS, 2 LUMPS), isubs, CO2_included, InitialData_use (1 or 2), sondeflag (0), wsb.  The slab's
 * variables are per-cell state of kind SUEWS_KIND_ESTM_STATE (cbl_blh_m, cbl_tm_K, cbl_qm_kgkg, cbl_tp_K, cbl_qp_kgkg, cbl_cm,
 * cbl_gamt_Km, cbl_gamq_kgkgm).  ini is IniCBLdata, the rows of CBL_initial_data.txt ([nline x 8] row-major: DOY zi0 gamt
 * gamq theta+ q+ theta q), shared by the batch.  Output: the "BL" group of get_block ([nsteps x 22]: the five date columns,
 * z theta q theta+ q+ Temp_C rh QH_use QE_use Press_hPa avu1 UStar avdens lv_J_kg avcp gamt gamq), policy SUEWS_OUT_FULL. */
int suews_b200_set_cbl_initial(suews_b200_handle h, const double *ini, int nline);

/* Advance all cells over forcing rows [step0, step0 + nsteps).  One kernel launch.  Asynchronous
 * with respect to the host; suews_b200_sync waits and returns the first fatal cell error (0 if none). */
int suews_b200_advance(suews_b200_handle h, int step0, int nsteps, int out_flags);
/* A setter called after an advance keeps the run: the host copy of the state is refreshed from the device first, and the
 * running sums and per-cell diagnostics carry over the re-upload.  suews_b200_restart forgets the run instead (fresh sums
 * and diagnostics, state as set by the caller) -- what the reference-shaped calls do with their resident handle. */
int suews_b200_restart(suews_b200_handle h);
/* tstep_prev seen by the steps of a launch (SUEWS_cal_DailyState's first-step-of-day test, dailystate.f95:266): rolls = 1
 * (default) -> tstep after the first step of a run; rolls = 0 -> the set_clock value for every step, as
 * SUEWS_cal_multitsteps passes its INTENT(IN) tstep_prev to every SUEWS_cal_Main call (driver.f95:3445). */
int suews_b200_set_tstep_prev_rolls(suews_b200_handle h, int rolls);
int suews_b200_sync(suews_b200_handle h);
/* Output policy P1.  set_output_freq: ResolutionFilesOut in seconds (default 3600); every advance with SUEWS_OUT_AGG
 * aggregates its own rows [step0, step0 + nsteps) into INT(nsteps / nlinesOut) rows (a trailing partial period is
 * dropped, as in the reference).  get_agg_block: one cell's [nrows x 87] block, column-major, datetime columns =
 * the last step of each period (rows a cell did not reach because it stopped on a fatal error are NaN);
 * get_agg_column: one column (0-based, 0-4 datetime) of one row for all cells. */
int suews_b200_set_output_freq(suews_b200_handle h, int out_freq_s);
int suews_b200_get_agg_block(suews_b200_handle h, int cell, double *block, int nrows);
int suews_b200_get_agg_column(suews_b200_handle h, int row, int col, double *vals);
/* The same launch also keeps the DailyState line of every day that ends inside it (it exists on the last step of a day
 * only, dailystate.f95:1094-1111): agg_daily_rows reports how many, get_agg_daily_block returns one cell's [nrows x 50]
 * block, column-major -- the rows SUEWS_Output writes to <site>_DailyState.txt. */
int suews_b200_agg_daily_rows(suews_b200_handle h, int *nrows);
int suews_b200_get_agg_daily_block(suews_b200_handle h, int cell, double *block, int nrows);

/* Results.  sums: [ncell][82] (cell-major).  Full blocks (after SUEWS_OUT_FULL): per cell a
 * column-major [nsteps x ncol] block as SUEWS_cal_multitsteps returns it (87/103/138/50/31/32 columns). */
int suews_b200_get_sums(suews_b200_handle h, double *sums);
/* selected columns only (0-based indices into the 82-column SUEWS line), [ncell][ncols]; device -> pinned host */
int suews_b200_get_sums_cols(suews_b200_handle h, const int *cols, int ncols, double *out);
int suews_b200_device_sync(suews_b200_handle h);
/* measured FP64 FMA throughput of the device in TFLOP/s (compute-roofline denominator) */
int suews_b200_fp64_peak(int device, double *tflops);
int suews_b200_reset_sums(suews_b200_handle h);
int suews_b200_get_block(suews_b200_handle h, const char *group /* "SUEWS","snow","RSL","DailyState","SOLWEIG","ESTM" */, int cell,
                         double *block, int nsteps, int ncol);
int suews_b200_get_diag(suews_b200_handle h, int *fatal_code_per_cell, unsigned long long *warn_mask2_per_cell,
                        long long *tsurf_iterations_per_cell);
/* device time (ms) of the last advance, measured with CUDA events on the library's stream */
int suews_b200_last_advance_ms(suews_b200_handle h, float *ms);
/* raw device pointers for callers that keep results on the GPU (multi-GPU hosts): d_sums is [82][cell_stride] in cell
 * order; d_state is [field][cell_stride] in the library's thread order (the cost-sorted order the kernel walks, so that its
 * loads coalesce) -- use suews_b200_get_state for cell order */
int suews_b200_device_ptrs(suews_b200_handle h, void **d_state, void **d_sums, long long *cell_stride);

/* ---- reference-shaped entry points --------------------------------------------------------- */
/* The Fortran positional argument lists, every argument by reference (INTEGER -> int*, REAL(KIND(1d0))
 * -> double*), arrays column-major: args[k] is the k-th dummy argument of the reference subroutine,
 * inputs, INTENT(INOUT) state and INTENT(OUT) lines/blocks alike (200 for SUEWS_cal_multitsteps,
 * driver.f95:2814-2862; 220 for SUEWS_cal_Main, driver.f95:45-91).  nargs must equal that count.
 * INOUT arguments are updated in place and the output lines/blocks are filled exactly as the reference
 * fills them; the two assumed-shape AnOHM/ESTM inputs (MetForcingData_grid, Ts5mindata_ir) and any
 * unwanted output block may be NULL.  The tables (name, type, intent, extent) are generated from the
 * reference's SUBROUTINE statements into suews_b200/csrc/suews_refargs.inc; *_argnames returns the names
 * in order.  Both calls run the one cell on the device through the batched path.  Return: 0, the
 * ErrorHint number of a fatal condition (outputs up to the failing step are still returned), or SUEWS_E_*. */
int suews_cal_multitsteps_c(void *const *args, int nargs);
int suews_cal_main_c(void *const *args, int nargs);
int suews_cal_multitsteps_argnames(const char *const **names);
int suews_cal_main_argnames(const char *const **names);
/* argument i (0-based) of SUEWS_cal_Main (which = 0) or SUEWS_cal_multitsteps (which = 1): INTEGER or REAL,
 * intent (0 in, 1 inout, 2 out), element count, and whether that count is per forcing row (x len_sim) */
int suews_refapi_arginfo(int which, int i, const char **name, int *is_int, int *intent, int *elements, int *per_row);
/* The same two calls under the symbol names and calling convention gfortran gives an external procedure
 * (suews_cal_main_, suews_cal_multitsteps_: 220 / 200 pointer parameters) are exported too; their
 * prototypes are SUEWS_CAL_MAIN_PARAMS / SUEWS_CAL_MULTITSTEPS_PARAMS in suews_refargs.inc.  They return
 * nothing, like the SUBROUTINEs they replace: check suews_last_error afterwards. */
/* forget the process-wide hidden SAVE state (Lat_vap switches, NARP alb1, MODULE ESTM_data) -- a fresh reference process */
void suews_refapi_reset_hidden(void);
/* MODULE ESTM_data (suews_phys_estm.f95:1-176) for the reference-shaped calls.  The reference's SUEWS_cal_Main reads
 * ESTM's parameters and element temperatures from that module, not from its arguments; the host fills the module
 * (ESTM_initials :589, SUEWS_Translate, ESTM_translate :637) before the call.  Here the host does the same through
 * these two functions, by the module's variable names (the ESTM class fields, AreaWall and the ESTM state fields
 * listed at suews_b200_set_estm_forcing); state variables are updated by every SUEWS_cal_Main / _multitsteps call
 * with StorageHeatMethod 4/14 and can be read back.  Ts5mindata_ir is the argument of those calls. */
int suews_estm_data_set(const char *name, const double *vals, int n);
/* MetForcingData_grid for the reference-shaped calls with StorageHeatMethod 3 (AnOHM): the argument is assumed-shape in
 * the reference (driver.f95:277) and cannot travel through a pointer table, so the host hands the grid's forcing block
 * over here ([nrows x 24] column-major, copied; NULL forgets it).  Without it suews_cal_multitsteps_c uses its own
 * MetForcingBlock (the reference passes an uninitialised 10 x 10 placeholder there, driver.f95:3163) and suews_cal_main_c
 * returns 1003.  AnOHM's SAVE'd locals are process-wide, like the reference's (readable through suews_estm_data_get
 * under the names anohm_id_save, anohm_coef_id_save, anohm_coef_grid_save, anohm_coeff_grid_day). */
int suews_anohm_set_met_grid(const double *met, int nrows);
int suews_estm_data_get(const char *name, double *vals, int n);

int suews_cal_sunposition_c(double year, double idectime, double UTC, double locationlatitude,
                            double locationlongitude, double locationaltitude, double *sunazimuth,
                            double *sunzenith);
int suews_output_size(int *nVar);
int suews_output_name_n(int i, char name[16], char group[16], char aggreg[16], int *outlevel);

/* The text form of an output block, as SUEWS_Output_Init + SUEWS_Write_txt write it (suews_ctrl_output.f95:737-1023):
 * block is column-major [nrows x ncol] (5 datetime columns first; e.g. what suews_b200_get_agg_block or a
 * SUEWS_cal_multitsteps block aggregated by the caller holds), group one of "SUEWS", "snow", "ESTM", "RSL", "SOLWEIG",
 * "DailyState", outlevel the WriteOutOption level (0, 1, 2).  Writes at most buflen - 1 bytes + NUL into buf (buf may be
 * NULL to size it) and returns the full length, or a negative SUEWS_E_* value. */
long long suews_write_txt(const double *block, int nrows, int ncol, const char *group, int outlevel, int with_header, char *buf,
                          long long buflen);

/* The InitialConditions namelist NextInitial writes at the end of a year / run (suews_ctrl_init.f95:2022-2149), in
 * gfortran's list-directed layout.  v54: DaysSinceRain (HDD_id(12)), Temp_C0 (HDD_id(9)), GDD_id(1:2), LAI_id(1:3),
 * albEveTr_id, albDecTr_id, albGrass_id, DecidCap_id, porosity_id, soilstore_id(1:6), state_id(1:7), SnowWater(1:7),
 * SnowPack(1:7), SnowFrac(1:7), SnowDens(1:7), SnowAlb; the snow entries are written when snow_use = 1.
 * Returns the text length (buf may be NULL), or a negative SUEWS_E_* value. */
long long suews_write_initial_conditions(const double *v54, int snow_use, char *buf, long long buflen);

/* Diagnostic: compares the library's own exp/log (suews_b200/csrc/dev_math.cuh) with the CUDA math library
 * bit for bit on n pseudo-random arguments; returns the mismatch counts (expected 0 and 0). */
int suews_b200_selftest_math(long long n, unsigned long long seed, long long *bad_exp, long long *bad_log);
/* Diagnostic: the wet-bulb search of Lat_vap (suews_util_meteo.f95:277-346) with its single-precision screening of the
 * three-way decision against the plain loop, on n pseudo-random (Temp_C, RH, Press, cp, switch state) tuples; returns the
 * number of tuples whose latent heat, exit switches or warnings differ (expected 0). */
int suews_b200_selftest_latvap(long long n, unsigned long long seed, long long *bad);

/* last error of the calling thread: code and human-readable message */
int suews_last_error(int *code, char *buf, int len);

/* ---- several GPUs of one box behind one handle ---------------------------------------------------
 * Cells are independent (driver.f95:478-486: the inter-grid terms are zero; Test_SUEWS.py:384 checks multi-grid ==
 * single-grid), so a run is partitioned in contiguous blocks of the cell index, one block per device, with no exchange
 * while the steps advance: one host thread drives every device through its own stream (all calls below are
 * asynchronous launches), the forcing block is replicated (20 MB per year), and the only cross-device traffic is the
 * final gather of the reduced outputs -- the per-cell running sums, 8 B per cell and column -- to the first device over
 * NVLink (cudaMemcpyPeerAsync) and from there to the host.  This is what the reference's serial grid loop
 * (suews_program.f95:606-623) becomes on an 8-GPU box; a Fortran or Python host needs no process per GPU.
 * devices == NULL: devices 0 .. ndev-1.  A device may be listed more than once (two blocks on one GPU).
 * Setters take the arrays of the WHOLE run, [ncell][n] cell-major, exactly like the single-device calls. */
typedef struct suews_b200_multi_s *suews_b200_multi;
int suews_b200_multi_create(int ncell, int nclass, int ndev, const int *devices, suews_b200_multi *out);
int suews_b200_multi_destroy(suews_b200_multi m);
int suews_b200_multi_ndev(suews_b200_multi m, int *ndev);
/* block d of the partition: its single-device handle, first cell and cell count */
int suews_b200_multi_part(suews_b200_multi m, int d, suews_b200_handle *h, int *cell0, int *ncell);
int suews_b200_multi_set_class_param(suews_b200_multi m, int cls, const char *name, const double *vals, int n);
int suews_b200_multi_set_cell_param(suews_b200_multi m, const char *name, const double *vals, int n);
int suews_b200_multi_set_cell_class(suews_b200_multi m, const int *class_of_cell);
int suews_b200_multi_set_state(suews_b200_multi m, const char *name, const double *vals, int n);
int suews_b200_multi_get_state(suews_b200_multi m, const char *name, double *vals, int n);
int suews_b200_multi_set_clock(suews_b200_multi m, int dt_since_start, int tstep_prev);
int suews_b200_multi_upload(suews_b200_multi m);
int suews_b200_multi_set_forcing(suews_b200_multi m, const double *met, int nsteps);
/* the other forcing / input blocks a batch shares (same arguments as the single-device calls): a copy per device */
int suews_b200_multi_set_forcing_orig(suews_b200_multi m, const double *met_orig, int nrows, int resolution_in_s, int disagg_method,
                                      int rain_disagg_method, int raw_units, int smd_method);
int suews_b200_multi_apply_kdown_zen(suews_b200_multi m, double lat, double lng, double timezone, double alt);
int suews_b200_multi_set_estm_forcing(suews_b200_multi m, const double *ts5, int nsteps);
int suews_b200_multi_set_estm_forcing_orig(suews_b200_multi m, const double *ts_orig, int nrows, int resolution_in_s, int disagg_method);
int suews_b200_multi_set_anohm_forcing(suews_b200_multi m, const double *met, int nrows);
int suews_b200_multi_set_cbl_initial(suews_b200_multi m, const double *ini, int nline);
/* suews_b200_get_block for cell `cell` of the whole run, from the device that holds it */
int suews_b200_multi_get_block(suews_b200_multi m, const char *group, int cell, double *block, int nsteps, int ncol);
int suews_b200_multi_advance(suews_b200_multi m, int step0, int nsteps, int out_flags);   /* every device, asynchronous */
int suews_b200_multi_sync(suews_b200_multi m);
/* device time of the last advance: the slowest device's (CUDA events on each device's launch stream) */
int suews_b200_multi_last_advance_ms(suews_b200_multi m, float *ms);
/* The final gather: the selected columns of the running sums of every cell, [ncell][ncols] cell-major on the host.  Each
 * device packs its block in cell order, the blocks travel device -> first device (peer copies) and the whole array goes
 * to the host in one copy.  gather_ms (may be NULL): wall time of the call in milliseconds. */
int suews_b200_multi_gather_sums(suews_b200_multi m, const int *cols, int ncols, double *out, float *gather_ms);
int suews_b200_multi_get_diag(suews_b200_multi m, int *fatal, unsigned long long *warn2, long long *iters);

#ifdef __cplusplus
}
#endif
#endif
