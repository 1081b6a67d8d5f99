// suews_refapi.cpp -- the reference-shaped entry points of libsuews_b200.so.
//
// The reference exposes its per-cell physics as two Fortran subroutines with ~200 positional
// arguments (SUEWS_cal_Main, suews_ctrl_driver.f95:45; SUEWS_cal_multitsteps, driver.f95:2814) plus the
// small f2py helpers output_name_n / output_size (driver.f95:2778-2812).  These wrappers accept exactly
// those argument lists -- every argument by reference, arrays column-major, as gfortran and f2py pass
// them -- and run the cell on the device through the batched API of suews_b200.h.  Nothing here
// computes physics; a missing or unusable device is an error (SUEWS_E_NO_DEVICE), never a CPU fallback.
//
// Hidden state: the reference keeps three SAVE'd variables across calls (switch1/switch2 of Lat_vap,
// suews_util_meteo.f95:229-231, and alb1 of NARP, suews_phys_narp.f95:331 as compiled).  They are
// process-wide there and are process-wide here (g_hidden).
#include "../../include/suews_b200.h"
#include "suews_refargs.inc"

#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <strings.h>
#include <vector>

extern "C" int suews_set_last_error(int code, const char *msg);   // suews_b200.cu

namespace {

enum { SU_ARG_INT = 0, SU_ARG_DBL = 1 };
enum { SU_ARG_IN = 0, SU_ARG_INOUT = 1, SU_ARG_OUT = 2 };

struct ArgDesc {
   const char *name;
   int type, intent, n, per_row;
};

#define X(name, type, intent, n, per_row) {#name, type, intent, n, per_row},
const ArgDesc kMainArgs[] = {SUEWS_CAL_MAIN_ARGS(X)};
const ArgDesc kMultiArgs[] = {SUEWS_CAL_MULTITSTEPS_ARGS(X)};
#undef X
const int kNMain = (int)(sizeof(kMainArgs) / sizeof(kMainArgs[0]));
const int kNMulti = (int)(sizeof(kMultiArgs) / sizeof(kMultiArgs[0]));

#define X(name, type, intent, n, per_row) #name,
const char *const kMainNames[] = {SUEWS_CAL_MAIN_ARGS(X)};
const char *const kMultiNames[] = {SUEWS_CAL_MULTITSTEPS_ARGS(X)};
#undef X

struct VarRow {
   const char *header, *unit, *aggreg, *group;
   int level;
   const char *fmt;   // 'i0004,1X' / 'f10.4,1X' ... (suews_ctrl_output.f95:39-46)
};
#define V(h, u, a, g, l, f) {h, u, a, g, l, f},
const VarRow kVarList[] = {
#include "suews_varlist.inc"
};
#undef V
const int kNVar = (int)(sizeof(kVarList) / sizeof(kVarList[0]));
const int kVarListAllSize = 500;   // TYPE(varAttr) :: varListAll(500), suews_ctrl_output.f95:68

int find_arg(const ArgDesc *tab, int n, const char *name)
{
   for (int i = 0; i < n; i++)
      if (strcasecmp(tab[i].name, name) == 0) return i;
   return -1;
}

// one resident single-cell batch, reused by every reference-shaped call of the process
std::mutex g_mu;
suews_b200_handle g_h = nullptr;
struct {
   double alb1;
   double lv_switch;
} g_hidden = {0.0, 0.0};

// MODULE ESTM_data (suews_phys_estm.f95:1-176) is process-wide in the reference; so is this copy of it, keyed by
// lower-cased variable name.  Parameters are set by the host (suews_estm_data_set) where the reference runs
// ESTM_initials / SUEWS_Translate; the element temperatures come back here after every call.
std::map<std::string, std::vector<double>> g_estm;
// MetForcingData_grid for AnOHM (suews_anohm_set_met_grid): [g_an_rows x 24] column-major, empty = none
std::vector<double> g_an_met;
int g_an_rows = 0;
std::string lower(const char *s)
{
   std::string r(s ? s : "");
   for (auto &c : r) c = (char)tolower((unsigned char)c);
   return r;
}

int ensure_handle()
{
   if (!g_h) {
      int rc = suews_b200_create(1, 1, 0, &g_h);
      if (rc) return rc;
      suews_b200_set_tstep_prev_rolls(g_h, 0);   // tstep_prev is INTENT(IN): the same value for every step of a call
   }
   return suews_b200_restart(g_h);   // every call passes all of its arguments: nothing of the previous call is kept
}

// AnOHM's SAVE'd locals (suews_phys_anohm.f95:79, :215-216) are process-wide in the reference: they ride in the same
// name-keyed store as MODULE ESTM_data, created on the first call that selects StorageHeatMethod 3
void anohm_saves()
{
   if (g_estm.count("anohm_id_save")) return;
   g_estm["anohm_id_save"] = {0.0};
   g_estm["anohm_coef_id_save"] = {0.0};
   g_estm["anohm_coef_grid_save"] = {0.0};
   g_estm["anohm_coeff_grid_day"] = std::vector<double>(21, -999.0);
}

// Push every argument that is a library field (parameters and evolving state) into the batch.
int push_fields(const ArgDesc *tab, int ntab, void *const *args)
{
   std::vector<double> tmp;
   for (int i = 0; i < ntab; i++) {
      const ArgDesc &a = tab[i];
      if (a.intent == SU_ARG_OUT || a.per_row || a.n <= 0) continue;
      int kind = -1;
      int n = suews_b200_field_info(a.name, &kind);
      if (n <= 0) continue;   // clock / grid id / AnOHM-ESTM inputs: handled by the caller or unused on this path
      if (n != a.n) return SUEWS_E_BAD_SIZE;
      if (!args[i]) return SUEWS_E_BAD_ARG;
      tmp.resize((size_t)n);
      if (a.type == SU_ARG_INT)
         for (int k = 0; k < n; k++) tmp[(size_t)k] = (double)((const int *)args[i])[k];
      else
         memcpy(tmp.data(), args[i], sizeof(double) * (size_t)n);
      int rc;
      if (kind == SUEWS_KIND_CLASS_INT || kind == SUEWS_KIND_CLASS_DBL) rc = suews_b200_set_class_param(g_h, 0, a.name, tmp.data(), n);
      else if (kind == SUEWS_KIND_CELL_DBL) rc = suews_b200_set_cell_param(g_h, a.name, tmp.data(), n);
      else rc = suews_b200_set_state(g_h, a.name, tmp.data(), n);
      if (rc) return rc;
   }
   // MODULE ESTM_data: parameters and element temperatures
   for (const auto &kv : g_estm) {
      int kind = -1;
      const int n = suews_b200_field_info(kv.first.c_str(), &kind);
      if (n <= 0 || n != (int)kv.second.size()) return SUEWS_E_BAD_SIZE;
      int rc;
      if (kind == SUEWS_KIND_CLASS_INT || kind == SUEWS_KIND_CLASS_DBL) rc = suews_b200_set_class_param(g_h, 0, kv.first.c_str(), kv.second.data(), n);
      else if (kind == SUEWS_KIND_CELL_DBL) rc = suews_b200_set_cell_param(g_h, kv.first.c_str(), kv.second.data(), n);
      else rc = suews_b200_set_state(g_h, kv.first.c_str(), kv.second.data(), n);
      if (rc) return rc;
   }
   // hidden SAVE'd state and the library's own per-cell scratch
   double z10[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
   int rc = suews_b200_set_state(g_h, "alb1_hidden", &g_hidden.alb1, 1);
   if (!rc) rc = suews_b200_set_state(g_h, "lv_switch", &g_hidden.lv_switch, 1);
   if (!rc) rc = suews_b200_set_state(g_h, "rsl_cache", z10, 10);
   if (!rc) rc = suews_b200_set_cell_param(g_h, "dTemp_C", z10, 1);
   if (!rc) rc = suews_b200_set_cell_param(g_h, "dRH", z10, 1);
   return rc;
}

// Copy the evolved state back into the caller's INTENT(INOUT) arguments.
int pull_state(const ArgDesc *tab, int ntab, void *const *args)
{
   std::vector<double> tmp;
   for (int i = 0; i < ntab; i++) {
      const ArgDesc &a = tab[i];
      if (a.intent != SU_ARG_INOUT || a.n <= 0) continue;
      int kind = -1;
      int n = suews_b200_field_info(a.name, &kind);
      if (n <= 0 || kind != SUEWS_KIND_STATE_DBL) continue;
      tmp.resize((size_t)n);
      int rc = suews_b200_get_state(g_h, a.name, tmp.data(), n);
      if (rc) return rc;
      if (a.type == SU_ARG_INT)
         for (int k = 0; k < n; k++) ((int *)args[i])[k] = (int)tmp[(size_t)k];
      else
         memcpy(args[i], tmp.data(), sizeof(double) * (size_t)n);
   }
   int rc = suews_b200_get_state(g_h, "alb1_hidden", &g_hidden.alb1, 1);
   if (!rc) rc = suews_b200_get_state(g_h, "lv_switch", &g_hidden.lv_switch, 1);
   for (auto &kv : g_estm) {   // the element temperatures ESTM left behind
      int kind = -1;
      const int n = suews_b200_field_info(kv.first.c_str(), &kind);
      if (rc || kind != SUEWS_KIND_ESTM_STATE) continue;
      rc = suews_b200_get_state(g_h, kv.first.c_str(), kv.second.data(), n);
   }
   return rc;
}

const int kBlkCols[6] = {87, 103, 32, 138, 31, 50};   // SUEWS, snow, ESTM, RSL, SOLWEIG, DailyState (suews_ctrl_const.f95:49-55)

// Run len_sim rows of a [len_sim x 24] forcing block for the single cell and fill the six output blocks.
int run_rows(const double *met, int len_sim, int dt_since_start, int tstep_prev, double *const blk[6], const double *ts5_ir,
             bool estm_block)
{
   int rc = suews_b200_set_clock(g_h, dt_since_start, tstep_prev);
   if (!rc) rc = suews_b200_upload(g_h);
   if (!rc) rc = suews_b200_set_forcing(g_h, met, len_sim);
   if (!rc && g_an_rows > 0) rc = suews_b200_set_anohm_forcing(g_h, g_an_met.data(), g_an_rows);
   if (!rc && ts5_ir) {   // Ts5mindata_ir is one row, used for every step of the block (driver.f95:3044, :3475)
      std::vector<double> ts((size_t)len_sim * 13);
      for (int c = 0; c < 13; c++)
         for (int r = 0; r < len_sim; r++) ts[(size_t)r + (size_t)len_sim * c] = ts5_ir[c];
      rc = suews_b200_set_estm_forcing(g_h, ts.data(), len_sim);
   }
   if (!rc) rc = suews_b200_advance(g_h, 0, len_sim, SUEWS_OUT_FULL);
   if (rc) return rc;
   // a physics fatal code (< 2000, the reference's ErrorHint number) stops that cell only: rows up to the
   // failing step are still returned, then the code; library errors return at once
   const int fatal_rc = suews_b200_sync(g_h);
   if (fatal_rc >= SUEWS_E_NO_DEVICE) return fatal_rc;
   static const char *const grp[6] = {"SUEWS", "snow", "ESTM", "RSL", "SOLWEIG", "DailyState"};
   for (int g = 0; g < 6; g++) {
      if (!blk[g]) continue;
      if (g != 2 || estm_block) {
         rc = suews_b200_get_block(g_h, grp[g], 0, blk[g], len_sim, kBlkCols[g]);
         if (rc) return rc;
      } else {
         // the ESTM block is only written when StorageHeatMethod = 4 (driver.f95:2503); otherwise date columns + -999
         const size_t n = (size_t)len_sim;
         for (size_t c = 0; c < (size_t)kBlkCols[g]; c++)
            for (size_t r = 0; r < n; r++)
               blk[g][r + n * c] = (c < 5 && blk[0]) ? blk[0][r + n * c] : -999.0;
      }
   }
   return fatal_rc;
}

}   // namespace

extern "C" {

int suews_cal_multitsteps_argnames(const char *const **names)
{
   if (names) *names = kMultiNames;
   return kNMulti;
}
int suews_cal_main_argnames(const char *const **names)
{
   if (names) *names = kMainNames;
   return kNMain;
}
int suews_refapi_arginfo(int which, int i, const char **name, int *is_int, int *intent, int *elements, int *per_row)
{
   const ArgDesc *tab = which ? kMultiArgs : kMainArgs;
   const int n = which ? kNMulti : kNMain;
   if (i < 0 || i >= n) return SUEWS_E_BAD_ARG;
   if (name) *name = tab[i].name;
   if (is_int) *is_int = (tab[i].type == SU_ARG_INT);
   if (intent) *intent = tab[i].intent;
   if (elements) *elements = tab[i].n;
   if (per_row) *per_row = tab[i].per_row;
   return SUEWS_OK;
}

int suews_cal_multitsteps_c(void *const *args, int nargs)
{
   if (!args || nargs != kNMulti) return SUEWS_E_BAD_ARG;
   std::lock_guard<std::mutex> lk(g_mu);
   int rc = ensure_handle();
   if (rc) return rc;
   const double *met = (const double *)args[0];
   const int *len_sim = (const int *)args[1];
   const int i_dt = find_arg(kMultiArgs, kNMulti, "dt_since_start"), i_tp = find_arg(kMultiArgs, kNMulti, "tstep_prev"),
             i_ts = find_arg(kMultiArgs, kNMulti, "tstep");
   if (!met || !len_sim || *len_sim <= 0 || !args[i_dt] || !args[i_tp] || !args[i_ts]) return SUEWS_E_BAD_ARG;
   rc = push_fields(kMultiArgs, kNMulti, args);
   if (rc) return rc;
   double *blk[6];
   for (int g = 0; g < 6; g++) blk[g] = (double *)args[kNMulti - 6 + g];
   std::vector<double> own_main;
   if (!blk[0]) {   // the SOLWEIG/ESTM date columns are copied from the main block
      own_main.resize((size_t)*len_sim * 87);
      blk[0] = own_main.data();
   }
   const int i_shm = find_arg(kMultiArgs, kNMulti, "StorageHeatMethod"), i_ts5 = find_arg(kMultiArgs, kNMulti, "Ts5mindata_ir");
   const int shm = (i_shm >= 0 && args[i_shm]) ? *(const int *)args[i_shm] : 0;
   const double *ts5 = ((shm == 4 || shm == 14) && i_ts5 >= 0) ? (const double *)args[i_ts5] : nullptr;
   if (shm == 3) anohm_saves();
   rc = shm == 3 ? push_fields(kMultiArgs, kNMulti, args) : 0;
   if (rc) return rc;
   const int run_rc = run_rows(met, *len_sim, *(const int *)args[i_dt], *(const int *)args[i_tp], blk, ts5, shm == 4);
   if (run_rc >= SUEWS_E_NO_DEVICE) return run_rc;
   rc = pull_state(kMultiArgs, kNMulti, args);
   if (rc) return rc;
   *(int *)args[i_dt] += *len_sim * *(const int *)args[i_ts];   // driver.f95:3482
   return run_rc;
}

int suews_cal_main_c(void *const *args, int nargs)
{
   if (!args || nargs != kNMain) return SUEWS_E_BAD_ARG;
   std::lock_guard<std::mutex> lk(g_mu);
   int rc = ensure_handle();
   if (rc) return rc;
   // one forcing row in the MetForcingBlock column layout (driver.f95:3177-3201)
   double met[24];
   for (int c = 0; c < 24; c++) met[c] = -999.0;
   static const struct { const char *name; int col; } kMet[] = {
      {"iy", 0}, {"id", 1}, {"it", 2}, {"imin", 3}, {"qn1_obs", 4}, {"qs_obs", 7}, {"qf_obs", 8}, {"avU1", 9}, {"avRh", 10},
      {"Temp_C", 11}, {"Press_hPa", 12}, {"Precip", 13}, {"avkdn", 14}, {"snowFrac_obs", 15}, {"ldown_obs", 16},
      {"fcld_obs", 17}, {"wu_m3", 18}, {"xsmd", 19}, {"LAI_obs", 20}};
   for (const auto &m : kMet) {
      int i = find_arg(kMainArgs, kNMain, m.name);
      if (i < 0 || !args[i]) return SUEWS_E_BAD_ARG;
      met[m.col] = (kMainArgs[i].type == SU_ARG_INT) ? (double)*(const int *)args[i] : *(const double *)args[i];
   }
   const int i_sec = find_arg(kMainArgs, kNMain, "isec"), i_dt = find_arg(kMainArgs, kNMain, "dt_since_start"),
             i_tp = find_arg(kMainArgs, kNMain, "tstep_prev");
   if (!args[i_sec] || !args[i_dt] || !args[i_tp]) return SUEWS_E_BAD_ARG;
   if (*(const int *)args[i_sec] != 0) return SUEWS_E_NOT_IMPL;   // forcing rows carry whole minutes (driver.f95:3181)
   rc = push_fields(kMainArgs, kNMain, args);
   if (rc) return rc;
   double b0[87], b1[103], b2[32], b3[138], b4[31], b5[50];
   double *blk[6] = {b0, b1, b2, b3, b4, b5};
   const int i_shm = find_arg(kMainArgs, kNMain, "StorageHeatMethod"), i_ts5 = find_arg(kMainArgs, kNMain, "Ts5mindata_ir");
   const int shm = (i_shm >= 0 && args[i_shm]) ? *(const int *)args[i_shm] : 0;
   const double *ts5 = ((shm == 4 || shm == 14) && i_ts5 >= 0) ? (const double *)args[i_ts5] : nullptr;
   if (shm == 3) {
      if (g_an_rows <= 0)
         return suews_set_last_error(1003, "StorageHeatMethod 3 (AnOHM): hand MetForcingData_grid over with suews_anohm_set_met_grid before suews_cal_main");
      anohm_saves();
      rc = push_fields(kMainArgs, kNMain, args);
      if (rc) return rc;
   }
   const int run_rc = run_rows(met, 1, *(const int *)args[i_dt], *(const int *)args[i_tp], blk, ts5, shm == 4 || shm == 14);
   if (run_rc >= SUEWS_E_NO_DEVICE) return run_rc;
   rc = pull_state(kMainArgs, kNMain, args);
   if (rc) return rc;
   double **out = (double **)(args + (kNMain - 7));   // datetimeLine, SUEWS, Snow, ESTM, RSL, SOLWEIG, DailyState
   if (out[0]) memcpy(out[0], b0, sizeof(double) * 5);
   if (out[1]) memcpy(out[1], b0 + 5, sizeof(double) * 82);
   if (out[2]) memcpy(out[2], b1 + 5, sizeof(double) * 98);
   if (out[3]) memcpy(out[3], b2 + 5, sizeof(double) * 27);   // dataOutLineESTM: -999 unless ESTM ran (driver.f95:1545)
   if (out[4]) memcpy(out[4], b3 + 5, sizeof(double) * 133);
   if (out[5]) memcpy(out[5], b4 + 5, sizeof(double) * 26);
   if (out[6]) memcpy(out[6], b5 + 5, sizeof(double) * 45);
   return run_rc;
}

// gfortran external-procedure ABI (lower-case name + underscore, everything by reference): a reference build
// that drops its own SUEWS_cal_Main / SUEWS_cal_multitsteps links against these without source changes
// beyond the CALL site.  Errors: suews_last_error().
// A non-zero return is the reference's STOP (ErrorHint fatal, suews_ctrl_error.f95:392-415): the message goes to stderr
// and the process ends, unless SUEWS_B200_NO_STOP is set, in which case the caller must check suews_last_error().
static void fortran_stop(int rc)
{
   if (rc == 0 || getenv("SUEWS_B200_NO_STOP")) return;
   char msg[256];
   int code = 0;
   suews_last_error(&code, msg, (int)sizeof(msg));
   fprintf(stderr, " ERROR! Program stopped! See problems.txt.\n %s\nSTOP Fatal error in SUEWS!\n", msg);
   exit(1);
}
void suews_cal_main_(SUEWS_CAL_MAIN_PARAMS)
{
   void *p[] = {SUEWS_CAL_MAIN_PTRS};
   fortran_stop(suews_cal_main_c(p, (int)(sizeof(p) / sizeof(p[0]))));
}
void suews_cal_multitsteps_(SUEWS_CAL_MULTITSTEPS_PARAMS)
{
   void *p[] = {SUEWS_CAL_MULTITSTEPS_PTRS};
   fortran_stop(suews_cal_multitsteps_c(p, (int)(sizeof(p) / sizeof(p[0]))));
}

int suews_output_size(int *nVar)
{
   if (nVar) *nVar = kVarListAllSize;
   return SUEWS_OK;
}

// output_name_n (driver.f95:2778-2800): 1-based i; blank beyond the filled part of varListAll and for i >= 500.
// Fortran CHARACTER(len=15) becomes a 16-byte NUL-terminated C string.
int suews_output_name_n(int i, char name[16], char group[16], char aggreg[16], int *outlevel)
{
   if (!name || !group || !aggreg || !outlevel) return SUEWS_E_BAD_ARG;
   memset(name, 0, 16);
   memset(group, 0, 16);
   memset(aggreg, 0, 16);
   *outlevel = 0;
   if (i < kVarListAllSize && i > 0 && i <= kNVar) {
      const VarRow &v = kVarList[i - 1];
      strncpy(name, v.header, 15);
      strncpy(group, v.group, 15);
      strncpy(aggreg, v.aggreg, 15);
      *outlevel = v.level;
   }
   return SUEWS_OK;
}

// ---- SUEWS_Output_Init + SUEWS_Write_txt (suews_ctrl_output.f95:737-1023): the text form of an output block.
// Columns: the five datetime variables followed by the variables of `group` in varListAll order; those with
// level <= outlevel are written.  Header fields are right-aligned in the width of the column's format and every
// field but the last is followed by one blank ('...,1X': a trailing X edit descriptor moves the position without
// writing, so lines do not end in a blank).  Reals go through
// Fortran's Fw.d edit descriptor: correctly rounded decimal, asterisks when the value does not fit, no leading
// zero dropped while the field has room (gfortran).  Returns the number of bytes the text needs (excluding the
// terminating NUL) and writes at most buflen - 1 of them; a negative value is an error.
long long suews_write_txt(const double *block, int nrows, int ncol, const char *group, int outlevel, int with_header, char *buf,
                          long long buflen)
{
   if (!block || !group || nrows < 0) return -SUEWS_E_BAD_ARG;
   std::vector<const VarRow *> cols;
   for (int i = 0; i < kNVar; i++)
      if (strcmp(kVarList[i].group, "datetime") == 0) cols.push_back(&kVarList[i]);
   for (int i = 0; i < kNVar; i++)
      if (strcasecmp(kVarList[i].group, group) == 0) cols.push_back(&kVarList[i]);
   if ((int)cols.size() <= 5) return -SUEWS_E_BAD_NAME;
   if ((int)cols.size() > ncol) cols.resize((size_t)ncol);   // varListAll may list more names than the block has columns
   long long pos = 0;
   auto put = [&](const char *t, size_t n) {
      for (size_t k = 0; k < n; k++, pos++)
         if (buf && pos < buflen - 1) buf[pos] = t[k];
   };
   auto width_dec = [](const char *fmt, int &w, int &d, bool &is_int) {
      is_int = (fmt[0] == 'i' || fmt[0] == 'I');
      w = atoi(fmt + 1);
      const char *dot = strchr(fmt, '.');
      d = (dot && !is_int) ? atoi(dot + 1) : 0;
   };
   char tmp[512];
   if (with_header) {
      for (size_t c = 0; c < cols.size(); c++) {
         if (cols[c]->level > outlevel) continue;
         int w, d; bool is_int;
         width_dec(cols[c]->fmt, w, d, is_int);
         const int n = snprintf(tmp, sizeof tmp, "%*s ", w, cols[c]->header);   // ADJUSTR into (aW,1x); longer names are cut
         if ((int)strlen(cols[c]->header) > w) { memcpy(tmp, cols[c]->header, (size_t)w); tmp[w] = ' '; put(tmp, (size_t)w + 1); }
         else put(tmp, (size_t)n);
      }
      if (pos > 0) pos--;   // the trailing 1X writes nothing
      put("\n", 1);
   }
   for (int r = 0; r < nrows; r++) {
      for (size_t c = 0; c < cols.size(); c++) {
         if (cols[c]->level > outlevel) continue;
         int w, d; bool is_int;
         width_dec(cols[c]->fmt, w, d, is_int);
         const double v = block[(size_t)r + (size_t)nrows * c];
         int n;
         if (is_int || c < 4) n = snprintf(tmp, sizeof tmp, "%*d", w, (int)v);   // INT(dataOutSel(i, 1:4))
         else {
            n = snprintf(tmp, sizeof tmp, "%*.*f", w, d, v);
            if (n == w + 1 && tmp[0] == '0') { memmove(tmp, tmp + 1, (size_t)n); n--; }                       // 0.xxxx -> .xxxx
            else if (n == w + 1 && tmp[0] == '-' && tmp[1] == '0') { memmove(tmp + 1, tmp + 2, (size_t)n - 1); n--; }   // -0.x -> -.x
         }
         if (n != w) { memset(tmp, '*', (size_t)w); n = w; }
         tmp[n] = ' ';
         put(tmp, (size_t)n + 1);
      }
      if (pos > 0) pos--;
      put("\n", 1);
   }
   if (buf && buflen > 0) buf[pos < buflen - 1 ? pos : buflen - 1] = 0;
   return pos;
}

// ---- NextInitial (suews_ctrl_init.f95:2022-2149): the InitialConditions namelist a following run starts from, in
// gfortran's list-directed layout: a leading blank, INTEGER in 12 columns, REAL(8) with 17 significant digits in 26
// columns -- F editing followed by five blanks for 0.1 <= |x| < 1e17 (and for 0), 1PE25.17E3 otherwise.
static int fmt_list_real(double v, char *out /* at least 128 bytes */)
{
   char num[96];
   int n;
   const double a = fabs(v);
   if (v == 0.0 || (a >= 0.1 && a < 1e17)) {
      int k = 1;   // 10^(k-1) <= a < 10^k after rounding to 17 significant digits
      if (v != 0.0) {
         char e[64];
         snprintf(e, sizeof e, "%.16e", a);   // d.dddddddddddddddde+XX : the decimal exponent after rounding
         k = atoi(strchr(e, 'e') + 1) + 1;
      }
      if (k < 0) k = 0;
      n = snprintf(num, sizeof num, "%.*f", 17 - k, v);
      return snprintf(out, 128, "%*s     ", 21, num);
   }
   char e[64];
   snprintf(e, sizeof e, "%.16E", v);
   char *ex = strchr(e, 'E');
   const int x = atoi(ex + 1);
   *ex = 0;
   n = snprintf(num, sizeof num, "%sE%c%03d", e, x < 0 ? '-' : '+', x < 0 ? -x : x);
   (void)n;
   return snprintf(out, 128, "%*s", 26, num);
}

long long suews_write_initial_conditions(const double *v54, int snow_use, char *buf, long long buflen)
{
   if (!v54) return -SUEWS_E_BAD_ARG;
   static const char *const surf[7] = {"Paved", "Bldgs", "EveTr", "DecTr", "Grass", "BSoil", "Water"};
   std::string t = " &InitialConditions\n";
   char f[128], nm[64];
   snprintf(f, sizeof f, " DaysSinceRain=%12d\n", (int)v54[0]);
   t += f;
   static const char *const head[11] = {"Temp_C0", "GDD_1_0", "GDD_2_0", "LAIinitialEveTr", "LAIinitialDecTr", "LAIinitialGrass",
                                         "AlbEveTr0", "AlbDecTr0", "AlbGrass0", "DecidCap0", "Porosity0"};
   int k = 1;
   auto line = [&](const char *name, double v) {
      fmt_list_real(v, f);
      t += " ";
      t += name;
      t += "=";
      t += f;
      t += "\n";
   };
   for (int i = 0; i < 11; i++) line(head[i], v54[k++]);
   for (int i = 0; i < 6; i++) { snprintf(nm, sizeof nm, "SoilStore%sState", surf[i]); line(nm, v54[k++]); }
   for (int i = 0; i < 7; i++) { snprintf(nm, sizeof nm, "%sState", surf[i]); line(nm, v54[k++]); }
   if (snow_use == 1) {   // only when the snow part is running (init.f95:2106)
      for (int i = 0; i < 7; i++) { snprintf(nm, sizeof nm, "SnowWater%sState", surf[i]); line(nm, v54[k++]); }
      for (int i = 0; i < 7; i++) { snprintf(nm, sizeof nm, "SnowPack%s", surf[i]); line(nm, v54[k++]); }
      for (int i = 0; i < 7; i++) { snprintf(nm, sizeof nm, "SnowFrac%s", surf[i]); line(nm, v54[k++]); }
      for (int i = 0; i < 7; i++) { snprintf(nm, sizeof nm, "SnowDens%s", surf[i]); line(nm, v54[k++]); }
      line("SnowAlb0", v54[k++]);
   }
   t += " /\n";
   if (buf && buflen > 0) {
      const size_t n = t.size() < (size_t)buflen - 1 ? t.size() : (size_t)buflen - 1;
      memcpy(buf, t.data(), n);
      buf[n] = 0;
   }
   return (long long)t.size();
}

void suews_refapi_reset_hidden(void)
{
   std::lock_guard<std::mutex> lk(g_mu);
   g_hidden.alb1 = 0.0;
   g_hidden.lv_switch = 0.0;
   g_estm.clear();
}

// MODULE ESTM_data for the reference-shaped calls: a variable by its module name (a class / cell / ESTM-state
// field of the library that is not a SUEWS_cal_Main argument)
int suews_estm_data_set(const char *name, const double *vals, int n)
{
   int kind = -1;
   const int nf = suews_b200_field_info(name, &kind);
   if (nf <= 0 || !vals) return SUEWS_E_BAD_NAME;
   if (nf != n) return SUEWS_E_BAD_SIZE;
   std::lock_guard<std::mutex> lk(g_mu);
   g_estm[lower(name)].assign(vals, vals + n);
   return SUEWS_OK;
}
int suews_anohm_set_met_grid(const double *met, int nrows)
{
   std::lock_guard<std::mutex> lk(g_mu);
   if (!met || nrows <= 0) {
      g_an_met.clear();
      g_an_rows = 0;
      return SUEWS_OK;
   }
   g_an_met.assign(met, met + (size_t)nrows * 24);
   g_an_rows = nrows;
   return SUEWS_OK;
}

int suews_estm_data_get(const char *name, double *vals, int n)
{
   std::lock_guard<std::mutex> lk(g_mu);
   auto it = g_estm.find(lower(name));
   if (it == g_estm.end() || !vals) return SUEWS_E_BAD_NAME;
   if ((int)it->second.size() != n) return SUEWS_E_BAD_SIZE;
   memcpy(vals, it->second.data(), sizeof(double) * (size_t)n);
   return SUEWS_OK;
}

}   // extern "C"
