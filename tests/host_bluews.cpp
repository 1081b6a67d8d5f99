// Host build of the device-side BLUEWS stage (suews_b200/csrc/dev_bluews.cuh) for the CPU tests: the source the kernel
// compiles, with the device qualifiers defined away, replayed step by step on the inputs the oracle's own run handed to its
// BLUEWS (they are all on the BL output line) -- so the device code is compared with the oracle bit for bit without a GPU.
// stab_psi_heat (dev_atmos.cuh, checked against the oracle by the GPU parity tests) is supplied by the caller.
// Test infrastructure; built by the test into a temporary directory.
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#define F(c) ((double)(c##f))
#define SU_INL static inline
#define SU_HOT static inline
#define SU_DEVNI static
#define SU_COLD static
#define SU_DEV static inline
#define __expf expf
#define __fdividef(a, b) ((a) / (b))
#include "../suews_b200/csrc/suews_fields.h"
struct SuDiag { int fatal; unsigned long long warn[2]; };
static inline void su_warn(SuDiag &d, int code) { d.warn[(code >> 6) & 1] |= (1ull << (code & 63)); }
static inline void su_fatal(SuDiag &d, int code) { if (d.fatal == 0) d.fatal = code; }
static double su_pow_general(double x, double y) { return pow(x, y); }
static double (*g_psih)(int, double);
static double stab_psi_heat(int sm, double zl) { return g_psih(sm, zl); }
#include "../suews_b200/csrc/dev_latvap.cuh"
#include "../suews_b200/csrc/dev_bluews.cuh"

// cfg: EntrainmentType QH_choice isubs CO2_included InitialData_use sondeflag tstep StabilityMethod AerodynamicResistanceMethod
// clock: [nsteps][4] iy id it imin;  drive: [nsteps][10] avkdn Press_hPa avU1 qh_use qe_use UStar avdens lv_J_kg avcp zl_meas
// state8 (in/out): blh_m tm_K qm_kgkg tp_K qp_kgkg cm gamt_Km gamq_kgkgm;  lines: [nsteps][17] or NULL (policy P0)
extern "C" int bluews_host_run(int nsteps, const int *clock, const double *drive, const int *cfg, double wsb, const double *ini,
                               int nline, double (*psih)(int, double), double *state8, double *lines, unsigned long long *warn2)
{
   static SuewsClassParams P;   // (zero-initialised: only the fields BLUEWS reads are set)
   static SuewsEstmState S;
   g_psih = psih;
   P.EntrainmentType = cfg[0]; P.QH_choice = cfg[1]; P.isubs = cfg[2]; P.CO2_included = cfg[3]; P.InitialData_use = cfg[4];
   P.sondeflag = cfg[5]; P.tstep = cfg[6]; P.StabilityMethod = cfg[7]; P.AerodynamicResistanceMethod = cfg[8]; P.CBLuse = 1;
   P.wsb = wsb;
   S.cbl_blh_m = state8[0]; S.cbl_tm_K = state8[1]; S.cbl_qm_kgkg = state8[2]; S.cbl_tp_K = state8[3]; S.cbl_qp_kgkg = state8[4];
   S.cbl_cm = state8[5]; S.cbl_gamt_Km = state8[6]; S.cbl_gamq_kgkgm = state8[7];
   SuDiag dg = {0, {0, 0}};
   for (int i = 0; i < nsteps && !dg.fatal; i++) {
      SuewsStepRec R = {};
      const double *d = drive + (size_t)i * 10;
      R.iy = clock[4 * i]; R.id = clock[4 * i + 1]; R.it = clock[4 * i + 2]; R.imin = clock[4 * i + 3];
      R.avkdn = d[0]; R.Press_hPa = d[1]; R.avU1 = d[2];
      BLUEWS(dg, P, S, ini, nline, R, d[3], d[4], d[5], d[6], d[7], d[8], d[9], lines ? lines + (size_t)i * 17 : nullptr);
   }
   state8[0] = S.cbl_blh_m; state8[1] = S.cbl_tm_K; state8[2] = S.cbl_qm_kgkg; state8[3] = S.cbl_tp_K; state8[4] = S.cbl_qp_kgkg;
   state8[5] = S.cbl_cm; state8[6] = S.cbl_gamt_Km; state8[7] = S.cbl_gamq_kgkgm;
   if (warn2) { warn2[0] = dg.warn[0]; warn2[1] = dg.warn[1]; }
   return dg.fatal;
}
