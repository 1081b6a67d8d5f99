// suews_kernel_inst.cu -- ONE instance of the timestep kernel (suews_kernel.cuh) and its launcher.  The Makefile compiles
// this file once per instance with -DSU_INST_K / _SNAME / _SV / _ENAME / _EV / _B / _MINB (output kind; snowUse and ESTM
// switches as launcher-name token and template value; block size and blocks per SM), each into its own object, so that the
// instances -- one to two minutes of ptxas each -- build side by side.
#include "suews_kernel.cuh"

#if !defined(SU_INST_K) || !defined(SU_INST_SV) || !defined(SU_INST_EV) || !defined(SU_INST_B) || !defined(SU_INST_MINB)
#error "compile with the SU_INST_* definitions of the Makefile"
#endif

#ifdef SU_PHASE_CLOCKS
static cudaError_t su_read_clk(unsigned long long *out16, int reset)
{
   cudaError_t e = cudaMemcpyFromSymbol(out16, su_phase_clk, sizeof(unsigned long long) * 16);
   if (e == cudaSuccess && reset) {
      unsigned long long z[16] = {0};
      e = cudaMemcpyToSymbol(su_phase_clk, z, sizeof(z));
   }
   return e;
}
#endif

cudaError_t SU_LAUNCHER_NAME(SU_INST_K, SU_INST_SNAME, SU_INST_ENAME, SU_INST_BNAME)(int nslot, cudaStream_t stream, const SuRunArgs &a)
{
   const int grid = (nslot + SU_INST_B - 1) / SU_INST_B;   // the launcher knows the block size it was compiled with
   // shared memory: 5 KB of staged forcing + the per-thread records of the hot per-surface arrays (168 B per thread,
   // 172 KB for a 1024-thread block); what is left of the SM's 256 KB serves the thread frames as L1 (per device: the
   // attributes belong to the function on the current device)
#ifdef SU_PHASE_CLOCKS
   su_register_clk_reader(su_read_clk);
#endif
   const size_t dyn = SU_HOT_BYTES(SU_INST_B);
   static unsigned long long attr_devs = 0;
   int dev = 0;
   cudaGetDevice(&dev);
   if (!((attr_devs >> (dev & 63)) & 1ull)) {
      cudaError_t e = cudaFuncSetAttribute(suews_run_kernel<SU_INST_K, SU_INST_SV, SU_INST_EV, SU_INST_B, SU_INST_MINB>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      if (e != cudaSuccess) return e;
      attr_devs |= 1ull << (dev & 63);
   }
#ifdef SU_CONST_CLASS
   cudaMemcpyToSymbolAsync(su_c_class, a.classes, sizeof(SuewsClassParams), 0, cudaMemcpyDeviceToDevice, stream);
#endif
   suews_run_kernel<SU_INST_K, SU_INST_SV, SU_INST_EV, SU_INST_B, SU_INST_MINB><<<grid, SU_INST_B, dyn, stream>>>(a);
   return cudaGetLastError();
}
