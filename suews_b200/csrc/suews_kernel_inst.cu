// suews_kernel_inst.cu -- ONE instance of the timestep kernel (suews_kernel.cuh) and its launcher.  The Makefile compiles
// this file once per instance with -DSU_INST_K / _SNAME / _SV / _ENAME / _EV / _B / _MINB (output kind; snowUse and ESTM
// switches as launcher-name token and template value; block size and blocks per SM), each into its own object, so that the
// instances -- one to two minutes of ptxas each -- build side by side.
#include "suews_kernel.cuh"

#if !defined(SU_INST_K) || !defined(SU_INST_SV) || !defined(SU_INST_EV) || !defined(SU_INST_B) || !defined(SU_INST_MINB)
#error "compile with the SU_INST_* definitions of the Makefile"
#endif

cudaError_t SU_LAUNCHER_NAME(SU_INST_K, SU_INST_SNAME, SU_INST_ENAME, SU_INST_B)(int grid, cudaStream_t stream, const SuRunArgs &a)
{
   // the kernel stages 11 KB of forcing in shared memory and lives off L1 for its stack: ask for the largest L1 (per
   // device: the attribute belongs to the function on the current device)
   static unsigned long long carveout_devs = 0;
   int dev = 0;
   cudaGetDevice(&dev);
   if (!((carveout_devs >> (dev & 63)) & 1ull) && !getenv("SUEWS_B200_DEFAULT_CARVEOUT")) {
      cudaFuncSetAttribute(suews_run_kernel<SU_INST_K, SU_INST_SV, SU_INST_EV, SU_INST_B, SU_INST_MINB>,
                           cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
      carveout_devs |= 1ull << (dev & 63);
   }
   suews_run_kernel<SU_INST_K, SU_INST_SV, SU_INST_EV, SU_INST_B, SU_INST_MINB><<<grid, SU_INST_B, 0, stream>>>(a);
   return cudaGetLastError();
}
