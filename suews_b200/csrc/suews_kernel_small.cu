// suews_kernel_small.cu -- the 256-thread / two-blocks-per-SM instantiation of the timestep kernel (suews_kernel.cuh),
// used for batches that cannot fill the chip with 1024-cell blocks.  Its own translation unit: the two geometries are
// ~1.5 minutes of ptxas each and compile in parallel.
#include "suews_kernel.cuh"

cudaError_t suews_launch_small(bool full, int grid, cudaStream_t stream, const SuRunArgs &a)
{
   static bool carveout_set = false;
   if (!carveout_set) {
      cudaFuncSetAttribute(suews_run_kernel<true, SU_BLOCK_SMALL, SU_MINBLOCKS_SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxL1);
      cudaFuncSetAttribute(suews_run_kernel<false, SU_BLOCK_SMALL, SU_MINBLOCKS_SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxL1);
      carveout_set = true;
   }
   if (full) suews_run_kernel<true, SU_BLOCK_SMALL, SU_MINBLOCKS_SMALL><<<grid, SU_BLOCK_SMALL, 0, stream>>>(a);
   else suews_run_kernel<false, SU_BLOCK_SMALL, SU_MINBLOCKS_SMALL><<<grid, SU_BLOCK_SMALL, 0, stream>>>(a);
   return cudaGetLastError();
}
