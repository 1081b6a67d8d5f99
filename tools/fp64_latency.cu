// FP64 pipe probe for the launch-geometry choice: cycles per dependent DFMA as a function of warps per SM and of the
// independent chains per thread.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void chain(double *out, int n, long long *cyc)
{
   double a[ILP];
   for (int k = 0; k < ILP; k++) a[k] = 1.0 + threadIdx.x * 1e-9 + k;
   const double m = 1.0000001, c = 1e-9;
   __syncthreads();
   long long t0 = clock64();
   for (int i = 0; i < n; i++) {
#pragma unroll
      for (int k = 0; k < ILP; k++) a[k] = fma(a[k], m, c);
   }
   long long t1 = clock64();
   double s = 0;
   for (int k = 0; k < ILP; k++) s += a[k];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
   if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
   double *out; long long *cyc, h;
   cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
   const int n = 20000;
   int threads[] = {32, 128, 256, 512, 768, 1024};
   for (int t : threads) {
      chain<1><<<148, t>>>(out, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      double c1 = (double)h / n;
      chain<2><<<148, t>>>(out, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      double c2 = (double)h / n;
      chain<4><<<148, t>>>(out, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      double c4 = (double)h / n;
      printf("threads/SM %4d (warps/scheduler %4.1f): cycles per loop trip ILP1 %.2f  ILP2 %.2f  ILP4 %.2f  -> DFMA/clk/SM %.1f %.1f %.1f\n", t, t / 128.0,
             c1, c2, c4, t / c1, 2 * t / c2, 4 * t / c4);
   }
   return 0;
}
