// Throughput / latency of MUFU.RSQ64H and MUFU.RCP64H (the seeds of double sqrt / reciprocal) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu64 mufu64.cu && ./mufu64
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int ILP>
__global__ void k(double* out, int iters, long long* cyc) {
  double x[ILP];
  for (int i = 0; i < ILP; i++) x[i] = 1.5 + threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      double y;
      if (MODE == 0) asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i]));
      else if (MODE == 1) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i]));
      else y = __fma_rn(x[i], 1.0000001, 0.5);
      x[i] = y + 1.0 * (MODE != 2);   // dependent chain through one DADD (MODE 2: the DFMA itself)
      if (MODE == 2) x[i] = y;
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE, int ILP>
void run(const char* name, int warps) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<MODE, ILP><<<1, 32 * warps>>>(out, iters, cyc);
  k<MODE, ILP><<<1, 32 * warps>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-10s ILP %d warps/SM %2d (per SMSP %.1f): %.1f cycles per instruction-slot per warp, %.2f cycles per warp-instruction per SMSP\n", name, ILP, warps,
         warps / 4.0, (double)h / iters / ILP, (double)h / iters / ILP / (warps / 4.0 < 1 ? 1 : warps / 4.0));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 4, 8, 16}) { run<0, 1>("rsq64h", w); run<0, 4>("rsq64h", w); }
  for (int w : {1, 4, 8, 16}) { run<1, 1>("rcp64h", w); run<1, 4>("rcp64h", w); }
  for (int w : {1, 4, 8, 16}) { run<2, 1>("dfma", w); run<2, 4>("dfma", w); }
  return 0;
}
