#include <cstdio>
#include <cuda_runtime.h>
// throughput of full-mask redux.max.u32: W warps per SM, 4 independent chains per warp
__global__ void k(const unsigned* in, unsigned* out, int iters) {
  const int lane = threadIdx.x & 31;
  unsigned a = in[lane], b = a * 3u, c = a * 5u, d = a * 7u;
  for (int i = 0; i < iters; i++) {
    a = __reduce_max_sync(0xffffffffu, a + lane) ^ i;
    b = __reduce_max_sync(0xffffffffu, b + lane) ^ i;
    c = __reduce_max_sync(0xffffffffu, c + lane) ^ i;
    d = __reduce_max_sync(0xffffffffu, d + lane) ^ i;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
__global__ void kshfl(const unsigned* in, unsigned* out, int iters) {
  const int lane = threadIdx.x & 31;
  unsigned a = in[lane], b = a * 3u, c = a * 5u, d = a * 7u;
  for (int i = 0; i < iters; i++) {
    a = max(a, __shfl_xor_sync(0xffffffffu, a, 4)) + i; b = max(b, __shfl_xor_sync(0xffffffffu, b, 4)) + i;
    c = max(c, __shfl_xor_sync(0xffffffffu, c, 4)) + i; d = max(d, __shfl_xor_sync(0xffffffffu, d, 4)) + i;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}
int main() {
  unsigned h[32], *d, *o;
  for (int i = 0; i < 32; i++) h[i] = (i * 2654435761u) >> 4;
  cudaMalloc(&d, 128); cudaMalloc(&o, 148 * 1024 * 4);
  cudaMemcpy(d, h, 128, cudaMemcpyHostToDevice);
  const int iters = 20000;
  for (int wps : {4, 8, 16, 32}) {
    for (int which = 0; which < 2; which++) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      if (which) kshfl<<<148, wps * 32>>>(d, o, 100); else k<<<148, wps * 32>>>(d, o, 100);
      cudaEventRecord(e0);
      if (which) kshfl<<<148, wps * 32>>>(d, o, iters); else k<<<148, wps * 32>>>(d, o, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double per_sm_per_clk = (double)wps * 4 * iters / (ms * 1e-3 * 1.965e9);
      printf("%s, %2d warps/SM x ILP4: %.3f warp-ops/clk/SM (%.2f per sub-partition)\n", which ? "shfl.bfly+max" : "redux.max.u32 ", wps, per_sm_per_clk, per_sm_per_clk / 4);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
