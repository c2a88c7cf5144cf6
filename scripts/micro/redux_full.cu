#include <cstdio>
#include <cuda_runtime.h>
// latency of a full-mask redux.max chain vs the 3-stage 64-bit butterfly, one warp
__global__ void k(const unsigned* in, unsigned* out, long long* cyc) {
  const int lane = threadIdx.x & 31;
  unsigned long long x = ((unsigned long long)in[lane] << 32) | in[(lane * 7) & 31];
  long long t0 = clock64();
  unsigned long long a = x;
#pragma unroll 4
  for (int i = 0; i < 1024; i++) {  // 64-bit max over the warp by two full-mask 32-bit redux
    unsigned hi = (unsigned)(a >> 32), lo = (unsigned)a;
    unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    a = (((unsigned long long)mh << 32) | ml) + (unsigned long long)(lane + i);
  }
  long long t1 = clock64();
  unsigned long long b = x;
#pragma unroll 4
  for (int i = 0; i < 1024; i++) {  // 8-lane butterfly (what the kernel does)
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { unsigned long long g = __shfl_xor_sync(0xffffffffu, b, o); b = g > b ? g : b; }
    b += (unsigned long long)(lane + i);
  }
  long long t2 = clock64();
  unsigned c = in[lane];
#pragma unroll 4
  for (int i = 0; i < 1024; i++) c = __reduce_max_sync(0xffffffffu, c + i) ^ (unsigned)lane;
  long long t3 = clock64();
  out[lane] = (unsigned)(a ^ b) + c;
  if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
}
int main() {
  unsigned h[32], *d, *o; long long* c; long long hc[3];
  for (int i = 0; i < 32; i++) h[i] = (i * 2654435761u) >> 4;
  cudaMalloc(&d, 128); cudaMalloc(&o, 128); cudaMalloc(&c, 24);
  cudaMemcpy(d, h, 128, cudaMemcpyHostToDevice);
  k<<<1, 32>>>(d, o, c); k<<<1, 32>>>(d, o, c);
  cudaMemcpy(hc, c, 24, cudaMemcpyDeviceToHost);
  printf("64-bit max via 2 full-mask redux : %.1f cycles/op (dependent chain, 1 warp)\n", hc[0] / 1024.0);
  printf("64-bit max via 3-stage butterfly : %.1f cycles/op\n", hc[1] / 1024.0);
  printf("32-bit full-mask redux.max       : %.1f cycles/op\n", hc[2] / 1024.0);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
