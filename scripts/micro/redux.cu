// Does redux.sync work with four disjoint 8-lane member masks issued by one converged warp, and what does it cost?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const unsigned* in, unsigned* out, long long* cyc) {
  const int lane = threadIdx.x & 31;
  const unsigned tmask = 0xffu << (lane & ~7);
  unsigned v = in[threadIdx.x];
  unsigned r = __reduce_max_sync(tmask, v);
  out[threadIdx.x] = r;
  long long t0 = clock64();
  unsigned acc = v;
#pragma unroll 8
  for (int i = 0; i < 1024; i++) acc = __reduce_max_sync(tmask, acc + i) ^ (unsigned)lane;
  long long t1 = clock64();
  out[32 + threadIdx.x] = acc;
  // 64-bit max via two 32-bit redux
  unsigned long long x = ((unsigned long long)in[threadIdx.x] << 32) | in[(threadIdx.x * 7) & 31];
  long long t2 = clock64();
  unsigned long long a = x;
#pragma unroll 4
  for (int i = 0; i < 1024; i++) {
    unsigned hi = (unsigned)(a >> 32), lo = (unsigned)a;
    unsigned mh = __reduce_max_sync(tmask, hi);
    unsigned ml = __reduce_max_sync(tmask, hi == mh ? lo : 0u);
    a = (((unsigned long long)mh << 32) | ml) + (unsigned long long)(lane + i);
  }
  long long t3 = clock64();
  // shuffle version
  unsigned long long b = x;
#pragma unroll 4
  for (int i = 0; i < 1024; i++) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { unsigned long long g = __shfl_xor_sync(0xffffffffu, b, o); b = g > b ? g : b; }
    b += (unsigned long long)(lane + i);
  }
  long long t4 = clock64();
  out[64 + threadIdx.x] = (unsigned)(a ^ b);
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; cyc[2] = t4 - t3; }
}
int main() {
  unsigned h[32], *d, *o; long long* c; long long hc[3];
  for (int i = 0; i < 32; i++) h[i] = (i * 2654435761u) >> 4;
  cudaMalloc(&d, 128); cudaMalloc(&o, 96 * 4); cudaMalloc(&c, 24);
  cudaMemcpy(d, h, 128, cudaMemcpyHostToDevice);
  k<<<1, 32>>>(d, o, c);
  unsigned ho[96];
  cudaMemcpy(ho, o, 96 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, 24, cudaMemcpyDeviceToHost);
  int ok = 1;
  for (int t = 0; t < 4; t++) { unsigned m = 0; for (int i = 0; i < 8; i++) m = h[t * 8 + i] > m ? h[t * 8 + i] : m; for (int i = 0; i < 8; i++) ok &= (ho[t * 8 + i] == m); }
  printf("tile-masked redux.max correct: %d  (%s)\n", ok, cudaGetErrorString(cudaDeviceSynchronize()));
  printf("redux.max.u32 chain: %.1f cycles/op; u64 max via 2 redux: %.1f cycles; u64 max via 3-level shuffle: %.1f cycles\n", hc[0] / 1024.0, hc[1] / 1024.0, hc[2] / 1024.0);
}
