// Microbenchmarks behind DESIGN.md's latency model: dependent-chain latency and per-SM throughput of the
// FP64 and shuffle instructions the ADMM loop is made of.   nvcc -arch=sm_100a -O3 lat.cu -o lat && ./lat
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b) {
  double x = a + threadIdx.x, y = b;
  unsigned long long k = __double_as_longlong(x);
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) {
    if (OP == 0) x = __fma_rn(x, y, y);
    if (OP == 1) x = __dadd_rn(x, y);
    if (OP == 2) x = __dmul_rn(x, y);
    if (OP == 3) x = __shfl_xor_sync(0xffffffffu, x, 1);
    if (OP == 4) { unsigned long long g = __shfl_xor_sync(0xffffffffu, k, 4); k = g > k ? g : k + 1; }
    if (OP == 5) x = (x < y) ? __dadd_rn(x, 1.0) : x;            // DSETP + predicated op
    if (OP == 6) x = x / y;
    if (OP == 7) x = sqrt(x) + y;
    if (OP == 8) { unsigned b = __ballot_sync(0xffffffffu, x < y); x = __dadd_rn(x, (double)(b & 1)); }
  }
  long long t1 = clock64();
  if (OP == 4) x = __longlong_as_double(k);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// throughput: ILP independent DFMA chains per thread, many warps
template <int ILP>
__global__ void tput(double* out, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) x[j] = a + j + threadIdx.x;
  for (int i = 0; i < N_IT; i++) {
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = __fma_rn(x[j], b, b);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8 * 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"DFMA", "DADD", "DMUL", "SHFL64", "SHFL64+umax", "DSETP+pred DADD", "DDIV", "DSQRT+DADD", "DSETP+VOTE+I2F+DADD"};
  long long h;
#define RUN(OP) chain<OP><<<1, 32>>>(out, cyc, 1.0, 1.0000001); chain<OP><<<1, 32>>>(out, cyc, 1.0, 1.0000001); \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-22s %.2f cycles/op (dependent chain, 1 warp)\n", names[OP], (double)h / N_IT);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps : {4, 8, 16, 32}) {
    tput<4><<<148, warps * 32>>>(out, 1.0, 1.0000001);
    cudaEventRecord(e0); tput<4><<<148, warps * 32>>>(out, 1.0, 1.0000001); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 148 * warps * 32 * 4.0 * N_IT;
    printf("DFMA throughput, %2d warps/SM x ILP4: %.2f TFLOP/s  (%.2f DFMA/clk/SM at 1.965 GHz)\n", warps, fl / ms / 1e9,
           fl / 2 / (ms * 1e-3) / 148 / 1.965e9);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
