// Dependent-issue latency of DFMA / DADD / DMUL / 64-bit SHFL / LDS on one warp (cycles per op).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double a, double b) {
  __shared__ double sm[64];
  sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
  __syncthreads();
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a);
  }
  long long t1 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { x += a; x += b; x += a; x += b; }
  long long t2 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { x *= b; x *= a; x *= b; x *= a; }
  long long t3 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31); x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31); x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  }
  long long t4 = clock64();
  int idx = threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    idx = (int)sm[idx & 63] & 63; idx = (int)sm[idx & 63] & 63; idx = (int)sm[idx & 63] & 63; idx = (int)sm[idx & 63] & 63;
  }
  long long t5 = clock64();
  float y = (float)x;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { y = fmaf(y, 1.0001f, 0.5f); y = fmaf(y, 1.0001f, 0.5f); y = fmaf(y, 1.0001f, 0.5f); y = fmaf(y, 1.0001f, 0.5f); }
  long long t6 = clock64();
  if (threadIdx.x == 0) {
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5;
  }
  out[threadIdx.x] = x + idx + y;
}
int main() {
  double *out; long long *cyc, h[6];
  cudaMalloc(&out, 256); cudaMalloc(&cyc, 64);
  for (int r = 0; r < 2; ++r) { k<<<1, 32>>>(out, cyc, 1e-9, 1.0000001); cudaDeviceSynchronize(); }
  cudaMemcpy(h, cyc, 48, cudaMemcpyDeviceToHost);
  const char *names[6] = {"DFMA", "DADD", "DMUL", "SHFL.64", "LDS+cvt", "FFMA"};
  for (int i = 0; i < 6; ++i) printf("%-8s dependent latency: %.1f cycles/op\n", names[i], h[i] / 1024.0);
  return 0;
}
