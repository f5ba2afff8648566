// Dependent-issue latency of FP64 / shuffle / LDS instructions on one warp (B200).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  __shared__ double sm[1024];
  sm[threadIdx.x] = threadIdx.x;
  __syncthreads();
  double x = threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
  long long t1 = clock64();
  double y = x;
  for (int i = 0; i < iters; ++i) { y += __shfl_xor_sync(0xffffffffu, y, 1); y += __shfl_xor_sync(0xffffffffu, y, 2); }
  long long t2 = clock64();
  int idx = threadIdx.x;
  for (int i = 0; i < iters; ++i) { idx = (int)sm[idx & 1023] & 1023; idx = (int)sm[(idx + 1) & 1023]; }
  long long t3 = clock64();
  float f = (float)x;
  for (int i = 0; i < iters; ++i) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f)); }
  long long t4 = clock64();
  double z = y;
  for (int i = 0; i < iters; ++i) { z = z + a; z = z * b; z = z + a; z = z * b; }
  long long t5 = clock64();
  out[threadIdx.x] = x + y + idx + f + z;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
  const int iters = 10000;
  for (int th : {32, 128, 416}) {
    k<<<1, th>>>(out, cyc, iters, 1.0000001, 1e-9);
    long long h[5]; cudaMemcpy(h, cyc, 40, cudaMemcpyDeviceToHost);
    printf("threads %3d: dependent DFMA %.1f cyc, SHFL.64+DADD %.1f cyc, LDS.64(+cvt) %.1f cyc, MUFU.RCP %.1f cyc, DADD/DMUL %.1f cyc\n", th, h[0] / (4.0 * iters), h[1] / (2.0 * iters),
           h[2] / (2.0 * iters), h[3] / (2.0 * iters), h[4] / (4.0 * iters));
  }
  return 0;
}
