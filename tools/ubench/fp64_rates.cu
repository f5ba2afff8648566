// Micro-benchmark (B200): per-SM issue rates of the instruction classes the log-likelihood kernel is made of.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rates fp64_rates.cu && ./fp64_rates
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(1024) k(double* out, int iters, double a, double b, int ia) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  int i0 = threadIdx.x + ia, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
  for (int i = 0; i < iters; ++i) {
    if (KIND == 0) {  // DFMA
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    } else if (KIND == 1) {  // I2F.F64 (+ IADD to keep it from being hoisted)
      x0 += (double)i0; x1 += (double)i1; x2 += (double)i2; x3 += (double)i3;
      i0 += ia; i1 += ia; i2 += ia; i3 += ia;
    } else if (KIND == 2) {  // magic int->double: LOP + DADD
      x0 += __hiloint2double(0x43300000, i0 ^ 0x80000000) - 4503601774854144.0;
      x1 += __hiloint2double(0x43300000, i1 ^ 0x80000000) - 4503601774854144.0;
      x2 += __hiloint2double(0x43300000, i2 ^ 0x80000000) - 4503601774854144.0;
      x3 += __hiloint2double(0x43300000, i3 ^ 0x80000000) - 4503601774854144.0;
      i0 += ia; i1 += ia; i2 += ia; i3 += ia;
    } else if (KIND == 3) {  // f32 rcp + newton: F2F, MUFU, F2F, 2 DFMA
      double r0 = (double)__frcp_rn((float)x0); r0 = fma(r0, fma(-x0, r0, 1.0), r0); x0 += r0;
      double r1 = (double)__frcp_rn((float)x1); r1 = fma(r1, fma(-x1, r1, 1.0), r1); x1 += r1;
      double r2 = (double)__frcp_rn((float)x2); r2 = fma(r2, fma(-x2, r2, 1.0), r2); x2 += r2;
      double r3 = (double)__frcp_rn((float)x3); r3 = fma(r3, fma(-x3, r3, 1.0), r3); x3 += r3;
    } else if (KIND == 4) {  // DADD
      x0 += a; x1 += a; x2 += a; x3 += a; x4 += b; x5 += b; x6 += b; x7 += b;
    } else if (KIND == 5) {  // DFMA + IMAD interleaved (dual issue?)
      x0 = fma(x0, a, b); i0 = i0 * ia + 1; x1 = fma(x1, a, b); i1 = i1 * ia + 1;
      x2 = fma(x2, a, b); i2 = i2 * ia + 1; x3 = fma(x3, a, b); i3 = i3 * ia + 1;
    } else if (KIND == 6) {  // MUFU.RCP f32 approx (rcp.approx.ftz) + conversions
      float f0 = (float)x0, f1 = (float)x1, f2 = (float)x2, f3 = (float)x3;
      asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(f0)); asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(f1));
      asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(f2)); asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(f3));
      x0 += (double)f0; x1 += (double)f1; x2 += (double)f2; x3 += (double)f3;
    } else if (KIND == 7) {  // DMUL
      x0 *= a; x1 *= a; x2 *= a; x3 *= a; x4 *= a; x5 *= a; x6 *= a; x7 *= a;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + i0 + i1 + i2 + i3;
}

template <int KIND>
void run(const char* name, double ops_per_iter, int threads) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 1024);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<sms, threads>>>(out, 100, 1.0000001, 1e-9, 3);
  cudaEventRecord(e0);
  k<KIND><<<sms, threads>>>(out, iters, 1.0000001, 1e-9, 3);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-44s threads %4d: %.2f thread-ops/clk/SM  (%.3f ms)\n", name, threads, ops_per_iter * iters * threads / cycles, ms);
  cudaFree(out);
}

int main() {
  for (int th : {128, 256, 512, 1024}) {
    run<0>("DFMA", 8, th);
    run<4>("DADD", 8, th);
    run<7>("DMUL", 8, th);
    run<1>("I2F.F64.S32 + DADD + IADD (conversions)", 4, th);
    run<2>("magic int->double (LOP+DADD) + DADD + IADD", 4, th);
    run<3>("frcp f32 + Newton (F2F,MUFU,F2F,2 DFMA,DADD)", 4, th);
    run<6>("rcp.approx + F2F x2 + DADD", 4, th);
    run<5>("DFMA + IMAD interleaved (pairs)", 4, th);
  }
  return 0;
}
