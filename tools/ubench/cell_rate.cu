// Micro-benchmark (B200): cycles per cell of the S->E cell arithmetic (cell.cuh: cell_fast) on register / shared-memory
// resident data, as a function of warps per SM and cells in flight per thread -- isolates the arithmetic from the copy pipeline.
#include <cstdio>
#include <cuda_runtime.h>
#include "cell.cuh"

template <int CNT, bool VAL, int MODE>
__global__ void __launch_bounds__(1024) k(const uint4* __restrict__ src, double* out, long long* cyc, int iters, double psi, double pat, const ll_coefs K) {
  __shared__ double2 tab[128];
  __shared__ uint4 cells[12 * 384 / 4];  // 18 KB: one "day" worth of packed cells per warp pattern
  if (threadIdx.x < 128) tab[threadIdx.x] = make_double2(1.0 / (1.0 + (threadIdx.x + 0.5) / 128.0), 0.01 * threadIdx.x);
  for (int i = threadIdx.x; i < 12 * 384 / 4; i += blockDim.x) cells[i] = src[i];
  __syncthreads();
  const unsigned long long magic = (unsigned long long)__double_as_longlong(K.k[13]);
  const int lane = threadIdx.x & 31;
  double val = 0.0, psig = 0.0, row[CNT], col = 0.0;
  double pm[CNT];
  for (int j = 0; j < CNT; ++j) { row[j] = 0.0; pm[j] = 1e-6 * (1 + j + lane); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    double yd[CNT], rd[CNT], X[CNT], bw[CNT], term[CNT], gge[CNT];
#pragma unroll
    for (int j = 0; j < CNT; ++j) {
      const uint4 w = cells[(lane + 32 * j + it * 7) % (12 * 384 / 4)];
      double Id;
      if (MODE == 0) {        // IMAD.WIDE magic
        Id = uint_to_double_wide(w.x & 0x00ffffffu, magic);
        rd[j] = uint_to_double_wide(w.y & 0x00ffffffu, magic);
        yd[j] = uint_to_double_wide(__byte_perm(__byte_perm(w.x, 0u, 0x4443), w.y, 0x3270), magic);
      } else if (MODE == 1) { // I2F on the conversion unit
        Id = (double)(w.x & 0x00ffffffu);
        rd[j] = (double)(w.y & 0x00ffffffu);
        yd[j] = (double)__byte_perm(__byte_perm(w.x, 0u, 0x4443), w.y, 0x3270);
      } else {                // plain magic (LOP + MOV + DADD)
        Id = uint_to_double_magic(w.x & 0x00ffffffu);
        rd[j] = uint_to_double_magic(w.y & 0x00ffffffu);
        yd[j] = uint_to_double_magic(__byte_perm(__byte_perm(w.x, 0u, 0x4443), w.y, 0x3270));
      }
      bw[j] = __hiloint2double((int)w.w, (int)w.z);
      X[j] = fma(psi, bw[j], Id);
    }
    bool all_fast = true;
#pragma unroll
    for (int j = 0; j < CNT; ++j) {
      term[j] = 0.0; gge[j] = 0.0;
      all_fast &= cell_fast<true, VAL>(yd[j], rd[j], X[j], pat * pm[j], 1e-9, tab, K, term[j], gge[j]);
    }
    if (!all_fast) val += 1.0;
#pragma unroll
    for (int j = 0; j < CNT; ++j) {
      val += term[j];
      const double h = gge[j] * X[j];
      row[j] += h;
      psig = fma(gge[j], bw[j], psig);
      col += h;
    }
  }
  long long t1 = clock64();
  double s = val + psig + col;
  for (int j = 0; j < CNT; ++j) s += row[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CNT, bool VAL, int MODE>
void run(const char* name, int threads, const uint4* src, double* out, long long* cyc) {
  const int iters = 2000;
  k<CNT, VAL, MODE><<<148, threads>>>(src, out, cyc, iters, 0.5, 0.3, LL_COEFS);
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double cells = (double)iters * CNT * threads;
  printf("%-34s threads %4d cells/iter %2d: %.2f cycles per cell per SM (%.2f per SMSP-cell)\n", name, threads, CNT, h / cells, 4.0 * h / cells);
}

int main() {
  uint4* src; double* out; long long* cyc;
  cudaMalloc(&src, 12 * 384 / 4 * sizeof(uint4)); cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  uint4* h = new uint4[12 * 384 / 4];
  for (int i = 0; i < 12 * 384 / 4; ++i) {
    const unsigned y = 20 + i % 17, I = 500 + i % 301, r = 100000 + i;
    const double bcw = 123.456 + i;
    unsigned long long b; memcpy(&b, &bcw, 8);
    h[i] = make_uint4((I & 0xffffff) | (y << 24), (r & 0xffffff) | ((y >> 8) << 24), (unsigned)b, (unsigned)(b >> 32));
  }
  cudaMemcpy(src, h, 12 * 384 / 4 * sizeof(uint4), cudaMemcpyHostToDevice);
  for (int th : {384, 768}) {
    run<4, false, 0>("grad, IMAD.WIDE conversions", th, src, out, cyc);
    run<8, false, 0>("grad, IMAD.WIDE conversions", th, src, out, cyc);
    run<4, false, 1>("grad, I2F conversions", th, src, out, cyc);
    run<8, false, 1>("grad, I2F conversions", th, src, out, cyc);
    run<4, false, 2>("grad, MOV magic conversions", th, src, out, cyc);
    run<4, true, 0>("value+grad, IMAD.WIDE", th, src, out, cyc);
    run<12, false, 0>("grad, IMAD.WIDE conversions", th, src, out, cyc);
  }
  if (cudaDeviceSynchronize() != cudaSuccess) printf("error\n");
  return 0;
}
