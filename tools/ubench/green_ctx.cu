// Probe: do green contexts (CUDA 12.4+ driver API) partition the SMs of a B200 for kernels launched with the RUNTIME API on
// green-context streams, with memory allocated in the primary context?  Prints the SM sets two concurrent kernels ran on.
//   nvcc -arch=sm_100a -o green_ctx green_ctx.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <set>
#include <vector>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("%s -> %s\n", #x, s_); return 1; } } while (0)
#define RT(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(r_)); return 1; } } while (0)

__global__ void spin(int* smid, long long* t0, long long* t1, long long cycles) {
  unsigned id;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  long long g;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
  const long long c0 = clock64();
  while (clock64() - c0 < cycles) { }
  if (threadIdx.x == 0) {
    smid[blockIdx.x] = (int)id;
    t0[blockIdx.x] = g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    t1[blockIdx.x] = g;
  }
}

int main(int argc, char** argv) {
  const int want = argc > 1 ? atoi(argv[1]) : 40;
  RT(cudaFree(0));
  CUdevice dev;
  CK(cuDeviceGet(&dev, 0));
  CUdevResource all;
  CK(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("device SMs %u\n", all.sm.smCount);
  CUdevResource part[1], rest;
  unsigned n = 1;
  CK(cuDevSmResourceSplitByCount(part, &n, &all, &rest, 0, want));
  printf("split: groups %u, group0 %u SMs, remaining %u SMs\n", n, part[0].sm.smCount, rest.sm.smCount);
  CUdevResourceDesc dA, dB;
  CK(cuDevResourceGenerateDesc(&dA, &part[0], 1));
  CK(cuDevResourceGenerateDesc(&dB, &rest, 1));
  CUgreenCtx gA, gB;
  CK(cuGreenCtxCreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CK(cuGreenCtxCreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream sA, sB;
  CK(cuGreenCtxStreamCreate(&sA, gA, CU_STREAM_NON_BLOCKING, 0));
  CK(cuGreenCtxStreamCreate(&sB, gB, CU_STREAM_NON_BLOCKING, 0));
  const int nblk = 400;
  int *smA, *smB;
  long long *tA0, *tA1, *tB0, *tB1;
  RT(cudaMalloc(&smA, nblk * 4)); RT(cudaMalloc(&smB, nblk * 4));
  RT(cudaMalloc(&tA0, nblk * 8)); RT(cudaMalloc(&tA1, nblk * 8)); RT(cudaMalloc(&tB0, nblk * 8)); RT(cudaMalloc(&tB1, nblk * 8));
  // big CTAs (one per SM): 1024 threads
  cudaEvent_t e0, e1;
  RT(cudaEventCreate(&e0)); RT(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    spin<<<nblk, 1024, 0, (cudaStream_t)sA>>>(smA, tA0, tA1, 200000);
    spin<<<nblk, 1024, 0, (cudaStream_t)sB>>>(smB, tB0, tB1, 200000);
    RT(cudaGetLastError());
    // cross-stream event from a green stream to the legacy default stream
    RT(cudaEventRecord(e0, (cudaStream_t)sA));
    RT(cudaStreamWaitEvent((cudaStream_t)sB, e0, 0));
    RT(cudaDeviceSynchronize());
  }
  std::vector<int> hA(nblk), hB(nblk);
  std::vector<long long> a0(nblk), a1(nblk), b0(nblk), b1(nblk);
  RT(cudaMemcpy(hA.data(), smA, nblk * 4, cudaMemcpyDeviceToHost)); RT(cudaMemcpy(hB.data(), smB, nblk * 4, cudaMemcpyDeviceToHost));
  RT(cudaMemcpy(a0.data(), tA0, nblk * 8, cudaMemcpyDeviceToHost)); RT(cudaMemcpy(a1.data(), tA1, nblk * 8, cudaMemcpyDeviceToHost));
  RT(cudaMemcpy(b0.data(), tB0, nblk * 8, cudaMemcpyDeviceToHost)); RT(cudaMemcpy(b1.data(), tB1, nblk * 8, cudaMemcpyDeviceToHost));
  std::set<int> SA(hA.begin(), hA.end()), SB(hB.begin(), hB.end());
  int common = 0;
  for (int x : SA) common += SB.count(x);
  long long amin = a0[0], amax = a1[0], bmin = b0[0], bmax = b1[0];
  for (int i = 0; i < nblk; ++i) { amin = std::min(amin, a0[i]); amax = std::max(amax, a1[i]); bmin = std::min(bmin, b0[i]); bmax = std::max(bmax, b1[i]); }
  printf("kernel A ran on %zu SMs, kernel B on %zu SMs, common %d\n", SA.size(), SB.size(), common);
  printf("A: [%lld, %lld] us   B: [%lld, %lld] us (relative to A start)\n", 0LL, (amax - amin) / 1000, (bmin - amin) / 1000, (bmax - amin) / 1000);
  int devrt = -1;
  RT(cudaGetDevice(&devrt));
  printf("ok (runtime device %d)\n", devrt);
  return 0;
}
