// Micro-benchmark (B200): per-SM feed rate of cp.async.bulk (1-D TMA) from L2 into a shared-memory ring, as a function of
// copy size, copies per stage, ring depth, and whether the region was last WRITTEN by this CTA (st.global) or by a memset.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tma.cuh"

__global__ void __launch_bounds__(416, 1) feed_kernel(unsigned char* base, size_t region, int reps, int stage_bytes, int nstage, int split, int dirty,
                                                       long long* cyc) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ uint64_t full[8], empty[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 12); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned char* src = base + (size_t)blockIdx.x * region;
  if (dirty) {  // the CTA writes its own region first (generic proxy), as the trajectory kernel's evaluation 0 does
    for (size_t o = (size_t)tid * 16; o < region; o += 416 * 16) *reinterpret_cast<uint4*>(src + o) = make_uint4(tid, 1, 2, 3);
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
  const int nchunk = (int)(region / stage_bytes);
  const int total = nchunk * reps;
  long long t0 = clock64();
  if (warp == 12) {
    if (lane == 0)
      for (int g = 0; g < total; ++g) {
        const int st = g % nstage;
        if (g >= nstage) mbar_wait(&empty[st], (unsigned)((g / nstage - 1) & 1));
        mbar_expect_tx(&full[st], stage_bytes);
        const int part = stage_bytes / split;
        for (int k = 0; k < split; ++k)
          bulk_load_1d(ring + (size_t)st * stage_bytes + (size_t)k * part, src + (size_t)(g % nchunk) * stage_bytes + (size_t)k * part, part, &full[st]);
      }
  } else {
    unsigned acc = 0;
    for (int g = 0; g < total; ++g) {
      const int st = g % nstage;
      mbar_wait(&full[st], (unsigned)((g / nstage) & 1));
      acc += ring[(size_t)st * stage_bytes + tid * 16];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (acc == 0xffffffffu) cyc[1] = acc;
    if (tid == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
  }
}

int main() {
  unsigned char* buf;
  const size_t maxbytes = (size_t)200 << 20;
  cudaMalloc(&buf, maxbytes);
  cudaMemset(buf, 1, maxbytes);
  long long* cyc;
  cudaMalloc(&cyc, 64);
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 8;
  struct cfg { int ncta, stage, nstage, split, dirty; };
  const cfg cfgs[] = {{37, 49152, 4, 1, 1}, {37, 49152, 3, 1, 1}, {37, 49152, 2, 1, 1}, {37, 65536, 3, 1, 1}, {37, 98304, 2, 1, 1}, {37, 49152, 4, 2, 1},
                      {148, 49152, 4, 1, 1}, {148, 65536, 3, 1, 1}, {148, 98304, 2, 1, 1}, {37, 16384, 8, 1, 1}, {37, 4096, 8, 1, 1}, {37, 24576, 6, 2, 1}};
  for (const cfg& c : cfgs) {
    const size_t region = (size_t)589824 / c.stage * c.stage;
    cudaMemset(buf, 2, maxbytes);
    feed_kernel<<<c.ncta, 416, (size_t)c.stage * c.nstage>>>(buf, region, reps, c.stage, c.nstage, c.split, c.dirty, cyc);
    long long h[2];
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    if (cudaGetLastError() != cudaSuccess) { printf("error\n"); return 1; }
    printf("ctas %3d stage %5d B x %d stages, %2d copies/stage, %s: %.1f B/clk/SM  (%lld cycles per pass of %zu KB)\n", c.ncta, c.stage, c.nstage, c.split,
           c.dirty ? "written by the CTA" : "memset", (double)region * reps / h[0], h[0] / reps, region / 1024);
  }
  return 0;
}
