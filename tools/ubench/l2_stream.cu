// Micro-benchmark (B200): can per-CTA working sets that add up to most of the 126 MB L2 be re-streamed from L2?
// Each of `ncta` persistent CTAs owns a private region of `kb` KB and streams it `reps` times through a shared-memory
// ring with 1-D bulk copies (one producer thread; consumers only wait and release).  Reports aggregate bytes/s per pass:
// pass 0 comes from HBM, later passes from L2 if the working set stays resident.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../covid19uk_b200/csrc -o l2_stream l2_stream.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tma.cuh"

#define STAGES 4
#define STAGE_BYTES 32768

__global__ void __launch_bounds__(160, 1) stream_kernel(const unsigned char* base, size_t region, int reps, long long* t_pass, int npass_rec) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned char* src = base + (size_t)blockIdx.x * region;
  const int nchunk = (int)(region / STAGE_BYTES);
  const int total = nchunk * reps;
  if (warp == 4) {
    if (lane == 0)
      for (int g = 0; g < total; ++g) {
        const int st = g % STAGES;
        if (g >= STAGES) mbar_wait(&empty[st], (unsigned)((g / STAGES - 1) & 1));
        mbar_expect_tx(&full[st], STAGE_BYTES);
        bulk_load_1d(ring + (size_t)st * STAGE_BYTES, src + (size_t)(g % nchunk) * STAGE_BYTES, STAGE_BYTES, &full[st]);
      }
  } else {
    unsigned acc = 0;
    for (int g = 0; g < total; ++g) {
      const int st = g % STAGES;
      mbar_wait(&full[st], (unsigned)((g / STAGES) & 1));
      acc += ring[(size_t)st * STAGE_BYTES + tid * 16];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
      if (tid == 0 && (g % nchunk) == nchunk - 1 && g / nchunk < npass_rec) t_pass[(size_t)blockIdx.x * npass_rec + g / nchunk] = clock64();
    }
    if (acc == 0xffffffffu) t_pass[0] = acc;
  }
}

int main(int argc, char** argv) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int reps = 17;
  unsigned char* buf;
  const size_t maxbytes = (size_t)400 << 20;
  cudaMalloc(&buf, maxbytes);
  cudaMemset(buf, 1, maxbytes);
  long long* tp;
  cudaMalloc(&tp, sizeof(long long) * 4096 * reps);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * STAGE_BYTES);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int kbs[] = {160, 320, 448, 640, 768, 896, 1280, 2048};
  const int nctas[] = {sms, sms / 2};
  for (int nc : nctas)
    for (int kb : kbs) {
      const size_t region = (size_t)kb * 1024 / STAGE_BYTES * STAGE_BYTES;
      if (region * nc > maxbytes) continue;
      float best = 1e9f;
      for (int it = 0; it < 3; ++it) {
        // evict: touch another 300 MB region? (memset of the whole buffer)
        cudaMemset(buf, it + 1, maxbytes);
        cudaEventRecord(e0);
        stream_kernel<<<nc, 160, STAGES * STAGE_BYTES>>>(buf, region, reps, tp, reps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double total = (double)region * nc * reps;
      // first pass ~ HBM rate: estimate steady-state rate from (reps-1) passes assuming pass 0 at 6.5 TB/s
      const double t0 = (double)region * nc / 6.5e12;
      printf("ctas %3d x %4d KB = %6.1f MB working set: %.3f ms for %d passes -> %.2f TB/s overall, ~%.2f TB/s for passes 1..%d\n", nc, kb,
             region * nc / 1048576.0, best, reps, total / (best * 1e-3) / 1e12, (total - (double)region * nc) / (best * 1e-3 - t0) / 1e12, reps - 1);
    }
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  return 0;
}
