// K2 on the 5th-generation tensor cores: the commuting contraction  Bc = I . Cs  ([B*T, Mp] x [Mp, Mp], model_spec.py:262)
// as an EXACT integer GEMM on tcgen05 (kind::i8, int32 accumulators in TMEM) -- an error-free splitting ("Ozaki scheme"):
//
//   I[r][j]   = sum_a 128^a A_a[r][j]                 A_a in [0,127]    (na <= 3 digit planes: populations < 2^21)
//   Cs[j][i]  = 2^e_i sum_c 128^-(c+1) D_c[j][i] + delta   D_c in [-64,64], |delta| <= 2^e_i 128^-nb / 2  (nb = 7 planes,
//                                                          per-column scale 2^e_i chosen at model creation)
//   Bc[r][i]  = 2^e_i sum_s 128^(s-1) G_s[r][i],      G_s = sum_{a-c=s} A_a D_c   (int32, exact: |G_s| < 2^24)
//
// Every partial product is exact in int32; the only rounding is the truncation of Cs to 49 bits below its column maximum
// and the nine FP64 additions of the epilogue -- a relative error of ~1e-14 on Bc, far inside the 1e-10 budget of the
// log-probability (tests/test_gpu_contract.py compares against the FP64 DMMA kernel of contract.cu).
// 21 int8 GEMMs of 2 M^2 T flop each replace one FP64 GEMM: at 4.5 Pop/s (int8) vs 35 Tflop/s (FP64 DMMA) that is ~5x
// less tensor time for the same result.
//
// Kernel (one CTA per SM, persistent over 128 x 128 output tiles):
//   warps 0-7  (a) split the tile's 128 x K int32 slab of I into na int8 planes in shared memory, canonical K-major
//                  no-swizzle UMMA layout (8 x 16 B core matrices), fence.proxy.async, arrive
//              (b) epilogue: tcgen05.ld finished accumulator groups, FMA them into 64 FP64 registers per thread with
//                  weight 128^(s-1), scale by 2^e_i, store the tile
//   warp 8     producer: streams the pre-split planes of Cs (laid out on the host in the same canonical layout) with
//              1-D bulk copies into a 3-stage ring (half a plane per stage)
//   warp 9     allocates TMEM (512 columns = 4 accumulator slots of 128), one elected lane issues tcgen05.mma:
//              plane-major order (c outer, a inner) so that only na+1 accumulator groups are live at a time:
//              group s = a - c uses slot s mod 4; after plane c group na-1-c is complete and is committed to the epilogue.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "seir_internal.cuh"
#include "tma.cuh"

#define I8_BM 128
#define I8_BN 128
#define I8_NB 7        // digit planes of Cs
#define I8_STAGES 3
#define I8_EPI_THREADS 256
#define I8_THREADS (I8_EPI_THREADS + 64)

// ---- tcgen05 wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes); LBO = byte distance between the two
// K-adjacent core matrices of one MMA (128), SBO = distance between 8-row groups.  (cute::UMMA::SmemDescriptor bit layout.)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor: c_format S32 (2) [4,6), a/b_format INT8 (1) [7,10) / [10,13), K-major both, N>>3 [17,23), M>>4 [24,29)
#define I8_IDESC ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_BN >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24))

__global__ void __launch_bounds__(I8_THREADS, 1) seir_contract_i8_kernel(long long R, int Mp, int na, int ntiles,
                                                                         const int* __restrict__ Ix,
                                                                         const signed char* __restrict__ Bd,
                                                                         const double* __restrict__ colscale, double* __restrict__ Bc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full_b[I8_STAGES], empty_b[I8_STAGES], a_ready, a_free, t_full[4], t_empty[4];
  __shared__ uint32_t tmem_base_s;
  const int K = Mp, KH = Mp / 2;                      // one B stage = half a plane: 128 columns x KH bytes
  const int plane_a = I8_BM * K;                      // bytes of one A digit plane
  const int stage_b = I8_BN * KH;
  unsigned char* smA = smem;                          // [na][plane_a]
  unsigned char* smB = smem + (size_t)na * plane_a;   // [I8_STAGES][stage_b]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncol_tiles = Mp / I8_BN;
  const int ksteps_h = KH / 32;                       // MMAs (K = 32 bytes) per half plane

  if (tid == 0) {
    for (int s = 0; s < I8_STAGES; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    mbar_init(&a_ready, I8_EPI_THREADS / 32);
    mbar_init(&a_free, 1);
    for (int s = 0; s < 4; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], I8_EPI_THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {  // TMEM: 512 columns (4 slots x 128 int32 accumulator columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 8) {
    // ================= A-digit extraction + epilogue (256 threads) =================
    const int half = warp >> 2;             // which 64 of the tile's 128 columns this warp drains
    const int lane_base = (warp & 3) * 32;  // TMEM lanes (= tile rows) this warp may access
    uint32_t ph_afree = 0, ph_full[4] = {0, 0, 0, 0};
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int rb = tile / ncol_tiles, ct = tile - rb * ncol_tiles;
      const long long r0 = (long long)rb * I8_BM;
      // ---- (a) digits of I: unit = (row r, 16-byte K chunk kc); a lane group of 8 consecutive rows x 4 chunks writes
      //      512 contiguous bytes per plane (bank-conflict free)
      if (it > 0) { mbar_wait(&a_free, ph_afree); ph_afree ^= 1u; }
      const int nkc = K / 16;
      for (int u = tid; u < (I8_BM / 8) * (nkc / 4) * 32; u += I8_EPI_THREADS) {
        const int l = u & 31, blk = u >> 5;
        const int rgrp = blk / (nkc / 4), kq = blk - rgrp * (nkc / 4);
        const int r = rgrp * 8 + (l & 7), kc = kq * 4 + (l >> 3);
        const long long gr = r0 + r;
        int4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          v[q] = gr < R ? __ldg(reinterpret_cast<const int4*>(Ix + gr * Mp + kc * 16 + q * 4)) : make_int4(0, 0, 0, 0);
        const uint32_t off = (uint32_t)rgrp * (uint32_t)(K * 8) + (uint32_t)kc * 128u + (uint32_t)(l & 7) * 16u;
        for (int a = 0; a < na; ++a) {
          const int sh = 7 * a;
          uint4 o;
          o.x = ((v[0].x >> sh) & 127) | (((v[0].y >> sh) & 127) << 8) | (((v[0].z >> sh) & 127) << 16) | (((v[0].w >> sh) & 127) << 24);
          o.y = ((v[1].x >> sh) & 127) | (((v[1].y >> sh) & 127) << 8) | (((v[1].z >> sh) & 127) << 16) | (((v[1].w >> sh) & 127) << 24);
          o.z = ((v[2].x >> sh) & 127) | (((v[2].y >> sh) & 127) << 8) | (((v[2].z >> sh) & 127) << 16) | (((v[2].w >> sh) & 127) << 24);
          o.w = ((v[3].x >> sh) & 127) | (((v[3].y >> sh) & 127) << 8) | (((v[3].z >> sh) & 127) << 16) | (((v[3].w >> sh) & 127) << 24);
          *reinterpret_cast<uint4*>(smA + (size_t)a * plane_a + off) = o;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready);
      // ---- (b) epilogue: groups complete in the order s = na-1, na-2, ..., -(I8_NB-1)
      double out[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) out[j] = 0.0;
      for (int s = na - 1; s >= -(I8_NB - 1); --s) {
        const int slot = s & 3;
        mbar_wait(&t_full[slot], ph_full[slot]);
        ph_full[slot] ^= 1u;
        tc_fence_after();
        const double w = ldexp(1.0, 7 * (s - 1));  // 128^(s-1), exact
        const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(slot * I8_BN + half * 64);
        uint32_t v0[32], v1[32];
        tc_ld32(taddr, v0);
        tc_ld32(taddr + 32u, v1);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[slot]);  // the slot may be overwritten
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          out[j] = fma((double)(int)v0[j], w, out[j]);
          out[32 + j] = fma((double)(int)v1[j], w, out[32 + j]);
        }
      }
      const long long row = r0 + lane_base + lane;
      if (row < R) {
        const int c0 = ct * I8_BN + half * 64;
        double* dst = Bc + row * Mp + c0;
#pragma unroll
        for (int j = 0; j < 64; j += 2)
          *reinterpret_cast<double2*>(dst + j) = make_double2(out[j] * colscale[c0 + j], out[j + 1] * colscale[c0 + j + 1]);
      }
    }
  } else if (warp == 8) {
    // ================= producer: planes of Cs =================
    if (lane == 0) {
      uint32_t n = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int ct = tile % ncol_tiles;
        for (int c = 0; c < I8_NB; ++c)
          for (int h = 0; h < 2; ++h, ++n) {
            const int st = n % I8_STAGES;
            if (n >= I8_STAGES) mbar_wait(&empty_b[st], ((n / I8_STAGES) - 1) & 1u);
            mbar_expect_tx(&full_b[st], (unsigned)stage_b);
            bulk_load_1d(smB + (size_t)st * stage_b, Bd + (((size_t)c * ncol_tiles + ct) * 2 + h) * stage_b, (unsigned)stage_b, &full_b[st]);
          }
      }
    }
  } else if (lane == 0) {
    // ================= MMA issuer (warp 9, one thread) =================
    uint32_t n = 0, ph_aready = 0;
    uint32_t hosted[4] = {0, 0, 0, 0};  // groups started in each accumulator slot so far (over all tiles)
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      mbar_wait(&a_ready, ph_aready);
      ph_aready ^= 1u;
      tc_fence_after();
      for (int c = 0; c < I8_NB; ++c) {
        for (int h = 0; h < 2; ++h, ++n) {
          const int st = n % I8_STAGES;
          mbar_wait(&full_b[st], (n / I8_STAGES) & 1u);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(smB + (size_t)st * stage_b);
          for (int a = 0; a < na; ++a) {
            const int s = a - c, slot = s & 3;
            const bool first_pair = (c == 0 || a == 0);  // group s receives its first digit pair in this plane
            if (first_pair && h == 0) {
              if (hosted[slot] > 0) {  // slot reuse: the epilogue must have drained the group that lived here before
                mbar_wait(&t_empty[slot], (hosted[slot] - 1) & 1u);
                tc_fence_after();
              }
              hosted[slot] += 1;
            }
            const uint32_t a_addr = smem_u32(smA + (size_t)a * plane_a);
            for (int kk = 0; kk < ksteps_h; ++kk) {
              const uint32_t acc = (first_pair && h == 0 && kk == 0) ? 0u : 1u;
              tc_mma_i8(tmem_base + (uint32_t)(slot * I8_BN), umma_desc(a_addr + (uint32_t)(h * ksteps_h + kk) * 256u, (uint32_t)K * 8u),
                        umma_desc(b_addr + (uint32_t)kk * 256u, (uint32_t)KH * 8u), I8_IDESC, acc);
            }
          }
          tc_commit(&empty_b[st]);  // the stage is free once these MMAs have read it
        }
        // plane c done: group na-1-c has all its pairs; after the last plane every remaining group has
        const int s_hi = na - 1 - c, s_lo = (c == I8_NB - 1) ? -(I8_NB - 1) : s_hi;
        for (int s = s_hi; s >= s_lo; --s) tc_commit(&t_full[s & 3]);
      }
      tc_commit(&a_free);  // the A planes may be overwritten for the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ---- host: split Cs into digit planes, once per model --------------------------------------------------------------
// Returns 0 and fills the device arrays when the integer path applies (Mp a multiple of 128, at most 384, populations
// below 2^21), 1 when it does not (the FP64 DMMA kernel of contract.cu is used), < 0 on a CUDA error.
int seir_contract_i8_setup(seir_model* m, const double* h_cs /*[Mp][Mp], Cs[j][i]*/, double max_population) {
  m->i8_na = 0;
  const int Mp = m->Mp;
  if (Mp % I8_BN != 0 || Mp > 384) return 1;
  int na = 1;
  while (na < 6 && ldexp(1.0, 7 * na) <= max_population) ++na;  // infectious counts never exceed the population
  if (na > 3) return 1;
  const int KH = Mp / 2, nct = Mp / I8_BN;
  const size_t stage_b = (size_t)I8_BN * KH;
  std::vector<signed char> bd((size_t)I8_NB * nct * 2 * stage_b, 0);
  std::vector<double> scale(Mp, 1.0);
  for (int i = 0; i < Mp; ++i) {
    double cmax = 0.0;
    for (int j = 0; j < Mp; ++j) cmax = fmax(cmax, fabs(h_cs[(size_t)j * Mp + i]));
    int e = 0;
    if (cmax > 0.0) {
      frexp(cmax, &e);  // cmax = f 2^e, f in [0.5, 1)
      e += 1;           // |Cs / 2^e| < 0.5: every digit, the first included, lies in [-64, 64]
    }
    scale[i] = ldexp(1.0, e);
    const int ct = i / I8_BN, nn = i % I8_BN;
    for (int j = 0; j < Mp; ++j) {
      double r = ldexp(h_cs[(size_t)j * Mp + i], -e);  // exact scaling
      const int h = j / KH, kl = j % KH;
      const size_t off = (size_t)(nn / 8) * ((size_t)KH * 8) + (size_t)(kl / 16) * 128 + (size_t)(nn % 8) * 16 + (size_t)(kl % 16);
      for (int c = 0; c < I8_NB; ++c) {
        const double t = r * 128.0;   // exact
        const double d = nearbyint(t);
        bd[(((size_t)c * nct + ct) * 2 + h) * stage_b + off] = (signed char)d;
        r = t - d;                    // exact, in [-0.5, 0.5]
      }
    }
  }
  SEIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->d_cs_i8), bd.size()));
  SEIR_CUDA(cudaMemcpy(m->d_cs_i8, bd.data(), bd.size(), cudaMemcpyHostToDevice));
  SEIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->d_cs_scale), sizeof(double) * Mp));
  SEIR_CUDA(cudaMemcpy(m->d_cs_scale, scale.data(), sizeof(double) * Mp, cudaMemcpyHostToDevice));
  m->i8_na = na;
  return 0;
}

int seir_launch_contract_i8(seir_chains* c, cudaStream_t s) {
  const seir_model* m = c->model;
  const long long R = (long long)c->B * m->T;
  const int ntiles = (int)((R + I8_BM - 1) / I8_BM) * (m->Mp / I8_BN);
  const size_t smem = (size_t)m->i8_na * I8_BM * m->Mp + (size_t)I8_STAGES * I8_BN * (m->Mp / 2);
  static int sms = 0;
  static size_t attr = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (attr != smem) {
    SEIR_CUDA(cudaFuncSetAttribute(seir_contract_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  seir_contract_i8_kernel<<<ntiles < sms ? ntiles : sms, I8_THREADS, smem, s>>>(R, m->Mp, m->i8_na, ntiles, c->d_I, m->d_cs_i8,
                                                                                m->d_cs_scale, c->d_Bc);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_contract_i8_kernel");
}
