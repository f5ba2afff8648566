// K2 on the 5th-generation tensor cores: the commuting contraction  Bc = I . Cs  ([B*T, Mp] x [Mp, Mp], model_spec.py:262)
// as an EXACT integer GEMM on tcgen05 (kind::i8, int32 accumulators in TMEM) -- an error-free splitting ("Ozaki scheme"):
//
//   I[r][j]   = sum_a 256^a A_a[r][j]                 A_a in [0,255] (unsigned bytes; na <= 3 planes: counts < 2^24;
//                                                     planes that are all zero in a tile are skipped at run time)
//   Cs[j][i]  = 2^e_i sum_c 256^-(c+1) D_c[j][i] + delta   D_c in [-128,127] (balanced radix-256 digits of the 48-bit
//                                                     fixed-point value round(Cs 2^(48-e_i)), 2^e_i > 4 max_j |Cs[j][i]|),
//                                                     |delta| <= 2^(e_i-49),
//                                                     nb = 6 planes, per-column scale 2^e_i chosen at model creation
//   Bc[r][i]  = 2^e_i sum_s 256^(s-1) G_s[r][i],      G_s = sum_{a-c=s} A_a D_c   (int32, exact: |G_s| < 2^26)
//
// Every partial product is exact in int32; the only rounding is the truncation of Cs to 46 bits below its column maximum
// and the FP64 additions of the epilogue -- a relative error of ~1e-14 on Bc, far inside the 1e-10 budget of the
// log-probability (tests/test_gpu_contract.py compares against the FP64 DMMA kernel of contract.cu).
// At most 18 (typically 12: infectious counts below 65536) int8 GEMMs of 2 M^2 T flop each replace one FP64 GEMM.
//
// Two kernels:
//   seir_i8_split_kernel     one pass over I: unsigned byte planes per 128-row tile in the K-major 128-byte-swizzled UMMA
//                            layout ([K/128] blocks of [128 rows x 128 B]) + per-tile "plane holds anything" flags
//   seir_contract_i8_kernel  one CTA per SM, persistent over 128 x 128 output tiles:
//     warps 0-7  epilogue: tcgen05.ld finished accumulator groups, FMA them into 64 FP64 registers per thread with weight
//                256^(s-1), release the A region, scale by 2^e_i, store the rows with 32-byte stores
//     warp 8     producer: 1-D bulk copies of the tile's digit planes of I (one copy per plane) and of the pre-split planes
//                of Cs (laid out on the host in the same layout; 4-stage ring, one 16 KB block per stage)
//     warp 9     allocates TMEM (512 columns = 4 accumulator slots of 128), one thread issues tcgen05.mma (kind::i8,
//                M = N = 128, K = 32) in plane-major order (c outer, a inner) so that at most na+1 <= 4 accumulator groups
//                are live: group s = a - c uses slot s mod 4; after plane c group na_t-1-c is complete and is committed
//                to the epilogue.
// Measured (UK size, 256 chains, B200): 104 us for both kernels vs 200 us for the FP64 DMMA kernel.  Per 128 x 128 tile:
// 144 MMAs in ~13 us (~150 cycles per MMA with both operands in shared memory -- the same with the no-swizzle core-matrix
// layout and with 128-byte swizzle: the rate of the instruction, not the layout), ~4 us of output stores; 504 tiles on
// 148 SMs = 4 rounds.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "seir_internal.cuh"
#include "tma.cuh"

#define I8_BM 128
#define I8_BN 128
#define I8_NB 6        // digit planes of Cs (balanced radix 256)
#define I8_STAGES 4
#define I8_KB 128        // bytes of K per swizzled row (one 128-byte swizzle span = 4 MMAs of K = 32)
#define I8_KBLOCK (I8_BM * I8_KB)  // bytes of one [128 rows x 128 B] operand block
#define I8_EPI_THREADS 256
#define I8_THREADS (I8_EPI_THREADS + 64)
#define I8_STAGE_OUT 0

// phase timestamps of one CTA's second tile (build with SEIR_NVCC_EXTRA=-DSEIR_I8_DEBUG, read with seir_debug_i8)
#ifdef SEIR_I8_DEBUG
__device__ long long g_i8dbg[64];
#define TM(k) do { if (blockIdx.x == 5) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_i8dbg[k] = t_; } } while (0)
#else
#define TM(k) do { } while (0)
#endif

__device__ __forceinline__ double i8_int_to_double(int k) {  // exact int32 -> double with one FP64 add
  return __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0;  // 2^52 + 2^31
}

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

// K-major, 128-byte swizzle (cute::UMMA::SmemDescriptor bit layout; layout type SWIZZLE_128B = 2): an operand block is
// [rows][128 bytes of K]; 8-row atoms of 1024 bytes (SBO), inside an atom the 16-byte chunk c of row r sits at chunk c ^ r
// (Swizzle<3,4,3>).  An MMA consumes K = 32 bytes: its start address is the block base + 32 * (k step inside the block);
// LBO is not used by swizzled K-major operands (set to 1 as CUTLASS does).  Blocks must be 1024-byte aligned.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of (row r, K byte k) inside an operand made of [K / 128] consecutive blocks of [nrows x 128 B]
__host__ __device__ __forceinline__ size_t sw128_offset(int r, int k, int nrows) {
  return (size_t)(k >> 7) * ((size_t)nrows * 128) + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 +
         (size_t)((((k & 127) >> 4) ^ (r & 7)) << 4) + (size_t)(k & 15);
}
// cute::UMMA::InstrDescriptor: c_format S32 (2) [4,6), a_format UINT8 (0) [7,10), b_format INT8 (1) [10,13), K-major both,
// N>>3 [17,23), M>>4 [24,29)
#define I8_IDESC ((2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(I8_BN >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24))

// ---- digit planes of the infectious counts: one pass over I (33 MB at the UK size, 256 chains) ---------------------
// grid (row tiles, 4): a CTA splits a quarter of a 128-row tile; unit = (row r, 16-byte K chunk kc); a lane group of 8 consecutive rows x 4 chunks writes 512
// bytes per plane.  Planes are stored per row tile in the 128-byte-swizzled K-major UMMA layout, so the contraction kernel
// brings a plane into shared memory with ONE bulk copy; flags[rt][a] says whether plane a of the tile holds anything.
// kmajor = 0: [a][K block][16 KB] (a plane is contiguous: seir_contract_i8_kernel brings it in with one copy);
// kmajor = 1: [K block][a][16 KB] (the planes of a K block are contiguous: one copy per ring stage of the K-outermost kernel).
__global__ void __launch_bounds__(I8_EPI_THREADS) seir_i8_split_kernel(long long R, int Mp, int na, const int* __restrict__ Ix,
                                                                       unsigned char* __restrict__ planes, int* __restrict__ flags, int kmajor) {
  __shared__ int s_nz[4];
  const int tid = threadIdx.x, lane = tid & 31, rt = blockIdx.x, K = Mp;
  const long long r0 = (long long)rt * I8_BM;
  const size_t plane_a = (size_t)I8_BM * K;
  unsigned char* dst = planes + (size_t)rt * na * plane_a;
  if (tid < 4) s_nz[tid] = 0;
  __syncthreads();
  const int nkc = K / 16;
  const int nunits = (I8_BM / 8) * (nkc / 4) * 32;
  uint32_t nz[3] = {0u, 0u, 0u};
  const int uq = nunits / 4, ulo = blockIdx.y * uq, uhi = blockIdx.y == 3 ? nunits : ulo + uq;  // (nunits is a multiple of 128)
  for (int u0 = ulo + tid; u0 < uhi; u0 += 4 * I8_EPI_THREADS) {  // 4 units = 16 independent 16-byte loads in flight per thread
    int4 v[4][4];
    uint32_t off[4];
    int kblk[4];
#pragma unroll
    for (int b4 = 0; b4 < 4; ++b4) {
      const int u = u0 + b4 * I8_EPI_THREADS;
      const int l = u & 31, blk = u >> 5;
      const int rgrp = blk / (nkc / 4), kq = blk - rgrp * (nkc / 4);
      const int r = rgrp * 8 + (l & 7), kc = kq * 4 + (l >> 3);
      const long long gr = r0 + r;
      const bool in = u < uhi && gr < R;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        v[b4][q] = in ? __ldg(reinterpret_cast<const int4*>(Ix + gr * Mp + kc * 16 + q * 4)) : make_int4(0, 0, 0, 0);
      // (block-local offset; the K block index kc / 8 is applied per plane below)
      off[b4] = (uint32_t)sw128_offset(r, (kc * 16) & 127, I8_BM);
      kblk[b4] = (kc * 16) >> 7;
    }
#pragma unroll
    for (int b4 = 0; b4 < 4; ++b4) {
      if (u0 + b4 * I8_EPI_THREADS >= uhi) break;
      for (int a = 0; a < na; ++a) {
        const int sh = 8 * a;
        uint4 o;
        o.x = ((v[b4][0].x >> sh) & 255) | (((v[b4][0].y >> sh) & 255) << 8) | (((v[b4][0].z >> sh) & 255) << 16) | (((v[b4][0].w >> sh) & 255) << 24);
        o.y = ((v[b4][1].x >> sh) & 255) | (((v[b4][1].y >> sh) & 255) << 8) | (((v[b4][1].z >> sh) & 255) << 16) | (((v[b4][1].w >> sh) & 255) << 24);
        o.z = ((v[b4][2].x >> sh) & 255) | (((v[b4][2].y >> sh) & 255) << 8) | (((v[b4][2].z >> sh) & 255) << 16) | (((v[b4][2].w >> sh) & 255) << 24);
        o.w = ((v[b4][3].x >> sh) & 255) | (((v[b4][3].y >> sh) & 255) << 8) | (((v[b4][3].z >> sh) & 255) << 16) | (((v[b4][3].w >> sh) & 255) << 24);
        const size_t blk = kmajor ? ((size_t)kblk[b4] * na + a) * I8_KBLOCK : (size_t)a * plane_a + (size_t)kblk[b4] * I8_KBLOCK;
        *reinterpret_cast<uint4*>(dst + blk + off[b4]) = o;
        nz[a] |= o.x | o.y | o.z | o.w;
      }
    }
  }
  for (int a = 1; a < na; ++a)
    if (__any_sync(0xffffffffu, nz[a] != 0u) && lane == 0) atomicOr(&s_nz[a], 1);
  __syncthreads();
  if (tid < 4 && (tid == 0 || s_nz[tid])) atomicOr(flags + rt * 4 + tid, 1);  // (flags zeroed by the launcher)
}

__global__ void __launch_bounds__(I8_THREADS, 1) seir_contract_i8_kernel(long long R, int Mp, int na, int ntiles,
                                                                         const unsigned char* __restrict__ planes,
                                                                         const int* __restrict__ flags,
                                                                         const signed char* __restrict__ Bd,
                                                                         const double* __restrict__ colscale, double* __restrict__ Bc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full_b[I8_STAGES], empty_b[I8_STAGES], a_ready, a_region_free, t_full[4], t_empty[4];
  __shared__ uint32_t tmem_base_s;
  const int K = Mp, nkb = Mp / I8_KB;                 // one B stage = one [128 columns x 128 B] block of a plane
  const int plane_a = I8_BM * K;                      // bytes of one A digit plane
  const int stage_b = I8_KBLOCK;
  unsigned char* smA = smem;                          // [na][plane_a]
  const size_t a_region = (size_t)na * plane_a > (size_t)I8_STAGE_OUT ? (size_t)na * plane_a : (size_t)I8_STAGE_OUT;
  unsigned char* smB = smem + a_region;               // [I8_STAGES][stage_b]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncol_tiles = Mp / I8_BN;

  if (tid == 0) {
    for (int s = 0; s < I8_STAGES; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    mbar_init(&a_ready, 1);
    mbar_init(&a_region_free, 1);  // one arrival per tile: the commit behind the tile's last MMA (the MMA issuer)
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
    // ================= epilogue (256 threads) =================
    const int half = warp >> 2;             // which 64 of the tile's 128 columns this warp drains
    const int lane_base = (warp & 3) * 32;  // TMEM lanes (= tile rows) this warp may access
    uint32_t ph_full[4] = {0, 0, 0, 0};
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int rb = tile / ncol_tiles, ct = tile - rb * ncol_tiles;
      const long long r0 = (long long)rb * I8_BM;
      int na_t = 1;  // planes of this row tile that hold anything (seir_i8_split_kernel)
      for (int a = 1; a < na; ++a)
        if (__ldg(flags + rb * 4 + a)) na_t = a + 1;
      if (tid == 0 && it == 1) TM(0);
      // ---- (b) epilogue: groups complete in the order s = na-1, na-2, ..., -(I8_NB-1)
      double out[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) out[j] = 0.0;
      for (int s = na_t - 1; s >= -(I8_NB - 1); --s) {
        const int slot = s & 3;
        mbar_wait(&t_full[slot], ph_full[slot]);
        ph_full[slot] ^= 1u;
        tc_fence_after();
        if (tid == 0 && it == 1) TM(10 + (na_t - 1 - s));
        const double w = ldexp(1.0, 8 * (s - 1));  // 256^(s-1), exact
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
          // int32 -> double by the 2^52 + 2^31 trick (one DADD on the FP64 pipe, exact): the conversion unit (I2F.F64) runs
          // at a quarter of the FP64 rate and this loop is 448 conversions per thread and tile (tools/ubench/fp64_rates.cu)
          out[j] = fma(i8_int_to_double((int)v0[j]), w, out[j]);
          out[32 + j] = fma(i8_int_to_double((int)v1[j]), w, out[32 + j]);
        }
      }
      if (tid == 0 && it == 1) TM(3);
      // (the A planes were released by the commit behind the tile's last MMA: the producer refills them and the next tile's
      //  first groups run while this tile is still being drained and stored)
      // ---- store: thread <-> row, 64 consecutive columns as sixteen 32-byte stores (st.global.v4.f64: every store fills
      //      a whole 32-byte sector; a 16-byte store pattern took 8.4 us per tile, a shared-memory transposition 3.5 us but
      //      held the A region until the end)
      {
        const int c0 = ct * I8_BN + half * 64;
        const long long row = r0 + lane_base + lane;
        if (row < R) {
          double* dst = Bc + row * Mp + c0;
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const double x0 = out[j] * colscale[c0 + j], x1 = out[j + 1] * colscale[c0 + j + 1], x2 = out[j + 2] * colscale[c0 + j + 2],
                         x3 = out[j + 3] * colscale[c0 + j + 3];
            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
          }
        }
      }
      if (tid == 0 && it == 1) TM(4);
    }
  } else if (warp == 8) {
    // ================= producer: planes of Cs =================
    if (seir_elect_one()) {
      uint32_t n = 0, itp = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++itp) {
        const int rb = tile / ncol_tiles, ct = tile - rb * ncol_tiles;
        int issued = 0;
        for (int c = 0; c < I8_NB; ++c)
          for (int h = 0; h < nkb; ++h, ++n) {
            if (issued++ == I8_STAGES) {
              // (the first ring-full of Cs stages of this tile is already on its way: it does not depend on the A region)
              // digit planes of the row tile: the A region is free once every epilogue warp has drained the previous tile
              if (itp > 0) mbar_wait(&a_region_free, (itp - 1) & 1u);
              int na_t = 1;
              for (int a = 1; a < na; ++a)
                if (__ldg(flags + rb * 4 + a)) na_t = a + 1;
              mbar_expect_tx(&a_ready, (unsigned)(na_t * plane_a));
              for (int a = 0; a < na_t; ++a)
                bulk_load_1d(smA + (size_t)a * plane_a, planes + ((size_t)rb * na + a) * plane_a, (unsigned)plane_a, &a_ready);
            }
            const int st = n % I8_STAGES;
            if (n >= I8_STAGES) mbar_wait(&empty_b[st], ((n / I8_STAGES) - 1) & 1u);
#ifdef SEIR_I8_EXPERIMENT_NOB  // timing experiment only (wrong results): 16 bytes per stage instead of the half plane
            mbar_expect_tx(&full_b[st], 16u);
            bulk_load_1d(smB + (size_t)st * stage_b, Bd + (((size_t)c * ncol_tiles + ct) * nkb + h) * stage_b, 16u, &full_b[st]);
#else
            mbar_expect_tx(&full_b[st], (unsigned)stage_b);
            bulk_load_1d(smB + (size_t)st * stage_b, Bd + (((size_t)c * ncol_tiles + ct) * nkb + h) * stage_b, (unsigned)stage_b, &full_b[st]);
#endif
          }
      }
    }
  } else if (seir_elect_one()) {
    // ================= MMA issuer (warp 9, one elected thread) =================
    uint32_t n = 0, ph_aready = 0;
    uint32_t hosted[4] = {0, 0, 0, 0};  // groups started in each accumulator slot so far (over all tiles)
    uint64_t a_desc_plane[3];
    for (int a = 0; a < 3; ++a) a_desc_plane[a] = umma_desc(smem_u32(smA + (size_t)(a < na ? a : 0) * plane_a));
    int itm = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++itm) {
      if (itm == 1) TM(30);
      mbar_wait(&a_ready, ph_aready);
      ph_aready ^= 1u;
      tc_fence_after();
      if (itm == 1) TM(31);
      int na_t = 1;
      for (int a = 1; a < na; ++a)
        if (__ldg(flags + (tile / ncol_tiles) * 4 + a)) na_t = a + 1;
      for (int c = 0; c < I8_NB; ++c) {
        for (int h = 0; h < nkb; ++h, ++n) {
          const int st = n % I8_STAGES;
          if (itm == 1 && c == 2 && h < 2) TM(50 + 3 * h);
          mbar_wait(&full_b[st], (n / I8_STAGES) & 1u);
          tc_fence_after();
          if (itm == 1 && c == 2 && h < 2) TM(51 + 3 * h);
          const uint64_t b_desc0 = umma_desc(smem_u32(smB + (size_t)st * stage_b));
          for (int a = 0; a < na_t; ++a) {
            const int s = a - c, slot = s & 3;
            const bool first_pair = (c == 0 || a == 0);  // group s receives its first digit pair in this plane
            if (first_pair && h == 0) {
              if (hosted[slot] > 0) {  // slot reuse: the epilogue must have drained the group that lived here before
                mbar_wait(&t_empty[slot], (hosted[slot] - 1) & 1u);
                tc_fence_after();
              }
              hosted[slot] += 1;
            }
            // K block h of plane a; a K step of 32 bytes advances both start addresses by 2 descriptor units (low field, no
            // carry: shared-memory addresses stay below 2^18)
            const uint64_t a_desc0 = a_desc_plane[a] + (uint64_t)h * (I8_KBLOCK >> 4);
            const uint32_t d_addr = tmem_base + (uint32_t)(slot * I8_BN);
            tc_mma_i8(d_addr, a_desc0, b_desc0, I8_IDESC, (first_pair && h == 0) ? 0u : 1u);
#pragma unroll
            for (int kk = 1; kk < I8_KB / 32; ++kk) tc_mma_i8(d_addr, a_desc0 + (uint64_t)kk * 2u, b_desc0 + (uint64_t)kk * 2u, I8_IDESC, 1u);
#ifdef SEIR_I8_EXPERIMENT_REP  // timing experiment only (wrong results): extra MMAs per step to separate per-MMA from per-stage cost
            for (int rep = 0; rep < SEIR_I8_EXPERIMENT_REP; ++rep)
              for (int kk = 0; kk < I8_KB / 32; ++kk) tc_mma_i8(d_addr, a_desc0 + (uint64_t)kk * 2u, b_desc0 + (uint64_t)kk * 2u, I8_IDESC, 1u);
#endif
          }
          if (itm == 1 && c == 2 && h < 2) TM(52 + 3 * h);
          tc_commit(&empty_b[st]);  // the stage is free once these MMAs have read it
        }
        // plane c done: group na-1-c has all its pairs; after the last plane every remaining group has
        const int s_hi = na_t - 1 - c, s_lo = (c == I8_NB - 1) ? -(I8_NB - 1) : s_hi;
        for (int s = s_hi; s >= s_lo; --s) tc_commit(&t_full[s & 3]);
        if (c == I8_NB - 1) tc_commit(&a_region_free);  // every MMA of the tile has read its A planes once this commit arrives
        if (itm == 1) TM(40 + c);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ---- the same contraction for long K (Mp > 384: BASELINE.json configs[4], 2000 regions) ------------------------------
// The digit planes of a 128-row tile no longer fit shared memory (3 x 128 x 2048 B), so the loop order turns around: K blocks
// outermost, and inside a K block EVERY digit pair (a, c).  All groups s = a - c in [-5, 2] are then live for the whole
// tile: 8 groups x 64 int32 columns = the 512 TMEM columns, hence 128 x 64 output tiles.  One ring stage = K block h of the
// tile's na_t planes of I (16 KB each) + of the 6 digit planes of the tile's 64 columns of Cs (8 KB each: the first or
// second half of the [128 columns x 128 B] block the host laid out); two stages of 96 KB.  Every operand byte is fetched
// once per tile.  Same exact integers, same epilogue order of additions as seir_contract_i8_kernel.
#define I8L_BN 64
#define I8L_STAGES 2
#define I8L_A_BYTES (3 * I8_KBLOCK)               // up to three planes of I per stage
#define I8L_B_BYTES (I8_NB * I8L_BN * I8_KB)     // six half blocks of Cs per stage
#define I8L_STAGE_BYTES (I8L_A_BYTES + I8L_B_BYTES)
#define I8L_IDESC ((2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(I8L_BN >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24))

__global__ void __launch_bounds__(I8_THREADS, 1) seir_contract_i8_longk_kernel(long long R, int Mp, int na, int ntiles,
                                                                               const unsigned char* __restrict__ planes,
                                                                               const int* __restrict__ flags,
                                                                               const signed char* __restrict__ Bd,
                                                                               const double* __restrict__ colscale, double* __restrict__ Bc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full_b[I8L_STAGES], empty_b[I8L_STAGES], acc_full, acc_free;
  __shared__ uint32_t tmem_base_s;
  const int nkb = Mp / I8_KB;
  const size_t plane_a = (size_t)I8_BM * Mp;  // bytes of one digit plane of a row tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncol_tiles = Mp / I8L_BN;
  if (tid == 0) {
    for (int s = 0; s < I8L_STAGES; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    mbar_init(&acc_full, 1);                      // the commit behind the tile's last MMA
    mbar_init(&acc_free, I8_EPI_THREADS / 32);   // every epilogue warp has read the accumulators
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 8) {
    // ================= epilogue =================
    const int half = warp >> 2;             // which 32 of the tile's 64 columns this warp drains
    const int lane_base = (warp & 3) * 32;  // TMEM lanes (= tile rows) this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int rb = tile / ncol_tiles, ct = tile - rb * ncol_tiles;
      const long long r0 = (long long)rb * I8_BM;
      int na_t = 1;
      for (int a = 1; a < na; ++a)
        if (__ldg(flags + rb * 4 + a)) na_t = a + 1;
      mbar_wait(&acc_full, (uint32_t)it & 1u);
      tc_fence_after();
      double out[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) out[j] = 0.0;
      for (int s = na_t - 1; s >= -(I8_NB - 1); --s) {
        const double w = ldexp(1.0, 8 * (s - 1));  // 256^(s-1), exact
        const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)((s + I8_NB - 1) * I8L_BN + half * 32);
        uint32_t v[32];
        tc_ld32(taddr, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) out[j] = fma(i8_int_to_double((int)v[j]), w, out[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_free);  // the next tile's MMAs may overwrite the accumulators
      const int c0 = ct * I8L_BN + half * 32;
      const long long row = r0 + lane_base + lane;
      if (row < R) {
        double* dst = Bc + row * Mp + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const double x0 = out[j] * colscale[c0 + j], x1 = out[j + 1] * colscale[c0 + j + 1], x2 = out[j + 2] * colscale[c0 + j + 2],
                       x3 = out[j + 3] * colscale[c0 + j + 3];
          asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
        }
      }
    }
  } else if (warp == 8) {
    // ================= producer =================
    if (seir_elect_one()) {
      uint32_t n = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int rb = tile / ncol_tiles, ct = tile - rb * ncol_tiles;
        int na_t = 1;
        for (int a = 1; a < na; ++a)
          if (__ldg(flags + rb * 4 + a)) na_t = a + 1;
        const unsigned char* a_src = planes + (size_t)rb * na * plane_a;  // [K block][plane][16 KB] (seir_i8_split_kernel, kmajor)
        const signed char* b_src = Bd + (size_t)ct * nkb * I8L_B_BYTES;   // [K block][digit plane][8 KB]
        for (int h = 0; h < nkb; ++h, ++n) {
          const int st = n % I8L_STAGES;
          if (n >= I8L_STAGES) mbar_wait(&empty_b[st], ((n / I8L_STAGES) - 1) & 1u);
          unsigned char* sa = smem + (size_t)st * I8L_STAGE_BYTES;
          unsigned char* sb = sa + I8L_A_BYTES;
          // TWO copies per stage: a 1-D bulk copy costs ~700 cycles + its bytes and the copies of an SM do not overlap
          // (tools/ubench/tma_feed.cu) -- with one copy per plane (8 per stage) the kernel was bound by their number
          mbar_expect_tx(&full_b[st], (unsigned)(na_t * I8_KBLOCK + I8L_B_BYTES));
          bulk_load_1d(sa, a_src + (size_t)h * na * I8_KBLOCK, (unsigned)(na_t * I8_KBLOCK), &full_b[st]);
          bulk_load_1d(sb, b_src + (size_t)h * I8L_B_BYTES, (unsigned)I8L_B_BYTES, &full_b[st]);
        }
      }
    }
  } else if (seir_elect_one()) {
    // ================= MMA issuer (warp 9, one elected thread) =================
    uint32_t n = 0;
    int itm = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++itm) {
      int na_t = 1;
      for (int a = 1; a < na; ++a)
        if (__ldg(flags + (tile / ncol_tiles) * 4 + a)) na_t = a + 1;
      if (itm > 0) {  // the epilogue has drained the previous tile
        mbar_wait(&acc_free, (uint32_t)(itm - 1) & 1u);
        tc_fence_after();
      }
      for (int h = 0; h < nkb; ++h, ++n) {
        const int st = n % I8L_STAGES;
        mbar_wait(&full_b[st], (n / I8L_STAGES) & 1u);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)st * I8L_STAGE_BYTES), sb = sa + I8L_A_BYTES;
        for (int c = 0; c < I8_NB; ++c) {
          const uint64_t b_desc0 = umma_desc(sb + (uint32_t)c * (I8L_BN * I8_KB));
          for (int a = 0; a < na_t; ++a) {
            const uint64_t a_desc0 = umma_desc(sa + (uint32_t)a * I8_KBLOCK);
            const uint32_t d_addr = tmem_base + (uint32_t)((a - c + I8_NB - 1) * I8L_BN);
            const bool first = h == 0 && (c == 0 || a == 0);  // group a - c receives its first digit pair
            // (instruction order, measured at the UK size: these runs of 4 K steps per digit pair 71.6 us; all pairs of a group back to
            //  back, runs of up to 12, 72.5 us; K step outermost, a different accumulator group with every instruction, 87.8 us)
            tc_mma_i8(d_addr, a_desc0, b_desc0, I8L_IDESC, first ? 0u : 1u);
#pragma unroll
            for (int kk = 1; kk < I8_KB / 32; ++kk) tc_mma_i8(d_addr, a_desc0 + (uint64_t)kk * 2u, b_desc0 + (uint64_t)kk * 2u, I8L_IDESC, 1u);
          }
        }
        tc_commit(&empty_b[st]);  // the stage is free once these MMAs have read it
      }
      tc_commit(&acc_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ---- long K, 128 x 128 tiles in TWO passes over K ---------------------------------------------------------------------
// The K-outermost kernel above delivers operands at ~55 B/clk: an N = 64 instruction reads 4 KB of I for 2 KB of Cs.  With
// N = 128 the same bytes of I serve twice the columns, but eight live groups of 128 columns do not fit the 512 TMEM columns --
// so the groups are split: pass 0 accumulates the four most significant groups s in [s_top - 3, s_top] (digit planes 0..3 of
// Cs), pass 1 the rest (planes 4 - s_top .. 5), each pass a full sweep over the K blocks with four 128-column accumulators.
// The planes of I are fetched twice (L2), every digit pair is still multiplied once: -25 % operand bytes per MAC.  Used where
// the tile count does not quantise (Mp > 384: thousands of tiles per SM); at the UK size 504 tiles are 3.4 rounds that cost 4.
#define I8T_A_BYTES (3 * I8_KBLOCK)
#define I8T_B_BYTES (4 * I8_KBLOCK)
#define I8T_STAGE_BYTES (I8T_A_BYTES + I8T_B_BYTES)
#define I8T_STAGES 2

__global__ void __launch_bounds__(I8_THREADS, 1) seir_contract_i8_twopass_kernel(long long R, int Mp, int na, int ntiles,
                                                                                 const unsigned char* __restrict__ planes,
                                                                                 const int* __restrict__ flags,
                                                                                 const signed char* __restrict__ Bd,
                                                                                 const double* __restrict__ colscale, double* __restrict__ Bc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full_b[I8T_STAGES], empty_b[I8T_STAGES], acc_full, acc_free;
  __shared__ uint32_t tmem_base_s;
  const int nkb = Mp / I8_KB;
  const size_t plane_a = (size_t)I8_BM * Mp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncol_tiles = Mp / I8_BN;
  if (tid == 0) {
    for (int s = 0; s < I8T_STAGES; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    mbar_init(&acc_full, 1);                      // the commit behind a pass's last MMA
    mbar_init(&acc_free, I8_EPI_THREADS / 32);   // every epilogue warp has read the pass's accumulators
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // pass p of a tile with na_t planes: groups s in [s_lo, s_hi], digit planes of Cs c in [c_lo, c_hi]
  auto pass_range = [](int na_t, int p, int& s_lo, int& s_hi, int& c_lo, int& c_hi) {
    const int s_top = na_t - 1;
    if (p == 0) { s_hi = s_top; s_lo = s_top - 3; c_lo = 0; c_hi = 3; }
    else { s_hi = s_top - 4; s_lo = -(I8_NB - 1); c_lo = 4 - s_top; c_hi = I8_NB - 1; }
  };

  if (warp < 8) {
    // ================= epilogue =================
    const int half = warp >> 2;             // which 64 of the tile's 128 columns this warp drains
    const int lane_base = (warp & 3) * 32;  // TMEM lanes (= tile rows) this warp may access
    uint32_t npass = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int rb = tile / ncol_tiles, ct = tile - rb * ncol_tiles;
      const long long r0 = (long long)rb * I8_BM;
      int na_t = 1;
      for (int a = 1; a < na; ++a)
        if (__ldg(flags + rb * 4 + a)) na_t = a + 1;
      double out[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) out[j] = 0.0;
      for (int p = 0; p < 2; ++p, ++npass) {
        int s_lo, s_hi, c_lo, c_hi;
        pass_range(na_t, p, s_lo, s_hi, c_lo, c_hi);
        mbar_wait(&acc_full, npass & 1u);
        tc_fence_after();
        for (int s = s_hi; s >= s_lo; --s) {  // (most significant group first over both passes: the order of the other kernels)
          const double w = ldexp(1.0, 8 * (s - 1));  // 256^(s-1), exact
          const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)((s_hi - s) * I8_BN + half * 64);
          uint32_t v0[32], v1[32];
          tc_ld32(taddr, v0);
          tc_ld32(taddr + 32u, v1);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            out[j] = fma(i8_int_to_double((int)v0[j]), w, out[j]);
            out[32 + j] = fma(i8_int_to_double((int)v1[j]), w, out[32 + j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_free);  // the next pass may overwrite the accumulators
      }
      const int c0 = ct * I8_BN + half * 64;
      const long long row = r0 + lane_base + lane;
      if (row < R) {
        double* dst = Bc + row * Mp + c0;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          const double x0 = out[j] * colscale[c0 + j], x1 = out[j + 1] * colscale[c0 + j + 1], x2 = out[j + 2] * colscale[c0 + j + 2],
                       x3 = out[j + 3] * colscale[c0 + j + 3];
          asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "d"(x0), "d"(x1), "d"(x2), "d"(x3) : "memory");
        }
      }
    }
  } else if (warp == 8) {
    // ================= producer =================
    if (seir_elect_one()) {
      uint32_t n = 0;
      const int nct = Mp / I8_BN;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int rb = tile / ncol_tiles, ct = tile - rb * ncol_tiles;
        int na_t = 1;
        for (int a = 1; a < na; ++a)
          if (__ldg(flags + rb * 4 + a)) na_t = a + 1;
        const unsigned char* a_src = planes + (size_t)rb * na * plane_a;  // [K block][plane][16 KB] (seir_i8_split_kernel, kmajor)
        for (int p = 0; p < 2; ++p) {
          int s_lo, s_hi, c_lo, c_hi;
          pass_range(na_t, p, s_lo, s_hi, c_lo, c_hi);
          const int nc = c_hi - c_lo + 1;
          for (int h = 0; h < nkb; ++h, ++n) {
            const int st = n % I8T_STAGES;
            if (n >= I8T_STAGES) mbar_wait(&empty_b[st], ((n / I8T_STAGES) - 1) & 1u);
            unsigned char* sa = smem + (size_t)st * I8T_STAGE_BYTES;
            unsigned char* sb = sa + I8T_A_BYTES;
            mbar_expect_tx(&full_b[st], (unsigned)((na_t + nc) * I8_KBLOCK));
            bulk_load_1d(sa, a_src + (size_t)h * na * I8_KBLOCK, (unsigned)(na_t * I8_KBLOCK), &full_b[st]);
            for (int c = 0; c < nc; ++c)  // plane-major host layout [plane][column tile][K block][16 KB]
              bulk_load_1d(sb + (size_t)c * I8_KBLOCK, Bd + (((size_t)(c_lo + c) * nct + ct) * nkb + h) * I8_KBLOCK, (unsigned)I8_KBLOCK, &full_b[st]);
          }
        }
      }
    }
  } else if (seir_elect_one()) {
    // ================= MMA issuer (warp 9, one elected thread) =================
    uint32_t n = 0, npass = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int na_t = 1;
      for (int a = 1; a < na; ++a)
        if (__ldg(flags + (tile / ncol_tiles) * 4 + a)) na_t = a + 1;
      for (int p = 0; p < 2; ++p, ++npass) {
        int s_lo, s_hi, c_lo, c_hi;
        pass_range(na_t, p, s_lo, s_hi, c_lo, c_hi);
        if (npass > 0) {  // the epilogue has drained the previous pass
          mbar_wait(&acc_free, (npass - 1) & 1u);
          tc_fence_after();
        }
        uint32_t started = 0;  // groups of this pass that have received their first digit pair
        for (int h = 0; h < nkb; ++h, ++n) {
          const int st = n % I8T_STAGES;
          mbar_wait(&full_b[st], (n / I8T_STAGES) & 1u);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)st * I8T_STAGE_BYTES), sb = sa + I8T_A_BYTES;
          for (int c = c_lo; c <= c_hi; ++c) {
            const uint64_t b_desc0 = umma_desc(sb + (uint32_t)(c - c_lo) * I8_KBLOCK);
            for (int a = 0; a < na_t; ++a) {
              const int s = a - c;
              if (s < s_lo || s > s_hi) continue;  // the other pass's group
              const uint32_t slot = (uint32_t)(s_hi - s);
              const uint64_t a_desc0 = umma_desc(sa + (uint32_t)a * I8_KBLOCK);
              const uint32_t d_addr = tmem_base + slot * I8_BN;
              const bool first = !((started >> slot) & 1u);
              started |= 1u << slot;
              tc_mma_i8(d_addr, a_desc0, b_desc0, I8_IDESC, first ? 0u : 1u);
#pragma unroll
              for (int kk = 1; kk < I8_KB / 32; ++kk) tc_mma_i8(d_addr, a_desc0 + (uint64_t)kk * 2u, b_desc0 + (uint64_t)kk * 2u, I8_IDESC, 1u);
            }
          }
          tc_commit(&empty_b[st]);  // the stage is free once these MMAs have read it
        }
        tc_commit(&acc_full);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ---- host: split Cs into digit planes, once per model --------------------------------------------------------------
// Returns 0 and fills the device arrays when the integer path applies (Mp a multiple of 128, at most 4096, populations
// below 2^24), 1 when it does not (the FP64 DMMA kernel of contract.cu is used), < 0 on a CUDA error.
int seir_contract_i8_setup(seir_model* m, const double* h_cs /*[Mp][Mp], Cs[j][i]*/, double max_population) {
  m->i8_na = 0;
  const int Mp = m->Mp;
  if (Mp % I8_BN != 0 || Mp > 4096) return 1;  // (Mp <= 384: seir_contract_i8_kernel; longer K: seir_contract_i8_longk_kernel)
  int na = 1;
  while (na < 6 && ldexp(1.0, 8 * na) <= max_population) ++na;  // infectious counts never exceed the population
  if (na > 3) return 1;
  const int nct = Mp / I8_BN;
  const size_t plane_b = (size_t)I8_BN * Mp;  // one plane of one column tile: [Mp / 128] blocks of [128 columns x 128 B]
  std::vector<signed char> bd((size_t)I8_NB * nct * plane_b, 0);
  std::vector<double> scale(Mp, 1.0);
  for (int i = 0; i < Mp; ++i) {
    double cmax = 0.0;
    for (int j = 0; j < Mp; ++j) cmax = fmax(cmax, fabs(h_cs[(size_t)j * Mp + i]));
    int e = 0;
    if (cmax > 0.0) {
      frexp(cmax, &e);  // cmax = f 2^e, f in [0.5, 1)
      e += 2;           // |Cs / 2^e| < 0.25: the 48-bit fixed-point value stays below 2^46 and its top digit inside [-64, 64]
    }
    scale[i] = ldexp(1.0, e);
    const int ct = i / I8_BN, nn = i % I8_BN;
    for (int j = 0; j < Mp; ++j) {
      long long W = llrint(ldexp(h_cs[(size_t)j * Mp + i], 8 * I8_NB - e));  // round(Cs 2^(48-e)), |W| <= 2^46
      const size_t off = sw128_offset(nn, j, I8_BN);
      for (int c = I8_NB - 1; c >= 0; --c) {  // balanced digits, least significant (plane nb-1) first
        long long d = ((W + 128) & 255) - 128;  // in [-128, 127], congruent to W mod 256
        bd[((size_t)c * nct + ct) * plane_b + off] = (signed char)d;
        W = (W - d) / 256;                      // exact
      }
      // (W is now 0: |top digit| <= 64 because |Cs / 2^e| < 0.25)
    }
  }
  SEIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->d_cs_i8), bd.size()));
  SEIR_CUDA(cudaMemcpy(m->d_cs_i8, bd.data(), bd.size(), cudaMemcpyHostToDevice));
  {  // the same digits regrouped for the K-outermost kernel: [Mp/64 column tiles][K blocks][6 planes][8 KB half block]
    const int nkb = Mp / I8_KB, half = I8L_BN * I8_KB;
    std::vector<signed char> bl(bd.size());
    for (int c = 0; c < I8_NB; ++c)
      for (int ct = 0; ct < 2 * nct; ++ct)
        for (int h = 0; h < nkb; ++h)
          memcpy(&bl[(((size_t)ct * nkb + h) * I8_NB + c) * half],
                 &bd[(((size_t)c * nct + (ct >> 1)) * nkb + h) * I8_KBLOCK + (size_t)(ct & 1) * half], (size_t)half);
    SEIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->d_cs_i8l), bl.size()));
    SEIR_CUDA(cudaMemcpy(m->d_cs_i8l, bl.data(), bl.size(), cudaMemcpyHostToDevice));
  }
  SEIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->d_cs_scale), sizeof(double) * Mp));
  SEIR_CUDA(cudaMemcpy(m->d_cs_scale, scale.data(), sizeof(double) * Mp, cudaMemcpyHostToDevice));
  m->i8_na = na;
  return 0;
}

int seir_launch_contract_i8(seir_chains* c, cudaStream_t s) { return seir_launch_contract_i8_range(c, s, seir_all(c)); }

// rows of chains [r.b0, r.b0 + r.nb): the first row must start a 128-row tile (callers check seir_contract_range_ok)
int seir_launch_contract_i8_range(seir_chains* c, cudaStream_t s, seir_range r) {
  const seir_model* m = c->model;
  const long long Rall = (long long)c->B * m->T, R = (long long)r.nb * m->T, row0 = (long long)r.b0 * m->T;
  if (row0 % I8_BM != 0) return seir_set_error(SEIR_ERR_BAD_ARG, "seir_launch_contract_i8_range: chain range does not start a row tile");
  static int force_longk = -1;
  if (force_longk < 0) {
    const char* e = getenv("SEIR_I8_LONGK");  // 0: the plane-major 128 x 128 kernel where it applies (Mp <= 384; A/B timing)
    force_longk = e ? atoi(e) : 1;
  }
  // the K-outermost kernel is the default for every shape: UK, 256 chains 62 us against 86 us (every operand byte fetched once,
  // 7 rounds of 1008 half-width tiles instead of 4 rounds of 504)
  const bool longk = m->Mp > 384 || force_longk > 0;
  static int use_twopass = -1;
  if (use_twopass < 0) {
    const char* e = getenv("SEIR_I8_TWOPASS");  // 0: the single-pass 128 x 64 kernel also for long K (A/B timing; same results)
    use_twopass = e ? atoi(e) : 1;
  }
  const bool twopass = m->Mp > 384 && use_twopass > 0;
  const int ntiles = (int)((R + I8_BM - 1) / I8_BM) * (m->Mp / ((longk && !twopass) ? I8L_BN : I8_BN));
  size_t a_region = (size_t)m->i8_na * I8_BM * m->Mp;
#if I8_STAGE_OUT > 0
  if (a_region < (size_t)I8_STAGE_OUT) a_region = I8_STAGE_OUT;
#endif
  const size_t smem = twopass ? (size_t)I8T_STAGES * I8T_STAGE_BYTES + 1024
                     : longk  ? (size_t)I8L_STAGES * I8L_STAGE_BYTES + 1024
                              : a_region + (size_t)I8_STAGES * I8_KBLOCK + 1024;  // (+ slack to align the dynamic base to 1024 bytes)
  const int sms = m->sms;
  static size_t attr_dev[SEIR_MAX_DEVICES][3] = {{0}};  // (the opt-in is per device)
  size_t& attr = attr_dev[c->model->device % SEIR_MAX_DEVICES][twopass ? 2 : (longk ? 1 : 0)];
  if (attr != smem) {
    if (twopass) SEIR_CUDA(cudaFuncSetAttribute(seir_contract_i8_twopass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else if (longk) SEIR_CUDA(cudaFuncSetAttribute(seir_contract_i8_longk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else SEIR_CUDA(cudaFuncSetAttribute(seir_contract_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int nrt_all = (int)((Rall + I8_BM - 1) / I8_BM), nrt = (int)((R + I8_BM - 1) / I8_BM), rt0 = (int)(row0 / I8_BM);
  if (!c->d_i8_planes) {
    SEIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_i8_planes), (size_t)nrt_all * m->i8_na * I8_BM * m->Mp));
    SEIR_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_i8_flags), sizeof(int) * 4 * (size_t)nrt_all));
    c->bytes += (int64_t)((size_t)nrt_all * m->i8_na * I8_BM * m->Mp + sizeof(int) * 4 * (size_t)nrt_all);
  }
  unsigned char* planes = c->d_i8_planes + (size_t)rt0 * m->i8_na * I8_BM * m->Mp;
  int* flags = c->d_i8_flags + 4 * (size_t)rt0;
  const size_t cell0 = (size_t)row0 * m->Mp;
  SEIR_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * 4 * (size_t)nrt, s));
  seir_i8_split_kernel<<<dim3(nrt, 4), I8_EPI_THREADS, 0, s>>>(R, m->Mp, m->i8_na, c->d_I + cell0, planes, flags, longk ? 1 : 0);
  if (twopass)
    seir_contract_i8_twopass_kernel<<<ntiles < sms ? ntiles : sms, I8_THREADS, smem, s>>>(R, m->Mp, m->i8_na, ntiles, planes, flags, m->d_cs_i8,
                                                                                          m->d_cs_scale, c->d_Bc + cell0);
  else if (longk)
    seir_contract_i8_longk_kernel<<<ntiles < sms ? ntiles : sms, I8_THREADS, smem, s>>>(R, m->Mp, m->i8_na, ntiles, planes, flags, m->d_cs_i8l,
                                                                                        m->d_cs_scale, c->d_Bc + cell0);
  else
    seir_contract_i8_kernel<<<ntiles < sms ? ntiles : sms, I8_THREADS, smem, s>>>(R, m->Mp, m->i8_na, ntiles, planes, flags, m->d_cs_i8,
                                                                                  m->d_cs_scale, c->d_Bc + cell0);
  seir_count_launch(2);
  return seir_cuda_check(cudaGetLastError(), "seir_contract_i8_kernel");
}

#ifdef SEIR_I8_DEBUG
extern "C" int seir_debug_i8(long long* h) { return (int)cudaMemcpyFromSymbol(h, g_i8dbg, sizeof(long long) * 64); }
#endif
