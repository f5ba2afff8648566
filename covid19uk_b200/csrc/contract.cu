// K2: the commuting-matrix contraction of the force of infection (model_spec.py:262),
//     Bc[b,t,i] = sum_j Cstar[i,j] * I[b,t,j] / N[j],
// batched over chains and days as one GEMM  [B*T, Mp] x [Mp, Mp].  Cstar is symmetric, so the day slab of
// infectious counts is the row-major left operand; the right operand is Cs[j][i] = Cstar[j][i] / N[j]
// (1/N folded into the matrix once at model creation).
//
// FP64 tensor-core path: mma.sync.aligned.m8n8k4.f64 (DMMA).  ncu on the v1 FMA kernel showed the FP64
// pipe at 32 % with issue slots at 23 % (profiles/r01_v1_ncu_full_summary.md): operand delivery, not
// HBM (DRAM 1 %), was the limit, which is the case the north star names for DMMA.
//   CTA tile 64 x 64, 4 warps, warp tile 32 x 32 (4 x 4 MMA tiles, 32 accumulator doubles / thread),
//   BK = 16, two shared-memory stages, next tile's global loads issued before the current tile's MMAs.
// Measured 31.4 TFLOP/s = 89 % of cuBLAS DGEMM 4096^3 on the same GPU.  Tried and measured equal (197-209 us at the UK
// size, B = 256): row tiles of 40..80 rows chosen per launch against wave quantisation (2016 CTAs on 592 slots), warp
// strips of full tile height, and mma.sync.m16n8k16.f64 -- which ptxas lowers to eight DMMA.8x8x4 on sm_100a, the only
// FP64 tensor shape in the SASS.  The kernel sits at the DMMA issue rate, not at a tiling or scheduling loss.
#include <stdlib.h>

#include "seir_internal.cuh"

#define CT_BM 64
#define CT_BN 64
#define CT_BK 16
#define CT_AS (CT_BK + 4)  // A tile row stride (doubles): fragment loads hit 16 distinct bank pairs
#define CT_BS (CT_BN + 4)  // B tile row stride

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128) seir_contract_kernel(long long R, int Mp, const int* __restrict__ Ix,
                                                            const double* __restrict__ cs, double* __restrict__ Bc) {
  __shared__ __align__(16) double As[2][CT_BM][CT_AS];
  __shared__ __align__(16) double Bs[2][CT_BK][CT_BS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;  // warp tile origin inside the CTA tile
  const long long r0 = (long long)blockIdx.y * CT_BM;
  const int c0 = blockIdx.x * CT_BN;

  // global -> register staging: A 64 rows x 16 ints = 256 int4 (2 / thread); B 16 x 64 doubles = 512 double2 (4 / thread)
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;  // rows a_row and a_row + 32
  const int b_k = tid >> 5, b_c = (tid & 31) * 2;   // k rows b_k, +4, +8, +12
  int4 ra[2];
  double2 rb[4];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long r = r0 + a_row + 32 * h;
      ra[h] = (r < R) ? __ldg(reinterpret_cast<const int4*>(Ix + r * Mp + k0 + a_k)) : make_int4(0, 0, 0, 0);
    }
#pragma unroll
    for (int h = 0; h < 4; ++h)
      rb[h] = __ldg(reinterpret_cast<const double2*>(cs + (size_t)(k0 + b_k + 4 * h) * Mp + c0 + b_c));
  };
  auto store_tile = [&](int st) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double* dst = &As[st][a_row + 32 * h][a_k];
      *reinterpret_cast<double2*>(dst) = make_double2((double)ra[h].x, (double)ra[h].y);
      *reinterpret_cast<double2*>(dst + 2) = make_double2((double)ra[h].z, (double)ra[h].w);
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) *reinterpret_cast<double2*>(&Bs[st][b_k + 4 * h][b_c]) = rb[h];
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = Mp / CT_BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int st = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * CT_BK);
#pragma unroll
    for (int kk = 0; kk < CT_BK; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[st][wr + i * 8 + g][kk + tig];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[st][kk + tig][wc + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (kt + 1 < nk) store_tile(st ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = r0 + wr + i * 8 + g;
    if (r < R) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<double2*>(Bc + r * Mp + c0 + wc + j * 8 + tig * 2) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

static bool contract_uses_i8(const seir_model* m) {
  static int use_i8 = -1;
  if (use_i8 < 0) {
    const char* e = getenv("SEIR_CONTRACT_I8");  // 0: always the FP64 DMMA kernel below; default: the exact int8 tcgen05 kernel
    use_i8 = e ? atoi(e) : 1;                    // (contract_i8.cu) wherever it applies (Mp a multiple of 128, <= 4096)
  }
  return use_i8 && m->i8_na > 0;
}

int seir_launch_contract(seir_chains* c, cudaStream_t s) { return seir_launch_contract_range(c, s, seir_all(c)); }

// can the contraction run on the rows of chains [b0, ...) alone?  (the int8 kernel works on 128-row tiles)
bool seir_contract_range_ok(const seir_chains* c, int b0) {
  return !contract_uses_i8(c->model) || ((long long)b0 * c->model->T) % 128 == 0;
}

int seir_launch_contract_range(seir_chains* c, cudaStream_t s, seir_range r) {
  if (contract_uses_i8(c->model)) return seir_launch_contract_i8_range(c, s, r);
  return seir_launch_contract_f64_range(c, s, r);
}

int seir_launch_contract_f64(seir_chains* c, cudaStream_t s) { return seir_launch_contract_f64_range(c, s, seir_all(c)); }

int seir_launch_contract_f64_range(seir_chains* c, cudaStream_t s, seir_range r) {
  const seir_model* m = c->model;
  const long long R = (long long)r.nb * m->T;
  const size_t cell0 = (size_t)r.b0 * m->T * m->Mp;
  dim3 grid(m->Mp / CT_BN, (unsigned)((R + CT_BM - 1) / CT_BM));
  seir_contract_kernel<<<grid, 128, 0, s>>>(R, m->Mp, c->d_I + cell0, m->d_cs, c->d_Bc + cell0);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_contract_kernel");
}
