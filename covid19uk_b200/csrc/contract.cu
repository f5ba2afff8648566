// K2: the commuting-matrix contraction of the force of infection (model_spec.py:262),
//     Bc[b,t,i] = sum_j Cstar[i,j] * I[b,t,j] / N[j],
// batched over chains and days as one GEMM  [B*T, Mp] x [Mp, Mp]  (Cstar is symmetric, so the day slab
// of infectious counts is the row-major left operand and Cstar itself the right operand).
//
// v1: shared-memory tiled FP64 FMA kernel, 64x64x16 tiles, 4x4 register micro-tiles.
#include "seir_internal.cuh"

#define CT_BM 64
#define CT_BN 64
#define CT_BK 16

__global__ void __launch_bounds__(256) seir_contract_kernel(long long R, int Mp, const int* __restrict__ Ix,
                                                            const double* __restrict__ rN, const double* __restrict__ cstar,
                                                            double* __restrict__ Bc) {
  __shared__ __align__(16) double As[CT_BK][CT_BM + 4];
  __shared__ __align__(16) double Bs[CT_BK][CT_BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long r0 = (long long)blockIdx.y * CT_BM;
  const int c0 = blockIdx.x * CT_BN;

  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  const int a_row = tid >> 2, a_k = (tid & 3) * 4;   // 64 rows x 4 int4
  const int b_k = tid >> 4, b_c = (tid & 15) * 4;    // 16 k x 16 groups of 4 doubles
  const long long ar = r0 + a_row;

  for (int k0 = 0; k0 < Mp; k0 += CT_BK) {
    int4 ai = make_int4(0, 0, 0, 0);
    if (ar < R) ai = *reinterpret_cast<const int4*>(Ix + ar * Mp + k0 + a_k);
    const double2 n01 = *reinterpret_cast<const double2*>(rN + k0 + a_k);
    const double2 n23 = *reinterpret_cast<const double2*>(rN + k0 + a_k + 2);
    const double2 b01 = *reinterpret_cast<const double2*>(cstar + (size_t)(k0 + b_k) * Mp + c0 + b_c);
    const double2 b23 = *reinterpret_cast<const double2*>(cstar + (size_t)(k0 + b_k) * Mp + c0 + b_c + 2);
    __syncthreads();
    As[a_k + 0][a_row] = (double)ai.x * n01.x;
    As[a_k + 1][a_row] = (double)ai.y * n01.y;
    As[a_k + 2][a_row] = (double)ai.z * n23.x;
    As[a_k + 3][a_row] = (double)ai.w * n23.y;
    *reinterpret_cast<double2*>(&Bs[b_k][b_c]) = b01;
    *reinterpret_cast<double2*>(&Bs[b_k][b_c + 2]) = b23;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CT_BK; ++k) {
      const double2 a01 = *reinterpret_cast<const double2*>(&As[k][ty * 4]);
      const double2 a23 = *reinterpret_cast<const double2*>(&As[k][ty * 4 + 2]);
      const double2 q01 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4]);
      const double2 q23 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4 + 2]);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y};
      const double q[4] = {q01.x, q01.y, q23.x, q23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], q[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = r0 + ty * 4 + i;
    if (r < R) {
      double* dst = Bc + r * Mp + c0 + tx * 4;
      *reinterpret_cast<double2*>(dst) = make_double2(acc[i][0], acc[i][1]);
      *reinterpret_cast<double2*>(dst + 2) = make_double2(acc[i][2], acc[i][3]);
    }
  }
}

int seir_launch_contract(seir_chains* c, cudaStream_t s) {
  const seir_model* m = c->model;
  const long long R = (long long)c->B * m->T;
  dim3 grid(m->Mp / CT_BN, (unsigned)((R + CT_BM - 1) / CT_BM));
  seir_contract_kernel<<<grid, 256, 0, s>>>(R, m->Mp, c->d_I, m->d_rN, m->d_cstar, c->d_Bc);
  seir_count_launch(1);
  return seir_cuda_check(cudaGetLastError(), "seir_contract_kernel");
}
