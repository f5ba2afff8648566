// One (day, metapopulation) cell of the S->E chain-binomial term and of its parameter gradient: the innermost
// arithmetic shared by the streaming log-likelihood kernels (loglik.cu) and the persistent HMC trajectory kernel
// (hmc_traj.cu).
//
//   lam = exp(a_t + beta la_m + sigma s_m) (I + psi W_t Bc) / N_m + eps        (model_spec.py:257-266)
//   term = y log(1 - exp(-lam dt)) - (S - y) lam dt                            (chain binomial, tex:254-268)
//
// Measured on a B200 (tools/ubench/fp64_rates.cu, thread-operations per clock per SM): DFMA/DADD/DMUL 57-62,
// I2F.F64.S32 ~16, F2F.F32.F64 + MUFU.RCP + F2F.F64.F32 ~5 for the group, __frcp_rn (IEEE reciprocal: a software
// sequence) ~4.  The round-1 cell spent as long in three int->double conversions and one __frcp_rn as in its ~40 FP64
// operations.  Here nothing goes through the conversion unit:
//   * int -> double by the 2^52 trick: the integer is dropped into the low word of a double whose high word is
//     0x43300000 and 2^52 (+2^31 for signed values) is subtracted -- one DADD on the FP64 pipe, exact;
//   * 1/x (gradient only, tolerance 1e-8): the FP32 seed is built from the double's words with two integer operations
//     (truncation of the mantissa, exponent re-bias), MUFU.RCP, the result is widened back with three integer
//     operations, then ONE Newton step in FP64: relative error < 2^-43;
//   * log x (value only): 128-bucket table range reduction + degree-7 log1p (as in round 1);
//   * log(1-e^-x) / 1/expm1(x) by even / odd power series for x < 0.05 (truncation < 1e-17).
// Anything outside the fast range (x >= 0.05, x < 2^-126, x <= 0, NaN) takes the library path: NaN log for x < 0 like the
// reference's log(1 - exp(-x)).
#pragma once
#include <stdint.h>

// Polynomial / series coefficients travel as a by-value kernel parameter: they sit in constant bank 0 and FP64
// instructions take them directly as c[0x0][offset] operands (as literals the compiler rebuilds each 64-bit immediate
// with two moves per use; a __constant__ array costs a load per use).
struct ll_coefs {
  double k[14];
};
static const ll_coefs LL_COEFS = {{
    0.14285714285714285, -0.16666666666666666, 0.2, -0.25, 0.3333333333333333, -0.5,  // log1p(r) = r + r^2 (c5 + r (c4 + ...))
    0.6931471805599453,                                                                // ln 2
    5.511463844797178e-06, -3.472222222222222e-04, 0.041666666666666664,             // log(1-e^-x) - log x + x/2, even powers
    3.306878306878307e-05, -1.388888888888889e-03, 0.08333333333333333,              // 1/expm1(x) - 1/x + 1/2, odd powers
    4503599627370496.0}};                                                              // 2^52 (bit pattern for uint_to_double_wide)

#define CELL_HI_LO 0x38100000  // high word of 2^-126: below it the FP32 seed of the reciprocal would be subnormal
#define CELL_HI_UP 0x3FA99999  // high word of 0.05 (0x3FA999999999999A): the series are used below it

__device__ __forceinline__ double int_to_double_magic(int k) {  // exact int32 -> double with one FP64 add
  return __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0;  // 2^52 + 2^31
}
__device__ __forceinline__ double uint_to_double_magic(unsigned k) {  // exact uint32 -> double with one FP64 add
  return __hiloint2double(0x43300000, (int)k) - 4503599627370496.0;  // 2^52
}
// The same with the 2^52 bit pattern (0x4330000000000000) held in a 64-bit register: ONE integer instruction
// (IMAD.WIDE.U32 k * 1 + pattern) builds the double instead of a move per half, then the FP64 subtraction.
__device__ __forceinline__ double uint_to_double_wide(unsigned k, unsigned long long pattern) {
  unsigned long long bits;
  asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(bits) : "r"(k), "l"(pattern));
  return __longlong_as_double((long long)bits) - 4503599627370496.0;
}

// 1/x for a positive normal double inside the FP32 range (2^-126 <= x < 2^127), relative error < 2^-43
__device__ __forceinline__ double rcp_seeded(double x, int hi, int lo) {
  const unsigned fb = __funnelshift_l((unsigned)lo, (unsigned)(hi - 0x38000000), 3);  // float(x), mantissa truncated
  float rf;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(__uint_as_float(fb)));
  const unsigned rb = __float_as_uint(rf);
  const double rc = __hiloint2double((int)(rb >> 3) + 0x38000000, (int)(rb << 29));
  return fma(rc, fma(-x, rc, 1.0), rc);
}

// Inputs: yd = y, rd = S - y, X = I + psi W_t Bc (all exact / already formed), e = dt exp(a_t) pm_m, epsdt = eps dt.
// Outputs: val += term (VAL);  GRAD: gg_e = (d term / d x) * e, so that h = gg_e * X is the cell's d/d(log-rate) and
// gg_e * (W_t Bc) its d/d psi.
template <bool GRAD, bool VAL>
__device__ __forceinline__ void cell_eval(double yd, double rd, double X, double e, double epsdt, const double2* __restrict__ tab,
                                          const ll_coefs& K_, double& val, double& gg_e) {
  const double* K = K_.k;
  const double x = fma(e, X, epsdt);
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const double x2 = x * x;
  double term = 0.0, gg = 0.0;
  if (VAL) {
    const int ex = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 tc = tab[(hi >> 13) & 127];  // {1/c rounded, -log(1/c rounded)}
    const double r = fma(m, tc.x, -1.0);
    const double r2 = r * r;
    double p = fma(r, K[0], K[1]);
    p = fma(r, p, K[2]);
    p = fma(r, p, K[3]);
    p = fma(r, p, K[4]);
    p = fma(r, p, K[5]);
    const double lg = fma(int_to_double_magic(ex), K[6], tc.y) + fma(r2, p, r);
    term = fma(yd, lg + fma(x2, fma(x2, fma(x2, K[7], K[8]), K[9]), -0.5 * x), -rd * x);
  }
  if (GRAD) {
    const double rc = rcp_seeded(x, hi, lo);
    gg = fma(yd, (rc - 0.5) + x * fma(x2, fma(x2, K[10], K[11]), K[12]), -rd);
  }
  const bool fast = (unsigned)(hi - CELL_HI_LO) < (unsigned)(CELL_HI_UP - CELL_HI_LO);
  if (__builtin_expect(!fast, 0)) {
    const double em = expm1(-x);  // -(1-exp(-x)) = -p
    term = -rd * x;
    gg = -rd;
    if (yd > 0.0) {
      if (VAL) term += yd * log(-em);
      if (GRAD) gg += yd * (1.0 + em) / (-em);
    }
  }
  if (VAL) val += term;
  if (GRAD) gg_e = gg * e;
}

// ---- branch-free form: several cells of a thread in flight ------------------------------------------------------------------
// cell_eval ends in a (rare) branch to the library path, which closes the basic block: the compiler cannot interleave the
// dependent chains of consecutive cells across it, and the kernels ran at the latency of one cell after the other.
// cell_fast computes the fast path unconditionally (garbage, but no trap, outside its range) and reports whether the cell is
// inside the range; the caller evaluates a whole group of cells, then takes ONE branch for the group and repairs the
// out-of-range cells with cell_slow.
template <bool GRAD, bool VAL>
__device__ __forceinline__ bool cell_fast(double yd, double rd, double X, double e, double epsdt, const double2* __restrict__ tab,
                                          const ll_coefs& K_, double& term, double& gg_e) {
  const double* K = K_.k;
  const double x = fma(e, X, epsdt);
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const double x2 = x * x;
  if (VAL) {
    const int ex = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 tc = tab[(hi >> 13) & 127];  // {1/c rounded, -log(1/c rounded)}
    const double r = fma(m, tc.x, -1.0);
    const double r2 = r * r;
    double p = fma(r, K[0], K[1]);
    p = fma(r, p, K[2]);
    p = fma(r, p, K[3]);
    p = fma(r, p, K[4]);
    p = fma(r, p, K[5]);
    const double lg = fma(int_to_double_magic(ex), K[6], tc.y) + fma(r2, p, r);
    term = fma(yd, lg + fma(x2, fma(x2, fma(x2, K[7], K[8]), K[9]), -0.5 * x), -rd * x);
  }
  if (GRAD) {
    const double rc = rcp_seeded(x, hi, lo);
    gg_e = fma(yd, (rc - 0.5) + x * fma(x2, fma(x2, K[10], K[11]), K[12]), -rd) * e;
  }
  return (unsigned)(hi - CELL_HI_LO) < (unsigned)(CELL_HI_UP - CELL_HI_LO);
}

template <bool GRAD, bool VAL>
__device__ __forceinline__ void cell_slow(double yd, double rd, double X, double e, double epsdt, double& term, double& gg_e) {
  const double x = fma(e, X, epsdt);
  const double em = expm1(-x);  // -(1-exp(-x)) = -p
  double t = -rd * x, gg = -rd;
  if (yd > 0.0) {
    if (VAL) t += yd * log(-em);
    if (GRAD) gg += yd * (1.0 + em) / (-em);
  }
  if (VAL) term = t;
  if (GRAD) gg_e = gg * e;
}

// by-value form of the library path (a by-reference interface forces the caller's result arrays into local memory)
template <bool VAL>
__device__ __noinline__ double2 cell_slow_v(double yd, double rd, double X, double e, double epsdt) {
  double t = 0.0, g = 0.0;
  cell_slow<true, VAL>(yd, rd, X, e, epsdt, t, g);
  return make_double2(t, g);
}
