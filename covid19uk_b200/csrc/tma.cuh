// 1-D bulk asynchronous copies (cp.async.bulk -> SASS UBLKCP, the TMA unit without a tensor map) completing on
// shared-memory mbarriers: the staging primitive of the streaming kernels (loglik.cu, ingest.cu).
#pragma once
#include <stdint.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spins > (1u << 24)) __trap();  // a copy that never lands must fail loudly, not hang the device
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// One lane of a fully active warp (elect.sync).  Producers of bulk copies and issuers of tcgen05.mma must be chosen THIS way:
// behind `if (lane == 0)` the compiler treats the branch as divergent and wraps every uniform-datapath instruction
// (UBLKCP, UTCIMMA, ...) in an ELECT / BRA.U.ANY loop over the active lanes -- several extra instructions and a backward
// branch per copy / MMA (cuobjdump -sass; contract_i8.cu: 116 -> 100 cycles per MMA step).
__device__ __forceinline__ bool seir_elect_one() {
  unsigned p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
