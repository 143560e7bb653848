// Device helpers: mbarrier + 1-D TMA bulk copies (sm_100a), group reductions.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pcb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// global -> shared bulk copy through the TMA unit; completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
// Programmatic dependent launch (the sweep's kernel chain, csrc/pcb_api.cu): `pdl_trigger` lets the next kernel of the
// stream start its prologue while this grid still runs, `pdl_wait` holds a kernel until everything before it in the
// stream has completed and its writes are visible.  Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// reductions over the P consecutive lanes (P power of two <= 32) that share one tile row
__device__ __forceinline__ double group_max(double v, int P) {
  for (int off = P >> 1; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
__device__ __forceinline__ double group_sum(double v, int P) {
  for (int off = P >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
// exclusive prefix sum over the P lanes of a group; `lane_in_group` in [0,P)
__device__ __forceinline__ double group_excl_scan(double v, int P, int lane_in_group) {
  double incl = v;
  for (int off = 1; off < P; off <<= 1) {
    double o = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane_in_group >= off) incl += o;
  }
  // shift rather than subtract: `incl - v` cancels catastrophically when v dwarfs the prefix
  const double excl = __shfl_up_sync(0xffffffffu, incl, 1);
  return lane_in_group == 0 ? 0.0 : excl;
}

// (m, s) pairs representing log(sum) = m + log(s); combine is associative
struct LsePair {
  double m, s;
};
__device__ __forceinline__ LsePair lse_combine(LsePair a, LsePair b) {
  // a is the earlier prefix, b the later one
  if (a.m == -INFINITY) return b;
  if (b.m == -INFINITY) return a;
  LsePair r;
  if (a.m >= b.m) {
    r.m = a.m;
    r.s = a.s + b.s * exp(b.m - a.m);
  } else {
    r.m = b.m;
    r.s = a.s * exp(a.m - b.m) + b.s;
  }
  return r;
}

}  // namespace pcb
