// Shared-memory staging for the HBM-bound row kernels (add+LayerNorm, combine): a producer thread streams chunks of
// rows into a ring of shared-memory stages with bulk asynchronous copies (cp.async.bulk, completion counted on an
// mbarrier) while the consumer warps work out of shared memory.  The bytes in flight per SM are then
// (stages - 1) x stage size regardless of how long the arithmetic of a row takes — with one-warp-per-row register
// kernels the loads of a warp only overlap the arithmetic of OTHER warps, and at ~130 registers per thread there are
// too few of those to cover the DRAM latency (ncu, profiles/r02e: 23 % warps active, 18 % of DRAM peak).
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace b200 {

// 1-D bulk copy global -> shared; dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Ring bookkeeping shared by producer and consumers (both walk the same chunk sequence).
struct RingPos {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int stages) {
    if (++stage == stages) { stage = 0; phase ^= 1u; }
  }
};

}  // namespace b200
