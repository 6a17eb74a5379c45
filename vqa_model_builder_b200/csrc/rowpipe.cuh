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

// keep-bits of the 8 elements 8*idx8 .. 8*idx8+7 (bit q set = element kept); same stream as drop_scales8
__device__ __forceinline__ uint32_t drop_keep8(const DropState& d, unsigned long long idx8) {
  const uint4 r = philox4x32(d.key, idx8, d.site);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m |= ((w[i] & 0xFFFFu) >= d.thr ? 1u : 0u) << (2 * i);
    m |= ((w[i] >> 16) >= d.thr ? 1u : 0u) << (2 * i + 1);
  }
  return m;
}

__device__ __forceinline__ void unpack8(const uint4& t, float (&v)[8]) {
  const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 t;
  t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
  t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
  return t;
}

}  // namespace b200
