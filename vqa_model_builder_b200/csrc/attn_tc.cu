// Fused attention core on the 5th-gen tensor cores (bf16 in, fp32 softmax/accumulate), forward and backward.
// One CTA per (batch, head, 128-row query tile): the query rows form the M=128 dimension of every UMMA; keys/values
// are walked in tiles of 128 (S <= 384 forward; any T).  Q/K/V/dO tiles arrive by TMA into 128B-swizzled shared memory; S = Q K^T, dP = dO V^T,
// O/dQ/dK/dV accumulate in TMEM; the 128 threads (thread = TMEM lane = matrix row) do the softmax math in
// registers and hand P / dS back to the tensor core through shared memory written in the same swizzled layout.
// The [B,H,T,S] score tensor never exists in HBM.
//
// Layout trick used throughout: a [rows x 64-column chunk] tile with 128-byte rows and the 128B XOR swizzle is at
// the same time a K-major operand (M/N = rows, K = columns) and an MN-major operand (K = rows, M/N = columns), so
// one copy of P (or dS, Q, K, dO) serves both A*B and A^T*B products.
#include <cuda.h>

#include "gemm_common.cuh"
#include "tc_ptx.cuh"

namespace b200 {

namespace {

constexpr int ROWS = 128;              // query rows per CTA == kv rows per tile == TMA box rows
constexpr int CHB = ROWS * 128;        // bytes of one [128 rows x 64 bf16] chunk
constexpr float LOG2E = 1.4426950408889634f;

struct AttnTcArgs {
  int B, H, T, S, dh;
  float scale;
  const uint8_t* key_pad;   // [B,S] 1 = ignore, or null
  int causal;               // 1: key s is visible to query t only if s <= t (decoder self-attention)
  bf16* o; int ldo;         // fwd out
  float* lse;               // [B,H,T]
  // backward
  const bf16* o_in; int ldo_in;
  const bf16* d_o; int lddo;
  bf16* dq; int lddq;
  bf16* dk; int lddk;
  bf16* dv; int lddv;
  // attention-probability dropout; element index = ((b*H+h)*T + t) * (ntiles*128) + s
  const unsigned long long* drop_state; float drop_p; unsigned int drop_site;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16-byte unit `u` (0..7) of row `r` inside a swizzled [128 x 64] chunk
__device__ __forceinline__ uint32_t swz(int r, int u) { return (uint32_t)(r * 128 + ((u ^ (r & 7)) << 4)); }

// store 32 consecutive columns (32-col block `c32` of a 128-col tile) of row r as bf16 into a swizzled tile buffer
__device__ __forceinline__ void store_row32(uint8_t* tile, int r, int c32, const float (&v)[32]) {
  uint8_t* chunk = tile + (c32 >> 1) * CHB;
  const int u0 = (c32 & 1) * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 t;
    t.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]); t.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
    t.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); t.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(chunk + swz(r, u0 + j)) = t;
  }
}

__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  ptx::tmem_ld_32x32(taddr, r);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ uint32_t idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// K-major operand: k-step kk covers columns 16kk..16kk+15 -> chunk kk/4, 32-byte step inside the 128-byte row
__device__ __forceinline__ uint64_t desc_k(uint32_t tile, int kk, int ch = CHB) {
  return ptx::umma_smem_desc(tile + (kk >> 2) * ch + (kk & 3) * 32, 16, 1024);
}
// MN-major operand: k-step kk covers rows 16kk..16kk+15 (2048 bytes); 64-wide MN chunks are CHB apart
__device__ __forceinline__ uint64_t desc_mn(uint32_t tile, int kk, int ch = CHB) {
  return ptx::umma_smem_desc(tile + kk * 2048, ch, 1024);
}

// Key visibility as bit masks: the per-element "col < n_valid && !key_pad[col] && !(causal && col > row)" test was a
// byte load, two compares and a branch per score element (ncu, profiles/r02l: 40 % of the forward kernel's
// instructions were ISETP / BRA / BSSY / BSYNC / LDG).  The CTA builds one 32-bit word per 32 key columns once
// (valid and not padded); the causal part is a per-row word computed from the row index.
__device__ __forceinline__ void build_key_mask(uint32_t* words, const uint8_t* kp, int S, int cols, int tid, int nthr) {
  for (int col = tid; col < cols; col += nthr) {     // cols and nthr are multiples of 32: whole warps iterate together
    const bool ok = col < S && !(kp != nullptr && kp[col] != 0);
    const uint32_t w = __ballot_sync(0xffffffffu, ok);
    if ((tid & 31) == 0) words[col >> 5] = w;
  }
}
__device__ __forceinline__ uint32_t causal_word(int causal, int row, int col0) {
  if (!causal) return 0xffffffffu;
  const int d = row - col0;                          // columns col0 .. col0+31 are visible while col <= row
  return d >= 31 ? 0xffffffffu : (d < 0 ? 0u : ((2u << d) - 1u));
}

// Geometry of the two kernel flavours.  R = 128: the general kernel (query tiles and kv tiles of 128 rows, 128 threads).
// R = 64: T <= 64 and S <= 64 (the question / image-patch shapes of the fusion block): TMA boxes of 64 rows, 64
// threads, half the shared memory and a quarter of the TMEM, so 3 (forward) / 2 (backward) CTAs share an SM and
// hide each other's TMA -> MMA -> softmax -> MMA latency chain.  The UMMAs keep M = 128: accumulator rows 64..127
// are computed from whatever follows the 64-row operand in shared memory and are never read back.
// P may take Q's shared-memory buffer when there is a single kv tile and Q is at least as large as a P tile
template <int NT, int R>
__host__ __device__ constexpr bool fwd_alias_possible() { return NT == 1 && R == 128; }
template <int NT, int R>
__host__ __device__ inline bool fwd_alias(int dh) { return fwd_alias_possible<NT, R>() && dh > 64; }

template <int R> struct BwdGeo {
  static constexpr int NTHR = (R == 128) ? 256 : 64;     // threads of the backward kernel: two per row for R = 128
};
template <int R> struct Geo {
  static constexpr int CH = R * 128;                 // bytes of one TMA-loaded [R rows x 64 bf16] chunk
  static constexpr int PT = (R == 64) ? 8192 : 2 * CHB;   // bytes of a P / dS tile ([R x 64] or [128 x 128])
  static constexpr int SLACK = (R == 64) ? 8192 : 0;      // readable bytes behind the last operand (M = 128 over-read)
};

struct Smem {
  uint32_t base;     // 1024-aligned shared address
  uint8_t* ptr;
};
__device__ __forceinline__ Smem align_smem(uint8_t* raw) {
  const uint32_t r = ptx::smem_u32(raw);
  const uint32_t b = (r + 1023u) & ~1023u;
  return Smem{b, raw + (b - r)};
}

// ============================================ forward ==============================================
// smem: Q[nch] | K/V[nch] (one buffer: every K tile is consumed by its S = Q K^T product before the first V tile is
// loaded, and the V loads are issued behind the MMA that frees the buffer) | P[2] chunks, then barriers.  The shared
// buffer is what lets two 128-row CTAs (four 64-row ones) share an SM and hide each other's TMA -> MMA -> softmax ->
// MMA chain.  TMEM: S tiles at columns 128*j, O at 128*NT.
template <int NT, int R>
__global__ void __launch_bounds__(R)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                   const __grid_constant__ CUtensorMap vmap, const AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const Smem sm = align_smem(smem_raw);
  const int nch = (a.dh + 63) / 64;
  constexpr int CH = Geo<R>::CH;
  constexpr bool SMALL = (R == 64);
  static_assert(!SMALL || NT == 1, "the 64-row flavour handles a single kv tile");
  // Single-kv-tile 128-row flavour: Q is dead once S = Q K^T has completed and S is dead once P sits in shared
  // memory, so P takes Q's buffer and O takes S's TMEM columns: 64 KB of shared memory and 128 TMEM columns per CTA,
  // three CTAs per SM hide each other's TMA -> MMA -> softmax -> MMA chain (two before).
  const bool alias = fwd_alias<NT, R>(a.dh);
  const uint32_t sQ = sm.base, sK = sQ + nch * CH, sV = sK, sP = alias ? sQ : sK + nch * CH;
  uint8_t* pP = alias ? sm.ptr : sm.ptr + 2 * nch * CH;
  const int tail = 2 * nch * CH + (alias ? 0 : Geo<R>::PT + Geo<R>::SLACK);
  const uint32_t bars = sm.base + tail;
  const uint32_t bar_q = bars, bar_k = bars + 8, bar_v = bars + 16, bar_mma = bars + 24;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm.ptr + tail + 32);
  uint32_t* s_ok = reinterpret_cast<uint32_t*>(sm.ptr + tail + 64);   // [NT*R/32]
  // 64-row flavour and the aliased 128-row one: O overwrites S -> 128 columns
  constexpr uint32_t TCOLS = (SMALL || NT == 1) ? 128 : 512;
  constexpr uint32_t O_COL = (SMALL || NT == 1) ? 0 : NT * 128;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int q0 = blockIdx.y * R;         // first query row of this CTA's tile (blockIdx.y > 0 only when T > 128)
  const int ksteps_d = a.dh / 16;

  if (tid == 0) {
    ptx::prefetch_tensormap(&qmap);
    ptx::prefetch_tensormap(&kmap);
    ptx::prefetch_tensormap(&vmap);
    ptx::mbar_init(bar_q, 1);
    ptx::mbar_init(bar_k, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), TCOLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_trigger();   // setup above overlapped the previous kernel's tail; global traffic starts below
  pdl_wait();
  uint32_t ph_k = 0, ph_v = 0, ph_mma = 0;

  // ---- S_j = Q K_j^T for every kv tile (K buffer reused serially; the tensor-core work here is tiny) ----
  if (tid == 0) {
    ptx::mbar_arrive_expect_tx(bar_q, nch * CH);
    for (int c = 0; c < nch; ++c) ptx::tma_load_2d(sQ + c * CH, &qmap, bar_q, h * a.dh + 64 * c, b * a.T + q0);
  }
  for (int j = 0; j < NT; ++j) {
    const int n_valid = min(R, a.S - j * R);
    if (n_valid <= 0) break;
    const int n16 = (n_valid + 15) & ~15;
    if (tid == 0) {
      ptx::mbar_arrive_expect_tx(bar_k, nch * CH);
      for (int c = 0; c < nch; ++c)
        ptx::tma_load_2d(sK + c * CH, &kmap, bar_k, h * a.dh + 64 * c, b * a.S + j * R);
      if (j == 0) ptx::mbar_wait(bar_q, 0);
      ptx::mbar_wait(bar_k, ph_k);
      ptx::tc_fence_after();
      const uint32_t id = idesc(128, n16, false, false);
      for (int kk = 0; kk < ksteps_d; ++kk)
        ptx::umma_bf16(tmem + j * 128, desc_k(sQ, kk, CH), desc_k(sK, kk, CH), id, kk > 0 ? 1u : 0u);
      ptx::umma_commit(bar_mma);
    }
    ph_k ^= 1;
    ptx::mbar_wait(bar_mma, ph_mma);   // every thread: S_j complete and the K buffer is free again
    ph_mma ^= 1;
  }
  ptx::tc_fence_after();
  if (tid == 0) {   // prefetch V_0 while the softmax statistics are computed
    ptx::mbar_arrive_expect_tx(bar_v, nch * CH);
    for (int c = 0; c < nch; ++c) ptx::tma_load_2d(sV + c * CH, &vmap, bar_v, h * a.dh + 64 * c, b * a.S);
  }

  // ---- softmax: thread = query row -------------------------------------------------------------------
  const int r = tid;
  const int rg = q0 + r;                 // query row inside the (batch, head) problem
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const uint8_t* kp = a.key_pad ? a.key_pad + (long long)b * a.S : nullptr;
  build_key_mask(s_ok, kp, a.S, NT * R, tid, R);
  __syncthreads();
  float m = -INFINITY;
  for (int j = 0; j < NT; ++j) {
    const int n_valid = min(R, a.S - j * R);
    if (n_valid <= 0) break;
    for (int c = 0; c * 32 < n_valid; ++c) {
      float v[32];
      ld32(lane_base + j * 128 + c * 32, v);
      const uint32_t okb = s_ok[j * (R / 32) + c] & causal_word(a.causal, rg, j * R + c * 32);
#pragma unroll
      for (int i = 0; i < 32; ++i) m = fmaxf(m, (okb >> i) & 1u ? v[i] : -INFINITY);
    }
  }
  const float sl2 = a.scale * LOG2E;
  const float m_s = (m == -INFINITY) ? 0.f : m * sl2;
  const DropState ds = drop_load(a.drop_state, a.drop_p, a.drop_site);
  const unsigned long long drow = ((unsigned long long)blockIdx.x * a.T + rg) * (unsigned long long)(((a.S + ROWS - 1) / ROWS) * ROWS);
  float l = 0.f;
  for (int j = 0; j < NT; ++j) {
    const int n_valid = min(R, a.S - j * R);
    if (n_valid <= 0) break;
    const int n16 = (n_valid + 15) & ~15;
    for (int c = 0; c * 32 < n16; ++c) {
      float v[32];
      ld32(lane_base + j * 128 + c * 32, v);
      const uint32_t okb = s_ok[j * (R / 32) + c] & causal_word(a.causal, rg, j * R + c * 32);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float p = (okb >> i) & 1u ? ex2(fmaf(v[i], sl2, -m_s)) : 0.f;
        // round to bf16 first so the normaliser matches the probabilities the tensor core actually sees
        const float pr = __bfloat162float(__float2bfloat16_rn(p));
        l += pr;
        v[i] = pr;
      }
      if (ds.on) {   // dropped probabilities feed P.V; the normaliser keeps the undropped sum
#pragma unroll
        for (int i8 = 0; i8 < 32; i8 += 8) {
          float sc[8];
          drop_scales8(ds, (drow + (unsigned long long)(j * ROWS + c * 32 + i8)) >> 3, sc);
#pragma unroll
          for (int q = 0; q < 8; ++q) v[i8 + q] *= sc[q];
        }
      }
      store_row32(pP, r, c, v);
    }
    fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      ptx::mbar_wait(bar_v, ph_v);
      const uint32_t id = idesc(128, a.dh, false, true);
      for (int kk = 0; kk < n16 / 16; ++kk)
        ptx::umma_bf16(tmem + O_COL, desc_k(sP, kk), desc_mn(sV, kk, CH), id, (j > 0 || kk > 0) ? 1u : 0u);
      ptx::umma_commit(bar_mma);
    }
    ph_v ^= 1;
    ptx::mbar_wait(bar_mma, ph_mma);   // P and V buffers are free again
    ph_mma ^= 1;
    if (tid == 0 && j + 1 < NT && a.S - (j + 1) * R > 0) {
      ptx::mbar_arrive_expect_tx(bar_v, nch * CH);
      for (int c = 0; c < nch; ++c)
        ptx::tma_load_2d(sV + c * CH, &vmap, bar_v, h * a.dh + 64 * c, b * a.S + (j + 1) * R);
    }
  }
  ptx::tc_fence_after();

  // ---- epilogue: O / l -> bf16 -> global; log-sum-exp for backward -----------------------------------------
  const float inv = l > 0.f ? 1.f / l : 0.f;
  {
    bf16* orow = a.o + ((long long)b * a.T + rg) * a.ldo + h * a.dh;
    for (int c = 0; c < a.dh / 32; ++c) {
      float v[32];
      ld32(lane_base + O_COL + c * 32, v);   // .aligned: executed by the whole warp, stores are predicated
      if (rg < a.T) {
#pragma unroll
        for (int jv = 0; jv < 4; ++jv) {
          uint4 t;
          t.x = pack_bf16x2(v[8 * jv + 0] * inv, v[8 * jv + 1] * inv); t.y = pack_bf16x2(v[8 * jv + 2] * inv, v[8 * jv + 3] * inv);
          t.z = pack_bf16x2(v[8 * jv + 4] * inv, v[8 * jv + 5] * inv); t.w = pack_bf16x2(v[8 * jv + 6] * inv, v[8 * jv + 7] * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + jv * 8) = t;
        }
      }
    }
    if (rg < a.T) a.lse[((long long)b * a.H + h) * a.T + rg] = (l > 0.f) ? m * a.scale + __logf(l) : -INFINITY;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, TCOLS);
}

// ============================================ backward =============================================
// smem: Q[nch] | dO[nch] | K[nch] | V[nch] | P | dS.  TMEM (R = 128): S at 0, dP at 128, dQ at 256; dV / dK take the
// S / dP columns once P / dS sit in shared memory.  R = 64: S at 0, dP at 64; dV at 0 and dK at 128 are drained first,
// then dQ reuses column 0 (256 columns in all, two CTAs per SM).
//
// More than 128 query rows (vision-token queries of ViT-B/16 / DINOv2, the single-stream fusion's 300+ tokens): the
// grid grows a second dimension and every CTA owns ONE output tile completely, so nothing is accumulated across CTAs
// (no atomics, deterministic):
//   role Q  (blockIdx.y <  q_tiles): query tile i fixed, key tiles streamed  -> dQ_i   (S, dP, dS_ij, dQ_i += dS_ij K_j)
//   role KV (blockIdx.y >= q_tiles): key tile j fixed, query tiles streamed  -> dV_j, dK_j accumulate in TMEM columns
//                                    256 / 384 over the query tiles (S / dP columns are rewritten every step)
// S and dP are computed by both roles; with T <= 128 there is one CTA per (batch, head) doing everything (role ALL).
enum { ROLE_ALL = 0, ROLE_Q = 1, ROLE_KV = 2 };

template <int R>
__global__ void __launch_bounds__(BwdGeo<R>::NTHR, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                   const __grid_constant__ CUtensorMap vmap, const __grid_constant__ CUtensorMap domap,
                   const AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const Smem sm = align_smem(smem_raw);
  const int nch = (a.dh + 63) / 64;
  constexpr int CH = Geo<R>::CH, PT = Geo<R>::PT;
  constexpr bool SMALL = (R == 64);
  constexpr int PCH = SMALL ? PT : CHB;    // pitch between the 64-column chunks of the P / dS tiles
  const uint32_t sQ = sm.base, sdO = sQ + nch * CH, sK = sdO + nch * CH, sV = sK + nch * CH;
  const uint32_t sP = sV + nch * CH, sdS = sP + PT;
  uint8_t* pP = sm.ptr + 4 * nch * CH;
  uint8_t* pdS = pP + PT;
  const uint32_t bars = sdS + PT + Geo<R>::SLACK;
  const uint32_t bar_q = bars, bar_kv = bars + 8, bar_mma = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm.ptr + 4 * nch * CH + 2 * PT + Geo<R>::SLACK + 32);
  uint32_t* s_ok = reinterpret_cast<uint32_t*>(sm.ptr + 4 * nch * CH + 2 * PT + Geo<R>::SLACK + 64);   // [ntiles*R/32]
  constexpr uint32_t TCOLS = SMALL ? 256 : 512, C_S = 0, C_DP = SMALL ? 64 : 128, C_DQ = SMALL ? 0 : 256;

  // 128-row flavour: TWO threads per row (256 threads).  One CTA per SM is all the 192 KB of operand tiles allow, and
  // with one warp per scheduler the softmax / dS arithmetic and the TMEM drains were a serial chain per row; the
  // second half-CTA takes every other 32-column chunk of S / dP, the other one of the dV / dK drains and part of dQ.
  constexpr int NTHR = BwdGeo<R>::NTHR, HALVES = NTHR / R;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int half = tid / R;                 // warp-uniform
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int ksteps_d = a.dh / 16;
  const int ntiles = (a.S + R - 1) / R;
  const int q_tiles = (a.T + R - 1) / R;
  const int role = q_tiles == 1 ? ROLE_ALL : ((int)blockIdx.y < q_tiles ? ROLE_Q : ROLE_KV);
  const int i_fixed = role == ROLE_Q ? (int)blockIdx.y : 0;
  const int j_fixed = role == ROLE_KV ? (int)blockIdx.y - q_tiles : 0;
  const int nsteps = role == ROLE_KV ? q_tiles : ntiles;
  const uint32_t C_DV = role == ROLE_KV ? 256u : 0u, C_DK = role == ROLE_KV ? 384u : 128u;

  if (tid == 0) {
    ptx::prefetch_tensormap(&qmap);
    ptx::prefetch_tensormap(&kmap);
    ptx::prefetch_tensormap(&vmap);
    ptx::prefetch_tensormap(&domap);
    ptx::mbar_init(bar_q, 1);
    ptx::mbar_init(bar_kv, 1);
    ptx::mbar_init(bar_mma, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), TCOLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_trigger();
  pdl_wait();
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // a warp reads TMEM lanes 32 (w % 4) ..
  uint32_t ph_q = 0, ph_kv = 0, ph_mma = 0;

  const int r = tid % R;
  const float sl2 = a.scale * LOG2E;
  const uint8_t* kp = a.key_pad ? a.key_pad + (long long)b * a.S : nullptr;
  build_key_mask(s_ok, kp, a.S, ntiles * R, tid, NTHR);
  __syncthreads();
  const DropState ds = drop_load(a.drop_state, a.drop_p, a.drop_site);
  const unsigned long long spad = (unsigned long long)(((a.S + ROWS - 1) / ROWS) * ROWS);
  // per-row state of the current query tile
  float delta = 0.f, lse_l2 = 0.f;
  bool row_ok = false;
  int rg = r;
  unsigned long long drow = 0;

  for (int st = 0; st < nsteps; ++st) {
    const int i = role == ROLE_KV ? st : i_fixed;        // query tile
    const int j = role == ROLE_KV ? j_fixed : st;        // key tile
    const bool load_q = st == 0 || role == ROLE_KV;
    const bool load_kv = st == 0 || role != ROLE_KV;
    const int q0 = i * R;
    const int ksteps_t = (min(a.T - q0, R) + 15) / 16;
    const int n_valid = min(R, a.S - j * R);
    const int n16 = (n_valid + 15) & ~15;
    if (tid == 0) {
      if (load_q) {
        ptx::mbar_arrive_expect_tx(bar_q, 2 * nch * CH);
        for (int c = 0; c < nch; ++c) {
          ptx::tma_load_2d(sQ + c * CH, &qmap, bar_q, h * a.dh + 64 * c, b * a.T + q0);
          ptx::tma_load_2d(sdO + c * CH, &domap, bar_q, h * a.dh + 64 * c, b * a.T + q0);
        }
      }
      if (load_kv) {
        ptx::mbar_arrive_expect_tx(bar_kv, 2 * nch * CH);
        for (int c = 0; c < nch; ++c) {
          ptx::tma_load_2d(sK + c * CH, &kmap, bar_kv, h * a.dh + 64 * c, b * a.S + j * R);
          ptx::tma_load_2d(sV + c * CH, &vmap, bar_kv, h * a.dh + 64 * c, b * a.S + j * R);
        }
      }
    }
    if (load_q) {
      // per-row statistics of this query tile: delta = sum_d dO*O, lse
      rg = q0 + r;
      row_ok = rg < a.T;
      delta = 0.f;
      lse_l2 = 0.f;
      if (row_ok) {
        const bf16* orow = a.o_in + ((long long)b * a.T + rg) * a.ldo_in + h * a.dh;
        const bf16* grow = a.d_o + ((long long)b * a.T + rg) * a.lddo + h * a.dh;
        const int dsplit = (a.dh / 8 + HALVES - 1) / HALVES * 8;          // each half-CTA sums its share of the head dim
        for (int d = half * dsplit; d < min(a.dh, (half + 1) * dsplit); d += 8) {
          Vec16<bf16> ov, gv;
          ov.load(orow + d);
          gv.load(grow + d);
#pragma unroll
          for (int u = 0; u < 8; ++u) delta = fmaf(ov.v[u], gv.v[u], delta);
        }
        lse_l2 = a.lse[((long long)b * a.H + h) * a.T + rg] * LOG2E;
      }
      drow = ((unsigned long long)blockIdx.x * a.T + rg) * spad;
      if (HALVES == 2) {                                                  // the P buffer is free between steps
        reinterpret_cast<float*>(pP)[half * R + r] = delta;
        __syncthreads();
        delta = reinterpret_cast<float*>(pP)[r] + reinterpret_cast<float*>(pP)[R + r];
        __syncthreads();                                                  // before P is written
      }
    }
    if (tid == 0) {
      if (load_q) ptx::mbar_wait(bar_q, ph_q);
      if (load_kv) ptx::mbar_wait(bar_kv, ph_kv);
      ptx::tc_fence_after();
      const uint32_t id = idesc(128, n16, false, false);
      for (int kk = 0; kk < ksteps_d; ++kk)   // S = Q K^T
        ptx::umma_bf16(tmem + C_S, desc_k(sQ, kk, CH), desc_k(sK, kk, CH), id, kk > 0 ? 1u : 0u);
      for (int kk = 0; kk < ksteps_d; ++kk)   // dP = dO V^T
        ptx::umma_bf16(tmem + C_DP, desc_k(sdO, kk, CH), desc_k(sV, kk, CH), id, kk > 0 ? 1u : 0u);
      ptx::umma_commit(bar_mma);
    }
    if (load_q) ph_q ^= 1;
    if (load_kv) ph_kv ^= 1;
    ptx::mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    ptx::tc_fence_after();

    // P = exp(S*scale - lse), dS = P * (dP - delta) * scale  -> swizzled smem tiles (bf16)
    for (int c = half; c < R / 32; c += HALVES) {
      float s[32], dp[32];
      if (c * 32 < n16) {
        ld32(lane_base + C_S + c * 32, s);
        ld32(lane_base + C_DP + c * 32, dp);
      }
      float keep[32];
#pragma unroll
      for (int i8 = 0; i8 < 32; i8 += 8) {
        float sc[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        if (ds.on) drop_scales8(ds, (drow + (unsigned long long)(j * ROWS + c * 32 + i8)) >> 3, sc);
#pragma unroll
        for (int q = 0; q < 8; ++q) keep[i8 + q] = sc[q];
      }
      const uint32_t okb = row_ok ? (s_ok[j * (R / 32) + c] & causal_word(a.causal, rg, j * R + c * 32)) : 0u;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const bool ok = (okb >> e) & 1u;
        const float p = ok ? ex2(fmaf(s[e], sl2, -lse_l2)) : 0.f;
        s[e] = p * keep[e];                                               // dropped P feeds dV = P^T dO
        dp[e] = ok ? p * (dp[e] * keep[e] - delta) * a.scale : 0.f;       // dS = P (dP*mask/(1-p) - delta)
      }
      // columns >= n16 are written as zeros too: the dV/dK products read all 128 columns of these tiles (M dim)
      store_row32(pP, r, c, s);
      store_row32(pdS, r, c, dp);
    }
    fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      ptx::tc_fence_after();
      if (role != ROLE_Q) {
        const uint32_t acc0 = (role == ROLE_KV && st > 0) ? 1u : 0u;   // role KV: sums over the query tiles
        const uint32_t id_t = idesc(128, a.dh, true, true);    // A^T B with A = P / dS (MN-major), B = dO / Q (MN-major)
        for (int kk = 0; kk < ksteps_t; ++kk)   // dV = P^T dO
          ptx::umma_bf16(tmem + C_DV, desc_mn(sP, kk, PCH), desc_mn(sdO, kk, CH), id_t, kk > 0 ? 1u : acc0);
        for (int kk = 0; kk < ksteps_t; ++kk)   // dK = dS^T Q
          ptx::umma_bf16(tmem + C_DK, desc_mn(sdS, kk, PCH), desc_mn(sQ, kk, CH), id_t, kk > 0 ? 1u : acc0);
      }
      if (!SMALL && role != ROLE_KV) {
        const uint32_t id_q = idesc(128, a.dh, false, true);   // dQ += dS K  (A = dS K-major, B = K MN-major)
        for (int kk = 0; kk < n16 / 16; ++kk)
          ptx::umma_bf16(tmem + C_DQ, desc_k(sdS, kk, PCH), desc_mn(sK, kk, CH), id_q, (st > 0 || kk > 0) ? 1u : 0u);
      }
      ptx::umma_commit(bar_mma);
    }
    ptx::mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    ptx::tc_fence_after();
    // dV, dK rows of this key tile: thread = kv row (role ALL: every step; role KV: once, after the last query tile)
    if (role == ROLE_ALL || (role == ROLE_KV && st == nsteps - 1)) {
      const int srow = j * R + r;
      const bool ok = r < n_valid;
      bf16* dvrow = a.dv + ((long long)b * a.S + srow) * a.lddv + h * a.dh;
      bf16* dkrow = a.dk + ((long long)b * a.S + srow) * a.lddk + h * a.dh;
      for (int c = 0; c < a.dh / 32; ++c) {
#pragma unroll
        for (int which = 0; which < 2; ++which) {         // 0: dV, 1: dK; with two half-CTAs each drains one of them
          if (HALVES == 2 && which != half) continue;
          float v[32];
          ld32(lane_base + (which == 0 ? C_DV : C_DK) + c * 32, v);
          bf16* orow = which == 0 ? dvrow : dkrow;
          if (ok) {
#pragma unroll
            for (int jv = 0; jv < 4; ++jv) {
              uint4 t;
              t.x = pack_bf16x2(v[8 * jv + 0], v[8 * jv + 1]); t.y = pack_bf16x2(v[8 * jv + 2], v[8 * jv + 3]);
              t.z = pack_bf16x2(v[8 * jv + 4], v[8 * jv + 5]); t.w = pack_bf16x2(v[8 * jv + 6], v[8 * jv + 7]);
              *reinterpret_cast<uint4*>(orow + c * 32 + jv * 8) = t;
            }
          }
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();   // TMEM S/dP regions and the Q/dO/K/V/P/dS buffers may be overwritten by the next step
  }

  if (SMALL) {   // dV / dK are drained: dQ = dS K takes their columns (single kv tile in this flavour)
    if (tid == 0) {
      ptx::tc_fence_after();
      const int n16 = (min(R, a.S) + 15) & ~15;
      const uint32_t id_q = idesc(128, a.dh, false, true);
      for (int kk = 0; kk < n16 / 16; ++kk)
        ptx::umma_bf16(tmem + C_DQ, desc_k(sdS, kk, PCH), desc_mn(sK, kk, CH), id_q, kk > 0 ? 1u : 0u);
      ptx::umma_commit(bar_mma);
    }
    ptx::mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
  }
  // dQ rows (rg / row_ok still describe the fixed query tile in roles ALL and Q)
  ptx::tc_fence_after();
  if (role != ROLE_KV) {
    bf16* dqrow = a.dq + ((long long)b * a.T + rg) * a.lddq + h * a.dh;
    for (int c = half; c < a.dh / 32; c += HALVES) {
      float v[32];
      ld32(lane_base + C_DQ + c * 32, v);
      if (row_ok) {
#pragma unroll
        for (int jv = 0; jv < 4; ++jv) {
          uint4 t;
          t.x = pack_bf16x2(v[8 * jv + 0], v[8 * jv + 1]); t.y = pack_bf16x2(v[8 * jv + 2], v[8 * jv + 3]);
          t.z = pack_bf16x2(v[8 * jv + 4], v[8 * jv + 5]); t.w = pack_bf16x2(v[8 * jv + 6], v[8 * jv + 7]);
          *reinterpret_cast<uint4*>(dqrow + c * 32 + jv * 8) = t;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, TCOLS);
}

template <int R> size_t fwd_smem_aliased(int dh) {     // P in Q's buffer (single kv tile, see fwd_alias)
  return (size_t)2 * ((dh + 63) / 64) * Geo<R>::CH + 1024 + 128;
}
template <int R> size_t fwd_smem(int dh) {
  return (size_t)2 * ((dh + 63) / 64) * Geo<R>::CH + Geo<R>::PT + Geo<R>::SLACK + 1024 + 128;
}
template <int R> size_t bwd_smem(int dh) {
  return (size_t)4 * ((dh + 63) / 64) * Geo<R>::CH + 2 * Geo<R>::PT + Geo<R>::SLACK + 1024 + 128;
}
bool small_shape(int T, int S) { return T <= 64 && S <= 64; }

}  // namespace

bool attn_tc_supported(int T, int S, int dh, int ldq, int ldk, int ldv, const void* q, const void* k, const void* v) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return S <= 3 * ROWS && dh % 32 == 0 && dh <= 128 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 &&
         al(q) && al(k) && al(v);
}

int launch_attn_fwd_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const uint8_t* key_pad,
                       int causal, void* o, int ldo, float* lse, int B, int H, int T, int S, int dh, float scale,
                       const unsigned long long* drop_state, float drop_p, unsigned int drop_site, cudaStream_t stream) {
  CUtensorMap qm, km, vm;
  const long long cols = (long long)H * dh;
  const bool small = small_shape(T, S);
  const int box = small ? 64 : ROWS;
  if (int rc = make_tma_map_bf16(&qm, q, cols, (long long)B * T, ldq, box)) return rc;
  if (int rc = make_tma_map_bf16(&km, k, cols, (long long)B * S, ldk, box)) return rc;
  if (int rc = make_tma_map_bf16(&vm, v, cols, (long long)B * S, ldv, box)) return rc;
  AttnTcArgs a{};
  a.B = B; a.H = H; a.T = T; a.S = S; a.dh = dh; a.scale = scale; a.key_pad = key_pad; a.causal = causal;
  a.o = (bf16*)o; a.ldo = ldo; a.lse = lse;
  a.drop_state = drop_state; a.drop_p = drop_p; a.drop_site = drop_site;
  const int nt = (S + ROWS - 1) / ROWS;
  static bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<128>(128)));
    B200_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<3, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<128>(128)));
    B200_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<64>(128)));
    B200_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<1, 64>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    B200_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<1, 128>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = true;
  }
  const int q_tiles = (T + ROWS - 1) / ROWS;     // one CTA per (batch, head, 128-row query tile)
  if (small) launch_kernel(attn_fwd_tc_kernel<1, 64>, dim3(B * H), dim3(64), fwd_smem<64>(dh), stream, qm, km, vm, a);
  else if (nt == 1)
    launch_kernel(attn_fwd_tc_kernel<1, 128>, dim3(B * H, q_tiles), dim3(128),
                  fwd_alias<1, 128>(dh) ? fwd_smem_aliased<128>(dh) : fwd_smem<128>(dh), stream, qm, km, vm, a);
  else launch_kernel(attn_fwd_tc_kernel<3, 128>, dim3(B * H, q_tiles), dim3(128), fwd_smem<128>(dh), stream, qm, km, vm, a);
  B200_LAUNCH_CHECK("attn_fwd_tc_kernel");
  count_launch();
  return 0;
}

int launch_attn_bwd_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const uint8_t* key_pad,
                       int causal, const void* o, int ldo, const void* d_o, int lddo, const float* lse, void* dq, int lddq, void* dk,
                       int lddk, void* dv, int lddv, int B, int H, int T, int S, int dh, float scale,
                       const unsigned long long* drop_state, float drop_p, unsigned int drop_site, cudaStream_t stream) {
  CUtensorMap qm, km, vm, dom;
  const long long cols = (long long)H * dh;
  const bool small = small_shape(T, S);
  const int box = small ? 64 : ROWS;
  if (int rc = make_tma_map_bf16(&qm, q, cols, (long long)B * T, ldq, box)) return rc;
  if (int rc = make_tma_map_bf16(&km, k, cols, (long long)B * S, ldk, box)) return rc;
  if (int rc = make_tma_map_bf16(&vm, v, cols, (long long)B * S, ldv, box)) return rc;
  if (int rc = make_tma_map_bf16(&dom, d_o, cols, (long long)B * T, lddo, box)) return rc;
  AttnTcArgs a{};
  a.B = B; a.H = H; a.T = T; a.S = S; a.dh = dh; a.scale = scale; a.key_pad = key_pad; a.causal = causal;
  a.lse = const_cast<float*>(lse);
  a.o_in = (const bf16*)o; a.ldo_in = ldo; a.d_o = (const bf16*)d_o; a.lddo = lddo;
  a.dq = (bf16*)dq; a.lddq = lddq; a.dk = (bf16*)dk; a.lddk = lddk; a.dv = (bf16*)dv; a.lddv = lddv;
  a.drop_state = drop_state; a.drop_p = drop_p; a.drop_site = drop_site;
  static bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem<128>(128)));
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem<64>(128)));
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = true;
  }
  if (small) launch_kernel(attn_bwd_tc_kernel<64>, dim3(B * H), dim3(BwdGeo<64>::NTHR), bwd_smem<64>(dh), stream, qm, km, vm, dom, a);
  else {
    // T > 128: q_tiles CTAs own a dQ tile each, k_tiles CTAs a dK / dV tile each (see the kernel's header comment)
    const int q_tiles = (T + ROWS - 1) / ROWS, k_tiles = (S + ROWS - 1) / ROWS;
    const dim3 grid(B * H, q_tiles == 1 ? 1 : q_tiles + k_tiles);
    launch_kernel(attn_bwd_tc_kernel<128>, grid, dim3(BwdGeo<128>::NTHR), bwd_smem<128>(dh), stream, qm, km, vm, dom, a);
  }
  B200_LAUNCH_CHECK("attn_bwd_tc_kernel");
  count_launch();
  return 0;
}

}  // namespace b200
