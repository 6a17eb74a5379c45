// bf16 GEMM on the 5th-gen tensor cores: TMA -> 128B-swizzled shared memory -> tcgen05.mma -> TMEM ->
// tcgen05.ld epilogue.  One kernel serves the dense Linear layers of the fusion blocks and the grouped
// expert FFN (forward, dgrad, wgrad) through operand major-ness flags and a tile->expert map.
//
// Persistent, warp-specialised CTA (one per SM, 320 threads):
//   warp 0      TMA producer (one elected lane) — ring of STAGES {A,B} k-blocks, full/empty mbarriers
//   warp 1      UMMA issuer (one elected lane); owns the TMEM allocation
//   warps 2..9  epilogue: two 128 x BN fp32 accumulators live in TMEM, so the epilogue of tile i overlaps the
//               main loop of tile i+1 (tmem_full / tmem_empty mbarriers).  Warp w reads TMEM lanes
//               32*(w%4)..+31 (its hardware quadrant) and every PARTS-th 32-column chunk of the tile:
//               tcgen05.ld -> bias / activation / residual in registers -> transpose through a padded
//               shared-memory staging tile -> row-contiguous 16-byte global stores (full 32-byte sectors).
#include <cuda.h>
#include <stdlib.h>

#include "gemm_common.cuh"
#include "tc_ptx.cuh"

namespace b200 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;
constexpr int CHUNK_BYTES = 64 * BK * 2;  // one 64(MN) x 64(K) MN-major TMA box
#ifndef B200_EPI_WARPS
#define B200_EPI_WARPS 8
#endif
constexpr int NUM_EPI_WARPS = B200_EPI_WARPS;      // multiple of 4: one warp per TMEM lane quadrant and column part
constexpr int EPI_PARTS = NUM_EPI_WARPS / 4;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
// Epilogue staging (per warp, 4 KB): a 32 x 32 tile goes thread-row -> shared -> row-contiguous global accesses.
//   bf16 tiles: 64-byte rows at a pitch of 80 bytes (the 16-byte pad spreads the banks);
//   fp32 tiles: 128-byte rows at a pitch of 128 bytes, the 16-byte units of row r XOR-swizzled by (r & 7).
// Keeping the buffer at 4 KB (no padded fp32 rows, no bias copy: a lane holds one bias value and broadcasts it with
// shuffles) is what leaves room for a FOURTH 48 KB pipeline stage of the 128 x 256 tile: with three stages only two
// k-blocks are in flight behind the one being multiplied (0.55 us of cover for ~1 us of TMA latency) and the tensor
// pipe idled half of the time (ncu: sm__pipe_tensor_cycles_active 44-53 % on the large GEMMs, profiles/r02e).
constexpr int STG_PITCH_BF16 = 80;
constexpr int STG_BYTES = 32 * 128;                // per epilogue warp

template <int BN> constexpr int num_stages() { return BN == 256 ? 4 : (BN == 128 ? 6 : 8); }
template <int BN> constexpr int stage_bytes() { return A_BYTES + BN * BK * 2; }
template <int BN> constexpr int smem_bytes() {
  return num_stages<BN>() * stage_bytes<BN>() + NUM_EPI_WARPS * STG_BYTES + 1024 + 256;
}
static_assert(smem_bytes<256>() <= 227 * 1024 && smem_bytes<128>() <= 227 * 1024 && smem_bytes<64>() <= 227 * 1024,
              "pipeline stages + epilogue staging must fit the 227 KB a CTA can own");

// Optional in-kernel timeline (build with -DB200_GEMM_TRACE, scripts/gemm_trace.py): CTA 0 of every launch records
// %globaltimer / clock64 at the phase boundaries of its first tile.  Not part of the shipped library.
#ifdef B200_GEMM_TRACE
__device__ unsigned long long g_trace[256 * 64];
__device__ unsigned int g_trace_n;
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define B200_TRACE(idx_)                                                                      \
  do {                                                                                         \
    if (blockIdx.x == 0 && tr_slot < 256u) {                                                   \
      g_trace[tr_slot * 64 + (idx_)] = trace_now();                                            \
      g_trace[tr_slot * 64 + 16 + (idx_)] = (unsigned long long)clock64();                     \
    }                                                                                          \
  } while (0)
// arrival of k-block kb_ (< 32) of the first tile, seen by the MMA issuer
#define B200_TRACE_KB(kb_)                                                                     \
  do {                                                                                         \
    if (blockIdx.x == 0 && tr_slot < 256u && (kb_) < 32)                                       \
      g_trace[tr_slot * 64 + 32 + (kb_)] = (unsigned long long)clock64();                      \
  } while (0)
#else
#define B200_TRACE(idx_) do { } while (0)
#define B200_TRACE_KB(kb_) do { } while (0)
#endif

// ---- thread-block clusters / distributed shared memory (split-K of small GEMMs, see launch_gemm_tc) -------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// every thread of every CTA of the cluster (warp-converged): orders prior (distributed) shared-memory accesses
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// split-phase form: arrive early (kernel prologue), wait where the guarantee is needed
__device__ __forceinline__ void cluster_arrive_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
// byte offset of the 16-byte unit j4 (4 columns) of (partial kr >= 1, 32-column chunk c, row) in the leader's reduction
// buffer: rows are the fastest index, so the 32 lanes of a warp (= 32 consecutive rows) touch 512 contiguous bytes
__device__ __forceinline__ uint32_t red_off(int kr, int c, int j4, int row, int nchunk, int red_rows) {
  return (uint32_t)(((((kr - 1) * nchunk + c) * 8 + j4) * red_rows + row) * 16);
}

struct TileInfo {
  int valid, group, m_tile, n_tile;
  int a_mn0, a_k0, b_mn0, b_k0, k_begin, k_blocks;
};

template <int BN, bool B_MN>
__device__ __forceinline__ TileInfo get_tile(const GemmArgs& p, int tile, int m_tiles, int n_tiles) {
  TileInfo t;
  const int per_z = m_tiles * n_tiles;
  const int z = tile / per_z, rem = tile - z * per_z;
  t.m_tile = rem / n_tiles;
  t.n_tile = rem - t.m_tile * n_tiles;
  t.valid = 1;
  t.group = 0;
  t.a_mn0 = t.m_tile * BM; t.a_k0 = 0; t.b_mn0 = t.n_tile * BN; t.b_k0 = 0;
  t.k_begin = 0;
  t.k_blocks = (p.K + BK - 1) / BK;
  if (p.mode == GEMM_GROUP_ROWS) {
    t.group = p.tile_group[t.m_tile];
    if (t.group < 0) { t.valid = 0; return t; }
    if (B_MN) t.b_k0 = t.group * p.b_group_rows; else t.b_mn0 += t.group * p.b_group_rows;
  } else if (p.mode == GEMM_GROUP_WGRAD) {
    t.group = z;
    const int r0 = p.group_off[z], r1 = p.group_off[z + 1];
    t.a_k0 = t.b_k0 = r0;
    t.k_blocks = (r1 - r0 + BK - 1) / BK;
  } else if (p.k_splits > 1) {
    const int per = (t.k_blocks + p.k_splits - 1) / p.k_splits;
    t.k_begin = z * per;
    t.k_blocks = min(per, t.k_blocks - t.k_begin);
    if (t.k_blocks <= 0) t.valid = 0;
  }
  return t;
}

// GELU(x) = x * Phi(x), Phi(x) = 0.5 * (1 + erf(x / sqrt 2)), with erf from Abramowitz & Stegun 7.1.26
// (|error| < 1.5e-7, far below bf16 resolution).  Written on the complementary tail so that no sign fix-up and no
// IEEE reciprocal (whose slow path costs a branch per element) is needed:
//   t = 1 / (1 + p |x| / sqrt 2),  q = 0.5 * poly(t) * t * exp(-x^2 / 2),  Phi = x >= 0 ? 1 - q : q
// 16 instructions per element, two of them MUFU (rcp.approx, ex2.approx).
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Phi(x) and e = exp(-x^2/2)
__device__ __forceinline__ float gelu_cdf(float x, float& e) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
  e = ex2_approx(x * x * (-0.5f * 1.4426950408889634f));
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  const float q = poly * t * e;
  return x >= 0.f ? 1.0f - q : q;
}
// Activation applied to a 32-value register tile with the activation kind resolved ONCE per tile (a per-element
// runtime switch made every element pay for all branches: 100+ instructions per GELU).
template <int ACT>
__device__ __forceinline__ float act1_fwd(float x) {
  if (ACT == B200_ACT_GELU) {
    float e;
    return x * gelu_cdf(x, e);
  }
  if (ACT == B200_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == B200_ACT_SILU) return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
  if (ACT == B200_ACT_TANH) return tanhf(x);
  return x;
}
template <int ACT>
__device__ __forceinline__ float act1_bwd(float x) {
  if (ACT == B200_ACT_GELU) {   // Phi(x) + x * phi(x), phi(x) = exp(-x^2/2) / sqrt(2 pi): the exponential is shared
    float e;
    const float cdf = gelu_cdf(x, e);
    return fmaf(x * 0.39894228040143267794f, e, cdf);
  }
  if (ACT == B200_ACT_RELU) return x > 0.f ? 1.f : 0.f;
  if (ACT == B200_ACT_SILU) {
    const float sg = rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
    return sg * (1.0f + x * (1.0f - sg));
  }
  if (ACT == B200_ACT_TANH) {
    const float th = tanhf(x);
    return 1.0f - th * th;
  }
  return 1.f;
}
// activation and its derivative from one evaluation (GELU: Phi and the exponential are shared)
template <int ACT>
__device__ __forceinline__ void act1_fwd_d(float x, float& y, float& d) {
  if (ACT == B200_ACT_GELU) {
    float e;
    const float cdf = gelu_cdf(x, e);
    y = x * cdf;
    d = fmaf(x * 0.39894228040143267794f, e, cdf);
  } else if (ACT == B200_ACT_SILU) {
    const float sg = rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
    y = x * sg;
    d = sg * (1.0f + x * (1.0f - sg));
  } else {
    y = act1_fwd<ACT>(x);
    d = act1_bwd<ACT>(x);
  }
}
// 32 pre-activations -> packed bf16 activation and packed bf16 backward factor (derivative x dropout keep-scale),
// eight elements at a time so that the fp32 inputs die as the packed outputs appear (no extra live registers: a
// separate 32-float derivative tile made the whole kernel spill).
template <int ACT>
__device__ __forceinline__ void tile_act_fwd_d_packed(const float (&v)[32], uint32_t (&ypk)[16], uint32_t (&dpk)[16],
                                                      const DropState& dsr, long long row, int ldo, int col0) {
  const unsigned long long base = (unsigned long long)row * (unsigned long long)ldo + (unsigned long long)col0;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    float sc[8];
    if (dsr.on) {
      if ((base & 7ull) == 0) {
        drop_scales8(dsr, (base + j) >> 3, sc);
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) sc[q] = drop_scale1(dsr, base + j + q);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) sc[q] = 1.f;
    }
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      float y0, d0, y1, d1;
      act1_fwd_d<ACT>(v[j + q], y0, d0);
      act1_fwd_d<ACT>(v[j + q + 1], y1, d1);
      ypk[(j + q) >> 1] = pack_bf16x2(y0 * sc[q], y1 * sc[q + 1]);
      dpk[(j + q) >> 1] = pack_bf16x2(d0 * sc[q], d1 * sc[q + 1]);
    }
  }
}
__device__ __forceinline__ void tile_act_fwd_d_packed(const float (&v)[32], uint32_t (&ypk)[16], uint32_t (&dpk)[16],
                                                      int act, const DropState& dsr, long long row, int ldo, int col0) {
  switch (act) {
    case B200_ACT_GELU: tile_act_fwd_d_packed<B200_ACT_GELU>(v, ypk, dpk, dsr, row, ldo, col0); break;
    case B200_ACT_RELU: tile_act_fwd_d_packed<B200_ACT_RELU>(v, ypk, dpk, dsr, row, ldo, col0); break;
    case B200_ACT_SILU: tile_act_fwd_d_packed<B200_ACT_SILU>(v, ypk, dpk, dsr, row, ldo, col0); break;
    case B200_ACT_TANH: tile_act_fwd_d_packed<B200_ACT_TANH>(v, ypk, dpk, dsr, row, ldo, col0); break;
    default: tile_act_fwd_d_packed<B200_ACT_NONE>(v, ypk, dpk, dsr, row, ldo, col0); break;
  }
}
template <int ACT>
__device__ __forceinline__ void tile_act_fwd(float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = act1_fwd<ACT>(v[j]);
}
template <int ACT>
__device__ __forceinline__ void tile_act_bwd(float (&v)[32], const float (&aux)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] *= act1_bwd<ACT>(aux[j]);
}
__device__ __forceinline__ void tile_act_fwd(float (&v)[32], int act) {
  switch (act) {
    case B200_ACT_GELU: tile_act_fwd<B200_ACT_GELU>(v); break;
    case B200_ACT_RELU: tile_act_fwd<B200_ACT_RELU>(v); break;
    case B200_ACT_SILU: tile_act_fwd<B200_ACT_SILU>(v); break;
    case B200_ACT_TANH: tile_act_fwd<B200_ACT_TANH>(v); break;
    default: break;
  }
}
__device__ __forceinline__ void tile_act_bwd(float (&v)[32], const float (&aux)[32], int act) {
  switch (act) {
    case B200_ACT_GELU: tile_act_bwd<B200_ACT_GELU>(v, aux); break;
    case B200_ACT_RELU: tile_act_bwd<B200_ACT_RELU>(v, aux); break;
    case B200_ACT_SILU: tile_act_bwd<B200_ACT_SILU>(v, aux); break;
    case B200_ACT_TANH: tile_act_bwd<B200_ACT_TANH>(v, aux); break;
    default: break;
  }
}

// ---- epilogue helpers: a warp moves a 32-row x 32-column tile between registers (thread = row) and global
// memory (row-contiguous 16-byte accesses) through its padded staging buffer -----------------------------------
// byte offset of 16-byte unit `u` of staged row `r`
template <int ELEM_BYTES>
__device__ __forceinline__ int stg_off(int r, int u) {
  return ELEM_BYTES == 2 ? r * STG_PITCH_BF16 + 16 * u : r * 128 + 16 * (u ^ (r & 7));
}

template <int ELEM_BYTES>  // 2 (bf16) or 4 (fp32)
__device__ __forceinline__ void stage_store_tile(uint8_t* stg, int lane, const float (&v)[32], void* gbase, long long ld,
                                                 long long row0, int rows_ok, int col0) {
  constexpr int ROW_BYTES = 32 * ELEM_BYTES;       // 64 or 128
  constexpr int LPR = ROW_BYTES / 16;              // lanes per row: 4 or 8
  constexpr int RPI = 32 / LPR;                    // rows per instruction: 8 or 4
  if (ELEM_BYTES == 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 t;
      t.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]); t.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
      t.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); t.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
      *reinterpret_cast<uint4*>(stg + stg_off<2>(lane, j)) = t;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(stg + stg_off<4>(lane, j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  __syncwarp();
  const int sub = lane % LPR, rsel = lane / LPR;
  // running pointer (one 64-bit add per row instead of a 64-bit multiply-add chain)
  uint8_t* dst = reinterpret_cast<uint8_t*>(gbase) + ((row0 + rsel) * ld + col0) * ELEM_BYTES + 16 * sub;
  const long long dstep = (long long)RPI * ld * ELEM_BYTES;
#pragma unroll
  for (int i = 0; i < 32 / RPI; ++i) {
    const int r = i * RPI + rsel;
    if (r < rows_ok) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(stg + stg_off<ELEM_BYTES>(r, sub));
    dst += dstep;
  }
  __syncwarp();
}

// the same for a tile whose 32 bf16 values per row are already packed (16 words per thread)
__device__ __forceinline__ void stage_store_packed_bf16(uint8_t* stg, int lane, const uint32_t (&pk)[16], void* gbase,
                                                        long long ld, long long row0, int rows_ok, int col0) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(stg + stg_off<2>(lane, j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  __syncwarp();
  const int sub = lane % 4, rsel = lane / 4;
  uint8_t* dst = reinterpret_cast<uint8_t*>(gbase) + ((row0 + rsel) * ld + col0) * 2 + 16 * sub;
  const long long dstep = 8ll * ld * 2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + rsel;
    if (r < rows_ok) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(stg + stg_off<2>(r, sub));
    dst += dstep;
  }
  __syncwarp();
}

__device__ __forceinline__ void stage_accum_tile(uint8_t* stg, int lane, const float (&v)[32], float* gbase, long long ld,
                                                 long long row0, int rows_ok, int col0) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + stg_off<4>(lane, j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  for (int r = 0; r < rows_ok; ++r)   // one row (32 consecutive floats) per warp instruction
    atomicAdd(gbase + (row0 + r) * ld + col0 + lane,
              *reinterpret_cast<const float*>(stg + stg_off<4>(r, lane >> 2) + 4 * (lane & 3)));
  __syncwarp();
}

__device__ __forceinline__ void stage_load_tile_bf16(uint8_t* stg, int lane, float (&v)[32], const void* gbase,
                                                     long long ld, long long row0, int rows_ok, int col0) {
  const int sub = lane % 4, rsel = lane / 4;
  const uint8_t* src = reinterpret_cast<const uint8_t*>(gbase) + ((row0 + rsel) * ld + col0) * 2 + 16 * sub;
  const long long sstep = 8ll * ld * 2;
  uint4 t[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {      // four independent loads in flight
    t[i] = make_uint4(0, 0, 0, 0);
    if (i * 8 + rsel < rows_ok) t[i] = __ldg(reinterpret_cast<const uint4*>(src));
    src += sstep;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stg + stg_off<2>(i * 8 + rsel, sub)) = t[i];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 t = *reinterpret_cast<const uint4*>(stg + stg_off<2>(lane, j));
    const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[8 * j + 0] = a.x; v[8 * j + 1] = a.y; v[8 * j + 2] = b.x; v[8 * j + 3] = b.y;
    v[8 * j + 4] = c.x; v[8 * j + 5] = c.y; v[8 * j + 6] = d.x; v[8 * j + 7] = d.y;
  }
  __syncwarp();
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const GemmArgs p, const int m_tiles_arg, const int n_tiles, const int total_tiles_arg,
               const int a_tx_bytes /* bytes one k-block of A brings in: A_BYTES, less for a short-row box */,
               const int kc /* > 1: cluster of kc CTAs splits K of ONE tile; rank 0 reduces + runs the epilogue */,
               const int red_rows /* kc > 1: rows of the tile that exist (32 or 64) */,
               const int red_base /* kc > 1: byte offset of the reduction buffer inside the pipeline-stage area; >= 0 and
                                     beyond the stages this launch's short pipelines touch: partial tiles may be
                                     pushed without waiting for the leader (one cluster barrier); < 0: offset 0, two */) {
  constexpr int STAGES = num_stages<BN>();
  constexpr int STAGE = stage_bytes<BN>();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* stg_all = smem + STAGES * STAGE;
  const uint32_t bar0 = base + STAGES * STAGE + NUM_EPI_WARPS * STG_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bar0 - base) + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef B200_GEMM_TRACE
  __shared__ unsigned int tr_slot_s;
  unsigned int tr_slot = 0xffffffffu;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    tr_slot_s = atomicAdd(&g_trace_n, 1u);
    tr_slot = tr_slot_s;
    B200_TRACE(0);
  }
#endif

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tma_a);
    ptx::prefetch_tensormap(&tma_b);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full_bar(a), 1);
      ptx::mbar_init(tmem_empty_bar(a), NUM_EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 2 * BN);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Split-K cluster whose reduction buffer lies behind the stages in use: the only thing the first cluster barrier has
  // to guarantee is that the leader CTA is running before anyone stores into its shared memory.  Arrive here, wait
  // (long since satisfied) right before the remote stores: a barrier phase without latency.
  if (kc > 1 && red_base >= 0) cluster_arrive_relaxed();
#ifdef B200_GEMM_TRACE
  if (blockIdx.x == 0) tr_slot = tr_slot_s;
  if (threadIdx.x == 0) B200_TRACE(1);
#endif
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) B200_TRACE(2);
  // Expert-parallel receive buffers are sized for the worst case; the rows actually in use are only known on the
  // device (tiles are numbered m-major, so bounding m_tiles bounds the tile walk).
  int m_tiles = m_tiles_arg, total_tiles = total_tiles_arg;
  if (p.rows_used != nullptr) {
    const int used_tiles = (__ldg(p.rows_used) + BM - 1) / BM;
    if (used_tiles < m_tiles) { m_tiles = used_tiles; total_tiles = used_tiles * n_tiles; }
  }

  // Cluster split-K: grid = tiles * kc, the kc CTAs of a cluster share a tile and take consecutive K ranges (get_tile's
  // split logic with z = rank).  Each CTA walks exactly one work item.
  const int kc_rank = kc > 1 ? (int)cluster_ctarank() : 0;
  const int tile0 = kc > 1 ? (int)(blockIdx.x / kc) + kc_rank * (m_tiles * n_tiles) : (int)blockIdx.x;
  const int tile_step = kc > 1 ? total_tiles : (int)gridDim.x;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int it = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const TileInfo t = get_tile<BN, B_MN>(p, tile, m_tiles, n_tiles);
        if (!t.valid) continue;
        for (int kb = 0; kb < t.k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(empty_bar(s), ph ^ 1u);
          ptx::mbar_arrive_expect_tx(full_bar(s), (uint32_t)(a_tx_bytes + BN * BK * 2));
          const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
          const int kc = (t.k_begin + kb) * BK;
          if (!A_MN) {
            ptx::tma_load_2d(sa, &tma_a, full_bar(s), t.a_k0 + kc, t.a_mn0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              ptx::tma_load_2d(sa + c * CHUNK_BYTES, &tma_a, full_bar(s), t.a_mn0 + 64 * c, t.a_k0 + kc);
          }
          if (!B_MN) {
            ptx::tma_load_2d(sb, &tma_b, full_bar(s), t.b_k0 + kc, t.b_mn0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              ptx::tma_load_2d(sb + c * CHUNK_BYTES, &tma_b, full_bar(s), t.b_mn0 + 64 * c, t.b_k0 + kc);
          }
        }
      }
    }
    if (kc > 1) {          // every thread of the cluster takes part in the reduction barriers (epilogue branch)
      __syncwarp();
      if (red_base < 0) cluster_sync_all(); else cluster_wait();
      cluster_sync_all();
    }
  } else if (warp == 1) {
    // ================= UMMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int it = 0, acc_it = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const TileInfo t = get_tile<BN, B_MN>(p, tile, m_tiles, n_tiles);
        if (!t.valid || t.k_blocks == 0) continue;
        const int acc = acc_it & 1;
        const uint32_t acc_ph = (acc_it >> 1) & 1;
        ptx::mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1u);   // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < t.k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(full_bar(s), ph);
          ptx::tc_fence_after();
          if (it == 0) B200_TRACE(3);
          if (acc_it == 0) B200_TRACE_KB(kb);
          const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: step 16 elements (32 B) inside the 128-byte swizzle row.
            // MN-major: step 16 k-rows (2048 B); LBO = distance between 64-wide MN chunks.
            const uint64_t ad = A_MN ? ptx::umma_smem_desc(sa + k * (UMMA_K * 128), CHUNK_BYTES, 1024)
                                     : ptx::umma_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
            const uint64_t bd = B_MN ? ptx::umma_smem_desc(sb + k * (UMMA_K * 128), CHUNK_BYTES, 1024)
                                     : ptx::umma_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
            ptx::umma_bf16(tmem_d, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(s));       // frees the smem slot once these MMAs retire
        }
        ptx::umma_commit(tmem_full_bar(acc));   // accumulator complete -> epilogue
        if (acc_it == 0) B200_TRACE(4);
        ++acc_it;
      }
    }
    if (kc > 1) {
      __syncwarp();
      if (red_base < 0) cluster_sync_all(); else cluster_wait();
      cluster_sync_all();
    }
  } else {
    // ================= epilogue warps =================
    const int ew = warp - 2;                 // 0..7
    const int quad = warp & 3;               // TMEM lane quadrant this warp may read
    const int part = ew >> 2;                // this warp handles chunks part, part + EPI_PARTS, ...
    constexpr int NCHUNK = BN / 32;
    uint8_t* stg = stg_all + ew * STG_BYTES;
    const bool out_bf16 = !p.out_f32;
    // fast path: every 16-byte group of a row is either fully inside or fully outside the matrix
    const bool vec_ok = (p.N % 8 == 0) && (p.ldo % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                        (p.ld_aux % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.aux_in) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.aux_out) & 15) == 0) &&
                        (p.mode != GEMM_GROUP_WGRAD || p.out_group_elems % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    const DropState drop = drop_load(p.drop_state, p.drop_p, p.drop_site);
    int acc_it = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const TileInfo t = get_tile<BN, B_MN>(p, tile, m_tiles, n_tiles);
      if (!t.valid) {
        if (kc > 1) {        // never leave the cluster's barriers short of a CTA (the host keeps every rank busy)
          __syncwarp();
          if (red_base < 0) cluster_sync_all(); else cluster_wait();
          cluster_sync_all();
        }
        continue;
      }
      const bool have_acc = t.k_blocks > 0;
      const int acc = acc_it & 1;
      const uint32_t acc_ph = (acc_it >> 1) & 1;
      const int ncol0 = t.n_tile * BN;                          // first column of the tile
      const float* bias = p.bias;
      if (bias != nullptr && p.mode == GEMM_GROUP_ROWS) bias += (long long)t.group * p.N;
      if (have_acc) {
        ptx::mbar_wait(tmem_full_bar(acc), acc_ph);
        ptx::tc_fence_after();
      }
      if (acc_it == 0 && ew == 0 && lane == 0) B200_TRACE(5);
      const long long row0 = (long long)t.m_tile * BM + quad * 32;
      const long long my_row = row0 + lane;
      long long rows_left = (long long)p.M - row0;
      const int rows_ok = rows_left >= 32 ? 32 : (rows_left > 0 ? (int)rows_left : 0);
      const long long obase = (p.mode == GEMM_GROUP_WGRAD) ? (long long)t.group * p.out_group_elems : 0ll;
      const bool red_warp = kc > 1 && quad * 32 < red_rows;      // this warp's TMEM lanes hold rows that exist
      const uint32_t red_at = red_base > 0 ? (uint32_t)red_base : 0u;
      if (kc > 1) {
        // #1 (only when the buffer overlaps stages the leader's pipeline uses): every CTA's MMAs have completed (its
        // epilogue warps waited for the accumulator), so the leader's pipeline stages are free to receive the partials
        __syncwarp();
        if (red_base < 0) cluster_sync_all(); else cluster_wait();
        if (kc_rank != 0) {
          if (red_warp && have_acc) {
            const uint32_t rbase = mapa_shared(base + red_at, 0);
#pragma unroll 1
            for (int c = part; c < NCHUNK; c += EPI_PARTS) {
              uint32_t r[32];
              ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                st_cluster_v4(rbase + red_off(kc_rank, c, j4, quad * 32 + lane, NCHUNK, red_rows),
                              __uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1]), __uint_as_float(r[4 * j4 + 2]),
                              __uint_as_float(r[4 * j4 + 3]));
            }
          }
          __syncwarp();
          cluster_sync_all();        // #2: partial tiles have landed in the leader
          ++acc_it;
          continue;                  // only the leader runs the epilogue
        }
        __syncwarp();
        cluster_sync_all();          // #2
      }
#pragma unroll 1
      for (int c = part, k = 0; c < NCHUNK; c += EPI_PARTS, ++k) {
        const int col0 = ncol0 + c * 32;
        float v[32];
        if (have_acc) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
          ptx::tmem_ld_wait();
          if (acc_it == 0 && ew == 0 && lane == 0 && k == 0) B200_TRACE(8);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (red_warp) {                          // leader of a split-K cluster: add the other CTAs' partial tiles
          for (int kr = 1; kr < kc; ++kr) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 q4 = *reinterpret_cast<const float4*>(smem + red_at + red_off(kr, c, j4, quad * 32 + lane, NCHUNK, red_rows));
              v[4 * j4] += q4.x; v[4 * j4 + 1] += q4.y; v[4 * j4 + 2] += q4.z; v[4 * j4 + 3] += q4.w;
            }
          }
        }
        if (col0 >= p.N) continue;               // warp-uniform
        if (vec_ok && col0 + 32 <= p.N) {
          if (bias != nullptr) {      // every lane needs the same 32 values: 8 uniform 16-byte loads (L1 broadcast)
            const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t4 = __ldg(b4 + j);
              v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
            }
          }
          if (acc_it == 0 && ew == 0 && lane == 0 && k == 0) B200_TRACE(9);
          if (p.epi == B200_EPI_ACT_D && out_bf16) {
            uint32_t ypk[16], dpk[16];
            tile_act_fwd_d_packed(v, ypk, dpk, p.act, drop, my_row, p.ldo, col0);
            if (p.aux_out != nullptr) stage_store_packed_bf16(stg, lane, dpk, p.aux_out, p.ld_aux, row0, rows_ok, col0);
            stage_store_packed_bf16(stg, lane, ypk, reinterpret_cast<bf16*>(p.out) + obase, p.ldo, row0, rows_ok, col0);
            continue;
          } else if (p.epi == B200_EPI_ACT_D) {      // fp32 output (not used by the drop-in modules): generic path
            epilogue_store<bf16, 32>(p, t.group, my_row, col0, v, my_row < p.M);
            continue;
          } else if (p.epi == B200_EPI_ACT) {
            if (p.aux_out != nullptr) stage_store_tile<2>(stg, lane, v, p.aux_out, p.ld_aux, row0, rows_ok, col0);
            tile_act_fwd(v, p.act);
            if (drop.on) apply_dropout_row<32>(drop, my_row, p.ldo, col0, v);
          } else if (p.epi == B200_EPI_MUL) {
            float aux[32];
            stage_load_tile_bf16(stg, lane, aux, p.aux_in, p.ld_aux, row0, rows_ok, col0);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= aux[j];
          } else if (p.epi == B200_EPI_ADD || p.epi == B200_EPI_DACT) {
            float aux[32];
            stage_load_tile_bf16(stg, lane, aux, p.aux_in, p.ld_aux, row0, rows_ok, col0);
            if (p.epi == B200_EPI_ADD) {
              if (drop.on) apply_dropout_row<32>(drop, my_row, p.ldo, col0, v);   // dropout(acc+bias) + residual
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += aux[j];
            } else {
              tile_act_bwd(v, aux, p.act);
              if (drop.on) apply_dropout_row<32>(drop, my_row, p.ldo, col0, v);
            }
          }
          if (p.epi == B200_EPI_ACCUM)
            stage_accum_tile(stg, lane, v, reinterpret_cast<float*>(p.out) + obase, p.ldo, row0, rows_ok, col0);
          else if (out_bf16)
            stage_store_tile<2>(stg, lane, v, reinterpret_cast<bf16*>(p.out) + obase, p.ldo, row0, rows_ok, col0);
          else
            stage_store_tile<4>(stg, lane, v, reinterpret_cast<float*>(p.out) + obase, p.ldo, row0, rows_ok, col0);
          if (acc_it == 0 && ew == 0 && lane == 0 && k == 0) B200_TRACE(10);
        } else {
          // ragged / unaligned tile: per-thread row path (bias handled inside)
          epilogue_store<bf16, 32>(p, t.group, my_row, col0, v, my_row < p.M);
        }
      }
      if (acc_it == 0 && ew == 0 && lane == 0) B200_TRACE(6);
      if (have_acc) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tmem_empty_bar(acc));   // accumulator may be overwritten
        ++acc_it;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) B200_TRACE(7);
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 2 * BN);
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

}  // namespace

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows of pitch `pitch` elements; the box is
// 64 (inner, = 128 B swizzle span) x box_outer.  Shared with the attention kernels.
int make_tma_map_bf16(void* map_out, const void* ptr, long long inner, long long outer, long long pitch,
                      int box_outer) {
  CUtensorMap* m = reinterpret_cast<CUtensorMap*>(map_out);
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return B200_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  cuuint32_t es[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%lld outer=%lld pitch=%lld box_outer=%d", (int)r,
              ptr, inner, outer, pitch, box_outer);
    return B200_ERR_CUDA;
  }
  return 0;
}

namespace {

// launch_kernel (common.cuh) plus a thread-block cluster of `cluster_x` CTAs along x
template <typename... KP, typename... A>
cudaError_t launch_kernel_cluster(void (*kern)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  int cluster_x, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster_x;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KP>(args)...);
}

template <int BN, bool A_MN, bool B_MN>
int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, int m_tiles, int n_tiles,
                   int total_tiles, int a_tx_bytes, int kc, int red_rows, cudaStream_t stream) {
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BN>()));
    configured = true;
  }
  if (kc > 1) {     // one CTA per (tile, K range): total_tiles = tiles * kc, clusters of kc consecutive CTAs
    B200_CHECK_ARG((size_t)(kc - 1) * BN * red_rows * 4 <= (size_t)num_stages<BN>() * stage_bytes<BN>(),
                   "gemm: split-K reduction buffer does not fit the pipeline stages (kc=%d BN=%d rows=%d)", kc, BN, red_rows);
    // Each CTA runs K / kc k-blocks through stages 0 .. kb-1 once; a reduction buffer behind them can be written while
    // the leader is still multiplying (no barrier before the partial tiles are pushed).
    const int kb_per_cta = ((args.K + BK - 1) / BK + kc - 1) / kc;
    const size_t red_bytes = (size_t)(kc - 1) * BN * red_rows * 4;
    const size_t area = (size_t)num_stages<BN>() * stage_bytes<BN>();
    const size_t used = (size_t)(kb_per_cta < num_stages<BN>() ? kb_per_cta : num_stages<BN>()) * stage_bytes<BN>();
    const int red_base = (used + red_bytes <= area && used > 0) ? (int)(area - red_bytes) / 1024 * 1024 : -1;
    launch_kernel_cluster(kern, dim3(total_tiles), dim3(NUM_THREADS), smem_bytes<BN>(), stream, kc, ta, tb, args, m_tiles,
                          n_tiles, total_tiles, a_tx_bytes, kc, red_rows, red_base >= (int)used ? red_base : -1);
    B200_LAUNCH_CHECK("gemm_tc_kernel (cluster split-K)");
    count_launch();
    return 0;
  }
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  launch_kernel(kern, dim3(grid), dim3(NUM_THREADS), smem_bytes<BN>(), stream, ta, tb, args, m_tiles, n_tiles, total_tiles,
                a_tx_bytes, 1, 0, -1);
  B200_LAUNCH_CHECK("gemm_tc_kernel");
  count_launch();
  return 0;
}

template <int BN>
int launch_bn(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, int m_tiles,
              int n_tiles, int total_tiles, int a_tx, int kc, int red_rows, cudaStream_t stream) {
  if (!a_mn && !b_mn)
    return launch_variant<BN, false, false>(ta, tb, args, m_tiles, n_tiles, total_tiles, a_tx, kc, red_rows, stream);
  if (!a_mn && b_mn)
    return launch_variant<BN, false, true>(ta, tb, args, m_tiles, n_tiles, total_tiles, a_tx, kc, red_rows, stream);
  if (a_mn && b_mn)
    return launch_variant<BN, true, true>(ta, tb, args, m_tiles, n_tiles, total_tiles, a_tx, kc, red_rows, stream);
  return launch_variant<BN, true, false>(ta, tb, args, m_tiles, n_tiles, total_tiles, a_tx, kc, red_rows, stream);
}

}  // namespace

int launch_gemm_tc(const void* A, int lda, int a_layout, long long a_mn_extent, long long a_k_extent,
                   const void* B, int ldb, int b_layout, long long b_mn_extent, long long b_k_extent,
                   GemmArgs args, int grid_m_tiles, int groups, cudaStream_t stream) {
  const bool a_mn = a_layout == B200_LAYOUT_MN, b_mn = b_layout == B200_LAYOUT_MN;
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                 "gemm: operand base pointers must be 16-byte aligned");
  B200_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm: bf16 operand pitches must be multiples of 8 (lda=%d ldb=%d)",
                 lda, ldb);

  // Tile-N / split-K selection by a small cost model (unit: the time of one 128x256x64 k-block, ~0.27 us):
  //   time = fill + tiles_per_cta * max(k_blocks * c(BN), epilogue(BN)) + epilogue(BN)
  // c(BN) is the per-k-block time of a tile (256: tensor-pipe bound; 128 / 64: bound by the shared-memory fill
  // rate, 32 / 24 KB per k-block), the epilogue of tile i overlaps the main loop of tile i+1, the last one is
  // exposed.  Split-K (atomic fp32 accumulation) is available to dense wgrad GEMMs only.
  const int sms = num_sms();
  const long long m_tiles = grid_m_tiles;
  const int zdim = args.mode == GEMM_GROUP_WGRAD ? groups : 1;
  // grouped wgrad reduces over data-dependent row segments: use the mean segment length
  const int kblocks = args.mode == GEMM_GROUP_WGRAD ? (int)((a_k_extent / (groups > 0 ? groups : 1) + BK - 1) / BK)
                                                    : (args.K + BK - 1) / BK;
  const bool can_split = args.mode == GEMM_DENSE && args.epi == B200_EPI_ACCUM;
  // epilogue of a 128 x 256 tile in k-block units (8 warps, 32 x 32 chunks): GELU / GELU' ~ 1 000 warp instructions per
  // chunk (profiles/r01o_ncu_source_gemm_act_after.txt) = about 9 k-blocks; it overlaps the next tile's main loop
  const float epi256 = (args.epi == B200_EPI_ACT || args.epi == B200_EPI_DACT || args.epi == B200_EPI_ACT_D) ? 9.f
                       : (args.epi == B200_EPI_ACCUM ? 8.f
                          : ((args.epi == B200_EPI_ADD || args.epi == B200_EPI_MUL) ? 5.f : 3.f));
  int bn = 64, splits = 1;
  float best = 1e30f;
  static const int forced_bn = []() { const char* e = getenv("B200VQA_GEMM_BN"); return e ? atoi(e) : 0; }();
  for (int cand : {256, 128, 64}) {
    if (forced_bn && cand != forced_bn) continue;
    const float c = cand == 256 ? 1.0f : (cand == 128 ? 0.68f : 0.51f);
    const float epi = epi256 * (float)cand / 256.f + 1.5f;   // + per-tile fixed cost (barrier waits, bias staging)
    const long long tiles = m_tiles * ((args.N + cand - 1) / cand) * zdim;
    for (int s_ = 1; s_ <= (can_split ? 16 : 1); ++s_) {
      if (s_ > 1 && kblocks / s_ < 4) break;
      const long long items = tiles * s_;
      const long long ctas = items < sms ? items : sms;
      const float tpc = (float)((items + ctas - 1) / ctas);
      const float kb = (float)((kblocks + s_ - 1) / s_);
      const float main_t = kb * c;
      const float t = 6.f + tpc * (main_t > epi ? main_t : epi) + epi + (s_ > 1 ? 1.0f * s_ : 0.f);
      if (t < best * 0.999f) { best = t; bn = cand; splits = s_; }
    }
  }
  // Cluster split-K for single-row-tile GEMMs with few rows (M = batch: the CLS-row GEMMs of the last fusion layer,
  // the pooled vectors, the classifier head).  Their main loop is bound by 140 clocks per 128-row UMMA x k-steps on a
  // handful of SMs (timeline in DESIGN.md, section 4) while 130+ SMs idle: kc CTAs of a thread-block cluster take K / kc
  // each, push their partial tiles (only the rows that exist) into the leader's free pipeline stages through distributed
  // shared memory, and the leader runs the unchanged epilogue.  No atomics, no workspace, deterministic.
  int kc = 1;
  static const bool cluster_on = []() { const char* e = getenv("B200VQA_GEMM_CLUSTER"); return !(e && e[0] == '0'); }();
  if (cluster_on && !forced_bn && !a_mn && m_tiles == 1 && a_mn_extent <= 64 && args.mode == GEMM_DENSE &&
      args.epi != B200_EPI_ACCUM && kblocks >= 4) {
    const int rows = a_mn_extent <= 32 ? 32 : 64;
    float best_t = 1e30f;
    int best_bn = bn, best_kc = 1;
    for (int cand : {64, 128}) {
      const int tiles = (args.N + cand - 1) / cand;
      const size_t stage_cap = cand == 64 ? (size_t)num_stages<64>() * stage_bytes<64>() : (size_t)num_stages<128>() * stage_bytes<128>();
      for (int c : {8, 4, 2}) {
        if (kblocks % c != 0 || tiles * c > sms) continue;
        if ((size_t)(c - 1) * cand * rows * 4 > stage_cap) continue;
        // k-blocks per CTA at the UMMA instruction floor + the reduction (barriers, partial tiles) + epilogue chunks
        const float t = (float)(kblocks / c) * (cand == 64 ? 561.f : 585.f) + 900.f + 300.f * c + 2000.f * (cand / 64);
        if (t < best_t) { best_t = t; best_bn = cand; best_kc = c; }
      }
    }
    const float t_plain = (float)kblocks * (bn == 64 ? 561.f : (bn == 128 ? 585.f : 800.f)) + 2000.f * (bn / 64);
    if (best_kc > 1 && best_t < t_plain) {
      bn = best_bn;
      kc = best_kc;
      splits = kc;
    }
  }
  args.k_splits = splits;
  if (args.epi == B200_EPI_ACCUM && args.mode == GEMM_DENSE) {
    if (splits > 1)   // partial sums are added atomically into a zeroed buffer
      B200_CUDA(cudaMemsetAsync(args.out, 0, (size_t)args.M * args.ldo * sizeof(float), stream));
    else              // one CTA owns each output tile: plain fp32 stores
      args.epi = B200_EPI_NONE;
  }

  CUtensorMap ta, tb;
  int rc;
  // A single row tile with few rows (the CLS-row GEMMs of the last fusion layer, the pooled vectors of the
  // classification pipeline: M = batch): the TMA box covers only 32 or 64 rows.  A k-block of a 128-row box costs the
  // producer the same whether its rows are in bounds or zero-filled (timeline: 525 clocks per k-block at M = 32 and at
  // M = 2048); the UMMA still multiplies 128 rows, rows beyond the box hold whatever the stage held before and only
  // reach accumulator rows that are never stored.
  int a_box = BM;
  if (!a_mn && m_tiles == 1 && args.mode == GEMM_DENSE) a_box = a_mn_extent <= 32 ? 32 : (a_mn_extent <= 64 ? 64 : BM);
  const int a_tx = a_mn ? A_BYTES : a_box * BK * 2;
  if (!a_mn) rc = make_tma_map_bf16(&ta, A, a_k_extent, a_mn_extent, lda, a_box);
  else rc = make_tma_map_bf16(&ta, A, a_mn_extent, a_k_extent, lda, BK);
  if (rc) return rc;
  if (!b_mn) rc = make_tma_map_bf16(&tb, B, b_k_extent, b_mn_extent, ldb, bn);
  else rc = make_tma_map_bf16(&tb, B, b_mn_extent, b_k_extent, ldb, BK);
  if (rc) return rc;

  const int n_tiles = (args.N + bn - 1) / bn;
  const int total = (int)(m_tiles * n_tiles * (args.mode == GEMM_GROUP_WGRAD ? groups : splits));
  const int red_rows = kc > 1 ? a_box : 0;
  if (bn == 256) return launch_bn<256>(a_mn, b_mn, ta, tb, args, (int)m_tiles, n_tiles, total, a_tx, 1, 0, stream);
  if (bn == 128) return launch_bn<128>(a_mn, b_mn, ta, tb, args, (int)m_tiles, n_tiles, total, a_tx, kc, red_rows, stream);
  return launch_bn<64>(a_mn, b_mn, ta, tb, args, (int)m_tiles, n_tiles, total, a_tx, kc, red_rows, stream);
}

}  // namespace b200

#ifdef B200_GEMM_TRACE
// copies the timeline out and resets the launch counter: out[launch * 64 + {0..15: globaltimer ns, 16..31: clock64, 32..63: clock64 at the arrival of k-block 0..31}]
extern "C" int b200_debug_gemm_trace(unsigned long long* host_out, unsigned int* n_out) {
  unsigned int zero = 0;
  if (cudaMemcpyFromSymbol(host_out, b200::g_trace, sizeof(unsigned long long) * 256 * 64) != cudaSuccess) return 1;
  if (cudaMemcpyFromSymbol(n_out, b200::g_trace_n, sizeof(unsigned int)) != cudaSuccess) return 1;
  if (cudaMemcpyToSymbol(b200::g_trace_n, &zero, sizeof(unsigned int)) != cudaSuccess) return 1;
  return 0;
}
#endif
