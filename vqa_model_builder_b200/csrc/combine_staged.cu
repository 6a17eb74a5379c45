// MOE combine (+ output LayerNorm), forward and backward, bf16, top-k <= 2, as persistent shared-memory-staged kernels.
// The K expert rows of a token are scattered over the padded expert layout (dest_row), so a chunk of 8 tokens is
// gathered by up to 16 row-sized bulk asynchronous copies issued by the lanes of warp 0 (destination indices are
// prefetched one chunk ahead); the upstream-gradient rows of the chunk are contiguous and arrive in one copy.
// Backward reads every expert row ONCE from HBM: the recomputed weighted sum (LayerNorm input), the LayerNorm
// backward, dz = w * ds and d_w = <ds, z> all work out of the staged copy (the register kernel of dispatch.cu read z
// twice).  dgamma / dbeta partial sums live in registers and are folded once per block.
#include "rowops.cuh"
#include "rowpipe.cuh"

namespace b200 {

int launch_ln_staged_reduce(const float* part, const int* part_group, int entries, int D, int G, int nz, float* dgamma,
                            float* dbeta, float* dcol, cudaStream_t stream);

namespace {

constexpr int CS_WARPS = 8;           // tokens per chunk
constexpr int CS_THREADS = CS_WARPS * 32;
constexpr int CS_MAX_STAGES = 8;
constexpr int CS_HDR = 128;           // full[8] + empty[8] mbarriers

struct TokenMeta {   // per-token scalars, prefetched one chunk ahead
  int d[2];
  float w[2];
  float mean, rstd;
};

template <int KK, bool STATS>
__device__ __forceinline__ TokenMeta load_meta(const int* __restrict__ dest_row, const float* __restrict__ w,
                                               const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                               int n) {
  TokenMeta m;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    m.d[k] = k < KK ? __ldg(dest_row + (long long)n * KK + k) : -1;
    m.w[k] = k < KK ? __ldg(w + (long long)n * KK + k) : 0.f;
  }
  m.mean = STATS ? __ldg(mean_in + n) : 0.f;
  m.rstd = STATS ? __ldg(rstd_in + n) : 0.f;
  return m;
}

// ---------------------------------------------------------------------------------------------------------------
// forward: out[n] = LN( sum_k w[n,k] * z[dest_row[n,k]] ) * gamma + beta
// ---------------------------------------------------------------------------------------------------------------
template <int NV, int KK>
__global__ void __launch_bounds__(CS_THREADS, 2)
combine_fwd_staged_kernel(const bf16* __restrict__ z, const int* __restrict__ dest_row, const float* __restrict__ w,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int N,
                          bf16* __restrict__ out, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                          int stages) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int D = NV * 256;
  constexpr uint32_t ROWB = D * sizeof(bf16);
  constexpr uint32_t STAGEB = CS_WARPS * KK * ROWB;
  const uint32_t full = ptx::smem_u32(smem), empty = full + 8 * CS_MAX_STAGES;
  unsigned char* ring = smem + CS_HDR;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      ptx::mbar_init(full + 8 * i, 1);
      ptx::mbar_init(empty + 8 * i, CS_WARPS);
    }
    ptx::fence_mbar_init();
  }
  pdl_trigger();
  pdl_wait();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = ((N + gridDim.x - 1) / gridDim.x + CS_WARPS - 1) / CS_WARPS * CS_WARPS;
  const int t_begin = blockIdx.x * q, t_end = min(N, t_begin + q);
  const int nchunks = t_end > t_begin ? (t_end - t_begin + CS_WARPS - 1) / CS_WARPS : 0;

  // feeder state (warp 0): destination row of pair `lane` of the next chunk to issue
  int idx_nx = -1;
  auto fetch_idx = [&](int c) {
    const long long p = (long long)(t_begin + c * CS_WARPS) * KK + lane;
    idx_nx = (c < nchunks && lane < CS_WARPS * KK && p < (long long)t_end * KK) ? __ldg(dest_row + p) : -1;
  };
  auto issue = [&](int c) {   // whole warp 0
    const int st = c % stages;
    ptx::mbar_wait(empty + 8 * st, (((uint32_t)(c / stages)) & 1u) ^ 1u);
    const int my = idx_nx;
    fetch_idx(c + 1);
    const bool valid = my >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const uint32_t bar = full + 8 * st;
    if (lane == 0) ptx::mbar_arrive_expect_tx(bar, (uint32_t)__popc(m) * ROWB);
    __syncwarp();
    if (valid) bulk_g2s(ptx::smem_u32(ring + (size_t)st * STAGEB + (size_t)lane * ROWB), z + (long long)my * D, ROWB, bar);
  };
  if (warp == 0) {
    fetch_idx(0);
    for (int c = 0; c < min(stages - 1, nchunks); ++c) issue(c);
  }

  float gam[NV][8], bet[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    load_param<8>(gamma, lane + 32 * j, gam[j]);
    load_param<8>(beta, lane + 32 * j, bet[j]);
  }
  TokenMeta nx;
  nx.d[0] = nx.d[1] = -1; nx.w[0] = nx.w[1] = 0.f; nx.mean = nx.rstd = 0.f;
  if (t_begin + warp < t_end) nx = load_meta<KK, false>(dest_row, w, nullptr, nullptr, t_begin + warp);
  RingPos rp;
  for (int c = 0; c < nchunks; ++c) {
    if (warp == 0 && c + stages - 1 < nchunks) issue(c + stages - 1);
    const int t0 = t_begin + c * CS_WARPS;
    const int n = t0 + warp;
    const bool has = n < t_end;
    const TokenMeta cur = nx;
    if (n + CS_WARPS < t_end) nx = load_meta<KK, false>(dest_row, w, nullptr, nullptr, n + CS_WARPS);
    ptx::mbar_wait(full + 8 * rp.stage, rp.phase);
    if (has) {
      const unsigned char* sb = ring + (size_t)rp.stage * STAGEB + (size_t)warp * KK * ROWB;
      float acc[NV][8];
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int vi = lane + 32 * j;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[j][u] = 0.f;
#pragma unroll
        for (int k = 0; k < KK; ++k)
          if (cur.d[k] >= 0) {
            float zv[8];
            unpack8(*reinterpret_cast<const uint4*>(sb + (size_t)k * ROWB + vi * 16), zv);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[j][u] = fmaf(cur.w[k], zv[u], acc[j][u]);
          }
#pragma unroll
        for (int u = 0; u < 8; ++u) sum += acc[j][u];
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(empty + 8 * rp.stage);
      const float mean = warp_sum(sum) * (1.0f / D);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc[j][u] -= mean;
          sq = fmaf(acc[j][u], acc[j][u], sq);
        }
      const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) o[u] = fmaf(acc[j][u] * rstd, gam[j][u], bet[j][u]);
        *reinterpret_cast<uint4*>(out + (long long)n * D + (lane + 32 * j) * 8) = pack8(o);
      }
      if (lane == 0) {
        mean_out[n] = mean;
        rstd_out[n] = rstd;
      }
    } else {
      if (lane == 0) ptx::mbar_arrive(empty + 8 * rp.stage);
    }
    rp.advance(stages);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward: ds = LN'(dout), dz[dest_row[n,k]] = w[n,k] * ds, d_w[n,k] = <ds, z[dest_row[n,k]]>, dgamma / dbeta partials
// ---------------------------------------------------------------------------------------------------------------
template <int NV, int KK>
__global__ void __launch_bounds__(CS_THREADS, 1)
combine_bwd_staged_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ z, const int* __restrict__ dest_row,
                          const float* __restrict__ w, const float* __restrict__ mean_in,
                          const float* __restrict__ rstd_in, const float* __restrict__ gamma, int N,
                          bf16* __restrict__ dz, float* __restrict__ d_w, float* __restrict__ part,
                          int* __restrict__ part_group, int stages) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int D = NV * 256;
  constexpr uint32_t ROWB = D * sizeof(bf16);
  constexpr uint32_t STAGEB = CS_WARPS * (KK + 1) * ROWB;    // [8 dout rows][8*KK expert rows]
  const uint32_t full = ptx::smem_u32(smem), empty = full + 8 * CS_MAX_STAGES;
  unsigned char* ring = smem + CS_HDR;
  float* red = reinterpret_cast<float*>(ring + (size_t)stages * STAGEB);   // [8][2][D]
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      ptx::mbar_init(full + 8 * i, 1);
      ptx::mbar_init(empty + 8 * i, CS_WARPS);
    }
    ptx::fence_mbar_init();
  }
  pdl_trigger();
  pdl_wait();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = ((N + gridDim.x - 1) / gridDim.x + CS_WARPS - 1) / CS_WARPS * CS_WARPS;
  const int t_begin = blockIdx.x * q, t_end = min(N, t_begin + q);
  const int nchunks = t_end > t_begin ? (t_end - t_begin + CS_WARPS - 1) / CS_WARPS : 0;
  if (nchunks == 0) {
    if (threadIdx.x == 0) part_group[blockIdx.x] = -1;
    return;
  }

  int idx_nx = -1;
  auto fetch_idx = [&](int c) {
    const long long p = (long long)(t_begin + c * CS_WARPS) * KK + lane;
    idx_nx = (c < nchunks && lane < CS_WARPS * KK && p < (long long)t_end * KK) ? __ldg(dest_row + p) : -1;
  };
  auto issue = [&](int c) {   // whole warp 0
    const int st = c % stages;
    ptx::mbar_wait(empty + 8 * st, (((uint32_t)(c / stages)) & 1u) ^ 1u);
    const int my = idx_nx;
    fetch_idx(c + 1);
    const int t0 = t_begin + c * CS_WARPS;
    const int ntok = min(CS_WARPS, t_end - t0);
    const bool valid = my >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const uint32_t bar = full + 8 * st;
    unsigned char* dst = ring + (size_t)st * STAGEB;
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar, (uint32_t)(__popc(m) + ntok) * ROWB);
      bulk_g2s(ptx::smem_u32(dst), dout + (long long)t0 * D, (uint32_t)ntok * ROWB, bar);
    }
    __syncwarp();
    if (valid) bulk_g2s(ptx::smem_u32(dst + (size_t)(CS_WARPS + lane) * ROWB), z + (long long)my * D, ROWB, bar);
  };
  if (warp == 0) {
    fetch_idx(0);
    for (int c = 0; c < min(stages - 1, nchunks); ++c) issue(c);
  }

  float gam[NV][8], acc_g[NV][8], acc_b[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    load_param<8>(gamma, lane + 32 * j, gam[j]);
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc_g[j][u] = 0.f; acc_b[j][u] = 0.f; }
  }
  TokenMeta nx;
  nx.d[0] = nx.d[1] = -1; nx.w[0] = nx.w[1] = 0.f; nx.mean = nx.rstd = 0.f;
  if (t_begin + warp < t_end) nx = load_meta<KK, true>(dest_row, w, mean_in, rstd_in, t_begin + warp);
  RingPos rp;
  for (int c = 0; c < nchunks; ++c) {
    if (warp == 0 && c + stages - 1 < nchunks) issue(c + stages - 1);
    const int t0 = t_begin + c * CS_WARPS;
    const int n = t0 + warp;
    const bool has = n < t_end;
    const TokenMeta cur = nx;
    if (n + CS_WARPS < t_end) nx = load_meta<KK, true>(dest_row, w, mean_in, rstd_in, n + CS_WARPS);
    ptx::mbar_wait(full + 8 * rp.stage, rp.phase);
    if (has) {
      const unsigned char* sd = ring + (size_t)rp.stage * STAGEB + (size_t)warp * ROWB;
      const unsigned char* sz = ring + (size_t)rp.stage * STAGEB + (size_t)(CS_WARPS + warp * KK) * ROWB;
      float xh[NV][8], gg[NV][8];
      float s1 = 0.f, s2 = 0.f;
      const float nmr = -cur.mean * cur.rstd;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int vi = lane + 32 * j;
        float dv[8], sv[8];
        unpack8(*reinterpret_cast<const uint4*>(sd + vi * 16), dv);
#pragma unroll
        for (int u = 0; u < 8; ++u) sv[u] = 0.f;
#pragma unroll
        for (int k = 0; k < KK; ++k)
          if (cur.d[k] >= 0) {
            float zv[8];
            unpack8(*reinterpret_cast<const uint4*>(sz + (size_t)k * ROWB + vi * 16), zv);
#pragma unroll
            for (int u = 0; u < 8; ++u) sv[u] = fmaf(cur.w[k], zv[u], sv[u]);
          }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float xhat = fmaf(sv[u], cur.rstd, nmr);
          const float d = dv[u];
          acc_g[j][u] = fmaf(d, xhat, acc_g[j][u]);
          acc_b[j][u] += d;
          const float t = d * gam[j][u];
          xh[j][u] = xhat;
          gg[j][u] = t;
          s1 += t;
          s2 = fmaf(t, xhat, s2);
        }
      }
      s1 = warp_sum(s1) * (1.0f / D);
      s2 = warp_sum(s2) * (1.0f / D);
      const float c1 = -cur.rstd * s1, c2 = -cur.rstd * s2;
      float dot[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int vi = lane + 32 * j;
        float ds[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) ds[u] = fmaf(xh[j][u], c2, fmaf(gg[j][u], cur.rstd, c1));
#pragma unroll
        for (int k = 0; k < KK; ++k)
          if (cur.d[k] >= 0) {
            float zv[8], o[8];
            unpack8(*reinterpret_cast<const uint4*>(sz + (size_t)k * ROWB + vi * 16), zv);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              dot[k] = fmaf(ds[u], zv[u], dot[k]);
              o[u] = cur.w[k] * ds[u];
            }
            *reinterpret_cast<uint4*>(dz + (long long)cur.d[k] * D + vi * 8) = pack8(o);
          }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(empty + 8 * rp.stage);
#pragma unroll
      for (int k = 0; k < KK; ++k) {
        const float t = warp_sum(dot[k]);
        if (lane == 0) d_w[(long long)n * KK + k] = cur.d[k] >= 0 ? t : 0.f;
      }
    } else {
      if (lane == 0) ptx::mbar_arrive(empty + 8 * rp.stage);
    }
    rp.advance(stages);
  }

  // fold the block's dgamma / dbeta partials
  float* mine = red + (size_t)warp * 2 * D;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int vi = lane + 32 * j;
#pragma unroll
    for (int u = 0; u < 8; u += 4) {
      *reinterpret_cast<float4*>(mine + vi * 8 + u) =
          make_float4(acc_g[j][u], acc_g[j][u + 1], acc_g[j][u + 2], acc_g[j][u + 3]);
      *reinterpret_cast<float4*>(mine + D + vi * 8 + u) =
          make_float4(acc_b[j][u], acc_b[j][u + 1], acc_b[j][u + 2], acc_b[j][u + 3]);
    }
  }
  __syncthreads();
  float* outp = part + (size_t)blockIdx.x * 2 * D;
  for (int i = threadIdx.x; i < 2 * D; i += CS_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int wp = 0; wp < CS_WARPS; ++wp) t += red[(size_t)wp * 2 * D + i];
    outp[i] = t;
  }
  if (threadIdx.x == 0) part_group[blockIdx.x] = 0;
}

inline int staged_grid(int N, int per_sm) {
  const int chunks = (N + CS_WARPS - 1) / CS_WARPS;
  const int cap = num_sms() * per_sm;
  return chunks < cap ? chunks : cap;
}

inline int pick_stages(size_t budget, size_t fixed, size_t stage, int chunks_per_block) {
  if (fixed + 2 * stage > budget) return 0;
  size_t st = (budget - fixed) / stage;
  if (st > CS_MAX_STAGES) st = CS_MAX_STAGES;
  const size_t need = chunks_per_block < 2 ? 2 : (size_t)chunks_per_block + 1;
  if (st > need) st = need;
  return (int)st;
}

}  // namespace

bool ln_staged_enabled();

size_t combine_bwd_staged_ws(int D) { return (size_t)160 * (2 * (size_t)D * sizeof(float) + sizeof(int)) + 512; }

int launch_combine_fwd_staged(const bf16* z, const int* dest_row, const float* w, const float* gamma,
                              const float* beta, float eps, int N, int K, int D, bf16* out, float* mean, float* rstd,
                              cudaStream_t stream) {
  if (!ln_staged_enabled() || gamma == nullptr || K < 1 || K > 2 || D % 256 != 0 || D > 1024) return -1;
  const int grid = staged_grid(N, 2);
  const int q = ((N + grid - 1) / grid + CS_WARPS - 1) / CS_WARPS;
  const size_t stage = (size_t)CS_WARPS * K * D * sizeof(bf16);
  const int stages = pick_stages(108 * 1024, CS_HDR + 128, stage, q);
  if (stages < 2) return -1;
  const size_t smem = CS_HDR + 128 + (size_t)stages * stage;
#define B200_CF(NVC, KC)                                                                                            \
  {                                                                                                                 \
    static bool attr_set = false;                                                                                   \
    if (!attr_set) {                                                                                                \
      B200_CUDA(cudaFuncSetAttribute(combine_fwd_staged_kernel<NVC, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     227 * 1024));                                                                  \
      attr_set = true;                                                                                              \
    }                                                                                                               \
    launch_kernel(combine_fwd_staged_kernel<NVC, KC>, dim3(grid), dim3(CS_THREADS), smem, stream, z, dest_row, w,   \
                  gamma, beta, eps, N, out, mean, rstd, stages);                                                    \
  }
#define B200_CF_K(NVC) \
  if (K == 1) B200_CF(NVC, 1) else B200_CF(NVC, 2)
  switch (D / 256) {
    case 1: B200_CF_K(1) break;
    case 2: B200_CF_K(2) break;
    case 3: B200_CF_K(3) break;
    default: B200_CF_K(4) break;
  }
#undef B200_CF_K
#undef B200_CF
  B200_LAUNCH_CHECK("combine_fwd_staged_kernel");
  count_launch();
  return 0;
}

int launch_combine_bwd_staged(const bf16* dout, const bf16* z, const int* dest_row, const float* w, const float* mean,
                              const float* rstd, const float* gamma, int N, int K, int D, bf16* dz, float* d_w,
                              float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                              cudaStream_t stream) {
  if (!ln_staged_enabled() || gamma == nullptr || K < 1 || K > 2 || D % 256 != 0 || D > 1024) return -1;
  const int grid = staged_grid(N, 1);
  const int q = ((N + grid - 1) / grid + CS_WARPS - 1) / CS_WARPS;
  const size_t stage = (size_t)CS_WARPS * (K + 1) * D * sizeof(bf16);
  const size_t red = (size_t)CS_WARPS * 2 * D * sizeof(float);
  const int stages = pick_stages(220 * 1024, CS_HDR + 128 + red, stage, q);
  if (stages < 2) return -1;
  const size_t smem = CS_HDR + 128 + (size_t)stages * stage + red;
  const size_t part_bytes = ((size_t)grid * 2 * D * sizeof(float) + 255) / 256 * 256;
  if (workspace_bytes < part_bytes + (size_t)grid * sizeof(int)) return -1;
  float* part = (float*)workspace;
  int* part_group = (int*)((char*)workspace + part_bytes);
#define B200_CB(NVC, KC)                                                                                            \
  {                                                                                                                 \
    static bool attr_set = false;                                                                                   \
    if (!attr_set) {                                                                                                \
      B200_CUDA(cudaFuncSetAttribute(combine_bwd_staged_kernel<NVC, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     227 * 1024));                                                                  \
      attr_set = true;                                                                                              \
    }                                                                                                               \
    launch_kernel(combine_bwd_staged_kernel<NVC, KC>, dim3(grid), dim3(CS_THREADS), smem, stream, dout, z, dest_row, \
                  w, mean, rstd, gamma, N, dz, d_w, part, part_group, stages);                                      \
  }
#define B200_CB_K(NVC) \
  if (K == 1) B200_CB(NVC, 1) else B200_CB(NVC, 2)
  switch (D / 256) {
    case 1: B200_CB_K(1) break;
    case 2: B200_CB_K(2) break;
    case 3: B200_CB_K(3) break;
    default: B200_CB_K(4) break;
  }
#undef B200_CB_K
#undef B200_CB
  B200_LAUNCH_CHECK("combine_bwd_staged_kernel");
  count_launch();
  return launch_ln_staged_reduce(part, part_group, grid, D, 1, 2, dgamma, dbeta, nullptr, stream);
}

}  // namespace b200
