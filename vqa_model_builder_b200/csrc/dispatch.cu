// MOE dispatch / combine: routing plan (histogram -> scan -> stable scatter), capacity masking, row
// permute / un-permute with 128-bit accesses, weighted combine fused with the output LayerNorm (fwd + bwd).
// Everything stays on the device: no counts are ever read back by the host.
#include "rowops.cuh"

namespace b200 {

// combine_staged.cu: persistent shared-memory-staged bf16 kernels for top-k <= 2 (return -1 when not covered)
size_t combine_bwd_staged_ws(int D);
int launch_combine_fwd_staged(const bf16* z, const int* dest_row, const float* w, const float* gamma,
                              const float* beta, float eps, int N, int K, int D, bf16* out, float* mean, float* rstd,
                              cudaStream_t stream);
int launch_combine_bwd_staged(const bf16* dout, const bf16* z, const int* dest_row, const float* w, const float* mean,
                              const float* rstd, const float* gamma, int N, int K, int D, bf16* dz, float* d_w,
                              float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                              cudaStream_t stream);

int launch_ln_param_reduce(const float* part, int blocks, int rows_per_block, int D, const int* tile_group, int G,
                           float* dgamma, float* dbeta, cudaStream_t stream, float* dcol = nullptr);

namespace {

constexpr int PLAN_CHUNK = 1024;  // (token,slot) pairs per block
constexpr int MAX_E = 64;

// ---- plan stage 1: per-chunk expert histogram ------------------------------------------------------
__global__ void __launch_bounds__(PLAN_CHUNK)
plan_count_kernel(const int* __restrict__ idx, int NK, int E, int* __restrict__ block_counts) {
  pdl_trigger();
  pdl_wait();
  __shared__ int hist[MAX_E];
  if (threadIdx.x < MAX_E) hist[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * PLAN_CHUNK + threadIdx.x;
  if (i < NK) {
    const int e = idx[i];
    if (e >= 0 && e < E) atomicAdd(&hist[e], 1);
  }
  __syncthreads();
  if (threadIdx.x < E) block_counts[blockIdx.x * E + threadIdx.x] = hist[threadIdx.x];
}

// ---- plan stage 2 (one block): exclusive scan over chunks, offsets, tile map ------------------------
__global__ void __launch_bounds__(1024)
plan_scan_kernel(int* __restrict__ block_counts /* in: counts, out: exclusive bases */, int chunks, int E, int Rmax,
                 int* __restrict__ counts, int* __restrict__ cmp_off, int* __restrict__ pad_off,
                 int* __restrict__ tile_group) {
  pdl_trigger();
  pdl_wait();
  __shared__ int s_pad[MAX_E + 1];
  const int t = threadIdx.x;
  if (t < E) {
    int run = 0;
    for (int b = 0; b < chunks; ++b) {
      const int c = block_counts[b * E + t];
      block_counts[b * E + t] = run;
      run += c;
    }
    counts[t] = run;
  }
  __syncthreads();
  if (t == 0) {
    int c = 0, p = 0;
    for (int e = 0; e < E; ++e) {
      cmp_off[e] = c;
      pad_off[e] = p;
      s_pad[e] = p;
      c += counts[e];
      p += (counts[e] + B200_GROUP_TILE - 1) / B200_GROUP_TILE * B200_GROUP_TILE;
    }
    cmp_off[E] = c;
    pad_off[E] = p;
    s_pad[E] = p;
  }
  __syncthreads();
  const int tiles = Rmax / B200_GROUP_TILE;
  for (int tile = t; tile < tiles; tile += blockDim.x) {
    const int r = tile * B200_GROUP_TILE;
    int g = -1;
    for (int e = 0; e < E; ++e)
      if (r >= s_pad[e] && r < s_pad[e + 1]) g = e;
    tile_group[tile] = g;
  }
}

// ---- plan stage 3: stable rank inside the chunk, scatter the maps ------------------------------------
__global__ void __launch_bounds__(PLAN_CHUNK)
plan_scatter_kernel(const int* __restrict__ idx, int NK, int E, const int* __restrict__ block_base,
                    const int* __restrict__ cmp_off, const int* __restrict__ pad_off, int* __restrict__ dest_row,
                    int* __restrict__ cmp_pos, int* __restrict__ row_src, int* __restrict__ cmp_src) {
  pdl_trigger();
  pdl_wait();
  __shared__ int warp_hist[PLAN_CHUNK / 32][MAX_E];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = threadIdx.x; j < (PLAN_CHUNK / 32) * MAX_E; j += blockDim.x) (&warp_hist[0][0])[j] = 0;
  __syncthreads();
  const int i = blockIdx.x * PLAN_CHUNK + threadIdx.x;
  int e = -1;
  if (i < NK) {
    e = idx[i];
    if (e < 0 || e >= E) e = -1;
  }
  // peers = lanes of this warp routed to the same expert; lower lanes come first (stable)
  const unsigned peers = __match_any_sync(0xffffffffu, e >= 0 ? e : (MAX_E + lane));
  const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
  if (e >= 0 && rank_in_warp == 0) warp_hist[warp][e] = __popc(peers);
  __syncthreads();
  if (i < NK) {
    if (e >= 0) {
      int before = 0;
      for (int w = 0; w < warp; ++w) before += warp_hist[w][e];
      const int k = block_base[blockIdx.x * E + e] + before + rank_in_warp;
      const int d = pad_off[e] + k;
      dest_row[i] = d;
      cmp_pos[i] = cmp_off[e] + k;
      row_src[d] = i;
      if (cmp_src != nullptr) cmp_src[cmp_off[e] + k] = i;
    } else {
      dest_row[i] = -1;
      cmp_pos[i] = -1;
    }
  }
}

// ---- the whole plan in one block when every (token, slot) pair fits one chunk (NK <= 1024: the classification
// pipelines route B pooled vectors) -- histogram, offsets, tile map, stable scatter and the row_src fill that the
// three-kernel path does with a memset: one launch instead of three kernels and a memset node on the forward chain.
__global__ void __launch_bounds__(PLAN_CHUNK)
plan_small_kernel(const int* __restrict__ idx, int NK, int E, int Rmax, int* __restrict__ counts,
                  int* __restrict__ cmp_off, int* __restrict__ pad_off, int* __restrict__ tile_group,
                  int* __restrict__ dest_row, int* __restrict__ cmp_pos, int* __restrict__ row_src,
                  int* __restrict__ cmp_src) {
  pdl_trigger();
  pdl_wait();
  __shared__ int warp_hist[PLAN_CHUNK / 32][MAX_E];
  __shared__ int s_cmp[MAX_E + 1], s_pad[MAX_E + 1];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int j = t; j < (PLAN_CHUNK / 32) * MAX_E; j += blockDim.x) (&warp_hist[0][0])[j] = 0;
  for (int r = t; r < Rmax; r += blockDim.x) row_src[r] = -1;
  if (cmp_src != nullptr && t < NK) cmp_src[t] = -1;
  __syncthreads();
  int e = -1;
  if (t < NK) {
    e = idx[t];
    if (e < 0 || e >= E) e = -1;
  }
  const unsigned peers = __match_any_sync(0xffffffffu, e >= 0 ? e : (MAX_E + lane));
  const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
  if (e >= 0 && rank_in_warp == 0) warp_hist[warp][e] = __popc(peers);
  __syncthreads();
  if (t == 0) {
    int c = 0, p = 0;
    for (int x = 0; x < E; ++x) {
      int n = 0;
      for (int w = 0; w < PLAN_CHUNK / 32; ++w) n += warp_hist[w][x];
      counts[x] = n;
      cmp_off[x] = c;
      pad_off[x] = p;
      s_cmp[x] = c;
      s_pad[x] = p;
      c += n;
      p += (n + B200_GROUP_TILE - 1) / B200_GROUP_TILE * B200_GROUP_TILE;
    }
    cmp_off[E] = c;
    pad_off[E] = p;
    s_cmp[E] = c;
    s_pad[E] = p;
  }
  __syncthreads();
  const int tiles = Rmax / B200_GROUP_TILE;
  for (int tile = t; tile < tiles; tile += blockDim.x) {
    const int r = tile * B200_GROUP_TILE;
    int g = -1;
    for (int x = 0; x < E; ++x)
      if (r >= s_pad[x] && r < s_pad[x + 1]) g = x;
    tile_group[tile] = g;
  }
  if (t < NK) {
    if (e >= 0) {
      int before = 0;
      for (int w = 0; w < warp; ++w) before += warp_hist[w][e];
      const int k = before + rank_in_warp;
      const int d = s_pad[e] + k;
      dest_row[t] = d;
      cmp_pos[t] = s_cmp[e] + k;
      row_src[d] = t;
      if (cmp_src != nullptr) cmp_src[s_cmp[e] + k] = t;
    } else {
      dest_row[t] = -1;
      cmp_pos[t] = -1;
    }
  }
}

// ---- capacity masking (SparseMOELayer) -----------------------------------------------------------------
__global__ void capacity_init_kernel(const float* __restrict__ w, int NK, float* __restrict__ w_eff,
                                     uint8_t* __restrict__ keep) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < NK) {
    w_eff[i] = w[i];
    keep[i] = 1;
  }
}
__global__ void __launch_bounds__(1024)
capacity_kernel(const float* __restrict__ w, const int* __restrict__ counts, const int* __restrict__ pad_off,
                const int* __restrict__ row_src, int capacity, float* __restrict__ w_eff, uint8_t* __restrict__ keep) {
  pdl_trigger();
  pdl_wait();
  const int e = blockIdx.x;
  const int c = counts[e];
  if (c <= capacity) return;
  const int r0 = pad_off[e];
  for (int a = threadIdx.x; a < c; a += blockDim.x) {
    const int sa = row_src[r0 + a];
    const float wa = w[sa];
    int rank = 0;
    for (int b = 0; b < c; ++b) {
      const float wb = w[row_src[r0 + b]];
      rank += (wb > wa) || (wb == wa && b < a);
    }
    if (rank >= capacity) {
      w_eff[sa] = 0.f;
      keep[sa] = 0;
    }
  }
}

// ---- permute / un-permute --------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
permute_kernel(const T* __restrict__ x, const int* __restrict__ row_src, const int* __restrict__ pad_off, int E, int K,
               int Rmax, int D, T* __restrict__ xp) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int rows = min(Rmax, pad_off[E]);
  const int nv = D / VT;
  for (int r = warp; r < rows; r += nwarps) {
    const int src = row_src[r];
    T* dst = xp + (long long)r * D;
    if (src >= 0) {
      const T* s = x + (long long)(src / K) * D;
      for (int v = lane; v < nv; v += 32)
        *reinterpret_cast<uint4*>(dst + v * VT) = __ldg(reinterpret_cast<const uint4*>(s + v * VT));
    } else {
      for (int v = lane; v < nv; v += 32) *reinterpret_cast<uint4*>(dst + v * VT) = make_uint4(0, 0, 0, 0);
    }
  }
}

template <typename T, int NV>
__global__ void __launch_bounds__(256)
unpermute_kernel(const T* __restrict__ dxp, const int* __restrict__ dest_row, const T* __restrict__ add, int N, int K,
                 int D, T* __restrict__ dx) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int n = warp; n < N; n += nwarps) {
    RowRegs<T, NV> acc;
    if (add != nullptr) acc.load(add + (long long)n * D, D, lane);
    else acc.zero();
    for (int k = 0; k < K; ++k) {
      const int d = dest_row[n * K + k];
      if (d >= 0) acc.axpy(dxp + (long long)d * D, 1.f, D, lane);
    }
    acc.store(dx + (long long)n * D, D, lane);
  }
}

// ---- combine + output LayerNorm ------------------------------------------------------------------------------
template <typename T, int NV>
__global__ void __launch_bounds__(256)
combine_fwd_kernel(const T* __restrict__ z, const int* __restrict__ dest_row, const float* __restrict__ w,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int N, int K, int D,
                   T* __restrict__ out, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nv = D / VT;
  for (int n = warp; n < N; n += nwarps) {
    RowRegs<T, NV> acc;
    acc.zero();
    for (int k = 0; k < K; ++k) {
      const int d = dest_row[n * K + k];
      const float wk = w[n * K + k];
      if (d >= 0 && wk != 0.f) acc.axpy(z + (long long)d * D, wk, D, lane);
    }
    if (gamma == nullptr) {      // plain weighted sum (HierarchicalMOE applies output_proj before its LayerNorm)
      acc.store(out + (long long)n * D, D, lane);
      continue;
    }
    const float mean = acc.sum(D, lane) / D;
    const float var = acc.sumsq_centered(mean, D, lane) / D;
    const float rstd = rsqrtf(var + eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nv) {
        float gv[VT], bv[VT];
        load_param<VT>(gamma, vi, gv);
        load_param<VT>(beta, vi, bv);
#pragma unroll
        for (int u = 0; u < VT; ++u) acc.v[j][u] = (acc.v[j][u] - mean) * rstd * gv[u] + bv[u];
      }
    }
    acc.store(out + (long long)n * D, D, lane);
    if (lane == 0) {
      mean_out[n] = mean;
      rstd_out[n] = rstd;
    }
  }
}

constexpr int CMB_WARPS = 8;
constexpr int CMB_TOKENS_PER_BLOCK = 8;  // workspace bound (tokens per block is 8 * tpw, tpw >= 1)

// Each warp walks `tpw` tokens keeping the output_norm dgamma/dbeta partials in registers.
template <typename T, int NV>
__global__ void __launch_bounds__(CMB_WARPS * 32)
combine_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ z, const int* __restrict__ dest_row,
                   const float* __restrict__ w, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                   const float* __restrict__ gamma, int N, int K, int D, T* __restrict__ dz, float* __restrict__ d_w,
                   float* __restrict__ part, int tpw) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  extern __shared__ float red[];  // [CMB_WARPS][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = D / VT;
  float* acc_g = red + (warp * 2 + 0) * D;   // per-warp dgamma / dbeta accumulators in shared memory
  float* acc_b = red + (warp * 2 + 1) * D;
  for (int d = lane; d < D; d += 32) { acc_g[d] = 0.f; acc_b[d] = 0.f; }
  __syncwarp();
  const int n0 = blockIdx.x * (CMB_WARPS * tpw);
  for (int i = 0; i < tpw; ++i) {
    const int n = n0 + i * CMB_WARPS + warp;
    if (n >= N) break;
    RowRegs<T, NV> s, g;
    s.zero();
    for (int k = 0; k < K; ++k) {
      const int d = dest_row[n * K + k];
      const float wk = w[n * K + k];
      if (d >= 0 && wk != 0.f) s.axpy(z + (long long)d * D, wk, D, lane);
    }
    g.load(dout + (long long)n * D, D, lane);
    const bool norm = gamma != nullptr;
    const float mean = norm ? mean_in[n] : 0.f, rstd = norm ? rstd_in[n] : 1.f;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nv && norm) {
        float gv[VT];
        load_param<VT>(gamma, vi, gv);
        float* ag = acc_g + vi * VT;
        float* ab = acc_b + vi * VT;
#pragma unroll
        for (int u = 0; u < VT; u += 4) {      // 16-byte read-modify-write of the per-warp accumulators
          float4 tg = *reinterpret_cast<float4*>(ag + u), tb = *reinterpret_cast<float4*>(ab + u);
          float* pg = reinterpret_cast<float*>(&tg);
          float* pb = reinterpret_cast<float*>(&tb);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float xhat = (s.v[j][u + q] - mean) * rstd;
            const float d = g.v[j][u + q];
            pg[q] = fmaf(d, xhat, pg[q]);
            pb[q] += d;
            const float gg = d * gv[u + q];
            s.v[j][u + q] = xhat;
            g.v[j][u + q] = gg;
            s1 += gg;
            s2 = fmaf(gg, xhat, s2);
          }
          *reinterpret_cast<float4*>(ag + u) = tg;
          *reinterpret_cast<float4*>(ab + u) = tb;
        }
      }
    }
    if (norm) {
      s1 = warp_sum(s1) / D;
      s2 = warp_sum(s2) / D;
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int u = 0; u < VT; ++u) g.v[j][u] = rstd * (g.v[j][u] - s1 - s.v[j][u] * s2);  // ds
    }                                                                                      // else ds = dout
    for (int k = 0; k < K; ++k) {
      const int d = dest_row[n * K + k];
      const float wk = w[n * K + k];
      float dot = 0.f;
      if (d >= 0) {
        const T* zr = z + (long long)d * D;
        T* dzr = dz + (long long)d * D;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int vi = lane + 32 * j;
          if (vi < nv) {
            Vec16<T> zv, o;
            zv.load(zr + vi * VT);
#pragma unroll
            for (int u = 0; u < VT; ++u) {
              dot = fmaf(g.v[j][u], zv.v[u], dot);
              o.v[u] = wk * g.v[j][u];
            }
            o.store(dzr + vi * VT);
          }
        }
      }
      dot = warp_sum(dot);
      if (lane == 0) d_w[n * K + k] = (d >= 0) ? dot : 0.f;
    }
  }

  __syncthreads();
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
    const int which = c / D, d = c % D;
    float sum = 0.f;
#pragma unroll
    for (int wp = 0; wp < CMB_WARPS; ++wp) sum += red[(wp * 2 + which) * D + d];
    part[((long long)blockIdx.x * 2 + which) * D + d] = sum;
  }
}

// rows whose combine weight was zeroed (capacity) or that are padding get a zero gradient
template <typename T>
__global__ void __launch_bounds__(256)
zero_unwritten_rows_kernel(const int* __restrict__ row_src, const float* __restrict__ w, int Rmax, int D,
                           T* __restrict__ dz) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nv = D / VT;
  for (int r = warp; r < Rmax; r += nwarps) {
    const int src = row_src[r];
    if (src < 0) {
      T* dst = dz + (long long)r * D;
      for (int v = lane; v < nv; v += 32) *reinterpret_cast<uint4*>(dst + v * VT) = make_uint4(0, 0, 0, 0);
    }
  }
}

inline int row_grid(int rows) {
  int blocks = (rows + 7) / 8;  // 8 warps per 256-thread block
  const int cap = num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return blocks;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_moe_max_rows(int NK, int E) {
  // every expert segment is padded to a multiple of 128 rows
  const long long r = ((long long)NK + (long long)E * (B200_GROUP_TILE - 1) + B200_GROUP_TILE - 1) /
                      B200_GROUP_TILE * B200_GROUP_TILE;
  return (int)(r < B200_GROUP_TILE ? B200_GROUP_TILE : r);
}

size_t b200_moe_plan_ws(int NK, int E) {
  const size_t chunks = (size_t)(NK + PLAN_CHUNK - 1) / PLAN_CHUNK;
  return chunks * (size_t)E * sizeof(int);
}

int b200_moe_plan(const int32_t* idx, int NK, int E, int Rmax, int32_t* counts, int32_t* cmp_off, int32_t* pad_off,
                  int32_t* dest_row, int32_t* cmp_pos, int32_t* row_src, int32_t* tile_group, int32_t* cmp_src,
                  void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(NK > 0 && E > 0 && E <= MAX_E, "moe_plan: need NK>0 and 0<E<=%d (NK=%d E=%d)", MAX_E, NK, E);
  B200_CHECK_ARG(Rmax % B200_GROUP_TILE == 0 && Rmax >= b200_moe_max_rows(NK, E),
                 "moe_plan: Rmax=%d must be a multiple of 128 and >= %d", Rmax, b200_moe_max_rows(NK, E));
  B200_CHECK_ARG(workspace_bytes >= b200_moe_plan_ws(NK, E), "moe_plan: workspace too small");
  const int chunks = (NK + PLAN_CHUNK - 1) / PLAN_CHUNK;
  int* block_counts = (int*)workspace;
  if (chunks == 1) {
    launch_kernel(plan_small_kernel, dim3(1), dim3(PLAN_CHUNK), 0, stream, idx, NK, E, Rmax, counts, cmp_off, pad_off,
                  tile_group, dest_row, cmp_pos, row_src, cmp_src);
    B200_LAUNCH_CHECK("plan_small_kernel");
    count_launch(1);
    return 0;
  }
  B200_CUDA(cudaMemsetAsync(row_src, 0xFF, (size_t)Rmax * sizeof(int), stream));
  if (cmp_src != nullptr) B200_CUDA(cudaMemsetAsync(cmp_src, 0xFF, (size_t)NK * sizeof(int), stream));
  launch_kernel(plan_count_kernel, dim3(chunks), dim3(PLAN_CHUNK), 0, stream, idx, NK, E, block_counts);
  B200_LAUNCH_CHECK("plan_count_kernel");
  launch_kernel(plan_scan_kernel, dim3(1), dim3(1024), 0, stream, block_counts, chunks, E, Rmax, counts, cmp_off, pad_off, tile_group);
  B200_LAUNCH_CHECK("plan_scan_kernel");
  launch_kernel(plan_scatter_kernel, dim3(chunks), dim3(PLAN_CHUNK), 0, stream, idx, NK, E, block_counts, cmp_off, pad_off, dest_row, cmp_pos,
                                                         row_src, cmp_src);
  B200_LAUNCH_CHECK("plan_scatter_kernel");
  count_launch(3);
  return 0;
}

int b200_moe_capacity(const int32_t* idx, const float* w, const int32_t* counts, const int32_t* pad_off,
                      const int32_t* row_src, int NK, int E, int capacity, float* w_eff, uint8_t* keep, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  (void)idx;
  B200_CHECK_ARG(NK > 0 && E > 0 && capacity >= 0, "moe_capacity: bad arguments");
  launch_kernel(capacity_init_kernel, dim3((NK + 255) / 256), dim3(256), 0, stream, w, NK, w_eff, keep);
  B200_LAUNCH_CHECK("capacity_init_kernel");
  launch_kernel(capacity_kernel, dim3(E), dim3(1024), 0, stream, w, counts, pad_off, row_src, capacity, w_eff, keep);
  B200_LAUNCH_CHECK("capacity_kernel");
  count_launch(2);
  return 0;
}

#define B200_ROW_DISPATCH(dtype, D, name)                                                                   \
  B200_CHECK_ARG((dtype) == B200_BF16 ? RowRegs<bf16>::supported(D) : RowRegs<float>::supported(D),         \
                 name ": D=%d unsupported (bf16: D%%8==0 && D<=2048; fp32: D%%4==0 && D<=1024)", D)

int b200_moe_permute(const void* x, const int32_t* row_src, const int32_t* pad_off, int E, int K, int Rmax, int D,
                     int dtype, void* xp, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_ROW_DISPATCH(dtype, D, "moe_permute");
  const int blocks = row_grid(Rmax);
  if (dtype == B200_BF16)
    launch_kernel(permute_kernel<bf16>, dim3(blocks), dim3(256), 0, stream, (const bf16*)x, row_src, pad_off, E, K, Rmax, D, (bf16*)xp);
  else
    launch_kernel(permute_kernel<float>, dim3(blocks), dim3(256), 0, stream, (const float*)x, row_src, pad_off, E, K, Rmax, D, (float*)xp);
  B200_LAUNCH_CHECK("permute_kernel");
  count_launch();
  return 0;
}

int b200_moe_unpermute(const void* dxp, const int32_t* dest_row, const void* add, int N, int K, int D, int dtype,
                       void* dx, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_ROW_DISPATCH(dtype, D, "moe_unpermute");
  const int blocks = row_grid(N);
  if (dtype == B200_BF16) {
    B200_NV_SWITCH(row_nv<bf16>(D), launch_kernel(unpermute_kernel<bf16, NV>, dim3(blocks), dim3(256), 0, stream, 
        (const bf16*)dxp, dest_row, (const bf16*)add, N, K, D, (bf16*)dx));
  } else {
    B200_NV_SWITCH(row_nv<float>(D), launch_kernel(unpermute_kernel<float, NV>, dim3(blocks), dim3(256), 0, stream, 
        (const float*)dxp, dest_row, (const float*)add, N, K, D, (float*)dx));
  }
  B200_LAUNCH_CHECK("unpermute_kernel");
  count_launch();
  return 0;
}

int b200_moe_combine_fwd(const void* z, const int32_t* dest_row, const float* w, const float* gamma, const float* beta,
                         float eps, int N, int K, int D, int dtype, void* out, float* mean, float* rstd,
                         void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_ROW_DISPATCH(dtype, D, "moe_combine_fwd");
  if (dtype == B200_BF16) {
    const int rc = launch_combine_fwd_staged((const bf16*)z, dest_row, w, gamma, beta, eps, N, K, D, (bf16*)out, mean, rstd,
                                             stream);
    if (rc >= 0) return rc;
  }
  const int blocks = row_grid(N);
  if (dtype == B200_BF16) {
    B200_NV_SWITCH(row_nv<bf16>(D), launch_kernel(combine_fwd_kernel<bf16, NV>, dim3(blocks), dim3(256), 0, stream, 
        (const bf16*)z, dest_row, w, gamma, beta, eps, N, K, D, (bf16*)out, mean, rstd));
  } else {
    B200_NV_SWITCH(row_nv<float>(D), launch_kernel(combine_fwd_kernel<float, NV>, dim3(blocks), dim3(256), 0, stream, 
        (const float*)z, dest_row, w, gamma, beta, eps, N, K, D, (float*)out, mean, rstd));
  }
  B200_LAUNCH_CHECK("combine_fwd_kernel");
  count_launch();
  return 0;
}

size_t b200_moe_combine_bwd_ws(int N, int D) {
  const size_t blocks = (size_t)(N + CMB_TOKENS_PER_BLOCK - 1) / CMB_TOKENS_PER_BLOCK;
  const size_t a = blocks * 2 * (size_t)D * sizeof(float), b = combine_bwd_staged_ws(D);
  return a > b ? a : b;
}

int b200_moe_combine_bwd(const void* dout, const void* z, const int32_t* dest_row, const float* w, const float* mean,
                         const float* rstd, const float* gamma, const int32_t* row_src, int N, int K, int D, int Rmax,
                         int dtype, void* dz, float* d_w, float* dgamma, float* dbeta, void* workspace,
                         size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_ROW_DISPATCH(dtype, D, "moe_combine_bwd");
  B200_CHECK_ARG(workspace_bytes >= b200_moe_combine_bwd_ws(N, D), "moe_combine_bwd: workspace too small");
  int tpw = 1;
  while (tpw < 16 && (N + CMB_WARPS * tpw * 2 - 1) / (CMB_WARPS * tpw * 2) >= 2 * num_sms()) tpw *= 2;
  const int tokens_per_block = CMB_WARPS * tpw;
  const int blocks = (N + tokens_per_block - 1) / tokens_per_block;
  const size_t smem = (size_t)CMB_WARPS * 2 * D * sizeof(float);
  B200_CHECK_ARG(smem <= 160 * 1024, "moe_combine_bwd: D=%d too large", D);
  float* part = (float*)workspace;
  const int zb = row_grid(Rmax);
  if (dtype == B200_BF16) {
    launch_kernel(zero_unwritten_rows_kernel<bf16>, dim3(zb), dim3(256), 0, stream, row_src, w, Rmax, D, (bf16*)dz);
    const int rc = launch_combine_bwd_staged((const bf16*)dout, (const bf16*)z, dest_row, w, mean, rstd, gamma, N, K, D,
                                             (bf16*)dz, d_w, dgamma, dbeta, workspace, workspace_bytes, stream);
    if (rc >= 0) {
      if (rc == 0) count_launch();     // the zeroing kernel above
      return rc;
    }
    B200_NV_SWITCH(row_nv<bf16>(D), {
      if (smem > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(combine_bwd_kernel<bf16, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      launch_kernel(combine_bwd_kernel<bf16, NV>, dim3(blocks), dim3(CMB_WARPS * 32), smem, stream, 
          (const bf16*)dout, (const bf16*)z, dest_row, w, mean, rstd, gamma, N, K, D, (bf16*)dz, d_w, part, tpw);
    });
  } else {
    launch_kernel(zero_unwritten_rows_kernel<float>, dim3(zb), dim3(256), 0, stream, row_src, w, Rmax, D, (float*)dz);
    B200_NV_SWITCH(row_nv<float>(D), {
      if (smem > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(combine_bwd_kernel<float, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      launch_kernel(combine_bwd_kernel<float, NV>, dim3(blocks), dim3(CMB_WARPS * 32), smem, stream, 
          (const float*)dout, (const float*)z, dest_row, w, mean, rstd, gamma, N, K, D, (float*)dz, d_w, part, tpw);
    });
  }
  B200_LAUNCH_CHECK("combine_bwd_kernel");
  count_launch(2);
  if (gamma == nullptr) return 0;       // plain weighted sum: no LayerNorm parameters
  return launch_ln_param_reduce(part, blocks, tokens_per_block, D, nullptr, 1, dgamma, dbeta, stream);
}

}  // extern "C"
