// Expert-parallel dispatch / return fused with their collective: the permute kernels write token rows straight into
// the destination rank's receive buffer through NVLink peer mappings (16-byte st.global on peer pointers) — there
// is no separate all-to-all and no host-side split computation.  Buffers are symmetric allocations whose peer
// pointers come from torch.distributed._symmetric_memory; cross-rank ordering is a stream-ordered barrier between
// the phases (push counts | barrier | layout + dispatch | barrier | experts | return | barrier | combine).
#include "common.cuh"

namespace b200 {
namespace {

constexpr int EP_MAX_E = 64;       // experts handled by the single-block layout kernel (same bound as the router)

// every rank writes its per-global-expert pair counts into row `me` of every peer's count table [W, E]
__global__ void ep_push_counts_kernel(const int* __restrict__ counts, int* const* __restrict__ peer_tabs, int me, int W,
                                      int E) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W * E) {
    const int r = i / E, e = i % E;
    peer_tabs[r][me * E + e] = counts[e];
  }
}

// From the full table tab[s][e] (pairs rank s routes to global expert e) every rank derives, without talking to
// anybody, the layout of EVERY owner's grouped-GEMM input: expert segments in ascending expert order, each padded to
// 128 rows, rows inside a segment ordered by (source rank, token) — the canonical (expert, token) order of the
// unsharded layer on the concatenated batch.
//   send_base[e]  : row in owner(e)'s padded buffer where MY rows for global expert e start
//   pad_off2[le]  : my own padded segment offsets (local experts), pad_off2[El] = rows in use; pad_off2[El+1+le] =
//                   routed rows of local expert le (array of 2*El + 1 ints)
//   tile_group2[t]: local expert of my 128-row tile t (-1 beyond the rows in use)
//   row_home[i]   : home_rank * nk_cap + compact position at home of my padded row i (-1: padding)
__global__ void __launch_bounds__(1024)
ep_layout_kernel(const int* __restrict__ tab, int me, int W, int E, int El, int Rcap, int nk_cap,
                 int* __restrict__ send_base, int* __restrict__ pad_off2, int* __restrict__ tile_group2,
                 int* __restrict__ row_home) {
  pdl_trigger();
  pdl_wait();
  __shared__ int s_tot[EP_MAX_E], s_pad[EP_MAX_E + 1];
  __shared__ int s_src[EP_MAX_E * 8 + 8], s_home[EP_MAX_E * 8];   // [le][s] prefix inside the segment / position at home
  __shared__ int s_poff[EP_MAX_E + 1];
  const int t = threadIdx.x;
  if (t < E) {
    int tot = 0;
    for (int s = 0; s < W; ++s) tot += tab[s * E + t];
    s_tot[t] = tot;
  }
  __syncthreads();
  if (t < E) {      // padded offset of expert t inside its owner's buffer
    const int r = t / El;
    int p = 0;
    for (int e2 = r * El; e2 < t; ++e2) p += (s_tot[e2] + B200_GROUP_TILE - 1) / B200_GROUP_TILE * B200_GROUP_TILE;
    s_pad[t] = p;
    int before = 0;
    for (int s = 0; s < me; ++s) before += tab[s * E + t];
    if (blockIdx.x == 0) send_base[t] = p + before;
  }
  __syncthreads();
  if (t <= El) {
    int p;
    if (t < El) p = s_pad[me * El + t];
    else {
      const int last = me * El + El - 1;
      p = s_pad[last] + (s_tot[last] + B200_GROUP_TILE - 1) / B200_GROUP_TILE * B200_GROUP_TILE;
    }
    s_poff[t] = p;
    if (blockIdx.x == 0) {
      pad_off2[t] = p;
      if (t < El) pad_off2[El + 1 + t] = s_tot[me * El + t];    // routed rows of local expert t (behind the offsets)
    }
  }
  for (int i = t; i < El * W; i += blockDim.x) {      // (le, s): prefix of sources inside the segment, home position
    const int le = i / W, s = i % W, e = me * El + le;
    int pre = 0;
    for (int s2 = 0; s2 < s; ++s2) pre += tab[s2 * E + e];
    s_src[le * (W + 1) + s] = pre;
    if (s == W - 1) s_src[le * (W + 1) + W] = pre + tab[s * E + e];
    int home = 0;
    for (int e2 = 0; e2 < e; ++e2) home += tab[s * E + e2];
    s_home[le * W + s] = home;
  }
  __syncthreads();
  const int used = s_poff[El];
  const int tiles = Rcap / B200_GROUP_TILE;
  // every block derives the (tiny) tables above; the per-row maps are spread over the grid
  const int gt = blockIdx.x * blockDim.x + t, gstride = gridDim.x * blockDim.x;
  for (int tile = gt; tile < tiles; tile += gstride) {
    const int r = tile * B200_GROUP_TILE;
    int g = -1;
    if (r < used)
      for (int le = 0; le < El; ++le)
        if (r >= s_poff[le] && r < s_poff[le + 1]) g = le;
    tile_group2[tile] = g;
  }
  for (int i = gt; i < Rcap; i += gstride) {
    int v = -1;
    if (i < used) {
      int le = 0;
      while (le + 1 < El && s_poff[le + 1] <= i) ++le;
      const int j = i - s_poff[le];
      if (j < s_tot[me * El + le]) {
        int s = 0;
        while (s + 1 < W && s_src[le * (W + 1) + s + 1] <= j) ++s;
        const int pos = s_home[le * W + s] + (j - s_src[le * (W + 1) + s]);
        v = (pos < nk_cap) ? s * nk_cap + pos : -1;
      }
    }
    row_home[i] = v;
  }
}

// one warp per compact source row r (canonical expert order): copy it STRAIGHT into the owner's grouped-GEMM input
// (padded layout) through the peer mapping; the local padding rows of my own buffer are zeroed by the same kernel.
template <typename T>
__global__ void __launch_bounds__(256)
ep_dispatch_kernel(const T* __restrict__ src, const int* __restrict__ cmp_src, const int* __restrict__ cmp_off,
                   const int* __restrict__ send_base, const int* __restrict__ pad_off2, T* const* __restrict__ peer_bufs,
                   int me, int K, int E, int El, int D, int Rcap) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int total = cmp_off[E];
  const int nv = D / VT;
  for (int r = warp; r < total; r += nwarps) {
    int e = 0;
    while (e + 1 < E && cmp_off[e + 1] <= r) ++e;          // expert of this row (E <= 64)
    const int dest_rank = e / El;
    const int dest_row = send_base[e] + (r - cmp_off[e]);
    if (dest_row >= Rcap) continue;                         // beyond the receiver's capacity: dropped
    const int s = cmp_src != nullptr ? cmp_src[r] / K : r;
    const T* from = src + (long long)s * D;
    T* to = peer_bufs[dest_rank] + (long long)dest_row * D;   // peer-mapped address: the store crosses NVLink
    for (int v = lane; v < nv; v += 32)
      *reinterpret_cast<uint4*>(to + v * VT) = __ldg(reinterpret_cast<const uint4*>(from + v * VT));
  }
  // padding rows of my own segments (nobody else writes them): the rows behind the routed rows of every local expert
  // up to the next multiple of 128; pad_off2[El + 1 + le] holds the routed rows of local expert le
  T* mine = peer_bufs[me];
  for (int p = warp; p < El * B200_GROUP_TILE; p += nwarps) {
    const int le = p / B200_GROUP_TILE, j = p % B200_GROUP_TILE;
    const int row = pad_off2[le] + pad_off2[El + 1 + le] + j;
    if (row >= pad_off2[le + 1] || row >= Rcap) continue;
    T* to = mine + (long long)row * D;
    for (int v = lane; v < nv; v += 32) *reinterpret_cast<uint4*>(to + v * VT) = make_uint4(0, 0, 0, 0);
  }
}

// one warp per padded row i of my expert buffers: send it back to its home rank at its compact position
template <typename T>
__global__ void __launch_bounds__(256)
ep_return_kernel(const T* __restrict__ rows, const int* __restrict__ row_home, const int* __restrict__ pad_off2,
                 T* const* __restrict__ peer_rets, int El, int D, int Rcap, int nk_cap) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int total = min(pad_off2[El], Rcap);
  const int nv = D / VT;
  for (int i = warp; i < total; i += nwarps) {
    const int h = row_home[i];
    if (h < 0) continue;
    const int home_rank = h / nk_cap, home_row = h - home_rank * nk_cap;
    const T* from = rows + (long long)i * D;
    T* to = peer_rets[home_rank] + (long long)home_row * D;
    for (int v = lane; v < nv; v += 32)
      *reinterpret_cast<uint4*>(to + v * VT) = __ldg(reinterpret_cast<const uint4*>(from + v * VT));
  }
}

inline int ep_grid(int rows) {
  int blocks = (rows + 7) / 8;
  const int cap = num_sms() * 8;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : blocks;
}

// ---- data-parallel gradient all-reduce over peer memory (two-shot: reduce-scatter + all-gather in one kernel) ----
// Every rank owns a contiguous 1/W slice of the symmetric fp32 buffer: it loads that slice from all W ranks over
// NVLink (ranks summed in a fixed order, so every rank ends up with bit-identical values), scales, and stores the
// result into the same slice of every rank's buffer.  Cross-rank ordering is the caller's stream-ordered barrier
// before (all contributions written) and after (all stores landed).
constexpr int AR_MAX_W = 8;
struct PeerPtrs { float* p[AR_MAX_W]; };

__global__ void __launch_bounds__(512)
p2p_allreduce_kernel(const PeerPtrs peers, int me, int W, long long nvec, float scale) {
  pdl_trigger();
  pdl_wait();
  const long long per = (nvec + W - 1) / W;
  const long long v0 = (long long)me * per, v1 = min(nvec, v0 + per);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = v0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; v < v1; v += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < AR_MAX_W; ++r) {
      if (r < W) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(peers.p[r]) + v);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
    }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
#pragma unroll
    for (int r = 0; r < AR_MAX_W; ++r)
      if (r < W) __stcg(reinterpret_cast<float4*>(peers.p[r]) + v, acc);
  }
}

// ---- the same all-reduce through the NVSwitch (NVLS): `mc` is the MULTICAST address of the symmetric buffer.  One
// multimem.ld_reduce returns the sum over all ranks' copies of 4 floats, computed inside the switch; one multimem.st
// broadcasts the result to every rank's copy.  Every rank handles 1/W of the range, so per GPU the NVLink carries the
// buffer once in each direction, independent of W.
__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
constexpr int NVLS_UNROLL = 4;
__global__ void __launch_bounds__(512)
nvls_allreduce_kernel(float* mc, int me, int W, long long nvec, float scale) {
  pdl_trigger();
  pdl_wait();
  const long long per = (nvec + W - 1) / W;
  const long long v0 = (long long)me * per, v1 = min(nvec, v0 + per);
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long v = v0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; v + (NVLS_UNROLL - 1) * stride < v1; v += NVLS_UNROLL * stride) {
    float4 a[NVLS_UNROLL];
#pragma unroll
    for (int u = 0; u < NVLS_UNROLL; ++u) a[u] = mc_ld_reduce(mc + 4 * (v + u * stride));
#pragma unroll
    for (int u = 0; u < NVLS_UNROLL; ++u) {
      a[u].x *= scale; a[u].y *= scale; a[u].z *= scale; a[u].w *= scale;
      mc_st(mc + 4 * (v + u * stride), a[u]);
    }
  }
  for (; v < v1; v += stride) {
    float4 a = mc_ld_reduce(mc + 4 * v);
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    mc_st(mc + 4 * v, a);
  }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_ep_push_counts(const int32_t* counts, void* const* peer_tabs, int me, int W, int E, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(W > 0 && E > 0 && me >= 0 && me < W, "ep_push_counts: bad arguments");
  launch_kernel(ep_push_counts_kernel, dim3((W * E + 255) / 256), dim3(256), 0, stream, counts, (int* const*)peer_tabs,
                me, W, E);
  B200_LAUNCH_CHECK("ep_push_counts_kernel");
  count_launch();
  return 0;
}

int b200_ep_layout(const int32_t* tab, int me, int W, int E, int Rcap, int nk_cap, int32_t* send_base,
                   int32_t* pad_off2, int32_t* tile_group2, int32_t* row_home, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(W > 0 && W <= 8 && E > 0 && E % W == 0 && E <= EP_MAX_E && Rcap > 0 && Rcap % B200_GROUP_TILE == 0 &&
                     nk_cap > 0 && (long long)W * nk_cap < 2147483647ll,
                 "ep_layout: bad arguments (W=%d E=%d Rcap=%d nk_cap=%d; W<=8, E<=%d, E %% W == 0)", W, E, Rcap,
                 nk_cap, EP_MAX_E);
  int blocks = (Rcap + 4095) / 4096;        // ~4 rows per thread
  if (blocks > 64) blocks = 64;
  launch_kernel(ep_layout_kernel, dim3(blocks), dim3(1024), 0, stream, tab, me, W, E, E / W, Rcap, nk_cap, send_base,
                pad_off2, tile_group2, row_home);
  B200_LAUNCH_CHECK("ep_layout_kernel");
  count_launch();
  return 0;
}

int b200_ep_dispatch(const void* src, const int32_t* cmp_src, const int32_t* cmp_off, const int32_t* send_base,
                     const int32_t* pad_off2, void* const* peer_bufs, int me, int K, int NK, int E, int El, int D,
                     int Rcap, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(NK > 0 && E > 0 && El > 0 && D % (dtype == B200_BF16 ? 8 : 4) == 0, "ep_dispatch: bad arguments");
  if (dtype == B200_BF16)
    launch_kernel(ep_dispatch_kernel<bf16>, dim3(ep_grid(NK)), dim3(256), 0, stream, (const bf16*)src, cmp_src, cmp_off,
                  send_base, pad_off2, (bf16* const*)peer_bufs, me, K, E, El, D, Rcap);
  else
    launch_kernel(ep_dispatch_kernel<float>, dim3(ep_grid(NK)), dim3(256), 0, stream, (const float*)src, cmp_src,
                  cmp_off, send_base, pad_off2, (float* const*)peer_bufs, me, K, E, El, D, Rcap);
  B200_LAUNCH_CHECK("ep_dispatch_kernel");
  count_launch();
  return 0;
}

int b200_ep_return(const void* rows, const int32_t* row_home, const int32_t* pad_off2, void* const* peer_rets,
                   int El, int D, int Rcap, int nk_cap, int rows_hint, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(El > 0 && Rcap > 0 && nk_cap > 0 && D % (dtype == B200_BF16 ? 8 : 4) == 0, "ep_return: bad arguments");
  const int grid = ep_grid(rows_hint > 0 && rows_hint < Rcap ? rows_hint : Rcap);
  if (dtype == B200_BF16)
    launch_kernel(ep_return_kernel<bf16>, dim3(grid), dim3(256), 0, stream, (const bf16*)rows, row_home, pad_off2,
                  (bf16* const*)peer_rets, El, D, Rcap, nk_cap);
  else
    launch_kernel(ep_return_kernel<float>, dim3(grid), dim3(256), 0, stream, (const float*)rows, row_home, pad_off2,
                  (float* const*)peer_rets, El, D, Rcap, nk_cap);
  B200_LAUNCH_CHECK("ep_return_kernel");
  count_launch();
  return 0;
}

int b200_p2p_allreduce_f32(const unsigned long long* peer_bufs_host, int me, int W, long long offset,
                           long long count, float scale, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(peer_bufs_host != nullptr && W >= 1 && W <= AR_MAX_W && me >= 0 && me < W,
                 "p2p_allreduce: bad world (W=%d, me=%d; at most %d ranks)", W, me, AR_MAX_W);
  B200_CHECK_ARG(count > 0 && count % 4 == 0 && offset >= 0 && offset % 4 == 0,
                 "p2p_allreduce: offset / count must be multiples of 4 floats");
  PeerPtrs pp{};
  for (int r = 0; r < W; ++r) {
    B200_CHECK_ARG((peer_bufs_host[r] & 15ull) == 0, "p2p_allreduce: peer buffers must be 16-byte aligned");
    pp.p[r] = reinterpret_cast<float*>(peer_bufs_host[r]) + offset;
  }
  const long long nvec = count / 4;
  long long blocks = (nvec / W + 511) / 512;
  const long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  launch_kernel(p2p_allreduce_kernel, dim3((unsigned)blocks), dim3(512), 0, stream, pp, me, W, nvec, scale);
  B200_LAUNCH_CHECK("p2p_allreduce_kernel");
  count_launch();
  return 0;
}

int b200_nvls_allreduce_f32(void* multicast_ptr, int me, int W, long long offset, long long count, float scale,
                            int max_blocks, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(multicast_ptr != nullptr && W >= 1 && me >= 0 && me < W, "nvls_allreduce: bad arguments");
  B200_CHECK_ARG(count > 0 && count % 4 == 0 && offset >= 0 && offset % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) == 0,
                 "nvls_allreduce: offset / count must be multiples of 4 floats, the buffer 16-byte aligned");
  const long long nvec = count / 4;
  long long blocks = (nvec / W + 512 * NVLS_UNROLL - 1) / (512 * NVLS_UNROLL);
  const long long cap = max_blocks > 0 ? max_blocks : 64;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  launch_kernel(nvls_allreduce_kernel, dim3((unsigned)blocks), dim3(512), 0, stream,
                reinterpret_cast<float*>(multicast_ptr) + offset, me, W, nvec, scale);
  B200_LAUNCH_CHECK("nvls_allreduce_kernel");
  count_launch();
  return 0;
}

}  // extern "C"
