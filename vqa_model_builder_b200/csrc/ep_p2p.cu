// Expert-parallel dispatch / return fused with their collective: the permute kernels write token rows straight into
// the destination rank's receive buffer through NVLink peer mappings (16-byte st.global on peer pointers) — there
// is no separate all-to-all and no host-side split computation.  Buffers are symmetric allocations whose peer
// pointers come from torch.distributed._symmetric_memory; cross-rank ordering is a stream-ordered barrier between
// the phases (push counts | barrier | layout + dispatch | barrier | experts | return | barrier | combine).
#include "common.cuh"

namespace b200 {
namespace {

constexpr int EP_MAX_SEG = 1024;   // W * E_local segments handled by the single-block layout kernel

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

// From the full table tab[s][e] (pairs rank s routes to global expert e), one block derives
//   send_off[e]   : row in the owner's receive buffer where MY block for expert e starts  (order: [source][local e])
//   seg_off[g]    : prefix over my receive segments g = s*El + le   (seg_off[W*El] = rows I receive)
//   home_off[g]   : position of segment g inside source s's compact (expert-sorted) order
//   idx_recv[i]   : local expert of received row i, -1 beyond the received rows (feeds b200_moe_plan)
__global__ void __launch_bounds__(1024)
ep_layout_kernel(const int* __restrict__ tab, int me, int W, int E, int El, int cap, int* __restrict__ send_off,
                 int* __restrict__ seg_off, int* __restrict__ home_off, int* __restrict__ idx_recv) {
  pdl_trigger();
  pdl_wait();
  __shared__ int s_seg[EP_MAX_SEG + 1];
  const int t = threadIdx.x;
  const int nseg = W * El;
  if (t < E) {   // send offsets: one thread per global expert
    const int r = t / El, le = t % El;
    int off = 0;
    for (int s = 0; s < me; ++s)
      for (int l2 = 0; l2 < El; ++l2) off += tab[s * E + r * El + l2];
    for (int l2 = 0; l2 < le; ++l2) off += tab[me * E + r * El + l2];
    send_off[t] = off;
  }
  if (t < nseg) {   // home offsets: prefix of the source's counts up to my expert
    const int s = t / El, le = t % El;
    int off = 0;
    for (int e2 = 0; e2 < me * El + le; ++e2) off += tab[s * E + e2];
    home_off[t] = off;
  }
  if (t == 0) {
    int run = 0;
    for (int g = 0; g < nseg; ++g) {
      s_seg[g] = run;
      run += tab[(g / El) * E + me * El + (g % El)];
    }
    s_seg[nseg] = run;
  }
  __syncthreads();
  for (int g = t; g <= nseg; g += blockDim.x) seg_off[g] = s_seg[g];
  const int total = min(s_seg[nseg], cap);
  for (int i = t; i < cap; i += blockDim.x) {
    int v = -1;
    if (i < total) {
      int g = 0;
      while (g + 1 <= nseg && s_seg[g + 1] <= i) ++g;
      v = g % El;
    }
    idx_recv[i] = v;
  }
}

// one warp per compact source row r (canonical expert order): copy it into the owner's receive buffer
template <typename T>
__global__ void __launch_bounds__(256)
ep_dispatch_kernel(const T* __restrict__ src, const int* __restrict__ row_src_c, const int* __restrict__ cmp_off,
                   const int* __restrict__ send_off, T* const* __restrict__ peer_bufs, int K, int E, int El, int D,
                   int cap) {
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
    const int dest_row = send_off[e] + (r - cmp_off[e]);
    if (dest_row >= cap) continue;                          // beyond the receiver's capacity: dropped
    const int s = row_src_c != nullptr ? row_src_c[r] / K : r;
    const T* from = src + (long long)s * D;
    T* to = peer_bufs[dest_rank] + (long long)dest_row * D;   // peer-mapped address: the store crosses NVLink
    for (int v = lane; v < nv; v += 32)
      *reinterpret_cast<uint4*>(to + v * VT) = __ldg(reinterpret_cast<const uint4*>(from + v * VT));
  }
}

// one warp per received row i: send rows[row_map[i]] back to its home rank at its compact position
template <typename T>
__global__ void __launch_bounds__(256)
ep_return_kernel(const T* __restrict__ rows, const int* __restrict__ row_map, const int* __restrict__ seg_off,
                 const int* __restrict__ home_off, T* const* __restrict__ peer_rets, int W, int El, int D, int cap) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nseg = W * El;
  const int total = min(seg_off[nseg], cap);
  const int nv = D / VT;
  for (int i = warp; i < total; i += nwarps) {
    int g = 0;
    while (g + 1 < nseg && seg_off[g + 1] <= i) ++g;
    const int home_rank = g / El;
    const int home_row = home_off[g] + (i - seg_off[g]);
    const int s = row_map != nullptr ? row_map[i] : i;
    if (s < 0) continue;
    const T* from = rows + (long long)s * D;
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

int b200_ep_layout(const int32_t* tab, int me, int W, int E, int cap, int32_t* send_off, int32_t* seg_off,
                   int32_t* home_off, int32_t* idx_recv, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(W > 0 && E > 0 && E % W == 0 && E <= 1024 && W * (E / W) <= EP_MAX_SEG && cap > 0,
                 "ep_layout: bad arguments (W=%d E=%d cap=%d)", W, E, cap);
  launch_kernel(ep_layout_kernel, dim3(1), dim3(1024), 0, stream, tab, me, W, E, E / W, cap, send_off, seg_off, home_off,
                idx_recv);
  B200_LAUNCH_CHECK("ep_layout_kernel");
  count_launch();
  return 0;
}

int b200_ep_dispatch(const void* src, const int32_t* row_src_c, const int32_t* cmp_off, const int32_t* send_off,
                     void* const* peer_bufs, int K, int NK, int E, int El, int D, int cap, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(NK > 0 && E > 0 && El > 0 && D % (dtype == B200_BF16 ? 8 : 4) == 0, "ep_dispatch: bad arguments");
  if (dtype == B200_BF16)
    launch_kernel(ep_dispatch_kernel<bf16>, dim3(ep_grid(NK)), dim3(256), 0, stream, (const bf16*)src, row_src_c, cmp_off,
                  send_off, (bf16* const*)peer_bufs, K, E, El, D, cap);
  else
    launch_kernel(ep_dispatch_kernel<float>, dim3(ep_grid(NK)), dim3(256), 0, stream, (const float*)src, row_src_c,
                  cmp_off, send_off, (float* const*)peer_bufs, K, E, El, D, cap);
  B200_LAUNCH_CHECK("ep_dispatch_kernel");
  count_launch();
  return 0;
}

int b200_ep_return(const void* rows, const int32_t* row_map, const int32_t* seg_off, const int32_t* home_off,
                   void* const* peer_rets, int W, int El, int D, int cap, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(W > 0 && El > 0 && cap > 0 && D % (dtype == B200_BF16 ? 8 : 4) == 0, "ep_return: bad arguments");
  if (dtype == B200_BF16)
    launch_kernel(ep_return_kernel<bf16>, dim3(ep_grid(cap)), dim3(256), 0, stream, (const bf16*)rows, row_map, seg_off,
                  home_off, (bf16* const*)peer_rets, W, El, D, cap);
  else
    launch_kernel(ep_return_kernel<float>, dim3(ep_grid(cap)), dim3(256), 0, stream, (const float*)rows, row_map, seg_off,
                  home_off, (float* const*)peer_rets, W, El, D, cap);
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

}  // extern "C"
