// MOE router, common case (bf16 token rows, E <= 8 experts, no noise, D % 128 == 0) on the tensor cores.
//
// The router is a [N, D] x [D, E] product with E = 8: 6 144 MACs per 1.5 KB token row.  On the FMA pipe that is
// ~190 warp instructions per token — more issue time than the row takes to stream from HBM (ncu, profiles/r02e:
// 14 % of DRAM peak, FMA-bound).  Routing decisions must come from fp32 gate weights (bit-exact expert indices
// whatever the activation dtype), so the weights are split into THREE bf16 terms, W = hi + mid + lo exactly
// (3 x 8 mantissa bits), and the logits are three bf16 tensor-core products accumulated in fp32:
// x (bf16, exact) times each term is exact in fp32, only the summation order differs from an fp32 FMA chain.
//
// Warp-level mma.sync (m16n8k16 / m16n8k8) is the right tool here, not tcgen05: the N dimension is 8, the operand
// that streams (x) is read exactly once, and a 16-token tile needs 144 MMAs — there is nothing to stage through
// TMA/TMEM; the kernel is bound by the token stream.  A-fragments are loaded straight from global memory with
// 16-byte loads: the k index of an MMA is a free permutation of the columns as long as both operands use the same
// one, so thread (g, t) of a warp takes columns [32c + 8t, 32c + 8t + 8) of rows g and g + 8 — two full 32-byte
// sectors per row and warp load — and the weight fragments are pre-arranged in shared memory in the same order.
//
// Backward: dl [16 tokens x 8 experts] comes out of the softmax / top-k backward in exactly the A-fragment layout;
// dx = dl . W runs as m16n8k8 MMAs with dl and W each split into two bf16 terms (three products, error ~2^-17,
// below the bf16 rounding of dx), column-permuted so that every thread owns 8 consecutive output columns (16-byte
// stores).
#include <atomic>

#include "rowops.cuh"

namespace b200 {

namespace {

constexpr int RM_WARPS = 8;
constexpr int RM_MAX_E = 64;   // partial layout shared with router_finalize_kernel: [grid][3][64]
constexpr int RM_BATCH = 8;
// Tickets of the fused finalize: the block of a forward launch that finishes LAST folds the per-block statistics
// (counts, sum of probabilities, load-balance loss) -- what a second single-block kernel used to do.  Device-global ring
// (zero at load, reset by the folding block); launches in flight at the same time on different streams get different slots.
constexpr int RM_TICKETS = 256;
__device__ unsigned int g_rm_tickets[RM_TICKETS];    // 32-column chunks per load batch: 2 x 8 x 16 B per thread in flight, double-buffered

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}

// v = hi + mid + lo with every term a bf16 (round-to-nearest residuals: 24 mantissa bits covered)
__device__ __forceinline__ void split3(float v, bf16& hi, bf16& mid, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(hi);
  mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);
  lo = __float2bfloat16_rn(r2);
}
__device__ __forceinline__ uint32_t pack2(bf16 a, bf16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
// Epilogue of one token row held by the 4 lanes of a quad (lane t owns experts 2t, 2t+1): softmax, top-k, outputs.
__device__ __forceinline__ void route_row(float l0, float l1, int n, bool ok, int t, int E, int K,
                                          int* __restrict__ idx, float* __restrict__ w, float* __restrict__ topk_sum,
                                          float* __restrict__ probs, float (&cnt)[2], float (&ps)[2]) {
  const int e0 = 2 * t, e1 = 2 * t + 1;
  const bool v0 = e0 < E, v1 = e1 < E;
  float m = fmaxf(v0 ? l0 : -INFINITY, v1 ? l1 : -INFINITY);
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  const float x0 = v0 ? __expf(l0 - m) : 0.f, x1 = v1 ? __expf(l1 - m) : 0.f;
  float s = x0 + x1;
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  const float p0 = x0 / s, p1 = x1 / s;
  if (ok) {
    if (v1 && (E & 1) == 0) {
      *reinterpret_cast<float2*>(probs + (long long)n * E + e0) = make_float2(p0, p1);
    } else {
      if (v0) probs[(long long)n * E + e0] = p0;
      if (v1) probs[(long long)n * E + e1] = p1;
    }
    if (v0) ps[0] += p0;
    if (v1) ps[1] += p1;
  }
  float a0 = v0 ? p0 : -1.f, a1 = v1 ? p1 : -1.f;
  float sel_sum = 0.f, my_w[2] = {0.f, 0.f};
  int my_i[2] = {0, 0};
#pragma unroll
  for (int k = 0; k < 8; ++k) {      // descending; exact ties -> lowest expert index
    if (k < K) {
      float bv = a0;
      int bi = e0;
      if (a1 > bv) { bv = a1; bi = e1; }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      sel_sum += bv;
      if ((k & 3) == t) { my_w[k >> 2] = bv; my_i[k >> 2] = bi; }
      if (bi == e0) { a0 = -2.f; if (ok) cnt[0] += 1.f; }
      if (bi == e1) { a1 = -2.f; if (ok) cnt[1] += 1.f; }
    }
  }
  if (ok) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = t + 4 * j;
      if (k < K) {
        idx[(long long)n * K + k] = my_i[j];
        w[(long long)n * K + k] = my_w[j] / sel_sum;
      }
    }
    if (t == 0) topk_sum[n] = sel_sum;
  }
}

__global__ void __launch_bounds__(RM_WARPS * 32, 1)
router_fwd_mma_kernel(const bf16* __restrict__ x, const float* __restrict__ w_gate, int N, int D, int E, int K,
                      int* __restrict__ idx, float* __restrict__ w, float* __restrict__ topk_sum,
                      float* __restrict__ probs, float* __restrict__ part /* [grid][3][64] */, float lb_weight,
                      float* __restrict__ counts, float* __restrict__ psum, float* __restrict__ loss,
                      unsigned int* __restrict__ ticket) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* ws = reinterpret_cast<uint4*>(smem_raw);     // [3 splits][D/32 chunks][32 lanes]
  __shared__ float red[RM_WARPS][2][8];
  pdl_trigger();
  pdl_wait();
  const int chunks = D >> 5;
  // stage the split gate weights in fragment order: unit (c, lane): expert g = lane >> 2, columns 32c + 8t .. +7
  auto stage_weights = [&]() {
#pragma unroll 2
    for (int u = threadIdx.x; u < chunks * 32; u += blockDim.x) {
      const int c = u >> 5, ln = u & 31, eg = ln >> 2, et = ln & 3;
      float v[8];
      if (eg < E) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(w_gate + (long long)eg * D + c * 32 + et * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(w_gate + (long long)eg * D + c * 32 + et * 8 + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      bf16 h[8], m[8], l[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) split3(v[i], h[i], m[i], l[i]);
      ws[(0 * chunks + c) * 32 + ln] = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
      ws[(1 * chunks + c) * 32 + ln] = make_uint4(pack2(m[0], m[1]), pack2(m[2], m[3]), pack2(m[4], m[5]), pack2(m[6], m[7]));
      ws[(2 * chunks + c) * 32 + ln] = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
    }
    __syncthreads();
  };
  bool staged = false;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = (N + 15) >> 4;
  float cnt[2] = {0.f, 0.f}, ps[2] = {0.f, 0.f};
  // every warp of the block runs the same number of iterations (staging has a block-wide barrier in the first one)
  const int iters = (ntiles - blockIdx.x * RM_WARPS + gridDim.x * RM_WARPS - 1) / (gridDim.x * RM_WARPS);
  for (int it = 0; it < iters; ++it) {
    const int tile = blockIdx.x * RM_WARPS + warp + it * gridDim.x * RM_WARPS;
    const int n0 = tile * 16 + g, n1 = n0 + 8;
    const bool ok0 = n0 < N, ok1 = n1 < N;
    const uint4* r0 = reinterpret_cast<const uint4*>(x + (long long)(ok0 ? n0 : 0) * D) + t;
    const uint4* r1 = reinterpret_cast<const uint4*>(x + (long long)(ok1 ? n1 : 0) * D) + t;
    float acc[3][4];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[s][i] = 0.f;
    // batches of RM_BATCH chunks: the loads of batch b+1 are in flight while batch b multiplies
    uint4 a0[RM_BATCH], a1[RM_BATCH];
#pragma unroll
    for (int i = 0; i < RM_BATCH; ++i) {
      a0[i] = (ok0 && i < chunks) ? __ldg(r0 + 4 * i) : make_uint4(0, 0, 0, 0);
      a1[i] = (ok1 && i < chunks) ? __ldg(r1 + 4 * i) : make_uint4(0, 0, 0, 0);
    }
    if (!staged) {          // first tile: the weight staging below overlaps the first batch of row loads
      stage_weights();
      staged = true;
    }
    for (int cb = 0; cb < chunks; cb += RM_BATCH) {
      uint4 b0[RM_BATCH], b1[RM_BATCH];
#pragma unroll
      for (int i = 0; i < RM_BATCH; ++i) {
        const bool in = cb + RM_BATCH + i < chunks;
        b0[i] = (in && ok0) ? __ldg(r0 + 4 * (cb + RM_BATCH + i)) : make_uint4(0, 0, 0, 0);
        b1[i] = (in && ok1) ? __ldg(r1 + 4 * (cb + RM_BATCH + i)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < RM_BATCH; ++i) {
        if (cb + i < chunks) {
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const uint4 wv = ws[(s * chunks + cb + i) * 32 + lane];
            mma_16816(acc[s], a0[i].x, a1[i].x, a0[i].y, a1[i].y, wv.x, wv.y);
            mma_16816(acc[s], a0[i].z, a1[i].z, a0[i].w, a1[i].w, wv.z, wv.w);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < RM_BATCH; ++i) { a0[i] = b0[i]; a1[i] = b1[i]; }
    }
    float lg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) lg[i] = (acc[2][i] + acc[1][i]) + acc[0][i];
    route_row(lg[0], lg[1], n0, ok0, t, E, K, idx, w, topk_sum, probs, cnt, ps);
    route_row(lg[2], lg[3], n1, ok1, t, E, K, idx, w, topk_sum, probs, cnt, ps);
  }
  // statistics: fold the 8 row groups of the warp, then the warps of the block
#pragma unroll
  for (int o = 4; o <= 16; o <<= 1) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      cnt[j] += __shfl_xor_sync(0xffffffffu, cnt[j], o);
      ps[j] += __shfl_xor_sync(0xffffffffu, ps[j], o);
    }
  }
  if (lane < 4) {
    red[warp][0][2 * lane] = cnt[0]; red[warp][0][2 * lane + 1] = cnt[1];
    red[warp][1][2 * lane] = ps[0];  red[warp][1][2 * lane + 1] = ps[1];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * RM_MAX_E; c += blockDim.x) {
    const int which = c / RM_MAX_E, e = c % RM_MAX_E;
    float s = 0.f;
    if (which < 2 && e < 8) {
#pragma unroll
      for (int wp = 0; wp < RM_WARPS; ++wp) s += red[wp][which][e];
    }
    part[((long long)blockIdx.x * 3 + which) * RM_MAX_E + e] = s;
  }
  // ---- finalize in the last block to arrive: counts[e], psum[e], loss = lb_weight * E * sum_e (counts/N) (psum/N) ----
  if (ticket == nullptr) return;       // statistics are folded by router_finalize_kernel (B200VQA_ROUTER_FOLD=0)
  __shared__ bool s_last;
  __shared__ float s_tot[2][8];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int arrived = atomicAdd(ticket, 1u);
    s_last = (arrived == gridDim.x - 1);
    if (s_last) *ticket = 0u;          // ready for the next launch that is handed this slot
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    // 16 (statistic, expert) pairs x 16 threads each: fixed strided partial sums, then a shuffle tree (deterministic)
    const int pair = threadIdx.x >> 4, l16 = threadIdx.x & 15;
    const int which = pair >> 3, e = pair & 7;
    float s = 0.f;
    for (int b = l16; b < (int)gridDim.x; b += 16) s += __ldcg(part + ((long long)b * 3 + which) * RM_MAX_E + e);
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (l16 == 0) s_tot[which][e] = s;
  }
  __syncthreads();
  if (threadIdx.x < E) {
    counts[threadIdx.x] = s_tot[0][threadIdx.x];
    if (psum != nullptr) psum[threadIdx.x] = s_tot[1][threadIdx.x];
  }
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int i = 0; i < E; ++i) acc += (s_tot[0][i] / (float)N) * (s_tot[1][i] / (float)N);
    loss[0] = lb_weight * (float)E * acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward: dl (softmax / top-k / aux-loss backward) and dx = dl . W
// ---------------------------------------------------------------------------------------------------------------
// d logits of one token row for the quad lane t (experts 2t, 2t+1)
__device__ __forceinline__ void dl_row(int n, bool ok, int t, int E, int K, float lb_scale,
                                       const int* __restrict__ idx, const float* __restrict__ w,
                                       const float* __restrict__ topk_sum, const float* __restrict__ probs,
                                       const float* __restrict__ counts, const float* __restrict__ d_w,
                                       const float* __restrict__ d_probs, bool aux, float* __restrict__ dl_out,
                                       float& o0, float& o1) {
  const int e0 = 2 * t, e1 = 2 * t + 1;
  const bool v0 = ok && e0 < E, v1 = ok && e1 < E;
  const float p0 = v0 ? probs[(long long)n * E + e0] : 0.f, p1 = v1 ? probs[(long long)n * E + e1] : 0.f;
  float dq0 = 0.f, dq1 = 0.f;
  if (d_w != nullptr && ok) {
    const float ssum = topk_sum[n];
    float dot = 0.f;
    for (int k = 0; k < K; ++k) dot = fmaf(d_w[(long long)n * K + k], w[(long long)n * K + k], dot);
    for (int k = 0; k < K; ++k) {
      const int ik = idx[(long long)n * K + k];
      const float gk = (d_w[(long long)n * K + k] - dot) / ssum;
      if (ik == e0) dq0 += gk;
      if (ik == e1) dq1 += gk;
    }
  }
  float qdq = p0 * dq0 + p1 * dq1;
  qdq += __shfl_xor_sync(0xffffffffu, qdq, 1);
  qdq += __shfl_xor_sync(0xffffffffu, qdq, 2);
  float d0 = p0 * (dq0 - qdq), d1 = p1 * (dq1 - qdq);
  if (aux) {     // uniform over the warp
    float dp0 = v0 ? lb_scale * counts[e0] : 0.f, dp1 = v1 ? lb_scale * counts[e1] : 0.f;
    if (d_probs != nullptr) {
      if (v0) dp0 += d_probs[(long long)n * E + e0];
      if (v1) dp1 += d_probs[(long long)n * E + e1];
    }
    float pdp = p0 * dp0 + p1 * dp1;
    pdp += __shfl_xor_sync(0xffffffffu, pdp, 1);
    pdp += __shfl_xor_sync(0xffffffffu, pdp, 2);
    d0 += p0 * (dp0 - pdp);
    d1 += p1 * (dp1 - pdp);
  }
  if (v0) dl_out[(long long)n * E + e0] = d0;
  if (v1) dl_out[(long long)n * E + e1] = d1;
  o0 = v0 ? d0 : 0.f;
  o1 = v1 ? d1 : 0.f;
}

__global__ void __launch_bounds__(RM_WARPS * 32, 2)
router_bwd_mma_kernel(const float* __restrict__ w_gate, float lb_weight, int N, int D, int E, int K,
                      const int* __restrict__ idx, const float* __restrict__ w, const float* __restrict__ topk_sum,
                      const float* __restrict__ probs, const float* __restrict__ counts,
                      const float* __restrict__ d_w, const float* __restrict__ d_loss,
                      const float* __restrict__ d_probs, bf16* __restrict__ dx, float* __restrict__ dl_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* wf = reinterpret_cast<uint32_t*>(smem_raw);   // [2 splits][D/32 groups][4 tiles][32 lanes]
  pdl_trigger();
  pdl_wait();
  const int groups = D >> 5;
  // B fragment of n-tile i of column group b for lane (g, t): experts (2t, 2t+1) at physical column
  // 32b + 8(g >> 1) + 2i + (g & 1), so that the C fragments of the 4 tiles give every thread 8 consecutive columns
  for (int u = threadIdx.x; u < groups * 128; u += blockDim.x) {
    const int ln = u & 31, i = (u >> 5) & 3, b = u >> 7, g = ln >> 2, t = ln & 3;
    const int col = 32 * b + 8 * (g >> 1) + 2 * i + (g & 1);
    const float v0 = 2 * t < E ? __ldg(w_gate + (long long)(2 * t) * D + col) : 0.f;
    const float v1 = 2 * t + 1 < E ? __ldg(w_gate + (long long)(2 * t + 1) * D + col) : 0.f;
    const bf16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
    const bf16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
    wf[u] = pack2(h0, h1);
    wf[groups * 128 + u] = pack2(l0, l1);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const float gl = d_loss != nullptr ? d_loss[0] : 0.f;
  const bool aux = gl != 0.f || d_probs != nullptr;
  const float lb_scale = gl * lb_weight * (float)E / ((float)N * (float)N);
  const int ntiles = (N + 15) >> 4;
  for (int tile = blockIdx.x * RM_WARPS + warp; tile < ntiles; tile += gridDim.x * RM_WARPS) {
    const int n0 = tile * 16 + g, n1 = n0 + 8;
    const bool ok0 = n0 < N, ok1 = n1 < N;
    float d00, d01, d10, d11;
    dl_row(n0, ok0, t, E, K, lb_scale, idx, w, topk_sum, probs, counts, d_w, d_probs, aux, dl_out, d00, d01);
    dl_row(n1, ok1, t, E, K, lb_scale, idx, w, topk_sum, probs, counts, d_w, d_probs, aux, dl_out, d10, d11);
    const bf16 h00 = __float2bfloat16_rn(d00), h01 = __float2bfloat16_rn(d01);
    const bf16 h10 = __float2bfloat16_rn(d10), h11 = __float2bfloat16_rn(d11);
    const uint32_t ah0 = pack2(h00, h01), ah1 = pack2(h10, h11);
    const uint32_t al0 = pack2(__float2bfloat16_rn(d00 - __bfloat162float(h00)), __float2bfloat16_rn(d01 - __bfloat162float(h01)));
    const uint32_t al1 = pack2(__float2bfloat16_rn(d10 - __bfloat162float(h10)), __float2bfloat16_rn(d11 - __bfloat162float(h11)));
    bf16* o0 = dx + (long long)n0 * D + 8 * t;
    bf16* o1 = dx + (long long)n1 * D + 8 * t;
#pragma unroll 2
    for (int b = 0; b < groups; ++b) {
      float c[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
        const uint32_t bh = wf[(b * 4 + i) * 32 + lane], bl = wf[groups * 128 + (b * 4 + i) * 32 + lane];
        mma_1688(c[i], al0, al1, bh);
        mma_1688(c[i], ah0, ah1, bl);
        mma_1688(c[i], ah0, ah1, bh);
      }
      if (ok0)
        *reinterpret_cast<uint4*>(o0 + 32 * b) = make_uint4(pack_bf16x2(c[0][0], c[0][1]), pack_bf16x2(c[1][0], c[1][1]),
                                                            pack_bf16x2(c[2][0], c[2][1]), pack_bf16x2(c[3][0], c[3][1]));
      if (ok1)
        *reinterpret_cast<uint4*>(o1 + 32 * b) = make_uint4(pack_bf16x2(c[0][2], c[0][3]), pack_bf16x2(c[1][2], c[1][3]),
                                                            pack_bf16x2(c[2][2], c[2][3]), pack_bf16x2(c[3][2], c[3][3]));
    }
  }
}

inline int mma_grid(int N, int per_sm) {
  const int need = (((N + 15) >> 4) + RM_WARPS - 1) / RM_WARPS;
  const int cap = num_sms() * per_sm;
  return need < cap ? (need < 1 ? 1 : need) : cap;
}

}  // namespace

bool router_mma_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VQA_ROUTER_MMA");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// the last block of the forward kernel folds the statistics (default) / a separate finalize launch does (A/B testing)
bool router_fold_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VQA_ROUTER_FOLD");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static bool router_mma_covers(int D, int E, int K) { return router_mma_enabled() && D % 128 == 0 && D <= 2048 && E <= 8 && K <= 8; }

// returns the number of blocks launched (> 0), or -1 when the shape is not covered, or -2 on a launch error
int launch_router_fwd_mma(const bf16* x, const float* w_gate, int N, int D, int E, int K, int* idx, float* w,
                          float* topk_sum, float* probs, float* part, float lb_weight, float* counts, float* psum,
                          float* loss, cudaStream_t stream) {
  if (!router_mma_covers(D, E, K)) return -1;
  const size_t smem = (size_t)3 * D * 16;       // 3 splits x D/32 chunks x 32 lanes x 16 B
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(router_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
      return -2;
    attr_set = true;
  }
  const int grid = mma_grid(N, 1);
  static std::atomic<unsigned int> next_ticket{0};
  unsigned int* tk = nullptr;
  if (router_fold_enabled()) {
    if (cudaGetSymbolAddress((void**)&tk, g_rm_tickets) != cudaSuccess) return -2;
    tk += next_ticket.fetch_add(1u) % RM_TICKETS;
  }
  launch_kernel(router_fwd_mma_kernel, dim3(grid), dim3(RM_WARPS * 32), smem, stream, x, w_gate, N, D, E, K, idx, w,
                topk_sum, probs, part, lb_weight, counts, psum, loss, tk);
  if (cudaGetLastError() != cudaSuccess) return -2;
  return grid;
}

int launch_router_bwd_mma(const float* w_gate, float lb_weight, int N, int D, int E, int K, const int* idx,
                          const float* w, const float* topk_sum, const float* probs, const float* counts,
                          const float* d_w, const float* d_loss, const float* d_probs, bf16* dx, float* dl_out,
                          cudaStream_t stream) {
  if (!router_mma_covers(D, E, K)) return -1;
  const size_t smem = (size_t)2 * D * 16;       // 2 splits x D/32 groups x 4 tiles x 32 lanes x 4 B
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(router_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
      return -2;
    attr_set = true;
  }
  const int grid = mma_grid(N, 2);
  launch_kernel(router_bwd_mma_kernel, dim3(grid), dim3(RM_WARPS * 32), smem, stream, w_gate, lb_weight, N, D, E, K, idx,
                w, topk_sum, probs, counts, d_w, d_loss, d_probs, dx, dl_out);
  if (cudaGetLastError() != cudaSuccess) return -2;
  return 0;
}

}  // namespace b200
