// MOE router: gate dot-products, (noisy) softmax, top-k, renormalisation and the Switch load-balance
// statistics in one pass over x; backward produces dx and the gate-weight gradients.
// One warp per token, fp32 throughout (routing decisions must not depend on the activation dtype),
// gate weights staged in shared memory, lane e owns expert e (and e+32).
#include "rowops.cuh"

namespace b200 {

int launch_partial_reduce(const float* part, int blocks, int rows_per_block, int width, const int* tile_group, int G,
                          float* out, cudaStream_t stream);
// router_mma.cu: tensor-core kernels for the common case (bf16 rows, E <= 8, no noise, D % 128 == 0)
int launch_router_fwd_mma(const bf16* x, const float* w_gate, int N, int D, int E, int K, int* idx, float* w,
                          float* topk_sum, float* probs, float* part, float lb_weight, float* counts, float* psum,
                          float* loss, cudaStream_t stream);
bool router_fold_enabled();
int launch_router_bwd_mma(const float* w_gate, float lb_weight, int N, int D, int E, int K, const int* idx,
                          const float* w, const float* topk_sum, const float* probs, const float* counts,
                          const float* d_w, const float* d_loss, const float* d_probs, bf16* dx, float* dl_out,
                          cudaStream_t stream);

namespace {

constexpr int RT_MAX_E = 64;
constexpr int RT_WARPS = 8;

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(__expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

// Gate weights live in shared memory in a lane-interleaved float4 layout so that a warp's reads are conflict-free:
//   unit(e, j, q, lane) = ((e*NV + j)*(VT/4) + q)*32 + lane   holds elements d = (lane + 32j)*VT + 4q .. +3 of row e
// (the row-major staging used before produced 8-way bank conflicts: 19.6 M conflicts per call in ncu).
template <int VT, int NV>
__device__ __forceinline__ void stage_gate_weights(const float* __restrict__ w, float* __restrict__ ws, int E, int D) {
  // one 16-byte unit per iteration (4 consecutive weights of a row), several independent loads in flight: the
  // scalar version was a chain of dependent-latency iterations (~15 us per block for E = 8, D = 768)
  constexpr int QV = VT / 4;
  float4* ws4 = reinterpret_cast<float4*>(ws);
  const int units = E * NV * QV * 32;
#pragma unroll 4
  for (int u = threadIdx.x; u < units; u += blockDim.x) {
    const int lane = u & 31, t = u >> 5;
    const int q = t % QV, j = (t / QV) % NV, e = t / (QV * NV);
    const int d = (lane + 32 * j) * VT + 4 * q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d < D) v = __ldg(reinterpret_cast<const float4*>(w + (long long)e * D + d));   // D % 4 == 0: whole unit valid
    ws4[u] = v;
  }
}

// dot products of one token row (held in registers, loaded once) with the E staged weight rows.
// On return lane e (and slot 1: e+32) holds logit e.  EB > 0: E <= EB is a compile-time bound, the expert loop is
// unrolled and the EB warp reductions run interleaved (5 dependent shuffle rounds instead of 5 * E).
template <typename T, int NV, int EB>
__device__ __forceinline__ void gate_dots(const RowRegs<T, NV>& x, const float* __restrict__ ws, int E, int lane,
                                          float& l0, float& l1) {
  constexpr int VT = Vec16<T>::N;
  const float4* ws4 = reinterpret_cast<const float4*>(ws);
  l0 = 0.f;
  l1 = 0.f;
  if (EB > 0) {
    float s[EB > 0 ? EB : 1];
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      s[e] = 0.f;
      if (e < E) {
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
          for (int q = 0; q < VT / 4; ++q) {
            const float4 wv = ws4[((e * NV + j) * (VT / 4) + q) * 32 + lane];
            s[e] = fmaf(x.v[j][4 * q + 0], wv.x, s[e]);
            s[e] = fmaf(x.v[j][4 * q + 1], wv.y, s[e]);
            s[e] = fmaf(x.v[j][4 * q + 2], wv.z, s[e]);
            s[e] = fmaf(x.v[j][4 * q + 3], wv.w, s[e]);
          }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int e = 0; e < EB; ++e) s[e] += __shfl_xor_sync(0xffffffffu, s[e], o);
#pragma unroll
    for (int e = 0; e < EB; ++e)
      if (e == lane) l0 = s[e];
    return;
  }
  for (int e = 0; e < E; ++e) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int q = 0; q < VT / 4; ++q) {
        const float4 wv = ws4[((e * NV + j) * (VT / 4) + q) * 32 + lane];
        s = fmaf(x.v[j][4 * q + 0], wv.x, s);
        s = fmaf(x.v[j][4 * q + 1], wv.y, s);
        s = fmaf(x.v[j][4 * q + 2], wv.z, s);
        s = fmaf(x.v[j][4 * q + 3], wv.w, s);
      }
    s = warp_sum(s);
    if ((e & 31) == lane) {
      if (e < 32) l0 = s; else l1 = s;
    }
  }
}

// Token-blocked variants for the common case (bf16 rows, E <= 8, no noise): a warp handles TB consecutive tokens at
// once so every staged weight vector read from shared memory feeds TB rows.  With one token per warp the kernel was
// bound by shared-memory bandwidth (E * D * 4 bytes of weight reads per token: 35 % of HBM peak at best).  The
// accumulation order per lane is the same as in gate_dots, so both paths give bit-identical logits.
constexpr int RT_TB = 4;
template <int NV, int EB, int TB>
__device__ __forceinline__ void gate_dots_block(const bf16* __restrict__ x, const float* __restrict__ ws, int base,
                                                int N, int D, int E, int lane, float& out) {
  const float4* ws4 = reinterpret_cast<const float4*>(ws);
  const int nv = D / 8;
  uint4 raw[TB][NV];
#pragma unroll
  for (int t = 0; t < TB; ++t)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int vi = lane + 32 * j;
      raw[t][j] = (base + t < N && vi < nv)
                      ? __ldg(reinterpret_cast<const uint4*>(x + (long long)(base + t) * D + vi * 8))
                      : make_uint4(0, 0, 0, 0);
    }
  float s[TB][EB];
#pragma unroll
  for (int t = 0; t < TB; ++t)
#pragma unroll
    for (int e = 0; e < EB; ++e) s[t][e] = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float xv[TB][4];
#pragma unroll
      for (int t = 0; t < TB; ++t) {
        const float2 a = unpack_bf16x2(q == 0 ? raw[t][j].x : raw[t][j].z);
        const float2 b = unpack_bf16x2(q == 0 ? raw[t][j].y : raw[t][j].w);
        xv[t][0] = a.x; xv[t][1] = a.y; xv[t][2] = b.x; xv[t][3] = b.y;
      }
#pragma unroll
      for (int e = 0; e < EB; ++e) {
        if (e < E) {
          const float4 wv = ws4[((e * NV + j) * 2 + q) * 32 + lane];
#pragma unroll
          for (int t = 0; t < TB; ++t) {
            s[t][e] = fmaf(xv[t][0], wv.x, s[t][e]);
            s[t][e] = fmaf(xv[t][1], wv.y, s[t][e]);
            s[t][e] = fmaf(xv[t][2], wv.z, s[t][e]);
            s[t][e] = fmaf(xv[t][3], wv.w, s[t][e]);
          }
        }
      }
    }
  // Transpose-reduce: 32 partial sums per lane (index i = t * EB + e) -> lane l ends up with the FULL sum of index l.
  // Round with offset o: the lane keeps the half of its values whose index has bit o equal to its own lane bit and
  // receives the partner's partial sums of that half: 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 5 * 32.
  static_assert(TB * EB == 32, "one value per lane");
  float v[32];
#pragma unroll
  for (int t = 0; t < TB; ++t)
#pragma unroll
    for (int e = 0; e < EB; ++e) v[t * EB + e] = s[t][e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float keep = up ? v[i + o] : v[i];
      const float send = up ? v[i] : v[i + o];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  out = v[0];
}

// dx rows of TB consecutive tokens: dx[t] = sum_e dl[t][e] * Wg[e,:]  (dl from shared memory, same layout)
template <int NV, int EB, int TB>
__device__ __forceinline__ void gate_dx_block(const float* __restrict__ ws, const float (*dl_s)[RT_MAX_E], int base,
                                              int N, int D, int E, int lane, bf16* __restrict__ dx) {
  const float4* ws4 = reinterpret_cast<const float4*>(ws);
  const int nv = D / 8;
  float dl[TB][EB];
#pragma unroll
  for (int t = 0; t < TB; ++t)
#pragma unroll
    for (int e = 0; e < EB; ++e) dl[t][e] = (e < E) ? dl_s[t][e] : 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float o[TB][8];
#pragma unroll
    for (int t = 0; t < TB; ++t)
#pragma unroll
      for (int u = 0; u < 8; ++u) o[t][u] = 0.f;
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      if (e < E) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 wv = ws4[((e * NV + j) * 2 + q) * 32 + lane];
#pragma unroll
          for (int t = 0; t < TB; ++t) {
            o[t][4 * q + 0] = fmaf(dl[t][e], wv.x, o[t][4 * q + 0]);
            o[t][4 * q + 1] = fmaf(dl[t][e], wv.y, o[t][4 * q + 1]);
            o[t][4 * q + 2] = fmaf(dl[t][e], wv.z, o[t][4 * q + 2]);
            o[t][4 * q + 3] = fmaf(dl[t][e], wv.w, o[t][4 * q + 3]);
          }
        }
      }
    }
    const int vi = lane + 32 * j;
    if (vi < nv) {
#pragma unroll
      for (int t = 0; t < TB; ++t) {
        if (base + t < N) {
          uint4 pk;
          pk.x = pack_bf16x2(o[t][0], o[t][1]); pk.y = pack_bf16x2(o[t][2], o[t][3]);
          pk.z = pack_bf16x2(o[t][4], o[t][5]); pk.w = pack_bf16x2(o[t][6], o[t][7]);
          *reinterpret_cast<uint4*>(dx + (long long)(base + t) * D + vi * 8) = pk;
        }
      }
    }
  }
}

// softmax over the E logits spread as (lane, slot); invalid slots must hold -inf
__device__ __forceinline__ void warp_softmax(float& a, float& b) {
  const float m = warp_max(fmaxf(a, b));
  const float ea = __expf(a - m), eb = __expf(b - m);
  const float s = warp_sum(ea + eb);
  a = ea / s;
  b = eb / s;
}

template <typename T, int NV, int EB>
__global__ void __launch_bounds__(RT_WARPS * 32)
router_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w_gate, const float* __restrict__ w_noise,
                  const float* __restrict__ eps, float noise_std, int N, int D, int E, int K, int* __restrict__ idx,
                  float* __restrict__ w, float* __restrict__ topk_sum, float* __restrict__ probs,
                  float* __restrict__ probs_noisy, float* __restrict__ part /* [grid][3][RT_MAX_E] */) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  constexpr int DP = NV * 32 * VT;
  extern __shared__ float smem[];
  float* wg = smem;                                  // [E][DP] lane-interleaved
  float* wn = smem + (size_t)E * DP;                 // noisy only
  __shared__ float red[RT_WARPS][3][RT_MAX_E];
  const bool noisy = (eps != nullptr);
  stage_gate_weights<VT, NV>(w_gate, wg, E, D);
  if (noisy) stage_gate_weights<VT, NV>(w_noise, wn, E, D);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool v0 = lane < E, v1 = lane + 32 < E;
  float cnt0 = 0.f, cnt1 = 0.f, ps0 = 0.f, ps1 = 0.f, ns0 = 0.f, ns1 = 0.f;

  constexpr bool FAST = (EB == 8) && (sizeof(T) == 2) && (NV <= 4);   // D <= 1024: the fully unrolled blocked path
  constexpr int TBC = FAST ? RT_TB : 1;
  const bool fast = FAST && !noisy;
  const int tb = fast ? TBC : 1;
  for (int base = (blockIdx.x * RT_WARPS + warp) * tb; base < N; base += gridDim.x * RT_WARPS * tb) {
   if constexpr (FAST) if (fast) {
     // Token-blocked path: after the transpose-reduce lane l holds the logit of (token base + l / 8, expert l % 8),
     // so softmax, top-k and the statistics of the 4 tokens run at once in 8-lane groups (3 shuffle rounds each).
     float logit;
     gate_dots_block<NV, EB, TBC>(reinterpret_cast<const bf16*>(x), wg, base, N, D, E, lane, logit);
     const int e = lane & 7, n = base + (lane >> 3);
     const bool tok_ok = n < N, ve = e < E;
     float m = ve ? logit : -INFINITY;
#pragma unroll
     for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
     const float ex = ve ? __expf(logit - m) : 0.f;
     float sum = ex;
#pragma unroll
     for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
     const float p = ex / sum;
     if (tok_ok && ve) { probs[(long long)n * E + e] = p; ps0 += p; }
     float a = ve ? p : -1.f, sel_sum = 0.f, my_w = 0.f;
     int my_i = 0;
     for (int k = 0; k < K; ++k) {       // descending; exact ties -> lowest expert index
       float bv = a;
       int bi = e;
#pragma unroll
       for (int o = 4; o > 0; o >>= 1) {
         const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
         const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
         if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
       }
       sel_sum += bv;
       if (e == k) { my_w = bv; my_i = bi; }
       if (bi == e) { a = -2.f; if (tok_ok) cnt0 += 1.f; }
     }
     if (tok_ok && e < K) {
       idx[(long long)n * K + e] = my_i;
       w[(long long)n * K + e] = my_w / sel_sum;
     }
     if (tok_ok && e == 0) topk_sum[n] = sel_sum;
     continue;
   }
   {
    const int n = base;
    RowRegs<T, NV> xr;
    float c0, c1;  // clean logits -> clean probs
    xr.load(x + (long long)n * D, D, lane);
    gate_dots<T, NV, EB>(xr, wg, E, lane, c0, c1);
    float q0 = c0, q1 = c1;  // logits used for selection
    if (noisy) {
      float u0, u1;
      gate_dots<T, NV, EB>(xr, wn, E, lane, u0, u1);
      const float sp0 = softplus_f(u0), sp1 = softplus_f(u1);
      if (v0) { q0 = c0 + eps[(long long)n * E + lane] * sp0 * noise_std; ns0 += sp0; }
      if (v1) { q1 = c1 + eps[(long long)n * E + lane + 32] * sp1 * noise_std; ns1 += sp1; }
    }
    if (!v0) { c0 = -INFINITY; q0 = -INFINITY; }
    if (!v1) { c1 = -INFINITY; q1 = -INFINITY; }
    warp_softmax(c0, c1);
    if (noisy) warp_softmax(q0, q1);
    else { q0 = c0; q1 = c1; }
    if (v0) { probs[(long long)n * E + lane] = c0; ps0 += c0; }
    if (v1) { probs[(long long)n * E + lane + 32] = c1; ps1 += c1; }
    if (noisy) {
      if (v0) probs_noisy[(long long)n * E + lane] = q0;
      if (v1) probs_noisy[(long long)n * E + lane + 32] = q1;
    }
    // top-k by repeated warp arg-max (descending; exact ties -> lowest expert index)
    float a0 = v0 ? q0 : -1.f, a1 = v1 ? q1 : -1.f;
    float sel_sum = 0.f;
    float my_w = 0.f;
    int my_i = 0;
    for (int k = 0; k < K; ++k) {
      float bv = a0;
      int bi = lane;
      if (a1 > bv) { bv = a1; bi = lane + 32; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      sel_sum += bv;
      if (lane == k) { my_w = bv; my_i = bi; }
      if (bi == lane) { a0 = -2.f; cnt0 += 1.f; }
      if (bi == lane + 32) { a1 = -2.f; cnt1 += 1.f; }
    }
    if (lane < K) {
      idx[(long long)n * K + lane] = my_i;
      w[(long long)n * K + lane] = my_w / sel_sum;
    }
    if (lane == 0) topk_sum[n] = sel_sum;
   }
  }
  if (FAST && fast) {   // the 4 token groups of a warp each counted expert (lane & 7): fold them onto lanes 0..7
    cnt0 += __shfl_xor_sync(0xffffffffu, cnt0, 8);
    cnt0 += __shfl_xor_sync(0xffffffffu, cnt0, 16);
    ps0 += __shfl_xor_sync(0xffffffffu, ps0, 8);
    ps0 += __shfl_xor_sync(0xffffffffu, ps0, 16);
    if (lane >= 8) { cnt0 = 0.f; ps0 = 0.f; }
  }

  // block partials: [0] counts, [1] sum of clean probs, [2] sum of softplus noise scales
  red[warp][0][lane] = cnt0; red[warp][0][lane + 32] = cnt1;
  red[warp][1][lane] = ps0;  red[warp][1][lane + 32] = ps1;
  red[warp][2][lane] = ns0;  red[warp][2][lane + 32] = ns1;
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * RT_MAX_E; c += blockDim.x) {
    const int which = c / RT_MAX_E, e = c % RT_MAX_E;
    float s = 0.f;
#pragma unroll
    for (int wp = 0; wp < RT_WARPS; ++wp) s += red[wp][which][e];
    part[((long long)blockIdx.x * 3 + which) * RT_MAX_E + e] = s;
  }
}

// counts[e], loss = lb_weight * E * sum_e (counts[e]/N) * (psum[e]/N), mean softplus noise scale
__global__ void __launch_bounds__(1024)
router_finalize_kernel(const float* __restrict__ part, int blocks, int N, int E, float lb_weight,
                       float* __restrict__ counts, float* __restrict__ psum, float* __restrict__ loss,
                       float* __restrict__ noise_scale_mean) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_red[16][3][RT_MAX_E];
  __shared__ float s_cnt[RT_MAX_E], s_ps[RT_MAX_E], s_ns[RT_MAX_E];
  const int e = threadIdx.x & (RT_MAX_E - 1), y = threadIdx.x / RT_MAX_E;   // 64 experts x 16 partial lanes
  float c = 0.f, p = 0.f, q = 0.f;
  for (int b = y; b < blocks; b += 16) {
    c += part[((long long)b * 3 + 0) * RT_MAX_E + e];
    p += part[((long long)b * 3 + 1) * RT_MAX_E + e];
    q += part[((long long)b * 3 + 2) * RT_MAX_E + e];
  }
  s_red[y][0][e] = c; s_red[y][1][e] = p; s_red[y][2][e] = q;
  __syncthreads();
  if (y == 0) {
    c = p = q = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { c += s_red[j][0][e]; p += s_red[j][1][e]; q += s_red[j][2][e]; }
    s_cnt[e] = c; s_ps[e] = p; s_ns[e] = q;
    if (e < E) {
      counts[e] = c;
      if (psum != nullptr) psum[e] = p;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f, ns = 0.f;
    for (int i = 0; i < E; ++i) {
      acc += (s_cnt[i] / (float)N) * (s_ps[i] / (float)N);
      ns += s_ns[i];
    }
    loss[0] = lb_weight * (float)E * acc;
    if (noise_scale_mean != nullptr) noise_scale_mean[0] = ns / ((float)N * (float)E);
  }
}

// ---- backward --------------------------------------------------------------------------------------------
// per token: d logits (clean) and d noise-logits, then dx = dl Wg + du Wn.  dl/du are also written to the
// workspace for the weight-gradient reduction.
template <typename T, int NV, int EB>
__global__ void __launch_bounds__(RT_WARPS * 32)
router_bwd_kernel(const T* __restrict__ x, const float* __restrict__ w_gate, const float* __restrict__ w_noise,
                  const float* __restrict__ eps, float noise_std, float lb_weight, int N, int D, int E, int K,
                  const int* __restrict__ idx, const float* __restrict__ w, const float* __restrict__ topk_sum,
                  const float* __restrict__ probs, const float* __restrict__ probs_noisy,
                  const float* __restrict__ counts, const float* __restrict__ d_w, const float* __restrict__ d_loss,
                  const float* __restrict__ d_probs, T* __restrict__ dx, float* __restrict__ dl_out,
                  float* __restrict__ du_out) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  constexpr int DP = NV * 32 * VT;
  extern __shared__ float smem[];
  float* wg = smem;
  float* wn = smem + (size_t)E * DP;
  constexpr bool FAST = (EB == 8) && (sizeof(T) == 2) && (NV <= 4);   // D <= 1024: the fully unrolled blocked path
  constexpr int TBC = FAST ? RT_TB : 1;
  __shared__ float s_dl[RT_WARPS][TBC][RT_MAX_E], s_du[RT_WARPS][RT_MAX_E];
  const bool noisy = (eps != nullptr);
  stage_gate_weights<VT, NV>(w_gate, wg, E, D);
  if (noisy) stage_gate_weights<VT, NV>(w_noise, wn, E, D);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool v0 = lane < E, v1 = lane + 32 < E;
  const float gl = (d_loss != nullptr) ? d_loss[0] : 0.f;

  const bool fast = FAST && !noisy;
  const int tb = fast ? TBC : 1;
  for (int base = (blockIdx.x * RT_WARPS + warp) * tb; base < N; base += gridDim.x * RT_WARPS * tb) {
   if (FAST && fast) {      // rows beyond N contribute nothing
#pragma unroll
     for (int t = 0; t < TBC; ++t) { s_dl[warp][t][lane] = 0.f; s_dl[warp][t][lane + 32] = 0.f; }
   }
#pragma unroll
   for (int t = 0; t < TBC; ++t) {
    const int n = base + t;
    if (t >= tb || n >= N) break;
    const float* qrow = (noisy ? probs_noisy : probs) + (long long)n * E;
    const float q0 = v0 ? qrow[lane] : 0.f, q1 = v1 ? qrow[lane + 32] : 0.f;
    // gradient wrt the selection probabilities q through w_k = q_{i_k} / sum_j q_{i_j}
    float dq0 = 0.f, dq1 = 0.f;
    if (d_w != nullptr) {
      const float ssum = topk_sum[n];
      float dot = 0.f;
      for (int k = 0; k < K; ++k) dot = fmaf(d_w[(long long)n * K + k], w[(long long)n * K + k], dot);
      for (int k = 0; k < K; ++k) {
        const int ik = idx[(long long)n * K + k];
        const float g = (d_w[(long long)n * K + k] - dot) / ssum;
        if (ik == lane) dq0 += g;
        if (ik == lane + 32) dq1 += g;
      }
    }
    float qdq = warp_sum(q0 * dq0 + q1 * dq1);
    float dsel0 = q0 * (dq0 - qdq), dsel1 = q1 * (dq1 - qdq);  // d (selection logits)
    // aux loss: d loss / d p_clean[n,e] = lb_weight * E * (counts[e]/N) / N; plus any gradient that reached
    // router_probs directly (aux_outputs['router_probs'] is differentiable in the reference, router.py:140)
    float dc0 = 0.f, dc1 = 0.f;
    if (gl != 0.f || d_probs != nullptr) {
      const float* prow = probs + (long long)n * E;
      const float p0 = v0 ? prow[lane] : 0.f, p1 = v1 ? prow[lane + 32] : 0.f;
      const float sc = gl * lb_weight * (float)E / ((float)N * (float)N);
      float dp0 = v0 ? sc * counts[lane] : 0.f, dp1 = v1 ? sc * counts[lane + 32] : 0.f;
      if (d_probs != nullptr) {
        if (v0) dp0 += d_probs[(long long)n * E + lane];
        if (v1) dp1 += d_probs[(long long)n * E + lane + 32];
      }
      const float pdp = warp_sum(p0 * dp0 + p1 * dp1);
      dc0 = p0 * (dp0 - pdp);
      dc1 = p1 * (dp1 - pdp);
    }
    float du0 = 0.f, du1 = 0.f;
    if (noisy) {
      RowRegs<T, NV> xr;
      xr.load(x + (long long)n * D, D, lane);
      float u0, u1;
      gate_dots<T, NV, EB>(xr, wn, E, lane, u0, u1);
      if (v0) du0 = dsel0 * eps[(long long)n * E + lane] * noise_std * sigmoid_f(u0);
      if (v1) du1 = dsel1 * eps[(long long)n * E + lane + 32] * noise_std * sigmoid_f(u1);
    }
    const float dl0 = dsel0 + dc0, dl1 = dsel1 + dc1;
    s_dl[warp][t][lane] = dl0; s_dl[warp][t][lane + 32] = dl1;
    s_du[warp][lane] = du0; s_du[warp][lane + 32] = du1;
    if (v0) dl_out[(long long)n * E + lane] = dl0;
    if (v1) dl_out[(long long)n * E + lane + 32] = dl1;
    if (noisy) {
      if (v0) du_out[(long long)n * E + lane] = du0;
      if (v1) du_out[(long long)n * E + lane + 32] = du1;
    }
    __syncwarp();
    if (FAST && fast) continue;   // dx of the whole token block is produced below
    // dx = sum_e dl[e] Wg[e,:] + du[e] Wn[e,:]   (same conflict-free weight layout)
    RowRegs<T, NV> o;
    o.zero();
    const float4* wg4 = reinterpret_cast<const float4*>(wg);
    const float4* wn4 = reinterpret_cast<const float4*>(wn);
#pragma unroll
    for (int e = 0; e < (EB > 0 ? EB : RT_MAX_E); ++e) {
      if (e >= E) break;
      const float a = s_dl[warp][t][e];
      const float bb = noisy ? s_du[warp][e] : 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int q = 0; q < VT / 4; ++q) {
          const int unit = ((e * NV + j) * (VT / 4) + q) * 32 + lane;
          const float4 wv = wg4[unit];
          o.v[j][4 * q + 0] = fmaf(a, wv.x, o.v[j][4 * q + 0]);
          o.v[j][4 * q + 1] = fmaf(a, wv.y, o.v[j][4 * q + 1]);
          o.v[j][4 * q + 2] = fmaf(a, wv.z, o.v[j][4 * q + 2]);
          o.v[j][4 * q + 3] = fmaf(a, wv.w, o.v[j][4 * q + 3]);
          if (noisy) {
            const float4 nvv = wn4[unit];
            o.v[j][4 * q + 0] = fmaf(bb, nvv.x, o.v[j][4 * q + 0]);
            o.v[j][4 * q + 1] = fmaf(bb, nvv.y, o.v[j][4 * q + 1]);
            o.v[j][4 * q + 2] = fmaf(bb, nvv.z, o.v[j][4 * q + 2]);
            o.v[j][4 * q + 3] = fmaf(bb, nvv.w, o.v[j][4 * q + 3]);
          }
        }
    }
    o.store(dx + (long long)n * D, D, lane);
    __syncwarp();
   }
   if (FAST && fast) {
     __syncwarp();
     gate_dx_block<NV, (FAST ? EB : 8), TBC>(wg, s_dl[warp], base, N, D, E, lane, reinterpret_cast<bf16*>(dx));
     __syncwarp();
   }
  }
}

// dW[e][d] partial over a chunk of tokens: part[chunk][E][D]; thread per column.
constexpr int RW_CHUNK = 64;
template <typename T, int EB>   // EB = expert-count bucket (8/16/32/64): a runtime bound made every thread issue 64
__global__ void __launch_bounds__(128)   // predicated FMAs per token even for E = 8
router_wgrad_kernel(const T* __restrict__ x, const float* __restrict__ dl, int N, int D, int E,
                    float* __restrict__ part) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_dl[32][EB];
  const int d = blockIdx.x * 128 + threadIdx.x;
  const int n0 = blockIdx.y * RW_CHUNK, n1 = min(N, n0 + RW_CHUNK);
  float acc[EB];
#pragma unroll
  for (int e = 0; e < EB; ++e) acc[e] = 0.f;
  for (int nb = n0; nb < n1; nb += 32) {
    const int cnt = min(32, n1 - nb);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * EB; i += blockDim.x) {
      const int tt = i / EB, ee = i % EB;
      s_dl[tt][ee] = ee < E ? dl[(long long)(nb + tt) * E + ee] : 0.f;
    }
    __syncthreads();
    if (d < D) {
      int t = 0;
      for (; t + 4 <= cnt; t += 4) {   // four independent row loads in flight
        float xv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) xv[q] = to_f32<T>(x[(long long)(nb + t + q) * D + d]);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int e = 0; e < EB; ++e) acc[e] = fmaf(s_dl[t + q][e], xv[q], acc[e]);
      }
      for (; t < cnt; ++t) {
        const float xv = to_f32<T>(x[(long long)(nb + t) * D + d]);
#pragma unroll
        for (int e = 0; e < EB; ++e) acc[e] = fmaf(s_dl[t][e], xv, acc[e]);
      }
    }
  }
  if (d < D) {
#pragma unroll
    for (int e = 0; e < EB; ++e)
      if (e < E) part[((long long)blockIdx.y * E + e) * D + d] = acc[e];
  }
}
__global__ void router_wgrad_reduce_kernel(const float* __restrict__ part, int chunks, int ED, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ED) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += part[(long long)c * ED + i];
  out[i] = s;
}

// grid of the token-blocked path: RT_TB tokens per warp and pass, two resident blocks per SM
inline int router_grid_blocked(int N) {
  int blocks = (N + RT_WARPS * RT_TB - 1) / (RT_WARPS * RT_TB);
  const int cap = num_sms() * 2;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : blocks;
}

inline int router_grid(int N) {
  int blocks = (N + RT_WARPS - 1) / RT_WARPS;
  const int cap = num_sms() * 6;   // ~2 tokens per warp at config-5 scale: more independent row loads in flight
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : blocks;
}

template <typename K>
int set_smem(K kern, size_t bytes) {
  // the kernels also hold a few KB of static shared memory: opt in well before the 48 KB default limit is reached
  if (bytes > 32 * 1024) B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

size_t b200_router_ws(int N, int E) {
  (void)E;
  return (size_t)router_grid(N) * 3 * RT_MAX_E * sizeof(float);
}

int b200_router_fwd(const void* x, int dtype, const float* w_gate, const float* w_noise, const float* eps,
                    float noise_std, float lb_weight, int N, int D, int E, int K, int32_t* idx, float* w,
                    float* topk_sum, float* probs, float* probs_noisy, float* counts, float* psum, float* loss,
                    float* noise_scale_mean, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(N > 0 && D > 0 && E > 0 && E <= RT_MAX_E && K > 0 && K <= E && K <= 32,
                 "router_fwd: bad shape N=%d D=%d E=%d K=%d (E<=64, K<=min(E,32))", N, D, E, K);
  B200_CHECK_ARG((eps == nullptr) || (w_noise != nullptr && probs_noisy != nullptr),
                 "router_fwd: noisy routing needs w_noise and probs_noisy");
  B200_CHECK_ARG(workspace_bytes >= b200_router_ws(N, E), "router_fwd: workspace too small");
  B200_CHECK_ARG(((uintptr_t)w_gate & 15) == 0 && ((uintptr_t)w_noise & 15) == 0 && D % 4 == 0,
                 "router_fwd: gate weights must be 16-byte aligned rows (D %% 4 == 0)");
  B200_CHECK_ARG(dtype == B200_BF16 ? D % 8 == 0 : D % 4 == 0, "router_fwd: D=%d not vectorisable", D);
  B200_CHECK_ARG(dtype == B200_BF16 ? RowRegs<bf16>::supported(D) : RowRegs<float>::supported(D),
                 "router_fwd: D=%d unsupported", D);
  const int nvb = dtype == B200_BF16 ? row_nv<bf16>(D) : row_nv<float>(D);
  const size_t smem = (size_t)E * nvb * 32 * (dtype == B200_BF16 ? 8 : 4) * sizeof(float) * (eps != nullptr ? 2 : 1);
  B200_CHECK_ARG(smem <= 200 * 1024, "router_fwd: gate weights (%zu B) do not fit in shared memory", smem);
  int blocks = router_grid(N);
  if (dtype == B200_BF16 && E <= 8 && eps == nullptr && nvb <= 4) blocks = router_grid_blocked(N);   // <= router_grid(N)
  float* part = (float*)workspace;
  if (dtype == B200_BF16 && eps == nullptr) {
    // statistics and the load-balance loss are folded by the last block of the same launch (ticket)
    const int mb = launch_router_fwd_mma((const bf16*)x, w_gate, N, D, E, K, idx, w, topk_sum, probs, part, lb_weight,
                                         counts, psum, loss, stream);
    if (mb == -2) return cuda_fail(cudaGetLastError(), "launch router_fwd_mma_kernel");
    if (mb > 0) {
      if (!router_fold_enabled()) {
        launch_kernel(router_finalize_kernel, dim3(1), dim3(1024), 0, stream, part, mb, N, E, lb_weight, counts, psum, loss,
                      (float*)nullptr);
        B200_LAUNCH_CHECK("router_finalize_kernel");
        count_launch(1);
      }
      count_launch(1);
      return 0;
    }
  }
  if (dtype == B200_BF16) {
    B200_NV_SWITCH(nvb, {
      if (E <= 8) {
        if (int rc = set_smem(router_fwd_kernel<bf16, NV, 8>, smem)) return rc;
        launch_kernel(router_fwd_kernel<bf16, NV, 8>, dim3(blocks), dim3(RT_WARPS * 32), smem, stream,
            (const bf16*)x, w_gate, w_noise, eps, noise_std, N, D, E, K, idx, w, topk_sum, probs, probs_noisy, part);
      } else {
        if (int rc = set_smem(router_fwd_kernel<bf16, NV, 0>, smem)) return rc;
        launch_kernel(router_fwd_kernel<bf16, NV, 0>, dim3(blocks), dim3(RT_WARPS * 32), smem, stream,
            (const bf16*)x, w_gate, w_noise, eps, noise_std, N, D, E, K, idx, w, topk_sum, probs, probs_noisy, part);
      }
    });
  } else {
    B200_NV_SWITCH(nvb, {
      if (int rc = set_smem(router_fwd_kernel<float, NV, 0>, smem)) return rc;
      launch_kernel(router_fwd_kernel<float, NV, 0>, dim3(blocks), dim3(RT_WARPS * 32), smem, stream,
          (const float*)x, w_gate, w_noise, eps, noise_std, N, D, E, K, idx, w, topk_sum, probs, probs_noisy, part);
    });
  }
  B200_LAUNCH_CHECK("router_fwd_kernel");
  launch_kernel(router_finalize_kernel, dim3(1), dim3(1024), 0, stream, part, blocks, N, E, lb_weight, counts, psum, loss,
                                                     eps != nullptr ? noise_scale_mean : nullptr);
  B200_LAUNCH_CHECK("router_finalize_kernel");
  count_launch(2);
  return 0;
}

size_t b200_router_bwd_ws(int N, int D, int E) {
  const size_t chunks = (size_t)(N + RW_CHUNK - 1) / RW_CHUNK;
  // dl [N,E] + du [N,E] + partial weight grads [chunks][E][D]
  return ((size_t)2 * N * E + chunks * (size_t)E * D) * sizeof(float);
}

int b200_router_bwd(const void* x, int dtype, const float* w_gate, const float* w_noise, const float* eps,
                    float noise_std, float lb_weight, int N, int D, int E, int K, const int32_t* idx, const float* w,
                    const float* topk_sum, const float* probs, const float* probs_noisy, const float* counts,
                    const float* d_w, const float* d_loss, const float* d_probs, void* dx, float* d_w_gate,
                    float* d_w_noise, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(N > 0 && D > 0 && E > 0 && E <= RT_MAX_E && K > 0 && K <= E, "router_bwd: bad shape");
  B200_CHECK_ARG(workspace_bytes >= b200_router_bwd_ws(N, D, E), "router_bwd: workspace too small");
  B200_CHECK_ARG(((uintptr_t)w_gate & 15) == 0 && ((uintptr_t)w_noise & 15) == 0 && D % 4 == 0,
                 "router_bwd: gate weights must be 16-byte aligned rows (D %% 4 == 0)");
  B200_CHECK_ARG(dtype == B200_BF16 ? D % 8 == 0 : D % 4 == 0, "router_bwd: D=%d not vectorisable", D);
  const bool noisy = eps != nullptr;
  B200_CHECK_ARG(!noisy || (w_noise != nullptr && probs_noisy != nullptr && d_w_noise != nullptr),
                 "router_bwd: noisy routing needs w_noise, probs_noisy, d_w_noise");
  B200_CHECK_ARG(dtype == B200_BF16 ? RowRegs<bf16>::supported(D) : RowRegs<float>::supported(D),
                 "router_bwd: D=%d unsupported", D);
  const int nvb = dtype == B200_BF16 ? row_nv<bf16>(D) : row_nv<float>(D);
  const size_t smem = (size_t)E * nvb * 32 * (dtype == B200_BF16 ? 8 : 4) * sizeof(float) * (noisy ? 2 : 1);
  B200_CHECK_ARG(smem <= 200 * 1024, "router_bwd: gate weights do not fit in shared memory");
  float* dl = (float*)workspace;
  float* du = dl + (size_t)N * E;
  float* part = du + (size_t)N * E;
  int blocks = router_grid(N);
  if (dtype == B200_BF16 && E <= 8 && !noisy && nvb <= 4) blocks = router_grid_blocked(N);
  const int chunks = (N + RW_CHUNK - 1) / RW_CHUNK;
  dim3 wg_grid((D + 127) / 128, chunks);
  const int ED = E * D;
  int mma_rc = -1;
  if (dtype == B200_BF16 && !noisy) {
    mma_rc = launch_router_bwd_mma(w_gate, lb_weight, N, D, E, K, idx, w, topk_sum, probs, counts, d_w, d_loss, d_probs,
                                   (bf16*)dx, dl, stream);
    if (mma_rc == -2) return cuda_fail(cudaGetLastError(), "launch router_bwd_mma_kernel");
  }
  if (mma_rc == 0) {
    // dl and dx are done: only the weight gradient below remains
  } else if (dtype == B200_BF16) {
    B200_NV_SWITCH(nvb, {
      if (E <= 8) {
        if (int rc = set_smem(router_bwd_kernel<bf16, NV, 8>, smem)) return rc;
        launch_kernel(router_bwd_kernel<bf16, NV, 8>, dim3(blocks), dim3(RT_WARPS * 32), smem, stream,
            (const bf16*)x, w_gate, w_noise, eps, noise_std, lb_weight, N, D, E, K, idx, w, topk_sum, probs,
            probs_noisy, counts, d_w, d_loss, d_probs, (bf16*)dx, dl, du);
      } else {
        if (int rc = set_smem(router_bwd_kernel<bf16, NV, 0>, smem)) return rc;
        launch_kernel(router_bwd_kernel<bf16, NV, 0>, dim3(blocks), dim3(RT_WARPS * 32), smem, stream,
            (const bf16*)x, w_gate, w_noise, eps, noise_std, lb_weight, N, D, E, K, idx, w, topk_sum, probs,
            probs_noisy, counts, d_w, d_loss, d_probs, (bf16*)dx, dl, du);
      }
    });
  } else {
    B200_NV_SWITCH(nvb, {
      if (int rc = set_smem(router_bwd_kernel<float, NV, 0>, smem)) return rc;
      launch_kernel(router_bwd_kernel<float, NV, 0>, dim3(blocks), dim3(RT_WARPS * 32), smem, stream, 
          (const float*)x, w_gate, w_noise, eps, noise_std, lb_weight, N, D, E, K, idx, w, topk_sum, probs, probs_noisy,
          counts, d_w, d_loss, d_probs, (float*)dx, dl, du);
    });
  }
  B200_LAUNCH_CHECK("router_bwd_kernel");
  count_launch();
  for (int pass = 0; pass < (noisy ? 2 : 1); ++pass) {
    const float* g = pass == 0 ? dl : du;
    float* out = pass == 0 ? d_w_gate : d_w_noise;
#define B200_RW(TT, EBV) launch_kernel(router_wgrad_kernel<TT, EBV>, dim3(wg_grid), dim3(128), 0, stream, (const TT*)x, g, N, D, E, part)
    if (dtype == B200_BF16) {
      if (E <= 8) B200_RW(bf16, 8); else if (E <= 16) B200_RW(bf16, 16); else if (E <= 32) B200_RW(bf16, 32); else B200_RW(bf16, 64);
    } else {
      if (E <= 8) B200_RW(float, 8); else if (E <= 16) B200_RW(float, 16); else if (E <= 32) B200_RW(float, 32); else B200_RW(float, 64);
    }
#undef B200_RW
    B200_LAUNCH_CHECK("router_wgrad_kernel");
    count_launch();
    if (int rc = launch_partial_reduce(part, chunks, 1, ED, nullptr, 1, out, stream)) return rc;
  }
  return 0;
}

}  // extern "C"
