// Fused multi-head attention core on the CUDA cores (fp32 math): softmax(scale * Q K^T + key_pad) V with
// online softmax over key chunks; the [B,H,T,S] score tensor lives only in shared memory / registers.
// This is the full-precision engine (fp32 validation mode) and the first-generation bf16 path.
//   forward : grid (B*H, q-tiles); one warp per query row, keys of a chunk spread over lanes.
//   backward: two passes over (query tile, key chunk) pairs with 4x4 register-tiled shared-memory GEMMs:
//             pass 0 owns a key chunk and accumulates dK,dV over query tiles; pass 1 owns a query tile and
//             accumulates dQ over key chunks (S/P recomputed from the saved log-sum-exp; no atomics).
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

namespace {

constexpr int FA_WARPS = 8;
constexpr int FA_SC = 128;  // keys per chunk (forward)
constexpr int FA_TQ = 32;   // query rows per CTA (forward): 4 per warp

template <typename T>
__global__ void __launch_bounds__(FA_WARPS * 32)
attn_fwd_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k, int ldk, const T* __restrict__ v, int ldv,
                const uint8_t* __restrict__ key_pad, T* __restrict__ o, int ldo, float* __restrict__ lse, int H, int Tq,
                int S, int dh, float scale, const unsigned long long* drop_state, float drop_p, unsigned int drop_site, int causal) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const DropState ds = drop_load(drop_state, drop_p, drop_site);
  const unsigned long long s_pad = (unsigned long long)((S + 127) / 128) * 128;   // attention dropout index pitch
  const int LD = dh + 1;
  float* Ks = sm;                          // [FA_SC][LD]
  float* Vs = Ks + FA_SC * LD;             // [FA_SC][LD]
  float* Qs = Vs + FA_SC * LD;             // [FA_WARPS][dh]
  float* Ps = Qs + FA_WARPS * dh;          // [FA_WARPS][FA_SC]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_base = blockIdx.y * FA_TQ;
  constexpr int RPW = FA_TQ / FA_WARPS;  // rows per warp
  constexpr int DPL = 4;                 // output dims per lane (dh <= 128)

  float m_run[RPW], l_run[RPW], acc[RPW][DPL];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    m_run[r] = -INFINITY;
    l_run[r] = 0.f;
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[r][d] = 0.f;
  }

  for (int s0 = 0; s0 < S; s0 += FA_SC) {
    const int sc = min(FA_SC, S - s0);
    __syncthreads();
    for (int i = threadIdx.x; i < sc * dh; i += blockDim.x) {
      const int j = i / dh, d = i % dh;
      const long long row = (long long)b * S + s0 + j;
      Ks[j * LD + d] = to_f32<T>(k[row * ldk + h * dh + d]);
      Vs[j * LD + d] = to_f32<T>(v[row * ldv + h * dh + d]);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int t = t_base + warp * RPW + r;
      if (t >= Tq) continue;  // warp-uniform
      float* qs = Qs + warp * dh;
      float* ps = Ps + warp * FA_SC;
      __syncwarp();
      for (int d = lane; d < dh; d += 32)
        qs[d] = to_f32<T>(q[((long long)b * Tq + t) * ldq + h * dh + d]) * scale;
      __syncwarp();
      float sj[FA_SC / 32];
      float cmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < FA_SC / 32; ++jj) {
        const int j = lane + 32 * jj;
        float s = -INFINITY;
        if (j < sc && !(key_pad != nullptr && key_pad[(long long)b * S + s0 + j]) && !(causal && s0 + j > t)) {
          s = 0.f;
          const float* kr = Ks + j * LD;
          for (int d = 0; d < dh; ++d) s = fmaf(qs[d], kr[d], s);
        }
        sj[jj] = s;
        cmax = fmaxf(cmax, s);
      }
      cmax = warp_max(cmax);
      const float m_new = fmaxf(m_run[r], cmax);
      const float corr = (m_run[r] == -INFINITY) ? 0.f : __expf(m_run[r] - m_new);
      float psum = 0.f;
#pragma unroll
      for (int jj = 0; jj < FA_SC / 32; ++jj) {
        const float p = (sj[jj] == -INFINITY) ? 0.f : __expf(sj[jj] - m_new);
        float keep = 1.f;
        if (ds.on)
          keep = drop_scale1(ds, ((unsigned long long)blockIdx.x * Tq + t) * s_pad + (unsigned long long)(s0 + lane + 32 * jj));
        ps[lane + 32 * jj] = p * keep;   // dropped probabilities feed P.V; the normaliser uses the undropped sum
        psum += p;
      }
      psum = warp_sum(psum);
      l_run[r] = l_run[r] * corr + psum;
      m_run[r] = m_new;
      __syncwarp();
#pragma unroll
      for (int dd = 0; dd < DPL; ++dd) {
        const int d = lane + 32 * dd;
        float a = acc[r][dd] * corr;
        if (d < dh) {
          for (int j = 0; j < sc; ++j) a = fmaf(ps[j], Vs[j * LD + d], a);
        }
        acc[r][dd] = a;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int t = t_base + warp * RPW + r;
    if (t >= Tq) continue;
    const float inv = 1.f / l_run[r];
#pragma unroll
    for (int dd = 0; dd < DPL; ++dd) {
      const int d = lane + 32 * dd;
      if (d < dh) o[((long long)b * Tq + t) * ldo + h * dh + d] = from_f32<T>(acc[r][dd] * inv);
    }
    if (lane == 0) lse[((long long)b * H + h) * Tq + t] = m_run[r] + __logf(l_run[r]);
  }
}

// ---- backward ------------------------------------------------------------------------------------------

// C[i][j] (+)= sum_k A(i,k) * B(k,j) on shared-memory operands, 4x4 register tiles, 256 threads.
template <bool ACCUM>
__device__ __forceinline__ void smem_gemm(float* C, int ldc, int M, int N, const float* A, int a_si, int a_sk,
                                          const float* B, int b_sk, int b_sj, int Kred, float alpha) {
  const int tiles_n = N / 4, tiles = (M / 4) * tiles_n;
  for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
    const int i0 = (t / tiles_n) * 4, j0 = (t % tiles_n) * 4;
    float c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    for (int kk = 0; kk < Kred; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = A[(i0 + i) * a_si + kk * a_sk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = B[kk * b_sk + (j0 + j) * b_sj];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = fmaf(a[i], b[j], c[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float* dst = C + (i0 + i) * ldc + j0 + j;
        if (ACCUM) *dst += alpha * c[i][j];
        else *dst = alpha * c[i][j];
      }
  }
}

// PASS 0: blockIdx.y = key chunk -> dK, dV.   PASS 1: blockIdx.y = query tile -> dQ.
template <typename T, int PASS, int BT>
__global__ void __launch_bounds__(256)
attn_bwd_kernel(const T* __restrict__ q, int ldq, const T* __restrict__ k, int ldk, const T* __restrict__ v, int ldv,
                const uint8_t* __restrict__ key_pad, const T* __restrict__ o, int ldo, const T* __restrict__ d_o,
                int lddo, const float* __restrict__ lse, T* __restrict__ dq, int lddq, T* __restrict__ dk, int lddk,
                T* __restrict__ dv, int lddv, int H, int Tq, int S, int dh, float scale,
                const unsigned long long* drop_state, float drop_p, unsigned int drop_site, int causal) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const DropState ds = drop_load(drop_state, drop_p, drop_site);
  const unsigned long long s_pad = (unsigned long long)((S + 127) / 128) * 128;
  const int LD = dh + 1, LS = BT + 1;
  float* Qs = sm;                 // [BT][LD]
  float* dOs = Qs + BT * LD;      // [BT][LD]
  float* Ks = dOs + BT * LD;      // [BT][LD]
  float* Vs = Ks + BT * LD;       // [BT][LD]
  float* Ps = Vs + BT * LD;       // [BT][LS]   P, later dS
  float* dPs = Ps + BT * LS;      // [BT][LS]
  float* Acc0 = dPs + BT * LS;    // [BT][LD]   PASS0: dK   PASS1: dQ
  float* Acc1 = Acc0 + BT * LD;   // [BT][LD]   PASS0: dV
  float* delta = Acc1 + BT * LD;  // [BT]
  float* lses = delta + BT;       // [BT]

  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int own0 = blockIdx.y * BT;  // first key (PASS 0) or first query row (PASS 1) owned by this CTA
  const int inner_len = PASS == 0 ? Tq : S;

  for (int i = threadIdx.x; i < BT * LD; i += blockDim.x) {
    Acc0[i] = 0.f;
    if (PASS == 0) Acc1[i] = 0.f;
  }

  auto load_q_tile = [&](int t0) {
    for (int i = threadIdx.x; i < BT * dh; i += blockDim.x) {
      const int r = i / dh, d = i % dh;
      const int t = t0 + r;
      float qv = 0.f, gv = 0.f;
      if (t < Tq) {
        qv = to_f32<T>(q[((long long)b * Tq + t) * ldq + h * dh + d]);
        gv = to_f32<T>(d_o[((long long)b * Tq + t) * lddo + h * dh + d]);
      }
      Qs[r * LD + d] = qv;
      dOs[r * LD + d] = gv;
    }
    // delta[r] = sum_d dO * O ; 4 threads per row
    {
      const int r = threadIdx.x / 4, part = threadIdx.x % 4;
      const int t = t0 + r;
      float s = 0.f;
      if (t < Tq && r < BT)
        for (int d = part; d < dh; d += 4)
          s = fmaf(to_f32<T>(d_o[((long long)b * Tq + t) * lddo + h * dh + d]),
                   to_f32<T>(o[((long long)b * Tq + t) * ldo + h * dh + d]), s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (part == 0 && r < BT) {
        delta[r] = s;
        lses[r] = (t < Tq) ? lse[((long long)b * H + h) * Tq + t] : 0.f;
      }
    }
  };
  auto load_k_tile = [&](int s0) {
    for (int i = threadIdx.x; i < BT * dh; i += blockDim.x) {
      const int r = i / dh, d = i % dh;
      const int s = s0 + r;
      float kv = 0.f, vv = 0.f;
      if (s < S) {
        kv = to_f32<T>(k[((long long)b * S + s) * ldk + h * dh + d]);
        vv = to_f32<T>(v[((long long)b * S + s) * ldv + h * dh + d]);
      }
      Ks[r * LD + d] = kv;
      Vs[r * LD + d] = vv;
    }
  };

  if (PASS == 0) load_k_tile(own0); else load_q_tile(own0);

  for (int in0 = 0; in0 < inner_len; in0 += BT) {
    __syncthreads();
    if (PASS == 0) load_q_tile(in0); else load_k_tile(in0);
    __syncthreads();
    const int t0 = PASS == 0 ? in0 : own0, s0 = PASS == 0 ? own0 : in0;
    // S = scale * Q K^T ; dP = dO V^T
    smem_gemm<false>(Ps, LS, BT, BT, Qs, LD, 1, Ks, 1, LD, dh, scale);
    smem_gemm<false>(dPs, LS, BT, BT, dOs, LD, 1, Vs, 1, LD, dh, 1.f);
    __syncthreads();
    // P = exp(S - lse) (0 where masked / out of range); dS = P * (dP - delta)
    for (int i = threadIdx.x; i < BT * BT; i += blockDim.x) {
      const int r = i / BT, c = i % BT;
      const int t = t0 + r, s = s0 + c;
      float p = 0.f;
      if (t < Tq && s < S && !(key_pad != nullptr && key_pad[(long long)b * S + s]) && !(causal && s > t))
        p = __expf(Ps[r * LS + c] - lses[r]);
      float keep = 1.f;
      if (ds.on && p != 0.f)
        keep = drop_scale1(ds, ((unsigned long long)blockIdx.x * Tq + t) * s_pad + (unsigned long long)s);
      const float dsv = p * (dPs[r * LS + c] * keep - delta[r]);
      Ps[r * LS + c] = p * keep;
      dPs[r * LS + c] = dsv;
    }
    __syncthreads();
    if (PASS == 0) {
      // dV[j][d] += sum_i P[i][j] dO[i][d] ; dK[j][d] += scale * sum_i dS[i][j] Q[i][d]
      smem_gemm<true>(Acc1, LD, BT, dh, Ps, 1, LS, dOs, LD, 1, BT, 1.f);
      smem_gemm<true>(Acc0, LD, BT, dh, dPs, 1, LS, Qs, LD, 1, BT, scale);
    } else {
      // dQ[i][d] += scale * sum_j dS[i][j] K[j][d]
      smem_gemm<true>(Acc0, LD, BT, dh, dPs, LS, 1, Ks, LD, 1, BT, scale);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BT * dh; i += blockDim.x) {
    const int r = i / dh, d = i % dh;
    if (PASS == 0) {
      const int s = own0 + r;
      if (s < S) {
        dk[((long long)b * S + s) * lddk + h * dh + d] = from_f32<T>(Acc0[r * LD + d]);
        dv[((long long)b * S + s) * lddv + h * dh + d] = from_f32<T>(Acc1[r * LD + d]);
      }
    } else {
      const int t = own0 + r;
      if (t < Tq) dq[((long long)b * Tq + t) * lddq + h * dh + d] = from_f32<T>(Acc0[r * LD + d]);
    }
  }
}

size_t fwd_smem(int dh) { return ((size_t)2 * FA_SC * (dh + 1) + FA_WARPS * dh + FA_WARPS * FA_SC) * sizeof(float); }
size_t bwd_smem(int dh, int bt) { return ((size_t)6 * bt * (dh + 1) + 2 * bt * (bt + 1) + 2 * bt) * sizeof(float); }

template <typename K>
int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

}  // namespace

// tensor-core path (attn_tc.cu)
bool attn_tc_supported(int T, int S, int dh, int ldq, int ldk, int ldv, const void* q, const void* k, const void* v);
int launch_attn_fwd_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const uint8_t* key_pad,
                       int causal, void* o, int ldo, float* lse, int B, int H, int T, int S, int dh, float scale,
                       const unsigned long long* drop_state, float drop_p, unsigned int drop_site, cudaStream_t stream);
int launch_attn_bwd_tc(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const uint8_t* key_pad,
                       int causal, const void* o, int ldo, const void* d_o, int lddo, const float* lse, void* dq, int lddq, void* dk,
                       int lddk, void* dv, int lddv, int B, int H, int T, int S, int dh, float scale,
                       const unsigned long long* drop_state, float drop_p, unsigned int drop_site, cudaStream_t stream);

static bool use_tc_attention() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VQA_ATTN");   // "simt" forces the CUDA-core kernels (A/B testing)
    v = (e != nullptr && strcmp(e, "simt") == 0) ? 0 : 1;
  }
  return v == 1;
}
}  // namespace b200

using namespace b200;

extern "C" {

int b200_attn_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const uint8_t* key_pad,
                  int causal, void* o, int ldo, float* lse, int B, int H, int T, int S, int dh, float scale, int dtype,
                  const b200_dropout_t* drop, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool don = drop != nullptr && drop->p > 0.f;
  const unsigned long long* dst = don ? drop->rng_state : nullptr;
  const float dpp = don ? drop->p : 0.f;
  const unsigned int dsite = don ? drop->site : 0u;
  B200_CHECK_ARG(B > 0 && H > 0 && T > 0 && S > 0 && dh > 0 && dh <= 128 && dh % 4 == 0,
                 "attn_fwd: bad shape B=%d H=%d T=%d S=%d dh=%d (dh<=128, dh%%4==0)", B, H, T, S, dh);
  B200_CHECK_ARG(!causal || T == S, "attn_fwd: the causal mask needs T == S (got %d, %d)", T, S);
  if (dtype == B200_BF16 && use_tc_attention() && attn_tc_supported(T, S, dh, ldq, ldk, ldv, q, k, v) &&
      ldo % 8 == 0 && ((uintptr_t)o & 15) == 0)
    return launch_attn_fwd_tc(q, ldq, k, ldk, v, ldv, key_pad, causal, o, ldo, lse, B, H, T, S, dh, scale, dst, dpp, dsite,
                              stream);
  dim3 grid(B * H, (T + FA_TQ - 1) / FA_TQ);
  const size_t smem = fwd_smem(dh);
  if (dtype == B200_BF16) {
    if (int rc = set_smem(attn_fwd_kernel<bf16>, smem)) return rc;
    launch_kernel(attn_fwd_kernel<bf16>, dim3(grid), dim3(FA_WARPS * 32), smem, stream, (const bf16*)q, ldq, (const bf16*)k, ldk,
                                                                 (const bf16*)v, ldv, key_pad, (bf16*)o, ldo, lse, H, T,
                                                                 S, dh, scale, dst, dpp, dsite, causal);
  } else {
    if (int rc = set_smem(attn_fwd_kernel<float>, smem)) return rc;
    launch_kernel(attn_fwd_kernel<float>, dim3(grid), dim3(FA_WARPS * 32), smem, stream, (const float*)q, ldq, (const float*)k, ldk,
                                                                  (const float*)v, ldv, key_pad, (float*)o, ldo, lse, H,
                                                                  T, S, dh, scale, dst, dpp, dsite, causal);
  }
  B200_LAUNCH_CHECK("attn_fwd_kernel");
  count_launch();
  return 0;
}

int b200_attn_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const uint8_t* key_pad,
                  int causal, const void* o, int ldo, const void* d_o, int lddo, const float* lse, void* dq, int lddq, void* dk,
                  int lddk, void* dv, int lddv, int B, int H, int T, int S, int dh, float scale, int dtype,
                  const b200_dropout_t* drop, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool don = drop != nullptr && drop->p > 0.f;
  const unsigned long long* dst = don ? drop->rng_state : nullptr;
  const float dpp = don ? drop->p : 0.f;
  const unsigned int dsite = don ? drop->site : 0u;
  B200_CHECK_ARG(B > 0 && H > 0 && T > 0 && S > 0 && dh > 0 && dh <= 128 && dh % 4 == 0,
                 "attn_bwd: bad shape B=%d H=%d T=%d S=%d dh=%d (dh<=128, dh%%4==0)", B, H, T, S, dh);
  B200_CHECK_ARG(!causal || T == S, "attn_bwd: the causal mask needs T == S (got %d, %d)", T, S);
  if (dtype == B200_BF16 && use_tc_attention() && attn_tc_supported(T, S, dh, ldq, ldk, ldv, q, k, v) &&
      ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 &&
      (((uintptr_t)o | (uintptr_t)d_o | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15) == 0)
    return launch_attn_bwd_tc(q, ldq, k, ldk, v, ldv, key_pad, causal, o, ldo, d_o, lddo, lse, dq, lddq, dk, lddk, dv, lddv, B,
                              H, T, S, dh, scale, dst, dpp, dsite, stream);
  const int bt = bwd_smem(dh, 64) <= 200 * 1024 ? 64 : 32;
  const size_t smem = bwd_smem(dh, bt);
  dim3 g0(B * H, (S + bt - 1) / bt), g1(B * H, (T + bt - 1) / bt);
#define B200_ATTN_BWD(TT, BTV)                                                                                   \
  do {                                                                                                           \
    if (int rc = set_smem(attn_bwd_kernel<TT, 0, BTV>, smem)) return rc;                                         \
    if (int rc = set_smem(attn_bwd_kernel<TT, 1, BTV>, smem)) return rc;                                         \
    launch_kernel(attn_bwd_kernel<TT, 0, BTV>, dim3(g0), dim3(256), smem, stream,                                                       \
        (const TT*)q, ldq, (const TT*)k, ldk, (const TT*)v, ldv, key_pad, (const TT*)o, ldo, (const TT*)d_o,     \
        lddo, lse, (TT*)dq, lddq, (TT*)dk, lddk, (TT*)dv, lddv, H, T, S, dh, scale, dst, dpp, dsite, causal);  \
    B200_LAUNCH_CHECK("attn_bwd_kernel<0>");                                                                     \
    launch_kernel(attn_bwd_kernel<TT, 1, BTV>, dim3(g1), dim3(256), smem, stream,                                                       \
        (const TT*)q, ldq, (const TT*)k, ldk, (const TT*)v, ldv, key_pad, (const TT*)o, ldo, (const TT*)d_o,     \
        lddo, lse, (TT*)dq, lddq, (TT*)dk, lddk, (TT*)dv, lddv, H, T, S, dh, scale, dst, dpp, dsite, causal);  \
    B200_LAUNCH_CHECK("attn_bwd_kernel<1>");                                                                     \
  } while (0)
  if (dtype == B200_BF16) {
    if (bt == 64) B200_ATTN_BWD(bf16, 64); else B200_ATTN_BWD(bf16, 32);
  } else {
    if (bt == 64) B200_ATTN_BWD(float, 64); else B200_ATTN_BWD(float, 32);
  }
#undef B200_ATTN_BWD
  count_launch(2);
  return 0;
}

}  // extern "C"
