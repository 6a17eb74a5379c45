// Generative-decoder glue (generative_vqa_model.py:342-451,572-591): token embedding + sinusoidal
// positions + dropout in one gather pass, its scatter-add backward, and the label-smoothed cross-entropy over the
// 64 000-way vocabulary (one streaming pass for the loss, one for the logit gradient).  All HBM-bound.
#include "rowops.cuh"

namespace b200 {

namespace {

// out[n,:] = dropout(table[ids[n],:] + pos[n % T,:])
template <typename T>
__global__ void __launch_bounds__(256)
embed_fwd_kernel(const int* __restrict__ ids, const T* __restrict__ table, const float* __restrict__ pos,
                 T* __restrict__ out, int N, int Tlen, int D, int V, const unsigned long long* drop_state, float drop_p,
                 unsigned int drop_site) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const DropState ds = drop_load(drop_state, drop_p, drop_site);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nv = D / VT;
  for (int n = warp; n < N; n += nwarps) {
    int id = ids[n];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const T* src = table + (long long)id * D;
    const float* pr = pos + (long long)(n % Tlen) * D;
    for (int vi = lane; vi < nv; vi += 32) {
      Vec16<T> t;
      t.load(src + vi * VT);
      float pv[VT];
      load_param<VT>(pr, vi, pv);
#pragma unroll
      for (int u = 0; u < VT; ++u) t.v[u] += pv[u];
      if (ds.on) {
        const unsigned long long base = (unsigned long long)n * D + (unsigned long long)vi * VT;
#pragma unroll
        for (int u = 0; u < VT; u += 4) {
          float sc[4];
          drop_scales4(ds, (base + u) >> 2, sc);
#pragma unroll
          for (int q = 0; q < 4; ++q) t.v[u + q] *= sc[q];
        }
      }
      t.store(out + (long long)n * D + vi * VT);
    }
  }
}

// dtable[ids[n],:] += dropout_mask * dout[n,:]   (fp32 atomics: several positions may hold the same token)
template <typename T>
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const int* __restrict__ ids, const T* __restrict__ dout, float* __restrict__ dtable, int N, int D,
                 int V, const unsigned long long* drop_state, float drop_p, unsigned int drop_site) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const DropState ds = drop_load(drop_state, drop_p, drop_site);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nv = D / VT;
  for (int n = warp; n < N; n += nwarps) {
    int id = ids[n];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    float* dst = dtable + (long long)id * D;
    for (int vi = lane; vi < nv; vi += 32) {
      Vec16<T> t;
      t.load(dout + (long long)n * D + vi * VT);
      if (ds.on) {
        const unsigned long long base = (unsigned long long)n * D + (unsigned long long)vi * VT;
#pragma unroll
        for (int u = 0; u < VT; u += 4) {
          float sc[4];
          drop_scales4(ds, (base + u) >> 2, sc);
#pragma unroll
          for (int q = 0; q < 4; ++q) t.v[u + q] *= sc[q];
        }
      }
#pragma unroll
      for (int u = 0; u < VT; ++u) atomicAdd(dst + vi * VT + u, t.v[u]);
    }
  }
}

// ---- cross-entropy ---------------------------------------------------------------------------------------------
constexpr int CE_THREADS = 512;
constexpr float CE_LOG2E = 1.4426950408889634f, CE_LN2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One block per row: online (max, sum-exp) in the log2 domain and the plain sum of the logits in ONE pass over the
// row.  loss_row = lse - (1 - eps) * z[label] - eps * mean(z)   (nn.CrossEntropyLoss with label_smoothing = eps).
template <typename T>
__global__ void __launch_bounds__(CE_THREADS)
ce_fwd_kernel(const T* __restrict__ logits, long long ld, const int* __restrict__ labels, int C, int ignore_index,
              float smoothing, float* __restrict__ loss_rows, float* __restrict__ lse_out) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  __shared__ float s_m[CE_THREADS / 32], s_s[CE_THREADS / 32], s_z[CE_THREADS / 32];
  const int r = blockIdx.x;
  const T* row = logits + (long long)r * ld;
  const int label = labels[r];
  float m = -INFINITY, s = 0.f, zs = 0.f;
  const int nvec = C / VT;
  for (int vi = threadIdx.x; vi < nvec; vi += CE_THREADS) {
    Vec16<T> t;
    t.load(row + vi * VT);
    float vm = t.v[0];
#pragma unroll
    for (int u = 1; u < VT; ++u) vm = fmaxf(vm, t.v[u]);
    vm *= CE_LOG2E;
    if (vm > m) {
      s *= fast_ex2(m - vm);
      m = vm;
    }
#pragma unroll
    for (int u = 0; u < VT; ++u) {
      s += fast_ex2(fmaf(t.v[u], CE_LOG2E, -m));
      zs += t.v[u];
    }
  }
  for (int c = nvec * VT + threadIdx.x; c < C; c += CE_THREADS) {   // tail when C is not a multiple of the vector
    const float z = to_f32<T>(row[c]);
    const float zl = z * CE_LOG2E;
    if (zl > m) {
      s *= fast_ex2(m - zl);
      m = zl;
    }
    s += fast_ex2(zl - m);
    zs += z;
  }
  // warp, then block combine of (m, s) pairs
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    const float nm = fmaxf(m, om);
    s = (m == -INFINITY ? 0.f : s * fast_ex2(m - nm)) + (om == -INFINITY ? 0.f : os * fast_ex2(om - nm));
    m = nm;
    zs += __shfl_xor_sync(0xffffffffu, zs, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_m[warp] = m; s_s[warp] = s; s_z[warp] = zs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = -INFINITY;
    for (int w = 0; w < CE_THREADS / 32; ++w) M = fmaxf(M, s_m[w]);
    float S = 0.f, Z = 0.f;
    for (int w = 0; w < CE_THREADS / 32; ++w) {
      if (s_m[w] != -INFINITY) S += s_s[w] * fast_ex2(s_m[w] - M);
      Z += s_z[w];
    }
    const float lse = (M + log2f(S)) * CE_LN2;
    lse_out[r] = lse;
    float loss = 0.f;
    if (label != ignore_index && label >= 0 && label < C) {
      const float zy = to_f32<T>(row[label]);
      loss = lse - (1.f - smoothing) * zy - smoothing * (Z / (float)C);
    }
    loss_rows[r] = loss;
  }
}

// loss = sum(loss_rows) / n_valid (fixed order: deterministic); n_valid is kept for the backward pass
__global__ void __launch_bounds__(1024)
ce_finalize_kernel(const float* __restrict__ loss_rows, const int* __restrict__ labels, int R, int C, int ignore_index,
                   float* __restrict__ loss, float* __restrict__ n_valid) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_l[32], s_n[32];
  float l = 0.f, n = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const int y = labels[r];
    if (y != ignore_index && y >= 0 && y < C) {
      l += loss_rows[r];
      n += 1.f;
    }
  }
  l = warp_sum(l);
  n = warp_sum(n);
  if ((threadIdx.x & 31) == 0) { s_l[threadIdx.x >> 5] = l; s_n[threadIdx.x >> 5] = n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float L = 0.f, Nn = 0.f;
    for (int w = 0; w < 32; ++w) { L += s_l[w]; Nn += s_n[w]; }
    n_valid[0] = Nn;
    loss[0] = Nn > 0.f ? L / Nn : 0.f;     // all rows ignored: torch gives nan; 0 keeps a training loop alive
  }
}

// dlogits[r,c] = g * (softmax(z)[c] - (1 - eps) [c == y] - eps / C), g = dloss / n_valid; ignored rows get zeros
template <typename T>
__global__ void __launch_bounds__(CE_THREADS)
ce_bwd_kernel(const T* __restrict__ logits, long long ld, const int* __restrict__ labels, const float* __restrict__ lse,
              int C, int ignore_index, float smoothing, const float* __restrict__ dloss,
              const float* __restrict__ n_valid, T* __restrict__ dlogits, long long ldd) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const int r = blockIdx.x;
  const T* row = logits + (long long)r * ld;
  T* out = dlogits + (long long)r * ldd;
  const int label = labels[r];
  const bool valid = label != ignore_index && label >= 0 && label < C;
  const float nv_ = n_valid[0];
  const float g = (valid && nv_ > 0.f) ? dloss[0] / nv_ : 0.f;
  const float l2 = lse[r] * CE_LOG2E;
  const float un = smoothing / (float)C, on = 1.f - smoothing;
  const int nvec = C / VT;
  for (int vi = threadIdx.x; vi < nvec; vi += CE_THREADS) {
    Vec16<T> t;
    t.load(row + vi * VT);
    const int c0 = vi * VT;
#pragma unroll
    for (int u = 0; u < VT; ++u) {
      const float p = fast_ex2(fmaf(t.v[u], CE_LOG2E, -l2));
      t.v[u] = g * (p - un - (c0 + u == label ? on : 0.f));
    }
    t.store(out + vi * VT);
  }
  for (int c = nvec * VT + threadIdx.x; c < C; c += CE_THREADS) {
    const float p = fast_ex2(fmaf(to_f32<T>(row[c]), CE_LOG2E, -l2));
    out[c] = from_f32<T>(g * (p - un - (c == label ? on : 0.f)));
  }
}

inline int warp_grid(int rows) {
  int blocks = (rows + 7) / 8;
  const int cap = num_sms() * 8;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : blocks;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_embed_fwd(const int32_t* ids, const void* table, const float* pos, void* out, int N, int T, int D, int V,
                   int dtype, const b200_dropout_t* drop, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(N > 0 && T > 0 && D > 0 && V > 0 && D % (dtype == B200_BF16 ? 8 : 4) == 0,
                 "embed_fwd: bad shape N=%d T=%d D=%d V=%d", N, T, D, V);
  const bool don = drop != nullptr && drop->p > 0.f;
  const unsigned long long* dst = don ? drop->rng_state : nullptr;
  const float dp = don ? drop->p : 0.f;
  const unsigned int dsite = don ? drop->site : 0u;
  if (dtype == B200_BF16)
    launch_kernel(embed_fwd_kernel<bf16>, dim3(warp_grid(N)), dim3(256), 0, stream, ids, (const bf16*)table, pos,
                  (bf16*)out, N, T, D, V, dst, dp, dsite);
  else
    launch_kernel(embed_fwd_kernel<float>, dim3(warp_grid(N)), dim3(256), 0, stream, ids, (const float*)table, pos,
                  (float*)out, N, T, D, V, dst, dp, dsite);
  B200_LAUNCH_CHECK("embed_fwd_kernel");
  count_launch();
  return 0;
}

int b200_embed_bwd(const int32_t* ids, const void* dout, float* dtable, int N, int D, int V, int dtype,
                   const b200_dropout_t* drop, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(N > 0 && D > 0 && V > 0 && D % (dtype == B200_BF16 ? 8 : 4) == 0, "embed_bwd: bad shape N=%d D=%d", N, D);
  const bool don = drop != nullptr && drop->p > 0.f;
  const unsigned long long* dst = don ? drop->rng_state : nullptr;
  const float dp = don ? drop->p : 0.f;
  const unsigned int dsite = don ? drop->site : 0u;
  if (dtype == B200_BF16)
    launch_kernel(embed_bwd_kernel<bf16>, dim3(warp_grid(N)), dim3(256), 0, stream, ids, (const bf16*)dout, dtable, N, D,
                  V, dst, dp, dsite);
  else
    launch_kernel(embed_bwd_kernel<float>, dim3(warp_grid(N)), dim3(256), 0, stream, ids, (const float*)dout, dtable, N,
                  D, V, dst, dp, dsite);
  B200_LAUNCH_CHECK("embed_bwd_kernel");
  count_launch();
  return 0;
}

int b200_ce_fwd(const void* logits, long long ld, const int32_t* labels, int R, int C, int ignore_index, float smoothing,
                int dtype, float* loss_rows, float* lse, float* loss, float* n_valid, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(R > 0 && C > 0 && ld >= C && smoothing >= 0.f && smoothing < 1.f, "ce_fwd: bad shape R=%d C=%d", R, C);
  B200_CHECK_ARG(((uintptr_t)logits & 15) == 0 && (ld * (dtype == B200_BF16 ? 2 : 4)) % 16 == 0,
                 "ce_fwd: logits rows must be 16-byte aligned");
  if (dtype == B200_BF16)
    launch_kernel(ce_fwd_kernel<bf16>, dim3(R), dim3(CE_THREADS), 0, stream, (const bf16*)logits, ld, labels, C,
                  ignore_index, smoothing, loss_rows, lse);
  else
    launch_kernel(ce_fwd_kernel<float>, dim3(R), dim3(CE_THREADS), 0, stream, (const float*)logits, ld, labels, C,
                  ignore_index, smoothing, loss_rows, lse);
  B200_LAUNCH_CHECK("ce_fwd_kernel");
  launch_kernel(ce_finalize_kernel, dim3(1), dim3(1024), 0, stream, (const float*)loss_rows, labels, R, C, ignore_index,
                loss, n_valid);
  B200_LAUNCH_CHECK("ce_finalize_kernel");
  count_launch(2);
  return 0;
}

int b200_ce_bwd(const void* logits, long long ld, const int32_t* labels, const float* lse, int R, int C,
                int ignore_index, float smoothing, int dtype, const float* dloss, const float* n_valid, void* dlogits,
                long long ldd, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(R > 0 && C > 0 && ld >= C && ldd >= C, "ce_bwd: bad shape R=%d C=%d", R, C);
  B200_CHECK_ARG((((uintptr_t)logits | (uintptr_t)dlogits) & 15) == 0 && (ld * (dtype == B200_BF16 ? 2 : 4)) % 16 == 0 &&
                     (ldd * (dtype == B200_BF16 ? 2 : 4)) % 16 == 0,
                 "ce_bwd: logits rows must be 16-byte aligned");
  if (dtype == B200_BF16)
    launch_kernel(ce_bwd_kernel<bf16>, dim3(R), dim3(CE_THREADS), 0, stream, (const bf16*)logits, ld, labels, lse, C,
                  ignore_index, smoothing, dloss, n_valid, (bf16*)dlogits, ldd);
  else
    launch_kernel(ce_bwd_kernel<float>, dim3(R), dim3(CE_THREADS), 0, stream, (const float*)logits, ld, labels, lse, C,
                  ignore_index, smoothing, dloss, n_valid, (float*)dlogits, ldd);
  B200_LAUNCH_CHECK("ce_bwd_kernel");
  count_launch();
  return 0;
}

}  // extern "C"
