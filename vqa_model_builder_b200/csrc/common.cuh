// Shared device/host helpers for the b200vqa kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200vqa.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------
// Error plumbing: every C-ABI entry returns 0 or a B200_ERR_* code; the message is thread-local.
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define B200_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      ::b200::set_error(__VA_ARGS__);        \
      return B200_ERR_INVALID;               \
    }                                        \
  } while (0)

#define B200_CUDA(expr)                                          \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return ::b200::cuda_fail(_e, #expr);  \
  } while (0)

#define B200_LAUNCH_CHECK(name)                                         \
  do {                                                                  \
    cudaError_t _e = cudaGetLastError();                                \
    if (_e != cudaSuccess) return ::b200::cuda_fail(_e, "launch " name); \
  } while (0)

int num_sms();
void count_launch(int n = 1);

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel of the library is launched with the programmatic-stream-
// serialization attribute and starts with griddepcontrol.launch_dependents + griddepcontrol.wait, so the launch
// latency and the prologue (barrier init, TMEM allocation, descriptor prefetch) of kernel N+1 overlap the tail of
// kernel N.  All global-memory traffic stays behind the wait.  B200VQA_PDL=0 disables the attribute.
// ---------------------------------------------------------------------------------------------
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KP, typename... A>
inline cudaError_t launch_kernel(void (*kern)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KP>(args)...);
}

// ---------------------------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// A 16-byte packet of T viewed as floats: 4 floats (T=float) or 8 floats (T=bf16).
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const bf16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  __device__ __forceinline__ void store(bf16* p) const {
    uint4 t;
    t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
    t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// activations (FeedForwardExpert supports gelu(erf) | relu | silu | tanh, expert_types.py:44-51)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case B200_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
    case B200_ACT_RELU: return x > 0.f ? x : 0.f;
    case B200_ACT_SILU: return x / (1.0f + __expf(-x));
    case B200_ACT_TANH: return tanhf(x);
    default: return x;
  }
}
__device__ __forceinline__ float act_bwd(float x, int act) {  // d act(x) / dx
  switch (act) {
    case B200_ACT_GELU: {
      float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
      float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
      return cdf + x * pdf;
    }
    case B200_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case B200_ACT_SILU: {
      float s = 1.0f / (1.0f + __expf(-x));
      return s * (1.0f + x * (1.0f - s));
    }
    case B200_ACT_TANH: {
      float t = tanhf(x);
      return 1.0f - t * t;
    }
    default: return 1.f;
  }
}

// ---------------------------------------------------------------------------------------------
// Counter-based RNG for dropout: Philox4x32-7 keyed by (seed, stream), counter = element index/8.  Seven rounds are
// the fewest with which Philox4x32 passes BigCrush (Salmon et al., SC'11, table 2; curand's default of 10 is a safety
// margin): the generator sits in the epilogues of the GEMMs and in the attention softmax, where its ~9 instructions
// per round and 8 elements are a visible share of the per-element budget (ncu, profiles/r02l).
// The same (seed, stream, index) reproduces the same keep-mask in backward without storing it.
// ---------------------------------------------------------------------------------------------
constexpr int PHILOX_ROUNDS = 7;
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr, uint32_t stream) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0x9E3779B9u;
#pragma unroll
  for (int r = 0; r < PHILOX_ROUNDS; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// Dropout state resolved once per kernel from the ABI struct (device memory read for seed / offset).
struct DropState {
  unsigned long long key;   // seed mixed with the step offset
  float p, inv_keep;
  unsigned int site, thr;   // element dropped when its 16-bit uniform < thr  (thr = round(p * 65536))
  bool on;
};
__device__ __forceinline__ DropState drop_load(const unsigned long long* rng_state, float p, unsigned int site) {
  DropState d;
  d.on = (rng_state != nullptr) && (p > 0.f);
  d.p = p;
  d.inv_keep = d.on ? 1.0f / (1.0f - p) : 1.0f;
  d.site = site;
  d.thr = (unsigned int)fminf(p * 65536.0f + 0.5f, 65535.0f);
  d.key = d.on ? (rng_state[0] ^ (rng_state[1] * 0x9E3779B97F4A7C15ull)) : 0ull;
  return d;
}
// One Philox4x32-7 call serves 8 consecutive elements (16 random bits each: p is resolved to 2^-16):
// keep-scales of the elements 8*idx8 .. 8*idx8+7.
__device__ __forceinline__ void drop_scales8(const DropState& d, unsigned long long idx8, float (&s)[8]) {
  const uint4 r = philox4x32(d.key, idx8, d.site);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s[2 * i] = (w[i] & 0xFFFFu) < d.thr ? 0.f : d.inv_keep;
    s[2 * i + 1] = (w[i] >> 16) < d.thr ? 0.f : d.inv_keep;
  }
}
// keep-scales of the 4 elements 4*idx4 .. 4*idx4+3 (one half of the 8-element group)
__device__ __forceinline__ void drop_scales4(const DropState& d, unsigned long long idx4, float (&s)[4]) {
  const uint4 r = philox4x32(d.key, idx4 >> 1, d.site);
  const bool hi = (idx4 & 1ull) != 0;
  const uint32_t w0 = hi ? r.z : r.x, w1 = hi ? r.w : r.y;
  s[0] = (w0 & 0xFFFFu) < d.thr ? 0.f : d.inv_keep;
  s[1] = (w0 >> 16) < d.thr ? 0.f : d.inv_keep;
  s[2] = (w1 & 0xFFFFu) < d.thr ? 0.f : d.inv_keep;
  s[3] = (w1 >> 16) < d.thr ? 0.f : d.inv_keep;
}
__device__ __forceinline__ float drop_scale1(const DropState& d, unsigned long long idx) {
  float s[4];
  drop_scales4(d, idx >> 2, s);
  const int k = (int)(idx & 3);
  return k == 0 ? s[0] : k == 1 ? s[1] : k == 2 ? s[2] : s[3];
}

}  // namespace b200
