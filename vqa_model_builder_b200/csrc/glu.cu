// Gated linear unit of GatedLinearExpert (reference: src/modeling/moe/expert_types.py:448-515):
//   [value | gate] = fc1(x)  ->  h = dropout(value * sigmoid(gate)),   forward and backward.
// HBM-bound elementwise kernels between the two (grouped) GEMMs of the expert: 16-byte accesses, one thread per
// 16-byte packet of h.  With a tile map (grouped experts) the rows of unused 128-row tiles are skipped.
#include "rowops.cuh"

namespace b200 {
namespace {

__device__ __forceinline__ float sigmoid_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

template <typename T>
__device__ __forceinline__ void keep_scales(const DropState& ds, unsigned long long base, float (&sc)[Vec16<T>::N]) {
  constexpr int VT = Vec16<T>::N;
  if (!ds.on) {
#pragma unroll
    for (int u = 0; u < VT; ++u) sc[u] = 1.f;
    return;
  }
  if (VT == 8) {
    float s8[8];
    drop_scales8(ds, base >> 3, s8);
#pragma unroll
    for (int u = 0; u < VT; ++u) sc[u] = s8[u];
  } else {
    float s4[4];
    drop_scales4(ds, base >> 2, s4);
#pragma unroll
    for (int u = 0; u < VT; ++u) sc[u] = s4[u % 4];
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
glu_fwd_kernel(const T* __restrict__ pre, int ld_pre, T* __restrict__ h, int R, int F, const int* __restrict__ tile_group,
               const unsigned long long* drop_state, float drop_p, unsigned int drop_site) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const DropState ds = drop_load(drop_state, drop_p, drop_site);
  const int vpr = F / VT;                                  // packets per row
  const long long total = (long long)R * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / vpr), c = (int)(i - (long long)r * vpr) * VT;
    if (tile_group != nullptr && tile_group[r / B200_GROUP_TILE] < 0) continue;
    Vec16<T> v, g, o;
    v.load(pre + (long long)r * ld_pre + c);
    g.load(pre + (long long)r * ld_pre + F + c);
    float sc[VT];
    keep_scales<T>(ds, (unsigned long long)r * F + c, sc);
#pragma unroll
    for (int u = 0; u < VT; ++u) o.v[u] = v.v[u] * sigmoid_fast(g.v[u]) * sc[u];
    o.store(h + (long long)r * F + c);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
glu_bwd_kernel(const T* __restrict__ dh, const T* __restrict__ pre, int ld_pre, T* __restrict__ dpre, int R, int F,
               const int* __restrict__ tile_group, const unsigned long long* drop_state, float drop_p,
               unsigned int drop_site) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const DropState ds = drop_load(drop_state, drop_p, drop_site);
  const int vpr = F / VT;
  const long long total = (long long)R * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / vpr), c = (int)(i - (long long)r * vpr) * VT;
    if (tile_group != nullptr && tile_group[r / B200_GROUP_TILE] < 0) continue;
    Vec16<T> d, v, g, dv, dg;
    d.load(dh + (long long)r * F + c);
    v.load(pre + (long long)r * ld_pre + c);
    g.load(pre + (long long)r * ld_pre + F + c);
    float sc[VT];
    keep_scales<T>(ds, (unsigned long long)r * F + c, sc);
#pragma unroll
    for (int u = 0; u < VT; ++u) {
      const float s = sigmoid_fast(g.v[u]);
      const float t = d.v[u] * sc[u];
      dv.v[u] = t * s;
      dg.v[u] = t * v.v[u] * s * (1.0f - s);
    }
    dv.store(dpre + (long long)r * ld_pre + c);
    dg.store(dpre + (long long)r * ld_pre + F + c);
  }
}

inline int glu_grid(long long packets) {
  long long b = (packets + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int b200_glu_fwd(const void* pre, int ld_pre, void* h, int R, int F, int dtype, const int32_t* tile_group,
                 const b200_dropout_t* drop, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int vt = dtype == B200_BF16 ? 8 : 4;
  B200_CHECK_ARG(R > 0 && F > 0 && F % vt == 0 && ld_pre >= 2 * F && ld_pre % vt == 0, "glu_fwd: bad shape R=%d F=%d ld=%d",
                 R, F, ld_pre);
  const bool don = drop != nullptr && drop->p > 0.f;
  const unsigned long long* dst = don ? drop->rng_state : nullptr;
  const float dp = don ? drop->p : 0.f;
  const unsigned int site = don ? drop->site : 0u;
  const int grid = glu_grid((long long)R * (F / vt));
  if (dtype == B200_BF16)
    launch_kernel(glu_fwd_kernel<bf16>, dim3(grid), dim3(256), 0, stream, (const bf16*)pre, ld_pre, (bf16*)h, R, F,
                  tile_group, dst, dp, site);
  else
    launch_kernel(glu_fwd_kernel<float>, dim3(grid), dim3(256), 0, stream, (const float*)pre, ld_pre, (float*)h, R, F,
                  tile_group, dst, dp, site);
  B200_LAUNCH_CHECK("glu_fwd_kernel");
  count_launch();
  return 0;
}

int b200_glu_bwd(const void* dh, const void* pre, int ld_pre, void* dpre, int R, int F, int dtype,
                 const int32_t* tile_group, const b200_dropout_t* drop, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int vt = dtype == B200_BF16 ? 8 : 4;
  B200_CHECK_ARG(R > 0 && F > 0 && F % vt == 0 && ld_pre >= 2 * F && ld_pre % vt == 0, "glu_bwd: bad shape R=%d F=%d ld=%d",
                 R, F, ld_pre);
  const bool don = drop != nullptr && drop->p > 0.f;
  const unsigned long long* dst = don ? drop->rng_state : nullptr;
  const float dp = don ? drop->p : 0.f;
  const unsigned int site = don ? drop->site : 0u;
  const int grid = glu_grid((long long)R * (F / vt));
  if (dtype == B200_BF16)
    launch_kernel(glu_bwd_kernel<bf16>, dim3(grid), dim3(256), 0, stream, (const bf16*)dh, (const bf16*)pre, ld_pre,
                  (bf16*)dpre, R, F, tile_group, dst, dp, site);
  else
    launch_kernel(glu_bwd_kernel<float>, dim3(grid), dim3(256), 0, stream, (const float*)dh, (const float*)pre, ld_pre,
                  (float*)dpre, R, F, tile_group, dst, dp, site);
  B200_LAUNCH_CHECK("glu_bwd_kernel");
  count_launch();
  return 0;
}

}  // extern "C"
