// Row-in-registers helpers for the HBM-bound row kernels (LayerNorm, combine): one warp owns one row of
// D elements, moved with 16-byte vector accesses; lane l holds vectors l, l+32, l+64, ...
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int ROW_MAXV = 8;  // vectors per lane: D <= 8*32*8 = 2048 (bf16) or 8*32*4 = 1024 (fp32)

// NV = vectors per lane is a compile-time parameter: sizing every row kernel for the worst case (8) cost ~100
// registers per thread and capped the HBM-bound kernels at 25% occupancy (ncu, profiles/r01e).
template <typename T, int NV = ROW_MAXV>
struct RowRegs {
  static constexpr int VT = Vec16<T>::N;
  float v[NV][VT];

  __host__ __device__ __forceinline__ static bool supported(int D) { return D % VT == 0 && D / VT <= 32 * ROW_MAXV; }

  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int u = 0; u < VT; ++u) v[i][u] = 0.f;
  }
  __device__ __forceinline__ void load(const T* row, int D, int lane) {
    const int nv = D / VT;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nv) {
        Vec16<T> t;
        t.load(row + vi * VT);
#pragma unroll
        for (int u = 0; u < VT; ++u) v[i][u] = t.v[u];
      } else {
#pragma unroll
        for (int u = 0; u < VT; ++u) v[i][u] = 0.f;
      }
    }
  }
  // v += scale * row
  __device__ __forceinline__ void axpy(const T* row, float scale, int D, int lane) {
    const int nv = D / VT;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nv) {
        Vec16<T> t;
        t.load(row + vi * VT);
#pragma unroll
        for (int u = 0; u < VT; ++u) v[i][u] = fmaf(scale, t.v[u], v[i][u]);
      }
    }
  }
  __device__ __forceinline__ void store(T* row, int D, int lane) const {
    const int nv = D / VT;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + 32 * i;
      if (vi < nv) {
        Vec16<T> t;
#pragma unroll
        for (int u = 0; u < VT; ++u) t.v[u] = v[i][u];
        t.store(row + vi * VT);
      }
    }
  }
  __device__ __forceinline__ float sum(int D, int lane) const {
    const int nv = D / VT;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < nv) {
#pragma unroll
        for (int u = 0; u < VT; ++u) s += v[i][u];
      }
    return warp_sum(s);
  }
  __device__ __forceinline__ float sumsq_centered(float mean, int D, int lane) const {
    const int nv = D / VT;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < nv) {
#pragma unroll
        for (int u = 0; u < VT; ++u) {
          const float d = v[i][u] - mean;
          s = fmaf(d, d, s);
        }
      }
    return warp_sum(s);
  }
};

// fp32 parameter vector slice aligned with RowRegs' element mapping
template <int VT>
__device__ __forceinline__ void load_param(const float* p, int vi, float (&out)[VT]) {
#pragma unroll
  for (int u = 0; u < VT; u += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + vi * VT + u));
    out[u] = t.x; out[u + 1] = t.y; out[u + 2] = t.z; out[u + 3] = t.w;
  }
}

}  // namespace b200

namespace b200 {
// vectors-per-lane bucket for a row of D elements of T (supported instantiations: 1,2,3,4,6,8)
template <typename T>
inline int row_nv(int D) {
  const int need = (D / Vec16<T>::N + 31) / 32;
  const int buckets[6] = {1, 2, 3, 4, 6, 8};
  for (int b : buckets)
    if (need <= b) return b;
  return 8;
}
}  // namespace b200

#define B200_NV_SWITCH(nv, ...)                                   \
  switch (nv) {                                                   \
    case 1: { constexpr int NV = 1; __VA_ARGS__; } break;         \
    case 2: { constexpr int NV = 2; __VA_ARGS__; } break;         \
    case 3: { constexpr int NV = 3; __VA_ARGS__; } break;         \
    case 4: { constexpr int NV = 4; __VA_ARGS__; } break;         \
    case 6: { constexpr int NV = 6; __VA_ARGS__; } break;         \
    default: { constexpr int NV = 8; __VA_ARGS__; } break;        \
  }
