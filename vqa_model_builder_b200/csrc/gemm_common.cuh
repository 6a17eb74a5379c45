// GEMM argument block + epilogue shared by the tcgen05 (bf16) and SIMT (fp32) GEMM kernels.
#pragma once
#include "common.cuh"

namespace b200 {

enum { GEMM_DENSE = 0, GEMM_GROUP_ROWS = 1, GEMM_GROUP_WGRAD = 2 };

struct GemmArgs {
  int M, N, K;          // output rows bound, output cols bound, reduction length (dense / per group)
  int mode;             // GEMM_*
  int epi, act;         // B200_EPI_*, B200_ACT_*
  int out_f32;          // output element type is fp32 (else same as input type)
  int ldo, ld_aux;      // row pitches (elements)
  void* out;
  const void* aux_in;
  void* aux_out;
  const float* bias;
  const int* tile_group;      // GROUP_ROWS: expert of each 128-row tile
  const int* rows_used;       // GROUP_ROWS (optional, device): rows in use; 128-row tiles beyond it are not visited
  const int* group_off;       // GROUP_WGRAD: padded row offsets [G+1]
  int b_group_rows;           // GROUP_ROWS: per-group coordinate offset in B (N for K-layout, K for MN-layout)
  long long out_group_elems;  // GROUP_WGRAD: output elements per group
  int k_splits;               // DENSE: split-K factor (EPI_ACCUM only when > 1)
  // SIMT-only raw operand description (tcgen05 kernel uses tensor maps instead)
  const void* A; const void* B;
  long long sa_m, sa_k, sb_n, sb_k;  // element strides
  long long b_group_elems;           // GROUP_ROWS: elements per group in B
  // dropout on the activation output (EPI_ACT) / its backward (EPI_DACT); element index = row*ldo + col
  const unsigned long long* drop_state; float drop_p; unsigned int drop_site;
};

// multiply CNT consecutive columns (col0 % 4 == 0, CNT % 4 == 0) of one row by their keep-scales
template <int CNT>
__device__ __forceinline__ void apply_dropout_row(const DropState& d, long long row, int ldo, int col0, float (&v)[CNT]) {
  const unsigned long long base = (unsigned long long)row * (unsigned long long)ldo + (unsigned long long)col0;
  if ((base & 7ull) == 0 && CNT % 8 == 0) {
#pragma unroll
    for (int j = 0; j < CNT; j += 8) {
      float sc[8];
      drop_scales8(d, (base + j) >> 3, sc);
#pragma unroll
      for (int q = 0; q < 8; ++q) v[j + q] *= sc[q];
    }
  } else if ((base & 3ull) == 0) {
#pragma unroll
    for (int j = 0; j < CNT; j += 4) {
      float sc[4];
      drop_scales4(d, (base + j) >> 2, sc);
#pragma unroll
      for (int q = 0; q < 4; ++q) v[j + q] *= sc[q];
    }
  } else {
#pragma unroll
    for (int j = 0; j < CNT; ++j) v[j] *= drop_scale1(d, base + j);
  }
}

// Apply the epilogue to CNT consecutive columns [col0, col0+CNT) of one output row and store them.
// T = activation storage type (bf16 or float) used for aux tensors and non-fp32 outputs.
template <typename T, int CNT>
__device__ __forceinline__ void epilogue_store(const GemmArgs& p, int group, long long row, int col0,
                                               float (&acc)[CNT], bool row_ok) {
  if (!row_ok || col0 >= p.N) return;
  const bool full = (col0 + CNT <= p.N);
  const float* bias = p.bias;
  if (bias != nullptr && p.mode == GEMM_GROUP_ROWS) bias += (long long)group * p.N;
  const long long obase = (p.mode == GEMM_GROUP_WGRAD ? (long long)group * p.out_group_elems : 0ll) +
                          row * (long long)p.ldo + col0;
  const long long abase = row * (long long)p.ld_aux + col0;

  if (bias != nullptr) {
#pragma unroll
    for (int j = 0; j < CNT; ++j)
      if (full || col0 + j < p.N) acc[j] += __ldg(bias + col0 + j);
  }

  constexpr int VT = Vec16<T>::N;
  if (p.epi == B200_EPI_ACT_D) {      // out = dropout(act(pre)), aux_out = act'(pre) * keep-scale
    float dv[CNT];
#pragma unroll
    for (int j = 0; j < CNT; ++j) {
      dv[j] = act_bwd(acc[j], p.act);
      acc[j] = act_fwd(acc[j], p.act);
    }
    if (p.drop_state != nullptr && p.drop_p > 0.f) {
      const DropState ds = drop_load(p.drop_state, p.drop_p, p.drop_site);
      apply_dropout_row<CNT>(ds, row, p.ldo, col0, acc);
      apply_dropout_row<CNT>(ds, row, p.ldo, col0, dv);
    }
    if (p.aux_out != nullptr) {
      T* ao = reinterpret_cast<T*>(p.aux_out) + abase;
#pragma unroll
      for (int j = 0; j < CNT; ++j)
        if (col0 + j < p.N) ao[j] = from_f32<T>(dv[j]);
    }
  } else if (p.epi == B200_EPI_ACT) {
    if (p.aux_out != nullptr) {
      T* ao = reinterpret_cast<T*>(p.aux_out) + abase;
      if (full && (CNT % VT == 0) && (p.ld_aux % VT == 0) && (col0 % VT == 0)) {
#pragma unroll
        for (int j = 0; j < CNT; j += VT) {
          Vec16<T> v;
#pragma unroll
          for (int u = 0; u < VT; ++u) v.v[u] = acc[j + u];
          v.store(ao + j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < CNT; ++j)
          if (col0 + j < p.N) ao[j] = from_f32<T>(acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < CNT; ++j) acc[j] = act_fwd(acc[j], p.act);
    if (p.drop_state != nullptr && p.drop_p > 0.f) {
      const DropState ds = drop_load(p.drop_state, p.drop_p, p.drop_site);
      apply_dropout_row<CNT>(ds, row, p.ldo, col0, acc);
    }
  } else if (p.epi == B200_EPI_ADD || p.epi == B200_EPI_DACT || p.epi == B200_EPI_MUL) {
    const T* ai = reinterpret_cast<const T*>(p.aux_in) + abase;
    float aux[CNT];
    if (full && (CNT % VT == 0) && (p.ld_aux % VT == 0) && (col0 % VT == 0)) {
#pragma unroll
      for (int j = 0; j < CNT; j += VT) {
        Vec16<T> v;
        v.load(ai + j);
#pragma unroll
        for (int u = 0; u < VT; ++u) aux[j + u] = v.v[u];
      }
    } else {
#pragma unroll
      for (int j = 0; j < CNT; ++j) aux[j] = (col0 + j < p.N) ? to_f32<T>(ai[j]) : 0.f;
    }
    if (p.epi == B200_EPI_ADD) {
      if (p.drop_state != nullptr && p.drop_p > 0.f) {   // out = dropout(acc + bias) + residual
        const DropState ds = drop_load(p.drop_state, p.drop_p, p.drop_site);
        apply_dropout_row<CNT>(ds, row, p.ldo, col0, acc);
      }
#pragma unroll
      for (int j = 0; j < CNT; ++j) acc[j] += aux[j];
    } else if (p.epi == B200_EPI_MUL) {
#pragma unroll
      for (int j = 0; j < CNT; ++j) acc[j] *= aux[j];
    } else {
#pragma unroll
      for (int j = 0; j < CNT; ++j) acc[j] *= act_bwd(aux[j], p.act);
      if (p.drop_state != nullptr && p.drop_p > 0.f) {
        const DropState ds = drop_load(p.drop_state, p.drop_p, p.drop_site);
        apply_dropout_row<CNT>(ds, row, p.ldo, col0, acc);
      }
    }
  }

  if (p.epi == B200_EPI_ACCUM) {
    float* o = reinterpret_cast<float*>(p.out) + obase;
#pragma unroll
    for (int j = 0; j < CNT; ++j)
      if (full || col0 + j < p.N) atomicAdd(o + j, acc[j]);
    return;
  }

  if (p.out_f32) {
    float* o = reinterpret_cast<float*>(p.out) + obase;
    if (full && (CNT % 4 == 0) && (p.ldo % 4 == 0) && (col0 % 4 == 0) &&
        (p.mode != GEMM_GROUP_WGRAD || p.out_group_elems % 4 == 0)) {
#pragma unroll
      for (int j = 0; j < CNT; j += 4)
        *reinterpret_cast<float4*>(o + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < CNT; ++j)
        if (col0 + j < p.N) o[j] = acc[j];
    }
  } else {
    T* o = reinterpret_cast<T*>(p.out) + obase;
    if (full && (CNT % VT == 0) && (p.ldo % VT == 0) && (col0 % VT == 0)) {
#pragma unroll
      for (int j = 0; j < CNT; j += VT) {
        Vec16<T> v;
#pragma unroll
        for (int u = 0; u < VT; ++u) v.v[u] = acc[j + u];
        v.store(o + j);
      }
    } else {
#pragma unroll
      for (int j = 0; j < CNT; ++j)
        if (col0 + j < p.N) o[j] = from_f32<T>(acc[j]);
    }
  }
}

// implemented in gemm_tc.cu / gemm_simt.cu
// K-layout operand: [mn_extent, k_extent] row-major; MN-layout operand: [k_extent, mn_extent] row-major.
int launch_gemm_tc(const void* A, int lda, int a_layout, long long a_mn_extent, long long a_k_extent,
                   const void* B, int ldb, int b_layout, long long b_mn_extent, long long b_k_extent,
                   GemmArgs args, int grid_m_tiles, int groups, cudaStream_t stream);
int launch_gemm_simt(GemmArgs args, int grid_m_tiles, int groups, cudaStream_t stream);
// 128B-swizzled 2-D bf16 tensor map (box = 64 inner elements x box_outer rows); map_out is a CUtensorMap*
int make_tma_map_bf16(void* map_out, const void* ptr, long long inner, long long outer, long long pitch,
                      int box_outer);

}  // namespace b200
