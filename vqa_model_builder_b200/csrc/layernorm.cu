// Fused residual-add + LayerNorm, forward and backward.  HBM-bound: one warp per row, 16-byte accesses,
// statistics in fp32.  Optional per-expert affine parameters through the tile->expert map.
#include "rowops.cuh"

namespace b200 {

namespace {

constexpr int LN_WARPS = 8;
constexpr int LN_ROWS_PER_BLOCK = 8;  // one row per warp; divides B200_GROUP_TILE so a block sees one expert

template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32)
add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
                  const float* __restrict__ beta, const int* __restrict__ tile_group, float eps,
                  T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int R, int D) {
  constexpr int VT = Vec16<T>::N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = D / VT;
  for (int i = 0; i < LN_ROWS_PER_BLOCK / LN_WARPS; ++i) {
    const int r = blockIdx.x * LN_ROWS_PER_BLOCK + warp * (LN_ROWS_PER_BLOCK / LN_WARPS) + i;
    if (r >= R) return;
    int g = 0;
    if (tile_group != nullptr) {
      g = tile_group[r / B200_GROUP_TILE];
      if (g < 0) continue;
    }
    RowRegs<T> row;
    row.load(x + (long long)r * D, D, lane);
    if (res != nullptr) row.axpy(res + (long long)r * D, 1.f, D, lane);
    const float mean = row.sum(D, lane) / D;
    const float var = row.sumsq_centered(mean, D, lane) / D;
    const float rstd = rsqrtf(var + eps);
    const float* gm = gamma + (long long)g * D;
    const float* bt = beta + (long long)g * D;
#pragma unroll
    for (int j = 0; j < ROW_MAXV; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nv) {
        float gv[VT], bv[VT];
        load_param<VT>(gm, vi, gv);
        load_param<VT>(bt, vi, bv);
#pragma unroll
        for (int u = 0; u < VT; ++u) row.v[j][u] = (row.v[j][u] - mean) * rstd * gv[u] + bv[u];
      }
    }
    row.store(y + (long long)r * D, D, lane);
    if (lane == 0) {
      mean_out[r] = mean;
      rstd_out[r] = rstd;
    }
  }
}

// dsum = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma.
// Per-block partial dgamma/dbeta are written to part[block][2][D]; a second kernel reduces them per group.
template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32)
add_ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                  const float* __restrict__ gamma, const int* __restrict__ tile_group, T* __restrict__ dsum,
                  float* __restrict__ part, int R, int D) {
  constexpr int VT = Vec16<T>::N;
  extern __shared__ float red[];  // [LN_WARPS][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = D / VT;
  float dg[ROW_MAXV][VT], db[ROW_MAXV][VT];
#pragma unroll
  for (int j = 0; j < ROW_MAXV; ++j)
#pragma unroll
    for (int u = 0; u < VT; ++u) dg[j][u] = db[j][u] = 0.f;

  for (int i = 0; i < LN_ROWS_PER_BLOCK / LN_WARPS; ++i) {
    const int r = blockIdx.x * LN_ROWS_PER_BLOCK + warp * (LN_ROWS_PER_BLOCK / LN_WARPS) + i;
    if (r >= R) break;
    int g = 0;
    if (tile_group != nullptr) {
      g = tile_group[r / B200_GROUP_TILE];
      if (g < 0) continue;
    }
    RowRegs<T> xr, gr;
    xr.load(x + (long long)r * D, D, lane);
    if (res != nullptr) xr.axpy(res + (long long)r * D, 1.f, D, lane);
    gr.load(dy + (long long)r * D, D, lane);
    const float mean = mean_in[r], rstd = rstd_in[r];
    const float* gm = gamma + (long long)g * D;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < ROW_MAXV; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nv) {
        float gv[VT];
        load_param<VT>(gm, vi, gv);
#pragma unroll
        for (int u = 0; u < VT; ++u) {
          const float xhat = (xr.v[j][u] - mean) * rstd;
          const float d = gr.v[j][u];
          dg[j][u] = fmaf(d, xhat, dg[j][u]);
          db[j][u] += d;
          const float gg = d * gv[u];
          xr.v[j][u] = xhat;
          gr.v[j][u] = gg;
          s1 += gg;
          s2 = fmaf(gg, xhat, s2);
        }
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int j = 0; j < ROW_MAXV; ++j)
#pragma unroll
      for (int u = 0; u < VT; ++u) gr.v[j][u] = rstd * (gr.v[j][u] - s1 - xr.v[j][u] * s2);
    gr.store(dsum + (long long)r * D, D, lane);
  }

  // cross-warp reduction of the parameter-gradient partials
#pragma unroll
  for (int j = 0; j < ROW_MAXV; ++j) {
    const int vi = lane + 32 * j;
    if (vi < nv) {
#pragma unroll
      for (int u = 0; u < VT; ++u) {
        red[(warp * 2 + 0) * D + vi * VT + u] = dg[j][u];
        red[(warp * 2 + 1) * D + vi * VT + u] = db[j][u];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
    const int which = c / D, d = c % D;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) s += red[(w * 2 + which) * D + d];
    part[((long long)blockIdx.x * 2 + which) * D + d] = s;
  }
}

// out_z[g][c] = sum over partial blocks b of group g of part[b][z][c]   (z = blockIdx.z selects dgamma / dbeta).
// 32 columns x 8 partial-lanes per block: loads are independent across threads instead of one serial chain.
__global__ void __launch_bounds__(256)
partial_reduce_kernel(const float* __restrict__ part, int blocks, int rows_per_block, int width, int nz,
                      const int* __restrict__ tile_group, float* __restrict__ out0, float* __restrict__ out1) {
  __shared__ float red[8][33];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + x;
  const int g = blockIdx.y, z = blockIdx.z;
  float s = 0.f;
  if (c < width) {
    for (int b = y; b < blocks; b += 8) {
      const int bg = tile_group ? tile_group[(b * rows_per_block) / B200_GROUP_TILE] : 0;
      if (bg == g) s += part[((long long)b * nz + z) * width + c];
    }
  }
  red[y][x] = s;
  __syncthreads();
  if (y == 0 && c < width) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][x];
    (z == 0 ? out0 : out1)[(long long)g * width + c] = t;
  }
}

}  // namespace

// shared with dispatch.cu (combine backward uses the same partial layout)
int launch_ln_param_reduce(const float* part, int blocks, int rows_per_block, int D, const int* tile_group, int G,
                           float* dgamma, float* dbeta, cudaStream_t stream) {
  dim3 grid((D + 31) / 32, G, 2);
  partial_reduce_kernel<<<grid, 256, 0, stream>>>(part, blocks, rows_per_block, D, 2, tile_group, dgamma, dbeta);
  B200_LAUNCH_CHECK("partial_reduce_kernel");
  count_launch();
  return 0;
}

// single-array variant (column sums)
int launch_partial_reduce(const float* part, int blocks, int rows_per_block, int width, const int* tile_group, int G,
                          float* out, cudaStream_t stream) {
  dim3 grid((width + 31) / 32, G, 1);
  partial_reduce_kernel<<<grid, 256, 0, stream>>>(part, blocks, rows_per_block, width, 1, tile_group, out, out);
  B200_LAUNCH_CHECK("partial_reduce_kernel");
  count_launch();
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_add_ln_fwd(const void* x, const void* res, const float* gamma, const float* beta,
                    const int32_t* tile_group, float eps, void* y, float* mean, float* rstd, int R, int D, int dtype,
                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(R > 0 && D > 0, "add_ln_fwd: bad shape R=%d D=%d", R, D);
  const int blocks = (R + LN_ROWS_PER_BLOCK - 1) / LN_ROWS_PER_BLOCK;
  if (dtype == B200_BF16) {
    B200_CHECK_ARG(RowRegs<bf16>::supported(D), "add_ln_fwd: D=%d unsupported for bf16 (need D%%8==0, D<=2048)", D);
    add_ln_fwd_kernel<bf16><<<blocks, LN_WARPS * 32, 0, stream>>>((const bf16*)x, (const bf16*)res, gamma, beta, tile_group,
                                                        eps, (bf16*)y, mean, rstd, R, D);
  } else {
    B200_CHECK_ARG(RowRegs<float>::supported(D), "add_ln_fwd: D=%d unsupported for fp32 (need D%%4==0, D<=1024)", D);
    add_ln_fwd_kernel<float><<<blocks, LN_WARPS * 32, 0, stream>>>((const float*)x, (const float*)res, gamma, beta,
                                                         tile_group, eps, (float*)y, mean, rstd, R, D);
  }
  B200_LAUNCH_CHECK("add_ln_fwd_kernel");
  count_launch();
  return 0;
}

size_t b200_add_ln_bwd_ws(int R, int D) {
  const size_t blocks = (size_t)(R + LN_ROWS_PER_BLOCK - 1) / LN_ROWS_PER_BLOCK;
  return blocks * 2 * (size_t)D * sizeof(float);
}

int b200_add_ln_bwd(const void* dy, const void* x, const void* res, const float* mean, const float* rstd,
                    const float* gamma, const int32_t* tile_group, int G, void* dsum, float* dgamma, float* dbeta,
                    int R, int D, int dtype, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(R > 0 && D > 0 && G > 0, "add_ln_bwd: bad shape R=%d D=%d G=%d", R, D, G);
  B200_CHECK_ARG(workspace_bytes >= b200_add_ln_bwd_ws(R, D), "add_ln_bwd: workspace too small");
  const int blocks = (R + LN_ROWS_PER_BLOCK - 1) / LN_ROWS_PER_BLOCK;
  float* part = (float*)workspace;
  const size_t smem = (size_t)LN_WARPS * 2 * D * sizeof(float);
  if (tile_group != nullptr)  // unused tiles leave their partial slots untouched
    B200_CUDA(cudaMemsetAsync(part, 0, b200_add_ln_bwd_ws(R, D), stream));
  if (dtype == B200_BF16) {
    B200_CHECK_ARG(RowRegs<bf16>::supported(D), "add_ln_bwd: D=%d unsupported for bf16", D);
    if (smem > 48 * 1024)
      B200_CUDA(cudaFuncSetAttribute(add_ln_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    add_ln_bwd_kernel<bf16><<<blocks, LN_WARPS * 32, smem, stream>>>((const bf16*)dy, (const bf16*)x, (const bf16*)res, mean,
                                                           rstd, gamma, tile_group, (bf16*)dsum, part, R, D);
  } else {
    B200_CHECK_ARG(RowRegs<float>::supported(D), "add_ln_bwd: D=%d unsupported for fp32", D);
    if (smem > 48 * 1024)
      B200_CUDA(cudaFuncSetAttribute(add_ln_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    add_ln_bwd_kernel<float><<<blocks, LN_WARPS * 32, smem, stream>>>((const float*)dy, (const float*)x, (const float*)res,
                                                            mean, rstd, gamma, tile_group, (float*)dsum, part, R, D);
  }
  B200_LAUNCH_CHECK("add_ln_bwd_kernel");
  count_launch();
  return launch_ln_param_reduce(part, blocks, LN_ROWS_PER_BLOCK, D, tile_group, G, dgamma, dbeta, stream);
}

}  // extern "C"
