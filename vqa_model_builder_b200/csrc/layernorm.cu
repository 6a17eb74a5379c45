// Fused residual-add + LayerNorm, forward and backward.  HBM-bound: one warp per row, 16-byte accesses,
// statistics in fp32.  Optional per-expert affine parameters through the tile->expert map.
#include "rowops.cuh"

namespace b200 {

// layernorm_staged.cu: persistent shared-memory-staged bf16 kernels (return -1 when a shape is not covered)
bool ln_staged_enabled();
size_t ln_bwd_staged_ws(int R, int D);
int launch_add_ln_fwd_staged(const bf16* x, const bf16* res, const float* gamma, const float* beta,
                             const int* tile_group, float eps, bf16* y, float* mean, float* rstd, int R, int D,
                             const unsigned long long* dst, float dp, unsigned int dsite, int drop_target,
                             cudaStream_t stream);
int launch_add_ln_bwd_staged(const bf16* dy, const bf16* x, const bf16* res, const float* mean, const float* rstd,
                             const float* gamma, const int* tile_group, int G, bf16* dsum, float* dgamma, float* dbeta,
                             float* d_colsum, int R, int D, const unsigned long long* dst, float dp, unsigned int dsite,
                             int drop_target, bf16* d_dropped, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream);

namespace {

constexpr int LN_WARPS = 8;
constexpr int LN_ROWS_PER_BLOCK = 8;  // workspace bound: the backward never uses fewer than 8 rows per block

// multiply a register-resident row by its dropout keep-scales (element index = r*D + d)
template <typename T, int NV>
__device__ __forceinline__ void drop_row(RowRegs<T, NV>& row, const DropState& ds, int r, int D, int lane) {
  constexpr int VT = Vec16<T>::N;
  const int nv = D / VT;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int vi = lane + 32 * j;
    if (vi < nv) {
      const unsigned long long base = (unsigned long long)r * D + (unsigned long long)vi * VT;
      if (VT == 8) {      // bf16 packets: 8 elements = one Philox call (D % 8 == 0, so base % 8 == 0)
        float sc[8];
        drop_scales8(ds, base >> 3, sc);
#pragma unroll
        for (int q = 0; q < 8; ++q) row.v[j][q % VT] *= sc[q];
      } else {
#pragma unroll
        for (int u = 0; u < VT; u += 4) {
          float sc[4];
          drop_scales4(ds, (base + u) >> 2, sc);
#pragma unroll
          for (int q = 0; q < 4; ++q) row.v[j][u + q] *= sc[q];
        }
      }
    }
  }
}

// row = drop?(x) + drop?(res)
template <typename T, int NV>
__device__ __forceinline__ void load_sum_row(RowRegs<T, NV>& row, const T* x, const T* res, const DropState& ds,
                                             int drop_target, int r, int D, int lane) {
  row.load(x + (long long)r * D, D, lane);
  if (ds.on && drop_target == 1) drop_row<T, NV>(row, ds, r, D, lane);
  if (res != nullptr) {
    if (ds.on && drop_target == 2) {
      RowRegs<T, NV> rr;
      rr.load(res + (long long)r * D, D, lane);
      drop_row<T, NV>(rr, ds, r, D, lane);
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int u = 0; u < Vec16<T>::N; ++u) row.v[j][u] += rr.v[j][u];
    } else {
      row.axpy(res + (long long)r * D, 1.f, D, lane);
    }
  }
}

template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
                  const float* __restrict__ beta, const int* __restrict__ tile_group, float eps,
                  T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int R, int D,
                  const unsigned long long* drop_state, float drop_p, unsigned int drop_site, int drop_target) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = D / VT;
  const DropState ds = drop_load(drop_state, drop_p, drop_site);
  const int r = blockIdx.x * LN_WARPS + warp;
  if (r >= R) return;
  int g = 0;
  if (tile_group != nullptr) {
    g = tile_group[r / B200_GROUP_TILE];
    if (g < 0) return;
  }
  RowRegs<T, NV> row;
  load_sum_row<T, NV>(row, x, res, ds, drop_target, r, D, lane);
  const float mean = row.sum(D, lane) / D;
  const float var = row.sumsq_centered(mean, D, lane) / D;
  const float rstd = rsqrtf(var + eps);
  const float* gm = gamma + (long long)g * D;
  const float* bt = beta + (long long)g * D;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
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

// dsum = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma.
// Each warp walks `rpw` consecutive rows keeping its dgamma/dbeta partials in registers; the block (8 warps,
// 8*rpw rows, always inside one 128-row expert tile) reduces them through shared memory into part[block][2][D];
// a second kernel reduces the partials per group.
template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
add_ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                  const float* __restrict__ gamma, const int* __restrict__ tile_group, T* __restrict__ dsum,
                  float* __restrict__ part, int R, int D, int rpw, const unsigned long long* drop_state, float drop_p,
                  unsigned int drop_site, int drop_target, T* __restrict__ d_dropped, int nz) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  extern __shared__ float red[];  // [LN_WARPS][nz][D]; nz = 3 adds the column sums of the branch gradient
  const DropState ds = drop_load(drop_state, drop_p, drop_site);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = D / VT;
  // per-warp dgamma / dbeta accumulators live in shared memory (each lane owns its own 16-byte slots, so no
  // synchronisation is needed): keeping them in registers cost ~50 registers and left one block per SM
  float* acc_g = red + (warp * nz + 0) * D;
  float* acc_b = red + (warp * nz + 1) * D;
  float* acc_c = red + (warp * nz + 2) * D;      // only touched when nz == 3
  for (int d = lane; d < D; d += 32) { acc_g[d] = 0.f; acc_b[d] = 0.f; }
  if (nz == 3)
    for (int d = lane; d < D; d += 32) acc_c[d] = 0.f;
  __syncwarp();

  const int row0 = blockIdx.x * (LN_WARPS * rpw);
  int g = 0;
  bool live = true;
  if (tile_group != nullptr) {
    g = tile_group[row0 / B200_GROUP_TILE];
    live = g >= 0;
    if (!live) g = 0;
  }
  const float* gm = gamma + (long long)g * D;
  for (int i = 0; i < rpw && live; ++i) {
    const int r = row0 + i * LN_WARPS + warp;   // interleaved so the 8 warps stream adjacent rows
    if (r >= R) break;
    RowRegs<T, NV> xr, gr;
    load_sum_row<T, NV>(xr, x, res, ds, drop_target, r, D, lane);
    gr.load(dy + (long long)r * D, D, lane);
    const float mean = mean_in[r], rstd = rstd_in[r];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nv) {
        float gv[VT];
        load_param<VT>(gm, vi, gv);
        float* ag = acc_g + vi * VT;
        float* ab = acc_b + vi * VT;
#pragma unroll
        for (int u = 0; u < VT; u += 4) {
          float4 tg = *reinterpret_cast<float4*>(ag + u), tb = *reinterpret_cast<float4*>(ab + u);
          float* pg = reinterpret_cast<float*>(&tg);
          float* pb = reinterpret_cast<float*>(&tb);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float xhat = (xr.v[j][u + q] - mean) * rstd;
            const float d = gr.v[j][u + q];
            pg[q] = fmaf(d, xhat, pg[q]);
            pb[q] += d;
            const float gg = d * gv[u + q];
            xr.v[j][u + q] = xhat;
            gr.v[j][u + q] = gg;
            s1 += gg;
            s2 = fmaf(gg, xhat, s2);
          }
          *reinterpret_cast<float4*>(ag + u) = tg;
          *reinterpret_cast<float4*>(ab + u) = tb;
        }
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int u = 0; u < VT; ++u) gr.v[j][u] = rstd * (gr.v[j][u] - s1 - xr.v[j][u] * s2);
    gr.store(dsum + (long long)r * D, D, lane);
    if (ds.on && d_dropped != nullptr) {   // gradient of the operand that went through dropout
      drop_row<T, NV>(gr, ds, r, D, lane);
      gr.store(d_dropped + (long long)r * D, D, lane);
    }
    if (nz == 3) {   // column sums of the branch gradient = bias gradient of the Linear that produced the branch
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int vi = lane + 32 * j;
        if (vi < nv) {
          float* ac = acc_c + vi * VT;
#pragma unroll
          for (int u = 0; u < VT; u += 4) {
            float4 t = *reinterpret_cast<float4*>(ac + u);
            t.x += gr.v[j][u]; t.y += gr.v[j][u + 1]; t.z += gr.v[j][u + 2]; t.w += gr.v[j][u + 3];
            *reinterpret_cast<float4*>(ac + u) = t;
          }
        }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < nz * D; c += blockDim.x) {
    const int which = c / D, d = c % D;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) s += red[(w * nz + which) * D + d];
    part[((long long)blockIdx.x * nz + which) * D + d] = s;
  }
}

// rows per warp for the backward kernels: enough blocks to fill the machine, few enough partials to keep the
// second-stage reduction small.  Power of two <= 16 so a block (8*rpw rows) never straddles a 128-row tile.
inline int ln_bwd_rpw(int R) {
  int rpw = 1;
  while (rpw < 16 && (R + LN_WARPS * rpw * 2 - 1) / (LN_WARPS * rpw * 2) >= 2 * num_sms()) rpw *= 2;
  return rpw;
}

// out_z[g][c] = sum over partial blocks b of group g of part[b][z][c]   (z = blockIdx.z selects dgamma / dbeta).
// 32 columns x 8 partial-lanes per block: loads are independent across threads instead of one serial chain.
// Expert segments are contiguous tile ranges, so the block first finds [first,last] tile of its group (parallel scan
// of the tile map) and then walks only that block range — no per-iteration indirection through tile_group.
__global__ void __launch_bounds__(1024)
partial_reduce_kernel(const float* __restrict__ part, int blocks, int rows_per_block, int width, int nz,
                      const int* __restrict__ tile_group, int tiles, float* __restrict__ out0,
                      float* __restrict__ out1, float* __restrict__ out2) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32][33];
  __shared__ int s_lo, s_hi;
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;     // 32 columns x 32 partial lanes
  const int c = blockIdx.x * 32 + x;
  const int g = blockIdx.y, z = blockIdx.z;
  int b_lo = 0, b_hi = blocks;
  if (tile_group != nullptr) {
    if (threadIdx.x == 0) { s_lo = tiles; s_hi = -1; }
    __syncthreads();
    for (int t = threadIdx.x; t < tiles; t += blockDim.x)
      if (tile_group[t] == g) { atomicMin(&s_lo, t); atomicMax(&s_hi, t); }
    __syncthreads();
    const int bpt = B200_GROUP_TILE / rows_per_block;   // partial blocks per 128-row tile
    b_lo = s_lo * bpt;
    b_hi = min(blocks, (s_hi + 1) * bpt);
  }
  // four independent loads in flight per thread; the order of the additions is fixed (deterministic result)
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < width) {
    const float* p = part + (long long)z * width + c;
    const long long step = (long long)nz * width;
    int b = b_lo + y;
    for (; b + 96 < b_hi; b += 128) {
      s0 += p[(long long)b * step];
      s1 += p[(long long)(b + 32) * step];
      s2 += p[(long long)(b + 64) * step];
      s3 += p[(long long)(b + 96) * step];
    }
    for (; b < b_hi; b += 32) s0 += p[(long long)b * step];
  }
  red[y][x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (y == 0 && c < width) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) t += red[j][x];
    (z == 0 ? out0 : (z == 1 ? out1 : out2))[(long long)g * width + c] = t;
  }
}

}  // namespace

// shared with dispatch.cu (combine backward uses the same partial layout)
int launch_ln_param_reduce(const float* part, int blocks, int rows_per_block, int D, const int* tile_group, int G,
                           float* dgamma, float* dbeta, cudaStream_t stream, float* dcol) {
  const int nz = dcol != nullptr ? 3 : 2;
  dim3 grid((D + 31) / 32, G, nz);
  const int tiles = (int)(((long long)blocks * rows_per_block + B200_GROUP_TILE - 1) / B200_GROUP_TILE);
  launch_kernel(partial_reduce_kernel, dim3(grid), dim3(1024), 0, stream, part, blocks, rows_per_block, D, nz, tile_group, tiles, dgamma, dbeta, dcol);
  B200_LAUNCH_CHECK("partial_reduce_kernel");
  count_launch();
  return 0;
}

// single-array variant (column sums)
int launch_partial_reduce(const float* part, int blocks, int rows_per_block, int width, const int* tile_group, int G,
                          float* out, cudaStream_t stream) {
  dim3 grid((width + 31) / 32, G, 1);
  const int tiles = (int)(((long long)blocks * rows_per_block + B200_GROUP_TILE - 1) / B200_GROUP_TILE);
  launch_kernel(partial_reduce_kernel, dim3(grid), dim3(1024), 0, stream, part, blocks, rows_per_block, width, 1, tile_group, tiles, out, out, out);
  B200_LAUNCH_CHECK("partial_reduce_kernel");
  count_launch();
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_add_ln_fwd(const void* x, const void* res, const float* gamma, const float* beta,
                    const int32_t* tile_group, float eps, void* y, float* mean, float* rstd, int R, int D, int dtype,
                    const b200_dropout_t* drop, int drop_target, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(R > 0 && D > 0, "add_ln_fwd: bad shape R=%d D=%d", R, D);
  const bool don = drop != nullptr && drop->p > 0.f && drop_target != 0;
  const unsigned long long* dst = don ? drop->rng_state : nullptr;
  const float dp = don ? drop->p : 0.f;
  const unsigned int dsite = don ? drop->site : 0u;
  const int blocks = (R + LN_WARPS - 1) / LN_WARPS;
  if (dtype == B200_BF16 && ln_staged_enabled()) {
    const int rc = launch_add_ln_fwd_staged((const bf16*)x, (const bf16*)res, gamma, beta, tile_group, eps, (bf16*)y, mean,
                                            rstd, R, D, dst, dp, dsite, drop_target, stream);
    if (rc >= 0) return rc;
  }
  if (dtype == B200_BF16) {
    B200_CHECK_ARG(RowRegs<bf16>::supported(D), "add_ln_fwd: D=%d unsupported for bf16 (need D%%8==0, D<=2048)", D);
    B200_NV_SWITCH(row_nv<bf16>(D), launch_kernel(add_ln_fwd_kernel<bf16, NV>, dim3(blocks), dim3(LN_WARPS * 32), 0, stream, 
        (const bf16*)x, (const bf16*)res, gamma, beta, tile_group, eps, (bf16*)y, mean, rstd, R, D, dst, dp, dsite, drop_target));
  } else {
    B200_CHECK_ARG(RowRegs<float>::supported(D), "add_ln_fwd: D=%d unsupported for fp32 (need D%%4==0, D<=1024)", D);
    B200_NV_SWITCH(row_nv<float>(D), launch_kernel(add_ln_fwd_kernel<float, NV>, dim3(blocks), dim3(LN_WARPS * 32), 0, stream, 
        (const float*)x, (const float*)res, gamma, beta, tile_group, eps, (float*)y, mean, rstd, R, D, dst, dp, dsite, drop_target));
  }
  B200_LAUNCH_CHECK("add_ln_fwd_kernel");
  count_launch();
  return 0;
}

size_t b200_add_ln_bwd_ws(int R, int D) {
  const size_t blocks = (size_t)(R + LN_ROWS_PER_BLOCK - 1) / LN_ROWS_PER_BLOCK;
  const size_t a = blocks * 3 * (size_t)D * sizeof(float), b = ln_bwd_staged_ws(R, D);
  return a > b ? a : b;
}

int b200_add_ln_bwd(const void* dy, const void* x, const void* res, const float* mean, const float* rstd,
                    const float* gamma, const int32_t* tile_group, int G, void* dsum, float* dgamma, float* dbeta,
                    float* d_colsum, int R, int D, int dtype, const b200_dropout_t* drop, int drop_target,
                    void* d_dropped, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(R > 0 && D > 0 && G > 0, "add_ln_bwd: bad shape R=%d D=%d G=%d", R, D, G);
  const bool don = drop != nullptr && drop->p > 0.f && drop_target != 0;
  const unsigned long long* dst = don ? drop->rng_state : nullptr;
  const float dp = don ? drop->p : 0.f;
  const unsigned int dsite = don ? drop->site : 0u;
  B200_CHECK_ARG(workspace_bytes >= b200_add_ln_bwd_ws(R, D), "add_ln_bwd: workspace too small");
  if (dtype == B200_BF16 && ln_staged_enabled()) {
    const int rc = launch_add_ln_bwd_staged((const bf16*)dy, (const bf16*)x, (const bf16*)res, mean, rstd, gamma, tile_group,
                                            G, (bf16*)dsum, dgamma, dbeta, d_colsum, R, D, dst, dp, dsite, drop_target,
                                            (bf16*)d_dropped, workspace, workspace_bytes, stream);
    if (rc >= 0) return rc;
  }
  const int rpw = ln_bwd_rpw(R);
  const int rows_per_block = LN_WARPS * rpw;
  const int blocks = (R + rows_per_block - 1) / rows_per_block;
  float* part = (float*)workspace;
  const int nz = d_colsum != nullptr ? 3 : 2;
  const size_t smem = (size_t)LN_WARPS * nz * D * sizeof(float);
  B200_CHECK_ARG(smem <= 160 * 1024, "add_ln_bwd: D=%d too large", D);
  if (dtype == B200_BF16) {
    B200_CHECK_ARG(RowRegs<bf16>::supported(D), "add_ln_bwd: D=%d unsupported for bf16", D);
    B200_NV_SWITCH(row_nv<bf16>(D), {
      if (smem > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(add_ln_bwd_kernel<bf16, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      launch_kernel(add_ln_bwd_kernel<bf16, NV>, dim3(blocks), dim3(LN_WARPS * 32), smem, stream, 
          (const bf16*)dy, (const bf16*)x, (const bf16*)res, mean, rstd, gamma, tile_group, (bf16*)dsum, part, R, D, rpw,
          dst, dp, dsite, drop_target, (bf16*)d_dropped, nz);
    });
  } else {
    B200_CHECK_ARG(RowRegs<float>::supported(D), "add_ln_bwd: D=%d unsupported for fp32", D);
    B200_NV_SWITCH(row_nv<float>(D), {
      if (smem > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(add_ln_bwd_kernel<float, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      launch_kernel(add_ln_bwd_kernel<float, NV>, dim3(blocks), dim3(LN_WARPS * 32), smem, stream, 
          (const float*)dy, (const float*)x, (const float*)res, mean, rstd, gamma, tile_group, (float*)dsum, part, R, D, rpw,
          dst, dp, dsite, drop_target, (float*)d_dropped, nz);
    });
  }
  B200_LAUNCH_CHECK("add_ln_bwd_kernel");
  count_launch();
  return launch_ln_param_reduce(part, blocks, rows_per_block, D, tile_group, G, dgamma, dbeta, stream, d_colsum);
}

}  // extern "C"
