// fp32 GEMM on the CUDA cores (validation mode: the 1e-4 fp32 parity target needs full-precision
// products, which the bf16/tf32 tensor pipes cannot give).  Same argument block, modes and epilogues as
// the tcgen05 kernel; operands are addressed through element strides so one kernel covers forward,
// dgrad and wgrad layouts.  64x64x16 tiles, 256 threads, 4x4 register micro-tiles.
#include "gemm_common.cuh"

namespace b200 {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmArgs p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  int group = 0;
  const float* A = reinterpret_cast<const float*>(p.A);
  const float* B = reinterpret_cast<const float*>(p.B);
  int K = p.K;
  if (p.mode == GEMM_GROUP_ROWS) {
    group = p.tile_group[m0 / B200_GROUP_TILE];
    if (group < 0) return;
    B += (long long)group * p.b_group_elems;
  } else if (p.mode == GEMM_GROUP_WGRAD) {
    group = blockIdx.z;
    const int r0 = p.group_off[group], r1 = p.group_off[group + 1];
    A += (long long)r0 * p.sa_k;
    B += (long long)r0 * p.sb_k;
    K = r1 - r0;
  }

  const int tx = tid % 16, ty = tid / 16;  // micro-tile: rows ty*4.., cols tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kc = (p.sa_k == 1), b_kc = (p.sb_k == 1);
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int it = 0; it < (TM * TK) / 256; ++it) {
      const int e = tid + it * 256;
      const int mm = a_kc ? e / TK : e % TM, kk = a_kc ? e % TK : e / TM;
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < p.M && gk < K) ? __ldg(A + (long long)gm * p.sa_m + (long long)gk * p.sa_k) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < (TN * TK) / 256; ++it) {
      const int e = tid + it * 256;
      const int nn = b_kc ? e / TK : e % TN, kk = b_kc ? e % TK : e / TN;
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < p.N && gk < K) ? __ldg(B + (long long)gn * p.sb_n + (long long)gk * p.sb_k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long row = m0 + ty * 4 + i;
    epilogue_store<float, 4>(p, group, row, n0 + tx * 4, acc[i], row < p.M);
  }
}

}  // namespace

int launch_gemm_simt(GemmArgs args, int grid_m_tiles, int groups, cudaStream_t stream) {
  // grid_m_tiles is given in 128-row units (shared with the tcgen05 path)
  dim3 grid((args.N + TN - 1) / TN, grid_m_tiles * (B200_GROUP_TILE / TM),
            args.mode == GEMM_GROUP_WGRAD ? groups : 1);
  args.k_splits = 1;
  launch_kernel(gemm_simt_kernel, dim3(grid), dim3(256), 0, stream, args);
  B200_LAUNCH_CHECK("gemm_simt_kernel");
  count_launch();
  return 0;
}

}  // namespace b200
