// bf16 GEMM on the 5th-gen tensor cores: TMA -> 128B-swizzled shared memory -> tcgen05.mma -> TMEM ->
// tcgen05.ld epilogue.  One kernel serves the dense Linear layers of the fusion blocks and the grouped
// expert FFN (forward, dgrad, wgrad) through operand major-ness flags and a tile->expert map.
//
// CTA = 128 threads: warp 0 lane 0 is the TMA producer, warp 1 lane 0 issues the UMMAs (and warp 1 owns
// the TMEM allocation); afterwards all four warps drain the 128 x BN fp32 accumulator (warp w owns TMEM
// lanes 32w..32w+31 = output rows) through the fused epilogue.  Tiles are not persistent: small problems
// keep >= 2 CTAs per SM resident (BN=64) so one CTA's epilogue overlaps another's main loop.
#include <cuda.h>

#include "gemm_common.cuh"
#include "tc_ptx.cuh"

namespace b200 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int CHUNK_BYTES = 64 * BK * 2;  // one 64(MN) x 64(K) MN-major TMA box

template <int BN> constexpr int stage_bytes() { return A_BYTES + BN * BK * 2; }
template <int BN> constexpr int smem_bytes() { return STAGES * stage_bytes<BN>() + 1024 + 256; }

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  constexpr int STAGE = stage_bytes<BN>();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  const uint32_t bar0 = base + STAGES * STAGE;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y;

  // ---- which problem does this CTA work on ------------------------------------------------------
  int group = 0;
  int a_mn0 = m_tile * BM, a_k0 = 0, b_mn0 = n_tile * BN, b_k0 = 0;
  int k_begin = 0, k_blocks = (p.K + BK - 1) / BK;
  if (p.mode == GEMM_GROUP_ROWS) {
    group = p.tile_group[m_tile];
    if (group < 0) return;
    if (B_MN) b_k0 = group * p.b_group_rows; else b_mn0 += group * p.b_group_rows;
  } else if (p.mode == GEMM_GROUP_WGRAD) {
    group = blockIdx.z;
    const int r0 = p.group_off[group], r1 = p.group_off[group + 1];
    a_k0 = b_k0 = r0;
    k_blocks = (r1 - r0 + BK - 1) / BK;
  } else if (p.k_splits > 1) {
    const int per = (k_blocks + p.k_splits - 1) / p.k_splits;
    k_begin = blockIdx.z * per;
    k_blocks = min(per, k_blocks - k_begin);
    if (k_blocks <= 0) return;
  }
  const bool have_acc = k_blocks > 0;

  // ---- one-time setup ---------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tma_a);
    ptx::prefetch_tensormap(&tma_b);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), BN);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- main loop: producer / MMA issuer ---------------------------------------------------------
  if (warp == 0 && lane == 0) {
    for (int kb = 0; kb < k_blocks; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      ptx::mbar_wait(empty_bar(s), ph ^ 1u);
      ptx::mbar_arrive_expect_tx(full_bar(s), STAGE);
      const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
      const int kc = (k_begin + kb) * BK;
      if (!A_MN) {
        ptx::tma_load_2d(sa, &tma_a, full_bar(s), a_k0 + kc, a_mn0);
      } else {
#pragma unroll
        for (int c = 0; c < BM / 64; ++c)
          ptx::tma_load_2d(sa + c * CHUNK_BYTES, &tma_a, full_bar(s), a_mn0 + 64 * c, a_k0 + kc);
      }
      if (!B_MN) {
        ptx::tma_load_2d(sb, &tma_b, full_bar(s), b_k0 + kc, b_mn0);
      } else {
#pragma unroll
        for (int c = 0; c < BN / 64; ++c)
          ptx::tma_load_2d(sb + c * CHUNK_BYTES, &tma_b, full_bar(s), b_mn0 + 64 * c, b_k0 + kc);
      }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, A_MN, B_MN);
    for (int kb = 0; kb < k_blocks; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      ptx::mbar_wait(full_bar(s), ph);
      ptx::tc_fence_after();
      const uint32_t sa = base + s * STAGE, sb = sa + A_BYTES;
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k) {
        // K-major: step 16 elements (32 B) inside the 128-byte swizzle row.
        // MN-major: step 16 k-rows (2048 B); LBO = distance between 64-wide MN chunks.
        const uint64_t ad = A_MN ? ptx::umma_smem_desc(sa + k * (UMMA_K * 128), CHUNK_BYTES, 1024)
                                 : ptx::umma_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
        const uint64_t bd = B_MN ? ptx::umma_smem_desc(sb + k * (UMMA_K * 128), CHUNK_BYTES, 1024)
                                 : ptx::umma_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
        ptx::umma_bf16(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
      }
      ptx::umma_commit(empty_bar(s));  // frees the smem slot once these MMAs retire
    }
    if (have_acc) ptx::umma_commit(tmem_full_bar);
  }
  __syncwarp();

  // ---- epilogue: TMEM -> registers -> fused op -> global ----------------------------------------
  if (have_acc) {
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
  }
  const long long row = (long long)m_tile * BM + warp * 32 + lane;
  const bool row_ok = row < p.M;
#pragma unroll 1
  for (int c = 0; c < BN / 32; ++c) {
    float acc[32];
    if (have_acc) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    }
    epilogue_store<bf16, 32>(p, group, row, n_tile * BN + c * 32, acc, row_ok);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, BN);
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows of pitch `pitch` elements; the box is
// 64 (inner, = 128 B swizzle span) x box_outer.
int make_map(CUtensorMap* m, const void* ptr, long long inner, long long outer, long long pitch, int box_outer) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return B200_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  cuuint32_t es[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%lld outer=%lld pitch=%lld box_outer=%d", (int)r,
              ptr, inner, outer, pitch, box_outer);
    return B200_ERR_CUDA;
  }
  return 0;
}

template <int BN, bool A_MN, bool B_MN>
int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, dim3 grid,
                   cudaStream_t stream) {
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BN>()));
    configured = true;
  }
  kern<<<grid, 128, smem_bytes<BN>(), stream>>>(ta, tb, args);
  B200_LAUNCH_CHECK("gemm_tc_kernel");
  count_launch();
  return 0;
}

template <int BN>
int launch_bn(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args,
              dim3 grid, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_variant<BN, false, false>(ta, tb, args, grid, stream);
  if (!a_mn && b_mn) return launch_variant<BN, false, true>(ta, tb, args, grid, stream);
  if (a_mn && b_mn) return launch_variant<BN, true, true>(ta, tb, args, grid, stream);
  return launch_variant<BN, true, false>(ta, tb, args, grid, stream);
}

}  // namespace

int launch_gemm_tc(const void* A, int lda, int a_layout, long long a_mn_extent, long long a_k_extent,
                   const void* B, int ldb, int b_layout, long long b_mn_extent, long long b_k_extent,
                   GemmArgs args, int grid_m_tiles, int groups, cudaStream_t stream) {
  const bool a_mn = a_layout == B200_LAYOUT_MN, b_mn = b_layout == B200_LAYOUT_MN;
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                 "gemm: operand base pointers must be 16-byte aligned");
  B200_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm: bf16 operand pitches must be multiples of 8 (lda=%d ldb=%d)",
                 lda, ldb);

  // tile-N selection: largest tile that still gives every SM a CTA
  const int sms = num_sms();
  const long long m_tiles = grid_m_tiles;
  int splits = 1;
  auto ctas = [&](int bn) { return m_tiles * ((args.N + bn - 1) / bn) * (args.mode == GEMM_GROUP_WGRAD ? groups : 1); };
  int bn = 64;
  if (ctas(256) >= sms) bn = 256;
  else if (ctas(128) >= sms) bn = 128;
  if (args.mode == GEMM_DENSE && args.epi == B200_EPI_ACCUM) {
    const int kblocks = (args.K + BK - 1) / BK;
    while (ctas(bn) * splits < sms && kblocks / (splits * 2) >= 4 && splits < 16) splits *= 2;
  }
  args.k_splits = splits;

  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = make_map(&ta, A, a_k_extent, a_mn_extent, lda, BM);
  else rc = make_map(&ta, A, a_mn_extent, a_k_extent, lda, BK);
  if (rc) return rc;
  if (!b_mn) rc = make_map(&tb, B, b_k_extent, b_mn_extent, ldb, bn);
  else rc = make_map(&tb, B, b_mn_extent, b_k_extent, ldb, BK);
  if (rc) return rc;

  dim3 grid((args.N + bn - 1) / bn, (unsigned)m_tiles,
            args.mode == GEMM_GROUP_WGRAD ? groups : splits);
  if (bn == 256) return launch_bn<256>(a_mn, b_mn, ta, tb, args, grid, stream);
  if (bn == 128) return launch_bn<128>(a_mn, b_mn, ta, tb, args, grid, stream);
  return launch_bn<64>(a_mn, b_mn, ta, tb, args, grid, stream);
}

}  // namespace b200
