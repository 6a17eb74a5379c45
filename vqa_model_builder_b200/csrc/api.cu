// C-ABI surface: lifecycle, error reporting, casts, column sums and the GEMM entry points.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "gemm_common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static int g_num_sms = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return B200_ERR_CUDA;
}
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      g_num_sms = 148;
  }
  return g_num_sms;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VQA_PDL");
    v = (e != nullptr && strcmp(e, "0") == 0) ? 0 : 1;
  }
  return v == 1;
}

// ---- cast ---------------------------------------------------------------------------------------
template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, long long n) {
  pdl_trigger();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float v[4];
      if constexpr (sizeof(S) == 4) {
        float4 t = *reinterpret_cast<const float4*>(src + i);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        uint2 t = *reinterpret_cast<const uint2*>(src + i);
        float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
      }
      if constexpr (sizeof(D) == 4) {
        *reinterpret_cast<float4*>(dst + i) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        uint2 t;
        t.x = pack_bf16x2(v[0], v[1]);
        t.y = pack_bf16x2(v[2], v[3]);
        *reinterpret_cast<uint2*>(dst + i) = t;
      }
    } else {
      for (long long j = i; j < n; ++j) dst[j] = from_f32<D>(to_f32<S>(src[j]));
    }
  }
}

// ---- column sums (bias gradients): one launch, deterministic ------------------------------------------------
// Every block writes the partial sums of its 64 rows; the block that finishes LAST within a column slab (ticket
// counter, __threadfence reduction) folds the slab's partials in block order, so the result does not depend on
// scheduling.  Tickets live in a device-global ring (zero at load, reset by the folding block): launches in flight
// at the same time on different streams get different slots.
constexpr int CS_ROWS = 128;  // rows per partial block of the grouped path; divides B200_GROUP_TILE
constexpr int TICKETS = 8192;
__device__ unsigned int g_tickets[TICKETS];

unsigned int ticket_slots(int n) {
  static std::atomic<unsigned int> next{0};
  unsigned int base = next.fetch_add((unsigned int)n) % TICKETS;
  if (base + (unsigned int)n > TICKETS) base = 0;
  return base;
}

// true in exactly one block per ticket: the last of `total` blocks to arrive (its reads see all partials)
__device__ __forceinline__ bool last_arrival(unsigned int* ticket, unsigned int total) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == total - 1);
    if (s_last) *ticket = 0u;       // ready for the next launch that is handed this slot
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

// fold part[b][c] over the row blocks b: out[g][c] for every group g (groups are contiguous tile ranges of the
// padded expert layout; tile_group == nullptr: one group)
// sum of part[b][c] over b = b0, b0 + step, ... < blocks (fixed order: deterministic)
__device__ __forceinline__ float fold_strided(const float* __restrict__ part, int blocks, int N, int c, int b0, int step) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int b = b0;
  for (; b + 3 * step < blocks; b += 4 * step) {
    s0 += __ldcg(part + (long long)b * N + c);
    s1 += __ldcg(part + (long long)(b + step) * N + c);
    s2 += __ldcg(part + (long long)(b + 2 * step) * N + c);
    s3 += __ldcg(part + (long long)(b + 3 * step) * N + c);
  }
  for (; b < blocks; b += step) s0 += __ldcg(part + (long long)b * N + c);
  return (s0 + s1) + (s2 + s3);
}

__device__ __forceinline__ void fold_partials(const float* __restrict__ part, int blocks, int N, int c,
                                              const int* __restrict__ tile_group, int G, float* __restrict__ out) {
  if (tile_group == nullptr) {
    out[c] = fold_strided(part, blocks, N, c, 0, 1);
    return;
  }
  for (int g = 0; g < G; ++g) out[(long long)g * N + c] = 0.f;
  constexpr int BPT = B200_GROUP_TILE / CS_ROWS;
  int cur = -1;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) {
    const int g = tile_group[b / BPT];
    if (g != cur) {
      if (cur >= 0) out[(long long)cur * N + c] = s;
      cur = g;
      s = 0.f;
    }
    if (g >= 0) s += __ldcg(part + (long long)b * N + c);
  }
  if (cur >= 0) out[(long long)cur * N + c] = s;
}

// 8 warps x 16-byte vectors: block (slab, rb) sums rows [rb*rpb, rb*rpb+rpb) of a 32*VT-column slab; four row loads per
// warp are in flight at a time.  With `drop` the kernel first applies the dropout keep-scales of the site (element
// index = row * N + col), stores the scaled rows to `gout` and sums THOSE: the backward of "dropout(x W^T + b) +
// residual" needs both dropout(dy) (operand of the dgrad / wgrad GEMMs) and its column sums (the bias gradient).
constexpr int CS_WARPS = 16;
template <typename T>
__global__ void __launch_bounds__(CS_WARPS * 32)
colsum_stage1(const T* __restrict__ x, int R, int N, int rpb, float* __restrict__ part,
              const int* __restrict__ tile_group, int G, float* __restrict__ out, unsigned int* __restrict__ tickets,
              T* __restrict__ gout, const unsigned long long* drop_state, float drop_p, unsigned int drop_site) {
  pdl_trigger();
  pdl_wait();
  constexpr int VT = Vec16<T>::N;
  __shared__ float red[CS_WARPS][32 * VT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = (blockIdx.x * 32 + lane) * VT;
  const int r0 = blockIdx.y * rpb, r1 = min(R, r0 + rpb);
  const DropState ds = drop_load(drop_state, drop_p, drop_site);
  float acc[VT];
#pragma unroll
  for (int u = 0; u < VT; ++u) acc[u] = 0.f;
  // rows of unused 128-row tiles (expert-parallel buffers are sized for the worst case) are never read
  const bool live = tile_group == nullptr || tile_group[r0 / B200_GROUP_TILE] >= 0;
  if (col < N && live) {
    constexpr int INF = 8;                    // row loads in flight per warp (16 warps x 32 lanes x 8 x 16 B = 64 KB / CTA)
    for (int rb = r0 + warp; rb < r1; rb += INF * CS_WARPS) {
      Vec16<T> v[INF];
#pragma unroll
      for (int q = 0; q < INF; ++q) {
        const int r = rb + CS_WARPS * q;
        if (r < r1) v[q].load(x + (long long)r * N + col);
      }
#pragma unroll
      for (int q = 0; q < INF; ++q) {
        const int r = rb + CS_WARPS * q;
        if (r < r1) {
          if (ds.on) {
            const unsigned long long base = (unsigned long long)r * N + col;
            if (VT == 8) {
              float sc[8];
              drop_scales8(ds, base >> 3, sc);
#pragma unroll
              for (int u = 0; u < 8; ++u) v[q].v[u % VT] *= sc[u];
            } else {
              float sc[4];
              drop_scales4(ds, base >> 2, sc);
#pragma unroll
              for (int u = 0; u < 4; ++u) v[q].v[u % VT] *= sc[u];
            }
            v[q].store(gout + (long long)r * N + col);
          }
#pragma unroll
          for (int u = 0; u < VT; ++u) acc[u] += v[q].v[u];
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < VT; ++u) red[warp][lane * VT + u] = acc[u];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VT; c += blockDim.x) {
    const int gc = blockIdx.x * 32 * VT + c;
    if (gc < N) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < CS_WARPS; ++w) s += red[w][c];
      part[(long long)blockIdx.y * N + gc] = s;
    }
  }
  if (tickets == nullptr) return;                     // grouped: a second kernel folds per expert in parallel
  if (!last_arrival(tickets + blockIdx.x, gridDim.y)) return;
  // fold the (at most 64) partials of this column slab: every thread sums a strided subset of the row blocks of one
  // column, the subsets are combined through shared memory in a fixed order
  constexpr int COLS = 32 * VT;
  constexpr int SUB = (CS_WARPS * 32) / COLS >= 1 ? (CS_WARPS * 32) / COLS : 1;     // threads per column
  float* fold = &red[0][0];
  __syncthreads();
  for (int i = threadIdx.x; i < COLS * SUB; i += blockDim.x) {
    const int c = i % COLS, sub = i / COLS;
    const int gc = blockIdx.x * COLS + c;
    fold[sub * COLS + c] = gc < N ? fold_strided(part, gridDim.y, N, gc, sub, SUB) : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < COLS; c += blockDim.x) {
    const int gc = blockIdx.x * COLS + c;
    if (gc < N) {
      float s = 0.f;
#pragma unroll
      for (int sub = 0; sub < SUB; ++sub) s += fold[sub * COLS + c];
      out[gc] = s;
    }
  }
}
// scalar fallback for widths that are not a multiple of the vector length
template <typename T>
__global__ void colsum_stage1_scalar(const T* __restrict__ x, int R, int N, int rpb, float* __restrict__ part,
                                     const int* __restrict__ tile_group, int G, float* __restrict__ out,
                                     unsigned int* __restrict__ tickets) {
  pdl_trigger();
  pdl_wait();
  const int col = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * rpb, r1 = min(R, r0 + rpb);
  if (col < N) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += to_f32<T>(x[(long long)r * N + col]);
    part[(long long)blockIdx.y * N + col] = s;
  }
  if (tickets == nullptr) return;
  if (!last_arrival(tickets + blockIdx.x, gridDim.y)) return;
  if (col < N) fold_partials(part, gridDim.y, N, col, tile_group, G, out);
}

int launch_partial_reduce(const float* part, int blocks, int rows_per_block, int width, const int* tile_group, int G,
                          float* out, cudaStream_t stream);

__global__ void dropout_mask_kernel(const unsigned long long* st, float p, unsigned int site, long long n, float* out) {
  pdl_trigger();
  pdl_wait();
  const DropState d = drop_load(st, p, site);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = d.on ? drop_scale1(d, (unsigned long long)i) : 1.f;
}

template <typename T>
__global__ void dropout_apply_kernel(const T* __restrict__ x, T* __restrict__ out, long long n,
                                     const unsigned long long* st, float p, unsigned int site) {
  pdl_trigger();
  pdl_wait();
  const DropState d = drop_load(st, p, site);
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    float sc[4];
    drop_scales4(d, (unsigned long long)i >> 2, sc);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (i + q < n) out[i + q] = from_f32<T>(to_f32<T>(x[i + q]) * sc[q]);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_dropout_apply(const void* x, void* out, long long n, int dtype, const b200_dropout_t* drop, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(drop != nullptr && n > 0, "dropout_apply: bad arguments");
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  if (dtype == B200_BF16)
    launch_kernel(dropout_apply_kernel<bf16>, dim3((int)blocks), dim3(256), 0, stream, (const bf16*)x, (bf16*)out, n, drop->rng_state, drop->p, drop->site);
  else
    launch_kernel(dropout_apply_kernel<float>, dim3((int)blocks), dim3(256), 0, stream, (const float*)x, (float*)out, n, drop->rng_state, drop->p, drop->site);
  B200_LAUNCH_CHECK("dropout_apply_kernel");
  count_launch();
  return 0;
}

int b200_init(int device) {
  B200_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B200_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("b200vqa requires an sm_100 (Blackwell B200) device; device %d is sm_%d%d", device, prop.major,
              prop.minor);
    return B200_ERR_UNSUPPORTED;
  }
  return 0;
}
const char* b200_last_error_string(void) { return g_err; }
int b200_abi_version(void) { return B200VQA_ABI_VERSION; }
long long b200_launch_count(void) { return g_launches.load(); }
void b200_reset_launch_count(void) { g_launches.store(0); }

int b200_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n <= 0) return 0;
  B200_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "cast: pointers must be 16-byte aligned");
  const int threads = 256;
  long long blocks = (n / 4 + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (src_dtype == B200_F32 && dst_dtype == B200_BF16)
    launch_kernel(cast_kernel<float, bf16>, dim3((int)blocks), dim3(threads), 0, stream, (const float*)src, (bf16*)dst, n);
  else if (src_dtype == B200_BF16 && dst_dtype == B200_F32)
    launch_kernel(cast_kernel<bf16, float>, dim3((int)blocks), dim3(threads), 0, stream, (const bf16*)src, (float*)dst, n);
  else if (src_dtype == B200_F32 && dst_dtype == B200_F32)
    launch_kernel(cast_kernel<float, float>, dim3((int)blocks), dim3(threads), 0, stream, (const float*)src, (float*)dst, n);
  else if (src_dtype == B200_BF16 && dst_dtype == B200_BF16)
    launch_kernel(cast_kernel<bf16, bf16>, dim3((int)blocks), dim3(threads), 0, stream, (const bf16*)src, (bf16*)dst, n);
  else {
    set_error("cast: bad dtypes %d -> %d", src_dtype, dst_dtype);
    return B200_ERR_INVALID;
  }
  B200_LAUNCH_CHECK("cast_kernel");
  count_launch();
  return 0;
}

int b200_dropout_mask(const b200_dropout_t* drop, long long n, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(drop != nullptr && n > 0, "dropout_mask: bad arguments");
  launch_kernel(dropout_mask_kernel, dim3(256), dim3(256), 0, stream, drop->rng_state, drop->p, drop->site, n, out);
  B200_LAUNCH_CHECK("dropout_mask_kernel");
  count_launch();
  return 0;
}

size_t b200_colsum_ws(int R, int N) {
  const size_t blocks = (size_t)(R + CS_ROWS - 1) / CS_ROWS;
  return blocks * (size_t)N * sizeof(float);
}
static int launch_colsum(const void* x, void* gout, const b200_dropout_t* drop, int dtype, int R, int N,
                         const int32_t* tile_group, int G, float* out, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream) {
  B200_CHECK_ARG(R > 0 && N > 0 && G > 0, "colsum: bad shape R=%d N=%d G=%d", R, N, G);
  B200_CHECK_ARG(workspace_bytes >= b200_colsum_ws(R, N), "colsum: workspace too small");
  const bool don = drop != nullptr && drop->p > 0.f && gout != nullptr;
  const int vt = dtype == B200_F32 ? 4 : 8;
  const bool vec = (N % vt == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)gout & 15) == 0);
  B200_CHECK_ARG(!don || vec, "dropout_colsum: needs 16-byte aligned rows (N %% %d == 0)", vt);
  // Ungrouped: ONE launch — at most 64 row blocks (rows per block grows with R), the last-arriving block of a column
  // slab folds the partials in block order (deterministic).  Grouped (per-expert sums): 64-row blocks, which never
  // straddle a 128-row expert tile, and a second kernel folds per expert in parallel.
  int rpb = CS_ROWS;
  const bool single = tile_group == nullptr;
  const int slabs = vec ? (N + 32 * vt - 1) / (32 * vt) : (N + 127) / 128;
  if (single) {
    // row blocks: at most 64 (the serial part of the fold), and slabs x row-blocks close to ONE full wave of equally
    // loaded CTAs (a 171-CTA grid on 148 SMs took two rounds: profiles/r02f)
    int rb = num_sms() / slabs;
    if (rb > 64) rb = 64;
    if (rb < 1) rb = 1;
    rpb = ((R + rb - 1) / rb + CS_WARPS - 1) / CS_WARPS * CS_WARPS;
    if (rpb < CS_ROWS) rpb = CS_ROWS;
  }
  const int blocks = (R + rpb - 1) / rpb;
  float* part = (float*)workspace;
  dim3 g1(slabs, blocks);
  unsigned int* tk = nullptr;
  if (single) {
    B200_CUDA(cudaGetSymbolAddress((void**)&tk, g_tickets));
    tk += ticket_slots((int)g1.x);
  }
  const unsigned long long* dst = don ? drop->rng_state : nullptr;
  const float dp = don ? drop->p : 0.f;
  const unsigned int dsite = don ? drop->site : 0u;
  if (vec) {
    if (dtype == B200_F32) launch_kernel(colsum_stage1<float>, dim3(g1), dim3(CS_WARPS * 32), 0, stream, (const float*)x, R, N, rpb, part, tile_group, G, out, tk, (float*)gout, dst, dp, dsite);
    else launch_kernel(colsum_stage1<bf16>, dim3(g1), dim3(CS_WARPS * 32), 0, stream, (const bf16*)x, R, N, rpb, part, tile_group, G, out, tk, (bf16*)gout, dst, dp, dsite);
  } else {
    if (dtype == B200_F32) launch_kernel(colsum_stage1_scalar<float>, dim3(g1), dim3(128), 0, stream, (const float*)x, R, N, rpb, part, tile_group, G, out, tk);
    else launch_kernel(colsum_stage1_scalar<bf16>, dim3(g1), dim3(128), 0, stream, (const bf16*)x, R, N, rpb, part, tile_group, G, out, tk);
  }
  B200_LAUNCH_CHECK("colsum_stage1");
  count_launch();
  if (single) return 0;
  return launch_partial_reduce(part, blocks, CS_ROWS, N, tile_group, G, out, stream);
}

int b200_colsum(const void* x, int dtype, int R, int N, const int32_t* tile_group, int G, float* out,
                void* workspace, size_t workspace_bytes, void* stream_) {
  return launch_colsum(x, nullptr, nullptr, dtype, R, N, tile_group, G, out, workspace, workspace_bytes,
                       (cudaStream_t)stream_);
}

int b200_dropout_colsum(const void* x, void* out, int dtype, int R, int N, const b200_dropout_t* drop, float* colsum,
                        void* workspace, size_t workspace_bytes, void* stream_) {
  B200_CHECK_ARG(drop != nullptr && out != nullptr && colsum != nullptr, "dropout_colsum: bad arguments");
  return launch_colsum(x, out, drop, dtype, R, N, nullptr, 1, colsum, workspace, workspace_bytes, (cudaStream_t)stream_);
}

static int check_epi(int epi, int act, const void* aux_in, int out_dtype) {
  B200_CHECK_ARG(epi >= B200_EPI_NONE && epi <= B200_EPI_MUL, "gemm: bad epilogue %d", epi);
  B200_CHECK_ARG(act >= B200_ACT_NONE && act <= B200_ACT_TANH, "gemm: bad activation %d", act);
  B200_CHECK_ARG(!(epi == B200_EPI_ADD || epi == B200_EPI_DACT || epi == B200_EPI_MUL) || aux_in != nullptr,
                 "gemm: epilogue needs aux_in");
  B200_CHECK_ARG(epi != B200_EPI_ACCUM || out_dtype == B200_F32, "gemm: ACCUM epilogue needs fp32 output");
  return 0;
}

int b200_gemm(const void* A, int lda, int a_layout, const void* B, int ldb, int b_layout, void* out, int ldo,
              int M, int N, int K, int dtype, int out_dtype, const float* bias, int epi, int act,
              const void* aux_in, void* aux_out, int ld_aux, const b200_dropout_t* drop, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  B200_CHECK_ARG(dtype == B200_F32 || dtype == B200_BF16, "gemm: bad dtype %d", dtype);
  B200_CHECK_ARG(dtype == B200_BF16 || out_dtype == B200_F32, "gemm: fp32 inputs need fp32 output");
  if (int rc = check_epi(epi, act, aux_in, out_dtype)) return rc;
  GemmArgs a{};
  a.M = M; a.N = N; a.K = K; a.mode = GEMM_DENSE; a.epi = epi; a.act = act;
  a.out_f32 = (out_dtype == B200_F32); a.ldo = ldo; a.ld_aux = ld_aux; a.out = out; a.aux_in = aux_in;
  a.aux_out = aux_out; a.bias = bias; a.k_splits = 1;
  if (drop != nullptr && drop->p > 0.f) { a.drop_state = drop->rng_state; a.drop_p = drop->p; a.drop_site = drop->site; }
  const int m_tiles = (M + B200_GROUP_TILE - 1) / B200_GROUP_TILE;
  if (dtype == B200_BF16) {   // (EPI_ACCUM: the launcher zeroes `out` when it decides to split K)
    return launch_gemm_tc(A, lda, a_layout, M, K, B, ldb, b_layout, N, K, a, m_tiles, 1, stream);
  }
  a.A = A; a.B = B;
  if (epi == B200_EPI_ACCUM) a.epi = B200_EPI_NONE;   // the fp32 kernel never splits K: plain stores
  a.sa_m = a_layout == B200_LAYOUT_K ? lda : 1; a.sa_k = a_layout == B200_LAYOUT_K ? 1 : lda;
  a.sb_n = b_layout == B200_LAYOUT_K ? ldb : 1; a.sb_k = b_layout == B200_LAYOUT_K ? 1 : ldb;
  return launch_gemm_simt(a, m_tiles, 1, stream);
}

int b200_ggemm(const void* A, int lda, const void* B, int b_layout, void* out, int ldo, int R, int N, int K, int G,
               const int32_t* tile_group, const int32_t* rows_used, int dtype, int out_dtype, const float* bias, int epi,
               int act, const void* aux_in, void* aux_out, int ld_aux, const b200_dropout_t* drop, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(R > 0 && R % B200_GROUP_TILE == 0, "ggemm: R=%d must be a positive multiple of %d", R,
                 B200_GROUP_TILE);
  B200_CHECK_ARG(N > 0 && K > 0 && G > 0 && tile_group != nullptr, "ggemm: bad arguments");
  B200_CHECK_ARG(epi != B200_EPI_ACCUM, "ggemm: ACCUM epilogue not supported");
  B200_CHECK_ARG(dtype == B200_BF16 || out_dtype == B200_F32, "ggemm: fp32 inputs need fp32 output");
  if (int rc = check_epi(epi, act, aux_in, out_dtype)) return rc;
  GemmArgs a{};
  a.M = R; a.N = N; a.K = K; a.mode = GEMM_GROUP_ROWS; a.epi = epi; a.act = act;
  a.out_f32 = (out_dtype == B200_F32); a.ldo = ldo; a.ld_aux = ld_aux; a.out = out; a.aux_in = aux_in;
  a.aux_out = aux_out; a.bias = bias; a.tile_group = tile_group; a.rows_used = rows_used; a.k_splits = 1;
  if (drop != nullptr && drop->p > 0.f) { a.drop_state = drop->rng_state; a.drop_p = drop->p; a.drop_site = drop->site; }
  a.b_group_rows = (b_layout == B200_LAYOUT_K) ? N : K;
  a.b_group_elems = (long long)N * K;
  const int m_tiles = R / B200_GROUP_TILE;
  if (dtype == B200_BF16) {
    if (b_layout == B200_LAYOUT_K)  // B stacked [G*N, K]
      return launch_gemm_tc(A, lda, B200_LAYOUT_K, R, K, B, K, b_layout, (long long)G * N, K, a, m_tiles, G, stream);
    // B stacked [G*K, N]
    return launch_gemm_tc(A, lda, B200_LAYOUT_K, R, K, B, N, b_layout, N, (long long)G * K, a, m_tiles, G, stream);
  }
  a.A = A; a.B = B; a.sa_m = lda; a.sa_k = 1;
  a.sb_n = b_layout == B200_LAYOUT_K ? K : 1; a.sb_k = b_layout == B200_LAYOUT_K ? 1 : N;
  return launch_gemm_simt(a, m_tiles, G, stream);
}

int b200_ggemm_wgrad(const void* A, int lda, const void* B, int ldb, float* out, int Mo, int No, int R, int G,
                     const int32_t* group_off, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(R > 0 && R % B200_GROUP_TILE == 0 && Mo > 0 && No > 0 && G > 0 && group_off != nullptr,
                 "ggemm_wgrad: bad arguments");
  GemmArgs a{};
  a.M = Mo; a.N = No; a.K = 0; a.mode = GEMM_GROUP_WGRAD; a.epi = B200_EPI_NONE; a.act = B200_ACT_NONE;
  a.out_f32 = 1; a.ldo = No; a.out = out; a.group_off = group_off; a.out_group_elems = (long long)Mo * No;
  a.k_splits = 1;
  const int m_tiles = (Mo + B200_GROUP_TILE - 1) / B200_GROUP_TILE;
  if (dtype == B200_BF16)
    return launch_gemm_tc(A, lda, B200_LAYOUT_MN, Mo, R, B, ldb, B200_LAYOUT_MN, No, R, a, m_tiles, G, stream);
  a.A = A; a.B = B; a.sa_m = 1; a.sa_k = lda; a.sb_n = 1; a.sb_k = ldb;
  return launch_gemm_simt(a, m_tiles, G, stream);
}

}  // extern "C"
