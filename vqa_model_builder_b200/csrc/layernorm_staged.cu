// Residual-add + LayerNorm, forward and backward, bf16, as persistent shared-memory-staged kernels (rowpipe.cuh).
// One block of 8 warps per SM (backward) or two (forward): lane 0 of warp 0 keeps stages-1 chunks of 8 rows in flight
// with bulk asynchronous copies into a ring of shared-memory stages (a ninth, dedicated producer warp would put three
// warps on one scheduler partition and cap every thread at 168 registers); each warp takes one row of a chunk.
// Backward keeps the dgamma / dbeta / bias-gradient partial sums of a block in REGISTERS (one block per SM leaves 255
// registers per thread) and folds them through shared memory once per (block, group) instead of once per row.
// Rows are handed out as equal contiguous ranges of the LIVE rows (unused 128-row tiles of a padded expert layout are
// skipped), so there is no partial last wave.  fp32 rows and D > 1024 stay on the register kernels of layernorm.cu.
#include "rowops.cuh"
#include "rowpipe.cuh"

namespace b200 {

int launch_ln_staged_reduce(const float* part, const int* part_group, int entries, int D, int G, int nz, float* dgamma,
                            float* dbeta, float* dcol, cudaStream_t stream);

namespace {

constexpr int ST_WARPS = 8;                        // consumer warps = rows per chunk
constexpr int ST_THREADS = ST_WARPS * 32;
constexpr int ST_MAX_STAGES = 8;
constexpr int ST_HDR = 256;                        // barriers + counters

struct StagedSmem {
  uint32_t full, empty;       // shared-window addresses of the barrier arrays
  int* s_count;
  int* grp;                   // tile -> group (copy of tile_group)
  int* live;                  // compacted list of live tiles
  unsigned char* stages;      // ring
  float* red;                 // [ST_WARPS][nz][D] (backward only)
};

__device__ __forceinline__ StagedSmem carve(unsigned char* smem, int tiles_pad, int stages, size_t stage_bytes) {
  StagedSmem s;
  s.full = ptx::smem_u32(smem);
  s.empty = s.full + 8 * ST_MAX_STAGES;
  s.s_count = reinterpret_cast<int*>(smem + 16 * ST_MAX_STAGES);
  s.grp = reinterpret_cast<int*>(smem + ST_HDR);
  s.live = s.grp + tiles_pad;
  s.stages = smem + ST_HDR + (size_t)tiles_pad * 8;
  s.red = reinterpret_cast<float*>(s.stages + (size_t)stages * stage_bytes);
  return s;
}

// Copies the tile->group map into shared memory and compacts the live tiles; returns the number of live rows.
__device__ __forceinline__ int live_rows(const int* __restrict__ tile_group, int tiles, int R, const StagedSmem& s) {
  if (tile_group == nullptr) return R;
  for (int t = threadIdx.x; t < tiles; t += blockDim.x) s.grp[t] = tile_group[t];
  __syncthreads();
  if (threadIdx.x < 32) {
    int base = 0;
    for (int t0 = 0; t0 < tiles; t0 += 32) {
      const int t = t0 + threadIdx.x;
      const bool f = t < tiles && s.grp[t] >= 0;
      const unsigned m = __ballot_sync(0xffffffffu, f);
      if (f) s.live[base + __popc(m & ((1u << threadIdx.x) - 1u))] = t;
      base += __popc(m);
    }
    if (threadIdx.x == 0) *s.s_count = base;
  }
  __syncthreads();
  return *s.s_count * B200_GROUP_TILE;
}

__device__ __forceinline__ int rows_per_block(int live, int grid) {
  const int q = (live + grid - 1) / grid;
  return (q + ST_WARPS - 1) / ST_WARPS * ST_WARPS;
}

// ------------------------------------------------------------------------------------------------------------
// forward: y = LN(drop?(x) + drop?(res)) * gamma[g] + beta[g]
// ------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(ST_THREADS, 2)
add_ln_fwd_staged_kernel(const bf16* __restrict__ x, const bf16* __restrict__ res, const float* __restrict__ gamma,
                         const float* __restrict__ beta, const int* __restrict__ tile_group, int tiles, int tiles_pad,
                         float eps, bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                         int R, int D, int stages, const unsigned long long* drop_state, float drop_p,
                         unsigned int drop_site, int drop_target) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nops = res != nullptr ? 2 : 1;
  const size_t op_bytes = (size_t)ST_WARPS * D * sizeof(bf16);
  const StagedSmem s = carve(smem, tiles_pad, stages, op_bytes * nops);
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      ptx::mbar_init(s.full + 8 * i, 1);
      ptx::mbar_init(s.empty + 8 * i, ST_WARPS);
    }
    ptx::fence_mbar_init();
  }
  pdl_trigger();
  pdl_wait();
  __syncthreads();
  const int live = live_rows(tile_group, tiles, R, s);
  const int q = rows_per_block(live, gridDim.x);
  const int v_begin = blockIdx.x * q, v_end = min(live, v_begin + q);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int nv = NV * 32;   // the staged kernels cover D == NV * 256 exactly: no per-vector guards

  // ---- feeder (warp 0, lane 0): chunk c of this block goes to stage c % stages
  const int nchunks = v_end > v_begin ? (v_end - v_begin + ST_WARPS - 1) / ST_WARPS : 0;
  auto issue = [&](int c) {
    const int v = v_begin + c * ST_WARPS;
    const int n = min(ST_WARPS, v_end - v);
    const long long phys = tile_group != nullptr ? (long long)s.live[v >> 7] * B200_GROUP_TILE + (v & 127) : v;
    const int st = c % stages;
    ptx::mbar_wait(s.empty + 8 * st, (((uint32_t)(c / stages)) & 1u) ^ 1u);
    const uint32_t bytes = (uint32_t)(n * D * sizeof(bf16));
    const uint32_t bar = s.full + 8 * st;
    ptx::mbar_arrive_expect_tx(bar, bytes * nops);
    const uint32_t dst = ptx::smem_u32(s.stages + (size_t)st * op_bytes * nops);
    bulk_g2s(dst, x + phys * D, bytes, bar);
    if (res != nullptr) bulk_g2s(dst + (uint32_t)op_bytes, res + phys * D, bytes, bar);
  };
  if (threadIdx.x == 0)
    for (int c = 0; c < min(stages - 1, nchunks); ++c) issue(c);

  // ---- consumers
  const DropState ds = drop_load(drop_state, drop_p, drop_site);
  float gam[NV][8], bet[NV][8];
  int g_cur = -2;
  RingPos rp;
  int ci = 0;
  for (int v = v_begin; v < v_end; v += ST_WARPS, ++ci) {
    if (threadIdx.x == 0 && ci + stages - 1 < nchunks) issue(ci + stages - 1);
    __syncwarp();
    const int n = min(ST_WARPS, v_end - v);
    int g = 0;
    long long phys = v;
    if (tile_group != nullptr) {
      const int t = s.live[v >> 7];
      g = s.grp[t];
      phys = (long long)t * B200_GROUP_TILE + (v & 127);
    }
    if (g != g_cur) {
      g_cur = g;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int vi = lane + 32 * j;
        if (vi < nv) {
          load_param<8>(gamma + (long long)g * D, vi, gam[j]);
          load_param<8>(beta + (long long)g * D, vi, bet[j]);
        }
      }
    }
    ptx::mbar_wait(s.full + 8 * rp.stage, rp.phase);
    if (warp < n) {
      const long long r = phys + warp;
      const unsigned char* sb = s.stages + (size_t)rp.stage * op_bytes * nops + (size_t)warp * D * sizeof(bf16);
      float xv[NV][8];
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int vi = lane + 32 * j;
        if (vi < nv) {
          unpack8(*reinterpret_cast<const uint4*>(sb + vi * 16), xv[j]);
          uint32_t keep = 0xFFu;
          if (ds.on && drop_target != 0) keep = drop_keep8(ds, (unsigned long long)r * nv + vi);
          if (ds.on && drop_target == 1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[j][u] *= (keep >> u) & 1u ? ds.inv_keep : 0.f;
          }
          if (res != nullptr) {
            float rv[8];
            unpack8(*reinterpret_cast<const uint4*>(sb + op_bytes + vi * 16), rv);
            if (ds.on && drop_target == 2) {
#pragma unroll
              for (int u = 0; u < 8; ++u) xv[j][u] = fmaf(rv[u], (keep >> u) & 1u ? ds.inv_keep : 0.f, xv[j][u]);
            } else {
#pragma unroll
              for (int u = 0; u < 8; ++u) xv[j][u] += rv[u];
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) sum += xv[j][u];
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) xv[j][u] = 0.f;
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s.empty + 8 * rp.stage);   // the row is in registers: release the stage early
      const float mean = warp_sum(sum) / D;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (lane + 32 * j < nv) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            xv[j][u] -= mean;
            sq = fmaf(xv[j][u], xv[j][u], sq);
          }
        }
      const float rstd = rsqrtf(warp_sum(sq) / D + eps);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int vi = lane + 32 * j;
        if (vi < nv) {
          float o[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) o[u] = fmaf(xv[j][u] * rstd, gam[j][u], bet[j][u]);
          *reinterpret_cast<uint4*>(y + r * D + vi * 8) = pack8(o);
        }
      }
      if (lane == 0) {
        mean_out[r] = mean;
        rstd_out[r] = rstd;
      }
    } else {
      if (lane == 0) ptx::mbar_arrive(s.empty + 8 * rp.stage);
    }
    rp.advance(stages);
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
template <int NV, bool COLSUM>
__global__ void __launch_bounds__(ST_THREADS, 1)
add_ln_bwd_staged_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ res,
                         const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                         const float* __restrict__ gamma, const int* __restrict__ tile_group, int tiles, int tiles_pad,
                         bf16* __restrict__ dsum, float* __restrict__ part, int* __restrict__ part_group, int R, int D,
                         int slots, int stages, const unsigned long long* drop_state, float drop_p,
                         unsigned int drop_site, int drop_target, bf16* __restrict__ d_dropped) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NZ = COLSUM ? 3 : 2;
  const int nops = res != nullptr ? 3 : 2;
  const size_t op_bytes = (size_t)ST_WARPS * D * sizeof(bf16);
  const StagedSmem s = carve(smem, tiles_pad, stages, op_bytes * nops);
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      ptx::mbar_init(s.full + 8 * i, 1);
      ptx::mbar_init(s.empty + 8 * i, ST_WARPS);
    }
    ptx::fence_mbar_init();
  }
  pdl_trigger();
  pdl_wait();
  __syncthreads();
  const int live = live_rows(tile_group, tiles, R, s);
  const int q = rows_per_block(live, gridDim.x);
  const int v_begin = blockIdx.x * q, v_end = min(live, v_begin + q);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int nv = NV * 32;   // the staged kernels cover D == NV * 256 exactly: no per-vector guards

  // ---- feeder (warp 0, lane 0): chunk c of this block goes to stage c % stages
  const int nchunks = v_end > v_begin ? (v_end - v_begin + ST_WARPS - 1) / ST_WARPS : 0;
  auto issue = [&](int c) {
    const int v = v_begin + c * ST_WARPS;
    const int n = min(ST_WARPS, v_end - v);
    const long long phys = tile_group != nullptr ? (long long)s.live[v >> 7] * B200_GROUP_TILE + (v & 127) : v;
    const int st = c % stages;
    ptx::mbar_wait(s.empty + 8 * st, (((uint32_t)(c / stages)) & 1u) ^ 1u);
    const uint32_t bytes = (uint32_t)(n * D * sizeof(bf16));
    const uint32_t bar = s.full + 8 * st;
    ptx::mbar_arrive_expect_tx(bar, bytes * nops);
    const uint32_t dst = ptx::smem_u32(s.stages + (size_t)st * op_bytes * nops);
    bulk_g2s(dst, dy + phys * D, bytes, bar);
    bulk_g2s(dst + (uint32_t)op_bytes, x + phys * D, bytes, bar);
    if (res != nullptr) bulk_g2s(dst + 2 * (uint32_t)op_bytes, res + phys * D, bytes, bar);
  };
  if (threadIdx.x == 0)
    for (int c = 0; c < min(stages - 1, nchunks); ++c) issue(c);

  // ---- consumers
  const DropState ds = drop_load(drop_state, drop_p, drop_site);
  const bool dropping = ds.on && drop_target != 0;
  constexpr int NC = COLSUM ? NV : 1;
  float gam[NV][8], acc_g[NV][8], acc_b[NV][8], acc_c[NC][8];
  int g_cur = -2, slot = 0;
  const int ctid = threadIdx.x;   // 0..255

  auto flush = [&](int g) {
    __syncthreads();
    float* mine = s.red + (size_t)warp * NZ * D;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nv) {
#pragma unroll
        for (int u = 0; u < 8; u += 4) {
          *reinterpret_cast<float4*>(mine + vi * 8 + u) =
              make_float4(acc_g[j][u], acc_g[j][u + 1], acc_g[j][u + 2], acc_g[j][u + 3]);
          *reinterpret_cast<float4*>(mine + D + vi * 8 + u) =
              make_float4(acc_b[j][u], acc_b[j][u + 1], acc_b[j][u + 2], acc_b[j][u + 3]);
          if (COLSUM)
            *reinterpret_cast<float4*>(mine + 2 * D + vi * 8 + u) =
                make_float4(acc_c[COLSUM ? j : 0][u], acc_c[COLSUM ? j : 0][u + 1], acc_c[COLSUM ? j : 0][u + 2],
                            acc_c[COLSUM ? j : 0][u + 3]);
        }
      }
    }
    __syncthreads();
    float* out = part + ((size_t)blockIdx.x * slots + slot) * NZ * D;
    for (int i = ctid; i < NZ * D; i += ST_WARPS * 32) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < ST_WARPS; ++w) t += s.red[(size_t)w * NZ * D + i];
      out[i] = t;
    }
    if (ctid == 0) part_group[blockIdx.x * slots + slot] = g;
    ++slot;
  };

  RingPos rp;
  int ci = 0;
  float mean_nx = 0.f, rstd_nx = 0.f;
  if (v_begin + warp < v_end) {
    const long long p0 =
        (tile_group != nullptr ? (long long)s.live[v_begin >> 7] * B200_GROUP_TILE + (v_begin & 127) : v_begin) + warp;
    mean_nx = __ldg(mean_in + p0);
    rstd_nx = __ldg(rstd_in + p0);
  }
  for (int v = v_begin; v < v_end; v += ST_WARPS, ++ci) {
    if (threadIdx.x == 0 && ci + stages - 1 < nchunks) issue(ci + stages - 1);
    __syncwarp();
    const int n = min(ST_WARPS, v_end - v);
    int g = 0;
    long long phys = v;
    if (tile_group != nullptr) {
      const int t = s.live[v >> 7];
      g = s.grp[t];
      phys = (long long)t * B200_GROUP_TILE + (v & 127);
    }
    if (g != g_cur) {
      if (g_cur >= 0) flush(g_cur);
      g_cur = g;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int vi = lane + 32 * j;
        if (vi < nv) load_param<8>(gamma + (long long)g * D, vi, gam[j]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc_g[j][u] = 0.f;
          acc_b[j][u] = 0.f;
          if (COLSUM) acc_c[COLSUM ? j : 0][u] = 0.f;
        }
      }
    }
    const long long r = phys + warp;
    // statistics of this warp's row in the NEXT chunk are fetched now: their DRAM latency hides behind this chunk
    const float mean = mean_nx, rstd = rstd_nx;
    {
      const int vn = v + ST_WARPS;
      if (vn + warp < v_end) {
        const long long pn = (tile_group != nullptr ? (long long)s.live[vn >> 7] * B200_GROUP_TILE + (vn & 127) : vn) + warp;
        mean_nx = __ldg(mean_in + pn);
        rstd_nx = __ldg(rstd_in + pn);
      }
    }
    ptx::mbar_wait(s.full + 8 * rp.stage, rp.phase);
    if (warp < n) {
      const unsigned char* sb = s.stages + (size_t)rp.stage * op_bytes * nops + (size_t)warp * D * sizeof(bf16);
      float xh[NV][8], gg[NV][8];
      uint32_t keep[NV];
      float s1 = 0.f, s2 = 0.f;
      const float nmr = -mean * rstd;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int vi = lane + 32 * j;
        keep[j] = 0xFFu;
        float dv[8];
        unpack8(*reinterpret_cast<const uint4*>(sb + vi * 16), dv);
        unpack8(*reinterpret_cast<const uint4*>(sb + op_bytes + vi * 16), xh[j]);
        if (dropping) keep[j] = drop_keep8(ds, (unsigned long long)r * nv + vi);
        if (dropping && drop_target == 1) {
#pragma unroll
          for (int u = 0; u < 8; ++u) xh[j][u] *= (keep[j] >> u) & 1u ? ds.inv_keep : 0.f;
        }
        if (res != nullptr) {
          float rv[8];
          unpack8(*reinterpret_cast<const uint4*>(sb + 2 * op_bytes + vi * 16), rv);
          if (dropping && drop_target == 2) {
#pragma unroll
            for (int u = 0; u < 8; ++u) xh[j][u] = fmaf(rv[u], (keep[j] >> u) & 1u ? ds.inv_keep : 0.f, xh[j][u]);
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) xh[j][u] += rv[u];
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float xhat = fmaf(xh[j][u], rstd, nmr);
          const float d = dv[u];
          acc_g[j][u] = fmaf(d, xhat, acc_g[j][u]);
          acc_b[j][u] += d;
          const float t = d * gam[j][u];
          xh[j][u] = xhat;
          gg[j][u] = t;
          s1 += t;
          s2 = fmaf(t, xhat, s2);
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s.empty + 8 * rp.stage);   // operands are in registers: release the stage
      s1 = warp_sum(s1) * (1.0f / D);
      s2 = warp_sum(s2) * (1.0f / D);
      const float c1 = -rstd * s1, c2 = -rstd * s2;     // dsum = rstd*gg - rstd*s1 - xhat*rstd*s2
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int vi = lane + 32 * j;
        float o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) o[u] = fmaf(xh[j][u], c2, fmaf(gg[j][u], rstd, c1));
        *reinterpret_cast<uint4*>(dsum + r * D + vi * 8) = pack8(o);
        if (dropping) {
#pragma unroll
          for (int u = 0; u < 8; ++u) o[u] *= (keep[j] >> u) & 1u ? ds.inv_keep : 0.f;
          if (d_dropped != nullptr) *reinterpret_cast<uint4*>(d_dropped + r * D + vi * 8) = pack8(o);
        }
        if (COLSUM) {
#pragma unroll
          for (int u = 0; u < 8; ++u) acc_c[COLSUM ? j : 0][u] += o[u];
        }
      }
    } else {
      if (lane == 0) ptx::mbar_arrive(s.empty + 8 * rp.stage);
    }
    rp.advance(stages);
  }
  if (g_cur >= 0) flush(g_cur);
  if (ctid == 0)
    for (int k = slot; k < slots; ++k) part_group[blockIdx.x * slots + k] = -1;
}

// out_z[g][c] = sum over the (block, slot) entries e with part_group[e] == g of part[e][z][c], in entry order.
__global__ void __launch_bounds__(1024)
ln_staged_reduce_kernel(const float* __restrict__ part, const int* __restrict__ part_group, int entries, int D, int nz,
                        float* __restrict__ out0, float* __restrict__ out1, float* __restrict__ out2) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32][33];
  const int xl = threadIdx.x & 31, yl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + xl;
  const int g = blockIdx.y, z = blockIdx.z;
  float s0 = 0.f;
  if (c < D)
    for (int e = yl; e < entries; e += 32)
      if (part_group[e] == g) s0 += part[((size_t)e * nz + z) * D + c];
  red[yl][xl] = s0;
  __syncthreads();
  if (yl == 0 && c < D) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) t += red[j][xl];
    (z == 0 ? out0 : (z == 1 ? out1 : out2))[(size_t)g * D + c] = t;
  }
}

struct Geom {
  int grid, slots, stages, tiles, tiles_pad;
  size_t smem;
};

inline bool staged_geom(int R, int D, int nops, bool grouped, int blocks_per_sm, size_t red_bytes, Geom& gm) {
  if (D % 256 != 0 || D > 1024) return false;    // D == NV * 256 exactly (NV = 1..4)
  if (grouped && R % B200_GROUP_TILE != 0) return false;
  gm.tiles = grouped ? R / B200_GROUP_TILE : 0;
  if (gm.tiles > 8192) return false;
  gm.tiles_pad = (gm.tiles + 31) / 32 * 32;
  const int chunks = (R + ST_WARPS - 1) / ST_WARPS;
  gm.grid = chunks < num_sms() * blocks_per_sm ? chunks : num_sms() * blocks_per_sm;
  const int q = ((R + gm.grid - 1) / gm.grid + ST_WARPS - 1) / ST_WARPS * ST_WARPS;
  gm.slots = grouped ? q / B200_GROUP_TILE + 2 : 1;
  const size_t stage = (size_t)nops * ST_WARPS * D * sizeof(bf16);
  const size_t budget = (size_t)(blocks_per_sm == 1 ? 220 : 108) * 1024;
  const size_t fixed = ST_HDR + (size_t)gm.tiles_pad * 8 + red_bytes + 128;
  if (fixed + 2 * stage > budget) return false;
  size_t st = (budget - fixed) / stage;
  const size_t need = (size_t)(q / ST_WARPS);        // never more stages than chunks per block
  if (st > need) st = need < 2 ? 2 : need;
  gm.stages = (int)(st > ST_MAX_STAGES ? ST_MAX_STAGES : st);
  gm.smem = fixed + gm.stages * stage;
  return true;
}

}  // namespace

bool ln_staged_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VQA_LN_STAGED");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

size_t ln_bwd_staged_ws(int R, int D) {
  // entries = grid * (q/128 + 2) with grid <= min(chunks, SMs) and q ~ R/grid rounded up to 8 rows:
  // grid*q/128 <= (R + 8*grid)/128, so entries <= R/128 + 3.1*grid + 1 whatever the SM count (<= 160) is
  const long long chunks = (R + ST_WARPS - 1) / ST_WARPS;
  const long long grid = chunks < 160 ? chunks : 160;
  const size_t entries = (size_t)(R / B200_GROUP_TILE + (31 * grid + 9) / 10 + 1);
  return entries * (3 * (size_t)D * sizeof(float) + sizeof(int)) + 512;
}

// returns 0 on success, -1 when the shape is not covered (caller falls back), > 0 on error
int launch_add_ln_fwd_staged(const bf16* x, const bf16* res, const float* gamma, const float* beta,
                             const int* tile_group, float eps, bf16* y, float* mean, float* rstd, int R, int D,
                             const unsigned long long* dst, float dp, unsigned int dsite, int drop_target,
                             cudaStream_t stream) {
  Geom gm;
  if (!staged_geom(R, D, res != nullptr ? 2 : 1, tile_group != nullptr, 2, 0, gm)) return -1;
  const int nvv = row_nv<bf16>(D);
  if (nvv > 4) return -1;
#define B200_FWD_CASE(NVC)                                                                                        \
  case NVC: {                                                                                                     \
    static bool attr_set = false;                                                                                 \
    if (!attr_set) {                                                                                              \
      B200_CUDA(cudaFuncSetAttribute(add_ln_fwd_staged_kernel<NVC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     227 * 1024));                                                                \
      attr_set = true;                                                                                            \
    }                                                                                                             \
    launch_kernel(add_ln_fwd_staged_kernel<NVC>, dim3(gm.grid), dim3(ST_THREADS), gm.smem, stream, x, res, gamma, \
                  beta, tile_group, gm.tiles, gm.tiles_pad, eps, y, mean, rstd, R, D, gm.stages, dst, dp, dsite,  \
                  drop_target);                                                                                   \
  } break;
  switch (nvv) {
    B200_FWD_CASE(1)
    B200_FWD_CASE(2)
    B200_FWD_CASE(3)
    B200_FWD_CASE(4)
    default: return -1;
  }
#undef B200_FWD_CASE
  B200_LAUNCH_CHECK("add_ln_fwd_staged_kernel");
  count_launch();
  return 0;
}

int launch_add_ln_bwd_staged(const bf16* dy, const bf16* x, const bf16* res, const float* mean, const float* rstd,
                             const float* gamma, const int* tile_group, int G, bf16* dsum, float* dgamma, float* dbeta,
                             float* d_colsum, int R, int D, const unsigned long long* dst, float dp, unsigned int dsite,
                             int drop_target, bf16* d_dropped, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream) {
  Geom gm;
  const int nz = d_colsum != nullptr ? 3 : 2;
  const size_t red_bytes = (size_t)ST_WARPS * nz * D * sizeof(float);
  if (!staged_geom(R, D, res != nullptr ? 3 : 2, tile_group != nullptr, 1, red_bytes, gm)) return -1;
  const int nvv = row_nv<bf16>(D);
  if (nvv > 4) return -1;
  const int entries = gm.grid * gm.slots;
  const size_t need = (size_t)entries * (nz * (size_t)D * sizeof(float) + sizeof(int)) + 256;
  if (workspace_bytes < need) return -1;
  float* part = (float*)workspace;
  int* part_group = (int*)((char*)workspace + ((size_t)entries * nz * D * sizeof(float) + 255) / 256 * 256);
#define B200_BWD_LAUNCH(NVC, CS)                                                                                   \
  {                                                                                                                \
    static bool attr_set = false;                                                                                  \
    if (!attr_set) {                                                                                               \
      B200_CUDA(cudaFuncSetAttribute(add_ln_bwd_staged_kernel<NVC, CS>,                                            \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));                    \
      attr_set = true;                                                                                             \
    }                                                                                                              \
    launch_kernel(add_ln_bwd_staged_kernel<NVC, CS>, dim3(gm.grid), dim3(ST_THREADS), gm.smem, stream, dy, x, res, \
                  mean, rstd, gamma, tile_group, gm.tiles, gm.tiles_pad, dsum, part, part_group, R, D, gm.slots,   \
                  gm.stages, dst, dp, dsite, drop_target, d_dropped);                                              \
  }
#define B200_BWD_CASE(NVC)                                 \
  case NVC:                                                \
    if (nz == 3) B200_BWD_LAUNCH(NVC, true)                \
    else B200_BWD_LAUNCH(NVC, false)                       \
    break;
  switch (nvv) {
    B200_BWD_CASE(1)
    B200_BWD_CASE(2)
    B200_BWD_CASE(3)
    B200_BWD_CASE(4)
    default: return -1;
  }
#undef B200_BWD_CASE
#undef B200_BWD_LAUNCH
  B200_LAUNCH_CHECK("add_ln_bwd_staged_kernel");
  count_launch();
  return launch_ln_staged_reduce(part, part_group, entries, D, G, nz, dgamma, dbeta, d_colsum, stream);
}

int launch_ln_staged_reduce(const float* part, const int* part_group, int entries, int D, int G, int nz, float* dgamma,
                            float* dbeta, float* dcol, cudaStream_t stream) {
  dim3 grid((D + 31) / 32, G, nz);
  launch_kernel(ln_staged_reduce_kernel, grid, dim3(1024), 0, stream, part, part_group, entries, D, nz, dgamma, dbeta,
                dcol);
  B200_LAUNCH_CHECK("ln_staged_reduce_kernel");
  count_launch();
  return 0;
}

}  // namespace b200
