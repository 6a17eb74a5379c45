/*
 * b200vqa.h — C-ABI of libb200vqa.so: the sm_100a kernels behind the AutoViVQA fusion + MOE hot path.
 *
 * The reference (richardnguyen0715/vqa-model-builder) has no native/FFI layer: its hot path is Python
 * nn.Modules whose arithmetic is ATen library calls.  Each entry point below replaces the group of ATen
 * calls issued by the cited reference lines; the Python drop-in modules in vqa_model_builder_b200/ bind
 * these symbols with ctypes (see INTEGRATION.md for the reference-side binding).
 *
 * Conventions
 *   - every function returns 0 on success or a B200_ERR_* code; b200_last_error_string() gives the
 *     thread-local message.  No exception crosses the ABI.
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch caching allocator); the library never
 *     allocates or frees device memory and never synchronises: launches are asynchronous on `stream`
 *     (a cudaStream_t passed as void*), so calls are CUDA-graph capturable.
 *   - `dtype` is B200_F32 or B200_BF16 and names the activation storage type; router outputs, LayerNorm
 *     statistics, softmax statistics and all parameter gradients are always fp32.
 *   - activations are dense row-major; row pitches must be multiples of 16 bytes.
 *   - integer maps are int32.
 */
#ifndef B200VQA_H_
#define B200VQA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VQA_ABI_VERSION 2

enum { B200_OK = 0, B200_ERR_INVALID = 1, B200_ERR_CUDA = 2, B200_ERR_UNSUPPORTED = 3 };
enum { B200_F32 = 0, B200_BF16 = 1 };
/* operand storage for GEMMs: K = reduction dimension contiguous ([rows, k]); MN = [k, rows] */
enum { B200_LAYOUT_K = 0, B200_LAYOUT_MN = 1 };
enum { B200_ACT_NONE = 0, B200_ACT_GELU = 1, B200_ACT_RELU = 2, B200_ACT_SILU = 3, B200_ACT_TANH = 4 };
/* GEMM epilogues */
enum {
  B200_EPI_NONE = 0,     /* out = acc (+bias)                                              */
  B200_EPI_ACT = 1,      /* pre = acc+bias; aux_out = pre (if given); out = act(pre)        */
  B200_EPI_ADD = 2,      /* out = acc (+bias) + aux_in          (residual add)              */
  B200_EPI_DACT = 3,     /* out = acc * act'(aux_in)            (backward through act)      */
  B200_EPI_ACCUM = 4,    /* out(fp32) = acc, K may be split (zeroed + atomic partial sums)   */
  B200_EPI_ACT_D = 5,    /* pre = acc+bias; out = dropout(act(pre)); aux_out = act'(pre) * keep-scale: the factor
                            the backward pass multiplies by (derivative and dropout mask, evaluated once, while the
                            exponential of the activation is at hand)                                            */
  B200_EPI_MUL = 6       /* out = acc * aux_in               (backward through EPI_ACT_D: no erf, no RNG)        */
};

#define B200_GROUP_TILE 128 /* expert row segments are padded to this many rows */

/* Dropout (nn.Dropout / MHA attention dropout in train mode: vqa_model.py:258-277, expert_types.py:56,81-83).
 * Masks are never stored: every kernel regenerates keep/drop decisions from a counter-based Philox4x32-7 stream
 * keyed by (rng_state[0] = seed, rng_state[1] = step offset, site, element index), so forward and backward of the same
 * site see the same mask (one Philox call covers 8 consecutive elements, 16 random bits each: p is resolved to
 * 2^-16).  rng_state is a DEVICE pointer to two uint64 (read at execution time: CUDA-graph replays
 * get fresh masks when the offset is advanced on the device).  drop == NULL or p == 0 disables dropout.            */
typedef struct {
  const unsigned long long* rng_state;
  float p;
  unsigned int site;
} b200_dropout_t;

/* ---- lifecycle ------------------------------------------------------------------------------- */
int b200_init(int device);
const char* b200_last_error_string(void);
int b200_abi_version(void);
/* kernels launched by this library since the last reset (host-side counter; bench.py's gpu_launches) */
long long b200_launch_count(void);
void b200_reset_launch_count(void);

/* keep-scales (0 or 1/(1-p)) of elements [0,n) of a dropout site, as fp32 — test/debug aid that materialises the
 * mask the fused kernels regenerate on the fly.                                                              */
int b200_dropout_mask(const b200_dropout_t* drop, long long n, float* out, void* stream);
/* out[i] = x[i] * keep_scale(i)  (element index i = row*ld + col of a dense [.., ld] tensor): backward of a dropout
 * that was fused into a GEMM epilogue with a residual add.                                                     */
int b200_dropout_apply(const void* x, void* out, long long n, int dtype, const b200_dropout_t* drop, void* stream);

/* ---- elementwise ----------------------------------------------------------------------------- */
/* dtype conversion (autocast's weight/activation casts; torch/amp/autocast_mode, called around every
 * nn.Linear in the reference under training_pipeline.py:457). */
int b200_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, void* stream);
/* out[g, :] = sum over rows r with group(r)==g of x[r, :]   (bias gradients).  tile_group==NULL: G=1.
 * row_limit (device int, may be NULL) bounds the rows actually used. workspace >= b200_colsum_ws(R,N) */
size_t b200_colsum_ws(int R, int N);
int b200_colsum(const void* x, int dtype, int R, int N, const int32_t* tile_group, int G, float* out,
                void* workspace, size_t workspace_bytes, void* stream);
/* out[r,c] = x[r,c] * keep_scale(r*N + c) and colsum[c] = sum_r out[r,c] in one pass: the backward of
 * "dropout(x W^T + b) + residual" (nn.TransformerEncoderLayer's dropout1/dropout2, generative_vqa_model.py:203-214)
 * needs dropout(dy) as GEMM operand and its column sums as the bias gradient.  workspace >= b200_colsum_ws(R,N). */
int b200_dropout_colsum(const void* x, void* out, int dtype, int R, int N, const b200_dropout_t* drop, float* colsum,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- GEMM ------------------------------------------------------------------------------------ */
/* out[M,N] = A[M,K] * B[N,K]^T with epilogue.  Replaces nn.Linear fwd/bwd (vqa_model.py:258-271,
 * expert_types.py:54-55, fusion_approaches.py:210-241) — F.linear -> cuBLAS in the reference.
 * a_layout/b_layout select [rows,K] (B200_LAYOUT_K) or [K,rows] (B200_LAYOUT_MN) storage, which gives
 * forward (K,K), dgrad (K,MN) and wgrad (MN,MN) from one kernel.  dtype BF16 -> tcgen05/TMEM/TMA kernel,
 * F32 -> SIMT fp32 kernel (validation mode).  out_dtype may be F32 for BF16 inputs (wgrad).
 * lda/ldb/ldo/ld_aux are row pitches in elements. bias is fp32 [N] or NULL.  With `drop`, EPI_ACT applies
 * dropout after the activation, EPI_DACT multiplies by the same mask, EPI_ADD computes dropout(acc+bias)+aux_in
 * (element index = row*ldo + col).                                                                        */
int b200_gemm(const void* A, int lda, int a_layout, const void* B, int ldb, int b_layout, void* out,
              int ldo, int M, int N, int K, int dtype, int out_dtype, const float* bias, int epi,
              int act, const void* aux_in, void* aux_out, int ld_aux, const b200_dropout_t* drop, void* stream);

/* Grouped GEMM over expert row segments (FeedForwardExpert fc1/fc2 fwd + dgrad, expert_types.py:79-83,
 * evaluated sparsely instead of moe_layer.py:151-168's dense loop).  A is [R, K] (permuted rows, each
 * expert's segment padded to B200_GROUP_TILE rows); tile_group[R/128] gives the expert of every 128-row
 * tile (-1 = unused tile).  B is the stacked expert weight [G, N, K] (b_layout K) or [G, K, N] (MN).
 * bias is [G, N] fp32 or NULL.  rows_used (device int32, may be NULL) bounds the rows in use: 128-row tiles at or
 * beyond it are not visited (expert-parallel receive buffers are sized for the worst case).          */
int b200_ggemm(const void* A, int lda, const void* B, int b_layout, void* out, int ldo, int R, int N,
               int K, int G, const int32_t* tile_group, const int32_t* rows_used, int dtype, int out_dtype,
               const float* bias, int epi, int act, const void* aux_in, void* aux_out, int ld_aux,
               const b200_dropout_t* drop, void* stream);

/* Grouped weight gradient: out[g] (fp32 [Mo, No]) = A[rows of g, :Mo]^T * B[rows of g, :No], rows of g =
 * [group_off[g], group_off[g+1]) (device int32, multiples of 128; pad rows must be zero in A or B).  */
int b200_ggemm_wgrad(const void* A, int lda, const void* B, int ldb, float* out, int Mo, int No, int R,
                     int G, const int32_t* group_off, int dtype, void* stream);

/* ---- gated linear unit ------------------------------------------------------------------------- */
/* GatedLinearExpert (expert_types.py:448-515): pre = fc1(x) is [R, 2F] = [value | gate] (row pitch ld_pre);
 * h[r,c] = dropout(value[r,c] * sigmoid(gate[r,c]))  (dropout element index = r*F + c).  Backward writes
 * dpre = [dh*mask*sigmoid(gate) | dh*mask*value*sigmoid'(gate)].  tile_group (may be NULL) skips the rows of unused
 * 128-row tiles of a grouped expert layout.  fp32 mode uses the same fast sigmoid (ex2/rcp.approx, rel. err ~1e-7). */
int b200_glu_fwd(const void* pre, int ld_pre, void* h, int R, int F, int dtype, const int32_t* tile_group,
                 const b200_dropout_t* drop, void* stream);
int b200_glu_bwd(const void* dh, const void* pre, int ld_pre, void* dpre, int R, int F, int dtype,
                 const int32_t* tile_group, const b200_dropout_t* drop, void* stream);

/* ---- LayerNorm (+ residual) ------------------------------------------------------------------- */
/* y = LN(x + res) * gamma[g] + beta[g]   (res may be NULL).  nn.LayerNorm after the residual adds at
 * vqa_model.py:301,305,309, expert_types.py:85-90, moe_layer.py:171, fusion_approaches.py:268-279.
 * tile_group (may be NULL -> group 0) picks per-expert affine parameters gamma/beta [G, D].
 * Saves mean/rstd [R] for backward.  drop_target: 0 none, 1 dropout on x, 2 dropout on res (the branch).  */
int b200_add_ln_fwd(const void* x, const void* res, const float* gamma, const float* beta,
                    const int32_t* tile_group, float eps, void* y, float* mean, float* rstd, int R, int D,
                    int dtype, const b200_dropout_t* drop, int drop_target, void* stream);
/* dsum = d(x+res); dgamma/dbeta [G, D] fp32 (overwritten).  With dropout, d_dropped receives the gradient of the
 * dropped operand (dsum * mask / (1-p)).  d_colsum [G, D] (or NULL) receives the column sums of that gradient
 * (d_dropped with dropout, else dsum): the bias gradient of the Linear that produced the operand, for free.
 * workspace >= b200_add_ln_bwd_ws(R, D)                                                                  */
size_t b200_add_ln_bwd_ws(int R, int D);
int b200_add_ln_bwd(const void* dy, const void* x, const void* res, const float* mean, const float* rstd,
                    const float* gamma, const int32_t* tile_group, int G, void* dsum, float* dgamma,
                    float* dbeta, float* d_colsum, int R, int D, int dtype, const b200_dropout_t* drop,
                    int drop_target, void* d_dropped, void* workspace, size_t workspace_bytes, void* stream);

/* ---- attention ------------------------------------------------------------------------------- */
/* o[b,t,h,:] = softmax_s(scale * q[b,t,h,:].k[b,s,h,:] + mask) v[b,s,h,:]; nn.MultiheadAttention core
 * (vqa_model.py:300,304; fusion_approaches.py:262-277; TransformerEncoderLayer self-attn in
 * generative_vqa_model.py:203-214).  q/k/v are [B, T|S, H, dh] views with row pitches ldq/ldk/ldv
 * (elements; lets q,k,v alias one packed in-proj output).  key_pad [B,S] uint8, 1 = ignore, or NULL.
 * causal != 0 (needs T == S): key s is visible to query t only if s <= t — the tgt_mask of the generative decoder's
 * self-attention (generative_vqa_model.py:404-406,448-451; nn.TransformerDecoderLayer).
 * lse [B,H,T] fp32 is saved for backward.  The [B,H,T,S] score tensor never touches HBM.  `drop` applies
 * attention-probability dropout (element index = ((b*H+h)*T+t)*S+s).                                 */
int b200_attn_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                  const uint8_t* key_pad, int causal, void* o, int ldo, float* lse, int B, int H, int T, int S, int dh,
                  float scale, int dtype, const b200_dropout_t* drop, void* stream);
int b200_attn_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                  const uint8_t* key_pad, int causal, const void* o, int ldo, const void* d_o, int lddo,
                  const float* lse, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv, int B,
                  int H, int T, int S, int dh, float scale, int dtype, const b200_dropout_t* drop, void* stream);

/* ---- generative decoder glue ---------------------------------------------------------------- */
/* out[n,:] = dropout(table[ids[n],:] + pos[n % T,:]): nn.Embedding + PositionalEncoding (+ its dropout) of the answer
 * decoder (generative_vqa_model.py:400-402, 453-476).  ids int32 [N] (N = B*T, clamped to [0,V)), table [V,D] in the
 * compute dtype, pos fp32 [>= T, D] (the registered `pe` buffer).  Dropout element index = n*D + d.              */
int b200_embed_fwd(const int32_t* ids, const void* table, const float* pos, void* out, int N, int T, int D, int V,
                   int dtype, const b200_dropout_t* drop, void* stream);
/* dtable[ids[n],:] += mask * dout[n,:]  (fp32 atomics into a caller-zeroed or accumulating [V,D] buffer).       */
int b200_embed_bwd(const int32_t* ids, const void* dout, float* dtable, int N, int D, int V, int dtype,
                   const b200_dropout_t* drop, void* stream);
/* nn.CrossEntropyLoss(ignore_index, label_smoothing) with mean reduction over the non-ignored rows
 * (generative_vqa_model.py:508-511,585-587; classification loss of vqa_model.py:705-713 with smoothing 0).
 * logits [R, C] with row pitch ld (elements), labels int32 [R].  One streaming pass: loss_rows[R], lse[R] (saved for
 * backward), loss[1] = sum(loss_rows) / n_valid, n_valid[1].  Backward: one pass, dlogits = dloss/n_valid *
 * (softmax - (1-eps) onehot - eps/C), zeros for ignored rows; dlogits may alias logits.                          */
int b200_ce_fwd(const void* logits, long long ld, const int32_t* labels, int R, int C, int ignore_index, float smoothing,
                int dtype, float* loss_rows, float* lse, float* loss, float* n_valid, void* stream);
int b200_ce_bwd(const void* logits, long long ld, const int32_t* labels, const float* lse, int R, int C,
                int ignore_index, float smoothing, int dtype, const float* dloss, const float* n_valid, void* dlogits,
                long long ldd, void* stream);

/* ---- MOE router ------------------------------------------------------------------------------ */
/* TopKRouter / NoisyTopKRouter forward (router.py:105-178, 287-366): logits = x Wg^T in fp32,
 * optional noise eps * softplus(x Wn^T) * noise_std, softmax, top-k (descending, lowest index wins
 * exact ties), renormalise; clean probs and the Switch load-balance loss
 *   loss = lb_weight * E * sum_e (count_e / N) * (sum_n p_clean[n,e] / N).
 * Outputs: idx int32 [N,K], w fp32 [N,K], topk_sum fp32 [N], probs fp32 [N,E] (clean),
 * probs_noisy fp32 [N,E] (only when eps != NULL), counts fp32 [E], psum fp32 [E] = sum_n p_clean[n,e] (may be
 * NULL; with counts it is what ranks all-reduce for a global aux loss), loss fp32 [1], noise_scale_mean fp32 [1]
 * (only when eps != NULL).  workspace >= b200_router_ws(N, E).                                          */
size_t b200_router_ws(int N, int E);
int b200_router_fwd(const void* x, int dtype, const float* w_gate, const float* w_noise, const float* eps,
                    float noise_std, float lb_weight, int N, int D, int E, int K, int32_t* idx, float* w,
                    float* topk_sum, float* probs, float* probs_noisy, float* counts, float* psum, float* loss,
                    float* noise_scale_mean, void* workspace, size_t workspace_bytes, void* stream);
/* Backward of the above.  d_w [N,K] (may be NULL), d_loss device fp32 scalar (may be NULL), d_probs [N,E] fp32
 * (may be NULL): gradient that reached the clean probabilities directly (aux_outputs['router_probs'] is
 * differentiable in the reference, router.py:140,328; e.g. solvers/losses/vqa_losses.py:543-573).
 * Produces dx [N,D] (overwrite), d_w_gate [E,D] fp32, d_w_noise [E,D] fp32 (if noisy).
 * workspace >= b200_router_bwd_ws(N, D, E).                                                         */
size_t b200_router_bwd_ws(int N, int D, int E);
int b200_router_bwd(const void* x, int dtype, const float* w_gate, const float* w_noise, const float* eps,
                    float noise_std, float lb_weight, int N, int D, int E, int K, const int32_t* idx,
                    const float* w, const float* topk_sum, const float* probs, const float* probs_noisy,
                    const float* counts, const float* d_w, const float* d_loss, const float* d_probs, void* dx,
                    float* d_w_gate, float* d_w_noise, void* workspace, size_t workspace_bytes, void* stream);

/* ---- MOE dispatch / combine ------------------------------------------------------------------- */
/* Upper bound of padded rows for NK (token,slot) pairs over E experts (host-side, no sync).        */
int b200_moe_max_rows(int NK, int E);
/* Routing plan (replaces nonzero/any/len host syncs of moe_layer.py:151-160, 317-337).  idx [NK] int32
 * (entries <0 or >=E are dropped).  Canonical order = stable sort of the flattened (n,k) list by expert
 * id (== (expert asc, token asc), the order SparseMOELayer's nonzero() produces, moe_layer.py:326).
 * counts[E], cmp_off[E+1] (compact offsets), pad_off[E+1] (offsets with every segment padded to 128),
 * dest_row[NK] (row in the padded layout, -1 dropped), cmp_pos[NK] (position in the compact canonical
 * order, -1 dropped), row_src[Rmax] (flattened (n,k) of each padded row, -1 = padding),
 * tile_group[Rmax/128] (expert per 128-row tile, -1 unused), cmp_src[NK] (may be NULL: flattened (n,k) of each
 * compact position, -1 beyond the routed pairs — the inverse of cmp_pos, used by the expert-parallel dispatch).
 * workspace >= b200_moe_plan_ws(NK,E).                                                              */
size_t b200_moe_plan_ws(int NK, int E);
int b200_moe_plan(const int32_t* idx, int NK, int E, int Rmax, int32_t* counts, int32_t* cmp_off,
                  int32_t* pad_off, int32_t* dest_row, int32_t* cmp_pos, int32_t* row_src,
                  int32_t* tile_group, int32_t* cmp_src, void* workspace, size_t workspace_bytes, void* stream);
/* SparseMOELayer capacity (moe_layer.py:329-337): for experts with count > capacity keep the `capacity`
 * largest combine weights (ties: lower token first); others get w_eff = 0 and keep[...] = 0.        */
int b200_moe_capacity(const int32_t* idx, const float* w, const int32_t* counts, const int32_t* pad_off,
                      const int32_t* row_src, int NK, int E, int capacity, float* w_eff, uint8_t* keep,
                      void* stream);
/* xp[r,:] = x[row_src[r] / K, :] (zeros for padding rows).  128-bit vectorised row gather.           */
int b200_moe_permute(const void* x, const int32_t* row_src, const int32_t* pad_off, int E, int K, int Rmax,
                     int D, int dtype, void* xp, void* stream);
/* dx[n,:] = sum_k dxp[dest_row[n,k], :] (+ add[n,:] if add != NULL)   — backward of permute.        */
int b200_moe_unpermute(const void* dxp, const int32_t* dest_row, const void* add, int N, int K, int D,
                       int dtype, void* dx, void* stream);
/* out[n,:] = LN_out( sum_k w[n,k] * z[dest_row[n,k], :] )   (moe_layer.py:163-171).  gamma == NULL: the plain
 * weighted sum without LayerNorm (HierarchicalMOE, moe_layer.py:489-543; mean / rstd / dgamma / dbeta unused).   */
int b200_moe_combine_fwd(const void* z, const int32_t* dest_row, const float* w, const float* gamma,
                         const float* beta, float eps, int N, int K, int D, int dtype, void* out,
                         float* mean, float* rstd, void* stream);
/* dz[dest_row[n,k],:] = w[n,k] * ds[n,:], d_w[n,k] = <ds[n,:], z[dest_row[n,k],:]>, padding rows of dz
 * zeroed; dgamma/dbeta [D] fp32.  workspace >= b200_moe_combine_bwd_ws(N, D).                       */
size_t b200_moe_combine_bwd_ws(int N, int D);
int b200_moe_combine_bwd(const void* dout, const void* z, const int32_t* dest_row, const float* w,
                         const float* mean, const float* rstd, const float* gamma,
                         const int32_t* row_src, int N, int K, int D, int Rmax, int dtype, void* dz,
                         float* d_w, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- expert parallelism over NVLink peer memory ------------------------------------------------------
 * Dispatch / return fused with their collective (north star: "experts sharded expert-parallel ... all-to-all over
 * NVLink"; the reference only has the placeholder moe_utils.py:194-254): token rows are written straight into the
 * owner rank's grouped-GEMM input through peer-mapped pointers (symmetric allocations; `peer_*` are DEVICE arrays
 * of W device pointers, one per rank).  No host-side split sizes, no staging buffer, no second permute; phases are
 * separated by a stream-ordered cross-rank barrier owned by the caller:
 *   push counts | barrier | layout, dispatch | barrier | grouped FFN, return | barrier | combine.
 *   b200_ep_push_counts : counts[E] (pairs per GLOBAL expert on this rank) -> row `me` of every peer's table [W,E]
 *   b200_ep_layout      : table -> send_base[E] (row in owner(e)'s padded buffer where my rows for expert e start),
 *                         pad_off2[2*El+1] (my padded segment offsets [El+1], then routed rows per local expert [El]),
 *                         tile_group2[Rcap/128] (local expert per tile, -1 unused), row_home[Rcap]
 *                         (home_rank * nk_cap + compact position at home of my padded row, -1 padding).
 *                         Segment order = (expert, source rank, token) = the canonical order of the unsharded layer on
 *                         the concatenated batch.
 *   b200_ep_dispatch    : compact (expert-sorted) row r of `src` (src[cmp_src[r]/K] when cmp_src is given, else src[r])
 *                         -> peer_bufs[owner(e)][send_base[e] + r - cmp_off[e]]; also zeroes my own padding rows
 *   b200_ep_return      : my padded row i -> peer_rets[home_rank][home position]   (rows_hint sizes the grid)      */
int b200_ep_push_counts(const int32_t* counts, void* const* peer_tabs, int me, int W, int E, void* stream);
int b200_ep_layout(const int32_t* tab, int me, int W, int E, int Rcap, int nk_cap, int32_t* send_base,
                   int32_t* pad_off2, int32_t* tile_group2, int32_t* row_home, void* stream);
int b200_ep_dispatch(const void* src, const int32_t* cmp_src, const int32_t* cmp_off, const int32_t* send_base,
                     const int32_t* pad_off2, void* const* peer_bufs, int me, int K, int NK, int E, int El, int D,
                     int Rcap, int dtype, void* stream);
int b200_ep_return(const void* rows, const int32_t* row_home, const int32_t* pad_off2, void* const* peer_rets,
                   int El, int D, int Rcap, int nk_cap, int rows_hint, int dtype, void* stream);
/* Data-parallel gradient all-reduce over peer memory (SURVEY 8(e) collective 3, hand-rolled).
 * b200_p2p_allreduce_f32: peer_bufs is a HOST array of the W ranks' device addresses of one symmetric fp32 buffer;
 * elements [offset, offset+count) are replaced on every rank by scale * (sum over ranks), summed in rank order
 * (two-shot: every rank reduces 1/W of the range with peer loads and stores the result to every peer).
 * b200_nvls_allreduce_f32: the same through the NVSwitch: multicast_ptr is the MULTICAST mapping of the buffer;
 * multimem.ld_reduce sums the W copies inside the switch, multimem.st broadcasts the result (max_blocks <= 0: 64).
 * The caller places a cross-rank barrier on the stream before and after either call.                          */
int b200_p2p_allreduce_f32(const unsigned long long* peer_bufs, int me, int W, long long offset, long long count,
                           float scale, void* stream);
int b200_nvls_allreduce_f32(void* multicast_ptr, int me, int W, long long offset, long long count, float scale,
                            int max_blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VQA_H_ */
