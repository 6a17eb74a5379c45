"""torch.autograd Functions over the C-ABI kernels.

Each Function is the fwd/bwd pair of one fused stage of the hot path.  Activations are 2-D row-major
[rows, features] tensors in the compute dtype (bf16 or fp32); parameters arrive twice: the fp32
nn.Parameter (so autograd routes the gradient) and its compute-dtype view from the ParamSlab (what the
kernel reads).  Parameter gradients are always fp32.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib
from .runtime import aux_fork, aux_join, aux_on
from ._lib import (ACT_CODES, ACT_NONE, EPI_ACCUM, EPI_ACT, EPI_ACT_D, EPI_ADD, EPI_DACT, EPI_MUL, EPI_NONE, GROUP_TILE, LAYOUT_K,
                   LAYOUT_MN, call, dropout_arg, dtype_code, query, stream_ptr)


_LIB_DTYPES = (torch.float32, torch.bfloat16)


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ---- parameter-gradient buffers ----------------------------------------------------------------------------------
# Every backward below writes the parameter gradients of its stage into ONE flat fp32 buffer.  Under data parallelism
# the buffers come from a gradient arena in symmetric memory (parallel.GradArena): the wgrad GEMMs then write straight
# into the buffer the peer-memory / NVLS all-reduce works on — no staging copy, no re-pointing of .grad.
_ARENA = [None]
_LOCAL = [False]


def set_grad_arena(arena) -> None:
    _ARENA[0] = arena


class local_grads:
    """Context: gradients of Functions whose FORWARD runs inside are rank-local (expert-parallel shards) and must not
    be placed in the all-reduce arena."""

    def __enter__(self):
        self.prev = _LOCAL[0]
        _LOCAL[0] = True

    def __exit__(self, *a):
        _LOCAL[0] = self.prev


def grad_buffer(n: int, device, local: bool = False) -> torch.Tensor:
    arena = _ARENA[0]
    if arena is not None and not local:
        t = arena.take(int(n), device)
        if t is not None:
            return t
    return torch.empty(int(n), dtype=torch.float32, device=device)


def _rows(x: torch.Tensor) -> Tuple[int, int]:
    """(row pitch, dtype code) of a 2-D activation whose last dim is contiguous."""
    assert x.dim() == 2 and x.stride(1) == 1, "activations must be 2-D with a contiguous last dim"
    return x.stride(0), dtype_code(x.dtype)


def cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """dtype conversion with the library's kernel (no autograd).  float16 (the reference trains under fp16 autocast,
    training_pipeline.py:457) is not a compute type of the library: fp16 tensors are converted by torch at the module
    boundary, on entry to the bf16 compute type and on exit back to the caller's dtype."""
    if x.dtype == dtype:
        return x
    _lib.ensure_device(x)
    if x.dtype not in _LIB_DTYPES or dtype not in _LIB_DTYPES:
        return x.to(dtype)
    x = x.contiguous()
    out = torch.empty_like(x, dtype=dtype)
    call("b200_cast", x, dtype_code(x.dtype), out, dtype_code(dtype), x.numel(), stream_ptr())
    return out


class CastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src_dtype = x.dtype
        return cast(x, dtype)

    @staticmethod
    def backward(ctx, g):
        return cast(g, ctx.src_dtype), None


class DropoutFn(torch.autograd.Function):
    """Stand-alone dropout with a regenerated mask (rarely needed: dropout is normally fused into a neighbour)."""

    @staticmethod
    def forward(ctx, x, drop):
        ctx.drop = drop
        return dropout_apply(x, drop)

    @staticmethod
    def backward(ctx, g):
        return dropout_apply(g, ctx.drop), None


def to_compute(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return x if x.dtype == dtype else CastFn.apply(x, dtype)


# ---- raw GEMM helpers (no autograd) ------------------------------------------------------------------
def gemm(a: torch.Tensor, a_layout: int, b: torch.Tensor, b_layout: int, M: int, N: int, K: int, *,
         out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None, bias=None, epi=EPI_NONE,
         act=ACT_NONE, aux_in=None, aux_out=None, drop=None) -> torch.Tensor:
    lda, dt = _rows(a)
    ldb, _ = _rows(b)
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype or a.dtype, device=a.device)
    ld_aux = 0
    for t in (aux_in, aux_out):
        if t is not None:
            ld_aux = t.stride(0)
    call("b200_gemm", a, lda, a_layout, b, ldb, b_layout, out, out.stride(0), M, N, K, dt, dtype_code(out.dtype),
         bias, epi, act, aux_in, aux_out, ld_aux, dropout_arg(drop), stream_ptr())
    return out


def dropout_apply(x: torch.Tensor, drop) -> torch.Tensor:
    """x * keep-mask of a site (element index = flat index of the dense tensor)."""
    x = x.contiguous()
    out = torch.empty_like(x)
    call("b200_dropout_apply", x, out, x.numel(), dtype_code(x.dtype), dropout_arg(drop), stream_ptr())
    return out


def dropout_colsum(x: torch.Tensor, drop, out_sum: torch.Tensor) -> torch.Tensor:
    """(x * keep-mask of the site, column sums of that product written to out_sum) in one pass."""
    x = x.contiguous()
    R, N = x.shape
    out = torch.empty_like(x)
    nb = query("b200_colsum_ws", R, N)
    ws = _ws(nb, x.device)
    call("b200_dropout_colsum", x, out, dtype_code(x.dtype), R, N, dropout_arg(drop), out_sum, ws, nb, stream_ptr())
    return out


def colsum(x: torch.Tensor, tile_group=None, G: int = 1) -> torch.Tensor:
    R, N = x.shape
    assert x.is_contiguous()
    out = torch.empty((G, N), dtype=torch.float32, device=x.device)
    nb = query("b200_colsum_ws", R, N)
    ws = _ws(nb, x.device)
    call("b200_colsum", x, dtype_code(x.dtype), R, N, tile_group, G, out, ws, nb, stream_ptr())
    return out if tile_group is not None else out[0]


# ---- bias gradients handed over by the kernel that produced dy -------------------------------------------------
def _attach_colsum(t: torch.Tensor, cs: torch.Tensor) -> None:
    """add_ln_bwd already summed the columns of the branch gradient it wrote; the Linear that produced the branch
    picks the sums up from the gradient tensor instead of launching a column-sum kernel.  The tensor's version
    counter is stored so an in-place accumulation by autograd invalidates the hand-over."""
    t._b200_colsum = (cs, t._version)


def _take_colsum(t: torch.Tensor, n: int) -> Optional[torch.Tensor]:
    tag = getattr(t, "_b200_colsum", None)
    if tag is None:
        return None
    cs, version = tag
    if version != t._version or cs.numel() != n or cs.device != t.device:
        return None
    return cs


# ---- Linear (+bias, + optional residual) ---------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = x W^T + b, or y = dropout(x W^T + b) + residual.  nn.Linear / MHA in-proj / out-proj
    (vqa_model.py:258-263; TransformerEncoderLayer's x + dropout1(sa(x)))."""

    @staticmethod
    def forward(ctx, x, weight, bias, w_c, residual, drop=None, passthrough=False):
        """passthrough=True additionally returns x itself: a caller that also feeds x to a residual connection uses
        the returned alias there, so both gradients of x arrive in this backward and the sum is fused into the
        dgrad GEMM's epilogue instead of a separate accumulation kernel."""
        _lib.ensure_device(x)
        M, K = x.shape
        N = w_c.shape[0]
        epi = EPI_ADD if residual is not None else EPI_NONE
        if residual is None:
            drop = None
        out = None
        if residual is None and N % 8 != 0:     # e.g. a vocabulary that is not a multiple of 8: pad the row pitch
            out = torch.empty((M, (N + 7) // 8 * 8), dtype=x.dtype, device=x.device)[:, :N]
        y = gemm(x, LAYOUT_K, w_c, LAYOUT_K, M, N, K, out=out, bias=bias, epi=epi, aux_in=residual, drop=drop)
        ctx.save_for_backward(x, w_c)
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.drop = drop
        ctx.passthrough = passthrough
        return (y, x) if passthrough else y

    @staticmethod
    def backward(ctx, dy, dx_alias=None):
        x, w_c = ctx.saved_tensors
        M, K = x.shape
        N = w_c.shape[0]
        pre_db = _take_colsum(dy, N) if ctx.drop is None else None
        dy = _rows_aligned(dy)
        dx = dw = db = None
        side = side2 = None
        g = None
        need_db = False
        epi = EPI_NONE
        if ctx.needs_input_grad[1]:
            need_db = ctx.has_bias and pre_db is None
            flat = grad_buffer(N * K + (N if need_db else 0), x.device)
            dw = flat[:N * K].view(N, K)
            if need_db:
                db = flat[N * K:]
            if ctx.drop is not None:       # gradient of the pre-dropout GEMM output (+ its column sums, same pass)
                g = dropout_colsum(dy, ctx.drop, db) if need_db else dropout_apply(dy, ctx.drop)
            else:
                g = dy
            # split-K accumulates atomically when the [N,K] tile grid cannot fill the machine
            epi = EPI_ACCUM if x.dtype == torch.bfloat16 else EPI_NONE
            if ctx.needs_input_grad[0]:     # fork here (g is ready); the side work is ENQUEUED after the dgrad GEMM
                side = aux_fork(x.device)
                side2 = aux_fork(x.device, 1) if need_db and ctx.drop is None else None
            if ctx.has_bias and pre_db is not None:
                db = pre_db
        if g is None:
            g = dropout_apply(dy, ctx.drop) if ctx.drop is not None else dy
        if dw is None and ctx.has_bias and ctx.needs_input_grad[2]:
            db = pre_db if pre_db is not None else colsum(g)
        # critical path first: when both GEMMs are ready at the same time the one launched first gets the SMs
        if ctx.needs_input_grad[0]:
            if dx_alias is not None:
                dx = gemm(g, LAYOUT_K, w_c, LAYOUT_MN, M, K, N, epi=EPI_ADD, aux_in=_as_rows(dx_alias, x))
            else:
                dx = gemm(g, LAYOUT_K, w_c, LAYOUT_MN, M, K, N)
        if dw is not None:
            with aux_on(side):          # parameter gradients run beside the dgrad GEMM
                gemm(g, LAYOUT_MN, x, LAYOUT_MN, N, K, M, out=dw, epi=epi)
            if need_db and ctx.drop is None:
                with aux_on(side2):
                    _colsum_into(g, db)
        aux_join(side)
        aux_join(side2)
        return dx, dw, db, None, (dy if ctx.has_res else None), None, None


def _rows_aligned(t: torch.Tensor) -> torch.Tensor:
    """A 2-D operand whose rows start on 16-byte boundaries (row pitch padded to 8 elements when needed)."""
    if t.stride(1) == 1 and (t.stride(0) * t.element_size()) % 16 == 0 and t.data_ptr() % 16 == 0:
        return t
    if (t.shape[1] * t.element_size()) % 16 == 0:
        return t.contiguous()
    buf = torch.zeros((t.shape[0], (t.shape[1] + 7) // 8 * 8), dtype=t.dtype, device=t.device)
    buf[:, :t.shape[1]] = t
    return buf[:, :t.shape[1]]


def _as_rows(g: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """gradient of a pass-through alias as a contiguous [rows, features] tensor in the activation dtype"""
    g = g.reshape(like.shape)
    if g.dtype != like.dtype:
        g = g.to(like.dtype)
    return g.contiguous()


def _colsum_into(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    R, N = x.shape
    nb = query("b200_colsum_ws", R, N)
    ws = _ws(nb, x.device)
    call("b200_colsum", x, dtype_code(x.dtype), R, N, None, 1, out, ws, nb, stream_ptr())
    return out


# ---- two-layer FFN with fused activation ----------------------------------------------------------------
class FFNFn(torch.autograd.Function):
    """y = drop_in(act(x W1^T + b1)) W2^T + b2, optionally y = drop_out(...) + residual
    (ffn of vqa_model.py:265-271; TransformerEncoderLayer FF).  GEMM-1's epilogue evaluates the activation AND its
    derivative (x the dropout keep-scale) while the exponential is at hand and stores the latter (`dact`) instead of
    the pre-activation; GEMM-2's dgrad epilogue is then a plain multiply (no erf, no RNG: it is no longer
    epilogue-bound)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w1_c, w2_c, act, residual, drop_in=None, drop_out=None, passthrough=False):
        _lib.ensure_device(x)
        M, D = x.shape
        F = w1_c.shape[0]
        Do = w2_c.shape[0]
        pre = torch.empty((M, F), dtype=x.dtype, device=x.device)       # holds act'(pre) * keep-scale
        h = gemm(x, LAYOUT_K, w1_c, LAYOUT_K, M, F, D, bias=b1, epi=EPI_ACT_D, act=act, aux_out=pre, drop=drop_in)
        epi = EPI_ADD if residual is not None else EPI_NONE
        if residual is None:
            drop_out = None
        y = gemm(h, LAYOUT_K, w2_c, LAYOUT_K, M, Do, F, bias=b2, epi=epi, aux_in=residual, drop=drop_out)
        ctx.save_for_backward(x, pre, h, w1_c, w2_c)
        ctx.act = act
        ctx.has_res = residual is not None
        ctx.drops = (drop_in, drop_out)
        return (y, x) if passthrough else y     # see LinearFn: alias of x for the caller's residual connection

    @staticmethod
    def backward(ctx, dy, dx_alias=None):
        x, pre, h, w1_c, w2_c = ctx.saved_tensors
        drop_in, drop_out = ctx.drops
        M, D = x.shape
        F = w1_c.shape[0]
        Do = w2_c.shape[0]
        pre_db2 = _take_colsum(dy, Do) if drop_out is None else None
        dy = dy.contiguous()
        bf = x.dtype == torch.bfloat16
        wepi = EPI_ACCUM if bf else EPI_NONE
        flat = grad_buffer(F * D + F + Do * F + Do, x.device)
        dw1 = flat[:F * D].view(F, D)
        db1 = flat[F * D:F * D + F]
        dw2 = flat[F * D + F:F * D + F + Do * F].view(Do, F)
        db2 = flat[F * D + F + Do * F:]
        # dropout(dy) and its column sums (= db2) come out of one pass
        g = dropout_colsum(dy, drop_out, db2) if drop_out is not None else dy
        # critical path (current stream): dpre -> dx;  auxiliary streams: dW2 | db2, then (after dpre) dW1 | db1.
        # Forks are taken when the operand is ready, the side work is enqueued AFTER the critical GEMM it runs beside.
        side = aux_fork(x.device)
        side2 = None
        if drop_out is None:
            if pre_db2 is None:
                side2 = aux_fork(x.device, 1)
            else:
                db2 = pre_db2
        dpre = gemm(g, LAYOUT_K, w2_c, LAYOUT_MN, M, F, Do, epi=EPI_MUL, aux_in=pre)
        with aux_on(side):
            gemm(g, LAYOUT_MN, h, LAYOUT_MN, Do, F, M, out=dw2, epi=wepi)
        if drop_out is None and pre_db2 is None:
            with aux_on(side2):
                _colsum_into(g, db2)
        side = aux_fork(x.device)       # dW1 / db1 read dpre
        side2 = aux_fork(x.device, 1)
        dx = None
        if ctx.needs_input_grad[0]:
            if dx_alias is not None:
                dx = gemm(dpre, LAYOUT_K, w1_c, LAYOUT_MN, M, D, F, epi=EPI_ADD, aux_in=_as_rows(dx_alias, x))
            else:
                dx = gemm(dpre, LAYOUT_K, w1_c, LAYOUT_MN, M, D, F)
        with aux_on(side):
            gemm(dpre, LAYOUT_MN, x, LAYOUT_MN, F, D, M, out=dw1, epi=wepi)
        with aux_on(side2):
            _colsum_into(dpre, db1)
        aux_join(side)
        aux_join(side2)
        return dx, dw1, db1, dw2, db2, None, None, None, (dy if ctx.has_res else None), None, None, None


# ---- gated linear unit ---------------------------------------------------------------------------------------------
ACT_GLU = -1     # Python-side marker: the hidden activation of an expert bank is a gated linear unit


def glu_fwd(pre: torch.Tensor, F: int, drop=None, tile_group=None) -> torch.Tensor:
    R = pre.shape[0]
    h = torch.empty((R, F), dtype=pre.dtype, device=pre.device)
    call("b200_glu_fwd", pre, pre.stride(0), h, R, F, dtype_code(pre.dtype), tile_group, dropout_arg(drop), stream_ptr())
    return h


def glu_bwd(dh: torch.Tensor, pre: torch.Tensor, F: int, drop=None, tile_group=None) -> torch.Tensor:
    R = pre.shape[0]
    dpre = torch.empty_like(pre)
    call("b200_glu_bwd", dh.contiguous(), pre, pre.stride(0), dpre, R, F, dtype_code(pre.dtype), tile_group,
         dropout_arg(drop), stream_ptr())
    return dpre


class GLUFn(torch.autograd.Function):
    """h = dropout(value * sigmoid(gate)) with [value | gate] = pre [R, 2F]  (GatedLinearExpert, expert_types.py:497-501)."""

    @staticmethod
    def forward(ctx, pre, drop=None):
        _lib.ensure_device(pre)
        pre = pre.contiguous()
        F = pre.shape[1] // 2
        ctx.save_for_backward(pre)
        ctx.cfg = (F, drop)
        return glu_fwd(pre, F, drop)

    @staticmethod
    def backward(ctx, dh):
        (pre,) = ctx.saved_tensors
        F, drop = ctx.cfg
        return glu_bwd(dh, pre, F, drop), None


# ---- residual add + LayerNorm ------------------------------------------------------------------------------
class AddLNFn(torch.autograd.Function):
    """y = LayerNorm(x + dropout(branch))   (post-LN residual blocks, vqa_model.py:301,305,309); branch may be None."""

    @staticmethod
    def forward(ctx, x, branch, gamma, beta, eps, drop=None):
        _lib.ensure_device(x)
        x = x.contiguous()
        if branch is not None:
            branch = branch.contiguous()
        else:
            drop = None
        R, D = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(R, dtype=torch.float32, device=x.device)
        rstd = torch.empty(R, dtype=torch.float32, device=x.device)
        call("b200_add_ln_fwd", x, branch, gamma, beta, None, float(eps), y, mean, rstd, R, D, dtype_code(x.dtype),
             dropout_arg(drop), 2 if drop is not None else 0, stream_ptr())
        ctx.save_for_backward(x, branch, mean, rstd, gamma)
        ctx.has_branch = branch is not None
        ctx.drop = drop
        return y

    @staticmethod
    def backward(ctx, dy):
        x, branch, mean, rstd, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        R, D = x.shape
        dsum = torch.empty_like(x)
        dbranch = torch.empty_like(x) if ctx.drop is not None else None
        flat = grad_buffer((3 if ctx.has_branch else 2) * D, x.device)
        dcol = flat[2 * D:] if ctx.has_branch else None     # column sums of the branch gradient (its bias gradient)
        nb = query("b200_add_ln_bwd_ws", R, D)
        ws = _ws(nb, x.device)
        call("b200_add_ln_bwd", dy, x, branch, mean, rstd, gamma, None, 1, dsum, flat[:D], flat[D:2 * D], dcol, R, D,
             dtype_code(x.dtype), dropout_arg(ctx.drop), 2 if ctx.drop is not None else 0, dbranch, ws, nb,
             stream_ptr())
        if not ctx.has_branch:
            return dsum, None, flat[:D], flat[D:2 * D], None, None
        dbr = dbranch if dbranch is not None else dsum
        _attach_colsum(dbr, dcol)
        return dsum, dbr, flat[:D], flat[D:2 * D], None, None


# ---- generative decoder glue ------------------------------------------------------------------
class EmbedFn(torch.autograd.Function):
    """x[n] = dropout(E[ids[n]] + pe[n % T])  (nn.Embedding + PositionalEncoding, generative_vqa_model.py:400-402,
    453-476).  `table_c` is the compute-dtype copy of the embedding matrix; the gradient goes to the fp32 master."""

    @staticmethod
    def forward(ctx, ids, table, table_c, pos, T, drop=None):
        _lib.ensure_device(table_c)
        ids = ids.reshape(-1).to(torch.int32).contiguous()
        N = ids.numel()
        V, D = table_c.shape
        out = torch.empty((N, D), dtype=table_c.dtype, device=table_c.device)
        call("b200_embed_fwd", ids, table_c, pos.contiguous(), out, N, int(T), D, V, dtype_code(table_c.dtype),
             dropout_arg(drop), stream_ptr())
        ctx.save_for_backward(ids)
        ctx.cfg = (N, D, V, drop)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        N, D, V, drop = ctx.cfg
        dout = dout.contiguous()
        dtable = grad_buffer(V * D, dout.device).view(V, D)
        dtable.zero_()
        call("b200_embed_bwd", ids, dout, dtable, N, D, V, dtype_code(dout.dtype), dropout_arg(drop), stream_ptr())
        return None, dtable, None, None, None, None


class CrossEntropyFn(torch.autograd.Function):
    """nn.CrossEntropyLoss(ignore_index, label_smoothing), mean over the non-ignored rows
    (generative_vqa_model.py:508-511,585-587; vqa_model.py:705-713): one streaming pass over the logits for the loss,
    one for the gradient."""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index, smoothing):
        _lib.ensure_device(logits)
        assert logits.dim() == 2 and logits.stride(1) == 1
        R, C = logits.shape
        labels = labels.reshape(-1).to(torch.int32).contiguous()
        dev = logits.device
        loss_rows = torch.empty(R, dtype=torch.float32, device=dev)
        lse = torch.empty(R, dtype=torch.float32, device=dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)      # loss, n_valid
        call("b200_ce_fwd", logits, logits.stride(0), labels, R, C, int(ignore_index), float(smoothing),
             dtype_code(logits.dtype), loss_rows, lse, out[:1], out[1:], stream_ptr())
        ctx.save_for_backward(logits, labels, lse, out)
        ctx.cfg = (int(ignore_index), float(smoothing))
        return out[0]

    @staticmethod
    def backward(ctx, dloss):
        logits, labels, lse, out = ctx.saved_tensors
        ignore_index, smoothing = ctx.cfg
        R, C = logits.shape
        dl = dloss.reshape(1).to(torch.float32).contiguous()
        dlogits = torch.empty((R, logits.stride(0)), dtype=logits.dtype, device=logits.device)[:, :C]
        call("b200_ce_bwd", logits, logits.stride(0), labels, lse, R, C, ignore_index, smoothing,
             dtype_code(logits.dtype), dl, out[1:], dlogits, dlogits.stride(0), stream_ptr())
        return dlogits, None, None, None


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100,
                  label_smoothing: float = 0.0) -> torch.Tensor:
    """Fused replacement of F.cross_entropy(logits.view(-1, C), labels.view(-1), ignore_index, label_smoothing)."""
    C = logits.shape[-1]
    l2 = logits.reshape(-1, C)
    if l2.dtype == torch.float16:
        l2 = l2.float()
    if l2.stride(1) != 1 or (l2.stride(0) * l2.element_size()) % 16 != 0 or l2.data_ptr() % 16 != 0:
        pitch = (C + 7) // 8 * 8
        buf = torch.zeros((l2.shape[0], pitch), dtype=l2.dtype, device=l2.device)
        buf[:, :C] = l2
        l2 = buf[:, :C]
    return CrossEntropyFn.apply(l2, labels, ignore_index, label_smoothing)


# ---- attention -----------------------------------------------------------------------------------------------
class AttentionFn(torch.autograd.Function):
    """Multi-head attention core on packed projections.
    self-attention : q_src = kv_src = qkv [B*T, 3D]  (q | k | v column blocks)
    cross-attention: q_src = q [B*T, D], kv_src = kv [B*S, 2D] (k | v column blocks)."""

    @staticmethod
    def forward(ctx, q_src, kv_src, key_pad, B, T, S, H, self_attn, drop=None, causal=False):
        _lib.ensure_device(q_src)
        es = q_src.element_size()
        if self_attn:
            D = q_src.shape[1] // 3
            qp, kp, vp = q_src.data_ptr(), q_src.data_ptr() + D * es, q_src.data_ptr() + 2 * D * es
            ldq = ldk = ldv = q_src.stride(0)
        else:
            D = q_src.shape[1]
            qp, kp, vp = q_src.data_ptr(), kv_src.data_ptr(), kv_src.data_ptr() + D * es
            ldq, ldk = q_src.stride(0), kv_src.stride(0)
            ldv = ldk
        dh = D // H
        scale = 1.0 / float(dh) ** 0.5
        o = torch.empty((B * T, D), dtype=q_src.dtype, device=q_src.device)
        lse = torch.empty((B, H, T), dtype=torch.float32, device=q_src.device)
        call("b200_attn_fwd", qp, ldq, kp, ldk, vp, ldv, key_pad, 1 if causal else 0, o, D, lse, B, H, T, S, dh, scale,
             dtype_code(q_src.dtype), dropout_arg(drop), stream_ptr())
        ctx.save_for_backward(q_src, kv_src if not self_attn else None, key_pad, o, lse)
        ctx.dims = (B, T, S, H, D, dh, scale, self_attn)
        ctx.drop = drop
        ctx.causal = 1 if causal else 0
        return o

    @staticmethod
    def backward(ctx, do):
        q_src, kv_src, key_pad, o, lse = ctx.saved_tensors
        B, T, S, H, D, dh, scale, self_attn = ctx.dims
        do = do.contiguous()
        es = q_src.element_size()
        if self_attn:
            dqkv = torch.empty((B * T, 3 * D), dtype=q_src.dtype, device=q_src.device)
            qp, kp, vp = q_src.data_ptr(), q_src.data_ptr() + D * es, q_src.data_ptr() + 2 * D * es
            ldq = ldk = ldv = q_src.stride(0)
            dqp, dkp, dvp = dqkv.data_ptr(), dqkv.data_ptr() + D * es, dqkv.data_ptr() + 2 * D * es
            ldd = 3 * D
            call("b200_attn_bwd", qp, ldq, kp, ldk, vp, ldv, key_pad, ctx.causal, o, D, do, D, lse, dqp, ldd, dkp, ldd,
                 dvp, ldd, B, H, T, S, dh, scale, dtype_code(q_src.dtype), dropout_arg(ctx.drop), stream_ptr())
            return dqkv, None, None, None, None, None, None, None, None, None
        dq = torch.empty((B * T, D), dtype=q_src.dtype, device=q_src.device)
        dkv = torch.empty((B * S, 2 * D), dtype=q_src.dtype, device=q_src.device)
        qp, kp, vp = q_src.data_ptr(), kv_src.data_ptr(), kv_src.data_ptr() + D * es
        call("b200_attn_bwd", qp, q_src.stride(0), kp, kv_src.stride(0), vp, kv_src.stride(0), key_pad, ctx.causal, o, D,
             do, D, lse, dq, D, dkv.data_ptr(), 2 * D, dkv.data_ptr() + D * es, 2 * D, B, H, T, S, dh, scale,
             dtype_code(q_src.dtype), dropout_arg(ctx.drop), stream_ptr())
        return dq, dkv, None, None, None, None, None, None, None, None


# ---- MOE router ----------------------------------------------------------------------------------------------
class RouterFn(torch.autograd.Function):
    """TopK / NoisyTopK routing (router.py:105-178, 287-366).  Returns (weights [N,K] f32, indices [N,K] i32,
    load-balance loss [1] f32, clean probs [N,E] f32, mean noise scale [1] f32, topk_sum, counts).
    With `stats_group` (a torch.distributed group) the load-balance statistics (pairs per expert, summed
    probabilities: 2E floats) are all-reduced so the loss equals the single-process loss on the concatenated batch."""

    @staticmethod
    def forward(ctx, x, w_gate, w_noise, eps, noise_std, lb_weight, K, stats_group=None):
        _lib.ensure_device(x)
        x = x.contiguous()
        N, D = x.shape
        E = w_gate.shape[0]
        dev = x.device
        noisy = eps is not None
        idx = torch.empty((N, K), dtype=torch.int32, device=dev)
        w = torch.empty((N, K), dtype=torch.float32, device=dev)
        topk_sum = torch.empty(N, dtype=torch.float32, device=dev)
        probs = torch.empty((N, E), dtype=torch.float32, device=dev)
        probs_noisy = torch.empty((N, E), dtype=torch.float32, device=dev) if noisy else None
        stats = torch.empty(2 * E, dtype=torch.float32, device=dev)
        counts, psum = stats[:E], stats[E:]
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        nsm = torch.zeros(1, dtype=torch.float32, device=dev)
        nb = query("b200_router_ws", N, E)
        ws = _ws(nb, dev)
        wg = w_gate.detach().contiguous()
        wn = w_noise.detach().contiguous() if noisy else None
        if noisy:
            eps = eps.contiguous()
        call("b200_router_fwd", x, dtype_code(x.dtype), wg, wn, eps, float(noise_std), float(lb_weight), N, D, E, K,
             idx, w, topk_sum, probs, probs_noisy, counts, psum, loss, nsm, ws, nb, stream_ptr())
        counts_bwd = counts
        if stats_group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(stats_group)
            if world > 1:
                local_counts = counts.clone()
                dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=stats_group)
                n_tot = float(N * world)
                loss = (float(lb_weight) * E / (n_tot * n_tot)) * (stats[:E] * stats[E:]).sum().reshape(1)
                # true partial: d loss / d p_local[n,e] = lb*E*C_e/N_tot^2.  Gradients are AVERAGED over ranks
                # afterwards (DDP convention), so each rank scales its share by W: lb*E*C_e/(N_tot*N); the kernel
                # computes lb*E*counts[e]/N^2, hence counts[e] := C_e * N / N_tot.
                counts_bwd = stats[:E] * (float(N) / n_tot)
                counts = local_counts
        ctx.save_for_backward(x, wg, wn, eps, idx, w, topk_sum, probs, probs_noisy, counts_bwd.contiguous())
        ctx.cfg = (float(noise_std), float(lb_weight), N, D, E, K, noisy)
        ctx.mark_non_differentiable(idx, nsm, topk_sum, counts)
        return w, idx, loss, probs, nsm, topk_sum, counts

    @staticmethod
    def backward(ctx, d_w, _d_idx, d_loss, d_probs, _d_nsm, _d_ts, _d_counts):
        x, wg, wn, eps, idx, w, topk_sum, probs, probs_noisy, counts = ctx.saved_tensors
        noise_std, lb_weight, N, D, E, K, noisy = ctx.cfg
        dev = x.device
        dx = torch.empty_like(x)
        flat = grad_buffer((2 if noisy else 1) * E * D, dev)
        dwg = flat[:E * D].view(E, D)
        dwn = flat[E * D:].view(E, D) if noisy else None
        nb = query("b200_router_bwd_ws", N, D, E)
        ws = _ws(nb, dev)
        if d_w is not None:
            d_w = d_w.contiguous().float()
        if d_loss is not None:
            d_loss = d_loss.contiguous().float()
        if d_probs is not None:      # a loss built on aux_outputs['router_probs'] (router.py:140 is differentiable)
            d_probs = d_probs.contiguous().float()
        call("b200_router_bwd", x, dtype_code(x.dtype), wg, wn, eps, noise_std, lb_weight, N, D, E, K, idx, w,
             topk_sum, probs, probs_noisy, counts, d_w, d_loss, d_probs, dx, dwg, dwn, ws, nb, stream_ptr())
        return dx, dwg, dwn, None, None, None, None, None


# ---- MOE dispatch -> grouped expert FFN -> combine ---------------------------------------------------------------
class RoutingPlan:
    """Device-resident maps produced by b200_moe_plan (no host reads)."""

    def __init__(self, idx: torch.Tensor, E: int, with_cmp_src: bool = False):
        _lib.ensure_device(idx)
        idx = idx.contiguous()
        if idx.dtype != torch.int32:
            idx = idx.to(torch.int32)
        self.idx = idx
        self.NK = idx.numel()
        self.E = E
        dev = idx.device
        self.Rmax = query("b200_moe_max_rows", self.NK, E)
        i32 = dict(dtype=torch.int32, device=dev)
        self.counts = torch.empty(E, **i32)
        self.cmp_off = torch.empty(E + 1, **i32)
        self.pad_off = torch.empty(E + 1, **i32)
        self.dest_row = torch.empty(self.NK, **i32)
        self.cmp_pos = torch.empty(self.NK, **i32)
        self.row_src = torch.empty(self.Rmax, **i32)
        self.tile_group = torch.empty(self.Rmax // GROUP_TILE, **i32)
        self.cmp_src = torch.empty(self.NK, **i32) if with_cmp_src else None    # inverse of cmp_pos (EP dispatch)
        nb = query("b200_moe_plan_ws", self.NK, E)
        ws = _ws(nb, dev)
        call("b200_moe_plan", idx, self.NK, E, self.Rmax, self.counts, self.cmp_off, self.pad_off, self.dest_row,
             self.cmp_pos, self.row_src, self.tile_group, self.cmp_src, ws, nb, stream_ptr())

    def apply_capacity(self, w: torch.Tensor, capacity: int) -> Tuple[torch.Tensor, torch.Tensor]:
        w = w.contiguous()
        w_eff = torch.empty_like(w)
        keep = torch.empty(self.NK, dtype=torch.uint8, device=w.device)
        call("b200_moe_capacity", self.idx, w, self.counts, self.pad_off, self.row_src, self.NK, self.E, int(capacity),
             w_eff, keep, stream_ptr())
        return w_eff, keep


class PermuteFn(torch.autograd.Function):
    """xp[r] = x[row_src[r] // K] (zeros on padding rows); backward sums the K copies back per token.
    `row_src`/`offsets`/`dest` describe either the padded layout (pad_off, dest_row) or the compact one
    (cmp_off, cmp_pos) of a RoutingPlan."""

    @staticmethod
    def forward(ctx, x, row_src, offsets, dest, E, K, rows):
        _lib.ensure_device(x)
        x = x.contiguous()
        N, D = x.shape
        xp = torch.empty((rows, D), dtype=x.dtype, device=x.device)
        call("b200_moe_permute", x, row_src, offsets, E, K, rows, D, dtype_code(x.dtype), xp, stream_ptr())
        ctx.save_for_backward(dest)
        ctx.cfg = (N, K, D)
        return xp

    @staticmethod
    def backward(ctx, dxp):
        (dest,) = ctx.saved_tensors
        N, K, D = ctx.cfg
        dxp = dxp.contiguous()
        dx = torch.empty((N, D), dtype=dxp.dtype, device=dxp.device)
        call("b200_moe_unpermute", dxp, dest, None, N, K, D, dtype_code(dxp.dtype), dx, stream_ptr())
        return dx, None, None, None, None, None, None


class GatherRowsFn(torch.autograd.Function):
    """y[i] = z[dest[i]]  (inverse of a K=1 permute; used to return expert outputs to arrival order under
    expert parallelism).  Backward scatters with the matching row_src map (padding rows zero)."""

    @staticmethod
    def forward(ctx, z, dest, row_src, offsets, E):
        _lib.ensure_device(z)
        z = z.contiguous()
        n = dest.numel()
        D = z.shape[1]
        y = torch.empty((n, D), dtype=z.dtype, device=z.device)
        call("b200_moe_unpermute", z, dest, None, n, 1, D, dtype_code(z.dtype), y, stream_ptr())
        ctx.save_for_backward(row_src, offsets)
        ctx.cfg = (E, z.shape[0], D)
        return y

    @staticmethod
    def backward(ctx, dy):
        row_src, offsets = ctx.saved_tensors
        E, rows, D = ctx.cfg
        dy = dy.contiguous()
        dz = torch.empty((rows, D), dtype=dy.dtype, device=dy.device)
        call("b200_moe_permute", dy, row_src, offsets, E, 1, rows, D, dtype_code(dy.dtype), dz, stream_ptr())
        return dz, None, None, None, None


class ExpertFFNFn(torch.autograd.Function):
    """Grouped FeedForwardExpert over permuted rows (expert_types.py:75-92, all experts in two grouped GEMMs):
        z[r] = LN_e( fc2_e(act(fc1_e xp[r])) (+ xp[r]) ),  e = tile_group[r / 128]
    Inputs after `eps`: flattened per-expert fp32 Parameters in the order
        fc1.weight x E, fc1.bias x E, fc2.weight x E, fc2.bias x E, ln.weight x E, ln.bias x E."""

    @staticmethod
    def forward(ctx, xp, tile_group, pad_off, stacks, act, residual, eps, drops, *expert_params):
        _lib.ensure_device(xp)
        drop_in, drop_out = drops if drops is not None else (None, None)
        R, D = xp.shape
        w1s, b1s, w2s, b2s, lng, lnb = stacks  # [E,F1,D] c, [E,F1] f32, [E,Do,F] c, [E,Do] f32, [E,Do] f32 x2
        E, F1 = w1s.shape[0], w1s.shape[1]     # F1 = F, or 2F for gated (GLU) experts: fc1 emits [value | gate]
        Do, F = w2s.shape[1], w2s.shape[2]
        gated = act == ACT_GLU
        dt = dtype_code(xp.dtype)
        st = stream_ptr()
        dev = xp.device
        rows_used = pad_off[E:E + 1]          # rows in use (device): tiles beyond are never visited
        ctx.local = _LOCAL[0]
        pre = torch.empty((R, F1), dtype=xp.dtype, device=dev)
        if gated:
            call("b200_ggemm", xp, D, w1s, LAYOUT_K, pre, F1, R, F1, D, E, tile_group, rows_used, dt, dt, b1s, EPI_NONE,
                 ACT_NONE, None, None, 0, None, st)
            h = glu_fwd(pre, F, drop_in, tile_group)
        else:
            h = torch.empty((R, F), dtype=xp.dtype, device=dev)
            call("b200_ggemm", xp, D, w1s, LAYOUT_K, h, F, R, F, D, E, tile_group, rows_used, dt, dt, b1s, EPI_ACT_D, act,
                 None, pre, F, dropout_arg(drop_in), st)      # `pre` receives act'(pre) * keep-scale (see FFNFn)
        y2 = torch.empty((R, Do), dtype=xp.dtype, device=dev)
        call("b200_ggemm", h, F, w2s, LAYOUT_K, y2, Do, R, Do, F, E, tile_group, rows_used, dt, dt, b2s, EPI_NONE, ACT_NONE, None,
             None, 0, None, st)
        z = torch.empty((R, Do), dtype=xp.dtype, device=dev)
        mean_e = torch.empty(R, dtype=torch.float32, device=dev)
        rstd_e = torch.empty(R, dtype=torch.float32, device=dev)
        call("b200_add_ln_fwd", y2, xp if residual else None, lng, lnb, tile_group, float(eps), z, mean_e, rstd_e, R,
             Do, dt, dropout_arg(drop_out), 1 if drop_out is not None else 0, st)
        ctx.save_for_backward(xp, pre, h, y2, mean_e, rstd_e, w1s, w2s, lng, tile_group, pad_off)
        ctx.cfg = (R, D, E, F, Do, act, residual, F1)
        ctx.drops = (drop_in, drop_out)
        return z

    @staticmethod
    def backward(ctx, dz):
        xp, pre, h, y2, mean_e, rstd_e, w1s, w2s, lng, tile_group, pad_off = ctx.saved_tensors
        R, D, E, F, Do, act, residual, F1 = ctx.cfg
        gated = act == ACT_GLU
        dz = dz.contiguous()
        dev = dz.device
        dt = dtype_code(dz.dtype)
        st = stream_ptr()
        rows_used = pad_off[E:E + 1]
        # one flat fp32 buffer for every parameter gradient of the expert bank (one DP all-reduce bucket)
        sizes = [E * F1 * D, E * F1, E * Do * F, E * Do, E * Do, E * Do]
        flat = grad_buffer(sum(sizes), dev, local=ctx.local)
        offs = [0]
        for sz in sizes:
            offs.append(offs[-1] + sz)
        dw1, db1, dw2, db2, dlng, dlnb = [flat[offs[i]:offs[i + 1]] for i in range(6)]
        drop_in, drop_out = ctx.drops
        dsum = torch.empty((R, Do), dtype=dz.dtype, device=dev)            # grad of (drop(y2) + xp): residual path
        dr = torch.empty((R, Do), dtype=dz.dtype, device=dev) if drop_out is not None else dsum   # grad of y2
        nb = query("b200_add_ln_bwd_ws", R, Do)
        ws = _ws(nb, dev)
        call("b200_add_ln_bwd", dz, y2, xp if residual else None, mean_e, rstd_e, lng, tile_group, E, dsum, dlng, dlnb,
             db2, R, Do, dt, dropout_arg(drop_out), 1 if drop_out is not None else 0,
             dr if drop_out is not None else None, ws, nb, st)      # db2 = per-expert column sums of dr, fused
        nb = query("b200_colsum_ws", R, max(F1, Do))
        ws = _ws(nb, dev)
        # critical path (current stream): dpre -> dxp;  auxiliary streams: dW2, then (after dpre) dW1 | db1.  Forks are
        # taken when the operand is ready; the side work is enqueued AFTER the critical GEMM it runs beside.
        side = aux_fork(dev)
        if gated:
            dh = torch.empty((R, F), dtype=dz.dtype, device=dev)
            call("b200_ggemm", dr, Do, w2s, LAYOUT_MN, dh, F, R, F, Do, E, tile_group, rows_used, dt, dt, None, EPI_NONE,
                 ACT_NONE, None, None, 0, None, st)
            dpre = glu_bwd(dh, pre, F, drop_in, tile_group)
        else:
            dpre = torch.empty((R, F), dtype=dz.dtype, device=dev)
            call("b200_ggemm", dr, Do, w2s, LAYOUT_MN, dpre, F, R, F, Do, E, tile_group, rows_used, dt, dt, None, EPI_MUL,
                 ACT_NONE, pre, None, F, None, st)
        with aux_on(side):
            call("b200_ggemm_wgrad", dr, Do, h, F, dw2, Do, F, R, E, pad_off, dt, stream_ptr())
        side = aux_fork(dev)
        side2 = aux_fork(dev, 1)
        dxp = None
        if ctx.needs_input_grad[0]:
            dxp = torch.empty((R, D), dtype=dz.dtype, device=dev)
            if residual:
                call("b200_ggemm", dpre, F1, w1s, LAYOUT_MN, dxp, D, R, D, F1, E, tile_group, rows_used, dt, dt, None, EPI_ADD,
                     ACT_NONE, dsum, None, Do, None, st)
            else:
                call("b200_ggemm", dpre, F1, w1s, LAYOUT_MN, dxp, D, R, D, F1, E, tile_group, rows_used, dt, dt, None,
                     EPI_NONE, ACT_NONE, None, None, 0, None, st)
        with aux_on(side):
            call("b200_ggemm_wgrad", dpre, F1, xp, D, dw1, F1, D, R, E, pad_off, dt, stream_ptr())
        with aux_on(side2):
            call("b200_colsum", dpre, dt, R, F1, tile_group, E, db1, ws, nb, stream_ptr())
        aux_join(side)
        aux_join(side2)
        # hand every per-expert Parameter its slice of the flat buffer
        grads: List[torch.Tensor] = []
        for buf, shape in ((dw1, (F1, D)), (db1, (F1,)), (dw2, (Do, F)), (db2, (Do,)), (dlng, (Do,)), (dlnb, (Do,))):
            per = buf.view(E, *shape)
            grads.extend(per[e] for e in range(E))
        return (dxp, None, None, None, None, None, None, None, *grads)


class CombineFn(torch.autograd.Function):
    """out[n] = LN_out( sum_k w[n,k] * z[dest[n,k]] )   (moe_layer.py:163-171), dest < 0 entries skipped."""

    @staticmethod
    def forward(ctx, z, w, dest, row_src, out_gamma, out_beta, eps):
        _lib.ensure_device(z)
        z = z.contiguous()
        w = w.contiguous()
        N, K = w.shape
        Do = z.shape[1]
        dev = z.device
        out = torch.empty((N, Do), dtype=z.dtype, device=dev)
        norm = out_gamma is not None        # None: plain weighted sum (HierarchicalMOE normalises after output_proj)
        mean_o = torch.empty(N, dtype=torch.float32, device=dev) if norm else None
        rstd_o = torch.empty(N, dtype=torch.float32, device=dev) if norm else None
        call("b200_moe_combine_fwd", z, dest, w, out_gamma, out_beta, float(eps), N, K, Do, dtype_code(z.dtype), out,
             mean_o, rstd_o, stream_ptr())
        ctx.save_for_backward(z, w, dest, row_src, mean_o, rstd_o, out_gamma)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, w, dest, row_src, mean_o, rstd_o, out_gamma = ctx.saved_tensors
        N, K = w.shape
        R, Do = z.shape
        dout = dout.contiguous()
        dev = dout.device
        dz = torch.empty((R, Do), dtype=dout.dtype, device=dev)
        d_w = torch.empty((N, K), dtype=torch.float32, device=dev)
        norm = out_gamma is not None
        flat = grad_buffer(2 * Do, dev) if norm else None
        nb = query("b200_moe_combine_bwd_ws", N, Do)
        ws = _ws(nb, dev)
        call("b200_moe_combine_bwd", dout, z, dest, w, mean_o, rstd_o, out_gamma, row_src, N, K, Do, R,
             dtype_code(dout.dtype), dz, d_w, flat[:Do] if norm else None, flat[Do:] if norm else None, ws, nb,
             stream_ptr())
        return dz, d_w, None, None, (flat[:Do] if norm else None), (flat[Do:] if norm else None), None


class DenseCombineFn(torch.autograd.Function):
    """Combine for heterogeneous (PyTorch) experts, VQAMOELayer (moe_layer.py:551-692 via :146-171):
    ys [E, N, D] holds every expert's output; out[n] = LN_out(sum_k w[n,k] * ys[idx[n,k], n])."""

    @staticmethod
    def forward(ctx, ys, w, idx, out_gamma, out_beta, eps):
        _lib.ensure_device(ys)
        E, N, D = ys.shape
        K = idx.shape[1]
        ys = ys.contiguous()
        w = w.contiguous().float()
        idx = idx.to(torch.int64)
        n_ar = torch.arange(N, device=ys.device).unsqueeze(1)
        valid = (idx >= 0) & (idx < E)
        dest = torch.where(valid, idx * N + n_ar, torch.full_like(idx, -1)).to(torch.int32).contiguous()
        out = torch.empty((N, D), dtype=ys.dtype, device=ys.device)
        norm = out_gamma is not None
        mean = torch.empty(N, dtype=torch.float32, device=ys.device) if norm else None
        rstd = torch.empty(N, dtype=torch.float32, device=ys.device) if norm else None
        call("b200_moe_combine_fwd", ys, dest, w, out_gamma, out_beta, float(eps), N, K, D, dtype_code(ys.dtype), out,
             mean, rstd, stream_ptr())
        ctx.save_for_backward(ys, dest, w, mean, rstd, out_gamma)
        ctx.cfg = (E, N, K, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        ys, dest, w, mean, rstd, out_gamma = ctx.saved_tensors
        E, N, K, D = ctx.cfg
        dev = dout.device
        dout = dout.contiguous()
        # rows of ys that no token selected get a zero gradient: mark them as "padding" for the kernel
        row_src = torch.full((E * N,), -1, dtype=torch.int32, device=dev)
        flat_dest = dest.view(-1).to(torch.int64)
        ok = flat_dest >= 0
        row_src[flat_dest[ok]] = torch.arange(N * K, device=dev, dtype=torch.int32)[ok]
        dys = torch.empty_like(ys)
        d_w = torch.empty((N, K), dtype=torch.float32, device=dev)
        norm = out_gamma is not None
        flat = grad_buffer(2 * D, dev) if norm else None
        nb = query("b200_moe_combine_bwd_ws", N, D)
        ws = _ws(nb, dev)
        call("b200_moe_combine_bwd", dout, ys, dest, w, mean, rstd, out_gamma, row_src, N, K, D, E * N,
             dtype_code(ys.dtype), dys, d_w, flat[:D] if norm else None, flat[D:] if norm else None, ws, nb,
             stream_ptr())
        return dys, d_w, None, (flat[:D] if norm else None), (flat[D:] if norm else None), None


# ---- cross-attention projections from one packed in_proj ------------------------------------------------------
class CrossProjFn(torch.autograd.Function):
    """nn.MultiheadAttention with key != query (vqa_model.py:304; fusion_approaches.py:262-277) slices its
    packed in_proj_weight [3D, D]: q = x Wq^T + bq (rows 0:D), [k|v] = kv W_kv^T + b_kv (rows D:3D).
    One Function so the packed parameter receives one gradient written in place (no slice/cat kernels)."""

    @staticmethod
    def forward(ctx, x, kv, in_w, in_b, in_w_c, passthrough=False):
        _lib.ensure_device(x)
        M, D = x.shape
        Mk = kv.shape[0]
        bq = in_b[:D] if in_b is not None else None
        bkv = in_b[D:] if in_b is not None else None
        kvp = torch.empty((Mk, 2 * D), dtype=kv.dtype, device=kv.device)
        side = aux_fork(x.device)       # the image-patch projection is independent of the question stream
        with aux_on(side):
            gemm(kv, LAYOUT_K, in_w_c[D:], LAYOUT_K, Mk, 2 * D, D, bias=bkv, out=kvp)
        q = gemm(x, LAYOUT_K, in_w_c[:D], LAYOUT_K, M, D, D, bias=bq)
        aux_join(side)
        ctx.save_for_backward(x, kv, in_w_c)
        ctx.has_bias = in_b is not None
        return (q, kvp, x) if passthrough else (q, kvp)    # see LinearFn: alias of x for the residual connection

    @staticmethod
    def backward(ctx, dq, dkvp, dx_alias=None):
        x, kv, in_w_c = ctx.saved_tensors
        M, D = x.shape
        Mk = kv.shape[0]
        dq = dq.contiguous()
        dkvp = dkvp.contiguous()
        bf = x.dtype == torch.bfloat16
        wepi = EPI_ACCUM if bf else EPI_NONE
        flat = grad_buffer(3 * D * D + 3 * D, x.device)
        dw = flat[:3 * D * D].view(3 * D, D)
        db = flat[3 * D * D:]
        side = aux_fork(x.device)       # parameter gradients beside the two dgrad GEMMs (enqueued after them)
        side2 = aux_fork(x.device, 1)
        dx = None
        if ctx.needs_input_grad[0]:
            if dx_alias is not None:
                dx = gemm(dq, LAYOUT_K, in_w_c[:D], LAYOUT_MN, M, D, D, epi=EPI_ADD, aux_in=_as_rows(dx_alias, x))
            else:
                dx = gemm(dq, LAYOUT_K, in_w_c[:D], LAYOUT_MN, M, D, D)
        dkv = gemm(dkvp, LAYOUT_K, in_w_c[D:], LAYOUT_MN, Mk, D, 2 * D) if ctx.needs_input_grad[1] else None
        with aux_on(side):
            gemm(dq, LAYOUT_MN, x, LAYOUT_MN, D, D, M, out=dw[:D], epi=wepi)
            gemm(dkvp, LAYOUT_MN, kv, LAYOUT_MN, 2 * D, D, Mk, out=dw[D:], epi=wepi)
        with aux_on(side2):
            _colsum_into(dq, db[:D])
            _colsum_into(dkvp, db[D:])
        aux_join(side)
        aux_join(side2)
        return dx, dkv, dw, (db if ctx.has_bias else None), None, None


# ---- multi-layer perceptron with fused activations (answer head) -----------------------------------------------
class MLPFn(torch.autograd.Function):
    """h_{i+1} = dropout_i(act_i(h_i W_i^T + b_i)), last layer without activation (AnswerHead, vqa_model.py:451-477).
    Every activation (+dropout) runs in its GEMM's epilogue; its derivative (x the regenerated mask) in the epilogue
    of the next layer's dgrad GEMM, exactly as in FFNFn.  `params` = (w_0, b_0, wc_0, w_1, b_1, wc_1, ...): fp32
    Parameters for autograd plus the compute-dtype copy the kernels read.  The output row pitch is padded to a
    multiple of 8 elements so any number of classes satisfies the 16-byte operand rule of the tensor-core path; the
    returned logits are the [:, :n_out] view."""

    @staticmethod
    def forward(ctx, x, acts, drops, *params):
        _lib.ensure_device(x)
        L = len(params) // 3
        assert len(acts) == L and acts[-1] == ACT_NONE, "the last layer of the MLP must be linear"
        h = x.contiguous()
        saved = []
        for i in range(L):
            b, wc = params[3 * i + 1], params[3 * i + 2]
            M, K = h.shape
            N = wc.shape[0]
            if i < L - 1:
                pre = torch.empty((M, N), dtype=h.dtype, device=h.device)
                nxt = gemm(h, LAYOUT_K, wc, LAYOUT_K, M, N, K, bias=b, epi=EPI_ACT, act=acts[i], aux_out=pre,
                           drop=drops[i])
                saved += [h, pre]
                h = nxt
            else:
                npad = (N + 7) // 8 * 8
                buf = torch.empty((M, npad), dtype=h.dtype, device=h.device)
                gemm(h, LAYOUT_K, wc, LAYOUT_K, M, N, K, bias=b, out=buf[:, :N])
                saved += [h]
                out = buf[:, :N]
        ctx.save_for_backward(*saved, *[params[3 * i + 2] for i in range(L)])
        ctx.cfg = (L, tuple(acts), tuple(drops), [params[3 * i + 1] is not None for i in range(L)])
        return out

    @staticmethod
    def backward(ctx, dy):
        L, acts, drops, has_bias = ctx.cfg
        saved = ctx.saved_tensors
        wcs = saved[len(saved) - L:]
        hs = [saved[2 * i] for i in range(L - 1)] + [saved[2 * (L - 1)]]
        pres = [saved[2 * i + 1] for i in range(L - 1)]
        dev = dy.device
        bf = hs[0].dtype == torch.bfloat16
        wepi = EPI_ACCUM if bf else EPI_NONE
        M = dy.shape[0]
        n_out = wcs[-1].shape[0]
        npad = (n_out + 7) // 8 * 8
        gbuf = torch.empty((M, npad), dtype=hs[0].dtype, device=dev)
        gbuf[:, :n_out].copy_(dy)
        g = gbuf[:, :n_out]                       # row pitch padded: a legal tensor-core operand for any n_out
        g_dense = dy.to(hs[0].dtype).contiguous()  # for the column sums
        grads = [None] * (3 * L)
        dx = None
        for i in range(L - 1, -1, -1):
            h, wc = hs[i], wcs[i]
            N, K = wc.shape
            flat = grad_buffer(N * K + (N if has_bias[i] else 0), dev)
            dw = flat[:N * K].view(N, K)
            side = aux_fork(dev)                  # parameter gradients beside the dgrad GEMM (enqueued after it)
            side2 = aux_fork(dev, 1) if has_bias[i] else None
            g_w, g_b = g, g_dense
            grads[3 * i] = dw
            if i > 0:      # gradient wrt the previous layer's pre-activation: act' and the dropout mask in the epilogue
                g = gemm(g, LAYOUT_K, wc, LAYOUT_MN, M, K, N, epi=EPI_DACT, act=acts[i - 1], aux_in=pres[i - 1],
                         drop=drops[i - 1])
                g_dense = g
            elif ctx.needs_input_grad[0]:
                dx = gemm(g, LAYOUT_K, wc, LAYOUT_MN, M, K, N)
            with aux_on(side):
                gemm(g_w, LAYOUT_MN, h, LAYOUT_MN, N, K, M, out=dw, epi=wepi)
            if has_bias[i]:
                with aux_on(side2):
                    _colsum_into(g_b, flat[N * K:])
                grads[3 * i + 1] = flat[N * K:]
            aux_join(side)
            aux_join(side2)
        return (dx, None, None, *grads)
