"""Shared building blocks of the fusion modules: multi-head attention (self / cross) and the FFN, expressed
over the library's autograd Functions.  torch's nn.MultiheadAttention / nn.Linear / nn.LayerNorm objects are
used as PARAMETER CONTAINERS only (identical state_dict keys and initialisation to the reference); their
forward() is never called."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import ops
from .._lib import ACT_GELU
from ..slab import ParamSlab


def pad_mask_u8(mask: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """bool/int key-padding mask (True = ignore) -> contiguous uint8 for the attention kernels."""
    if mask is None:
        return None
    return mask.to(torch.uint8).contiguous()


def self_attention(x2, B, T, mha: nn.MultiheadAttention, slab: ParamSlab, key_pad_u8, residual=None, drop_attn=None,
                   drop_out=None, passthrough=False, causal=False):
    """out_proj(softmax(q k^T / sqrt(dh) + mask) v) over x2 [B*T, D]; optional (dropout +) residual fused into
    out_proj.  passthrough=True returns (out, alias of x2): use the alias for the residual connection around this
    branch so the two gradients of x2 are summed in the in-projection's dgrad epilogue."""
    cdt = x2.dtype
    xr = None
    if passthrough and x2.requires_grad:
        qkv, xr = ops.LinearFn.apply(x2, mha.in_proj_weight, mha.in_proj_bias,
                                     slab.compute_view(mha.in_proj_weight, cdt), None, None, True)
    else:
        qkv = ops.LinearFn.apply(x2, mha.in_proj_weight, mha.in_proj_bias,
                                 slab.compute_view(mha.in_proj_weight, cdt), None)
    ctx = ops.AttentionFn.apply(qkv, None, key_pad_u8, B, T, T, mha.num_heads, True, drop_attn, causal)
    out = ops.LinearFn.apply(ctx, mha.out_proj.weight, mha.out_proj.bias,
                             slab.compute_view(mha.out_proj.weight, cdt), residual, drop_out)
    return (out, xr if xr is not None else x2) if passthrough else out


def cross_attention(x2, kv2, B, T, S, mha: nn.MultiheadAttention, slab: ParamSlab, key_pad_u8, residual=None,
                    drop_attn=None, passthrough=False, drop_out=None):
    """Queries from x2 [B*T, D], keys/values from kv2 [B*S, D] (question -> image patches).  passthrough: see
    self_attention."""
    cdt = x2.dtype
    xr = None
    if passthrough and x2.requires_grad:
        q, kvp, xr = ops.CrossProjFn.apply(x2, kv2, mha.in_proj_weight, mha.in_proj_bias,
                                           slab.compute_view(mha.in_proj_weight, cdt), True)
    else:
        q, kvp = ops.CrossProjFn.apply(x2, kv2, mha.in_proj_weight, mha.in_proj_bias,
                                       slab.compute_view(mha.in_proj_weight, cdt))
    ctx = ops.AttentionFn.apply(q, kvp, key_pad_u8, B, T, S, mha.num_heads, False, drop_attn)
    out = ops.LinearFn.apply(ctx, mha.out_proj.weight, mha.out_proj.bias,
                             slab.compute_view(mha.out_proj.weight, cdt), residual, drop_out)
    return (out, xr if xr is not None else x2) if passthrough else out


def ffn(x2, lin1: nn.Linear, lin2: nn.Linear, slab: ParamSlab, act: int = ACT_GELU, residual=None, drop_in=None,
        drop_out=None, passthrough=False):
    cdt = x2.dtype
    if passthrough and x2.requires_grad:
        return ops.FFNFn.apply(x2, lin1.weight, lin1.bias, lin2.weight, lin2.bias,
                               slab.compute_view(lin1.weight, cdt), slab.compute_view(lin2.weight, cdt), act, residual,
                               drop_in, drop_out, True)
    out = ops.FFNFn.apply(x2, lin1.weight, lin1.bias, lin2.weight, lin2.bias, slab.compute_view(lin1.weight, cdt),
                          slab.compute_view(lin2.weight, cdt), act, residual, drop_in, drop_out)
    return (out, x2) if passthrough else out


def add_ln(x2, branch, ln: nn.LayerNorm, drop=None):
    return ops.AddLNFn.apply(x2, branch, ln.weight, ln.bias, ln.eps, drop)


def linear(x2, lin: nn.Linear, slab: ParamSlab, residual=None):
    return ops.LinearFn.apply(x2, lin.weight, lin.bias, slab.compute_view(lin.weight, x2.dtype), residual)


def param_groups(module: nn.Module, prefix: str = ""):
    """One slab group per parameter (no stacking needed for the fusion blocks)."""
    return [[(prefix + n, p)] for n, p in module.named_parameters()]
