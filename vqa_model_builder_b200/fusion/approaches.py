"""`src/modeling/fusion` API of the reference: BaseFusion, CrossAttentionFusion (+ CrossAttentionBlock) and the
create_fusion_model registry (fusion_approaches.py:16-281, 681-734).  QFormer / single-stream fusions are the
next rows of the scope table and are not provided here."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from .. import ops
from ..runtime import DropCtx, SlabOwner, alloc_sites, resolve_compute_dtype
from . import blocks


class BaseFusion(nn.Module):
    """fusion_approaches.py:16-56."""

    def __init__(self, vision_dim: int, text_dim: int, output_dim: int):
        super().__init__()
        self.vision_dim = vision_dim
        self.text_dim = text_dim
        self.output_dim = output_dim

    def get_output_dim(self) -> int:
        return self.output_dim


class CrossAttentionBlock(nn.Module):
    """Bidirectional block: text <- vision cross-attention (+LN, FFN, LN), then vision <- updated text.
    Masks here are True = VALID (inverted before the attention kernels), fusion_approaches.py:262-279."""

    def __init__(self, dim: int, num_heads: int, intermediate_dim: int, dropout: float = 0.1):
        super().__init__()

        def mha():
            return nn.MultiheadAttention(embed_dim=dim, num_heads=num_heads, dropout=dropout, batch_first=True)

        def mlp():
            return nn.Sequential(nn.Linear(dim, intermediate_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(intermediate_dim, dim), nn.Dropout(dropout))

        self.v2t_attention = mha()
        self.v2t_norm1 = nn.LayerNorm(dim)
        self.v2t_norm2 = nn.LayerNorm(dim)
        self.v2t_ffn = mlp()
        self.t2v_attention = mha()
        self.t2v_norm1 = nn.LayerNorm(dim)
        self.t2v_norm2 = nn.LayerNorm(dim)
        self.t2v_ffn = mlp()

    SITES = 6   # per direction: attention probabilities, ffn inner, ffn output

    def _block(self, v2, t2, B, V, T, vpad_u8, tpad_u8, slab, dc, k0=0):
        a, r = blocks.cross_attention(t2, v2, B, T, V, self.v2t_attention, slab, vpad_u8, drop_attn=dc.site(k0 + 0),
                                      passthrough=True)
        t2 = blocks.add_ln(r, a, self.v2t_norm1)
        f, r = blocks.ffn(t2, self.v2t_ffn[0], self.v2t_ffn[3], slab, drop_in=dc.site(k0 + 1), passthrough=True)
        t2 = blocks.add_ln(r, f, self.v2t_norm2, dc.site(k0 + 2))
        a, r = blocks.cross_attention(v2, t2, B, V, T, self.t2v_attention, slab, tpad_u8, drop_attn=dc.site(k0 + 3),
                                      passthrough=True)
        v2 = blocks.add_ln(r, a, self.t2v_norm1)
        f, r = blocks.ffn(v2, self.t2v_ffn[0], self.t2v_ffn[3], slab, drop_in=dc.site(k0 + 4), passthrough=True)
        v2 = blocks.add_ln(r, f, self.t2v_norm2, dc.site(k0 + 5))
        return v2, t2


class CrossAttentionFusion(SlabOwner, BaseFusion):
    """fusion_approaches.py:59-191."""

    def __init__(self, vision_dim: int = 768, text_dim: int = 768, output_dim: int = 768,
                 num_attention_heads: int = 8, num_layers: int = 4, intermediate_dim: int = 3072,
                 dropout: float = 0.1, fusion_method: str = "concat"):
        BaseFusion.__init__(self, vision_dim, text_dim, output_dim)
        self.num_attention_heads = num_attention_heads
        self.num_layers = num_layers
        self.fusion_method = fusion_method
        self.vision_projection = nn.Linear(vision_dim, output_dim) if vision_dim != output_dim else nn.Identity()
        self.text_projection = nn.Linear(text_dim, output_dim) if text_dim != output_dim else nn.Identity()
        self.cross_attention_layers = nn.ModuleList([
            CrossAttentionBlock(dim=output_dim, num_heads=num_attention_heads, intermediate_dim=intermediate_dim,
                                dropout=dropout) for _ in range(num_layers)])
        fin = output_dim * 2 if fusion_method == "concat" else output_dim
        self.fusion_layer = nn.Sequential(nn.Linear(fin, output_dim), nn.LayerNorm(output_dim), nn.GELU(),
                                          nn.Dropout(dropout), nn.Linear(output_dim, output_dim),
                                          nn.LayerNorm(output_dim))
        self.pooling = nn.AdaptiveAvgPool1d(1)
        self.dropout_p = float(dropout)
        self._sites = alloc_sites(CrossAttentionBlock.SITES * num_layers)

    def _slab_groups(self):
        return blocks.param_groups(self)

    def forward(self, vision_features: torch.Tensor, text_features: torch.Tensor,
                vision_mask: Optional[torch.Tensor] = None, text_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, V, _ = vision_features.shape
        T = text_features.shape[1]
        D = self.output_dim
        cdt = resolve_compute_dtype(text_features)
        slab = self._get_slab(text_features.device, cdt)
        v2 = ops.to_compute(vision_features.reshape(B * V, -1), cdt)
        t2 = ops.to_compute(text_features.reshape(B * T, -1), cdt)
        if isinstance(self.vision_projection, nn.Linear):
            v2 = blocks.linear(v2, self.vision_projection, slab)
        if isinstance(self.text_projection, nn.Linear):
            t2 = blocks.linear(t2, self.text_projection, slab)
        vpad = blocks.pad_mask_u8(~vision_mask.bool()) if vision_mask is not None else None
        tpad = blocks.pad_mask_u8(~text_mask.bool()) if text_mask is not None else None
        dc = DropCtx(self.training, self.dropout_p, text_features.device, self._sites)
        for li, layer in enumerate(self.cross_attention_layers):
            v2, t2 = layer._block(v2, t2, B, V, T, vpad, tpad, slab, dc, li * CrossAttentionBlock.SITES)
        # mask-unaware mean pooling over tokens, as in the reference (:172-173)
        vp = v2.view(B, V, D).float().mean(dim=1)
        tp = t2.view(B, T, D).float().mean(dim=1)
        if self.fusion_method == "concat":
            fused = torch.cat([vp, tp], dim=-1)
        elif self.fusion_method == "add":
            fused = vp + tp
        elif self.fusion_method == "multiply":
            fused = vp * tp
        else:
            raise ValueError(f"Unknown fusion method: {self.fusion_method}")
        fl = self.fusion_layer
        h = blocks.linear(ops.to_compute(fused.contiguous(), cdt), fl[0], slab)
        h = blocks.add_ln(h, None, fl[1])
        h = fl[3](torch.nn.functional.gelu(h))     # [B, D] elementwise (+ nn.Dropout) between two LayerNorms
        h = blocks.linear(h, fl[4], slab)
        h = blocks.add_ln(h, None, fl[5])
        return ops.to_compute(h, text_features.dtype)


_FUSIONS = {"cross_attention": CrossAttentionFusion}


def create_fusion_model(fusion_type: str, **kwargs) -> BaseFusion:
    """fusion_approaches.py:681-734 (registry restricted to the fusions implemented natively)."""
    if fusion_type not in _FUSIONS:
        raise ValueError(f"Unknown fusion type: {fusion_type}. Available types: {', '.join(_FUSIONS.keys())}")
    return _FUSIONS[fusion_type](**kwargs)
