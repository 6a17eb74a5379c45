"""The fusion the classification pipeline actually runs (reference: src/modeling/meta_arch/vqa_model.py,
CrossModalAttention :237-311 and MultimodalFusion :314-433)."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
from torch import nn

from .. import ops
from .._lib import ACT_RELU
from ..runtime import DropCtx, SlabOwner, alloc_sites, resolve_compute_dtype
from . import blocks


#: MultimodalFusion consumes only the CLS row of its last layer (vqa_model.py:386-387), so that layer's query
#: projections, output projections, FFN and LayerNorms are evaluated for ONE row per sample; its self-attention still
#: projects keys / values from all T rows and its cross-attention from all image patches (SURVEY 8(a) A2:
#: 1.48 instead of 2.40 GFLOP forward per sample at config 1).  Results for the consumed row are unchanged.
DEAD_ROW_ELIMINATION = os.environ.get("B200VQA_DEAD_ROWS", "1") != "0"


@dataclass
class FusionConfig:
    """vqa_config.py:86-105."""
    fusion_type: str = "cross_attention"
    hidden_dim: int = 512
    output_dim: int = 512
    num_heads: int = 8
    num_layers: int = 2
    dropout: float = 0.1
    use_layer_norm: bool = True


class CrossModalAttention(SlabOwner, nn.Module):
    """Post-LN block: x = LN1(q + SelfAttn(q)); x = LN2(x + CrossAttn(x, kv)); x = LN3(x + FFN(x))."""

    def __init__(self, embed_dim: int, num_heads: int = 8, dropout: float = 0.1):
        nn.Module.__init__(self)
        self.self_attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout, batch_first=True)
        self.cross_attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout, batch_first=True)
        self.ffn = nn.Sequential(nn.Linear(embed_dim, embed_dim * 4), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(embed_dim * 4, embed_dim), nn.Dropout(dropout))
        self.norm1 = nn.LayerNorm(embed_dim)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.norm3 = nn.LayerNorm(embed_dim)
        self.dropout = nn.Dropout(dropout)
        self.dropout_p = float(dropout)
        self._sites = alloc_sites(self.SITES)

    def _slab_groups(self):
        return blocks.param_groups(self)

    SITES = 6   # dropout sites: attn probs (self), attn out, attn probs (cross), cross out, ffn inner, ffn out

    def _block(self, x2, kv2, B, T, S, qmask_u8, kvmask_u8, slab, dc: DropCtx, k0: int = 0):
        # each branch returns an alias of its input for the residual connection (gradient sum fused into its dgrad)
        a, xr = blocks.self_attention(x2, B, T, self.self_attn, slab, qmask_u8, drop_attn=dc.site(k0 + 0),
                                      passthrough=True)
        x2 = blocks.add_ln(xr, a, self.norm1, dc.site(k0 + 1))
        c, xr = blocks.cross_attention(x2, kv2, B, T, S, self.cross_attn, slab, kvmask_u8, drop_attn=dc.site(k0 + 2),
                                       passthrough=True)
        x2 = blocks.add_ln(xr, c, self.norm2, dc.site(k0 + 3))
        f, xr = blocks.ffn(x2, self.ffn[0], self.ffn[3], slab, drop_in=dc.site(k0 + 4), passthrough=True)
        return blocks.add_ln(xr, f, self.norm3, dc.site(k0 + 5))

    def _block_cls(self, x2, kv2, B, T, S, qmask_u8, kvmask_u8, slab, dc: DropCtx, k0: int = 0):
        """The block for the CLS row only: returns row 0 of every sample's block output, [B, D].  Queries come from
        that row; the self-attention keys / values from all T rows of x2, the cross-attention ones from kv2."""
        D = x2.shape[1]
        x0 = x2.view(B, T, D)[:, 0, :].contiguous()
        a, xr = blocks.cross_attention(x0, x2, B, 1, T, self.self_attn, slab, qmask_u8, drop_attn=dc.site(k0 + 0),
                                       passthrough=True)
        x0 = blocks.add_ln(xr, a, self.norm1, dc.site(k0 + 1))
        c, xr = blocks.cross_attention(x0, kv2, B, 1, S, self.cross_attn, slab, kvmask_u8, drop_attn=dc.site(k0 + 2),
                                       passthrough=True)
        x0 = blocks.add_ln(xr, c, self.norm2, dc.site(k0 + 3))
        f, xr = blocks.ffn(x0, self.ffn[0], self.ffn[3], slab, drop_in=dc.site(k0 + 4), passthrough=True)
        return blocks.add_ln(xr, f, self.norm3, dc.site(k0 + 5))

    def forward(self, query: torch.Tensor, key_value: torch.Tensor, query_mask: Optional[torch.Tensor] = None,
                kv_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, T, D = query.shape
        S = key_value.shape[1]
        cdt = resolve_compute_dtype(query)
        slab = self._get_slab(query.device, cdt)
        x2 = ops.to_compute(query.reshape(B * T, D), cdt)
        kv2 = ops.to_compute(key_value.reshape(B * S, D), cdt)
        dc = DropCtx(self.training, self.dropout_p, query.device, self._sites)
        out = self._block(x2, kv2, B, T, S, blocks.pad_mask_u8(query_mask), blocks.pad_mask_u8(kv_mask), slab, dc)
        return ops.to_compute(out, query.dtype).view(B, T, D)


class MultimodalFusion(SlabOwner, nn.Module):
    """cross_attention: L CrossModalAttention layers on the text stream, CLS pooling, output_proj, LayerNorm.
    Other fusion_type values keep the reference's (small) branches: concat, bilinear, default add."""

    def __init__(self, config):
        nn.Module.__init__(self)
        self.config = config
        if config.fusion_type == "cross_attention":
            self.fusion_layers = nn.ModuleList([
                CrossModalAttention(config.hidden_dim, config.num_heads, config.dropout)
                for _ in range(config.num_layers)])
            self.output_proj = nn.Linear(config.hidden_dim, config.output_dim)
        elif config.fusion_type == "concat":
            self.fusion_layer = nn.Sequential(nn.Linear(config.hidden_dim * 2, config.hidden_dim), nn.ReLU(),
                                              nn.Dropout(config.dropout),
                                              nn.Linear(config.hidden_dim, config.output_dim))
        elif config.fusion_type == "bilinear":
            self.bilinear = nn.Bilinear(config.hidden_dim, config.hidden_dim, config.output_dim)
        else:
            self.fusion_layer = nn.Linear(config.hidden_dim, config.output_dim)
        self.layer_norm = nn.LayerNorm(config.output_dim) if config.use_layer_norm else None
        self._sites = alloc_sites(CrossModalAttention.SITES * max(1, getattr(config, "num_layers", 1)) + 2)

    def _slab_groups(self):
        return blocks.param_groups(self)

    @staticmethod
    def _pool(t: torch.Tensor) -> torch.Tensor:
        return t[:, 0, :] if t.dim() == 3 else t

    def forward(self, visual_features: torch.Tensor, text_features: torch.Tensor,
                visual_mask: Optional[torch.Tensor] = None, text_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        ft = self.config.fusion_type
        cdt = resolve_compute_dtype(text_features)
        slab = self._get_slab(text_features.device, cdt)
        if ft == "cross_attention":
            B, T, D = text_features.shape
            S = visual_features.shape[1]
            x2 = ops.to_compute(text_features.reshape(B * T, D), cdt)
            kv2 = ops.to_compute(visual_features.reshape(B * S, D), cdt)
            qm, km = blocks.pad_mask_u8(text_mask), blocks.pad_mask_u8(visual_mask)
            dc = DropCtx(self.training, float(self.config.dropout), text_features.device, self._sites)
            last = len(self.fusion_layers) - 1
            cls = None
            for li, layer in enumerate(self.fusion_layers):
                if li == last and DEAD_ROW_ELIMINATION:          # only the CLS row of the last layer is consumed
                    cls = layer._block_cls(x2, kv2, B, T, S, qm, km, slab, dc, li * CrossModalAttention.SITES)
                else:
                    x2 = layer._block(x2, kv2, B, T, S, qm, km, slab, dc, li * CrossModalAttention.SITES)
            if cls is None:
                cls = x2.view(B, T, D)[:, 0, :]                  # CLS position, strided rows (no copy)
            fused = blocks.linear(cls, self.output_proj, slab)
        elif ft == "concat":
            both = torch.cat([self._pool(visual_features), self._pool(text_features)], dim=-1)
            dc = DropCtx(self.training, float(self.config.dropout), text_features.device, self._sites)
            fused = blocks.ffn(ops.to_compute(both.contiguous(), cdt), self.fusion_layer[0], self.fusion_layer[3],
                               slab, act=ACT_RELU, drop_in=dc.site(0))
        elif ft == "bilinear":
            # 768^3-parameter tensor contraction; not on the cross-attention path (left to torch)
            fused = self.bilinear(self._pool(visual_features), self._pool(text_features))
            fused = ops.to_compute(fused.contiguous(), cdt)
        else:
            pooled = self._pool(visual_features) + self._pool(text_features)
            fused = blocks.linear(ops.to_compute(pooled.contiguous(), cdt), self.fusion_layer, slab)
        if self.layer_norm is not None:
            fused = blocks.add_ln(fused, None, self.layer_norm)
        return ops.to_compute(fused, text_features.dtype)
