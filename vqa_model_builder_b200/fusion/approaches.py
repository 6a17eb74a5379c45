"""`src/modeling/fusion` API of the reference: BaseFusion, CrossAttentionFusion (+ CrossAttentionBlock), QFormerFusion
(+ QFormerLayer), SingleStreamFusion and the create_fusion_model registry (fusion_approaches.py:16-734).  All three
fusions are wiring over the same kernels: tensor-core GEMMs with fused epilogues, fused attention, add + LayerNorm."""
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


class QFormerLayer(nn.Module):
    """fusion_approaches.py:402-513: queries self-attend, then cross-attend to the image tokens, then to the text
    tokens; every attention is followed by post-LN, an FFN and another post-LN.  Masks are True = VALID."""

    def __init__(self, dim: int, num_heads: int, intermediate_dim: int, dropout: float = 0.1):
        super().__init__()

        def mha():
            return nn.MultiheadAttention(embed_dim=dim, num_heads=num_heads, dropout=dropout, batch_first=True)

        def mlp():
            return nn.Sequential(nn.Linear(dim, intermediate_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(intermediate_dim, dim), nn.Dropout(dropout))

        self.self_attention = mha()
        self.self_norm1 = nn.LayerNorm(dim)
        self.self_norm2 = nn.LayerNorm(dim)
        self.self_ffn = mlp()
        self.vision_cross_attention = mha()
        self.vision_norm1 = nn.LayerNorm(dim)
        self.vision_norm2 = nn.LayerNorm(dim)
        self.vision_ffn = mlp()
        self.text_cross_attention = mha()
        self.text_norm1 = nn.LayerNorm(dim)
        self.text_norm2 = nn.LayerNorm(dim)
        self.text_ffn = mlp()

    SITES = 9   # per stage: attention probabilities, ffn inner, ffn output

    def _block(self, q2, v2, t2, B, Q, V, T, vpad_u8, tpad_u8, slab, dc, k0=0):
        a, r = blocks.self_attention(q2, B, Q, self.self_attention, slab, None, drop_attn=dc.site(k0 + 0),
                                     passthrough=True)
        q2 = blocks.add_ln(r, a, self.self_norm1)
        f, r = blocks.ffn(q2, self.self_ffn[0], self.self_ffn[3], slab, drop_in=dc.site(k0 + 1), passthrough=True)
        q2 = blocks.add_ln(r, f, self.self_norm2, dc.site(k0 + 2))
        a, r = blocks.cross_attention(q2, v2, B, Q, V, self.vision_cross_attention, slab, vpad_u8,
                                      drop_attn=dc.site(k0 + 3), passthrough=True)
        q2 = blocks.add_ln(r, a, self.vision_norm1)
        f, r = blocks.ffn(q2, self.vision_ffn[0], self.vision_ffn[3], slab, drop_in=dc.site(k0 + 4), passthrough=True)
        q2 = blocks.add_ln(r, f, self.vision_norm2, dc.site(k0 + 5))
        a, r = blocks.cross_attention(q2, t2, B, Q, T, self.text_cross_attention, slab, tpad_u8,
                                      drop_attn=dc.site(k0 + 6), passthrough=True)
        q2 = blocks.add_ln(r, a, self.text_norm1)
        f, r = blocks.ffn(q2, self.text_ffn[0], self.text_ffn[3], slab, drop_in=dc.site(k0 + 7), passthrough=True)
        return blocks.add_ln(r, f, self.text_norm2, dc.site(k0 + 8))


class QFormerFusion(SlabOwner, BaseFusion):
    """fusion_approaches.py:284-399 (BLIP-2 style): learnable query tokens read both modalities; output = mean over the
    queries of Linear(LayerNorm(queries))."""

    def __init__(self, vision_dim: int = 768, text_dim: int = 768, output_dim: int = 768, num_query_tokens: int = 32,
                 num_attention_heads: int = 8, num_layers: int = 6, intermediate_dim: int = 3072, dropout: float = 0.1):
        BaseFusion.__init__(self, vision_dim, text_dim, output_dim)
        self.num_query_tokens = num_query_tokens
        self.num_layers = num_layers
        self.query_tokens = nn.Parameter(torch.randn(1, num_query_tokens, output_dim))
        nn.init.trunc_normal_(self.query_tokens, std=0.02)
        self.vision_projection = nn.Linear(vision_dim, output_dim)
        self.text_projection = nn.Linear(text_dim, output_dim)
        self.qformer_layers = nn.ModuleList([
            QFormerLayer(dim=output_dim, num_heads=num_attention_heads, intermediate_dim=intermediate_dim,
                         dropout=dropout) for _ in range(num_layers)])
        self.output_projection = nn.Sequential(nn.LayerNorm(output_dim), nn.Linear(output_dim, output_dim))
        self.dropout_p = float(dropout)
        self._sites = alloc_sites(QFormerLayer.SITES * num_layers)

    def _slab_groups(self):
        return blocks.param_groups(self)

    def forward(self, vision_features: torch.Tensor, text_features: torch.Tensor,
                vision_mask: Optional[torch.Tensor] = None, text_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, V, _ = vision_features.shape
        T = text_features.shape[1]
        D, Q = self.output_dim, self.num_query_tokens
        cdt = resolve_compute_dtype(text_features)
        slab = self._get_slab(text_features.device, cdt)
        v2 = blocks.linear(ops.to_compute(vision_features.reshape(B * V, -1), cdt), self.vision_projection, slab)
        t2 = blocks.linear(ops.to_compute(text_features.reshape(B * T, -1), cdt), self.text_projection, slab)
        q2 = ops.to_compute(self.query_tokens.expand(B, -1, -1).reshape(B * Q, D), cdt)
        vpad = blocks.pad_mask_u8(~vision_mask.bool()) if vision_mask is not None else None
        tpad = blocks.pad_mask_u8(~text_mask.bool()) if text_mask is not None else None
        dc = DropCtx(self.training, self.dropout_p, text_features.device, self._sites)
        for li, layer in enumerate(self.qformer_layers):
            q2 = layer._block(q2, v2, t2, B, Q, V, T, vpad, tpad, slab, dc, li * QFormerLayer.SITES)
        q2 = blocks.add_ln(q2, None, self.output_projection[0])
        q2 = blocks.linear(q2, self.output_projection[1], slab)
        out = q2.view(B, Q, D).float().mean(dim=1)
        return out.to(text_features.dtype)


class SingleStreamFusion(SlabOwner, BaseFusion):
    """fusion_approaches.py:516-677 (ViLT style): [CLS] + image tokens + text tokens (+ modality and position
    embeddings) through pre-LN transformer encoder layers; output = LayerNorm(CLS).  Masks are True = VALID."""

    def __init__(self, vision_dim: int = 768, text_dim: int = 768, output_dim: int = 768, num_attention_heads: int = 12,
                 num_layers: int = 6, intermediate_dim: int = 3072, dropout: float = 0.1, max_vision_tokens: int = 200,
                 max_text_tokens: int = 128):
        BaseFusion.__init__(self, vision_dim, text_dim, output_dim)
        self.num_layers = num_layers
        self.max_vision_tokens = max_vision_tokens
        self.max_text_tokens = max_text_tokens
        self.vision_projection = nn.Linear(vision_dim, output_dim)
        self.text_projection = nn.Linear(text_dim, output_dim)
        self.modality_embeddings = nn.Embedding(2, output_dim)
        self.position_embeddings = nn.Parameter(torch.randn(1, max_vision_tokens + max_text_tokens, output_dim))
        nn.init.trunc_normal_(self.position_embeddings, std=0.02)
        layer = nn.TransformerEncoderLayer(d_model=output_dim, nhead=num_attention_heads,
                                           dim_feedforward=intermediate_dim, dropout=dropout, activation="gelu",
                                           batch_first=True, norm_first=True)
        self.transformer = nn.TransformerEncoder(layer, num_layers=num_layers)    # parameter container (same keys)
        self.norm = nn.LayerNorm(output_dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, output_dim))
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        self.dropout_p = float(dropout)
        self._sites = alloc_sites(4 * num_layers)

    def _slab_groups(self):
        return blocks.param_groups(self)

    def forward(self, vision_features: torch.Tensor, text_features: torch.Tensor,
                vision_mask: Optional[torch.Tensor] = None, text_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, V, _ = vision_features.shape
        T = text_features.shape[1]
        D = self.output_dim
        S = 1 + V + T
        cdt = resolve_compute_dtype(text_features)
        dev = text_features.device
        slab = self._get_slab(dev, cdt)
        v2 = blocks.linear(ops.to_compute(vision_features.reshape(B * V, -1), cdt), self.vision_projection, slab)
        t2 = blocks.linear(ops.to_compute(text_features.reshape(B * T, -1), cdt), self.text_projection, slab)
        emb = self.modality_embeddings.weight
        pos = self.position_embeddings[0, :S]
        # token assembly (elementwise, once per forward): type + position embeddings, [CLS] first
        vis = v2.view(B, V, D).float() + (emb[0] + pos[1:1 + V])
        txt = t2.view(B, T, D).float() + (emb[1] + pos[1 + V:])
        cls = (self.cls_token[0] + pos[:1]).expand(B, 1, D)
        x2 = ops.to_compute(torch.cat([cls, vis, txt], dim=1).reshape(B * S, D).contiguous(), cdt)
        pad = None
        if vision_mask is not None or text_mask is not None:
            vm = vision_mask.bool() if vision_mask is not None else torch.ones(B, V, dtype=torch.bool, device=dev)
            tm = text_mask.bool() if text_mask is not None else torch.ones(B, T, dtype=torch.bool, device=dev)
            pad = blocks.pad_mask_u8(~torch.cat([torch.ones(B, 1, dtype=torch.bool, device=dev), vm, tm], dim=1))
        dc = DropCtx(self.training, self.dropout_p, dev, self._sites)
        for li, layer in enumerate(self.transformer.layers):   # pre-LN: x += drop(SA(LN1 x)); x += drop(FF(LN2 x))
            h = blocks.add_ln(x2, None, layer.norm1)
            x2 = blocks.self_attention(h, B, S, layer.self_attn, slab, pad, residual=x2, drop_attn=dc.site(4 * li),
                                       drop_out=dc.site(4 * li + 1))
            h = blocks.add_ln(x2, None, layer.norm2)
            x2 = blocks.ffn(h, layer.linear1, layer.linear2, slab, residual=x2, drop_in=dc.site(4 * li + 2),
                            drop_out=dc.site(4 * li + 3))
        cls_rows = x2.view(B, S, D)[:, 0, :].contiguous()        # only the CLS row is normalised and returned
        out = blocks.add_ln(cls_rows, None, self.norm)
        return ops.to_compute(out, text_features.dtype)


_FUSIONS = {"cross_attention": CrossAttentionFusion, "qformer": QFormerFusion, "q_former": QFormerFusion,
            "single_stream": SingleStreamFusion, "vilt": SingleStreamFusion}


def create_fusion_model(fusion_type: str, **kwargs) -> BaseFusion:
    """fusion_approaches.py:681-734."""
    if fusion_type not in _FUSIONS:
        raise ValueError(f"Unknown fusion type: {fusion_type}. Available types: {', '.join(_FUSIONS.keys())}")
    return _FUSIONS[fusion_type](**kwargs)
