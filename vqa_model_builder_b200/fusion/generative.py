"""Fusion of the generative VQA model (reference: src/modeling/meta_arch/generative_vqa_model.py,
CrossModalFusion :193-339): concat [visual; question] -> L pre-LN transformer encoder layers (joint
self-attention = question <-> image attention) -> optional MOE layer on all tokens -> LayerNorm."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch
from torch import nn

from .. import ops
from ..moe.config import MOEConfig, RouterConfig
from ..moe.layers import MOELayer, SparseMOELayer, VQAMOELayer
from ..runtime import DropCtx, SlabOwner, alloc_sites, resolve_compute_dtype
from . import blocks


@dataclass
class GenerativeFusionConfig:
    """The subset of GenerativeVQAConfig (generative_vqa_model.py:36-88) that CrossModalFusion reads; any
    object with these attributes (e.g. the reference's own config) is accepted."""
    fusion_dim: int = 768
    fusion_num_heads: int = 8
    fusion_num_layers: int = 2
    fusion_dropout: float = 0.1
    decoder_ff_dim: int = 2048
    use_moe: bool = False
    moe_type: str = "standard"
    num_experts: int = 4
    num_experts_per_token: int = 2
    expert_capacity_factor: float = 1.25
    moe_loss_weight: float = 0.01
    moe_position: str = "fusion"
    num_vision_experts: int = 1
    num_text_experts: int = 1
    num_multimodal_experts: int = 1
    num_specialized_experts: int = 1
    vietnamese_optimized: bool = True


class CrossModalFusion(SlabOwner, nn.Module):
    def __init__(self, config):
        nn.Module.__init__(self)
        self.config = config
        self.use_moe = config.use_moe and config.moe_position in ["fusion", "both"]
        self.moe_type = getattr(config, "moe_type", "standard")
        self.layers = nn.ModuleList([
            nn.TransformerEncoderLayer(d_model=config.fusion_dim, nhead=config.fusion_num_heads,
                                       dim_feedforward=config.decoder_ff_dim, dropout=config.fusion_dropout,
                                       activation="gelu", batch_first=True, norm_first=True)
            for _ in range(config.fusion_num_layers)])
        self.moe_layer = None
        self.moe_aux_loss = 0.0
        #: keep the aux loss on the device (no .item() sync); the reference returns a Python float
        self.return_aux_tensor = False
        if self.use_moe:
            self._create_moe_layer(config)
        self.layer_norm = nn.LayerNorm(config.fusion_dim)
        self._sites = alloc_sites(4 * config.fusion_num_layers)

    def _create_moe_layer(self, config):
        if self.moe_type == "vqa":
            self.moe_layer = VQAMOELayer(
                input_dim=config.fusion_dim, hidden_dim=config.decoder_ff_dim, output_dim=config.fusion_dim,
                num_vision_experts=getattr(config, "num_vision_experts", 1),
                num_text_experts=getattr(config, "num_text_experts", 1),
                num_multimodal_experts=getattr(config, "num_multimodal_experts", 1),
                num_specialized_experts=getattr(config, "num_specialized_experts", 1),
                top_k=config.num_experts_per_token, dropout=config.fusion_dropout,
                vietnamese_optimized=getattr(config, "vietnamese_optimized", True))
            return
        if self.moe_type == "sparse":
            # the reference passes config= to SparseMOELayer, which it does not accept (TypeError,
            # generative_vqa_model.py:262); here the same settings are passed by keyword instead
            self.moe_layer = SparseMOELayer(
                input_dim=config.fusion_dim, hidden_dim=config.decoder_ff_dim, output_dim=config.fusion_dim,
                num_experts=config.num_experts, top_k=config.num_experts_per_token,
                capacity_factor=config.expert_capacity_factor, dropout=config.fusion_dropout, use_aux_loss=True)
            return
        rc = RouterConfig(router_type="topk", num_experts=config.num_experts, top_k=config.num_experts_per_token,
                          capacity_factor=config.expert_capacity_factor, load_balance_weight=config.moe_loss_weight,
                          use_aux_loss=True)
        self.moe_layer = MOELayer(config=MOEConfig(
            input_dim=config.fusion_dim, hidden_dim=config.decoder_ff_dim, output_dim=config.fusion_dim,
            num_experts=config.num_experts, num_experts_per_token=config.num_experts_per_token, router_config=rc,
            expert_dropout=config.fusion_dropout))

    def _slab_groups(self):
        groups = []
        for i, layer in enumerate(self.layers):
            groups += blocks.param_groups(layer, f"layers.{i}.")
        groups += blocks.param_groups(self.layer_norm, "layer_norm.")
        return groups

    def forward(self, visual_features: torch.Tensor, question_features: torch.Tensor,
                question_mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Union[float, torch.Tensor]]:
        B, V, D = visual_features.shape
        Tq = question_features.shape[1]
        S = V + Tq
        cdt = resolve_compute_dtype(question_features)
        slab = self._get_slab(question_features.device, cdt)
        fused = torch.cat([visual_features, question_features], dim=1)
        pad = None
        if question_mask is not None:  # visual tokens always attended; question: 1 = attend
            pad = torch.cat([torch.zeros(B, V, dtype=torch.bool, device=fused.device), ~question_mask.bool()], dim=1)
        pad = blocks.pad_mask_u8(pad)
        x2 = ops.to_compute(fused.reshape(B * S, D), cdt)
        dc = DropCtx(self.training, float(self.config.fusion_dropout), fused.device, self._sites)
        for li, layer in enumerate(self.layers):  # pre-LN: x += drop(SA(LN1 x)); x += drop(FF(LN2 x))
            h = blocks.add_ln(x2, None, layer.norm1)
            x2 = blocks.self_attention(h, B, S, layer.self_attn, slab, pad, residual=x2, drop_attn=dc.site(4 * li),
                                       drop_out=dc.site(4 * li + 1))
            h = blocks.add_ln(x2, None, layer.norm2)
            x2 = blocks.ffn(h, layer.linear1, layer.linear2, slab, residual=x2, drop_in=dc.site(4 * li + 2),
                            drop_out=dc.site(4 * li + 3))
        aux: Union[float, torch.Tensor] = 0.0
        if self.moe_layer is not None:
            x3 = self.moe_layer(x2.view(B, S, D))
            aux_t = self.moe_layer.get_aux_loss()
            if self.return_aux_tensor:
                aux = aux_t
            elif isinstance(aux_t, torch.Tensor):
                aux = aux_t.item() if aux_t.numel() == 1 else aux_t.mean().item()
            else:
                aux = float(aux_t) if aux_t else 0.0
            x2 = x3.reshape(B * S, D)
        out = blocks.add_ln(x2, None, self.layer_norm)
        return ops.to_compute(out, question_features.dtype).view(B, S, D), aux
