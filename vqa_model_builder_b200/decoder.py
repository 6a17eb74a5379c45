"""Answer decoder of the generative VQA model.

Reference: src/modeling/meta_arch/generative_vqa_model.py — TransformerDecoder :342-451 (token embedding +
sinusoidal positions, 6 pre-LN nn.TransformerDecoderLayer: causal self-attention, cross-attention to the fused
[B, 114, 768] memory, GELU FFN; final LayerNorm; 64 000-way output projection tied to the embedding),
PositionalEncoding :453-476, the label-smoothed cross-entropy :508-511,585-587.

torch's nn.TransformerDecoder / nn.Embedding / nn.LayerNorm objects are PARAMETER CONTAINERS only (same state_dict
keys and initialisers as the reference); the arithmetic runs on the library's kernels: embedding gather + positions +
dropout in one pass, tensor-core GEMMs with fused bias / GELU / dropout / residual epilogues, tcgen05 attention with a
causal flag, and a streaming cross-entropy over the vocabulary."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from . import ops
from .fusion import blocks
from .runtime import DropCtx, SlabOwner, alloc_sites, resolve_compute_dtype


@dataclass
class DecoderConfig:
    """The subset of GenerativeVQAConfig (generative_vqa_model.py:36-105) that the decoder and its loss read; any
    object with these attributes (e.g. the reference's own config) is accepted."""
    vocab_size: int = 64000
    decoder_hidden_dim: int = 768
    decoder_num_layers: int = 6
    decoder_num_heads: int = 8
    decoder_ff_dim: int = 2048
    decoder_dropout: float = 0.1
    max_answer_length: int = 64
    tie_word_embeddings: bool = True
    label_smoothing: float = 0.1


class PositionalEncoding(nn.Module):
    """generative_vqa_model.py:453-476: buffer `pe` [1, max_len, d_model]; x + pe[:, :T] then dropout.  The decoder's
    forward fuses both into the embedding gather; this module only keeps the buffer (state_dict key pos_encoding.pe)
    and the reference's stand-alone behaviour."""

    def __init__(self, d_model: int, dropout: float = 0.1, max_len: int = 512):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.dropout(x + self.pe[:, :x.size(1)])


class TransformerDecoder(SlabOwner, nn.Module):
    def __init__(self, config, embedding: Optional[nn.Embedding] = None):
        nn.Module.__init__(self)
        self.config = config
        D = config.decoder_hidden_dim
        self.embedding = embedding if embedding is not None else nn.Embedding(config.vocab_size, D)
        self.pos_encoding = PositionalEncoding(D, config.decoder_dropout, max_len=config.max_answer_length)
        layer = nn.TransformerDecoderLayer(d_model=D, nhead=config.decoder_num_heads,
                                           dim_feedforward=config.decoder_ff_dim, dropout=config.decoder_dropout,
                                           activation="gelu", batch_first=True, norm_first=True)
        self.decoder = nn.TransformerDecoder(layer, num_layers=config.decoder_num_layers)   # parameter container
        self.layer_norm = nn.LayerNorm(D)
        self.output_projection = nn.Linear(D, config.vocab_size, bias=False)
        if getattr(config, "tie_word_embeddings", True):
            self.output_projection.weight = self.embedding.weight
        # dropout sites: embedding, then per layer: self-attn probs, sa out, cross-attn probs, ca out, ffn in, ffn out
        self._sites = alloc_sites(1 + 6 * config.decoder_num_layers)

    def _slab_groups(self):
        return blocks.param_groups(self)      # named_parameters() lists a tied weight once

    def forward(self, encoder_hidden_states: torch.Tensor, decoder_input_ids: torch.Tensor,
                encoder_attention_mask: Optional[torch.Tensor] = None,
                decoder_attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B,S,D] memory, [B,T] token ids, masks with 1 = attend -> logits [B,T,V] (in the compute dtype: a fp32 copy
        of the bf16 logits would be 2 GB at B = 128)."""
        B, T = decoder_input_ids.shape
        S, D = encoder_hidden_states.shape[1], encoder_hidden_states.shape[2]
        if T > self.pos_encoding.pe.shape[1]:
            raise ValueError(f"decoder sequence length {T} exceeds max_answer_length {self.pos_encoding.pe.shape[1]}")
        dev = encoder_hidden_states.device
        cdt = resolve_compute_dtype(encoder_hidden_states)
        slab = self._get_slab(dev, cdt)
        dc = DropCtx(self.training, float(self.config.decoder_dropout), dev, self._sites)
        x2 = ops.EmbedFn.apply(decoder_input_ids, self.embedding.weight,
                               slab.compute_view(self.embedding.weight, cdt), self.pos_encoding.pe[0], T, dc.site(0))
        mem2 = ops.to_compute(encoder_hidden_states.reshape(B * S, D), cdt)
        mem_pad = blocks.pad_mask_u8(encoder_attention_mask == 0) if encoder_attention_mask is not None else None
        tgt_pad = blocks.pad_mask_u8(decoder_attention_mask == 0) if decoder_attention_mask is not None else None
        for li, layer in enumerate(self.decoder.layers):
            b = 1 + 6 * li      # pre-LN: x += drop(SA(LN1 x)); x += drop(CA(LN2 x, memory)); x += drop(FF(LN3 x))
            h = blocks.add_ln(x2, None, layer.norm1)
            x2 = blocks.self_attention(h, B, T, layer.self_attn, slab, tgt_pad, residual=x2, drop_attn=dc.site(b),
                                       drop_out=dc.site(b + 1), causal=True)
            h = blocks.add_ln(x2, None, layer.norm2)
            x2 = blocks.cross_attention(h, mem2, B, T, S, layer.multihead_attn, slab, mem_pad, residual=x2,
                                        drop_attn=dc.site(b + 2), drop_out=dc.site(b + 3))
            h = blocks.add_ln(x2, None, layer.norm3)
            x2 = blocks.ffn(h, layer.linear1, layer.linear2, slab, residual=x2, drop_in=dc.site(b + 4),
                            drop_out=dc.site(b + 5))
        h = blocks.add_ln(x2, None, self.layer_norm)
        w = self.output_projection.weight
        logits = ops.LinearFn.apply(h, w, None, slab.compute_view(w, cdt), None)
        return logits.view(B, T, -1)

    def _generate_causal_mask(self, seq_len: int, device: torch.device) -> torch.Tensor:
        """Kept for callers of the reference API (generative_vqa_model.py:448-451); the kernels take a causal flag."""
        mask = torch.triu(torch.ones(seq_len, seq_len, device=device), diagonal=1)
        return mask.masked_fill(mask == 1, float("-inf"))


class FusedCrossEntropyLoss(nn.Module):
    """Drop-in for nn.CrossEntropyLoss(ignore_index=-100, label_smoothing=eps) with mean reduction
    (generative_vqa_model.py:508-511; also the classification loss behind AnswerHead, vqa_model.py:705-713):
    forward(logits [N, C], target [N]) -> scalar.  One streaming pass over the logits each way."""

    def __init__(self, ignore_index: int = -100, label_smoothing: float = 0.0):
        super().__init__()
        self.ignore_index = int(ignore_index)
        self.label_smoothing = float(label_smoothing)

    def forward(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return ops.cross_entropy(logits, target, self.ignore_index, self.label_smoothing)
