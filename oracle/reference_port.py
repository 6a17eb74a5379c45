"""ORACLE — test infrastructure only.  Never imported by the product package.

CPU restatement (plain torch ops on CPU tensors, fp32 or fp64) of the reference's fusion + MOE algorithms,
written at the formula level over a flat state_dict so that it shares no code path with the CUDA product
and none with torch.nn.MultiheadAttention / nn.TransformerEncoderLayer.  Each function cites the reference
lines it restates.  The arithmetic of the reference lives in PyTorch (pinned torch 2.9.1 in poetry.lock;
this image has 2.11.0), so the reference defines no numbers of its own: PARITY IS PINNED by
tests/golden/*.npz, produced by running the reference's own modules in this container
(oracle/make_golden.py), and tests/test_oracle_golden.py checks this file against those vectors.

Gradients come from autograd over these formulas (state_dict tensors with requires_grad=True).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

SD = Dict[str, torch.Tensor]


def drop(x: torch.Tensor, p: float) -> torch.Tensor:
    """nn.Dropout in train mode (p = 0: identity).  Only the CPU baseline timing uses p > 0; parity runs use 0."""
    return torch.nn.functional.dropout(x, p, training=True) if p > 0 else x


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)          # biased variance, as nn.LayerNorm
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))    # nn.GELU() default = exact erf form


def activation(x: torch.Tensor, name: str) -> torch.Tensor:
    if name == "gelu":
        return gelu(x)
    if name == "relu":
        return torch.clamp(x, min=0)
    if name == "silu":
        return x * torch.sigmoid(x)
    if name == "tanh":
        return torch.tanh(x)
    raise ValueError(name)


def mha(sd: SD, p: str, q_in: torch.Tensor, kv_in: torch.Tensor, num_heads: int,
        key_padding_mask: Optional[torch.Tensor], pdrop: float = 0.0) -> torch.Tensor:
    """torch.nn.functional.multi_head_attention_forward semantics as used at vqa_model.py:300,304 and
    fusion_approaches.py:262-277: packed in_proj rows q=[0:D], k=[D:2D], v=[2D:3D]; q scaled by 1/sqrt(dh);
    bool key_padding_mask (True = ignore) -> -inf before softmax; out_proj.  Attention weights are discarded."""
    B, T, D = q_in.shape
    S = kv_in.shape[1]
    W, bias = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = q_in @ W[:D].t() + bias[:D]
    k = kv_in @ W[D:2 * D].t() + bias[D:2 * D]
    v = kv_in @ W[2 * D:].t() + bias[2 * D:]
    dh = D // num_heads
    q = q.view(B, T, num_heads, dh).transpose(1, 2) * (1.0 / math.sqrt(dh))
    k = k.view(B, S, num_heads, dh).transpose(1, 2)
    v = v.view(B, S, num_heads, dh).transpose(1, 2)
    scores = q @ k.transpose(-1, -2)                            # [B,H,T,S]
    if key_padding_mask is not None:
        scores = scores.masked_fill(key_padding_mask.bool()[:, None, None, :], float("-inf"))
    ctx = drop(torch.softmax(scores, dim=-1), pdrop) @ v            # attention-probability dropout
    ctx = ctx.transpose(1, 2).reshape(B, T, D)
    return ctx @ sd[p + "out_proj.weight"].t() + sd[p + "out_proj.bias"]


def ffn(sd: SD, p1: str, p2: str, x: torch.Tensor, act: str = "gelu", pdrop: float = 0.0) -> torch.Tensor:
    h = drop(activation(x @ sd[p1 + "weight"].t() + sd[p1 + "bias"], act), pdrop)
    return drop(h @ sd[p2 + "weight"].t() + sd[p2 + "bias"], pdrop)


# ---- A1: CrossModalAttention (vqa_model.py:279-311), dropout p = 0 -----------------------------------------
def cross_modal_attention(sd: SD, p: str, query, key_value, query_mask, kv_mask, num_heads: int, pdrop: float = 0.0):
    x = query
    x = layer_norm(x + drop(mha(sd, p + "self_attn.", x, x, num_heads, query_mask, pdrop), pdrop),
                   sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    x = layer_norm(x + drop(mha(sd, p + "cross_attn.", x, key_value, num_heads, kv_mask, pdrop), pdrop),
                   sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    x = layer_norm(x + ffn(sd, p + "ffn.0.", p + "ffn.3.", x, pdrop=pdrop), sd[p + "norm3.weight"],
                   sd[p + "norm3.bias"])
    return x


# ---- A2: MultimodalFusion (vqa_model.py:361-433) -------------------------------------------------------------
def multimodal_fusion(sd: SD, fusion_type: str, num_heads: int, num_layers: int, use_layer_norm: bool, visual, text,
                      visual_mask=None, text_mask=None, pdrop: float = 0.0):
    def pool(t):
        return t[:, 0, :] if t.dim() == 3 else t

    if fusion_type == "cross_attention":
        for l in range(num_layers):
            text = cross_modal_attention(sd, f"fusion_layers.{l}.", text, visual, text_mask, visual_mask, num_heads, pdrop)
        fused = text[:, 0, :] @ sd["output_proj.weight"].t() + sd["output_proj.bias"]
    elif fusion_type == "concat":
        both = torch.cat([pool(visual), pool(text)], dim=-1)
        fused = ffn(sd, "fusion_layer.0.", "fusion_layer.3.", both, act="relu")
    elif fusion_type == "bilinear":
        fused = torch.einsum("bi,oij,bj->bo", pool(visual), sd["bilinear.weight"], pool(text)) + sd["bilinear.bias"]
    else:  # 'add' and every unrecognised name ('mcan', 'mutan', 'attention') fall through to this branch
        fused = (pool(visual) + pool(text)) @ sd["fusion_layer.weight"].t() + sd["fusion_layer.bias"]
    if use_layer_norm:
        fused = layer_norm(fused, sd["layer_norm.weight"], sd["layer_norm.bias"])
    return fused


# ---- A3: CrossAttentionFusion / CrossAttentionBlock (fusion_approaches.py:143-188, 243-281) ------------------
def cross_attention_fusion(sd: SD, num_heads: int, num_layers: int, fusion_method: str, vision, text,
                           vision_mask=None, text_mask=None):
    if "vision_projection.weight" in sd:
        vision = vision @ sd["vision_projection.weight"].t() + sd["vision_projection.bias"]
    if "text_projection.weight" in sd:
        text = text @ sd["text_projection.weight"].t() + sd["text_projection.bias"]
    vpad = None if vision_mask is None else ~vision_mask.bool()    # masks here are True = valid
    tpad = None if text_mask is None else ~text_mask.bool()
    for l in range(num_layers):
        p = f"cross_attention_layers.{l}."
        text = layer_norm(text + mha(sd, p + "v2t_attention.", text, vision, num_heads, vpad),
                          sd[p + "v2t_norm1.weight"], sd[p + "v2t_norm1.bias"])
        text = layer_norm(text + ffn(sd, p + "v2t_ffn.0.", p + "v2t_ffn.3.", text), sd[p + "v2t_norm2.weight"],
                          sd[p + "v2t_norm2.bias"])
        vision = layer_norm(vision + mha(sd, p + "t2v_attention.", vision, text, num_heads, tpad),
                            sd[p + "t2v_norm1.weight"], sd[p + "t2v_norm1.bias"])
        vision = layer_norm(vision + ffn(sd, p + "t2v_ffn.0.", p + "t2v_ffn.3.", vision), sd[p + "t2v_norm2.weight"],
                            sd[p + "t2v_norm2.bias"])
    vp, tp = vision.mean(dim=1), text.mean(dim=1)               # mask-unaware means (:172-173)
    if fusion_method == "concat":
        fused = torch.cat([vp, tp], dim=-1)
    elif fusion_method == "add":
        fused = vp + tp
    else:
        fused = vp * tp
    h = fused @ sd["fusion_layer.0.weight"].t() + sd["fusion_layer.0.bias"]
    h = gelu(layer_norm(h, sd["fusion_layer.1.weight"], sd["fusion_layer.1.bias"]))
    h = h @ sd["fusion_layer.4.weight"].t() + sd["fusion_layer.4.bias"]
    return layer_norm(h, sd["fusion_layer.5.weight"], sd["fusion_layer.5.bias"])


# ---- N3: QFormerFusion / QFormerLayer (fusion_approaches.py:284-513) ------------------------------------------------
def qformer_fusion(sd: SD, num_heads: int, num_layers: int, vision, text, vision_mask=None, text_mask=None):
    v = vision @ sd["vision_projection.weight"].t() + sd["vision_projection.bias"]
    t = text @ sd["text_projection.weight"].t() + sd["text_projection.bias"]
    q = sd["query_tokens"].expand(vision.shape[0], -1, -1)
    vpad = None if vision_mask is None else ~vision_mask.bool()
    tpad = None if text_mask is None else ~text_mask.bool()
    for l in range(num_layers):
        p = f"qformer_layers.{l}."
        q = layer_norm(q + mha(sd, p + "self_attention.", q, q, num_heads, None), sd[p + "self_norm1.weight"],
                       sd[p + "self_norm1.bias"])
        q = layer_norm(q + ffn(sd, p + "self_ffn.0.", p + "self_ffn.3.", q), sd[p + "self_norm2.weight"],
                       sd[p + "self_norm2.bias"])
        q = layer_norm(q + mha(sd, p + "vision_cross_attention.", q, v, num_heads, vpad), sd[p + "vision_norm1.weight"],
                       sd[p + "vision_norm1.bias"])
        q = layer_norm(q + ffn(sd, p + "vision_ffn.0.", p + "vision_ffn.3.", q), sd[p + "vision_norm2.weight"],
                       sd[p + "vision_norm2.bias"])
        q = layer_norm(q + mha(sd, p + "text_cross_attention.", q, t, num_heads, tpad), sd[p + "text_norm1.weight"],
                       sd[p + "text_norm1.bias"])
        q = layer_norm(q + ffn(sd, p + "text_ffn.0.", p + "text_ffn.3.", q), sd[p + "text_norm2.weight"],
                       sd[p + "text_norm2.bias"])
    q = layer_norm(q, sd["output_projection.0.weight"], sd["output_projection.0.bias"])
    q = q @ sd["output_projection.1.weight"].t() + sd["output_projection.1.bias"]
    return q.mean(dim=1)


# ---- N3: SingleStreamFusion (fusion_approaches.py:516-677) ----------------------------------------------------------
def single_stream_fusion(sd: SD, num_heads: int, num_layers: int, vision, text, vision_mask=None, text_mask=None):
    B, V, _ = vision.shape
    T = text.shape[1]
    v = vision @ sd["vision_projection.weight"].t() + sd["vision_projection.bias"] + sd["modality_embeddings.weight"][0]
    t = text @ sd["text_projection.weight"].t() + sd["text_projection.bias"] + sd["modality_embeddings.weight"][1]
    x = torch.cat([sd["cls_token"].expand(B, -1, -1), v, t], dim=1)
    x = x + sd["position_embeddings"][:, :x.shape[1], :]
    pad = None
    if vision_mask is not None or text_mask is not None:
        vm = vision_mask.bool() if vision_mask is not None else torch.ones(B, V, dtype=torch.bool)
        tm = text_mask.bool() if text_mask is not None else torch.ones(B, T, dtype=torch.bool)
        pad = ~torch.cat([torch.ones(B, 1, dtype=torch.bool), vm, tm], dim=1)
    for l in range(num_layers):
        x = transformer_encoder_layer_prenorm(sd, f"transformer.layers.{l}.", x, num_heads, pad)
    return layer_norm(x, sd["norm.weight"], sd["norm.bias"])[:, 0, :]


# ---- A5: routers (router.py:105-178, 287-366) -------------------------------------------------------------------
def topk_router(sd: SD, p: str, x: torch.Tensor, top_k: int, lb_weight: float = 0.01,
                noise: Optional[torch.Tensor] = None, noise_std: float = 1.0):
    """Returns (weights [B,S,K], indices [B,S,K] int64, aux-loss scalar, clean probs [B,S,E], clean logits).
    `noise` (N(0,1) draw, [B,S,E]) switches on the NoisyTopKRouter train path."""
    Wg = sd[p + "gate.weight"]
    E = Wg.shape[0]
    clean = x @ Wg.t()
    logits = clean
    if noise is not None:
        scale = torch.nn.functional.softplus(x @ sd[p + "w_noise.weight"].t())
        logits = clean + noise * scale * noise_std
    sel = torch.softmax(logits, dim=-1)
    w, idx = torch.topk(sel, top_k, dim=-1)
    w = w / w.sum(dim=-1, keepdim=True)
    probs = torch.softmax(clean, dim=-1)
    n_tok = x.shape[0] * x.shape[1]
    counts = torch.bincount(idx.reshape(-1), minlength=E).to(probs.dtype)     # (token, slot) pairs per expert
    loss = lb_weight * E * torch.sum((counts / n_tok) * probs.reshape(n_tok, E).mean(dim=0))
    return w, idx, loss, probs, clean


# ---- A7: FeedForwardExpert (expert_types.py:75-92), dropout p = 0 ------------------------------------------------
def feed_forward_expert(sd: SD, p: str, x: torch.Tensor, act: str = "gelu", pdrop: float = 0.0) -> torch.Tensor:
    h = ffn(sd, p + "fc1.", p + "fc2.", x, act, pdrop)
    if x.shape[-1] == h.shape[-1]:
        h = h + x
    return layer_norm(h, sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"])


# ---- N4: GatedLinearExpert (expert_types.py:448-515), dropout p = 0 ------------------------------------------------
def gated_linear_expert(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    h = x @ sd[p + "fc1.weight"].t() + sd[p + "fc1.bias"]
    value, gate = h.chunk(2, dim=-1)
    h = value * torch.sigmoid(gate)
    h = h @ sd[p + "fc2.weight"].t() + sd[p + "fc2.bias"]
    if x.shape[-1] == h.shape[-1]:
        h = h + x
    return layer_norm(h, sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"])


def token_expert(sd: SD, p: str, x: torch.Tensor, kind: str = "feedforward", act: str = "gelu", pdrop: float = 0.0):
    return gated_linear_expert(sd, p, x) if kind == "glu" else feed_forward_expert(sd, p, x, act, pdrop)


# ---- N4: HierarchicalMOE (moe_layer.py:361-548) -------------------------------------------------------------------
def hierarchical_moe(sd: SD, x: torch.Tensor, G: int, Epg: int, Kg: int, Ke: int, kind: str = "feedforward",
                     ys: Optional[torch.Tensor] = None, lb_weight: float = 0.01):
    """The reference's loop structure: for every group slot k and group g that received a token, group g's expert
    router is evaluated (and its aux loss added) and the selected experts' outputs are accumulated with weight
    expert_weight * group_weight under the expert and group masks; then output_proj and output_norm.
    `ys` [G*Epg, B, S, D] supplies the expert outputs for heterogeneous groups (expert bodies as data)."""
    gw, gidx, gloss, gprobs, _ = topk_router(sd, "group_router.", x, Kg, lb_weight)
    out = torch.zeros(x.shape[0], x.shape[1], sd["output_norm.weight"].shape[0], dtype=x.dtype)
    total_aux = gloss
    for k in range(Kg):
        for g in range(G):
            gmask = (gidx[:, :, k] == g)
            if not bool(gmask.any()):
                continue
            ew, eidx, eloss, _, _ = topk_router(sd, f"expert_routers.{g}.", x, Ke, lb_weight)
            total_aux = total_aux + eloss
            gout = torch.zeros_like(out)
            for ek in range(Ke):
                for e in range(Epg):
                    emask = (eidx[:, :, ek] == e)
                    if not bool(emask.any()):
                        continue
                    y = ys[g * Epg + e] if ys is not None else token_expert(sd, f"expert_groups.{g}.{e}.", x, kind)
                    w = ew[:, :, ek] * gw[:, :, k]
                    gout = gout + y * w.unsqueeze(-1) * emask.unsqueeze(-1).to(x.dtype)
            out = out + gout * gmask.unsqueeze(-1).to(x.dtype)
    out = out @ sd["output_proj.weight"].t() + sd["output_proj.bias"]
    return layer_norm(out, sd["output_norm.weight"], sd["output_norm.bias"]), total_aux, gprobs


# ---- A6: MOELayer dense combine (moe_layer.py:146-171) -------------------------------------------------------------
def moe_layer(sd: SD, x: torch.Tensor, num_experts: int, top_k: int, lb_weight: float = 0.01, act: str = "gelu",
              noise=None, noise_std: float = 1.0, weights_indices=None, pdrop: float = 0.0, kind: str = "feedforward"):
    """The reference's algorithm verbatim in structure: every selected expert is evaluated on ALL tokens and masked
    by its routing weight; accumulation in ascending expert order from a zero tensor; output_norm."""
    if weights_indices is None:
        w, idx, loss, probs, _ = topk_router(sd, "router.", x, top_k, lb_weight, noise, noise_std)
    else:
        w, idx = weights_indices
        loss, probs = None, None
    out = torch.zeros(x.shape[0], x.shape[1], sd["output_norm.weight"].shape[0], dtype=x.dtype)
    for e in range(num_experts):
        hit = (idx == e)
        if not bool(hit.any()):
            continue
        we = (w * hit.to(w.dtype)).sum(dim=-1)
        out = out + token_expert(sd, f"experts.{e}.", x, kind, act, pdrop) * we.unsqueeze(-1)
    return layer_norm(out, sd["output_norm.weight"], sd["output_norm.bias"]), loss, probs, w, idx


# ---- A9: VQAMOELayer = MOELayer.forward over heterogeneous experts (moe_layer.py:146-171, 551-692) ----------------
def moe_combine_dense(ys: torch.Tensor, w: torch.Tensor, idx: torch.Tensor, norm_w: torch.Tensor,
                      norm_b: torch.Tensor) -> torch.Tensor:
    """The combine of MOELayer.forward given every expert's output on the whole input: ys [E, B, S, D] (the
    heterogeneous expert bodies are outside the hot path, so they enter as data); experts no token selected are
    skipped; accumulation in ascending expert order; output_norm."""
    E = ys.shape[0]
    out = torch.zeros_like(ys[0])
    for e in range(E):
        hit = (idx == e)
        if not bool(hit.any()):
            continue
        we = (w * hit.to(w.dtype)).sum(dim=-1)
        out = out + ys[e] * we.unsqueeze(-1)
    return layer_norm(out, norm_w, norm_b)


# ---- A8: SparseMOELayer token dispatch with capacity (moe_layer.py:281-352) ---------------------------------------
def sparse_moe_layer(sd: SD, x: torch.Tensor, num_experts: int, top_k: int, capacity_factor: float = 1.25,
                     lb_weight: float = 0.01, act: str = "gelu", noise=None, noise_std: float = 1.0):
    B, S, D = x.shape
    n_tok = B * S
    capacity = int(capacity_factor * n_tok * top_k / num_experts)
    w, idx, loss, probs, _ = topk_router(sd, "router.", x, top_k, lb_weight, noise, noise_std)
    xf, wf, idf = x.reshape(n_tok, D), w.reshape(n_tok, top_k), idx.reshape(n_tok, top_k)
    out = torch.zeros(n_tok, sd["output_norm.weight"].shape[0], dtype=x.dtype)
    for e in range(num_experts):
        hit = (idf == e)
        if not bool(hit.any()):
            continue
        tok = hit.any(dim=-1).nonzero(as_tuple=True)[0]               # ascending token order
        if tok.numel() > capacity:
            ew = (wf * hit.to(wf.dtype)).sum(dim=-1)
            _, keep = torch.topk(ew[tok], capacity)
            tok = tok[keep]
        sel_w = (wf[tok] * hit[tok].to(wf.dtype)).sum(dim=-1)
        y = feed_forward_expert(sd, f"experts.{e}.", xf[tok].unsqueeze(0), act).squeeze(0)
        out = out.index_add(0, tok, y * sel_w.unsqueeze(-1))
    out = layer_norm(out.view(B, S, -1), sd["output_norm.weight"], sd["output_norm.bias"])
    return out, loss, probs, w, idx


# ---- A4: CrossModalFusion (generative_vqa_model.py:286-339) -----------------------------------------------------------
def transformer_encoder_layer_prenorm(sd: SD, p: str, x, num_heads: int, key_padding_mask, pdrop: float = 0.0):
    """nn.TransformerEncoderLayer(norm_first=True, activation='gelu'): x += drop(SA(LN1 x)); x += drop(FF(LN2 x))
    (ffn() applies the inner and the outer dropout of the feed-forward block)."""
    h = layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    x = x + drop(mha(sd, p + "self_attn.", h, h, num_heads, key_padding_mask, pdrop), pdrop)
    h = layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    return x + ffn(sd, p + "linear1.", p + "linear2.", h, pdrop=pdrop)


def cross_modal_fusion(sd: SD, num_heads: int, num_layers: int, visual, question, question_mask=None,
                       moe: Optional[dict] = None, pdrop: float = 0.0):
    """`moe`: None or dict(num_experts, top_k, lb_weight) for the standard MOELayer variant (moe_layer.* keys)."""
    B, V, _ = visual.shape
    fused = torch.cat([visual, question], dim=1)
    pad = None
    if question_mask is not None:
        pad = torch.cat([torch.zeros(B, V, dtype=torch.bool), ~question_mask.bool()], dim=1)
    for l in range(num_layers):
        fused = transformer_encoder_layer_prenorm(sd, f"layers.{l}.", fused, num_heads, pad, pdrop)
    aux = None
    if moe is not None:
        sub = {k[len("moe_layer."):]: v for k, v in sd.items() if k.startswith("moe_layer.")}
        fused, aux, _, _, _ = moe_layer(sub, fused, moe["num_experts"], moe["top_k"], moe.get("lb_weight", 0.01),
                                        pdrop=pdrop)
    return layer_norm(fused, sd["layer_norm.weight"], sd["layer_norm.bias"]), aux


# ---- N2: TransformerDecoder + label-smoothed cross-entropy (generative_vqa_model.py:342-451, 453-476, 508-511,
# 585-587) ------------------------------------------------------------------------------------------------------------
def mha_masked(sd: SD, p: str, q_in, kv_in, num_heads: int, key_padding_mask, causal: bool, pdrop: float = 0.0):
    """mha() plus the additive causal mask of the decoder's self-attention (tgt_mask = triu(-inf, diagonal=1),
    generative_vqa_model.py:404-406,448-451).  The reference passes the padding masks as float 0 / -inf tensors
    (:411-430): added to the scores, which is the same as masked_fill(-inf)."""
    B, T, D = q_in.shape
    S = kv_in.shape[1]
    W, bias = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = q_in @ W[:D].t() + bias[:D]
    k = kv_in @ W[D:2 * D].t() + bias[D:2 * D]
    v = kv_in @ W[2 * D:].t() + bias[2 * D:]
    dh = D // num_heads
    q = q.view(B, T, num_heads, dh).transpose(1, 2) * (1.0 / math.sqrt(dh))
    k = k.view(B, S, num_heads, dh).transpose(1, 2)
    v = v.view(B, S, num_heads, dh).transpose(1, 2)
    scores = q @ k.transpose(-1, -2)
    if causal:
        scores = scores + torch.triu(torch.full((T, S), float("-inf"), dtype=scores.dtype), diagonal=1)
    if key_padding_mask is not None:
        scores = scores.masked_fill(key_padding_mask.bool()[:, None, None, :], float("-inf"))
    ctx = drop(torch.softmax(scores, dim=-1), pdrop) @ v
    ctx = ctx.transpose(1, 2).reshape(B, T, D)
    return ctx @ sd[p + "out_proj.weight"].t() + sd[p + "out_proj.bias"]


def sinusoidal_positions(max_len: int, d_model: int) -> torch.Tensor:
    """PositionalEncoding.pe (generative_vqa_model.py:458-467)."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def transformer_decoder(sd: SD, num_heads: int, num_layers: int, memory, ids, memory_mask=None, tgt_mask=None,
                        pdrop: float = 0.0):
    """TransformerDecoder.forward: embedding + positions (+dropout), L x nn.TransformerDecoderLayer(norm_first=True,
    gelu): x += drop(SA(LN1 x, causal, tgt padding)); x += drop(CA(LN2 x, memory, memory padding));
    x += drop(FF(LN3 x)); final LayerNorm; tied / untied output projection without bias.  Masks: 1 = attend."""
    E = sd["embedding.weight"]
    T = ids.shape[1]
    pe = sd["pos_encoding.pe"][0] if "pos_encoding.pe" in sd else sinusoidal_positions(T, E.shape[1])
    x = drop(E[ids] + pe[:T].to(E.dtype), pdrop)
    mem_pad = (memory_mask == 0) if memory_mask is not None else None
    tgt_pad = (tgt_mask == 0) if tgt_mask is not None else None
    for l in range(num_layers):
        p = f"decoder.layers.{l}."
        h = layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        x = x + drop(mha_masked(sd, p + "self_attn.", h, h, num_heads, tgt_pad, True, pdrop), pdrop)
        h = layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        x = x + drop(mha_masked(sd, p + "multihead_attn.", h, memory, num_heads, mem_pad, False, pdrop), pdrop)
        h = layer_norm(x, sd[p + "norm3.weight"], sd[p + "norm3.bias"])
        x = x + ffn(sd, p + "linear1.", p + "linear2.", h, pdrop=pdrop)
    x = layer_norm(x, sd["layer_norm.weight"], sd["layer_norm.bias"])
    return x @ sd["output_projection.weight"].t()


def smoothed_cross_entropy(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100,
                           smoothing: float = 0.0) -> torch.Tensor:
    """nn.CrossEntropyLoss(ignore_index, label_smoothing), mean over the non-ignored rows:
    loss_n = (1 - eps) * (lse - z_y) + eps * (lse - mean_c z_c)."""
    z = logits.reshape(-1, logits.shape[-1])
    y = labels.reshape(-1)
    valid = y != ignore_index
    lse = torch.logsumexp(z, dim=-1)
    zy = z.gather(1, y.clamp(min=0).unsqueeze(1)).squeeze(1)
    per = (1.0 - smoothing) * (lse - zy) + smoothing * (lse - z.mean(dim=-1))
    return (per * valid).sum() / valid.sum()
