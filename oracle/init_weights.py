"""ORACLE — test infrastructure only.  Random-init state_dicts with the reference's key layout, built from plain
torch.nn containers (same constructors, hence the same initialisers, as the reference modules:
vqa_model.py:258-277,331-359 and moe_layer.py:95-117) — used by the CPU baseline so it never touches the product."""
from __future__ import annotations

import torch
from torch import nn


def multimodal_fusion_sd(D: int, H: int, L: int, out_dim: int | None = None) -> dict:
    out_dim = out_dim or D
    sd = {}
    for l in range(L):
        p = f"fusion_layers.{l}."
        for name in ("self_attn", "cross_attn"):
            m = nn.MultiheadAttention(D, H, batch_first=True)
            for k, v in m.state_dict().items():
                sd[p + name + "." + k] = v
        f0, f3 = nn.Linear(D, 4 * D), nn.Linear(4 * D, D)
        for i, m in (("0", f0), ("3", f3)):
            sd[p + f"ffn.{i}.weight"], sd[p + f"ffn.{i}.bias"] = m.weight.detach(), m.bias.detach()
        for n in ("norm1", "norm2", "norm3"):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = torch.ones(D), torch.zeros(D)
    proj = nn.Linear(D, out_dim)
    sd["output_proj.weight"], sd["output_proj.bias"] = proj.weight.detach(), proj.bias.detach()
    sd["layer_norm.weight"], sd["layer_norm.bias"] = torch.ones(out_dim), torch.zeros(out_dim)
    return {k: v.detach().clone() for k, v in sd.items()}


def cross_modal_fusion_sd(D: int, H: int, L: int, F: int) -> dict:
    """CrossModalFusion without its MOE layer (generative_vqa_model.py:196-222): L pre-LN TransformerEncoderLayers
    (state_dict keys of nn.TransformerEncoderLayer) + the final LayerNorm."""
    sd = {}
    for l in range(L):
        p = f"layers.{l}."
        m = nn.MultiheadAttention(D, H, batch_first=True)
        for k, v in m.state_dict().items():
            sd[p + "self_attn." + k] = v
        l1, l2 = nn.Linear(D, F), nn.Linear(F, D)
        sd[p + "linear1.weight"], sd[p + "linear1.bias"] = l1.weight.detach(), l1.bias.detach()
        sd[p + "linear2.weight"], sd[p + "linear2.bias"] = l2.weight.detach(), l2.bias.detach()
        for n in ("norm1", "norm2"):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = torch.ones(D), torch.zeros(D)
    sd["layer_norm.weight"], sd["layer_norm.bias"] = torch.ones(D), torch.zeros(D)
    return {k: v.detach().clone() for k, v in sd.items()}


def moe_layer_sd(D: int, F: int, E: int, noisy: bool = False) -> dict:
    sd = {"router.gate.weight": nn.Linear(D, E, bias=False).weight.detach()}
    if noisy:
        sd["router.w_noise.weight"] = nn.Linear(D, E, bias=False).weight.detach()
    for e in range(E):
        fc1, fc2 = nn.Linear(D, F), nn.Linear(F, D)
        p = f"experts.{e}."
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = fc1.weight.detach(), fc1.bias.detach()
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = fc2.weight.detach(), fc2.bias.detach()
        sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"] = torch.ones(D), torch.zeros(D)
    sd["output_norm.weight"], sd["output_norm.bias"] = torch.ones(D), torch.zeros(D)
    return {k: v.detach().clone() for k, v in sd.items()}


def seeded_state_dict(template: dict, seed: int, keys=None) -> dict:
    """Deterministic weights for the golden vectors / fingerprints: values come from numpy's PCG64 stream (stable
    across torch versions), drawn in the order of `keys` (default: the template's own order).  LayerNorm scales are
    1 + 0.1 n, matrices n / sqrt(fan_in), vectors 0.1 n; integer tensors and 0-d buffers are copied.
    `template` maps names to tensors of the right shape (a module's state_dict())."""
    import numpy as np
    rng = np.random.default_rng(seed)
    sd = {}
    for k in (list(template) if keys is None else list(keys)):
        v = template[k]
        if not v.dtype.is_floating_point or v.dim() == 0:
            sd[k] = v.clone()
            continue
        parts = k.split(".")
        is_norm_w = k.endswith("weight") and (k.endswith("norm.weight") or ".norm" in k or
                                              (len(parts) >= 2 and parts[-2].startswith("norm")))
        if is_norm_w:
            a = 1.0 + 0.1 * rng.standard_normal(tuple(v.shape))
        elif v.dim() >= 2:
            a = rng.standard_normal(tuple(v.shape)) / np.sqrt(v.shape[-1])
        else:
            a = 0.1 * rng.standard_normal(tuple(v.shape))
        sd[k] = torch.tensor(a, dtype=torch.float32)
    return sd


def decoder_sd(D: int, H: int, L: int, F: int, vocab: int, max_len: int) -> dict:
    """TransformerDecoder (generative_vqa_model.py:345-381): embedding tied to the output projection, sinusoidal
    positions, L pre-LN nn.TransformerDecoderLayer key sets, the final LayerNorm."""
    from .reference_port import sinusoidal_positions
    emb = nn.Embedding(vocab, D).weight.detach().clone()
    sd = {"embedding.weight": emb, "output_projection.weight": emb,
          "pos_encoding.pe": sinusoidal_positions(max_len, D).unsqueeze(0)}
    for l in range(L):
        p = f"decoder.layers.{l}."
        for name in ("self_attn", "multihead_attn"):
            m = nn.MultiheadAttention(D, H, batch_first=True)
            for k, v in m.state_dict().items():
                sd[p + name + "." + k] = v.detach().clone()
        l1, l2 = nn.Linear(D, F), nn.Linear(F, D)
        sd[p + "linear1.weight"], sd[p + "linear1.bias"] = l1.weight.detach().clone(), l1.bias.detach().clone()
        sd[p + "linear2.weight"], sd[p + "linear2.bias"] = l2.weight.detach().clone(), l2.bias.detach().clone()
        for n in ("norm1", "norm2", "norm3"):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = torch.ones(D), torch.zeros(D)
    sd["layer_norm.weight"], sd["layer_norm.bias"] = torch.ones(D), torch.zeros(D)
    return sd
