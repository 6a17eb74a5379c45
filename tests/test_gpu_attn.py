"""GPU: fused attention core (online softmax, key-padding mask, packed q/k/v views) fwd + bwd against the
explicit softmax(QK^T)V math of nn.MultiheadAttention."""
import math

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

from vqa_model_builder_b200 import ops  # noqa: E402

DEV = "cuda"


def ref_attn(q, k, v, pad, H, causal=False):
    B, T, D = q.shape
    S = k.shape[1]
    dh = D // H
    qh = q.view(B, T, H, dh).transpose(1, 2) / math.sqrt(dh)
    kh = k.view(B, S, H, dh).transpose(1, 2)
    vh = v.view(B, S, H, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if pad is not None:
        s = s.masked_fill(pad.bool()[:, None, None, :], float("-inf"))
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(T, S, dtype=torch.bool, device=s.device), diagonal=1), float("-inf"))
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, T, D)


CASES = [(2, 4, 12, 7, 64), (32, 8, 64, 50, 768), (4, 8, 64, 257, 768), (3, 8, 114, 114, 768), (2, 8, 50, 64, 1024),
         (2, 4, 70, 130, 128),
         # 64-row flavour (T, S <= 64): odd sizes, d_h = 64 / 96 / 128, a single query, many co-resident CTAs
         (5, 8, 33, 17, 768), (3, 4, 64, 64, 256), (2, 8, 1, 50, 768), (3, 8, 64, 3, 1024), (200, 8, 64, 50, 768),
         (2, 8, 65, 64, 768), (2, 8, 64, 65, 768),
         # more than 128 query rows on the tensor cores (bf16): one CTA per 128-row query tile forward; backward splits
         # into dQ-tile and dK/dV-tile CTAs.  ViT-B/16 (197) and DINOv2 (257) patch tokens as queries, the single-stream
         # fusion's 328 tokens, a ragged last tile, three key tiles
         (2, 8, 197, 64, 768), (2, 8, 257, 40, 768), (3, 8, 129, 50, 768), (1, 8, 328, 328, 768), (2, 4, 300, 130, 128),
         (2, 2, 384, 384, 256)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,T,S,D", CASES)
def test_cross_attention(B, H, T, S, D, dtype):
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + T + S)
    q = torch.randn(B * T, D, generator=g, device=DEV).to(dtype).requires_grad_()
    kv = torch.randn(B * S, 2 * D, generator=g, device=DEV).to(dtype).requires_grad_()
    pad = torch.zeros(B, S, dtype=torch.uint8, device=DEV)
    for b in range(B):
        pad[b, S - (b % 5):] = 1 if b % 5 else 0
    gout = torch.randn(B * T, D, generator=g, device=DEV).to(dtype)
    o = ops.AttentionFn.apply(q, kv, pad, B, T, S, H, False)
    (o.float() * gout.float()).sum().backward()
    qr = q.detach().double().requires_grad_()
    kvr = kv.detach().double().requires_grad_()
    ref = ref_attn(qr.view(B, T, D), kvr[:, :D].reshape(B, S, D), kvr[:, D:].reshape(B, S, D), pad, H)
    (ref.reshape(B * T, D) * gout.double()).sum().backward()
    t = 2e-5 if dtype == torch.float32 else 8e-3
    assert rel_err(o, ref.reshape(B * T, D)) < t, rel_err(o, ref.reshape(B * T, D))
    assert rel_err(q.grad, qr.grad) < t, ("dq", rel_err(q.grad, qr.grad))
    assert rel_err(kv.grad, kvr.grad) < t, ("dkv", rel_err(kv.grad, kvr.grad))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_self_attention_packed(dtype):
    B, H, T, D = 8, 8, 64, 768
    g = torch.Generator(device=DEV).manual_seed(3)
    qkv = torch.randn(B * T, 3 * D, generator=g, device=DEV).to(dtype).requires_grad_()
    pad = torch.zeros(B, T, dtype=torch.uint8, device=DEV)
    pad[:, 40:] = 1
    gout = torch.randn(B * T, D, generator=g, device=DEV).to(dtype)
    o = ops.AttentionFn.apply(qkv, None, pad, B, T, T, H, True)
    (o.float() * gout.float()).sum().backward()
    r = qkv.detach().double().requires_grad_()
    ref = ref_attn(r[:, :D].reshape(B, T, D), r[:, D:2 * D].reshape(B, T, D), r[:, 2 * D:].reshape(B, T, D), pad, H)
    (ref.reshape(B * T, D) * gout.double()).sum().backward()
    t = 2e-5 if dtype == torch.float32 else 8e-3
    assert rel_err(o, ref.reshape(B * T, D)) < t
    assert rel_err(qkv.grad, r.grad) < t


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,T,D", [(3, 8, 64, 768), (2, 4, 9, 64), (2, 8, 114, 768), (2, 8, 33, 768), (1, 2, 200, 64),
                                     (2, 8, 328, 768), (3, 4, 257, 128)])
def test_causal_self_attention(B, H, T, D, dtype):
    """Decoder self-attention (generative_vqa_model.py:404-406): causal mask combined with a key-padding mask, on the
    64-row and 128-row tensor-core flavours (bf16; T > 128 runs one CTA per query tile), the SIMT kernels (fp32) —
    forward and backward."""
    g = torch.Generator(device=DEV).manual_seed(B * 100 + T)
    qkv = torch.randn(B * T, 3 * D, generator=g, device=DEV).to(dtype).requires_grad_()
    pad = torch.zeros(B, T, dtype=torch.uint8, device=DEV)
    for b in range(B):
        pad[b, T - b:] = 1 if b else 0           # trailing padding; position 0 always valid
    gout = torch.randn(B * T, D, generator=g, device=DEV).to(dtype)
    o = ops.AttentionFn.apply(qkv, None, pad, B, T, T, H, True, None, True)
    (o.float() * gout.float()).sum().backward()
    r = qkv.detach().double().requires_grad_()
    ref = ref_attn(r[:, :D].reshape(B, T, D), r[:, D:2 * D].reshape(B, T, D), r[:, 2 * D:].reshape(B, T, D), pad, H,
                   causal=True)
    (ref.reshape(B * T, D) * gout.double()).sum().backward()
    t = 2e-5 if dtype == torch.float32 else 8e-3
    assert rel_err(o, ref.reshape(B * T, D)) < t
    assert rel_err(qkv.grad, r.grad) < t
