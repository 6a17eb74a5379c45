"""GPU: train-mode dropout.  Masks are regenerated inside the kernels from a counter-based RNG; b200_dropout_mask
materialises the same mask so every fused site can be checked EXACTLY (forward and backward) against torch math."""
import math

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

from vqa_model_builder_b200 import _lib, fusion, moe, ops, runtime  # noqa: E402
from vqa_model_builder_b200._lib import ACT_GELU, EPI_ACT, EPI_ACT_D, EPI_ADD, EPI_DACT, EPI_MUL, LAYOUT_K  # noqa: E402
import vqa_model_builder_b200 as pkg  # noqa: E402

DEV = "cuda"


def state(seed=1234, off=7):
    return torch.tensor([seed, off], dtype=torch.int64, device=DEV)


def mask_of(drop, n):
    out = torch.empty(n, dtype=torch.float32, device=DEV)
    _lib.call("b200_dropout_mask", _lib.dropout_arg(drop), n, out, _lib.stream_ptr())
    return out


def test_mask_statistics_and_determinism():
    st = state()
    m = mask_of((st, 0.1, 3), 1 << 20)
    keep = (m > 0).float().mean().item()
    assert abs(keep - 0.9) < 2e-3
    assert torch.all((m == 0) | (torch.abs(m - 1 / 0.9) < 1e-6))
    assert torch.equal(m, mask_of((st, 0.1, 3), 1 << 20))
    assert not torch.equal(m, mask_of((st, 0.1, 4), 1 << 20))                 # other site
    assert not torch.equal(m, mask_of((state(off=8), 0.1, 3), 1 << 20))       # next step


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_epilogue_dropout_exact(dtype):
    M, N, K = 256, 192, 128
    g = torch.Generator(device=DEV).manual_seed(0)
    a = torch.randn(M, K, generator=g, device=DEV).to(dtype)
    b = (torch.randn(N, K, generator=g, device=DEV) * 0.1).to(dtype)
    aux = torch.randn(M, N, generator=g, device=DEV).to(dtype)
    drop = (state(), 0.25, 11)
    mask = mask_of(drop, M * N).view(M, N).double()
    acc = a.double() @ b.double().t()
    t = 2e-6 if dtype == torch.float32 else 6e-3
    pre = torch.empty(M, N, dtype=dtype, device=DEV)
    h = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, epi=EPI_ACT, act=ACT_GELU, aux_out=pre, drop=drop)
    assert rel_err(h, torch.nn.functional.gelu(acc) * mask) < t
    assert (h[mask == 0] == 0).all()
    x = aux.double().requires_grad_()
    torch.nn.functional.gelu(x).sum().backward()
    d = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, epi=EPI_DACT, act=ACT_GELU, aux_in=aux, drop=drop)
    assert rel_err(d, acc * x.grad * mask) < t
    r = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, epi=EPI_ADD, aux_in=aux, drop=drop)
    assert rel_err(r, acc * mask + aux.double()) < t
    # activation + saved backward factor (derivative x keep-scale), and the plain-multiply backward epilogue
    xa = acc.clone().requires_grad_()
    torch.nn.functional.gelu(xa).sum().backward()
    fac = torch.empty(M, N, dtype=dtype, device=DEV)
    h2 = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, epi=EPI_ACT_D, act=ACT_GELU, aux_out=fac, drop=drop)
    assert rel_err(h2, torch.nn.functional.gelu(acc) * mask) < t
    assert rel_err(fac, xa.grad * mask) < t
    assert (fac[mask == 0] == 0).all()
    mlt = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, epi=EPI_MUL, aux_in=aux)
    assert rel_err(mlt, acc * aux.double()) < t


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_add_ln_branch_dropout_exact(dtype):
    R, D = 300, 768
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(R, D, generator=g, device=DEV).to(dtype).requires_grad_()
    br = torch.randn(R, D, generator=g, device=DEV).to(dtype).requires_grad_()
    gamma = (1 + 0.1 * torch.randn(D, generator=g, device=DEV)).requires_grad_()
    beta = (0.1 * torch.randn(D, generator=g, device=DEV)).requires_grad_()
    gout = torch.randn(R, D, generator=g, device=DEV).to(dtype)
    drop = (state(), 0.1, 5)
    mask = mask_of(drop, R * D).view(R, D).double()
    y = ops.AddLNFn.apply(x, br, gamma, beta, 1e-5, drop)
    (y.float() * gout.float()).sum().backward()
    xr, brr = x.detach().double().requires_grad_(), br.detach().double().requires_grad_()
    gr, btr = gamma.detach().double().requires_grad_(), beta.detach().double().requires_grad_()
    ref = torch.nn.functional.layer_norm(xr + brr * mask, (D,), gr, btr, 1e-5)
    (ref * gout.double()).sum().backward()
    t = 1e-5 if dtype == torch.float32 else 6e-3
    assert rel_err(y, ref) < t
    assert rel_err(x.grad, xr.grad) < t and rel_err(br.grad, brr.grad) < t
    assert (br.grad[mask == 0] == 0).all()
    assert rel_err(gamma.grad, gr.grad) < (1e-5 if dtype == torch.float32 else 2e-3)


@pytest.mark.parametrize("dtype,B,H,T,S,D", [(torch.bfloat16, 4, 8, 64, 50, 768), (torch.bfloat16, 2, 8, 40, 257, 768),
                                              (torch.bfloat16, 2, 8, 197, 200, 768),   # query tiles x key tiles
                                              (torch.float32, 2, 4, 12, 7, 64), (torch.float32, 2, 8, 64, 150, 768)])
def test_attention_probability_dropout_exact(dtype, B, H, T, S, D):
    g = torch.Generator(device=DEV).manual_seed(2)
    q = torch.randn(B * T, D, generator=g, device=DEV).to(dtype).requires_grad_()
    kv = torch.randn(B * S, 2 * D, generator=g, device=DEV).to(dtype).requires_grad_()
    gout = torch.randn(B * T, D, generator=g, device=DEV).to(dtype)
    drop = (state(), 0.2, 9)
    s_pad = (S + 127) // 128 * 128
    mask = mask_of(drop, B * H * T * s_pad).view(B, H, T, s_pad)[..., :S].double()
    o = ops.AttentionFn.apply(q, kv, None, B, T, S, H, False, drop)
    (o.float() * gout.float()).sum().backward()
    dh = D // H
    qr, kvr = q.detach().double().requires_grad_(), kv.detach().double().requires_grad_()
    qh = qr.view(B, T, H, dh).transpose(1, 2) / math.sqrt(dh)
    kh = kvr[:, :D].reshape(B, S, H, dh).transpose(1, 2)
    vh = kvr[:, D:].reshape(B, S, H, dh).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2), -1) * mask
    ref = (p @ vh).transpose(1, 2).reshape(B * T, D)
    (ref * gout.double()).sum().backward()
    t = 2e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(o, ref) < t, rel_err(o, ref)
    assert rel_err(q.grad, qr.grad) < t, rel_err(q.grad, qr.grad)
    assert rel_err(kv.grad, kvr.grad) < t, rel_err(kv.grad, kvr.grad)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_modules_train_mode_dropout(mode):
    """Train mode with p=0.1: stochastic across steps, reproducible after reseeding, eval mode unaffected,
    gradients finite; gradient accumulation (two forwards before the backwards) keeps each forward's own masks."""
    pkg.set_compute_dtype(mode)
    try:
        torch.manual_seed(0)
        m = fusion.MultimodalFusion(fusion.FusionConfig("cross_attention", 256, 256, 4, 2, 0.1, True)).to(DEV)
        layer = moe.MOELayer(input_dim=256, hidden_dim=512, output_dim=256, num_experts=4, top_k=2, dropout=0.1).to(DEV)
        vis = torch.randn(4, 10, 256, device=DEV)
        txt = torch.randn(4, 16, 256, device=DEV, requires_grad=True)

        def run():
            return layer(m(vis, txt).unsqueeze(1))

        m.eval(); layer.eval()
        e1, e2 = run(), run()
        assert torch.equal(e1, e2)
        m.train(); layer.train()
        runtime.reseed_dropout(123)
        t1, t2 = run(), run()
        assert not torch.equal(t1, t2) and not torch.equal(t1, e1)
        runtime.reseed_dropout(123)
        t1b = run()
        assert torch.equal(t1, t1b)
        # gradient accumulation: grads of forward A must not depend on forward B having run in between
        runtime.reseed_dropout(7)
        a = run()
        ga, = torch.autograd.grad(a.float().square().sum(), txt)
        runtime.reseed_dropout(7)
        a2 = run()
        _b = run()
        ga2, = torch.autograd.grad(a2.float().square().sum(), txt)
        assert torch.equal(ga, ga2)
        assert torch.isfinite(ga).all()
    finally:
        pkg.set_compute_dtype("auto")
