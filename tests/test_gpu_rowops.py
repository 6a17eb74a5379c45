"""GPU: HBM-bound row kernels — casts, column sums, residual+LayerNorm fwd/bwd — against torch."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

from vqa_model_builder_b200 import _lib, ops  # noqa: E402

DEV = "cuda"


def test_cast_roundtrip_and_colsum():
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(1000, 777, generator=g, device=DEV)
    xb = ops.cast(x, torch.bfloat16)
    assert torch.equal(xb, x.to(torch.bfloat16))
    assert torch.equal(ops.cast(xb, torch.float32), xb.float())
    for t in (x, xb):
        assert rel_err(ops.colsum(t.contiguous()), t.double().sum(0)) < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("R,D", [(2048, 768), (37, 64), (515, 1024)])
def test_add_ln_fwd_bwd(R, D, dtype):
    g = torch.Generator(device=DEV).manual_seed(R + D)
    x = torch.randn(R, D, generator=g, device=DEV).to(dtype)
    br = torch.randn(R, D, generator=g, device=DEV).to(dtype)
    gamma = (1 + 0.1 * torch.randn(D, generator=g, device=DEV)).requires_grad_()
    beta = (0.1 * torch.randn(D, generator=g, device=DEV)).requires_grad_()
    gout = torch.randn(R, D, generator=g, device=DEV).to(dtype)
    xr, brr = x.double().requires_grad_(), br.double().requires_grad_()
    gr, btr = gamma.detach().double().requires_grad_(), beta.detach().double().requires_grad_()
    ref = torch.nn.functional.layer_norm(xr + brr, (D,), gr, btr, 1e-5)
    (ref * gout.double()).sum().backward()
    xx, bb = x.clone().requires_grad_(), br.clone().requires_grad_()
    y = ops.AddLNFn.apply(xx, bb, gamma, beta, 1e-5)
    (y.float() * gout.float()).sum().backward()
    t = 1e-5 if dtype == torch.float32 else 6e-3
    assert rel_err(y, ref) < t
    assert rel_err(xx.grad, xr.grad) < t and rel_err(bb.grad, brr.grad) < t
    assert rel_err(gamma.grad, gr.grad) < (1e-5 if dtype == torch.float32 else 2e-3)
    assert rel_err(beta.grad, btr.grad) < 1e-5
    # no-branch variant
    y2 = ops.AddLNFn.apply(x, None, gamma.detach(), beta.detach(), 1e-5)
    assert rel_err(y2, torch.nn.functional.layer_norm(x.double(), (D,), gr, btr, 1e-5)) < t
