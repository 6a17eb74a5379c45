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


@pytest.mark.parametrize("R", [1, 63, 64, 4096, 4097, 20000, 116736])
@pytest.mark.parametrize("N", [768, 100])
def test_colsum_single_launch_and_two_stage_paths(R, N):
    """Ungrouped column sums are ONE launch for any R (rows per block grow with R; the last-arriving block of a column
    slab folds at most 64 partials in block order); N = 100 exercises the scalar (non-vectorised) variant for bf16."""
    g = torch.Generator(device=DEV).manual_seed(R + N)
    x = torch.randn(R, N, generator=g, device=DEV)
    for t in (x, x.to(torch.bfloat16)):
        _lib.reset_launch_count()
        got = ops.colsum(t.contiguous())
        launches = _lib.launch_count()
        assert rel_err(got, t.double().sum(0)) < 2e-6
        assert launches == 1, launches
    # repeated calls reuse (and re-arm) the ticket counters
    for _ in range(3):
        assert rel_err(ops.colsum(x), x.double().sum(0)) < 2e-6


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_dropout_colsum_matches_mask_then_sum(dtype):
    """b200_dropout_colsum: out = x * keep-scales of the site, colsum = column sums of out, in one pass."""
    from vqa_model_builder_b200 import runtime
    R, N = 1000, 768
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(R, N, generator=g, device=DEV).to(dtype)
    st = runtime.dropout_state(torch.device(DEV)).clone()
    drop = (st, 0.1, 77)
    mask = torch.empty(R * N, dtype=torch.float32, device=DEV)
    _lib.call("b200_dropout_mask", _lib.dropout_arg(drop), R * N, mask, _lib.stream_ptr())
    cs = torch.empty(N, dtype=torch.float32, device=DEV)
    out = ops.dropout_colsum(x, drop, cs)
    want = (x.double() * mask.view(R, N).double())
    assert rel_err(out, want) < (1e-6 if dtype == torch.float32 else 4e-3)
    assert rel_err(cs, want.sum(0)) < 2e-6          # sums are taken before the products are rounded to the row dtype
    assert torch.equal(out, ops.dropout_apply(x, drop))


def test_colsum_grouped_by_expert_tiles():
    g = torch.Generator(device=DEV).manual_seed(5)
    tiles = torch.tensor([0, 0, 2, 3, 3, 3, -1, -1], dtype=torch.int32, device=DEV)   # expert 1 owns no tile
    x = torch.randn(tiles.numel() * 128, 256, generator=g, device=DEV).to(torch.bfloat16)
    got = ops.colsum(x, tile_group=tiles, G=4)
    for e in range(4):
        rows = torch.cat([x[t * 128:(t + 1) * 128] for t in range(tiles.numel()) if int(tiles[t]) == e] or
                         [x[:0]])
        assert rel_err(got[e], rows.double().sum(0)) < 2e-6 or float(rows.abs().sum()) == 0.0
    assert float(got[1].abs().max()) == 0.0
