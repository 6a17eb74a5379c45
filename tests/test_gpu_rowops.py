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


@pytest.mark.parametrize("D", [256, 768, 1024])
@pytest.mark.parametrize("drop_target", [0, 1, 2])
def test_grouped_add_ln_staged_path_against_torch(D, drop_target):
    """Per-expert add+LayerNorm over a padded expert layout (the persistent shared-memory-staged bf16 kernels:
    D == 256 * k): dead 128-row tiles in the middle and at the end, a group change inside a block's row range,
    dropout on either operand (mask materialised by b200_dropout_mask), dgamma / dbeta / bias-gradient sums per group."""
    G = 3
    tiles = torch.tensor([0, 0, -1, 1, 2, 2, 2, -1, -1], dtype=torch.int32, device=DEV)
    R = tiles.numel() * 128
    g = torch.Generator(device=DEV).manual_seed(D + drop_target)
    x = torch.randn(R, D, generator=g, device=DEV).bfloat16()
    res = torch.randn(R, D, generator=g, device=DEV).bfloat16()
    dy = torch.randn(R, D, generator=g, device=DEV).bfloat16()
    gamma = 1 + 0.1 * torch.randn(G, D, generator=g, device=DEV)
    beta = 0.1 * torch.randn(G, D, generator=g, device=DEV)
    st = torch.tensor([77, 5], dtype=torch.int64, device=DEV)
    drop = (st, 0.2, 9) if drop_target else None
    y = torch.zeros_like(x)
    mean = torch.zeros(R, device=DEV)
    rstd = torch.zeros(R, device=DEV)
    _lib.call("b200_add_ln_fwd", x, res, gamma, beta, tiles, 1e-5, y, mean, rstd, R, D, _lib.BF16,
              _lib.dropout_arg(drop), drop_target, _lib.stream_ptr())
    dsum = torch.zeros_like(x)
    ddrop = torch.zeros_like(x)
    dgam, dbet, dcol = (torch.zeros(G, D, device=DEV) for _ in range(3))
    nb = _lib.query("b200_add_ln_bwd_ws", R, D)
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    _lib.call("b200_add_ln_bwd", dy, x, res, mean, rstd, gamma, tiles, G, dsum, dgam, dbet, dcol, R, D, _lib.BF16,
              _lib.dropout_arg(drop), drop_target, ddrop if drop_target else None, ws, nb, _lib.stream_ptr())
    torch.cuda.synchronize()
    mask = torch.ones(R, D, dtype=torch.float64, device=DEV)
    if drop_target:
        m = torch.empty(R * D, dtype=torch.float32, device=DEV)
        _lib.call("b200_dropout_mask", _lib.dropout_arg(drop), R * D, m, _lib.stream_ptr())
        mask = m.view(R, D).double()
    live = (tiles >= 0).repeat_interleave(128)
    grp = tiles.clamp(min=0).long().repeat_interleave(128)
    xr, rr = x.double().requires_grad_(), res.double().requires_grad_()
    gr, br = gamma.double().requires_grad_(), beta.double().requires_grad_()
    xs = (xr * mask if drop_target == 1 else xr) + (rr * mask if drop_target == 2 else rr)
    ref = torch.nn.functional.layer_norm(xs, (D,)) * gr[grp] + br[grp]
    (ref * dy.double())[live].sum().backward()
    assert rel_err(y[live], ref[live]) < 6e-3
    dxs = rr.grad if drop_target == 1 else xr.grad      # gradient of the sum = gradient of the operand NOT dropped
    assert rel_err(dsum[live], dxs[live]) < 6e-3
    if drop_target:
        dropped = xr.grad if drop_target == 1 else rr.grad
        assert rel_err(ddrop[live], dropped[live]) < 6e-3
        assert (ddrop[live][mask[live] == 0] == 0).all()
    assert rel_err(dgam, gr.grad) < 3e-3 and rel_err(dbet, br.grad) < 1e-5
    want_col = (xr.grad if drop_target == 1 else rr.grad if drop_target == 2 else dxs)
    col_ref = torch.zeros(G, D, dtype=torch.float64, device=DEV).index_add_(0, grp[live], want_col[live])
    assert rel_err(dcol, col_ref) < 3e-3


@pytest.mark.parametrize("D,K", [(768, 2), (256, 1), (1024, 2), (768, 3), (64, 2)])
def test_combine_fwd_bwd_against_torch(D, K):
    """Weighted combine + output LayerNorm over scattered expert rows (bf16): the staged kernels (D == 256 k, K <= 2)
    and the register kernels (everything else) against torch fp64; dropped pairs (dest < 0), zero weights (capacity),
    a token count that is not a multiple of the 8-token chunk."""
    N = 1003
    R = N * K + 300
    g = torch.Generator(device=DEV).manual_seed(D * K)
    z = torch.randn(R, D, generator=g, device=DEV).bfloat16()
    w = torch.rand(N, K, generator=g, device=DEV)
    perm = torch.randperm(R, generator=g, device=DEV)[:N * K].view(N, K).to(torch.int32)
    dest = perm.clone()
    dest[5, 0] = -1
    dest[77, K - 1] = -1
    w[9, 0] = 0.0
    gamma = (1 + 0.1 * torch.randn(D, generator=g, device=DEV)).requires_grad_()
    beta = (0.1 * torch.randn(D, generator=g, device=DEV)).requires_grad_()
    dout = torch.randn(N, D, generator=g, device=DEV).bfloat16()
    row_src = torch.full((R,), -1, dtype=torch.int32, device=DEV)
    flat = torch.arange(N * K, device=DEV, dtype=torch.int32).view(N, K)
    ok = dest >= 0
    row_src[dest[ok].long()] = flat[ok]
    zz = z.clone().requires_grad_()
    ww = w.clone().requires_grad_()
    out = ops.CombineFn.apply(zz, ww, dest.contiguous(), row_src, gamma, beta, 1e-5)
    (out.float() * dout.float()).sum().backward()
    zr, wr = z.double().requires_grad_(), w.double().requires_grad_()
    gr, br = gamma.detach().double().requires_grad_(), beta.detach().double().requires_grad_()
    rows = zr[dest.clamp(min=0).long()] * ok.unsqueeze(-1)
    ref = torch.nn.functional.layer_norm((wr.unsqueeze(-1) * rows).sum(1), (D,), gr, br, 1e-5)
    (ref * dout.double()).sum().backward()
    assert rel_err(out, ref) < 6e-3
    assert rel_err(zz.grad, zr.grad) < 8e-3
    assert (zz.grad[row_src < 0] == 0).all()
    assert rel_err(ww.grad, wr.grad * ok) < 6e-3
    assert rel_err(gamma.grad, gr.grad) < 3e-3 and rel_err(beta.grad, br.grad) < 1e-5
