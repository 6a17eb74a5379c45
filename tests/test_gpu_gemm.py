"""GPU: the tcgen05/TMA bf16 GEMM and the fp32 SIMT GEMM through the C-ABI, every operand layout, epilogue and
grouped mode, against torch matmul on the same (bf16-representable) inputs."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

from vqa_model_builder_b200 import _lib, ops  # noqa: E402
from vqa_model_builder_b200._lib import (ACT_GELU, ACT_NONE, ACT_RELU, EPI_ACCUM, EPI_ACT, EPI_ADD, EPI_DACT,  # noqa
                                         EPI_NONE, LAYOUT_K, LAYOUT_MN)

DEV = "cuda"
SHAPES = [(128, 64, 64), (256, 128, 192), (304, 200, 136), (2048, 768, 768), (2048, 2304, 768), (96, 3072, 768),
          (128, 64, 32), (32, 768, 768)]


def mk(shape, dtype, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device=DEV) * scale).to(dtype)


def tol(dtype):
    return 2e-6 if dtype == torch.float32 else 5e-3   # bf16: output rounding to bf16 dominates


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_layouts(M, N, K, dtype):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K)
    a = mk((M, K), dtype, g)
    b = mk((N, K), dtype, g)
    ref = a.double() @ b.double().t()
    # forward layout: both K-major
    out = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K)
    assert out.dtype == dtype
    assert rel_err(out, ref) < tol(dtype), ("KK", rel_err(out, ref))
    # dgrad layout: B stored [K, N]
    bt = b.t().contiguous()
    out = ops.gemm(a, LAYOUT_K, bt, LAYOUT_MN, M, N, K)
    assert rel_err(out, ref) < tol(dtype), ("K,MN", rel_err(out, ref))
    # wgrad layout: both stored [K, rows]; fp32 output
    at = a.t().contiguous()
    out = ops.gemm(at, LAYOUT_MN, bt, LAYOUT_MN, M, N, K, out_dtype=torch.float32)
    assert out.dtype == torch.float32
    assert rel_err(out, ref) < (2e-6 if dtype == torch.float32 else 1e-5), ("MN,MN", rel_err(out, ref))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_epilogues(dtype):
    M, N, K = 384, 320, 256
    g = torch.Generator(device=DEV).manual_seed(11)
    a, b = mk((M, K), dtype, g), mk((N, K), dtype, g, 0.1)
    bias = torch.randn(N, generator=g, device=DEV)
    aux = mk((M, N), dtype, g)
    acc = a.double() @ b.double().t()
    out = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, bias=bias)
    assert rel_err(out, acc + bias.double()) < tol(dtype)
    for act, fn in ((ACT_GELU, torch.nn.functional.gelu), (ACT_RELU, torch.relu)):
        pre = torch.empty((M, N), dtype=dtype, device=DEV)
        out = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, bias=bias, epi=EPI_ACT, act=act, aux_out=pre)
        assert rel_err(pre, acc + bias.double()) < tol(dtype)
        assert rel_err(out, fn(acc + bias.double())) < tol(dtype)
    out = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, bias=bias, epi=EPI_ADD, aux_in=aux)
    assert rel_err(out, acc + bias.double() + aux.double()) < tol(dtype)
    x = aux.double().requires_grad_()
    torch.nn.functional.gelu(x).sum().backward()
    out = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, epi=EPI_DACT, act=ACT_GELU, aux_in=aux)
    assert rel_err(out, acc * x.grad) < tol(dtype)


def test_gemm_split_k_accumulate():
    M, N, K = 768, 768, 4096     # 36 output tiles -> split-K kicks in
    g = torch.Generator(device=DEV).manual_seed(5)
    at, bt = mk((K, M), torch.bfloat16, g), mk((K, N), torch.bfloat16, g)
    out = torch.full((M, N), 7.0, dtype=torch.float32, device=DEV)   # must be overwritten, not accumulated onto
    ops.gemm(at, LAYOUT_MN, bt, LAYOUT_MN, M, N, K, out=out, epi=EPI_ACCUM)
    ref = at.double().t() @ bt.double()
    assert rel_err(out, ref) < 1e-5


@pytest.mark.parametrize("M", [1, 7, 32, 33, 64])
@pytest.mark.parametrize("N,K", [(768, 768), (768, 3072), (3072, 768), (104, 256), (768, 320), (64, 1024)])
def test_gemm_small_m_cluster_split_k(M, N, K):
    """Single-row-tile GEMMs with M <= 64 split K over a thread-block cluster (partial tiles reduced through
    distributed shared memory into the CTA that runs the epilogue): every epilogue, both weight layouts, K with and
    without a power-of-two number of k-blocks (K = 320: five k-blocks, no split), ragged N."""
    from vqa_model_builder_b200._lib import EPI_ACT_D, EPI_MUL
    dtype = torch.bfloat16
    g = torch.Generator(device=DEV).manual_seed(M * 131 + N + K)
    a, b = mk((M, K), dtype, g), mk((N, K), dtype, g, 0.05)
    bias = torch.randn(N, generator=g, device=DEV)
    aux = mk((M, N), dtype, g)
    acc = a.double() @ b.double().t()
    out = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, bias=bias)
    assert rel_err(out, acc + bias.double()) < tol(dtype), rel_err(out, acc + bias.double())
    out = ops.gemm(a, LAYOUT_K, b.t().contiguous(), LAYOUT_MN, M, N, K)            # dgrad layout
    assert rel_err(out, acc) < tol(dtype), rel_err(out, acc)
    out = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, bias=bias, epi=EPI_ADD, aux_in=aux)
    assert rel_err(out, acc + bias.double() + aux.double()) < tol(dtype)
    out = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, epi=EPI_MUL, aux_in=aux)
    assert rel_err(out, acc * aux.double()) < tol(dtype)
    if N % 8 == 0:
        dact = torch.empty((M, N), dtype=dtype, device=DEV)
        out = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, bias=bias, epi=EPI_ACT_D, act=ACT_GELU, aux_out=dact)
        x = (acc + bias.double()).requires_grad_()
        y = torch.nn.functional.gelu(x)
        y.sum().backward()
        assert rel_err(out, y) < tol(dtype)
        assert rel_err(dact, x.grad) < tol(dtype)
    # fp32 output of the same products (classifier logits)
    out = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, bias=bias, out_dtype=torch.float32)
    assert rel_err(out, acc + bias.double()) < 1e-5


def test_gemm_strided_rows():
    """A = CLS rows of a [B, T, D] activation (pitch T*D) — the pooled Linear of MultimodalFusion."""
    B, T, D, N = 32, 64, 768, 768
    g = torch.Generator(device=DEV).manual_seed(9)
    for dtype in (torch.bfloat16, torch.float32):
        x = mk((B, T, D), dtype, g)
        w = mk((N, D), dtype, g, 0.05)
        cls = x[:, 0, :]
        out = ops.gemm(cls, LAYOUT_K, w, LAYOUT_K, B, N, D)
        assert rel_err(out, cls.double() @ w.double().t()) < tol(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_grouped_gemm_and_wgrad(dtype):
    E, N_tok, Kk, D, F = 8, 700, 2, 128, 256
    g = torch.Generator(device=DEV).manual_seed(21)
    idx = torch.stack([torch.randperm(E, generator=g, device=DEV)[:Kk] for _ in range(N_tok)]).to(torch.int32)
    idx[5, 1] = -1                                              # dropped pair
    plan = ops.RoutingPlan(idx, E)
    R = plan.Rmax
    x = mk((N_tok, D), dtype, g)
    xp = torch.empty((R, D), dtype=dtype, device=DEV)
    _lib.call("b200_moe_permute", x, plan.row_src, plan.pad_off, E, Kk, R, D, _lib.dtype_code(dtype), xp,
              _lib.stream_ptr())
    row_src = plan.row_src.cpu()
    tile_group = plan.tile_group.cpu()
    pad_off = plan.pad_off.cpu().tolist()
    assert pad_off[-1] <= R
    used = pad_off[-1]
    # permuted rows are copies of their tokens, padding rows are zero
    for r in range(0, used, 37):
        s = int(row_src[r])
        want = x[s // Kk] if s >= 0 else torch.zeros(D, dtype=dtype, device=DEV)
        assert torch.equal(xp[r], want)
    w1 = mk((E, F, D), dtype, g, 0.1)
    b1 = torch.randn(E, F, generator=g, device=DEV)
    dt = _lib.dtype_code(dtype)
    h = torch.empty((R, F), dtype=dtype, device=DEV)
    pre = torch.empty((R, F), dtype=dtype, device=DEV)
    _lib.call("b200_ggemm", xp, D, w1, LAYOUT_K, h, F, R, F, D, E, plan.tile_group, None, dt, dt, b1, EPI_ACT, ACT_GELU, None,
              pre, F, None, _lib.stream_ptr())
    for e in range(E):
        r0, r1 = pad_off[e], pad_off[e + 1]
        if r1 == r0:
            continue
        ref = xp[r0:r1].double() @ w1[e].double().t() + b1[e].double()
        assert rel_err(pre[r0:r1], ref) < tol(dtype), ("pre", e)
        assert rel_err(h[r0:r1], torch.nn.functional.gelu(ref)) < tol(dtype), ("h", e)
    # dgrad layout: B = w1 read as [F(k), D(n)] per expert
    dx = torch.empty((R, D), dtype=dtype, device=DEV)
    _lib.call("b200_ggemm", h, F, w1, LAYOUT_MN, dx, D, R, D, F, E, plan.tile_group, plan.pad_off[E:E + 1], dt, dt, None, EPI_NONE, ACT_NONE,
              None, None, 0, None, _lib.stream_ptr())
    for e in range(E):
        r0, r1 = pad_off[e], pad_off[e + 1]
        if r1 > r0:
            assert rel_err(dx[r0:r1], h[r0:r1].double() @ w1[e].double()) < tol(dtype), ("dgrad", e)
    # wgrad: dW[e] = h_e^T xp_e  (padding rows of xp are zero)
    dw = torch.empty((E, F, D), dtype=torch.float32, device=DEV)
    _lib.call("b200_ggemm_wgrad", h, F, xp, D, dw, F, D, R, E, plan.pad_off, dt, _lib.stream_ptr())
    for e in range(E):
        r0, r1 = pad_off[e], pad_off[e + 1]
        ref = h[r0:r1].double().t() @ xp[r0:r1].double()
        if r1 > r0:
            assert rel_err(dw[e], ref) < 1e-5, ("wgrad", e, rel_err(dw[e], ref))
        else:
            assert float(dw[e].abs().max()) == 0.0
