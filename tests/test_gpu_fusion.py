"""GPU: the fusion drop-ins against the reference's golden vectors (fp32 1e-4, bf16 1e-2) and against the CPU
oracle at the benchmark shapes (config 1/2: B=32, T=64, V=50, D=768, H=8, L=2)."""
import pytest
import torch

from conftest import bf16_representable, leafs, load_golden, rel_err, round_sd_for_bf16
from oracle import reference_port as rp

pytestmark = pytest.mark.gpu

from vqa_model_builder_b200 import fusion  # noqa: E402
import vqa_model_builder_b200 as pkg  # noqa: E402

DEV = "cuda"
MODES = [("fp32", 1e-4), ("bf16", 1e-2)]


def worst_grad(module, golden_grads):
    return max((rel_err(p.grad, golden_grads[k]), k) for k, p in module.named_parameters()
               if float(golden_grads[k].norm()) > 0)


def reference_for(mode, g, names, run):
    """fp32: the reference's golden vectors.  bf16: the (golden-pinned) oracle on bf16-representable copies of the
    same weights and inputs, so both sides start from identical values."""
    sd = g["sd"]
    inputs = [g[n] for n in names]
    if mode == "fp32":
        return sd, inputs, None
    sd = round_sd_for_bf16(sd)
    inputs = [bf16_representable(t) for t in inputs]
    sdr = leafs(sd)
    leaf_in = [t.clone().requires_grad_() for t in inputs]
    out = run(sdr, *leaf_in)
    (out * g["gout"]).sum().backward()
    ref = dict(out=out.detach(), d_in=[t.grad for t in leaf_in],
               grads={k: v.grad for k, v in sdr.items() if v.grad is not None})
    return sd, inputs, ref


@pytest.mark.parametrize("mode,tol", MODES)
def test_multimodal_fusion_cross_attention_golden(mode, tol):
    g = load_golden("multimodal_fusion_xattn")
    B, T, V, D, H, L = [int(v) for v in g["cfg"]]
    pad = ~g["text_valid"]
    sd, (vis0, txt0), ref = reference_for(
        mode, g, ("visual", "text"),
        lambda s, v, t: rp.multimodal_fusion(s, "cross_attention", H, L, True, v, t, None, pad))
    if ref is None:
        ref = dict(out=g["out"], d_in=[g["d_visual"], g["d_text"]], grads=g["grads"])
    pkg.set_compute_dtype(mode)
    try:
        m = fusion.MultimodalFusion(fusion.FusionConfig("cross_attention", D, D, H, L, 0.0, True)).to(DEV)
        m.load_state_dict(sd)
        m.train()
        vis = vis0.to(DEV).requires_grad_()
        txt = txt0.to(DEV).requires_grad_()
        out = m(vis, txt, text_mask=pad.to(DEV))
        (out * g["gout"].to(DEV)).sum().backward()
    finally:
        pkg.set_compute_dtype("auto")
    assert out.shape == (B, D)
    assert rel_err(out, ref["out"]) < tol, rel_err(out, ref["out"])
    assert rel_err(vis.grad, ref["d_in"][0]) < tol, rel_err(vis.grad, ref["d_in"][0])
    assert rel_err(txt.grad, ref["d_in"][1]) < tol, rel_err(txt.grad, ref["d_in"][1])
    w = worst_grad(m, ref["grads"])
    assert w[0] < tol, w


def test_multimodal_fusion_other_branches_golden():
    for ft in ("concat", "add"):
        g = load_golden(f"multimodal_fusion_{ft}")
        m = fusion.MultimodalFusion(fusion.FusionConfig(ft, 64, 64, 4, 2, 0.0, True)).to(DEV)
        m.load_state_dict(g["sd"])
        out = m(g["visual"].to(DEV), g["text"].to(DEV))
        assert rel_err(out, g["out"]) < 1e-4, ft


@pytest.mark.parametrize("mode,tol", MODES)
def test_cross_attention_fusion_golden(mode, tol):
    g = load_golden("cross_attention_fusion")
    B, T, V, D, H, L, I = [int(v) for v in g["cfg"]]
    sd, (vis0, txt0), ref = reference_for(
        mode, g, ("vision", "text"),
        lambda s, v, t: rp.cross_attention_fusion(s, H, L, "concat", v, t, None, g["text_valid"]))
    if ref is None:
        ref = dict(out=g["out"], d_in=[g["d_vision"], g["d_text"]], grads=g["grads"])
    pkg.set_compute_dtype(mode)
    try:
        m = fusion.CrossAttentionFusion(D, D, D, H, L, I, 0.0, "concat").to(DEV)
        m.load_state_dict(sd)
        m.train()
        vis = vis0.to(DEV).requires_grad_()
        txt = txt0.to(DEV).requires_grad_()
        out = m(vis, txt, text_mask=g["text_valid"].to(DEV))
        (out * g["gout"].to(DEV)).sum().backward()
    finally:
        pkg.set_compute_dtype("auto")
    assert rel_err(out, ref["out"]) < tol, rel_err(out, ref["out"])
    assert rel_err(vis.grad, ref["d_in"][0]) < tol, rel_err(vis.grad, ref["d_in"][0])
    assert rel_err(txt.grad, ref["d_in"][1]) < tol, rel_err(txt.grad, ref["d_in"][1])
    w = worst_grad(m, ref["grads"])
    assert w[0] < tol, w


@pytest.mark.parametrize("mode,tol", MODES)
def test_cross_modal_fusion_with_moe_golden(mode, tol):
    g = load_golden("cross_modal_fusion_moe")
    B, V, T, D, H, F, E = [int(v) for v in g["cfg"]]
    cfg = fusion.GenerativeFusionConfig(fusion_dim=D, fusion_num_heads=H, fusion_num_layers=2, fusion_dropout=0.0,
                                        decoder_ff_dim=F, use_moe=True, num_experts=E, num_experts_per_token=2)
    sd, (vis0, q0), ref = reference_for(
        mode, g, ("visual", "question"),
        lambda s, v, q: rp.cross_modal_fusion(s, H, 2, v, q, g["question_valid"], moe=dict(num_experts=E, top_k=2))[0])
    if ref is None:
        ref = dict(out=g["out"], d_in=[g["d_visual"], g["d_question"]], grads=g["grads"])
    pkg.set_compute_dtype(mode)
    try:
        m = fusion.CrossModalFusion(cfg).to(DEV)
        m.load_state_dict(sd)
        m.train()
        vis = vis0.to(DEV).requires_grad_()
        q = q0.to(DEV).requires_grad_()
        out, aux = m(vis, q, g["question_valid"].long().to(DEV))
        (out * g["gout"].to(DEV)).sum().backward()
    finally:
        pkg.set_compute_dtype("auto")
    assert isinstance(aux, float)
    if mode == "fp32":
        assert abs(aux - float(g["aux"])) < 1e-6
    assert rel_err(out, ref["out"]) < tol, rel_err(out, ref["out"])
    assert rel_err(vis.grad, ref["d_in"][0]) < tol, rel_err(vis.grad, ref["d_in"][0])
    assert rel_err(q.grad, ref["d_in"][1]) < tol, rel_err(q.grad, ref["d_in"][1])
    w = worst_grad(m, ref["grads"])
    assert w[0] < tol, w


@pytest.mark.parametrize("mode,tol", MODES)
def test_multimodal_fusion_at_benchmark_shape_vs_oracle(mode, tol):
    """config 1/2 shapes; oracle runs on the CPU on the same (bf16-representable in bf16 mode) values."""
    B, T, V, D, H, L = 8, 64, 50, 768, 8, 2
    torch.manual_seed(0)
    m = fusion.MultimodalFusion(fusion.FusionConfig("cross_attention", D, D, H, L, 0.0, True))
    gen = torch.Generator().manual_seed(1234)
    vis = torch.randn(B, V, D, generator=gen)
    txt = torch.randn(B, T, D, generator=gen)
    lens = torch.randint(8, T + 1, (B,), generator=torch.Generator().manual_seed(4321))
    valid = torch.arange(T)[None, :] < lens[:, None]
    gout = torch.randn(B, D, generator=torch.Generator().manual_seed(99))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    if mode == "bf16":
        sd = {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 else v) for k, v in sd.items()}
        vis, txt = vis.to(torch.bfloat16).float(), txt.to(torch.bfloat16).float()
        m.load_state_dict(sd)
    sdr = {k: v.clone().requires_grad_() for k, v in sd.items()}
    vr, tr = vis.clone().requires_grad_(), txt.clone().requires_grad_()
    ref = rp.multimodal_fusion(sdr, "cross_attention", H, L, True, vr, tr, None, ~valid)
    (ref * gout).sum().backward()
    pkg.set_compute_dtype(mode)
    try:
        m = m.to(DEV).train()
        vg, tg = vis.to(DEV).requires_grad_(), txt.to(DEV).requires_grad_()
        out = m(vg, tg, text_mask=~valid.to(DEV))
        (out * gout.to(DEV)).sum().backward()
    finally:
        pkg.set_compute_dtype("auto")
    assert rel_err(out, ref) < tol, rel_err(out, ref)
    assert rel_err(vg.grad, vr.grad) < tol, rel_err(vg.grad, vr.grad)
    assert rel_err(tg.grad, tr.grad) < tol, rel_err(tg.grad, tr.grad)
    worst = max((rel_err(p.grad, sdr[k].grad), k) for k, p in m.named_parameters())
    assert worst[0] < tol, worst
