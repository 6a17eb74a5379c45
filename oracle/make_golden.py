"""Generates tests/golden/*.npz by running the REFERENCE's own modules (imported from /root/reference) on seeded
synthetic inputs, forward + backward, CPU fp32.  Run here (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Weights and inputs come from numpy's PCG64 streams (stable across torch versions) and are loaded into the reference
modules via load_state_dict; dropout is 0 so train-mode code paths are deterministic."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

REF = os.environ.get("B200VQA_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, str(Path(__file__).resolve().parent.parent))
sys.dont_write_bytecode = True
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def rnd_state_dict(module: torch.nn.Module, seed: int) -> dict:
    from oracle.init_weights import seeded_state_dict
    return seeded_state_dict(module.state_dict(), seed)


def save(name: str, **arrays):
    OUT.mkdir(parents=True, exist_ok=True)
    flat = {}
    for k, v in arrays.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat[f"{k}/{kk}"] = vv.detach().numpy() if isinstance(vv, torch.Tensor) else np.asarray(vv)
        else:
            flat[k] = v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    np.savez_compressed(OUT / f"{name}.npz", **flat)
    print(name, sum(a.nbytes for a in flat.values()) // 1024, "KiB")


def grads_of(module):
    return {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in module.named_parameters()}


def lengths_mask(rng, B, T):
    lens = rng.integers(low=max(2, T // 3), high=T + 1, size=B)
    m = torch.zeros(B, T, dtype=torch.bool)
    for b, n in enumerate(lens):
        m[b, :n] = True
    return m  # True = valid


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    from src.modeling.meta_arch.vqa_model import MultimodalFusion
    from src.modeling.meta_arch.vqa_config import FusionConfig
    from src.modeling.fusion.fusion_approaches import CrossAttentionFusion
    from src.modeling.moe.moe_layer import MOELayer, SparseMOELayer
    from src.modeling.moe.router import TopKRouter, NoisyTopKRouter
    from src.modeling.meta_arch.generative_vqa_model import CrossModalFusion, GenerativeVQAConfig

    rng = np.random.default_rng(20261018)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)

    # ---- MultimodalFusion, cross_attention (A1+A2) -------------------------------------------------------
    B, T, V, D, H, L = 3, 12, 7, 64, 4, 2
    m = MultimodalFusion(FusionConfig(fusion_type="cross_attention", hidden_dim=D, output_dim=D, num_heads=H,
                                      num_layers=L, dropout=0.0, use_layer_norm=True))
    sd = rnd_state_dict(m, 1)
    m.load_state_dict(sd)
    m.train()
    vis = f32(rng.standard_normal((B, V, D))).requires_grad_()
    txt = f32(rng.standard_normal((B, T, D))).requires_grad_()
    valid = lengths_mask(rng, B, T)
    gout = f32(rng.standard_normal((B, D)))
    out = m(vis, txt, text_mask=~valid)           # call-site polarity: True = PAD (vqa_model.py:662-666)
    (out * gout).sum().backward()
    save("multimodal_fusion_xattn", cfg=np.array([B, T, V, D, H, L]), sd=sd, visual=vis, text=txt,
         text_valid=valid, gout=gout, out=out, d_visual=vis.grad, d_text=txt.grad, grads=grads_of(m))

    # ---- MultimodalFusion other branches (forward only) ------------------------------------------------------
    for ft in ("concat", "add"):
        m = MultimodalFusion(FusionConfig(fusion_type=ft, hidden_dim=D, output_dim=D, num_heads=H, num_layers=L,
                                          dropout=0.0, use_layer_norm=True))
        sd = rnd_state_dict(m, 2)
        m.load_state_dict(sd)
        out = m(vis.detach(), txt.detach())
        save(f"multimodal_fusion_{ft}", sd=sd, visual=vis, text=txt, out=out)

    # ---- CrossAttentionFusion (A3) ---------------------------------------------------------------------------
    Bc, Tc, Vc, Dc, Hc, Lc, Ic = 2, 20, 36, 64, 4, 2, 128   # token counts of examples/fusion_examples.py:24-29
    m = CrossAttentionFusion(vision_dim=Dc, text_dim=Dc, output_dim=Dc, num_attention_heads=Hc, num_layers=Lc,
                             intermediate_dim=Ic, dropout=0.0, fusion_method="concat")
    sd = rnd_state_dict(m, 3)
    m.load_state_dict(sd)
    m.train()
    vis2 = f32(rng.standard_normal((Bc, Vc, Dc))).requires_grad_()
    txt2 = f32(rng.standard_normal((Bc, Tc, Dc))).requires_grad_()
    tvalid = torch.ones(Bc, Tc, dtype=torch.bool)
    tvalid[:, -5:] = False                                   # last 5 text tokens padded (fusion_examples.py:113-127)
    gout2 = f32(rng.standard_normal((Bc, Dc)))
    out = m(vis2, txt2, text_mask=tvalid)
    (out * gout2).sum().backward()
    save("cross_attention_fusion", cfg=np.array([Bc, Tc, Vc, Dc, Hc, Lc, Ic]), sd=sd, vision=vis2, text=txt2,
         text_valid=tvalid, gout=gout2, out=out, d_vision=vis2.grad, d_text=txt2.grad, grads=grads_of(m))

    # ---- routers (A5) ------------------------------------------------------------------------------------------
    Br, Sr, Dr, Er, Kr = 4, 32, 64, 8, 2
    xr = f32(rng.standard_normal((Br, Sr, Dr))).requires_grad_()
    r = TopKRouter(Dr, Er, top_k=Kr)
    sd = rnd_state_dict(r, 4)
    r.load_state_dict(sd)
    w, idx, aux = r(xr)
    gw = f32(rng.standard_normal(w.shape))
    ((w * gw).sum() + 3.0 * aux["load_balance_loss"]).backward()
    save("topk_router", cfg=np.array([Br, Sr, Dr, Er, Kr]), sd=sd, x=xr, w=w, idx=idx, loss=aux["load_balance_loss"],
         probs=aux["router_probs"], gw=gw, d_x=xr.grad, grads=grads_of(r))

    xr2 = f32(rng.standard_normal((Br, Sr, Dr))).requires_grad_()
    nr = NoisyTopKRouter(Dr, Er, top_k=Kr, noise_std=1.0)
    sd = rnd_state_dict(nr, 5)
    nr.load_state_dict(sd)
    nr.train()
    eps = f32(rng.standard_normal((Br, Sr, Er)))
    orig = torch.randn_like
    torch.randn_like = lambda t, **kw: eps.to(t.dtype)       # feed the recorded N(0,1) draw (router.py:308)
    try:
        w, idx, aux = nr(xr2)
    finally:
        torch.randn_like = orig
    gw2 = f32(rng.standard_normal(w.shape))
    ((w * gw2).sum() + 3.0 * aux["load_balance_loss"]).backward()
    save("noisy_router", cfg=np.array([Br, Sr, Dr, Er, Kr]), sd=sd, x=xr2, eps=eps, w=w, idx=idx,
         loss=aux["load_balance_loss"], probs=aux["router_probs"], noise_scale=aux["noise_scale"], gw=gw2,
         d_x=xr2.grad, grads=grads_of(nr))

    # ---- MOELayer, homogeneous FFN experts (A6+A7) -------------------------------------------------------------------
    Bm, Sm, Dm, Fm, Em, Km = 4, 16, 64, 128, 4, 2
    m = MOELayer(input_dim=Dm, hidden_dim=Fm, output_dim=Dm, num_experts=Em, top_k=Km, dropout=0.0)
    sd = rnd_state_dict(m, 6)
    m.load_state_dict(sd)
    m.train()
    xm = f32(rng.standard_normal((Bm, Sm, Dm))).requires_grad_()
    gm = f32(rng.standard_normal((Bm, Sm, Dm)))
    out = m(xm)
    ((out * gm).sum() + 2.0 * m.get_aux_loss()).backward()
    save("moe_layer", cfg=np.array([Bm, Sm, Dm, Fm, Em, Km]), sd=sd, x=xm, gout=gm, out=out,
         loss=m.get_aux_loss(), probs=m.aux_outputs["router_probs"], d_x=xm.grad, grads=grads_of(m))

    # ---- SparseMOELayer with an active capacity limit (A8) ------------------------------------------------------------
    m = SparseMOELayer(input_dim=Dm, hidden_dim=Fm, output_dim=Dm, num_experts=Em, top_k=Km, capacity_factor=0.6,
                       dropout=0.0)
    sd = rnd_state_dict(m, 7)
    m.load_state_dict(sd)
    m.eval()                                                # eval: NoisyTopKRouter is noise-free
    xs = f32(rng.standard_normal((Bm, Sm, Dm)))
    with torch.no_grad():
        out = m(xs)
    save("sparse_moe_layer", cfg=np.array([Bm, Sm, Dm, Fm, Em, Km]), capacity_factor=np.array(0.6), sd=sd, x=xs,
         out=out)

    # ---- CrossModalFusion with the standard MOE layer (A4) -------------------------------------------------------------
    Bg, Vg, Tg, Dg, Hg, Fg, Eg = 2, 10, 14, 64, 4, 128, 4
    cfg = GenerativeVQAConfig()
    cfg.fusion_dim, cfg.fusion_num_heads, cfg.fusion_num_layers, cfg.fusion_dropout = Dg, Hg, 2, 0.0
    cfg.decoder_ff_dim, cfg.use_moe, cfg.moe_type, cfg.moe_position = Fg, True, "standard", "fusion"
    cfg.num_experts, cfg.num_experts_per_token = Eg, 2
    m = CrossModalFusion(cfg)
    sd = rnd_state_dict(m, 8)
    m.load_state_dict(sd)
    m.train()
    visg = f32(rng.standard_normal((Bg, Vg, Dg))).requires_grad_()
    qg = f32(rng.standard_normal((Bg, Tg, Dg))).requires_grad_()
    qvalid = lengths_mask(rng, Bg, Tg)
    gg = f32(rng.standard_normal((Bg, Vg + Tg, Dg)))
    out, aux = m(visg, qg, qvalid.long())
    (out * gg).sum().backward()
    save("cross_modal_fusion_moe", cfg=np.array([Bg, Vg, Tg, Dg, Hg, Fg, Eg]), sd=sd, visual=visg, question=qg,
         question_valid=qvalid, gout=gg, out=out, aux=np.array(aux), d_visual=visg.grad, d_question=qg.grad,
         grads=grads_of(m))


if __name__ == "__main__" and not ({"--r2", "--n4", "--n3", "--n2"} & set(sys.argv)):
    main()


# =====================================================================================================================
# Round-2 fixtures (own RNG stream, so the fixtures above stay byte-identical when this file is re-run):
#   sparse_moe_layer_train : A8 in TRAIN mode — injected router noise, active capacity limit, forward + backward
#   vqa_moe_layer          : A9 — the reference's own VQAMOELayer (heterogeneous experts): router outputs, every
#                            expert's output, combined output, and the gradients that flow through router + combine
#   fp_*                   : fingerprints at the BENCHMARK dimensions (D=768, H=8, d_h=96, F=2048): full outputs,
#                            sha256 of the routing indices, gradient norms + seeded probes.  Weights and inputs are
#                            regenerated from numpy seeds by the tests (tests/fingerprint_util.py), not stored.
# =====================================================================================================================
def seeded_inputs(seed: int, *shapes):
    rng = np.random.default_rng(seed)
    return [torch.tensor(rng.standard_normal(s), dtype=torch.float32) for s in shapes]


from oracle.fingerprint import fingerprint, sha_int  # noqa: E402


def main_r2():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    from src.modeling.meta_arch.vqa_model import MultimodalFusion
    from src.modeling.meta_arch.vqa_config import FusionConfig
    from src.modeling.moe.moe_layer import MOELayer, SparseMOELayer, VQAMOELayer
    from src.modeling.meta_arch.generative_vqa_model import CrossModalFusion, GenerativeVQAConfig

    rng = np.random.default_rng(20261019)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)

    # ---- A8: SparseMOELayer, train mode, noise injected, capacity active, fwd + bwd ------------------------------
    Bm, Sm, Dm, Fm, Em, Km = 4, 24, 64, 128, 4, 2
    m = SparseMOELayer(input_dim=Dm, hidden_dim=Fm, output_dim=Dm, num_experts=Em, top_k=Km, capacity_factor=0.7,
                       dropout=0.0)
    sd = rnd_state_dict(m, 21)
    m.load_state_dict(sd)
    m.train()
    xs = f32(rng.standard_normal((Bm, Sm, Dm))).requires_grad_()
    eps = f32(rng.standard_normal((Bm, Sm, Em)))
    gs = f32(rng.standard_normal((Bm, Sm, Dm)))
    orig = torch.randn_like
    torch.randn_like = lambda t, **kw: eps.to(t.dtype)       # router.py:308
    try:
        out = m(xs)
    finally:
        torch.randn_like = orig
    ((out * gs).sum() + 2.0 * m.get_aux_loss()).backward()
    save("sparse_moe_layer_train", cfg=np.array([Bm, Sm, Dm, Fm, Em, Km]), capacity_factor=np.array(0.7), sd=sd,
         x=xs, eps=eps, gout=gs, out=out, loss=m.get_aux_loss(), probs=m.aux_outputs["router_probs"], d_x=xs.grad,
         grads=grads_of(m))

    # ---- A9: the reference's VQAMOELayer (2/2/2/2 heterogeneous experts, NoisyTopK), train mode, injected noise --
    Bv, Sv, Dv, Fv, Kv = 3, 5, 64, 128, 2
    m = VQAMOELayer(input_dim=Dv, hidden_dim=Fv, output_dim=Dv, num_vision_experts=2, num_text_experts=2,
                    num_multimodal_experts=2, num_specialized_experts=2, top_k=Kv, dropout=0.0)
    Ev = m.num_experts
    torch.manual_seed(5)
    for p in m.parameters():                                 # the experts keep their own initialisers' structure;
        if p.dim() >= 2:                                     # the values only need to be deterministic here
            torch.nn.init.normal_(p, std=1.0 / np.sqrt(p.shape[-1]))
    m.router.load_state_dict(rnd_state_dict(m.router, 22))
    rn = np.random.default_rng(23)
    m.output_norm.load_state_dict({"weight": torch.tensor(1.0 + 0.1 * rn.standard_normal(Dv), dtype=torch.float32),
                                   "bias": torch.tensor(0.1 * rn.standard_normal(Dv), dtype=torch.float32)})
    m.train()
    xv = f32(rng.standard_normal((Bv, Sv, Dv))).requires_grad_()
    epsv = f32(rng.standard_normal((Bv, Sv, Ev)))
    gv = f32(rng.standard_normal((Bv, Sv, Dv)))
    ys, kinds = {}, []
    hooks = []
    for e, ex in enumerate(m.experts):
        kinds.append(type(ex).__name__)

        def hook(mod, inp, outp, e=e):
            outp.retain_grad()
            ys[e] = outp
        hooks.append(ex.register_forward_hook(hook))
    torch.randn_like = lambda t, **kw: epsv.to(t.dtype) if t.shape == epsv.shape else orig(t, **kw)
    try:
        out = m(xv)
        w_v, idx_v, _ = m.router(xv)                         # same noise -> same routing as inside forward
    finally:
        torch.randn_like = orig
    ((out * gv).sum() + 2.0 * m.get_aux_loss()).backward()
    N = Bv * Sv
    ys_all = torch.zeros(Ev, N, Dv)
    d_ys = torch.zeros(Ev, N, Dv)
    used = np.zeros(Ev, dtype=np.int64)
    for e, t in ys.items():
        ys_all[e] = t.detach().reshape(N, Dv)
        d_ys[e] = t.grad.reshape(N, Dv)
        used[e] = 1
    # router-path-only input gradient: experts fed a detached copy of x (their outputs do not depend on the router)
    for h in hooks:
        h.remove()
    m.zero_grad()
    xv2 = xv.detach().clone().requires_grad_()
    restore = []
    for ex in m.experts:
        f = ex.forward
        restore.append((ex, f))
        ex.forward = (lambda f: (lambda t, **kw: f(t.detach(), **kw)))(f)
    torch.randn_like = lambda t, **kw: epsv.to(t.dtype) if t.shape == epsv.shape else orig(t, **kw)
    try:
        out2 = m(xv2)
    finally:
        torch.randn_like = orig
    ((out2 * gv).sum() + 2.0 * m.get_aux_loss()).backward()
    for ex, f in restore:
        ex.forward = f
    g = grads_of(m)
    save("vqa_moe_layer", cfg=np.array([Bv, Sv, Dv, Fv, Ev, Kv]), expert_kinds=np.array(kinds), used=used,
         router_sd=m.router.state_dict(), norm_sd=m.output_norm.state_dict(), x=xv, eps=epsv, gout=gv, w=w_v,
         idx=idx_v, probs=m.aux_outputs["router_probs"], loss=m.get_aux_loss(), ys=ys_all, d_ys=d_ys, out=out,
         d_x_router=xv2.grad,
         grads={k: v for k, v in g.items() if k.startswith("router.") or k.startswith("output_norm.")})

    # ---- fingerprints at the benchmark dimensions -----------------------------------------------------------------
    D, H, F = 768, 8, 2048
    # A2: MultimodalFusion cross_attention, cfg1/2 shapes (T=64, V=50, L=2), 8 samples
    B, T, V, L = 8, 64, 50, 2
    m = MultimodalFusion(FusionConfig(fusion_type="cross_attention", hidden_dim=D, output_dim=D, num_heads=H,
                                      num_layers=L, dropout=0.0, use_layer_norm=True))
    sd = rnd_state_dict(m, 31)
    m.load_state_dict(sd)
    m.train()
    vis, txt, gout = seeded_inputs(32, (B, V, D), (B, T, D), (B, D))
    lens = np.random.default_rng(33).integers(8, T + 1, size=B)
    valid = torch.arange(T)[None, :] < torch.tensor(lens)[:, None]
    vis.requires_grad_(); txt.requires_grad_()
    out = m(vis, txt, text_mask=~valid)
    (out * gout).sum().backward()
    tensors = {"d_visual": vis.grad, "d_text": txt.grad}
    tensors.update({f"grads/{k}": v for k, v in grads_of(m).items()})
    save("fp_multimodal_fusion_d768", cfg=np.array([B, T, V, D, H, L]), lens=lens, out=out, sd_keys=np.array(list(sd)),
         **fingerprint(tensors))

    # A6/A7: MOELayer on [32,114,768] (the cfg5 token layout), E=8, top-2, F=2048
    B, S, E, K = 32, 114, 8, 2
    m = MOELayer(input_dim=D, hidden_dim=F, output_dim=D, num_experts=E, top_k=K, dropout=0.0)
    sd = rnd_state_dict(m, 41)
    m.load_state_dict(sd)
    m.train()
    x, gout = seeded_inputs(42, (B, S, D), (B, S, D))
    x.requires_grad_()
    out = m(x)
    ((out * gout).sum() + 2.0 * m.get_aux_loss()).backward()
    _, idx, _ = m.router(x.detach())
    tensors = {"out": out, "d_x": x.grad, "probs": m.aux_outputs["router_probs"]}
    tensors.update({f"grads/{k}": v for k, v in grads_of(m).items()})
    save("fp_moe_layer_d768", cfg=np.array([B, S, D, F, E, K]), loss=m.get_aux_loss(), idx_sha256=sha_int(idx),
         idx_head=idx.reshape(-1, K)[:64], sd_keys=np.array(list(sd)), **fingerprint(tensors))

    # A4: CrossModalFusion + standard MOE layer, cfg5 shapes (V=50, Tq=64 -> 114 tokens), 4 samples
    B, V, Tq, E = 4, 50, 64, 8
    cfg = GenerativeVQAConfig()
    cfg.fusion_dim, cfg.fusion_num_heads, cfg.fusion_num_layers, cfg.fusion_dropout = D, H, 2, 0.0
    cfg.decoder_ff_dim, cfg.use_moe, cfg.moe_type, cfg.moe_position = F, True, "standard", "fusion"
    cfg.num_experts, cfg.num_experts_per_token = E, 2
    m = CrossModalFusion(cfg)
    sd = rnd_state_dict(m, 51)
    m.load_state_dict(sd)
    m.train()
    vis, q, gout = seeded_inputs(52, (B, V, D), (B, Tq, D), (B, V + Tq, D))
    lens = np.random.default_rng(53).integers(8, Tq + 1, size=B)
    qvalid = torch.arange(Tq)[None, :] < torch.tensor(lens)[:, None]
    vis.requires_grad_(); q.requires_grad_()
    out, aux = m(vis, q, qvalid.long())
    (out * gout).sum().backward()
    tensors = {"out": out, "d_visual": vis.grad, "d_question": q.grad}
    tensors.update({f"grads/{k}": v for k, v in grads_of(m).items()})
    save("fp_cross_modal_fusion_d768", cfg=np.array([B, V, Tq, D, H, F, E]), lens=lens, aux=np.array(aux),
         sd_keys=np.array(list(sd)), **fingerprint(tensors))


def main_n4():
    """GatedLinearExpert banks and HierarchicalMOE, from the reference's own modules."""
    torch.manual_seed(0)
    torch.set_num_threads(8)
    from src.modeling.moe.moe_layer import MOELayer, HierarchicalMOE
    rng = np.random.default_rng(20261020)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)

    # ---- MOELayer over GatedLinearExperts ----------------------------------------------------------------------------
    B, S, D, F, E, K = 3, 20, 64, 96, 4, 2
    m = MOELayer(input_dim=D, hidden_dim=F, output_dim=D, num_experts=E, top_k=K, expert_type="glu", dropout=0.0)
    sd = rnd_state_dict(m, 71)
    m.load_state_dict(sd)
    m.train()
    x = f32(rng.standard_normal((B, S, D))).requires_grad_()
    gout = f32(rng.standard_normal((B, S, D)))
    out = m(x)
    ((out * gout).sum() + 2.0 * m.get_aux_loss()).backward()
    save("glu_moe_layer", cfg=np.array([B, S, D, F, E, K]), sd=sd, x=x, gout=gout, out=out, loss=m.get_aux_loss(),
         d_x=x.grad, grads=grads_of(m))

    # ---- HierarchicalMOE, homogeneous FFN groups -----------------------------------------------------------------------
    B, S, D, F, G, Epg, Kg, Ke = 3, 16, 64, 96, 3, 2, 2, 2
    m = HierarchicalMOE(input_dim=D, hidden_dim=F, output_dim=D, num_expert_groups=G, experts_per_group=Epg,
                        top_k_groups=Kg, top_k_experts=Ke, dropout=0.0, expert_types=["feedforward"] * G)
    sd = rnd_state_dict(m, 72)
    m.load_state_dict(sd)
    m.train()
    x = f32(rng.standard_normal((B, S, D))).requires_grad_()
    gout = f32(rng.standard_normal((B, S, D)))
    out = m(x)
    ((out * gout).sum() + 2.0 * m.get_aux_loss()).backward()
    save("hierarchical_moe_ffn", cfg=np.array([B, S, D, F, G, Epg, Kg, Ke]), sd=sd, x=x, gout=gout, out=out,
         loss=m.get_aux_loss(), d_x=x.grad, grads=grads_of(m))

    # ---- HierarchicalMOE, default (heterogeneous) groups: expert bodies recorded as data -------------------------------
    B, S, D, F, G, Epg, Kg, Ke = 2, 6, 64, 96, 4, 2, 2, 1
    m = HierarchicalMOE(input_dim=D, hidden_dim=F, output_dim=D, num_expert_groups=G, experts_per_group=Epg,
                        top_k_groups=Kg, top_k_experts=Ke, dropout=0.0)
    torch.manual_seed(9)
    for p in m.parameters():
        if p.dim() >= 2:
            torch.nn.init.normal_(p, std=1.0 / np.sqrt(p.shape[-1]))
    kinds = [type(e).__name__ for grp in m.expert_groups for e in grp]
    m.train()
    x = f32(rng.standard_normal((B, S, D))).requires_grad_()
    gout = f32(rng.standard_normal((B, S, D)))
    ys, hooks = {}, []
    flat = [e for grp in m.expert_groups for e in grp]
    for i, ex in enumerate(flat):
        def hook(mod, inp, outp, i=i):
            if i not in ys:                       # the reference re-runs an expert per (slot, group): identical outputs
                outp.retain_grad()
                ys[i] = []
            ys[i].append(outp)
        hooks.append(ex.register_forward_hook(hook))
    out = m(x)
    ((out * gout).sum() + 2.0 * m.get_aux_loss()).backward()
    N = B * S
    ys_all, d_ys, used = torch.zeros(G * Epg, N, D), torch.zeros(G * Epg, N, D), np.zeros(G * Epg, dtype=np.int64)
    for i, lst in ys.items():
        ys_all[i] = lst[0].detach().reshape(N, D)
        used[i] = 1
    for h in hooks:
        h.remove()
    # router/combine-only gradients: second pass with the experts fed a detached input and every call's output gradient
    # summed per expert
    m.zero_grad()
    x2 = x.detach().clone().requires_grad_()
    leaves = {i: ys_all[i].view(B, S, D).clone().requires_grad_() for i in range(G * Epg)}
    restore = []
    for i, ex in enumerate(flat):
        f = ex.forward
        restore.append((ex, f))
        ex.forward = (lambda i: (lambda t, **kw: leaves[i]))(i)
    out2 = m(x2)
    ((out2 * gout).sum() + 2.0 * m.get_aux_loss()).backward()
    for ex, f in restore:
        ex.forward = f
    for i in range(G * Epg):
        if leaves[i].grad is not None:
            d_ys[i] = leaves[i].grad.reshape(N, D)
    g = grads_of(m)
    keep = {k: v for k, v in g.items() if not k.startswith("expert_groups.")}
    rsd = {k: v for k, v in m.state_dict().items() if not k.startswith("expert_groups.")}
    assert float((out2 - out).abs().max()) < 1e-6
    save("hierarchical_moe_default", cfg=np.array([B, S, D, F, G, Epg, Kg, Ke]), expert_kinds=np.array(kinds), used=used,
         sd=rsd, x=x, gout=gout, ys=ys_all, d_ys=d_ys, out=out, loss=m.get_aux_loss(), d_x_router=x2.grad, grads=keep)


def main_n3():
    """QFormerFusion and SingleStreamFusion from the reference's fusion registry."""
    torch.manual_seed(0)
    torch.set_num_threads(8)
    from src.modeling.fusion.fusion_approaches import create_fusion_model
    rng = np.random.default_rng(20261021)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    B, V, T, D, H, L, I = 2, 36, 20, 64, 4, 2, 128            # token counts of examples/fusion_examples.py:24-29
    vis = f32(rng.standard_normal((B, V, D)))
    txt = f32(rng.standard_normal((B, T, D)))
    tvalid = torch.ones(B, T, dtype=torch.bool)
    tvalid[:, -5:] = False
    vvalid = torch.ones(B, V, dtype=torch.bool)
    vvalid[1, -7:] = False
    gout = f32(rng.standard_normal((B, D)))
    for name, kw, seed in (("qformer", dict(num_query_tokens=8, num_attention_heads=H, num_layers=L, intermediate_dim=I), 81),
                           ("single_stream", dict(num_attention_heads=H, num_layers=L, intermediate_dim=I,
                                                  max_vision_tokens=40, max_text_tokens=24), 82)):
        m = create_fusion_model(name, vision_dim=D, text_dim=D, output_dim=D, dropout=0.0, **kw)
        sd = rnd_state_dict(m, seed)
        m.load_state_dict(sd)
        m.train()
        v, t = vis.clone().requires_grad_(), txt.clone().requires_grad_()
        out = m(v, t, vision_mask=vvalid, text_mask=tvalid)
        (out * gout).sum().backward()
        save(f"{name}_fusion", cfg=np.array([B, V, T, D, H, L, I]), sd=sd, vision=v, text=t, vision_valid=vvalid,
             text_valid=tvalid, gout=gout, out=out, d_vision=v.grad, d_text=t.grad, grads=grads_of(m))


def main_n2():
    """the reference's TransformerDecoder (tied 97-way vocabulary, 2 layers) and its label-smoothed
    cross-entropy with ignored positions, forward + backward."""
    torch.manual_seed(0)
    torch.set_num_threads(8)
    from src.modeling.meta_arch.generative_vqa_model import GenerativeVQAConfig, TransformerDecoder
    rng = np.random.default_rng(20261022)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    B, T, S, D, H, L, F, V = 3, 9, 11, 64, 4, 2, 128, 97
    cfg = GenerativeVQAConfig(hidden_size=D, num_decoder_layers=L, num_attention_heads=H, decoder_ff_dim=F,
                              decoder_dropout=0.0, max_answer_length=16, vocab_size=V, tie_word_embeddings=True,
                              label_smoothing=0.1)
    emb = torch.nn.Embedding(V, D)
    m = TransformerDecoder(cfg, embedding=emb)
    sd = rnd_state_dict(m, 91)
    sd["output_projection.weight"] = sd["embedding.weight"]          # tied
    sd["pos_encoding.pe"] = m.state_dict()["pos_encoding.pe"]        # the sinusoid buffer is not random
    m.load_state_dict(sd)
    m.train()
    memory = f32(rng.standard_normal((B, S, D))).requires_grad_()
    ids = torch.tensor(rng.integers(0, V, size=(B, T)), dtype=torch.long)
    mem_mask = torch.ones(B, S)
    mem_mask[1, -3:] = 0
    tgt_mask = torch.ones(B, T)
    tgt_mask[0, -2:] = 0
    tgt_mask[2, -4:] = 0
    labels = torch.tensor(rng.integers(0, V, size=(B, T)), dtype=torch.long)
    labels[tgt_mask == 0] = -100
    logits = m(memory, ids, encoder_attention_mask=mem_mask, decoder_attention_mask=tgt_mask)
    loss = torch.nn.CrossEntropyLoss(ignore_index=-100, label_smoothing=0.1)(logits.view(-1, V), labels.view(-1))
    loss.backward()
    save("generative_decoder", cfg=np.array([B, T, S, D, H, L, F, V]), sd=sd, memory=memory.detach(), ids=ids,
         mem_mask=mem_mask, tgt_mask=tgt_mask, labels=labels, logits=logits, loss=loss, d_memory=memory.grad,
         grads=grads_of(m))


if __name__ == "__main__" and "--n2" in sys.argv:
    main_n2()
if __name__ == "__main__" and "--n3" in sys.argv:
    main_n3()
if __name__ == "__main__" and "--r2" in sys.argv:
    main_r2()
if __name__ == "__main__" and "--n4" in sys.argv:
    main_n4()
