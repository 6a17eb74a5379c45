"""CPU: the oracle restatement (oracle/reference_port.py, oracle/routing_np.py) against the golden vectors
produced by the reference's own modules (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import torch

from conftest import load_golden, rel_err
from oracle import reference_port as rp
from oracle import routing_np

TOL = 2e-5  # fp32 CPU vs fp32 CPU, different op order


def _leafs(sd):
    return {k: v.clone().requires_grad_(v.dtype.is_floating_point and v.dim() > 0) for k, v in sd.items()}


def _check_grads(sd, golden_grads, tol=TOL):
    for k, g in golden_grads.items():
        got = sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])
        if float(g.norm()) < 1e-5:        # (numerically) zero gradient, e.g. a top-1 router whose two experts got the
            assert float((got - g).norm()) < 1e-5, k      # same number of tokens: compare absolutely
        else:
            assert rel_err(got, g) < tol, (k, rel_err(got, g))


def test_multimodal_fusion_cross_attention():
    g = load_golden("multimodal_fusion_xattn")
    B, T, V, D, H, L = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    vis = g["visual"].clone().requires_grad_()
    txt = g["text"].clone().requires_grad_()
    out = rp.multimodal_fusion(sd, "cross_attention", H, L, True, vis, txt, None, ~g["text_valid"])
    assert rel_err(out, g["out"]) < TOL
    (out * g["gout"]).sum().backward()
    assert rel_err(vis.grad, g["d_visual"]) < TOL
    assert rel_err(txt.grad, g["d_text"]) < TOL
    _check_grads(sd, g["grads"])


def test_multimodal_fusion_other_branches():
    for ft in ("concat", "add"):
        g = load_golden(f"multimodal_fusion_{ft}")
        out = rp.multimodal_fusion(g["sd"], ft, 4, 2, True, g["visual"], g["text"])
        assert rel_err(out, g["out"]) < TOL, ft


def test_cross_attention_fusion():
    g = load_golden("cross_attention_fusion")
    B, T, V, D, H, L, I = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    vis = g["vision"].clone().requires_grad_()
    txt = g["text"].clone().requires_grad_()
    out = rp.cross_attention_fusion(sd, H, L, "concat", vis, txt, None, g["text_valid"])
    assert rel_err(out, g["out"]) < TOL
    (out * g["gout"]).sum().backward()
    assert rel_err(vis.grad, g["d_vision"]) < TOL
    assert rel_err(txt.grad, g["d_text"]) < TOL
    _check_grads(sd, g["grads"])


def test_topk_router():
    g = load_golden("topk_router")
    B, S, D, E, K = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    x = g["x"].clone().requires_grad_()
    w, idx, loss, probs, _ = rp.topk_router(sd, "", x, K, 0.01)
    assert torch.equal(idx, g["idx"])
    assert rel_err(w, g["w"]) < TOL and rel_err(probs, g["probs"]) < TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-7
    ((w * g["gw"]).sum() + 3.0 * loss).backward()
    assert rel_err(x.grad, g["d_x"]) < TOL
    _check_grads(sd, g["grads"])


def test_noisy_router():
    g = load_golden("noisy_router")
    B, S, D, E, K = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    x = g["x"].clone().requires_grad_()
    w, idx, loss, probs, _ = rp.topk_router(sd, "", x, K, 0.01, noise=g["eps"], noise_std=1.0)
    assert torch.equal(idx, g["idx"])
    assert rel_err(w, g["w"]) < TOL and rel_err(probs, g["probs"]) < TOL
    ((w * g["gw"]).sum() + 3.0 * loss).backward()
    assert rel_err(x.grad, g["d_x"]) < TOL
    _check_grads(sd, g["grads"])


def test_moe_layer():
    g = load_golden("moe_layer")
    B, S, D, F, E, K = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    x = g["x"].clone().requires_grad_()
    out, loss, probs, w, idx = rp.moe_layer(sd, x, E, K)
    assert rel_err(out, g["out"]) < TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-7
    ((out * g["gout"]).sum() + 2.0 * loss).backward()
    assert rel_err(x.grad, g["d_x"]) < TOL
    _check_grads(sd, g["grads"])


def test_sparse_moe_layer_capacity():
    g = load_golden("sparse_moe_layer")
    B, S, D, F, E, K = [int(v) for v in g["cfg"]]
    out, loss, probs, w, idx = rp.sparse_moe_layer(g["sd"], g["x"], E, K, capacity_factor=float(g["capacity_factor"]))
    assert rel_err(out, g["out"]) < TOL
    # the capacity limit really is active in this fixture
    counts = np.bincount(idx.reshape(-1).numpy(), minlength=E)
    assert counts.max() > int(float(g["capacity_factor"]) * B * S * K / E)


def test_cross_modal_fusion_with_moe():
    g = load_golden("cross_modal_fusion_moe")
    B, V, T, D, H, F, E = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    vis = g["visual"].clone().requires_grad_()
    q = g["question"].clone().requires_grad_()
    out, aux = rp.cross_modal_fusion(sd, H, 2, vis, q, g["question_valid"], moe=dict(num_experts=E, top_k=2))
    assert rel_err(out, g["out"]) < TOL
    assert abs(float(aux) - float(g["aux"])) < 1e-6
    (out * g["gout"]).sum().backward()
    assert rel_err(vis.grad, g["d_visual"]) < TOL
    assert rel_err(q.grad, g["d_question"]) < TOL
    _check_grads(sd, g["grads"])


def test_routing_plan_matches_reference_dispatch_order():
    """Canonical map == per-expert nonzero() order of SparseMOELayer (moe_layer.py:326) on the golden routing."""
    g = load_golden("moe_layer")
    B, S, D, F, E, K = [int(v) for v in g["cfg"]]
    _, idx, _, _, _ = rp.topk_router({"gate.weight": g["sd"]["router.gate.weight"]}, "", g["x"], K)
    flat = idx.reshape(-1, K)
    plan = routing_np.routing_plan(flat.numpy(), E)
    pos = 0
    for e in range(E):
        toks = (flat == e).any(dim=-1).nonzero(as_tuple=True)[0]
        for t in toks.tolist():
            k = int((flat[t] == e).nonzero()[0])
            assert plan["cmp_pos"][t * K + k] == pos
            pos += 1
        assert plan["cmp_off"][e + 1] == pos
    assert (plan["pad_off"] % 128 == 0).all()
    src = plan["row_src"]
    assert (np.sort(src[src >= 0]) == np.arange(B * S * K)).all()


def test_routing_plan_drops_masked_entries():
    idx = np.array([[0, 1], [-1, 1], [2, -1], [1, 0]])
    plan = routing_np.routing_plan(idx, 3)
    assert plan["counts"].tolist() == [2, 3, 1]
    assert plan["dest_row"][2] == -1 and plan["dest_row"][5] == -1
    assert plan["cmp_pos"].tolist() == [0, 2, -1, 3, 5, -1, 4, 1]


def test_capacity_keep():
    idx = np.array([[0], [0], [0], [1]])
    w = np.array([[0.2], [0.9], [0.5], [1.0]], dtype=np.float32)
    keep = routing_np.capacity_keep(idx, w, 2, capacity=2)
    assert keep.tolist() == [0, 1, 1, 1]


# ---- round-2 fixtures ---------------------------------------------------------------------------------------------
def test_sparse_moe_layer_train_mode_noise_capacity_backward():
    """A8 in train mode: injected router noise, active capacity limit, forward + backward."""
    g = load_golden("sparse_moe_layer_train")
    B, S, D, F, E, K = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    x = g["x"].clone().requires_grad_()
    out, loss, probs, w, idx = rp.sparse_moe_layer(sd, x, E, K, capacity_factor=float(g["capacity_factor"]),
                                                   noise=g["eps"], noise_std=1.0)
    counts = np.bincount(idx.reshape(-1).numpy(), minlength=E)
    assert counts.max() > int(float(g["capacity_factor"]) * B * S * K / E), "capacity limit must be active"
    assert rel_err(out, g["out"]) < TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-7
    ((out * g["gout"]).sum() + 2.0 * loss).backward()
    assert rel_err(x.grad, g["d_x"]) < TOL
    _check_grads(sd, g["grads"])


def test_vqa_moe_layer_router_and_combine():
    """A9: the reference's VQAMOELayer.  Its expert bodies are data here (recorded outputs); router (noisy, train mode)
    and the dense combine + output_norm are restated and must reproduce the layer output and every gradient that flows
    through them."""
    g = load_golden("vqa_moe_layer")
    B, S, D, F, E, K = [int(v) for v in g["cfg"]]
    assert list(g["expert_kinds"]) == ["VisionExpert"] * 2 + ["TextExpert"] * 2 + ["MultimodalExpert"] * 2 + \
        ["SegmentationExpert", "ObjectDetectionExpert"]
    sd = _leafs({"gate.weight": g["router_sd"]["gate.weight"], "w_noise.weight": g["router_sd"]["w_noise.weight"],
                 "nw": g["norm_sd"]["weight"], "nb": g["norm_sd"]["bias"]})
    x = g["x"].clone().requires_grad_()
    ys = g["ys"].view(E, B, S, D).clone().requires_grad_()
    w, idx, loss, probs, _ = rp.topk_router(sd, "", x, K, 0.01, noise=g["eps"], noise_std=1.0)
    assert torch.equal(idx, g["idx"])
    assert rel_err(w, g["w"]) < TOL and rel_err(probs, g["probs"]) < TOL
    out = rp.moe_combine_dense(ys, w, idx, sd["nw"], sd["nb"])
    assert rel_err(out, g["out"]) < TOL
    ((out * g["gout"]).sum() + 2.0 * loss).backward()
    assert rel_err(x.grad, g["d_x_router"]) < TOL
    used = g["used"].bool()
    assert rel_err(ys.grad.view(E, B * S, D)[used], g["d_ys"][used]) < TOL
    for k_ref, k in (("router.gate.weight", "gate.weight"), ("router.w_noise.weight", "w_noise.weight"),
                     ("output_norm.weight", "nw"), ("output_norm.bias", "nb")):
        assert rel_err(sd[k].grad, g["grads"][k_ref]) < TOL, k_ref


def _fp_check(tensors, g, tol):
    from oracle.fingerprint import compare
    from conftest import fp_flat
    errs = compare(tensors, fp_flat(g))
    worst = max((v, k) for k, v in errs.items())
    assert worst[0] < tol, worst


def test_fingerprint_multimodal_fusion_d768():
    """A2 at the benchmark dimensions (D=768, H=8, d_h=96, T=64, V=50): oracle vs the reference's fingerprint."""
    from conftest import seeded_normal
    from oracle.init_weights import multimodal_fusion_sd, seeded_state_dict
    g = load_golden("fp_multimodal_fusion_d768")
    B, T, V, D, H, L = [int(v) for v in g["cfg"]]
    sd = _leafs(seeded_state_dict(multimodal_fusion_sd(D, H, L), 31, keys=list(g["sd_keys"])))
    vis, txt, gout = seeded_normal(32, (B, V, D), (B, T, D), (B, D))
    valid = torch.arange(T)[None, :] < g["lens"][:, None]
    vis.requires_grad_(); txt.requires_grad_()
    out = rp.multimodal_fusion(sd, "cross_attention", H, L, True, vis, txt, None, ~valid)
    assert rel_err(out, g["out"]) < TOL
    (out * gout).sum().backward()
    tensors = {"d_visual": vis.grad, "d_text": txt.grad}
    tensors.update({f"grads/{k}": v.grad for k, v in sd.items()})
    _fp_check(tensors, g, 5e-5)


def test_fingerprint_moe_layer_d768():
    """A5-A7 at D=768 / F=2048 on [32,114,768]: routing indices by sha256, outputs and ALL gradients by probes."""
    from conftest import seeded_normal
    from oracle.fingerprint import sha_int
    from oracle.init_weights import moe_layer_sd, seeded_state_dict
    g = load_golden("fp_moe_layer_d768")
    B, S, D, F, E, K = [int(v) for v in g["cfg"]]
    tmpl = moe_layer_sd(D, F, E)
    for e in range(E):      # buffers of BaseExpert (base_expert.py:50-51) are part of the reference's key order
        tmpl[f"experts.{e}.usage_count"] = torch.tensor(0.0)
        tmpl[f"experts.{e}.total_tokens"] = torch.tensor(0.0)
    keys = list(g["sd_keys"])
    assert set(keys) == set(tmpl)
    sd = _leafs(seeded_state_dict(tmpl, 41, keys=keys))
    x, gout = seeded_normal(42, (B, S, D), (B, S, D))
    x.requires_grad_()
    out, loss, probs, w, idx = rp.moe_layer(sd, x, E, K)
    assert np.array_equal(sha_int(idx), g["idx_sha256"].numpy())
    assert abs(float(loss) - float(g["loss"])) < 1e-7
    ((out * gout).sum() + 2.0 * loss).backward()
    tensors = {"out": out, "d_x": x.grad, "probs": probs}
    tensors.update({f"grads/{k}": v.grad for k, v in sd.items() if v.requires_grad})
    _fp_check(tensors, g, 5e-5)


def test_fingerprint_cross_modal_fusion_d768():
    """A4 at the cfg5 shapes (V=50, Tq=64 -> 114 tokens, D=768, F=2048, 8 experts): oracle vs the reference."""
    from conftest import seeded_normal
    from oracle.init_weights import seeded_state_dict
    from vqa_model_builder_b200 import fusion       # parameter shapes only (construction needs no GPU)
    g = load_golden("fp_cross_modal_fusion_d768")
    B, V, Tq, D, H, F, E = [int(v) for v in g["cfg"]]
    cfg = fusion.GenerativeFusionConfig(fusion_dim=D, fusion_num_heads=H, fusion_num_layers=2, fusion_dropout=0.0,
                                        decoder_ff_dim=F, use_moe=True, moe_type="standard", num_experts=E,
                                        num_experts_per_token=2)
    tmpl = fusion.CrossModalFusion(cfg).state_dict()
    keys = list(g["sd_keys"])
    assert set(keys) == set(tmpl)
    sd = _leafs(seeded_state_dict(tmpl, 51, keys=keys))
    vis, q, gout = seeded_normal(52, (B, V, D), (B, Tq, D), (B, V + Tq, D))
    qvalid = torch.arange(Tq)[None, :] < g["lens"][:, None]
    vis.requires_grad_(); q.requires_grad_()
    out, aux = rp.cross_modal_fusion(sd, H, 2, vis, q, qvalid, moe=dict(num_experts=E, top_k=2))
    assert abs(float(aux) - float(g["aux"])) < 1e-6
    (out * gout).sum().backward()
    tensors = {"out": out, "d_visual": vis.grad, "d_question": q.grad}
    tensors.update({f"grads/{k}": v.grad for k, v in sd.items() if v.requires_grad})
    _fp_check(tensors, g, 5e-5)


# ---- GatedLinearExpert banks, HierarchicalMOE ---------------------------------------------------------
def test_glu_moe_layer():
    g = load_golden("glu_moe_layer")
    B, S, D, F, E, K = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    x = g["x"].clone().requires_grad_()
    out, loss, probs, w, idx = rp.moe_layer(sd, x, E, K, kind="glu")
    assert rel_err(out, g["out"]) < TOL
    ((out * g["gout"]).sum() + 2.0 * loss).backward()
    assert rel_err(x.grad, g["d_x"]) < TOL
    _check_grads(sd, g["grads"])


def test_hierarchical_moe_homogeneous_groups():
    g = load_golden("hierarchical_moe_ffn")
    B, S, D, F, G, Epg, Kg, Ke = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    x = g["x"].clone().requires_grad_()
    out, loss, _ = rp.hierarchical_moe(sd, x, G, Epg, Kg, Ke)
    assert rel_err(out, g["out"]) < TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    ((out * g["gout"]).sum() + 2.0 * loss).backward()
    assert rel_err(x.grad, g["d_x"]) < TOL
    _check_grads(sd, g["grads"])


def test_hierarchical_moe_default_groups_router_and_combine():
    g = load_golden("hierarchical_moe_default")
    B, S, D, F, G, Epg, Kg, Ke = [int(v) for v in g["cfg"]]
    assert list(g["expert_kinds"]) == ["VisionExpert"] * 2 + ["TextExpert"] * 2 + ["MultimodalExpert"] * 2 + \
        ["FeedForwardExpert"] * 2
    sd = _leafs(g["sd"])
    x = g["x"].clone().requires_grad_()
    ys = g["ys"].view(G * Epg, B, S, D).clone().requires_grad_()
    out, loss, _ = rp.hierarchical_moe(sd, x, G, Epg, Kg, Ke, ys=ys)
    assert rel_err(out, g["out"]) < TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    ((out * g["gout"]).sum() + 2.0 * loss).backward()
    assert rel_err(x.grad, g["d_x_router"]) < TOL
    used = g["used"].bool()
    assert rel_err(ys.grad.view(G * Epg, B * S, D)[used], g["d_ys"][used]) < TOL
    _check_grads(sd, g["grads"], tol=1e-4)     # aux-only gradients of the top-1 expert routers are ~1e-3 in norm


# ---- QFormerFusion, SingleStreamFusion ---------------------------------------------------------------
def test_qformer_fusion():
    g = load_golden("qformer_fusion")
    B, V, T, D, H, L, I = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    vis, txt = g["vision"].clone().requires_grad_(), g["text"].clone().requires_grad_()
    out = rp.qformer_fusion(sd, H, L, vis, txt, g["vision_valid"], g["text_valid"])
    assert rel_err(out, g["out"]) < TOL
    (out * g["gout"]).sum().backward()
    assert rel_err(vis.grad, g["d_vision"]) < TOL and rel_err(txt.grad, g["d_text"]) < TOL
    _check_grads(sd, g["grads"], tol=5e-5)


def test_single_stream_fusion():
    g = load_golden("single_stream_fusion")
    B, V, T, D, H, L, I = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    vis, txt = g["vision"].clone().requires_grad_(), g["text"].clone().requires_grad_()
    out = rp.single_stream_fusion(sd, H, L, vis, txt, g["vision_valid"], g["text_valid"])
    assert rel_err(out, g["out"]) < TOL
    (out * g["gout"]).sum().backward()
    assert rel_err(vis.grad, g["d_vision"]) < TOL and rel_err(txt.grad, g["d_text"]) < TOL
    _check_grads(sd, g["grads"], tol=5e-5)


def test_generative_decoder_and_smoothed_cross_entropy():
    """the restated TransformerDecoder (causal + padding masks, cross-attention to the fused memory,
    tied output projection) and label-smoothed CE with ignored positions against the reference's own run."""
    g = load_golden("generative_decoder")
    B, T, S, D, H, L, F, V = [int(v) for v in g["cfg"]]
    sd = _leafs(g["sd"])
    mem = g["memory"].clone().requires_grad_()
    logits = rp.transformer_decoder(sd, H, L, mem, g["ids"].long(), g["mem_mask"], g["tgt_mask"])
    assert rel_err(logits, g["logits"]) < TOL
    loss = rp.smoothed_cross_entropy(logits, g["labels"].long(), -100, 0.1)
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    assert rel_err(mem.grad, g["d_memory"]) < TOL
    tied = sd["embedding.weight"].grad + sd["output_projection.weight"].grad
    assert rel_err(tied, g["grads"]["embedding.weight"]) < TOL
    _check_grads(sd, {k: v for k, v in g["grads"].items() if k != "embedding.weight"})
