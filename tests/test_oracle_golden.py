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
        if float(g.norm()) == 0.0:
            assert float(got.norm()) < 1e-6, k
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
