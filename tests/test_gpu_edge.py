"""GPU: edge cases of the MOE / fusion path — single token, top-1, many experts, experts that receive nothing,
unaligned token counts, no masks, fully valid masks, query length 1, large-V attention tiles, noisy routing in train
mode, SparseMOE capacity in train mode."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import reference_port as rp
from oracle import routing_np

pytestmark = pytest.mark.gpu

from vqa_model_builder_b200 import fusion, moe, ops  # noqa: E402
import vqa_model_builder_b200 as pkg  # noqa: E402

DEV = "cuda"


def oracle_moe(layer, x, E, K, **kw):
    sd = {k: v.detach().cpu().clone() for k, v in layer.state_dict().items()}
    return rp.moe_layer(sd, x.detach().cpu(), E, K, **kw)


@pytest.mark.parametrize("B,S,D,F,E,K", [(1, 1, 64, 128, 4, 2), (3, 7, 64, 96, 8, 1), (2, 33, 128, 256, 64, 4),
                                         (5, 13, 768, 2048, 32, 2), (1, 2, 64, 128, 16, 2)])
def test_moe_layer_shapes(B, S, D, F, E, K):
    torch.manual_seed(B * 100 + E)
    layer = moe.MOELayer(input_dim=D, hidden_dim=F, output_dim=D, num_experts=E, top_k=K, dropout=0.0).to(DEV).train()
    x = torch.randn(B, S, D, device=DEV, requires_grad=True)
    out = layer(x)
    ref, loss, probs, w, idx = oracle_moe(layer, x, E, K)
    assert rel_err(out, ref) < 1e-4, rel_err(out, ref)
    assert abs(float(layer.get_aux_loss()) - float(loss)) < 1e-6
    got_idx = layer.last_plan.idx.view(-1, K).cpu().numpy()
    _, amb = routing_np.topk_with_ties(probs.reshape(-1, E).numpy(), K)
    assert (np.sort(got_idx, -1)[~amb] == np.sort(idx.reshape(-1, K).numpy(), -1)[~amb]).all()
    (out.square().mean() + layer.get_aux_loss()).backward()
    assert torch.isfinite(x.grad).all()
    # experts that received no token get an exactly-zero weight gradient
    counts = layer.last_plan.counts.cpu().numpy()
    for e, c in enumerate(counts):
        if c == 0:
            assert float(layer.experts[e].fc1.weight.grad.abs().max()) == 0.0


def test_moe_non_residual_experts_and_activations():
    """input_dim != output_dim drops the expert residual (expert_types.py:76); relu / silu / tanh activations."""
    for act in ("relu", "silu", "tanh"):
        torch.manual_seed(1)
        layer = moe.MOELayer(input_dim=64, hidden_dim=96, output_dim=128, num_experts=4, top_k=2, dropout=0.0)
        for i in range(4):
            layer.experts[i] = moe.FeedForwardExpert(64, 96, 128, expert_id=i, dropout=0.0, activation=act)
        layer = layer.to(DEV)
        x = torch.randn(2, 9, 64, device=DEV)
        out = layer(x)
        sd = {k: v.detach().cpu() for k, v in layer.state_dict().items()}
        ref, *_ = rp.moe_layer(sd, x.cpu(), 4, 2, act=act)
        assert out.shape == (2, 9, 128)
        assert rel_err(out, ref) < 1e-4, (act, rel_err(out, ref))


def test_sparse_moe_train_mode_noise_and_capacity():
    torch.manual_seed(3)
    E, K, D = 4, 2, 64
    layer = moe.SparseMOELayer(input_dim=D, hidden_dim=128, output_dim=D, num_experts=E, top_k=K, capacity_factor=0.5,
                               dropout=0.0).to(DEV).train()
    x = torch.randn(4, 16, D, device=DEV, requires_grad=True)
    noise = torch.randn(4, 16, E, device=DEV)
    orig = layer.router.forward
    layer.router.forward = lambda t, **kw: orig(t, noise=noise)
    out = layer(x)
    sd = {k: v.detach().cpu() for k, v in layer.state_dict().items()}
    ref, *_ = rp.sparse_moe_layer(sd, x.detach().cpu(), E, K, capacity_factor=0.5, noise=noise.cpu(), noise_std=1.0)
    assert rel_err(out, ref) < 1e-4, rel_err(out, ref)
    out.sum().backward()
    assert torch.isfinite(x.grad).all() and layer.router.w_noise.weight.grad is not None


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_fusion_without_masks_and_query_length_one(mode):
    pkg.set_compute_dtype(mode)
    try:
        torch.manual_seed(0)
        D, H = 128, 4
        m = fusion.MultimodalFusion(fusion.FusionConfig("cross_attention", D, D, H, 1, 0.0, True)).to(DEV)
        for T, V in ((1, 1), (5, 3), (64, 257), (128, 50)):
            vis = torch.randn(2, V, D, device=DEV)
            txt = torch.randn(2, T, D, device=DEV)
            out = m(vis, txt)                         # no masks at all
            full = torch.zeros(2, T, dtype=torch.bool, device=DEV)
            out2 = m(vis, txt, text_mask=full)        # all-valid mask == no mask
            assert torch.equal(out, out2)
            sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
            ref = rp.multimodal_fusion(sd, "cross_attention", H, 1, True, vis.cpu(), txt.cpu())
            assert rel_err(out, ref) < (1e-4 if mode == "fp32" else 2e-2), (T, V, rel_err(out, ref))
    finally:
        pkg.set_compute_dtype("auto")


def test_cross_attention_fusion_vision_queries_longer_than_tile():
    """Vision stream as queries with V = 257 > 128 rows: three query tiles on the tensor-core attention."""
    torch.manual_seed(0)
    D, H = 128, 4
    pkg.set_compute_dtype("bf16")
    try:
        m = fusion.CrossAttentionFusion(D, D, D, H, 1, 256, 0.0, "add").to(DEV)
        vis = torch.randn(2, 257, D, device=DEV)
        txt = torch.randn(2, 20, D, device=DEV)
        out = m(vis, txt)
        sd = {k: v.detach().cpu().to(torch.bfloat16).float() if v.dim() >= 2 else v.detach().cpu()
              for k, v in m.state_dict().items()}
        ref = rp.cross_attention_fusion(sd, H, 1, "add", vis.cpu().to(torch.bfloat16).float(),
                                        txt.cpu().to(torch.bfloat16).float())
        assert rel_err(out, ref) < 2e-2, rel_err(out, ref)
    finally:
        pkg.set_compute_dtype("auto")


def test_deepcopy_to_device_and_state_dict_roundtrip():
    import copy
    torch.manual_seed(0)
    layer = moe.MOELayer(input_dim=64, hidden_dim=128, output_dim=64, num_experts=4, top_k=2, dropout=0.0).to(DEV)
    x = torch.randn(2, 8, 64, device=DEV)
    y0 = layer(x)
    clone = copy.deepcopy(layer)
    assert torch.equal(clone(x), y0)
    sd = {k: v.cpu() for k, v in layer.state_dict().items()}
    fresh = moe.MOELayer(input_dim=64, hidden_dim=128, output_dim=64, num_experts=4, top_k=2, dropout=0.0)
    fresh.load_state_dict(sd)
    assert torch.equal(fresh.to(DEV)(x), y0)
    # an optimiser step through the slab-backed Parameters changes the output
    opt = torch.optim.SGD(layer.parameters(), lr=0.1)
    layer(x).square().mean().backward()
    opt.step()
    assert not torch.equal(layer(x), y0)


def test_prefetch_compute_weights_matches_in_line_cast():
    """runtime.prefetch_compute_weights: the bf16 weight copy refreshed on the side stream is the one the next forward
    uses (after an optimizer-style in-place update of the fp32 masters), eagerly and across repeated steps."""
    torch.manual_seed(0)
    pkg.set_compute_dtype("bf16")
    try:
        layer = moe.MOELayer(input_dim=256, hidden_dim=512, output_dim=256, num_experts=4, top_k=2, dropout=0.0).to(DEV).eval()
        x = torch.randn(4, 9, 256, device=DEV)
        y0 = layer(x).float()
        for step in range(3):
            with torch.no_grad():
                for p in layer.parameters():
                    p.mul_(1.01)                      # what an optimizer step does: in-place update of the masters
            assert pkg.prefetch_compute_weights(layer) == 1
            assert pkg.prefetch_compute_weights(layer) == 0          # already queued
            y1 = layer(x).float()                                    # waits for the side-stream copy
            layer.invalidate()
            y2 = layer(x).float()                                    # in-line cast of the same masters
            assert torch.equal(y1, y2), step
            assert not torch.equal(y1, y0)
    finally:
        pkg.set_compute_dtype("auto")
