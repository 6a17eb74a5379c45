"""CPU: host-side logic of the drop-in modules — constructor/state_dict contract against the reference's key
layout (recorded in the golden fixtures), factories, config plumbing, parameter slab, helper utilities."""
import copy

import numpy as np
import pytest
import torch

from conftest import load_golden
from vqa_model_builder_b200 import fusion, moe
from vqa_model_builder_b200.slab import ParamSlab


def test_state_dict_keys_match_reference_multimodal_fusion():
    g = load_golden("multimodal_fusion_xattn")
    m = fusion.MultimodalFusion(fusion.FusionConfig(hidden_dim=64, output_dim=64, num_heads=4, num_layers=2))
    assert set(m.state_dict()) == set(g["sd"])
    m.load_state_dict(g["sd"])
    for k, v in m.state_dict().items():
        assert v.shape == g["sd"][k].shape


def test_state_dict_keys_match_reference_cross_attention_fusion():
    g = load_golden("cross_attention_fusion")
    m = fusion.CrossAttentionFusion(64, 64, 64, 4, 2, 128)
    assert set(m.state_dict()) == set(g["sd"])
    m.load_state_dict(g["sd"])


def test_state_dict_keys_match_reference_moe_layers():
    g = load_golden("moe_layer")
    m = moe.MOELayer(input_dim=64, hidden_dim=128, output_dim=64, num_experts=4, top_k=2)
    assert set(m.state_dict()) == set(g["sd"])
    m.load_state_dict(g["sd"])
    s = moe.SparseMOELayer(input_dim=64, hidden_dim=128, output_dim=64, num_experts=4, top_k=2)
    assert set(s.state_dict()) == set(load_golden("sparse_moe_layer")["sd"])


def test_state_dict_keys_match_reference_cross_modal_fusion():
    g = load_golden("cross_modal_fusion_moe")
    cfg = fusion.GenerativeFusionConfig(fusion_dim=64, fusion_num_heads=4, decoder_ff_dim=128, use_moe=True,
                                        num_experts=4)
    m = fusion.CrossModalFusion(cfg)
    assert set(m.state_dict()) == set(g["sd"])
    m.load_state_dict(g["sd"])
    assert isinstance(m.moe_layer, moe.MOELayer) and m.moe_type == "standard"


def test_moe_config_overrides_and_attributes():
    rc = moe.RouterConfig(router_type="noisy_topk", load_balance_weight=0.05)
    m = moe.MOELayer(config=moe.MOEConfig(input_dim=32, hidden_dim=64, output_dim=32, num_experts=4,
                                          num_experts_per_token=1, router_config=rc))
    assert isinstance(m.router, moe.NoisyTopKRouter) and m.router.load_balance_weight == 0.05
    assert (m.input_dim, m.hidden_dim, m.output_dim, m.num_experts, m.top_k) == (32, 64, 32, 4, 1)
    assert len(m.experts) == 4 and type(m.experts[0]).__name__ == "FeedForwardExpert"
    assert m.get_expert_usage() == {0: 0.0, 1: 0.0, 2: 0.0, 3: 0.0}
    assert float(m.get_aux_loss()) == 0.0
    assert m._homogeneous()
    m.experts[1] = torch.nn.Identity()
    assert not m._homogeneous()


def test_create_router_filters_kwargs_and_rejects_unknown():
    r = moe.create_router("topk", 16, 4, top_k=2, noise_std=0.3, use_aux_loss=False)
    assert isinstance(r, moe.TopKRouter) and r.use_aux_loss is False
    r = moe.create_router("expert_choice", 16, 4, top_k=2, capacity_factor=2.0)
    assert isinstance(r, moe.ExpertChoiceRouter) and r.capacity_factor == 2.0
    with pytest.raises(ValueError):
        moe.create_router("top_k", 16, 4)      # the reference's vqa_config default is not a registry key either
    with pytest.raises(ValueError):
        moe.create_expert("nope", 8, 8, 8)
    with pytest.raises(ValueError):
        fusion.create_fusion_model("mcan")
    assert type(fusion.create_fusion_model("qformer", vision_dim=32, text_dim=32, output_dim=32, num_query_tokens=4,
                                           num_attention_heads=2, num_layers=1, intermediate_dim=64)).__name__ == "QFormerFusion"


def test_vqa_moe_layer_accepts_explicit_heterogeneous_experts():
    ex = [torch.nn.Linear(16, 16), torch.nn.Sequential(torch.nn.Linear(16, 16), torch.nn.Tanh())]
    m = moe.VQAMOELayer(input_dim=16, hidden_dim=32, output_dim=16, top_k=1, experts=ex)
    assert m.num_experts == 2 and isinstance(m.router, moe.NoisyTopKRouter) and not m._homogeneous()
    with pytest.raises(RuntimeError):
        moe.VQAMOELayer(input_dim=16, hidden_dim=32, output_dim=16)


def test_param_slab_packs_views_and_survives_moves():
    m = moe.MOELayer(input_dim=32, hidden_dim=64, output_dim=32, num_experts=4)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    slab = ParamSlab(m._slab_groups())
    slab.ensure(torch.device("cpu"))
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k])
    ex = list(m.experts)
    assert slab.contiguous_run([e.fc1.weight for e in ex]) and slab.contiguous_run([e.layer_norm.bias for e in ex])
    stacked = slab.span(ex[0].fc1.weight, 4 * 64 * 32, torch.float32).view(4, 64, 32)
    assert stacked.data_ptr() == ex[0].fc1.weight.data_ptr()
    assert torch.equal(stacked[2], ex[2].fc1.weight)
    with torch.no_grad():
        ex[3].fc1.weight.add_(1.0)          # optimiser-style in-place update is visible through the slab
    assert torch.equal(stacked[3], ex[3].fc1.weight)
    assert slab._packed()
    m2 = copy.deepcopy(m)
    assert "_slab" not in m2.__dict__
    m.double().float()                       # .to()-style re-allocation breaks the packing; ensure() repairs it
    assert not slab._packed()
    slab.ensure(torch.device("cpu"))
    assert slab._packed()


def test_moe_utils():
    idx = torch.tensor([[[0, 1], [1, 2]], [[1, 0], [3, 1]]])
    probs = torch.softmax(torch.arange(16.0).view(2, 2, 4) / 7, dim=-1)
    assert moe.compute_expert_capacity(64, 8, 2, 1.25) == 20
    util = moe.get_expert_utilization(idx, 4)
    assert util == {0: 0.25, 1: 0.5, 2: 0.125, 3: 0.125}
    frac = torch.tensor([2.0, 4.0, 1.0, 1.0]) / 4
    want = 0.01 * 4 * torch.sum(frac * probs.view(4, 4).mean(0))
    assert torch.allclose(moe.compute_load_balance_loss(probs, idx, 4), want)
    a = moe.analyze_routing_patterns(probs, idx, 4)
    co = np.array(a["expert_co_selection"])
    assert co[0, 1] == 2 and co[1, 0] == 2 and co[1, 2] == 1 and co[1, 3] == 1 and co.sum() == 8
    d = moe.ExpertDropout(4, 0.5).eval()
    w = torch.rand(2, 2, 2)
    assert d(w, idx)[0] is w


def test_compute_dtype_policy():
    import vqa_model_builder_b200 as pkg
    x32, x16 = torch.zeros(1), torch.zeros(1, dtype=torch.bfloat16)
    assert pkg.resolve_compute_dtype(x32) == torch.float32 and pkg.resolve_compute_dtype(x16) == torch.bfloat16
    pkg.set_compute_dtype("bf16")
    try:
        assert pkg.resolve_compute_dtype(x32) == torch.bfloat16
    finally:
        pkg.set_compute_dtype("auto")
    with pytest.raises(ValueError):
        pkg.set_compute_dtype("fp8")


def test_answer_head_state_dict_layout_matches_reference_sequential():
    """AnswerHead keeps the reference's nn.Sequential layout (vqa_model.py:451-465): Linear / ReLU / Dropout triples,
    so checkpoints keyed classifier.{0,3,6,...} load unchanged."""
    from vqa_model_builder_b200 import heads
    cfg = heads.AnswerHeadConfig(num_answers=3001, hidden_dims=[512, 256], dropout=0.3)
    head = heads.AnswerHead(cfg, 768)
    assert list(head.state_dict().keys()) == ["classifier.0.weight", "classifier.0.bias", "classifier.3.weight",
                                              "classifier.3.bias", "classifier.6.weight", "classifier.6.bias"]
    assert head.classifier[6].weight.shape == (3001, 256)
    with pytest.raises(RuntimeError):          # CUDA-only product path: no silent CPU fallback
        head(torch.randn(2, 768))
