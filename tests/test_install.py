"""CPU, only where the reference tree is present (this container): install() rebinds the reference's symbols and
the reference's own construction code then builds our modules."""
import os
import sys

import pytest

REF = os.environ.get("B200VQA_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "modeling")),
                                reason="reference tree not available")


@pytest.fixture()
def installed():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from vqa_model_builder_b200 import install as inst
    done = inst.install()
    yield done
    inst.uninstall()
    sys.path.remove(REF)


def test_install_rebinds_and_reference_code_builds_our_modules(installed):
    from vqa_model_builder_b200 import fusion, moe
    import src.modeling.meta_arch.generative_vqa_model as gen
    import src.modeling.meta_arch.vqa_model as vqa
    import src.modeling.moe as ref_moe
    from src.modeling.meta_arch.vqa_config import FusionConfig
    assert "src.modeling.meta_arch.vqa_model.MultimodalFusion" in installed
    assert vqa.MultimodalFusion is fusion.MultimodalFusion and ref_moe.VQAMOELayer is moe.VQAMOELayer
    assert gen.MOELayer is moe.MOELayer and gen.CrossModalFusion is fusion.CrossModalFusion
    # the reference's FusionConfig dataclass drives our MultimodalFusion unchanged
    m = vqa.MultimodalFusion(FusionConfig(hidden_dim=64, output_dim=64, num_heads=4, num_layers=1))
    assert isinstance(m, fusion.MultimodalFusion) and len(m.fusion_layers) == 1
    # `--use-moe` path: VQAMOELayer built with the reference's heterogeneous expert classes, our router
    layer = ref_moe.VQAMOELayer(input_dim=64, hidden_dim=128, output_dim=64, num_vision_experts=1, num_text_experts=1,
                                num_multimodal_experts=1, num_specialized_experts=1, top_k=2)
    assert isinstance(layer, moe.VQAMOELayer) and isinstance(layer.router, moe.NoisyTopKRouter)
    assert [type(e).__name__ for e in layer.experts] == ["VisionExpert", "TextExpert", "MultimodalExpert",
                                                         "SegmentationExpert"]
    assert type(layer.experts[0]).__module__.startswith("src.modeling.moe")
    # generative fusion built from the reference's own config object
    cfg = gen.GenerativeVQAConfig()
    cfg.fusion_dim, cfg.fusion_num_heads, cfg.decoder_ff_dim, cfg.use_moe, cfg.num_experts = 64, 4, 128, True, 4
    f = gen.CrossModalFusion(cfg)
    assert isinstance(f, fusion.CrossModalFusion) and isinstance(f.moe_layer, moe.MOELayer)
    # answer head: the reference's AnswerHeadConfig drives ours; same state_dict keys as the reference's own class
    from vqa_model_builder_b200 import heads
    from src.modeling.meta_arch.vqa_config import AnswerHeadConfig
    assert vqa.AnswerHead is heads.AnswerHead
    head = vqa.AnswerHead(AnswerHeadConfig(num_answers=37), 64)
    assert isinstance(head, heads.AnswerHead) and head.classifier[-1].out_features == 37
    # the reference's GenerativeVQAModel constructs our decoder (generative_vqa_model.py:503) with its
    # own config object and the shared answer embedding; state_dict keys equal the reference class's
    from vqa_model_builder_b200 import decoder as b200_decoder
    assert gen.TransformerDecoder is b200_decoder.TransformerDecoder
    dcfg = gen.GenerativeVQAConfig(hidden_size=64, num_decoder_layers=1, num_attention_heads=4, decoder_ff_dim=128,
                                   vocab_size=50, max_answer_length=12)
    import torch
    emb = torch.nn.Embedding(50, 64)
    dec = gen.TransformerDecoder(dcfg, embedding=emb)
    assert isinstance(dec, b200_decoder.TransformerDecoder) and dec.output_projection.weight is emb.weight
    # registry: cross_attention -> ours, others stay the reference's
    import src.modeling.fusion as ref_fusion
    assert isinstance(ref_fusion.create_fusion_model("cross_attention", vision_dim=64, text_dim=64, output_dim=64,
                                                     num_attention_heads=4, num_layers=1, intermediate_dim=128),
                      fusion.CrossAttentionFusion)


def test_uninstall_restores():
    sys.path.insert(0, REF)
    try:
        from vqa_model_builder_b200 import install as inst
        import src.modeling.meta_arch.vqa_model as vqa
        orig = vqa.MultimodalFusion
        inst.install()
        assert vqa.MultimodalFusion is not orig
        inst.uninstall()
        assert vqa.MultimodalFusion is orig
    finally:
        sys.path.remove(REF)
