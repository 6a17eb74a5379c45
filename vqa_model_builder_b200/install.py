"""Swap the B200-native modules into the reference package (`src.modeling...`), symbol by symbol.

The reference has no plugin mechanism: its pipelines import classes by name.  `install()` rebinds exactly the names
SURVEY 8(b) lists, so `python -m src.core.vqa_pipeline --use-moe` (fusion=cross_attention) and the generative
pipeline build our modules with their own, unchanged, construction code.  `uninstall()` restores the originals.

    import vqa_model_builder_b200.install as b200_install
    b200_install.install()          # requires the reference on PYTHONPATH (import name `src`)
"""
from __future__ import annotations

import importlib
from typing import Dict, List, Tuple

from . import decoder as _decoder
from . import fusion as _fusion
from . import heads as _heads
from . import moe as _moe

_saved: List[Tuple[object, str, object]] = []


def _bind(module, name: str, value) -> None:
    _saved.append((module, name, getattr(module, name, None)))
    setattr(module, name, value)


def install(verbose: bool = False) -> Dict[str, str]:
    """Rebind the reference's fusion / MOE symbols to the B200 implementations.  Returns {qualified name: class}."""
    if _saved:
        return {}
    ref_moe = importlib.import_module("src.modeling.moe")
    ref_layer = importlib.import_module("src.modeling.moe.moe_layer")
    ref_router = importlib.import_module("src.modeling.moe.router")
    ref_experts = importlib.import_module("src.modeling.moe.expert_types")
    ref_special = importlib.import_module("src.modeling.moe.specialized_experts")
    ref_vqa = importlib.import_module("src.modeling.meta_arch.vqa_model")
    ref_gen = importlib.import_module("src.modeling.meta_arch.generative_vqa_model")
    ref_fusion = importlib.import_module("src.modeling.fusion")
    ref_fusion_impl = importlib.import_module("src.modeling.fusion.fusion_approaches")

    # heterogeneous expert bodies stay the reference's own PyTorch modules (out of kernel scope)
    _moe.VQAMOELayer.expert_factories = {
        "vision": ref_experts.VisionExpert, "text": ref_experts.TextExpert, "multimodal": ref_experts.MultimodalExpert,
        "specialized": [ref_special.SegmentationExpert, ref_special.ObjectDetectionExpert, ref_special.OCRExpert,
                        ref_special.SceneUnderstandingExpert],
    }
    for kind, cls in (("vision", ref_experts.VisionExpert), ("text", ref_experts.TextExpert),
                      ("multimodal", ref_experts.MultimodalExpert)):      # 'feedforward' and 'glu' are native
        _moe.register_expert_type(kind, cls)

    done: Dict[str, str] = {}

    def bind_all(mod, names):
        for n, v in names.items():
            if hasattr(mod, n):
                _bind(mod, n, v)
                done[f"{mod.__name__}.{n}"] = f"{v.__module__}.{getattr(v, '__name__', v)}"

    moe_syms = {"MOELayer": _moe.MOELayer, "SparseMOELayer": _moe.SparseMOELayer, "VQAMOELayer": _moe.VQAMOELayer,
                "TopKRouter": _moe.TopKRouter, "NoisyTopKRouter": _moe.NoisyTopKRouter,
                "create_router": _moe.create_router, "FeedForwardExpert": _moe.FeedForwardExpert,
                "GatedLinearExpert": _moe.GatedLinearExpert, "HierarchicalMOE": _moe.HierarchicalMOE,
                "create_expert": _moe.create_expert}
    bind_all(ref_moe, moe_syms)              # `from src.modeling.moe import VQAMOELayer` (vqa_model.py:529)
    bind_all(ref_layer, moe_syms)
    bind_all(ref_router, {k: moe_syms[k] for k in ("TopKRouter", "NoisyTopKRouter", "create_router")})
    bind_all(ref_experts, {k: moe_syms[k] for k in ("FeedForwardExpert", "GatedLinearExpert", "create_expert")})
    bind_all(ref_vqa, {"MultimodalFusion": _fusion.MultimodalFusion,           # vqa_model.py:503
                       "CrossModalAttention": _fusion.CrossModalAttention,     # vqa_model.py:331
                       "AnswerHead": _heads.AnswerHead})                       # vqa_model.py:436
    bind_all(ref_gen, {"MOELayer": _moe.MOELayer, "VQAMOELayer": _moe.VQAMOELayer,   # generative_vqa_model.py:23
                       "SparseMOELayer": _moe.SparseMOELayer, "CrossModalFusion": _fusion.CrossModalFusion,
                       "TransformerDecoder": _decoder.TransformerDecoder,            # :342
                       "PositionalEncoding": _decoder.PositionalEncoding})           # :453
    fus_syms = {"CrossAttentionFusion": _fusion.CrossAttentionFusion, "CrossAttentionBlock": _fusion.CrossAttentionBlock,
                "QFormerFusion": _fusion.QFormerFusion, "QFormerLayer": _fusion.QFormerLayer,
                "SingleStreamFusion": _fusion.SingleStreamFusion, "create_fusion_model": _fusion.create_fusion_model}
    bind_all(ref_fusion, fus_syms)          # the whole registry (fusion_approaches.py:719-725) is native
    bind_all(ref_fusion_impl, fus_syms)
    if verbose:
        for k, v in done.items():
            print(f"[b200] {k} -> {v}")
    return done


def uninstall() -> None:
    while _saved:
        mod, name, old = _saved.pop()
        if old is None:
            delattr(mod, name)
        else:
            setattr(mod, name, old)
    _moe.VQAMOELayer.expert_factories = {}
