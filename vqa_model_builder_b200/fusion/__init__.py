"""Fusion package: the reference's `src/modeling/fusion` names plus the meta_arch fusion modules."""
from .approaches import (BaseFusion, CrossAttentionBlock, CrossAttentionFusion, QFormerFusion, QFormerLayer,
                         SingleStreamFusion, create_fusion_model)
from .cross_modal import CrossModalAttention, FusionConfig, MultimodalFusion
from .generative import CrossModalFusion, GenerativeFusionConfig

__all__ = ["BaseFusion", "CrossAttentionFusion", "CrossAttentionBlock", "create_fusion_model", "CrossModalAttention",
           "MultimodalFusion", "FusionConfig", "CrossModalFusion", "GenerativeFusionConfig", "QFormerFusion", "QFormerLayer",
           "SingleStreamFusion"]
