"""MOE package: same public names as the reference's src/modeling/moe (moe/__init__.py:6-49) for the
components on the hot path."""
from .config import ExpertConfig, MOEConfig, RouterConfig, VQAMOEConfig
from .experts import (BaseExpert, ExpertWithCapacity, FeedForwardExpert, GatedLinearExpert, create_expert,
                      register_expert_type)
from .layers import HierarchicalMOE, MOELayer, SparseMOELayer, VQAMOELayer
from .router import BaseRouter, ExpertChoiceRouter, NoisyTopKRouter, SoftRouter, TopKRouter, create_router
from .utils import (ExpertDropout, ExpertParallelWrapper, analyze_routing_patterns, compute_expert_capacity,
                    compute_expert_entropy, compute_load_balance_loss, compute_router_z_loss,
                    get_expert_utilization, load_moe_checkpoint, save_moe_checkpoint)

__all__ = [
    "BaseExpert", "ExpertWithCapacity", "FeedForwardExpert", "GatedLinearExpert", "create_expert",
    "register_expert_type",
    "BaseRouter", "TopKRouter", "SoftRouter", "NoisyTopKRouter", "ExpertChoiceRouter", "create_router",
    "MOELayer", "SparseMOELayer", "VQAMOELayer", "HierarchicalMOE",
    "MOEConfig", "ExpertConfig", "RouterConfig", "VQAMOEConfig",
    "compute_expert_capacity", "compute_load_balance_loss", "compute_router_z_loss", "get_expert_utilization",
    "compute_expert_entropy", "ExpertDropout", "ExpertParallelWrapper", "save_moe_checkpoint",
    "load_moe_checkpoint", "analyze_routing_patterns",
]
