"""Experts of the MOE layer (reference: src/modeling/moe/base_expert.py, expert_types.py:14-92).

FeedForwardExpert is the homogeneous FFN expert the grouped tcgen05 GEMM targets.  Its nn.Linear /
nn.LayerNorm children are parameter containers (same state_dict keys and init as the reference); the
arithmetic runs in the library's kernels — grouped over all experts inside MOELayer, or through the dense
GEMM path when an expert is called on its own."""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import torch
from torch import nn

from .. import ops
from .._lib import ACT_CODES
from ..runtime import DropCtx, alloc_sites, resolve_compute_dtype


class BaseExpert(nn.Module):
    """base_expert.py:12-114: dims, expert_id and the usage buffers kept in the state_dict."""

    def __init__(self, input_dim: int = 768, hidden_dim: int = 3072, output_dim: int = 768,
                 expert_id: Optional[int] = None, dropout: float = 0.1):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.output_dim = output_dim
        self.expert_id = expert_id
        self.dropout_rate = dropout
        self.register_buffer("usage_count", torch.tensor(0.0))
        self.register_buffer("total_tokens", torch.tensor(0.0))

    def update_usage_stats(self, num_tokens: int):
        self.usage_count += 1
        self.total_tokens += num_tokens

    def get_usage_ratio(self) -> float:
        if self.total_tokens == 0:
            return 0.0
        return (self.usage_count / self.total_tokens).item()

    def reset_usage_stats(self):
        self.usage_count.zero_()
        self.total_tokens.zero_()

    def get_expert_info(self) -> Dict[str, Any]:
        return {"expert_id": self.expert_id, "input_dim": self.input_dim, "hidden_dim": self.hidden_dim,
                "output_dim": self.output_dim, "usage_ratio": self.get_usage_ratio(),
                "num_parameters": sum(p.numel() for p in self.parameters())}


class ExpertWithCapacity(BaseExpert):
    """base_expert.py:117-169."""

    def __init__(self, input_dim: int = 768, hidden_dim: int = 3072, output_dim: int = 768,
                 expert_id: Optional[int] = None, dropout: float = 0.1, capacity: Optional[int] = None):
        super().__init__(input_dim, hidden_dim, output_dim, expert_id, dropout)
        self.capacity = capacity

    def apply_capacity_constraint(self, x: torch.Tensor, routing_weights: torch.Tensor
                                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        if self.capacity is None or x.size(0) <= self.capacity:
            return x, routing_weights, torch.arange(x.size(0), device=x.device)
        _, keep = torch.topk(routing_weights, self.capacity)
        return x[keep], routing_weights[keep], keep


class FeedForwardExpert(BaseExpert):
    """LN( drop(fc2(drop(act(fc1 x)))) + x )  — residual only when input_dim == output_dim
    (expert_types.py:75-92)."""

    def __init__(self, input_dim: int = 768, hidden_dim: int = 3072, output_dim: int = 768,
                 expert_id: Optional[int] = None, dropout: float = 0.1, activation: str = "gelu"):
        super().__init__(input_dim, hidden_dim, output_dim, expert_id, dropout)
        acts = {"gelu": nn.GELU(), "relu": nn.ReLU(), "silu": nn.SiLU(), "tanh": nn.Tanh()}
        self.activation_name = activation if activation in acts else "gelu"
        self.activation = acts[self.activation_name]
        self.fc1 = nn.Linear(input_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, output_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(output_dim)
        self._sites = alloc_sites(2)

    @property
    def act_code(self) -> int:
        return ACT_CODES[self.activation_name]

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        lead = x.shape[:-1]
        cdt = resolve_compute_dtype(x)
        x2 = ops.to_compute(x.reshape(-1, x.shape[-1]), cdt)
        w1, w2 = self.fc1.weight, self.fc2.weight
        w1c = w1.detach() if cdt == torch.float32 else ops.cast(w1.detach(), cdt)
        w2c = w2.detach() if cdt == torch.float32 else ops.cast(w2.detach(), cdt)
        dc = DropCtx(self.training, float(self.dropout_rate), x.device, self._sites)
        h = ops.FFNFn.apply(x2, w1, self.fc1.bias, w2, self.fc2.bias, w1c, w2c, self.act_code, None, dc.site(0), None)
        if x.shape[-1] == self.output_dim:   # LN(dropout(h) + x)
            y = ops.AddLNFn.apply(x2, h, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps, dc.site(1))
        else:
            if dc.on:
                h = ops.DropoutFn.apply(h, dc.site(1))
            y = ops.AddLNFn.apply(h, None, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps, None)
        return ops.to_compute(y, x.dtype).view(*lead, self.output_dim)


class GatedLinearExpert(BaseExpert):
    """expert_types.py:448-515: [value | gate] = fc1(x) (2 * hidden wide); h = drop(value * sigmoid(gate));
    out = LN( drop(fc2 h) + x ) (residual only when input_dim == output_dim).  Token-wise, so MOELayer evaluates a bank
    of these through the same grouped GEMMs as FeedForwardExpert (ops.ExpertFFNFn with ACT_GLU)."""

    activation_name = "glu"

    def __init__(self, input_dim: int = 768, hidden_dim: int = 3072, output_dim: int = 768,
                 expert_id: Optional[int] = None, dropout: float = 0.1):
        super().__init__(input_dim, hidden_dim, output_dim, expert_id, dropout)
        self.fc1 = nn.Linear(input_dim, hidden_dim * 2)
        self.fc2 = nn.Linear(hidden_dim, output_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(output_dim)
        self._sites = alloc_sites(2)

    @property
    def act_code(self) -> int:
        return ops.ACT_GLU

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        lead = x.shape[:-1]
        cdt = resolve_compute_dtype(x)
        x2 = ops.to_compute(x.reshape(-1, x.shape[-1]), cdt)
        w1, w2 = self.fc1.weight, self.fc2.weight
        w1c = w1.detach() if cdt == torch.float32 else ops.cast(w1.detach(), cdt)
        w2c = w2.detach() if cdt == torch.float32 else ops.cast(w2.detach(), cdt)
        dc = DropCtx(self.training, float(self.dropout_rate), x.device, self._sites)
        pre = ops.LinearFn.apply(x2, w1, self.fc1.bias, w1c, None)
        h = ops.GLUFn.apply(pre, dc.site(0))
        h = ops.LinearFn.apply(h, w2, self.fc2.bias, w2c, None)
        if x.shape[-1] == self.output_dim:   # LN(dropout(h) + x)
            y = ops.AddLNFn.apply(x2, h, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps, dc.site(1))
        else:
            if dc.on:
                h = ops.DropoutFn.apply(h, dc.site(1))
            y = ops.AddLNFn.apply(h, None, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps, None)
        return ops.to_compute(y, x.dtype).view(*lead, self.output_dim)


_EXPERTS = {"feedforward": FeedForwardExpert, "glu": GatedLinearExpert}


def register_expert_type(name: str, cls) -> None:
    """Heterogeneous experts (vision/text/multimodal/specialised, expert_types.py:95-515) are out of the
    kernel scope; `install()` registers the reference's own classes here so create_expert keeps its API."""
    _EXPERTS[name] = cls


def create_expert(expert_type: str, input_dim: int, hidden_dim: int, output_dim: int,
                  expert_id: Optional[int] = None, **kwargs) -> BaseExpert:
    """expert_types.py:518-557."""
    if expert_type not in _EXPERTS:
        raise ValueError(f"Unknown expert type: {expert_type}. Available: {list(_EXPERTS.keys())}")
    return _EXPERTS[expert_type](input_dim=input_dim, hidden_dim=hidden_dim, output_dim=output_dim,
                                 expert_id=expert_id, **kwargs)
