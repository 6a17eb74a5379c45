"""Routers of the MOE layer behind the reference's router API (src/modeling/moe/router.py).

TopKRouter / NoisyTopKRouter run the fused sm_100a router kernel (gate dot products in fp32, softmax,
top-k, renormalisation, load-balance statistics in one pass).  Return convention is the reference's:
(routing_weights [B,S,K] float32, expert_indices [B,S,K] int64, aux_outputs dict)."""
from __future__ import annotations

import inspect
from typing import Any, Dict, Optional, Tuple

import torch
from torch import nn

from .. import ops
from ..runtime import resolve_compute_dtype


class BaseRouter(nn.Module):
    """Common state: gate = Linear(input_dim, num_experts, bias=False)  (router.py:19-40)."""

    def __init__(self, input_dim: int, num_experts: int, top_k: int = 2):
        super().__init__()
        self.input_dim = input_dim
        self.num_experts = num_experts
        self.top_k = top_k
        self.gate = nn.Linear(input_dim, num_experts, bias=False)

    def forward(self, x: torch.Tensor, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError

    def _compute_routing_logits(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 gate logits via the library GEMM path (used by the soft / expert-choice routers)."""
        lead = x.shape[:-1]
        x2 = ops.to_compute(x.reshape(-1, x.shape[-1]), torch.float32)
        w = self.gate.weight
        y = ops.LinearFn.apply(x2, w, None, w.detach(), None)
        return y.view(*lead, self.num_experts)


class _FusedTopK(BaseRouter):
    """Shared implementation of TopKRouter and NoisyTopKRouter."""

    noise_std: float = 0.0

    def _route(self, x: torch.Tensor, noisy: bool, eps: Optional[torch.Tensor] = None):
        if x.dim() != 3:
            raise ValueError("router expects [batch, seq, dim]")
        B, S, D = x.shape
        cdt = resolve_compute_dtype(x)
        x2 = ops.to_compute(x.reshape(B * S, D), cdt)
        w_noise = None
        if noisy:
            w_noise = self.w_noise.weight
            if eps is None:
                # same generator consumption as the reference's torch.randn_like(clean_logits) (router.py:308)
                eps = torch.randn(B, S, self.num_experts, device=x.device, dtype=torch.float32)
            eps = eps.reshape(B * S, self.num_experts).to(torch.float32)
        lb = float(self.load_balance_weight) if self.use_aux_loss else 0.0
        w, idx32, loss, probs, nsm, _ts, _cnt = ops.RouterFn.apply(
            x2, self.gate.weight, w_noise, eps, float(self.noise_std), lb, int(self.top_k),
            getattr(self, "stats_group", None))
        idx = idx32.to(torch.int64).view(B, S, self.top_k)
        aux: Dict[str, Any] = {}
        if self.use_aux_loss:
            aux["load_balance_loss"] = loss.reshape(())
            aux["router_probs"] = probs.view(B, S, self.num_experts)
        # int32 copy for the dispatch kernels; only trusted while `idx` is the tensor handed onwards
        aux["_b200_idx32"] = (idx, idx32)
        return w.view(B, S, self.top_k), idx, aux, nsm


class TopKRouter(_FusedTopK):
    """router.py:76-178."""

    def __init__(self, input_dim: int, num_experts: int, top_k: int = 2, use_aux_loss: bool = True,
                 load_balance_weight: float = 0.01):
        super().__init__(input_dim, num_experts, top_k)
        self.use_aux_loss = use_aux_loss
        self.load_balance_weight = load_balance_weight

    def forward(self, x: torch.Tensor, **kwargs) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, Any]]:
        w, idx, aux, _ = self._route(x, noisy=False)
        return w, idx, aux


class NoisyTopKRouter(_FusedTopK):
    """router.py:251-366: train-time logits += randn * softplus(w_noise x) * noise_std; aux on clean logits."""

    def __init__(self, input_dim: int, num_experts: int, top_k: int = 2, noise_std: float = 1.0,
                 use_aux_loss: bool = True, load_balance_weight: float = 0.01):
        super().__init__(input_dim, num_experts, top_k)
        self.noise_std = noise_std
        self.use_aux_loss = use_aux_loss
        self.load_balance_weight = load_balance_weight
        self.w_noise = nn.Linear(input_dim, num_experts, bias=False)

    def forward(self, x: torch.Tensor, noise: Optional[torch.Tensor] = None, **kwargs):
        """`noise` optionally injects the N(0,1) draw (parity tests feed both sides the same tensor)."""
        noisy = self.training
        w, idx, aux, nsm = self._route(x, noisy=noisy, eps=noise if noisy else None)
        if self.use_aux_loss:
            aux["noise_scale"] = nsm.reshape(()) if noisy else 0.0
        return w, idx, aux


class SoftRouter(BaseRouter):
    """router.py:181-248: all experts, softmax(logits / temperature) weights."""

    def __init__(self, input_dim: int, num_experts: int, temperature: float = 1.0):
        super().__init__(input_dim, num_experts, num_experts)
        self.temperature = temperature

    def forward(self, x: torch.Tensor, **kwargs):
        logits = self._compute_routing_logits(x) / self.temperature
        weights = torch.softmax(logits, dim=-1)
        idx = torch.arange(self.num_experts, device=x.device).expand(x.size(0), x.size(1), -1)
        ent = -(weights * torch.log(weights + 1e-10)).sum(-1).mean()
        return weights, idx, {"router_probs": weights, "entropy": ent}


class ExpertChoiceRouter(BaseRouter):
    """router.py:369-449: experts pick their top-`capacity` tokens (softmax over the token axis)."""

    def __init__(self, input_dim: int, num_experts: int, capacity_factor: float = 1.25):
        super().__init__(input_dim, num_experts, 1)
        self.capacity_factor = capacity_factor

    def forward(self, x: torch.Tensor, **kwargs):
        B, S, _ = x.shape
        n_tok = B * S
        capacity = int(self.capacity_factor * n_tok / self.num_experts)
        scores = torch.softmax(self._compute_routing_logits(x), dim=1)
        flat = scores.reshape(n_tok, self.num_experts)
        chosen = torch.zeros(n_tok, dtype=torch.long, device=x.device)
        weights = torch.zeros(n_tok, device=x.device)
        for e in range(self.num_experts):  # later experts overwrite earlier ones, as in the reference
            top_s, top_i = torch.topk(flat[:, e], min(capacity, n_tok), dim=0)
            chosen[top_i] = e
            weights[top_i] = top_s
        return weights.view(B, S, 1), chosen.view(B, S, 1), {"router_probs": scores, "capacity": capacity}


_ROUTERS = {"topk": TopKRouter, "soft": SoftRouter, "noisy_topk": NoisyTopKRouter,
            "expert_choice": ExpertChoiceRouter}


def create_router(router_type: str, input_dim: int, num_experts: int, **kwargs) -> BaseRouter:
    """Factory with kwarg filtering (router.py:452-494): unknown keys for the chosen class are dropped."""
    if router_type not in _ROUTERS:
        raise ValueError(f"Unknown router type: {router_type}. Available: {list(_ROUTERS.keys())}")
    cls = _ROUTERS[router_type]
    accepted = set(inspect.signature(cls.__init__).parameters) - {"self"}
    return cls(input_dim, num_experts, **{k: v for k, v in kwargs.items() if k in accepted})
