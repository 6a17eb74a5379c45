"""Host-side MOE helpers with the reference's names and return conventions
(src/modeling/moe/moe_utils.py:12-341).  These are monitoring / bookkeeping utilities, not hot-path code;
they are vectorised (one device->host read instead of one per expert or per token)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import nn


def compute_expert_capacity(num_tokens: int, num_experts: int, top_k: int, capacity_factor: float = 1.25) -> int:
    return int(capacity_factor * num_tokens * top_k / num_experts)


def _expert_counts(expert_indices: torch.Tensor, num_experts: int) -> torch.Tensor:
    flat = expert_indices.reshape(-1)
    flat = flat[(flat >= 0) & (flat < num_experts)]
    return torch.bincount(flat.to(torch.int64), minlength=num_experts).to(torch.float32)


def compute_load_balance_loss(router_probs: torch.Tensor, expert_indices: torch.Tensor, num_experts: int,
                              weight: float = 0.01) -> torch.Tensor:
    """weight * E * sum_e (tokens routed to e / N) * (mean prob of e)   (moe_utils.py:35-76)."""
    n_tok = router_probs.shape[0] * router_probs.shape[1]
    frac = _expert_counts(expert_indices, num_experts).to(router_probs.device) / n_tok
    mean_p = router_probs.reshape(n_tok, num_experts).mean(dim=0)
    return weight * num_experts * torch.sum(frac * mean_p)


def compute_router_z_loss(router_logits: torch.Tensor, weight: float = 0.001) -> torch.Tensor:
    return weight * torch.logsumexp(router_logits, dim=-1).pow(2).mean()


def get_expert_utilization(expert_indices: torch.Tensor, num_experts: int) -> Dict[int, float]:
    total = expert_indices.numel()
    counts = _expert_counts(expert_indices, num_experts).tolist()
    return {e: counts[e] / total for e in range(num_experts)}


def compute_expert_entropy(router_probs: torch.Tensor) -> torch.Tensor:
    return (-(router_probs * torch.log(router_probs + 1e-10)).sum(dim=-1)).mean()


class ExpertDropout(nn.Module):
    """Drops whole experts at train time and renormalises the surviving weights (moe_utils.py:142-191)."""

    def __init__(self, num_experts: int, drop_rate: float = 0.1):
        super().__init__()
        self.num_experts = num_experts
        self.drop_rate = drop_rate

    def forward(self, expert_weights: torch.Tensor, expert_indices: torch.Tensor
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        if not self.training or self.drop_rate == 0:
            return expert_weights, expert_indices
        alive = torch.bernoulli(torch.full((self.num_experts,), 1 - self.drop_rate, device=expert_indices.device))
        kept = expert_weights * alive[expert_indices]
        return kept / (kept.sum(dim=-1, keepdim=True) + 1e-10), expert_indices


class ExpertParallelWrapper(nn.Module):
    """API-compatible stand-in for the reference's device-placement wrapper (moe_utils.py:194-254): expert i
    lives on device_ids[i // (E / len(device_ids))] and activations are moved to it and back.  The B200-native
    expert parallelism (one process per GPU, NCCL all-to-all of permuted rows) is parallel.ExpertParallelMOE."""

    def __init__(self, experts: nn.ModuleList, device_ids: Optional[List[int]] = None):
        super().__init__()
        self.experts = experts
        self.num_experts = len(experts)
        if device_ids is None:
            device_ids = list(range(torch.cuda.device_count()))
        self.device_ids = device_ids
        if device_ids:
            per = max(1, self.num_experts // len(device_ids))
            for i, expert in enumerate(experts):
                expert.to(f"cuda:{device_ids[min(i // per, len(device_ids) - 1)]}")

    def forward(self, x: torch.Tensor, expert_id: int, **kwargs) -> torch.Tensor:
        expert = self.experts[expert_id]
        dev = next(expert.parameters()).device
        return expert(x.to(dev), **kwargs).to(x.device)


def save_moe_checkpoint(moe_layer: nn.Module, path: str, additional_info: Optional[Dict] = None):
    ckpt = {"state_dict": moe_layer.state_dict(), "num_experts": moe_layer.num_experts,
            "input_dim": moe_layer.input_dim, "output_dim": moe_layer.output_dim}
    if additional_info:
        ckpt["additional_info"] = additional_info
    torch.save(ckpt, path)


def load_moe_checkpoint(moe_layer: nn.Module, path: str, strict: bool = True) -> Dict:
    ckpt = torch.load(path, map_location="cpu")
    moe_layer.load_state_dict(ckpt["state_dict"], strict=strict)
    return ckpt.get("additional_info", {})


def analyze_routing_patterns(router_probs: torch.Tensor, expert_indices: torch.Tensor, num_experts: int
                             ) -> Dict[str, Any]:
    """Utilisation, entropy, extreme-probability means and the symmetric expert co-selection matrix."""
    flat = expert_indices.reshape(-1, expert_indices.size(-1)).to(torch.int64)
    K = flat.size(-1)
    co = torch.zeros(num_experts * num_experts, device=flat.device)
    for a in range(K):
        for b in range(a + 1, K):
            pair = flat[:, a] * num_experts + flat[:, b]
            co += torch.bincount(pair, minlength=num_experts * num_experts).to(co.dtype)
    co = co.view(num_experts, num_experts)
    co = co + co.t()
    return {
        "expert_utilization": get_expert_utilization(expert_indices, num_experts),
        "routing_entropy": compute_expert_entropy(router_probs).item(),
        "max_prob_mean": router_probs.max(dim=-1).values.mean().item(),
        "min_prob_mean": router_probs.min(dim=-1).values.mean().item(),
        "expert_co_selection": co.cpu().numpy().tolist(),
    }
