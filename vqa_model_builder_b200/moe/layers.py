"""MOE layers behind the reference's module API (src/modeling/moe/moe_layer.py).

MOELayer.forward keeps the reference's contract — call `self.router(x)` (whatever module or monkey-patched
callable that is), then combine expert outputs with the routing weights and apply `output_norm` — but for
homogeneous FeedForwardExperts the dense "every expert on every token" loop (moe_layer.py:151-168) becomes
plan -> permute -> grouped tcgen05 FFN -> weighted combine + LayerNorm, with no host synchronisation.
Heterogeneous experts stay PyTorch modules; only routing and the combine run in the library."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import nn

from .. import ops
from ..runtime import DropCtx, SlabOwner, alloc_sites, resolve_compute_dtype
from .config import MOEConfig
from .experts import FeedForwardExpert, GatedLinearExpert, create_expert
from .router import NoisyTopKRouter, create_router


class MOELayer(SlabOwner, nn.Module):
    """moe_layer.py:29-196."""

    def __init__(self, config: Optional[MOEConfig] = None, input_dim: int = 768, hidden_dim: int = 3072,
                 output_dim: int = 768, num_experts: int = 8, top_k: int = 2, router_type: str = "topk",
                 expert_type: str = "feedforward", dropout: float = 0.1, use_aux_loss: bool = True,
                 load_balance_weight: float = 0.01):
        nn.Module.__init__(self)
        if config is not None:  # config overrides dims/top_k/dropout/router settings but not expert_type
            input_dim, hidden_dim, output_dim = config.input_dim, config.hidden_dim, config.output_dim
            num_experts, top_k, dropout = config.num_experts, config.num_experts_per_token, config.expert_dropout
            if config.router_config:
                router_type = config.router_config.router_type
                use_aux_loss = config.router_config.use_aux_loss
                load_balance_weight = config.router_config.load_balance_weight
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.num_experts, self.top_k = num_experts, top_k
        self.router = create_router(router_type=router_type, input_dim=input_dim, num_experts=num_experts,
                                    top_k=top_k, use_aux_loss=use_aux_loss, load_balance_weight=load_balance_weight)
        self.experts = nn.ModuleList([
            create_expert(expert_type=expert_type, input_dim=input_dim, hidden_dim=hidden_dim, output_dim=output_dim,
                          expert_id=i, dropout=dropout) for i in range(num_experts)])
        self.output_norm = nn.LayerNorm(output_dim)
        self.aux_outputs: Dict[str, Any] = {}
        self.capacity_factor: Optional[float] = None   # set by SparseMOELayer

    # -- parameter slab -------------------------------------------------------------------------------------
    def _flat_experts(self) -> List[nn.Module]:
        return list(self.experts)

    def _homogeneous(self) -> bool:
        """all experts are token-wise FFN experts of one kind and shape (FeedForwardExpert or GatedLinearExpert)"""
        ex = self._flat_experts()
        if not ex or type(ex[0]) not in (FeedForwardExpert, GatedLinearExpert):
            return False
        e0 = ex[0]
        return all(type(e) is type(e0) and (e.input_dim, e.hidden_dim, e.output_dim, e.activation_name) ==
                   (e0.input_dim, e0.hidden_dim, e0.output_dim, e0.activation_name) for e in ex)

    def _slab_groups(self) -> List[List[Tuple[str, nn.Parameter]]]:
        ex = self._flat_experts()
        groups = []
        for attr in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "layer_norm.weight", "layer_norm.bias"):
            mod, leaf = attr.split(".")
            groups.append([(f"experts.{i}.{attr}", getattr(getattr(e, mod), leaf)) for i, e in enumerate(ex)])
        return groups

    def _expert_params(self) -> List[nn.Parameter]:
        ex = self._flat_experts()
        params: List[nn.Parameter] = []
        for attr in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "layer_norm.weight", "layer_norm.bias"):
            mod, leaf = attr.split(".")
            params.extend(getattr(getattr(e, mod), leaf) for e in ex)
        return params

    def _expert_stacks(self, device, cdt):
        """Stacked views of the expert parameters inside the slab (no copy)."""
        ex = self._flat_experts()
        E, D, F, Do = len(ex), ex[0].input_dim, ex[0].hidden_dim, ex[0].output_dim
        F1 = ex[0].fc1.out_features          # = F, or 2F for gated experts ([value | gate])
        slab = self._get_slab(device, cdt)
        return (slab.span(ex[0].fc1.weight, E * F1 * D, cdt).view(E, F1, D),
                slab.span(ex[0].fc1.bias, E * F1, torch.float32).view(E, F1),
                slab.span(ex[0].fc2.weight, E * Do * F, cdt).view(E, Do, F),
                slab.span(ex[0].fc2.bias, E * Do, torch.float32).view(E, Do),
                slab.span(ex[0].layer_norm.weight, E * Do, torch.float32).view(E, Do),
                slab.span(ex[0].layer_norm.bias, E * Do, torch.float32).view(E, Do))

    # -- forward ----------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        B, S, D = x.shape
        weights, indices, aux = self.router(x)
        self.aux_outputs = aux
        if self._homogeneous():
            return self._forward_grouped(x, weights, indices, aux)
        return self._forward_dense(x, weights, indices, mask, kwargs)

    def _plan(self, indices: torch.Tensor, aux: Dict[str, Any], num_experts: Optional[int] = None) -> "ops.RoutingPlan":
        stash = aux.get("_b200_idx32") if isinstance(aux, dict) else None
        if stash is not None and stash[0] is indices:
            idx32 = stash[1]
        else:
            idx32 = indices.reshape(-1, indices.shape[-1]).to(torch.int32)
        return ops.RoutingPlan(idx32, num_experts if num_experts is not None else self.num_experts)

    def _forward_grouped(self, x, weights, indices, aux, norm: bool = True) -> torch.Tensor:
        B, S, D = x.shape
        N = B * S
        K = indices.shape[-1]
        cdt = resolve_compute_dtype(x)
        ex = self._flat_experts()
        E = len(ex)
        F, Do = ex[0].hidden_dim, ex[0].output_dim
        w1s, b1s, w2s, b2s, lng, lnb = self._expert_stacks(x.device, cdt)
        x2 = ops.to_compute(x.reshape(N, D), cdt)
        plan = self._plan(indices, aux, E)
        w2d = weights.reshape(N, K).to(torch.float32)
        if self.capacity_factor is not None:
            capacity = int(self.capacity_factor * N * self.top_k / self.num_experts)
            w_eff, keep = plan.apply_capacity(w2d.detach(), capacity)
            # dropped (token, expert) pairs leave the graph: zero weight, zero gradient
            w2d = w2d * keep.view(N, K).to(w2d.dtype)
        params = self._expert_params()
        xp = ops.PermuteFn.apply(x2, plan.row_src, plan.pad_off, plan.dest_row, E, K, plan.Rmax)
        if "_sites" not in self.__dict__:
            self.__dict__["_sites"] = alloc_sites(2)
        dc = DropCtx(self.training, float(ex[0].dropout_rate), x.device, self.__dict__["_sites"])
        z = ops.ExpertFFNFn.apply(xp, plan.tile_group, plan.pad_off, (w1s, b1s, w2s, b2s, lng, lnb), ex[0].act_code,
                                  D == Do, ex[0].layer_norm.eps, (dc.site(0), dc.site(1)) if dc.on else None, *params)
        if norm:
            out = ops.CombineFn.apply(z, w2d, plan.dest_row, plan.row_src, self.output_norm.weight,
                                      self.output_norm.bias, self.output_norm.eps)
        else:       # plain weighted sum (HierarchicalMOE projects before it normalises)
            out = ops.CombineFn.apply(z, w2d, plan.dest_row, plan.row_src, None, None, 0.0)
        self.last_plan = plan
        return ops.to_compute(out, x.dtype).view(B, S, Do)

    def _forward_dense(self, x, weights, indices, mask, kwargs, norm: bool = True) -> torch.Tensor:
        """Heterogeneous experts see the whole sequence (they contain attention), exactly as in the reference;
        experts no token selected are skipped with ONE host read for all experts (reference: one per expert)."""
        B, S, D = x.shape
        N = B * S
        experts = self._flat_experts()
        E = len(experts)
        used = torch.zeros(E + 1, dtype=torch.bool, device=x.device)
        used[indices.reshape(-1).clamp(min=-1, max=E - 1) + 1] = True
        used = used[1:].tolist()
        outs = []
        zero = None
        for e, expert in enumerate(experts):
            if used[e]:
                outs.append(expert(x, mask=mask, **kwargs).reshape(N, -1))
            else:
                if zero is None:
                    zero = torch.zeros(N, self.output_dim, dtype=x.dtype, device=x.device)
                outs.append(zero)
        cdt = resolve_compute_dtype(x)
        ys = torch.stack([ops.to_compute(o.contiguous(), cdt) for o in outs], dim=0)
        if norm:
            out = ops.DenseCombineFn.apply(ys, weights.reshape(N, -1), indices.reshape(N, -1), self.output_norm.weight,
                                           self.output_norm.bias, self.output_norm.eps)
        else:
            out = ops.DenseCombineFn.apply(ys, weights.reshape(N, -1), indices.reshape(N, -1), None, None, 0.0)
        return ops.to_compute(out, x.dtype).view(B, S, self.output_dim)

    # -- reference accessors ------------------------------------------------------------------------------------
    def get_aux_loss(self) -> torch.Tensor:
        if "load_balance_loss" in self.aux_outputs:
            return self.aux_outputs["load_balance_loss"]
        return torch.tensor(0.0)

    def get_expert_usage(self) -> Dict[int, float]:
        return {i: e.get_usage_ratio() for i, e in enumerate(self.experts)}


class SparseMOELayer(MOELayer):
    """moe_layer.py:199-358: NoisyTopKRouter + capacity factor.  Same dispatch kernels; pairs beyond an
    expert's capacity (lowest combine weights) are masked instead of synchronising on counts."""

    def __init__(self, input_dim: int = 768, hidden_dim: int = 3072, output_dim: int = 768, num_experts: int = 8,
                 top_k: int = 2, capacity_factor: float = 1.25, dropout: float = 0.1, use_aux_loss: bool = True,
                 expert_type: str = "feedforward"):
        nn.Module.__init__(self)
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.num_experts, self.top_k = num_experts, top_k
        self.capacity_factor = capacity_factor
        self.router = NoisyTopKRouter(input_dim=input_dim, num_experts=num_experts, top_k=top_k,
                                      use_aux_loss=use_aux_loss)
        self.experts = nn.ModuleList([
            create_expert(expert_type=expert_type, input_dim=input_dim, hidden_dim=hidden_dim, output_dim=output_dim,
                          expert_id=i, dropout=dropout) for i in range(num_experts)])
        self.output_norm = nn.LayerNorm(output_dim)
        self.aux_outputs: Dict[str, Any] = {}

    def _compute_capacity(self, num_tokens: int) -> int:
        return int(self.capacity_factor * num_tokens * self.top_k / self.num_experts)


class VQAMOELayer(MOELayer):
    """moe_layer.py:551-692: NoisyTopKRouter over a heterogeneous expert list (vision / text / multimodal /
    specialised).  The expert bodies are outside the kernel scope (they are attention-bearing sequence modules);
    pass them in with `experts=[...]`, or let `install()` bind the reference's expert classes so the reference
    constructor signature builds them."""

    expert_factories: Dict[str, Any] = {}

    def __init__(self, input_dim: int = 768, hidden_dim: int = 3072, output_dim: int = 768,
                 num_vision_experts: int = 2, num_text_experts: int = 2, num_multimodal_experts: int = 2,
                 num_specialized_experts: int = 2, top_k: int = 2, dropout: float = 0.1,
                 vietnamese_optimized: bool = True, experts: Optional[List[nn.Module]] = None):
        nn.Module.__init__(self)
        total = num_vision_experts + num_text_experts + num_multimodal_experts + num_specialized_experts
        if experts is not None:
            total = len(experts)
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.num_experts, self.top_k = total, top_k
        self.capacity_factor = None
        self.router = NoisyTopKRouter(input_dim=input_dim, num_experts=total, top_k=top_k, use_aux_loss=True)
        if experts is None:
            experts = self._build_reference_experts(input_dim, hidden_dim, output_dim, num_vision_experts,
                                                    num_text_experts, num_multimodal_experts,
                                                    num_specialized_experts, dropout, vietnamese_optimized)
        self.experts = nn.ModuleList(experts)
        self.output_norm = nn.LayerNorm(output_dim)
        self.aux_outputs: Dict[str, Any] = {}

    @classmethod
    def _build_reference_experts(cls, input_dim, hidden_dim, output_dim, n_vis, n_txt, n_mm, n_spec, dropout,
                                 vietnamese_optimized) -> List[nn.Module]:
        f = cls.expert_factories
        need = ["vision", "text", "multimodal", "specialized"]
        if any(k not in f for k in need):
            raise RuntimeError(
                "VQAMOELayer needs the heterogeneous expert classes: call vqa_model_builder_b200.install() with the "
                "reference on PYTHONPATH, or pass experts=[...] explicitly")
        out: List[nn.Module] = []
        eid = 0
        common = dict(input_dim=input_dim, hidden_dim=hidden_dim, output_dim=output_dim, dropout=dropout)
        for kind, count in (("vision", n_vis), ("text", n_txt), ("multimodal", n_mm)):
            for _ in range(count):
                out.append(f[kind](expert_id=eid, **common))
                eid += 1
        spec = f["specialized"]  # ordered list cycled as in moe_layer.py:659-683
        for i in range(n_spec):
            klass = spec[i % len(spec)]
            kw = dict(common, expert_id=eid)
            if klass.__name__ == "OCRExpert":
                kw["vietnamese_optimized"] = vietnamese_optimized
            out.append(klass(**kw))
            eid += 1
        return out


class HierarchicalMOE(MOELayer):
    """moe_layer.py:361-548: two-level routing — a TopKRouter over expert GROUPS, then one TopKRouter per group over the
    experts of that group; out = LN(output_proj(sum over selected (group, expert) pairs of gw * ew * expert(x))).

    The reference walks four nested Python loops with a host sync (`.any()`) per (slot, group) and per (slot, expert) and
    re-runs whole experts for every pair.  Here every token's top_k_groups x top_k_experts (group, expert) pairs are
    flattened to expert ids g * experts_per_group + e with weights gw * ew and go through the same routing plan ->
    permute -> grouped FFN -> weighted combine as MOELayer when the experts are homogeneous FFN / GLU experts (one
    grouped GEMM over all groups); heterogeneous groups keep their PyTorch expert bodies and use the dense combine.
    The auxiliary loss keeps the reference's accounting: the group router's loss plus, for every (slot k, group g) that
    received at least one token, the loss of group g's expert router — evaluated on the device, without a host sync."""

    def __init__(self, input_dim: int = 768, hidden_dim: int = 3072, output_dim: int = 768, num_expert_groups: int = 4,
                 experts_per_group: int = 4, top_k_groups: int = 2, top_k_experts: int = 1, dropout: float = 0.1,
                 expert_types: Optional[List[str]] = None):
        nn.Module.__init__(self)
        from .router import TopKRouter
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.num_expert_groups, self.experts_per_group = num_expert_groups, experts_per_group
        self.top_k_groups, self.top_k_experts = top_k_groups, top_k_experts
        self.num_experts, self.top_k = num_expert_groups * experts_per_group, top_k_groups * top_k_experts
        self.capacity_factor = None
        if expert_types is None:
            expert_types = (["vision", "text", "multimodal", "feedforward"] * num_expert_groups)[:num_expert_groups]
        self.group_router = TopKRouter(input_dim=input_dim, num_experts=num_expert_groups, top_k=top_k_groups,
                                       use_aux_loss=True)
        self.expert_routers = nn.ModuleList([
            TopKRouter(input_dim=input_dim, num_experts=experts_per_group, top_k=top_k_experts, use_aux_loss=True)
            for _ in range(num_expert_groups)])
        self.expert_groups = nn.ModuleList([
            nn.ModuleList([create_expert(expert_type=expert_types[g], input_dim=input_dim, hidden_dim=hidden_dim,
                                         output_dim=output_dim, expert_id=g * experts_per_group + e, dropout=dropout)
                           for e in range(experts_per_group)])
            for g in range(num_expert_groups)])
        self.output_proj = nn.Linear(output_dim, output_dim)
        self.output_norm = nn.LayerNorm(output_dim)
        self.aux_outputs: Dict[str, Any] = {}

    def _flat_experts(self) -> List[nn.Module]:
        return [e for group in self.expert_groups for e in group]

    def _slab_groups(self) -> List[List[Tuple[str, nn.Parameter]]]:
        groups = []
        if self._homogeneous():
            ex = self._flat_experts()
            for attr in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "layer_norm.weight", "layer_norm.bias"):
                mod, leaf = attr.split(".")
                groups.append([(f"expert_groups.{i // self.experts_per_group}.{i % self.experts_per_group}.{attr}",
                                getattr(getattr(e, mod), leaf)) for i, e in enumerate(ex)])
        groups.append([("output_proj.weight", self.output_proj.weight)])
        return groups

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        B, S, D = x.shape
        G, Epg, Kg, Ke = self.num_expert_groups, self.experts_per_group, self.top_k_groups, self.top_k_experts
        gw, gidx, gaux = self.group_router(x)                         # [B,S,Kg]
        routed = [r(x) for r in self.expert_routers]                  # G x ([B,S,Ke], [B,S,Ke], aux)
        ew = torch.stack([r[0] for r in routed], dim=2)               # [B,S,G,Ke]
        eidx = torch.stack([r[1] for r in routed], dim=2)             # [B,S,G,Ke]
        sel = gidx.clamp(min=0).unsqueeze(-1).expand(B, S, Kg, Ke)    # group of every (slot, expert-slot) pair
        ew_sel = torch.gather(ew, 2, sel)                             # [B,S,Kg,Ke]
        eidx_sel = torch.gather(eidx, 2, sel)
        valid = (gidx >= 0).unsqueeze(-1) & (eidx_sel >= 0)
        flat_idx = torch.where(valid, sel * Epg + eidx_sel, torch.full_like(eidx_sel, -1)).reshape(B, S, Kg * Ke)
        flat_w = (gw.unsqueeze(-1) * ew_sel).reshape(B, S, Kg * Ke)
        # aux: group loss + the expert-router loss of every (slot, group) that received a token (moe_layer.py:489-501)
        total_aux = gaux.get("load_balance_loss", torch.zeros((), device=x.device))
        hits = (gidx.reshape(-1, Kg).unsqueeze(-1) == torch.arange(G, device=x.device)).any(dim=0)     # [Kg, G]
        eaux = torch.stack([r[2].get("load_balance_loss", torch.zeros((), device=x.device)) for r in routed])
        total_aux = total_aux + (hits.to(eaux.dtype) * eaux.unsqueeze(0)).sum()
        self.aux_outputs = {"load_balance_loss": total_aux, "group_probs": gaux.get("router_probs", None)}
        if self._homogeneous():
            mixed = self._forward_grouped(x, flat_w, flat_idx, {}, norm=False)
        else:
            mixed = self._forward_dense(x, flat_w, flat_idx, mask, kwargs, norm=False)
        from ..fusion import blocks
        cdt = resolve_compute_dtype(x)
        slab = self._get_slab(x.device, cdt)
        m2 = ops.to_compute(mixed.reshape(B * S, self.output_dim), cdt)
        y = blocks.linear(m2, self.output_proj, slab)
        y = blocks.add_ln(y, None, self.output_norm)
        return ops.to_compute(y, x.dtype).view(B, S, self.output_dim)
