"""Answer prediction head (the classification tail right behind the MOE layer).
Reference: src/modeling/meta_arch/vqa_model.py:436-477 (AnswerHead), vqa_config.py:153-168 (AnswerHeadConfig)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import torch
from torch import nn

from . import ops
from ._lib import ACT_NONE, ACT_RELU
from .fusion import blocks
from .runtime import DropCtx, SlabOwner, alloc_sites, resolve_compute_dtype


@dataclass
class AnswerHeadConfig:
    """vqa_config.py:153-168."""
    num_answers: int = 3000
    hidden_dims: List[int] = field(default_factory=lambda: [512, 256])
    dropout: float = 0.3
    use_sigmoid: bool = False
    classifier_type: str = "mlp"


class AnswerHead(SlabOwner, nn.Module):
    """[Linear -> ReLU -> Dropout] x len(hidden_dims) -> Linear(num_answers); `classifier` keeps the reference's
    nn.Sequential layout, so state_dict keys (classifier.0.weight, classifier.3.weight, ...) are unchanged."""

    def __init__(self, config, input_dim: int):
        nn.Module.__init__(self)
        self.config = config
        layers, prev = [], input_dim
        for hidden in config.hidden_dims:
            layers.extend([nn.Linear(prev, hidden), nn.ReLU(), nn.Dropout(config.dropout)])
            prev = hidden
        layers.append(nn.Linear(prev, config.num_answers))
        self.classifier = nn.Sequential(*layers)
        self._sites = alloc_sites(max(1, len(config.hidden_dims)))

    def _slab_groups(self):
        return blocks.param_groups(self)

    def _linears(self) -> List[nn.Linear]:
        return [m for m in self.classifier if isinstance(m, nn.Linear)]

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        lead = features.shape[:-1]
        x = features.reshape(-1, features.shape[-1])
        cdt = resolve_compute_dtype(x)
        slab = self._get_slab(x.device, cdt)
        lins = self._linears()
        dc = DropCtx(self.training, float(self.config.dropout), x.device, self._sites)
        acts = [ACT_RELU] * (len(lins) - 1) + [ACT_NONE]
        drops = [dc.site(i) for i in range(len(lins) - 1)] + [None]
        params = []
        for lin in lins:
            params += [lin.weight, lin.bias, slab.compute_view(lin.weight, cdt)]
        logits = ops.MLPFn.apply(ops.to_compute(x.contiguous(), cdt), acts, drops, *params)
        return ops.to_compute(logits, features.dtype).reshape(*lead, logits.shape[-1])
