"""One process per GPU over torch.distributed (NCCL on NVLink 5 / NVSwitch; gloo for CPU tests).

Data parallel: the fusion + MOE path is independent per sample, so ranks need no data-path collective in
forward; after backward the parameter gradients are summed across ranks.  Every autograd Function of this
package returns its parameter gradients as views of ONE flat fp32 buffer per call, so the all-reduce works on a
handful of large buckets (one per fused stage) instead of one message per parameter.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> tuple:
    """(rank, world, local_rank) from the torchrun environment; initialises the default process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def grad_buckets(params: Iterable[torch.nn.Parameter]) -> List[torch.Tensor]:
    """Distinct gradient storages to reduce: the flat per-stage buffers when .grad tensors are views of one,
    otherwise the .grad tensors themselves.  Deterministic order (first appearance) on every rank."""
    seen: Dict[int, torch.Tensor] = {}
    order: List[torch.Tensor] = []
    for p in params:
        g = p.grad
        if g is None:
            continue
        base = g._base if g._base is not None else g
        key = base.data_ptr()
        if key not in seen:
            seen[key] = base
            order.append(base)
    return order


def allreduce_gradients(params: Iterable[torch.nn.Parameter], average: bool = True,
                        group: Optional[dist.ProcessGroup] = None) -> int:
    """Sum (or average) gradients across ranks, bucket by bucket.  Returns the number of collectives issued."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    buckets = grad_buckets(params)
    for b in buckets:
        dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
        if average:
            b.div_(world)
    return len(buckets)


def expert_owner(expert: int, num_experts: int, world: int) -> int:
    """Contiguous-block expert placement for expert parallelism: expert e lives on rank e // (E / W)."""
    if num_experts % world != 0:
        raise ValueError(f"expert parallelism needs num_experts ({num_experts}) divisible by world size ({world})")
    return expert // (num_experts // world)
