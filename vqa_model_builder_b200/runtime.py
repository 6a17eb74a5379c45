"""Compute-dtype policy and small shared helpers for the drop-in modules."""
from __future__ import annotations

import contextlib
import copy
import os
from typing import List, Optional, Tuple

import torch
from torch import nn

from .slab import ParamSlab

_COMPUTE = "auto"


def set_compute_dtype(mode: str) -> None:
    """'auto' (bf16 under autocast or for bf16 inputs, else fp32), 'bf16' or 'fp32'."""
    global _COMPUTE
    if mode not in ("auto", "bf16", "fp32"):
        raise ValueError(f"compute dtype must be auto|bf16|fp32, got {mode}")
    _COMPUTE = mode


def get_compute_dtype_mode() -> str:
    return _COMPUTE


def resolve_compute_dtype(x: torch.Tensor) -> torch.dtype:
    """The reference trains under fp16 autocast (training_pipeline.py:457); the B200 path computes in bf16
    with fp32 accumulation, statistics, routing and parameter gradients."""
    if _COMPUTE == "bf16":
        return torch.bfloat16
    if _COMPUTE == "fp32":
        return torch.float32
    if torch.is_autocast_enabled() or x.dtype in (torch.bfloat16, torch.float16):
        return torch.bfloat16
    return torch.float32


# ---------------------------------------------------------------------------------------------------------
# Auxiliary stream: work that is off the critical path of backward (weight / bias gradients) or independent of the
# query stream (key/value projections of the image patches) is enqueued on a second stream so that it fills the SMs
# a small dgrad GEMM leaves idle.  Fork = aux waits for everything enqueued so far on the current stream; join =
# current stream waits for aux.  Every fork is joined before the enclosing autograd Function returns, so consumers
# (autograd's accumulation, optimizers, collectives) never see an unfinished gradient and the pattern is
# CUDA-graph capturable.  B200VQA_AUX_STREAM=0 (or set_aux_stream(False)) serialises everything on one stream.
# ---------------------------------------------------------------------------------------------------------
_AUX = [os.environ.get("B200VQA_AUX_STREAM", "1") != "0"]
_aux_streams = {}
_prefetch_streams = {}


def set_aux_stream(enabled: bool) -> None:
    _AUX[0] = bool(enabled)


def aux_fork(device: torch.device, lane: int = 0) -> Optional["torch.cuda.Stream"]:
    """Order the device's auxiliary stream `lane` after the current stream and return it (None when disabled).
    Lane 0 carries the weight-gradient GEMMs, lane 1 the bias-gradient column sums: a column sum is a handful of
    blocks that fits beside two GEMMs, queued behind the wgrad GEMM it only made the join wait longer."""
    if not _AUX[0]:
        return None
    key = (device.index if device.index is not None else torch.cuda.current_device(), lane)
    s = _aux_streams.get(key)
    if s is None:
        s = torch.cuda.Stream(device=device)
        _aux_streams[key] = s
    s.wait_stream(torch.cuda.current_stream(device))
    return s


def aux_on(s):
    """Context: kernels launched inside go to the auxiliary stream `s` (no-op context when s is None)."""
    return torch.cuda.stream(s) if s is not None else contextlib.nullcontext()


def aux_join(s) -> None:
    if s is not None:
        torch.cuda.current_stream(s.device).wait_stream(s)


class SlabOwner:
    """Mixin: lazily packs the module's parameters into a ParamSlab (rebuilt if parameters were replaced,
    moved by .to(), or the module was deep-copied)."""

    def _slab_groups(self) -> List[List[Tuple[str, nn.Parameter]]]:
        raise NotImplementedError

    def _get_slab(self, device: torch.device, dtype: torch.dtype) -> ParamSlab:
        groups = self._slab_groups()
        flat = [p for g in groups for _, p in g]
        slab: Optional[ParamSlab] = self.__dict__.get("_slab")
        if slab is None or len(slab.params) != len(flat) or any(a is not b for a, b in zip(slab.params, flat)):
            slab = ParamSlab(groups)
            self.__dict__["_slab"] = slab
        slab.refresh(device, dtype)
        return slab

    def invalidate(self) -> None:
        """Mark the bf16 compute copy of the weights stale.  Staleness is normally detected through the parameters'
        version counters; writes through `p.data` (EMA swap-in for evaluation, Lookahead's `p.data.copy_(slow)`,
        solvers/optimizers/vqa_optimizers.py:312) do not bump them, so such code calls this (or `invalidate_all`)."""
        slab = self.__dict__.get("_slab")
        if slab is not None:
            slab.dirty = True

    def train(self, mode: bool = True):
        # train()/eval() switches are where weight swaps (EMA, checkpoint averaging) usually happen: re-cast once
        self.invalidate()
        return super().train(mode)

    def __deepcopy__(self, memo):
        # drop the slab (it is rebuilt on first use) so deepcopy does not duplicate the flat buffers
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_slab", "last_plan"):       # rebuilt lazily / per-forward scratch
                continue
            if k == "aux_outputs":                # holds autograd-tracked tensors of the last forward
                new.__dict__[k] = {}
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new


def prefetch_compute_weights(*modules: nn.Module) -> int:
    """Start refreshing the bf16 compute copies of the given drop-in modules (and their drop-in children) on a
    side stream; returns how many slabs were queued.  Call it at the top of a training step (after
    `optimizer.step()` changed the fp32 masters) for modules that run LATER in the step — e.g. the MOE layer and the
    answer decoder while the fusion runs first — so that their cast (one pass over 6 bytes per parameter) overlaps the
    first module's kernels instead of sitting in front of their own first GEMM.  Each module's forward waits for its own
    copy only.  Every prefetched module must run its forward in the same step (inside a CUDA-graph capture an unused
    prefetch would leave the auxiliary stream un-joined).  With the auxiliary stream disabled this is a no-op: the
    forward casts as usual."""
    n = 0
    for root in modules:
        for m in root.modules():
            if not isinstance(m, SlabOwner):
                continue
            slab: Optional[ParamSlab] = m.__dict__.get("_slab")
            if slab is None or slab.master is None or slab._ready is not None:
                continue                       # never ran (nothing to refresh yet) or already queued
            if not _AUX[0]:
                return n
            dev = slab.master.device
            key = dev.index if dev.index is not None else torch.cuda.current_device()
            s = _prefetch_streams.get(key)      # its own stream: the fork / join pairs of the auxiliary stream inside
            if s is None:                       # the autograd Functions must not end up waiting for a weight cast
                s = torch.cuda.Stream(device=dev)
                _prefetch_streams[key] = s
            s.wait_stream(torch.cuda.current_stream(dev))
            slab.prefetch(dev, s)
            n += 1
    return n


def invalidate_all(model: nn.Module) -> None:
    """invalidate() on every drop-in module of `model` (call after editing weights through `.data`)."""
    for m in model.modules():
        if isinstance(m, SlabOwner):
            m.invalidate()


# ---------------------------------------------------------------------------------------------------------
# Dropout: counter-based masks regenerated inside the kernels (include/b200vqa.h: b200_dropout_t)
# ---------------------------------------------------------------------------------------------------------
_rng_states = {}
_next_site = [1]
_seed_override = [None]


def alloc_sites(n: int) -> int:
    """Reserve `n` consecutive dropout site ids (called once per module at construction)."""
    base = _next_site[0]
    _next_site[0] += n
    return base


def dropout_state(device: torch.device) -> torch.Tensor:
    """Per-device int64 [2] = (seed, step offset)."""
    key = (device.type, device.index)
    st = _rng_states.get(key)
    if st is None:
        seed = _seed_override[0] if _seed_override[0] is not None else torch.initial_seed()
        st = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=device)
        _rng_states[key] = st
    return st


def reseed_dropout(seed: int) -> None:
    """Restart the dropout streams of every device (existing and future) from `seed`, step offset 0."""
    _seed_override[0] = seed
    for st in _rng_states.values():
        st[0] = seed & 0x7FFFFFFFFFFFFFFF
        st[1] = 0


class DropCtx:
    """Dropout context of one module forward: advances the device-side step offset (graph-capturable in-place add)
    and snapshots (seed, offset) so the backward of this forward regenerates the same masks even if other forwards
    ran in between (gradient accumulation)."""

    def __init__(self, training: bool, p: float, device: torch.device, base_site: int):
        self.on = bool(training) and p > 0.0
        self.p = float(p)
        self.base = base_site
        self.snap = None
        if self.on:
            st = dropout_state(device)
            st[1:].add_(1)
            self.snap = st.clone()

    def site(self, k: int):
        return (self.snap, self.p, self.base + k) if self.on else None
