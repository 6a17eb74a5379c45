"""Parameter slabs: every parameter of a drop-in module lives in ONE contiguous fp32 buffer in HBM.

Why: (1) the grouped expert GEMMs need the experts' weights stacked [E, F, D] while the reference's
state_dict exposes them as per-expert tensors (`experts.{e}.fc1.weight`, moe_layer.py:105-114) — the
per-expert nn.Parameters stay, as views into the slab; (2) the bf16 compute copy of all weights is
refreshed by one cast launch instead of one per nn.Linear (what autocast does in the reference,
training_pipeline.py:457).  Gradients are NOT kept in a persistent slab (gradient accumulation would alias
it): every autograd Function returns views of one flat per-call buffer, which data-parallel training
all-reduces as a bucket (parallel.py).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import torch

from . import _lib

_ALIGN = 64  # elements; keeps every parameter 256-byte aligned (TMA needs 16)

#: re-cast the bf16 compute copy on EVERY forward even when no parameter changed (what autocast does in the
#: reference, one cast per nn.Linear call); benchmarks set this so a step without an optimizer update still
#: pays for the cast.
ALWAYS_REFRESH = False


class ParamSlab:
    def __init__(self, groups: Iterable[List[Tuple[str, torch.nn.Parameter]]]):
        """`groups`: lists of (name, parameter).  Parameters of one group have equal numel and are packed
        back to back (so a group is one stacked [len(group), ...] tensor); group starts are aligned."""
        self.names: List[str] = []
        self.params: List[torch.nn.Parameter] = []
        self.offsets: List[int] = []
        off = 0
        for group in groups:
            group = list(group)
            if not group:
                continue
            n0 = group[0][1].numel()
            for n, p in group:
                if p.numel() != n0:
                    raise ValueError(f"slab group member {n} has numel {p.numel()} != {n0}")
                self.names.append(n)
                self.params.append(p)
                self.offsets.append(off)
                off += n0
            off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = max(off, _ALIGN)
        self._index: Dict[int, int] = {id(p): i for i, p in enumerate(self.params)}
        self.master: torch.Tensor | None = None   # fp32 [total]
        self.shadow: torch.Tensor | None = None   # bf16 [total]
        self._versions: List[int] = []
        #: set by SlabOwner.invalidate(): forces the next refresh to re-cast (writes through `p.data`, e.g. an EMA
        #: weight swap or Lookahead's slow-weight copy, do not bump the version counters checked below)
        self.dirty = False
        #: set by prefetch(): the event the consuming stream waits for instead of casting again
        self._ready = None

    # -- packing --------------------------------------------------------------------------------------
    def _packed(self) -> bool:
        if self.master is None:
            return False
        base = self.master.data_ptr()
        dev = self.master.device
        for p, off in zip(self.params, self.offsets):
            if p.device != dev or p.dtype != torch.float32 or not p.is_contiguous():
                return False
            if p.data_ptr() != base + off * 4:
                return False
        return True

    def pack(self, device: torch.device) -> None:
        """(Re)build the slab on `device` and re-point every Parameter's storage into it."""
        master = torch.zeros(self.total, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = master[off:off + p.numel()].view(p.shape)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
        self.master = master
        self.shadow = None
        self._versions = []

    def ensure(self, device: torch.device) -> None:
        if self.master is None or self.master.device != device or not self._packed():
            self.pack(device)

    # -- views ----------------------------------------------------------------------------------------
    def index_of(self, p: torch.nn.Parameter) -> int:
        return self._index[id(p)]

    def span(self, first: torch.nn.Parameter, count_elems: int, dtype: torch.dtype) -> torch.Tensor:
        """Flat view of `count_elems` elements starting at parameter `first` (for stacked expert weights)."""
        off = self.offsets[self.index_of(first)]
        buf = self.shadow if dtype == torch.bfloat16 else self.master
        return buf[off:off + count_elems]

    def compute_view(self, p: torch.nn.Parameter, dtype: torch.dtype) -> torch.Tensor:
        i = self.index_of(p)
        off = self.offsets[i]
        buf = self.shadow if dtype == torch.bfloat16 else self.master
        return buf[off:off + p.numel()].view(p.shape)

    def contiguous_run(self, plist: List[torch.nn.Parameter]) -> bool:
        """True when plist occupies consecutive back-to-back slots (stackable without a copy)."""
        if not plist:
            return False
        n = plist[0].numel()
        i0 = self.index_of(plist[0])
        for j, p in enumerate(plist):
            if p.numel() != n or self.offsets[self.index_of(p)] != self.offsets[i0] + j * n:
                return False
        return True

    # -- per-step maintenance -------------------------------------------------------------------------
    def refresh(self, device: torch.device, dtype: torch.dtype) -> None:
        """Make the compute copy current: fp32 uses the master in place; bf16 re-casts when stale."""
        self.ensure(device)
        if dtype != torch.bfloat16:
            return
        if self._ready is not None:          # the copy is being made on the auxiliary stream (prefetch): wait, not cast
            torch.cuda.current_stream(device).wait_event(self._ready)
            self._ready = None
            return
        versions = [p._version for p in self.params]
        capturing = torch.cuda.is_current_stream_capturing()
        if self.shadow is None:
            self.shadow = torch.empty(self.total, dtype=torch.bfloat16, device=device)
            self._versions = []
        if capturing or ALWAYS_REFRESH or self.dirty or versions != self._versions:
            _lib.call("b200_cast", self.master, _lib.F32, self.shadow, _lib.BF16, self.total, _lib.stream_ptr())
            self._versions = versions
            self.dirty = False

    def prefetch(self, device: torch.device, stream) -> None:
        """Re-cast the bf16 compute copy on `stream` (ordered after the work already enqueued on the current stream)
        and remember the completion event: the next refresh() on the consuming stream waits for it and skips its own
        cast.  See runtime.prefetch_compute_weights."""
        self.ensure(device)
        if self.shadow is None:
            self.shadow = torch.empty(self.total, dtype=torch.bfloat16, device=device)
        with torch.cuda.stream(stream):
            _lib.call("b200_cast", self.master, _lib.F32, self.shadow, _lib.BF16, self.total, _lib.stream_ptr())
            ev = torch.cuda.Event()
            ev.record(stream)
        self._versions = [p._version for p in self.params]
        self.dirty = False
        self._ready = ev
