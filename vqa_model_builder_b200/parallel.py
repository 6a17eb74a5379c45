"""One process per GPU over torch.distributed (NCCL on NVLink 5 / NVSwitch; gloo for CPU tests).

Data parallel: the fusion + MOE path is independent per sample, so ranks need no data-path collective in
forward; after backward the parameter gradients are summed across ranks.  Every autograd Function of this
package returns its parameter gradients as views of ONE flat fp32 buffer per call, so the all-reduce works on a
handful of large buckets (one per fused stage) instead of one message per parameter.
"""
from __future__ import annotations

import contextlib
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
        import datetime
        # short collective timeout: a mismatched collective must fail fast, not hold the GPUs for ten minutes
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=int(os.environ.get("B200VQA_PG_TIMEOUT", "120"))))
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


def _reduction_units(grads, seen: Optional[set] = None):
    """Tensors to all-reduce for a list of gradients.  The autograd Functions of this package return the parameter
    gradients of one stage as views of ONE flat fp32 buffer (all experts of a bank, weight + bias of a Linear, ...):
    such gradients are replaced by their base buffer, reduced once.  Grouping follows the view structure (`_base`),
    which is identical on every rank; storage addresses are only used to drop duplicates within this rank.  `seen`
    carries the duplicates filter across calls (buckets of one backward pass)."""
    units, local = [], set()
    for g in grads:
        base = g._base if g._base is not None else g
        if base is not g and not (base.dim() == 1 and base.is_contiguous() and base.dtype == g.dtype):
            base = g
        key = (base.data_ptr(), base.numel())
        if key in local or (seen is not None and key in seen):
            continue
        local.add(key)
        if seen is not None:
            seen.add(key)
        units.append(base)
    return units


def allreduce_gradients(params: Iterable[torch.nn.Parameter], average: bool = True,
                        group: Optional[dist.ProcessGroup] = None, seen: Optional[set] = None) -> int:
    """Sum (or average) gradients across ranks: one all-reduce per flat gradient buffer (see `_reduction_units`),
    issued inside a single NCCL group (one fused launch).  Returns the number of tensors reduced."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    grads = _reduction_units([p.grad for p in params if p.grad is not None], seen)
    if not grads:
        return 0
    op = dist.ReduceOp.AVG if (average and grads[0].is_cuda) else dist.ReduceOp.SUM
    done = False
    if grads[0].is_cuda:   # NCCL: fuse into one group launch
        try:
            from torch.distributed.distributed_c10d import _coalescing_manager
            with _coalescing_manager(group=group, device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=op, group=group)
            done = True
        except (ImportError, TypeError):
            done = False
    if not done:
        for g in grads:
            dist.all_reduce(g, op=op, group=group)
    if average and op == dist.ReduceOp.SUM:
        for g in grads:
            g.div_(world)
    return len(grads)


class OverlappedGradReducer:
    """Data-parallel gradient all-reduce overlapped with backward (what DDP's buckets do, for these modules).

    `buckets` lists parameters in the order their gradients become complete during backward (last layer first).
    A post-accumulate hook counts the gradients of a bucket; when the last one has landed the bucket is all-reduced
    on a communication stream ordered after the compute stream, while backward of the earlier layers continues.
    `finish()` (after `backward()`) makes the compute stream wait for the communication stream.  The pattern holds no
    host synchronisation, so a whole step including its collectives can be captured in one CUDA graph."""

    def __init__(self, buckets, average: bool = True, group: Optional[dist.ProcessGroup] = None,
                 transport: str = "nccl"):
        """transport='nccl' (default): coalesced NCCL all-reduce per bucket.
        transport='p2p': the bucket's gradient buffers are staged in a symmetric-memory buffer and reduced by
        `b200_p2p_allreduce_f32` (every rank reduces its 1/W slice with loads from all peers over NVLink and stores
        the result into every peer's buffer); the parameters' .grad are re-pointed at the reduced copies, so nothing
        is copied back.  Verified on 2xB200 (`scripts/dp_check.py`: bit-identical to NCCL) and CUDA-graph capturable,
        but at the benchmark's size it is slower than NCCL (1.66 vs 1.53 ms/step): one staging copy per gradient
        buffer and two cross-rank barriers per bucket cost more than the kernel saves."""
        if transport not in ("nccl", "p2p"):
            raise ValueError("transport must be 'nccl' or 'p2p'")
        self.transport = transport
        self._p2p = {}              # bucket index -> (symmetric buffer, handle, host pointer array, layout)
        self._reduced = {}          # flat buffer key -> (symmetric buffer, offset) for this backward pass
        self._keep = []             # staged gradient buffers of this pass (see _p2p_reduce)
        self.buckets = [[p for p in b if p.requires_grad] for b in buckets]
        self.buckets = [b for b in self.buckets if b]
        self.average = average
        self.group = group
        self.enabled = True
        self.comm = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._left = [len(b) for b in self.buckets]
        self._fired = [set() for _ in self.buckets]     # parameters whose gradient has landed in this pass
        self._seen = set()          # flat buffers already reduced in this backward pass (may span two buckets)
        self._handles = []
        for bi, bucket in enumerate(self.buckets):
            for p in bucket:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook(bi)))

    def _hook(self, bi: int):
        def fire(param):
            if not self.enabled or id(param) in self._fired[bi]:
                return                      # a second accumulation into the same parameter must not count twice
            self._fired[bi].add(id(param))
            self._left[bi] -= 1
            if self._left[bi] == 0:
                self._launch(bi)
        return fire

    def _launch(self, bi: int) -> None:
        if self.comm is None:                      # CPU / gloo: no streams
            if self.transport == "p2p" and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                self._p2p_reduce(bi)               # staged path with a plain all-reduce (host-logic tests)
            else:
                allreduce_gradients(self.buckets[bi], self.average, self.group, self._seen)
            return
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            if self.transport == "p2p" and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                self._p2p_reduce(bi)
            else:
                allreduce_gradients(self.buckets[bi], self.average, self.group, self._seen)

    def _p2p_reduce(self, bi: int) -> None:
        import ctypes
        from . import _lib
        params = [p for p in self.buckets[bi] if p.grad is not None]
        units = _reduction_units([p.grad for p in params], self._seen)
        for u in units:
            if u.dtype != torch.float32:
                raise RuntimeError("p2p gradient all-reduce expects fp32 gradients")
        if not units:
            self._repoint(params)
            return
        group = self.group if self.group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        st = self._p2p.get(bi)
        sizes = [(u.numel() + 3) // 4 * 4 for u in units]
        on_gpu = units[0].is_cuda
        if st is None or st["sizes"] != sizes:        # first use (eager warm-up): collective allocation
            total = sum(sizes)
            if on_gpu:
                import torch.distributed._symmetric_memory as symm_mem
                if hasattr(symm_mem, "enable_symm_mem_for_group"):
                    try:
                        symm_mem.enable_symm_mem_for_group(group.group_name)
                    except Exception:
                        pass
                buf = symm_mem.empty(total, dtype=torch.float32, device=units[0].device)
                hdl = symm_mem.rendezvous(buf, group)
                # the tensor may sit at an offset inside its symmetric allocation: apply this rank's offset to all
                delta = buf.data_ptr() - int(hdl.buffer_ptrs[rank])
                host_ptrs = (ctypes.c_ulonglong * world)(*[int(a) + delta for a in hdl.buffer_ptrs])
            else:                                     # CPU stand-in: same staging / re-pointing, library all-reduce
                buf, hdl, host_ptrs = torch.empty(total, dtype=torch.float32), None, None
            buf.zero_()
            st = {"buf": buf, "hdl": hdl, "ptrs": host_ptrs, "sizes": sizes, "total": total}
            self._p2p[bi] = st
            if hdl is not None:
                hdl.barrier(channel=0)
        buf, hdl = st["buf"], st["hdl"]
        offs, off = [], 0
        for u, sz in zip(units, sizes):               # stage this rank's contribution
            buf[off:off + u.numel()].copy_(u.reshape(-1))
            offs.append(off)
            off += sz
        # The staged buffers are kept alive until finish(): re-pointing .grad below would free them, and (1) their
        # memory must outlive the copy enqueued on the communication stream, (2) a later gradient allocated at the
        # same address with the same size would collide with the (address, size) keys of this pass — it would be
        # taken for "already reduced" and re-pointed at another parameter's data.
        self._keep.extend(units)
        if hdl is not None:
            hdl.barrier(channel=0)                    # every rank's contribution is in place
            _lib.call("b200_p2p_allreduce_f32", st["ptrs"], rank, world, 0, st["total"],
                      (1.0 / world) if self.average else 1.0, _lib.stream_ptr())
            hdl.barrier(channel=0)                    # every peer's stores into this rank's buffer have landed
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
            if self.average:
                buf.div_(world)
        for u, o in zip(units, offs):
            self._reduced[(u.data_ptr(), u.numel())] = (buf, o)
        self._repoint(params)

    def _repoint(self, params) -> None:
        """Point the gradients at the reduced copies in symmetric memory (same shapes, offsets inside their flat
        buffer preserved); a flat buffer may have been reduced with an earlier bucket."""
        for p in params:
            g = p.grad
            base = g._base if g._base is not None else g
            if base is not g and not (base.dim() == 1 and base.is_contiguous() and base.dtype == g.dtype):
                base = g
            hit = self._reduced.get((base.data_ptr(), base.numel()))
            if hit is None:
                continue
            buf, off = hit
            if not g.is_contiguous():                  # odd view: copy the reduced values back instead
                base.reshape(-1).copy_(buf[off:off + base.numel()])
                continue
            start = off + (g.data_ptr() - base.data_ptr()) // g.element_size()
            p.grad = buf[start:start + g.numel()].view(g.shape)

    def finish(self) -> None:
        """Call once after backward(): reduces buckets whose hooks did not all fire (unused parameters), joins the
        communication stream and re-arms the counters."""
        if self.enabled:
            for bi, left in enumerate(self._left):
                if left > 0:
                    self._launch(bi)
            if self.comm is not None:
                torch.cuda.current_stream().wait_stream(self.comm)
        self._left = [len(b) for b in self.buckets]
        self._fired = [set() for _ in self.buckets]
        self._seen = set()
        self._reduced = {}
        self._keep = []

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


def expert_owner(expert: int, num_experts: int, world: int) -> int:
    """Contiguous-block expert placement for expert parallelism: expert e lives on rank e // (E / W)."""
    if num_experts % world != 0:
        raise ValueError(f"expert parallelism needs num_experts ({num_experts}) divisible by world size ({world})")
    return expert // (num_experts // world)


# ---------------------------------------------------------------------------------------------------------
# Expert parallelism: experts sharded across ranks, tokens exchanged with all-to-all over NVLink (NCCL)
# ---------------------------------------------------------------------------------------------------------
class AllToAllRows(torch.autograd.Function):
    """Variable-split all-to-all of row blocks; backward is the all-to-all with the splits swapped."""

    @staticmethod
    def forward(ctx, x, in_splits, out_splits, group):
        x = x.contiguous()
        out = torch.empty((sum(out_splits),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_to_all_single(out, x, list(out_splits), list(in_splits), group=group)
        ctx.splits = (list(in_splits), list(out_splits), group, x.shape[0])
        return out

    @staticmethod
    def backward(ctx, g):
        in_splits, out_splits, group, n_in = ctx.splits
        g = g.contiguous()
        gx = torch.empty((n_in,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        dist.all_to_all_single(gx, g, in_splits, out_splits, group=group)
        return gx, None, None, None


def exchange_counts(counts: torch.Tensor, world: int, group=None):
    """counts [E] (pairs routed to every GLOBAL expert on this rank) -> (send [W, E_local], recv [W, E_local]) as
    Python lists: recv[s][e] = rows rank s sends to my local expert e.  One small all-to-all + one host read."""
    E = counts.numel()
    send = counts.to(torch.int64).view(world, E // world).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    both = torch.stack([send, recv]).tolist()
    return both[0], both[1]


class ExpertParallelMOELayer(torch.nn.Module):
    """MOELayer with its FeedForwardExperts sharded over the ranks of `group` (expert e on rank e // (E/W)).

    forward: router (replicated, global load-balance statistics) -> routing plan -> rows in canonical
    (expert, token) order, which is also destination-rank order -> all-to-all -> grouped expert FFN on the owner ->
    all-to-all back -> weighted combine + output_norm on the token's home rank.  Gradients of the sharded experts are
    complete locally (every token routed to an expert reached its owner); replicated parameters (router gate,
    output_norm) are all-reduced like any data-parallel parameter."""

    def __init__(self, full_layer, group=None):
        super().__init__()
        from .moe.layers import MOELayer
        assert isinstance(full_layer, MOELayer) and full_layer._homogeneous(), "needs homogeneous FeedForwardExperts"
        if getattr(full_layer, "capacity_factor", None) is not None:
            raise ValueError("expert parallelism does not implement SparseMOELayer's capacity_factor (the per-expert "
                             "top-C selection needs the weights of every rank's tokens); use MOELayer")
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        E = full_layer.num_experts
        if E % self.world != 0:
            raise ValueError(f"num_experts={E} not divisible by world size {self.world}")
        self.num_experts = E
        self.experts_per_rank = E // self.world
        lo = self.rank * self.experts_per_rank
        self.local = full_layer
        # keep only the local shard of the expert bank
        full_layer.experts = torch.nn.ModuleList(list(full_layer.experts)[lo:lo + self.experts_per_rank])
        full_layer.router.stats_group = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
        self.input_dim, self.hidden_dim, self.output_dim = full_layer.input_dim, full_layer.hidden_dim, full_layer.output_dim
        self.top_k = full_layer.top_k
        self.aux_outputs = {}

    def replicated_parameters(self):
        return list(self.local.router.parameters()) + list(self.local.output_norm.parameters())

    def expert_parameters(self):
        return list(self.local.experts.parameters())

    def get_aux_loss(self):
        return self.local.get_aux_loss()

    def forward(self, x: torch.Tensor, mask=None, **kwargs) -> torch.Tensor:
        from . import ops
        from .runtime import resolve_compute_dtype
        L = self.local
        B, S, D = x.shape
        N, K, E, W, El = B * S, self.top_k, self.num_experts, self.world, self.experts_per_rank
        weights, indices, aux = L.router(x)
        L.aux_outputs = self.aux_outputs = aux
        cdt = resolve_compute_dtype(x)
        x2 = ops.to_compute(x.reshape(N, D), cdt)
        stash = aux.get("_b200_idx32")
        idx32 = stash[1] if stash is not None and stash[0] is indices else indices.reshape(N, K).to(torch.int32)
        plan = ops.RoutingPlan(idx32, E)
        NK = plan.NK
        # compact canonical layout (no padding on the wire)
        ar = torch.arange(NK, device=x.device, dtype=torch.int32)
        valid = plan.cmp_pos >= 0
        row_src_c = torch.full((NK,), -1, dtype=torch.int32, device=x.device)
        row_src_c[plan.cmp_pos[valid].long()] = ar[valid]
        xc = ops.PermuteFn.apply(x2, row_src_c, plan.cmp_off, plan.cmp_pos, E, K, NK)
        send, recv = exchange_counts(plan.counts, W, self.group)
        in_splits = [sum(r) for r in send]
        out_splits = [sum(r) for r in recv]
        n_send, n_recv = sum(in_splits), sum(out_splits)
        got = AllToAllRows.apply(xc[:n_send], in_splits, out_splits, self.group)
        if n_recv > 0:
            flat_counts = torch.tensor([c for r in recv for c in r], device=x.device)
            ids = torch.arange(El, device=x.device, dtype=torch.int32).repeat(W)
            idx_recv = torch.repeat_interleave(ids, flat_counts, output_size=n_recv).view(n_recv, 1)
            plan2 = ops.RoutingPlan(idx_recv, El)
            stacks = L._expert_stacks(x.device, cdt)
            ex0 = L.experts[0]
            xp2 = ops.PermuteFn.apply(got, plan2.row_src, plan2.pad_off, plan2.dest_row, El, 1, plan2.Rmax)
            from .runtime import DropCtx, alloc_sites
            if "_sites" not in self.__dict__:
                self.__dict__["_sites"] = alloc_sites(2)
            dc = DropCtx(self.training, float(ex0.dropout_rate), x.device, self.__dict__["_sites"])
            z2 = ops.ExpertFFNFn.apply(xp2, plan2.tile_group, plan2.pad_off, stacks, ex0.act_code,
                                       D == ex0.output_dim, ex0.layer_norm.eps,
                                       (dc.site(0), dc.site(1)) if dc.on else None, *L._expert_params())
            zc = ops.GatherRowsFn.apply(z2, plan2.dest_row, plan2.row_src, plan2.pad_off, El)
        else:
            # no row reached this rank's experts: keep BOTH all-to-alls in the autograd graph (a detached tensor here
            # would skip their backward collectives on this rank while the peers block in theirs)
            zc = got.sum(dim=1, keepdim=True).expand(0, self.output_dim).contiguous()
        back = AllToAllRows.apply(zc, out_splits, in_splits, self.group)
        w2d = weights.reshape(N, K).to(torch.float32)
        out = ops.CombineFn.apply(back, w2d, plan.cmp_pos, row_src_c[:max(n_send, 1)], L.output_norm.weight,
                                  L.output_norm.bias, L.output_norm.eps)
        self.last_plan = plan
        return ops.to_compute(out, x.dtype).view(B, S, self.output_dim)


def finish_gradients(replicated, expert_sharded, group=None) -> None:
    """After backward under DP(+EP): average replicated-parameter gradients across ranks; scale the sharded
    experts' gradients by 1/W so both follow the same mean-over-ranks loss convention."""
    if not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    allreduce_gradients(replicated, average=True, group=group)
    for p in expert_sharded:
        if p.grad is not None:
            p.grad.div_(world)


# ---------------------------------------------------------------------------------------------------------
# Expert parallelism with dispatch/return FUSED with the collective: kernels store rows straight into the owner GPU's
# grouped-GEMM input over NVLink (symmetric memory) — no all_to_all call, no send/receive staging, no second permute,
# no host-side split sizes, no host sync: the layer is asynchronous on the stream and CUDA-graph capturable.
# ---------------------------------------------------------------------------------------------------------
def _symm_alloc(nbytes: int, group, device):
    """(uint8 tensor in symmetric memory, handle, list of the W ranks' addresses of this tensor)."""
    import torch.distributed._symmetric_memory as symm_mem
    group = group if group is not None else dist.group.WORLD
    if hasattr(symm_mem, "enable_symm_mem_for_group"):
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
    buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
    hdl = symm_mem.rendezvous(buf, group)
    rank = dist.get_rank(group)
    # the tensor may sit at an offset inside its symmetric allocation: the offset is the same on every rank
    delta = buf.data_ptr() - int(hdl.buffer_ptrs[rank])
    return buf, hdl, [int(a) + delta for a in hdl.buffer_ptrs], delta


class _P2PState:
    """Symmetric buffers of one EP layer + the device arrays of peer pointers the kernels index."""

    def __init__(self, group, world: int, rank: int, E: int, D: int, nk_cap: int, dtype: torch.dtype, device):
        from . import _lib
        self.world, self.rank, self.E, self.D, self.nk_cap, self.dtype = world, rank, E, D, nk_cap, dtype
        self.El = E // world
        # worst case: every pair of every rank lands on this rank (exact routing, nothing is ever dropped)
        self.Rcap = _lib.query("b200_moe_max_rows", world * nk_cap, self.El)
        es = torch.empty((), dtype=dtype).element_size()
        al = lambda n: (n + 255) // 256 * 256
        sizes = {"tab": al(world * E * 4), "recv_x": al(self.Rcap * D * es), "recv_g": al(self.Rcap * D * es),
                 "ret_z": al(nk_cap * D * es), "ret_dx": al(nk_cap * D * es)}
        self.offs, total = {}, 0
        for k, v in sizes.items():
            self.offs[k] = total
            total += v
        self.buf, self.hdl, bases, _ = _symm_alloc(total, group, device)
        self.ptrs = {k: torch.tensor([b + off for b in bases], dtype=torch.int64, device=device)
                     for k, off in self.offs.items()}
        self.buf.zero_()
        self.live = 0           # forwards whose backward has not run yet (their saved tensors alias the buffers)
        self.barrier()

    def local(self, key: str):
        off = self.offs[key]
        if key == "tab":
            return self.buf[off:off + self.world * self.E * 4].view(torch.int32)
        rows = self.Rcap if key.startswith("recv") else self.nk_cap
        es = torch.empty((), dtype=self.dtype).element_size()
        return self.buf[off:off + rows * self.D * es].view(self.dtype).view(rows, self.D)

    def barrier(self):
        self.hdl.barrier(channel=0)


class _EPMeta:
    __slots__ = ("plan", "send_base", "pad_off2", "tile_group2", "row_home", "K", "El", "N", "zero_copy")


class _P2PDispatchFn(torch.autograd.Function):
    """token rows -> the owners' grouped-GEMM inputs (peer stores into the padded layout); backward brings the input
    gradients of those rows home and sums the K copies per token."""

    @staticmethod
    def forward(ctx, x2, st, meta):
        from . import _lib
        plan = meta.plan
        x2 = x2.contiguous()
        _lib.call("b200_ep_dispatch", x2, plan.cmp_src, plan.cmp_off, meta.send_base, meta.pad_off2, st.ptrs["recv_x"],
                  st.rank, meta.K, plan.NK, plan.E, meta.El, st.D, st.Rcap, _lib.dtype_code(x2.dtype), _lib.stream_ptr())
        st.barrier()
        ctx.st, ctx.meta = st, meta
        out = st.local("recv_x")
        return out if meta.zero_copy else out.clone()

    @staticmethod
    def backward(ctx, d_recv):
        from . import _lib
        st, meta = ctx.st, ctx.meta
        d_recv = d_recv.contiguous()
        _lib.call("b200_ep_return", d_recv, meta.row_home, meta.pad_off2, st.ptrs["ret_dx"], meta.El, st.D, st.Rcap,
                  st.nk_cap, meta.plan.NK, _lib.dtype_code(d_recv.dtype), _lib.stream_ptr())
        st.barrier()
        dx = torch.empty((meta.N, st.D), dtype=d_recv.dtype, device=d_recv.device)
        _lib.call("b200_moe_unpermute", st.local("ret_dx"), meta.plan.cmp_pos, None, meta.N, meta.K, st.D,
                  _lib.dtype_code(d_recv.dtype), dx, _lib.stream_ptr())
        st.live = max(0, st.live - 1)
        return dx, None, None


class _P2PReturnFn(torch.autograd.Function):
    """expert outputs (padded layout on the owner) -> the tokens' home ranks at their compact positions; backward
    sends the combine gradients to the owners with the dispatch kernel."""

    @staticmethod
    def forward(ctx, z2, st, meta):
        from . import _lib
        z2 = z2.contiguous()
        _lib.call("b200_ep_return", z2, meta.row_home, meta.pad_off2, st.ptrs["ret_z"], meta.El, st.D, st.Rcap,
                  st.nk_cap, meta.plan.NK, _lib.dtype_code(z2.dtype), _lib.stream_ptr())
        st.barrier()
        ctx.st, ctx.meta = st, meta
        out = st.local("ret_z")[:meta.plan.NK]
        return out if meta.zero_copy else out.clone()

    @staticmethod
    def backward(ctx, d_back):
        from . import _lib
        st, meta = ctx.st, ctx.meta
        plan = meta.plan
        d_back = d_back.contiguous()
        _lib.call("b200_ep_dispatch", d_back, None, plan.cmp_off, meta.send_base, meta.pad_off2, st.ptrs["recv_g"],
                  st.rank, meta.K, plan.NK, plan.E, meta.El, st.D, st.Rcap, _lib.dtype_code(d_back.dtype),
                  _lib.stream_ptr())
        st.barrier()
        return st.local("recv_g"), None, None


class P2PExpertParallelMOELayer(ExpertParallelMOELayer):
    """ExpertParallelMOELayer whose dispatch / return are single kernels over NVLink peer memory.

    forward : router -> plan -> push counts | barrier | layout, dispatch (rows land in the owner's 128-row-padded
              grouped-GEMM input) | barrier | grouped expert FFN | return | barrier | combine + output_norm
    backward: the same two kernels in the opposite direction (2 barriers).
    `max_tokens` (tokens per rank and forward, identical on all ranks) sizes the symmetric buffers; without it they are
    sized from the first forward, agreed across ranks by one all-reduce.  With `zero_copy` (default) the tensors saved
    for backward alias the symmetric buffers: a second forward before the backward of the first raises."""

    def __init__(self, full_layer, group=None, max_tokens: Optional[int] = None, zero_copy: bool = True):
        super().__init__(full_layer, group)
        self.max_tokens = max_tokens
        self.zero_copy = zero_copy

    def _state(self, N: int, K: int, D: int, dtype, device) -> _P2PState:
        st = self.__dict__.get("_p2p")
        if st is not None and st.dtype == dtype and st.D == D:
            if N * K > st.nk_cap:
                raise RuntimeError(f"P2PExpertParallelMOELayer: {N} tokens exceed the buffers sized for "
                                   f"{st.nk_cap // K}; construct the layer with max_tokens=")
            return st
        n_cap = self.max_tokens if self.max_tokens is not None else N
        if self.max_tokens is None and self.world > 1:     # every rank must size the same: agree on the maximum once
            t = torch.tensor([n_cap], dtype=torch.int64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            n_cap = int(t.item())
        st = _P2PState(self.group, self.world, self.rank, self.num_experts, D, n_cap * K, dtype, device)
        self.__dict__["_p2p"] = st
        return st

    def forward(self, x: torch.Tensor, mask=None, **kwargs) -> torch.Tensor:
        from . import _lib, ops
        from .runtime import DropCtx, alloc_sites, resolve_compute_dtype
        L = self.local
        B, S, D = x.shape
        N, K, E, W, El = B * S, self.top_k, self.num_experts, self.world, self.experts_per_rank
        weights, indices, aux = L.router(x)
        L.aux_outputs = self.aux_outputs = aux
        cdt = resolve_compute_dtype(x)
        x2 = ops.to_compute(x.reshape(N, D), cdt)
        stash = aux.get("_b200_idx32")
        idx32 = stash[1] if stash is not None and stash[0] is indices else indices.reshape(N, K).to(torch.int32)
        plan = ops.RoutingPlan(idx32, E, with_cmp_src=True)
        dev = x.device
        st = self._state(N, K, D, cdt, dev)
        if self.zero_copy and torch.is_grad_enabled() and x2.requires_grad:
            if st.live > 0:
                raise RuntimeError("P2PExpertParallelMOELayer(zero_copy=True): forward called again before the backward "
                                   "of the previous forward; use zero_copy=False for such schedules")
            st.live += 1
        sp = _lib.stream_ptr()
        # phase 1: publish my per-expert counts in every rank's table, then derive every layout on the device
        _lib.call("b200_ep_push_counts", plan.counts, st.ptrs["tab"], self.rank, W, E, sp)
        st.barrier()
        i32 = dict(dtype=torch.int32, device=dev)
        meta = _EPMeta()
        meta.plan, meta.K, meta.El, meta.N, meta.zero_copy = plan, K, El, N, self.zero_copy
        meta.send_base = torch.empty(E, **i32)
        meta.pad_off2 = torch.empty(2 * El + 1, **i32)
        meta.tile_group2 = torch.empty(st.Rcap // 128, **i32)
        meta.row_home = torch.empty(st.Rcap, **i32)
        _lib.call("b200_ep_layout", st.local("tab"), self.rank, W, E, st.Rcap, st.nk_cap, meta.send_base, meta.pad_off2,
                  meta.tile_group2, meta.row_home, sp)
        # phase 2: permute fused with the exchange: rows land in the owners' grouped-GEMM inputs
        xp2 = _P2PDispatchFn.apply(x2, st, meta)
        stacks = L._expert_stacks(dev, cdt)
        ex0 = L.experts[0]
        if "_sites" not in self.__dict__:
            self.__dict__["_sites"] = alloc_sites(2)
        dc = DropCtx(self.training, float(ex0.dropout_rate), dev, self.__dict__["_sites"])
        with ops.local_grads():          # the shard's gradients are complete locally: keep them out of the DP arena
            z2 = ops.ExpertFFNFn.apply(xp2, meta.tile_group2, meta.pad_off2, stacks, ex0.act_code,
                                       D == ex0.output_dim, ex0.layer_norm.eps,
                                       (dc.site(0), dc.site(1)) if dc.on else None, *L._expert_params())
        # phase 3: expert outputs go straight back to the tokens' home ranks (compact canonical positions)
        back = _P2PReturnFn.apply(z2, st, meta)
        w2d = weights.reshape(N, K).to(torch.float32)
        out = ops.CombineFn.apply(back, w2d, plan.cmp_pos, plan.cmp_src, L.output_norm.weight, L.output_norm.bias,
                                  L.output_norm.eps)
        self.last_plan = plan
        self.last_meta = meta
        return ops.to_compute(out, x.dtype).view(B, S, self.output_dim)


# ---------------------------------------------------------------------------------------------------------
# Data-parallel gradients in symmetric memory: the wgrad kernels write into the buffer the all-reduce works on
# ---------------------------------------------------------------------------------------------------------
class GradArena:
    """Bump allocator over one fp32 buffer in symmetric memory.  While installed (`ops.set_grad_arena`), every backward
    of this package takes its flat parameter-gradient buffer from here, so the gradients of a step lie back to back in
    backward order at offsets that are identical on every rank and in every step (`reset()` at the start of a step).
    `ArenaGradReducer` then all-reduces ranges of the arena in place — through the NVSwitch (multimem.ld_reduce /
    multimem.st on the multicast mapping) when the allocation has one, else with peer loads — without any staging
    copy and without touching `.grad`."""

    ALIGN = 32      # floats: ranges stay 128-byte aligned

    def __init__(self, numel: int, group=None, device=None):
        self.numel = (int(numel) + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if device is not None and torch.device(device).type == "cuda":
            raw, self.hdl, bases, delta = _symm_alloc(self.numel * 4, self.group, device)
            self.buf = raw.view(torch.float32)
            self.peer_ptrs = bases
            mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
            self.multicast_ptr = mc + delta if mc else 0
        else:       # CPU stand-in for the gloo host-logic tests: same bump allocation, library all-reduce on ranges
            self.buf, self.hdl, self.peer_ptrs, self.multicast_ptr = torch.empty(self.numel), None, [], 0
        self.device = self.buf.device
        self.off = 0
        self.overflow = 0
        self.buf.zero_()
        if self.hdl is not None:
            self.hdl.barrier(channel=0)

    def reset(self) -> None:
        self.off = 0

    def take(self, n: int, device) -> Optional[torch.Tensor]:
        n_al = (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if torch.device(device) != self.device or self.off + n_al > self.numel:
            self.overflow += n_al
            return None
        t = self.buf[self.off:self.off + n]
        self.off += n_al
        return t

    def offset_of(self, t: torch.Tensor) -> int:
        """element offset of a tensor inside the arena, or -1"""
        d = t.data_ptr() - self.buf.data_ptr()
        return d // 4 if 0 <= d < self.numel * 4 else -1


class ArenaMeter:
    """Stand-in arena that only measures: install it for one eager step to learn how many floats a step needs."""

    def __init__(self):
        self.total = 0
        self.device = None

    def reset(self):
        self.total = 0

    def take(self, n: int, device):
        self.total += (n + GradArena.ALIGN - 1) // GradArena.ALIGN * GradArena.ALIGN
        return None


class ArenaGradReducer:
    """Overlapped data-parallel all-reduce over a GradArena.  `buckets` lists parameters in backward order; when the
    gradients of a bucket have landed, the arena range [cursor, end of the bucket's last gradient) is all-reduced in
    place on the communication stream (the sweep covers every allocated float exactly once, whatever the bucket
    boundaries), while backward continues.  Gradients that are not in the arena (produced by torch, or taken after an
    overflow) fall back to NCCL.  No host synchronisation: the step stays CUDA-graph capturable."""

    def __init__(self, arena: GradArena, buckets, average: bool = True, max_blocks: int = 0):
        self.arena = arena
        self.buckets = [[p for p in b if p.requires_grad] for b in buckets]
        self.buckets = [b for b in self.buckets if b]
        self.average = average
        self.max_blocks = max_blocks
        self.enabled = True
        self.comm = torch.cuda.Stream() if arena.hdl is not None else None
        self.cursor = 0
        self.ranges = []            # (lo, hi) reduced in the current pass (diagnostics / tests)
        self._left = [len(b) for b in self.buckets]
        self._fired = [set() for _ in self.buckets]
        self._handles = []
        for bi, bucket in enumerate(self.buckets):
            for p in bucket:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook(bi)))
        import ctypes
        self._host_ptrs = (ctypes.c_ulonglong * arena.world)(*arena.peer_ptrs) if arena.peer_ptrs else None

    def _hook(self, bi: int):
        def fire(param):
            if not self.enabled or id(param) in self._fired[bi]:
                return
            self._fired[bi].add(id(param))
            self._left[bi] -= 1
            if self._left[bi] == 0:
                self._launch(self.buckets[bi])
        return fire

    def _reduce_range(self, lo: int, hi: int) -> None:
        from . import _lib
        a = self.arena
        if hi <= lo:
            return
        self.ranges.append((lo, hi))
        scale = (1.0 / a.world) if self.average else 1.0
        if a.hdl is None:                       # CPU stand-in
            dist.all_reduce(a.buf[lo:hi], op=dist.ReduceOp.SUM, group=a.group)
            if self.average:
                a.buf[lo:hi].div_(a.world)
            return
        a.hdl.barrier(channel=1)                # every rank's gradients of this range are written
        if a.multicast_ptr:
            _lib.call("b200_nvls_allreduce_f32", a.multicast_ptr, a.rank, a.world, lo, hi - lo, scale,
                      self.max_blocks, _lib.stream_ptr())
        else:
            _lib.call("b200_p2p_allreduce_f32", self._host_ptrs, a.rank, a.world, lo, hi - lo, scale,
                      _lib.stream_ptr())
        a.hdl.barrier(channel=1)                # every rank's stores into this rank's copy have landed

    def _launch(self, params) -> None:
        a = self.arena
        hi, outside = self.cursor, []
        for p in params:
            g = p.grad
            if g is None:
                continue
            off = a.offset_of(g)
            if off < 0:
                outside.append(p)
                continue
            end = (off + g.numel() + a.ALIGN - 1) // a.ALIGN * a.ALIGN
            hi = max(hi, end)
        if self.comm is not None:
            self.comm.wait_stream(torch.cuda.current_stream())
        with (torch.cuda.stream(self.comm) if self.comm is not None else contextlib.nullcontext()):
            self._reduce_range(self.cursor, hi)
            if outside:
                allreduce_gradients(outside, self.average, a.group)
        self.cursor = max(self.cursor, hi)

    def finish(self) -> None:
        """Call once after backward(): reduces what the hooks did not cover, joins the communication stream."""
        if self.enabled:
            pending = [p for bi, b in enumerate(self.buckets) if self._left[bi] > 0 for p in b
                       if id(p) not in self._fired[bi]]
            if self.comm is not None:
                self.comm.wait_stream(torch.cuda.current_stream())
            with (torch.cuda.stream(self.comm) if self.comm is not None else contextlib.nullcontext()):
                self._reduce_range(self.cursor, self.arena.off)
                outside = [p for p in pending if p.grad is not None and self.arena.offset_of(p.grad) < 0]
                if outside:
                    allreduce_gradients(outside, self.average, self.arena.group)
            if self.comm is not None:
                torch.cuda.current_stream().wait_stream(self.comm)
        self.cursor = 0
        self.ranges = []
        self._left = [len(b) for b in self.buckets]
        self._fired = [set() for _ in self.buckets]

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
