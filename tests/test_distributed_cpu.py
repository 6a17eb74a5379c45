"""CPU, world_size 2 over gloo: host-side logic of the multi-GPU path — gradient bucketing + all-reduce,
the autograd all-to-all used by expert parallelism, count exchange and expert placement."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vqa_model_builder_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    try:
        parallel.init_distributed("gloo")
        fn(rank, world)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def run2(fn):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fn, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, status in res:
        assert status == "ok", f"rank {rank}: {status}"


def _dp_buckets(rank, world):
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.zeros(4, 3))
    b = torch.nn.Parameter(torch.zeros(5))
    c = torch.nn.Parameter(torch.zeros(2))
    flat = torch.arange(17, dtype=torch.float32) * (rank + 1)        # one stage returned views of one flat buffer
    a.grad, b.grad = flat[:12].view(4, 3), flat[12:]
    c.grad = torch.full((2,), float(rank + 1))
    buckets = parallel.grad_buckets([a, b, c])
    assert len(buckets) == 2 and buckets[0].data_ptr() == flat.data_ptr()
    n = parallel.allreduce_gradients([a, b, c], average=True)
    assert n == 2          # a.grad and b.grad are views of one flat buffer: reduced once, as a whole
    assert torch.allclose(a.grad, (torch.arange(12, dtype=torch.float32) * 1.5).view(4, 3))
    assert torch.allclose(b.grad, torch.arange(12, 17, dtype=torch.float32) * 1.5)
    assert torch.allclose(c.grad, torch.full((2,), 1.5))


def _a2a_autograd(rank, world):
    # rank r sends (r+1) rows to rank 0 and 2 rows to rank 1
    in_splits = [rank + 1, 2]
    counts = torch.tensor([rank + 1, 2])
    send, recv = parallel.exchange_counts(counts, world)
    assert send == [[rank + 1], [2]]
    assert recv == ([[1], [2]] if rank == 0 else [[2], [2]])
    out_splits = [r[0] for r in recv]
    x = (torch.arange(sum(in_splits) * 3, dtype=torch.float32).view(-1, 3) + 100 * rank).requires_grad_()
    y = parallel.AllToAllRows.apply(x, in_splits, out_splits, None)
    assert y.shape[0] == sum(out_splits)
    if rank == 0:   # got my own first row and rank 1's first two rows
        assert torch.equal(y[0], x[0].detach()) and float(y[1, 0]) == 100.0
    (y * (rank + 1)).sum().backward()
    # rows sent to rank d come back scaled by (d+1)
    want = torch.cat([torch.full((in_splits[0], 3), 1.0), torch.full((in_splits[1], 3), 2.0)])
    assert torch.equal(x.grad, want)


def test_dp_gradient_buckets_allreduce_world2():
    run2(_dp_buckets)


def test_all_to_all_rows_autograd_world2():
    run2(_a2a_autograd)


def test_expert_owner_placement():
    assert [parallel.expert_owner(e, 32, 8) for e in (0, 3, 4, 31)] == [0, 0, 1, 7]
    assert parallel.expert_owner(7, 8, 8) == 7
    with pytest.raises(ValueError):
        parallel.expert_owner(0, 6, 4)


def _overlapped_reducer(rank, world):
    """Bucketed all-reduce driven by post-accumulate hooks: averaged gradients equal the full-batch gradients,
    buckets fire in backward order, unused parameters are still reduced by finish(), counters re-arm."""
    torch.manual_seed(0)
    l1, l2 = torch.nn.Linear(6, 5), torch.nn.Linear(5, 3)
    unused = torch.nn.Parameter(torch.ones(2))
    fired = []
    red = parallel.OverlappedGradReducer([list(l2.parameters()), list(l1.parameters()), [unused]])
    orig = red._launch
    red._launch = lambda bi: (fired.append(bi), orig(bi))[1]
    x_all = torch.arange(4 * 6, dtype=torch.float32).view(4, 6) / 10.0
    for it in range(2):
        for p in list(l1.parameters()) + list(l2.parameters()) + [unused]:
            p.grad = None
        unused.grad = torch.full((2,), float(rank))          # a gradient no backward pass touches
        x = x_all[rank * 2:(rank + 1) * 2]
        l2(torch.tanh(l1(x))).square().mean().backward()
        red.finish()
        ref1, ref2 = torch.nn.Linear(6, 5), torch.nn.Linear(5, 3)
        ref1.load_state_dict(l1.state_dict())
        ref2.load_state_dict(l2.state_dict())
        ref2(torch.tanh(ref1(x_all))).square().mean().backward()
        for p, r in zip(list(l1.parameters()) + list(l2.parameters()), list(ref1.parameters()) + list(ref2.parameters())):
            assert torch.allclose(p.grad, r.grad, atol=1e-6), (it, p.grad, r.grad)
        assert torch.allclose(unused.grad, torch.full((2,), 0.5))
        assert fired == [0, 1, 2], fired
        fired.clear()
    red.remove()


def test_overlapped_grad_reducer_world2():
    run2(_overlapped_reducer)


def _staged_reducer(rank, world):
    """transport='p2p' host logic (CPU stand-in for the symmetric buffer): gradients that are views of flat
    per-stage buffers are staged once per buffer, reduced, and re-pointed with their offsets preserved — also when
    a flat buffer is shared by parameters of two buckets."""
    a = torch.nn.Parameter(torch.zeros(4, 3))
    b = torch.nn.Parameter(torch.zeros(5))
    c = torch.nn.Parameter(torch.zeros(2, 2))
    d = torch.nn.Parameter(torch.zeros(3))
    red = parallel.OverlappedGradReducer([[a, c], [b, d]], transport="p2p")
    red.enabled = True
    for it in range(2):
        flat = torch.arange(17, dtype=torch.float32) * (rank + 1) + it      # a and b share one flat buffer
        a.grad, b.grad = flat[:12].view(4, 3), flat[12:]
        c.grad = torch.full((2, 2), float(rank + 1))
        d.grad = torch.arange(3, dtype=torch.float32) * (rank + 2)
        red._launch(0)
        red._launch(1)
        red.finish()
        want = torch.arange(17, dtype=torch.float32) * 1.5 + it
        assert torch.allclose(a.grad, want[:12].view(4, 3)), a.grad
        assert torch.allclose(b.grad, want[12:]), b.grad
        assert torch.allclose(c.grad, torch.full((2, 2), 1.5))
        assert torch.allclose(d.grad, torch.arange(3, dtype=torch.float32) * 2.5)
    red.remove()


def test_staged_p2p_reducer_host_logic_world2():
    run2(_staged_reducer)


class _ToyLinearFn(torch.autograd.Function):
    """Mimics the package's Functions: parameter gradients are views of ONE flat buffer taken from ops.grad_buffer
    (the gradient arena when one is installed)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return x @ w.t() + b

    @staticmethod
    def backward(ctx, dy):
        from vqa_model_builder_b200 import ops
        x, w = ctx.saved_tensors
        flat = ops.grad_buffer(w.numel() + b_numel(w), x.device)
        dw = flat[:w.numel()].view_as(w)
        db = flat[w.numel():]
        dw.copy_(dy.t() @ x)
        db.copy_(dy.sum(0))
        return dy @ w, dw, db


def b_numel(w):
    return w.shape[0]


def _arena_reducer(rank, world):
    """Gradient arena + sweep reducer (CPU stand-in for the symmetric buffer): gradients are allocated back to back
    in backward order at identical offsets on both ranks; every float is reduced exactly once whatever the bucket
    boundaries; parameters whose gradients live outside the arena fall back to the library all-reduce; the arena is
    re-used from offset 0 in the next step."""
    from vqa_model_builder_b200 import ops
    torch.manual_seed(0)
    l1, l2, l3 = torch.nn.Linear(6, 5), torch.nn.Linear(5, 4), torch.nn.Linear(4, 3)
    outside = torch.nn.Parameter(torch.ones(3))                  # gradient produced by torch, not in the arena
    meter = parallel.ArenaMeter()
    ops.set_grad_arena(meter)
    x_all = torch.arange(4 * 6, dtype=torch.float32).view(4, 6) / 10.0

    def fwd(x):
        h = torch.tanh(_ToyLinearFn.apply(x, l1.weight, l1.bias))
        h = torch.tanh(_ToyLinearFn.apply(h, l2.weight, l2.bias))
        return (_ToyLinearFn.apply(h, l3.weight, l3.bias) * outside).square().mean()

    x = x_all[rank * 2:(rank + 1) * 2]
    fwd(x).backward()
    assert meter.total == 3 * 32 + 0 or meter.total % 32 == 0     # three flat buffers, each padded to 32 floats
    arena = parallel.GradArena(meter.total, device=torch.device("cpu"))
    ops.set_grad_arena(arena)
    # bucket boundaries deliberately cut across the flat buffers: (l3.weight) | (l3.bias, l2.*) | (l1.*, outside)
    red = parallel.ArenaGradReducer(arena, [[l3.weight], [l3.bias, l2.weight, l2.bias],
                                            [l1.weight, l1.bias, outside]], average=True)
    ps = list(l1.parameters()) + list(l2.parameters()) + list(l3.parameters()) + [outside]
    for it in range(2):
        for p in ps:
            p.grad = None
        arena.reset()
        fwd(x).backward()
        assert all(arena.offset_of(p.grad) >= 0 for p in ps[:-1]) and arena.offset_of(outside.grad) < 0
        ranges = list(red.ranges)
        red.finish()
        # the sweep is monotone and covers [0, arena.off) without overlap
        covered = sorted(ranges)
        assert covered[0][0] == 0 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
        assert covered[-1][1] == arena.off or arena.off == covered[-1][1]
        r1, r2, r3 = torch.nn.Linear(6, 5), torch.nn.Linear(5, 4), torch.nn.Linear(4, 3)
        for a, b in ((r1, l1), (r2, l2), (r3, l3)):
            a.load_state_dict(b.state_dict())
        ro = torch.ones(3, requires_grad=True)
        # mean over ranks of the local losses
        tot = 0.0
        for r in range(world):
            xr = x_all[r * 2:(r + 1) * 2]
            tot = tot + (r3(torch.tanh(r2(torch.tanh(r1(xr))))) * ro).square().mean() / world
        tot.backward()
        for p, q in zip(ps, list(r1.parameters()) + list(r2.parameters()) + list(r3.parameters()) + [ro]):
            assert torch.allclose(p.grad, q.grad, atol=1e-6), (it, p.grad, q.grad)
    red.remove()
    ops.set_grad_arena(None)


def test_gradient_arena_sweep_reducer_world2():
    run2(_arena_reducer)
