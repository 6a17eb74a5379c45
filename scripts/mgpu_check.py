#!/usr/bin/env python
"""2+ GPUs (torchrun --nproc-per-node N scripts/mgpu_check.py): the two multi-GPU mechanisms against plain collectives.

  DP : gradient arena in symmetric memory + in-place all-reduce (NVLS multimem when the allocation has a multicast
       mapping, else peer loads) vs. parameter-by-parameter NCCL all-reduce of the same step (dropout off).
  EP : P2PExpertParallelMOELayer (fused dispatch / return over NVLink, experts sharded) vs. the unsharded MOELayer on
       the concatenated batch: outputs, aux loss, input gradients, expert gradients (owner shard), replicated gradients
       — eager and replayed from a CUDA graph."""
import copy
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vqa_model_builder_b200 as pkg  # noqa: E402
from vqa_model_builder_b200 import fusion, moe, ops, parallel  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def dp_check(rank, world, dev):
    pkg.set_compute_dtype("bf16")
    torch.manual_seed(0)
    fus = fusion.MultimodalFusion(fusion.FusionConfig("cross_attention", 768, 768, 8, 2, 0.0, True)).to(dev).train()
    layer = moe.MOELayer(input_dim=768, hidden_dim=2048, output_dim=768, num_experts=8, top_k=2, dropout=0.0).to(dev).train()
    params = list(fus.parameters()) + list(layer.parameters())
    g = torch.Generator().manual_seed(100 + rank)
    vis = torch.randn(8, 50, 768, generator=g).to(dev)
    txt = torch.randn(8, 64, 768, generator=g).to(dev)

    def step():
        for p in params:
            p.grad = None
        o = layer(fus(vis, txt).unsqueeze(1))
        (o.float().square().mean() + layer.get_aux_loss()).backward()

    step()
    torch.cuda.synchronize()
    want = []
    for p in params:
        t = p.grad.detach().clone()
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
        want.append(t)
    meter = parallel.ArenaMeter()
    ops.set_grad_arena(meter)
    step()
    arena = parallel.GradArena(meter.total + 1024, device=dev)
    ops.set_grad_arena(arena)
    buckets = [list(layer.parameters()) + [p for n, p in fus.named_parameters() if not n.startswith("fusion_layers.")]]
    for blk in reversed(list(fus.fusion_layers)):
        buckets.append(list(blk.parameters()))
    red = parallel.ArenaGradReducer(arena, buckets, average=True)
    worst = 0.0
    for it in range(3):
        arena.reset()
        step()
        red.finish()
        torch.cuda.synchronize()
        inside = sum(arena.offset_of(p.grad) >= 0 for p in params)
        for p, w_ in zip(params, want):
            worst = max(worst, float((p.grad - w_).abs().max()) / (float(w_.abs().max()) + 1e-12))
    red.remove()
    ops.set_grad_arena(None)
    return dict(worst=worst, multicast=bool(arena.multicast_ptr), inside=inside, params=len(params),
                arena_floats=arena.off, overflow=arena.overflow)


def ep_check(rank, world, dev, mode, graph):
    pkg.set_compute_dtype(mode)
    tol = 2e-4 if mode == "fp32" else 2e-2
    B, S, D, F, E, K = 4, 57, 256, 512, 8, 2
    torch.manual_seed(0)
    full = moe.MOELayer(input_dim=D, hidden_dim=F, output_dim=D, num_experts=E, top_k=K, dropout=0.0).to(dev).train()
    ep = parallel.P2PExpertParallelMOELayer(copy.deepcopy(full), max_tokens=B * S)
    g = torch.Generator(device="cpu").manual_seed(1)
    x_all = torch.randn(world * B, S, D, generator=g).to(dev)
    gout_all = torch.randn(world * B, S, D, generator=g).to(dev)
    if mode == "bf16":
        x_all = x_all.to(torch.bfloat16).float()
    xr = x_all.clone().requires_grad_()
    out_ref = full(xr)
    ((out_ref * gout_all).sum() / world + full.get_aux_loss()).backward()
    xs = x_all[rank * B:(rank + 1) * B].clone().requires_grad_()
    gs = gout_all[rank * B:(rank + 1) * B]
    eparams = list(ep.parameters())
    out_buf = torch.zeros(B, S, D, device=dev)

    def step():
        for p in eparams:
            p.grad = None
        xs.grad = None
        out = ep(xs)
        # mean-over-ranks convention with SUM reduction: local loss / W
        (((out * gs).sum() + ep.get_aux_loss()) / world).backward()
        out_buf.copy_(out.detach())

    step()
    if graph:
        step()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            step()
        for _ in range(3):
            gr.replay()
    torch.cuda.synchronize()
    for p in ep.replicated_parameters():                  # replicated: sum over ranks
        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
    errs = {"out": rel(out_buf, out_ref[rank * B:(rank + 1) * B]),
            "dx": rel(xs.grad, xr.grad[rank * B:(rank + 1) * B])}
    aux_err = abs(float(ep.get_aux_loss()) - float(full.get_aux_loss()))
    lo = rank * (E // world)
    full_experts = list(full.experts)
    for i, e_loc in enumerate(ep.local.experts):
        for (n, p), (_, q) in zip(e_loc.named_parameters(), full_experts[lo + i].named_parameters()):
            errs[f"expert{lo + i}.{n}"] = rel(p.grad, q.grad)
    errs["output_norm.weight"] = rel(ep.local.output_norm.weight.grad, full.output_norm.weight.grad)
    errs["router.gate.weight"] = rel(ep.local.router.gate.weight.grad, full.router.gate.weight.grad)
    wk = max(errs, key=lambda k: errs[k])
    ok = errs[wk] < tol and aux_err < 1e-6
    return dict(ok=ok, worst=(wk, errs[wk]), aux_err=aux_err, mode=mode, graph=graph)


def main():
    rank, world, local = parallel.init_distributed("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    # never the legacy default stream: AccumulateGrad nodes created on it (and kept alive by aux_outputs / gradient
    # hooks) would synchronise with it inside a CUDA-graph capture and invalidate the capture
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    results = {}
    for name, fn in (("dp_arena", lambda: dp_check(rank, world, dev)),
                     ("ep_fp32", lambda: ep_check(rank, world, dev, "fp32", False)),
                     ("ep_bf16", lambda: ep_check(rank, world, dev, "bf16", False)),
                     ("ep_bf16_graph", lambda: ep_check(rank, world, dev, "bf16", True))):
        try:
            results[name] = fn()
        except Exception as e:
            import traceback
            results[name] = dict(ok=False, error=f"{type(e).__name__}: {e}", tb=traceback.format_exc()[-1500:])
        torch.cuda.synchronize()
        dist.barrier()
    dp = results["dp_arena"]
    ok = all(r.get("ok", True) for r in results.values()) and "error" not in dp and dp.get("worst", 1.0) < 1e-5
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        for k, v in results.items():
            print(k, v)
        print(f"MGPU_CHECK {'PASS' if int(flag.item()) else 'FAIL'} world={world}")
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
