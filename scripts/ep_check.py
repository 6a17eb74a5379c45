#!/usr/bin/env python
"""Expert-parallel + data-parallel parity check (run under torchrun on W GPUs):
every rank also evaluates the full, unsharded MOELayer on the CONCATENATED batch; the EP layer's outputs, input
gradients, expert gradients (owner shard) and replicated gradients must match it.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/ep_check.py"""
import copy
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vqa_model_builder_b200 as pkg  # noqa: E402
from vqa_model_builder_b200 import moe, parallel  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    rank, world, local = parallel.init_distributed("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    mode = os.environ.get("EP_MODE", "fp32")
    pkg.set_compute_dtype(mode)
    tol = 2e-4 if mode == "fp32" else 2e-2
    B, S, D, F, E, K = 4, 57, 256, 512, 8, 2
    torch.manual_seed(0)
    full = moe.MOELayer(input_dim=D, hidden_dim=F, output_dim=D, num_experts=E, top_k=K, dropout=0.0).to(dev).train()
    transport = os.environ.get("EP_TRANSPORT", "nccl")
    cls = parallel.P2PExpertParallelMOELayer if transport == "p2p" else parallel.ExpertParallelMOELayer
    ep = cls(copy.deepcopy(full))
    g = torch.Generator(device="cpu").manual_seed(1)
    x_all = torch.randn(world * B, S, D, generator=g).to(dev)
    gout_all = torch.randn(world * B, S, D, generator=g).to(dev)
    # ---- reference: full layer, whole batch, loss = sum(out*gout)/W + aux  (mean-over-ranks convention) ----
    xr = x_all.clone().requires_grad_()
    out_ref = full(xr)
    ((out_ref * gout_all).sum() / world + full.get_aux_loss()).backward()
    # ---- EP: this rank's slice ----
    xs = x_all[rank * B:(rank + 1) * B].clone().requires_grad_()
    out = ep(xs)
    loss = (out * gout_all[rank * B:(rank + 1) * B]).sum() + ep.get_aux_loss()
    loss.backward()
    parallel.finish_gradients(ep.replicated_parameters(), ep.expert_parameters())
    errs = {"out": rel(out, out_ref[rank * B:(rank + 1) * B]),
            "aux": abs(float(ep.get_aux_loss()) - float(full.get_aux_loss())),
            # local loss = sum_local + aux with DDP-mean gradient convention => xs.grad / W == reference slice
            "dx": rel(xs.grad / world, xr.grad[rank * B:(rank + 1) * B])}
    lo = rank * (E // world)
    full_experts = list(full.experts)
    for i, e_loc in enumerate(ep.local.experts):
        for (n, p), (_, q) in zip(e_loc.named_parameters(), full_experts[lo + i].named_parameters()):
            errs[f"expert{lo + i}.{n}"] = rel(p.grad, q.grad)
    errs["output_norm.weight"] = rel(ep.local.output_norm.weight.grad, full.output_norm.weight.grad)
    errs["router.gate.weight"] = rel(ep.local.router.gate.weight.grad, full.router.gate.weight.grad)
    worst_key = max((k for k in errs if k != "aux"), key=lambda k: errs[k])
    ok = errs[worst_key] < tol and errs["aux"] < 1e-6
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"[rank {rank}] aux ep={float(ep.get_aux_loss()):.8f} full={float(full.get_aux_loss()):.8f} "
          f"counts={ep.last_plan.counts.tolist()}", flush=True)
    print(f"[rank {rank}] transport={transport} mode={mode} worst {worst_key}={errs[worst_key]:.3e} out={errs['out']:.3e} aux_abs={errs['aux']:.2e} "
          f"dx={errs['dx']:.3e}", flush=True)
    if rank == 0:
        print("EP_CHECK", "PASS" if int(flag) == 1 else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
