#!/usr/bin/env python
"""Expert-parallel MOE layer throughput (fwd+bwd) under torchrun: NCCL all-to-all transport vs the fused peer-memory
dispatch, config-5 token count per rank (B=128/W... here --batch per rank), bf16.  Device-timed (max over ranks).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/ep_bench.py"""
import argparse
import copy
import json
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vqa_model_builder_b200 as pkg  # noqa: E402
from vqa_model_builder_b200 import moe, parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)     # samples per rank (x 114 tokens)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = parallel.init_distributed("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    pkg.set_compute_dtype("bf16")
    D, F, E, K, S = 768, 2048, 8, 2, 114
    torch.manual_seed(0)
    full = moe.MOELayer(input_dim=D, hidden_dim=F, output_dim=D, num_experts=E, top_k=K, dropout=0.0).to(dev).train()
    x = torch.randn(args.batch, S, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    gout = torch.randn(args.batch, S, D, device=dev, dtype=torch.bfloat16)
    res = {}
    for name, cls in (("nccl_all_to_all", parallel.ExpertParallelMOELayer), ("p2p_fused", parallel.P2PExpertParallelMOELayer)):
        layer = cls(copy.deepcopy(full))

        def step():
            for p in layer.parameters():
                p.grad = None
            x.grad = None
            out = layer(x)
            ((out * gout).sum() + layer.get_aux_loss()).backward()
            parallel.finish_gradients(layer.replicated_parameters(), layer.expert_parameters())

        for _ in range(5):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            step()
        e.record()
        e.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / args.steps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[name] = float(ms)
        dist.barrier()
    if rank == 0:
        tok = args.batch * S * world
        for k, v in res.items():
            print(f"{k:18s} {v:8.3f} ms/step  {args.batch * world / (v / 1e3):10.0f} samples/s  ({tok} tokens/step, W={world})")
        print(json.dumps({"world": world, "batch_per_rank": args.batch, "ms_per_step": res}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
