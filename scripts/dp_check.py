#!/usr/bin/env python
"""2+ GPUs: the peer-memory gradient all-reduce (OverlappedGradReducer transport='p2p') against NCCL on the same
step (dropout off so both passes see identical gradients).  torchrun --nproc-per-node N scripts/dp_check.py"""
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vqa_model_builder_b200 as pkg  # noqa: E402
from vqa_model_builder_b200 import fusion, moe, parallel  # noqa: E402


def main():
    rank, world, local = parallel.init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pkg.set_compute_dtype("bf16")
    torch.manual_seed(0)
    fus = fusion.MultimodalFusion(fusion.FusionConfig("cross_attention", 768, 768, 8, 2, 0.0, True)).to(dev).train()
    layer = moe.MOELayer(input_dim=768, hidden_dim=2048, output_dim=768, num_experts=8, top_k=2, dropout=0.0).to(dev).train()
    params = list(fus.parameters()) + list(layer.parameters())
    g = torch.Generator().manual_seed(100 + rank)
    vis = torch.randn(8, 50, 768, generator=g).to(dev)
    txt = torch.randn(8, 64, 768, generator=g).to(dev)
    buckets = [list(layer.parameters()), list(fus.parameters())]

    def run(transport):
        red = parallel.OverlappedGradReducer(buckets, transport=transport)
        out = []
        for _ in range(2):      # second pass re-uses the symmetric buffers
            for p in params:
                p.grad = None
            o = layer(fus(vis, txt).unsqueeze(1))
            (o.float().square().mean() + layer.get_aux_loss()).backward()
            red.finish()
            torch.cuda.synchronize()
            out = [p.grad.detach().clone() for p in params]
        red.remove()
        return out

    # expected: local gradients averaged parameter by parameter with plain collectives
    for p in params:
        p.grad = None
    o = layer(fus(vis, txt).unsqueeze(1))
    (o.float().square().mean() + layer.get_aux_loss()).backward()
    torch.cuda.synchronize()
    want, local_g = [], []
    for p in params:
        t = p.grad.detach().clone()
        local_g.append(t.clone())
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
        want.append(t)
    ref = run("nccl")
    got = run("p2p")
    worst = 0.0
    names = [n for n, _ in fus.named_parameters()] + ["moe." + n for n, _ in layer.named_parameters()]
    for n, w_, a, b in zip(names, want, ref, got):
        sc = float(w_.abs().max()) + 1e-12
        da, db = float((a - w_).abs().max()) / sc, float((b - w_).abs().max()) / sc
        if max(da, db) > 1e-6 and rank == 0:
            print(f"  mismatch {n}: nccl-vs-expected {da:.3e}  p2p-vs-expected {db:.3e}  shape {tuple(a.shape)}")
            lg = local_g[names.index(n)]
            print("    expected", w_.flatten()[:4].tolist(), "\n    p2p     ", b.flatten()[:4].tolist(),
                  "\n    local   ", lg.flatten()[:4].tolist(), "\n    p2p==local:", bool(torch.equal(b, lg)),
                  " p2p==2*expected:", bool(torch.allclose(b, 2 * w_, rtol=1e-5, atol=1e-9)))
            for n2, w2 in zip(names, want):      # does the wrong tensor equal some other parameter's gradient?
                if w2.shape == b.shape and n2 != n and torch.allclose(b, w2, rtol=1e-5, atol=1e-9):
                    print("    p2p result equals the expected gradient of", n2)
        worst = max(worst, da, db)
    ok = worst < 1e-6
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"DP_CHECK {'PASS' if int(flag.item()) else 'FAIL'} world={world} worst rel diff p2p vs nccl = {worst:.3e}")
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    import os
    os._exit(0)


if __name__ == "__main__":
    main()
