#!/usr/bin/env python
"""Attention core fwd+bwd at shapes with more than 128 query rows (vision-token queries, single-stream fusion) and the
fusion block's own shapes: tensor-core kernels vs the CUDA-core (SIMT) kernels they replace for those shapes.
Run twice: `python scripts/attn_bench.py` and `B200VQA_ATTN=simt python scripts/attn_bench.py`.  CUDA events around
fwd and bwd separately, median of --iters, L2 flushed between iterations."""
import argparse
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vqa_model_builder_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = "cuda"
    mode = os.environ.get("B200VQA_ATTN", "tc")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    shapes = [(32, 8, 64, 50, 768, False), (128, 8, 114, 114, 768, False), (32, 8, 197, 64, 768, False),
              (32, 8, 257, 64, 768, False), (32, 8, 64, 257, 768, False), (32, 8, 328, 328, 768, False),
              (32, 8, 328, 328, 768, True)]
    for B, H, T, S, D, causal in shapes:
        q = torch.randn(B * T, D, generator=g, device=dev).to(torch.bfloat16).requires_grad_()
        kv = torch.randn(B * S, 2 * D, generator=g, device=dev).to(torch.bfloat16).requires_grad_()
        gout = torch.randn(B * T, D, generator=g, device=dev).to(torch.bfloat16)
        tf, tb = [], []
        for it in range(args.iters + 2):
            flush.fill_(0)
            q.grad = kv.grad = None
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            o = ops.AttentionFn.apply(q, kv, None, B, T, S, H, False, None, causal)
            e[1].record()
            o.backward(gout)
            e[2].record()
            e[2].synchronize()
            if it >= 2:
                tf.append(e[0].elapsed_time(e[1]))
                tb.append(e[1].elapsed_time(e[2]))
        tf.sort()
        tb.sort()
        f, b = tf[len(tf) // 2] * 1e3, tb[len(tb) // 2] * 1e3
        flops = 4.0 * B * H * T * S * (D // H)
        print(f"[{mode}] B={B:4d} H={H} T={T:4d} S={S:4d} dh={D // H} causal={int(causal)}: fwd {f:8.1f} us "
              f"({flops / f / 1e6:7.1f} TFLOP/s)  bwd {b:8.1f} us ({2.5 * flops / b / 1e6:7.1f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    main()
