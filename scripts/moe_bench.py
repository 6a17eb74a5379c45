#!/usr/bin/env python
"""MOE layer at config-5 scale (B=128 samples x 114 tokens = 14 592 tokens, D=768, F=2048, E=8, top-2), bf16,
fwd+bwd: per-kernel CUDA-event times (GPU parked while the host enqueues, so events bracket back-to-back kernels)
and the roofline of every kernel family:
  tensor: grouped expert GEMMs (algorithmic FLOPs = 3 * 2 * N*K_top * 2*D*F)        vs measured bf16 peak
  HBM   : router / permute / combine(+LN) / per-expert LN (algorithmic bytes) vs measured copy bandwidth
Usage: python scripts/moe_bench.py [--batch 128] [--iters 5]"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vqa_model_builder_b200 as pkg  # noqa: E402
from vqa_model_builder_b200 import _lib, moe  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--seq", type=int, default=114)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--experts", type=int, default=8)
    args = ap.parse_args()
    dev = "cuda"
    D, F, E, K = 768, 2048, args.experts, 2
    B, S = args.batch, args.seq
    N = B * S
    pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else \
        {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    pkg.set_compute_dtype("bf16")
    torch.manual_seed(0)
    layer = moe.MOELayer(input_dim=D, hidden_dim=F, output_dim=D, num_experts=E, top_k=K, dropout=0.0).to(dev).train()
    x = torch.randn(B, S, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    gout = torch.randn(B, S, D, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        for p in layer.parameters():
            p.grad = None
        x.grad = None
        out = layer(x)
        ((out * gout).sum() + layer.get_aux_loss()).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    # whole-layer time (eager launches queued behind a parked GPU)
    ts = []
    for _ in range(args.iters):
        flush.fill_(0)
        torch.cuda._sleep(100_000_000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    from vqa_model_builder_b200 import runtime as _rt
    _rt.set_aux_stream(False)        # one stream: per-call event times must not overlap each other
    _lib.PROFILE = []
    for _ in range(args.iters):
        flush.fill_(0)
        torch.cuda._sleep(100_000_000)
        step()
    torch.cuda.synchronize()
    kern = {}
    for name, s, e, scal in _lib.PROFILE:
        kern.setdefault(name, []).append(s.elapsed_time(e))
    _lib.PROFILE = None
    per = {k: sum(v) / args.iters for k, v in kern.items()}
    cnt = {k: len(v) // args.iters for k, v in kern.items()}
    NK = N * K
    es = 2  # bf16
    # algorithmic work per step (fwd + bwd), SURVEY 8(d)
    work = {
        "b200_ggemm": ("tensor", 2.0 * NK * D * F * 2 * 2),            # fwd fc1,fc2 + dgrad fc2,fc1
        "b200_ggemm_wgrad": ("tensor", 2.0 * NK * D * F * 2),          # dW1, dW2
        "b200_router_fwd": ("hbm", N * (D * es + 4 * E + 8 * K + 4)),
        "b200_router_bwd": ("hbm", N * (2 * D * es + 4 * E + 8 * K) + N * D * es),   # dx kernel + wgrad pass re-reads x
        "b200_moe_permute": ("hbm", N * D * es + NK * D * es),
        "b200_moe_unpermute": ("hbm", NK * D * es + N * D * es),
        "b200_moe_combine_fwd": ("hbm", NK * D * es + 4 * NK + N * D * es),
        "b200_moe_combine_bwd": ("hbm", N * D * es + 2 * NK * D * es + 8 * NK),      # dout, z rows, dz rows
        "b200_add_ln_fwd": ("hbm", 3 * NK * D * es),                                   # y2, xp, z
        "b200_add_ln_bwd": ("hbm", 4 * NK * D * es),                                   # dz, y2, xp, dr
        "b200_colsum": ("hbm", NK * (D + F) * es),
    }
    rows = []
    for k, ms in sorted(per.items(), key=lambda kv: -kv[1]):
        bound, amount = work.get(k, (None, None))
        if bound == "tensor":
            ach = amount / (ms / 1e3) / 1e12
            rows.append((k, cnt[k], ms, f"{ach:8.1f} TFLOP/s", f"{ach / pk['bf16_tflops_sustained']:.3f} of measured tensor peak"))
        elif bound == "hbm":
            ach = amount / (ms / 1e3) / 1e9
            rows.append((k, cnt[k], ms, f"{ach:8.1f} GB/s", f"{ach / pk['hbm_gbs']:.3f} of measured HBM peak"))
        else:
            rows.append((k, cnt[k], ms, "", ""))
    total_ms = sorted(ts)[len(ts) // 2]
    flops = 3 * 2.0 * NK * 2 * D * F
    print(f"MOELayer fwd+bwd  tokens={N} (B={B} x S={S}) E={E} top{K} D={D} F={F} bf16: {total_ms:.3f} ms/step, "
          f"{B / (total_ms / 1e3):.0f} samples/s, {flops / (total_ms / 1e3) / 1e12:.1f} TFLOP/s layer-level")
    print(f"{'kernel':26s} {'n':>3s} {'ms/step':>9s}  achieved")
    for k, n, ms, a, f in rows:
        print(f"{k:26s} {n:3d} {ms:9.4f}  {a}  {f}")
    print(json.dumps({"ms_per_step": total_ms, "kernels": per}))


if __name__ == "__main__":
    main()
