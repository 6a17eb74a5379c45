#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM through the C-ABI: the dense shapes of the fusion block (config 1/2, M=2048
tokens), a saturating batch (M=65536) and the grouped expert FFN of config 5.  CUDA events, L2 flushed between
launches.  Usage: python scripts/gemm_bench.py [--only NAME] [--iters N]"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vqa_model_builder_b200 import _lib, ops  # noqa: E402
from vqa_model_builder_b200._lib import (ACT_GELU, EPI_ACCUM, EPI_ACT, EPI_DACT, EPI_NONE, LAYOUT_K, LAYOUT_MN)  # noqa


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    cases = []
    for M in (2048, 65536):
        cases += [(f"fwd_qkv_M{M}", "KK", M, 2304, 768, EPI_NONE), (f"fwd_proj_M{M}", "KK", M, 768, 768, EPI_NONE),
                  (f"fwd_ffn1_M{M}", "KK", M, 3072, 768, EPI_ACT), (f"fwd_ffn2_M{M}", "KK", M, 768, 3072, EPI_NONE),
                  (f"dgrad_ffn2_M{M}", "KMN", M, 3072, 768, EPI_DACT), (f"dgrad_ffn1_M{M}", "KMN", M, 768, 3072, EPI_NONE),
                  (f"wgrad_ffn1_M{M}", "MNMN", 3072, 768, M, EPI_ACCUM), (f"wgrad_proj_M{M}", "MNMN", 768, 768, M, EPI_ACCUM)]
    M = 14592          # config 5: 128 samples x 114 tokens, encoder FFN 768 <-> 2048
    cases += [(f"fwd_qkv_M{M}", "KK", M, 2304, 768, EPI_NONE), (f"fwd_ffn1_M{M}", "KK", M, 2048, 768, EPI_ACT),
              (f"fwd_ffn2_M{M}", "KK", M, 768, 2048, EPI_NONE), (f"dgrad_ffn2_M{M}", "KMN", M, 2048, 768, EPI_DACT),
              (f"dgrad_ffn1_M{M}", "KMN", M, 768, 2048, EPI_NONE), (f"wgrad_ffn1_M{M}", "MNMN", 2048, 768, M, EPI_ACCUM)]
    res = {}
    for name, lay, M, N, K, epi in cases:
        if args.only and args.only not in name:
            continue
        a = torch.randn((M, K) if lay != "MNMN" else (K, M), generator=g, device=dev).to(torch.bfloat16)
        b = torch.randn((N, K) if lay == "KK" else (K, N), generator=g, device=dev).to(torch.bfloat16) * 0.05
        out_dtype = torch.float32 if lay == "MNMN" else torch.bfloat16
        out = torch.empty((M, N), dtype=out_dtype, device=dev)
        aux = torch.randn((M, N), generator=g, device=dev).to(torch.bfloat16) if epi in (EPI_ACT, EPI_DACT) else None
        bias = torch.randn(N, generator=g, device=dev) if lay == "KK" else None
        al = LAYOUT_MN if lay == "MNMN" else LAYOUT_K
        bl = LAYOUT_K if lay == "KK" else LAYOUT_MN

        def run():
            ops.gemm(a, al, b, bl, M, N, K, out=out, bias=bias, epi=epi, act=ACT_GELU,
                     aux_in=aux if epi == EPI_DACT else None, aux_out=aux if epi == EPI_ACT else None)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            if not args.no_flush:
                flush.fill_(0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            run()
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        ms = ts[len(ts) // 2]
        tf = 2.0 * M * N * K / (ms / 1e3) / 1e12
        res[name] = dict(ms=ms, tflops=tf)
        print(f"{name:22s} M={M:6d} N={N:5d} K={K:6d}  {ms * 1e3:9.1f} us  {tf:8.1f} TFLOP/s", flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
