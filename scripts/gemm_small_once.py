#!/usr/bin/env python
"""A few launches of the CLS-row FFN2 GEMM (M = 32, N = 768, K = 3072, bias + residual epilogue) for ncu: with the
default settings this is a thread-block-cluster split-K launch (kc = 8)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vqa_model_builder_b200 import ops  # noqa: E402
from vqa_model_builder_b200._lib import EPI_ADD, LAYOUT_K  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
M, N, K = 32, 768, 3072
a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
b = (torch.randn(N, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16)
res = torch.randn(M, N, generator=g, device="cuda").to(torch.bfloat16)
bias = torch.randn(N, generator=g, device="cuda")
for _ in range(6):
    out = ops.gemm(a, LAYOUT_K, b, LAYOUT_K, M, N, K, bias=bias, epi=EPI_ADD, aux_in=res)
torch.cuda.synchronize()
ref = a.double() @ b.double().t() + bias.double() + res.double()
print("rel err", float((out.double() - ref).norm() / ref.norm()))
