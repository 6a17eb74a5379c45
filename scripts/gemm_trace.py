#!/usr/bin/env python
"""Where does a small GEMM's time go?  Needs the instrumented library:
    B200VQA_GEMM_TRACE=1 python -m vqa_model_builder_b200._build && B200VQA_GEMM_TRACE=1 python scripts/gemm_trace.py
A chain of identical GEMMs is captured in a CUDA graph (programmatic dependent launch edges, like a bench step) and
replayed; CTA 0 of every launch records %globaltimer at: 0 kernel entry, 1 setup done (barriers, TMEM), 2 after
griddepcontrol.wait, 3 first k-block landed, 4 last MMA committed, 5 epilogue sees the accumulator, 6 first tile stored,
7 CTA done, 8 / 9 / 10 first chunk of the first tile: accumulator in registers / bias added / stores issued.  Printed: per-launch chain time (events) and the median phase durations in microseconds."""
import ctypes
import os
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
assert os.environ.get("B200VQA_GEMM_TRACE") == "1", "set B200VQA_GEMM_TRACE=1"
from vqa_model_builder_b200 import _lib, ops  # noqa: E402
from vqa_model_builder_b200._lib import (ACT_GELU, EPI_ACCUM, EPI_ACT_D, EPI_NONE, LAYOUT_K, LAYOUT_MN)  # noqa: E402


def main():
    dev = "cuda"
    lib = _lib.load()
    lib.b200_debug_gemm_trace.restype = ctypes.c_int
    lib.b200_debug_gemm_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    buf = (ctypes.c_ulonglong * (256 * 64))()
    n = ctypes.c_uint(0)
    g = torch.Generator(device=dev).manual_seed(0)
    cases = [("fwd_proj_M2048", "KK", 2048, 768, 768, EPI_NONE), ("fwd_proj_M32", "KK", 32, 768, 768, EPI_NONE),
             ("fwd_ffn2_M32", "KK", 32, 768, 3072, EPI_NONE), ("fwd_ffn1_M32", "KK", 32, 3072, 768, EPI_ACT_D),
             ("fwd_qkv_M2048", "KK", 2048, 2304, 768, EPI_NONE), ("fwd_ffn1_M2048", "KK", 2048, 3072, 768, EPI_ACT_D),
             ("fwd_ffn2_M2048", "KK", 2048, 768, 3072, EPI_NONE), ("dgrad_proj_M2048", "KMN", 2048, 768, 768, EPI_NONE),
             ("wgrad_proj_M2048", "MNMN", 768, 768, 2048, EPI_ACCUM)]
    CH = 24
    for name, lay, M, N, K, epi in cases:
        a = torch.randn((M, K) if lay != "MNMN" else (K, M), generator=g, device=dev).to(torch.bfloat16)
        b = torch.randn((N, K) if lay == "KK" else (K, N), generator=g, device=dev).to(torch.bfloat16) * 0.05
        out = torch.empty((M, N), dtype=torch.float32 if lay == "MNMN" else torch.bfloat16, device=dev)
        aux = torch.empty((M, N), dtype=torch.bfloat16, device=dev) if epi == EPI_ACT_D else None
        bias = torch.randn(N, generator=g, device=dev) if lay == "KK" else None
        al = LAYOUT_MN if lay == "MNMN" else LAYOUT_K
        bl = LAYOUT_K if lay == "KK" else LAYOUT_MN

        def run():
            ops.gemm(a, al, b, bl, M, N, K, out=out, bias=bias, epi=epi, act=ACT_GELU, aux_out=aux)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(3):
                run()
            st.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for _ in range(CH):
                    run()
            gr.replay()
            st.synchronize()
            lib.b200_debug_gemm_trace(buf, ctypes.byref(n))          # reset
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            gr.replay()
            e1.record(st)
            st.synchronize()
        chain_us = e0.elapsed_time(e1) * 1e3 / CH
        lib.b200_debug_gemm_trace(buf, ctypes.byref(n))
        rows = [[buf[i * 64 + k] for k in range(64)] for i in range(min(n.value, 256))]
        rows.sort(key=lambda r: r[0])
        mid = rows[4:-2]
        ph = lambda i, j: statistics.median((r[j] - r[i]) / 1e3 for r in mid)   # noqa: E731
        gap = statistics.median((rows[k + 1][2] - rows[k][7]) / 1e3 for k in range(4, len(rows) - 3))
        period = statistics.median((rows[k + 1][7] - rows[k][7]) / 1e3 for k in range(4, len(rows) - 3))
        clk = lambda i, j: statistics.median((r[16 + j] - r[16 + i]) for r in mid)   # noqa: E731
        print(f"{name:18s} M={M} N={N} K={K} epi={epi}: chain {chain_us:6.2f} us/launch (graph of {CH}), "
              f"exit-to-exit {period:5.2f} | entry->setup {ph(0, 1):5.2f} setup->dep-wait-done {ph(1, 2):5.2f} "
              f"wait->first-kblock {ph(2, 3):5.2f} mainloop {ph(3, 4):5.2f} commit->epilogue-sees {ph(4, 5):5.2f} "
              f"epilogue(first tile) {ph(5, 6):5.2f} ->cta-done {ph(6, 7):5.2f} | prev-exit -> this wait-done {gap:5.2f} "
              f"| clocks: mainloop {clk(3, 4):.0f} epilogue {clk(5, 6):.0f} [first chunk: tmem-ld {clk(5, 8):.0f} bias {clk(8, 9):.0f} "
              f"stage+store {clk(9, 10):.0f}]", flush=True)
        nkb = min(32, (K + 63) // 64)
        r = mid[len(mid) // 2]
        print("    k-block arrivals (clocks after the first, one launch): " +
              " ".join(str(int(r[32 + k] - r[32])) for k in range(nkb) if r[32 + k] >= r[32]), flush=True)


if __name__ == "__main__":
    main()
