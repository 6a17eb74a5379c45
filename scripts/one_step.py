#!/usr/bin/env python
"""One eager fwd+bwd step of a bench.py configuration between cudaProfilerStart/Stop — the window ncu captures with
`--profile-from-start off` (launch lists per step for profiles/).  usage: one_step.py [--config C] [--batch B]"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import vqa_model_builder_b200 as pkg  # noqa: E402
from vqa_model_builder_b200 import runtime, slab  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--moe-parallel", default="ep")
    ap.add_argument("--single-stream", action="store_true", help="serialise the auxiliary stream (cleaner per-kernel view)")
    args = ap.parse_args()
    c = dict(bench.CONFIGS[args.config])
    if args.batch:
        c["B"] = args.batch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    pkg.set_compute_dtype("bf16")
    slab.ALWAYS_REFRESH = True
    if args.single_stream:
        runtime.set_aux_stream(False)
    wl = bench.Workload(c, dev, 0, 1, args)
    for _ in range(3):
        wl.step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    wl.step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("one step done, loss", float(wl.loss_d))


if __name__ == "__main__":
    main()
