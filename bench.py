#!/usr/bin/env python
"""bench.py — fusion + MOE fwd+bwd samples/sec on N B200s (contract in the build brief, section 4).

Workload at every N: BASELINE.json configs[1] per GPU (weak scaling):
  MultimodalFusion(cross_attention, D=768, H=8, L=2) on visual [32,50,768] + text [32,64,768] (random valid lengths)
  -> MOE layer (8 experts, top-2, F=2048) on the pooled [32,1,768] vector, bf16 compute, fwd + bwd.
The MOE layer is the homogeneous-FFN MOELayer (the north star's grouped-GEMM path): the reference's `--use-moe`
VQAMOELayer fills its expert list with heterogeneous attention modules that are outside the hot-path scope and
whose classes live only in the reference tree (absent on the GPU box).

  python bench.py [--gpus N] [--steps K] [--warmup W]        # this repo's CUDA path
  python bench.py --impl reference ...                        # the reference algorithm on the host CPU (oracle port)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GEMM_DRAM_BYTES_PER_LAUNCH = 10.72e6   # measured, see roofline.traffic below
CFG = dict(B=32, T=64, V=50, D=768, H=8, L=2, E=8, K=2, F=2048, dropout=0.1)
METRIC = "fusion+MOE fwd+bwd samples/sec"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def algorithmic_flops(c) -> dict:
    """MACs per sample, 1 MAC = 2 FLOP, fwd+bwd = 3 x fwd."""
    T, V, D, L, K, F, E = c["T"], c["V"], c["D"], c["L"], c["K"], c["F"], c["E"]
    layer_gemm = 14 * T * D * D + 2 * V * D * D            # in/out projections + FFN(4D)
    layer_attn = 2 * T * D * (T + V)                       # QK^T and PV, self + cross
    pool = D * D
    moe = K * 2 * D * F + D * E                            # one token per sample
    gemm = (L * layer_gemm + pool + K * 2 * D * F) * 2 * 3
    total = (L * (layer_gemm + layer_attn) + pool + moe) * 2 * 3
    return dict(total=total, gemm=gemm)


def synth_inputs(c, rank: int):
    g = torch.Generator().manual_seed(1234 + rank)
    vis = torch.randn(c["B"], c["V"], c["D"], generator=g)
    txt = torch.randn(c["B"], c["T"], c["D"], generator=g)
    lens = torch.randint(8, c["T"] + 1, (c["B"],), generator=torch.Generator().manual_seed(4321 + rank))
    pad = ~(torch.arange(c["T"])[None, :] < lens[:, None])       # True = PAD; position 0 always valid
    return vis, txt, pad


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm (oracle port: dense MOE loop, explicit-softmax MHA) on the host CPU
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(c, threads: int):
    from oracle import init_weights
    from oracle import reference_port as rp
    torch.set_num_threads(threads)
    torch.set_flush_denormal(True)   # give the CPU arm its best case: softmax tails underflow into denormals
    torch.manual_seed(0)
    sd_f = {k: v.requires_grad_() for k, v in init_weights.multimodal_fusion_sd(c["D"], c["H"], c["L"]).items()}
    sd_m = {k: v.requires_grad_() for k, v in init_weights.moe_layer_sd(c["D"], c["F"], c["E"]).items()}
    vis, txt, pad = synth_inputs(c, 0)
    vis.requires_grad_()
    txt.requires_grad_()

    def step():
        for t in list(sd_f.values()) + list(sd_m.values()) + [vis, txt]:
            t.grad = None
        fused = rp.multimodal_fusion(sd_f, "cross_attention", c["H"], c["L"], True, vis, txt, None, pad,
                                     pdrop=c["dropout"])
        out, aux, _, _, _ = rp.moe_layer(sd_m, fused.unsqueeze(1), c["E"], c["K"], pdrop=c["dropout"])
        (out.float().square().mean() + aux).backward()

    return step


def time_cpu(step, warmup: int, steps: int):
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args, c):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    try:
        import psutil
        threads = psutil.cpu_count(logical=False) or threads
    except Exception:
        pass
    step = cpu_reference_step_factory(c, threads)
    steps = max(1, min(args.steps, 8))
    ts = time_cpu(step, max(1, min(args.warmup, 2)), steps)
    ms = 1e3 * sum(ts) / len(ts)
    val = c["B"] / (ms / 1e3)
    sample = f"{steps} full steps of the B={c['B']} workload, fp32, {threads} threads, train mode dropout {c['dropout']}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(c, 1),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(c, n):
    b = c["B"]
    tag = "configs[1]" if b == CFG["B"] else f"configs[1] shapes at saturating batch B={b} (not the headline configuration)"
    return {"workload": f"{tag}: MultimodalFusion(cross_attention D768 H8 L2) B{b} T64 V50 + MOELayer(E8 top2 F2048 "
                        f"homogeneous FFN experts) on [{b},1,768], fwd+bwd, per GPU",
            "global_batch": c["B"] * n, "per_gpu_batch": c["B"], "dropout": c["dropout"], "parallelism": f"dp{n}",
            "l2": "flushed between timed steps (256 MiB write)", "cuda_graph": True}


# ---------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args, c):
    import torch.distributed as dist

    import vqa_model_builder_b200 as pkg
    from vqa_model_builder_b200 import _lib, fusion, moe, parallel, slab

    rank, world, local = parallel.init_distributed()
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback (use --impl reference)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pkg.set_compute_dtype("bf16")
    slab.ALWAYS_REFRESH = True       # pay autocast's per-step weight cast even without an optimizer update
    torch.manual_seed(0)
    fus = fusion.MultimodalFusion(fusion.FusionConfig("cross_attention", c["D"], c["D"], c["H"], c["L"],
                                                      c["dropout"], True)).to(dev).train()
    layer = moe.MOELayer(input_dim=c["D"], hidden_dim=c["F"], output_dim=c["D"], num_experts=c["E"], top_k=c["K"],
                         dropout=c["dropout"]).to(dev).train()
    params = list(fus.parameters()) + list(layer.parameters())
    vis_h, txt_h, pad_h = [t.pin_memory() for t in synth_inputs(c, rank)]
    vis = vis_h.to(dev).requires_grad_()
    txt = txt_h.to(dev).requires_grad_()
    pad = pad_h.to(dev)
    loss_h = torch.zeros(1, dtype=torch.float32).pin_memory()
    loss_d = torch.zeros(1, dtype=torch.float32, device=dev)

    # Data parallel: gradients are all-reduced bucket by bucket on a communication stream as soon as backward has
    # produced them (MOE layer first, then fusion layer 2, then layer 1), overlapped with the rest of backward.
    reducer = None
    if world > 1:
        tail = [p for n, p in fus.named_parameters() if not n.startswith("fusion_layers.")]
        buckets = [list(layer.parameters()), tail]
        for blk in reversed(list(fus.fusion_layers)):     # backward order: FFN, cross-attention, self-attention
            buckets.append(list(blk.ffn.parameters()) + list(blk.norm3.parameters()))
            buckets.append(list(blk.cross_attn.parameters()) + list(blk.norm2.parameters()))
            buckets.append(list(blk.self_attn.parameters()) + list(blk.norm1.parameters()))
        reducer = parallel.OverlappedGradReducer(buckets, transport=args.dp_transport)

    def step():
        for p in params:
            p.grad = None
        vis.grad = None
        txt.grad = None
        fused = fus(vis, txt, text_mask=pad)
        out = layer(fused.unsqueeze(1))
        loss = out.float().square().mean() + layer.get_aux_loss()
        loss.backward()
        if reducer is not None:
            reducer.finish()
        loss_d.copy_(loss.detach().reshape(1))

    # ---- eager warm-up (also configures kernels, creates the NCCL communicator), launch count per step ----
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        _lib.reset_launch_count()
        if reducer is not None:
            reducer.enabled = False
        step()
        if reducer is not None:
            reducer.enabled = True
        torch.cuda.synchronize()
        launches_per_step = _lib.launch_count()
    torch.cuda.current_stream().wait_stream(side)
    if world > 1:
        dist.barrier()

    # The whole step, collectives included (NCCL kernels are graph-capturable), is captured in one CUDA graph.
    use_graph = not args.no_graph
    graph = None
    if use_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
        except Exception as e:  # capture unsupported for this configuration: time eagerly
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches",
                      file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    if world > 1:   # every rank must take the same path (graph or eager): agree on the slower one
        ok = torch.tensor([1 if graph is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            graph = None

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            step()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    clocks.__enter__()
    # ~1 s of load so nvidia-smi (100 ms period) sees the clocks under load.  The iteration count must be the SAME
    # on every rank (each step issues collectives), so it is fixed, not wall-clock driven.
    for _ in range(max(args.warmup, 3) + 400):
        run_step()
    barrier()

    # ---- device-resident throughput ("value") ----
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)            # evict L2 (126 MB) between timed steps; outside the timed interval
        starts[i].record()
        run_step()
        ends[i].record()
    barrier()
    dev_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(dev_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    value = c["B"] * world * args.steps / (total_ms / 1e3)

    # ---- end-to-end through the module API with host buffers ("e2e") ----
    e2e_s = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e2e_e = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    # Input upload is software-pipelined: the pinned host batch of step i+1 is copied to a device staging buffer
    # on a copy stream while step i computes (one H2D per step inside the timed intervals; step 0 pays for its own
    # and the next one).  The step itself starts with a device-to-device copy staging -> the graph's input buffers.
    copy_stream = torch.cuda.Stream()
    vis_s, txt_s, pad_s = torch.empty_like(vis), torch.empty_like(txt), torch.empty_like(pad)
    ready, free = torch.cuda.Event(), torch.cuda.Event()
    main = torch.cuda.current_stream()

    def upload():
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free)
            vis_s.copy_(vis_h, non_blocking=True)
            txt_s.copy_(txt_h, non_blocking=True)
            pad_s.copy_(pad_h, non_blocking=True)
            ready.record(copy_stream)

    barrier()
    free.record(main)
    with torch.no_grad():
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            e2e_s[i].record()
            if i == 0:
                upload()
            main.wait_event(ready)
            vis.copy_(vis_s, non_blocking=True)
            txt.copy_(txt_s, non_blocking=True)
            pad.copy_(pad_s, non_blocking=True)
            free.record(main)
            if i + 1 < args.steps:
                upload()
            with torch.enable_grad():
                run_step()
            loss_h.copy_(loss_d, non_blocking=True)
            e2e_e[i].record()
            e2e_e[i].synchronize()          # the caller reads the loss every step (training_pipeline.py:484)
    barrier()
    e2e_ms = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(e2e_s, e2e_e))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = c["B"] * world * args.steps / (float(e2e_ms.item()) / 1e3)
    clocks.__exit__()
    h2d = vis_h.numel() * 4 + txt_h.numel() * 4 + pad_h.numel()
    assert torch.isfinite(loss_h).all(), "non-finite loss"

    # ---- per-kernel timing pass (eager, CUDA events around every library call) for the roofline ----
    kern = {}
    if rank == 0:
        _lib.PROFILE = []
        if reducer is not None:
            reducer.enabled = False  # rank-0-only pass: no collectives here
        from vqa_model_builder_b200 import runtime as _rt
        _rt.set_aux_stream(False)   # one stream: per-kernel event times must not overlap each other
        for _ in range(3):
            flush.fill_(1)
            torch.cuda._sleep(60_000_000)   # park the GPU (~30 ms) so the host enqueues the whole step first:
            step()                          # events then bracket back-to-back kernels, not launch gaps
        torch.cuda.synchronize()
        for name, s, e, _ in _lib.PROFILE:
            kern.setdefault(name, []).append(s.elapsed_time(e))
        if args.detail:     # per-call table of the dense GEMMs of the last profiled step (stderr)
            calls = [(sc, s.elapsed_time(e)) for name, s, e, sc in _lib.PROFILE if name == "b200_gemm"]
            calls = calls[-(len(calls) // 3):]
            for sc, ms in calls:
                lda, al, ldb, bl, ldo, M, N, K, dt, odt, epi, act = sc[:12]
                print(f"[gemm] M={M:6d} N={N:5d} K={K:6d} layouts={al}{bl} epi={epi} out={'f32' if odt == 0 else 'bf16'} "
                      f"{ms * 1e3:8.1f} us {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", file=sys.stderr)
        _lib.PROFILE = None
        _rt.set_aux_stream(True)
    out = None
    if rank == 0:
        pk = peaks()
        fl = algorithmic_flops(c)
        gemm_ms = (sum(kern.get("b200_gemm", [])) + sum(kern.get("b200_ggemm", [])) +
                   sum(kern.get("b200_ggemm_wgrad", []))) / 3.0
        gemm_launches = (len(kern.get("b200_gemm", [])) + len(kern.get("b200_ggemm", [])) +
                         len(kern.get("b200_ggemm_wgrad", []))) // 3
        achieved = fl["gemm"] * c["B"] / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        per_kernel = {k: {"launches_per_step": len(v) // 3, "ms_per_step": sum(v) / 3.0} for k, v in kern.items()}
        # CPU baseline (oracle port) on a bounded sample: 3 full steps of the same workload
        threads = os.cpu_count() or 1
        try:
            import psutil
            threads = psutil.cpu_count(logical=False) or threads
        except Exception:
            pass
        cpu_val = None
        if not args.no_cpu_baseline:
            cstep = cpu_reference_step_factory(c, threads)
            ts = time_cpu(cstep, 1, 3)
            cpu_val = c["B"] / (sum(ts) / len(ts))
        out = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(c, world),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "pipeline": "pinned host batch i+1 uploaded on a copy stream during step i; loss read back every step"},
            "gpu_launches": int(launches_per_step) * args.steps,
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (all dense + grouped GEMM launches of a step)",
                         "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tf_sustained"],
                         # DRAM bytes (read + write) per GEMM launch, ncu, mean over the 51 launches of a step of
                         # the default configuration (profiles/r01m_launches_dram_step_cfg2.md); null otherwise
                         "traffic": GEMM_DRAM_BYTES_PER_LAUNCH if c["B"] == CFG["B"] else None,
                         "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "peak_source": pk["source"],
                         "launches_per_step": gemm_launches, "kernel_ms_per_step": gemm_ms,
                         "algorithmic_gflop_per_step": fl["gemm"] * c["B"] / 1e9},
            "cpu_baseline": {"value": cpu_val, "unit": "samples/s", "cores": threads, "kind": "port",
                             "sample": f"3 full fwd+bwd steps of the same B=32 workload (oracle port, fp32, train mode dropout {c['dropout']})"},
            "kernels": per_kernel, "cuda_graph": graph is not None,
            "algorithmic_gflop_per_step_total": fl["total"] * c["B"] / 1e9,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        # The captured graph holds NCCL kernels; tearing the communicator down under a live graph was observed to
        # block.  Drain the device, meet the other ranks, then leave without running destructors.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dp-transport", default="nccl", choices=["nccl", "p2p"],
                    help="data-parallel gradient all-reduce: NCCL (default) or the peer-memory kernel over symmetric "
                         "memory (correct, but slower at this size: DESIGN.md section 6)")
    ap.add_argument("--detail", action="store_true", help="print a per-call table of the dense GEMMs (stderr)")
    ap.add_argument("--batch", type=int, default=CFG["B"],
                    help="per-GPU batch; the default is the named configuration, larger values give the "
                         "saturating-batch roofline SURVEY 8(d) asks for beside it")
    args = ap.parse_args()
    c = dict(CFG)
    c["B"] = args.batch
    if args.impl == "reference":
        run_reference(args, c)
    else:
        run_ours(args, c)


if __name__ == "__main__":
    main()
