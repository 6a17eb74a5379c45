#!/usr/bin/env python
"""bench.py — fusion + MOE fwd+bwd samples/sec on N B200s (contract in the build brief, section 4).

Headline workload at every N (weak scaling, per GPU): BASELINE.json configs[1]
  MultimodalFusion(cross_attention, D=768, H=8, L=2) on visual [32,50,768] + text [32,64,768] (random valid lengths)
  -> MOE layer (8 experts, top-2, F=2048) on the pooled [32,1,768] vector, bf16 compute, fwd + bwd, dropout 0.1.
The MOE layer is the homogeneous-FFN MOELayer (the north star's grouped-GEMM path): the reference's `--use-moe`
VQAMOELayer fills its expert list with heterogeneous attention modules that are outside the hot-path scope and
whose classes live only in the reference tree (absent on the GPU box).

`--config {1,3,4,5}` selects the other BASELINE.json configurations (same JSON line); by default the headline line
also carries short runs of configs 3, 4 and 5 under "other_configs" (`--other-configs none` switches that off).

On N > 1 GPUs (one process per GPU, torchrun): the batch is data parallel; the FUSION parameters are replicated and
their gradients all-reduced over NVLink (through the NVSwitch with multimem instructions on a gradient arena in
symmetric memory, overlapped with backward); the EXPERTS are sharded expert-parallel (expert e on rank e // (E/N)),
tokens exchanged by the fused dispatch / return kernels over NVLink peer memory, so expert gradients need no
all-reduce.  The first step of every multi-GPU run is checked against the unsharded layer ("ep_check").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C]       # this repo's CUDA path
  python bench.py --impl reference ...                                     # the reference algorithm on the host CPU
"""
from __future__ import annotations

import argparse
import copy
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

METRIC = "fusion+MOE fwd+bwd samples/sec"
COMMON = dict(D=768, H=8, L=2, K=2, F=2048, T=64, dropout=0.1)
CONFIGS = {
    1: dict(COMMON, kind="classification", tag="configs[1]", B=32, V=50, E=8,
            desc="MultimodalFusion(cross_attention D768 H8 L2) T64 V50 (CLIP ViT-B/32) + MOELayer(E8 top2 F2048) on the pooled [B,1,768]"),
    3: dict(COMMON, kind="classification", tag="configs[2]", B=32, V=257, E=16,
            desc="DINOv2-B token count V=257, cross-attention stack (the reference has no MCAN: fusion_type='mcan' falls "
                 "back to add, SURVEY F8) + MOELayer(E16 top2 F2048), 32 samples per GPU (global 256 on 8 GPUs)"),
    4: dict(COMMON, kind="classification", tag="configs[3]", B=32, V=49, E=32,
            desc="Swin-B token count V=49 + PhoBERT-large projected to D768, cross-attention + MOELayer(E32 top2 F2048), "
                 "experts expert-parallel over the GPUs"),
    5: dict(COMMON, kind="generative", tag="configs[4]", B=128, V=50, E=8,
            desc="CrossModalFusion (2 pre-LN encoder layers over [visual;question] = 114 tokens, ff 2048) + MOELayer(E8 "
                 "top2 F2048) on all B*114 tokens; decoder excluded (out of scope, SURVEY 8(d))"),
    # configs[4] followed by the answer decoder and the label-smoothed cross-entropy (the generative
    # pipeline's whole trainable path behind the encoders); not a BASELINE.json configuration, selected explicitly
    6: dict(COMMON, kind="generative", tag="configs[4] + decoder (SURVEY 8(f) N2)", B=128, V=50, E=8,
            dec=dict(L=6, Ta=64, vocab=64000, smoothing=0.1, moe_loss_weight=0.01),
            desc="CrossModalFusion + MOELayer(E8 top2 F2048) on B*114 tokens -> TransformerDecoder (6 pre-LN layers, "
                 "causal self-attention + cross-attention to the 114 fused tokens, 64 answer tokens, tied 64000-way "
                 "projection) -> label-smoothed cross-entropy"),
}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def measured_traffic(cfg_id: int, batch: int):
    """DRAM bytes per GEMM launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the GEMM launches of
    one step) recorded from the last committed ncu launch list; None when no capture exists for this workload."""
    p = ROOT / "profiles" / "gemm_dram_traffic.json"
    if not p.exists():
        return None, None
    d = json.loads(p.read_text())
    rec = d.get(f"config{cfg_id}_B{batch}")
    return (rec["bytes_per_launch"], rec["source"]) if rec else (None, None)


def algorithmic_flops(c) -> dict:
    """MACs per sample, 1 MAC = 2 FLOP, fwd+bwd = 3 x fwd.  `done` counts the work the kernels actually
    execute (last-layer dead-row elimination: only the CLS row of the last fusion layer is consumed,
    vqa_model.py:386-387); `full` is the reference's count.  roofline.achieved uses `done`."""
    T, V, D, L, K, F, E = c["T"], c["V"], c["D"], c["L"], c["K"], c["F"], c["E"]
    if c["kind"] == "generative":
        S = V + T
        layer_gemm = 4 * S * D * D + 2 * S * D * F
        layer_attn = 2 * S * S * D
        moe_g = S * K * 2 * D * F
        gemm = (L * layer_gemm + moe_g) * 6
        total = (L * (layer_gemm + layer_attn) + moe_g + S * D * E) * 6
        if c.get("dec"):
            d = c["dec"]
            Ta, Ld, Vv = d["Ta"], d["L"], d["vocab"]
            dec_gemm = Ld * (6 * Ta * D * D + 2 * S * D * D + 2 * Ta * D * F) + Ta * D * Vv
            dec_attn = Ld * (2 * Ta * Ta * D + 2 * Ta * S * D)
            gemm += dec_gemm * 6
            total += (dec_gemm + dec_attn) * 6
        return dict(total=total, gemm=gemm, total_full=total, gemm_full=gemm)
    layer_gemm = 14 * T * D * D + 2 * V * D * D            # in/out projections + FFN(4D)
    layer_attn = 2 * T * D * (T + V)                       # QK^T and PV, self + cross
    pool = D * D
    moe = K * 2 * D * F + D * E                            # one token per sample
    gemm_full = (L * layer_gemm + pool + K * 2 * D * F) * 6
    total_full = (L * (layer_gemm + layer_attn) + pool + moe) * 6
    from vqa_model_builder_b200.fusion import cross_modal
    if getattr(cross_modal, "DEAD_ROW_ELIMINATION", False) and L >= 1:
        # last layer: self-attention K/V projection of all T rows + the image K/V projection stay; everything that
        # only feeds rows 1..T-1 of the layer output is skipped (Q, out-proj, FFN, LayerNorms for one row)
        last_gemm = 2 * T * D * D + 2 * V * D * D + 12 * D * D
        last_attn = 2 * D * (T + V)
        gemm = ((L - 1) * layer_gemm + last_gemm + pool + K * 2 * D * F) * 6
        total = ((L - 1) * (layer_gemm + layer_attn) + last_gemm + last_attn + pool + moe) * 6
    else:
        gemm, total = gemm_full, total_full
    return dict(total=total, gemm=gemm, total_full=total_full, gemm_full=gemm_full)


def synth_inputs(c, rank: int):
    """Encoder output features.  Under the configuration's bf16 AMP the encoders' projection Linears
    (vqa_model.py:128-129, 231-232) hand the fusion bf16 tensors, so the synthetic features are bf16 (the CPU arm
    computes in fp32 on the same bf16-representable values)."""
    g = torch.Generator().manual_seed(1234 + rank)
    vis = torch.randn(c["B"], c["V"], c["D"], generator=g).to(torch.bfloat16)
    txt = torch.randn(c["B"], c["T"], c["D"], generator=g).to(torch.bfloat16)
    lens = torch.randint(8, c["T"] + 1, (c["B"],), generator=torch.Generator().manual_seed(4321 + rank))
    pad = ~(torch.arange(c["T"])[None, :] < lens[:, None])       # True = PAD; position 0 always valid
    return vis, txt, pad


def synth_answers(c, rank: int):
    """Teacher-forcing inputs of the decoder: answer token ids, their attention mask (1 = token) and the labels
    (-100 on padding), generative_vqa_model.py:540-587."""
    d = c["dec"]
    g = torch.Generator().manual_seed(777 + rank)
    ids = torch.randint(0, d["vocab"], (c["B"], d["Ta"]), generator=g)
    labels = torch.randint(0, d["vocab"], (c["B"], d["Ta"]), generator=g)
    lens = torch.randint(4, d["Ta"] + 1, (c["B"],), generator=g)
    mask = (torch.arange(d["Ta"])[None, :] < lens[:, None]).long()
    labels[mask == 0] = -100
    return ids, mask, labels


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


def host_threads() -> int:
    threads = os.cpu_count() or 1
    try:
        import psutil
        threads = psutil.cpu_count(logical=False) or threads
    except Exception:
        pass
    return threads


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm (oracle port: dense MOE loop, explicit-softmax MHA) on the host CPU
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(c, threads: int, batch: int):
    """One fwd+bwd of the workload's algorithm as the reference computes it (dense MOE: every expert on every token),
    fp32, train mode with the workload's dropout, on `batch` samples."""
    from oracle import init_weights
    from oracle import reference_port as rp
    torch.set_num_threads(threads)
    torch.set_flush_denormal(True)   # give the CPU arm its best case: softmax tails underflow into denormals
    torch.manual_seed(0)
    D, H, L, E, K, F = c["D"], c["H"], c["L"], c["E"], c["K"], c["F"]
    cb = dict(c, B=batch)
    vis, txt, pad = synth_inputs(cb, 0)
    vis, txt = vis.float().requires_grad_(), txt.float().requires_grad_()
    sd_m = {k: v.requires_grad_() for k, v in init_weights.moe_layer_sd(D, F, E).items()}
    if c["kind"] == "generative":
        sd_f = {k: v.requires_grad_() for k, v in init_weights.cross_modal_fusion_sd(D, H, L, F).items()}
        sd_f.update({"moe_layer." + k: v for k, v in sd_m.items()})
        leaves = list(sd_f.values()) + [vis, txt]
        if c.get("dec"):
            d = c["dec"]
            sd_d = {k: (v.requires_grad_() if v.dtype.is_floating_point and k != "pos_encoding.pe" else v)
                    for k, v in init_weights.decoder_sd(D, H, d["L"], F, d["vocab"], d["Ta"]).items()}
            ids, amask, labels = synth_answers(cb, 0)
            enc_mask = torch.cat([torch.ones(batch, c["V"]), (~pad).float()], dim=1)
            leaves += [v for v in sd_d.values() if v.requires_grad]

            def step_dec():
                for t in leaves:
                    t.grad = None
                out, aux = rp.cross_modal_fusion(sd_f, H, L, vis, txt, ~pad, moe=dict(num_experts=E, top_k=K),
                                                 pdrop=c["dropout"])
                logits = rp.transformer_decoder(sd_d, H, d["L"], out, ids, enc_mask, amask, pdrop=c["dropout"])
                loss = rp.smoothed_cross_entropy(logits, labels, -100, d["smoothing"]) + d["moe_loss_weight"] * aux
                loss.backward()
            return step_dec

        def step():
            for t in leaves:
                t.grad = None
            out, aux = rp.cross_modal_fusion(sd_f, H, L, vis, txt, ~pad, moe=dict(num_experts=E, top_k=K),
                                             pdrop=c["dropout"])
            (out.float().square().mean() + aux).backward()
        return step
    sd_f = {k: v.requires_grad_() for k, v in init_weights.multimodal_fusion_sd(D, H, L).items()}
    leaves = list(sd_f.values()) + list(sd_m.values()) + [vis, txt]

    def step():
        for t in leaves:
            t.grad = None
        fused = rp.multimodal_fusion(sd_f, "cross_attention", H, L, True, vis, txt, None, pad, pdrop=c["dropout"])
        out, aux, _, _, _ = rp.moe_layer(sd_m, fused.unsqueeze(1), E, K, pdrop=c["dropout"])
        (out.float().square().mean() + aux).backward()

    return step


def cpu_sample_batch(c) -> int:
    """Bounded CPU sample: the full batch for the classification configs, 8 samples (x114 tokens through 8 dense
    experts) for the generative one — scaled linearly to samples/s."""
    return c["B"] if c["kind"] == "classification" else min(c["B"], 8)


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
    threads = host_threads()
    bs = cpu_sample_batch(c)
    step = cpu_reference_step_factory(c, threads, bs)
    steps = max(1, args.steps)
    ts = time_cpu(step, max(0, min(args.warmup, 2)), steps)
    ms = 1e3 * sum(ts) / len(ts)
    val = bs / (ms / 1e3)
    sample = (f"{steps} fwd+bwd steps on {bs} of the {c['B']} samples of the workload's batch (dense reference algorithm, "
              f"oracle port, fp32, {threads} threads, train mode dropout {c['dropout']}); samples/s scales linearly")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(c, args.config, 1, "none"),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(c, cfg_id, n, moe_parallel):
    b = c["B"]
    named = b == CONFIGS[cfg_id]["B"]
    tag = c["tag"] if named else f"{c['tag']} shapes at batch B={b} per GPU (not the named configuration)"
    par = f"dp{n}" if moe_parallel != "ep" else f"dp{n} (fusion, router) + ep{n} (experts: {c['E'] // n} per GPU, fused NVLink dispatch)"
    return {"workload": f"{tag}: {c['desc']}, B={b} per GPU, fwd+bwd", "config_id": cfg_id,
            "global_batch": b * n, "per_gpu_batch": b, "dropout": c["dropout"], "parallelism": par,
            "l2": "flushed between timed steps (256 MiB write)", "cuda_graph": True}


# ---------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------
class Workload:
    """Modules + one training step (fwd + bwd [+ gradient exchange]) of a configuration on this rank."""

    def __init__(self, c, dev, rank, world, args):
        from vqa_model_builder_b200 import fusion, moe, ops, parallel
        self.c, self.dev, self.rank, self.world = c, dev, rank, world
        self.parallel, self.ops = parallel, ops
        torch.manual_seed(0)                       # same initial weights on every rank
        D, H, L, E, K, F = c["D"], c["H"], c["L"], c["E"], c["K"], c["F"]
        self.moe_parallel = "none"
        if world > 1:
            self.moe_parallel = "ep" if (args.moe_parallel == "ep" and E % world == 0) else "dp"
        self.ep_check = None
        if c["kind"] == "generative":
            cfg = fusion.GenerativeFusionConfig(fusion_dim=D, fusion_num_heads=H, fusion_num_layers=L,
                                                fusion_dropout=c["dropout"], decoder_ff_dim=F, use_moe=True,
                                                moe_type="standard", num_experts=E, num_experts_per_token=K)
            self.fus = fusion.CrossModalFusion(cfg).to(dev).train()
            self.fus.return_aux_tensor = True      # the reference's .item() on the aux loss is a host sync per step
            layer = self.fus.moe_layer
            self.dec = None
            if c.get("dec"):
                from vqa_model_builder_b200.decoder import DecoderConfig, TransformerDecoder
                d = c["dec"]
                self.dec = TransformerDecoder(DecoderConfig(
                    vocab_size=d["vocab"], decoder_hidden_dim=D, decoder_num_layers=d["L"], decoder_num_heads=H,
                    decoder_ff_dim=F, decoder_dropout=c["dropout"], max_answer_length=d["Ta"],
                    label_smoothing=d["smoothing"])).to(dev).train()
                ids, amask, labels = synth_answers(c, rank)
                self.ans = (ids.to(dev), amask.to(dev), labels.to(dev))
        else:
            self.fus = fusion.MultimodalFusion(fusion.FusionConfig("cross_attention", D, D, H, L, c["dropout"],
                                                                   True)).to(dev).train()
            layer = moe.MOELayer(input_dim=D, hidden_dim=F, output_dim=D, num_experts=E, top_k=K,
                                 dropout=c["dropout"]).to(dev).train()
        self.full_copy = None
        if self.moe_parallel == "ep":
            self.full_copy = copy.deepcopy(layer).eval()             # unsharded twin for the first-step check
            tokens = c["B"] * (c["V"] + c["T"] if c["kind"] == "generative" else 1)
            layer = parallel.P2PExpertParallelMOELayer(layer, max_tokens=tokens)
            if c["kind"] == "generative":
                self.fus.moe_layer = layer
        self.layer = layer
        vis_h, txt_h, pad_h = [t.pin_memory() for t in synth_inputs(c, rank)]
        self.host = (vis_h, txt_h, pad_h)
        self.vis = vis_h.to(dev).requires_grad_()
        self.txt = txt_h.to(dev).requires_grad_()
        self.pad = pad_h.to(dev)
        self.loss_d = torch.zeros(1, dtype=torch.float32, device=dev)
        if c["kind"] == "generative":
            self.replicated = [p for n, p in self.fus.named_parameters() if ".experts." not in n]
            self.sharded = list(layer.expert_parameters()) if self.moe_parallel == "ep" else []
            if self.moe_parallel != "ep":
                self.replicated = list(self.fus.parameters())
            if self.dec is not None:
                self.replicated = self.replicated + list(self.dec.parameters())
        else:
            rep_moe = list(layer.replicated_parameters()) if self.moe_parallel == "ep" else list(layer.parameters())
            self.replicated = list(self.fus.parameters()) + rep_moe
            self.sharded = list(layer.expert_parameters()) if self.moe_parallel == "ep" else []
        self.params = self.replicated + self.sharded
        import vqa_model_builder_b200 as pkg
        self.pkg = pkg
        self.prefetch = not getattr(args, "no_prefetch", False)
        self.later_modules = [self.layer] + ([self.dec] if getattr(self, "dec", None) is not None else [])
        self.reducer = None
        self.arena = None

    # -- gradient exchange -----------------------------------------------------------------------------------------
    def buckets(self):
        """parameters in the order their gradients become complete during backward"""
        c, fus, layer = self.c, self.fus, self.layer
        moe_rep = [p for p in self.replicated if any(p is q for q in layer.parameters())]
        if c["kind"] == "generative":
            out = []
            if self.dec is not None:       # backward reaches the decoder first: tied projection, then layers 5..0
                dec = self.dec
                out.append(list(dec.layer_norm.parameters()) + [dec.embedding.weight])
                for blk in reversed(list(dec.decoder.layers)):
                    out.append(list(blk.parameters()))
            out.append(list(fus.layer_norm.parameters()) + moe_rep)
            for blk in reversed(list(fus.layers)):
                out.append(list(blk.parameters()))
            return out
        tail = [p for n, p in fus.named_parameters() if not n.startswith("fusion_layers.")]
        out = [moe_rep + tail]
        for blk in reversed(list(fus.fusion_layers)):     # backward order: FFN, cross-attention, self-attention
            out.append(list(blk.ffn.parameters()) + list(blk.norm3.parameters()) +
                       list(blk.cross_attn.parameters()) + list(blk.norm2.parameters()))
            out.append(list(blk.self_attn.parameters()) + list(blk.norm1.parameters()))
        return out

    def setup_gradient_exchange(self, transport: str):
        if self.world == 1:
            return
        parallel, ops = self.parallel, self.ops
        if transport == "nccl":
            self.reducer = parallel.OverlappedGradReducer(self.buckets(), average=False)
            return
        meter = parallel.ArenaMeter()                 # one eager step tells how many floats a step's gradients need
        ops.set_grad_arena(meter)
        self.step()
        torch.cuda.synchronize()
        self.arena = parallel.GradArena(meter.total + 4096, device=self.dev)
        ops.set_grad_arena(self.arena)
        self.reducer = parallel.ArenaGradReducer(self.arena, self.buckets(), average=False)

    # -- one step ----------------------------------------------------------------------------------------------------
    def forward_loss(self):
        c = self.c
        if c["kind"] == "generative":
            out, aux = self.fus(self.vis, self.txt, (~self.pad).long())
            if self.dec is not None:
                from vqa_model_builder_b200 import ops
                d = c["dec"]
                ids, amask, labels = self.ans
                enc_mask = torch.cat([torch.ones(c["B"], c["V"], dtype=torch.long, device=self.dev),
                                      (~self.pad).long()], dim=1)
                logits = self.dec(out, ids, encoder_attention_mask=enc_mask, decoder_attention_mask=amask)
                ce = ops.cross_entropy(logits.view(-1, logits.shape[-1]), labels.view(-1), -100, d["smoothing"])
                return ce + d["moe_loss_weight"] * aux
            return out.float().square().mean() + aux
        fused = self.fus(self.vis, self.txt, text_mask=self.pad)
        out = self.layer(fused.unsqueeze(1))
        return out.float().square().mean() + self.layer.get_aux_loss()

    def step(self):
        for p in self.params:
            p.grad = None
        self.vis.grad = None
        self.txt.grad = None
        if self.arena is not None:
            self.arena.reset()
        if self.prefetch:
            # modules that run later in the step get their bf16 weight copies refreshed on a side stream while the
            # fusion's first kernels run (the public API a trainer calls after optimizer.step())
            self.pkg.prefetch_compute_weights(*self.later_modules)
        loss = self.forward_loss()
        if self.world > 1:
            # mean-over-ranks loss with SUM gradient exchange: replicated gradients are summed by the all-reduce,
            # the sharded experts' gradients are complete locally (every token routed to an expert reached its owner)
            (loss / self.world).backward()
        else:
            loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.loss_d.copy_(loss.detach().reshape(1))

    # -- first-step check of the sharded layer against its unsharded twin ---------------------------------------------
    def check_expert_parallel(self):
        """The EP layer on this rank's tokens must equal the unsharded layer on the concatenated batch (same weights
        on every rank): outputs, input gradients, bit-equal expert indices."""
        import torch.distributed as dist
        if self.moe_parallel != "ep":
            return None
        c, W = self.c, self.world
        S = c["V"] + c["T"] if c["kind"] == "generative" else 1
        g = torch.Generator(device="cpu").manual_seed(77 + self.rank)
        x = torch.randn(c["B"], S, c["D"], generator=g).to(self.dev).to(torch.bfloat16)
        gout = torch.randn(c["B"], S, c["D"], generator=g).to(self.dev)
        xs = [torch.empty_like(x) for _ in range(W)]
        dist.all_gather(xs, x)
        was = self.layer.training
        self.layer.eval()                                   # dropout off on both sides
        xl = x.clone().requires_grad_()
        out = self.layer(xl)
        (out.float() * gout).sum().backward()
        x_all = torch.cat(xs, dim=0).requires_grad_()
        ref = self.full_copy(x_all)
        lo = self.rank * c["B"]
        (ref[lo:lo + c["B"]].float() * gout).sum().backward()
        # the twin's input gradient holds only this rank's loss term; the EP layer's too (other ranks' terms flow
        # into THEIR inputs) -> directly comparable on this rank's slice, expert path + router path
        rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
        K = c["K"]
        idx_ep = self.layer.last_plan.idx.view(-1, K)
        idx_ref = self.full_copy.last_plan.idx.view(-1, K)[lo * S:(lo + c["B"]) * S]
        res = {"out_rel_err": rel(out, ref[lo:lo + c["B"]]), "dx_rel_err": rel(xl.grad, x_all.grad[lo:lo + c["B"]]),
               "indices_equal": bool(torch.equal(idx_ep, idx_ref)),
               "aux_abs_err": abs(float(self.layer.get_aux_loss()) - float(self.full_copy.get_aux_loss()))}
        ok = res["out_rel_err"] < 2e-2 and res["dx_rel_err"] < 2e-2 and res["indices_equal"] and res["aux_abs_err"] < 1e-5
        flag = torch.tensor([1 if ok else 0], device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        res["status"] = "pass" if int(flag.item()) == 1 else "FAIL"
        self.layer.train(was)
        for p in self.layer.parameters():
            p.grad = None
        self.full_copy = None                                # free the twin
        torch.cuda.empty_cache()
        self.ep_check = res
        return res


def measure(wl: Workload, args, steps: int, warm_iters: int, want_e2e: bool, want_kernels: bool):
    """Times `steps` graph-replayed steps (L2 flushed between them), optionally the e2e pipeline and the per-entry-point
    kernel table.  Returns a dict of raw measurements (rank-max already applied)."""
    import torch.distributed as dist
    from vqa_model_builder_b200 import _lib
    from vqa_model_builder_b200 import runtime as _rt
    dev, world, rank, c = wl.dev, wl.world, wl.rank, wl.c

    # ---- eager warm-up (configures kernels, sizes the symmetric buffers), launch count per step ----
    for _ in range(3):
        wl.step()
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    wl.step()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count()
    if world > 1:
        dist.barrier()

    # The whole step, collectives included, is captured in one CUDA graph.
    graph = None
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=torch.cuda.current_stream()):
                wl.step()
        except Exception as e:  # capture unsupported for this configuration: time eagerly
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches",
                      file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    if world > 1:   # every rank must take the same path (graph or eager)
        ok = torch.tensor([1 if graph is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            graph = None

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            wl.step()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # load so nvidia-smi (100 ms period) sees the clocks under load; fixed count (every step issues collectives)
    for _ in range(warm_iters):
        run_step()
    barrier()

    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    barrier()
    for i in range(steps):
        flush.fill_(i & 0xFF)            # evict L2 (126 MB) between timed steps; outside the timed interval
        starts[i].record()
        run_step()
        ends[i].record()
    barrier()
    total_ms = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(starts, ends))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    res = dict(ms_per_step=total_ms / steps, value=c["B"] * world * steps / (total_ms / 1e3),
               launches_per_step=int(launches_per_step), cuda_graph=graph is not None)

    if want_e2e:
        # Input upload is software-pipelined: the pinned host batch of step i+1 is copied to a device staging buffer
        # on a copy stream while step i computes (one H2D per step inside the timed intervals; step 0 pays for its
        # own and the next one).  The step itself starts with a device-to-device copy staging -> the graph's inputs.
        vis_h, txt_h, pad_h = wl.host
        loss_h = torch.zeros(1, dtype=torch.float32).pin_memory()
        e2e_s = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e2e_e = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        copy_stream = torch.cuda.Stream()
        vis_s, txt_s, pad_s = torch.empty_like(wl.vis), torch.empty_like(wl.txt), torch.empty_like(wl.pad)
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
            for i in range(steps):
                flush.fill_(i & 0xFF)
                e2e_s[i].record()
                if i == 0:
                    upload()
                main.wait_event(ready)
                wl.vis.copy_(vis_s, non_blocking=True)
                wl.txt.copy_(txt_s, non_blocking=True)
                wl.pad.copy_(pad_s, non_blocking=True)
                free.record(main)
                if i + 1 < steps:
                    upload()
                with torch.enable_grad():
                    run_step()
                loss_h.copy_(wl.loss_d, non_blocking=True)
                e2e_e[i].record()
                e2e_e[i].synchronize()          # the caller reads the loss every step (training_pipeline.py:484)
        barrier()
        e2e_ms = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(e2e_s, e2e_e))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        res["e2e_value"] = c["B"] * world * steps / (float(e2e_ms.item()) / 1e3)
        res["h2d"] = vis_h.numel() * vis_h.element_size() + txt_h.numel() * txt_h.element_size() + pad_h.numel()
        assert torch.isfinite(loss_h).all(), "non-finite loss"

    # ---- per-kernel timing pass (eager, CUDA events around every library call) for the roofline ----
    # Runs on every rank (each step issues cross-rank barriers / collectives); rank 0's table is reported.
    if want_kernels:
        kern = {}
        _lib.PROFILE = []
        _rt.set_aux_stream(False)   # one stream: per-kernel event times must not overlap each other
        for _ in range(3):
            flush.fill_(1)
            torch.cuda._sleep(60_000_000)   # park the GPU (~30 ms) so the host enqueues the whole step first:
            wl.step()                       # events then bracket back-to-back kernels, not launch gaps
        torch.cuda.synchronize()
        for name, s, e, _ in _lib.PROFILE:
            kern.setdefault(name, []).append(s.elapsed_time(e))
        if args.detail and rank == 0:     # per-call table of the dense GEMMs of the last profiled step (stderr)
            calls = [(sc, s.elapsed_time(e)) for name, s, e, sc in _lib.PROFILE if name == "b200_gemm"]
            calls = calls[-(len(calls) // 3):]
            for sc, ms in calls:
                lda, al, ldb, bl, ldo, M, N, K = sc[:8]
                print(f"[gemm] M={M:6d} N={N:5d} K={K:6d} layouts={al}{bl} epi={sc[10]} "
                      f"{ms * 1e3:8.1f} us {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", file=sys.stderr)
        _lib.PROFILE = None
        _rt.set_aux_stream(True)
        res["kernels"] = {k: {"launches_per_step": len(v) // 3, "ms_per_step": sum(v) / 3.0} for k, v in kern.items()}
    del graph
    return res


def roofline_of(c, cfg_id, kernels, pk):
    fl = algorithmic_flops(c)
    names = ("b200_gemm", "b200_ggemm", "b200_ggemm_wgrad")
    gemm_ms = sum(kernels.get(n, {}).get("ms_per_step", 0.0) for n in names)
    launches = sum(kernels.get(n, {}).get("launches_per_step", 0) for n in names)
    achieved = fl["gemm"] * c["B"] / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    traffic, src = measured_traffic(cfg_id, c["B"])
    return {"bound": "tensor", "kernel": "gemm_tc_kernel (all dense + grouped GEMM launches of a step)",
            "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
            "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
            "traffic_source": src, "peak_source": pk["source"] + " (sustained bf16, MEASURED_PEAKS.json)",
            "launches_per_step": launches, "kernel_ms_per_step": gemm_ms,
            "algorithmic_gflop_per_step": fl["gemm"] * c["B"] / 1e9,
            "algorithmic_gflop_per_step_without_dead_row_elimination": fl["gemm_full"] * c["B"] / 1e9}


def hbm_rooflines(c, kernels, pk):
    """Achieved fraction of the measured HBM bandwidth for the MOE row kernels (SURVEY 8(d) algorithmic bytes; only
    meaningful when the MOE layer sees many tokens, i.e. the generative configuration)."""
    if c["kind"] != "generative":
        return None
    N = c["B"] * (c["V"] + c["T"])
    D, E, K, F, es = c["D"], c["E"], c["K"], c["F"], 2
    NK = N * K
    work = {"b200_router_fwd": N * (D * es + 4 * E + 8 * K + 4),
            "b200_router_bwd": N * (2 * D * es + 4 * E + 8 * K) + N * D * es,
            "b200_moe_permute": N * D * es + NK * D * es, "b200_moe_unpermute": NK * D * es + N * D * es,
            "b200_moe_combine_fwd": NK * D * es + 4 * NK + N * D * es,
            "b200_moe_combine_bwd": N * D * es + 2 * NK * D * es + 8 * NK}
    out = {}
    for k, nbytes in work.items():
        ms = kernels.get(k, {}).get("ms_per_step")
        if ms:
            gbs = nbytes / (ms / 1e3) / 1e9
            out[k] = {"achieved_gbs": gbs, "frac_of_measured_hbm": gbs / pk["hbm"], "ms_per_step": ms}
    return out


def run_ours(args, c):
    import torch.distributed as dist

    import vqa_model_builder_b200 as pkg
    from vqa_model_builder_b200 import parallel, slab

    # CPU baseline (oracle port) on a bounded sample: N=1 only, before any GPU work or process group exists
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    cpu_val, threads, cpu_bs = None, host_threads(), cpu_sample_batch(c)
    if world_env == 1 and not args.no_cpu_baseline and torch.cuda.is_available():
        cstep = cpu_reference_step_factory(c, threads, cpu_bs)
        ts = time_cpu(cstep, 1, 3)
        cpu_val = cpu_bs / (sum(ts) / len(ts))
        del cstep

    rank, world, local = parallel.init_distributed()
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback (use --impl reference)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pkg.set_compute_dtype("bf16")
    slab.ALWAYS_REFRESH = True       # pay autocast's per-step weight cast even without an optimizer update
    # Nothing below runs on the legacy default stream: autograd's AccumulateGrad nodes remember the stream they were
    # created on (the gradient hooks and aux_outputs keep them alive across steps), and a node created on the default
    # stream would synchronise with it inside the CUDA-graph capture, which invalidates the capture.
    # The step runs (and is captured) on a HIGH-priority stream: the library's auxiliary streams (weight / bias
    # gradients) have the default, lower priority, so when a critical-path dgrad GEMM and a wgrad GEMM become ready
    # together the block scheduler hands the free SMs to the dgrad first (kernel nodes keep the priority in a graph).
    torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1 if args.main_priority == "high" else 0))

    wl = Workload(c, dev, rank, world, args)
    ep_res = wl.check_expert_parallel() if world > 1 else None
    wl.setup_gradient_exchange(args.dp_transport)

    clocks = ClockSampler(local)
    clocks.__enter__()
    m = measure(wl, args, args.steps, max(args.warmup, 3) + 400, want_e2e=True, want_kernels=True)
    clocks.__exit__()
    pk = peaks()
    transport = None
    if world > 1:
        transport = "nccl" if wl.arena is None else ("nvls multimem all-reduce on the gradient arena" if wl.arena.multicast_ptr
                                                     else "peer-load all-reduce on the gradient arena (no multicast mapping)")

    # ---- short runs of the other configurations (same process, fresh modules) ----
    others = {}
    if args.other_configs != "none":
        from vqa_model_builder_b200 import ops as _ops
        for oc in [int(x) for x in args.other_configs.split(",") if x]:
            if oc == args.config or oc not in CONFIGS:
                continue
            try:
                if wl is not None:
                    if wl.reducer is not None:
                        wl.reducer.remove()
                    _ops.set_grad_arena(None)
                    wl = None
                    torch.cuda.empty_cache()
                c2 = dict(CONFIGS[oc])
                w2 = Workload(c2, dev, rank, world, args)
                ep2 = w2.check_expert_parallel() if world > 1 else None
                w2.setup_gradient_exchange(args.dp_transport)
                m2 = measure(w2, args, max(3, min(args.steps, 10)), 30, want_e2e=False, want_kernels=True)
                rec = {"config": workload_config(c2, oc, world, w2.moe_parallel), "value": m2["value"],
                       "unit": "samples/s", "ms_per_step": m2["ms_per_step"], "steps": max(3, min(args.steps, 10)),
                       "gpu_launches_per_step": m2["launches_per_step"], "cuda_graph": m2["cuda_graph"],
                       "roofline": roofline_of(c2, oc, m2["kernels"], pk), "ep_check": ep2,
                       "hbm_kernels": hbm_rooflines(c2, m2["kernels"], pk), "kernels": m2["kernels"]}
                others[str(oc)] = rec
                if w2.reducer is not None:
                    w2.reducer.remove()
                _ops.set_grad_arena(None)
                del w2
                torch.cuda.empty_cache()
            except Exception as e:      # a secondary record must never take the headline line down
                others[str(oc)] = {"error": f"{type(e).__name__}: {e}"[:300]}
                if world > 1:
                    break               # ranks may have diverged: stop issuing collectives

    if rank == 0:
        out = {
            "metric": METRIC, "value": m["value"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(c, args.config, world, wl.moe_parallel if wl is not None else
                                      ("ep" if world > 1 and args.moe_parallel == "ep" and c["E"] % world == 0 else "dp")),
            "clocks": clocks.summary(),
            "e2e": {"value": m["e2e_value"], "unit": "samples/s", "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": 4,
                    "pipeline": "pinned host batch i+1 uploaded on a copy stream during step i; loss read back every step"},
            "gpu_launches": m["launches_per_step"] * args.steps, "gpu_launches_per_step": m["launches_per_step"],
            "roofline": roofline_of(c, args.config, m["kernels"], pk),
            "cpu_baseline": {"value": cpu_val, "unit": "samples/s", "cores": threads, "kind": "port",
                             "sample": (f"3 fwd+bwd steps on {cpu_bs} of the {c['B']} samples of the same workload (dense "
                                        f"reference algorithm, oracle port, fp32, train mode dropout {c['dropout']}), run "
                                        "before the GPU timing; N=1 only") if cpu_val is not None else
                                       "not run (N>1: reported at N=1 only)"},
            "kernels": m["kernels"], "hbm_kernels": hbm_rooflines(c, m["kernels"], pk), "cuda_graph": m["cuda_graph"],
            "algorithmic_gflop_per_step_total": algorithmic_flops(c)["total"] * c["B"] / 1e9,
            "ep_check": ep_res, "gradient_exchange": transport, "other_configs": others,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        # The captured graphs hold collective kernels; tearing the communicator down under a live graph was observed
        # to block.  Drain the device, meet the other ranks, then leave without running destructors.
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
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration: 1 = configs[1] (headline), 3 = configs[2] (V=257, E=16), "
                         "4 = configs[3] (E=32, expert parallel), 5 = configs[4] (generative fusion + MOE on B*114 tokens)")
    ap.add_argument("--other-configs", default="3,4,5",
                    help="comma list of further configurations measured briefly and reported under 'other_configs' "
                         "('none' to skip)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--main-priority", default="high", choices=["high", "normal"],
                    help="CUDA priority of the stream the step runs on (auxiliary streams are always 'normal')")
    ap.add_argument("--no-prefetch", action="store_true",
                    help="do not refresh the later modules' bf16 weight copies on a side stream (A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dp-transport", default="arena", choices=["arena", "nccl"],
                    help="data-parallel gradient all-reduce: gradient arena in symmetric memory reduced through the "
                         "NVSwitch / peer loads (default), or bucketed NCCL")
    ap.add_argument("--moe-parallel", default="ep", choices=["ep", "dp"],
                    help="N>1: experts sharded expert-parallel (default, when E divides by N) or replicated")
    ap.add_argument("--detail", action="store_true", help="print a per-call table of the dense GEMMs (stderr)")
    ap.add_argument("--batch", type=int, default=None,
                    help="per-GPU batch; the default is the named configuration, larger values give the "
                         "saturating-batch roofline SURVEY 8(d) asks for beside it")
    args = ap.parse_args()
    c = dict(CONFIGS[args.config])
    if args.batch is not None:
        c["B"] = args.batch
    if args.impl == "reference":
        run_reference(args, c)
    else:
        run_ours(args, c)


if __name__ == "__main__":
    main()
