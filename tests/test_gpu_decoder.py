"""GPU: SURVEY 8(f) N2 — the generative answer decoder (causal self-attention, cross-attention to the fused memory,
tied 64 000-way projection) and the label-smoothed cross-entropy, through the C-ABI:
  * against the golden of the reference's own TransformerDecoder + nn.CrossEntropyLoss (fp32 1e-4), and against the
    golden-pinned oracle on bf16-representable values (bf16 1e-2), forward + backward incl. the tied embedding gradient;
  * the streaming cross-entropy at the real vocabulary size against torch;
  * embedding gather + positions + dropout and its scatter-add backward against torch (mask materialised);
  * decoder shapes of the benchmark (D=768, H=8, T=64, S=114) against the oracle."""
import numpy as np
import pytest
import torch

from conftest import bf16_representable, leafs, load_golden, rel_err, round_sd_for_bf16
from oracle import reference_port as rp
from oracle.init_weights import seeded_state_dict

pytestmark = pytest.mark.gpu

from vqa_model_builder_b200 import _lib, ops  # noqa: E402
from vqa_model_builder_b200.decoder import DecoderConfig, FusedCrossEntropyLoss, TransformerDecoder  # noqa: E402
import vqa_model_builder_b200 as pkg  # noqa: E402

DEV = "cuda"
MODES = [("fp32", 1e-4), ("bf16", 1e-2)]


class computing:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        pkg.set_compute_dtype(self.mode)

    def __exit__(self, *a):
        pkg.set_compute_dtype("auto")


def _oracle(sd, H, L, mem0, ids, mem_mask, tgt_mask, labels, smoothing):
    sdr, mr = leafs(sd), mem0.clone().requires_grad_()
    logits = rp.transformer_decoder(sdr, H, L, mr, ids, mem_mask, tgt_mask)
    loss = rp.smoothed_cross_entropy(logits, labels, -100, smoothing)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sdr.items() if v.requires_grad}
    grads["embedding.weight"] = grads["embedding.weight"] + grads.pop("output_projection.weight")
    return dict(logits=logits.detach(), loss=loss.detach(), d_memory=mr.grad, grads=grads)


def _run(dec, sd, mem0, ids, mem_mask, tgt_mask, labels, smoothing):
    dec.load_state_dict(sd)
    dec.to(DEV).train()
    mem = mem0.to(DEV).requires_grad_()
    logits = dec(mem, ids.to(DEV), encoder_attention_mask=mem_mask.to(DEV), decoder_attention_mask=tgt_mask.to(DEV))
    loss = FusedCrossEntropyLoss(-100, smoothing)(logits.view(-1, logits.shape[-1]), labels.to(DEV).view(-1))
    loss.backward()
    return logits, loss, mem.grad, {k: p.grad for k, p in dec.named_parameters()}


@pytest.mark.parametrize("mode,tol", MODES)
def test_decoder_matches_reference_golden(mode, tol):
    g = load_golden("generative_decoder")
    B, T, S, D, H, L, F, V = [int(v) for v in g["cfg"]]
    cfg = DecoderConfig(vocab_size=V, decoder_hidden_dim=D, decoder_num_layers=L, decoder_num_heads=H, decoder_ff_dim=F,
                        decoder_dropout=0.0, max_answer_length=16, label_smoothing=0.1)
    sd, mem0 = g["sd"], g["memory"]
    ref = dict(logits=g["logits"], loss=g["loss"], d_memory=g["d_memory"], grads=g["grads"])
    if mode == "bf16":
        sd, mem0 = round_sd_for_bf16(sd), bf16_representable(mem0)
        ref = _oracle(sd, H, L, mem0, g["ids"].long(), g["mem_mask"], g["tgt_mask"], g["labels"].long(), 0.1)
    with computing(mode):
        dec = TransformerDecoder(cfg)
        assert set(dec.state_dict().keys()) == set(g["sd"].keys())
        logits, loss, dmem, grads = _run(dec, sd, mem0, g["ids"].long(), g["mem_mask"], g["tgt_mask"],
                                         g["labels"].long(), 0.1)
    assert logits.shape == (B, T, V)
    assert rel_err(logits, ref["logits"]) < tol, rel_err(logits, ref["logits"])
    assert abs(float(loss) - float(ref["loss"])) < tol * max(1.0, abs(float(ref["loss"])))
    assert rel_err(dmem, ref["d_memory"]) < tol, rel_err(dmem, ref["d_memory"])
    errs = {k: rel_err(grads[k], v) for k, v in ref["grads"].items() if k in grads and float(v.norm()) > 1e-6}
    assert "embedding.weight" in errs and len(errs) >= 38
    bad = {k: e for k, e in errs.items() if e > (tol if mode == "fp32" else 2e-2)}
    assert not bad, bad


@pytest.mark.parametrize("mode,tol", MODES)
def test_decoder_benchmark_shape_vs_oracle(mode, tol):
    """D=768, H=8 (d_h=96), T=64 answer tokens, S=114 fused tokens (the 128-row attention flavour for the
    cross-attention, the 64-row flavour for the causal self-attention), 2 layers, 1 000-way vocabulary."""
    B, T, S, D, H, L, F, V = 3, 64, 114, 768, 8, 2, 2048, 1000
    cfg = DecoderConfig(vocab_size=V, decoder_hidden_dim=D, decoder_num_layers=L, decoder_num_heads=H, decoder_ff_dim=F,
                        decoder_dropout=0.0, max_answer_length=64, label_smoothing=0.1)
    rng = np.random.default_rng(5)
    with computing(mode):
        dec = TransformerDecoder(cfg)
        sd = seeded_state_dict(dec.state_dict(), 17, keys=[k for k in dec.state_dict() if k != "pos_encoding.pe"])
        sd["pos_encoding.pe"] = dec.state_dict()["pos_encoding.pe"].clone()
        sd["output_projection.weight"] = sd["embedding.weight"]
        mem0 = torch.tensor(rng.standard_normal((B, S, D)), dtype=torch.float32)
        if mode == "bf16":
            sd, mem0 = round_sd_for_bf16(sd), bf16_representable(mem0)
        ids = torch.tensor(rng.integers(0, V, size=(B, T)), dtype=torch.long)
        labels = torch.tensor(rng.integers(0, V, size=(B, T)), dtype=torch.long)
        mem_mask = torch.ones(B, S)
        mem_mask[0, -20:] = 0
        tgt_mask = torch.ones(B, T)
        tgt_mask[1, -30:] = 0
        labels[tgt_mask == 0] = -100
        ref = _oracle(sd, H, L, mem0, ids, mem_mask, tgt_mask, labels, 0.1)
        logits, loss, dmem, grads = _run(dec, sd, mem0, ids, mem_mask, tgt_mask, labels, 0.1)
    t2 = tol if mode == "fp32" else 2e-2
    assert rel_err(logits, ref["logits"]) < t2, rel_err(logits, ref["logits"])
    assert abs(float(loss) - float(ref["loss"])) < t2 * abs(float(ref["loss"]))
    assert rel_err(dmem, ref["d_memory"]) < t2, rel_err(dmem, ref["d_memory"])
    errs = {k: rel_err(grads[k], v) for k, v in ref["grads"].items() if k in grads and float(v.norm()) > 1e-6}
    bad = {k: e for k, e in errs.items() if e > (3e-4 if mode == "fp32" else 3e-2)}
    assert not bad, bad


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("R,C,smoothing", [(257, 64000, 0.1), (33, 3001, 0.0), (8, 97, 0.2)])
def test_cross_entropy_against_torch(R, C, smoothing, dtype):
    g = torch.Generator(device=DEV).manual_seed(R + C)
    z = (3.0 * torch.randn(R, C, generator=g, device=DEV)).to(dtype)
    y = torch.randint(0, C, (R,), generator=g, device=DEV)
    y[::7] = -100
    zz = z.clone().requires_grad_()
    loss = ops.cross_entropy(zz, y, -100, smoothing)
    (loss * 1.7).backward()
    zr = z.double().requires_grad_()
    ref = torch.nn.functional.cross_entropy(zr, y, ignore_index=-100, label_smoothing=smoothing)
    (ref * 1.7).backward()
    assert abs(float(loss) - float(ref)) < 2e-5 * abs(float(ref)) + 1e-6
    t = 1e-5 if dtype == torch.float32 else 6e-3       # the gradient is stored in the logits' dtype
    assert rel_err(zz.grad, zr.grad) < t, rel_err(zz.grad, zr.grad)
    assert (zz.grad[::7] == 0).all()
    # all rows ignored: zero loss and zero gradient instead of torch's nan
    z2 = z.clone().requires_grad_()
    l2 = ops.cross_entropy(z2, torch.full_like(y, -100), -100, smoothing)
    l2.backward()
    assert float(l2) == 0.0 and float(z2.grad.abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_embedding_positions_dropout_fwd_bwd(dtype):
    B, T, D, V, p = 5, 13, 256, 300, 0.25
    g = torch.Generator(device=DEV).manual_seed(3)
    table = torch.randn(V, D, generator=g, device=DEV)
    tc = table.to(dtype)
    pos = torch.randn(32, D, generator=g, device=DEV)
    ids = torch.randint(0, V, (B, T), generator=g, device=DEV)
    ids[0, :4] = 7                                     # repeated token: the backward must accumulate
    st = torch.tensor([99, 3], dtype=torch.int64, device=DEV)
    drop = (st, p, 21)
    tm = table.clone().requires_grad_()
    out = ops.EmbedFn.apply(ids, tm, tc, pos, T, drop)
    gout = torch.randn(B * T, D, generator=g, device=DEV).to(dtype)
    (out.float() * gout.float()).sum().backward()
    m = torch.empty(B * T * D, dtype=torch.float32, device=DEV)
    _lib.call("b200_dropout_mask", _lib.dropout_arg(drop), B * T * D, m, _lib.stream_ptr())
    mask = m.view(B * T, D).double()
    tr = tc.double().requires_grad_()
    ref = (tr[ids.view(-1)] + pos[:T].double().repeat(B, 1)) * mask
    (ref * gout.double()).sum().backward()
    t = 1e-6 if dtype == torch.float32 else 6e-3
    assert rel_err(out, ref) < t
    assert rel_err(tm.grad, tr.grad) < 1e-5
