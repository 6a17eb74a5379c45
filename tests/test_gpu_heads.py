"""GPU: AnswerHead against the reference's own layer stack (plain torch nn.Sequential, the exact
module structure of vqa_model.py:451-465) — logits and every gradient, fp32 1e-4 / bf16 1e-2, including class counts
that are not a multiple of 8 (the answer vocabulary size comes from the dataset)."""
import copy

import pytest
import torch
from torch import nn

from conftest import rel_err

pytestmark = pytest.mark.gpu

import vqa_model_builder_b200 as pkg  # noqa: E402
from vqa_model_builder_b200 import heads  # noqa: E402

DEV = "cuda"


def reference_head(cfg, input_dim):
    layers, prev = [], input_dim
    for hd in cfg.hidden_dims:
        layers.extend([nn.Linear(prev, hd), nn.ReLU(), nn.Dropout(cfg.dropout)])
        prev = hd
    layers.append(nn.Linear(prev, cfg.num_answers))
    return nn.Sequential(*layers)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("num_answers,hidden,batch", [(3000, [512, 256], 32), (3001, [512, 256], 5), (10, [64], 1),
                                                      (77, [], 9)])
def test_answer_head_parity(mode, num_answers, hidden, batch):
    torch.manual_seed(num_answers + batch)
    cfg = heads.AnswerHeadConfig(num_answers=num_answers, hidden_dims=hidden, dropout=0.0)
    ours = heads.AnswerHead(cfg, 768).to(DEV).train()
    ref = reference_head(cfg, 768).double()
    x = torch.randn(batch, 768)
    if mode == "bf16":      # both sides start from bf16-representable values
        with torch.no_grad():
            for p in ours.parameters():
                p.copy_(p.to(torch.bfloat16).float())
        x = x.to(torch.bfloat16).float()
    ref.load_state_dict({"" + k.replace("classifier.", ""): v.detach().cpu().double()
                         for k, v in ours.state_dict().items()})
    assert [k for k in ours.state_dict()] == ["classifier." + k for k in ref.state_dict()]
    pkg.set_compute_dtype(mode)
    try:
        xg = x.to(DEV).requires_grad_()
        logits = ours(xg)
        assert logits.shape == (batch, num_answers) and logits.dtype == torch.float32
        # upstream gradient fixed on both sides (a softmax loss would feed the logits' own bf16 rounding back into
        # the gradient and measure the loss, not the head); bf16-representable like every other input
        gout = torch.randn(batch, num_answers)
        if mode == "bf16":
            gout = gout.to(torch.bfloat16).float()
        (logits * gout.to(DEV)).sum().backward()
    finally:
        pkg.set_compute_dtype("auto")
    xr = x.double().requires_grad_()
    if mode == "fp32":
        lr = ref(xr)
    else:
        # bf16 mode stores every hidden activation in bf16.  ReLU is not smooth: a pre-activation within that
        # rounding error of zero switches a whole unit on or off, so the oracle must see the same stored values
        # (straight-through rounding) — otherwise ~0.3 % of the units flip and the gradients differ by several
        # per cent at batch 32 although both sides are "right".
        h = xr
        for m in ref:
            h = m(h)
            if isinstance(m, nn.ReLU):
                h = h + (h.detach().to(torch.bfloat16).double() - h.detach())
        lr = h
    (lr * gout.double()).sum().backward()
    tol = 1e-4 if mode == "fp32" else 1e-2
    assert rel_err(logits, lr) < tol, rel_err(logits, lr)
    errs = {n: rel_err(p.grad, q.grad) for (n, p), q in zip(ours.named_parameters(), ref.parameters())}
    errs["input"] = rel_err(xg.grad, xr.grad)
    assert max(errs.values()) < tol, errs


def test_answer_head_cross_entropy_step_fp32():
    torch.manual_seed(3)
    cfg = heads.AnswerHeadConfig(num_answers=3001, hidden_dims=[512, 256], dropout=0.0)
    ours = heads.AnswerHead(cfg, 768).to(DEV).train()
    ref = reference_head(cfg, 768).double()
    ref.load_state_dict({k.replace("classifier.", ""): v.detach().cpu().double() for k, v in ours.state_dict().items()})
    x = torch.randn(16, 768)
    tgt = torch.randint(0, 3001, (16,))
    xg = x.to(DEV).requires_grad_()
    loss = torch.nn.functional.cross_entropy(ours(xg), tgt.to(DEV))
    loss.backward()
    xr = x.double().requires_grad_()
    lref = torch.nn.functional.cross_entropy(ref(xr), tgt)
    lref.backward()
    assert abs(float(loss) - float(lref)) < 1e-5
    assert rel_err(xg.grad, xr.grad) < 1e-4
    for p, q in zip(ours.parameters(), ref.parameters()):
        assert rel_err(p.grad, q.grad) < 1e-4


def test_answer_head_train_mode_dropout_and_deepcopy():
    torch.manual_seed(0)
    cfg = heads.AnswerHeadConfig(num_answers=100, hidden_dims=[256, 128], dropout=0.5)
    head = heads.AnswerHead(cfg, 768).to(DEV)
    x = torch.randn(64, 768, device=DEV)
    head.eval()
    a, b = head(x), head(x)
    assert torch.equal(a, b)                      # eval: deterministic, no dropout
    head.train()
    c, d = head(x), head(x)
    assert not torch.equal(c, d)                  # train: fresh masks every forward
    c.sum().backward()
    assert all(torch.isfinite(p.grad).all() for p in head.parameters())
    twin = copy.deepcopy(head).eval()
    assert torch.equal(twin(x), a)
