import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name: str) -> dict:
    """npz -> nested dict of torch tensors ('sd/..', 'grads/..' prefixes become sub-dicts)."""
    raw = np.load(GOLDEN / f"{name}.npz")
    out: dict = {}
    for k in raw.files:
        a = np.array(raw[k])
        v = torch.from_numpy(a) if a.dtype.kind in "fiub" else a      # string arrays (key lists) stay numpy
        if "/" in k:
            head, tail = k.split("/", 1)
            out.setdefault(head, {})[tail] = v
        else:
            out[k] = v
    return out


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a - b||_F / ||b||_F in fp64 (the relative-error measure of the parity targets)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)


def bf16_representable(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def round_sd_for_bf16(sd: dict) -> dict:
    """What the bf16 compute path sees: matrices are rounded to bf16, vectors (biases, LayerNorm affine) and
    buffers stay fp32.  Both implementations then start from identical values."""
    return {k: (bf16_representable(v) if v.dtype.is_floating_point and v.dim() >= 2 else v.clone())
            for k, v in sd.items()}


def leafs(sd: dict) -> dict:
    return {k: v.clone().requires_grad_(v.dtype.is_floating_point and v.dim() > 0) for k, v in sd.items()}


def seeded_normal(seed: int, *shapes):
    """The input generator of oracle/make_golden.py's fingerprint fixtures (numpy PCG64, stable across versions)."""
    rng = np.random.default_rng(seed)
    return [torch.tensor(rng.standard_normal(s), dtype=torch.float32) for s in shapes]


def fp_flat(g: dict) -> dict:
    """load_golden() nests 'norm/<k>' and 'probe/<k>' one level; oracle.fingerprint.compare wants them flat."""
    out = {}
    for head in ("norm", "probe"):
        for k, v in g[head].items():
            out[f"{head}/{k}"] = v.numpy() if isinstance(v, torch.Tensor) else v
    return out
