"""ORACLE — test infrastructure only.  Fingerprints of large tensors: Frobenius norm + seeded probes.

The reference's outputs and gradients at the benchmark dimensions (D=768, F=2048) are too large to commit in full
(SURVEY 8(c) golden-vector policy), so oracle/make_golden.py stores, per tensor, its norm and 256 values at seeded
positions; tests re-derive the same positions and compare
    ||probe(got) - probe(ref)|| / ||probe(ref)||   (a sampled estimate of the full relative error)
and the norms."""
from __future__ import annotations

import hashlib
from typing import Dict

import numpy as np
import torch

N_PROBES = 256
SEED0 = 1000


def probe_positions(numel: int, seed: int, n: int = N_PROBES) -> np.ndarray:
    if numel <= n:
        return np.arange(numel)
    return np.random.default_rng(seed).integers(0, numel, size=n)


def probe(t: torch.Tensor, seed: int, n: int = N_PROBES) -> np.ndarray:
    flat = t.detach().float().cpu().reshape(-1).numpy()
    return flat[probe_positions(flat.size, seed, n)].copy()


def fingerprint(tensors: Dict[str, torch.Tensor], seed0: int = SEED0) -> Dict[str, np.ndarray]:
    """name -> norm / probes; probe seeds follow the sorted key order."""
    out = {}
    for i, k in enumerate(sorted(tensors)):
        t = tensors[k]
        out[f"norm/{k}"] = np.array(float(t.detach().double().norm()))
        out[f"probe/{k}"] = probe(t, seed0 + i)
    return out


def compare(tensors: Dict[str, torch.Tensor], fp: Dict[str, np.ndarray], seed0: int = SEED0) -> Dict[str, float]:
    """Relative probe error per tensor (norm mismatch folded in as max).  `fp` holds 'norm/<k>' and 'probe/<k>'."""
    errs = {}
    names = sorted(k[len("probe/"):] for k in fp if k.startswith("probe/"))
    assert names == sorted(tensors), (set(names) ^ set(tensors))
    for i, k in enumerate(names):
        ref = np.asarray(fp[f"probe/{k}"], dtype=np.float64)
        got = probe(tensors[k], seed0 + i).astype(np.float64)
        den = np.linalg.norm(ref)
        if den == 0.0:
            errs[k] = float(np.linalg.norm(got))
            continue
        e = float(np.linalg.norm(got - ref) / den)
        rn = float(fp[f"norm/{k}"])
        gn = float(tensors[k].detach().double().norm())
        errs[k] = max(e, abs(gn - rn) / rn if rn > 0 else 0.0)
    return errs


def sha_int(t: torch.Tensor) -> np.ndarray:
    h = hashlib.sha256(t.detach().cpu().to(torch.int64).contiguous().numpy().tobytes()).digest()
    return np.frombuffer(h, dtype=np.uint8).copy()
