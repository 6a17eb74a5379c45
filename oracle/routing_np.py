"""ORACLE — test infrastructure only.  Integer side of the MOE path in numpy: the canonical permutation map,
expert offsets, padded layout and capacity selection implied by the reference's dispatch
(SparseMOELayer.forward, moe_layer.py:317-346: per expert `nonzero()` of the token mask, i.e. rows ordered by
(expert id ascending, token index ascending) == stable sort of the flattened (token, slot) list by expert id).
Bit-exact targets for the CUDA plan kernels."""
from __future__ import annotations

import numpy as np

GROUP_TILE = 128


def routing_plan(idx: np.ndarray, num_experts: int, rmax: int | None = None):
    """idx: int array [N, K] (entries outside [0, E) are dropped, as ablation masking sets -1).
    Returns dict(counts[E], cmp_off[E+1], pad_off[E+1], cmp_pos[N*K], dest_row[N*K], row_src[Rmax], tile_group)."""
    flat = np.asarray(idx).reshape(-1).astype(np.int64)
    nk = flat.size
    valid = (flat >= 0) & (flat < num_experts)
    counts = np.bincount(flat[valid], minlength=num_experts).astype(np.int32)
    cmp_off = np.zeros(num_experts + 1, np.int32)
    cmp_off[1:] = np.cumsum(counts)
    padded = (counts + GROUP_TILE - 1) // GROUP_TILE * GROUP_TILE
    pad_off = np.zeros(num_experts + 1, np.int32)
    pad_off[1:] = np.cumsum(padded)
    if rmax is None:
        rmax = max_rows(nk, num_experts)
    order = np.argsort(np.where(valid, flat, num_experts), kind="stable")     # stable: token order inside an expert
    cmp_pos = np.full(nk, -1, np.int32)
    dest_row = np.full(nk, -1, np.int32)
    n_valid = int(valid.sum())
    ranks = np.arange(n_valid, dtype=np.int32)
    src = order[:n_valid]
    cmp_pos[src] = ranks
    e_of = flat[src]
    dest_row[src] = (ranks - cmp_off[e_of] + pad_off[e_of]).astype(np.int32)
    row_src = np.full(rmax, -1, np.int32)
    row_src[dest_row[src]] = src.astype(np.int32)
    tile_group = np.full(rmax // GROUP_TILE, -1, np.int32)
    for e in range(num_experts):
        tile_group[pad_off[e] // GROUP_TILE:pad_off[e + 1] // GROUP_TILE] = e
    return dict(counts=counts, cmp_off=cmp_off, pad_off=pad_off, cmp_pos=cmp_pos, dest_row=dest_row,
                row_src=row_src, tile_group=tile_group)


def max_rows(nk: int, num_experts: int) -> int:
    r = (nk + num_experts * (GROUP_TILE - 1) + GROUP_TILE - 1) // GROUP_TILE * GROUP_TILE
    return max(r, GROUP_TILE)


def capacity_keep(idx: np.ndarray, w: np.ndarray, num_experts: int, capacity: int) -> np.ndarray:
    """keep[N*K] (uint8): SparseMOELayer keeps, per over-subscribed expert, the `capacity` (token, slot) pairs with
    the largest combine weight (moe_layer.py:329-337); exact-weight ties resolve to the lower token index here."""
    flat = np.asarray(idx).reshape(-1)
    wf = np.asarray(w, dtype=np.float32).reshape(-1)
    keep = np.ones(flat.size, np.uint8)
    for e in range(num_experts):
        members = np.nonzero(flat == e)[0]
        if members.size > capacity:
            order = np.lexsort((members, -wf[members]))      # by weight desc, then position asc
            keep[members[order[capacity:]]] = 0
    return keep


def topk_with_ties(probs: np.ndarray, k: int, tol: float = 1e-6):
    """Top-k expert ids per row (descending) plus a mask of rows whose selection is ambiguous within `tol`
    (the k-th and (k+1)-th probabilities, or two selected ones, closer than tol) — those rows are exempt from the
    bit-exact index comparison (north star)."""
    p = np.asarray(probs, dtype=np.float64)
    order = np.argsort(-p, axis=-1, kind="stable")
    top = order[..., :k]
    sorted_p = np.take_along_axis(p, order, axis=-1)
    gaps = sorted_p[..., :k] - sorted_p[..., 1:k + 1] if p.shape[-1] > k else sorted_p[..., :k - 1] - sorted_p[..., 1:k]
    ambiguous = (np.abs(gaps) < tol).any(axis=-1)
    return top, ambiguous
