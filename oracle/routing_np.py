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


def ep_layout(tab: np.ndarray, me: int, rcap: int, nk_cap: int):
    """Expert-parallel layout derived from the table tab[s][e] = (token, slot) pairs rank s routes to global expert e
    (experts sharded contiguous-block, expert e on rank e // (E / W)).  On the owner, expert segments lie in ascending
    expert order, each padded to 128 rows, rows inside a segment ordered by (source rank, canonical position at the
    source) — the canonical (expert, token) order of the unsharded layer on the concatenated batch.
    Returns what rank `me` needs: send_base[E], pad_off2[2*El+1] (offsets, then routed rows per local expert),
    tile_group2[rcap/128], row_home[rcap] (home_rank * nk_cap + compact position at home, -1 padding)."""
    tab = np.asarray(tab, dtype=np.int64)
    W, E = tab.shape
    El = E // W
    tot = tab.sum(axis=0)
    padded = (tot + GROUP_TILE - 1) // GROUP_TILE * GROUP_TILE
    pad = np.zeros(E, np.int64)                      # padded offset of expert e inside its owner's buffer
    for e in range(E):
        r = e // El
        pad[e] = padded[r * El:e].sum()
    send_base = (pad + tab[:me].sum(axis=0)).astype(np.int32)
    pad_off2 = np.zeros(2 * El + 1, np.int32)
    for le in range(El):
        pad_off2[le] = pad[me * El + le]
        pad_off2[El + 1 + le] = tot[me * El + le]
    pad_off2[El] = pad[me * El + El - 1] + padded[me * El + El - 1]
    tile_group2 = np.full(rcap // GROUP_TILE, -1, np.int32)
    row_home = np.full(rcap, -1, np.int32)
    cmp_off = np.concatenate([np.zeros((W, 1), np.int64), np.cumsum(tab, axis=1)], axis=1)   # per source rank
    for le in range(El):
        e = me * El + le
        tile_group2[pad_off2[le] // GROUP_TILE:pad_off2[le + 1] // GROUP_TILE] = le
        row = pad_off2[le]
        for s in range(W):
            n = int(tab[s, e])
            row_home[row:row + n] = s * nk_cap + cmp_off[s, e] + np.arange(n)
            row += n
    return dict(send_base=send_base, pad_off2=pad_off2, tile_group2=tile_group2, row_home=row_home)
