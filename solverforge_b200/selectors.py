"""Host-side candidate enumeration in the reference's pull order (the SoA rows the score kernels
consume). Mirrors, for SelectionOrder::Original:

  ChangeMoveSelector            solverforge-solver/src/heuristic/selector/move_selector/change.rs:66-104,246-307
  NearbyListChangeMoveSelector  heuristic/selector/list_kernel/nearby_change.rs:102-232
  bounded stable top-k          heuristic/selector/nearby_list_support.rs:3-34
  MatrixDistanceMeter           crates/solverforge-cvrp/src/meters.rs:10-28

CandidateId == row index (move_selector/borrowed.rs:396-430).
"""
from __future__ import annotations

import numpy as np

from .instances import change_neighbourhood  # noqa: F401  (ChangeMove order)

_I64_MAX = np.iinfo(np.int64).max


def nearby_list_change_rows(offsets: np.ndarray, elems: np.ndarray, matrix: np.ndarray, max_nearby: int = 20
                            ) -> np.ndarray:
    """rows[n][4] = (src_entity, src_position, dst_entity, dst_position) uint32 in canonical order.

    For every source (entity, position) in order: scan the intra-list destinations 0..=len (skipping
    position and position+1), then every other entity's 0..=len, measure
    distance(src element, element at min(dst_position, len-1)) — infinite (dropped) when that list is
    empty or the cell is negative / UNREACHABLE — and keep the max_nearby smallest, ties in scan order.
    """
    offsets = np.asarray(offsets, dtype=np.int64)
    elems = np.asarray(elems, dtype=np.int64)
    n_owners = len(offsets) - 1
    lens = np.diff(offsets)
    # every destination slot (entity, position 0..=len) with its reference element, in entity order
    slot_e = np.repeat(np.arange(n_owners), lens + 1)
    slot_first = np.concatenate([[0], np.cumsum(lens + 1)])[:-1]
    slot_p = np.arange(len(slot_e)) - np.repeat(slot_first, lens + 1)
    slot_len = lens[slot_e]
    has_ref = slot_len > 0
    ref_pos = np.minimum(slot_p, np.maximum(slot_len - 1, 0))
    ref_elem = np.where(has_ref, elems[np.minimum(offsets[slot_e] + ref_pos, max(len(elems) - 1, 0))]
                        if len(elems) else 0, 0)
    out = []
    for se in range(n_owners):
        slen = int(lens[se])
        if slen == 0:
            continue
        intra = np.flatnonzero(slot_e == se)
        inter = np.flatnonzero(slot_e != se)
        order = np.concatenate([intra, inter])            # scan order: own list first
        o_e, o_p, o_ref, o_has = slot_e[order], slot_p[order], ref_elem[order], has_ref[order]
        own = np.arange(len(order)) < len(intra)
        for sp in range(slen):
            x = elems[offsets[se] + sp]
            cell = matrix[x, o_ref]
            finite = o_has & (cell >= 0) & (cell != _I64_MAX)
            keep = finite & ~(own & ((o_p == sp) | (o_p == sp + 1)))
            idx = np.flatnonzero(keep)
            if len(idx) == 0:
                continue
            dist = cell[idx].astype(np.float64)
            top = idx[np.argsort(dist, kind="stable")[:max_nearby]]
            rows = np.empty((len(top), 4), dtype=np.uint32)
            rows[:, 0] = se
            rows[:, 1] = sp
            rows[:, 2] = o_e[top]
            rows[:, 3] = o_p[top]
            out.append(rows)
    return np.concatenate(out) if out else np.zeros((0, 4), dtype=np.uint32)
