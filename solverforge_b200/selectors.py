"""Host-side candidate enumeration in the reference's pull order (the SoA rows the score kernels
consume). Mirrors:

  MoveStreamContext             solverforge-solver/src/heuristic/selector/move_selector/iter.rs:14-207
  ChangeMoveSelector            heuristic/selector/move_selector/change.rs:66-104,246-307
  NearbyListChangeMoveSelector  heuristic/selector/list_kernel/nearby_change.rs:102-232
  bounded stable top-k          heuristic/selector/nearby_list_support.rs:3-34
  MatrixDistanceMeter           crates/solverforge-cvrp/src/meters.rs:10-28

CandidateId == row index (move_selector/borrowed.rs:396-430). Pure integer arithmetic on splitmix64,
so the order is reproducible bit for bit. (The device-side generator, sfgpu_step_nearby_list_change,
covers the canonical order; seeded orders are enumerated here and scored with sfgpu_step_list_change.)
"""
from __future__ import annotations

from dataclasses import dataclass
from math import gcd

import numpy as np

from .instances import change_neighbourhood  # noqa: F401  (canonical ChangeMove order, vectorised)

_I64_MAX = np.iinfo(np.int64).max
_M64 = (1 << 64) - 1
ORIGINAL, RANDOM, SHUFFLED = 0, 1, 2


def splitmix64(v: int) -> int:
    v = (v + 0x9E3779B97F4A7C15) & _M64
    v = ((v ^ (v >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    v = ((v ^ (v >> 27)) * 0x94D049BB133111EB) & _M64
    return v ^ (v >> 31)


@dataclass(frozen=True)
class MoveStreamContext:
    """iter.rs:14-184 — seeded, stateless index maps for every selector leaf."""
    step_index: int = 0
    step_seed: int = 0
    selection_order: int = ORIGINAL

    def is_canonical(self) -> bool:
        return self.selection_order == ORIGINAL

    def mixed_seed(self, salt: int) -> int:
        return splitmix64(self.step_seed ^ ((self.step_index * 0x9E3779B97F4A7C15) & _M64) ^ salt)

    def random_index(self, length: int, salt: int) -> int:
        return 0 if length <= 1 else self.mixed_seed(salt) % length

    def random_stride(self, length: int, salt: int) -> int:
        if length <= 1:
            return 1
        stride = self.mixed_seed(salt) % (length - 1) + 1
        while gcd(stride, length) != 1:
            stride = 1 if stride == length - 1 else stride + 1
        return stride

    def selection_index(self, offset: int, length: int, salt: int) -> int:
        if self.selection_order == RANDOM:
            return self.random_index(length, salt ^ ((offset * 0xD1B54A32D192ED03) & _M64))
        if self.selection_order == SHUFFLED:
            start = self.random_index(length, salt)
            stride = self.random_stride(length, salt ^ 0xA24BAED4963EE407)
            return (start + offset * stride) % length
        return offset

    def selection_index_without_replacement(self, offset: int, length: int, salt: int) -> int:
        """iter.rs:130-147: outer source dimensions are walked once — Random and Shuffled share the strided order."""
        if self.is_canonical():
            return offset
        start = self.random_index(length, salt)
        stride = self.random_stride(length, salt ^ 0xA24BAED4963EE407)
        return (start + offset * stride) % length


def change_move_rows(values: np.ndarray, n_values: int, allows_unassigned: bool = True,
                     ctx: MoveStreamContext = MoveStreamContext(), descriptor_index: int = 0,
                     variable_index: int = 0) -> np.ndarray:
    """rows[n][2] int64 = (entity, to_value | -1). change.rs:66-104: per (permuted) entity its
    (permuted) values, then the to-None move when the entity is currently assigned."""
    if ctx.is_canonical():
        return change_neighbourhood(np.asarray(values), n_values, allows_unassigned)
    n = len(values)
    entity_salt = 0xC4A46E0000000001 ^ (descriptor_index << 32) ^ variable_index
    out = []
    for eo in range(n):
        e = ctx.selection_index(eo, n, entity_salt)
        value_salt = 0xC4A46E0000000000 ^ e ^ (descriptor_index << 32) ^ variable_index
        for vo in range(n_values):
            out.append((e, ctx.selection_index(vo, n_values, value_salt)))
        if allows_unassigned and values[e] >= 0:
            out.append((e, -1))
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def swap_move_rows(n_entities: int, ctx: MoveStreamContext = MoveStreamContext(), descriptor_index: int = 0,
                   variable_index: int = 0) -> np.ndarray:
    """rows[n][2] = (left_entity, right_entity) in the pull order of SwapMoveSelector over one entity class
    (heuristic/selector/move_selector/swap.rs:64-100,196-233): the left and the right entity lists are permuted
    independently by the stream context; every pair with left < right is a SwapMove, left-major."""
    salt = (descriptor_index << 32) ^ variable_index
    n = n_entities
    left = np.array([ctx.selection_index(o, n, 0x5A09000000000001 ^ salt) for o in range(n)], dtype=np.int64)
    right = np.array([ctx.selection_index(o, n, 0x5A09000000000002 ^ salt) for o in range(n)], dtype=np.int64)
    li, ri = np.meshgrid(left, right, indexing="ij")
    keep = li < ri
    return np.stack([li[keep], ri[keep]], axis=1).astype(np.int64)


def nearby_list_change_rows(offsets: np.ndarray, elems: np.ndarray, matrix: np.ndarray, max_nearby: int = 20,
                            ctx: MoveStreamContext = MoveStreamContext(), descriptor_index: int = 0) -> np.ndarray:
    """rows[n][4] = (src_entity, src_position, dst_entity, dst_position) uint32 in pull order.

    For every source (entity, position) in (permuted) order: scan the intra-list destinations 0..=len
    (skipping position and position+1), then every other entity's 0..=len in the (permuted) entity
    order, measure distance(src element, element at min(dst_position, len-1)) — infinite (dropped)
    when that list is empty or the cell is negative / UNREACHABLE — and keep the max_nearby smallest,
    ties in scan order.
    """
    offsets = np.asarray(offsets, dtype=np.int64)
    elems = np.asarray(elems, dtype=np.int64)
    n_owners = len(offsets) - 1
    lens = np.diff(offsets)
    entity_salt = 0xA1EA2B17C4A40001 ^ descriptor_index
    if n_owners <= 1 or ctx.is_canonical():
        entities = np.arange(n_owners)
    else:
        entities = np.array([ctx.selection_index(o, n_owners, entity_salt) for o in range(n_owners)])
    # destination slots (entity, position 0..=len) with their reference element, in entity scan order
    e_lens = lens[entities]
    slot_rank = np.repeat(np.arange(n_owners), e_lens + 1)      # index into `entities`
    slot_e = entities[slot_rank]
    slot_first = np.concatenate([[0], np.cumsum(e_lens + 1)])[:-1]
    slot_p = np.arange(len(slot_e)) - np.repeat(slot_first, e_lens + 1)
    slot_len = lens[slot_e]
    has_ref = slot_len > 0
    ref_pos = np.minimum(slot_p, np.maximum(slot_len - 1, 0))
    if len(elems):
        ref_elem = np.where(has_ref, elems[np.minimum(offsets[slot_e] + ref_pos, len(elems) - 1)], 0)
    else:
        ref_elem = np.zeros(len(slot_e), dtype=np.int64)
    out = []
    for si in range(n_owners):
        se = int(entities[si])
        slen = int(lens[se])
        if slen == 0:
            continue
        intra = np.flatnonzero(slot_rank == si)
        inter = np.flatnonzero(slot_rank != si)
        order = np.concatenate([intra, inter])            # scan order: own list first
        o_e, o_p, o_ref, o_has = slot_e[order], slot_p[order], ref_elem[order], has_ref[order]
        own = np.arange(len(order)) < len(intra)
        source_salt = 0xA1EA2B17C4A40002 ^ se ^ descriptor_index
        for po in range(slen):
            sp = ctx.selection_index(po, slen, source_salt)
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


def nearby_list_swap_rows(offsets: np.ndarray, elems: np.ndarray, matrix: np.ndarray, max_nearby: int = 20,
                          ctx: MoveStreamContext = MoveStreamContext(), descriptor_index: int = 0) -> np.ndarray:
    """rows[n][4] = (first_entity, first_position, second_entity, second_position) uint32 in the pull order of
    NearbyListSwapMoveSelector (heuristic/selector/nearby_list_swap.rs:165-213, cursor
    list_kernel/nearby_swap.rs:99-262): for every source (entity order and position order from the stream
    context) the destinations are the later positions of the same list, then every position of the entities
    later in the entity ORDER; distance(source element, destination element), infinite cells dropped; the
    max_nearby smallest, ties in scan order; sources without destinations are skipped.
    """
    offsets = np.asarray(offsets, dtype=np.int64)
    elems = np.asarray(elems, dtype=np.int64)
    n_owners = len(offsets) - 1
    lens = np.diff(offsets)
    entity_salt = 0xA1EA25A090000001 ^ descriptor_index
    if n_owners <= 1 or ctx.is_canonical():
        entities = np.arange(n_owners)
    else:
        entities = np.array([ctx.selection_index(o, n_owners, entity_salt) for o in range(n_owners)])
    out = []
    for si in range(n_owners):
        se = int(entities[si])
        slen = int(lens[se])
        if slen == 0:
            continue
        later = entities[si + 1:]
        l_e = np.repeat(later, lens[later])
        l_first = np.concatenate([[0], np.cumsum(lens[later])])[:-1]
        l_p = np.arange(len(l_e)) - np.repeat(l_first, lens[later])
        source_salt = 0xA1EA25A090000002 ^ se ^ descriptor_index
        for po in range(slen):
            sp = ctx.selection_index(po, slen, source_salt)
            own_p = np.arange(sp + 1, slen)
            d_e = np.concatenate([np.full(len(own_p), se, dtype=np.int64), l_e])
            d_p = np.concatenate([own_p, l_p])
            if len(d_e) == 0:
                continue
            x = elems[offsets[se] + sp]
            cell = matrix[x, elems[offsets[d_e] + d_p]]
            idx = np.flatnonzero((cell >= 0) & (cell != _I64_MAX))
            if len(idx) == 0:
                continue
            top = idx[np.argsort(cell[idx].astype(np.float64), kind="stable")[:max_nearby]]
            rows = np.empty((len(top), 4), dtype=np.uint32)
            rows[:, 0] = se
            rows[:, 1] = sp
            rows[:, 2] = d_e[top]
            rows[:, 3] = d_p[top]
            out.append(rows)
    return np.concatenate(out) if out else np.zeros((0, 4), dtype=np.uint32)


def sublist_change_rows(offsets: np.ndarray, min_size: int = 1, max_size: int = 3,
                        ctx: MoveStreamContext = MoveStreamContext(), descriptor_index: int = 0) -> np.ndarray:
    """rows[n][5] = (src_entity, start, end, dst_entity, dst_position) in the pull order of
    SublistChangeMoveSelector (heuristic/selector/sublist_change.rs:166-205, cursor
    list_kernel/sublist_change.rs:103-268; defaults 1..=3 from solverforge-config move_selector.rs:713-715):
    entities in stream order; per source every start, every valid size; intra-list destinations over the
    post-removal list (skipping the segment's own start) before the insertions into the other entities."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    lens = np.diff(offsets)
    ents = [ctx.selection_index(o, n, 0x5B157C4A46E00001 ^ descriptor_index) if n > 1 else o for o in range(n)]
    out = []
    for si, se in enumerate(ents):
        slen = int(lens[se])
        if slen < min_size:
            continue
        for so in range(slen):
            start = ctx.selection_index(so, slen, 0x5B157C4A46E00002 ^ se ^ descriptor_index)
            max_valid = min(max_size, slen - start)
            size_count = max_valid - min_size + 1 if max_valid >= min_size else 0
            for zo in range(size_count):
                size = min_size + ctx.selection_index(zo, size_count, 0x5B157C4A46E00003 ^ se ^ start)
                end, post = start + size, slen - size
                for po in range(post + 1):
                    dp = ctx.selection_index(po, post + 1, 0x5B157C4A46E00004 ^ se ^ start)
                    if dp != start:
                        out.append((se, start, end, se, dp))
                for di, de in enumerate(ents):
                    if di == si:
                        continue
                    dlen = int(lens[de])
                    for po in range(dlen + 1):
                        out.append((se, start, end, de, ctx.selection_index(po, dlen + 1, 0x5B157C4A46E00005 ^ se ^ de ^ start)))
    return np.array(out, dtype=np.uint32).reshape(-1, 5)


def sublist_swap_rows(offsets: np.ndarray, min_size: int = 1, max_size: int = 3,
                      ctx: MoveStreamContext = MoveStreamContext(), descriptor_index: int = 0) -> np.ndarray:
    """rows[n][6] = (first_entity, start1, end1, second_entity, start2, end2) in the pull order of
    SublistSwapMoveSelector (heuristic/selector/sublist_swap.rs, cursor list_kernel/sublist_swap.rs:28-318): first
    segments in entity / start / size stream order; second segments from the same entity onwards; inside one list
    only segments that start at or after the end of the first one."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    lens = np.diff(offsets)
    ents = [ctx.selection_index(o, n, 0x5B1575A090000001 ^ descriptor_index) if n > 1 else o for o in range(n)]

    def segments(e):
        ln = int(lens[e])
        out = []
        if ln < min_size:
            return out
        for so in range(ln):
            start = ctx.selection_index(so, ln, 0x5B1575A090000002 ^ e ^ descriptor_index)
            max_valid = min(max_size, ln - start)
            if max_valid < min_size:
                continue
            count = max_valid - min_size + 1
            for zo in range(count):
                out.append((start, start + min_size + ctx.selection_index(zo, count, 0x5B1575A090000003 ^ e ^ start)))
        return out

    segs = [segments(e) for e in ents]
    out = []
    for fi in range(n):
        for (s1, t1) in segs[fi]:
            for si in range(fi, n):
                for (s2, t2) in segs[si]:
                    if fi == si and (s2 < t1 or (s1, t1) == (s2, t2)):
                        continue
                    out.append((ents[fi], s1, t1, ents[si], s2, t2))
    return np.array(out, dtype=np.uint32).reshape(-1, 6)


def _binomial(n: int, k: int) -> int:
    if k > n:
        return 0
    k = min(k, n - k)
    r = 1
    for i in range(k):
        r = r * (n - i) // (i + 1)
    return r


def k_opt_cut_combination_at(k: int, length: int, min_seg: int, rank: int):
    """heuristic/selector/k_opt/iterators.rs:110-160 — the cut positions of the combination with this rank."""
    choice_count = length - (k + 1) * min_seg + k
    cuts, start = [], 0
    for position in range(k):
        remaining, maximum = k - position - 1, choice_count - (k - position)
        for cand in range(start, maximum + 1):
            suffix = _binomial(choice_count - cand - 1, remaining)
            if rank < suffix:
                cuts.append(cand + min_seg + position * (min_seg - 1))
                start = cand + 1
                break
            rank -= suffix
    return cuts


def k_opt_pattern_count(k: int) -> int:
    """Non-identity reconnection patterns of k-opt (k_opt_reconnection.rs:218-262): (k-1)! * 2^(k-1) - 1."""
    f = 1
    for i in range(2, k):
        f *= i
    return f * (1 << (k - 1)) - 1


def k_opt_rows(offsets: np.ndarray, k: int = 3, min_seg: int = 1, ctx: MoveStreamContext = MoveStreamContext(),
               descriptor_index: int = 0) -> np.ndarray:
    """rows[n][k + 2] = (entity, cut_0 .. cut_{k-1}, pattern index) in the pull order of KOptMoveSelector
    (heuristic/selector/list_kernel/k_opt/full.rs:34-98): per entity the (cut combination rank) x (pattern)
    product pulled through selection_index; entities in the without-replacement stream order (full.rs:48-51).
    The device walks the same cursor (SFGPU_FAM_K_OPT) and scores these rows (sfgpu_score_k_opt)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    lens = np.diff(offsets)
    n_pat = k_opt_pattern_count(k)
    out = []
    n_owners = len(lens)
    for eo in range(n_owners):
        e = ctx.selection_index_without_replacement(eo, n_owners, 0x4B0F7E1171000001 ^ descriptor_index)
        ln = int(lens[e])
        combos = _binomial(ln - (k + 1) * min_seg + k, k) if ln >= (k + 1) * min_seg else 0
        move_count = combos * n_pat
        for off in range(move_count):
            sel = ctx.selection_index(off, move_count, 0x4B0F7E1171000002 ^ descriptor_index ^ e)
            out.append([e] + k_opt_cut_combination_at(k, ln, min_seg, sel // n_pat) + [sel % n_pat])
    return np.array(out, dtype=np.uint32).reshape(-1, k + 2)


def list_reverse_rows(offsets: np.ndarray, ctx: MoveStreamContext = MoveStreamContext(),
                      descriptor_index: int = 0) -> np.ndarray:
    """rows[n][4] = (entity, start, end, 0) uint32 in the pull order of ListReverseMoveSelector
    (heuristic/selector/list_reverse.rs:139-172, cursor list_kernel/reverse.rs:66-108): entities in stream
    order; per entity with len >= 2 every start and every end in start+2..=len, both in stream order."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n_owners = len(offsets) - 1
    lens = np.diff(offsets)
    out = []
    for o in range(n_owners):
        e = ctx.selection_index(o, n_owners, 0x11572A0700000001 ^ descriptor_index)
        ln = int(lens[e])
        if ln < 2:
            continue
        for so in range(ln):
            start = ctx.selection_index(so, ln, 0x11572A0700000002 ^ e ^ descriptor_index)
            end_count = max(ln - (start + 1), 0)
            for eo in range(end_count):
                end = start + 2 + ctx.selection_index(eo, end_count, 0x11572A0700000003 ^ e ^ start)
                out.append((e, start, end, 0))
    return np.array(out, dtype=np.uint32).reshape(-1, 4)
