"""Synthetic instances of the BASELINE.json configs: integer-only generators driven
by a splitmix64 stream, so the oracle, the CUDA path and the benchmark all see identical data."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64_stream(seed: int, n: int) -> np.ndarray:
    """n outputs of splitmix64 seeded with `seed` (state += golden; mix)."""
    with np.errstate(over="ignore"):
        state = np.uint64(seed) + _GOLDEN * np.arange(1, n + 1, dtype=np.uint64)
        z = state
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


@dataclass
class GraphColoringInstance:
    n: int
    k: int
    row_ptr: np.ndarray  # uint32 [n+1], symmetric adjacency
    col: np.ndarray      # uint32 [nnz]
    color: np.ndarray    # int32 [n], -1 = unassigned


def graph_coloring(n: int = 10_000, m: int = 50_000, k: int = 8, seed_edges: int = 42, seed_colors: int = 43,
                   unassigned_permille: int = 10) -> GraphColoringInstance:
    """C2: m undirected edges drawn uniformly without self-loops/duplicates, stored symmetric."""
    need = m
    draws = splitmix64_stream(seed_edges, 4 * m + 64)
    a = (draws[0::2] % np.uint64(n)).astype(np.int64)
    b = (draws[1::2] % np.uint64(n)).astype(np.int64)
    keep = a != b
    lo, hi = np.minimum(a, b)[keep], np.maximum(a, b)[keep]
    key = lo * n + hi
    _, first = np.unique(key, return_index=True)
    first.sort()
    first = first[:need]
    if len(first) < need:
        raise ValueError("graph too dense for the draw budget")
    lo, hi = lo[first], hi[first]
    src = np.concatenate([lo, hi])
    dst = np.concatenate([hi, lo])
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    row_ptr = np.zeros(n + 1, dtype=np.uint32)
    np.add.at(row_ptr, src + 1, 1)
    row_ptr = np.cumsum(row_ptr, dtype=np.uint64).astype(np.uint32)
    cs = splitmix64_stream(seed_colors, 2 * n)
    color = (cs[:n] % np.uint64(k)).astype(np.int32)
    color[(cs[n:] % np.uint64(1000)) < np.uint64(unassigned_permille)] = -1
    return GraphColoringInstance(n, k, row_ptr, dst.astype(np.uint32), color)


def graph_coloring_colors(inst: GraphColoringInstance, seed: int, unassigned_permille: int = 10) -> np.ndarray:
    """Another seeded colouring of the same graph (replica-specific start)."""
    cs = splitmix64_stream(seed, 2 * inst.n)
    color = (cs[:inst.n] % np.uint64(inst.k)).astype(np.int32)
    color[(cs[inst.n:] % np.uint64(1000)) < np.uint64(unassigned_permille)] = -1
    return color


@dataclass
class NQueensInstance:
    n: int
    row: np.ndarray  # int32 [n]


def nqueens(n: int = 64, seed: int | None = None) -> NQueensInstance:
    """C1: queen i sits in column i; start row (i*7+3) % n, or uniform from `seed`."""
    if seed is None:
        row = ((np.arange(n) * 7 + 3) % n).astype(np.int32)
    else:
        row = (splitmix64_stream(seed, n) % np.uint64(n)).astype(np.int32)
    return NQueensInstance(n, row)


@dataclass
class ClusterInstance:
    n: int
    n_teams: int
    team: np.ndarray  # int32 [n], -1 = unassigned
    joins: tuple      # ((arity, soft weight), ...) keyed self-joins on the team, arity 2..5


def cluster(n: int = 60, n_teams: int = 5, seed: int = 17, unassigned_permille: int = 80,
            joins=((2, 1), (3, 10), (4, 100), (5, 1000))) -> ClusterInstance:
    """Task clustering (the shape of the reference's tri / quad / penta known-answer tests,
    constraint/tests/{tri,quad,penta}_incr.rs, as a planning model): tasks keyed by their team."""
    r = splitmix64_stream(seed, 2 * n)
    team = (r[:n] % np.uint64(n_teams)).astype(np.int32)
    team[(r[n:] % np.uint64(1000)) < unassigned_permille] = -1
    return ClusterInstance(n, n_teams, team, tuple(joins))


@dataclass
class CvrpInstance:
    dim: int              # locations incl. depot 0
    n_routes: int
    capacity: int
    depot: int
    demands: np.ndarray   # int32 [dim]
    matrix: np.ndarray    # int64 [dim, dim]
    offsets: np.ndarray   # uint32 [n_routes+1]
    elems: np.ndarray     # uint32 [dim-1]


def cvrp(n_customers: int = 1000, n_routes: int = 80, seed: int = 7) -> CvrpInstance:
    """C3: coordinates uniform in [0,1000)^2, rounded Euclidean int64 matrix, demands 1..=20,
    capacity = ceil(sum/n_routes * 1.15), round-robin initial routes in id order."""
    dim = n_customers + 1
    s = splitmix64_stream(seed, 3 * dim)
    x = (s[0:dim] % np.uint64(1000)).astype(np.int64)
    y = (s[dim:2 * dim] % np.uint64(1000)).astype(np.int64)
    demands = (s[2 * dim:3 * dim] % np.uint64(20)).astype(np.int32) + 1
    demands[0] = 0
    dx = x[:, None] - x[None, :]
    dy = y[:, None] - y[None, :]
    matrix = np.rint(np.sqrt((dx * dx + dy * dy).astype(np.float64))).astype(np.int64)
    total = int(demands.sum())
    capacity = -(-(total * 115) // (n_routes * 100))
    routes = [[] for _ in range(n_routes)]
    for c in range(1, dim):
        routes[(c - 1) % n_routes].append(c)
    offsets = np.zeros(n_routes + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum([len(r) for r in routes])
    elems = np.array([c for r in routes for c in r], dtype=np.uint32)
    return CvrpInstance(dim, n_routes, capacity, 0, demands, matrix, offsets, elems)


def perturb_routes(inst: CvrpInstance, seed: int, n_moves: int = 64):
    """Replica-specific start: applies n_moves seeded relocations to the base routes."""
    routes = [list(inst.elems[inst.offsets[r]:inst.offsets[r + 1]]) for r in range(inst.n_routes)]
    s = splitmix64_stream(seed, 4 * n_moves)
    for i in range(n_moves):
        se = int(s[4 * i] % np.uint64(inst.n_routes))
        if not routes[se]:
            continue
        sp = int(s[4 * i + 1] % np.uint64(len(routes[se])))
        de = int(s[4 * i + 2] % np.uint64(inst.n_routes))
        v = routes[se].pop(sp)
        dp = int(s[4 * i + 3] % np.uint64(len(routes[de]) + 1))
        routes[de].insert(dp, v)
    offsets = np.zeros(inst.n_routes + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum([len(r) for r in routes])
    elems = np.array([c for r in routes for c in r], dtype=np.uint32)
    return offsets, elems


@dataclass
class JobShopInstance:
    n_ops: int
    n_machines: int
    job: np.ndarray          # uint32 [n_ops]
    step: np.ndarray         # uint32 [n_ops]
    machine_idx: np.ndarray  # int32 [n_ops]
    seq_offsets: np.ndarray  # uint32 [n_machines+1]
    seq_elems: np.ndarray    # uint32


def job_shop(n_jobs: int = 200, n_steps: int = 20, n_machines: int = 20, seed: int = 11,
             unassigned_permille: int = 0) -> JobShopInstance:
    """C4: operation id = job*n_steps + step; machine uniform (seed); every assigned operation sits
    in its machine's sequence in id order."""
    n = n_jobs * n_steps
    ids = np.arange(n, dtype=np.uint32)
    s = splitmix64_stream(seed, 2 * n)
    mach = (s[:n] % np.uint64(n_machines)).astype(np.int32)
    if unassigned_permille:
        mach[(s[n:] % np.uint64(1000)) < np.uint64(unassigned_permille)] = -1
    seqs = [ids[mach == m] for m in range(n_machines)]
    offsets = np.zeros(n_machines + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum([len(q) for q in seqs])
    elems = np.concatenate(seqs).astype(np.uint32) if n else np.zeros(0, np.uint32)
    return JobShopInstance(n, n_machines, (ids // n_steps).astype(np.uint32), (ids % n_steps).astype(np.uint32), mach,
                           offsets, elems)


def job_shop_machines(inst: JobShopInstance, seed: int, unassigned_permille: int = 0) -> np.ndarray:
    """Another seeded machine assignment of the same operations (replica-specific start; the machine sequences
    of the list variable stay those of the base instance)."""
    s = splitmix64_stream(seed, 2 * inst.n_ops)
    mach = (s[:inst.n_ops] % np.uint64(inst.n_machines)).astype(np.int32)
    if unassigned_permille:
        mach[(s[inst.n_ops:] % np.uint64(1000)) < np.uint64(unassigned_permille)] = -1
    return mach


def change_neighbourhood(values: np.ndarray, n_values: int, allows_unassigned: bool = True) -> np.ndarray:
    """Canonical ChangeMove order (move_selector/change.rs:66-104): per entity every value, then the
    to-None move when the entity is assigned. Returns rows[n][2] int64 (entity, to_value)."""
    n = len(values)
    ent = np.repeat(np.arange(n, dtype=np.int64), n_values + 1)
    val = np.tile(np.concatenate([np.arange(n_values, dtype=np.int64), [-1]]), n)
    keep = np.ones(len(ent), dtype=bool)
    if allows_unassigned:
        keep[(val == -1) & (np.repeat(values, n_values + 1) < 0)] = False
    else:
        keep[val == -1] = False
    return np.stack([ent[keep], val[keep]], axis=1)


@dataclass
class ShiftInstance:
    n_shifts: int
    n_nurses: int
    day: np.ndarray        # int64
    slot: np.ndarray       # uint32
    required: np.ndarray   # uint8
    hours: np.ndarray      # int64 (0 hours exercises load_balance's zero-metric skip)
    nurse_idx: np.ndarray  # int32, -1 = unassigned
    target: int = 4


def shift_scheduling(n_days: int = 14, slots_per_day: int = 3, n_nurses: int = 6, seed: int = 21,
                     unassigned_permille: int = 150) -> ShiftInstance:
    """examples/minimal-shift-scheduling scaled up: shift id = day*slots + slot."""
    n = n_days * slots_per_day
    ids = np.arange(n)
    s = splitmix64_stream(seed, 4 * n)
    nurse = (s[:n] % np.uint64(n_nurses)).astype(np.int32)
    nurse[(s[n:2 * n] % np.uint64(1000)) < np.uint64(unassigned_permille)] = -1
    required = ((s[2 * n:3 * n] % np.uint64(4)) != 0).astype(np.uint8)
    hours = (s[3 * n:4 * n] % np.uint64(4)).astype(np.int64) * 4   # 0, 4, 8, 12
    return ShiftInstance(n, n_nurses, (ids // slots_per_day).astype(np.int64), (ids % slots_per_day).astype(np.uint32),
                         required, hours, nurse)


@dataclass
class RosterInstance:
    n_shifts: int
    n_nurses: int
    n_days: int
    limit: int               # daily hour limit per nurse
    required: np.ndarray     # uint8 [n_shifts]
    span_ptr: np.ndarray     # uint32 [n_shifts + 1]: projected rows of shift i are span_*[span_ptr[i]:span_ptr[i+1]]
    span_day: np.ndarray     # int64 [rows]
    span_hours: np.ndarray   # int64 [rows]
    nurse_idx: np.ndarray    # int32 [n_shifts], -1 = unassigned


def roster(n_shifts: int = 120, n_nurses: int = 7, n_days: int = 10, seed: int = 33, limit: int = 10,
           unassigned_permille: int = 120, max_spans: int = 3) -> RosterInstance:
    """Shifts that span 0..max_spans days (projected rows, MAX_EMITS <= 8): a night shift contributes hours to
    two days, some shifts emit two rows for the SAME day (split shift), a few emit none."""
    s = splitmix64_stream(seed, 4 * n_shifts + 2 * n_shifts * max_spans)
    nurse = (s[:n_shifts] % np.uint64(n_nurses)).astype(np.int32)
    nurse[(s[n_shifts:2 * n_shifts] % np.uint64(1000)) < np.uint64(unassigned_permille)] = -1
    required = ((s[2 * n_shifts:3 * n_shifts] % np.uint64(4)) != 0).astype(np.uint8)
    n_spans = (s[3 * n_shifts:4 * n_shifts] % np.uint64(max_spans + 1)).astype(np.int64)
    rest = s[4 * n_shifts:]
    ptr = np.zeros(n_shifts + 1, dtype=np.uint32)
    ptr[1:] = np.cumsum(n_spans)
    days, hours = [], []
    k = 0
    for i in range(n_shifts):
        start = int(rest[k] % np.uint64(n_days))
        for j in range(int(n_spans[i])):
            split = int(rest[k + 1] % np.uint64(5)) == 0          # same day twice
            days.append(min(start + (0 if split else j), n_days - 1))
            hours.append(int(rest[k + 1] % np.uint64(7)) + 2)
            k += 2
    return RosterInstance(n_shifts, n_nurses, n_days, limit, required, ptr, np.array(days, dtype=np.int64),
                          np.array(hours, dtype=np.int64), nurse)


@dataclass
class AvailabilityInstance:
    """Shift x Employee (the fixture shape of the reference's cross-bi tests, constraint/tests/cross_bi_incr.rs:17-165)
    plus skills / hours / contracts for the authored pair-weight and multi-row-per-key joins."""
    n_shifts: int
    n_employees: int
    day: np.ndarray          # int64 per shift
    required: np.ndarray     # int64 per shift: required skill
    hours: np.ndarray        # int64 per shift
    employee: np.ndarray     # int32 per shift, -1 = unassigned (the planning variable)
    skill: np.ndarray        # int64 per employee
    un_ptr: np.ndarray       # uint32 CSR over employees: unavailable days
    un_days: np.ndarray      # uint32
    contracts: np.ndarray    # int64 [n_contracts, 4] = employee, from, to, fee


def availability(n_shifts: int = 60, n_employees: int = 7, n_days: int = 14, seed: int = 51,
                 unassigned_permille: int = 120) -> AvailabilityInstance:
    s = splitmix64_stream(seed, 6 * n_shifts + 4 * n_employees * 4 + 8)
    at = 0

    def take(n):
        nonlocal at
        out = s[at:at + n]
        at += n
        return out
    day = (take(n_shifts) % np.uint64(n_days)).astype(np.int64)
    required = (take(n_shifts) % np.uint64(5)).astype(np.int64)
    hours = (take(n_shifts) % np.uint64(3)).astype(np.int64) * 4 + 4
    employee = (take(n_shifts) % np.uint64(n_employees)).astype(np.int32)
    employee[(take(n_shifts) % np.uint64(1000)) < np.uint64(unassigned_permille)] = -1
    skill = (take(n_employees) % np.uint64(5)).astype(np.int64)
    un_lists = []
    pick = take(n_employees * 4)
    for e in range(n_employees):
        k = int(pick[4 * e] % np.uint64(4))       # 0..3 unavailable days
        un_lists.append(sorted({int(pick[4 * e + 1 + j] % np.uint64(n_days)) for j in range(k)}))
    un_ptr = np.concatenate([[0], np.cumsum([len(x) for x in un_lists])]).astype(np.uint32)
    un_days = np.array([d for x in un_lists for d in x], dtype=np.uint32)
    cs = take(n_employees * 4)
    contracts = []
    for e in range(n_employees):
        for j in range(int(cs[4 * e] % np.uint64(3))):   # 0..2 contract rows per employee
            lo = int(cs[4 * e + 1 + j] % np.uint64(n_days))
            contracts.append([e, lo, min(n_days - 1, lo + 4), 3 + 2 * j])
    contracts = np.array(contracts, dtype=np.int64).reshape(-1, 4)
    return AvailabilityInstance(n_shifts, n_employees, day, required, hours, employee, skill, un_ptr, un_days, contracts)
