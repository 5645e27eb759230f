"""The four config models authored with the device ConstraintFactory — each function is the
GPU counterpart of one reference `define_constraints()`.

  graph colouring  examples/scalar-graph-coloring/src/domain/graph_coloring.rs:21-44
  n-queens         examples/nqueens/src/domain/board.rs:21-47
  CVRP             crates/solverforge/tests/list_clarke_wright_publication/domain/publication_plan.rs:51-65
                   (+ authored capacity / distance constraints)
  mixed job-shop   examples/mixed-job-shop/src/domain/job_shop_plan.rs:28-69
                   (+ authored grouped-complement load constraint)
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .api import (AdjacentEqual, ConstraintFactory, Count, EqualId, EqualKey, EqualVarToRow, GpuScoreDirector,
                  HardSoftScore, ListSum, PathCost, Sum, hard, soft)
from .instances import CvrpInstance, GraphColoringInstance, JobShopInstance, NQueensInstance


def graph_coloring_director(inst: GraphColoringInstance, n_replicas: int = 1, colors=None, device: int = 0,
                            stream=None, flags: int = 0) -> GpuScoreDirector:
    d = GpuScoreDirector(n_replicas, device, stream, flags)
    d.add_collection("colors", inst.k, -1)
    nodes = d.add_collection("nodes", inst.n, 0)
    d.add_scalar_variable(nodes, "color_idx", inst.k, allows_unassigned=True)
    adj = d.add_csr("neighbors", inst.row_ptr, inst.col)
    f = ConstraintFactory(d)
    f.for_each(nodes).unassigned().penalize(HardSoftScore.ONE_HARD).named("Unassigned color")
    f.for_each(nodes).join(f.for_each(nodes), AdjacentEqual(adj)).penalize(HardSoftScore.ONE_HARD).named(
        "Adjacent color conflict")
    d.set_scalar_state(inst.color if colors is None else colors)
    d.commit()
    return d


def nqueens_director(inst: NQueensInstance, n_replicas: int = 1, rows=None, device: int = 0,
                     stream=None, flags: int = 0) -> GpuScoreDirector:
    """`left.column < right.column && (same row || same diagonal)` is the disjoint union of three
    equal-key joins: row, row + column, row - column."""
    d = GpuScoreDirector(n_replicas, device, stream, flags)
    d.add_collection("rows", inst.n, -1)
    queens = d.add_collection("queens", inst.n, 0)
    d.add_scalar_variable(queens, "row_idx", inst.n, allows_unassigned=True)
    column = d.add_column(queens, "column", np.arange(inst.n))
    f = ConstraintFactory(d)
    f.for_each(queens).unassigned().penalize(HardSoftScore.ONE_HARD).named("Unassigned queen")
    for name, mul in (("row", 0), ("diag+", 1), ("diag-", -1)):
        f.for_each(queens).join(f.for_each(queens), EqualKey(column, mul, 1)).penalize(HardSoftScore.ONE_HARD).named(
            f"Queen conflict ({name})")
    d.set_scalar_state(inst.row if rows is None else rows)
    d.commit()
    return d


def cluster_director(inst, n_replicas: int = 1, team=None, device: int = 0, stream=None,
                     flags: int = 0) -> GpuScoreDirector:
    """Keyed tri / quad / penta self-joins (`for_each(tasks).join(equal(team)).join(equal(team))…`,
    constraint/nary_incremental/higher_arity/shared.rs): a bucket of n rows holds C(n, arity) tuples."""
    d = GpuScoreDirector(n_replicas, device, stream, flags)
    d.add_collection("teams", inst.n_teams, -1)
    tasks = d.add_collection("tasks", inst.n, 0)
    d.add_scalar_variable(tasks, "team_idx", inst.n_teams, allows_unassigned=True)
    f = ConstraintFactory(d)
    f.for_each(tasks).unassigned().penalize(HardSoftScore.ONE_HARD).named("Unassigned task")
    for arity, weight in inst.joins:
        s = f.for_each(tasks).join(f.for_each(tasks), EqualKey())
        for _ in range(arity - 2):
            s = s.join(f.for_each(tasks), EqualKey())
        s.penalize(HardSoftScore(0, weight)).named(f"Cluster arity {arity}")
    d.set_scalar_state(inst.team if team is None else team)
    d.commit()
    return d


def cvrp_director(inst: CvrpInstance, n_replicas: int = 1, offsets=None, elems=None, device: int = 0,
                  stream=None, flags: int = 0, distance_weight: int = 1, capacity_weight: int = 1) -> GpuScoreDirector:
    """distance_weight / capacity_weight scale the two authored weights (tests of the arithmetic-width switch)."""
    d = GpuScoreDirector(n_replicas, device, stream, flags)
    locations = d.add_collection("locations", inst.dim, -1)       # matrix rows: depot + customers
    customers = d.add_collection("customers", inst.dim - 1, -1)   # problem facts, id = 1..n
    routes = d.add_collection("routes", inst.n_routes, 0)
    d.add_list_variable(routes, locations, "visits")
    cust_id = d.add_column(customers, "id", np.array([i for i in range(inst.dim) if i != inst.depot]))
    demand = d.add_column(locations, "demand", inst.demands)
    dist = d.add_matrix("distance_matrix", inst.matrix, cost_semantics=True)
    f = ConstraintFactory(d)
    f.for_each(customers).if_not_exists(f.for_each(routes).flattened(), EqualId(cust_id)).penalize(
        HardSoftScore.ONE_HARD).named("all_customers_assigned")
    f.for_each(routes).penalize(hard(L.W_EXCESS, capacity_weight, inst.capacity), ListSum(demand)).named("vehicle_capacity")
    f.for_each(routes).penalize(soft(L.W_LINEAR, distance_weight, 0), PathCost(dist, inst.depot)).named("total_distance")
    d.set_list_state(inst.offsets if offsets is None else offsets, inst.elems if elems is None else elems)
    d.commit()
    return d


def job_shop_director(inst: JobShopInstance, n_replicas: int = 1, machine_idx=None, with_complement: bool = True,
                      device: int = 0, stream=None, flags: int = 0) -> GpuScoreDirector:
    d = GpuScoreDirector(n_replicas, device, stream, flags)
    machines = d.add_collection("machines", inst.n_machines, -1)
    ops = d.add_collection("operations", inst.n_ops, 0)
    seqs = d.add_collection("machine_sequences", inst.n_machines, 1)
    d.add_scalar_variable(ops, "machine_idx", inst.n_machines, allows_unassigned=True)
    d.add_list_variable(seqs, ops, "operations")
    job = d.add_column(ops, "job", inst.job)
    f = ConstraintFactory(d)
    f.for_each(ops).unassigned().penalize(HardSoftScore.ONE_HARD).named("Unassigned operation machine")
    f.for_each(ops).if_not_exists(f.for_each(seqs).flattened(), EqualId()).penalize(HardSoftScore.ONE_HARD).named(
        "Unscheduled operation")
    f.for_each(ops).join(f.for_each(ops), EqualKey(job, inst.n_machines, 1)).penalize(HardSoftScore.ONE_SOFT).named(
        "Same job machine reuse")
    if with_complement:
        f.for_each(ops).join(machines, EqualVarToRow()).group_by(Count()).complement(machines, 0).penalize(
            soft(L.W_SQUARE, 1, 0)).named("Machine load balance")
    d.set_scalar_state(inst.machine_idx if machine_idx is None else machine_idx)
    d.set_list_state(inst.seq_offsets, inst.seq_elems)
    d.commit()
    return d


def shift_scheduling_director(inst, n_replicas: int = 1, nurse_idx=None, device: int = 0, stream=None,
                              flags: int = 0, with_load_balance: bool = True, presence_days: int = 0) -> GpuScoreDirector:
    """examples/minimal-shift-scheduling/src/domain/schedule.rs:21-84 (all four constraints, incl. "Long work
    streaks" over the consecutive_runs collector) + an authored load_balance constraint."""
    from .api import ConsecutiveRuns, IndexedPresence, LoadBalance
    d = GpuScoreDirector(n_replicas, device, stream, flags)
    nurses = d.add_collection("nurses", inst.n_nurses, -1)
    shifts = d.add_collection("shifts", inst.n_shifts, 0)
    d.add_scalar_variable(shifts, "nurse_idx", inst.n_nurses, allows_unassigned=True)
    day = d.add_column(shifts, "day", inst.day)
    required = d.add_column(shifts, "required", inst.required)
    hours = d.add_column(shifts, "hours", inst.hours)
    f = ConstraintFactory(d)
    f.for_each(shifts).filter(required).unassigned().penalize(HardSoftScore.ONE_HARD).named("Unassigned required shift")
    f.for_each(shifts).join(f.for_each(shifts), EqualKey(day, inst.n_nurses, 1)).penalize(
        HardSoftScore.ONE_HARD).named("One shift per nurse day")
    f.for_each(shifts).assigned().group_by(ConsecutiveRuns(day, int(np.max(inst.day)) + 1)).penalize(
        soft(L.W_EXCESS, 1, 2)).named("Long work streaks")
    if presence_days > 0:
        # the indexed_presence example of the reference (solverforge-macros/tests/ui/pass/
        # solverforge_constraints_indexed_presence.rs:13-55) + an authored distinct-days constraint
        n_points = max(int(np.max(inst.day)) + 1, presence_days, 7)   # the views' ranges must lie inside the table
        f.for_each(shifts).assigned().group_by(IndexedPresence(day, n_points, "complement_runs", 0, presence_days)).penalize(
            soft(L.W_EXCESS, 1, 1)).named("Rest streaks")
        f.for_each(shifts).assigned().group_by(IndexedPresence(day, n_points, "any_in", 5, 7)).penalize(
            soft(L.W_CONST, 1, 0)).named("Weekend work")
        f.for_each(shifts).assigned().group_by(IndexedPresence(day, n_points, "count")).penalize(
            soft(L.W_LINEAR, 2, 0)).named("Days worked")
    f.for_each(shifts).assigned().group_by(Count()).complement(nurses, 0).penalize(
        soft(L.W_ABSDIFF, 1, inst.target)).named("Balanced workload")
    if with_load_balance:
        f.for_each(shifts).assigned().group_by(LoadBalance(hours)).penalize(soft(L.W_LINEAR, 1, 0)).named("Fair hours")
    d.set_scalar_state(inst.nurse_idx if nurse_idx is None else nurse_idx)
    d.commit()
    return d


def roster_director(inst, n_replicas: int = 1, nurse_idx=None, device: int = 0, stream=None,
                    flags: int = 0) -> GpuScoreDirector:
    """Projected scoring rows (`.project(..)`, stream/projected_stream/): every assigned shift emits one row per
    spanned day; grouped by (nurse, day) with sum(hours) / count(), plus a per-row terminal."""
    from .api import Projection
    d = GpuScoreDirector(n_replicas, device, stream, flags)
    d.add_collection("nurses", inst.n_nurses, -1)
    shifts = d.add_collection("shifts", inst.n_shifts, 0)
    d.add_scalar_variable(shifts, "nurse_idx", inst.n_nurses, allows_unassigned=True)
    required = d.add_column(shifts, "required", inst.required)
    spans = d.add_csr("spans", inst.span_ptr, inst.span_day)
    rows = Projection(spans, inst.n_days, inst.span_hours)
    f = ConstraintFactory(d)
    f.for_each(shifts).filter(required).unassigned().penalize(HardSoftScore.ONE_HARD).named("Unassigned required shift")
    f.for_each(shifts).project(rows).group_by(Sum(L.NO_COLUMN)).penalize(hard(L.W_EXCESS, 1, inst.limit)).named("Daily hours")
    f.for_each(shifts).project(rows).group_by(Count()).penalize(soft(L.W_SQUARE, 1, 0)).named("Fragmented days")
    f.for_each(shifts).project(rows).join_self().penalize(HardSoftScore.ONE_HARD).named("Double booking")
    f.for_each(shifts).project(rows).penalize(soft(L.W_LINEAR, 1, 0)).named("Worked hours")
    d.set_scalar_state(inst.nurse_idx if nurse_idx is None else nurse_idx)
    d.commit()
    return d


def availability_director(inst, n_replicas: int = 1, employee=None, device: int = 0, stream=None,
                          flags: int = 0) -> GpuScoreDirector:
    """General cross-collection joins with pair filters and pair weights (SFGPU_K_JOIN_EXPR): the Shift x Employee
    fixture of the reference's cross-bi tests (constraint/tests/cross_bi_incr.rs:63-88 "Unavailable employee",
    :308-341 index-aware filter) plus authored pair-weight and multi-row-per-key joins — the same five constraints
    as the oracle's AvailabilityModel."""
    from .api import EqualVarToKey, Expr
    d = GpuScoreDirector(n_replicas, device, stream, flags)
    employees = d.add_collection("employees", inst.n_employees, -1)
    contracts = d.add_collection("contracts", max(len(inst.contracts), 1), -1)
    shifts = d.add_collection("shifts", inst.n_shifts, 0)
    d.add_scalar_variable(shifts, "employee", inst.n_employees, allows_unassigned=True)
    day = d.add_column(shifts, "day", inst.day)
    required = d.add_column(shifts, "required", inst.required)
    hours = d.add_column(shifts, "hours", inst.hours)
    skill = d.add_column(employees, "skill", inst.skill)
    unavailable = d.add_csr("unavailable_days", inst.un_ptr, inst.un_days)
    n_c = len(inst.contracts)
    pad = np.zeros(max(n_c, 1), dtype=np.int64)
    c_from, c_to, c_fee = pad.copy(), pad.copy(), pad.copy()
    if n_c:
        c_from[:], c_to[:], c_fee[:] = inst.contracts[:, 1], inst.contracts[:, 2], inst.contracts[:, 3]
    c_from = d.add_column(contracts, "from", c_from)
    c_to = d.add_column(contracts, "to", c_to)
    c_fee = d.add_column(contracts, "fee", c_fee)
    # bucket CSR: employee id -> its contract rows
    order = np.argsort(inst.contracts[:, 0], kind="stable") if n_c else np.zeros(0, dtype=np.int64)
    counts = np.bincount(inst.contracts[:, 0], minlength=inst.n_employees) if n_c else np.zeros(inst.n_employees, dtype=np.int64)
    by_employee = d.add_csr("contracts_by_employee", np.concatenate([[0], np.cumsum(counts)]), order)
    f = ConstraintFactory(d)
    f.for_each(shifts).unassigned().penalize(HardSoftScore.ONE_HARD).named("Unassigned shift")
    joined = f.for_each(shifts).join(f.for_each(employees), EqualVarToRow())
    joined.filter(Expr.csr_contains(unavailable, Expr.b_index(), Expr.a(day))).penalize(HardSoftScore.ONE_HARD) \
        .named("Unavailable employee")
    joined.filter(Expr.a(required) > Expr.b(skill)) \
        .penalize(soft(L.W_LINEAR, 1, 0), (Expr.a(required) - Expr.b(skill)) * Expr.a(hours)).named("Skill gap")
    joined.filter(((Expr.a_index() + Expr.b_index() * 2) % 3).eq(0)).penalize(soft(L.W_LINEAR, 1, 0), Expr.a(day)) \
        .named("Indexed pairs")
    f.for_each(shifts).join(f.for_each(contracts), EqualVarToKey(by_employee)) \
        .filter((Expr.a(day) < Expr.b(c_from)) | (Expr.a(day) > Expr.b(c_to))) \
        .penalize(soft(L.W_LINEAR, 1, 0), Expr.b(c_fee)).named("Contract window")
    d.set_scalar_state(inst.employee if employee is None else employee)
    d.commit()
    return d


def pairs_director(demand, prio, bucket, n_buckets: int, n_replicas: int = 1, device: int = 0, stream=None,
                   flags: int = 0) -> GpuScoreDirector:
    """Keyed self-joins of entity rows with pair filters / pair weights, undirected and directed
    (SFGPU_K_PAIR_KEY_EXPR): the Work{bucket, demand} fixture of the reference's projected self-join tests
    (constraint/tests/projected/self_join.rs:108-290) with the bucket as the planning variable — the four constraints
    of the oracle's PairsModel."""
    from .api import EqualKeyExpr, Expr
    bucket = np.asarray(bucket)
    n = bucket.shape[-1]
    d = GpuScoreDirector(n_replicas, device, stream, flags)
    d.add_collection("buckets", n_buckets, -1)
    work = d.add_collection("work", n, 0)
    d.add_scalar_variable(work, "bucket", n_buckets, allows_unassigned=True)
    dem = d.add_column(work, "demand", demand)
    pri = d.add_column(work, "prio", prio)
    f = ConstraintFactory(d)
    f.for_each(work).unassigned().penalize(HardSoftScore.ONE_HARD).named("Unassigned work")
    same_bucket = EqualKeyExpr(Expr.a_value(), n_buckets)
    f.for_each(work).join(f.for_each(work), same_bucket).filter(Expr.a(dem) < Expr.b(dem)) \
        .penalize(HardSoftScore(0, 1)).named("projected duplicate bucket")
    f.for_each(work).join(f.for_each(work), same_bucket).penalize(soft(L.W_LINEAR, 1, 0), abs(Expr.a(pri) - Expr.b(pri))) \
        .named("priority spread")
    f.for_each(work).join(f.for_each(work), EqualKeyExpr(Expr.a_value(), n_buckets, right_key=Expr.a(dem))) \
        .penalize(soft(L.W_LINEAR, 1, 0), Expr.a_value() * 10 + Expr.b_value()).named("projected parent child")
    d.set_scalar_state(bucket)
    d.commit()
    return d
