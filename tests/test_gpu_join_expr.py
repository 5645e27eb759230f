"""GPU parity of the general cross-collection join with pair filter / pair weight expressions (SFGPU_K_JOIN_EXPR,
the reference's CrossBiConstraint, constraint/cross_bi_incremental/state.rs:260-460) and of the direct if_exists
(constraint/exists.rs:167-272): the reference's own known answers (tests/golden/reference_kats.json) and the
availability model against the oracle's closure-based CrossBiConstraint on every move kind. Bit-exact."""
import json
import os

import numpy as np
import pytest

from solverforge_b200 import (ConstraintFactory, EqualVarToRow, Expr, ForageParams, GpuScoreDirector, HardSoftScore, instances,
                              models, soft)
from solverforge_b200 import _lib as L
from tests import oracle_lib
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


def _csr(lists):
    ptr = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.uint32)
    return ptr, np.array([v for x in lists for v in x], dtype=np.uint32)


def test_reference_kats_cross_bi_pair_filter_and_index_aware_filter():
    k = GOLDEN["cross_bi_pair_filter"][0]
    d = GpuScoreDirector(1)
    employees = d.add_collection("employees", k["n_employees"], 1)
    shifts = d.add_collection("shifts", len(k["shift_employee"]), 0)
    d.add_scalar_variable(shifts, "employee_id", k["n_employees"])
    day = d.add_column(shifts, "day", k["shift_day"])
    un = d.add_csr("unavailable_days", *_csr(k["unavailable"]))
    f = ConstraintFactory(d)
    f.for_each(shifts).join(f.for_each(employees), EqualVarToRow()) \
        .filter(Expr.csr_contains(un, Expr.b_index(), Expr.a(day))).penalize(HardSoftScore(0, 1)).named("Unavailable employee")
    d.set_scalar_state(k["shift_employee"])
    assert d.commit()[0].tolist() == [0, k["soft"]]
    # on_retract(entity 0): the pair of shift 0 disappears (cross_bi_incr.rs:252-264)
    sc, ok = d.score_change(np.array([[0, -1]]))
    assert ok[0] == 1 and sc[0].tolist() == [0, k["retract_row0_soft"]]
    k = GOLDEN["cross_bi_pair_filter"][1]
    d = GpuScoreDirector(1)
    employees = d.add_collection("employees", k["n_employees"], 1)
    shifts = d.add_collection("shifts", len(k["shift_employee"]), 0)
    d.add_scalar_variable(shifts, "employee_id", k["n_employees"])
    day = d.add_column(shifts, "day", k["shift_day"])
    f = ConstraintFactory(d)
    f.for_each(shifts).join(f.for_each(employees), EqualVarToRow()) \
        .filter(Expr.a_index().eq(1) & Expr.b_index().eq(0)).penalize(soft(L.W_LINEAR, 1, 0), Expr.a(day)) \
        .named("indexed cross path")
    d.set_scalar_state(k["shift_employee"])
    assert d.commit()[0].tolist() == [0, k["indexed_soft"]]


def test_reference_kat_direct_exists():
    k = GOLDEN["exists_direct"][0]
    for avail, want in ((k["available_before"], k["soft_before"]), (k["available_after"], k["soft_after"])):
        d = GpuScoreDirector(1)
        workers = d.add_collection("workers", k["n_workers"], -1)
        tasks = d.add_collection("tasks", len(k["assignee"]), 0)
        d.add_scalar_variable(tasks, "assignee", k["n_workers"])
        unavailable = d.add_column(workers, "unavailable", [0 if a else 1 for a in avail])
        f = ConstraintFactory(d)
        f.for_each(tasks).assigned().if_exists(f.for_each(workers).filter(unavailable), EqualVarToRow()) \
            .penalize(HardSoftScore(0, 1)).named("unavailable worker")
        d.set_scalar_state(k["assignee"])
        assert d.commit()[0].tolist() == [0, want]
    # if_not_exists is the complement on the assigned tasks
    d = GpuScoreDirector(1)
    workers = d.add_collection("workers", 2, -1)
    tasks = d.add_collection("tasks", 3, 0)
    d.add_scalar_variable(tasks, "assignee", 2)
    unavailable = d.add_column(workers, "unavailable", [1, 0])
    f = ConstraintFactory(d)
    f.for_each(tasks).assigned().if_not_exists(f.for_each(workers).filter(unavailable), EqualVarToRow()) \
        .penalize(HardSoftScore(0, 1)).named("available worker")
    d.set_scalar_state([0, 0, 1])
    assert d.commit()[0].tolist() == [0, -1]


@pytest.mark.parametrize("seed", [51, 52])
def test_availability_joins_match_the_oracle_on_every_move_kind(seed):
    inst = instances.availability(60, 7, 14, seed=seed)
    R = 3
    starts = np.stack([instances.availability(60, 7, 14, seed=seed + 10 * (r + 1)).employee for r in range(R)])
    d = models.availability_director(inst, R, employee=starts)
    assert d.scalar_program() == -1     # expression joins run on the constraint-table interpreter
    oracles = [Oracle.availability(inst, starts[r]) for r in range(R)]
    for r, o in enumerate(oracles):
        assert d.calculate_score()[r].tolist() == o.committed_score().tolist() == d.fresh_score()[r].tolist()
    # full ChangeMove neighbourhood of every replica
    rows = [o.enumerate_change() for o in oracles]
    offs = np.concatenate([[0], np.cumsum([len(x) for x in rows])]).astype(np.uint64)
    sc, ok = d.score_change(np.concatenate(rows), offs)
    for r, o in enumerate(oracles):
        so, oko = o.score_change(rows[r])
        assert np.array_equal(ok[offs[r]:offs[r + 1]], oko)
        assert np.array_equal(sc[offs[r]:offs[r + 1]], so)
    # swaps and compound edits (sequential overlay) against replica 0
    o = oracles[0]
    swaps = o.enumerate_swap()[:600]
    sc, ok = d.score_swap(np.concatenate([swaps] * R), np.arange(R + 1, dtype=np.uint64) * len(swaps))
    for r in range(R):
        so, oko = oracles[r].score_swap(swaps)
        assert np.array_equal(ok[r * len(swaps):(r + 1) * len(swaps)], oko) and np.array_equal(sc[r * len(swaps):(r + 1) * len(swaps)], so)
    rng = np.random.default_rng(seed)
    n_c = 150
    sizes = rng.integers(1, 5, size=n_c)
    eo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    er = np.stack([rng.integers(0, inst.n_shifts, size=eo[-1]), rng.integers(-1, inst.n_employees, size=eo[-1])], axis=1)
    d1 = models.availability_director(inst, 1, employee=starts[0][None, :])
    sc, ok = d1.score_compound(eo, er)
    so, oko = o.score_compound(eo, er)
    assert np.array_equal(ok, oko) and np.array_equal(sc[ok == 1], so[oko == 1])
    # device-enumerated step + commit chain: committed == fresh == oracle after every step
    for step in range(12):
        last = d.calculate_score()
        ref = np.concatenate([last, last], axis=1)
        seeds = [100 + step, 200 + step, 300 + step]
        idx, best, ev, win = d.step_change(ForageParams(1, 1, 0), step_seeds=seeds, ref_scores=ref, apply=True)
        for r, orc in enumerate(oracles):
            rws = orc.enumerate_change()
            so, oko = orc.score_change(rws)
            out = oracle_lib.replay_step(so, oko, [0, 0], last[r], last[r], seeds[r], 2, 1, True, 0)
            assert int(ev[r]) == out[2]
            if out[0]:
                assert int(idx[r]) == out[1] and best[r].tolist() == so[out[1]].tolist()
                orc.apply_change(*[int(x) for x in rws[out[1]]])
            else:
                assert idx[r] == 0xFFFFFFFF
        now = d.calculate_score()
        for r, orc in enumerate(oracles):
            assert now[r].tolist() == orc.committed_score().tolist() == d.fresh_score()[r].tolist()


def test_expression_errors_are_loud():
    d = GpuScoreDirector(1)
    a = d.add_collection("a", 3, 0)
    d.add_scalar_variable(a, "v", 2)
    with pytest.raises(L.SfgpuError):
        d.add_expr(Expr([(L.X_ADD, 0, 0)]))                       # stack underflow
    with pytest.raises(L.SfgpuError):
        d.add_expr(Expr([(L.X_CONST, 0, 1), (L.X_CONST, 0, 2)]))  # two results
    with pytest.raises(L.SfgpuError):
        d.add_expr(Expr([(L.X_A_COL, 99, 0)]))                    # unknown column
    with pytest.raises(L.SfgpuError):
        d.add_expr(Expr([(99, 0, 0)]))                            # unknown op


def _pairs_kat_director(k, directed):
    from solverforge_b200 import EqualKeyExpr
    d = GpuScoreDirector(1)
    d.add_collection("buckets", k["n_buckets"], -1)
    work = d.add_collection("work", len(k["bucket"]), 0)
    d.add_scalar_variable(work, "bucket", k["n_buckets"])
    return d, work


def test_reference_kats_projected_self_join_with_pair_filter_and_directed_join():
    """constraint/tests/projected/self_join.rs:108-290 as SFGPU_K_PAIR_KEY_EXPR."""
    from solverforge_b200 import EqualKeyExpr
    k = GOLDEN["projected_self_join_pairs"][0]
    d, work = _pairs_kat_director(k, False)
    dem = d.add_column(work, "demand", k["demand"])
    f = ConstraintFactory(d)
    f.for_each(work).join(f.for_each(work), EqualKeyExpr(Expr.a_value(), k["n_buckets"])).filter(Expr.a(dem) < Expr.b(dem)) \
        .penalize(HardSoftScore(0, 1)).named("projected duplicate bucket")
    d.set_scalar_state(k["bucket"])
    assert d.commit()[0].tolist() == [0, k["soft"]]
    sc, ok = d.score_change(np.array([k["change"]]))
    assert ok[0] == 1 and sc[0].tolist() == [0, k["soft_after"]]
    d.apply_change(np.array([k["change"]]))
    assert d.calculate_score()[0].tolist() == d.fresh_score()[0].tolist() == [0, k["soft_after"]]
    for k in GOLDEN["projected_self_join_pairs"][1:]:
        for demand, want in ((k["demand"], k["directed_soft"]),) + (((k["demand_after"], k["directed_soft_after"]),)
                                                                    if "demand_after" in k else ()):
            d, work = _pairs_kat_director(k, True)
            dem = d.add_column(work, "demand", demand)
            f = ConstraintFactory(d)
            f.for_each(work).join(f.for_each(work), EqualKeyExpr(Expr.a_value(), k["n_buckets"], right_key=Expr.a(dem))) \
                .penalize(soft(L.W_LINEAR, 1, 0), Expr.a_value() * 10 + Expr.b_value()).named("projected parent child")
            d.set_scalar_state(k["bucket"])
            assert d.commit()[0].tolist() == [0, want], k["cite"]


@pytest.mark.parametrize("seed", [7, 8])
def test_pair_key_expression_joins_match_the_oracle_on_every_move_kind(seed):
    rng = np.random.default_rng(seed)
    n, kb, R = 70, 6, 3
    demand = rng.integers(-1, kb, size=n)
    prio = rng.integers(0, 9, size=n)
    starts = rng.integers(-1, kb, size=(R, n)).astype(np.int32)
    d = models.pairs_director(demand, prio, starts, kb, R)
    oracles = [Oracle.pairs(demand, prio, starts[r], kb) for r in range(R)]
    for r, o in enumerate(oracles):
        assert d.calculate_score()[r].tolist() == o.committed_score().tolist() == d.fresh_score()[r].tolist()
    rows = [o.enumerate_change() for o in oracles]
    offs = np.concatenate([[0], np.cumsum([len(x) for x in rows])]).astype(np.uint64)
    sc, ok = d.score_change(np.concatenate(rows), offs)
    for r, o in enumerate(oracles):
        so, oko = o.score_change(rows[r])
        assert np.array_equal(ok[offs[r]:offs[r + 1]], oko) and np.array_equal(sc[offs[r]:offs[r + 1]], so)
    swaps = oracles[0].enumerate_swap()[:800]
    sc, ok = d.score_swap(np.concatenate([swaps] * R), np.arange(R + 1, dtype=np.uint64) * len(swaps))
    for r in range(R):
        so, oko = oracles[r].score_swap(swaps)
        sl = slice(r * len(swaps), (r + 1) * len(swaps))
        assert np.array_equal(ok[sl], oko) and np.array_equal(sc[sl], so)
    n_c = 200
    sizes = rng.integers(1, 6, size=n_c)
    eo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    er = np.stack([rng.integers(0, n, size=eo[-1]), rng.integers(-1, kb, size=eo[-1])], axis=1)
    d1 = models.pairs_director(demand, prio, starts[0][None, :], kb, 1)
    sc, ok = d1.score_compound(eo, er)
    so, oko = oracles[0].score_compound(eo, er)
    assert np.array_equal(ok, oko) and np.array_equal(sc[ok == 1], so[oko == 1])
    # the default scalar search (Change + Swap union, seeded Random) commits through the list maintenance
    desc = GpuScoreDirector.default_scalar_union(window=16)
    for step in range(15):
        last = d.calculate_score()
        ref = np.concatenate([last, last], axis=1)
        seeds = [500 + step, 600 + step, 700 + step]
        idx, best, ev, win, flags = d.step_union(desc, ForageParams(1, 1, 40), step_seeds=seeds, step_indices=[step] * R,
                                                 ref_scores=ref, apply=True)
        for r, o in enumerate(oracles):
            out, child, local, data, scs = oracle_lib.union_step(o, [(L.FAM_CHANGE,), (L.FAM_SWAP,)], L.UNION_STRATIFIED_RANDOM,
                                                                 L.ORDER_RANDOM, step, seeds[r], last[r], last[r], 0, 40, True, 0)
            assert int(ev[r]) == out[2] and flags[r] == 0
            if out[0]:
                assert int(idx[r]) == out[1]
                ci, j = int(child[out[1]]), int(local[out[1]])
                a, b = [int(x) for x in data[ci][0][j]]
                (o.apply_change if ci == 0 else o.apply_swap)(a, b)
            else:
                assert idx[r] == 0xFFFFFFFF
        now = d.calculate_score()
        for r, o in enumerate(oracles):
            assert now[r].tolist() == o.committed_score().tolist() == d.fresh_score()[r].tolist()
