"""CPU tests of the oracle: pinned against the reference's known-answer vectors (oracle/kat_main.cpp)
and against its own FullAssert invariant (incremental == evaluate_all,
solverforge-solver/src/scope/solver/scope_core.rs:642-653) on the config models."""
import subprocess

import numpy as np
import pytest

from solverforge_b200 import instances
from tests import oracle_lib
from tests.oracle_lib import Oracle


def test_reference_known_answer_vectors():
    oracle_lib.lib()
    out = subprocess.run([oracle_lib.KAT], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("KAT OK")


def test_nqueens_64_full_neighbourhood():
    # config C1: 64 x (64 values + None) = 4160 pulls, 64 not doable
    q = instances.nqueens(64)
    o = Oracle.nqueens(q)
    assert np.array_equal(o.committed_score(), o.evaluate_all())
    rows = o.enumerate_change()
    assert len(rows) == 4160
    assert np.array_equal(rows, instances.change_neighbourhood(q.row, q.n))
    scores, doable = o.score_change(rows)
    assert int((doable == 0).sum()) == 64
    # brute force (test_utils.rs:110-131 calculate_conflicts) on a few candidates
    def conflicts(row):
        c = 0
        for i in range(len(row)):
            for j in range(i + 1, len(row)):
                if row[i] >= 0 and row[j] >= 0 and (row[i] == row[j] or abs(row[i] - row[j]) == j - i):
                    c += 1
        return c
    for i in (0, 65, 1234, 4159):
        e, v = rows[i]
        r = q.row.copy()
        r[e] = v
        if doable[i]:
            assert scores[i][0] == -(conflicts(r) + int((r < 0).sum()))
    assert np.array_equal(o.committed_score(), o.evaluate_all())


@pytest.mark.parametrize("seed", [1, 2])
def test_graph_coloring_incremental_equals_full(seed):
    g = instances.graph_coloring(300, 1200, 4, seed_edges=seed, seed_colors=seed + 10, unassigned_permille=50)
    o = Oracle.graph_coloring(g)
    rows = instances.change_neighbourhood(g.color, g.k)
    scores, doable = o.score_change(rows)
    color = g.color.copy()
    pick = instances.splitmix64_stream(seed, 40) % np.uint64(len(rows))
    for i in pick:
        e, v = rows[int(i)]
        cur = o.score_change(np.array([[e, v]]))
        if not cur[1][0]:
            continue
        o.apply_change(e, v)
        color[e] = v
        assert np.array_equal(o.committed_score(), cur[0][0])
        assert np.array_equal(o.committed_score(), o.evaluate_all())
    fresh = Oracle.graph_coloring(g, color)
    assert np.array_equal(fresh.committed_score(), o.committed_score())


def test_cvrp_moves_round_trip():
    c = instances.cvrp(60, 6, seed=3)
    o = Oracle.cvrp(c)
    base = o.committed_score().copy()
    rows = o.enumerate_nearby_list_change(8)
    assert len(rows) == 60 * 8
    scores, doable = o.score_list_change(rows)
    assert doable.all()
    assert np.array_equal(o.committed_score(), base)  # do/undo restores the committed score
    for i in (0, 17, 300, 479):
        o.apply_list_change(*rows[i])
        assert np.array_equal(o.committed_score(), o.evaluate_all())
        rows2 = o.enumerate_nearby_list_change(8)
        s2, d2 = o.score_list_change(rows2[:50])
        assert np.array_equal(o.committed_score(), o.evaluate_all())
    # list swaps, including adjacent / same-route / cross-route
    swaps = np.array([[0, 0, 0, 1], [0, 0, 0, 3], [0, 1, 2, 0], [1, 2, 1, 2], [5, 0, 4, 99]])
    s, d = o.score_list_swap(swaps)
    assert list(d) == [1, 1, 1, 0, 0]
    o.apply_list_swap(*swaps[0])
    assert np.array_equal(o.committed_score(), s[0])
    assert np.array_equal(o.committed_score(), o.evaluate_all())


def test_job_shop_grouped_complement():
    j = instances.job_shop(12, 5, 4, seed=5, unassigned_permille=100)
    o = Oracle.job_shop(j)
    assert np.array_equal(o.committed_score(), o.evaluate_all())
    # soft = -(same job+machine pairs) - sum(load^2)
    m = j.machine_idx
    loads = np.bincount(m[m >= 0], minlength=j.n_machines)
    pairs = 0
    for a in range(j.n_ops):
        for b in range(a + 1, j.n_ops):
            if j.job[a] == j.job[b] and m[a] >= 0 and m[a] == m[b]:
                pairs += 1
    sc = o.committed_score()
    assert sc[1] == -(pairs + int((loads ** 2).sum()))
    assert sc[0] == -2 * int((m < 0).sum())  # unassigned operations are also unscheduled (not in any sequence)
    rows = instances.change_neighbourhood(m, j.n_machines)
    scores, doable = o.score_change(rows)
    for i in (0, 7, 33):
        if doable[i]:
            o.apply_change(*rows[i])
            assert np.array_equal(o.committed_score(), scores[i])
            assert np.array_equal(o.committed_score(), o.evaluate_all())
            scores, doable = o.score_change(rows)


def test_replay_matches_reference_forager_rules():
    # AcceptedCount(2) + HillClimbing: stops after the 2nd accepted pull (forager.rs:240-250)
    scores = np.array([[0, -12], [0, -8], [0, -9], [0, -1]])
    out = oracle_lib.replay_step(scores, [1, 1, 1, 1], [0, -10], [0, -10], [0, -10], 7, 0, 2, True, 0)
    assert out == (1, 1, 3, 3, 2)
    # BestScore + accept-all, First tie break
    scores = np.array([[0, -5], [0, -3], [0, -3], [0, -9]])
    out = oracle_lib.replay_step(scores, [1, 1, 1, 0], [0, 0], [0, 0], [0, 0], 1, 2, 0, False, 3)
    assert out[:3] == (1, 1, 4) and out[3] == 3


def test_shift_scheduling_invariant_and_unfairness():
    inst = instances.shift_scheduling(seed=21)
    o = Oracle.shift_scheduling(inst)
    assert np.array_equal(o.committed_score(), o.evaluate_all())
    # independent recompute of every constraint
    m = inst.nurse_idx
    hard = -int(((m < 0) & (inst.required != 0)).sum())
    for a in range(inst.n_shifts):
        for b in range(a + 1, inst.n_shifts):
            if inst.day[a] == inst.day[b] and m[a] >= 0 and m[a] == m[b]:
                hard -= 1
    counts = np.bincount(m[m >= 0], minlength=inst.n_nurses)
    soft = -int(np.abs(counts - inst.target).sum())
    # Long work streaks (schedule.rs:43-58): per nurse, runs of consecutive worked days, excess over 2 days
    for nurse in range(inst.n_nurses):
        days = sorted(set(int(x) for x in inst.day[m == nurse]))
        run = 0
        for i, day in enumerate(days):
            run = run + 1 if i > 0 and days[i - 1] + 1 == day else 1
            if i + 1 == len(days) or days[i + 1] != day + 1:
                soft -= max(0, run - 2)
    loads = {}
    for i in range(inst.n_shifts):
        if m[i] >= 0 and inst.hours[i] != 0:
            loads[int(m[i])] = loads.get(int(m[i]), 0) + int(inst.hours[i])
    if loads:
        v = np.array(list(loads.values()), dtype=np.float64)
        n = len(v)
        tmp = float(-(int(v.sum()) ** 2)) / n + float(int((v * v).sum())) if n > 1 else float(-(int(v.sum()) ** 2)) + float(int((v * v).sum()))
        import math
        unf = int(math.floor(math.sqrt(tmp) + 0.5))
        soft -= unf
    assert o.committed_score().tolist() == [hard, soft]
    rows = o.enumerate_change()
    s, ok = o.score_change(rows)
    for i in (0, 5, 50, 100):
        if ok[i]:
            o.apply_change(*rows[i])
            assert np.array_equal(o.committed_score(), o.evaluate_all())


def test_fast_cpu_baseline_matches_the_oracle():
    from tests.oracle_lib import FastCvrp
    c = instances.cvrp(150, 9, seed=4)
    offs, el = instances.perturb_routes(c, 3, 80)
    o = Oracle.cvrp(c, offs, el)
    f = FastCvrp(c, offs, el)
    rows = o.enumerate_nearby_list_change(15)
    extra = np.array([[0, 0, 0, 0], [0, 0, 0, 1], [0, 99, 1, 0], [1, 0, 2, 99], [3, 1, 3, 0]], dtype=np.uint32)
    rows = np.concatenate([rows, extra])
    so, oko = o.score_list_change(rows)
    sf, okf = f.score(rows)
    assert np.array_equal(okf, oko) and np.array_equal(sf, so)
    assert f.bench(rows, 2, 0.2) > 0


def test_fast_scalar_checkers_match_the_oracle():
    """The O(1) / O(degree) CPU checkers of the scalar configs (oracle/fast_cpu.cpp) are only trusted because they
    reproduce the reference-faithful oracle on every candidate of small instances, unassigned and not-doable
    rows included."""
    from tests.oracle_lib import FastGraphColoring, FastJobShop
    for seed in (1, 2, 3):
        g = instances.graph_coloring(300, 1200, 5, seed_edges=seed, seed_colors=seed + 10, unassigned_permille=60)
        o = Oracle.graph_coloring(g)
        rows = np.concatenate([instances.change_neighbourhood(g.color, g.k), [[5, int(g.color[5])], [7, -1]]])
        so, oko = o.score_change(rows)
        f = FastGraphColoring(g)
        sf, okf, com = f.score_change(g.color, rows)
        assert np.array_equal(okf, oko) and np.array_equal(sf, so)
        assert np.array_equal(com, o.committed_score())
        for with_c in (True, False):
            j = instances.job_shop(30, 6, 5, seed=seed, unassigned_permille=80)
            oj = Oracle.job_shop(j, with_complement=with_c)
            jr = np.concatenate([instances.change_neighbourhood(j.machine_idx, j.n_machines), [[3, int(j.machine_idx[3])], [4, -1]]])
            so, oko = oj.score_change(jr)
            fj = FastJobShop(j, with_c)
            sf, okf, com = fj.score_change(j.machine_idx, jr)
            assert np.array_equal(okf, oko) and np.array_equal(sf, so)
            assert np.array_equal(com, oj.committed_score())
    # FastCvrp rebinding keeps the converted matrix and follows the new routes
    from tests.oracle_lib import FastCvrp
    c = instances.cvrp(120, 7, seed=5)
    f = FastCvrp(c)
    for seed in (11, 12):
        offs, el = instances.perturb_routes(c, seed, 50)
        o = Oracle.cvrp(c, offs, el)
        f.set_routes(offs, el)
        rows = o.enumerate_nearby_list_change(12)
        so, oko = o.score_list_change(rows)
        sf, okf = f.score(rows)
        assert np.array_equal(okf, oko) and np.array_equal(sf, so)
        assert np.array_equal(f.committed_score(), o.committed_score())


def test_roster_projected_rows_incremental_equals_fresh():
    """Projected multi-emit rows in the oracle (ProjectedUni / ProjectedGrouped): cached == evaluate_all
    along a random walk of committed change moves (FullAssert invariant, scope_core.rs:642-653)."""
    from solverforge_b200 import instances
    inst = instances.roster(60, 4, 6, seed=2)
    o = Oracle.roster(inst)
    assert o.committed_score().tolist() == o.evaluate_all().tolist()
    r = instances.splitmix64_stream(12, 400)
    for i in range(200):
        e = int(r[2 * i] % np.uint64(inst.n_shifts))
        v = int(r[2 * i + 1] % np.uint64(inst.n_nurses + 1)) - 1
        o.apply_change(e, v)
        assert o.committed_score().tolist() == o.evaluate_all().tolist(), f"step {i}"


def test_sublist_moves_incremental_equals_fresh():
    """SublistChange / SublistSwap / ListReverse walks: the oracle's retained score equals a fresh evaluation after
    every applied move and do/undo evaluation leaves the committed score untouched (inverse layouts of
    segment_layout.rs)."""
    c = instances.cvrp(36, 5, seed=9)
    o = Oracle.cvrp(c)
    for step in range(6):
        base = o.committed_score().copy()
        for enum, score, apply in ((lambda: o.enumerate_sublist_change(1, 3), o.score_sublist_change, o.apply_sublist_change),
                                   (lambda: o.enumerate_sublist_swap(1, 3), o.score_sublist_swap, o.apply_sublist_swap),
                                   (lambda: o.enumerate_list_reverse()[:, :3], o.score_list_reverse, o.apply_list_reverse)):
            sample = enum()[step::37]   # enumerated against the current state
            s, d = score(sample if sample.shape[1] != 3 else np.concatenate([sample, np.zeros((len(sample), 1), sample.dtype)], axis=1))
            assert d.all()
            assert np.array_equal(o.committed_score(), base)
            pick = int(np.lexsort((s[:, 1], s[:, 0]))[-1])
            apply(*sample[pick])
            assert np.array_equal(o.committed_score(), s[pick])
            assert np.array_equal(o.committed_score(), o.evaluate_all())
            base = o.committed_score().copy()


def test_cluster_and_presence_models_incremental_equals_fresh():
    """Keyed tri / quad / penta joins (higher_arity/shared.rs) and the indexed_presence views along ChangeMove walks."""
    c = instances.cluster(26, 3, seed=5)
    o = Oracle.cluster(c)
    assert np.array_equal(o.committed_score(), o.evaluate_all())
    s_inst = instances.shift_scheduling(n_days=9, slots_per_day=2, n_nurses=3, seed=6, unassigned_permille=200)
    p = Oracle.shift_scheduling(s_inst, with_load_balance=False, presence_days=9)
    assert np.array_equal(p.committed_score(), p.evaluate_all())
    for o_ in (o, p):
        for step in range(8):
            rows = o_.enumerate_change()
            s, d = o_.score_change(rows[step::11])
            k = int(np.flatnonzero(d)[step % max(int(d.sum()), 1)])
            e, v = rows[step::11][k]
            o_.apply_change(int(e), int(v))
            assert np.array_equal(o_.committed_score(), s[k])
            assert np.array_equal(o_.committed_score(), o_.evaluate_all())
