"""GPU parity of the union step (sfgpu_step_union, sfgpu_union.cuh) against the reference's default list local search
as the oracle restates it: every leaf cursor in the seeded SelectionOrder of the step's MoveStreamContext
(move_selector/iter.rs:109-125), the children interleaved by the UnionScheduler (decorator/vec_union.rs:190-366),
acceptor + forager replayed in union pull order (phase/candidates.rs:66-282). Winner CandidateId, score,
moves_evaluated, the winning move and the committed state are compared bit for bit."""
import numpy as np
import pytest

from solverforge_b200 import ForageParams, GpuScoreDirector, instances, models
from solverforge_b200 import _lib as L
from tests import oracle_lib
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu

DEFAULT = [(L.FAM_NEARBY_LIST_CHANGE, 20), (L.FAM_NEARBY_LIST_SWAP, 20), (L.FAM_SUBLIST_CHANGE, 1, 3),
           (L.FAM_SUBLIST_SWAP, 1, 3), (L.FAM_LIST_REVERSE,)]


def _setup(n=46, routes=7, seed=31, R=3, coarse=True, moves=30):
    c = instances.cvrp(n, routes, seed=seed)
    if coarse:
        c.matrix = (c.matrix // 40) * 40     # many equal distances and scores: tie rules matter
    starts = [instances.perturb_routes(c, 40 + r, moves + 10 * r) for r in range(R)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    return c, d, oracles


def _check(d, oracles, children, union_order, order, seeds, steps, acceptor, okind, ties, limit, dl=0, window=0,
           max_window=0, weights=None):
    kids = [tuple(ch) + ((weights[i],) if weights else ()) for i, ch in enumerate(children)]
    if weights:
        kids = [(ch[0], ch[1] if len(ch) > 1 else 0, ch[2] if len(ch) > 2 else 0, weights[i]) for i, ch in enumerate(children)]
    desc = GpuScoreDirector.union_desc(kids, union_order, order, window, max_window)
    base = d.calculate_score()
    ref = np.concatenate([base + [0, dl], base + [0, dl - 7]], axis=1)
    idx, best, ev, win, flags = d.step_union(desc, ForageParams(acceptor, ties, limit), step_seeds=seeds, step_indices=steps,
                                             ref_scores=ref)
    for r, o in enumerate(oracles):
        out, child, local, data, sc = oracle_lib.union_step(o, children, union_order, order, steps[r], seeds[r], ref[r][:2],
                                                            ref[r][2:], 0 if limit else 2, max(limit, 1), bool(ties), okind,
                                                            weights)
        what = f"r={r} union={union_order} order={order} acc={acceptor} ties={ties} limit={limit} window={window}"
        assert flags[r] == 0, what
        assert int(ev[r]) == out[2], what + " moves_evaluated"
        if out[0]:
            t = out[1]
            assert int(idx[r]) == t, what
            assert best[r].tolist() == sc[t].tolist(), what
            c, j = int(child[t]), int(local[t])
            assert win[r].tolist() == [children[c][0], c] + data[c][1][j].tolist() + [j, 0], what
        else:
            assert idx[r] == 0xFFFFFFFF, what


@pytest.mark.parametrize("order", [L.ORDER_ORIGINAL, L.ORDER_RANDOM, L.ORDER_SHUFFLED])
@pytest.mark.parametrize("family", [0, 1, 2, 3, 4])
def test_single_family_cursor_in_every_selection_order(family, order):
    """One child, sequential union: the device walker of every family against the oracle's cursor."""
    _, d, oracles = _setup()
    child = [DEFAULT[family]]
    seeds, steps = [5, 77, 0xDEADBEEF], [0, 3, 900]
    for acceptor, okind, dl in ((0, 3, 0), (1, 0, -30), (2, 1, 0)):
        for ties in (0, 1):
            for limit in (1, 17, 300):
                _check(d, oracles, child, L.UNION_SEQUENTIAL, order, seeds, steps, acceptor, okind, ties, limit, dl)


@pytest.mark.parametrize("union_order", [L.UNION_SEQUENTIAL, L.UNION_ROUND_ROBIN, L.UNION_ROTATING_ROUND_ROBIN,
                                         L.UNION_RANDOM, L.UNION_STRATIFIED_RANDOM])
def test_default_list_union_matches_oracle(union_order):
    _, d, oracles = _setup()
    seeds, steps = [11, 12345, 0xFEEDF00D], [0, 1, 77]
    for order in (L.ORDER_ORIGINAL, L.ORDER_RANDOM, L.ORDER_SHUFFLED):
        for acceptor, okind, dl in ((0, 3, 0), (1, 0, -30), (2, 1, 0)):
            for ties, limit in ((1, 1), (0, 40), (1, 256), (1, 3000)):
                _check(d, oracles, DEFAULT, union_order, order, seeds, steps, acceptor, okind, ties, limit, dl,
                       max_window=1 << 14)


def test_union_windows_grow_until_the_forager_quits_and_weights():
    """Tiny first window: the step needs several passes; weighted children (Random / StratifiedRandom only)."""
    _, d, oracles = _setup(n=30, routes=5, seed=3, R=2)
    seeds, steps = [9, 10], [4, 5]
    for limit in (5, 200, 2000):
        _check(d, oracles, DEFAULT, L.UNION_STRATIFIED_RANDOM, L.ORDER_RANDOM, seeds, steps, 2, 1, 1, limit, window=4,
               max_window=1 << 16)
        _check(d, oracles, DEFAULT, L.UNION_STRATIFIED_RANDOM, L.ORDER_RANDOM, seeds, steps, 2, 1, 1, limit, window=8,
               max_window=1 << 16, weights=[3, 1, 2, 1, 5])
        _check(d, oracles, DEFAULT, L.UNION_RANDOM, L.ORDER_SHUFFLED, seeds, steps, 1, 0, 0, limit, dl=-30, window=16,
               max_window=1 << 16, weights=[1, 4, 1, 2, 1])


def test_union_whole_stream_best_score_and_degenerate_routes():
    """BestScore forager (never quits): the whole union stream of a small instance, with an empty and a singleton route."""
    c = instances.cvrp(20, 5, seed=36)
    offs, el = instances.perturb_routes(c, 4, 20)
    lists = [el[offs[i]:offs[i + 1]].tolist() for i in range(5)]
    lists[0] += lists[2]
    lists[2] = []
    while len(lists[4]) > 1:
        lists[1].append(lists[4].pop())
    offs = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.uint32)
    el = np.array([x for l in lists for x in l], dtype=np.uint32)
    d = models.cvrp_director(c, 1, offsets=offs[None, :], elems=el)
    o = Oracle.cvrp(c, offs, el)
    for order in (L.ORDER_ORIGINAL, L.ORDER_RANDOM, L.ORDER_SHUFFLED):
        for union_order in (L.UNION_SEQUENTIAL, L.UNION_STRATIFIED_RANDOM):
            _check(d, [o], DEFAULT, union_order, order, [21], [2], 0, 3, 1, 0, window=64, max_window=1 << 17)


def test_union_apply_chain_follows_the_oracle():
    """sfgpu_step_union with apply_winners over 25 steps: committed == fresh == oracle after every step."""
    _, d, oracles = _setup(n=40, routes=6, seed=8, R=2, coarse=False)
    desc = GpuScoreDirector.default_list_union()
    apply_fn = {0: "apply_list_change", 1: "apply_list_swap", 2: "apply_sublist_change", 3: "apply_sublist_swap",
                4: "apply_list_reverse"}
    for step in range(25):
        last = d.calculate_score()
        ref = np.concatenate([last, last], axis=1)
        seeds = [1000 + 7 * step, 2000 + step]
        steps = [step, step]
        idx, best, ev, win, flags = d.step_union(desc, ForageParams(2, 1, 32), step_seeds=seeds, step_indices=steps,
                                                 ref_scores=ref, apply=True)
        for r, o in enumerate(oracles):
            out, child, local, data, sc = oracle_lib.union_step(o, DEFAULT, L.UNION_STRATIFIED_RANDOM, L.ORDER_RANDOM, step,
                                                                seeds[r], last[r], last[r], 0, 32, True, 1)
            assert int(ev[r]) == out[2] and flags[r] == 0
            assert out[0] and int(idx[r]) == out[1]
            c, j = int(child[out[1]]), int(local[out[1]])
            getattr(o, apply_fn[DEFAULT[c][0]])(*[int(x) for x in data[c][0][j]])
        got = d.calculate_score()
        for r, o in enumerate(oracles):
            assert got[r].tolist() == o.committed_score().tolist() == d.fresh_score()[r].tolist()


def test_solve_union_tracks_the_step_calls():
    """The device-resident loop (sfgpu_solve_union) == the same loop driven call by call with the stated seeds."""
    from solverforge_b200.selectors import splitmix64
    c, d, _ = _setup(n=60, routes=6, seed=5, R=3, coarse=False)
    _, d2, _ = _setup(n=60, routes=6, seed=5, R=3, coarse=False)
    desc = GpuScoreDirector.default_list_union()
    n_steps, late = 40, 5
    best, ev, acc, ovf = d.solve_union(desc, n_steps, acceptor=2, late_size=late, tie_mode=1, accepted_limit=24, seed_base=99)
    R = 3
    hist = [[d2.calculate_score()[r].copy() for _ in range(late)] for r in range(R)]
    best2 = d2.calculate_score().copy()
    evs = np.zeros(R, dtype=np.int64)
    for t in range(n_steps):
        last = d2.calculate_score()
        ref = np.concatenate([last, np.stack([hist[r][t % late] for r in range(R)])], axis=1)
        seeds = [splitmix64(99 ^ ((r * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)) ^ t) for r in range(R)]
        idx, b, e, win, flags = d2.step_union(desc, ForageParams(2, 1, 24), step_seeds=seeds, step_indices=[t] * R,
                                              ref_scores=ref, apply=True)
        now = d2.calculate_score()
        for r in range(R):
            hist[r][t % late] = now[r].copy()
            if tuple(now[r]) > tuple(best2[r]):
                best2[r] = now[r]
        evs += e.astype(np.int64)
    assert ovf.tolist() == [0] * R
    assert best.tolist() == best2.tolist()
    assert ev.astype(np.int64).tolist() == evs.tolist()
    assert d.calculate_score().tolist() == d2.calculate_score().tolist()


def test_solve_union_default_policy_trajectory_cvrp100_200_steps():
    """The reference's default list local search (Random leaves, StratifiedRandom union, LateAcceptance(400),
    AcceptedCount(256)) for 200 steps on CVRP-100: the device-resident loop lands on the oracle's trajectory —
    same committed solution, best score and moves_evaluated — given the loop's stated step seeds."""
    from solverforge_b200.selectors import splitmix64
    c = instances.cvrp(100, 8, seed=77)
    start = instances.perturb_routes(c, 5, 60)
    d = models.cvrp_director(c, 1, offsets=start[0][None, :], elems=start[1])
    o = Oracle.cvrp(c, *start)
    desc = GpuScoreDirector.default_list_union()
    n_steps, late, limit, seed_base = 200, 400, 256, 4242
    best, ev, acc, ovf = d.solve_union(desc, n_steps, acceptor=2, late_size=late, tie_mode=1, accepted_limit=limit,
                                       seed_base=seed_base)
    apply_fn = {0: "apply_list_change", 1: "apply_list_swap", 2: "apply_sublist_change", 3: "apply_sublist_swap",
                4: "apply_list_reverse"}
    init = o.committed_score().copy()
    hist = [init.copy() for _ in range(late)]
    best_o, evaluated, committed = init.copy(), 0, 0
    for t in range(n_steps):
        last = o.committed_score().copy()
        seed = splitmix64(seed_base ^ 0 ^ t)
        out, child, local, data, sc = oracle_lib.union_step(o, DEFAULT, L.UNION_STRATIFIED_RANDOM, L.ORDER_RANDOM, t, seed, last,
                                                            hist[t % late], 0, limit, True, 1)
        evaluated += out[2]
        if out[0]:
            ci, j = int(child[out[1]]), int(local[out[1]])
            getattr(o, apply_fn[DEFAULT[ci][0]])(*[int(x) for x in data[ci][0][j]])
            committed += 1
        now = o.committed_score().copy()
        hist[t % late] = now
        if tuple(now) > tuple(best_o):
            best_o = now
    assert ovf.tolist() == [0]
    assert int(ev[0]) == evaluated and int(acc[0]) == committed
    assert best[0].tolist() == best_o.tolist()
    assert d.calculate_score()[0].tolist() == o.committed_score().tolist() == d.fresh_score()[0].tolist()


def test_solve_union_simulated_annealing_follows_the_oracle():
    """SimulatedAnnealing inside the union loop (windows re-run from the state at the start of the step)."""
    from solverforge_b200.selectors import splitmix64
    from tests.oracle_lib import OracleAcceptor
    c = instances.cvrp(40, 5, seed=12)
    start = instances.perturb_routes(c, 3, 30)
    d = models.cvrp_director(c, 1, offsets=start[0][None, :], elems=start[1])
    o = Oracle.cvrp(c, *start)
    desc = GpuScoreDirector.default_list_union(window=8)
    n_steps, limit, seed_base, samples = 40, 12, 99, 20
    best, ev, acc, ovf = d.solve_union(desc, n_steps, acceptor=6, late_size=samples, tie_mode=1, accepted_limit=limit,
                                       seed_base=seed_base, acceptor_real=0.9)
    apply_fn = {0: "apply_list_change", 1: "apply_list_swap", 2: "apply_sublist_change", 3: "apply_sublist_swap",
                4: "apply_list_reverse"}
    sa = OracleAcceptor(OracleAcceptor.SIMULATED_ANNEALING, size=samples, real=0.9)
    init = o.committed_score().copy()
    sa.phase_started(init)
    best_o, evaluated, committed = init.copy(), 0, 0
    for t in range(n_steps):
        last = o.committed_score().copy()
        seed = splitmix64(seed_base ^ 0 ^ t)
        kids = oracle_lib.union_children(o, DEFAULT, t, seed, L.ORDER_RANDOM)
        child, local = oracle_lib.union_pull_order([len(k[1]) for k in kids], L.UNION_STRATIFIED_RANDOM, t, seed, L.ORDER_RANDOM)
        sc = np.zeros((len(child), 2), dtype=np.int64)
        ok = np.zeros(len(child), dtype=np.uint8)
        for ci, k in enumerate(kids):
            sel = child == ci
            sc[sel] = k[2][local[sel]]
            ok[sel] = k[3][local[sel]]
        out = sa.step(sc, ok, best_o, last, seed, 0, limit, True)
        evaluated += out[2]
        if out[0]:
            ci, j = int(child[out[1]]), int(local[out[1]])
            getattr(o, apply_fn[DEFAULT[ci][0]])(*[int(x) for x in kids[ci][0][j]])
            committed += 1
        now = o.committed_score().copy()
        if tuple(now) > tuple(best_o):
            best_o = now
    assert ovf.tolist() == [0]
    assert int(ev[0]) == evaluated and int(acc[0]) == committed
    assert best[0].tolist() == best_o.tolist()
    assert d.calculate_score()[0].tolist() == o.committed_score().tolist() == d.fresh_score()[0].tolist()


SCALAR = [(L.FAM_CHANGE,), (L.FAM_SWAP,)]


def _gc(R=3, n=120, m=400, k=4):
    g = instances.graph_coloring(n, m, k, seed_edges=6, seed_colors=9, unassigned_permille=100)
    colors = np.stack([instances.graph_coloring(n, m, k, seed_edges=6, seed_colors=50 + r, unassigned_permille=100).color
                       for r in range(R)])
    return g, colors, models.graph_coloring_director(g, R, colors=colors), [Oracle.graph_coloring(g, colors[r]) for r in range(R)]


@pytest.mark.parametrize("order", [L.ORDER_ORIGINAL, L.ORDER_RANDOM, L.ORDER_SHUFFLED])
def test_scalar_change_and_swap_cursors_in_every_selection_order(order):
    """The Change / Swap families of a scalar model (move_selector/change.rs:246-307, swap.rs:196-233) walked on device."""
    g, colors, d, oracles = _gc()
    seeds, steps = [5, 77, 0xDEADBEEF], [0, 3, 900]
    for child in ([SCALAR[0]], [SCALAR[1]], SCALAR):
        for union_order in (L.UNION_SEQUENTIAL, L.UNION_STRATIFIED_RANDOM):
            for acceptor, okind, dl in ((0, 3, 0), (1, 0, -2), (2, 1, 0)):
                for ties, limit in ((0, 1), (1, 25), (1, 600)):
                    _check(d, oracles, child, union_order, order, seeds, steps, acceptor, okind, ties, limit, dl, max_window=1 << 14)


def test_default_scalar_search_on_device_follows_the_oracle():
    """The reference's default for a plain scalar model — union[Change, Swap] with Random leaves, StratifiedRandom,
    SimulatedAnnealing, AcceptedCount(1) (default_local_search/policy.rs:48-81) — as a device-resident loop."""
    from solverforge_b200.selectors import splitmix64
    from tests.oracle_lib import OracleAcceptor
    g, colors, d, oracles = _gc(R=2)
    desc = GpuScoreDirector.default_scalar_union(window=4)
    n_steps, seed_base, samples = 120, 31, 16
    best, ev, acc, ovf = d.solve_union(desc, n_steps, acceptor=6, late_size=samples, tie_mode=1, accepted_limit=1,
                                       seed_base=seed_base, acceptor_real=0.97)
    final, state = d.calculate_score(), d.scalar_state()
    for r, o in enumerate(oracles):
        sa = OracleAcceptor(OracleAcceptor.SIMULATED_ANNEALING, size=samples, real=0.97)
        init = o.committed_score().copy()
        sa.phase_started(init)
        best_o, evaluated, committed = init.copy(), 0, 0
        cur = colors[r].copy()
        for t in range(n_steps):
            last = o.committed_score().copy()
            seed = splitmix64(seed_base ^ ((r * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)) ^ t)
            kids = oracle_lib.union_children(o, SCALAR, t, seed, L.ORDER_RANDOM)
            child, local = oracle_lib.union_pull_order([len(k[1]) for k in kids], L.UNION_STRATIFIED_RANDOM, t, seed, L.ORDER_RANDOM,
                                                       limit=4000)
            sc = np.zeros((len(child), 2), dtype=np.int64)
            ok = np.zeros(len(child), dtype=np.uint8)
            for ci, k in enumerate(kids):
                sel = child == ci
                sc[sel] = k[2][local[sel]]
                ok[sel] = k[3][local[sel]]
            out = sa.step(sc, ok, best_o, last, seed, 0, 1, True)
            assert out[2] < 4000            # the forager quit inside the enumerated prefix
            evaluated += out[2]
            if out[0]:
                ci, j = int(child[out[1]]), int(local[out[1]])
                a, b = [int(x) for x in kids[ci][0][j]]
                if ci == 0:
                    o.apply_change(a, b)
                    cur[a] = b
                else:
                    o.apply_swap(a, b)
                    cur[a], cur[b] = cur[b], cur[a]
                committed += 1
            now = o.committed_score().copy()
            if tuple(now) > tuple(best_o):
                best_o = now
        assert int(ovf[r]) == 0
        assert int(ev[r]) == evaluated and int(acc[r]) == committed, f"r={r}"
        assert best[r].tolist() == best_o.tolist()
        assert final[r].tolist() == o.committed_score().tolist()
        assert np.array_equal(state[r], cur)
    assert np.array_equal(d.fresh_score(), final)


def test_windowed_nearby_loop_equals_the_whole_neighbourhood_loop():
    """sfgpu_solve_nearby_list_change with the windowed flag runs as a one-child union in SelectionOrder::Original
    (windowed speculation): same trajectory as the whole-neighbourhood loop."""
    c = instances.cvrp(120, 9, seed=44)
    R = 4
    starts = [instances.perturb_routes(c, 60 + r, 40) for r in range(R)]
    offs, el = np.stack([s[0] for s in starts]), np.concatenate([s[1] for s in starts])
    out = {}
    for mode in ("windowed", "whole"):
        d = models.cvrp_director(c, R, offsets=offs, elems=el)
        best, ev, acc = d.solve_nearby_list_change(80, 20, 2, 7, 1, 48, seed_base=321, windowed=mode == "windowed")
        out[mode] = (best.tolist(), ev.tolist(), acc.tolist(), d.calculate_score().tolist(), d.list_state()[1].tolist())
        assert np.array_equal(d.fresh_score(), d.calculate_score())
    assert out["windowed"] == out["whole"]
