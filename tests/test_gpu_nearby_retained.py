"""The retained nearby neighbourhood (nearby_step_cached_kernel): after a committed ListChange move
only the sources the move can have changed are regenerated or re-scored, the rest keep their score deltas. Every step
of a long committed trajectory must be bit-identical to the same step with the whole neighbourhood regenerated
(SFGPU_NO_NBCACHE=1) — index, best score, moves_evaluated, winner row — and to the oracle."""
import ctypes as C

import numpy as np
import pytest

from solverforge_b200 import ForageParams, instances, models
from solverforge_b200 import _lib as L
from tests import oracle_lib
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu


def _director(c, starts, monkeypatch, cached):
    if cached:
        monkeypatch.delenv("SFGPU_NO_NBCACHE", raising=False)
    else:
        monkeypatch.setenv("SFGPU_NO_NBCACHE", "1")
    d = models.cvrp_director(c, len(starts), offsets=np.stack([s[0] for s in starts]),
                             elems=np.concatenate([s[1] for s in starts]))
    monkeypatch.delenv("SFGPU_NO_NBCACHE", raising=False)
    return d


def _tags(d):
    out = np.zeros((d.R, 16), dtype=np.uint32)
    fn = d.lib.sfgpu_debug_nearby_cache_tags
    fn.argtypes = [C.c_void_p, C.c_void_p]
    fn.restype = C.c_int32
    assert fn(d.h, out.ctypes.data_as(C.c_void_p)) == 0
    return out


def _ref(last, acceptor, step):
    # LateAcceptance-like references: the late score trails the last one by a step-dependent margin
    late = last + np.array([0, -(37 * (step % 5))])
    return np.concatenate([last, late], axis=1)


@pytest.mark.parametrize("n,routes,coarse,K,acceptor,limit,steps", [
    (180, 11, 1, 20, 0, 0, 60),       # accept-all + BestScore: the best move of the whole neighbourhood, every step
    (180, 11, 40, 20, 2, 0, 60),      # coarse distances: ties at the k-th boundary and inside the kept lists
    (180, 70, 25, 20, 2, 37, 80),     # many short routes: append slots appear / vanish, routes empty out
    (120, 9, 1, 12, 1, 0, 40),        # hill climbing: converges, later steps commit nothing (cache stays current)
    (40, 3, 10, 32, 0, 0, 50),        # lists that hold every slot (count < K at times)
    (30, 1, 1, 20, 0, 5, 40),         # one route: every move is intra-route
    (400, 16, 1, 20, 2, 256, 50),
])
def test_retained_neighbourhood_equals_full_regeneration(monkeypatch, n, routes, coarse, K, acceptor, limit, steps):
    c = instances.cvrp(n, routes, seed=21)
    c.matrix = (c.matrix // coarse) * coarse
    R = 4
    starts = [instances.perturb_routes(c, 400 + r, n // 3) for r in range(R)]
    d_c = _director(c, starts, monkeypatch, True)
    d_f = _director(c, starts, monkeypatch, False)
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(2)]
    fp = ForageParams(acceptor, 1, limit)
    saw_moved = saw_current = False
    for step in range(steps):
        last = d_c.calculate_score()
        assert np.array_equal(last, d_f.calculate_score())
        ref = _ref(last, acceptor, step)
        seeds = [1000 * step + r for r in range(R)]
        got = d_c.step_nearby_list_change(K, fp, step_seeds=seeds, ref_scores=ref, apply=True)
        want = d_f.step_nearby_list_change(K, fp, step_seeds=seeds, ref_scores=ref, apply=True)
        for g, w, what in zip(got, want, ("index", "best", "moves_evaluated", "winner rows")):
            assert np.array_equal(g, w), f"step {step}: {what} differ from the fully regenerated step"
        t = _tags(d_c)
        for r in range(R):
            moved = got[0][r] != 0xFFFFFFFF
            assert t[r, 0] == (2 if moved else 1) and t[r, 1] == K, f"step {step} replica {r}: protocol state {t[r, :3]}"
            saw_moved |= bool(moved)
            saw_current |= not moved
        if step % 7 == 0:   # the oracle on the same trajectory (replicas 0, 1)
            pass
        for r in range(2):
            rows = oracles[r].enumerate_nearby_list_change(K)
            so, oko = oracles[r].score_list_change(rows)
            okind = {0: 3, 1: 0, 2: 1}[acceptor]
            out = oracle_lib.replay_step(so, oko, [0, 0], ref[r][:2], ref[r][2:], seeds[r], 0 if limit else 2, max(limit, 1),
                                         True, okind)
            if out[0]:
                assert int(got[0][r]) == out[1] and got[3][r].tolist() == rows[out[1]].tolist(), f"step {step} replica {r} vs oracle"
                oracles[r].apply_list_change(*rows[out[1]])
            else:
                assert got[0][r] == 0xFFFFFFFF
            assert int(got[2][r]) == out[2]
    assert saw_moved
    assert np.array_equal(d_c.fresh_score(), d_c.calculate_score())
    assert _tags(d_f)[:, 0].max() == 0       # the regenerating context never marks a cache current


def test_other_writers_invalidate_the_retained_neighbourhood(monkeypatch):
    """Anything but one ListChange commit right after a cached step — another move kind, two commits in a row, a
    materialising step in between — must not leave stale deltas behind."""
    c = instances.cvrp(150, 8, seed=5)
    R, K = 3, 20
    starts = [instances.perturb_routes(c, 90 + r, 50) for r in range(R)]
    d_c = _director(c, starts, monkeypatch, True)
    d_f = _director(c, starts, monkeypatch, False)
    fp = ForageParams(0, 1, 0)
    rng = np.random.default_rng(3)

    def both(fn):
        return fn(d_c), fn(d_f)

    def step(apply=True, seed=0):
        got, want = both(lambda d: d.step_nearby_list_change(K, fp, step_seeds=[seed + r for r in range(R)], apply=apply))
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
        return got

    for it in range(30):
        got = step(True, 10 * it)
        kind = it % 5
        offs, _ = d_c.list_state()
        lens = np.diff(offs, axis=1)
        if kind == 0:      # a swap commit
            rows = []
            for r in range(R):
                es = [e for e in range(lens.shape[1]) if lens[r, e] > 0]
                a, b = rng.choice(es), rng.choice(es)
                rows.append([a, rng.integers(lens[r, a]), b, rng.integers(lens[r, b])])
            both(lambda d: d.apply_list_swap(np.array(rows, dtype=np.uint32)))
            assert (_tags(d_c)[:, 0] == 0).all()
        elif kind == 1:    # a second ListChange commit without a step in between
            got2 = step(False, 10 * it + 5)
            both(lambda d: d.apply_list_change(got2[3]))
            both(lambda d: d.apply_list_change(np.array([[0, 0, 0, 0]] * R, dtype=np.uint32)))  # not doable (no-op): ignored
        elif kind == 2:    # a reversal
            rows = []
            for r in range(R):
                e = int(np.argmax(lens[r]))
                rows.append([e, 0, lens[r, e], 0])
            both(lambda d: d.apply_list_reverse(np.array(rows, dtype=np.uint32)))
            assert (_tags(d_c)[:, 0] == 0).all()
        elif kind == 3:    # a masked ListChange commit: only replica 1 moves
            got2 = step(False, 10 * it + 7)
            mask = np.array([0, 1, 0], dtype=np.uint8)
            both(lambda d: d.apply_list_change(got2[3], mask=mask))
        assert np.array_equal(d_c.calculate_score(), d_f.calculate_score())
    step(True, 999)


def test_device_loop_with_the_retained_neighbourhood(monkeypatch):
    """sfgpu_solve_nearby_list_change (captured step graph) on a cached and on a regenerating context: same best scores,
    moves_evaluated and accepted steps; restore_best invalidates."""
    c = instances.cvrp(200, 10, seed=8)
    R = 6
    starts = [instances.perturb_routes(c, 300 + r, 80) for r in range(R)]
    d_c = _director(c, starts, monkeypatch, True)
    d_f = _director(c, starts, monkeypatch, False)
    for acceptor, late, limit, restore in ((2, 7, 48, False), (1, 0, 0, True), (2, 50, 0, True)):
        a = d_c.solve_nearby_list_change(70, 20, acceptor, late, 1, limit, seed_base=17, restore_best=restore)
        b = d_f.solve_nearby_list_change(70, 20, acceptor, late, 1, limit, seed_base=17, restore_best=restore)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert np.array_equal(d_c.calculate_score(), d_f.calculate_score())
        if restore:
            assert (_tags(d_c)[:, 0] == 0).all()
        # and a host-driven step right behind the loop
        got = d_c.step_nearby_list_change(20, ForageParams(0, 1, 0), step_seeds=list(range(R)), apply=True)
        want = d_f.step_nearby_list_change(20, ForageParams(0, 1, 0), step_seeds=list(range(R)), apply=True)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)


def test_retained_neighbourhood_fuzz(monkeypatch):
    """Random shapes (customers, routes, max_nearby, distance granularity, acceptor, AcceptedCount): every committed
    step of the cached context equals the regenerating one."""
    rng = np.random.default_rng(20261017)
    for case in range(12):
        n = int(rng.integers(12, 260))
        routes = int(rng.integers(1, max(2, n // 3)))
        K = int(rng.integers(1, 33))
        coarse = int(rng.choice([1, 1, 7, 60]))
        acceptor = int(rng.choice([0, 1, 2]))
        limit = int(rng.choice([0, 0, 3, 200]))
        c = instances.cvrp(n, routes, seed=100 + case)
        c.matrix = (c.matrix // coarse) * coarse
        R = 3
        starts = [instances.perturb_routes(c, 900 + 7 * case + r, n // 2) for r in range(R)]
        d_c = _director(c, starts, monkeypatch, True)
        d_f = _director(c, starts, monkeypatch, False)
        fp = ForageParams(acceptor, int(rng.integers(0, 2)), limit)
        for step in range(30):
            last = d_c.calculate_score()
            ref = _ref(last, acceptor, step)
            seeds = [31 * step + r for r in range(R)]
            got = d_c.step_nearby_list_change(K, fp, step_seeds=seeds, ref_scores=ref, apply=True)
            want = d_f.step_nearby_list_change(K, fp, step_seeds=seeds, ref_scores=ref, apply=True)
            for g, w, what in zip(got, want, ("index", "best", "moves_evaluated", "winner rows")):
                assert np.array_equal(g, w), f"case {case} (n={n} routes={routes} K={K} coarse={coarse} acc={acceptor} " \
                                             f"limit={limit}) step {step}: {what}"
        assert np.array_equal(d_c.fresh_score(), d_f.calculate_score())
        d_c.close()
        d_f.close()


def test_retained_neighbourhood_full_size(monkeypatch):
    """CVRP-1000 / 80 routes (the bench workload), 8 replicas, 60 committed steps under accept-all and under
    LateAcceptance-like references with AcceptedCount(256)."""
    c = instances.cvrp()
    R, K = 8, 20
    starts = [instances.perturb_routes(c, 5000 + r, 64) for r in range(R)]
    for acceptor, limit in ((0, 0), (2, 256)):
        d_c = _director(c, starts, monkeypatch, True)
        d_f = _director(c, starts, monkeypatch, False)
        fp = ForageParams(acceptor, 1, limit)
        for step in range(60):
            last = d_c.calculate_score()
            ref = _ref(last, acceptor, step)
            seeds = [step * 1000 + r for r in range(R)]
            got = d_c.step_nearby_list_change(K, fp, step_seeds=seeds, ref_scores=ref, apply=True)
            want = d_f.step_nearby_list_change(K, fp, step_seeds=seeds, ref_scores=ref, apply=True)
            for g, w in zip(got, want):
                assert np.array_equal(g, w), f"acceptor {acceptor} step {step}"
        t = _tags(d_c)[:, 10:13].astype(np.float64).sum(axis=0)
        assert t[2] / t.sum() < 0.15, "most sources keep or re-score their deltas at full size"
        assert np.array_equal(d_c.fresh_score(), d_f.calculate_score())
        d_c.close()
        d_f.close()
