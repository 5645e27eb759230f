"""Seeded fuzzing of the device paths against the oracle: random instance shapes (route counts from a
single route to more routes than a batch holds, empty routes, coarse distances with many ties, max_nearby
1..32), random forager / acceptor settings. Everything bit-exact, through the C ABI."""
import os

import numpy as np
import pytest

from solverforge_b200 import ForageParams, instances, models
from solverforge_b200.instances import splitmix64_stream
from tests import oracle_lib
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu
FUZZ = int(os.environ.get("SFGPU_FUZZ_CASES", "24"))   # widen for a deep run: SFGPU_FUZZ_CASES=400


def _cfg(seed):
    s = [int(x) for x in splitmix64_stream(seed, 12)]
    n = 12 + s[0] % 150
    routes = 1 + s[1] % min(n, 60)
    coarse = [1, 1, 7, 40, 200][s[2] % 5]
    K = 1 + s[3] % 32
    perturb = s[4] % 80
    acceptor = s[5] % 4
    limit = [0, 0, 1, 5, 60, 100000][s[6] % 6]
    ties = s[7] % 2
    return n, routes, coarse, K, perturb, acceptor, limit, ties, s[8]


@pytest.mark.parametrize("seed", range(FUZZ))
def test_fuzz_nearby_change_and_swap_steps(seed):
    n, routes, coarse, K, perturb, acceptor, limit, ties, step_seed = _cfg(1000 + seed)
    c = instances.cvrp(n, routes, seed=seed + 3)
    c.matrix = (c.matrix // coarse) * coarse
    start = instances.perturb_routes(c, seed, perturb)
    d = models.cvrp_director(c, 1, offsets=start[0][None, :], elems=start[1])
    o = Oracle.cvrp(c, *start)
    base = d.calculate_score()[0]
    assert base.tolist() == o.committed_score().tolist()
    okind = {0: 3, 1: 0, 2: 1}.get(acceptor)
    ref = np.concatenate([base + [0, -(step_seed % 50)], base + [0, -(step_seed % 170)]])
    what = f"n={n} routes={routes} coarse={coarse} K={K} acc={acceptor} limit={limit} ties={ties}"
    for swap in (False, True):
        rows = o.enumerate_nearby_list_swap(K) if swap else o.enumerate_nearby_list_change(K)
        if len(rows):
            so, oko = (o.score_list_swap if swap else o.score_list_change)(rows)
        else:
            so, oko = np.zeros((0, 2), np.int64), np.zeros(0, np.uint8)
        step = d.step_nearby_list_swap if swap else d.step_nearby_list_change
        idx, best, ev, win = step(K, ForageParams(acceptor, ties, limit), step_seeds=[step_seed], ref_scores=[ref])
        if okind is not None:
            out = oracle_lib.replay_step(so, oko, [0, 0], ref[:2], ref[2:], step_seed, 0 if limit else 2, max(limit, 1),
                                         bool(ties), okind)
        else:  # predicate 3: score > last || score >= threshold (GreatDeluge form)
            acc = oracle_lib.OracleAcceptor(oracle_lib.OracleAcceptor.GREAT_DELUGE, real=0.0)
            acc.phase_started(ref[2:])
            out = acc.step(so, oko, base, ref[:2], step_seed, 0 if limit else 2, max(limit, 1), bool(ties))
        assert int(ev[0]) == out[2], what + f" swap={swap} moves_evaluated"
        if out[0]:
            assert int(idx[0]) == out[1], what + f" swap={swap}"
            assert best[0].tolist() == so[out[1]].tolist(), what
            assert win[0].tolist() == rows[out[1]].tolist(), what
        else:
            assert idx[0] == 0xFFFFFFFF, what
    # rows-resident fused step over the relocation batch
    rows = o.enumerate_nearby_list_change(K)
    if len(rows) and okind is not None:
        so, oko = o.score_list_change(rows)
        s, ok = d.score_list_change(rows)
        assert np.array_equal(s, so) and np.array_equal(ok, oko), what


@pytest.mark.parametrize("seed", range(max(10, FUZZ // 3)))
def test_fuzz_scalar_models(seed):
    s = [int(x) for x in splitmix64_stream(500 + seed, 8)]
    which = s[0] % 4
    if which == 3:   # the shipped shift-scheduling example (keyed pairs, complemented count, consecutive runs)
        sh = instances.shift_scheduling(n_days=3 + s[1] % 14, slots_per_day=1 + s[2] % 4, n_nurses=2 + s[3] % 6, seed=seed,
                                        unassigned_permille=s[4] % 400)
        o, d = Oracle.shift_scheduling(sh, with_load_balance=False), models.shift_scheduling_director(sh, with_load_balance=False)
    elif which == 0:
        n = 20 + s[1] % 300
        g = instances.graph_coloring(n, min(n * (1 + s[2] % 6), n * (n - 1) // 2), 2 + s[3] % 9, seed_edges=seed,
                                     seed_colors=seed + 1, unassigned_permille=s[4] % 300)
        o, d = Oracle.graph_coloring(g), models.graph_coloring_director(g)
    elif which == 1:
        j = instances.job_shop(5 + s[1] % 30, 2 + s[2] % 9, 2 + s[3] % 7, seed=seed, unassigned_permille=s[4] % 200)
        o, d = Oracle.job_shop(j), models.job_shop_director(j)
    else:
        r = instances.roster(10 + s[1] % 150, 2 + s[2] % 7, 2 + s[3] % 9, seed=seed, limit=4 + s[4] % 9)
        o, d = Oracle.roster(r), models.roster_director(r)
    assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
    for step in range(6):
        rows = o.enumerate_change()
        so, oko = o.score_change(rows)
        sg, okg = d.score_change(rows)
        assert np.array_equal(okg, oko) and np.array_equal(sg, so), f"which={which} seed={seed} step={step}"
        last = d.calculate_score()
        idx, best, ev, win = d.step_change(ForageParams(1 + step % 2, 1, [0, 7][step % 2]), step_seeds=[s[5] + step],
                                           ref_scores=np.concatenate([last, last], axis=1), apply=True)
        out = oracle_lib.replay_step(so, oko, [0, 0], last[0], last[0], s[5] + step, [2, 0][step % 2], [1, 7][step % 2],
                                     True, step % 2)
        if out[0]:
            assert int(idx[0]) == out[1], f"which={which} seed={seed} step={step}"
            o.apply_change(*rows[out[1]])
        else:
            assert idx[0] == 0xFFFFFFFF
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
        assert d.fresh_score()[0].tolist() == o.committed_score().tolist()


@pytest.mark.parametrize("seed", range(max(FUZZ // 2, 6)))
def test_fuzz_sublist_steps(seed):
    """Device-enumerated SublistChange / SublistSwap steps on random shapes: one to many routes, empty and
    one-element routes, coarse distances (ties), random segment sizes, acceptors, limits and tie modes."""
    s = [int(x) for x in splitmix64_stream(5000 + seed, 12)]
    n = 8 + s[0] % 40
    routes = 1 + s[1] % min(n, 9)
    coarse = [1, 7, 40, 200][s[2] % 4]
    lo = 1 + s[3] % 3
    hi = lo + s[4] % 4
    acceptor = s[5] % 3
    limit = [0, 0, 1, 5, 60, 100000][s[6] % 6]
    ties = s[7] % 2
    step_seed = s[8]
    c = instances.cvrp(n, routes, seed=seed + 11)
    c.matrix = (c.matrix // coarse) * coarse
    start = instances.perturb_routes(c, seed, s[9] % 60)
    d = models.cvrp_director(c, 1, offsets=start[0][None, :], elems=start[1])
    o = Oracle.cvrp(c, *start)
    base = d.calculate_score()[0]
    okind = {0: 3, 1: 0, 2: 1}[acceptor]
    ref = np.concatenate([base + [0, -(step_seed % 50)], base + [0, -(step_seed % 170)]])
    what = f"n={n} routes={routes} coarse={coarse} sizes={lo}..{hi} acc={acceptor} limit={limit} ties={ties}"
    for swap in (False, True):
        rows = o.enumerate_sublist_swap(lo, hi) if swap else o.enumerate_sublist_change(lo, hi)
        if len(rows):
            so, oko = (o.score_sublist_swap if swap else o.score_sublist_change)(rows)
        else:
            so, oko = np.zeros((0, 2), np.int64), np.zeros(0, np.uint8)
        step = d.step_sublist_swap if swap else d.step_sublist_change
        idx, best, ev, win = step(lo, hi, ForageParams(acceptor, ties, limit), step_seeds=[step_seed], ref_scores=[ref])
        out = oracle_lib.replay_step(so, oko, [0, 0], ref[:2], ref[2:], step_seed, 0 if limit else 2, max(limit, 1),
                                     bool(ties), okind)
        assert int(ev[0]) == out[2], what + f" swap={swap} moves_evaluated"
        if out[0]:
            assert int(idx[0]) == out[1], what + f" swap={swap}"
            assert best[0].tolist() == so[out[1]].tolist(), what + f" swap={swap}"
            assert win[0].tolist() == rows[out[1]].astype(np.int64).tolist(), what + f" swap={swap}"
        else:
            assert idx[0] == 0xFFFFFFFF, what + f" swap={swap}"
    # the ListReverse neighbourhood of the same state
    rows = o.enumerate_list_reverse()
    if len(rows):
        so, oko = o.score_list_reverse(rows)
    else:
        so, oko = np.zeros((0, 2), np.int64), np.zeros(0, np.uint8)
    idx, best, ev, win = d.step_list_reverse(ForageParams(acceptor, ties, limit), step_seeds=[step_seed], ref_scores=[ref])
    out = oracle_lib.replay_step(so, oko, [0, 0], ref[:2], ref[2:], step_seed, 0 if limit else 2, max(limit, 1), bool(ties),
                                 okind)
    assert int(ev[0]) == out[2], what + " reverse moves_evaluated"
    if out[0]:
        assert int(idx[0]) == out[1] and best[0].tolist() == so[out[1]].tolist(), what + " reverse"
        assert win[0].tolist() == rows[out[1]][:3].astype(np.int64).tolist(), what + " reverse"
    else:
        assert idx[0] == 0xFFFFFFFF, what + " reverse"
