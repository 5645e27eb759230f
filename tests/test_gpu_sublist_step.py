"""GPU parity of the device-enumerated SublistChange step (sfgpu_step_sublist_change, sfgpu_index_step.cuh) against
the oracle's SublistChangeMoveSelector order (list_kernel/sublist_change.rs:103-268), scores and candidate-loop
replay: pull index of the winner, moves_evaluated, best score, the winning move and the committed state. Bit-exact."""
import numpy as np
import pytest

from solverforge_b200 import ForageParams, instances, models
from tests import oracle_lib
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu


def _check_step(d, oracles, sizes, seeds, acceptor, okind, ties, limit, dl=0, swap=False):
    lo, hi = sizes
    base = d.calculate_score()
    ref = np.concatenate([base + [0, dl], base + [0, dl - 7]], axis=1)
    step = d.step_sublist_swap if swap else d.step_sublist_change
    idx, best, ev, win = step(lo, hi, ForageParams(acceptor, ties, limit), step_seeds=seeds, ref_scores=ref)
    for r, o in enumerate(oracles):
        rows = o.enumerate_sublist_swap(lo, hi) if swap else o.enumerate_sublist_change(lo, hi)
        if len(rows):
            so, oko = o.score_sublist_swap(rows) if swap else o.score_sublist_change(rows)
        else:
            so, oko = np.zeros((0, 2), np.int64), np.zeros(0, np.uint8)
        out = oracle_lib.replay_step(so, oko, [0, 0], ref[r][:2], ref[r][2:], seeds[r], 0 if limit else 2, max(limit, 1),
                                     bool(ties), okind)
        what = f"replica={r} swap={swap} sizes={sizes} acc={acceptor} ties={ties} limit={limit} dl={dl}"
        assert int(ev[r]) == out[2], what + " moves_evaluated"
        if out[0]:
            assert int(idx[r]) == out[1], what
            assert best[r].tolist() == so[out[1]].tolist(), what
            assert win[r].tolist() == rows[out[1]].astype(np.int64).tolist(), what
        else:
            assert idx[r] == 0xFFFFFFFF and win[r].tolist() == [-1] * (6 if swap else 5), what


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("sizes", [(1, 3), (2, 2), (3, 6)])
def test_sublist_change_step_matches_oracle(sizes, generic):
    """generic = SFGPU_CTX_GENERIC_KERNELS: the cursor walks with the constraint-table delta instead of the
    CVRP-shaped one with hoisted source invariants."""
    from solverforge_b200 import _lib as L
    c = instances.cvrp(46, 7, seed=31)
    c.matrix = (c.matrix // 40) * 40     # coarse distances: many equal scores, the tie rule matters
    R = 3
    starts = [instances.perturb_routes(c, 40 + r, 30 + 10 * r) for r in range(R)]
    offs = np.stack([s[0] for s in starts])
    el = np.concatenate([s[1] for s in starts])
    d = models.cvrp_director(c, R, offsets=offs, elems=el, flags=L.CTX_GENERIC_KERNELS if generic else 0)
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    for r in range(R):
        assert d.calculate_score()[r].tolist() == oracles[r].committed_score().tolist()
    seeds = [5, 77, 0xDEADBEEF]
    for acceptor, okind in ((0, 3), (1, 0), (2, 1)):
        for ties in (0, 1):
            for limit in ((0, 17) if generic else (0, 1, 17, 900, 10 ** 7)):
                _check_step(d, oracles, sizes, seeds, acceptor, okind, ties, limit, dl=0 if acceptor != 1 else -30)


@pytest.mark.parametrize("sizes", [(1, 3), (2, 2), (2, 5)])
def test_sublist_swap_step_matches_oracle(sizes):
    """sfgpu_step_sublist_swap vs SublistSwapMoveSelector (list_kernel/sublist_swap.rs:28-318) + candidate loop."""
    c = instances.cvrp(34, 6, seed=35)
    c.matrix = (c.matrix // 40) * 40
    R = 3
    starts = [instances.perturb_routes(c, 50 + r, 20 + 15 * r) for r in range(R)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    seeds = [6, 78, 0xFEEDF00D]
    for acceptor, okind in ((0, 3), (1, 0), (2, 1)):
        for ties in (0, 1):
            for limit in (0, 1, 23, 1500, 10 ** 7):
                _check_step(d, oracles, sizes, seeds, acceptor, okind, ties, limit, dl=0 if acceptor != 1 else -30, swap=True)


def test_sublist_swap_step_apply_chain_and_degenerate_routes():
    c = instances.cvrp(26, 5, seed=36)
    offs, el = instances.perturb_routes(c, 4, 20)
    lists = [el[offs[i]:offs[i + 1]].tolist() for i in range(5)]
    lists[0] += lists[2]
    lists[2] = []                         # an empty route in the middle
    while len(lists[4]) > 1:              # the last route keeps one element
        lists[1].append(lists[4].pop())
    offs = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.uint32)
    el = np.array([x for l in lists for x in l], dtype=np.uint32)
    d = models.cvrp_director(c, 1, offsets=offs[None, :], elems=el)
    o = Oracle.cvrp(c, offs, el)
    for step in range(6):
        last = d.calculate_score()
        ref = np.concatenate([last, last], axis=1)
        rows = o.enumerate_sublist_swap(1, 3)
        so, oko = o.score_sublist_swap(rows)
        idx, best, ev, win = d.step_sublist_swap(1, 3, ForageParams(1, 1, 0), step_seeds=[500 + step], ref_scores=ref,
                                                 apply=True)
        out = oracle_lib.replay_step(so, oko, [0, 0], last[0], last[0], 500 + step, 2, 1, True, 0)
        assert int(ev[0]) == out[2] == len(rows)
        if not out[0]:
            assert idx[0] == 0xFFFFFFFF
            break
        assert int(idx[0]) == out[1] and win[0].tolist() == rows[out[1]].astype(np.int64).tolist()
        o.apply_sublist_swap(*rows[out[1]])
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist() == d.fresh_score()[0].tolist()
    one = instances.cvrp(12, 1, seed=2)
    d1 = models.cvrp_director(one)
    o1 = Oracle.cvrp(one)
    _check_step(d1, [o1], (1, 3), [9], 0, 3, 1, 0, swap=True)
    _check_step(d1, [o1], (2, 4), [9], 0, 3, 1, 7, swap=True)
    idx, best, ev, win = d1.step_sublist_swap(20, 25, ForageParams(0, 1, 0), step_seeds=[1])
    assert idx[0] == 0xFFFFFFFF and int(ev[0]) == 0


def test_sublist_swap_step_full_size_properties():
    """CVRP-1000 / 80: ~4.4 M segment pairs per replica, never materialised; moves_evaluated equals the closed-form
    pair count and the winner re-scores to the reported best through the rows-resident call."""
    c = instances.cvrp()
    R = 2
    starts = [instances.perturb_routes(c, 80 + r, 200) for r in range(R)]
    offs = np.stack([s[0] for s in starts])
    d = models.cvrp_director(c, R, offsets=offs, elems=np.concatenate([s[1] for s in starts]))
    idx, best, ev, win = d.step_sublist_swap(1, 3, ForageParams(0, 1, 0), step_seeds=[3, 4])
    for r in range(R):
        lens = np.diff(offs[r]).astype(np.int64)
        segs = [[(s, s + z) for s in range(ln) for z in range(1, min(3, ln - s) + 1)] for ln in lens]
        nseg = np.array([len(x) for x in segs])
        later = nseg.sum() - np.cumsum(nseg)
        want = 0
        for e, sg in enumerate(segs):
            starts_e = np.array([a for a, _ in sg])
            for (a, b) in sg:
                want += int((starts_e >= b).sum()) + int(later[e])
        assert int(ev[r]) == want
    s, ok = d.score_sublist_swap(win, np.arange(R + 1, dtype=np.uint64))
    assert ok.tolist() == [1] * R and np.array_equal(s, best)


def test_sublist_change_step_apply_chain_and_degenerate_routes():
    """A single route, an empty route and a route shorter than the minimum segment; winners committed on device."""
    c = instances.cvrp(30, 5, seed=33)
    offs, el = instances.perturb_routes(c, 3, 20)
    # move everything of route 4 into route 0 (route 4 empty) by rebuilding the lists
    lens = np.diff(offs).tolist()
    lists = [el[offs[i]:offs[i + 1]].tolist() for i in range(5)]
    lists[0] += lists[4]
    lists[4] = []
    while len(lists[3]) > 1:              # route 3 keeps one element
        lists[1].append(lists[3].pop())
    offs = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.uint32)
    el = np.array([x for l in lists for x in l], dtype=np.uint32)
    d = models.cvrp_director(c, 1, offsets=offs[None, :], elems=el)
    o = Oracle.cvrp(c, offs, el)
    for step in range(6):
        last = d.calculate_score()
        ref = np.concatenate([last, last], axis=1)
        rows = o.enumerate_sublist_change(2, 3)
        so, oko = o.score_sublist_change(rows)
        idx, best, ev, win = d.step_sublist_change(2, 3, ForageParams(1, 1, 0), step_seeds=[400 + step], ref_scores=ref,
                                                   apply=True)
        out = oracle_lib.replay_step(so, oko, [0, 0], last[0], last[0], 400 + step, 2, 1, True, 0)
        assert int(ev[0]) == out[2] == len(rows)
        if not out[0]:
            assert idx[0] == 0xFFFFFFFF
            break
        assert int(idx[0]) == out[1] and win[0].tolist() == rows[out[1]].astype(np.int64).tolist()
        o.apply_sublist_change(*rows[out[1]])
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist() == d.fresh_score()[0].tolist()
    # one route only: intra-list candidates only
    one = instances.cvrp(12, 1, seed=2)
    d1 = models.cvrp_director(one)
    o1 = Oracle.cvrp(one)
    _check_step(d1, [o1], (1, 3), [9], 0, 3, 1, 0)
    _check_step(d1, [o1], (1, 3), [9], 0, 3, 1, 5)
    # segments longer than every route: empty neighbourhood, a step without a winner
    idx, best, ev, win = d1.step_sublist_change(20, 25, ForageParams(0, 1, 0), step_seeds=[1])
    assert idx[0] == 0xFFFFFFFF and int(ev[0]) == 0


def test_sublist_change_step_full_size_properties():
    """CVRP-1000 / 80 (BASELINE C3 shape): ~3.2 M candidates per replica, never materialised. The winner's score
    must equal the score of that move through the rows-resident call, no candidate may beat it (sampled), and
    moves_evaluated equals the closed-form neighbourhood size."""
    c = instances.cvrp()
    R = 2
    starts = [instances.perturb_routes(c, 60 + r, 200) for r in range(R)]
    offs = np.stack([s[0] for s in starts])
    el = np.concatenate([s[1] for s in starts])
    d = models.cvrp_director(c, R, offsets=offs, elems=el)
    idx, best, ev, win = d.step_sublist_change(1, 3, ForageParams(0, 1, 0), step_seeds=[3, 4])
    for r in range(R):
        lens = np.diff(offs[r]).astype(np.int64)
        T = int(lens.sum()) + len(lens) - 1
        want = sum(int(T - size) for ln in lens for s in range(ln) for size in range(1, min(3, ln - s) + 1))
        assert int(ev[r]) == want
    offs_c = np.arange(R + 1, dtype=np.uint64)
    s, ok = d.score_sublist_change(win, offs_c)
    assert ok.tolist() == [1] * R and np.array_equal(s, best)
    # a random sample of replica 0's neighbourhood through the rows-resident call: nothing beats the winner
    lens = np.diff(offs[0]).astype(np.int64)
    rnd = instances.splitmix64_stream(123, 5 * 30000).reshape(-1, 5)
    se = (rnd[:, 0] % np.uint64(len(lens))).astype(np.int64)
    keep = lens[se] > 0
    se, rnd = se[keep], rnd[keep]
    start = (rnd[:, 1] % lens[se].astype(np.uint64)).astype(np.int64)
    size = np.minimum((rnd[:, 2] % np.uint64(3)).astype(np.int64) + 1, lens[se] - start)
    de = (rnd[:, 3] % np.uint64(len(lens))).astype(np.int64)
    room = np.where(de == se, lens[se] - size, lens[de])
    dp = (rnd[:, 4] % (room + 1).astype(np.uint64)).astype(np.int64)
    rows0 = np.stack([se, start, start + size, de, dp], axis=1)
    d0 = models.cvrp_director(c, 1, offsets=offs[0][None, :], elems=starts[0][1])
    s0, ok0 = d0.score_sublist_change(rows0)
    assert ok0.sum() > 20000
    s0 = s0[ok0 == 1]
    better = (s0[:, 0] > best[0][0]) | ((s0[:, 0] == best[0][0]) & (s0[:, 1] > best[0][1]))
    assert not better.any()


def test_list_reverse_step_matches_oracle():
    """sfgpu_step_list_reverse vs ListReverseMoveSelector (list_kernel/reverse.rs:66-108) + candidate loop, symmetric
    and asymmetric costs, empty / one-element routes, apply chain."""
    c = instances.cvrp(40, 6, seed=37)
    r = instances.splitmix64_stream(8, c.dim * c.dim).reshape(c.dim, c.dim)
    c.matrix = (c.matrix // 30) * 30 + (r % np.uint64(3)).astype(np.int64)   # asymmetric, many ties
    np.fill_diagonal(c.matrix, 0)
    R = 3
    starts = [instances.perturb_routes(c, 90 + q, 25 + 10 * q) for q in range(R)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    oracles = [Oracle.cvrp(c, *starts[q]) for q in range(R)]
    seeds = [4, 99, 0xABCDEF]
    for acceptor, okind in ((0, 3), (1, 0), (2, 1)):
        for ties in (0, 1):
            for limit in (0, 1, 9, 10 ** 6):
                base = d.calculate_score()
                dl = -20 if acceptor == 1 else 0
                ref = np.concatenate([base + [0, dl], base + [0, dl - 7]], axis=1)
                idx, best, ev, win = d.step_list_reverse(ForageParams(acceptor, ties, limit), step_seeds=seeds, ref_scores=ref)
                for q, o in enumerate(oracles):
                    rows = o.enumerate_list_reverse()
                    so, oko = o.score_list_reverse(rows)
                    out = oracle_lib.replay_step(so, oko, [0, 0], ref[q][:2], ref[q][2:], seeds[q], 0 if limit else 2,
                                                 max(limit, 1), bool(ties), okind)
                    what = f"replica={q} acc={acceptor} ties={ties} limit={limit}"
                    assert int(ev[q]) == out[2], what
                    if out[0]:
                        assert int(idx[q]) == out[1] and best[q].tolist() == so[out[1]].tolist(), what
                        assert win[q].tolist() == rows[out[1]][:3].astype(np.int64).tolist(), what
                    else:
                        assert idx[q] == 0xFFFFFFFF, what
    for step in range(5):   # committed winners
        last = d.calculate_score()
        idx, best, ev, win = d.step_list_reverse(ForageParams(1, 1, 0), step_seeds=[600 + step] * R,
                                                 ref_scores=np.concatenate([last, last], axis=1), apply=True)
        for q, o in enumerate(oracles):
            rows = o.enumerate_list_reverse()
            so, oko = o.score_list_reverse(rows)
            out = oracle_lib.replay_step(so, oko, [0, 0], last[q], last[q], 600 + step, 2, 1, True, 0)
            if out[0]:
                assert int(idx[q]) == out[1]
                o.apply_list_reverse(*rows[out[1]])
            else:
                assert idx[q] == 0xFFFFFFFF
            assert d.calculate_score()[q].tolist() == o.committed_score().tolist() == d.fresh_score()[q].tolist()
    one = instances.cvrp(3, 3, seed=1)   # every route holds one element: empty neighbourhood
    d1 = models.cvrp_director(one)
    idx, best, ev, win = d1.step_list_reverse(ForageParams(0, 1, 0), step_seeds=[1])
    assert idx[0] == 0xFFFFFFFF and int(ev[0]) == 0
