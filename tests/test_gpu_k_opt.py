"""GPU parity of the intra-list KOptMove (heuristic/move/k_opt.rs:11-110, reconnection patterns
k_opt_reconnection.rs:63-262) and of the device KOptCursor (list_kernel/k_opt/full.rs:34-98, SFGPU_FAM_K_OPT): every
move of the oracle's selector scored on device, the cursor in every selection order, winners applied. Bit-exact."""
import numpy as np
import pytest

from solverforge_b200 import ForageParams, GpuScoreDirector, instances, models
from solverforge_b200 import _lib as L
from tests import oracle_lib
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu


def _asym(c, seed):
    """asymmetric costs: reversed segments change their inner legs"""
    rng = np.random.default_rng(seed)
    c.matrix = c.matrix + rng.integers(0, 40, size=c.matrix.shape)
    np.fill_diagonal(c.matrix, 0)
    return c


@pytest.mark.parametrize("k,min_seg", [(2, 1), (3, 1), (3, 2), (4, 1), (5, 1)])
def test_every_k_opt_move_scores_like_the_oracle(k, min_seg):
    c = _asym(instances.cvrp(16 if k >= 4 else 24, 3, seed=61), 5)
    R = 2
    starts = [instances.perturb_routes(c, 70 + r, 12) for r in range(R)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    rows, scores = [], []
    for r in range(R):
        o = Oracle.cvrp(c, *starts[r])
        rw = o.enumerate_k_opt(k, min_seg)
        assert len(rw) > 0
        rows.append(rw)
        scores.append(o.score_k_opt(rw, k))
    offs = np.concatenate([[0], np.cumsum([len(x) for x in rows])]).astype(np.uint64)
    sc, ok = d.score_k_opt(np.concatenate(rows), k, offs)
    for r in range(R):
        assert np.array_equal(ok[offs[r]:offs[r + 1]], scores[r][1])
        assert np.array_equal(sc[offs[r]:offs[r + 1]], scores[r][0])
    # not doable rows: unordered cuts, a cut beyond the list
    bad = rows[0][:2].copy()
    bad[0, 1], bad[0, 2] = bad[0, 2], bad[0, 1]
    bad[1, k] = 60000
    sc, ok = d.score_k_opt(np.concatenate([bad, np.zeros((0, k + 2), np.uint32)]), k, np.array([0, 2, 2], dtype=np.uint64))
    assert ok.tolist() == [0, 0]


def test_reference_three_opt_tour_245_moves():
    """The 8-city single tour of the reference's selector KAT (245 = C(7, 3) x 7 3-opt moves, the oracle pins the count
    and the order in kat_main.cpp): scores of all 245 moves."""
    c = instances.cvrp(8, 1, seed=9)
    d = models.cvrp_director(c)
    o = Oracle.cvrp(c)
    rows = o.enumerate_k_opt(3, 1)
    assert len(rows) == 245
    sc, ok = d.score_k_opt(rows, 3)
    so, oko = o.score_k_opt(rows, 3)
    assert np.array_equal(ok, oko) and np.array_equal(sc, so)


@pytest.mark.parametrize("order", [L.ORDER_ORIGINAL, L.ORDER_RANDOM, L.ORDER_SHUFFLED])
def test_k_opt_cursor_on_device_in_every_selection_order(order):
    c = _asym(instances.cvrp(30, 4, seed=62), 6)
    R = 3
    starts = [instances.perturb_routes(c, 90 + r, 15) for r in range(R)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    child = [(L.FAM_K_OPT, 3, 1)]
    desc = GpuScoreDirector.union_desc(child, L.UNION_SEQUENTIAL, order, 64, 1 << 15)
    seeds, steps = [3, 99, 0xABCDEF], [0, 2, 40]
    base = d.calculate_score()
    for acceptor, okind, dl in ((0, 3, 0), (1, 0, -25)):
        for ties, limit in ((1, 0), (0, 1), (1, 37), (1, 700)):
            ref = np.concatenate([base + [0, dl], base + [0, dl]], axis=1)
            idx, best, ev, win, flags = d.step_union(desc, ForageParams(acceptor, ties, limit), step_seeds=seeds,
                                                     step_indices=steps, ref_scores=ref)
            for r, o in enumerate(oracles):
                out, chd, loc, data, sc = oracle_lib.union_step(o, child, L.UNION_SEQUENTIAL, order, steps[r], seeds[r], ref[r][:2],
                                                                ref[r][2:], 0 if limit else 2, max(limit, 1), bool(ties), okind)
                assert flags[r] == 0 and int(ev[r]) == out[2]
                if out[0]:
                    assert int(idx[r]) == out[1] and best[r].tolist() == sc[out[1]].tolist()
                    assert win[r].tolist() == [L.FAM_K_OPT, 0] + data[0][1][int(loc[out[1]])].tolist() + [int(loc[out[1]]), 0]
                else:
                    assert idx[r] == 0xFFFFFFFF


def test_k_opt_apply_chain_and_union_with_the_default_families():
    c = _asym(instances.cvrp(36, 5, seed=63), 7)
    start = instances.perturb_routes(c, 11, 20)
    d = models.cvrp_director(c, 1, offsets=start[0][None, :], elems=start[1])
    o = Oracle.cvrp(c, *start)
    children = [(L.FAM_NEARBY_LIST_CHANGE, 10), (L.FAM_K_OPT, 3, 1), (L.FAM_LIST_REVERSE,), (L.FAM_K_OPT, 2, 2)]
    desc = GpuScoreDirector.union_desc(children, L.UNION_STRATIFIED_RANDOM, L.ORDER_RANDOM, 32, 1 << 14)
    apply_fn = {0: "apply_list_change", 4: "apply_list_reverse"}
    kopt = 0
    for step in range(30):
        last = d.calculate_score()
        ref = np.concatenate([last, last], axis=1)
        idx, best, ev, win, flags = d.step_union(desc, ForageParams(1, 1, 20), step_seeds=[700 + step], step_indices=[step],
                                                 ref_scores=ref, apply=True)
        out, chd, loc, data, sc = oracle_lib.union_step(o, children, L.UNION_STRATIFIED_RANDOM, L.ORDER_RANDOM, step, 700 + step,
                                                        last[0], last[0], 0, 20, True, 0)
        assert int(ev[0]) == out[2] and flags[0] == 0
        if not out[0]:
            assert idx[0] == 0xFFFFFFFF
            continue
        assert int(idx[0]) == out[1]
        ci, j = int(chd[out[1]]), int(loc[out[1]])
        fam = children[ci][0]
        if fam == L.FAM_K_OPT:
            o.apply_k_opt(data[ci][0][j], children[ci][1])
            kopt += 1
        else:
            getattr(o, apply_fn[fam])(*[int(x) for x in data[ci][0][j]])
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist() == d.fresh_score()[0].tolist()
        offs, el = d.list_state()
    assert kopt > 0
    # the rows-resident apply call
    rows = o.enumerate_k_opt(3, 1)
    so, oko = o.score_k_opt(rows, 3)
    pick = int(np.flatnonzero(oko)[len(rows) // 2])
    d.apply_k_opt(rows[pick:pick + 1], 3)
    o.apply_k_opt(rows[pick], 3)
    assert d.calculate_score()[0].tolist() == o.committed_score().tolist() == d.fresh_score()[0].tolist()
