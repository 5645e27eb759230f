"""CPU checks of the union-cursor checker the GPU tests rely on: the oracle's UnionScheduler
(heuristic/selector/decorator/vec_union.rs:190-366) against the reference's own union tests
(decorator/vec_union/tests.rs:204-297) and the per-family cursors against the host selectors."""
import numpy as np

from solverforge_b200 import instances, selectors
from solverforge_b200.selectors import MoveStreamContext
from tests import oracle_lib
from tests.oracle_lib import Oracle


def _drain(kids, union_order, **kw):
    child, local = oracle_lib.union_pull_order([len(k) for k in kids], union_order, **kw)
    return [kids[c][j] for c, j in zip(child.tolist(), local.tolist())]


def test_union_scheduler_reference_vectors():
    assert _drain([[1, 2, 3], [10, 11]], 0) == [1, 2, 3, 10, 11]                         # sequential
    assert _drain([[1, 2, 3], [], [10], [20, 21]], 1) == [1, 10, 20, 2, 21, 3]           # round robin skips empty / exhausted
    ctx = MoveStreamContext(0, 2, selectors.RANDOM)
    offset = ctx.random_index(3, 0xA11CE5E1EC700001)
    want = {0: [1, 10, 20, 2, 11, 21], 1: [10, 20, 1, 11, 21, 2], 2: [20, 1, 10, 21, 2, 11]}[offset]
    assert _drain([[1, 2], [10, 11], [20, 21]], 2, step_index=0, step_seed=2, order=1) == want   # rotating round robin
    got = _drain([[1, 2], [10, 11], [20, 21]], 4, step_index=3, step_seed=42, order=1)           # stratified random
    assert len(got) == 6 and {1, 10, 20} == set(got[:3]) and {2, 11, 21} == set(got[3:])
    # weights: a child with weight w is pulled w times per round of sum(w) pulls (smooth weighted round-robin)
    child, _ = oracle_lib.union_pull_order([100, 100, 100], 4, step_seed=9, order=1, weights=[3, 1, 2], limit=60)
    assert np.bincount(child, minlength=3).tolist() == [30, 10, 20]
    for at in range(0, 60, 6):
        assert np.bincount(child[at:at + 6], minlength=3).tolist() == [3, 1, 2]
    # the pull limit truncates, exhausted children drop out
    child, local = oracle_lib.union_pull_order([2, 50], 4, step_seed=5, order=1)
    assert len(child) == 52 and np.bincount(child).tolist() == [2, 50]


def test_oracle_cursors_match_the_host_selectors_in_seeded_orders():
    """The oracle's family cursors (the checker of the device walkers) == the Python host selectors, order by order."""
    c = instances.cvrp(30, 4, seed=17)
    offs, el = instances.perturb_routes(c, 3, 25)
    o = Oracle.cvrp(c, offs, el)
    for order in (selectors.ORIGINAL, selectors.RANDOM, selectors.SHUFFLED):
        for step_index, seed in ((0, 1), (7, 0xBEEF)):
            ctx = MoveStreamContext(step_index, seed, order)
            assert np.array_equal(o.enumerate_nearby_list_change(8, step_index, seed, order),
                                  selectors.nearby_list_change_rows(offs, el, c.matrix, 8, ctx))
            assert np.array_equal(o.enumerate_list_reverse(step_index, seed, order), selectors.list_reverse_rows(offs, ctx))
            assert np.array_equal(o.enumerate_k_opt(3, 1, step_index, seed, order), selectors.k_opt_rows(offs, 3, 1, ctx))
