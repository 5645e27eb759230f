"""Monomorphised scalar scoring programs (sfgpu_spec.cuh) against the constraint-table interpreter and the
oracle: the same model committed twice (default context / SFGPU_CTX_GENERIC_KERNELS), full ChangeMove
neighbourhood through the rows-resident call and through the fused device step. Bit-exact.

Reference counterpart: monomorphised ConstraintSet tuples,
solverforge-scoring/src/api/constraint_set/incremental.rs:339-408."""
import numpy as np
import pytest

from solverforge_b200 import ForageParams, instances, models
from solverforge_b200 import _lib as L
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu


def _pair(factory, *args, **kw):
    return factory(*args, **kw), factory(*args, flags=L.CTX_GENERIC_KERNELS, **kw)


def _check(spec, gen, oracle, expect_spec=True, steps=4):
    assert gen.scalar_program() == -1
    assert (spec.scalar_program() >= 0) == expect_spec
    for step in range(steps):
        rows = oracle.enumerate_change()
        so, oko = oracle.score_change(rows)
        # ragged extras the reference would reject before scoring: out-of-range entity / value -> not doable
        extra = np.array([[2 ** 31 - 1, 0], [0, 2 ** 20], [rows[0][0], rows[0][1]]], dtype=np.int64)
        for d in (spec, gen):
            s, ok = d.score_change(np.concatenate([rows, extra]))
            assert np.array_equal(ok[:len(rows)], oko) and ok[-3:].tolist() == [0, 0, int(oko[0])]
            assert np.array_equal(s[:len(rows)], so) and s[-3:].tolist() == [[0, 0], [0, 0], so[0].tolist()]
        base = spec.calculate_score()
        ref = np.concatenate([base, base], axis=1)
        outs = []
        for d in (spec, gen):
            for params in (ForageParams(0, 1, 0), ForageParams(1, 1, 7), ForageParams(2, 0, 0)):
                outs.append([np.asarray(x).tolist() for x in d.step_change(params, step_seeds=[17 + step], ref_scores=ref)])
        assert outs[:3] == outs[3:]
        # commit the BestScore winner on both and on the oracle, then go round again
        win = None
        for d in (spec, gen):
            idx, best, ev, win = d.step_change(ForageParams(0, 1, 0), step_seeds=[17 + step], ref_scores=ref, apply=True)
        if int(idx[0]) == 0xFFFFFFFF:
            break
        oracle.apply_change(int(win[0][0]), int(np.int32(win[0][1])))
        want = oracle.committed_score().tolist()
        assert spec.calculate_score()[0].tolist() == want
        assert gen.calculate_score()[0].tolist() == want
        assert spec.fresh_score()[0].tolist() == want


def test_graph_coloring_spec():
    g = instances.graph_coloring(900, 4000, 6, seed_edges=5, seed_colors=6, unassigned_permille=70)
    _check(*_pair(models.graph_coloring_director, g), Oracle.graph_coloring(g))


def test_nqueens_spec():
    q = instances.nqueens(40, seed=8)
    _check(*_pair(models.nqueens_director, q), Oracle.nqueens(q))


@pytest.mark.parametrize("with_complement", [False, True])
def test_job_shop_spec(with_complement):
    j = instances.job_shop(40, 9, 6, seed=4, unassigned_permille=50)
    spec, gen = _pair(models.job_shop_director, j, with_complement=with_complement)
    _check(spec, gen, Oracle.job_shop(j, with_complement=with_complement))


def test_unspecialised_programs_keep_the_interpreter():
    s = instances.shift_scheduling(seed=31)
    _check(*_pair(models.shift_scheduling_director, s), Oracle.shift_scheduling(s), expect_spec=False, steps=2)


def test_spec_multi_replica_full_size_c2():
    """BASELINE C2 shape with several replicas: spec == interpreter on all 90k candidates per replica."""
    g = instances.graph_coloring()
    R = 3
    colors = np.stack([instances.graph_coloring(seed_colors=50 + r).color for r in range(R)])
    spec, gen = _pair(models.graph_coloring_director, g, R, colors=colors)
    assert spec.scalar_program() >= 0
    def change_rows(color):   # canonical ChangeMoveSelector order (change.rs:66-104): k values, then to-None when assigned
        out = []
        for e in range(g.n):
            out += [(e, v) for v in range(g.k)]
            if color[e] >= 0:
                out.append((e, -1))
        return np.array(out, dtype=np.int64)

    per = [change_rows(colors[r]) for r in range(R)]
    offs = np.concatenate([[0], np.cumsum([len(p) for p in per])]).astype(np.uint64)
    allrows = np.concatenate(per)
    s1, ok1 = spec.score_change(allrows, offs)
    s2, ok2 = gen.score_change(allrows, offs)
    assert np.array_equal(s1, s2) and np.array_equal(ok1, ok2)
    o1 = Oracle.graph_coloring(g, colors[1])   # one oracle only: its predicate join initialises in O(n^2)
    assert np.array_equal(o1.enumerate_change(), per[1])
    so, oko = o1.score_change(per[1][::7])
    lo = int(offs[1])
    assert np.array_equal(s1[lo:int(offs[2])][::7], so) and np.array_equal(ok1[lo:int(offs[2])][::7], oko)
