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


def _rows_step(d, rows, offs, params, seeds, ref, materialise):
    """sfgpu_step_change_rows through torch device buffers; returns (index, best, evaluated[, scores, doable])."""
    import torch
    R, n = d.R, len(rows)
    t_rows = torch.from_numpy(np.ascontiguousarray(rows.astype(np.int64).astype(np.uint32)).view(np.int32)).cuda()
    t_offs = torch.from_numpy(np.asarray(offs, dtype=np.uint64).view(np.int64)).cuda()
    t_seeds = torch.tensor(seeds, dtype=torch.int64).cuda()
    t_ref = torch.from_numpy(np.ascontiguousarray(ref, dtype=np.int64)).cuda()
    t_idx = torch.zeros(R, dtype=torch.int32).cuda()
    t_best = torch.zeros((R, 2), dtype=torch.int64).cuda()
    t_ev = torch.zeros(R, dtype=torch.int32).cuda()
    t_sc = torch.zeros((n, 2), dtype=torch.int64).cuda()
    t_ok = torch.zeros(n, dtype=torch.uint8).cuda()
    torch.cuda.synchronize()
    d.step_change_rows_device(n, t_offs.data_ptr(), t_rows.data_ptr(), params, t_seeds.data_ptr(), t_ref.data_ptr(),
                              t_sc.data_ptr() if materialise else 0, t_ok.data_ptr() if materialise else 0,
                              t_idx.data_ptr(), t_best.data_ptr(), t_ev.data_ptr())
    d.synchronize()
    out = [t_idx.cpu().numpy().view(np.uint32), t_best.cpu().numpy(), t_ev.cpu().numpy().view(np.uint32)]
    if materialise:
        out += [t_sc.cpu().numpy(), t_ok.cpu().numpy()]
    return out


@pytest.mark.parametrize("model", ["graph_coloring", "job_shop", "job_shop_plain", "nqueens", "shift"])
def test_fused_rows_step_equals_score_plus_argbest(model, monkeypatch):
    """sfgpu_step_change_rows (monomorphised kernel with fused forager partials, int32 or int64 deltas) picks the
    same winner as scoring + the ordered argbest replay for every acceptor / tie mode / limit, with and without
    materialised scores, on several replicas with ragged batches; the int32 programs equal the int64 ones."""
    R = 3
    if model == "graph_coloring":
        inst = instances.graph_coloring(700, 3000, 5, seed_edges=3, seed_colors=4, unassigned_permille=60)
        states = np.stack([instances.graph_coloring_colors(inst, 4 + r, 60) for r in range(R)])
        make = lambda **kw: models.graph_coloring_director(inst, R, colors=states, **kw)
        k = inst.k
    elif model.startswith("job_shop"):
        inst = instances.job_shop(50, 8, 7, seed=6, unassigned_permille=40)
        states = np.stack([instances.job_shop_machines(inst, 6 + r, 40) for r in range(R)])
        make = lambda **kw: models.job_shop_director(inst, R, machine_idx=states, with_complement=model == "job_shop", **kw)
        k = inst.n_machines
    elif model == "nqueens":
        inst = instances.nqueens(30, seed=2)
        states = np.stack([instances.nqueens(30, seed=2 + r).row for r in range(R)])
        make = lambda **kw: models.nqueens_director(inst, R, rows=states, **kw)
        k = inst.n
    else:
        inst = instances.shift_scheduling(seed=9)
        states = np.stack([instances.shift_scheduling(seed=9 + r).nurse_idx for r in range(R)])
        make = lambda **kw: models.shift_scheduling_director(inst, R, nurse_idx=states, **kw)
        k = inst.n_nurses
    d = make()
    monkeypatch.setenv("SFGPU_NO_NARROW", "1")
    wide = make()
    monkeypatch.delenv("SFGPU_NO_NARROW")
    per = [instances.change_neighbourhood(states[r], k) for r in range(R)]
    per[1] = per[1][: len(per[1]) // 3]                      # ragged: a short batch
    per[2] = np.concatenate([per[2], per[2][:50], [[0, k + 5], [10 ** 6, 0]]])   # duplicates (ties) + not-doable rows
    offs = np.concatenate([[0], np.cumsum([len(p) for p in per])]).astype(np.uint64)
    rows = np.concatenate(per)
    s_ref, ok_ref = wide.score_change(rows, offs)
    s_n, ok_n = d.score_change(rows, offs)
    assert np.array_equal(s_n, s_ref) and np.array_equal(ok_n, ok_ref)
    base = d.calculate_score()
    late = base - np.array([[1, 3], [0, 40], [2, 0]])
    ref = np.concatenate([base, late], axis=1)
    seeds = [5, 77, 123456789]
    for acceptor in (0, 1, 2, 3):
        for tie in (0, 1):
            for limit in (0, 5):
                p = ForageParams(acceptor, tie, limit)
                want = d.argbest(s_ref, ok_ref, offs, p, seeds, ref)
                for dd in (d, wide):
                    for mat in (True, False):
                        got = _rows_step(dd, rows, offs, p, seeds, ref, mat)
                        assert np.array_equal(got[0], want[0]), (model, acceptor, tie, limit, mat)
                        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
                        if mat:
                            assert np.array_equal(got[3], s_ref) and np.array_equal(got[4], ok_ref)
