"""GPU parity of the device-side ChangeMove neighbourhood + fused step (sfgpu_step_change) against the
oracle's ChangeMoveSelector order, scores and candidate-loop replay. Bit-exact."""
import numpy as np
import pytest

from solverforge_b200 import ForageParams, instances, models
from solverforge_b200 import _lib as L
from tests import oracle_lib
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu


def _run_and_check(d, oracle, n_entities, k, seeds=(3, 11), limits=(0, 1, 9, 300, 10 ** 6)):
    import torch
    dev = torch.device("cuda")
    S = n_entities * (k + 1)
    t_off = torch.zeros(2, dtype=torch.int64, device=dev)
    t_rows = torch.zeros((S, 2), dtype=torch.int32, device=dev)
    t_scores = torch.zeros((S, 2), dtype=torch.int64, device=dev)
    t_doable = torch.full((S,), 7, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()  # the director launches on its own stream: order after torch's fills
    d.step_change(ForageParams(0, 1, 0), step_seeds=[1], out_offsets_ptr=t_off.data_ptr(),
                  out_rows_ptr=t_rows.data_ptr(), out_scores_ptr=t_scores.data_ptr(), out_doable_ptr=t_doable.data_ptr())
    want = oracle.enumerate_change()
    so, oko = oracle.score_change(want)
    n = len(want)
    rows = t_rows.cpu().numpy()
    assert t_off.cpu().tolist() == [0, S]
    assert np.array_equal(rows[:n], want.astype(np.int32)), "generated ChangeMove order differs from the reference"
    assert np.array_equal(t_doable.cpu().numpy()[:n], oko)
    assert np.array_equal(t_scores.cpu().numpy()[:n], so)
    assert (t_doable.cpu().numpy()[n:] == 0).all() and (rows[n:, 0] == -1).all()
    base = d.calculate_score()[0]
    for seed in seeds:
        for acceptor, okind in ((0, 3), (1, 0), (2, 1)):
            for ties in (0, 1):
                for limit in limits:
                    ref = np.concatenate([base + [0, -1], base + [-1, 0]])[None, :]
                    idx, best, ev, win = d.step_change(ForageParams(acceptor, ties, limit), step_seeds=[seed],
                                                       ref_scores=ref)
                    out = oracle_lib.replay_step(so, oko, [0, 0], ref[0][:2], ref[0][2:], seed, 0 if limit else 2,
                                                 max(limit, 1), bool(ties), okind)
                    what = f"seed={seed} acc={acceptor} ties={ties} limit={limit}"
                    assert int(ev[0]) == out[2], what + " moves_evaluated"
                    if out[0]:
                        assert int(idx[0]) == out[1], what
                        assert best[0].tolist() == so[out[1]].tolist(), what
                        assert win[0].tolist() == want[out[1]].tolist(), what
                    else:
                        assert idx[0] == 0xFFFFFFFF, what


def test_graph_coloring_change_step():
    g = instances.graph_coloring(700, 3000, 5, seed_edges=3, seed_colors=4, unassigned_permille=80)
    _run_and_check(models.graph_coloring_director(g), Oracle.graph_coloring(g), g.n, g.k)


def test_job_shop_change_step():
    j = instances.job_shop(30, 8, 5, seed=3, unassigned_permille=60)
    _run_and_check(models.job_shop_director(j), Oracle.job_shop(j), j.n_ops, j.n_machines, seeds=(5,))


def test_shift_and_nqueens_change_step():
    s = instances.shift_scheduling(seed=25)
    _run_and_check(models.shift_scheduling_director(s), Oracle.shift_scheduling(s), s.n_shifts, s.n_nurses, seeds=(2,))
    q = instances.nqueens(24, seed=5)
    _run_and_check(models.nqueens_director(q), Oracle.nqueens(q), q.n, q.n, seeds=(9,), limits=(0, 50))


def test_change_step_loop_with_apply_tracks_oracle():
    g = instances.graph_coloring(400, 1600, 4, seed_edges=8, seed_colors=9, unassigned_permille=100)
    R = 3
    colors = np.stack([instances.graph_coloring(400, 1600, 4, seed_edges=8, seed_colors=20 + r).color for r in range(R)])
    d = models.graph_coloring_director(g, R, colors=colors)
    oracles = [Oracle.graph_coloring(g, colors[r]) for r in range(R)]
    for step in range(12):
        last = d.calculate_score()
        ref = np.concatenate([last, last], axis=1)
        idx, best, ev, win = d.step_change(ForageParams(1, 1, 0), step_seeds=[40 + step] * R, ref_scores=ref, apply=True)
        after = d.calculate_score()
        for r in range(R):
            rows = oracles[r].enumerate_change()
            so, oko = oracles[r].score_change(rows)
            out = oracle_lib.replay_step(so, oko, [0, 0], last[r], last[r], 40 + step, 2, 1, True, 0)
            if out[0]:
                assert int(idx[r]) == out[1], f"step {step} replica {r}"
                oracles[r].apply_change(*rows[out[1]])
            else:
                assert idx[r] == 0xFFFFFFFF
            assert after[r].tolist() == oracles[r].committed_score().tolist()
        assert np.array_equal(d.fresh_score(), after)


def test_device_resident_change_loop_equals_step_by_step():
    from solverforge_b200.selectors import splitmix64
    g = instances.graph_coloring(300, 1200, 4, seed_edges=2, seed_colors=3, unassigned_permille=150)
    R, steps, late = 2, 40, 7
    colors = np.stack([instances.graph_coloring(300, 1200, 4, seed_edges=2, seed_colors=30 + r, unassigned_permille=150).color
                       for r in range(R)])
    loop = models.graph_coloring_director(g, R, colors=colors)
    ref = models.graph_coloring_director(g, R, colors=colors)
    seed_base = 77
    best, evaluated, committed = loop.solve_change(steps, 2, late, 1, 25, seed_base)
    init = ref.calculate_score()
    history = [[init[r].copy() for _ in range(late)] for r in range(R)]
    hidx = [0] * R
    best_h = init.copy()
    ev_h = np.zeros(R, dtype=np.uint64)
    for t in range(steps):
        last = ref.calculate_score()
        refs = np.stack([np.concatenate([last[r], history[r][hidx[r]]]) for r in range(R)])
        seeds = [splitmix64(seed_base ^ ((r * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)) ^ t) for r in range(R)]
        idx, b, ev, win = ref.step_change(ForageParams(2, 1, 25), step_seeds=seeds, ref_scores=refs, apply=True)
        after = ref.calculate_score()
        for r in range(R):
            history[r][hidx[r]] = after[r].copy()
            hidx[r] = (hidx[r] + 1) % late
            ev_h[r] += ev[r]
            if (after[r][0], after[r][1]) > (best_h[r][0], best_h[r][1]):
                best_h[r] = after[r]
    assert np.array_equal(loop.calculate_score(), ref.calculate_score())
    assert np.array_equal(loop.scalar_state(), ref.scalar_state())
    assert np.array_equal(best, best_h) and np.array_equal(evaluated, ev_h)
    assert np.array_equal(loop.fresh_score(), loop.calculate_score())
    assert (best[:, 0] > init[:, 0]).all()      # hard score improved (conflicts / unassigned removed)


@pytest.mark.parametrize("kind,okind,size,real,limit", [
    (3, 4, 0, 0.01, 0), (4, 5, 2, 0.0, 15), (5, 6, 5, 0.02, 0),
])
def test_scalar_device_loop_stateful_acceptors_follow_the_oracle(kind, okind, size, real, limit):
    """sfgpu_solve_change with GreatDeluge / StepCountingHillClimbing / DiversifiedLateAcceptance against the
    oracle's stateful acceptors driving the oracle's ChangeMoveSelector + scoring + forager."""
    from solverforge_b200.selectors import splitmix64
    from tests.oracle_lib import OracleAcceptor
    g = instances.graph_coloring(150, 500, 4, seed_edges=6, seed_colors=9, unassigned_permille=100)
    R, steps = 2, 25
    colors = np.stack([instances.graph_coloring(150, 500, 4, seed_edges=6, seed_colors=50 + r, unassigned_permille=100).color
                       for r in range(R)])
    loop = models.graph_coloring_director(g, R, colors=colors)
    seed_base = 31337
    best, evaluated, committed = loop.solve_change(steps, kind, max(size, 1), 1, limit, seed_base, acceptor_real=real,
                                                   step_count_limit=size)
    final = loop.calculate_score()
    state = loop.scalar_state()
    for r in range(R):
        o = Oracle.graph_coloring(g, colors[r])
        acc = OracleAcceptor(okind, size=size, real=real)
        init = o.committed_score()
        acc.phase_started(init)
        best_o, ev_o, steps_o = init.copy(), 0, 0
        cur = colors[r].copy()
        for t in range(steps):
            last = o.committed_score()
            rows = o.enumerate_change()
            so, oko = o.score_change(rows)
            seed = splitmix64(seed_base ^ ((r * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)) ^ t)
            out = acc.step(so, oko, best_o, last, seed, 0 if limit else 2, max(limit, 1), True)
            ev_o += out[2]
            if out[0]:
                e, v = rows[out[1]]
                o.apply_change(e, v)
                cur[int(e)] = int(v)
                steps_o += 1
            now = o.committed_score()
            if (now[0], now[1]) > (best_o[0], best_o[1]):
                best_o = now.copy()
        assert final[r].tolist() == o.committed_score().tolist(), f"kind={kind} r={r}"
        assert best[r].tolist() == best_o.tolist()
        assert int(evaluated[r]) == ev_o and int(committed[r]) == steps_o
        assert np.array_equal(state[r], cur)
    assert np.array_equal(loop.fresh_score(), final)


@pytest.mark.parametrize("tenures,aspiration,limit", [((5, 0, 0, 0), True, 0), ((0, 3, 0, 0), True, 12), ((0, 0, 7, 0), False, 0),
                                                      ((0, 0, 0, 4), True, 0), ((6, 2, 9, 3), True, 25), ((64, 0, 0, 1), False, 1)])
def test_scalar_device_loop_tabu_search_follows_the_oracle(tenures, aspiration, limit):
    """sfgpu_solve_change with TabuSearch (acceptor 7) against the oracle's TabuSearchAcceptor (tabu_search.rs:103-237)
    fed the ChangeMove signatures of the oracle's own selector: entity / value / move / undo-move memories, FIFO
    tenure, aspiration — same trajectory, best score, moves_evaluated, committed steps."""
    from solverforge_b200.selectors import splitmix64
    from tests.oracle_lib import OracleAcceptor, move_signatures
    g = instances.graph_coloring(120, 400, 4, seed_edges=16, seed_colors=19, unassigned_permille=100)
    R, steps = 2, 40
    colors = np.stack([instances.graph_coloring(120, 400, 4, seed_edges=16, seed_colors=70 + r, unassigned_permille=100).color
                       for r in range(R)])
    loop = models.graph_coloring_director(g, R, colors=colors)
    seed_base = 4242
    packed = tenures[0] | (tenures[1] << 8) | (tenures[2] << 16) | (tenures[3] << 24)
    best, evaluated, committed = loop.solve_change(steps, 7, packed, 1, limit, seed_base, step_count_limit=1 if aspiration else 0)
    final = loop.calculate_score()
    state = loop.scalar_state()
    for r in range(R):
        o = Oracle.graph_coloring(g, colors[r])
        acc = OracleAcceptor(OracleAcceptor.TABU, tabu=list(tenures), aspiration=aspiration)
        init = o.committed_score()
        acc.phase_started(init)
        best_o, ev_o, steps_o = init.copy(), 0, 0
        cur = colors[r].copy()
        for t in range(steps):
            last = o.committed_score()
            rows = o.enumerate_change()
            so, oko = o.score_change(rows)
            sigs = move_signatures(o, 0, rows)
            seed = splitmix64(seed_base ^ ((r * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)) ^ t)
            out = acc.step(so, oko, best_o, last, seed, 0 if limit else 2, max(limit, 1), True, signatures=sigs)
            ev_o += out[2]
            if out[0]:
                e, v = rows[out[1]]
                o.apply_change(e, v)
                cur[int(e)] = int(v)
                steps_o += 1
            now = o.committed_score()
            if (now[0], now[1]) > (best_o[0], best_o[1]):
                best_o = now.copy()
        what = f"tenures={tenures} aspiration={aspiration} limit={limit} r={r}"
        assert int(evaluated[r]) == ev_o and int(committed[r]) == steps_o, what
        assert final[r].tolist() == o.committed_score().tolist(), what
        assert best[r].tolist() == best_o.tolist(), what
        assert np.array_equal(state[r], cur), what
    assert np.array_equal(loop.fresh_score(), final)
    with pytest.raises(L.SfgpuError):
        loop.solve_change(2, 7, 0, 1, 0, 1)           # no tabu dimension
    with pytest.raises(L.SfgpuError):
        loop.solve_change(2, 7, 65, 1, 0, 1)          # tenure above the device memory


@pytest.mark.parametrize("limit,samples,decay,never_hard", [(0, 16, 0.9, False), (40, 128, 0.0, False), (7, 24, 0.95, True),
                                                            (1, 8, 0.9, False)])
def test_scalar_device_loop_simulated_annealing_follows_the_oracle(limit, samples, decay, never_hard):
    """sfgpu_solve_change with SimulatedAnnealing (acceptor 6) against the oracle's SimulatedAnnealingAcceptor
    (acceptor/simulated_annealing.rs:338-431: calibration over the evaluated worsening candidates, Boltzmann test,
    decay) fed the same stated uniform stream: same trajectory, best score, moves_evaluated, committed steps."""
    from solverforge_b200.selectors import splitmix64
    from tests.oracle_lib import OracleAcceptor
    g = instances.graph_coloring(150, 500, 4, seed_edges=6, seed_colors=9, unassigned_permille=100)
    R, steps = 2, 30
    colors = np.stack([instances.graph_coloring(150, 500, 4, seed_edges=6, seed_colors=50 + r, unassigned_permille=100).color
                       for r in range(R)])
    loop = models.graph_coloring_director(g, R, colors=colors)
    seed_base = 777
    best, evaluated, committed = loop.solve_change(steps, 6, samples, 1, limit, seed_base, acceptor_real=decay,
                                                   step_count_limit=1 if never_hard else 0)
    final = loop.calculate_score()
    state = loop.scalar_state()
    for r in range(R):
        o = Oracle.graph_coloring(g, colors[r])
        acc = OracleAcceptor(OracleAcceptor.SIMULATED_ANNEALING, size=samples, real=decay, aspiration=2 if never_hard else 1)
        init = o.committed_score()
        acc.phase_started(init)
        best_o, ev_o, steps_o = init.copy(), 0, 0
        cur = colors[r].copy()
        for t in range(steps):
            last = o.committed_score()
            rows = o.enumerate_change()
            so, oko = o.score_change(rows)
            seed = splitmix64(seed_base ^ ((r * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)) ^ t)
            out = acc.step(so, oko, best_o, last, seed, 0 if limit else 2, max(limit, 1), True)
            ev_o += out[2]
            if out[0]:
                e, v = rows[out[1]]
                o.apply_change(e, v)
                cur[int(e)] = int(v)
                steps_o += 1
            now = o.committed_score()
            if (now[0], now[1]) > (best_o[0], best_o[1]):
                best_o = now.copy()
        what = f"limit={limit} samples={samples} r={r}"
        assert int(evaluated[r]) == ev_o and int(committed[r]) == steps_o, what
        assert final[r].tolist() == o.committed_score().tolist(), what
        assert best[r].tolist() == best_o.tolist(), what
        assert np.array_equal(state[r], cur), what
    assert np.array_equal(loop.fresh_score(), final)
