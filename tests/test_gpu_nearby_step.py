"""GPU parity of the device-side nearby neighbourhood + fused step (sfgpu_step_nearby_list_change):
generated rows, their order (CandidateId), scores, acceptor/forager replay and committed winners
against the oracle's NearbyListChangeMoveSelector + candidate loop restatement. Bit-exact."""
import numpy as np
import pytest

from solverforge_b200 import ForageParams, instances, models
from solverforge_b200 import _lib as L
from tests import oracle_lib
from tests.oracle_lib import Oracle, OracleAcceptor

pytestmark = pytest.mark.gpu


def _materialise(d, R, cap, K):
    import torch
    dev = torch.device("cuda")
    S = cap * K
    bufs = (torch.zeros(R + 1, dtype=torch.int64, device=dev), torch.zeros((R * S, 4), dtype=torch.int32, device=dev),
            torch.zeros((R * S, 2), dtype=torch.int64, device=dev), torch.zeros(R * S, dtype=torch.uint8, device=dev))
    torch.cuda.synchronize()  # the director launches on its own stream: order after torch's fills
    return bufs


def _check_generated(t_off, t_rows, t_scores, t_doable, R, cap, K, oracles):
    rows = t_rows.cpu().numpy().view(np.uint32)
    scores = t_scores.cpu().numpy()
    doable = t_doable.cpu().numpy()
    off = t_off.cpu().numpy()
    S = cap * K
    assert off.tolist() == [r * S for r in range(R + 1)]
    out = []
    for r in range(R):
        want = oracles[r].enumerate_nearby_list_change(K)
        blk = slice(r * S, (r + 1) * S)
        ok = doable[blk] == 1
        got = rows[blk][ok]
        assert got.shape == want.shape, f"replica {r}: {got.shape} vs {want.shape}"
        assert np.array_equal(got, want), f"replica {r}: generated rows differ from the reference selector order"
        so, oko = oracles[r].score_list_change(want)
        assert oko.all()
        assert np.array_equal(scores[blk][ok], so), f"replica {r}: scores differ"
        # sentinels everywhere else
        assert (rows[blk][~ok][:, 0] == 0xFFFFFFFF).all()
        out.append((want, so))
    return out


@pytest.mark.parametrize("coarse,n_routes", [(1, 11), (40, 11), (1, 70), (25, 90)])
def test_generated_neighbourhood_and_step_match_oracle(coarse, n_routes):
    # many short routes: most neighbours are route-last, so a batch of 28 expands to more than 32 slot keys
    # (the overflow half of the compaction buffer) and empty routes appear after perturbation
    c = instances.cvrp(180, n_routes, seed=13)
    c.matrix = (c.matrix // coarse) * coarse       # coarse => many equal distances => tie order matters
    R, K = 3, 20
    starts = [instances.perturb_routes(c, 70 + r, 60) for r in range(R)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    cap = 180
    bufs = _materialise(d, R, cap, K)
    base = d.calculate_score()
    idx, best, ev, win = d.step_nearby_list_change(K, ForageParams(0, 1, 0), step_seeds=[5, 6, 7],
                                                   out_offsets_ptr=bufs[0].data_ptr(), out_rows_ptr=bufs[1].data_ptr(),
                                                   out_scores_ptr=bufs[2].data_ptr(), out_doable_ptr=bufs[3].data_ptr())
    gen = _check_generated(*bufs, R, cap, K, oracles)
    for seed in (5, 9):
        for acceptor, okind in ((0, 3), (1, 0), (2, 1)):
            for ties in (0, 1):
                for limit in (0, 1, 37, 256, 100000):
                    ref = np.stack([np.concatenate([base[r] + [0, -20], base[r] + [0, -45]]) for r in range(R)])
                    idx, best, ev, win = d.step_nearby_list_change(K, ForageParams(acceptor, ties, limit),
                                                                   step_seeds=[seed] * R, ref_scores=ref)
                    for r in range(R):
                        rows, so = gen[r]
                        out = oracle_lib.replay_step(so, np.ones(len(so), np.uint8), [0, 0], ref[r][:2], ref[r][2:],
                                                     seed, 0 if limit else 2, max(limit, 1), bool(ties), okind)
                        what = f"coarse={coarse} seed={seed} acc={acceptor} ties={ties} limit={limit} r={r}"
                        assert int(ev[r]) == out[2], what + " moves_evaluated"
                        if out[0]:
                            assert int(idx[r]) == out[1], what
                            assert best[r].tolist() == so[out[1]].tolist(), what
                            assert win[r].tolist() == rows[out[1]].tolist(), what + " winner row"
                        else:
                            assert idx[r] == 0xFFFFFFFF, what


def test_step_loop_on_device_tracks_oracle_for_many_steps():
    c = instances.cvrp(120, 9, seed=3)
    R, K = 2, 12
    starts = [instances.perturb_routes(c, 30 + r, 25) for r in range(R)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    for step in range(25):
        last = d.calculate_score()
        ref = np.concatenate([last, last], axis=1)
        # hill climbing + best score: commit the winner on device in the same call
        idx, best, ev, win = d.step_nearby_list_change(K, ForageParams(1, 1, 0), step_seeds=[100 + step] * R,
                                                       ref_scores=ref, apply=True)
        after = d.calculate_score()
        for r in range(R):
            rows = oracles[r].enumerate_nearby_list_change(K)
            so, oko = oracles[r].score_list_change(rows)
            out = oracle_lib.replay_step(so, oko, [0, 0], last[r], last[r], 100 + step, 2, 1, True, 0)
            if out[0]:
                assert int(idx[r]) == out[1], f"step {step} replica {r}"
                oracles[r].apply_list_change(*rows[out[1]])
            else:
                assert idx[r] == 0xFFFFFFFF
            assert after[r].tolist() == oracles[r].committed_score().tolist(), f"step {step} replica {r}"
        assert np.array_equal(d.fresh_score(), after)


def test_full_size_cvrp_1000_generation_matches_oracle():
    c = instances.cvrp()
    R, K = 2, 20
    starts = [(c.offsets, c.elems), instances.perturb_routes(c, 77, 300)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    bufs = _materialise(d, R, 1000, K)
    idx, best, ev, win = d.step_nearby_list_change(K, ForageParams(0, 1, 0), step_seeds=[1, 2],
                                                   out_offsets_ptr=bufs[0].data_ptr(), out_rows_ptr=bufs[1].data_ptr(),
                                                   out_scores_ptr=bufs[2].data_ptr(), out_doable_ptr=bufs[3].data_ptr())
    gen = _check_generated(*bufs, R, 1000, K, oracles)
    for r in range(R):
        rows, so = gen[r]
        out = oracle_lib.replay_step(so, np.ones(len(so), np.uint8), [0, 0], [0, 0], [0, 0], r + 1, 2, 1, True, 3)
        assert int(idx[r]) == out[1] and int(ev[r]) == 20000
        assert win[r].tolist() == rows[out[1]].tolist()


def test_unsupported_models_are_rejected_not_emulated():
    g = instances.graph_coloring(50, 100, 3)
    d = models.graph_coloring_director(g)
    with pytest.raises(L.SfgpuError) as e:
        d.step_nearby_list_change(20)
    assert e.value.code == L.E_UNSUPPORTED
    c = instances.cvrp(30, 4, seed=2)
    c.matrix = c.matrix.copy()
    c.matrix[3, 7] = -1            # non-finite cell: the fast program (and device generation) is refused
    d2 = models.cvrp_director(c)
    with pytest.raises(L.SfgpuError) as e:
        d2.step_nearby_list_change(20)
    assert e.value.code == L.E_UNSUPPORTED
    d3 = models.cvrp_director(instances.cvrp(30, 4, seed=2))
    with pytest.raises(L.SfgpuError):
        d3.step_nearby_list_change(33)


def _solve_seed(seed_base, r, t):
    from solverforge_b200.selectors import splitmix64
    return splitmix64(seed_base ^ ((r * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)) ^ t)


@pytest.mark.parametrize("acceptor,limit", [(1, 0), (2, 0), (2, 40)])
def test_device_resident_loop_equals_step_by_step_and_oracle(acceptor, limit):
    c = instances.cvrp(100, 8, seed=9)
    R, K, steps, late = 3, 10, 37, 5
    starts = [instances.perturb_routes(c, 11 + r, 30) for r in range(R)]
    offs = np.stack([s[0] for s in starts])
    elems = np.concatenate([s[1] for s in starts])
    loop = models.cvrp_director(c, R, offsets=offs, elems=elems)
    ref = models.cvrp_director(c, R, offsets=offs, elems=elems)
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    seed_base = 0xC0FFEE
    best, evaluated, committed = loop.solve_nearby_list_change(steps, K, acceptor, late, 1, limit, seed_base)
    # the same loop driven one step at a time through sfgpu_step_nearby_list_change (already pinned to the
    # oracle above) with the acceptor state kept on the host
    init = ref.calculate_score()
    history = [[init[r].copy() for _ in range(late)] for r in range(R)]
    hidx = [0] * R
    best_h = init.copy()
    ev_h = np.zeros(R, dtype=np.uint64)
    acc_h = np.zeros(R, dtype=np.uint64)
    for t in range(steps):
        last = ref.calculate_score()
        refs = np.stack([np.concatenate([last[r], history[r][hidx[r]]]) for r in range(R)])
        seeds = [_solve_seed(seed_base, r, t) for r in range(R)]
        idx, b, ev, win = ref.step_nearby_list_change(K, ForageParams(acceptor, 1, limit), step_seeds=seeds,
                                                      ref_scores=refs, apply=True)
        after = ref.calculate_score()
        for r in range(R):
            if t < 6:   # oracle cross-check of the first steps
                rows = oracles[r].enumerate_nearby_list_change(K)
                so, oko = oracles[r].score_list_change(rows)
                out = oracle_lib.replay_step(so, oko, [0, 0], last[r], history[r][hidx[r]], seeds[r],
                                             0 if limit else 2, max(limit, 1), True, 0 if acceptor == 1 else 1)
                if out[0]:
                    assert int(idx[r]) == out[1]
                    oracles[r].apply_list_change(*rows[out[1]])
                assert after[r].tolist() == oracles[r].committed_score().tolist()
            history[r][hidx[r]] = after[r].copy()
            hidx[r] = (hidx[r] + 1) % late
            ev_h[r] += ev[r]
            acc_h[r] += 1 if idx[r] != 0xFFFFFFFF else 0
            if (after[r][0], after[r][1]) > (best_h[r][0], best_h[r][1]):
                best_h[r] = after[r]
    assert np.array_equal(loop.calculate_score(), ref.calculate_score()), "final committed scores differ"
    lo, le = loop.list_state()
    ro, re_ = ref.list_state()
    assert np.array_equal(lo, ro) and np.array_equal(le, re_), "final solutions differ"
    assert np.array_equal(best, best_h)
    assert np.array_equal(evaluated, ev_h) and np.array_equal(committed, acc_h)
    assert np.array_equal(loop.fresh_score(), loop.calculate_score())
    assert (best[:, 0] >= init[:, 0]).all()


def test_device_loop_restore_best_and_improves():
    c = instances.cvrp(200, 12, seed=7)
    d = models.cvrp_director(c, 2)
    init = d.calculate_score()
    best, ev, acc = d.solve_nearby_list_change(300, 20, 2, 50, 1, 64, seed_base=5, restore_best=True)
    now = d.calculate_score()
    assert np.array_equal(now, best)                 # working solution := best solution
    assert np.array_equal(d.fresh_score(), best)     # and it is a consistent state
    for r in range(2):
        assert (best[r][0], best[r][1]) > (init[r][0], init[r][1])
    offs, el = d.list_state()
    assert sorted(el[0][:200].tolist()) == list(range(1, 201))


@pytest.mark.parametrize("kind,okind,size,real,limit", [
    (3, OracleAcceptor.GREAT_DELUGE, 0, 0.002, 0),
    (3, OracleAcceptor.GREAT_DELUGE, 0, 0.0005, 25),
    (4, OracleAcceptor.STEP_COUNTING, 3, 0.0, 0),
    (4, OracleAcceptor.STEP_COUNTING, 2, 0.0, 30),
    (5, OracleAcceptor.DIVERSIFIED_LATE, 4, 0.01, 0),
    (5, OracleAcceptor.DIVERSIFIED_LATE, 6, 0.003, 40),
])
def test_device_loop_stateful_acceptors_follow_the_oracle_trajectory(kind, okind, size, real, limit):
    """GreatDeluge / StepCountingHillClimbing / DiversifiedLateAcceptance kept on the device
    (sfgpu_solve_nearby_list_change) against the oracle's stateful restatement of
    acceptor/{great_deluge,step_counting,diversified_late_acceptance}.rs driving the oracle's own
    selector + scoring + forager, step by step: same final solution, best score and counters."""
    c = instances.cvrp(70, 6, seed=19)
    R, K, steps = 2, 8, 30
    starts = [instances.perturb_routes(c, 40 + r, 20) for r in range(R)]
    offs = np.stack([s[0] for s in starts])
    elems = np.concatenate([s[1] for s in starts])
    loop = models.cvrp_director(c, R, offsets=offs, elems=elems)
    seed_base = 0xBEEF
    best, evaluated, committed = loop.solve_nearby_list_change(
        steps, K, kind, max(size, 1), 1, limit, seed_base, acceptor_real=real, step_count_limit=size)
    final = loop.calculate_score()
    for r in range(R):
        o = Oracle.cvrp(c, *starts[r])
        acc = OracleAcceptor(okind, size=size, real=real)
        init = o.committed_score()
        acc.phase_started(init)
        best_o = init.copy()
        ev_o = 0
        steps_o = 0
        for t in range(steps):
            last = o.committed_score()
            rows = o.enumerate_nearby_list_change(K)
            so, oko = o.score_list_change(rows)
            out = acc.step(so, oko, best_o, last, _solve_seed(seed_base, r, t), 0 if limit else 2, max(limit, 1), True)
            ev_o += out[2]
            if out[0]:
                o.apply_list_change(*rows[out[1]])
                steps_o += 1
            now = o.committed_score()
            if (now[0], now[1]) > (best_o[0], best_o[1]):
                best_o = now.copy()
        what = f"kind={kind} replica={r}"
        assert final[r].tolist() == o.committed_score().tolist(), what
        assert best[r].tolist() == best_o.tolist(), what
        assert int(evaluated[r]) == ev_o and int(committed[r]) == steps_o, what
    assert np.array_equal(loop.fresh_score(), final)


@pytest.mark.parametrize("tenures,aspiration,limit,n,routes,K", [
    ((3, 0, 0, 0), True, 0, 70, 6, 8), ((0, 6, 0, 0), True, 20, 70, 6, 8), ((0, 0, 9, 0), False, 0, 70, 6, 8),
    ((0, 0, 0, 5), True, 0, 70, 6, 8), ((2, 4, 6, 3), True, 30, 90, 9, 12), ((1, 0, 0, 2), True, 0, 12, 2, 20),
])
def test_device_loop_tabu_search_follows_the_oracle_trajectory(tenures, aspiration, limit, n, routes, K):
    """sfgpu_solve_nearby_list_change with TabuSearch (acceptor 7) against the oracle's TabuSearchAcceptor fed the
    ListChangeMove signatures (source / destination entity, moved element, move id with the adjusted destination, undo
    id) of the oracle's own selector. The last case has fewer candidates per source than max_nearby (padded rows)."""
    from tests.oracle_lib import move_signatures
    c = instances.cvrp(n, routes, seed=23)
    R, steps = 2, 35
    starts = [instances.perturb_routes(c, 60 + r, n // 3) for r in range(R)]
    loop = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    seed_base = 0xFACE
    packed = tenures[0] | (tenures[1] << 8) | (tenures[2] << 16) | (tenures[3] << 24)
    best, evaluated, committed = loop.solve_nearby_list_change(steps, K, 7, packed, 1, limit, seed_base,
                                                               step_count_limit=1 if aspiration else 0)
    final = loop.calculate_score()
    for r in range(R):
        o = Oracle.cvrp(c, *starts[r])
        acc = OracleAcceptor(OracleAcceptor.TABU, tabu=list(tenures), aspiration=aspiration)
        init = o.committed_score()
        acc.phase_started(init)
        best_o, ev_o, steps_o = init.copy(), 0, 0
        for t in range(steps):
            last = o.committed_score()
            rows = o.enumerate_nearby_list_change(K)
            so, oko = o.score_list_change(rows)
            sigs = move_signatures(o, 2, rows)
            out = acc.step(so, oko, best_o, last, _solve_seed(seed_base, r, t), 0 if limit else 2, max(limit, 1), True,
                           signatures=sigs)
            ev_o += out[2]
            if out[0]:
                o.apply_list_change(*rows[out[1]])
                steps_o += 1
            now = o.committed_score()
            if (now[0], now[1]) > (best_o[0], best_o[1]):
                best_o = now.copy()
        what = f"tenures={tenures} aspiration={aspiration} limit={limit} replica={r}"
        assert int(evaluated[r]) == ev_o and int(committed[r]) == steps_o, what
        assert final[r].tolist() == o.committed_score().tolist(), what
        assert best[r].tolist() == best_o.tolist(), what
    assert np.array_equal(loop.fresh_score(), final)


@pytest.mark.parametrize("limit,samples,decay,never_hard,n,routes,K", [(0, 16, 0.9, False, 70, 6, 8), (30, 64, 0.0, False, 90, 9, 12),
                                                                       (5, 24, 0.95, True, 70, 6, 8), (1, 8, 0.9, False, 12, 2, 20)])
def test_device_loop_simulated_annealing_follows_the_oracle_trajectory(limit, samples, decay, never_hard, n, routes, K):
    """sfgpu_solve_nearby_list_change with SimulatedAnnealing (acceptor 6) against the oracle's SimulatedAnnealingAcceptor
    fed the same stated uniform stream over the oracle's nearby ListChange cursor."""
    c = instances.cvrp(n, routes, seed=29)
    R, steps = 2, 35
    starts = [instances.perturb_routes(c, 80 + r, n // 3) for r in range(R)]
    loop = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    seed_base = 0xABCD
    best, evaluated, committed = loop.solve_nearby_list_change(steps, K, 6, samples, 1, limit, seed_base, acceptor_real=decay,
                                                               step_count_limit=1 if never_hard else 0)
    final = loop.calculate_score()
    for r in range(R):
        o = Oracle.cvrp(c, *starts[r])
        acc = OracleAcceptor(OracleAcceptor.SIMULATED_ANNEALING, size=samples, real=decay, aspiration=2 if never_hard else 1)
        init = o.committed_score()
        acc.phase_started(init)
        best_o, ev_o, steps_o = init.copy(), 0, 0
        for t in range(steps):
            last = o.committed_score()
            rows = o.enumerate_nearby_list_change(K)
            so, oko = o.score_list_change(rows)
            out = acc.step(so, oko, best_o, last, _solve_seed(seed_base, r, t), 0 if limit else 2, max(limit, 1), True)
            ev_o += out[2]
            if out[0]:
                o.apply_list_change(*rows[out[1]])
                steps_o += 1
            now = o.committed_score()
            if (now[0], now[1]) > (best_o[0], best_o[1]):
                best_o = now.copy()
        what = f"limit={limit} samples={samples} replica={r}"
        assert int(evaluated[r]) == ev_o and int(committed[r]) == steps_o, what
        assert final[r].tolist() == o.committed_score().tolist(), what
        assert best[r].tolist() == best_o.tolist(), what
    assert np.array_equal(loop.fresh_score(), final)


def test_great_deluge_form_on_the_fused_step():
    """forage acceptor 3 (score > last || score >= threshold) on sfgpu_step_nearby_list_change."""
    c = instances.cvrp(60, 5, seed=4)
    d = models.cvrp_director(c)
    o = Oracle.cvrp(c)
    K = 12
    base = d.calculate_score()[0]
    rows = o.enumerate_nearby_list_change(K)
    so, oko = o.score_list_change(rows)
    for dl, dt in ((0, -30), (0, 0), (-1, 0), (0, 40), (1, 0)):
        last = base + [0, dl * 10]
        water = base + [0, dt]
        idx, best, ev, win = d.step_nearby_list_change(K, ForageParams(3, 1, 0), step_seeds=[11],
                                                       ref_scores=[np.concatenate([last, water])])
        acc = OracleAcceptor(OracleAcceptor.GREAT_DELUGE, real=0.0)
        acc.phase_started(water)   # water level := threshold, rain 0
        out = acc.step(so, oko, base, last, 11, 2, 1, True)
        if out[0]:
            assert int(idx[0]) == out[1] and best[0].tolist() == so[out[1]].tolist()
        else:
            assert idx[0] == 0xFFFFFFFF


def _swap_pull_layout(total, K):
    """per-source candidate counts of the nearby swap cursor: min(K, total - 1 - f)."""
    return [min(K, total - 1 - f) for f in range(total)]


@pytest.mark.parametrize("n,routes,K,coarse", [(150, 9, 20, 1), (150, 9, 20, 30), (40, 3, 32, 1), (90, 6, 5, 7)])
def test_nearby_list_swap_generation_and_step_match_oracle(n, routes, K, coarse):
    """sfgpu_step_nearby_list_swap: generated ListSwapMove rows in the reference cursor order
    (list_kernel/nearby_swap.rs), their scores, forager/acceptor replay and the committed winner."""
    c = instances.cvrp(n, routes, seed=31)
    c.matrix = (c.matrix // coarse) * coarse
    R = 2
    starts = [instances.perturb_routes(c, 90 + r, 40) for r in range(R)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    cap = n
    bufs = _materialise(d, R, cap, K)
    base = d.calculate_score()
    d.step_nearby_list_swap(K, ForageParams(0, 1, 0), step_seeds=[1, 2], out_offsets_ptr=bufs[0].data_ptr(),
                            out_rows_ptr=bufs[1].data_ptr(), out_scores_ptr=bufs[2].data_ptr(),
                            out_doable_ptr=bufs[3].data_ptr())
    rows = bufs[1].cpu().numpy().view(np.uint32)
    scores = bufs[2].cpu().numpy()
    doable = bufs[3].cpu().numpy()
    S = cap * K
    gen = []
    for r in range(R):
        want = oracles[r].enumerate_nearby_list_swap(K)
        blk = slice(r * S, (r + 1) * S)
        ok = doable[blk] == 1
        got = rows[blk][ok]
        assert got.shape == want.shape, f"replica {r}: {got.shape} vs {want.shape}"
        assert np.array_equal(got, want), f"replica {r}: generated swap rows differ from the reference cursor order"
        # fixed stride of K rows per source, existing rows first
        counts = _swap_pull_layout(n, K)
        per_src = ok.reshape(cap, K).sum(axis=1)
        assert per_src.tolist() == counts
        so, oko = oracles[r].score_list_swap(want)
        assert oko.all()
        assert np.array_equal(scores[blk][ok], so), f"replica {r}: swap scores differ"
        gen.append((want, so))
    for seed in (3, 8):
        for acceptor, okind in ((0, 3), (1, 0), (2, 1)):
            for ties in (0, 1):
                for limit in (0, 1, 50, 100000):
                    ref = np.stack([np.concatenate([base[r] + [0, -15], base[r] + [0, -60]]) for r in range(R)])
                    idx, best, ev, win = d.step_nearby_list_swap(K, ForageParams(acceptor, ties, limit),
                                                                 step_seeds=[seed] * R, ref_scores=ref)
                    for r in range(R):
                        want, so = gen[r]
                        out = oracle_lib.replay_step(so, np.ones(len(so), np.uint8), [0, 0], ref[r][:2], ref[r][2:],
                                                     seed, 0 if limit else 2, max(limit, 1), bool(ties), okind)
                        what = f"seed={seed} acc={acceptor} ties={ties} limit={limit} r={r}"
                        assert int(ev[r]) == out[2], what + " moves_evaluated"
                        if out[0]:
                            assert int(idx[r]) == out[1], what
                            assert best[r].tolist() == so[out[1]].tolist(), what
                            assert win[r].tolist() == want[out[1]].tolist(), what + " winner row"
                        else:
                            assert idx[r] == 0xFFFFFFFF, what
    # committing the winners on device keeps cached == fresh == oracle
    last = d.calculate_score()
    idx, best, ev, win = d.step_nearby_list_swap(K, ForageParams(1, 0, 0), step_seeds=[5] * R,
                                                 ref_scores=np.concatenate([last, last], axis=1), apply=True)
    after = d.calculate_score()
    for r in range(R):
        if idx[r] != 0xFFFFFFFF:
            oracles[r].apply_list_swap(*win[r])
        assert after[r].tolist() == oracles[r].committed_score().tolist()
    assert np.array_equal(d.fresh_score(), after)


def test_full_size_cvrp_1000_nearby_swap_matches_oracle():
    """C3 size (1000 customers / 80 routes, max_nearby 20): the whole swap neighbourhood, every score, the winner."""
    c = instances.cvrp()
    K = 20
    start = instances.perturb_routes(c, 78, 300)
    d = models.cvrp_director(c, 1, offsets=start[0][None, :], elems=start[1])
    o = Oracle.cvrp(c, *start)
    bufs = _materialise(d, 1, 1000, K)
    idx, best, ev, win = d.step_nearby_list_swap(K, ForageParams(0, 1, 0), step_seeds=[9],
                                                 out_offsets_ptr=bufs[0].data_ptr(), out_rows_ptr=bufs[1].data_ptr(),
                                                 out_scores_ptr=bufs[2].data_ptr(), out_doable_ptr=bufs[3].data_ptr())
    want = o.enumerate_nearby_list_swap(K)
    ok = bufs[3].cpu().numpy() == 1
    got = bufs[1].cpu().numpy().view(np.uint32)[ok]
    assert len(want) == sum(_swap_pull_layout(1000, K))
    assert np.array_equal(got, want)
    so, oko = o.score_list_swap(want)
    assert oko.all() and np.array_equal(bufs[2].cpu().numpy()[ok], so)
    out = oracle_lib.replay_step(so, np.ones(len(so), np.uint8), [0, 0], [0, 0], [0, 0], 9, 2, 1, True, 3)
    assert int(idx[0]) == out[1] and int(ev[0]) == len(want)
    assert win[0].tolist() == want[out[1]].tolist()
