"""GPU edge cases the reference's tests also poke at: empty and ragged batches, tiny models where the
neighbourhood is smaller than max_nearby, single-value ranges, everything unassigned. Oracle parity."""
import numpy as np
import pytest

from solverforge_b200 import ForageParams, instances, models
from solverforge_b200.instances import CvrpInstance
from tests import oracle_lib
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu


def _tiny_cvrp(routes, dim, seed=1):
    s = instances.splitmix64_stream(seed, dim * dim + dim)
    m = (s[:dim * dim] % np.uint64(50)).astype(np.int64).reshape(dim, dim) + 1   # asymmetric on purpose
    np.fill_diagonal(m, 0)
    demands = (s[dim * dim:] % np.uint64(5)).astype(np.int32) + 1
    demands[0] = 0
    offs = np.cumsum([0] + [len(r) for r in routes]).astype(np.uint32)
    el = np.array([x for r in routes for x in r], dtype=np.uint32)
    return CvrpInstance(dim, len(routes), 6, 0, demands, m, offs, el)


@pytest.mark.parametrize("routes", [
    [[1, 2], [3]],                      # 5 slots - 2 = 3 candidates per source < max_nearby
    [[1], [], [2, 3, 4], []],           # empty routes, single-element route
    [[1, 2, 3, 4, 5, 6]],               # one route only: intra moves only
    [[1], [2]],
])
def test_tiny_neighbourhoods_match_oracle(routes):
    import torch
    dim = 1 + sum(len(r) for r in routes)
    c = _tiny_cvrp(routes, dim)
    o = Oracle.cvrp(c)
    d = models.cvrp_director(c)
    assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
    K = 20
    cap = dim - 1
    dev = torch.device("cuda")
    t_rows = torch.zeros((cap * K, 4), dtype=torch.int32, device=dev)
    t_scores = torch.zeros((cap * K, 2), dtype=torch.int64, device=dev)
    t_doable = torch.zeros(cap * K, dtype=torch.uint8, device=dev)
    t_off = torch.zeros(2, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()  # the director launches on its own stream: order after torch's fills
    idx, best, ev, win = d.step_nearby_list_change(K, ForageParams(0, 0, 0), out_offsets_ptr=t_off.data_ptr(),
                                                   out_rows_ptr=t_rows.data_ptr(), out_scores_ptr=t_scores.data_ptr(),
                                                   out_doable_ptr=t_doable.data_ptr())
    want = o.enumerate_nearby_list_change(K)
    ok = t_doable.cpu().numpy() == 1
    got = t_rows.cpu().numpy().view(np.uint32)[ok]
    assert np.array_equal(got, want), (got.tolist(), want.tolist())
    so, oko = o.score_list_change(want) if len(want) else (np.zeros((0, 2), np.int64), np.zeros(0, np.uint8))
    assert np.array_equal(t_scores.cpu().numpy()[ok], so)
    assert int(ev[0]) == len(want)
    if len(want):
        out = oracle_lib.replay_step(so, oko, [0, 0], [0, 0], [0, 0], 0, 2, 1, False, 3)
        assert int(idx[0]) == out[1] and win[0].tolist() == want[out[1]].tolist()
    else:
        assert idx[0] == 0xFFFFFFFF
    # the generic scoring path agrees too, including out-of-range rows
    rows = np.concatenate([want.reshape(-1, 4), np.array([[0, 0, 0, 0], [9, 0, 0, 0], [0, 7, 0, 0]], dtype=np.uint32)])
    s, okk = d.score_list_change(rows)
    assert okk[-3:].tolist() == [0, 0, 0]


def test_empty_and_ragged_batches():
    import torch
    c = instances.cvrp(40, 4, seed=3)
    R = 3
    d = models.cvrp_director(c, R)
    o = Oracle.cvrp(c)
    rows = o.enumerate_nearby_list_change(5)
    # replica 1 gets no candidates at all
    co = np.array([0, len(rows), len(rows), 2 * len(rows)], dtype=np.uint64)
    batch = np.concatenate([rows, rows])
    s, ok = d.score_list_change(batch, co)
    so, oko = o.score_list_change(rows)
    assert np.array_equal(s[:len(rows)], so) and np.array_equal(s[len(rows):], so)
    idx, best, ev = d.argbest(s, ok, co, ForageParams(0, 0, 0))
    assert idx[1] == 0xFFFFFFFF and ev[1] == 0 and idx[0] == idx[2]
    dev = torch.device("cuda")
    t_off = torch.from_numpy(co.view(np.int64)).to(dev)
    t_rows = torch.from_numpy(batch.view(np.int32)).to(dev)
    t_idx = torch.zeros(R, dtype=torch.int32, device=dev)
    t_best = torch.zeros((R, 2), dtype=torch.int64, device=dev)
    t_ev = torch.zeros(R, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()  # the director launches on its own stream: order after torch's copies
    d.step_list_change_device(len(batch), t_off.data_ptr(), t_rows.data_ptr(), ForageParams(0, 0, 0), 0, 0, 0, 0,
                              t_idx.data_ptr(), t_best.data_ptr(), t_ev.data_ptr())
    d.synchronize()
    assert t_idx.cpu().numpy().view(np.uint32).tolist() == idx.tolist()
    # a completely empty batch is a step with no winner, not an error
    empty = np.zeros(R + 1, dtype=np.uint64)
    t_off0 = torch.from_numpy(empty.view(np.int64)).to(dev)
    torch.cuda.synchronize()
    d.step_list_change_device(0, t_off0.data_ptr(), t_rows.data_ptr(), ForageParams(0, 0, 0), 0, 0, 0, 0,
                              t_idx.data_ptr(), t_best.data_ptr(), t_ev.data_ptr())
    d.synchronize()
    assert (t_idx.cpu().numpy().view(np.uint32) == 0xFFFFFFFF).all()
    s0, ok0 = d.score_list_change(np.zeros((0, 4), dtype=np.uint32), empty)
    assert s0.shape == (0, 2)


def test_scalar_extremes():
    # one value only; everything unassigned; fewer entities than a warp
    g = instances.graph_coloring(7, 9, 1, seed_edges=2, seed_colors=3)
    for colors in (np.full(7, -1, np.int32), np.zeros(7, np.int32), np.array([0, -1, 0, -1, 0, 0, -1], np.int32)):
        o = Oracle.graph_coloring(g, colors)
        d = models.graph_coloring_director(g, colors=colors)
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
        rows = o.enumerate_change()
        so, oko = o.score_change(rows)
        s, ok = d.score_change(rows)
        assert np.array_equal(ok, oko) and np.array_equal(s, so)
        idx, best, ev, win = d.step_change(ForageParams(0, 0, 0))
        assert int(ev[0]) == len(rows)
        out = oracle_lib.replay_step(so, oko, [0, 0], [0, 0], [0, 0], 0, 2, 1, False, 3)
        if out[0]:
            assert int(idx[0]) == out[1]
        else:
            assert idx[0] == 0xFFFFFFFF


@pytest.mark.parametrize("scale,demand_scale,env", [
    (1, 1, {}), (100, 1, {}), (150000, 1, {}), (1, 3000, {}), (1, 1, {"SFGPU_NO_COMPACT": "1"}),
    (1, 1, {"SFGPU_NO_RELABEL": "1"})])
def test_arithmetic_width_paths_and_far_acceptor_references(scale, demand_scale, env, monkeypatch):
    """scale 1: uint16 cells + 32-bit deltas + compact 8-byte records; 100: int32 cells + 64-bit deltas
    (cost >= 65536); 150000: costs near the 2^28 fast-path limit; demand_scale 3000: values beyond int16
    (16-byte records with 32-bit arithmetic); env switches the compact records / the internal relabelling
    off. Acceptor references far outside the int32 delta range exercise the clamped relative thresholds."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    c = instances.cvrp(90, 7, seed=21)
    c.matrix = c.matrix * scale
    c.demands = (c.demands.astype(np.int64) * demand_scale).astype(np.int32)
    c.capacity = int(c.capacity) * demand_scale
    R, K = 2, 20
    starts = [instances.perturb_routes(c, 5 + r, 30) for r in range(R)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    base = d.calculate_score()
    for r in range(R):
        assert base[r].tolist() == oracles[r].committed_score().tolist()
    gen = []
    for r in range(R):
        rows = oracles[r].enumerate_nearby_list_change(K)
        so, oko = oracles[r].score_list_change(rows)
        gen.append((rows, so, oko))
    # materialised scoring path
    allrows = np.concatenate([g[0] for g in gen])
    offs = np.cumsum([0] + [len(g[0]) for g in gen]).astype(np.uint64)
    s, ok = d.score_list_change(allrows, offs)
    assert np.array_equal(s, np.concatenate([g[1] for g in gen])) and ok.all()
    far = 1 << 45
    refs = [np.array([0, -far, 0, -far]), np.array([0, far, 0, far]), np.array([-3, far, 3, -far]),
            np.array([0, -7 * scale, 0, -19 * scale]), np.array([far, 0, -far, 0])]
    for ref_d in refs:
        for acceptor, okind in ((1, 0), (2, 1)):
            ref = np.stack([np.concatenate([base[r], base[r]]) + ref_d for r in range(R)])
            idx, best, ev, win = d.step_nearby_list_change(K, ForageParams(acceptor, 1, 0), step_seeds=[3] * R,
                                                           ref_scores=ref)
            for r in range(R):
                rows, so, oko = gen[r]
                out = oracle_lib.replay_step(so, oko, [0, 0], ref[r][:2], ref[r][2:], 3, 2, 1, True, okind)
                what = f"scale={scale} ref={ref_d.tolist()} acc={acceptor} r={r}"
                if out[0]:
                    assert int(idx[r]) == out[1], what
                    assert best[r].tolist() == so[out[1]].tolist(), what
                    assert win[r].tolist() == rows[out[1]].tolist(), what
                else:
                    assert idx[r] == 0xFFFFFFFF, what


def test_sync_best_single_rank_is_the_lexicographic_max_over_replicas():
    """sfgpu_sync_best without a communicator: device reduction over the replicas' committed scores (or a given
    score array), first replica on ties — the N = 1 leg of the best-score sync (the NCCL leg runs in bench.py)."""
    import torch
    c = instances.cvrp(60, 5, seed=3)
    R = 7
    starts = [instances.perturb_routes(c, 40 + r, 20 + 5 * r) for r in range(R)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    sc = d.calculate_score()
    want = max(range(R), key=lambda r: (sc[r][0], sc[r][1], -r))
    best, owner, rep = d.sync_best()
    assert best == (int(sc[want][0]), int(sc[want][1])) and owner == 0 and rep == want
    # an explicit device score array, far outside the packed key's range, with a tie (first replica wins)
    arr = np.array([[-(1 << 40), 3], [5, -(1 << 45)], [5, -(1 << 45)], [4, 1 << 50], [5, -(1 << 45) - 1], [0, 0], [-1, 9]],
                   dtype=np.int64)
    t = torch.from_numpy(arr).cuda()
    torch.cuda.synchronize()
    best, owner, rep = d.sync_best(scores_ptr=t.data_ptr())
    assert best == (5, -(1 << 45)) and rep == 1


def test_thirty_two_constraints_like_the_reference_tuple_limit():
    """The reference's ConstraintSet tuples go up to 32 members (api/constraint_set/incremental.rs:339-408); so does a
    device model. 31 uni constraints with distinct weights + one grouped constraint; the 33rd is rejected loudly."""
    from solverforge_b200 import ConstraintFactory, Count, EqualVarToRow, GpuScoreDirector, HardSoftScore, soft
    from solverforge_b200 import _lib as L
    n, k = 40, 5
    rng = np.random.default_rng(3)
    d = GpuScoreDirector(2)
    vals = d.add_collection("values", k, -1)
    ents = d.add_collection("entities", n, 0)
    d.add_scalar_variable(ents, "v", k)
    cols = [rng.integers(0, 9, size=n) for _ in range(31)]
    f = ConstraintFactory(d)
    for i, c in enumerate(cols):
        f.for_each(ents).assigned().penalize(soft(L.W_LINEAR, i + 1, 0), x=d.add_column(ents, f"c{i}", c)).named(f"uni {i}")
    f.for_each(ents).join(f.for_each(vals), EqualVarToRow()).group_by(Count()).penalize(soft(L.W_SQUARE, 1, 0)).named("load")
    with pytest.raises(L.SfgpuError):
        f.for_each(ents).unassigned().penalize(HardSoftScore.ONE_HARD).named("one too many")
    start = rng.integers(-1, k, size=(2, n)).astype(np.int32)
    d.set_scalar_state(start)
    got = d.commit()

    def score(v):
        s = 0
        for i, c in enumerate(cols):
            s -= (i + 1) * int(c[v >= 0].sum())
        s -= int((np.bincount(v[v >= 0], minlength=k) ** 2).sum())
        return s
    for r in range(2):
        assert got[r].tolist() == [0, score(start[r])]
    rows = np.array([[e, v] for e in range(n) for v in range(-1, k)], dtype=np.int64)
    sc, ok = d.score_change(np.concatenate([rows, rows]), np.array([0, len(rows), 2 * len(rows)], dtype=np.uint64))
    for r in range(2):
        for i, (e, v) in enumerate(rows):
            j = r * len(rows) + i
            if start[r][e] == v:
                assert ok[j] == 0
                continue
            after = start[r].copy()
            after[e] = v
            assert ok[j] == 1 and sc[j].tolist() == [0, score(after)]


def test_int32_delta_bound_switches_exactly_at_the_edge():
    """The fast list kernels run their deltas in int32 when a commit-time bound proves every delta fits
    (|w_dist| * 4 * max cell + |w_cap| * 2 * sum |demand| < 2^31 - 1). Weights just below, at and above that edge —
    where a wrapped 32-bit product would be off by 2^32 — all score like the oracle (weights are linear per level)."""
    c = instances.cvrp(80, 6, seed=23)
    R, K = 2, 20
    starts = [instances.perturb_routes(c, 9 + r, 60) for r in range(R)]
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    gen = []
    for r in range(R):
        rows = oracles[r].enumerate_nearby_list_change(K)
        so, oko = oracles[r].score_list_change(rows)
        gen.append((rows, so, oko))
    mx = int(c.matrix.max())
    tot = int(np.abs(c.demands.astype(np.int64)).sum())
    for w_cap in (1, 100000):
        edge = (2 ** 31 - 1 - w_cap * 2 * tot) // (4 * mx)      # largest distance weight with bound < 2^31 - 1 (or one above)
        for w_dist in (edge - 1, edge, edge + 1, edge + 2, 3 * edge, 1000 * edge):
            d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]),
                                     distance_weight=w_dist, capacity_weight=w_cap)
            base = d.calculate_score()
            for r in range(R):
                oc = oracles[r].committed_score()
                assert base[r].tolist() == [int(oc[0]) * w_cap, int(oc[1]) * w_dist]
            allrows = np.concatenate([g[0] for g in gen])
            offs = np.cumsum([0] + [len(g[0]) for g in gen]).astype(np.uint64)
            s, ok = d.score_list_change(allrows, offs)
            want = np.concatenate([g[1] for g in gen]) * np.array([w_cap, w_dist], dtype=np.int64)
            assert ok.all() and np.array_equal(s, want), f"w_dist={w_dist} w_cap={w_cap}"
            # the fused device-generated step (REDUX / 64-bit shuffle reductions) picks the oracle's winner
            idx, best, ev, win = d.step_nearby_list_change(K, ForageParams(0, 0, 0), step_seeds=[1] * R)
            for r in range(R):
                rows, so, oko = gen[r]
                sw = so * np.array([w_cap, w_dist], dtype=np.int64)
                out = oracle_lib.replay_step(sw, oko, [0, 0], [0, 0], [0, 0], 1, 2, 1, False, 3)
                assert out[0] and int(idx[r]) == out[1] and best[r].tolist() == sw[out[1]].tolist(), f"w_dist={w_dist}"
            del d
