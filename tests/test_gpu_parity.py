"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded
inputs, bit-exact (tolerance 0: every level is int64), plus the reference's own known-answer
vectors and size-independent properties at BASELINE.json's full sizes."""
import json
import os

import numpy as np
import pytest

from solverforge_b200 import (ConstraintFactory, Count, EqualId, EqualKey, EqualVarToRow, ForageParams,
                              GpuScoreDirector, HardSoftScore, Sum, instances, models, soft)
from solverforge_b200 import _lib as L
from tests import oracle_lib
from tests.oracle_lib import Oracle

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


def _eq(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} mismatches, first at {bad[0]}: gpu={a[tuple(bad[0])]} oracle={b[tuple(bad[0])]}")


# ------------------------------------------------------------------ reference known answers
def _scalar_director(values, n_values):
    d = GpuScoreDirector(1)
    d.add_collection("values", n_values, -1)
    ent = d.add_collection("entities", len(values), 0)
    d.add_scalar_variable(ent, "var", n_values, True)
    return d, ent


def test_reference_kats_uni_unassigned():
    for v in GOLDEN["uni_unassigned"]:
        d, ent = _scalar_director(v["values"], v["n_values"])
        ConstraintFactory(d).for_each(ent).unassigned().penalize(HardSoftScore.ONE_SOFT).named("Unassigned")
        d.set_scalar_state(v["values"])
        assert d.commit()[0].tolist() == [0, v["soft"]], v["cite"]
        if "change" in v:
            s, ok = d.score_change(np.array([v["change"]]))
            assert ok[0] == 1 and s[0].tolist() == [0, v["soft_after"]], v["cite"]
            d.apply_change(np.array([v["change"]]))
            assert d.calculate_score()[0].tolist() == [0, v["soft_after"]]
            assert d.fresh_score()[0].tolist() == [0, v["soft_after"]]


def test_reference_kats_grouped():
    for v in GOLDEN["grouped_count"]:
        d, ent = _scalar_director(v["employee"], v["n_values"])
        w = soft(L.W_SQUARE if v["weight"] == "square" else L.W_LINEAR, 1, 0)
        g = ConstraintFactory(d).for_each(ent).group_by(Count())
        (g.penalize(w) if v["impact"] == "penalty" else g.reward(w)).named("Workload")
        d.set_scalar_state(v["employee"])
        assert d.commit()[0].tolist() == [0, v["soft"]], v["cite"]
    for v in GOLDEN["cross_grouped_sum"]:
        d, ent = _scalar_director(v["employee"], v["n_values"])
        ones = d.add_column(ent, "one", np.ones(len(v["employee"])))
        ConstraintFactory(d).for_each(ent).join(0, EqualVarToRow()).group_by(Sum(ones)).penalize(
            soft(L.W_SQUARE, 1, 0)).named("grouped assigned shift count")
        d.set_scalar_state(v["employee"])
        assert d.commit()[0].tolist() == [0, v["soft"]], v["cite"]
        s, ok = d.score_change(np.array([v["change"]]))
        assert s[0].tolist() == [0, v["soft_after"]], v["cite"]
    for v in GOLDEN["cross_complemented_grouped"]:
        d, ent = _scalar_director(v["employee"], v["n_values"])
        ones = d.add_column(ent, "one", np.ones(len(v["employee"])))
        ConstraintFactory(d).for_each(ent).join(0, EqualVarToRow()).group_by(Sum(ones)).complement(
            0, v["default"]).penalize(soft(L.W_LINEAR, 1, 0)).named("complemented cross grouped shift count")
        d.set_scalar_state(v["employee"])
        assert d.commit()[0].tolist() == [0, v["soft"]], v["cite"]
        s, ok = d.score_change(np.array([v["change"]]))
        assert s[0].tolist() == [0, v["soft_after"]], v["cite"]
        d.apply_change(np.array([v["change"]]))
        assert d.fresh_score()[0].tolist() == [0, v["soft_after"]]


def test_reference_kats_projected_grouped_complement():
    # a single-emit projection keyed by the planning variable lowers to GROUP with a per-key weight offset
    for v in GOLDEN["projected_grouped_complement"]:
        for demand, want in ((v["demand"], v["soft"]), (v["demand_after"], v["soft_after"])):
            d = GpuScoreDirector(1)
            buckets = d.add_collection("capacity", v["n_values"], -1)
            work = d.add_collection("work", len(v["bucket"]), 0)
            d.add_scalar_variable(work, "bucket", v["n_values"], True)
            dem = d.add_column(work, "demand", demand)
            off = d.add_column(buckets, "bucket_x10", v["key_offset"])
            ConstraintFactory(d).for_each(work).join(buckets, EqualVarToRow()).group_by(Sum(dem)).complement(
                buckets, v["default"]).penalize(soft(L.W_LINEAR, 1, 0), key_offset_column=off).named(
                "projected demand by capacity bucket")
            d.set_scalar_state(v["bucket"])
            assert d.commit()[0].tolist() == [0, want], v["cite"]
        # moving the only work row to bucket 1: bucket 0 falls back to its default
        s, ok = d.score_change(np.array([[0, 1]]))
        assert ok[0] == 1 and s[0].tolist() == [0, -(0 + 3) - (10 + 7)]
        d.apply_change(np.array([[0, 1]]))
        assert d.fresh_score()[0].tolist() == [0, -20]


def test_reference_kats_exists_and_self_join():
    for v in GOLDEN["exists_flattened"]:
        for routes, want in ((v["routes_before"], v["soft_before"]), (v["routes_after"], v["soft_after"])):
            d = GpuScoreDirector(1)
            loc = d.add_collection("locations", v["n_locations"], -1)
            cust = d.add_collection("customers", len(v["customer_ids"]), -1)
            rts = d.add_collection("routes", len(routes), 0)
            d.add_list_variable(rts, loc, "visits")
            cid = d.add_column(cust, "id", v["customer_ids"])
            f = ConstraintFactory(d)
            f.for_each(cust).if_not_exists(f.for_each(rts).flattened(), EqualId(cid)).penalize(
                HardSoftScore.ONE_SOFT).named("missing assignment")
            offs = np.cumsum([0] + [len(r) for r in routes])
            d.set_list_state(offs, np.array([x for r in routes for x in r], dtype=np.uint32))
            assert d.commit()[0].tolist() == [0, want], v["cite"]
    for v in GOLDEN["self_join_bi"]:
        d, ent = _scalar_director(v["row"], v["n_values"])
        f = ConstraintFactory(d)
        f.for_each(ent).join(f.for_each(ent), EqualKey()).penalize(HardSoftScore.ONE_SOFT).named("Row conflict")
        d.set_scalar_state(v["row"])
        assert d.commit()[0].tolist() == [0, v["soft"]], v["cite"]


# ------------------------------------------------------------------ config models vs oracle
def test_nqueens_full_neighbourhood_matches_oracle():
    for inst in (instances.nqueens(64), instances.nqueens(64, seed=1), instances.nqueens(8, seed=3)):
        o = Oracle.nqueens(inst)
        d = models.nqueens_director(inst)
        _eq(d.calculate_score()[0], o.committed_score(), "initial")
        rows = o.enumerate_change()
        s, ok = d.score_change(rows)
        so, oko = o.score_change(rows)
        _eq(ok, oko, "doable")
        _eq(s, so, "scores")


def test_graph_coloring_matches_oracle_and_apply_chain():
    g = instances.graph_coloring(2000, 9000, 6, seed_edges=5, seed_colors=6, unassigned_permille=30)
    o = Oracle.graph_coloring(g)
    d = models.graph_coloring_director(g)
    _eq(d.calculate_score()[0], o.committed_score(), "initial")
    rows = o.enumerate_change()
    s, ok = d.score_change(rows)
    so, oko = o.score_change(rows)
    _eq(ok, oko, "doable")
    _eq(s, so, "scores")
    # commit a chain of winners; committed == fresh == oracle after every step (FullAssert)
    for step in range(6):
        idx, best, ev = d.argbest(s, ok, params=ForageParams(0, 1, 0), step_seeds=[1000 + step])
        w = int(idx[0])
        assert ev[0] == len(rows)
        out = oracle_lib.replay_step(so, oko, [0, 0], [0, 0], [0, 0], 1000 + step, 2, 0, True, 3)
        assert out[0] == 1 and out[1] == w, "winner differs from the oracle's forager replay"
        d.apply_change(rows[w][None, :])
        o.apply_change(*rows[w])
        _eq(d.calculate_score()[0], o.committed_score(), "committed")
        _eq(d.fresh_score()[0], o.evaluate_all(), "fresh")
        _eq(best[0], o.committed_score(), "winner score")
        rows = o.enumerate_change()
        s, ok = d.score_change(rows)
        so, oko = o.score_change(rows)
        _eq(s, so, f"scores after step {step}")
        _eq(ok, oko)
    _eq(d.scalar_state()[0][:50], [int(x) for x in d.scalar_state()[0][:50]])


def test_graph_coloring_swap_and_compound_match_oracle():
    g = instances.graph_coloring(600, 3000, 5, seed_edges=9, seed_colors=10, unassigned_permille=40)
    o = Oracle.graph_coloring(g)
    d = models.graph_coloring_director(g)
    r = instances.splitmix64_stream(77, 4000)
    swaps = np.stack([r[:1000] % np.uint64(g.n), r[1000:2000] % np.uint64(g.n)], axis=1).astype(np.int64)
    s, ok = d.score_swap(swaps)
    so, oko = o.score_swap(swaps)
    _eq(ok, oko, "swap doable")
    _eq(s, so, "swap scores")
    # compound moves of 1..4 edits, including repeated entities and neighbours edited together
    n_c = 500
    sizes = (r[2000:2000 + n_c] % np.uint64(4)).astype(np.int64) + 1
    eo = np.concatenate([[0], np.cumsum(sizes)])
    tot = int(eo[-1])
    rr = instances.splitmix64_stream(78, 2 * tot)
    ent = (rr[:tot] % np.uint64(g.n)).astype(np.int64)
    # bias towards neighbours so overlays matter
    for i in range(1, tot, 3):
        lo, hi = g.row_ptr[ent[i - 1]], g.row_ptr[ent[i - 1] + 1]
        if hi > lo:
            ent[i] = g.col[lo]
    val = (rr[tot:] % np.uint64(g.k + 1)).astype(np.int64) - 1
    edits = np.stack([ent, val], axis=1)
    s, ok = d.score_compound(eo, edits)
    so, oko = o.score_compound(eo, edits)
    _eq(ok, oko, "compound doable")
    _eq(s, so, "compound scores")
    # improvement gates of evaluate_candidate (evaluation.rs:76-111) on the compound batch: every third move
    # requires hard improvement, every fifth score improvement; replay on device vs the oracle loop
    gates = np.zeros(n_c, dtype=np.uint8)
    gates[::3] |= 1
    gates[::5] |= 2
    base = d.calculate_score()[0]
    for acceptor, okind in ((0, 3), (1, 0), (2, 1)):
        for limit in (0, 3, 40):
            for dl in (0, -2):
                last = base + [dl, 0]
                late = base + [dl - 1, 0]
                ref = [np.concatenate([last, late])]
                idx, best, ev = d.argbest(s, ok, None, ForageParams(acceptor, 1, limit), [21], ref, gates=gates)
                out = oracle_lib.replay_step(so, oko, [0, 0], last, late, 21, 0 if limit else 2, max(limit, 1), True,
                                             okind, gates=gates)
                what = f"gated acceptor={acceptor} limit={limit} dl={dl}"
                assert int(ev[0]) == out[2], what
                if out[0]:
                    assert int(idx[0]) == out[1], what
                    assert best[0].tolist() == so[out[1]].tolist(), what
                else:
                    assert idx[0] == 0xFFFFFFFF, what


def test_job_shop_matches_oracle():
    j = instances.job_shop(40, 10, 6, seed=11, unassigned_permille=30)
    for with_c in (True, False):
        o = Oracle.job_shop(j, with_complement=with_c)
        d = models.job_shop_director(j, with_complement=with_c)
        _eq(d.calculate_score()[0], o.committed_score(), "initial")
        rows = o.enumerate_change()
        s, ok = d.score_change(rows)
        so, oko = o.score_change(rows)
        _eq(ok, oko)
        _eq(s, so, "change scores")
        r = instances.splitmix64_stream(5, 600)
        swaps = np.stack([r[:300] % np.uint64(j.n_ops), r[300:] % np.uint64(j.n_ops)], axis=1).astype(np.int64)
        s2, ok2 = d.score_swap(swaps)
        so2, oko2 = o.score_swap(swaps)
        _eq(ok2, oko2)
        _eq(s2, so2, "swap scores")
        for i in (3, 100, 777):
            if ok[i]:
                d.apply_change(rows[i][None, :])
                o.apply_change(*rows[i])
                _eq(d.calculate_score()[0], o.committed_score())
                _eq(d.fresh_score()[0], o.evaluate_all())
        # list moves on machine_sequences only touch the flattened exists constraint (delta 0)
        lrows = np.array([[0, 0, 1, 0], [2, 1, 2, 3], [1, 0, 1, 1]])
        s3, ok3 = d.score_list_change(lrows)
        so3, oko3 = o.score_list_change(lrows)
        _eq(ok3, oko3)
        _eq(s3, so3, "list change on job shop")


def test_shift_scheduling_load_balance_and_masked_uni_match_oracle():
    # examples/minimal-shift-scheduling + load_balance: unfairness = round(sqrt(f64)) must be bit-identical
    for seed in (21, 22, 23):
        inst = instances.shift_scheduling(seed=seed)
        o = Oracle.shift_scheduling(inst)
        d = models.shift_scheduling_director(inst)
        _eq(d.calculate_score()[0], o.committed_score(), "initial")
        for step in range(8):
            rows = o.enumerate_change()
            s, ok = d.score_change(rows)
            so, oko = o.score_change(rows)
            _eq(ok, oko, "doable")
            _eq(s, so, f"scores step {step}")
            idx, best, _ = d.argbest(s, ok, params=ForageParams(0, 1, 0), step_seeds=[seed * 100 + step])
            w = int(idx[0])
            d.apply_change(rows[w][None, :])
            o.apply_change(*rows[w])
            _eq(d.calculate_score()[0], o.committed_score(), "committed")
            _eq(d.fresh_score()[0], o.evaluate_all(), "fresh")
        with pytest.raises(L.SfgpuError) as e:
            d.score_swap(np.array([[0, 1]]))
        assert e.value.code == L.E_UNSUPPORTED


def test_cvrp_matches_oracle_lists_swaps_and_apply():
    c = instances.cvrp(200, 12, seed=7)
    o = Oracle.cvrp(c)
    d = models.cvrp_director(c)
    _eq(d.calculate_score()[0], o.committed_score(), "initial")
    for step in range(5):
        rows = o.enumerate_nearby_list_change(20)
        # plus ragged / not-doable rows (out-of-range positions, intra no-ops, empty targets)
        extra = np.array([[0, 0, 0, 0], [0, 0, 0, 1], [0, 999, 1, 0], [1, 0, 2, 999], [3, 1, 3, 0], [3, 0, 3, 2]],
                         dtype=np.uint32)
        rows = np.concatenate([rows, extra])
        s, ok = d.score_list_change(rows)
        so, oko = o.score_list_change(rows)
        _eq(ok, oko, "doable")
        _eq(s, so, "list change scores")
        idx, best, _ = d.argbest(s, ok, params=ForageParams(0, 0, 0))
        w = int(idx[0])
        out = oracle_lib.replay_step(so, oko, [0, 0], [0, 0], [0, 0], 0, 2, 0, False, 3)
        assert out[0] == 1 and out[1] == w, "First tie-break winner differs from the oracle replay"
        d.apply_list_change(rows[w][None, :])
        o.apply_list_change(*rows[w])
        _eq(d.calculate_score()[0], o.committed_score(), "committed")
        _eq(d.fresh_score()[0], o.evaluate_all(), "fresh")
    r = instances.splitmix64_stream(3, 4000)
    sw = np.stack([r[:1000] % np.uint64(12), r[1000:2000] % np.uint64(20), r[2000:3000] % np.uint64(12),
                   r[3000:] % np.uint64(20)], axis=1).astype(np.uint32)
    sw[::7, 2] = sw[::7, 0]            # same-route swaps
    sw[::14, 3] = sw[::14, 1] + 1      # adjacent positions
    s, ok = d.score_list_swap(sw)
    so, oko = o.score_list_swap(sw)
    _eq(ok, oko, "swap doable")
    _eq(s, so, "list swap scores")
    for i in np.flatnonzero(ok)[:4]:
        cur, okc = d.score_list_swap(sw[i][None, :])
        if not okc[0]:
            continue
        d.apply_list_swap(sw[i][None, :])
        o.apply_list_swap(*sw[i])
        _eq(d.calculate_score()[0], o.committed_score(), "swap committed")
        _eq(d.fresh_score()[0], o.evaluate_all())
    offs, elems = d.list_state()
    assert offs[0][-1] == 200 and sorted(elems[0][:200].tolist()) == list(range(1, 201))


def test_cvrp_empty_routes_and_unreachable_cells():
    c = instances.cvrp(30, 6, seed=2)
    c.matrix = c.matrix.copy()
    c.matrix[3, 7] = np.iinfo(np.int64).max   # UNREACHABLE
    c.matrix[5, 9] = -4                       # negative => MAX_SAFE_LEG_COST
    routes = [[1, 2, 3, 7, 9], [], [4, 5, 9 + 1], [], list(range(11, 31)), [6, 8]]
    routes[2] = [4, 5, 10]
    offs = np.cumsum([0] + [len(r) for r in routes]).astype(np.uint32)
    elems = np.array([x for r in routes for x in r], dtype=np.uint32)
    o = Oracle.cvrp(c, offs, elems)
    d = models.cvrp_director(c, offsets=offs, elems=elems)
    _eq(d.calculate_score()[0], o.committed_score(), "initial with unreachable legs")
    rows = np.array([[0, 0, 1, 0], [5, 0, 1, 0], [5, 1, 3, 0], [0, 2, 0, 5], [0, 3, 2, 1], [2, 1, 0, 4], [4, 0, 1, 0],
                     [1, 0, 0, 0], [0, 4, 0, 0]], dtype=np.uint32)
    s, ok = d.score_list_change(rows)
    so, oko = o.score_list_change(rows)
    _eq(ok, oko)
    _eq(s, so, "empty route / unreachable scores")
    # empty a route completely, then refill it
    for mv in ([5, 0, 1, 0], [5, 0, 1, 1], [1, 0, 5, 0]):
        d.apply_list_change(np.array([mv], dtype=np.uint32))
        o.apply_list_change(*mv)
        _eq(d.calculate_score()[0], o.committed_score())
        _eq(d.fresh_score()[0], o.evaluate_all())


def test_hard_soft_decimal_score_is_exact_on_device():
    # HardSoftDecimalScore = the same int64 pair pre-scaled by 100000 (hard_soft_decimal.rs:14,45-48):
    # authoring the weights scaled must give exactly 100000 x the HardSoftScore result, candidate by candidate
    from solverforge_b200 import HardSoftDecimalScore
    g = instances.graph_coloring(500, 2200, 5, seed_edges=12, seed_colors=13, unassigned_permille=60)
    plain = models.graph_coloring_director(g)
    dec = GpuScoreDirector(1)
    dec.add_collection("colors", g.k, -1)
    nodes = dec.add_collection("nodes", g.n, 0)
    dec.add_scalar_variable(nodes, "color_idx", g.k, True)
    adj = dec.add_csr("neighbors", g.row_ptr, g.col)
    from solverforge_b200 import AdjacentEqual
    f = ConstraintFactory(dec)
    one_hard = HardSoftDecimalScore.of(1, 0)
    assert (one_hard.hard, one_hard.soft) == (100000, 0)
    f.for_each(nodes).unassigned().penalize(one_hard).named("Unassigned color")
    f.for_each(nodes).join(f.for_each(nodes), AdjacentEqual(adj)).penalize(one_hard).named("Adjacent color conflict")
    dec.set_scalar_state(g.color)
    assert np.array_equal(dec.commit()[0], plain.calculate_score()[0] * 100000)
    rows = instances.change_neighbourhood(g.color, g.k)
    sp, okp = plain.score_change(rows)
    sd, okd = dec.score_change(rows)
    _eq(okd, okp)
    _eq(sd, sp * 100000, "decimal scores")


def test_fast_list_kernel_equals_generic_kernel():
    c = instances.cvrp(300, 20, seed=4)
    offs, el = instances.perturb_routes(c, 9, 200)
    o = Oracle.cvrp(c, offs, el)
    rows = o.enumerate_nearby_list_change(20)
    r = instances.splitmix64_stream(8, 8000)
    rnd = np.stack([r[:2000] % np.uint64(20), r[2000:4000] % np.uint64(40), r[4000:6000] % np.uint64(20),
                    r[6000:] % np.uint64(40)], axis=1).astype(np.uint32)      # many not-doable / out-of-range rows
    rows = np.concatenate([rows, rnd])
    fast = models.cvrp_director(c, offsets=offs, elems=el)
    gen = models.cvrp_director(c, offsets=offs, elems=el, flags=L.CTX_GENERIC_KERNELS)
    sf, okf = fast.score_list_change(rows)
    sg, okg = gen.score_list_change(rows)
    so, oko = o.score_list_change(rows)
    _eq(okf, okg, "fast vs generic doable")
    _eq(sf, sg, "fast vs generic scores")
    _eq(okf, oko)
    _eq(sf, so, "fast vs oracle")
    for i in np.flatnonzero(okf)[[5, 900, 4000]]:
        fast.apply_list_change(rows[i][None, :])
        gen.apply_list_change(rows[i][None, :])
        o.apply_list_change(*rows[i])
        sf, okf = fast.score_list_change(rows)
        so, oko = o.score_list_change(rows)
        _eq(sf, so, "fast after apply")
        _eq(okf, oko)
    _eq(fast.calculate_score(), gen.calculate_score())


def test_fused_step_matches_unfused_and_oracle_replay():
    import torch
    c = instances.cvrp(150, 10, seed=5)
    c.matrix = (c.matrix // 40) * 40          # coarse distances => many equal scores => tie rule matters
    R = 4
    starts = [instances.perturb_routes(c, 50 + r, 40) for r in range(R)]
    d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    batches = [oracles[r].enumerate_nearby_list_change(12) for r in range(R)]
    co = np.concatenate([[0], np.cumsum([len(b) for b in batches])]).astype(np.uint64)
    rows = np.concatenate(batches).astype(np.uint32)
    n = len(rows)
    dev = torch.device("cuda")
    t_off = torch.from_numpy(co.view(np.int64)).to(dev)
    t_rows = torch.from_numpy(rows.view(np.int32)).to(dev)
    t_scores = torch.zeros((n, 2), dtype=torch.int64, device=dev)
    t_doable = torch.zeros(n, dtype=torch.uint8, device=dev)
    t_idx = torch.zeros(R, dtype=torch.int32, device=dev)
    t_best = torch.zeros((R, 2), dtype=torch.int64, device=dev)
    t_ev = torch.zeros(R, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()  # the director launches on its own stream: order after torch's copies
    so = [oracles[r].score_list_change(batches[r]) for r in range(R)]
    base = d.calculate_score()
    for seed in (3, 4):
        for acceptor, okind in ((0, 3), (1, 0), (2, 1)):
            for ties in (0, 1):
                for limit in (0, 5):
                    for materialise in (True, False):
                        # reference scores: last step = committed score + a little slack, late = worse
                        ref = np.stack([np.concatenate([base[r] + [0, -30], base[r] + [0, -60]]) for r in range(R)])
                        t_ref = torch.from_numpy(ref).to(dev)
                        t_seed = torch.full((R,), seed, dtype=torch.int64, device=dev)
                        t_idx.fill_(-7)
                        torch.cuda.synchronize()
                        d.step_list_change_device(n, t_off.data_ptr(), t_rows.data_ptr(),
                                                  ForageParams(acceptor, ties, limit), t_seed.data_ptr(),
                                                  t_ref.data_ptr(), t_scores.data_ptr() if materialise else 0,
                                                  t_doable.data_ptr() if materialise else 0, t_idx.data_ptr(),
                                                  t_best.data_ptr(), t_ev.data_ptr())
                        d.synchronize()
                        idx = t_idx.cpu().numpy().view(np.uint32)
                        best = t_best.cpu().numpy()
                        for r in range(R):
                            out = oracle_lib.replay_step(so[r][0], so[r][1], [0, 0], ref[r][:2], ref[r][2:], seed,
                                                         0 if limit else 2, max(limit, 1), bool(ties), okind)
                            what = f"seed={seed} acc={acceptor} ties={ties} limit={limit} mat={materialise} r={r}"
                            if out[0]:
                                assert int(idx[r]) == out[1], what
                                assert best[r].tolist() == so[r][0][out[1]].tolist(), what
                            else:
                                assert idx[r] == 0xFFFFFFFF, what
                            assert int(t_ev.cpu()[r]) == out[2], what
                        if materialise:
                            sc = t_scores.cpu().numpy()
                            for r in range(R):
                                _eq(sc[int(co[r]):int(co[r + 1])], so[r][0], "materialised scores")
    # winners applied straight from the batch on device
    d.step_list_change_device(n, t_off.data_ptr(), t_rows.data_ptr(), ForageParams(0, 0, 0), 0, 0, 0, 0,
                              t_idx.data_ptr(), t_best.data_ptr(), t_ev.data_ptr())
    d.apply_winners_device(2, t_off.data_ptr(), t_rows.data_ptr(), t_idx.data_ptr())
    d.synchronize()
    idx = t_idx.cpu().numpy().view(np.uint32)
    after = d.calculate_score()
    for r in range(R):
        oracles[r].apply_list_change(*batches[r][int(idx[r])])
        _eq(after[r], oracles[r].committed_score(), "apply_winners")
    _eq(d.fresh_score(), after)


def test_replicas_are_independent():
    c = instances.cvrp(120, 8, seed=7)
    R = 5
    starts = [instances.perturb_routes(c, 100 + r, 30) for r in range(R)]
    offs = np.stack([s[0] for s in starts])
    elems = np.concatenate([s[1] for s in starts])
    d = models.cvrp_director(c, R, offsets=offs, elems=elems)
    oracles = [Oracle.cvrp(c, *starts[r]) for r in range(R)]
    init = d.calculate_score()
    for r in range(R):
        _eq(init[r], oracles[r].committed_score(), f"replica {r} initial")
    batches = [oracles[r].enumerate_nearby_list_change(5 + r) for r in range(R)]  # ragged batch sizes
    co = np.concatenate([[0], np.cumsum([len(b) for b in batches])]).astype(np.uint64)
    rows = np.concatenate(batches)
    s, ok = d.score_list_change(rows, co)
    for r in range(R):
        so, oko = oracles[r].score_list_change(batches[r])
        _eq(s[co[r]:co[r + 1]], so, f"replica {r} scores")
        _eq(ok[co[r]:co[r + 1]], oko)
    idx, best, ev = d.argbest(s, ok, co, ForageParams(0, 1, 0), step_seeds=np.arange(R) + 5)
    win = np.stack([rows[int(co[r]) + int(idx[r])] for r in range(R)])
    mask = np.array([1, 0, 1, 1, 0], dtype=np.uint8)
    d.apply_list_change(win, mask)
    after = d.calculate_score()
    for r in range(R):
        if mask[r]:
            oracles[r].apply_list_change(*win[r])
        _eq(after[r], oracles[r].committed_score(), f"replica {r} after masked apply")
    _eq(d.fresh_score(), after, "fresh == committed for every replica")


def test_argbest_replays_forager_and_acceptor_rules():
    rng = instances.splitmix64_stream(123, 6000)
    n = 3000
    hard = -(rng[:n] % np.uint64(3)).astype(np.int64)
    softv = -(rng[n:2 * n] % np.uint64(5)).astype(np.int64)       # many ties
    scores = np.stack([hard, softv], axis=1)
    doable = ((rng[:n] >> np.uint64(20)) % np.uint64(10) != 0).astype(np.uint8)
    g = instances.graph_coloring(50, 100, 3)
    d = models.graph_coloring_director(g)
    last, late = [-1, -2], [-1, -4]
    for seed in (1, 2, 3, 99):
        for acceptor, okind in ((0, 3), (1, 0), (2, 1)):
            for limit, fkind in ((0, 2), (1, 0), (7, 0), (256, 0)):
                for ties in (0, 1):
                    idx, best, ev = d.argbest(scores, doable, params=ForageParams(acceptor, ties, limit),
                                              step_seeds=[seed], ref_scores=[last + late])
                    out = oracle_lib.replay_step(scores, doable, [0, 0], last, late, seed, fkind, max(limit, 1),
                                                 bool(ties), okind)
                    what = f"seed={seed} acceptor={acceptor} limit={limit} ties={ties}"
                    if out[0]:
                        assert int(idx[0]) == out[1], what
                        assert best[0].tolist() == scores[out[1]].tolist(), what
                    else:
                        assert idx[0] == 0xFFFFFFFF, what
                    assert int(ev[0]) == out[2], what + " moves_evaluated"
    # nothing accepted at all
    idx, best, ev = d.argbest(scores, np.zeros(n, np.uint8))
    assert idx[0] == 0xFFFFFFFF


# ------------------------------------------------------------------ full BASELINE sizes
def test_full_size_c2_every_row_of_every_replica():
    """C2 at BASELINE size, no slices: every candidate of six differently coloured replicas against the O(degree)
    CPU checker (oracle/fast_cpu.cpp, itself pinned to the oracle in tests/test_oracle.py), a brute-force
    recompute of sampled candidates and the reference-faithful oracle on a window."""
    from tests.oracle_lib import FastGraphColoring
    g = instances.graph_coloring()            # 10 000 / 50 000 / k=8
    R = 6
    colors = np.stack([g.color] + [instances.graph_coloring_colors(g, 43 + r) for r in range(1, R)])
    d = models.graph_coloring_director(g, R, colors=colors)
    per = [instances.change_neighbourhood(colors[r], g.k) for r in range(R)]
    offs = np.concatenate([[0], np.cumsum([len(x) for x in per])]).astype(np.uint64)
    rows = np.concatenate(per)
    s, ok = d.score_change(rows, offs)
    assert len(per[0]) == 10_000 * 8 + int((g.color >= 0).sum())
    f = FastGraphColoring(g)
    base = d.calculate_score()
    _eq(base, d.fresh_score(), "committed == fresh")
    for r in range(R):
        sl = slice(int(offs[r]), int(offs[r + 1]))
        sf, okf, com = f.score_change(colors[r], per[r])
        _eq(ok[sl], okf, f"doable replica {r}")
        _eq(s[sl], sf, f"scores replica {r}")
        _eq(base[r], com, f"committed replica {r}")
        assert int((ok[sl] == 0).sum()) == int((colors[r] >= 0).sum())     # exactly the move-to-current-value rows
    # brute-force full recompute of sampled candidates (independent of every engine)
    def full(color):
        un = int((color < 0).sum())
        src = np.repeat(np.arange(g.n), np.diff(g.row_ptr.astype(np.int64)))
        dst = g.col.astype(np.int64)
        m = (src < dst) & (color[src] >= 0) & (color[src] == color[dst])
        return -(un + int(m.sum()))
    assert base[0][0] == full(g.color)
    for i in instances.splitmix64_stream(9, 60) % np.uint64(len(per[0])):
        e, v = per[0][int(i)]
        if ok[int(i)]:
            c2 = g.color.copy()
            c2[e] = v
            assert s[int(i)][0] == full(c2)
    # the reference-faithful engine (O(n) work per candidate) on a window
    o = Oracle.graph_coloring(g)
    sl = slice(40_000, 40_600)
    so, oko = o.score_change(per[0][sl])
    _eq(s[sl], so, "oracle slice")
    _eq(ok[sl], oko)


def test_full_size_c3_parity_and_round_trip():
    c = instances.cvrp()                      # 1000 customers / 80 vehicles
    o = Oracle.cvrp(c)
    d = models.cvrp_director(c)
    rows = o.enumerate_nearby_list_change(20)
    assert len(rows) == 20_000
    s, ok = d.score_list_change(rows)
    so, oko = o.score_list_change(rows)
    _eq(ok, oko)
    _eq(s, so, "CVRP-1000 neighbourhood")
    # move -> inverse move returns to the committed score and state
    base = d.calculate_score()[0].copy()
    offs0, el0 = d.list_state()
    mv = rows[np.flatnonzero(rows[:, 0] != rows[:, 2])[4321]]
    d.apply_list_change(mv[None, :])
    back = np.array([[mv[2], mv[3], mv[0], mv[1]]], dtype=np.uint32)
    d.apply_list_change(back)
    _eq(d.calculate_score()[0], base, "round trip score")
    offs1, el1 = d.list_state()
    _eq(offs1, offs0)
    _eq(el1, el0)


def test_full_size_c4_every_row_of_every_replica():
    """C4 at BASELINE size, no slices: all 84 000 candidates of six replicas with different machine assignments
    (a few operations unassigned) against the O(1) CPU checker, plus the oracle on a window."""
    from tests.oracle_lib import FastJobShop
    j = instances.job_shop()                  # 200 x 20 / 20 machines
    R = 6
    mach = np.stack([j.machine_idx] + [instances.job_shop_machines(j, 11 + r, unassigned_permille=5 * r) for r in range(1, R)])
    d = models.job_shop_director(j, R, machine_idx=mach)
    per = [instances.change_neighbourhood(mach[r], j.n_machines) for r in range(R)]
    assert len(per[0]) == 84_000
    offs = np.concatenate([[0], np.cumsum([len(x) for x in per])]).astype(np.uint64)
    s, ok = d.score_change(np.concatenate(per), offs)
    base = d.calculate_score()
    _eq(base, d.fresh_score())
    f = FastJobShop(j)
    for r in range(R):
        sl = slice(int(offs[r]), int(offs[r + 1]))
        sf, okf, com = f.score_change(mach[r], per[r])
        _eq(ok[sl], okf, f"doable replica {r}")
        _eq(s[sl], sf, f"scores replica {r}")
        _eq(base[r], com, f"committed replica {r}")
    o = Oracle.job_shop(j)
    _eq(base[0], o.committed_score())
    sl = slice(21_000, 21_400)
    so, oko = o.score_change(per[0][sl])
    _eq(s[sl], so, "job-shop slice")
    _eq(ok[sl], oko)


def test_projected_rows_roster_matches_oracle():
    """`.project(..)` multi-emit rows (stream/projected_stream/, constraint/projected/{uni.rs,grouped/}):
    grouped sum with an EXCESS weight, grouped count with a SQUARE weight, and a per-row terminal; change,
    swap and compound candidates, then committed moves with cached == fresh == oracle."""
    inst = instances.roster(160, 6, 9, seed=5)
    o = Oracle.roster(inst)
    d = models.roster_director(inst)
    assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
    assert d.fresh_score()[0].tolist() == o.evaluate_all().tolist()
    rows = o.enumerate_change()
    s, ok = d.score_change(rows)
    so, oko = o.score_change(rows)
    _eq(ok, oko, "projected change doable")
    _eq(s, so, "projected change scores")
    r = instances.splitmix64_stream(99, 3000)
    swaps = np.stack([r[:400] % np.uint64(inst.n_shifts), r[400:800] % np.uint64(inst.n_shifts)], axis=1).astype(np.int64)
    s, ok = d.score_swap(swaps)
    so, oko = o.score_swap(swaps)
    _eq(ok, oko, "projected swap doable")
    _eq(s, so, "projected swap scores")
    # compound candidates: several shifts moved at once, often onto the same (nurse, day) group
    n_c = 300
    sizes = (r[800:800 + n_c] % np.uint64(4)).astype(np.int64) + 1
    eo = np.concatenate([[0], np.cumsum(sizes)])
    tot = int(eo[-1])
    rr = instances.splitmix64_stream(98, 2 * tot)
    ent = (rr[:tot] % np.uint64(inst.n_shifts)).astype(np.int64)
    val = (rr[tot:] % np.uint64(3)).astype(np.int64) - 1            # few nurses => shared groups
    edits = np.stack([ent, val], axis=1)
    s, ok = d.score_compound(eo, edits)
    so, oko = o.score_compound(eo, edits)
    _eq(ok, oko, "projected compound doable")
    _eq(s, so, "projected compound scores")
    # a short hill-climbing trajectory on device, replayed on the oracle
    for step in range(8):
        last = d.calculate_score()
        idx, best, ev, win = d.step_change(ForageParams(1, 1, 0), step_seeds=[60 + step],
                                           ref_scores=np.concatenate([last, last], axis=1), apply=True)
        rows = o.enumerate_change()
        so, oko = o.score_change(rows)
        out = oracle_lib.replay_step(so, oko, [0, 0], last[0], last[0], 60 + step, 2, 1, True, 0)
        if out[0]:
            assert int(idx[0]) == out[1], f"step {step}"
            o.apply_change(*rows[out[1]])
        else:
            assert idx[0] == 0xFFFFFFFF
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
        assert d.fresh_score()[0].tolist() == o.committed_score().tolist()


@pytest.mark.parametrize("asymmetric", [False, True])
def test_list_reverse_moves_match_oracle(asymmetric):
    """ListReverseMove (2-opt segment reversal, heuristic/move/list_kernel/reverse.rs): the whole reverse
    neighbourhood in the selector's order, not-doable rows, forager replay and committed winners."""
    from solverforge_b200 import selectors
    c = instances.cvrp(70, 6, seed=12)
    if asymmetric:   # inner legs of a reversed segment change cost when d(a,b) != d(b,a)
        r = instances.splitmix64_stream(5, c.dim * c.dim).reshape(c.dim, c.dim)
        c.matrix = c.matrix + (r % np.uint64(9)).astype(np.int64)
        np.fill_diagonal(c.matrix, 0)
    offs, el = instances.perturb_routes(c, 8, 35)
    o = Oracle.cvrp(c, offs, el)
    d = models.cvrp_director(c, 1, offsets=offs[None, :], elems=el)
    rows = selectors.list_reverse_rows(offs)
    assert np.array_equal(rows, o.enumerate_list_reverse())
    bad = np.array([[0, 1, 2, 0], [0, 0, 999, 0], [99, 0, 2, 0], [1, 3, 1, 0]], dtype=np.uint32)
    allrows = np.concatenate([rows, bad])
    s, ok = d.score_list_reverse(allrows)
    so, oko = o.score_list_reverse(allrows[:-2])          # the oracle indexes lists directly: keep entities valid
    _eq(ok[:-2], oko, "reverse doable")
    _eq(s[:-2], so, "reverse scores")
    assert ok[-4:].tolist() == [0, 0, 0, 0]
    for step in range(6):
        rows = o.enumerate_list_reverse()
        so, oko = o.score_list_reverse(rows)
        sg, okg = d.score_list_reverse(rows)
        _eq(sg, so, f"reverse scores step {step}")
        last = d.calculate_score()
        idx, best, ev = d.argbest(sg, okg, None, ForageParams(1, 1, 0), [70 + step], [np.concatenate([last[0], last[0]])])
        out = oracle_lib.replay_step(so, oko, [0, 0], last[0], last[0], 70 + step, 2, 1, True, 0)
        if not out[0]:
            assert idx[0] == 0xFFFFFFFF
            break
        assert int(idx[0]) == out[1]
        d.apply_list_reverse(rows[out[1]][None, :])
        o.apply_list_reverse(*rows[out[1]])
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
        assert d.fresh_score()[0].tolist() == o.committed_score().tolist()
        lo, le = d.list_state()
        assert best[0].tolist() == o.committed_score().tolist()


@pytest.mark.parametrize("asymmetric", [False, True])
def test_sublist_change_moves_match_oracle(asymmetric):
    """SublistChangeMove (heuristic/move/list_kernel/sublist_change.rs): the whole neighbourhood (sizes 1..=3) in
    the selector's order, intra-list moves in post-removal coordinates, emptied / empty routes, not-doable rows,
    forager replay and committed winners tracked against the oracle."""
    from solverforge_b200 import selectors
    c = instances.cvrp(60, 7, seed=14)
    if asymmetric:
        r = instances.splitmix64_stream(6, c.dim * c.dim).reshape(c.dim, c.dim)
        c.matrix = c.matrix + (r % np.uint64(9)).astype(np.int64)
        np.fill_diagonal(c.matrix, 0)
    offs, el = instances.perturb_routes(c, 9, 40)
    # one short route (emptied by a size-2 move) and one empty route
    lens = np.diff(offs).tolist()
    o = Oracle.cvrp(c, offs, el)
    d = models.cvrp_director(c, 1, offsets=offs[None, :], elems=el)
    rows = selectors.sublist_change_rows(offs, 1, 3)
    assert np.array_equal(rows, o.enumerate_sublist_change(1, 3))
    s, ok = d.score_sublist_change(rows)
    so, oko = o.score_sublist_change(rows)
    _eq(ok, oko, "sublist change doable")
    _eq(s, so, "sublist change scores")
    assert ok.all()
    # rows the reference rejects in is_doable (sublist_change.rs:17-49); entities stay valid for the oracle
    L0 = lens[0]
    bad = np.array([[0, 1, 1, 1, 0], [0, 0, L0 + 1, 1, 0], [0, 0, 2, 0, 0], [0, 0, 2, 0, L0 - 1], [0, 0, 1, 1, lens[1] + 1],
                    [0, 1, 3, 0, 1]], dtype=np.int64)
    sb, okb = d.score_sublist_change(bad)
    sob, okob = o.score_sublist_change(bad)
    _eq(okb, okob, "sublist change not-doable rows")
    assert okb.tolist() == [0, 0, 0, 0, 0, 0]
    sx, okx = d.score_sublist_change(np.array([[99, 0, 1, 0, 0], [0, 0, 1, 99, 0]]))
    assert okx.tolist() == [0, 0]
    for step in range(8):
        rows = o.enumerate_sublist_change(1, 3)
        so, oko = o.score_sublist_change(rows)
        sg, okg = d.score_sublist_change(rows)
        _eq(sg, so, f"sublist change scores step {step}")
        _eq(okg, oko, f"sublist change doable step {step}")
        last = d.calculate_score()
        idx, best, ev = d.argbest(sg, okg, None, ForageParams(1, 1, 0), [90 + step], [np.concatenate([last[0], last[0]])])
        out = oracle_lib.replay_step(so, oko, [0, 0], last[0], last[0], 90 + step, 2, 1, True, 0)
        if not out[0]:
            assert idx[0] == 0xFFFFFFFF
            break
        assert int(idx[0]) == out[1]
        d.apply_sublist_change(rows[out[1]][None, :])
        o.apply_sublist_change(*rows[out[1]])
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
        assert d.fresh_score()[0].tolist() == o.committed_score().tolist()
        assert best[0].tolist() == o.committed_score().tolist()
    # a forced inter-route relocation that empties a route, then one into the empty route
    lo, le = d.list_state()
    lens = np.diff(lo[0]).tolist()
    short = int(np.argmin([x if x > 0 else 10 ** 9 for x in lens]))
    if lens[short] <= 3:
        other = (short + 1) % len(lens)
        mv = np.array([[short, 0, lens[short], other, 0]])
        sg, okg = d.score_sublist_change(mv)
        so, oko = o.score_sublist_change(mv)
        _eq(sg, so, "emptying move score")
        d.apply_sublist_change(mv)
        o.apply_sublist_change(*mv[0])
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist() == d.fresh_score()[0].tolist()
        back = np.array([[other, 0, 2, short, 0]])
        sg, okg = d.score_sublist_change(back)
        so, oko = o.score_sublist_change(back)
        _eq(sg, so, "move into the empty route")
        d.apply_sublist_change(back)
        o.apply_sublist_change(*back[0])
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist() == d.fresh_score()[0].tolist()


@pytest.mark.parametrize("asymmetric", [False, True])
def test_sublist_swap_moves_match_oracle(asymmetric):
    """SublistSwapMove (heuristic/move/list_kernel/sublist_swap.rs): the whole neighbourhood (sizes 1..=3, unequal
    sizes, adjacent and distant segments inside one list, segments at route ends), rows given late-segment-first,
    not-doable rows, forager replay and committed winners tracked against the oracle."""
    from solverforge_b200 import selectors
    c = instances.cvrp(44, 6, seed=15)
    if asymmetric:
        r = instances.splitmix64_stream(7, c.dim * c.dim).reshape(c.dim, c.dim)
        c.matrix = c.matrix + (r % np.uint64(9)).astype(np.int64)
        np.fill_diagonal(c.matrix, 0)
    offs, el = instances.perturb_routes(c, 10, 30)
    lens = np.diff(offs).tolist()
    o = Oracle.cvrp(c, offs, el)
    d = models.cvrp_director(c, 1, offsets=offs[None, :], elems=el)
    rows = selectors.sublist_swap_rows(offs, 1, 3)
    assert np.array_equal(rows, o.enumerate_sublist_swap(1, 3))
    s, ok = d.score_sublist_swap(rows)
    so, oko = o.score_sublist_swap(rows)
    _eq(ok, oko, "sublist swap doable")
    _eq(s, so, "sublist swap scores")
    assert ok.all()
    # the same moves with the two segments exchanged in the row (later segment first): same scores
    flipped = rows[:, [3, 4, 5, 0, 1, 2]]
    sf, okf = d.score_sublist_swap(flipped)
    _eq(sf, so, "flipped sublist swap scores")
    sof, okof = o.score_sublist_swap(flipped[:2000])
    _eq(sof, so[:2000], "oracle agrees on flipped rows")
    L0 = lens[0]
    bad = np.array([[0, 1, 4, 0, 2, 5], [0, 1, 1, 0, 2, 3], [0, 0, 2, 0, 2, L0 + 1], [0, 0, 2, 1, 0, lens[1] + 1],
                    [0, 0, 3, 0, 2, 4], [0, 2, 4, 0, 2, 4]], dtype=np.int64)
    sb, okb = d.score_sublist_swap(bad)
    sob, okob = o.score_sublist_swap(bad)
    _eq(okb, okob, "sublist swap not-doable rows")
    assert okb.tolist() == [0, 0, 0, 0, 0, 0]
    assert d.score_sublist_swap(np.array([[99, 0, 1, 0, 0, 1], [0, 0, 1, 99, 0, 1]]))[1].tolist() == [0, 0]
    for step in range(8):
        rows = o.enumerate_sublist_swap(1, 3)
        so, oko = o.score_sublist_swap(rows)
        sg, okg = d.score_sublist_swap(rows)
        _eq(sg, so, f"sublist swap scores step {step}")
        last = d.calculate_score()
        idx, best, ev = d.argbest(sg, okg, None, ForageParams(1, 1, 0), [110 + step], [np.concatenate([last[0], last[0]])])
        out = oracle_lib.replay_step(so, oko, [0, 0], last[0], last[0], 110 + step, 2, 1, True, 0)
        if not out[0]:
            assert idx[0] == 0xFFFFFFFF
            break
        assert int(idx[0]) == out[1]
        mv = rows[out[1]] if step % 2 == 0 else rows[out[1]][[3, 4, 5, 0, 1, 2]]
        d.apply_sublist_swap(mv[None, :])
        o.apply_sublist_swap(*mv)
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
        assert d.fresh_score()[0].tolist() == o.committed_score().tolist()
        assert best[0].tolist() == o.committed_score().tolist()
    # forced unequal exchanges (not necessarily improving): inter-list 3 <-> 1 and intra-list 1 <-> 3 with a gap
    lo, le = d.list_state()
    lens = np.diff(lo[0]).tolist()
    big = [i for i, x in enumerate(lens) if x >= 6]
    for mv in ([big[0], 0, 3, big[1], lens[big[1]] - 1, lens[big[1]]], [big[0], 0, 1, big[0], 2, 5],
               [big[1], 3, 5, big[1], 0, 3]):
        mv = np.array(mv)
        sg, okg = d.score_sublist_swap(mv[None, :])
        so, oko = o.score_sublist_swap(mv[None, :])
        _eq(sg, so, f"forced swap {mv.tolist()}")
        assert okg[0] == 1
        d.apply_sublist_swap(mv[None, :])
        o.apply_sublist_swap(*mv)
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist() == d.fresh_score()[0].tolist()


@pytest.mark.parametrize("joins", [((3, 1),), ((2, 1), (3, 10), (4, 100), (5, 1000)), ((5, 7), (4, 3))])
def test_higher_arity_keyed_joins_match_oracle(joins):
    """Keyed tri / quad / penta self-joins (constraint/nary_incremental/higher_arity/shared.rs): C(n, arity) tuples
    per bucket. Reference KATs (tri_incr.rs:22-138 and the quad / penta twins) as columns, then the cluster model:
    full ChangeMove neighbourhood, swaps, compound moves with overlays, fused device step, committed winners."""
    for v in GOLDEN["higher_arity_joins"]:   # committed golden vectors of the reference's tri / quad / penta tests
        k = instances.ClusterInstance(len(v["team"]), v["n_values"], np.array(v["team"], dtype=np.int32), ((v["arity"], 1),))
        dk = models.cluster_director(k)
        assert dk.calculate_score()[0].tolist() == [0, v["soft"]] == dk.fresh_score()[0].tolist(), v["cite"]
        dk.apply_change(np.array([[0, -1]]))                      # retract row 0: its tuples go, one unassigned task
        assert dk.calculate_score()[0].tolist() == [-1, v["retract_row0_soft"]] == dk.fresh_score()[0].tolist(), v["cite"]
    c = instances.cluster(44, 4, seed=23, joins=joins)   # buckets of ~10 rows: the oracle materialises every tuple
    o = Oracle.cluster(c)
    d = models.cluster_director(c)
    assert d.scalar_program() == -1 or all(a == 2 for a, _ in joins)
    assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
    r = instances.splitmix64_stream(91, 6000)
    for step in range(5):
        rows = o.enumerate_change()
        so, oko = o.score_change(rows)
        s, ok = d.score_change(rows)
        _eq(ok, oko, f"change doable step {step}")
        _eq(s, so, f"change scores step {step}")
        swaps = np.stack([r[:200] % np.uint64(c.n), r[600:800] % np.uint64(c.n)], axis=1).astype(np.int64)
        s2, ok2 = d.score_swap(swaps)
        so2, oko2 = o.score_swap(swaps)
        _eq(ok2, oko2, "swap doable")
        _eq(s2, so2, "swap scores")
        sizes = (r[1200:1350] % np.uint64(4)).astype(np.int64) + 1
        eo = np.concatenate([[0], np.cumsum(sizes)])
        tot = int(eo[-1])
        ent = (r[1500:1500 + tot] % np.uint64(c.n // 3)).astype(np.int64)      # few entities: edits overlap
        val = (r[3000:3000 + tot] % np.uint64(c.n_teams + 1)).astype(np.int64) - 1
        s3, ok3 = d.score_compound(eo, np.stack([ent, val], axis=1))
        so3, oko3 = o.score_compound(eo, np.stack([ent, val], axis=1))
        _eq(ok3, oko3, "compound doable")
        _eq(s3, so3, "compound scores")
        last = d.calculate_score()
        idx, best, ev, win = d.step_change(ForageParams(0, 1, 0), step_seeds=[300 + step],
                                           ref_scores=np.concatenate([last, last], axis=1), apply=True)
        out = oracle_lib.replay_step(so, oko, [0, 0], last[0], last[0], 300 + step, 2, 1, True, 3)
        assert out[0] and int(idx[0]) == out[1]
        o.apply_change(int(rows[out[1]][0]), int(rows[out[1]][1]))
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist() == d.fresh_score()[0].tolist()
        r = instances.splitmix64_stream(92 + step, 6000)


def test_scalar_swap_neighbourhood_steps_match_oracle():
    """The whole SwapMoveSelector neighbourhood (move_selector/swap.rs, canonical and seeded orders) generated by the
    host selector, scored on device, winner by the device forager replay, committed; tracked against the oracle."""
    from solverforge_b200 import selectors
    from solverforge_b200.selectors import MoveStreamContext
    g = instances.graph_coloring(120, 420, 4, seed_edges=12, seed_colors=13, unassigned_permille=60)
    o = Oracle.graph_coloring(g)
    d = models.graph_coloring_director(g)
    for step, order in enumerate((selectors.ORIGINAL, selectors.SHUFFLED, selectors.RANDOM, selectors.ORIGINAL)):
        ctx = MoveStreamContext(step, 900 + step, order)
        rows = selectors.swap_move_rows(g.n, ctx)
        assert np.array_equal(rows, o.enumerate_swap(step, 900 + step, order))
        s, ok = d.score_swap(rows)
        so, oko = o.score_swap(rows)
        _eq(ok, oko, f"swap doable step {step}")
        _eq(s, so, f"swap scores step {step}")
        last = d.calculate_score()
        for limit in (0, 25):
            idx, best, ev = d.argbest(s, ok, None, ForageParams(1, 1, limit), [70 + step], [np.concatenate([last[0], last[0]])])
            out = oracle_lib.replay_step(so, oko, [0, 0], last[0], last[0], 70 + step, 0 if limit else 2, max(limit, 1), True, 0)
            assert int(ev[0]) == out[2]
            assert (int(idx[0]) == out[1]) if out[0] else idx[0] == 0xFFFFFFFF
        if not out[0]:
            break
        d.apply_swap(rows[out[1]][None, :])
        o.apply_swap(*rows[out[1]])
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist() == d.fresh_score()[0].tolist()


def test_consecutive_runs_collector_matches_oracle():
    """group_by(nurse, consecutive_runs(day)) — "Long work streaks" of examples/minimal-shift-scheduling
    (stream/collector/runs.rs): change, swap and compound candidates (several shifts of one candidate landing
    on adjacent days of one nurse), committed trajectories with cached == fresh == oracle."""
    for seed in (3, 8):
        inst = instances.shift_scheduling(n_days=12, slots_per_day=3, n_nurses=4, seed=seed, unassigned_permille=250)
        o = Oracle.shift_scheduling(inst, with_load_balance=False)
        d = models.shift_scheduling_director(inst, with_load_balance=False)
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
        rows = o.enumerate_change()
        s, ok = d.score_change(rows)
        so, oko = o.score_change(rows)
        _eq(ok, oko, "runs change doable")
        _eq(s, so, "runs change scores")
        r = instances.splitmix64_stream(40 + seed, 4000)
        swaps = np.stack([r[:300] % np.uint64(inst.n_shifts), r[300:600] % np.uint64(inst.n_shifts)], axis=1).astype(np.int64)
        s, ok = d.score_swap(swaps)
        so, oko = o.score_swap(swaps)
        _eq(ok, oko, "runs swap doable")
        _eq(s, so, "runs swap scores")
        n_c = 300
        sizes = (r[600:600 + n_c] % np.uint64(5)).astype(np.int64) + 1
        eo = np.concatenate([[0], np.cumsum(sizes)])
        tot = int(eo[-1])
        rr = instances.splitmix64_stream(41 + seed, 2 * tot)
        ent = (rr[:tot] % np.uint64(inst.n_shifts)).astype(np.int64)
        for i in range(1, tot, 2):                       # neighbouring days in one candidate
            ent[i] = min(int(ent[i - 1]) + 3, inst.n_shifts - 1)
        val = (rr[tot:] % np.uint64(3)).astype(np.int64) - 1
        edits = np.stack([ent, val], axis=1)
        s, ok = d.score_compound(eo, edits)
        so, oko = o.score_compound(eo, edits)
        _eq(ok, oko, "runs compound doable")
        _eq(s, so, "runs compound scores")
        for step in range(10):
            last = d.calculate_score()
            idx, best, ev, win = d.step_change(ForageParams(2, 1, 0), step_seeds=[90 + step],
                                               ref_scores=np.concatenate([last, last + [0, -3]], axis=1), apply=True)
            rows = o.enumerate_change()
            so, oko = o.score_change(rows)
            out = oracle_lib.replay_step(so, oko, [0, 0], last[0], last[0] + [0, -3], 90 + step, 2, 1, True, 1)
            if out[0]:
                assert int(idx[0]) == out[1], f"step {step}"
                o.apply_change(*rows[out[1]])
            assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
            assert d.fresh_score()[0].tolist() == o.committed_score().tolist()


def test_indexed_presence_collector_matches_oracle():
    """group_by(nurse, indexed_presence(day)) scored through complement_runs(0..H), any_in(5..7) and count()
    (stream/collector/indexed_presence.rs; the reference's macro example
    solverforge-macros/tests/ui/pass/solverforge_constraints_indexed_presence.rs): the known answer of that example
    as columns, then change, swap and compound candidates and committed trajectories, cached == fresh == oracle.
    Groups that lose their last item disappear (their complement runs are not scored)."""
    kat = instances.shift_scheduling(n_days=5, slots_per_day=1, n_nurses=2, seed=1)
    kat.n_shifts, kat.day, kat.slot = 3, kat.day[:3], kat.slot[:3]
    kat.required, kat.hours = np.zeros(3, dtype=np.uint8), kat.hours[:3]
    kat.nurse_idx, kat.target = np.zeros(3, dtype=np.int32), 0
    gv = GOLDEN["indexed_presence"][0]
    assert kat.day.tolist() == gv["days"] and kat.nurse_idx.tolist() == gv["nurse"]
    dk = models.shift_scheduling_director(kat, with_load_balance=False, presence_days=gv["horizon"])
    assert dk.calculate_score()[0].tolist() == [0, gv["model_soft"]] == dk.fresh_score()[0].tolist()
    dk.apply_change(np.array([gv["change"]]))
    assert dk.calculate_score()[0].tolist() == [0, gv["model_soft_after"]] == dk.fresh_score()[0].tolist()
    for seed in (3, 8):
        inst = instances.shift_scheduling(n_days=12, slots_per_day=3, n_nurses=4, seed=seed, unassigned_permille=250)
        o = Oracle.shift_scheduling(inst, with_load_balance=False, presence_days=12)
        d = models.shift_scheduling_director(inst, with_load_balance=False, presence_days=12)
        assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
        rows = o.enumerate_change()
        s, ok = d.score_change(rows)
        so, oko = o.score_change(rows)
        _eq(ok, oko, "presence change doable")
        _eq(s, so, "presence change scores")
        r = instances.splitmix64_stream(40 + seed, 4000)
        swaps = np.stack([r[:300] % np.uint64(inst.n_shifts), r[300:600] % np.uint64(inst.n_shifts)], axis=1).astype(np.int64)
        s, ok = d.score_swap(swaps)
        so, oko = o.score_swap(swaps)
        _eq(ok, oko, "presence swap doable")
        _eq(s, so, "presence swap scores")
        n_c = 300
        sizes = (r[600:600 + n_c] % np.uint64(5)).astype(np.int64) + 1
        eo = np.concatenate([[0], np.cumsum(sizes)])
        tot = int(eo[-1])
        rr = instances.splitmix64_stream(41 + seed, 2 * tot)
        ent = (rr[:tot] % np.uint64(inst.n_shifts)).astype(np.int64)
        for i in range(1, tot, 2):                       # neighbouring days in one candidate
            ent[i] = min(int(ent[i - 1]) + 3, inst.n_shifts - 1)
        val = (rr[tot:] % np.uint64(3)).astype(np.int64) - 1
        edits = np.stack([ent, val], axis=1)
        s, ok = d.score_compound(eo, edits)
        so, oko = o.score_compound(eo, edits)
        _eq(ok, oko, "presence compound doable")
        _eq(s, so, "presence compound scores")
        for step in range(10):
            last = d.calculate_score()
            idx, best, ev, win = d.step_change(ForageParams(2, 1, 0), step_seeds=[90 + step],
                                               ref_scores=np.concatenate([last, last + [0, -3]], axis=1), apply=True)
            rows = o.enumerate_change()
            so, oko = o.score_change(rows)
            out = oracle_lib.replay_step(so, oko, [0, 0], last[0], last[0] + [0, -3], 90 + step, 2, 1, True, 1)
            if out[0]:
                assert int(idx[0]) == out[1], f"step {step}"
                o.apply_change(*rows[out[1]])
            assert d.calculate_score()[0].tolist() == o.committed_score().tolist()
            assert d.fresh_score()[0].tolist() == o.committed_score().tolist()
