"""CPU tests: the product's host-side selectors reproduce the reference pull order (oracle
restatement of MoveStreamContext / ChangeMoveSelector / NearbyListChangeMoveSelector) for every
selection order."""
import numpy as np
import pytest

from solverforge_b200 import instances, selectors
from solverforge_b200.selectors import MoveStreamContext
from tests.oracle_lib import Oracle


@pytest.mark.parametrize("order", [selectors.ORIGINAL, selectors.RANDOM, selectors.SHUFFLED])
def test_change_move_order_matches_reference(order):
    g = instances.graph_coloring(97, 300, 5, seed_edges=3, seed_colors=4, unassigned_permille=100)
    o = Oracle.graph_coloring(g)
    for step_index, seed in ((0, 0), (3, 12345), (17, 0xDEADBEEFCAFE)):
        want = o.enumerate_change(step_index, seed, order)
        got = selectors.change_move_rows(g.color, g.k, True, MoveStreamContext(step_index, seed, order))
        assert np.array_equal(got, want)


@pytest.mark.parametrize("order", [selectors.ORIGINAL, selectors.RANDOM, selectors.SHUFFLED])
def test_nearby_list_change_order_matches_reference(order):
    c = instances.cvrp(70, 7, seed=5)
    c.matrix = (c.matrix // 30) * 30                   # ties exercise the stable top-k
    offs, el = instances.perturb_routes(c, 4, 40)
    o = Oracle.cvrp(c, offs, el)
    for step_index, seed in ((0, 0), (5, 987654321), (40, 0xABCDEF0123456789)):
        want = o.enumerate_nearby_list_change(9, step_index, seed, order)
        got = selectors.nearby_list_change_rows(offs, el, c.matrix, 9, MoveStreamContext(step_index, seed, order))
        assert np.array_equal(got, want), f"order={order} step={step_index}"


def test_nearby_handles_empty_routes_and_unreachable_cells():
    c = instances.cvrp(20, 5, seed=2)
    c.matrix = c.matrix.copy()
    c.matrix[3, 7] = np.iinfo(np.int64).max
    c.matrix[5, :] = -1
    routes = [[1, 2, 3, 7, 9], [], [4, 5, 10], [], list(range(11, 21)) + [6, 8]]
    offs = np.cumsum([0] + [len(r) for r in routes]).astype(np.uint32)
    el = np.array([x for r in routes for x in r], dtype=np.uint32)
    o = Oracle.cvrp(c, offs, el)
    assert np.array_equal(selectors.nearby_list_change_rows(offs, el, c.matrix, 6), o.enumerate_nearby_list_change(6))


@pytest.mark.parametrize("order", [selectors.ORIGINAL, selectors.RANDOM, selectors.SHUFFLED])
def test_nearby_list_swap_order_matches_reference(order):
    c = instances.cvrp(60, 6, seed=8)
    c.matrix = (c.matrix // 25) * 25
    offs, el = instances.perturb_routes(c, 6, 30)
    o = Oracle.cvrp(c, offs, el)
    for step_index, seed in ((0, 0), (2, 55555), (31, 0x123456789ABCDEF)):
        want = o.enumerate_nearby_list_swap(7, step_index, seed, order)
        got = selectors.nearby_list_swap_rows(offs, el, c.matrix, 7, MoveStreamContext(step_index, seed, order))
        assert np.array_equal(got, want), f"order={order} step={step_index}"


def test_nearby_swap_handles_empty_routes_and_unreachable_cells():
    c = instances.cvrp(20, 5, seed=2)
    c.matrix = c.matrix.copy()
    c.matrix[3, 7] = np.iinfo(np.int64).max
    c.matrix[5, :] = -1
    routes = [[1, 2, 3, 7, 9], [], [4, 5, 10], [], list(range(11, 21)) + [6, 8]]
    offs = np.cumsum([0] + [len(r) for r in routes]).astype(np.uint32)
    el = np.array([x for r in routes for x in r], dtype=np.uint32)
    o = Oracle.cvrp(c, offs, el)
    assert np.array_equal(selectors.nearby_list_swap_rows(offs, el, c.matrix, 6), o.enumerate_nearby_list_swap(6))


@pytest.mark.parametrize("order", [selectors.ORIGINAL, selectors.RANDOM, selectors.SHUFFLED])
def test_list_reverse_order_matches_reference(order):
    c = instances.cvrp(40, 6, seed=4)
    offs, el = instances.perturb_routes(c, 2, 25)
    o = Oracle.cvrp(c, offs, el)
    for step_index, seed in ((0, 0), (9, 4242), (77, 0xFEEDFACE12345)):
        want = o.enumerate_list_reverse(step_index, seed, order)
        got = selectors.list_reverse_rows(offs, MoveStreamContext(step_index, seed, order))
        assert np.array_equal(got, want), f"order={order} step={step_index}"


@pytest.mark.parametrize("order", [selectors.ORIGINAL, selectors.RANDOM, selectors.SHUFFLED])
def test_sublist_change_order_matches_reference(order):
    """SublistChangeMoveSelector pull order (list_kernel/sublist_change.rs:103-268) incl. routes shorter than the
    minimum segment and empty routes."""
    c = instances.cvrp(40, 7, seed=4)
    offs, el = instances.perturb_routes(c, 2, 25)
    o = Oracle.cvrp(c, offs, el)
    for (lo, hi) in ((1, 3), (2, 2), (3, 9)):
        for step_index, seed in ((0, 0), (9, 4242)):
            want = o.enumerate_sublist_change(lo, hi, step_index, seed, order)
            got = selectors.sublist_change_rows(offs, lo, hi, MoveStreamContext(step_index, seed, order))
            assert np.array_equal(got, want), f"order={order} step={step_index} sizes={lo}..{hi}"
    # reference KAT (selector/tests/sublist_neighborhood.rs:133-187)
    kat = selectors.sublist_change_rows(np.array([0, 3, 5]), 2, 2)
    assert kat.tolist() == [[0, 0, 2, 0, 1], [0, 0, 2, 1, 0], [0, 0, 2, 1, 1], [0, 0, 2, 1, 2], [0, 1, 3, 0, 0],
                            [0, 1, 3, 1, 0], [0, 1, 3, 1, 1], [0, 1, 3, 1, 2], [1, 0, 2, 0, 0], [1, 0, 2, 0, 1],
                            [1, 0, 2, 0, 2], [1, 0, 2, 0, 3]]


@pytest.mark.parametrize("order", [selectors.ORIGINAL, selectors.RANDOM, selectors.SHUFFLED])
def test_sublist_swap_order_matches_reference(order):
    """SublistSwapMoveSelector pull order (list_kernel/sublist_swap.rs:28-318)."""
    c = instances.cvrp(36, 6, seed=5)
    offs, el = instances.perturb_routes(c, 3, 25)
    o = Oracle.cvrp(c, offs, el)
    for (lo, hi) in ((1, 3), (2, 2), (4, 9)):
        for step_index, seed in ((0, 0), (9, 4242)):
            want = o.enumerate_sublist_swap(lo, hi, step_index, seed, order)
            got = selectors.sublist_swap_rows(offs, lo, hi, MoveStreamContext(step_index, seed, order))
            assert np.array_equal(got, want), f"order={order} step={step_index} sizes={lo}..{hi}"
    # reference KAT (selector/tests/sublist_neighborhood.rs:315-364)
    kat = selectors.sublist_swap_rows(np.array([0, 4, 7]), 2, 2)
    assert kat.tolist() == [[0, 0, 2, 0, 2, 4], [0, 0, 2, 1, 0, 2], [0, 0, 2, 1, 1, 3], [0, 1, 3, 1, 0, 2],
                            [0, 1, 3, 1, 1, 3], [0, 2, 4, 1, 0, 2], [0, 2, 4, 1, 1, 3]]


@pytest.mark.parametrize("order", [selectors.ORIGINAL, selectors.RANDOM, selectors.SHUFFLED])
def test_scalar_swap_order_matches_reference(order):
    """SwapMoveSelector pull order (move_selector/swap.rs:64-100,196-233)."""
    g = instances.graph_coloring(60, 150, 4, seed_edges=2, seed_colors=3)
    o = Oracle.graph_coloring(g)
    for step_index, seed in ((0, 0), (9, 4242), (77, 0xFEEDFACE12345)):
        want = o.enumerate_swap(step_index, seed, order)
        got = selectors.swap_move_rows(g.n, MoveStreamContext(step_index, seed, order))
        assert np.array_equal(got, want), f"order={order} step={step_index}"
    assert selectors.swap_move_rows(4).tolist() == [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]
    assert len(selectors.swap_move_rows(g.n)) == g.n * (g.n - 1) // 2


def test_sublist_selector_golden_vectors():
    """tests/golden/reference_kats.json: the reference's canonical sublist change / swap orders."""
    import json
    import os
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
    for v in golden["sublist_selectors"]:
        fn = selectors.sublist_change_rows if v["kind"] == "change" else selectors.sublist_swap_rows
        assert fn(np.array(v["offsets"]), v["min"], v["max"]).tolist() == v["rows"], v["cite"]


def test_segment_row_packing_round_trip():
    """The 4-word device rows of the sublist moves ({entity, start | size << 24, ...}, SFGPU_SEG in sfgpu.h)."""
    from solverforge_b200 import _lib as L
    from solverforge_b200.api import GpuScoreDirector
    rows = selectors.sublist_change_rows(np.array([0, 5, 9, 9, 12]), 1, 3)
    packed = GpuScoreDirector.pack_sublist_change(rows)
    assert packed.shape == (len(rows), 4)
    assert np.array_equal(packed[:, 1] & 0xFFFFFF, rows[:, 1]) and np.array_equal(packed[:, 1] >> 24, rows[:, 2] - rows[:, 1])
    assert np.array_equal(packed[:, [0, 2, 3]], rows[:, [0, 3, 4]].astype(np.int64))
    sw = selectors.sublist_swap_rows(np.array([0, 5, 9, 9, 12]), 1, 3)
    ps = GpuScoreDirector.pack_sublist_swap(sw)
    assert np.array_equal(ps[:, 1] >> 24, sw[:, 2] - sw[:, 1]) and np.array_equal(ps[:, 3] >> 24, sw[:, 5] - sw[:, 4])
    assert np.array_equal(ps[:, 3] & 0xFFFFFF, sw[:, 4])
    with pytest.raises(L.SfgpuError):
        GpuScoreDirector.pack_sublist_change(np.array([[0, 0, 300, 1, 0]]))      # size > 255
    with pytest.raises(L.SfgpuError):
        GpuScoreDirector.pack_sublist_change(np.array([[0, 5, 3, 1, 0]]))        # end < start
    with pytest.raises(L.SfgpuError):
        GpuScoreDirector.pack_sublist_swap(np.array([[0, 1 << 24, (1 << 24) + 1, 1, 0, 1]]))  # start beyond the packing


@pytest.mark.parametrize("order", [selectors.ORIGINAL, selectors.RANDOM, selectors.SHUFFLED])
def test_k_opt_order_matches_reference(order):
    """KOptMoveSelector pull order (list_kernel/k_opt/full.rs:34-98, cut unranking k_opt/iterators.rs:110-160); the
    oracle scores the rows (CPU groundwork — no device path yet). Reference KAT: 245 moves for one 8-city tour."""
    assert selectors.k_opt_pattern_count(2) == 1 and selectors.k_opt_pattern_count(3) == 7
    assert selectors.k_opt_pattern_count(4) == 47 and selectors.k_opt_pattern_count(5) == 383
    assert len(selectors.k_opt_rows(np.array([0, 8]), 3)) == 245
    assert selectors.k_opt_rows(np.array([0, 8]), 3)[0].tolist() == [0, 1, 2, 3, 0]
    c = instances.cvrp(22, 3, seed=6)
    offs, el = instances.perturb_routes(c, 5, 12)
    o = Oracle.cvrp(c, offs, el)
    for k, min_seg in ((2, 1), (3, 1), (3, 2)):
        for step_index, seed in ((0, 0), (4, 99)):
            want = o.enumerate_k_opt(k, min_seg, step_index, seed, order)
            got = selectors.k_opt_rows(offs, k, min_seg, MoveStreamContext(step_index, seed, order))
            assert np.array_equal(got, want), f"k={k} min_seg={min_seg} order={order}"
    rows = o.enumerate_k_opt(2, 1)
    s2, d2 = o.score_k_opt(rows, 2)
    assert d2.all()
    # 2-opt == ListReverseMove on the same window [cut_0, cut_1)
    rev = np.stack([rows[:, 0], rows[:, 1], rows[:, 2], np.zeros(len(rows), dtype=np.uint32)], axis=1)
    keep = rows[:, 2] > rows[:, 1] + 1          # a one-element window is not a doable reversal
    sr, dr = o.score_list_reverse(rev[keep])
    assert dr.all() and np.array_equal(sr, s2[keep])
