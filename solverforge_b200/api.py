"""Host-side mirror of the reference's scoring interface above the C ABI.

The reference plugs its scorer in as ``ScoreDirector::with_descriptor(solution, constraints, ..)``
(solverforge-solver/src/run.rs:552-557) and authors constraints with the fluent
``ConstraintFactory::new().for_each(..)...penalize(..).named(..)`` builder
(solverforge-scoring/src/stream/factory.rs:43-73). This module keeps those names and shapes;
closures become column expressions that lower to ``sfgpu_constraint_desc`` rows.

All compute happens in libsfgpu (CUDA); nothing here scores anything on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib as L


@dataclass(frozen=True, order=True)
class HardSoftScore:
    """solverforge-core/src/score/hard_soft.rs:35-153 — ordering is hard, then soft."""
    hard: int = 0
    soft: int = 0

    @staticmethod
    def of(hard: int, soft: int) -> "HardSoftScore":
        return HardSoftScore(hard, soft)

    @staticmethod
    def of_hard(h: int) -> "HardSoftScore":
        return HardSoftScore(h, 0)

    @staticmethod
    def of_soft(s: int) -> "HardSoftScore":
        return HardSoftScore(0, s)

    def is_feasible(self) -> bool:
        return self.hard >= 0

    def __add__(self, o):
        return HardSoftScore(self.hard + o.hard, self.soft + o.soft)

    def __sub__(self, o):
        return HardSoftScore(self.hard - o.hard, self.soft - o.soft)

    def __neg__(self):
        return HardSoftScore(-self.hard, -self.soft)

    def __str__(self):
        return f"{self.hard}hard/{self.soft}soft"


HardSoftScore.ZERO = HardSoftScore(0, 0)
HardSoftScore.ONE_HARD = HardSoftScore(1, 0)
HardSoftScore.ONE_SOFT = HardSoftScore(0, 1)


class HardSoftDecimalScore(HardSoftScore):
    """Same layout, levels pre-scaled by 100000 (hard_soft_decimal.rs:14,45-48)."""
    SCALE = 100000

    @staticmethod
    def of(hard: int, soft: int) -> "HardSoftDecimalScore":
        return HardSoftDecimalScore(hard * 100000, soft * 100000)

    @staticmethod
    def of_scaled(hard: int, soft: int) -> "HardSoftDecimalScore":
        return HardSoftDecimalScore(hard, soft)


@dataclass(frozen=True)
class WeightFn:
    """weight(x) on one score level: the device form of a ``penalize(|..| Score)`` closure."""
    fn: int
    level: int
    a: int
    b: int = 0


def _const_weight(score: HardSoftScore) -> WeightFn:
    if score.hard != 0 and score.soft != 0:
        raise L.SfgpuError(L.E_UNSUPPORTED, "a constant weight must sit on one level")
    return WeightFn(L.W_CONST, 0 if score.hard != 0 else 1, score.hard if score.hard != 0 else score.soft)


def hard(fn: int = L.W_LINEAR, a: int = 1, b: int = 0) -> WeightFn:
    return WeightFn(fn, 0, a, b)


def soft(fn: int = L.W_LINEAR, a: int = 1, b: int = 0) -> WeightFn:
    return WeightFn(fn, 1, a, b)


def _i64(v: int) -> int:
    """two's-complement view of an unsigned 64-bit value (ctypes int64 fields)"""
    return v - (1 << 64) if v >= (1 << 63) else v


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class GpuScoreDirector:
    """Owns one ``sfgpu_ctx``: R replicas of the planning state + the constraint program.

    Reference counterpart: ``ScoreDirector<S, C>`` (director/score_director/incremental.rs:64-224)
    and, for the batch seam the reference lacks, ``evaluate_candidate`` (evaluation.rs:20-115)
    applied to a whole neighbourhood at once.
    """

    def __init__(self, n_replicas: int = 1, device: int = 0, stream: Optional[int] = None, flags: int = 0):
        self.lib = L.load()
        self.R = int(n_replicas)
        h = C.c_void_p()
        rc = self.lib.sfgpu_ctx_create(device, flags, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != L.OK:
            raise L.SfgpuError(rc, self.lib.sfgpu_last_error(None).decode())
        self.h = h
        self._check(self.lib.sfgpu_model_begin(self.h, self.R))
        self.coll_rows: dict[int, int] = {}
        self.scalar_var: Optional[int] = None
        self.list_var: Optional[int] = None
        self.n_owners = 0
        self.n_entities = 0
        self._keep = []

    # ---- plumbing -------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != L.OK:
            raise L.SfgpuError(rc, self.lib.sfgpu_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.sfgpu_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- model ----------------------------------------------------------------------
    def add_collection(self, name: str, n_rows: int, descriptor_index: int = -1) -> int:
        out = C.c_uint32()
        self._check(self.lib.sfgpu_add_collection(self.h, name.encode(), n_rows, descriptor_index, C.byref(out)))
        self.coll_rows[out.value] = n_rows
        return out.value

    def add_column(self, collection: int, name: str, values) -> int:
        v = np.ascontiguousarray(values, dtype=np.int64)
        if v.shape != (self.coll_rows[collection],):
            raise L.SfgpuError(L.E_INVALID, "column length != collection rows")
        out = C.c_uint32()
        self._check(self.lib.sfgpu_add_column_i64(self.h, collection, name.encode(), _ptr(v), C.byref(out)))
        self._col_host = getattr(self, "_col_host", {})
        self._col_host[out.value] = v.copy()
        return out.value

    def add_scalar_variable(self, collection: int, name: str, n_values: int, allows_unassigned: bool = True) -> int:
        out = C.c_uint32()
        self._check(self.lib.sfgpu_add_scalar_variable(self.h, collection, name.encode(), n_values,
                                                       1 if allows_unassigned else 0, C.byref(out)))
        self.scalar_var = out.value
        self.n_entities = self.coll_rows[collection]
        return out.value

    def add_list_variable(self, owner_collection: int, element_collection: int, name: str) -> int:
        out = C.c_uint32()
        self._check(self.lib.sfgpu_add_list_variable(self.h, owner_collection, element_collection, name.encode(),
                                                     C.byref(out)))
        self.list_var = out.value
        self.n_owners = self.coll_rows[owner_collection]
        return out.value

    def add_csr(self, name: str, row_ptr, col_idx) -> int:
        rp = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        ci = np.ascontiguousarray(col_idx, dtype=np.uint32)
        out = C.c_uint32()
        self._check(self.lib.sfgpu_add_csr(self.h, name.encode(), len(rp) - 1, _ptr(rp), _ptr(ci), C.byref(out)))
        self._csr_host = getattr(self, "_csr_host", {})
        self._csr_host[out.value] = (rp.copy(), ci.copy())
        return out.value

    def add_matrix(self, name: str, values, cost_semantics: bool = False) -> int:
        m = np.ascontiguousarray(values, dtype=np.int64)
        out = C.c_uint32()
        self._check(self.lib.sfgpu_add_matrix_i64(self.h, name.encode(), m.shape[0], m.shape[1], _ptr(m),
                                                  1 if cost_semantics else 0, C.byref(out)))
        return out.value

    def add_expr(self, expr: "Expr") -> int:
        """Registers a column expression (sfgpu_add_expr) and returns its id."""
        ops = (L.ExprOp * len(expr.ops))(*[L.ExprOp(o, a, i) for o, a, i in expr.ops])
        out = C.c_uint32()
        self._check(self.lib.sfgpu_add_expr(self.h, ops, len(expr.ops), C.byref(out)))
        return out.value

    def add_constraint(self, kind: int, impact: int, weight: WeightFn, collection: int = 0, variable: int = 0,
                       aux0: int = L.NO_COLUMN, aux1: int = L.NO_COLUMN, p0: int = 0, p1: int = 0, name: str = "") -> int:
        d = L.ConstraintDesc(kind, impact, L.Weight(weight.fn, weight.level, weight.a, weight.b), collection,
                             variable, aux0, aux1, p0, p1, name.encode())
        out = C.c_uint32()
        self._check(self.lib.sfgpu_add_constraint(self.h, C.byref(d), C.byref(out)))
        return out.value

    def set_scalar_state(self, values):
        v = np.ascontiguousarray(values, dtype=np.int32)
        per = 1 if v.ndim == 2 else 0
        self._check(self.lib.sfgpu_set_scalar_state(self.h, self.scalar_var, _ptr(v), per))

    def set_list_state(self, offsets, elems):
        """offsets: [n_owners+1] (broadcast) or [R][n_owners+1]; elems: concatenated copies."""
        o = np.ascontiguousarray(offsets, dtype=np.uint32)
        e = np.ascontiguousarray(elems, dtype=np.uint32)
        per = 1 if o.ndim == 2 else 0
        self._check(self.lib.sfgpu_set_list_state(self.h, self.list_var, _ptr(o), _ptr(e), per))

    def commit(self) -> np.ndarray:
        """Freezes the model and runs initialize_all; returns committed scores [R, 2]."""
        out = np.zeros((self.R, 2), dtype=np.int64)
        self._check(self.lib.sfgpu_model_commit(self.h, _ptr(out)))
        return out

    # ---- Director surface --------------------------------------------------------------
    def calculate_score(self) -> np.ndarray:
        """Committed (cached) score of every replica — Director::calculate_score."""
        out = np.zeros((self.R, 2), dtype=np.int64)
        self._check(self.lib.sfgpu_committed_scores(self.h, _ptr(out)))
        return out

    def fresh_score(self) -> np.ndarray:
        """Stateless full recompute — Director::fresh_score / ConstraintSet::evaluate_all."""
        out = np.zeros((self.R, 2), dtype=np.int64)
        self._check(self.lib.sfgpu_evaluate_all(self.h, _ptr(out)))
        return out

    # ---- batched candidate scoring (host buffers) ---------------------------------------
    def _offsets(self, cand_offsets, n: int) -> np.ndarray:
        if cand_offsets is None:
            if self.R != 1:
                raise L.SfgpuError(L.E_INVALID, "cand_offsets is required when R > 1")
            return np.array([0, n], dtype=np.uint64)
        o = np.ascontiguousarray(cand_offsets, dtype=np.uint64)
        if o.shape != (self.R + 1,):
            raise L.SfgpuError(L.E_INVALID, "cand_offsets must have R+1 entries")
        return o

    def _score(self, fn, rows: np.ndarray, words: int, cand_offsets):
        rows = np.ascontiguousarray(rows, dtype=np.uint32).reshape(-1, words)
        n = rows.shape[0]
        offs = self._offsets(cand_offsets, n)
        scores = np.zeros((n, 2), dtype=np.int64)
        doable = np.zeros(n, dtype=np.uint8)
        self._check(fn(self.h, 0, n, _ptr(offs), _ptr(rows), _ptr(scores), _ptr(doable)))
        return scores, doable

    def score_change(self, rows, cand_offsets=None):
        """rows[n][2] = (entity, to_value) with -1 (0xFFFFFFFF) for None."""
        return self._score(self.lib.sfgpu_score_change, np.asarray(rows).astype(np.int64).astype(np.uint32), 2,
                           cand_offsets)

    def score_swap(self, rows, cand_offsets=None):
        return self._score(self.lib.sfgpu_score_swap, rows, 2, cand_offsets)

    def score_list_change(self, rows, cand_offsets=None):
        return self._score(self.lib.sfgpu_score_list_change, rows, 4, cand_offsets)

    def score_list_swap(self, rows, cand_offsets=None):
        return self._score(self.lib.sfgpu_score_list_swap, rows, 4, cand_offsets)

    def score_list_reverse(self, rows, cand_offsets=None):
        """rows[n][3 or 4] = (entity, start, end[, 0]): ListReverseMove (2-opt segment reversal)."""
        rows = np.asarray(rows).astype(np.int64)
        if rows.ndim == 2 and rows.shape[1] == 3:
            rows = np.concatenate([rows, np.zeros((len(rows), 1), dtype=np.int64)], axis=1)
        return self._score(self.lib.sfgpu_score_list_reverse, rows, 4, cand_offsets)

    @staticmethod
    def pack_sublist_change(rows) -> np.ndarray:
        """(src_entity, start, end, dst_entity, dst_position) -> the 4-word device row
        {src_entity, start | size << 24, dst_entity, dst_position} (SFGPU_SEG in sfgpu.h)."""
        rows = np.asarray(rows).astype(np.int64).reshape(-1, 5)
        size = rows[:, 2] - rows[:, 1]
        if len(rows) and (size.min() < 0 or size.max() > 255 or rows[:, 1].max() >= 1 << 24):
            raise L.SfgpuError(L.E_INVALID, "sublist segment outside the packed range (size 0..255, start < 2^24)")
        return np.stack([rows[:, 0], rows[:, 1] | (size << 24), rows[:, 3], rows[:, 4]], axis=1)

    def score_sublist_change(self, rows, cand_offsets=None):
        """rows[n][5] = (src_entity, start, end, dst_entity, dst_position): SublistChangeMove."""
        return self._score(self.lib.sfgpu_score_sublist_change, self.pack_sublist_change(rows), 4, cand_offsets)

    @staticmethod
    def pack_sublist_swap(rows) -> np.ndarray:
        """(first_entity, start1, end1, second_entity, start2, end2) -> {e1, start1 | size1 << 24, e2, start2 | size2 << 24}."""
        rows = np.asarray(rows).astype(np.int64).reshape(-1, 6)
        n1, n2 = rows[:, 2] - rows[:, 1], rows[:, 5] - rows[:, 4]
        if len(rows) and (min(n1.min(), n2.min()) < 0 or max(n1.max(), n2.max()) > 255 or
                          max(rows[:, 1].max(), rows[:, 4].max()) >= 1 << 24):
            raise L.SfgpuError(L.E_INVALID, "sublist segment outside the packed range (size 0..255, start < 2^24)")
        return np.stack([rows[:, 0], rows[:, 1] | (n1 << 24), rows[:, 3], rows[:, 4] | (n2 << 24)], axis=1)

    def score_sublist_swap(self, rows, cand_offsets=None):
        """rows[n][6] = (first_entity, start1, end1, second_entity, start2, end2): SublistSwapMove."""
        return self._score(self.lib.sfgpu_score_sublist_swap, self.pack_sublist_swap(rows), 4, cand_offsets)

    @staticmethod
    def pack_k_opt(rows, k: int) -> np.ndarray:
        """(entity, cut_0 .. cut_{k-1}, pattern) rows -> the packed device rows of sfgpu_score_k_opt."""
        r = np.asarray(rows).astype(np.int64).reshape(-1, k + 2)
        c = np.zeros((len(r), 5), dtype=np.int64)
        c[:, :k] = r[:, 1:1 + k]
        out = np.stack([r[:, 0] | (k << 28), c[:, 0] | (c[:, 1] << 16), c[:, 2] | (c[:, 3] << 16), c[:, 4] | (r[:, k + 1] << 16)],
                       axis=1)
        return np.ascontiguousarray(out.astype(np.uint32))

    def score_k_opt(self, rows, k: int = 3, cand_offsets=None):
        """rows[n][k + 2] = (entity, cuts.., pattern index) — KOptMove on one list (sfgpu_score_k_opt)."""
        return self._score(self.lib.sfgpu_score_k_opt, self.pack_k_opt(rows, k), 4, cand_offsets)

    def apply_k_opt(self, rows, k: int = 3, mask=None):
        return self._apply(self.lib.sfgpu_apply_k_opt, self.pack_k_opt(rows, k), 4, mask)

    def score_compound(self, edit_offsets, edit_rows, cand_offsets=None):
        eo = np.ascontiguousarray(edit_offsets, dtype=np.uint64)
        rows = np.ascontiguousarray(np.asarray(edit_rows).astype(np.int64).astype(np.uint32)).reshape(-1, 2)
        n = len(eo) - 1
        offs = self._offsets(cand_offsets, n)
        scores = np.zeros((n, 2), dtype=np.int64)
        doable = np.zeros(n, dtype=np.uint8)
        self._check(self.lib.sfgpu_score_compound(self.h, 0, n, _ptr(offs), _ptr(eo), _ptr(rows), _ptr(scores),
                                                  _ptr(doable)))
        return scores, doable

    # ---- device-pointer variants (torch tensors / raw pointers, asynchronous) ------------
    def score_device(self, kind: str, n: int, offsets_ptr: int, rows_ptr: int, scores_ptr: int, doable_ptr: int):
        fn = getattr(self.lib, f"sfgpu_score_{kind}")
        self._check(fn(self.h, L.DEVICE_IO, n, C.c_void_p(offsets_ptr), C.c_void_p(rows_ptr), C.c_void_p(scores_ptr),
                       C.c_void_p(doable_ptr)))

    def argbest_device(self, params: "ForageParams", offsets_ptr, scores_ptr, doable_ptr, seeds_ptr, ref_ptr,
                       index_ptr, best_ptr, evaluated_ptr):
        fp = L.ForageParams(params.acceptor, params.tie_mode, params.accepted_limit, 0)
        v = lambda p: C.c_void_p(p) if p else None
        self._check(self.lib.sfgpu_argbest(self.h, L.DEVICE_IO, C.byref(fp), v(offsets_ptr), v(scores_ptr),
                                           v(doable_ptr), v(seeds_ptr), v(ref_ptr), v(index_ptr), v(best_ptr),
                                           v(evaluated_ptr)))

    def step_list_change_device(self, n: int, offsets_ptr: int, rows_ptr: int, params: "ForageParams", seeds_ptr: int,
                                ref_ptr: int, scores_ptr: int, doable_ptr: int, index_ptr: int, best_ptr: int,
                                evaluated_ptr: int):
        """Fused score + acceptor/forager replay (sfgpu_step_list_change); scores_ptr/doable_ptr may be 0."""
        fp = L.ForageParams(params.acceptor, params.tie_mode, params.accepted_limit, 0)
        v = lambda p: C.c_void_p(p) if p else None
        self._check(self.lib.sfgpu_step_list_change(self.h, n, v(offsets_ptr), v(rows_ptr), C.byref(fp), v(seeds_ptr),
                                                    v(ref_ptr), v(scores_ptr), v(doable_ptr), v(index_ptr),
                                                    v(best_ptr), v(evaluated_ptr)))

    def step_change_rows_device(self, n: int, offsets_ptr: int, rows_ptr: int, params: "ForageParams", seeds_ptr: int,
                                ref_ptr: int, scores_ptr: int, doable_ptr: int, index_ptr: int, best_ptr: int,
                                evaluated_ptr: int):
        """Fused score + acceptor/forager replay over resident ChangeMove rows (sfgpu_step_change_rows);
        scores_ptr/doable_ptr may be 0."""
        fp = L.ForageParams(params.acceptor, params.tie_mode, params.accepted_limit, 0)
        v = lambda p: C.c_void_p(p) if p else None
        self._check(self.lib.sfgpu_step_change_rows(self.h, n, v(offsets_ptr), v(rows_ptr), C.byref(fp), v(seeds_ptr),
                                                    v(ref_ptr), v(scores_ptr), v(doable_ptr), v(index_ptr),
                                                    v(best_ptr), v(evaluated_ptr)))

    def step_nearby_list_swap(self, max_nearby: int = 20, params: "ForageParams" = None, step_seeds=None,
                              ref_scores=None, apply: bool = False, out_rows_ptr: int = 0, out_scores_ptr: int = 0,
                              out_doable_ptr: int = 0, out_offsets_ptr: int = 0):
        """One whole step over the nearby list-swap neighbourhood (sfgpu_step_nearby_list_swap); same
        conventions and return values as step_nearby_list_change."""
        return self.step_nearby_list_change(max_nearby, params, step_seeds, ref_scores, apply, out_rows_ptr,
                                            out_scores_ptr, out_doable_ptr, out_offsets_ptr,
                                            _fn=self.lib.sfgpu_step_nearby_list_swap)

    def step_nearby_list_change(self, max_nearby: int = 20, params: "ForageParams" = None, step_seeds=None,
                                ref_scores=None, apply: bool = False, out_rows_ptr: int = 0, out_scores_ptr: int = 0,
                                out_doable_ptr: int = 0, out_offsets_ptr: int = 0, _fn=None):
        """One whole local-search step on device (sfgpu_step_nearby_list_change) with HOST per-replica
        arrays: returns (index[R], best[R,2], moves_evaluated[R], winner_rows[R,4])."""
        params = params or ForageParams()
        fp = L.ForageParams(params.acceptor, params.tie_mode, params.accepted_limit, 0)
        seeds = None if step_seeds is None else np.ascontiguousarray(step_seeds, dtype=np.uint64)
        ref = None if ref_scores is None else np.ascontiguousarray(ref_scores, dtype=np.int64).reshape(self.R, 4)
        idx = np.zeros(self.R, dtype=np.uint32)
        best = np.zeros((self.R, 2), dtype=np.int64)
        ev = np.zeros(self.R, dtype=np.uint32)
        win = np.zeros((self.R, 4), dtype=np.uint32)
        v = lambda p: C.c_void_p(p) if p else None
        self._check((_fn or self.lib.sfgpu_step_nearby_list_change)(
            self.h, 0, max_nearby, C.byref(fp), _ptr(seeds), _ptr(ref), v(out_offsets_ptr), v(out_rows_ptr),
            v(out_scores_ptr), v(out_doable_ptr), _ptr(idx), _ptr(best), _ptr(ev), _ptr(win), 1 if apply else 0))
        return idx, best, ev, win

    def step_nearby_list_change_device(self, max_nearby: int, params: "ForageParams", seeds_ptr: int, ref_ptr: int,
                                       index_ptr: int, best_ptr: int, evaluated_ptr: int, winner_rows_ptr: int,
                                       apply: bool = False, out_offsets_ptr: int = 0, out_rows_ptr: int = 0,
                                       out_scores_ptr: int = 0, out_doable_ptr: int = 0):
        fp = L.ForageParams(params.acceptor, params.tie_mode, params.accepted_limit, 0)
        v = lambda p: C.c_void_p(p) if p else None
        self._check(self.lib.sfgpu_step_nearby_list_change(
            self.h, L.DEVICE_IO, max_nearby, C.byref(fp), v(seeds_ptr), v(ref_ptr), v(out_offsets_ptr),
            v(out_rows_ptr), v(out_scores_ptr), v(out_doable_ptr), v(index_ptr), v(best_ptr), v(evaluated_ptr),
            v(winner_rows_ptr), 1 if apply else 0))

    def step_change(self, params: "ForageParams" = None, step_seeds=None, ref_scores=None, apply: bool = False,
                    out_rows_ptr: int = 0, out_scores_ptr: int = 0, out_doable_ptr: int = 0, out_offsets_ptr: int = 0):
        """One whole step of a scalar model on device (sfgpu_step_change): full ChangeMove neighbourhood +
        scoring + forager. Returns (index[R], best[R,2], moves_evaluated[R], winner_rows[R,2])."""
        params = params or ForageParams()
        fp = L.ForageParams(params.acceptor, params.tie_mode, params.accepted_limit, 0)
        seeds = None if step_seeds is None else np.ascontiguousarray(step_seeds, dtype=np.uint64)
        ref = None if ref_scores is None else np.ascontiguousarray(ref_scores, dtype=np.int64).reshape(self.R, 4)
        idx = np.zeros(self.R, dtype=np.uint32)
        best = np.zeros((self.R, 2), dtype=np.int64)
        ev = np.zeros(self.R, dtype=np.uint32)
        win = np.zeros((self.R, 2), dtype=np.uint32)
        v = lambda p: C.c_void_p(p) if p else None
        self._check(self.lib.sfgpu_step_change(self.h, 0, C.byref(fp), _ptr(seeds), _ptr(ref), v(out_offsets_ptr),
                                               v(out_rows_ptr), v(out_scores_ptr), v(out_doable_ptr), _ptr(idx),
                                               _ptr(best), _ptr(ev), _ptr(win), 1 if apply else 0))
        return idx, best, ev, win.view(np.int32)

    def step_sublist_change(self, min_size: int = 1, max_size: int = 3, params: "ForageParams" = None, step_seeds=None,
                            ref_scores=None, apply: bool = False):
        """One whole step over the SublistChange neighbourhood, enumerated on device by pull index
        (sfgpu_step_sublist_change). Returns (index[R], best[R,2], moves_evaluated[R], winner_rows[R,5]) with
        winner rows unpacked to (src_entity, start, end, dst_entity, dst_position); -1 when there is no winner."""
        params = params or ForageParams()
        fp = L.ForageParams(params.acceptor, params.tie_mode, params.accepted_limit, 0)
        seeds = None if step_seeds is None else np.ascontiguousarray(step_seeds, dtype=np.uint64)
        ref = None if ref_scores is None else np.ascontiguousarray(ref_scores, dtype=np.int64).reshape(self.R, 4)
        idx = np.zeros(self.R, dtype=np.uint32)
        best = np.zeros((self.R, 2), dtype=np.int64)
        ev = np.zeros(self.R, dtype=np.uint32)
        win = np.zeros((self.R, 4), dtype=np.uint32)
        self._check(self.lib.sfgpu_step_sublist_change(self.h, 0, min_size, max_size, C.byref(fp), _ptr(seeds),
                                                       _ptr(ref), _ptr(idx), _ptr(best), _ptr(ev), _ptr(win),
                                                       1 if apply else 0))
        w = win.astype(np.int64)
        start, size = w[:, 1] & 0xFFFFFF, w[:, 1] >> 24
        rows = np.stack([w[:, 0], start, start + size, w[:, 2], w[:, 3]], axis=1)
        rows[idx == 0xFFFFFFFF] = -1
        return idx, best, ev, rows

    def step_list_reverse(self, params: "ForageParams" = None, step_seeds=None, ref_scores=None, apply: bool = False):
        """One whole step over the ListReverse (2-opt) neighbourhood, enumerated on device (sfgpu_step_list_reverse).
        Winner rows come back as (entity, start, end); -1 when there is no winner."""
        params = params or ForageParams()
        fp = L.ForageParams(params.acceptor, params.tie_mode, params.accepted_limit, 0)
        seeds = None if step_seeds is None else np.ascontiguousarray(step_seeds, dtype=np.uint64)
        ref = None if ref_scores is None else np.ascontiguousarray(ref_scores, dtype=np.int64).reshape(self.R, 4)
        idx = np.zeros(self.R, dtype=np.uint32)
        best = np.zeros((self.R, 2), dtype=np.int64)
        ev = np.zeros(self.R, dtype=np.uint32)
        win = np.zeros((self.R, 4), dtype=np.uint32)
        self._check(self.lib.sfgpu_step_list_reverse(self.h, 0, C.byref(fp), _ptr(seeds), _ptr(ref), _ptr(idx), _ptr(best),
                                                     _ptr(ev), _ptr(win), 1 if apply else 0))
        rows = win.astype(np.int64)[:, :3]
        rows[idx == 0xFFFFFFFF] = -1
        return idx, best, ev, rows

    def step_sublist_swap(self, min_size: int = 1, max_size: int = 3, params: "ForageParams" = None, step_seeds=None,
                          ref_scores=None, apply: bool = False):
        """One whole step over the SublistSwap neighbourhood, enumerated on device (sfgpu_step_sublist_swap). Winner
        rows come back as (first_entity, start1, end1, second_entity, start2, end2); -1 when there is no winner."""
        params = params or ForageParams()
        fp = L.ForageParams(params.acceptor, params.tie_mode, params.accepted_limit, 0)
        seeds = None if step_seeds is None else np.ascontiguousarray(step_seeds, dtype=np.uint64)
        ref = None if ref_scores is None else np.ascontiguousarray(ref_scores, dtype=np.int64).reshape(self.R, 4)
        idx = np.zeros(self.R, dtype=np.uint32)
        best = np.zeros((self.R, 2), dtype=np.int64)
        ev = np.zeros(self.R, dtype=np.uint32)
        win = np.zeros((self.R, 4), dtype=np.uint32)
        self._check(self.lib.sfgpu_step_sublist_swap(self.h, 0, min_size, max_size, C.byref(fp), _ptr(seeds), _ptr(ref),
                                                     _ptr(idx), _ptr(best), _ptr(ev), _ptr(win), 1 if apply else 0))
        w = win.astype(np.int64)
        s1, n1, s2, n2 = w[:, 1] & 0xFFFFFF, w[:, 1] >> 24, w[:, 3] & 0xFFFFFF, w[:, 3] >> 24
        rows = np.stack([w[:, 0], s1, s1 + n1, w[:, 2], s2, s2 + n2], axis=1)
        rows[idx == 0xFFFFFFFF] = -1
        return idx, best, ev, rows

    @staticmethod
    def union_desc(children, union_order: int = L.UNION_STRATIFIED_RANDOM, selection_order: int = L.ORDER_RANDOM,
                   window: int = 0, max_window: int = 0) -> "L.UnionDesc":
        """children: [(family, p0, p1[, weight])] — see sfgpu_union_desc. The reference's default list policy is
        `default_list_union()`."""
        d = L.UnionDesc()
        d.n_children = len(children)
        d.union_order = union_order
        d.selection_order = selection_order
        d.window = window
        d.max_window = max_window
        for i, ch in enumerate(children):
            d.children[i].family = ch[0]
            d.children[i].p0 = ch[1] if len(ch) > 1 else 0
            d.children[i].p1 = ch[2] if len(ch) > 2 else 0
            d.children[i].weight = ch[3] if len(ch) > 3 else 1
        return d

    @staticmethod
    def default_list_union(max_nearby: int = 20, window: int = 0, max_window: int = 0) -> "L.UnionDesc":
        """The device-enumerable rows of the reference's default list policy table
        (runtime/compiler/default_local_search/policy/list.rs:24-33): NearbyChange, NearbySwap, SublistChange,
        SublistSwap, Reverse; seeded Random leaves interleaved by StratifiedRandom."""
        return GpuScoreDirector.union_desc(
            [(L.FAM_NEARBY_LIST_CHANGE, max_nearby), (L.FAM_NEARBY_LIST_SWAP, max_nearby), (L.FAM_SUBLIST_CHANGE, 1, 3),
             (L.FAM_SUBLIST_SWAP, 1, 3), (L.FAM_LIST_REVERSE,)], L.UNION_STRATIFIED_RANDOM, L.ORDER_RANDOM, window,
            max_window)

    @staticmethod
    def default_scalar_union(window: int = 0, max_window: int = 0) -> "L.UnionDesc":
        """The reference's default selectors of a plain scalar model (policy/scalar.rs:64-108): ChangeMoveSelector +
        SwapMoveSelector, seeded Random leaves, StratifiedRandom union."""
        return GpuScoreDirector.union_desc([(L.FAM_CHANGE,), (L.FAM_SWAP,)], L.UNION_STRATIFIED_RANDOM, L.ORDER_RANDOM, window,
                                           max_window)

    def step_union(self, desc: "L.UnionDesc", params: "ForageParams" = None, step_seeds=None, step_indices=None,
                   ref_scores=None, apply: bool = False):
        """One step over a union of list neighbourhoods in the reference's seeded pull order (sfgpu_step_union).
        Returns (index[R], best[R,2], moves_evaluated[R], winner_rows[R,8], flags[R]); winner row =
        {family, child, packed row[4], child-local pull index, 0}."""
        params = params or ForageParams()
        fp = L.ForageParams(params.acceptor, params.tie_mode, params.accepted_limit, 0)
        seeds = None if step_seeds is None else np.ascontiguousarray(step_seeds, dtype=np.uint64)
        steps = None if step_indices is None else np.ascontiguousarray(step_indices, dtype=np.uint64)
        ref = None if ref_scores is None else np.ascontiguousarray(ref_scores, dtype=np.int64).reshape(self.R, 4)
        idx = np.zeros(self.R, dtype=np.uint32)
        best = np.zeros((self.R, 2), dtype=np.int64)
        ev = np.zeros(self.R, dtype=np.uint32)
        win = np.zeros((self.R, 8), dtype=np.uint32)
        flags = np.zeros(self.R, dtype=np.uint32)
        self._check(self.lib.sfgpu_step_union(self.h, 0, C.byref(desc), C.byref(fp), _ptr(seeds), _ptr(steps), _ptr(ref),
                                              _ptr(idx), _ptr(best), _ptr(ev), _ptr(win), _ptr(flags), 1 if apply else 0))
        return idx, best, ev, win, flags

    def solve_union(self, desc: "L.UnionDesc", n_steps: int, acceptor: int = 2, late_size: int = 400, tie_mode: int = 1,
                    accepted_limit: int = 256, seed_base: int = 0, restore_best: bool = False,
                    acceptor_real: float = 0.0, step_count_limit: int = 0):
        """Device-resident loop over the union step (sfgpu_solve_union): returns (best_scores[R,2],
        moves_evaluated[R], committed_steps[R], window_overflows[R]); self.last_pulls_scored[R] = union pulls scored
        over all window passes (speculation included)."""
        p = L.SolveParams(0, n_steps, acceptor, late_size, tie_mode, accepted_limit, seed_base,
                          1 if restore_best else 0, 0, acceptor_real, step_count_limit)
        best = np.zeros((self.R, 2), dtype=np.int64)
        ev = np.zeros(self.R, dtype=np.uint64)
        acc = np.zeros(self.R, dtype=np.uint64)
        ovf = np.zeros(self.R, dtype=np.uint64)
        pulls = np.zeros(self.R, dtype=np.uint64)
        self._check(self.lib.sfgpu_solve_union(self.h, C.byref(desc), C.byref(p), _ptr(best), _ptr(ev), _ptr(acc), _ptr(ovf),
                                               _ptr(pulls)))
        self.last_pulls_scored = pulls
        return best, ev, acc, ovf

    def solve_nearby_list_change(self, n_steps: int, max_nearby: int = 20, acceptor: int = 2, late_size: int = 400,
                                 tie_mode: int = 1, accepted_limit: int = 0, seed_base: int = 0,
                                 restore_best: bool = False, acceptor_real: float = 0.0,
                                 step_count_limit: int = 0, windowed: bool = False):
        """Device-resident local-search loop (sfgpu_solve_nearby_list_change): returns
        (best_scores[R,2], moves_evaluated[R], committed_steps[R]). acceptor: 1 HillClimbing,
        2 LateAcceptance(late_size), 3 GreatDeluge(acceptor_real = rain_speed),
        4 StepCountingHillClimbing(step_count_limit), 5 DiversifiedLateAcceptance(late_size,
        acceptor_real = tolerance)."""
        p = L.SolveParams(max_nearby, n_steps, acceptor, late_size, tie_mode, accepted_limit, seed_base,
                          1 if restore_best else 0, 1 if windowed else 0, acceptor_real, step_count_limit)
        best = np.zeros((self.R, 2), dtype=np.int64)
        ev = np.zeros(self.R, dtype=np.uint64)
        acc = np.zeros(self.R, dtype=np.uint64)
        self._check(self.lib.sfgpu_solve_nearby_list_change(self.h, C.byref(p), _ptr(best), _ptr(ev), _ptr(acc)))
        return best, ev, acc

    def solve_change(self, n_steps: int, acceptor: int = 2, late_size: int = 400, tie_mode: int = 1,
                     accepted_limit: int = 0, seed_base: int = 0, restore_best: bool = False,
                     acceptor_real: float = 0.0, step_count_limit: int = 0):
        """Device-resident loop over the full ChangeMove neighbourhood (sfgpu_solve_change); acceptors as
        solve_nearby_list_change."""
        p = L.SolveParams(0, n_steps, acceptor, late_size, tie_mode, accepted_limit, seed_base,
                          1 if restore_best else 0, 0, acceptor_real, step_count_limit)
        best = np.zeros((self.R, 2), dtype=np.int64)
        ev = np.zeros(self.R, dtype=np.uint64)
        acc = np.zeros(self.R, dtype=np.uint64)
        self._check(self.lib.sfgpu_solve_change(self.h, C.byref(p), _ptr(best), _ptr(ev), _ptr(acc)))
        return best, ev, acc

    def apply_winners_device(self, move_kind: int, offsets_ptr: int, rows_ptr: int, index_ptr: int):
        self._check(self.lib.sfgpu_apply_winners(self.h, move_kind, C.c_void_p(offsets_ptr), C.c_void_p(rows_ptr),
                                                 C.c_void_p(index_ptr)))

    def pack_best_keys_device(self, keys_ptr: int):
        self._check(self.lib.sfgpu_pack_best_keys(self.h, C.c_void_p(keys_ptr)))

    def sync_best(self, comm=None, scores_ptr: int = 0):
        """Best (hard, soft) over every replica of every rank of `comm` (an ncclComm_t created with
        sfgpu_comm_init_rank; None = this context only), its owner rank and replica (sfgpu_sync_best).
        scores_ptr: device pointer to R (hard, soft) pairs, 0 = the committed scores."""
        best = np.zeros(2, dtype=np.int64)
        owner, rep = C.c_int32(), C.c_uint32()
        self._check(self.lib.sfgpu_sync_best(self.h, comm, L.DEVICE_IO if scores_ptr else 0,
                                             C.c_void_p(scores_ptr) if scores_ptr else None, _ptr(best),
                                             C.byref(owner), C.byref(rep)))
        return (int(best[0]), int(best[1])), owner.value, rep.value

    # ---- forager / acceptor replay on device ------------------------------------------
    def argbest(self, scores, doable, cand_offsets=None, params: "ForageParams" = None, step_seeds=None,
                ref_scores=None, gates=None):
        """Acceptor + forager replay over scored rows (sfgpu_argbest / sfgpu_argbest_gated). gates[i]: bit 0 =
        requires_hard_improvement, bit 1 = requires_score_improvement (evaluation.rs:76-111)."""
        params = params or ForageParams()
        scores = np.ascontiguousarray(scores, dtype=np.int64).reshape(-1, 2)
        doable = np.ascontiguousarray(doable, dtype=np.uint8)
        offs = self._offsets(cand_offsets, scores.shape[0])
        seeds = None if step_seeds is None else np.ascontiguousarray(step_seeds, dtype=np.uint64)
        ref = None if ref_scores is None else np.ascontiguousarray(ref_scores, dtype=np.int64).reshape(self.R, 4)
        idx = np.zeros(self.R, dtype=np.uint32)
        best = np.zeros((self.R, 2), dtype=np.int64)
        ev = np.zeros(self.R, dtype=np.uint32)
        fp = L.ForageParams(params.acceptor, params.tie_mode, params.accepted_limit, 0)
        if gates is not None:
            g = np.ascontiguousarray(gates, dtype=np.uint8)
            self._check(self.lib.sfgpu_argbest_gated(self.h, 0, C.byref(fp), _ptr(offs), _ptr(scores), _ptr(doable),
                                                     _ptr(g), _ptr(seeds), _ptr(ref), _ptr(idx), _ptr(best), _ptr(ev)))
            return idx, best, ev
        self._check(self.lib.sfgpu_argbest(self.h, 0, C.byref(fp), _ptr(offs), _ptr(scores), _ptr(doable),
                                           _ptr(seeds), _ptr(ref), _ptr(idx), _ptr(best), _ptr(ev)))
        return idx, best, ev

    # ---- committing moves ---------------------------------------------------------------
    def _apply(self, fn, rows, words, mask):
        rows = np.ascontiguousarray(np.asarray(rows).astype(np.int64).astype(np.uint32)).reshape(self.R, words)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._check(fn(self.h, 0, _ptr(rows), _ptr(m)))

    def apply_change(self, rows, mask=None):
        self._apply(self.lib.sfgpu_apply_change, rows, 2, mask)

    def apply_swap(self, rows, mask=None):
        self._apply(self.lib.sfgpu_apply_swap, rows, 2, mask)

    def apply_list_change(self, rows, mask=None):
        self._apply(self.lib.sfgpu_apply_list_change, rows, 4, mask)

    def apply_list_swap(self, rows, mask=None):
        self._apply(self.lib.sfgpu_apply_list_swap, rows, 4, mask)

    def apply_list_reverse(self, rows, mask=None):
        rows = np.asarray(rows).astype(np.int64).reshape(self.R, -1)
        if rows.shape[1] == 3:
            rows = np.concatenate([rows, np.zeros((self.R, 1), dtype=np.int64)], axis=1)
        self._apply(self.lib.sfgpu_apply_list_reverse, rows, 4, mask)

    def apply_sublist_change(self, rows, mask=None):
        """rows[R][5] = (src_entity, start, end, dst_entity, dst_position), one per replica."""
        self._apply(self.lib.sfgpu_apply_sublist_change, self.pack_sublist_change(rows), 4, mask)

    def apply_sublist_swap(self, rows, mask=None):
        """rows[R][6] = (first_entity, start1, end1, second_entity, start2, end2), one per replica."""
        self._apply(self.lib.sfgpu_apply_sublist_swap, self.pack_sublist_swap(rows), 4, mask)

    # ---- state read-back ------------------------------------------------------------------
    def scalar_state(self) -> np.ndarray:
        out = np.zeros((self.R, self.n_entities), dtype=np.int32)
        self._check(self.lib.sfgpu_get_scalar_state(self.h, self.scalar_var, _ptr(out)))
        return out

    def list_state(self):
        cap = C.c_uint32()
        self._check(self.lib.sfgpu_list_capacity(self.h, self.list_var, C.byref(cap)))
        offs = np.zeros((self.R, self.n_owners + 1), dtype=np.uint32)
        elems = np.zeros((self.R, cap.value), dtype=np.uint32)
        self._check(self.lib.sfgpu_get_list_state(self.h, self.list_var, _ptr(offs), _ptr(elems)))
        return offs, elems

    def last_kernel_ns(self) -> int:
        out = C.c_uint64()
        self._check(self.lib.sfgpu_last_kernel_ns(self.h, C.byref(out)))
        return out.value

    def kernel_times_ns(self, max_n: int = 512) -> np.ndarray:
        """Device durations of the most recent scoring-kernel launches (oldest first)."""
        out = np.zeros(max_n, dtype=np.uint64)
        n = C.c_uint32()
        self._check(self.lib.sfgpu_kernel_times_ns(self.h, max_n, _ptr(out), C.byref(n)))
        return out[:n.value]

    def launch_count(self) -> int:
        out = C.c_uint64()
        self._check(self.lib.sfgpu_launch_count(self.h, C.byref(out)))
        return out.value

    def synchronize(self):
        self._check(self.lib.sfgpu_synchronize(self.h))

    def scalar_program(self) -> int:
        """>= 0: the monomorphised scalar scoring kernel the model runs; -1: the interpreter."""
        out = C.c_int32()
        self._check(self.lib.sfgpu_scalar_program(self.h, C.byref(out)))
        return out.value


@dataclass
class ForageParams:
    """acceptor: 0 accept-all, 1 HillClimbing, 2 LateAcceptance; accepted_limit 0 = BestScore forager,
    N = AcceptedCount(N); tie_mode 1 = reservoir ties (forager.rs:99-155)."""
    acceptor: int = 0
    tie_mode: int = 1
    accepted_limit: int = 0


# ---------------------------------------------------------------------------------------------
# Fluent ConstraintFactory (stream/factory.rs:43-73, uni_stream/, cross_bi_stream/, grouped.rs)
# ---------------------------------------------------------------------------------------------
class ConstraintFactory:
    def __init__(self, director: GpuScoreDirector):
        self.d = director

    def for_each(self, collection: int) -> "UniStream":
        return UniStream(self.d, collection)


@dataclass
class AdjacentEqual:
    """|l, r| l.id < r.id && l.neighbors.contains(r.id) && l.var.is_some() && l.var == r.var"""
    csr: int


@dataclass
class EqualKey:
    """equal(key) joiner with key(e) = column[e]*col_mul + var[e]*var_mul, plus l.id < r.id && var.is_some()"""
    column: int = L.NO_COLUMN
    col_mul: int = 0
    var_mul: int = 1
    arity: int = 2   # 3 / 4 / 5 after further .join(.., same joiner): tri / quad / penta self-join


class Expr:
    """Column expression over a joined pair (a, b): the data form of the reference's pair filter / pair weight
    closures (stream/filter/adapters.rs:63-93). Built with the static constructors and Python operators; lowers to
    the postfix program of sfgpu_add_expr."""

    def __init__(self, ops):
        self.ops = list(ops)

    @staticmethod
    def const(v: int) -> "Expr":
        return Expr([(L.X_CONST, 0, int(v))])

    @staticmethod
    def a(column: int) -> "Expr":
        return Expr([(L.X_A_COL, column, 0)])

    @staticmethod
    def b(column: int) -> "Expr":
        return Expr([(L.X_B_COL, column, 0)])

    @staticmethod
    def a_index() -> "Expr":
        return Expr([(L.X_A_IDX, 0, 0)])

    @staticmethod
    def b_index() -> "Expr":
        return Expr([(L.X_B_IDX, 0, 0)])

    @staticmethod
    def value() -> "Expr":
        return Expr([(L.X_VALUE, 0, 0)])

    @staticmethod
    def a_value() -> "Expr":
        """planning value of row a (the left row of a self-join)"""
        return Expr([(L.X_A_VAL, 0, 0)])

    @staticmethod
    def b_value() -> "Expr":
        return Expr([(L.X_B_VAL, 0, 0)])

    @staticmethod
    def _lift(x) -> "Expr":
        return x if isinstance(x, Expr) else Expr.const(x)

    def _bin(self, other, op) -> "Expr":
        return Expr(self.ops + Expr._lift(other).ops + [(op, 0, 0)])

    def __add__(self, o): return self._bin(o, L.X_ADD)
    def __sub__(self, o): return self._bin(o, L.X_SUB)
    def __mul__(self, o): return self._bin(o, L.X_MUL)
    def __mod__(self, o): return self._bin(o, L.X_MOD)
    def __neg__(self): return Expr(self.ops + [(L.X_NEG, 0, 0)])
    def __abs__(self): return Expr(self.ops + [(L.X_ABS, 0, 0)])
    def eq(self, o): return self._bin(o, L.X_EQ)
    def ne(self, o): return self._bin(o, L.X_NE)
    def __lt__(self, o): return self._bin(o, L.X_LT)
    def __le__(self, o): return self._bin(o, L.X_LE)
    def __gt__(self, o): return self._bin(o, L.X_GT)
    def __ge__(self, o): return self._bin(o, L.X_GE)
    def __and__(self, o): return self._bin(o, L.X_AND)
    def __or__(self, o): return self._bin(o, L.X_OR)
    def __invert__(self): return Expr(self.ops + [(L.X_NOT, 0, 0)])
    def min(self, o): return self._bin(o, L.X_MIN)
    def max(self, o): return self._bin(o, L.X_MAX)

    @staticmethod
    def csr_contains(csr: int, row, x) -> "Expr":
        """csr.row(row).contains(x), e.g. employee.unavailable_days.contains(shift.day)"""
        return Expr(Expr._lift(row).ops + Expr._lift(x).ops + [(L.X_CSR_CONTAINS, csr, 0)])

    @staticmethod
    def select(cond, then, other) -> "Expr":
        return Expr(Expr._lift(cond).ops + Expr._lift(then).ops + Expr._lift(other).ops + [(L.X_SELECT, 0, 0)])


@dataclass
class EqualKeyExpr:
    """Self-join of the entity rows on key expressions of one row (Expr over a(..) / a_value() / a_index()):
    equal(key) when right_key is None (every pair once, left = lower index; nary_incremental/bi.rs, projected/bi.rs),
    equal_bi(left_key, right_key) otherwise (ordered pairs, projected/directed_bi.rs). A key outside [0, n_keys) is
    None."""
    left_key: "Expr"
    n_keys: int
    right_key: Optional["Expr"] = None


@dataclass
class EqualVarToKey:
    """equal_bi(|a| a.var, |b| Some(b.key)) where several B rows may share a key: bucket_csr row k lists the B rows
    whose key is k (cross_bi_incremental/state.rs: b rows indexed by key)."""
    bucket_csr: int


@dataclass
class EqualVarToRow:
    """equal_bi(|e| e.var, |v| Some(v.row)) — joins an entity with the value row it is assigned to."""
    pass


@dataclass
class EqualId:
    """equal_bi(|a| a.id, |element| *element); id column or the row index."""
    column: int = L.NO_COLUMN


@dataclass
class PathCost:
    matrix: int
    depot: int


@dataclass
class ListSum:
    column: int


@dataclass
class Projection:
    """Data form of a `Projection<A>` (stream/projected_stream/source.rs:13-24): an ASSIGNED entity e emits one
    row per entry j of csr row e — key offset csr.col[j] (< keys_per_value), amount amounts[j] (or 1). The group
    key of a row is var[e] * keys_per_value + key offset, e.g. (nurse, day). MAX_EMITS <= 8."""
    csr: int
    keys_per_value: int
    amounts: Optional[np.ndarray] = None   # int64 per csr entry


@dataclass
class Count:
    pass


@dataclass
class Sum:
    column: int


@dataclass
class ConsecutiveRuns:
    """consecutive_runs(|e| e.point) (stream/collector/runs.rs): the weight applies to every run's point_count
    and the group scores their sum. point_column: entity column with 0 <= point < n_points."""
    point_column: int
    n_points: int


@dataclass
class IndexedPresence:
    """indexed_presence(|e| e.point) (stream/collector/indexed_presence.rs) scored through one of its views:
    "complement_runs" — the weight applies to every run of absent points inside [lo, hi) and the group scores their
    sum; "any_in" — weight(1) when a point of [lo, hi) is present; "count" — weight(distinct points)."""
    point_column: int
    n_points: int
    view: str = "count"
    lo: int = 0
    hi: int = 0


@dataclass
class LoadBalance:
    metric_column: int = L.NO_COLUMN


class _Terminal:
    def __init__(self, d: GpuScoreDirector, **kw):
        self.d, self.kw = d, kw

    def named(self, name: str) -> int:
        return self.d.add_constraint(name=name, **self.kw)


class UniStream:
    def __init__(self, d: GpuScoreDirector, collection: int, filt: int = 2, mask: int = L.NO_COLUMN):
        self.d, self.collection, self.filt, self.mask = d, collection, filt, mask

    def unassigned(self) -> "UniStream":
        return UniStream(self.d, self.collection, 0, self.mask)

    def assigned(self) -> "UniStream":
        return UniStream(self.d, self.collection, 1, self.mask)

    def filter(self, mask_column: int) -> "UniStream":
        """`.filter(|e| e.flag)` over a static 0/1 fact column."""
        return UniStream(self.d, self.collection, self.filt, mask_column)

    def flattened(self) -> "UniStream":
        return UniStream(self.d, self.collection, self.filt, self.mask)

    def _impact(self, impact: int, weight, x=None, by_value: bool = False) -> _Terminal:
        if isinstance(x, PathCost):
            return _Terminal(self.d, kind=L.K_LIST_PATH_COST, impact=impact, weight=weight, collection=self.collection,
                             variable=L.LIST_VAR, aux0=x.matrix, p0=x.depot)
        if isinstance(x, ListSum):
            return _Terminal(self.d, kind=L.K_LIST_SUM, impact=impact, weight=weight, collection=self.collection,
                             variable=L.LIST_VAR, aux0=x.column)
        w = _const_weight(weight) if isinstance(weight, HardSoftScore) else weight
        return _Terminal(self.d, kind=L.K_UNI, impact=impact, weight=w, collection=self.collection,
                         aux0=L.NO_COLUMN if x is None else x, aux1=self.mask, p0=self.filt, p1=1 if by_value else 0)

    def penalize(self, weight, x=None, by_value: bool = False) -> _Terminal:
        return self._impact(L.PENALTY, weight, x, by_value)

    def reward(self, weight, x=None, by_value: bool = False) -> _Terminal:
        return self._impact(L.REWARD, weight, x, by_value)

    def join(self, other, joiner) -> "BiStream":
        return BiStream(self.d, self.collection, other, joiner)

    def if_exists(self, other: "UniStream", joiner) -> "ExistsStream":
        if isinstance(joiner, (EqualVarToRow, EqualVarToKey)):
            return DirectExistsStream(self.d, self.collection, other, joiner, 0)
        return ExistsStream(self.d, self.collection, joiner, 0)

    def if_not_exists(self, other: "UniStream", joiner) -> "ExistsStream":
        if isinstance(joiner, (EqualVarToRow, EqualVarToKey)):
            return DirectExistsStream(self.d, self.collection, other, joiner, 1)
        return ExistsStream(self.d, self.collection, joiner, 1)

    def group_by(self, collector) -> "GroupedStream":
        # for_each(E).group_by(|e| e.var, collector) over assigned entities
        return GroupedStream(self.d, self.collection, collector)

    def project(self, projection: "Projection") -> "ProjectedStream":
        """for_each(E).project(P) — projected scoring rows (stream/projected_stream/uni.rs)."""
        return ProjectedStream(self.d, self.collection, projection)


class ProjectedStream:
    def __init__(self, d, collection, projection: "Projection"):
        self.d, self.collection, self.p = d, collection, projection

    def _impact(self, impact, weight: "WeightFn") -> _Terminal:
        # .project(P).penalize(|row| w(row.amount)) scores every emitted row on its own
        # (constraint/projected/uni.rs:61-263): with a LINEAR weight (b = 0) or a CONST weight the rows of an
        # entity sum to one per-entity weight, i.e. a uni constraint over assigned entities.
        if weight.fn not in (L.W_CONST, L.W_LINEAR) or (weight.fn == L.W_LINEAR and weight.b != 0):
            raise L.SfgpuError(L.E_UNSUPPORTED, "projected row weights must be CONST or LINEAR (a * amount)")
        rp, _ = self.d._csr_host[self.p.csr]
        nnz = int(rp[-1])
        amt = np.ones(nnz, dtype=np.int64) if self.p.amounts is None else np.asarray(self.p.amounts, dtype=np.int64)
        per_row = np.full(nnz, weight.a, dtype=np.int64) if weight.fn == L.W_CONST else amt
        sums = np.add.reduceat(np.concatenate([per_row, [0]]), rp[:-1].astype(np.int64)) if nnz else np.zeros(len(rp) - 1, np.int64)
        sums = np.where(np.diff(rp.astype(np.int64)) > 0, sums, 0)
        col = self.d.add_column(self.collection, f"projected_row_sum_{self.p.csr}_{id(self) & 0xFFFF}", sums)
        w = WeightFn(L.W_LINEAR, weight.level, 1 if weight.fn == L.W_CONST else weight.a, 0)
        return _Terminal(self.d, kind=L.K_UNI, impact=impact, weight=w, collection=self.collection, aux0=col,
                         aux1=L.NO_COLUMN, p0=1, p1=0)

    def penalize(self, weight: "WeightFn") -> _Terminal:
        return self._impact(L.PENALTY, weight)

    def reward(self, weight: "WeightFn") -> _Terminal:
        return self._impact(L.REWARD, weight)

    def join_self(self) -> "ProjectedPairStream":
        """.join(equal(|row| row.key)) — keyed self-join of the projected rows (constraint/projected/bi.rs): every
        unordered pair of rows with equal keys, rows of one entity included."""
        return ProjectedPairStream(self.d, self.collection, self.p)

    def group_by(self, collector) -> "ProjectedGroupedStream":
        """.group_by(|row| row.key, count() | sum(|row| row.amount))"""
        if not isinstance(collector, (Count, Sum)):
            raise L.SfgpuError(L.E_UNSUPPORTED, "projected group_by supports count() and sum(amount)")
        return ProjectedGroupedStream(self.d, self.collection, self.p, collector)


class ProjectedPairStream:
    def __init__(self, d, collection, projection: "Projection"):
        self.d, self.collection, self.p = d, collection, projection

    def _impact(self, impact, weight: HardSoftScore) -> _Terminal:
        # n rows in a key group form n(n-1)/2 pairs: the grouped count with the triangular weight
        w = _const_weight(weight)
        return _Terminal(self.d, kind=L.K_PROJECT_GROUP, impact=impact, weight=WeightFn(L.W_PAIRS, w.level, w.a, 0),
                         collection=self.collection, aux0=self.p.csr, aux1=L.NO_COLUMN, p0=self.p.keys_per_value, p1=0)

    def penalize(self, weight: HardSoftScore) -> _Terminal:
        return self._impact(L.PENALTY, weight)

    def reward(self, weight: HardSoftScore) -> _Terminal:
        return self._impact(L.REWARD, weight)


class ProjectedGroupedStream:
    def __init__(self, d, collection, projection: "Projection", collector):
        self.d, self.collection, self.p, self.collector = d, collection, projection, collector

    def _impact(self, impact, weight: "WeightFn") -> _Terminal:
        aux1 = L.NO_COLUMN
        if isinstance(self.collector, Sum):
            rp, _ = self.d._csr_host[self.p.csr]
            nnz = int(rp[-1])
            rows = self.d.add_collection(f"projected_rows_{self.p.csr}_{id(self) & 0xFFFF}", max(nnz, 1), -1)
            amt = np.zeros(max(nnz, 1), dtype=np.int64)
            amt[:nnz] = np.asarray(self.p.amounts, dtype=np.int64)
            aux1 = self.d.add_column(rows, "amount", amt)
        return _Terminal(self.d, kind=L.K_PROJECT_GROUP, impact=impact, weight=weight, collection=self.collection,
                         aux0=self.p.csr, aux1=aux1, p0=self.p.keys_per_value, p1=0)

    def penalize(self, weight: "WeightFn") -> _Terminal:
        return self._impact(L.PENALTY, weight)

    def reward(self, weight: "WeightFn") -> _Terminal:
        return self._impact(L.REWARD, weight)


class ExistsStream:
    def __init__(self, d, collection, joiner: EqualId, mode: int):
        self.d, self.collection, self.joiner, self.mode = d, collection, joiner, mode

    def penalize(self, weight: HardSoftScore) -> _Terminal:
        return _Terminal(self.d, kind=L.K_EXISTS_FLAT, impact=L.PENALTY, weight=_const_weight(weight),
                         collection=self.collection, variable=L.LIST_VAR, aux0=self.joiner.column, p0=self.mode)

    def reward(self, weight: HardSoftScore) -> _Terminal:
        return _Terminal(self.d, kind=L.K_EXISTS_FLAT, impact=L.REWARD, weight=_const_weight(weight),
                         collection=self.collection, variable=L.LIST_VAR, aux0=self.joiner.column, p0=self.mode)


class DirectExistsStream:
    """for_each(A).if_[not_]exists(for_each(B).filter(fb), equal_bi(|a| a.var, |b| Some(b.key))) — the direct
    (non-flattened) exists of the reference (constraint/exists.rs:167-272) with the planning variable as the A key and
    a fact collection on the B side. B never changes during a solve, so the per-key count of filter-passing B rows is
    a constant table: the constraint lowers on the host to a uni constraint over the per-value indicator
    [b_count(key) > 0] (exists) / [b_count(key) == 0] (not exists) — exists.rs:148-159."""

    def __init__(self, d, collection, other: "UniStream", joiner, mode: int):
        self.d, self.collection, self.other, self.joiner, self.mode = d, collection, other, joiner, mode

    def _impact(self, impact, weight: HardSoftScore) -> _Terminal:
        d = self.d
        n_b = d.coll_rows[self.other.collection]
        passing = np.ones(n_b, dtype=np.int64) if self.other.mask == L.NO_COLUMN else \
            (d._col_host[self.other.mask] != 0).astype(np.int64)
        if isinstance(self.joiner, EqualVarToKey):
            rp, ci = d._csr_host[self.joiner.bucket_csr]
            count = np.array([int(passing[ci[rp[k]:rp[k + 1]]].sum()) for k in range(len(rp) - 1)], dtype=np.int64)
            values_coll = None
        else:
            count = passing
            values_coll = self.other.collection
        indicator = (count > 0).astype(np.int64) if self.mode == 0 else (count == 0).astype(np.int64)
        if values_coll is None:
            values_coll = d.add_collection(f"exists_keys_{len(d.coll_rows)}", len(indicator), -1)
        col = d.add_column(values_coll, "exists_indicator", indicator)
        w = _const_weight(weight)
        return _Terminal(d, kind=L.K_UNI, impact=impact, weight=WeightFn(L.W_LINEAR, w.level, w.a, 0),
                         collection=self.collection, aux0=col, p0=1, p1=1)

    def penalize(self, weight: HardSoftScore) -> _Terminal:
        return self._impact(L.PENALTY, weight)

    def reward(self, weight: HardSoftScore) -> _Terminal:
        return self._impact(L.REWARD, weight)


class BiStream:
    def __init__(self, d, collection, other, joiner, pair_filter: Optional["Expr"] = None):
        self.d, self.collection, self.other, self.joiner = d, collection, other, joiner
        self.pair_filter = pair_filter

    def filter(self, expr: "Expr") -> "BiStream":
        """Pair filter |a, b, index of a, index of b| as a column expression (general cross-collection joins:
        the entity's variable on the A side, a fact collection on the B side)."""
        if not isinstance(self.joiner, (EqualVarToRow, EqualVarToKey, EqualKeyExpr)):
            raise L.SfgpuError(L.E_UNSUPPORTED, "pair filters need the var -> row / var -> key / key-expression joiner")
        f = expr if self.pair_filter is None else (self.pair_filter & expr)
        return BiStream(self.d, self.collection, self.other, self.joiner, f)

    def _impact_expr(self, impact, weight, x: Optional["Expr"]) -> _Terminal:
        """weight: HardSoftScore (constant per pair) or WeightFn of the pair expression x."""
        j = self.joiner
        w = weight if isinstance(weight, WeightFn) else _const_weight(weight)
        fid = L.NO_COLUMN if self.pair_filter is None else self.d.add_expr(self.pair_filter)
        xid = L.NO_COLUMN if x is None else self.d.add_expr(x)
        if isinstance(j, EqualKeyExpr):
            kl = self.d.add_expr(j.left_key)
            kr = L.NO_COLUMN if j.right_key is None else self.d.add_expr(j.right_key)
            return _Terminal(self.d, kind=L.K_PAIR_KEY_EXPR, impact=impact, weight=w, collection=self.collection, aux0=fid,
                             aux1=xid, p0=_i64(kl | (kr << 32)), p1=j.n_keys)
        return _Terminal(self.d, kind=L.K_JOIN_EXPR, impact=impact, weight=w, collection=self.collection, aux0=fid, aux1=xid,
                         p0=j.bucket_csr if isinstance(j, EqualVarToKey) else -1, p1=self.other.collection)

    def _impact(self, impact, weight, x: Optional["Expr"] = None) -> _Terminal:
        j = self.joiner
        if isinstance(j, (EqualVarToRow, EqualVarToKey, EqualKeyExpr)):
            return self._impact_expr(impact, weight, x)
        w = _const_weight(weight)
        if isinstance(j, AdjacentEqual):
            return _Terminal(self.d, kind=L.K_PAIR_CSR_EQUAL, impact=impact, weight=w, collection=self.collection,
                             aux0=j.csr)
        if isinstance(j, EqualKey):
            return _Terminal(self.d, kind=L.K_PAIR_KEY_EQUAL, impact=impact, weight=w, collection=self.collection,
                             aux0=j.column, aux1=j.arity, p0=j.col_mul, p1=j.var_mul)
        raise L.SfgpuError(L.E_UNSUPPORTED, f"joiner {type(j).__name__} is not expressible on device")

    def penalize(self, weight, x: Optional["Expr"] = None) -> _Terminal:
        return self._impact(L.PENALTY, weight, x)

    def reward(self, weight, x: Optional["Expr"] = None) -> _Terminal:
        return self._impact(L.REWARD, weight, x)

    def join(self, other, joiner) -> "BiStream":
        """Third / fourth / fifth member of a keyed self-join (TriConstraintStream .. PentaConstraintStream,
        nary_incremental/higher_arity/shared.rs): same collection, same equal(key) joiner, index-ordered tuples."""
        j = self.joiner
        if not isinstance(j, EqualKey) or not isinstance(joiner, EqualKey) or \
                (joiner.column, joiner.col_mul, joiner.var_mul) != (j.column, j.col_mul, j.var_mul):
            raise L.SfgpuError(L.E_UNSUPPORTED, "higher-arity joins need the same equal(key) joiner on every member")
        if j.arity >= 5:
            raise L.SfgpuError(L.E_UNSUPPORTED, "joins beyond penta are not part of the reference API")
        return BiStream(self.d, self.collection, other, EqualKey(j.column, j.col_mul, j.var_mul, j.arity + 1))

    def group_by(self, collector) -> "GroupedStream":
        if not isinstance(self.joiner, EqualVarToRow):
            raise L.SfgpuError(L.E_UNSUPPORTED, "group_by over a join needs the var -> value-row joiner")
        return GroupedStream(self.d, self.collection, collector)


class GroupedStream:
    def __init__(self, d, collection, collector, complemented: bool = False, default: int = 0):
        self.d, self.collection, self.collector = d, collection, collector
        self.complemented, self.default = complemented, default

    def complement(self, targets: int, default: int = 0) -> "GroupedStream":
        return GroupedStream(self.d, self.collection, self.collector, True, default)

    def _impact(self, impact, weight: WeightFn, key_offset_column: int = L.NO_COLUMN) -> _Terminal:
        c = self.collector
        if isinstance(c, ConsecutiveRuns):
            if self.complemented:
                raise L.SfgpuError(L.E_UNSUPPORTED, "consecutive_runs with a complement is not expressible on device")
            return _Terminal(self.d, kind=L.K_RUNS, impact=impact, weight=weight, collection=self.collection,
                             aux0=c.point_column, p0=c.n_points)
        if isinstance(c, IndexedPresence):
            if self.complemented:
                raise L.SfgpuError(L.E_UNSUPPORTED, "indexed_presence with a complement is not expressible on device")
            view = {"complement_runs": 1, "any_in": 2, "count": 3}[c.view]
            return _Terminal(self.d, kind=L.K_RUNS, impact=impact, weight=weight, collection=self.collection,
                             aux0=c.point_column, aux1=view, p0=c.n_points, p1=c.lo | (c.hi << 32))
        if isinstance(c, LoadBalance):
            return _Terminal(self.d, kind=L.K_LOAD_BALANCE, impact=impact, weight=weight, collection=self.collection,
                             aux0=c.metric_column)
        col = L.NO_COLUMN if isinstance(c, Count) else c.column
        return _Terminal(self.d, kind=L.K_GROUP, impact=impact, weight=weight, collection=self.collection, aux0=col,
                         aux1=key_offset_column, p0=1 if self.complemented else 0, p1=self.default)

    def penalize(self, weight: WeightFn, key_offset_column: int = L.NO_COLUMN) -> _Terminal:
        """key_offset_column: per-value column replacing the weight's b for that key (|key, result| weights)."""
        return self._impact(L.PENALTY, weight, key_offset_column)

    def reward(self, weight: WeightFn, key_offset_column: int = L.NO_COLUMN) -> _Terminal:
        return self._impact(L.REWARD, weight, key_offset_column)
