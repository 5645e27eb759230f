"""ctypes wrapper over oracle/_build/liboracle.so — TEST INFRASTRUCTURE (checker / CPU baseline).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this. The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
KAT = os.path.join(ORACLE_DIR, "_build", "kat")

_lib = None
_P = C.c_void_p


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        for name in ("sfo_gc_create", "sfo_nq_create", "sfo_cvrp_create", "sfo_js_create", "sfo_shift_create",
                     "sfo_roster_create", "sfo_cluster_create", "sfo_availability_create", "sfo_pairs_create"):
            getattr(l, name).restype = _P
        l.sfo_gc_create.argtypes = [C.c_uint32, C.c_uint32, _P, _P, _P]
        l.sfo_nq_create.argtypes = [C.c_uint32, _P]
        l.sfo_cluster_create.argtypes = [C.c_uint32, C.c_uint32, _P, C.c_uint32, _P]
        l.sfo_cvrp_create.argtypes = [C.c_uint32, C.c_uint32, C.c_int64, C.c_uint32, _P, _P, _P, _P]
        l.sfo_js_create.argtypes = [C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, C.c_int]
        l.sfo_shift_create.argtypes = [C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int64]
        l.sfo_roster_create.argtypes = [C.c_uint32, C.c_uint32, C.c_int64, C.c_int64, _P, _P, _P, _P, _P]
        l.sfo_availability_create.argtypes = [C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, _P, _P, C.c_uint32, _P]
        l.sfo_pairs_create.argtypes = [C.c_uint32, C.c_uint32, _P, _P, _P]
        l.sfo_destroy.argtypes = [_P]
        l.sfo_committed_score.argtypes = [_P, _P]
        l.sfo_evaluate_all.argtypes = [_P, _P]
        l.sfo_score_calculations.argtypes = [_P]
        l.sfo_score_calculations.restype = C.c_uint64
        for name in ("sfo_score_change", "sfo_score_swap"):
            getattr(l, name).argtypes = [_P, C.c_uint64, _P, _P, _P, _P, _P]
        l.sfo_score_compound.argtypes = [_P, C.c_uint64, _P, _P, _P, _P, _P, _P]
        for name in ("sfo_score_list_change", "sfo_score_list_swap"):
            getattr(l, name).argtypes = [_P, C.c_uint64, _P, _P, _P, _P, _P, _P, _P]
        l.sfo_apply_change.argtypes = [_P, C.c_uint32, C.c_int32]
        l.sfo_apply_swap.argtypes = [_P, C.c_uint32, C.c_uint32]
        l.sfo_apply_compound.argtypes = [_P, C.c_uint32, _P, _P]
        l.sfo_apply_list_change.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        l.sfo_apply_list_swap.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        l.sfo_enumerate_change.argtypes = [_P, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, _P, _P]
        l.sfo_enumerate_change.restype = C.c_int64
        l.sfo_enumerate_swap.argtypes = [_P, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, _P, _P]
        l.sfo_enumerate_swap.restype = C.c_int64
        l.sfo_enumerate_k_opt.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, _P]
        l.sfo_enumerate_k_opt.restype = C.c_int64
        l.sfo_score_k_opt.argtypes = [_P, C.c_uint32, C.c_uint64, _P, _P, _P, _P]
        l.sfo_apply_k_opt.argtypes = [_P, C.c_uint32, _P]
        l.sfo_enumerate_nearby_list_change.argtypes = [_P, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, _P,
                                                       _P, _P, _P]
        l.sfo_enumerate_nearby_list_change.restype = C.c_int64
        l.sfo_enumerate_nearby_list_swap.argtypes = l.sfo_enumerate_nearby_list_change.argtypes
        l.sfo_enumerate_nearby_list_swap.restype = C.c_int64
        l.sfo_replay_step.argtypes = [C.c_uint64, _P, _P, _P, _P, _P, _P, C.c_uint64, C.c_int, C.c_uint64, C.c_int,
                                      C.c_int, _P]
        l.sfo_score_list_reverse.argtypes = [_P, C.c_uint64, _P, _P, _P, _P, _P, _P]
        l.sfo_apply_list_reverse.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_uint32]
        l.sfo_enumerate_list_reverse.argtypes = [_P, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, _P, _P, _P]
        l.sfo_enumerate_list_reverse.restype = C.c_int64
        l.sfo_score_sublist_change.argtypes = [_P, C.c_uint64, _P, _P, _P, _P, _P, _P, _P, _P]
        l.sfo_apply_sublist_change.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        l.sfo_enumerate_sublist_change.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int,
                                                   C.c_uint64, _P, _P, _P, _P, _P]
        l.sfo_enumerate_sublist_change.restype = C.c_int64
        l.sfo_score_sublist_swap.argtypes = [_P, C.c_uint64, _P, _P, _P, _P, _P, _P, _P, _P, _P]
        l.sfo_apply_sublist_swap.argtypes = [_P] + [C.c_uint32] * 6
        l.sfo_enumerate_sublist_swap.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int,
                                                 C.c_uint64, _P, _P, _P, _P, _P, _P]
        l.sfo_enumerate_sublist_swap.restype = C.c_int64
        l.sfo_replay_step_gated.argtypes = [C.c_uint64, _P, _P, _P, _P, _P, _P, _P, C.c_uint64, C.c_int, C.c_uint64,
                                            C.c_int, C.c_int, _P]
        l.sfo_acceptor_create.restype = _P
        l.sfo_acceptor_create.argtypes = [C.c_int, C.c_uint64, C.c_double, _P, C.c_int]
        l.sfo_acceptor_destroy.argtypes = [_P]
        l.sfo_acceptor_phase_started.argtypes = [_P, _P]
        l.sfo_acceptor_step.argtypes = [_P, C.c_uint64, _P, _P, _P, _P, _P, _P, C.c_uint64, C.c_int, C.c_uint64,
                                        C.c_int, _P]
        l.sfo_move_signatures.argtypes = [_P, C.c_int, C.c_uint64, _P, _P]
        l.sfo_union_pull_order.argtypes = [C.c_uint32, _P, _P, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, _P, _P]
        l.sfo_union_pull_order.restype = C.c_int64
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(_P)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class Oracle:
    def __init__(self, handle):
        self.h = handle
        self.l = lib()

    def __del__(self):
        if getattr(self, "h", None):
            self.l.sfo_destroy(self.h)
            self.h = None

    # ---- constructors -------------------------------------------------------------------
    @staticmethod
    def graph_coloring(inst, colors=None) -> "Oracle":
        c = np.ascontiguousarray(inst.color if colors is None else colors, dtype=np.int32)
        return Oracle(lib().sfo_gc_create(inst.n, inst.k, _p(_u32(inst.row_ptr)), _p(_u32(inst.col)), _p(c)))

    @staticmethod
    def cluster(inst, team=None) -> "Oracle":
        t = np.ascontiguousarray(inst.team if team is None else team, dtype=np.int32)
        j = np.ascontiguousarray(np.array(inst.joins, dtype=np.int64).reshape(-1))
        return Oracle(lib().sfo_cluster_create(inst.n, inst.n_teams, _p(t), len(inst.joins), _p(j)))

    @staticmethod
    def availability(inst, employee=None) -> "Oracle":
        e = np.ascontiguousarray(inst.employee if employee is None else employee, dtype=np.int32)
        day, req, hrs = (np.ascontiguousarray(x, dtype=np.int64) for x in (inst.day, inst.required, inst.hours))
        skill = np.ascontiguousarray(inst.skill, dtype=np.int64)
        ptr = np.ascontiguousarray(inst.un_ptr, dtype=np.uint32)
        days = np.ascontiguousarray(inst.un_days, dtype=np.int64)
        con = np.ascontiguousarray(inst.contracts, dtype=np.int64).reshape(-1)
        return Oracle(lib().sfo_availability_create(inst.n_shifts, inst.n_employees, _p(day), _p(req), _p(hrs), _p(e), _p(skill),
                                                    _p(ptr), _p(days), len(inst.contracts), _p(con)))

    @staticmethod
    def pairs(demand, prio, bucket, n_buckets) -> "Oracle":
        d, p = np.ascontiguousarray(demand, dtype=np.int64), np.ascontiguousarray(prio, dtype=np.int64)
        b = np.ascontiguousarray(bucket, dtype=np.int32)
        return Oracle(lib().sfo_pairs_create(len(b), n_buckets, _p(d), _p(p), _p(b)))

    @staticmethod
    def nqueens(inst, rows=None) -> "Oracle":
        r = np.ascontiguousarray(inst.row if rows is None else rows, dtype=np.int32)
        return Oracle(lib().sfo_nq_create(inst.n, _p(r)))

    @staticmethod
    def cvrp(inst, offsets=None, elems=None) -> "Oracle":
        o = _u32(inst.offsets if offsets is None else offsets)
        e = _u32(inst.elems if elems is None else elems)
        dm = np.ascontiguousarray(inst.demands, dtype=np.int32)
        mx = np.ascontiguousarray(inst.matrix, dtype=np.int64)
        return Oracle(lib().sfo_cvrp_create(inst.dim, inst.n_routes, inst.capacity, inst.depot, _p(dm), _p(mx), _p(o),
                                            _p(e)))

    @staticmethod
    def job_shop(inst, machine_idx=None, with_complement=True) -> "Oracle":
        m = np.ascontiguousarray(inst.machine_idx if machine_idx is None else machine_idx, dtype=np.int32)
        return Oracle(lib().sfo_js_create(inst.n_ops, inst.n_machines, _p(_u32(inst.job)), _p(_u32(inst.step)), _p(m),
                                          _p(_u32(inst.seq_offsets)), _p(_u32(inst.seq_elems)),
                                          1 if with_complement else 0))

    @staticmethod
    def shift_scheduling(inst, nurse_idx=None, with_load_balance=True, presence_days=0) -> "Oracle":
        n = np.ascontiguousarray(inst.nurse_idx if nurse_idx is None else nurse_idx, dtype=np.int32)
        return Oracle(lib().sfo_shift_create(inst.n_shifts, inst.n_nurses, _p(np.ascontiguousarray(inst.day, np.int64)),
                                             _p(_u32(inst.slot)), _p(np.ascontiguousarray(inst.required, np.uint8)),
                                             _p(np.ascontiguousarray(inst.hours, np.int64)), _p(n), inst.target,
                                             1 if with_load_balance else 0, presence_days))

    @staticmethod
    def roster(inst, nurse_idx=None) -> "Oracle":
        n = np.ascontiguousarray(inst.nurse_idx if nurse_idx is None else nurse_idx, dtype=np.int32)
        return Oracle(lib().sfo_roster_create(inst.n_shifts, inst.n_nurses, inst.n_days, inst.limit,
                                              _p(np.ascontiguousarray(inst.required, np.uint8)), _p(_u32(inst.span_ptr)),
                                              _p(np.ascontiguousarray(inst.span_day, np.int64)),
                                              _p(np.ascontiguousarray(inst.span_hours, np.int64)), _p(n)))

    # ---- scores -------------------------------------------------------------------------
    def committed_score(self) -> np.ndarray:
        out = np.zeros(2, dtype=np.int64)
        self.l.sfo_committed_score(self.h, _p(out))
        return out

    def evaluate_all(self) -> np.ndarray:
        out = np.zeros(2, dtype=np.int64)
        self.l.sfo_evaluate_all(self.h, _p(out))
        return out

    def score_calculations(self) -> int:
        return int(self.l.sfo_score_calculations(self.h))

    def _out(self, n):
        return np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.uint8)

    def score_change(self, rows):
        rows = np.asarray(rows, dtype=np.int64).reshape(-1, 2)
        e, v = _u32(rows[:, 0]), np.ascontiguousarray(rows[:, 1], dtype=np.int32)
        h, s, d = self._out(len(e))
        self.l.sfo_score_change(self.h, len(e), _p(e), _p(v), _p(h), _p(s), _p(d))
        return np.stack([h, s], axis=1), d

    def score_swap(self, rows):
        rows = np.asarray(rows, dtype=np.int64).reshape(-1, 2)
        a, b = _u32(rows[:, 0]), _u32(rows[:, 1])
        h, s, d = self._out(len(a))
        self.l.sfo_score_swap(self.h, len(a), _p(a), _p(b), _p(h), _p(s), _p(d))
        return np.stack([h, s], axis=1), d

    def score_compound(self, edit_offsets, edit_rows):
        eo = _u32(edit_offsets)
        rows = np.asarray(edit_rows, dtype=np.int64).reshape(-1, 2)
        e, v = _u32(rows[:, 0]), np.ascontiguousarray(rows[:, 1], dtype=np.int32)
        n = len(eo) - 1
        h, s, d = self._out(n)
        self.l.sfo_score_compound(self.h, n, _p(eo), _p(e), _p(v), _p(h), _p(s), _p(d))
        return np.stack([h, s], axis=1), d

    def _score_list(self, fn, rows):
        rows = np.asarray(rows, dtype=np.int64).reshape(-1, 4)
        c = [_u32(rows[:, i]) for i in range(4)]
        h, s, d = self._out(len(rows))
        fn(self.h, len(rows), _p(c[0]), _p(c[1]), _p(c[2]), _p(c[3]), _p(h), _p(s), _p(d))
        return np.stack([h, s], axis=1), d

    def score_list_change(self, rows):
        return self._score_list(self.l.sfo_score_list_change, rows)

    def score_list_swap(self, rows):
        return self._score_list(self.l.sfo_score_list_swap, rows)

    def score_list_reverse(self, rows):
        rows = np.asarray(rows, dtype=np.int64).reshape(-1, 4)
        c = [_u32(rows[:, i]) for i in range(3)]
        h, s, d = self._out(len(rows))
        self.l.sfo_score_list_reverse(self.h, len(rows), _p(c[0]), _p(c[1]), _p(c[2]), _p(h), _p(s), _p(d))
        return np.stack([h, s], axis=1), d

    def score_sublist_change(self, rows):
        rows = np.asarray(rows, dtype=np.int64).reshape(-1, 5)
        c = [_u32(rows[:, i]) for i in range(5)]
        h, s, d = self._out(len(rows))
        self.l.sfo_score_sublist_change(self.h, len(rows), _p(c[0]), _p(c[1]), _p(c[2]), _p(c[3]), _p(c[4]), _p(h), _p(s),
                                        _p(d))
        return np.stack([h, s], axis=1), d

    def apply_sublist_change(self, se, start, end, de, dp):
        self.l.sfo_apply_sublist_change(self.h, int(se), int(start), int(end), int(de), int(dp))

    def enumerate_sublist_change(self, min_size=1, max_size=3, step_index=0, step_seed=0, order=0) -> np.ndarray:
        n = self.l.sfo_enumerate_sublist_change(self.h, min_size, max_size, step_index, step_seed, order, 0, None, None,
                                                None, None, None)
        c = [np.zeros(n, dtype=np.uint32) for _ in range(5)]
        self.l.sfo_enumerate_sublist_change(self.h, min_size, max_size, step_index, step_seed, order, n, _p(c[0]),
                                            _p(c[1]), _p(c[2]), _p(c[3]), _p(c[4]))
        return np.stack(c, axis=1)

    def score_sublist_swap(self, rows):
        rows = np.asarray(rows, dtype=np.int64).reshape(-1, 6)
        c = [_u32(rows[:, i]) for i in range(6)]
        h, s, d = self._out(len(rows))
        self.l.sfo_score_sublist_swap(self.h, len(rows), *[_p(x) for x in c], _p(h), _p(s), _p(d))
        return np.stack([h, s], axis=1), d

    def apply_sublist_swap(self, e1, s1, t1, e2, s2, t2):
        self.l.sfo_apply_sublist_swap(self.h, int(e1), int(s1), int(t1), int(e2), int(s2), int(t2))

    def enumerate_sublist_swap(self, min_size=1, max_size=3, step_index=0, step_seed=0, order=0) -> np.ndarray:
        n = self.l.sfo_enumerate_sublist_swap(self.h, min_size, max_size, step_index, step_seed, order, 0, None, None,
                                              None, None, None, None)
        c = [np.zeros(n, dtype=np.uint32) for _ in range(6)]
        self.l.sfo_enumerate_sublist_swap(self.h, min_size, max_size, step_index, step_seed, order, n,
                                          *[_p(x) for x in c])
        return np.stack(c, axis=1)

    def apply_list_reverse(self, e, start, end, *_):
        self.l.sfo_apply_list_reverse(self.h, int(e), int(start), int(end))

    def enumerate_list_reverse(self, step_index=0, step_seed=0, order=0) -> np.ndarray:
        n = self.l.sfo_enumerate_list_reverse(self.h, step_index, step_seed, order, 0, None, None, None)
        c = [np.zeros(n, dtype=np.uint32) for _ in range(3)]
        self.l.sfo_enumerate_list_reverse(self.h, step_index, step_seed, order, n, _p(c[0]), _p(c[1]), _p(c[2]))
        return np.stack(c + [np.zeros(n, dtype=np.uint32)], axis=1)

    # ---- apply --------------------------------------------------------------------------
    def apply_change(self, e, v):
        self.l.sfo_apply_change(self.h, int(e), int(v))

    def apply_swap(self, a, b):
        self.l.sfo_apply_swap(self.h, int(a), int(b))

    def apply_list_change(self, se, sp, de, dp):
        self.l.sfo_apply_list_change(self.h, int(se), int(sp), int(de), int(dp))

    def apply_list_swap(self, e1, p1, e2, p2):
        self.l.sfo_apply_list_swap(self.h, int(e1), int(p1), int(e2), int(p2))

    # ---- candidate order -------------------------------------------------------------------
    def enumerate_change(self, step_index=0, step_seed=0, order=0) -> np.ndarray:
        n = self.l.sfo_enumerate_change(self.h, step_index, step_seed, order, 0, None, None)
        e = np.zeros(n, dtype=np.uint32)
        v = np.zeros(n, dtype=np.int32)
        self.l.sfo_enumerate_change(self.h, step_index, step_seed, order, n, _p(e), _p(v))
        return np.stack([e.astype(np.int64), v.astype(np.int64)], axis=1)

    def enumerate_k_opt(self, k=3, min_seg=1, step_index=0, step_seed=0, order=0) -> np.ndarray:
        n = self.l.sfo_enumerate_k_opt(self.h, k, min_seg, step_index, step_seed, order, 0, None)
        rows = np.zeros((n, k + 2), dtype=np.uint32)
        self.l.sfo_enumerate_k_opt(self.h, k, min_seg, step_index, step_seed, order, n, _p(rows))
        return rows

    def score_k_opt(self, rows, k=3):
        rows = np.ascontiguousarray(rows, dtype=np.uint32).reshape(-1, k + 2)
        h, s, d = self._out(len(rows))
        self.l.sfo_score_k_opt(self.h, k, len(rows), _p(rows), _p(h), _p(s), _p(d))
        return np.stack([h, s], axis=1), d

    def apply_k_opt(self, row, k=3):
        r = np.ascontiguousarray(row, dtype=np.uint32)
        self.l.sfo_apply_k_opt(self.h, k, _p(r))

    def enumerate_swap(self, step_index=0, step_seed=0, order=0) -> np.ndarray:
        n = self.l.sfo_enumerate_swap(self.h, step_index, step_seed, order, 0, None, None)
        a = np.zeros(n, dtype=np.uint32)
        b = np.zeros(n, dtype=np.uint32)
        self.l.sfo_enumerate_swap(self.h, step_index, step_seed, order, n, _p(a), _p(b))
        return np.stack([a.astype(np.int64), b.astype(np.int64)], axis=1)

    def enumerate_nearby_list_swap(self, max_nearby=20, step_index=0, step_seed=0, order=0) -> np.ndarray:
        n = self.l.sfo_enumerate_nearby_list_swap(self.h, max_nearby, step_index, step_seed, order, 0, None, None,
                                                  None, None)
        c = [np.zeros(n, dtype=np.uint32) for _ in range(4)]
        self.l.sfo_enumerate_nearby_list_swap(self.h, max_nearby, step_index, step_seed, order, n, _p(c[0]),
                                              _p(c[1]), _p(c[2]), _p(c[3]))
        return np.stack(c, axis=1)

    def enumerate_nearby_list_change(self, max_nearby=20, step_index=0, step_seed=0, order=0) -> np.ndarray:
        n = self.l.sfo_enumerate_nearby_list_change(self.h, max_nearby, step_index, step_seed, order, 0, None, None,
                                                    None, None)
        c = [np.zeros(n, dtype=np.uint32) for _ in range(4)]
        self.l.sfo_enumerate_nearby_list_change(self.h, max_nearby, step_index, step_seed, order, n, _p(c[0]),
                                                _p(c[1]), _p(c[2]), _p(c[3]))
        return np.stack(c, axis=1)


def replay_step(scores, doable, best_score, last_step_score, late_score, step_seed, forager_kind, accepted_limit,
                random_ties, acceptor_kind, gates=None):
    """Oracle replay of phase/candidates.rs:66-282. Returns (has_winner, winner, moves_evaluated,
    score_calculations, moves_accepted)."""
    scores = np.asarray(scores, dtype=np.int64).reshape(-1, 2)
    h = np.ascontiguousarray(scores[:, 0])
    s = np.ascontiguousarray(scores[:, 1])
    d = np.ascontiguousarray(doable, dtype=np.uint8)
    out = np.zeros(5, dtype=np.uint64)
    b = np.asarray(best_score, dtype=np.int64)
    l_ = np.asarray(last_step_score, dtype=np.int64)
    t = np.asarray(late_score, dtype=np.int64)
    if gates is not None:
        g = np.ascontiguousarray(gates, dtype=np.uint8)
        lib().sfo_replay_step_gated(len(h), _p(h), _p(s), _p(d), _p(g), _p(b), _p(l_), _p(t), step_seed, forager_kind,
                                    accepted_limit, 1 if random_ties else 0, acceptor_kind, _p(out))
        return tuple(int(x) for x in out)
    lib().sfo_replay_step(len(h), _p(h), _p(s), _p(d), _p(b), _p(l_), _p(t), step_seed, forager_kind, accepted_limit,
                          1 if random_ties else 0, acceptor_kind, _p(out))
    return tuple(int(x) for x in out)


def union_pull_order(sizes, union_order, step_index=0, step_seed=0, order=0, weights=None, limit=None):
    """UnionScheduler of the reference (vec_union.rs:190-366) over children of the given sizes: (child[], local[])."""
    sz = np.ascontiguousarray(sizes, dtype=np.uint64)
    w = np.ones(len(sz), dtype=np.uint64) if weights is None else np.ascontiguousarray(weights, dtype=np.uint64)
    cap = int(sz.sum()) if limit is None else int(limit)
    child = np.zeros(max(cap, 1), dtype=np.uint32)
    local = np.zeros(max(cap, 1), dtype=np.uint64)
    n = lib().sfo_union_pull_order(len(sz), _p(sz), _p(w), union_order, step_index, step_seed, order, cap, _p(child),
                                   _p(local))
    return child[:n], local[:n].astype(np.int64)


# families of sfgpu_union_child: (enumerate, score, pack to the device row format)
def union_children(o: "Oracle", children, step_index, step_seed, order):
    """Per child: (rows as the oracle enumerates them, packed device rows[n,4], scores[n,2], doable[n])."""
    out = []
    for ch in children:
        fam = ch[0]
        if fam == 0:
            rows = o.enumerate_nearby_list_change(ch[1], step_index, step_seed, order)
            sc, ok = o.score_list_change(rows) if len(rows) else (np.zeros((0, 2), np.int64), np.zeros(0, np.uint8))
            packed = rows.astype(np.int64)
        elif fam == 1:
            rows = o.enumerate_nearby_list_swap(ch[1], step_index, step_seed, order)
            sc, ok = o.score_list_swap(rows) if len(rows) else (np.zeros((0, 2), np.int64), np.zeros(0, np.uint8))
            packed = rows.astype(np.int64)
        elif fam == 2:
            rows = o.enumerate_sublist_change(ch[1], ch[2], step_index, step_seed, order)
            sc, ok = o.score_sublist_change(rows) if len(rows) else (np.zeros((0, 2), np.int64), np.zeros(0, np.uint8))
            r = rows.astype(np.int64).reshape(-1, 5)
            packed = np.stack([r[:, 0], r[:, 1] | ((r[:, 2] - r[:, 1]) << 24), r[:, 3], r[:, 4]], axis=1)
        elif fam == 3:
            rows = o.enumerate_sublist_swap(ch[1], ch[2], step_index, step_seed, order)
            sc, ok = o.score_sublist_swap(rows) if len(rows) else (np.zeros((0, 2), np.int64), np.zeros(0, np.uint8))
            r = rows.astype(np.int64).reshape(-1, 6)
            packed = np.stack([r[:, 0], r[:, 1] | ((r[:, 2] - r[:, 1]) << 24), r[:, 3], r[:, 4] | ((r[:, 5] - r[:, 4]) << 24)],
                              axis=1)
        elif fam == 6:
            rows = o.enumerate_change(step_index, step_seed, order)
            sc, ok = o.score_change(rows) if len(rows) else (np.zeros((0, 2), np.int64), np.zeros(0, np.uint8))
            r = rows.astype(np.int64).reshape(-1, 2)
            packed = np.stack([r[:, 0], r[:, 1] & 0xFFFFFFFF, np.zeros(len(r), np.int64), np.zeros(len(r), np.int64)], axis=1)
        elif fam == 7:
            rows = o.enumerate_swap(step_index, step_seed, order)
            sc, ok = o.score_swap(rows) if len(rows) else (np.zeros((0, 2), np.int64), np.zeros(0, np.uint8))
            r = rows.astype(np.int64).reshape(-1, 2)
            packed = np.stack([r[:, 0], r[:, 1], np.zeros(len(r), np.int64), np.zeros(len(r), np.int64)], axis=1)
        elif fam == 5:
            k = ch[1]
            rows = o.enumerate_k_opt(k, ch[2], step_index, step_seed, order)
            sc, ok = o.score_k_opt(rows, k) if len(rows) else (np.zeros((0, 2), np.int64), np.zeros(0, np.uint8))
            r = rows.astype(np.int64).reshape(-1, k + 2)
            c5 = np.zeros((len(r), 5), dtype=np.int64)
            c5[:, :k] = r[:, 1:1 + k]
            packed = np.stack([r[:, 0] | (k << 28), c5[:, 0] | (c5[:, 1] << 16), c5[:, 2] | (c5[:, 3] << 16),
                               c5[:, 4] | (r[:, k + 1] << 16)], axis=1)
        else:
            rows = o.enumerate_list_reverse(step_index, step_seed, order)
            sc, ok = o.score_list_reverse(rows) if len(rows) else (np.zeros((0, 2), np.int64), np.zeros(0, np.uint8))
            packed = rows.astype(np.int64).reshape(-1, 4)
        out.append((rows, packed.reshape(-1, 4), np.asarray(sc).reshape(-1, 2), np.asarray(ok)))
    return out


def union_step(o: "Oracle", children, union_order, order, step_index, step_seed, last, late, forager_kind, accepted_limit,
               random_ties, acceptor_kind, weights=None):
    """One reference step over the union cursor: returns (replay outcome, child[], local[], per-child data)."""
    kids = union_children(o, children, step_index, step_seed, order)
    child, local = union_pull_order([len(k[1]) for k in kids], union_order, step_index, step_seed, order, weights)
    sc = np.zeros((len(child), 2), dtype=np.int64)
    ok = np.zeros(len(child), dtype=np.uint8)
    for c, k in enumerate(kids):
        sel = child == c
        sc[sel] = k[2][local[sel]]
        ok[sel] = k[3][local[sel]]
    out = replay_step(sc, ok, [0, 0], last, late, step_seed, forager_kind, accepted_limit, random_ties, acceptor_kind)
    return out, child, local, kids, sc


class OracleAcceptor:
    """Stateful restatement of one reference acceptor (acceptor/*.rs) for multi-step trajectories.
    kind: 0 HillClimbing, 1 LateAcceptance(size), 3 AcceptAll, 4 GreatDeluge(real = rain_speed),
    5 StepCountingHillClimbing(size = limit), 6 DiversifiedLateAcceptance(size, real = tolerance),
    7 TabuSearch(tabu = entity/value/move/undo tenures, aspiration)."""
    HILL_CLIMBING, LATE_ACCEPTANCE, ACCEPT_ALL, GREAT_DELUGE, STEP_COUNTING, DIVERSIFIED_LATE, TABU = 0, 1, 3, 4, 5, 6, 7
    SIMULATED_ANNEALING = 2   # size = calibration samples, real = decay rate; aspiration=2: never accept hard regression

    def __init__(self, kind, size=0, real=0.0, tabu=None, aspiration=True):
        self.l = lib()
        t = np.asarray(tabu if tabu is not None else [0, 0, 0, 0], dtype=np.uint64)
        self.h = self.l.sfo_acceptor_create(kind, size, float(real), _p(t), int(aspiration))

    def __del__(self):
        if getattr(self, "h", None):
            self.l.sfo_acceptor_destroy(self.h)
            self.h = None

    def phase_started(self, initial):
        i = np.asarray(initial, dtype=np.int64)
        self.l.sfo_acceptor_phase_started(self.h, _p(i))

    def step(self, scores, doable, best_score, last_step_score, step_seed, forager_kind, accepted_limit, random_ties,
             signatures=None):
        """replay + acceptor.step_ended; returns (has_winner, winner, moves_evaluated, score_calcs, accepted)."""
        scores = np.asarray(scores, dtype=np.int64).reshape(-1, 2)
        h = np.ascontiguousarray(scores[:, 0])
        s = np.ascontiguousarray(scores[:, 1])
        d = np.ascontiguousarray(doable, dtype=np.uint8)
        b = np.asarray(best_score, dtype=np.int64)
        l_ = np.asarray(last_step_score, dtype=np.int64)
        out = np.zeros(5, dtype=np.uint64)
        sg = None if signatures is None else np.ascontiguousarray(signatures, dtype=np.uint64)
        rc = self.l.sfo_acceptor_step(self.h, len(h), _p(h), _p(s), _p(d), None if sg is None else _p(sg), _p(b), _p(l_),
                                      step_seed, forager_kind, accepted_limit, 1 if random_ties else 0, _p(out))
        assert rc == 0
        return tuple(int(x) for x in out)


def move_signatures(oracle: "Oracle", move_kind: int, rows) -> np.ndarray:
    """tabu signatures (19 u64 each) of ChangeMove (kind 0) / ListChangeMove (kind 2) rows."""
    rows = _u32(rows)
    n = len(rows)
    out = np.zeros((n, 19), dtype=np.uint64)
    lib().sfo_move_signatures(oracle.h, move_kind, n, _p(rows), _p(out))
    return out


# ---- the stronger O(1)-delta CPU baseline (oracle/fast_cpu.cpp) ---------------------------------
FAST_LIB = os.path.join(ORACLE_DIR, "_build", "libfastcpu.so")
_fast = None


def fast_lib() -> C.CDLL:
    global _fast
    if _fast is None:
        if not os.path.exists(FAST_LIB):
            build()
        l = C.CDLL(FAST_LIB)
        l.sfo_fast_cvrp_create.restype = _P
        l.sfo_fast_cvrp_create.argtypes = [C.c_uint32, C.c_uint32, C.c_int64, C.c_uint32, _P, _P, _P, _P]
        l.sfo_fast_cvrp_destroy.argtypes = [_P]
        l.sfo_fast_cvrp_score.argtypes = [_P, C.c_uint64, _P, _P, _P]
        l.sfo_fast_cvrp_bench.restype = C.c_double
        l.sfo_fast_cvrp_bench.argtypes = [_P, C.c_uint64, _P, C.c_uint32, C.c_double]
        l.sfo_fast_cvrp_set_routes.argtypes = [_P, _P, _P, _P]
        l.sfo_fast_cvrp_committed.argtypes = [_P, _P]
        l.sfo_fast_gc_create.restype = _P
        l.sfo_fast_gc_create.argtypes = [C.c_uint32, C.c_uint32, _P, _P]
        l.sfo_fast_gc_destroy.argtypes = [_P]
        l.sfo_fast_gc_score.argtypes = [_P, _P, C.c_uint64, _P, _P, _P, _P]
        l.sfo_fast_js_create.restype = _P
        l.sfo_fast_js_create.argtypes = [C.c_uint32, C.c_uint32, _P, C.c_int64, C.c_int]
        l.sfo_fast_js_destroy.argtypes = [_P]
        l.sfo_fast_js_score.argtypes = [_P, _P, C.c_uint64, _P, _P, _P, _P]
        _fast = l
    return _fast


class FastCvrp:
    """Read-only O(1)-delta CPU scorer for CVRP list-change batches (same algorithm as the GPU fast path)."""

    def __init__(self, inst, offsets=None, elems=None):
        self.l = fast_lib()
        o = _u32(inst.offsets if offsets is None else offsets)
        e = _u32(inst.elems if elems is None else elems)
        dm = np.ascontiguousarray(inst.demands, dtype=np.int32)
        self._dm = dm
        mx = np.ascontiguousarray(inst.matrix, dtype=np.int64)
        self.h = self.l.sfo_fast_cvrp_create(inst.dim, inst.n_routes, inst.capacity, inst.depot, _p(dm), _p(mx), _p(o),
                                             _p(e))
        if not self.h:
            raise ValueError("matrix does not fit the int32 fast path")

    def __del__(self):
        if getattr(self, "h", None):
            self.l.sfo_fast_cvrp_destroy(self.h)
            self.h = None

    def score(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint32).reshape(-1, 4)
        sc = np.zeros((len(rows), 2), dtype=np.int64)
        ok = np.zeros(len(rows), dtype=np.uint8)
        self.l.sfo_fast_cvrp_score(self.h, len(rows), _p(rows), _p(sc), _p(ok))
        return sc, ok

    def set_routes(self, offsets, elems):
        """Rebinds the scorer to another route state (keeps the converted matrix)."""
        o, e = _u32(offsets), _u32(elems)
        self.l.sfo_fast_cvrp_set_routes(self.h, _p(o), _p(e), _p(self._dm))

    def committed_score(self) -> np.ndarray:
        out = np.zeros(2, dtype=np.int64)
        self.l.sfo_fast_cvrp_committed(self.h, _p(out))
        return out

    def bench(self, rows, n_threads: int, seconds: float) -> float:
        rows = np.ascontiguousarray(rows, dtype=np.uint32).reshape(-1, 4)
        return float(self.l.sfo_fast_cvrp_bench(self.h, len(rows), _p(rows), n_threads, seconds))


class FastGraphColoring:
    """O(degree) CPU checker for graph-colouring ChangeMove batches (oracle/fast_cpu.cpp); checked against the
    oracle on small instances."""

    def __init__(self, inst):
        self.l = fast_lib()
        self.n = inst.n
        rp, ci = _u32(inst.row_ptr), _u32(inst.col)
        self.h = self.l.sfo_fast_gc_create(inst.n, inst.k, _p(rp), _p(ci))

    def __del__(self):
        if getattr(self, "h", None):
            self.l.sfo_fast_gc_destroy(self.h)
            self.h = None

    def score_change(self, colors, rows):
        """Returns (scores[n,2], doable[n], committed[2]) for rows[n][2] = (entity, to_value)."""
        col = np.ascontiguousarray(colors, dtype=np.int32)
        r = np.ascontiguousarray(np.asarray(rows).astype(np.int64).astype(np.int32)).reshape(-1, 2)
        sc = np.zeros((len(r), 2), dtype=np.int64)
        ok = np.zeros(len(r), dtype=np.uint8)
        com = np.zeros(2, dtype=np.int64)
        self.l.sfo_fast_gc_score(self.h, _p(col), len(r), _p(r), _p(sc), _p(ok), _p(com))
        return sc, ok, com


class FastJobShop:
    """O(1) CPU checker for job-shop ChangeMove batches (oracle/fast_cpu.cpp)."""

    def __init__(self, inst, with_complement=True):
        self.l = fast_lib()
        job = _u32(inst.job)
        unscheduled = int(inst.n_ops - len(np.unique(inst.seq_elems)))
        self.h = self.l.sfo_fast_js_create(inst.n_ops, inst.n_machines, _p(job), unscheduled, 1 if with_complement else 0)

    def __del__(self):
        if getattr(self, "h", None):
            self.l.sfo_fast_js_destroy(self.h)
            self.h = None

    def score_change(self, machine_idx, rows):
        mi = np.ascontiguousarray(machine_idx, dtype=np.int32)
        r = np.ascontiguousarray(np.asarray(rows).astype(np.int64).astype(np.int32)).reshape(-1, 2)
        sc = np.zeros((len(r), 2), dtype=np.int64)
        ok = np.zeros(len(r), dtype=np.uint8)
        com = np.zeros(2, dtype=np.int64)
        self.l.sfo_fast_js_score(self.h, _p(mi), len(r), _p(r), _p(sc), _p(ok), _p(com))
        return sc, ok, com
