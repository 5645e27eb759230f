"""ctypes binding of libsfgpu.so (the C ABI in include/sfgpu.h).

The library is built in-tree by ``__graft_entry__.build()`` (``make -C solverforge_b200/csrc``).
There is no CPU fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsfgpu.so")

OK = 0
E_INVALID, E_UNSUPPORTED, E_CUDA, E_NCCL, E_OOM, E_STATE = -1, -2, -3, -4, -5, -6
DEVICE_IO = 1
SYNC_ASYNC = 2
CTX_LEGACY_DEFAULT_STREAM, CTX_GENERIC_KERNELS = 1, 2

W_CONST, W_LINEAR, W_SQUARE, W_EXCESS, W_ABSDIFF, W_PAIRS = 0, 1, 2, 3, 4, 5
PENALTY, REWARD = 0, 1
K_UNI, K_PAIR_CSR_EQUAL, K_PAIR_KEY_EQUAL, K_EXISTS_FLAT, K_GROUP = 1, 2, 3, 4, 5
K_LIST_PATH_COST, K_LIST_SUM, K_LOAD_BALANCE, K_PROJECT_GROUP, K_RUNS = 6, 7, 8, 9, 10
K_JOIN_EXPR, K_PAIR_KEY_EXPR = 11, 12
X_CONST, X_A_COL, X_B_COL, X_A_IDX, X_B_IDX, X_VALUE, X_A_VAL, X_B_VAL = 1, 2, 3, 4, 5, 6, 7, 8
X_ADD, X_SUB, X_MUL, X_NEG, X_ABS, X_MIN, X_MAX, X_MOD = 10, 11, 12, 13, 14, 15, 16, 17
X_EQ, X_NE, X_LT, X_LE, X_GT, X_GE = 20, 21, 22, 23, 24, 25
X_AND, X_OR, X_NOT, X_CSR_CONTAINS, X_SELECT = 30, 31, 32, 40, 41
NO_COLUMN = 0xFFFFFFFF
LIST_VAR = 0x80000000
MAX_EDITS = 8


class Weight(C.Structure):
    _fields_ = [("fn", C.c_int32), ("level", C.c_int32), ("a", C.c_int64), ("b", C.c_int64)]


class ConstraintDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("impact", C.c_int32),
        ("weight", Weight),
        ("collection", C.c_uint32),
        ("variable", C.c_uint32),
        ("aux0", C.c_uint32),
        ("aux1", C.c_uint32),
        ("p0", C.c_int64),
        ("p1", C.c_int64),
        ("name", C.c_char_p),
    ]


class SolveParams(C.Structure):
    _fields_ = [("max_nearby", C.c_uint32), ("n_steps", C.c_uint32), ("acceptor", C.c_int32),
                ("late_size", C.c_uint32), ("tie_mode", C.c_int32), ("accepted_limit", C.c_uint32),
                ("seed_base", C.c_uint64), ("restore_best", C.c_int32), ("reserved", C.c_int32),
                ("acceptor_real", C.c_double), ("step_count_limit", C.c_uint64)]


class ExprOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("arg", C.c_uint32), ("imm", C.c_int64)]


class UnionChild(C.Structure):
    _fields_ = [("family", C.c_int32), ("p0", C.c_uint32), ("p1", C.c_uint32), ("reserved", C.c_uint32),
                ("weight", C.c_uint64)]


class UnionDesc(C.Structure):
    _fields_ = [("n_children", C.c_uint32), ("union_order", C.c_int32), ("selection_order", C.c_int32),
                ("window", C.c_uint32), ("max_window", C.c_uint32), ("reserved", C.c_uint32),
                ("children", UnionChild * 8)]


FAM_NEARBY_LIST_CHANGE, FAM_NEARBY_LIST_SWAP, FAM_SUBLIST_CHANGE, FAM_SUBLIST_SWAP, FAM_LIST_REVERSE = 0, 1, 2, 3, 4
FAM_K_OPT, FAM_CHANGE, FAM_SWAP = 5, 6, 7
ORDER_ORIGINAL, ORDER_RANDOM, ORDER_SHUFFLED = 0, 1, 2
UNION_SEQUENTIAL, UNION_ROUND_ROBIN, UNION_ROTATING_ROUND_ROBIN, UNION_RANDOM, UNION_STRATIFIED_RANDOM = 0, 1, 2, 3, 4


class ForageParams(C.Structure):
    _fields_ = [("acceptor", C.c_int32), ("tie_mode", C.c_int32), ("accepted_limit", C.c_uint32),
                ("reserved", C.c_uint32)]


# every symbol include/sfgpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "sfgpu_ctx_create": (C.c_int32, [C.c_int32, C.c_uint64, _P, C.POINTER(_P)]),
    "sfgpu_ctx_destroy": (C.c_int32, [_P]),
    "sfgpu_last_error": (C.c_char_p, [_P]),
    "sfgpu_abi_version": (C.c_int32, []),
    "sfgpu_synchronize": (C.c_int32, [_P]),
    "sfgpu_model_begin": (C.c_int32, [_P, C.c_uint32]),
    "sfgpu_add_collection": (C.c_int32, [_P, C.c_char_p, C.c_uint32, C.c_int32, C.POINTER(C.c_uint32)]),
    "sfgpu_add_column_i64": (C.c_int32, [_P, C.c_uint32, C.c_char_p, _P, C.POINTER(C.c_uint32)]),
    "sfgpu_add_scalar_variable": (C.c_int32, [_P, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int32,
                                              C.POINTER(C.c_uint32)]),
    "sfgpu_add_list_variable": (C.c_int32, [_P, C.c_uint32, C.c_uint32, C.c_char_p, C.POINTER(C.c_uint32)]),
    "sfgpu_add_csr": (C.c_int32, [_P, C.c_char_p, C.c_uint32, _P, _P, C.POINTER(C.c_uint32)]),
    "sfgpu_add_matrix_i64": (C.c_int32, [_P, C.c_char_p, C.c_uint32, C.c_uint32, _P, C.c_int32,
                                         C.POINTER(C.c_uint32)]),
    "sfgpu_add_expr": (C.c_int32, [_P, C.POINTER(ExprOp), C.c_uint32, C.POINTER(C.c_uint32)]),
    "sfgpu_add_constraint": (C.c_int32, [_P, C.POINTER(ConstraintDesc), C.POINTER(C.c_uint32)]),
    "sfgpu_set_scalar_state": (C.c_int32, [_P, C.c_uint32, _P, C.c_int32]),
    "sfgpu_set_list_state": (C.c_int32, [_P, C.c_uint32, _P, _P, C.c_int32]),
    "sfgpu_model_commit": (C.c_int32, [_P, _P]),
    "sfgpu_score_change": (C.c_int32, [_P, C.c_uint32, C.c_uint64, _P, _P, _P, _P]),
    "sfgpu_score_swap": (C.c_int32, [_P, C.c_uint32, C.c_uint64, _P, _P, _P, _P]),
    "sfgpu_score_compound": (C.c_int32, [_P, C.c_uint32, C.c_uint64, _P, _P, _P, _P, _P]),
    "sfgpu_score_list_change": (C.c_int32, [_P, C.c_uint32, C.c_uint64, _P, _P, _P, _P]),
    "sfgpu_score_list_swap": (C.c_int32, [_P, C.c_uint32, C.c_uint64, _P, _P, _P, _P]),
    "sfgpu_score_list_reverse": (C.c_int32, [_P, C.c_uint32, C.c_uint64, _P, _P, _P, _P]),
    "sfgpu_score_sublist_change": (C.c_int32, [_P, C.c_uint32, C.c_uint64, _P, _P, _P, _P]),
    "sfgpu_score_sublist_swap": (C.c_int32, [_P, C.c_uint32, C.c_uint64, _P, _P, _P, _P]),
    "sfgpu_score_k_opt": (C.c_int32, [_P, C.c_uint32, C.c_uint64, _P, _P, _P, _P]),
    "sfgpu_argbest": (C.c_int32, [_P, C.c_uint32, C.POINTER(ForageParams), _P, _P, _P, _P, _P, _P, _P, _P]),
    "sfgpu_argbest_gated": (C.c_int32, [_P, C.c_uint32, C.POINTER(ForageParams), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "sfgpu_step_list_change": (C.c_int32, [_P, C.c_uint64, _P, _P, C.POINTER(ForageParams), _P, _P, _P, _P, _P, _P,
                                           _P]),
    "sfgpu_step_change_rows": (C.c_int32, [_P, C.c_uint64, _P, _P, C.POINTER(ForageParams), _P, _P, _P, _P, _P, _P,
                                           _P]),
    "sfgpu_step_nearby_list_change": (C.c_int32, [_P, C.c_uint32, C.c_uint32, C.POINTER(ForageParams), _P, _P, _P, _P,
                                                  _P, _P, _P, _P, _P, _P, C.c_int32]),
    "sfgpu_step_nearby_list_swap": (C.c_int32, [_P, C.c_uint32, C.c_uint32, C.POINTER(ForageParams), _P, _P, _P, _P, _P,
                                                _P, _P, _P, _P, _P, C.c_int32]),
    "sfgpu_step_change": (C.c_int32, [_P, C.c_uint32, C.POINTER(ForageParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                      C.c_int32]),
    "sfgpu_step_sublist_change": (C.c_int32, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(ForageParams), _P, _P,
                                              _P, _P, _P, _P, C.c_int32]),
    "sfgpu_step_list_reverse": (C.c_int32, [_P, C.c_uint32, C.POINTER(ForageParams), _P, _P, _P, _P, _P, _P, C.c_int32]),
    "sfgpu_step_sublist_swap": (C.c_int32, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(ForageParams), _P, _P,
                                            _P, _P, _P, _P, C.c_int32]),
    "sfgpu_step_union": (C.c_int32, [_P, C.c_uint32, C.POINTER(UnionDesc), C.POINTER(ForageParams), _P, _P, _P, _P, _P, _P,
                                     _P, _P, C.c_int32]),
    "sfgpu_solve_union": (C.c_int32, [_P, C.POINTER(UnionDesc), C.POINTER(SolveParams), _P, _P, _P, _P, _P]),
    "sfgpu_solve_nearby_list_change": (C.c_int32, [_P, C.POINTER(SolveParams), _P, _P, _P]),
    "sfgpu_solve_change": (C.c_int32, [_P, C.POINTER(SolveParams), _P, _P, _P]),
    "sfgpu_apply_change": (C.c_int32, [_P, C.c_uint32, _P, _P]),
    "sfgpu_apply_swap": (C.c_int32, [_P, C.c_uint32, _P, _P]),
    "sfgpu_apply_list_change": (C.c_int32, [_P, C.c_uint32, _P, _P]),
    "sfgpu_apply_list_swap": (C.c_int32, [_P, C.c_uint32, _P, _P]),
    "sfgpu_apply_list_reverse": (C.c_int32, [_P, C.c_uint32, _P, _P]),
    "sfgpu_apply_sublist_change": (C.c_int32, [_P, C.c_uint32, _P, _P]),
    "sfgpu_apply_sublist_swap": (C.c_int32, [_P, C.c_uint32, _P, _P]),
    "sfgpu_apply_k_opt": (C.c_int32, [_P, C.c_uint32, _P, _P]),
    "sfgpu_apply_winners": (C.c_int32, [_P, C.c_int32, _P, _P, _P]),
    "sfgpu_committed_scores": (C.c_int32, [_P, _P]),
    "sfgpu_evaluate_all": (C.c_int32, [_P, _P]),
    "sfgpu_get_scalar_state": (C.c_int32, [_P, C.c_uint32, _P]),
    "sfgpu_get_list_state": (C.c_int32, [_P, C.c_uint32, _P, _P]),
    "sfgpu_list_capacity": (C.c_int32, [_P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "sfgpu_pack_best_keys": (C.c_int32, [_P, _P]),
    "sfgpu_comm_unique_id": (C.c_int32, [_P]),
    "sfgpu_comm_init_rank": (C.c_int32, [C.c_int32, _P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "sfgpu_comm_destroy": (C.c_int32, [_P]),
    "sfgpu_sync_best": (C.c_int32, [_P, _P, C.c_uint32, _P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]),
    "sfgpu_last_kernel_ns": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "sfgpu_kernel_times_ns": (C.c_int32, [_P, C.c_uint32, _P, C.POINTER(C.c_uint32)]),
    "sfgpu_launch_count": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "sfgpu_scalar_program": (C.c_int32, [_P, C.POINTER(C.c_int32)]),
}

_lib = None


class SfgpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libsfgpu error {code}: {message}")
        self.code = code


def load() -> C.CDLL:
    """Loads libsfgpu.so and binds every declared symbol; raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
