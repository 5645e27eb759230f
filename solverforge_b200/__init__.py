"""solverforge_b200 — B200-native batched re-scoring of local-search candidate moves.

Drop-in for the hot path of SolverForge's incremental ConstraintStream scoring
(solverforge-scoring + solverforge-solver's evaluate_candidate loop). The product path is
libsfgpu.so (hand-written sm_100a CUDA behind the C ABI in include/sfgpu.h); this package is
the host-side mirror of the reference interface above it. There is no CPU scoring fallback.
"""
from . import _lib
from .api import (AdjacentEqual, ConsecutiveRuns, ConstraintFactory, Count, EqualId, EqualKey, EqualKeyExpr, EqualVarToKey, EqualVarToRow, Expr,
                  ForageParams, GpuScoreDirector, HardSoftDecimalScore, HardSoftScore, ListSum, LoadBalance, PathCost, Projection, Sum,
                  WeightFn, hard, soft)

__all__ = [
    "AdjacentEqual", "ConsecutiveRuns", "ConstraintFactory", "Count", "EqualId", "EqualKey", "EqualKeyExpr", "EqualVarToKey", "EqualVarToRow", "Expr", "ForageParams",
    "GpuScoreDirector", "HardSoftDecimalScore", "HardSoftScore", "ListSum", "LoadBalance", "PathCost", "Projection",
    "Sum",
    "WeightFn", "hard", "soft",
]
