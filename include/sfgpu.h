/*
 * sfgpu.h — C ABI of libsfgpu: batched, read-only re-scoring of local-search candidate moves
 * on B200 (sm_100a). This is the drop-in boundary for SolverForge's incremental
 * ConstraintStream scoring path; a Rust `solverforge-gpu-sys` crate (or cgo/ctypes) binds
 * exactly these symbols (INTEGRATION.md shows the Rust side).
 *
 * Reference interfaces this replaces (paths under the reference's crates/):
 *   ScoreDirector::with_descriptor / calculate_score / before|after_variable_changed
 *       solverforge-scoring/src/director/score_director/incremental.rs:64-82,141-224
 *   ConstraintSet::{initialize_all,evaluate_all,on_insert_all,on_retract_all}
 *       solverforge-scoring/src/api/constraint_set/incremental.rs:152-212
 *   evaluate_candidate (do -> score -> undo per candidate)
 *       solverforge-solver/src/phase/localsearch/evaluation.rs:20-115
 *   BestCandidate::consider / reservoir_pick (forager tie rule)
 *       solverforge-solver/src/phase/localsearch/forager.rs:99-155
 *
 * Conventions
 *   - every function returns int32: 0 = OK, <0 = SFGPU_E_*; message via sfgpu_last_error().
 *   - nothing unwinds across the boundary; there is NO CPU fallback: without a CUDA device
 *     every compute entry point fails with SFGPU_E_CUDA.
 *   - the caller owns host pointers for the duration of the call only; the library owns all
 *     device memory it allocates. With SFGPU_DEVICE_IO the data pointers of a call are DEVICE
 *     pointers owned by the caller (already resident in HBM) and the call is asynchronous on
 *     the context's stream.
 *   - one context per solve; a context is not thread-safe but may move between threads
 *     (reference: Director: Send, not Sync — solverforge-scoring/src/director/traits.rs:27).
 *   - a context holds R independent replicas (seeded restarts) of the planning state; static
 *     facts (adjacency, matrices, columns) are shared by all replicas.
 *   - scores are (hard, soft) int64 pairs, HardSoftScore / HardSoftDecimalScore layout
 *       solverforge-core/src/score/hard_soft.rs:35-38, hard_soft_decimal.rs:45-48.
 */
#ifndef SFGPU_H
#define SFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFGPU_ABI_VERSION 1

#define SFGPU_OK 0
#define SFGPU_E_INVALID (-1)     /* bad argument / out of range */
#define SFGPU_E_UNSUPPORTED (-2) /* model shape not expressible on device (never a silent fallback) */
#define SFGPU_E_CUDA (-3)
#define SFGPU_E_NCCL (-4)
#define SFGPU_E_OOM (-5)
#define SFGPU_E_STATE (-6)       /* call order violated (e.g. score before commit) */

/* call flags */
#define SFGPU_DEVICE_IO 1u /* data pointers are device pointers; call is stream-asynchronous */
#define SFGPU_SYNC_ASYNC 2u /* sfgpu_sync_best: outputs are DEVICE pointers too and the stream is not synchronised */

typedef struct sfgpu_ctx sfgpu_ctx;

/* ---- context -------------------------------------------------------------------------- */
/* cuda_stream: a cudaStream_t to launch on (e.g. torch's current stream), or NULL to create
 * a private non-blocking stream (NULL + SFGPU_CTX_LEGACY_DEFAULT_STREAM = the legacy stream 0). */
#define SFGPU_CTX_LEGACY_DEFAULT_STREAM 1ull
/* always use the generic constraint-interpreting kernels (parity testing of the fast paths) */
#define SFGPU_CTX_GENERIC_KERNELS 2ull
int32_t sfgpu_ctx_create(int32_t device, uint64_t flags, void* cuda_stream, sfgpu_ctx** out);
int32_t sfgpu_ctx_destroy(sfgpu_ctx* ctx);
const char* sfgpu_last_error(const sfgpu_ctx* ctx);
int32_t sfgpu_abi_version(void);
int32_t sfgpu_synchronize(sfgpu_ctx* ctx);

/* ---- model ---------------------------------------------------------------------------- */
/* descriptor_index: declaration order of the #[planning_entity_collection]
 * (solverforge-macros/src/planning_model/metadata.rs:9-44); -1 = problem fact (ChangeSource::Static). */
int32_t sfgpu_model_begin(sfgpu_ctx* ctx, uint32_t n_replicas);
int32_t sfgpu_add_collection(sfgpu_ctx* ctx, const char* name, uint32_t n_rows, int32_t descriptor_index,
                             uint32_t* out_collection);
/* static int64 fact column of a collection (n_rows values), shared by all replicas */
int32_t sfgpu_add_column_i64(sfgpu_ctx* ctx, uint32_t collection, const char* name, const int64_t* values,
                             uint32_t* out_column);
/* scalar planning variable: value in [0, n_values) or -1 (None)
 * (#[planning_variable(allows_unassigned)], examples/scalar-graph-coloring/src/domain/node.rs) */
int32_t sfgpu_add_scalar_variable(sfgpu_ctx* ctx, uint32_t collection, const char* name, uint32_t n_values,
                                  int32_t allows_unassigned, uint32_t* out_variable);
/* list planning variable: each owner row holds an ordered list of element row indices
 * (#[planning_list_variable], crates/solverforge-cvrp/src/solution.rs:11-16) */
int32_t sfgpu_add_list_variable(sfgpu_ctx* ctx, uint32_t owner_collection, uint32_t element_collection,
                                const char* name, uint32_t* out_variable);
/* CSR over one collection: adjacency (Node::neighbors; PAIR_CSR_EQUAL needs column indices < n_rows) or any
 * per-row list of int values (Employee::unavailable_days) for SFGPU_X_CSR_CONTAINS / JOIN_EXPR buckets */
int32_t sfgpu_add_csr(sfgpu_ctx* ctx, const char* name, uint32_t n_rows, const uint32_t* row_ptr,
                      const uint32_t* col_idx, uint32_t* out_csr);
/* dense int64 matrix. cost_semantics 1 applies ProblemData::distance_cost at upload: a cell v is
 * kept when 0 <= v != INT64_MAX, else replaced by INT64_MAX/4
 * (crates/solverforge-cvrp/src/problem_data.rs:6-12,28-47). */
int32_t sfgpu_add_matrix_i64(sfgpu_ctx* ctx, const char* name, uint32_t rows, uint32_t cols,
                             const int64_t* values, int32_t cost_semantics, uint32_t* out_matrix);

/* weight(x) on one score level */
#define SFGPU_W_CONST 0  /* a */
#define SFGPU_W_LINEAR 1 /* a*x + b */
#define SFGPU_W_SQUARE 2 /* a*x*x + b */
#define SFGPU_W_EXCESS 3 /* a*max(0, x - b) */
#define SFGPU_W_ABSDIFF 4 /* a*|x - b| */
#define SFGPU_W_PAIRS 5   /* a*x*(x-1)/2: unordered pairs among x rows (keyed self-join of projected rows) */
typedef struct sfgpu_weight {
  int32_t fn;
  int32_t level; /* 0 = hard, 1 = soft */
  int64_t a, b;
} sfgpu_weight;

#define SFGPU_PENALTY 0
#define SFGPU_REWARD 1

/* constraint kinds — each is one ConstraintStream shape of the reference */
/* for_each(E)[.unassigned()|.filter(assigned)].penalize(w(x)); x = const | entity column | value column
 *   solverforge-scoring/src/constraint/incremental.rs:97-156
 *   p0: filter 0 = unassigned, 1 = assigned, 2 = always; aux0: column id or UINT32_MAX (x = 0);
 *   p1: 0 = column indexed by entity row, 1 = column indexed by assigned value;
 *   aux1: entity mask column id or UINT32_MAX — `.filter(|e| e.flag)`: rows whose mask is 0 never match */
#define SFGPU_K_UNI 1
/* for_each(E).join(E, |l,r| l.id < r.id && l.neighbors.contains(r.id) && l.var.is_some() && l.var == r.var)
 *   predicate cross-bi, solverforge-scoring/src/constraint/cross_bi_incremental/, stream/join_target.rs:83-110
 *   aux0: csr id; weight must be CONST */
#define SFGPU_K_PAIR_CSR_EQUAL 2
/* for_each(E).join(E, equal(key)).filter(l.id < r.id && var.is_some()); key(e) = column[e]*p0 + var[e]*p1
 *   keyed self-join, constraint/nary_incremental/bi.rs:78-206 (and the keyed form of predicate joins)
 *   aux0: key column id or UINT32_MAX (column = 0); weight must be CONST;
 *   aux1: join arity — UINT32_MAX / 0 / 2 = pairs; 3, 4, 5 = the keyed tri / quad / penta self-joins
 *   `.join(equal(key)).join(equal(key))…` with index-ordered tuples a < b < c … and a constant weight per tuple
 *   (constraint/nary_incremental/higher_arity/shared.rs:28-417): a bucket of n rows holds C(n, arity) tuples */
#define SFGPU_K_PAIR_KEY_EQUAL 3
/* for_each(A).if_exists|if_not_exists(for_each(owners).flattened(list), equal(a.row, element))
 *   constraint/exists.rs:126-417; p0: 0 = exists, 1 = not exists; weight must be CONST */
#define SFGPU_K_EXISTS_FLAT 4
/* for_each(E).join(values, equal(var, Some(v.row))).group_by(v.row, count()|sum(column))
 *   [.complement(values, default p1)].penalize(w(result))
 *   constraint/grouped/, cross_grouped/, cross_complemented_grouped/
 *   aux0: summed entity column id or UINT32_MAX (count); p0: 1 = complemented; p1: default result;
 *   aux1: per-value column that replaces the weight offset b for that key, or UINT32_MAX — key-dependent
 *   weights |key, result| (e.g. max(0, result - capacity[key])). A single-emit `.project(..)` row keyed by
 *   the planning variable (stream/projected_stream/source.rs:13-24) lowers to this kind as well. */
#define SFGPU_K_GROUP 5
/* for_each(owners).penalize(w(sum of matrix legs depot -> list... -> depot)); empty list = 0
 *   uni constraint whose weight walks the route; aux0: matrix id; p0: depot row */
#define SFGPU_K_LIST_PATH_COST 6
/* for_each(owners).penalize(w(sum of column[element])) (C3-ii: EXCESS weight with b = capacity)
 *   aux0: element column id */
#define SFGPU_K_LIST_SUM 7
/* for_each(E).group_by((), load_balance(var, metric column|1)).penalize(w(unfairness))
 *   stream/collector/load_balance.rs:104-240; aux0: metric column id or UINT32_MAX (metric = 1) */
#define SFGPU_K_LOAD_BALANCE 8
/* for_each(E).project(P).group_by(key, count()|sum(amount)).penalize(w(result)) with P::MAX_EMITS <= 8:
 *   projected scoring rows, stream/projected_stream/source.rs:13-147 (Projection, RowCoordinate),
 *   constraint/projected/grouped/{state.rs,terminal.rs}. An ASSIGNED entity e emits one row per entry j
 *   of its csr row: key = var[e] * p0 + csr.col[j] (p0 = number of key offsets per value, e.g. days),
 *   amount = aux1 column[j] (one row per csr entry) or 1 (count). Groups with no rows score nothing
 *   (grouped/state.rs:320-332); two rows of one entity may share a group.
 *   aux0: csr id (rows = entities); aux1: amount column id or UINT32_MAX. The keyed self-join of projected
 *   rows, .project(P).join(equal(key)).penalize(CONST) (constraint/projected/bi.rs: every unordered pair of
 *   rows with equal keys, rows of one entity included), is this kind with count() and SFGPU_W_PAIRS.
 *   A projected UNI terminal
 *   (.project(P).penalize(w(row))) lowers on the host to SFGPU_K_UNI over the per-entity sum of row weights
 *   (constraint/projected/uni.rs:61-263). */
#define SFGPU_K_PROJECT_GROUP 9

/* for_each(E).filter(assigned).group_by(var, consecutive_runs(point column)).penalize(sum over runs of
 *   w(run.point_count)): stream/collector/runs.rs:14-229 (unique points as maximal runs of consecutive
 *   integers; duplicates do not lengthen a run) — "Long work streaks" of examples/minimal-shift-scheduling
 *   (schedule.rs:43-58) is EXCESS with b = 2. aux0: entity column with the point (0 <= point < p0);
 *   p0: number of points.
 *   aux1 selects a view of the indexed_presence collector over the same per-group point counts
 *   (stream/collector/indexed_presence.rs:6-147; UINT32_MAX / 0 = consecutive_runs as above):
 *   1 = sum over complement_runs(lo..hi) of w(run.point_count); 2 = any_in(lo..hi) ? w(1) : 0;
 *   3 = w(count()) of the distinct points; p1 = lo | hi << 32 with 0 <= lo <= hi <= p0. Groups without items
 *   do not exist and score nothing. */
#define SFGPU_K_RUNS 10
/* for_each(A).join(for_each(B), equal(a.var, b.key)).filter(pair filter).penalize(pair weight) — the general
 * cross-collection join of the reference (constraint/cross_bi_incremental/state.rs:260-460, stream/join_target.rs:
 * 57-110, stream/filter/adapters.rs:63-93) for the planning-model form: the A side is the entity collection, its join
 * key is the scalar planning variable (None joins nothing) and the B side is a fact collection. Pair filter and
 * pair weight are column expressions (sfgpu_add_expr) over (a, b, index of a, index of b) — closures cannot run
 * on a GPU. No retained state: a ChangeMove of a re-evaluates the pairs of its old and its new key.
 *   collection = A; p1 = B collection;
 *   p0 = -1: b.key is the row index of B (a.var references a B row), else a CSR id mapping every key value to the
 *        B rows that carry it (several B rows per key);
 *   aux0 = pair filter expression id or UINT32_MAX (every joined pair matches);
 *   aux1 = pair weight expression id or UINT32_MAX: the pair scores weight(x) with x = the expression's value
 *        (x = 0 without one), e.g. {SFGPU_W_LINEAR, level, 1, 0} scores x itself, SFGPU_W_CONST a constant. */
#define SFGPU_K_JOIN_EXPR 11
/* Keyed self-join of entity rows with a pair filter and a pair weight, undirected or directed — the reference's
 * `for_each(E).join(equal(key)).filter(|l, r| ..).penalize(|l, r| ..)` (constraint/nary_incremental/bi.rs:78-206 with an
 * arbitrary filter / weight) and its projected forms `.project(P).join(equal(key))..` (constraint/projected/bi.rs) and
 * `.project(P).join(equal_bi(left_key, right_key))..` (constraint/projected/directed_bi.rs) for single-emit projections:
 * every ASSIGNED entity is one row; its key(s) are column expressions of the row (SFGPU_X_A_COL, SFGPU_X_A_VAL,
 * SFGPU_X_A_IDX, constants, arithmetic); a key outside [0, p1) means "no key" (Option::None).
 *   undirected (right key expression UINT32_MAX): every pair of rows with equal keys once, left = the lower entity
 *     index (projected/bi.rs:116-152, bi.rs:107-126);
 *   directed: every ordered pair (l, r), l != r, with left_key(l) == right_key(r) (directed_bi_incremental.rs:26-45).
 * Retained per replica: the rows of every key as intrusive doubly linked lists (the reference's rows_by_key maps).
 *   p0 = left key expression | (uint64) right key expression << 32; p1 = number of keys (dense);
 *   aux0 = pair filter expression or UINT32_MAX; aux1 = pair weight expression or UINT32_MAX (x = 0); in both,
 *   A_* reads the left row, B_* the right row. */
#define SFGPU_K_PAIR_KEY_EXPR 12

/* Column expressions: a postfix program over an int64 stack (depth <= 8); booleans are 0 / 1. */
#define SFGPU_X_CONST 1        /* push imm */
#define SFGPU_X_A_COL 2        /* push column[arg][a] (a column of collection A) */
#define SFGPU_X_B_COL 3        /* push column[arg][b] (a column of collection B) */
#define SFGPU_X_A_IDX 4        /* push index of a */
#define SFGPU_X_B_IDX 5        /* push index of b */
#define SFGPU_X_VALUE 6        /* push the join key (value of a's planning variable) */
#define SFGPU_X_A_VAL 7        /* push the planning value of row a (the left row of a self-join) */
#define SFGPU_X_B_VAL 8        /* push the planning value of row b (the right row of a self-join; -1 for a fact row) */
#define SFGPU_X_ADD 10
#define SFGPU_X_SUB 11
#define SFGPU_X_MUL 12
#define SFGPU_X_NEG 13
#define SFGPU_X_ABS 14
#define SFGPU_X_MIN 15
#define SFGPU_X_MAX 16
#define SFGPU_X_MOD 17         /* Euclidean-free: x % y with the sign of x, 0 when y == 0 */
#define SFGPU_X_EQ 20
#define SFGPU_X_NE 21
#define SFGPU_X_LT 22
#define SFGPU_X_LE 23
#define SFGPU_X_GT 24
#define SFGPU_X_GE 25
#define SFGPU_X_AND 30
#define SFGPU_X_OR 31
#define SFGPU_X_NOT 32
#define SFGPU_X_CSR_CONTAINS 40 /* pops x, then row: pushes csr[arg].row(row).contains(x) (e.g. unavailable_days) */
#define SFGPU_X_SELECT 41       /* pops else, then, cond: pushes cond ? then : else */
typedef struct sfgpu_expr_op {
  int32_t op;
  uint32_t arg;
  int64_t imm;
} sfgpu_expr_op;
/* registers a program (validated: stack discipline, one result); ids are per context */
int32_t sfgpu_add_expr(sfgpu_ctx* ctx, const sfgpu_expr_op* ops, uint32_t n_ops, uint32_t* out_expr);

typedef struct sfgpu_constraint_desc {
  int32_t kind;
  int32_t impact;      /* SFGPU_PENALTY / SFGPU_REWARD (constraint/incremental.rs:70-76) */
  sfgpu_weight weight;
  uint32_t collection; /* source collection (A side) */
  uint32_t variable;   /* scalar or list variable the constraint reads */
  uint32_t aux0;
  uint32_t aux1;
  int64_t p0, p1;
  const char* name;
} sfgpu_constraint_desc;

int32_t sfgpu_add_constraint(sfgpu_ctx* ctx, const sfgpu_constraint_desc* desc, uint32_t* out_constraint);

/* planning state upload. per_replica 0: one copy broadcast to every replica; 1: [R] copies.
 * scalar: int32 values[n_rows]; list: offsets[n_owners+1] + elems[offsets[n_owners]] per copy
 * (per_replica 1: offsets is [R][n_owners+1], elems is the concatenation of the R copies). */
int32_t sfgpu_set_scalar_state(sfgpu_ctx* ctx, uint32_t variable, const int32_t* values, int32_t per_replica);
int32_t sfgpu_set_list_state(sfgpu_ctx* ctx, uint32_t variable, const uint32_t* offsets, const uint32_t* elems,
                             int32_t per_replica);
/* freezes the model, lays out HBM, runs initialize_all on every replica
 * (ScoreDirector::calculate_score first call, incremental.rs:141-149); out_scores [R][2] or NULL */
int32_t sfgpu_model_commit(sfgpu_ctx* ctx, int64_t* out_scores);

/* ---- candidate batches ---------------------------------------------------------------- */
/* A batch holds the candidates of all R replicas back to back; replica r owns rows
 * [cand_offsets[r], cand_offsets[r+1]); n_candidates == cand_offsets[R] (passed separately so the
 * device-pointer path never reads an offset back). Rows are packed so one candidate is one aligned vector
 * load. Outputs: out_scores[n][2] = score of the working solution AFTER the move (what
 * evaluate_candidate returns); out_doable[n] = Move::is_doable. Not-doable rows get score (0,0).
 *
 * scalar rows: {uint32 entity, int32 to_value(-1 = None)}  == ScalarEdit
 *   (solverforge-solver/src/planning/scalar/candidate.rs:6-12); ChangeMove semantics change.rs:125-175 */
int32_t sfgpu_score_change(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets, const uint32_t* rows,
                           int64_t* out_scores, uint8_t* out_doable);
/* swap rows: {uint32 left_entity, uint32 right_entity} (heuristic/move/swap.rs:140-215) */
int32_t sfgpu_score_swap(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets, const uint32_t* rows,
                         int64_t* out_scores, uint8_t* out_doable);
/* compound: candidate i owns ScalarEdit rows [edit_offsets[i], edit_offsets[i+1])
 * (CompoundScalarMove, heuristic/move/compound_scalar.rs:245-319); at most SFGPU_MAX_EDITS per candidate */
#define SFGPU_MAX_EDITS 8
int32_t sfgpu_score_compound(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                             const uint64_t* edit_offsets, const uint32_t* edit_rows, int64_t* out_scores,
                             uint8_t* out_doable);
/* list change rows: {src_entity, src_position, dst_entity, dst_position} uint32 x4
 * (ListChangeMove, heuristic/move/list_kernel/change.rs:15-153) */
int32_t sfgpu_score_list_change(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                                const uint32_t* rows, int64_t* out_scores, uint8_t* out_doable);
/* list swap rows: {first_entity, first_position, second_entity, second_position}
 * (ListSwapMove, heuristic/move/list_kernel/swap.rs:16-110) */
int32_t sfgpu_score_list_swap(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets, const uint32_t* rows,
                              int64_t* out_scores, uint8_t* out_doable);

/* rows[n][4] = {entity, start, end, 0}: ListReverseMove reverses [start, end) of one list (2-opt segment
 * reversal, heuristic/move/list_kernel/reverse.rs:21-58; doable iff end > start + 1 && end <= len) */
int32_t sfgpu_score_list_reverse(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                                 const uint32_t* rows, int64_t* out_scores, uint8_t* out_doable);

/* rows[n][4] = {src_entity, start | size << 24, dst_entity, dst_position}: SublistChangeMove relocates the
 * contiguous segment [start, start + size) (heuristic/move/list_kernel/sublist_change.rs:17-125). For an
 * intra-list move dst_position is a position of the list AFTER the removal. Doable iff size >= 1,
 * start + size <= len(src), dst_position <= (intra ? len - size : len(dst)) and not (intra && dst_position ==
 * start). Positions are limited to 2^24 and segment sizes to 255 by the packing (reference default: 1..=3,
 * solverforge-config/src/move_selector.rs:713-715). SFGPU_SEG(start, size) packs the second word. */
#define SFGPU_SEG(start, size) (((uint32_t)(start) & 0xFFFFFFu) | ((uint32_t)(size) << 24))
int32_t sfgpu_score_sublist_change(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                                   const uint32_t* rows, int64_t* out_scores, uint8_t* out_doable);

/* rows[n][4] = {first_entity, start1 | size1 << 24, second_entity, start2 | size2 << 24}: SublistSwapMove
 * exchanges two contiguous segments, sizes may differ (heuristic/move/list_kernel/sublist_swap.rs:17-170).
 * Doable iff both sizes >= 1, both segments inside their lists and, inside one list, the segments do not overlap. */
int32_t sfgpu_score_sublist_swap(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                                 const uint32_t* rows, int64_t* out_scores, uint8_t* out_doable);

/* ---- winner selection on device ------------------------------------------------------- */
/* KOptMove on one list (heuristic/move/k_opt.rs:11-110): cuts c_0 < .. < c_{k-1} (2 <= k <= 5) split the list into
 * k + 1 segments; reconnection pattern p re-orders the middle segments and reverses some of them — p indexes
 * enumerate_reconnections(k) (k_opt_reconnection.rs:218-262: every order of the middle segments, lexicographic,
 * times every reversal mask, the identity dropped; 1 / 7 / 47 / 383 patterns for k = 2..5, the 3-opt table of
 * k_opt/selector.rs:111-115 included). rows[n][4] = {entity | k << 28, c0 | c1 << 16, c2 | c3 << 16,
 * c4 | p << 16} (positions < 65536). */
#define SFGPU_KOPT_ROW0(entity, k) (((uint32_t)(entity) & 0x0FFFFFFFu) | ((uint32_t)(k) << 28))
int32_t sfgpu_score_k_opt(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                          const uint32_t* rows, int64_t* out_scores, uint8_t* out_doable);

/* Per replica: replay of acceptor + forager over the scored rows in pull order
 * (phase/candidates.rs:66-282 with BestCandidate::consider, forager.rs:99-155).
 * acceptor: 0 = accept every doable candidate, 1 = HillClimbing (score > last_step_score,
 *           acceptor/hill_climbing.rs:33-42), 2 = LateAcceptance form (score >= last_step_score ||
 *           score >= threshold, late_acceptance.rs:89-101), 3 = GreatDeluge form (score >
 *           last_step_score || score >= threshold, great_deluge.rs:53-68). The stateful acceptors reduce
 *           to these per step: StepCountingHillClimbing = 0 or 1 (step_counting.rs:56-67);
 *           DiversifiedLateAcceptance = 2 with threshold min(late score, best - |best| * tolerance)
 *           (diversified_late_acceptance.rs:72-101). Tabu and SimulatedAnnealing need per-move metadata /
 *           a random stream: they replay on the host over the materialised scores.
 * forager : accepted_limit 0 = BestScore (never quits early); N > 0 = AcceptedCount(N)
 * tie_mode: 0 = ScoreTieBreak::First, 1 = reservoir (random_ties)
 * ref_scores[R][2][2] = {last_step_score, threshold (late score / water level)}; step_seeds[R].
 * out_index[R] = winning pull index inside the replica's range, or UINT32_MAX when none accepted;
 * out_best[R][2] its score; out_evaluated[R] = moves_evaluated (pulls up to the quit point). */
typedef struct sfgpu_forage_params {
  int32_t acceptor;
  int32_t tie_mode;
  uint32_t accepted_limit;
  uint32_t reserved;
} sfgpu_forage_params;
int32_t sfgpu_argbest(sfgpu_ctx* ctx, uint32_t flags, const sfgpu_forage_params* params,
                      const uint64_t* cand_offsets, const int64_t* scores, const uint8_t* doable,
                      const uint64_t* step_seeds, const int64_t* ref_scores, uint32_t* out_index,
                      int64_t* out_best, uint32_t* out_evaluated);

/* sfgpu_argbest with the per-candidate improvement gates of evaluate_candidate
 * (phase/localsearch/evaluation.rs:76-111): gates[i] bit 0 = Move::requires_hard_improvement (the candidate
 * only reaches the acceptor when hard_score_delta(last_step_score, score) is Improving), bit 1 =
 * Move::requires_score_improvement (only when score > last_step_score). Gated-out candidates still count as
 * evaluated. CompoundScalarMove carries the first flag (heuristic/move/compound_scalar.rs:340). gates may be
 * NULL; it lives where scores / doable live (host, or device with SFGPU_DEVICE_IO). */
int32_t sfgpu_argbest_gated(sfgpu_ctx* ctx, uint32_t flags, const sfgpu_forage_params* params,
                            const uint64_t* cand_offsets, const int64_t* scores, const uint8_t* doable,
                            const uint8_t* gates, const uint64_t* step_seeds, const int64_t* ref_scores,
                            uint32_t* out_index, int64_t* out_best, uint32_t* out_evaluated);

/* Fused step for list-change batches (DEVICE pointers, stream-asynchronous): scores every candidate and
 * replays acceptor + forager in the same pass — evaluate_candidates (phase/candidates.rs:47-285) for a
 * whole neighbourhood. out_scores / out_doable may both be NULL: then per-candidate scores are never
 * written to HBM (only the winner leaves the kernel). Same outputs as sfgpu_argbest. */
int32_t sfgpu_step_list_change(sfgpu_ctx* ctx, uint64_t n_candidates, const uint64_t* cand_offsets,
                               const uint32_t* rows, const sfgpu_forage_params* params, const uint64_t* step_seeds,
                               const int64_t* ref_scores, int64_t* out_scores, uint8_t* out_doable,
                               uint32_t* out_index, int64_t* out_best, uint32_t* out_evaluated);

/* The same fused step over resident ChangeMove rows {entity, to_value} of a scalar model (DEVICE pointers,
 * stream-asynchronous): models whose scalar constraint tuple has a monomorphised program (sfgpu_scalar_program
 * >= 0) score and reduce in one pass — the scores are never re-read; others, and AcceptedCount(N), materialise
 * the scores and run the ordered replay. out_scores / out_doable may both be NULL. Same outputs as sfgpu_argbest. */
int32_t sfgpu_step_change_rows(sfgpu_ctx* ctx, uint64_t n_candidates, const uint64_t* cand_offsets,
                               const uint32_t* rows, const sfgpu_forage_params* params, const uint64_t* step_seeds,
                               const int64_t* ref_scores, int64_t* out_scores, uint8_t* out_doable,
                               uint32_t* out_index, int64_t* out_best, uint32_t* out_evaluated);

/* Whole local-search step on device: generates the nearby list-change neighbourhood of every replica
 * (NearbyListChangeMoveSelector, heuristic/selector/list_kernel/nearby_change.rs:102-232 with the matrix
 * distance meter crates/solverforge-cvrp/src/meters.rs:10-28; canonical SelectionOrder::Original), scores
 * it, replays acceptor + forager and optionally commits the winners — evaluate_candidates + pick + apply
 * (phase/step.rs:30-225) without the neighbourhood ever crossing PCIe.
 *   flags & SFGPU_DEVICE_IO selects where the small per-replica arrays live (step_seeds, ref_scores,
 *   out_index, out_best, out_evaluated, out_winner_rows); without it they are HOST pointers.
 *   out_cand_offsets / out_rows / out_scores / out_doable are optional DEVICE buffers that receive the
 *   generated batch: replica r owns rows [r*S, (r+1)*S), S = list_capacity * max_nearby, source position f
 *   owns rows f*max_nearby .. +max_nearby in pull order; rows that do not exist are not-doable sentinels.
 *   out_index is the reference pull index (CandidateId): source_position * candidates_per_source + rank.
 *   out_winner_rows[R][4] = the winning ListChangeMove of each replica (sentinel 0xFFFFFFFF when none).
 * Requires the fast list program (SFGPU_E_UNSUPPORTED otherwise) and max_nearby <= 32.
 * Retained neighbourhood: when nothing is materialised (out_rows NULL) the library keeps every source's score
 * deltas between calls; after a winner committed through apply_winners (or sfgpu_apply_list_change /
 * sfgpu_apply_winners of one ListChange move) the next call regenerates only the sources that move can have changed
 * and re-scores the candidates into the two touched routes — results are bit-identical to a full regeneration;
 * any other writer of the planning state invalidates the kept deltas (environment SFGPU_NO_NBCACHE=1, read at
 * commit, turns it off).
 * With host pointers and apply_winners the call returns once the winners have reached the host; the commit kernel is
 * still ordered on the context's stream ahead of any later call (sfgpu_synchronize waits for it). */
int32_t sfgpu_step_nearby_list_change(sfgpu_ctx* ctx, uint32_t flags, uint32_t max_nearby,
                                      const sfgpu_forage_params* params, const uint64_t* step_seeds,
                                      const int64_t* ref_scores, uint64_t* out_cand_offsets, uint32_t* out_rows,
                                      int64_t* out_scores, uint8_t* out_doable, uint32_t* out_index,
                                      int64_t* out_best, uint32_t* out_evaluated, uint32_t* out_winner_rows,
                                      int32_t apply_winners);

/* The same whole step over the nearby list-SWAP neighbourhood (NearbyListSwapMoveSelector,
 * heuristic/selector/nearby_list_swap.rs:165-213 + list_kernel/nearby_swap.rs:99-262, canonical order): for
 * every source position the destinations are the later positions of its own list and every position of the
 * later entities, the max_nearby nearest by the matrix meter (stable ties); sources without destinations are
 * skipped, so out_index (CandidateId) counts only existing candidates. Scored as ListSwapMove
 * (heuristic/move/list_kernel/swap.rs). Pointer conventions and the materialised layout (fixed stride of
 * max_nearby rows per source, sentinels elsewhere) as sfgpu_step_nearby_list_change; out_winner_rows[R][4] =
 * {first_entity, first_position, second_entity, second_position}. */
int32_t sfgpu_step_nearby_list_swap(sfgpu_ctx* ctx, uint32_t flags, uint32_t max_nearby,
                                    const sfgpu_forage_params* params, const uint64_t* step_seeds,
                                    const int64_t* ref_scores, uint64_t* out_cand_offsets, uint32_t* out_rows,
                                    int64_t* out_scores, uint8_t* out_doable, uint32_t* out_index, int64_t* out_best,
                                    uint32_t* out_evaluated, uint32_t* out_winner_rows, int32_t apply_winners);

/* The same whole step for scalar models: generates the full ChangeMove neighbourhood of every replica in
 * the canonical order of ChangeMoveSelector (heuristic/selector/move_selector/change.rs:66-104,246-307:
 * entities in order, per entity every value then the to-None move when it is assigned), scores it with
 * every scalar constraint kind, replays acceptor + forager, optionally commits the winners. Pointer
 * conventions as sfgpu_step_nearby_list_change; the materialised batch gives replica r the rows
 * [r*S, (r+1)*S), S = n_entities * (n_values + 1), in pull order, padded with not-doable sentinels;
 * out_winner_rows[R][2] = the winning ScalarEdit {entity, to_value} (entity 0xFFFFFFFF when none). */
int32_t sfgpu_step_change(sfgpu_ctx* ctx, uint32_t flags, const sfgpu_forage_params* params,
                          const uint64_t* step_seeds, const int64_t* ref_scores, uint64_t* out_cand_offsets,
                          uint32_t* out_rows, int64_t* out_scores, uint8_t* out_doable, uint32_t* out_index,
                          int64_t* out_best, uint32_t* out_evaluated, uint32_t* out_winner_rows,
                          int32_t apply_winners);

/* The same whole step over the SublistChange neighbourhood (SublistChangeMoveSelector, SelectionOrder::Original,
 * heuristic/selector/sublist_change.rs:166-205 + list_kernel/sublist_change.rs:103-268): per source entity every
 * segment start and size min_size..=max_size (reference defaults 1..=3), intra-list destinations first, then every
 * position of every other entity. The neighbourhood (CVRP-1000: ~3.2 M candidates per replica) is never
 * materialised: candidates are decoded from their pull index on device, scored, and only the winner leaves the
 * SM. out_index = CandidateId (pull index) or UINT32_MAX; out_winner_rows[R][4] = the packed SublistChange row
 * {src_entity, start | size << 24, dst_entity, dst_position}; apply_winners commits it. With SFGPU_DEVICE_IO
 * every pointer is a device pointer and the call is asynchronous. */
int32_t sfgpu_step_sublist_change(sfgpu_ctx* ctx, uint32_t flags, uint32_t min_size, uint32_t max_size,
                                  const sfgpu_forage_params* params, const uint64_t* step_seeds,
                                  const int64_t* ref_scores, uint32_t* out_index, int64_t* out_best,
                                  uint32_t* out_evaluated, uint32_t* out_winner_rows, int32_t apply_winners);

/* The same for the ListReverse (2-opt segment reversal) neighbourhood (ListReverseMoveSelector, SelectionOrder::Original,
 * heuristic/selector/list_reverse.rs:139-172 + list_kernel/reverse.rs:66-108): per entity every start and every end
 * in start + 2 ..= len. out_winner_rows[R][4] = {entity, start, end, 0}. */
int32_t sfgpu_step_list_reverse(sfgpu_ctx* ctx, uint32_t flags, const sfgpu_forage_params* params,
                                const uint64_t* step_seeds, const int64_t* ref_scores, uint32_t* out_index,
                                int64_t* out_best, uint32_t* out_evaluated, uint32_t* out_winner_rows,
                                int32_t apply_winners);

/* The same for the SublistSwap neighbourhood (SublistSwapMoveSelector, SelectionOrder::Original,
 * list_kernel/sublist_swap.rs:28-318): first segments in entity / start / size order, each paired with the later
 * non-overlapping segments of its own list and every segment of the later entities (CVRP-1000: ~4.4 M pairs per
 * replica). out_winner_rows[R][4] = {first_entity, start1 | size1 << 24, second_entity, start2 | size2 << 24}. */
int32_t sfgpu_step_sublist_swap(sfgpu_ctx* ctx, uint32_t flags, uint32_t min_size, uint32_t max_size,
                                const sfgpu_forage_params* params, const uint64_t* step_seeds,
                                const int64_t* ref_scores, uint32_t* out_index, int64_t* out_best,
                                uint32_t* out_evaluated, uint32_t* out_winner_rows, int32_t apply_winners);

/* ---- union of neighbourhoods in seeded pull order (the reference's default list local search) -------------
 * One step over a VecUnionSelector of list move families (heuristic/selector/decorator/vec_union.rs:204-366)
 * exactly as the reference pulls it: every leaf walks its cursor in SelectionOrder `selection_order`
 * (MoveStreamContext::selection_index, move_selector/iter.rs:109-125 — Random is with replacement, Shuffled a
 * strided permutation), the children are interleaved by `union_order` (the default list policy,
 * runtime/compiler/default_local_search/policy/list.rs:24-33, is Random leaves + StratifiedRandom), and the
 * acceptor + forager replay stops where the reference stops. Only a prefix of the union stream is generated:
 * each child emits its first `window` candidates, the step is complete when the forager quit inside that
 * window (AcceptedCount) or every cursor ended; otherwise the window grows by x8 up to `max_window` per child.
 * A step still incomplete at max_window picks the best of what it saw and sets bit 0 of out_flags[r]
 * (BestScore over a multi-million sublist neighbourhood is not what this call is for — use the per-family
 * whole-neighbourhood steps above).
 *   families : canonical cursors restated per family — NearbyListChange / NearbyListSwap (p0 = max_nearby <= 32),
 *              SublistChange / SublistSwap (p0 = min_size, p1 = max_size), ListReverse, KOpt (p0 = k, p1 =
 *              min_segment_len; winner row = the packed k-opt row of sfgpu_score_k_opt).
 *   step_indices[R] (may be NULL = 0) and step_seeds[R] are the MoveStreamContext of each replica's step.
 *   out_index = CandidateId of the union cursor (pull index, vec_union.rs:447-455) or UINT32_MAX;
 *   out_winner_rows[R][8] = {family, child, row[4], child-local pull index, 0}.
 *   Scalar models take the Change / Swap families — the reference's default for plain scalar models is
 *   union[ChangeMoveSelector(Random), SwapMoveSelector(Random)], StratifiedRandom, SimulatedAnnealing +
 *   AcceptedCount(1) (runtime/compiler/default_local_search/policy.rs:48-81, policy/scalar.rs:64-108).
 * Pointer conventions as sfgpu_step_nearby_list_change (HOST arrays unless SFGPU_DEVICE_IO). Needs the fast
 * list program when a nearby family is present. */
#define SFGPU_FAM_NEARBY_LIST_CHANGE 0
#define SFGPU_FAM_NEARBY_LIST_SWAP 1
#define SFGPU_FAM_SUBLIST_CHANGE 2
#define SFGPU_FAM_SUBLIST_SWAP 3
#define SFGPU_FAM_LIST_REVERSE 4
#define SFGPU_FAM_CHANGE 6 /* ChangeMoveSelector of a scalar model (move_selector/change.rs:246-307); rows {entity, to_value} */
#define SFGPU_FAM_SWAP 7   /* SwapMoveSelector (move_selector/swap.rs:196-233); rows {left, right} */
#define SFGPU_FAM_K_OPT 5 /* KOptMoveSelector, p0 = k (2..5), p1 = min_segment_len; list_kernel/k_opt/full.rs:34-98 */
#define SFGPU_ORDER_ORIGINAL 0
#define SFGPU_ORDER_RANDOM 1
#define SFGPU_ORDER_SHUFFLED 2
#define SFGPU_UNION_SEQUENTIAL 0
#define SFGPU_UNION_ROUND_ROBIN 1
#define SFGPU_UNION_ROTATING_ROUND_ROBIN 2
#define SFGPU_UNION_RANDOM 3
#define SFGPU_UNION_STRATIFIED_RANDOM 4
#define SFGPU_UNION_MAX_CHILDREN 8
typedef struct sfgpu_union_child {
  int32_t family;
  uint32_t p0, p1;
  uint32_t reserved;
  uint64_t weight; /* UnionWeighting::Equal = 1 */
} sfgpu_union_child;
typedef struct sfgpu_union_desc {
  uint32_t n_children;
  int32_t union_order;     /* SFGPU_UNION_* */
  int32_t selection_order; /* SFGPU_ORDER_* of every leaf */
  uint32_t window;         /* first window per child (0 = 64) */
  uint32_t max_window;     /* largest window per child (0 = 4096) */
  uint32_t reserved;
  sfgpu_union_child children[SFGPU_UNION_MAX_CHILDREN];
} sfgpu_union_desc;
int32_t sfgpu_step_union(sfgpu_ctx* ctx, uint32_t flags, const sfgpu_union_desc* desc,
                         const sfgpu_forage_params* params, const uint64_t* step_seeds, const uint64_t* step_indices,
                         const int64_t* ref_scores, uint32_t* out_index, int64_t* out_best, uint32_t* out_evaluated,
                         uint32_t* out_winner_rows, uint32_t* out_flags, int32_t apply_winners);
/* The device-resident loop (below) over this union step: step t of replica r uses step_index = t and
 * step_seed = splitmix64(seed_base ^ r * 0x9E3779B97F4A7C15 ^ t); three window passes per step: the
 * replica's adaptive window (what its previous step needed per child plus half, starting at `window`), four
 * times that, then max_window — each only for the replicas whose forager had not quit in the pass before. out_window_overflows (may be NULL) counts the steps per replica that hit max_window;
 * out_pulls_scored (may be NULL) the union pulls scored per replica over all passes (speculation included —
 * compare with out_moves_evaluated, the pulls the reference's loop would have evaluated). */
struct sfgpu_solve_params;
int32_t sfgpu_solve_union(sfgpu_ctx* ctx, const sfgpu_union_desc* desc, const struct sfgpu_solve_params* params,
                          int64_t* out_best_scores, uint64_t* out_moves_evaluated, uint64_t* out_accepted_steps,
                          uint64_t* out_window_overflows, uint64_t* out_pulls_scored);

/* Device-resident local-search loop: n_steps whole steps (seed, neighbourhood, scoring, acceptor, forager,
 * commit, acceptor.step_ended, best-solution tracking) without a host round trip, captured in a CUDA
 * graph — solve_local_search_with_resources (phase/localsearch/phase.rs:237-320) for every replica.
 * acceptor: 1 HillClimbing, 2 LateAcceptance(late_size), 3 GreatDeluge(acceptor_real = rain_speed),
 * 4 StepCountingHillClimbing(step_count_limit), 5 DiversifiedLateAcceptance(late_size, acceptor_real =
 * tolerance), 6 SimulatedAnnealing (every loop; acceptor_real = decay rate, 0 = the
 * default 0.999985; late_size = calibration sample size, 0 = 128; step_count_limit bit 0 =
 * HardRegressionPolicy::NeverAcceptHardRegression), 7 TabuSearch (sfgpu_solve_change and sfgpu_solve_nearby_list_change; late_size packs the four
 * tenures as bytes: entity | value << 8 | move << 16 | undo_move << 24, each <= 64, 0 = dimension off, at least one
 * set; step_count_limit bit 0 = aspiration enabled: a tabu move whose score beats the best score is accepted;
 * tabu_search.rs:103-237 with the ChangeMove / ListChangeMove signatures of heuristic/move/change.rs:189-220 and
 * list_kernel/change.rs:155-204) —
 * acceptor/{hill_climbing,late_acceptance,great_deluge,step_counting,diversified_late_acceptance,simulated_annealing,
 * tabu_search}.rs, state kept per replica on the device.
 * SimulatedAnnealing replays is_accepted in pull order over the step's scores (calibration from the first
 * worsening candidates the phase evaluates, Boltzmann test on the first differing level, geometric decay once
 * calibrated). The reference draws its uniforms from rand::SmallRng (third party, unpinned); here draw j of a
 * step is (splitmix64(step_seed ^ 0x5A17EA11EA1DF00D ^ j * 0x9E3779B97F4A7C15) >> 11) * 2^-53, consumed exactly
 * where the reference draws (worsening candidates whose level temperature exceeds 1e-9). The reference draws step seeds from
 * rand::StdRng (unpinned third-party stream); here step t of replica r uses
 * splitmix64(seed_base ^ r * 0x9E3779B97F4A7C15 ^ t), so a trajectory is reproducible and each of its
 * steps can be checked against the oracle, but it is not the reference's trajectory.
 * Host outputs (may be NULL): best score per replica, moves_evaluated and committed steps per replica.
 * restore_best != 0 makes the best solution of each replica its working solution at the end
 * (read it with sfgpu_get_list_state). */
typedef struct sfgpu_solve_params {
  uint32_t max_nearby;
  uint32_t n_steps;
  int32_t acceptor;
  uint32_t late_size;
  int32_t tie_mode;
  uint32_t accepted_limit;
  uint64_t seed_base;
  int32_t restore_best;
  int32_t reserved;          /* bit 0 (sfgpu_solve_nearby_list_change, AcceptedCount): windowed speculation — only the prefix of
                                the cursor the forager consumes is generated (one-child union, same winners) */
  double acceptor_real;      /* GreatDeluge rain_speed / DiversifiedLateAcceptance tolerance */
  uint64_t step_count_limit; /* StepCountingHillClimbing */
} sfgpu_solve_params;
int32_t sfgpu_solve_nearby_list_change(sfgpu_ctx* ctx, const sfgpu_solve_params* params, int64_t* out_best_scores,
                                       uint64_t* out_moves_evaluated, uint64_t* out_accepted_steps);
/* the same loop over the full ChangeMove neighbourhood of a scalar model (max_nearby is ignored) */
int32_t sfgpu_solve_change(sfgpu_ctx* ctx, const sfgpu_solve_params* params, int64_t* out_best_scores,
                           uint64_t* out_moves_evaluated, uint64_t* out_accepted_steps);

/* ---- committing the winner ------------------------------------------------------------ */
/* One row per replica (same packing as the score calls); mask[r] == 0 skips replica r
 * (mask may be NULL). Updates the replica's planning state, its retained aggregates and its
 * committed score (Move::do_move on the committed director, phase/step.rs:140-142). */
int32_t sfgpu_apply_change(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask);
int32_t sfgpu_apply_swap(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask);
int32_t sfgpu_apply_list_change(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask);
int32_t sfgpu_apply_list_swap(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask);
int32_t sfgpu_apply_list_reverse(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask);
int32_t sfgpu_apply_sublist_change(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask);
int32_t sfgpu_apply_sublist_swap(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask);
int32_t sfgpu_apply_k_opt(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask);
/* apply the winner found by sfgpu_argbest straight from the batch, no host round trip:
 * row = batch_rows[cand_offsets[r] + index[r]]; replicas with index == UINT32_MAX are skipped.
 * move_kind: 0 change, 1 swap, 2 list change, 3 list swap, 4 list reverse, 5 sublist change, 6 sublist swap, 7 k-opt. All pointers are device pointers. */
int32_t sfgpu_apply_winners(sfgpu_ctx* ctx, int32_t move_kind, const uint64_t* cand_offsets,
                            const uint32_t* batch_rows, const uint32_t* index);

/* ---- reading state back --------------------------------------------------------------- */
int32_t sfgpu_committed_scores(sfgpu_ctx* ctx, int64_t* out_scores /* [R][2] host */);
/* stateless full recompute on a scratch copy (ConstraintSet::evaluate_all; FullAssert invariant
 * cached == fresh, solverforge-solver/src/scope/solver/scope_core.rs:642-653) */
int32_t sfgpu_evaluate_all(sfgpu_ctx* ctx, int64_t* out_scores /* [R][2] host */);
int32_t sfgpu_get_scalar_state(sfgpu_ctx* ctx, uint32_t variable, int32_t* out_values /* [R][n_rows] host */);
int32_t sfgpu_get_list_state(sfgpu_ctx* ctx, uint32_t variable, uint32_t* out_offsets /* [R][n_owners+1] */,
                             uint32_t* out_elems /* [R][n_elements_capacity] */);
int32_t sfgpu_list_capacity(sfgpu_ctx* ctx, uint32_t variable, uint32_t* out_capacity);

/* ---- replicas across GPUs -------------------------------------------------------------- */
/* Order-preserving packed key of each replica's committed score for a MAX all-reduce:
 * key = ((hard + 2^22) << 40) | (soft + 2^39), valid for hard in [-2^22, 2^22), soft in [-2^39, 2^39)
 * (out-of-range levels saturate — sfgpu_sync_best below is exact for every score). out_keys is a DEVICE pointer to R int64 (e.g. a torch tensor the
 * caller then passes to torch.distributed.all_reduce(MAX) / ncclAllReduce). */
int32_t sfgpu_pack_best_keys(sfgpu_ctx* ctx, int64_t* out_keys);

/* Best-score sync for hosts without torch (SURVEY §8b `sfgpu_sync_best`, §8e): NCCL resolved at run time
 * (dlopen of libnccl.so.2; without it these calls fail with SFGPU_E_NCCL, everything else works).
 * One communicator per process / GPU: rank 0 creates the 128-byte id and ships it to the other ranks by any
 * host channel (the reference's SolverManager jobs share a process; across processes: MPI, a file, a socket). */
int32_t sfgpu_comm_unique_id(uint8_t* out_id128);
int32_t sfgpu_comm_init_rank(int32_t n_ranks, const uint8_t* id128, int32_t rank, int32_t device, void** out_comm);
int32_t sfgpu_comm_destroy(void* comm);
/* All ranks learn the best score over every replica of every rank, its owner rank (lowest rank on ties) and the
 * replica index inside the owner. scores: DEVICE pointer (flags & SFGPU_DEVICE_IO) to R (hard, soft) pairs —
 * e.g. the best-so-far scores of a solve loop — or NULL for the committed scores. One device reduction over the
 * replicas, one ncclAllGather of 24 B per rank on the context's stream, a local lexicographic max: exact over
 * the whole int64 range (no packed key). nccl_comm NULL = single rank (no collective). Synchronises the stream —
 * unless flags & SFGPU_SYNC_ASYNC: then out_best[2] / out_owner_rank / out_owner_replica are DEVICE pointers, the
 * reduction over the gathered records runs on the device too and the call returns without draining the stream (a
 * solver loop keeps queueing steps behind the collective). */
int32_t sfgpu_sync_best(sfgpu_ctx* ctx, void* nccl_comm, uint32_t flags, const int64_t* scores, int64_t* out_best,
                        int32_t* out_owner_rank, uint32_t* out_owner_replica);

/* ---- timing ----------------------------------------------------------------------------- */
/* device time of the most recent scoring kernel launch of this context (CUDA events recorded on
 * the launching stream around the kernel), in nanoseconds; synchronizes the stream. */
int32_t sfgpu_last_kernel_ns(sfgpu_ctx* ctx, uint64_t* out_ns);
/* device times (ns) of the most recent <= max_n scoring-kernel launches, oldest first: every score /
 * step call records a CUDA event pair on the launching stream around its dominant kernel only
 * (ring of 512), so a benchmark can read per-kernel durations of its timed region afterwards. */
int32_t sfgpu_kernel_times_ns(sfgpu_ctx* ctx, uint32_t max_n, uint64_t* out_ns, uint32_t* out_n);
/* number of kernels this context has launched since creation */
int32_t sfgpu_launch_count(sfgpu_ctx* ctx, uint64_t* out_count);
/* Which scalar scoring program the committed model runs: >= 0 = index of the monomorphised kernel
 * (the tuple of scalar constraint kinds has a template instantiation — the counterpart of the reference's
 * monomorphised ConstraintSet tuples, solverforge-scoring/src/api/constraint_set/incremental.rs:339-408),
 * -1 = the constraint-table interpreter (any program). Results are identical either way. */
int32_t sfgpu_scalar_program(sfgpu_ctx* ctx, int32_t* out_program);

#ifdef __cplusplus
}
#endif
#endif /* SFGPU_H */
