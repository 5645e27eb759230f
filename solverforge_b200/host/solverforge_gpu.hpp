// solverforge_gpu.hpp — C++ host-side mirror of the reference's scoring interface, above the C ABI.
//
// The reference is Rust; this image has no Rust toolchain, so the host layer a Rust
// `solverforge-gpu` crate would provide is written in C++17 with the same names, argument
// meaning and error behaviour (INTEGRATION.md shows the Rust binding):
//
//   HardSoftScore / HardSoftDecimalScore   solverforge-core/src/score/hard_soft.rs:35-153, hard_soft_decimal.rs:14-221
//   ConstraintFactory .for_each(..)...     solverforge-scoring/src/stream/factory.rs:43-73
//   Director surface (calculate_score, fresh_score, ...)   solverforge-scoring/src/director/traits.rs:27-95
//   ScalarEdit                             solverforge-solver/src/planning/scalar/candidate.rs:6-12
//   Acceptor / LocalSearchForager replay   solverforge-solver/src/phase/localsearch/{acceptor/*,forager.rs,phase/candidates.rs:66-282}
//
// Everything that touches scores runs in libsfgpu (CUDA). Errors surface as sf::GpuError; there is no
// CPU scoring fallback (north_star). The host replay below only *orders* already-computed scores, which
// is what the reference's acceptor/forager do.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sfgpu.h"

namespace sf {

struct GpuError : std::runtime_error {
  int code;
  GpuError(int c, const std::string& m) : std::runtime_error("libsfgpu error " + std::to_string(c) + ": " + m), code(c) {}
};

struct HardSoftScore {
  int64_t hard = 0, soft = 0;
  static constexpr HardSoftScore of(int64_t h, int64_t s) { return {h, s}; }
  static constexpr HardSoftScore of_hard(int64_t h) { return {h, 0}; }
  static constexpr HardSoftScore of_soft(int64_t s) { return {0, s}; }
  static constexpr HardSoftScore ZERO() { return {0, 0}; }
  static constexpr HardSoftScore ONE_HARD() { return {1, 0}; }
  static constexpr HardSoftScore ONE_SOFT() { return {0, 1}; }
  bool is_feasible() const { return hard >= 0; }
  HardSoftScore operator+(HardSoftScore o) const { return {hard + o.hard, soft + o.soft}; }
  HardSoftScore operator-(HardSoftScore o) const { return {hard - o.hard, soft - o.soft}; }
  HardSoftScore operator-() const { return {-hard, -soft}; }
  bool operator==(HardSoftScore o) const { return hard == o.hard && soft == o.soft; }
  bool operator!=(HardSoftScore o) const { return !(*this == o); }
  bool operator<(HardSoftScore o) const { return hard != o.hard ? hard < o.hard : soft < o.soft; }
  bool operator>(HardSoftScore o) const { return o < *this; }
  bool operator<=(HardSoftScore o) const { return !(o < *this); }
  bool operator>=(HardSoftScore o) const { return !(*this < o); }
  int64_t level(int l) const { return l == 0 ? hard : soft; }
  // Score::abs / Score::multiply (solverforge-core/src/score/macros.rs:61-71)
  HardSoftScore abs() const { return {hard < 0 ? -hard : hard, soft < 0 ? -soft : soft}; }
  HardSoftScore multiply(double m) const {
    return {(int64_t)std::round((double)hard * m), (int64_t)std::round((double)soft * m)};
  }
};
// HardSoftDecimalScore shares the layout with levels pre-scaled by 100000.
struct HardSoftDecimalScore : HardSoftScore {
  static constexpr int64_t SCALE = 100000;
  static constexpr HardSoftScore of(int64_t h, int64_t s) { return {h * SCALE, s * SCALE}; }
  static constexpr HardSoftScore of_scaled(int64_t h, int64_t s) { return {h, s}; }
};

struct ScalarEdit {  // planning/scalar/candidate.rs:6-12 (descriptor + variable are fixed per model here)
  uint32_t entity_index;
  int32_t to_value;  // -1 = None
};
struct ScalarSwap {  // heuristic/move/swap.rs: the two entities exchange their values
  uint32_t left_entity, right_entity;
};
struct ListChange {  // heuristic/move/list_kernel/change.rs:15-21
  uint32_t source_entity, source_position, destination_entity, destination_position;
};
struct ListSwap {  // heuristic/move/list_kernel/swap.rs:16-30
  uint32_t first_entity, first_position, second_entity, second_position;
};
struct ListReverse {  // heuristic/move/list_kernel/reverse.rs:14-19: reverses [start, end) of one list
  uint32_t entity, start, end, reserved = 0;
};
// heuristic/move/sublist_change.rs: relocates [start, start + size) of source_entity to dest_entity at
// dest_position (a position of the post-removal list when both entities are the same)
struct SublistChange {
  uint32_t source_entity, packed_segment, dest_entity, dest_position;
  SublistChange(uint32_t se, uint32_t start, uint32_t end, uint32_t de, uint32_t dp)
      : source_entity(se), packed_segment(SFGPU_SEG(start, end - start)), dest_entity(de), dest_position(dp) {}
};
// heuristic/move/sublist_swap.rs: exchanges [first_start, first_end) and [second_start, second_end)
struct SublistSwap {
  uint32_t first_entity, packed_first, second_entity, packed_second;
  SublistSwap(uint32_t e1, uint32_t s1, uint32_t t1, uint32_t e2, uint32_t s2, uint32_t t2)
      : first_entity(e1), packed_first(SFGPU_SEG(s1, t1 - s1)), second_entity(e2), packed_second(SFGPU_SEG(s2, t2 - s2)) {}
};
// acceptor + forager of one fused device step (sfgpu_forage_params) and what it returns per replica
struct StepParams {
  int acceptor = 0;            // 0 accept all, 1 > last, 2 >= last || >= threshold, 3 > last || >= threshold
  bool random_ties = true;     // ScoreTieBreak: reservoir (forager.rs:99-155) or first
  uint32_t accepted_limit = 0; // 0 = BestScore, N = AcceptedCount(N)
};
struct StepResult {
  std::vector<uint32_t> index;      // winning CandidateId per replica (UINT32_MAX: none accepted)
  std::vector<struct HardSoftScore> best;
  std::vector<uint32_t> evaluated;  // moves_evaluated
  std::vector<uint32_t> winner_rows;  // [R][4] list moves / [R][2] scalar edits
};
struct SolveParams {  // sfgpu_solve_params: device-resident loop (phase.rs:237-320)
  uint32_t max_nearby = 20, n_steps = 0;
  int acceptor = 2;            // 1 HillClimbing, 2 LateAcceptance, 3 GreatDeluge, 4 StepCountingHillClimbing,
                               // 5 DiversifiedLateAcceptance, 6 SimulatedAnnealing (late_size = calibration samples,
                               // acceptor_real = decay), 7 TabuSearch (late_size = tabu_tenures(..), step_count_limit
                               // bit 0 = aspiration)
  uint32_t late_size = 400;
  // TabuSearchAcceptor::new(entity_tabu, value_tabu, move_tabu, undo_move_tabu) packed for acceptor 7 (each <= 64)
  static uint32_t tabu_tenures(uint32_t entity, uint32_t value, uint32_t move, uint32_t undo_move) {
    return entity | (value << 8) | (move << 16) | (undo_move << 24);
  }
  bool random_ties = true;
  uint32_t accepted_limit = 0;
  uint64_t seed_base = 0;
  bool restore_best = false;
  double acceptor_real = 0.0;  // rain_speed / tolerance
  uint64_t step_count_limit = 0;
};
struct SolveResult {
  std::vector<struct HardSoftScore> best;
  std::vector<uint64_t> moves_evaluated, accepted_steps;
};


struct Weight {
  sfgpu_weight w;
  static Weight constant(HardSoftScore s) {
    if (s.hard != 0 && s.soft != 0) throw GpuError(SFGPU_E_UNSUPPORTED, "a constant weight must sit on one level");
    return {{SFGPU_W_CONST, s.hard != 0 ? 0 : 1, s.hard != 0 ? s.hard : s.soft, 0}};
  }
  static Weight hard(int fn, int64_t a = 1, int64_t b = 0) { return {{fn, 0, a, b}}; }
  static Weight soft(int fn, int64_t a = 1, int64_t b = 0) { return {{fn, 1, a, b}}; }
};

class GpuScoreDirector;

// joiners / collectors (device forms of the reference closures)
struct AdjacentEqual { uint32_t csr; };
struct EqualKey { uint32_t column = UINT32_MAX; int64_t col_mul = 0, var_mul = 1; };
struct EqualVarToRow {};
struct EqualId { uint32_t column = UINT32_MAX; };
struct PathCost { uint32_t matrix; uint32_t depot; };
struct ListSum { uint32_t column; };
struct Count {};
struct Sum { uint32_t column; };
struct LoadBalance { uint32_t metric_column = UINT32_MAX; };
// consecutive_runs(|e| e.point) (stream/collector/runs.rs): the weight applies to every run's point_count
struct ConsecutiveRuns { uint32_t point_column; uint32_t n_points; };
// Data form of a `Projection<A>` (stream/projected_stream/source.rs:13-24): an ASSIGNED entity e emits one row per
// entry j of csr row e with key offset csr.col[j] (< keys_per_value) and amount amounts[j] (empty = 1 each); the
// group key of a row is var[e] * keys_per_value + key offset. MAX_EMITS <= 8.
struct Projection {
  uint32_t csr;
  uint32_t keys_per_value;
  std::vector<uint32_t> row_ptr;  // the csr's row pointers (host copy, for the per-row terminal)
  std::vector<int64_t> amounts;
};

struct Terminal {
  GpuScoreDirector* d;
  sfgpu_constraint_desc desc;
  uint32_t named(const std::string& name);
};

struct GroupedStream {
  GpuScoreDirector* d;
  uint32_t collection, column;
  bool load_balance, complemented = false;
  int64_t default_result = 0;
  GroupedStream complement(uint32_t /*targets*/, int64_t def) const {
    GroupedStream g = *this;
    g.complemented = true;
    g.default_result = def;
    return g;
  }
  Terminal impact(int imp, Weight w, uint32_t key_offset_column = UINT32_MAX) const {
    sfgpu_constraint_desc c{};
    c.kind = load_balance ? SFGPU_K_LOAD_BALANCE : SFGPU_K_GROUP;
    c.impact = imp;
    c.weight = w.w;
    c.collection = collection;
    c.aux0 = column;
    c.aux1 = key_offset_column;
    c.p0 = complemented ? 1 : 0;
    c.p1 = default_result;
    return {d, c};
  }
  // key_offset_column: per-value column replacing the weight's b for that key (|key, result| weights)
  Terminal penalize(Weight w, uint32_t key_offset_column = UINT32_MAX) const {
    return impact(SFGPU_PENALTY, w, key_offset_column);
  }
  Terminal reward(Weight w, uint32_t key_offset_column = UINT32_MAX) const {
    return impact(SFGPU_REWARD, w, key_offset_column);
  }
};

struct BiStream {
  GpuScoreDirector* d;
  uint32_t collection;
  int joiner_kind;  // 0 adjacent, 1 key, 2 var->row
  AdjacentEqual adj{};
  EqualKey key{};
  Terminal impact(int imp, HardSoftScore w) const {
    sfgpu_constraint_desc c{};
    c.impact = imp;
    c.weight = Weight::constant(w).w;
    c.collection = collection;
    if (joiner_kind == 0) {
      c.kind = SFGPU_K_PAIR_CSR_EQUAL;
      c.aux0 = adj.csr;
    } else if (joiner_kind == 1) {
      c.kind = SFGPU_K_PAIR_KEY_EQUAL;
      c.aux0 = key.column;
      c.p0 = key.col_mul;
      c.p1 = key.var_mul;
    } else {
      throw GpuError(SFGPU_E_UNSUPPORTED, "joiner is not expressible on device without group_by");
    }
    return {d, c};
  }
  Terminal penalize(HardSoftScore w) const { return impact(SFGPU_PENALTY, w); }
  Terminal reward(HardSoftScore w) const { return impact(SFGPU_REWARD, w); }
  GroupedStream group_by(Count) const { return {d, collection, UINT32_MAX, false}; }
  GroupedStream group_by(Sum s) const { return {d, collection, s.column, false}; }
};

struct ExistsStream {
  GpuScoreDirector* d;
  uint32_t collection, column;
  int mode;
  Terminal impact(int imp, HardSoftScore w) const {
    sfgpu_constraint_desc c{};
    c.kind = SFGPU_K_EXISTS_FLAT;
    c.impact = imp;
    c.weight = Weight::constant(w).w;
    c.collection = collection;
    c.variable = 0x80000000u;
    c.aux0 = column;
    c.p0 = mode;
    return {d, c};
  }
  Terminal penalize(HardSoftScore w) const { return impact(SFGPU_PENALTY, w); }
  Terminal reward(HardSoftScore w) const { return impact(SFGPU_REWARD, w); }
};

struct UniStream {
  GpuScoreDirector* d;
  uint32_t collection;
  int filter = 2;  // 0 unassigned, 1 assigned, 2 always
  uint32_t mask = UINT32_MAX;  // `.filter(|e| e.flag)` over a static 0/1 column
  UniStream unassigned() const { return {d, collection, 0, mask}; }
  UniStream assigned() const { return {d, collection, 1, mask}; }
  UniStream filtered(uint32_t mask_column) const { return {d, collection, filter, mask_column}; }
  UniStream flattened() const { return *this; }
  Terminal uni(int imp, Weight w, uint32_t column, bool by_value) const {
    sfgpu_constraint_desc c{};
    c.kind = SFGPU_K_UNI;
    c.impact = imp;
    c.weight = w.w;
    c.collection = collection;
    c.aux0 = column;
    c.aux1 = mask;
    c.p0 = filter;
    c.p1 = by_value ? 1 : 0;
    return {d, c};
  }
  Terminal penalize(HardSoftScore w) const { return uni(SFGPU_PENALTY, Weight::constant(w), UINT32_MAX, false); }
  Terminal reward(HardSoftScore w) const { return uni(SFGPU_REWARD, Weight::constant(w), UINT32_MAX, false); }
  Terminal penalize(Weight w, uint32_t column, bool by_value = false) const { return uni(SFGPU_PENALTY, w, column, by_value); }
  Terminal penalize(Weight w, PathCost pc) const {
    sfgpu_constraint_desc c{};
    c.kind = SFGPU_K_LIST_PATH_COST;
    c.impact = SFGPU_PENALTY;
    c.weight = w.w;
    c.collection = collection;
    c.variable = 0x80000000u;
    c.aux0 = pc.matrix;
    c.p0 = pc.depot;
    return {d, c};
  }
  Terminal penalize(Weight w, ListSum ls) const {
    sfgpu_constraint_desc c{};
    c.kind = SFGPU_K_LIST_SUM;
    c.impact = SFGPU_PENALTY;
    c.weight = w.w;
    c.collection = collection;
    c.variable = 0x80000000u;
    c.aux0 = ls.column;
    return {d, c};
  }
  BiStream join(const UniStream&, AdjacentEqual j) const { return {d, collection, 0, j, {}}; }
  BiStream join(const UniStream&, EqualKey j) const { return {d, collection, 1, {}, j}; }
  BiStream join(uint32_t /*values*/, EqualVarToRow) const { return {d, collection, 2, {}, {}}; }
  ExistsStream if_exists(const UniStream&, EqualId j) const { return {d, collection, j.column, 0}; }
  ExistsStream if_not_exists(const UniStream&, EqualId j) const { return {d, collection, j.column, 1}; }
  GroupedStream group_by(Count) const { return {d, collection, UINT32_MAX, false}; }
  GroupedStream group_by(Sum s) const { return {d, collection, s.column, false}; }
  GroupedStream group_by(LoadBalance lb) const { return {d, collection, lb.metric_column, true}; }
  // group_by(var, consecutive_runs(point)).penalize(sum over runs of w(point_count)) -> SFGPU_K_RUNS
  Terminal penalize_runs(ConsecutiveRuns r, Weight w) const {
    sfgpu_constraint_desc c{};
    c.kind = SFGPU_K_RUNS;
    c.impact = SFGPU_PENALTY;
    c.weight = w.w;
    c.collection = collection;
    c.aux0 = r.point_column;
    c.aux1 = UINT32_MAX;
    c.p0 = r.n_points;
    return {d, c};
  }
};

// for_each(E).project(P)... — projected scoring rows (stream/projected_stream/uni.rs)
struct ProjectedGroupedStream {
  GpuScoreDirector* d;
  uint32_t collection;
  Projection p;
  bool sum;
  Terminal impact(int imp, Weight w) const;  // lowers to SFGPU_K_PROJECT_GROUP
  Terminal penalize(Weight w) const { return impact(SFGPU_PENALTY, w); }
  Terminal reward(Weight w) const { return impact(SFGPU_REWARD, w); }
};
struct ProjectedStream {
  GpuScoreDirector* d;
  uint32_t collection;
  Projection p;
  // .penalize(|row| w(row.amount)): every row scored on its own (constraint/projected/uni.rs); CONST or
  // LINEAR (b = 0) weights sum per entity into one uni constraint over assigned entities
  Terminal impact(int imp, Weight w) const;
  Terminal penalize(Weight w) const { return impact(SFGPU_PENALTY, w); }
  Terminal reward(Weight w) const { return impact(SFGPU_REWARD, w); }
  // .join(equal(key)).penalize(CONST): every unordered pair of rows with equal keys (constraint/projected/bi.rs)
  Terminal penalize_pairs(HardSoftScore w) const {
    ProjectedGroupedStream g{d, collection, p, false};
    Weight cw = Weight::constant(w);
    cw.w.fn = SFGPU_W_PAIRS;
    return g.impact(SFGPU_PENALTY, cw);
  }
  ProjectedGroupedStream group_by(Count) const { return {d, collection, p, false}; }
  ProjectedGroupedStream group_by(Sum) const { return {d, collection, p, true}; }  // sums the projection's amounts
};
inline ProjectedStream project(const UniStream& s, Projection p) { return {s.d, s.collection, std::move(p)}; }

struct ConstraintFactory {
  GpuScoreDirector* d;
  explicit ConstraintFactory(GpuScoreDirector& dir) : d(&dir) {}
  UniStream for_each(uint32_t collection) const { return {d, collection, 2, UINT32_MAX}; }
};

// RAII owner of one sfgpu_ctx — the Director of R replicas.
class GpuScoreDirector {
 public:
  explicit GpuScoreDirector(uint32_t n_replicas = 1, int device = 0, void* stream = nullptr, uint64_t flags = 0)
      : R_(n_replicas) {
    int rc = sfgpu_ctx_create(device, flags, stream, &ctx_);
    if (rc != SFGPU_OK) throw GpuError(rc, sfgpu_last_error(nullptr));
    check(sfgpu_model_begin(ctx_, n_replicas));
  }
  ~GpuScoreDirector() {
    if (ctx_) sfgpu_ctx_destroy(ctx_);
  }
  GpuScoreDirector(const GpuScoreDirector&) = delete;
  GpuScoreDirector& operator=(const GpuScoreDirector&) = delete;

  uint32_t replicas() const { return R_; }
  sfgpu_ctx* raw() const { return ctx_; }
  void check(int rc) const {
    if (rc != SFGPU_OK) throw GpuError(rc, sfgpu_last_error(ctx_));
  }

  uint32_t add_collection(const std::string& name, uint32_t n_rows, int32_t descriptor_index = -1) {
    uint32_t id;
    check(sfgpu_add_collection(ctx_, name.c_str(), n_rows, descriptor_index, &id));
    return id;
  }
  uint32_t add_column(uint32_t coll, const std::string& name, const std::vector<int64_t>& v) {
    uint32_t id;
    check(sfgpu_add_column_i64(ctx_, coll, name.c_str(), v.data(), &id));
    return id;
  }
  uint32_t add_scalar_variable(uint32_t coll, const std::string& name, uint32_t n_values, bool allows_unassigned) {
    uint32_t id;
    check(sfgpu_add_scalar_variable(ctx_, coll, name.c_str(), n_values, allows_unassigned ? 1 : 0, &id));
    return id;
  }
  // working_solution() of the scalar variable: values[R][n_entities], -1 = None
  std::vector<std::vector<int32_t>> scalar_state(uint32_t n_entities) {
    std::vector<int32_t> flat((size_t)R_ * n_entities);
    check(sfgpu_get_scalar_state(ctx_, 0, flat.data()));
    std::vector<std::vector<int32_t>> out(R_);
    for (uint32_t r = 0; r < R_; ++r) out[r].assign(flat.begin() + (size_t)r * n_entities, flat.begin() + (size_t)(r + 1) * n_entities);
    return out;
  }
  uint32_t add_list_variable(uint32_t owners, uint32_t elements, const std::string& name) {
    uint32_t id;
    check(sfgpu_add_list_variable(ctx_, owners, elements, name.c_str(), &id));
    return id;
  }
  uint32_t add_csr(const std::string& name, const std::vector<uint32_t>& row_ptr, const std::vector<uint32_t>& col) {
    uint32_t id;
    check(sfgpu_add_csr(ctx_, name.c_str(), (uint32_t)row_ptr.size() - 1, row_ptr.data(), col.data(), &id));
    return id;
  }
  uint32_t add_matrix(const std::string& name, uint32_t rows, uint32_t cols, const std::vector<int64_t>& v,
                      bool cost_semantics) {
    uint32_t id;
    check(sfgpu_add_matrix_i64(ctx_, name.c_str(), rows, cols, v.data(), cost_semantics ? 1 : 0, &id));
    return id;
  }
  void set_scalar_state(const std::vector<int32_t>& values, bool per_replica = false) {
    check(sfgpu_set_scalar_state(ctx_, 0, values.data(), per_replica ? 1 : 0));
  }
  void set_list_state(const std::vector<uint32_t>& offsets, const std::vector<uint32_t>& elems, bool per_replica = false) {
    check(sfgpu_set_list_state(ctx_, 0x80000000u, offsets.data(), elems.data(), per_replica ? 1 : 0));
  }
  std::vector<HardSoftScore> commit() {
    std::vector<HardSoftScore> s(R_);
    check(sfgpu_model_commit(ctx_, reinterpret_cast<int64_t*>(s.data())));
    return s;
  }
  // Director::calculate_score / fresh_score
  std::vector<HardSoftScore> calculate_score() {
    std::vector<HardSoftScore> s(R_);
    check(sfgpu_committed_scores(ctx_, reinterpret_cast<int64_t*>(s.data())));
    return s;
  }
  std::vector<HardSoftScore> fresh_score() {
    std::vector<HardSoftScore> s(R_);
    check(sfgpu_evaluate_all(ctx_, reinterpret_cast<int64_t*>(s.data())));
    return s;
  }
  // the batch seam: scores of a whole neighbourhood (evaluate_candidate for every pull)
  void score_candidates(const std::vector<ScalarEdit>& batch, const std::vector<uint64_t>& cand_offsets,
                        std::vector<HardSoftScore>& scores, std::vector<uint8_t>& doable) {
    scores.resize(batch.size());
    doable.resize(batch.size());
    check(sfgpu_score_change(ctx_, 0, batch.size(), cand_offsets.data(), reinterpret_cast<const uint32_t*>(batch.data()),
                             reinterpret_cast<int64_t*>(scores.data()), doable.data()));
  }
  void score_candidates(const std::vector<ListChange>& batch, const std::vector<uint64_t>& cand_offsets,
                        std::vector<HardSoftScore>& scores, std::vector<uint8_t>& doable) {
    scores.resize(batch.size());
    doable.resize(batch.size());
    check(sfgpu_score_list_change(ctx_, 0, batch.size(), cand_offsets.data(),
                                  reinterpret_cast<const uint32_t*>(batch.data()),
                                  reinterpret_cast<int64_t*>(scores.data()), doable.data()));
  }
  void score_candidates(const std::vector<ScalarSwap>& batch, const std::vector<uint64_t>& cand_offsets,
                        std::vector<HardSoftScore>& scores, std::vector<uint8_t>& doable) {
    scores.resize(batch.size());
    doable.resize(batch.size());
    check(sfgpu_score_swap(ctx_, 0, batch.size(), cand_offsets.data(), reinterpret_cast<const uint32_t*>(batch.data()),
                           reinterpret_cast<int64_t*>(scores.data()), doable.data()));
  }
  void score_candidates(const std::vector<ListSwap>& batch, const std::vector<uint64_t>& cand_offsets,
                        std::vector<HardSoftScore>& scores, std::vector<uint8_t>& doable) {
    scores.resize(batch.size());
    doable.resize(batch.size());
    check(sfgpu_score_list_swap(ctx_, 0, batch.size(), cand_offsets.data(),
                                reinterpret_cast<const uint32_t*>(batch.data()),
                                reinterpret_cast<int64_t*>(scores.data()), doable.data()));
  }
  void score_candidates(const std::vector<ListReverse>& batch, const std::vector<uint64_t>& cand_offsets,
                        std::vector<HardSoftScore>& scores, std::vector<uint8_t>& doable) {
    scores.resize(batch.size());
    doable.resize(batch.size());
    check(sfgpu_score_list_reverse(ctx_, 0, batch.size(), cand_offsets.data(),
                                   reinterpret_cast<const uint32_t*>(batch.data()),
                                   reinterpret_cast<int64_t*>(scores.data()), doable.data()));
  }
  void apply(const std::vector<ListReverse>& one_per_replica, const uint8_t* mask = nullptr) {
    check(sfgpu_apply_list_reverse(ctx_, 0, reinterpret_cast<const uint32_t*>(one_per_replica.data()), mask));
  }
  void score_candidates(const std::vector<SublistChange>& batch, const std::vector<uint64_t>& cand_offsets,
                        std::vector<HardSoftScore>& scores, std::vector<uint8_t>& doable) {
    scores.resize(batch.size());
    doable.resize(batch.size());
    check(sfgpu_score_sublist_change(ctx_, 0, batch.size(), cand_offsets.data(),
                                     reinterpret_cast<const uint32_t*>(batch.data()),
                                     reinterpret_cast<int64_t*>(scores.data()), doable.data()));
  }
  void apply(const std::vector<SublistChange>& one_per_replica, const uint8_t* mask = nullptr) {
    check(sfgpu_apply_sublist_change(ctx_, 0, reinterpret_cast<const uint32_t*>(one_per_replica.data()), mask));
  }
  void score_candidates(const std::vector<SublistSwap>& batch, const std::vector<uint64_t>& cand_offsets,
                        std::vector<HardSoftScore>& scores, std::vector<uint8_t>& doable) {
    scores.resize(batch.size());
    doable.resize(batch.size());
    check(sfgpu_score_sublist_swap(ctx_, 0, batch.size(), cand_offsets.data(),
                                   reinterpret_cast<const uint32_t*>(batch.data()),
                                   reinterpret_cast<int64_t*>(scores.data()), doable.data()));
  }
  void apply(const std::vector<SublistSwap>& one_per_replica, const uint8_t* mask = nullptr) {
    check(sfgpu_apply_sublist_swap(ctx_, 0, reinterpret_cast<const uint32_t*>(one_per_replica.data()), mask));
  }
  // whole step over the SublistChange neighbourhood, enumerated on device (sfgpu_step_sublist_change)
  void step_sublist_change(uint32_t min_size, uint32_t max_size, const StepParams& p, const std::vector<uint64_t>& step_seeds,
                           const std::vector<int64_t>& ref_scores, std::vector<uint32_t>& out_index,
                           std::vector<HardSoftScore>& out_best, std::vector<uint32_t>& out_evaluated,
                           std::vector<uint32_t>& out_winner_rows, bool apply_winners) {
    sfgpu_forage_params fp{p.acceptor, p.random_ties ? 1 : 0, p.accepted_limit, 0};
    const size_t R = step_seeds.size();
    out_index.resize(R);
    out_best.resize(R);
    out_evaluated.resize(R);
    out_winner_rows.resize(R * 4);
    check(sfgpu_step_sublist_change(ctx_, 0, min_size, max_size, &fp, step_seeds.data(),
                                    ref_scores.empty() ? nullptr : ref_scores.data(), out_index.data(),
                                    reinterpret_cast<int64_t*>(out_best.data()), out_evaluated.data(),
                                    out_winner_rows.data(), apply_winners ? 1 : 0));
  }
  // the same over the SublistSwap neighbourhood (sfgpu_step_sublist_swap)
  void step_sublist_swap(uint32_t min_size, uint32_t max_size, const StepParams& p, const std::vector<uint64_t>& step_seeds,
                         const std::vector<int64_t>& ref_scores, std::vector<uint32_t>& out_index,
                         std::vector<HardSoftScore>& out_best, std::vector<uint32_t>& out_evaluated,
                         std::vector<uint32_t>& out_winner_rows, bool apply_winners) {
    sfgpu_forage_params fp{p.acceptor, p.random_ties ? 1 : 0, p.accepted_limit, 0};
    const size_t R = step_seeds.size();
    out_index.resize(R);
    out_best.resize(R);
    out_evaluated.resize(R);
    out_winner_rows.resize(R * 4);
    check(sfgpu_step_sublist_swap(ctx_, 0, min_size, max_size, &fp, step_seeds.data(),
                                  ref_scores.empty() ? nullptr : ref_scores.data(), out_index.data(),
                                  reinterpret_cast<int64_t*>(out_best.data()), out_evaluated.data(),
                                  out_winner_rows.data(), apply_winners ? 1 : 0));
  }
  // CompoundScalarMove batch: candidate i owns edits [edit_offsets[i], edit_offsets[i+1]) (compound_scalar.rs:289-319)
  void score_compound(const std::vector<uint64_t>& edit_offsets, const std::vector<ScalarEdit>& edits,
                      const std::vector<uint64_t>& cand_offsets, std::vector<HardSoftScore>& scores,
                      std::vector<uint8_t>& doable) {
    const size_t n = edit_offsets.size() - 1;
    scores.resize(n);
    doable.resize(n);
    check(sfgpu_score_compound(ctx_, 0, n, cand_offsets.data(), edit_offsets.data(),
                               reinterpret_cast<const uint32_t*>(edits.data()),
                               reinterpret_cast<int64_t*>(scores.data()), doable.data()));
  }
  // acceptor + forager replay on device over scored rows (sfgpu_argbest[_gated]); ref = {last_step, threshold}
  StepResult argbest(const std::vector<HardSoftScore>& scores, const std::vector<uint8_t>& doable,
                     const std::vector<uint64_t>& cand_offsets, StepParams p, const std::vector<uint64_t>& step_seeds,
                     const std::vector<HardSoftScore>& ref_pairs = {}, const uint8_t* gates = nullptr) {
    StepResult out = make_result(0);
    sfgpu_forage_params fp{p.acceptor, p.random_ties ? 1 : 0, p.accepted_limit, 0};
    check(sfgpu_argbest_gated(ctx_, 0, &fp, cand_offsets.data(), reinterpret_cast<const int64_t*>(scores.data()),
                              doable.data(), gates, step_seeds.empty() ? nullptr : step_seeds.data(),
                              ref_pairs.empty() ? nullptr : reinterpret_cast<const int64_t*>(ref_pairs.data()),
                              out.index.data(), reinterpret_cast<int64_t*>(out.best.data()), out.evaluated.data()));
    return out;
  }
  // whole steps on device (evaluate_candidates + pick [+ apply], phase/step.rs:30-225)
  StepResult step_change(StepParams p, const std::vector<uint64_t>& step_seeds,
                         const std::vector<HardSoftScore>& ref_pairs = {}, bool apply_winners = false) {
    StepResult out = make_result(2);
    sfgpu_forage_params fp{p.acceptor, p.random_ties ? 1 : 0, p.accepted_limit, 0};
    check(sfgpu_step_change(ctx_, 0, &fp, step_seeds.empty() ? nullptr : step_seeds.data(),
                            ref_pairs.empty() ? nullptr : reinterpret_cast<const int64_t*>(ref_pairs.data()), nullptr,
                            nullptr, nullptr, nullptr, out.index.data(), reinterpret_cast<int64_t*>(out.best.data()),
                            out.evaluated.data(), out.winner_rows.data(), apply_winners ? 1 : 0));
    return out;
  }
  StepResult step_nearby_list_change(uint32_t max_nearby, StepParams p, const std::vector<uint64_t>& step_seeds,
                                     const std::vector<HardSoftScore>& ref_pairs = {}, bool apply_winners = false,
                                     bool swap_moves = false) {
    StepResult out = make_result(4);
    sfgpu_forage_params fp{p.acceptor, p.random_ties ? 1 : 0, p.accepted_limit, 0};
    auto fn = swap_moves ? sfgpu_step_nearby_list_swap : sfgpu_step_nearby_list_change;
    check(fn(ctx_, 0, max_nearby, &fp, step_seeds.empty() ? nullptr : step_seeds.data(),
             ref_pairs.empty() ? nullptr : reinterpret_cast<const int64_t*>(ref_pairs.data()), nullptr, nullptr, nullptr,
             nullptr, out.index.data(), reinterpret_cast<int64_t*>(out.best.data()), out.evaluated.data(),
             out.winner_rows.data(), apply_winners ? 1 : 0));
    return out;
  }
  StepResult step_nearby_list_swap(uint32_t max_nearby, StepParams p, const std::vector<uint64_t>& step_seeds,
                                   const std::vector<HardSoftScore>& ref_pairs = {}, bool apply_winners = false) {
    return step_nearby_list_change(max_nearby, p, step_seeds, ref_pairs, apply_winners, true);
  }
  // device-resident loops (solve_local_search_with_resources for every replica)
  SolveResult solve(const SolveParams& p, bool scalar_model) {
    sfgpu_solve_params sp{};
    sp.max_nearby = p.max_nearby;
    sp.n_steps = p.n_steps;
    sp.acceptor = p.acceptor;
    sp.late_size = p.late_size;
    sp.tie_mode = p.random_ties ? 1 : 0;
    sp.accepted_limit = p.accepted_limit;
    sp.seed_base = p.seed_base;
    sp.restore_best = p.restore_best ? 1 : 0;
    sp.acceptor_real = p.acceptor_real;
    sp.step_count_limit = p.step_count_limit;
    SolveResult out;
    out.best.resize(R_);
    out.moves_evaluated.resize(R_);
    out.accepted_steps.resize(R_);
    auto fn = scalar_model ? sfgpu_solve_change : sfgpu_solve_nearby_list_change;
    check(fn(ctx_, &sp, reinterpret_cast<int64_t*>(out.best.data()), out.moves_evaluated.data(),
             out.accepted_steps.data()));
    return out;
  }
  // ---- union of neighbourhoods in the reference's seeded pull order (sfgpu_step_union / sfgpu_solve_union):
  // VecUnionSelector over move families (decorator/vec_union.rs:204-366), leaves in SelectionOrder `selection_order`
  static sfgpu_union_desc union_desc(const std::vector<sfgpu_union_child>& children, int union_order = SFGPU_UNION_STRATIFIED_RANDOM,
                                     int selection_order = SFGPU_ORDER_RANDOM, uint32_t window = 0, uint32_t max_window = 0) {
    sfgpu_union_desc d{};
    d.n_children = (uint32_t)children.size();
    d.union_order = union_order;
    d.selection_order = selection_order;
    d.window = window;
    d.max_window = max_window;
    for (size_t i = 0; i < children.size() && i < SFGPU_UNION_MAX_CHILDREN; ++i) d.children[i] = children[i];
    return d;
  }
  // the device-enumerable rows of the default list policy table (default_local_search/policy/list.rs:24-33)
  static sfgpu_union_desc default_list_union(uint32_t max_nearby = 20) {
    return union_desc({{SFGPU_FAM_NEARBY_LIST_CHANGE, max_nearby, 0, 0, 1}, {SFGPU_FAM_NEARBY_LIST_SWAP, max_nearby, 0, 0, 1},
                       {SFGPU_FAM_SUBLIST_CHANGE, 1, 3, 0, 1}, {SFGPU_FAM_SUBLIST_SWAP, 1, 3, 0, 1}, {SFGPU_FAM_LIST_REVERSE, 0, 0, 0, 1}});
  }
  // the default selectors of a plain scalar model (default_local_search/policy/scalar.rs:64-108)
  static sfgpu_union_desc default_scalar_union() {
    return union_desc({{SFGPU_FAM_CHANGE, 0, 0, 0, 1}, {SFGPU_FAM_SWAP, 0, 0, 0, 1}});
  }
  // one step; winner_rows[R][8] = {family, child, row[4], child-local pull index, 0}; flags bit 0 = window overflow
  StepResult step_union(const sfgpu_union_desc& desc, StepParams p, const std::vector<uint64_t>& step_seeds,
                        const std::vector<uint64_t>& step_indices, const std::vector<HardSoftScore>& ref_pairs = {},
                        bool apply_winners = false, std::vector<uint32_t>* flags = nullptr) {
    sfgpu_forage_params fp{p.acceptor, p.random_ties ? 1 : 0, p.accepted_limit, 0};
    StepResult out = make_result(8);
    std::vector<uint32_t> fl(R_, 0);
    check(sfgpu_step_union(ctx_, 0, &desc, &fp, step_seeds.empty() ? nullptr : step_seeds.data(),
                           step_indices.empty() ? nullptr : step_indices.data(),
                           ref_pairs.empty() ? nullptr : reinterpret_cast<const int64_t*>(ref_pairs.data()), out.index.data(),
                           reinterpret_cast<int64_t*>(out.best.data()), out.evaluated.data(), out.winner_rows.data(), fl.data(),
                           apply_winners ? 1 : 0));
    if (flags) *flags = fl;
    return out;
  }
  // the device-resident loop over the union step (acceptor 6 = SimulatedAnnealing)
  SolveResult solve_union(const sfgpu_union_desc& desc, const SolveParams& p, std::vector<uint64_t>* window_overflows = nullptr) {
    sfgpu_solve_params sp{};
    sp.n_steps = p.n_steps;
    sp.acceptor = p.acceptor;
    sp.late_size = p.late_size;
    sp.tie_mode = p.random_ties ? 1 : 0;
    sp.accepted_limit = p.accepted_limit;
    sp.seed_base = p.seed_base;
    sp.restore_best = p.restore_best ? 1 : 0;
    sp.acceptor_real = p.acceptor_real;
    sp.step_count_limit = p.step_count_limit;
    SolveResult out;
    out.best.resize(R_);
    out.moves_evaluated.resize(R_);
    out.accepted_steps.resize(R_);
    std::vector<uint64_t> ovf(R_, 0);
    check(sfgpu_solve_union(ctx_, &desc, &sp, reinterpret_cast<int64_t*>(out.best.data()), out.moves_evaluated.data(),
                            out.accepted_steps.data(), ovf.data(), nullptr));
    if (window_overflows) *window_overflows = ovf;
    return out;
  }
  // pair filter / pair weight of SFGPU_K_JOIN_EXPR as a postfix column expression
  uint32_t add_expr(const std::vector<sfgpu_expr_op>& ops) {
    uint32_t out = 0;
    check(sfgpu_add_expr(ctx_, ops.data(), (uint32_t)ops.size(), &out));
    return out;
  }
  // working_solution() of the list variable: per replica (offsets[n_owners + 1], elems[capacity])
  void list_state(uint32_t n_owners, std::vector<uint32_t>& offsets, std::vector<uint32_t>& elems) {
    uint32_t cap = 0;
    check(sfgpu_list_capacity(ctx_, 0x80000000u, &cap));
    offsets.assign((size_t)R_ * (n_owners + 1), 0);
    elems.assign((size_t)R_ * cap, 0);
    check(sfgpu_get_list_state(ctx_, 0x80000000u, offsets.data(), elems.data()));
  }
  void apply(const std::vector<ScalarSwap>& one_per_replica, const uint8_t* mask = nullptr) {
    check(sfgpu_apply_swap(ctx_, 0, reinterpret_cast<const uint32_t*>(one_per_replica.data()), mask));
  }
  void apply(const std::vector<ListSwap>& one_per_replica, const uint8_t* mask = nullptr) {
    check(sfgpu_apply_list_swap(ctx_, 0, reinterpret_cast<const uint32_t*>(one_per_replica.data()), mask));
  }
  void apply(const std::vector<ScalarEdit>& one_per_replica, const uint8_t* mask = nullptr) {
    check(sfgpu_apply_change(ctx_, 0, reinterpret_cast<const uint32_t*>(one_per_replica.data()), mask));
  }
  void apply(const std::vector<ListChange>& one_per_replica, const uint8_t* mask = nullptr) {
    check(sfgpu_apply_list_change(ctx_, 0, reinterpret_cast<const uint32_t*>(one_per_replica.data()), mask));
  }

 private:
  StepResult make_result(uint32_t row_words) const {
    StepResult out;
    out.index.assign(R_, UINT32_MAX);
    out.best.resize(R_);
    out.evaluated.assign(R_, 0);
    out.winner_rows.assign((size_t)R_ * (row_words ? row_words : 1), UINT32_MAX);
    return out;
  }
  sfgpu_ctx* ctx_ = nullptr;
  uint32_t R_;
};

inline Terminal ProjectedStream::impact(int imp, Weight w) const {
  if (w.w.fn != SFGPU_W_CONST && !(w.w.fn == SFGPU_W_LINEAR && w.w.b == 0))
    throw GpuError(SFGPU_E_UNSUPPORTED, "projected row weights must be CONST or LINEAR (a * amount)");
  const size_t n = p.row_ptr.size() - 1;
  std::vector<int64_t> sums(n, 0);
  for (size_t e = 0; e < n; ++e)
    for (uint32_t j = p.row_ptr[e]; j < p.row_ptr[e + 1]; ++j)
      sums[e] += w.w.fn == SFGPU_W_CONST ? w.w.a : (p.amounts.empty() ? 1 : p.amounts[j]);
  const uint32_t col = d->add_column(collection, "projected_row_sum", sums);
  sfgpu_constraint_desc c{};
  c.kind = SFGPU_K_UNI;
  c.impact = imp;
  c.weight = w.w;
  c.weight.fn = SFGPU_W_LINEAR;
  c.weight.a = w.w.fn == SFGPU_W_CONST ? 1 : w.w.a;
  c.weight.b = 0;
  c.collection = collection;
  c.aux0 = col;
  c.aux1 = UINT32_MAX;
  c.p0 = 1;  // assigned entities only: unassigned ones emit no rows
  return {d, c};
}
inline Terminal ProjectedGroupedStream::impact(int imp, Weight w) const {
  sfgpu_constraint_desc c{};
  c.kind = SFGPU_K_PROJECT_GROUP;
  c.impact = imp;
  c.weight = w.w;
  c.collection = collection;
  c.aux0 = p.csr;
  c.aux1 = UINT32_MAX;
  if (sum) {
    const uint32_t nnz = p.row_ptr.back();
    std::vector<int64_t> amt(std::max<uint32_t>(nnz, 1), 0);
    for (uint32_t j = 0; j < nnz; ++j) amt[j] = p.amounts.empty() ? 1 : p.amounts[j];
    const uint32_t rows = d->add_collection("projected_rows", (uint32_t)amt.size(), -1);
    c.aux1 = d->add_column(rows, "amount", amt);
  }
  c.p0 = p.keys_per_value;
  return {d, c};
}

inline uint32_t Terminal::named(const std::string& name) {
  desc.name = name.c_str();
  uint32_t id;
  d->check(sfgpu_add_constraint(d->raw(), &desc, &id));
  return id;
}

// ---------------------------------------------------------------------------------------------
// Acceptor / forager replay over batched scores, in pull order (phase/candidates.rs:66-282).
// ---------------------------------------------------------------------------------------------
inline uint64_t splitmix64(uint64_t v) {
  v += 0x9E3779B97F4A7C15ull;
  v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
  v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
  return v ^ (v >> 31);
}
inline bool reservoir_pick(uint64_t step_seed, uint64_t equal_count) {  // forager.rs:143-148
  return splitmix64(step_seed ^ (equal_count * 0x9E3779B97F4A7C15ull) ^ 0xF04A63E239B74D11ull) % equal_count == 0;
}

// Tabu metadata of one candidate (heuristic/move/metadata.rs:12-107). Ids are compared for equality
// only, so the raw indices stand in for the reference's SipHash-ed names and Debug strings.
struct MoveTabuSignature {
  uint64_t scope = 0;  // (descriptor_index, variable)
  std::vector<uint64_t> entity_ids, destination_value_ids, move_id, undo_move_id;
};
constexpr uint64_t TABU_NONE_ID = UINT64_MAX;
// ChangeMove::tabu_signature (change.rs:189-220); from/to = value index, -1 = None
inline MoveTabuSignature change_signature(uint64_t descriptor, uint64_t entity, int64_t from, int64_t to) {
  MoveTabuSignature g;
  const uint64_t f = from < 0 ? TABU_NONE_ID : (uint64_t)from, t = to < 0 ? TABU_NONE_ID : (uint64_t)to;
  g.scope = descriptor << 32;
  g.move_id = {descriptor, 0, entity, f, t};
  g.undo_move_id = {descriptor, 0, entity, t, f};
  g.entity_ids = {entity};
  g.destination_value_ids = {t};
  return g;
}
// change_tabu_signature (list_kernel/change.rs:155-204); moved = the element at (src_e, src_p)
inline MoveTabuSignature list_change_signature(uint64_t descriptor, uint64_t se, uint64_t sp, uint64_t de, uint64_t dp,
                                               uint64_t moved) {
  MoveTabuSignature g;
  const uint64_t adj = (se == de && dp > sp) ? dp - 1 : dp;
  g.scope = descriptor << 32;
  g.move_id = {descriptor, 0, se, sp, de, adj, moved};
  g.undo_move_id = {descriptor, 0, de, adj, se, sp, moved};
  g.entity_ids = {se};
  if (se != de) g.entity_ids.push_back(de);
  g.destination_value_ids = {moved};
  return g;
}

// How one step of an acceptor is expressed for the fused device steps (sfgpu_forage_params.acceptor
// + ref_scores {last_step, threshold}); acceptors that need per-move metadata or a random stream are
// not expressible and replay on the host over the materialised scores.
struct DeviceAcceptance {
  bool expressible = false;
  uint32_t acceptor = 0;  // 0 accept all, 1 move > last, 2 move >= last || move >= threshold,
                          // 3 move > last || move >= threshold
  HardSoftScore threshold;
};

struct Acceptor {  // acceptor/traits.rs:13-43
  virtual ~Acceptor() = default;
  virtual bool requires_move_signatures() const { return false; }
  virtual void phase_started(HardSoftScore) {}
  virtual void phase_ended() {}
  virtual bool is_accepted(HardSoftScore last_step, HardSoftScore move, const MoveTabuSignature* sig = nullptr) = 0;
  virtual void step_ended(HardSoftScore, const MoveTabuSignature* accepted = nullptr) { (void)accepted; }
  virtual DeviceAcceptance device_form(HardSoftScore /*last_step*/) const { return {}; }
};
struct HillClimbingAcceptor : Acceptor {  // hill_climbing.rs:33-42
  bool is_accepted(HardSoftScore last, HardSoftScore mv, const MoveTabuSignature* = nullptr) override {
    return mv > last;
  }
  DeviceAcceptance device_form(HardSoftScore last) const override { return {true, 1, last}; }
};
struct LateAcceptanceAcceptor : Acceptor {  // late_acceptance.rs:89-126
  size_t size;
  std::vector<HardSoftScore> history;
  size_t idx = 0;
  explicit LateAcceptanceAcceptor(size_t n = 400) : size(n) {}
  void phase_started(HardSoftScore initial) override {
    history.assign(size, initial);
    idx = 0;
  }
  bool is_accepted(HardSoftScore last, HardSoftScore mv, const MoveTabuSignature* = nullptr) override {
    return mv >= last || mv >= history[idx];
  }
  void step_ended(HardSoftScore step, const MoveTabuSignature* = nullptr) override {
    history[idx] = step;
    idx = (idx + 1) % size;
  }
  DeviceAcceptance device_form(HardSoftScore) const override { return {true, 2, history[idx]}; }
};
struct GreatDelugeAcceptor : Acceptor {  // great_deluge.rs:52-101
  double rain_speed;
  bool has_water = false;
  HardSoftScore water_level, initial_abs;
  explicit GreatDelugeAcceptor(double rain = 0.001) : rain_speed(rain) {}
  void phase_started(HardSoftScore initial) override {
    has_water = true;
    water_level = initial;
    initial_abs = initial.abs();
  }
  void phase_ended() override { has_water = false; }
  bool is_accepted(HardSoftScore last, HardSoftScore mv, const MoveTabuSignature* = nullptr) override {
    return mv > last || !has_water || mv >= water_level;
  }
  void step_ended(HardSoftScore, const MoveTabuSignature* = nullptr) override {
    if (has_water) water_level = water_level + initial_abs.multiply(rain_speed);
  }
  DeviceAcceptance device_form(HardSoftScore) const override {
    return has_water ? DeviceAcceptance{true, 3, water_level} : DeviceAcceptance{true, 0, {}};
  }
};
struct StepCountingHillClimbingAcceptor : Acceptor {  // step_counting.rs:55-104
  uint64_t step_count_limit, steps_since_improvement = 0;
  bool has_best = false;
  HardSoftScore best;
  explicit StepCountingHillClimbingAcceptor(uint64_t limit = 100) : step_count_limit(limit) {}
  void phase_started(HardSoftScore initial) override {
    has_best = true;
    best = initial;
    steps_since_improvement = 0;
  }
  void phase_ended() override {
    has_best = false;
    steps_since_improvement = 0;
  }
  bool is_accepted(HardSoftScore last, HardSoftScore mv, const MoveTabuSignature* = nullptr) override {
    return mv > last || steps_since_improvement < step_count_limit;
  }
  void step_ended(HardSoftScore step, const MoveTabuSignature* = nullptr) override {
    if (!has_best || step > best) {
      has_best = true;
      best = step;
      steps_since_improvement = 0;
    } else {
      steps_since_improvement++;
    }
  }
  DeviceAcceptance device_form(HardSoftScore) const override {
    // "> last || anything" while under the limit, "> last" afterwards: predicate 3 with a -inf / +inf threshold
    // (one predicate code for every replica of a batch)
    const int64_t inf = 1ll << 62;
    return {true, 3u, steps_since_improvement < step_count_limit ? HardSoftScore{-inf, -inf} : HardSoftScore{inf, inf}};
  }
};
struct DiversifiedLateAcceptanceAcceptor : Acceptor {  // diversified_late_acceptance.rs:71-140
  size_t size;
  double tolerance;
  std::vector<HardSoftScore> history;
  size_t idx = 0;
  bool has_best = false;
  HardSoftScore best;
  explicit DiversifiedLateAcceptanceAcceptor(size_t n = 400, double tol = 0.01) : size(n), tolerance(tol) {}
  void phase_started(HardSoftScore initial) override {
    history.assign(size, initial);
    idx = 0;
    has_best = true;
    best = initial;
  }
  HardSoftScore floor_score() const {  // the weaker of the late score and best - |best| * tolerance
    HardSoftScore t = history[idx];
    if (has_best) {
      const HardSoftScore d = best - best.abs().multiply(tolerance);
      if (d < t) t = d;
    }
    return t;
  }
  bool is_accepted(HardSoftScore last, HardSoftScore mv, const MoveTabuSignature* = nullptr) override {
    if (mv >= last || mv >= history[idx]) return true;
    return has_best && mv >= best - best.abs().multiply(tolerance);
  }
  void step_ended(HardSoftScore step, const MoveTabuSignature* = nullptr) override {
    if (!has_best || step > best) {
      has_best = true;
      best = step;
    }
    history[idx] = step;
    idx = (idx + 1) % size;
  }
  DeviceAcceptance device_form(HardSoftScore) const override { return {true, 2, floor_score()}; }
};
struct TabuSearchAcceptor : Acceptor {  // tabu_search.rs:103-237; tenure 0 = dimension off (None)
  template <class T>
  struct Memory {  // TabuMemory :64-101
    size_t tenure = 0;
    std::vector<T> entries;
    bool contains(const T& e) const { return std::find(entries.begin(), entries.end(), e) != entries.end(); }
    void record(const T& e) {
      if (!tenure) return;
      if (entries.size() >= tenure) entries.erase(entries.begin());
      entries.push_back(e);
    }
  };
  Memory<std::pair<uint64_t, uint64_t>> entity_memory, value_memory;
  Memory<std::vector<uint64_t>> move_memory, reverse_move_memory;
  bool aspiration_enabled;
  bool has_best = false;
  HardSoftScore best;
  TabuSearchAcceptor(size_t entity_tabu, size_t value_tabu, size_t move_tabu, size_t undo_move_tabu,
                     bool aspiration = true)
      : aspiration_enabled(aspiration) {
    if (!entity_tabu && !value_tabu && !move_tabu && !undo_move_tabu)
      throw std::invalid_argument("tabu_search requires at least one tabu dimension");  // :35-41
    entity_memory.tenure = entity_tabu;
    value_memory.tenure = value_tabu;
    move_memory.tenure = move_tabu;
    reverse_move_memory.tenure = undo_move_tabu;
  }
  bool requires_move_signatures() const override { return true; }
  void clear() {
    entity_memory.entries.clear();
    value_memory.entries.clear();
    move_memory.entries.clear();
    reverse_move_memory.entries.clear();
  }
  void phase_started(HardSoftScore initial) override {
    clear();
    has_best = true;
    best = initial;
  }
  void phase_ended() override {
    clear();
    has_best = false;
  }
  bool is_tabu(const MoveTabuSignature& g) const {
    for (uint64_t e : g.entity_ids)
      if (entity_memory.contains({g.scope, e})) return true;
    for (uint64_t v : g.destination_value_ids)
      if (value_memory.contains({g.scope, v})) return true;
    return move_memory.contains(g.move_id) || reverse_move_memory.contains(g.move_id);
  }
  bool is_accepted(HardSoftScore, HardSoftScore mv, const MoveTabuSignature* sig = nullptr) override {
    if (!sig) throw std::logic_error("tabu search requires move signatures");
    const bool aspirational = has_best && aspiration_enabled && mv > best;
    return aspirational || !is_tabu(*sig);
  }
  void step_ended(HardSoftScore step, const MoveTabuSignature* accepted = nullptr) override {
    if (accepted) {
      for (uint64_t e : accepted->entity_ids) entity_memory.record({accepted->scope, e});
      for (uint64_t v : accepted->destination_value_ids) value_memory.record({accepted->scope, v});
      move_memory.record(accepted->move_id);
      reverse_move_memory.record(accepted->undo_move_id);
    }
    if (!has_best || step > best) {
      has_best = true;
      best = step;
    }
  }
};
struct SimulatedAnnealingAcceptor : Acceptor {  // simulated_annealing.rs:338-431 (uniform stream injected)
  std::function<double()> uniform;
  double decay = 0.999985, hill_climbing_temperature = 1e-9, fallback = 1.0, target = 0.80;
  size_t calibration_samples = 128, seen = 0;
  double sum[2] = {0, 0}, temperature[2] = {0, 0};
  size_t count[2] = {0, 0};
  bool calibrated = false;
  explicit SimulatedAnnealingAcceptor(std::function<double()> u) : uniform(std::move(u)) {}
  bool is_accepted(HardSoftScore last, HardSoftScore mv, const MoveTabuSignature* = nullptr) override {
    if (mv >= last) return true;
    int lvl = mv.hard != last.hard ? 0 : 1;
    double delta = (double)(mv.level(lvl) - last.level(lvl));
    if (!calibrated) {
      sum[lvl] += std::fabs(delta);
      count[lvl]++;
      if (++seen < calibration_samples) return false;
      for (int l = 0; l < 2; ++l)
        temperature[l] = count[l] ? std::max((sum[l] / (double)count[l]) / -std::log(target), fallback) : fallback;
      calibrated = true;
    }
    if (temperature[lvl] <= hill_climbing_temperature) return false;
    return uniform() < std::exp(delta / temperature[lvl]);
  }
  void step_ended(HardSoftScore, const MoveTabuSignature* = nullptr) override {
    if (calibrated)
      for (double& t : temperature) t = std::max(t * decay, hill_climbing_temperature);
  }
};

struct ForagerConfig {
  // forager.rs:167-425 + forager/improving.rs:17-227
  enum Kind { AcceptedCount, FirstAccepted, BestScore, FirstBestScoreImproving, FirstLastStepScoreImproving } kind = BestScore;
  size_t accepted_count_limit = 1;  // AcceptedCount(N); optional horizon of FirstLastStepScoreImproving
  bool has_improving_limit = false;
  bool random_ties = true;
};
struct StepOutcome {
  bool has_winner = false;
  size_t winner = 0;  // CandidateId == pull index (move_selector/borrowed.rs:396-430)
  HardSoftScore score;
  uint64_t moves_evaluated = 0, score_calculations = 0, moves_accepted = 0;
};

// Replays phase/candidates.rs:66-282 over batched scores: pull order, quit-early, acceptor, forager.
// `signature(i)` supplies the tabu metadata of candidate i for acceptors that require it; gates[i] bit 0 =
// Move::requires_hard_improvement, bit 1 = Move::requires_score_improvement (may be null).
inline StepOutcome replay_step(const HardSoftScore* scores, const uint8_t* doable, size_t n, HardSoftScore best_score,
                               HardSoftScore last_step, uint64_t step_seed, const ForagerConfig& fc,
                               Acceptor& acceptor,
                               const std::function<MoveTabuSignature(size_t)>& signature = nullptr,
                               const uint8_t* gates = nullptr) {
  if (acceptor.requires_move_signatures() && !signature)
    throw std::logic_error("this acceptor requires move signatures");
  StepOutcome out;
  uint64_t equal_count = 0;
  size_t accepted = 0;
  bool found_improving = false;
  auto consider = [&](size_t i) {  // BestCandidate::consider, forager.rs:99-141
    if (!out.has_winner || scores[i] > out.score) {
      out.has_winner = true;
      out.winner = i;
      out.score = scores[i];
      equal_count = 1;
    } else if (scores[i] == out.score) {
      equal_count++;
      if (fc.random_ties && reservoir_pick(step_seed, equal_count)) out.winner = i;
    }
  };
  auto replace = [&](size_t i) {
    out.has_winner = true;
    out.winner = i;
    out.score = scores[i];
    equal_count = 1;
  };
  for (size_t i = 0; i < n; ++i) {
    bool quit = false;
    switch (fc.kind) {
      case ForagerConfig::BestScore: quit = false; break;
      case ForagerConfig::FirstAccepted: quit = out.has_winner; break;
      case ForagerConfig::AcceptedCount: quit = accepted >= fc.accepted_count_limit; break;
      case ForagerConfig::FirstBestScoreImproving: quit = found_improving; break;
      case ForagerConfig::FirstLastStepScoreImproving:
        quit = found_improving || (fc.has_improving_limit && accepted >= fc.accepted_count_limit);
        break;
    }
    if (quit) break;
    out.moves_evaluated++;
    if (!doable[i]) continue;
    out.score_calculations++;
    if (gates) {  // evaluation.rs:76-111: improvement gates run before the acceptor sees the candidate
      if ((gates[i] & 1) && !(scores[i].hard > last_step.hard)) continue;  // requires_hard_improvement
      if ((gates[i] & 2) && !(scores[i] > last_step)) continue;            // requires_score_improvement
    }
    if (signature) {
      const MoveTabuSignature g = signature(i);
      if (!acceptor.is_accepted(last_step, scores[i], &g)) continue;
    } else if (!acceptor.is_accepted(last_step, scores[i])) {
      continue;
    }
    out.moves_accepted++;
    switch (fc.kind) {
      case ForagerConfig::FirstAccepted:
        if (!out.has_winner) replace(i);
        break;
      case ForagerConfig::AcceptedCount:
        accepted++;
        consider(i);
        break;
      case ForagerConfig::BestScore: consider(i); break;
      case ForagerConfig::FirstBestScoreImproving:  // improving.rs:86-95
        if (scores[i] > best_score) {
          found_improving = true;
          replace(i);
        } else {
          consider(i);
        }
        break;
      case ForagerConfig::FirstLastStepScoreImproving:  // improving.rs:197-211
        accepted++;
        if (scores[i] > last_step) {
          found_improving = true;
          replace(i);
        } else {
          consider(i);
        }
        break;
    }
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
// Local-search phase over R replicas with one host-side acceptor per replica and the whole step on the
// device: solve_local_search_with_resources / execute_step (phase/localsearch/phase.rs:237-320,
// phase/step.rs:30-225) — phase_started, then per step: reference scores from the acceptor state ->
// neighbourhood + scoring + forager + commit on the GPU -> acceptor.step_ended(step score) -> best score.
// Acceptors must reduce to a device predicate per step (device_form); tabu / simulated annealing use
// replay_step over materialised scores instead.
class LocalSearch {
 public:
  enum Neighbourhood { Change, NearbyListChange, NearbyListSwap };
  LocalSearch(GpuScoreDirector& d, const std::function<std::unique_ptr<Acceptor>()>& make_acceptor, StepParams forager)
      : d_(d), forager_(forager) {
    for (uint32_t r = 0; r < d.replicas(); ++r) acceptors_.push_back(make_acceptor());
  }
  void phase_started() {
    best_ = d_.calculate_score();
    for (uint32_t r = 0; r < d_.replicas(); ++r) acceptors_[r]->phase_started(best_[r]);
  }
  void phase_ended() {
    for (auto& a : acceptors_) a->phase_ended();
  }
  // step seeds: one per replica (the reference draws them from the solver's StdRng, phase/step.rs:60-64)
  StepResult step(Neighbourhood n, const std::vector<uint64_t>& step_seeds, uint32_t max_nearby = 20) {
    const std::vector<HardSoftScore> last = d_.calculate_score();
    std::vector<HardSoftScore> refs(2 * last.size());
    StepParams p = forager_;
    for (size_t r = 0; r < last.size(); ++r) {
      const DeviceAcceptance f = acceptors_[r]->device_form(last[r]);
      if (!f.expressible) throw GpuError(SFGPU_E_UNSUPPORTED, "acceptor has no per-step device form: use replay_step");
      if (r > 0 && (int)f.acceptor != p.acceptor)
        throw GpuError(SFGPU_E_UNSUPPORTED, "replicas of one batch must share the acceptor predicate");
      p.acceptor = (int)f.acceptor;
      refs[2 * r] = last[r];
      refs[2 * r + 1] = f.threshold;
    }
    StepResult res = n == Change ? d_.step_change(p, step_seeds, refs, true)
                                 : d_.step_nearby_list_change(max_nearby, p, step_seeds, refs, true, n == NearbyListSwap);
    const std::vector<HardSoftScore> now = d_.calculate_score();
    for (size_t r = 0; r < now.size(); ++r) {
      acceptors_[r]->step_ended(now[r]);
      if (now[r] > best_[r]) best_[r] = now[r];
    }
    return res;
  }
  const std::vector<HardSoftScore>& best_scores() const { return best_; }
  Acceptor& acceptor(uint32_t r) { return *acceptors_[r]; }

 private:
  GpuScoreDirector& d_;
  StepParams forager_;
  std::vector<std::unique_ptr<Acceptor>> acceptors_;
  std::vector<HardSoftScore> best_;
};

}  // namespace sf
