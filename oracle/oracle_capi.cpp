// ORACLE — TEST INFRASTRUCTURE ONLY (see score.hpp header).
//
// C ABI over the oracle so tests/ and bench.py's cpu_baseline / --impl reference legs can drive
// it through ctypes. Each score call performs exactly the reference's evaluate_candidate
// (phase/localsearch/evaluation.rs:20-115): is_doable -> snapshot -> do_move -> calculate_score
// -> undo_move -> restore, through the retained incremental constraint state.
#include <cstring>

#include "models.hpp"

using namespace sfo;

namespace {
inline OptVal opt(int32_t v) { return v < 0 ? std::nullopt : OptVal((size_t)v); }

template <class MakeMove>
int score_batch(void* h, uint64_t n, MakeMove mk, int64_t* hard, int64_t* soft, uint8_t* doable) {
  auto* m = static_cast<OracleModel*>(h);
  m->calculate_score();
  for (uint64_t i = 0; i < n; ++i) {
    auto ev = m->evaluate(mk(i));
    bool ok = ev.kind != EvalKind::NotDoable;
    if (doable) doable[i] = ok ? 1 : 0;
    hard[i] = ok ? ev.score.hard : 0;
    soft[i] = ok ? ev.score.soft : 0;
  }
  return 0;
}
}  // namespace

extern "C" {

void* sfo_gc_create(uint32_t n, uint32_t k, const uint32_t* row_ptr, const uint32_t* col, const int32_t* color) {
  GraphColoring g;
  g.n_colors = k;
  g.nodes.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    g.nodes[i].id = i;
    g.nodes[i].color_idx = opt(color[i]);
    for (uint32_t j = row_ptr[i]; j < row_ptr[i + 1]; ++j) g.nodes[i].neighbors.push_back(col[j]);
  }
  return new GraphColoringModel(std::move(g));
}

// task clustering: joins = n_joins (arity, weight) pairs
void* sfo_cluster_create(uint32_t n, uint32_t n_teams, const int32_t* team, uint32_t n_joins, const int64_t* joins) {
  ClusterPlan p;
  p.n_teams = n_teams;
  for (uint32_t i = 0; i < n; ++i) p.tasks.push_back({i, opt(team[i])});
  std::vector<std::pair<size_t, int64_t>> js;
  for (uint32_t j = 0; j < n_joins; ++j) js.push_back({(size_t)joins[2 * j], joins[2 * j + 1]});
  return new ClusterModel(std::move(p), js);
}

void* sfo_nq_create(uint32_t n, const int32_t* row) {
  Board b;
  b.n_rows = n;
  for (uint32_t i = 0; i < n; ++i) b.queens.push_back({i, i, opt(row[i])});
  return new NQueensModel(std::move(b));
}

void* sfo_cvrp_create(uint32_t dim, uint32_t n_routes, int64_t capacity, uint32_t depot, const int32_t* demands,
                      const int64_t* matrix, const uint32_t* offsets, const uint32_t* elems) {
  auto pd = std::make_shared<ProblemData>();
  pd->capacity = capacity;
  pd->depot = depot;
  pd->demands.assign(demands, demands + dim);
  pd->distance_matrix.resize(dim);
  for (uint32_t i = 0; i < dim; ++i) pd->distance_matrix[i].assign(matrix + (size_t)i * dim, matrix + (size_t)(i + 1) * dim);
  CvrpPlan p;
  p.shared = pd;
  for (uint32_t i = 0; i < dim; ++i)
    if (i != depot) p.customers.push_back({i});
  for (uint32_t r = 0; r < n_routes; ++r) {
    Route rt{r, {}, pd.get()};
    for (uint32_t j = offsets[r]; j < offsets[r + 1]; ++j) rt.visits.push_back(elems[j]);
    p.routes.push_back(std::move(rt));
  }
  return new CvrpModel(std::move(p));
}

void* sfo_js_create(uint32_t n_ops, uint32_t n_machines, const uint32_t* job, const uint32_t* step,
                    const int32_t* machine_idx, const uint32_t* seq_offsets, const uint32_t* seq_elems,
                    int with_complement) {
  JobShopPlan p;
  for (uint32_t m = 0; m < n_machines; ++m) p.machines.push_back({m});
  for (uint32_t i = 0; i < n_ops; ++i) p.operations.push_back({i, job[i], step[i], opt(machine_idx[i])});
  for (uint32_t m = 0; m < n_machines; ++m) {
    MachineSequence s{m, {}};
    for (uint32_t j = seq_offsets[m]; j < seq_offsets[m + 1]; ++j) s.operations.push_back(seq_elems[j]);
    p.machine_sequences.push_back(std::move(s));
  }
  return new JobShopModel(std::move(p), with_complement != 0);
}

void* sfo_shift_create(uint32_t n_shifts, uint32_t n_nurses, const int64_t* day, const uint32_t* slot,
                       const uint8_t* required, const int64_t* hours, const int32_t* nurse_idx, int64_t target,
                       int with_load_balance, int64_t presence_days) {
  ShiftSchedule s;
  for (uint32_t i = 0; i < n_nurses; ++i) s.nurses.push_back({i});
  for (uint32_t i = 0; i < n_shifts; ++i)
    s.shifts.push_back({i, day[i], slot[i], required[i] != 0, hours[i], opt(nurse_idx[i])});
  return new ShiftModel(std::move(s), target, with_load_balance != 0, presence_days);
}

// roster with projected rows: span rows of shift i are span_day/span_hours[span_ptr[i] .. span_ptr[i+1])
void* sfo_roster_create(uint32_t n_shifts, uint32_t n_nurses, int64_t n_days, int64_t limit, const uint8_t* required,
                        const uint32_t* span_ptr, const int64_t* span_day, const int64_t* span_hours,
                        const int32_t* nurse_idx) {
  Roster s;
  s.n_nurses = n_nurses;
  s.n_days = n_days;
  for (uint32_t i = 0; i < n_shifts; ++i) {
    RShift sh{i, required[i] != 0, {}, opt(nurse_idx[i])};
    for (uint32_t j = span_ptr[i]; j < span_ptr[i + 1]; ++j) sh.spans.push_back({span_day[j], span_hours[j]});
    s.shifts.push_back(std::move(sh));
  }
  return new RosterModel(std::move(s), limit);
}

// unavailable: CSR over employees (row_ptr[n_emp + 1], days); contracts: [n][4] = {employee, from, to, fee}
void* sfo_availability_create(uint32_t n_shifts, uint32_t n_emp, const int64_t* day, const int64_t* required,
                              const int64_t* hours, const int32_t* employee, const int64_t* skill, const uint32_t* un_ptr,
                              const int64_t* un_days, uint32_t n_contracts, const int64_t* contracts) {
  AvSchedule s;
  for (uint32_t i = 0; i < n_shifts; ++i) s.shifts.push_back({i, day[i], required[i], hours[i], opt(employee[i])});
  for (uint32_t e = 0; e < n_emp; ++e) {
    AvEmployee emp{e, skill[e], {}};
    for (uint32_t j = un_ptr[e]; j < un_ptr[e + 1]; ++j) emp.unavailable_days.push_back(un_days[j]);
    s.employees.push_back(std::move(emp));
  }
  for (uint32_t k = 0; k < n_contracts; ++k)
    s.contracts.push_back({(size_t)contracts[4 * k], contracts[4 * k + 1], contracts[4 * k + 2], contracts[4 * k + 3]});
  return new AvailabilityModel(std::move(s));
}

void* sfo_pairs_create(uint32_t n, uint32_t n_buckets, const int64_t* demand, const int64_t* prio, const int32_t* bucket) {
  PwPlan p;
  p.n_buckets = n_buckets;
  for (uint32_t i = 0; i < n; ++i) p.work.push_back({i, demand[i], prio[i], opt(bucket[i])});
  return new PairsModel(std::move(p));
}

void sfo_destroy(void* h) { delete static_cast<OracleModel*>(h); }

int sfo_committed_score(void* h, int64_t out[2]) {
  Sc s = static_cast<OracleModel*>(h)->calculate_score();
  out[0] = s.hard;
  out[1] = s.soft;
  return 0;
}
int sfo_evaluate_all(void* h, int64_t out[2]) {
  Sc s = static_cast<OracleModel*>(h)->fresh_score();
  out[0] = s.hard;
  out[1] = s.soft;
  return 0;
}
uint64_t sfo_score_calculations(void* h) { return static_cast<OracleModel*>(h)->score_calculations(); }

int sfo_score_change(void* h, uint64_t n, const uint32_t* e, const int32_t* v, int64_t* hard, int64_t* soft,
                     uint8_t* doable) {
  size_t d = static_cast<OracleModel*>(h)->scalar_desc();
  return score_batch(h, n, [&](uint64_t i) { return Move::change(d, e[i], opt(v[i])); }, hard, soft, doable);
}
int sfo_score_swap(void* h, uint64_t n, const uint32_t* l, const uint32_t* r, int64_t* hard, int64_t* soft,
                   uint8_t* doable) {
  size_t d = static_cast<OracleModel*>(h)->scalar_desc();
  return score_batch(h, n, [&](uint64_t i) { return Move::swap(d, l[i], r[i]); }, hard, soft, doable);
}
int sfo_score_compound(void* h, uint64_t n, const uint32_t* offs, const uint32_t* e, const int32_t* v, int64_t* hard,
                       int64_t* soft, uint8_t* doable) {
  size_t d = static_cast<OracleModel*>(h)->scalar_desc();
  return score_batch(
      h, n,
      [&](uint64_t i) {
        std::vector<ScalarEdit> edits;
        for (uint32_t j = offs[i]; j < offs[i + 1]; ++j) edits.push_back({d, e[j], opt(v[j])});
        return Move::compound(std::move(edits));
      },
      hard, soft, doable);
}
int sfo_score_list_change(void* h, uint64_t n, const uint32_t* se, const uint32_t* sp, const uint32_t* de,
                          const uint32_t* dp, int64_t* hard, int64_t* soft, uint8_t* doable) {
  size_t d = static_cast<OracleModel*>(h)->list_desc();
  return score_batch(h, n, [&](uint64_t i) { return Move::list_change(d, se[i], sp[i], de[i], dp[i]); }, hard, soft,
                     doable);
}
int sfo_score_list_swap(void* h, uint64_t n, const uint32_t* e1, const uint32_t* p1, const uint32_t* e2,
                        const uint32_t* p2, int64_t* hard, int64_t* soft, uint8_t* doable) {
  size_t d = static_cast<OracleModel*>(h)->list_desc();
  return score_batch(h, n, [&](uint64_t i) { return Move::list_swap(d, e1[i], p1[i], e2[i], p2[i]); }, hard, soft,
                     doable);
}

// ListReverseMove rows {entity, start, end}: reverses [start, end) (heuristic/move/list_kernel/reverse.rs)
int sfo_score_list_reverse(void* h, uint64_t n, const uint32_t* e, const uint32_t* start, const uint32_t* end,
                           int64_t* hard, int64_t* soft, uint8_t* doable) {
  size_t d = static_cast<OracleModel*>(h)->list_desc();
  return score_batch(h, n, [&](uint64_t i) { return Move::list_reverse(d, e[i], start[i], end[i]); }, hard, soft, doable);
}
// SublistChangeMove rows {src_entity, start, end, dst_entity, dst_position} (heuristic/move/list_kernel/sublist_change.rs)
int sfo_score_sublist_change(void* h, uint64_t n, const uint32_t* se, const uint32_t* start, const uint32_t* end,
                             const uint32_t* de, const uint32_t* dp, int64_t* hard, int64_t* soft, uint8_t* doable) {
  size_t d = static_cast<OracleModel*>(h)->list_desc();
  return score_batch(h, n, [&](uint64_t i) { return move_sublist_change(d, se[i], start[i], end[i], de[i], dp[i]); },
                     hard, soft, doable);
}
int sfo_apply_sublist_change(void* h, uint32_t se, uint32_t start, uint32_t end, uint32_t de, uint32_t dp) {
  auto* m = static_cast<OracleModel*>(h);
  m->apply(move_sublist_change(m->list_desc(), se, start, end, de, dp));
  return 0;
}
// SublistSwapMove rows {first_entity, start1, end1, second_entity, start2, end2} (heuristic/move/list_kernel/sublist_swap.rs)
int sfo_score_sublist_swap(void* h, uint64_t n, const uint32_t* e1, const uint32_t* s1, const uint32_t* t1,
                           const uint32_t* e2, const uint32_t* s2, const uint32_t* t2, int64_t* hard, int64_t* soft,
                           uint8_t* doable) {
  size_t d = static_cast<OracleModel*>(h)->list_desc();
  return score_batch(h, n, [&](uint64_t i) { return move_sublist_swap(d, e1[i], s1[i], t1[i], e2[i], s2[i], t2[i]); },
                     hard, soft, doable);
}
int sfo_apply_sublist_swap(void* h, uint32_t e1, uint32_t s1, uint32_t t1, uint32_t e2, uint32_t s2, uint32_t t2) {
  auto* m = static_cast<OracleModel*>(h);
  m->apply(move_sublist_swap(m->list_desc(), e1, s1, t1, e2, s2, t2));
  return 0;
}
int sfo_apply_list_reverse(void* h, uint32_t e, uint32_t start, uint32_t end) {
  auto* m = static_cast<OracleModel*>(h);
  m->apply(Move::list_reverse(m->list_desc(), e, start, end));
  return 0;
}
int sfo_apply_change(void* h, uint32_t e, int32_t v) {
  auto* m = static_cast<OracleModel*>(h);
  m->apply(Move::change(m->scalar_desc(), e, opt(v)));
  return 0;
}
int sfo_apply_swap(void* h, uint32_t l, uint32_t r) {
  auto* m = static_cast<OracleModel*>(h);
  m->apply(Move::swap(m->scalar_desc(), l, r));
  return 0;
}
int sfo_apply_compound(void* h, uint32_t n_edits, const uint32_t* e, const int32_t* v) {
  auto* m = static_cast<OracleModel*>(h);
  std::vector<ScalarEdit> edits;
  for (uint32_t j = 0; j < n_edits; ++j) edits.push_back({m->scalar_desc(), e[j], opt(v[j])});
  m->apply(Move::compound(std::move(edits)));
  return 0;
}
int sfo_apply_list_change(void* h, uint32_t se, uint32_t sp, uint32_t de, uint32_t dp) {
  auto* m = static_cast<OracleModel*>(h);
  m->apply(Move::list_change(m->list_desc(), se, sp, de, dp));
  return 0;
}
int sfo_apply_list_swap(void* h, uint32_t e1, uint32_t p1, uint32_t e2, uint32_t p2) {
  auto* m = static_cast<OracleModel*>(h);
  m->apply(Move::list_swap(m->list_desc(), e1, p1, e2, p2));
  return 0;
}

static MoveStreamContext make_ctx(uint64_t step_index, uint64_t step_seed, int order) {
  MoveStreamContext c;
  c.step_index = step_index;
  c.step_seed = step_seed;
  c.order = order == 1 ? SelectionOrder::Random : order == 2 ? SelectionOrder::Shuffled : SelectionOrder::Original;
  return c;
}

// union pull order (vec_union.rs UnionScheduler) over children of the given sizes: writes (child, child-local
// index) pairs; returns the number of pulls (<= cap). union_order: 0 Sequential, 1 RoundRobin,
// 2 RotatingRoundRobin, 3 Random, 4 StratifiedRandom.
int64_t sfo_union_pull_order(uint32_t n_children, const uint64_t* sizes, const uint64_t* weights, int union_order,
                             uint64_t step_index, uint64_t step_seed, int order, uint64_t cap, uint32_t* out_child,
                             uint64_t* out_local) {
  std::vector<size_t> sz(sizes, sizes + n_children);
  std::vector<uint64_t> w(weights, weights + n_children);
  const UnionOrder uo = union_order == 1   ? UnionOrder::RoundRobin
                        : union_order == 2 ? UnionOrder::RotatingRoundRobin
                        : union_order == 3 ? UnionOrder::Random
                        : union_order == 4 ? UnionOrder::StratifiedRandom
                                           : UnionOrder::Sequential;
  auto pulls = union_pull_order(sz, uo, make_ctx(step_index, step_seed, order), w, (size_t)cap);
  for (size_t i = 0; i < pulls.size(); ++i) {
    out_child[i] = (uint32_t)pulls[i].first;
    out_local[i] = (uint64_t)pulls[i].second;
  }
  return (int64_t)pulls.size();
}

// Returns the number of candidates (may exceed cap; only the first cap are written).
int64_t sfo_enumerate_change(void* h, uint64_t step_index, uint64_t step_seed, int order, uint64_t cap, uint32_t* e,
                             int32_t* v) {
  auto moves = static_cast<OracleModel*>(h)->enumerate_scalar(make_ctx(step_index, step_seed, order));
  for (size_t i = 0; i < moves.size() && i < cap; ++i) {
    e[i] = (uint32_t)moves[i].a;
    v[i] = moves[i].to ? (int32_t)*moves[i].to : -1;
  }
  return (int64_t)moves.size();
}
int64_t sfo_enumerate_swap(void* h, uint64_t step_index, uint64_t step_seed, int order, uint64_t cap, uint32_t* l,
                           uint32_t* r) {
  auto moves = static_cast<OracleModel*>(h)->enumerate_scalar_swap(make_ctx(step_index, step_seed, order));
  for (size_t i = 0; i < moves.size() && i < cap; ++i) {
    l[i] = (uint32_t)moves[i].a;
    r[i] = (uint32_t)moves[i].b;
  }
  return (int64_t)moves.size();
}
int64_t sfo_enumerate_nearby_list_change(void* h, uint32_t max_nearby, uint64_t step_index, uint64_t step_seed,
                                         int order, uint64_t cap, uint32_t* se, uint32_t* sp, uint32_t* de,
                                         uint32_t* dp) {
  auto moves = static_cast<OracleModel*>(h)->enumerate_list(max_nearby, make_ctx(step_index, step_seed, order));
  for (size_t i = 0; i < moves.size() && i < cap; ++i) {
    se[i] = (uint32_t)moves[i].a;
    sp[i] = (uint32_t)moves[i].b;
    de[i] = (uint32_t)moves[i].c;
    dp[i] = (uint32_t)moves[i].d;
  }
  return (int64_t)moves.size();
}

int64_t sfo_enumerate_sublist_change(void* h, uint32_t min_size, uint32_t max_size, uint64_t step_index,
                                     uint64_t step_seed, int order, uint64_t cap, uint32_t* se, uint32_t* start,
                                     uint32_t* end, uint32_t* de, uint32_t* dp) {
  auto moves = static_cast<OracleModel*>(h)->enumerate_sublist_change(min_size, max_size, make_ctx(step_index, step_seed, order));
  for (size_t i = 0; i < moves.size() && i < cap; ++i) {
    se[i] = (uint32_t)moves[i].a;
    start[i] = (uint32_t)moves[i].b;
    end[i] = (uint32_t)moves[i].c;
    de[i] = (uint32_t)moves[i].d;
    dp[i] = (uint32_t)moves[i].e;
  }
  return (int64_t)moves.size();
}

int64_t sfo_enumerate_sublist_swap(void* h, uint32_t min_size, uint32_t max_size, uint64_t step_index, uint64_t step_seed,
                                   int order, uint64_t cap, uint32_t* e1, uint32_t* s1, uint32_t* t1, uint32_t* e2,
                                   uint32_t* s2, uint32_t* t2) {
  auto moves = static_cast<OracleModel*>(h)->enumerate_sublist_swap(min_size, max_size, make_ctx(step_index, step_seed, order));
  for (size_t i = 0; i < moves.size() && i < cap; ++i) {
    e1[i] = (uint32_t)moves[i].a;
    s1[i] = (uint32_t)moves[i].b;
    t1[i] = (uint32_t)moves[i].c;
    e2[i] = (uint32_t)moves[i].d;
    s2[i] = (uint32_t)moves[i].e;
    t2[i] = (uint32_t)moves[i].f;
  }
  return (int64_t)moves.size();
}

// KOptMove rows: k + 2 words each = {entity, cut_0 .. cut_{k-1}, pattern index in enumerate_reconnections(k)}
int64_t sfo_enumerate_k_opt(void* h, uint32_t k, uint32_t min_seg, uint64_t step_index, uint64_t step_seed, int order,
                            uint64_t cap, uint32_t* rows) {
  auto moves = static_cast<OracleModel*>(h)->enumerate_k_opt(k, min_seg, make_ctx(step_index, step_seed, order));
  for (size_t i = 0; i < moves.size() && i < cap; ++i) {
    uint32_t* r = rows + i * (k + 2);
    r[0] = (uint32_t)moves[i].a;
    for (uint32_t c = 0; c < k; ++c) r[1 + c] = (uint32_t)moves[i].cuts[c];
    r[k + 1] = (uint32_t)moves[i].f;
  }
  return (int64_t)moves.size();
}
// scores k-opt rows of the layout above (do / score / undo per candidate)
int sfo_score_k_opt(void* h, uint32_t k, uint64_t n, const uint32_t* rows, int64_t* hard, int64_t* soft, uint8_t* doable) {
  size_t d = static_cast<OracleModel*>(h)->list_desc();
  const auto patterns = enumerate_reconnections(k);
  return score_batch(h, n, [&](uint64_t i) {
    const uint32_t* r = rows + i * (k + 2);
    std::vector<size_t> cuts(r + 1, r + 1 + k);
    return move_k_opt(d, r[0], cuts, patterns[r[k + 1] % patterns.size()]);
  }, hard, soft, doable);
}

int sfo_apply_k_opt(void* h, uint32_t k, const uint32_t* row) {
  auto* m = static_cast<OracleModel*>(h);
  const auto patterns = enumerate_reconnections(k);
  std::vector<size_t> cuts(row + 1, row + 1 + k);
  m->apply(move_k_opt(m->list_desc(), row[0], cuts, patterns[row[k + 1] % patterns.size()]));
  return 0;
}

int64_t sfo_enumerate_list_reverse(void* h, uint64_t step_index, uint64_t step_seed, int order, uint64_t cap,
                                   uint32_t* e, uint32_t* start, uint32_t* end) {
  auto moves = static_cast<OracleModel*>(h)->enumerate_list_reverse(make_ctx(step_index, step_seed, order));
  for (size_t i = 0; i < moves.size() && i < cap; ++i) {
    e[i] = (uint32_t)moves[i].a;
    start[i] = (uint32_t)moves[i].b;
    end[i] = (uint32_t)moves[i].c;
  }
  return (int64_t)moves.size();
}

int64_t sfo_enumerate_nearby_list_swap(void* h, uint32_t max_nearby, uint64_t step_index, uint64_t step_seed,
                                       int order, uint64_t cap, uint32_t* se, uint32_t* sp, uint32_t* de,
                                       uint32_t* dp) {
  auto moves = static_cast<OracleModel*>(h)->enumerate_list_swap(max_nearby, make_ctx(step_index, step_seed, order));
  for (size_t i = 0; i < moves.size() && i < cap; ++i) {
    se[i] = (uint32_t)moves[i].a;
    sp[i] = (uint32_t)moves[i].b;
    de[i] = (uint32_t)moves[i].c;
    dp[i] = (uint32_t)moves[i].d;
  }
  return (int64_t)moves.size();
}

// Replay of the candidate loop (phase/candidates.rs:66-282) over precomputed evaluations.
// forager: 0 AcceptedCount(limit) 1 FirstAccepted 2 BestScore 3 FirstBestScoreImproving
//          4 FirstLastStepScoreImproving; acceptor: 0 HillClimbing 1 LateAcceptance(late_score) 3 AcceptAll.
// out[0]=has_winner out[1]=winner out[2]=moves_evaluated out[3]=score_calculations out[4]=moves_accepted
int sfo_replay_step(uint64_t n, const int64_t* hard, const int64_t* soft, const uint8_t* doable,
                    const int64_t best_score[2], const int64_t last_step_score[2], const int64_t late_score[2],
                    uint64_t step_seed, int forager_kind, uint64_t accepted_limit, int random_ties, int acceptor_kind,
                    uint64_t out[5]) {
  Forager<Sc> fg;
  fg.kind = (ForagerKind)forager_kind;
  fg.accepted_count_limit = accepted_limit;
  fg.best.random_ties = random_ties != 0;
  Acceptor<Sc> ac;
  ac.kind = (AcceptorKind)acceptor_kind;
  if (ac.kind == AcceptorKind::LateAcceptance) {
    ac.history.assign(1, Sc::of(late_score[0], late_score[1]));
    ac.history_idx = 0;
  }
  auto o = replay_step<Sc>(
      n,
      [&](size_t i) {
        return CandidateEvaluation<Sc>{doable[i] ? EvalKind::Scored : EvalKind::NotDoable, Sc::of(hard[i], soft[i])};
      },
      Sc::of(best_score[0], best_score[1]), Sc::of(last_step_score[0], last_step_score[1]), step_seed, fg, ac);
  out[0] = o.has_winner;
  out[1] = o.winner;
  out[2] = o.moves_evaluated;
  out[3] = o.score_calculations;
  out[4] = o.moves_accepted;
  return 0;
}

// sfo_replay_step with the improvement gates of evaluate_candidate (evaluation.rs:62-111): gates[i] bit 0 =
// requires_hard_improvement, bit 1 = requires_score_improvement, judged against last_step_score.
int sfo_replay_step_gated(uint64_t n, const int64_t* hard, const int64_t* soft, const uint8_t* doable,
                          const uint8_t* gates, const int64_t best_score[2], const int64_t last_step_score[2],
                          const int64_t late_score[2], uint64_t step_seed, int forager_kind, uint64_t accepted_limit,
                          int random_ties, int acceptor_kind, uint64_t out[5]) {
  Forager<Sc> fg;
  fg.kind = (ForagerKind)forager_kind;
  fg.accepted_count_limit = accepted_limit;
  fg.best.random_ties = random_ties != 0;
  Acceptor<Sc> ac;
  ac.kind = (AcceptorKind)acceptor_kind;
  if (ac.kind == AcceptorKind::LateAcceptance) {
    ac.history.assign(1, Sc::of(late_score[0], late_score[1]));
    ac.history_idx = 0;
  }
  const Sc last = Sc::of(last_step_score[0], last_step_score[1]);
  auto o = replay_step<Sc>(
      n,
      [&](size_t i) {
        const Sc sc = Sc::of(hard[i], soft[i]);
        if (!doable[i]) return CandidateEvaluation<Sc>{EvalKind::NotDoable, sc};
        if ((gates[i] & 1) && hard_score_delta(last, sc) != HardDelta::Improving)
          return CandidateEvaluation<Sc>{EvalKind::RejectedByHardImprovement, sc};
        if ((gates[i] & 2) && sc <= last) return CandidateEvaluation<Sc>{EvalKind::RejectedByScoreImprovement, sc};
        return CandidateEvaluation<Sc>{EvalKind::Scored, sc};
      },
      Sc::of(best_score[0], best_score[1]), last, step_seed, fg, ac);
  out[0] = o.has_winner;
  out[1] = o.winner;
  out[2] = o.moves_evaluated;
  out[3] = o.score_calculations;
  out[4] = o.moves_accepted;
  return 0;
}

// ---- stateful acceptors (acceptor/*.rs) for multi-step trajectories ---------------------------
// kind: AcceptorKind order (0 HillClimbing, 1 LateAcceptance, 3 AcceptAll, 4 GreatDeluge,
// 5 StepCountingHillClimbing, 6 DiversifiedLateAcceptance, 7 TabuSearch).
// size = late-acceptance size / step_count_limit; real = rain_speed / tolerance;
// tabu[4] = entity, value, move, undo-move tenures (0 = off); aspiration flag.
void* sfo_acceptor_create(int kind, uint64_t size, double real, const uint64_t tabu[4], int aspiration) {
  auto* a = new Acceptor<Sc>();
  a->kind = (AcceptorKind)kind;
  a->history.assign(size ? size : 1, Sc::zero());
  a->step_count_limit = size;
  a->rain_speed = real;
  a->tolerance = real;
  if (tabu) {
    a->entity_memory.tenure = tabu[0];
    a->value_memory.tenure = tabu[1];
    a->move_memory.tenure = tabu[2];
    a->reverse_move_memory.tenure = tabu[3];
  }
  a->aspiration_enabled = aspiration != 0;
  if (a->kind == AcceptorKind::SimulatedAnnealing) {  // size = calibration samples, real = decay rate (0 = defaults)
    if (size) a->calibration_samples = size;
    if (real > 0.0) a->decay = real;
    a->never_accept_hard_regression = aspiration == 2;
  }
  return a;
}
void sfo_acceptor_destroy(void* h) { delete (Acceptor<Sc>*)h; }
void sfo_acceptor_phase_started(void* h, const int64_t initial[2]) {
  auto* a = (Acceptor<Sc>*)h;
  a->phase_started(Sc::of(initial[0], initial[1]), a->history.size());
}
// tabu signatures of ListChangeMove rows against the handle's working solution, flattened:
// sig[i] = {scope, n_entities, e0, e1, value, move_id[7], undo_move_id[7]} (18 u64 per row)
static TabuSignature unpack_sig(const uint64_t* p) {
  TabuSignature t;
  t.scope = p[0];
  for (uint64_t k = 0; k < p[1]; ++k) t.entity_ids.push_back(p[2 + k]);
  t.value_ids = {p[4]};
  t.move_id.assign(p + 5, p + 12);
  t.undo_move_id.assign(p + 12, p + 19);
  return t;
}
// One step with the stateful acceptor: replay, then acceptor.step_ended(step_score, accepted signature)
// where step_score = winner's score, or last_step_score when nothing was picked (step.rs:122-221).
// sigs may be null (19 u64 per candidate otherwise). out as sfo_replay_step.
int sfo_acceptor_step(void* h, uint64_t n, const int64_t* hard, const int64_t* soft, const uint8_t* doable,
                      const uint64_t* sigs, const int64_t best_score[2], const int64_t last_step_score[2],
                      uint64_t step_seed, int forager_kind, uint64_t accepted_limit, int random_ties, uint64_t out[5]) {
  auto* a = (Acceptor<Sc>*)h;
  Forager<Sc> fg;
  fg.kind = (ForagerKind)forager_kind;
  fg.accepted_count_limit = accepted_limit;
  fg.best.random_ties = random_ties != 0;
  std::vector<TabuSignature> ts;
  if (sigs)
    for (uint64_t i = 0; i < n; ++i) ts.push_back(unpack_sig(sigs + 19 * i));
  auto eval = [&](size_t i) {
    return CandidateEvaluation<Sc>{doable[i] ? EvalKind::Scored : EvalKind::NotDoable, Sc::of(hard[i], soft[i])};
  };
  const Sc best = Sc::of(best_score[0], best_score[1]), last = Sc::of(last_step_score[0], last_step_score[1]);
  // SimulatedAnnealing: the uniform stream is injected (the reference's SmallRng is third party and unpinned):
  // draw j of a step = (splitmix64(step_seed ^ 0x5A17EA11EA1DF00D ^ j * 0x9E3779B97F4A7C15) >> 11) * 2^-53, the
  // stream include/sfgpu.h states for the device loop.
  uint64_t draw = 0;
  a->uniform = [&draw, step_seed]() {
    const uint64_t x = splitmix64(step_seed ^ 0x5A17EA11EA1DF00Dull ^ (draw * 0x9E3779B97F4A7C15ull));
    ++draw;
    return (double)(x >> 11) * 0x1.0p-53;
  };
  StepOutcome<Sc> o;
  try {
    if (sigs)
      o = replay_step<Sc>(n, eval, best, last, step_seed, fg, *a, [&](size_t i) { return &ts[i]; });
    else
      o = replay_step<Sc>(n, eval, best, last, step_seed, fg, *a);
  } catch (const std::exception&) {
    return -1;
  }
  a->step_ended(o.has_winner ? o.winner_score : last, (o.has_winner && sigs) ? &ts[o.winner] : nullptr);
  out[0] = o.has_winner;
  out[1] = o.winner;
  out[2] = o.moves_evaluated;
  out[3] = o.score_calculations;
  out[4] = o.moves_accepted;
  return 0;
}
// packs tabu_signature of n moves of the model behind `model` (19 u64 each, layout above).
// move_kind: 0 ChangeMove rows {entity, to_value(-1 = None)}, 2 ListChangeMove rows {se, sp, de, dp}
int sfo_move_signatures(void* model, int move_kind, uint64_t n, const uint32_t* rows, uint64_t* out_sigs) {
  auto* m = static_cast<OracleModel*>(model);
  for (uint64_t i = 0; i < n; ++i) {
    Move mv = move_kind == 0 ? Move::change(m->scalar_desc(), rows[2 * i], opt((int32_t)rows[2 * i + 1]))
                             : Move::list_change(m->list_desc(), rows[4 * i], rows[4 * i + 1], rows[4 * i + 2],
                                                 rows[4 * i + 3]);
    TabuSignature t = m->signature(mv);
    uint64_t* p = out_sigs + 19 * i;
    for (int k = 0; k < 19; ++k) p[k] = 0;
    p[0] = t.scope;
    p[1] = t.entity_ids.size();
    for (size_t k = 0; k < t.entity_ids.size() && k < 2; ++k) p[2 + k] = t.entity_ids[k];
    p[4] = t.value_ids.empty() ? TABU_NONE_ID : t.value_ids[0];
    for (size_t k = 0; k < t.move_id.size() && k < 7; ++k) p[5 + k] = t.move_id[k];
    for (size_t k = 0; k < t.undo_move_id.size() && k < 7; ++k) p[12 + k] = t.undo_move_id[k];
  }
  return 0;
}

}  // extern "C"
