// ORACLE — TEST INFRASTRUCTURE ONLY (see score.hpp header).
//
// The four BASELINE.json config models, authored against the oracle's restatement of the
// reference ConstraintStream API, closure for closure:
//   C1 n-queens        examples/nqueens/src/domain/board.rs:21-47
//   C2 graph colouring examples/scalar-graph-coloring/src/domain/graph_coloring.rs:21-44
//   C3 CVRP            constraint (i) verbatim from crates/solverforge/tests/list_clarke_wright_publication/
//                      domain/publication_plan.rs:51-65; (ii) capacity and (iii) distance are authored by us
//                      (the reference ships no CVRP constraints — SURVEY §0.1-5) over
//                      crates/solverforge-cvrp/src/problem_data.rs:14-47 semantics (distance_cost)
//   C4 mixed job-shop  examples/mixed-job-shop/src/domain/job_shop_plan.rs:28-69 plus one authored
//                      grouped-complement constraint
#pragma once
#include <memory>

#include "director.hpp"

namespace sfo {

using Sc = HardSoftScore;

// Type-erased handle used by the C API.
struct OracleModel {
  virtual ~OracleModel() = default;
  virtual Sc calculate_score() = 0;
  virtual Sc fresh_score() = 0;
  virtual CandidateEvaluation<Sc> evaluate(const Move& m) = 0;
  virtual void apply(const Move& m) = 0;
  virtual std::vector<Move> enumerate_scalar(MoveStreamContext ctx) { return {}; }
  virtual std::vector<Move> enumerate_scalar_swap(MoveStreamContext ctx) { return {}; }
  virtual std::vector<Move> enumerate_list(size_t max_nearby, MoveStreamContext ctx) { return {}; }
  virtual std::vector<Move> enumerate_list_swap(size_t max_nearby, MoveStreamContext ctx) { return {}; }
  virtual std::vector<Move> enumerate_list_reverse(MoveStreamContext ctx) { return {}; }
  virtual std::vector<Move> enumerate_sublist_change(size_t min_size, size_t max_size, MoveStreamContext ctx) { return {}; }
  virtual std::vector<Move> enumerate_sublist_swap(size_t min_size, size_t max_size, MoveStreamContext ctx) { return {}; }
  virtual std::vector<Move> enumerate_k_opt(size_t k, size_t min_seg, MoveStreamContext ctx) { return {}; }
  virtual size_t scalar_desc() const { return 0; }
  virtual size_t list_desc() const { return 0; }
  virtual uint64_t score_calculations() const = 0;
  virtual TabuSignature signature(const Move& m) = 0;
};

template <class S>
struct ModelImpl : OracleModel {
  ScoreDirector<S, Sc> dir;
  Sc calculate_score() override { return dir.calculate_score(); }
  Sc fresh_score() override { return dir.fresh_score(); }
  CandidateEvaluation<Sc> evaluate(const Move& m) override {
    return evaluate_candidate(m, dir, dir.calculate_score());
  }
  void apply(const Move& m) override {
    dir.calculate_score();
    do_move(m, dir);
  }
  uint64_t score_calculations() const override { return dir.score_calculations; }
  TabuSignature signature(const Move& m) override { return tabu_signature(m, dir); }
};

struct ConstKey {
  uint8_t operator()(...) const { return 0; }
};

// ------------------------------------------------------------------------------------------- C2
struct GcNode {
  size_t id;
  std::vector<size_t> neighbors;
  OptVal color_idx;
};
struct GraphColoring {
  std::vector<GcNode> nodes;
  size_t n_colors = 0;
};
inline const std::vector<GcNode>& gc_nodes(const GraphColoring& s) { return s.nodes; }

struct GraphColoringModel final : ModelImpl<GraphColoring> {
  GraphColoringModel(GraphColoring sol) {
    dir.working = std::move(sol);
    dir.access.get = [](const GraphColoring& s, size_t, size_t e) { return s.nodes[e].color_idx; };
    dir.access.set = [](GraphColoring& s, size_t, size_t e, OptVal v) { s.nodes[e].color_idx = v; };
    dir.access.entity_count = [](const GraphColoring& s, size_t) { return s.nodes.size(); };
    Source<GraphColoring, GcNode> src{gc_nodes, ChangeSource::Desc(0)};
    auto uf = [](const GraphColoring&, const GcNode& n) { return !n.color_idx.has_value(); };
    auto uw = [](const GcNode&) { return Sc::ONE_HARD(); };
    dir.constraints.add(std::make_unique<UniConstraint<GraphColoring, GcNode, Sc, decltype(uf), decltype(uw)>>(
        "Unassigned color", Impact::Penalty, src, uf, uw, true));
    auto pf = [](const GraphColoring&, const GcNode& l, const GcNode& r, size_t, size_t) {
      return l.id < r.id && std::find(l.neighbors.begin(), l.neighbors.end(), r.id) != l.neighbors.end() &&
             l.color_idx.has_value() && l.color_idx == r.color_idx;
    };
    auto pw = [](const GraphColoring&, const GcNode&, const GcNode&, size_t, size_t) { return Sc::ONE_HARD(); };
    dir.constraints.add(
        std::make_unique<CrossBiConstraint<GraphColoring, GcNode, GcNode, uint8_t, Sc, ConstKey, ConstKey,
                                           decltype(pf), decltype(pw)>>(
            "Adjacent color conflict", Impact::Penalty, src, src, ConstKey{}, ConstKey{}, pf, pw, true));
  }
  std::vector<Move> enumerate_scalar(MoveStreamContext ctx) override {
    return enumerate_change_moves(dir.working, dir.access, 0, 0, dir.working.n_colors, true, ctx);
  }
  std::vector<Move> enumerate_scalar_swap(MoveStreamContext ctx) override {
    return enumerate_swap_moves(dir.working, dir.access, 0, 0, ctx);
  }
};

// ------------------------------------------------------------------------- task clustering (authored)
// The shape of the reference's tri / quad / penta known-answer tests (constraint/tests/{tri,quad,penta}_incr.rs:
// tasks keyed by team) as a planning model: Task{team_idx: Option}, "Unassigned task" (1 hard) and one keyed
// self-join per requested arity: for_each(tasks).join(equal(team)).join(equal(team))[..] filtered to assigned
// tasks, penalised arity_weight soft per tuple.
struct ClusterTask {
  size_t id;
  OptVal team_idx;
};
struct ClusterPlan {
  std::vector<ClusterTask> tasks;
  size_t n_teams = 0;
};
inline const std::vector<ClusterTask>& cluster_tasks(const ClusterPlan& s) { return s.tasks; }

struct ClusterModel final : ModelImpl<ClusterPlan> {
  ClusterModel(ClusterPlan sol, const std::vector<std::pair<size_t, int64_t>>& joins /* (arity, soft weight) */) {
    dir.working = std::move(sol);
    dir.access.get = [](const ClusterPlan& s, size_t, size_t e) { return s.tasks[e].team_idx; };
    dir.access.set = [](ClusterPlan& s, size_t, size_t e, OptVal v) { s.tasks[e].team_idx = v; };
    dir.access.entity_count = [](const ClusterPlan& s, size_t) { return s.tasks.size(); };
    Source<ClusterPlan, ClusterTask> src{cluster_tasks, ChangeSource::Desc(0)};
    auto uf = [](const ClusterPlan&, const ClusterTask& t) { return !t.team_idx.has_value(); };
    auto uw = [](const ClusterTask&) { return Sc::ONE_HARD(); };
    dir.constraints.add(std::make_unique<UniConstraint<ClusterPlan, ClusterTask, Sc, decltype(uf), decltype(uw)>>(
        "Unassigned task", Impact::Penalty, src, uf, uw, true));
    for (auto& j : joins) {
      auto kf = [](const ClusterTask& t) { return t.team_idx.has_value() ? (int64_t)*t.team_idx : (int64_t)-1; };
      auto ff = [](const ClusterPlan&, const std::vector<ClusterTask>& es, const NaryTuple& t) {
        return es[t[0]].team_idx.has_value();
      };
      const int64_t wt = j.second;
      auto wf = [wt](const ClusterPlan&, const std::vector<ClusterTask>&, const NaryTuple&) { return Sc{0, wt}; };
      if (j.first == 2) {
        auto f2 = [](const ClusterPlan&, const ClusterTask& l, const ClusterTask&, size_t, size_t) { return l.team_idx.has_value(); };
        auto w2 = [wt](const ClusterPlan&, const ClusterTask&, const ClusterTask&) { return Sc{0, wt}; };
        dir.constraints.add(std::make_unique<SelfJoinBiConstraint<ClusterPlan, ClusterTask, int64_t, Sc, decltype(kf), decltype(f2), decltype(w2)>>(
            "Cluster pairs", Impact::Penalty, src, kf, f2, w2, false));
      } else {
        dir.constraints.add(std::make_unique<SelfJoinNaryConstraint<ClusterPlan, ClusterTask, int64_t, Sc, decltype(kf), decltype(ff), decltype(wf)>>(
            "Cluster arity " + std::to_string(j.first), Impact::Penalty, j.first, src, kf, ff, wf, false));
      }
    }
  }
  std::vector<Move> enumerate_scalar(MoveStreamContext ctx) override {
    return enumerate_change_moves(dir.working, dir.access, 0, 0, dir.working.n_teams, true, ctx);
  }
};

// ------------------------------------------------------------------------------------------- C1
struct Queen {
  size_t id, column;
  OptVal row_idx;
};
struct Board {
  std::vector<Queen> queens;
  size_t n_rows = 0;
};
inline const std::vector<Queen>& board_queens(const Board& s) { return s.queens; }

struct NQueensModel final : ModelImpl<Board> {
  NQueensModel(Board sol) {
    dir.working = std::move(sol);
    dir.access.get = [](const Board& s, size_t, size_t e) { return s.queens[e].row_idx; };
    dir.access.set = [](Board& s, size_t, size_t e, OptVal v) { s.queens[e].row_idx = v; };
    dir.access.entity_count = [](const Board& s, size_t) { return s.queens.size(); };
    Source<Board, Queen> src{board_queens, ChangeSource::Desc(0)};
    auto uf = [](const Board&, const Queen& q) { return !q.row_idx.has_value(); };
    auto uw = [](const Queen&) { return Sc::ONE_HARD(); };
    dir.constraints.add(std::make_unique<UniConstraint<Board, Queen, Sc, decltype(uf), decltype(uw)>>(
        "Unassigned queen", Impact::Penalty, src, uf, uw, true));
    auto pf = [](const Board&, const Queen& l, const Queen& r, size_t, size_t) {
      if (l.column >= r.column) return false;
      if (!l.row_idx || !r.row_idx) return false;
      size_t lr = *l.row_idx, rr = *r.row_idx;
      size_t dr = lr > rr ? lr - rr : rr - lr;
      size_t dc = l.column > r.column ? l.column - r.column : r.column - l.column;
      return lr == rr || dr == dc;
    };
    auto pw = [](const Board&, const Queen&, const Queen&, size_t, size_t) { return Sc::ONE_HARD(); };
    dir.constraints.add(std::make_unique<CrossBiConstraint<Board, Queen, Queen, uint8_t, Sc, ConstKey, ConstKey,
                                                           decltype(pf), decltype(pw)>>(
        "Queen conflict", Impact::Penalty, src, src, ConstKey{}, ConstKey{}, pf, pw, true));
  }
  std::vector<Move> enumerate_scalar(MoveStreamContext ctx) override {
    return enumerate_change_moves(dir.working, dir.access, 0, 0, dir.working.n_rows, true, ctx);
  }
};

// ------------------------------------------------------------------------------------------- C3
// crates/solverforge-cvrp/src/problem_data.rs:6-47
constexpr int64_t UNREACHABLE = std::numeric_limits<int64_t>::max();
constexpr int64_t MAX_SAFE_LEG_COST = std::numeric_limits<int64_t>::max() / 4;
struct ProblemData {
  int64_t capacity = 0;
  size_t depot = 0;
  std::vector<int32_t> demands;
  std::vector<std::vector<int64_t>> distance_matrix;
  std::optional<int64_t> finite_distance(size_t from, size_t to) const {
    if (from >= distance_matrix.size() || to >= distance_matrix[from].size()) return std::nullopt;
    int64_t v = distance_matrix[from][to];
    if (v >= 0 && v != UNREACHABLE) return v;
    return std::nullopt;
  }
  int64_t distance_cost(size_t from, size_t to) const { return finite_distance(from, to).value_or(MAX_SAFE_LEG_COST); }
};
struct Customer {
  size_t id;
};
struct Route {
  size_t id;
  std::vector<size_t> visits;
  const ProblemData* data;
};
struct CvrpPlan {
  std::vector<Customer> customers;
  std::vector<Route> routes;
  std::shared_ptr<ProblemData> shared;
};
inline const std::vector<Customer>& cvrp_customers(const CvrpPlan& s) { return s.customers; }
inline const std::vector<Route>& cvrp_routes(const CvrpPlan& s) { return s.routes; }

struct CvrpModel final : ModelImpl<CvrpPlan> {
  CvrpModel(CvrpPlan sol) {
    dir.working = std::move(sol);
    dir.access.list = [](CvrpPlan& s, size_t, size_t e) -> std::vector<size_t>& { return s.routes[e].visits; };
    dir.access.entity_count = [](const CvrpPlan& s, size_t) { return s.routes.size(); };
    Source<CvrpPlan, Customer> cust{cvrp_customers, ChangeSource::Stat()};
    Source<CvrpPlan, Route> routes{cvrp_routes, ChangeSource::Desc(0)};
    // (i) all_customers_assigned — publication_plan.rs:51-65
    auto ka = [](const Customer& c) { return c.id; };
    auto kb = [](const size_t& assigned) { return assigned; };
    auto fa = [](const CvrpPlan&, const Customer&) { return true; };
    auto fp = [](const CvrpPlan&, const Route&) { return true; };
    auto fl = [](const Route& r) -> const std::vector<size_t>& { return r.visits; };
    auto w1 = [](const Customer&) { return Sc::ONE_HARD(); };
    dir.constraints.add(std::make_unique<ExistsConstraint<CvrpPlan, Customer, Route, size_t, size_t, Sc, decltype(ka),
                                                          decltype(kb), decltype(fa), decltype(fp), decltype(fl),
                                                          decltype(w1)>>(
        "all_customers_assigned", Impact::Penalty, ExistenceMode::NotExists, cust, routes, ka, kb, fa, fp, fl, w1,
        true));
    // (ii) capacity: of_hard(max(0, load - capacity))
    auto always = [](const CvrpPlan&, const Route&) { return true; };
    auto wcap = [](const Route& r) {
      int64_t load = 0;
      for (size_t v : r.visits) load += r.data->demands[v];
      return Sc::of_hard(std::max<int64_t>(0, load - r.data->capacity));
    };
    dir.constraints.add(std::make_unique<UniConstraint<CvrpPlan, Route, Sc, decltype(always), decltype(wcap)>>(
        "vehicle_capacity", Impact::Penalty, routes, always, wcap, true));
    // (iii) distance: depot -> visits... -> depot via distance_cost; empty route costs 0
    auto wdist = [](const Route& r) {
      if (r.visits.empty()) return Sc::zero();
      int64_t total = 0;
      size_t prev = r.data->depot;
      for (size_t v : r.visits) {
        total = wadd(total, r.data->distance_cost(prev, v));
        prev = v;
      }
      total = wadd(total, r.data->distance_cost(prev, r.data->depot));
      return Sc::of_soft(total);
    };
    dir.constraints.add(std::make_unique<UniConstraint<CvrpPlan, Route, Sc, decltype(always), decltype(wdist)>>(
        "total_distance", Impact::Penalty, routes, always, wdist, false));
  }
  // crates/solverforge-cvrp/src/meters.rs:10-28 (MatrixDistanceMeter)
  static double meter(CvrpPlan& s, size_t se, size_t sp, size_t de, size_t dp) {
    auto& sv = s.routes[se].visits;
    auto& dv = s.routes[de].visits;
    if (sp >= sv.size() || dp >= dv.size()) return std::numeric_limits<double>::infinity();
    auto d = s.shared->finite_distance(sv[sp], dv[dp]);
    return d ? (double)*d : std::numeric_limits<double>::infinity();
  }
  std::vector<Move> enumerate_list(size_t max_nearby, MoveStreamContext ctx) override {
    return enumerate_nearby_list_change_moves(dir.working, dir.access, 0, max_nearby, ctx, meter);
  }
  std::vector<Move> enumerate_list_swap(size_t max_nearby, MoveStreamContext ctx) override {
    return enumerate_nearby_list_swap_moves(dir.working, dir.access, 0, max_nearby, ctx, meter);
  }
  std::vector<Move> enumerate_list_reverse(MoveStreamContext ctx) override {
    return enumerate_list_reverse_moves(dir.working, dir.access, 0, ctx);
  }
  std::vector<Move> enumerate_sublist_change(size_t min_size, size_t max_size, MoveStreamContext ctx) override {
    return enumerate_sublist_change_moves(dir.working, dir.access, 0, min_size, max_size, ctx);
  }
  std::vector<Move> enumerate_sublist_swap(size_t min_size, size_t max_size, MoveStreamContext ctx) override {
    return enumerate_sublist_swap_moves(dir.working, dir.access, 0, min_size, max_size, ctx);
  }
  std::vector<Move> enumerate_k_opt(size_t k, size_t min_seg, MoveStreamContext ctx) override {
    return enumerate_k_opt_moves(dir.working, dir.access, 0, k, min_seg, ctx);
  }
};

// ------------------------------------------------------------------------------------------- C4
struct Operation {
  size_t id, job, step;
  OptVal machine_idx;
};
struct Machine {
  size_t id;
};
struct MachineSequence {
  size_t id;
  std::vector<size_t> operations;
};
struct JobShopPlan {
  std::vector<Machine> machines;
  std::vector<Operation> operations;
  std::vector<MachineSequence> machine_sequences;
};
inline const std::vector<Machine>& js_machines(const JobShopPlan& s) { return s.machines; }
inline const std::vector<Operation>& js_operations(const JobShopPlan& s) { return s.operations; }
inline const std::vector<MachineSequence>& js_sequences(const JobShopPlan& s) { return s.machine_sequences; }

struct OptHash {
  size_t operator()(const OptVal& v) const { return v ? std::hash<size_t>()(*v) + 1 : 0; }
};

struct JobShopModel final : ModelImpl<JobShopPlan> {
  // descriptor indices = declaration order of entity collections (job_shop_plan.rs:15-22):
  // operations = 0, machine_sequences = 1.
  explicit JobShopModel(JobShopPlan sol, bool with_grouped_complement = true) {
    dir.working = std::move(sol);
    dir.access.get = [](const JobShopPlan& s, size_t, size_t e) { return s.operations[e].machine_idx; };
    dir.access.set = [](JobShopPlan& s, size_t, size_t e, OptVal v) { s.operations[e].machine_idx = v; };
    dir.access.list = [](JobShopPlan& s, size_t, size_t e) -> std::vector<size_t>& {
      return s.machine_sequences[e].operations;
    };
    dir.access.entity_count = [](const JobShopPlan& s, size_t d) {
      return d == 0 ? s.operations.size() : s.machine_sequences.size();
    };
    Source<JobShopPlan, Operation> ops{js_operations, ChangeSource::Desc(0)};
    Source<JobShopPlan, MachineSequence> seqs{js_sequences, ChangeSource::Desc(1)};
    Source<JobShopPlan, Machine> machines{js_machines, ChangeSource::Stat()};
    auto uf = [](const JobShopPlan&, const Operation& o) { return !o.machine_idx.has_value(); };
    auto uw = [](const Operation&) { return Sc::ONE_HARD(); };
    dir.constraints.add(std::make_unique<UniConstraint<JobShopPlan, Operation, Sc, decltype(uf), decltype(uw)>>(
        "Unassigned operation machine", Impact::Penalty, ops, uf, uw, true));
    auto ka = [](const Operation& o) { return o.id; };
    auto kb = [](const size_t& assigned) { return assigned; };
    auto fa = [](const JobShopPlan&, const Operation&) { return true; };
    auto fp = [](const JobShopPlan&, const MachineSequence&) { return true; };
    auto fl = [](const MachineSequence& m) -> const std::vector<size_t>& { return m.operations; };
    auto w1 = [](const Operation&) { return Sc::ONE_HARD(); };
    dir.constraints.add(
        std::make_unique<ExistsConstraint<JobShopPlan, Operation, MachineSequence, size_t, size_t, Sc, decltype(ka),
                                          decltype(kb), decltype(fa), decltype(fp), decltype(fl), decltype(w1)>>(
            "Unscheduled operation", Impact::Penalty, ExistenceMode::NotExists, ops, seqs, ka, kb, fa, fp, fl, w1,
            true));
    auto pf = [](const JobShopPlan&, const Operation& l, const Operation& r, size_t, size_t) {
      return l.id < r.id && l.job == r.job && l.machine_idx.has_value() && l.machine_idx == r.machine_idx;
    };
    auto pw = [](const JobShopPlan&, const Operation&, const Operation&, size_t, size_t) { return Sc::ONE_SOFT(); };
    dir.constraints.add(
        std::make_unique<CrossBiConstraint<JobShopPlan, Operation, Operation, uint8_t, Sc, ConstKey, ConstKey,
                                           decltype(pf), decltype(pw)>>(
            "Same job machine reuse", Impact::Penalty, ops, ops, ConstKey{}, ConstKey{}, pf, pw, false));
    if (with_grouped_complement) {
      // for_each(operations).join((machines, equal_bi(op.machine_idx, Some(m.id))))
      //   .group_by(|_, m| m.id, count()).complement(machines, |m| m.id, |_| 0usize)
      //   .penalize(|_, load| of_soft(load^2))
      auto jka = [](const Operation& o) { return o.machine_idx; };
      auto jkb = [](const Machine& m) { return OptVal(m.id); };
      auto jf = [](const JobShopPlan&, const Operation&, const Machine&, size_t, size_t) { return true; };
      auto gk = [](const Operation&, const Machine& m) { return m.id; };
      auto vf = [](const Operation&, const Machine&) { return (char)0; };
      auto kt = [](const Machine& m) { return m.id; };
      auto df = [](const Machine&) { return (size_t)0; };
      auto gw = [](const size_t&, const size_t& load) { return Sc::of_soft((int64_t)(load * load)); };
      dir.constraints.add(
          std::make_unique<CrossComplementedGroupedConstraint<
              JobShopPlan, Operation, Machine, Machine, OptVal, size_t, Sc, CountAcc, decltype(jka), decltype(jkb),
              decltype(jf), decltype(gk), decltype(vf), decltype(kt), decltype(df), decltype(gw), OptHash>>(
              "Machine load balance", Impact::Penalty, ops, machines, machines, jka, jkb, jf, gk, vf, kt, df, gw,
              false));
    }
  }
  std::vector<Move> enumerate_scalar(MoveStreamContext ctx) override {
    return enumerate_change_moves(dir.working, dir.access, 0, 0, dir.working.machines.size(), true, ctx);
  }
  size_t list_desc() const override { return 1; }
};

// ------------------------------------------------------------------------------ shift scheduling
// examples/minimal-shift-scheduling/src/domain/schedule.rs:21-84 — all four constraints verbatim
// ("Long work streaks" uses the consecutive_runs collector, stream/collector/runs.rs), plus
// one authored load_balance constraint over stream/collector/load_balance.rs (metric = `hours`).
struct SNurse {
  size_t id;
};
struct SShift {
  size_t id;
  int64_t day;
  size_t slot;
  bool required;
  int64_t hours;
  OptVal nurse_idx;
};
struct ShiftSchedule {
  std::vector<SNurse> nurses;
  std::vector<SShift> shifts;
};
inline const std::vector<SNurse>& ss_nurses(const ShiftSchedule& s) { return s.nurses; }
inline const std::vector<SShift>& ss_shifts(const ShiftSchedule& s) { return s.shifts; }

struct ShiftModel final : ModelImpl<ShiftSchedule> {
  explicit ShiftModel(ShiftSchedule sol, int64_t target = 4, bool with_load_balance = true, int64_t presence_days = 0) {
    dir.working = std::move(sol);
    dir.access.get = [](const ShiftSchedule& s, size_t, size_t e) { return s.shifts[e].nurse_idx; };
    dir.access.set = [](ShiftSchedule& s, size_t, size_t e, OptVal v) { s.shifts[e].nurse_idx = v; };
    dir.access.entity_count = [](const ShiftSchedule& s, size_t) { return s.shifts.size(); };
    Source<ShiftSchedule, SShift> shifts{ss_shifts, ChangeSource::Desc(0)};
    Source<ShiftSchedule, SNurse> nurses{ss_nurses, ChangeSource::Stat()};
    auto uf = [](const ShiftSchedule&, const SShift& s) { return s.required && !s.nurse_idx.has_value(); };
    auto uw = [](const SShift&) { return Sc::ONE_HARD(); };
    dir.constraints.add(std::make_unique<UniConstraint<ShiftSchedule, SShift, Sc, decltype(uf), decltype(uw)>>(
        "Unassigned required shift", Impact::Penalty, shifts, uf, uw, true));
    auto pf = [](const ShiftSchedule&, const SShift& l, const SShift& r, size_t, size_t) {
      return l.id < r.id && l.day == r.day && l.nurse_idx.has_value() && l.nurse_idx == r.nurse_idx;
    };
    auto pw = [](const ShiftSchedule&, const SShift&, const SShift&, size_t, size_t) { return Sc::ONE_HARD(); };
    dir.constraints.add(
        std::make_unique<CrossBiConstraint<ShiftSchedule, SShift, SShift, uint8_t, Sc, ConstKey, ConstKey,
                                           decltype(pf), decltype(pw)>>(
            "One shift per nurse day", Impact::Penalty, shifts, shifts, ConstKey{}, ConstKey{}, pf, pw, true));
    // Long work streaks (schedule.rs:43-58): group_by(nurse, consecutive_runs(day))
    //   .penalize(|_, runs| of_soft(sum over runs of max(0, point_count - 2)))
    auto sf_ = [](const ShiftSchedule&, const SShift& s) { return s.nurse_idx.has_value(); };
    auto sk = [](const SShift& s) { return s.nurse_idx.value_or((size_t)-1); };
    auto sv = [](const SShift& s) { return s.day; };
    auto sw = [](const size_t&, const Runs& runs) {
      int64_t excess = 0;
      for (auto& r : runs.runs) excess += r.point_count > 2 ? (int64_t)r.point_count - 2 : 0;
      return Sc::of_soft(excess);
    };
    dir.constraints.add(
        std::make_unique<GroupedConstraint<ShiftSchedule, SShift, size_t, Sc, RunsAcc, decltype(sf_), decltype(sk),
                                           decltype(sv), decltype(sw)>>("Long work streaks", Impact::Penalty, shifts,
                                                                        sf_, sk, sv, sw, false));
    if (presence_days > 0) {
      // the three constraints of the reference's indexed_presence example
      // (solverforge-macros/tests/ui/pass/solverforge_constraints_indexed_presence.rs:13-55) over a horizon of
      // presence_days days: rest streaks (complement runs, excess over 1), weekend work (any_in(5..7)), and an
      // authored one on the number of distinct days worked
      const int64_t H = presence_days;
      auto w_rest = [H](const size_t&, const IndexedPresence& p) {
        int64_t t = 0;
        for (auto& r : p.complement_runs(0, H).runs) t += r.point_count > 1 ? (int64_t)r.point_count - 1 : 0;
        return Sc::of_soft(t);
      };
      auto w_weekend = [](const size_t&, const IndexedPresence& p) { return Sc::of_soft(p.any_in(5, 7) ? 1 : 0); };
      auto w_days = [](const size_t&, const IndexedPresence& p) { return Sc::of_soft(2 * (int64_t)p.count()); };
      dir.constraints.add(
          std::make_unique<GroupedConstraint<ShiftSchedule, SShift, size_t, Sc, IndexedPresenceAcc, decltype(sf_),
                                             decltype(sk), decltype(sv), decltype(w_rest)>>(
              "Rest streaks", Impact::Penalty, shifts, sf_, sk, sv, w_rest, false));
      dir.constraints.add(
          std::make_unique<GroupedConstraint<ShiftSchedule, SShift, size_t, Sc, IndexedPresenceAcc, decltype(sf_),
                                             decltype(sk), decltype(sv), decltype(w_weekend)>>(
              "Weekend work", Impact::Penalty, shifts, sf_, sk, sv, w_weekend, false));
      dir.constraints.add(
          std::make_unique<GroupedConstraint<ShiftSchedule, SShift, size_t, Sc, IndexedPresenceAcc, decltype(sf_),
                                             decltype(sk), decltype(sv), decltype(w_days)>>(
              "Days worked", Impact::Penalty, shifts, sf_, sk, sv, w_days, false));
    }
    // Balanced workload: group_by(nurse, count()).complement(nurses, id, 0).penalize(|count - target|)
    auto jka = [](const SShift& s) { return s.nurse_idx; };
    auto jkb = [](const SNurse& n) { return OptVal(n.id); };
    auto jf = [](const ShiftSchedule&, const SShift&, const SNurse&, size_t, size_t) { return true; };
    auto gk = [](const SShift&, const SNurse& n) { return n.id; };
    auto vf = [](const SShift&, const SNurse&) { return (char)0; };
    auto kt = [](const SNurse& n) { return n.id; };
    auto df = [](const SNurse&) { return (size_t)0; };
    auto gw = [target](const size_t&, const size_t& count) {
      int64_t d = (int64_t)count - target;
      return Sc::of_soft(d < 0 ? -d : d);
    };
    dir.constraints.add(
        std::make_unique<CrossComplementedGroupedConstraint<
            ShiftSchedule, SShift, SNurse, SNurse, OptVal, size_t, Sc, CountAcc, decltype(jka), decltype(jkb),
            decltype(jf), decltype(gk), decltype(vf), decltype(kt), decltype(df), decltype(gw), OptHash>>(
            "Balanced workload", Impact::Penalty, shifts, nurses, nurses, jka, jkb, jf, gk, vf, kt, df, gw, false));
    if (!with_load_balance) return;  // the example as shipped (swap / compound candidates stay expressible)
    // authored: group_by(|_| (), load_balance(|s| nurse, |s| hours)).penalize(|_, lb| of_soft(lb.unfairness()))
    auto lf = [](const ShiftSchedule&, const SShift& s) { return s.nurse_idx.has_value(); };
    auto lk = [](const SShift&) { return (char)0; };
    auto lv = [](const SShift& s) { return std::make_pair((int64_t)*s.nurse_idx, s.hours); };
    auto lw = [](const char&, const int64_t& unfairness) { return Sc::of_soft(unfairness); };
    dir.constraints.add(
        std::make_unique<GroupedConstraint<ShiftSchedule, SShift, char, Sc, LoadBalanceAcc, decltype(lf), decltype(lk),
                                           decltype(lv), decltype(lw)>>("Fair hours", Impact::Penalty, shifts, lf, lk,
                                                                        lv, lw, false));
  }
  std::vector<Move> enumerate_scalar(MoveStreamContext ctx) override {
    return enumerate_change_moves(dir.working, dir.access, 0, 0, dir.working.nurses.size(), true, ctx);
  }
};

// ------------------------------------------------------------------------------ roster (projected rows)
// Authored with the reference's `.project(..)` API (stream/projected_stream/uni.rs): a shift spans one
// or more days; when it is assigned it emits one row per spanned day, Row{nurse, day, hours}
// (MAX_EMITS = 8). Constraints:
//   1 for_each(shifts).filter(required && unassigned).penalize(ONE_HARD)
//   2 .project(rows).group_by(|r| (r.nurse, r.day), sum(|r| r.hours)).penalize(hard max(0, total - limit))
//   3 .project(rows).group_by(|r| (r.nurse, r.day), count()).penalize(soft count^2)
//   4 .project(rows).penalize(|r| soft r.hours)       (projected uni terminal, every row scored on its own)
//   5 .project(rows).join(equal(key)).penalize(ONE_HARD)   (projected keyed self-join)
struct RShift {
  size_t id;
  bool required;
  std::vector<std::pair<int64_t, int64_t>> spans;  // (day, hours) rows, in emit order
  OptVal nurse_idx;
};
struct RRow {
  size_t nurse;
  int64_t day, hours;
};
struct Roster {
  std::vector<RShift> shifts;
  size_t n_nurses = 0;
  int64_t n_days = 0;
};
inline const std::vector<RShift>& rs_shifts(const Roster& s) { return s.shifts; }

struct RosterModel final : ModelImpl<Roster> {
  explicit RosterModel(Roster sol, int64_t limit) {
    dir.working = std::move(sol);
    dir.access.get = [](const Roster& s, size_t, size_t e) { return s.shifts[e].nurse_idx; };
    dir.access.set = [](Roster& s, size_t, size_t e, OptVal v) { s.shifts[e].nurse_idx = v; };
    dir.access.entity_count = [](const Roster& s, size_t) { return s.shifts.size(); };
    Source<Roster, RShift> shifts{rs_shifts, ChangeSource::Desc(0)};
    auto uf = [](const Roster&, const RShift& s) { return s.required && !s.nurse_idx.has_value(); };
    auto uw = [](const RShift&) { return Sc::ONE_HARD(); };
    dir.constraints.add(std::make_unique<UniConstraint<Roster, RShift, Sc, decltype(uf), decltype(uw)>>(
        "Unassigned required shift", Impact::Penalty, shifts, uf, uw, true));
    auto project = [](const RShift& s, std::vector<RRow>& out) {
      if (!s.nurse_idx) return;
      for (auto& sp : s.spans) out.push_back({*s.nurse_idx, sp.first, sp.second});
    };
    auto always = [](const Roster&, const RRow&) { return true; };
    const int64_t n_days = dir.working.n_days;
    auto key = [n_days](const RRow& r) { return (int64_t)r.nurse * n_days + r.day; };
    auto hours = [](const RRow& r) { return r.hours; };
    auto hw = [limit](const int64_t&, const int64_t& total) { return Sc::of_hard(total > limit ? total - limit : 0); };
    dir.constraints.add(
        std::make_unique<ProjectedGroupedConstraint<Roster, RShift, RRow, int64_t, Sc, SumAcc, decltype(project),
                                                    decltype(always), decltype(key), decltype(hours), decltype(hw)>>(
            "Daily hours", Impact::Penalty, shifts, project, always, key, hours, hw, true));
    auto unit = [](const RRow&) { return (char)0; };
    auto cw = [](const int64_t&, const size_t& n) { return Sc::of_soft((int64_t)(n * n)); };
    dir.constraints.add(
        std::make_unique<ProjectedGroupedConstraint<Roster, RShift, RRow, int64_t, Sc, CountAcc, decltype(project),
                                                    decltype(always), decltype(key), decltype(unit), decltype(cw)>>(
            "Fragmented days", Impact::Penalty, shifts, project, always, key, unit, cw, false));
    // 5 .project(rows).join(equal(|r| (r.nurse, r.day))).penalize(ONE_HARD): every pair of rows of one nurse
    //   on one day (constraint/projected/bi.rs), rows of the same shift included
    auto pf = [](const RRow&, const RRow&) { return true; };
    auto pw = [](const RRow&, const RRow&) { return Sc::ONE_HARD(); };
    dir.constraints.add(
        std::make_unique<ProjectedBiConstraint<Roster, RShift, RRow, int64_t, Sc, decltype(project), decltype(always),
                                               decltype(key), decltype(pf), decltype(pw)>>(
            "Double booking", Impact::Penalty, shifts, project, always, key, pf, pw, true));
    auto rw = [](const RRow& r) { return Sc::of_soft(r.hours); };
    dir.constraints.add(
        std::make_unique<ProjectedUniConstraint<Roster, RShift, RRow, Sc, decltype(project), decltype(always), decltype(rw)>>(
            "Worked hours", Impact::Penalty, shifts, project, always, rw, false));
  }
  std::vector<Move> enumerate_scalar(MoveStreamContext ctx) override {
    return enumerate_change_moves(dir.working, dir.access, 0, 0, dir.working.n_nurses, true, ctx);
  }
};

// ---------------------------------------------------------------------------------------------
// The fixture of the reference's cross-bi known-answer tests (constraint/tests/cross_bi_incr.rs:17-165:
// Shift{employee_id: Option, day} x Employee{id, unavailable_days}) as a planning model, with the reference's own
// constraints and authored pair-weight / index-aware / multi-row-per-key variants of the same CrossBiConstraint:
//   1 "Unassigned shift"      for_each(shifts).filter(employee.is_none()).penalize(1 hard)
//   2 "Unavailable employee"  join(employees, equal_bi(shift.employee, Some(employee.id)))
//                             .filter(employee.unavailable_days.contains(shift.day)).penalize(1 hard)  [:63-88]
//   3 "Skill gap"             same join, filter shift.required > employee.skill,
//                             penalize((required - skill) * shift.hours soft) — a pair weight reading both sides
//   4 "Indexed pairs"         same join, filter (shift_idx + 2 * employee_idx) % 3 == 0, penalize(shift.day soft)
//                             — the index-aware filter of :308-341
//   5 "Contract window"       join(contracts, equal_bi(shift.employee, Some(contract.employee))) — several
//                             contract rows per employee —, filter day outside [from, to], penalize(contract.fee soft)
struct AvShift {
  size_t id;
  int64_t day, required, hours;
  OptVal employee;
};
struct AvEmployee {
  size_t id;
  int64_t skill;
  std::vector<int64_t> unavailable_days;
};
struct AvContract {
  size_t employee;
  int64_t from, to, fee;
};
struct AvSchedule {
  std::vector<AvShift> shifts;
  std::vector<AvEmployee> employees;
  std::vector<AvContract> contracts;
};
inline const std::vector<AvShift>& av_shifts(const AvSchedule& s) { return s.shifts; }
inline const std::vector<AvEmployee>& av_employees(const AvSchedule& s) { return s.employees; }
inline const std::vector<AvContract>& av_contracts(const AvSchedule& s) { return s.contracts; }

struct AvailabilityModel final : ModelImpl<AvSchedule> {
  explicit AvailabilityModel(AvSchedule sol) {
    dir.working = std::move(sol);
    dir.access.get = [](const AvSchedule& s, size_t, size_t e) { return s.shifts[e].employee; };
    dir.access.set = [](AvSchedule& s, size_t, size_t e, OptVal v) { s.shifts[e].employee = v; };
    dir.access.entity_count = [](const AvSchedule& s, size_t) { return s.shifts.size(); };
    Source<AvSchedule, AvShift> shifts{av_shifts, ChangeSource::Desc(0)};
    Source<AvSchedule, AvEmployee> employees{av_employees, ChangeSource::Stat()};
    Source<AvSchedule, AvContract> contracts{av_contracts, ChangeSource::Stat()};
    auto uf = [](const AvSchedule&, const AvShift& s) { return !s.employee.has_value(); };
    auto uw = [](const AvShift&) { return Sc::ONE_HARD(); };
    dir.constraints.add(std::make_unique<UniConstraint<AvSchedule, AvShift, Sc, decltype(uf), decltype(uw)>>(
        "Unassigned shift", Impact::Penalty, shifts, uf, uw, true));
    auto ka = [](const AvShift& s) { return s.employee.has_value() ? (int64_t)*s.employee : (int64_t)-1; };
    auto kb = [](const AvEmployee& e) { return (int64_t)e.id; };
    auto f2 = [](const AvSchedule&, const AvShift& s, const AvEmployee& e, size_t, size_t) {
      return s.employee.has_value() && std::find(e.unavailable_days.begin(), e.unavailable_days.end(), s.day) != e.unavailable_days.end();
    };
    auto w2 = [](const AvSchedule&, const AvShift&, const AvEmployee&, size_t, size_t) { return Sc::ONE_HARD(); };
    dir.constraints.add(std::make_unique<CrossBiConstraint<AvSchedule, AvShift, AvEmployee, int64_t, Sc, decltype(ka),
                                                           decltype(kb), decltype(f2), decltype(w2)>>(
        "Unavailable employee", Impact::Penalty, shifts, employees, ka, kb, f2, w2, true));
    auto f3 = [](const AvSchedule&, const AvShift& s, const AvEmployee& e, size_t, size_t) { return s.required > e.skill; };
    auto w3 = [](const AvSchedule&, const AvShift& s, const AvEmployee& e, size_t, size_t) {
      return Sc::of_soft((s.required - e.skill) * s.hours);
    };
    dir.constraints.add(std::make_unique<CrossBiConstraint<AvSchedule, AvShift, AvEmployee, int64_t, Sc, decltype(ka),
                                                           decltype(kb), decltype(f3), decltype(w3)>>(
        "Skill gap", Impact::Penalty, shifts, employees, ka, kb, f3, w3, false));
    auto f4 = [](const AvSchedule&, const AvShift&, const AvEmployee&, size_t ia, size_t ib) { return (ia + 2 * ib) % 3 == 0; };
    auto w4 = [](const AvSchedule&, const AvShift& s, const AvEmployee&, size_t, size_t) { return Sc::of_soft(s.day); };
    dir.constraints.add(std::make_unique<CrossBiConstraint<AvSchedule, AvShift, AvEmployee, int64_t, Sc, decltype(ka),
                                                           decltype(kb), decltype(f4), decltype(w4)>>(
        "Indexed pairs", Impact::Penalty, shifts, employees, ka, kb, f4, w4, false));
    auto kc = [](const AvContract& c) { return (int64_t)c.employee; };
    auto f5 = [](const AvSchedule&, const AvShift& s, const AvContract& c, size_t, size_t) { return s.day < c.from || s.day > c.to; };
    auto w5 = [](const AvSchedule&, const AvShift&, const AvContract& c, size_t, size_t) { return Sc::of_soft(c.fee); };
    dir.constraints.add(std::make_unique<CrossBiConstraint<AvSchedule, AvShift, AvContract, int64_t, Sc, decltype(ka),
                                                           decltype(kc), decltype(f5), decltype(w5)>>(
        "Contract window", Impact::Penalty, shifts, contracts, ka, kc, f5, w5, false));
  }
  std::vector<Move> enumerate_scalar(MoveStreamContext ctx) override {
    return enumerate_change_moves(dir.working, dir.access, 0, 0, dir.working.employees.size(), true, ctx);
  }
  std::vector<Move> enumerate_scalar_swap(MoveStreamContext ctx) override {
    return enumerate_swap_moves(dir.working, dir.access, 0, 0, ctx);
  }
};

// ---------------------------------------------------------------------------------------------
// The fixture of the reference's projected self-join tests (constraint/tests/projected/self_join.rs: Work{bucket,
// demand}) as a planning model — the bucket is the planning variable — with the reference's own constraints:
//   1 "Unassigned work"             1 hard
//   2 "projected duplicate bucket"  .project(entry).join(equal(bucket)).filter(left.delta < right.delta).penalize(1)   [:108-154]
//   3 "priority spread"             the same keyed self-join with a pair weight |left.prio - right.prio| (authored)
//   4 "projected parent child"      .project(row).join(equal_bi(left.group, right.parent)).penalize(left.group * 10 +
//                                   right.group), group = bucket, parent = (demand >= 0).then(demand)               [:156-290]
struct PwWork {
  size_t id;
  int64_t demand, prio;
  OptVal bucket;
};
struct PwPlan {
  std::vector<PwWork> work;
  size_t n_buckets = 0;
};
inline const std::vector<PwWork>& pw_work(const PwPlan& s) { return s.work; }

struct PairsModel final : ModelImpl<PwPlan> {
  explicit PairsModel(PwPlan sol) {
    dir.working = std::move(sol);
    dir.access.get = [](const PwPlan& s, size_t, size_t e) { return s.work[e].bucket; };
    dir.access.set = [](PwPlan& s, size_t, size_t e, OptVal v) { s.work[e].bucket = v; };
    dir.access.entity_count = [](const PwPlan& s, size_t) { return s.work.size(); };
    Source<PwPlan, PwWork> src{pw_work, ChangeSource::Desc(0)};
    auto uf = [](const PwPlan&, const PwWork& w) { return !w.bucket.has_value(); };
    auto uw = [](const PwWork&) { return Sc::ONE_HARD(); };
    dir.constraints.add(std::make_unique<UniConstraint<PwPlan, PwWork, Sc, decltype(uf), decltype(uw)>>(
        "Unassigned work", Impact::Penalty, src, uf, uw, true));
    auto kf = [](const PwWork& w) { return w.bucket.has_value() ? (int64_t)*w.bucket : (int64_t)-1; };
    auto f2 = [](const PwPlan&, const PwWork& l, const PwWork& r, size_t, size_t) {
      return l.bucket.has_value() && l.demand < r.demand;
    };
    auto w2 = [](const PwPlan&, const PwWork&, const PwWork&) { return Sc::of_soft(1); };
    dir.constraints.add(std::make_unique<SelfJoinBiConstraint<PwPlan, PwWork, int64_t, Sc, decltype(kf), decltype(f2), decltype(w2)>>(
        "projected duplicate bucket", Impact::Penalty, src, kf, f2, w2, false));
    auto f3 = [](const PwPlan&, const PwWork& l, const PwWork&, size_t, size_t) { return l.bucket.has_value(); };
    auto w3 = [](const PwPlan&, const PwWork& l, const PwWork& r) { return Sc::of_soft(std::llabs(l.prio - r.prio)); };
    dir.constraints.add(std::make_unique<SelfJoinBiConstraint<PwPlan, PwWork, int64_t, Sc, decltype(kf), decltype(f3), decltype(w3)>>(
        "priority spread", Impact::Penalty, src, kf, f3, w3, false));
    // directed: left key = group, right key = parent; rows exist for assigned work only; a row never pairs with itself
    auto kl = [](const PwWork& w) { return w.bucket.has_value() ? (int64_t)*w.bucket : (int64_t)-1; };
    auto kr = [](const PwWork& w) { return (w.bucket.has_value() && w.demand >= 0) ? w.demand : (int64_t)-2; };
    auto f4 = [](const PwPlan&, const PwWork& l, const PwWork& r, size_t il, size_t ir) {
      return il != ir && l.bucket.has_value() && r.bucket.has_value();
    };
    auto w4 = [](const PwPlan&, const PwWork& l, const PwWork& r, size_t, size_t) {
      return Sc::of_soft((int64_t)(*l.bucket * 10 + *r.bucket));
    };
    dir.constraints.add(std::make_unique<CrossBiConstraint<PwPlan, PwWork, PwWork, int64_t, Sc, decltype(kl), decltype(kr),
                                                           decltype(f4), decltype(w4)>>(
        "projected parent child", Impact::Penalty, src, src, kl, kr, f4, w4, false));
  }
  std::vector<Move> enumerate_scalar(MoveStreamContext ctx) override {
    return enumerate_change_moves(dir.working, dir.access, 0, 0, dir.working.n_buckets, true, ctx);
  }
  std::vector<Move> enumerate_scalar_swap(MoveStreamContext ctx) override {
    return enumerate_swap_moves(dir.working, dir.access, 0, 0, ctx);
  }
};

}  // namespace sfo
