// ORACLE — TEST INFRASTRUCTURE ONLY (see score.hpp header).
//
// Pins the oracle against the reference's own known-answer tests. Every block re-encodes one
// reference test (fixture + asserted numbers), cited as file:line under /root/reference/crates.
// Exit code 0 and a final "KAT OK n" line mean every vector matched.
#include <cstdio>
#include <cstdlib>
#include <set>

#include "models.hpp"

using namespace sfo;

static int g_checks = 0, g_fail = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    ++g_checks;                                                            \
    if (!(cond)) {                                                         \
      ++g_fail;                                                            \
      std::fprintf(stderr, "KAT FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
    }                                                                      \
  } while (0)

// ---------------------------------------------------------------- scores
static void kat_scores() {
  // solverforge-core/src/score/tests/hard_soft_score.rs:3-57
  auto s = HardSoftScore::of(-2, -100);
  CHECK(s.hard == -2 && s.soft == -100);
  CHECK(HardSoftScore::of(0, -1000).is_feasible());
  CHECK(HardSoftScore::of(10, -50).is_feasible());
  CHECK(!HardSoftScore::of(-1, 0).is_feasible());
  CHECK(HardSoftScore::of(0, -1000) > HardSoftScore::of(-1, 0));
  CHECK(HardSoftScore::of(0, -50) > HardSoftScore::of(0, -100));
  CHECK(HardSoftScore::of(-1, -1000) > HardSoftScore::of(-2, 0));
  auto s1 = HardSoftScore::of(-1, -100), s2 = HardSoftScore::of(-1, -50);
  CHECK(s1 + s2 == HardSoftScore::of(-2, -150));
  CHECK(s1 - s2 == HardSoftScore::of(0, -50));
  CHECK(-s1 == HardSoftScore::of(1, 100));
  CHECK(s1.str() == "-1hard/-100soft");
  // solverforge-core/src/score/tests/hard_soft_decimal_score.rs:3-60
  auto d = HardSoftDecimalScore::of(-2, -100);
  CHECK(d.hard == -200000 && d.soft == -10000000);
  auto ds = HardSoftDecimalScore::of_scaled(-30500, -208250);
  CHECK(ds.hard == -30500 && ds.soft == -208250);
  CHECK(!HardSoftDecimalScore::of_scaled(-1, 0).is_feasible());
  CHECK(HardSoftDecimalScore::of(0, -1000) > HardSoftDecimalScore::of(-1, 0));
  auto d1 = HardSoftDecimalScore::of(-1, -100), d2 = HardSoftDecimalScore::of(-1, -50);
  CHECK(d1 + d2 == HardSoftDecimalScore::of(-2, -150));
  CHECK(d1 - d2 == HardSoftDecimalScore::of(0, -50));
  // solverforge-solver/src/phase/hard_delta.rs:11-35
  CHECK(hard_score_delta(HardSoftScore::of(-2, 0), HardSoftScore::of(-1, -9)) == HardDelta::Improving);
  CHECK(hard_score_delta(HardSoftScore::of(-2, 0), HardSoftScore::of(-2, 5)) == HardDelta::Neutral);
  CHECK(hard_score_delta(HardSoftScore::of(-2, 0), HardSoftScore::of(-3, 5)) == HardDelta::Worse);
  CHECK(hard_score_delta(SoftScore::of(1), SoftScore::of(2)) == HardDelta::None);
}

// ---------------------------------------------------------------- uni + director
struct TestSolution {
  std::vector<OptVal> values;
};
static const std::vector<OptVal>& ts_values(const TestSolution& s) { return s.values; }

static void kat_director() {
  // solverforge-scoring/src/director/tests/score_director.rs:63-115
  auto make = [](std::vector<OptVal> v) {
    auto d = std::make_unique<ScoreDirector<TestSolution, SoftScore>>();
    d->working.values = std::move(v);
    auto f = [](const TestSolution&, const OptVal& v) { return !v.has_value(); };
    auto w = [](const OptVal&) { return SoftScore::of(1); };
    d->constraints.add(std::make_unique<UniConstraint<TestSolution, OptVal, SoftScore, decltype(f), decltype(w)>>(
        "Unassigned", Impact::Penalty, Source<TestSolution, OptVal>{ts_values, ChangeSource::Desc(0)}, f, w, false));
    return d;
  };
  {
    auto d = make({OptVal(1), std::nullopt, std::nullopt, OptVal(2)});
    CHECK(!d->initialized);
    CHECK(d->calculate_score() == SoftScore::of(-2));
    CHECK(d->initialized);
  }
  {
    auto d = make({OptVal(1), std::nullopt});
    CHECK(d->calculate_score() == SoftScore::of(-1));
    CHECK(d->calculate_score() == SoftScore::of(-1));
  }
  {
    auto d = make({OptVal(1), std::nullopt, OptVal(2)});
    CHECK(d->calculate_score() == SoftScore::of(-1));
    d->before_variable_changed(0, 1);
    d->working.values[1] = OptVal(3);
    d->after_variable_changed(0, 1);
    CHECK(d->calculate_score() == SoftScore::of(0));
    CHECK(d->fresh_score() == SoftScore::of(0));
    // out-of-range entity => zero delta (constraint/incremental.rs:128-130)
    d->before_variable_changed(0, 99);
    d->after_variable_changed(0, 99);
    CHECK(d->calculate_score() == SoftScore::of(0));
    // foreign descriptor => ignored (collection_extract.rs:83-93)
    d->before_variable_changed(7, 0);
    d->after_variable_changed(7, 0);
    CHECK(d->calculate_score() == SoftScore::of(0));
  }
}

// ---------------------------------------------------------------- cross-bi
struct Shift {
  OptVal employee_id;
  int day;
};
struct Employee {
  size_t id;
  std::vector<int> unavailable_days;
};
struct Schedule {
  std::vector<Shift> shifts;
  std::vector<Employee> employees;
};
static const std::vector<Shift>& sch_shifts(const Schedule& s) { return s.shifts; }
static const std::vector<Employee>& sch_employees(const Schedule& s) { return s.employees; }
struct OptKeyHash {
  size_t operator()(const OptVal& v) const { return v ? *v + 1 : 0; }
};

static void kat_cross_bi() {
  // solverforge-scoring/src/constraint/tests/cross_bi_incr.rs:122-165 fixtures
  auto sample = [] {
    return Schedule{{{OptVal(0), 5}, {OptVal(0), 6}}, {{0, {5}}}};
  };
  auto two_emp = [] {
    return Schedule{{{OptVal(0), 5}, {OptVal(0), 6}}, {{0, {}}, {1, {}}}};
  };
  Source<Schedule, Shift> ss{sch_shifts, ChangeSource::Desc(0)};
  Source<Schedule, Employee> se{sch_employees, ChangeSource::Desc(1)};
  auto ka = [](const Shift& s) { return s.employee_id; };
  auto kb = [](const Employee& e) { return OptVal(e.id); };
  auto f = [](const Schedule&, const Shift& s, const Employee& e, size_t, size_t) {
    return s.employee_id.has_value() &&
           std::find(e.unavailable_days.begin(), e.unavailable_days.end(), s.day) != e.unavailable_days.end();
  };
  auto w = [](const Schedule&, const Shift&, const Employee&, size_t, size_t) { return SoftScore::of(1); };
  using C = CrossBiConstraint<Schedule, Shift, Employee, OptVal, SoftScore, decltype(ka), decltype(kb), decltype(f),
                              decltype(w), OptKeyHash>;
  {
    // :193-201 evaluate without initialize == -1; match_count 1 (:205-217)
    C c("Unavailable employee", Impact::Penalty, ss, se, ka, kb, f, w, false);
    auto sc = sample();
    CHECK(c.evaluate(sc) == SoftScore::of(-1));
    CHECK(c.match_count(sc) == 1);
  }
  {
    // :252-264 incremental retract/insert: -1, +1, -1; :167-191 unrelated descriptor => zero
    C c("Unavailable employee", Impact::Penalty, ss, se, ka, kb, f, w, false);
    auto sc = sample();
    CHECK(c.initialize(sc) == SoftScore::of(-1));
    CHECK(c.on_insert(sc, 0, 2) == SoftScore::zero());
    CHECK(c.on_retract(sc, 0, 0) == SoftScore::of(1));
    CHECK(c.on_insert(sc, 0, 0) == SoftScore::of(-1));
  }
  {
    // :267-279 B-side retract / mutate / insert
    C c("Unavailable employee", Impact::Penalty, ss, se, ka, kb, f, w, false);
    auto sc = sample();
    SoftScore total = c.initialize(sc);
    CHECK(total == SoftScore::of(-1));
    total = total + c.on_retract(sc, 0, 1);
    sc.employees[0].unavailable_days = {6};
    total = total + c.on_insert(sc, 0, 1);
    CHECK(total == SoftScore::of(-1));
    CHECK(total == c.evaluate(sc));
  }
  {
    // :308-341 index-aware filter: shift_idx == 1 && employee_idx == 0, weight = shift.day => -6
    auto fi = [](const Schedule&, const Shift&, const Employee&, size_t si, size_t ei) { return si == 1 && ei == 0; };
    auto wi = [](const Schedule&, const Shift& s, const Employee&, size_t, size_t) { return SoftScore::of(s.day); };
    CrossBiConstraint<Schedule, Shift, Employee, OptVal, SoftScore, decltype(ka), decltype(kb), decltype(fi),
                      decltype(wi), OptKeyHash>
        c("indexed cross path", Impact::Penalty, ss, se, ka, kb, fi, wi, false);
    auto sc = two_emp();
    CHECK(c.match_count(sc) == 1);
    CHECK(c.evaluate(sc) == SoftScore::of(-6));
  }
  {
    // :283-304 cross group_by (sum of 1 per pair, penalise count^2): -4 -> -2 after moving shift 1
    auto tf = [](const Schedule&, const Shift&, const Employee&, size_t, size_t) { return true; };
    auto gk = [](const Shift&, const Employee& e) { return e.id; };
    auto vf = [](const Shift&, const Employee&) { return (int64_t)1; };
    auto kt = [](const Employee& e) { return e.id; };
    auto df = [](const Employee&) { return (int64_t)0; };
    auto gw = [](const size_t&, const int64_t& c) { return SoftScore::of(c * c); };
    CrossComplementedGroupedConstraint<Schedule, Shift, Employee, Employee, OptVal, size_t, SoftScore, SumAcc,
                                       decltype(ka), decltype(kb), decltype(tf), decltype(gk), decltype(vf),
                                       decltype(kt), decltype(df), decltype(gw), OptKeyHash>
        c("grouped assigned shift count", Impact::Penalty, ss, se, se, ka, kb, tf, gk, vf, kt, df, gw, false,
          /*with_complement=*/false);
    auto sc = two_emp();
    CHECK(c.match_count(sc) == 1);
    CHECK(c.evaluate(sc) == SoftScore::of(-4));
    SoftScore total = c.initialize(sc);
    CHECK(total == SoftScore::of(-4));
    total = total + c.on_retract(sc, 1, 0);
    sc.shifts[1].employee_id = OptVal(1);
    total = total + c.on_insert(sc, 1, 0);
    CHECK(total == SoftScore::of(-2));
    CHECK(total == c.evaluate(sc));
  }
}

// ---------------------------------------------------------------- self-join bi
struct BQueen {
  int64_t row, col;
};
struct BSolution {
  std::vector<BQueen> queens;
};
static const std::vector<BQueen>& bq(const BSolution& s) { return s.queens; }
static void kat_self_join() {
  // solverforge-scoring/src/constraint/tests/bi_incr.rs:21-200
  Source<BSolution, BQueen> src{bq, ChangeSource::Desc(0)};
  auto key = [](const BQueen& q) { return q.row; };
  auto f = [](const BSolution&, const BQueen& a, const BQueen& b, size_t, size_t) { return a.col < b.col; };
  auto w = [](const BSolution&, const BQueen&, const BQueen&) { return SoftScore::of(1); };
  using C = SelfJoinBiConstraint<BSolution, BQueen, int64_t, SoftScore, decltype(key), decltype(f), decltype(w)>;
  {
    C c("Row conflict", Impact::Penalty, src, key, f, w, false);
    BSolution s{{{0, 0}, {1, 1}, {2, 2}}};
    CHECK(c.evaluate(s) == SoftScore::of(0));
    CHECK(c.match_count(s) == 0);
  }
  {
    C c("Row conflict", Impact::Penalty, src, key, f, w, false);
    BSolution s{{{0, 0}, {0, 1}, {2, 2}}};
    CHECK(c.evaluate(s) == SoftScore::of(-1));
    CHECK(c.match_count(s) == 1);
    c.initialize(s);
    c.reset();
    CHECK(c.on_insert(s, 0, 0) == SoftScore::of(0));
    CHECK(c.on_insert(s, 1, 0) == SoftScore::of(-1));
    CHECK(c.on_insert(s, 2, 0) == SoftScore::of(0));
    CHECK(c.on_retract(s, 0, 0) == SoftScore::of(1));
  }
  {
    // :172-200 dynamic weight |b.col - a.col| => -3
    auto wd = [](const BSolution&, const BQueen& a, const BQueen& b) { return SoftScore::of(std::llabs(b.col - a.col)); };
    SelfJoinBiConstraint<BSolution, BQueen, int64_t, SoftScore, decltype(key), decltype(f), decltype(wd)> c(
        "Column distance", Impact::Penalty, src, key, f, wd, false);
    BSolution s{{{0, 0}, {0, 3}}};
    CHECK(c.evaluate(s) == SoftScore::of(-3));
  }
  {
    // :143-170 reward, adjacent columns => +2
    auto fr = [](const BSolution&, const BQueen& a, const BQueen& b, size_t, size_t) {
      return a.col < b.col && std::llabs(a.col - b.col) == 1;
    };
    auto w2 = [](const BSolution&, const BQueen&, const BQueen&) { return SoftScore::of(2); };
    SelfJoinBiConstraint<BSolution, BQueen, int64_t, SoftScore, decltype(key), decltype(fr), decltype(w2)> c(
        "Adjacent queens", Impact::Reward, src, key, fr, w2, false);
    BSolution s{{{0, 0}, {0, 1}}};
    CHECK(c.evaluate(s) == SoftScore::of(2));
  }
}

// ---------------------------------------------------------------- exists
struct Task {
  OptVal assignee;
};
struct Worker {
  size_t id;
  bool available;
};
struct TaskSchedule {
  std::vector<Task> tasks;
  std::vector<Worker> workers;
};
static const std::vector<Task>& tsk(const TaskSchedule& s) { return s.tasks; }
static const std::vector<Worker>& wrk(const TaskSchedule& s) { return s.workers; }
struct CustomerState {
  std::vector<size_t> customers;
  std::vector<std::vector<size_t>> routes;
};
static const std::vector<size_t>& cs_c(const CustomerState& s) { return s.customers; }
static const std::vector<std::vector<size_t>>& cs_r(const CustomerState& s) { return s.routes; }
struct TaggedItem {
  size_t key;
  bool enabled;
};
struct TaggedItems {
  std::vector<TaggedItem> items;
};
static const std::vector<TaggedItem>& ti(const TaggedItems& s) { return s.items; }

static void kat_exists() {
  {
    // solverforge-scoring/src/constraint/tests/exists.rs:34-84: 0 -> -2 after worker 0 becomes unavailable
    auto ka = [](const Task& t) { return t.assignee; };
    auto kb = [](const Worker& w) { return OptVal(w.id); };
    auto fa = [](const TaskSchedule&, const Task& t) { return t.assignee.has_value(); };
    auto fp = [](const TaskSchedule&, const Worker& w) { return !w.available; };
    auto fl = [](const Worker& w) { return std::vector<Worker>{w}; };
    auto w1 = [](const Task&) { return SoftScore::of(1); };
    ExistsConstraint<TaskSchedule, Task, Worker, Worker, OptVal, SoftScore, decltype(ka), decltype(kb), decltype(fa),
                     decltype(fp), decltype(fl), decltype(w1), OptKeyHash>
        c("unavailable worker", Impact::Penalty, ExistenceMode::Exists, {tsk, ChangeSource::Stat()},
          {wrk, ChangeSource::Desc(0)}, ka, kb, fa, fp, fl, w1, false);
    TaskSchedule s{{{OptVal(0)}, {OptVal(0)}, {OptVal(1)}}, {{0, true}, {1, true}}};
    SoftScore total = c.initialize(s);
    CHECK(total == SoftScore::of(0));
    total = total + c.on_retract(s, 0, 0);
    s.workers[0].available = false;
    total = total + c.on_insert(s, 0, 0);
    CHECK(total == c.evaluate(s));
    CHECK(total == SoftScore::of(-2));
  }
  {
    // exists.rs:100-135 flattened not-exists: -3 -> 0 after route gets [1,2,3]
    auto ka = [](const size_t& c) { return c; };
    auto kb = [](const size_t& a) { return a; };
    auto fa = [](const CustomerState&, const size_t&) { return true; };
    auto fp = [](const CustomerState&, const std::vector<size_t>&) { return true; };
    auto fl = [](const std::vector<size_t>& r) -> const std::vector<size_t>& { return r; };
    auto w1 = [](const size_t&) { return SoftScore::of(1); };
    ExistsConstraint<CustomerState, size_t, std::vector<size_t>, size_t, size_t, SoftScore, decltype(ka), decltype(kb),
                     decltype(fa), decltype(fp), decltype(fl), decltype(w1)>
        c("missing assignment", Impact::Penalty, ExistenceMode::NotExists, {cs_c, ChangeSource::Stat()},
          {cs_r, ChangeSource::Desc(0)}, ka, kb, fa, fp, fl, w1, false);
    CustomerState s{{1, 2, 3}, {{}}};
    SoftScore total = c.initialize(s);
    CHECK(total == SoftScore::of(-3));
    total = total + c.on_retract(s, 0, 0);
    s.routes[0] = {1, 2, 3};
    total = total + c.on_insert(s, 0, 0);
    CHECK(total == c.evaluate(s));
    CHECK(total == SoftScore::of(0));
  }
  {
    // exists.rs:150-190 same-source exists: -2 -> 0
    auto ka = [](const TaggedItem& i) { return i.key; };
    auto kb = [](const TaggedItem& i) { return i.key; };
    auto fa = [](const TaggedItems&, const TaggedItem&) { return true; };
    auto fp = [](const TaggedItems&, const TaggedItem& i) { return i.enabled; };
    auto fl = [](const TaggedItem& i) { return std::vector<TaggedItem>{i}; };
    auto w1 = [](const TaggedItem&) { return SoftScore::of(1); };
    ExistsConstraint<TaggedItems, TaggedItem, TaggedItem, TaggedItem, size_t, SoftScore, decltype(ka), decltype(kb),
                     decltype(fa), decltype(fp), decltype(fl), decltype(w1)>
        c("key has enabled peer", Impact::Penalty, ExistenceMode::Exists, {ti, ChangeSource::Desc(0)},
          {ti, ChangeSource::Desc(0)}, ka, kb, fa, fp, fl, w1, false);
    TaggedItems s{{{1, false}, {1, true}}};
    SoftScore total = c.initialize(s);
    CHECK(total == SoftScore::of(-2));
    total = total + c.on_retract(s, 1, 0);
    s.items[1].enabled = false;
    total = total + c.on_insert(s, 1, 0);
    CHECK(total == c.evaluate(s));
    CHECK(total == SoftScore::of(0));
  }
}

// ---------------------------------------------------------------- grouped
struct GShift {
  size_t employee_id;
};
struct GSolution {
  std::vector<GShift> shifts;
};
static const std::vector<GShift>& gs(const GSolution& s) { return s.shifts; }
static void kat_grouped() {
  // solverforge-scoring/src/constraint/tests/grouped.rs:26-150
  Source<GSolution, GShift> src{gs, ChangeSource::Desc(0)};
  auto tf = [](const GSolution&, const GShift&) { return true; };
  auto key = [](const GShift& s) { return s.employee_id; };
  auto val = [](const GShift&) { return (char)0; };
  {
    auto w = [](const size_t&, const size_t& c) { return SoftScore::of((int64_t)(c * c)); };
    GroupedConstraint<GSolution, GShift, size_t, SoftScore, CountAcc, decltype(tf), decltype(key), decltype(val),
                      decltype(w)>
        c("Workload", Impact::Penalty, src, tf, key, val, w, false);
    GSolution s{{{1}, {1}, {1}, {2}}};
    CHECK(c.evaluate(s) == SoftScore::of(-10));
  }
  {
    auto w = [](const size_t&, const size_t& c) { return SoftScore::of((int64_t)c); };
    GroupedConstraint<GSolution, GShift, size_t, SoftScore, CountAcc, decltype(tf), decltype(key), decltype(val),
                      decltype(w)>
        c("Workload", Impact::Penalty, src, tf, key, val, w, false);
    GSolution s{{{1}, {1}, {2}}};
    CHECK(c.initialize(s) == SoftScore::of(-3));
    CHECK(c.on_retract(s, 0, 0) == SoftScore::of(1));
    CHECK(c.on_insert(s, 0, 0) == SoftScore::of(-1));
    GroupedConstraint<GSolution, GShift, size_t, SoftScore, CountAcc, decltype(tf), decltype(key), decltype(val),
                      decltype(w)>
        r("Collaboration", Impact::Reward, src, tf, key, val, w, false);
    GSolution s2{{{1}, {1}}};
    CHECK(r.evaluate(s2) == SoftScore::of(2));
  }
  {
    // solverforge-scoring/src/stream/collector/load_balance.rs:38-55: loads 2,1 => unfairness 1
    LoadBalanceAcc acc;
    acc.accumulate({0, 1});
    acc.accumulate({0, 1});
    acc.accumulate({1, 1});
    CHECK(acc.result() == 1);
    auto r = acc.accumulate({1, 1});  // 2,2 => 0
    CHECK(acc.result() == 0);
    acc.retract(r);
    CHECK(acc.result() == 1);
    // count.rs:85-98 / sum.rs:176-191 round trips
    CountAcc ca;
    ca.accumulate(0);
    ca.accumulate(0);
    CHECK(ca.result() == 2);
    ca.retract(0);
    CHECK(ca.result() == 1);
    SumAcc sa;
    auto r1 = sa.accumulate(5);
    sa.accumulate(7);
    CHECK(sa.result() == 12);
    sa.retract(r1);
    CHECK(sa.result() == 7);
  }
}

// ---------------------------------------------------------------- cross complemented grouped
struct CEmployee {
  size_t id;
};
struct CShift {
  OptVal employee_id;
};
struct CTarget {
  size_t employee_id;
};
struct CSchedule {
  std::vector<CShift> shifts;
  std::vector<CEmployee> employees;
  std::vector<CTarget> targets;
};
static const std::vector<CShift>& c_sh(const CSchedule& s) { return s.shifts; }
static const std::vector<CEmployee>& c_em(const CSchedule& s) { return s.employees; }
static const std::vector<CTarget>& c_tg(const CSchedule& s) { return s.targets; }

static void kat_cross_complemented() {
  // solverforge-scoring/src/constraint/tests/cross_complemented_grouped.rs:58-215
  auto ka = [](const CShift& s) { return s.employee_id; };
  auto kb = [](const CEmployee& e) { return OptVal(e.id); };
  auto tf = [](const CSchedule&, const CShift&, const CEmployee&, size_t, size_t) { return true; };
  auto gk = [](const CShift&, const CEmployee& e) { return e.id; };
  auto vf = [](const CShift&, const CEmployee&) { return (int64_t)1; };
  auto kt = [](const CTarget& t) { return t.employee_id; };
  auto df = [](const CTarget&) { return (int64_t)5; };
  auto gw = [](const size_t&, const int64_t& c) { return SoftScore::of(c); };
  using C = CrossComplementedGroupedConstraint<CSchedule, CShift, CEmployee, CTarget, OptVal, size_t, SoftScore, SumAcc,
                                               decltype(ka), decltype(kb), decltype(tf), decltype(gk), decltype(vf),
                                               decltype(kt), decltype(df), decltype(gw), OptKeyHash>;
  Source<CSchedule, CShift> ss{c_sh, ChangeSource::Desc(0)};
  Source<CSchedule, CEmployee> se{c_em, ChangeSource::Desc(1)};
  Source<CSchedule, CTarget> st{c_tg, ChangeSource::Desc(2)};
  auto two = [] { return CSchedule{{{OptVal(0)}, {OptVal(0)}}, {{0}, {1}}, {{0}, {1}}}; };
  {
    C c("complemented", Impact::Penalty, ss, se, st, ka, kb, tf, gk, vf, kt, df, gw, false);
    auto s = two();
    CHECK(c.match_count(s) == 2);
    CHECK(c.evaluate(s) == SoftScore::of(-7));
  }
  {
    C c("complemented", Impact::Penalty, ss, se, st, ka, kb, tf, gk, vf, kt, df, gw, false);
    auto s = two();
    SoftScore total = c.initialize(s);
    CHECK(total == SoftScore::of(-7));
    total = total + c.on_retract(s, 1, 0);
    s.shifts[1].employee_id = OptVal(1);
    total = total + c.on_insert(s, 1, 0);
    CHECK(total == SoftScore::of(-2));
    CHECK(total == c.evaluate(s));
  }
  {
    C c("complemented", Impact::Penalty, ss, se, st, ka, kb, tf, gk, vf, kt, df, gw, false);
    auto s = two();
    SoftScore total = c.initialize(s);
    total = total + c.on_retract(s, 0, 1);
    s.employees[0].id = 2;
    total = total + c.on_insert(s, 0, 1);
    CHECK(total == SoftScore::of(-10));
    CHECK(total == c.evaluate(s));
  }
  {
    C c("complemented", Impact::Penalty, ss, se, st, ka, kb, tf, gk, vf, kt, df, gw, false);
    auto s = two();
    SoftScore total = c.initialize(s);
    s.targets.push_back({2});
    total = total + c.on_insert(s, 2, 2);
    CHECK(total == SoftScore::of(-12));
    CHECK(total == c.evaluate(s));
  }
  {
    // :191-215 filtered join + filtered complement sources: -6 -> -10
    Source<CSchedule, CEmployee> sef{c_em, ChangeSource::Desc(1),
                                     [](const CSchedule&, const CEmployee& e) { return e.id != 0; }};
    Source<CSchedule, CTarget> stf{c_tg, ChangeSource::Desc(2),
                                   [](const CSchedule&, const CTarget& t) { return t.employee_id != 2; }};
    C c("filtered complemented", Impact::Penalty, ss, sef, stf, ka, kb, tf, gk, vf, kt, df, gw, false);
    CSchedule s{{{OptVal(0)}, {OptVal(1)}}, {{0}, {1}}, {{0}, {1}, {2}}};
    CHECK(c.match_count(s) == 2);
    CHECK(c.evaluate(s) == SoftScore::of(-6));
    SoftScore total = c.initialize(s);
    CHECK(total == SoftScore::of(-6));
    total = total + c.on_retract(s, 1, 1);
    s.employees[1].id = 0;
    total = total + c.on_insert(s, 1, 1);
    CHECK(total == SoftScore::of(-10));
    CHECK(total == c.evaluate(s));
  }
}

// ---------------------------------------------------------------- projected rows (single emit)
struct PWork {
  size_t bucket;
  int64_t demand;
  bool enabled = true;
};
struct PCapacity {
  size_t bucket;
};
struct PPlan {
  std::vector<PWork> work;
  std::vector<PCapacity> capacity;
};
static const std::vector<PWork>& pp_work(const PPlan& s) { return s.work; }
static const std::vector<PCapacity>& pp_cap(const PPlan& s) { return s.capacity; }
static void kat_projected() {
  // solverforge-scoring/src/constraint/tests/projected/updates.rs:196-278: Work -> Entry{bucket, delta}
  // (MAX_EMITS = 1), group_by(bucket, sum(delta)).complement(capacity buckets, default 3)
  // .penalize(bucket*10 + demand). A single-emit projection is a per-row map, so the grouped join over
  // (work, capacity bucket) with a key-dependent weight restates it.
  auto ka = [](const PWork& w) { return w.bucket; };
  auto kb = [](const PCapacity& c) { return c.bucket; };
  auto tf = [](const PPlan&, const PWork&, const PCapacity&, size_t, size_t) { return true; };
  auto gk = [](const PWork&, const PCapacity& c) { return c.bucket; };
  auto vf = [](const PWork& w, const PCapacity&) { return w.demand; };
  auto kt = [](const PCapacity& c) { return c.bucket; };
  auto df = [](const PCapacity&) { return (int64_t)3; };
  auto gw = [](const size_t& bucket, const int64_t& demand) { return SoftScore::of((int64_t)bucket * 10 + demand); };
  CrossComplementedGroupedConstraint<PPlan, PWork, PCapacity, PCapacity, size_t, size_t, SoftScore, SumAcc, decltype(ka),
                                     decltype(kb), decltype(tf), decltype(gk), decltype(vf), decltype(kt), decltype(df),
                                     decltype(gw)>
      c("projected demand by capacity bucket", Impact::Penalty, {pp_work, ChangeSource::Desc(0)},
        {pp_cap, ChangeSource::Desc(1)}, {pp_cap, ChangeSource::Desc(1)}, ka, kb, tf, gk, vf, kt, df, gw, false);
  PPlan plan{{{0, 5}}, {{0}, {1}}};
  CHECK(c.match_count(plan) == 2);
  CHECK(c.evaluate(plan) == SoftScore::of(-18));
  SoftScore total = c.initialize(plan);
  CHECK(total == SoftScore::of(-18));
  total = total + c.on_retract(plan, 0, 0);
  plan.work[0].demand = 7;
  total = total + c.on_insert(plan, 0, 0);
  CHECK(total == SoftScore::of(-20));
  CHECK(total == c.evaluate(plan));
}

struct PEntry {
  size_t bucket;
  int64_t delta;
};
static void kat_projected_multi_emit() {
  auto always = [](const PPlan&, const PEntry&) { return true; };
  {  // constraint/tests/projected/self_join.rs:26-55: zero and multiple outputs (WorkTwoEntries, MAX_EMITS = 2)
    auto two = [](const PWork& w, std::vector<PEntry>& out) {
      if (!w.enabled) return;
      out.push_back({w.bucket, w.demand});
      out.push_back({w.bucket + 1, w.demand});
    };
    auto wt = [](const PEntry& e) { return SoftScore::of(e.delta); };
    ProjectedUniConstraint<PPlan, PWork, PEntry, SoftScore, decltype(two), decltype(always), decltype(wt)> c(
        "projected work", Impact::Penalty, {pp_work, ChangeSource::Desc(0)}, two, always, wt, false);
    PPlan plan{{{0, 3, true}, {0, 100, false}}, {}};
    CHECK(c.match_count(plan) == 2);
    CHECK(c.evaluate(plan) == SoftScore::of(-6));
    SoftScore total = c.initialize(plan);
    CHECK(total == SoftScore::of(-6));
    total = total + c.on_retract(plan, 1, 0);  // enabling the second entity emits two more rows
    plan.work[1].enabled = true;
    total = total + c.on_insert(plan, 1, 0);
    CHECK(total == SoftScore::of(-206) && total == c.evaluate(plan));
    total = total + c.on_retract(plan, 0, 1);  // foreign descriptor: ignored (collection_extract.rs:83-93)
    total = total + c.on_insert(plan, 0, 1);
    CHECK(total == SoftScore::of(-206));
  }
  auto one = [](const PWork& w, std::vector<PEntry>& out) { out.push_back({w.bucket, w.demand}); };
  auto kf = [](const PEntry& e) { return e.bucket; };
  auto vf = [](const PEntry& e) { return e.delta; };
  {  // localization.rs:4-36: previous outputs are retracted before an update
    auto gw = [](const size_t&, const int64_t& d) { return SoftScore::of(d > 0 ? d : 0); };
    ProjectedGroupedConstraint<PPlan, PWork, PEntry, size_t, SoftScore, SumAcc, decltype(one), decltype(always),
                               decltype(kf), decltype(vf), decltype(gw)>
        c("demand", Impact::Penalty, {pp_work, ChangeSource::Desc(0)}, one, always, kf, vf, gw, false);
    PPlan plan{{{0, 5, true}}, {}};
    SoftScore total = c.initialize(plan);
    CHECK(total == SoftScore::of(-5));
    total = total + c.on_retract(plan, 0, 0);
    plan.work[0].demand = 2;
    total = total + c.on_insert(plan, 0, 0);
    CHECK(total == SoftScore::of(-2));
    CHECK(total == c.evaluate(plan));
  }
  {  // localization.rs:38-70: the grouped weight can use the key
    auto gw = [](const size_t& b, const int64_t& d) { return SoftScore::of((int64_t)b + (d > 0 ? d : 0)); };
    ProjectedGroupedConstraint<PPlan, PWork, PEntry, size_t, SoftScore, SumAcc, decltype(one), decltype(always),
                               decltype(kf), decltype(vf), decltype(gw)>
        c("key weighted demand", Impact::Penalty, {pp_work, ChangeSource::Desc(0)}, one, always, kf, vf, gw, false);
    PPlan plan{{{2, 3, true}, {4, 1, true}}, {}};
    CHECK(c.evaluate(plan) == SoftScore::of(-10));
  }
  {  // self_join.rs:109-154: projected rows self-join by key, pair filter on (left, right) in coordinate order
    auto pf = [](const PEntry& l, const PEntry& r) { return l.delta < r.delta; };
    auto pw = [](const PEntry&, const PEntry&) { return SoftScore::of(1); };
    ProjectedBiConstraint<PPlan, PWork, PEntry, size_t, SoftScore, decltype(one), decltype(always), decltype(kf),
                          decltype(pf), decltype(pw)>
        c("projected duplicate bucket", Impact::Penalty, {pp_work, ChangeSource::Desc(0)}, one, always, kf, pf, pw, false);
    PPlan plan{{{0, 1, true}, {0, 2, true}, {1, 3, true}}, {}};
    SoftScore total = c.initialize(plan);
    CHECK(c.match_count(plan) == 1);
    CHECK(total == SoftScore::of(-1));
    total = total + c.on_retract(plan, 2, 0);
    plan.work[2].bucket = 0;
    total = total + c.on_insert(plan, 2, 0);
    CHECK(c.match_count(plan) == 3);
    CHECK(total == SoftScore::of(-3));
    CHECK(total == c.evaluate(plan));
  }
  {  // two rows of one entity in the same group (support.rs OrderedWorkEntries) with a non-linear weight:
     // the group is re-scored once per notification, incremental == evaluate
    auto ordered = [](const PWork& w, std::vector<PEntry>& out) {
      out.push_back({w.bucket, w.demand});
      out.push_back({w.bucket, w.demand + 10});
    };
    auto gw = [](const size_t&, const int64_t& d) { return SoftScore::of(d > 12 ? (d - 12) * (d - 12) : 0); };
    ProjectedGroupedConstraint<PPlan, PWork, PEntry, size_t, SoftScore, SumAcc, decltype(ordered), decltype(always),
                               decltype(kf), decltype(vf), decltype(gw)>
        c("ordered", Impact::Penalty, {pp_work, ChangeSource::Desc(0)}, ordered, always, kf, vf, gw, false);
    PPlan plan{{{0, 1, true}, {0, 2, true}, {1, 4, true}}, {}};
    SoftScore total = c.initialize(plan);
    CHECK(total == c.evaluate(plan));
    CHECK(total == SoftScore::of(-((26 - 12) * (26 - 12) + (18 - 12) * (18 - 12))));
    total = total + c.on_retract(plan, 1, 0);
    plan.work[1].bucket = 1;
    total = total + c.on_insert(plan, 1, 0);
    CHECK(total == c.evaluate(plan));
    CHECK(c.match_count(plan) == 2);
  }
}

// stream/collector/tests/collector.rs:266-331 — consecutive_runs
static void kat_runs_collector() {
  RunsAcc empty;
  CHECK(empty.result().runs.empty() && empty.result().point_count == 0 && empty.result().item_count == 0);
  RunsAcc one;
  for (int64_t v : {3, 1, 2}) one.accumulate(v);
  auto r1 = one.result();
  CHECK(r1.runs.size() == 1 && r1.runs[0].start == 1 && r1.runs[0].end == 3 && r1.runs[0].point_count == 3 &&
        r1.runs[0].item_count == 3);
  RunsAcc many;
  for (int64_t v : {8, 1, 2, 4, 5, 10}) many.accumulate(v);
  auto r2 = many.result();
  CHECK(r2.runs.size() == 4);
  CHECK(r2.runs[0].start == 1 && r2.runs[0].end == 2 && r2.runs[1].start == 4 && r2.runs[1].end == 5);
  CHECK(r2.runs[2].start == 8 && r2.runs[2].end == 8 && r2.runs[3].start == 10 && r2.runs[3].end == 10);
  RunsAcc dup;
  for (int64_t v : {1, 1, 2, 4, 4, 4}) dup.accumulate(v);
  auto r3 = dup.result();
  CHECK(r3.point_count == 3 && r3.item_count == 6);
  CHECK(r3.runs[0].point_count == 2 && r3.runs[0].item_count == 3 && r3.runs[1].point_count == 1 &&
        r3.runs[1].item_count == 3);
  dup.retract(1);  // one of the two items at point 1: the point stays
  CHECK(dup.result().runs[0].point_count == 2 && dup.result().runs[0].item_count == 2);
  dup.retract(1);  // the point goes: the run shrinks to {2}
  CHECK(dup.result().runs[0].start == 2 && dup.result().runs[0].point_count == 1);
  dup.retract(99);  // unknown point: ignored (:151-154)
  CHECK(dup.result().item_count == 4);
}

// ---------------------------------------------------------------- selectors / foragers
static void kat_nearby_sort() {
  // solverforge-solver/src/heuristic/selector/nearby_list_support.rs:51-73
  for (size_t len = 0; len < 96; ++len) {
    uint32_t state = 0x9E3779B9u ^ (uint32_t)len;
    std::vector<NearbyCandidate> cands;
    for (size_t i = 0; i < len; ++i) {
      state = state * 1664525u + 1013904223u;
      double dist = i % 17 == 0 ? -0.0 : (double)(state % 11);
      cands.push_back({i / 7, i, dist});
    }
    for (size_t k = 0; k <= len + 2; ++k) {
      auto expected = cands;
      std::stable_sort(expected.begin(), expected.end(),
                       [](const NearbyCandidate& l, const NearbyCandidate& r) { return l.distance < r.distance; });
      if (expected.size() > k) expected.resize(k);
      auto actual = cands;
      sort_and_limit_nearby_candidates(actual, k);
      bool same = actual.size() == expected.size();
      for (size_t i = 0; same && i < actual.size(); ++i)
        same = actual[i].entity == expected[i].entity && actual[i].position == expected[i].position;
      CHECK(same);
    }
  }
}

static void kat_moves_and_loop() {
  // list_kernel/change.rs:46-75 doability + do/undo round trip on a tiny CVRP (score invariant
  // incremental == evaluate_all, scope/solver/scope_core.rs:642-653).
  auto pd = std::make_shared<ProblemData>();
  pd->capacity = 3;
  pd->depot = 0;
  pd->demands = {0, 1, 2, 3, 1};
  pd->distance_matrix = {{0, 2, 3, 4, 5}, {2, 0, 6, 7, 8}, {3, 6, 0, 9, 1}, {4, 7, 9, 0, 2}, {5, 8, 1, 2, 0}};
  CvrpPlan plan;
  plan.shared = pd;
  for (size_t i = 1; i <= 4; ++i) plan.customers.push_back({i});
  plan.routes = {{0, {1, 2}, pd.get()}, {1, {3}, pd.get()}, {2, {}, pd.get()}};
  CvrpModel m(plan);
  // customer 4 unassigned: -1 hard; route 1 load 3 (ok), route 0 load 3 (ok); distance 2+6+3 + 4+4 = 19
  CHECK(m.calculate_score() == Sc::of(-1, -19));
  CHECK(m.fresh_score() == Sc::of(-1, -19));
  CHECK(!is_doable(Move::list_change(0, 0, 0, 0, 0), m.dir));
  CHECK(!is_doable(Move::list_change(0, 0, 0, 0, 1), m.dir));
  CHECK(is_doable(Move::list_change(0, 0, 0, 0, 2), m.dir));
  CHECK(!is_doable(Move::list_change(0, 0, 2, 1, 0), m.dir));
  CHECK(!is_doable(Move::list_change(0, 0, 0, 1, 2), m.dir));
  CHECK(is_doable(Move::list_change(0, 0, 0, 2, 0), m.dir));
  auto ev = m.evaluate(Move::list_change(0, 0, 1, 1, 1));  // move customer 2 after 3: load 5 > 3 => -2 hard
  CHECK(ev.kind == EvalKind::Scored);
  CHECK(ev.score == Sc::of(-1 - 2, -(2 + 2 + 4 + 9 + 3)));
  CHECK(m.calculate_score() == Sc::of(-1, -19));
  m.apply(Move::list_change(0, 0, 1, 1, 1));
  CHECK(m.calculate_score() == m.fresh_score());
  CHECK(m.calculate_score() == Sc::of(-3, -20));
  {  // heuristic/selector/tests/nearby_list.rs:302-347: nearby list swap, unique pairs in stable order
    CvrpPlan p2;
    p2.shared = pd;
    p2.customers = plan.customers;
    p2.routes = {{0, {1, 2}, pd.get()}, {1, {3, 4}, pd.get()}};
    CvrpModel m2(p2);
    auto equal = [](const CvrpPlan&, size_t, size_t, size_t, size_t) { return 1.0; };
    auto mv = enumerate_nearby_list_swap_moves(m2.dir.working, m2.dir.access, 0, 4, MoveStreamContext{}, equal);
    const size_t want[6][4] = {{0, 0, 0, 1}, {0, 0, 1, 0}, {0, 0, 1, 1}, {0, 1, 1, 0}, {0, 1, 1, 1}, {1, 0, 1, 1}};
    CHECK(mv.size() == 6);
    for (size_t i = 0; i < 6 && i < mv.size(); ++i)
      CHECK(mv[i].a == want[i][0] && mv[i].b == want[i][1] && mv[i].c == want[i][2] && mv[i].d == want[i][3]);
  }
  {  // heuristic/move/tests/list_reverse.rs:63-160: segment reversal, doability; selector order
     // heuristic/selector/tests/list_precedence.rs:215-223 (len 3: (0,2), (0,3), (1,3))
    CvrpPlan p3;
    p3.shared = pd;
    p3.customers = plan.customers;
    p3.routes = {{0, {1, 2, 3, 4}, pd.get()}, {1, {}, pd.get()}};
    CvrpModel m3(p3);
    CHECK(is_doable(Move::list_reverse(0, 0, 1, 4), m3.dir));
    CHECK(!is_doable(Move::list_reverse(0, 0, 1, 2), m3.dir));   // single element
    CHECK(!is_doable(Move::list_reverse(0, 0, 1, 10), m3.dir));  // out of bounds
    const Sc before = m3.calculate_score();
    auto evr = m3.evaluate(Move::list_reverse(0, 0, 0, 4));
    CHECK(evr.kind == EvalKind::Scored);
    CHECK(m3.calculate_score() == before);  // undo restores the cached score
    m3.apply(Move::list_reverse(0, 0, 1, 4));
    CHECK((m3.dir.working.routes[0].visits == std::vector<size_t>{1, 4, 3, 2}));
    CHECK(m3.calculate_score() == m3.fresh_score());
    m3.apply(Move::list_reverse(0, 0, 0, 4));
    CHECK((m3.dir.working.routes[0].visits == std::vector<size_t>{2, 3, 4, 1}));
    CHECK(m3.calculate_score() == m3.fresh_score());
    CvrpPlan p4 = p3;
    p4.routes = {{0, {1, 2, 3}, pd.get()}};
    CvrpModel m4(p4);
    auto rv = m4.enumerate_list_reverse({});
    const size_t wantr[3][2] = {{0, 2}, {0, 3}, {1, 3}};
    CHECK(rv.size() == 3);
    for (size_t i = 0; i < 3 && i < rv.size(); ++i) CHECK(rv[i].a == 0 && rv[i].b == wantr[i][0] && rv[i].c == wantr[i][1]);
  }
  {  // heuristic/move/tests/sublist_change.rs:80-266 (relocation forward / backward / inter-list, doability)
     // heuristic/selector/tests/sublist_neighborhood.rs:133-187 (canonical segment order)
    CvrpPlan p5;
    p5.shared = pd;
    p5.customers = plan.customers;
    p5.routes = {{0, {1, 2, 3, 4, 5, 6}, pd.get()}, {1, {}, pd.get()}};
    {
      CvrpModel m5(p5);
      const Move fwd = move_sublist_change(0, 0, 1, 3, 0, 4);
      CHECK(is_doable(fwd, m5.dir));
      const Sc before = m5.calculate_score();
      auto ev = m5.evaluate(fwd);
      CHECK(ev.kind == EvalKind::Scored);
      CHECK(m5.calculate_score() == before);  // undo (inverse layout) restores list and cached score
      CHECK((m5.dir.working.routes[0].visits == std::vector<size_t>{1, 2, 3, 4, 5, 6}));
      m5.apply(fwd);
      CHECK((m5.dir.working.routes[0].visits == std::vector<size_t>{1, 4, 5, 6, 2, 3}));
      CHECK(m5.calculate_score() == m5.fresh_score() && m5.calculate_score() == ev.score);
      m5.apply(sublist_change_inverse(fwd));
      CHECK((m5.dir.working.routes[0].visits == std::vector<size_t>{1, 2, 3, 4, 5, 6}));
      m5.apply(move_sublist_change(0, 0, 3, 5, 0, 1));  // backward
      CHECK((m5.dir.working.routes[0].visits == std::vector<size_t>{1, 4, 5, 2, 3, 6}));
      CHECK(m5.calculate_score() == m5.fresh_score());
    }
    {
      CvrpPlan p6 = p5;
      p6.routes = {{0, {1, 2, 3, 4}, pd.get()}, {1, {5, 6}, pd.get()}};
      CvrpModel m6(p6);
      const Move inter = move_sublist_change(0, 0, 1, 3, 1, 1);
      CHECK(is_doable(inter, m6.dir));
      auto ev = m6.evaluate(inter);
      m6.apply(inter);
      CHECK((m6.dir.working.routes[0].visits == std::vector<size_t>{1, 4}));
      CHECK((m6.dir.working.routes[1].visits == std::vector<size_t>{5, 2, 3, 6}));
      CHECK(m6.calculate_score() == m6.fresh_score() && m6.calculate_score() == ev.score);
      m6.apply(sublist_change_inverse(inter));
      CHECK((m6.dir.working.routes[0].visits == std::vector<size_t>{1, 2, 3, 4}));
      CHECK((m6.dir.working.routes[1].visits == std::vector<size_t>{5, 6}));
      // tabu: the undo id of a move is the move id of its inverse (sublist_change.rs tests :270-378)
      auto sg = m6.signature(inter);
      m6.apply(inter);
      auto rs = m6.signature(sublist_change_inverse(inter));
      CHECK(sg.move_id != sg.undo_move_id && sg.undo_move_id == rs.move_id);
    }
    {
      CvrpPlan p7 = p5;
      p7.routes = {{0, {1, 2, 3}, pd.get()}};
      CvrpModel m7(p7);
      CHECK(!is_doable(move_sublist_change(0, 0, 2, 2, 0, 0), m7.dir));   // empty range
      CHECK(!is_doable(move_sublist_change(0, 0, 1, 10, 0, 0), m7.dir));  // out of bounds
      CvrpPlan p8 = p5;
      p8.routes = {{0, {1, 2, 3, 4, 5}, pd.get()}};
      CvrpModel m8(p8);
      CHECK(!is_doable(move_sublist_change(0, 0, 1, 4, 0, 1), m8.dir));   // destination == source start
      CHECK(!is_doable(move_sublist_change(0, 0, 1, 4, 0, 3), m8.dir));   // beyond the post-removal list
      CHECK(is_doable(move_sublist_change(0, 0, 1, 4, 0, 2), m8.dir));
    }
    {
      CvrpPlan p9 = p5;
      p9.routes = {{0, {1, 2, 3}, pd.get()}, {1, {4, 5}, pd.get()}};
      CvrpModel m9(p9);
      auto sv = m9.enumerate_sublist_change(2, 2, {});
      const size_t want[12][5] = {{0, 0, 2, 0, 1}, {0, 0, 2, 1, 0}, {0, 0, 2, 1, 1}, {0, 0, 2, 1, 2},
                                  {0, 1, 3, 0, 0}, {0, 1, 3, 1, 0}, {0, 1, 3, 1, 1}, {0, 1, 3, 1, 2},
                                  {1, 0, 2, 0, 0}, {1, 0, 2, 0, 1}, {1, 0, 2, 0, 2}, {1, 0, 2, 0, 3}};
      CHECK(sv.size() == 12);
      for (size_t i = 0; i < 12 && i < sv.size(); ++i) {
        CHECK(sv[i].a == want[i][0] && sv[i].b == want[i][1] && sv[i].c == want[i][2] && sv[i].d == want[i][3] &&
              sv[i].e == want[i][4]);
        CHECK(is_doable(sv[i], m9.dir));  // sublist_neighborhood.rs:190-218
      }
    }
  }
  {  // heuristic/move/tests/sublist_swap.rs:80-236 (inter / intra exchange, doability), :238-350 (unequal lengths:
     // the undo id is the move id of the inverse layout); selector/tests/sublist_neighborhood.rs:315-364
    CvrpPlan q;
    q.shared = pd;
    q.customers = plan.customers;
    q.routes = {{0, {1, 2, 3, 4}, pd.get()}, {1, {5, 6, 7}, pd.get()}};
    {
      CvrpModel m(q);
      const Move sw = move_sublist_swap(0, 0, 1, 3, 1, 0, 2);
      CHECK(is_doable(sw, m.dir));
      const Sc before = m.calculate_score();
      auto ev = m.evaluate(sw);
      CHECK(ev.kind == EvalKind::Scored && m.calculate_score() == before);
      CHECK((m.dir.working.routes[0].visits == std::vector<size_t>{1, 2, 3, 4}));
      m.apply(sw);
      CHECK((m.dir.working.routes[0].visits == std::vector<size_t>{1, 5, 6, 4}));
      CHECK((m.dir.working.routes[1].visits == std::vector<size_t>{2, 3, 7}));
      CHECK(m.calculate_score() == m.fresh_score() && m.calculate_score() == ev.score);
      m.apply(sublist_swap_inverse(sw));
      CHECK((m.dir.working.routes[0].visits == std::vector<size_t>{1, 2, 3, 4}));
      CHECK((m.dir.working.routes[1].visits == std::vector<size_t>{5, 6, 7}));
      // unequal lengths across lists
      const Move un = move_sublist_swap(0, 0, 1, 2, 1, 0, 3);
      auto sg = m.signature(un);
      auto evu = m.evaluate(un);
      m.apply(un);
      CHECK((m.dir.working.routes[0].visits == std::vector<size_t>{1, 5, 6, 7, 3, 4}));
      CHECK((m.dir.working.routes[1].visits == std::vector<size_t>{2}));
      CHECK(m.calculate_score() == m.fresh_score() && m.calculate_score() == evu.score);
      auto rs = m.signature(sublist_swap_inverse(un));
      CHECK(sg.move_id != sg.undo_move_id && sg.undo_move_id == rs.move_id);
    }
    {
      CvrpPlan q2 = q;
      q2.routes = {{0, {1, 2, 3, 4, 5, 6, 7, 8}, pd.get()}};
      CvrpModel m(q2);
      const Move sw = move_sublist_swap(0, 0, 1, 3, 0, 5, 7);
      CHECK(is_doable(sw, m.dir));
      auto ev = m.evaluate(sw);
      CHECK((m.dir.working.routes[0].visits == std::vector<size_t>{1, 2, 3, 4, 5, 6, 7, 8}));
      m.apply(sw);
      CHECK((m.dir.working.routes[0].visits == std::vector<size_t>{1, 6, 7, 4, 5, 2, 3, 8}));
      CHECK(m.calculate_score() == m.fresh_score() && m.calculate_score() == ev.score);
      m.apply(sublist_swap_inverse(sw));
      CHECK((m.dir.working.routes[0].visits == std::vector<size_t>{1, 2, 3, 4, 5, 6, 7, 8}));
      // unequal lengths inside one list, later segment given first
      const Move un = move_sublist_swap(0, 0, 5, 8, 0, 1, 3);
      auto sg = m.signature(un);
      auto evu = m.evaluate(un);
      CHECK((m.dir.working.routes[0].visits == std::vector<size_t>{1, 2, 3, 4, 5, 6, 7, 8}));
      m.apply(un);
      CHECK((m.dir.working.routes[0].visits == std::vector<size_t>{1, 6, 7, 8, 4, 5, 2, 3}));
      CHECK(m.calculate_score() == m.fresh_score() && m.calculate_score() == evu.score);
      auto rs = m.signature(sublist_swap_inverse(un));
      CHECK(sg.undo_move_id == rs.move_id);
      CHECK(!is_doable(move_sublist_swap(0, 0, 1, 4, 0, 2, 5), m.dir));   // overlapping
      CHECK(!is_doable(move_sublist_swap(0, 0, 1, 1, 0, 2, 3), m.dir));   // empty range
      CHECK(!is_doable(move_sublist_swap(0, 0, 0, 2, 0, 2, 10), m.dir));  // out of bounds
    }
    {
      CvrpPlan q3 = q;
      q3.routes = {{0, {1, 2, 3, 4}, pd.get()}, {1, {5, 6, 7}, pd.get()}};
      CvrpModel m(q3);
      auto sv = m.enumerate_sublist_swap(2, 2, {});
      const size_t want[7][6] = {{0, 0, 2, 0, 2, 4}, {0, 0, 2, 1, 0, 2}, {0, 0, 2, 1, 1, 3}, {0, 1, 3, 1, 0, 2},
                                 {0, 1, 3, 1, 1, 3}, {0, 2, 4, 1, 0, 2}, {0, 2, 4, 1, 1, 3}};
      CHECK(sv.size() == 7);
      for (size_t i = 0; i < 7 && i < sv.size(); ++i) {
        CHECK(sv[i].a == want[i][0] && sv[i].b == want[i][1] && sv[i].c == want[i][2] && sv[i].d == want[i][3] &&
              sv[i].e == want[i][4] && sv[i].f == want[i][5]);
        CHECK(is_doable(sv[i], m.dir));
      }
    }
  }
  {  // constraint/tests/tri_incr.rs:22-138, quad_incr.rs, penta_incr.rs: tuples per team bucket, retract / insert
    struct T { uint32_t team; };
    struct Sol { std::vector<T> tasks; };
    using SS = SoftScore;
    auto mk = [](size_t arity) {
      Source<Sol, T> src{+[](const Sol& s) -> const std::vector<T>& { return s.tasks; }, ChangeSource::Desc(0)};
      auto kf = [](const T& t) { return t.team; };
      auto ff = [](const Sol&, const std::vector<T>&, const NaryTuple&) { return true; };
      auto wf = [](const Sol&, const std::vector<T>&, const NaryTuple&) { return SS::of(1); };
      return SelfJoinNaryConstraint<Sol, T, uint32_t, SS, decltype(kf), decltype(ff), decltype(wf)>(
          "Cluster", Impact::Penalty, arity, src, kf, ff, wf, false);
    };
    {
      auto c = mk(3);
      CHECK(c.evaluate(Sol{{{1}, {1}, {1}, {2}}}) == SS::of(-1));       // tri_incr.rs:22-57
      CHECK(c.evaluate(Sol{{{1}, {1}, {1}, {1}}}) == SS::of(-4));       // :59-94, C(4,3)
      Sol s{{{1}, {1}, {1}}};
      CHECK(c.initialize(s) == SS::of(-1));                             // :96-138
      CHECK(c.on_retract(s, 0, 0) == SS::of(1));
      CHECK(c.on_insert(s, 0, 0) == SS::of(-1));
      CHECK(c.on_insert(s, 0, 1) == SS::of(0));                         // foreign descriptor: not localised
    }
    {
      auto c = mk(4);
      CHECK(c.evaluate(Sol{{{1}, {1}, {1}, {1}, {2}}}) == SS::of(-1));
      CHECK(c.evaluate(Sol{{{1}, {1}, {1}, {1}, {1}}}) == SS::of(-5));  // C(5,4)
      Sol s{{{1}, {1}, {1}, {1}}};
      CHECK(c.initialize(s) == SS::of(-1));
      CHECK(c.on_retract(s, 2, 0) == SS::of(1));
      CHECK(c.on_insert(s, 2, 0) == SS::of(-1));
    }
    {
      auto c = mk(5);
      CHECK(c.evaluate(Sol{{{1}, {1}, {1}, {1}, {1}, {2}}}) == SS::of(-1));
      CHECK(c.evaluate(Sol{{{1}, {1}, {1}, {1}, {1}, {1}}}) == SS::of(-6));  // C(6,5)
      Sol s{{{3}, {3}, {3}, {3}, {3}, {3}}};
      CHECK(c.initialize(s) == SS::of(-6));
      CHECK(c.on_retract(s, 5, 0) == SS::of(5));                        // the C(5,4) tuples that contain row 5
      s.tasks[5].team = 4;
      CHECK(c.on_insert(s, 5, 0) == SS::of(0));
      CHECK(c.evaluate(s) == SS::of(-1));
    }
    // the planning model: incremental score == fresh score along a walk of ChangeMoves
    ClusterPlan cp;
    cp.n_teams = 3;
    for (size_t i = 0; i < 14; ++i) cp.tasks.push_back({i, i % 5 == 4 ? OptVal{} : OptVal{i % 3}});
    ClusterModel cm(cp, {{2, 1}, {3, 10}, {4, 100}, {5, 1000}});
    CHECK(cm.calculate_score() == cm.fresh_score());
    auto mvs = cm.enumerate_scalar({});
    for (size_t i = 0; i < mvs.size(); i += 7) {
      if (!is_doable(mvs[i], cm.dir)) continue;
      auto ev = cm.evaluate(mvs[i]);
      cm.apply(mvs[i]);
      CHECK(cm.calculate_score() == ev.score && cm.calculate_score() == cm.fresh_score());
      mvs = cm.enumerate_scalar({});
    }
  }
  {  // stream/collector/tests/collector.rs:333-395 (indexed_presence) and the macro example
     // solverforge-macros/tests/ui/pass/solverforge_constraints_indexed_presence.rs:57-64 (initialize_all == -2)
    IndexedPresenceAcc acc;
    for (int64_t v : {4, 2, 3, 7, 7}) acc.accumulate(v);
    IndexedPresence r = acc.result();
    CHECK(r.contains(3) && !r.contains(5) && r.count() == 4 && r.item_count() == 5);
    CHECK(r.count_in(2, 5) == 3 && r.any_in(7, 8));
    Runs rr = r.runs();
    CHECK(rr.runs.size() == 2 && rr.runs[0].start == 2 && rr.runs[0].end == 4 && rr.runs[1].start == 7 &&
          rr.runs[1].item_count == 2);
    IndexedPresenceAcc a2;
    for (int64_t v : {0, 2, 5}) a2.accumulate(v);
    Runs cr = a2.result().complement_runs(0, 7);
    CHECK(cr.runs.size() == 3 && cr.runs[0].start == 1 && cr.runs[0].end == 1 && cr.runs[1].start == 3 &&
          cr.runs[1].end == 4 && cr.runs[2].start == 6 && cr.runs[2].end == 6);
    IndexedPresenceAcc a3;
    a3.accumulate(-1);
    a3.accumulate(-1);
    a3.accumulate(0);
    a3.retract(-1);
    CHECK(a3.result().contains(-1) && a3.result().item_count() == 2);
    a3.retract(-1);
    CHECK(!a3.result().contains(-1));
    a3.reset();
    CHECK(a3.result().is_empty() && a3.result().item_count() == 0);
    // the macro example: one nurse works days 0, 1, 2 of a 5-day horizon: streak excess 1 (runs collector form)
    // + rest excess 1 = -2; the model below holds the rest / weekend / days-worked constraints
    ShiftSchedule ss;
    ss.nurses = {{0}, {1}};
    for (size_t i = 0; i < 3; ++i) ss.shifts.push_back({i, (int64_t)i, 0, false, 8, OptVal{0}});
    ShiftModel pm(ss, 0, false, 5);
    // unassigned 0, one-per-day 0, long streaks -1, rest streaks -1, weekend 0, days worked -6, balanced |3-0| + |0-0| = -3
    CHECK(pm.calculate_score() == (Sc{0, -11}));
    CHECK(pm.calculate_score() == pm.fresh_score());
    pm.apply(Move::change(0, 1, OptVal{1}));  // nurse 1 takes day 1: nurse 0 {0, 2}, nurse 1 {1}
    // streaks 0; rest: nurse 0 gaps {1}, {3,4} -> 1, nurse 1 gaps {0}, {2,3,4} -> 2; days worked -6; balanced 2 + 1
    CHECK(pm.calculate_score() == (Sc{0, -12}));
    CHECK(pm.calculate_score() == pm.fresh_score());
  }
  {  // move_selector/swap.rs:64-100: canonical SwapMove order = left-major pairs with left < right; size() = n(n-1)/2
    GraphColoring g4;
    g4.n_colors = 2;
    for (size_t i = 0; i < 4; ++i) g4.nodes.push_back({i, {}, OptVal{i % 2}});
    GraphColoringModel gm(g4);
    auto sw = gm.enumerate_scalar_swap({});
    const size_t want[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    CHECK(sw.size() == 6);
    for (size_t i = 0; i < 6 && i < sw.size(); ++i) CHECK(sw[i].a == want[i][0] && sw[i].b == want[i][1]);
    CHECK(is_doable(sw[0], gm.dir) && !is_doable(sw[1], gm.dir));  // swap.rs:140-157: equal values are not doable
  }
  {  // heuristic/move/k_opt_reconnection_tests.rs:4-47 (pattern counts, no identity) and move/tests/k_opt.rs:86-222
    CHECK(enumerate_reconnections(2).size() == 1 && enumerate_reconnections(2)[0].should_reverse(1));
    auto three = enumerate_reconnections(3);
    CHECK(three.size() == 7);
    const uint8_t t_order[7][4] = {{0, 1, 2, 3}, {0, 1, 2, 3}, {0, 1, 2, 3}, {0, 2, 1, 3}, {0, 2, 1, 3}, {0, 2, 1, 3}, {0, 2, 1, 3}};
    const uint8_t t_mask[7] = {0b0010, 0b0100, 0b0110, 0b0000, 0b0010, 0b0100, 0b0110};  // THREE_OPT_RECONNECTIONS
    for (size_t i = 0; i < 7 && i < three.size(); ++i) {
      CHECK(three[i].reverse_mask == t_mask[i] && three[i].k() == 3);
      for (size_t j = 0; j < 4; ++j) CHECK(three[i].order[j] == t_order[i][j]);
    }
    CHECK(enumerate_reconnections(4).size() == 47 && enumerate_reconnections(5).size() == 383);
    for (size_t k = 2; k <= 5; ++k)
      for (auto& r : enumerate_reconnections(k)) CHECK(!r.is_identity());
    CvrpPlan kp;
    kp.shared = pd;
    kp.customers = plan.customers;
    kp.routes = {{0, {1, 2, 3, 4, 5, 6, 7, 8}, pd.get()}};
    CvrpModel km(kp);
    const Move swap_bc = move_k_opt(0, 0, {2, 4, 6}, three[3]);
    CHECK(is_doable(swap_bc, km.dir));
    const Sc before = km.calculate_score();
    auto ev = km.evaluate(swap_bc);
    CHECK(ev.kind == EvalKind::Scored && km.calculate_score() == before);
    CHECK((km.dir.working.routes[0].visits == std::vector<size_t>{1, 2, 3, 4, 5, 6, 7, 8}));
    km.apply(swap_bc);
    CHECK((km.dir.working.routes[0].visits == std::vector<size_t>{1, 2, 5, 6, 3, 4, 7, 8}));
    CHECK(km.calculate_score() == km.fresh_score() && km.calculate_score() == ev.score);
    CvrpModel km2(kp);
    km2.apply(move_k_opt(0, 0, {2, 4, 6}, three[0]));
    CHECK((km2.dir.working.routes[0].visits == std::vector<size_t>{1, 2, 4, 3, 5, 6, 7, 8}));
    CHECK(km2.calculate_score() == km2.fresh_score());
    CHECK(!is_doable(move_k_opt(0, 0, {2, 4, 10}, three[0]), km2.dir));  // cut beyond the list
    CHECK(!is_doable(move_k_opt(0, 0, {4, 2, 6}, three[0]), km2.dir));   // cuts not increasing
    // heuristic/selector/k_opt/tests.rs:80-170: cut combinations and the whole 3-opt neighbourhood of one tour
    CHECK(binomial(5, 2) == 10 && binomial(7, 3) == 35 && binomial(10, 5) == 252);
    CHECK(count_cut_combinations(3, 8, 1) == 35 && count_cut_combinations(3, 3, 1) == 0);
    std::vector<size_t> cc;
    CHECK(cut_combination_at(3, 8, 1, 0, cc) && cc == (std::vector<size_t>{1, 2, 3}));
    CHECK(cut_combination_at(3, 8, 1, 34, cc) && cc == (std::vector<size_t>{5, 6, 7}));
    CHECK(!cut_combination_at(3, 8, 1, 35, cc));
    {
      // CutCombinationIterator order (iterators.rs:12-91) == rank order
      std::vector<size_t> pos{1, 2, 3};
      for (size_t rank = 0; rank < 35; ++rank) {
        CHECK(cut_combination_at(3, 8, 1, rank, cc) && cc == pos);
        for (int i = 2; i >= 0; --i) {
          const size_t max_pos = 8 - 1 * (3 - (size_t)i);
          if (pos[(size_t)i] < max_pos) {
            pos[(size_t)i] += 1;
            for (size_t j = (size_t)i + 1; j < 3; ++j) pos[j] = pos[j - 1] + 1;
            break;
          }
        }
      }
    }
    CvrpModel km4(kp);
    auto kmoves = enumerate_k_opt_moves(km4.dir.working, km4.dir.access, 0, 3, 1, {});
    CHECK(kmoves.size() == 245);  // 35 cut combinations x 7 patterns
    for (auto& mv : kmoves) CHECK(is_doable(mv, km4.dir));
    for (size_t i = 0; i < kmoves.size(); i += 31) {
      auto evk = km4.evaluate(kmoves[i]);
      CHECK(evk.kind == EvalKind::Scored && km4.calculate_score() == km4.fresh_score());
    }
    // a 2-opt pattern is the segment reversal: same score as ListReverseMove on the same window
    CvrpModel km3(kp);
    auto two = enumerate_reconnections(2);
    auto e2 = km3.evaluate(move_k_opt(0, 0, {2, 6}, two[0]));
    auto er = km3.evaluate(Move::list_reverse(0, 0, 2, 6));
    CHECK(e2.kind == EvalKind::Scored && er.kind == EvalKind::Scored && e2.score == er.score);
  }
  // forager.rs:99-155: first of equal scores kept unless the reservoir pick fires
  BestCandidate<Sc> bc;
  bc.reset(42);
  bc.consider(0, Sc::of(0, -5));
  bc.consider(1, Sc::of(0, -7));
  CHECK(bc.index == 0);
  bc.consider(2, Sc::of(0, -3));
  CHECK(bc.index == 2 && bc.equal_count == 1);
  bc.random_ties = false;
  bc.consider(3, Sc::of(0, -3));
  CHECK(bc.index == 2 && bc.equal_count == 2);
  // phase/localsearch/phase/tests/foraging.rs: accepted-count horizon stops evaluation
  Forager<Sc> fg;
  fg.kind = ForagerKind::AcceptedCount;
  fg.accepted_count_limit = 2;
  Acceptor<Sc> acc;
  acc.kind = AcceptorKind::HillClimbing;
  std::vector<Sc> scores = {Sc::of(0, -12), Sc::of(0, -8), Sc::of(0, -9), Sc::of(0, -1)};
  auto out = replay_step<Sc>(
      scores.size(), [&](size_t i) { return CandidateEvaluation<Sc>{EvalKind::Scored, scores[i]}; }, Sc::of(0, -10),
      Sc::of(0, -10), 7, fg, acc);
  CHECK(out.moves_evaluated == 3);
  CHECK(out.moves_accepted == 2);
  CHECK(out.has_winner && out.winner == 1);
}

// ---------------------------------------------------------------- acceptors
// solverforge-solver/src/phase/localsearch/acceptor/{tests.rs, great_deluge/tests.rs,
// step_counting/tests.rs, diversified_late_acceptance/tests.rs}
static TabuSignature kat_sig(uint64_t scope, std::vector<uint64_t> ent, std::vector<uint64_t> val,
                             std::vector<uint64_t> mv, std::vector<uint64_t> undo) {
  TabuSignature t;
  t.scope = scope;
  t.entity_ids = std::move(ent);
  t.value_ids = std::move(val);
  t.move_id = std::move(mv);
  t.undo_move_id = std::move(undo);
  return t;
}
static Acceptor<SoftScore> kat_tabu(size_t e, size_t v, size_t m, size_t u, bool aspiration) {
  Acceptor<SoftScore> a;
  a.kind = AcceptorKind::TabuSearch;
  a.entity_memory.tenure = e;
  a.value_memory.tenure = v;
  a.move_memory.tenure = m;
  a.reverse_move_memory.tenure = u;
  a.aspiration_enabled = aspiration;
  return a;
}
static void kat_acceptors() {
  auto S = [](int64_t v) { return SoftScore::of(v); };
  {  // tests.rs:82-111 hill climbing; :113-121 late acceptance
    Acceptor<SoftScore> hc;
    hc.kind = AcceptorKind::HillClimbing;
    CHECK(hc.is_accepted(S(-10), S(-5)));
    CHECK(!hc.is_accepted(S(-5), S(-10)));
    CHECK(!hc.is_accepted(S(-5), S(-5)));
    Acceptor<SoftScore> la;
    la.kind = AcceptorKind::LateAcceptance;
    la.phase_started(S(-10), 5);
    CHECK(la.is_accepted(S(-10), S(-5)));
    CHECK(la.is_accepted(S(-10), S(-10)));
    CHECK(!la.is_accepted(S(-10), S(-15)));
  }
  {  // great_deluge/tests.rs:22-77
    Acceptor<SoftScore> gd;
    gd.kind = AcceptorKind::GreatDeluge;
    gd.rain_speed = 0.001;
    gd.phase_started(S(-100));
    CHECK(gd.is_accepted(S(-100), S(-50)));
    CHECK(gd.is_accepted(S(-100), S(-100)));
    CHECK(gd.is_accepted(S(-100), S(-90)));
    CHECK(!gd.is_accepted(S(-100), S(-110)));
    Acceptor<SoftScore> g2;
    g2.kind = AcceptorKind::GreatDeluge;
    g2.rain_speed = 0.1;
    g2.phase_started(S(-100));
    CHECK(g2.is_accepted(S(-100), S(-100)));
    CHECK(!g2.is_accepted(S(-100), S(-101)));
    g2.step_ended(S(-100));  // water -100 + round(100 * 0.1) = -90
    CHECK(g2.is_accepted(S(-90), S(-90)));
    CHECK(!g2.is_accepted(S(-90), S(-91)));
    g2.step_ended(S(-90));
    CHECK(g2.is_accepted(S(-80), S(-80)));
    CHECK(!g2.is_accepted(S(-80), S(-81)));
    g2.phase_ended();
    g2.phase_started(S(-50));
    CHECK(g2.is_accepted(S(-50), S(-50)));
    CHECK(!g2.is_accepted(S(-50), S(-51)));
  }
  {  // step_counting/tests.rs:22-101
    Acceptor<SoftScore> sc;
    sc.kind = AcceptorKind::StepCountingHillClimbing;
    sc.step_count_limit = 5;
    sc.phase_started(S(-100));
    CHECK(sc.is_accepted(S(-100), S(-50)));
    CHECK(sc.is_accepted(S(-100), S(-110)));
    Acceptor<SoftScore> s3;
    s3.kind = AcceptorKind::StepCountingHillClimbing;
    s3.step_count_limit = 3;
    s3.phase_started(S(-100));
    CHECK(s3.is_accepted(S(-100), S(-110)));
    s3.step_ended(S(-110));
    CHECK(s3.is_accepted(S(-110), S(-120)));
    s3.step_ended(S(-120));
    CHECK(s3.is_accepted(S(-120), S(-130)));
    s3.step_ended(S(-130));
    CHECK(!s3.is_accepted(S(-130), S(-140)));
    s3.phase_started(S(-100));  // test_resets_on_improvement
    s3.step_ended(S(-110));
    s3.step_ended(S(-120));
    CHECK(s3.steps_since_improvement == 2);
    s3.step_ended(S(-50));
    CHECK(s3.steps_since_improvement == 0);
    s3.step_ended(S(-60));
    s3.step_ended(S(-70));
    CHECK(s3.is_accepted(S(-70), S(-80)));
    Acceptor<SoftScore> s2;
    s2.kind = AcceptorKind::StepCountingHillClimbing;
    s2.step_count_limit = 2;
    s2.phase_started(S(-100));
    s2.step_ended(S(-110));
    s2.step_ended(S(-120));
    CHECK(!s2.is_accepted(S(-120), S(-130)));
    CHECK(s2.is_accepted(S(-120), S(-50)));
    s2.phase_ended();
    s2.phase_started(S(-200));
    CHECK(s2.steps_since_improvement == 0);
  }
  {  // diversified_late_acceptance/tests.rs:17-77
    Acceptor<SoftScore> d;
    d.kind = AcceptorKind::DiversifiedLateAcceptance;
    d.tolerance = 0.1;
    d.phase_started(S(-100), 5);
    CHECK(d.is_accepted(S(-100), S(-90)));
    d.phase_started(S(-100), 3);
    CHECK(d.is_accepted(S(-90), S(-100)));  // equals the late score
    d.step_ended(S(-80));
    d.step_ended(S(-70));
    d.step_ended(S(-60));
    CHECK(d.is_accepted(S(-60), S(-65)));  // best -60, threshold -60 - round(6.0) = -66
    CHECK(d.is_accepted(S(-60), S(-75)));  // history cycled: late = -80
    Acceptor<SoftScore> d2;
    d2.kind = AcceptorKind::DiversifiedLateAcceptance;
    d2.tolerance = 0.05;
    d2.phase_started(S(-100), 3);
    d2.step_ended(S(-40));
    d2.step_ended(S(-40));
    d2.step_ended(S(-40));
    CHECK(!d2.is_accepted(S(-40), S(-50)));  // threshold -42
  }
  {  // tests.rs:123-137 entity tabu + aspiration
    auto a = kat_tabu(3, 0, 0, 0, true);
    auto first = kat_sig(1, {7}, {}, {10}, {11}), second = kat_sig(1, {7}, {}, {12}, {13});
    a.phase_started(S(-10));
    a.step_ended(S(-9), &first);
    CHECK(!a.is_accepted(S(-9), S(-9), &second));
    CHECK(a.is_accepted(S(-9), S(-5), &second));
  }
  {  // tests.rs:139-150 value tabu
    auto a = kat_tabu(0, 2, 0, 0, false);
    auto first = kat_sig(1, {}, {42}, {10}, {11}), second = kat_sig(1, {}, {42}, {12}, {13});
    a.phase_started(S(-10));
    a.step_ended(S(-9), &first);
    CHECK(!a.is_accepted(S(-9), S(-8), &second));
  }
  {  // tests.rs:152-166 exact move + undo move
    auto a = kat_tabu(0, 0, 2, 2, false);
    auto committed = kat_sig(1, {}, {}, {10, 20}, {30, 40});
    auto exact = kat_sig(1, {}, {}, {10, 20}, {99}), undo = kat_sig(1, {}, {}, {30, 40}, {10, 20});
    a.phase_started(S(-10));
    a.step_ended(S(-9), &committed);
    CHECK(!a.is_accepted(S(-9), S(-8), &exact));
    CHECK(!a.is_accepted(S(-9), S(-8), &undo));
  }
  {  // tests.rs:168-186 undo memory matches the candidate's move identity
    auto a = kat_tabu(0, 0, 0, 2, false);
    auto committed = kat_sig(1, {}, {}, {10}, {20}), reverse = kat_sig(1, {}, {}, {20}, {10});
    auto unrelated = kat_sig(1, {}, {}, {30}, {20});
    a.phase_started(S(-10));
    a.step_ended(S(-9), &committed);
    CHECK(!a.is_accepted(S(-9), S(-8), &reverse));
    CHECK(a.is_accepted(S(-9), S(-8), &unrelated));
  }
  {  // tests.rs:188-201 signatures required; :203-218 memories cleared at phase end
    auto a = kat_tabu(1, 0, 0, 0, true);
    a.phase_started(S(-10));
    bool threw = false;
    try {
      a.is_accepted(S(-10), S(-9), nullptr);
    } catch (const std::logic_error&) {
      threw = true;
    }
    CHECK(threw);
    CHECK(a.requires_move_signatures());
    auto b = kat_tabu(1, 1, 1, 1, false);
    auto sig = kat_sig(1, {1}, {2}, {3}, {4});
    b.phase_started(S(-10));
    b.step_ended(S(-9), &sig);
    CHECK(!b.is_accepted(S(-9), S(-8), &sig));
    b.phase_ended();
    b.phase_started(S(-10));
    CHECK(b.is_accepted(S(-10), S(-9), &sig));
  }
  {  // tests.rs:220-231 move-only policy; :258-270 scopes do not collide
    auto a = kat_tabu(0, 0, 10, 0, true);
    auto committed = kat_sig(1, {}, {}, {10, 20}, {30, 40}), repeated = kat_sig(1, {}, {}, {10, 20}, {99});
    a.phase_started(S(-10));
    a.step_ended(S(-9), &committed);
    CHECK(!a.is_accepted(S(-9), S(-9), &repeated));
    auto b = kat_tabu(2, 2, 0, 0, false);
    auto first = kat_sig(1, {7}, {42}, {10}, {11}), second = kat_sig(2, {7}, {42}, {12}, {13});
    b.phase_started(S(-10));
    b.step_ended(S(-9), &first);
    CHECK(b.is_accepted(S(-9), S(-8), &second));
  }
  {  // Score::multiply / abs (macros.rs:61-71): half away from zero on every level
    CHECK(HardSoftScore::of(-3, 25).multiply(0.5) == HardSoftScore::of(-2, 13));
    CHECK(HardSoftScore::of(-3, 25).abs() == HardSoftScore::of(3, 25));
    CHECK(SoftScore::of(-100).multiply(0.001).v == 0);
    CHECK(SoftScore::of(-15).multiply(0.1).v == -2);
  }
}

// ---------------------------------------------------------------- union scheduler
static void kat_union() {
  // heuristic/selector/decorator/vec_union/tests.rs:204-297: children hold the listed values
  auto drain = [](std::vector<std::vector<int>> kids, UnionOrder order, MoveStreamContext ctx) {
    std::vector<size_t> sizes;
    for (auto& k : kids) sizes.push_back(k.size());
    std::vector<int> out;
    for (auto& pr : union_pull_order(sizes, order, ctx, std::vector<uint64_t>(kids.size(), 1))) out.push_back(kids[pr.first][pr.second]);
    return out;
  };
  CHECK((drain({{1, 2, 3}, {10, 11}}, UnionOrder::Sequential, MoveStreamContext{}) == std::vector<int>{1, 2, 3, 10, 11}));
  CHECK((drain({{1, 2, 3}, {}, {10}, {20, 21}}, UnionOrder::RoundRobin, MoveStreamContext{}) ==
         std::vector<int>{1, 10, 20, 2, 21, 3}));
  {
    MoveStreamContext ctx;
    ctx.step_index = 0;
    ctx.step_seed = 2;
    const size_t offset = ctx.random_index(3, 0xA11CE5E1EC700001ull);  // start_offset of a seeded context
    std::vector<std::vector<int>> expected = {{1, 10, 20, 2, 11, 21}, {10, 20, 1, 11, 21, 2}, {20, 1, 10, 21, 2, 11}};
    ctx.order = SelectionOrder::Random;
    CHECK(drain({{1, 2}, {10, 11}, {20, 21}}, UnionOrder::RotatingRoundRobin, ctx) == expected[offset]);
  }
  {
    MoveStreamContext ctx;
    ctx.step_index = 3;
    ctx.step_seed = 42;
    auto v = drain({{1, 2}, {10, 11}, {20, 21}}, UnionOrder::StratifiedRandom, ctx);
    CHECK(v.size() == 6);
    // smooth weighted round-robin with equal weights: every child once per round, in one strided order
    CHECK((std::set<int>{v[0], v[1], v[2]} == std::set<int>{1, 10, 20}));
    CHECK((std::set<int>{v[3], v[4], v[5]} == std::set<int>{2, 11, 21}));
    CHECK(v[3] / 10 == v[0] / 10 && v[4] / 10 == v[1] / 10 && v[5] / 10 == v[2] / 10);
  }
  {  // uneven children: an exhausted child drops out, the others keep their relative order
    MoveStreamContext ctx;
    ctx.step_seed = 7;
    auto v = drain({{1}, {10, 11, 12}, {20, 21}}, UnionOrder::StratifiedRandom, ctx);
    CHECK(v.size() == 6);
    std::vector<int> c1;
    for (int x : v)
      if (x >= 10 && x < 20) c1.push_back(x);
    CHECK((c1 == std::vector<int>{10, 11, 12}));
    auto w = drain({{1}, {10, 11, 12}, {20, 21}}, UnionOrder::Random, ctx);
    CHECK(w.size() == 6);
  }
}

int main() {
  kat_scores();
  kat_director();
  kat_cross_bi();
  kat_self_join();
  kat_exists();
  kat_grouped();
  kat_cross_complemented();
  kat_projected();
  kat_projected_multi_emit();
  kat_runs_collector();
  kat_nearby_sort();
  kat_moves_and_loop();
  kat_acceptors();
  kat_union();
  if (g_fail) {
    std::printf("KAT FAILED %d of %d\n", g_fail, g_checks);
    return 1;
  }
  std::printf("KAT OK %d\n", g_checks);
  return 0;
}
