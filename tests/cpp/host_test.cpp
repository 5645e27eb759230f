// Tests of the C++ host mirror (solverforge_b200/host/solverforge_gpu.hpp) against the oracle.
//   host_test replay   CPU only: acceptor/forager replay vs the oracle's restatement
//   host_test gpu      needs a B200: ConstraintFactory -> libsfgpu -> scores vs the oracle, bit-exact
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../oracle/models.hpp"
#include "../../solverforge_b200/host/solverforge_gpu.hpp"

static int g_fail = 0, g_checks = 0;
#define CHECK(c)                                                   \
  do {                                                             \
    ++g_checks;                                                    \
    if (!(c)) {                                                    \
      ++g_fail;                                                    \
      std::fprintf(stderr, "FAIL %s:%d %s\n", __FILE__, __LINE__, #c); \
    }                                                              \
  } while (0)

static int test_replay() {
  uint64_t s = 12345;
  for (int trial = 0; trial < 200; ++trial) {
    size_t n = 1 + sfo::splitmix64(s++) % 400;
    std::vector<sf::HardSoftScore> scores(n);
    std::vector<uint8_t> doable(n);
    for (size_t i = 0; i < n; ++i) {
      scores[i] = {-(int64_t)(sfo::splitmix64(s++) % 3), -(int64_t)(sfo::splitmix64(s++) % 6)};
      doable[i] = sfo::splitmix64(s++) % 8 != 0;
    }
    sf::HardSoftScore last{-1, -3};
    uint64_t seed = sfo::splitmix64(s++);
    for (int fk = 0; fk < 5; ++fk)
      for (int ak = 0; ak < 2; ++ak)
        for (int ties = 0; ties < 2; ++ties) {
          sf::ForagerConfig fc;
          fc.kind = fk == 0 ? sf::ForagerConfig::AcceptedCount : fk == 1 ? sf::ForagerConfig::FirstAccepted : fk == 2 ? sf::ForagerConfig::BestScore : fk == 3 ? sf::ForagerConfig::FirstBestScoreImproving : sf::ForagerConfig::FirstLastStepScoreImproving;
          fc.has_improving_limit = (trial % 2) == 1;
          fc.accepted_count_limit = 1 + trial % 9;
          fc.random_ties = ties;
          sf::HillClimbingAcceptor hc;
          sf::LateAcceptanceAcceptor la(4);
          la.phase_started({-1, -5});
          sf::Acceptor& acc = ak == 0 ? (sf::Acceptor&)hc : (sf::Acceptor&)la;
          const sf::HardSoftScore best_so_far{-1, -2};
          // improvement gates on some trials (evaluation.rs:76-111)
          std::vector<uint8_t> gates(n, 0);
          const bool gated = trial % 3 == 0;
          if (gated)
            for (size_t i = 0; i < n; ++i) gates[i] = (uint8_t)(sfo::splitmix64(seed + i) % 4);
          auto got = sf::replay_step(scores.data(), doable.data(), n, best_so_far, last, seed, fc, acc, nullptr,
                                     gated ? gates.data() : nullptr);
          sfo::Forager<sfo::Sc> of;
          of.kind = fk == 0 ? sfo::ForagerKind::AcceptedCount : fk == 1 ? sfo::ForagerKind::FirstAccepted : fk == 2 ? sfo::ForagerKind::BestScore : fk == 3 ? sfo::ForagerKind::FirstBestScoreImproving : sfo::ForagerKind::FirstLastStepScoreImproving;
          of.has_improving_limit = fc.has_improving_limit;
          of.accepted_count_limit = fc.accepted_count_limit;
          of.best.random_ties = ties;
          sfo::Acceptor<sfo::Sc> oa;
          oa.kind = ak == 0 ? sfo::AcceptorKind::HillClimbing : sfo::AcceptorKind::LateAcceptance;
          if (ak == 1) {
            oa.history.assign(4, sfo::Sc::of(-1, -5));
            oa.history_idx = 0;
          }
          auto want = sfo::replay_step<sfo::Sc>(
              n,
              [&](size_t i) {
                const sfo::Sc sc = sfo::Sc::of(scores[i].hard, scores[i].soft), ol = sfo::Sc::of(last.hard, last.soft);
                sfo::EvalKind k = doable[i] ? sfo::EvalKind::Scored : sfo::EvalKind::NotDoable;
                if (k == sfo::EvalKind::Scored && gated) {
                  if ((gates[i] & 1) && sfo::hard_score_delta(ol, sc) != sfo::HardDelta::Improving)
                    k = sfo::EvalKind::RejectedByHardImprovement;
                  else if ((gates[i] & 2) && sc <= ol)
                    k = sfo::EvalKind::RejectedByScoreImprovement;
                }
                return sfo::CandidateEvaluation<sfo::Sc>{k, sc};
              },
              sfo::Sc::of(best_so_far.hard, best_so_far.soft), sfo::Sc::of(last.hard, last.soft), seed, of, oa);
          CHECK(got.has_winner == want.has_winner);
          if (want.has_winner) CHECK(got.winner == want.winner);
          CHECK(got.moves_evaluated == want.moves_evaluated);
          CHECK(got.score_calculations == want.score_calculations);
          CHECK(got.moves_accepted == want.moves_accepted);
        }
  }
  return 0;
}

// Multi-step trajectories: every stateful acceptor of the host mirror against the oracle's restatement
// on the same random candidate streams (scores, doability, tabu signatures), including the device
// form (acceptor code + threshold) that the fused GPU steps take.
static bool device_accepts(const sf::DeviceAcceptance& d, sf::HardSoftScore last, sf::HardSoftScore mv) {
  switch (d.acceptor) {
    case 0: return true;
    case 1: return mv > last;
    case 2: return mv >= last || mv >= d.threshold;
    default: return mv > last || mv >= d.threshold;
  }
}
static int test_acceptor_trajectories() {
  uint64_t s = 777;
  for (int kind = 0; kind < 6; ++kind)
    for (int trial = 0; trial < 30; ++trial) {
      sf::HillClimbingAcceptor hc;
      sf::LateAcceptanceAcceptor la(3 + trial % 5);
      sf::GreatDelugeAcceptor gd(trial % 2 ? 0.05 : 0.001);
      sf::StepCountingHillClimbingAcceptor sc(1 + trial % 4);
      sf::DiversifiedLateAcceptanceAcceptor dla(2 + trial % 6, trial % 3 ? 0.1 : 0.01);
      sf::TabuSearchAcceptor tabu(trial % 4, (trial / 2) % 3, 1 + trial % 3, trial % 2 ? 2 : 0, trial % 5 != 0);
      sf::Acceptor* accs[6] = {&hc, &la, &gd, &sc, &dla, &tabu};
      sf::Acceptor& acc = *accs[kind];
      sfo::Acceptor<sfo::Sc> oa;
      const sfo::AcceptorKind kinds[6] = {sfo::AcceptorKind::HillClimbing, sfo::AcceptorKind::LateAcceptance,
                                          sfo::AcceptorKind::GreatDeluge, sfo::AcceptorKind::StepCountingHillClimbing,
                                          sfo::AcceptorKind::DiversifiedLateAcceptance, sfo::AcceptorKind::TabuSearch};
      oa.kind = kinds[kind];
      oa.rain_speed = gd.rain_speed;
      oa.step_count_limit = sc.step_count_limit;
      oa.tolerance = dla.tolerance;
      oa.entity_memory.tenure = tabu.entity_memory.tenure;
      oa.value_memory.tenure = tabu.value_memory.tenure;
      oa.move_memory.tenure = tabu.move_memory.tenure;
      oa.reverse_move_memory.tenure = tabu.reverse_move_memory.tenure;
      oa.aspiration_enabled = tabu.aspiration_enabled;
      sf::HardSoftScore last{-2, -(int64_t)(40 + trial)};
      sf::HardSoftScore best = last;
      acc.phase_started(last);
      oa.phase_started(sfo::Sc::of(last.hard, last.soft), kind == 1 ? la.size : dla.size);
      for (int step = 0; step < 40; ++step) {
        const size_t n = 1 + sfo::splitmix64(s++) % 60;
        std::vector<sf::HardSoftScore> scores(n);
        std::vector<uint8_t> doable(n);
        std::vector<sf::MoveTabuSignature> sigs(n);
        std::vector<sfo::TabuSignature> osigs(n);
        for (size_t i = 0; i < n; ++i) {
          scores[i] = {last.hard + (int64_t)(sfo::splitmix64(s++) % 3) - 1, last.soft + (int64_t)(sfo::splitmix64(s++) % 21) - 11};
          doable[i] = sfo::splitmix64(s++) % 8 != 0;
          const uint64_t e = sfo::splitmix64(s++) % 6, from = sfo::splitmix64(s++) % 4, to = sfo::splitmix64(s++) % 4;
          sigs[i] = sf::change_signature(0, e, (int64_t)from - 1, (int64_t)to - 1);
          osigs[i].scope = sigs[i].scope;
          osigs[i].entity_ids = sigs[i].entity_ids;
          osigs[i].value_ids = sigs[i].destination_value_ids;
          osigs[i].move_id = sigs[i].move_id;
          osigs[i].undo_move_id = sigs[i].undo_move_id;
        }
        const uint64_t seed = sfo::splitmix64(s++);
        sf::ForagerConfig fc;
        fc.kind = step % 3 == 0 ? sf::ForagerConfig::AcceptedCount : sf::ForagerConfig::BestScore;
        fc.accepted_count_limit = 1 + step % 7;
        sfo::Forager<sfo::Sc> of;
        of.kind = step % 3 == 0 ? sfo::ForagerKind::AcceptedCount : sfo::ForagerKind::BestScore;
        of.accepted_count_limit = fc.accepted_count_limit;
        // device form agrees with is_accepted for every candidate (before the replay mutates nothing)
        const sf::DeviceAcceptance df = acc.device_form(last);
        CHECK(df.expressible == (kind != 5));
        if (df.expressible)
          for (size_t i = 0; i < n; ++i) CHECK(device_accepts(df, last, scores[i]) == acc.is_accepted(last, scores[i], &sigs[i]));
        auto got = sf::replay_step(scores.data(), doable.data(), n, best, last, seed, fc, acc,
                                   [&](size_t i) { return sigs[i]; });
        auto want = sfo::replay_step<sfo::Sc>(
            n,
            [&](size_t i) {
              return sfo::CandidateEvaluation<sfo::Sc>{doable[i] ? sfo::EvalKind::Scored : sfo::EvalKind::NotDoable,
                                                       sfo::Sc::of(scores[i].hard, scores[i].soft)};
            },
            sfo::Sc::of(best.hard, best.soft), sfo::Sc::of(last.hard, last.soft), seed, of, oa,
            [&](size_t i) { return &osigs[i]; });
        CHECK(got.has_winner == want.has_winner);
        CHECK(got.moves_accepted == want.moves_accepted);
        CHECK(got.moves_evaluated == want.moves_evaluated);
        if (want.has_winner) {
          CHECK(got.winner == want.winner);
          last = scores[got.winner];
          if (last > best) best = last;
          acc.step_ended(last, &sigs[got.winner]);
          oa.step_ended(sfo::Sc::of(last.hard, last.soft), &osigs[want.winner]);
        } else {
          acc.step_ended(last, nullptr);
          oa.step_ended(sfo::Sc::of(last.hard, last.soft), nullptr);
        }
      }
      acc.phase_ended();
      oa.phase_ended();
    }
  return 0;
}

static int test_gpu() {
  // graph colouring, 400 nodes, through the C++ ConstraintFactory
  const uint32_t n = 400, k = 5;
  std::vector<std::vector<uint32_t>> adj(n);
  uint64_t s = 7;
  for (int e = 0; e < 1500; ++e) {
    uint32_t a = sfo::splitmix64(s++) % n, b = sfo::splitmix64(s++) % n;
    if (a == b) continue;
    adj[a].push_back(b);
    adj[b].push_back(a);
  }
  std::vector<uint32_t> rp(n + 1, 0), ci;
  sfo::GraphColoring g;
  g.n_colors = k;
  g.nodes.resize(n);
  std::vector<int32_t> color(n);
  for (uint32_t i = 0; i < n; ++i) {
    color[i] = (int32_t)(sfo::splitmix64(s++) % (k + 1)) - 1;
    g.nodes[i] = {i, {}, color[i] < 0 ? sfo::OptVal() : sfo::OptVal((size_t)color[i])};
    for (uint32_t b : adj[i]) {
      ci.push_back(b);
      g.nodes[i].neighbors.push_back(b);
    }
    rp[i + 1] = (uint32_t)ci.size();
  }
  sfo::GraphColoringModel oracle(g);
  sf::GpuScoreDirector d(1);
  d.add_collection("colors", k, -1);
  uint32_t nodes = d.add_collection("nodes", n, 0);
  d.add_scalar_variable(nodes, "color_idx", k, true);
  uint32_t csr = d.add_csr("neighbors", rp, ci);
  sf::ConstraintFactory f(d);
  f.for_each(nodes).unassigned().penalize(sf::HardSoftScore::ONE_HARD()).named("Unassigned color");
  f.for_each(nodes).join(f.for_each(nodes), sf::AdjacentEqual{csr}).penalize(sf::HardSoftScore::ONE_HARD()).named("Adjacent color conflict");
  d.set_scalar_state(color);
  auto init = d.commit();
  auto os = oracle.calculate_score();
  CHECK(init[0].hard == os.hard && init[0].soft == os.soft);
  auto moves = oracle.enumerate_scalar({});
  std::vector<sf::ScalarEdit> batch;
  for (auto& m : moves) batch.push_back({(uint32_t)m.a, m.to ? (int32_t)*m.to : -1});
  std::vector<sf::HardSoftScore> scores;
  std::vector<uint8_t> doable;
  d.score_candidates(batch, {0, batch.size()}, scores, doable);
  for (size_t i = 0; i < moves.size(); ++i) {
    auto ev = oracle.evaluate(moves[i]);
    bool ok = ev.kind != sfo::EvalKind::NotDoable;
    CHECK(ok == (doable[i] != 0));
    if (ok) CHECK(scores[i].hard == ev.score.hard && scores[i].soft == ev.score.soft);
  }
  sf::HillClimbingAcceptor hc;
  sf::ForagerConfig fc;
  auto out = sf::replay_step(scores.data(), doable.data(), scores.size(), init[0], init[0], 99, fc, hc);
  if (out.has_winner) {
    d.apply(std::vector<sf::ScalarEdit>{batch[out.winner]});
    oracle.apply(moves[out.winner]);
    auto c = d.calculate_score()[0], fr = d.fresh_score()[0];
    auto o2 = oracle.calculate_score();
    CHECK(c == out.score && c == fr);
    CHECK(c.hard == o2.hard && c.soft == o2.soft);
  }
  // Tabu search over GPU scores: the acceptor needs per-move signatures, so it replays on the host
  // over the materialised batch (tabu_search.rs:170-235) — same trajectory as the oracle's tabu
  // acceptor over the oracle's own scores and signatures.
  {
    sf::TabuSearchAcceptor tabu(3, 0, 4, 4, true);
    sfo::Acceptor<sfo::Sc> ot;
    ot.kind = sfo::AcceptorKind::TabuSearch;
    ot.entity_memory.tenure = 3;
    ot.move_memory.tenure = 4;
    ot.reverse_move_memory.tenure = 4;
    auto start = d.calculate_score()[0];
    tabu.phase_started(start);
    ot.phase_started(sfo::Sc::of(start.hard, start.soft));
    sf::HardSoftScore best = start;
    std::vector<int32_t> cur = d.scalar_state(n)[0];
    for (int step = 0; step < 12; ++step) {
      auto mv = oracle.enumerate_scalar({});
      std::vector<sf::ScalarEdit> b2;
      for (auto& m : mv) b2.push_back({(uint32_t)m.a, m.to ? (int32_t)*m.to : -1});
      d.score_candidates(b2, {0, b2.size()}, scores, doable);
      auto last = d.calculate_score()[0];
      sf::ForagerConfig fc2;
      fc2.kind = sf::ForagerConfig::AcceptedCount;
      fc2.accepted_count_limit = 50;
      auto got = sf::replay_step(scores.data(), doable.data(), scores.size(), best, last, 500 + step, fc2, tabu,
                                 [&](size_t i) { return sf::change_signature(0, b2[i].entity_index, cur[b2[i].entity_index], b2[i].to_value); });
      sfo::Forager<sfo::Sc> of;
      of.kind = sfo::ForagerKind::AcceptedCount;
      of.accepted_count_limit = 50;
      std::vector<sfo::TabuSignature> osig(mv.size());
      for (size_t i = 0; i < mv.size(); ++i) osig[i] = oracle.signature(mv[i]);
      auto want = sfo::replay_step<sfo::Sc>(
          mv.size(), [&](size_t i) { return oracle.evaluate(mv[i]); }, sfo::Sc::of(best.hard, best.soft),
          sfo::Sc::of(last.hard, last.soft), 500 + step, of, ot, [&](size_t i) { return &osig[i]; });
      CHECK(got.has_winner == want.has_winner);
      CHECK(got.moves_accepted == want.moves_accepted);
      if (!want.has_winner) break;
      CHECK(got.winner == want.winner);
      const sf::MoveTabuSignature sig = sf::change_signature(0, b2[got.winner].entity_index, cur[b2[got.winner].entity_index], b2[got.winner].to_value);
      d.apply(std::vector<sf::ScalarEdit>{b2[got.winner]});
      cur[b2[got.winner].entity_index] = b2[got.winner].to_value;
      oracle.apply(mv[want.winner]);
      auto now = d.calculate_score()[0];
      auto onow = oracle.calculate_score();
      CHECK(now.hard == onow.hard && now.soft == onow.soft);
      tabu.step_ended(now, &sig);
      ot.step_ended(onow, &osig[want.winner]);
      if (now > best) best = now;
    }
  }
  // projected scoring rows through the C++ factory (`.project(..)`): roster with shifts spanning days
  {
    const uint32_t n_shifts = 90, n_nurses = 5;
    const int64_t n_days = 7, limit = 9;
    sfo::Roster ro;
    ro.n_nurses = n_nurses;
    ro.n_days = n_days;
    std::vector<uint32_t> ptr(1, 0), days;
    std::vector<int64_t> hours, required(n_shifts);
    std::vector<int32_t> nurse(n_shifts);
    uint64_t q = 4242;
    for (uint32_t i = 0; i < n_shifts; ++i) {
      required[i] = sfo::splitmix64(q++) % 3 != 0;
      nurse[i] = (int32_t)(sfo::splitmix64(q++) % (n_nurses + 1)) - 1;
      sfo::RShift sh{i, required[i] != 0, {}, nurse[i] < 0 ? sfo::OptVal() : sfo::OptVal((size_t)nurse[i])};
      const uint32_t spans = sfo::splitmix64(q++) % 4;
      const int64_t start = sfo::splitmix64(q++) % n_days;
      for (uint32_t j = 0; j < spans; ++j) {
        const int64_t day = std::min<int64_t>(start + (j % 2), n_days - 1), h = 2 + sfo::splitmix64(q++) % 6;
        sh.spans.push_back({day, h});
        days.push_back((uint32_t)day);
        hours.push_back(h);
      }
      ptr.push_back((uint32_t)days.size());
      ro.shifts.push_back(sh);
    }
    sfo::RosterModel orc(ro, limit);
    sf::GpuScoreDirector rd(1);
    rd.add_collection("nurses", n_nurses, -1);
    uint32_t shifts = rd.add_collection("shifts", n_shifts, 0);
    rd.add_scalar_variable(shifts, "nurse_idx", n_nurses, true);
    uint32_t req = rd.add_column(shifts, "required", required);
    uint32_t spans = rd.add_csr("spans", ptr, days);
    sf::Projection rows{spans, (uint32_t)n_days, ptr, hours};
    sf::ConstraintFactory rf(rd);
    rf.for_each(shifts).filtered(req).unassigned().penalize(sf::HardSoftScore::ONE_HARD()).named("Unassigned required shift");
    sf::project(rf.for_each(shifts), rows).group_by(sf::Sum{UINT32_MAX}).penalize(sf::Weight::hard(SFGPU_W_EXCESS, 1, limit)).named("Daily hours");
    sf::project(rf.for_each(shifts), rows).group_by(sf::Count{}).penalize(sf::Weight::soft(SFGPU_W_SQUARE, 1, 0)).named("Fragmented days");
    sf::project(rf.for_each(shifts), rows).penalize_pairs(sf::HardSoftScore::ONE_HARD()).named("Double booking");
    sf::project(rf.for_each(shifts), rows).penalize(sf::Weight::soft(SFGPU_W_LINEAR, 1, 0)).named("Worked hours");
    rd.set_scalar_state(nurse);
    auto ri = rd.commit();
    auto ros = orc.calculate_score();
    CHECK(ri[0].hard == ros.hard && ri[0].soft == ros.soft);
    auto rmoves = orc.enumerate_scalar({});
    std::vector<sf::ScalarEdit> rb;
    for (auto& m : rmoves) rb.push_back({(uint32_t)m.a, m.to ? (int32_t)*m.to : -1});
    rd.score_candidates(rb, {0, rb.size()}, scores, doable);
    for (size_t i = 0; i < rmoves.size(); ++i) {
      auto ev = orc.evaluate(rmoves[i]);
      bool ok = ev.kind != sfo::EvalKind::NotDoable;
      CHECK(ok == (doable[i] != 0));
      if (ok) CHECK(scores[i].hard == ev.score.hard && scores[i].soft == ev.score.soft);
    }
  }
  // Whole local-search phases through the C++ host mirror (sf::LocalSearch: host acceptor state, device
  // steps) on a small CVRP against the oracle's selector + scoring + acceptor + forager, step by step —
  // hill climbing, late acceptance, great deluge, step counting, diversified late acceptance; nearby
  // list change and nearby list swap neighbourhoods.
  {
    const uint32_t n_cust = 60, n_routes = 5, dim = n_cust + 1;
    auto pd = std::make_shared<sfo::ProblemData>();
    pd->depot = 0;
    pd->demands.assign(dim, 0);
    std::vector<int64_t> xs(dim), ys(dim), flat((size_t)dim * dim), demand64(dim, 0);
    uint64_t q = 9001;
    int64_t total = 0;
    for (uint32_t i = 0; i < dim; ++i) {
      xs[i] = sfo::splitmix64(q++) % 300;
      ys[i] = sfo::splitmix64(q++) % 300;
      if (i) {
        pd->demands[i] = 1 + (int32_t)(sfo::splitmix64(q++) % 9);
        demand64[i] = pd->demands[i];
        total += pd->demands[i];
      }
    }
    pd->capacity = total / n_routes + 3;
    pd->distance_matrix.assign(dim, std::vector<int64_t>(dim, 0));
    for (uint32_t i = 0; i < dim; ++i)
      for (uint32_t j = 0; j < dim; ++j) {
        const double dx = (double)(xs[i] - xs[j]), dy = (double)(ys[i] - ys[j]);
        pd->distance_matrix[i][j] = flat[(size_t)i * dim + j] = (int64_t)std::llround(std::sqrt(dx * dx + dy * dy));
      }
    std::vector<std::vector<size_t>> routes(n_routes);
    for (uint32_t c = 1; c < dim; ++c) routes[sfo::splitmix64(q++) % n_routes].push_back(c);
    std::vector<uint32_t> offs(1, 0), el;
    for (auto& rt : routes) {
      for (size_t c : rt) el.push_back((uint32_t)c);
      offs.push_back((uint32_t)el.size());
    }
    std::vector<int64_t> ids;
    for (uint32_t c = 1; c < dim; ++c) ids.push_back(c);
    for (int kind = 0; kind < 5; ++kind)
      for (int swap_moves = 0; swap_moves < 2; ++swap_moves) {
        sfo::CvrpPlan plan;
        plan.shared = pd;
        for (uint32_t c = 1; c < dim; ++c) plan.customers.push_back({c});
        for (uint32_t r = 0; r < n_routes; ++r) plan.routes.push_back({r, routes[r], pd.get()});
        sfo::CvrpModel orc(plan);
        sf::GpuScoreDirector cd(1);
        uint32_t locations = cd.add_collection("locations", dim, -1);
        uint32_t customers = cd.add_collection("customers", n_cust, -1);
        uint32_t rts = cd.add_collection("routes", n_routes, 0);
        cd.add_list_variable(rts, locations, "visits");
        uint32_t cid = cd.add_column(customers, "id", ids);
        uint32_t dem = cd.add_column(locations, "demand", demand64);
        uint32_t mat = cd.add_matrix("distance_matrix", dim, dim, flat, true);
        sf::ConstraintFactory cf(cd);
        cf.for_each(customers).if_not_exists(cf.for_each(rts).flattened(), sf::EqualId{cid}).penalize(sf::HardSoftScore::ONE_HARD()).named("all_customers_assigned");
        cf.for_each(rts).penalize(sf::Weight::hard(SFGPU_W_EXCESS, 1, pd->capacity), sf::ListSum{dem}).named("vehicle_capacity");
        cf.for_each(rts).penalize(sf::Weight::soft(SFGPU_W_LINEAR, 1, 0), sf::PathCost{mat, 0}).named("total_distance");
        cd.set_list_state(offs, el);
        auto c0 = cd.commit();
        auto o0 = orc.calculate_score();
        CHECK(c0[0].hard == o0.hard && c0[0].soft == o0.soft);
        auto make = [&]() -> std::unique_ptr<sf::Acceptor> {
          switch (kind) {
            case 0: return std::make_unique<sf::HillClimbingAcceptor>();
            case 1: return std::make_unique<sf::LateAcceptanceAcceptor>(4);
            case 2: return std::make_unique<sf::GreatDelugeAcceptor>(0.002);
            case 3: return std::make_unique<sf::StepCountingHillClimbingAcceptor>(2);
            default: return std::make_unique<sf::DiversifiedLateAcceptanceAcceptor>(3, 0.01);
          }
        };
        sf::StepParams fp;
        fp.accepted_limit = swap_moves ? 0 : 30;  // AcceptedCount(30) for relocations, BestScore for swaps
        sf::LocalSearch ls(cd, make, fp);
        ls.phase_started();
        sfo::Acceptor<sfo::Sc> oa;
        const sfo::AcceptorKind kinds[5] = {sfo::AcceptorKind::HillClimbing, sfo::AcceptorKind::LateAcceptance,
                                            sfo::AcceptorKind::GreatDeluge, sfo::AcceptorKind::StepCountingHillClimbing,
                                            sfo::AcceptorKind::DiversifiedLateAcceptance};
        oa.kind = kinds[kind];
        oa.rain_speed = 0.002;
        oa.step_count_limit = 2;
        oa.tolerance = 0.01;
        oa.phase_started(o0, kind == 1 ? 4 : 3);
        sfo::Sc obest = o0;
        for (int step = 0; step < 14; ++step) {
          const uint64_t seed = 7000 + step;
          auto res = ls.step(swap_moves ? sf::LocalSearch::NearbyListSwap : sf::LocalSearch::NearbyListChange, {seed}, 8);
          auto mv = swap_moves ? orc.enumerate_list_swap(8, {}) : orc.enumerate_list(8, {});
          sfo::Forager<sfo::Sc> of;
          of.kind = swap_moves ? sfo::ForagerKind::BestScore : sfo::ForagerKind::AcceptedCount;
          of.accepted_count_limit = 30;
          const sfo::Sc olast = orc.calculate_score();
          auto want = sfo::replay_step<sfo::Sc>(
              mv.size(), [&](size_t i) { return orc.evaluate(mv[i]); }, obest, olast, seed, of, oa);
          CHECK(want.has_winner == (res.index[0] != UINT32_MAX));
          if (want.has_winner) {
            CHECK(res.index[0] == want.winner);
            CHECK(res.evaluated[0] == want.moves_evaluated);
            orc.apply(mv[want.winner]);
          }
          const sfo::Sc onow = orc.calculate_score();
          oa.step_ended(onow);
          if (onow > obest) obest = onow;
          auto now = cd.calculate_score()[0];
          CHECK(now.hard == onow.hard && now.soft == onow.soft);
        }
        CHECK(ls.best_scores()[0].hard == obest.hard && ls.best_scores()[0].soft == obest.soft);
        auto fr = cd.fresh_score()[0], cm = cd.calculate_score()[0];
        CHECK(fr == cm);
        // the device-resident loop through the mirror runs and keeps cached == fresh
        if (kind == 0 && !swap_moves) {
          // device-enumerated SublistChange step (sfgpu_step_sublist_change) vs the oracle's selector order and
          // do / score / undo evaluation: BestScore, first of equal scores
          auto moves = orc.enumerate_sublist_change(1, 3, {});
          bool any = false;
          size_t widx = 0;
          int64_t bh = 0, bs = 0;
          for (size_t i = 0; i < moves.size(); ++i) {
            auto ev = orc.evaluate(moves[i]);
            if (ev.kind != sfo::EvalKind::Scored) continue;
            if (!any || ev.score.hard > bh || (ev.score.hard == bh && ev.score.soft > bs)) {
              bh = ev.score.hard;
              bs = ev.score.soft;
              widx = i;
              any = true;
            }
          }
          std::vector<uint32_t> s_idx, s_ev, s_win;
          std::vector<sf::HardSoftScore> s_best;
          sf::StepParams stp;
          stp.acceptor = 0;
          stp.random_ties = false;
          stp.accepted_limit = 0;
          cd.step_sublist_change(1, 3, stp, {1}, {}, s_idx, s_best, s_ev, s_win, false);
          CHECK(any && s_idx[0] == widx);
          CHECK(s_ev[0] == moves.size());
          CHECK(s_best[0].hard == bh && s_best[0].soft == bs);
          CHECK(s_win[0] == moves[widx].a && (s_win[1] & 0xFFFFFFu) == moves[widx].b && s_win[2] == moves[widx].d &&
                s_win[3] == moves[widx].e);
        }
        if (kind == 1 && !swap_moves) {
          // the union step through the mirror: three families in seeded Random order under StratifiedRandom, against
          // the oracle's cursors + UnionScheduler + candidate loop (LateAcceptance form, AcceptedCount(12))
          sfo::MoveStreamContext sctx;
          sctx.step_index = 3;
          sctx.step_seed = 4242;
          sctx.order = sfo::SelectionOrder::Random;
          std::vector<std::vector<sfo::Move>> kids = {orc.enumerate_list(8, sctx), orc.enumerate_list_reverse(sctx),
                                                      orc.enumerate_sublist_change(1, 2, sctx)};
          auto pulls = sfo::union_pull_order({kids[0].size(), kids[1].size(), kids[2].size()}, sfo::UnionOrder::StratifiedRandom,
                                             sctx, {1, 1, 1});
          const sfo::Sc olast = orc.calculate_score();
          sfo::Forager<sfo::Sc> of;
          of.kind = sfo::ForagerKind::AcceptedCount;
          of.accepted_count_limit = 12;
          sfo::Acceptor<sfo::Sc> la;
          la.kind = sfo::AcceptorKind::LateAcceptance;
          la.phase_started(olast, 1);
          auto want = sfo::replay_step<sfo::Sc>(
              pulls.size(), [&](size_t i) { return orc.evaluate(kids[pulls[i].first][pulls[i].second]); }, olast, olast, 4242, of, la);
          auto ud = sf::GpuScoreDirector::union_desc({{SFGPU_FAM_NEARBY_LIST_CHANGE, 8, 0, 0, 1}, {SFGPU_FAM_LIST_REVERSE, 0, 0, 0, 1},
                                                      {SFGPU_FAM_SUBLIST_CHANGE, 1, 2, 0, 1}});
          sf::StepParams up;
          up.acceptor = 2;
          up.random_ties = true;
          up.accepted_limit = 12;
          std::vector<uint32_t> uflags;
          sf::HardSoftScore last_c{olast.hard, olast.soft};
          auto ures = cd.step_union(ud, up, {4242}, {3}, {last_c, last_c}, false, &uflags);
          CHECK(uflags[0] == 0);
          CHECK(want.has_winner && ures.index[0] == want.winner);
          CHECK(ures.evaluated[0] == want.moves_evaluated);
          CHECK(ures.best[0].hard == want.winner_score.hard && ures.best[0].soft == want.winner_score.soft);
          CHECK(ures.winner_rows[1] == pulls[want.winner].first && ures.winner_rows[6] == pulls[want.winner].second);
        }
        sf::SolveParams sp;
        sp.n_steps = 10;
        sp.max_nearby = 8;
        sp.acceptor = kind + 1;
        sp.late_size = 4;
        sp.acceptor_real = kind == 2 ? 0.002 : 0.01;
        sp.step_count_limit = 2;
        sp.seed_base = 11;
        if (!swap_moves) {
          auto sr = cd.solve(sp, false);
          CHECK(sr.best[0] >= cm);
          CHECK(cd.fresh_score()[0] == cd.calculate_score()[0]);
          // and the default list search (union, seeded Random leaves, StratifiedRandom) as a resident loop
          sp.accepted_limit = 16;
          std::vector<uint64_t> ovf;
          auto su = cd.solve_union(sf::GpuScoreDirector::default_list_union(8), sp, &ovf);
          CHECK(ovf[0] == 0 && su.best[0] >= sr.best[0]);
          CHECK(cd.fresh_score()[0] == cd.calculate_score()[0]);
        }
      }
  }
  // error behaviour: unknown constraint kind is rejected, never emulated
  try {
    sf::GpuScoreDirector bad(1);
    sfgpu_constraint_desc cd{};
    cd.kind = 99;
    int rc = sfgpu_add_constraint(bad.raw(), &cd, nullptr);
    CHECK(rc == SFGPU_E_UNSUPPORTED);
  } catch (const sf::GpuError&) {
    CHECK(false);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  if (!std::strcmp(argv[1], "replay")) {
    test_replay();
    test_acceptor_trajectories();
  }
  else if (!std::strcmp(argv[1], "gpu")) test_gpu();
  else return 2;
  if (g_fail) {
    std::printf("HOST TEST FAILED %d of %d\n", g_fail, g_checks);
    return 1;
  }
  std::printf("HOST TEST OK %d\n", g_checks);
  return 0;
}
