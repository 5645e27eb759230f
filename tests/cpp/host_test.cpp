// Tests of the C++ host mirror (solverforge_b200/host/solverforge_gpu.hpp) against the oracle.
//   host_test replay   CPU only: acceptor/forager replay vs the oracle's restatement
//   host_test gpu      needs a B200: ConstraintFactory -> libsfgpu -> scores vs the oracle, bit-exact
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
          auto got = sf::replay_step(scores.data(), doable.data(), n, best_so_far, last, seed, fc, acc);
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
                return sfo::CandidateEvaluation<sfo::Sc>{doable[i] ? sfo::EvalKind::Scored : sfo::EvalKind::NotDoable,
                                                         sfo::Sc::of(scores[i].hard, scores[i].soft)};
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
  if (!std::strcmp(argv[1], "replay")) test_replay();
  else if (!std::strcmp(argv[1], "gpu")) test_gpu();
  else return 2;
  if (g_fail) {
    std::printf("HOST TEST FAILED %d of %d\n", g_fail, g_checks);
    return 1;
  }
  std::printf("HOST TEST OK %d\n", g_checks);
  return 0;
}
