// ORACLE DIRECTORY — TEST / BENCH INFRASTRUCTURE ONLY (never linked by the product).
//
// A STRONGER CPU baseline than the reference-faithful engine: the same
// read-only O(1) delta the GPU fast path computes for CVRP list-change candidates (per-route sums,
// per-position removal gains, per-slot gap costs, two matrix lookups per candidate), in plain C++.
// It exists so the GPU/CPU ratio reported by bench.py is not inflated by the reference's O(route)
// closures and do/undo protocol. Results are checked against the oracle in tests/test_oracle.py.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {
struct FastCvrp {
  uint32_t dim, n_routes, depot;
  int64_t capacity;
  std::vector<int32_t> mat;       // distance_cost applied, int32
  std::vector<int32_t> demand;
  std::vector<uint32_t> base, len, elems;
  std::vector<int64_t> rsum;
  std::vector<int32_t> rem;       // per flat position
  std::vector<uint32_t> slot_a, slot_b;
  std::vector<int32_t> slot_gap;  // per slot base[e] + e + p
  int64_t hard = 0, soft = 0;
};
}  // namespace

extern "C" {

void* sfo_fast_cvrp_create(uint32_t dim, uint32_t n_routes, int64_t capacity, uint32_t depot, const int32_t* demands,
                           const int64_t* matrix, const uint32_t* offsets, const uint32_t* elems) {
  auto* f = new FastCvrp();
  f->dim = dim;
  f->n_routes = n_routes;
  f->depot = depot;
  f->capacity = capacity;
  f->mat.resize((size_t)dim * dim);
  for (size_t i = 0; i < f->mat.size(); ++i) {
    int64_t v = matrix[i];
    if (!(v >= 0 && v != INT64_MAX)) v = INT64_MAX / 4;
    if (v > INT32_MAX) {  // the fast path is int32-only, like the GPU one
      delete f;
      return nullptr;
    }
    f->mat[i] = (int32_t)v;
  }
  f->demand.assign(demands, demands + dim);
  f->base.resize(n_routes);
  f->len.resize(n_routes);
  f->rsum.assign(n_routes, 0);
  uint32_t total = offsets[n_routes];
  f->elems.assign(elems, elems + total);
  f->rem.resize(total);
  f->slot_a.resize(total + n_routes);
  f->slot_b.resize(total + n_routes);
  f->slot_gap.resize(total + n_routes);
  std::vector<bool> seen(dim, false);
  for (uint32_t r = 0; r < n_routes; ++r) {
    uint32_t b = offsets[r], l = offsets[r + 1] - b;
    f->base[r] = b;
    f->len[r] = l;
    int64_t cost = 0;
    for (uint32_t p = 0; p <= l; ++p) {
      uint32_t a_el = p > 0 ? elems[b + p - 1] : depot, b_el = p < l ? elems[b + p] : depot;
      f->slot_a[b + r + p] = a_el;
      f->slot_b[b + r + p] = b_el;
      f->slot_gap[b + r + p] = l > 0 ? f->mat[(size_t)a_el * dim + b_el] : 0;
      if (l > 0) cost += f->mat[(size_t)a_el * dim + b_el];
      if (p < l) {
        uint32_t x = b_el, nx = p + 1 < l ? elems[b + p + 1] : depot;
        f->rem[b + p] = -f->mat[(size_t)a_el * dim + x] - f->mat[(size_t)x * dim + nx] +
                        (l > 1 ? f->mat[(size_t)a_el * dim + nx] : 0);
        f->rsum[r] += demands[x];
        seen[x] = true;
      }
    }
    f->soft -= cost;
    f->hard -= std::max<int64_t>(0, f->rsum[r] - capacity);
  }
  for (uint32_t c = 0; c < dim; ++c)
    if (c != depot && !seen[c]) f->hard -= 1;
  return f;
}

void sfo_fast_cvrp_destroy(void* h) { delete static_cast<FastCvrp*>(h); }

// scores n list-change rows {se, sp, de, dp}; out_scores[n][2], out_doable[n]
void sfo_fast_cvrp_score(const void* h, uint64_t n, const uint32_t* rows, int64_t* out_scores, uint8_t* out_doable) {
  const FastCvrp& f = *static_cast<const FastCvrp*>(h);
  const uint32_t dim = f.dim;
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t se = rows[4 * i], sp = rows[4 * i + 1], de = rows[4 * i + 2], dp = rows[4 * i + 3];
    bool ok = se < f.n_routes && de < f.n_routes;
    int64_t h2 = 0, s2 = 0;
    if (ok) {
      const uint32_t slen = f.len[se], dlen = f.len[de];
      ok = sp < slen && dp <= dlen && !(se == de && (dp == sp || dp == sp + 1));
      if (ok) {
        const uint32_t fp = f.base[se] + sp, sl = f.base[de] + de + dp;
        const uint32_t x = f.elems[fp];
        const int32_t d = f.rem[fp] + f.mat[(size_t)f.slot_a[sl] * dim + x] + f.mat[(size_t)x * dim + f.slot_b[sl]] -
                          f.slot_gap[sl];
        s2 = f.soft - d;
        h2 = f.hard;
        if (se != de) {
          const int64_t v = f.demand[x], e0 = f.rsum[se] - f.capacity, e1 = f.rsum[de] - f.capacity;
          h2 -= (std::max<int64_t>(e0 - v, 0) - std::max<int64_t>(e0, 0)) +
                (std::max<int64_t>(e1 + v, 0) - std::max<int64_t>(e1, 0));
        }
      }
    }
    out_scores[2 * i] = ok ? h2 : 0;
    out_scores[2 * i + 1] = ok ? s2 : 0;
    out_doable[i] = ok ? 1 : 0;
  }
}

// candidates/s of n_threads threads each re-scoring the same batch against its own copy for `seconds`
double sfo_fast_cvrp_bench(const void* h, uint64_t n, const uint32_t* rows, uint32_t n_threads, double seconds) {
  std::atomic<uint64_t> total{0};
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> ths;
  for (uint32_t t = 0; t < n_threads; ++t)
    ths.emplace_back([&, t] {
      FastCvrp copy = *static_cast<const FastCvrp*>(h);  // one solver per thread, like the reference
      std::vector<int64_t> sc(n * 2);
      std::vector<uint8_t> ok(n);
      uint64_t done = 0;
      for (;;) {
        sfo_fast_cvrp_score(&copy, n, rows, sc.data(), ok.data());
        done += n;
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() >= seconds) break;
      }
      total += done;
    });
  for (auto& th : ths) th.join();
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return (double)total.load() / dt;
}

}  // extern "C"
