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

// ---------------------------------------------------------------------------------------------------------
// O(1) / O(degree) CPU checkers for the scalar configs, so the full-size parity tests and the bench gate can
// compare EVERY candidate of every replica (the reference-faithful oracle walks whole collections per
// candidate: 4·10^4 predicate calls per graph-colouring candidate). Same role and status as FastCvrp above:
// test infrastructure, itself checked against the oracle on small instances (tests/test_oracle.py).
//
// C2  examples/scalar-graph-coloring/src/domain/graph_coloring.rs:21-44:
//       hard = -#unassigned - #{(a < b) adjacent, both assigned, same colour}
// C4  examples/mixed-job-shop/src/domain/job_shop_plan.rs:28-69 + the authored grouped complement:
//       hard = -#unassigned - #unscheduled;  soft = -#{(l < r) same job, same machine} - sum_m count(m)^2
namespace {
struct FastGc {
  uint32_t n, k;
  std::vector<uint32_t> rp, col;  // symmetric adjacency without duplicates
};
struct FastJs {
  uint32_t n, n_machines, n_jobs;
  std::vector<uint32_t> job;
  int64_t unscheduled;
  int with_complement;
};
}  // namespace

extern "C" {

// rebinds a FastCvrp to another route state (keeps the converted matrix)
void sfo_fast_cvrp_set_routes(void* h, const uint32_t* offsets, const uint32_t* elems, const int32_t* demands) {
  FastCvrp* f = static_cast<FastCvrp*>(h);
  const uint32_t dim = f->dim, n_routes = f->n_routes, depot = f->depot;
  const uint32_t total = offsets[n_routes];
  f->elems.assign(elems, elems + total);
  f->rem.assign(total, 0);
  f->slot_a.assign(total + n_routes, 0);
  f->slot_b.assign(total + n_routes, 0);
  f->slot_gap.assign(total + n_routes, 0);
  f->rsum.assign(n_routes, 0);
  f->hard = f->soft = 0;
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
    f->hard -= std::max<int64_t>(0, f->rsum[r] - f->capacity);
  }
  for (uint32_t c = 0; c < dim; ++c)
    if (c != depot && !seen[c]) f->hard -= 1;
}

void sfo_fast_cvrp_committed(const void* h, int64_t* out2) {
  const FastCvrp& f = *static_cast<const FastCvrp*>(h);
  out2[0] = f.hard;
  out2[1] = f.soft;
}

void* sfo_fast_gc_create(uint32_t n, uint32_t k, const uint32_t* row_ptr, const uint32_t* col) {
  auto* g = new FastGc();
  g->n = n;
  g->k = k;
  g->rp.assign(n + 1, 0);
  // `left.neighbors.contains(right.id)` is a set test: de-duplicate and symmetrise like the predicate does
  std::vector<std::vector<uint32_t>> adj(n);
  for (uint32_t a = 0; a < n; ++a)
    for (uint32_t j = row_ptr[a]; j < row_ptr[a + 1]; ++j) {
      const uint32_t b = col[j];
      if (a < b && b < n) {
        adj[a].push_back(b);
        adj[b].push_back(a);
      }
    }
  for (uint32_t a = 0; a < n; ++a) {
    std::sort(adj[a].begin(), adj[a].end());
    adj[a].erase(std::unique(adj[a].begin(), adj[a].end()), adj[a].end());
    g->rp[a + 1] = g->rp[a] + (uint32_t)adj[a].size();
    g->col.insert(g->col.end(), adj[a].begin(), adj[a].end());
  }
  return g;
}
void sfo_fast_gc_destroy(void* h) { delete static_cast<FastGc*>(h); }

// colors[n] (-1 = unassigned); rows[n_rows][2] = (entity, to_value (-1 = None)) as int32
void sfo_fast_gc_score(const void* h, const int32_t* colors, uint64_t n_rows, const int32_t* rows, int64_t* out_scores,
                       uint8_t* out_doable, int64_t* out_committed) {
  const FastGc& g = *static_cast<const FastGc*>(h);
  int64_t hard = 0;
  for (uint32_t a = 0; a < g.n; ++a) {
    if (colors[a] < 0) {
      hard -= 1;
      continue;
    }
    for (uint32_t j = g.rp[a]; j < g.rp[a + 1]; ++j)
      if (g.col[j] > a && colors[g.col[j]] == colors[a]) hard -= 1;
  }
  if (out_committed) {
    out_committed[0] = hard;
    out_committed[1] = 0;
  }
  for (uint64_t i = 0; i < n_rows; ++i) {
    const int64_t e = rows[2 * i];
    int32_t nv = rows[2 * i + 1];
    if (nv < 0) nv = -1;
    bool ok = e >= 0 && e < (int64_t)g.n && nv < (int32_t)g.k;
    int64_t h2 = hard;
    if (ok) {
      const int32_t ov = colors[e];
      ok = ov != nv;
      if (ok) {
        h2 += (ov < 0 ? 1 : 0) - (nv < 0 ? 1 : 0);
        for (uint32_t j = g.rp[e]; j < g.rp[e + 1]; ++j) {
          const int32_t c = colors[g.col[j]];
          if (c < 0) continue;
          if (c == ov) h2 += 1;
          if (c == nv) h2 -= 1;
        }
      }
    }
    out_scores[2 * i] = ok ? h2 : 0;
    out_scores[2 * i + 1] = 0;
    out_doable[i] = ok ? 1 : 0;
  }
}

void* sfo_fast_js_create(uint32_t n_ops, uint32_t n_machines, const uint32_t* job, int64_t unscheduled, int with_complement) {
  auto* f = new FastJs();
  f->n = n_ops;
  f->n_machines = n_machines;
  f->job.assign(job, job + n_ops);
  f->n_jobs = 0;
  for (uint32_t j : f->job) f->n_jobs = std::max(f->n_jobs, j + 1);
  f->unscheduled = unscheduled;
  f->with_complement = with_complement;
  return f;
}
void sfo_fast_js_destroy(void* h) { delete static_cast<FastJs*>(h); }

void sfo_fast_js_score(const void* h, const int32_t* machine, uint64_t n_rows, const int32_t* rows, int64_t* out_scores,
                       uint8_t* out_doable, int64_t* out_committed) {
  const FastJs& f = *static_cast<const FastJs*>(h);
  std::vector<int64_t> jm((size_t)f.n_jobs * f.n_machines, 0), load(f.n_machines, 0);
  int64_t hard = -f.unscheduled, soft = 0;
  for (uint32_t e = 0; e < f.n; ++e) {
    if (machine[e] < 0) {
      hard -= 1;
      continue;
    }
    jm[(size_t)f.job[e] * f.n_machines + machine[e]] += 1;
    load[machine[e]] += 1;
  }
  for (int64_t c : jm) soft -= c * (c - 1) / 2;
  if (f.with_complement)
    for (int64_t c : load) soft -= c * c;
  if (out_committed) {
    out_committed[0] = hard;
    out_committed[1] = soft;
  }
  for (uint64_t i = 0; i < n_rows; ++i) {
    const int64_t e = rows[2 * i];
    int32_t nv = rows[2 * i + 1];
    if (nv < 0) nv = -1;
    bool ok = e >= 0 && e < (int64_t)f.n && nv < (int32_t)f.n_machines;
    int64_t h2 = hard, s2 = soft;
    if (ok) {
      const int32_t ov = machine[e];
      ok = ov != nv;
      if (ok) {
        h2 += (ov < 0 ? 1 : 0) - (nv < 0 ? 1 : 0);
        const int64_t* row = jm.data() + (size_t)f.job[e] * f.n_machines;
        if (ov >= 0) {
          s2 += row[ov] - 1;
          if (f.with_complement) s2 -= (load[ov] - 1) * (load[ov] - 1) - load[ov] * load[ov];
        }
        if (nv >= 0) {
          s2 -= row[nv];
          if (f.with_complement) s2 -= (load[nv] + 1) * (load[nv] + 1) - load[nv] * load[nv];
        }
      }
    }
    out_scores[2 * i] = ok ? h2 : 0;
    out_scores[2 * i + 1] = ok ? s2 : 0;
    out_doable[i] = ok ? 1 : 0;
  }
}

}  // extern "C"
