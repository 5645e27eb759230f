// CUDA kernels of libsfgpu (sm_100a). All integer/index work, HBM/L2-bound; no tensor cores.
//
// Key idea: the GPU never replays do/undo. For a candidate m against the
// immutable committed state S of its replica it computes
//     score(m) = committed + sum_k [ contrib_k(S (+) m, touched) - contrib_k(S, touched) ]
// read-only, so every candidate is independent. The retained aggregates of the reference
// (per-key counts, per-group accumulators, per-route sums) live in the replica block and are
// only updated when the winning move is applied.
#pragma once
#include "sfgpu_dev.cuh"

// =============================================================================================
// scalar edits (ChangeMove / SwapMove / CompoundScalarMove)
// =============================================================================================
struct EditDev {
  uint32_t e;
  int32_t old_v, new_v;
};

__device__ __forceinline__ int32_t var_at(const int32_t* var, const EditDev* prev, int n_prev, uint32_t x) {
  int32_t v = var[x];
  for (int i = 0; i < n_prev; ++i)
    if (prev[i].e == x) v = prev[i].new_v;
  return v;
}

__device__ __forceinline__ int64_t col_at(const void* col, uint32_t i) {
  return col ? ((const int64_t*)col)[i] : 0;
}

// key(e, v) of a PAIR_KEY_EQUAL constraint, relative to the table origin p2
__device__ __forceinline__ int64_t pair_key(const ConsDev& c, uint32_t e, int32_t v) {
  return col_at(c.g0, e) * c.p0 + (int64_t)v * c.p1 - c.p2;
}

__device__ __forceinline__ int64_t uni_x(const ConsDev& c, uint32_t e, int32_t v) {
  if (!c.g0) return 0;
  if (c.flags & SFGPU_CF_COL_BY_VALUE) return v >= 0 ? ((const int64_t*)c.g0)[v] : 0;
  return ((const int64_t*)c.g0)[e];
}
__device__ __forceinline__ int64_t uni_contrib(const ConsDev& c, uint32_t e, int32_t v) {
  bool pass = c.p0 == 0 ? v < 0 : (c.p0 == 1 ? v >= 0 : true);
  if (c.g1 && ((const int64_t*)c.g1)[e] == 0) pass = false;  // entity mask column
  return pass ? weight_eval(c.w, uni_x(c, e, v)) : 0;
}

// score of a group with `count` rows and accumulated `sum` (count() result == count)
// `key` = the group's value row; a per-value column (g1) may replace the weight offset b, which is
// how key-dependent weights such as max(0, demand - capacity[key]) or key*10 + demand are expressed
__device__ __forceinline__ int64_t group_score(const ConsDev& c, uint32_t key, int64_t count, int64_t sum) {
  bool counting = c.g0 == nullptr;
  WeightDev w = c.w;
  if (c.g1) w.b = ((const int64_t*)c.g1)[key];
  if (count > 0) return weight_eval(w, counting ? count : sum);
  if (c.flags & SFGPU_CF_COMPLEMENT) return weight_eval(w, c.p1);
  return 0;
}

// load_balance.rs:165-183
__device__ __forceinline__ int64_t lb_unfairness(int64_t n, int64_t sum, int64_t sumsq) {
  if (n == 0) return 0;
  double frac = (double)(-(sum * sum));
  double tmp = n == 1 ? frac + (double)sumsq : frac / (double)n + (double)sumsq;
  return (int64_t)round(sqrt(tmp));
}

// indexed_presence views of one group (stream/collector/indexed_presence.rs): cntf(q) = items at point q.
// c.p0: 1 = complement_runs(lo..hi) scored run by run, 2 = any_in(lo..hi), 3 = count() of distinct points;
// c.p1 / c.p2 = lo / hi. A group without items does not exist and scores nothing (grouped/state.rs:349-365).
template <class F>
__device__ __forceinline__ int64_t presence_group_score(const ConsDev& c, F cntf) {
  const int32_t np = (int32_t)c.n0, lo = (int32_t)c.p1, hi = (int32_t)c.p2;
  int64_t items = 0, distinct = 0, gaps = 0, gap = 0;
  bool any = false;
  for (int32_t q = 0; q < np; ++q) {
    const int32_t n = cntf(q);
    items += n;
    distinct += n > 0 ? 1 : 0;
    if (q >= lo && q < hi) {
      any = any || n > 0;
      if (n == 0) {
        ++gap;
      } else {
        if (gap > 0) gaps += weight_eval(c.w, gap);
        gap = 0;
      }
    }
  }
  if (gap > 0) gaps += weight_eval(c.w, gap);
  if (items == 0) return 0;
  if (c.p0 == 1) return gaps;
  if (c.p0 == 2) return any ? weight_eval(c.w, 1) : 0;
  return weight_eval(c.w, distinct);
}

// SFGPU_K_PAIR_KEY_EXPR: every joined pair that involves row (e, v), against the committed lists overlaid with the
// edits prev[0..n_prev) of the same candidate (rows of edited entities are taken from the overlay, not from the lists)
static __device__ int64_t pke_contrib(const DevModel& m, const ConsDev& c, const char* gst, const int32_t* var, const EditDev* prev,
                                      int n_prev, uint32_t e, int32_t v) {
  if (v < 0) return 0;
  const ExprTables t{(const int64_t* const*)c.g1, (const uint32_t* const*)c.g2};
  const PairKeyLists L = pke_lists(c, const_cast<char*>(gst), m.n_entities);
  const bool directed = pke_directed(c);
  auto edited = [&](uint32_t x) {
    for (int i = 0; i < n_prev; ++i)
      if (prev[i].e == x) return true;
    return false;
  };
  auto last_edit = [&](int i) {  // the overlay row of an entity is its LAST edit
    for (int j = i + 1; j < n_prev; ++j)
      if (prev[j].e == prev[i].e) return false;
    return true;
  };
  int64_t total = 0;
  if (!directed) {
    const int64_t k = pke_key(c, t, 0, e, v);
    if (k < 0) return 0;
    for (int32_t x = L.headL[k]; x >= 0; x = L.nxtL[x]) {
      const uint32_t mth = (uint32_t)x;
      if (mth == e || edited(mth)) continue;
      total += mth < e ? pke_pair(c, t, mth, var[mth], e, v) : pke_pair(c, t, e, v, mth, var[mth]);
    }
    for (int i = 0; i < n_prev; ++i) {
      const uint32_t pe = prev[i].e;
      if (pe == e || !last_edit(i) || pke_key(c, t, 0, pe, prev[i].new_v) != k) continue;
      total += pe < e ? pke_pair(c, t, pe, prev[i].new_v, e, v) : pke_pair(c, t, e, v, pe, prev[i].new_v);
    }
    return total;
  }
  const int64_t kl = pke_key(c, t, 0, e, v), kr = pke_key(c, t, 1, e, v);
  if (kl >= 0) {  // e as the left row: right rows whose right key equals e's left key
    for (int32_t x = L.headR[kl]; x >= 0; x = L.nxtR[x])
      if ((uint32_t)x != e && !edited((uint32_t)x)) total += pke_pair(c, t, e, v, (uint32_t)x, var[x]);
    for (int i = 0; i < n_prev; ++i)
      if (prev[i].e != e && last_edit(i) && pke_key(c, t, 1, prev[i].e, prev[i].new_v) == kl)
        total += pke_pair(c, t, e, v, prev[i].e, prev[i].new_v);
  }
  if (kr >= 0) {  // e as the right row
    for (int32_t x = L.headL[kr]; x >= 0; x = L.nxtL[x])
      if ((uint32_t)x != e && !edited((uint32_t)x)) total += pke_pair(c, t, (uint32_t)x, var[x], e, v);
    for (int i = 0; i < n_prev; ++i)
      if (prev[i].e != e && last_edit(i) && pke_key(c, t, 0, prev[i].e, prev[i].new_v) == kr)
        total += pke_pair(c, t, prev[i].e, prev[i].new_v, e, v);
  }
  return total;
}

// Delta of ONE edit (e: old -> new) against the replica state `st` overlaid with the edits
// prev[0..n_prev) that were already applied by the same candidate.
// `st` = the replica block as the kernel sees it (possibly the staged shared-memory prefix); `gst` = the
// same block in global memory, for retained sections too large to stage (CSR partner-value counts).
static __device__ void scalar_edit_delta(const DevModel& m, const char* st, const char* gst, const EditDev* prev, int n_prev, EditDev cur,
                                  Score2& d) {
  if (cur.old_v == cur.new_v) return;
  const int32_t* var = (const int32_t*)(st + m.off_var);
  for (uint32_t k = 0; k < m.n_cons; ++k) {
    const ConsDev& c = m.cons[k];
    switch (c.kind) {
      case SFGPU_K_UNI: {
        add_level(d, c, uni_contrib(c, cur.e, cur.new_v) - uni_contrib(c, cur.e, cur.old_v));
        break;
      }
      case SFGPU_K_PAIR_KEY_EXPR: {
        add_level(d, c, pke_contrib(m, c, gst, var, prev, n_prev, cur.e, cur.new_v) -
                            pke_contrib(m, c, gst, var, prev, n_prev, cur.e, cur.old_v));
        break;
      }
      case SFGPU_K_JOIN_EXPR: {  // pairs of one A row depend on that row only: no overlay needed
        add_level(d, c, join_expr_contrib(c, cur.e, cur.new_v) - join_expr_contrib(c, cur.e, cur.old_v));
        break;
      }
      case SFGPU_K_PAIR_CSR_EQUAL: {
        // g0 = partner row_ptr, g1 = partner ids (symmetrised, deduplicated at commit)
        const uint32_t* rp = (const uint32_t*)c.g0;
        const uint32_t* ci = (const uint32_t*)c.g1;
        int64_t cnt = 0;
        if (c.off0 != 0xFFFFFFFFu && n_prev == 0) {
          // retained partner-value counts cc[e][v] (the analogue of the reference's retained match rows):
          // O(1) instead of a walk over the partners
          const uint16_t* cc = (const uint16_t*)(gst + c.off0) + (size_t)cur.e * m.n_values;
          cnt = (cur.new_v >= 0 ? (int64_t)cc[cur.new_v] : 0) - (cur.old_v >= 0 ? (int64_t)cc[cur.old_v] : 0);
        } else {
          uint32_t lo = rp[cur.e], hi = rp[cur.e + 1];
          for (uint32_t j = lo; j < hi; ++j) {
            int32_t vp = var_at(var, prev, n_prev, ci[j]);
            cnt += (cur.new_v >= 0 && vp == cur.new_v) ? 1 : 0;
            cnt -= (cur.old_v >= 0 && vp == cur.old_v) ? 1 : 0;
          }
        }
        add_level(d, c, cnt * c.w.a);
        break;
      }
      case SFGPU_K_PAIR_KEY_EQUAL: {
        const int32_t* tab = (const int32_t*)(st + c.off0);
        const int ar = (int)c.pad;  // join arity 2..5: a bucket of n rows holds C(n, arity) tuples
        int64_t cnt = 0;
        if (cur.new_v >= 0) {
          int64_t key = pair_key(c, cur.e, cur.new_v);
          int64_t n = tab[key];
          for (int i = 0; i < n_prev; ++i) {
            n += (prev[i].new_v >= 0 && pair_key(c, prev[i].e, prev[i].new_v) == key) ? 1 : 0;
            n -= (prev[i].old_v >= 0 && pair_key(c, prev[i].e, prev[i].old_v) == key) ? 1 : 0;
          }
          cnt += choose_small(n, ar - 1);
        }
        if (cur.old_v >= 0) {
          int64_t key = pair_key(c, cur.e, cur.old_v);
          int64_t n = tab[key];
          for (int i = 0; i < n_prev; ++i) {
            n += (prev[i].new_v >= 0 && pair_key(c, prev[i].e, prev[i].new_v) == key) ? 1 : 0;
            n -= (prev[i].old_v >= 0 && pair_key(c, prev[i].e, prev[i].old_v) == key) ? 1 : 0;
          }
          cnt -= choose_small(n - 1, ar - 1);
        }
        add_level(d, c, cnt * c.w.a);
        break;
      }
      case SFGPU_K_GROUP: {
        const int32_t* gc = (const int32_t*)(st + c.off0);
        const int64_t* gs = (const int64_t*)(st + c.off1);
        int64_t x = c.g0 ? ((const int64_t*)c.g0)[cur.e] : 1;
        int64_t delta = 0;
        for (int side = 0; side < 2; ++side) {
          int32_t v = side == 0 ? cur.old_v : cur.new_v;
          if (v < 0) continue;
          int64_t cn = gc[v], sm = gs[v];
          for (int i = 0; i < n_prev; ++i) {
            int64_t xi = c.g0 ? ((const int64_t*)c.g0)[prev[i].e] : 1;
            if (prev[i].new_v == v) { cn += 1; sm += xi; }
            if (prev[i].old_v == v) { cn -= 1; sm -= xi; }
          }
          int64_t before = group_score(c, (uint32_t)v, cn, sm);
          int64_t after = side == 0 ? group_score(c, (uint32_t)v, cn - 1, sm - x) : group_score(c, (uint32_t)v, cn + 1, sm + x);
          delta += after - before;
        }
        add_level(d, c, delta);
        break;
      }
      case SFGPU_K_LOAD_BALANCE: {
        // single group; balanced key = assigned value; metric x (zero metrics are skipped,
        // load_balance.rs:196-200). Overlay-free: compound candidates touching a load-balance
        // constraint are rejected at the API (SFGPU_E_UNSUPPORTED) unless n_prev == 0.
        const int64_t* loads = (const int64_t*)(st + c.off0);
        const int32_t* icnt = (const int32_t*)(st + c.off1);
        const int64_t* agg = (const int64_t*)(st + c.off2);  // {sum, sumsq, nkeys}
        int64_t x = c.g0 ? ((const int64_t*)c.g0)[cur.e] : 1;
        if (x == 0) break;
        int64_t sum = agg[0], sumsq = agg[1], nk = agg[2];
        int64_t before = weight_eval(c.w, lb_unfairness(nk, sum, sumsq));
        if (cur.old_v >= 0) {
          int64_t l = loads[cur.old_v];
          int64_t nl = icnt[cur.old_v] == 1 ? 0 : l - x;
          if (icnt[cur.old_v] == 1) nk -= 1;
          sumsq += nl * nl - l * l;
          sum += nl - l;
        }
        if (cur.new_v >= 0) {
          int64_t l = loads[cur.new_v];  // 0 when the key has no items
          if (icnt[cur.new_v] == 0) nk += 1;
          int64_t nl = l + x;
          sumsq += nl * nl - l * l;
          sum += x;
        }
        add_level(d, c, weight_eval(c.w, lb_unfairness(nk, sum, sumsq)) - before);
        break;
      }
      case SFGPU_K_RUNS: {
        // cnt[v][point] items of group v at each point. Leaving a point whose count drops to zero splits
        // the run around it (L + 1 + R -> L, R); entering an empty point merges its neighbours.
        const int32_t* cnt = (const int32_t*)(st + c.off0);
        const int32_t np = (int32_t)c.n0;
        const int32_t pt = (int32_t)((const int64_t*)c.g0)[cur.e];
        int64_t delta = 0;
        if (c.p0 != 0) {  // indexed_presence views: re-score the (at most two) touched groups from their rows
          for (int side = 0; side < 2; ++side) {
            const int32_t v = side == 0 ? cur.old_v : cur.new_v;
            if (v < 0) continue;
            const int32_t* row = cnt + (size_t)v * np;
            auto at = [&](int32_t q) {  // earlier edits of this candidate included
              int32_t n = row[q];
              for (int i = 0; i < n_prev; ++i) {
                if ((int32_t)((const int64_t*)c.g0)[prev[i].e] != q) continue;
                if (prev[i].new_v == v) n += 1;
                if (prev[i].old_v == v) n -= 1;
              }
              return n;
            };
            const int32_t dn = side == 0 ? -1 : 1;
            delta += presence_group_score(c, [&](int32_t q) { return at(q) + (q == pt ? dn : 0); }) -
                     presence_group_score(c, at);
          }
          add_level(d, c, delta);
          break;
        }
        for (int side = 0; side < 2; ++side) {
          const int32_t v = side == 0 ? cur.old_v : cur.new_v;
          if (v < 0) continue;
          const int32_t* row = cnt + (size_t)v * np;
          auto at = [&](int32_t q) {  // items at point q of group v, earlier edits of this candidate included
            int32_t n = row[q];
            for (int i = 0; i < n_prev; ++i) {
              const int32_t pq = (int32_t)((const int64_t*)c.g0)[prev[i].e];
              if (pq != q) continue;
              if (prev[i].new_v == v) n += 1;
              if (prev[i].old_v == v) n -= 1;
            }
            return n;
          };
          const int32_t here = at(pt);
          if (side == 0 ? here != 1 : here != 0) continue;  // the set of points does not change
          int64_t L = 0, R = 0;
          for (int32_t q = pt - 1; q >= 0 && at(q) > 0; --q) ++L;
          for (int32_t q = pt + 1; q < np && at(q) > 0; ++q) ++R;
          const int64_t merged = weight_eval(c.w, L + 1 + R);
          const int64_t split = (L > 0 ? weight_eval(c.w, L) : 0) + (R > 0 ? weight_eval(c.w, R) : 0);
          delta += side == 0 ? split - merged : merged - split;
        }
        add_level(d, c, delta);
        break;
      }
      case SFGPU_K_PROJECT_GROUP: {
        // projected rows of ONE entity change groups together: collect the affected groups first, then
        // re-score each once (two rows of an entity may share a group; weights need not be linear)
        const uint32_t* rp = (const uint32_t*)c.g0;
        const longlong2* em = (const longlong2*)c.g1;  // {key offset, amount}
        const int32_t* gc = (const int32_t*)(st + c.off0);
        const int64_t* gs = (const int64_t*)(st + c.off1);
        const uint32_t stride = c.n0;
        uint32_t keys[16];
        int32_t dcn[16];
        int64_t dsm[16];
        int ng = 0;
        for (int side = 0; side < 2; ++side) {
          const int32_t v = side == 0 ? cur.old_v : cur.new_v;
          if (v < 0) continue;
          for (uint32_t j = rp[cur.e]; j < rp[cur.e + 1]; ++j) {
            const longlong2 row = em[j];
            const uint32_t key = (uint32_t)v * stride + (uint32_t)row.x;
            int g = 0;
            while (g < ng && keys[g] != key) ++g;
            if (g == ng) {
              keys[ng] = key;
              dcn[ng] = 0;
              dsm[ng] = 0;
              ++ng;
            }
            dcn[g] += side == 0 ? -1 : 1;
            dsm[g] += side == 0 ? -row.y : row.y;
          }
        }
        int64_t delta = 0;
        for (int g = 0; g < ng; ++g) {
          int64_t cn = gc[keys[g]], sm = gs[keys[g]];
          for (int i = 0; i < n_prev; ++i)  // edits already applied by the same compound candidate
            for (uint32_t j = rp[prev[i].e]; j < rp[prev[i].e + 1]; ++j) {
              const longlong2 row = em[j];
              if (prev[i].new_v >= 0 && (uint32_t)prev[i].new_v * stride + (uint32_t)row.x == keys[g]) { cn += 1; sm += row.y; }
              if (prev[i].old_v >= 0 && (uint32_t)prev[i].old_v * stride + (uint32_t)row.x == keys[g]) { cn -= 1; sm -= row.y; }
            }
          const int64_t before = cn > 0 ? weight_eval(c.w, sm) : 0;
          const int64_t after = cn + dcn[g] > 0 ? weight_eval(c.w, sm + dsm[g]) : 0;
          delta += after - before;
        }
        add_level(d, c, delta);
        break;
      }
      default: break;  // list-only kinds do not react to scalar edits (ChangeSource routing)
    }
  }
}

enum { MODE_CHANGE = 0, MODE_SWAP = 1, MODE_COMPOUND = 2 };

// ChangeMove {entity, to_value}: doable iff the target differs from the current value (change.rs:125-139)
__device__ __forceinline__ bool score_change_row(const DevModel& m, const char* st, const char* gst, uint2 row,
                                                 Score2& d) {
  const int32_t* var = (const int32_t*)(st + m.off_var);
  d.hard = 0;
  d.soft = 0;
  uint32_t e = row.x;
  int32_t nv = (int32_t)row.y;
  if (e >= m.n_entities || nv >= (int32_t)m.n_values) return false;
  if (nv < 0) nv = SFGPU_NONE;
  int32_t ov = var[e];
  if (ov == nv) return false;
  scalar_edit_delta(m, st, gst, nullptr, 0, EditDev{e, ov, nv}, d);
  return true;
}

// one thread = one candidate. Returns doable; d = score delta.
template <int MODE>
__device__ __forceinline__ bool score_scalar_candidate(const DevModel& m, const char* st, const char* gst, const uint32_t* rows,
                                                       const uint64_t* edit_offsets, uint64_t i, Score2& d) {
  const int32_t* var = (const int32_t*)(st + m.off_var);
  d.hard = 0;
  d.soft = 0;
  if (MODE == MODE_CHANGE) {
    return score_change_row(m, st, gst, ((const uint2*)rows)[i], d);
  } else if (MODE == MODE_SWAP) {
    uint2 row = ((const uint2*)rows)[i];
    if (row.x >= m.n_entities || row.y >= m.n_entities) return false;
    int32_t lv = var[row.x], rv = var[row.y];
    if (lv == rv) return false;  // swap.rs:140-157
    EditDev e0{row.x, lv, rv};
    scalar_edit_delta(m, st, gst, nullptr, 0, e0, d);
    scalar_edit_delta(m, st, gst, &e0, 1, EditDev{row.y, rv, lv}, d);
    return true;
  } else {
    uint64_t lo = edit_offsets[i], hi = edit_offsets[i + 1];
    int n = (int)(hi - lo);
    if (n <= 0 || n > SFGPU_MAX_EDITS) return false;
    EditDev ed[SFGPU_MAX_EDITS];
    bool changes = false;
    for (int j = 0; j < n; ++j) {
      uint2 row = ((const uint2*)rows)[lo + j];
      if (row.x >= m.n_entities || (int32_t)row.y >= (int32_t)m.n_values) return false;
      int32_t nv = (int32_t)row.y < 0 ? SFGPU_NONE : (int32_t)row.y;
      changes |= var[row.x] != nv;  // compound_scalar.rs:245-261 (against the unmodified solution)
      ed[j].e = row.x;
      ed[j].new_v = nv;
      ed[j].old_v = var_at(var, ed, j, row.x);
    }
    if (!changes) return false;
    for (int j = 0; j < n; ++j) scalar_edit_delta(m, st, gst, ed, j, ed[j], d);
    return true;
  }
}

// grid = (chunks, R). Each CTA stages its replica block (TMA bulk copy) and scores a strided
// share of the replica's candidate range.
template <int MODE, bool STAGED>
__global__ void __launch_bounds__(256) score_scalar_kernel(const __grid_constant__ DevModel m,
                                                           const uint64_t* __restrict__ cand_offsets,
                                                           const uint32_t* __restrict__ rows,
                                                           const uint64_t* __restrict__ edit_offsets,
                                                           int64_t* __restrict__ out_scores,
                                                           uint8_t* __restrict__ out_doable) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  const uint32_t r = blockIdx.y;
  const char* gblock = m.state + (size_t)r * m.block_bytes;
  const char* st = gblock;
  if (STAGED) {
    stage_block(smem, gblock, m.stage_bytes, &bar);
    st = smem;
  }
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  const uint64_t lo = cand_offsets[r], hi = cand_offsets[r + 1];
  if (MODE == MODE_CHANGE) {
    // four rows per thread per trip, loaded before any is scored: four independent 64-bit loads in
    // flight hide the DRAM latency of the streamed batch
    constexpr int U = 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < hi; base += stride * U) {
      uint2 row[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint64_t i = base + u * stride;
        row[u] = i < hi ? __ldcs((const uint2*)rows + i) : make_uint2(0xFFFFFFFFu, 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint64_t i = base + u * stride;
        if (i >= hi) break;
        Score2 d;
        const bool ok = score_change_row(m, st, gblock, row[u], d);
        longlong2 o;
        o.x = ok ? ch + d.hard : 0;
        o.y = ok ? csf + d.soft : 0;
        __stcs((longlong2*)out_scores + i, o);
        out_doable[i] = ok ? 1 : 0;
      }
    }
  } else {
    for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi;
         i += (uint64_t)gridDim.x * blockDim.x) {
      Score2 d;
      bool ok = score_scalar_candidate<MODE>(m, st, gblock, rows, edit_offsets, i, d);
      longlong2 o;
      o.x = ok ? ch + d.hard : 0;
      o.y = ok ? csf + d.soft : 0;
      ((longlong2*)out_scores)[i] = o;
      out_doable[i] = ok ? 1 : 0;
    }
  }
}

// =============================================================================================
// list moves (ListChangeMove / ListSwapMove)
// =============================================================================================
__device__ __forceinline__ int64_t mat_at(const ConsDev& c, uint32_t from, uint32_t to) {
  size_t idx = (size_t)from * c.n0 + to;
  if (c.flags & SFGPU_CF_MATRIX_I32) return ((const int32_t*)c.g0)[idx];
  return ((const int64_t*)c.g0)[idx];
}

// ListChange delta. Returns doable (list_kernel/change.rs:46-75).
__device__ __forceinline__ bool list_change_delta(const DevModel& m, const char* st, uint4 row, Score2& d) {
  const uint32_t* off = (const uint32_t*)(st + m.off_offsets);
  const uint32_t* el = (const uint32_t*)(st + m.off_elems);
  d.hard = 0;
  d.soft = 0;
  const uint32_t se = row.x, sp = row.y, de = row.z, dp = row.w;
  if (se >= m.n_owners || de >= m.n_owners) return false;
  const uint32_t sb = off[se], slen = off[se + 1] - sb;
  if (sp >= slen) return false;
  const bool intra = se == de;
  const uint32_t db = off[de], dlen = off[de + 1] - db;
  if (dp > (intra ? slen : dlen)) return false;
  if (intra && (dp == sp || dp == sp + 1)) return false;
  const uint32_t x = el[sb + sp];
  for (uint32_t k = 0; k < m.n_cons; ++k) {
    const ConsDev& c = m.cons[k];
    if (c.kind == SFGPU_K_LIST_PATH_COST) {
      const uint32_t depot = (uint32_t)c.p0;
      const uint32_t prev = sp > 0 ? el[sb + sp - 1] : depot;
      const uint32_t next = sp + 1 < slen ? el[sb + sp + 1] : depot;
      const uint32_t a = dp > 0 ? el[db + dp - 1] : depot;
      const uint32_t b = dp < dlen ? el[db + dp] : depot;
      // removal; an emptied route costs 0 (no depot->depot leg)
      int64_t rem = -mat_at(c, prev, x) - mat_at(c, x, next) + (slen > 1 ? mat_at(c, prev, next) : 0);
      // insertion; inserting into an empty route has no old leg to remove
      int64_t ins = mat_at(c, a, x) + mat_at(c, x, b) - ((intra || dlen > 0) ? mat_at(c, a, b) : 0);
      const int64_t* rcost = (const int64_t*)(st + c.off0);
      if (intra) {
        int64_t old_c = rcost[se];
        add_level(d, c, weight_eval(c.w, old_c + rem + ins) - weight_eval(c.w, old_c));
      } else {
        int64_t os = rcost[se], od = rcost[de];
        add_level(d, c, weight_eval(c.w, os + rem) - weight_eval(c.w, os) + weight_eval(c.w, od + ins) -
                            weight_eval(c.w, od));
      }
    } else if (c.kind == SFGPU_K_LIST_SUM) {
      if (!intra) {
        const int64_t* rsum = (const int64_t*)(st + c.off0);
        int64_t v = ((const int64_t*)c.g0)[x];
        int64_t ss = rsum[se], ds = rsum[de];
        add_level(d, c, weight_eval(c.w, ss - v) - weight_eval(c.w, ss) + weight_eval(c.w, ds + v) -
                            weight_eval(c.w, ds));
      }
    }
    // EXISTS_FLAT: a relocation keeps the multiset of flattened keys => per-key counts unchanged => 0.
  }
  return true;
}

// ListSwap delta (list_kernel/swap.rs:31-57 doability).
__device__ __forceinline__ bool list_swap_delta(const DevModel& m, const char* st, uint4 row, Score2& d) {
  const uint32_t* off = (const uint32_t*)(st + m.off_offsets);
  const uint32_t* el = (const uint32_t*)(st + m.off_elems);
  d.hard = 0;
  d.soft = 0;
  const uint32_t e1 = row.x, p1 = row.y, e2 = row.z, p2 = row.w;
  if (e1 >= m.n_owners || e2 >= m.n_owners) return false;
  const uint32_t b1 = off[e1], len1 = off[e1 + 1] - b1;
  const uint32_t b2 = off[e2], len2 = off[e2 + 1] - b2;
  if (p1 >= len1 || p2 >= len2) return false;
  const bool intra = e1 == e2;
  if (intra && p1 == p2) return false;
  const uint32_t x1 = el[b1 + p1], x2 = el[b2 + p2];
  if (x1 == x2) return false;
  const uint32_t f1 = b1 + p1, f2 = b2 + p2;  // flat positions
  for (uint32_t k = 0; k < m.n_cons; ++k) {
    const ConsDev& c = m.cons[k];
    if (c.kind == SFGPU_K_LIST_PATH_COST) {
      const uint32_t depot = (uint32_t)c.p0;
      // legs touched: (e1,p1) (e1,p1+1) (e2,p2) (e2,p2+1); leg (e,i) joins position i-1 and i
      // (position -1 and len are the depot). Patched accessor sees the swapped values.
      int64_t delta1 = 0, delta2 = 0;
      for (int t = 0; t < 4; ++t) {
        const bool first = t < 2;
        const uint32_t base = first ? b1 : b2, len = first ? len1 : len2;
        const uint32_t leg = (first ? p1 : p2) + (t & 1);
        if (!first && intra && (leg == p1 || leg == p1 + 1)) continue;  // already counted
        const uint32_t fa = base + leg - 1, fb = base + leg;
        const uint32_t oa = leg > 0 ? el[fa] : depot;
        const uint32_t ob = leg < len ? el[fb] : depot;
        const uint32_t na = leg > 0 ? (fa == f1 ? x2 : (fa == f2 ? x1 : oa)) : depot;
        const uint32_t nb = leg < len ? (fb == f1 ? x2 : (fb == f2 ? x1 : ob)) : depot;
        int64_t dl = mat_at(c, na, nb) - mat_at(c, oa, ob);
        if (first) delta1 += dl; else delta2 += dl;
      }
      const int64_t* rcost = (const int64_t*)(st + c.off0);
      if (intra) {
        int64_t oc = rcost[e1];
        add_level(d, c, weight_eval(c.w, oc + delta1 + delta2) - weight_eval(c.w, oc));
      } else {
        int64_t o1 = rcost[e1], o2 = rcost[e2];
        add_level(d, c, weight_eval(c.w, o1 + delta1) - weight_eval(c.w, o1) + weight_eval(c.w, o2 + delta2) -
                            weight_eval(c.w, o2));
      }
    } else if (c.kind == SFGPU_K_LIST_SUM) {
      if (!intra) {
        const int64_t* rsum = (const int64_t*)(st + c.off0);
        int64_t v1 = ((const int64_t*)c.g0)[x1], v2 = ((const int64_t*)c.g0)[x2];
        int64_t s1 = rsum[e1], s2 = rsum[e2];
        add_level(d, c, weight_eval(c.w, s1 - v1 + v2) - weight_eval(c.w, s1) + weight_eval(c.w, s2 - v2 + v1) -
                            weight_eval(c.w, s2));
      }
    }
  }
  return true;
}

// ListReverseMove {entity, start, end}: reverses [start, end) of one list (2-opt segment reversal,
// heuristic/move/list_kernel/reverse.rs:21-58). Only the path cost can change: the two boundary legs are
// replaced and every inner leg is traversed the other way (zero net effect on a symmetric matrix).
__device__ __forceinline__ bool list_reverse_delta(const DevModel& m, const char* st, uint4 row, Score2& d) {
  const uint32_t* off = (const uint32_t*)(st + m.off_offsets);
  const uint32_t* el = (const uint32_t*)(st + m.off_elems);
  d.hard = 0;
  d.soft = 0;
  const uint32_t e = row.x, start = row.y, end = row.z;
  if (e >= m.n_owners) return false;
  const uint32_t b = off[e], len = off[e + 1] - b;
  if (!(end > start + 1 && end <= len)) return false;  // reverse.rs:21-33
  for (uint32_t k = 0; k < m.n_cons; ++k) {
    const ConsDev& c = m.cons[k];
    if (c.kind != SFGPU_K_LIST_PATH_COST) continue;
    const uint32_t depot = (uint32_t)c.p0;
    const uint32_t a = start > 0 ? el[b + start - 1] : depot, z = end < len ? el[b + end] : depot;
    const uint32_t xs = el[b + start], xe = el[b + end - 1];
    int64_t delta = mat_at(c, a, xe) + mat_at(c, xs, z) - mat_at(c, a, xs) - mat_at(c, xe, z);
    for (uint32_t i = start; i + 1 < end; ++i) delta += mat_at(c, el[b + i + 1], el[b + i]) - mat_at(c, el[b + i], el[b + i + 1]);
    const int64_t oc = ((const int64_t*)(st + c.off0))[e];
    add_level(d, c, weight_eval(c.w, oc + delta) - weight_eval(c.w, oc));
  }
  return true;
}

// SublistChangeMove {src_entity, start | size << 24, dst_entity, dst_position}: relocates the contiguous segment
// [start, start + size) (heuristic/move/list_kernel/sublist_change.rs:17-125). For an intra-list move the
// destination is a position of the list AFTER the removal. The segment keeps its direction, so its inner legs
// move with it; per-route sums change by the segment's total.
#define SFGPU_SEG_POS(w) ((w) & 0xFFFFFFu)
#define SFGPU_SEG_SIZE(w) ((w) >> 24)
__device__ __forceinline__ bool list_sublist_change_delta(const DevModel& m, const char* st, uint4 row, Score2& d) {
  const uint32_t* off = (const uint32_t*)(st + m.off_offsets);
  const uint32_t* el = (const uint32_t*)(st + m.off_elems);
  d.hard = 0;
  d.soft = 0;
  const uint32_t se = row.x, start = SFGPU_SEG_POS(row.y), size = SFGPU_SEG_SIZE(row.y), de = row.z, dp = row.w;
  if (se >= m.n_owners || de >= m.n_owners || size == 0) return false;
  const uint32_t sb = off[se], slen = off[se + 1] - sb;
  const uint32_t end = start + size;
  if (end > slen) return false;
  const bool intra = se == de;
  const uint32_t db = off[de], dlen = off[de + 1] - db;
  if (dp > (intra ? slen - size : dlen)) return false;
  if (intra && dp == start) return false;
  const uint32_t first = el[sb + start], last = el[sb + end - 1];
  for (uint32_t k = 0; k < m.n_cons; ++k) {
    const ConsDev& c = m.cons[k];
    if (c.kind == SFGPU_K_LIST_PATH_COST) {
      const uint32_t depot = (uint32_t)c.p0;
      const uint32_t prev = start > 0 ? el[sb + start - 1] : depot;
      const uint32_t next = end < slen ? el[sb + end] : depot;
      const int64_t* rcost = (const int64_t*)(st + c.off0);
      // closing the gap; a route emptied by the move costs 0 (no depot->depot leg)
      const int64_t gap = -mat_at(c, prev, first) - mat_at(c, last, next) + (slen > size ? mat_at(c, prev, next) : 0);
      if (intra) {
        // neighbours of the insertion point in the post-removal list
        const uint32_t ia = dp > 0 ? dp - 1 : 0, ib = dp;
        const uint32_t a = dp > 0 ? el[sb + (ia < start ? ia : ia + size)] : depot;
        const uint32_t b = dp < slen - size ? el[sb + (ib < start ? ib : ib + size)] : depot;
        const int64_t ins = mat_at(c, a, first) + mat_at(c, last, b) - mat_at(c, a, b);
        const int64_t oc = rcost[se];
        add_level(d, c, weight_eval(c.w, oc + gap + ins) - weight_eval(c.w, oc));
      } else {
        int64_t inner = 0;
        for (uint32_t i = start; i + 1 < end; ++i) inner += mat_at(c, el[sb + i], el[sb + i + 1]);
        const uint32_t a = dp > 0 ? el[db + dp - 1] : depot;
        const uint32_t b = dp < dlen ? el[db + dp] : depot;
        const int64_t ins = mat_at(c, a, first) + mat_at(c, last, b) - (dlen > 0 ? mat_at(c, a, b) : 0);
        const int64_t os = rcost[se], od = rcost[de];
        add_level(d, c, weight_eval(c.w, os + gap - inner) - weight_eval(c.w, os) + weight_eval(c.w, od + ins + inner) -
                            weight_eval(c.w, od));
      }
    } else if (c.kind == SFGPU_K_LIST_SUM) {
      if (!intra) {
        const int64_t* rsum = (const int64_t*)(st + c.off0);
        int64_t v = 0;
        for (uint32_t i = start; i < end; ++i) v += ((const int64_t*)c.g0)[el[sb + i]];
        const int64_t ss = rsum[se], ds = rsum[de];
        add_level(d, c, weight_eval(c.w, ss - v) - weight_eval(c.w, ss) + weight_eval(c.w, ds + v) -
                            weight_eval(c.w, ds));
      }
    }
    // EXISTS_FLAT: a relocation keeps the multiset of flattened keys => 0.
  }
  return true;
}

// SublistSwapMove {first_entity, start1 | size1 << 24, second_entity, start2 | size2 << 24}: exchanges two
// contiguous segments, possibly of different sizes (heuristic/move/list_kernel/sublist_swap.rs:17-170). Inside
// one list the segments must not overlap. Both segments keep their direction.
__device__ __forceinline__ bool list_sublist_swap_delta(const DevModel& m, const char* st, uint4 row, Score2& d) {
  const uint32_t* off = (const uint32_t*)(st + m.off_offsets);
  const uint32_t* el = (const uint32_t*)(st + m.off_elems);
  d.hard = 0;
  d.soft = 0;
  const uint32_t e1 = row.x, s1 = SFGPU_SEG_POS(row.y), n1 = SFGPU_SEG_SIZE(row.y);
  const uint32_t e2 = row.z, s2 = SFGPU_SEG_POS(row.w), n2 = SFGPU_SEG_SIZE(row.w);
  if (e1 >= m.n_owners || e2 >= m.n_owners || n1 == 0 || n2 == 0) return false;
  const uint32_t b1 = off[e1], len1 = off[e1 + 1] - b1;
  const uint32_t b2 = off[e2], len2 = off[e2 + 1] - b2;
  const uint32_t t1 = s1 + n1, t2 = s2 + n2;
  if (t1 > len1 || t2 > len2) return false;
  const bool intra = e1 == e2;
  if (intra && s1 < t2 && s2 < t1) return false;
  for (uint32_t k = 0; k < m.n_cons; ++k) {
    const ConsDev& c = m.cons[k];
    if (c.kind == SFGPU_K_LIST_PATH_COST) {
      const uint32_t depot = (uint32_t)c.p0;
      const int64_t* rcost = (const int64_t*)(st + c.off0);
      if (intra) {
        // early segment E = [es, et), late segment L = [ls, lt), gap M = [et, ls) (may be empty)
        const bool first_early = s1 < s2;
        const uint32_t es = first_early ? s1 : s2, et = first_early ? t1 : t2;
        const uint32_t ls = first_early ? s2 : s1, lt = first_early ? t2 : t1;
        const uint32_t pE = es > 0 ? el[b1 + es - 1] : depot, nL = lt < len1 ? el[b1 + lt] : depot;
        const uint32_t fE = el[b1 + es], lE = el[b1 + et - 1], fL = el[b1 + ls], lL = el[b1 + lt - 1];
        int64_t delta = mat_at(c, pE, fL) + mat_at(c, lE, nL) - mat_at(c, pE, fE) - mat_at(c, lL, nL);
        if (et == ls) {
          delta += mat_at(c, lL, fE) - mat_at(c, lE, fL);
        } else {
          const uint32_t m0 = el[b1 + et], mk = el[b1 + ls - 1];
          delta += mat_at(c, lL, m0) + mat_at(c, mk, fE) - mat_at(c, lE, m0) - mat_at(c, mk, fL);
        }
        const int64_t oc = rcost[e1];
        add_level(d, c, weight_eval(c.w, oc + delta) - weight_eval(c.w, oc));
      } else {
        const uint32_t p1 = s1 > 0 ? el[b1 + s1 - 1] : depot, x1 = t1 < len1 ? el[b1 + t1] : depot;
        const uint32_t p2 = s2 > 0 ? el[b2 + s2 - 1] : depot, x2 = t2 < len2 ? el[b2 + t2] : depot;
        const uint32_t f1 = el[b1 + s1], l1 = el[b1 + t1 - 1], f2 = el[b2 + s2], l2 = el[b2 + t2 - 1];
        int64_t in1 = 0, in2 = 0;
        for (uint32_t i = s1; i + 1 < t1; ++i) in1 += mat_at(c, el[b1 + i], el[b1 + i + 1]);
        for (uint32_t i = s2; i + 1 < t2; ++i) in2 += mat_at(c, el[b2 + i], el[b2 + i + 1]);
        const int64_t d1 = mat_at(c, p1, f2) + mat_at(c, l2, x1) + in2 - mat_at(c, p1, f1) - mat_at(c, l1, x1) - in1;
        const int64_t d2 = mat_at(c, p2, f1) + mat_at(c, l1, x2) + in1 - mat_at(c, p2, f2) - mat_at(c, l2, x2) - in2;
        const int64_t o1 = rcost[e1], o2 = rcost[e2];
        add_level(d, c, weight_eval(c.w, o1 + d1) - weight_eval(c.w, o1) + weight_eval(c.w, o2 + d2) - weight_eval(c.w, o2));
      }
    } else if (c.kind == SFGPU_K_LIST_SUM) {
      if (!intra) {
        const int64_t* rsum = (const int64_t*)(st + c.off0);
        int64_t v1 = 0, v2 = 0;
        for (uint32_t i = s1; i < t1; ++i) v1 += ((const int64_t*)c.g0)[el[b1 + i]];
        for (uint32_t i = s2; i < t2; ++i) v2 += ((const int64_t*)c.g0)[el[b2 + i]];
        const int64_t a1 = rsum[e1], a2 = rsum[e2];
        add_level(d, c, weight_eval(c.w, a1 - v1 + v2) - weight_eval(c.w, a1) + weight_eval(c.w, a2 - v2 + v1) -
                            weight_eval(c.w, a2));
      }
    }
  }
  return true;
}

// KOptMove on one list (heuristic/move/k_opt.rs:11-110): cuts c_0 < .. < c_{k-1} split the list into k + 1 segments;
// the reconnection pattern re-orders the middle segments and reverses some of them (k_opt_reconnection.rs:63-262, the
// pattern index counts enumerate_reconnections(k) — every order of the middle segments, every reversal mask, the
// identity dropped). Row {entity | k << 28, c0 | c1 << 16, c2 | c3 << 16, c4 | pattern << 16}. Only the path cost can
// change: the junction legs and, on an asymmetric matrix, the inner legs of the reversed segments.
#define SFGPU_KOPT_MAX 5
struct KOptDecoded {
  uint32_t e, k, bounds[SFGPU_KOPT_MAX + 2], order[SFGPU_KOPT_MAX + 1], rev;  // rev: bit s = segment s reversed
};
__device__ __forceinline__ bool k_opt_decode(const DevModel& m, const uint32_t* off, uint4 row, KOptDecoded& q) {
  q.e = row.x & 0x0FFFFFFFu;
  q.k = row.x >> 28;
  if (q.e >= m.n_owners || q.k < 2 || q.k > SFGPU_KOPT_MAX) return false;
  const uint32_t len = off[q.e + 1] - off[q.e];
  const uint32_t c[5] = {row.y & 0xFFFFu, row.y >> 16, row.z & 0xFFFFu, row.z >> 16, row.w & 0xFFFFu};
  q.bounds[0] = 0;
  for (uint32_t i = 0; i < q.k; ++i) {
    if (c[i] > len || (i > 0 && c[i] <= c[i - 1])) return false;  // k_opt.rs:11-38
    q.bounds[i + 1] = c[i];
  }
  q.bounds[q.k + 1] = len;
  // pattern -> (permutation of the middle segments 1..k-1 in lexicographic order, reversal mask)
  const uint32_t raw = (row.w >> 16) + 1;  // raw index 0 is the identity, which the reference drops
  const uint32_t mid = q.k - 1;
  uint32_t perm = raw >> mid;
  q.rev = (raw & ((1u << mid) - 1)) << 1;
  uint32_t fact = 1;
  for (uint32_t i = 2; i <= mid; ++i) fact *= i;
  if (perm >= fact) return false;
  uint32_t avail[SFGPU_KOPT_MAX];
  for (uint32_t i = 0; i < mid; ++i) avail[i] = i + 1;
  q.order[0] = 0;
  for (uint32_t i = 0; i < mid; ++i) {
    fact /= (mid - i);
    const uint32_t pick = perm / fact;
    perm %= fact;
    q.order[i + 1] = avail[pick];
    for (uint32_t j = pick; j + 1 < mid - i; ++j) avail[j] = avail[j + 1];
  }
  q.order[q.k] = q.k;
  return true;
}
__device__ __forceinline__ bool list_k_opt_delta(const DevModel& m, const char* st, uint4 row, Score2& d) {
  const uint32_t* off = (const uint32_t*)(st + m.off_offsets);
  const uint32_t* el = (const uint32_t*)(st + m.off_elems);
  d.hard = 0;
  d.soft = 0;
  KOptDecoded q;
  if (!k_opt_decode(m, off, row, q)) return false;
  const uint32_t b = off[q.e];
  for (uint32_t kk = 0; kk < m.n_cons; ++kk) {
    const ConsDev& c = m.cons[kk];
    if (c.kind != SFGPU_K_LIST_PATH_COST) continue;
    const uint32_t depot = (uint32_t)c.p0;
    int64_t delta = 0;
    uint32_t new_tail = depot, old_tail = depot;
    bool any = false;
    for (uint32_t pos = 0; pos <= q.k; ++pos) {
      // the old route visits segment `pos`, the new one segment order[pos]
      const uint32_t so = pos, sn = q.order[pos];
      if (q.bounds[so + 1] > q.bounds[so]) {
        delta -= mat_at(c, old_tail, el[b + q.bounds[so]]);
        old_tail = el[b + q.bounds[so + 1] - 1];
        any = true;
      }
      if (q.bounds[sn + 1] > q.bounds[sn]) {
        const bool rv = (q.rev >> sn) & 1;
        const uint32_t lo = b + q.bounds[sn], hi = b + q.bounds[sn + 1] - 1;
        delta += mat_at(c, new_tail, rv ? el[hi] : el[lo]);
        new_tail = rv ? el[lo] : el[hi];
        if (rv)
          for (uint32_t i = lo; i < hi; ++i) delta += mat_at(c, el[i + 1], el[i]) - mat_at(c, el[i], el[i + 1]);
      }
    }
    if (any) delta += mat_at(c, new_tail, depot) - mat_at(c, old_tail, depot);
    const int64_t oc = ((const int64_t*)(st + c.off0))[q.e];
    add_level(d, c, weight_eval(c.w, oc + delta) - weight_eval(c.w, oc));
  }
  return true;
}

enum { LMODE_CHANGE = 0, LMODE_SWAP = 1, LMODE_REVERSE = 2, LMODE_SUBLIST_CHANGE = 3, LMODE_SUBLIST_SWAP = 4, LMODE_K_OPT = 5 };

template <int LMODE, bool STAGED>
__global__ void __launch_bounds__(256) score_list_kernel(const __grid_constant__ DevModel m,
                                                         const uint64_t* __restrict__ cand_offsets,
                                                         const uint32_t* __restrict__ rows,
                                                         int64_t* __restrict__ out_scores,
                                                         uint8_t* __restrict__ out_doable) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  const uint32_t r = blockIdx.y;
  const char* gblock = m.state + (size_t)r * m.block_bytes;
  const char* st = gblock;
  if (STAGED) {
    stage_block(smem, gblock, m.stage_bytes, &bar);
    st = smem;
  }
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  const uint64_t lo = cand_offsets[r], hi = cand_offsets[r + 1];
  for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint4 row = ((const uint4*)rows)[i];  // one 128-bit load per candidate
    Score2 d;
    bool ok = LMODE == LMODE_CHANGE
                  ? list_change_delta(m, st, row, d)
                  : (LMODE == LMODE_SWAP ? list_swap_delta(m, st, row, d)
                                         : (LMODE == LMODE_REVERSE
                                                ? list_reverse_delta(m, st, row, d)
                                                : (LMODE == LMODE_SUBLIST_CHANGE
                                                       ? list_sublist_change_delta(m, st, row, d)
                                                       : (LMODE == LMODE_SUBLIST_SWAP ? list_sublist_swap_delta(m, st, row, d)
                                                                                      : list_k_opt_delta(m, st, row, d)))));
    longlong2 o;
    o.x = ok ? ch + d.hard : 0;
    o.y = ok ? csf + d.soft : 0;
    ((longlong2*)out_scores)[i] = o;
    out_doable[i] = ok ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Fast list-change path. Preconditions (checked at commit, DevModel::fast_list): at most one
// LIST_PATH_COST constraint with a LINEAR weight over an int32 matrix whose cells are < 2^28, at
// most one LIST_SUM constraint whose column fits int32, any number of EXISTS_FLAT constraints
// (relocations never change their score). The replica block then carries packed records that
// are rebuilt whenever a move is committed, and one candidate costs four 128-bit LDS, two matrix
// gathers, one 128-bit LDG and one 128-bit STG.
// ---------------------------------------------------------------------------------------------
// block-parallel over insertion slots (one thread per slot g = base + owner + position, so the record stores are
// coalesced and every thread waits for one round of gathers); s_off = shared scratch of n_owners + 1 words
// s_el (optional) = a shared-memory copy of the element array: every slot reads three neighbouring elements before its
// matrix gathers, and the copy takes that dependent global round trip out of each slot's chain
__device__ __forceinline__ void build_fast_records(const DevModel& m, char* st, uint32_t* s_off,
                                                   const uint32_t* s_el = nullptr) {
  const uint32_t* off = (const uint32_t*)(st + m.off_offsets);
  const uint32_t* el = s_el ? s_el : (const uint32_t*)(st + m.off_elems);
  RouteRec* rr = (RouteRec*)(st + m.off_route_rec);
  PosRec* pr = (PosRec*)(st + m.off_pos_rec);
  SlotRec* sr = (SlotRec*)(st + m.off_slot_rec);
  const ConsDev* pc = m.fast_pc >= 0 ? &m.cons[m.fast_pc] : nullptr;
  const ConsDev* ls = m.fast_ls >= 0 ? &m.cons[m.fast_ls] : nullptr;
  const int32_t* mat = pc ? (const int32_t*)pc->g0 : nullptr;
  const uint32_t dim = pc ? pc->n0 : 0;
  const uint32_t depot = pc ? (uint32_t)pc->p0 : 0;
  uint32_t* pos_of = m.nearby_ok ? (uint32_t*)(st + m.off_pos_of) : nullptr;
  const uint32_t* __restrict__ rl = m.relabel;
  uint2* r8 = m.compact_bytes ? (uint2*)(st + m.off_route8) : nullptr;
  uint2* p8 = m.compact_bytes ? (uint2*)(st + m.off_pos8) : nullptr;
  uint2* s8 = m.compact_bytes ? (uint2*)(st + m.off_slot8) : nullptr;
  const uint32_t n_owners = m.n_owners;
  for (uint32_t o = threadIdx.x; o <= n_owners; o += blockDim.x) s_off[o] = off[o];
  if (pos_of)
    for (uint32_t i = threadIdx.x; i < m.n_elem_rows; i += blockDim.x) pos_of[i] = 0xFFFFFFFFu;
  __syncthreads();
  const uint32_t n_slots = s_off[n_owners] + n_owners;
#pragma unroll 2
  for (uint32_t g = threadIdx.x; g < n_slots; g += blockDim.x) {
    uint32_t lo = 0, hi = n_owners;  // owner of slot g: the largest o with s_off[o] + o <= g
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_off[mid] + mid <= g) lo = mid; else hi = mid;
    }
    const uint32_t o = lo, b = s_off[o], len = s_off[o + 1] - b, p = g - b - o;
    const uint32_t a_el = p > 0 ? el[b + p - 1] : depot;
    const uint32_t b_el = p < len ? el[b + p] : depot;
    SlotRec sl;
    sl.a = rl ? rl[a_el] : a_el;  // ids stored in records address the relabelled matrices only
    sl.b = rl ? rl[b_el] : b_el;
    sl.gap = (pc && len > 0) ? mat[a_el * dim + b_el] : 0;
    sl.where = (o << 16) | p;
    sr[g] = sl;
    if (s8) s8[g] = make_uint2(sl.a | (sl.b << 16), (uint32_t)sl.gap);
    if (p < len) {
      const uint32_t x = b_el;
      const uint32_t nx = p + 1 < len ? el[b + p + 1] : depot;
      PosRec q;
      q.elem = rl ? rl[x] : x;
      q.rem = pc ? (-mat[a_el * dim + x] - mat[x * dim + nx] + (len > 1 ? mat[a_el * dim + nx] : 0)) : 0;
      q.val = ls ? (int32_t)((const int64_t*)ls->g0)[x] : 0;
      q.owner = o;
      pr[b + p] = q;
      if (p8) p8[b + p] = make_uint2(q.elem | ((uint32_t)(uint16_t)(int16_t)q.val << 16), (uint32_t)q.rem);
      if (pos_of) pos_of[q.elem] = (o << 16) | p;
    }
  }
  // per-route records; the per-route sum is the retained LIST_SUM aggregate the caller has just brought up to date
  const int64_t* rsum = ls ? (const int64_t*)(st + ls->off0) : nullptr;
  for (uint32_t o = threadIdx.x; o < n_owners; o += blockDim.x) {
    const uint32_t b = s_off[o], len = s_off[o + 1] - b;
    RouteRec r;
    r.base = b;
    r.len = len;
    r.sum = rsum ? rsum[o] : 0;
    rr[o] = r;
    if (r8) r8[o] = make_uint2(b | (len << 16), (uint32_t)(int32_t)r.sum);
  }
}

// acceptor predicate on a score (1 HillClimbing: move > last; 2 LateAcceptance: >= last || >= late;
// 3 GreatDeluge form: > last || >= threshold; see sfgpu_forage_params)
__device__ __forceinline__ bool accept_score(int acceptor, int64_t h, int64_t s, int64_t lh, int64_t ls, int64_t th,
                                             int64_t ts) {
  if (acceptor == 0) return true;
  if (acceptor == 1) return score_less(lh, ls, h, s);
  if (acceptor == 3) return score_less(lh, ls, h, s) || !score_less(h, s, th, ts);  // > last || >= threshold
  return !score_less(h, s, lh, ls) || !score_less(h, s, th, ts);
}

// Arithmetic width of the fast list kernels. uint16 cells are only selected for narrow models (every
// fast-path delta provably inside int32, sfgpu_api.cu commit), so those kernels run the whole delta,
// the acceptor test and the running best on 32-bit registers; wider models keep int64. Products and
// sums go through the unsigned type (two's-complement wrap, exact whenever the result fits).
template <typename CELL>
struct FastArith {
  typedef int64_t S;
  typedef uint64_t US;
};
template <>
struct FastArith<uint16_t> {
  typedef int32_t S;
  typedef uint32_t US;
};
template <typename S>
__device__ __forceinline__ S clamp_to(int64_t v, int64_t lim) {
  return (S)(v < -lim ? -lim : (v > lim ? lim : v));
}
// 32-bit threshold of an acceptor reference score relative to the committed score: deltas are
// strictly inside (INT32_MIN, INT32_MAX), so a clamped threshold compares like the exact one
__device__ __forceinline__ void rel_threshold(int64_t ref, int64_t committed, int32_t& out) {
  const int64_t d = ref - committed;
  out = d < (int64_t)INT32_MIN ? INT32_MIN : (d > (int64_t)INT32_MAX ? INT32_MAX : (int32_t)d);
}
__device__ __forceinline__ void rel_threshold(int64_t ref, int64_t committed, int64_t& out) { out = ref - committed; }
template <typename S>
__device__ __forceinline__ bool lex_less(S h1, S s1, S h2, S s2) {
  return h1 != h2 ? h1 < h2 : s1 < s2;
}
// acceptor predicate on score deltas against thresholds relative to the committed score
template <typename S>
__device__ __forceinline__ bool accept_delta(int acceptor, S h, S s, S lh, S ls, S th, S ts) {
  if (acceptor == 0) return true;
  if (acceptor == 1) return lex_less(lh, ls, h, s);
  if (acceptor == 3) return lex_less(lh, ls, h, s) || !lex_less(h, s, th, ts);
  return !lex_less(h, s, lh, ls) || !lex_less(h, s, th, ts);
}
// LIST_SUM delta of moving value v from a route with sum ss to one with sum sd (a, b pre-narrowed)
template <int SUM_FN, typename S, typename US>
__device__ __forceinline__ S list_sum_delta(S a, S b, S v, S ss, S sd) {
  if (SUM_FN == SFGPU_W_EXCESS) {
    const S e0 = (S)((US)ss - (US)b), e1 = (S)((US)sd - (US)b);
    const S z = 0;
    const US d = ((US)max((S)((US)e0 - (US)v), z) - (US)max(e0, z)) + ((US)max((S)((US)e1 + (US)v), z) - (US)max(e1, z));
    return (S)(d * (US)a);
  }
  if (SUM_FN == SFGPU_W_SQUARE) {
    const US s0 = (US)ss, s1 = (US)sd, uv = (US)v;
    return (S)((US)a * (((s0 - uv) * (s0 - uv) - s0 * s0) + ((s1 + uv) * (s1 + uv) - s1 * s1)));
  }
  return 0;  // LINEAR: -a*v + a*v = 0; CONST: the number of routes is unchanged
}

struct ForageArgs {
  ForageDev f;
  const int64_t* ref_scores;  // [R][4] = {last_step, late} or null
  ChunkPartial* partials;     // [R][gridDim.x]
  int32_t row_kind;           // rows the finish kernel re-derives when scores were not materialised:
                              // 0 = ListChange {se, sp, de, dp}, 1 = ScalarEdit {entity, to_value}
};

template <int SUM_FN /* -1: no LIST_SUM constraint */, int U = 2, int MINB = 3, bool FORAGE = false, typename CELL = int32_t,
          bool COMPACT = false>
__global__ void __launch_bounds__(256, MINB) score_list_change_fast_kernel(const __grid_constant__ DevModel m,
                                                                     const uint64_t* __restrict__ cand_offsets,
                                                                     const uint32_t* __restrict__ rows,
                                                                     int64_t* __restrict__ out_scores,
                                                                     uint8_t* __restrict__ out_doable,
                                                                     const ForageArgs fa) {
  typedef typename FastArith<CELL>::S S;
  typedef typename FastArith<CELL>::US US;
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  const uint32_t r = blockIdx.y;
  if (COMPACT)  // committed score (first 16 bytes of the block) + the 8-byte records
    stage_block2(smem, m.state + (size_t)r * m.block_bytes + m.off_score, 16, smem + 16,
                 m.state + (size_t)r * m.block_bytes + m.off_route8, m.compact_bytes, &bar);
  else
    stage_block(smem, m.state + (size_t)r * m.block_bytes, m.fast_score_bytes, &bar);
  const int64_t* cs = (const int64_t*)(smem + (COMPACT ? 0 : m.off_score));
  const int64_t ch = cs[0], csf = cs[1];
  // fused forager partial (FORAGE): this thread's best accepted score delta, multiplicity, first
  // row; acceptor references become thresholds on the delta
  S tb_h = 0, tb_s = 0, f_lh = 0, f_ls = 0, f_th = 0, f_ts = 0;
  uint32_t tb_n = 0, tb_first = 0xFFFFFFFFu, tb_second = 0xFFFFFFFFu, t_acc = 0;
  if (FORAGE && fa.ref_scores) {
    rel_threshold(fa.ref_scores[r * 4 + 0], ch, f_lh);
    rel_threshold(fa.ref_scores[r * 4 + 1], csf, f_ls);
    rel_threshold(fa.ref_scores[r * 4 + 2], ch, f_th);
    rel_threshold(fa.ref_scores[r * 4 + 3], csf, f_ts);
  } else if (FORAGE) {
    rel_threshold(0, ch, f_lh);
    rel_threshold(0, csf, f_ls);
    rel_threshold(0, ch, f_th);
    rel_threshold(0, csf, f_ts);
  }
  // one branch-free form for every acceptor code: accept(d) = (A < d) || (d >= B), lexicographic.
  //   0 accept all: A = +inf, B = -inf      1 d > last: A = last, B = +inf
  //   2 d >= last || d >= t: A = +inf, B = min(last, t)      3 d > last || d >= t: A = last, B = t
  // (deltas are strictly inside the representable range, so +-inf never compare equal)
  const S S_MAX = sizeof(S) == 4 ? (S)INT32_MAX : (S)INT64_MAX, S_MIN = sizeof(S) == 4 ? (S)INT32_MIN : (S)INT64_MIN;
  S a_h = S_MAX, a_s = S_MAX, b_h = S_MIN, b_s = S_MIN;
  if (FORAGE) {
    const int acc = fa.f.acceptor;
    if (acc == 1 || acc == 3) {
      a_h = f_lh;
      a_s = f_ls;
    }
    if (acc == 1) {
      b_h = S_MAX;
      b_s = S_MAX;
    } else if (acc == 2) {
      const bool l_lt = lex_less<S>(f_lh, f_ls, f_th, f_ts);
      b_h = l_lt ? f_lh : f_th;
      b_s = l_lt ? f_ls : f_ts;
    } else if (acc == 3) {
      b_h = f_th;
      b_s = f_ts;
    }
  }
  const uint4* rr = (const uint4*)(smem + m.off_route_rec);
  const uint4* pr = (const uint4*)(smem + m.off_pos_rec);
  const uint4* sr = (const uint4*)(smem + m.off_slot_rec);
  const uint2* rr8 = (const uint2*)(smem + 16);
  const uint2* pr8 = (const uint2*)(smem + 16 + (m.off_pos8 - m.off_route8));
  const uint2* sr8 = (const uint2*)(smem + 16 + (m.off_slot8 - m.off_route8));
  const uint32_t n_owners = m.n_owners;
  // path cost: LINEAR weight => weight(old + d) - weight(old) = a * d
  const bool has_pc = m.fast_pc >= 0;
  const ConsDev& pc = m.cons[has_pc ? m.fast_pc : 0];
  const CELL* __restrict__ mrow = (const CELL*)m.fm_row;
  const CELL* __restrict__ mcol = (const CELL*)m.fm_col;
  const uint32_t dim = pc.n0;
  const S pc_a = has_pc ? (S)(pc.sign < 0 ? -pc.w.a : pc.w.a) : 0;
  const bool pc_hard = pc.w.level == 0;
  // level routing as 0/1 multipliers (one IMAD per level instead of a select + add)
  const S pc_mh = pc_hard ? 1 : 0, pc_ms = pc_hard ? 0 : 1;
  const ConsDev& ls = m.cons[SUM_FN >= 0 ? m.fast_ls : 0];
  // narrow models: |route sum| < 2^30, so a threshold clamped to +-2^30 gives the same excesses
  const S ls_a = (S)(ls.sign < 0 ? -ls.w.a : ls.w.a);
  const S ls_b = sizeof(S) == 4 ? clamp_to<S>(ls.w.b, 1ll << 30) : (S)ls.w.b;
  const bool ls_hard = ls.w.level == 0;
  const S ls_mh = ls_hard ? 1 : 0, ls_ms = ls_hard ? 0 : 1;
  // contiguous chunk per CTA (keeps pull order inside a chunk for the fused forager partials);
  // inside the chunk every index is 32-bit (a replica's pull indices are 32-bit by contract)
  const uint64_t lo = cand_offsets[r], hi = cand_offsets[r + 1];
  const uint64_t per = (hi - lo + gridDim.x - 1) / gridDim.x;
  const uint64_t c_lo64 = lo + per * blockIdx.x < hi ? lo + per * blockIdx.x : hi;
  const uint64_t c_hi64 = c_lo64 + per < hi ? c_lo64 + per : hi;
  const uint32_t n_c = (uint32_t)(c_hi64 - c_lo64);
  const uint32_t first_base = (uint32_t)(c_lo64 - lo);
  // U candidates per thread per trip with the next trip's rows prefetched into registers: the
  // DRAM latency of the streamed rows is hidden by 2*U independent 128-bit loads per thread.
  const uint4* __restrict__ rows4 = (const uint4*)rows + c_lo64;
  longlong2* __restrict__ scores_c = out_scores ? (longlong2*)out_scores + c_lo64 : nullptr;
  uint8_t* __restrict__ doable_c = out_doable ? out_doable + c_lo64 : nullptr;
  const uint32_t stride = blockDim.x * U;
  uint4 nxt[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t i = threadIdx.x + u * blockDim.x;
    nxt[u] = i < n_c ? __ldcs(rows4 + i) : make_uint4(0xFFFFFFFFu, 0, 0xFFFFFFFFu, 0);
  }
  for (uint32_t base = threadIdx.x; base < n_c; base += stride) {
    uint4 cur[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      cur[u] = nxt[u];
      const uint32_t i = base + stride + u * blockDim.x;
      nxt[u] = i < n_c ? __ldcs(rows4 + i) : make_uint4(0xFFFFFFFFu, 0, 0xFFFFFFFFu, 0);
    }
    // phase 1: route records + doability
    bool ok[U];
    uint32_t pidx[U], sidx[U];
    S sums[U], sumd[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t se = cur[u].x, sp = cur[u].y, de = cur[u].z, dp = cur[u].w;
      ok[u] = se < n_owners && de < n_owners;
      uint4 rs, rd;  // {base, len, sum lo, sum hi}
      if (COMPACT) {
        const uint2 a8 = rr8[ok[u] ? se : 0], b8 = rr8[ok[u] ? de : 0];
        rs = make_uint4(a8.x & 0xFFFFu, a8.x >> 16, a8.y, 0);
        rd = make_uint4(b8.x & 0xFFFFu, b8.x >> 16, b8.y, 0);
      } else {
        rs = rr[ok[u] ? se : 0];
        rd = rr[ok[u] ? de : 0];
      }
      ok[u] = ok[u] && sp < rs.y && dp <= rd.y && !(se == de && (dp == sp || dp == sp + 1));
      pidx[u] = ok[u] ? rs.x + sp : 0;
      sidx[u] = ok[u] ? rd.x + de + dp : 0;
      if (SUM_FN >= 0) {
        sums[u] = sizeof(S) == 4 ? (S)rs.z : (S)(((uint64_t)rs.w << 32) | rs.z);
        sumd[u] = sizeof(S) == 4 ? (S)rd.z : (S)(((uint64_t)rd.w << 32) | rd.z);
      }
    }
    // phase 2: position / slot records, then the two matrix gathers
    uint4 p[U], sl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (COMPACT) {
        const uint2 p8 = pr8[pidx[u]], s8 = sr8[sidx[u]];
        p[u] = make_uint4(p8.x & 0xFFFFu, p8.y, (uint32_t)(int32_t)(int16_t)(p8.x >> 16), 0);  // {elem, rem, val}
        sl[u] = make_uint4(s8.x & 0xFFFFu, s8.x >> 16, s8.y, 0);                              // {a, b, gap}
      } else {
        p[u] = pr[pidx[u]];
        sl[u] = sr[sidx[u]];
      }
    }
    int32_t m0[U], m1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // both gathers address row x (of the transpose and of the matrix): the candidates of one
      // source share two short rows, so a warp's gathers fall into few 128 B lines
      m0[u] = has_pc ? (int32_t)__ldg(mcol + p[u].x * dim + sl[u].x) : 0;
      m1[u] = has_pc ? (int32_t)__ldg(mrow + p[u].x * dim + sl[u].y) : 0;
    }
    // phase 3: deltas + stores
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t i = base + u * blockDim.x;
      if (i >= n_c) continue;
      S dh = 0, ds = 0;
      if (has_pc) {
        const US d = (US)pc_a * (US)(S)((int32_t)p[u].y + m0[u] + m1[u] - (int32_t)sl[u].z);
        dh = (S)(d * (US)pc_mh);
        ds = (S)(d * (US)pc_ms);
      }
      if (SUM_FN >= 0 && cur[u].x != cur[u].z) {
        const US d = (US)list_sum_delta<SUM_FN, S, US>(ls_a, ls_b, (S)(int32_t)p[u].z, sums[u], sumd[u]);
        dh = (S)((US)dh + d * (US)ls_mh);
        ds = (S)((US)ds + d * (US)ls_ms);
      }
      if (!FORAGE || out_scores) {
        longlong2 o;
        o.x = ok[u] ? ch + (int64_t)dh : 0;
        o.y = ok[u] ? csf + (int64_t)ds : 0;
        __stcs(scores_c + i, o);
        doable_c[i] = ok[u] ? 1 : 0;
      }
      if (FORAGE && ok[u] && (lex_less<S>(a_h, a_s, dh, ds) || !lex_less<S>(dh, ds, b_h, b_s))) {
        t_acc++;
        if (tb_n == 0 || lex_less<S>(tb_h, tb_s, dh, ds)) {
          tb_h = dh;
          tb_s = ds;
          tb_n = 1;
          tb_first = first_base + i;
          tb_second = 0xFFFFFFFFu;
        } else if (tb_h == dh && tb_s == ds) {
          // rows of one thread are visited in increasing pull order: tb_first stays the minimum
          if (tb_n == 1) tb_second = first_base + i;
          tb_n++;
        }
      }
    }
  }
  if (FORAGE) {
    // merge: better score wins; equal scores add multiplicities and keep the earliest row
    __shared__ int64_t sh_h[8], sh_s[8];
    __shared__ uint32_t sh_n[8], sh_f[8], sh_a[8];
    for (int o = 16; o > 0; o >>= 1) {
      const S oh = __shfl_down_sync(0xffffffffu, tb_h, o), os = __shfl_down_sync(0xffffffffu, tb_s, o);
      const uint32_t on = __shfl_down_sync(0xffffffffu, tb_n, o), of = __shfl_down_sync(0xffffffffu, tb_first, o);
      const uint32_t osec = __shfl_down_sync(0xffffffffu, tb_second, o);
      t_acc += __shfl_down_sync(0xffffffffu, t_acc, o);
      if (on && (!tb_n || lex_less<S>(tb_h, tb_s, oh, os))) {
        tb_h = oh; tb_s = os; tb_n = on; tb_first = of; tb_second = osec;
      } else if (on && tb_n && oh == tb_h && os == tb_s) {
        tb_n += on;
        tb_second = min(max(tb_first, of), min(tb_second, osec));  // second smallest of the four indices
        tb_first = min(tb_first, of);
      }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ uint32_t sh_2[8];
    if (lane == 0) {
      sh_h[warp] = ch + (int64_t)tb_h; sh_s[warp] = csf + (int64_t)tb_s; sh_n[warp] = tb_n; sh_f[warp] = tb_first;
      sh_a[warp] = t_acc;
      sh_2[warp] = tb_second;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      ChunkPartial cp{0, 0, 0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu};
      for (int w = 0; w < 8; ++w) {
        cp.n_accepted += sh_a[w];
        if (!sh_n[w]) continue;
        if (!cp.n_best || score_less(cp.best_h, cp.best_s, sh_h[w], sh_s[w])) {
          cp.best_h = sh_h[w]; cp.best_s = sh_s[w]; cp.n_best = sh_n[w]; cp.first_idx = sh_f[w];
          cp.second_idx = sh_2[w];
        } else if (sh_h[w] == cp.best_h && sh_s[w] == cp.best_s) {
          cp.n_best += sh_n[w];
          cp.second_idx = min(max(cp.first_idx, sh_f[w]), min(cp.second_idx, sh_2[w]));
          cp.first_idx = min(cp.first_idx, sh_f[w]);
        }
      }
      fa.partials[(size_t)r * gridDim.x + blockIdx.x] = cp;
    }
  }
}

// =============================================================================================
// block reductions
// =============================================================================================
__device__ __forceinline__ int64_t block_sum_i64(int64_t v, int64_t* scratch /* >= 32 */) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = threadIdx.x < nw ? scratch[threadIdx.x] : 0;
  if (warp == 0)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (threadIdx.x == 0) scratch[0] = v;
  __syncthreads();
  v = scratch[0];
  __syncthreads();
  return v;
}

// inclusive block scan of a 32-bit count; returns this thread's inclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_scan_u32(uint32_t v, uint32_t* scratch /* >= 33 */, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  __syncthreads();
  if (lane == 31) scratch[warp] = x;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  if (warp == 0) {
    uint32_t w = lane < nw ? scratch[lane] : 0;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    scratch[lane] = w;
  }
  __syncthreads();
  uint32_t base = warp > 0 ? scratch[warp - 1] : 0;
  *total = scratch[nw - 1];
  __syncthreads();
  return x + base;
}

