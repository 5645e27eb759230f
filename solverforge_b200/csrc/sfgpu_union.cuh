// Union of list neighbourhoods in the reference's seeded pull order, speculated window by window.
//
// Reference: the default list local search opens ONE cursor per step — a VecUnionSelector over the capability-
// matched families (runtime/compiler/default_local_search/policy/list.rs:24-33: nearby change, nearby swap,
// sublist change, sublist swap, reverse, ...), leaves in seeded SelectionOrder::Random, children interleaved by
// UnionSelectionOrder::StratifiedRandom (heuristic/selector/decorator/vec_union.rs:204-366) — and pulls
// candidates one by one until the forager quits (AcceptedCount(N), phase/candidates.rs:66). Only a PREFIX of the
// union stream is ever evaluated, so scoring whole neighbourhoods is waste. Here:
//
//   1. walk kernels: every child cursor is walked in its own pull order (MoveStreamContext::selection_index,
//      move_selector/iter.rs:109-125, pure splitmix64) and emits its first `window` candidate rows — the
//      nearby families with one warp per source (static neighbour-list walk + warp top-K, ranks instead of
//      entity ids so that the with-replacement entity order of SelectionOrder::Random is exact), the
//      index families with one sequential walker per (replica, child);
//   2. union_schedule_kernel: the UnionScheduler replayed per replica over the emitted counts -> union pull t
//      = (child, child-local index), stopping where a child's window (not its cursor) runs out;
//   3. union_score_kernel: pull t of replica r scored with the generic read-only list deltas, materialised in
//      union pull order; argbest_kernel replays acceptor + forager (AcceptedCount cut, reservoir ties);
//   4. union_pick_kernel: the step is complete when the forager quit inside the window or every child cursor
//      ended; otherwise the host / the resident loop re-runs the replica with a larger window.
//
// CandidateId of the union cursor = union pull index (vec_union.rs:447-455).
#pragma once
#include "sfgpu_nearby.cuh"

// x % len for a 64-bit hash and a 32-bit length. Lengths below 2^16 (entity counts, route lengths, value ranges) take
// three 32-bit remainders instead of the software 64-bit division: x = hi * 2^32 + lo, so
// x mod len = ((hi mod len) * (2^32 mod len) + lo mod len) mod len with every product below 2^32.
__device__ __forceinline__ uint32_t mod_u64(uint64_t x, uint32_t len) {
  if (len < 65536u) {
    const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
    const uint32_t c = (0u - len) % len;  // 2^32 mod len in 32-bit arithmetic
    return ((hi % len) * c + lo % len) % len;
  }
  return (uint32_t)(x % (uint64_t)len);
}

// MoveStreamContext (move_selector/iter.rs:14-207)
struct StreamCtx {
  uint64_t step_index, step_seed;
  int order;
  __device__ __forceinline__ uint64_t mixed(uint64_t salt) const {
    return splitmix64_dev(step_seed ^ (step_index * 0x9E3779B97F4A7C15ull) ^ salt);
  }
  __device__ __forceinline__ uint32_t random_index(uint32_t len, uint64_t salt) const {
    return len <= 1 ? 0u : mod_u64(mixed(salt), len);
  }
  __device__ __forceinline__ uint32_t random_stride(uint32_t len, uint64_t salt) const {
    if (len <= 1) return 1;
    uint32_t stride = mod_u64(mixed(salt), len - 1) + 1;
    while (true) {
      uint32_t a = stride, b = len;
      while (b) {
        const uint32_t t = a % b;
        a = b;
        b = t;
      }
      if (a == 1) break;
      stride = stride == len - 1 ? 1 : stride + 1;
    }
    return stride;
  }
};
__device__ __forceinline__ StreamCtx stream_ctx(const UnionArgs& a, uint32_t r) {
  StreamCtx c;
  c.step_seed = a.step_seeds ? a.step_seeds[r] : 0;
  c.step_index = a.step_indices ? a.step_indices[a.step_index_shared ? 0 : r] : 0;
  c.order = a.order;
  return c;
}

// rows a child of replica r may emit in this pass
__device__ __forceinline__ uint32_t union_cap(const UnionArgs& a, uint32_t r) {
  if (!a.win_r) return a.window;
  const uint64_t w = (uint64_t)a.win_r[r] << a.win_shift;
  return w < a.window ? (uint32_t)w : a.window;
}

// selection_index(offset, len, salt) for one (len, salt): the Shuffled start / stride are computed once
struct SelMap {
  const StreamCtx* c;
  uint64_t salt;
  uint32_t len, start, stride;
  __device__ __forceinline__ SelMap(const StreamCtx& cx, uint32_t len_, uint64_t salt_) : c(&cx), salt(salt_), len(len_) {
    start = 0;
    stride = 1;
    if (cx.order == SFGPU_ORDER_SHUFFLED) {
      start = cx.random_index(len, salt);
      stride = cx.random_stride(len, salt ^ 0xA24BAED4963EE407ull);
    }
  }
  __device__ __forceinline__ uint32_t at(uint32_t offset) const {
    if (c->order == SFGPU_ORDER_RANDOM) return c->random_index(len, salt ^ ((uint64_t)offset * 0xD1B54A32D192ED03ull));
    if (c->order == SFGPU_ORDER_SHUFFLED) return mod_u64((uint64_t)start + (uint64_t)offset * stride, len);
    return offset;
  }
};

// ---- sequential walkers (one thread per replica and child) ---------------------------------------------------

struct RowSink {
  uint4* rows;
  uint32_t cap, n;
  bool more;
  __device__ __forceinline__ bool push(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {  // false: window full
    if (n == cap) {
      more = true;
      return false;
    }
    rows[n++] = make_uint4(a, b, c, d);
    return true;
  }
};

// ReverseCursor: heuristic/selector/list_reverse.rs:139-172 + list_kernel/reverse.rs:66-108
__device__ inline void walk_reverse(const StreamCtx& cx, uint32_t desc, const uint32_t* off, uint32_t n, RowSink& out) {
  const SelMap em(cx, n, 0x11572A0700000001ull ^ (uint64_t)desc);
  for (uint32_t o = 0; o < n; ++o) {
    const uint32_t e = em.at(o);
    const uint32_t len = off[e + 1] - off[e];
    if (len < 2) continue;
    const SelMap sm(cx, len, 0x11572A0700000002ull ^ (uint64_t)e ^ (uint64_t)desc);
    for (uint32_t so = 0; so < len; ++so) {
      const uint32_t start = sm.at(so);
      const uint32_t end_count = len > start + 1 ? len - (start + 1) : 0;
      const SelMap zm(cx, end_count, 0x11572A0700000003ull ^ (uint64_t)e ^ (uint64_t)start);
      for (uint32_t eo = 0; eo < end_count; ++eo)
        if (!out.push(e, start, start + 2 + zm.at(eo), 0)) return;
    }
  }
}

// SublistChangeCursor: heuristic/selector/sublist_change.rs:166-205 + list_kernel/sublist_change.rs:103-268
__device__ inline void walk_sublist_change(const StreamCtx& cx, uint32_t desc, const uint32_t* off, uint32_t n,
                                           uint32_t min_size, uint32_t max_size, RowSink& out) {
  const SelMap em(cx, n, 0x5B157C4A46E00001ull ^ (uint64_t)desc);
  for (uint32_t si = 0; si < n; ++si) {
    const uint32_t se = n <= 1 ? si : em.at(si);
    const uint32_t slen = off[se + 1] - off[se];
    if (slen < min_size) continue;
    const SelMap sm(cx, slen, 0x5B157C4A46E00002ull ^ (uint64_t)se ^ (uint64_t)desc);
    for (uint32_t so = 0; so < slen; ++so) {
      const uint32_t start = sm.at(so);
      const uint32_t max_valid = min(max_size, slen - start);
      const uint32_t size_count = max_valid >= min_size ? max_valid - min_size + 1 : 0;
      const SelMap zm(cx, size_count, 0x5B157C4A46E00003ull ^ (uint64_t)se ^ (uint64_t)start);
      for (uint32_t zo = 0; zo < size_count; ++zo) {
        const uint32_t size = min_size + zm.at(zo);
        const uint32_t post = slen - size;
        const uint32_t seg = SFGPU_SEG(start, size);
        const SelMap im(cx, post + 1, 0x5B157C4A46E00004ull ^ (uint64_t)se ^ (uint64_t)start);
        for (uint32_t po = 0; po <= post; ++po) {
          const uint32_t dp = im.at(po);
          if (dp == start) continue;
          if (!out.push(se, seg, se, dp)) return;
        }
        for (uint32_t di = 0; di < n; ++di) {
          if (di == si) continue;
          const uint32_t de = n <= 1 ? di : em.at(di);
          const uint32_t dlen = off[de + 1] - off[de];
          const SelMap xm(cx, dlen + 1, 0x5B157C4A46E00005ull ^ (uint64_t)se ^ (uint64_t)de ^ (uint64_t)start);
          for (uint32_t po = 0; po <= dlen; ++po)
            if (!out.push(se, seg, de, xm.at(po))) return;
        }
      }
    }
  }
}

// SublistSwapCursor: heuristic/selector/sublist_swap.rs:160-190 + list_kernel/sublist_swap.rs:28-318. The segments
// of rank idx come in (start order, size order); a first segment pairs with the segments of its own rank that start
// at or after its end, then with every segment of the later ranks.
__device__ inline void walk_sublist_swap(const StreamCtx& cx, uint32_t desc, const uint32_t* off, uint32_t n,
                                         uint32_t min_size, uint32_t max_size, RowSink& out) {
  const SelMap em(cx, n, 0x5B1575A090000001ull ^ (uint64_t)desc);
  for (uint32_t fi = 0; fi < n; ++fi) {
    const uint32_t e1 = n <= 1 ? fi : em.at(fi);
    const uint32_t len1 = off[e1 + 1] - off[e1];
    if (len1 < min_size) continue;
    const SelMap sm1(cx, len1, 0x5B1575A090000002ull ^ (uint64_t)e1 ^ (uint64_t)desc);
    for (uint32_t so1 = 0; so1 < len1; ++so1) {
      const uint32_t s1 = sm1.at(so1);
      const uint32_t mv1 = min(max_size, len1 - s1);
      if (mv1 < min_size) continue;
      const uint32_t cnt1 = mv1 - min_size + 1;
      const SelMap zm1(cx, cnt1, 0x5B1575A090000003ull ^ (uint64_t)e1 ^ (uint64_t)s1);
      for (uint32_t zo1 = 0; zo1 < cnt1; ++zo1) {
        const uint32_t t1 = s1 + min_size + zm1.at(zo1);
        for (uint32_t si = fi; si < n; ++si) {
          const uint32_t e2 = n <= 1 ? si : em.at(si);
          const uint32_t len2 = off[e2 + 1] - off[e2];
          if (len2 < min_size) continue;
          const SelMap sm2(cx, len2, 0x5B1575A090000002ull ^ (uint64_t)e2 ^ (uint64_t)desc);
          for (uint32_t so2 = 0; so2 < len2; ++so2) {
            const uint32_t s2 = sm2.at(so2);
            const uint32_t mv2 = min(max_size, len2 - s2);
            if (mv2 < min_size) continue;
            const uint32_t cnt2 = mv2 - min_size + 1;
            const SelMap zm2(cx, cnt2, 0x5B1575A090000003ull ^ (uint64_t)e2 ^ (uint64_t)s2);
            for (uint32_t zo2 = 0; zo2 < cnt2; ++zo2) {
              const uint32_t t2 = s2 + min_size + zm2.at(zo2);
              if (fi == si && (s2 < t1 || (s1 == s2 && t1 == t2))) continue;
              if (!out.push(e1, SFGPU_SEG(s1, t1 - s1), e2, SFGPU_SEG(s2, t2 - s2))) return;
            }
          }
        }
      }
    }
  }
}

// KOptCursor: heuristic/selector/list_kernel/k_opt/full.rs:34-98 — entities in the without-replacement stream order
// (iter.rs:130-147: Random and Shuffled both walk start + offset * stride), per entity the (cut combination rank) x
// (reconnection pattern) product pulled through selection_index; cut_combination_at: k_opt/iterators.rs:118-176.
__device__ __forceinline__ uint64_t binomial_dev(uint32_t n, uint32_t k) {
  if (k > n) return 0;
  if (k > n - k) k = n - k;
  uint64_t r = 1;
  for (uint32_t i = 0; i < k; ++i) r = r * (n - i) / (i + 1);
  return r;
}
__device__ inline void walk_k_opt(const StreamCtx& cx, uint32_t desc, const uint32_t* off, uint32_t n, uint32_t k,
                                  uint32_t min_seg, RowSink& out) {
  uint32_t n_pat = 1;  // (k-1)! * 2^(k-1) - 1 reconnection patterns
  for (uint32_t i = 2; i < k; ++i) n_pat *= i;
  n_pat = (n_pat << (k - 1)) - 1;
  uint32_t start = 0, stride = 1;
  if (cx.order != SFGPU_ORDER_ORIGINAL) {
    const uint64_t salt = 0x4B0F7E1171000001ull ^ (uint64_t)desc;
    start = cx.random_index(n, salt);
    stride = cx.random_stride(n, salt ^ 0xA24BAED4963EE407ull);
  }
  for (uint32_t o = 0; o < n; ++o) {
    const uint32_t e = cx.order == SFGPU_ORDER_ORIGINAL ? o : (uint32_t)(((uint64_t)start + (uint64_t)o * stride) % n);
    const uint32_t len = off[e + 1] - off[e];
    if (len < (k + 1) * min_seg) continue;
    const uint32_t choice_count = len - (k + 1) * min_seg + k;
    const uint64_t combos = binomial_dev(choice_count, k);
    const uint64_t move_count = combos * n_pat;
    if (move_count == 0 || move_count > 0xFFFFFFFFull) continue;
    const SelMap mm(cx, (uint32_t)move_count, 0x4B0F7E1171000002ull ^ (uint64_t)desc ^ (uint64_t)e);
    for (uint32_t mo = 0; mo < (uint32_t)move_count; ++mo) {
      const uint32_t sel = mm.at(mo);
      uint64_t rank = sel / n_pat;
      const uint32_t pattern = sel % n_pat;
      uint32_t cuts[5] = {0, 0, 0, 0, 0};
      uint32_t first = 0;
      for (uint32_t position = 0; position < k; ++position) {
        const uint32_t remaining = k - position - 1, maximum = choice_count - (k - position);
        for (uint32_t cand = first; cand <= maximum; ++cand) {
          const uint64_t suffix = binomial_dev(choice_count - cand - 1, remaining);
          if (rank < suffix) {
            cuts[position] = cand + min_seg + position * (min_seg - 1);
            first = cand + 1;
            break;
          }
          rank -= suffix;
        }
      }
      if (!out.push(e | (k << 28), cuts[0] | (cuts[1] << 16), cuts[2] | (cuts[3] << 16), cuts[4] | (pattern << 16))) return;
    }
  }
}

// ChangeMoveSelector cursor (heuristic/selector/move_selector/change.rs:246-307): entities in stream order, per entity
// its values in stream order (canonical contexts keep the value order), then the to-None move when the entity is
// assigned. Rows {entity, to_value (0xFFFFFFFF = None), 0, 0}.
__device__ inline void walk_change(const StreamCtx& cx, uint32_t desc, const int32_t* var, uint32_t n, uint32_t k,
                                   bool with_none, RowSink& out) {
  const uint64_t base = ((uint64_t)desc << 32);  // ^ variable_index (0)
  const SelMap em(cx, n, 0xC4A46E0000000001ull ^ base);
  for (uint32_t eo = 0; eo < n; ++eo) {
    const uint32_t e = em.at(eo);
    const SelMap vm(cx, k, 0xC4A46E0000000000ull ^ (uint64_t)e ^ base);
    for (uint32_t vo = 0; vo < k; ++vo)
      if (!out.push(e, vm.at(vo), 0, 0)) return;
    if (with_none && var[e] >= 0)
      if (!out.push(e, 0xFFFFFFFFu, 0, 0)) return;
  }
}

// SwapMoveSelector cursor over one entity class (move_selector/swap.rs:196-233): the left and the right entity lists
// are permuted independently; every (left, right) with left < right is a SwapMove, left-major. Rows {left, right, 0, 0}.
// The right-hand scan is spread over the warp: most (left, right) pairs of a high left entity fail left < right, so
// the scan is a rejection loop — 32 pairs per trip, matches compacted in pull order.
__device__ inline void walk_swap_warp(const StreamCtx& cx, uint32_t desc, uint32_t n, uint4* rows, uint32_t cap, uint32_t& n_out,
                                      bool& more) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t base = ((uint64_t)desc << 32);
  const SelMap lm(cx, n, 0x5A09000000000001ull ^ base), rm(cx, n, 0x5A09000000000002ull ^ base);
  n_out = 0;
  more = false;
  for (uint32_t lo = 0; lo < n; ++lo) {
    const uint32_t l = lm.at(lo);
    for (uint32_t rb = 0; rb < n; rb += 32) {
      const uint32_t ro = rb + lane;
      const uint32_t r = ro < n ? rm.at(ro) : 0;
      const bool hit = ro < n && l < r;
      const uint32_t hits = __ballot_sync(0xffffffffu, hit);
      if (!hits) continue;
      const uint32_t at = n_out + __popc(hits & ((1u << lane) - 1));
      if (hit && at < cap) rows[at] = make_uint4(l, r, 0, 0);
      n_out += __popc(hits);
      if (n_out > cap) {
        n_out = cap;
        more = true;
        return;
      }
    }
  }
}

// grid = R, 32 threads; lane 0 walks (the cursors are sequential state machines; replicas run side by side)
__global__ void __launch_bounds__(32) union_walk_index_kernel(const __grid_constant__ DevModel m, const UnionArgs a,
                                                              const uint32_t child) {
  const uint32_t r = blockIdx.x;
  if (a.done && a.done[r]) return;
  const char* st = m.state + (size_t)r * m.block_bytes;
  const uint32_t* off = (const uint32_t*)(st + m.off_offsets);
  const StreamCtx cx = stream_ctx(a, r);
  if (a.child[child].family == SFGPU_FAM_SWAP) {  // the whole warp scans
    uint32_t n_out;
    bool more;
    walk_swap_warp(cx, a.sdesc, m.n_entities, (uint4*)a.rows + ((size_t)r * a.n_children + child) * a.window, union_cap(a, r),
                   n_out, more);
    if (threadIdx.x == 0) {
      a.n_emit[(size_t)r * a.n_children + child] = n_out;
      a.ended[(size_t)r * a.n_children + child] = more ? 0 : 1;
    }
    return;
  }
  if (threadIdx.x != 0) return;
  RowSink out;
  out.rows = (uint4*)a.rows + ((size_t)r * a.n_children + child) * a.window;
  out.cap = union_cap(a, r);
  out.n = 0;
  out.more = false;
  const UnionChildDev& c = a.child[child];
  if (c.family == SFGPU_FAM_CHANGE)
    walk_change(cx, a.sdesc, (const int32_t*)(st + m.off_var), m.n_entities, m.n_values, m.allows_unassigned != 0, out);
  else if (c.family == SFGPU_FAM_LIST_REVERSE) walk_reverse(cx, a.desc, off, m.n_owners, out);
  else if (c.family == SFGPU_FAM_K_OPT) walk_k_opt(cx, a.desc, off, m.n_owners, c.p0, c.p1, out);
  else if (c.family == SFGPU_FAM_SUBLIST_CHANGE) walk_sublist_change(cx, a.desc, off, m.n_owners, c.p0, c.p1, out);
  else walk_sublist_swap(cx, a.desc, off, m.n_owners, c.p0, c.p1, out);
  a.n_emit[(size_t)r * a.n_children + child] = out.n;
  a.ended[(size_t)r * a.n_children + child] = out.more ? 0 : 1;
}

// ---- nearby families by rank --------------------------------------------------------------------------------
// Per (replica, step) tables in shared memory: ent[i] = entity at rank i of the (seeded) entity order — with
// SelectionOrder::Random an entity may hold several ranks or none —, slot_first / pos_first = prefix sums over
// ranks of len + 1 / len, and the rank lists of every entity (occ_head / occ_next, increasing rank).
struct RankTables {
  uint32_t* ent;         // [n]
  uint32_t* slot_first;  // [n + 1]
  uint32_t* pos_first;   // [n + 1]
  uint32_t* occ_head;    // [n] by entity
  uint32_t* occ_next;    // [n] by rank
  uint32_t max_occ, n;
};
__device__ __forceinline__ uint32_t rank_tables_words(uint32_t n) { return 5 * n + 2; }

// all threads of the block; ends with a barrier. Global loads (the route lengths) are issued by all threads at once;
// the serial part (prefix sums, rank lists) touches shared memory only.
__device__ inline void build_rank_tables(RankTables& t, uint32_t* smem_words, const StreamCtx& cx, uint64_t entity_salt,
                                         const uint4* rr, uint32_t n, uint32_t* s_max_occ) {
  t.n = n;
  t.ent = smem_words;
  t.slot_first = t.ent + n;
  t.pos_first = t.slot_first + n + 1;
  t.occ_head = t.pos_first + n + 1;
  t.occ_next = t.occ_head + n;
  const SelMap em(cx, n, entity_salt);
  for (uint32_t o = threadIdx.x; o < n; o += blockDim.x) {
    const uint32_t e = n <= 1 ? o : em.at(o);
    t.ent[o] = e;
    t.occ_next[o] = rr[e].y;  // route length of rank o, parked here until the prefix pass
    t.occ_head[o] = UNION_NONE;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t slots = 0, pos = 0;
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t len = t.occ_next[i];
      t.slot_first[i] = slots;
      t.pos_first[i] = pos;
      slots += len + 1;
      pos += len;
    }
    t.slot_first[n] = slots;
    t.pos_first[n] = pos;
    uint32_t mo = 1;
    if (cx.order == SFGPU_ORDER_RANDOM) {  // with replacement: an entity may hold several ranks
      for (uint32_t i = n; i-- > 0;) {     // reverse, so every list is in increasing rank order
        const uint32_t e = t.ent[i];
        t.occ_next[i] = t.occ_head[e];
        t.occ_head[e] = i;
      }
      for (uint32_t e = 0; e < n; ++e) {
        uint32_t k = 0;
        for (uint32_t i = t.occ_head[e]; i != UNION_NONE; i = t.occ_next[i]) ++k;
        mo = max(mo, k);
      }
    } else {  // a permutation: exactly one rank per entity
      for (uint32_t i = 0; i < n; ++i) {
        t.occ_head[t.ent[i]] = i;
        t.occ_next[i] = UNION_NONE;
      }
    }
    *s_max_occ = mo;
  }
  __syncthreads();
  t.max_occ = *s_max_occ;
}

// largest i in [0, n) with first[i] <= v (first is non-decreasing, first[0] = 0)
__device__ __forceinline__ uint32_t rank_of_prefix(const uint32_t* first, uint32_t n, uint32_t v) {
  uint32_t lo = 0, hi = n;  // invariant: first[lo] <= v, answer in [lo, hi)
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (first[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// merges the (<= 64) keys k0 / k1 of this trip into the kept sorted top-32 L
template <typename KEY>
__device__ __forceinline__ void merge_keys(KEY& L, bool& first, KEY k0, KEY k1, uint32_t lane, KEY* buf) {
  const KEY MAXK = KeyTraits<KEY>::maxkey();
  const uint32_t m0 = __ballot_sync(0xffffffffu, k0 != MAXK), m1 = __ballot_sync(0xffffffffu, k1 != MAXK);
  const uint32_t total = __popc(m0) + __popc(m1);
  if (total == 0) return;
  const uint32_t lt = (1u << lane) - 1;
  const uint32_t at = __popc(m0 & lt) + __popc(m1 & lt);
  buf[lane] = MAXK;
  buf[lane + 32] = MAXK;
  __syncwarp();
  if (k0 != MAXK) buf[at] = k0;
  if (k1 != MAXK) buf[at + (k0 != MAXK ? 1 : 0)] = k1;
  __syncwarp();
  KEY b = warp_sort32_adaptive(buf[lane], lane);
  L = first ? b : warp_merge32(L, b, lane);
  first = false;
  if (total > 32) {
    b = warp_sort32(buf[lane + 32], lane);
    L = warp_merge32(L, b, lane);
  }
  __syncwarp();
}

// NearbyListChange: lane l returns the l-th nearest destination slot key of source (rank si, position sp) — the
// scan order of list_kernel/nearby_change.rs:132-195 (own rank first, skipping sp and sp + 1, then the other ranks
// in order; key = distance << scan_bits | scan index, so ties keep scan order like the stable bounded sort).
template <typename KEY, typename CELL>
__device__ __forceinline__ KEY ranked_change_gen(const DevModel& m, const NearbyView& v, const RankTables& t,
                                                 uint32_t scan_bits, uint32_t si, uint32_t sp, uint32_t K,
                                                 uint32_t lane, KEY* __restrict__ buf) {
  const KEY MAXK = KeyTraits<KEY>::maxkey();
  const uint32_t se = t.ent[si];
  const uint4 rsrc = v.rr[se];
  const uint32_t slen = rsrc.y;
  const uint32_t x = v.pr[rsrc.x + sp].x;
  const uint32_t* __restrict__ nb = m.nbr + (size_t)x * m.nbr_stride;
  const CELL* __restrict__ mrow = (const CELL*)m.fm_row + (size_t)x * m.cons[m.fast_pc].n0;
  KEY L = MAXK;
  bool first = true;
  uint32_t d_next = 0;
  // the source element itself is the reference of slots of the OTHER ranks its entity holds (with-replacement entity
  // order): the static neighbour list does not contain x, so those keys are merged first
  if (t.max_occ > 1) {
    const KEY dxx = (KEY)(uint32_t)__ldg(mrow + x);
    for (uint32_t rank = t.occ_head[se]; rank != UNION_NONE; rank = t.occ_next[rank]) {
      if (rank == si) continue;
      KEY k0 = MAXK, k1 = MAXK;
      if (lane == 0) {
        const uint32_t base = t.slot_first[rank] + (rank < si ? slen + 1 : 0);
        k0 = (dxx << scan_bits) | (KEY)(base + sp);
        if (sp + 1 == slen) k1 = (dxx << scan_bits) | (KEY)(base + slen);
      }
      merge_keys<KEY>(L, first, k0, k1, lane, buf);
    }
  }
  for (uint32_t start = 0; start < m.nbr_stride; start += NB_BATCH) {
    if (start > 0) {
      const KEY kth = __shfl_sync(0xffffffffu, L, K - 1);
      if (kth != MAXK && (KEY)d_next > (kth >> scan_bits)) break;  // strictly farther: cannot enter the top K
    }
    const uint32_t idx = start + lane;
    uint32_t y = 0, dyu = 0;
    if (lane <= NB_BATCH && idx < m.nbr_stride) {
      y = __ldg(nb + idx);
      dyu = (uint32_t)__ldg(mrow + y);
    }
    d_next = __shfl_sync(0xffffffffu, dyu, NB_BATCH);
    uint32_t rank = UNION_NONE, py = 0, elen = 0;
    if (lane < NB_BATCH && idx < m.nbr_stride) {
      const uint32_t where = v.pos_of[y];
      if (where != 0xFFFFFFFFu) {
        const uint32_t e = where >> 16;
        py = where & 0xFFFFu;
        elen = v.rr[e].y;
        rank = t.occ_head[e];
      }
    }
    for (uint32_t occ = 0; occ < t.max_occ; ++occ) {
      KEY k0 = MAXK, k1 = MAXK;
      if (rank != UNION_NONE) {
        const bool own = rank == si;
        const uint32_t base = own ? 0 : t.slot_first[rank] + (rank < si ? slen + 1 : 0);
        const KEY dy = (KEY)dyu;
        if (!(own && (py == sp || py == sp + 1))) k0 = (dy << scan_bits) | (KEY)(base + py);
        if (py + 1 == elen && !(own && (elen == sp || elen == sp + 1))) k1 = (dy << scan_bits) | (KEY)(base + elen);
        rank = t.occ_next[rank];
      }
      merge_keys<KEY>(L, first, k0, k1, lane, buf);
    }
  }
  return lane < K ? L : MAXK;
}

// scan index of a change key -> destination (rank, position)
__device__ __forceinline__ void ranked_change_decode(const RankTables& t, uint32_t si, uint32_t slen, uint32_t scan,
                                                     uint32_t& rank, uint32_t& p) {
  if (scan <= slen) {
    rank = si;
    p = scan;
    return;
  }
  const uint32_t u = scan - (slen + 1);
  const uint32_t val = u < t.slot_first[si] ? u : u + slen + 1;
  rank = rank_of_prefix(t.slot_first, t.n, val);
  p = val - t.slot_first[rank];
}

// NearbyListSwap (list_kernel/nearby_swap.rs:99-262): destinations of source (rank si, position sp) are the later
// positions of its own rank, then every position of the later ranks; destination ordinal d IS the scan index.
template <typename KEY, typename CELL>
__device__ __forceinline__ KEY ranked_swap_gen(const DevModel& m, const NearbyView& v, const RankTables& t,
                                               uint32_t scan_bits, uint32_t si, uint32_t sp, uint32_t K,
                                               uint32_t eligible, uint32_t lane, KEY* __restrict__ buf) {
  const KEY MAXK = KeyTraits<KEY>::maxkey();
  const uint32_t se = t.ent[si];
  const uint4 rsrc = v.rr[se];
  const uint32_t slen = rsrc.y, own_later = slen - 1 - sp;
  const uint32_t x = v.pr[rsrc.x + sp].x;
  const CELL* __restrict__ mrow = (const CELL*)m.fm_row + (size_t)x * m.cons[m.fast_pc].n0;
  KEY L = MAXK;
  bool first = true;
  if (eligible <= 64) {
    for (uint32_t base = 0; base < eligible; base += 32) {
      const uint32_t d = base + lane;
      KEY k = MAXK;
      if (d < eligible) {
        uint32_t rank = si, p = sp + 1 + d;
        if (d >= own_later) {
          const uint32_t val = d - own_later + t.pos_first[si + 1];
          rank = rank_of_prefix(t.pos_first, t.n, val);
          // ranks with empty lists share a prefix value: take the last rank with this prefix (the one that owns val)
          p = val - t.pos_first[rank];
        }
        const uint32_t y = v.pr[v.rr[t.ent[rank]].x + p].x;
        k = ((KEY)(uint32_t)__ldg(mrow + y) << scan_bits) | (KEY)d;
      }
      merge_keys<KEY>(L, first, k, MAXK, lane, buf);
    }
    return lane < K ? L : MAXK;
  }
  const uint32_t* __restrict__ nb = m.nbr + (size_t)x * m.nbr_stride;
  uint32_t d_next = 0;
  if (t.max_occ > 1) {  // x itself at the later ranks of its entity (see ranked_change_gen)
    const KEY dxx = (KEY)(uint32_t)__ldg(mrow + x);
    for (uint32_t rank = t.occ_head[se]; rank != UNION_NONE; rank = t.occ_next[rank]) {
      if (rank <= si) continue;
      KEY k = MAXK;
      if (lane == 0) k = (dxx << scan_bits) | (KEY)(own_later + (t.pos_first[rank] - t.pos_first[si + 1]) + sp);
      merge_keys<KEY>(L, first, k, MAXK, lane, buf);
    }
  }
  for (uint32_t start = 0; start < m.nbr_stride; start += 31) {
    if (start > 0) {
      const KEY kth = __shfl_sync(0xffffffffu, L, K - 1);
      if (kth != MAXK && (KEY)d_next > (kth >> scan_bits)) break;
    }
    const uint32_t idx = start + lane;
    uint32_t dyu = 0, rank = UNION_NONE, py = 0;
    if (idx < m.nbr_stride) {
      const uint32_t y = __ldg(nb + idx);
      dyu = (uint32_t)__ldg(mrow + y);
      const uint32_t where = v.pos_of[y];
      if (lane < 31 && where != 0xFFFFFFFFu) {
        py = where & 0xFFFFu;
        rank = t.occ_head[where >> 16];
      }
    }
    d_next = __shfl_sync(0xffffffffu, dyu, 31);
    for (uint32_t occ = 0; occ < t.max_occ; ++occ) {
      KEY k = MAXK;
      if (rank != UNION_NONE) {
        if (rank == si) {
          if (py > sp) k = ((KEY)dyu << scan_bits) | (KEY)(py - sp - 1);
        } else if (rank > si) {
          k = ((KEY)dyu << scan_bits) | (KEY)(own_later + (t.pos_first[rank] - t.pos_first[si + 1]) + py);
        }
        rank = t.occ_next[rank];
      }
      merge_keys<KEY>(L, first, k, MAXK, lane, buf);
    }
  }
  return lane < K ? L : MAXK;
}

// source ordinal q (pull order over sources: ranks in order, positions of a rank in seeded order) -> (rank, po)
__device__ __forceinline__ void source_at(const RankTables& t, uint32_t q, uint32_t& si, uint32_t& po) {
  // first rank whose range [pos_first[i], pos_first[i + 1]) holds q: empty ranks are skipped by taking the LAST
  // rank with pos_first <= q
  si = rank_of_prefix(t.pos_first, t.n, q);
  po = q - t.pos_first[si];
}

// grid = R, 256 threads (8 warps, one source per warp at a time). dynamic smem: rank tables.
template <typename KEY, typename CELL, int MOVE>
__global__ void __launch_bounds__(256) union_walk_nearby_kernel(const __grid_constant__ DevModel m, const UnionArgs a,
                                                                const uint32_t child) {
  extern __shared__ __align__(16) uint32_t u_smem[];
  __shared__ KEY s_buf[8][64];
  __shared__ uint32_t s_max_occ, s_scan[33];
  __shared__ uint32_t s_src[256][3];  // swap: (rank, sp, first row) of the sources of this trip
  const uint32_t r = blockIdx.x;
  if (a.done && a.done[r]) return;
  const char* st = m.state + (size_t)r * m.block_bytes;
  NearbyView v;
  v.rr = (const uint4*)(st + m.off_route_rec);
  v.pr = (const uint4*)(st + m.off_pos_rec);
  v.sr = (const uint4*)(st + m.off_slot_rec);
  v.pos_of = (const uint32_t*)(st + m.off_pos_of);
  const StreamCtx cx = stream_ctx(a, r);
  const uint32_t n = m.n_owners, K = a.child[child].p0, M = union_cap(a, r);
  RankTables t;
  build_rank_tables(t, u_smem, cx, (MOVE == MOVE_SWAP ? 0xA1EA25A090000001ull : 0xA1EA2B17C4A40001ull) ^ (uint64_t)a.desc,
                    v.rr, n, &s_max_occ);
  const uint64_t source_salt = (MOVE == MOVE_SWAP ? 0xA1EA25A090000002ull : 0xA1EA2B17C4A40002ull) ^ (uint64_t)a.desc;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* rows = (uint4*)a.rows + ((size_t)r * a.n_children + child) * a.window;
  const uint32_t n_src = t.pos_first[n];
  const KEY mask = ((KEY)1 << a.scan_bits) - 1;
  if (MOVE == MOVE_CHANGE) {
    // every source yields the same number of candidates: the slots of the non-empty ranks minus its own two
    uint32_t slots = 0;
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t len = t.pos_first[i + 1] - t.pos_first[i];
      slots += len ? len + 1 : 0;
    }
    const uint32_t valid = slots >= 2 ? slots - 2 : 0;
    const uint32_t cnt = valid < K ? valid : K;
    const uint64_t all = (uint64_t)n_src * cnt;
    if (threadIdx.x == 0) {
      a.n_emit[(size_t)r * a.n_children + child] = (uint32_t)(all < M ? all : M);
      a.ended[(size_t)r * a.n_children + child] = all <= M ? 1 : 0;
    }
    if (cnt == 0) return;
    const uint32_t need = (uint32_t)min((uint64_t)n_src, ((uint64_t)M + cnt - 1) / cnt);
    for (uint32_t q = warp; q < need; q += 8) {
      uint32_t si, po;
      source_at(t, q, si, po);
      const uint32_t se = t.ent[si], slen = t.pos_first[si + 1] - t.pos_first[si];
      const uint32_t sp = SelMap(cx, slen, source_salt ^ (uint64_t)se).at(po);
      const KEY key = ranked_change_gen<KEY, CELL>(m, v, t, a.scan_bits, si, sp, K, lane, s_buf[warp]);
      const size_t at = (size_t)q * cnt + lane;
      if (lane < cnt && at < M) {
        uint32_t rank, p;
        ranked_change_decode(t, si, slen, (uint32_t)(key & mask), rank, p);
        rows[at] = key == KeyTraits<KEY>::maxkey() ? make_uint4(0xFFFFFFFFu, 0, 0xFFFFFFFFu, 0) : make_uint4(se, sp, t.ent[rank], p);
      }
    }
  } else {
    // sources yield min(K, later positions) candidates; sources without destinations are skipped
    uint32_t emitted = 0;
    bool full = false;
    for (uint32_t base = 0; base < n_src && !full; base += 256) {
      const uint32_t q = base + threadIdx.x;
      uint32_t cnt = 0, si = 0, sp = 0;
      if (q < n_src) {
        uint32_t po;
        source_at(t, q, si, po);
        const uint32_t se = t.ent[si], slen = t.pos_first[si + 1] - t.pos_first[si];
        sp = SelMap(cx, slen, source_salt ^ (uint64_t)se).at(po);
        const uint32_t eligible = (slen - 1 - sp) + (t.pos_first[n] - t.pos_first[si + 1]);
        cnt = eligible < K ? eligible : K;
      }
      uint32_t tot;
      const uint32_t incl = block_scan_u32(cnt, s_scan, &tot);
      s_src[threadIdx.x][0] = si;
      s_src[threadIdx.x][1] = sp;
      s_src[threadIdx.x][2] = emitted + incl - cnt;
      __syncthreads();
      const uint32_t in_trip = min(256u, n_src - base);
      for (uint32_t j = warp; j < in_trip; j += 8) {
        const uint32_t at0 = s_src[j][2];
        if (at0 >= M) break;  // row offsets grow with j
        const uint32_t sj = s_src[j][0], spj = s_src[j][1];
        const uint32_t se = t.ent[sj], slen = t.pos_first[sj + 1] - t.pos_first[sj];
        const uint32_t own_later = slen - 1 - spj;
        const uint32_t eligible = own_later + (t.pos_first[n] - t.pos_first[sj + 1]);
        const uint32_t c = eligible < K ? eligible : K;
        if (c == 0) continue;
        const KEY key = ranked_swap_gen<KEY, CELL>(m, v, t, a.scan_bits, sj, spj, K, eligible, lane, s_buf[warp]);
        if (lane < c && at0 + lane < M) {
          const uint32_t d = (uint32_t)(key & mask);
          uint32_t rank = sj, p = spj + 1 + d;
          if (d >= own_later) {
            const uint32_t val = d - own_later + t.pos_first[sj + 1];
            rank = rank_of_prefix(t.pos_first, n, val);
            p = val - t.pos_first[rank];
          }
          rows[at0 + lane] = key == KeyTraits<KEY>::maxkey() ? make_uint4(0xFFFFFFFFu, 0, 0xFFFFFFFFu, 0) : make_uint4(se, spj, t.ent[rank], p);
        }
      }
      emitted += tot;
      full = emitted > M;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      a.n_emit[(size_t)r * a.n_children + child] = emitted < M ? emitted : M;
      // `full` left the loop early: candidates remain. Otherwise every source was visited.
      a.ended[(size_t)r * a.n_children + child] = emitted <= M ? 1 : 0;
    }
  }
}

// ---- the scheduler (vec_union.rs:190-366) replayed over the emitted counts -------------------------------------
// one thread per replica. A pull from a child whose WINDOW is used up (cursor not ended) stops the schedule before
// the scheduler state changes: the next pass resumes from scratch with a larger window. Every per-child array is
// indexed by compile-time constants after unrolling (the dynamic child of a pull is found with predicated sweeps),
// so the whole scheduler state lives in registers; StratifiedRandom keeps its arrays in the strided child order
// (position p holds child cid[p]) because that is the order its ties are broken in.
#define UNION_FOR(p) _Pragma("unroll") for (int p = 0; p < UNION_MAX_CHILDREN; ++p)
template <int UORDER>
__global__ void __launch_bounds__(64) union_schedule_kernel(const UnionArgs a, const uint32_t R) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (r == 0) a.offsets[R] = (uint64_t)R * a.t_cap;
  a.offsets[r] = (uint64_t)r * a.t_cap;
  if (a.done && a.done[r]) return;
  const int C = (int)a.n_children;
  const StreamCtx cx = stream_ctx(a, r);
  uint32_t offset = 0, stride = 1;
  if (UORDER == SFGPU_UNION_ROTATING_ROUND_ROBIN || UORDER == SFGPU_UNION_STRATIFIED_RANDOM)
    offset = cx.random_index(C, 0xA11CE5E1EC700001ull);
  if (UORDER == SFGPU_UNION_STRATIFIED_RANDOM) stride = cx.random_stride(C, 0xA11CE5E1EC700002ull);
  uint32_t at[UNION_MAX_CHILDREN], ne[UNION_MAX_CHILDREN], cid[UNION_MAX_CHILDREN];
  bool ended[UNION_MAX_CHILDREN], exhausted[UNION_MAX_CHILDREN];
  int64_t weighted[UNION_MAX_CHILDREN], weight[UNION_MAX_CHILDREN];
  int live = 0;
  int64_t total_live_weight = 0;
  UNION_FOR(p) {
    at[p] = 0;
    weighted[p] = 0;
    const bool in = p < C;
    const uint32_t c = !in ? 0u : (UORDER == SFGPU_UNION_STRATIFIED_RANDOM ? (offset + (uint32_t)p * stride) % (uint32_t)C : (uint32_t)p);
    cid[p] = c;
    ne[p] = in ? a.n_emit[(size_t)r * C + c] : 0u;
    ended[p] = in ? a.ended[(size_t)r * C + c] != 0 : true;
    weight[p] = in ? (int64_t)a.child[c].weight : 0;
    exhausted[p] = weight[p] == 0;
    if (!exhausted[p]) ++live;
    total_live_weight += weight[p];
  }
  int current = UORDER == SFGPU_UNION_STRATIFIED_RANDOM ? 0 : (int)offset;
  uint64_t random_draw = 0;
  uint32_t* sched = a.sched + (size_t)r * a.t_cap;
  uint32_t t = 0;
  bool window_out = false;
  // Fast path. With equal weights RoundRobin, RotatingRoundRobin and StratifiedRandom (smooth weighted round-robin
  // from an all-zero state) pull the live children cyclically — positions `current`, current + 1, .. — and are back in
  // their initial state after every full cycle. So as many full cycles as the shortest child allows are written in
  // one go; the per-pull replay below continues from there (exhaustion, window end, unequal weights, other orders).
  if (UORDER == SFGPU_UNION_ROUND_ROBIN || UORDER == SFGPU_UNION_ROTATING_ROUND_ROBIN || UORDER == SFGPU_UNION_STRATIFIED_RANDOM) {
    bool equal = live == C;
    uint32_t cycles = 0xFFFFFFFFu;
    UNION_FOR(p) if (p < C) {
      equal = equal && weight[p] == 1;
      cycles = min(cycles, ne[p]);
    }
    if (equal && cycles > 0 && cycles != 0xFFFFFFFFu) {
      uint32_t word[UNION_MAX_CHILDREN];  // child of the q-th pull of a cycle
      UNION_FOR(q) {
        word[q] = 0;
        UNION_FOR(p) if (p == (current + q) % C) word[q] = cid[p] << 28;
      }
      for (uint32_t i = 0; i < cycles; ++i) {
        UNION_FOR(q) if (q < C) sched[t + q] = word[q] | i;
        t += (uint32_t)C;
      }
      UNION_FOR(p) if (p < C) at[p] = cycles;
    }
  }
  while (t < a.t_cap && !window_out) {
    int pick = -1;
    // one scheduler decision: `sel` = the position asked for its next candidate
    while (pick < 0 && !window_out) {
      int sel = -1;
      if (UORDER == SFGPU_UNION_SEQUENTIAL) {
        if (current >= C) break;
        sel = current;
      } else if (UORDER == SFGPU_UNION_ROUND_ROBIN || UORDER == SFGPU_UNION_ROTATING_ROUND_ROBIN) {
        if (live == 0) break;
        bool ex = false;
        UNION_FOR(p) if (p == current) ex = exhausted[p];
        if (ex) {
          current = (current + 1) % C;
          continue;
        }
        sel = current;
      } else if (UORDER == SFGPU_UNION_RANDOM) {
        if (live == 0) break;
        const uint64_t draw = cx.mixed(0xA11CE5E1EC701000ull + random_draw) % (uint64_t)total_live_weight;
        uint64_t cumulative = 0;
        UNION_FOR(p) {
          if (!exhausted[p]) {
            cumulative += (uint64_t)weight[p];
            if (sel < 0 && draw < cumulative) sel = p;
          }
        }
      } else {
        if (live == 0) break;
        int64_t sel_w = 0;
        UNION_FOR(p) {
          if (!exhausted[p]) {
            const int64_t w = weighted[p] + weight[p];
            if (sel < 0 || w > sel_w) {
              sel = p;
              sel_w = w;
            }
          }
        }
      }
      bool has = false, end = false;
      UNION_FOR(p) if (p == sel) {
        has = at[p] < ne[p];
        end = ended[p];
      }
      if (!has && !end) {  // the window (not the cursor) ran out: stop before any state changes
        window_out = true;
        break;
      }
      if (UORDER == SFGPU_UNION_ROUND_ROBIN || UORDER == SFGPU_UNION_ROTATING_ROUND_ROBIN) current = (current + 1) % C;
      if (UORDER == SFGPU_UNION_RANDOM) ++random_draw;
      if (UORDER == SFGPU_UNION_STRATIFIED_RANDOM) {
        UNION_FOR(p) {
          if (!exhausted[p]) weighted[p] += weight[p];
          if (p == sel) weighted[p] -= total_live_weight;
        }
      }
      if (has) {
        pick = sel;
      } else if (UORDER == SFGPU_UNION_SEQUENTIAL) {
        ++current;
      } else {
        UNION_FOR(p) if (p == sel) {
          exhausted[p] = true;
          total_live_weight -= weight[p];
        }
        --live;
      }
    }
    if (pick < 0) break;
    uint32_t word = 0;
    UNION_FOR(p) if (p == pick) {
      word = (cid[p] << 28) | at[p];
      ++at[p];
    }
    sched[t++] = word;
  }
  a.n_sched[r] = t;
  // the stream ended iff no window cut happened and no child can deliver any more
  bool any_left = window_out;
  UNION_FOR(p) if (p < C && weight[p] != 0 && (at[p] < ne[p] || !ended[p])) any_left = true;
  a.stream_end[r] = any_left ? 0 : 1;
}

// ---- scoring of the scheduled pulls --------------------------------------------------------------------------
__device__ __forceinline__ bool union_delta(const DevModel& m, const char* st, int family, uint4 row, Score2& d) {
  switch (family) {
    case SFGPU_FAM_NEARBY_LIST_CHANGE: return list_change_delta(m, st, row, d);
    case SFGPU_FAM_NEARBY_LIST_SWAP: return list_swap_delta(m, st, row, d);
    case SFGPU_FAM_LIST_REVERSE: return list_reverse_delta(m, st, row, d);
    case SFGPU_FAM_SUBLIST_CHANGE: return list_sublist_change_delta(m, st, row, d);
    case SFGPU_FAM_K_OPT: return list_k_opt_delta(m, st, row, d);
    case SFGPU_FAM_CHANGE: return score_change_row(m, st, st, make_uint2(row.x, row.y), d);
    case SFGPU_FAM_SWAP: {
      const uint2 r2 = make_uint2(row.x, row.y);
      return score_scalar_candidate<MODE_SWAP>(m, st, st, (const uint32_t*)&r2, nullptr, 0, d);
    }
    default: return list_sublist_swap_delta(m, st, row, d);
  }
}

// Every emitted row of one child scored by child-local index (grid = (chunks, R), 128 threads; one kernel per move
// family, so a warp runs one delta function). The replica block is read in place: a window holds a few dozen rows per
// child, far too few to amortise staging the block.
template <int FAMILY>
__global__ void __launch_bounds__(128) union_score_child_kernel(const __grid_constant__ DevModel m, const UnionArgs a,
                                                                const uint32_t child) {
  const uint32_t r = blockIdx.y;
  if (a.done && a.done[r]) return;
  const uint32_t n = a.n_emit[(size_t)r * a.n_children + child];
  if (blockIdx.x * blockDim.x >= n) return;
  const char* st = m.state + (size_t)r * m.block_bytes;
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  const size_t base = ((size_t)r * a.n_children + child) * a.window;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const uint4 row = ((const uint4*)a.rows)[base + j];
    Score2 d;
    const bool ok = union_delta(m, st, FAMILY, row, d);
    longlong2 o;
    o.x = ok ? ch + d.hard : 0;
    o.y = ok ? csf + d.soft : 0;
    ((longlong2*)a.child_scores)[base + j] = o;
    a.child_doable[base + j] = ok ? 1 : 0;
  }
}

// scores into union pull order: pull t of replica r = (child, child-local index) of the schedule
__global__ void __launch_bounds__(256) union_gather_kernel(const UnionArgs a) {
  const uint32_t r = blockIdx.y;
  if (a.done && a.done[r]) return;
  const uint32_t T = a.n_sched[r];
  const uint32_t* sched = a.sched + (size_t)r * a.t_cap;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const uint32_t s = sched[t];
    const size_t src = ((size_t)r * a.n_children + (s >> 28)) * a.window + (s & 0x0FFFFFFFu);
    const size_t q = (size_t)r * a.t_cap + t;
    ((longlong2*)a.scores)[q] = ((const longlong2*)a.child_scores)[src];
    a.doable[q] = a.child_doable[src];
  }
}

// one thread per replica, after argbest: completeness of the step, winner row, per-replica apply kind
__global__ void union_pick_kernel(const UnionArgs a, const uint32_t R, const uint32_t last_pass, uint32_t* out_index,
                                  const uint32_t* out_evaluated, uint32_t* out_winner_rows /* [R][8] or null */,
                                  uint32_t* apply_rows /* [R][4] */, int32_t* apply_kinds /* [R] */,
                                  uint32_t* out_flags /* [R] or null */, uint32_t* pending /* [1] */,
                                  uint64_t* overflow_acc /* [2][R] or null: overflows, pulls scored */) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (a.done[r]) return;
  const uint32_t T = a.n_sched[r];
  if (overflow_acc) overflow_acc[R + r] += T;
  const bool quit = a.f.accepted_limit > 0 && out_evaluated[r] < T;
  const bool complete = a.stream_end[r] != 0 || quit;
  if (!complete && !last_pass) {
    atomicAdd(pending, 1u);
    return;
  }
  a.done[r] = 1;
  if (a.next_win) {
    // next step's first window: what this step needed per child plus half, at least 16
    const uint32_t per = (out_evaluated[r] + a.n_children - 1) / a.n_children;
    a.next_win[r] = min(a.win_max, max(16u, per + per / 2 + 8));
  }
  if (out_flags) out_flags[r] = complete ? 0u : 1u;  // bit 0: the window limit cut the step short
  if (overflow_acc && !complete) overflow_acc[r] += 1;
  const uint32_t idx = out_index[r];
  uint4 row = make_uint4(0xFFFFFFFFu, 0, 0xFFFFFFFFu, 0);
  int32_t kind = -1;
  uint32_t c = 0, j = 0, family = 0xFFFFFFFFu;
  if (idx != 0xFFFFFFFFu) {
    const uint32_t s = a.sched[(size_t)r * a.t_cap + idx];
    c = s >> 28;
    j = s & 0x0FFFFFFFu;
    row = ((const uint4*)a.rows)[((size_t)r * a.n_children + c) * a.window + j];
    family = (uint32_t)a.child[c].family;
    kind = family == SFGPU_FAM_CHANGE ? 0
           : family == SFGPU_FAM_SWAP   ? 1
           : family == SFGPU_FAM_NEARBY_LIST_CHANGE ? 2
           : family == SFGPU_FAM_NEARBY_LIST_SWAP ? 3
           : family == SFGPU_FAM_LIST_REVERSE     ? 4
           : family == SFGPU_FAM_SUBLIST_CHANGE   ? 5
           : family == SFGPU_FAM_K_OPT            ? 7
                                                  : 6;
  }
  ((uint4*)apply_rows)[r] = row;
  apply_kinds[r] = kind;
  if (out_winner_rows) {
    uint32_t* w = out_winner_rows + (size_t)r * 8;
    w[0] = family; w[1] = c; w[2] = row.x; w[3] = row.y; w[4] = row.z; w[5] = row.w; w[6] = j; w[7] = 0;
  }
}

__global__ void union_reset_kernel(uint32_t* done, uint32_t* pending, uint32_t R) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) done[r] = 0;
  if (r == 0) *pending = 0;
}

__global__ void union_fill_kernel(uint32_t* dst, uint32_t value, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = value;
}

// sets the condition of a later window pass (an IF node of the step graph): taken when replicas are still pending
__global__ void union_cond_kernel(cudaGraphConditionalHandle handle, const uint32_t* pending) {
  cudaGraphSetConditional(handle, *pending ? 1u : 0u);
}
