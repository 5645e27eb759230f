// Device-side enumeration of a list neighbourhood by PULL INDEX, fused with scoring and the forager replay.
//
// The sublist neighbourhoods are too large to materialise (CVRP-1000 / 80 routes, sizes 1..=3: ~3.2 M
// SublistChange candidates per replica and step = 100 MB of rows + scores per replica), so nothing is written to
// HBM: a thread decodes candidate `idx` of the canonical selector order straight from the staged route lengths,
// scores it against the staged replica block and keeps a forager partial; only chunk partials and the winner
// leave the SM. The finish kernel replays the AcceptedCount cut and the tie rule by re-decoding one chunk in order.
//
// SublistChangeNb: SublistChangeMoveSelector, SelectionOrder::Original
//   (heuristic/selector/sublist_change.rs:166-205 + list_kernel/sublist_change.rs:103-268): entities in order; per
//   source every segment start, every valid size min..=max; intra-list destinations 0..=len-size of the
//   post-removal list except the segment's own start, then every position 0..=len of every other entity in entity
//   order. A segment of `size` therefore owns T - size candidates, T = total elements + entities - 1.
#pragma once
#include "sfgpu_kernels.cuh"


struct SublistChangeNb {
  const uint32_t* off;   // staged route offsets
  uint32_t* ent_base;    // [n + 1] first pull index of every source entity
  uint32_t* dst_pre;     // [n + 1] prefix of (len + 1): destination slots before every entity
  uint32_t n, T, min_size, max_size, n_sizes, size_sum;

  static __host__ __device__ __forceinline__ size_t table_words(uint32_t n_owners, uint32_t) {
    return (size_t)2 * (n_owners + 1);
  }
  // candidates of one source route
  __device__ __forceinline__ uint32_t per_start(uint32_t mv) const {  // sizes min..=mv of one start
    if (mv < min_size) return 0;
    const uint32_t c = mv - min_size + 1;
    return c * T - (min_size + mv) * c / 2;
  }
  __device__ __forceinline__ uint32_t route_count(uint32_t slen) const {
    if (slen < min_size) return 0;
    const uint32_t nf = slen >= max_size ? slen - max_size + 1 : 0;  // starts that fit every size
    uint32_t tot = nf * per_start(max_size);
    for (uint32_t start = nf; start < slen; ++start) tot += per_start(slen - start);
    return tot;
  }
  // tables in shared memory; all threads of the CTA must call. scratch: 2 * (n + 1) uint32
  __device__ __forceinline__ void build(const DevModel& m, const char* st, uint32_t* scratch, uint32_t mn, uint32_t mx) {
    off = (const uint32_t*)(st + m.off_offsets);
    n = m.n_owners;
    ent_base = scratch;
    dst_pre = scratch + n + 1;
    min_size = mn;
    max_size = mx;
    n_sizes = mx - mn + 1;
    size_sum = (mn + mx) * n_sizes / 2;
    T = off[n] + n - 1;
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
      const uint32_t slen = off[e + 1] - off[e];
      ent_base[e + 1] = route_count(slen);
      dst_pre[e + 1] = slen + 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      ent_base[0] = 0;
      dst_pre[0] = 0;
      for (uint32_t e = 0; e < n; ++e) {
        ent_base[e + 1] += ent_base[e];
        dst_pre[e + 1] += dst_pre[e];
      }
    }
    __syncthreads();
  }
  __device__ __forceinline__ uint32_t total() const { return ent_base[n]; }
  // last e with tab[e] <= v (tab non-decreasing, tab[0] = 0, v < tab[n])
  __device__ __forceinline__ uint32_t find(const uint32_t* tab, uint32_t v) const {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (tab[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
  }
  // candidate `idx` (< total()) as the packed device row {src_e, start | size << 24, dst_e, dst_position}
  __device__ __forceinline__ uint4 decode(uint32_t idx) const {
    const uint32_t e = find(ent_base, idx);
    uint32_t r = idx - ent_base[e];
    const uint32_t slen = off[e + 1] - off[e];
    const uint32_t nf = slen >= max_size ? slen - max_size + 1 : 0;
    const uint32_t k_full = n_sizes * T - size_sum;
    uint32_t start, mv;
    if (r < nf * k_full) {
      start = r / k_full;
      r -= start * k_full;
      mv = max_size;
    } else {
      r -= nf * k_full;
      start = nf;
      for (;;) {
        mv = slen - start;
        const uint32_t k = per_start(mv);
        if (r < k) break;
        r -= k;
        ++start;
      }
    }
    uint32_t size = min_size;
    while (size < mv && r >= T - size) {
      r -= T - size;
      ++size;
    }
    const uint32_t post = slen - size;  // intra-list destinations (own start skipped)
    if (r < post) return make_uint4(e, start | (size << 24), e, r < start ? r : r + 1);
    uint32_t k = r - post;                     // slot among the other entities' (len + 1) positions
    if (k >= dst_pre[e]) k += slen + 1;        // jump over the source's own slots
    const uint32_t de = find(dst_pre, k);
    return make_uint4(e, start | (size << 24), de, k - dst_pre[de]);
  }
  __device__ __forceinline__ bool delta(const DevModel& m, const char* st, uint4 row, Score2& d) const {
    return list_sublist_change_delta(m, st, row, d);
  }

  // ---- cursor: a thread walks a run of consecutive pull indices. Consecutive candidates differ only in the
  // destination until the segment changes (~every T candidates), so the decode is incremental and, for
  // CVRP-shaped programs (m.fast_list: at most one linear path cost and one per-route sum), everything that
  // depends on the source segment only is computed once per segment. Same integer arithmetic as the generic
  // delta, regrouped — the finish kernel re-scores single candidates with the generic one.
  static constexpr bool kHasCursor = true;
  struct Cursor {
    uint32_t e, start, size, slen, sb;   // source segment
    uint32_t de, dp, dlen, db;           // destination
    uint32_t first, last;
    int64_t gap, inner, v;               // path: closing the gap, inner legs; sum: segment total
    int64_t os, ss, w_ss, ls_src_after;  // retained per-route aggregates of the source and their weights
  };
  __device__ __forceinline__ void set_dest(Cursor& c, uint32_t de, uint32_t dp) const {
    c.de = de;
    c.dp = dp;
    c.db = off[de];
    c.dlen = off[de + 1] - c.db;
  }
  __device__ __forceinline__ void seek(const DevModel& m, const char* st, uint32_t idx, Cursor& c) const {
    const uint4 row = decode(idx);
    c.e = row.x;
    c.start = SFGPU_SEG_POS(row.y);
    c.size = SFGPU_SEG_SIZE(row.y);
    c.sb = off[c.e];
    c.slen = off[c.e + 1] - c.sb;
    set_dest(c, row.z, row.w);
    if (!m.fast_list) return;
    const uint32_t* el = (const uint32_t*)(st + m.off_elems);
    const uint32_t end = c.start + c.size;
    c.first = el[c.sb + c.start];
    c.last = el[c.sb + end - 1];
    c.gap = c.inner = c.v = c.os = c.ss = c.w_ss = c.ls_src_after = 0;
    if (m.fast_pc >= 0) {
      const ConsDev& pc = m.cons[m.fast_pc];
      const uint32_t depot = (uint32_t)pc.p0;
      const uint32_t prev = c.start > 0 ? el[c.sb + c.start - 1] : depot;
      const uint32_t next = end < c.slen ? el[c.sb + end] : depot;
      c.gap = -mat_at(pc, prev, c.first) - mat_at(pc, c.last, next) + (c.slen > c.size ? mat_at(pc, prev, next) : 0);
      for (uint32_t i = c.start; i + 1 < end; ++i) c.inner += mat_at(pc, el[c.sb + i], el[c.sb + i + 1]);
      c.os = ((const int64_t*)(st + pc.off0))[c.e];
    }
    if (m.fast_ls >= 0) {
      const ConsDev& ls = m.cons[m.fast_ls];
      for (uint32_t i = c.start; i < end; ++i) c.v += ((const int64_t*)ls.g0)[el[c.sb + i]];
      c.ss = ((const int64_t*)(st + ls.off0))[c.e];
      c.w_ss = weight_eval(ls.w, c.ss);
      c.ls_src_after = weight_eval(ls.w, c.ss - c.v);
    }
  }
  // cursor -> candidate next_idx = current + 1
  __device__ __forceinline__ void advance(const DevModel& m, const char* st, uint32_t next_idx, Cursor& c) const {
    if (c.de == c.e) {  // intra-list destinations 0..=post except the segment's own start
      uint32_t dp = c.dp + 1;
      if (dp == c.start) ++dp;
      if (dp <= c.slen - c.size) {
        c.dp = dp;
        return;
      }
      const uint32_t de = c.e == 0 ? 1 : 0;
      if (de >= n) return seek(m, st, next_idx, c);
      return set_dest(c, de, 0);
    }
    if (c.dp < c.dlen) {
      ++c.dp;
      return;
    }
    uint32_t de = c.de + 1;
    if (de == c.e) ++de;
    if (de >= n) return seek(m, st, next_idx, c);
    set_dest(c, de, 0);
  }
  __device__ __forceinline__ bool eval(const DevModel& m, const char* st, const Cursor& c, Score2& d) const {
    if (!m.fast_list)
      return list_sublist_change_delta(m, st, make_uint4(c.e, c.start | (c.size << 24), c.de, c.dp), d);
    d.hard = 0;
    d.soft = 0;
    const uint32_t* el = (const uint32_t*)(st + m.off_elems);
    const bool intra = c.de == c.e;
    if (m.fast_pc >= 0) {  // linear weight (fast_list): w(x + delta) - w(x) = a * delta
      const ConsDev& pc = m.cons[m.fast_pc];
      const uint32_t depot = (uint32_t)pc.p0;
      uint32_t a, b;
      int64_t delta;
      if (intra) {
        const uint32_t ia = c.dp > 0 ? c.dp - 1 : 0, ib = c.dp;
        a = c.dp > 0 ? el[c.sb + (ia < c.start ? ia : ia + c.size)] : depot;
        b = c.dp < c.slen - c.size ? el[c.sb + (ib < c.start ? ib : ib + c.size)] : depot;
        delta = c.gap + mat_at(pc, a, c.first) + mat_at(pc, c.last, b) - mat_at(pc, a, b);
      } else {
        a = c.dp > 0 ? el[c.db + c.dp - 1] : depot;
        b = c.dp < c.dlen ? el[c.db + c.dp] : depot;
        // source: gap - inner; destination: legs + inner; the inner legs cancel under a linear weight
        delta = c.gap + mat_at(pc, a, c.first) + mat_at(pc, c.last, b) - (c.dlen > 0 ? mat_at(pc, a, b) : 0);
      }
      add_level(d, pc, pc.w.a * delta);
    }
    if (m.fast_ls >= 0 && !intra) {
      const ConsDev& ls = m.cons[m.fast_ls];
      const int64_t ds = ((const int64_t*)(st + ls.off0))[c.de];
      add_level(d, ls, c.ls_src_after - c.w_ss + weight_eval(ls.w, ds + c.v) - weight_eval(ls.w, ds));
    }
    return true;
  }
};

// SublistSwapNb: SublistSwapMoveSelector, SelectionOrder::Original (heuristic/selector/sublist_swap.rs +
//   list_kernel/sublist_swap.rs:28-318): first segments in entity / start / size order; for each, the second
//   segments are those of the same list that start at or after the first one's end, then every segment of every
//   later entity. With G(m) = number of segments inside a list suffix of length m, a first segment (e, s, size)
//   owns G(len - s - size) + (segments of the entities after e) candidates. Tables: pull index of the first
//   candidate per flat element position, and the segment-count prefix per entity.
struct SublistSwapNb {
  const uint32_t* off;
  uint32_t* pos_base;  // [elems + 1]
  uint32_t* seg_pre;   // [n + 1] segments of the entities before e
  uint32_t n, elems, min_size, max_size, n_sizes;

  static __host__ __device__ __forceinline__ size_t table_words(uint32_t n_owners, uint32_t elem_cap) {
    return (size_t)elem_cap + 1 + n_owners + 1 + 40;
  }
  __device__ __forceinline__ uint32_t sizes_at(uint32_t m) const {  // valid sizes with m elements left
    const uint32_t mv = m < max_size ? m : max_size;
    return mv >= min_size ? mv - min_size + 1 : 0;
  }
  __device__ __forceinline__ uint32_t G(uint32_t m) const {  // segments inside a suffix of length m
    if (m < min_size) return 0;
    const uint32_t mm = m < max_size ? m : max_size;
    const uint32_t t = mm - min_size + 1;
    return t * (t + 1) / 2 + (m > max_size ? (m - max_size) * n_sizes : 0);
  }
  __device__ __forceinline__ uint32_t entity_of(uint32_t p) const {  // last e with off[e] <= p and a non-empty route
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (off[mid] <= p) lo = mid; else hi = mid;
    }
    return lo;
  }
  __device__ __forceinline__ void build(const DevModel& m, const char* st, uint32_t* scratch, uint32_t mn, uint32_t mx) {
    off = (const uint32_t*)(st + m.off_offsets);
    n = m.n_owners;
    elems = off[n];
    min_size = mn;
    max_size = mx;
    n_sizes = mx - mn + 1;
    pos_base = scratch;
    seg_pre = scratch + elems + 1;
    uint32_t* scan = seg_pre + n + 1;  // 33 words for the block scan
    if (threadIdx.x == 0) {
      seg_pre[0] = 0;
      for (uint32_t e = 0; e < n; ++e) seg_pre[e + 1] = seg_pre[e] + G(off[e + 1] - off[e]);
    }
    __syncthreads();
    // candidates owned by the first segments that start at flat position p, then an exclusive prefix over p
    uint32_t carry = 0;
    for (uint32_t base = 0; base < elems; base += blockDim.x) {
      const uint32_t p = base + threadIdx.x;
      uint32_t cnt = 0;
      if (p < elems) {
        const uint32_t e = entity_of(p), L = off[e + 1] - off[e], s = p - off[e];
        const uint32_t later = seg_pre[n] - seg_pre[e + 1];
        const uint32_t ns = sizes_at(L - s);
        for (uint32_t k = 0; k < ns; ++k) cnt += G(L - s - (min_size + k)) + later;
      }
      uint32_t tot;
      const uint32_t incl = block_scan_u32(cnt, scan, &tot);
      if (p < elems) pos_base[p] = carry + incl - cnt;
      carry += tot;
      __syncthreads();
    }
    if (threadIdx.x == 0) pos_base[elems] = carry;
    __syncthreads();
  }
  __device__ __forceinline__ uint32_t total() const { return pos_base[elems]; }
  // the r-th segment (start asc, size asc) among those of a list of length L that start at or after `from`
  __device__ __forceinline__ uint32_t segment_at(uint32_t L, uint32_t from, uint32_t r) const {
    const uint32_t m = L - from;
    const uint32_t nf = m >= max_size ? m - max_size + 1 : 0;
    uint32_t start, size;
    if (r < nf * n_sizes) {
      start = from + r / n_sizes;
      size = min_size + r % n_sizes;
    } else {
      r -= nf * n_sizes;
      start = from + nf;
      for (;;) {
        const uint32_t c = sizes_at(L - start);
        if (r < c) break;
        r -= c;
        ++start;
      }
      size = min_size + r;
    }
    return start | (size << 24);
  }
  __device__ __forceinline__ uint4 decode(uint32_t idx) const {
    uint32_t lo = 0, hi = elems;  // last p with pos_base[p] <= idx (pos_base[elems] = total > idx)
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (pos_base[mid] <= idx) lo = mid; else hi = mid;
    }
    const uint32_t p = lo;
    uint32_t r = idx - pos_base[p];
    const uint32_t e = entity_of(p), L = off[e + 1] - off[e], s1 = p - off[e];
    const uint32_t later = seg_pre[n] - seg_pre[e + 1];
    uint32_t size1 = min_size;
    for (;;) {
      const uint32_t cnt = G(L - s1 - size1) + later;
      if (r < cnt) break;
      r -= cnt;
      ++size1;
    }
    const uint32_t own = G(L - s1 - size1);
    if (r < own) return make_uint4(e, s1 | (size1 << 24), e, segment_at(L, s1 + size1, r));
    const uint32_t q = seg_pre[e + 1] + (r - own);  // rank among the segments of all entities
    uint32_t a = e + 1, b = n;                     // last si with seg_pre[si] <= q
    while (b - a > 1) {
      const uint32_t mid = (a + b) >> 1;
      if (seg_pre[mid] <= q) a = mid; else b = mid;
    }
    return make_uint4(e, s1 | (size1 << 24), a, segment_at(off[a + 1] - off[a], 0, q - seg_pre[a]));
  }
  __device__ __forceinline__ bool delta(const DevModel& m, const char* st, uint4 row, Score2& d) const {
    return list_sublist_swap_delta(m, st, row, d);
  }

  // ---- cursor (see SublistChangeNb): consecutive pull indices keep the first segment and step the second one
  // through size, start and entity; everything that depends on the first segment only is hoisted.
  static constexpr bool kHasCursor = true;
  struct Cursor {
    uint32_t e1, s1, n1, b1, L1;
    uint32_t e2, s2, n2, b2, L2;
    uint32_t p1, x1, f1, l1;         // first segment: element before / after, first / last element
    int64_t c_pf, c_lx;              // legs (p1, f1) and (l1, x1)
    int64_t v1, a1, w_a1;            // per-route sum constraint: segment total, route sum, its weight
  };
  __device__ __forceinline__ void set_second(Cursor& c, uint32_t e2, uint32_t s2, uint32_t n2) const {
    c.e2 = e2;
    c.s2 = s2;
    c.n2 = n2;
    c.b2 = off[e2];
    c.L2 = off[e2 + 1] - c.b2;
  }
  __device__ __forceinline__ void seek(const DevModel& m, const char* st, uint32_t idx, Cursor& c) const {
    const uint4 row = decode(idx);
    c.e1 = row.x;
    c.s1 = SFGPU_SEG_POS(row.y);
    c.n1 = SFGPU_SEG_SIZE(row.y);
    c.b1 = off[c.e1];
    c.L1 = off[c.e1 + 1] - c.b1;
    set_second(c, row.z, SFGPU_SEG_POS(row.w), SFGPU_SEG_SIZE(row.w));
    if (!m.fast_list) return;
    const uint32_t* el = (const uint32_t*)(st + m.off_elems);
    const uint32_t t1 = c.s1 + c.n1;
    c.f1 = el[c.b1 + c.s1];
    c.l1 = el[c.b1 + t1 - 1];
    c.p1 = c.x1 = 0;
    c.c_pf = c.c_lx = c.v1 = c.a1 = c.w_a1 = 0;
    if (m.fast_pc >= 0) {
      const ConsDev& pc = m.cons[m.fast_pc];
      const uint32_t depot = (uint32_t)pc.p0;
      c.p1 = c.s1 > 0 ? el[c.b1 + c.s1 - 1] : depot;
      c.x1 = t1 < c.L1 ? el[c.b1 + t1] : depot;
      c.c_pf = mat_at(pc, c.p1, c.f1);
      c.c_lx = mat_at(pc, c.l1, c.x1);
    }
    if (m.fast_ls >= 0) {
      const ConsDev& ls = m.cons[m.fast_ls];
      for (uint32_t i = c.s1; i < t1; ++i) c.v1 += ((const int64_t*)ls.g0)[el[c.b1 + i]];
      c.a1 = ((const int64_t*)(st + ls.off0))[c.e1];
      c.w_a1 = weight_eval(ls.w, c.a1);
    }
  }
  __device__ __forceinline__ void advance(const DevModel& m, const char* st, uint32_t next_idx, Cursor& c) const {
    if (c.n2 < max_size && c.s2 + c.n2 + 1 <= c.L2) {  // next size at the same start
      ++c.n2;
      return;
    }
    if (c.s2 + 1 + min_size <= c.L2) {  // next start of the same list
      ++c.s2;
      c.n2 = min_size;
      return;
    }
    uint32_t e = c.e2 + 1;  // next entity that holds a segment
    while (e < n && off[e + 1] - off[e] < min_size) ++e;
    if (e >= n) return seek(m, st, next_idx, c);
    set_second(c, e, 0, min_size);
  }
  __device__ __forceinline__ bool eval(const DevModel& m, const char* st, const Cursor& c, Score2& d) const {
    if (!m.fast_list)
      return list_sublist_swap_delta(m, st, make_uint4(c.e1, c.s1 | (c.n1 << 24), c.e2, c.s2 | (c.n2 << 24)), d);
    d.hard = 0;
    d.soft = 0;
    const uint32_t* el = (const uint32_t*)(st + m.off_elems);
    const bool intra = c.e1 == c.e2;
    const uint32_t t1 = c.s1 + c.n1, t2 = c.s2 + c.n2;
    if (m.fast_pc >= 0) {  // linear weight: the inner legs of both segments cancel, w(x + delta) - w(x) = a * delta
      const ConsDev& pc = m.cons[m.fast_pc];
      const uint32_t depot = (uint32_t)pc.p0;
      const uint32_t f2 = el[c.b2 + c.s2], l2 = el[c.b2 + t2 - 1];
      const uint32_t x2 = t2 < c.L2 ? el[c.b2 + t2] : depot;
      int64_t delta;
      if (intra) {  // the first segment is the earlier one (second.start >= first.end)
        delta = mat_at(pc, c.p1, f2) + mat_at(pc, c.l1, x2) - c.c_pf - mat_at(pc, l2, x2);
        if (t1 == c.s2) {
          delta += mat_at(pc, l2, c.f1) - mat_at(pc, c.l1, f2);
        } else {
          const uint32_t mk = el[c.b1 + c.s2 - 1];
          delta += mat_at(pc, l2, c.x1) + mat_at(pc, mk, c.f1) - c.c_lx - mat_at(pc, mk, f2);
        }
      } else {
        const uint32_t p2 = c.s2 > 0 ? el[c.b2 + c.s2 - 1] : depot;
        delta = mat_at(pc, c.p1, f2) + mat_at(pc, l2, c.x1) - c.c_pf - c.c_lx + mat_at(pc, p2, c.f1) +
                mat_at(pc, c.l1, x2) - mat_at(pc, p2, f2) - mat_at(pc, l2, x2);
      }
      add_level(d, pc, pc.w.a * delta);
    }
    if (m.fast_ls >= 0 && !intra) {
      const ConsDev& ls = m.cons[m.fast_ls];
      // the per-position records of the fast list path carry the element's column value (staged with the block)
      const PosRec* pr = (const PosRec*)(st + m.off_pos_rec);
      int64_t v2 = 0;
      for (uint32_t i = c.s2; i < t2; ++i) v2 += pr[c.b2 + i].val;
      const int64_t a2 = ((const int64_t*)(st + ls.off0))[c.e2];
      add_level(d, ls, weight_eval(ls.w, c.a1 - c.v1 + v2) - c.w_a1 + weight_eval(ls.w, a2 - v2 + c.v1) -
                           weight_eval(ls.w, a2));
    }
    return true;
  }
};

// ReverseNb: ListReverseMoveSelector, SelectionOrder::Original (heuristic/selector/list_reverse.rs:139-172 +
//   list_kernel/reverse.rs:66-108): per entity with len >= 2 every start and every end in start + 2 ..= len, so a
//   list of length L owns L (L - 1) / 2 candidates and start s owns L - 1 - s of them. Rows {entity, start, end, 0}.
struct ReverseNb {
  const uint32_t* off;
  uint32_t* ent_base;  // [n + 1]
  uint32_t n;
  static constexpr bool kHasCursor = false;
  static __host__ __device__ __forceinline__ size_t table_words(uint32_t n_owners, uint32_t) { return (size_t)n_owners + 1; }
  __device__ __forceinline__ void build(const DevModel& m, const char* st, uint32_t* scratch, uint32_t, uint32_t) {
    off = (const uint32_t*)(st + m.off_offsets);
    n = m.n_owners;
    ent_base = scratch;
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
      const uint32_t L = off[e + 1] - off[e];
      ent_base[e + 1] = L >= 2 ? L * (L - 1) / 2 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      ent_base[0] = 0;
      for (uint32_t e = 0; e < n; ++e) ent_base[e + 1] += ent_base[e];
    }
    __syncthreads();
  }
  __device__ __forceinline__ uint32_t total() const { return ent_base[n]; }
  __device__ __forceinline__ uint4 decode(uint32_t idx) const {
    uint32_t lo = 0, hi = n;  // last e with ent_base[e] <= idx
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (ent_base[mid] <= idx) lo = mid; else hi = mid;
    }
    const uint32_t e = lo, L = off[e + 1] - off[e];
    // start s owns L - 1 - s candidates, so B(s) = s (2L - 1 - s) / 2 precede it: invert the triangular number
    // in closed form (double sqrt, exact for L < 2^24) and fix the estimate up by one step either way
    const uint32_t r = idx - ent_base[e];
    const double t = 2.0 * (double)L - 1.0;
    uint32_t s = (uint32_t)((t - sqrt(t * t - 8.0 * (double)r)) * 0.5);
    if (s > L - 2) s = L - 2;
    auto before = [&](uint32_t q) { return (uint32_t)(((uint64_t)q * (2ull * L - 1 - q)) / 2); };
    while (s > 0 && before(s) > r) --s;
    while (s + 1 <= L - 2 && before(s + 1) <= r) ++s;
    return make_uint4(e, s, s + 2 + (r - before(s)), 0);
  }
  __device__ __forceinline__ bool delta(const DevModel& m, const char* st, uint4 row, Score2& d) const {
    return list_reverse_delta(m, st, row, d);
  }
};

// grid = (chunks, R): chunk c scores pull indices [c * per_chunk, (c + 1) * per_chunk) of its replica.
// Dynamic shared memory: [staged block (STAGED)] [NB::table_words(..) uint32].
template <bool STAGED, class NB>
__global__ void __launch_bounds__(256, 3) index_step_kernel(const __grid_constant__ DevModel m, const IndexStepArgs a) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  __shared__ int64_t sh_h[8], sh_s[8];
  __shared__ uint32_t sh_n[8], sh_f[8], sh_a[8];
  const uint32_t r = blockIdx.y;
  const char* gblock = m.state + (size_t)r * m.block_bytes;
  const char* st = gblock;
  uint32_t* tables = (uint32_t*)smem;
  if (STAGED) {
    stage_block(smem, gblock, m.stage_bytes, &bar);
    st = smem;
    tables = (uint32_t*)(smem + m.stage_bytes);
  }
  NB nb;
  nb.build(m, st, tables, a.min_size, a.max_size);
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  int64_t lh = 0, ls = 0, th = 0, ts = 0;
  if (a.ref_scores) {
    lh = a.ref_scores[r * 4 + 0];
    ls = a.ref_scores[r * 4 + 1];
    th = a.ref_scores[r * 4 + 2];
    ts = a.ref_scores[r * 4 + 3];
  }
  const uint32_t total = nb.total();
  const uint64_t c_lo64 = (uint64_t)blockIdx.x * a.per_chunk;
  const uint32_t c_lo = c_lo64 < total ? (uint32_t)c_lo64 : total;
  const uint32_t c_hi = (uint64_t)c_lo + a.per_chunk < total ? c_lo + a.per_chunk : total;
  int64_t tb_h = 0, tb_s = 0;
  uint32_t tb_n = 0, tb_first = 0xFFFFFFFFu, t_acc = 0;
  auto consider = [&](uint32_t idx, const Score2& d) {
    const int64_t oh = ch + d.hard, os = csf + d.soft;
    if (!accept_score(a.f.acceptor, oh, os, lh, ls, th, ts)) return;
    t_acc++;
    if (tb_n == 0 || score_less(tb_h, tb_s, oh, os)) {
      tb_h = oh;
      tb_s = os;
      tb_n = 1;
      tb_first = idx;
    } else if (tb_h == oh && tb_s == os) {
      tb_n++;
    }
  };
  if constexpr (NB::kHasCursor) {
    // each thread walks one run of consecutive pull indices
    const uint32_t run = a.per_chunk / blockDim.x;
    uint32_t idx = c_lo + threadIdx.x * run;
    const uint32_t end = idx < c_hi ? (idx + run < c_hi ? idx + run : c_hi) : idx;
    if (idx < end) {
      typename NB::Cursor cur;
      nb.seek(m, st, idx, cur);
      for (;;) {
        Score2 d;
        if (nb.eval(m, st, cur, d)) consider(idx, d);
        if (++idx >= end) break;
        nb.advance(m, st, idx, cur);
      }
    }
  } else {
    for (uint32_t idx = c_lo + threadIdx.x; idx < c_hi; idx += blockDim.x) {
      Score2 d;
      if (nb.delta(m, st, nb.decode(idx), d)) consider(idx, d);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t oh = __shfl_down_sync(0xffffffffu, tb_h, o), os = __shfl_down_sync(0xffffffffu, tb_s, o);
    const uint32_t on = __shfl_down_sync(0xffffffffu, tb_n, o), of = __shfl_down_sync(0xffffffffu, tb_first, o);
    t_acc += __shfl_down_sync(0xffffffffu, t_acc, o);
    if (on && (!tb_n || score_less(tb_h, tb_s, oh, os))) {
      tb_h = oh; tb_s = os; tb_n = on; tb_first = of;
    } else if (on && tb_n && oh == tb_h && os == tb_s) {
      tb_n += on;
      tb_first = min(tb_first, of);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh_h[warp] = tb_h; sh_s[warp] = tb_s; sh_n[warp] = tb_n; sh_f[warp] = tb_first; sh_a[warp] = t_acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ChunkPartial cp{0, 0, 0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu};
    for (int w = 0; w < 8; ++w) {
      cp.n_accepted += sh_a[w];
      if (!sh_n[w]) continue;
      if (!cp.n_best || score_less(cp.best_h, cp.best_s, sh_h[w], sh_s[w])) {
        cp.best_h = sh_h[w]; cp.best_s = sh_s[w]; cp.n_best = sh_n[w]; cp.first_idx = sh_f[w];
      } else if (sh_h[w] == cp.best_h && sh_s[w] == cp.best_s) {
        cp.n_best += sh_n[w];
        cp.first_idx = min(cp.first_idx, sh_f[w]);
      }
    }
    a.partials[(size_t)r * gridDim.x + blockIdx.x] = cp;
  }
}

// Ordered search over pull indices [lo, hi): the `want`-th (1-based) candidate that is accepted (mode 0) or
// accepted and equal to (bh, bs) (mode 1). All threads participate; result in *s_out (untouched when absent).
template <class NB>
__device__ __forceinline__ void index_find(const DevModel& m, const char* st, const NB& nb, const ForageDev& f,
                                           uint32_t lo, uint32_t hi, int mode, int64_t bh, int64_t bs, uint32_t want,
                                           int64_t ch, int64_t csf, int64_t lh, int64_t ls, int64_t th, int64_t ts,
                                           uint32_t* scratch, uint32_t* s_out) {
  uint32_t seen = 0;
  for (uint32_t base = lo; base < hi; base += blockDim.x) {
    const uint32_t idx = base + threadIdx.x;
    uint32_t hit = 0;
    if (idx < hi) {
      Score2 d;
      if (nb.delta(m, st, nb.decode(idx), d)) {
        const int64_t oh = ch + d.hard, os = csf + d.soft;
        hit = (accept_score(f.acceptor, oh, os, lh, ls, th, ts) && (mode == 0 || (oh == bh && os == bs))) ? 1 : 0;
      }
    }
    uint32_t tot;
    const uint32_t incl = block_scan_u32(hit, scratch, &tot);
    if (hit && seen + incl == want) *s_out = idx;
    seen += tot;
    __syncthreads();
    if (seen >= want) break;
  }
}

// One CTA per replica: AcceptedCount cut, best over chunks, tie rule, winner (forager.rs:70-155, 167-425).
// Dynamic shared memory: NB::table_words(..) uint32.
template <class NB>
__global__ void __launch_bounds__(256) index_finish_kernel(const __grid_constant__ DevModel m, const IndexStepArgs a,
                                                           uint32_t n_chunks, uint32_t* __restrict__ out_index,
                                                           int64_t* __restrict__ out_best,
                                                           uint32_t* __restrict__ out_evaluated,
                                                           uint32_t* __restrict__ out_winner_rows) {
  extern __shared__ __align__(16) uint32_t tables[];
  __shared__ uint32_t scratch[33];
  __shared__ uint32_t s_idx, s_cut_chunk, s_cut_rank, s_any, s_cstar, s_jstar;
  __shared__ int64_t s_bh, s_bs;
  __shared__ ChunkPartial s_cutp;
  const uint32_t r = blockIdx.x;
  const char* st = m.state + (size_t)r * m.block_bytes;
  NB nb;
  nb.build(m, st, tables, a.min_size, a.max_size);
  const uint32_t total = nb.total();
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  const ChunkPartial* P = a.partials + (size_t)r * n_chunks;
  int64_t lh = 0, ls = 0, th = 0, ts = 0;
  if (a.ref_scores) {
    lh = a.ref_scores[r * 4 + 0];
    ls = a.ref_scores[r * 4 + 1];
    th = a.ref_scores[r * 4 + 2];
    ts = a.ref_scores[r * 4 + 3];
  }
  auto chunk_lo = [&](uint32_t c) { return (uint32_t)min((uint64_t)c * a.per_chunk, (uint64_t)total); };
  auto chunk_hi = [&](uint32_t c) { return (uint32_t)min((uint64_t)(c + 1) * a.per_chunk, (uint64_t)total); };
  uint32_t limit_idx = 0xFFFFFFFFu;  // last pull index that counts
  uint32_t n_eff = n_chunks;
  if (threadIdx.x == 0) {
    s_cut_chunk = n_chunks;
    s_cutp = ChunkPartial{0, 0, 0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu};
  }
  __syncthreads();
  if (a.f.accepted_limit > 0) {
    if (threadIdx.x == 0) {
      uint32_t seen = 0;
      for (uint32_t c = 0; c < n_chunks; ++c) {
        if (seen + P[c].n_accepted >= a.f.accepted_limit) {
          s_cut_chunk = c;
          s_cut_rank = a.f.accepted_limit - seen;
          break;
        }
        seen += P[c].n_accepted;
      }
    }
    __syncthreads();
    if (s_cut_chunk < n_chunks) {
      const uint32_t c_lo = chunk_lo(s_cut_chunk), c_hi = chunk_hi(s_cut_chunk);
      index_find(m, st, nb, a.f, c_lo, c_hi, 0, 0, 0, s_cut_rank, ch, csf, lh, ls, th, ts, scratch, &s_idx);
      __syncthreads();
      limit_idx = s_idx;
      n_eff = s_cut_chunk + 1;
      // partial of the cut chunk truncated at the limit
      int64_t tb_h = 0, tb_s = 0;
      uint32_t tb_n = 0, tb_first = 0xFFFFFFFFu;
      for (uint32_t idx = c_lo + threadIdx.x; idx <= limit_idx && idx < c_hi; idx += blockDim.x) {
        Score2 d;
        if (!nb.delta(m, st, nb.decode(idx), d)) continue;
        const int64_t oh = ch + d.hard, os = csf + d.soft;
        if (!accept_score(a.f.acceptor, oh, os, lh, ls, th, ts)) continue;
        if (tb_n == 0 || score_less(tb_h, tb_s, oh, os)) {
          tb_h = oh; tb_s = os; tb_n = 1; tb_first = idx;
        } else if (tb_h == oh && tb_s == os) {
          tb_n++;
        }
      }
      for (uint32_t t = 0; t < blockDim.x; ++t) {  // serialised merge (rare path)
        if (threadIdx.x == t && tb_n) {
          if (!s_cutp.n_best || score_less(s_cutp.best_h, s_cutp.best_s, tb_h, tb_s)) {
            s_cutp.best_h = tb_h; s_cutp.best_s = tb_s; s_cutp.n_best = tb_n; s_cutp.first_idx = tb_first;
          } else if (s_cutp.best_h == tb_h && s_cutp.best_s == tb_s) {
            s_cutp.n_best += tb_n;
            s_cutp.first_idx = min(s_cutp.first_idx, tb_first);
          }
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  const uint32_t cut_chunk = s_cut_chunk;
  const uint32_t evaluated = limit_idx == 0xFFFFFFFFu ? total : limit_idx + 1;
  if (threadIdx.x == 0) {
    int64_t bh = 0, bs = 0;
    uint32_t any = 0;
    for (uint32_t c = 0; c < n_eff; ++c) {
      const ChunkPartial p = c == cut_chunk ? s_cutp : P[c];
      if (!p.n_best) continue;
      if (!any || score_less(bh, bs, p.best_h, p.best_s)) {
        bh = p.best_h;
        bs = p.best_s;
      }
      any = 1;
    }
    s_bh = bh;
    s_bs = bs;
    s_any = any;
    s_jstar = 1;
  }
  __syncthreads();
  const int64_t bh = s_bh, bs = s_bs;
  if (!s_any) {
    if (threadIdx.x == 0) {
      out_index[r] = 0xFFFFFFFFu;
      out_best[r * 2] = 0;
      out_best[r * 2 + 1] = 0;
      if (out_evaluated) out_evaluated[r] = evaluated;
      if (out_winner_rows) ((uint4*)out_winner_rows)[r] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    }
    return;
  }
  uint32_t mtot = 0;
  for (uint32_t c = 0; c < n_eff; ++c) {
    const ChunkPartial p = c == cut_chunk ? s_cutp : P[c];
    if (p.n_best && p.best_h == bh && p.best_s == bs) mtot += p.n_best;
  }
  if (a.f.tie_mode == 1) {  // reservoir_pick replay: the last k with splitmix64(..) % k == 0 wins (forager.rs:143-155)
    const uint64_t seed = a.step_seeds ? a.step_seeds[r] : 0;
    uint32_t best_k = 1;
    for (uint32_t kk = 2 + threadIdx.x; kk <= mtot; kk += blockDim.x) {
      const uint64_t mixed = splitmix64_dev(seed ^ ((uint64_t)kk * 0x9E3779B97F4A7C15ull) ^ 0xF04A63E239B74D11ull);
      if (mixed % kk == 0) best_k = kk;
    }
    atomicMax(&s_jstar, best_k);
  }
  __syncthreads();
  const uint32_t want = s_jstar;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t before = 0;
    for (uint32_t c = 0; c < n_eff; ++c) {
      const ChunkPartial p = c == cut_chunk ? s_cutp : P[c];
      const uint32_t nb_ = (p.n_best && p.best_h == bh && p.best_s == bs) ? p.n_best : 0;
      if (before + nb_ >= want) {
        s_cstar = c;
        s_jstar = want - before;
        s_idx = p.first_idx;
        break;
      }
      before += nb_;
    }
  }
  __syncthreads();
  if (s_jstar > 1) {
    const uint32_t c_lo = chunk_lo(s_cstar);
    const uint32_t c_hi = min(chunk_hi(s_cstar), limit_idx == 0xFFFFFFFFu ? 0xFFFFFFFFu : limit_idx + 1);
    const uint32_t j = s_jstar;
    __syncthreads();
    index_find(m, st, nb, a.f, c_lo, c_hi, 1, bh, bs, j, ch, csf, lh, ls, th, ts, scratch, &s_idx);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const uint32_t idx = s_idx;
    out_index[r] = idx;
    out_best[r * 2] = bh;
    out_best[r * 2 + 1] = bs;
    if (out_evaluated) out_evaluated[r] = evaluated;
    if (out_winner_rows) ((uint4*)out_winner_rows)[r] = nb.decode(idx);
  }
}
