// Device-side nearby list-change neighbourhood, fused with scoring and the forager replay.
//
// Reference: NearbyListChangeMoveSelector — solverforge-solver/src/heuristic/selector/
//   list_kernel/nearby_change.rs:102-232 (scan order, skip rules, min(dp, len-1) reference element),
//   nearby_list_support.rs:3-34 (stable bounded top-k: ties keep scan order),
//   crates/solverforge-cvrp/src/meters.rs:10-28 (MatrixDistanceMeter: finite cell as f64).
//
// The reference scans ~n slots per source element. Here the distance from source x to a slot depends
// only on the slot's reference element y, and every routed element y is the reference of slot
// (route(y), pos(y)) plus, when it is last in its route, of the append slot (route(y), len). So walking
// the STATIC list of x's neighbours in increasing distance (built once at commit) and mapping each
// neighbour through the per-replica inverse index pos_of[] enumerates slots in increasing distance;
// the walk stops as soon as max_nearby slots are held and the next neighbour is strictly farther than
// the current max_nearby-th. Ties are ordered by the slot's scan index, exactly like the reference's
// stable insertion sort. One warp handles one source: NB_BATCH neighbours per trip are expanded to
// slot keys (distance << scan_bits | scan index), compacted through shared memory, sorted with a
// 32-lane bitonic network and merged into the kept top-32. The first `count` lanes then score one
// candidate each (same record math as score_list_change_fast_kernel) and feed the forager partial.
//
// Canonical order only (SelectionOrder::Original): pull index = source_flat_position * count + lane.
#pragma once
#include "sfgpu_kernels.cuh"

#define NB_BATCH 28  // neighbours expanded per trip: 28 + their append slots rarely exceed 32 keys


template <typename KEY>
struct KeyTraits;
template <>
struct KeyTraits<uint32_t> {
  static __device__ __forceinline__ uint32_t maxkey() { return 0xFFFFFFFFu; }
};
template <>
struct KeyTraits<uint64_t> {
  static __device__ __forceinline__ uint64_t maxkey() { return 0xFFFFFFFFFFFFFFFFull; }
};

// bitonic sort of 32 keys, one per lane, ascending by lane
template <typename KEY>
__device__ __forceinline__ KEY warp_sort32(KEY k, const uint32_t lane) {
#pragma unroll
  for (uint32_t size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      const KEY p = __shfl_xor_sync(0xffffffffu, k, stride);
      const bool up = (lane & size) == 0 || size == 32;
      const bool lower = (lane & stride) == 0;
      k = (lower == up) ? (k < p ? k : p) : (k > p ? k : p);
    }
  }
  return k;
}

// ascending sort of 32 nearly-sorted keys: odd-even transposition rounds until sorted. The batch keys
// arrive in increasing distance (the static neighbour list), so only ties in distance are out of
// order and one or two rounds suffice; any input is sorted correctly (at most 16 rounds).
template <typename KEY>
__device__ __forceinline__ KEY warp_sort32_adaptive(KEY k, const uint32_t lane) {
  const bool odd = lane & 1;
  const uint32_t partner = (odd ? lane + 1 : lane - 1) & 31;
  const bool edge = lane == 0 || lane == 31;
  while (true) {
    const KEY nx = __shfl_down_sync(0xffffffffu, k, 1);
    if (!__any_sync(0xffffffffu, lane < 31 && nx < k)) break;
    KEY p = __shfl_xor_sync(0xffffffffu, k, 1);  // pairs (0,1), (2,3), ...
    k = odd ? (k > p ? k : p) : (k < p ? k : p);
    p = __shfl_sync(0xffffffffu, k, partner);     // pairs (1,2), (3,4), ...
    if (!edge) k = odd ? (k < p ? k : p) : (k > p ? k : p);
  }
  return k;
}

// merges a sorted batch b into the kept sorted list L: the 32 smallest of the union, ascending
template <typename KEY>
__device__ __forceinline__ KEY warp_merge32(KEY L, KEY b, const uint32_t lane) {
  const KEY rev = __shfl_sync(0xffffffffu, b, 31 - lane);
  KEY v = L < rev ? L : rev;  // bitonic sequence holding the 32 smallest of the union
#pragma unroll
  for (uint32_t stride = 16; stride > 0; stride >>= 1) {
    const KEY p = __shfl_xor_sync(0xffffffffu, v, stride);
    const bool lower = (lane & stride) == 0;
    v = lower ? (v < p ? v : p) : (v > p ? v : p);
  }
  return v;
}

struct NearbyView {  // pointers into the (staged or global) fast section of one replica block
  const uint4* rr;
  const uint4* pr;
  const uint4* sr;
  const uint32_t* pos_of;
};

// Returns, in lane l, the l-th nearest destination slot key of source flat position f
// (maxkey when fewer than l+1 exist). All 32 lanes participate. `buf` = 64 KEY slots of this warp.
template <typename KEY, typename CELL>
__device__ __forceinline__ KEY nearby_gen_source(const DevModel& m, const NearbyView& v, const uint32_t scan_bits,
                                                 const uint32_t f, const uint32_t K, const uint32_t lane,
                                                 KEY* __restrict__ buf, uint32_t& x, uint32_t& se, uint32_t& sp,
                                                 uint4& prec, uint4& rsrc) {
  const KEY MAXK = KeyTraits<KEY>::maxkey();
  prec = v.pr[f];
  x = prec.x;
  se = prec.w;
  rsrc = v.rr[se];
  sp = f - rsrc.x;
  const uint32_t slen = rsrc.y;
  const uint32_t g_own = rsrc.x + se;  // first slot index of the source's own route
  const uint32_t* __restrict__ nb = m.nbr + (size_t)x * m.nbr_stride;
  const CELL* __restrict__ mrow = (const CELL*)m.fm_row + (size_t)x * m.cons[m.fast_pc].n0;
  KEY L = MAXK;
  uint32_t d_next = 0;  // distance of the first neighbour of the next trip (loaded by lane NB_BATCH)
  for (uint32_t start = 0; start < m.nbr_stride; start += NB_BATCH) {
    if (start > 0) {
      const KEY kth = __shfl_sync(0xffffffffu, L, K - 1);
      if (kth != MAXK && (KEY)d_next > (kth >> scan_bits)) break;  // strictly farther: cannot enter the top K
    }
    KEY k0 = MAXK, k1 = MAXK;
    const uint32_t idx = start + lane;
    uint32_t y = 0, dyu = 0;
    if (lane <= NB_BATCH && idx < m.nbr_stride) {
      y = __ldg(nb + idx);
      dyu = (uint32_t)__ldg(mrow + y);
    }
    d_next = __shfl_sync(0xffffffffu, dyu, NB_BATCH);
    if (lane < NB_BATCH && idx < m.nbr_stride) {
      const uint32_t where = v.pos_of[y];
      if (where != 0xFFFFFFFFu) {
        const uint32_t e = where >> 16, py = where & 0xFFFFu;
        const KEY dy = (KEY)dyu;
        const uint4 re = v.rr[e];
        const uint32_t g = re.x + e + py;
        const bool own = e == se;
        // slot (e, py): the element at py is its reference
        if (!(own && (py == sp || py == sp + 1))) {
          const uint32_t scan = own ? py : (g < g_own ? g + slen + 1 : g);
          k0 = (dy << scan_bits) | scan;
        }
        // append slot (e, len): its reference is the last element
        if (py + 1 == re.y) {
          const uint32_t dp = re.y;
          if (!(own && (dp == sp || dp == sp + 1))) {
            const uint32_t scan = own ? dp : (g + 1 < g_own ? g + 1 + slen + 1 : g + 1);
            k1 = (dy << scan_bits) | scan;
          }
        }
      }
    }
    // compaction in list order (an append-slot key right behind its element's key): the sequence
    // stays sorted by distance, only equal distances may be out of order
    const uint32_t m0 = __ballot_sync(0xffffffffu, k0 != MAXK), m1 = __ballot_sync(0xffffffffu, k1 != MAXK);
    const uint32_t total = __popc(m0) + __popc(m1);
    const uint32_t lt = (1u << lane) - 1;
    const uint32_t at = __popc(m0 & lt) + __popc(m1 & lt);
    buf[lane] = MAXK;
    buf[lane + 32] = MAXK;
    __syncwarp();
    if (k0 != MAXK) buf[at] = k0;
    if (k1 != MAXK) buf[at + (k0 != MAXK ? 1 : 0)] = k1;
    __syncwarp();
    KEY b = warp_sort32_adaptive(buf[lane], lane);
    L = (start == 0) ? b : warp_merge32(L, b, lane);
    if (total > 32) {  // rare: more than 32 slots from one batch
      b = warp_sort32(buf[lane + 32], lane);
      L = warp_merge32(L, b, lane);
    }
    __syncwarp();
  }
  return lane < K ? L : MAXK;
}

// per-launch constants of the fast record math, narrowed once (FastArith)
template <typename S>
struct NearbyConsts {
  S pc_a, ls_a, ls_b;
  bool pc_hard, ls_hard;
  uint32_t dim;
};
template <int SUM_FN, typename S>
__device__ __forceinline__ NearbyConsts<S> nearby_consts(const DevModel& m) {
  NearbyConsts<S> c;
  const ConsDev& pc = m.cons[m.fast_pc];
  c.pc_a = (S)(pc.sign < 0 ? -pc.w.a : pc.w.a);
  c.pc_hard = pc.w.level == 0;
  c.dim = pc.n0;
  const ConsDev& ls = m.cons[SUM_FN >= 0 ? m.fast_ls : 0];
  c.ls_a = (S)(ls.sign < 0 ? -ls.w.a : ls.w.a);
  c.ls_b = sizeof(S) == 4 ? clamp_to<S>(ls.w.b, 1ll << 30) : (S)ls.w.b;
  c.ls_hard = ls.w.level == 0;
  return c;
}

// score delta of candidate (source x -> insertion slot g) with the fast-path record math; returns the destination
// (e, dp) and the slot's reference element (| append << 31) too. d_ref = d(x, reference element). Identical
// arithmetic to score_list_change_fast_kernel.
template <int SUM_FN, typename CELL, typename S>
__device__ __forceinline__ void nearby_score_slot(const DevModel& m, const NearbyConsts<S>& c, const NearbyView& v,
                                                  const uint32_t g, const uint32_t d_ref, const uint32_t x,
                                                  const uint32_t se, const uint4 prec, const uint4 rsrc, S& dh, S& ds,
                                                  uint32_t& de, uint32_t& dp, uint32_t& ident) {
  typedef typename FastArith<CELL>::US US;
  const uint4 s = v.sr[g];
  de = s.w >> 16;
  dp = s.w & 0xFFFFu;
  const uint4 rd = v.rr[de];
  dh = 0;
  ds = 0;
  ident = dp < rd.y ? s.y : (s.x | 0x80000000u);
  {
    const CELL* __restrict__ mrow = (const CELL*)m.fm_row + (size_t)x * c.dim;
    const CELL* __restrict__ mcol = (const CELL*)m.fm_col + (size_t)x * c.dim;
    // d(x, b): for a non-append slot b is the slot's reference element, whose distance the caller holds
    const int32_t xb = dp < rd.y ? (int32_t)d_ref : (int32_t)__ldg(mrow + s.y);
    const int32_t ax = (int32_t)__ldg(mcol + s.x);
    const S d = (S)((US)c.pc_a * (US)(S)((int32_t)prec.y + ax + xb - (int32_t)s.z));
    if (c.pc_hard) dh += d; else ds += d;
  }
  if (SUM_FN >= 0 && se != de) {
    const S ss = sizeof(S) == 4 ? (S)rsrc.z : (S)(((uint64_t)rsrc.w << 32) | rsrc.z);
    const S sd = sizeof(S) == 4 ? (S)rd.z : (S)(((uint64_t)rd.w << 32) | rd.z);
    const S d = list_sum_delta<SUM_FN, S, US>(c.ls_a, c.ls_b, (S)(int32_t)prec.z, ss, sd);
    if (c.ls_hard) dh += d; else ds += d;
  }
}

// flat slot index of the scan index held in the low bits of a key (inverse of the scan order of nearby_gen_source)
__device__ __forceinline__ uint32_t nearby_slot_of_scan(const uint32_t scan, const uint32_t se, const uint4 rsrc) {
  const uint32_t slen = rsrc.y, g_own = rsrc.x + se;
  return scan <= slen ? g_own + scan : (scan - (slen + 1) < g_own ? scan - (slen + 1) : scan);
}

template <int SUM_FN, typename KEY, typename CELL, typename S>
__device__ __forceinline__ void nearby_score(const DevModel& m, const NearbyConsts<S>& c, const NearbyView& v,
                                             const uint32_t scan_bits, const KEY key, const uint32_t x,
                                             const uint32_t se, const uint4 prec, const uint4 rsrc, S& dh, S& ds,
                                             uint32_t& de, uint32_t& dp) {
  uint32_t ident;
  const uint32_t g = nearby_slot_of_scan((uint32_t)(key & (((KEY)1 << scan_bits) - 1)), se, rsrc);
  nearby_score_slot<SUM_FN, CELL, S>(m, c, v, g, (uint32_t)(key >> scan_bits), x, se, prec, rsrc, dh, ds, de, dp, ident);
}

// warp-level lexicographic max of the deltas (dh, ds) over the lanes with `acc`: returns the
// multiplicity mask of the best; accm = accepted lanes. 32-bit deltas use two REDUX instructions.
__device__ __forceinline__ uint32_t warp_best_mask(bool acc, int32_t dh, int32_t ds, uint32_t& accm) {
  accm = __ballot_sync(0xffffffffu, acc);
  if (!accm) return 0;
  const int32_t bh = __reduce_max_sync(0xffffffffu, acc ? dh : INT32_MIN);
  const bool top = acc && dh == bh;
  const int32_t bs = __reduce_max_sync(0xffffffffu, top ? ds : INT32_MIN);
  return __ballot_sync(0xffffffffu, top && ds == bs);
}
__device__ __forceinline__ uint32_t warp_best_mask(bool acc, int64_t dh, int64_t ds, uint32_t& accm) {
  accm = __ballot_sync(0xffffffffu, acc);
  if (!accm) return 0;
  int64_t bh = dh, bs = ds;
  uint32_t any = acc ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t ph = __shfl_xor_sync(0xffffffffu, bh, o), ps = __shfl_xor_sync(0xffffffffu, bs, o);
    const uint32_t pa = __shfl_xor_sync(0xffffffffu, any, o);
    if (pa && (!any || score_less(bh, bs, ph, ps))) {
      bh = ph;
      bs = ps;
    }
    any |= pa;
  }
  return __ballot_sync(0xffffffffu, acc && dh == bh && ds == bs);
}

// ---- nearby list SWAP (NearbySwapCursor, list_kernel/nearby_swap.rs:99-262) -------------------
// Destinations of source (se, sp): the later positions of its own list, then every position of the
// later entities — exactly the flat positions g > f in scan order — ranked by d(x, element at g)
// with the stable bounded top-k (ties keep scan order = flat position). Source f has
// min(K, total - 1 - f) candidates; sources without destinations are skipped, so pull indices are
// a prefix sum over sources (swap_prefix).

__device__ __forceinline__ uint32_t swap_count(uint32_t f, uint32_t total, uint32_t K) {
  const uint32_t e = total - 1 - f;
  return e < K ? e : K;
}
// number of candidates pulled before source f
__device__ __forceinline__ uint32_t swap_prefix(uint32_t f, uint32_t total, uint32_t K) {
  const uint32_t n_full = total > K ? total - K : 0;  // sources that yield K candidates
  if (f <= n_full) return f * K;
  // sources n_full .. f-1 yield total-1-f' each
  const uint32_t cnt = f - n_full, first = total - 1 - n_full, last = total - f;  // first down to last
  return n_full * K + (uint32_t)(((uint64_t)(first + last) * cnt) / 2);
}

template <typename KEY, typename CELL>
__device__ __forceinline__ KEY nearby_swap_gen_source(const DevModel& m, const NearbyView& v, const uint32_t scan_bits,
                                                      const uint32_t f, const uint32_t K, const uint32_t total,
                                                      const uint32_t lane, KEY* __restrict__ buf, uint32_t& x,
                                                      uint32_t& se, uint32_t& sp, uint4& prec, uint4& rsrc) {
  const KEY MAXK = KeyTraits<KEY>::maxkey();
  prec = v.pr[f];
  x = prec.x;
  se = prec.w;
  rsrc = v.rr[se];
  sp = f - rsrc.x;
  const CELL* __restrict__ mrow = (const CELL*)m.fm_row + (size_t)x * m.cons[m.fast_pc].n0;
  const uint32_t eligible = total - 1 - f;
  KEY L = MAXK;
  if (eligible <= 64) {
    // few destinations: rank them all directly (the walk below would scan most of the list)
    for (uint32_t base = 0; base < eligible; base += 32) {
      const uint32_t g = f + 1 + base + lane;
      KEY k = MAXK;
      if (base + lane < eligible) k = ((KEY)(uint32_t)__ldg(mrow + v.pr[g].x) << scan_bits) | g;
      k = warp_sort32(k, lane);
      L = base == 0 ? k : warp_merge32(L, k, lane);
    }
    return lane < K ? L : MAXK;
  }
  const uint32_t* __restrict__ nb = m.nbr + (size_t)x * m.nbr_stride;
  uint32_t d_next = 0;
  for (uint32_t start = 0; start < m.nbr_stride; start += 31) {
    if (start > 0) {
      const KEY kth = __shfl_sync(0xffffffffu, L, K - 1);
      if (kth != MAXK && (KEY)d_next > (kth >> scan_bits)) break;  // strictly farther: cannot enter the top K
    }
    const uint32_t idx = start + lane;
    uint32_t dyu = 0;
    KEY k = MAXK;
    if (idx < m.nbr_stride) {
      const uint32_t y = __ldg(nb + idx);
      dyu = (uint32_t)__ldg(mrow + y);
      const uint32_t where = v.pos_of[y];
      if (lane < 31 && where != 0xFFFFFFFFu) {
        const uint32_t g = v.rr[where >> 16].x + (where & 0xFFFFu);
        if (g > f) k = ((KEY)dyu << scan_bits) | g;
      }
    }
    d_next = __shfl_sync(0xffffffffu, dyu, 31);
    k = warp_sort32(k, lane);
    L = start == 0 ? k : warp_merge32(L, k, lane);
  }
  return lane < K ? L : MAXK;
}

// score delta of swapping the source element x (flat f) with the element at flat position g of `key`
// — packed-record restatement of list_swap_delta (ListSwapMove::do_move, list_kernel/swap.rs): the four
// legs around both positions change; adjacent positions share one leg.
template <int SUM_FN, typename KEY, typename CELL, typename S>
__device__ __forceinline__ void nearby_swap_score(const DevModel& m, const NearbyConsts<S>& c, const NearbyView& v,
                                                  const uint32_t scan_bits, const KEY key, const uint32_t f,
                                                  const uint32_t x, const uint32_t se, const uint4 prec,
                                                  const uint4 rsrc, S& dh, S& ds, uint32_t& de, uint32_t& dp) {
  typedef typename FastArith<CELL>::US US;
  const uint32_t g = (uint32_t)(key & (((KEY)1 << scan_bits) - 1));
  const uint4 py = v.pr[g];
  const uint32_t y = py.x;
  de = py.w;
  const uint4 rd = v.rr[de];
  dp = g - rd.x;
  const uint4 sx0 = v.sr[f + se], sx1 = v.sr[f + se + 1];  // legs (a, x) and (x, b)
  const uint4 sy0 = v.sr[g + de], sy1 = v.sr[g + de + 1];  // legs (c, y) and (y, d)
  const CELL* __restrict__ xrow = (const CELL*)m.fm_row + (size_t)x * c.dim;
  const CELL* __restrict__ xcol = (const CELL*)m.fm_col + (size_t)x * c.dim;
  const CELL* __restrict__ yrow = (const CELL*)m.fm_row + (size_t)y * c.dim;
  const CELL* __restrict__ ycol = (const CELL*)m.fm_col + (size_t)y * c.dim;
  int32_t dl;
  if (g == f + 1 && de == se) {
    // a x y d -> a y x d
    dl = (int32_t)__ldg(ycol + sx0.x) + (int32_t)__ldg(yrow + x) + (int32_t)__ldg(xrow + sy1.y) - (int32_t)sx0.z -
         (int32_t)sx1.z - (int32_t)sy1.z;
  } else {
    dl = (int32_t)__ldg(ycol + sx0.x) + (int32_t)__ldg(yrow + sx1.y) - (int32_t)sx0.z - (int32_t)sx1.z +
         (int32_t)__ldg(xcol + sy0.x) + (int32_t)__ldg(xrow + sy1.y) - (int32_t)sy0.z - (int32_t)sy1.z;
  }
  dh = 0;
  ds = 0;
  const S d = (S)((US)c.pc_a * (US)(S)dl);
  if (c.pc_hard) dh += d; else ds += d;
  if (SUM_FN >= 0 && se != de) {
    const S ss = sizeof(S) == 4 ? (S)rsrc.z : (S)(((uint64_t)rsrc.w << 32) | rsrc.z);
    const S sd = sizeof(S) == 4 ? (S)rd.z : (S)(((uint64_t)rd.w << 32) | rd.z);
    // x leaves se and y enters it: the net transfer se -> de is val(x) - val(y)
    const S d2 = list_sum_delta<SUM_FN, S, US>(c.ls_a, c.ls_b, (S)((US)(S)(int32_t)prec.z - (US)(S)(int32_t)py.z), ss, sd);
    if (c.ls_hard) dh += d2; else ds += d2;
  }
}

// number of candidates every source yields: min(K, slots of non-empty routes - 2), computed by one warp (every lane
// returns it)
__device__ __forceinline__ uint32_t nearby_count_warp(const DevModel& m, const uint4* rr, uint32_t K, uint32_t lane) {
  uint32_t slots = 0;
  for (uint32_t o = lane; o < m.n_owners; o += 32) {
    const uint32_t len = rr[o].y;
    slots += len ? len + 1 : 0;
  }
  slots = __reduce_add_sync(0xffffffffu, slots);
  const uint32_t valid = slots >= 2 ? slots - 2 : 0;
  return valid < K ? valid : K;
}

// grid = (chunks, R), 256 threads. Warp w of chunk c handles sources c_lo + w, c_lo + w + 8, ...
template <int SUM_FN, typename KEY, typename CELL, int MOVE = MOVE_CHANGE>
__global__ void __launch_bounds__(256) nearby_step_kernel(const __grid_constant__ DevModel m, const NearbyArgs a) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_count;
  __shared__ KEY s_buf[8][64];
  const uint32_t r = blockIdx.y;
  stage_block(smem, m.state + (size_t)r * m.block_bytes, m.fast_stage_bytes, &bar);
  NearbyView v;
  v.rr = (const uint4*)(smem + m.off_route_rec);
  v.pr = (const uint4*)(smem + m.off_pos_rec);
  v.sr = (const uint4*)(smem + m.off_slot_rec);
  v.pos_of = (const uint32_t*)(smem + m.off_pos_of);
  typedef typename FastArith<CELL>::S S;
  const int64_t* cs = (const int64_t*)(smem + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  if (threadIdx.x < 32) {
    const uint32_t c = nearby_count_warp(m, v.rr, a.max_nearby, threadIdx.x);
    if (threadIdx.x == 0) s_count = c;
  }
  __syncthreads();
  const uint32_t count = s_count;
  const uint32_t total = v.rr[m.n_owners - 1].x + v.rr[m.n_owners - 1].y;  // routed elements
  // acceptor references as thresholds on the score delta (rel_threshold)
  S lh, ls, th, ts;
  rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 0] : 0, ch, lh);
  rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 1] : 0, csf, ls);
  rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 2] : 0, ch, th);
  rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 3] : 0, csf, ts);
  const NearbyConsts<S> nc = nearby_consts<SUM_FN, S>(m);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t per = (total + gridDim.x - 1) / gridDim.x;
  const uint32_t c_lo = per * blockIdx.x, c_hi = min(c_lo + per, total);
  const size_t out_base = (size_t)r * m.elem_cap * a.max_nearby;
  if (a.out_offsets && blockIdx.x == 0 && threadIdx.x == 0) {
    a.out_offsets[r] = out_base;
    if (r == m.R - 1) a.out_offsets[m.R] = out_base + (size_t)m.elem_cap * a.max_nearby;
  }
  if (a.out_rows && blockIdx.x == 0) {
    // padding of the materialised batch: unrouted capacity [total, elem_cap) holds not-doable rows
    for (size_t q = (size_t)total * a.max_nearby + threadIdx.x; q < (size_t)m.elem_cap * a.max_nearby; q += blockDim.x) {
      ((uint4*)a.out_rows)[out_base + q] = make_uint4(0xFFFFFFFFu, 0, 0xFFFFFFFFu, 0);
      if (a.out_scores) {
        ((longlong2*)a.out_scores)[out_base + q] = make_longlong2(0, 0);
        a.out_doable[out_base + q] = 0;
      }
    }
  }
  for (uint32_t f = c_lo + warp; f < c_hi; f += 8) {
    uint32_t x, se, sp;
    uint4 prec, rsrc;
    KEY key;
    if (MOVE == MOVE_SWAP)
      key = nearby_swap_gen_source<KEY, CELL>(m, v, a.scan_bits, f, a.max_nearby, total, lane, s_buf[warp], x, se, sp,
                                              prec, rsrc);
    else
      key = nearby_gen_source<KEY, CELL>(m, v, a.scan_bits, f, a.max_nearby, lane, s_buf[warp], x, se, sp, prec, rsrc);
    const uint32_t cnt = MOVE == MOVE_SWAP ? swap_count(f, total, a.max_nearby) : count;
    const bool have = lane < cnt && key != KeyTraits<KEY>::maxkey();
    S dh = 0, ds = 0;
    uint32_t de = 0, dp = 0;
    if (have) {
      if (MOVE == MOVE_SWAP)
        nearby_swap_score<SUM_FN, KEY, CELL, S>(m, nc, v, a.scan_bits, key, f, x, se, prec, rsrc, dh, ds, de, dp);
      else
        nearby_score<SUM_FN, KEY, CELL, S>(m, nc, v, a.scan_bits, key, x, se, prec, rsrc, dh, ds, de, dp);
    }
    const bool acc = have && accept_delta<S>(a.f.acceptor, dh, ds, lh, ls, th, ts);
    uint32_t accm;
    const uint32_t eq = warp_best_mask(acc, dh, ds, accm);
    if (lane == (eq ? __ffs(eq) - 1 : 0)) {
      SrcPartial p;
      p.best_h = eq ? ch + (int64_t)dh : 0;
      p.best_s = eq ? csf + (int64_t)ds : 0;
      p.n_best = __popc(eq);
      p.n_accepted = __popc(accm);
      p.first_lane = eq ? __ffs(eq) - 1 : 0;
      p.pad = 0;
      a.partials[(size_t)r * m.elem_cap + f] = p;
    }
    if (a.out_rows && lane < a.max_nearby) {
      // materialised batch: fixed stride of max_nearby rows per source; missing rows are not doable
      const size_t q = out_base + (size_t)f * a.max_nearby + lane;
      ((uint4*)a.out_rows)[q] = have ? make_uint4(se, sp, de, dp) : make_uint4(0xFFFFFFFFu, 0, 0xFFFFFFFFu, 0);
      if (a.out_scores) {
        longlong2 o2;
        o2.x = have ? ch + (int64_t)dh : 0;
        o2.y = have ? csf + (int64_t)ds : 0;
        ((longlong2*)a.out_scores)[q] = o2;
        a.out_doable[q] = have ? 1 : 0;
      }
    }
  }
}

// ---- retained neighbourhood (incremental step) ------------------------------------------------------------------
// A committed ListChange move x: (A, p) -> (B, q) leaves most of the neighbourhood as it was. Per (replica, source
// element) the kernel below keeps the score deltas of the source's candidates, the reference element of every
// candidate slot and the k-th kept distance, and after a commit redoes only what the move can have changed:
//   tier 3 (generate + score, as nearby_step_kernel): sources in route A or B (own legs, own exclusions), and sources
//           within their k-th distance of a "hot" element — the moved element (its slot changes route, so its place
//           among equal distances changes) or an element whose append slot appeared / vanished (old / new last
//           element of A and B). Nothing else can enter, leave or reorder a kept list: the slot set is one slot per
//           routed element plus one append slot per non-empty route, distances are static, and the scan order of two
//           slots of unmoved elements never changes (routes keep their index, positions shift together).
//   tier 2 (score only, from the kept reference elements through pos_of): sources with a candidate into A or B
//           (8-bit route code per lane) — positions, legs around the edit and the route loads changed, the slot set
//           did not; only those lanes are re-scored.
//   tier 1 (kept deltas): everything else. The acceptor thresholds move every step, so acceptance, the best and its
//           multiplicity are always recomputed from the deltas.
// Keys are not kept: the finish kernel regenerates the one or two sources whose lanes matter. The protocol lives in
// nbc_tag (sfgpu_dev.cuh): nearby_finish_kernel marks the cache current, apply_list_kernel records the one committed
// move, every other writer of a replica block (init, restore, other move kinds) invalidates.
// Work is laid out for throughput, not one warp per source: (A) one THREAD per source classifies it from its meta
// word; tier 2 scans the kept reference elements (each carries its route), re-scores the few that sit in route A or B
// through pos_of and stores their deltas; tiers 1 and 2 then fold the kept deltas into the source's partial on the
// spot (batched independent loads, no cross-lane traffic). Tier-3 sources go to a shared-memory work list and (B) one
// warp per such source generates and scores as nearby_step_kernel does.

// kept reference element of a candidate slot: element | append << 15 | route << 16 (n_elem_rows <= 32768)
__device__ __forceinline__ uint32_t nbc_ident(uint32_t ident31, uint32_t route) {
  return (ident31 & 0x7FFFu) | ((ident31 >> 31) << 15) | (route << 16);
}

// the partial of one source from its kept deltas, sequentially in lane order (== warp_best_mask over the lanes);
// branch-free so the 32 sources of a warp stay converged, NBC_FOLD_BATCH loads in flight at a time
#define NBC_FOLD_BATCH 10
struct NbcFold {
  int32_t bh = INT32_MIN, bs = INT32_MIN;  // below every delta (deltas are strictly inside the int32 range)
  uint32_t n_best = 0, n_acc = 0, first = 0;
  __device__ __forceinline__ void add(const int32_t dx, const int32_t dy, const uint32_t j, const bool valid, const int acceptor,
                                      const int32_t lh, const int32_t ls, const int32_t th, const int32_t ts) {
    const bool in = valid && dx != INT32_MIN && accept_delta<int32_t>(acceptor, dx, dy, lh, ls, th, ts);
    const bool better = in && lex_less(bh, bs, dx, dy);
    const bool equal = in && dx == bh && dy == bs;
    n_acc += in ? 1u : 0u;
    n_best = better ? 1u : n_best + (equal ? 1u : 0u);
    first = better ? j : first;
    bh = better ? dx : bh;
    bs = better ? dy : bs;
  }
  __device__ __forceinline__ void store(const int64_t ch, const int64_t csf, SrcPartial* out) const {
    SrcPartial p;
    p.best_h = n_best ? ch + (int64_t)bh : 0;
    p.best_s = n_best ? csf + (int64_t)bs : 0;
    p.n_best = n_best;
    p.n_accepted = n_acc;
    p.first_lane = first;
    p.pad = 0;
    *out = p;
  }
};
// KC > 0: the row holds exactly KC candidates (count == K == KC, KC % 4 == 0): 128-bit loads, fully unrolled
template <int KC>
__device__ __forceinline__ void nbc_fold_source(const int2* cd, const uint32_t count, const int acceptor, const int32_t lh,
                                                const int32_t ls, const int32_t th, const int32_t ts, const int64_t ch,
                                                const int64_t csf, SrcPartial* out) {
  NbcFold fo;
  if (KC > 0) {
    const int4* __restrict__ c4 = (const int4*)cd;
#pragma unroll
    for (int j0 = 0; j0 < KC / 2; j0 += 5) {
      int4 d[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) d[u] = c4[j0 + u < KC / 2 ? j0 + u : KC / 2 - 1];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        if (j0 + u < KC / 2) {
          fo.add(d[u].x, d[u].y, 2 * (j0 + u), true, acceptor, lh, ls, th, ts);
          fo.add(d[u].z, d[u].w, 2 * (j0 + u) + 1, true, acceptor, lh, ls, th, ts);
        }
      }
    }
  } else {
    for (uint32_t j0 = 0; j0 < count; j0 += NBC_FOLD_BATCH) {
      int2 d[NBC_FOLD_BATCH];
#pragma unroll
      for (uint32_t u = 0; u < NBC_FOLD_BATCH; ++u) d[u] = cd[min(j0 + u, count - 1)];
#pragma unroll
      for (uint32_t u = 0; u < NBC_FOLD_BATCH; ++u) fo.add(d[u].x, d[u].y, j0 + u, j0 + u < count, acceptor, lh, ls, th, ts);
    }
  }
  fo.store(ch, csf, out);
}

// lanes (bit j = candidate j) whose 8-bit route code equals code_a or code_b; codes = 7 words, 4 lanes each; lanes
// 28..31 have no code and always count as hits
__device__ __forceinline__ uint32_t nbc_route_hits(const uint32_t* codes, const uint32_t code_a, const uint32_t code_b,
                                                   const uint32_t count) {
  const uint32_t va = code_a * 0x01010101u, vb = code_b * 0x01010101u;
  uint32_t hits = count > 28 ? 0xF0000000u : 0u;
#pragma unroll
  for (uint32_t w = 0; w < 7; ++w) {
    const uint32_t eq = __vcmpeq4(codes[w], va) | __vcmpeq4(codes[w], vb);  // 0xFF per equal byte
    hits |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * w);
  }
  return hits & (count >= 32 ? 0xFFFFFFFFu : (1u << count) - 1);
}

template <int SUM_FN, int KC>
__global__ void __launch_bounds__(256, 4) nearby_step_cached_kernel(const __grid_constant__ DevModel m, const NearbyArgs a) {
  typedef uint32_t KEY;
  typedef uint16_t CELL;
  typedef int32_t S;
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_n2, s_n3;
  __shared__ uint16_t s_list3[256];  // offsets (f - tile base) of the tile's tier-3 sources
  const uint32_t r = blockIdx.y;
  // protocol words and acceptor references: in flight while the records are staged
  const uint32_t* __restrict__ tag = m.nbc_tag + (size_t)r * NBC_WORDS;
  const uint4 tag0 = *(const uint4*)tag, tag1 = *(const uint4*)(tag + 4), tag2 = *(const uint4*)(tag + 8);
  int64_t refs[4] = {0, 0, 0, 0};
  if (a.ref_scores) {
#pragma unroll
    for (int q = 0; q < 4; ++q) refs[q] = a.ref_scores[r * 4 + q];
  }
  stage_block(smem, m.state + (size_t)r * m.block_bytes, m.fast_stage_bytes, &bar);
  NearbyView v;
  v.rr = (const uint4*)(smem + m.off_route_rec);
  v.pr = (const uint4*)(smem + m.off_pos_rec);
  v.sr = (const uint4*)(smem + m.off_slot_rec);
  v.pos_of = (const uint32_t*)(smem + m.off_pos_of);
  const int64_t* cs = (const int64_t*)(smem + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t count = nearby_count_warp(m, v.rr, a.max_nearby, lane), K = a.max_nearby;
  const uint32_t total = v.rr[m.n_owners - 1].x + v.rr[m.n_owners - 1].y;
  S lh, ls, th, ts;
  rel_threshold(refs[0], ch, lh);
  rel_threshold(refs[1], csf, ls);
  rel_threshold(refs[2], ch, th);
  rel_threshold(refs[3], csf, ts);
  const NearbyConsts<S> nc = nearby_consts<SUM_FN, S>(m);
  const uint32_t per = (total + gridDim.x - 1) / gridDim.x;
  const uint32_t c_lo = per * blockIdx.x, c_hi = min(c_lo + per, total);
  const uint32_t t_state = tag0.x;                                 // NBC_STATE, NBC_K, NBC_COUNT, NBC_A
  const bool valid = t_state != 0 && tag0.y == K && tag0.z == count;
  const bool moved = valid && t_state == 2;
  const uint32_t tA = tag0.w, tB = tag1.x;                         // NBC_B, then the five hot elements
  const uint32_t hot[NBC_N_HOT] = {moved ? tag1.y : 0xFFFFFFFFu, moved ? tag1.z : 0xFFFFFFFFu, moved ? tag1.w : 0xFFFFFFFFu,
                                   moved ? tag2.x : 0xFFFFFFFFu, moved ? tag2.y : 0xFFFFFFFFu};
  const size_t src0 = (size_t)r * m.n_elem_rows;
  SrcPartial* __restrict__ P = a.partials + (size_t)r * m.elem_cap;
  const CELL* __restrict__ fm_row = (const CELL*)m.fm_row;
  for (uint32_t base = c_lo; base < c_hi; base += 256) {
    if (threadIdx.x == 0) {
      s_n2 = 0;
      s_n3 = 0;
    }
    __syncthreads();
    // ---- (A) classify; tier 2 re-scores its candidates into routes A / B; tiers 1 and 2 fold their kept deltas ------
    {
      const uint32_t f = base + threadIdx.x;
      if (f < c_hi) {
        const uint4 prec = v.pr[f];
        const uint32_t x = prec.x, se = prec.w;
        int tier = 3;
        uint32_t hits = 0;  // lanes whose slot sits in route A or B
        if (valid) {
          tier = 1;
          if (moved) {
            const uint4 m0 = a.c_meta[(src0 + x) * 2], m1 = a.c_meta[(src0 + x) * 2 + 1];
            const CELL* __restrict__ mrow = fm_row + (size_t)x * nc.dim;
            bool near = false;
#pragma unroll
            for (int h = 0; h < NBC_N_HOT; ++h)
              near |= hot[h] != 0xFFFFFFFFu && (uint32_t)__ldg(mrow + hot[h]) <= m0.x;
            const uint32_t codes[7] = {m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
            hits = nbc_route_hits(codes, tA & 255u, tB & 255u, count);
            if (se == tA || se == tB || near) tier = 3;
            else if (hits) tier = 2;
          }
        }
        int2* cd = a.c_delta + (src0 + x) * K;
        if (tier == 2) {
          const uint32_t* __restrict__ ci = a.c_ident + (src0 + x) * K;
          const uint4 rsrc = v.rr[se];
          const CELL* __restrict__ mrow = fm_row + (size_t)x * nc.dim;
          while (hits) {
            const uint32_t j = __ffs(hits) - 1;
            hits &= hits - 1;
            const uint32_t id = ci[j];
            if (id == 0xFFFFFFFFu) continue;
            const uint32_t y = id & 0x7FFFu;
            const uint32_t where = v.pos_of[y];
            const uint32_t e = where >> 16, py = where & 0xFFFFu;
            S dh, ds;
            uint32_t de, dp, ident;
            nearby_score_slot<SUM_FN, CELL, S>(m, nc, v, v.rr[e].x + e + py + ((id >> 15) & 1), (uint32_t)__ldg(mrow + y), x,
                                               se, prec, rsrc, dh, ds, de, dp, ident);
            cd[j] = make_int2(dh, ds);
          }
          atomicAdd(&s_n2, 1u);
        }
        if (tier != 3) nbc_fold_source<KC>(cd, count, a.f.acceptor, lh, ls, th, ts, ch, csf, P + f);
        else s_list3[atomicAdd(&s_n3, 1u)] = (uint16_t)threadIdx.x;  // appended to the replica's work list below
      }
    }
    __syncthreads();
    const uint32_t n2 = s_n2, n3 = s_n3;
    if (threadIdx.x == 0) {  // running totals of the three tiers (test hook / profiles)
      uint32_t* tg = m.nbc_tag + (size_t)r * NBC_WORDS;
      atomicAdd(tg + NBC_STATS, min(c_hi - base, 256u) - n2 - n3);
      atomicAdd(tg + NBC_STATS + 1, n2);
      atomicAdd(tg + NBC_STATS + 2, n3);
    }
    // tier-3 sources of this tile -> the replica's work list (nearby_regen_kernel)
    if (n3) {
      __shared__ uint32_t s_at;
      if (threadIdx.x == 0) s_at = atomicAdd(a.c_work + (size_t)r * (m.elem_cap + 1), n3);
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < n3; i += 256) a.c_work[(size_t)r * (m.elem_cap + 1) + 1 + s_at + i] = base + s_list3[i];
    }
    if (base + 256 < c_hi) __syncthreads();  // the next tile reuses the work list
  }
}

// (B) the tier-3 sources of every replica, generated and scored one warp per source as nearby_step_kernel does; the
// work list of replica r is c_work[r] = {n, flat source positions...}. grid = (G, R): the warps of the G CTAs of a
// replica stride over its list; CTAs without work leave before staging anything.
template <int SUM_FN>
__global__ void __launch_bounds__(256) nearby_regen_kernel(const __grid_constant__ DevModel m, const NearbyArgs a) {
  typedef uint32_t KEY;
  typedef uint16_t CELL;
  typedef int32_t S;
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  __shared__ KEY s_buf[8][64];
  const uint32_t r = blockIdx.y;
  const uint32_t* __restrict__ work = a.c_work + (size_t)r * (m.elem_cap + 1);
  const uint32_t n3 = work[0];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // three or more sources per warp before another CTA (and its 38 KB of staging) joins in
  const uint32_t active = min(max((n3 + 23) / 24, 1u), gridDim.x);
  if (n3 == 0 || blockIdx.x >= active) return;
  int64_t refs[4] = {0, 0, 0, 0};
  if (a.ref_scores) {
#pragma unroll
    for (int q = 0; q < 4; ++q) refs[q] = a.ref_scores[r * 4 + q];
  }
  stage_block(smem, m.state + (size_t)r * m.block_bytes, m.fast_stage_bytes, &bar);
  NearbyView v;
  v.rr = (const uint4*)(smem + m.off_route_rec);
  v.pr = (const uint4*)(smem + m.off_pos_rec);
  v.sr = (const uint4*)(smem + m.off_slot_rec);
  v.pos_of = (const uint32_t*)(smem + m.off_pos_of);
  const int64_t* cs = (const int64_t*)(smem + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  const uint32_t count = nearby_count_warp(m, v.rr, a.max_nearby, lane), K = a.max_nearby;
  S lh, ls, th, ts;
  rel_threshold(refs[0], ch, lh);
  rel_threshold(refs[1], csf, ls);
  rel_threshold(refs[2], ch, th);
  rel_threshold(refs[3], csf, ts);
  const NearbyConsts<S> nc = nearby_consts<SUM_FN, S>(m);
  const size_t src0 = (size_t)r * m.n_elem_rows;
  SrcPartial* __restrict__ P = a.partials + (size_t)r * m.elem_cap;
  for (uint32_t i = blockIdx.x * 8 + warp; i < n3; i += active * 8) {
    const uint32_t f = work[1 + i];
    uint32_t x, se, sp;
    uint4 prec, rsrc;
    const KEY key = nearby_gen_source<KEY, CELL>(m, v, a.scan_bits, f, K, lane, s_buf[warp], x, se, sp, prec, rsrc);
    const bool have = lane < count && key != 0xFFFFFFFFu;
    S dh = INT32_MIN, ds = 0;
    uint32_t de = 0, dp = 0, ident = 0xFFFFFFFFu;
    if (have) {
      nearby_score_slot<SUM_FN, CELL, S>(m, nc, v, nearby_slot_of_scan(key & ((1u << a.scan_bits) - 1), se, rsrc),
                                         key >> a.scan_bits, x, se, prec, rsrc, dh, ds, de, dp, ident);
      ident = nbc_ident(ident, de);
    }
    if (lane < K) {
      a.c_delta[(src0 + x) * K + lane] = make_int2(dh, ds);
      a.c_ident[(src0 + x) * K + lane] = ident;
    }
    // meta: k-th kept distance + the 8-bit route code of every lane (4 lanes per word)
    uint32_t cw = have ? (de & 255u) << (8 * (lane & 3)) : 0u;
    cw |= __shfl_xor_sync(0xffffffffu, cw, 1);
    cw |= __shfl_xor_sync(0xffffffffu, cw, 2);
    uint32_t codes[7];
#pragma unroll
    for (uint32_t w = 0; w < 7; ++w) codes[w] = __shfl_sync(0xffffffffu, cw, 4 * w);
    const uint32_t kth = __shfl_sync(0xffffffffu, key, K - 1);
    if (lane == 0) {
      // a list that is not full (count < K, or fewer slots than K) holds every slot: any change matters
      a.c_meta[(src0 + x) * 2] = make_uint4((count < K || kth == 0xFFFFFFFFu) ? 0xFFFFFFFFu : kth >> a.scan_bits, codes[0],
                                            codes[1], codes[2]);
      a.c_meta[(src0 + x) * 2 + 1] = make_uint4(codes[3], codes[4], codes[5], codes[6]);
    }
    const bool acc = have && accept_delta<S>(a.f.acceptor, dh, ds, lh, ls, th, ts);
    uint32_t accm;
    const uint32_t eq = warp_best_mask(acc, dh, ds, accm);
    if (lane == (eq ? __ffs(eq) - 1 : 0)) {
      SrcPartial p;
      p.best_h = eq ? ch + (int64_t)dh : 0;
      p.best_s = eq ? csf + (int64_t)ds : 0;
      p.n_best = __popc(eq);
      p.n_accepted = __popc(accm);
      p.first_lane = eq ? __ffs(eq) - 1 : 0;
      p.pad = 0;
      P[f] = p;
    }
  }
}

// the lanes of source f from the retained rows of the cached step (instead of regenerating the source): delta, and for
// the winner its destination through the kept reference element. Only valid right behind nearby_step_cached_kernel.
template <typename S>
__device__ __forceinline__ bool nbc_source_lane(const DevModel& m, const NearbyArgs& a, const NearbyView& v, const uint32_t r,
                                                const uint32_t f, const uint32_t lane, const uint32_t count, uint32_t& se,
                                                uint32_t& sp, S& dh, S& ds, uint32_t& id) {
  const uint4 prec = v.pr[f];
  se = prec.w;
  sp = f - v.rr[se].x;
  dh = 0;
  ds = 0;
  id = 0xFFFFFFFFu;
  if (lane >= a.max_nearby) return false;
  const size_t at = ((size_t)r * m.n_elem_rows + prec.x) * a.max_nearby + lane;
  const int2 d = a.c_delta[at];
  id = a.c_ident[at];
  dh = (S)d.x;
  ds = (S)d.y;
  return lane < count && d.x != INT32_MIN;
}
__device__ __forceinline__ void nbc_destination(const NearbyView& v, const uint32_t id, uint32_t& de, uint32_t& dp) {
  const uint32_t where = v.pos_of[id & 0x7FFFu];
  de = where >> 16;
  dp = (where & 0xFFFFu) + ((id >> 15) & 1u);
}

// One CTA per replica: ordered replay over the per-source partials (AcceptedCount cut, best, tie
// rule), regeneration of the one or two sources whose lanes matter, winner row out.
template <int SUM_FN, typename KEY, typename CELL, int MOVE = MOVE_CHANGE>
__global__ void __launch_bounds__(256) nearby_finish_kernel(const __grid_constant__ DevModel m, const NearbyArgs a,
                                                            uint32_t* __restrict__ out_index,
                                                            int64_t* __restrict__ out_best,
                                                            uint32_t* __restrict__ out_evaluated,
                                                            uint32_t* __restrict__ out_winner_rows) {
  __shared__ uint32_t scratch[33];
  __shared__ int64_t s_h[8], s_s[8];
  __shared__ uint32_t s_flag[8];
  __shared__ uint32_t s_fcut, s_lcut, s_evaluated, s_any, s_fstar, s_jstar, s_count;
  __shared__ int64_t s_bh, s_bs;
  __shared__ SrcPartial s_cutp;  // truncated partial of the cut source
  __shared__ KEY s_buf[64];
  const uint32_t r = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const char* st = m.state + (size_t)r * m.block_bytes;
  NearbyView v;
  v.rr = (const uint4*)(st + m.off_route_rec);
  v.pr = (const uint4*)(st + m.off_pos_rec);
  v.sr = (const uint4*)(st + m.off_slot_rec);
  v.pos_of = (const uint32_t*)(st + m.off_pos_of);
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  const uint32_t total = v.rr[m.n_owners - 1].x + v.rr[m.n_owners - 1].y;
  const SrcPartial* P = a.partials + (size_t)r * m.elem_cap;
  typedef typename FastArith<CELL>::S S;
  S lh, ls, th, ts;
  rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 0] : 0, ch, lh);
  rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 1] : 0, csf, ls);
  rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 2] : 0, ch, th);
  rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 3] : 0, csf, ts);
  const NearbyConsts<S> nc = nearby_consts<SUM_FN, S>(m);
  if (warp == 0) {  // the route lengths sit in global memory here: one load per lane instead of a serial walk
    const uint32_t c = nearby_count_warp(m, v.rr, a.max_nearby, lane);
    if (lane == 0) s_count = c;
  }
  if (threadIdx.x == 0) {
    s_fcut = total;       // sources [0, s_fcut) count fully; source s_fcut only up to lane s_lcut
    s_lcut = 0;
    s_cutp.n_best = 0;
    s_cutp.n_accepted = 0;
  }
  __syncthreads();
  const uint32_t count = s_count;
  if (a.cache && threadIdx.x == 0) {  // the cached step kernel of this launch pair brought the retained deltas up to date
    uint32_t* tag = m.nbc_tag + (size_t)r * NBC_WORDS;
    tag[NBC_K] = a.max_nearby;
    tag[NBC_COUNT] = count;
    tag[NBC_STATE] = 1;
    a.c_work[(size_t)r * (m.elem_cap + 1)] = 0;  // the regeneration work list of the next step starts empty
  }
  // ---- AcceptedCount(N): the step ends right after the N-th accepted pull -------------------
  if (a.f.accepted_limit > 0) {
    uint32_t seen = 0;
    bool found = false;
    for (uint32_t base = 0; base < total && !found; base += blockDim.x) {
      const uint32_t f = base + threadIdx.x;
      const uint32_t na = f < total ? P[f].n_accepted : 0;
      uint32_t tot;
      const uint32_t incl = block_scan_u32(na, scratch, &tot);
      if (na && seen + incl - na < a.f.accepted_limit && a.f.accepted_limit <= seen + incl) {
        s_fcut = f;
        s_lcut = a.f.accepted_limit - (seen + incl - na);  // rank (1-based) of the cutting accepted lane
      }
      seen += tot;
      __syncthreads();
      found = seen >= a.f.accepted_limit;
    }
    __syncthreads();
    if (s_fcut < total && warp == 0) {
      // regenerate the cut source; keep lanes up to the s_lcut-th accepted one
      uint32_t x, se, sp;
      uint4 prec, rsrc;
      S dh = 0, ds = 0;
      uint32_t de = 0, dp = 0;
      bool have;
      if (a.cache) {  // the cached step kernel holds this source's deltas
        uint32_t id;
        have = nbc_source_lane<S>(m, a, v, r, s_fcut, lane, count, se, sp, dh, ds, id);
      } else {
        KEY key;
        if (MOVE == MOVE_SWAP)
          key = nearby_swap_gen_source<KEY, CELL>(m, v, a.scan_bits, s_fcut, a.max_nearby, total, lane, s_buf, x, se, sp,
                                                  prec, rsrc);
        else
          key = nearby_gen_source<KEY, CELL>(m, v, a.scan_bits, s_fcut, a.max_nearby, lane, s_buf, x, se, sp, prec, rsrc);
        const uint32_t cnt = MOVE == MOVE_SWAP ? swap_count(s_fcut, total, a.max_nearby) : count;
        have = lane < cnt && key != KeyTraits<KEY>::maxkey();
        if (have) {
          if (MOVE == MOVE_SWAP)
            nearby_swap_score<SUM_FN, KEY, CELL, S>(m, nc, v, a.scan_bits, key, s_fcut, x, se, prec, rsrc, dh, ds, de, dp);
          else
            nearby_score<SUM_FN, KEY, CELL, S>(m, nc, v, a.scan_bits, key, x, se, prec, rsrc, dh, ds, de, dp);
        }
      }
      const int64_t oh = ch + (int64_t)dh, os = csf + (int64_t)ds;
      bool acc = have && accept_delta<S>(a.f.acceptor, dh, ds, lh, ls, th, ts);
      const uint32_t cut_rank = s_lcut;  // read by every lane before one lane overwrites it below
      __syncwarp();
      uint32_t mm = __ballot_sync(0xffffffffu, acc);
      for (uint32_t t = 1; t < cut_rank; ++t) mm &= mm - 1;
      const uint32_t cut_lane = __ffs(mm) - 1;
      acc = acc && lane <= cut_lane;
      uint32_t accm;
      const uint32_t eq = warp_best_mask(acc, dh, ds, accm);
      if (lane == (eq ? __ffs(eq) - 1 : 0)) {
        s_cutp.best_h = oh;
        s_cutp.best_s = os;
        s_cutp.n_best = __popc(eq);
        s_cutp.n_accepted = cut_rank;
        s_cutp.first_lane = eq ? __ffs(eq) - 1 : 0;
        s_lcut = cut_lane;  // from here on: last counted lane of the cut source
      }
    }
    __syncthreads();
  }
  const uint32_t fcut = s_fcut;
  const uint32_t n_src = fcut < total ? fcut + 1 : total;  // sources that contribute
  // ---- best accepted score ---------------------------------------------------------------------
  int64_t bh = 0, bs = 0;
  uint32_t any = 0;
  for (uint32_t f = threadIdx.x; f < n_src; f += blockDim.x) {
    const SrcPartial p = (f == fcut) ? s_cutp : P[f];
    if (!p.n_best) continue;
    if (!any || score_less(bh, bs, p.best_h, p.best_s)) {
      bh = p.best_h;
      bs = p.best_s;
    }
    any = 1;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t ph = __shfl_xor_sync(0xffffffffu, bh, o), ps = __shfl_xor_sync(0xffffffffu, bs, o);
    const uint32_t pa = __shfl_xor_sync(0xffffffffu, any, o);
    if (pa && (!any || score_less(bh, bs, ph, ps))) {
      bh = ph;
      bs = ps;
    }
    any |= pa;
  }
  if (lane == 0) {
    s_h[warp] = bh;
    s_s[warp] = bs;
    s_flag[warp] = any;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t h = 0, s2 = 0;
    uint32_t an = 0;
    for (int w = 0; w < 8; ++w)
      if (s_flag[w] && (!an || score_less(h, s2, s_h[w], s_s[w]))) {
        h = s_h[w];
        s2 = s_s[w];
        an = 1;
      }
    s_bh = h;
    s_bs = s2;
    s_any = an;
    if (MOVE == MOVE_SWAP)
      s_evaluated = fcut < total ? swap_prefix(fcut, total, a.max_nearby) + s_lcut + 1 : swap_prefix(total, total, a.max_nearby);
    else
      s_evaluated = fcut < total ? fcut * count + s_lcut + 1 : total * count;
  }
  __syncthreads();
  bh = s_bh;
  bs = s_bs;
  if (!s_any) {
    if (threadIdx.x == 0) {
      out_index[r] = 0xFFFFFFFFu;
      out_best[r * 2] = 0;
      out_best[r * 2 + 1] = 0;
      if (out_evaluated) out_evaluated[r] = s_evaluated;
      if (out_winner_rows) ((uint4*)out_winner_rows)[r] = make_uint4(0xFFFFFFFFu, 0, 0xFFFFFFFFu, 0);
    }
    return;
  }
  // ---- multiplicity of the best and the tie rule -----------------------------------------------
  uint32_t mloc = 0;
  for (uint32_t f = threadIdx.x; f < n_src; f += blockDim.x) {
    const SrcPartial p = (f == fcut) ? s_cutp : P[f];
    if (p.n_best && p.best_h == bh && p.best_s == bs) mloc += p.n_best;
  }
  uint32_t mtot;
  block_scan_u32(mloc, scratch, &mtot);
  if (threadIdx.x == 0) s_jstar = 1;
  __syncthreads();
  if (a.f.tie_mode == 1) {
    const uint64_t seed = a.step_seeds ? a.step_seeds[r] : 0;
    uint32_t best_k = 1;
    for (uint32_t k = 2 + threadIdx.x; k <= mtot; k += blockDim.x) {
      const uint64_t mixed = splitmix64_dev(seed ^ ((uint64_t)k * 0x9E3779B97F4A7C15ull) ^ 0xF04A63E239B74D11ull);
      if (mixed % k == 0) best_k = k;
    }
    atomicMax(&s_jstar, best_k);
    __syncthreads();
  }
  const uint32_t want = s_jstar;
  __syncthreads();
  // ---- source holding the want-th best occurrence (sources are in pull order) -------------------
  {
    uint32_t seen = 0;
    bool found = false;
    for (uint32_t base = 0; base < n_src && !found; base += blockDim.x) {
      const uint32_t f = base + threadIdx.x;
      uint32_t nb = 0;
      if (f < n_src) {
        const SrcPartial p = (f == fcut) ? s_cutp : P[f];
        nb = (p.n_best && p.best_h == bh && p.best_s == bs) ? p.n_best : 0;
      }
      uint32_t tot;
      const uint32_t incl = block_scan_u32(nb, scratch, &tot);
      if (nb && seen + incl - nb < want && want <= seen + incl) {
        s_fstar = f;
        s_jstar = want - (seen + incl - nb);  // rank inside the source
      }
      seen += tot;
      __syncthreads();
      found = seen >= want;
    }
  }
  __syncthreads();
  if (warp == 0) {
    const uint32_t fstar = s_fstar, j = s_jstar;
    uint32_t x, se, sp;
    uint4 prec, rsrc;
    S dh = 0, ds = 0;
    uint32_t de = 0, dp = 0, id = 0xFFFFFFFFu;
    bool have;
    if (a.cache) {
      have = nbc_source_lane<S>(m, a, v, r, fstar, lane, count, se, sp, dh, ds, id) && !(fstar == fcut && lane > s_lcut);
    } else {
      KEY key;
      if (MOVE == MOVE_SWAP)
        key = nearby_swap_gen_source<KEY, CELL>(m, v, a.scan_bits, fstar, a.max_nearby, total, lane, s_buf, x, se, sp, prec,
                                                rsrc);
      else
        key = nearby_gen_source<KEY, CELL>(m, v, a.scan_bits, fstar, a.max_nearby, lane, s_buf, x, se, sp, prec, rsrc);
      const uint32_t cnt = MOVE == MOVE_SWAP ? swap_count(fstar, total, a.max_nearby) : count;
      have = lane < cnt && key != KeyTraits<KEY>::maxkey() && !(fstar == fcut && lane > s_lcut);
      if (have) {
        if (MOVE == MOVE_SWAP)
          nearby_swap_score<SUM_FN, KEY, CELL, S>(m, nc, v, a.scan_bits, key, fstar, x, se, prec, rsrc, dh, ds, de, dp);
        else
          nearby_score<SUM_FN, KEY, CELL, S>(m, nc, v, a.scan_bits, key, x, se, prec, rsrc, dh, ds, de, dp);
      }
    }
    const int64_t oh = ch + (int64_t)dh, os = csf + (int64_t)ds;
    const bool hit = have && oh == bh && os == bs && accept_delta<S>(a.f.acceptor, dh, ds, lh, ls, th, ts);
    uint32_t mm = __ballot_sync(0xffffffffu, hit);
    for (uint32_t t = 1; t < j; ++t) mm &= mm - 1;
    const uint32_t wl = __ffs(mm) - 1;
    if (lane == wl) {
      if (a.cache) nbc_destination(v, id, de, dp);
      out_index[r] = (MOVE == MOVE_SWAP ? swap_prefix(fstar, total, a.max_nearby) : fstar * count) + wl;
      out_best[r * 2] = bh;
      out_best[r * 2 + 1] = bs;
      if (out_evaluated) out_evaluated[r] = s_evaluated;
      if (out_winner_rows) ((uint4*)out_winner_rows)[r] = make_uint4(se, sp, de, dp);
    }
  }
}
