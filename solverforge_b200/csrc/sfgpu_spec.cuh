// Program-specialised ChangeMove scoring for scalar models.
//
// The generic score_scalar_kernel interprets the constraint table per candidate (a switch per
// constraint, parameters re-read from the kernel-parameter bank, ~250 instructions and ~30 branches
// per candidate); measurements show that interpreter overhead, not memory, bounds it.
// The reference gets the same effect from monomorphised ConstraintSet tuples
// (solverforge-scoring/src/api/constraint_set/incremental.rs:339-408). Here the kernel is a template
// over the (sorted) kinds of the scalar constraints of the model: every parameter is hoisted into
// registers before the candidate loop, there is no switch, and the four candidates of a trip are
// straight-line code, so all their table gathers are in flight together. Instantiated for the
// constraint tuples of the reference's scalar examples; other programs keep the interpreter.
//
// Deltas are the same closed forms as scalar_edit_delta with n_prev == 0 (ChangeMove).
#pragma once
#include "sfgpu_kernels.cuh"

template <int KIND>
struct SpecCons;

// impact sign and score level of one constraint (constraint/incremental.rs:70-85), held in registers
struct SpecRoute {
  bool neg, hard;
  __device__ __forceinline__ void route_init(const ConsDev& c) {
    neg = c.sign < 0;
    hard = c.w.level == 0;
  }
  // constraints whose delta is linear in the weight fold the impact sign into the weight once
  __device__ __forceinline__ int64_t signed_weight(int64_t w) const { return neg ? -w : w; }
  __device__ __forceinline__ int64_t apply_sign(int64_t v) const { return neg ? -v : v; }
  // v already carries the impact sign
  __device__ __forceinline__ void add(int64_t& dh, int64_t& ds, int64_t v) const {
    if (hard) dh += v; else ds += v;
  }
};

template <>
struct SpecCons<0> {  // empty slot
  __device__ __forceinline__ SpecCons(const DevModel&, int, const char*, const char*) {}
  __device__ __forceinline__ int64_t delta(uint32_t, int32_t, int32_t) const { return 0; }
  __device__ __forceinline__ void add(int64_t&, int64_t&, int64_t) const {}
};

template <>
struct SpecCons<SFGPU_K_UNI> : SpecRoute {  // constraint/incremental.rs:97-156
  const int64_t* col;
  const int64_t* mask;
  WeightDev w;
  int32_t filt;
  bool by_value;
  int64_t w0;  // weight when there is no column
  __device__ __forceinline__ SpecCons(const DevModel& m, int k, const char*, const char*) {
    const ConsDev& c = m.cons[k];
    col = (const int64_t*)c.g0;
    mask = (const int64_t*)c.g1;
    w = c.w;
    filt = (int32_t)c.p0;
    by_value = (c.flags & SFGPU_CF_COL_BY_VALUE) != 0;
    w0 = weight_eval(c.w, 0);
    route_init(c);
  }
  __device__ __forceinline__ int64_t contrib(uint32_t e, int32_t v, bool masked_out) const {
    const bool pass = (filt == 0 ? v < 0 : (filt == 1 ? v >= 0 : true)) && !masked_out;
    if (!col) return pass ? w0 : 0;
    const int64_t x = by_value ? (v >= 0 ? col[v] : 0) : col[e];
    return pass ? weight_eval(w, x) : 0;
  }
  __device__ __forceinline__ int64_t delta(uint32_t e, int32_t ov, int32_t nv) const {
    const bool out = mask && mask[e] == 0;
    return apply_sign(contrib(e, nv, out) - contrib(e, ov, out));
  }
};

// pseudo-kind (not part of the ABI): a uni constraint without weight column and without entity mask —
// `for_each(E).unassigned().penalize(ONE_HARD)` of every scalar example. The weight is one register.
#define SPEC_K_UNI_CONST 33
template <>
struct SpecCons<SPEC_K_UNI_CONST> : SpecRoute {
  int32_t filt;
  int64_t w0;
  __device__ __forceinline__ SpecCons(const DevModel& m, int k, const char*, const char*) {
    const ConsDev& c = m.cons[k];
    filt = (int32_t)c.p0;
    route_init(c);
    w0 = signed_weight(weight_eval(c.w, 0));
  }
  __device__ __forceinline__ int64_t delta(uint32_t, int32_t ov, int32_t nv) const {
    // pass(v): filt 0 = unassigned, 1 = assigned, 2 = always
    const int32_t pn = filt == 2 ? 1 : ((nv < 0) == (filt == 0) ? 1 : 0);
    const int32_t po = filt == 2 ? 1 : ((ov < 0) == (filt == 0) ? 1 : 0);
    return (int64_t)(pn - po) * w0;
  }
};

template <>
struct SpecCons<SFGPU_K_PAIR_CSR_EQUAL> : SpecRoute {  // retained partner-value counts cc[e][v] (global, unstaged)
  const uint16_t* cc;
  uint32_t k;
  int64_t a;
  __device__ __forceinline__ SpecCons(const DevModel& m, int idx, const char*, const char* gblock) {
    const ConsDev& c = m.cons[idx];
    cc = (const uint16_t*)(gblock + c.off0);
    k = m.n_values;
    route_init(c);
    a = signed_weight(c.w.a);
  }
  __device__ __forceinline__ int64_t delta(uint32_t e, int32_t ov, int32_t nv) const {
    const uint16_t* row = cc + (size_t)e * k;
    const int32_t cn = nv >= 0 ? (int32_t)row[nv] : 0, co = ov >= 0 ? (int32_t)row[ov] : 0;
    return (int64_t)(cn - co) * a;
  }
};

template <>
struct SpecCons<SFGPU_K_PAIR_KEY_EQUAL> : SpecRoute {  // keyed self-join, per-key counts (staged)
  const int32_t* tab;
  const int64_t* col;
  uint32_t p0, p1, p2;  // key arithmetic modulo 2^32: the final index lies in [0, table length)
  int64_t a;
  __device__ __forceinline__ SpecCons(const DevModel& m, int idx, const char* st, const char*) {
    const ConsDev& c = m.cons[idx];
    tab = (const int32_t*)(st + c.off0);
    col = (const int64_t*)c.g0;
    p0 = (uint32_t)c.p0;
    p1 = (uint32_t)c.p1;
    p2 = (uint32_t)c.p2;
    route_init(c);
    a = signed_weight(c.w.a);
  }
  __device__ __forceinline__ int64_t delta(uint32_t e, int32_t ov, int32_t nv) const {
    const uint32_t base = (col ? (uint32_t)col[e] : 0u) * p0 - p2;
    const int32_t cn = nv >= 0 ? tab[base + (uint32_t)nv * p1] : 0;
    const int32_t co = ov >= 0 ? tab[base + (uint32_t)ov * p1] - 1 : 0;
    return (int64_t)(cn - co) * a;
  }
};

template <>
struct SpecCons<SFGPU_K_GROUP> : SpecRoute {  // grouped count / sum, optional complement and per-key offset (staged)
  const int32_t* gc;
  const int64_t* gs;
  const int64_t* col;
  const int64_t* key_off;
  WeightDev w;
  bool complement;
  bool square_uniform;  // a*x*x + b with the empty group scoring exactly w(0): the delta has a closed form
  int64_t dflt, empty0;
  __device__ __forceinline__ SpecCons(const DevModel& m, int idx, const char* st, const char*) {
    const ConsDev& c = m.cons[idx];
    gc = (const int32_t*)(st + c.off0);
    gs = (const int64_t*)(st + c.off1);
    col = (const int64_t*)c.g0;
    key_off = (const int64_t*)c.g1;
    w = c.w;
    complement = (c.flags & SFGPU_CF_COMPLEMENT) != 0;
    dflt = c.p1;
    empty0 = complement ? weight_eval(c.w, c.p1) : 0;
    square_uniform = c.w.fn == SFGPU_W_SQUARE && !key_off && empty0 == c.w.b;
    route_init(c);
  }
  template <int FN>
  __device__ __forceinline__ int64_t wfn(int64_t wb, int64_t x) const {
    WeightDev ww;
    ww.fn = FN;
    ww.a = w.a;
    ww.b = wb;
    return weight_eval(ww, x);
  }
  // score of an empty group: the complement's default result, or nothing (grouped/state.rs:349-365)
  template <int FN>
  __device__ __forceinline__ int64_t empty(int64_t wb) const {
    return key_off ? (complement ? wfn<FN>(wb, dflt) : 0) : empty0;
  }
  // e sits in group `ov`, so that group has count >= 1 before the edit; group `nv` has count >= 1 after it
  template <int FN>
  __device__ __forceinline__ int64_t delta_fn(uint32_t e, int32_t ov, int32_t nv) const {
    const int64_t x = col ? col[e] : 1;
    int64_t v = 0;
    if (ov >= 0) {
      const int64_t cn = gc[ov], sm = col ? gs[ov] : cn, wb = key_off ? key_off[ov] : w.b;
      v += (cn > 1 ? wfn<FN>(wb, sm - x) : empty<FN>(wb)) - wfn<FN>(wb, sm);
    }
    if (nv >= 0) {
      const int64_t cn = gc[nv], sm = col ? gs[nv] : cn, wb = key_off ? key_off[nv] : w.b;
      v += wfn<FN>(wb, sm + x) - (cn > 0 ? wfn<FN>(wb, sm) : empty<FN>(wb));
    }
    return v;
  }
  // one (warp-uniform) dispatch on the weight function per candidate instead of one per evaluation
  __device__ __forceinline__ int64_t delta(uint32_t e, int32_t ov, int32_t nv) const {
    if (square_uniform) {
      // every group, empty or not, scores a*s*s + b (s = its count or sum, 0 when empty), so moving x from
      // group o to group n changes the total by a*((so - x)^2 - so^2 + (sn + x)^2 - sn^2)
      const int64_t x = col ? col[e] : 1;
      int64_t v = 0;
      if (ov >= 0) {
        const int64_t so = col ? gs[ov] : (int64_t)gc[ov];
        v += x * (x - 2 * so);
      }
      if (nv >= 0) {
        const int64_t sn = col ? gs[nv] : (int64_t)gc[nv];
        v += x * (x + 2 * sn);
      }
      return apply_sign(w.a * v);
    }
    switch (w.fn) {
      case SFGPU_W_CONST: return apply_sign(delta_fn<SFGPU_W_CONST>(e, ov, nv));
      case SFGPU_W_LINEAR: return apply_sign(delta_fn<SFGPU_W_LINEAR>(e, ov, nv));
      case SFGPU_W_SQUARE: return apply_sign(delta_fn<SFGPU_W_SQUARE>(e, ov, nv));
      case SFGPU_W_ABSDIFF: return apply_sign(delta_fn<SFGPU_W_ABSDIFF>(e, ov, nv));
      case SFGPU_W_PAIRS: return apply_sign(delta_fn<SFGPU_W_PAIRS>(e, ov, nv));
      default: return apply_sign(delta_fn<SFGPU_W_EXCESS>(e, ov, nv));
    }
  }
};

// ---- narrow (int32) forms of the kinds that have one --------------------------------------------------------
// Selected at commit when every delta of a ChangeMove provably fits int32 with headroom (sfgpu_scalar.cu,
// scalar_narrow_bound): table cells, column values (ConsDev::g2), weights and every compared quantity are then
// below 2^30 in magnitude, sums / products go through uint32 (two's-complement ring, exact whenever the result
// fits) and the 64-bit committed score is only added for the store. Same closed forms as the wide structs.
template <int KIND>
struct SpecConsN;

struct SpecRouteN {
  bool hard;
  __device__ __forceinline__ void add(int32_t& dh, int32_t& ds, int32_t v) const {
    if (hard) dh += v; else ds += v;
  }
};

template <>
struct SpecConsN<0> {
  __device__ __forceinline__ SpecConsN(const DevModel&, int, const char*, const char*) {}
  __device__ __forceinline__ int32_t delta(uint32_t, int32_t, int32_t) const { return 0; }
  __device__ __forceinline__ void add(int32_t&, int32_t&, int32_t) const {}
};

template <>
struct SpecConsN<SPEC_K_UNI_CONST> : SpecRouteN {
  int32_t filt, w0;
  __device__ __forceinline__ SpecConsN(const DevModel& m, int k, const char*, const char*) {
    const ConsDev& c = m.cons[k];
    filt = (int32_t)c.p0;
    hard = c.w.level == 0;
    const int64_t w = weight_eval(c.w, 0);
    w0 = (int32_t)(c.sign < 0 ? -w : w);
  }
  __device__ __forceinline__ int32_t delta(uint32_t, int32_t ov, int32_t nv) const {
    const int32_t pn = filt == 2 ? 1 : ((nv < 0) == (filt == 0) ? 1 : 0);
    const int32_t po = filt == 2 ? 1 : ((ov < 0) == (filt == 0) ? 1 : 0);
    return (pn - po) * w0;
  }
};

template <>
struct SpecConsN<SFGPU_K_PAIR_CSR_EQUAL> : SpecRouteN {
  const uint16_t* cc;
  uint32_t k;
  int32_t a;
  __device__ __forceinline__ SpecConsN(const DevModel& m, int idx, const char*, const char* gblock) {
    const ConsDev& c = m.cons[idx];
    cc = (const uint16_t*)(gblock + c.off0);
    k = m.n_values;
    hard = c.w.level == 0;
    a = (int32_t)(c.sign < 0 ? -c.w.a : c.w.a);
  }
  __device__ __forceinline__ int32_t delta(uint32_t e, int32_t ov, int32_t nv) const {
    const uint16_t* row = cc + (size_t)e * k;
    const int32_t cn = nv >= 0 ? (int32_t)row[nv] : 0, co = ov >= 0 ? (int32_t)row[ov] : 0;
    return (cn - co) * a;
  }
};

template <>
struct SpecConsN<SFGPU_K_PAIR_KEY_EQUAL> : SpecRouteN {
  const int32_t* tab;
  const int32_t* col;
  uint32_t p0, p1, p2;
  int32_t a;
  __device__ __forceinline__ SpecConsN(const DevModel& m, int idx, const char* st, const char*) {
    const ConsDev& c = m.cons[idx];
    tab = (const int32_t*)(st + c.off0);
    col = (const int32_t*)c.g2;
    p0 = (uint32_t)c.p0;
    p1 = (uint32_t)c.p1;
    p2 = (uint32_t)c.p2;
    hard = c.w.level == 0;
    a = (int32_t)(c.sign < 0 ? -c.w.a : c.w.a);
  }
  __device__ __forceinline__ int32_t delta(uint32_t e, int32_t ov, int32_t nv) const {
    const uint32_t base = (col ? (uint32_t)__ldg(col + e) : 0u) * p0 - p2;
    const int32_t cn = nv >= 0 ? tab[base + (uint32_t)nv * p1] : 0;
    const int32_t co = ov >= 0 ? tab[base + (uint32_t)ov * p1] - 1 : 0;
    return (cn - co) * a;
  }
};

template <>
struct SpecConsN<SFGPU_K_GROUP> : SpecRouteN {  // count() / sum(col), optional complement; no per-key offsets
  const int32_t* gc;
  const int32_t* gs_lo;  // low words of the int64 sums (they fit: narrow bound)
  const int32_t* col;
  int32_t fn, a, b, dflt, empty0;
  bool neg, complement, counting, square_uniform;
  __device__ __forceinline__ int32_t w32(int32_t x) const {
    const uint32_t ua = (uint32_t)a, ux = (uint32_t)x;
    switch (fn) {
      case SFGPU_W_CONST: return a;
      case SFGPU_W_LINEAR: return (int32_t)(ua * ux + (uint32_t)b);
      case SFGPU_W_SQUARE: return (int32_t)(ua * ux * ux + (uint32_t)b);
      case SFGPU_W_ABSDIFF: {
        const int32_t d = x - b;
        return (int32_t)(ua * (uint32_t)(d < 0 ? -d : d));
      }
      default: {
        const int32_t d = x - b;
        return d > 0 ? (int32_t)(ua * (uint32_t)d) : 0;
      }
    }
  }
  __device__ __forceinline__ SpecConsN(const DevModel& m, int idx, const char* st, const char*) {
    const ConsDev& c = m.cons[idx];
    gc = (const int32_t*)(st + c.off0);
    gs_lo = (const int32_t*)(st + c.off1);
    counting = c.g0 == nullptr;
    col = (const int32_t*)c.g2;
    fn = c.w.fn;
    a = (int32_t)c.w.a;
    b = (int32_t)c.w.b;
    complement = (c.flags & SFGPU_CF_COMPLEMENT) != 0;
    dflt = (int32_t)c.p1;
    empty0 = complement ? w32(dflt) : 0;
    square_uniform = c.w.fn == SFGPU_W_SQUARE && empty0 == b;
    neg = c.sign < 0;
    hard = c.w.level == 0;
  }
  __device__ __forceinline__ int32_t delta(uint32_t e, int32_t ov, int32_t nv) const {
    const int32_t x = counting ? 1 : __ldg(col + e);
    int32_t v = 0;
    if (square_uniform) {
      if (ov >= 0) {
        const int32_t so = counting ? gc[ov] : gs_lo[2 * ov];
        v += (int32_t)((uint32_t)x * (uint32_t)(x - 2 * so));
      }
      if (nv >= 0) {
        const int32_t sn = counting ? gc[nv] : gs_lo[2 * nv];
        v += (int32_t)((uint32_t)x * (uint32_t)(x + 2 * sn));
      }
      v = (int32_t)((uint32_t)a * (uint32_t)v);
    } else {
      if (ov >= 0) {
        const int32_t cn = gc[ov], sm = counting ? cn : gs_lo[2 * ov];
        v += (cn > 1 ? w32(sm - x) : empty0) - w32(sm);
      }
      if (nv >= 0) {
        const int32_t cn = gc[nv], sm = counting ? cn : gs_lo[2 * nv];
        v += w32(sm + x) - (cn > 0 ? w32(sm) : empty0);
      }
    }
    return neg ? -v : v;
  }
};

// The scoring program of a model as the kernels see it: delta(e, old, new) of one ChangeMove-shaped edit.
// InterpProg walks the constraint table (any program); SpecProg is the monomorphised tuple; SpecProgN its
// int32 form. S = the type score deltas are computed (and forager partials compared) in.
struct InterpProg {
  typedef int64_t S;
  const DevModel& m;
  const char* st;
  const char* gst;
  __device__ __forceinline__ InterpProg(const DevModel& m_, const SpecIdx&, const char* st_, const char* gst_)
      : m(m_), st(st_), gst(gst_) {}
  __device__ __forceinline__ void delta(uint32_t e, int32_t ov, int32_t nv, int64_t& dh, int64_t& ds) const {
    Score2 d{0, 0};
    scalar_edit_delta(m, st, gst, nullptr, 0, EditDev{e, ov, nv}, d);
    dh = d.hard;
    ds = d.soft;
  }
};

template <int K0, int K1, int K2, int K3>
struct SpecProg {
  typedef int64_t S;
  const SpecCons<K0> c0;
  const SpecCons<K1> c1;
  const SpecCons<K2> c2;
  const SpecCons<K3> c3;
  __device__ __forceinline__ SpecProg(const DevModel& m, const SpecIdx& idx, const char* st, const char* gst)
      : c0(m, idx.k[0], st, gst), c1(m, idx.k[1], st, gst), c2(m, idx.k[2], st, gst), c3(m, idx.k[3], st, gst) {}
  __device__ __forceinline__ void delta(uint32_t e, int32_t ov, int32_t nv, int64_t& dh, int64_t& ds) const {
    dh = 0;
    ds = 0;
    c0.add(dh, ds, c0.delta(e, ov, nv));
    c1.add(dh, ds, c1.delta(e, ov, nv));
    c2.add(dh, ds, c2.delta(e, ov, nv));
    c3.add(dh, ds, c3.delta(e, ov, nv));
  }
};

template <int K0, int K1, int K2, int K3>
struct SpecProgN {
  typedef int32_t S;
  const SpecConsN<K0> c0;
  const SpecConsN<K1> c1;
  const SpecConsN<K2> c2;
  const SpecConsN<K3> c3;
  __device__ __forceinline__ SpecProgN(const DevModel& m, const SpecIdx& idx, const char* st, const char* gst)
      : c0(m, idx.k[0], st, gst), c1(m, idx.k[1], st, gst), c2(m, idx.k[2], st, gst), c3(m, idx.k[3], st, gst) {}
  __device__ __forceinline__ void delta(uint32_t e, int32_t ov, int32_t nv, int32_t& dh, int32_t& ds) const {
    dh = 0;
    ds = 0;
    c0.add(dh, ds, c0.delta(e, ov, nv));
    c1.add(dh, ds, c1.delta(e, ov, nv));
    c2.add(dh, ds, c2.delta(e, ov, nv));
    c3.add(dh, ds, c3.delta(e, ov, nv));
  }
};

#ifndef SFGPU_SPECN_CTAS
#define SFGPU_SPECN_CTAS 4
#endif
// resident CTAs per SM the rows-resident kernel is compiled for: light programs are latency-bound on their table
// gathers and gain from more CTAs; heavy ones keep their registers
template <class PROG>
struct SpecOccupancy {
  static constexpr int min_ctas = 1;
};
template <>
struct SpecOccupancy<SpecProg<SFGPU_K_PAIR_CSR_EQUAL, SPEC_K_UNI_CONST, 0, 0>> {
  static constexpr int min_ctas = 5;
};
template <int K0, int K1, int K2, int K3>
struct SpecOccupancy<SpecProgN<K0, K1, K2, K3>> {
  static constexpr int min_ctas = SFGPU_SPECN_CTAS;
};

// Score deltas as the forager compares them. Wide programs keep the (hard, soft) int64 pair; int32 programs pack
// the pair into ONE int64 key k = hard * 2^32 + soft (a true 64-bit sum): |soft| < 2^31, so numeric order and
// equality of keys are exactly the lexicographic order and equality of the pairs, and the acceptor test and the
// running best cost one 64-bit compare each. Thresholds clamped to int32 (rel_threshold) pack the same way.
template <typename S>
struct DeltaKey;
template <>
struct DeltaKey<int64_t> {
  int64_t h, s;
  __device__ __forceinline__ static DeltaKey make(int64_t dh, int64_t ds) { return DeltaKey{dh, ds}; }
  __device__ __forceinline__ static DeltaKey lowest() { return DeltaKey{INT64_MIN, INT64_MIN}; }
  __device__ __forceinline__ static DeltaKey highest() { return DeltaKey{INT64_MAX, INT64_MAX}; }
  __device__ __forceinline__ bool less(const DeltaKey& o) const { return h != o.h ? h < o.h : s < o.s; }
  __device__ __forceinline__ bool equal(const DeltaKey& o) const { return h == o.h && s == o.s; }
  __device__ __forceinline__ int64_t hard() const { return h; }
  __device__ __forceinline__ int64_t soft() const { return s; }
  __device__ __forceinline__ DeltaKey shfl_down(int o) const {
    return DeltaKey{__shfl_down_sync(0xffffffffu, h, o), __shfl_down_sync(0xffffffffu, s, o)};
  }
};
template <>
struct DeltaKey<int32_t> {
  int64_t k;
  __device__ __forceinline__ static DeltaKey make(int32_t dh, int32_t ds) {
    return DeltaKey{(int64_t)(((uint64_t)(uint32_t)dh << 32) + (uint64_t)(int64_t)ds)};
  }
  __device__ __forceinline__ static DeltaKey lowest() { return DeltaKey{INT64_MIN}; }
  __device__ __forceinline__ static DeltaKey highest() { return DeltaKey{INT64_MAX}; }
  __device__ __forceinline__ bool less(const DeltaKey& o) const { return k < o.k; }
  __device__ __forceinline__ bool equal(const DeltaKey& o) const { return k == o.k; }
  __device__ __forceinline__ int64_t soft() const { return (int64_t)(int32_t)(uint32_t)k; }
  __device__ __forceinline__ int64_t hard() const { return (k - soft()) >> 32; }
  __device__ __forceinline__ DeltaKey shfl_down(int o) const { return DeltaKey{__shfl_down_sync(0xffffffffu, k, o)}; }
};

// Rows-resident ChangeMove scoring with a monomorphised program: grid = (chunks, R), every CTA stages its replica
// block (TMA bulk copy) and scores ONE CONTIGUOUS chunk of the replica's rows (pull order inside a chunk), U rows per
// thread per trip, the next trip's rows prefetched into registers while the current ones are scored. FORAGE: the
// kernel also keeps the forager partial of its chunk (best accepted score, multiplicity, first and second row —
// ChunkPartial, same contract as score_list_change_fast_kernel) so the step never re-reads the scores; out_scores /
// out_doable may then be null.
template <class PROG, bool FORAGE>
__global__ void __launch_bounds__(256, SpecOccupancy<PROG>::min_ctas)
spec_change_kernel(const __grid_constant__ DevModel m, const SpecIdx idx, const uint64_t* __restrict__ cand_offsets,
                   const uint32_t* __restrict__ rows, int64_t* __restrict__ out_scores, uint8_t* __restrict__ out_doable,
                   const ForageArgs fa) {
  typedef typename PROG::S S;
  typedef DeltaKey<S> Key;
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  const uint32_t r = blockIdx.y;
  const char* gblock = m.state + (size_t)r * m.block_bytes;
  // the first rows of the chunk are requested before the replica block is staged: both latencies overlap
  const uint64_t lo = cand_offsets[r], hi = cand_offsets[r + 1];
  const uint64_t per = (hi - lo + gridDim.x - 1) / gridDim.x;
  const uint64_t c_lo64 = lo + per * blockIdx.x < hi ? lo + per * blockIdx.x : hi;
  const uint64_t c_hi64 = c_lo64 + per < hi ? c_lo64 + per : hi;
  const uint32_t n_c = (uint32_t)(c_hi64 - c_lo64);
  const uint32_t first_base = (uint32_t)(c_lo64 - lo);
  const uint2* __restrict__ rows2 = (const uint2*)rows + c_lo64;
  constexpr int U = 4;
  uint2 nxt[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t i = threadIdx.x + u * 256;
    nxt[u] = i < n_c ? __ldcs(rows2 + i) : make_uint2(0xFFFFFFFFu, 0);
  }
  stage_block(smem, gblock, m.stage_bytes, &bar);
  const char* st = smem;
  const int32_t* var = (const int32_t*)(st + m.off_var);
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  const PROG prog(m, idx, st, gblock);
  const uint32_t n_entities = m.n_entities;
  const int32_t n_values = (int32_t)m.n_values;
  // acceptor as one branch-free form on deltas: accept(d) = (A < d) || (d >= B) (see score_list_change_fast_kernel)
  Key kA = Key::highest(), kB = Key::lowest(), tb = Key::lowest();
  uint32_t tb_n = 0, tb_first = 0xFFFFFFFFu, tb_second = 0xFFFFFFFFu, t_acc = 0;
  if (FORAGE) {
    S f_lh, f_ls, f_th, f_ts;
    rel_threshold(fa.ref_scores ? fa.ref_scores[r * 4 + 0] : 0, ch, f_lh);
    rel_threshold(fa.ref_scores ? fa.ref_scores[r * 4 + 1] : 0, csf, f_ls);
    rel_threshold(fa.ref_scores ? fa.ref_scores[r * 4 + 2] : 0, ch, f_th);
    rel_threshold(fa.ref_scores ? fa.ref_scores[r * 4 + 3] : 0, csf, f_ts);
    const Key kl = Key::make(f_lh, f_ls), kt = Key::make(f_th, f_ts);
    const int acc = fa.f.acceptor;
    if (acc == 1 || acc == 3) kA = kl;
    if (acc == 1) kB = Key::highest();
    else if (acc == 2) kB = kl.less(kt) ? kl : kt;
    else if (acc == 3) kB = kt;
  }
  longlong2* __restrict__ scores_c = out_scores ? (longlong2*)out_scores + c_lo64 : nullptr;
  uint8_t* __restrict__ doable_c = out_doable ? out_doable + c_lo64 : nullptr;
  const bool store = !FORAGE || out_scores != nullptr;
  const uint32_t stride = 256 * U;
  for (uint32_t base = threadIdx.x; base < n_c; base += stride) {
    uint2 row[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      row[u] = nxt[u];
      const uint32_t i = base + stride + u * 256;
      nxt[u] = i < n_c ? __ldcs(rows2 + i) : make_uint2(0xFFFFFFFFu, 0);
    }
    uint32_t e[U];
    int32_t nv[U], ov[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {  // ChangeMove doability (change.rs:125-139)
      nv[u] = (int32_t)row[u].y;
      ok[u] = row[u].x < n_entities && nv[u] < n_values;
      e[u] = ok[u] ? row[u].x : 0;
      if (nv[u] < 0) nv[u] = SFGPU_NONE;
      ov[u] = var[e[u]];
      ok[u] = ok[u] && ov[u] != nv[u];
      if (!ok[u]) nv[u] = ov[u];  // null edit: every table index stays in range
    }
    S dh[U], ds[U];
#pragma unroll
    for (int u = 0; u < U; ++u) prog.delta(e[u], ov[u], nv[u], dh[u], ds[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t i = base + u * 256;
      if (i >= n_c) break;
      if (store) {
        longlong2 o;
        o.x = ok[u] ? ch + (int64_t)dh[u] : 0;
        o.y = ok[u] ? csf + (int64_t)ds[u] : 0;
        __stcs(scores_c + i, o);
        doable_c[i] = ok[u] ? 1 : 0;
      }
      if (FORAGE) {
        const Key k = Key::make(dh[u], ds[u]);
        if (ok[u] && (kA.less(k) || !k.less(kB))) {
          t_acc++;
          if (tb_n == 0 || tb.less(k)) {
            tb = k;
            tb_n = 1;
            tb_first = first_base + i;
            tb_second = 0xFFFFFFFFu;
          } else if (tb.equal(k)) {
            if (tb_n == 1) tb_second = first_base + i;  // a thread's rows come in increasing pull order
            tb_n++;
          }
        }
      }
    }
  }
  if (FORAGE) {
    __shared__ int64_t sh_h[8], sh_s[8];
    __shared__ uint32_t sh_n[8], sh_f[8], sh_a[8], sh_2[8];
    for (int o = 16; o > 0; o >>= 1) {
      const Key ok_ = tb.shfl_down(o);
      const uint32_t on = __shfl_down_sync(0xffffffffu, tb_n, o), of = __shfl_down_sync(0xffffffffu, tb_first, o);
      const uint32_t osec = __shfl_down_sync(0xffffffffu, tb_second, o);
      t_acc += __shfl_down_sync(0xffffffffu, t_acc, o);
      if (on && (!tb_n || tb.less(ok_))) {
        tb = ok_; tb_n = on; tb_first = of; tb_second = osec;
      } else if (on && tb_n && tb.equal(ok_)) {
        tb_n += on;
        tb_second = min(max(tb_first, of), min(tb_second, osec));  // second smallest of the four indices
        tb_first = min(tb_first, of);
      }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
      sh_h[warp] = tb_n ? ch + tb.hard() : 0; sh_s[warp] = tb_n ? csf + tb.soft() : 0; sh_n[warp] = tb_n; sh_f[warp] = tb_first;
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
