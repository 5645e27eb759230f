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

// The scoring program of a model as the kernels see it: delta(e, old, new) of one ChangeMove-shaped edit.
// InterpProg walks the constraint table (any program); SpecProg is the monomorphised tuple.
struct InterpProg {
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

// resident CTAs per SM the rows-resident kernel is compiled for: light programs are latency-bound on their table
// gathers and gain from a fifth CTA (<= 51 registers); heavy ones keep their registers
template <class PROG>
struct SpecOccupancy {
  static constexpr int min_ctas = 1;
};
template <>
struct SpecOccupancy<SpecProg<SFGPU_K_PAIR_CSR_EQUAL, SPEC_K_UNI_CONST, 0, 0>> {
  static constexpr int min_ctas = 5;
};

// Rows-resident ChangeMove scoring with a monomorphised program; same contract as
// score_scalar_kernel<MODE_CHANGE, true> (grid = (chunks, R), staged replica block).
template <class PROG>
__global__ void __launch_bounds__(256, SpecOccupancy<PROG>::min_ctas) spec_change_kernel(const __grid_constant__ DevModel m, const SpecIdx idx,
                                                          const uint64_t* __restrict__ cand_offsets,
                                                          const uint32_t* __restrict__ rows,
                                                          int64_t* __restrict__ out_scores,
                                                          uint8_t* __restrict__ out_doable) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  const uint32_t r = blockIdx.y;
  const char* gblock = m.state + (size_t)r * m.block_bytes;
  stage_block(smem, gblock, m.stage_bytes, &bar);
  const char* st = smem;
  const int32_t* var = (const int32_t*)(st + m.off_var);
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  const PROG prog(m, idx, st, gblock);
  const uint32_t n_entities = m.n_entities;
  const int32_t n_values = (int32_t)m.n_values;
  const uint64_t lo = cand_offsets[r], hi = cand_offsets[r + 1];
  constexpr int U = 4;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < hi; base += stride * U) {
    uint2 row[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = base + u * stride;
      row[u] = i < hi ? __ldcs((const uint2*)rows + i) : make_uint2(0xFFFFFFFFu, 0);
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
    int64_t dh[U], ds[U];
#pragma unroll
    for (int u = 0; u < U; ++u) prog.delta(e[u], ov[u], nv[u], dh[u], ds[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = base + u * stride;
      if (i >= hi) break;
      longlong2 o;
      o.x = ok[u] ? ch + dh[u] : 0;
      o.y = ok[u] ? csf + ds[u] : 0;
      __stcs((longlong2*)out_scores + i, o);
      out_doable[i] = ok[u] ? 1 : 0;
    }
  }
}
