// Device-side ChangeMove neighbourhood fused with scoring and the forager replay — the scalar
// counterpart of sfgpu_nearby.cuh.
//
// Reference: ChangeMoveSelector (solverforge-solver/src/heuristic/selector/move_selector/change.rs:66-104,
// 246-307), canonical SelectionOrder::Original: entities in order; per entity every value 0..k-1 in
// order, then the to-None move when allows_unassigned and the entity is currently assigned. Pull index
// of (e, v) = e*k + #assigned entities before e + v (v = k for the to-None move). ChangeMove::is_doable:
// target != current (change.rs:125-139). One thread = one entity; it scores its k(+1) candidates with
// the generic scalar_edit_delta (every scalar constraint kind) and keeps a forager partial.
#pragma once
#include "sfgpu_kernels.cuh"
#include "sfgpu_spec.cuh"


// number of currently assigned entities in [0, end) — block-wide, all threads must call
__device__ __forceinline__ uint32_t block_count_assigned(const int32_t* var, uint32_t end, uint32_t* scratch) {
  uint32_t c = 0;
  for (uint32_t i = threadIdx.x; i < end; i += blockDim.x) c += var[i] >= 0 ? 1 : 0;
  uint32_t tot;
  block_scan_u32(c, scratch, &tot);
  return tot;
}

// PROG = the scoring program (sfgpu_spec.cuh): InterpProg for any constraint table, SpecProg<..> for the
// monomorphised tuples.
template <bool STAGED, class PROG>
__global__ void __launch_bounds__(256) change_step_kernel(const __grid_constant__ DevModel m, const ChangeStepArgs a,
                                                          const SpecIdx idx) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t scratch[33];
  __shared__ int64_t sh_h[8], sh_s[8];
  __shared__ uint32_t sh_n[8], sh_f[8], sh_a[8];
  const uint32_t r = blockIdx.y;
  const char* gblock = m.state + (size_t)r * m.block_bytes;
  const char* st = gblock;
  if (STAGED) {
    stage_block(smem, gblock, m.stage_bytes, &bar);
    st = smem;
  }
  const int32_t* var = (const int32_t*)(st + m.off_var);
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  const PROG prog(m, idx, st, gblock);
  typedef typename PROG::S S;
  typedef DeltaKey<S> Key;
  const uint32_t k = m.n_values, n = m.n_entities;
  const bool with_none = m.allows_unassigned != 0;
  // acceptor on score deltas: accept(d) = (A < d) || (d >= B), thresholds relative to the committed score
  // (see spec_change_kernel; the int32 programs compare one packed 64-bit key)
  Key kA = Key::highest(), kB = Key::lowest(), tb = Key::lowest();
  {
    S f_lh, f_ls, f_th, f_ts;
    rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 0] : 0, ch, f_lh);
    rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 1] : 0, csf, f_ls);
    rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 2] : 0, ch, f_th);
    rel_threshold(a.ref_scores ? a.ref_scores[r * 4 + 3] : 0, csf, f_ts);
    const Key kl = Key::make(f_lh, f_ls), kt = Key::make(f_th, f_ts);
    const int acc = a.f.acceptor;
    if (acc == 1 || acc == 3) kA = kl;
    if (acc == 1) kB = Key::highest();
    else if (acc == 2) kB = kl.less(kt) ? kl : kt;
    else if (acc == 3) kB = kt;
  }
  const uint32_t c_lo = blockIdx.x * a.ents_per_cta, c_hi = min(c_lo + a.ents_per_cta, n);
  const size_t stride = (size_t)n * (k + 1);  // padded rows per replica in the materialised batch
  if (a.out_offsets && blockIdx.x == 0 && threadIdx.x == 0) {
    a.out_offsets[r] = (size_t)r * stride;
    if (r == m.R - 1) a.out_offsets[m.R] = (size_t)m.R * stride;
  }
  uint32_t assigned_before = 0;
  if (a.out_rows) assigned_before = block_count_assigned(var, c_lo, scratch);
  if (a.out_rows && blockIdx.x == 0) {
    // padding of the materialised batch: rows beyond this replica's neighbourhood are not doable
    const uint32_t total_assigned = with_none ? block_count_assigned(var, n, scratch) : 0;
    if (a.out_counts && threadIdx.x == 0) a.out_counts[r] = n * k + total_assigned;
    for (size_t q = (size_t)n * k + total_assigned + threadIdx.x; q < stride; q += blockDim.x) {
      ((uint2*)a.out_rows)[(size_t)r * stride + q] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
      if (a.out_scores) {
        ((longlong2*)a.out_scores)[(size_t)r * stride + q] = make_longlong2(0, 0);
        a.out_doable[(size_t)r * stride + q] = 0;
      }
    }
  }
  uint32_t tb_n = 0, tb_first = 0xFFFFFFFFu, t_acc = 0;
  for (uint32_t g = c_lo; g < c_hi; g += blockDim.x) {
    const uint32_t e = g + threadIdx.x;
    const bool live = e < c_hi;
    const int32_t old = live ? var[e] : SFGPU_NONE;
    uint32_t base = 0;
    if (a.out_rows) {
      // pull index of this entity's first candidate: e*k + #assigned before e
      uint32_t tot;
      const uint32_t incl = block_scan_u32((live && old >= 0 && with_none) ? 1u : 0u, scratch, &tot);
      base = e * k + assigned_before + incl - ((live && old >= 0 && with_none) ? 1u : 0u);
      assigned_before += tot;
    }
    if (!live) continue;
    const uint32_t n_cand = k + ((with_none && old >= 0) ? 1 : 0);
    // four candidates of the entity at a time: their table gathers are independent, so they overlap
    constexpr uint32_t U = 4;
    for (uint32_t v0 = 0; v0 < n_cand; v0 += U) {
      S dh[U], ds[U];
      int32_t nv[U];
      bool ok[U];
#pragma unroll
      for (uint32_t u = 0; u < U; ++u) {
        const uint32_t v = v0 + u;
        nv[u] = v < k ? (int32_t)v : SFGPU_NONE;
        ok[u] = v < n_cand && nv[u] != old;
        prog.delta(e, old, ok[u] ? nv[u] : old, dh[u], ds[u]);  // null edit when not doable
      }
#pragma unroll
      for (uint32_t u = 0; u < U; ++u) {
        const uint32_t v = v0 + u;
        if (v >= n_cand) break;
        if (a.out_rows) {
          const size_t q = (size_t)r * stride + base + v;
          ((uint2*)a.out_rows)[q] = make_uint2(e, (uint32_t)nv[u]);
          if (a.out_scores) {
            ((longlong2*)a.out_scores)[q] = make_longlong2(ok[u] ? ch + (int64_t)dh[u] : 0, ok[u] ? csf + (int64_t)ds[u] : 0);
            a.out_doable[q] = ok[u] ? 1 : 0;
          }
        }
        const Key kk = Key::make(dh[u], ds[u]);
        if (ok[u] && (kA.less(kk) || !kk.less(kB))) {
          t_acc++;
          if (tb_n == 0 || tb.less(kk)) {
            tb = kk;
            tb_n = 1;
            tb_first = e * (k + 1) + v;
          } else if (tb.equal(kk)) {
            tb_n++;
          }
        }
      }
    }
  }
  // block merge of the forager partial (better score wins; equal adds multiplicity, keeps earliest)
  for (int o = 16; o > 0; o >>= 1) {
    const Key ok_ = tb.shfl_down(o);
    const uint32_t on = __shfl_down_sync(0xffffffffu, tb_n, o), of = __shfl_down_sync(0xffffffffu, tb_first, o);
    t_acc += __shfl_down_sync(0xffffffffu, t_acc, o);
    if (on && (!tb_n || tb.less(ok_))) {
      tb = ok_; tb_n = on; tb_first = of;
    } else if (on && tb_n && tb.equal(ok_)) {
      tb_n += on;
      tb_first = min(tb_first, of);
    }
  }
  const int64_t tb_h = tb_n ? ch + tb.hard() : 0, tb_s = tb_n ? csf + tb.soft() : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh_h[warp] = tb_h; sh_s[warp] = tb_s; sh_n[warp] = tb_n; sh_f[warp] = tb_first; sh_a[warp] = t_acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ChunkPartial cp{0, 0, 0, 0, 0xFFFFFFFFu, 0};
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

// Ordered search inside one chunk of entities [c_lo, c_hi): the `want`-th candidate (1-based, pull
// order) satisfying pred(accepted, score). Returns e*(k+1)+v via s_out. All 256 threads participate.
// mode 0: accepted candidates (AcceptedCount cut); mode 1: accepted candidates equal to (bh, bs).
template <class PROG>
__device__ __forceinline__ void chunk_find(const DevModel& m, const char* st, const PROG& prog, const ForageDev& f, uint32_t c_lo,
                                           uint32_t c_hi, uint32_t limit_code, int mode, int64_t bh, int64_t bs,
                                           uint32_t want, int64_t lh, int64_t ls, int64_t th, int64_t ts,
                                           uint32_t* scratch, uint32_t* s_out) {
  const int32_t* var = (const int32_t*)(st + m.off_var);
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  const int64_t ch = cs[0], csf = cs[1];
  const uint32_t k = m.n_values;
  const bool with_none = m.allows_unassigned != 0;
  uint32_t seen = 0;
  for (uint32_t g = c_lo; g < c_hi; g += blockDim.x) {
    const uint32_t e = g + threadIdx.x;
    const bool live = e < c_hi;
    const int32_t old = live ? var[e] : SFGPU_NONE;
    const uint32_t n_cand = live ? k + ((with_none && old >= 0) ? 1 : 0) : 0;
    uint32_t mine = 0;
    for (uint32_t v = 0; v < n_cand; ++v) {
      if (e * (k + 1) + v > limit_code) break;  // beyond the AcceptedCount cut
      const int32_t nv = v < k ? (int32_t)v : SFGPU_NONE;
      if (nv == old) continue;
      typename PROG::S dh, ds;
      prog.delta(e, old, nv, dh, ds);
      const int64_t oh = ch + (int64_t)dh, os = csf + (int64_t)ds;
      if (accept_score(f.acceptor, oh, os, lh, ls, th, ts) && (mode == 0 || (oh == bh && os == bs))) mine++;
    }
    uint32_t tot;
    const uint32_t incl = block_scan_u32(mine, scratch, &tot);
    const uint32_t excl = seen + incl - mine;
    if (mine && excl < want && want <= excl + mine) {
      // this thread's (want - excl)-th hit
      uint32_t hit = 0;
      for (uint32_t v = 0; v < n_cand; ++v) {
        if (e * (k + 1) + v > limit_code) break;
        const int32_t nv = v < k ? (int32_t)v : SFGPU_NONE;
        if (nv == old) continue;
        typename PROG::S dh, ds;
        prog.delta(e, old, nv, dh, ds);
        const int64_t oh = ch + (int64_t)dh, os = csf + (int64_t)ds;
        if (accept_score(f.acceptor, oh, os, lh, ls, th, ts) && (mode == 0 || (oh == bh && os == bs))) {
          if (++hit == want - excl) {
            *s_out = e * (k + 1) + v;
            break;
          }
        }
      }
    }
    seen += tot;
    __syncthreads();
    if (seen >= want) break;
  }
}

// One CTA per replica: combine chunk partials, AcceptedCount cut, tie rule, winner. The ordered searches inside a chunk
// (AcceptedCount cut, j-th of several equal bests) re-score that chunk with the same program the step kernel ran
// (PROG, replica block staged): with the interpreter on the unstaged block they cost 3-4 x the whole step kernel.
template <bool STAGED, class PROG>
__global__ void __launch_bounds__(256) change_finish_kernel(const __grid_constant__ DevModel m, const ChangeStepArgs a,
                                                            const SpecIdx idx, uint32_t n_chunks,
                                                            uint32_t* __restrict__ out_index,
                                                            int64_t* __restrict__ out_best,
                                                            uint32_t* __restrict__ out_evaluated,
                                                            uint32_t* __restrict__ out_winner_rows) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t scratch[33];
  __shared__ uint32_t s_code, s_cut_chunk, s_cut_rank, s_any, s_cstar, s_jstar;
  __shared__ int64_t s_bh, s_bs;
  __shared__ ChunkPartial s_cutp;
  const uint32_t r = blockIdx.x;
  const char* gblock = m.state + (size_t)r * m.block_bytes;
  const char* st = gblock;
  if (STAGED) {
    stage_block(smem, gblock, m.stage_bytes, &bar);
    st = smem;
  }
  const PROG prog(m, idx, st, gblock);
  const int32_t* var = (const int32_t*)(st + m.off_var);
  const uint32_t k = m.n_values, n = m.n_entities;
  const ChunkPartial* P = a.partials + (size_t)r * n_chunks;
  int64_t lh = 0, ls = 0, th = 0, ts = 0;
  if (a.ref_scores) {
    lh = a.ref_scores[r * 4 + 0];
    ls = a.ref_scores[r * 4 + 1];
    th = a.ref_scores[r * 4 + 2];
    ts = a.ref_scores[r * 4 + 3];
  }
  uint32_t limit_code = 0xFFFFFFFFu;  // e*(k+1)+v of the last pull that counts
  uint32_t n_eff = n_chunks;
  if (threadIdx.x == 0) {
    s_cut_chunk = n_chunks;
    s_cutp = ChunkPartial{0, 0, 0, 0, 0xFFFFFFFFu, 0};
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
      const uint32_t c_lo = s_cut_chunk * a.ents_per_cta, c_hi = min(c_lo + a.ents_per_cta, n);
      chunk_find(m, st, prog, a.f, c_lo, c_hi, 0xFFFFFFFFu, 0, 0, 0, s_cut_rank, lh, ls, th, ts, scratch, &s_code);
      __syncthreads();
      limit_code = s_code;
      n_eff = s_cut_chunk + 1;
      // truncated partial of the cut chunk: best / multiplicity / first among pulls <= limit_code
      {
        const int64_t* cs = (const int64_t*)(st + m.off_score);
        const int64_t ch = cs[0], csf = cs[1];
        const bool with_none = m.allows_unassigned != 0;
        int64_t tb_h = 0, tb_s = 0;
        uint32_t tb_n = 0, tb_first = 0xFFFFFFFFu;
        for (uint32_t e = c_lo + threadIdx.x; e < c_hi; e += blockDim.x) {
          const int32_t old = var[e];
          const uint32_t n_cand = k + ((with_none && old >= 0) ? 1 : 0);
          for (uint32_t v = 0; v < n_cand; ++v) {
            if (e * (k + 1) + v > limit_code) break;
            const int32_t nv = v < k ? (int32_t)v : SFGPU_NONE;
            if (nv == old) continue;
            typename PROG::S dh, ds;
            prog.delta(e, old, nv, dh, ds);
            const int64_t oh = ch + (int64_t)dh, os = csf + (int64_t)ds;
            if (!accept_score(a.f.acceptor, oh, os, lh, ls, th, ts)) continue;
            if (tb_n == 0 || score_less(tb_h, tb_s, oh, os)) {
              tb_h = oh; tb_s = os; tb_n = 1; tb_first = e * (k + 1) + v;
            } else if (tb_h == oh && tb_s == os) {
              tb_n++;
              tb_first = min(tb_first, e * (k + 1) + v);
            }
          }
        }
        // serialised merge through shared memory (rare path)
        for (uint32_t t = 0; t < blockDim.x; ++t) {
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
  }
  __syncthreads();
  const uint32_t cut_chunk = s_cut_chunk;
  // moves_evaluated: pulls up to the cut (pull index + 1), or the whole neighbourhood
  uint32_t evaluated;
  {
    const uint32_t e_end = limit_code == 0xFFFFFFFFu ? n : limit_code / (k + 1);
    const uint32_t before = block_count_assigned(var, e_end, scratch);
    const bool with_none = m.allows_unassigned != 0;
    evaluated = limit_code == 0xFFFFFFFFu ? n * k + (with_none ? before : 0)
                                           : e_end * k + (with_none ? before : 0) + limit_code % (k + 1) + 1;
  }
  // best over chunks
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
      if (out_winner_rows) ((uint2*)out_winner_rows)[r] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    }
    return;
  }
  uint32_t mtot = 0;
  for (uint32_t c = 0; c < n_eff; ++c) {
    const ChunkPartial p = c == cut_chunk ? s_cutp : P[c];
    if (p.n_best && p.best_h == bh && p.best_s == bs) mtot += p.n_best;
  }
  if (a.f.tie_mode == 1) {
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
      const uint32_t nb = (p.n_best && p.best_h == bh && p.best_s == bs) ? p.n_best : 0;
      if (before + nb >= want) {
        s_cstar = c;
        s_jstar = want - before;
        s_code = p.first_idx;
        break;
      }
      before += nb;
    }
  }
  __syncthreads();
  if (s_jstar > 1) {
    const uint32_t c_lo = s_cstar * a.ents_per_cta, c_hi = min(c_lo + a.ents_per_cta, n);
    const uint32_t j = s_jstar;
    __syncthreads();
    chunk_find(m, st, prog, a.f, c_lo, c_hi, limit_code, 1, bh, bs, j, lh, ls, th, ts, scratch, &s_code);
    __syncthreads();
  }
  const uint32_t code = s_code;
  const uint32_t we = code / (k + 1), wv = code % (k + 1);
  const uint32_t before = block_count_assigned(var, we, scratch);
  if (threadIdx.x == 0) {
    out_index[r] = we * k + (m.allows_unassigned ? before : 0) + wv;
    out_best[r * 2] = bh;
    out_best[r * 2 + 1] = bs;
    if (out_evaluated) out_evaluated[r] = evaluated;
    if (out_winner_rows) ((uint2*)out_winner_rows)[r] = make_uint2(we, wv < k ? wv : 0xFFFFFFFFu);
  }
}
