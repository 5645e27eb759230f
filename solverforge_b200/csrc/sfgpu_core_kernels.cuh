// Non-template kernels of libsfgpu: compiled once, in sfgpu_api.cu (the other translation units reach them
// through the sfgpu_launch_* wrappers declared in sfgpu_ctx.hpp).
#pragma once
#include "sfgpu_kernels.cuh"

// Finishes the fused forager: one warp per replica combines the chunk partials, applies the tie
// rule (largest firing k <= m, BestCandidate::consider) and, only when the winner is not the first
// best row of its chunk, rescans that chunk in pull order (scores re-derived with the generic
// per-candidate function, or read back when they were materialised).
__global__ void __launch_bounds__(256) forage_finish_kernel(const __grid_constant__ DevModel m, ForageArgs fa,
                                                            uint32_t n_chunks,
                                                            const uint64_t* __restrict__ cand_offsets,
                                                            const uint32_t* __restrict__ rows,
                                                            const int64_t* __restrict__ scores,
                                                            const uint8_t* __restrict__ doable,
                                                            const uint64_t* __restrict__ step_seeds,
                                                            uint32_t* __restrict__ out_index,
                                                            int64_t* __restrict__ out_best,
                                                            uint32_t* __restrict__ out_evaluated) {
  __shared__ int64_t s_bh, s_bs;
  __shared__ uint32_t s_any, s_cstar, s_j, s_winner, s_warp_cnt[8];
  const uint32_t r = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const ChunkPartial* cp = fa.partials + (size_t)r * n_chunks;
  const uint64_t lo = cand_offsets[r], hi = cand_offsets[r + 1];
  if (warp == 0) {
    // global best over chunks
    int64_t bh = 0, bs = 0;
    uint32_t any = 0;
    for (uint32_t c = lane; c < n_chunks; c += 32) {
      if (!cp[c].n_best) continue;
      if (!any || score_less(bh, bs, cp[c].best_h, cp[c].best_s)) {
        bh = cp[c].best_h;
        bs = cp[c].best_s;
      }
      any = 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const int64_t oh = __shfl_xor_sync(0xffffffffu, bh, o), os = __shfl_xor_sync(0xffffffffu, bs, o);
      const uint32_t oa = __shfl_xor_sync(0xffffffffu, any, o);
      if (oa && (!any || score_less(bh, bs, oh, os))) {
        bh = oh;
        bs = os;
      }
      any |= oa;
    }
    uint32_t mtot = 0;
    for (uint32_t c = lane; c < n_chunks; c += 32)
      if (any && cp[c].n_best && cp[c].best_h == bh && cp[c].best_s == bs) mtot += cp[c].n_best;
    for (int o = 16; o > 0; o >>= 1) mtot += __shfl_xor_sync(0xffffffffu, mtot, o);
    // BestCandidate::consider: the winner is the occurrence with the largest firing k <= m
    uint32_t want = 1;
    if (fa.f.tie_mode == 1) {
      const uint64_t seed = step_seeds ? step_seeds[r] : 0;
      for (uint32_t k = 2 + lane; k <= mtot; k += 32) {
        const uint64_t mixed = splitmix64_dev(seed ^ ((uint64_t)k * 0x9E3779B97F4A7C15ull) ^ 0xF04A63E239B74D11ull);
        if (mixed % k == 0) want = k;
      }
      for (int o = 16; o > 0; o >>= 1) want = max(want, __shfl_xor_sync(0xffffffffu, want, o));
    }
    if (lane == 0) {
      // chunk holding the want-th occurrence (chunks are contiguous and in pull order)
      uint32_t c_star = 0, before = 0;
      for (uint32_t c = 0; any && c < n_chunks; ++c) {
        const uint32_t nb = (cp[c].n_best && cp[c].best_h == bh && cp[c].best_s == bs) ? cp[c].n_best : 0;
        if (before + nb >= want) {
          c_star = c;
          break;
        }
        before += nb;
      }
      s_bh = bh;
      s_bs = bs;
      s_any = any;
      s_cstar = c_star;
      s_j = want - before;  // 1-based rank inside the chunk
      s_winner = any ? cp[c_star].first_idx : 0xFFFFFFFFu;
      if (any && s_j == 2 && cp[c_star].second_idx != 0xFFFFFFFFu) {  // the chunk's second best row is recorded
        s_winner = cp[c_star].second_idx;
        s_j = 1;
      }
      if (out_evaluated) out_evaluated[r] = (uint32_t)(hi - lo);
    }
  }
  __syncthreads();
  const int64_t bh = s_bh, bs = s_bs;
  const uint32_t j = s_j;
  if (s_any && j > 1) {
    // the winner is not the chunk's first best row: ordered rescan of that chunk, 4 consecutive
    // rows per thread so the loads are independent and the pull order is thread order
    const uint64_t per = (hi - lo + n_chunks - 1) / n_chunks;
    const uint64_t c_lo = lo + per * s_cstar;
    const uint64_t c_hi = c_lo + per < hi ? c_lo + per : hi;
    const char* st = m.state + (size_t)r * m.block_bytes;
    const int64_t* cs = (const int64_t*)(st + m.off_score);
    int64_t lh = 0, ls = 0, th = 0, ts = 0;
    if (fa.ref_scores) {
      lh = fa.ref_scores[r * 4 + 0];
      ls = fa.ref_scores[r * 4 + 1];
      th = fa.ref_scores[r * 4 + 2];
      ts = fa.ref_scores[r * 4 + 3];
    }
    // every thread owns a contiguous run of rows (thread order == pull order) and first collects the
    // hit mask of its whole run with independent loads; one block scan then locates the j-th hit.
    // Super-blocks of 256 * 64 rows keep the mask in one 64-bit register.
    uint32_t seen = 0;
    for (uint64_t base = c_lo; base < c_hi; base += 256 * 64) {
      const uint64_t b_hi = base + 256 * 64 < c_hi ? base + 256 * 64 : c_hi;
      const uint32_t run = (uint32_t)((b_hi - base + 255) / 256);  // rows per thread, <= 64
      const uint64_t t_lo = base + (uint64_t)threadIdx.x * run;
      uint64_t hits = 0;  // bit q set: row t_lo + q is an accepted best row
      if (scores) {
        // eight independent 128-bit loads in flight per thread: the rescan is latency-, not bandwidth-bound
        for (uint32_t q0 = 0; q0 < run; q0 += 8) {
          longlong2 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint64_t i = t_lo + q0 + u;
            v[u] = (q0 + u < run && i < b_hi) ? __ldcs((const longlong2*)scores + i) : make_longlong2(bh - 1, 0);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint64_t i = t_lo + q0 + u;
            if (v[u].x == bh && v[u].y == bs && q0 + u < run && i < b_hi && doable[i] != 0 &&
                accept_score(fa.f.acceptor, bh, bs, lh, ls, th, ts))
              hits |= 1ull << (q0 + u);
          }
        }
      } else {
        for (uint32_t q = 0; q < run; ++q) {
          const uint64_t i = t_lo + q;
          if (i >= b_hi) break;
          Score2 d;
          const bool ok = fa.row_kind == 1 ? score_change_row(m, st, st, ((const uint2*)rows)[i], d)
                                           : list_change_delta(m, st, ((const uint4*)rows)[i], d);
          const int64_t h = cs[0] + d.hard, s2 = cs[1] + d.soft;
          if (ok && h == bh && s2 == bs && accept_score(fa.f.acceptor, h, s2, lh, ls, th, ts)) hits |= 1ull << q;
        }
      }
      const uint32_t mine = __popcll(hits);
      // inclusive scan over the 256 threads
      uint32_t incl = mine;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      if (lane == 31) s_warp_cnt[warp] = incl;
      __syncthreads();
      uint32_t warp_base = 0, total = 0;
      for (uint32_t w = 0; w < 8; ++w) {
        if (w < warp) warp_base += s_warp_cnt[w];
        total += s_warp_cnt[w];
      }
      const uint32_t excl = seen + warp_base + incl - mine;
      if (mine && excl < j && j <= excl + mine) {
        uint64_t mm = hits;
        for (uint32_t t = 1; t < j - excl; ++t) mm &= mm - 1;
        s_winner = (uint32_t)(t_lo + (__ffsll((long long)mm) - 1) - lo);
      }
      seen += total;
      __syncthreads();
      if (seen >= j) break;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    out_index[r] = s_winner;
    out_best[r * 2] = s_any ? bh : 0;
    out_best[r * 2 + 1] = s_any ? bs : 0;
  }
}

// First phase of the two-phase argbest over MATERIALISED scores (BestScore forager): grid (chunks, R);
// every CTA reduces one contiguous chunk of a replica's rows to a ChunkPartial (best accepted score, its
// multiplicity, first and second row, accepted count); forage_finish_kernel then replays the tie rule.
// One CTA per replica (argbest_kernel) leaves most SMs idle when R is small and the rows are many.
__global__ void __launch_bounds__(256) argbest_partial_kernel(ForageDev f, const uint64_t* __restrict__ cand_offsets,
                                                              const int64_t* __restrict__ scores,
                                                              const uint8_t* __restrict__ doable,
                                                              const int64_t* __restrict__ ref_scores,
                                                              ChunkPartial* __restrict__ partials) {
  const uint32_t r = blockIdx.y;
  const uint64_t lo = cand_offsets[r], hi = cand_offsets[r + 1];
  const uint64_t per = (hi - lo + gridDim.x - 1) / gridDim.x;
  const uint64_t c_lo = lo + per * blockIdx.x < hi ? lo + per * blockIdx.x : hi;
  const uint64_t c_hi = c_lo + per < hi ? c_lo + per : hi;
  const int64_t lh = ref_scores ? ref_scores[r * 4 + 0] : 0, ls = ref_scores ? ref_scores[r * 4 + 1] : 0;
  const int64_t th = ref_scores ? ref_scores[r * 4 + 2] : 0, ts = ref_scores ? ref_scores[r * 4 + 3] : 0;
  int64_t tb_h = 0, tb_s = 0;
  uint32_t tb_n = 0, tb_first = 0xFFFFFFFFu, tb_second = 0xFFFFFFFFu, t_acc = 0;
  for (uint64_t base = c_lo + threadIdx.x; base < c_hi; base += 256 * 4) {
    longlong2 v[4];
    uint8_t ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t i = base + (uint64_t)u * 256;
      v[u] = i < c_hi ? __ldcs((const longlong2*)scores + i) : make_longlong2(0, 0);
      ok[u] = i < c_hi ? doable[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t i = base + (uint64_t)u * 256;
      if (!ok[u] || !accept_score(f.acceptor, v[u].x, v[u].y, lh, ls, th, ts)) continue;
      t_acc++;
      if (tb_n == 0 || score_less(tb_h, tb_s, v[u].x, v[u].y)) {
        tb_h = v[u].x;
        tb_s = v[u].y;
        tb_n = 1;
        tb_first = (uint32_t)(i - lo);
        tb_second = 0xFFFFFFFFu;
      } else if (tb_h == v[u].x && tb_s == v[u].y) {
        if (tb_n == 1) tb_second = (uint32_t)(i - lo);  // a thread's rows come in increasing pull order
        tb_n++;
      }
    }
  }
  __shared__ int64_t sh_h[8], sh_s[8];
  __shared__ uint32_t sh_n[8], sh_f[8], sh_2[8], sh_a[8];
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t oh = __shfl_down_sync(0xffffffffu, tb_h, o), os = __shfl_down_sync(0xffffffffu, tb_s, o);
    const uint32_t on = __shfl_down_sync(0xffffffffu, tb_n, o), of = __shfl_down_sync(0xffffffffu, tb_first, o);
    const uint32_t osec = __shfl_down_sync(0xffffffffu, tb_second, o);
    t_acc += __shfl_down_sync(0xffffffffu, t_acc, o);
    if (on && (!tb_n || score_less(tb_h, tb_s, oh, os))) {
      tb_h = oh; tb_s = os; tb_n = on; tb_first = of; tb_second = osec;
    } else if (on && tb_n && oh == tb_h && os == tb_s) {
      tb_n += on;
      tb_second = min(max(tb_first, of), min(tb_second, osec));
      tb_first = min(tb_first, of);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh_h[warp] = tb_h; sh_s[warp] = tb_s; sh_n[warp] = tb_n; sh_f[warp] = tb_first; sh_2[warp] = tb_second;
    sh_a[warp] = t_acc;
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
    partials[(size_t)r * gridDim.x + blockIdx.x] = cp;
  }
}

// =============================================================================================
// initialize_all / evaluate_all: one CTA per replica, full recompute + (re)build of the retained
// aggregates inside the block it is pointed at (the live state at commit, a scratch copy for
// sfgpu_evaluate_all).  Reference: IncrementalConstraint::evaluate of every constraint kind.
// =============================================================================================
__global__ void __launch_bounds__(256) init_kernel(const __grid_constant__ DevModel m, char* __restrict__ state) {
  extern __shared__ __align__(16) uint32_t init_smem[];  // n_owners + 1 words (build_fast_records)
  __shared__ int64_t scratch[32];
  const uint32_t r = blockIdx.x;
  char* st = state + (size_t)r * m.block_bytes;
  if (state == m.state && threadIdx.x == 0) nbc_invalidate(m.nbc_tag, r);
  const int32_t* var = (const int32_t*)(st + m.off_var);
  const uint32_t* off = (const uint32_t*)(st + m.off_offsets);
  const uint32_t* el = (const uint32_t*)(st + m.off_elems);
  int64_t hard = 0, soft = 0;
  for (uint32_t k = 0; k < m.n_cons; ++k) {
    const ConsDev& c = m.cons[k];
    int64_t local = 0;
    switch (c.kind) {
      case SFGPU_K_UNI:
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x) local += uni_contrib(c, e, var[e]);
        break;
      case SFGPU_K_JOIN_EXPR:
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x) local += join_expr_contrib(c, e, var[e]);
        break;
      case SFGPU_K_PAIR_KEY_EXPR: {
        const ExprTables t{(const int64_t* const*)c.g1, (const uint32_t* const*)c.g2};
        const PairKeyLists L = pke_lists(c, st, m.n_entities);
        const bool directed = pke_directed(c);
        const uint32_t nk = (uint32_t)c.p1;
        for (uint32_t i = threadIdx.x; i < nk; i += blockDim.x) L.headL[i] = L.headR[i] = -1;
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x) L.nxtL[e] = L.prvL[e] = L.nxtR[e] = L.prvR[e] = -1;
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x) {  // lock-free push; list order does not matter
          const int64_t kl = pke_key(c, t, 0, e, var[e]);
          if (kl >= 0) L.nxtL[e] = atomicExch(&L.headL[kl], (int32_t)e);
          const int64_t kr = directed ? pke_key(c, t, 1, e, var[e]) : -1;
          if (kr >= 0) L.nxtR[e] = atomicExch(&L.headR[kr], (int32_t)e);
        }
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x) {
          if (L.nxtL[e] >= 0) L.prvL[L.nxtL[e]] = (int32_t)e;
          if (L.nxtR[e] >= 0) L.prvR[L.nxtR[e]] = (int32_t)e;
        }
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x) {
          const int64_t kl = pke_key(c, t, 0, e, var[e]);
          if (kl < 0) continue;
          if (directed) {
            for (int32_t x = L.headR[kl]; x >= 0; x = L.nxtR[x])
              if ((uint32_t)x != e) local += pke_pair(c, t, e, var[e], (uint32_t)x, var[x]);
          } else {
            for (int32_t x = L.headL[kl]; x >= 0; x = L.nxtL[x])
              if ((uint32_t)x > e) local += pke_pair(c, t, e, var[e], (uint32_t)x, var[x]);
          }
        }
        break;
      }
      case SFGPU_K_PAIR_CSR_EQUAL: {
        const uint32_t* rp = (const uint32_t*)c.g0;
        const uint32_t* ci = (const uint32_t*)c.g1;
        uint16_t* cc = c.off0 != 0xFFFFFFFFu ? (uint16_t*)(st + c.off0) : nullptr;
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x) {
          int32_t v = var[e];
          if (cc) {
            uint16_t* row = cc + (size_t)e * m.n_values;
            for (uint32_t q = 0; q < m.n_values; ++q) row[q] = 0;
            for (uint32_t j = rp[e]; j < rp[e + 1]; ++j) {
              int32_t vp = var[ci[j]];
              if (vp >= 0) row[vp] += 1;
            }
          }
          if (v < 0) continue;
          for (uint32_t j = rp[e]; j < rp[e + 1]; ++j)
            if (ci[j] > e && var[ci[j]] == v) local += c.w.a;
        }
        break;
      }
      case SFGPU_K_PAIR_KEY_EQUAL: {
        int32_t* tab = (int32_t*)(st + c.off0);
        for (uint32_t i = threadIdx.x; i < c.n0; i += blockDim.x) tab[i] = 0;
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x)
          if (var[e] >= 0) atomicAdd(&tab[pair_key(c, e, var[e])], 1);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < c.n0; i += blockDim.x) {
          int64_t n = tab[i];
          local += choose_small(n, (int)c.pad) * c.w.a;
        }
        break;
      }
      case SFGPU_K_GROUP: {
        int32_t* gc = (int32_t*)(st + c.off0);
        unsigned long long* gs = (unsigned long long*)(st + c.off1);
        for (uint32_t i = threadIdx.x; i < m.n_values; i += blockDim.x) {
          gc[i] = 0;
          gs[i] = 0;
        }
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x)
          if (var[e] >= 0) {
            atomicAdd(&gc[var[e]], 1);
            atomicAdd(&gs[var[e]], (unsigned long long)(c.g0 ? ((const int64_t*)c.g0)[e] : 1));
          }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < m.n_values; i += blockDim.x)
          local += group_score(c, i, gc[i], (int64_t)gs[i]);
        break;
      }
      case SFGPU_K_RUNS: {
        int32_t* cnt = (int32_t*)(st + c.off0);
        const uint32_t np = c.n0;
        for (uint32_t i = threadIdx.x; i < m.n_values * np; i += blockDim.x) cnt[i] = 0;
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x)
          if (var[e] >= 0) atomicAdd(&cnt[(size_t)var[e] * np + (uint32_t)((const int64_t*)c.g0)[e]], 1);
        __syncthreads();
        for (uint32_t v = threadIdx.x; v < m.n_values; v += blockDim.x) {
          if (c.p0 != 0) {
            const int32_t* row = cnt + (size_t)v * np;
            local += presence_group_score(c, [&](int32_t q) { return row[q]; });
            continue;
          }
          int64_t run = 0;
          for (uint32_t q = 0; q <= np; ++q) {
            if (q < np && cnt[(size_t)v * np + q] > 0) {
              ++run;
            } else {
              if (run > 0) local += weight_eval(c.w, run);
              run = 0;
            }
          }
        }
        break;
      }
      case SFGPU_K_PROJECT_GROUP: {
        int32_t* gc = (int32_t*)(st + c.off0);
        unsigned long long* gs = (unsigned long long*)(st + c.off1);
        const uint32_t* rp = (const uint32_t*)c.g0;
        const longlong2* em = (const longlong2*)c.g1;
        const uint32_t n_groups = m.n_values * c.n0;
        for (uint32_t i = threadIdx.x; i < n_groups; i += blockDim.x) {
          gc[i] = 0;
          gs[i] = 0;
        }
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x)
          if (var[e] >= 0)
            for (uint32_t j = rp[e]; j < rp[e + 1]; ++j) {
              const longlong2 row = em[j];
              const uint32_t key = (uint32_t)var[e] * c.n0 + (uint32_t)row.x;
              atomicAdd(&gc[key], 1);
              atomicAdd(&gs[key], (unsigned long long)row.y);
            }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_groups; i += blockDim.x)
          if (gc[i] > 0) local += weight_eval(c.w, (int64_t)gs[i]);
        break;
      }
      case SFGPU_K_LOAD_BALANCE: {
        unsigned long long* loads = (unsigned long long*)(st + c.off0);
        int32_t* icnt = (int32_t*)(st + c.off1);
        int64_t* agg = (int64_t*)(st + c.off2);
        for (uint32_t i = threadIdx.x; i < m.n_values; i += blockDim.x) {
          loads[i] = 0;
          icnt[i] = 0;
        }
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < m.n_entities; e += blockDim.x) {
          int64_t x = c.g0 ? ((const int64_t*)c.g0)[e] : 1;
          if (var[e] >= 0 && x != 0) {
            atomicAdd(&icnt[var[e]], 1);
            atomicAdd(&loads[var[e]], (unsigned long long)x);
          }
        }
        __syncthreads();
        int64_t s = 0, sq = 0, nk = 0;
        for (uint32_t i = threadIdx.x; i < m.n_values; i += blockDim.x) {
          int64_t l = (int64_t)loads[i];
          s += l;
          sq += l * l;
          nk += icnt[i] > 0 ? 1 : 0;
        }
        s = block_sum_i64(s, scratch);
        sq = block_sum_i64(sq, scratch);
        nk = block_sum_i64(nk, scratch);
        if (threadIdx.x == 0) {
          agg[0] = s;
          agg[1] = sq;
          agg[2] = nk;
          local = weight_eval(c.w, lb_unfairness(nk, s, sq));
        }
        break;
      }
      case SFGPU_K_EXISTS_FLAT: {
        int32_t* bc = (int32_t*)(st + c.off0);
        for (uint32_t i = threadIdx.x; i < m.n_elem_rows; i += blockDim.x) bc[i] = 0;
        __syncthreads();
        const uint32_t total = off[m.n_owners];
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x)
          if (el[i] < m.n_elem_rows) atomicAdd(&bc[el[i]], 1);
        __syncthreads();
        // A rows: key_a = key column (g0) or the row index; n0 = |A|
        for (uint32_t a = threadIdx.x; a < c.n0; a += blockDim.x) {
          int64_t key = c.g0 ? ((const int64_t*)c.g0)[a] : (int64_t)a;
          int32_t n = (key >= 0 && key < (int64_t)m.n_elem_rows) ? bc[key] : 0;
          bool match = c.p0 == 0 ? n > 0 : n == 0;
          if (match) local += c.w.a;
        }
        break;
      }
      case SFGPU_K_LIST_PATH_COST: {
        int64_t* rcost = (int64_t*)(st + c.off0);
        const uint32_t depot = (uint32_t)c.p0;
        for (uint32_t o = threadIdx.x; o < m.n_owners; o += blockDim.x) {
          int64_t cost = 0;
          uint32_t b = off[o], e = off[o + 1];
          if (e > b) {
            uint32_t prev = depot;
            for (uint32_t i = b; i < e; ++i) {
              cost += mat_at(c, prev, el[i]);
              prev = el[i];
            }
            cost += mat_at(c, prev, depot);
          }
          rcost[o] = cost;
          local += weight_eval(c.w, cost);
        }
        break;
      }
      case SFGPU_K_LIST_SUM: {
        int64_t* rsum = (int64_t*)(st + c.off0);
        for (uint32_t o = threadIdx.x; o < m.n_owners; o += blockDim.x) {
          int64_t s = 0;
          for (uint32_t i = off[o]; i < off[o + 1]; ++i) s += ((const int64_t*)c.g0)[el[i]];
          rsum[o] = s;
          local += weight_eval(c.w, s);
        }
        break;
      }
      default: break;
    }
    int64_t tot = block_sum_i64(local, scratch);
    tot = c.sign < 0 ? -tot : tot;
    if (c.w.level == 0) hard += tot; else soft += tot;
  }
  if (threadIdx.x == 0) {
    int64_t* cs = (int64_t*)(st + m.off_score);
    cs[0] = hard;
    cs[1] = soft;
  }
  if (m.fast_list) {
    __syncthreads();
    build_fast_records(m, st, init_smem);
  }
}

// =============================================================================================
// apply: one CTA per replica commits one move (Move::do_move on the committed director).
// =============================================================================================
static __device__ void apply_scalar_edit(const DevModel& m, char* st, EditDev cur) {
  // thread 0 only; the retained tables move with the variable
  int32_t* var = (int32_t*)(st + m.off_var);
  for (uint32_t k = 0; k < m.n_cons; ++k) {
    const ConsDev& c = m.cons[k];
    if (c.kind == SFGPU_K_PAIR_CSR_EQUAL && c.off0 != 0xFFFFFFFFu) {
      uint16_t* cc = (uint16_t*)(st + c.off0);
      const uint32_t* rp = (const uint32_t*)c.g0;
      const uint32_t* ci = (const uint32_t*)c.g1;
      for (uint32_t j = rp[cur.e]; j < rp[cur.e + 1]; ++j) {
        uint16_t* row = cc + (size_t)ci[j] * m.n_values;
        if (cur.old_v >= 0) row[cur.old_v] -= 1;
        if (cur.new_v >= 0) row[cur.new_v] += 1;
      }
    } else if (c.kind == SFGPU_K_PAIR_KEY_EXPR) {
      const ExprTables t{(const int64_t* const*)c.g1, (const uint32_t* const*)c.g2};
      const PairKeyLists L = pke_lists(c, st, m.n_entities);
      for (int side = 0; side < (pke_directed(c) ? 2 : 1); ++side) {
        int32_t* head = side ? L.headR : L.headL;
        int32_t* nxt = side ? L.nxtR : L.nxtL;
        int32_t* prv = side ? L.prvR : L.prvL;
        const int64_t ko = pke_key(c, t, side, cur.e, cur.old_v), kn = pke_key(c, t, side, cur.e, cur.new_v);
        if (ko >= 0) {  // unlink
          const int32_t p = prv[cur.e], n = nxt[cur.e];
          if (p >= 0) nxt[p] = n; else head[ko] = n;
          if (n >= 0) prv[n] = p;
          nxt[cur.e] = prv[cur.e] = -1;
        }
        if (kn >= 0) {  // link at the head
          const int32_t n = head[kn];
          nxt[cur.e] = n;
          prv[cur.e] = -1;
          if (n >= 0) prv[n] = (int32_t)cur.e;
          head[kn] = (int32_t)cur.e;
        }
      }
    } else if (c.kind == SFGPU_K_PAIR_KEY_EQUAL) {
      int32_t* tab = (int32_t*)(st + c.off0);
      if (cur.old_v >= 0) tab[pair_key(c, cur.e, cur.old_v)] -= 1;
      if (cur.new_v >= 0) tab[pair_key(c, cur.e, cur.new_v)] += 1;
    } else if (c.kind == SFGPU_K_GROUP) {
      int32_t* gc = (int32_t*)(st + c.off0);
      int64_t* gs = (int64_t*)(st + c.off1);
      int64_t x = c.g0 ? ((const int64_t*)c.g0)[cur.e] : 1;
      if (cur.old_v >= 0) { gc[cur.old_v] -= 1; gs[cur.old_v] -= x; }
      if (cur.new_v >= 0) { gc[cur.new_v] += 1; gs[cur.new_v] += x; }
    } else if (c.kind == SFGPU_K_RUNS) {
      int32_t* cnt = (int32_t*)(st + c.off0);
      const uint32_t pt = (uint32_t)((const int64_t*)c.g0)[cur.e];
      if (cur.old_v >= 0) cnt[(size_t)cur.old_v * c.n0 + pt] -= 1;
      if (cur.new_v >= 0) cnt[(size_t)cur.new_v * c.n0 + pt] += 1;
    } else if (c.kind == SFGPU_K_PROJECT_GROUP) {
      int32_t* gc = (int32_t*)(st + c.off0);
      int64_t* gs = (int64_t*)(st + c.off1);
      const uint32_t* rp = (const uint32_t*)c.g0;
      const longlong2* em = (const longlong2*)c.g1;
      for (uint32_t j = rp[cur.e]; j < rp[cur.e + 1]; ++j) {
        const longlong2 row = em[j];
        if (cur.old_v >= 0) { gc[(uint32_t)cur.old_v * c.n0 + row.x] -= 1; gs[(uint32_t)cur.old_v * c.n0 + row.x] -= row.y; }
        if (cur.new_v >= 0) { gc[(uint32_t)cur.new_v * c.n0 + row.x] += 1; gs[(uint32_t)cur.new_v * c.n0 + row.x] += row.y; }
      }
    } else if (c.kind == SFGPU_K_LOAD_BALANCE) {
      int64_t* loads = (int64_t*)(st + c.off0);
      int32_t* icnt = (int32_t*)(st + c.off1);
      int64_t* agg = (int64_t*)(st + c.off2);
      int64_t x = c.g0 ? ((const int64_t*)c.g0)[cur.e] : 1;
      if (x == 0) continue;
      if (cur.old_v >= 0) {
        int64_t l = loads[cur.old_v];
        icnt[cur.old_v] -= 1;
        int64_t nl = icnt[cur.old_v] == 0 ? 0 : l - x;
        if (icnt[cur.old_v] == 0) agg[2] -= 1;
        agg[1] += nl * nl - l * l;
        agg[0] += nl - l;
        loads[cur.old_v] = nl;
      }
      if (cur.new_v >= 0) {
        int64_t l = loads[cur.new_v];
        if (icnt[cur.new_v] == 0) agg[2] += 1;
        icnt[cur.new_v] += 1;
        int64_t nl = l + x;
        agg[1] += nl * nl - l * l;
        agg[0] += x;
        loads[cur.new_v] = nl;
      }
    }
  }
  var[cur.e] = cur.new_v;
}

// rows: one per replica (mask / index select). kind 0 change, 1 swap.
__global__ void apply_scalar_kernel(const __grid_constant__ DevModel m, int kind, const uint32_t* __restrict__ rows,
                                    const uint8_t* __restrict__ mask, const uint64_t* __restrict__ cand_offsets,
                                    const uint32_t* __restrict__ index, const int32_t* __restrict__ kinds = nullptr) {
  const uint32_t r = blockIdx.x;
  if (threadIdx.x != 0) return;
  if (mask && !mask[r]) return;
  uint64_t ri = r;
  if (index) {
    if (index[r] == 0xFFFFFFFFu) return;
    ri = cand_offsets[r] + index[r];
  }
  if (kinds) {  // union of neighbourhoods: one move kind per replica, rows of 4 words; list kinds go to apply_list_kernel
    kind = kinds[r];
    if (kind < 0 || kind > 1) return;
    rows += (size_t)r * 4;
    ri = 0;
  }
  char* st = m.state + (size_t)r * m.block_bytes;
  int64_t* cs = (int64_t*)(st + m.off_score);
  Score2 d;
  bool ok = kind == 0 ? score_scalar_candidate<MODE_CHANGE>(m, st, st, rows, nullptr, ri, d)
                      : score_scalar_candidate<MODE_SWAP>(m, st, st, rows, nullptr, ri, d);
  if (!ok) return;
  const int32_t* var = (const int32_t*)(st + m.off_var);
  uint2 row = ((const uint2*)rows)[ri];
  if (kind == 0) {
    int32_t nv = (int32_t)row.y < 0 ? SFGPU_NONE : (int32_t)row.y;
    apply_scalar_edit(m, st, EditDev{row.x, var[row.x], nv});
  } else {
    int32_t lv = var[row.x], rv = var[row.y];
    apply_scalar_edit(m, st, EditDev{row.x, lv, rv});
    apply_scalar_edit(m, st, EditDev{row.y, rv, lv});
  }
  cs[0] += d.hard;
  cs[1] += d.soft;
}

// kind 2 list change, 3 list swap, 4 list reverse, 5 sublist change, 6 sublist swap, 7 k-opt. Dynamic smem: elem_cap uint32 (old element copy).
__global__ void __launch_bounds__(256, 4) apply_list_kernel(const __grid_constant__ DevModel m, int kind,
                                                         const uint32_t* __restrict__ rows,
                                                         const uint8_t* __restrict__ mask,
                                                         const uint64_t* __restrict__ cand_offsets,
                                                         const uint32_t* __restrict__ index,
                                                         const int32_t* __restrict__ kinds = nullptr) {
  extern __shared__ __align__(16) uint32_t old_el[];
  __shared__ int s_ok;
  __shared__ uint32_t s_hot[3];  // retained nearby neighbourhood: the moved element and the old last elements of both routes
  const uint32_t r = blockIdx.x;
  if (mask && !mask[r]) return;
  if (kinds) {  // one move kind per replica (union of neighbourhoods); < 0 = no winner, 0 / 1 = scalar moves
    kind = kinds[r];
    if (kind < 2) return;
  }
  uint64_t ri = r;
  if (index) {
    if (index[r] == 0xFFFFFFFFu) return;
    ri = cand_offsets[r] + index[r];
  }
  char* st = m.state + (size_t)r * m.block_bytes;
  uint32_t* off = (uint32_t*)(st + m.off_offsets);
  uint32_t* el = (uint32_t*)(st + m.off_elems);
  const uint4 row = ((const uint4*)rows)[ri];
  const uint32_t total = off[m.n_owners];
  for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) old_el[i] = el[i];
  if (threadIdx.x == 0) {
    Score2 d;
    bool ok = kind == 2 ? list_change_delta(m, st, row, d)
                        : (kind == 3 ? list_swap_delta(m, st, row, d)
                                     : (kind == 4 ? list_reverse_delta(m, st, row, d)
                                                  : (kind == 5 ? list_sublist_change_delta(m, st, row, d)
                                                               : (kind == 6 ? list_sublist_swap_delta(m, st, row, d)
                                                                            : list_k_opt_delta(m, st, row, d)))));
    s_ok = ok ? 1 : 0;
    if (ok && kind == 2 && m.nbc_tag) {
      s_hot[0] = el[off[row.x] + row.y];
      s_hot[1] = off[row.x + 1] > off[row.x] ? el[off[row.x + 1] - 1] : 0xFFFFFFFFu;
      s_hot[2] = off[row.z + 1] > off[row.z] ? el[off[row.z + 1] - 1] : 0xFFFFFFFFu;
    }
    if (ok && (kind == 4 || kind == 7)) {  // a reversal / k-opt keeps every per-route sum
      int64_t* cs = (int64_t*)(st + m.off_score);
      cs[0] += d.hard;
      cs[1] += d.soft;
    } else if (ok) {
      // retained per-route aggregates
      const uint32_t e1 = row.x, p1 = kind >= 5 ? SFGPU_SEG_POS(row.y) : row.y, e2 = row.z, p2 = row.w;
      const uint32_t x1 = el[off[e1] + p1];
      for (uint32_t k = 0; k < m.n_cons; ++k) {
        const ConsDev& c = m.cons[k];
        if (c.kind == SFGPU_K_LIST_SUM && e1 != e2) {
          int64_t* rsum = (int64_t*)(st + c.off0);
          int64_t v1 = ((const int64_t*)c.g0)[x1];
          if (kind == 5) {
            for (uint32_t i = 1; i < SFGPU_SEG_SIZE(row.y); ++i) v1 += ((const int64_t*)c.g0)[el[off[e1] + p1 + i]];
            rsum[e1] -= v1;
            rsum[e2] += v1;
          } else if (kind == 6) {
            for (uint32_t i = 1; i < SFGPU_SEG_SIZE(row.y); ++i) v1 += ((const int64_t*)c.g0)[el[off[e1] + p1 + i]];
            int64_t v2 = 0;
            for (uint32_t i = 0; i < SFGPU_SEG_SIZE(row.w); ++i) v2 += ((const int64_t*)c.g0)[el[off[e2] + SFGPU_SEG_POS(row.w) + i]];
            rsum[e1] += v2 - v1;
            rsum[e2] += v1 - v2;
          } else if (kind == 2) {
            rsum[e1] -= v1;
            rsum[e2] += v1;
          } else {
            int64_t v2 = ((const int64_t*)c.g0)[el[off[e2] + p2]];
            rsum[e1] += v2 - v1;
            rsum[e2] += v1 - v2;
          }
        }
      }
      int64_t* cs = (int64_t*)(st + m.off_score);
      cs[0] += d.hard;
      cs[1] += d.soft;
    }
  }
  __syncthreads();
  if (!s_ok) return;
  if (kind == 4) {
    const uint32_t b = off[row.x];
    for (uint32_t i = row.y + threadIdx.x; i < row.z; i += blockDim.x) el[b + i] = old_el[b + row.y + (row.z - 1 - i)];
  } else if (kind == 7) {
    // k-opt: the route is rebuilt from its segments in the new order (k_opt.rs:40-85)
    KOptDecoded q;
    k_opt_decode(m, off, row, q);
    const uint32_t b = off[q.e], len = q.bounds[q.k + 1];
    for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
      uint32_t at = 0;
      for (uint32_t pos = 0; pos <= q.k; ++pos) {
        const uint32_t sg = q.order[pos], n = q.bounds[sg + 1] - q.bounds[sg];
        if (i < at + n) {
          const uint32_t j = i - at;
          el[b + i] = old_el[b + (((q.rev >> sg) & 1) ? q.bounds[sg + 1] - 1 - j : q.bounds[sg] + j)];
          break;
        }
        at += n;
      }
    }
  } else if (kind == 6) {
    // routes are contiguous in the flat element array, so both the intra- and the inter-list exchange are a swap
    // of two disjoint flat segments: [.. Ea) late [Ea + ne, La) early [La + nl ..)
    const uint32_t A = off[row.x] + SFGPU_SEG_POS(row.y), nA = SFGPU_SEG_SIZE(row.y);
    const uint32_t B = off[row.z] + SFGPU_SEG_POS(row.w), nB = SFGPU_SEG_SIZE(row.w);
    const bool a_early = A < B;
    const uint32_t Ea = a_early ? A : B, ne = a_early ? nA : nB, La = a_early ? B : A, nl = a_early ? nB : nA;
    const uint32_t oe = a_early ? row.x : row.z, ol = a_early ? row.z : row.x;  // owners of the early / late segment
    const uint32_t mid = La - Ea - ne;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
      uint32_t v;
      if (i < Ea || i >= La + nl) v = old_el[i];
      else if (i < Ea + nl) v = old_el[La + (i - Ea)];
      else if (i < Ea + nl + mid) v = old_el[Ea + ne + (i - Ea - nl)];
      else v = old_el[Ea + (i - Ea - nl - mid)];
      el[i] = v;
    }
    __syncthreads();
    for (uint32_t o = threadIdx.x; o <= m.n_owners; o += blockDim.x)
      if (o > oe && o <= ol) off[o] = off[o] + nl - ne;
  } else if (kind == 3) {
    if (threadIdx.x == 0) {
      uint32_t f1 = off[row.x] + row.y, f2 = off[row.z] + row.w;
      el[f1] = old_el[f2];
      el[f2] = old_el[f1];
    }
  } else {
    // relocation of n consecutive elements (n = 1: ListChange, whose destination is given in pre-removal
    // coordinates; a sublist change gives it in post-removal coordinates already)
    const uint32_t se = row.x, sp = kind == 5 ? SFGPU_SEG_POS(row.y) : row.y, de = row.z, dp = row.w;
    const uint32_t n = kind == 5 ? SFGPU_SEG_SIZE(row.y) : 1;
    const uint32_t S = off[se] + sp;
    const uint32_t adj = (kind != 5 && se == de && dp > sp) ? dp - 1 : dp;
    const uint32_t T = off[de] - (de > se ? n : 0) + adj;  // insertion index in the post-removal array
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
      uint32_t v;
      if (i >= T && i < T + n) v = old_el[S + (i - T)];
      else {
        uint32_t j = i >= T + n ? i - n : i;  // index in post-removal array
        v = old_el[j < S ? j : j + n];
      }
      el[i] = v;
    }
    __syncthreads();
    for (uint32_t o = threadIdx.x; o <= m.n_owners; o += blockDim.x) {
      uint32_t v = off[o];
      v = v - (o > se ? n : 0) + (o > de ? n : 0);
      // all reads of off[] above happened before this barrier-separated write phase
      off[o] = v;
    }
  }
  __syncthreads();
  // per-route path costs are recomputed for the (at most two) touched routes: one warp per route, one leg per lane
  for (uint32_t k = 0; k < m.n_cons; ++k) {
    const ConsDev& c = m.cons[k];
    if (c.kind != SFGPU_K_LIST_PATH_COST) continue;
    int64_t* rcost = (int64_t*)(st + c.off0);
    const uint32_t depot = (uint32_t)c.p0;
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w < 2) {
      const bool one_route = kind == 4 || kind == 7;
      const uint32_t o = w == 0 || one_route ? (kind == 7 ? (row.x & 0x0FFFFFFFu) : row.x) : row.z;
      if (!(w == 1 && (row.x == row.z || one_route))) {
        const uint32_t b = off[o], e = off[o + 1];
        int64_t cost = 0;
        if (e > b)  // legs: depot -> el[b], el[i] -> el[i + 1], el[e - 1] -> depot
          for (uint32_t i = b + lane; i <= e; i += 32) cost += mat_at(c, i > b ? el[i - 1] : depot, i < e ? el[i] : depot);
        for (int s = 16; s > 0; s >>= 1) cost += __shfl_down_sync(0xffffffffu, cost, s);
        if (lane == 0) rcost[o] = cost;
      }
    }
  }
  if (m.nbc_tag && threadIdx.x == 0) {
    // retained nearby neighbourhood (sfgpu_nearby.cuh): one ListChange commit right after a cached step is recorded —
    // the two routes, the moved element and every element whose append slot appeared or vanished; anything else
    // (another move kind, a second commit without a step in between) invalidates the cache
    uint32_t* tag = m.nbc_tag + (size_t)r * NBC_WORDS;
    if (kind == 2 && tag[NBC_STATE] == 1) {
      const uint32_t* __restrict__ rl = m.relabel;
      tag[NBC_A] = row.x;
      tag[NBC_B] = row.z;
      tag[NBC_HOT] = rl ? rl[s_hot[0]] : s_hot[0];
      for (uint32_t w = 0; w < 2; ++w) {
        const uint32_t o = w == 0 ? row.x : row.z;
        const uint32_t was = s_hot[1 + w], is = off[o + 1] > off[o] ? el[off[o + 1] - 1] : 0xFFFFFFFFu;
        const bool same = was == is;
        tag[NBC_HOT + 1 + 2 * w] = (same || was == 0xFFFFFFFFu) ? 0xFFFFFFFFu : (rl ? rl[was] : was);
        tag[NBC_HOT + 2 + 2 * w] = (same || is == 0xFFFFFFFFu) ? 0xFFFFFFFFu : (rl ? rl[is] : is);
      }
      tag[NBC_STATE] = 2;
    } else {
      tag[NBC_STATE] = 0;
    }
  }
  if (m.fast_list) {
    __syncthreads();
    // the element copy is dead by now (its first max(elem_cap, n_owners + 1) words become the offsets scratch); the new
    // element array is staged behind it
    uint32_t* s_el = old_el + max(m.elem_cap, m.n_owners + 1);
    const uint32_t n_now = off[m.n_owners];
    for (uint32_t i = threadIdx.x; i < n_now; i += blockDim.x) s_el[i] = el[i];
    build_fast_records(m, st, old_el, s_el);  // its first barrier orders the staging above
  }
}

// =============================================================================================
// argbest: one CTA per replica replays acceptor + forager over the scored rows in pull order.
// =============================================================================================
__device__ __forceinline__ bool accepted_at(const ForageDev& f, const int64_t* scores, const uint8_t* doable,
                                            uint64_t i, int64_t lh, int64_t ls, int64_t th, int64_t ts) {
  if (!doable[i]) return false;
  if (f.acceptor == 0 && !f.gates) return true;
  longlong2 s = ((const longlong2*)scores)[i];
  if (f.gates) {
    // a gated candidate that does not improve on last_step_score is evaluated but never reaches the
    // acceptor (RejectedByHardImprovement / RejectedByScoreImprovement); hard_score_delta of a
    // two-level score is Improving iff the hard level rose (phase/hard_delta.rs:19-34)
    const uint8_t g = f.gates[i];
    if ((g & 1) && !(s.x > lh)) return false;
    if ((g & 2) && !score_less(lh, ls, s.x, s.y)) return false;
  }
  return accept_score(f.acceptor, s.x, s.y, lh, ls, th, ts);
}

__global__ void __launch_bounds__(1024) argbest_kernel(ForageDev f, const uint64_t* __restrict__ cand_offsets,
                                                       const int64_t* __restrict__ scores,
                                                       const uint8_t* __restrict__ doable,
                                                       const uint64_t* __restrict__ step_seeds,
                                                       const int64_t* __restrict__ ref_scores,
                                                       uint32_t* __restrict__ out_index,
                                                       int64_t* __restrict__ out_best,
                                                       uint32_t* __restrict__ out_evaluated,
                                                       const uint32_t* __restrict__ counts = nullptr,
                                                       const uint32_t* __restrict__ skip = nullptr) {
  __shared__ uint32_t scratch[33];
  __shared__ int64_t s_h[32], s_s[32];
  __shared__ uint64_t s_end;
  __shared__ uint32_t s_pick;
  const uint32_t r = blockIdx.x;
  if (skip && skip[r]) return;
  // counts: replica r owns rows [cand_offsets[r], cand_offsets[r] + counts[r]) of a fixed-stride batch
  const uint64_t lo = cand_offsets[r], hi0 = counts ? lo + counts[r] : cand_offsets[r + 1];
  const int64_t lh = ref_scores ? ref_scores[r * 4 + 0] : 0, ls = ref_scores ? ref_scores[r * 4 + 1] : 0;
  const int64_t th = ref_scores ? ref_scores[r * 4 + 2] : 0, ts = ref_scores ? ref_scores[r * 4 + 3] : 0;
  const uint64_t seed = step_seeds ? step_seeds[r] : 0;
  uint64_t hi = hi0;
  // pass 0 (AcceptedCount(N) only): the step stops right after the N-th accepted pull
  if (f.accepted_limit > 0) {
    if (threadIdx.x == 0) s_end = hi0;
    __syncthreads();
    uint32_t seen = 0;
    for (uint64_t base = lo; base < hi0; base += blockDim.x) {
      uint64_t i = base + threadIdx.x;
      uint32_t a = (i < hi0 && accepted_at(f, scores, doable, i, lh, ls, th, ts)) ? 1 : 0;
      uint32_t tot;
      uint32_t incl = block_scan_u32(a, scratch, &tot);
      if (a && seen + incl == f.accepted_limit) s_end = i + 1;
      seen += tot;
      __syncthreads();
      if (seen >= f.accepted_limit) break;
    }
    __syncthreads();
    hi = s_end;
  }
  // pass 1: best accepted score
  int64_t bh = INT64_MIN, bs = INT64_MIN;
  uint32_t cnt = 0;
  for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    if (!accepted_at(f, scores, doable, i, lh, ls, th, ts)) continue;
    longlong2 s = ((const longlong2*)scores)[i];
    if (cnt == 0 || score_less(bh, bs, s.x, s.y)) {
      bh = s.x;
      bs = s.y;
    }
    cnt = 1;
  }
  // warp + block lexicographic max (threads without a candidate carry INT64_MIN pairs)
  for (int o = 16; o > 0; o >>= 1) {
    int64_t oh = __shfl_down_sync(0xffffffffu, bh, o), os = __shfl_down_sync(0xffffffffu, bs, o);
    uint32_t oc = __shfl_down_sync(0xffffffffu, cnt, o);
    if (oc && (!cnt || score_less(bh, bs, oh, os))) {
      bh = oh;
      bs = os;
    }
    cnt |= oc;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_h[warp] = bh;
    s_s[warp] = bs;
    scratch[warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    int64_t h = 0, s = 0;
    uint32_t any = 0;
    for (int w = 0; w < nw; ++w)
      if (scratch[w] && (!any || score_less(h, s, s_h[w], s_s[w]))) {
        h = s_h[w];
        s = s_s[w];
        any = 1;
      }
    s_h[0] = h;
    s_s[0] = s;
    s_pick = any;
  }
  __syncthreads();
  const int64_t mh = s_h[0], ms = s_s[0];
  const uint32_t any = s_pick;
  __syncthreads();
  if (out_evaluated && threadIdx.x == 0) out_evaluated[r] = (uint32_t)(hi - lo);
  if (!any) {
    if (threadIdx.x == 0) {
      out_index[r] = 0xFFFFFFFFu;
      out_best[r * 2] = 0;
      out_best[r * 2 + 1] = 0;
    }
    return;
  }
  // pass 2: number of accepted rows equal to the best score
  uint32_t eq = 0;
  for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    if (!accepted_at(f, scores, doable, i, lh, ls, th, ts)) continue;
    longlong2 s = ((const longlong2*)scores)[i];
    eq += (s.x == mh && s.y == ms) ? 1 : 0;
  }
  uint32_t m_total;
  block_scan_u32(eq, scratch, &m_total);
  // which occurrence wins: First => 1; reservoir => the largest k in [1, m] with pick(seed, k)
  // (BestCandidate::consider replaces the selection at every k whose reservoir_pick fires)
  if (threadIdx.x == 0) s_pick = 1;
  __syncthreads();
  if (f.tie_mode == 1) {
    uint32_t best_k = 1;
    for (uint32_t k = 2 + threadIdx.x; k <= m_total; k += blockDim.x) {
      uint64_t mixed = splitmix64_dev(seed ^ ((uint64_t)k * 0x9E3779B97F4A7C15ull) ^ 0xF04A63E239B74D11ull);
      if (mixed % k == 0) best_k = k;
    }
    atomicMax(&s_pick, best_k);
    __syncthreads();
  }
  const uint32_t want = s_pick;
  __syncthreads();
  // pass 3: locate the want-th occurrence in pull order
  uint32_t seen = 0;
  for (uint64_t base = lo; base < hi; base += blockDim.x) {
    uint64_t i = base + threadIdx.x;
    uint32_t a = 0;
    if (i < hi && accepted_at(f, scores, doable, i, lh, ls, th, ts)) {
      longlong2 s = ((const longlong2*)scores)[i];
      a = (s.x == mh && s.y == ms) ? 1 : 0;
    }
    uint32_t tot;
    uint32_t incl = block_scan_u32(a, scratch, &tot);
    if (a && seen + incl == want) {
      out_index[r] = (uint32_t)(i - lo);
      out_best[r * 2] = mh;
      out_best[r * 2 + 1] = ms;
    }
    seen += tot;
    if (seen >= want) break;
  }
}

__global__ void pack_keys_kernel(const __grid_constant__ DevModel m, int64_t* __restrict__ out_keys) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m.R) return;
  const int64_t* cs = (const int64_t*)(m.state + (size_t)r * m.block_bytes + m.off_score);
  int64_t h = cs[0], s = cs[1];
  const int64_t HB = (int64_t)1 << 22, SB = (int64_t)1 << 39;  // 23 + 40 = 63 bits: stays positive in int64
  h = h < -HB ? -HB : (h >= HB ? HB - 1 : h);
  s = s < -SB ? -SB : (s >= SB ? SB - 1 : s);
  out_keys[r] = (int64_t)(((uint64_t)(h + HB) << 40) | (uint64_t)(s + SB));
}

