// Device-resident local-search loop for nearby list-change models: every step (seed, acceptor
// reference scores, neighbourhood, scoring, forager, commit, acceptor bookkeeping, best-solution
// tracking) runs on the GPU with no host round trip; steps are captured in a CUDA graph.
//
// Reference loop: solve_local_search_with_resources — solverforge-solver/src/phase/localsearch/phase.rs:237-320,
// execute_step phase/step.rs:30-225 (seed -> step_started -> candidates -> pick -> apply ->
// update_best_solution -> acceptor.step_ended). Acceptors: hill_climbing.rs:33-42,
// late_acceptance.rs:89-126 (history written every step, also when nothing was accepted),
// great_deluge.rs, step_counting.rs, diversified_late_acceptance.rs (state per replica in acc_state).
// The reference draws step seeds from rand::StdRng (third party, unpinned — SURVEY §0.1-6); here
// step t of replica r uses splitmix64(seed_base ^ r * 0x9E3779B97F4A7C15 ^ t), stated in the API.
#pragma once
#include "sfgpu_dev.cuh"

struct SolveState {
  uint64_t* step_counter;   // [1] steps executed so far
  uint64_t* step_seeds;     // [R]
  int64_t* ref_scores;      // [R][4] {last_step, late}
  int64_t* history;         // [R][late_size] x 2
  uint32_t* hist_idx;       // [R]
  int64_t* best_scores;     // [R][2]
  char* best_state;         // [R][block_bytes] snapshot of the best solution of each replica
  uint64_t* evaluated;      // [R] moves_evaluated accumulated
  uint64_t* accepted_steps; // [R] steps that committed a move
  uint32_t* out_index;      // [R] per-step outputs of the finish kernel
  int64_t* out_best;
  uint32_t* out_evaluated;
  uint32_t* winner_rows;    // [R][4]
  int64_t* acc_state;       // [R][4]: GreatDeluge {water h, s, increment h, s}; StepCounting {steps_since, ...}
  uint64_t seed_base;
  uint32_t late_size;
  int32_t acceptor;         // sfgpu_solve_params.acceptor (1..5)
  double real;              // rain_speed / tolerance
  uint64_t step_count_limit;
  SaState* sa_cur;          // [R] state at the start of the step
  SaState* sa_nxt;          // [R] state after the step's evaluated candidates (committed by solve_post_kernel)
  SaParams sa;
  struct TabuState* tabu;   // [R] TabuSearch memories (acceptor 7)
  uint32_t tabu_tenure[4];  // entity, value, move, undo-move tenures (0 = dimension off)
  uint32_t tabu_aspiration;
};

// TabuSearchAcceptor memories of one replica (tabu_search.rs:64-101: FIFO of at most `tenure` entries; membership does
// not depend on the order, so each memory is a ring: slot = (head + k) % tenure). Entries of the scalar ChangeMove
// (heuristic/move/change.rs:189-220): entity id, destination value id (0xFFFFFFFF = None), move id (entity, from, to),
// undo move id (entity, to, from); the scope (descriptor, variable) is the same for every move of the loop. Entries of
// ListChangeMove (list_kernel/change.rs:155-204): source and destination entity, moved element, move id (source entity,
// source position, destination entity, adjusted destination position, element), undo id (the same with both ends swapped).
#define TABU_CAP 64
struct TabuState {
  uint32_t n[4], head[4];
  uint32_t ent[TABU_CAP], val[TABU_CAP];
  uint32_t mov[TABU_CAP][5], und[TABU_CAP][5];  // scalar moves use the first three words
};

// forage acceptor code (sfgpu_forage_params) that one step of the solve-level acceptor reduces to
__host__ __device__ inline int solve_forage_code(int acceptor) {
  switch (acceptor) {
    case 1: return 1;
    case 2: return 2;
    case 3: return 3;  // GreatDeluge: > last || >= water
    case 4: return 3;  // StepCounting: > last || >= (-inf | +inf)
    case 6: return 0;  // SimulatedAnnealing: sa_accept_kernel leaves the accepted candidates as the doable ones
    case 7: return 0;  // TabuSearch: tabu_accept_kernel does the same
    default: return 2; // DiversifiedLateAcceptance: >= last || >= min(late, best - |best| * tol)
  }
}
// Score::multiply (macros.rs:61-63): (x as f64 * m).round() as i64, half away from zero
__device__ __forceinline__ int64_t mul_round_dev(int64_t x, double m) { return (int64_t)round(__dmul_rn((double)x, m)); }
__device__ __forceinline__ int64_t abs64_dev(int64_t x) { return x < 0 ? -x : x; }
#define SOLVE_INF (1ll << 62)

// once per solve: acceptor.phase_started (history filled with the initial score), best = initial
__global__ void solve_init_kernel(const __grid_constant__ DevModel m, SolveState s) {
  const uint32_t r = blockIdx.x;
  const char* st = m.state + (size_t)r * m.block_bytes;
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  for (uint32_t i = threadIdx.x; i < s.late_size; i += blockDim.x) {
    s.history[((size_t)r * s.late_size + i) * 2] = cs[0];
    s.history[((size_t)r * s.late_size + i) * 2 + 1] = cs[1];
  }
  for (uint32_t i = threadIdx.x; i < m.block_bytes / 16; i += blockDim.x)
    ((uint4*)(s.best_state + (size_t)r * m.block_bytes))[i] = ((const uint4*)st)[i];
  if (threadIdx.x == 0) {
    s.hist_idx[r] = 0;
    if (s.acceptor == 3) {  // great_deluge.rs:70-73: water = initial, increment = |initial| * rain_speed
      s.acc_state[r * 4 + 0] = cs[0];
      s.acc_state[r * 4 + 1] = cs[1];
      s.acc_state[r * 4 + 2] = mul_round_dev(abs64_dev(cs[0]), s.real);
      s.acc_state[r * 4 + 3] = mul_round_dev(abs64_dev(cs[1]), s.real);
    } else {
      s.acc_state[r * 4 + 0] = 0;  // step_counting.rs:69-72: steps_since_improvement
    }
    if (s.acceptor == 6) {
      SaState z{};
      s.sa_cur[r] = z;
      s.sa_nxt[r] = z;
    }
    if (s.acceptor == 7)  // tabu_search.rs:188-194 phase_started: memories cleared, best = initial
      for (int q = 0; q < 4; ++q) s.tabu[r].n[q] = s.tabu[r].head[q] = 0;
    s.best_scores[r * 2] = cs[0];
    s.best_scores[r * 2 + 1] = cs[1];
    s.evaluated[r] = 0;
    s.accepted_steps[r] = 0;
    if (r == 0) *s.step_counter = 0;
  }
}

// before the neighbourhood: step seed + the scores the acceptor compares against
__global__ void solve_prep_kernel(const __grid_constant__ DevModel m, SolveState s) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m.R) return;
  const uint64_t t = *s.step_counter;
  s.step_seeds[r] = splitmix64_dev(s.seed_base ^ ((uint64_t)r * 0x9E3779B97F4A7C15ull) ^ t);
  const int64_t* cs = (const int64_t*)(m.state + (size_t)r * m.block_bytes + m.off_score);
  s.ref_scores[r * 4 + 0] = cs[0];
  s.ref_scores[r * 4 + 1] = cs[1];
  const size_t h = ((size_t)r * s.late_size + s.hist_idx[r]) * 2;
  int64_t th = s.history[h], ts = s.history[h + 1];
  if (s.acceptor == 3) {
    th = s.acc_state[r * 4 + 0];
    ts = s.acc_state[r * 4 + 1];
  } else if (s.acceptor == 4) {  // step_counting.rs:56-67: everything passes while under the limit
    th = ts = (uint64_t)s.acc_state[r * 4 + 0] < s.step_count_limit ? -SOLVE_INF : SOLVE_INF;
  } else if (s.acceptor == 5) {  // diversified_late_acceptance.rs:91-98: best - |best| * tolerance
    const int64_t bh = s.best_scores[r * 2], bs = s.best_scores[r * 2 + 1];
    const int64_t dh = bh - mul_round_dev(abs64_dev(bh), s.real), ds = bs - mul_round_dev(abs64_dev(bs), s.real);
    if (score_less(dh, ds, th, ts)) {
      th = dh;
      ts = ds;
    }
  }
  s.ref_scores[r * 4 + 2] = th;
  s.ref_scores[r * 4 + 3] = ts;
}

// after the commit: acceptor.step_ended, statistics, best-solution snapshot
__global__ void __launch_bounds__(256) solve_post_kernel(const __grid_constant__ DevModel m, SolveState s) {
  __shared__ int improved;
  const uint32_t r = blockIdx.x;
  const char* st = m.state + (size_t)r * m.block_bytes;
  const int64_t* cs = (const int64_t*)(st + m.off_score);
  if (threadIdx.x == 0) {
    const size_t h = ((size_t)r * s.late_size + s.hist_idx[r]) * 2;
    s.history[h] = cs[0];  // late_acceptance.rs:116-126: written every step
    s.history[h + 1] = cs[1];
    s.hist_idx[r] = (s.hist_idx[r] + 1) % s.late_size;
    s.evaluated[r] += s.out_evaluated[r];
    if (s.out_index[r] != 0xFFFFFFFFu) s.accepted_steps[r] += 1;
    improved = score_less(s.best_scores[r * 2], s.best_scores[r * 2 + 1], cs[0], cs[1]) ? 1 : 0;
    if (s.acceptor == 3) {  // great_deluge.rs:76-85: the water rises every step
      s.acc_state[r * 4 + 0] += s.acc_state[r * 4 + 2];
      s.acc_state[r * 4 + 1] += s.acc_state[r * 4 + 3];
    } else if (s.acceptor == 4) {  // step_counting.rs:74-90 (its best score == the phase's best score)
      s.acc_state[r * 4 + 0] = improved ? 0 : s.acc_state[r * 4 + 0] + 1;
    } else if (s.acceptor == 6) {  // simulated_annealing.rs:416-431: no decay while calibrating
      SaState z = s.sa_nxt[r];
      if (z.calibrated)
        for (int l = 0; l < 2; ++l) {
          z.temp[l] = __dmul_rn(z.temp[l], s.sa.decay);
          if (z.temp[l] < s.sa.hc_temp) z.temp[l] = s.sa.hc_temp;
        }
      s.sa_cur[r] = z;
    }
    if (improved) {
      s.best_scores[r * 2] = cs[0];
      s.best_scores[r * 2 + 1] = cs[1];
    }
    if (r == 0) *s.step_counter += 1;
  }
  __syncthreads();
  if (improved)
    for (uint32_t i = threadIdx.x; i < m.block_bytes / 16; i += blockDim.x)
      ((uint4*)(s.best_state + (size_t)r * m.block_bytes))[i] = ((const uint4*)st)[i];
}

// optional at the end: working solution := best solution
__global__ void solve_restore_best_kernel(const __grid_constant__ DevModel m, SolveState s) {
  const uint32_t r = blockIdx.x;
  char* st = m.state + (size_t)r * m.block_bytes;
  if (threadIdx.x == 0) nbc_invalidate(m.nbc_tag, r);
  for (uint32_t i = threadIdx.x; i < m.block_bytes / 16; i += blockDim.x)
    ((uint4*)st)[i] = ((const uint4*)(s.best_state + (size_t)r * m.block_bytes))[i];
}


// inclusive block scan of a 32-bit count (same contract as block_scan_u32 of sfgpu_kernels.cuh)
__device__ __forceinline__ uint32_t solve_scan_u32(uint32_t v, uint32_t* scratch /* >= 33 */, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t x = v;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  __syncthreads();
  if (lane == 31) scratch[warp] = x;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  if (warp == 0) {
    uint32_t w = lane < nw ? scratch[lane] : 0;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    scratch[lane] = w;
  }
  __syncthreads();
  const uint32_t base = warp > 0 ? scratch[warp - 1] : 0;
  *total = scratch[nw - 1];
  __syncthreads();
  return x + base;
}

// uniform stream of the SimulatedAnnealing acceptor. The reference draws from rand::SmallRng (third party, unpinned);
// here draw j of a step is a pure function of the step seed, stated in the API, and only worsening candidates whose
// level temperature is above the hill-climbing temperature consume a draw — exactly where the reference draws.
__device__ __forceinline__ double sa_uniform(uint64_t step_seed, uint32_t j) {
  const uint64_t x = splitmix64_dev(step_seed ^ 0x5A17EA11EA1DF00Dull ^ ((uint64_t)j * 0x9E3779B97F4A7C15ull));
  return (double)(x >> 11) * 0x1.0p-53;
}

// One CTA per replica over its materialised, pull-ordered scores: replays SimulatedAnnealingAcceptor::is_accepted in
// pull order up to the AcceptedCount cut and leaves doable[i] = accepted (argbest_kernel then runs with acceptor 0).
// counts: rows of replica r (fixed-stride batch) or null (cand_offsets[r + 1]); skip[r] != 0: nothing to do.
__global__ void __launch_bounds__(1024) sa_accept_kernel(const uint64_t* __restrict__ cand_offsets,
                                                         const uint32_t* __restrict__ counts,
                                                         const uint32_t* __restrict__ skip,
                                                         const int64_t* __restrict__ scores, uint8_t* __restrict__ doable,
                                                         const int64_t* __restrict__ ref_scores,
                                                         const uint64_t* __restrict__ step_seeds, const SaState* cur,
                                                         SaState* nxt, const SaParams p, const uint32_t accepted_limit) {
  __shared__ uint32_t scratch[33];
  __shared__ SaState z;
  __shared__ uint32_t s_cut, s_cal;
  __shared__ unsigned long long s_sum[2];
  __shared__ uint32_t s_cnt[2];
  const uint32_t r = blockIdx.x;
  if (skip && skip[r]) return;
  const uint64_t lo = cand_offsets[r], hi = counts ? lo + counts[r] : cand_offsets[r + 1];
  const int64_t lh = ref_scores[r * 4 + 0], ls = ref_scores[r * 4 + 1];
  const uint64_t seed = step_seeds ? step_seeds[r] : 0;
  if (threadIdx.x == 0) z = cur[r];
  __syncthreads();
  uint32_t accepted_before = 0, draws_before = 0;
  for (uint64_t base = lo; base < hi; base += blockDim.x) {
    const uint64_t i = base + threadIdx.x;
    const bool in = i < hi;
    bool good = false, worse = false;
    int lvl = 0;
    int64_t delta = 0;
    if (in && doable[i]) {
      const longlong2 sc = ((const longlong2*)scores)[i];
      if (!score_less(sc.x, sc.y, lh, ls)) good = true;  // move_score >= last_step_score
      else {
        lvl = sc.x != lh ? 0 : 1;
        delta = lvl == 0 ? sc.x - lh : sc.y - ls;  // < 0
        worse = !(p.never_hard && lvl == 0);       // a barred hard regression is rejected before it is sampled
      }
    }
    uint32_t from = 0;  // positions of this chunk below `from` are settled (calibrating mode decided them)
    bool acc = good;
    if (!z.calibrated) {
      uint32_t g_tot, w_tot;
      const uint32_t g_incl = solve_scan_u32(good ? 1u : 0u, scratch, &g_tot);
      const uint32_t w_incl = solve_scan_u32(worse ? 1u : 0u, scratch, &w_tot);
      if (threadIdx.x == 0) {
        s_cut = 0xFFFFFFFFu;
        s_cal = 0xFFFFFFFFu;
        s_sum[0] = s_sum[1] = 0;
        s_cnt[0] = s_cnt[1] = 0;
      }
      __syncthreads();
      const uint32_t seen = z.cnt[0] + z.cnt[1];
      if (accepted_limit && good && accepted_before + g_incl == accepted_limit) s_cut = threadIdx.x;
      if (worse && seen + w_incl == p.sample_size) s_cal = threadIdx.x;
      __syncthreads();
      const uint32_t cut = s_cut, cal = s_cal;
      const uint32_t last = cut < cal ? cut : cal;  // samples are recorded up to here (inclusive)
      if (worse && threadIdx.x <= last) {
        atomicAdd(&s_sum[lvl], (unsigned long long)(-delta));
        atomicAdd(&s_cnt[lvl], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int l = 0; l < 2; ++l) {
          z.sum[l] += (int64_t)s_sum[l];
          z.cnt[l] += s_cnt[l];
        }
        if (cal < cut) {  // CalibrationState::temperatures
          const double denom = p.neg_log_target;
          for (int l = 0; l < 2; ++l) {
            double t = p.fallback;
            if (z.cnt[l]) {
              t = ((double)z.sum[l] / (double)z.cnt[l]) / denom;
              if (t < p.fallback) t = p.fallback;
            }
            z.temp[l] = t;
          }
          z.calibrated = 1;
        }
      }
      __syncthreads();
      if (cut < cal) {  // the forager quits inside the calibrating part
        if (in && threadIdx.x <= cut) doable[i] = acc ? 1 : 0;
        break;
      }
      if (cal == 0xFFFFFFFFu) {  // still calibrating after this chunk: worsening candidates are rejected
        if (in) doable[i] = acc ? 1 : 0;
        accepted_before += g_tot;
        continue;
      }
      // calibration completed at `cal`: that candidate and the later ones of the chunk go through the Boltzmann test
      if (in && threadIdx.x < cal) doable[i] = acc ? 1 : 0;
      uint32_t gb_tot;
      solve_scan_u32((good && threadIdx.x < cal) ? 1u : 0u, scratch, &gb_tot);
      accepted_before += gb_tot;
      from = cal;
    }
    // calibrated: worsening candidates above the hill-climbing temperature draw in pull order
    const bool active = threadIdx.x >= from;
    const double T = z.temp[lvl];
    const bool draws = active && worse && T > p.hc_temp;
    uint32_t d_tot;
    const uint32_t d_incl = solve_scan_u32(draws ? 1u : 0u, scratch, &d_tot);
    if (draws) acc = sa_uniform(seed, draws_before + d_incl - 1) < exp((double)delta / T);
    uint32_t a_tot;
    const uint32_t a_incl = solve_scan_u32((active && acc) ? 1u : 0u, scratch, &a_tot);
    if (threadIdx.x == 0) s_cut = 0xFFFFFFFFu;
    __syncthreads();
    if (accepted_limit && active && acc && accepted_before + a_incl == accepted_limit) s_cut = threadIdx.x;
    __syncthreads();
    const uint32_t cut = s_cut;
    if (in && active && threadIdx.x <= cut) doable[i] = acc ? 1 : 0;
    if (cut != 0xFFFFFFFFu) break;
    accepted_before += a_tot;
    draws_before += d_tot;
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) nxt[r] = z;
}


// ---- TabuSearch (acceptor/tabu_search.rs:103-237) over the materialised ChangeMove batch of a scalar step ---------
// is_accepted = aspirational (aspiration enabled and move score > best score) or not tabu; tabu = the move's entity is
// in the entity memory, or its destination value in the value memory, or its move id in the move memory or in the
// undo-move memory (:150-186). One CTA per replica; leaves doable[i] = accepted (argbest then runs with acceptor 0).
__global__ void __launch_bounds__(256) tabu_accept_kernel(const __grid_constant__ DevModel m, SolveState s,
                                                          const uint64_t* __restrict__ cand_offsets,
                                                          const uint32_t* __restrict__ counts,
                                                          const uint32_t* __restrict__ rows,
                                                          const int64_t* __restrict__ scores, uint8_t* __restrict__ doable) {
  __shared__ TabuState z;
  const uint32_t r = blockIdx.x;
  for (uint32_t i = threadIdx.x; i < sizeof(TabuState) / 4; i += blockDim.x) ((uint32_t*)&z)[i] = ((const uint32_t*)(s.tabu + r))[i];
  __syncthreads();
  const int32_t* var = (const int32_t*)(m.state + (size_t)r * m.block_bytes + m.off_var);
  const int64_t bh = s.best_scores[r * 2], bs = s.best_scores[r * 2 + 1];
  const uint64_t lo = cand_offsets[r], hi = lo + counts[r];
  for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    if (!doable[i]) continue;
    const uint2 row = ((const uint2*)rows)[i];
    const uint32_t e = row.x, to = (int32_t)row.y < 0 ? 0xFFFFFFFFu : row.y;
    const int32_t fv = var[e];
    const uint32_t from = fv < 0 ? 0xFFFFFFFFu : (uint32_t)fv;
    bool tabu = false;
    for (uint32_t k = 0; k < z.n[0]; ++k) tabu |= z.ent[k] == e;
    for (uint32_t k = 0; k < z.n[1]; ++k) tabu |= z.val[k] == to;
    for (uint32_t k = 0; k < z.n[2]; ++k) tabu |= z.mov[k][0] == e && z.mov[k][1] == from && z.mov[k][2] == to;
    for (uint32_t k = 0; k < z.n[3]; ++k) tabu |= z.und[k][0] == e && z.und[k][1] == from && z.und[k][2] == to;
    if (tabu) {
      const longlong2 sc = ((const longlong2*)scores)[i];
      const bool aspirational = s.tabu_aspiration && score_less(bh, bs, sc.x, sc.y);
      if (!aspirational) doable[i] = 0;
    }
  }
}

// step_ended (:204-236): the accepted move's entity, destination value, move id and undo move id enter the memories
// (before the commit kernel: `from` is still the committed value). One thread per replica.
__global__ void tabu_record_kernel(const __grid_constant__ DevModel m, SolveState s, const uint64_t* __restrict__ cand_offsets,
                                   const uint32_t* __restrict__ rows) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m.R) return;
  const uint32_t idx = s.out_index[r];
  if (idx == 0xFFFFFFFFu) return;
  const uint2 row = ((const uint2*)rows)[cand_offsets[r] + idx];
  const int32_t* var = (const int32_t*)(m.state + (size_t)r * m.block_bytes + m.off_var);
  const uint32_t e = row.x, to = (int32_t)row.y < 0 ? 0xFFFFFFFFu : row.y;
  const int32_t fv = var[e];
  const uint32_t from = fv < 0 ? 0xFFFFFFFFu : (uint32_t)fv;
  TabuState& z = s.tabu[r];
  auto slot = [&](int q) -> uint32_t {  // FIFO: the oldest entry leaves when the memory is full
    const uint32_t t = s.tabu_tenure[q];
    if (z.n[q] < t) return z.n[q]++;
    const uint32_t at = z.head[q];
    z.head[q] = (at + 1) % t;
    return at;
  };
  if (s.tabu_tenure[0]) z.ent[slot(0)] = e;
  if (s.tabu_tenure[1]) z.val[slot(1)] = to;
  if (s.tabu_tenure[2]) {
    const uint32_t k = slot(2);
    z.mov[k][0] = e;
    z.mov[k][1] = from;
    z.mov[k][2] = to;
  }
  if (s.tabu_tenure[3]) {
    const uint32_t k = slot(3);
    z.und[k][0] = e;
    z.und[k][1] = to;
    z.und[k][2] = from;
  }
}


// ---- TabuSearch over the materialised nearby ListChange batch of a list step ----------------------------------------
// rows {se, sp, de, dp}: fixed stride of max_nearby rows per source position, sentinel rows (not doable) where a source
// has fewer candidates. A candidate is tabu when its source or destination entity is in the entity memory, the moved
// element in the value memory, or its move id in the move / undo-move memory. Also leaves counts[r] = rows of the routed
// sources (the ordered replay runs over those) and n_per[r] = candidates per source (for moves_evaluated).
__global__ void __launch_bounds__(256) tabu_accept_list_kernel(const __grid_constant__ DevModel m, SolveState s,
                                                               const uint32_t max_nearby, uint32_t* __restrict__ counts,
                                                               uint32_t* __restrict__ n_per,
                                                               const uint32_t* __restrict__ rows,
                                                               const int64_t* __restrict__ scores,
                                                               uint8_t* __restrict__ doable) {
  __shared__ TabuState z;
  const uint32_t r = blockIdx.x;
  for (uint32_t i = threadIdx.x; i < sizeof(TabuState) / 4; i += blockDim.x) ((uint32_t*)&z)[i] = ((const uint32_t*)(s.tabu + r))[i];
  __syncthreads();
  const char* st = m.state + (size_t)r * m.block_bytes;
  const uint32_t* off = (const uint32_t*)(st + m.off_offsets);
  const uint32_t* el = (const uint32_t*)(st + m.off_elems);
  const uint32_t total = off[m.n_owners];
  const size_t lo = (size_t)r * m.elem_cap * max_nearby, hi = lo + (size_t)total * max_nearby;
  if (threadIdx.x == 0) {
    uint32_t slots = 0;  // candidates per source: min(max_nearby, slots of non-empty routes - 2)
    for (uint32_t o = 0; o < m.n_owners; ++o) slots += off[o + 1] > off[o] ? off[o + 1] - off[o] + 1 : 0;
    const uint32_t valid = slots >= 2 ? slots - 2 : 0;
    n_per[r] = valid < max_nearby ? valid : max_nearby;
    counts[r] = total * max_nearby;
  }
  if (s.acceptor != 7) return;  // SimulatedAnnealing only needs the counts: sa_accept_kernel decides
  const int64_t bh = s.best_scores[r * 2], bs = s.best_scores[r * 2 + 1];
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    if (!doable[i]) continue;
    const uint4 row = ((const uint4*)rows)[i];
    const uint32_t se = row.x, sp = row.y, de = row.z, dp = row.w;
    const uint32_t moved = el[off[se] + sp];
    const uint32_t adj = (se == de && dp > sp) ? dp - 1 : dp;
    bool tabu = false;
    for (uint32_t k = 0; k < z.n[0]; ++k) tabu |= z.ent[k] == se || z.ent[k] == de;
    for (uint32_t k = 0; k < z.n[1]; ++k) tabu |= z.val[k] == moved;
    for (uint32_t k = 0; k < z.n[2]; ++k)
      tabu |= z.mov[k][0] == se && z.mov[k][1] == sp && z.mov[k][2] == de && z.mov[k][3] == adj && z.mov[k][4] == moved;
    for (uint32_t k = 0; k < z.n[3]; ++k)
      tabu |= z.und[k][0] == se && z.und[k][1] == sp && z.und[k][2] == de && z.und[k][3] == adj && z.und[k][4] == moved;
    if (tabu) {
      const longlong2 sc = ((const longlong2*)scores)[i];
      if (!(s.tabu_aspiration && score_less(bh, bs, sc.x, sc.y))) doable[i] = 0;
    }
  }
}

// step_ended for the list loop + moves_evaluated in reference pulls (the replay counted padded rows). One thread per replica.
__global__ void tabu_record_list_kernel(const __grid_constant__ DevModel m, SolveState s, const uint32_t max_nearby,
                                        const uint32_t* __restrict__ n_per, const uint32_t* __restrict__ rows) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m.R) return;
  {  // padded position q = f * max_nearby + l  ->  pull f * n_per + min(l, n_per - 1)
    const uint32_t e_pad = s.out_evaluated[r], per = n_per[r];
    if (e_pad && per) {
      const uint32_t f = (e_pad - 1) / max_nearby, l = (e_pad - 1) % max_nearby;
      s.out_evaluated[r] = f * per + (l < per ? l : per - 1) + 1;
    } else {
      s.out_evaluated[r] = 0;
    }
  }
  const uint32_t idx = s.out_index[r];
  if (idx == 0xFFFFFFFFu || s.acceptor != 7) return;
  const uint4 row = ((const uint4*)rows)[(size_t)r * m.elem_cap * max_nearby + idx];
  const char* st = m.state + (size_t)r * m.block_bytes;
  const uint32_t* off = (const uint32_t*)(st + m.off_offsets);
  const uint32_t* el = (const uint32_t*)(st + m.off_elems);
  const uint32_t se = row.x, sp = row.y, de = row.z, dp = row.w;
  const uint32_t moved = el[off[se] + sp];
  const uint32_t adj = (se == de && dp > sp) ? dp - 1 : dp;
  TabuState& z = s.tabu[r];
  auto slot = [&](int q) -> uint32_t {
    const uint32_t t = s.tabu_tenure[q];
    if (z.n[q] < t) return z.n[q]++;
    const uint32_t at = z.head[q];
    z.head[q] = (at + 1) % t;
    return at;
  };
  if (s.tabu_tenure[0]) {  // every entity id of the signature is recorded in turn
    z.ent[slot(0)] = se;
    if (de != se) z.ent[slot(0)] = de;
  }
  if (s.tabu_tenure[1]) z.val[slot(1)] = moved;
  if (s.tabu_tenure[2]) {
    const uint32_t k = slot(2);
    z.mov[k][0] = se; z.mov[k][1] = sp; z.mov[k][2] = de; z.mov[k][3] = adj; z.mov[k][4] = moved;
  }
  if (s.tabu_tenure[3]) {
    const uint32_t k = slot(3);
    z.und[k][0] = de; z.und[k][1] = adj; z.und[k][2] = se; z.und[k][3] = sp; z.und[k][4] = moved;
  }
}
