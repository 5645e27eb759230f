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
};

// forage acceptor code (sfgpu_forage_params) that one step of the solve-level acceptor reduces to
__host__ __device__ inline int solve_forage_code(int acceptor) {
  switch (acceptor) {
    case 1: return 1;
    case 2: return 2;
    case 3: return 3;  // GreatDeluge: > last || >= water
    case 4: return 3;  // StepCounting: > last || >= (-inf | +inf)
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
  for (uint32_t i = threadIdx.x; i < m.block_bytes / 16; i += blockDim.x)
    ((uint4*)st)[i] = ((const uint4*)(s.best_state + (size_t)r * m.block_bytes))[i];
}
