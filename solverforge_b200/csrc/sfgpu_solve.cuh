// Device-resident local-search loop for nearby list-change models: every step (seed, acceptor
// reference scores, neighbourhood, scoring, forager, commit, acceptor bookkeeping, best-solution
// tracking) runs on the GPU with no host round trip; steps are captured in a CUDA graph.
//
// Reference loop: solve_local_search_with_resources — solverforge-solver/src/phase/localsearch/phase.rs:237-320,
// execute_step phase/step.rs:30-225 (seed -> step_started -> candidates -> pick -> apply ->
// update_best_solution -> acceptor.step_ended). Acceptors: hill_climbing.rs:33-42,
// late_acceptance.rs:89-126 (history written every step, also when nothing was accepted).
// The reference draws step seeds from rand::StdRng (third party, unpinned — SURVEY §0.1-6); here
// step t of replica r uses splitmix64(seed_base ^ r * 0x9E3779B97F4A7C15 ^ t), stated in the API.
#pragma once
#include "sfgpu_nearby.cuh"

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
  uint64_t seed_base;
  uint32_t late_size;
  int32_t acceptor;
};

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
  s.ref_scores[r * 4 + 2] = s.history[h];
  s.ref_scores[r * 4 + 3] = s.history[h + 1];
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
