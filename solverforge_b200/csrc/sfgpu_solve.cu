// Device-resident local-search loop of libsfgpu (sfgpu_solve.cuh): whole steps without a host round trip.
#include "sfgpu_ctx.hpp"
#include "sfgpu_solve.cuh"

using namespace sfgpu_host;

int sfgpu_launch_sa_accept(sfgpu_ctx* ctx, const uint64_t* d_offs, const uint32_t* d_counts, const uint32_t* d_skip,
                           const int64_t* d_scores, uint8_t* d_doable, const int64_t* d_ref, const uint64_t* d_seeds,
                           const SaState* cur, SaState* nxt, const SaParams& p, uint32_t accepted_limit) {
  sa_accept_kernel<<<ctx->dm.R, 1024, 0, ctx->stream>>>(d_offs, d_counts, d_skip, d_scores, d_doable, d_ref, d_seeds, cur, nxt, p,
                                                        accepted_limit);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}

namespace {
int solve_impl(sfgpu_ctx* ctx, const sfgpu_solve_params* p, bool scalar, int64_t* out_best_scores,
               uint64_t* out_moves_evaluated, uint64_t* out_accepted_steps, const sfgpu_union_desc* udesc = nullptr,
               uint64_t* out_window_overflows = nullptr, uint64_t* out_pulls_scored = nullptr) {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (!p) return fail(ctx, SFGPU_E_INVALID, "null params");
  const DevModel& dm = ctx->dm;
  // AcceptedCount(N) consumes a prefix of the cursor: with params->reserved bit 0 the canonical nearby loop runs as a
  // one-child union in SelectionOrder::Original, which generates and scores only the window the forager reaches — same
  // pulls, same winners, same moves_evaluated (tests/test_gpu_union.py). It wins while the forager quits within a few
  // hundred pulls; a replica that needs most of its neighbourhood is faster in the whole-neighbourhood step (many CTAs
  // per replica), which stays the default.
  sfgpu_union_desc windowed{};
  if (!scalar && !udesc && (p->reserved & 1) && p->accepted_limit > 0 && p->max_nearby >= 1 && p->max_nearby <= 32 &&
      dm.nearby_ok && !ctx->force_generic) {
    windowed.n_children = 1;
    windowed.union_order = SFGPU_UNION_SEQUENTIAL;
    windowed.selection_order = SFGPU_ORDER_ORIGINAL;
    windowed.children[0].family = SFGPU_FAM_NEARBY_LIST_CHANGE;
    windowed.children[0].p0 = p->max_nearby;
    windowed.children[0].weight = 1;
    windowed.max_window = std::max<uint32_t>(4096u, dm.elem_cap * p->max_nearby);  // the whole neighbourhood fits
    udesc = &windowed;
  }
  if (scalar) {
    if (!dm.has_scalar) return fail(ctx, SFGPU_E_STATE, "model has no scalar variable");
    if ((uint64_t)dm.n_entities * (dm.n_values + 1) >= 0xFFFFFFFFull)
      return fail(ctx, SFGPU_E_UNSUPPORTED, "neighbourhood too large for 32-bit pull indices");
  } else if (!udesc) {
    if (!dm.nearby_ok || ctx->force_generic)
      return fail(ctx, SFGPU_E_UNSUPPORTED, "device-resident loop needs the fast list program (see sfgpu_step_nearby_list_change)");
    if (p->max_nearby == 0 || p->max_nearby > 32) return fail(ctx, SFGPU_E_UNSUPPORTED, "max_nearby must be in [1, 32]");
  }
  if (p->acceptor < 1 || p->acceptor > 7)
    return fail(ctx, SFGPU_E_INVALID,
                "acceptor: 1 HillClimbing, 2 LateAcceptance, 3 GreatDeluge, 4 StepCountingHillClimbing, "
                "5 DiversifiedLateAcceptance, 6 SimulatedAnnealing, 7 TabuSearch");
  const bool tabu = p->acceptor == 7;
  uint32_t tabu_tenure[4] = {0, 0, 0, 0};
  if (tabu) {
    if (udesc) return fail(ctx, SFGPU_E_UNSUPPORTED, "TabuSearch runs in sfgpu_solve_change and sfgpu_solve_nearby_list_change");
    for (int q = 0; q < 4; ++q) tabu_tenure[q] = (p->late_size >> (8 * q)) & 0xFFu;
    if (!(tabu_tenure[0] | tabu_tenure[1] | tabu_tenure[2] | tabu_tenure[3]))
      return fail(ctx, SFGPU_E_INVALID, "tabu_search requires at least one tabu dimension");  // tabu_search.rs:35-41
    for (int q = 0; q < 4; ++q)
      if (tabu_tenure[q] > TABU_CAP) return fail(ctx, SFGPU_E_UNSUPPORTED, "tabu tenure above 64");
  }
  const bool sa = p->acceptor == 6;
  if (sa && !(p->acceptor_real >= 0.0 && p->acceptor_real <= 1.0))
    return fail(ctx, SFGPU_E_INVALID, "simulated_annealing decay_rate must be finite and in (0, 1] (0 = default)");
  if ((p->acceptor == 3 || p->acceptor == 5) && !(p->acceptor_real >= 0.0 && p->acceptor_real <= 1e6))
    return fail(ctx, SFGPU_E_INVALID, "acceptor_real (rain_speed / tolerance) must be a finite value >= 0");
  if (p->tie_mode < 0 || p->tie_mode > 1) return fail(ctx, SFGPU_E_INVALID, "bad tie_mode");
  CU(cudaSetDevice(ctx->device));
  const uint32_t R = dm.R;
  const uint32_t late = tabu ? 1u : std::max<uint32_t>(p->late_size, 1);
  auto a16 = [](size_t v) { return (v + 15) / 16 * 16; };
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = a16(o + bytes); return at; };
  const size_t o_cnt = take(8), o_seed = take((size_t)R * 8), o_ref = take((size_t)R * 32);
  const size_t o_hist = take((size_t)R * late * 16), o_hidx = take((size_t)R * 4), o_bests = take((size_t)R * 16);
  const size_t o_eval = take((size_t)R * 8), o_acc = take((size_t)R * 8), o_idx = take((size_t)R * 4);
  const size_t o_ob = take((size_t)R * 16), o_oe = take((size_t)R * 4), o_win = take((size_t)R * 16);
  const size_t o_accst = take((size_t)R * 32);
  const size_t o_ovf = take((size_t)R * 16);
  const size_t o_sa = take((size_t)R * sizeof(SaState) * 2), o_cnts = take((size_t)R * 4);
  const size_t o_tabu = take(tabu ? (size_t)R * sizeof(TabuState) : 16);
  const size_t o_snap = take((size_t)R * dm.block_bytes);
  if (o > ctx->solve_bytes) {
    if (ctx->solve_buf) cudaFree(ctx->solve_buf);
    ctx->solve_buf = nullptr;
    ctx->solve_bytes = 0;
    CU(cudaMalloc(&ctx->solve_buf, o));
    ctx->solve_bytes = o;
  }
  char* b = (char*)ctx->solve_buf;
  SolveState s{};
  s.step_counter = (uint64_t*)(b + o_cnt);
  s.step_seeds = (uint64_t*)(b + o_seed);
  s.ref_scores = (int64_t*)(b + o_ref);
  s.history = (int64_t*)(b + o_hist);
  s.hist_idx = (uint32_t*)(b + o_hidx);
  s.best_scores = (int64_t*)(b + o_bests);
  s.evaluated = (uint64_t*)(b + o_eval);
  s.accepted_steps = (uint64_t*)(b + o_acc);
  s.out_index = (uint32_t*)(b + o_idx);
  s.out_best = (int64_t*)(b + o_ob);
  s.out_evaluated = (uint32_t*)(b + o_oe);
  s.winner_rows = (uint32_t*)(b + o_win);
  s.best_state = b + o_snap;
  s.seed_base = p->seed_base;
  s.late_size = late;
  s.acceptor = p->acceptor;
  s.acc_state = (int64_t*)(b + o_accst);
  s.real = p->acceptor_real;
  s.step_count_limit = p->step_count_limit;
  s.sa_cur = (SaState*)(b + o_sa);
  s.sa_nxt = s.sa_cur + R;
  s.tabu = (TabuState*)(b + o_tabu);
  if (tabu) CU(cudaMemsetAsync(s.tabu, 0, (size_t)R * sizeof(TabuState), ctx->stream));  // the kernels stage whole memories
  for (int q = 0; q < 4; ++q) s.tabu_tenure[q] = tabu_tenure[q];
  s.tabu_aspiration = (uint32_t)(p->step_count_limit & 1);
  if (sa) {  // simulated_annealing.rs:11-15 defaults; late_size = calibration sample size, step_count_limit bit 0 =
             // HardRegressionPolicy::NeverAcceptHardRegression
    s.sa.decay = p->acceptor_real > 0.0 ? p->acceptor_real : 0.999985;
    s.sa.hc_temp = 1.0e-9;
    s.sa.fallback = 1.0;
    s.sa.neg_log_target = -std::log(0.80);
    s.sa.sample_size = p->late_size ? p->late_size : 128;
    s.sa.never_hard = (int32_t)(p->step_count_limit & 1);
  }
  uint32_t* d_counts = (uint32_t*)(b + o_cnts);
  const int forage_code = solve_forage_code(p->acceptor);
  uint64_t* d_overflows = (uint64_t*)(b + o_ovf);
  CU(cudaMemsetAsync(d_overflows, 0, (size_t)R * 16, ctx->stream));  // [R] overflows, [R] pulls scored
  UnionPlan& plan = ctx->union_plan;
  if (udesc) {
    sfgpu_forage_params fp{forage_code, p->tie_mode, p->accepted_limit, 0};
    rc = sfgpu_union_prepare(ctx, udesc, &fp, plan);
    if (rc) return rc;
    plan.a.step_seeds = s.step_seeds;
    plan.a.step_indices = s.step_counter;
    plan.a.step_index_shared = 1;
    plan.a.ref_scores = s.ref_scores;
    plan.sa = sa;
    plan.sa_cur = s.sa_cur;
    plan.sa_nxt = s.sa_nxt;
    plan.sa_params = s.sa;
    rc = sfgpu_union_reset_windows(ctx, plan);
    if (rc) return rc;
  }
  NearbyArgs a{};
  ChangeStepArgs ca{};
  uint32_t* d_nper = nullptr;  // list TabuSearch: candidates per source of every replica
  uint32_t c_chunks = 0;
  if (scalar) {
    uint32_t per = 0;
    sfgpu_change_step_chunks(ctx, &per, &c_chunks);
    rc = ensure_partials(ctx, (size_t)R * c_chunks * sizeof(ChunkPartial));
    if (rc) return rc;
    ca.f = ForageDev{forage_code, p->tie_mode, p->accepted_limit};
    ca.ents_per_cta = per;
    ca.step_seeds = s.step_seeds;
    ca.ref_scores = s.ref_scores;
    ca.partials = (ChunkPartial*)ctx->partials;
    if (sa || tabu) {  // the acceptor replays over the materialised, pull-ordered scores
      const size_t stride = (size_t)dm.n_entities * (dm.n_values + 1);
      const size_t need = (size_t)R * stride * (8 + 16 + 1) + (size_t)(R + 1) * 8 + 64;
      rc = ensure_staging(ctx, 64, need);
      if (rc) return rc;
      char* q = (char*)ctx->dscr;
      ca.out_scores = (int64_t*)q;
      ca.out_rows = (uint32_t*)(q + (size_t)R * stride * 16);
      ca.out_offsets = (uint64_t*)(q + (size_t)R * stride * 24);
      ca.out_doable = (uint8_t*)(q + (size_t)R * stride * 24 + (size_t)(R + 1) * 8);
      ca.out_counts = d_counts;
    }
  } else if (!udesc) {
    rc = ensure_partials(ctx, (size_t)R * dm.elem_cap * sizeof(SrcPartial));
    if (rc) return rc;
    rc = sfgpu_nearby_prepare_cache(ctx, p->max_nearby);  // retained neighbourhood: allocated before the capture
    if (rc) return rc;
    a.f = ForageDev{forage_code, p->tie_mode, p->accepted_limit};
    a.max_nearby = p->max_nearby;
    a.step_seeds = s.step_seeds;
    a.ref_scores = s.ref_scores;
    a.partials = (SrcPartial*)ctx->partials;
    if (tabu || sa) {  // TabuSearch / SimulatedAnnealing decide over the materialised batch (rows, scores, doable), like the scalar loop
      const size_t stride = (size_t)dm.elem_cap * p->max_nearby;
      const size_t need = (size_t)R * stride * (16 + 16 + 1) + (size_t)(R + 1) * 8 + (size_t)R * 4 + 64;
      rc = ensure_staging(ctx, 64, need);
      if (rc) return rc;
      char* q = (char*)ctx->dscr;
      a.out_scores = (int64_t*)q;
      a.out_rows = (uint32_t*)(q + (size_t)R * stride * 16);
      a.out_offsets = (uint64_t*)(q + (size_t)R * stride * 32);
      d_nper = (uint32_t*)(q + (size_t)R * stride * 32 + (size_t)(R + 1) * 8);
      a.out_doable = (uint8_t*)(q + (size_t)R * stride * 32 + (size_t)(R + 1) * 8 + (size_t)R * 4);
    }
  }
  solve_init_kernel<<<R, 256, 0, ctx->stream>>>(dm, s);
  ctx->launches++;
  CU(cudaGetLastError());
  auto one_step = [&]() -> int {
    solve_prep_kernel<<<(R + 127) / 128, 128, 0, ctx->stream>>>(dm, s);
    int rc2;
    if (scalar) {
      rc2 = sfgpu_launch_change_step(ctx, ca, c_chunks, s.out_index, s.out_best, s.out_evaluated, s.winner_rows);
      if (rc2) return rc2;
      if (sa || tabu) {  // the fused partials of the step kernel are ignored: the acceptor decides over the batch
        if (sa) {
          rc2 = sfgpu_launch_sa_accept(ctx, ca.out_offsets, d_counts, nullptr, ca.out_scores, ca.out_doable, s.ref_scores,
                                       s.step_seeds, s.sa_cur, s.sa_nxt, s.sa, p->accepted_limit);
          if (rc2) return rc2;
        } else {
          tabu_accept_kernel<<<R, 256, 0, ctx->stream>>>(dm, s, ca.out_offsets, d_counts, ca.out_rows, ca.out_scores, ca.out_doable);
        }
        rc2 = sfgpu_launch_argbest_counts(ctx, ForageDev{0, p->tie_mode, p->accepted_limit, nullptr}, ca.out_offsets, d_counts,
                                          nullptr, ca.out_scores, ca.out_doable, s.step_seeds, s.ref_scores, s.out_index,
                                          s.out_best, s.out_evaluated);
        if (rc2) return rc2;
        if (tabu) tabu_record_kernel<<<(R + 127) / 128, 128, 0, ctx->stream>>>(dm, s, ca.out_offsets, ca.out_rows);
        rc2 = sfgpu_launch_apply_scalar(ctx, 0, ca.out_rows, nullptr, ca.out_offsets, s.out_index);
      } else {
        rc2 = sfgpu_launch_apply_scalar(ctx, 0, s.winner_rows, nullptr, nullptr, nullptr);
      }
    } else if (udesc) {
      // three window passes: the replica's adaptive window (what its previous step needed plus half), four times
      // that, then max_window — each only for the replicas whose forager had not quit in the pass before
      rc2 = sfgpu_union_begin_step(ctx, plan);
      if (rc2) return rc2;
      for (int k = 0; k < 3; ++k) {
        bool used = false;
        if (k > 0) {  // inside a captured step graph the later passes are IF nodes: skipped when nothing is pending
          rc2 = sfgpu_union_conditional_pass(ctx, plan, (uint32_t)k, plan.wmax, k == 2, s.out_index, s.out_best, s.out_evaluated,
                                             d_overflows, k == 1 ? 2 : 0, &used);
          if (rc2) return rc2;
        }
        if (!used)
          rc2 = sfgpu_union_launch_pass(ctx, plan, plan.wmax, k == 2, s.out_index, s.out_best, s.out_evaluated, nullptr, nullptr,
                                        d_overflows, true, k == 1 ? 2 : 0, (uint32_t)k);
        if (rc2) return rc2;
      }
      rc2 = sfgpu_launch_apply_list_kinds(ctx, plan.apply_rows, plan.apply_kinds);
    } else {
      rc2 = sfgpu_launch_nearby(ctx, a, s.out_index, s.out_best, s.out_evaluated, s.winner_rows, MOVE_CHANGE);
      if (rc2) return rc2;
      if (tabu || sa) {  // the fused forager's pick is replaced: acceptor over the batch, ordered replay, record, commit by row index
        tabu_accept_list_kernel<<<R, 256, 0, ctx->stream>>>(dm, s, p->max_nearby, d_counts, d_nper, a.out_rows, a.out_scores,
                                                            a.out_doable);
        if (sa) {
          rc2 = sfgpu_launch_sa_accept(ctx, a.out_offsets, d_counts, nullptr, a.out_scores, a.out_doable, s.ref_scores,
                                       s.step_seeds, s.sa_cur, s.sa_nxt, s.sa, p->accepted_limit);
          if (rc2) return rc2;
        }
        rc2 = sfgpu_launch_argbest_counts(ctx, ForageDev{0, p->tie_mode, p->accepted_limit, nullptr}, a.out_offsets, d_counts,
                                          nullptr, a.out_scores, a.out_doable, s.step_seeds, s.ref_scores, s.out_index,
                                          s.out_best, s.out_evaluated);
        if (rc2) return rc2;
        tabu_record_list_kernel<<<(R + 127) / 128, 128, 0, ctx->stream>>>(dm, s, p->max_nearby, d_nper, a.out_rows);
        rc2 = sfgpu_launch_apply_list(ctx, 2, a.out_rows, nullptr, a.out_offsets, s.out_index);
      } else
      rc2 = sfgpu_launch_apply_list(ctx, 2, s.winner_rows, nullptr, nullptr, nullptr);
    }
    if (rc2) return rc2;
    solve_post_kernel<<<R, 256, 0, ctx->stream>>>(dm, s);
    return SFGPU_OK;
  };
  // steps are captured once into a CUDA graph of `per_graph` steps and replayed
  const uint32_t per_graph = std::min<uint32_t>(p->n_steps, 16);
  uint32_t done = 0;
  if (per_graph >= 2) {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    int rc2 = SFGPU_OK;
    for (uint32_t i = 0; i < per_graph && rc2 == SFGPU_OK; ++i) rc2 = one_step();
    cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc2 != SFGPU_OK || ce != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      return fail(ctx, SFGPU_E_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(ce));
    }
    CU(cudaGraphInstantiate(&exec, graph, 0));
    for (; done + per_graph <= p->n_steps; done += per_graph) {
      CU(cudaGraphLaunch(exec, ctx->stream));
      ctx->launches += (5ull + (a.cache ? 1 : 0)) * per_graph;
    }
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
  }
  for (; done < p->n_steps; ++done) {
    rc = one_step();
    if (rc) return rc;
    ctx->launches += 5 + (a.cache ? 1 : 0);
  }
  if (p->restore_best) {
    solve_restore_best_kernel<<<R, 256, 0, ctx->stream>>>(dm, s);
    ctx->launches++;
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  if (out_best_scores) CU(cudaMemcpy(out_best_scores, s.best_scores, (size_t)R * 16, cudaMemcpyDeviceToHost));
  if (out_moves_evaluated) CU(cudaMemcpy(out_moves_evaluated, s.evaluated, (size_t)R * 8, cudaMemcpyDeviceToHost));
  if (out_accepted_steps) CU(cudaMemcpy(out_accepted_steps, s.accepted_steps, (size_t)R * 8, cudaMemcpyDeviceToHost));
  if (out_window_overflows) CU(cudaMemcpy(out_window_overflows, d_overflows, (size_t)R * 8, cudaMemcpyDeviceToHost));
  if (out_pulls_scored) CU(cudaMemcpy(out_pulls_scored, d_overflows + R, (size_t)R * 8, cudaMemcpyDeviceToHost));
  return SFGPU_OK;
}

}  // namespace

extern "C" {

int32_t sfgpu_solve_nearby_list_change(sfgpu_ctx* ctx, const sfgpu_solve_params* p, int64_t* out_best_scores,
                                       uint64_t* out_moves_evaluated, uint64_t* out_accepted_steps) try {
  return solve_impl(ctx, p, false, out_best_scores, out_moves_evaluated, out_accepted_steps);
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_solve_change(sfgpu_ctx* ctx, const sfgpu_solve_params* p, int64_t* out_best_scores,
                           uint64_t* out_moves_evaluated, uint64_t* out_accepted_steps) try {
  return solve_impl(ctx, p, true, out_best_scores, out_moves_evaluated, out_accepted_steps);
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_solve_union(sfgpu_ctx* ctx, const sfgpu_union_desc* desc, const sfgpu_solve_params* p, int64_t* out_best_scores,
                          uint64_t* out_moves_evaluated, uint64_t* out_accepted_steps, uint64_t* out_window_overflows,
                          uint64_t* out_pulls_scored) try {
  if (!desc) return fail(ctx, SFGPU_E_INVALID, "null union description");
  return solve_impl(ctx, p, false, out_best_scores, out_moves_evaluated, out_accepted_steps, desc, out_window_overflows,
                    out_pulls_scored);
} SFGPU_API_CATCH(ctx)

}  // extern "C"
