// Union of list neighbourhoods in seeded pull order (sfgpu_union.cuh): host side of sfgpu_step_union and the
// per-step launcher the device-resident loop (sfgpu_solve.cu) captures.
#include "sfgpu_ctx.hpp"
#include "sfgpu_union.cuh"

using namespace sfgpu_host;

namespace {
bool is_nearby(int family) { return family == SFGPU_FAM_NEARBY_LIST_CHANGE || family == SFGPU_FAM_NEARBY_LIST_SWAP; }
bool is_scalar_family(int family) { return family == SFGPU_FAM_CHANGE || family == SFGPU_FAM_SWAP; }
}  // namespace

int sfgpu_union_prepare(sfgpu_ctx* ctx, const sfgpu_union_desc* desc, const sfgpu_forage_params* params, UnionPlan& plan) {
  const DevModel& dm = ctx->dm;
  if (!desc || !params) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  if (desc->n_children == 0 || desc->n_children > SFGPU_UNION_MAX_CHILDREN)
    return fail(ctx, SFGPU_E_INVALID, "union needs 1..8 children");
  if (desc->union_order < 0 || desc->union_order > SFGPU_UNION_STRATIFIED_RANDOM)
    return fail(ctx, SFGPU_E_INVALID, "bad union_order");
  if (desc->selection_order < 0 || desc->selection_order > SFGPU_ORDER_SHUFFLED)
    return fail(ctx, SFGPU_E_INVALID, "bad selection_order");
  if (params->acceptor < 0 || params->acceptor > 3 || params->tie_mode < 0 || params->tie_mode > 1)
    return fail(ctx, SFGPU_E_INVALID, "bad forage params");
  bool any_nearby = false;
  for (uint32_t c = 0; c < desc->n_children; ++c) {
    const sfgpu_union_child& ch = desc->children[c];
    if (ch.family < 0 || ch.family > SFGPU_FAM_SWAP) return fail(ctx, SFGPU_E_INVALID, "unknown move family");
    if (is_scalar_family(ch.family) ? !dm.has_scalar : !dm.has_list)
      return fail(ctx, SFGPU_E_STATE, "the model has no variable of the kind this move family edits");
    if (ch.family == SFGPU_FAM_SWAP && ctx->has_load_balance)
      return fail(ctx, SFGPU_E_UNSUPPORTED, "swap moves over a load_balance constraint");
    if (is_nearby(ch.family)) {
      any_nearby = true;
      if (ch.p0 == 0 || ch.p0 > 32) return fail(ctx, SFGPU_E_UNSUPPORTED, "max_nearby must be in [1, 32]");
    } else if (is_scalar_family(ch.family)) {
      // no parameters
    } else if (ch.family == SFGPU_FAM_K_OPT) {
      if (ch.p0 < 2 || ch.p0 > 5 || ch.p1 == 0) return fail(ctx, SFGPU_E_INVALID, "k-opt: 2 <= k <= 5, min_segment_len >= 1");
      if (dm.elem_cap >= 65536) return fail(ctx, SFGPU_E_UNSUPPORTED, "k-opt rows hold 16-bit cut positions");
    } else if (ch.family != SFGPU_FAM_LIST_REVERSE) {
      if (ch.p0 == 0 || ch.p1 < ch.p0 || ch.p1 > 255) return fail(ctx, SFGPU_E_INVALID, "sublist sizes: 1 <= min <= max <= 255");
    }
    if (desc->union_order != SFGPU_UNION_RANDOM && desc->union_order != SFGPU_UNION_STRATIFIED_RANDOM && ch.weight != 1)
      return fail(ctx, SFGPU_E_INVALID, "union weights require random or stratified_random selection order");
    if (ch.weight > (1ull << 40)) return fail(ctx, SFGPU_E_INVALID, "union weight too large");
  }
  if (any_nearby && (!dm.nearby_ok || ctx->force_generic || dm.fast_pc < 0))
    return fail(ctx, SFGPU_E_UNSUPPORTED,
                "nearby families need the fast list program (int32 path-cost matrix as the distance meter, every cell finite)");
  if (any_nearby && rank_tables_words_host(dm.n_owners) * 4 + 8192 > (size_t)ctx->max_smem_optin)
    return fail(ctx, SFGPU_E_UNSUPPORTED, "too many list owners for the shared-memory rank tables");
  const uint32_t w0 = desc->window ? desc->window : 64;
  const uint32_t wmax = std::max(w0, desc->max_window ? desc->max_window : 4096u);
  if (wmax >= (1u << 28)) return fail(ctx, SFGPU_E_INVALID, "max_window must be below 2^28");
  CU(cudaSetDevice(ctx->device));
  const uint32_t R = dm.R, C = desc->n_children;
  const size_t T = (size_t)C * wmax;
  auto a256 = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = a256(o + bytes); return at; };
  const size_t o_rows = take((size_t)R * T * 16), o_sched = take((size_t)R * T * 4), o_scores = take((size_t)R * T * 16);
  const size_t o_cscores = take((size_t)R * T * 16), o_cdoable = take((size_t)R * T);
  const size_t o_doable = take((size_t)R * T), o_emit = take((size_t)R * C * 4), o_ended = take((size_t)R * C * 4);
  const size_t o_nsched = take((size_t)R * 4), o_send = take((size_t)R * 4), o_offs = take((size_t)(R + 1) * 8);
  const size_t o_done = take((size_t)R * 4), o_pend = take(16), o_arows = take((size_t)R * 16), o_akind = take((size_t)R * 4);
  const size_t o_win = take((size_t)R * 4);
  if (o > ctx->union_bytes) {
    if (ctx->union_buf) cudaFree(ctx->union_buf);
    ctx->union_buf = nullptr;
    ctx->union_bytes = 0;
    CU(cudaMalloc(&ctx->union_buf, o));
    ctx->union_bytes = o;
  }
  char* b = (char*)ctx->union_buf;
  UnionArgs& a = plan.a;
  a = UnionArgs{};
  a.f = ForageDev{params->acceptor, params->tie_mode, params->accepted_limit, nullptr};
  a.n_children = C;
  a.union_order = desc->union_order;
  a.order = desc->selection_order;
  a.desc = ctx->lvars.empty() ? 0u : (uint32_t)std::max(0, ctx->colls[ctx->lvars[0].owner_coll].descriptor);
  a.sdesc = ctx->svars.empty() ? 0u : (uint32_t)std::max(0, ctx->colls[ctx->svars[0].coll].descriptor);
  a.scan_bits = 32;
  for (uint32_t c = 0; c < C; ++c) a.child[c] = UnionChildDev{desc->children[c].family, desc->children[c].p0, desc->children[c].p1, 0, desc->children[c].weight};
  a.rows = (uint32_t*)(b + o_rows);
  a.sched = (uint32_t*)(b + o_sched);
  a.scores = (int64_t*)(b + o_scores);
  a.doable = (uint8_t*)(b + o_doable);
  a.child_scores = (int64_t*)(b + o_cscores);
  a.child_doable = (uint8_t*)(b + o_cdoable);
  a.n_emit = (uint32_t*)(b + o_emit);
  a.ended = (uint32_t*)(b + o_ended);
  a.n_sched = (uint32_t*)(b + o_nsched);
  a.stream_end = (uint32_t*)(b + o_send);
  a.offsets = (uint64_t*)(b + o_offs);
  a.done = (uint32_t*)(b + o_done);
  plan.pending = (uint32_t*)(b + o_pend);
  plan.apply_rows = (uint32_t*)(b + o_arows);
  plan.apply_kinds = (int32_t*)(b + o_akind);
  plan.win_state = (uint32_t*)(b + o_win);
  a.win_max = wmax;
  plan.w0 = w0;
  plan.wmax = wmax;
  while (ctx->aux_streams.size() < 3 * (size_t)(C - 1) + 2) {  // three disjoint sets (one per window pass) + 2 body streams
    cudaStream_t aux = nullptr;
    CU(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
    ctx->aux_streams.push_back(aux);
  }
  while (ctx->aux_events.size() < 3 * (size_t)C) {
    cudaEvent_t ev = nullptr;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ctx->aux_events.push_back(ev);
  }
  if (!plan.configured) {
    const int tb = (int)(rank_tables_words_host(dm.n_owners) * 4);
#define UW_ATTR(CELL, MOVE) \
  CU(cudaFuncSetAttribute(union_walk_nearby_kernel<uint64_t, CELL, MOVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, tb))
    UW_ATTR(uint16_t, MOVE_CHANGE);
    UW_ATTR(uint16_t, MOVE_SWAP);
    UW_ATTR(int32_t, MOVE_CHANGE);
    UW_ATTR(int32_t, MOVE_SWAP);
    plan.configured = true;
  }
  return SFGPU_OK;
}

// resident loop: every replica starts with the first window
int sfgpu_union_reset_windows(sfgpu_ctx* ctx, UnionPlan& plan) {
  const uint32_t R = ctx->dm.R;
  union_fill_kernel<<<(R + 255) / 256, 256, 0, ctx->stream>>>(plan.win_state, plan.w0, R);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}

int sfgpu_union_begin_step(sfgpu_ctx* ctx, UnionPlan& plan) {
  const uint32_t R = ctx->dm.R;
  union_reset_kernel<<<(R + 255) / 256, 256, 0, ctx->stream>>>(plan.a.done, plan.pending, R);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}

// one window pass: walk every child, schedule, score, replay, pick
int sfgpu_union_launch_pass(sfgpu_ctx* ctx, UnionPlan& plan, uint32_t window, bool last_pass, uint32_t* d_idx, int64_t* d_best,
                            uint32_t* d_eval, uint32_t* d_win8, uint32_t* d_flags, uint64_t* d_overflow_acc, bool adaptive,
                            uint32_t win_shift, uint32_t stream_set) {
  const DevModel& dm = ctx->dm;
  const uint32_t R = dm.R;
  UnionArgs a = plan.a;
  a.window = window;
  a.win_r = adaptive && !last_pass ? plan.win_state : nullptr;  // the last pass always offers the whole window
  a.next_win = adaptive ? plan.win_state : nullptr;
  a.win_shift = win_shift;
  a.t_cap = a.n_children * window;
  // every child walks and scores on its own stream (child 0 on the context's): the cursors are independent and each
  // is a latency-bound kernel of one small CTA per replica
  const uint32_t chunks = std::max<uint32_t>(1, std::min<uint32_t>((window + 127) / 128,
                                                                   std::max<uint32_t>(2, (uint32_t)ctx->sm_count * 8 / std::max(R, 1u))));
  const bool fork = a.n_children > 1;
  // side streams and events of this pass (disjoint per pass: a conditional pass is captured into its own body graph)
  cudaStream_t* side = ctx->aux_streams.data() + (size_t)stream_set * (a.n_children - 1);
  cudaEvent_t* evs = ctx->aux_events.data() + (size_t)stream_set * a.n_children;
  if (fork) CU(cudaEventRecord(evs[0], ctx->stream));
  for (uint32_t c = 0; c < a.n_children; ++c) {
    cudaStream_t cs = c == 0 ? ctx->stream : side[c - 1];
    if (c > 0) CU(cudaStreamWaitEvent(cs, evs[0], 0));
    const int fam = a.child[c].family;
    if (is_nearby(fam)) {
      const size_t tb = rank_tables_words_host(dm.n_owners) * 4;
      if (fam == SFGPU_FAM_NEARBY_LIST_CHANGE) {
        if (dm.fm_u16) union_walk_nearby_kernel<uint64_t, uint16_t, MOVE_CHANGE><<<R, 256, tb, cs>>>(dm, a, c);
        else union_walk_nearby_kernel<uint64_t, int32_t, MOVE_CHANGE><<<R, 256, tb, cs>>>(dm, a, c);
      } else {
        if (dm.fm_u16) union_walk_nearby_kernel<uint64_t, uint16_t, MOVE_SWAP><<<R, 256, tb, cs>>>(dm, a, c);
        else union_walk_nearby_kernel<uint64_t, int32_t, MOVE_SWAP><<<R, 256, tb, cs>>>(dm, a, c);
      }
    } else {
      union_walk_index_kernel<<<R, 32, 0, cs>>>(dm, a, c);
    }
    const dim3 g(chunks, R);
    switch (fam) {
      case SFGPU_FAM_NEARBY_LIST_CHANGE: union_score_child_kernel<SFGPU_FAM_NEARBY_LIST_CHANGE><<<g, 128, 0, cs>>>(dm, a, c); break;
      case SFGPU_FAM_NEARBY_LIST_SWAP: union_score_child_kernel<SFGPU_FAM_NEARBY_LIST_SWAP><<<g, 128, 0, cs>>>(dm, a, c); break;
      case SFGPU_FAM_SUBLIST_CHANGE: union_score_child_kernel<SFGPU_FAM_SUBLIST_CHANGE><<<g, 128, 0, cs>>>(dm, a, c); break;
      case SFGPU_FAM_SUBLIST_SWAP: union_score_child_kernel<SFGPU_FAM_SUBLIST_SWAP><<<g, 128, 0, cs>>>(dm, a, c); break;
      case SFGPU_FAM_LIST_REVERSE: union_score_child_kernel<SFGPU_FAM_LIST_REVERSE><<<g, 128, 0, cs>>>(dm, a, c); break;
      case SFGPU_FAM_CHANGE: union_score_child_kernel<SFGPU_FAM_CHANGE><<<g, 128, 0, cs>>>(dm, a, c); break;
      case SFGPU_FAM_SWAP: union_score_child_kernel<SFGPU_FAM_SWAP><<<g, 128, 0, cs>>>(dm, a, c); break;
      default: union_score_child_kernel<SFGPU_FAM_K_OPT><<<g, 128, 0, cs>>>(dm, a, c); break;
    }
    if (c > 0) {
      CU(cudaEventRecord(evs[c], cs));
      CU(cudaStreamWaitEvent(ctx->stream, evs[c], 0));
    }
  }
  switch (a.union_order) {
    case SFGPU_UNION_SEQUENTIAL: union_schedule_kernel<SFGPU_UNION_SEQUENTIAL><<<(R + 63) / 64, 64, 0, ctx->stream>>>(a, R); break;
    case SFGPU_UNION_ROUND_ROBIN: union_schedule_kernel<SFGPU_UNION_ROUND_ROBIN><<<(R + 63) / 64, 64, 0, ctx->stream>>>(a, R); break;
    case SFGPU_UNION_ROTATING_ROUND_ROBIN:
      union_schedule_kernel<SFGPU_UNION_ROTATING_ROUND_ROBIN><<<(R + 63) / 64, 64, 0, ctx->stream>>>(a, R);
      break;
    case SFGPU_UNION_RANDOM: union_schedule_kernel<SFGPU_UNION_RANDOM><<<(R + 63) / 64, 64, 0, ctx->stream>>>(a, R); break;
    default: union_schedule_kernel<SFGPU_UNION_STRATIFIED_RANDOM><<<(R + 63) / 64, 64, 0, ctx->stream>>>(a, R); break;
  }
  {
    const uint32_t gchunks = std::max<uint32_t>(1, std::min<uint32_t>((a.t_cap + 255) / 256, std::max<uint32_t>(2, (uint32_t)ctx->sm_count * 4 / std::max(R, 1u))));
    union_gather_kernel<<<dim3(gchunks, R), 256, 0, ctx->stream>>>(a);
  }
  ctx->launches += 2 * a.n_children + 2;
  CU(cudaGetLastError());
  if (plan.sa) {
    int rc0 = sfgpu_launch_sa_accept(ctx, a.offsets, a.n_sched, a.done, a.scores, a.doable, a.ref_scores, a.step_seeds, plan.sa_cur,
                                     plan.sa_nxt, plan.sa_params, a.f.accepted_limit);
    if (rc0) return rc0;
  }
  ctx->argbest_stride_hint = (adaptive && !last_pass) ? (uint64_t)a.n_children * plan.w0 * 4 : a.t_cap;
  int rc = sfgpu_launch_argbest_counts(ctx, a.f, a.offsets, a.n_sched, a.done, a.scores, a.doable, a.step_seeds, a.ref_scores,
                                       d_idx, d_best, d_eval);
  if (rc) return rc;
  union_pick_kernel<<<(R + 127) / 128, 128, 0, ctx->stream>>>(a, R, last_pass ? 1u : 0u, d_idx, d_eval, d_win8, plan.apply_rows,
                                                              plan.apply_kinds, d_flags, plan.pending, d_overflow_acc);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}

// A later window pass inside a captured step graph: an IF node whose body holds the pass, taken only when the pass
// before left replicas pending (cudaGraphConditionalHandle set on the device by union_cond_kernel). The body is
// captured from a dedicated stream into the node's body graph. Returns SFGPU_OK with *used = false when the stream is
// not capturing (the caller then launches the pass unconditionally).
int sfgpu_union_conditional_pass(sfgpu_ctx* ctx, UnionPlan& plan, uint32_t pass_index, uint32_t window, bool last_pass, uint32_t* d_idx,
                                 int64_t* d_best, uint32_t* d_eval, uint64_t* d_overflow_acc, uint32_t win_shift, bool* used) {
  *used = false;
  static const bool disabled = getenv("SFGPU_NO_COND") != nullptr;  // tuning / debugging knob
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (disabled || cudaStreamIsCapturing(ctx->stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusActive) return SFGPU_OK;
  cudaGraph_t graph = nullptr;
  const cudaGraphNode_t* deps = nullptr;
  size_t n_deps = 0;
  CU(cudaStreamGetCaptureInfo(ctx->stream, &st, nullptr, &graph, &deps, &n_deps));
  cudaGraphConditionalHandle handle;
  CU(cudaGraphConditionalHandleCreate(&handle, graph, 0, cudaGraphCondAssignDefault));
  union_cond_kernel<<<1, 1, 0, ctx->stream>>>(handle, plan.pending);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamGetCaptureInfo(ctx->stream, &st, nullptr, &graph, &deps, &n_deps));
  cudaGraphNodeParams np{};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = handle;
  np.conditional.type = cudaGraphCondTypeIf;
  np.conditional.size = 1;
  cudaGraphNode_t node = nullptr;
  CU(cudaGraphAddNode(&node, graph, deps, n_deps, &np));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  CU(cudaStreamUpdateCaptureDependencies(ctx->stream, &node, 1, cudaStreamSetCaptureDependencies));
  // the body: zero the pending counter, then the pass — captured from a body stream into the node's graph
  cudaStream_t outer = ctx->stream;
  cudaStream_t bs = ctx->aux_streams[ctx->aux_streams.size() - 1 - (pass_index & 1)];
  CU(cudaStreamBeginCaptureToGraph(bs, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  ctx->stream = bs;
  union_fill_kernel<<<1, 1, 0, bs>>>(plan.pending, 0u, 1u);
  int rc = sfgpu_union_launch_pass(ctx, plan, window, last_pass, d_idx, d_best, d_eval, nullptr, nullptr, d_overflow_acc, true, win_shift,
                                   pass_index);
  ctx->stream = outer;
  cudaError_t ce = cudaStreamEndCapture(bs, nullptr);
  if (rc) return rc;
  if (ce != cudaSuccess) return fail(ctx, SFGPU_E_CUDA, std::string("conditional pass capture: ") + cudaGetErrorString(ce));
  *used = true;
  return SFGPU_OK;
}

extern "C" {

int32_t sfgpu_step_union(sfgpu_ctx* ctx, uint32_t flags, const sfgpu_union_desc* desc, const sfgpu_forage_params* params,
                         const uint64_t* step_seeds, const uint64_t* step_indices, const int64_t* ref_scores,
                         uint32_t* out_index, int64_t* out_best, uint32_t* out_evaluated, uint32_t* out_winner_rows,
                         uint32_t* out_flags, int32_t apply_winners) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (!out_index || !out_best) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  if (params && params->acceptor != 0 && !ref_scores) return fail(ctx, SFGPU_E_INVALID, "acceptor needs ref_scores");
  UnionPlan& plan = ctx->union_plan;
  plan.sa = false;
  rc = sfgpu_union_prepare(ctx, desc, params, plan);
  if (rc) return rc;
  const uint32_t R = ctx->dm.R;
  const bool dev_io = (flags & SFGPU_DEVICE_IO) != 0;
  SmallIo io;
  rc = small_io_begin(ctx, io, dev_io, 32, step_seeds, ref_scores, out_index, out_best, out_evaluated, out_winner_rows);
  if (rc) return rc;
  // step indices and flags ride in a second small staging area
  uint64_t* d_steps = nullptr;
  uint32_t* d_flags = nullptr;
  if (dev_io) {
    d_steps = const_cast<uint64_t*>(step_indices);
    d_flags = out_flags;
    if (!io.d_eval) return fail(ctx, SFGPU_E_INVALID, "the device path needs out_evaluated");
  } else {
    rc = ensure_staging(ctx, (size_t)R * 16, (size_t)R * 16);
    if (rc) return rc;
    if (step_indices) {
      memcpy(ctx->pin, step_indices, (size_t)R * 8);
      CU(cudaMemcpyAsync(ctx->dscr, ctx->pin, (size_t)R * 8, cudaMemcpyHostToDevice, ctx->stream));
      d_steps = (uint64_t*)ctx->dscr;
    }
    d_flags = (uint32_t*)((char*)ctx->dscr + (size_t)R * 8);
  }
  plan.a.step_seeds = io.d_seeds;
  plan.a.step_indices = d_steps;
  plan.a.step_index_shared = 0;
  plan.a.ref_scores = io.d_ref;
  rc = sfgpu_union_begin_step(ctx, plan);
  if (rc) return rc;
  ev_begin(ctx);
  for (uint32_t window = plan.w0;;) {
    const bool last = window >= plan.wmax;
    rc = sfgpu_union_launch_pass(ctx, plan, window, last, io.d_idx, io.d_best, io.d_eval, io.d_win, d_flags, nullptr, false, 0, 0);
    if (rc) return rc;
    if (last) break;
    uint32_t pending = 0;
    CU(cudaMemcpyAsync(&pending, plan.pending, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (pending == 0) break;
    CU(cudaMemsetAsync(plan.pending, 0, 4, ctx->stream));
    window = (uint32_t)std::min<uint64_t>((uint64_t)window * 8, plan.wmax);
  }
  ev_end(ctx);
  if (apply_winners) {
    rc = sfgpu_launch_apply_list_kinds(ctx, plan.apply_rows, plan.apply_kinds);
    if (rc) return rc;
  }
  if (!dev_io && out_flags) {
    CU(cudaMemcpyAsync((char*)ctx->pin + (size_t)R * 8, d_flags, (size_t)R * 4, cudaMemcpyDeviceToHost, ctx->stream));
    rc = small_io_end(ctx, io, out_index, out_best, out_evaluated, out_winner_rows);
    if (rc) return rc;
    memcpy(out_flags, (char*)ctx->pin + (size_t)R * 8, (size_t)R * 4);
    return SFGPU_OK;
  }
  return small_io_end(ctx, io, out_index, out_best, out_evaluated, out_winner_rows);
} SFGPU_API_CATCH(ctx)

}  // extern "C"

// test hook (not part of the ABI): the rows child `child` of replica `replica` emitted in the last pass
extern "C" int32_t sfgpu_debug_union_rows(sfgpu_ctx* ctx, uint32_t replica, uint32_t child, uint32_t window, uint32_t* out_rows,
                                          uint32_t cap, uint32_t* out_n, uint32_t* out_ended) {
  const UnionPlan& plan = ctx->union_plan;
  const UnionArgs& a = plan.a;
  uint32_t n = 0, ended = 0;
  cudaStreamSynchronize(ctx->stream);
  cudaMemcpy(&n, a.n_emit + (size_t)replica * a.n_children + child, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&ended, a.ended + (size_t)replica * a.n_children + child, 4, cudaMemcpyDeviceToHost);
  *out_n = n;
  *out_ended = ended;
  const uint32_t k = n < cap ? n : cap;
  cudaMemcpy(out_rows, a.rows + ((size_t)replica * a.n_children + child) * window * 4, (size_t)k * 16, cudaMemcpyDeviceToHost);
  return SFGPU_OK;
}
