// Scalar-move kernels of libsfgpu: ChangeMove / SwapMove / CompoundScalarMove batches (interpreter and
// monomorphised programs) and the whole ChangeMove step generated on device.
#include "sfgpu_ctx.hpp"
#include "sfgpu_change_step.cuh"

using namespace sfgpu_host;

namespace {

// Monomorphised scalar programs: (sorted) kinds of the scalar constraints -> kernel instantiations.
// The tuples of the reference's scalar examples (graph colouring, n-queens, job shop) and their prefixes.
typedef void (*SpecScoreFn)(const DevModel, const SpecIdx, const uint64_t*, const uint32_t*, int64_t*, uint8_t*);
typedef void (*SpecStepFn)(const DevModel, const ChangeStepArgs, const SpecIdx);
struct SpecEntry {
  int k[4];
  SpecScoreFn score;
  SpecStepFn step;
};
#define SPEC_ENTRY(a, b, c, d) \
  { {a, b, c, d}, spec_change_kernel<SpecProg<a, b, c, d>>, change_step_kernel<true, SpecProg<a, b, c, d>> }
#define U_ SFGPU_K_UNI
#define C_ SFGPU_K_PAIR_CSR_EQUAL
#define K_ SFGPU_K_PAIR_KEY_EQUAL
#define G_ SFGPU_K_GROUP
#define UC SPEC_K_UNI_CONST
const SpecEntry g_spec[] = {  // kinds ascending; UC (uni without column / mask) sorts last
    SPEC_ENTRY(U_, 0, 0, 0),   SPEC_ENTRY(UC, 0, 0, 0),    // unassigned only
    SPEC_ENTRY(U_, C_, 0, 0),  SPEC_ENTRY(C_, UC, 0, 0),   // graph colouring
    SPEC_ENTRY(U_, K_, 0, 0),  SPEC_ENTRY(K_, UC, 0, 0),
    SPEC_ENTRY(U_, G_, 0, 0),  SPEC_ENTRY(G_, UC, 0, 0),
    SPEC_ENTRY(U_, C_, G_, 0), SPEC_ENTRY(C_, G_, UC, 0),
    SPEC_ENTRY(U_, K_, G_, 0), SPEC_ENTRY(K_, G_, UC, 0),  // job shop
    SPEC_ENTRY(U_, K_, K_, 0), SPEC_ENTRY(K_, K_, UC, 0),
    SPEC_ENTRY(U_, K_, K_, K_), SPEC_ENTRY(K_, K_, K_, UC),  // n-queens
    SPEC_ENTRY(U_, K_, G_, G_), SPEC_ENTRY(K_, G_, G_, UC),
    SPEC_ENTRY(U_, U_, K_, G_), SPEC_ENTRY(U_, K_, G_, UC),
};
#undef U_
#undef C_
#undef K_
#undef G_
#undef UC
#undef SPEC_ENTRY

}  // namespace

int sfgpu_configure_scalar(sfgpu_ctx* ctx) {
  DevModel& dm = ctx->dm;
  if (ctx->staged) {
    int bytes = (int)dm.stage_bytes;
    CU(cudaFuncSetAttribute(score_scalar_kernel<MODE_CHANGE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CU(cudaFuncSetAttribute(change_step_kernel<true, InterpProg>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CU(cudaFuncSetAttribute(score_scalar_kernel<MODE_SWAP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CU(cudaFuncSetAttribute(score_scalar_kernel<MODE_COMPOUND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  }
  // monomorphised scalar program: every constraint that reacts to scalar edits must be one of the
  // specialised kinds and the sorted tuple one of the instantiated ones; otherwise the interpreter
  ctx->spec_id = -1;
  if (ctx->staged && dm.n_values > 0 && !ctx->force_generic && !getenv("SFGPU_NO_SPEC")) {
    std::vector<std::pair<int, int>> sc;  // (kind, index)
    bool ok = true;
    for (uint32_t k = 0; k < dm.n_cons; ++k) {
      const ConsDev& c = dm.cons[k];
      switch (c.kind) {
        case SFGPU_K_UNI: sc.push_back({(!c.g0 && !c.g1) ? SPEC_K_UNI_CONST : SFGPU_K_UNI, (int)k}); break;
        case SFGPU_K_PAIR_KEY_EQUAL:
          if (c.pad != 2) ok = false;  // tri / quad / penta joins keep the interpreter
          sc.push_back({c.kind, (int)k});
          break;
        case SFGPU_K_GROUP: sc.push_back({c.kind, (int)k}); break;
        case SFGPU_K_PAIR_CSR_EQUAL:
          if (c.off0 == 0xFFFFFFFFu) ok = false;  // no retained partner-value counts
          sc.push_back({c.kind, (int)k});
          break;
        case SFGPU_K_LOAD_BALANCE: case SFGPU_K_RUNS: case SFGPU_K_PROJECT_GROUP: ok = false; break;
        default: break;  // list-only kinds do not react to scalar edits
      }
    }
    if (ok && !sc.empty() && sc.size() <= 4) {
      std::stable_sort(sc.begin(), sc.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
      int kinds[4] = {0, 0, 0, 0};
      for (size_t i = 0; i < sc.size(); ++i) kinds[i] = sc[i].first;
      for (size_t t = 0; t < sizeof(g_spec) / sizeof(g_spec[0]); ++t) {
        if (memcmp(g_spec[t].k, kinds, sizeof(kinds)) != 0) continue;
        ctx->spec_id = (int)t;
        for (size_t i = 0; i < 4; ++i) ctx->spec_idx.k[i] = i < sc.size() ? sc[i].second : -1;
        CU(cudaFuncSetAttribute((const void*)g_spec[t].score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dm.stage_bytes));
        CU(cudaFuncSetAttribute((const void*)g_spec[t].step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dm.stage_bytes));
        break;
      }
    }
  }
  return SFGPU_OK;
}

// entities per CTA of the generated ChangeMove step: amortise the state staging, but cover the machine when
// replicas are few
void sfgpu_change_step_chunks(const sfgpu_ctx* ctx, uint32_t* out_per, uint32_t* out_chunks) {
  const DevModel& dm = ctx->dm;
  uint32_t per = 2048;
  while (per > 256 && (uint64_t)((dm.n_entities + per - 1) / per) * dm.R < (uint64_t)ctx->sm_count * 4) per /= 2;
  *out_per = per;
  *out_chunks = (dm.n_entities + per - 1) / per;
}


// kind: 0 change, 1 swap, 2 compound (ScoreKind of sfgpu_api.cu)
int sfgpu_launch_score_scalar(sfgpu_ctx* ctx, int kind, uint64_t n_total, const uint64_t* d_offs, const uint32_t* d_rows,
                              const uint64_t* d_edit_offs, int64_t* d_scores, uint8_t* d_doable) {
  const DevModel& dm = ctx->dm;
  const uint32_t threads = 256;
  dim3 grid(chunks_for(ctx, n_total, dm.R, threads), dm.R);
  size_t smem = ctx->staged ? dm.stage_bytes : 0;
  ev_begin(ctx);
#define LAUNCH_SCALAR(MODE)                                                                                    \
  if (ctx->staged)                                                                                             \
    score_scalar_kernel<MODE, true><<<grid, threads, smem, ctx->stream>>>(dm, d_offs, d_rows, d_edit_offs,    \
                                                                          d_scores, d_doable);                 \
  else                                                                                                         \
    score_scalar_kernel<MODE, false><<<grid, threads, 0, ctx->stream>>>(dm, d_offs, d_rows, d_edit_offs,      \
                                                                        d_scores, d_doable)
  switch (kind) {
    case 0:
      if (ctx->spec_id >= 0)
        g_spec[ctx->spec_id].score<<<grid, threads, smem, ctx->stream>>>(dm, ctx->spec_idx, d_offs, d_rows, d_scores, d_doable);
      else
        LAUNCH_SCALAR(MODE_CHANGE);
      break;
    case 1: LAUNCH_SCALAR(MODE_SWAP); break;
    default: LAUNCH_SCALAR(MODE_COMPOUND); break;
  }
  ev_end(ctx);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}

// generate + score + forage (two kernels) on the context's stream; the dominant kernel is bracketed by the
// event ring when `timed`
int sfgpu_launch_change_step(sfgpu_ctx* ctx, const ChangeStepArgs& a, uint32_t chunks, uint32_t* d_idx, int64_t* d_best,
                             uint32_t* d_eval, uint32_t* d_win) {
  const DevModel& dm = ctx->dm;
  dim3 grid(chunks, dm.R);
  if (ctx->spec_id >= 0)
    g_spec[ctx->spec_id].step<<<grid, 256, dm.stage_bytes, ctx->stream>>>(dm, a, ctx->spec_idx);
  else if (ctx->staged)
    change_step_kernel<true, InterpProg><<<grid, 256, dm.stage_bytes, ctx->stream>>>(dm, a, ctx->spec_idx);
  else
    change_step_kernel<false, InterpProg><<<grid, 256, 0, ctx->stream>>>(dm, a, ctx->spec_idx);
  change_finish_kernel<<<dm.R, 256, 0, ctx->stream>>>(dm, a, chunks, d_idx, d_best, d_eval, d_win);
  ctx->launches += 2;
  CU(cudaGetLastError());
  return SFGPU_OK;
}

extern "C" {

// ------------------------------------------------------------------------------------------
// Whole step for scalar models: ChangeMove neighbourhood generation + scoring + forager on device.
int32_t sfgpu_step_change(sfgpu_ctx* ctx, uint32_t flags, const sfgpu_forage_params* params,
                          const uint64_t* step_seeds, const int64_t* ref_scores, uint64_t* out_cand_offsets,
                          uint32_t* out_rows, int64_t* out_scores, uint8_t* out_doable, uint32_t* out_index,
                          int64_t* out_best, uint32_t* out_evaluated, uint32_t* out_winner_rows,
                          int32_t apply_winners) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  const DevModel& dm = ctx->dm;
  if (!params || !out_index || !out_best) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  if (!dm.has_scalar) return fail(ctx, SFGPU_E_STATE, "model has no scalar variable");
  if ((uint64_t)dm.n_entities * (dm.n_values + 1) >= 0xFFFFFFFFull)
    return fail(ctx, SFGPU_E_UNSUPPORTED, "neighbourhood too large for 32-bit pull indices");
  if (params->acceptor < 0 || params->acceptor > 3 || params->tie_mode < 0 || params->tie_mode > 1)
    return fail(ctx, SFGPU_E_INVALID, "bad forage params");
  if (params->acceptor != 0 && !ref_scores) return fail(ctx, SFGPU_E_INVALID, "acceptor needs ref_scores");
  if (out_scores && (!out_rows || !out_doable)) return fail(ctx, SFGPU_E_INVALID, "out_scores needs out_rows and out_doable");
  CU(cudaSetDevice(ctx->device));
  const uint32_t R = dm.R;
  const bool dev_io = (flags & SFGPU_DEVICE_IO) != 0;
  SmallIo io;
  rc = small_io_begin(ctx, io, dev_io, 8, step_seeds, ref_scores, out_index, out_best, out_evaluated, out_winner_rows);
  if (rc) return rc;
  if (dev_io && apply_winners && !io.d_win)
    return fail(ctx, SFGPU_E_INVALID, "apply_winners needs out_winner_rows on the device path");
  ChangeStepArgs a{};
  a.f = ForageDev{params->acceptor, params->tie_mode, params->accepted_limit};
  a.step_seeds = io.d_seeds;
  a.ref_scores = io.d_ref;
  a.out_rows = out_rows;
  a.out_scores = out_scores;
  a.out_doable = out_doable;
  a.out_offsets = out_cand_offsets;
  uint32_t per = 0, chunks = 0;
  sfgpu_change_step_chunks(ctx, &per, &chunks);
  a.ents_per_cta = per;
  rc = ensure_partials(ctx, (size_t)R * chunks * sizeof(ChunkPartial));
  if (rc) return rc;
  a.partials = (ChunkPartial*)ctx->partials;
  ev_begin(ctx);
  rc = sfgpu_launch_change_step(ctx, a, chunks, io.d_idx, io.d_best, io.d_eval, io.d_win);
  ev_end(ctx);
  if (rc) return rc;
  if (apply_winners) {
    rc = sfgpu_launch_apply_scalar(ctx, 0, io.d_win, nullptr, nullptr, nullptr);
    if (rc) return rc;
  }
  return small_io_end(ctx, io, out_index, out_best, out_evaluated, out_winner_rows);
} SFGPU_API_CATCH(ctx)

}  // extern "C"
