// Scalar-move kernels of libsfgpu: ChangeMove / SwapMove / CompoundScalarMove batches (interpreter and
// monomorphised programs) and the whole ChangeMove step generated on device.
#include "sfgpu_ctx.hpp"
#include "sfgpu_spec.cuh"
#include "sfgpu_spec_list.h"

using namespace sfgpu_host;

namespace {

// Monomorphised scalar programs: (sorted) kinds of the scalar constraints -> kernel instantiations.
// The tuples of the reference's scalar examples (graph colouring, n-queens, job shop) and their prefixes.
// Every tuple has the wide (int64) kernels; tuples made only of kinds with an int32 form also carry the narrow ones.
typedef void (*SpecScoreFn)(const DevModel, const SpecIdx, const uint64_t*, const uint32_t*, int64_t*, uint8_t*, const ForageArgs);
struct SpecEntry {
  int k[4];
  SpecScoreFn score, fused;      // rows-resident: scores only / scores + forager partials
  SpecScoreFn score_n, fused_n;  // int32 forms, or null
};
#define SPEC_WIDE(a, b, c, d) \
  { {a, b, c, d}, spec_change_kernel<SpecProg<a, b, c, d>, false>, spec_change_kernel<SpecProg<a, b, c, d>, true>, nullptr, nullptr }
#define SPEC_BOTH(a, b, c, d) \
  { {a, b, c, d}, spec_change_kernel<SpecProg<a, b, c, d>, false>, spec_change_kernel<SpecProg<a, b, c, d>, true>, \
    spec_change_kernel<SpecProgN<a, b, c, d>, false>, spec_change_kernel<SpecProgN<a, b, c, d>, true> }
const SpecEntry g_spec[] = {SFGPU_SPEC_TUPLES(SPEC_WIDE, SPEC_BOTH)};
#undef SPEC_WIDE
#undef SPEC_BOTH

// Upper bound of |delta| of one ChangeMove over the scalar constraints of a monomorphised program, or a negative
// value when some quantity the int32 structs compare or store may not fit (then the program stays wide).
// Sums / products wrap in uint32 and are exact whenever the true result fits, so only the FINAL deltas, the
// stored operands (weights, offsets, column values, table cells) and the operands of comparisons need a bound.
double scalar_narrow_bound(const sfgpu_ctx* ctx) {
  const DevModel& dm = ctx->dm;
  const double LIM = 1073741824.0;  // 2^30
  double total = 0;
  for (int i = 0; i < 4; ++i) {
    const int k = ctx->spec_idx.k[i];
    if (k < 0) continue;
    const ConsDev& c = dm.cons[k];
    const sfgpu_constraint_desc& d = ctx->cons[k].d;
    const double a = std::fabs((double)c.w.a), b = std::fabs((double)c.w.b);
    if (a >= LIM || b >= LIM) return -1;
    switch (c.kind) {
      case SFGPU_K_UNI:
        if (c.g0 || c.g1) return -1;  // only the column-free form has an int32 struct
        total += a + b + a * b;  // |weight(0)| for every weight function
        break;
      case SFGPU_K_PAIR_CSR_EQUAL: total += a * (double)dm.n_entities; break;
      case SFGPU_K_PAIR_KEY_EQUAL:
        if (c.g0 && !c.g2) return -1;  // key column does not fit int32
        if (std::fabs((double)c.p0) >= LIM || std::fabs((double)c.p1) >= LIM || std::fabs((double)c.p2) >= LIM) return -1;
        total += a * (double)dm.n_entities;
        break;
      case SFGPU_K_GROUP: {
        if (c.g1) return -1;                      // per-key weight offsets keep the wide struct
        if (c.g0 && !c.g2) return -1;
        if (c.w.fn == SFGPU_W_PAIRS) return -1;   // x (x - 1) / 2 is not a ring expression
        double X = (double)dm.n_entities + 1, xmax = 1;  // largest |group aggregate| after an edit, largest |x|
        if (c.g0) {
          X = 0;
          xmax = 0;
          for (int64_t v : ctx->cols[d.aux0].host) {
            X += std::fabs((double)v);
            xmax = std::max(xmax, std::fabs((double)v));
          }
          X += xmax;
        }
        X = std::max(X, std::fabs((double)c.p1));  // the complement's default result
        if (X + b >= LIM) return -1;               // compared quantity x - b
        double w;
        switch (c.w.fn) {
          case SFGPU_W_CONST: w = a; break;
          case SFGPU_W_LINEAR: w = a * X + b; break;
          case SFGPU_W_SQUARE: w = a * X * X + b; break;
          default: w = a * (X + b); break;  // EXCESS, ABSDIFF
        }
        total += 4 * w + 4 * a * xmax * (xmax + 2 * X);  // generic form / closed square form
        break;
      }
      default: return -1;
    }
  }
  return total;
}

}  // namespace

int sfgpu_configure_scalar(sfgpu_ctx* ctx) {
  DevModel& dm = ctx->dm;
  if (ctx->staged) {
    int bytes = (int)dm.stage_bytes;
    CU(cudaFuncSetAttribute(score_scalar_kernel<MODE_CHANGE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
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
        case SFGPU_K_LOAD_BALANCE: case SFGPU_K_RUNS: case SFGPU_K_PROJECT_GROUP: case SFGPU_K_JOIN_EXPR:
        case SFGPU_K_PAIR_KEY_EXPR: ok = false; break;
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
        CU(cudaFuncSetAttribute((const void*)g_spec[t].fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dm.stage_bytes));
        ctx->spec_narrow = false;
        if (g_spec[t].score_n && !getenv("SFGPU_NO_NARROW")) {
          const double bound = scalar_narrow_bound(ctx);
          if (bound >= 0 && bound < 1073741824.0) {
            ctx->spec_narrow = true;
            CU(cudaFuncSetAttribute((const void*)g_spec[t].score_n, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dm.stage_bytes));
            CU(cudaFuncSetAttribute((const void*)g_spec[t].fused_n, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dm.stage_bytes));
          }
        }
        break;
      }
    }
  }
  return sfgpu_configure_scalar_step(ctx);  // the generated step + finish kernels live in sfgpu_scalar_step.cu
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


// chunks per replica of the rows-resident monomorphised kernel: contiguous chunks, fat enough to amortise the
// staging of the replica block, enough of them to cover the machine when replicas are few
static uint32_t spec_chunks(const sfgpu_ctx* ctx, uint64_t n_total) {
  const DevModel& dm = ctx->dm;
  const uint64_t per_replica = (n_total + dm.R - 1) / dm.R;
  static const int override_chunks = getenv("SFGPU_SPEC_CHUNKS") ? atoi(getenv("SFGPU_SPEC_CHUNKS")) : 0;  // tuning knob
  if (override_chunks > 0) return (uint32_t)override_chunks;
  uint32_t chunks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((per_replica + 8191) / 8192, 64));
  while ((uint64_t)chunks * dm.R < (uint64_t)ctx->sm_count * 6 && (uint64_t)chunks * 1024 < per_replica) chunks *= 2;
  return chunks;
}

// kind: 0 change, 1 swap, 2 compound (ScoreKind of sfgpu_api.cu). forage != nullptr (monomorphised ChangeMove
// programs only): the kernel also emits per-chunk forager partials and *out_chunks receives the chunk count.
int sfgpu_launch_score_scalar(sfgpu_ctx* ctx, int kind, uint64_t n_total, const uint64_t* d_offs, const uint32_t* d_rows,
                              const uint64_t* d_edit_offs, int64_t* d_scores, uint8_t* d_doable, ForageArgs* forage,
                              uint32_t* out_chunks) {
  const DevModel& dm = ctx->dm;
  const uint32_t threads = 256;
  dim3 grid(chunks_for(ctx, n_total, dm.R, threads), dm.R);
  size_t smem = ctx->staged ? dm.stage_bytes : 0;
  if (forage && !(kind == 0 && ctx->spec_id >= 0)) return fail(ctx, SFGPU_E_STATE, "fused forager needs a monomorphised ChangeMove program");
  if (kind == 0 && ctx->spec_id >= 0) {
    const uint32_t chunks = spec_chunks(ctx, n_total);
    const SpecEntry& se = g_spec[ctx->spec_id];
    dim3 sgrid(chunks, dm.R);
    if (forage) {
      int rc = ensure_partials(ctx, (size_t)chunks * dm.R * sizeof(ChunkPartial));
      if (rc) return rc;
      forage->partials = (ChunkPartial*)ctx->partials;
      if (out_chunks) *out_chunks = chunks;
    }
    ev_begin(ctx);
    const SpecScoreFn fn = forage ? (ctx->spec_narrow ? se.fused_n : se.fused) : (ctx->spec_narrow ? se.score_n : se.score);
    fn<<<sgrid, threads, smem, ctx->stream>>>(dm, ctx->spec_idx, d_offs, d_rows, d_scores, d_doable, forage ? *forage : ForageArgs{});
    ev_end(ctx);
    ctx->launches++;
    CU(cudaGetLastError());
    return SFGPU_OK;
  }
  ev_begin(ctx);
#define LAUNCH_SCALAR(MODE)                                                                                    \
  if (ctx->staged)                                                                                             \
    score_scalar_kernel<MODE, true><<<grid, threads, smem, ctx->stream>>>(dm, d_offs, d_rows, d_edit_offs,    \
                                                                          d_scores, d_doable);                 \
  else                                                                                                         \
    score_scalar_kernel<MODE, false><<<grid, threads, 0, ctx->stream>>>(dm, d_offs, d_rows, d_edit_offs,      \
                                                                        d_scores, d_doable)
  switch (kind) {
    case 0: LAUNCH_SCALAR(MODE_CHANGE); break;
    case 1: LAUNCH_SCALAR(MODE_SWAP); break;
    default: LAUNCH_SCALAR(MODE_COMPOUND); break;
  }
  ev_end(ctx);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}

int sfgpu_launch_forage_finish(sfgpu_ctx* ctx, const ForageArgs& fa, uint32_t chunks, const uint64_t* d_offs,
                               const uint32_t* d_rows, const int64_t* d_scores, const uint8_t* d_doable,
                               const uint64_t* d_seeds, uint32_t* d_idx, int64_t* d_best, uint32_t* d_eval);
int sfgpu_launch_argbest_ordered(sfgpu_ctx* ctx, const ForageDev& f, const uint64_t* d_offs, const int64_t* d_scores,
                                 const uint8_t* d_doable, const uint64_t* d_seeds, const int64_t* d_ref, uint32_t* d_idx,
                                 int64_t* d_best, uint32_t* d_eval);

extern "C" {

// ------------------------------------------------------------------------------------------
// Fused step over resident ChangeMove rows: score every candidate and replay acceptor + forager in one pass
// (the scalar counterpart of sfgpu_step_list_change). Device pointers only.
int32_t sfgpu_step_change_rows(sfgpu_ctx* ctx, uint64_t n_candidates, const uint64_t* cand_offsets, const uint32_t* rows,
                               const sfgpu_forage_params* params, const uint64_t* step_seeds, const int64_t* ref_scores,
                               int64_t* out_scores, uint8_t* out_doable, uint32_t* out_index, int64_t* out_best,
                               uint32_t* out_evaluated) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (!cand_offsets || !rows || !params || !out_index || !out_best) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  if ((out_scores == nullptr) != (out_doable == nullptr))
    return fail(ctx, SFGPU_E_INVALID, "out_scores and out_doable are given together or not at all");
  if (params->acceptor < 0 || params->acceptor > 3 || params->tie_mode < 0 || params->tie_mode > 1)
    return fail(ctx, SFGPU_E_INVALID, "bad forage params");
  if (params->acceptor != 0 && !ref_scores) return fail(ctx, SFGPU_E_INVALID, "acceptor needs ref_scores");
  const DevModel& dm = ctx->dm;
  if (!dm.has_scalar) return fail(ctx, SFGPU_E_STATE, "model has no scalar variable");
  CU(cudaSetDevice(ctx->device));
  if (n_candidates == 0) return fail(ctx, SFGPU_E_INVALID, "empty batch");
  if (ctx->spec_id >= 0 && params->accepted_limit == 0) {
    ForageArgs fa{};
    fa.f = ForageDev{params->acceptor, params->tie_mode, params->accepted_limit};
    fa.ref_scores = ref_scores;
    fa.row_kind = 1;
    uint32_t chunks = 0;
    rc = sfgpu_launch_score_scalar(ctx, 0, n_candidates, cand_offsets, rows, nullptr, out_scores, out_doable, &fa, &chunks);
    if (rc) return rc;
    return sfgpu_launch_forage_finish(ctx, fa, chunks, cand_offsets, rows, out_scores, out_doable, step_seeds, out_index,
                                      out_best, out_evaluated);
  }
  // interpreter programs / AcceptedCount(N): materialise the scores, then the ordered replay kernel
  int64_t* d_scores = out_scores;
  uint8_t* d_doable = out_doable;
  if (!d_scores) {
    const size_t need = n_candidates * 16 + (n_candidates + 15) / 16 * 16;
    rc = ensure_staging(ctx, 64, need);
    if (rc) return rc;
    d_scores = (int64_t*)ctx->dscr;
    d_doable = (uint8_t*)ctx->dscr + n_candidates * 16;
  }
  rc = sfgpu_launch_score_scalar(ctx, 0, n_candidates, cand_offsets, rows, nullptr, d_scores, d_doable, nullptr, nullptr);
  if (rc) return rc;
  ForageDev f{params->acceptor, params->tie_mode, params->accepted_limit};
  return sfgpu_launch_argbest_ordered(ctx, f, cand_offsets, d_scores, d_doable, step_seeds, ref_scores, out_index, out_best,
                                      out_evaluated);
} SFGPU_API_CATCH(ctx)

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
