// List-move scoring kernels of libsfgpu: the generic constraint-interpreting kernels for every list move kind,
// the fast ListChange kernel of CVRP-shaped programs, and the fused score + forager step over resident rows.
#include "sfgpu_ctx.hpp"
#include "sfgpu_kernels.cuh"

using namespace sfgpu_host;

int sfgpu_launch_forage_finish(sfgpu_ctx* ctx, const ForageArgs& fa, uint32_t chunks, const uint64_t* d_offs,
                               const uint32_t* d_rows, const int64_t* d_scores, const uint8_t* d_doable,
                               const uint64_t* d_seeds, uint32_t* d_idx, int64_t* d_best, uint32_t* d_eval);
int sfgpu_launch_argbest_ordered(sfgpu_ctx* ctx, const ForageDev& f, const uint64_t* d_offs, const int64_t* d_scores,
                                 const uint8_t* d_doable, const uint64_t* d_seeds, const int64_t* d_ref, uint32_t* d_idx,
                                 int64_t* d_best, uint32_t* d_eval);

int sfgpu_configure_list(sfgpu_ctx* ctx) {
  DevModel& dm = ctx->dm;
  if (ctx->staged && dm.has_list) {
    int bytes = (int)dm.stage_bytes;
    CU(cudaFuncSetAttribute(score_list_kernel<LMODE_CHANGE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CU(cudaFuncSetAttribute(score_list_kernel<LMODE_SWAP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CU(cudaFuncSetAttribute(score_list_kernel<LMODE_REVERSE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CU(cudaFuncSetAttribute(score_list_kernel<LMODE_SUBLIST_CHANGE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CU(cudaFuncSetAttribute(score_list_kernel<LMODE_SUBLIST_SWAP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CU(cudaFuncSetAttribute(score_list_kernel<LMODE_K_OPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  }
  if (dm.fast_list) {
    {
      int bytes = (int)dm.fast_stage_bytes;
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<-1, 2, 3, false, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<-1, 2, 4, false, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<-1, 2, 3, true, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<-1, 2, 4, true, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_CONST, 2, 3, false, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_CONST, 2, 4, false, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_CONST, 2, 3, true, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_CONST, 2, 4, true, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_LINEAR, 2, 3, false, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_LINEAR, 2, 4, false, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_LINEAR, 2, 3, true, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_LINEAR, 2, 4, true, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_SQUARE, 2, 3, false, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_SQUARE, 2, 4, false, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_SQUARE, 2, 3, true, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_SQUARE, 2, 4, true, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_EXCESS, 2, 3, false, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_EXCESS, 2, 4, false, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_EXCESS, 2, 3, true, int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CU(cudaFuncSetAttribute(score_list_change_fast_kernel<SFGPU_W_EXCESS, 2, 4, true, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    }
  }
  if (dm.fast_list && dm.compact_bytes) {
    const int bytes = (int)(16 + dm.compact_bytes);
#define COMPACT_ATTR(FN)                                                                                              \
  CU(cudaFuncSetAttribute(score_list_change_fast_kernel<FN, 1, 4, false, uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
  CU(cudaFuncSetAttribute(score_list_change_fast_kernel<FN, 1, 4, true, uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
      COMPACT_ATTR(-1);
      COMPACT_ATTR(SFGPU_W_CONST);
      COMPACT_ATTR(SFGPU_W_LINEAR);
      COMPACT_ATTR(SFGPU_W_SQUARE);
      COMPACT_ATTR(SFGPU_W_EXCESS);
  }
  return SFGPU_OK;
}

// kind: 3 list change, 4 list swap, 5 list reverse, 6 sublist change, 7 sublist swap (ScoreKind of sfgpu_api.cu).
// forage != nullptr (fast list path only): the kernel also emits per-chunk forager partials into
// forage->partials and *out_chunks receives the chunk count the finishing kernel needs.
int sfgpu_launch_score_list(sfgpu_ctx* ctx, int kind, uint64_t n_total, const uint64_t* d_offs, const uint32_t* d_rows,
                            int64_t* d_scores, uint8_t* d_doable, ForageArgs* forage, uint32_t* out_chunks) {
  const DevModel& dm = ctx->dm;
  const uint32_t threads = 256;
  dim3 grid(chunks_for(ctx, n_total, dm.R, threads), dm.R);
  size_t smem = ctx->staged ? dm.stage_bytes : 0;
  ev_begin(ctx);
#define LAUNCH_LIST(MODE)                                                                                      \
  if (ctx->staged)                                                                                             \
    score_list_kernel<MODE, true><<<grid, threads, smem, ctx->stream>>>(dm, d_offs, d_rows, d_scores, d_doable); \
  else                                                                                                         \
    score_list_kernel<MODE, false><<<grid, threads, 0, ctx->stream>>>(dm, d_offs, d_rows, d_scores, d_doable)
  switch (kind) {
    case 3:
      if (dm.fast_list && !ctx->force_generic) {
        // contiguous chunk per CTA; fewer, fatter CTAs amortise the 16 B/record staging
        uint64_t per_replica = (n_total + dm.R - 1) / dm.R;
        // ~10k candidates per CTA: the 35 KB record staging is paid once per CTA (measured: 2 chunks of
        // 10 000 beat 4 x 5 000 by 7 %); with few replicas split further until the machine is covered
        uint32_t chunks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((per_replica + 10239) / 10240, 64));
        while ((uint64_t)chunks * dm.R < (uint64_t)ctx->sm_count * 6 && (uint64_t)chunks * 512 < per_replica) chunks *= 2;
        static const int chunk_override = getenv("SFGPU_FAST_CHUNKS") ? atoi(getenv("SFGPU_FAST_CHUNKS")) : 0;  // tuning knob
        if (chunk_override > 0) chunks = (uint32_t)chunk_override;
        dim3 fgrid(chunks, dm.R);
        if (forage) {
          size_t need = (size_t)chunks * dm.R * sizeof(ChunkPartial);
          if (need > ctx->partials_bytes) {
            if (ctx->partials) cudaFree(ctx->partials);
            ctx->partials = nullptr;
            ctx->partials_bytes = 0;
            CU(cudaMalloc(&ctx->partials, need));
            ctx->partials_bytes = need;
          }
          forage->partials = (ChunkPartial*)ctx->partials;
          if (out_chunks) *out_chunks = chunks;
        }
        size_t fsm = dm.fast_score_bytes;
        int fn = dm.fast_ls >= 0 ? dm.cons[dm.fast_ls].w.fn : -1;
#define FASTK(FN)                                                                                              \
  if (dm.compact_bytes && dm.fm_u16) {                                                                         \
    const size_t csm = 16 + dm.compact_bytes;                                                                  \
    if (forage) score_list_change_fast_kernel<FN, 1, 4, true, uint16_t, true><<<fgrid, threads, csm, ctx->stream>>>(dm, d_offs, d_rows, d_scores, d_doable, *forage); \
    else score_list_change_fast_kernel<FN, 1, 4, false, uint16_t, true><<<fgrid, threads, csm, ctx->stream>>>(dm, d_offs, d_rows, d_scores, d_doable, ForageArgs{}); \
  } else if (forage)                                                                                           \
    if (dm.fm_u16) score_list_change_fast_kernel<FN, 2, 4, true, uint16_t><<<fgrid, threads, fsm, ctx->stream>>>(dm, d_offs, d_rows, d_scores, d_doable, *forage); \
    else score_list_change_fast_kernel<FN, 2, 3, true, int32_t><<<fgrid, threads, fsm, ctx->stream>>>(dm, d_offs, d_rows, d_scores, d_doable, *forage); \
  else                                                                                                         \
    if (dm.fm_u16) score_list_change_fast_kernel<FN, 2, 4, false, uint16_t><<<fgrid, threads, fsm, ctx->stream>>>(dm, d_offs, d_rows, d_scores, d_doable, ForageArgs{}); \
    else score_list_change_fast_kernel<FN, 2, 3, false, int32_t><<<fgrid, threads, fsm, ctx->stream>>>(dm, d_offs, d_rows, d_scores, d_doable, ForageArgs{})
        switch (fn) {
          case -1: FASTK(-1); break;
          case SFGPU_W_CONST: FASTK(SFGPU_W_CONST); break;
          case SFGPU_W_LINEAR: FASTK(SFGPU_W_LINEAR); break;
          case SFGPU_W_SQUARE: FASTK(SFGPU_W_SQUARE); break;
          default: FASTK(SFGPU_W_EXCESS); break;
        }
      } else {
        LAUNCH_LIST(LMODE_CHANGE);
      }
      break;
    case 4: LAUNCH_LIST(LMODE_SWAP); break;
    case 5: LAUNCH_LIST(LMODE_REVERSE); break;
    case 6: LAUNCH_LIST(LMODE_SUBLIST_CHANGE); break;
    case 8: LAUNCH_LIST(LMODE_K_OPT); break;
    default: LAUNCH_LIST(LMODE_SUBLIST_SWAP); break;
  }
  ev_end(ctx);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}

extern "C" {

// ------------------------------------------------------------------------------------------
// Fused step: score every candidate and replay acceptor + forager in one call. Device pointers only.
int32_t sfgpu_step_list_change(sfgpu_ctx* ctx, uint64_t n_candidates, const uint64_t* cand_offsets,
                               const uint32_t* rows, const sfgpu_forage_params* params, const uint64_t* step_seeds,
                               const int64_t* ref_scores, int64_t* out_scores, uint8_t* out_doable,
                               uint32_t* out_index, int64_t* out_best, uint32_t* out_evaluated) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (!cand_offsets || !rows || !params || !out_index || !out_best) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  if ((out_scores == nullptr) != (out_doable == nullptr))
    return fail(ctx, SFGPU_E_INVALID, "out_scores and out_doable are given together or not at all");
  if (params->acceptor < 0 || params->acceptor > 3 || params->tie_mode < 0 || params->tie_mode > 1)
    return fail(ctx, SFGPU_E_INVALID, "bad forage params");
  if (params->acceptor != 0 && !ref_scores) return fail(ctx, SFGPU_E_INVALID, "acceptor needs ref_scores");
  const DevModel& dm = ctx->dm;
  if (!dm.has_list) return fail(ctx, SFGPU_E_STATE, "model has no list variable");
  CU(cudaSetDevice(ctx->device));
  const bool fused = dm.fast_list && !ctx->force_generic && params->accepted_limit == 0;
  if (fused) {
    ForageArgs fa{};
    fa.f = ForageDev{params->acceptor, params->tie_mode, params->accepted_limit};
    fa.ref_scores = ref_scores;
    uint32_t chunks = 0;
    rc = sfgpu_launch_score_list(ctx, 3, n_candidates, cand_offsets, rows, out_scores, out_doable, &fa, &chunks);
    if (rc) return rc;
    return sfgpu_launch_forage_finish(ctx, fa, chunks, cand_offsets, rows, out_scores, out_doable, step_seeds, out_index,
                                      out_best, out_evaluated);
  }
  // unfused: materialise scores (caller's buffers or internal scratch), then the ordered replay kernel
  int64_t* d_scores = out_scores;
  uint8_t* d_doable = out_doable;
  if (!d_scores) {
    size_t need = n_candidates * 16 + (n_candidates + 15) / 16 * 16;
    rc = ensure_staging(ctx, 64, need);
    if (rc) return rc;
    d_scores = (int64_t*)ctx->dscr;
    d_doable = (uint8_t*)ctx->dscr + n_candidates * 16;
  }
  rc = sfgpu_launch_score_list(ctx, 3, n_candidates, cand_offsets, rows, d_scores, d_doable, nullptr, nullptr);
  if (rc) return rc;
  ForageDev f{params->acceptor, params->tie_mode, params->accepted_limit};
  return sfgpu_launch_argbest_ordered(ctx, f, cand_offsets, d_scores, d_doable, step_seeds, ref_scores, out_index, out_best,
                                      out_evaluated);
} SFGPU_API_CATCH(ctx)

}  // extern "C"
