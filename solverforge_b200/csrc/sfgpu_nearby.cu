// Device-side nearby neighbourhoods of libsfgpu (sfgpu_nearby.cuh): generation + scoring + forager replay.
#include "sfgpu_ctx.hpp"
#include "sfgpu_nearby.cuh"

using namespace sfgpu_host;

int sfgpu_configure_nearby(sfgpu_ctx* ctx) {
  DevModel& dm = ctx->dm;
  int bytes = (int)dm.fast_stage_bytes;
#define NB_ATTR(FN, KEY, CELL)                                                                                        \
  CU(cudaFuncSetAttribute(nearby_step_kernel<FN, KEY, CELL>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
  CU(cudaFuncSetAttribute(nearby_step_kernel<FN, KEY, CELL, MOVE_SWAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
#define NB_ATTR4(FN) NB_ATTR(FN, uint32_t, uint16_t); NB_ATTR(FN, uint32_t, int32_t); NB_ATTR(FN, uint64_t, uint16_t); NB_ATTR(FN, uint64_t, int32_t)
    NB_ATTR4(-1);
    NB_ATTR4(SFGPU_W_SQUARE);
    NB_ATTR4(SFGPU_W_EXCESS);
#define NBC_ATTR(FN)                                                                                                    \
  CU(cudaFuncSetAttribute(nearby_step_cached_kernel<FN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));      \
  CU(cudaFuncSetAttribute(nearby_step_cached_kernel<FN, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
  NBC_ATTR(-1);
  NBC_ATTR(SFGPU_W_SQUARE);
  NBC_ATTR(SFGPU_W_EXCESS);
  CU(cudaFuncSetAttribute(nearby_regen_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CU(cudaFuncSetAttribute(nearby_regen_kernel<SFGPU_W_SQUARE>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CU(cudaFuncSetAttribute(nearby_regen_kernel<SFGPU_W_EXCESS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return SFGPU_OK;
}

// Retained neighbourhood buffers (NearbyArgs::c_*): sized for max_nearby K, allocated outside any stream capture by the
// callers of sfgpu_launch_nearby; a reallocation invalidates every replica's cache.
int sfgpu_nearby_prepare_cache(sfgpu_ctx* ctx, uint32_t K) {
  const DevModel& dm = ctx->dm;
  if (!dm.nbc_tag || ctx->nbc_off || !ctx->nb_key32 || !dm.fm_u16 || dm.n_elem_rows > 32768 || K <= ctx->nbc_K) return SFGPU_OK;
  const size_t n_src = (size_t)dm.R * dm.n_elem_rows;
  if (ctx->nbc_buf) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->nbc_buf);
  }
  ctx->nbc_buf = nullptr;
  ctx->nbc_K = 0;
  const size_t work_bytes = (size_t)dm.R * (dm.elem_cap + 1) * 4;
  CU(cudaMalloc(&ctx->nbc_buf, n_src * (2 * sizeof(uint4) + (size_t)K * (sizeof(int2) + 4)) + work_bytes));
  CU(cudaMemsetAsync((char*)ctx->nbc_buf + n_src * (2 * sizeof(uint4) + (size_t)K * (sizeof(int2) + 4)), 0, work_bytes, ctx->stream));
  ctx->nbc_K = K;
  CU(cudaMemsetAsync(dm.nbc_tag, 0, (size_t)dm.R * NBC_WORDS * 4, ctx->stream));
  return SFGPU_OK;
}

// generate + score + forage (two kernels) on the context's stream
int sfgpu_launch_nearby(sfgpu_ctx* ctx, NearbyArgs& a, uint32_t* d_idx, int64_t* d_best, uint32_t* d_eval,
                        uint32_t* d_win, int move) {
  const DevModel& dm = ctx->dm;
  const uint32_t R = dm.R;
  // sources per CTA: 8 warps, >= 24 sources each when there is enough work
  // few fat CTAs when there are many replicas (amortises the record staging); with few replicas
  // spread the sources over the machine: down to one source per warp
  uint32_t chunks = std::max<uint32_t>(1, std::min<uint32_t>((dm.elem_cap + 255) / 256, 64));
  const uint32_t max_chunks = std::max<uint32_t>(1, (dm.elem_cap + 7) / 8);
  while ((uint64_t)chunks * R < (uint64_t)ctx->sm_count * 2 && chunks * 2 <= max_chunks) chunks *= 2;
  if (const char* cv = getenv("SFGPU_NB_CHUNKS")) chunks = std::max(1, atoi(cv));  // tuning experiments
  dim3 grid(chunks, R);
  size_t smem = dm.fast_stage_bytes;
  int fn = dm.fast_ls >= 0 ? dm.cons[dm.fast_ls].w.fn : -1;
#define NEARBYK(FN, KEY, CELL)                                                                       \
  if (move == MOVE_SWAP) {                                                                           \
    nearby_step_kernel<FN, KEY, CELL, MOVE_SWAP><<<grid, 256, smem, ctx->stream>>>(dm, a);            \
    nearby_finish_kernel<FN, KEY, CELL, MOVE_SWAP><<<R, 256, 0, ctx->stream>>>(dm, a, d_idx, d_best, d_eval, d_win); \
  } else {                                                                                           \
    nearby_step_kernel<FN, KEY, CELL><<<grid, 256, smem, ctx->stream>>>(dm, a);                       \
    nearby_finish_kernel<FN, KEY, CELL><<<R, 256, 0, ctx->stream>>>(dm, a, d_idx, d_best, d_eval, d_win); \
  }
#define NEARBYK4(FN)                                                                                 \
  if (ctx->nb_key32) {                                                                               \
    if (dm.fm_u16) { NEARBYK(FN, uint32_t, uint16_t); } else { NEARBYK(FN, uint32_t, int32_t); }     \
  } else {                                                                                           \
    if (dm.fm_u16) { NEARBYK(FN, uint64_t, uint16_t); } else { NEARBYK(FN, uint64_t, int32_t); }     \
  }
  a.scan_bits = ctx->nb_scan_bits;
  a.cache = 0;
  // retained neighbourhood: ListChange on the narrow uint16 / 32-bit-key program, nothing materialised
  if (move == MOVE_CHANGE && ctx->nb_key32 && dm.fm_u16 && !a.out_rows && ctx->nbc_buf && a.max_nearby <= ctx->nbc_K) {
    const size_t n_src = (size_t)R * dm.n_elem_rows;  // laid out for the allocated K; the kernels stride by the K of the call
    a.c_meta = (uint4*)ctx->nbc_buf;
    a.c_delta = (int2*)((char*)ctx->nbc_buf + n_src * 2 * sizeof(uint4));
    a.c_ident = (uint32_t*)((char*)ctx->nbc_buf + n_src * (2 * sizeof(uint4) + (size_t)ctx->nbc_K * sizeof(int2)));
    a.c_work = (uint32_t*)((char*)ctx->nbc_buf + n_src * (2 * sizeof(uint4) + (size_t)ctx->nbc_K * (sizeof(int2) + 4)));
    a.cache = 1;
    // the reference's default max_nearby (20) with full lists folds its kept deltas fully unrolled; whether a replica's
    // lists are full (count == K) is only known on the device, so the kernel falls back per replica
#define NEARBYC(FN)                                                                                    \
  if (a.max_nearby == 20) nearby_step_cached_kernel<FN, 20><<<grid, 256, smem, ctx->stream>>>(dm, a);  \
  else nearby_step_cached_kernel<FN, 0><<<grid, 256, smem, ctx->stream>>>(dm, a);                      \
  nearby_regen_kernel<FN><<<grid, 256, smem, ctx->stream>>>(dm, a);                                  \
  nearby_finish_kernel<FN, uint32_t, uint16_t><<<R, 256, 0, ctx->stream>>>(dm, a, d_idx, d_best, d_eval, d_win);
    if (fn == SFGPU_W_EXCESS) { NEARBYC(SFGPU_W_EXCESS) }
    else if (fn == SFGPU_W_SQUARE) { NEARBYC(SFGPU_W_SQUARE) }
    else { NEARBYC(-1) }
    return SFGPU_OK;
  }
  if (fn == SFGPU_W_EXCESS) { NEARBYK4(SFGPU_W_EXCESS) }
  else if (fn == SFGPU_W_SQUARE) { NEARBYK4(SFGPU_W_SQUARE) }
  else { NEARBYK4(-1) }  // no LIST_SUM, or a LINEAR / CONST weight whose relocation delta is 0
  return SFGPU_OK;
}

// ------------------------------------------------------------------------------------------
// Whole local-search step on device: nearby list-change neighbourhood generation + scoring + forager.
namespace {
int step_nearby_impl(sfgpu_ctx* ctx, int move, uint32_t flags, uint32_t max_nearby, const sfgpu_forage_params* params,
                     const uint64_t* step_seeds, const int64_t* ref_scores, uint64_t* out_cand_offsets,
                     uint32_t* out_rows, int64_t* out_scores, uint8_t* out_doable, uint32_t* out_index,
                     int64_t* out_best, uint32_t* out_evaluated, uint32_t* out_winner_rows, int32_t apply_winners) {
  int rc = check_committed(ctx);
  if (rc) return rc;
  const DevModel& dm = ctx->dm;
  if (!params || !out_index || !out_best) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  if (!dm.nearby_ok || ctx->force_generic)
    return fail(ctx, SFGPU_E_UNSUPPORTED,
                "device-side nearby neighbourhood needs the fast list program (int32 path-cost matrix as the "
                "distance meter, every cell finite); enumerate on the host and call sfgpu_step_list_change");
  if (max_nearby == 0 || max_nearby > 32) return fail(ctx, SFGPU_E_UNSUPPORTED, "max_nearby must be in [1, 32]");
  if (params->acceptor < 0 || params->acceptor > 3 || params->tie_mode < 0 || params->tie_mode > 1)
    return fail(ctx, SFGPU_E_INVALID, "bad forage params");
  if (params->acceptor != 0 && !ref_scores) return fail(ctx, SFGPU_E_INVALID, "acceptor needs ref_scores");
  if (out_scores && (!out_rows || !out_doable)) return fail(ctx, SFGPU_E_INVALID, "out_scores needs out_rows and out_doable");
  CU(cudaSetDevice(ctx->device));
  const uint32_t R = dm.R;
  const bool dev_io = (flags & SFGPU_DEVICE_IO) != 0;
  rc = ensure_partials(ctx, (size_t)R * dm.elem_cap * sizeof(SrcPartial));
  if (rc) return rc;
  if (move == MOVE_CHANGE && !out_rows) {
    rc = sfgpu_nearby_prepare_cache(ctx, max_nearby);
    if (rc) return rc;
  }
  SmallIo io;
  rc = small_io_begin(ctx, io, dev_io, 16, step_seeds, ref_scores, out_index, out_best, out_evaluated, out_winner_rows);
  if (rc) return rc;
  if (dev_io && apply_winners && !io.d_win)
    return fail(ctx, SFGPU_E_INVALID, "apply_winners needs out_winner_rows on the device path");
  NearbyArgs a{};
  a.f = ForageDev{params->acceptor, params->tie_mode, params->accepted_limit};
  a.max_nearby = max_nearby;
  a.step_seeds = io.d_seeds;
  a.ref_scores = io.d_ref;
  a.partials = (SrcPartial*)ctx->partials;
  a.out_rows = out_rows;
  a.out_scores = out_scores;
  a.out_doable = out_doable;
  a.out_offsets = out_cand_offsets;
  ev_begin(ctx);
  rc = sfgpu_launch_nearby(ctx, a, io.d_idx, io.d_best, io.d_eval, io.d_win, move);
  if (rc) return rc;
  ev_end(ctx);
  ctx->launches += a.cache ? 3 : 2;
  CU(cudaGetLastError());
  if (apply_winners) {
    rc = small_io_results_ready(ctx, io);  // the read-back overlaps the commit
    if (rc) return rc;
    // a replica without a winner carries the sentinel row (owner 0xFFFFFFFF): not doable, skipped
    rc = sfgpu_launch_apply_list(ctx, move == MOVE_SWAP ? 3 : 2, io.d_win, nullptr, nullptr, nullptr);
    if (rc) return rc;
  }
  return small_io_end(ctx, io, out_index, out_best, out_evaluated, out_winner_rows);
}
}  // namespace

extern "C" {

int32_t sfgpu_step_nearby_list_change(sfgpu_ctx* ctx, uint32_t flags, uint32_t max_nearby,
                                      const sfgpu_forage_params* params, const uint64_t* step_seeds,
                                      const int64_t* ref_scores, uint64_t* out_cand_offsets, uint32_t* out_rows,
                                      int64_t* out_scores, uint8_t* out_doable, uint32_t* out_index,
                                      int64_t* out_best, uint32_t* out_evaluated, uint32_t* out_winner_rows,
                                      int32_t apply_winners) try {
  return step_nearby_impl(ctx, MOVE_CHANGE, flags, max_nearby, params, step_seeds, ref_scores, out_cand_offsets,
                          out_rows, out_scores, out_doable, out_index, out_best, out_evaluated, out_winner_rows,
                          apply_winners);
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_step_nearby_list_swap(sfgpu_ctx* ctx, uint32_t flags, uint32_t max_nearby,
                                    const sfgpu_forage_params* params, const uint64_t* step_seeds,
                                    const int64_t* ref_scores, uint64_t* out_cand_offsets, uint32_t* out_rows,
                                    int64_t* out_scores, uint8_t* out_doable, uint32_t* out_index, int64_t* out_best,
                                    uint32_t* out_evaluated, uint32_t* out_winner_rows, int32_t apply_winners) try {
  if (ctx && ctx->dm.fast_pc < 0)
    return fail(ctx, SFGPU_E_UNSUPPORTED, "nearby list swap needs a path-cost constraint (its matrix is the distance meter)");
  return step_nearby_impl(ctx, MOVE_SWAP, flags, max_nearby, params, step_seeds, ref_scores, out_cand_offsets, out_rows,
                          out_scores, out_doable, out_index, out_best, out_evaluated, out_winner_rows, apply_winners);
} SFGPU_API_CATCH(ctx)

}  // extern "C"

// test hook (not part of the ABI): the retained-neighbourhood protocol words of every replica, [R][16]
extern "C" int32_t sfgpu_debug_nearby_cache_tags(sfgpu_ctx* ctx, uint32_t* out_words) {
  if (!ctx || !ctx->dm.nbc_tag) return SFGPU_E_UNSUPPORTED;
  cudaStreamSynchronize(ctx->stream);
  cudaMemcpy(out_words, ctx->dm.nbc_tag, (size_t)ctx->dm.R * NBC_WORDS * 4, cudaMemcpyDeviceToHost);
  return SFGPU_OK;
}
