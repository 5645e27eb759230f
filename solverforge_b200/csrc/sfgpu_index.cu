// Neighbourhoods enumerated on device by pull index (sfgpu_index_step.cuh): SublistChange, SublistSwap, ListReverse.
#include "sfgpu_ctx.hpp"
#include "sfgpu_index_step.cuh"

using namespace sfgpu_host;

namespace {
// Whole step over a neighbourhood enumerated on device by pull index (sfgpu_index_step.cuh). NB = the decoder,
// apply_kind = the apply_list_kernel kind of its rows, ub = an upper bound of the candidates of one replica.
template <class NB>
int step_index_neighbourhood(sfgpu_ctx* ctx, uint32_t flags, uint32_t min_size, uint32_t max_size, int apply_kind, uint64_t ub,
                             const sfgpu_forage_params* params, const uint64_t* step_seeds, const int64_t* ref_scores,
                             uint32_t* out_index, int64_t* out_best, uint32_t* out_evaluated, uint32_t* out_winner_rows,
                             int32_t apply_winners) {
  int rc = check_committed(ctx);
  if (rc) return rc;
  DevModel dm = ctx->dm;
  if (ctx->force_generic) dm.fast_list = 0;  // SFGPU_CTX_GENERIC_KERNELS: cursor walks with the generic delta
  if (!params || !out_index || !out_best) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  if (!dm.has_list) return fail(ctx, SFGPU_E_STATE, "model has no list variable");
  if (min_size < 1 || max_size < min_size || max_size > 255)
    return fail(ctx, SFGPU_E_INVALID, "segment sizes must satisfy 1 <= min <= max <= 255");
  if (params->acceptor < 0 || params->acceptor > 3 || params->tie_mode < 0 || params->tie_mode > 1)
    return fail(ctx, SFGPU_E_INVALID, "bad forage params");
  if (params->acceptor != 0 && !ref_scores) return fail(ctx, SFGPU_E_INVALID, "acceptor needs ref_scores");
  if (ub >= 0xFFFFFFFFull || dm.elem_cap >= (1u << 24))
    return fail(ctx, SFGPU_E_UNSUPPORTED, "neighbourhood too large for 32-bit pull indices");
  const size_t table_bytes = NB::table_words(dm.n_owners, dm.elem_cap) * 4;
  static const bool no_stage = getenv("SFGPU_INDEX_UNSTAGED") != nullptr;  // tuning knob
  const bool staged = !no_stage && ctx->staged && dm.stage_bytes + table_bytes + 1024 <= (size_t)ctx->max_smem_optin;
  if (table_bytes + 1024 > (size_t)ctx->max_smem_optin)
    return fail(ctx, SFGPU_E_UNSUPPORTED, "list variable too large for the shared-memory index tables");
  CU(cudaSetDevice(ctx->device));
  const uint32_t R = dm.R;
  const bool dev_io = (flags & SFGPU_DEVICE_IO) != 0;
  SmallIo io;
  rc = small_io_begin(ctx, io, dev_io, 16, step_seeds, ref_scores, out_index, out_best, out_evaluated, out_winner_rows);
  if (rc) return rc;
  if (dev_io && apply_winners && !io.d_win)
    return fail(ctx, SFGPU_E_INVALID, "apply_winners needs out_winner_rows on the device path");
  uint32_t *d_idx = io.d_idx, *d_eval = io.d_eval, *d_win = io.d_win;
  int64_t* d_best = io.d_best;
  IndexStepArgs a{};
  a.f = ForageDev{params->acceptor, params->tie_mode, params->accepted_limit};
  a.min_size = min_size;
  a.max_size = max_size;
  a.step_seeds = io.d_seeds;
  a.ref_scores = io.d_ref;
  // candidates per CTA: amortise staging + table build, but cover the machine when replicas are few
  uint32_t per = 16384;
  while (per > 1024 && ((ub + per - 1) / per) * R < (uint64_t)ctx->sm_count * 4) per /= 2;
  a.per_chunk = per;
  const uint32_t chunks = (uint32_t)std::min<uint64_t>((ub + per - 1) / per, 65535);
  if ((uint64_t)chunks * per < ub) return fail(ctx, SFGPU_E_UNSUPPORTED, "neighbourhood needs more than 65535 chunks");
  rc = ensure_partials(ctx, (size_t)R * chunks * sizeof(ChunkPartial));
  if (rc) return rc;
  a.partials = (ChunkPartial*)ctx->partials;
  dim3 grid(chunks, R);
  ev_begin(ctx);
  if (staged) {
    const int bytes = (int)(dm.stage_bytes + table_bytes);
    CU(cudaFuncSetAttribute(index_step_kernel<true, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    index_step_kernel<true, NB><<<grid, 256, bytes, ctx->stream>>>(dm, a);
  } else {
    CU(cudaFuncSetAttribute(index_step_kernel<false, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table_bytes));
    index_step_kernel<false, NB><<<grid, 256, table_bytes, ctx->stream>>>(dm, a);
  }
  ev_end(ctx);
  CU(cudaFuncSetAttribute(index_finish_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table_bytes));
  index_finish_kernel<NB><<<R, 256, table_bytes, ctx->stream>>>(dm, a, chunks, d_idx, d_best, d_eval, d_win);
  ctx->launches += 2;
  CU(cudaGetLastError());
  if (apply_winners) {
    rc = sfgpu_launch_apply_list(ctx, apply_kind, d_win, nullptr, nullptr, nullptr);
    if (rc) return rc;
  }
  return small_io_end(ctx, io, out_index, out_best, out_evaluated, out_winner_rows);
}
}  // namespace

extern "C" {

int32_t sfgpu_step_sublist_change(sfgpu_ctx* ctx, uint32_t flags, uint32_t min_size, uint32_t max_size,
                                  const sfgpu_forage_params* params, const uint64_t* step_seeds,
                                  const int64_t* ref_scores, uint32_t* out_index, int64_t* out_best,
                                  uint32_t* out_evaluated, uint32_t* out_winner_rows, int32_t apply_winners) try {
  if (!ctx) return SFGPU_E_INVALID;
  // every element starts at most (max - min + 1) segments, each with fewer than elements + entities destinations
  const DevModel& dm = ctx->dm;
  const uint64_t ub = (uint64_t)dm.elem_cap * (max_size >= min_size ? max_size - min_size + 1 : 1) *
                      ((uint64_t)dm.elem_cap + dm.n_owners);
  return step_index_neighbourhood<SublistChangeNb>(ctx, flags, min_size, max_size, 5, ub, params, step_seeds, ref_scores,
                                                   out_index, out_best, out_evaluated, out_winner_rows, apply_winners);
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_step_list_reverse(sfgpu_ctx* ctx, uint32_t flags, const sfgpu_forage_params* params,
                                const uint64_t* step_seeds, const int64_t* ref_scores, uint32_t* out_index,
                                int64_t* out_best, uint32_t* out_evaluated, uint32_t* out_winner_rows,
                                int32_t apply_winners) try {
  if (!ctx) return SFGPU_E_INVALID;
  const DevModel& dm = ctx->dm;
  const uint64_t ub = (uint64_t)dm.elem_cap * dm.elem_cap / 2 + 1;  // one list holding every element
  return step_index_neighbourhood<ReverseNb>(ctx, flags, 1, 1, 4, ub, params, step_seeds, ref_scores, out_index, out_best,
                                             out_evaluated, out_winner_rows, apply_winners);
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_step_sublist_swap(sfgpu_ctx* ctx, uint32_t flags, uint32_t min_size, uint32_t max_size,
                                const sfgpu_forage_params* params, const uint64_t* step_seeds,
                                const int64_t* ref_scores, uint32_t* out_index, int64_t* out_best,
                                uint32_t* out_evaluated, uint32_t* out_winner_rows, int32_t apply_winners) try {
  if (!ctx) return SFGPU_E_INVALID;
  // unordered pairs of segments: fewer than (segments)^2 / 2 + segments
  const DevModel& dm = ctx->dm;
  const uint64_t segs = (uint64_t)dm.elem_cap * (max_size >= min_size ? max_size - min_size + 1 : 1);
  const uint64_t ub = segs * segs / 2 + segs;
  return step_index_neighbourhood<SublistSwapNb>(ctx, flags, min_size, max_size, 6, ub, params, step_seeds, ref_scores,
                                                 out_index, out_best, out_evaluated, out_winner_rows, apply_winners);
} SFGPU_API_CATCH(ctx)

}  // extern "C"
