// change_finish_kernel (sfgpu_change_step.cuh) over every monomorphised program: the third translation unit of the
// scalar programs (sfgpu_scalar.cu: rows-resident scoring, sfgpu_scalar_step.cu: generated step).
#include "sfgpu_ctx.hpp"
#include "sfgpu_change_step.cuh"
#include "sfgpu_spec_list.h"

using namespace sfgpu_host;

namespace {
typedef void (*SpecFinishFn)(const DevModel, const ChangeStepArgs, const SpecIdx, uint32_t, uint32_t*, int64_t*, uint32_t*, uint32_t*);
struct FinishEntry {
  SpecFinishFn finish, finish_n;  // int64 / int32 program (or null)
};
#define FIN_WIDE(a, b, c, d) { change_finish_kernel<true, SpecProg<a, b, c, d>>, nullptr }
#define FIN_BOTH(a, b, c, d) { change_finish_kernel<true, SpecProg<a, b, c, d>>, change_finish_kernel<true, SpecProgN<a, b, c, d>> }
const FinishEntry g_finish[] = {SFGPU_SPEC_TUPLES(FIN_WIDE, FIN_BOTH)};  // same order as g_spec of sfgpu_scalar.cu
#undef FIN_WIDE
#undef FIN_BOTH
}  // namespace

int sfgpu_configure_scalar_finish(sfgpu_ctx* ctx) {
  const DevModel& dm = ctx->dm;
  const int bytes = (int)dm.stage_bytes;
  if (ctx->staged) CU(cudaFuncSetAttribute(change_finish_kernel<true, InterpProg>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (ctx->spec_id >= 0) {
    const FinishEntry& e = g_finish[ctx->spec_id];
    CU(cudaFuncSetAttribute((const void*)e.finish, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (ctx->spec_narrow) CU(cudaFuncSetAttribute((const void*)e.finish_n, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  }
  return SFGPU_OK;
}

int sfgpu_launch_change_finish(sfgpu_ctx* ctx, const ChangeStepArgs& a, uint32_t chunks, uint32_t* d_idx, int64_t* d_best,
                               uint32_t* d_eval, uint32_t* d_win) {
  const DevModel& dm = ctx->dm;
  if (ctx->spec_id >= 0)
    (ctx->spec_narrow ? g_finish[ctx->spec_id].finish_n : g_finish[ctx->spec_id].finish)<<<dm.R, 256, dm.stage_bytes, ctx->stream>>>(
        dm, a, ctx->spec_idx, chunks, d_idx, d_best, d_eval, d_win);
  else if (ctx->staged)
    change_finish_kernel<true, InterpProg><<<dm.R, 256, dm.stage_bytes, ctx->stream>>>(dm, a, ctx->spec_idx, chunks, d_idx, d_best, d_eval, d_win);
  else
    change_finish_kernel<false, InterpProg><<<dm.R, 256, 0, ctx->stream>>>(dm, a, ctx->spec_idx, chunks, d_idx, d_best, d_eval, d_win);
  ctx->launches += 1;
  CU(cudaGetLastError());
  return SFGPU_OK;
}
