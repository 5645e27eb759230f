// Generated ChangeMove step of scalar models (sfgpu_change_step.cuh): change_step_kernel over every monomorphised
// program — a translation unit of its own so it compiles side by side with sfgpu_scalar.cu and sfgpu_scalar_finish.cu.
#include "sfgpu_ctx.hpp"
#include "sfgpu_change_step.cuh"
#include "sfgpu_spec_list.h"

using namespace sfgpu_host;

namespace {
typedef void (*SpecStepFn)(const DevModel, const ChangeStepArgs, const SpecIdx);
struct StepEntry {
  SpecStepFn step, step_n;  // int64 / int32 program (or null)
};
#define STEP_WIDE(a, b, c, d) { change_step_kernel<true, SpecProg<a, b, c, d>>, nullptr }
#define STEP_BOTH(a, b, c, d) { change_step_kernel<true, SpecProg<a, b, c, d>>, change_step_kernel<true, SpecProgN<a, b, c, d>> }
const StepEntry g_step[] = {SFGPU_SPEC_TUPLES(STEP_WIDE, STEP_BOTH)};  // same order as g_spec of sfgpu_scalar.cu
#undef STEP_WIDE
#undef STEP_BOTH
}  // namespace

// dynamic shared memory of the step / finish kernels the context will launch (ctx->spec_id / spec_narrow are set)
int sfgpu_configure_scalar_step(sfgpu_ctx* ctx) {
  const DevModel& dm = ctx->dm;
  const int bytes = (int)dm.stage_bytes;
  if (ctx->staged) CU(cudaFuncSetAttribute(change_step_kernel<true, InterpProg>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (ctx->spec_id >= 0) {
    const StepEntry& e = g_step[ctx->spec_id];
    CU(cudaFuncSetAttribute((const void*)e.step, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (ctx->spec_narrow) CU(cudaFuncSetAttribute((const void*)e.step_n, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  }
  return sfgpu_configure_scalar_finish(ctx);
}

// generate + score + forage (two kernels) on the context's stream; the dominant kernel is bracketed by the
// event ring when `timed`
int sfgpu_launch_change_step(sfgpu_ctx* ctx, const ChangeStepArgs& a, uint32_t chunks, uint32_t* d_idx, int64_t* d_best,
                             uint32_t* d_eval, uint32_t* d_win) {
  const DevModel& dm = ctx->dm;
  dim3 grid(chunks, dm.R);
  if (ctx->spec_id >= 0)
    (ctx->spec_narrow ? g_step[ctx->spec_id].step_n : g_step[ctx->spec_id].step)<<<grid, 256, dm.stage_bytes, ctx->stream>>>(dm, a, ctx->spec_idx);
  else if (ctx->staged)
    change_step_kernel<true, InterpProg><<<grid, 256, dm.stage_bytes, ctx->stream>>>(dm, a, ctx->spec_idx);
  else
    change_step_kernel<false, InterpProg><<<grid, 256, 0, ctx->stream>>>(dm, a, ctx->spec_idx);
  ctx->launches += 1;
  CU(cudaGetLastError());
  return sfgpu_launch_change_finish(ctx, a, chunks, d_idx, d_best, d_eval, d_win);
}
