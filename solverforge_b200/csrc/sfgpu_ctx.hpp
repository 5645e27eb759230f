// Host-side context of libsfgpu shared by the translation units of the library (one per kernel family so
// the build parallelises): model under construction, device model, staging buffers, error plumbing.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "sfgpu_dev.cuh"


namespace sfgpu_host {

struct Collection {
  std::string name;
  uint32_t n_rows;
  int32_t descriptor;
};
struct Column {
  uint32_t coll;
  std::vector<int64_t> host;
  int64_t* dev = nullptr;
  int32_t* dev32 = nullptr;  // narrowed copy (uploaded when every value fits int32 with two bits of headroom)
};
struct Csr {
  uint32_t n_rows;
  std::vector<uint32_t> row_ptr, col;
};
struct Matrix {
  uint32_t rows, cols;
  std::vector<int64_t> host;
  void* dev = nullptr;
  bool i32 = false;
};
struct ScalarVar {
  uint32_t coll, n_values;
  int allows_unassigned;
  std::vector<int32_t> init;  // [n] or [R][n]
  bool per_replica = false;
};
struct ListVar {
  uint32_t owner_coll, elem_coll;
  std::vector<uint32_t> offsets, elems;
  bool per_replica = false;
};
struct ExprHost {
  std::vector<sfgpu_expr_op> ops;
};
struct ConsHost {
  sfgpu_constraint_desc d;
  std::string name;
};

}  // namespace sfgpu_host

// buffers and arguments of the union step (sfgpu_union.cu), sized for the largest window
struct UnionPlan {
  UnionArgs a{};
  uint32_t* pending = nullptr;     // [1] replicas that need a larger window
  uint32_t* apply_rows = nullptr;  // [R][4] winner rows
  int32_t* apply_kinds = nullptr;  // [R] apply_list_kernel kind of each winner, -1 = none
  uint32_t* win_state = nullptr;   // [R] adaptive first window of each replica (resident loop)
  uint32_t w0 = 64, wmax = 4096;
  bool configured = false;
  // SimulatedAnnealing in the resident loop (sfgpu_solve.cuh): sa_accept_kernel runs between scoring and the replay
  bool sa = false;
  SaState* sa_cur = nullptr;
  SaState* sa_nxt = nullptr;
  SaParams sa_params{};
};

struct sfgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  // model under construction
  bool building = false, committed = false;
  uint32_t R = 0;
  std::vector<sfgpu_host::Collection> colls;
  std::vector<sfgpu_host::Column> cols;
  std::vector<sfgpu_host::Csr> csrs;
  std::vector<sfgpu_host::Matrix> mats;
  std::vector<sfgpu_host::ScalarVar> svars;
  std::vector<sfgpu_host::ListVar> lvars;
  std::vector<sfgpu_host::ConsHost> cons;
  std::vector<sfgpu_host::ExprHost> exprs;
  const void* expr_cols_dev = nullptr;  // column / CSR pointer tables of the column expressions
  const void* expr_csrs_dev = nullptr;
  // device
  DevModel dm{};
  char* scratch_state = nullptr;
  std::vector<void*> dev_allocs;
  int max_smem_optin = 0;
  int sm_count = 0;
  bool staged = false;
  bool has_load_balance = false;
  bool nb_key32 = false;      // nearby keys fit 32 bits
  uint32_t nb_scan_bits = 24;
  bool force_generic = false;  // SFGPU_CTX_GENERIC_KERNELS: never take a specialised fast path (testing)
  int spec_id = -1;            // monomorphised scalar program (sfgpu_spec.cuh), -1 = interpreter
  bool spec_narrow = false;    // the program runs its deltas in int32 (every delta provably fits, commit-time bound)
  SpecIdx spec_idx{{-1, -1, -1, -1}};
  // staging for host-pointer calls
  void* pin = nullptr;
  size_t pin_bytes = 0;
  void* dscr = nullptr;
  size_t dscr_bytes = 0;
  void* partials = nullptr;  // fused forager chunk partials
  void* nbc_buf = nullptr;   // retained nearby neighbourhood (nearby_step_cached_kernel): deltas, reference elements, meta
  uint32_t nbc_K = 0;        // max_nearby the buffers are sized for
  bool nbc_off = false;      // SFGPU_NO_NBCACHE=1: every step regenerates its whole neighbourhood
  void* solve_buf = nullptr;  // device-resident loop state
  void* union_buf = nullptr;  // union step buffers
  size_t union_bytes = 0;
  UnionPlan union_plan;
  uint64_t argbest_stride_hint = 0;  // rows per replica of the fixed-stride batch the next counted replay runs over
  std::vector<cudaStream_t> aux_streams;  // the children of a union walk side by side (fork / join by events)
  std::vector<cudaEvent_t> aux_events;    // [0] fork, [1 + i] join of aux stream i
  std::vector<uint32_t> relabel_host, inverse_host;  // element id <-> internal id of the fast records
  size_t solve_bytes = 0;
  cudaStream_t io_stream = nullptr;  // early result read-back of a whole-step call, side by side with its commit kernel
  cudaEvent_t io_ready = nullptr, io_done = nullptr;
  void* small_pin = nullptr;  // per-replica seeds / winners of the host-pointer step call
  void* small_dev = nullptr;
  size_t small_bytes = 0;
  size_t partials_bytes = 0;
  void* sync_dev = nullptr;  // best-score sync: {hard, soft, replica} of this rank + the gathered records
  void* sync_pin = nullptr;
  size_t sync_bytes = 0;
  // ring of CUDA event pairs around the dominant (scoring) kernel of each call
  static constexpr uint32_t EV_RING = 512;
  std::vector<cudaEvent_t> ev_a, ev_b;
  uint64_t ev_count = 0;  // scoring launches recorded so far
  uint64_t launches = 0;
};

namespace sfgpu_host {

inline thread_local std::string g_noctx_err;

inline int fail(sfgpu_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_noctx_err = msg;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? SFGPU_E_OOM : SFGPU_E_CUDA,               \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                             \
  } while (0)

// Nothing unwinds across the C boundary: every extern "C" entry point is a function-try-block
// that ends with this handler. Building the message may itself run out of memory: then only the code is returned.
#define SFGPU_API_CATCH(CTX)                                                                        \
  catch (const std::bad_alloc&) {                                                                   \
    try {                                                                                           \
      return sfgpu_host::fail(CTX, SFGPU_E_OOM, "out of host memory");                              \
    } catch (...) {                                                                                 \
      return SFGPU_E_OOM;                                                                           \
    }                                                                                               \
  }                                                                                                 \
  catch (const std::exception& e_) {                                                                \
    try {                                                                                           \
      return sfgpu_host::fail(CTX, SFGPU_E_INVALID, std::string("internal error: ") + e_.what());   \
    } catch (...) {                                                                                 \
      return SFGPU_E_INVALID;                                                                       \
    }                                                                                               \
  }                                                                                                 \
  catch (...) {                                                                                     \
    return SFGPU_E_INVALID;                                                                         \
  }

template <class T>
int dev_upload(sfgpu_ctx* ctx, const T* host, size_t n, T** out) {
  void* p = nullptr;
  CU(cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)));
  ctx->dev_allocs.push_back(p);
  if (n) CU(cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *out = (T*)p;
  return SFGPU_OK;
}

inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }
// dynamic shared memory of apply_list_kernel: the element copy, reused as the owner table of build_fast_records
inline size_t apply_smem_bytes(const DevModel& dm) {  // old element copy / offsets scratch + the new element array
  return ((size_t)std::max(dm.elem_cap, dm.n_owners + 1) + dm.elem_cap) * 4;
}
// apply_list_kernel is one CTA per replica and a chain of short barrier-separated phases: with more replicas than the
// machine holds 256-thread CTAs (4 per SM at 64 registers) half-size CTAs keep every replica resident in one wave
inline unsigned apply_threads(const sfgpu_ctx* ctx) { return ctx->dm.R > 4u * (unsigned)ctx->sm_count ? 128u : 256u; }

inline int ensure_staging(sfgpu_ctx* ctx, size_t pin_bytes, size_t dev_bytes) {
  if (pin_bytes > ctx->pin_bytes) {
    if (ctx->pin) cudaFreeHost(ctx->pin);
    ctx->pin = nullptr;
    ctx->pin_bytes = 0;
    CU(cudaMallocHost(&ctx->pin, pin_bytes));
    ctx->pin_bytes = pin_bytes;
  }
  if (dev_bytes > ctx->dscr_bytes) {
    if (ctx->dscr) cudaFree(ctx->dscr);
    ctx->dscr = nullptr;
    ctx->dscr_bytes = 0;
    CU(cudaMalloc(&ctx->dscr, dev_bytes));
    CU(cudaMemsetAsync(ctx->dscr, 0, dev_bytes, ctx->stream));  // alignment gaps between its arrays travel in read-backs
    ctx->dscr_bytes = dev_bytes;
  }
  return SFGPU_OK;
}

inline void ev_begin(sfgpu_ctx* ctx) {
  if (ctx->ev_a.empty()) {
    ctx->ev_a.resize(sfgpu_ctx::EV_RING);
    ctx->ev_b.resize(sfgpu_ctx::EV_RING);
    for (uint32_t i = 0; i < sfgpu_ctx::EV_RING; ++i) {
      cudaEventCreate(&ctx->ev_a[i]);
      cudaEventCreate(&ctx->ev_b[i]);
    }
  }
  cudaEventRecord(ctx->ev_a[ctx->ev_count % sfgpu_ctx::EV_RING], ctx->stream);
}
inline void ev_end(sfgpu_ctx* ctx) {
  cudaEventRecord(ctx->ev_b[ctx->ev_count % sfgpu_ctx::EV_RING], ctx->stream);
  ctx->ev_count++;
}

inline int check_committed(sfgpu_ctx* ctx) {
  if (!ctx) return SFGPU_E_INVALID;
  if (!ctx->committed) return fail(ctx, SFGPU_E_STATE, "model not committed");
  return SFGPU_OK;
}

// chunks per replica so that the grid covers the machine a few times over (measured on C2: more, smaller
// CTAs beat fewer fat ones for the generic kernels even though each stages the replica block again)
inline uint32_t chunks_for(const sfgpu_ctx* ctx, uint64_t n_total, uint32_t R, uint32_t threads) {
  uint64_t per_replica = (n_total + R - 1) / std::max<uint32_t>(R, 1);
  uint64_t max_chunks = std::max<uint64_t>(1, (per_replica + threads - 1) / threads);
  uint64_t target_ctas = (uint64_t)ctx->sm_count * 8;
  uint64_t want = std::max<uint64_t>(1, (target_ctas + R - 1) / R);
  uint64_t amort = std::max<uint64_t>(1, per_replica / (threads * 4ull));
  uint64_t chunks = std::min(max_chunks, std::max(want, std::min<uint64_t>(amort, want * 4)));
  return (uint32_t)std::min<uint64_t>(chunks, 65535);
}

inline int ensure_partials(sfgpu_ctx* ctx, size_t need) {
  if (need > ctx->partials_bytes) {
    if (ctx->partials) cudaFree(ctx->partials);
    ctx->partials = nullptr;
    ctx->partials_bytes = 0;
    CU(cudaMalloc(&ctx->partials, need));
    ctx->partials_bytes = need;
  }
  return SFGPU_OK;
}

// Small per-replica arrays of the whole-step calls (seeds, acceptor references in; winner index, score,
// moves_evaluated, winner row out). Host pointers are staged through pinned memory; with SFGPU_DEVICE_IO the
// caller's device pointers are used as they are.
struct SmallIo {
  const uint64_t* d_seeds = nullptr;
  const int64_t* d_ref = nullptr;
  uint32_t* d_idx = nullptr;
  int64_t* d_best = nullptr;
  uint32_t* d_eval = nullptr;
  bool early = false;  // results already on their way back (small_io_results_ready)
  uint32_t* d_win = nullptr;
  size_t o_idx = 0, o_best = 0, o_eval = 0, o_win = 0, total = 0;
  uint32_t win_bytes = 16;
  bool dev_io = false;
};
inline int small_io_begin(sfgpu_ctx* ctx, SmallIo& io, bool dev_io, uint32_t win_bytes, const uint64_t* step_seeds,
                          const int64_t* ref_scores, uint32_t* out_index, int64_t* out_best, uint32_t* out_evaluated,
                          uint32_t* out_winner_rows) {
  const uint32_t R = ctx->dm.R;
  auto a16 = [](size_t v) { return (v + 15) / 16 * 16; };
  const size_t o_seed = 0, o_ref = a16((size_t)R * 8);
  io.o_idx = a16(o_ref + (size_t)R * 32);
  io.o_best = a16(io.o_idx + (size_t)R * 4);
  io.o_eval = a16(io.o_best + (size_t)R * 16);
  io.o_win = a16(io.o_eval + (size_t)R * 4);
  io.total = a16(io.o_win + (size_t)R * win_bytes);
  io.win_bytes = win_bytes;
  io.dev_io = dev_io;
  io.d_seeds = step_seeds;
  io.d_ref = ref_scores;
  io.d_idx = out_index;
  io.d_best = out_best;
  io.d_eval = out_evaluated;
  io.d_win = out_winner_rows;
  if (dev_io) return SFGPU_OK;
  if (io.total > ctx->small_bytes) {
    if (ctx->small_pin) cudaFreeHost(ctx->small_pin);
    if (ctx->small_dev) cudaFree(ctx->small_dev);
    ctx->small_pin = ctx->small_dev = nullptr;
    ctx->small_bytes = 0;
    CU(cudaMallocHost(&ctx->small_pin, io.total));
    CU(cudaMalloc(&ctx->small_dev, io.total));
    CU(cudaMemsetAsync(ctx->small_dev, 0, io.total, ctx->stream));  // the alignment gaps between the arrays travel in the read-back
    ctx->small_bytes = io.total;
  }
  char* pin = (char*)ctx->small_pin;
  char* dv = (char*)ctx->small_dev;
  if (step_seeds) memcpy(pin + o_seed, step_seeds, (size_t)R * 8);
  if (ref_scores) memcpy(pin + o_ref, ref_scores, (size_t)R * 32);
  if (step_seeds || ref_scores) CU(cudaMemcpyAsync(dv, pin, io.o_idx, cudaMemcpyHostToDevice, ctx->stream));
  io.d_seeds = step_seeds ? (const uint64_t*)(dv + o_seed) : nullptr;
  io.d_ref = ref_scores ? (const int64_t*)(dv + o_ref) : nullptr;
  io.d_idx = (uint32_t*)(dv + io.o_idx);
  io.d_best = (int64_t*)(dv + io.o_best);
  io.d_eval = (uint32_t*)(dv + io.o_eval);
  io.d_win = (uint32_t*)(dv + io.o_win);
  return SFGPU_OK;
}
// The winners are final once the finish kernel has run: called between it and the commit kernel, this reads them back on
// a side stream so the copy and the host's wake-up overlap the commit. small_io_end then waits for the copy only; the
// commit stays ordered on the context's stream, ahead of whatever the next call enqueues.
inline int small_io_results_ready(sfgpu_ctx* ctx, SmallIo& io) {
  if (io.dev_io) return SFGPU_OK;
  if (!ctx->io_stream) {
    CU(cudaStreamCreateWithFlags(&ctx->io_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ctx->io_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->io_done, cudaEventDisableTiming));
  }
  char* pin = (char*)ctx->small_pin;
  char* dv = (char*)ctx->small_dev;
  CU(cudaEventRecord(ctx->io_ready, ctx->stream));
  CU(cudaStreamWaitEvent(ctx->io_stream, ctx->io_ready, 0));
  CU(cudaMemcpyAsync(pin + io.o_idx, dv + io.o_idx, io.total - io.o_idx, cudaMemcpyDeviceToHost, ctx->io_stream));
  CU(cudaEventRecord(ctx->io_done, ctx->io_stream));
  io.early = true;
  return SFGPU_OK;
}
inline int small_io_end(sfgpu_ctx* ctx, const SmallIo& io, uint32_t* out_index, int64_t* out_best,
                        uint32_t* out_evaluated, uint32_t* out_winner_rows) {
  if (io.dev_io) return SFGPU_OK;
  const uint32_t R = ctx->dm.R;
  char* pin = (char*)ctx->small_pin;
  char* dv = (char*)ctx->small_dev;
  if (io.early) {
    CU(cudaEventSynchronize(ctx->io_done));
  } else {
    CU(cudaMemcpyAsync(pin + io.o_idx, dv + io.o_idx, io.total - io.o_idx, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  memcpy(out_index, pin + io.o_idx, (size_t)R * 4);
  memcpy(out_best, pin + io.o_best, (size_t)R * 16);
  if (out_evaluated) memcpy(out_evaluated, pin + io.o_eval, (size_t)R * 4);
  if (out_winner_rows) memcpy(out_winner_rows, pin + io.o_win, (size_t)R * io.win_bytes);
  return SFGPU_OK;
}

}  // namespace sfgpu_host

// ---- functions one translation unit provides to the others (each kernel is launched and configured in
// exactly one TU) ---------------------------------------------------------------------------------------------
struct NearbyArgs;
struct ChangeStepArgs;
// sfgpu_api.cu
int sfgpu_launch_apply_list(sfgpu_ctx* ctx, int kind, const uint32_t* d_rows, const uint8_t* d_mask,
                            const uint64_t* d_offsets, const uint32_t* d_index);
int sfgpu_launch_apply_scalar(sfgpu_ctx* ctx, int kind, const uint32_t* d_rows, const uint8_t* d_mask,
                              const uint64_t* d_offsets, const uint32_t* d_index);
struct ForageDev;
int sfgpu_launch_apply_list_kinds(sfgpu_ctx* ctx, const uint32_t* d_rows, const int32_t* d_kinds);
int sfgpu_launch_argbest_counts(sfgpu_ctx* ctx, const ForageDev& f, const uint64_t* d_offs, const uint32_t* d_counts,
                                const uint32_t* d_skip, const int64_t* d_scores, const uint8_t* d_doable,
                                const uint64_t* d_seeds, const int64_t* d_ref, uint32_t* d_idx, int64_t* d_best,
                                uint32_t* d_eval);
// sfgpu_union.cu
inline size_t rank_tables_words_host(uint32_t n_owners) { return 5 * (size_t)n_owners + 2; }
int sfgpu_union_prepare(sfgpu_ctx* ctx, const sfgpu_union_desc* desc, const sfgpu_forage_params* params, UnionPlan& plan);
int sfgpu_union_begin_step(sfgpu_ctx* ctx, UnionPlan& plan);
int sfgpu_union_launch_pass(sfgpu_ctx* ctx, UnionPlan& plan, uint32_t window, bool last_pass, uint32_t* d_idx, int64_t* d_best,
                            uint32_t* d_eval, uint32_t* d_win8, uint32_t* d_flags, uint64_t* d_overflow_acc, bool adaptive,
                            uint32_t win_shift = 0, uint32_t stream_set = 0);
int sfgpu_union_conditional_pass(sfgpu_ctx* ctx, UnionPlan& plan, uint32_t pass_index, uint32_t window, bool last_pass, uint32_t* d_idx,
                                 int64_t* d_best, uint32_t* d_eval, uint64_t* d_overflow_acc, uint32_t win_shift, bool* used);
int sfgpu_union_reset_windows(sfgpu_ctx* ctx, UnionPlan& plan);
// sfgpu_solve.cu
int sfgpu_launch_sa_accept(sfgpu_ctx* ctx, const uint64_t* d_offs, const uint32_t* d_counts, const uint32_t* d_skip,
                           const int64_t* d_scores, uint8_t* d_doable, const int64_t* d_ref, const uint64_t* d_seeds,
                           const SaState* cur, SaState* nxt, const SaParams& p, uint32_t accepted_limit);
// sfgpu_scalar.cu
int sfgpu_configure_scalar(sfgpu_ctx* ctx);
void sfgpu_change_step_chunks(const sfgpu_ctx* ctx, uint32_t* out_per, uint32_t* out_chunks);
int sfgpu_configure_scalar_step(sfgpu_ctx* ctx);  // sfgpu_scalar_step.cu
int sfgpu_configure_scalar_finish(sfgpu_ctx* ctx);  // sfgpu_scalar_finish.cu
int sfgpu_launch_change_finish(sfgpu_ctx* ctx, const ChangeStepArgs& a, uint32_t chunks, uint32_t* d_idx, int64_t* d_best,
                               uint32_t* d_eval, uint32_t* d_win);
int sfgpu_launch_change_step(sfgpu_ctx* ctx, const ChangeStepArgs& a, uint32_t chunks, uint32_t* d_idx, int64_t* d_best,
                             uint32_t* d_eval, uint32_t* d_win);
// sfgpu_list.cu
int sfgpu_configure_list(sfgpu_ctx* ctx);
// sfgpu_nearby.cu
int sfgpu_configure_nearby(sfgpu_ctx* ctx);
int sfgpu_nearby_prepare_cache(sfgpu_ctx* ctx, uint32_t K);
int sfgpu_launch_nearby(sfgpu_ctx* ctx, NearbyArgs& a, uint32_t* d_idx, int64_t* d_best, uint32_t* d_eval, uint32_t* d_win,
                        int move);
