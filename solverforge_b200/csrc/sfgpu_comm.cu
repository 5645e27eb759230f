// Best-score sync across GPUs (replicas only, one tiny collective per sync) for hosts without
// torch: thin wrappers over NCCL resolved at run time (dlopen — libsfgpu has no link-time NCCL dependency and
// loads without it; the sync entry points then fail with SFGPU_E_NCCL).
//
// Reference analogue: up to 16 concurrent seeded jobs of one SolverManager
// (solverforge-solver/src/manager/solver_manager/manager.rs:22,93-146); the reference has no distributed solve.
//
// Every rank contributes {hard, soft, replica} of its best replica; one ncclAllGather of 24 B per rank and a
// local lexicographic max (lowest rank wins ties) — exact for the whole int64 range, unlike a packed MAX key.
#include <dlfcn.h>

#include "sfgpu_ctx.hpp"

using namespace sfgpu_host;

namespace {

struct NcclId {  // ncclUniqueId: 128 opaque bytes, passed by value to ncclCommInitRank
  char bytes[128];
};
typedef int (*nccl_comm_destroy_t)(void*);
typedef int (*nccl_comm_count_t)(void*, int*);
typedef int (*nccl_comm_user_rank_t)(void*, int*);
typedef int (*nccl_all_gather_t)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*nccl_get_error_string_t)(int);

struct NcclApi {
  void* handle = nullptr;
  int (*get_unique_id)(NcclId*) = nullptr;
  int (*comm_init_rank)(void**, int, NcclId, int) = nullptr;
  nccl_comm_destroy_t comm_destroy = nullptr;
  nccl_comm_count_t comm_count = nullptr;
  nccl_comm_user_rank_t comm_user_rank = nullptr;
  nccl_all_gather_t all_gather = nullptr;
  nccl_get_error_string_t get_error_string = nullptr;
  bool tried = false;
};
NcclApi g_nccl;
const int NCCL_INT64 = 4;  // ncclDataType_t::ncclInt64 (nccl.h)

bool nccl_load() {
  if (g_nccl.tried) return g_nccl.handle != nullptr;
  g_nccl.tried = true;
  // the copy the process already loaded (torch bundles its own) wins; then the usual sonames
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return false;
  g_nccl.get_unique_id = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
  g_nccl.comm_init_rank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
  g_nccl.comm_destroy = (nccl_comm_destroy_t)dlsym(h, "ncclCommDestroy");
  g_nccl.comm_count = (nccl_comm_count_t)dlsym(h, "ncclCommCount");
  g_nccl.comm_user_rank = (nccl_comm_user_rank_t)dlsym(h, "ncclCommUserRank");
  g_nccl.all_gather = (nccl_all_gather_t)dlsym(h, "ncclAllGather");
  g_nccl.get_error_string = (nccl_get_error_string_t)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_destroy || !g_nccl.comm_count ||
      !g_nccl.comm_user_rank || !g_nccl.all_gather)
    return false;
  g_nccl.handle = h;
  return true;
}

int nccl_fail(sfgpu_ctx* ctx, const char* what, int code) {
  std::string msg = std::string(what) + ": ";
  msg += g_nccl.get_error_string ? g_nccl.get_error_string(code) : "NCCL error";
  return fail(ctx, SFGPU_E_NCCL, msg);
}

// one CTA: lexicographic max over the replicas' scores (the given [R][2] array, or the committed scores);
// out = {hard, soft, replica} (first replica wins ties)
__global__ void __launch_bounds__(256) rank_best_kernel(const __grid_constant__ DevModel m, const int64_t* __restrict__ scores,
                                                        int64_t* __restrict__ out) {
  __shared__ int64_t s_h[8], s_s[8];
  __shared__ uint32_t s_r[8];
  int64_t bh = INT64_MIN, bs = INT64_MIN;
  uint32_t br = 0xFFFFFFFFu;
  for (uint32_t r = threadIdx.x; r < m.R; r += blockDim.x) {
    const int64_t* cs = scores ? scores + (size_t)r * 2 : (const int64_t*)(m.state + (size_t)r * m.block_bytes + m.off_score);
    const int64_t h = cs[0], s = cs[1];
    if (br == 0xFFFFFFFFu || score_less(bh, bs, h, s)) {  // a thread's replicas come in increasing order
      bh = h;
      bs = s;
      br = r;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t oh = __shfl_down_sync(0xffffffffu, bh, o), os = __shfl_down_sync(0xffffffffu, bs, o);
    const uint32_t orr = __shfl_down_sync(0xffffffffu, br, o);
    if (orr != 0xFFFFFFFFu && (br == 0xFFFFFFFFu || score_less(bh, bs, oh, os) || (oh == bh && os == bs && orr < br))) {
      bh = oh;
      bs = os;
      br = orr;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_h[warp] = bh;
    s_s[warp] = bs;
    s_r[warp] = br;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_r[w] != 0xFFFFFFFFu &&
          (br == 0xFFFFFFFFu || score_less(bh, bs, s_h[w], s_s[w]) || (s_h[w] == bh && s_s[w] == bs && s_r[w] < br))) {
        bh = s_h[w];
        bs = s_s[w];
        br = s_r[w];
      }
    out[0] = bh;
    out[1] = bs;
    out[2] = (int64_t)br;
  }
}

// lexicographic max over the gathered {hard, soft, replica} records (lowest rank on ties), results left on the device
__global__ void gathered_best_kernel(const int64_t* __restrict__ recv, int n_ranks, int64_t* out_best, int32_t* out_rank,
                                     uint32_t* out_replica) {
  if (threadIdx.x != 0) return;
  int best = 0;
  for (int g = 1; g < n_ranks; ++g)
    if (recv[g * 3] != recv[best * 3] ? recv[g * 3] > recv[best * 3] : recv[g * 3 + 1] > recv[best * 3 + 1]) best = g;
  out_best[0] = recv[best * 3];
  out_best[1] = recv[best * 3 + 1];
  if (out_rank) *out_rank = best;
  if (out_replica) *out_replica = (uint32_t)recv[best * 3 + 2];
}

}  // namespace

extern "C" {

int32_t sfgpu_comm_unique_id(uint8_t* out_id128) try {
  if (!out_id128) return SFGPU_E_INVALID;
  if (!nccl_load()) return fail(nullptr, SFGPU_E_NCCL, "NCCL is not available (libnccl.so.2 could not be loaded)");
  NcclId id;
  const int rc = g_nccl.get_unique_id(&id);
  if (rc) return nccl_fail(nullptr, "ncclGetUniqueId", rc);
  memcpy(out_id128, id.bytes, 128);
  return SFGPU_OK;
} catch (...) {
  return SFGPU_E_OOM;
}

int32_t sfgpu_comm_init_rank(int32_t n_ranks, const uint8_t* id128, int32_t rank, int32_t device, void** out_comm) try {
  if (!id128 || !out_comm || n_ranks < 1 || rank < 0 || rank >= n_ranks) return SFGPU_E_INVALID;
  *out_comm = nullptr;
  if (!nccl_load()) return fail(nullptr, SFGPU_E_NCCL, "NCCL is not available (libnccl.so.2 could not be loaded)");
  sfgpu_ctx* ctx = nullptr;
  CU(cudaSetDevice(device));
  NcclId id;
  memcpy(id.bytes, id128, 128);
  void* comm = nullptr;
  const int rc = g_nccl.comm_init_rank(&comm, n_ranks, id, rank);
  if (rc) return nccl_fail(nullptr, "ncclCommInitRank", rc);
  *out_comm = comm;
  return SFGPU_OK;
} catch (...) {
  return SFGPU_E_OOM;
}

int32_t sfgpu_comm_destroy(void* comm) try {
  if (!comm) return SFGPU_E_INVALID;
  if (!nccl_load()) return fail(nullptr, SFGPU_E_NCCL, "NCCL is not available");
  const int rc = g_nccl.comm_destroy(comm);
  if (rc) return nccl_fail(nullptr, "ncclCommDestroy", rc);
  return SFGPU_OK;
} catch (...) {
  return SFGPU_E_OOM;
}

int32_t sfgpu_sync_best(sfgpu_ctx* ctx, void* nccl_comm, uint32_t flags, const int64_t* scores, int64_t* out_best,
                        int32_t* out_owner_rank, uint32_t* out_owner_replica) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (!out_best) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  if (scores && !(flags & SFGPU_DEVICE_IO)) return fail(ctx, SFGPU_E_INVALID, "scores must be a device pointer (SFGPU_DEVICE_IO) or NULL");
  CU(cudaSetDevice(ctx->device));
  int n_ranks = 1, my_rank = 0;
  if (nccl_comm) {
    if (!nccl_load()) return fail(ctx, SFGPU_E_NCCL, "NCCL is not available (libnccl.so.2 could not be loaded)");
    rc = g_nccl.comm_count(nccl_comm, &n_ranks);
    if (rc) return nccl_fail(ctx, "ncclCommCount", rc);
    rc = g_nccl.comm_user_rank(nccl_comm, &my_rank);
    if (rc) return nccl_fail(ctx, "ncclCommUserRank", rc);
  }
  const size_t need = (size_t)(n_ranks + 1) * 24;
  if (need > ctx->sync_bytes) {
    if (ctx->sync_dev) cudaFree(ctx->sync_dev);
    if (ctx->sync_pin) cudaFreeHost(ctx->sync_pin);
    ctx->sync_dev = ctx->sync_pin = nullptr;
    ctx->sync_bytes = 0;
    CU(cudaMalloc(&ctx->sync_dev, need));
    CU(cudaMallocHost(&ctx->sync_pin, need));
    ctx->sync_bytes = need;
  }
  int64_t* send = (int64_t*)ctx->sync_dev;
  int64_t* recv = send + 3;
  rank_best_kernel<<<1, 256, 0, ctx->stream>>>(ctx->dm, scores, send);
  ctx->launches++;
  CU(cudaGetLastError());
  if (nccl_comm) {
    rc = g_nccl.all_gather(send, recv, 3, NCCL_INT64, nccl_comm, ctx->stream);
    if (rc) return nccl_fail(ctx, "ncclAllGather", rc);
  } else {
    CU(cudaMemcpyAsync(recv, send, 24, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (flags & SFGPU_SYNC_ASYNC) {
    // outputs stay on the device and the stream is not drained: the next step's kernels queue right behind
    gathered_best_kernel<<<1, 32, 0, ctx->stream>>>(recv, n_ranks, out_best, out_owner_rank, out_owner_replica);
    ctx->launches++;
    CU(cudaGetLastError());
    return SFGPU_OK;
  }
  int64_t* host = (int64_t*)ctx->sync_pin;
  CU(cudaMemcpyAsync(host, recv, (size_t)n_ranks * 24, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  int best = 0;
  for (int g = 1; g < n_ranks; ++g)
    if (host[g * 3] != host[best * 3] ? host[g * 3] > host[best * 3] : host[g * 3 + 1] > host[best * 3 + 1]) best = g;
  out_best[0] = host[best * 3];
  out_best[1] = host[best * 3 + 1];
  if (out_owner_rank) *out_owner_rank = best;
  if (out_owner_replica) *out_owner_replica = (uint32_t)host[best * 3 + 2];
  (void)my_rank;
  return SFGPU_OK;
} catch (const std::bad_alloc&) {
  return SFGPU_E_OOM;
} catch (...) {
  return SFGPU_E_INVALID;
}

}  // extern "C"
