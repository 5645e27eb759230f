// libsfgpu — host side of the C ABI declared in include/sfgpu.h: context, model lowering, HBM layout, the
// generic entry points (score, argbest, apply, state read-back). The kernel families live in their own
// translation units (sfgpu_scalar.cu, sfgpu_list.cu, sfgpu_nearby.cu, sfgpu_index.cu, sfgpu_solve.cu) so the
// build parallelises. No CPU scoring path exists anywhere in the library: every compute entry point launches
// a CUDA kernel or fails with SFGPU_E_CUDA.
#include "sfgpu_ctx.hpp"
#include "sfgpu_core_kernels.cuh"

using namespace sfgpu_host;

extern "C" {

int32_t sfgpu_abi_version(void) { return SFGPU_ABI_VERSION; }

const char* sfgpu_last_error(const sfgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_noctx_err.c_str(); }

int32_t sfgpu_ctx_create(int32_t device, uint64_t flags, void* cuda_stream, sfgpu_ctx** out) {
  if (!out) return SFGPU_E_INVALID;
  *out = nullptr;
  sfgpu_ctx* ctx = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, SFGPU_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                           " (libsfgpu has no CPU fallback)");
  if (device < 0 || device >= n) return fail(nullptr, SFGPU_E_INVALID, "device index out of range");
  CU(cudaSetDevice(device));
  ctx = new sfgpu_ctx();
  ctx->device = device;
  ctx->force_generic = (flags & SFGPU_CTX_GENERIC_KERNELS) != 0;
  if (cuda_stream || (flags & SFGPU_CTX_LEGACY_DEFAULT_STREAM)) {
    ctx->stream = (cudaStream_t)cuda_stream;  // NULL + flag = the legacy default stream
  } else {
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete ctx;
      return fail(nullptr, SFGPU_E_CUDA, cudaGetErrorString(e));
    }
    ctx->own_stream = true;
  }
  cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  *out = ctx;
  return SFGPU_OK;
}

int32_t sfgpu_ctx_destroy(sfgpu_ctx* ctx) try {
  if (!ctx) return SFGPU_E_INVALID;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (void* p : ctx->dev_allocs) cudaFree(p);
  if (ctx->pin) cudaFreeHost(ctx->pin);
  if (ctx->dscr) cudaFree(ctx->dscr);
  if (ctx->partials) cudaFree(ctx->partials);
  if (ctx->nbc_buf) cudaFree(ctx->nbc_buf);
  if (ctx->solve_buf) cudaFree(ctx->solve_buf);
  if (ctx->union_buf) cudaFree(ctx->union_buf);
  for (auto s_ : ctx->aux_streams) cudaStreamDestroy(s_);
  for (auto e_ : ctx->aux_events) cudaEventDestroy(e_);
  if (ctx->io_stream) cudaStreamDestroy(ctx->io_stream);
  if (ctx->io_ready) cudaEventDestroy(ctx->io_ready);
  if (ctx->io_done) cudaEventDestroy(ctx->io_done);
  if (ctx->small_pin) cudaFreeHost(ctx->small_pin);
  if (ctx->small_dev) cudaFree(ctx->small_dev);
  if (ctx->sync_dev) cudaFree(ctx->sync_dev);
  if (ctx->sync_pin) cudaFreeHost(ctx->sync_pin);
  for (auto e : ctx->ev_a) cudaEventDestroy(e);
  for (auto e : ctx->ev_b) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_synchronize(sfgpu_ctx* ctx) try {
  if (!ctx) return SFGPU_E_INVALID;
  CU(cudaStreamSynchronize(ctx->stream));
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_model_begin(sfgpu_ctx* ctx, uint32_t n_replicas) try {
  if (!ctx) return SFGPU_E_INVALID;
  if (ctx->committed || ctx->building) return fail(ctx, SFGPU_E_STATE, "model already begun");
  if (n_replicas == 0 || n_replicas > 65535) return fail(ctx, SFGPU_E_INVALID, "n_replicas must be in [1, 65535]");
  ctx->building = true;
  ctx->R = n_replicas;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_add_collection(sfgpu_ctx* ctx, const char* name, uint32_t n_rows, int32_t descriptor_index,
                             uint32_t* out_collection) try {
  if (!ctx || !out_collection) return SFGPU_E_INVALID;
  if (!ctx->building) return fail(ctx, SFGPU_E_STATE, "sfgpu_model_begin first");
  ctx->colls.push_back({name ? name : "", n_rows, descriptor_index});
  *out_collection = (uint32_t)ctx->colls.size() - 1;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_add_column_i64(sfgpu_ctx* ctx, uint32_t collection, const char* name, const int64_t* values,
                             uint32_t* out_column) try {
  (void)name;
  if (!ctx || !values || !out_column) return SFGPU_E_INVALID;
  if (!ctx->building) return fail(ctx, SFGPU_E_STATE, "sfgpu_model_begin first");
  if (collection >= ctx->colls.size()) return fail(ctx, SFGPU_E_INVALID, "unknown collection");
  Column c;
  c.coll = collection;
  c.host.assign(values, values + ctx->colls[collection].n_rows);
  ctx->cols.push_back(std::move(c));
  *out_column = (uint32_t)ctx->cols.size() - 1;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_add_scalar_variable(sfgpu_ctx* ctx, uint32_t collection, const char* name, uint32_t n_values,
                                  int32_t allows_unassigned, uint32_t* out_variable) try {
  (void)name;
  if (!ctx || !out_variable) return SFGPU_E_INVALID;
  if (!ctx->building) return fail(ctx, SFGPU_E_STATE, "sfgpu_model_begin first");
  if (collection >= ctx->colls.size()) return fail(ctx, SFGPU_E_INVALID, "unknown collection");
  if (!ctx->svars.empty()) return fail(ctx, SFGPU_E_UNSUPPORTED, "one scalar planning variable per model");
  if (n_values == 0 || n_values > 0x7FFFFFFFu) return fail(ctx, SFGPU_E_INVALID, "n_values out of range");
  ctx->svars.push_back({collection, n_values, allows_unassigned, {}, false});
  *out_variable = 0;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_add_list_variable(sfgpu_ctx* ctx, uint32_t owner_collection, uint32_t element_collection,
                                const char* name, uint32_t* out_variable) try {
  (void)name;
  if (!ctx || !out_variable) return SFGPU_E_INVALID;
  if (!ctx->building) return fail(ctx, SFGPU_E_STATE, "sfgpu_model_begin first");
  if (owner_collection >= ctx->colls.size() || element_collection >= ctx->colls.size())
    return fail(ctx, SFGPU_E_INVALID, "unknown collection");
  if (!ctx->lvars.empty()) return fail(ctx, SFGPU_E_UNSUPPORTED, "one list planning variable per model");
  ctx->lvars.push_back({owner_collection, element_collection, {}, {}, false});
  *out_variable = 0x80000000u;  // list variables live in their own id space
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_add_csr(sfgpu_ctx* ctx, const char* name, uint32_t n_rows, const uint32_t* row_ptr,
                      const uint32_t* col_idx, uint32_t* out_csr) try {
  (void)name;
  if (!ctx || !row_ptr || !out_csr) return SFGPU_E_INVALID;
  if (!ctx->building) return fail(ctx, SFGPU_E_STATE, "sfgpu_model_begin first");
  Csr c;
  c.n_rows = n_rows;
  c.row_ptr.assign(row_ptr, row_ptr + n_rows + 1);
  uint32_t nnz = row_ptr[n_rows];
  if (nnz && !col_idx) return SFGPU_E_INVALID;
  c.col.assign(col_idx, col_idx + nnz);
  ctx->csrs.push_back(std::move(c));
  *out_csr = (uint32_t)ctx->csrs.size() - 1;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_add_matrix_i64(sfgpu_ctx* ctx, const char* name, uint32_t rows, uint32_t cols,
                             const int64_t* values, int32_t cost_semantics, uint32_t* out_matrix) try {
  (void)name;
  if (!ctx || !values || !out_matrix) return SFGPU_E_INVALID;
  if (!ctx->building) return fail(ctx, SFGPU_E_STATE, "sfgpu_model_begin first");
  Matrix m;
  m.rows = rows;
  m.cols = cols;
  m.host.assign(values, values + (size_t)rows * cols);
  if (cost_semantics == 1) {
    // ProblemData::distance_cost (problem_data.rs:28-47)
    const int64_t UNREACH = std::numeric_limits<int64_t>::max();
    for (auto& v : m.host)
      if (!(v >= 0 && v != UNREACH)) v = UNREACH / 4;
  }
  ctx->mats.push_back(std::move(m));
  *out_matrix = (uint32_t)ctx->mats.size() - 1;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_add_expr(sfgpu_ctx* ctx, const sfgpu_expr_op* ops, uint32_t n_ops, uint32_t* out_expr) try {
  if (!ctx || !ops || !out_expr) return SFGPU_E_INVALID;
  if (!ctx->building) return fail(ctx, SFGPU_E_STATE, "sfgpu_model_begin first");
  if (n_ops == 0 || n_ops > 4096) return fail(ctx, SFGPU_E_INVALID, "expression length must be in [1, 4096]");
  int depth = 0;
  for (uint32_t i = 0; i < n_ops; ++i) {
    int pops = 0, pushes = 1;
    switch (ops[i].op) {
      case SFGPU_X_CONST: case SFGPU_X_A_IDX: case SFGPU_X_B_IDX: case SFGPU_X_VALUE: case SFGPU_X_A_VAL: case SFGPU_X_B_VAL: break;
      case SFGPU_X_A_COL: case SFGPU_X_B_COL:
        if (ops[i].arg >= ctx->cols.size()) return fail(ctx, SFGPU_E_INVALID, "expression reads an unknown column");
        break;
      case SFGPU_X_NEG: case SFGPU_X_ABS: case SFGPU_X_NOT: pops = 1; break;
      case SFGPU_X_ADD: case SFGPU_X_SUB: case SFGPU_X_MUL: case SFGPU_X_MIN: case SFGPU_X_MAX: case SFGPU_X_MOD:
      case SFGPU_X_EQ: case SFGPU_X_NE: case SFGPU_X_LT: case SFGPU_X_LE: case SFGPU_X_GT: case SFGPU_X_GE:
      case SFGPU_X_AND: case SFGPU_X_OR: pops = 2; break;
      case SFGPU_X_CSR_CONTAINS:
        if (ops[i].arg >= ctx->csrs.size()) return fail(ctx, SFGPU_E_INVALID, "expression reads an unknown csr");
        pops = 2;
        break;
      case SFGPU_X_SELECT: pops = 3; break;
      default: return fail(ctx, SFGPU_E_UNSUPPORTED, "unknown expression op (not expressible on device)");
    }
    if (depth < pops) return fail(ctx, SFGPU_E_INVALID, "expression stack underflow");
    depth += pushes - pops;
    if (depth > 8) return fail(ctx, SFGPU_E_UNSUPPORTED, "expression stack deeper than 8");
  }
  if (depth != 1) return fail(ctx, SFGPU_E_INVALID, "an expression leaves exactly one value");
  ExprHost e;
  e.ops.assign(ops, ops + n_ops);
  ctx->exprs.push_back(std::move(e));
  *out_expr = (uint32_t)ctx->exprs.size() - 1;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_add_constraint(sfgpu_ctx* ctx, const sfgpu_constraint_desc* desc, uint32_t* out_constraint) try {
  if (!ctx || !desc) return SFGPU_E_INVALID;
  if (!ctx->building) return fail(ctx, SFGPU_E_STATE, "sfgpu_model_begin first");
  if (ctx->cons.size() >= SFGPU_MAX_CONS) return fail(ctx, SFGPU_E_UNSUPPORTED, "too many constraints");
  if (desc->kind < SFGPU_K_UNI || desc->kind > SFGPU_K_PAIR_KEY_EXPR)
    return fail(ctx, SFGPU_E_UNSUPPORTED, "unknown constraint kind (not expressible on device)");
  if (desc->weight.fn < SFGPU_W_CONST || desc->weight.fn > SFGPU_W_PAIRS || desc->weight.level < 0 ||
      desc->weight.level > 1)
    return fail(ctx, SFGPU_E_INVALID, "bad weight");
  bool needs_const = desc->kind == SFGPU_K_PAIR_CSR_EQUAL || desc->kind == SFGPU_K_PAIR_KEY_EQUAL ||
                     desc->kind == SFGPU_K_EXISTS_FLAT;
  if (needs_const && desc->weight.fn != SFGPU_W_CONST)
    return fail(ctx, SFGPU_E_UNSUPPORTED, "pair/exists constraints take a constant weight");
  ConsHost c;
  c.d = *desc;
  c.name = desc->name ? desc->name : "";
  c.d.name = nullptr;
  ctx->cons.push_back(std::move(c));
  if (out_constraint) *out_constraint = (uint32_t)ctx->cons.size() - 1;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_set_scalar_state(sfgpu_ctx* ctx, uint32_t variable, const int32_t* values, int32_t per_replica) try {
  if (!ctx || !values) return SFGPU_E_INVALID;
  if (!ctx->building) return fail(ctx, SFGPU_E_STATE, "state is uploaded before sfgpu_model_commit");
  if (variable != 0 || ctx->svars.empty()) return fail(ctx, SFGPU_E_INVALID, "unknown scalar variable");
  ScalarVar& v = ctx->svars[0];
  size_t n = ctx->colls[v.coll].n_rows;
  size_t total = per_replica ? n * ctx->R : n;
  for (size_t i = 0; i < total; ++i)
    if (values[i] >= (int32_t)v.n_values) return fail(ctx, SFGPU_E_INVALID, "scalar value out of range");
  v.init.assign(values, values + total);
  for (auto& x : v.init)
    if (x < 0) x = SFGPU_NONE;
  v.per_replica = per_replica != 0;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_set_list_state(sfgpu_ctx* ctx, uint32_t variable, const uint32_t* offsets, const uint32_t* elems,
                             int32_t per_replica) try {
  if (!ctx || !offsets) return SFGPU_E_INVALID;
  if (!ctx->building) return fail(ctx, SFGPU_E_STATE, "state is uploaded before sfgpu_model_commit");
  if (variable != 0x80000000u || ctx->lvars.empty()) return fail(ctx, SFGPU_E_INVALID, "unknown list variable");
  ListVar& v = ctx->lvars[0];
  uint32_t no = ctx->colls[v.owner_coll].n_rows, ne = ctx->colls[v.elem_coll].n_rows;
  uint32_t copies = per_replica ? ctx->R : 1;
  v.offsets.assign(offsets, offsets + (size_t)copies * (no + 1));
  size_t total = 0;
  for (uint32_t c = 0; c < copies; ++c) {
    const uint32_t* o = offsets + (size_t)c * (no + 1);
    if (o[0] != 0) return fail(ctx, SFGPU_E_INVALID, "list offsets must start at 0");
    for (uint32_t i = 0; i < no; ++i)
      if (o[i + 1] < o[i]) return fail(ctx, SFGPU_E_INVALID, "list offsets must be non-decreasing");
    total += o[no];
  }
  if (total && !elems) return SFGPU_E_INVALID;
  v.elems.assign(elems, elems + total);
  for (uint32_t e : v.elems)
    if (e >= ne) return fail(ctx, SFGPU_E_INVALID, "list element out of range");
  {
    // an element sits in at most one list of a replica (a list variable partitions its elements): duplicates
    // would silently corrupt the per-element inverse index and the flattened exists counts
    std::vector<uint8_t> seen(ne);
    size_t at = 0;
    for (uint32_t c = 0; c < copies; ++c) {
      std::fill(seen.begin(), seen.end(), 0);
      const uint32_t cnt = offsets[(size_t)c * (no + 1) + no];
      for (uint32_t i = 0; i < cnt; ++i) {
        const uint32_t e = v.elems[at + i];
        if (seen[e]) return fail(ctx, SFGPU_E_INVALID, "list element appears twice in one replica");
        seen[e] = 1;
      }
      at += cnt;
    }
  }
  v.per_replica = per_replica != 0;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

// device tables of the column expressions: pointers of every column and every CSR, uploaded once per model
static int expr_tables(sfgpu_ctx* ctx) {
  if (ctx->expr_cols_dev) return SFGPU_OK;
  std::vector<const int64_t*> cols;
  for (auto& c : ctx->cols) cols.push_back(c.dev);
  if (cols.empty()) cols.push_back(nullptr);
  std::vector<const uint32_t*> csrs;
  for (auto& g : ctx->csrs) {
    uint32_t *rp = nullptr, *ci = nullptr;
    int rc = dev_upload(ctx, g.row_ptr.data(), g.row_ptr.size(), &rp);
    if (rc) return rc;
    rc = dev_upload(ctx, g.col.data(), g.col.size(), &ci);
    if (rc) return rc;
    csrs.push_back(rp);
    csrs.push_back(ci);
  }
  if (csrs.empty()) csrs.push_back(nullptr);
  const int64_t** dc = nullptr;
  const uint32_t** dg = nullptr;
  int rc = dev_upload(ctx, cols.data(), cols.size(), &dc);
  if (rc) return rc;
  rc = dev_upload(ctx, csrs.data(), csrs.size(), &dg);
  if (rc) return rc;
  ctx->expr_cols_dev = dc;
  ctx->expr_csrs_dev = dg;
  return SFGPU_OK;
}

int32_t sfgpu_model_commit(sfgpu_ctx* ctx, int64_t* out_scores) try {
  if (!ctx) return SFGPU_E_INVALID;
  if (!ctx->building) return fail(ctx, SFGPU_E_STATE, "sfgpu_model_begin first");
  CU(cudaSetDevice(ctx->device));
  DevModel& dm = ctx->dm;
  memset(&dm, 0, sizeof(dm));
  dm.R = ctx->R;
  // upload shared static data
  for (auto& c : ctx->cols) {
    int rc = dev_upload(ctx, c.host.data(), c.host.size(), &c.dev);
    if (rc) return rc;
    bool fits = true;
    for (int64_t v : c.host)
      if (v <= -(1ll << 30) || v >= (1ll << 30)) fits = false;
    if (fits) {
      std::vector<int32_t> narrow(c.host.begin(), c.host.end());
      rc = dev_upload(ctx, narrow.data(), narrow.size(), &c.dev32);
      if (rc) return rc;
    }
  }
  for (auto& m : ctx->mats) {
    bool fits = true;
    for (int64_t v : m.host)
      if (v < 0 || v > 0x7FFFFFFF) {
        fits = false;
        break;
      }
    m.i32 = fits;
    if (fits) {
      std::vector<int32_t> narrow(m.host.begin(), m.host.end());
      int32_t* p = nullptr;
      int rc = dev_upload(ctx, narrow.data(), narrow.size(), &p);
      if (rc) return rc;
      m.dev = p;
    } else {
      int64_t* p = nullptr;
      int rc = dev_upload(ctx, m.host.data(), m.host.size(), &p);
      if (rc) return rc;
      m.dev = p;
    }
  }
  // block layout: [score][scalar var][list offsets][list elems][staged retained][unstaged retained]
  uint32_t off = 0;
  dm.off_score = off;
  off += 16;
  if (!ctx->svars.empty()) {
    ScalarVar& v = ctx->svars[0];
    dm.has_scalar = 1;
    dm.n_entities = ctx->colls[v.coll].n_rows;
    dm.n_values = v.n_values;
    dm.allows_unassigned = v.allows_unassigned;
    dm.off_var = off;
    off = align_up(off + dm.n_entities * 4, 16);
    if (v.init.empty()) v.init.assign(dm.n_entities, SFGPU_NONE);
  }
  dm.fast_pc = dm.fast_ls = -1;
  if (!ctx->lvars.empty()) {
    ListVar& v = ctx->lvars[0];
    dm.has_list = 1;
    dm.n_owners = ctx->colls[v.owner_coll].n_rows;
    dm.n_elem_rows = ctx->colls[v.elem_coll].n_rows;
    if (v.offsets.empty()) v.offsets.assign(dm.n_owners + 1, 0);
    uint32_t copies = v.per_replica ? ctx->R : 1;
    uint32_t cap = 0;
    for (uint32_t c = 0; c < copies; ++c) cap = std::max(cap, v.offsets[(size_t)c * (dm.n_owners + 1) + dm.n_owners]);
    dm.elem_cap = std::max(cap, 1u);  // relocations and swaps keep the element count
    // fast list path eligibility (see score_list_change_fast_kernel)
    bool fast = true;
    for (uint32_t k = 0; k < ctx->cons.size() && fast; ++k) {
      const sfgpu_constraint_desc& d = ctx->cons[k].d;
      if (d.kind == SFGPU_K_LIST_PATH_COST) {
        if (dm.fast_pc >= 0 || d.weight.fn != SFGPU_W_LINEAR || d.aux0 >= ctx->mats.size()) { fast = false; break; }
        const Matrix& mt = ctx->mats[d.aux0];
        int64_t mx = 0;
        for (int64_t c : mt.host) mx = std::max(mx, c);
        if (!mt.i32 || mx >= (1 << 28) || (uint64_t)mt.rows * mt.cols >= (1ull << 31)) { fast = false; break; }
        dm.fast_pc = (int32_t)k;
      } else if (d.kind == SFGPU_K_LIST_SUM) {
        if (dm.fast_ls >= 0 || d.aux0 >= ctx->cols.size() || d.weight.fn == SFGPU_W_ABSDIFF || d.weight.fn == SFGPU_W_PAIRS) { fast = false; break; }
        for (int64_t c : ctx->cols[d.aux0].host)
          if (c < -(1ll << 30) || c > (1ll << 30)) fast = false;
        dm.fast_ls = (int32_t)k;
      }
    }
    if (fast && (dm.fast_pc >= 0 || dm.fast_ls >= 0)) {
      dm.fast_list = 1;
      // narrow model: every fast-path score delta provably fits int32 (lets the fused kernels
      // reduce deltas with 32-bit REDUX instead of 64-bit shuffles)
      bool narrow = true;
      double bound = 0;
      int64_t cell_max = 0, val_abs_max = 0;
      bool square = true;
      if (dm.fast_pc >= 0) {
        const sfgpu_constraint_desc& d = ctx->cons[dm.fast_pc].d;
        int64_t mx = 0;
        for (int64_t c : ctx->mats[d.aux0].host) mx = std::max(mx, c);
        cell_max = mx;
        square = ctx->mats[d.aux0].rows == ctx->mats[d.aux0].cols;
        bound += std::fabs((double)d.weight.a) * 4.0 * (double)mx;
      }
      if (dm.fast_ls >= 0) {
        const sfgpu_constraint_desc& d = ctx->cons[dm.fast_ls].d;
        double tot = 0;
        for (int64_t c : ctx->cols[d.aux0].host) tot += std::fabs((double)c);
        for (int64_t c : ctx->cols[d.aux0].host) val_abs_max = std::max<int64_t>(val_abs_max, c < 0 ? -c : c);
        if (d.weight.fn == SFGPU_W_EXCESS) bound += std::fabs((double)d.weight.a) * 2.0 * tot;
        else if (d.weight.fn == SFGPU_W_SQUARE) bound += std::fabs((double)d.weight.a) * 4.0 * tot * tot;
      }
      if (bound >= 2147483647.0) narrow = false;
      dm.fast_narrow = narrow ? 1 : 0;
      dm.off_route_rec = off;
      off += dm.n_owners * 16;
      dm.off_pos_rec = off;
      off += dm.elem_cap * 16;
      dm.off_slot_rec = off;
      off += (dm.elem_cap + dm.n_owners) * 16;
      // device-side nearby neighbourhood: needs the path-cost matrix as the distance meter, every
      // cell finite (guaranteed by the < 2^28 test above) and 16-bit owner / position fields
      if (dm.fast_pc >= 0 && dm.n_owners >= 1 && dm.n_owners < 65536 && dm.elem_cap < 65536) {
        dm.fast_score_bytes = off;  // the score kernels do not need pos_of
        dm.nearby_ok = 1;
        dm.off_pos_of = off;
        off = align_up(off + dm.n_elem_rows * 4, 16);
      }
      dm.fast_stage_bytes = off;
      if (!dm.nearby_ok) dm.fast_score_bytes = off;
      // compact 8-byte record copies for the score kernel (the uint16-cell / 32-bit arithmetic path
      // with 16-bit ids, positions and values); placed behind the staged range of the nearby kernels
      if (narrow && dm.fast_pc >= 0 && square && cell_max < 65536 && dm.n_elem_rows < 65536 &&
          dm.elem_cap + dm.n_owners < 65536 && val_abs_max < 32768 && !getenv("SFGPU_NO_COMPACT")) {
        dm.off_route8 = off;
        off += dm.n_owners * 8;
        dm.off_pos8 = off;
        off += dm.elem_cap * 8;
        dm.off_slot8 = off;
        off += (dm.elem_cap + dm.n_owners) * 8;
        off = align_up(off, 16);
        dm.compact_bytes = off - dm.off_route8;
      }
    } else {
      dm.fast_pc = dm.fast_ls = -1;
    }
    dm.off_offsets = off;
    off = align_up(off + (dm.n_owners + 1) * 4, 16);
    dm.off_elems = off;
    off = align_up(off + dm.elem_cap * 4, 16);
  }
  // constraints: staged sections first
  dm.n_cons = (uint32_t)ctx->cons.size();
  std::vector<uint32_t> unstaged_bytes(dm.n_cons, 0);
  for (uint32_t k = 0; k < dm.n_cons; ++k) {
    const sfgpu_constraint_desc& d = ctx->cons[k].d;
    ConsDev& c = dm.cons[k];
    c.kind = d.kind;
    c.sign = d.impact == SFGPU_REWARD ? 1 : -1;
    c.w = WeightDev{d.weight.fn, d.weight.level, d.weight.a, d.weight.b};
    auto column = [&](uint32_t id, uint32_t want_rows, const int64_t** out) -> int {
      *out = nullptr;
      if (id == 0xFFFFFFFFu) return SFGPU_OK;
      if (id >= ctx->cols.size()) return fail(ctx, SFGPU_E_INVALID, "unknown column in constraint " + ctx->cons[k].name);
      if (ctx->colls[ctx->cols[id].coll].n_rows != want_rows)
        return fail(ctx, SFGPU_E_INVALID, "column row count mismatch in constraint " + ctx->cons[k].name);
      *out = ctx->cols[id].dev;
      return SFGPU_OK;
    };
    bool scalar_kind = d.kind == SFGPU_K_UNI || d.kind == SFGPU_K_PAIR_CSR_EQUAL || d.kind == SFGPU_K_PAIR_KEY_EQUAL ||
                       d.kind == SFGPU_K_GROUP || d.kind == SFGPU_K_LOAD_BALANCE || d.kind == SFGPU_K_PROJECT_GROUP ||
                       d.kind == SFGPU_K_RUNS || d.kind == SFGPU_K_JOIN_EXPR || d.kind == SFGPU_K_PAIR_KEY_EXPR;
    if (scalar_kind && !dm.has_scalar) return fail(ctx, SFGPU_E_INVALID, "constraint needs a scalar variable");
    if (!scalar_kind && !dm.has_list) return fail(ctx, SFGPU_E_INVALID, "constraint needs a list variable");
    switch (d.kind) {
      case SFGPU_K_PAIR_KEY_EXPR: {
        int rc = expr_tables(ctx);
        if (rc) return rc;
        if (d.collection != ctx->svars[0].coll) return fail(ctx, SFGPU_E_INVALID, "PAIR_KEY_EXPR joins the entity collection with itself");
        if (d.p1 <= 0 || d.p1 > (1 << 24)) return fail(ctx, SFGPU_E_UNSUPPORTED, "PAIR_KEY_EXPR: p1 (number of keys) must be in [1, 2^24]");
        const uint32_t ids[4] = {(uint32_t)((uint64_t)d.p0 & 0xFFFFFFFFu), (uint32_t)((uint64_t)d.p0 >> 32), d.aux0, d.aux1};
        if (ids[0] == 0xFFFFFFFFu) return fail(ctx, SFGPU_E_INVALID, "PAIR_KEY_EXPR needs a left key expression");
        std::vector<sfgpu_expr_op> prog;
        uint32_t lens[4] = {0, 0, 0, 0};
        for (int w = 0; w < 4; ++w) {
          if (ids[w] == 0xFFFFFFFFu) continue;
          if (ids[w] >= ctx->exprs.size()) return fail(ctx, SFGPU_E_INVALID, "PAIR_KEY_EXPR: unknown expression");
          for (sfgpu_expr_op o : ctx->exprs[ids[w]].ops) {
            if ((o.op == SFGPU_X_A_COL || o.op == SFGPU_X_B_COL) && ctx->cols[o.arg].coll != d.collection)
              return fail(ctx, SFGPU_E_INVALID, "PAIR_KEY_EXPR: expressions read columns of the entity collection");
            if (w < 2 && (o.op == SFGPU_X_B_COL || o.op == SFGPU_X_B_IDX || o.op == SFGPU_X_B_VAL))
              return fail(ctx, SFGPU_E_INVALID, "PAIR_KEY_EXPR: a key expression reads one row (A_* operands only)");
            if (o.op == SFGPU_X_CSR_CONTAINS) o.imm = ctx->csrs[o.arg].n_rows;
            prog.push_back(o);
          }
          lens[w] = (uint32_t)ctx->exprs[ids[w]].ops.size();
          if (lens[w] > 0xFFFF) return fail(ctx, SFGPU_E_UNSUPPORTED, "expression too long");
        }
        sfgpu_expr_op* dprog = nullptr;
        rc = dev_upload(ctx, prog.data(), prog.size(), &dprog);
        if (rc) return rc;
        c.g0 = dprog;
        c.g1 = ctx->expr_cols_dev;
        c.g2 = ctx->expr_csrs_dev;
        c.pad = lens[0] | (lens[1] << 16);
        c.n0 = lens[2] | (lens[3] << 16);
        c.p1 = d.p1;
        unstaged_bytes[k] = align_up((uint32_t)(2 * d.p1 + 4 * (int64_t)dm.n_entities) * 4, 16);
        break;
      }
      case SFGPU_K_JOIN_EXPR: {
        int rc = expr_tables(ctx);
        if (rc) return rc;
        if (d.collection != ctx->svars[0].coll) return fail(ctx, SFGPU_E_INVALID, "JOIN_EXPR: the A side is the entity collection");
        if (d.p1 < 0 || (size_t)d.p1 >= ctx->colls.size()) return fail(ctx, SFGPU_E_INVALID, "JOIN_EXPR: unknown B collection");
        const uint32_t a_coll = d.collection, b_coll = (uint32_t)d.p1, b_rows = ctx->colls[b_coll].n_rows;
        if (d.p0 >= 0) {
          if ((size_t)d.p0 >= ctx->csrs.size()) return fail(ctx, SFGPU_E_INVALID, "JOIN_EXPR: unknown bucket csr");
          const Csr& g = ctx->csrs[d.p0];
          if (g.n_rows < dm.n_values) return fail(ctx, SFGPU_E_INVALID, "JOIN_EXPR: the bucket csr needs a row per key value");
          for (uint32_t v : g.col)
            if (v >= b_rows) return fail(ctx, SFGPU_E_INVALID, "JOIN_EXPR: bucket csr lists a B row out of range");
        } else if (b_rows < dm.n_values) {
          return fail(ctx, SFGPU_E_INVALID, "JOIN_EXPR: the variable's values must be rows of B");
        }
        std::vector<sfgpu_expr_op> prog;
        uint32_t lens[2] = {0, 0};
        const uint32_t ids[2] = {d.aux0, d.aux1};
        for (int w = 0; w < 2; ++w) {
          if (ids[w] == 0xFFFFFFFFu) continue;
          if (ids[w] >= ctx->exprs.size()) return fail(ctx, SFGPU_E_INVALID, "JOIN_EXPR: unknown expression");
          for (sfgpu_expr_op o : ctx->exprs[ids[w]].ops) {
            if (o.op == SFGPU_X_A_COL && ctx->cols[o.arg].coll != a_coll)
              return fail(ctx, SFGPU_E_INVALID, "JOIN_EXPR: A_COL reads a column of another collection");
            if (o.op == SFGPU_X_B_COL && ctx->cols[o.arg].coll != b_coll)
              return fail(ctx, SFGPU_E_INVALID, "JOIN_EXPR: B_COL reads a column of another collection");
            if (o.op == SFGPU_X_CSR_CONTAINS) o.imm = ctx->csrs[o.arg].n_rows;
            prog.push_back(o);
          }
          lens[w] = (uint32_t)ctx->exprs[ids[w]].ops.size();
        }
        if (lens[0] > 0xFFFF || lens[1] > 0xFFFF) return fail(ctx, SFGPU_E_UNSUPPORTED, "expression too long");
        sfgpu_expr_op* dprog = nullptr;
        if (prog.empty()) prog.push_back(sfgpu_expr_op{SFGPU_X_CONST, 0, 0});
        rc = dev_upload(ctx, prog.data(), prog.size(), &dprog);
        if (rc) return rc;
        c.g0 = dprog;
        c.g1 = ctx->expr_cols_dev;
        c.g2 = ctx->expr_csrs_dev;
        c.n0 = lens[0] | (lens[1] << 16);
        c.p0 = d.p0;
        break;
      }
      case SFGPU_K_UNI: {
        c.p0 = d.p0;
        const int64_t* col = nullptr;
        bool by_value = d.p1 == 1;
        uint32_t want = dm.n_entities;
        if (by_value && d.aux0 != 0xFFFFFFFFu && d.aux0 < ctx->cols.size())
          want = ctx->colls[ctx->cols[d.aux0].coll].n_rows;
        int rc = column(d.aux0, want, &col);
        if (rc) return rc;
        if (by_value && col && want < dm.n_values) return fail(ctx, SFGPU_E_INVALID, "value column too short");
        c.g0 = col;
        if (by_value) c.flags |= SFGPU_CF_COL_BY_VALUE;
        const int64_t* mask = nullptr;
        rc = column(d.aux1, dm.n_entities, &mask);
        if (rc) return rc;
        c.g1 = mask;
        break;
      }
      case SFGPU_K_PAIR_CSR_EQUAL: {
        if (d.aux0 >= ctx->csrs.size()) return fail(ctx, SFGPU_E_INVALID, "unknown csr");
        const Csr& g = ctx->csrs[d.aux0];
        if (g.n_rows != dm.n_entities) return fail(ctx, SFGPU_E_INVALID, "csr row count != entity count");
        for (uint32_t v : g.col)
          if (v >= g.n_rows) return fail(ctx, SFGPU_E_INVALID, "csr column index out of range");
        // partner lists: b is a partner of e iff the ordered pair (min,max) passes
        // `left.id < right.id && left.neighbors.contains(right.id)` (row index == id).
        std::vector<std::vector<uint32_t>> partners(g.n_rows);
        for (uint32_t a = 0; a < g.n_rows; ++a)
          for (uint32_t j = g.row_ptr[a]; j < g.row_ptr[a + 1]; ++j) {
            uint32_t b = g.col[j];
            if (a < b) {
              partners[a].push_back(b);
              partners[b].push_back(a);
            }
          }
        std::vector<uint32_t> rp(g.n_rows + 1, 0), ci;
        for (uint32_t a = 0; a < g.n_rows; ++a) {
          auto& p = partners[a];
          std::sort(p.begin(), p.end());
          p.erase(std::unique(p.begin(), p.end()), p.end());
          rp[a + 1] = rp[a] + (uint32_t)p.size();
          ci.insert(ci.end(), p.begin(), p.end());
        }
        uint32_t *drp = nullptr, *dci = nullptr;
        int rc = dev_upload(ctx, rp.data(), rp.size(), &drp);
        if (rc) return rc;
        rc = dev_upload(ctx, ci.data(), ci.size(), &dci);
        if (rc) return rc;
        c.g0 = drp;
        c.g1 = dci;
        // retained partner-value counts (uint16 [n_entities][n_values], unstaged): only when the table is
        // small enough and no entity has more than 65535 partners
        uint32_t maxdeg = 0;
        for (uint32_t a2 = 0; a2 < g.n_rows; ++a2) maxdeg = std::max(maxdeg, rp[a2 + 1] - rp[a2]);
        c.off0 = 0xFFFFFFFFu;
        if ((uint64_t)dm.n_entities * dm.n_values <= (1ull << 24) && maxdeg <= 65535)
          unstaged_bytes[k] = align_up(dm.n_entities * dm.n_values * 2, 16);
        break;
      }
      case SFGPU_K_PAIR_KEY_EQUAL: {
        const int64_t* col = nullptr;
        int rc = column(d.aux0, dm.n_entities, &col);
        if (rc) return rc;
        c.g0 = col;
        c.g2 = col ? ctx->cols[d.aux0].dev32 : nullptr;
        c.p0 = d.p0;
        c.p1 = d.p1;
        if (d.p1 == 0) return fail(ctx, SFGPU_E_INVALID, "PAIR_KEY_EQUAL needs p1 != 0 (the variable must enter the key)");
        // join arity: 2 (bi, the default) .. 5 (penta) — nary_incremental/higher_arity/{tri,quad,penta}.rs
        c.pad = (d.aux1 == 0xFFFFFFFFu || d.aux1 == 0) ? 2 : d.aux1;
        if (c.pad < 2 || c.pad > 5) return fail(ctx, SFGPU_E_INVALID, "PAIR_KEY_EQUAL: aux1 (join arity) must be 2..5");
        // key range over every (entity, value)
        int64_t kmin = std::numeric_limits<int64_t>::max(), kmax = std::numeric_limits<int64_t>::min();
        for (uint32_t e = 0; e < dm.n_entities; ++e) {
          int64_t base = (col ? ctx->cols[d.aux0].host[e] : 0) * d.p0;
          int64_t k0 = base, k1 = base + (int64_t)(dm.n_values - 1) * d.p1;
          kmin = std::min(kmin, std::min(k0, k1));
          kmax = std::max(kmax, std::max(k0, k1));
        }
        if (dm.n_entities == 0) kmin = kmax = 0;
        if (kmax - kmin + 1 > (1 << 24)) return fail(ctx, SFGPU_E_UNSUPPORTED, "join key range too wide for a dense table");
        c.p2 = kmin;
        c.n0 = (uint32_t)(kmax - kmin + 1);
        c.off0 = off;
        off = align_up(off + c.n0 * 4, 16);
        break;
      }
      case SFGPU_K_GROUP: {
        const int64_t* col = nullptr;
        int rc = column(d.aux0, dm.n_entities, &col);
        if (rc) return rc;
        c.g0 = col;
        c.g2 = col ? ctx->cols[d.aux0].dev32 : nullptr;
        if (d.p0 == 1) c.flags |= SFGPU_CF_COMPLEMENT;
        c.p1 = d.p1;
        if (d.aux1 != 0xFFFFFFFFu) {  // per-value weight offset column (key-dependent weight)
          if (d.aux1 >= ctx->cols.size() || ctx->colls[ctx->cols[d.aux1].coll].n_rows < dm.n_values)
            return fail(ctx, SFGPU_E_INVALID, "GROUP: aux1 must be a column with one row per value");
          c.g1 = ctx->cols[d.aux1].dev;
        }
        c.off0 = off;
        off = align_up(off + dm.n_values * 4, 16);
        c.off1 = off;
        off = align_up(off + dm.n_values * 8, 16);
        break;
      }
      case SFGPU_K_RUNS: {
        const int64_t* col = nullptr;
        int rc = column(d.aux0, dm.n_entities, &col);
        if (rc) return rc;
        if (!col) return fail(ctx, SFGPU_E_INVALID, "RUNS: aux0 must be the entity column holding the point");
        if (d.p0 <= 0 || (uint64_t)d.p0 * dm.n_values >= (1ull << 28))
          return fail(ctx, SFGPU_E_INVALID, "RUNS: p0 (number of points) out of range");
        for (int64_t v : ctx->cols[d.aux0].host)
          if (v < 0 || v >= d.p0) return fail(ctx, SFGPU_E_INVALID, "RUNS: point outside [0, p0)");
        c.g0 = col;
        c.n0 = (uint32_t)d.p0;
        // indexed_presence views (aux1): p1 = lo | hi << 32, a horizon / range inside [0, p0)
        if (d.aux1 != 0xFFFFFFFFu && d.aux1 != 0) {
          if (d.aux1 > 3) return fail(ctx, SFGPU_E_INVALID, "RUNS: aux1 (indexed_presence view) must be 0..3");
          const int64_t lo = d.p1 & 0xFFFFFFFFll, hi = (d.p1 >> 32) & 0xFFFFFFFFll;
          if (d.aux1 != 3 && (lo > hi || hi > d.p0)) return fail(ctx, SFGPU_E_INVALID, "RUNS: range outside [0, p0)");
          c.p0 = d.aux1;
          c.p1 = lo;
          c.p2 = hi;
        }
        c.off0 = off;
        off = align_up(off + dm.n_values * c.n0 * 4, 16);
        break;
      }
      case SFGPU_K_PROJECT_GROUP: {
        if (d.aux0 >= ctx->csrs.size()) return fail(ctx, SFGPU_E_INVALID, "PROJECT_GROUP: unknown csr");
        const Csr& g = ctx->csrs[d.aux0];
        if (g.n_rows != dm.n_entities) return fail(ctx, SFGPU_E_INVALID, "PROJECT_GROUP: csr row count != entity count");
        if (d.p0 <= 0 || (uint64_t)d.p0 * dm.n_values >= (1ull << 31))
          return fail(ctx, SFGPU_E_INVALID, "PROJECT_GROUP: p0 (key offsets per value) out of range");
        const uint32_t nnz = g.row_ptr[g.n_rows];
        const std::vector<int64_t>* amounts = nullptr;
        if (d.aux1 != 0xFFFFFFFFu) {
          if (d.aux1 >= ctx->cols.size() || ctx->cols[d.aux1].host.size() < nnz)
            return fail(ctx, SFGPU_E_INVALID, "PROJECT_GROUP: aux1 must be a column with one row per csr entry");
          amounts = &ctx->cols[d.aux1].host;
        }
        std::vector<int64_t> em((size_t)nnz * 2);
        for (uint32_t e = 0; e < g.n_rows; ++e) {
          if (g.row_ptr[e + 1] - g.row_ptr[e] > 8)
            return fail(ctx, SFGPU_E_UNSUPPORTED, "PROJECT_GROUP: more than 8 projected rows per entity (MAX_EMITS)");
          for (uint32_t j = g.row_ptr[e]; j < g.row_ptr[e + 1]; ++j) {
            if ((int64_t)g.col[j] >= d.p0) return fail(ctx, SFGPU_E_INVALID, "PROJECT_GROUP: key offset >= p0");
            em[(size_t)j * 2] = g.col[j];
            em[(size_t)j * 2 + 1] = amounts ? (*amounts)[j] : 1;
          }
        }
        if (em.empty()) em.assign(2, 0);
        int64_t* dem = nullptr;
        int rc = dev_upload(ctx, em.data(), em.size(), &dem);
        if (rc) return rc;
        uint32_t* drp = nullptr;
        rc = dev_upload(ctx, g.row_ptr.data(), g.row_ptr.size(), &drp);
        if (rc) return rc;
        c.g0 = drp;
        c.g1 = dem;
        c.n0 = (uint32_t)d.p0;
        c.off0 = off;
        off = align_up(off + dm.n_values * c.n0 * 4, 16);
        c.off1 = off;
        off = align_up(off + dm.n_values * c.n0 * 8, 16);
        break;
      }
      case SFGPU_K_LOAD_BALANCE: {
        const int64_t* col = nullptr;
        int rc = column(d.aux0, dm.n_entities, &col);
        if (rc) return rc;
        c.g0 = col;
        c.off0 = off;
        off = align_up(off + dm.n_values * 8, 16);
        c.off1 = off;
        off = align_up(off + dm.n_values * 4, 16);
        c.off2 = off;
        off = align_up(off + 24, 16);
        ctx->has_load_balance = true;
        break;
      }
      case SFGPU_K_EXISTS_FLAT: {
        if (d.collection >= ctx->colls.size()) return fail(ctx, SFGPU_E_INVALID, "unknown collection");
        const int64_t* col = nullptr;
        int rc = column(d.aux0, ctx->colls[d.collection].n_rows, &col);
        if (rc) return rc;
        c.g0 = col;
        c.n0 = ctx->colls[d.collection].n_rows;
        c.p0 = d.p0;
        unstaged_bytes[k] = align_up(dm.n_elem_rows * 4, 16);
        break;
      }
      case SFGPU_K_LIST_PATH_COST: {
        if (d.aux0 >= ctx->mats.size()) return fail(ctx, SFGPU_E_INVALID, "unknown matrix");
        const Matrix& mt = ctx->mats[d.aux0];
        if (mt.rows < dm.n_elem_rows || mt.cols < dm.n_elem_rows)
          return fail(ctx, SFGPU_E_INVALID, "matrix smaller than the element collection");
        if (d.p0 < 0 || d.p0 >= (int64_t)mt.rows) return fail(ctx, SFGPU_E_INVALID, "depot out of range");
        c.g0 = mt.dev;
        c.n0 = mt.cols;
        if (mt.i32) c.flags |= SFGPU_CF_MATRIX_I32;
        c.p0 = d.p0;
        c.off0 = off;
        off = align_up(off + dm.n_owners * 8, 16);
        break;
      }
      case SFGPU_K_LIST_SUM: {
        const int64_t* col = nullptr;
        int rc = column(d.aux0, dm.n_elem_rows, &col);
        if (rc) return rc;
        if (!col) return fail(ctx, SFGPU_E_INVALID, "LIST_SUM needs an element column");
        c.g0 = col;
        c.off0 = off;
        off = align_up(off + dm.n_owners * 8, 16);
        break;
      }
    }
  }
  dm.stage_bytes = align_up(off, 16);
  for (uint32_t k = 0; k < dm.n_cons; ++k)
    if (unstaged_bytes[k]) {
      dm.cons[k].off0 = off;
      off += unstaged_bytes[k];
    }
  dm.block_bytes = align_up(off, 128);
  size_t total = (size_t)dm.block_bytes * dm.R;
  CU(cudaMalloc((void**)&dm.state, total));
  ctx->dev_allocs.push_back(dm.state);
  CU(cudaMalloc((void**)&ctx->scratch_state, total));
  ctx->dev_allocs.push_back(ctx->scratch_state);
  // host image of the planning state
  {
    std::vector<char> img(total, 0);
    for (uint32_t r = 0; r < dm.R; ++r) {
      char* b = img.data() + (size_t)r * dm.block_bytes;
      if (dm.has_scalar) {
        const ScalarVar& v = ctx->svars[0];
        const int32_t* src = v.init.data() + (v.per_replica ? (size_t)r * dm.n_entities : 0);
        memcpy(b + dm.off_var, src, (size_t)dm.n_entities * 4);
      }
      if (dm.has_list) {
        const ListVar& v = ctx->lvars[0];
        size_t oc = v.per_replica ? (size_t)r * (dm.n_owners + 1) : 0;
        const uint32_t* o = v.offsets.data() + oc;
        size_t ebase = 0;
        if (v.per_replica)
          for (uint32_t q = 0; q < r; ++q) ebase += v.offsets[(size_t)q * (dm.n_owners + 1) + dm.n_owners];
        memcpy(b + dm.off_offsets, o, (size_t)(dm.n_owners + 1) * 4);
        if (o[dm.n_owners]) memcpy(b + dm.off_elems, v.elems.data() + ebase, (size_t)o[dm.n_owners] * 4);
      }
    }
    CU(cudaMemcpyAsync(dm.state, img.data(), total, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  ctx->staged = dm.stage_bytes + 1024 <= (uint32_t)ctx->max_smem_optin;
  {
    int rc = sfgpu_configure_scalar(ctx);  // monomorphised program choice + shared-memory opt-in
    if (rc) return rc;
  }
  if (dm.fast_list && dm.fast_stage_bytes + 1024 > (uint32_t)ctx->max_smem_optin) {
    dm.fast_list = 0;
    dm.nearby_ok = 0;
  }
  if (dm.fast_list && dm.fast_pc >= 0) {
    // gather-friendly copies of the path-cost matrix for the fast kernel: narrowest cell type that
    // holds every cost, plus the transpose (shared when the matrix is symmetric)
    const Matrix& mt = ctx->mats[ctx->cons[dm.fast_pc].d.aux0];
    int64_t mx = 0;
    bool sym = mt.rows == mt.cols;
    for (int64_t c : mt.host) mx = std::max(mx, c);
    for (uint32_t i = 0; sym && i < mt.rows; ++i)
      for (uint32_t j = i + 1; j < mt.cols; ++j)
        if (mt.host[(size_t)i * mt.cols + j] != mt.host[(size_t)j * mt.cols + i]) { sym = false; break; }
    // uint16 cells select the 32-bit arithmetic of the fast kernels (FastArith): narrow models only
    dm.fm_u16 = mx < 65536 && mt.rows == mt.cols && dm.fast_narrow ? 1 : 0;
    if (mt.rows != mt.cols) {
      dm.fm_row = mt.dev;  // rectangular: no transpose trick, plain int32 gathers
      dm.fm_col = nullptr;
    } else {
      // Internal relabelling for gather locality: a greedy nearest-neighbour tour from row 0 gives
      // elements that are close in cost neighbouring internal ids, so the ~20 candidates of one source
      // (all near it) read few 32-byte sectors of its matrix row. Only the record ids, these matrix
      // copies and the neighbour lists live in the internal id space; rows, state and outputs do not.
      const uint32_t n = mt.rows;
      std::vector<uint32_t> relabel(n), inverse(n);
      if (n <= 20000 && !getenv("SFGPU_NO_RELABEL")) {
        std::vector<uint8_t> seen(n, 0);
        uint32_t cur = 0;
        for (uint32_t k = 0; k < n; ++k) {
          seen[cur] = 1;
          inverse[k] = cur;
          relabel[cur] = k;
          const int64_t* row = mt.host.data() + (size_t)cur * n;
          int64_t best = INT64_MAX;
          uint32_t nxt = cur;
          for (uint32_t y = 0; y < n; ++y)
            if (!seen[y] && row[y] < best) {
              best = row[y];
              nxt = y;
            }
          cur = nxt;
        }
      } else {
        for (uint32_t i = 0; i < n; ++i) relabel[i] = inverse[i] = i;
      }
      while (relabel.size() < dm.n_elem_rows) {  // element rows beyond the matrix keep their own id
        inverse.push_back((uint32_t)relabel.size());
        relabel.push_back((uint32_t)relabel.size());
      }
      ctx->relabel_host = relabel;
      ctx->inverse_host = inverse;
      uint32_t* drl = nullptr;
      int rc = dev_upload(ctx, relabel.data(), relabel.size(), &drl);
      if (rc) return rc;
      dm.relabel = drl;
      if (dm.fm_u16) {
        std::vector<uint16_t> a(mt.host.size()), t(mt.host.size());
        for (uint32_t i = 0; i < n; ++i)
          for (uint32_t j = 0; j < n; ++j) {
            const uint16_t c = (uint16_t)mt.host[(size_t)i * n + j];
            a[(size_t)relabel[i] * n + relabel[j]] = c;
            t[(size_t)relabel[j] * n + relabel[i]] = c;
          }
        uint16_t *da = nullptr, *dt = nullptr;
        rc = dev_upload(ctx, a.data(), a.size(), &da);
        if (rc) return rc;
        dm.fm_row = da;
        dm.fm_col = da;
        if (!sym) {
          rc = dev_upload(ctx, t.data(), t.size(), &dt);
          if (rc) return rc;
          dm.fm_col = dt;
        }
      } else {
        std::vector<int32_t> a(mt.host.size()), t(mt.host.size());
        for (uint32_t i = 0; i < n; ++i)
          for (uint32_t j = 0; j < n; ++j) {
            const int32_t c = (int32_t)mt.host[(size_t)i * n + j];
            a[(size_t)relabel[i] * n + relabel[j]] = c;
            t[(size_t)relabel[j] * n + relabel[i]] = c;
          }
        int32_t *da = nullptr, *dt = nullptr;
        rc = dev_upload(ctx, a.data(), a.size(), &da);
        if (rc) return rc;
        dm.fm_row = da;
        dm.fm_col = da;
        if (!sym) {
          rc = dev_upload(ctx, t.data(), t.size(), &dt);
          if (rc) return rc;
          dm.fm_col = dt;
        }
      }
    }
    if (!dm.fm_col) {  // rectangular matrix: keep the generic kernel
      dm.fast_list = 0;
      dm.nearby_ok = 0;
    }
  }
  if (dm.nearby_ok) {
    // static neighbour lists: for every element row, the other rows by (distance, row) ascending
    const Matrix& mt = ctx->mats[ctx->cons[dm.fast_pc].d.aux0];
    const uint32_t nrow = dm.n_elem_rows;
    dm.nbr_stride = nrow > 0 ? nrow - 1 : 0;
    std::vector<uint32_t> nbr((size_t)nrow * dm.nbr_stride);
    std::vector<uint32_t> order(dm.nbr_stride);
    // (internal ids on both sides: row relabel[x] lists relabel[y]; equal distances may come in any
    // order, the kernels rank them by scan index)
    const std::vector<uint32_t>& rl = ctx->relabel_host;
    for (uint32_t x = 0; x < nrow; ++x) {
      uint32_t k = 0;
      for (uint32_t y = 0; y < nrow; ++y)
        if (y != x) order[k++] = y;
      const int64_t* row = mt.host.data() + (size_t)x * mt.cols;
      std::sort(order.begin(), order.end(), [row](uint32_t l, uint32_t r) {
        return row[l] != row[r] ? row[l] < row[r] : l < r;
      });
      uint32_t* dst = nbr.data() + (size_t)rl[x] * dm.nbr_stride;
      for (uint32_t q = 0; q < dm.nbr_stride; ++q) dst[q] = rl[order[q]];
    }
    uint32_t* dn = nullptr;
    int rc = dev_upload(ctx, nbr.data(), nbr.size(), &dn);
    if (rc) return rc;
    dm.nbr = dn;
    // key layout: distance bits above scan-index bits; 32-bit keys when both fit in 31 bits
    {
      const Matrix& mt2 = ctx->mats[ctx->cons[dm.fast_pc].d.aux0];
      int64_t mx = 0;
      for (int64_t c : mt2.host) mx = std::max(mx, c);
      uint32_t dbits = 1, sbits = 1;
      while ((1ll << dbits) <= mx) ++dbits;
      while ((1u << sbits) <= dm.elem_cap + dm.n_owners + 1) ++sbits;
      ctx->nb_key32 = dbits + sbits <= 31;
      ctx->nb_scan_bits = ctx->nb_key32 ? sbits : 24;
    }
  }
  if (dm.fast_list && dm.compact_bytes && (!dm.fm_u16 || 16 + dm.compact_bytes + 1024 > (uint32_t)ctx->max_smem_optin))
    dm.compact_bytes = 0;
  {
    int rc = sfgpu_configure_list(ctx);
    if (rc) return rc;
    if (dm.nearby_ok) {
      rc = sfgpu_configure_nearby(ctx);
      if (rc) return rc;
      // retained nearby neighbourhood: the per-replica protocol words exist from commit on (every state writer
      // checks them); the cache itself is allocated by the first cached step
      CU(cudaMalloc((void**)&dm.nbc_tag, (size_t)dm.R * NBC_WORDS * 4));
      ctx->dev_allocs.push_back(dm.nbc_tag);
      CU(cudaMemsetAsync(dm.nbc_tag, 0, (size_t)dm.R * NBC_WORDS * 4, ctx->stream));
      ctx->nbc_off = getenv("SFGPU_NO_NBCACHE") != nullptr;
    }
  }
  if (dm.has_list) {
    if (((size_t)dm.n_owners + 1) * 4 > 48 * 1024) {
      if (((size_t)dm.n_owners + 1) * 4 > (size_t)ctx->max_smem_optin) return fail(ctx, SFGPU_E_UNSUPPORTED, "too many list owners");
      CU(cudaFuncSetAttribute(init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)dm.n_owners + 1) * 4)));
    }
    int bytes = (int)apply_smem_bytes(dm);
    if (bytes > ctx->max_smem_optin) return fail(ctx, SFGPU_E_UNSUPPORTED, "list variable too large for the apply kernel");
    CU(cudaFuncSetAttribute(apply_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  }
  // initialize_all
  init_kernel<<<dm.R, 256, ((size_t)dm.n_owners + 1) * 4, ctx->stream>>>(dm, dm.state);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->building = false;
  ctx->committed = true;
  if (out_scores) return sfgpu_committed_scores(ctx, out_scores);
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

}  // extern "C"

// ------------------------------------------------------------------------------------------
// scoring
// ------------------------------------------------------------------------------------------
int sfgpu_launch_score_scalar(sfgpu_ctx* ctx, int kind, uint64_t n_total, const uint64_t* d_offs, const uint32_t* d_rows,
                              const uint64_t* d_edit_offs, int64_t* d_scores, uint8_t* d_doable, ForageArgs* forage,
                              uint32_t* out_chunks);
int sfgpu_launch_score_list(sfgpu_ctx* ctx, int kind, uint64_t n_total, const uint64_t* d_offs, const uint32_t* d_rows,
                            int64_t* d_scores, uint8_t* d_doable, ForageArgs* forage, uint32_t* out_chunks);

namespace {

enum ScoreKind { SK_CHANGE, SK_SWAP, SK_COMPOUND, SK_LIST_CHANGE, SK_LIST_SWAP, SK_LIST_REVERSE, SK_SUBLIST_CHANGE, SK_SUBLIST_SWAP,
                 SK_K_OPT };

int launch_score(sfgpu_ctx* ctx, ScoreKind kind, uint64_t n_total, const uint64_t* d_offs, const uint32_t* d_rows,
                 const uint64_t* d_edit_offs, int64_t* d_scores, uint8_t* d_doable) {
  if (kind <= SK_COMPOUND) return sfgpu_launch_score_scalar(ctx, (int)kind, n_total, d_offs, d_rows, d_edit_offs, d_scores, d_doable, nullptr, nullptr);
  return sfgpu_launch_score_list(ctx, (int)kind, n_total, d_offs, d_rows, d_scores, d_doable, nullptr, nullptr);
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Host-pointer path: stage through pinned memory, H2D, kernel, D2H, synchronize.
int score_entry(sfgpu_ctx* ctx, ScoreKind kind, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets, const uint32_t* rows,
                const uint64_t* edit_offsets, int64_t* out_scores, uint8_t* out_doable) {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (!cand_offsets || !out_scores || !out_doable) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  const DevModel& dm = ctx->dm;
  bool scalar = kind == SK_CHANGE || kind == SK_SWAP || kind == SK_COMPOUND;
  if (scalar && !dm.has_scalar) return fail(ctx, SFGPU_E_STATE, "model has no scalar variable");
  if (!scalar && !dm.has_list) return fail(ctx, SFGPU_E_STATE, "model has no list variable");
  if (kind == SK_COMPOUND && ctx->has_load_balance)
    return fail(ctx, SFGPU_E_UNSUPPORTED, "compound moves over a load_balance constraint");
  if (kind == SK_SWAP && ctx->has_load_balance)
    return fail(ctx, SFGPU_E_UNSUPPORTED, "swap moves over a load_balance constraint");
  CU(cudaSetDevice(ctx->device));
  const uint32_t row_words = scalar ? 2 : 4;
  if (flags & SFGPU_DEVICE_IO) {
    const uint64_t n_total = n_candidates;
    if (n_total == 0) return SFGPU_OK;
    if (!rows) return fail(ctx, SFGPU_E_INVALID, "null rows");
    return launch_score(ctx, kind, n_total, cand_offsets, rows, edit_offsets, out_scores, out_doable);
  }
  if (cand_offsets[0] != 0) return fail(ctx, SFGPU_E_INVALID, "cand_offsets[0] must be 0");
  for (uint32_t r = 0; r < dm.R; ++r)
    if (cand_offsets[r + 1] < cand_offsets[r]) return fail(ctx, SFGPU_E_INVALID, "cand_offsets must be non-decreasing");
  const uint64_t n = cand_offsets[dm.R];
  if (n != n_candidates) return fail(ctx, SFGPU_E_INVALID, "n_candidates != cand_offsets[R]");
  if (n == 0) return SFGPU_OK;
  if (!rows) return fail(ctx, SFGPU_E_INVALID, "null rows");
  uint64_t n_rows = n;
  if (kind == SK_COMPOUND) {
    if (!edit_offsets) return fail(ctx, SFGPU_E_INVALID, "null edit_offsets");
    n_rows = edit_offsets[n];
  }
  // layout of both staging areas: [offsets (R+1) u64][edit offsets (n+1) u64][rows][scores][doable]
  size_t o_off = 0;
  size_t o_eoff = (o_off + (dm.R + 1) * 8 + 15) / 16 * 16;
  size_t o_rows = (o_eoff + (kind == SK_COMPOUND ? (n + 1) * 8 : 0) + 15) / 16 * 16;
  size_t o_scores = (o_rows + n_rows * row_words * 4 + 15) / 16 * 16;
  size_t o_doable = o_scores + n * 16;
  size_t total = o_doable + (n + 15) / 16 * 16;
  // page-locked caller buffers are copied directly; pageable ones go through the pinned staging area
  const bool in_pinned = is_pinned(rows), out_pinned = is_pinned(out_scores) && is_pinned(out_doable);
  const size_t pin_need = (in_pinned ? o_rows : o_scores) + (out_pinned ? 0 : total - o_scores);
  rc = ensure_staging(ctx, std::max<size_t>(pin_need, 64), total);
  if (rc) return rc;
  char* pin = (char*)ctx->pin;
  char* dv = (char*)ctx->dscr;
  memcpy(pin + o_off, cand_offsets, (dm.R + 1) * 8);
  if (kind == SK_COMPOUND) memcpy(pin + o_eoff, edit_offsets, (n + 1) * 8);
  if (in_pinned) {
    CU(cudaMemcpyAsync(dv, pin, o_rows, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dv + o_rows, rows, n_rows * row_words * 4, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    memcpy(pin + o_rows, rows, n_rows * row_words * 4);
    CU(cudaMemcpyAsync(dv, pin, o_scores, cudaMemcpyHostToDevice, ctx->stream));
  }
  rc = launch_score(ctx, kind, n, (const uint64_t*)(dv + o_off), (const uint32_t*)(dv + o_rows),
                    kind == SK_COMPOUND ? (const uint64_t*)(dv + o_eoff) : nullptr, (int64_t*)(dv + o_scores),
                    (uint8_t*)(dv + o_doable));
  if (rc) return rc;
  if (out_pinned) {
    CU(cudaMemcpyAsync(out_scores, dv + o_scores, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_doable, dv + o_doable, n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  } else {
    char* stage = pin + (in_pinned ? o_rows : o_scores);
    CU(cudaMemcpyAsync(stage, dv + o_scores, total - o_scores, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy(out_scores, stage, n * 16);
    memcpy(out_doable, stage + (o_doable - o_scores), n);
  }
  return SFGPU_OK;
}

}  // namespace

extern "C" {

int32_t sfgpu_score_change(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets, const uint32_t* rows,
                           int64_t* out_scores, uint8_t* out_doable) try {
  return score_entry(ctx, SK_CHANGE, flags, n_candidates, cand_offsets, rows, nullptr, out_scores, out_doable);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_score_swap(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets, const uint32_t* rows,
                         int64_t* out_scores, uint8_t* out_doable) try {
  return score_entry(ctx, SK_SWAP, flags, n_candidates, cand_offsets, rows, nullptr, out_scores, out_doable);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_score_compound(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                             const uint64_t* edit_offsets, const uint32_t* edit_rows, int64_t* out_scores,
                             uint8_t* out_doable) try {
  return score_entry(ctx, SK_COMPOUND, flags, n_candidates, cand_offsets, edit_rows, edit_offsets, out_scores, out_doable);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_score_list_change(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                                const uint32_t* rows, int64_t* out_scores, uint8_t* out_doable) try {
  return score_entry(ctx, SK_LIST_CHANGE, flags, n_candidates, cand_offsets, rows, nullptr, out_scores, out_doable);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_score_list_swap(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets, const uint32_t* rows,
                              int64_t* out_scores, uint8_t* out_doable) try {
  return score_entry(ctx, SK_LIST_SWAP, flags, n_candidates, cand_offsets, rows, nullptr, out_scores, out_doable);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_score_list_reverse(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                                 const uint32_t* rows, int64_t* out_scores, uint8_t* out_doable) try {
  return score_entry(ctx, SK_LIST_REVERSE, flags, n_candidates, cand_offsets, rows, nullptr, out_scores, out_doable);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_score_sublist_change(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                                   const uint32_t* rows, int64_t* out_scores, uint8_t* out_doable) try {
  return score_entry(ctx, SK_SUBLIST_CHANGE, flags, n_candidates, cand_offsets, rows, nullptr, out_scores, out_doable);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_score_sublist_swap(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                                 const uint32_t* rows, int64_t* out_scores, uint8_t* out_doable) try {
  return score_entry(ctx, SK_SUBLIST_SWAP, flags, n_candidates, cand_offsets, rows, nullptr, out_scores, out_doable);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_score_k_opt(sfgpu_ctx* ctx, uint32_t flags, uint64_t n_candidates, const uint64_t* cand_offsets,
                          const uint32_t* rows, int64_t* out_scores, uint8_t* out_doable) try {
  return score_entry(ctx, SK_K_OPT, flags, n_candidates, cand_offsets, rows, nullptr, out_scores, out_doable);
} SFGPU_API_CATCH(ctx)

// ------------------------------------------------------------------------------------------
namespace {
// acceptor + forager replay over materialised scores: BestScore goes through chunk partials on the whole
// machine + the finish kernel; AcceptedCount(N) and gated batches use the one-CTA-per-replica kernel
int launch_argbest(sfgpu_ctx* ctx, const ForageDev& f, const uint64_t* d_offs, const int64_t* d_scores,
                   const uint8_t* d_doable, const uint64_t* d_seeds, const int64_t* d_ref, uint32_t* d_idx,
                   int64_t* d_best, uint32_t* d_eval, uint64_t n_hint) {
  const uint32_t R = ctx->dm.R;
  if (f.accepted_limit == 0 && !f.gates) {
    uint32_t chunks = std::max<uint32_t>(1, (uint32_t)((uint64_t)ctx->sm_count * 4 / R));
    const uint64_t per_rep = n_hint / std::max<uint32_t>(R, 1);
    while (chunks > 1 && per_rep / chunks < 2048) chunks /= 2;
    chunks = std::min<uint32_t>(chunks, 64);
    int rc = ensure_partials(ctx, (size_t)chunks * R * sizeof(ChunkPartial));
    if (rc) return rc;
    ForageArgs fa{f, d_ref, (ChunkPartial*)ctx->partials, 0};
    argbest_partial_kernel<<<dim3(chunks, R), 256, 0, ctx->stream>>>(f, d_offs, d_scores, d_doable, d_ref, fa.partials);
    forage_finish_kernel<<<R, 256, 0, ctx->stream>>>(ctx->dm, fa, chunks, d_offs, nullptr, d_scores, d_doable, d_seeds,
                                                    d_idx, d_best, d_eval);
    ctx->launches += 2;
  } else {
    argbest_kernel<<<R, 1024, 0, ctx->stream>>>(f, d_offs, d_scores, d_doable, d_seeds, d_ref, d_idx, d_best, d_eval);
    ctx->launches++;
  }
  CU(cudaGetLastError());
  return SFGPU_OK;
}
}  // namespace

int32_t sfgpu_argbest_gated(sfgpu_ctx* ctx, uint32_t flags, const sfgpu_forage_params* params,
                            const uint64_t* cand_offsets, const int64_t* scores, const uint8_t* doable,
                            const uint8_t* gates, const uint64_t* step_seeds, const int64_t* ref_scores,
                            uint32_t* out_index, int64_t* out_best, uint32_t* out_evaluated) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (!params || !cand_offsets || !scores || !doable || !out_index || !out_best)
    return fail(ctx, SFGPU_E_INVALID, "null pointer");
  if (params->acceptor < 0 || params->acceptor > 3 || params->tie_mode < 0 || params->tie_mode > 1)
    return fail(ctx, SFGPU_E_INVALID, "bad forage params");
  if ((params->acceptor != 0 || gates) && !ref_scores)
    return fail(ctx, SFGPU_E_INVALID, "acceptor / improvement gates need ref_scores (last_step_score)");
  CU(cudaSetDevice(ctx->device));
  const uint32_t R = ctx->dm.R;
  ForageDev f{params->acceptor, params->tie_mode, params->accepted_limit, gates};
  if (flags & SFGPU_DEVICE_IO)  // the candidate total is not known on the host: assume a saturating batch
    return launch_argbest(ctx, f, cand_offsets, scores, doable, step_seeds, ref_scores, out_index, out_best,
                          out_evaluated, (uint64_t)R << 20);
  const uint64_t n = cand_offsets[R];
  auto a16 = [](size_t v) { return (v + 15) / 16 * 16; };
  size_t o_off = 0, o_seed = a16(o_off + (R + 1) * 8), o_ref = a16(o_seed + R * 8), o_scores = a16(o_ref + R * 32);
  size_t o_doable = o_scores + n * 16;
  size_t o_gates = (o_doable + n + 15) / 16 * 16;
  size_t o_idx = (o_gates + (gates ? n : 0) + 15) / 16 * 16, o_best = o_idx + (R * 4 + 15) / 16 * 16, o_eval = o_best + R * 16;
  size_t total = o_eval + (R * 4 + 15) / 16 * 16;
  rc = ensure_staging(ctx, total, total);
  if (rc) return rc;
  char* pin = (char*)ctx->pin;
  char* dv = (char*)ctx->dscr;
  memcpy(pin + o_off, cand_offsets, (R + 1) * 8);
  if (step_seeds) memcpy(pin + o_seed, step_seeds, R * 8); else memset(pin + o_seed, 0, R * 8);
  if (ref_scores) memcpy(pin + o_ref, ref_scores, R * 32); else memset(pin + o_ref, 0, R * 32);
  memcpy(pin + o_scores, scores, n * 16);
  memcpy(pin + o_doable, doable, n);
  if (gates) {
    memcpy(pin + o_gates, gates, n);
    f.gates = (const uint8_t*)(dv + o_gates);
  }
  CU(cudaMemcpyAsync(dv, pin, o_idx, cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_argbest(ctx, f, (const uint64_t*)(dv + o_off), (const int64_t*)(dv + o_scores),
                      (const uint8_t*)(dv + o_doable), (const uint64_t*)(dv + o_seed), (const int64_t*)(dv + o_ref),
                      (uint32_t*)(dv + o_idx), (int64_t*)(dv + o_best), (uint32_t*)(dv + o_eval), n);
  if (rc) return rc;
  CU(cudaMemcpyAsync(pin + o_idx, dv + o_idx, total - o_idx, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(out_index, pin + o_idx, R * 4);
  memcpy(out_best, pin + o_best, R * 16);
  if (out_evaluated) memcpy(out_evaluated, pin + o_eval, R * 4);
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_argbest(sfgpu_ctx* ctx, uint32_t flags, const sfgpu_forage_params* params,
                      const uint64_t* cand_offsets, const int64_t* scores, const uint8_t* doable,
                      const uint64_t* step_seeds, const int64_t* ref_scores, uint32_t* out_index,
                      int64_t* out_best, uint32_t* out_evaluated) try {
  return sfgpu_argbest_gated(ctx, flags, params, cand_offsets, scores, doable, nullptr, step_seeds, ref_scores,
                             out_index, out_best, out_evaluated);
} SFGPU_API_CATCH(ctx)

// ------------------------------------------------------------------------------------------
namespace {
int apply_entry(sfgpu_ctx* ctx, int kind, uint32_t flags, const uint32_t* rows, const uint8_t* mask,
                const uint64_t* d_offsets, const uint32_t* d_index) {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (!rows) return fail(ctx, SFGPU_E_INVALID, "null rows");
  const DevModel& dm = ctx->dm;
  bool scalar = kind < 2;
  if (scalar && !dm.has_scalar) return fail(ctx, SFGPU_E_STATE, "model has no scalar variable");
  if (!scalar && !dm.has_list) return fail(ctx, SFGPU_E_STATE, "model has no list variable");
  if (kind == 1 && ctx->has_load_balance)
    return fail(ctx, SFGPU_E_UNSUPPORTED, "swap moves over a load_balance constraint");
  CU(cudaSetDevice(ctx->device));
  const uint32_t R = dm.R;
  const uint32_t* d_rows = rows;
  const uint8_t* d_mask = mask;
  if (!(flags & SFGPU_DEVICE_IO)) {
    size_t row_bytes = (size_t)R * (scalar ? 8 : 16);
    size_t total = row_bytes + (R + 15) / 16 * 16;
    rc = ensure_staging(ctx, total, total);
    if (rc) return rc;
    memcpy(ctx->pin, rows, row_bytes);
    if (mask) memcpy((char*)ctx->pin + row_bytes, mask, R);
    CU(cudaMemcpyAsync(ctx->dscr, ctx->pin, total, cudaMemcpyHostToDevice, ctx->stream));
    d_rows = (const uint32_t*)ctx->dscr;
    d_mask = mask ? (const uint8_t*)ctx->dscr + row_bytes : nullptr;
  }
  if (scalar)
    apply_scalar_kernel<<<R, 32, 0, ctx->stream>>>(dm, kind, d_rows, d_mask, d_offsets, d_index);
  else
    apply_list_kernel<<<R, apply_threads(ctx), apply_smem_bytes(dm), ctx->stream>>>(dm, kind, d_rows, d_mask, d_offsets, d_index);
  ctx->launches++;
  CU(cudaGetLastError());
  if (!(flags & SFGPU_DEVICE_IO)) CU(cudaStreamSynchronize(ctx->stream));
  return SFGPU_OK;
}
}  // namespace

int32_t sfgpu_apply_change(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask) try {
  return apply_entry(ctx, 0, flags, rows, mask, nullptr, nullptr);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_apply_swap(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask) try {
  return apply_entry(ctx, 1, flags, rows, mask, nullptr, nullptr);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_apply_list_change(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask) try {
  return apply_entry(ctx, 2, flags, rows, mask, nullptr, nullptr);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_apply_list_swap(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask) try {
  return apply_entry(ctx, 3, flags, rows, mask, nullptr, nullptr);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_apply_list_reverse(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask) try {
  return apply_entry(ctx, 4, flags, rows, mask, nullptr, nullptr);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_apply_sublist_change(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask) try {
  return apply_entry(ctx, 5, flags, rows, mask, nullptr, nullptr);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_apply_sublist_swap(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask) try {
  return apply_entry(ctx, 6, flags, rows, mask, nullptr, nullptr);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_apply_k_opt(sfgpu_ctx* ctx, uint32_t flags, const uint32_t* rows, const uint8_t* mask) try {
  return apply_entry(ctx, 7, flags, rows, mask, nullptr, nullptr);
} SFGPU_API_CATCH(ctx)
int32_t sfgpu_apply_winners(sfgpu_ctx* ctx, int32_t move_kind, const uint64_t* cand_offsets,
                            const uint32_t* batch_rows, const uint32_t* index) try {
  if (move_kind < 0 || move_kind > 7) return fail(ctx, SFGPU_E_INVALID, "bad move kind");
  if (!cand_offsets || !index) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  return apply_entry(ctx, move_kind, SFGPU_DEVICE_IO, batch_rows, nullptr, cand_offsets, index);
} SFGPU_API_CATCH(ctx)

// ------------------------------------------------------------------------------------------
namespace {
__global__ void gather_scores_kernel(const __grid_constant__ DevModel m, const char* __restrict__ state,
                                     int64_t* __restrict__ out) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m.R) return;
  const int64_t* cs = (const int64_t*)(state + (size_t)r * m.block_bytes + m.off_score);
  out[r * 2] = cs[0];
  out[r * 2 + 1] = cs[1];
}

int read_scores(sfgpu_ctx* ctx, const char* state, int64_t* out) {
  const uint32_t R = ctx->dm.R;
  int rc = ensure_staging(ctx, (size_t)R * 16, (size_t)R * 16);
  if (rc) return rc;
  gather_scores_kernel<<<(R + 127) / 128, 128, 0, ctx->stream>>>(ctx->dm, state, (int64_t*)ctx->dscr);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ctx->pin, ctx->dscr, (size_t)R * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->pin, (size_t)R * 16);
  return SFGPU_OK;
}
}  // namespace

int32_t sfgpu_committed_scores(sfgpu_ctx* ctx, int64_t* out_scores) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (!out_scores) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  CU(cudaSetDevice(ctx->device));
  return read_scores(ctx, ctx->dm.state, out_scores);
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_evaluate_all(sfgpu_ctx* ctx, int64_t* out_scores) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (!out_scores) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  CU(cudaSetDevice(ctx->device));
  const DevModel& dm = ctx->dm;
  CU(cudaMemcpyAsync(ctx->scratch_state, dm.state, (size_t)dm.block_bytes * dm.R, cudaMemcpyDeviceToDevice,
                     ctx->stream));
  init_kernel<<<dm.R, 256, ((size_t)dm.n_owners + 1) * 4, ctx->stream>>>(dm, ctx->scratch_state);
  ctx->launches++;
  CU(cudaGetLastError());
  return read_scores(ctx, ctx->scratch_state, out_scores);
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_get_scalar_state(sfgpu_ctx* ctx, uint32_t variable, int32_t* out_values) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  const DevModel& dm = ctx->dm;
  if (variable != 0 || !dm.has_scalar || !out_values) return fail(ctx, SFGPU_E_INVALID, "unknown scalar variable");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy2D(out_values, (size_t)dm.n_entities * 4, dm.state + dm.off_var, dm.block_bytes,
                  (size_t)dm.n_entities * 4, dm.R, cudaMemcpyDeviceToHost));
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_list_capacity(sfgpu_ctx* ctx, uint32_t variable, uint32_t* out_capacity) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (variable != 0x80000000u || !ctx->dm.has_list || !out_capacity) return fail(ctx, SFGPU_E_INVALID, "unknown list variable");
  *out_capacity = ctx->dm.elem_cap;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_get_list_state(sfgpu_ctx* ctx, uint32_t variable, uint32_t* out_offsets, uint32_t* out_elems) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  const DevModel& dm = ctx->dm;
  if (variable != 0x80000000u || !dm.has_list || !out_offsets || !out_elems)
    return fail(ctx, SFGPU_E_INVALID, "unknown list variable");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaMemcpy2D(out_offsets, (size_t)(dm.n_owners + 1) * 4, dm.state + dm.off_offsets, dm.block_bytes,
                  (size_t)(dm.n_owners + 1) * 4, dm.R, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy2D(out_elems, (size_t)dm.elem_cap * 4, dm.state + dm.off_elems, dm.block_bytes,
                  (size_t)dm.elem_cap * 4, dm.R, cudaMemcpyDeviceToHost));
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_pack_best_keys(sfgpu_ctx* ctx, int64_t* out_keys) try {
  int rc = check_committed(ctx);
  if (rc) return rc;
  if (!out_keys) return fail(ctx, SFGPU_E_INVALID, "null pointer");
  CU(cudaSetDevice(ctx->device));
  pack_keys_kernel<<<(ctx->dm.R + 127) / 128, 128, 0, ctx->stream>>>(ctx->dm, out_keys);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_last_kernel_ns(sfgpu_ctx* ctx, uint64_t* out_ns) try {
  uint32_t n = 0;
  return sfgpu_kernel_times_ns(ctx, 1, out_ns, &n);
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_kernel_times_ns(sfgpu_ctx* ctx, uint32_t max_n, uint64_t* out_ns, uint32_t* out_n) try {
  if (!ctx || !out_ns || !out_n) return SFGPU_E_INVALID;
  if (ctx->ev_count == 0) return fail(ctx, SFGPU_E_STATE, "no scoring kernel launched yet");
  const uint64_t avail = std::min<uint64_t>(ctx->ev_count, sfgpu_ctx::EV_RING);
  const uint32_t n = (uint32_t)std::min<uint64_t>(avail, max_n);
  for (uint32_t i = 0; i < n; ++i) {
    const uint64_t slot = (ctx->ev_count - n + i) % sfgpu_ctx::EV_RING;
    CU(cudaEventSynchronize(ctx->ev_b[slot]));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev_a[slot], ctx->ev_b[slot]));
    out_ns[i] = (uint64_t)((double)ms * 1e6);
  }
  *out_n = n;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_launch_count(sfgpu_ctx* ctx, uint64_t* out_count) try {
  if (!ctx || !out_count) return SFGPU_E_INVALID;
  *out_count = ctx->launches;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

int32_t sfgpu_scalar_program(sfgpu_ctx* ctx, int32_t* out_program) try {
  if (!ctx || !out_program) return SFGPU_E_INVALID;
  if (!ctx->committed) return fail(ctx, SFGPU_E_STATE, "model not committed");
  *out_program = ctx->spec_id;
  return SFGPU_OK;
} SFGPU_API_CATCH(ctx)

}  // extern "C"

// ---- launch wrappers for the other translation units -------------------------------------------------------
int sfgpu_launch_apply_list(sfgpu_ctx* ctx, int kind, const uint32_t* d_rows, const uint8_t* d_mask,
                            const uint64_t* d_offsets, const uint32_t* d_index) {
  const DevModel& dm = ctx->dm;
  apply_list_kernel<<<dm.R, apply_threads(ctx), apply_smem_bytes(dm), ctx->stream>>>(dm, kind, d_rows, d_mask, d_offsets, d_index);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}
int sfgpu_launch_apply_list_kinds(sfgpu_ctx* ctx, const uint32_t* d_rows, const int32_t* d_kinds) {
  const DevModel& dm = ctx->dm;
  if (dm.has_scalar) {
    apply_scalar_kernel<<<dm.R, 32, 0, ctx->stream>>>(dm, 0, d_rows, nullptr, nullptr, nullptr, d_kinds);
    ctx->launches++;
  }
  if (dm.has_list) apply_list_kernel<<<dm.R, apply_threads(ctx), apply_smem_bytes(dm), ctx->stream>>>(dm, 2, d_rows, nullptr, nullptr, nullptr, d_kinds);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}
int sfgpu_launch_argbest_counts(sfgpu_ctx* ctx, const ForageDev& f, const uint64_t* d_offs, const uint32_t* d_counts,
                                const uint32_t* d_skip, const int64_t* d_scores, const uint8_t* d_doable,
                                const uint64_t* d_seeds, const int64_t* d_ref, uint32_t* d_idx, int64_t* d_best,
                                uint32_t* d_eval) {
  // a window holds a few hundred pulls: a small CTA spends less time in the block scans' barriers
  const uint32_t stride = ctx->dm.R > 1 ? (uint32_t)std::min<uint64_t>(0xFFFFFFFFull, ctx->argbest_stride_hint) : 0;
  const int threads = (stride && stride <= 2048) ? 256 : 1024;
  argbest_kernel<<<ctx->dm.R, threads, 0, ctx->stream>>>(f, d_offs, d_scores, d_doable, d_seeds, d_ref, d_idx, d_best, d_eval,
                                                         d_counts, d_skip);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}
int sfgpu_launch_apply_scalar(sfgpu_ctx* ctx, int kind, const uint32_t* d_rows, const uint8_t* d_mask,
                              const uint64_t* d_offsets, const uint32_t* d_index) {
  const DevModel& dm = ctx->dm;
  apply_scalar_kernel<<<dm.R, 32, 0, ctx->stream>>>(dm, kind, d_rows, d_mask, d_offsets, d_index);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}
int sfgpu_launch_forage_finish(sfgpu_ctx* ctx, const ForageArgs& fa, uint32_t chunks, const uint64_t* d_offs,
                               const uint32_t* d_rows, const int64_t* d_scores, const uint8_t* d_doable,
                               const uint64_t* d_seeds, uint32_t* d_idx, int64_t* d_best, uint32_t* d_eval) {
  forage_finish_kernel<<<ctx->dm.R, 256, 0, ctx->stream>>>(ctx->dm, fa, chunks, d_offs, d_rows, d_scores, d_doable, d_seeds,
                                                          d_idx, d_best, d_eval);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}
int sfgpu_launch_argbest_ordered(sfgpu_ctx* ctx, const ForageDev& f, const uint64_t* d_offs, const int64_t* d_scores,
                                 const uint8_t* d_doable, const uint64_t* d_seeds, const int64_t* d_ref, uint32_t* d_idx,
                                 int64_t* d_best, uint32_t* d_eval) {
  argbest_kernel<<<ctx->dm.R, 1024, 0, ctx->stream>>>(f, d_offs, d_scores, d_doable, d_seeds, d_ref, d_idx, d_best, d_eval);
  ctx->launches++;
  CU(cudaGetLastError());
  return SFGPU_OK;
}
