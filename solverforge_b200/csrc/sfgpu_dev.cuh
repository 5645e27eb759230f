// Device-side model description and small device helpers shared by every kernel of libsfgpu.
// HBM layout: one contiguous "replica block" per replica holding the planning
// state and the retained aggregates of that replica; static facts are shared by all replicas.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sfgpu.h"

#define SFGPU_MAX_CONS 32  // constraint tuple size of the reference (api/constraint_set/incremental.rs:339-408: tuples up to 32)
#define SFGPU_NONE (-1)

struct WeightDev {
  int32_t fn, level;
  int64_t a, b;
};

struct ConsDev {
  int32_t kind;
  int32_t sign;  // -1 penalty, +1 reward (constraint/incremental.rs:70-76)
  WeightDev w;
  uint32_t off0, off1, off2;  // byte offsets of this constraint's retained sections in the replica block
  uint32_t n0;                // table length / stride
  const void* g0;             // shared static arrays (CSR row_ptr, column, matrix ...)
  const void* g1;
  const void* g2;             // int32 copy of the int64 column g0 when every value fits (narrow programs), or null
  int64_t p0, p1, p2;
  uint32_t flags;
  uint32_t pad;
};
#define SFGPU_CF_MATRIX_I32 1u   // g0 is int32 cells (cooked costs all fit in 31 bits)
#define SFGPU_CF_COMPLEMENT 2u
#define SFGPU_CF_COL_BY_VALUE 4u

// column expressions of SFGPU_K_JOIN_EXPR (sfgpu_expr_op), evaluated per joined pair
struct ExprTables {
  const int64_t* const* cols;   // device pointer of every column, by column id
  const uint32_t* const* csrs;  // {row_ptr, col} device pointers of every CSR, by 2 * csr id
};

struct DevModel {
  char* state;  // [R][block_bytes]
  uint32_t block_bytes;
  uint32_t stage_bytes;  // prefix of a block the scoring kernels stage into shared memory
  uint32_t R;
  uint32_t n_cons;
  uint32_t off_score;
  // scalar variable
  uint32_t has_scalar, n_entities, n_values, allows_unassigned, off_var;
  // list variable
  uint32_t has_list, n_owners, n_elem_rows, elem_cap, off_offsets, off_elems;
  // fast list path: packed per-route / per-position / per-slot records
  uint32_t fast_list;       // 1 when the constraint program fits the specialised list-change kernel
  uint32_t fast_stage_bytes;  // records + pos_of (nearby kernels)
  uint32_t fast_score_bytes;  // records only (score kernels)
  // compact 8-byte copies of the records (narrow models with 16-bit ids / positions / values): halves
  // the shared-memory wavefronts of the score kernel and its staged bytes. 0 = not available.
  uint32_t compact_bytes;     // Route8 + Pos8 + Slot8, contiguous from off_route8
  uint32_t off_route8, off_pos8, off_slot8;
  uint32_t off_route_rec, off_pos_rec, off_slot_rec;
  int32_t fast_pc, fast_ls;  // constraint indices of the path-cost / list-sum constraint, or -1
  // device-side nearby neighbourhood
  uint32_t nearby_ok;        // 1 when the fused generate+score+forage kernel can run on this model
  uint32_t fast_narrow;      // 1 when every fast-path delta fits int32 (multipliers, sums, offsets < 2^29)
  uint32_t fm_u16;           // 1: fm_row / fm_col hold uint16 cells (every cost < 65536), else int32
  const void* fm_row;        // path-cost matrix, row-major: d(x, .) is row x
  const void* fm_col;        // its transpose: d(., x) is row x (same array when the matrix is symmetric)
  const uint32_t* relabel;   // element id -> internal id of the fast records / fm_* / nbr (locality order), or null
  uint32_t off_pos_of;       // uint32[n_elem_rows]: (owner << 16 | position) of each element, or 0xFFFFFFFF
  uint32_t nbr_stride;       // entries per row of `nbr`
  const uint32_t* nbr;       // static: for every element row, the other rows sorted by (distance, row)
  // retained nearby neighbourhood (sfgpu_nearby.cuh, NBC_*): 16 words per replica — what the committed moves since the
  // last cached step touched; every kernel that writes a replica block keeps it honest. Null = no retained cache.
  uint32_t* nbc_tag;
  ConsDev cons[SFGPU_MAX_CONS];
};

// 16-byte records so one candidate needs four 128-bit shared-memory loads
struct RouteRec {   // per owner
  uint32_t base, len;
  int64_t sum;      // retained LIST_SUM aggregate (0 without a LIST_SUM constraint)
};
struct PosRec {     // per flat element position
  uint32_t elem;
  int32_t rem;      // path-cost delta of removing this element from its route
  int32_t val;      // LIST_SUM column value of the element
  uint32_t owner;   // owner (route) index of this position
};
struct SlotRec {    // per insertion slot (owner e, position p in 0..=len), index = base + e + p
  uint32_t a, b;    // neighbours of the slot (depot at the ends)
  int32_t gap;      // cost of the existing leg a -> b (0 for an empty route)
  uint32_t where;   // (owner << 16) | position of the slot
};

// Per (replica, source position) partial of the fused nearby step.
struct SrcPartial {
  int64_t best_h, best_s;
  uint32_t n_best, n_accepted;
  uint32_t first_lane, pad;
};

// indices into DevModel::cons of the (sorted) scalar constraints of a monomorphised program, -1 = none
struct SpecIdx {
  int32_t k[4];
};

struct Score2 {
  int64_t hard, soft;
};

// acceptor + forager parameters of one step (sfgpu_forage_params)
struct ForageDev {
  int32_t acceptor, tie_mode;
  uint32_t accepted_limit;
  // per-candidate improvement gates of evaluate_candidate (phase/localsearch/evaluation.rs:76-111), or
  // null: bit 0 = requires_hard_improvement, bit 1 = requires_score_improvement (argbest only)
  const uint8_t* gates;
};

// Per (replica, chunk) partial of the fused score+forage path: best accepted score of the chunk,
// how many accepted rows equal it, the first such row (pull index inside the replica) and the
// number of accepted rows.
struct ChunkPartial {
  int64_t best_h, best_s;
  uint32_t n_best, n_accepted;
  uint32_t first_idx;
  uint32_t second_idx;  // second accepted row equal to the best (0xFFFFFFFF when n_best < 2): the common
                        // tie of two rows never needs the ordered rescan
};

// ---- argument blocks of the whole-step kernels (defined here so host-only translation units can fill them) ----
#define MOVE_CHANGE 0
#define MOVE_SWAP 1
struct NearbyArgs {
  ForageDev f;
  uint32_t max_nearby;           // <= 32
  uint32_t scan_bits;            // low bits of a key that hold the scan index
  const uint64_t* step_seeds;    // [R] or null
  const int64_t* ref_scores;     // [R][4] or null
  SrcPartial* partials;          // [R][elem_cap]
  uint32_t* out_rows;            // [R][elem_cap * max_nearby][4] or null
  int64_t* out_scores;           // [R][elem_cap * max_nearby][2] or null
  uint8_t* out_doable;           // or null
  uint64_t* out_offsets;         // [R+1] or null: candidate offsets of the materialised batch
  // retained neighbourhood (nearby_step_cached_kernel): per (replica, source element) the score deltas of its
  // max_nearby candidates, the reference element of each candidate slot and {k-th distance, route bloom}
  uint32_t cache;                // 1: the cached kernel ran / runs for this step
  int2* c_delta;                 // [R][n_elem_rows][max_nearby] (dh, ds); dh == INT32_MIN: no candidate on this lane
  uint32_t* c_ident;             // [R][n_elem_rows][max_nearby] reference element | append << 15 | route << 16
  uint32_t* c_work;              // [R][elem_cap + 1] {n, flat positions of the sources to regenerate this step}
  uint4* c_meta;                 // [R][n_elem_rows][2] {k-th distance, 7 words of 8-bit route codes (4 lanes per word)}
};

// nbc_tag words of one replica
#define NBC_WORDS 16
#define NBC_STATE 0   // 0: cache invalid; 1: cache == committed state; 2: one ListChange move committed since
#define NBC_K 1       // max_nearby the cache was built with
#define NBC_COUNT 2   // candidates per source it was built with
#define NBC_A 3       // routes the committed move touched
#define NBC_B 4
#define NBC_HOT 5     // 5 words: the moved element and the elements whose append slot appeared / vanished (or NONE)
#define NBC_N_HOT 5
#define NBC_STATS 10  // 3 words: sources served from kept deltas / re-scored / regenerated since commit

// every kernel that rewrites a replica block outside the ListChange commit protocol calls this
__device__ __forceinline__ void nbc_invalidate(uint32_t* nbc_tag, uint32_t r) {
  if (nbc_tag) nbc_tag[(size_t)r * NBC_WORDS + NBC_STATE] = 0;
}

struct ChangeStepArgs {
  ForageDev f;
  uint32_t ents_per_cta;         // multiple of blockDim.x
  const uint64_t* step_seeds;    // [R] or null
  const int64_t* ref_scores;     // [R][4] or null
  ChunkPartial* partials;        // [R][gridDim.x]; first_idx holds e * (k + 1) + v
  uint32_t* out_rows;            // [R][n_entities * (k + 1)][2] or null (materialised batch, padded)
  int64_t* out_scores;
  uint8_t* out_doable;
  uint64_t* out_offsets;         // [R + 1] or null
  uint32_t* out_counts;          // [R] or null: rows of the replica's neighbourhood (the rest of its stride is padding)
};

struct IndexStepArgs {
  ForageDev f;
  uint32_t per_chunk;            // candidates per CTA (multiple of blockDim.x)
  uint32_t min_size, max_size;   // segment sizes
  const uint64_t* step_seeds;    // [R] or null
  const int64_t* ref_scores;     // [R][4] or null
  ChunkPartial* partials;        // [R][gridDim.x]
};

// SimulatedAnnealing (acceptor/simulated_annealing.rs): per-level temperatures, calibrated from the first
// `sample_size` worsening candidates the phase evaluates (CalibrationState, :57-88), Boltzmann acceptance on the first
// differing level (:338-375), geometric decay per step once calibrated (:416-431).
struct SaParams {
  double decay, hc_temp, fallback, neg_log_target;  // -ln(target acceptance), computed on the host
  uint32_t sample_size;
  int32_t never_hard;  // HardRegressionPolicy::NeverAcceptHardRegression
};
struct SaState {
  double temp[2];
  int64_t sum[2];
  uint32_t cnt[2];
  uint32_t calibrated, pad;
};

// ---- union of neighbourhoods (sfgpu_union.cuh) ----
#define UNION_MAX_CHILDREN 8
#define UNION_NONE 0xFFFFFFFFu

struct UnionChildDev {
  int32_t family;   // SFGPU_FAM_*
  uint32_t p0, p1;  // max_nearby | min_size, max_size
  uint32_t pad;
  uint64_t weight;
};

struct UnionArgs {
  ForageDev f;
  uint32_t n_children;
  int32_t union_order;      // SFGPU_UNION_*
  int32_t order;            // SFGPU_ORDER_* of every leaf (VecUnionSelector::child_context)
  uint32_t desc;            // descriptor index of the list owner collection (salts)
  uint32_t sdesc;           // descriptor index of the scalar variable's entity collection
  uint32_t window;          // rows a child may emit in this pass
  uint32_t t_cap;           // union pulls per replica in this pass (n_children * window)
  uint32_t scan_bits;       // low key bits holding the scan index (nearby families)
  UnionChildDev child[UNION_MAX_CHILDREN];
  const uint64_t* step_seeds;    // [R]
  const uint64_t* step_indices;  // [R] or null (0), or step_counter (one value for all replicas) with step_index_shared
  int32_t step_index_shared;
  const int64_t* ref_scores;     // [R][4] or null
  uint32_t* rows;                // [R][n_children][window][4]
  uint32_t* n_emit;              // [R][n_children]
  uint32_t* ended;               // [R][n_children] 1 = the child's cursor ended inside the window
  uint32_t* sched;               // [R][t_cap] child << 28 | child-local index
  uint32_t* n_sched;             // [R]
  uint32_t* stream_end;          // [R] 1 = the union stream ended inside the window
  int64_t* scores;               // [R][t_cap][2] in union pull order (gathered from child_scores)
  int64_t* child_scores;         // [R][n_children][window][2] by child-local index
  uint8_t* child_doable;         // [R][n_children][window]
  uint8_t* doable;               // [R][t_cap]
  uint64_t* offsets;             // [R + 1] = r * t_cap
  uint32_t* done;                // [R] step already complete (earlier pass): skip
  const uint32_t* win_r;         // [R] rows a child of replica r may emit in this pass (<= window), or null = window
  uint32_t* next_win;            // [R] or null: adaptive window of the replica's next step (resident loop)
  uint32_t win_max;              // largest adaptive window
  uint32_t win_shift;            // this pass offers win_r[r] << win_shift rows per child
};

__device__ __forceinline__ int64_t weight_eval(const WeightDev& w, int64_t x) {
  switch (w.fn) {
    case SFGPU_W_CONST: return w.a;
    case SFGPU_W_LINEAR: return w.a * x + w.b;
    case SFGPU_W_SQUARE: return w.a * x * x + w.b;
    case SFGPU_W_ABSDIFF: {
      int64_t d = x - w.b;
      return w.a * (d < 0 ? -d : d);
    }
    case SFGPU_W_PAIRS: return w.a * (x * (x - 1) / 2);
    default: {
      int64_t d = x - w.b;
      return d > 0 ? w.a * d : 0;
    }
  }
}

// postfix evaluation of a column expression for the pair (a, b) joined on key value v
__device__ __forceinline__ int64_t expr_eval(const sfgpu_expr_op* __restrict__ ops, uint32_t n, const ExprTables& t,
                                             uint32_t a, uint32_t b, int64_t v, int64_t va = -1, int64_t vb = -1) {
  int64_t s[8];
  int sp = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const sfgpu_expr_op o = ops[i];
    switch (o.op) {
      case SFGPU_X_CONST: s[sp++] = o.imm; break;
      case SFGPU_X_A_COL: s[sp++] = t.cols[o.arg][a]; break;
      case SFGPU_X_B_COL: s[sp++] = t.cols[o.arg][b]; break;
      case SFGPU_X_A_IDX: s[sp++] = (int64_t)a; break;
      case SFGPU_X_B_IDX: s[sp++] = (int64_t)b; break;
      case SFGPU_X_VALUE: s[sp++] = v; break;
      case SFGPU_X_A_VAL: s[sp++] = va; break;
      case SFGPU_X_B_VAL: s[sp++] = vb; break;
      case SFGPU_X_NEG: s[sp - 1] = -s[sp - 1]; break;
      case SFGPU_X_ABS: s[sp - 1] = s[sp - 1] < 0 ? -s[sp - 1] : s[sp - 1]; break;
      case SFGPU_X_NOT: s[sp - 1] = s[sp - 1] == 0 ? 1 : 0; break;
      case SFGPU_X_SELECT: {
        const int64_t e = s[sp - 1], th = s[sp - 2], c = s[sp - 3];
        sp -= 2;
        s[sp - 1] = c != 0 ? th : e;
        break;
      }
      case SFGPU_X_CSR_CONTAINS: {
        const int64_t x = s[sp - 1], row = s[sp - 2];
        --sp;
        const uint32_t* rp = t.csrs[2 * o.arg];
        const uint32_t* ci = t.csrs[2 * o.arg + 1];
        int64_t hit = 0;
        if (row >= 0 && row < o.imm)  // imm = row count of the CSR (set at commit)
          for (uint32_t j = rp[row]; j < rp[row + 1]; ++j) hit |= (int64_t)ci[j] == x ? 1 : 0;
        s[sp - 1] = hit;
        break;
      }
      default: {
        const int64_t y = s[sp - 1], x = s[sp - 2];
        --sp;
        int64_t r = 0;
        switch (o.op) {
          case SFGPU_X_ADD: r = x + y; break;
          case SFGPU_X_SUB: r = x - y; break;
          case SFGPU_X_MUL: r = x * y; break;
          case SFGPU_X_MIN: r = x < y ? x : y; break;
          case SFGPU_X_MAX: r = x > y ? x : y; break;
          case SFGPU_X_MOD: r = y == 0 ? 0 : x % y; break;
          case SFGPU_X_EQ: r = x == y; break;
          case SFGPU_X_NE: r = x != y; break;
          case SFGPU_X_LT: r = x < y; break;
          case SFGPU_X_LE: r = x <= y; break;
          case SFGPU_X_GT: r = x > y; break;
          case SFGPU_X_GE: r = x >= y; break;
          case SFGPU_X_AND: r = (x != 0) && (y != 0); break;
          case SFGPU_X_OR: r = (x != 0) || (y != 0); break;
        }
        s[sp - 1] = r;
      }
    }
  }
  return s[0];
}

// SFGPU_K_JOIN_EXPR: sum over the B rows joined to key v of [filter(a, b)] * weight(x(a, b))
// (CrossBiConstraint::evaluate restricted to one A row, cross_bi_incremental/state.rs:260-460)
__device__ __forceinline__ int64_t join_expr_contrib(const ConsDev& c, uint32_t a, int32_t v) {
  if (v < 0) return 0;  // key None joins nothing
  const sfgpu_expr_op* ops = (const sfgpu_expr_op*)c.g0;
  const ExprTables t{(const int64_t* const*)c.g1, (const uint32_t* const*)c.g2};
  const uint32_t nf = c.n0 & 0xFFFFu, nw = c.n0 >> 16;
  uint32_t lo = (uint32_t)v, hi = (uint32_t)v + 1;
  const uint32_t* bucket = nullptr;
  if (c.p0 >= 0) {
    const uint32_t* rp = t.csrs[2 * c.p0];
    bucket = t.csrs[2 * c.p0 + 1];
    lo = rp[v];
    hi = rp[v + 1];
  }
  int64_t total = 0;
  for (uint32_t j = lo; j < hi; ++j) {
    const uint32_t b = bucket ? bucket[j] : j;
    if (nf && expr_eval(ops, nf, t, a, b, v, v, -1) == 0) continue;
    total += weight_eval(c.w, nw ? expr_eval(ops + nf, nw, t, a, b, v, v, -1) : 0);
  }
  return total;
}

// ---- SFGPU_K_PAIR_KEY_EXPR: keyed self-join of entity rows with pair filter / weight expressions ----------------
// program layout: [left key][right key][pair filter][pair weight]; lengths: pad = lkl | lkr << 16, n0 = lf | lw << 16.
// retained section (int32): headL[n_keys] nxtL[n] prvL[n] headR[n_keys] nxtR[n] prvR[n]
struct PairKeyLists {
  int32_t *headL, *nxtL, *prvL, *headR, *nxtR, *prvR;
};
__device__ __forceinline__ PairKeyLists pke_lists(const ConsDev& c, char* block, uint32_t n_entities) {
  int32_t* base = (int32_t*)(block + c.off0);
  const uint32_t nk = (uint32_t)c.p1;
  PairKeyLists l;
  l.headL = base;
  l.nxtL = l.headL + nk;
  l.prvL = l.nxtL + n_entities;
  l.headR = l.prvL + n_entities;
  l.nxtR = l.headR + nk;
  l.prvR = l.nxtR + n_entities;
  return l;
}
__device__ __forceinline__ bool pke_directed(const ConsDev& c) { return (c.pad >> 16) != 0; }
// key of row (e, v): which = 0 left key, 1 right key; -1 = no key
__device__ __forceinline__ int64_t pke_key(const ConsDev& c, const ExprTables& t, int which, uint32_t e, int32_t v) {
  if (v < 0) return -1;
  const sfgpu_expr_op* ops = (const sfgpu_expr_op*)c.g0;
  const uint32_t lkl = c.pad & 0xFFFFu, lkr = c.pad >> 16;
  const int64_t k = which ? expr_eval(ops + lkl, lkr, t, e, e, v, v, v) : expr_eval(ops, lkl, t, e, e, v, v, v);
  return (k < 0 || k >= c.p1) ? -1 : k;
}
// score of the joined pair (left row l with value vl, right row r with value vr), 0 when the filter rejects it
__device__ __forceinline__ int64_t pke_pair(const ConsDev& c, const ExprTables& t, uint32_t l, int32_t vl, uint32_t r, int32_t vr) {
  const sfgpu_expr_op* ops = (const sfgpu_expr_op*)c.g0 + (c.pad & 0xFFFFu) + (c.pad >> 16);
  const uint32_t lf = c.n0 & 0xFFFFu, lw = c.n0 >> 16;
  if (lf && expr_eval(ops, lf, t, l, r, vl, vl, vr) == 0) return 0;
  return weight_eval(c.w, lw ? expr_eval(ops + lf, lw, t, l, r, vl, vl, vr) : 0);
}

// C(n, k) for the keyed self-joins of arity k + 1 (tuples that gain / lose one member of a bucket of n rows);
// exact while the result fits int64 (every prefix product of i consecutive integers is divisible by i!)
__device__ __forceinline__ int64_t choose_small(int64_t n, int k) {
  if (k == 1) return n;
  if (n < k) return 0;
  int64_t r = 1;
  for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
  return r;
}

// adds sign * v to the constraint's level
__device__ __forceinline__ void add_level(Score2& s, const ConsDev& c, int64_t v) {
  int64_t sv = c.sign < 0 ? -v : v;
  if (c.w.level == 0) s.hard += sv; else s.soft += sv;
}

__device__ __forceinline__ bool score_less(int64_t h1, int64_t s1, int64_t h2, int64_t s2) {
  return h1 != h2 ? h1 < h2 : s1 < s2;
}

__device__ __forceinline__ uint64_t splitmix64_dev(uint64_t v) {
  v += 0x9E3779B97F4A7C15ull;
  v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
  v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
  return v ^ (v >> 31);
}

// ---- TMA (bulk async copy) staging of a replica block into shared memory ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Two ranges behind one barrier: [src0, +bytes0) -> dst0 and [src1, +bytes1) -> dst1 (bytes1 <= 64 KB).
__device__ __forceinline__ void stage_block2(char* dst0, const char* src0, uint32_t bytes0, char* dst1,
                                             const char* src1, uint32_t bytes1, uint64_t* bar);

// Stage `bytes` (multiple of 16, 16 B aligned both sides) of global memory into shared memory
// with TMA bulk copies; every thread of the CTA must call this. `bar` is a shared mbarrier slot.
__device__ __forceinline__ void stage_block(char* smem_dst, const char* gmem_src, uint32_t bytes, uint64_t* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, bytes);
    const uint32_t piece = 32768;
    for (uint32_t o = 0; o < bytes; o += piece) {
      uint32_t n = bytes - o < piece ? bytes - o : piece;
      tma_bulk_g2s(smem_dst + o, gmem_src + o, n, bar);
    }
  }
  mbar_wait(bar, 0);
}

__device__ __forceinline__ void stage_block2(char* dst0, const char* src0, uint32_t bytes0, char* dst1,
                                             const char* src1, uint32_t bytes1, uint64_t* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, bytes0 + bytes1);
    tma_bulk_g2s(dst0, src0, bytes0, bar);
    const uint32_t piece = 32768;
    for (uint32_t o = 0; o < bytes1; o += piece) {
      uint32_t n = bytes1 - o < piece ? bytes1 - o : piece;
      tma_bulk_g2s(dst1 + o, src1 + o, n, bar);
    }
  }
  mbar_wait(bar, 0);
}
