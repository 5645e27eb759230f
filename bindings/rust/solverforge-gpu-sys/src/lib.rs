//! Raw bindings to libsfgpu — the B200 batched re-scoring library behind SolverForge's
//! `Director` / `ConstraintSet` surface. One declaration per symbol of `include/sfgpu.h`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub const SFGPU_OK: i32 = 0;
pub const SFGPU_E_INVALID: i32 = -1;
pub const SFGPU_E_UNSUPPORTED: i32 = -2;
pub const SFGPU_E_CUDA: i32 = -3;
pub const SFGPU_E_NCCL: i32 = -4;
pub const SFGPU_E_OOM: i32 = -5;
pub const SFGPU_E_STATE: i32 = -6;
pub const SFGPU_DEVICE_IO: u32 = 1;
pub const SFGPU_CTX_LEGACY_DEFAULT_STREAM: u64 = 1;
pub const SFGPU_CTX_GENERIC_KERNELS: u64 = 2;

pub const SFGPU_W_CONST: i32 = 0;
pub const SFGPU_W_LINEAR: i32 = 1;
pub const SFGPU_W_SQUARE: i32 = 2;
pub const SFGPU_W_EXCESS: i32 = 3;
pub const SFGPU_W_ABSDIFF: i32 = 4;
pub const SFGPU_W_PAIRS: i32 = 5;
pub const SFGPU_PENALTY: i32 = 0;
pub const SFGPU_REWARD: i32 = 1;
pub const SFGPU_K_UNI: i32 = 1;
pub const SFGPU_K_PAIR_CSR_EQUAL: i32 = 2;
pub const SFGPU_K_PAIR_KEY_EQUAL: i32 = 3;
pub const SFGPU_K_EXISTS_FLAT: i32 = 4;
pub const SFGPU_K_GROUP: i32 = 5;
pub const SFGPU_K_LIST_PATH_COST: i32 = 6;
pub const SFGPU_K_LIST_SUM: i32 = 7;
pub const SFGPU_K_LOAD_BALANCE: i32 = 8;
pub const SFGPU_K_PROJECT_GROUP: i32 = 9;
pub const SFGPU_K_RUNS: i32 = 10;

#[repr(C)]
pub struct sfgpu_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct sfgpu_weight {
    pub fn_: i32,
    pub level: i32,
    pub a: i64,
    pub b: i64,
}

#[repr(C)]
pub struct sfgpu_constraint_desc {
    pub kind: i32,
    pub impact: i32,
    pub weight: sfgpu_weight,
    pub collection: u32,
    pub variable: u32,
    pub aux0: u32,
    pub aux1: u32,
    pub p0: i64,
    pub p1: i64,
    pub name: *const c_char,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct sfgpu_expr_op {
    pub op: i32,
    pub arg: u32,
    pub imm: i64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct sfgpu_union_child {
    pub family: i32,
    pub p0: u32,
    pub p1: u32,
    pub reserved: u32,
    pub weight: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct sfgpu_union_desc {
    pub n_children: u32,
    pub union_order: i32,
    pub selection_order: i32,
    pub window: u32,
    pub max_window: u32,
    pub reserved: u32,
    pub children: [sfgpu_union_child; 8],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct sfgpu_forage_params {
    pub acceptor: i32,
    pub tie_mode: i32,
    pub accepted_limit: u32,
    pub reserved: u32,
}

/// `acceptor`: 1 HillClimbing, 2 LateAcceptance(`late_size`), 3 GreatDeluge(`acceptor_real` = rain speed),
/// 4 StepCountingHillClimbing(`step_count_limit`), 5 DiversifiedLateAcceptance(`late_size`, `acceptor_real` = tolerance),
/// 6 SimulatedAnnealing(`late_size` = calibration samples, `acceptor_real` = decay), 7 TabuSearch(`late_size` =
/// entity | value << 8 | move << 16 | undo_move << 24 tenures, `step_count_limit` bit 0 = aspiration).
/// `reserved` bit 0: windowed speculation for `sfgpu_solve_nearby_list_change` with AcceptedCount.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct sfgpu_solve_params {
    pub max_nearby: u32,
    pub n_steps: u32,
    pub acceptor: i32,
    pub late_size: u32,
    pub tie_mode: i32,
    pub accepted_limit: u32,
    pub seed_base: u64,
    pub restore_best: i32,
    pub reserved: i32,
    pub acceptor_real: f64,
    pub step_count_limit: u64,
}

extern "C" {
    pub fn sfgpu_abi_version() -> i32;
    pub fn sfgpu_ctx_create(device: i32, flags: u64, cuda_stream: *mut c_void, out: *mut *mut sfgpu_ctx) -> i32;
    pub fn sfgpu_ctx_destroy(ctx: *mut sfgpu_ctx) -> i32;
    pub fn sfgpu_last_error(ctx: *const sfgpu_ctx) -> *const c_char;
    pub fn sfgpu_synchronize(ctx: *mut sfgpu_ctx) -> i32;

    pub fn sfgpu_model_begin(ctx: *mut sfgpu_ctx, n_replicas: u32) -> i32;
    pub fn sfgpu_add_collection(ctx: *mut sfgpu_ctx, name: *const c_char, n_rows: u32, descriptor_index: i32, out: *mut u32) -> i32;
    pub fn sfgpu_add_column_i64(ctx: *mut sfgpu_ctx, collection: u32, name: *const c_char, values: *const i64, out: *mut u32) -> i32;
    pub fn sfgpu_add_scalar_variable(ctx: *mut sfgpu_ctx, collection: u32, name: *const c_char, n_values: u32, allows_unassigned: i32, out: *mut u32) -> i32;
    pub fn sfgpu_add_list_variable(ctx: *mut sfgpu_ctx, owner_collection: u32, element_collection: u32, name: *const c_char, out: *mut u32) -> i32;
    pub fn sfgpu_add_csr(ctx: *mut sfgpu_ctx, name: *const c_char, n_rows: u32, row_ptr: *const u32, col_idx: *const u32, out: *mut u32) -> i32;
    pub fn sfgpu_add_matrix_i64(ctx: *mut sfgpu_ctx, name: *const c_char, rows: u32, cols: u32, values: *const i64, cost_semantics: i32, out: *mut u32) -> i32;
    pub fn sfgpu_add_expr(ctx: *mut sfgpu_ctx, ops: *const sfgpu_expr_op, n_ops: u32, out_expr: *mut u32) -> i32;
    pub fn sfgpu_add_constraint(ctx: *mut sfgpu_ctx, desc: *const sfgpu_constraint_desc, out: *mut u32) -> i32;
    pub fn sfgpu_set_scalar_state(ctx: *mut sfgpu_ctx, variable: u32, values: *const i32, per_replica: i32) -> i32;
    pub fn sfgpu_set_list_state(ctx: *mut sfgpu_ctx, variable: u32, offsets: *const u32, elems: *const u32, per_replica: i32) -> i32;
    pub fn sfgpu_model_commit(ctx: *mut sfgpu_ctx, out_scores: *mut i64) -> i32;

    pub fn sfgpu_score_change(ctx: *mut sfgpu_ctx, flags: u32, n: u64, cand_offsets: *const u64, rows: *const u32, out_scores: *mut i64, out_doable: *mut u8) -> i32;
    pub fn sfgpu_score_swap(ctx: *mut sfgpu_ctx, flags: u32, n: u64, cand_offsets: *const u64, rows: *const u32, out_scores: *mut i64, out_doable: *mut u8) -> i32;
    pub fn sfgpu_score_compound(ctx: *mut sfgpu_ctx, flags: u32, n: u64, cand_offsets: *const u64, edit_offsets: *const u64, edit_rows: *const u32, out_scores: *mut i64, out_doable: *mut u8) -> i32;
    pub fn sfgpu_score_list_change(ctx: *mut sfgpu_ctx, flags: u32, n: u64, cand_offsets: *const u64, rows: *const u32, out_scores: *mut i64, out_doable: *mut u8) -> i32;
    pub fn sfgpu_score_list_swap(ctx: *mut sfgpu_ctx, flags: u32, n: u64, cand_offsets: *const u64, rows: *const u32, out_scores: *mut i64, out_doable: *mut u8) -> i32;
    pub fn sfgpu_score_list_reverse(ctx: *mut sfgpu_ctx, flags: u32, n: u64, cand_offsets: *const u64, rows: *const u32, out_scores: *mut i64, out_doable: *mut u8) -> i32;
    /// rows[n][4] = {src_entity, start | size << 24, dst_entity, dst_position} (SublistChangeMove)
    pub fn sfgpu_score_sublist_change(ctx: *mut sfgpu_ctx, flags: u32, n: u64, cand_offsets: *const u64, rows: *const u32, out_scores: *mut i64, out_doable: *mut u8) -> i32;
    /// rows[n][4] = {first_entity, start1 | size1 << 24, second_entity, start2 | size2 << 24} (SublistSwapMove)
    pub fn sfgpu_score_k_opt(ctx: *mut sfgpu_ctx, flags: u32, n_candidates: u64, cand_offsets: *const u64, rows: *const u32, out_scores: *mut i64, out_doable: *mut u8) -> i32;
    pub fn sfgpu_apply_k_opt(ctx: *mut sfgpu_ctx, flags: u32, rows: *const u32, mask: *const u8) -> i32;
    pub fn sfgpu_score_sublist_swap(ctx: *mut sfgpu_ctx, flags: u32, n: u64, cand_offsets: *const u64, rows: *const u32, out_scores: *mut i64, out_doable: *mut u8) -> i32;

    pub fn sfgpu_argbest(ctx: *mut sfgpu_ctx, flags: u32, params: *const sfgpu_forage_params, cand_offsets: *const u64, scores: *const i64, doable: *const u8, step_seeds: *const u64, ref_scores: *const i64, out_index: *mut u32, out_best: *mut i64, out_evaluated: *mut u32) -> i32;
    pub fn sfgpu_argbest_gated(ctx: *mut sfgpu_ctx, flags: u32, params: *const sfgpu_forage_params, cand_offsets: *const u64, scores: *const i64, doable: *const u8, gates: *const u8, step_seeds: *const u64, ref_scores: *const i64, out_index: *mut u32, out_best: *mut i64, out_evaluated: *mut u32) -> i32;
    pub fn sfgpu_step_list_change(ctx: *mut sfgpu_ctx, n: u64, cand_offsets: *const u64, rows: *const u32, params: *const sfgpu_forage_params, step_seeds: *const u64, ref_scores: *const i64, out_scores: *mut i64, out_doable: *mut u8, out_index: *mut u32, out_best: *mut i64, out_evaluated: *mut u32) -> i32;
    pub fn sfgpu_step_change_rows(ctx: *mut sfgpu_ctx, n: u64, cand_offsets: *const u64, rows: *const u32, params: *const sfgpu_forage_params, step_seeds: *const u64, ref_scores: *const i64, out_scores: *mut i64, out_doable: *mut u8, out_index: *mut u32, out_best: *mut i64, out_evaluated: *mut u32) -> i32;
    pub fn sfgpu_step_nearby_list_change(ctx: *mut sfgpu_ctx, flags: u32, max_nearby: u32, params: *const sfgpu_forage_params, step_seeds: *const u64, ref_scores: *const i64, out_cand_offsets: *mut u64, out_rows: *mut u32, out_scores: *mut i64, out_doable: *mut u8, out_index: *mut u32, out_best: *mut i64, out_evaluated: *mut u32, out_winner_rows: *mut u32, apply_winners: i32) -> i32;
    pub fn sfgpu_step_nearby_list_swap(ctx: *mut sfgpu_ctx, flags: u32, max_nearby: u32, params: *const sfgpu_forage_params, step_seeds: *const u64, ref_scores: *const i64, out_cand_offsets: *mut u64, out_rows: *mut u32, out_scores: *mut i64, out_doable: *mut u8, out_index: *mut u32, out_best: *mut i64, out_evaluated: *mut u32, out_winner_rows: *mut u32, apply_winners: i32) -> i32;
    pub fn sfgpu_step_sublist_change(ctx: *mut sfgpu_ctx, flags: u32, min_size: u32, max_size: u32, params: *const sfgpu_forage_params, step_seeds: *const u64, ref_scores: *const i64, out_index: *mut u32, out_best: *mut i64, out_evaluated: *mut u32, out_winner_rows: *mut u32, apply_winners: i32) -> i32;
    pub fn sfgpu_step_list_reverse(ctx: *mut sfgpu_ctx, flags: u32, params: *const sfgpu_forage_params, step_seeds: *const u64, ref_scores: *const i64, out_index: *mut u32, out_best: *mut i64, out_evaluated: *mut u32, out_winner_rows: *mut u32, apply_winners: i32) -> i32;
    pub fn sfgpu_step_sublist_swap(ctx: *mut sfgpu_ctx, flags: u32, min_size: u32, max_size: u32, params: *const sfgpu_forage_params, step_seeds: *const u64, ref_scores: *const i64, out_index: *mut u32, out_best: *mut i64, out_evaluated: *mut u32, out_winner_rows: *mut u32, apply_winners: i32) -> i32;
    pub fn sfgpu_step_change(ctx: *mut sfgpu_ctx, flags: u32, params: *const sfgpu_forage_params, step_seeds: *const u64, ref_scores: *const i64, out_cand_offsets: *mut u64, out_rows: *mut u32, out_scores: *mut i64, out_doable: *mut u8, out_index: *mut u32, out_best: *mut i64, out_evaluated: *mut u32, out_winner_rows: *mut u32, apply_winners: i32) -> i32;
    pub fn sfgpu_step_union(ctx: *mut sfgpu_ctx, flags: u32, desc: *const sfgpu_union_desc, params: *const sfgpu_forage_params, step_seeds: *const u64, step_indices: *const u64, ref_scores: *const i64, out_index: *mut u32, out_best: *mut i64, out_evaluated: *mut u32, out_winner_rows: *mut u32, out_flags: *mut u32, apply_winners: i32) -> i32;
    pub fn sfgpu_solve_union(ctx: *mut sfgpu_ctx, desc: *const sfgpu_union_desc, params: *const sfgpu_solve_params, out_best_scores: *mut i64, out_moves_evaluated: *mut u64, out_accepted_steps: *mut u64, out_window_overflows: *mut u64, out_pulls_scored: *mut u64) -> i32;
    pub fn sfgpu_solve_nearby_list_change(ctx: *mut sfgpu_ctx, params: *const sfgpu_solve_params, out_best_scores: *mut i64, out_moves_evaluated: *mut u64, out_accepted_steps: *mut u64) -> i32;
    pub fn sfgpu_solve_change(ctx: *mut sfgpu_ctx, params: *const sfgpu_solve_params, out_best_scores: *mut i64, out_moves_evaluated: *mut u64, out_accepted_steps: *mut u64) -> i32;

    pub fn sfgpu_apply_change(ctx: *mut sfgpu_ctx, flags: u32, rows: *const u32, mask: *const u8) -> i32;
    pub fn sfgpu_apply_swap(ctx: *mut sfgpu_ctx, flags: u32, rows: *const u32, mask: *const u8) -> i32;
    pub fn sfgpu_apply_list_change(ctx: *mut sfgpu_ctx, flags: u32, rows: *const u32, mask: *const u8) -> i32;
    pub fn sfgpu_apply_list_swap(ctx: *mut sfgpu_ctx, flags: u32, rows: *const u32, mask: *const u8) -> i32;
    pub fn sfgpu_apply_list_reverse(ctx: *mut sfgpu_ctx, flags: u32, rows: *const u32, mask: *const u8) -> i32;
    pub fn sfgpu_apply_sublist_change(ctx: *mut sfgpu_ctx, flags: u32, rows: *const u32, mask: *const u8) -> i32;
    pub fn sfgpu_apply_sublist_swap(ctx: *mut sfgpu_ctx, flags: u32, rows: *const u32, mask: *const u8) -> i32;
    pub fn sfgpu_apply_winners(ctx: *mut sfgpu_ctx, move_kind: i32, cand_offsets: *const u64, batch_rows: *const u32, index: *const u32) -> i32;

    pub fn sfgpu_committed_scores(ctx: *mut sfgpu_ctx, out_scores: *mut i64) -> i32;
    pub fn sfgpu_evaluate_all(ctx: *mut sfgpu_ctx, out_scores: *mut i64) -> i32;
    pub fn sfgpu_get_scalar_state(ctx: *mut sfgpu_ctx, variable: u32, out_values: *mut i32) -> i32;
    pub fn sfgpu_get_list_state(ctx: *mut sfgpu_ctx, variable: u32, out_offsets: *mut u32, out_elems: *mut u32) -> i32;
    pub fn sfgpu_list_capacity(ctx: *mut sfgpu_ctx, variable: u32, out_capacity: *mut u32) -> i32;
    pub fn sfgpu_pack_best_keys(ctx: *mut sfgpu_ctx, out_keys_device: *mut i64) -> i32;
    pub fn sfgpu_comm_unique_id(out_id128: *mut u8) -> i32;
    pub fn sfgpu_comm_init_rank(n_ranks: i32, id128: *const u8, rank: i32, device: i32, out_comm: *mut *mut c_void) -> i32;
    pub fn sfgpu_comm_destroy(comm: *mut c_void) -> i32;
    pub fn sfgpu_sync_best(ctx: *mut sfgpu_ctx, nccl_comm: *mut c_void, flags: u32, scores_device: *const i64,
                           out_best: *mut i64, out_owner_rank: *mut i32, out_owner_replica: *mut u32) -> i32;
    pub fn sfgpu_last_kernel_ns(ctx: *mut sfgpu_ctx, out_ns: *mut u64) -> i32;
    pub fn sfgpu_kernel_times_ns(ctx: *mut sfgpu_ctx, max_n: u32, out_ns: *mut u64, out_n: *mut u32) -> i32;
    pub fn sfgpu_launch_count(ctx: *mut sfgpu_ctx, out_count: *mut u64) -> i32;
    pub fn sfgpu_scalar_program(ctx: *mut sfgpu_ctx, out_program: *mut i32) -> i32;
}
