//! `solverforge-gpu`: plugs libsfgpu (B200, sm_100a) into SolverForge's local search.
//!
//! * [`GpuScoreDirector`] implements `solverforge_scoring::Director<S>`
//!   (solverforge-scoring/src/director/traits.rs:27-95) over one `sfgpu_ctx`: the committed score lives on the
//!   device, committed moves go through `sfgpu_apply_*`, `fresh_score` is `sfgpu_evaluate_all` (FullAssert).
//! * [`BatchDirector`] is the seam the reference lacks: `evaluate_candidates`
//!   (solverforge-solver/src/phase/localsearch/phase/candidates.rs:47-285) scores one candidate at a time; a
//!   batch of drained cursor rows is scored by one kernel launch and replayed in pull order.
//! * [`GpuScoreDirector::step_union`] / [`GpuScoreDirector::solve_union`] run whole steps / whole phases of
//!   the default list local search on the device (seeded leaves + StratifiedRandom union).
//!
//! Errors never unwind across the C boundary and there is no CPU fallback: `SFGPU_E_UNSUPPORTED` means the
//! model is not expressible on device and the caller keeps the stock `ScoreDirector`
//! (solverforge-solver/src/run.rs:552-557).
#![allow(clippy::too_many_arguments)]

use std::ffi::{CStr, CString};
use std::ptr;

use solverforge_core::domain::{PlanningSolution, SolutionDescriptor};
use solverforge_core::score::HardSoftScore;
use solverforge_gpu_sys as sys;
use solverforge_scoring::{ConstraintMetadata, Director, DirectorScoreState};

/// `SFGPU_E_*` with the library's message (`sfgpu_last_error`).
#[derive(Debug, Clone)]
pub struct GpuError {
    pub code: i32,
    pub message: String,
}

impl std::fmt::Display for GpuError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "libsfgpu error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for GpuError {}

pub type GpuResult<T> = Result<T, GpuError>;

/// One device context = one solve (R replicas). `Send`, not `Sync`, like `Director`.
pub struct GpuContext {
    raw: *mut sys::sfgpu_ctx,
    replicas: u32,
}
unsafe impl Send for GpuContext {}

impl GpuContext {
    pub fn new(device: i32, replicas: u32) -> GpuResult<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::sfgpu_ctx_create(device, 0, ptr::null_mut(), &mut raw) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sys::sfgpu_last_error(ptr::null())) };
            return Err(GpuError { code: rc, message: msg.to_string_lossy().into_owned() });
        }
        let ctx = Self { raw, replicas };
        ctx.check(unsafe { sys::sfgpu_model_begin(raw, replicas) })?;
        Ok(ctx)
    }

    fn check(&self, rc: i32) -> GpuResult<()> {
        if rc == 0 {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(sys::sfgpu_last_error(self.raw)) };
        Err(GpuError { code: rc, message: msg.to_string_lossy().into_owned() })
    }

    pub fn add_collection(&self, name: &str, rows: u32, descriptor_index: i32) -> GpuResult<u32> {
        let (n, mut out) = (CString::new(name).unwrap(), 0u32);
        self.check(unsafe { sys::sfgpu_add_collection(self.raw, n.as_ptr(), rows, descriptor_index, &mut out) })?;
        Ok(out)
    }
    pub fn add_column(&self, collection: u32, name: &str, values: &[i64]) -> GpuResult<u32> {
        let (n, mut out) = (CString::new(name).unwrap(), 0u32);
        self.check(unsafe { sys::sfgpu_add_column_i64(self.raw, collection, n.as_ptr(), values.as_ptr(), &mut out) })?;
        Ok(out)
    }
    pub fn add_scalar_variable(&self, collection: u32, name: &str, n_values: u32, allows_unassigned: bool) -> GpuResult<u32> {
        let (n, mut out) = (CString::new(name).unwrap(), 0u32);
        self.check(unsafe {
            sys::sfgpu_add_scalar_variable(self.raw, collection, n.as_ptr(), n_values, allows_unassigned as i32, &mut out)
        })?;
        Ok(out)
    }
    pub fn add_list_variable(&self, owners: u32, elements: u32, name: &str) -> GpuResult<u32> {
        let (n, mut out) = (CString::new(name).unwrap(), 0u32);
        self.check(unsafe { sys::sfgpu_add_list_variable(self.raw, owners, elements, n.as_ptr(), &mut out) })?;
        Ok(out)
    }
    pub fn add_csr(&self, name: &str, row_ptr: &[u32], col_idx: &[u32]) -> GpuResult<u32> {
        let (n, mut out) = (CString::new(name).unwrap(), 0u32);
        self.check(unsafe {
            sys::sfgpu_add_csr(self.raw, n.as_ptr(), row_ptr.len() as u32 - 1, row_ptr.as_ptr(), col_idx.as_ptr(), &mut out)
        })?;
        Ok(out)
    }
    pub fn add_matrix(&self, name: &str, rows: u32, cols: u32, values: &[i64], cost_semantics: bool) -> GpuResult<u32> {
        let (n, mut out) = (CString::new(name).unwrap(), 0u32);
        self.check(unsafe {
            sys::sfgpu_add_matrix_i64(self.raw, n.as_ptr(), rows, cols, values.as_ptr(), cost_semantics as i32, &mut out)
        })?;
        Ok(out)
    }
    /// Pair filter / pair weight of a cross-collection join as a postfix column expression
    /// (the closures of stream/filter/adapters.rs:63-93 cannot run on a GPU).
    pub fn add_expr(&self, ops: &[sys::sfgpu_expr_op]) -> GpuResult<u32> {
        let mut out = 0u32;
        self.check(unsafe { sys::sfgpu_add_expr(self.raw, ops.as_ptr(), ops.len() as u32, &mut out) })?;
        Ok(out)
    }
    pub fn add_constraint(&self, desc: &sys::sfgpu_constraint_desc) -> GpuResult<u32> {
        let mut out = 0u32;
        self.check(unsafe { sys::sfgpu_add_constraint(self.raw, desc, &mut out) })?;
        Ok(out)
    }
    pub fn set_scalar_state(&self, values: &[i32], per_replica: bool) -> GpuResult<()> {
        self.check(unsafe { sys::sfgpu_set_scalar_state(self.raw, 0, values.as_ptr(), per_replica as i32) })
    }
    pub fn set_list_state(&self, offsets: &[u32], elems: &[u32], per_replica: bool) -> GpuResult<()> {
        self.check(unsafe {
            sys::sfgpu_set_list_state(self.raw, 0x8000_0000, offsets.as_ptr(), elems.as_ptr(), per_replica as i32)
        })
    }
    /// `ConstraintSet::initialize_all` on every replica; returns the committed scores.
    pub fn commit(&self) -> GpuResult<Vec<HardSoftScore>> {
        let mut out = vec![0i64; 2 * self.replicas as usize];
        self.check(unsafe { sys::sfgpu_model_commit(self.raw, out.as_mut_ptr()) })?;
        Ok(out.chunks(2).map(|c| HardSoftScore::of(c[0], c[1])).collect())
    }
}

impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { sys::sfgpu_ctx_destroy(self.raw) };
    }
}

/// Move kinds of the packed SoA rows (CandidateId = row index, move_selector/borrowed.rs:396-430).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum BatchKind {
    Change,
    Swap,
    ListChange,
    ListSwap,
    ListReverse,
    SublistChange,
    SublistSwap,
    KOpt,
}

/// A drained cursor (or a chunk of it): `words` u32 per candidate, replica r owns rows
/// `offsets[r]..offsets[r + 1]`.
pub struct CandidateBatch<'a> {
    pub kind: BatchKind,
    pub rows: &'a [u32],
    pub offsets: &'a [u64],
}

/// The batch seam: one launch scores every pull of the batch; the caller replays acceptor + forager in
/// pull order (or asks the device to, [`GpuScoreDirector::argbest`]).
pub trait BatchDirector<S: PlanningSolution>: Director<S> {
    fn score_candidates(&mut self, batch: &CandidateBatch<'_>) -> GpuResult<(Vec<HardSoftScore>, Vec<u8>)>;
}

pub struct GpuScoreDirector<S: PlanningSolution> {
    ctx: GpuContext,
    working: S,
    descriptor: SolutionDescriptor,
    /// packs the winning move of the working solution into a device row after `after_variable_changed`
    /// (model-specific: which field is the planning variable)
    row_of: Box<dyn Fn(&S, usize, usize) -> (BatchKind, [u32; 4]) + Send>,
}

impl<S: PlanningSolution<Score = HardSoftScore>> GpuScoreDirector<S> {
    pub fn new(ctx: GpuContext, working: S, descriptor: SolutionDescriptor,
               row_of: Box<dyn Fn(&S, usize, usize) -> (BatchKind, [u32; 4]) + Send>) -> Self {
        Self { ctx, working, descriptor, row_of }
    }

    /// Acceptor + forager replay on the device over materialised scores (`BestCandidate::consider`,
    /// forager.rs:99-155; acceptor predicates 0..3 of `sfgpu_forage_params`).
    pub fn argbest(&mut self, params: &sys::sfgpu_forage_params, offsets: &[u64], scores: &[i64], doable: &[u8],
                   step_seeds: &[u64], ref_scores: &[i64]) -> GpuResult<(Vec<u32>, Vec<i64>, Vec<u32>)> {
        let r = self.ctx.replicas as usize;
        let (mut idx, mut best, mut ev) = (vec![0u32; r], vec![0i64; 2 * r], vec![0u32; r]);
        self.ctx.check(unsafe {
            sys::sfgpu_argbest(self.ctx.raw, 0, params, offsets.as_ptr(), scores.as_ptr(), doable.as_ptr(), step_seeds.as_ptr(),
                               ref_scores.as_ptr(), idx.as_mut_ptr(), best.as_mut_ptr(), ev.as_mut_ptr())
        })?;
        Ok((idx, best, ev))
    }

    /// One step of the reference's default list local search on the device: the union cursor in its seeded
    /// pull order, scored, replayed and (optionally) committed. Returns (CandidateId, score, moves_evaluated,
    /// winner rows [R][8], flags).
    pub fn step_union(&mut self, desc: &sys::sfgpu_union_desc, params: &sys::sfgpu_forage_params, step_seeds: &[u64],
                      step_indices: &[u64], ref_scores: &[i64], apply: bool)
                      -> GpuResult<(Vec<u32>, Vec<i64>, Vec<u32>, Vec<u32>, Vec<u32>)> {
        let r = self.ctx.replicas as usize;
        let (mut idx, mut best, mut ev) = (vec![0u32; r], vec![0i64; 2 * r], vec![0u32; r]);
        let (mut win, mut flags) = (vec![0u32; 8 * r], vec![0u32; r]);
        self.ctx.check(unsafe {
            sys::sfgpu_step_union(self.ctx.raw, 0, desc, params, step_seeds.as_ptr(), step_indices.as_ptr(), ref_scores.as_ptr(),
                                  idx.as_mut_ptr(), best.as_mut_ptr(), ev.as_mut_ptr(), win.as_mut_ptr(), flags.as_mut_ptr(),
                                  apply as i32)
        })?;
        Ok((idx, best, ev, win, flags))
    }

    /// The whole local-search phase for every replica without a host round trip
    /// (solve_local_search_with_resources, phase/localsearch/phase.rs:237-320).
    pub fn solve_union(&mut self, desc: &sys::sfgpu_union_desc, params: &sys::sfgpu_solve_params)
                       -> GpuResult<(Vec<HardSoftScore>, Vec<u64>, Vec<u64>)> {
        let r = self.ctx.replicas as usize;
        let (mut best, mut ev, mut steps) = (vec![0i64; 2 * r], vec![0u64; r], vec![0u64; r]);
        self.ctx.check(unsafe {
            sys::sfgpu_solve_union(self.ctx.raw, desc, params, best.as_mut_ptr(), ev.as_mut_ptr(), steps.as_mut_ptr(),
                                   ptr::null_mut(), ptr::null_mut())
        })?;
        Ok((best.chunks(2).map(|c| HardSoftScore::of(c[0], c[1])).collect(), ev, steps))
    }

    /// Best score over every replica of every rank (SolverManager jobs -> GPUs, manager.rs:22,93-146):
    /// one device reduction + one ncclAllGather of 24 B per rank.
    pub fn sync_best(&mut self, nccl_comm: *mut std::ffi::c_void) -> GpuResult<(HardSoftScore, i32, u32)> {
        let (mut best, mut rank, mut replica) = ([0i64; 2], 0i32, 0u32);
        self.ctx.check(unsafe {
            sys::sfgpu_sync_best(self.ctx.raw, nccl_comm, 0, ptr::null(), best.as_mut_ptr(), &mut rank, &mut replica)
        })?;
        Ok((HardSoftScore::of(best[0], best[1]), rank, replica))
    }
}

impl<S: PlanningSolution<Score = HardSoftScore>> BatchDirector<S> for GpuScoreDirector<S> {
    fn score_candidates(&mut self, batch: &CandidateBatch<'_>) -> GpuResult<(Vec<HardSoftScore>, Vec<u8>)> {
        let n = *batch.offsets.last().unwrap_or(&0);
        let (mut scores, mut doable) = (vec![0i64; 2 * n as usize], vec![0u8; n as usize]);
        let (c, o, r, s, d) = (self.ctx.raw, batch.offsets.as_ptr(), batch.rows.as_ptr(), scores.as_mut_ptr(), doable.as_mut_ptr());
        let rc = unsafe {
            match batch.kind {
                BatchKind::Change => sys::sfgpu_score_change(c, 0, n, o, r, s, d),
                BatchKind::Swap => sys::sfgpu_score_swap(c, 0, n, o, r, s, d),
                BatchKind::ListChange => sys::sfgpu_score_list_change(c, 0, n, o, r, s, d),
                BatchKind::ListSwap => sys::sfgpu_score_list_swap(c, 0, n, o, r, s, d),
                BatchKind::ListReverse => sys::sfgpu_score_list_reverse(c, 0, n, o, r, s, d),
                BatchKind::SublistChange => sys::sfgpu_score_sublist_change(c, 0, n, o, r, s, d),
                BatchKind::SublistSwap => sys::sfgpu_score_sublist_swap(c, 0, n, o, r, s, d),
                BatchKind::KOpt => sys::sfgpu_score_k_opt(c, 0, n, o, r, s, d),
            }
        };
        self.ctx.check(rc)?;
        Ok((scores.chunks(2).map(|c| HardSoftScore::of(c[0], c[1])).collect(), doable))
    }
}

impl<S: PlanningSolution<Score = HardSoftScore> + Clone + Send> Director<S> for GpuScoreDirector<S> {
    fn working_solution(&self) -> &S {
        &self.working
    }
    fn working_solution_mut(&mut self) -> &mut S {
        &mut self.working
    }
    /// incremental.rs:141-149: the committed score of replica 0
    fn calculate_score(&mut self) -> HardSoftScore {
        let mut out = vec![0i64; 2 * self.ctx.replicas as usize];
        let rc = unsafe { sys::sfgpu_committed_scores(self.ctx.raw, out.as_mut_ptr()) };
        self.ctx.check(rc).expect("sfgpu_committed_scores");
        let score = HardSoftScore::of(out[0], out[1]);
        self.working.set_score(Some(score));
        score
    }
    /// scope_core.rs:642-653 (FullAssert): stateless recompute on a scratch copy
    fn fresh_score(&self) -> Option<HardSoftScore> {
        let mut out = vec![0i64; 2 * self.ctx.replicas as usize];
        let rc = unsafe { sys::sfgpu_evaluate_all(self.ctx.raw, out.as_mut_ptr()) };
        self.ctx.check(rc).ok()?;
        Some(HardSoftScore::of(out[0], out[1]))
    }
    fn solution_descriptor(&self) -> &SolutionDescriptor {
        &self.descriptor
    }
    fn clone_working_solution(&self) -> S {
        self.working.clone()
    }
    /// retract happens on the device together with the insert: committed moves arrive as ONE apply call
    fn before_variable_changed(&mut self, _descriptor_index: usize, _entity_index: usize) {}
    fn after_variable_changed(&mut self, descriptor_index: usize, entity_index: usize) {
        let (kind, row) = (self.row_of)(&self.working, descriptor_index, entity_index);
        let (c, r) = (self.ctx.raw, row.as_ptr());
        let rc = unsafe {
            match kind {
                BatchKind::Change => sys::sfgpu_apply_change(c, 0, r, ptr::null()),
                BatchKind::Swap => sys::sfgpu_apply_swap(c, 0, r, ptr::null()),
                BatchKind::ListChange => sys::sfgpu_apply_list_change(c, 0, r, ptr::null()),
                BatchKind::ListSwap => sys::sfgpu_apply_list_swap(c, 0, r, ptr::null()),
                BatchKind::ListReverse => sys::sfgpu_apply_list_reverse(c, 0, r, ptr::null()),
                BatchKind::SublistChange => sys::sfgpu_apply_sublist_change(c, 0, r, ptr::null()),
                BatchKind::SublistSwap => sys::sfgpu_apply_sublist_swap(c, 0, r, ptr::null()),
                BatchKind::KOpt => sys::sfgpu_apply_k_opt(c, 0, r, ptr::null()),
            }
        };
        self.ctx.check(rc).expect("sfgpu_apply_*");
    }
    fn entity_count(&self, descriptor_index: usize) -> Option<usize> {
        self.descriptor.entity_count(&self.working, descriptor_index)
    }
    fn total_entity_count(&self) -> Option<usize> {
        self.descriptor.total_entity_count(&self.working)
    }
    fn constraint_metadata(&self) -> Vec<ConstraintMetadata<'_>> {
        Vec::new()
    }
    fn is_incremental(&self) -> bool {
        true
    }
    fn snapshot_score_state(&self) -> DirectorScoreState<HardSoftScore> {
        let s = self.working.score();
        DirectorScoreState { solution_score: s, committed_score: s, initialized: s.is_some() }
    }
    fn restore_score_state(&mut self, state: DirectorScoreState<HardSoftScore>) {
        self.working.set_score(state.solution_score);
    }
}
