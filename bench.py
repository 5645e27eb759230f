#!/usr/bin/env python
"""bench.py — candidate moves scored per second on the batched re-score path.

One "step" = one pass of the hot path over one batch: for each of R independent seeded replicas
(restarts) on this GPU — every one with its own perturbed start — every candidate of the replica's
neighbourhood is re-scored against the replica's committed state and the winner is reduced on device
(BestScore forager + reservoir tie rule). Headline workload: CVRP-1000 / 80 vehicles, nearby list-change
selector, max_nearby = 20 -> 20 000 candidates per replica (BASELINE.json configs[2], the config the
north_star target is quoted on). The default invocation also measures, as sub-lines under "extra",
graph colouring 10k / 50k (configs[1]), job-shop 200 x 20 with the grouped complement (configs[3]) and the
single-solver latency case R = 1 — each with its own roofline, cpu_baseline and e2e.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, libsfgpu)
  python bench.py --impl reference ...                     CPU arm: the oracle's reference-faithful
                                                           incremental engine on all host cores

Under torchrun (N > 1) every rank owns one GPU and its own replicas (replica r of rank g runs seed
1000 + g*R + r; no data-path collective: the path partitions by replica, scaling = weak). The only exchange
is the best-score sync of SURVEY §8(e): every K = min(sync_every, steps) steps, inside the timed region and
timed on its own ("sync_ms").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate_moves_scored_per_s"
UNIT = "candidates/s"
ROW_BYTES = {"cvrp": 16, "graph_coloring": 8, "job_shop": 8}
OUT_BYTES = 17  # 16 B score + 1 B doable
DEFAULT_R = {"cvrp": 1024, "graph_coloring": 192, "job_shop": 200}   # R x batch >= 2^24
SEED_BASE = 1000  # replica r uses random_seed = 1000 + r


def workload_label(name: str) -> str:
    return {
        "cvrp": "solverforge-cvrp 1000 customers / 80 vehicles, nearby-list selector (max_nearby=20), "
                "ListChangeMove batch, HardSoftScore",
        "graph_coloring": "scalar-graph-coloring 10k vertices / 50k edges / k=8, full ChangeMove neighbourhood",
        "job_shop": "mixed-job-shop 200 jobs x 20 machines + grouped-complement load, full ChangeMove neighbourhood",
    }[name]


class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(name, R):
    """DRAM bytes per launch of the dominant kernel from the committed ncu captures (profiles/), or None
    when no capture exists for this workload / replica count."""
    for fn in ("r02_traffic.json", "r01_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", fn))).get(f"{name}:{R}")
            if t:
                return t["dram_bytes_read"] + t["dram_bytes_write"]
        except Exception:
            pass
    return None


def d_state_bytes(name, inst) -> int:
    if name == "cvrp":
        return (inst.n_routes + 1) * 4 + (inst.dim - 1) * 4 + 2 * inst.n_routes * 8 + 16
    if name == "graph_coloring":
        return inst.n * 4 + 16
    return inst.n_ops * 4 + (inst.n_ops // 20) * inst.n_machines * 4 + inst.n_machines * 12 + 16


def shared_bytes(name, inst) -> int:
    if name == "cvrp":
        return inst.dim * inst.dim * 4 + inst.dim * 8
    if name == "graph_coloring":
        return (inst.n + 1) * 4 + len(inst.col) * 4
    return inst.n_ops * 8


# ------------------------------------------------------------------------------------------ workloads
def replica_start(name, inst, seed):
    """Planning state of the replica with random_seed `seed` (SEED_BASE itself is the unperturbed instance)."""
    from solverforge_b200 import instances
    if name == "cvrp":
        return (inst.offsets, inst.elems) if seed == SEED_BASE else instances.perturb_routes(inst, seed, 64)
    if name == "graph_coloring":
        return inst.color if seed == SEED_BASE else instances.graph_coloring_colors(inst, seed)
    return inst.machine_idx if seed == SEED_BASE else instances.job_shop_machines(inst, seed)


def host_rows(name, inst, state):
    from solverforge_b200 import instances, selectors
    if name == "cvrp":
        return selectors.nearby_list_change_rows(state[0], state[1], inst.matrix, 20)
    k = inst.k if name == "graph_coloring" else inst.n_machines
    return instances.change_neighbourhood(state, k).astype(np.int64).astype(np.uint32)


def make_instance(name):
    from solverforge_b200 import instances
    return {"cvrp": instances.cvrp, "graph_coloring": instances.graph_coloring, "job_shop": instances.job_shop}[name]()


# ------------------------------------------------------------------------------------------ CPU
def cpu_reference_pass(name, inst, state, rows, n_threads, min_seconds, max_candidates=None, want_scores=False,
                       passes=None):
    """Times the oracle's reference-faithful engine (evaluate_candidate do/score/undo per candidate) over the
    replica's neighbourhood on n_threads host threads (one independent solver per thread, as the reference runs
    one solve per rayon job). Runs until min_seconds elapsed, or exactly `passes` passes per thread.
    Returns (candidates/s, sample description, seconds[, scores, doable])."""
    from tests.oracle_lib import Oracle
    if max_candidates:
        rows = rows[:max_candidates]

    def make():
        if name == "cvrp":
            return Oracle.cvrp(inst, *state)
        if name == "graph_coloring":
            return Oracle.graph_coloring(inst, state)
        return Oracle.job_shop(inst, state)

    oracles = [make() for _ in range(n_threads)]
    for o in oracles:
        o.committed_score()
    score = (lambda o: o.score_list_change(rows)) if name == "cvrp" else (
        lambda o: o.score_change(rows.astype(np.int64).astype(np.int32)))
    counts = [0] * n_threads
    stop_at = [0.0]

    def work(i):
        done = 0
        while True:
            score(oracles[i])
            counts[i] += len(rows)
            done += 1
            if passes is not None:
                if done >= passes:
                    break
            elif time.perf_counter() >= stop_at[0]:
                break

    first = score(oracles[0])  # warm-up
    t0 = time.perf_counter()
    stop_at[0] = t0 + min_seconds
    ths = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    total = sum(counts)
    sample = f"{total} candidates ({len(rows)} per pass) in {dt:.1f} s on {n_threads} thread(s)"
    if want_scores:
        return total / dt, sample, dt, first[0], first[1]
    return total / dt, sample, dt


def run_reference(args):
    """The reference arm: the oracle port of solverforge-scoring + the evaluate_candidate loop on every host core.
    One step = every host thread re-scores the replica's neighbourhood `passes` times (a bounded sample of the
    workload: the reference runs one single-threaded solver per core); ms_per_step is the measured time of a step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    inst = make_instance(name)
    state = replica_start(name, inst, SEED_BASE)
    rows = host_rows(name, inst, state)
    cores = os.cpu_count() or 1
    cap = 600 if name != "cvrp" else None  # the reference's predicate joins are O(n) per candidate
    # calibrate passes so a step lasts about a second
    v0, _, _ = cpu_reference_pass(name, inst, state, rows, cores, 0.5, cap)
    per_pass = (len(rows) if cap is None else min(cap, len(rows))) * cores
    budget_s = max(0.25, min(2.0, 100.0 / max(args.steps + args.warmup, 1)))
    passes = max(1, int(round(v0 * budget_s / per_pass)))
    vals, times, sample = [], [], ""
    for i in range(args.warmup + args.steps):
        v, sample, dt = cpu_reference_pass(name, inst, state, rows, cores, 0.0, cap, passes=passes)
        if i >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = float(sum(per_pass * passes for _ in times) / sum(times))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_label(name), "engine": "oracle port of solverforge-scoring "
                   "(retained incremental constraints, do/score/undo per candidate)",
                   "step": f"{passes} pass(es) over the neighbourhood on each of {cores} threads = {per_pass * passes} candidates"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + " per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU
class Dist:
    """torch.distributed plumbing of the run (NCCL under torchrun, nothing at N = 1) + the library's own NCCL
    communicator for sfgpu_sync_best."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.comm = None
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=self.dev)
            if os.environ.get("BENCH_ABI_SYNC", "1") == "1":
                self._init_abi_comm()

    def _init_abi_comm(self):
        """The C-ABI communicator a non-torch host would create (sfgpu_comm_*): rank 0's unique id travels over
        the existing process group."""
        from solverforge_b200 import _lib as L
        torch, lib = self.torch, L.load()
        idt = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            if lib.sfgpu_comm_unique_id(buf) != 0:
                idt[:] = 255  # marks "unavailable": every rank falls back to the torch collective
            else:
                idt = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        idd = idt.to(self.dev)
        self.dist.broadcast(idd, 0)
        idh = idd.cpu().numpy()
        if int(idh.min()) == 255:
            return
        comm = C.c_void_p()
        rc = lib.sfgpu_comm_init_rank(self.world, idh.ctypes.data_as(C.c_void_p), self.rank, self.local, C.byref(comm))
        ok = torch.tensor([1 if rc == 0 else 0], device=self.dev)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            self.comm = comm

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def close(self):
        if self.comm is not None:
            from solverforge_b200 import _lib as L
            L.load().sfgpu_comm_destroy(self.comm)
            self.comm = None
        if self.world > 1:
            self.dist.destroy_process_group()


def run_workload(D: Dist, name: str, R: int, steps: int, warmup: int, args, primary: bool):
    """Builds R replicas of `name` on this rank's GPU, gates them against the CPU checkers, times `steps` steps.
    Returns the measurement dict (rank 0 content is what gets printed)."""
    import torch
    from solverforge_b200 import ForageParams, models, _lib as L
    lib = L.load()
    dev, world, rank = D.dev, D.world, D.rank
    inst = make_instance(name)
    seeds = [SEED_BASE + rank * R + r for r in range(R)]
    states = [replica_start(name, inst, s) for s in seeds]
    stream = torch.cuda.Stream(device=dev)  # a real (non-legacy) stream shared by torch and libsfgpu
    torch.cuda.set_stream(stream)
    fp = ForageParams(acceptor=0, tie_mode=int(os.environ.get("BENCH_TIE_MODE", "1")), accepted_limit=0)
    t_seeds = torch.tensor(seeds, dtype=torch.int64, device=dev)
    t_idx = torch.empty(R, dtype=torch.int32, device=dev)
    t_best = torch.empty((R, 2), dtype=torch.int64, device=dev)
    t_eval = torch.empty(R, dtype=torch.int32, device=dev)
    words = ROW_BYTES[name] // 4
    if name == "cvrp":
        offs = np.stack([s[0] for s in states])
        elems = np.concatenate([s[1] for s in states])
        d = models.cvrp_director(inst, R, offsets=offs, elems=elems, device=D.local, stream=stream.cuda_stream)
        # candidate rows: generated ON DEVICE in the reference's pull order (NearbyListChangeMoveSelector) and
        # materialised once — every replica has its own rows for its own routes
        S = (inst.dim - 1) * 20
        n = R * S
        t_rows = torch.empty((n, 4), dtype=torch.int32, device=dev)
        t_offsets = torch.empty(R + 1, dtype=torch.int64, device=dev)
        t_win = torch.empty((R, 4), dtype=torch.int32, device=dev)
        d.step_nearby_list_change_device(20, fp, t_seeds.data_ptr(), 0, t_idx.data_ptr(), t_best.data_ptr(), t_eval.data_ptr(),
                                         t_win.data_ptr(), False, t_offsets.data_ptr(), t_rows.data_ptr())
        kind = "list_change"
    else:
        if name == "graph_coloring":
            d = models.graph_coloring_director(inst, R, colors=np.stack(states), device=D.local, stream=stream.cuda_stream)
        else:
            d = models.job_shop_director(inst, R, machine_idx=np.stack(states), device=D.local, stream=stream.cuda_stream)
        per = [host_rows(name, inst, s) for s in states]
        offsets = np.concatenate([[0], np.cumsum([len(x) for x in per])]).astype(np.uint64)
        n = int(offsets[-1])
        rows_pin = torch.empty((n, words), dtype=torch.int32).pin_memory()
        rows_pin.numpy().view(np.uint32)[:] = np.concatenate(per)
        t_rows = rows_pin.to(dev)
        t_offsets = torch.from_numpy(offsets.view(np.int64)).to(dev)
        kind = "change"
    torch.cuda.synchronize()
    t_scores = torch.empty((n, 2), dtype=torch.int64, device=dev)
    t_doable = torch.empty(n, dtype=torch.uint8, device=dev)
    use_fused = name == "cvrp"
    has_row_step = hasattr(d, "step_change_rows_device")

    def step(ev_pair=None):
        if ev_pair:
            ev_pair[0].record(stream)
        if use_fused:
            # one pass: score every candidate (scores + doable are materialised in HBM) and emit forager
            # partials; a finish kernel then completes the BestScore + tie-rule replay
            d.step_list_change_device(n, t_offsets.data_ptr(), t_rows.data_ptr(), fp, t_seeds.data_ptr(), 0,
                                      t_scores.data_ptr(), t_doable.data_ptr(), t_idx.data_ptr(), t_best.data_ptr(),
                                      t_eval.data_ptr())
        elif has_row_step:
            d.step_change_rows_device(n, t_offsets.data_ptr(), t_rows.data_ptr(), fp, t_seeds.data_ptr(), 0,
                                      t_scores.data_ptr(), t_doable.data_ptr(), t_idx.data_ptr(), t_best.data_ptr(),
                                      t_eval.data_ptr())
        else:
            d.score_device(kind, n, t_offsets.data_ptr(), t_rows.data_ptr(), t_scores.data_ptr(), t_doable.data_ptr())
            d.argbest_device(fp, t_offsets.data_ptr(), t_scores.data_ptr(), t_doable.data_ptr(), t_seeds.data_ptr(), 0,
                             t_idx.data_ptr(), t_best.data_ptr(), t_eval.data_ptr())
        if ev_pair:
            ev_pair[1].record(stream)

    # ---- parity gate: an incorrect kernel is never timed. Every candidate of EVERY replica of this rank
    # against the O(1) CPU checker (oracle/fast_cpu.cpp, pinned to the oracle by tests/test_oracle.py); the first
    # replicas also against the reference-faithful oracle itself (whose output doubles as the cpu_baseline sample)
    step()
    torch.cuda.synchronize()
    g_scores, g_doable = t_scores.cpu().numpy(), t_doable.cpu().numpy()
    g_rows = t_rows.cpu().numpy().view(np.uint32)
    h_offsets = t_offsets.cpu().numpy().astype(np.int64)
    from tests.oracle_lib import FastCvrp, FastGraphColoring, FastJobShop
    fast = {"cvrp": lambda: FastCvrp(inst), "graph_coloring": lambda: FastGraphColoring(inst),
            "job_shop": lambda: FastJobShop(inst)}[name]()
    for r in range(R):
        sl = slice(int(h_offsets[r]), int(h_offsets[r + 1]))
        if name == "cvrp":
            fast.set_routes(*states[r])
            sf, okf = fast.score(g_rows[sl])
        else:
            sf, okf, _ = fast.score_change(states[r], g_rows[sl].view(np.int32))
        if not (np.array_equal(g_scores[sl], sf) and np.array_equal(g_doable[sl], okf)):
            raise SystemExit(f"bench.py: GPU scores of replica {r} differ from the CPU checker — refusing to time an incorrect kernel")
    cpu_v = cpu_sample = cpu_o1 = None
    n_oracle = min(R, 3 if name == "cvrp" else 1)
    cap = 600 if name != "cvrp" else None
    for r in range(n_oracle):
        rows_r = host_rows(name, inst, states[r])
        sl = slice(int(h_offsets[r]), int(h_offsets[r]) + len(rows_r))
        if not np.array_equal(g_rows[sl], rows_r):
            raise SystemExit(f"bench.py: device-generated candidate rows of replica {r} differ from the host selector")
        secs = (args.cpu_seconds if primary else 4.0) if (r == 0 and rank == 0) else 0.0
        v, sample, _, so, oko = cpu_reference_pass(name, inst, states[r], rows_r, 1, secs, cap, want_scores=True)
        m = len(so)
        if not (np.array_equal(g_scores[sl][:m], so) and np.array_equal(g_doable[sl][:m], oko)):
            raise SystemExit(f"bench.py: GPU scores of replica {r} differ from the oracle — refusing to time an incorrect kernel")
        if r == 0 and rank == 0:
            cpu_v, cpu_sample = v, sample
    if name == "cvrp" and rank == 0 and primary:
        # a second, STRONGER CPU figure so the GPU/CPU ratio is not read off the reference's O(route) closures
        # alone: the same read-only O(1) delta as the GPU fast path, plain C++ (oracle/fast_cpu.cpp)
        fast.set_routes(*states[0])
        cores = os.cpu_count() or 1
        r0 = g_rows[int(h_offsets[0]):int(h_offsets[1])]
        cpu_o1 = {"one_thread": fast.bench(r0, 1, 2.0), "all_threads": fast.bench(r0, cores, 3.0), "cores": cores,
                  "unit": UNIT, "what": "O(1)-delta list-change scorer (GPU fast-path algorithm) in C++, one solver per thread"}
    del g_scores, g_doable

    # ---- best-score sync (N > 1): the only collective of the path, every K steps, inside the timed region
    K_sync = max(1, min(args.sync_every, steps))
    # the result of a sync stays on the device (SFGPU_SYNC_ASYNC): the stream is not drained, the next steps queue
    # right behind the collective; it is read back (and checked against the host-synchronous call) after the timed region
    t_sync = torch.zeros(4, dtype=torch.int64, device=dev)   # {hard, soft, owner rank (i32), owner replica (u32)}

    def sync():
        if D.comm is not None:   # the C ABI a non-torch host uses: device reduce + one ncclAllGather of 24 B / rank
            rc = lib.sfgpu_sync_best(d.h, D.comm, L.DEVICE_IO | L.SYNC_ASYNC, C.c_void_p(t_best.data_ptr()),
                                     C.c_void_p(t_sync.data_ptr()), C.cast(C.c_void_p(t_sync.data_ptr() + 16), C.POINTER(C.c_int32)),
                                     C.cast(C.c_void_p(t_sync.data_ptr() + 24), C.POINTER(C.c_uint32)))
            if rc != 0:
                raise SystemExit("sfgpu_sync_best: " + lib.sfgpu_last_error(d.h).decode())
            return None
        from solverforge_b200 import replicas
        return replicas.sync_best_scores(t_best, group=None)

    sampler = ClockSampler(D.local)
    if rank == 0 and primary:
        sampler.start()
    for _ in range(warmup):
        step()
    if world > 1 and primary:
        for _ in range(3):   # warm the collective (NCCL connects its channels lazily) — untimed
            sync()
    if primary:
        # untimed: keep the GPU under the same load long enough for nvidia-smi (100 ms period) to see it
        t_spin = time.perf_counter()
        while time.perf_counter() - t_spin < 0.6:
            for _ in range(8):
                step()
            torch.cuda.synchronize()
    launches0 = d.launch_count()
    D.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sev = []
    ev0.record(stream)
    for i in range(steps):
        step(kev[i])
        if world > 1 and primary and (i + 1) % K_sync == 0:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            sync()
            b.record(stream)
            sev.append((a, b))
    ev1.record(stream)
    torch.cuda.synchronize()
    D.barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    call_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))  # whole call: score kernel + finish kernel
    sync_ms = float(np.mean([a.elapsed_time(b) for a, b in sev])) if sev else None
    # dominant kernel alone: the library records an event pair around the scoring kernel of every call
    # on the launching stream; read back the ones that belong to the timed region
    n_sync_launches = len(sev)
    kt = d.kernel_times_ns(min(steps, 512))
    kernel_ms = float(np.mean(kt)) / 1e6 if len(kt) else call_ms
    launches = d.launch_count() - launches0
    clocks = sampler.stop() if (rank == 0 and primary) else None
    if world > 1 and primary and D.comm is not None and sev:
        # the device-side result of the last in-region sync == the host-synchronous form of the same call
        best_h = np.zeros(2, dtype=np.int64)
        owner, owner_rep = C.c_int32(), C.c_uint32()
        rc = lib.sfgpu_sync_best(d.h, D.comm, L.DEVICE_IO, C.c_void_p(t_best.data_ptr()), best_h.ctypes.data_as(C.c_void_p),
                                 C.byref(owner), C.byref(owner_rep))
        got = t_sync.cpu().numpy()
        if rc != 0 or [int(got[0]), int(got[1])] != best_h.tolist() or int(got[2]) & 0xFFFFFFFF != owner.value:
            raise SystemExit("bench.py: asynchronous best-score sync disagrees with the synchronous call")

    # ---- e2e_host_rows: the call a stock reference cursor would feed — candidate rows in pinned HOST memory in,
    # every score + doable flag back to pinned host memory (sfgpu_score_*; PCIe-bound)
    e2e_host_rows = None
    if primary or name != "cvrp":
        rows_pin = torch.empty((n, words), dtype=torch.int32).pin_memory()
        rows_pin.copy_(t_rows.cpu())
        sc_pin = torch.empty((n, 2), dtype=torch.int64).pin_memory()
        ok_pin = torch.empty(n, dtype=torch.uint8).pin_memory()
        offs_h = np.ascontiguousarray(h_offsets.astype(np.uint64))
        fn = lib.sfgpu_score_list_change if name == "cvrp" else lib.sfgpu_score_change

        def host_rows_call():
            rc = fn(d.h, 0, n, offs_h.ctypes.data_as(C.c_void_p), C.c_void_p(rows_pin.data_ptr()),
                    C.c_void_p(sc_pin.data_ptr()), C.c_void_p(ok_pin.data_ptr()))
            if rc != 0:
                raise SystemExit("host-rows score call failed: " + lib.sfgpu_last_error(d.h).decode())

        host_rows_call()
        if not np.array_equal(sc_pin.numpy(), t_scores.cpu().numpy()):
            raise SystemExit("bench.py: host-rows scores differ from the device-resident scores")
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            host_rows_call()
        hr_ms = (time.perf_counter() - t0) * 1e3 / reps
        e2e_host_rows = {"value": n / (hr_ms / 1e3), "unit": UNIT, "ms_per_step": hr_ms,
                         "h2d_bytes_per_step": n * ROW_BYTES[name] + (R + 1) * 8, "d2h_bytes_per_step": n * OUT_BYTES,
                         "api": ("sfgpu_score_list_change" if name == "cvrp" else "sfgpu_score_change") +
                                " (pinned host rows in, scores + doable out)"}
        del rows_pin, sc_pin, ok_pin

    # ---- e2e: the same metric through the reference-facing call with HOST buffers, copies inside the timed
    # region. The whole step runs on device (sfgpu_step_nearby_list_change / sfgpu_step_change generate the
    # neighbourhood, score it, replay the forager): the per-replica step seeds go in (pinned H2D) and the winners
    # come back (D2H). For the list workload every step also COMMITS its winner on device (apply_winners = 1), as a
    # solver step does, so the planning state of every replica changes between timed steps and the retained
    # neighbourhood has real work to do: sources the committed move touched are regenerated or
    # re-scored, the rest keep their score deltas. The same loop with the cache off (every step regenerates all
    # 20 000 candidates per replica) is reported next to it.
    seeds_host = np.array(seeds, dtype=np.uint64)
    e2e_extra = {}
    if name == "cvrp":
        e2e_api = ("sfgpu_step_nearby_list_change, apply_winners = 1 (device-side neighbourhood + score + forager + commit; "
                   "host seeds in, winners out)")
        h2d, d2h = R * 8, R * (4 + 16 + 4 + 16)
        e2e_step = lambda dd=None, s_off=0: (dd or d).step_nearby_list_change(20, fp, step_seeds=seeds_host + np.uint64(s_off), apply=True)
        first = d.step_nearby_list_change(20, fp, step_seeds=seeds_host)      # not committed: compared below
    else:
        e2e_api = "sfgpu_step_change (device-side ChangeMove neighbourhood + score + forager; host seeds in, winners out)"
        h2d, d2h = R * 8, R * (4 + 16 + 4 + 8)
        e2e_step = lambda dd=None, s_off=0: d.step_change(fp, step_seeds=seeds_host)
        first = e2e_step()
    idx_e2e, best_e2e, ev_e2e, _ = first
    # same winners as the rows-resident path (replica starts, seeds and forager are identical)
    if not (np.array_equal(idx_e2e, t_idx.cpu().numpy().view(np.uint32)) and
            np.array_equal(best_e2e, t_best.cpu().numpy()) and int(ev_e2e.sum()) == n):
        raise SystemExit("bench.py: device-generated step disagrees with the rows-resident step")
    e2e_steps = max(5, min(steps, 50))
    d_full = None
    if name == "cvrp":
        # the regenerating twin: same starts, cache off (read at commit time); it must commit the same winners
        os.environ["SFGPU_NO_NBCACHE"] = "1"
        try:
            d_full = models.cvrp_director(inst, R, offsets=offs, elems=elems, device=D.local, stream=stream.cuda_stream)
        finally:
            os.environ.pop("SFGPU_NO_NBCACHE", None)
        for w in range(3):   # warm-up: the first cached step builds the retained rows
            g, f_ = e2e_step(d, 7000 + w), e2e_step(d_full, 7000 + w)
            if not all(np.array_equal(x, y) for x, y in zip(g, f_)):
                raise SystemExit("bench.py: retained-neighbourhood step disagrees with the fully regenerated step")
    torch.cuda.synchronize()
    D.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(d, i)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if d_full is not None:
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            last_full = e2e_step(d_full, i)
        torch.cuda.synchronize()
        full_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        if not np.array_equal(d.calculate_score(), d_full.calculate_score()):
            raise SystemExit("bench.py: retained-neighbourhood trajectory left the fully regenerated one")
        tags = np.zeros((R, 16), dtype=np.uint32)
        lib.sfgpu_debug_nearby_cache_tags.argtypes = [C.c_void_p, C.c_void_p]
        lib.sfgpu_debug_nearby_cache_tags(d.h, tags.ctypes.data_as(C.c_void_p))
        tiers = tags[:, 10:13].astype(np.float64).sum(axis=0)
        e2e_extra = {"committed_every_step": True,
                     "retained_neighbourhood": {"sources_kept": tiers[0] / tiers.sum(), "sources_rescored": tiers[1] / tiers.sum(),
                                                "sources_regenerated": tiers[2] / tiers.sum(),
                                                "note": "fractions since commit, incl. the first step that regenerates everything"},
                     "full_regeneration": {"ms_per_step": full_ms, "value": n * world / (full_ms / 1e3),
                                           "note": "same committed loop, SFGPU_NO_NBCACHE=1: identical winners and final scores"}}
        d_full.close()

    elapsed_ms, kernel_ms, e2e_ms, sync_max = D.max_([elapsed_ms, kernel_ms, e2e_ms, sync_ms or 0.0])
    total_cands = n * world
    peak, peak_src = measured_peak_gbs()
    alg_bytes = n * (ROW_BYTES[name] + OUT_BYTES) + R * d_state_bytes(name, inst) + shared_bytes(name, inst)
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    kernel_name = d.dominant_kernel_name(kind) if hasattr(d, "dominant_kernel_name") else (
        "score_list_change_fast_kernel (scores + forager partials; forage_finish_kernel excluded)" if use_fused
        else f"score kernel of sfgpu_score_{kind}")
    res = {
        "value": total_cands * steps / (elapsed_ms / 1e3), "unit": UNIT, "steps": steps, "warmup": warmup,
        "ms_per_step": elapsed_ms / steps,
        "config": {"workload": workload_label(name), "replicas_per_gpu": R, "distinct_starts": R,
                   "replica_seeds": f"{SEED_BASE} + rank * {R} + r (every replica its own perturbed start)",
                   "candidates_per_step_per_gpu": n,
                   "forager": "BestScore + reservoir ties, replayed on device (fused partials + finish kernel)",
                   "l2": f"inputs larger than L2 ({n * (ROW_BYTES[name] + OUT_BYTES) / 1e6:.0f} MB per step)" if n * 33 > 126e6
                         else f"{n * (ROW_BYTES[name] + OUT_BYTES) / 1e6:.1f} MB per step (L2-resident: single-solver latency case)",
                   "sync_every": K_sync if world > 1 else None,
                   "sync_api": ("sfgpu_sync_best (C ABI, own ncclComm: device reduce + ncclAllGather of 24 B / rank, result left on the device)" if D.comm is not None
                                else "replicas.sync_best_scores (torch.distributed all_gather)") if world > 1 else None,
                   "parity_gate": f"every candidate of all {R} replicas bit-identical to the O(1) CPU checker; "
                                  f"replicas 0..{n_oracle - 1} bit-identical to the oracle (rows and scores) before timing"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(name, R), "peak_source": peak_src, "kernel": kernel_name,
                     "kernel_ms": kernel_ms, "call_ms": call_ms, "algorithmic_bytes_per_launch": alg_bytes},
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port", "sample": cpu_sample},
        "e2e": dict({"value": total_cands / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                     "ms_per_step": e2e_ms, "api": e2e_api}, **e2e_extra),
        "e2e_host_rows": e2e_host_rows,
        "gpu_launches": int(launches),
    }
    if world > 1:
        res["sync"] = {"syncs_in_timed_region": n_sync_launches, "sync_ms": sync_max, "every": K_sync}
    if cpu_o1:
        res["cpu_baseline_o1_delta"] = cpu_o1
    if clocks:
        res["clocks"] = clocks
    if name == "graph_coloring" and not primary and args.loop_steps > 0:
        # the reference's DEFAULT search of a plain scalar model on device (sfgpu_solve_union): union[ChangeMoveSelector,
        # SwapMoveSelector] with seeded Random leaves, StratifiedRandom, SimulatedAnnealing + AcceptedCount(1)
        # (default_local_search/policy.rs:48-81) — a step ends at its first accepted pull, so this is steps/s, not
        # candidates/s: informational, outside every timed region
        try:
            desc = d.default_scalar_union(window=8)
            loop_steps = max(args.loop_steps, 256)
            d.solve_union(desc, 16, 6, 0, 1, 1, seed_base=500)
            d.synchronize()
            t0 = time.perf_counter()
            best_u, ev_u, acc_u, ovf_u = d.solve_union(desc, loop_steps, 6, 0, 1, 1, seed_base=1000)
            dt = time.perf_counter() - t0
            res["default_search"] = {
                "steps": loop_steps, "replicas": R, "ms_per_step": dt * 1e3 / loop_steps, "solver_steps_per_s": R * loop_steps / dt,
                "moves_evaluated_per_s": float(ev_u.sum()) / dt, "committed_steps": int(acc_u.sum()),
                "scored_over_evaluated": float(d.last_pulls_scored.sum()) / max(float(ev_u.sum()), 1.0),
                "window_overflows": int(ovf_u.sum()),
                "selector": "union[ChangeMoveSelector, SwapMoveSelector], Random leaves, StratifiedRandom",
                "acceptor": "SimulatedAnnealing(calibrated, decay 0.999985)", "forager": "AcceptedCount(1)"}
        except Exception as exc:  # informational only
            res["default_search"] = {"error": str(exc)[:200]}
    if primary:
        res["_director"], res["_inst"], res["_states"] = d, inst, states
    else:
        d.close()
    return res


def run_ours(args):
    import torch
    from solverforge_b200 import ForageParams, models
    D = Dist()
    name = args.workload
    R = args.replicas or DEFAULT_R[name]
    main = run_workload(D, name, R, args.steps, args.warmup, args, primary=True)
    d, inst, states = main.pop("_director"), main.pop("_inst"), main.pop("_states")

    # informational: the device-resident loop (reference defaults for list models: LateAcceptance(400) +
    # AcceptedCount(256), default_local_search/policy.rs:18-82) — whole steps incl. commit, no host round trip
    device_loop = None
    if name == "cvrp" and args.loop_steps > 0:
        # the committed e2e loop above moved every replica: the loops below start from the replicas' own starts again
        d.close()
        d = models.cvrp_director(inst, R, offsets=np.stack([s[0] for s in states]), elems=np.concatenate([s[1] for s in states]),
                                 device=D.local)
        d.solve_nearby_list_change(16, 20, 2, 400, 1, 256, seed_base=500)   # warm-up: buffers, graph capture
        d.synchronize()
        t0 = time.perf_counter()
        best_l, ev_l, acc_l = d.solve_nearby_list_change(args.loop_steps, 20, 2, 400, 1, 256, seed_base=1000)
        dt = time.perf_counter() - t0
        device_loop = {"steps": args.loop_steps, "replicas": R, "ms_per_step": dt * 1e3 / args.loop_steps,
                       "moves_evaluated_per_s": float(ev_l.sum()) / dt, "committed_steps": int(acc_l.sum()),
                       "selector": "NearbyListChange(20), SelectionOrder::Original, whole neighbourhood scored per step (retained between steps: only sources the committed move touched are regenerated)",
                       "acceptor": "LateAcceptance(400)", "forager": "AcceptedCount(256)",
                       "best_score_replica0": [int(best_l[0][0]), int(best_l[0][1])]}

    # the reference's DEFAULT list local search on device (sfgpu_solve_union): seeded Random leaves, StratifiedRandom
    # union of nearby change / nearby swap / sublist change / sublist swap / reverse, LateAcceptance(400) +
    # AcceptedCount(256); only a window of the union stream is generated and scored per step
    default_search = None
    if name == "cvrp" and args.loop_steps > 0:
        try:
            desc = d.default_list_union()
            d.solve_union(desc, 16, 2, 400, 1, 256, seed_base=500)     # warm-up: buffers, graph capture
            d.synchronize()
            t0 = time.perf_counter()
            best_u, ev_u, acc_u, ovf_u = d.solve_union(desc, args.loop_steps, 2, 400, 1, 256, seed_base=1000)
            dt = time.perf_counter() - t0
            pulls = float(d.last_pulls_scored.sum())
            default_search = {"steps": args.loop_steps, "replicas": R, "ms_per_step": dt * 1e3 / args.loop_steps,
                              "moves_evaluated_per_s": float(ev_u.sum()) / dt, "pulls_scored_per_s": pulls / dt,
                              "scored_over_evaluated": pulls / max(float(ev_u.sum()), 1.0),
                              "committed_steps": int(acc_u.sum()), "window_overflows": int(ovf_u.sum()),
                              "selector": "union[NearbyListChange(20), NearbyListSwap(20), SublistChange(1..=3), "
                                          "SublistSwap(1..=3), ListReverse], Random leaves, StratifiedRandom",
                              "acceptor": "LateAcceptance(400)", "forager": "AcceptedCount(256)",
                              "windows_per_child": "adaptive per replica (first 64), x4, then 4096"}
        except Exception as exc:  # informational only
            default_search = {"error": str(exc)[:200]}

    # informational, outside every timed region: the device-enumerated sublist neighbourhoods of the reference's
    # default list policy (SublistChange / SublistSwap, sizes 1..=3, ~3 M / ~3.8 M candidates per replica and
    # step, never materialised) — whole steps incl. commit through the host call, on a small replica count
    sublist_steps = None
    if name == "cvrp" and D.world == 1 and args.loop_steps > 0:
        try:
            Rs = min(R, 32)
            ds = models.cvrp_director(inst, Rs, offsets=np.stack([states[r][0] for r in range(Rs)]),
                                      elems=np.concatenate([states[r][1] for r in range(Rs)]), device=D.local)
            sublist_steps = {"replicas": Rs, "sizes": "1..=3"}
            for label, fn in (("sublist_change", ds.step_sublist_change), ("sublist_swap", ds.step_sublist_swap)):
                last = ds.calculate_score()
                ref = np.concatenate([last, last], axis=1)
                fn(1, 3, ForageParams(1, 1, 0), step_seeds=list(range(Rs)), ref_scores=ref)   # warm-up
                t0 = time.perf_counter()
                tot = 0
                for s_i in range(3):
                    idx_s, best_s, ev_s, win_s = fn(1, 3, ForageParams(1, 1, 0), step_seeds=[7 * s_i + r for r in range(Rs)],
                                                    ref_scores=ref, apply=True)
                    tot += int(ev_s.astype(np.int64).sum())
                dt = time.perf_counter() - t0
                sublist_steps[label] = {"candidates_per_s": tot / dt, "ms_per_step": dt * 1e3 / 3,
                                        "candidates_per_replica_step": tot // (3 * Rs)}
            del ds
        except Exception as exc:  # informational only: never fail the bench line
            sublist_steps = {"error": str(exc)[:200]}
    d.close()

    # sub-lines: the other single-GPU configs of BASELINE.json and the R = 1 latency case, measured
    # the same way (own parity gate, roofline, cpu_baseline, e2e). One GPU only: they are not part of the scaling run.
    extra = {}
    if D.world == 1 and not args.no_extra and name == "cvrp" and not args.replicas:
        sub_steps = max(5, min(args.steps, 50))
        for key, wl, r_ in (("c2", "graph_coloring", DEFAULT_R["graph_coloring"]), ("c4", "job_shop", DEFAULT_R["job_shop"]),
                            ("r1", "cvrp", 1)):
            try:
                sub = run_workload(D, wl, r_, sub_steps if key != "r1" else max(sub_steps, 50), args.warmup, args, primary=False)
                sub.update({"metric": METRIC, "n_gpus": 1, "dtype": "int64", "data": "synthetic"})
                if key == "r1":
                    sub["us_per_step"] = sub["ms_per_step"] * 1e3
                extra[key] = sub
            except SystemExit as exc:   # a failed gate of a sub-line must be visible, not fatal for the headline
                extra[key] = {"error": str(exc)[:300]}

    if D.rank == 0:
        line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic"}
        for k in ("config", "roofline", "cpu_baseline", "cpu_baseline_o1_delta", "e2e", "e2e_host_rows", "gpu_launches",
                  "clocks", "sync"):
            if k in main:
                line[k] = main[k]
        line["device_loop"] = device_loop
        line["default_search"] = default_search
        line["sublist_steps"] = sublist_steps
        line["extra"] = extra
        print(json.dumps(line))
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cvrp", choices=["cvrp", "graph_coloring", "job_shop"])
    ap.add_argument("--replicas", type=int, default=0,
                    help="independent seeded replicas per GPU per launch (0 = SURVEY §8d: 1024 / 192 / 200)")
    ap.add_argument("--loop-steps", type=int, default=512, help="steps of the device-resident loop demo (0 = skip)")
    ap.add_argument("--sync-every", type=int, default=64,
                    help="steps between best-score syncs (N > 1); SURVEY 8(d) C5: K = 64 (clamped to --steps)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="length of the cpu_baseline sample")
    ap.add_argument("--no-extra", action="store_true", help="skip the C2 / C4 / R=1 sub-lines")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
