#!/usr/bin/env python
"""bench.py — candidate moves scored per second on the batched re-score path.

One "step" = one pass of the hot path over one batch: for each of R independent seeded replicas
(restarts) on this GPU, every candidate of the replica's nearby-list-change neighbourhood
(CVRP-1000 / 80 vehicles, max_nearby = 20 -> 20 000 candidates per replica) is re-scored against
the replica's committed state and the winner is reduced on device (BestScore forager + tie rule).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, libsfgpu)
  python bench.py --impl reference ...                     CPU arm: the oracle's reference-faithful
                                                           incremental engine on all host cores

Under torchrun (N > 1) every rank owns one GPU and its own replicas (no data-path collective:
the path partitions by replica, scaling = weak); the only exchange is one 8-byte MAX all-reduce
of the best packed score per sync (NCCL).
"""
from __future__ import annotations

import argparse
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


def build_workload(name: str, distinct: int):
    """Returns (instance, list of (state, rows)) for `distinct` different replica starts."""
    from solverforge_b200 import instances, selectors
    starts = []
    if name == "cvrp":
        inst = instances.cvrp()
        for i in range(distinct):
            offs, el = (inst.offsets, inst.elems) if i == 0 else instances.perturb_routes(inst, 1000 + i, 64)
            rows = selectors.nearby_list_change_rows(offs, el, inst.matrix, 20)
            starts.append(((offs, el), rows))
    elif name == "graph_coloring":
        inst = instances.graph_coloring()
        for i in range(distinct):
            col = inst.color if i == 0 else instances.graph_coloring(seed_colors=43 + i).color
            starts.append((col, instances.change_neighbourhood(col, inst.k).astype(np.int64).astype(np.uint32)))
    elif name == "job_shop":
        inst = instances.job_shop()
        for i in range(distinct):
            m = inst.machine_idx if i == 0 else instances.job_shop(seed=11 + i).machine_idx
            starts.append((m, instances.change_neighbourhood(m, inst.n_machines).astype(np.int64).astype(np.uint32)))
    else:
        raise SystemExit(f"unknown workload {name}")
    return inst, starts


def workload_label(name: str) -> str:
    return {
        "cvrp": "solverforge-cvrp 1000 customers / 80 vehicles, nearby-list selector (max_nearby=20), "
                "ListChangeMove batch, HardSoftScore",
        "graph_coloring": "scalar-graph-coloring 10k vertices / 50k edges / k=8, full ChangeMove neighbourhood",
        "job_shop": "mixed-job-shop 200 jobs x 20 machines + grouped-complement load, full ChangeMove neighbourhood",
    }[name]


def algorithmic_bytes_per_candidate(name: str) -> float:
    # every candidate row read once + every result written once (+1 B doable);
    # replica state and shared facts are counted once per launch (added separately).
    return ROW_BYTES[name] + OUT_BYTES


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


# ------------------------------------------------------------------------------------------ CPU
def cpu_reference_pass(name, inst, start, n_threads, min_seconds, max_candidates=None, want_scores=False):
    """Times the oracle's reference-faithful engine (evaluate_candidate do/score/undo per candidate)
    over the replica's neighbourhood on n_threads host threads (one independent solver per thread,
    as the reference runs one solve per rayon job). Returns (candidates/s, sample description)."""
    from tests.oracle_lib import Oracle
    state, rows = start
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
        while True:
            score(oracles[i])
            counts[i] += len(rows)
            if time.perf_counter() >= stop_at[0]:
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
        return total / dt, sample, first[0], first[1]
    return total / dt, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inst, starts = build_workload(args.workload, 1)
    cores = os.cpu_count() or 1
    cap = 600 if args.workload != "cvrp" else None  # the reference's predicate joins are O(n) per candidate
    per_step = max(1.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
    vals = []
    for i in range(args.warmup + args.steps):
        v, sample = cpu_reference_pass(args.workload, inst, starts[0], cores, per_step, cap)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": workload_label(args.workload), "engine": "oracle port of solverforge-scoring "
                   "(retained incremental constraints, do/score/undo per candidate)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU
def run_ours(args):
    import torch
    import torch.distributed as dist
    from solverforge_b200 import ForageParams, models

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name, R = args.workload, args.replicas
    inst, starts = build_workload(name, min(args.distinct, R))
    D = len(starts)
    words = ROW_BYTES[name] // 4
    # replica r starts from start (r % D): its own state block and its own candidate rows in HBM
    counts = np.array([len(starts[r % D][1]) for r in range(R)], dtype=np.uint64)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    n = int(offsets[-1])
    rows_host = torch.empty((n, words), dtype=torch.int32).pin_memory()
    rows_np = rows_host.numpy().view(np.uint32)
    for r in range(R):
        rows_np[int(offsets[r]):int(offsets[r + 1])] = starts[r % D][1]
    stream = torch.cuda.Stream(device=dev)  # a real (non-legacy) stream shared by torch and libsfgpu
    torch.cuda.set_stream(stream)
    if name == "cvrp":
        offs = np.stack([starts[r % D][0][0] for r in range(R)])
        elems = np.concatenate([starts[r % D][0][1] for r in range(R)])
        d = models.cvrp_director(inst, R, offsets=offs, elems=elems, device=local, stream=stream.cuda_stream)
        kind, move_kind = "list_change", 2
    elif name == "graph_coloring":
        d = models.graph_coloring_director(inst, R, colors=np.stack([starts[r % D][0] for r in range(R)]),
                                           device=local, stream=stream.cuda_stream)
        kind, move_kind = "change", 0
    else:
        d = models.job_shop_director(inst, R, machine_idx=np.stack([starts[r % D][0] for r in range(R)]),
                                     device=local, stream=stream.cuda_stream)
        kind, move_kind = "change", 0

    t_offsets = torch.from_numpy(offsets.view(np.int64)).to(dev)
    t_rows = rows_host.to(dev)
    t_scores = torch.empty((n, 2), dtype=torch.int64, device=dev)
    t_doable = torch.empty(n, dtype=torch.uint8, device=dev)
    t_seeds = torch.arange(R, dtype=torch.int64, device=dev)
    t_idx = torch.empty(R, dtype=torch.int32, device=dev)
    t_best = torch.empty((R, 2), dtype=torch.int64, device=dev)
    t_eval = torch.empty(R, dtype=torch.int32, device=dev)
    t_keys = torch.empty(R, dtype=torch.int64, device=dev)
    fp = ForageParams(acceptor=0, tie_mode=int(os.environ.get("BENCH_TIE_MODE", "1")), accepted_limit=0)

    use_fused = name == "cvrp"

    def step(i, ev_pair=None):
        if ev_pair:
            ev_pair[0].record(stream)
        if use_fused:
            # one pass: score every candidate (scores + doable are materialised in HBM) and emit forager
            # partials; a one-warp-per-replica kernel then finishes the BestScore + tie-rule replay
            d.step_list_change_device(n, t_offsets.data_ptr(), t_rows.data_ptr(), fp, t_seeds.data_ptr(), 0,
                                      t_scores.data_ptr(), t_doable.data_ptr(), t_idx.data_ptr(), t_best.data_ptr(),
                                      t_eval.data_ptr())
            if ev_pair:
                ev_pair[1].record(stream)
        else:
            d.score_device(kind, n, t_offsets.data_ptr(), t_rows.data_ptr(), t_scores.data_ptr(), t_doable.data_ptr())
            if ev_pair:
                ev_pair[1].record(stream)
            d.argbest_device(fp, t_offsets.data_ptr(), t_scores.data_ptr(), t_doable.data_ptr(), t_seeds.data_ptr(), 0,
                             t_idx.data_ptr(), t_best.data_ptr(), t_eval.data_ptr())
        if world > 1 and (i + 1) % args.sync_every == 0:
            # best packed score over this rank's replicas, then one 8-byte MAX all-reduce (NCCL):
            # SURVEY 8(e) — the only collective of the path, every K steps
            dbg = os.environ.get("BENCH_DEBUG_SYNC")
            if dbg:
                import time as _t
                torch.cuda.synchronize()
                h0 = _t.perf_counter()
            best = (((t_best[:, 0] + (1 << 22)) << 40) | (t_best[:, 1] + (1 << 39))).max().reshape(1)
            if dbg:
                torch.cuda.synchronize()
                h1 = _t.perf_counter()
            dist.all_reduce(best, op=dist.ReduceOp.MAX)
            if dbg:
                h2 = _t.perf_counter()
                torch.cuda.synchronize()
                h3 = _t.perf_counter()
                print(f"[rank {rank}] sync at step {i}: pack {1e3 * (h1 - h0):.3f} ms, all_reduce call {1e3 * (h2 - h1):.3f} ms, "
                      f"drain {1e3 * (h3 - h2):.3f} ms", file=sys.stderr, flush=True)

    # cpu_baseline leg (rank 0): the oracle scores replica 0's batch on one host core; its output
    # doubles as the parity gate — an incorrect kernel is never timed.
    step(0)
    torch.cuda.synchronize()
    cpu_v = cpu_sample = cpu_o1 = None
    if rank == 0:
        cap = 600 if name != "cvrp" else None
        cpu_v, cpu_sample, so, oko = cpu_reference_pass(name, inst, starts[0], 1, 12.0, cap, want_scores=True)
        r0 = slice(0, len(so))
        if not (np.array_equal(t_scores[r0].cpu().numpy(), so) and np.array_equal(t_doable[r0].cpu().numpy(), oko)):
            raise SystemExit("bench.py: GPU scores differ from the oracle — refusing to time an incorrect kernel")
        if name == "cvrp":
            # a second, STRONGER CPU figure so the GPU/CPU ratio is not read off the reference's O(route)
            # closures alone: the same read-only O(1) delta as the GPU fast path, plain C++ (oracle/fast_cpu.cpp)
            from tests.oracle_lib import FastCvrp
            fc = FastCvrp(inst, *starts[0][0])
            cores = os.cpu_count() or 1
            cpu_o1 = {"one_thread": fc.bench(starts[0][1], 1, 2.0), "all_threads": fc.bench(starts[0][1], cores, 3.0),
                      "cores": cores, "unit": UNIT,
                      "what": "O(1)-delta list-change scorer (GPU fast-path algorithm) in C++, one solver per thread"}

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    # untimed: keep the GPU under the same load long enough for nvidia-smi (100 ms period) to see it
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.6:
        for i in range(8):
            step(i)
        torch.cuda.synchronize()
    launches0 = d.launch_count()
    if world > 1:
        # warm the collective itself (NCCL connects its channels lazily on the first all-reduce of a shape):
        # with a sync only every K steps no warm-up step would reach it otherwise
        # ... and torch loads the few elementwise / reduce kernels of the key packing lazily (tens of ms on
        # first use): run the exact sync expression untimed
        for _ in range(3):
            warm = (((t_best[:, 0] + (1 << 22)) << 40) | (t_best[:, 1] + (1 << 39))).max().reshape(1)
            dist.all_reduce(warm, op=dist.ReduceOp.MAX)
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record(stream)
    for i in range(args.steps):
        step(i, kev[i])
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    call_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))  # whole call: score kernel + finish kernel
    # dominant kernel alone: the library records an event pair around the scoring kernel of every call
    # on the launching stream; read back the ones that belong to the timed region
    kt = d.kernel_times_ns(min(args.steps, 512))
    kernel_ms = float(np.mean(kt)) / 1e6 if len(kt) else call_ms
    launches = d.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # e2e: the same metric through the reference-facing call with HOST buffers, copies inside the timed
    # region. CVRP: the whole step runs on device (sfgpu_step_nearby_list_change generates the nearby
    # neighbourhood, scores it, replays the forager) so only the per-replica step seeds go in (pinned
    # H2D) and the winners come back (D2H). Other workloads: sfgpu_score_* with pinned host rows in and
    # every score out.
    import ctypes as C
    if name == "cvrp":
        seeds_host = np.arange(R, dtype=np.uint64)
        e2e_api = "sfgpu_step_nearby_list_change (device-side neighbourhood + score + forager; host seeds in, winners out)"
        h2d, d2h = R * 8, R * (4 + 16 + 4 + 16)

        def e2e_step():
            return d.step_nearby_list_change(20, fp, step_seeds=seeds_host)

        idx_e2e, best_e2e, ev_e2e, win_e2e = e2e_step()
        # same winners as the rows-resident path (replica starts, seeds and forager are identical)
        if not (np.array_equal(idx_e2e, t_idx.cpu().numpy().view(np.uint32)) and
                np.array_equal(best_e2e, t_best.cpu().numpy()) and int(ev_e2e.sum()) == n):
            raise SystemExit("bench.py: device-generated step disagrees with the rows-resident step")
        e2e_steps = max(5, min(args.steps, 50))
    else:
        seeds_host = np.arange(R, dtype=np.uint64)
        e2e_api = "sfgpu_step_change (device-side ChangeMove neighbourhood + score + forager; host seeds in, winners out)"
        h2d, d2h = R * 8, R * (4 + 16 + 4 + 8)

        def e2e_step():
            return d.step_change(fp, step_seeds=seeds_host)

        idx_e2e, best_e2e, ev_e2e, win_e2e = e2e_step()
        if not (np.array_equal(idx_e2e, t_idx.cpu().numpy().view(np.uint32)) and
                np.array_equal(best_e2e, t_best.cpu().numpy()) and int(ev_e2e.sum()) == n):
            raise SystemExit("bench.py: device-generated step disagrees with the rows-resident step")
        e2e_steps = max(5, min(args.steps, 50))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps

    # informational: the device-resident loop (reference defaults for list models: LateAcceptance(400) +
    # AcceptedCount(256), default_local_search/policy.rs:18-82) — whole steps incl. commit, no host round trip
    device_loop = None
    if name == "cvrp" and args.loop_steps > 0:
        d.synchronize()
        t0 = time.perf_counter()
        best_l, ev_l, acc_l = d.solve_nearby_list_change(args.loop_steps, 20, 2, 400, 1, 256, seed_base=1000)
        dt = time.perf_counter() - t0
        device_loop = {"steps": args.loop_steps, "replicas": R, "ms_per_step": dt * 1e3 / args.loop_steps,
                       "moves_evaluated_per_s": float(ev_l.sum()) / dt, "committed_steps": int(acc_l.sum()),
                       "acceptor": "LateAcceptance(400)", "forager": "AcceptedCount(256)",
                       "best_score_replica0": [int(best_l[0][0]), int(best_l[0][1])]}

    # informational, outside every timed region: the device-enumerated sublist neighbourhoods of the reference's
    # default list policy (SublistChange / SublistSwap, sizes 1..=3, ~3 M / ~3.8 M candidates per replica and
    # step, never materialised) — whole steps incl. commit through the host call, on a small replica count
    sublist_steps = None
    if name == "cvrp" and world == 1 and args.loop_steps > 0:
        try:
            Rs = min(R, 32)
            ds = models.cvrp_director(inst, Rs, offsets=np.stack([starts[r % D][0][0] for r in range(Rs)]),
                                      elems=np.concatenate([starts[r % D][0][1] for r in range(Rs)]), device=local)
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

    t = torch.tensor([elapsed_ms, kernel_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, kernel_ms, e2e_ms = (float(x) for x in t.cpu())
    total_cands = n * world
    value = total_cands * args.steps / (elapsed_ms / 1e3)
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        state_bytes = R * d_state_bytes(name, inst)
        shared = shared_bytes(name, inst)
        alg_bytes = n * algorithmic_bytes_per_candidate(name) + state_bytes + shared
        achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": workload_label(name), "replicas_per_gpu": R, "distinct_starts": D,
                       "candidates_per_step_per_gpu": n, "forager": "BestScore + reservoir ties, replayed on device (fused partials + finish kernel)",
                       "l2": f"inputs larger than L2 ({n * (ROW_BYTES[name] + OUT_BYTES) / 1e6:.0f} MB per step)",
                       "sync_every": args.sync_every if world > 1 else None,
                       "parity_gate": "replica 0 bit-identical to the oracle before timing"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(name, R), "peak_source": peak_src,
                         "kernel": "score_list_change_fast_kernel (scores + forager partials; forage_finish_kernel excluded)" if use_fused
                         else f"score_{kind}_kernel", "kernel_ms": kernel_ms, "call_ms": call_ms,
                         "algorithmic_bytes_per_launch": alg_bytes},
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port", "sample": cpu_sample},
            "cpu_baseline_o1_delta": cpu_o1,
            "e2e": {"value": total_cands / (e2e_ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "api": e2e_api},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "device_loop": device_loop,
            "sublist_steps": sublist_steps,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measured_traffic(name, R):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None
    when no capture exists for this workload / replica count."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(f"{name}:{R}")
        return t["dram_bytes_read"] + t["dram_bytes_write"] if t else None
    except Exception:
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cvrp", choices=["cvrp", "graph_coloring", "job_shop"])
    ap.add_argument("--replicas", type=int, default=1024, help="independent seeded replicas per GPU per launch")
    ap.add_argument("--distinct", type=int, default=16, help="distinct replica starts (tiled over the replicas)")
    ap.add_argument("--loop-steps", type=int, default=64, help="steps of the device-resident loop demo (0 = skip)")
    ap.add_argument("--sync-every", type=int, default=64,
                    help="steps between NCCL best-score syncs (N > 1); SURVEY 8(d) C5: K = 64")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
