"""Times the committed nearby ListChange step (sfgpu_step_nearby_list_change, apply_winners=1) at the bench workload
with the retained neighbourhood and with full regeneration (SFGPU_NO_NBCACHE=1). Device-timed, informational."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from solverforge_b200 import ForageParams, models


def run(R, steps, cached, acceptor):
    if cached:
        os.environ.pop("SFGPU_NO_NBCACHE", None)
    else:
        os.environ["SFGPU_NO_NBCACHE"] = "1"
    inst = bench.make_instance("cvrp")
    states = [bench.replica_start("cvrp", inst, bench.SEED_BASE + r) for r in range(R)]
    d = models.cvrp_director(inst, R, offsets=np.stack([s[0] for s in states]), elems=np.concatenate([s[1] for s in states]))
    seeds = np.arange(R, dtype=np.uint64)
    fp = ForageParams(acceptor, 1, 0)
    out = []
    last = d.calculate_score()
    ref = np.concatenate([last, last], axis=1)
    for _ in range(5):
        d.step_nearby_list_change(20, fp, step_seeds=seeds, ref_scores=ref, apply=True)
    d.synchronize()
    t0 = time.perf_counter()
    for s in range(steps):
        r = d.step_nearby_list_change(20, fp, step_seeds=seeds + s, ref_scores=ref, apply=True)
        out.append((r[0].copy(), r[1].copy()))
    d.synchronize()
    dt = (time.perf_counter() - t0) / steps
    kt = d.kernel_times_ns(steps)
    stats = None
    if cached:
        import ctypes as C
        tags = np.zeros((R, 16), dtype=np.uint32)
        fn = d.lib.sfgpu_debug_nearby_cache_tags
        fn.argtypes = [C.c_void_p, C.c_void_p]
        fn(d.h, tags.ctypes.data_as(C.c_void_p))
        t = tags[:, 10:13].astype(np.float64).sum(axis=0)
        stats = (t / t.sum()).round(3).tolist()
    d.close()
    return dt * 1e3, float(np.mean(kt)) / 1e6, out, stats


if __name__ == "__main__":
    R = int(os.environ.get("R", "1024"))
    steps = int(os.environ.get("STEPS", "40"))
    for acceptor in (0, 1):
        a = run(R, steps, True, acceptor)
        b = run(R, steps, False, acceptor)
        same = all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(a[2], b[2]))
        moved = float(np.mean([(x[0] != 0xFFFFFFFF).mean() for x in a[2]]))
        print(f"acceptor {acceptor}: retained {a[0]:.3f} ms/step (kernels {a[1]:.3f} ms)  regenerated {b[0]:.3f} ms/step "
              f"(kernels {b[1]:.3f} ms)  same winners: {same}  replicas moved per step: {moved:.2f}  tiers kept/re-scored/regenerated: {a[3]}")
