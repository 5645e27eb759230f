"""Times sfgpu_step_change (device-generated ChangeMove neighbourhood + score + forager; host seeds in, winners out) at
the bench's C2 / C4 workloads. Informational; run under ncu for the per-kernel split."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from solverforge_b200 import ForageParams, models

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "job_shop"
    steps = int(os.environ.get("STEPS", "20"))
    R = bench.DEFAULT_R[name]
    inst = bench.make_instance(name)
    states = [bench.replica_start(name, inst, bench.SEED_BASE + r) for r in range(R)]
    if name == "graph_coloring":
        d = models.graph_coloring_director(inst, R, colors=np.stack(states))
    else:
        d = models.job_shop_director(inst, R, machine_idx=np.stack(states))
    seeds = np.arange(R, dtype=np.uint64)
    fp = ForageParams(0, 1, 0)
    for _ in range(3):
        d.step_change(fp, step_seeds=seeds)
    d.synchronize()
    t0 = time.perf_counter()
    for s in range(steps):
        d.step_change(fp, step_seeds=seeds)
    d.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f"{name}: {dt * 1e3:.3f} ms/step, kernels {float(np.mean(d.kernel_times_ns(steps))) / 1e6:.3f} ms")
