"""Times the device-resident default scalar search (union[Change, Swap], Random, StratifiedRandom, SimulatedAnnealing,
AcceptedCount(1)) on the bench's graph-colouring replicas. usage: union_bench_scalar.py [R] [steps] [window]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from solverforge_b200 import GpuScoreDirector, models  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
window = int(sys.argv[3]) if len(sys.argv) > 3 else 8
inst = B.make_instance("graph_coloring")
states = [B.replica_start("graph_coloring", inst, B.SEED_BASE + r) for r in range(R)]
d = models.graph_coloring_director(inst, R, colors=np.stack(states))
desc = GpuScoreDirector.default_scalar_union(window=window)
d.solve_union(desc, 16, 6, 0, 1, 1, seed_base=500)
d.synchronize()
t0 = time.perf_counter()
best, ev, acc, ovf = d.solve_union(desc, steps, 6, 0, 1, 1, seed_base=1000)
dt = time.perf_counter() - t0
print(f"R={R} steps={steps} window={window} ms_per_step={dt * 1e3 / steps:.3f} steps/s={R * steps / dt:.3e} "
      f"evaluated/replica-step={ev.sum() / (R * steps):.2f} overflows={int(ovf.sum())}")
