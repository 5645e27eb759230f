"""Times the device-resident default list local search (sfgpu_solve_union) on the bench's CVRP-1000 replicas.
usage: union_bench.py [R] [steps] [window] [max_window]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from solverforge_b200 import GpuScoreDirector, models  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
window = int(sys.argv[3]) if len(sys.argv) > 3 else 0
max_window = int(sys.argv[4]) if len(sys.argv) > 4 else 0
inst = B.make_instance("cvrp")
states = [B.replica_start("cvrp", inst, B.SEED_BASE + r) for r in range(R)]
d = models.cvrp_director(inst, R, offsets=np.stack([s[0] for s in states]), elems=np.concatenate([s[1] for s in states]))
desc = GpuScoreDirector.default_list_union(20, window, max_window)
d.solve_union(desc, 16, 2, 400, 1, 256, seed_base=500)
d.synchronize()
t0 = time.perf_counter()
best, ev, acc, ovf = d.solve_union(desc, steps, 2, 400, 1, 256, seed_base=1000)
dt = time.perf_counter() - t0
pulls = float(d.last_pulls_scored.sum())
print(f"R={R} steps={steps} window={window} ms_per_step={dt * 1e3 / steps:.3f} evaluated/s={ev.sum() / dt:.3e} "
      f"evaluated/replica-step={ev.sum() / (R * steps):.1f} scored/evaluated={pulls / max(ev.sum(), 1):.2f} overflows={int(ovf.sum())}")
