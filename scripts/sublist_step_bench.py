"""Throughput of the device-enumerated SublistChange step on CVRP-1000 / 80 (sizes 1..=3, ~3.2 M candidates per
replica and step, nothing materialised). Device-timed through the context's scoring-kernel events + wall clock
of the whole host call. Usage: python scripts/sublist_step_bench.py [replicas] [steps] [change|swap]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from solverforge_b200 import ForageParams, instances, models  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
swap = len(sys.argv) > 3 and sys.argv[3] == "swap"
c = instances.cvrp()
starts = [instances.perturb_routes(c, 60 + r % 16, 200) for r in range(R)]
d = models.cvrp_director(c, R, offsets=np.stack([s[0] for s in starts]), elems=np.concatenate([s[1] for s in starts]))
last = d.calculate_score()
ref = np.concatenate([last, last], axis=1)
seeds = list(range(R))
step = d.step_sublist_swap if swap else d.step_sublist_change
step(1, 3, ForageParams(1, 1, 0), step_seeds=seeds, ref_scores=ref)  # warm-up
t0 = time.perf_counter()
tot = 0
for s in range(steps):
    idx, best, ev, win = step(1, 3, ForageParams(1, 1, 0), step_seeds=[x + s for x in seeds], ref_scores=ref, apply=True)
    tot += int(ev.astype(np.int64).sum())
    last = d.calculate_score()
    ref = np.concatenate([last, last], axis=1)
dt = time.perf_counter() - t0
print({"neighbourhood": "sublist_swap" if swap else "sublist_change", "replicas": R, "steps": steps, "candidates": tot, "wall_s": round(dt, 4), "candidates_per_s": tot / dt,
       "ms_per_step": 1e3 * dt / steps, "best0": last[0].tolist()})
