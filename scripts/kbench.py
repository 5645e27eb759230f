#!/usr/bin/env python
"""Kernel-tuning microbenchmark (no parity gate — bench.py and the tests own correctness): times the dominant
kernel of one workload's rows-resident fused step. usage: kbench.py WORKLOAD [R] [steps]; knobs through the
SFGPU_* environment variables the library reads."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench as B  # noqa: E402
from solverforge_b200 import ForageParams, models  # noqa: E402

name = sys.argv[1]
R = int(sys.argv[2]) if len(sys.argv) > 2 else B.DEFAULT_R[name]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
mat = os.environ.get("KB_MATERIALISE", "1") == "1"
inst = B.make_instance(name)
seeds = [B.SEED_BASE + r for r in range(R)]
states = [B.replica_start(name, inst, s) for s in seeds]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
fp = ForageParams(0, 1, 0)
t_seeds = torch.tensor(seeds, dtype=torch.int64, device=dev)
t_idx = torch.empty(R, dtype=torch.int32, device=dev)
t_best = torch.empty((R, 2), dtype=torch.int64, device=dev)
t_eval = torch.empty(R, dtype=torch.int32, device=dev)
if name == "cvrp":
    d = models.cvrp_director(inst, R, offsets=np.stack([s[0] for s in states]), elems=np.concatenate([s[1] for s in states]),
                             stream=stream.cuda_stream)
    n = R * (inst.dim - 1) * 20
    t_rows = torch.empty((n, 4), dtype=torch.int32, device=dev)
    t_offsets = torch.empty(R + 1, dtype=torch.int64, device=dev)
    t_win = torch.empty((R, 4), dtype=torch.int32, device=dev)
    d.step_nearby_list_change_device(20, fp, t_seeds.data_ptr(), 0, t_idx.data_ptr(), t_best.data_ptr(), t_eval.data_ptr(),
                                     t_win.data_ptr(), False, t_offsets.data_ptr(), t_rows.data_ptr())
    fn = d.step_list_change_device
else:
    if name == "graph_coloring":
        d = models.graph_coloring_director(inst, R, colors=np.stack(states), stream=stream.cuda_stream)
    else:
        d = models.job_shop_director(inst, R, machine_idx=np.stack(states), stream=stream.cuda_stream)
    per = [B.host_rows(name, inst, s) for s in states]
    offsets = np.concatenate([[0], np.cumsum([len(x) for x in per])]).astype(np.uint64)
    n = int(offsets[-1])
    t_rows = torch.from_numpy(np.concatenate(per).view(np.int32)).to(dev)
    t_offsets = torch.from_numpy(offsets.view(np.int64)).to(dev)
    fn = d.step_change_rows_device
t_scores = torch.empty((n, 2), dtype=torch.int64, device=dev)
t_doable = torch.empty(n, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()


def step():
    fn(n, t_offsets.data_ptr(), t_rows.data_ptr(), fp, t_seeds.data_ptr(), 0, t_scores.data_ptr() if mat else 0,
       t_doable.data_ptr() if mat else 0, t_idx.data_ptr(), t_best.data_ptr(), t_eval.data_ptr())


for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps):
    step()
e1.record(stream)
torch.cuda.synchronize()
kt = d.kernel_times_ns(min(steps, 512))
alg = n * (B.ROW_BYTES[name] + B.OUT_BYTES) + R * B.d_state_bytes(name, inst) + B.shared_bytes(name, inst)
kms = float(np.mean(kt)) / 1e6
peak, _ = B.measured_peak_gbs()
print(f"{name} R={R} n={n} step_ms={e0.elapsed_time(e1) / steps:.4f} kernel_ms={kms:.4f} "
      f"frac={alg / (kms / 1e3) / 1e9 / peak:.3f} env={ {k: v for k, v in os.environ.items() if k.startswith(('SFGPU_', 'KB_'))} }")
