#!/usr/bin/env python
"""CVRP-1000 / 80 vehicles solved by R seeded replicas entirely on one B200: the model is authored with
the ConstraintFactory mirror, the local-search loop (nearby list-change neighbourhood, LateAcceptance(400)
+ AcceptedCount(256) — the reference's defaults for list models) runs device-resident.

    python examples/cvrp_device_solve.py --replicas 64 --steps 2000
    torchrun --nproc-per-node 8 examples/cvrp_device_solve.py     # one rank per GPU, NCCL best-score sync
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solverforge_b200 import instances, models, replicas  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=64)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--customers", type=int, default=1000)
    ap.add_argument("--vehicles", type=int, default=80)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    inst = instances.cvrp(args.customers, args.vehicles, seed=7)
    mine = replicas.partition_replicas(args.replicas * world, world, rank)
    starts = [instances.perturb_routes(inst, replicas.replica_seed(r), 32) for r in mine]
    d = models.cvrp_director(inst, len(mine), offsets=np.stack([s[0] for s in starts]),
                             elems=np.concatenate([s[1] for s in starts]), device=local)
    init = d.calculate_score()
    t0 = time.perf_counter()
    best, evaluated, committed = d.solve_nearby_list_change(args.steps, 20, acceptor=2, late_size=400,
                                                            accepted_limit=256, seed_base=1000 + rank,
                                                            restore_best=True)
    dt = time.perf_counter() - t0
    order = np.lexsort((best[:, 1], best[:, 0]))
    top = order[-1]
    print(f"[rank {rank}] {len(mine)} replicas x {args.steps} steps in {dt:.2f} s "
          f"({len(mine) * args.steps / dt:,.0f} solver steps/s, {evaluated.sum() / dt:,.0f} moves evaluated/s)")
    print(f"[rank {rank}] initial {init[top][0]}hard/{init[top][1]}soft -> best {best[top][0]}hard/{best[top][1]}soft "
          f"(replica {mine[top]}, fresh == committed: {np.array_equal(d.fresh_score(), d.calculate_score())})")
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
        key = torch.tensor([replicas.pack_score_key(int(best[top][0]), int(best[top][1]))], dtype=torch.int64,
                           device="cuda")
        gbest, owner = replicas.sync_best(key)
        if rank == 0:
            print(f"global best {replicas.unpack_score_key(gbest)} owned by rank {owner}")
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
