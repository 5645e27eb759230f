"""Latency of the path's only collective (8-byte MAX all-reduce) in isolation; run under torchrun."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = torch.zeros(1, dtype=torch.int64, device=dev)
for _ in range(5):
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
torch.cuda.synchronize()
for stream in (None, torch.cuda.Stream()):
    ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
    with ctx:
        for _ in range(3):
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(50):
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e1.record()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        if rank == 0:
            print(f"stream={'custom' if stream is not None else 'default'}: {e0.elapsed_time(e1) / 50 * 1000:.1f} us/allreduce (events), "
                  f"{(w1 - w0) / 50 * 1e6:.1f} us wall", flush=True)
dist.destroy_process_group()
