"""World-size-2 gloo tests of the multi-GPU host logic (replica partition + best-score sync)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from solverforge_b200 import replicas


def test_partition_and_key_order():
    assert replicas.partition_replicas(8, 2, 0) == [0, 2, 4, 6]
    assert replicas.partition_replicas(8, 2, 1) == [1, 3, 5, 7]
    assert sorted(sum((replicas.partition_replicas(13, 4, g) for g in range(4)), [])) == list(range(13))
    scores = [(-3, 0), (-3, -10), (0, -547906), (0, -1), (-1, 5), (0, 0), (2, -7)]
    keys = [replicas.pack_score_key(*s) for s in scores]
    assert sorted(range(len(scores)), key=lambda i: keys[i]) == sorted(range(len(scores)), key=lambda i: scores[i])
    for s in scores:
        assert replicas.unpack_score_key(replicas.pack_score_key(*s)) == s
    assert replicas.pack_score_key(-(1 << 40), 0) == replicas.pack_score_key(-(1 << 22), 0)  # saturates


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = {0: (-2, -100), 1: (0, -250)}[rank]      # rank 1 holds the better (feasible) score
    key = torch.tensor([replicas.pack_score_key(*local)], dtype=torch.int64)
    best, owner = replicas.sync_best(key)
    out[rank] = (replicas.unpack_score_key(best), owner)
    # a tie is owned by the lowest rank
    tie = torch.tensor([replicas.pack_score_key(0, -5)], dtype=torch.int64)
    out[rank + 10] = replicas.sync_best(tie)[1]
    # the exact (hard, soft) sync: scores far outside the packed key's range still order correctly
    big = {0: [[-(1 << 30), 5], [-(1 << 30), 9]], 1: [[-(1 << 30) - 1, 1 << 50], [-(1 << 31), 0]]}[rank]
    out[rank + 20] = replicas.sync_best_scores(torch.tensor(big, dtype=torch.int64))
    out[rank + 30] = replicas.sync_best_scores(torch.tensor([[3, -7]], dtype=torch.int64))   # tie -> lowest rank
    dist.destroy_process_group()


def test_sync_best_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        assert out[0] == ((0, -250), 1) and out[1] == ((0, -250), 1)
        assert out[10] == 0 and out[11] == 0
        assert out[20] == (-(1 << 30), 9, 0) and out[21] == (-(1 << 30), 9, 0)
        assert out[30] == (3, -7, 0) and out[31] == (3, -7, 0)
        assert replicas.key_saturates(-(1 << 30), 0) and not replicas.key_saturates(-89, -511236)
