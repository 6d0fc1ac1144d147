"""world_size-2 gloo tests (CPU) of the N>1 host logic: block sharding of pairs, the single
gather of per-pair results to rank 0, and the secondary mode (candidates of a pair spread over
ranks, all-gathered back into list order)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from ffsubsync_b200 import distributed as D
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # primary mode: 7 pairs over 2 ranks (uneven blocks), results = f(global pair index)
    n_pairs = 7
    lo, hi = D.shard_pairs(n_pairs, rank, world)
    local = torch.tensor([[100.0 + p, -p, p % 5] for p in range(lo, hi)], dtype=torch.float64)
    got = D.gather_pair_results(local, n_pairs, rank, world)
    if rank == 0:
        want = torch.tensor([[100.0 + p, -p, p % 5] for p in range(n_pairs)], dtype=torch.float64)
        assert torch.equal(got, want)
    else:
        assert got is None
    # secondary mode: B=3 pairs, K=7 candidates dealt round-robin over ranks
    B, K = 3, 7
    ks = D.shard_candidates(K, rank, world)
    loc = torch.tensor([[[10.0 * b + k, k - 3] for k in ks] for b in range(B)], dtype=torch.float64)
    full = D.allgather_candidate_results(loc, K, rank, world)
    want = torch.tensor([[[10.0 * b + k, k - 3] for k in range(K)] for b in range(B)], dtype=torch.float64)
    assert torch.equal(full, want)
    t = D.max_over_ranks(1.0 + rank)
    assert t == float(world)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
    q.put((rank, "ok"))


def test_gloo_world_size_2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(world)) == [(0, "ok"), (1, "ok")]


def test_shard_pairs_partitions_everything():
    from ffsubsync_b200.distributed import shard_candidates, shard_pairs
    for n in (0, 1, 7, 256, 4096, 4099):
        for world in (1, 2, 4, 8):
            spans = [shard_pairs(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert sorted(sum((shard_candidates(7, r, 4) for r in range(4)), [])) == list(range(7))
