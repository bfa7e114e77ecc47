"""Host-side logic of the multi-GPU path on CPU: shard ranges and the all-gather
exchange (world_size 2, gloo).  The per-shard scan and the merge are played by the
oracle here (the CUDA scan/merge kernels are covered by the -m gpu tests); what is
under test is vecgo_b200.sharded's partitioning + exchange plumbing and the claim
that (score, global row) over shards equals the single-segment order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vecgo_b200.sharded import exchange_topk, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 100, 10_000_000, 200_000_001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, dim, nq, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o

    rng = np.random.default_rng(5)
    x = rng.random((n, dim)).astype(np.float32)
    x[n - 3] = x[2]  # a cross-shard exact tie → must resolve by global row id
    q = rng.random((nq, dim)).astype(np.float32)
    lo, hi = shard_range(n, rank, world)
    seg = o.FlatOracle(dim=dim, metric=0, vectors=x[lo:hi])
    local, cnt = seg.search_batch(q, k)
    rows = torch.from_numpy((local["row"].astype(np.int64) + lo).astype(np.int32))  # global ids (row_base + local)
    scores = torch.from_numpy(local["score"].copy())
    all_rows, all_scores = exchange_topk(rows, scores)
    assert all_rows.shape == (world, nq, k)
    # merge on the host with the reference order (score, row)
    r = all_rows.numpy().astype(np.int64).transpose(1, 0, 2).reshape(nq, -1)
    s = all_scores.numpy().transpose(1, 0, 2).reshape(nq, -1)
    merged = np.stack([r[i][np.lexsort((r[i], s[i]))[:k]] for i in range(nq)])
    whole = o.FlatOracle(dim=dim, metric=0, vectors=x)
    want, _ = whole.search_batch(q, k)
    ok = np.array_equal(merged, want["row"].astype(np.int64))
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([ok]))
    dist.destroy_process_group()


def test_exchange_and_global_order_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 4001, 24, 7, 10, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert bool(np.load(tmp_path / f"ok{r}.npy")[0])


def test_exchange_world1_is_identity():
    rows = torch.arange(12, dtype=torch.int32).reshape(3, 4)
    scores = torch.rand(3, 4)
    a, b = exchange_topk(rows, scores)
    assert a.shape == (1, 3, 4) and torch.equal(a[0], rows) and torch.equal(b[0], scores)
