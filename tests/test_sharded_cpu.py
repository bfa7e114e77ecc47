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

from vecgo_b200.sharded import ShardedIndex, exchange_topk, owned_local_rows, shard_range


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


# ------------------------------------------------------------ quantized scan + rerank across shards
def test_owned_local_rows():
    rows = torch.tensor([[5, 9, 10, 19, 20, -1]], dtype=torch.int32)  # -1 = 0xFFFFFFFF (empty slot)
    local, owned = owned_local_rows(rows, row_base=10, nrows=10)
    assert owned.tolist() == [[False, False, True, True, False, False]]
    assert local.tolist() == [[-1, -1, 0, 9, -1, -1]]
    big = torch.tensor([[-2]], dtype=torch.int32)                      # global row 0xFFFFFFFE on a shard starting at 2^31
    local, owned = owned_local_rows(big, row_base=2 ** 31, nrows=2 ** 31 - 1)
    assert bool(owned[0, 0]) and int(local[0, 0]) == 2 ** 31 - 2


class _OracleShard:
    """Stands in for a DeviceIndex on CPU: SQ8 scan and float32 rerank by the oracle, same call signatures."""

    def __init__(self, o, x, codes, mins, inv, lo):
        self.o, self.rows, self.row_base, self.dim = o, len(x), lo, x.shape[1]
        self.seg = o.FlatOracle(dim=x.shape[1], metric=0, quant=1, codes=codes, mins=mins, inv=inv, vectors=x)

    @staticmethod
    def _view(ptr, shape, dtype):
        import ctypes

        n = int(np.prod(shape))
        buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def search_dev(self, d_q, nq, k, d_rows, d_scores, d_counts):
        q = self._view(d_q, (nq, self.dim), np.float32)
        out, cnt = self.seg.search_batch(q, k)
        rows = self._view(d_rows, (nq, k), np.uint32)
        sc = self._view(d_scores, (nq, k), np.float32)
        rows[:] = 0xFFFFFFFF
        for i in range(nq):
            c = int(cnt[i])
            rows[i, :c] = out["row"][i, :c] + self.row_base
            sc[i, :c] = out["score"][i, :c]
        self._view(d_counts, (nq,), np.int32)[:] = cnt

    def rerank_dev(self, d_q, nq, d_rows, r, d_scores):
        q = self._view(d_q, (nq, self.dim), np.float32)
        rows = self._view(d_rows, (nq, r), np.uint32)
        sc = self._view(d_scores, (nq, r), np.float32)
        for i in range(nq):
            ok = rows[i] < self.rows
            sc[i] = np.nan
            if ok.any():
                sc[i, ok] = self.seg.rerank(q[i], rows[i, ok])


def _numpy_merge(all_rows, all_scores, k_in, k_out, descending):
    w, nq, _ = all_rows.shape
    r = (all_rows.numpy().astype(np.int64) & 0xFFFFFFFF).transpose(1, 0, 2).reshape(nq, -1)
    s = all_scores.numpy().transpose(1, 0, 2).reshape(nq, -1)
    orow = np.full((nq, k_out), 0xFFFFFFFF, np.int64)
    osc = np.full((nq, k_out), np.nan, np.float32)
    cnt = np.zeros(nq, np.int32)
    for i in range(nq):
        live = r[i] != 0xFFFFFFFF
        rr, ss = r[i][live], s[i][live]
        order = np.lexsort((rr, -ss if descending else ss))[:k_out]
        orow[i, :len(order)], osc[i, :len(order)], cnt[i] = rr[order], ss[order], len(order)
    to32 = np.where(orow >= 2 ** 31, orow - 2 ** 32, orow).astype(np.int32)
    return torch.from_numpy(to32), torch.from_numpy(osc), torch.from_numpy(cnt)


def _rerank_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o

    n, dim, nq, r, k = 3000, 32, 6, 40, 5
    rng = np.random.default_rng(8)
    x = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    mins, maxs, sc, inv = (np.zeros(dim, np.float32) for _ in range(4))
    o.lib.vgo_sq8_train(o.fp(x), n, dim, o.fp(mins), o.fp(maxs), o.fp(sc), o.fp(inv))
    codes = np.clip((np.clip(x, mins, maxs) - mins) * sc + np.float32(0.5), 0, 255).astype(np.uint8)
    lo, hi = shard_range(n, rank, world)
    shard = _OracleShard(o, x[lo:hi], codes[lo:hi], mins, inv, lo)
    sh = ShardedIndex(shard, descending=False, merge=_numpy_merge)
    rows, scores, cnt = sh.search_rerank_dev(torch.from_numpy(q), nq, r, k)
    # single-segment reference semantics: approximate top-r over ALL rows, exact rerank, top-k by (score, row)
    whole = o.FlatOracle(dim=dim, metric=0, quant=1, codes=codes, mins=mins, inv=inv, vectors=x)
    approx, acnt = whole.search_batch(q, r)
    ok = True
    for i in range(nq):
        cand = approx["row"][i, : acnt[i]]
        ex = whole.rerank(q[i], cand)
        order = np.lexsort((cand, ex))[:k]
        ok &= np.array_equal(rows[i].numpy().astype(np.int64) & 0xFFFFFFFF, cand[order].astype(np.int64))
        ok &= np.array_equal(scores[i].numpy().view(np.uint32), ex[order].view(np.uint32))
    np.save(os.path.join(out_dir, f"rr{rank}.npy"), np.array([ok]))
    dist.destroy_process_group()


def test_sharded_search_rerank_matches_single_segment_world2(tmp_path):
    world = 2
    mp.spawn(_rerank_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert bool(np.load(tmp_path / f"rr{r}.npy")[0])
