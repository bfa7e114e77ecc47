"""Row sharding across GPUs: one process per GPU (torch.distributed), shard r owns
rows [r*N/W, (r+1)*N/W), every rank scans its shard for the whole query batch,
then ONE all-gather of the k best (score, global row) pairs per query and a
device-side merge (vg_topk_merge_dev) produce the global top-k on every rank.

This is the multi-segment merge of internal/engine/search.go:903-908 with the
segments living on different GPUs: global row id = row_base + local row, so the
reference's (score, SegmentID, RowID) tie-break is reproduced by (score, global
row) with one logical segment.  torch.distributed is plumbing
only: the exchange is k*8 bytes per query per rank over NCCL/NVLink.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of `rank` (first `total % world` shards get one extra row)."""
    base, rem = divmod(total_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def exchange_topk(rows, scores, group=None):
    """All-gather per-rank best-first lists.  rows/scores: torch tensors [nq, k] (int32 bit
    patterns of uint32 row ids / float32).  Returns ([W, nq, k], [W, nq, k]).  Works on any
    backend (NCCL on GPU; gloo on CPU for the host-logic tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return rows.unsqueeze(0), scores.unsqueeze(0)
    nq, k = rows.shape
    # concatenated-along-dim-0 output is the layout every backend (NCCL, gloo) accepts
    all_rows = torch.empty((world * nq, k), dtype=rows.dtype, device=rows.device)
    all_scores = torch.empty((world * nq, k), dtype=scores.dtype, device=scores.device)
    dist.all_gather_into_tensor(all_rows, rows.contiguous(), group=group)
    dist.all_gather_into_tensor(all_scores, scores.contiguous(), group=group)
    return all_rows.view(world, nq, k), all_scores.view(world, nq, k)


def pq_train_sharded(d_vecs, n: int, dim: int, m: int, k: int, iters: int, seed: int, codebooks, scales, offsets, centroids=None,
                     group=None):
    """ProductQuantizer.Train across GPUs (pq.go:68-143): the reference trains every subspace in its own goroutine
    (pq.go:79-140) — they are independent — so rank r trains subspaces shard_range(m, r, W) of the SAME training set
    (d_vecs: CUDA float32 [n, dim], replicated) and the slices are all-gathered: codebooks, scales and offsets are
    bit-identical to the single-GPU training.  Outputs are the caller's numpy arrays (codebooks int8 [m*k*ds], scales /
    offsets float32 [m], optional centroids float32 [m, k, ds]), filled on every rank."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from . import _lib as L

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ds = dim // m
    lo, hi = shard_range(m, rank, world)
    g = hi - lo
    cb = np.zeros(max(g, 1) * k * ds, np.int8)
    sc, of = np.zeros(max(g, 1), np.float32), np.zeros(max(g, 1), np.float32)
    cent = np.zeros((max(g, 1), k, ds), np.float32)
    if d_vecs.is_cuda:
        L.call("vg_set_stream", torch.cuda.current_stream(d_vecs.device).cuda_stream)
    L.call("vg_pq_train_range_dev", d_vecs.data_ptr(), n, dim, m, k, iters, seed, lo, hi, L.ptr(cb, L.i8p), L.ptr(sc, L.f32p),
           L.ptr(of, L.f32p), L.ptr(cent, L.f32p))
    if world == 1:
        parts = [(lo, hi, cb, sc, of, cent)]
    elif m % world == 0:
        # equal slices: ONE all-gather of the packed bytes (codebook | scales | offsets | float32 centroids)
        blob = np.concatenate([cb.view(np.uint8), sc.view(np.uint8), of.view(np.uint8), cent.reshape(-1).view(np.uint8)])
        mine = torch.from_numpy(blob).to(d_vecs.device)
        allb = torch.empty((world, blob.size), dtype=torch.uint8, device=d_vecs.device)
        dist.all_gather_into_tensor(allb.view(-1), mine, group=group)
        hb = allb.cpu().numpy()
        parts = []
        n_cb = g * k * ds
        for r in range(world):
            b = hb[r]
            parts.append((r * g, (r + 1) * g, b[:n_cb].view(np.int8), b[n_cb:n_cb + 4 * g].view(np.float32),
                          b[n_cb + 4 * g:n_cb + 8 * g].view(np.float32), b[n_cb + 8 * g:].view(np.float32).reshape(g, k, ds)))
    else:
        objs = [None] * world
        dist.all_gather_object(objs, (lo, hi, cb, sc, of, cent), group=group)
        parts = objs
    for lo_, hi_, cb_, sc_, of_, cent_ in parts:
        if hi_ <= lo_:
            continue
        codebooks[lo_ * k * ds:hi_ * k * ds] = cb_[:(hi_ - lo_) * k * ds]
        scales[lo_:hi_] = sc_[:hi_ - lo_]
        offsets[lo_:hi_] = of_[:hi_ - lo_]
        if centroids is not None:
            centroids[lo_:hi_] = cent_[:hi_ - lo_]


EMPTY_ROW = 0xFFFFFFFF


def owned_local_rows(rows_global, row_base: int, nrows: int):
    """Split a [nq, r] tensor of GLOBAL row ids (uint32 bit patterns in int32) into this shard's view:
    returns (local ids with 0xFFFFFFFF where the row lives on another shard, owned mask)."""
    import torch

    g = rows_global.to(torch.int64) & 0xFFFFFFFF
    local = g - int(row_base)
    owned = (local >= 0) & (local < int(nrows)) & (g != EMPTY_ROW)
    local = torch.where(owned, local, torch.full_like(local, EMPTY_ROW))
    return _as_u32_bits(local), owned


def _as_u32_bits(t):
    """int64 values in [0, 2^32) -> int32 tensor holding the same uint32 bit patterns."""
    import torch

    return torch.where(t >= 2 ** 31, t - 2 ** 32, t).to(torch.int32)


class ShardedIndex:
    """A DeviceIndex holding this rank's row shard + the cross-GPU merge."""

    def __init__(self, index, descending: bool, group=None, approx_descending: bool = False, merge=None):
        self.index = index
        self.descending = descending              # order of the segment metric (exact scores)
        self.approx_descending = approx_descending  # order of the codec's approximate scores (distances: ascending)
        self.group = group
        self._merge_fn = merge                    # test hook; default = vg_topk_merge_dev

    @staticmethod
    def _bind_stream(t):
        """The library must run on the stream torch (and NCCL) order their work on: rows / scores written by library
        kernels are read by torch ops and collectives, and their outputs are read by library kernels again.  Binding the
        calling thread's library stream to torch's current stream makes every step of a search one in-order queue."""
        if t.is_cuda:
            import torch

            from . import _lib as L

            L.call("vg_set_stream", torch.cuda.current_stream(t.device).cuda_stream)

    def _merge(self, all_rows, all_scores, k_in: int, k_out: int, descending: bool):
        import torch

        from . import _lib as L

        if self._merge_fn is not None:
            return self._merge_fn(all_rows, all_scores, k_in, k_out, descending)
        world, nq = all_rows.shape[0], all_rows.shape[1]
        dev = all_rows.device
        orow = torch.empty((nq, k_out), dtype=torch.int32, device=dev)
        osc = torch.empty((nq, k_out), dtype=torch.float32, device=dev)
        ocnt = torch.empty((nq,), dtype=torch.int32, device=dev)
        L.call("vg_topk_merge_dev", all_rows.contiguous().data_ptr(), all_scores.contiguous().data_ptr(), world, nq, k_in,
               int(descending), k_out, orow.data_ptr(), osc.data_ptr(), ocnt.data_ptr())
        return orow, osc, ocnt

    def _exchange_merge(self, rows, scores, counts, k_in: int, k_out: int, descending: bool):
        """Per-shard best-first lists -> the global best k_out on every rank.  On the GPU the exchange is ONE all-gather
        of 8-byte sortable keys (vg_topk_pack_dev) that vg_topk_merge_keys_dev merges directly; the two-tensor form
        (exchange_topk + a merge callback) serves the CPU / gloo tests of the host logic."""
        import torch
        import torch.distributed as dist

        from . import _lib as L

        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1 and k_in == k_out and self._merge_fn is None:
            return rows, scores, counts
        if self._merge_fn is not None or not rows.is_cuda:
            all_rows, all_scores = exchange_topk(rows, scores, self.group)
            return self._merge(all_rows, all_scores, k_in, k_out, descending)
        nq, dev = rows.shape[0], rows.device
        keys = torch.empty((nq, k_in), dtype=torch.int64, device=dev)
        L.call("vg_topk_pack_dev", rows.data_ptr(), scores.data_ptr(), nq * k_in, int(descending), keys.data_ptr())
        if world > 1:
            allk = torch.empty((world * nq, k_in), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allk, keys, group=self.group)
        else:
            allk = keys
        orow = torch.empty((nq, k_out), dtype=torch.int32, device=dev)
        osc = torch.empty((nq, k_out), dtype=torch.float32, device=dev)
        ocnt = torch.empty((nq,), dtype=torch.int32, device=dev)
        L.call("vg_topk_merge_keys_dev", allk.data_ptr(), world, nq, k_in, int(descending), k_out, orow.data_ptr(), osc.data_ptr(),
               ocnt.data_ptr())
        return orow, osc, ocnt

    def search_rerank_dev(self, d_queries, nq: int, r: int, k: int):
        """Quantized scan + exact rerank across shards with the reference's semantics (engine/search.go:188-192,
        913-973): the GLOBAL approximate top-r is reranked, not each shard's own top-r, so the ids are identical
        for every world size.  Two exchanges: (1) all-gather the per-shard approximate top-r and merge to the global
        top-r on every rank; (2) every rank scores exactly the rows it owns (Segment.Rerank), all-gather the exact
        (score, row) pairs and merge to the final top-k."""
        import torch

        self._bind_stream(d_queries)
        dev = d_queries.device
        rows = torch.empty((nq, r), dtype=torch.int32, device=dev)
        scores = torch.empty((nq, r), dtype=torch.float32, device=dev)
        counts = torch.empty((nq,), dtype=torch.int32, device=dev)
        self.index.search_dev(d_queries.data_ptr(), nq, r, rows.data_ptr(), scores.data_ptr(), counts.data_ptr())
        rows, scores, counts = self._exchange_merge(rows, scores, counts, r, r, self.approx_descending)
        local, owned = owned_local_rows(rows, self.index.row_base, self.index.rows)
        exact = torch.empty((nq, r), dtype=torch.float32, device=dev)
        self.index.rerank_dev(d_queries.data_ptr(), nq, local.contiguous().data_ptr(), r, exact.data_ptr())
        mine = torch.where(owned, rows, torch.full_like(rows, -1))  # -1 = 0xFFFFFFFF: not scored here
        return self._exchange_merge(mine, exact, counts, r, k, self.descending)

    def search_dev(self, d_queries, nq: int, k: int):
        """d_queries: CUDA float32 tensor [nq, dim] (replicated on every rank).
        Returns (rows int32-viewed-uint32 [nq,k], scores [nq,k], counts [nq]) — identical on all ranks."""
        import torch

        self._bind_stream(d_queries)
        dev = d_queries.device
        rows = torch.empty((nq, k), dtype=torch.int32, device=dev)
        scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
        counts = torch.empty((nq,), dtype=torch.int32, device=dev)
        self.index.search_dev(d_queries.data_ptr(), nq, k, rows.data_ptr(), scores.data_ptr(), counts.data_ptr())
        return self._exchange_merge(rows, scores, counts, k, k, self.descending)


class ShardGroup:
    """vg_shard_group_* (include/vecgo_cuda.h): row shards with the NCCL exchange and the merges INSIDE the library — the
    form a Go host binds.  `ShardGroup.single_process(devices)` drives all GPUs from this process (one shard handle per
    GPU); `ShardGroup.from_torch_distributed(device)` makes one member per process (the launcher — torch.distributed here
    — only ships the 128-byte NCCL id)."""

    def __init__(self, handle: int, world: int, members: int):
        self.handle, self.world, self.members = handle, world, members

    @classmethod
    def single_process(cls, devices):
        import ctypes as C

        import numpy as np

        from . import _lib as L

        d = np.ascontiguousarray(devices, np.int32)
        h = C.c_uint64()
        L.call("vg_shard_group_create", L.ptr(d, L.i32p), len(d), C.byref(h))
        return cls(h.value, len(d), len(d))

    @classmethod
    def from_torch_distributed(cls, device: int, group=None):
        import ctypes as C

        import numpy as np
        import torch
        import torch.distributed as dist

        from . import _lib as L

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ident = np.zeros(128, np.uint8)
        if rank == 0:
            L.call("vg_nccl_unique_id", L.ptr(ident, L.u8p))
        t = torch.from_numpy(ident)
        if dist.get_backend(group) == "nccl":
            t = t.cuda(device)
        dist.broadcast(t, src=0, group=group)
        ident = t.cpu().numpy().copy()
        h = C.c_uint64()
        L.call("vg_shard_group_create_rank", L.ptr(ident, L.u8p), rank, world, device, C.byref(h))
        return cls(h.value, world, 1)

    def _handles(self, indexes):
        import numpy as np

        idx = indexes if isinstance(indexes, (list, tuple)) else [indexes]
        if len(idx) != self.members:
            raise ValueError(f"this group drives {self.members} shard(s) from this process, got {len(idx)} handle(s)")
        return np.array([ix.handle for ix in idx], np.uint64)

    def search(self, indexes, queries, k: int, r: int = 0):
        """Host queries in, merged result out (from member 0).  r > 0: approximate top-r, exact rerank, final top-k."""
        import numpy as np

        from . import _lib as L

        h = self._handles(indexes)
        q = L.as_f32(queries)
        q = q.reshape(-1, q.shape[-1])
        nq = q.shape[0]
        rows = np.full((nq, k), 0xFFFFFFFF, np.uint32)
        scores = np.full((nq, k), np.nan, np.float32)
        counts = np.zeros(nq, np.int32)
        if r > 0:
            L.call("vg_shard_group_search_rerank", self.handle, L.ptr(h, L.u64p), L.ptr(q, L.f32p), nq, r, k, L.ptr(rows, L.u32p),
                   L.ptr(scores, L.f32p), L.ptr(counts, L.i32p))
        else:
            L.call("vg_shard_group_search", self.handle, L.ptr(h, L.u64p), L.ptr(q, L.f32p), nq, k, L.ptr(rows, L.u32p), L.ptr(scores, L.f32p),
                   L.ptr(counts, L.i32p))
        return rows, scores, counts

    def search_dev(self, indexes, d_queries, nq: int, k: int, r: int = 0):
        """Device form: d_queries = one CUDA float32 tensor [nq, dim] per member (a single tensor for a one-member group).
        Returns per member (rows int32-viewed-uint32 [nq,k], scores, counts), complete on return."""
        import ctypes as C

        import torch

        from . import _lib as L

        h = self._handles(indexes)
        qs = d_queries if isinstance(d_queries, (list, tuple)) else [d_queries]
        outs = []
        for t in qs:
            torch.cuda.current_stream(t.device).synchronize()   # the group runs on its own streams
            outs.append((torch.empty((nq, k), dtype=torch.int32, device=t.device), torch.empty((nq, k), dtype=torch.float32, device=t.device),
                         torch.empty((nq,), dtype=torch.int32, device=t.device)))
        arr = lambda ps: (C.c_void_p * len(ps))(*ps)  # noqa: E731
        qp, rp, sp, cp = arr([t.data_ptr() for t in qs]), arr([o[0].data_ptr() for o in outs]), arr([o[1].data_ptr() for o in outs]), \
            arr([o[2].data_ptr() for o in outs])
        if r > 0:
            L.call("vg_shard_group_search_rerank_dev", self.handle, L.ptr(h, L.u64p), qp, nq, r, k, rp, sp, cp)
        else:
            L.call("vg_shard_group_search_dev", self.handle, L.ptr(h, L.u64p), qp, nq, k, rp, sp, cp)
        return outs if isinstance(d_queries, (list, tuple)) else outs[0]

    def close(self):
        from . import _lib as L

        if self.handle:
            L.call("vg_shard_group_destroy", self.handle)
            self.handle = 0
