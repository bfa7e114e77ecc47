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


class ShardedIndex:
    """A DeviceIndex holding this rank's row shard + the cross-GPU merge."""

    def __init__(self, index, descending: bool, group=None):
        self.index = index
        self.descending = descending
        self.group = group

    def search_dev(self, d_queries, nq: int, k: int):
        """d_queries: CUDA float32 tensor [nq, dim] (replicated on every rank).
        Returns (rows int32-viewed-uint32 [nq,k], scores [nq,k], counts [nq]) — identical on all ranks."""
        import torch

        from . import _lib as L

        dev = d_queries.device
        rows = torch.empty((nq, k), dtype=torch.int32, device=dev)
        scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
        counts = torch.empty((nq,), dtype=torch.int32, device=dev)
        self.index.search_dev(d_queries.data_ptr(), nq, k, rows.data_ptr(), scores.data_ptr(), counts.data_ptr())
        all_rows, all_scores = exchange_topk(rows, scores, self.group)
        world = all_rows.shape[0]
        if world == 1:
            return rows, scores, counts
        orow = torch.empty_like(rows)
        osc = torch.empty_like(scores)
        ocnt = torch.empty_like(counts)
        L.call("vg_topk_merge_dev", all_rows.data_ptr(), all_scores.data_ptr(), world, nq, k, int(self.descending), k,
               orow.data_ptr(), osc.data_ptr(), ocnt.data_ptr())
        return orow, osc, ocnt
