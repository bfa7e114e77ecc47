"""Full-size (BASELINE.json shapes) property tests of the scan hot path on one B200.

The CPU oracle cannot score 10M-row segments in test time, so at the benchmark's own sizes the tensor-core paths are
checked through size-independent properties: (1) the filtered search returns bit for bit what this library's exact
CUDA-core scan returns (which the small-size tests pin to the oracle), (2) scores are sorted under the CandidateHeap
order with ties by row id, (3) searching two row shards and merging their top-k lists (vg_topk_merge, the multi-GPU
path) gives the result of the whole segment.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def vg():
    import vecgo_b200

    vecgo_b200._lib.call("vg_init", 0)
    vecgo_b200._lib.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
    return vecgo_b200


def bits(x):
    return np.ascontiguousarray(x, F).view(np.uint32)


def qtc_stats(vg):
    q, f = C.c_uint64(), C.c_uint64()
    vg._lib.call("vg_quant_tc_stats", C.byref(q), C.byref(f))
    return q.value, f.value


def build(vg, codec, n, dim, row_base=0, rows=None, seed=42):
    """Random codes generated on the device (uniform bytes; RaBitQ: random sign bits + norms in [20, 60))."""
    L = vg._lib
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    lo, hi = rows if rows is not None else (0, n)
    m = 96
    if codec == "sq8":
        ix = vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=hi - lo, row_base=row_base,
                                  sq8=(np.full(dim, -4, F), np.full(dim, 8 / 255, F)))
        cb = dim
    elif codec == "int4":
        ix = vg.index.DeviceIndex(codec=L.CODEC_INT4, metric=0, dim=dim, rows=hi - lo, row_base=row_base, int4=(np.full(dim, -4, F), np.full(dim, 8, F)))
        cb = dim // 2
    elif codec == "pq":
        ix = vg.index.DeviceIndex(codec=L.CODEC_PQ, metric=0, dim=dim, rows=hi - lo, row_base=row_base,
                                  pq=(rng.integers(-128, 128, m * 256 * (dim // m), dtype=np.int8), np.full(m, 0.01, F), np.zeros(m, F), m, 256))
        cb = m
    else:
        ix = vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=hi - lo, row_base=row_base)
        cb = dim // 8 + 4
    chunk = 1 << 20
    for c0 in range(0, n, chunk):  # chunk c of the WHOLE segment always gets the same bytes, whatever shard asks for it
        c1 = min(n, c0 + chunk)
        if c1 <= lo or c0 >= hi:
            continue
        g = torch.Generator(device=dev).manual_seed(seed * 1000 + c0 // chunk)
        codes = torch.randint(0, 256, (c1 - c0, cb), dtype=torch.uint8, device=dev, generator=g)
        if codec == "rabitq":
            norms = (torch.rand((c1 - c0,), device=dev, generator=g) * 40 + 20).to(torch.float32)
            codes[:, cb - 4:] = norms.view(torch.uint8).reshape(-1, 4)
        s, e = max(lo, c0), min(hi, c1)
        part = codes[s - c0:e - c0].contiguous()
        ix.upload_dev(e - s, d_codes=part.data_ptr(), row0=s - lo)
    torch.cuda.synchronize()
    return ix


@pytest.mark.parametrize("codec,n,dim,nq,k", [
    ("sq8", 10_000_000, 768, 32, 100),      # BASELINE configs[1]
    ("int4", 10_000_000, 768, 32, 100),     # BASELINE configs[1]
    ("pq", 25_000_000, 768, 32, 100),       # BASELINE configs[2], one GPU's shard of the 8-GPU layout
    ("rabitq", 12_500_000, 1536, 32, 1000), # BASELINE configs[3], one GPU's shard, rerank depth 1000
])
def test_full_size_filter_equals_exact_scan_and_shards_merge(vg, codec, n, dim, nq, k):
    L = vg._lib
    q = np.random.default_rng(43).standard_normal((nq, dim)).astype(F)
    ix = build(vg, codec, n, dim)
    try:
        before = qtc_stats(vg)
        rows, scores, counts = ix.search(q, k)
        after = qtc_stats(vg)
        assert after[0] - before[0] >= nq, "the full-size search did not go through the tensor-core filter"
        L.call("vg_flat_tc_enable", 0)
        try:
            e_rows, e_scores, e_counts = ix.search(q, k)
        finally:
            L.call("vg_flat_tc_enable", 1)
    finally:
        ix.close()
    # (1) identical to the exact CUDA-core scan
    assert np.array_equal(rows, e_rows) and np.array_equal(bits(scores), bits(e_scores)) and np.array_equal(counts, e_counts)
    # (2) CandidateHeap order: score ascending, ties by row id
    assert np.all(counts == k)
    assert np.all(np.diff(scores, axis=1) >= 0)
    tie = np.diff(scores, axis=1) == 0
    assert np.all(np.diff(rows.astype(np.int64), axis=1)[tie] > 0)
    # (3) two row shards + merge = the whole segment
    half = n // 2
    parts = []
    for lo, hi in ((0, half), (half, n)):
        sh = build(vg, codec, n, dim, row_base=lo, rows=(lo, hi))
        try:
            parts.append(sh.search(q[:8], k))
        finally:
            sh.close()
    all_rows = np.stack([p[0] for p in parts])
    all_scores = np.stack([p[1] for p in parts])
    m_rows, m_scores, m_counts = vg.index.topk_merge(all_rows, all_scores, False, k)
    assert np.array_equal(m_rows, rows[:8]) and np.array_equal(bits(m_scores), bits(scores[:8]))


@pytest.mark.parametrize("codec,n,dim", [("pq", 25_000_000, 768), ("sq8", 10_000_000, 768), ("int4", 10_000_000, 768),
                                         ("rabitq", 12_500_000, 1536)])
def test_full_size_repeatable_under_load(vg, codec, n, dim):
    """The same 512-query batch, searched again and again between full 10 000-query batches, must return the same bits
    every time (and the exact scan's).  Regression test of a timing-dependent hazard the 8-GPU merged-parity check of
    bench.py found: the kind::i8 PQ producers released a codebook-slice slot right behind their ld.shared gathers, and the
    next slice (an async-proxy bulk copy) could land while gathers were still in flight — about one batch in fourteen
    came back with one wrong neighbour, only on a GPU kept busy by large batches."""
    L = vg._lib
    dev = torch.device("cuda:0")
    k = 100
    ix = build(vg, codec, n, dim)
    g = torch.Generator(device=dev).manual_seed(5)
    q = torch.randn((10_000, dim), device=dev, generator=g)

    def run(nq):
        r = torch.empty((nq, k), dtype=torch.int32, device=dev)
        s = torch.empty((nq, k), dtype=torch.float32, device=dev)
        c = torch.empty((nq,), dtype=torch.int32, device=dev)
        ix.search_dev(q.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr())
        torch.cuda.synchronize()
        return r, s

    L.call("vg_flat_tc_enable", 0)
    try:
        er, es = run(64)
    finally:
        L.call("vg_flat_tc_enable", 1)
    first = run(512)
    assert torch.equal(first[0][:64], er) and torch.equal(first[1][:64].view(torch.int32), es.view(torch.int32))
    import os
    for _ in range(int(os.environ.get("VECGO_SOAK_ITERS", "8"))):
        run(10_000)
        again = run(512)
        assert torch.equal(again[0], first[0]) and torch.equal(again[1].view(torch.int32), first[1].view(torch.int32))
    ix.close()
