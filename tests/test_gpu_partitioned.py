"""IVF-partitioned segments (flat/segment.go:726-745) through the partition-grouped scan (vg_scan.cu scan_topk_partitioned):
every query scans only its nprobe probed partitions.  The result must be what the full scan with a per-row partition test
returns (VECGO_IVF_GROUPED=0 / vg_ivf_grouped_enable(0)) and what the oracle's flat.Search returns, bit for bit, with and
without a row bitmap, for float32, SQ8 and PQ segments — and at >= 1M rows."""
import time

import numpy as np
import pytest

from oracle import oracle as o

pytestmark = pytest.mark.gpu
F = np.float32


def bits(x):
    return np.ascontiguousarray(x, F).view(np.uint32)


@pytest.fixture(scope="module")
def vg():
    import vecgo_b200

    return vecgo_b200


def sections(vg, data, n, dim, P):
    cent = np.frombuffer(data, "<f4", P * dim, int.from_bytes(data[40:48], "little")).reshape(P, dim)
    poff = np.frombuffer(data, "<u4", P + 1, int.from_bytes(data[48:56], "little"))
    return cent, poff


def both_ways(vg, seg, q, k, nprobe, mask=None):
    L = vg._lib
    got = seg.Search(q, k, nprobes=nprobe, row_mask=mask)
    L.call("vg_ivf_grouped_enable", 0)
    try:
        full = seg.Search(q, k, nprobes=nprobe, row_mask=mask)
    finally:
        L.call("vg_ivf_grouped_enable", 1)
    assert np.array_equal(got[2], full[2])
    assert np.array_equal(got[0], full[0])
    assert np.array_equal(bits(got[1]), bits(full[1]))
    return got


@pytest.mark.parametrize("quant", ["none", "sq8", "pq"])
def test_partition_grouped_scan_matches_full_scan_and_oracle(vg, quant):
    rng = np.random.default_rng(11)
    n, dim, P, nq, k = 60_000, 64, 12, 37, 10
    v = (rng.standard_normal((n, dim)) + 2.0 * rng.integers(0, 3, (n, 1))).astype(F)
    q = (rng.standard_normal((nq, dim)) + 2.0 * rng.integers(0, 3, (nq, 1))).astype(F)
    qz = {"none": vg.flat.QuantizationNone, "sq8": vg.flat.QuantizationSQ8, "pq": vg.flat.QuantizationPQ}[quant]
    data = vg.flat.write_segment(segment_id=3, vectors=v, metric=0, k_partitions=P, seed=5, quantization=qz, pq_m=8, pq_iters=3)
    hdr = vg.flat.decode_header(data)
    assert hdr["num_partitions"] == P
    cent, poff = sections(vg, data, n, dim, P)
    seg = vg.flat.Segment.Open(data)
    mask_bits = rng.random(n) < 0.6
    mask = np.packbits(mask_bits, bitorder="little")
    for nprobe in (1, 3, P):
        for m in (None, mask):
            rows, scores, counts = both_ways(vg, seg, q, k, nprobe, m)
            if m is not None:
                live = rows[rows != 0xFFFFFFFF]
                assert mask_bits[live].all()
            if quant == "none" and m is None:
                vv = np.frombuffer(data, "<f4", n * dim, int.from_bytes(data[72:80], "little")).reshape(n, dim)
                so = o.FlatOracle(dim=dim, metric=0, vectors=vv, centroids=cent, partition_offsets=poff, segment_id=3)
                out, cnt = so.search_batch(q, k, nprobes=nprobe)
                for i in range(nq):
                    c = int(cnt[i])
                    assert int(counts[i]) == c
                    assert np.array_equal(rows[i, :c], out[i, :c]["row"])
                    assert np.array_equal(bits(scores[i, :c]), bits(out[i, :c]["score"]))
    seg.Close()


def test_partition_grouped_scan_one_million_rows(vg):
    """1M x 32-d float32 rows in 122 partitions (rows / 8192, the writer's rule), 512 queries, nprobe = 8: grouped vs full
    scan bit-identical, a sample of the queries against the oracle, and the grouped scan is the faster one."""
    rng = np.random.default_rng(12)
    n, dim, nq, k, nprobe = 1_000_000, 32, 512, 10, 8
    P = n // 8192
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    data = vg.flat.write_segment(segment_id=4, vectors=v, metric=0, k_partitions=P, seed=9, kmeans_iters=3)
    cent, poff = sections(vg, data, n, dim, P)
    seg = vg.flat.Segment.Open(data)
    L = vg._lib
    rows, scores, counts = both_ways(vg, seg, q, k, nprobe)
    t = {}
    for mode in (1, 0):
        L.call("vg_ivf_grouped_enable", mode)
        seg.Search(q, k, nprobes=nprobe)
        t0 = time.perf_counter()
        for _ in range(3):
            seg.Search(q, k, nprobes=nprobe)
        t[mode] = (time.perf_counter() - t0) / 3
    L.call("vg_ivf_grouped_enable", 1)
    print(f"1M rows, {P} partitions, nprobe {nprobe}, {nq} queries: grouped {t[1] * 1e3:.2f} ms, full scan {t[0] * 1e3:.2f} ms")
    assert t[1] < t[0]
    vv = np.frombuffer(data, "<f4", n * dim, int.from_bytes(data[72:80], "little")).reshape(n, dim)
    so = o.FlatOracle(dim=dim, metric=0, vectors=vv, centroids=cent, partition_offsets=poff, segment_id=4)
    out, cnt = so.search_batch(q[:16], k, nprobes=nprobe)
    for i in range(16):
        c = int(cnt[i])
        assert int(counts[i]) == c
        assert np.array_equal(rows[i, :c], out[i, :c]["row"])
        assert np.array_equal(bits(scores[i, :c]), bits(out[i, :c]["score"]))
    seg.Close()
