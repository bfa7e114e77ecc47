"""Small invocations of every hot-path kernel family for compute-sanitizer (tools/gpu_evidence.sh sanitize):
Flat search through the CTA-pair fp16 filter and the single-CTA TF32 filter, SQ8 / INT4 / PQ / RaBitQ / BQ through the
decode-GEMM filter, the exact CUDA-core scans, rerank, gather scoring, bounded L2, top-k merge, PQ training (tensor-core
assignment + exact k-means++ prefix).  Sizes are the smallest the filters accept: the tool replays every launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import vecgo_b200 as vg

L = vg._lib
F = np.float32
rng = np.random.default_rng(0)
which = sys.argv[1:] or ["flat", "quant", "misc", "train"]

if "flat" in which:
    n, dim, nq, k = 8192, 128, 32, 10
    x, q = rng.random((n, dim), dtype=F), rng.random((nq, dim), dtype=F)
    with vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        r1, s1, _ = ix.search(q, k)
        L.call("vg_flat_tc_enable", 0)
        r2, s2, _ = ix.search(q, k)
        L.call("vg_flat_tc_enable", 1)
        ix.rerank(q, r1)
        ix.l2_bounded(q, r1, np.full(nq, 10.0, F))
        # block-stat skipping: the CTA-pair filter (k > 16) over a tile list
        keep = np.array([True, False, False, True, True, False, True, False])
        r3, s3, _ = ix.search(q, 20, block_keep=keep)
        L.call("vg_flat_tc_enable", 0)
        r4, s4, _ = ix.search(q, 20, block_keep=keep)
        L.call("vg_flat_tc_enable", 1)
        assert np.array_equal(r3, r4) and np.array_equal(s3.view(np.uint32), s4.view(np.uint32))
    assert np.array_equal(r1, r2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32))
    # IVF-partitioned segment: partition-grouped scan
    data = vg.flat.write_segment(segment_id=1, vectors=x, metric=0, k_partitions=4, seed=1, kmeans_iters=2)
    seg = vg.flat.Segment.Open(data)
    g1 = seg.Search(q, k, nprobes=2)
    L.call("vg_ivf_grouped_enable", 0)
    g2 = seg.Search(q, k, nprobes=2)
    L.call("vg_ivf_grouped_enable", 1)
    seg.Close()
    assert np.array_equal(g1[0], g2[0]) and np.array_equal(g1[1].view(np.uint32), g2[1].view(np.uint32))
    print("flat ok", flush=True)
if "quant" in which:
    n, dim, nq, k = 8192, 128, 16, 10
    v, q = rng.standard_normal((n, dim)).astype(F), rng.standard_normal((nq, dim)).astype(F)
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(v)
    iq = vg.quantization.Int4Quantizer(dim)
    iq.Train(v)
    m = 16
    cb = rng.integers(-128, 128, m * 256 * (dim // m), dtype=np.int8)
    pq = (cb, np.full(m, 0.01, F), np.zeros(m, F), m, 256)
    cases = [
        dict(codec=L.CODEC_SQ8, sq8=(sq.mins, sq.invScales), codes=sq.EncodeBatch(v)),
        dict(codec=L.CODEC_INT4, int4=(iq.min, iq.diff), codes=iq.EncodeBatch(v)),
        dict(codec=L.CODEC_PQ, pq=pq, codes=rng.integers(0, 256, (n, m), dtype=np.uint8)),
        dict(codec=L.CODEC_RABITQ, codes=vg.quantization.RaBitQuantizer(dim).EncodeBatch(v)),
        dict(codec=L.CODEC_BQ, codes=vg.quantization.BinaryQuantizer(dim).EncodeBatch(v)),
    ]
    for c in cases:
        codes = c.pop("codes")
        with vg.index.DeviceIndex(metric=0, dim=dim, rows=n, **c) as ix:
            ix.upload(codes=codes)
            r1, s1, _ = ix.search(q, k)
            L.call("vg_flat_tc_enable", 0)
            r2, s2, _ = ix.search(q, k)
            L.call("vg_flat_tc_enable", 1)
            ix.score(q, r1)
            keep = np.array([True, False, True, True, False, False, True, False])
            r3, s3, _ = ix.search(q, k, block_keep=keep)     # tile list + skipped-group fill
            L.call("vg_flat_tc_enable", 0)
            r4, s4, _ = ix.search(q, k, block_keep=keep)
            L.call("vg_flat_tc_enable", 1)
            assert np.array_equal(r3, r4) and np.array_equal(s3.view(np.uint32), s4.view(np.uint32)), c["codec"]
            if c["codec"] == L.CODEC_SQ8:
                ix.set_host_vectors(v)                      # rerank from mapped host memory (staged gather)
                ix.rerank(q, r1)
        assert np.array_equal(r1, r2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32)), c["codec"]
    print("quant ok", flush=True)
if "misc" in which:
    rows = rng.integers(0, 1000, (3, 4, 8)).astype(np.uint32)
    vg.index.topk_merge(rows, rng.random((3, 4, 8), dtype=F), False, 8)
    vg.simd.SquaredL2Batch(rng.random(33, dtype=F), rng.random((7, 33), dtype=F), 33)
    print("misc ok", flush=True)
if "train" in which:
    v = rng.standard_normal((4096, 64)).astype(F)
    p = vg.quantization.ProductQuantizer(64, 8, 256)
    p.Train(v, iters=2, seed=3)
    print("train ok", flush=True)
