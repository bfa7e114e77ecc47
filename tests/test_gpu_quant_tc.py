"""GPU parity tests of the decode-GEMM tensor-core path of the quantized scans (vecgo_b200/csrc/vg_quant_tc.cu).

The fp16 GEMM over codes decoded inside the kernel only FILTERS; vg_index_search must return what
flat.(*Segment).Search returns on the SIMD path: row ids with ties by row id and float32 scores in the order of
simd.Sq8uL2BatchPerDimension / simd.Int4L2DistanceBatch / simd.PqAdcLookup — checked bit for bit against the CPU
oracle and against this library's exact CUDA-core scan — whatever the filter did, including a failing certificate.
"""
import ctypes as C
import os
import subprocess
import sys

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


def qtc_stats(vg):
    q, f = C.c_uint64(), C.c_uint64()
    vg._lib.call("vg_quant_tc_stats", C.byref(q), C.byref(f))
    return q.value, f.value


def check(rows, scores, counts, want):
    for i, w in enumerate(want):
        c = int(counts[i])
        assert c == len(w), (i, c, len(w))
        assert np.array_equal(rows[i, :c], w["row"]), i
        assert np.array_equal(bits(scores[i, :c]), bits(w["score"])), i
        assert np.all(rows[i, c:] == 0xFFFFFFFF)


def oracle_flat(queries, k, mask=None, **kw):
    seg = o.FlatOracle(**kw)
    out, cnt = seg.search_batch(queries, k, threads=8, mask=mask)
    return [out[i, : cnt[i]] for i in range(len(queries))]


def exact_scan(vg, make_index, q, k, **kw):
    """The same search on the exact CUDA-core scan (tensor-core filters off)."""
    vg._lib.call("vg_flat_tc_enable", 0)
    try:
        with make_index() as ix:
            return ix.search(q, k, **kw)
    finally:
        vg._lib.call("vg_flat_tc_enable", 1)


def random_pq(rng, dim, m, k=256):
    ds = dim // m
    cb = rng.integers(-128, 128, m * k * ds, dtype=np.int8)
    sc = (rng.random(m) * 0.02 + 0.005).astype(F)
    of = (rng.standard_normal(m) * 0.1).astype(F)
    return cb, sc, of


@pytest.mark.parametrize("n,dim,nq,k", [
    (20000, 768, 40, 10),     # lane-transposed layout VB=16, streamed 12 k-blocks
    (9000, 64, 300, 1),       # VB=4 layout, one k-block, two query tiles, ragged row tile
    (70000, 256, 64, 32),     # many row splits
    (150000, 128, 32, 100),   # k = 100 -> 200 candidate groups
    (300000, 128, 20, 1000),  # k = 1000 (the engine's refine depth): 2000 candidate groups
])
def test_sq8_tc_matches_oracle(vg, n, dim, nq, k):
    rng = np.random.default_rng(n + dim)
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(v)
    codes = sq.EncodeBatch(v)
    codes[n // 2] = codes[n // 3]   # duplicate rows -> equal scores -> tie by row id
    q[0] = v[n // 3]                # a query that sits on a stored row

    def make():
        ix = vg.index.DeviceIndex(codec=vg._lib.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales))
        ix.upload(codes=codes)
        return ix

    before = qtc_stats(vg)
    with make() as ix:
        rows, scores, counts = ix.search(q, k)
    after = qtc_stats(vg)
    assert after[0] - before[0] == nq, "the search did not go through the decode-GEMM filter"
    want = oracle_flat(q, k, dim=dim, metric=0, quant=1, codes=codes, mins=sq.mins, inv=sq.invScales)
    check(rows, scores, counts, want)
    r2, s2, c2 = exact_scan(vg, make, q, k)
    assert np.array_equal(rows, r2) and np.array_equal(bits(scores), bits(s2)) and np.array_equal(counts, c2)
    if k <= 32:
        assert after[1] - before[1] <= nq // 4, "too many certificate failures on benign data"


@pytest.mark.parametrize("n,dim,nq,k", [(30000, 768, 33, 10), (9000, 256, 20, 5), (12000, 64, 17, 10), (100000, 192, 32, 50),
                                        (280000, 256, 18, 1000)])
def test_int4_tc_matches_exact_scan(vg, n, dim, nq, k):
    rng = np.random.default_rng(n + dim + 1)
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    iq = vg.quantization.Int4Quantizer(dim)
    iq.Train(v)
    codes = iq.EncodeBatch(v)
    codes[n // 2] = codes[n // 3]

    def make():
        ix = vg.index.DeviceIndex(codec=vg._lib.CODEC_INT4, metric=0, dim=dim, rows=n, int4=(iq.min, iq.diff))
        ix.upload(codes=codes)
        return ix

    before = qtc_stats(vg)
    with make() as ix:
        rows, scores, counts = ix.search(q, k)
    after = qtc_stats(vg)
    assert after[0] - before[0] == nq
    r2, s2, c2 = exact_scan(vg, make, q, k)
    assert np.array_equal(rows, r2) and np.array_equal(bits(scores), bits(s2)) and np.array_equal(counts, c2)
    # and against the CPU oracle for a few queries (int4_avx512.c order)
    for i in range(min(nq, 6)):
        out = np.zeros(k, o.cand_dtype)
        c = o.lib.vgo_int4_search(o.fp(q[i]), o.bp(codes), n, dim, o.fp(iq.min), o.fp(iq.diff), k,
                                  o.fn_addr(o.lib.vgo_int4_l2_batch_a512), out.ctypes.data_as(C.POINTER(o.Cand)))
        assert np.array_equal(rows[i, :c], out[:c]["row"]) and np.array_equal(bits(scores[i, :c]), bits(out[:c]["score"]))


@pytest.mark.parametrize("n,dim,m,nq,k", [
    (20000, 768, 96, 33, 10),    # C3 shape: tiled code layout, dsub = 8
    (9000, 64, 8, 300, 1),       # one k-block
    (40000, 128, 8, 20, 10),     # dsub = 16
    (60000, 256, 32, 32, 100),   # k = 100
    (10000, 128, 16, 16, 10),     # m = 16, dsub = 8
    (300000, 768, 96, 17, 1000), # k = 1000 at the C3 code shape
])
def test_pq_tc_matches_oracle(vg, n, dim, m, nq, k):
    rng = np.random.default_rng(n + m)
    cb, sc, of = random_pq(rng, dim, m)
    codes = rng.integers(0, 256, (n, m), dtype=np.uint8)
    codes[n // 2] = codes[n // 3]
    q = (rng.standard_normal((nq, dim)) * 0.7).astype(F)

    def make():
        ix = vg.index.DeviceIndex(codec=vg._lib.CODEC_PQ, metric=0, dim=dim, rows=n, pq=(cb, sc, of, m, 256))
        ix.upload(codes=codes)
        return ix

    before = qtc_stats(vg)
    with make() as ix:
        rows, scores, counts = ix.search(q, k)
    after = qtc_stats(vg)
    assert after[0] - before[0] == nq
    want = oracle_flat(q[:8], k, dim=dim, metric=0, quant=2, codes=codes, pq=(cb, sc, of, m, 256))
    check(rows[:8], scores[:8], counts[:8], want)
    r2, s2, c2 = exact_scan(vg, make, q, k)
    assert np.array_equal(rows, r2) and np.array_equal(bits(scores), bits(s2)) and np.array_equal(counts, c2)


def test_tc_row_mask_row_base_and_chunked_upload(vg):
    n, dim, nq, k = 50000, 128, 48, 10
    rng = np.random.default_rng(77)
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(v)
    codes = sq.EncodeBatch(v)
    mask_bits = rng.random(n) < 0.3
    mask = np.packbits(mask_bits, bitorder="little")
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_SQ8, metric=0, dim=dim, rows=n, row_base=1000, sq8=(sq.mins, sq.invScales)) as ix:
        ix.upload(codes=codes[:20000])
        ix.upload(codes=codes[20000:], row0=20000)  # second upload invalidates the prepared norms
        before = qtc_stats(vg)
        rows, scores, counts = ix.search(q, k, row_mask=mask)
        assert qtc_stats(vg)[0] - before[0] == nq
    want = oracle_flat(q, k, dim=dim, metric=0, quant=1, codes=codes, mins=sq.mins, inv=sq.invScales, mask=mask)
    for i in range(nq):
        assert np.array_equal(rows[i], want[i]["row"] + 1000)
        assert np.array_equal(bits(scores[i]), bits(want[i]["score"]))
        assert mask_bits[rows[i] - 1000].all()


def test_tc_certificate_failure_falls_back_to_exact_scan(vg):
    """Thousands of identical rows: the k-th best ties with rows outside any candidate set, no certificate can hold."""
    n, dim, nq, k = 40000, 128, 32, 10
    rng = np.random.default_rng(3)
    v = rng.standard_normal((n, dim)).astype(F)
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(v)
    codes = sq.EncodeBatch(v)
    codes[5000:9000] = codes[5000]
    q = (v[5000] + 0.01 * rng.standard_normal((nq, dim))).astype(F)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales)) as ix:
        ix.upload(codes=codes)
        rows, scores, counts = ix.search(q, k)
        st = vg._lib.last_search_stats()
    # the certificate of the first pass cannot hold; the threshold pass lists all the tied rows and settles the queries
    # without the exact CUDA-core scan
    assert st["second_chance_queries"] > 0, "expected certificate failures"
    assert st["exact_rerun_queries"] == 0, st
    want = oracle_flat(q, k, dim=dim, metric=0, quant=1, codes=codes, mins=sq.mins, inv=sq.invScales)
    check(rows, scores, counts, want)
    assert np.array_equal(rows[0], np.arange(5000, 5000 + k))  # ties by row id


def test_tc_scaled_data_ranges(vg):
    """Power-of-two scaling keeps fp16 in range: values around 1e-4 and around 3e4 give the same ids as the exact scan."""
    n, dim, nq, k = 30000, 128, 16, 10
    rng = np.random.default_rng(11)
    for scale in (1e-4, 3e4):
        v = (rng.standard_normal((n, dim)) * scale).astype(F)
        q = (rng.standard_normal((nq, dim)) * scale).astype(F)
        sq = vg.quantization.ScalarQuantizer(dim)
        sq.Train(v)
        codes = sq.EncodeBatch(v)

        def make():
            ix = vg.index.DeviceIndex(codec=vg._lib.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales))
            ix.upload(codes=codes)
            return ix

        before = qtc_stats(vg)
        with make() as ix:
            rows, scores, counts = ix.search(q, k)
        after = qtc_stats(vg)
        assert after[0] - before[0] == nq and after[1] - before[1] == 0, scale
        r2, s2, c2 = exact_scan(vg, make, q, k)
        assert np.array_equal(rows, r2) and np.array_equal(bits(scores), bits(s2))


@pytest.mark.parametrize("n,dim,nq,k", [
    (20000, 1536, 33, 10),      # C4 dimension: 24 k-blocks
    (9000, 128, 300, 1),        # two k-blocks, two query tiles, ragged row tile
    (200000, 256, 32, 1000),    # rerank depth of C4: 2000 candidate groups
    (40000, 192, 20, 100),
])
def test_rabitq_tc_matches_oracle(vg, n, dim, nq, k):
    """RaBitQ estimator scan (rabitq.go:119-176): sign bits as exact +-1 fp16 operands, estimator in the epilogue."""
    rng = np.random.default_rng(n + dim + 5)
    v = (rng.standard_normal((n, dim)) * (0.5 + rng.random((n, 1)))).astype(F)   # a spread of row norms
    q = rng.standard_normal((nq, dim)).astype(F)
    v[n // 2] = v[n // 3]
    rq = vg.quantization.RaBitQuantizer(dim)
    codes = rq.EncodeBatch(v)

    def make():
        ix = vg.index.DeviceIndex(codec=vg._lib.CODEC_RABITQ, metric=0, dim=dim, rows=n)
        ix.upload(codes=codes, vectors=v)
        return ix

    before = qtc_stats(vg)
    with make() as ix:
        rows, scores, counts = ix.search(q, k)
        r2, s2, c2 = ix.search_rerank(q, k, min(10, k))
    after = qtc_stats(vg)
    assert after[0] - before[0] >= nq, "the scan did not go through the tensor-core filter"
    for i in range(min(nq, 5)):
        out = np.zeros(k, o.cand_dtype)
        c = o.lib.vgo_rabitq_search(o.fp(q[i]), o.bp(codes), n, dim, k, None, out.ctypes.data_as(C.POINTER(o.Cand)), None)
        assert c == counts[i]
        assert np.array_equal(rows[i, :c], out[:c]["row"]), i
        assert np.array_equal(bits(scores[i, :c]), bits(out[:c]["score"])), i
    vg._lib.call("vg_flat_tc_enable", 0)
    try:
        with make() as ix:
            e_rows, e_scores, e_counts = ix.search(q, k)
            e2, es2, ec2 = ix.search_rerank(q, k, min(10, k))
    finally:
        vg._lib.call("vg_flat_tc_enable", 1)
    assert np.array_equal(rows, e_rows) and np.array_equal(bits(scores), bits(e_scores)) and np.array_equal(counts, e_counts)
    assert np.array_equal(r2, e2) and np.array_equal(bits(s2), bits(es2)) and np.array_equal(c2, ec2)


@pytest.mark.parametrize("n,dim,nq,k", [
    (20000, 1536, 33, 10),      # C4 dimension
    (9000, 128, 300, 1),        # short codes: heavy ties at the k-th rank, many certificates fail -> exact re-run of those queries
    (200000, 256, 32, 1000),    # deep result lists
    (40000, 192, 20, 100),      # 6 sign words padded to 8 on the device
    (30000, 64, 40, 10),        # one k-block, Hamming distances in 0..64: ties everywhere
])
def test_bq_tc_matches_oracle(vg, n, dim, nq, k):
    """BinaryQuantizer Hamming scan (binary_quantizer.go / distance.Hamming): the +-1 GEMM gives D - 2 Hamming exactly, the
    certificate is a strict integer comparison, ties at the k-th rank go to the exact scan; ids ordered (Hamming, row)."""
    rng = np.random.default_rng(n + dim + 9)
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    v[n // 2] = v[n // 3]
    q[1] = v[7]                  # Hamming distance 0
    bq = vg.quantization.BinaryQuantizer(dim)
    codes = bq.EncodeBatch(v)
    mask_bits = rng.random(n) < 0.7
    mask = np.packbits(mask_bits, bitorder="little")

    def make():
        ix = vg.index.DeviceIndex(codec=vg._lib.CODEC_BQ, metric=0, dim=dim, rows=n, bq_threshold=0.0)
        ix.upload(codes=codes)
        return ix

    before = qtc_stats(vg)
    with make() as ix:
        rows, scores, counts = ix.search(q, k)
        mrows, mscores, mcounts = ix.search(q, k, row_mask=mask)
    after = qtc_stats(vg)
    assert after[0] - before[0] >= 2 * nq, "the scan did not go through the tensor-core filter"
    for i in range(min(nq, 6)):
        qc = bq.Encode(q[i])
        out = np.zeros(k, o.cand_dtype)
        c = o.lib.vgo_bq_search(o.bp(qc), o.bp(codes), n, codes.shape[1], k, None, out.ctypes.data_as(C.POINTER(o.Cand)))
        assert c == counts[i]
        assert np.array_equal(rows[i, :c], out[:c]["row"]), i
        assert np.array_equal(bits(scores[i, :c]), bits(out[:c]["score"])), i
    assert rows[1, 0] == 7 and scores[1, 0] == 0.0
    vg._lib.call("vg_flat_tc_enable", 0)
    try:
        with make() as ix:
            e_rows, e_scores, e_counts = ix.search(q, k)
            em_rows, em_scores, em_counts = ix.search(q, k, row_mask=mask)
    finally:
        vg._lib.call("vg_flat_tc_enable", 1)
    assert np.array_equal(rows, e_rows) and np.array_equal(bits(scores), bits(e_scores)) and np.array_equal(counts, e_counts)
    assert np.array_equal(mrows, em_rows) and np.array_equal(bits(mscores), bits(em_scores)) and np.array_equal(mcounts, em_counts)
    live = mrows[mrows != 0xFFFFFFFF]
    assert np.all(mask_bits[live])


def test_sign_codecs_rerun_only_failed_queries(vg):
    """A failed certificate of a RaBitQ / BQ query sends THAT query (with its own gathered sign words), not the batch, through
    the second chance — the threshold pass, which lists every tied row and settles it without the exact scan: duplicate rows
    beyond the candidate budget force failures for some queries only."""
    rng = np.random.default_rng(77)
    n, dim, nq, k = 60000, 128, 48, 10
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    # queries 0..7 have 400 rows at Hamming distance 0 148 rows apart, i.e. in 400 different groups (> 32 candidate groups): the
    # k-th exact distance ties tau, the certificate must fail and the exact scan must pick the 10 smallest row ids
    for j in range(8):
        v[j * 97 + 148 * np.arange(400)] = q[j]
    bq = vg.quantization.BinaryQuantizer(dim)
    codes = bq.EncodeBatch(v)
    before = qtc_stats(vg)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_BQ, metric=0, dim=dim, rows=n, bq_threshold=0.0) as ix:
        ix.upload(codes=codes)
        rows, scores, counts = ix.search(q, k)
        st = vg._lib.last_search_stats()
    after = qtc_stats(vg)
    assert after[0] - before[0] == nq
    assert 8 <= st["second_chance_queries"] < nq, "expected the planted queries (and only some queries) in the second chance"
    assert st["exact_rerun_queries"] == 0, st
    for i in range(nq):
        qc = bq.Encode(q[i])
        out = np.zeros(k, o.cand_dtype)
        c = o.lib.vgo_bq_search(o.bp(qc), o.bp(codes), n, codes.shape[1], k, None, out.ctypes.data_as(C.POINTER(o.Cand)))
        assert c == counts[i]
        assert np.array_equal(rows[i, :c], out[:c]["row"]), i
        assert np.array_equal(bits(scores[i, :c]), bits(out[:c]["score"])), i
    # RaBitQ: same planted duplicates (identical estimator values)
    rq = vg.quantization.RaBitQuantizer(dim)
    rcodes = rq.EncodeBatch(v)
    before = qtc_stats(vg)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_RABITQ, metric=0, dim=dim, rows=n) as ix:
        ix.upload(codes=rcodes, vectors=v)
        rows, scores, counts = ix.search(q, k)
        st = vg._lib.last_search_stats()
    after = qtc_stats(vg)
    assert after[0] - before[0] == nq
    assert 8 <= st["second_chance_queries"] < nq and st["exact_rerun_queries"] == 0, st
    for i in range(nq):
        out = np.zeros(k, o.cand_dtype)
        c = o.lib.vgo_rabitq_search(o.fp(q[i]), o.bp(rcodes), n, dim, k, None, out.ctypes.data_as(C.POINTER(o.Cand)), None)
        assert c == counts[i]
        assert np.array_equal(rows[i, :c], out[:c]["row"]), i
        assert np.array_equal(bits(scores[i, :c]), bits(out[:c]["score"])), i


def test_single_cta_kernels_in_subprocess():
    """The pair switches are read once per process: the single-CTA kernels (qtc_kernel, flat_tc_kernel) get their own process."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VECGO_QTC_PAIR="0", VECGO_FLAT_PAIR="0")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "tc_legacy_check.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "identical to the exact scan" in r.stdout


def test_opq_index_through_the_filter(vg):
    """OPQ = block rotation of the query (opq.go:196-214) + the PQ ADC scan: the rotated queries feed the decode-GEMM filter."""
    n, dim, m, nq, k, bs = 30000, 128, 16, 24, 10, 32
    rng = np.random.default_rng(123)
    cb, sc, of = random_pq(rng, dim, m)
    codes = rng.integers(0, 256, (n, m), dtype=np.uint8)
    rot = np.stack([np.linalg.qr(rng.standard_normal((bs, bs)))[0] for _ in range(dim // bs)]).astype(F)
    q = (rng.standard_normal((nq, dim)) * 0.7).astype(F)

    def make():
        ix = vg.index.DeviceIndex(codec=vg._lib.CODEC_OPQ, metric=0, dim=dim, rows=n, pq=(cb, sc, of, m, 256), opq=(rot, bs))
        ix.upload(codes=codes)
        return ix

    before = qtc_stats(vg)
    with make() as ix:
        rows, scores, counts = ix.search(q, k)
    assert qtc_stats(vg)[0] - before[0] == nq
    r2, s2, c2 = exact_scan(vg, make, q, k)
    assert np.array_equal(rows, r2) and np.array_equal(bits(scores), bits(s2)) and np.array_equal(counts, c2)


def test_tc_sparse_mask_returns_fewer_than_k(vg):
    """A row bitmap that keeps 5 rows with k = 10: the filter cannot certify (fewer than k scored rows), the exact scan answers."""
    n, dim, nq, k = 30000, 128, 20, 10
    rng = np.random.default_rng(8)
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(v)
    codes = sq.EncodeBatch(v)
    keep = np.zeros(n, bool)
    keep[[7, 4097, 12000, 12001, 29999]] = True
    mask = np.packbits(keep, bitorder="little")
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales)) as ix:
        ix.upload(codes=codes)
        rows, scores, counts = ix.search(q, k, row_mask=mask)
    want = oracle_flat(q, k, dim=dim, metric=0, quant=1, codes=codes, mins=sq.mins, inv=sq.invScales, mask=mask)
    check(rows, scores, counts, want)
    assert np.all(counts == 5)


def test_gather_scoring_matches_oracle(vg):
    """vg_index_score: the codec's own distance of each query to ITS candidate rows (DiskANN neighbour-list scoring,
    diskann/segment.go:511-588), bit for bit in the reference's arithmetic; rows past the end give NaN."""
    rng = np.random.default_rng(314)
    n, dim, nq, r = 3000, 768, 5, 37
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    cand = rng.integers(0, n, (nq, r)).astype(np.uint32)
    cand[0, 3] = n + 5  # out of range
    # SQ8 (Sq8uL2BatchPerDimension order)
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(v)
    codes = sq.EncodeBatch(v)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales)) as ix:
        ix.upload(codes=codes)
        got = ix.score(q, cand)
    assert np.isnan(got[0, 3])
    for i in range(nq):
        w = np.zeros(n, F)
        o.lib.vgo_sq8u_l2_batch_a512(o.fp(q[i]), o.bp(codes), o.fp(sq.mins), o.fp(sq.invScales), dim, n, o.fp(w))
        ok = cand[i] < n
        assert np.array_equal(bits(got[i][ok]), bits(w[cand[i][ok]])), i
    # INT4 (int4_avx512.c order)
    iq = vg.quantization.Int4Quantizer(dim)
    iq.Train(v)
    c4 = iq.EncodeBatch(v)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_INT4, metric=0, dim=dim, rows=n, int4=(iq.min, iq.diff)) as ix:
        ix.upload(codes=c4)
        got_lut = ix.score(q, cand)          # default: Int4Quantizer.L2Distance's live path (precomputed lookup table)
        ix.set_int4_score_mode(direct=True)  # simd.Int4L2Distance (nil-table fallback, int4.go:146)
        got = ix.score(q, cand)
    lut = np.zeros(dim * 16, F)
    o.lib.vgo_int4_build_lut(o.fp(iq.min), o.fp(iq.diff), dim, o.fp(lut))
    mine = np.zeros(dim * 16, F)
    vg._lib.call("vg_int4_build_lookup_table", vg._lib.ptr(iq.min, vg._lib.f32p), vg._lib.ptr(iq.diff, vg._lib.f32p), dim,
                 vg._lib.ptr(mine, vg._lib.f32p))
    assert np.array_equal(bits(mine), bits(lut))  # simd.BuildInt4LookupTable, kernels.go:94-103
    for i in range(nq):
        for j in range(r):
            if cand[i, j] < n:
                w = o.lib.vgo_int4_l2_a512(o.fp(q[i]), o.bp(c4[cand[i, j]]), dim, o.fp(iq.min), o.fp(iq.diff))
                assert bits(got[i, j]) == bits(F(w)), (i, j)
                w = o.lib.vgo_int4_l2_precomputed_a512(o.fp(q[i]), o.bp(c4[cand[i, j]]), dim, o.fp(lut))
                assert bits(got_lut[i, j]) == bits(F(w)), (i, j)
    # INT4 at an odd, non-multiple-of-16 dimension (row-major layout, scalar tail of the precomputed kernel)
    d2 = 131
    v2 = rng.standard_normal((500, d2)).astype(F)
    iq2 = vg.quantization.Int4Quantizer(d2)
    iq2.Train(v2)
    c42 = iq2.EncodeBatch(v2)
    q2 = rng.standard_normal((3, d2)).astype(F)
    cand2 = rng.integers(0, 500, (3, 11)).astype(np.uint32)
    lut2 = np.zeros(d2 * 16, F)
    o.lib.vgo_int4_build_lut(o.fp(iq2.min), o.fp(iq2.diff), d2, o.fp(lut2))
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_INT4, metric=0, dim=d2, rows=500, int4=(iq2.min, iq2.diff)) as ix:
        ix.upload(codes=c42)
        got2 = ix.score(q2, cand2)
    for i in range(3):
        for j in range(11):
            w = o.lib.vgo_int4_l2_precomputed_a512(o.fp(q2[i]), o.bp(c42[cand2[i, j]]), d2, o.fp(lut2))
            assert bits(got2[i, j]) == bits(F(w)), (i, j)
    # BQ (Hamming distance as float32: binary.go score = float32(popcount))
    bq = vg.quantization.BinaryQuantizer(dim)
    bq.Train(v)
    bc = bq.EncodeBatch(v)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_BQ, metric=0, dim=dim, rows=n, bq_threshold=float(bq.threshold)) as ix:
        ix.upload(codes=bc)
        gotb = ix.score(q, cand)
    qb = bq.EncodeBatch(q)
    for i in range(nq):
        for j in range(r):
            if cand[i, j] < n:
                h = int(np.unpackbits(np.bitwise_xor(qb[i], bc[cand[i, j]])).sum())
                assert gotb[i, j] == F(h), (i, j)
    assert np.isnan(gotb[0, 3])
    # PQ (generic table build + pqAdcLookupAvx512 order), tiled layout
    m = 96
    cb, sc, of = random_pq(rng, dim, m)
    pc = rng.integers(0, 256, (n, m), dtype=np.uint8)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_PQ, metric=0, dim=dim, rows=n, pq=(cb, sc, of, m, 256)) as ix:
        ix.upload(codes=pc)
        got = ix.score(q, cand)
    for i in range(nq):
        tab = np.zeros(m * 256, F)
        o.lib.vgo_pq_build_table(o.fp(q[i]), dim, m, 256, cb.ctypes.data_as(o.i8p), o.fp(sc), o.fp(of), o.fp(tab))
        for j in range(r):
            if cand[i, j] < n:
                w = o.lib.vgo_pq_adc_a512(o.fp(tab), o.bp(pc[cand[i, j]]), m)
                assert bits(got[i, j]) == bits(F(w)), (i, j)
    # RaBitQ (exact popcount + unfused Go estimator)
    rq = vg.quantization.RaBitQuantizer(dim)
    rc = rq.EncodeBatch(v)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_RABITQ, metric=0, dim=dim, rows=n) as ix:
        ix.upload(codes=rc)
        got = ix.score(q, cand)
    for i in range(nq):
        for j in range(r):
            if cand[i, j] < n:
                w = o.lib.vgo_rabitq_distance(o.fp(q[i]), dim, o.bp(rc[cand[i, j]]))
                assert bits(got[i, j]) == bits(F(w)), (i, j)
