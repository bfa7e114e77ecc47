"""GPU parity tests: every call goes through the C ABI (vecgo_b200/libvecgo_cuda.so)
and is compared with the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): bit-exact codes, popcounts and top-k row ids
(ties by row id); float32 distances within 1e-4 relative — in practice the
kernels reproduce the reference's AVX-512 summation order, so distances are
compared BIT-FOR-BIT and the 1e-4 bound is only the documented fallback.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as o

pytestmark = pytest.mark.gpu
F = np.float32
LENGTHS = [1, 3, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 128, 129, 191, 256, 768]


def bits(x):
    return np.ascontiguousarray(x, F).view(np.uint32)


@pytest.fixture(scope="module")
def vg():
    import vecgo_b200

    return vecgo_b200


# ----------------------------------------------------------------- simd table
def test_pair_kernels_bit_exact(vg):
    rng = np.random.default_rng(1)
    for n in LENGTHS + [1536]:
        a = rng.standard_normal((6, n)).astype(F)
        b = rng.standard_normal((6, n)).astype(F)
        want_d = np.array([o.lib.vgo_dot_a512(o.fp(a[i]), o.fp(b[i]), n) for i in range(6)], F)
        want_l = np.array([o.lib.vgo_sql2_a512(o.fp(a[i]), o.fp(b[i]), n) for i in range(6)], F)
        assert np.array_equal(bits(vg.simd.DotPairs(a, b)), bits(want_d)), n
        assert np.array_equal(bits(vg.simd.SquaredL2Pairs(a, b)), bits(want_l)), n


def test_simd_kat(vg):  # internal/simd/floats_test.go:11-33,55-74
    assert vg.simd.Dot(np.arange(1, 17), np.arange(1, 17)) == 1496.0
    assert vg.simd.Dot([1, -2, 3], [-4, 5, -6]) == -32.0
    assert vg.simd.SquaredL2([1, -2, 3], [-4, 5, -6]) == 155.0
    assert vg.simd.Dot(np.zeros(0, F), np.zeros(0, F)) == 0.0
    a = np.array([1, -2, 3], F)
    vg.simd.ScaleInPlace(a, -1.0)
    assert a.tolist() == [-1.0, 2.0, -3.0]
    assert vg.simd.Hamming(list(range(17)), [0xFF, 1, 0xFD, 3, 0xFB, 5, 0xF9, 7, 0xF7, 9, 0xF5, 0xB, 0xF3, 0xD, 0xF1, 0xF, 0xEF]) == 72
    table = np.array([[i * 1000 + j for j in range(256)] for i in range(16)], F).ravel()
    codes = np.array([17 * i for i in range(16)], np.uint8)
    assert vg.simd.PqAdcLookup(table, codes, 16) == sum(table[i * 256 + int(codes[i])] for i in range(16))


def test_batch_kernels_bit_exact(vg):
    rng = np.random.default_rng(2)
    for dim in LENGTHS:
        q = rng.standard_normal((5, dim)).astype(F)
        t = rng.standard_normal((37, dim)).astype(F)
        for mine, theirs in ((vg.simd.SquaredL2Batch, o.lib.vgo_sql2_batch_a512), (vg.simd.DotBatch, o.lib.vgo_dot_batch_a512)):
            got = mine(q, t, dim)
            for i in range(5):
                want = np.zeros(37, F)
                theirs(o.fp(q[i]), o.fp(t), dim, 37, o.fp(want))
                assert np.array_equal(bits(got[i]), bits(want)), dim


def test_sq8_kernel_bit_exact(vg):
    rng = np.random.default_rng(3)
    for dim in [1, 7, 8, 15, 16, 17, 31, 32, 33, 64, 100, 128, 256, 768]:
        q = (rng.random((9, dim)) * 2 - 1).astype(F)
        mins = (rng.random(dim) * 2 - 1).astype(F)
        inv = (rng.random(dim) * 0.02).astype(F)
        codes = rng.integers(0, 256, (70, dim), dtype=np.uint8)
        got = vg.simd.Sq8uL2BatchPerDimension(q, codes, mins, inv, dim)
        for i in range(9):
            want = np.zeros(70, F)
            o.lib.vgo_sq8u_l2_batch_a512(o.fp(q[i]), o.bp(codes), o.fp(mins), o.fp(inv), dim, 70, o.fp(want))
            assert np.array_equal(bits(got[i]), bits(want)), dim


def test_int4_kernel_bit_exact(vg):
    rng = np.random.default_rng(4)
    for dim in [2, 7, 16, 30, 32, 33, 64, 66, 96, 128, 130, 256, 768]:
        cs = (dim + 1) // 2
        q = rng.standard_normal((9, dim)).astype(F)
        minv = rng.standard_normal(dim).astype(F)
        diff = (rng.random(dim) * 3 + 0.1).astype(F)
        codes = rng.integers(0, 256, (41, cs), dtype=np.uint8)
        got = vg.simd.Int4L2DistanceBatch(q, codes, dim, 41, minv, diff)
        for i in range(9):
            want = np.zeros(41, F)
            o.lib.vgo_int4_l2_batch_a512(o.fp(q[i]), o.bp(codes), dim, 41, o.fp(minv), o.fp(diff), o.fp(want))
            assert np.array_equal(bits(got[i]), bits(want)), dim


def test_pq_adc_and_hamming_bit_exact(vg):
    rng = np.random.default_rng(5)
    for m in [1, 2, 7, 8, 15, 16, 17, 32, 48, 96, 100]:
        tables = (rng.random((3, m * 256)) * 2 - 1).astype(F)
        codes = rng.integers(0, 256, (50, m), dtype=np.uint8)
        got = vg.simd.PqAdcLookupBatch(tables, codes, m)
        for i in range(3):
            want = np.array([o.lib.vgo_pq_adc_a512(o.fp(tables[i]), o.bp(codes[j]), m) for j in range(50)], F)
            assert np.array_equal(bits(got[i]), bits(want)), m
    for n in [1, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 192]:
        a = rng.integers(0, 256, (4, n), dtype=np.uint8)
        b = rng.integers(0, 256, (33, n), dtype=np.uint8)
        got = vg.simd.HammingBatch(a, b, n)
        want = np.array([[o.lib.vgo_hamming(o.bp(a[i]), o.bp(b[j]), n) for j in range(33)] for i in range(4)])
        assert np.array_equal(got, want)


def test_normalize(vg):
    rng = np.random.default_rng(6)
    for dim in (3, 64, 100, 768):
        v = rng.standard_normal((5, dim)).astype(F)
        v[2] = 0
        got, ok = vg.distance.NormalizeL2Batch(v)
        for i in range(5):
            w = v[i].copy()
            r = o.lib.vgo_normalize_l2(o.fp(w), dim)
            assert bool(r) == bool(ok[i])
            assert np.array_equal(bits(got[i]), bits(w))


# ----------------------------------------------------------------- quantizers
def test_sq8_quantizer_bit_exact(vg):
    rng = np.random.default_rng(7)
    n, dim = 3000, 100
    v = rng.standard_normal((n, dim)).astype(F)
    v[:, 5] = 0.25  # constant dimension: min==max → max = min + 1e-6
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(v)
    mins, maxs, sc, inv = (np.zeros(dim, F) for _ in range(4))
    o.lib.vgo_sq8_train(o.fp(v), n, dim, o.fp(mins), o.fp(maxs), o.fp(sc), o.fp(inv))
    for a, b in ((sq.mins, mins), (sq.maxs, maxs), (sq.scales, sc), (sq.invScales, inv)):
        assert np.array_equal(bits(a), bits(b))
    x = (rng.standard_normal((500, dim)) * 1.5).astype(F)  # some values outside [min,max] → clamp
    got = sq.EncodeBatch(x)
    want = np.zeros_like(got)
    for i in range(len(x)):
        o.lib.vgo_sq8_encode(o.fp(x[i]), dim, o.fp(mins), o.fp(maxs), o.fp(sc), o.bp(want[i]))
    assert np.array_equal(got, want)
    dec = sq.DecodeBatch(got)
    wdec = np.zeros_like(dec)
    for i in range(len(x)):
        o.lib.vgo_sq8_decode(o.bp(want[i]), dim, o.fp(mins), o.fp(inv), o.fp(wdec[i]))
    assert np.array_equal(bits(dec), bits(wdec))
    with pytest.raises(Exception):
        vg.quantization.ScalarQuantizer(4).Encode(np.zeros(4, F))  # not trained
    with pytest.raises(Exception):
        sq.Encode(np.zeros(dim + 1, F))  # dimension mismatch
    sq2 = vg.quantization.ScalarQuantizer(1)
    sq2.UnmarshalBinary(sq.MarshalBinary())
    assert np.array_equal(bits(sq2.invScales), bits(sq.invScales))


def test_int4_bq_rabitq_encode_bit_exact(vg):
    rng = np.random.default_rng(8)
    n, dim = 800, 131
    v = rng.standard_normal((n, dim)).astype(F)
    v[:, 3] = 1.0
    iq = vg.quantization.Int4Quantizer(dim)
    iq.Train(v)
    minv, diff = np.zeros(dim, F), np.zeros(dim, F)
    o.lib.vgo_int4_train(o.fp(v), n, dim, o.fp(minv), o.fp(diff))
    assert np.array_equal(bits(iq.min), bits(minv)) and np.array_equal(bits(iq.diff), bits(diff))
    x = (rng.standard_normal((300, dim)) * 1.3).astype(F)
    got = iq.EncodeBatch(x)
    want = np.zeros_like(got)
    for i in range(len(x)):
        o.lib.vgo_int4_encode(o.fp(x[i]), dim, o.fp(minv), o.fp(diff), o.bp(want[i]))
    assert np.array_equal(got, want)
    dec = iq.DecodeBatch(got)
    wdec = np.zeros_like(dec)
    for i in range(len(x)):
        o.lib.vgo_int4_decode(o.bp(want[i]), dim, o.fp(minv), o.fp(diff), o.fp(wdec[i]))
    assert np.array_equal(bits(dec), bits(wdec))
    # BQ
    bq = vg.quantization.BinaryQuantizer(dim)
    bq.Train(v)
    assert bq.Threshold() == o.lib.vgo_bq_train(o.fp(v), n, dim)
    gb = bq.EncodeBatch(x)
    wb = np.zeros_like(gb)
    for i in range(len(x)):
        o.lib.vgo_bq_encode(o.fp(x[i]), dim, float(bq.Threshold()), o.bp(wb[i]))
    assert np.array_equal(gb, wb)
    e = vg.quantization.BinaryQuantizer(128).EncodeUint64(np.array([1.0 if i % 2 == 0 else -1.0 for i in range(128)], F))
    assert e[0] == 0x5555555555555555 and e[1] == 0x5555555555555555  # binary_test.go:9-39
    # RaBitQ: sign bits exact, norm bit-exact w.r.t. the AVX-512 dot order
    for d in (131, 1536):
        xx = rng.standard_normal((64, d)).astype(F)
        rq = vg.quantization.RaBitQuantizer(d)
        gr = rq.EncodeBatch(xx)
        wr = np.zeros_like(gr)
        for i in range(len(xx)):
            o.lib.vgo_rabitq_encode(o.fp(xx[i]), d, o.bp(wr[i]))
        assert np.array_equal(gr, wr)


def _random_pq(rng, dim, m, k=256):
    ds = dim // m
    cb = rng.integers(-128, 128, m * k * ds, dtype=np.int8)
    sc = (rng.random(m) * 0.02 + 0.005).astype(F)
    of = (rng.standard_normal(m) * 0.1).astype(F)
    return cb, sc, of


def test_pq_encode_table_decode_bit_exact(vg):
    rng = np.random.default_rng(9)
    for dim, m in ((64, 8), (96, 96), (768, 96), (100, 20)):
        cb, sc, of = _random_pq(rng, dim, m)
        pq = vg.quantization.ProductQuantizer(dim, m, 256)
        pq.SetCodebooks(cb, sc, of)
        x = rng.standard_normal((200, dim)).astype(F)
        got = pq.EncodeBatch(x)
        want = np.zeros_like(got)
        cbp = cb.ctypes.data_as(o.i8p)
        for i in range(len(x)):
            o.lib.vgo_pq_encode(o.fp(x[i]), dim, m, 256, cbp, o.fp(sc), o.fp(of), o.bp(want[i]))
        assert np.array_equal(got, want), (dim, m)
        tabs = pq.BuildDistanceTable(x[:7])
        for i in range(7):
            w = np.zeros(m * 256, F)
            o.lib.vgo_pq_build_table(o.fp(x[i]), dim, m, 256, cbp, o.fp(sc), o.fp(of), o.fp(w))
            assert np.array_equal(bits(tabs[i]), bits(w))
        dec = pq.DecodeBatch(got[:20])
        for i in range(20):
            w = np.zeros(dim, F)
            o.lib.vgo_pq_decode(o.bp(want[i]), dim, m, 256, cbp, o.fp(sc), o.fp(of), o.fp(w))
            assert np.array_equal(bits(dec[i]), bits(w))
    with pytest.raises(Exception):
        vg.quantization.ProductQuantizer(10, 3, 256)
    with pytest.raises(Exception):
        vg.quantization.ProductQuantizer(12, 3, 257)


# ----------------------------------------------------------------- top-k scans
def _check_topk(rows, scores, counts, want, k):
    nq = rows.shape[0]
    for i in range(nq):
        c = int(counts[i])
        w = want[i]
        assert c == len(w), (i, c, len(w))
        assert np.array_equal(rows[i, :c], w["row"]), i
        assert np.array_equal(bits(scores[i, :c]), bits(w["score"])), i
        assert np.all(rows[i, c:] == 0xFFFFFFFF)


def _oracle_flat(queries, k, **kw):
    seg = o.FlatOracle(**kw)
    out, cnt = seg.search_batch(queries, k, threads=8)
    return [out[i, : cnt[i]] for i in range(len(queries))]


@pytest.mark.parametrize("metric", [0, 2])
@pytest.mark.parametrize("n,dim,nq,k", [(1000, 128, 13, 10), (5000, 100, 5, 100), (257, 65, 3, 300), (40000, 32, 600, 10)])
def test_flat_f32_search(vg, metric, n, dim, nq, k):
    rng = np.random.default_rng(n + dim)
    x = rng.random((n, dim)).astype(F)
    q = rng.random((nq, dim)).astype(F)
    if metric == 2:
        x, _ = vg.distance.NormalizeL2Batch(x)
    x[n // 2] = x[n // 3]  # exact duplicate rows → equal scores → tie broken by row id
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_F32, metric=metric, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        rows, scores, counts = ix.search(q, k)
    _check_topk(rows, scores, counts, _oracle_flat(q, k, dim=dim, metric=metric, vectors=x), k)


def test_flat_config1_full(vg):
    """BASELINE configs[0]: Flat exact L2, 100k x 128 U[0,1), 1k queries, k=10."""
    n, dim, nq, k = 100_000, 128, 1000, 10
    x = np.random.default_rng(42).random((n, dim), dtype=F)
    q = np.random.default_rng(43).random((nq, dim), dtype=F)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_F32, metric=0, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        rows, scores, counts = ix.search(q, k)
        rr = ix.rerank(q[:50], rows[:50])
    want = _oracle_flat(q, k, dim=dim, metric=0, vectors=x)
    _check_topk(rows, scores, counts, want, k)
    assert np.array_equal(bits(rr), bits(scores[:50]))  # Rerank recomputes the same exact scores


@pytest.mark.parametrize("n,dim,nq,k", [(3000, 768, 20, 100), (2000, 128, 9, 10), (1500, 100, 4, 10), (700, 256, 3, 50)])
def test_sq8_search(vg, n, dim, nq, k):
    rng = np.random.default_rng(dim)
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(v)
    codes = sq.EncodeBatch(v)
    codes[7] = codes[3]
    for metric in (0, 2):  # L2 → Sq8uL2BatchPerDimension; dot → scalar Go sq.DotProduct (flat/segment.go:672-689)
        with vg.index.DeviceIndex(codec=vg._lib.CODEC_SQ8, metric=metric, dim=dim, rows=n, sq8=(sq.mins, sq.invScales)) as ix:
            ix.upload(codes=codes)
            rows, scores, counts = ix.search(q, k)
        want = _oracle_flat(q, k, dim=dim, metric=metric, quant=1, codes=codes, mins=sq.mins, inv=sq.invScales)
        _check_topk(rows, scores, counts, want, k)


@pytest.mark.parametrize("n,dim,nq,k", [(3000, 768, 20, 100), (1000, 96, 5, 10), (900, 131, 4, 10)])
def test_int4_search(vg, n, dim, nq, k):
    rng = np.random.default_rng(dim + 1)
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    iq = vg.quantization.Int4Quantizer(dim)
    iq.Train(v)
    codes = iq.EncodeBatch(v)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_INT4, metric=0, dim=dim, rows=n, int4=(iq.min, iq.diff)) as ix:
        ix.upload(codes=codes)
        rows, scores, counts = ix.search(q, k)
    want = []
    for i in range(nq):
        out = np.zeros(k, o.cand_dtype)
        c = o.lib.vgo_int4_search(o.fp(q[i]), o.bp(codes), n, dim, o.fp(iq.min), o.fp(iq.diff), k,
                                  o.fn_addr(o.lib.vgo_int4_l2_batch_a512), out.ctypes.data_as(C.POINTER(o.Cand)))
        want.append(out[:c])
    _check_topk(rows, scores, counts, want, k)


@pytest.mark.parametrize("n,dim,m,nq,k", [(4000, 768, 96, 11, 100), (1000, 64, 8, 5, 10), (800, 100, 20, 3, 10)])
def test_pq_adc_search(vg, n, dim, m, nq, k):
    rng = np.random.default_rng(m)
    cb, sc, of = _random_pq(rng, dim, m)
    codes = rng.integers(0, 256, (n, m), dtype=np.uint8)
    codes[11] = codes[2]
    q = rng.standard_normal((nq, dim)).astype(F)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_PQ, metric=0, dim=dim, rows=n, pq=(cb, sc, of, m, 256)) as ix:
        ix.upload(codes=codes)
        rows, scores, counts = ix.search(q, k)
    want = _oracle_flat(q, k, dim=dim, metric=0, quant=2, codes=codes, pq=(cb, sc, of, m, 256))
    _check_topk(rows, scores, counts, want, k)


@pytest.mark.parametrize("n,dim,nq,k", [(5000, 1536, 10, 1000), (2000, 128, 7, 10), (600, 100, 3, 10)])
def test_rabitq_and_bq_search(vg, n, dim, nq, k):
    rng = np.random.default_rng(dim + 2)
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    rq = vg.quantization.RaBitQuantizer(dim)
    codes = rq.EncodeBatch(v)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_RABITQ, metric=0, dim=dim, rows=n) as ix:
        ix.upload(codes=codes, vectors=v)
        rows, scores, counts = ix.search(q, k)
        r2, s2, c2 = ix.search_rerank(q, k, min(10, k))
    want, exact = [], []
    for i in range(nq):
        out = np.zeros(k, o.cand_dtype)
        c = o.lib.vgo_rabitq_search(o.fp(q[i]), o.bp(codes), n, dim, k, None, out.ctypes.data_as(C.POINTER(o.Cand)), None)
        want.append(out[:c])
        # engine refine: exact L2 of the approx top-k, best min(10,k) by (score,row)
        ex = np.array([o.lib.vgo_sql2_a512(o.fp(q[i]), o.fp(v[r]), dim) for r in out[:c]["row"]], F)
        order = np.lexsort((out[:c]["row"], ex))[: min(10, k)]
        exact.append((out[:c]["row"][order], ex[order]))
    _check_topk(rows, scores, counts, want, k)
    for i in range(nq):
        assert np.array_equal(r2[i, : c2[i]], exact[i][0]) and np.array_equal(bits(s2[i, : c2[i]]), bits(exact[i][1]))
    # BQ: Hamming score, ties (many!) broken by row id
    bq = vg.quantization.BinaryQuantizer(dim)
    bcodes = bq.EncodeBatch(v)
    kk = min(k, 50)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_BQ, metric=3, dim=dim, rows=n, bq_threshold=0.0) as ix:
        ix.upload(codes=bcodes)
        rows, scores, counts = ix.search(q, kk)
    want = []
    for i in range(nq):
        qc = bq.Encode(q[i])
        out = np.zeros(kk, o.cand_dtype)
        c = o.lib.vgo_bq_search(o.bp(qc), o.bp(bcodes), n, bcodes.shape[1], kk, None, out.ctypes.data_as(C.POINTER(o.Cand)))
        want.append(out[:c])
    _check_topk(rows, scores, counts, want, kk)


def test_row_mask_and_small_k_edge_cases(vg):
    rng = np.random.default_rng(20)
    n, dim = 777, 48
    x = rng.random((n, dim)).astype(F)
    q = rng.random((4, dim)).astype(F)
    mask_bits = rng.random(n) < 0.3
    mask = np.packbits(mask_bits, bitorder="little")
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_F32, metric=0, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        rows, scores, counts = ix.search(q, 1000, row_mask=mask)  # k > matching rows
        seg = o.FlatOracle(dim=dim, metric=0, vectors=x)
        out, cnt = seg.search_batch(q, 1000, mask=mask)
        _check_topk(rows, scores, counts, [out[i, : cnt[i]] for i in range(4)], 1000)
        assert int(counts[0]) == int(mask_bits.sum())
        rows, scores, counts = ix.search(q, 1)
        out, cnt = seg.search_batch(q, 1)
        _check_topk(rows, scores, counts, [out[i, : cnt[i]] for i in range(4)], 1)
        with pytest.raises(vg.VecgoError):
            ix.search(q, 0)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_F32, metric=0, dim=dim, rows=0) as ix:  # empty segment
        rows, scores, counts = ix.search(q, 5)
        assert counts.tolist() == [0, 0, 0, 0] and np.all(rows == 0xFFFFFFFF)


def test_topk_merge(vg):
    rng = np.random.default_rng(21)
    lists, nq, k = 8, 33, 100
    scores = np.sort(rng.integers(0, 500, (lists, nq, k)).astype(F), axis=2)
    rows = rng.permutation(lists * nq * k).astype(np.uint32).reshape(lists, nq, k)
    rows[3, :, 90:] = 0xFFFFFFFF  # short list
    for desc in (False, True):
        s = -scores if desc else scores
        orow, osc, ocnt = vg.index.topk_merge(rows, s, desc, 100)
        for qi in range(nq):
            r = rows[:, qi, :].ravel()
            sc = s[:, qi, :].ravel()
            live = r != 0xFFFFFFFF
            r, sc = r[live], sc[live]
            order = np.lexsort((r, -sc if desc else sc))[:100]
            assert ocnt[qi] == 100
            assert np.array_equal(orow[qi], r[order]) and np.array_equal(osc[qi], sc[order])


# ----------------------------------------------------------------- flat segment files
def test_flat_file_roundtrip(vg):
    rng = np.random.default_rng(30)
    n, dim = 2500, 64
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((6, dim)).astype(F)
    ids = (np.arange(n, dtype=np.uint64) * 7 + 1000)
    for quant in (vg.flat.QuantizationNone, vg.flat.QuantizationSQ8):
        data = vg.flat.write_segment(segment_id=9, vectors=v, ids=ids, metric=0, quantization=quant)
        seg = vg.flat.Segment.Open(data)
        assert seg.RowCount() == n and seg.ID() == 9
        rows, scores, counts = seg.Search(q, 10)
        if quant == vg.flat.QuantizationNone:
            want = _oracle_flat(q, 10, dim=dim, metric=0, vectors=v, segment_id=9)
        else:
            hdr_off = int.from_bytes(data[56:64], "little")
            mins = np.frombuffer(data, "<f4", dim, hdr_off)
            maxs = np.frombuffer(data, "<f4", dim, hdr_off + 4 * dim)
            sc, inv = np.zeros(dim, F), np.zeros(dim, F)
            o.lib.vgo_sq8_set_bounds(o.fp(mins.copy()), o.fp(maxs.copy()), dim, o.fp(sc), o.fp(inv))
            codes = np.frombuffer(data, np.uint8, n * dim, int.from_bytes(data[64:72], "little")).reshape(n, dim)
            want = _oracle_flat(q, 10, dim=dim, metric=0, quant=1, codes=codes, mins=mins, inv=inv)
        _check_topk(rows, scores, counts, want, 10)
        rr = seg.Rerank(q, rows)
        exact = np.array([[o.lib.vgo_sql2_a512(o.fp(q[i]), o.fp(v[r]), dim) for r in rows[i]] for i in range(6)], F)
        assert np.array_equal(bits(rr), bits(exact))
        assert np.array_equal(seg.FetchIDs(rows[0]), ids[rows[0]])
        seg.Close()
        bad = bytearray(data)
        bad[-9] ^= 0x40
        with pytest.raises(vg.VecgoError):
            vg.flat.Segment.Open(bytes(bad))  # checksum mismatch (flat/checksum_test.go)
        with pytest.raises(vg.VecgoError):
            vg.flat.Segment.Open(b"\x00" * 200)  # invalid magic
        with pytest.raises(vg.VecgoError):
            vg.flat.Segment.Open(data[: len(data) // 2], verify_checksum=False)  # truncated


def test_flat_pq_segment_finds_zero_vector(vg):
    """flat/pq_test.go:17-93: vec[j] = (i+j)*0.01, query = zeros → row 0 first."""
    n, dim = 1000, 32
    v = np.array([[(i + j) * 0.01 for j in range(dim)] for i in range(n)], F)
    data = vg.flat.write_segment(segment_id=1, vectors=v, metric=0, quantization=vg.flat.QuantizationPQ, pq_m=8, seed=3)
    seg = vg.flat.Segment.Open(data)
    rows, scores, counts = seg.Search(np.zeros((1, dim), F), 5)
    assert rows[0, 0] == 0 and counts[0] == 5
    seg.Close()


def test_partitioned_segment_nprobe(vg):
    rng = np.random.default_rng(31)
    n, dim, P = 4000, 32, 8
    v = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((9, dim)).astype(F)
    data = vg.flat.write_segment(segment_id=2, vectors=v, metric=0, k_partitions=P, seed=5)
    hdr = vg.flat.decode_header(data)
    assert hdr["num_partitions"] == P
    cent = np.frombuffer(data, "<f4", P * dim, int.from_bytes(data[40:48], "little")).reshape(P, dim)
    poff = np.frombuffer(data, "<u4", P + 1, int.from_bytes(data[48:56], "little"))
    vv = np.frombuffer(data, "<f4", n * dim, int.from_bytes(data[72:80], "little")).reshape(n, dim)
    seg = vg.flat.Segment.Open(data)
    for nprobe in (1, 3, 8):
        rows, scores, counts = seg.Search(q, 10, nprobes=nprobe)
        so = o.FlatOracle(dim=dim, metric=0, vectors=vv, centroids=cent, partition_offsets=poff, segment_id=2)
        out, cnt = so.search_batch(q, 10, nprobes=nprobe)
        _check_topk(rows, scores, counts, [out[i, : cnt[i]] for i in range(9)], 10)
    seg.Close()


# ----------------------------------------------------------------- training
def test_kmeans_train_matches_oracle(vg):
    rng = np.random.default_rng(40)
    n, dim, k = 3000, 24, 16
    v = (rng.standard_normal((n, dim)) + rng.integers(0, 4, (n, 1)) * 3).astype(F)
    init = rng.permutation(n)[:k].astype(np.int64)
    for metric in (0, 2):
        cent, assign, iters = vg.kmeans.TrainKMeans(v, dim, k, metric, 10, init_rows=init, seed=11, return_assign=True)
        wc = np.zeros((k, dim), F)
        wa = np.zeros(n, np.int32)
        wi = o.lib.vgo_kmeans_train(o.fp(v), n, dim, k, metric, 10, init.ctypes.data_as(o.i64p), 11, o.fp(wc),
                                    wa.ctypes.data_as(o.i32p))
        assert iters == wi
        assert np.array_equal(assign, wa)
        assert np.array_equal(bits(cent), bits(wc))
    a = vg.kmeans.AssignPartition(v[:100], wc, dim, 0)
    assert a.tolist() == [o.lib.vgo_kmeans_assign(o.fp(v[i]), o.fp(wc), dim, k, 0) for i in range(100)]
    got = vg.kmeans.FindClosestCentroids(v[:20], wc, dim, 3, 0)
    for i in range(20):
        w = np.zeros(3, np.int64)
        o.lib.vgo_find_closest_centroids(o.fp(v[i]), o.fp(wc), dim, k, 3, 0, w.ctypes.data_as(o.i64p))
        assert got[i].tolist() == w.tolist()


def test_pq_train_matches_oracle(vg):
    rng = np.random.default_rng(41)
    n, dim, m, k = 2000, 32, 4, 256
    v = rng.standard_normal((n, dim)).astype(F)
    pq = vg.quantization.ProductQuantizer(dim, m, k)
    pq.Train(v, iters=6, seed=77)
    ds = dim // m
    for s in range(m):
        cent = np.zeros((k, ds), F)
        o.lib.vgo_pq_kmeanspp_init(o.fp(v), n, dim, s * ds, ds, k, 77, s, o.fp(cent))
        assign = np.zeros(n, np.int32)
        o.lib.vgo_pq_lloyd(o.fp(v), n, dim, s * ds, ds, k, 6, 77, s, o.fp(cent), assign.ctypes.data_as(o.i32p))
        assert np.array_equal(bits(pq.centroids_f32[s]), bits(cent)), s
        cb = np.zeros(k * ds, np.int8)
        sc, of = np.zeros(1, F), np.zeros(1, F)
        o.lib.vgo_pq_quantize_centroids(o.fp(cent), k * ds, cb.ctypes.data_as(o.i8p), o.fp(sc), o.fp(of))
        assert np.array_equal(pq.codebooks[s * k * ds:(s + 1) * k * ds], cb)
        assert bits(pq.scales[s]) == bits(sc[0]) and bits(pq.offsets[s]) == bits(of[0])
    # pq_test.go:59-71: reconstruction MSE < 0.5 on N(0,1) data
    rec = pq.DecodeBatch(pq.EncodeBatch(v[:200]))
    assert float(np.mean((rec - v[:200]) ** 2)) < 0.5


# ----------------------------------------------------------------- OPQ (opq.go, svd.go)
def test_opq_procrustes_matches_oracle(vg):
    """computeProcrustesRotation: bit-exact with the restatement; R orthogonal with det +1 (opq_test.go:56-99)."""
    rng = np.random.default_rng(51)
    for n in (8, 32, 48):
        blocks = 3
        M = rng.standard_normal((blocks, n, n)).astype(F)
        M[1] = M[1] @ np.diag(np.linspace(1, -1, n)).astype(F)  # a block whose plain U V^T is a reflection
        R = np.zeros_like(M)
        sig = np.zeros((blocks, n), F)
        L = vg._lib
        L.call("vg_opq_procrustes", L.ptr(M, L.f32p), blocks, n, L.ptr(R, L.f32p), L.ptr(sig, L.f32p))
        for b in range(blocks):
            m = M[b].copy()
            want = np.zeros((n, n), F)
            ws = np.zeros(n, F)
            o.lib.vgo_opq_procrustes(o.fp(m), n, o.fp(want), o.fp(ws), None)
            assert np.array_equal(bits(R[b]), bits(want)), (n, b)
            assert np.array_equal(bits(sig[b]), bits(ws)), (n, b)
            assert np.allclose(R[b] @ R[b].T, np.eye(n), atol=1e-3)
            assert np.linalg.det(R[b].astype(np.float64)) > 0.9


@pytest.mark.parametrize("n,dim,m,rounds", [(600, 32, 8, 3), (500, 96, 12, 2)])
def test_opq_train_matches_oracle(vg, n, dim, m, rounds):
    """OptimizedProductQuantizer.Train against the restatement composed round by round (opq.go:89-193)."""
    k, pq_iters, seed = 256, 4, 99
    rng = np.random.default_rng(n + dim)
    v = (rng.random((n, dim)) * 2 - 1).astype(F)           # UniformRangeVectors as opq_test.go
    opq = vg.quantization.OptimizedProductQuantizer(dim, m, k, rounds)
    opq.Train(v, pq_iters=pq_iters, seed=seed)
    bs = int(o.lib.vgo_opq_block_size(dim, m))
    assert opq.blockSize == bs == (32 if dim > 64 else dim)
    ds, nb = dim // m, dim // bs
    rot = np.tile(np.eye(bs, dtype=F), (nb, 1, 1))
    cb = np.zeros(m * k * ds, np.int8)
    sc, of = np.zeros(m, F), np.zeros(m, F)
    for it in range(rounds):
        xr = np.zeros_like(v)
        for i in range(n):
            o.lib.vgo_opq_rotate(o.fp(v[i]), dim, bs, o.fp(rot), o.fp(xr[i]))
        for s in range(m):
            cent = np.zeros((k, ds), F)
            o.lib.vgo_pq_kmeanspp_init(o.fp(xr), n, dim, s * ds, ds, k, seed + it, s, o.fp(cent))
            assign = np.zeros(n, np.int32)
            o.lib.vgo_pq_lloyd(o.fp(xr), n, dim, s * ds, ds, k, pq_iters, seed + it, s, o.fp(cent), assign.ctypes.data_as(o.i32p))
            s1, o1 = np.zeros(1, F), np.zeros(1, F)
            o.lib.vgo_pq_quantize_centroids(o.fp(cent), k * ds, cb[s * k * ds:].ctypes.data_as(o.i8p), o.fp(s1), o.fp(o1))
            sc[s], of[s] = s1[0], o1[0]
        yh = np.zeros_like(v)
        codes = np.zeros(m, np.uint8)
        for i in range(n):
            o.lib.vgo_pq_encode(o.fp(xr[i]), dim, m, k, cb.ctypes.data_as(o.i8p), o.fp(sc), o.fp(of), o.bp(codes))
            o.lib.vgo_pq_decode(o.bp(codes), dim, m, k, cb.ctypes.data_as(o.i8p), o.fp(sc), o.fp(of), o.fp(yh[i]))
        M = np.zeros((nb, bs, bs), F)
        o.lib.vgo_opq_accumulate_m(o.fp(v), o.fp(yh), n, dim, bs, o.fp(M))
        for b in range(nb):
            o.lib.vgo_opq_procrustes(o.fp(M[b]), bs, o.fp(rot[b]), None, None)
    assert np.array_equal(bits(opq.rotations), bits(rot))
    assert np.array_equal(opq.pq.codebooks, cb)
    assert np.array_equal(bits(opq.pq.scales), bits(sc)) and np.array_equal(bits(opq.pq.offsets), bits(of))
    # opq_test.go:56-99 rotation orthogonality (0.1), :26-54 encode/decode shapes, :101-131 asymmetric distance
    for b in range(nb):
        assert np.allclose(opq.rotations[b] @ opq.rotations[b].T, np.eye(bs), atol=0.1)
    codes = opq.EncodeBatch(v[:5])
    assert codes.shape == (5, m)
    rec = opq.DecodeBatch(codes)
    want = np.zeros(dim, F)
    tmp = np.zeros(dim, F)
    for i in range(5):
        o.lib.vgo_pq_decode(o.bp(codes[i]), dim, m, k, cb.ctypes.data_as(o.i8p), o.fp(sc), o.fp(of), o.fp(tmp))
        o.lib.vgo_opq_unrotate(o.fp(tmp), dim, bs, o.fp(rot), o.fp(want))
        assert np.array_equal(bits(rec[i]), bits(want))
    d_self = opq.ComputeAsymmetricDistance(v[0], codes[0])
    d_other = opq.ComputeAsymmetricDistance(v[0], codes[1])
    assert 0 < d_self < d_other


def test_kmeanspp_parallel_prefix_equals_sequential_chain():
    """PQ training with the exact parallel prefix (default) and with the one-lane sequential chain
    (VECGO_KMEANSPP_SEQUENTIAL=1, its own process): bit-identical codebooks on data full of ties and zero distances."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "kmeanspp_ab.py"), "120000", "64", "8"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("kind", ["gauss", "ties", "scaled"])
def test_pq_training_tensor_core_assignment_equals_exact_assignment(kind):
    """8-dim subspaces x 256 centroids: the tcgen05 assignment with its gap certificate + exact re-evaluation must give
    the float32 centroids and int8 codebooks of the exact CUDA-core assignment, bit for bit (tools/pq_assign_ab.py)."""
    import subprocess
    import sys

    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "pq_assign_ab.py")
    r = subprocess.run([sys.executable, tool, "100000", "64", "8", "5", kind], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


def test_search_with_page_locked_and_pageable_host_buffers(vg):
    """vg_index_search DMA's page-locked query / result buffers directly and stages pageable ones: same answers."""
    import torch

    rng = np.random.default_rng(31)
    n, dim, nq, k = 20000, 96, 700, 10   # 700 x 96 x 4 B > the 64 KB direct-DMA threshold
    x = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_F32, metric=0, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        rows, scores, counts = ix.search(q, k)
        qp = torch.from_numpy(q).pin_memory()
        out_t = (torch.empty((nq, k), dtype=torch.int32).pin_memory(), torch.empty((nq, k), dtype=torch.float32).pin_memory(),
                 torch.empty((nq,), dtype=torch.int32).pin_memory())
        out = (out_t[0].numpy().view(np.uint32), out_t[1].numpy(), out_t[2].numpy())
        r2, s2, c2 = ix.search(qp.numpy(), k, out=out)
        assert r2 is out[0] and s2 is out[1] and c2 is out[2]
        with pytest.raises(ValueError):
            ix.search(q, k, out=(out[0][:, :5], out[1], out[2]))
    assert np.array_equal(rows, r2) and np.array_equal(bits(scores), bits(s2)) and np.array_equal(counts, c2)
    want = _oracle_flat(q[:8], k, dim=dim, metric=0, vectors=x)
    _check_topk(rows[:8], scores[:8], counts[:8], want, k)
