"""GPU parity tests of the tcgen05 Flat path (vecgo_b200/csrc/vg_flat_tc.cu) through the C ABI.

The tensor-core GEMM only FILTERS; what vg_index_search returns must be bit-identical to
flat.(*Segment).Search on the SIMD path: row ids with ties by row id, float32 scores in
simd.SquaredL2 / simd.Dot summation order (checked against the CPU oracle), whatever the
filter did — including the cases where its certificate fails and the exact scan takes over.
"""
import ctypes as C

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


def tc_stats(vg):
    q, f = C.c_uint64(), C.c_uint64()
    vg._lib.call("vg_flat_tc_stats", C.byref(q), C.byref(f))
    return q.value, f.value


def oracle_topk(q, k, **kw):
    seg = o.FlatOracle(**kw)
    out, cnt = seg.search_batch(q, k, threads=8, mask=kw.get("mask"))
    return [out[i, : cnt[i]] for i in range(len(q))]


def check(rows, scores, counts, want):
    for i, w in enumerate(want):
        c = int(counts[i])
        assert c == len(w), (i, c, len(w))
        assert np.array_equal(rows[i, :c], w["row"]), i
        assert np.array_equal(bits(scores[i, :c]), bits(w["score"])), i
        assert np.all(rows[i, c:] == 0xFFFFFFFF)


@pytest.mark.parametrize("metric", [0, 2])
@pytest.mark.parametrize("n,dim,nq,k", [
    (8192, 128, 256, 10),     # resident query tile, exact multiples
    (9000, 100, 33, 10),      # dim % 32 != 0 (TMA zero fill), ragged query / row tiles
    (20000, 36, 300, 1),      # k = 1, two query tiles
    (8200, 768, 17, 32),      # streamed query tile (dim > 128), k = 32 -> 64 candidates
    (70000, 64, 1000, 5),     # many row splits
])
def test_tc_search_matches_oracle(vg, metric, n, dim, nq, k):
    rng = np.random.default_rng(n * 7 + dim)
    x = (rng.random((n, dim)) - (0.5 if metric else 0.0)).astype(F)
    q = (rng.random((nq, dim)) - (0.5 if metric else 0.0)).astype(F)
    if metric == 2:
        x, _ = vg.distance.NormalizeL2Batch(x)
    x[n // 2] = x[n // 3]          # duplicate rows -> equal scores -> tie broken by row id
    q[0] = x[n // 3]               # a query equal to a stored row (score 0 / max dot)
    vg._lib.call("vg_flat_tc_enable", 1)
    before = tc_stats(vg)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_F32, metric=metric, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        rows, scores, counts = ix.search(q, k)
    after = tc_stats(vg)
    assert after[0] - before[0] == nq, "the search did not go through the tensor-core filter"
    check(rows, scores, counts, oracle_topk(q, k, dim=dim, metric=metric, vectors=x))


def test_tc_filter_error_within_certificate_bound(vg):
    """vg_flat_tc_candidates: the threshold tau and the selected row groups against float64 group minima and the bound E."""
    n, dim, nq, kc = 50_000, 128, 64, 32
    rng = np.random.default_rng(5)
    x = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_F32, metric=0, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        gids = np.zeros((nq, kc), np.uint32)
        cnt = np.zeros(nq, np.int32)
        tau = np.zeros(nq, F)
        G = C.c_int64()
        L = vg._lib
        L.call("vg_flat_tc_candidates", ix.handle, L.ptr(q, L.f32p), nq, kc, L.ptr(gids, L.u32p), L.ptr(cnt, L.i32p), L.ptr(tau, L.f32p),
               C.byref(G))
    G = int(G.value)
    assert G == 32 and np.all(cnt == kc)
    x64, q64 = x.astype(np.float64), q.astype(np.float64)
    s_true = np.sum(x64 * x64, 1)[None, :] - 2 * q64 @ x64.T
    qn, xmax = np.sum(q64 * q64, 1), np.sum(x64 * x64, 1).max()
    E = 1.125 / 256 * np.sqrt(qn * xmax) + (qn + xmax) / 16384 + (xmax + 2 * np.sqrt(qn * xmax)) * G / 2 ** 23
    ng = -(-n // G)
    top = np.argsort(s_true, axis=1, kind="stable")[:, :10]
    for i in range(nq):
        gm = np.pad(s_true[i], (0, ng * G - n), constant_values=np.inf).reshape(ng, G).min(1)
        ent = gids[i].astype(np.int64)
        crowded = (ent & 0x80000000) != 0                      # both the best and the second best of the group are <= tau
        sel = np.where(crowded, ent & 0x7FFFFFFF, ent // G)    # otherwise the entry is the group's arg-min ROW
        assert len(set(sel.tolist())) == kc
        named = ent[~crowded]
        assert np.all(s_true[i, named] <= gm[named // G] + 2 * E[i])  # the named row is the group's minimum (to within E)
        # tau is the kc-th smallest APPROXIMATE group minimum: within E of the true one
        assert abs(np.sort(gm)[kc - 1] - tau[i]) <= E[i]
        # every group whose true minimum is clearly below tau is selected, none clearly above is
        assert set(np.where(gm <= tau[i] - E[i])[0].tolist()) <= set(sel.tolist())
        assert np.all(gm[sel] <= tau[i] + E[i])
        assert set((top[i] // G).tolist()) <= set(sel.tolist())


def test_tc_certificate_failure_falls_back_to_exact_scan(vg):
    """Rows that are all (nearly) equidistant defeat any approximate filter: the certificate must fail and the exact
    scan must produce the reference answer (ties by row id)."""
    n, dim, nq, k = 9000, 64, 40, 10
    rng = np.random.default_rng(9)
    base = rng.random(dim).astype(F)
    x = np.tile(base, (n, 1))
    x[:, 0] += (np.arange(n) % 7).astype(F) * F(1e-6)   # thousands of exact ties and near-ties
    q = rng.random((nq, dim)).astype(F)
    vg._lib.call("vg_flat_tc_enable", 1)
    before = tc_stats(vg)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_F32, metric=0, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        rows, scores, counts = ix.search(q, k)
    after = tc_stats(vg)
    assert after[1] - before[1] == nq, "every query should have needed the exact re-run"
    check(rows, scores, counts, oracle_topk(q, k, dim=dim, metric=0, vectors=x))


def test_tc_row_mask(vg):
    n, dim, nq, k = 9000, 96, 50, 10
    rng = np.random.default_rng(11)
    x = rng.random((n, dim)).astype(F)
    q = rng.random((nq, dim)).astype(F)
    keep = rng.random(n) < 0.4
    mask = np.packbits(keep, bitorder="little")
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_F32, metric=0, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        rows, scores, counts = ix.search(q, k, row_mask=mask)
    seg = o.FlatOracle(dim=dim, metric=0, vectors=x)
    out, cnt = seg.search_batch(q, k, mask=mask)
    check(rows, scores, counts, [out[i, : cnt[i]] for i in range(nq)])
    assert keep[rows[:, :k].reshape(-1)].all()


def test_tc_equals_exact_scan_config1(vg):
    """BASELINE configs[0] shape: the filter path and the exact CUDA-core scan return identical bits."""
    n, dim, nq, k = 100_000, 128, 1000, 10
    x = np.random.default_rng(42).random((n, dim), dtype=F)
    q = np.random.default_rng(43).random((nq, dim), dtype=F)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_F32, metric=0, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        vg._lib.call("vg_flat_tc_enable", 1)
        r1, s1, c1 = ix.search(q, k)
        vg._lib.call("vg_flat_tc_enable", 0)
        r0, s0, c0 = ix.search(q, k)
        vg._lib.call("vg_flat_tc_enable", 1)
    assert np.array_equal(r0, r1) and np.array_equal(bits(s0), bits(s1)) and np.array_equal(c0, c1)


# ----------------------------------------------------------------- k-means assignment through the tensor cores
@pytest.mark.parametrize("metric", [0, 2])
@pytest.mark.parametrize("n,dim,k", [(4096, 64, 256), (5000, 128, 300), (3000, 768, 130)])
def test_kmeans_assignment_tc_matches_oracle(vg, metric, n, dim, k):
    """kmeans.AssignPartition for a batch (kmeans.go:142-196: argmin L2 / argmax dot, strict, first wins) goes through
    the tcgen05 filter when there are >= 1024 samples and >= 128 centroids; assignments must equal the scalar loop."""
    rng = np.random.default_rng(n + k)
    v = rng.standard_normal((n, dim)).astype(F)
    cent = rng.standard_normal((k, dim)).astype(F)
    cent[k // 2] = cent[k // 3]            # duplicate centroids: ties must resolve to the smaller id
    v[0] = cent[k // 3]
    vg._lib.call("vg_flat_tc_enable", 1)
    before = tc_stats(vg)
    got = vg.kmeans.AssignPartition(v, cent, dim, metric)
    after = tc_stats(vg)
    assert after[0] - before[0] == n, "the assignment did not go through the tensor-core filter"
    want = np.array([o.lib.vgo_kmeans_assign(o.fp(v[i]), o.fp(cent), dim, k, metric) for i in range(n)], np.int32)
    assert np.array_equal(got, want)
    assert got[0] == k // 3


def test_kmeans_train_tc_matches_exact_path(vg):
    """TrainKMeans with the filter on and off must produce bit-identical centroids, assignments and iteration counts."""
    rng = np.random.default_rng(77)
    n, dim, k = 6000, 64, 160
    v = (rng.standard_normal((n, dim)) + rng.integers(0, 6, (n, 1)) * 2).astype(F)
    init = rng.permutation(n)[:k].astype(np.int64)
    out = []
    for on in (1, 0):
        vg._lib.call("vg_flat_tc_enable", on)
        out.append(vg.kmeans.TrainKMeans(v, dim, k, 0, 5, init_rows=init, seed=3, return_assign=True))
    vg._lib.call("vg_flat_tc_enable", 1)
    assert np.array_equal(bits(out[0][0]), bits(out[1][0])) and np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]


@pytest.mark.parametrize("metric,k", [(0, 10), (2, 32)])
def test_tc_long_segment_short_vectors_min_only_epilogue(vg, metric, k):
    """>= 2^20 rows with d <= 256: the pair kernel keeps only group minima and the exact stage scores whole groups."""
    n, dim, nq = 1_100_000, 64, 40
    rng = np.random.default_rng(77)
    x = (rng.random((n, dim), dtype=F) - (0.5 if metric else 0.0)).astype(F)
    q = (rng.random((nq, dim), dtype=F) - (0.5 if metric else 0.0)).astype(F)
    if metric == 2:
        x, _ = vg.distance.NormalizeL2Batch(x)
    x[n // 2] = x[n // 3]
    vg._lib.call("vg_flat_tc_enable", 1)
    with vg.index.DeviceIndex(codec=vg._lib.CODEC_F32, metric=metric, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        before = tc_stats(vg)
        rows, scores, counts = ix.search(q, k)
        after = tc_stats(vg)
        vg._lib.call("vg_flat_tc_enable", 0)
        try:
            r2, s2, c2 = ix.search(q, k)
        finally:
            vg._lib.call("vg_flat_tc_enable", 1)
    assert after[0] - before[0] == nq
    assert np.array_equal(rows, r2) and np.array_equal(bits(scores), bits(s2)) and np.array_equal(counts, c2)
    check(rows[:4], scores[:4], counts[:4], oracle_topk(q[:4], k, dim=dim, metric=metric, vectors=x))
