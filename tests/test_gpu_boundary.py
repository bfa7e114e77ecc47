"""GPU tests of the drop-in boundary itself: thread-safety of a shared handle, per-handle device / stream,
the launch-only search + deferred resolve pair, per-call QueryStats counters, SquaredL2Bounded, and the hardening of
vg_flat_open / vg_index_create against hostile inputs.  Everything crosses the C ABI and is checked against the oracle."""
import ctypes as C
import struct
import threading

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


def oracle_topk(q, k, **kw):
    want, cnt = o.FlatOracle(**kw).search_batch(q, k, threads=8)
    return want, cnt


def same(rows, scores, want):
    return np.array_equal(rows, want["row"]) and np.array_equal(bits(scores), bits(want["score"]))


# ------------------------------------------------------------------ threads
def test_32_threads_share_handles(vg):
    """engine.go:365,1324: BatchSearch runs up to 100 goroutines against immutable segments.  32 host threads hammer
    vg_index_search / vg_index_rerank / vg_index_score on SHARED handles (Flat through the tensor-core filter, SQ8
    through the decode-GEMM filter, RaBitQ, a small exact-scan PQ index) while another thread creates, fills and closes
    unrelated handles; every result is compared with the oracle."""
    L = vg._lib
    rng = np.random.default_rng(7)
    n, dim = 20000, 128
    x = rng.random((n, dim), dtype=F)
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(x)
    codes = sq.EncodeBatch(x)
    rq = vg.quantization.RaBitQuantizer(dim)
    rc = rq.EncodeBatch(x)
    flat = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n)
    flat.upload(vectors=x)
    s8 = vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales))
    s8.upload(codes=codes, vectors=x)
    rb = vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=n)
    rb.upload(codes=rc)
    nthreads, rounds = 32, 3
    # per-thread work and expectations (computed up front: the oracle is not the thing under test)
    jobs = []
    for t in range(nthreads):
        kind = ("flat", "sq8", "rabitq", "flat1")[t % 4]
        nq = 1 if kind == "flat1" else 16 + (t % 3) * 8
        q = rng.random((nq, dim), dtype=F)
        k = 10
        if kind in ("flat", "flat1"):
            want, _ = oracle_topk(q, k, dim=dim, metric=0, vectors=x)
        elif kind == "sq8":
            want, _ = oracle_topk(q, k, dim=dim, metric=0, quant=1, codes=codes, mins=sq.mins, inv=sq.invScales)
        else:
            want = None
        jobs.append((kind, q, k, want))
    rb_want = {}
    for t, (kind, q, k, _) in enumerate(jobs):
        if kind == "rabitq":
            out = np.zeros((len(q), k), o.cand_dtype)
            for i in range(len(q)):
                c = o.lib.vgo_rabitq_search(o.fp(q[i]), o.bp(rc), n, dim, k, None, out[i].ctypes.data_as(C.POINTER(o.Cand)), None)
                assert c == k
            rb_want[t] = out
    errors = []
    stop = threading.Event()

    def worker(t):
        try:
            kind, q, k, want = jobs[t]
            for _ in range(rounds):
                if kind in ("flat", "flat1"):
                    rows, scores, _ = flat.search(q, k)
                    assert same(rows, scores, want), f"thread {t}: flat result differs"
                    rr = flat.rerank(q, rows)
                    assert np.array_equal(bits(rr), bits(scores)), f"thread {t}: rerank differs"
                elif kind == "sq8":
                    rows, scores, _ = s8.search(q, k)
                    assert same(rows, scores, want), f"thread {t}: sq8 result differs"
                    sc = s8.score(q, rows)
                    assert np.array_equal(bits(sc), bits(scores)), f"thread {t}: sq8 gather score differs"
                else:
                    rows, scores, _ = rb.search(q, k)
                    assert same(rows, scores, rb_want[t]), f"thread {t}: rabitq result differs"
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def churn():
        # a different handle is created, filled, searched and closed while the others are being searched
        r2 = np.random.default_rng(99)
        try:
            while not stop.is_set():
                y = r2.random((3000, 64), dtype=F)
                with vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=64, rows=3000) as ix:
                    ix.upload(vectors=y)
                    qq = r2.random((4, 64), dtype=F)
                    rows, scores, _ = ix.search(qq, 5)
                    want, _ = oracle_topk(qq, 5, dim=64, metric=0, vectors=y)
                    assert same(rows, scores, want), "churn handle result differs"
        except Exception as e:  # noqa: BLE001
            errors.append("churn: " + repr(e))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
    ch = threading.Thread(target=churn)
    ch.start()
    for t_ in th:
        t_.start()
    for t_ in th:
        t_.join()
    stop.set()
    ch.join()
    for ix in (flat, s8, rb):
        ix.close()
    assert not errors, errors[:3]


def test_handle_device_and_explicit_device_create(vg):
    L = vg._lib
    with vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=16, rows=64, device=0) as ix:
        assert ix.device() == 0
    n = C.c_int32()
    L.call("vg_device_count", C.byref(n))
    with pytest.raises(vg.VecgoError):
        vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=16, rows=64, device=n.value)  # ordinal out of range


# ------------------------------------------------------------------ async + resolve
@pytest.mark.parametrize("codec", ["flat", "sq8"])
def test_async_search_and_deferred_resolve(vg, codec):
    """vg_index_search_dev_async launches only; the flags say which queries still need vg_index_search_resolve.  Benign
    data: no flag set, results already final.  Tie-heavy data: flags set, resolve makes the result the oracle's."""
    import torch

    L = vg._lib
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    n, dim, nq, k = 16384, 128, 48, 10
    for ties in (False, True):
        x = rng.random((n, dim), dtype=F)
        if ties:
            x[1000:9000] = x[1000]  # 8000 identical rows: no certificate can separate the k-th best from the rest
        q = rng.random((nq, dim), dtype=F)
        if ties:
            q[:8] = x[1000] + F(0.001) * rng.standard_normal((8, dim)).astype(F)
        if codec == "flat":
            ix = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n)
            ix.upload(vectors=x)
            want, _ = oracle_topk(q, k, dim=dim, metric=0, vectors=x)
        else:
            sq = vg.quantization.ScalarQuantizer(dim)
            sq.Train(x)
            codes = sq.EncodeBatch(x)
            ix = vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales))
            ix.upload(codes=codes)
            want, _ = oracle_topk(q, k, dim=dim, metric=0, quant=1, codes=codes, mins=sq.mins, inv=sq.invScales)
        stream = torch.cuda.Stream()
        ix.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            dq = torch.from_numpy(q).to(dev)
            rows = torch.empty((nq, k), dtype=torch.int32, device=dev)
            scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
            counts = torch.empty((nq,), dtype=torch.int32, device=dev)
            flags = torch.full((nq,), -1, dtype=torch.int32, device=dev)
            stream.synchronize()
            ix.search_dev_async(dq.data_ptr(), nq, k, rows.data_ptr(), scores.data_ptr(), counts.data_ptr(), flags.data_ptr())
            stream.synchronize()
            f = flags.cpu().numpy()
            assert set(np.unique(f)) <= {0, 1}
            got_r = rows.cpu().numpy().view(np.uint32)
            got_s = scores.cpu().numpy()
            ok = f == 0
            # every query WITHOUT a flag is already final
            assert np.array_equal(got_r[ok], want["row"][ok]) and np.array_equal(bits(got_s[ok]), bits(want["score"][ok]))
            if not ties:
                assert f.sum() == 0
            else:
                assert f[:8].sum() > 0
            nres = ix.search_resolve(dq.data_ptr(), nq, k, rows.data_ptr(), scores.data_ptr(), counts.data_ptr(), flags.data_ptr())
            stream.synchronize()
            assert nres == int(f.sum())
            st = L.last_search_stats()
            if ties:
                assert st["second_chance_queries"] == nres and st["exact_rerun_queries"] <= nres
            assert same(rows.cpu().numpy().view(np.uint32), scores.cpu().numpy(), want)
        ix.close()


def test_search_stats_counters(vg):
    L = vg._lib
    rng = np.random.default_rng(12)
    n, dim, nq, k = 9000, 64, 20, 5
    x = rng.random((n, dim), dtype=F)
    with vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        ix.search(rng.random((nq, dim), dtype=F), k)
        st = L.last_search_stats()
        assert st["queries"] == nq and st["filter_queries"] == nq and st["distance_computations"] == nq * n
        assert st["exact_rerun_queries"] == 0
        ix.search(rng.random((3, dim), dtype=F), k)  # below the filter's batch threshold: exact scan
        st = L.last_search_stats()
        assert st["queries"] == 3 and st["filter_queries"] == 0 and st["distance_computations"] == 3 * n


# ------------------------------------------------------------------ bounded L2
def test_squared_l2_bounded_matches_oracle(vg):
    """simd.SquaredL2Bounded (bounded_l2_avx512.c:19-107): distance or the partial sum of the block that crossed the bound."""
    L = vg._lib
    rng = np.random.default_rng(13)
    for dim in (8, 63, 64, 65, 100, 128, 200, 768, 777):
        n, nq, r = 300, 4, 9
        x = rng.standard_normal((n, dim)).astype(F)
        q = rng.standard_normal((nq, dim)).astype(F)
        cand = rng.integers(0, n, (nq, r)).astype(np.uint32)
        cand[1, 2] = n + 3
        full = ((q[:, None, :].astype(np.float64) - x[np.minimum(cand, n - 1)]) ** 2).sum(-1)
        for per_pair in (False, True):
            bounds = (full * rng.uniform(0.2, 1.3, full.shape)).astype(F) if per_pair else (full.mean(1) * 0.8).astype(F)
            with vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n) as ix:
                ix.upload(vectors=x)
                got, ex = ix.l2_bounded(q, cand, bounds)
            assert np.isnan(got[1, 2]) and not ex[1, 2]
            for i in range(nq):
                for j in range(r):
                    if cand[i, j] >= n:
                        continue
                    b = bounds[i, j] if per_pair else bounds[i]
                    e = C.c_int32()
                    w = o.lib.vgo_squared_l2_bounded_a512(o.fp(q[i]), o.fp(x[cand[i, j]]), dim, F(b), C.byref(e))
                    assert bits(got[i, j]) == bits(F(w)) and bool(ex[i, j]) == bool(e.value), (dim, i, j)
    # the pairwise mirror of the kernel table
    a, b = rng.standard_normal((6, 130)).astype(F), rng.standard_normal((6, 130)).astype(F)
    bd = np.array([0.0, 50.0, 100.0, 260.0, np.inf, np.nan], F)
    out, ex = np.zeros(6, F), np.zeros(6, np.uint8)
    L.call("vg_simd_squared_l2_bounded", L.ptr(a, L.f32p), L.ptr(b, L.f32p), 6, 130, L.ptr(bd, L.f32p), L.ptr(out, L.f32p), L.ptr(ex, L.u8p))
    for i in range(6):
        e = C.c_int32()
        w = o.lib.vgo_squared_l2_bounded_a512(o.fp(a[i]), o.fp(b[i]), 130, bd[i], C.byref(e))
        assert bits(out[i]) == bits(F(w)) and int(ex[i]) == e.value


# ------------------------------------------------------------------ zero scores under the dot metric
def test_zero_dot_scores_are_positive_zero(vg):
    """Orthogonal / zero vectors under the dot metric: the reference's accumulators start at +0.0 and return +0.0; the
    descending key path must not hand back -0.0."""
    L = vg._lib
    n, dim = 64, 32
    x = np.zeros((n, dim), F)
    x[np.arange(n), np.arange(n) % 16] = 1.0          # rows live in dims 0..15
    q = np.zeros((2, dim), F)
    q[0, 20] = 1.0                                    # orthogonal to every row
    q[1, 3] = 2.0
    with vg.index.DeviceIndex(codec=L.CODEC_F32, metric=2, dim=dim, rows=n) as ix:
        ix.upload(vectors=x)
        rows, scores, counts = ix.search(q, 8)
    want, _ = oracle_topk(q, 8, dim=dim, metric=2, vectors=x)
    assert same(rows, scores, want)
    assert np.all(bits(scores[0]) == 0)               # +0.0, not 0x80000000


# ------------------------------------------------------------------ hostile inputs
def test_index_create_rejects_row_range_overflow(vg):
    L = vg._lib
    with pytest.raises(vg.VecgoError) as e:
        vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=8, rows=100, row_base=0xFFFFFFFE - 50)
    assert e.value.status == L.ERR_INVALID
    with vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=8, rows=100, row_base=0xFFFFFFFE - 100):
        pass


def test_flat_open_rejects_wrapping_offsets(vg):
    """A crafted header whose section offsets sit near 2^64 must fail with VG_ERR_FORMAT, not read out of bounds."""
    rng = np.random.default_rng(14)
    n, dim = 300, 16
    v = rng.standard_normal((n, dim)).astype(F)
    for quant in (vg.flat.QuantizationNone, vg.flat.QuantizationSQ8):
        data = bytearray(vg.flat.write_segment(segment_id=1, vectors=v, metric=0, quantization=quant))
        for field_off in (40, 48, 56, 64, 72, 80, 88):
            for evil in (0xFFFFFFFFFFFFFFF0, 0xFFFFFFFFFFFFFF00, 1 << 63, len(data) + 1):
                bad = bytearray(data)
                struct.pack_into("<Q", bad, field_off, evil)
                uses = {40: False, 48: False, 56: quant != 0, 64: quant != 0, 72: True, 80: True, 88: True}[field_off]
                if not uses:
                    continue
                with pytest.raises(vg.VecgoError) as e:
                    vg.flat.Segment.Open(bytes(bad), verify_checksum=False)
                assert e.value.status == vg._lib.ERR_FORMAT, (quant, field_off, hex(evil))
        # absurd row counts / dimensions whose products overflow
        bad = bytearray(data)
        struct.pack_into("<II", bad, 16, 0xFFFFFFFF, 0xFFFFFFFF)
        with pytest.raises(vg.VecgoError):
            vg.flat.Segment.Open(bytes(bad), verify_checksum=False)


# ------------------------------------------------------------------ PQ training split by subspace
@pytest.mark.parametrize("n,dim,m,parts", [(6000, 768, 96, 8), (5000, 64, 8, 3), (3000, 48, 6, 2)])
def test_pq_train_subspace_ranges_equal_whole(vg, n, dim, m, parts):
    """pq.go:79-140 trains every subspace in its own goroutine: subspaces [lo, hi) trained on their own (what one GPU of W
    does, vg_pq_train_range_dev) must give the same codebooks, scales, offsets and float32 centroids, bit for bit, as the
    slice of the whole training — including the seeded k-means++ draws and empty-cluster re-seeding (RNG stream = the
    subspace's number).  96 x 8-dim subspaces go through the tensor-core assignment on a column slice."""
    import torch

    L = vg._lib
    rng = np.random.default_rng(5)
    v = rng.standard_normal((n, dim)).astype(F)
    v[: n // 10] = v[0]  # duplicates: empty clusters appear and get re-seeded
    ds = dim // m
    dx = torch.from_numpy(v).cuda()
    cb, sc, of = np.zeros(m * 256 * ds, np.int8), np.zeros(m, F), np.zeros(m, F)
    cent = np.zeros((m, 256, ds), F)
    L.call("vg_pq_train_dev", dx.data_ptr(), n, dim, m, 256, 4, 9, L.ptr(cb, L.i8p), L.ptr(sc, L.f32p), L.ptr(of, L.f32p), L.ptr(cent, L.f32p))
    from vecgo_b200.sharded import shard_range

    for r in range(parts):
        lo, hi = shard_range(m, r, parts)
        g = hi - lo
        cb2, sc2, of2 = np.zeros(g * 256 * ds, np.int8), np.zeros(g, F), np.zeros(g, F)
        cent2 = np.zeros((g, 256, ds), F)
        L.call("vg_pq_train_range_dev", dx.data_ptr(), n, dim, m, 256, 4, 9, lo, hi, L.ptr(cb2, L.i8p), L.ptr(sc2, L.f32p), L.ptr(of2, L.f32p),
               L.ptr(cent2, L.f32p))
        assert np.array_equal(cb2, cb[lo * 256 * ds:hi * 256 * ds]), (r, "codebooks")
        assert np.array_equal(bits(sc2), bits(sc[lo:hi])) and np.array_equal(bits(of2), bits(of[lo:hi])), (r, "scales/offsets")
        assert np.array_equal(bits(cent2), bits(cent[lo:hi])), (r, "centroids")


# ------------------------------------------------------------------ shard groups: NCCL inside the C ABI
def _ngpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.skipif("_ngpus() < 2")
def test_shard_group_single_process_two_gpus(vg):
    """vg_shard_group_* driving two GPUs from ONE process (what a Go host does): SQ8 shards on cuda:0 / cuda:1, the
    all-gather and the merge inside the library; result = the oracle's search over the whole database.  Also the rerank
    form (global approximate top-r, exact rerank, top-k) and a float32 Flat group."""
    from vecgo_b200.sharded import ShardGroup, shard_range

    L = vg._lib
    rng = np.random.default_rng(21)
    n, dim, nq, k, r = 30000, 128, 40, 10, 60
    x = rng.standard_normal((n, dim)).astype(F)
    x[n - 5] = x[3]   # a cross-shard exact tie: resolved by the global row id
    q = rng.standard_normal((nq, dim)).astype(F)
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(x)
    codes = sq.EncodeBatch(x)
    W = 2
    grp = ShardGroup.single_process(list(range(W)))
    shards, flats = [], []
    for w in range(W):
        lo, hi = shard_range(n, w, W)
        ix = vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=hi - lo, row_base=lo, sq8=(sq.mins, sq.invScales), device=w)
        ix.upload(codes=codes[lo:hi], vectors=x[lo:hi])
        shards.append(ix)
        fx = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=hi - lo, row_base=lo, device=w)
        fx.upload(vectors=x[lo:hi])
        flats.append(fx)
    rows, scores, counts = grp.search(shards, q, k)
    want, _ = oracle_topk(q, k, dim=dim, metric=0, quant=1, codes=codes, mins=sq.mins, inv=sq.invScales)
    assert same(rows, scores, want)
    rows, scores, counts = grp.search(flats, q, k)
    want, _ = oracle_topk(q, k, dim=dim, metric=0, vectors=x)
    assert same(rows, scores, want)
    # rerank form: approximate top-r over ALL rows, exact float32 rerank, top-k by (score, row)
    rows, scores, counts = grp.search(shards, q, k, r=r)
    whole = o.FlatOracle(dim=dim, metric=0, quant=1, codes=codes, mins=sq.mins, inv=sq.invScales, vectors=x)
    approx, acnt = whole.search_batch(q, r)
    for i in range(nq):
        cand = approx["row"][i, : acnt[i]]
        ex = whole.rerank(q[i], cand)
        order = np.lexsort((cand, ex))[:k]
        assert np.array_equal(rows[i], cand[order]) and np.array_equal(bits(scores[i]), bits(ex[order])), i
    for ix in shards + flats:
        ix.close()
    grp.close()
    # a handle on the wrong GPU is refused, not searched
    grp = ShardGroup.single_process([0, 1])
    a = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=8, rows=64, device=0)
    b = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=8, rows=64, device=0)
    with pytest.raises(vg.VecgoError):
        grp.search([a, b], np.zeros((1, 8), F), 1)
    a.close()
    b.close()
    grp.close()


def test_rerank_from_host_resident_vectors(vg):
    """vg_index_set_host_vectors: the float32 rows stay in host memory (pageable numpy array, page-locked by the library;
    and a region the caller page-locked itself); rerank / search_rerank give the same bits as the device-resident path."""
    import torch

    L = vg._lib
    rng = np.random.default_rng(31)
    n, dim, nq, r, k = 20000, 96, 24, 50, 10
    x = rng.standard_normal((n, dim)).astype(F)
    q = rng.standard_normal((nq, dim)).astype(F)
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(x)
    codes = sq.EncodeBatch(x)
    with vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales)) as dev_ix:
        dev_ix.upload(codes=codes, vectors=x)
        rows, _, _ = dev_ix.search(q, r)
        want_scores = dev_ix.rerank(q, rows)
        want = dev_ix.search_rerank(q, r, k)
    pinned = torch.from_numpy(x.copy()).pin_memory()
    for host in (x.copy(), pinned.numpy()):
        with vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales)) as ix:
            ix.upload(codes=codes)
            with pytest.raises(vg.VecgoError):
                ix.rerank(q, rows)          # no float32 rows yet
            ix.set_host_vectors(host)
            got_scores = ix.rerank(q, rows)
            got = ix.search_rerank(q, r, k)
        assert np.array_equal(bits(got_scores), bits(want_scores))
        assert np.array_equal(got[0], want[0]) and np.array_equal(bits(got[1]), bits(want[1])) and np.array_equal(got[2], want[2])
    for i in range(nq):   # and against the oracle's Segment.Rerank
        ex = np.array([o.lib.vgo_sql2_a512(o.fp(q[i]), o.fp(x[rw]), dim) for rw in rows[i]], F)
        assert np.array_equal(bits(want_scores[i]), bits(ex))
    with vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n) as fx:
        with pytest.raises(vg.VecgoError):
            fx.set_host_vectors(x)          # a float32 index scans its rows: device-resident only
