"""Known-answer tests ported from the reference's own unit tests: they pin the
oracle (oracle/vecgo_oracle.c) and, where the ISA matters, the reference's C
kernels compiled into oracle/_ref.  Citations relative to /root/reference."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as o

F = np.float32


def f32(*v):
    return np.array(v, dtype=np.float32)


# internal/simd/floats_test.go:11-33
DOT_CASES = [
    ([1, 2, 3], [4, 5, 6], 32.0),
    ([-1, -2, -3], [-4, -5, -6], 32.0),
    ([1, 2, 3, 1, 2, 3], [4, 5, 6, 4, 5, 6], 64.0),
    ([1, -2, 3], [-4, 5, -6], -32.0),
    ([0, 0, 0], [0, 0, 0], 0.0),
    (list(range(1, 10)), list(range(1, 10)), 285.0),
    (list(range(1, 11)), list(range(1, 11)), 385.0),
    (list(range(1, 16)), list(range(1, 16)), 1240.0),
    (list(range(1, 17)), list(range(1, 17)), 1496.0),
]
# internal/simd/floats_test.go:55-74
L2_CASES = [
    ([1, 2, 3], [4, 5, 6], 27.0),
    ([-1, -2, -3], [-4, -5, -6], 27.0),
    ([1, 2, 3, 1, 2, 3], [4, 5, 6, 4, 5, 6], 54.0),
    ([1, -2, 3], [-4, 5, -6], 155.0),
    ([0, 0, 0], [0, 0, 0], 0.0),
]


@pytest.mark.parametrize("a,b,want", DOT_CASES)
def test_dot_kat(a, b, want):
    a, b = f32(*a), f32(*b)
    assert o.lib.vgo_dot_a512(o.fp(a), o.fp(b), len(a)) == want
    assert o.lib.vgo_dot_generic(o.fp(a), o.fp(b), len(a)) == want
    if o.ref is not None:
        assert o.ref_dot(a, b) == want


@pytest.mark.parametrize("a,b,want", L2_CASES)
def test_sql2_kat(a, b, want):
    a, b = f32(*a), f32(*b)
    assert o.lib.vgo_sql2_a512(o.fp(a), o.fp(b), len(a)) == want
    assert o.lib.vgo_sql2_generic(o.fp(a), o.fp(b), len(a)) == want
    if o.ref is not None:
        assert o.ref_sql2(a, b) == want


def test_dot_nil():  # floats_test.go:35-39
    z = np.zeros(1, F)
    assert o.lib.vgo_dot_a512(o.fp(z), o.fp(z), 0) == 0.0


def test_pq_adc_kat():  # floats_test.go:330-412
    table = (np.arange(512) % 256).astype(F)
    for codes, want in ([0, 0], 0.0), ([255, 255], 510.0), ([10, 20], 30.0):
        c = np.array(codes, np.uint8)
        assert o.lib.vgo_pq_adc_a512(o.fp(table), o.bp(c), 2) == want
        assert o.lib.vgo_pq_adc_generic(o.fp(table), o.bp(c), 2) == want
    for m, codes in (8, list(range(8))), (16, [17 * i for i in range(16)]):
        table = np.array([[i * 1000 + j for j in range(256)] for i in range(m)], F).ravel()
        c = np.array(codes, np.uint8)
        want = sum(table[i * 256 + codes[i]] for i in range(m))
        assert o.lib.vgo_pq_adc_a512(o.fp(table), o.bp(c), m) == want
        if o.ref is not None:
            assert o.ref_pq_adc(table, c, m) == want


def test_scale_kat():  # floats_test.go:441-464
    for inp, s, want in ([1, 2, 3], 2.0, [2, 4, 6]), ([1, 2, 3], 0.0, [0, 0, 0]), ([1, -2, 3], -1.0, [-1, 2, -3]):
        a = f32(*inp)
        o.lib.vgo_scale(o.fp(a), len(a), s)
        assert a.tolist() == [float(x) for x in want]


HAMMING_CASES = [  # floats_test.go:504-532
    ([], [], 0),
    ([0xFF, 0xAA], [0xFF, 0xAA], 0),
    ([0x00, 0xFF], [0xFF, 0x00], 16),
    ([0x0F], [0xF0], 8),
    ([1, 0, 0, 0, 0, 0, 0, 0], [0] * 8, 1),
    (list(range(17)), [0xFF, 1, 0xFD, 3, 0xFB, 5, 0xF9, 7, 0xF7, 9, 0xF5, 0xB, 0xF3, 0xD, 0xF1, 0xF, 0xEF], 72),
]


@pytest.mark.parametrize("a,b,want", HAMMING_CASES)
def test_hamming_kat(a, b, want):
    a = np.array(a + [0], np.uint8)
    b = np.array(b + [0], np.uint8)
    n = len(a) - 1
    assert o.lib.vgo_hamming(o.bp(a), o.bp(b), n) == want
    if o.ref is not None:
        assert o.ref.hammingAvx512(o.bp(a), o.bp(b), n) == want


def test_hamming_boundaries():  # floats_test.go:534-547 (deterministic inputs, portable verbatim)
    for n in [0, 1, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65]:
        a = np.array([((i * 131) & 0xFF) ^ ((i >> 1) & 0xFF) for i in range(n)] + [0], np.uint8)
        b = a ^ np.uint8(0x5A)
        want = sum(bin(int(x) ^ int(y)).count("1") for x, y in zip(a[:n], b[:n]))
        assert o.lib.vgo_hamming(o.bp(a), o.bp(b), n) == want
        if o.ref is not None:
            assert o.ref.hammingAvx512(o.bp(a), o.bp(b), n) == want


# ---------------------------------------------------------------- binary.go
def test_bq_basic():  # quantization/binary_test.go:9-39
    v = np.array([1.0 if i % 2 == 0 else -1.0 for i in range(128)], F)
    out = np.zeros(16, np.uint8)
    o.lib.vgo_bq_encode(o.fp(v), 128, 0.0, o.bp(out))
    w = out.view("<u8")
    assert w[0] == 0x5555555555555555 and w[1] == 0x5555555555555555


def test_bq_train():  # binary_test.go:41-62
    v = f32(1, 2, 3, 4, 5, 6, 7, 8)
    assert o.lib.vgo_bq_train(o.fp(v), 2, 4) == 4.5


def test_bq_threshold():  # binary_test.go:64-84
    v = f32(0.0, 0.4, 0.5, 0.6, 1.0, -1.0, 0.5, 0.49)
    out = np.zeros(8, np.uint8)
    o.lib.vgo_bq_encode(o.fp(v), 8, 0.5, o.bp(out))
    assert out.view("<u8")[0] == 0b01011100


def test_hamming_distance_words():  # binary_test.go:86-104
    cases = [([0], [0], 0), ([1], [0], 1), ([0xFF], [0], 8), ([0xFFFFFFFFFFFFFFFF], [0], 64),
             ([0x5555555555555555], [0xAAAAAAAAAAAAAAAA], 64), ([0, 0], [0xFFFFFFFFFFFFFFFF] * 2, 128)]
    for a, b, want in cases:
        a = np.array(a, "<u8").view(np.uint8)
        b = np.array(b, "<u8").view(np.uint8)
        assert o.lib.vgo_hamming(o.bp(a), o.bp(b), len(a)) == want


# -------------------------------------------------------------- quantizer.go
def test_sq8_train_kat():  # quantization/quantizer_test.go:8-37
    v = f32(-1.0, 0.0, 1.0, -0.5, 0.5, 2.0, -2.0, 1.0, 3.0)
    mins, maxs, sc, inv = (np.zeros(3, F) for _ in range(4))
    assert o.lib.vgo_sq8_train(o.fp(v), 3, 3, o.fp(mins), o.fp(maxs), o.fp(sc), o.fp(inv)) == 0
    assert mins[0] == -2.0 and maxs[0] == -0.5 and mins[2] == 1.0 and maxs[2] == 3.0


def test_sq8_l2_batch_kat():  # quantizer_test.go:230-271
    dim = 4
    mins = np.zeros(dim, F)
    maxs = np.full(dim, 10, F)
    sc = np.full(dim, F(255.0) / F(10.0), F)
    inv = np.full(dim, F(10.0) / F(255.0), F)
    q = f32(1, 2, 3, 4)
    codes = np.zeros((2, dim), np.uint8)
    o.lib.vgo_sq8_encode(o.fp(f32(1, 2, 3, 4)), dim, o.fp(mins), o.fp(maxs), o.fp(sc), o.bp(codes[0]))
    o.lib.vgo_sq8_encode(o.fp(f32(2, 3, 4, 5)), dim, o.fp(mins), o.fp(maxs), o.fp(sc), o.bp(codes[1]))
    out = np.zeros(2, F)
    o.lib.vgo_sq8u_l2_batch_a512(o.fp(q), o.bp(codes), o.fp(mins), o.fp(inv), dim, 2, o.fp(out))
    assert out[0] <= 0.1 and abs(out[1] - 4.0) <= 0.2


def test_sq8_encode_rounding():
    """uint8(x+0.5) truncation, clamp to [min,max] (quantizer.go:200-222)."""
    mins, maxs = f32(0.0), f32(255.0)
    sc, inv = np.zeros(1, F), np.zeros(1, F)
    o.lib.vgo_sq8_set_bounds(o.fp(mins), o.fp(maxs), 1, o.fp(sc), o.fp(inv))
    assert sc[0] == 1.0 and inv[0] == 1.0
    for v, want in (-5.0, 0), (0.49, 0), (0.5, 1), (1.5, 2), (254.5, 255), (300.0, 255):
        out = np.zeros(1, np.uint8)
        o.lib.vgo_sq8_encode(o.fp(f32(v)), 1, o.fp(mins), o.fp(maxs), o.fp(sc), o.bp(out))
        assert out[0] == want, v
    # SetBounds: diff < 1e-9 → scale = invScale = 0 (quantizer.go:62-69)
    o.lib.vgo_sq8_set_bounds(o.fp(f32(1.0)), o.fp(f32(1.0)), 1, o.fp(sc), o.fp(inv))
    assert sc[0] == 0 and inv[0] == 0


# ------------------------------------------------------------------- int4.go
def test_int4_pack_order_and_round():
    """high nibble = even dim, math.Round half away (int4.go:77-103)."""
    minv, diff = f32(0, 0, 0), f32(15, 15, 30)
    out = np.zeros(2, np.uint8)
    o.lib.vgo_int4_encode(o.fp(f32(3.0, 12.0, 5.0)), 3, o.fp(minv), o.fp(diff), o.bp(out))
    assert out[0] == (3 << 4) | 12
    assert out[1] == (3 << 4)  # 5/30*15 = 2.5 → 3 (half away from zero)
    dec = np.zeros(3, F)
    o.lib.vgo_int4_decode(o.bp(out), 3, o.fp(minv), o.fp(diff), o.fp(dec))
    assert dec.tolist() == [3.0, 12.0, 6.0]


def test_int4_train_zero_diff():  # int4.go:54-60
    v = f32(1, 5, 1, 7)
    minv, diff = np.zeros(2, F), np.zeros(2, F)
    o.lib.vgo_int4_train(o.fp(v), 2, 2, o.fp(minv), o.fp(diff))
    assert minv.tolist() == [1.0, 5.0] and diff.tolist() == [1.0, 2.0]


# ------------------------------------------------------------------ rabitq.go
def test_rabitq_encode_layout_and_distance():
    dim = 128
    rng = np.random.default_rng(7)
    v = rng.standard_normal(dim).astype(F)
    code = np.zeros(16 + 4, np.uint8)
    o.lib.vgo_rabitq_encode(o.fp(v), dim, o.bp(code))
    bits = np.unpackbits(code[:16], bitorder="little")[:dim]
    assert np.array_equal(bits.astype(bool), v >= 0)
    norm = code[16:].view("<f4")[0]
    assert abs(norm - np.linalg.norm(v.astype(np.float64))) < 1e-5 * norm  # rabitq_test.go norm round-trip 1e-5
    # distance to itself: hamming 0, norms equal → exactly 0; always ≥ 0 (rabitq_test.go:60-97)
    assert o.lib.vgo_rabitq_distance(o.fp(v), dim, o.bp(code)) == 0.0
    w = rng.standard_normal(dim).astype(F)
    assert o.lib.vgo_rabitq_distance(o.fp(w), dim, o.bp(code)) >= 0.0


# -------------------------------------------------------------- candidate heap
def test_heap_order_and_ties():  # internal/searcher/candidate_queue.go:12-38
    rng = np.random.default_rng(3)
    n = 500
    cands = np.zeros(n, o.cand_dtype)
    cands["seg"] = rng.integers(0, 3, n)
    cands["row"] = rng.permutation(n)
    cands["score"] = rng.integers(0, 20, n).astype(F)  # many ties
    for desc in (0, 1):
        for k in (1, 7, 100, 600):
            out = np.zeros(max(k, 1), o.cand_dtype)
            cnt = o.lib.vgo_heap_topk(cands.ctypes.data_as(C.POINTER(o.Cand)), n, k, desc,
                                      out.ctypes.data_as(C.POINTER(o.Cand)))
            key = np.lexsort((cands["row"], cands["seg"], -cands["score"] if desc else cands["score"]))
            want = cands[key][: min(k, n)]
            assert cnt == min(k, n)
            assert np.array_equal(out[:cnt]["row"], want["row"]) and np.array_equal(out[:cnt]["seg"], want["seg"])


# --------------------------------------------------------------------- kmeans
def test_kmeans_kat():  # internal/kmeans/kmeans_test.go:12-88 (two well separated blobs)
    pts = f32(0, 0, 0.1, 0.1, 0.2, 0.0, 10, 10, 10.1, 10.1, 9.9, 10.0).reshape(6, 2)
    cent = np.zeros((2, 2), F)
    assign = np.zeros(6, np.int32)
    init = np.array([0, 3], np.int64)
    it = o.lib.vgo_kmeans_train(o.fp(pts), 6, 2, 2, 0, 10, init.ctypes.data_as(o.i64p), 1, o.fp(cent),
                                assign.ctypes.data_as(o.i32p))
    assert it >= 1
    assert assign.tolist() == [0, 0, 0, 1, 1, 1]
    assert np.allclose(cent[0], [0.1, 1 / 30], atol=1e-6) and np.allclose(cent[1], [10.0, 10.0333333], atol=1e-5)
    assert o.lib.vgo_kmeans_assign(o.fp(f32(9, 9)), o.fp(cent), 2, 2, 0) == 1
    out = np.zeros(2, np.int64)
    assert o.lib.vgo_find_closest_centroids(o.fp(f32(1, 1)), o.fp(cent), 2, 2, 1, 0, out.ctypes.data_as(o.i64p)) == 1
    assert out[0] == 0


def test_crc32c_kat():
    d = np.frombuffer(b"123456789", np.uint8)
    assert o.lib.vgo_crc32c(o.bp(d), 9) == 0xE3069283  # CRC-32C check value


# ------------------------------------------------------------------ OPQ (opq.go, svd.go, opq_test.go)
def test_opq_block_size_rule():
    # NewOptimizedProductQuantizer opq.go:41-58: full dimension up to 64, else the multiple of the subvector size
    # that divides dim and is closest to 32 (first wins on ties)
    assert o.lib.vgo_opq_block_size(32, 8) == 32      # opq_test.go:11-24
    assert o.lib.vgo_opq_block_size(64, 8) == 64
    assert o.lib.vgo_opq_block_size(768, 96) == 32
    assert o.lib.vgo_opq_block_size(128, 8) == 32
    assert o.lib.vgo_opq_block_size(96, 2) == 48      # candidates 48, 96
    assert o.lib.vgo_opq_block_size(100, 1) == 100


def test_opq_svd_and_procrustes_properties():
    """svd.go: M = U diag(sigma) V^T; computeProcrustesRotation: R orthogonal, det +1, maximises tr(R^T M)."""
    rng = np.random.default_rng(3)
    for n in (4, 16, 32):
        M = rng.standard_normal((n, n)).astype(F)
        u = M.copy()
        R = np.zeros((n, n), F)
        sig = np.zeros(n, F)
        v = np.zeros((n, n), F)
        o.lib.vgo_opq_procrustes(o.fp(u), n, o.fp(R), o.fp(sig), o.fp(v))
        ref_sig = np.linalg.svd(M.astype(np.float64), compute_uv=False)
        assert np.allclose(np.sort(sig)[::-1], ref_sig, rtol=1e-3, atol=1e-4)
        assert np.allclose(R @ R.T, np.eye(n), atol=1e-4)
        assert np.linalg.det(R.astype(np.float64)) > 0.99
        # optimal rotation (Kabsch): trace(R^T M) equals sum(sigma) with the smallest one negated when det(U V^T) < 0
        U64, s64, Vt64 = np.linalg.svd(M.astype(np.float64))
        d = np.sign(np.linalg.det(U64 @ Vt64))
        best = s64[:-1].sum() + d * s64[-1]
        assert abs(np.trace(R.astype(np.float64).T @ M.astype(np.float64)) - best) < 1e-3 * best
    # identity in, identity out (the first Train round rotates by I, opq.go:98-101)
    eye = np.eye(8, dtype=F)
    R = np.zeros((8, 8), F)
    o.lib.vgo_opq_procrustes(o.fp(eye.copy()), 8, o.fp(R), None, None)
    assert np.array_equal(R, eye)


def test_opq_rotate_roundtrip():
    rng = np.random.default_rng(4)
    dim, bs = 96, 32
    q, _ = np.linalg.qr(rng.standard_normal((bs, bs)))
    rot = np.tile(q.astype(F), (dim // bs, 1, 1))
    x = rng.standard_normal(dim).astype(F)
    y, z = np.zeros(dim, F), np.zeros(dim, F)
    o.lib.vgo_opq_rotate(o.fp(x), dim, bs, o.fp(rot), o.fp(y))
    o.lib.vgo_opq_unrotate(o.fp(y), dim, bs, o.fp(rot), o.fp(z))
    assert np.allclose(y.reshape(-1, bs), x.reshape(-1, bs) @ q.T.astype(F), atol=1e-5)
    assert np.allclose(z, x, atol=1e-5)  # R^T R = I
