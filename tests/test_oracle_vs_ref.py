"""Pins the scalar restatement (oracle/vecgo_oracle.c) BIT-FOR-BIT to the
reference's own AVX-512 C kernels compiled into oracle/_ref, and repeats the
reference's SIMD≡generic tolerance tests (internal/simd/floats_test.go:195-276,
414-439,471-502; int8_test.go:11-55; int4_test.go)."""
import numpy as np
import pytest

from oracle import oracle as o

F = np.float32
needs_ref = pytest.mark.skipif(o.ref is None, reason="oracle/_ref not built or host lacks AVX-512")
LENGTHS = [1, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 79, 80, 96, 127, 128, 129, 191, 200, 256, 768,
           1000, 1536]


def bits(x):
    return np.asarray(x, F).view(np.uint32)


@needs_ref
def test_pair_kernels_bit_exact():
    rng = np.random.default_rng(1)
    for n in LENGTHS:
        for _ in range(4):
            a, b = rng.standard_normal(n).astype(F), rng.standard_normal(n).astype(F)
            assert bits(o.lib.vgo_dot_a512(o.fp(a), o.fp(b), n)) == bits(o.ref_dot(a, b)), n
            assert bits(o.lib.vgo_sql2_a512(o.fp(a), o.fp(b), n)) == bits(o.ref_sql2(a, b)), n


def test_pair_generic_equivalence():  # floats_test.go:195-216, tol 1e-4
    rng = np.random.default_rng(2)
    for n in [0, 1, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65]:
        a, b = rng.random(max(n, 1)).astype(F) * 2 - 1, rng.random(max(n, 1)).astype(F) * 2 - 1
        assert abs(o.lib.vgo_dot_a512(o.fp(a), o.fp(b), n) - o.lib.vgo_dot_generic(o.fp(a), o.fp(b), n)) <= 1e-4
        assert abs(o.lib.vgo_sql2_a512(o.fp(a), o.fp(b), n) - o.lib.vgo_sql2_generic(o.fp(a), o.fp(b), n)) <= 1e-4


@needs_ref
def test_batch_kernels_bit_exact():
    rng = np.random.default_rng(3)
    for dim in LENGTHS:
        n = 5
        q = rng.standard_normal(dim).astype(F)
        t = rng.standard_normal((n, dim)).astype(F)
        for mine, theirs in ((o.lib.vgo_sql2_batch_a512, o.ref.squaredL2BatchAvx512),
                             (o.lib.vgo_dot_batch_a512, o.ref.dotBatchAvx512)):
            x, y = np.zeros(n, F), np.zeros(n, F)
            mine(o.fp(q), o.fp(t), dim, n, o.fp(x))
            theirs(o.fp(q), o.fp(t), dim, n, o.fp(y))
            assert np.array_equal(bits(x), bits(y)), dim


@needs_ref
def test_sq8_bit_exact_and_generic_tol():
    rng = np.random.default_rng(4)
    for dim in [1, 7, 8, 15, 16, 17, 31, 32, 33, 64, 100, 768]:
        for n in (1, 2, 5):
            q = (rng.random(dim) * 2 - 1).astype(F)
            mins = (rng.random(dim) * 2 - 1).astype(F)
            inv = (rng.random(dim) * 2 - 1).astype(F)
            codes = rng.integers(0, 256, (n, dim), dtype=np.uint8)
            x, y, g = np.zeros(n, F), np.zeros(n, F), np.zeros(n, F)
            o.lib.vgo_sq8u_l2_batch_a512(o.fp(q), o.bp(codes), o.fp(mins), o.fp(inv), dim, n, o.fp(x))
            o.ref.sq8uL2BatchPerDimensionAvx512(o.fp(q), o.bp(codes), o.fp(mins), o.fp(inv), dim, n, o.fp(y))
            o.lib.vgo_sq8u_l2_batch_generic(o.fp(q), o.bp(codes), o.fp(mins), o.fp(inv), dim, n, o.fp(g))
            assert np.array_equal(bits(x), bits(y)), dim
            assert np.all(np.abs(x - g) <= np.maximum(5e-2, 1e-5 * np.abs(g)))  # floats_test.go:471-502 (5e-2 abs at dim<=33)


@needs_ref
def test_int4_bit_exact_and_generic_tol():
    rng = np.random.default_rng(5)
    for dim in [2, 6, 16, 30, 32, 34, 62, 64, 66, 96, 128, 130, 768, 1536, 7, 33, 65]:
        n = 4
        cs = (dim + 1) // 2
        q = rng.standard_normal(dim).astype(F)
        minv = rng.standard_normal(dim).astype(F)
        diff = (rng.random(dim) * 3 + 0.1).astype(F)
        codes = rng.integers(0, 256, (n, cs), dtype=np.uint8)
        x, y = np.zeros(n, F), np.zeros(n, F)
        o.lib.vgo_int4_l2_batch_a512(o.fp(q), o.bp(codes), dim, n, o.fp(minv), o.fp(diff), o.fp(x))
        o.ref.int4L2DistanceBatchAvx512(o.fp(q), o.bp(codes), dim, n, o.fp(minv), o.fp(diff), o.fp(y))
        assert np.array_equal(bits(x), bits(y)), dim
        one = np.zeros(1, F)
        o.ref.int4L2DistanceAvx512(o.fp(q), o.bp(codes[0]), dim, o.fp(minv), o.fp(diff), o.fp(one))
        assert bits(one[0]) == bits(o.lib.vgo_int4_l2_a512(o.fp(q), o.bp(codes[0]), dim, o.fp(minv), o.fp(diff)))
        g = o.lib.vgo_int4_l2_generic(o.fp(q), o.bp(codes[0]), dim, o.fp(minv), o.fp(diff))
        assert abs(g - x[0]) <= 1e-3 * max(1.0, abs(g))  # int4_test.go tolerance
        if dim % 2 == 0:
            lut = np.zeros(dim * 16, F)
            o.lib.vgo_int4_build_lut(o.fp(minv), o.fp(diff), dim, o.fp(lut))
            o.ref.int4L2DistancePrecomputedAvx512(o.fp(q), o.bp(codes[0]), dim, o.fp(lut), o.fp(one))
            assert bits(one[0]) == bits(o.lib.vgo_int4_l2_precomputed_a512(o.fp(q), o.bp(codes[0]), dim, o.fp(lut)))


@needs_ref
def test_pq_adc_bit_exact():
    rng = np.random.default_rng(6)
    for m in [1, 2, 7, 8, 9, 15, 16, 17, 32, 48, 96, 100]:
        table = (rng.random(m * 256) * 2 - 1).astype(F)
        codes = rng.integers(0, 256, m, dtype=np.uint8)
        a = o.lib.vgo_pq_adc_a512(o.fp(table), o.bp(codes), m)
        assert bits(a) == bits(o.ref_pq_adc(table, codes, m)), m
        assert abs(a - o.lib.vgo_pq_adc_generic(o.fp(table), o.bp(codes), m)) <= 1e-4  # floats_test.go:414-439


@needs_ref
def test_int8_helpers_match_ref_asm_within_tol():
    """The asm int8-PQ helpers are compiled but unregistered; the live path is
    the generic one we restate.  int8_test.go:11-55: dist tol max(1e-2,1e-6|x|)."""
    rng = np.random.default_rng(8)
    for sub in (4, 8, 16, 32):
        q = rng.standard_normal(sub).astype(F)
        code = rng.integers(-128, 128, sub, dtype=np.int8)
        sc, of = np.array([0.013], F), np.array([0.2], F)
        out = np.zeros(1, F)
        o.ref.squaredL2Int8DequantizedAvx512(o.fp(q), code.ctypes.data_as(o.i8p), sub, o.fp(sc), o.fp(of), o.fp(out))
        g = o.lib.vgo_sql2_int8_dequant(o.fp(q), code.ctypes.data_as(o.i8p), sub, sc[0], of[0])
        assert abs(g - out[0]) <= max(1e-2, 1e-6 * abs(g))


def test_flat_search_oracle_vs_numpy_bruteforce():
    """flat.Search restatement returns exactly the k best under (score,row)."""
    rng = np.random.default_rng(9)
    n, d, nq, k = 3000, 32, 8, 10
    x = rng.random((n, d)).astype(F)
    q = rng.random((nq, d)).astype(F)
    seg = o.FlatOracle(dim=d, metric=0, vectors=x)
    out, cnt = seg.search_batch(q, k)
    for i in range(nq):
        sc = np.array([o.lib.vgo_sql2_a512(o.fp(q[i]), o.fp(x[j]), d) for j in range(n)], F)
        order = np.lexsort((np.arange(n), sc))[:k]
        assert np.array_equal(out[i]["row"], order)
        assert np.array_equal(bits(out[i]["score"]), bits(sc[order]))
    if o.ref is not None:
        seg2 = o.FlatOracle(dim=d, metric=0, vectors=x, kernels=o.ref_kernels())
        out2, _ = seg2.search_batch(q, k, threads=4)
        assert np.array_equal(out2["row"], out["row"]) and np.array_equal(bits(out2["score"]), bits(out["score"]))


@needs_ref
def test_bounded_l2_bit_exact():
    """simd.SquaredL2Bounded (bounded_l2_avx512.c:19-107): the early-exit partial totals, the flag and the tails."""
    import ctypes as C

    rng = np.random.default_rng(21)
    for n in LENGTHS + [130, 136, 199, 640, 769]:
        for _ in range(4):
            a, b = rng.standard_normal(n).astype(F), rng.standard_normal(n).astype(F)
            full = float(np.sum((a.astype(np.float64) - b) ** 2))
            for bound in (np.inf, 0.0, full * 0.25, full * 0.5, full * 0.99, full * 1.01, -1.0, np.nan):
                ex1, ex2 = C.c_int32(-1), C.c_int32(-1)
                mine = o.lib.vgo_squared_l2_bounded_a512(o.fp(a), o.fp(b), n, F(bound), C.byref(ex1))
                theirs = np.zeros(1, F)
                o.ref.squaredL2BoundedAvx512(o.fp(a), o.fp(b), n, F(bound), o.fp(theirs), C.byref(ex2))
                assert bits(mine) == bits(theirs[0]), (n, bound)
                assert ex1.value == ex2.value, (n, bound)


def test_bounded_l2_generic_semantics():
    """distance_test / kernels.go:178-217: full distance when the bound is not exceeded, early exit otherwise; the SIMD
    and generic paths agree within the reference's 1e-4 relative tolerance on the full distance."""
    import ctypes as C

    rng = np.random.default_rng(22)
    for n in [1, 8, 63, 64, 65, 128, 200, 768]:
        a, b = rng.standard_normal(n).astype(F), rng.standard_normal(n).astype(F)
        ex = C.c_int32()
        full = o.lib.vgo_squared_l2_bounded_generic(o.fp(a), o.fp(b), n, F(np.inf), C.byref(ex))
        assert ex.value == 0
        assert abs(full - o.lib.vgo_sql2_generic(o.fp(a), o.fp(b), n)) <= 1e-4 * max(1.0, full)
        simd = o.lib.vgo_squared_l2_bounded_a512(o.fp(a), o.fp(b), n, F(np.inf), C.byref(ex))
        assert abs(simd - full) <= 1e-4 * max(1.0, full)
        part = o.lib.vgo_squared_l2_bounded_generic(o.fp(a), o.fp(b), n, F(full * 0.1), C.byref(ex))
        assert ex.value == 1 and part <= full * (1 + 1e-6)
