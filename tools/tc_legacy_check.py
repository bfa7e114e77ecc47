"""Parity of the single-CTA tensor-core kernels (VECGO_QTC_PAIR=0, VECGO_FLAT_PAIR=0: qtc_kernel, flat_tc_kernel) against the
exact CUDA-core scan.  The switches are read once per process, so tests/test_gpu_quant_tc.py runs this file in a subprocess."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

assert os.environ.get("VECGO_QTC_PAIR") == "0" and os.environ.get("VECGO_FLAT_PAIR") == "0", "run with both pair switches off"
import vecgo_b200 as vg

L = vg._lib
F = np.float32


def stats(name):
    a, b = C.c_uint64(), C.c_uint64()
    L.call(name, C.byref(a), C.byref(b))
    return a.value


def both(make, q, k):
    with make() as ix:
        got = ix.search(q, k)
    L.call("vg_flat_tc_enable", 0)
    try:
        with make() as ix:
            want = ix.search(q, k)
    finally:
        L.call("vg_flat_tc_enable", 1)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)) and np.array_equal(got[2], want[2])


rng = np.random.default_rng(1)
n, dim, nq, k = 20000, 768, 40, 10
v = rng.standard_normal((n, dim)).astype(F)
q = rng.standard_normal((nq, dim)).astype(F)
v[n // 2] = v[n // 3]

sq = vg.quantization.ScalarQuantizer(dim)
sq.Train(v)
codes = sq.EncodeBatch(v)


def mk_sq8():
    ix = vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales))
    ix.upload(codes=codes)
    return ix


s0 = stats("vg_quant_tc_stats")
both(mk_sq8, q, k)
assert stats("vg_quant_tc_stats") - s0 == nq

iq = vg.quantization.Int4Quantizer(dim)
iq.Train(v)
c4 = iq.EncodeBatch(v)


def mk_int4():
    ix = vg.index.DeviceIndex(codec=L.CODEC_INT4, metric=0, dim=dim, rows=n, int4=(iq.min, iq.diff))
    ix.upload(codes=c4)
    return ix


both(mk_int4, q, k)

m = 96
cb = rng.integers(-128, 128, m * 256 * (dim // m), dtype=np.int8)
sc = (rng.random(m) * 0.02 + 0.005).astype(F)
of = (rng.standard_normal(m) * 0.1).astype(F)
pc = rng.integers(0, 256, (n, m), dtype=np.uint8)


def mk_pq():
    ix = vg.index.DeviceIndex(codec=L.CODEC_PQ, metric=0, dim=dim, rows=n, pq=(cb, sc, of, m, 256))
    ix.upload(codes=pc)
    return ix


both(mk_pq, (q * 0.7).astype(F), k)


def mk_flat():
    ix = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n)
    ix.upload(vectors=v)
    return ix


f0 = stats("vg_flat_tc_stats")
both(mk_flat, q, k)
assert stats("vg_flat_tc_stats") - f0 == nq
print("single-CTA tensor-core kernels: ids and scores identical to the exact scan")
