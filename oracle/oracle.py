"""ctypes front-end for the CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``lib``  — oracle/libvecgo_oracle.so, our C restatement (vecgo_oracle.c).
* ``ref``  — oracle/_ref/libvecgo_simd_ref.so, the reference's own AVX-512 C
  kernels compiled by oracle/Makefile from /root/reference/internal/simd/src
  (None when the host CPU lacks AVX-512 or the file was never built).

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this.
Nothing in here touches the GPU, and nothing under vecgo_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libvecgo_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libvecgo_simd_ref.so")

f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u32p = C.POINTER(C.c_uint32)
i64 = C.c_int64
f32 = C.c_float


def build(force: bool = False) -> None:
    """Compile the restatement (and oracle/_ref when /root/reference exists)."""
    src = os.path.join(_HERE, "vecgo_oracle.c")
    stale = (not os.path.exists(_ORACLE_SO)) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src)
    need_ref = os.path.isdir("/root/reference/internal/simd/src") and not os.path.exists(_REF_SO)
    if force or stale or need_ref:
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)


def _cpu_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = set(line.split(":", 1)[1].split())
                    return {"avx512f", "avx512bw", "avx512dq", "avx512vl", "avx512_vpopcntdq"} <= fl
    except OSError:
        pass
    return False


class Cand(C.Structure):
    _fields_ = [("seg", C.c_uint32), ("row", C.c_uint32), ("score", C.c_float), ("approx", C.c_uint32)]


cand_dtype = np.dtype([("seg", "<u4"), ("row", "<u4"), ("score", "<f4"), ("approx", "<u4")])

SQ8_BATCH_FN = C.CFUNCTYPE(None, f32p, u8p, f32p, f32p, i64, i64, f32p)
PAIR_FN = C.CFUNCTYPE(C.c_float, f32p, f32p, i64)
ADC_FN = C.CFUNCTYPE(C.c_float, f32p, u8p, i64)
INT4_BATCH_FN = C.CFUNCTYPE(None, f32p, u8p, i64, i64, f32p, f32p, f32p)
HAMMING_FN = C.CFUNCTYPE(C.c_longlong, u8p, u8p, i64)


class Flat(C.Structure):
    _fields_ = [
        ("rows", i64), ("dim", i64), ("segment_id", C.c_uint32), ("metric", C.c_int), ("quant", C.c_int),
        ("vectors", f32p), ("codes", u8p), ("mins", f32p), ("inv", f32p),
        ("pq_m", i64), ("pq_k", i64), ("pq_codebooks", i8p), ("pq_scales", f32p), ("pq_offsets", f32p),
        ("num_partitions", i64), ("centroids", f32p), ("partition_offsets", u32p),
        ("sq8_batch", C.c_void_p), ("l2", C.c_void_p), ("dot", C.c_void_p), ("adc", C.c_void_p),
        ("l2_ref", C.c_void_p), ("dot_ref", C.c_void_p), ("adc_ref", C.c_void_p), ("adc_offsets", C.c_void_p),
    ]


build()
lib = C.CDLL(_ORACLE_SO)
ref = C.CDLL(_REF_SO) if (os.path.exists(_REF_SO) and _cpu_has_avx512()) else None


def _sig(fn, res, *args):
    fn.restype = res
    fn.argtypes = list(args)
    return fn


_sig(lib.vgo_dot_a512, f32, f32p, f32p, i64)
_sig(lib.vgo_sql2_a512, f32, f32p, f32p, i64)
_sig(lib.vgo_dot_generic, f32, f32p, f32p, i64)
_sig(lib.vgo_sql2_generic, f32, f32p, f32p, i64)
_sig(lib.vgo_sql2_batch_a512, None, f32p, f32p, i64, i64, f32p)
_sig(lib.vgo_dot_batch_a512, None, f32p, f32p, i64, i64, f32p)
_sig(lib.vgo_scale, None, f32p, i64, f32)
_sig(lib.vgo_sqrt, f32, f32)
_sig(lib.vgo_normalize_l2, C.c_int, f32p, i64)
_sig(lib.vgo_pq_adc_generic, f32, f32p, u8p, i64)
_sig(lib.vgo_pq_adc_a512, f32, f32p, u8p, i64)
_sig(lib.vgo_hamming, i64, u8p, u8p, i64)
_sig(lib.vgo_sq8u_l2_batch_generic, None, f32p, u8p, f32p, f32p, i64, i64, f32p)
_sig(lib.vgo_sq8u_l2_batch_a512, None, f32p, u8p, f32p, f32p, i64, i64, f32p)
_sig(lib.vgo_int4_l2_generic, f32, f32p, u8p, i64, f32p, f32p)
_sig(lib.vgo_int4_l2_a512, f32, f32p, u8p, i64, f32p, f32p)
_sig(lib.vgo_int4_l2_batch_a512, None, f32p, u8p, i64, i64, f32p, f32p, f32p)
_sig(lib.vgo_int4_build_lut, None, f32p, f32p, i64, f32p)
_sig(lib.vgo_int4_l2_precomputed_generic, f32, f32p, u8p, i64, f32p)
_sig(lib.vgo_int4_l2_precomputed_a512, f32, f32p, u8p, i64, f32p)
_sig(lib.vgo_squared_l2_bounded_a512, f32, f32p, f32p, i64, f32, i32p)
_sig(lib.vgo_squared_l2_bounded_generic, f32, f32p, f32p, i64, f32, i32p)
_sig(lib.vgo_sql2_int8_dequant, f32, f32p, i8p, i64, f32, f32)
_sig(lib.vgo_build_distance_table_int8, None, f32p, i8p, i64, f32, f32, i64, f32p)
_sig(lib.vgo_find_nearest_centroid_int8, i64, f32p, i8p, i64, i64, f32, f32)
_sig(lib.vgo_sq8_train, C.c_int, f32p, i64, i64, f32p, f32p, f32p, f32p)
_sig(lib.vgo_sq8_set_bounds, None, f32p, f32p, i64, f32p, f32p)
_sig(lib.vgo_sq8_encode, None, f32p, i64, f32p, f32p, f32p, u8p)
_sig(lib.vgo_sq8_decode, None, u8p, i64, f32p, f32p, f32p)
_sig(lib.vgo_sq8_l2_go, f32, f32p, u8p, i64, f32p, f32p)
_sig(lib.vgo_sq8_dot_go, f32, f32p, u8p, i64, f32p, f32p)
_sig(lib.vgo_int4_train, C.c_int, f32p, i64, i64, f32p, f32p)
_sig(lib.vgo_int4_encode, None, f32p, i64, f32p, f32p, u8p)
_sig(lib.vgo_int4_decode, None, u8p, i64, f32p, f32p, f32p)
_sig(lib.vgo_bq_train, f32, f32p, i64, i64)
_sig(lib.vgo_bq_encode, None, f32p, i64, f32, u8p)
_sig(lib.vgo_rabitq_encode, None, f32p, i64, u8p)
_sig(lib.vgo_rabitq_qnorm, f32, f32p, i64)
_sig(lib.vgo_rabitq_estimate, f32, f32, f32, i64, i64)
_sig(lib.vgo_rabitq_distance, f32, f32p, i64, u8p)
_sig(lib.vgo_pq_build_table, None, f32p, i64, i64, i64, i8p, f32p, f32p, f32p)
_sig(lib.vgo_pq_encode, None, f32p, i64, i64, i64, i8p, f32p, f32p, u8p)
_sig(lib.vgo_pq_decode, None, u8p, i64, i64, i64, i8p, f32p, f32p, f32p)
_sig(lib.vgo_pq_asymmetric, f32, f32p, u8p, i64, i64, i64, i8p, f32p, f32p)
_sig(lib.vgo_pq_quantize_centroids, None, f32p, i64, i8p, f32p, f32p)
_sig(lib.vgo_opq_block_size, i64, i64, i64)
_sig(lib.vgo_opq_rotate, None, f32p, i64, i64, f32p, f32p)
_sig(lib.vgo_opq_unrotate, None, f32p, i64, i64, f32p, f32p)
_sig(lib.vgo_opq_accumulate_m, None, f32p, f32p, i64, i64, i64, f32p)
_sig(lib.vgo_opq_procrustes, None, f32p, i64, f32p, f32p, f32p)
_sig(lib.vgo_splitmix64, C.c_uint64, C.c_uint64)
_sig(lib.vgo_rng_intn, i64, C.c_uint64, C.c_uint64, C.c_uint64, i64)
_sig(lib.vgo_rng_f32, f32, C.c_uint64, C.c_uint64, C.c_uint64)
_sig(lib.vgo_pq_kmeanspp_init, None, f32p, i64, i64, i64, i64, i64, C.c_uint64, C.c_uint64, f32p)
_sig(lib.vgo_pq_find_nearest, i64, f32p, f32p, i64, i64)
_sig(lib.vgo_pq_lloyd, i64, f32p, i64, i64, i64, i64, i64, i64, C.c_uint64, C.c_uint64, f32p, i32p)
_sig(lib.vgo_kmeans_assign, i64, f32p, f32p, i64, i64, C.c_int)
_sig(lib.vgo_kmeans_train, i64, f32p, i64, i64, i64, C.c_int, i64, i64p, C.c_uint64, f32p, i32p)
_sig(lib.vgo_find_closest_centroids, i64, f32p, f32p, i64, i64, i64, C.c_int, i64p)
_sig(lib.vgo_heap_topk, i64, C.POINTER(Cand), i64, i64, C.c_int, C.POINTER(Cand))
_sig(lib.vgo_flat_search, i64, C.POINTER(Flat), f32p, i64, i64, u8p, C.POINTER(Cand))
_sig(lib.vgo_flat_rerank, None, C.POINTER(Flat), f32p, u32p, i64, f32p)
_sig(lib.vgo_flat_search_batch, None, C.POINTER(Flat), f32p, i64, i64, i64, u8p, C.c_int, C.POINTER(Cand), i64p)
_sig(lib.vgo_int4_search, i64, f32p, u8p, i64, i64, f32p, f32p, i64, C.c_void_p, C.POINTER(Cand))
_sig(lib.vgo_rabitq_search, i64, f32p, u8p, i64, i64, i64, C.c_void_p, C.POINTER(Cand), i32p)
_sig(lib.vgo_bq_search, i64, u8p, u8p, i64, i64, i64, C.c_void_p, C.POINTER(Cand))
_sig(lib.vgo_int4_search_batch, None, f32p, i64, u8p, i64, i64, f32p, f32p, i64, C.c_void_p, C.c_int,
     C.POINTER(Cand), i64p)
_sig(lib.vgo_rabitq_search_batch, None, f32p, i64, u8p, i64, i64, i64, C.c_void_p, C.c_int, C.POINTER(Cand), i64p)
_sig(lib.vgo_crc32c, C.c_uint32, u8p, i64)

if ref is not None:
    _sig(ref.dotProductAvx512, None, f32p, f32p, i64, f32p)
    _sig(ref.squaredL2Avx512, None, f32p, f32p, i64, f32p)
    _sig(ref.pqAdcLookupAvx512, None, f32p, u8p, i64, f32p, C.c_void_p)
    _sig(ref.scaleAvx512, None, f32p, i64, f32p)
    _sig(ref.squaredL2BatchAvx512, None, f32p, f32p, i64, i64, f32p)
    _sig(ref.dotBatchAvx512, None, f32p, f32p, i64, i64, f32p)
    _sig(ref.sq8uL2BatchPerDimensionAvx512, None, f32p, u8p, f32p, f32p, i64, i64, f32p)
    _sig(ref.int4L2DistanceAvx512, None, f32p, u8p, i64, f32p, f32p, f32p)
    _sig(ref.int4L2DistancePrecomputedAvx512, None, f32p, u8p, i64, f32p, f32p)
    _sig(ref.int4L2DistanceBatchAvx512, None, f32p, u8p, i64, i64, f32p, f32p, f32p)
    _sig(ref.hammingAvx512, C.c_longlong, u8p, u8p, i64)
    _sig(ref.squaredL2BoundedAvx512, None, f32p, f32p, i64, f32, f32p, i32p)
    _sig(ref.squaredL2Int8DequantizedAvx512, None, f32p, i8p, i64, f32p, f32p, f32p)


# ---------------------------------------------------------------- helpers --
def fp(a):
    return a.ctypes.data_as(f32p)


def bp(a):
    return a.ctypes.data_as(u8p)


def c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def cu8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


# the pqAdcLookup asm needs offsets[i] = i*256 (kernels_amd64.go:38-44)
_ADC_OFFSETS = (np.arange(16, dtype=np.int32) * 256).copy()


# ------------------------------------------------------- reference kernels --
def ref_dot(a, b):
    out = np.zeros(1, np.float32)
    ref.dotProductAvx512(fp(a), fp(b), len(a), fp(out))
    return out[0]


def ref_sql2(a, b):
    out = np.zeros(1, np.float32)
    ref.squaredL2Avx512(fp(a), fp(b), len(a), fp(out))
    return out[0]


def ref_pq_adc(table, codes, m):
    out = np.zeros(1, np.float32)
    ref.pqAdcLookupAvx512(fp(table), bp(codes), m, fp(out), _ADC_OFFSETS.ctypes.data)
    return out[0]


def ref_hamming(a, b):
    return int(ref.hammingAvx512(bp(a), bp(b), len(a)))


def fn_addr(f):
    return C.cast(f, C.c_void_p).value


def oracle_kernels():
    """Function-pointer set: this file's restatements."""
    return dict(sq8_batch=fn_addr(lib.vgo_sq8u_l2_batch_a512), l2=fn_addr(lib.vgo_sql2_a512),
                dot=fn_addr(lib.vgo_dot_a512), adc=fn_addr(lib.vgo_pq_adc_a512))


class FlatOracle:
    """flat.(*Segment) restatement driver (internal/segment/flat/segment.go:447-781)."""

    def __init__(self, *, dim, metric=0, segment_id=0, vectors=None, quant=0, codes=None, mins=None, inv=None,
                 pq=None, centroids=None, partition_offsets=None, kernels=None):
        self.keep = []
        s = Flat()
        s.dim = dim
        s.metric = metric
        s.segment_id = segment_id
        s.quant = quant
        rows = None

        def keep(a):
            self.keep.append(a)
            return a

        if vectors is not None:
            v = keep(c32(vectors))
            rows = v.shape[0]
            s.vectors = fp(v)
        if codes is not None:
            c = keep(cu8(codes))
            rows = c.shape[0] if rows is None else rows
            s.codes = bp(c)
        if quant == 1:
            s.mins = fp(keep(c32(mins)))
            s.inv = fp(keep(c32(inv)))
        if quant == 2:
            cb, sc, of, m, k = pq
            s.pq_m, s.pq_k = m, k
            s.pq_codebooks = keep(np.ascontiguousarray(cb, np.int8)).ctypes.data_as(i8p)
            s.pq_scales = fp(keep(c32(sc)))
            s.pq_offsets = fp(keep(c32(of)))
        if centroids is not None:
            cen = keep(c32(centroids))
            s.num_partitions = cen.shape[0]
            s.centroids = fp(cen)
            s.partition_offsets = keep(np.ascontiguousarray(partition_offsets, np.uint32)).ctypes.data_as(u32p)
        s.rows = rows
        kk = kernels or oracle_kernels()
        s.sq8_batch, s.l2, s.dot, s.adc = kk["sq8_batch"], kk["l2"], kk["dot"], kk["adc"]
        s.l2_ref, s.dot_ref, s.adc_ref = kk.get("l2_ref"), kk.get("dot_ref"), kk.get("adc_ref")
        s.adc_offsets = _ADC_OFFSETS.ctypes.data
        self.s = s
        self.rows = rows
        self.dim = dim

    def search_batch(self, queries, k, nprobes=0, mask=None, threads=1):
        q = c32(queries)
        nq = q.shape[0]
        out = np.zeros((nq, max(k, 1)), dtype=cand_dtype)
        counts = np.zeros(nq, np.int64)
        mp = bp(cu8(mask)) if mask is not None else None
        lib.vgo_flat_search_batch(C.byref(self.s), fp(q), nq, k, nprobes, mp, threads,
                                  out.ctypes.data_as(C.POINTER(Cand)), counts.ctypes.data_as(i64p))
        return out[:, :k], counts

    def rerank(self, query, rows):
        q = c32(query)
        r = np.ascontiguousarray(rows, np.uint32)
        out = np.zeros(len(r), np.float32)
        lib.vgo_flat_rerank(C.byref(self.s), fp(q), r.ctypes.data_as(u32p), len(r), fp(out))
        return out


def ref_kernels():
    """Function-pointer set: the REAL reference kernels from oracle/_ref."""
    if ref is None:
        return None
    d = oracle_kernels()
    d.update(sq8_batch=fn_addr(ref.sq8uL2BatchPerDimensionAvx512), l2_ref=fn_addr(ref.squaredL2Avx512),
             dot_ref=fn_addr(ref.dotProductAvx512), adc_ref=fn_addr(ref.pqAdcLookupAvx512))
    return d
