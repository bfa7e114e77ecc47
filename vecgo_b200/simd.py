"""Mirror of vecgo's kernel table (internal/simd/kernels.go:39-123) over the CUDA library.

Same names and argument meaning as the Go functions so parity tests read like
the reference's own (internal/simd/floats_test.go).  Every call crosses the
C ABI once with the whole batch; the single-pair forms exist for tests only
(in production only batch/segment-level calls cross cgo).
"""
from __future__ import annotations

import numpy as np

from . import _lib as L

F = np.float32


def _pairs(fn, a, b):
    a, b = L.as_f32(a), L.as_f32(b)
    if a.shape != b.shape:
        raise ValueError("length mismatch")
    a2 = a.reshape(-1, a.shape[-1]) if a.ndim > 1 else a.reshape(1, -1)
    b2 = b.reshape(a2.shape)
    out = np.zeros(a2.shape[0], F)
    L.call(fn, L.ptr(a2, L.f32p), L.ptr(b2, L.f32p), a2.shape[0], a2.shape[1], L.ptr(out, L.f32p))
    return out


def Dot(a, b) -> np.float32:
    """simd.Dot (kernels.go:39-43)."""
    return _pairs("vg_simd_dot", a, b)[0]


def SquaredL2(a, b) -> np.float32:
    """simd.SquaredL2 (kernels.go:45-48)."""
    return _pairs("vg_simd_squared_l2", a, b)[0]


def DotPairs(a, b) -> np.ndarray:
    return _pairs("vg_simd_dot", a, b)


def SquaredL2Pairs(a, b) -> np.ndarray:
    return _pairs("vg_simd_squared_l2", a, b)


def ScaleInPlace(a: np.ndarray, scalar: float) -> None:
    """simd.ScaleInPlace (kernels.go:51)."""
    if a.size == 0:
        return
    buf = L.as_f32(a).copy()
    L.call("vg_simd_scale", L.ptr(buf, L.f32p), buf.size, float(scalar))
    a[...] = buf.reshape(a.shape)


def _batch(fn, queries, targets, dim):
    q = L.as_f32(queries).reshape(-1, dim) if dim else L.as_f32(queries).reshape(1, 0)
    t = L.as_f32(targets).reshape(-1, dim) if dim else L.as_f32(targets).reshape(0, 0)
    out = np.zeros((q.shape[0], t.shape[0]), F)
    L.call(fn, L.ptr(q, L.f32p), q.shape[0], L.ptr(t, L.f32p), t.shape[0], dim, L.ptr(out, L.f32p))
    return out


def DotBatch(query, targets, dim: int, out=None):
    """simd.DotBatch (kernels.go:61-63); `query` may hold several queries."""
    r = _batch("vg_simd_dot_batch", query, targets, dim)
    if out is not None:
        out[...] = r.reshape(out.shape)
    return r


def SquaredL2Batch(query, targets, dim: int, out=None):
    """simd.SquaredL2Batch (kernels.go:66-68)."""
    r = _batch("vg_simd_squared_l2_batch", query, targets, dim)
    if out is not None:
        out[...] = r.reshape(out.shape)
    return r


def Sq8uL2BatchPerDimension(query, codes, mins, invScales, dim: int, out=None):
    """simd.Sq8uL2BatchPerDimension (kernels.go:76-78)."""
    q = L.as_f32(query).reshape(-1, dim)
    c = L.as_u8(codes).reshape(-1, dim)
    mn, iv = L.as_f32(mins), L.as_f32(invScales)
    r = np.zeros((q.shape[0], c.shape[0]), F)
    L.call("vg_simd_sq8u_l2_batch", L.ptr(q, L.f32p), q.shape[0], L.ptr(c, L.u8p), c.shape[0], dim, L.ptr(mn, L.f32p),
           L.ptr(iv, L.f32p), L.ptr(r, L.f32p))
    if out is not None:
        out[...] = r.reshape(out.shape)
    return r


def Int4L2DistanceBatch(query, codes, dim: int, n: int, minVal, diff, out=None):
    """simd.Int4L2DistanceBatch (kernels.go:106-108)."""
    cs = (dim + 1) // 2
    q = L.as_f32(query).reshape(-1, dim)
    c = L.as_u8(codes).reshape(-1)[: n * cs].reshape(n, cs)
    mn, df = L.as_f32(minVal), L.as_f32(diff)
    r = np.zeros((q.shape[0], n), F)
    L.call("vg_simd_int4_l2_batch", L.ptr(q, L.f32p), q.shape[0], L.ptr(c, L.u8p), n, dim, L.ptr(mn, L.f32p), L.ptr(df, L.f32p),
           L.ptr(r, L.f32p))
    if out is not None:
        out[...] = r.reshape(out.shape)
    return r


def Int4L2Distance(query, code, minVal, diff) -> np.float32:
    """simd.Int4L2Distance (kernels.go:81-83)."""
    return Int4L2DistanceBatch(query, code, len(query), 1, minVal, diff)[0, 0]


def PqAdcLookupBatch(tables, codes, m: int) -> np.ndarray:
    """simd.PqAdcLookup for [nq] tables x [n] codes."""
    t = L.as_f32(tables).reshape(-1, m * 256) if m else L.as_f32(tables).reshape(1, 0)
    c = L.as_u8(codes).reshape(-1, m) if m else L.as_u8(codes).reshape(1, 0)
    r = np.zeros((t.shape[0], c.shape[0]), F)
    L.call("vg_simd_pq_adc_lookup", L.ptr(t, L.f32p), t.shape[0], L.ptr(c, L.u8p), c.shape[0], m, L.ptr(r, L.f32p))
    return r


def PqAdcLookup(table, codes, m: int) -> np.float32:
    """simd.PqAdcLookup (kernels.go:56-58)."""
    if m == 0:
        return F(0)
    return PqAdcLookupBatch(table, codes, m)[0, 0]


def HammingBatch(queries, codes, nbytes: int) -> np.ndarray:
    q = L.as_u8(queries).reshape(-1, nbytes) if nbytes else L.as_u8(queries).reshape(1, 0)
    c = L.as_u8(codes).reshape(-1, nbytes) if nbytes else L.as_u8(codes).reshape(1, 0)
    r = np.zeros((q.shape[0], c.shape[0]), np.int32)
    L.call("vg_simd_hamming", L.ptr(q, L.u8p), q.shape[0], L.ptr(c, L.u8p), c.shape[0], nbytes, L.ptr(r, L.i32p))
    return r


def Hamming(a, b) -> int:
    """simd.Hamming (kernels.go:71-73)."""
    a, b = L.as_u8(a).reshape(-1), L.as_u8(b).reshape(-1)
    if a.size == 0:
        return 0
    return int(HammingBatch(a, b, a.size)[0, 0])
