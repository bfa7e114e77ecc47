"""Mirror of vecgo's `distance` package (distance/distance.go) over the CUDA library."""
from __future__ import annotations

import enum

import numpy as np

from . import _lib as L
from . import simd


class Metric(enum.IntEnum):  # distance.go:68-73
    L2 = 0
    Cosine = 1
    Dot = 2
    Hamming = 3


MetricL2, MetricCosine, MetricDot, MetricHamming = Metric.L2, Metric.Cosine, Metric.Dot, Metric.Hamming


def Dot(a, b):
    return simd.Dot(a, b)


def SquaredL2(a, b):
    return simd.SquaredL2(a, b)


def Hamming(a, b) -> np.float32:
    return np.float32(simd.Hamming(a, b))


def NormalizeL2InPlace(v: np.ndarray) -> bool:
    """distance.NormalizeL2InPlace (distance.go:42-53). Returns False for zero norm / empty."""
    if v.size == 0:
        return False
    buf = L.as_f32(v).reshape(1, -1).copy()
    ok = np.zeros(1, np.uint8)
    L.call("vg_normalize_l2", L.ptr(buf, L.f32p), 1, buf.shape[1], L.ptr(ok, L.u8p))
    if ok[0]:
        v[...] = buf.reshape(v.shape)
    return bool(ok[0])


def NormalizeL2Batch(vecs: np.ndarray):
    """Row-wise NormalizeL2InPlace; returns (normalized copy, ok mask)."""
    buf = L.as_f32(vecs).copy()
    ok = np.zeros(buf.shape[0], np.uint8)
    if buf.size:
        L.call("vg_normalize_l2", L.ptr(buf, L.f32p), buf.shape[0], buf.shape[1], L.ptr(ok, L.u8p))
    return buf, ok.astype(bool)


def Provider(m: Metric):
    """distance.Provider (distance.go:97-106)."""
    if m == Metric.L2:
        return SquaredL2
    if m in (Metric.Cosine, Metric.Dot):
        return Dot
    raise ValueError(f"unsupported metric for float32: {m!r}")
