"""Mirror of internal/quantization (Quantizer interface + concrete quantizers)
over the CUDA library.  Method names, argument meaning and error behaviour
follow the Go types; batch variants (`*Batch`) take [n x dim] arrays and are
what production code should call (one C-ABI crossing per batch).
"""
from __future__ import annotations

import ctypes as C
import enum
import struct

import numpy as np

from . import _lib as L
from . import simd

F = np.float32


class Type(enum.IntEnum):  # internal/quantization/types.go:6-14
    NONE = 0
    PQ = 1
    OPQ = 2
    SQ8 = 3
    BQ = 4
    RaBitQ = 5
    INT4 = 6


class QuantizerError(ValueError):
    pass


def _rows(v, dim):
    a = L.as_f32(v)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.shape[1] != dim:
        raise QuantizerError("vector dimension mismatch")
    return a


# ===================================================================== SQ8
class ScalarQuantizer:
    """quantization.ScalarQuantizer (quantizer.go:26-327)."""

    def __init__(self, dimension: int):
        self.dimension = dimension
        self.trained = False
        self.mins = self.maxs = self.scales = self.invScales = None

    def Mins(self):
        return self.mins

    def Maxs(self):
        return self.maxs

    def SetBounds(self, mins, maxs):
        mins, maxs = L.as_f32(mins), L.as_f32(maxs)
        if len(mins) != self.dimension or len(maxs) != self.dimension:
            raise QuantizerError("dimension mismatch")
        self.mins, self.maxs = mins.copy(), maxs.copy()
        self.scales, self.invScales = np.zeros(self.dimension, F), np.zeros(self.dimension, F)
        L.call("vg_sq8_set_bounds", L.ptr(self.mins, L.f32p), L.ptr(self.maxs, L.f32p), self.dimension,
               L.ptr(self.scales, L.f32p), L.ptr(self.invScales, L.f32p))
        self.trained = True

    def Train(self, vectors):
        v = L.as_f32(vectors)
        if v.size == 0:
            raise QuantizerError("no vectors provided for training")
        if v.ndim != 2 or v.shape[1] != self.dimension:
            raise QuantizerError("vector dimension mismatch")
        d = self.dimension
        self.mins, self.maxs, self.scales, self.invScales = (np.zeros(d, F) for _ in range(4))
        L.call("vg_sq8_train", L.ptr(v, L.f32p), v.shape[0], d, L.ptr(self.mins, L.f32p), L.ptr(self.maxs, L.f32p),
               L.ptr(self.scales, L.f32p), L.ptr(self.invScales, L.f32p))
        self.trained = True

    def EncodeBatch(self, vectors) -> np.ndarray:
        if not self.trained:
            raise QuantizerError("ScalarQuantizer not trained")
        v = _rows(vectors, self.dimension)
        out = np.zeros(v.shape, np.uint8)
        L.call("vg_sq8_encode", L.ptr(v, L.f32p), v.shape[0], self.dimension, L.ptr(self.mins, L.f32p),
               L.ptr(self.maxs, L.f32p), L.ptr(self.scales, L.f32p), L.ptr(out, L.u8p))
        return out

    def Encode(self, v) -> np.ndarray:
        return self.EncodeBatch(v)[0]

    def DecodeBatch(self, codes) -> np.ndarray:
        if not self.trained:
            raise QuantizerError("ScalarQuantizer not trained")
        c = L.as_u8(codes)
        if c.ndim == 1:
            c = c.reshape(1, -1)
        if c.shape[1] != self.dimension:
            raise QuantizerError("vector dimension mismatch")
        out = np.zeros(c.shape, F)
        L.call("vg_sq8_decode", L.ptr(c, L.u8p), c.shape[0], self.dimension, L.ptr(self.mins, L.f32p),
               L.ptr(self.invScales, L.f32p), L.ptr(out, L.f32p))
        return out

    def Decode(self, b) -> np.ndarray:
        return self.DecodeBatch(b)[0]

    def L2DistanceBatch(self, q, codes, n: int, out=None):
        q = L.as_f32(q)
        if q.shape[-1] != self.dimension:
            raise QuantizerError("query dimension mismatch")
        c = L.as_u8(codes).reshape(-1)
        if c.size < n * self.dimension:
            raise QuantizerError("codes buffer too small")
        if out is not None and len(out) < n:
            raise QuantizerError("output buffer too small")
        r = simd.Sq8uL2BatchPerDimension(q, c[: n * self.dimension], self.mins, self.invScales, self.dimension)
        if out is not None:
            out[:n] = r[0]
        return r if q.ndim > 1 else r[0]

    def BytesPerDimension(self) -> int:
        return 1

    def Min(self, dim):
        return F(0) if (not self.trained or dim < 0 or dim >= self.dimension) else self.mins[dim]

    def Max(self, dim):
        return F(0) if (not self.trained or dim < 0 or dim >= self.dimension) else self.maxs[dim]

    def MarshalBinary(self) -> bytes:  # quantizer.go:275-290: dim u32, interleaved (min,max)
        if not self.trained:
            raise QuantizerError("ScalarQuantizer not trained")
        inter = np.empty(self.dimension * 2, "<f4")
        inter[0::2], inter[1::2] = self.mins, self.maxs
        return struct.pack("<I", self.dimension) + inter.tobytes()

    def UnmarshalBinary(self, data: bytes):  # quantizer.go:293-327 (note: Train's min==max rule, not SetBounds')
        if len(data) < 4:
            raise QuantizerError("invalid scalar quantizer binary length")
        d = struct.unpack_from("<I", data)[0]
        if len(data) != 4 + d * 8:
            raise QuantizerError("invalid scalar quantizer binary length for dimension")
        inter = np.frombuffer(data, "<f4", offset=4)
        self.dimension = d
        self.mins, self.maxs = inter[0::2].copy(), inter[1::2].copy()
        eq = self.mins == self.maxs
        self.maxs[eq] = self.mins[eq] + F(1e-6)
        rng = self.maxs - self.mins
        self.scales, self.invScales = (F(255.0) / rng).astype(F), (rng / F(255.0)).astype(F)
        self.trained = True


# ===================================================================== INT4
class Int4Quantizer:
    """quantization.Int4Quantizer (int4.go)."""

    def __init__(self, dim: int):
        self.dim = dim
        self.min = self.diff = None

    def Train(self, vectors):
        v = L.as_f32(vectors)
        if v.size == 0:
            raise QuantizerError("no vectors provided for training")
        self.dim = v.shape[1]
        self.min, self.diff = np.zeros(self.dim, F), np.zeros(self.dim, F)
        L.call("vg_int4_train", L.ptr(v, L.f32p), v.shape[0], self.dim, L.ptr(self.min, L.f32p), L.ptr(self.diff, L.f32p))

    def EncodeBatch(self, vectors) -> np.ndarray:
        a = L.as_f32(vectors)
        if a.ndim == 1:
            a = a.reshape(1, -1)
        if a.shape[1] != self.dim:
            raise QuantizerError("dimension mismatch")
        out = np.zeros((a.shape[0], (self.dim + 1) // 2), np.uint8)
        L.call("vg_int4_encode", L.ptr(a, L.f32p), a.shape[0], self.dim, L.ptr(self.min, L.f32p), L.ptr(self.diff, L.f32p),
               L.ptr(out, L.u8p))
        return out

    def Encode(self, v):
        return self.EncodeBatch(v)[0]

    def DecodeBatch(self, codes) -> np.ndarray:
        c = L.as_u8(codes)
        if c.ndim == 1:
            c = c.reshape(1, -1)
        if c.shape[1] != (self.dim + 1) // 2:
            raise QuantizerError("dimension mismatch")
        out = np.zeros((c.shape[0], self.dim), F)
        L.call("vg_int4_decode", L.ptr(c, L.u8p), c.shape[0], self.dim, L.ptr(self.min, L.f32p), L.ptr(self.diff, L.f32p),
               L.ptr(out, L.f32p))
        return out

    def Decode(self, b):
        return self.DecodeBatch(b)[0]

    def L2DistanceBatch(self, query, codes, n: int, out=None):
        cs = (self.dim + 1) // 2
        c = L.as_u8(codes).reshape(-1)
        if c.size < n * cs:
            raise QuantizerError("codes buffer too small")
        if out is not None and len(out) < n:
            raise QuantizerError("output buffer too small")
        r = simd.Int4L2DistanceBatch(query, c, self.dim, n, self.min, self.diff)
        if out is not None:
            out[:n] = r[0]
        return r if np.ndim(query) > 1 else r[0]

    def BytesPerDimension(self) -> int:
        return 0

    def MarshalBinary(self) -> bytes:  # int4.go:171-188
        return struct.pack("<I", self.dim) + self.min.astype("<f4").tobytes() + self.diff.astype("<f4").tobytes()

    def UnmarshalBinary(self, data: bytes):
        if len(data) < 4:
            raise QuantizerError("data too short")
        d = struct.unpack_from("<I", data)[0]
        if len(data) != 4 + d * 8:
            raise QuantizerError("data size mismatch")
        arr = np.frombuffer(data, "<f4", offset=4)
        self.dim, self.min, self.diff = d, arr[:d].copy(), arr[d:].copy()


# ===================================================================== BQ
class BinaryQuantizer:
    """quantization.BinaryQuantizer (binary.go)."""

    def __init__(self, dimension: int):
        self.dimension = dimension
        self.threshold = F(0.0)
        self.trained = False

    def WithThreshold(self, threshold: float):
        self.threshold = F(threshold)
        self.trained = True
        return self

    def Train(self, vectors):
        v = L.as_f32(vectors)
        if v.size == 0:
            raise QuantizerError("no vectors provided for training")
        t = np.zeros(1, F)
        L.call("vg_bq_train", L.ptr(v, L.f32p), v.shape[0], v.shape[1], L.ptr(t, L.f32p))
        self.threshold = t[0]
        self.trained = True

    def EncodeBatch(self, vectors) -> np.ndarray:
        v = _rows(vectors, self.dimension)
        nb = ((self.dimension + 63) // 64) * 8
        out = np.zeros((v.shape[0], nb), np.uint8)
        L.call("vg_bq_encode", L.ptr(v, L.f32p), v.shape[0], self.dimension, float(self.threshold), L.ptr(out, L.u8p))
        return out

    def Encode(self, v):
        return self.EncodeBatch(v)[0]

    def EncodeUint64(self, v):
        return self.Encode(v).view("<u8")

    def ComputeHammingDistance(self, query, codes_u64) -> int:
        q = self.EncodeUint64(query)
        return HammingDistance(q, codes_u64)

    def Decode(self, b):  # binary.go:175-189 — reconstruction is trivial host bit math
        bits = np.unpackbits(L.as_u8(b), bitorder="little")[: self.dimension]
        return np.where(bits > 0, self.threshold + F(0.5), self.threshold - F(0.5)).astype(F)

    def BytesPerDimension(self) -> int:
        return 0

    def BytesTotal(self) -> int:
        return (self.dimension + 7) // 8

    def Dimension(self):
        return self.dimension

    def Threshold(self):
        return self.threshold

    def IsTrained(self):
        return self.trained


def HammingDistance(a_u64, b_u64) -> int:
    """quantization.HammingDistance (binary.go:221-243)."""
    a = np.ascontiguousarray(a_u64, "<u8")
    b = np.ascontiguousarray(b_u64, "<u8")
    n = min(len(a), len(b))
    if n == 0:
        return 0
    return simd.Hamming(a[:n].view(np.uint8), b[:n].view(np.uint8))


def HammingDistanceBytes(a, b) -> int:
    a, b = L.as_u8(a), L.as_u8(b)
    n = min(len(a), len(b))
    if n == 0:
        return 0
    return simd.Hamming(a[:n], b[:n])


def NormalizedHammingDistance(a, b, dimension: int) -> np.float32:
    return F(HammingDistance(a, b)) / F(dimension)


# ===================================================================== RaBitQ
class RaBitQuantizer:
    """quantization.RaBitQuantizer (rabitq.go): sign bits ‖ float32 norm, no rotation."""

    def __init__(self, dimension: int):
        self.dimension = dimension
        self.threshold = F(0.0)

    def BytesTotal(self) -> int:
        return ((self.dimension + 63) // 64) * 8 + 4

    def BytesPerDimension(self) -> int:
        return 0

    def Train(self, vectors):
        return None

    def EncodeBatch(self, vectors) -> np.ndarray:
        v = _rows(vectors, self.dimension)
        out = np.zeros((v.shape[0], self.BytesTotal()), np.uint8)
        L.call("vg_rabitq_encode", L.ptr(v, L.f32p), v.shape[0], self.dimension, L.ptr(out, L.u8p))
        return out

    def Encode(self, v):
        return self.EncodeBatch(v)[0]

    def DistanceBatch(self, queries, codes) -> np.ndarray:
        """rq.Distance for every (query, code) pair, via a throw-away device index."""
        from .index import DeviceIndex

        c = L.as_u8(codes).reshape(-1, self.BytesTotal())
        q = _rows(queries, self.dimension)
        ix = DeviceIndex(codec=L.CODEC_RABITQ, metric=L.METRIC_L2, dim=self.dimension, rows=c.shape[0])
        try:
            ix.upload(codes=c)
            rows, scores, counts = ix.search(q, c.shape[0])
        finally:
            ix.close()
        out = np.zeros((q.shape[0], c.shape[0]), F)
        for i in range(q.shape[0]):
            out[i, rows[i, : counts[i]]] = scores[i, : counts[i]]
        return out

    def Distance(self, query, code) -> np.float32:
        c = L.as_u8(code)
        if c.size < self.BytesTotal():
            raise QuantizerError("invalid code length")
        return self.DistanceBatch(query, c[: self.BytesTotal()])[0, 0]


# ===================================================================== PQ
class ProductQuantizer:
    """quantization.ProductQuantizer (pq.go): int8 codebooks + per-subspace scale/offset."""

    def __init__(self, dimension: int, numSubvectors: int, numCentroids: int):
        if dimension <= 0 or numSubvectors <= 0:
            raise QuantizerError("dimension and numSubvectors must be positive")
        if dimension % numSubvectors != 0:
            raise QuantizerError("dimension must be divisible by numSubvectors")
        if numCentroids <= 0:
            raise QuantizerError("numCentroids must be positive")
        if numCentroids > 256:
            raise QuantizerError("numCentroids must be <= 256 for uint8 encoding")
        self.dimension, self.numSubvectors, self.numCentroids = dimension, numSubvectors, numCentroids
        self.subvectorDim = dimension // numSubvectors
        self.codebooks = np.zeros(numSubvectors * numCentroids * self.subvectorDim, np.int8)
        self.scales = np.zeros(numSubvectors, F)
        self.offsets = np.zeros(numSubvectors, F)
        self.centroids_f32 = None
        self.trained = False

    def _p(self):
        return (self.dimension, self.numSubvectors, self.numCentroids, L.ptr(self.codebooks, L.i8p), L.ptr(self.scales, L.f32p),
                L.ptr(self.offsets, L.f32p))

    def Train(self, vectors, iters: int = 20, seed: int = 0):
        """pq.go:68-143; 20 Lloyd iterations is the reference's hard-coded constant (pq.go:95)."""
        v = L.as_f32(vectors)
        if v.size == 0:
            raise QuantizerError("no vectors provided for training")
        if v.shape[1] != self.dimension:
            raise QuantizerError("vector dimension mismatch")
        self.centroids_f32 = np.zeros((self.numSubvectors, self.numCentroids, self.subvectorDim), F)
        L.call("vg_pq_train", L.ptr(v, L.f32p), v.shape[0], self.dimension, self.numSubvectors, self.numCentroids, iters, seed,
               L.ptr(self.codebooks, L.i8p), L.ptr(self.scales, L.f32p), L.ptr(self.offsets, L.f32p),
               L.ptr(self.centroids_f32, L.f32p))
        self.trained = True

    def SetCodebooks(self, codebooks, scales, offsets):
        self.codebooks = np.ascontiguousarray(codebooks, np.int8).reshape(-1)
        self.scales, self.offsets = L.as_f32(scales), L.as_f32(offsets)
        self.trained = True

    def Codebooks(self):
        return self.codebooks, self.scales, self.offsets

    def EncodeBatch(self, vectors) -> np.ndarray:
        if not self.trained:
            raise QuantizerError("ProductQuantizer not trained")
        v = _rows(vectors, self.dimension)
        out = np.zeros((v.shape[0], self.numSubvectors), np.uint8)
        L.call("vg_pq_encode", L.ptr(v, L.f32p), v.shape[0], *self._p(), L.ptr(out, L.u8p))
        return out

    def Encode(self, vec):
        return self.EncodeBatch(vec)[0]

    def DecodeBatch(self, codes) -> np.ndarray:
        if not self.trained:
            raise QuantizerError("ProductQuantizer not trained")
        c = L.as_u8(codes)
        if c.ndim == 1:
            c = c.reshape(1, -1)
        if c.shape[1] != self.numSubvectors:
            raise QuantizerError("invalid code length")
        out = np.zeros((c.shape[0], self.dimension), F)
        L.call("vg_pq_decode", L.ptr(c, L.u8p), c.shape[0], *self._p(), L.ptr(out, L.f32p))
        return out

    def Decode(self, codes):
        return self.DecodeBatch(codes)[0]

    def BuildDistanceTable(self, query) -> np.ndarray:
        q = L.as_f32(query)
        single = q.ndim == 1
        q = q.reshape(-1, q.shape[-1])
        if q.shape[1] != self.dimension:
            raise QuantizerError(f"query dimension mismatch: expected {self.dimension}, got {q.shape[1]}")
        out = np.zeros((q.shape[0], self.numSubvectors * self.numCentroids), F)
        L.call("vg_pq_build_distance_table", L.ptr(q, L.f32p), q.shape[0], *self._p(), L.ptr(out, L.f32p))
        return out[0] if single else out

    def AdcDistance(self, table, codes) -> np.float32:
        c = L.as_u8(codes)
        if c.size != self.numSubvectors:
            raise QuantizerError("codes length mismatch")
        return simd.PqAdcLookup(table, c, self.numSubvectors)

    def ComputeAsymmetricDistance(self, query, codes) -> np.float32:
        """pq.go:234-260: dist += SquaredL2Int8Dequantized(q_m, centroid_m) over subspaces in order.  Every term equals the
        distance-table entry (same scalar arithmetic, kernels.go:354-374); the sequential float32 sum is done on the host."""
        c = L.as_u8(codes)
        if c.size != self.numSubvectors:
            raise QuantizerError("codes length mismatch")
        t = self.BuildDistanceTable(query).reshape(self.numSubvectors, self.numCentroids)
        dist = F(0)
        for m in range(self.numSubvectors):
            dist = F(dist + t[m, int(c[m])])
        return dist

    def BytesPerVector(self) -> int:
        return self.numSubvectors

    def BytesPerDimension(self) -> int:
        return 0

    def NumSubvectors(self):
        return self.numSubvectors

    def NumCentroids(self):
        return self.numCentroids

    def IsTrained(self):
        return self.trained

    def CompressionRatio(self) -> float:
        return self.dimension * 4 / self.numSubvectors


class OptimizedProductQuantizer:
    """quantization.OptimizedProductQuantizer (opq.go): block-diagonal rotation + ProductQuantizer."""

    def __init__(self, dimension: int, numSubvectors: int, numCentroids: int, numIterations: int):
        self.pq = ProductQuantizer(dimension, numSubvectors, numCentroids)
        bs = C.c_int64()
        L.call("vg_opq_block_size", dimension, numSubvectors, C.byref(bs))
        self.blockSize = int(bs.value)
        self.numIterations = numIterations
        nb = dimension // self.blockSize
        self.rotations = np.tile(np.eye(self.blockSize, dtype=F), (nb, 1, 1))
        self.trained = False

    def Train(self, vectors, pq_iters: int = 20, seed: int = 0):
        """opq.go:89-193: numIterations rounds of rotate -> pq.Train -> Procrustes."""
        v = L.as_f32(vectors)
        if v.size == 0:
            raise QuantizerError("no vectors provided for training")
        if v.ndim != 2 or v.shape[1] != self.pq.dimension:
            raise QuantizerError("vector dimension mismatch")
        rot = np.ascontiguousarray(self.rotations, F)
        p = self.pq
        L.call("vg_opq_train", L.ptr(v, L.f32p), v.shape[0], p.dimension, p.numSubvectors, p.numCentroids, self.numIterations, pq_iters,
               seed, L.ptr(rot, L.f32p), L.ptr(p.codebooks, L.i8p), L.ptr(p.scales, L.f32p), L.ptr(p.offsets, L.f32p))
        self.rotations = rot
        p.trained = self.numIterations > 0
        self.trained = True

    def _check(self):
        if not self.trained:
            raise QuantizerError("OptimizedProductQuantizer not trained")

    def RotateBatch(self, vectors, inverse: bool = False) -> np.ndarray:
        v = L.as_f32(vectors).reshape(-1, self.pq.dimension)
        out = np.zeros_like(v)
        rot = np.ascontiguousarray(self.rotations, F)
        L.call("vg_opq_rotate", L.ptr(v, L.f32p), v.shape[0], self.pq.dimension, self.blockSize, L.ptr(rot, L.f32p), int(inverse),
               L.ptr(out, L.f32p))
        return out

    def EncodeBatch(self, vectors) -> np.ndarray:
        self._check()
        return self.pq.EncodeBatch(self.RotateBatch(vectors))

    def Encode(self, vec):
        return self.EncodeBatch(np.asarray(vec, F)[None, :])[0]

    def DecodeBatch(self, codes) -> np.ndarray:
        self._check()
        return self.RotateBatch(self.pq.DecodeBatch(codes), inverse=True)

    def Decode(self, codes):
        return self.DecodeBatch(np.asarray(codes, np.uint8)[None, :])[0]

    def ComputeAsymmetricDistance(self, query, codes) -> np.float32:
        """opq.go:265-282: rotate the query once, then the PQ asymmetric distance."""
        self._check()
        rq = self.RotateBatch(np.asarray(query, F)[None, :])[0]
        return self.pq.ComputeAsymmetricDistance(rq, codes)

    def BytesPerVector(self) -> int:
        return self.pq.BytesPerVector()

    def CompressionRatio(self) -> float:
        return self.pq.CompressionRatio()

    def IsTrained(self):
        return self.trained
