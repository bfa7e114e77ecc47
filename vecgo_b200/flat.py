"""Mirror of internal/segment/flat: the immutable flat segment (format.go,
segment.go) — open from file bytes onto the GPU, batched Search / Rerank —
plus a writer that produces byte-compatible files (writer.go:335-516) with
the quantizer sections encoded by the CUDA quantizers."""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from . import _lib as L
from . import kmeans as km
from .index import DeviceIndex
from .quantization import ProductQuantizer, ScalarQuantizer

MagicNumber = 0x56454331  # "VEC1"
Version = 1
HeaderSize = 152
QuantizationNone, QuantizationSQ8, QuantizationPQ = 0, 1, 2  # format.go:22-26


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli) of the body — file-format plumbing for the writer."""
    tbl = crc32c._tbl
    if tbl is None:
        tbl = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tbl.append(c)
        crc32c._tbl = tbl
    crc = 0xFFFFFFFF
    for b in data:
        crc = tbl[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


crc32c._tbl = None


def write_segment(*, segment_id: int, vectors, ids=None, metric: int = 0, quantization: int = QuantizationNone, pq_m: int = 0,
                  k_partitions: int = 0, kmeans_iters: int = 10, seed: int = 0, checksum: bool = True, pq_iters: int = 20) -> bytes:
    """flat.Writer.Flush (writer.go:99-516): optional k-means partitioning (rows reordered by
    partition), SQ8 / PQ train + encode on the GPU, packed little-endian sections."""
    v = L.as_f32(vectors)
    n, dim = v.shape
    ids = np.arange(n, dtype=np.uint64) if ids is None else np.ascontiguousarray(ids, np.uint64)
    cent = np.zeros((0, dim), np.float32)
    poff = np.zeros(0, np.uint32)
    if k_partitions > 1 and n >= k_partitions:
        init = np.random.default_rng(seed).permutation(n)[:k_partitions]
        cent = km.TrainKMeans(v, dim, k_partitions, metric, kmeans_iters, init_rows=init, seed=seed)
        assign = km.AssignPartition(v, cent, dim, metric)
        order = np.argsort(assign, kind="stable")
        v, ids = v[order], ids[order]
        counts = np.bincount(assign, minlength=k_partitions)
        poff = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    quant_meta = b""
    codes = b""
    if quantization == QuantizationSQ8:
        sq = ScalarQuantizer(dim)
        sq.Train(v)
        codes = sq.EncodeBatch(v).tobytes()
        quant_meta = sq.mins.astype("<f4").tobytes() + sq.maxs.astype("<f4").tobytes()
    elif quantization == QuantizationPQ:
        pq = ProductQuantizer(dim, pq_m, 256)
        pq.Train(v, iters=pq_iters, seed=seed)
        codes = pq.EncodeBatch(v).tobytes()
        quant_meta = (struct.pack("<II", pq_m, 256) + pq.scales.astype("<f4").tobytes() + pq.offsets.astype("<f4").tobytes()
                      + pq.codebooks.tobytes())
    nparts = cent.shape[0]
    body = bytearray()
    off = {}
    for name, blob in (("centroid", cent.astype("<f4").tobytes()), ("part", poff.astype("<u4").tobytes()), ("quant", quant_meta),
                       ("codes", codes), ("vec", v.astype("<f4").tobytes()), ("pk", ids.astype("<u8").tobytes())):
        off[name] = HeaderSize + len(body)
        body += blob
    off["meta"] = HeaderSize + len(body)
    body += np.zeros(n + 1, "<u4").tobytes()  # metadata offsets, empty blob
    off["stats"] = 0
    hdr = bytearray(HeaderSize)
    struct.pack_into("<IIQII", hdr, 0, MagicNumber, Version, segment_id, n, dim)
    hdr[24] = int(metric)
    struct.pack_into("<I", hdr, 28, nparts)
    hdr[32] = quantization
    struct.pack_into("<8Q", hdr, 40, off["centroid"], off["part"], off["quant"], off["codes"], off["vec"], off["pk"], off["meta"],
                     off["stats"])
    struct.pack_into("<I", hdr, 104, crc32c(bytes(body)) if checksum else 0)
    return bytes(hdr) + bytes(body)


def decode_header(data: bytes) -> dict:
    buf = np.frombuffer(data, np.uint8)
    h = L.FlatHeader()
    L.call("vg_flat_decode_header", L.ptr(buf, L.u8p), len(data), C.byref(h))
    return {f: getattr(h, f) for f, _ in L.FlatHeader._fields_}


class Segment:
    """flat.Segment (segment.go): Open / Search / Rerank / FetchIDs / Close."""

    def __init__(self, index: DeviceIndex, header: dict):
        self.index, self.header = index, header

    @classmethod
    def Open(cls, data: bytes, verify_checksum: bool = True) -> "Segment":
        buf = np.frombuffer(data, np.uint8)
        hdr = decode_header(data)
        h = C.c_uint64()
        L.call("vg_flat_open", L.ptr(buf, L.u8p), len(data), int(verify_checksum), C.byref(h))
        codec = {0: L.CODEC_F32, 1: L.CODEC_SQ8, 2: L.CODEC_PQ}[hdr["quantization_type"]]
        ix = DeviceIndex(codec=codec, metric=hdr["metric"], dim=hdr["dim"], rows=hdr["row_count"],
                         segment_id=hdr["segment_id"], _handle=h.value)
        return cls(ix, hdr)

    def ID(self):
        return self.header["segment_id"]

    def RowCount(self):
        return self.header["row_count"]

    def Metric(self):
        return self.header["metric"]

    def Search(self, queries, k: int, nprobes: int = 0, row_mask=None):
        return self.index.search(queries, k, nprobes=nprobes, row_mask=row_mask)

    def Rerank(self, queries, rows):
        return self.index.rerank(queries, rows)

    def FetchIDs(self, rows):
        return self.index.fetch_ids(rows)

    def Close(self):
        self.index.close()
