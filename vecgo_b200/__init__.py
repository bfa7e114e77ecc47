"""vecgo_b200 — B200-native (sm_100a) replacement for vecgo's vector-scan hot path.

The product is vecgo_b200/libvecgo_cuda.so (C ABI: include/vecgo_cuda.h).  The
modules here mirror the Go packages it plugs into:

    simd          internal/simd kernel table
    distance      distance package
    quantization  internal/quantization (SQ8, INT4, BQ, RaBitQ, PQ)
    kmeans        internal/kmeans
    index / flat  internal/segment/flat (Segment.Search / Rerank) + top-k merge
    sharded       row sharding across GPUs + NCCL all-gather merge
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from . import distance, flat, index, kmeans, quantization, sharded, simd  # noqa: F401
from ._lib import VecgoError, launch_count  # noqa: F401

__all__ = ["simd", "distance", "quantization", "kmeans", "index", "flat", "VecgoError", "launch_count"]
