"""Mirror of internal/kmeans/kmeans.go over the CUDA library."""
from __future__ import annotations

import numpy as np

from . import _lib as L


def TrainKMeans(vectors, dim: int, k: int, metric: int, maxIter: int, init_rows=None, seed: int = 0, return_assign=False):
    """kmeans.TrainKMeans (kmeans.go:16-138).  `init_rows` plays Go's rand.Perm(n)[:k]
    (explicit because the reference draws it from the unseeded global RNG)."""
    v = L.as_f32(vectors).reshape(-1, dim)
    n = v.shape[0]
    if n < k:
        return None  # Go: return nil, nil
    if init_rows is None:
        init_rows = np.random.default_rng(seed).permutation(n)[:k]
    init = np.ascontiguousarray(init_rows, np.int64)
    cent = np.zeros((k, dim), np.float32)
    assign = np.zeros(n, np.int32)
    iters = np.zeros(1, np.int64)
    L.call("vg_kmeans_train", L.ptr(v, L.f32p), n, dim, k, int(metric), maxIter, L.ptr(init, L.i64p), seed,
           L.ptr(cent, L.f32p), L.ptr(assign, L.i32p), L.ptr(iters, L.i64p))
    if return_assign:
        return cent, assign, int(iters[0])
    return cent


def AssignPartition(vec, centroids, dim: int, metric: int):
    """kmeans.AssignPartition (kmeans.go:142-196); accepts one vector or a batch."""
    v = L.as_f32(vec).reshape(-1, dim)
    c = L.as_f32(centroids).reshape(-1, dim)
    out = np.zeros(v.shape[0], np.int32)
    L.call("vg_kmeans_assign", L.ptr(v, L.f32p), v.shape[0], dim, L.ptr(c, L.f32p), c.shape[0], int(metric), L.ptr(out, L.i32p))
    return int(out[0]) if np.ndim(vec) == 1 else out


def FindClosestCentroids(query, centroids, dim: int, n: int, metric: int):
    """kmeans.FindClosestCentroids (kmeans.go:217-280); accepts one query or a batch."""
    q = L.as_f32(query).reshape(-1, dim)
    c = L.as_f32(centroids).reshape(-1, dim)
    n = min(n, c.shape[0])
    out = np.zeros((q.shape[0], n), np.int32)
    L.call("vg_kmeans_find_closest", L.ptr(q, L.f32p), q.shape[0], dim, L.ptr(c, L.f32p), c.shape[0], n, int(metric),
           L.ptr(out, L.i32p))
    return out[0] if np.ndim(query) == 1 else out
