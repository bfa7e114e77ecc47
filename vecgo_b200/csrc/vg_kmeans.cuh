// vg_kmeans.cuh — device entry points of vg_kmeans.cu.
#pragma once
#include "vg_common.cuh"

namespace vg {
// kmeans.FindClosestCentroids: out[q][0..np) = partition ids by (distance asc | dot desc, id asc).
vg_status dev_find_closest(const float *d_queries, int64_t nq, int64_t dim, const float *d_centroids, int64_t k, int64_t np,
                           int metric, int32_t *d_out, cudaStream_t st);
// ProductQuantizer.Train on device-resident vectors; cent [m][k][ds] f32, cb int8, sc / of [m] stay on the device.
vg_status dev_pq_train(const float *d_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t iters, uint64_t seed, DevBuf &cent,
                       DevBuf &cb, DevBuf &sc, DevBuf &of, cudaStream_t st);
// The same for subspaces [g0, g1) only (outputs hold g1 - g0 subspaces): what one GPU trains when the subspaces of a
// quantizer are split across GPUs.  Bit-identical to the corresponding slice of dev_pq_train.
vg_status dev_pq_train_range(const float *d_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t iters, uint64_t seed, int64_t g0,
                             int64_t g1, DevBuf &cent, DevBuf &cb, DevBuf &sc, DevBuf &of, cudaStream_t st);
}  // namespace vg
