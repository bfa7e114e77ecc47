// vg_pq_assign_tc.cuh — tensor-core assignment step of PQ training (8-dim subspaces, 256 centroids), exact through a
// gap certificate + exact re-evaluation of the uncertain pairs (vg_pq_assign_tc.cu).
#pragma once
#include "vg_common.cuh"

namespace vg {
namespace pqa {

struct Assigner {
    DevBuf x16;      // [n][G * 32] fp16: hi / lo split of the scaled samples + the constant slots (built once per training)
    DevBuf c16;      // [G / 2][256][64] fp16: centroid tiles of the current iteration
    DevBuf maxbits;  // [G] max ||x_m||^2, [G] max ||c_m||^2 (float bits), then the fallback counter
    DevBuf list;     // uncertified (sample, subspace) pairs of the current pass
    const float *vecs = nullptr;
    int64_t n = 0, dim = 0;
    int G = 0;
    float scale = 1.0f;
    bool ready = false;

    static bool supported(int64_t n, int64_t dim, int G, int K, int ds);
    // Builds the sample shadow; leaves `ready` false (exact CUDA-core path) for constant, non-finite or extreme data.
    vg_status prepare(const float *d_vecs, int64_t n, int64_t dim, int G, cudaStream_t st);
    // assign[g][i] = nearest centroid of sample i in subspace g, identical to the reference loop.
    vg_status assign(const float *d_cent, uint32_t *d_assign, cudaStream_t st);
    vg_status last_fallbacks(uint64_t *pairs, cudaStream_t st);
    vg_status account(cudaStream_t st);   // adds the last pass to the library-wide counters (reads 4 bytes back)
};
// (sample, subspace) pairs assigned on the tensor cores / re-evaluated exactly since the library was loaded
void stats(uint64_t *pairs, uint64_t *fallback_pairs);

}  // namespace pqa
}  // namespace vg
