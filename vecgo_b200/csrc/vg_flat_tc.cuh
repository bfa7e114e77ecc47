// vg_flat_tc.cuh — tensor-core (tcgen05, TF32) candidate filter for Flat search (vg_flat_tc.cu).
#pragma once
#include "vg_common.cuh"

namespace vg {
namespace tc {

constexpr int CAP = 128;  // candidate keys kept per query

struct FilterArgs {
    const float *d_queries = nullptr;  // [nq][dim]
    const float *d_vectors = nullptr;  // [rows][dim] row-major float32 (the segment's vector section)
    const float *d_xn = nullptr;       // [rows] squared norms (L2) — unused for dot
    const uint8_t *d_mask = nullptr;   // optional row bitmap (bit = 1 keeps the row), 4-byte aligned
    int64_t nq = 0, rows = 0, dim = 0;
    int kc = 0;                        // rank of the group minimum used as threshold: at least kc rows survive
    int is_dot = 0;
    uint32_t row_base = 0;
    // outputs
    float *d_tau = nullptr;                  // [ceil(nq/256)*256] threshold per query in s-space (s = ||x||^2 - 2q.x | -q.x)
    unsigned long long *d_cand = nullptr;    // [nq][CAP] keys (orderable(s_approx) << 32 | global row), unsorted
    int32_t *d_cand_cnt = nullptr;           // [nq] rows with s_approx <= tau (more than CAP = overflow)
};

bool supported(int64_t dim, int64_t rows, int64_t nq, int64_t k);
int candidates_for(int64_t k, int64_t dim);
// out[i] = ||v_i||^2; optional running maximum (as uint bits of a non-negative float).
vg_status sqnorms(const float *d_v, int64_t n, int64_t dim, float *d_out, unsigned int *d_max_bits, cudaStream_t st);
// pass 1 (group minima) → tau → pass 2 (collect rows with s <= tau).
vg_status filter(const FilterArgs &f, cudaStream_t st);
// Exact scores of the candidates in simd pair order, final top-k by (score,row), certificate → d_fail[q] (1 = re-run exactly).
vg_status finalize(const FilterArgs &f, int k, const float *d_qn, const unsigned int *d_xmax_bits, uint32_t *d_rows, float *d_scores,
                   int32_t *d_counts, int32_t *d_fail, cudaStream_t st);

}  // namespace tc
}  // namespace vg
