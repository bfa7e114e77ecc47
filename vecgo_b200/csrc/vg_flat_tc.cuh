// vg_flat_tc.cuh — tensor-core (tcgen05, TF32) candidate filter for Flat search (vg_flat_tc.cu).
#pragma once
#include <vector>

#include "vg_common.cuh"

namespace vg {
namespace tc {

struct FilterArgs {
    const float *d_queries = nullptr;  // [nq] rows of dim floats, q_stride floats apart
    int64_t q_stride = 0;              // 0 = dim
    const float *d_vectors = nullptr;  // [rows][dim] row-major float32 (the segment's vector section)
    const float *d_xn = nullptr;       // [rows] squared norms (L2) — unused for dot
    const uint8_t *d_mask = nullptr;   // optional row bitmap (bit = 1 keeps the row), 4-byte aligned
    int64_t nq = 0, rows = 0, dim = 0;
    // optional fp16 shadow of the vectors, [rows][dimp] halves = x * 2^x16_exp (dimp = dim rounded up to 64): when set
    // (and VECGO_FLAT_PAIR != 0) the filter runs the CTA-pair kind::f16 kernel instead of the single-CTA TF32 one
    const void *d_x16 = nullptr;
    int x16_exp = 0;
    int kc = 0;                        // number of row groups kept per query (>= k): at least kc rows have s <= tau
    int is_dot = 0;
    uint32_t row_base = 0;
    // outputs of filter()
    float *d_tau = nullptr;            // [nq] kc-th smallest group minimum in s-space (s = ||x||^2 - 2q.x | -q.x)
    uint32_t *d_gids = nullptr;        // [nq][kc] per selected group: its arg-min row, or 0x80000000|group when crowded (0xFFFFFFFF padded)
    int32_t *d_gcnt = nullptr;         // [nq] number of listed groups (< kc only when the segment has fewer groups)
};

bool supported(int64_t dim, int64_t rows, int64_t nq, int64_t k);
int candidates_for(int64_t k, int64_t dim);
// nearest-centroid assignment (k-means): k = 1 against a small centroid table
bool supported_assign(int64_t dim, int64_t rows, int64_t nq, int64_t q_stride);
int candidates_for_assign(int64_t rows);
// out[i] = ||v_i||^2; optional running maximum (as uint bits of a non-negative float).
vg_status sqnorms(const float *d_v, int64_t n, int64_t dim, int64_t stride, float *d_out, unsigned int *d_max_bits, cudaStream_t st);
int64_t group_rows(int64_t rows, int kc);  // rows per minimum group for a segment of `rows` rows
bool pair_enabled();                       // VECGO_FLAT_PAIR != 0
bool uses_pair(const FilterArgs &f);       // the CTA-pair fp16 kernel will run this filter (groups are then <= 128 rows)
int64_t filter_group_rows(const FilterArgs &f);
// x16[r][p] = half(x[r][p] * 2^sx_exp), zero padded to dimp columns.
vg_status make_shadow16(const float *d_x, int64_t rows, int64_t dim, int dimp, int sx_exp, void *d_x16, cudaStream_t st);
// tau and the kc best groups per query from the [nq][groups] (m1, m2) pairs a GEMM epilogue wrote (shared with vg_quant_tc.cu).
vg_status select_groups(const float2 *d_mins, int64_t groups, int64_t nq, int kc, int64_t G, float *d_tau, uint32_t *d_cand,
                        int32_t *d_gcnt, cudaStream_t st);
// CUtensorMap (`map` points at one) over a row-major [rows][cols] matrix of float32 or float16, 128-byte swizzled boxes.
vg_status tensor_map_2d(void *map, bool f16, const void *base, int64_t rows, int64_t cols, int64_t stride_elems, int box_cols, int box_rows);
vg_status tensor_map_2d_u8(void *map, const void *base, int64_t rows, int64_t cols, int64_t stride_bytes, int box_cols, int box_rows);
// TF32 GEMM with group-minimum epilogue → tau and the kc best groups per query.
vg_status filter(const FilterArgs &f, cudaStream_t st);
// Exact scores of the candidate rows in simd pair order, top-k by (score,row), certificate → d_fail[q] (1 = re-run exactly).
vg_status finalize(const FilterArgs &f, int k, const float *d_qn, const unsigned int *d_xmax_bits, uint32_t *d_rows, float *d_scores,
                   int32_t *d_counts, int32_t *d_fail, cudaStream_t st);

// One batch end to end (norms of the queries, filter, exact stage, certificate).
struct SearchIO {
    const float *d_queries = nullptr;
    int64_t q_stride = 0, nq = 0;
    const float *d_vectors = nullptr;
    int64_t rows = 0, dim = 0;
    const float *d_xn = nullptr;               // [rows] squared norms of the vectors
    const unsigned int *d_xmax_bits = nullptr; // their maximum (float bits)
    const void *d_x16 = nullptr;               // optional fp16 shadow (see FilterArgs)
    int x16_exp = 0;
    const uint8_t *d_mask = nullptr;
    int k = 0, is_dot = 0;
    uint32_t row_base = 0;
    uint32_t *d_rows = nullptr;                // [nq][k]
    float *d_scores = nullptr;                 // [nq][k]
    int32_t *d_counts = nullptr;               // [nq]
};
vg_status search(const SearchIO &io, int kc, std::vector<int32_t> &failed, cudaStream_t st);
// The same in two halves for callers that must not wait for the device: enqueue() launches the filter, the exact stage
// and the certificate (d_fail[q] = 1 where it did not hold) and returns; retry() takes the failed queries (read back by
// the caller whenever it synchronises anyway), gives them the second chance and leaves the still-unproven ones in `failed`.
vg_status enqueue(const SearchIO &io, int kc, int32_t *d_fail, cudaStream_t st);
vg_status retry(const SearchIO &io, int kc, std::vector<int32_t> &failed, cudaStream_t st);

// Single-launch search for short vectors (vg_flat_single.cu): phase 0 query preparation, sampled thresholds, threshold
// filter, exact stage and certificate in ONE cooperative kernel with grid-wide barriers — no intermediate plane, no
// host round trip.  dim <= 256, k <= 16, batches whose query tiles fit half the SM pairs.  VECGO_FLAT_SINGLE=0 disables.
namespace fs {
bool single_supported(int64_t dim, int64_t rows, int64_t nq, int64_t k);
vg_status single_enqueue(const SearchIO &io, int32_t *d_fail, cudaStream_t st);
}  // namespace fs
void count_queries(uint64_t n);

// Process-wide switch (default on; environment VECGO_FLAT_TC=0 turns it off) and counters.
bool enabled();
void set_enabled(bool on);
void stats(uint64_t *queries, uint64_t *fallbacks);

}  // namespace tc
}  // namespace vg
