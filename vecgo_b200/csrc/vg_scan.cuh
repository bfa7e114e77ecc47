// vg_scan.cuh — host-visible descriptors of the scan kernels (vg_scan.cu).
#pragma once
#include <vector>

#include "vg_common.cuh"
#include "vg_topk.cuh"

namespace vg {

// What a scan kernel reads.  Only the members of the active codec are set.
struct CodecParams {
    int codec = 0;     // VG_CODEC_*
    int variant = 0;   // VG_VAR_*
    int64_t dim = 0;
    int64_t row_bytes = 0;      // device row stride of `codes` in bytes
    const uint8_t *codes = nullptr;
    const float *vectors = nullptr;  // F32 rows [rows][dim]
    // SQ8 / INT4 per-dimension parameters (device)
    const float *p0 = nullptr;  // SQ8 mins  | INT4 min
    const float *p1 = nullptr;  // SQ8 inv   | INT4 diff
    // PQ
    int pq_m = 0, pq_k = 0, pq_dsub = 0;
    const int8_t *pq_codebooks = nullptr;
    const float *pq_scales = nullptr, *pq_offsets = nullptr;
    const float *pq_tables = nullptr;  // optional precomputed tables [nq][m*256]
    // RaBitQ / BQ
    const float *norms = nullptr;      // RaBitQ: [rows] stored f32 norm
    int words32 = 0;                   // sign words per row (u32)
    const uint32_t *q_words = nullptr; // prepared query sign words [nq][words32]
    const float *q_norms = nullptr;    // prepared query norms [nq] (RaBitQ)
};

enum {
    VG_VAR_PAIR = 0,     // F32: simd.SquaredL2 / simd.Dot order (flat.Search, Rerank)
    VG_VAR_BATCH = 1,    // F32: simd.SquaredL2Batch / DotBatch order (k-means)
    VG_VAR_GO_SCALAR = 2,// SQ8: sequential unfused Go loop (sq.L2Distance / sq.DotProduct)
    VG_VAR_PERM = 4      // codes are stored lane-transposed on device (fast path)
};

struct ScanArgs {
    const float *queries = nullptr;  // device [nq][dim]
    int64_t q_stride = 0;            // floats between consecutive queries (0 = dim); F32 codec only
    int64_t nq = 0;
    int64_t rows = 0;
    int64_t rows_per_split = 0;
    int splits = 1;
    int k = 0, C = 0, trigger = 0;
    int descending = 0;
    int is_dot = 0;                  // score = dot product instead of squared L2
    uint32_t row_base = 0;
    const uint8_t *mask = nullptr;   // optional row bitmap
    // IVF probing (optional): per query `nprobe` partition ids; rows of partition p are [part_off[p], part_off[p+1])
    const int32_t *probe = nullptr;
    int nprobe = 0;
    const uint32_t *part_off = nullptr;
    int num_parts = 0;
    // Partition-grouped form (scan_topk_partitioned): the batch is nq x nprobe VIRTUAL queries sorted by probed partition,
    // nprobe = 1, probe[v] = that partition; a block scans only the row ranges of its slots' partitions and writes
    // the sorted keys of virtual query v to partial + emit_index[v] * k.
    int by_partition = 0;
    const int32_t *emit_index = nullptr;
    // outputs (splits == 1)
    uint32_t *out_rows = nullptr;
    float *out_scores = nullptr;
    int32_t *out_counts = nullptr;
    // outputs (splits > 1): [nq][splits][k] sorted keys
    unsigned long long *partial = nullptr;
};

// Full top-k scan of one index (chooses tile kernel, row splits and merge).
vg_status scan_topk(const CodecParams &cp, ScanArgs a, cudaStream_t st);
// IVF-partitioned segment (flat/segment.go:726-745): every query scans only the rows of its `nprobe` probed partitions.
// The (query, partition) pairs are sorted by partition so that the four virtual queries of a block share a row range
// and concurrent blocks of a partition hit the same rows in L2; each pair keeps its own top-k, merged per query at the
// end.  Work and traffic are proportional to nprobe / num_partitions of the full scan; results are the full masked
// scan's, bit for bit.
vg_status scan_topk_partitioned(const CodecParams &cp, ScanArgs a, cudaStream_t st);
// Exact re-run of a subset of the batch (`which`: query indices into a.queries / a.out_*): gathers those queries,
// scans them with scan_topk and scatters the results back into a.out_rows / a.out_scores / a.out_counts.
vg_status scan_topk_subset(const CodecParams &cp, ScanArgs a, const std::vector<int32_t> &which, cudaStream_t st);
// dst[i] = src[idx[i]] (rows of `dim` floats, `stride` floats apart) and its inverse for (rows, scores, counts) results.
vg_status dev_gather_rows(const float *d_src, int64_t stride, const int32_t *d_idx, int64_t n, int64_t dim, float *d_dst, cudaStream_t st);
vg_status dev_scatter_results(const uint32_t *d_rows, const float *d_scores, const int32_t *d_counts, const int32_t *d_idx, int64_t n,
                              int64_t k, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts, cudaStream_t st);
// Dense distance matrix out[nq][n] (simd kernel-table mirrors, rerank, k-means); no top-k.
vg_status scan_dense(const CodecParams &cp, const float *d_queries, int64_t nq, int64_t n, int is_dot, float *d_out,
                     cudaStream_t st);
// Exact per-candidate scores: out[q][j] = dist(query q, vectors[rows[q][j]]) in simd pair order.
vg_status rerank_gather(const float *d_vectors, int64_t nrows, int64_t dim, const float *d_queries, int64_t nq,
                        const uint32_t *d_rows, int64_t r, int is_dot, float *d_out, cudaStream_t st);
// the same against rows in mapped host memory: staged through HBM chunk by chunk (see vg_scan.cu)
vg_status rerank_gather_host(const float *d_host_alias, int64_t nrows, int64_t dim, const float *d_queries, int64_t nq,
                             const uint32_t *d_rows, int64_t r, int is_dot, float *d_out, cudaStream_t st);
// simd.SquaredL2Bounded per (query, candidate row): out = the distance, or the partial sum at the 64-dim block where it
// first exceeded the bound (exceeded = 1).  Bounds: one per query, or one per pair when per_pair_bounds.
vg_status bounded_l2_gather(const float *d_vectors, int64_t nrows, int64_t dim, const float *d_queries, int64_t nq, const uint32_t *d_rows,
                            int64_t r, const float *d_bounds, int per_pair_bounds, float *d_out, uint8_t *d_exceeded, cudaStream_t st);
// Hamming matrix out[nq][n] between byte strings.
vg_status hamming_dense(const uint8_t *d_q, int64_t nq, const uint8_t *d_codes, int64_t n, int64_t nbytes, int32_t *d_out,
                        cudaStream_t st);
// Query preparation for BQ / RaBitQ scans: sign words (+ query norm in simd.Dot order).
vg_status prep_sign_queries(const float *d_queries, int64_t nq, int64_t dim, float threshold, uint32_t *d_words,
                            float *d_norms, cudaStream_t st);

}  // namespace vg
