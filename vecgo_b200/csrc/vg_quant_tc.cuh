// vg_quant_tc.cuh — decode-GEMM candidate filter for the quantized scans (SQ8, INT4, PQ/OPQ) on the tcgen05 tensor
// cores (vg_quant_tc.cu): codes are decoded to fp16 inside the kernel, straight into the swizzled shared-memory
// B tile, contracted against the fp16 query tile, and the survivors are re-scored exactly in the reference's order.
#pragma once
#include <vector>

#include "vg_common.cuh"
#include "vg_scan.cuh"

namespace vg {
namespace qtc {

// Per-index state of the filter (built once per index, after its codes were uploaded).
struct Prepared {
    DevBuf perm;      // int32 [dimp]: dimension held at position p along K of the B tile (-1 = padding)
    DevBuf wq;        // float [dimp]: decode weight w of that dimension (x^_d = mid_d + w_d b_d, b_d the integer the producer emits)
    DevBuf midp;      // float [dimp]: mid of that dimension
    DevBuf perm8, wq8, midp8;   // the same three in the K order of the kind::i8 decode (INT4, dim % 128 == 0); empty otherwise
    DevBuf xn;        // float [rows]: ||decode(row)||^2
    DevBuf xmax;      // uint  [4]: max ||decode(row)||^2, max ||decode(row) - mid||^2 (float bits)
    int dimp = 0;     // dim rounded up to a multiple of 64
    float mid_norm = 0.0f;  // ||mid||
    bool ready = false;
};

// Shapes the filter handles; everything else stays on the CUDA-core scan.
bool supported(const CodecParams &cp, int metric, int64_t rows, int64_t nq, int64_t k, int64_t num_partitions);

// Builds `pp` for the index described by `cp` (reads the codes: call after the upload).  h_* are host copies of the
// codec parameters: SQ8 (mins, invScales), INT4 (min, diff), PQ (scales, offsets).
vg_status prepare(const CodecParams &cp, int64_t rows, const float *h_p0, const float *h_p1, Prepared &pp, cudaStream_t st);

struct SearchIO {
    const float *d_queries = nullptr;  // [nq][dim] float32 (OPQ: already rotated)
    int64_t q_stride = 0, nq = 0;
    int64_t q_index0 = 0;              // index of the first query inside the prepared per-query arrays of `cp` (RaBitQ sign words / norms)
    int64_t rows = 0;
    const uint8_t *d_mask = nullptr;
    int k = 0;
    uint32_t row_base = 0;
    uint32_t *d_rows = nullptr;   // [nq][k]
    float *d_scores = nullptr;    // [nq][k]
    int32_t *d_counts = nullptr;  // [nq]
};
// Filter + exact stage + certificate for one batch.  `failed` lists the queries whose certificate did not hold: the
// caller re-runs those on the exact scan (scan_topk_subset).
vg_status search(const CodecParams &cp, const Prepared &pp, const SearchIO &io, std::vector<int32_t> &failed, cudaStream_t st);
// The same without waiting for the device: launches everything and leaves d_fail[q] = 1 where the certificate did not
// hold.  kc_scale = 1 is the normal pass; kc_scale = 2 is the second chance for queries that failed it (twice the
// candidate groups), available when second_chance_possible().  count_fallbacks() accounts queries that end on the exact scan.
vg_status enqueue(const CodecParams &cp, const Prepared &pp, const SearchIO &io, int kc_scale, int32_t *d_fail, cudaStream_t st);
bool second_chance_possible(const CodecParams &cp, int64_t rows, int64_t k);
// The stronger second chance: a THRESHOLD pass.  d_kth[q] = the k-th best exact score already known for query q (from the
// first pass; +inf if it found fewer than k rows).  Every row whose filter score could still beat it is listed by a
// threshold-collect GEMM epilogue and scored exactly, so the result is the exact scan's by construction — also on tightly
// clustered data where thousands of rows sit within the error bound of the k-th best and no certificate can hold.
// d_fail[q] = 1 only when a candidate list overflowed (then the exact CUDA-core scan serves the query).
vg_status enqueue_threshold(const CodecParams &cp, const Prepared &pp, const SearchIO &io, const float *d_kth, int32_t *d_fail,
                            cudaStream_t st);
bool threshold_pass_possible(const CodecParams &cp, int64_t rows, int64_t nq);
vg_status gather_kth(const float *d_scores, const int32_t *d_counts, const int32_t *d_idx, int64_t n, int k, float *d_kth, cudaStream_t st);
void count_fallbacks(uint64_t n);

// Quantized distance of query q to its r candidate rows d_rows[q][0..r) (local row ids; rows >= `rows` give NaN), in the
// reference's arithmetic.  RaBitQ needs cp.q_words / cp.q_norms (prep_sign_queries); OPQ queries must be rotated already.
// INT4: int4_lut = 1 scores like Int4Quantizer.L2Distance on a trained quantizer (simd.Int4L2DistancePrecomputed over the
// BuildInt4LookupTable values, int4.go:141-144), 0 like simd.Int4L2Distance (FMA dequantisation).
vg_status score_rows(const CodecParams &cp, int64_t rows, const float *d_queries, int64_t q_stride, int64_t nq, const uint32_t *d_rows,
                     int64_t r, float *d_out, int int4_lut, cudaStream_t st);

void stats(uint64_t *queries, uint64_t *fallbacks);
// Returns the accumulated CUDA-event time / launch count of the GEMM kernel since the last reset; enable = 1 / 0 turns
// the event pair around every GEMM launch on / off and resets the counters, enable < 0 only reads.
void profile(int enable, double *gemm_ms, uint64_t *gemm_launches);
bool i8_state();
void set_i8(bool on);   // SQ8 filter through tcgen05 kind::i8 (default on; VECGO_QTC_I8=0)

}  // namespace qtc
}  // namespace vg
