// vg_quant.cuh — device entry points of vg_quant.cu (all pointers are device memory).
#pragma once
#include "vg_common.cuh"

namespace vg {
vg_status dev_minmax(const float *d_vecs, int64_t n, int64_t dim, float *d_mins, float *d_maxs, cudaStream_t st);
vg_status dev_minmax_strided(const float *d_vecs, int64_t n, int64_t dim, int64_t stride, float *d_mins, float *d_maxs, cudaStream_t st);
vg_status dev_sq8_encode(const float *d_vecs, int64_t n, int64_t dim, const float *d_mins, const float *d_maxs,
                         const float *d_scales, uint8_t *d_codes, cudaStream_t st);
vg_status dev_sq8_decode(const uint8_t *d_codes, int64_t n, int64_t dim, const float *d_mins, const float *d_inv, float *d_vecs,
                         cudaStream_t st);
vg_status dev_int4_encode(const float *d_vecs, int64_t n, int64_t dim, const float *d_min, const float *d_diff, uint8_t *d_codes,
                          cudaStream_t st);
vg_status dev_int4_decode(const uint8_t *d_codes, int64_t n, int64_t dim, const float *d_min, const float *d_diff, float *d_vecs,
                          cudaStream_t st);
vg_status dev_mean_f64(const float *d_vecs, int64_t total, double *h_sum, cudaStream_t st);
vg_status dev_sign_encode(const float *d_vecs, int64_t n, int64_t dim, float threshold, bool with_norm, uint8_t *d_codes,
                          cudaStream_t st);
vg_status dev_pq_encode(const float *d_vecs, int64_t n, int64_t dim, int m, int k, const int8_t *d_cb, const float *d_scales,
                        const float *d_offsets, uint8_t *d_codes, cudaStream_t st);
vg_status dev_pq_decode(const uint8_t *d_codes, int64_t n, int64_t dim, int m, int k, const int8_t *d_cb, const float *d_scales,
                        const float *d_offsets, float *d_vecs, cudaStream_t st);
vg_status dev_pq_tables(const float *d_queries, int64_t nq, int64_t dim, int m, int k, const int8_t *d_cb, const float *d_scales,
                        const float *d_offsets, float *d_tables, cudaStream_t st);
vg_status dev_scale(float *d_a, int64_t n, float s, cudaStream_t st);
vg_status dev_normalize(float *d_v, int64_t n, int64_t dim, uint8_t *d_ok, cudaStream_t st);
vg_status dev_permute_sq8(const uint8_t *d_src, uint8_t *d_dst, int64_t n, int64_t dim, int vb, cudaStream_t st);
vg_status dev_permute_int4(const uint8_t *d_src, uint8_t *d_dst, int64_t n, int64_t cs, cudaStream_t st);
vg_status dev_permute_pq(const uint8_t *d_src, uint8_t *d_dst_base, int64_t row0, int64_t n, int m, int mpad, cudaStream_t st);
vg_status dev_split_sign(const uint8_t *d_src, int64_t n, int64_t nbytes, int64_t src_stride, int64_t dst_stride, uint8_t *d_bits,
                         float *d_norms, cudaStream_t st);
// OptimizedProductQuantizer (vg_opq.cu): rotateVector for n vectors; Procrustes rotations of `blocks` bs x bs matrices.
vg_status dev_opq_rotate(const float *d_v, int64_t n, int64_t dim, int bs, const float *d_rot, float *d_out, cudaStream_t st);
vg_status dev_opq_procrustes(const float *d_M, int blocks, int bs, float *d_R, float *d_sigma, cudaStream_t st);
}  // namespace vg
