/*
 * vecgo_cuda.h — C ABI of libvecgo_cuda.so: the B200 (sm_100a) replacement for
 * vecgo's vector-scan hot path.
 *
 * This is the boundary a cgo shim binds (INTEGRATION.md shows the Go side).
 * Plain pointers and sizes only; no C++/torch types.  Every entry point
 * returns VG_OK (0) or a negative vg_status; vg_last_error() returns the
 * thread-local message of the last failure on the calling thread.  Nothing
 * here ever falls back to the CPU: without a CUDA device every compute call
 * fails with VG_ERR_CUDA.
 *
 * Pointer naming: `h_` = host memory owned by the caller for the duration of
 * the call only (cgo rule: never retained); `d_` = device memory (e.g. a
 * torch tensor's data_ptr, or memory from vg_dev_alloc).  Handles are opaque.
 *
 * Reference interfaces replaced (paths relative to hupe1980/vecgo):
 *   internal/simd/kernels.go:39-123        kernel table  -> vg_simd_*
 *   distance/distance.go:13-53             Dot/SquaredL2/NormalizeL2InPlace
 *   internal/quantization/quantizer.go:12-24 Quantizer   -> vg_sq8_*, vg_int4_*, vg_bq_*, vg_rabitq_*, vg_pq_*
 *   internal/kmeans/kmeans.go:16,142,217   k-means       -> vg_kmeans_*
 *   internal/segment/segment.go:77-95      Segment.Search/Rerank -> vg_index_search / vg_index_rerank
 *   internal/segment/flat/segment.go:105   flat.Open     -> vg_flat_open
 *   internal/searcher/candidate_queue.go   top-k order   -> result order of every search, vg_topk_merge
 */
#ifndef VECGO_CUDA_H
#define VECGO_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

typedef int32_t vg_status;
enum {
    VG_OK = 0,
    VG_ERR_INVALID = -1,   /* bad argument / dimension mismatch (Go: errors.New("dimension mismatch")) */
    VG_ERR_CUDA = -2,      /* CUDA runtime failure, no device, out of memory */
    VG_ERR_STATE = -3,     /* quantizer not trained / handle closed */
    VG_ERR_FORMAT = -4,    /* flat segment file: bad magic/version/short file/checksum */
    VG_ERR_UNSUPPORTED = -5
};

/* distance.Metric — distance/distance.go:68-73 */
enum { VG_METRIC_L2 = 0, VG_METRIC_COSINE = 1, VG_METRIC_DOT = 2, VG_METRIC_HAMMING = 3 };

/* Scan codecs.  Values 0..6 follow quantization.Type (internal/quantization/types.go:6-14). */
enum {
    VG_CODEC_F32 = 0,    /* QuantizationNone: exact float32 rows */
    VG_CODEC_PQ = 1,     /* ProductQuantizer ADC (int8 codebooks, K=256) */
    VG_CODEC_OPQ = 2,    /* rotation + PQ */
    VG_CODEC_SQ8 = 3,    /* ScalarQuantizer, L2 via Sq8uL2BatchPerDimension */
    VG_CODEC_BQ = 4,     /* BinaryQuantizer, Hamming score */
    VG_CODEC_RABITQ = 5, /* RaBitQuantizer estimator */
    VG_CODEC_INT4 = 6    /* Int4Quantizer */
};

typedef uint64_t vg_index_t; /* device-resident scan index (one segment / one shard) */

/* ---------------------------------------------------------------- runtime */
const char *vg_last_error(void);
const char *vg_version(void);
vg_status vg_device_count(int32_t *count);
/* Devices, threads and streams.
 * - A handle (vg_index_t) lives on ONE GPU: every call on it runs there, whichever thread makes it, so one process can
 *   serve several GPUs (a Go host drives all GPUs of a box from one process).
 * - vg_init(device) selects the GPU for calls WITHOUT a handle (training, encoders, simd mirrors, vg_index_create) made
 *   by the calling thread; the first device initialised is the process default for threads that never call it.
 *   vg_index_create_on / vg_flat_open_on take the device explicitly (no thread state: what a Go shim should use,
 *   goroutines migrate between OS threads).
 * - Every call runs on exactly one CUDA stream: the handle's (vg_index_set_stream), else the calling thread's
 *   (vg_set_stream), else a stream leased from the device's pool for the duration of the call.  Concurrent callers
 *   therefore run on different streams with their own scratch: search / rerank / score entry points are re-entrant on
 *   a shared handle (flat segments are immutable, concurrent reads safe: flat/doc.go:52-55; Engine.BatchSearch runs
 *   up to 100 goroutines, engine.go:365,1324).  create / upload / close are serialised per handle by the caller.
 * - Host-pointer entry points return when their results are in the caller's buffers.  `_dev` entry points are
 *   stream-ordered on a caller-provided stream (handle or thread stream) and complete on return otherwise. */
vg_status vg_init(int32_t device);
/* Waits for the library's streams on the calling thread's device (and the thread's own stream, if set). */
vg_status vg_synchronize(void);
vg_status vg_dev_alloc(void **d_ptr, size_t bytes);
vg_status vg_dev_free(void *d_ptr);
vg_status vg_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes);
vg_status vg_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes);
/* Number of kernels this library has launched since load (for gpu_launches). */
uint64_t vg_launch_count(void);
/* Calls made by the CALLING THREAD run their kernels, copies and stream-ordered allocations on this CUDA stream
 * (cudaStream_t as integer).  0 is the legacy default stream (what PyTorch's default stream reports), NOT "unset": a
 * caller that prepares inputs on stream 0 and passes 0 gets correct ordering.  VG_STREAM_LIBRARY (~0) switches the
 * thread back to library-owned streams, which is also the state of a thread that never called this. */
#define VG_STREAM_LIBRARY (~0ull)
vg_status vg_set_stream(uint64_t cuda_stream);

/* ------------------------------------------------ simd kernel-table mirrors
 * Batch forms of internal/simd/kernels.go:39-123 — host in, host out, full
 * distance vectors/matrices (no top-k).  Arithmetic = the AVX-512 kernels
 * (internal/simd/src/*_avx512.c), bit for bit.
 * out is row-major [nq x n]. */
vg_status vg_simd_dot(const float *h_a, const float *h_b, int64_t n_pairs, int64_t dim, float *h_out);          /* simd.Dot per pair */
vg_status vg_simd_squared_l2(const float *h_a, const float *h_b, int64_t n_pairs, int64_t dim, float *h_out);   /* simd.SquaredL2 per pair */
vg_status vg_simd_dot_batch(const float *h_queries, int64_t nq, const float *h_targets, int64_t n, int64_t dim, float *h_out);        /* simd.DotBatch */
vg_status vg_simd_squared_l2_batch(const float *h_queries, int64_t nq, const float *h_targets, int64_t n, int64_t dim, float *h_out); /* simd.SquaredL2Batch */
vg_status vg_simd_sq8u_l2_batch(const float *h_queries, int64_t nq, const uint8_t *h_codes, int64_t n, int64_t dim,
                                const float *h_mins, const float *h_inv_scales, float *h_out);                   /* simd.Sq8uL2BatchPerDimension */
vg_status vg_simd_int4_l2_batch(const float *h_queries, int64_t nq, const uint8_t *h_codes, int64_t n, int64_t dim,
                                const float *h_min, const float *h_diff, float *h_out);                          /* simd.Int4L2DistanceBatch */
vg_status vg_simd_pq_adc_lookup(const float *h_tables, int64_t nq, const uint8_t *h_codes, int64_t n, int64_t m, float *h_out); /* simd.PqAdcLookup; table [nq][m*256] */
vg_status vg_simd_hamming(const uint8_t *h_queries, int64_t nq, const uint8_t *h_codes, int64_t n, int64_t nbytes, int32_t *h_out); /* simd.Hamming */
vg_status vg_simd_squared_l2_bounded(const float *h_a, const float *h_b, int64_t n_pairs, int64_t dim, const float *h_bounds, float *h_out,
                                     uint8_t *h_exceeded);                                                          /* simd.SquaredL2Bounded per pair */
vg_status vg_simd_scale(float *h_a, int64_t n, float scalar);                                                    /* simd.ScaleInPlace */
/* distance.NormalizeL2InPlace on each of n rows; h_ok[i]=0 for zero-norm rows (left untouched). */
vg_status vg_normalize_l2(float *h_vecs, int64_t n, int64_t dim, uint8_t *h_ok);

/* ------------------------------------------------------------- quantizers
 * Batch Train/Encode/Decode of internal/quantization; parameter arrays are
 * caller-owned host buffers, byte-compatible with the Go structs' fields. */
/* ScalarQuantizer.Train (quantizer.go:130-180): mins,maxs,scales,inv_scales [dim] */
vg_status vg_sq8_train(const float *h_vecs, int64_t n, int64_t dim, float *h_mins, float *h_maxs, float *h_scales, float *h_inv_scales);
/* ScalarQuantizer.SetBounds (quantizer.go:51-75) */
vg_status vg_sq8_set_bounds(const float *h_mins, const float *h_maxs, int64_t dim, float *h_scales, float *h_inv_scales);
vg_status vg_sq8_encode(const float *h_vecs, int64_t n, int64_t dim, const float *h_mins, const float *h_maxs, const float *h_scales, uint8_t *h_codes);
vg_status vg_sq8_decode(const uint8_t *h_codes, int64_t n, int64_t dim, const float *h_mins, const float *h_inv_scales, float *h_vecs);
/* Int4Quantizer (int4.go:29-132): min,diff [dim]; codes [n][(dim+1)/2] */
vg_status vg_int4_train(const float *h_vecs, int64_t n, int64_t dim, float *h_min, float *h_diff);
vg_status vg_int4_encode(const float *h_vecs, int64_t n, int64_t dim, const float *h_min, const float *h_diff, uint8_t *h_codes);
vg_status vg_int4_decode(const uint8_t *h_codes, int64_t n, int64_t dim, const float *h_min, const float *h_diff, float *h_vecs);
/* BinaryQuantizer (binary.go:59-154): threshold = float32(mean in float64); codes [n][ceil(dim/64)*8] */
vg_status vg_bq_train(const float *h_vecs, int64_t n, int64_t dim, float *h_threshold);
vg_status vg_bq_encode(const float *h_vecs, int64_t n, int64_t dim, float threshold, uint8_t *h_codes);
/* RaBitQuantizer.Encode (rabitq.go:51-78): codes [n][ceil(dim/64)*8 + 4] (sign bits ‖ f32 norm LE) */
vg_status vg_rabitq_encode(const float *h_vecs, int64_t n, int64_t dim, uint8_t *h_codes);
/* ProductQuantizer with given int8 codebooks [m][k][dim/m], scales[m], offsets[m] (pq.go:147-229,452-491) */
vg_status vg_pq_encode(const float *h_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, const int8_t *h_codebooks, const float *h_scales, const float *h_offsets, uint8_t *h_codes);
vg_status vg_pq_decode(const uint8_t *h_codes, int64_t n, int64_t dim, int64_t m, int64_t k, const int8_t *h_codebooks, const float *h_scales, const float *h_offsets, float *h_vecs);
vg_status vg_pq_build_distance_table(const float *h_queries, int64_t nq, int64_t dim, int64_t m, int64_t k, const int8_t *h_codebooks, const float *h_scales, const float *h_offsets, float *h_tables /* [nq][m*k] */);
/* ProductQuantizer.Train (pq.go:68-143,275-433): k-means++ (seeded) + `iters` Lloyd
 * iterations per subspace, then int8 codebook quantisation. */
vg_status vg_pq_train(const float *h_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t iters, uint64_t seed,
                      int8_t *h_codebooks, float *h_scales, float *h_offsets, float *h_centroids_f32 /* optional [m][k][dim/m] */);
/* ProductQuantizer.Train (internal/quantization/pq.go:68-143) on a training set that is already device-resident
 * (row-major [n][dim] float32, 16-byte aligned); same outputs, no host copy of the vectors inside the call. */
vg_status vg_pq_train_dev(const float *d_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t iters, uint64_t seed,
                      int8_t *h_codebooks, float *h_scales, float *h_offsets, float *h_centroids_f32 /* optional [m][k][dim/m] */);

/* The same for subspaces [subspace_lo, subspace_hi) only; outputs hold that many subspaces.  ProductQuantizer.Train runs
 * one goroutine per subspace (pq.go:79-140): they are independent, so GPU r of W trains subspaces [r*m/W, (r+1)*m/W) of
 * the same training set and the W slices are concatenated — bit-identical to the single-GPU training (every float32
 * sum still runs in sample order; the seeded k-means++ / re-seeding draws are indexed by the subspace's number). */
vg_status vg_pq_train_range_dev(const float *d_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t iters, uint64_t seed,
                                int64_t subspace_lo, int64_t subspace_hi, int8_t *h_codebooks, float *h_scales, float *h_offsets,
                                float *h_centroids_f32 /* optional */);

/* OptimizedProductQuantizer (opq.go:28-282, svd.go:13-216).  Rotations are [dim/block][block][block] float32,
 * block = vg_opq_block_size(dim, m) (NewOptimizedProductQuantizer's rule, opq.go:41-58).
 * vg_opq_train: `opq_iters` rounds of rotate -> ProductQuantizer.Train (`pq_iters` Lloyd iterations, seed + round)
 *   -> M_b = sum_i x_b^T Decode(Encode(Rx))_b in sample order -> Procrustes (one-sided Jacobi SVD, R = U V^T with the
 *   reflection fix).  Outputs the final rotations and the codebooks of the LAST ProductQuantizer.Train, as the reference.
 * vg_opq_rotate: rotateVector (inverse = 0, simd.Dot per output) or Decode's inverse rotation (inverse = 1).
 * vg_opq_procrustes: computeProcrustesRotation for `blocks` n x n matrices (h_sigma optional: singular values). */
vg_status vg_opq_block_size(int64_t dim, int64_t m, int64_t *block_size);
vg_status vg_opq_train(const float *h_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t opq_iters, int64_t pq_iters,
                       uint64_t seed, float *h_rotations, int8_t *h_codebooks, float *h_scales, float *h_offsets);
vg_status vg_opq_rotate(const float *h_vecs, int64_t n, int64_t dim, int64_t block, const float *h_rotations, int32_t inverse, float *h_out);
vg_status vg_opq_procrustes(const float *h_M, int64_t blocks, int64_t n, float *h_R, float *h_sigma);

/* Device-resident variants (d_vecs / d_codes in HBM; parameters still host arrays): used to
 * encode shards that never exist on the host (flat.Writer.Flush-style bulk encode). */
vg_status vg_minmax_dev(const float *d_vecs, int64_t n, int64_t dim, float *h_mins, float *h_maxs);
vg_status vg_sq8_encode_dev(const float *d_vecs, int64_t n, int64_t dim, const float *h_mins, const float *h_maxs, const float *h_scales, uint8_t *d_codes);
vg_status vg_int4_encode_dev(const float *d_vecs, int64_t n, int64_t dim, const float *h_min, const float *h_diff, uint8_t *d_codes);
vg_status vg_rabitq_encode_dev(const float *d_vecs, int64_t n, int64_t dim, uint8_t *d_codes);
vg_status vg_pq_encode_dev(const float *d_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, const int8_t *h_codebooks, const float *h_scales, const float *h_offsets, uint8_t *d_codes);

/* ----------------------------------------------------------------- k-means
 * internal/kmeans/kmeans.go. */
vg_status vg_kmeans_train(const float *h_vecs, int64_t n, int64_t dim, int64_t k, int32_t metric, int64_t max_iter,
                          const int64_t *h_init_rows /* [k] = Go's rand.Perm(n)[:k] */, uint64_t seed,
                          float *h_centroids, int32_t *h_assign /* optional [n] */, int64_t *iters_run);
vg_status vg_kmeans_assign(const float *h_vecs, int64_t n, int64_t dim, const float *h_centroids, int64_t k, int32_t metric, int32_t *h_assign); /* AssignPartition per row */
vg_status vg_kmeans_find_closest(const float *h_queries, int64_t nq, int64_t dim, const float *h_centroids, int64_t k, int64_t nprobe, int32_t metric, int32_t *h_out /* [nq][min(nprobe,k)] */);

/* ------------------------------------------------------- device scan index
 * One immutable, device-resident code matrix (a flat segment, or one row
 * shard of it).  Search semantics = flat.(*Segment).Search
 * (internal/segment/flat/segment.go:447-752): the k best rows under the
 * CandidateHeap order (score asc for L2 / desc otherwise, then row asc),
 * returned best-first.  Row ids returned are row_base + local row. */
typedef struct {
    int32_t codec;        /* VG_CODEC_* */
    int32_t metric;       /* VG_METRIC_* (ordering: L2 ascending, others descending) */
    int64_t dim;
    int64_t rows;
    uint32_t segment_id;
    uint32_t reserved;
    uint64_t row_base;    /* global id of local row 0 (row sharding) */
    /* codec parameters (host pointers, copied) */
    const float *sq8_mins, *sq8_inv_scales;          /* [dim] */
    const float *int4_min, *int4_diff;               /* [dim] */
    int64_t pq_m, pq_k;
    const int8_t *pq_codebooks;                      /* [m][k][dim/m] */
    const float *pq_scales, *pq_offsets;             /* [m] */
    const float *opq_rotation;                       /* optional [blocks][bs][bs] row-major */
    int64_t opq_block;
    float bq_threshold;                              /* BinaryQuantizer.threshold (queries are sign-coded with it) */
    uint32_t reserved2;
    /* IVF partitions (flat format): rows are partition-ordered */
    int64_t num_partitions;
    const float *centroids;                          /* [P][dim] */
    const uint32_t *partition_offsets;               /* [P+1] */
} vg_index_desc;

vg_status vg_index_create(const vg_index_desc *desc, vg_index_t *out);      /* on the calling thread's device */
vg_status vg_index_create_on(int32_t device, const vg_index_desc *desc, vg_index_t *out);
/* Calls on this handle run on `cuda_stream` (VG_STREAM_LIBRARY: back to thread / leased streams); which GPU holds it. */
vg_status vg_index_set_stream(vg_index_t idx, uint64_t cuda_stream);
vg_status vg_index_device(vg_index_t idx, int32_t *device);
/* Upload rows [row0,row0+n) of codes and/or float vectors from host memory
 * (e.g. the mmap'd segment) through the pinned staging ring.  Either may be
 * NULL.  Code row sizes: F32 none; SQ8 dim; INT4 (dim+1)/2; PQ m; BQ
 * ceil(dim/64)*8; RaBitQ ceil(dim/64)*8+4. */
vg_status vg_index_upload(vg_index_t idx, int64_t row0, int64_t n, const void *h_codes, const float *h_vectors);
/* Fill rows from device memory instead (d_codes / d_vectors already resident). */
vg_status vg_index_upload_dev(vg_index_t idx, int64_t row0, int64_t n, const void *d_codes, const float *d_vectors);
/* Rerank source in HOST memory: float32 rows [rows][dim] (e.g. the mmap'd vector section of the segment) that do not fit
 * next to the codes in HBM (100M x 1536-d = 614 GB).  The library page-locks and maps the region (or uses it as is when
 * the caller already page-locked it); vg_index_rerank* / vg_index_search_rerank / the shard-group rerank then gather the
 * candidate rows over the host link — same arithmetic and bits as the device-resident Segment.Rerank
 * (flat/segment.go:754-781).  Quantized indexes only; the region must stay valid and unchanged until vg_index_close.
 * Part of building the handle (like vg_index_upload): call it before the handle is shared with searching threads. */
vg_status vg_index_set_host_vectors(vg_index_t idx, const float *h_vectors, int64_t rows);
vg_status vg_index_close(vg_index_t idx);
vg_status vg_index_info(vg_index_t idx, int64_t *rows, int64_t *dim, int64_t *code_bytes_per_row, int64_t *device_bytes);

/* Batched Segment.Search.  h_row_mask: optional bitmap (bit r set = row r
 * allowed; segment.Filter.AsBitmap), ceil(rows/8) bytes.  nprobes as
 * model.SearchOptions.NProbes (only used when num_partitions > 1).
 * Outputs are [nq][k]; rows beyond out_counts[q] are 0xFFFFFFFF / NaN. */
vg_status vg_index_search(vg_index_t idx, const float *h_queries, int64_t nq, int64_t k, int64_t nprobes,
                          const uint8_t *h_row_mask, uint32_t *h_out_rows, float *h_out_scores, int32_t *h_out_counts);
/* Same with device-resident queries and outputs (no host copies of queries or results).  The tensor-core filters prove
 * their result per query with a certificate; this call reads the nq certificate flags back (ONE stream synchronisation)
 * and re-runs unproven queries before it returns, so the outputs are final on return.  The exact CUDA-core scan
 * (small / partitioned shapes) has no flags and stays stream-ordered. */
vg_status vg_index_search_dev(vg_index_t idx, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes,
                              const uint8_t *d_row_mask, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts);
/* The same search WITHOUT any host wait, for callers that keep the GPU queue full (needs a caller stream):
 * vg_index_search_dev_async only launches; d_unproven[q] (device, [nq]) is 1 where the certificate of query q did not
 * hold — that query's outputs are then candidates, not yet the reference's result.  vg_index_search_resolve, called
 * with the same arguments whenever the caller synchronises anyway, reads the flags, gives those queries the filter's
 * second chance and the exact scan, and overwrites their outputs; *n_resolved = how many needed it (0 on benign data). */
vg_status vg_index_search_dev_async(vg_index_t idx, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes,
                                    const uint8_t *d_row_mask, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts,
                                    int32_t *d_unproven);
vg_status vg_index_search_resolve(vg_index_t idx, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes,
                                  const uint8_t *d_row_mask, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts,
                                  const int32_t *d_unproven, int64_t *n_resolved);
/* Segment.Search with block-stat skipping (flat/segment.go:524-541,613-630): h_block_keep is a bitmap over the
 * BlockSize = 1024-row blocks of the segment, bit b set = scan block b, clear = the caller's filter.MatchesBlock /
 * matchesFilterSet verdict on block b's statistics was "cannot match" and the block is jumped over.  Only blocks wholly
 * inside the segment are skipped (the ragged last block is always scanned, as `i + BlockSize <= end`); the bitmap holds
 * ceil(floor(rows / 1024) / 8) bytes and always lives in HOST memory (it is the product of host metadata logic, and
 * distance_computations is counted from it).  The row bitmap (host / device as in vg_index_search / _dev) may be NULL.
 * On the device the verdicts are folded into the row bitmap; the tensor-core filters never fetch a 256-row tile
 * without an allowed row, so skipped blocks cost no HBM traffic.  NULL h_block_keep = vg_index_search[_dev]. */
vg_status vg_index_search_blocks(vg_index_t idx, const float *h_queries, int64_t nq, int64_t k, int64_t nprobes, const uint8_t *h_row_mask,
                                 const uint8_t *h_block_keep, uint32_t *h_out_rows, float *h_out_scores, int32_t *h_out_counts);
vg_status vg_index_search_blocks_dev(vg_index_t idx, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes, const uint8_t *d_row_mask,
                                     const uint8_t *h_block_keep, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts);
/* SQ8 filter through tcgen05 kind::i8 (raw code bytes as the unsigned B operand by TMA, query tile quantised to signed
 * 8-bit per query, int32 accumulators; dim % 128 == 0, row-major or lane-transposed codes).  Default on (VECGO_QTC_I8=0);
 * off = the fp16 decode-GEMM.  Results are identical either way (same exact stage, certificate with the measured
 * quantisation error of each query). */
vg_status vg_quant_tc_i8_enable(int32_t on);
vg_status vg_quant_tc_i8_state(int32_t *on);   /* 1 when SQ8 batches of eligible shapes (dim % 128 == 0, <= 1024; 6k <= 2048) take the kind::i8 filter */
/* Tile skipping of the quantized tensor-core filters when a row bitmap is given (default on; VECGO_TILE_SKIP=0): off
 * only for A/B measurements — results are identical either way. */
vg_status vg_tile_skip_enable(int32_t on);
/* IVF-partitioned segments (num_partitions > 1, flat/segment.go:726-745): a query scans only the rows of its nprobes
 * closest partitions.  The (query, partition) pairs of a batch are sorted by partition and scanned as independent
 * virtual queries over that partition's row range, then merged per query: work and HBM traffic are nprobes /
 * num_partitions of the full scan.  Off (VECGO_IVF_GROUPED=0) = the full scan with a per-row partition test — same
 * results, for A/B measurements only. */
vg_status vg_ivf_grouped_enable(int32_t on);
/* Counters of the last vg_index_search* / vg_index_search_resolve call made by the calling thread: what a Go caller adds
 * to searcher.FilterGateStats / model.QueryStats (flat/segment.go:448-471,553-591).  distance_computations counts the
 * (query, row) distance evaluations the reference's scan would report (rows visited per query, plus the rows of every
 * exact re-run); the rest describes how the batch was answered. */
typedef struct {
    uint64_t queries;               /* queries in the call */
    uint64_t distance_computations; /* FilterGateStats.DistanceComputations equivalent */
    uint64_t filter_queries;        /* answered through a tensor-core filter + certificate */
    uint64_t second_chance_queries; /* certificate failed once: filter re-run (Flat: twice the candidate groups; quantized: threshold pass) */
    uint64_t exact_rerun_queries;   /* no proof: re-run on the exact CUDA-core scan */
    uint64_t threshold_pass_queries;/* answered by the threshold pass (every row within the error bound of the k-th best listed
                                       and scored exactly): the second chance of the quantized filters, and the whole batch
                                       once an index has shown tightly clustered data */
} vg_search_stats;
vg_status vg_last_search_stats(vg_search_stats *out);
/* Batched Segment.Rerank (flat/segment.go:754-781): exact SquaredL2/Dot of each
 * query against its r candidate rows (local row ids); rows >= RowCount give NaN. */
vg_status vg_index_rerank(vg_index_t idx, const float *h_queries, int64_t nq, const uint32_t *h_rows, int64_t r, float *h_scores);
vg_status vg_index_rerank_dev(vg_index_t idx, const float *d_queries, int64_t nq, const uint32_t *d_rows, int64_t r, float *d_scores);
/* Quantized gather scoring: the codec's own distance of every query to ITS r candidate rows (local row ids; rows >=
 * RowCount give NaN), in the reference's arithmetic — what the DiskANN traversal evaluates per neighbour
 * (internal/segment/diskann/segment.go:511-588: pq.AdcDistance, int4.L2Distance, rabitq.Distance), batched.  SQ8
 * (L2), INT4, PQ / OPQ (K = 256), RaBitQ and BQ (Hamming distance as float32); a float32 index scores exactly
 * (= vg_index_rerank). */
vg_status vg_index_score(vg_index_t idx, const float *h_queries, int64_t nq, const uint32_t *h_rows, int64_t r, float *h_scores);
vg_status vg_index_score_dev(vg_index_t idx, const float *d_queries, int64_t nq, const uint32_t *d_rows, int64_t r, float *d_scores);
/* INT4 gather scoring follows Int4Quantizer.L2Distance (int4.go:136-147): a trained / unmarshalled quantizer always has
 * its lookup table, so the live path — and what DiskANN executes (diskann/segment.go:564,578) — is
 * simd.Int4L2DistancePrecomputed over simd.BuildInt4LookupTable's values: VG_INT4_SCORE_LUT, the default.
 * VG_INT4_SCORE_DIRECT selects simd.Int4L2Distance (FMA dequantisation, the nil-table fallback). */
enum { VG_INT4_SCORE_LUT = 0, VG_INT4_SCORE_DIRECT = 1 };
vg_status vg_index_set_int4_score_mode(vg_index_t idx, int32_t mode);
/* simd.BuildInt4LookupTable (internal/simd/kernels.go:94-103): h_table [dim*16]. */
vg_status vg_int4_build_lookup_table(const float *h_min, const float *h_diff, int64_t dim, float *h_table);
/* distance.SquaredL2Bounded / simd.SquaredL2Bounded (distance/distance.go:24-31, internal/simd/kernels.go:163-217, AVX-512
 * kernel internal/simd/src/bounded_l2_avx512.c:19-107) of every query against ITS r candidate rows of a float32 index:
 * scores[q][j] is the distance, or the partial sum at the 64-dimension block where it first exceeded the bound, in
 * which case exceeded[q][j] = 1.  Bounds: [nq] (one per query: the traversal's current worst result) or [nq][r] when
 * per_pair_bounds.  Rows >= RowCount give NaN / 0. */
vg_status vg_index_l2_bounded(vg_index_t idx, const float *h_queries, int64_t nq, const uint32_t *h_rows, int64_t r, const float *h_bounds,
                              int32_t per_pair_bounds, float *h_scores, uint8_t *h_exceeded);
vg_status vg_index_l2_bounded_dev(vg_index_t idx, const float *d_queries, int64_t nq, const uint32_t *d_rows, int64_t r,
                                  const float *d_bounds, int32_t per_pair_bounds, float *d_scores, uint8_t *d_exceeded);
/* Approximate scan to top-r, exact rerank, final top-k (engine refine path,
 * internal/engine/search.go:188-192,913-973), all on device. */
vg_status vg_index_search_rerank(vg_index_t idx, const float *h_queries, int64_t nq, int64_t r, int64_t k,
                                 uint32_t *h_out_rows, float *h_out_scores, int32_t *h_out_counts);

/* Tensor-core Flat filter (csrc/vg_flat_tc.cu).  vg_index_search / vg_index_search_dev on a float32 index run
 * flat.(*Segment).Search (internal/segment/flat/segment.go:447-752) as a tcgen05 TF32 GEMM whose epilogue keeps the
 * two smallest approximate scores of every group of G consecutive rows (and which row is the smallest); the arg-min
 * rows of the kc groups with the smallest minima are then scored EXACTLY in simd.SquaredL2 / simd.Dot order and an error-bound certificate proves that the result equals
 * the exact scan's; queries without a proof are re-run on the exact scan.
 * vg_flat_tc_enable(0) forces the exact CUDA-core scan (also: environment VECGO_FLAT_TC=0).
 * vg_flat_tc_stats: queries that went through the filter, and how many of them needed the exact re-run.
 * vg_flat_tc_candidates (diagnostic): per query the threshold tau (kc-th smallest group minimum of the APPROXIMATE
 * s-space score, L2: ||x||^2 - 2 q.x, dot: -q.x), and for each of the kc selected groups (h_groups [nq][kc], best
 * first, group g = rows [g*G, (g+1)*G)) either the row that attains the group minimum or 0x80000000|g when the
 * group's second smallest value is <= tau as well (then all its rows are scored exactly); *group_rows = G. */
vg_status vg_flat_tc_enable(int32_t on);
vg_status vg_flat_tc_stats(uint64_t *queries, uint64_t *fallbacks);
vg_status vg_flat_tc_candidates(vg_index_t idx, const float *h_queries, int64_t nq, int64_t kc, uint32_t *h_groups, int32_t *h_counts,
                                float *h_tau, int64_t *group_rows);

/* Tensor-core filter of the quantized scans (csrc/vg_quant_tc.cu).  vg_index_search / vg_index_search_dev on an
 * SQ8 / INT4 / PQ / OPQ index (L2, no IVF partitions, dim % 64 == 0, k <= 1024, batches of >= 16 queries over >= 8192
 * rows) run simd.Sq8uL2BatchPerDimension / simd.Int4L2DistanceBatch / simd.PqAdcLookup
 * (internal/segment/flat/segment.go:543-552,603-611) as a tcgen05 GEMM over codes that are decoded inside the
 * kernel — kind::i8 (8-bit integer operands, see vg_quant_tc_i8_enable) where the shape allows, fp16 otherwise; the candidates are re-scored in the reference's exact float32 order and a certificate proves the result
 * equals the exact scan's (queries without a proof are re-run on the exact CUDA-core scan).  Environment
 * VECGO_QUANT_TC=0 (or vg_flat_tc_enable(0)) forces the CUDA-core scan.  Counters: queries that went through the
 * filter and how many of them needed the exact re-run. */
vg_status vg_quant_tc_stats(uint64_t *queries, uint64_t *fallbacks);
/* PQ training (pq.go:347-386 assignClusters): (sample, subspace) pairs whose nearest centroid was found on the tensor
 * cores (8-dim subspaces x 256 centroids, vg_pq_assign_tc.cu) and how many of them failed the gap certificate and were
 * re-evaluated by the exact sequential-FMA loop, since the library was loaded.  VECGO_PQ_ASSIGN_TC=0 disables the path. */
vg_status vg_pq_assign_tc_stats(uint64_t *pairs, uint64_t *fallback_pairs);
/* Measurement aid: enable = 1 brackets every GEMM launch of the filter with a CUDA-event pair on its own stream and
 * resets the counters, 0 turns that off, < 0 only reads.  Returns the accumulated kernel time (ms) and launches. */
vg_status vg_quant_tc_profile(int32_t enable, double *gemm_ms, uint64_t *gemm_launches);

/* flat.Open (internal/segment/flat/segment.go:105-342) on the raw file bytes:
 * header decode, optional CRC32C verify, section views, staging to HBM. */
vg_status vg_flat_open(const uint8_t *h_file, size_t len, int32_t verify_checksum, vg_index_t *out);
vg_status vg_flat_open_on(int32_t device, const uint8_t *h_file, size_t len, int32_t verify_checksum, vg_index_t *out);
/* Header fields of an opened flat segment (format.go:28-51). */
typedef struct {
    uint64_t segment_id;
    uint32_t row_count, dim, metric, num_partitions, quantization_type, checksum;
} vg_flat_header;
vg_status vg_flat_decode_header(const uint8_t *h_file, size_t len, vg_flat_header *out);
/* model.ID column of the segment for the given local rows (FetchIDs). */
vg_status vg_index_fetch_ids(vg_index_t idx, const uint32_t *h_rows, int64_t n, uint64_t *h_ids);

/* ---------------------------------------------------------- top-k merging
 * Merge `lists` best-first candidate lists per query (per-shard results after
 * the NCCL allgather, or per-segment heaps: engine/search.go:903-908) into the
 * k best under the CandidateHeap order.  Layout [lists][nq][k_in]; entries
 * with row 0xFFFFFFFF are empty. */
vg_status vg_topk_merge_dev(const uint32_t *d_rows, const float *d_scores, int64_t lists, int64_t nq, int64_t k_in,
                            int32_t descending, int64_t k_out, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts);
/* Packed exchange format of the sharded search: one sortable 8-byte key per candidate ((score under the heap order) << 32
 * | row; empty = all ones).  vg_topk_pack_dev turns a shard's [nq][k] (rows, scores) into keys; the all-gather then
 * moves ONE buffer per rank, and vg_topk_merge_keys_dev merges the gathered [lists][nq][k_in] keys without a
 * conversion pass. */
vg_status vg_topk_pack_dev(const uint32_t *d_rows, const float *d_scores, int64_t n, int32_t descending, uint64_t *d_keys);
vg_status vg_topk_merge_keys_dev(const uint64_t *d_keys, int64_t lists, int64_t nq, int64_t k_in, int32_t descending, int64_t k_out,
                                 uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts);
vg_status vg_topk_merge(const uint32_t *h_rows, const float *h_scores, int64_t lists, int64_t nq, int64_t k_in,
                        int32_t descending, int64_t k_out, uint32_t *h_out_rows, float *h_out_scores, int32_t *h_out_counts);

/* ------------------------------------------------------------ shard groups
 * Row shards of one database on several GPUs with the merge exchange INSIDE the library (SURVEY 8b "Ownership": the
 * library owns the NCCL communicators; 8e): every GPU scans its shard (global row id = row_base + local row, so the
 * reference's (score, SegmentID, RowID) order is (score, global row) with one logical segment), ONE ncclAllGather moves
 * the per-shard top-k as 8-byte sortable keys, every GPU merges.  The rerank form keeps the reference's "global
 * approximate top-r, then exact" semantics (engine/search.go:188-192,913-973) with a second exchange, so ids do not
 * depend on the number of GPUs.  NCCL is loaded at run time (libnccl.so.2).
 *   single process (a Go host): vg_shard_group_create(devices, W); every call takes the W shard handles (handle i must
 *     live on devices[i]: vg_index_create_on) and fans out over W host threads; host results come from GPU 0.
 *   one process per GPU: vg_nccl_unique_id on rank 0 -> ship the 128 bytes -> vg_shard_group_create_rank everywhere; every
 *     call takes this process's ONE shard handle and must be made by all ranks.
 * The `_dev` forms take per-member device pointers ([members] arrays): queries [nq][dim] replicated on every GPU,
 * outputs [nq][k] filled on every GPU.  Calls on one group are serialised (collectives must be issued in one order). */
typedef uint64_t vg_shard_group_t;
vg_status vg_nccl_unique_id(uint8_t *id128);
vg_status vg_shard_group_create(const int32_t *devices, int32_t n, vg_shard_group_t *out);
vg_status vg_shard_group_create_rank(const uint8_t *id128, int32_t rank, int32_t world, int32_t device, vg_shard_group_t *out);
vg_status vg_shard_group_info(vg_shard_group_t group, int32_t *world, int32_t *members, int32_t *first_rank);
vg_status vg_shard_group_destroy(vg_shard_group_t group);
vg_status vg_shard_group_search(vg_shard_group_t group, const vg_index_t *shards, const float *h_queries, int64_t nq, int64_t k,
                                uint32_t *h_out_rows, float *h_out_scores, int32_t *h_out_counts);
vg_status vg_shard_group_search_rerank(vg_shard_group_t group, const vg_index_t *shards, const float *h_queries, int64_t nq, int64_t r, int64_t k,
                                       uint32_t *h_out_rows, float *h_out_scores, int32_t *h_out_counts);
vg_status vg_shard_group_search_dev(vg_shard_group_t group, const vg_index_t *shards, const float *const *d_queries, int64_t nq, int64_t k,
                                    uint32_t *const *d_out_rows, float *const *d_out_scores, int32_t *const *d_out_counts);
vg_status vg_shard_group_search_rerank_dev(vg_shard_group_t group, const vg_index_t *shards, const float *const *d_queries, int64_t nq, int64_t r,
                                           int64_t k, uint32_t *const *d_out_rows, float *const *d_out_scores, int32_t *const *d_out_counts);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* VECGO_CUDA_H */
