// vg_quant.cu — quantizer Train / Encode / Decode kernels (device side of
// internal/quantization).  Encoders restate the reference's scalar Go
// arithmetic: every float32 op individually rounded (Go on amd64 never fuses),
// truncating float→uint8 casts, math.Round = half away from zero.
#include "vg_quant.cuh"

namespace vg {

// ------------------------------------------------------------ per-dim min/max
// ScalarQuantizer.Train (quantizer.go:130-180) / Int4Quantizer.Train
// (int4.go:29-65).  min/max are order independent, so a column-parallel
// two-stage reduction is exact.
__global__ void __launch_bounds__(256) minmax_partial_kernel(const float *v, int64_t n, int64_t dim, int64_t stride, int64_t rows_per_block,
                                                             float *pmin, float *pmax) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dim) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > n) r1 = n;
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    for (int64_t r = r0; r < r1; r++) {
        const float x = v[r * stride + d];
        if (x < mn) mn = x;
        if (x > mx) mx = x;
    }
    pmin[(int64_t)blockIdx.y * dim + d] = mn;
    pmax[(int64_t)blockIdx.y * dim + d] = mx;
}
__global__ void __launch_bounds__(256) minmax_final_kernel(const float *pmin, const float *pmax, int64_t parts, int64_t dim,
                                                           float *mins, float *maxs) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dim) return;
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    for (int64_t p = 0; p < parts; p++) {
        const float a = pmin[p * dim + d], b = pmax[p * dim + d];
        if (a < mn) mn = a;
        if (b > mx) mx = b;
    }
    mins[d] = mn;
    maxs[d] = mx;
}
vg_status dev_minmax(const float *d_vecs, int64_t n, int64_t dim, float *d_mins, float *d_maxs, cudaStream_t st) {
    return dev_minmax_strided(d_vecs, n, dim, dim, d_mins, d_maxs, st);
}
// min / max of the first `dim` columns of rows that are `stride` floats apart (a column slice of a wider matrix)
vg_status dev_minmax_strided(const float *d_vecs, int64_t n, int64_t dim, int64_t stride, float *d_mins, float *d_maxs, cudaStream_t st) {
    int64_t parts = (n + 1023) / 1024;
    if (parts > 512) parts = 512;
    if (parts < 1) parts = 1;
    const int64_t rpb = (n + parts - 1) / parts;
    parts = (n + rpb - 1) / rpb;
    DevBuf scratch;
    VG_TRY(scratch.alloc((size_t)parts * dim * 8));
    float *pmin = scratch.as<float>(), *pmax = pmin + parts * dim;
    dim3 grid((unsigned)((dim + 255) / 256), (unsigned)parts);
    minmax_partial_kernel<<<grid, 256, 0, st>>>(d_vecs, n, dim, stride, rpb, pmin, pmax);
    VG_LAUNCHED();
    minmax_final_kernel<<<(unsigned)((dim + 255) / 256), 256, 0, st>>>(pmin, pmax, parts, dim, d_mins, d_maxs);
    VG_LAUNCHED();
    VG_CUDA(cudaStreamSynchronize(st));
    return VG_OK;
}

// ------------------------------------------------------------ SQ8
// EncodeInto (quantizer.go:200-222)
__global__ void __launch_bounds__(256) sq8_encode_kernel(const float *v, int64_t total, int64_t dim, const float *mins,
                                                         const float *maxs, const float *scales, uint8_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t d = i % dim;
    float val = v[i];
    const float mn = mins[d], mx = maxs[d];
    if (val < mn) val = mn;
    else if (val > mx) val = mx;
    const float nrm = __fmul_rn(__fsub_rn(val, mn), scales[d]);
    out[i] = (uint8_t)__float2uint_rz(__fadd_rn(nrm, 0.5f));
}
// DecodeInto (quantizer.go:241-248)
__global__ void __launch_bounds__(256) sq8_decode_kernel(const uint8_t *codes, int64_t total, int64_t dim, const float *mins,
                                                         const float *inv, float *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t d = i % dim;
    out[i] = __fadd_rn(__fmul_rn((float)codes[i], inv[d]), mins[d]);
}
vg_status dev_sq8_encode(const float *d_vecs, int64_t n, int64_t dim, const float *d_mins, const float *d_maxs,
                         const float *d_scales, uint8_t *d_codes, cudaStream_t st) {
    const int64_t total = n * dim;
    if (total <= 0) return VG_OK;
    sq8_encode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_vecs, total, dim, d_mins, d_maxs, d_scales, d_codes);
    VG_LAUNCHED();
    return VG_OK;
}
vg_status dev_sq8_decode(const uint8_t *d_codes, int64_t n, int64_t dim, const float *d_mins, const float *d_inv, float *d_vecs,
                         cudaStream_t st) {
    const int64_t total = n * dim;
    if (total <= 0) return VG_OK;
    sq8_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_codes, total, dim, d_mins, d_inv, d_vecs);
    VG_LAUNCHED();
    return VG_OK;
}

// ------------------------------------------------------------ INT4
// Encode (int4.go:68-106): norm = (v-min)/diff in f32, clamp [0,1],
// byte(math.Round(float64(norm)*15)); high nibble = even dim.
__device__ __forceinline__ uint32_t int4_quant(float v, float mn, float df) {
    float norm = __fdiv_rn(__fsub_rn(v, mn), df);
    if (norm < 0.0f) norm = 0.0f;
    else if (norm > 1.0f) norm = 1.0f;
    return (uint32_t)round((double)norm * 15.0);
}
__global__ void __launch_bounds__(256) int4_encode_kernel(const float *v, int64_t n, int64_t dim, const float *mn,
                                                          const float *df, uint8_t *out) {
    const int64_t cs = (dim + 1) / 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * cs) return;
    const int64_t r = i / cs, b = i - r * cs;
    const int64_t d = 2 * b;
    const uint32_t q1 = int4_quant(v[r * dim + d], mn[d], df[d]);
    uint32_t q2 = 0;
    if (d + 1 < dim) q2 = int4_quant(v[r * dim + d + 1], mn[d + 1], df[d + 1]);
    out[i] = (uint8_t)((q1 << 4) | (q2 & 0x0F));
}
// Decode (int4.go:109-132): float32(q)/15.0*diff + min, unfused, true divide.
__global__ void __launch_bounds__(256) int4_decode_kernel(const uint8_t *codes, int64_t n, int64_t dim, const float *mn,
                                                          const float *df, float *out) {
    const int64_t cs = (dim + 1) / 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * dim) return;
    const int64_t r = i / dim, d = i - r * dim;
    const uint32_t b = codes[r * cs + (d >> 1)];
    const float q = (float)((d & 1) ? (b & 0x0F) : (b >> 4));
    out[i] = __fadd_rn(__fmul_rn(__fdiv_rn(q, 15.0f), df[d]), mn[d]);
}
vg_status dev_int4_encode(const float *d_vecs, int64_t n, int64_t dim, const float *d_min, const float *d_diff, uint8_t *d_codes,
                          cudaStream_t st) {
    const int64_t total = n * ((dim + 1) / 2);
    if (total <= 0) return VG_OK;
    int4_encode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_vecs, n, dim, d_min, d_diff, d_codes);
    VG_LAUNCHED();
    return VG_OK;
}
vg_status dev_int4_decode(const uint8_t *d_codes, int64_t n, int64_t dim, const float *d_min, const float *d_diff, float *d_vecs,
                          cudaStream_t st) {
    const int64_t total = n * dim;
    if (total <= 0) return VG_OK;
    int4_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_codes, n, dim, d_min, d_diff, d_vecs);
    VG_LAUNCHED();
    return VG_OK;
}

// ------------------------------------------------------------ BQ / RaBitQ
// BinaryQuantizer.Train (binary.go:59-81): float64 sum of every value.  The
// reference adds sequentially; we reduce in float64 with a fixed tree, which
// agrees to ~1e-16 relative, then round once to float32 like the reference.
__global__ void __launch_bounds__(256) sum_f64_kernel(const float *v, int64_t total, double *partial) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        s += (double)v[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
vg_status dev_mean_f64(const float *d_vecs, int64_t total, double *h_sum, cudaStream_t st) {
    const int blocks = 256;
    DevBuf part;
    VG_TRY(part.alloc(blocks * sizeof(double)));
    sum_f64_kernel<<<blocks, 256, 0, st>>>(d_vecs, total, part.as<double>());
    VG_LAUNCHED();
    double h[256];
    VG_CUDA(cudaMemcpyAsync(h, part.p, sizeof h, cudaMemcpyDeviceToHost, st));
    VG_CUDA(cudaStreamSynchronize(st));
    double s = 0.0;
    for (int i = 0; i < blocks; i++) s += h[i];
    *h_sum = s;
    return VG_OK;
}

// Sign bits (binary.go:87-114, rabitq.go:61-72) + optional norm
// (rabitq.go:57-58: Sqrt(simd.Dot(v,v)), AVX-512 order).  Half-warp per row.
__global__ void __launch_bounds__(256) sign_encode_kernel(const float *v, int64_t n, int64_t dim, float threshold, int words32,
                                                          int64_t out_stride, uint8_t *out, bool with_norm) {
    const int64_t hwid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const bool live = hwid < n;
    const int64_t r = live ? hwid : n - 1;
    const float *x = v + r * dim;
    uint8_t *o = out + r * out_stride;
    if (live)
        for (int w = lane; w < words32; w += 16) {
            uint32_t bits = 0;
            for (int b = 0; b < 32; b++) {
                const int64_t d = (int64_t)w * 32 + b;
                if (d < dim && x[d] >= threshold) bits |= 1u << b;
            }
            o[w * 4 + 0] = (uint8_t)bits;
            o[w * 4 + 1] = (uint8_t)(bits >> 8);
            o[w * 4 + 2] = (uint8_t)(bits >> 16);
            o[w * 4 + 3] = (uint8_t)(bits >> 24);
        }
    if (with_norm) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        const int64_t epochs = dim >> 6;
        for (int64_t e = 0; e < epochs; e++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float t = x[e * 64 + j * 16 + lane];
                a[j] = __fmaf_rn(t, t, a[j]);
            }
        float tot = reduce16(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])));
        if (lane == 0 && live) {
            for (int64_t d = epochs * 64; d < dim; d++) tot = __fmaf_rn(x[d], x[d], tot);
            const float norm = (float)sqrt((double)tot);
            const uint32_t u = __float_as_uint(norm);
            uint8_t *p = o + (int64_t)words32 * 4;
            p[0] = (uint8_t)u;
            p[1] = (uint8_t)(u >> 8);
            p[2] = (uint8_t)(u >> 16);
            p[3] = (uint8_t)(u >> 24);
        }
    }
}
vg_status dev_sign_encode(const float *d_vecs, int64_t n, int64_t dim, float threshold, bool with_norm, uint8_t *d_codes,
                          cudaStream_t st) {
    if (n <= 0) return VG_OK;
    const int words32 = (int)(((dim + 63) / 64) * 2);
    const int64_t stride = (int64_t)words32 * 4 + (with_norm ? 4 : 0);
    const int64_t threads = n * 16;
    sign_encode_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_vecs, n, dim, threshold, words32, stride, d_codes,
                                                                          with_norm);
    VG_LAUNCHED();
    return VG_OK;
}

// ------------------------------------------------------------ PQ
// simd.SquaredL2Int8Dequantized (kernels.go:354-362): sequential, unfused.
__device__ __forceinline__ float pq_entry(const float *x, const int8_t *cb, int ds, float scale, float offset) {
    float sum = 0.0f;
    for (int i = 0; i < ds; i++) {
        const float vv = __fadd_rn(__fmul_rn((float)cb[i], scale), offset);
        const float d = __fsub_rn(x[i], vv);
        sum = __fadd_rn(sum, __fmul_rn(d, d));
    }
    return sum;
}
// ProductQuantizer.Encode (pq.go:147-182) → simd.FindNearestCentroidInt8
// (kernels.go:376-396): strict <, first minimum wins.  grid = (row chunks, m).
__global__ void __launch_bounds__(256) pq_encode_kernel(const float *v, int64_t n, int64_t dim, int m_total, int k, int ds,
                                                        const int8_t *codebooks, const float *scales, const float *offsets,
                                                        uint8_t *codes) {
    extern __shared__ __align__(16) unsigned char smem[];
    int8_t *cb = reinterpret_cast<int8_t *>(smem);
    const int m = blockIdx.y;
    for (int i = threadIdx.x; i < k * ds; i += blockDim.x) cb[i] = codebooks[(int64_t)m * k * ds + i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float *x = v + r * dim + (int64_t)m * ds;
    const float scale = scales[m], offset = offsets[m];
    int best = 0;
    float bd = pq_entry(x, cb, ds, scale, offset);
    for (int c = 1; c < k; c++) {
        const float d = pq_entry(x, cb + c * ds, ds, scale, offset);
        if (d < bd) {
            bd = d;
            best = c;
        }
    }
    codes[r * m_total + m] = (uint8_t)best;
}
// Decode (pq.go:185-229)
__global__ void __launch_bounds__(256) pq_decode_kernel(const uint8_t *codes, int64_t n, int64_t dim, int m_total, int k, int ds,
                                                        const int8_t *codebooks, const float *scales, const float *offsets,
                                                        float *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * dim) return;
    const int64_t r = i / dim, d = i - r * dim;
    const int m = (int)(d / ds), j = (int)(d - (int64_t)m * ds);
    const int c = codes[r * m_total + m];
    out[i] = __fadd_rn(__fmul_rn((float)codebooks[((int64_t)m * k + c) * ds + j], scales[m]), offsets[m]);
}
// BuildDistanceTable (pq.go:452-491): table[q][m*k + c]
__global__ void __launch_bounds__(256) pq_table_kernel(const float *queries, int64_t nq, int64_t dim, int m_total, int k, int ds,
                                                       const int8_t *codebooks, const float *scales, const float *offsets,
                                                       float *tables) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)m_total * k;
    if (i >= nq * per) return;
    const int64_t q = i / per, e = i - q * per;
    const int m = (int)(e / k), c = (int)(e - (int64_t)m * k);
    tables[i] = pq_entry(queries + q * dim + (int64_t)m * ds, codebooks + ((int64_t)m * k + c) * ds, ds, scales[m], offsets[m]);
}
vg_status dev_pq_encode(const float *d_vecs, int64_t n, int64_t dim, int m, int k, const int8_t *d_cb, const float *d_scales,
                        const float *d_offsets, uint8_t *d_codes, cudaStream_t st) {
    if (n <= 0) return VG_OK;
    const int ds = (int)(dim / m);
    const size_t sm = (size_t)k * ds;
    if (sm > 200 * 1024) return fail(VG_ERR_UNSUPPORTED, "PQ codebook of one subspace exceeds shared memory");
    VG_CUDA(cudaFuncSetAttribute(pq_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)m);
    pq_encode_kernel<<<grid, 256, sm, st>>>(d_vecs, n, dim, m, k, ds, d_cb, d_scales, d_offsets, d_codes);
    VG_LAUNCHED();
    return VG_OK;
}
vg_status dev_pq_decode(const uint8_t *d_codes, int64_t n, int64_t dim, int m, int k, const int8_t *d_cb, const float *d_scales,
                        const float *d_offsets, float *d_vecs, cudaStream_t st) {
    const int64_t total = n * dim;
    if (total <= 0) return VG_OK;
    pq_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_codes, n, dim, m, k, (int)(dim / m), d_cb, d_scales,
                                                                      d_offsets, d_vecs);
    VG_LAUNCHED();
    return VG_OK;
}
vg_status dev_pq_tables(const float *d_queries, int64_t nq, int64_t dim, int m, int k, const int8_t *d_cb, const float *d_scales,
                        const float *d_offsets, float *d_tables, cudaStream_t st) {
    const int64_t total = nq * m * k;
    if (total <= 0) return VG_OK;
    pq_table_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_queries, nq, dim, m, k, (int)(dim / m), d_cb, d_scales,
                                                                     d_offsets, d_tables);
    VG_LAUNCHED();
    return VG_OK;
}

// ------------------------------------------------------------ misc
// simd.ScaleInPlace (floats_avx512.c:174-217)
__global__ void __launch_bounds__(256) scale_kernel(float *a, int64_t n, float s) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = __fmul_rn(a[i], s);
}
vg_status dev_scale(float *d_a, int64_t n, float s, cudaStream_t st) {
    if (n <= 0) return VG_OK;
    scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_a, n, s);
    VG_LAUNCHED();
    return VG_OK;
}
// distance.NormalizeL2InPlace (distance.go:42-53): norm2 = simd.Dot(v,v) (AVX-512
// order), inv = 1/Sqrt(norm2) with Sqrt through float64, v *= inv.
__global__ void __launch_bounds__(256) normalize_kernel(float *v, int64_t n, int64_t dim, uint8_t *ok) {
    const int64_t hwid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const bool live = hwid < n;
    const int64_t r = live ? hwid : n - 1;
    float *x = v + r * dim;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t epochs = dim >> 6;
    for (int64_t e = 0; e < epochs; e++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float t = x[e * 64 + j * 16 + lane];
            a[j] = __fmaf_rn(t, t, a[j]);
        }
    float tot = reduce16(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])));
    if (lane == 0)
        for (int64_t d = epochs * 64; d < dim; d++) tot = __fmaf_rn(x[d], x[d], tot);
    tot = __shfl_sync(0xffffffffu, tot, 0, 16);
    if (!live) return;
    if (tot == 0.0f || dim == 0) {
        if (lane == 0 && ok) ok[r] = 0;
        return;
    }
    const float inv = __fdiv_rn(1.0f, (float)sqrt((double)tot));
    for (int64_t d = lane; d < dim; d += 16) x[d] = __fmul_rn(x[d], inv);
    if (lane == 0 && ok) ok[r] = 1;
}
vg_status dev_normalize(float *d_v, int64_t n, int64_t dim, uint8_t *d_ok, cudaStream_t st) {
    if (n <= 0) return VG_OK;
    const int64_t threads = n * 16;
    normalize_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_v, n, dim, d_ok);
    VG_LAUNCHED();
    return VG_OK;
}

// ------------------------------------------------------------ device layout permutes
// SQ8 fast path: inside each block of 16*VB dims move byte (step s, lane l)
// from 16*s + l to l*VB + s.
__global__ void __launch_bounds__(256) permute_sq8_kernel(const uint8_t *src, uint8_t *dst, int64_t n, int64_t dim, int vb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * dim) return;
    const int64_t r = i / dim, d = i - r * dim;
    const int blk = 16 * vb;
    const int64_t b = d / blk, o = d - b * blk;
    const int s = (int)(o >> 4), l = (int)(o & 15);
    dst[r * dim + b * blk + l * vb + s] = src[i];
}
// INT4 fast path (dim % 256 == 0): byte index within a 128-byte block is
// 32*e + 8*blk + p (epoch e, 16-dim block blk, lane pair p) → 16*p + 4*e + blk.
__global__ void __launch_bounds__(256) permute_int4_kernel(const uint8_t *src, uint8_t *dst, int64_t n, int64_t cs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * cs) return;
    const int64_t r = i / cs, o = i - r * cs;
    const int64_t b = o >> 7;
    const int w = (int)(o & 127);
    const int e = w >> 5, blk = (w >> 3) & 3, p = w & 7;
    dst[r * cs + b * 128 + 16 * p + 4 * e + blk] = src[i];
}
vg_status dev_permute_sq8(const uint8_t *d_src, uint8_t *d_dst, int64_t n, int64_t dim, int vb, cudaStream_t st) {
    const int64_t total = n * dim;
    if (total <= 0) return VG_OK;
    permute_sq8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_src, d_dst, n, dim, vb);
    VG_LAUNCHED();
    return VG_OK;
}
vg_status dev_permute_int4(const uint8_t *d_src, uint8_t *d_dst, int64_t n, int64_t cs, cudaStream_t st) {
    const int64_t total = n * cs;
    if (total <= 0) return VG_OK;
    permute_int4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_src, d_dst, n, cs);
    VG_LAUNCHED();
    return VG_OK;
}
// PQ fast path: rows are grouped in tiles of 32; inside a tile the 16 codes of
// chunk t (subspaces 16t..16t+15) of row r sit at t*512 + r*16, so one warp's
// 16-byte loads of a chunk cover 512 contiguous bytes (thread-per-row scan).
// mpad = M rounded up to 16; missing subspaces are zero bytes (never looked up).
__global__ void __launch_bounds__(256) permute_pq_kernel(const uint8_t *src, uint8_t *dst, int64_t row0, int64_t n, int m, int mpad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * mpad) return;
    const int64_t r = i / mpad;
    const int o = (int)(i - r * mpad);
    const int64_t g = row0 + r;
    dst[(g >> 5) * (32 * (int64_t)mpad) + (int64_t)(o >> 4) * 512 + (g & 31) * 16 + (o & 15)] = (o < m) ? src[r * m + o] : (uint8_t)0;
}
vg_status dev_permute_pq(const uint8_t *d_src, uint8_t *d_dst_base, int64_t row0, int64_t n, int m, int mpad, cudaStream_t st) {
    const int64_t total = n * mpad;
    if (total <= 0) return VG_OK;
    permute_pq_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_src, d_dst_base, row0, n, m, mpad);
    VG_LAUNCHED();
    return VG_OK;
}
// Sign-bit rows (BQ: bits; RaBitQ: bits ‖ f32 norm, 196 B for 1536-d) → a
// 16-byte aligned, zero padded bit plane (+ a separate norm column).
__global__ void __launch_bounds__(256) split_sign_kernel(const uint8_t *src, int64_t n, int64_t nbytes, int64_t src_stride,
                                                         int64_t dst_stride, uint8_t *bits, float *norms) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = dst_stride + (norms ? 4 : 0);
    if (i >= n * per) return;
    const int64_t r = i / per, o = i - r * per;
    if (o < dst_stride) bits[r * dst_stride + o] = (o < nbytes) ? src[r * src_stride + o] : (uint8_t)0;
    else reinterpret_cast<uint8_t *>(norms)[r * 4 + (o - dst_stride)] = src[r * src_stride + nbytes + (o - dst_stride)];
}
vg_status dev_split_sign(const uint8_t *d_src, int64_t n, int64_t nbytes, int64_t src_stride, int64_t dst_stride, uint8_t *d_bits,
                         float *d_norms, cudaStream_t st) {
    const int64_t total = n * (dst_stride + (d_norms ? 4 : 0));
    if (total <= 0) return VG_OK;
    split_sign_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_src, n, nbytes, src_stride, dst_stride, d_bits, d_norms);
    VG_LAUNCHED();
    return VG_OK;
}

}  // namespace vg
