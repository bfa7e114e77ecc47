/*
 * vecgo_oracle.c — CPU restatement of vecgo's vector-scan hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under vecgo_b200/ may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * CPU arm.  The shipped product path is the CUDA library (libvecgo_cuda.so).
 *
 * Parity pinning: every "_a512" function below restates the arithmetic of the
 * reference's AVX-512 C kernels (the ISA an x86 B200 host selects,
 * internal/simd/capability.go:145-155) in scalar C, in the same summation
 * order, with fmaf() exactly where the shipped assembly fuses.  tests/ checks
 * them BIT-FOR-BIT against the reference's own C sources compiled into
 * oracle/_ref (see oracle/Makefile) and against the reference's known-answer
 * tests (internal/simd/floats_test.go, quantization/*_test.go ...).
 * The "_generic" functions restate the pure-Go kernels
 * (internal/simd/kernels.go:223-396): float32 ops individually rounded,
 * never fused (Go on amd64 default GOAMD64=v1).
 *
 * Build: gcc -O2 -mfma -ffp-contract=off  (contract=off so ONLY the explicit
 * fmaf() calls fuse).
 *
 * All citations are relative to /root/reference.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VGO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* A.1  _mm512_reduce_add_ps as compiled into internal/simd/floats_avx512.s */
/*      :46-53 — i+8, i+4, then (s0+s2)+(s1+s3).                            */
/* ------------------------------------------------------------------------ */
static inline float reduce16(const float v[16]) {
    float t[8], s[4];
    for (int i = 0; i < 8; i++) t[i] = v[i] + v[i + 8];
    for (int i = 0; i < 4; i++) s[i] = t[i] + t[i + 4];
    float r0 = s[0] + s[2];
    float r1 = s[1] + s[3];
    return r0 + r1;
}

/* ------------------------------------------------------------------------ */
/* a1: simd.Dot / simd.SquaredL2, AVX-512 order                             */
/*     internal/simd/src/floats_avx512.c:12-65, 69-129                      */
/* ------------------------------------------------------------------------ */
VGO_API float vgo_dot_a512(const float *a, const float *b, int64_t n) {
    float acc[4][16];
    memset(acc, 0, sizeof acc);
    int64_t epoch = n / 64;
    for (int64_t e = 0; e < epoch; e++)
        for (int j = 0; j < 4; j++)
            for (int l = 0; l < 16; l++) {
                int64_t d = e * 64 + j * 16 + l;
                acc[j][l] = fmaf(a[d], b[d], acc[j][l]);
            }
    float c[16];
    for (int l = 0; l < 16; l++) {
        float s12 = acc[0][l] + acc[1][l];
        float s34 = acc[2][l] + acc[3][l];
        c[l] = s12 + s34;
    }
    float total = reduce16(c);
    for (int64_t i = epoch * 64; i < n; i++) total = fmaf(a[i], b[i], total);
    return total;
}

VGO_API float vgo_sql2_a512(const float *a, const float *b, int64_t n) {
    float acc[4][16];
    memset(acc, 0, sizeof acc);
    int64_t epoch = n / 64;
    for (int64_t e = 0; e < epoch; e++)
        for (int j = 0; j < 4; j++)
            for (int l = 0; l < 16; l++) {
                int64_t d = e * 64 + j * 16 + l;
                float df = a[d] - b[d];
                acc[j][l] = fmaf(df, df, acc[j][l]);
            }
    float c[16];
    for (int l = 0; l < 16; l++) {
        float s12 = acc[0][l] + acc[1][l];
        float s34 = acc[2][l] + acc[3][l];
        c[l] = s12 + s34;
    }
    float total = reduce16(c);
    for (int64_t i = epoch * 64; i < n; i++) {
        float df = a[i] - b[i];
        total = fmaf(df, df, total);
    }
    return total;
}

/* a2: simd.DotBatch / simd.SquaredL2Batch — internal/simd/src/batch_avx512.c:19-143.
 * Differs from the pair kernel for 16 <= dim%64: the 16-blocks accumulate into
 * the already combined accumulator before the reduce. */
static float batch_one_a512(const float *q, const float *t, int64_t dim, int is_l2) {
    float acc[4][16];
    memset(acc, 0, sizeof acc);
    int64_t j = 0;
    for (; j + 64 <= dim; j += 64)
        for (int a = 0; a < 4; a++)
            for (int l = 0; l < 16; l++) {
                int64_t d = j + a * 16 + l;
                if (is_l2) {
                    float df = q[d] - t[d];
                    acc[a][l] = fmaf(df, df, acc[a][l]);
                } else {
                    acc[a][l] = fmaf(q[d], t[d], acc[a][l]);
                }
            }
    float c[16];
    for (int l = 0; l < 16; l++) {
        float s12 = acc[0][l] + acc[1][l];
        float s34 = acc[2][l] + acc[3][l];
        c[l] = s12 + s34;
    }
    for (; j + 16 <= dim; j += 16)
        for (int l = 0; l < 16; l++) {
            if (is_l2) {
                float df = q[j + l] - t[j + l];
                c[l] = fmaf(df, df, c[l]);
            } else {
                c[l] = fmaf(q[j + l], t[j + l], c[l]);
            }
        }
    float total = reduce16(c);
    for (; j < dim; j++) {
        if (is_l2) {
            float df = q[j] - t[j];
            total = fmaf(df, df, total);
        } else {
            total = fmaf(q[j], t[j], total);
        }
    }
    return total;
}

VGO_API void vgo_sql2_batch_a512(const float *q, const float *targets, int64_t dim, int64_t n, float *out) {
    for (int64_t i = 0; i < n; i++) out[i] = batch_one_a512(q, targets + i * dim, dim, 1);
}
VGO_API void vgo_dot_batch_a512(const float *q, const float *targets, int64_t dim, int64_t n, float *out) {
    for (int64_t i = 0; i < n; i++) out[i] = batch_one_a512(q, targets + i * dim, dim, 0);
}

/* a3: simd.ScaleInPlace (floats_avx512.c:174-217) + distance.NormalizeL2InPlace
 * (distance/distance.go:42-53); simd.Sqrt goes through float64 (simd/doc.go:58-60). */
VGO_API void vgo_scale(float *a, int64_t n, float s) {
    for (int64_t i = 0; i < n; i++) a[i] = a[i] * s;
}
VGO_API float vgo_sqrt(float x) { return (float)sqrt((double)x); }
VGO_API int vgo_normalize_l2(float *v, int64_t n) {
    if (n == 0) return 0;
    float norm2 = vgo_dot_a512(v, v, n);
    if (norm2 == 0) return 0;
    float inv = 1.0f / vgo_sqrt(norm2);
    vgo_scale(v, n, inv);
    return 1;
}

/* ------------------------------------------------------------------------ */
/* Generic (pure-Go) kernels — internal/simd/kernels.go:223-396. Unfused.   */
/* ------------------------------------------------------------------------ */
VGO_API float vgo_dot_generic(const float *a, const float *b, int64_t n) {
    float r = 0;
    for (int64_t i = 0; i < n; i++) {
        float p = a[i] * b[i];
        r = r + p;
    }
    return r;
}
VGO_API float vgo_sql2_generic(const float *a, const float *b, int64_t n) {
    float r = 0;
    for (int64_t i = 0; i < n; i++) {
        float d = a[i] - b[i];
        float p = d * d;
        r = r + p;
    }
    return r;
}
VGO_API float vgo_pq_adc_generic(const float *table, const uint8_t *codes, int64_t m) {
    float s = 0;
    for (int64_t i = 0; i < m; i++) s = s + table[i * 256 + codes[i]];
    return s;
}
VGO_API int64_t vgo_hamming(const uint8_t *a, const uint8_t *b, int64_t n) {
    /* kernels.go:278-290; identical integer result on every ISA
     * (src/popcount_avx512.c:25-46). */
    int64_t total = 0;
    for (int64_t i = 0; i < n; i++) total += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return total;
}
VGO_API void vgo_sq8u_l2_batch_generic(const float *q, const uint8_t *codes, const float *mins,
                                       const float *inv, int64_t dim, int64_t n, float *out) {
    /* kernels.go:292-304 */
    for (int64_t i = 0; i < n; i++) {
        float sum = 0;
        for (int64_t d = 0; d < dim; d++) {
            float c = (float)codes[i * dim + d];
            float p = c * inv[d];
            float deq = mins[d] + p;
            float df = q[d] - deq;
            float sq = df * df;
            sum = sum + sq;
        }
        out[i] = sum;
    }
}
VGO_API float vgo_int4_l2_generic(const float *q, const uint8_t *code, int64_t dim, const float *minv,
                                  const float *diff) {
    /* kernels.go:306-324: float32(quant)/15.0*diff+min, val - query. */
    float sum = 0;
    for (int64_t i = 0; i < dim; i += 2) {
        uint8_t b = code[i / 2];
        float n1 = (float)((b >> 4) & 0x0F) / 15.0f;
        float v1 = n1 * diff[i];
        v1 = v1 + minv[i];
        float d1 = v1 - q[i];
        float s1 = d1 * d1;
        sum = sum + s1;
        if (i + 1 < dim) {
            float n2 = (float)(b & 0x0F) / 15.0f;
            float v2 = n2 * diff[i + 1];
            v2 = v2 + minv[i + 1];
            float d2 = v2 - q[i + 1];
            float s2 = d2 * d2;
            sum = sum + s2;
        }
    }
    return sum;
}
/* simd.BuildInt4LookupTable — kernels.go:96-105 */
VGO_API void vgo_int4_build_lut(const float *minv, const float *diff, int64_t dim, float *table) {
    for (int64_t d = 0; d < dim; d++)
        for (int q = 0; q < 16; q++) {
            float n = (float)q / 15.0f;
            float v = n * diff[d];
            table[d * 16 + q] = v + minv[d];
        }
}
/* kernels.go:326-345 (generic precomputed-LUT distance) */
VGO_API float vgo_int4_l2_precomputed_generic(const float *q, const uint8_t *code, int64_t dim, const float *lut) {
    float sum = 0;
    for (int64_t i = 0; i < dim; i += 2) {
        uint8_t b = code[i / 2];
        float d1 = lut[i * 16 + ((b >> 4) & 0x0F)] - q[i];
        float s1 = d1 * d1;
        sum = sum + s1;
        if (i + 1 < dim) {
            float d2 = lut[(i + 1) * 16 + (b & 0x0F)] - q[i + 1];
            float s2 = d2 * d2;
            sum = sum + s2;
        }
    }
    return sum;
}
/* The LIVE int8-PQ helpers (asm variants exist but are never registered,
 * kernels_amd64.go:46-71): kernels.go:354-396. */
VGO_API float vgo_sql2_int8_dequant(const float *q, const int8_t *code, int64_t n, float scale, float offset) {
    float sum = 0;
    for (int64_t i = 0; i < n; i++) {
        float p = (float)code[i] * scale;
        float v = p + offset;
        float d = q[i] - v;
        float s = d * d;
        sum = sum + s;
    }
    return sum;
}
VGO_API void vgo_build_distance_table_int8(const float *qsub, const int8_t *codebook, int64_t subdim, float scale,
                                           float offset, int64_t k, float *out) {
    if (subdim <= 0) return;
    for (int64_t c = 0; c < k; c++) out[c] = vgo_sql2_int8_dequant(qsub, codebook + c * subdim, subdim, scale, offset);
}
VGO_API int64_t vgo_find_nearest_centroid_int8(const float *qsub, const int8_t *codebook, int64_t subdim, int64_t k,
                                               float scale, float offset) {
    if (subdim <= 0 || k <= 0) return 0;
    int64_t best = 0;
    float bd = vgo_sql2_int8_dequant(qsub, codebook, subdim, scale, offset);
    for (int64_t c = 1; c < k; c++) {
        float d = vgo_sql2_int8_dequant(qsub, codebook + c * subdim, subdim, scale, offset);
        if (d < bd) {
            bd = d;
            best = c;
        }
    }
    return best;
}

/* ------------------------------------------------------------------------ */
/* a8: simd.Sq8uL2BatchPerDimension, AVX-512 order — src/sq8_avx512.c:59-104 */
/* ------------------------------------------------------------------------ */
VGO_API void vgo_sq8u_l2_batch_a512(const float *q, const uint8_t *codes, const float *mins, const float *inv,
                                    int64_t dim, int64_t n, float *out) {
    for (int64_t i = 0; i < n; i++) {
        const uint8_t *code = codes + i * dim;
        float s[16];
        memset(s, 0, sizeof s);
        int64_t j = 0;
        for (; j + 16 <= dim; j += 16)
            for (int l = 0; l < 16; l++) {
                float rec = fmaf((float)code[j + l], inv[j + l], mins[j + l]);
                float df = q[j + l] - rec;
                s[l] = fmaf(df, df, s[l]);
            }
        float total = reduce16(s);
        for (; j < dim; j++) {
            float rec = fmaf((float)code[j], inv[j], mins[j]);
            float df = q[j] - rec;
            total = fmaf(df, df, total);
        }
        out[i] = total;
    }
}

/* ------------------------------------------------------------------------ */
/* a9: simd.Int4L2DistanceBatch, AVX-512 order — src/int4_avx512.c:193-299.  */
/* Byte t holds dim 2t in the HIGH nibble and dim 2t+1 in the LOW nibble.    */
/* ------------------------------------------------------------------------ */
static inline float inv15(void) {
    union { uint32_t u; float f; } c;
    c.u = 0x3d888889u;
    return c.f;
}
static inline float int4_nib(const uint8_t *code, int64_t d) {
    uint8_t b = code[d / 2];
    return (float)((d & 1) ? (b & 0x0F) : ((b >> 4) & 0x0F));
}
VGO_API float vgo_int4_l2_a512(const float *q, const uint8_t *code, int64_t dim, const float *minv, const float *diff) {
    const float k = inv15();
    float s1[16], s2[16];
    memset(s1, 0, sizeof s1);
    memset(s2, 0, sizeof s2);
    int64_t i = 0;
    for (; i + 64 <= dim; i += 64)
        for (int blk = 0; blk < 4; blk++) {
            float *acc = (blk < 2) ? s1 : s2;
            for (int l = 0; l < 16; l++) {
                int64_t d = i + blk * 16 + l;
                float g = int4_nib(code, d) * k;
                float deq = fmaf(g, diff[d], minv[d]);
                float e = q[d] - deq;
                acc[l] = fmaf(e, e, acc[l]);
            }
        }
    for (; i + 32 <= dim; i += 32)
        for (int blk = 0; blk < 2; blk++)
            for (int l = 0; l < 16; l++) {
                int64_t d = i + blk * 16 + l;
                float g = int4_nib(code, d) * k;
                float deq = fmaf(g, diff[d], minv[d]);
                float e = q[d] - deq;
                s1[l] = fmaf(e, e, s1[l]);
            }
    float c[16];
    for (int l = 0; l < 16; l++) c[l] = s1[l] + s2[l];
    float total = reduce16(c);
    for (; i < dim; i++) {
        float g = int4_nib(code, i) * k;
        float deq = fmaf(g, diff[i], minv[i]);
        float e = q[i] - deq;
        total = fmaf(e, e, total);
    }
    return total;
}
VGO_API void vgo_int4_l2_batch_a512(const float *q, const uint8_t *codes, int64_t dim, int64_t n, const float *minv,
                                    const float *diff, float *out) {
    int64_t cs = (dim + 1) / 2;
    for (int64_t j = 0; j < n; j++) out[j] = vgo_int4_l2_a512(q, codes + j * cs, dim, minv, diff);
}
/* src/int4_avx512.c:141-189: single accumulator, 16 dims per step, LUT values */
VGO_API float vgo_int4_l2_precomputed_a512(const float *q, const uint8_t *code, int64_t dim, const float *lut) {
    float s[16];
    memset(s, 0, sizeof s);
    int64_t i = 0;
    for (; i + 16 <= dim; i += 16)
        for (int l = 0; l < 16; l++) {
            int64_t d = i + l;
            uint8_t b = code[d / 2];
            int nib = (d & 1) ? (b & 0x0F) : ((b >> 4) & 0x0F);
            float e = q[d] - lut[d * 16 + nib];
            s[l] = fmaf(e, e, s[l]);
        }
    float total = reduce16(s);
    for (; i < dim; i++) {
        uint8_t b = code[i / 2];
        int nib = (i & 1) ? (b & 0x0F) : ((b >> 4) & 0x0F);
        float e = q[i] - lut[i * 16 + nib];
        total = fmaf(e, e, total);
    }
    return total;
}

/* ------------------------------------------------------------------------ */
/* f4: simd.SquaredL2Bounded                                                 */
/*     AVX-512: internal/simd/src/bounded_l2_avx512.c:19-107 (registered at  */
/*     kernels_amd64.go:269); generic: internal/simd/kernels.go:178-217      */
/* ------------------------------------------------------------------------ */
static inline float hsum512_order(const float v[16]) { /* bounded_l2_avx512.c:6-17 */
    float t[8], u[4];
    for (int i = 0; i < 8; i++) t[i] = v[i] + v[i + 8];
    for (int i = 0; i < 4; i++) u[i] = t[i] + t[i + 4];
    float w0 = u[0] + u[1]; /* vhaddps */
    float w1 = u[2] + u[3];
    return w0 + w1;
}
VGO_API float vgo_squared_l2_bounded_a512(const float *a, const float *b, int64_t n, float bound, int32_t *exceeded) {
    float s[4][16];
    memset(s, 0, sizeof s);
    float total = 0.0f;
    int64_t i = 0;
    while (i + 64 <= n) {
        for (int j = 0; j < 4; j++)
            for (int l = 0; l < 16; l++) {
                float d = a[i + 16 * j + l] - b[i + 16 * j + l];
                s[j][l] = fmaf(d, d, s[j][l]);
            }
        i += 64;
        float c[16];
        for (int l = 0; l < 16; l++) {
            float x = s[0][l] + s[1][l];
            float y = s[2][l] + s[3][l];
            c[l] = x + y;
        }
        total = hsum512_order(c);
        if (total > bound) {
            *exceeded = 1;
            return total;
        }
    }
    {
        float c[16];
        for (int l = 0; l < 16; l++) {
            float x = s[0][l] + s[1][l];
            float y = s[2][l] + s[3][l];
            c[l] = x + y;
        }
        total = hsum512_order(c);
    }
    for (; i + 8 <= n; i += 8) {
        float sq[8], w[4];
        for (int l = 0; l < 8; l++) {
            float d = a[i + l] - b[i + l];
            sq[l] = d * d;
        }
        for (int l = 0; l < 4; l++) w[l] = sq[l] + sq[l + 4];
        float h0 = w[0] + w[1];
        float h1 = w[2] + w[3];
        float h = h0 + h1;
        total = total + h;
    }
    for (; i < n; i++) { /* shipped asm fuses the scalar tail: bounded_l2_avx512.s:94-114 */
        float d = a[i] - b[i];
        total = fmaf(d, d, total);
    }
    *exceeded = (total > bound) ? 1 : 0;
    return total;
}
VGO_API float vgo_squared_l2_bounded_generic(const float *a, const float *b, int64_t n, float bound, int32_t *exceeded) {
    float distance = 0.0f;
    int64_t i = 0;
    for (; i + 64 <= n; i += 64) {
        for (int64_t j = i; j < i + 64; j += 8) {
            float acc = 0.0f;
            for (int u = 0; u < 8; u++) {
                float d = a[j + u] - b[j + u];
                float p = d * d;
                acc = (u == 0) ? p : acc + p; /* d0*d0 + d1*d1 + ... left to right */
            }
            distance = distance + acc;
        }
        if (distance > bound) {
            *exceeded = 1;
            return distance;
        }
    }
    for (; i < n; i++) {
        float d = a[i] - b[i];
        float p = d * d;
        distance = distance + p;
    }
    *exceeded = distance > bound ? 1 : 0;
    return distance;
}

/* ------------------------------------------------------------------------ */
/* a10: simd.PqAdcLookup, AVX-512 order — src/floats_avx512.c:135-167        */
/* (table stride hard-wired to 256).                                         */
/* ------------------------------------------------------------------------ */
VGO_API float vgo_pq_adc_a512(const float *table, const uint8_t *codes, int64_t m) {
    float s[16];
    memset(s, 0, sizeof s);
    int64_t i = 0;
    for (; i + 16 <= m; i += 16)
        for (int l = 0; l < 16; l++) s[l] = s[l] + table[(i + l) * 256 + codes[i + l]];
    float total = reduce16(s);
    for (; i < m; i++) total = total + table[i * 256 + codes[i]];
    return total;
}

/* ------------------------------------------------------------------------ */
/* a7: ScalarQuantizer — internal/quantization/quantizer.go                  */
/* ------------------------------------------------------------------------ */
/* Train :130-180 */
VGO_API int vgo_sq8_train(const float *vecs, int64_t n, int64_t dim, float *mins, float *maxs, float *scales,
                          float *inv) {
    if (n <= 0) return -1;
    for (int64_t i = 0; i < dim; i++) {
        mins[i] = 3.40282346638528859811704183484516925440e+38f;
        maxs[i] = -3.40282346638528859811704183484516925440e+38f;
    }
    for (int64_t r = 0; r < n; r++)
        for (int64_t i = 0; i < dim; i++) {
            float v = vecs[r * dim + i];
            if (v < mins[i]) mins[i] = v;
            if (v > maxs[i]) maxs[i] = v;
        }
    for (int64_t i = 0; i < dim; i++) {
        if (mins[i] == maxs[i]) maxs[i] = mins[i] + 1e-6f;
        float range = maxs[i] - mins[i];
        scales[i] = 255.0f / range;
        inv[i] = range / 255.0f;
    }
    return 0;
}
/* SetBounds :51-75 */
VGO_API void vgo_sq8_set_bounds(const float *mins, const float *maxs, int64_t dim, float *scales, float *inv) {
    for (int64_t i = 0; i < dim; i++) {
        float diff = maxs[i] - mins[i];
        if (diff < 1e-9f) {
            scales[i] = 0;
            inv[i] = 0;
        } else {
            scales[i] = 255.0f / diff;
            inv[i] = diff / 255.0f;
        }
    }
}
/* EncodeInto :200-222 — truncating float32→uint8 */
VGO_API void vgo_sq8_encode(const float *v, int64_t dim, const float *mins, const float *maxs, const float *scales,
                            uint8_t *dst) {
    for (int64_t i = 0; i < dim; i++) {
        float val = v[i];
        if (val < mins[i]) val = mins[i];
        else if (val > maxs[i]) val = maxs[i];
        float t = val - mins[i];
        float nrm = t * scales[i];
        float r = nrm + 0.5f;
        dst[i] = (uint8_t)r;
    }
}
/* DecodeInto :241-248 */
VGO_API void vgo_sq8_decode(const uint8_t *b, int64_t dim, const float *mins, const float *inv, float *dst) {
    for (int64_t i = 0; i < dim; i++) {
        float p = (float)b[i] * inv[i];
        dst[i] = p + mins[i];
    }
}
/* L2Distance :78-91 (scalar Go, unfused) */
VGO_API float vgo_sq8_l2_go(const float *q, const uint8_t *code, int64_t dim, const float *mins, const float *inv) {
    float dist = 0;
    for (int64_t i = 0; i < dim; i++) {
        float p = (float)code[i] * inv[i];
        float val = mins[i] + p;
        float df = q[i] - val;
        float sq = df * df;
        dist = dist + sq;
    }
    return dist;
}
/* DotProduct :109-119 */
VGO_API float vgo_sq8_dot_go(const float *q, const uint8_t *code, int64_t dim, const float *mins, const float *inv) {
    float dot = 0;
    for (int64_t i = 0; i < dim; i++) {
        float p = (float)code[i] * inv[i];
        float val = mins[i] + p;
        float t = q[i] * val;
        dot = dot + t;
    }
    return dot;
}

/* ------------------------------------------------------------------------ */
/* a9: Int4Quantizer — internal/quantization/int4.go                         */
/* ------------------------------------------------------------------------ */
VGO_API int vgo_int4_train(const float *vecs, int64_t n, int64_t dim, float *minv, float *diff) {
    if (n <= 0) return -1;
    float *maxv = (float *)malloc(sizeof(float) * (size_t)dim);
    memcpy(minv, vecs, sizeof(float) * (size_t)dim);
    memcpy(maxv, vecs, sizeof(float) * (size_t)dim);
    for (int64_t r = 1; r < n; r++)
        for (int64_t i = 0; i < dim; i++) {
            float v = vecs[r * dim + i];
            if (v < minv[i]) minv[i] = v;
            if (v > maxv[i]) maxv[i] = v;
        }
    for (int64_t i = 0; i < dim; i++) {
        diff[i] = maxv[i] - minv[i];
        if (diff[i] == 0) diff[i] = 1.0f;
    }
    free(maxv);
    return 0;
}
static inline uint8_t int4_q(float v, float mn, float df) {
    float t = v - mn;
    float norm = t / df;
    if (norm < 0) norm = 0;
    else if (norm > 1) norm = 1;
    return (uint8_t)round((double)norm * 15.0); /* math.Round: half away from zero */
}
/* Encode :68-106 */
VGO_API void vgo_int4_encode(const float *v, int64_t dim, const float *minv, const float *diff, uint8_t *out) {
    for (int64_t i = 0; i < dim; i += 2) {
        uint8_t q1 = int4_q(v[i], minv[i], diff[i]);
        uint8_t q2 = 0;
        if (i + 1 < dim) q2 = int4_q(v[i + 1], minv[i + 1], diff[i + 1]);
        out[i / 2] = (uint8_t)((q1 << 4) | (q2 & 0x0F));
    }
}
/* Decode :109-132 */
VGO_API void vgo_int4_decode(const uint8_t *b, int64_t dim, const float *minv, const float *diff, float *out) {
    for (int64_t i = 0; i < dim; i++) {
        float n = int4_nib(b, i) / 15.0f;
        float v = n * diff[i];
        out[i] = v + minv[i];
    }
}

/* ------------------------------------------------------------------------ */
/* a11: BinaryQuantizer — internal/quantization/binary.go                    */
/* ------------------------------------------------------------------------ */
VGO_API float vgo_bq_train(const float *vecs, int64_t n, int64_t dim) {
    double sum = 0;
    int64_t cnt = 0;
    for (int64_t i = 0; i < n * dim; i++) {
        sum += (double)vecs[i];
        cnt++;
    }
    return cnt > 0 ? (float)(sum / (double)cnt) : 0.0f;
}
/* Encode :87-114 — bit i = (v[i] >= threshold), LSB first, padded to 64-bit words */
VGO_API void vgo_bq_encode(const float *v, int64_t dim, float threshold, uint8_t *out) {
    int64_t nbytes = ((dim + 63) / 64) * 8;
    memset(out, 0, (size_t)nbytes);
    for (int64_t i = 0; i < dim; i++)
        if (v[i] >= threshold) out[i / 8] |= (uint8_t)(1u << (i % 8));
}

/* ------------------------------------------------------------------------ */
/* a12: RaBitQuantizer — internal/quantization/rabitq.go                     */
/* ------------------------------------------------------------------------ */
VGO_API void vgo_rabitq_encode(const float *v, int64_t dim, uint8_t *out) {
    /* :51-78 — sign bits ‖ f32 norm (LE); norm = Sqrt(simd.Dot(v,v)) */
    int64_t nbytes = ((dim + 63) / 64) * 8;
    vgo_bq_encode(v, dim, 0.0f, out);
    float norm = vgo_sqrt(vgo_dot_a512(v, v, dim));
    memcpy(out + nbytes, &norm, 4);
}
VGO_API float vgo_rabitq_qnorm(const float *q, int64_t dim) { return vgo_sqrt(vgo_dot_a512(q, q, dim)); }
/* Distance :119-176 — unfused Go: (qn-yn)^2 + ((4*qn)*yn/float32(D))*hamming */
VGO_API float vgo_rabitq_estimate(float qn, float yn, int64_t dim, int64_t hamming) {
    float h = (float)hamming;
    float t1 = qn - yn;
    float t1sq = t1 * t1;
    float a = 4.0f * qn;
    a = a * yn;
    a = a / (float)dim;
    float t2 = a * h;
    return t1sq + t2;
}
VGO_API float vgo_rabitq_distance(const float *q, int64_t dim, const uint8_t *code) {
    int64_t nbytes = ((dim + 63) / 64) * 8;
    float yn;
    memcpy(&yn, code + nbytes, 4);
    float qn = vgo_rabitq_qnorm(q, dim);
    uint8_t *qb = (uint8_t *)malloc((size_t)nbytes);
    vgo_bq_encode(q, dim, 0.0f, qb);
    int64_t h = vgo_hamming(qb, code, nbytes);
    free(qb);
    return vgo_rabitq_estimate(qn, yn, dim, h);
}

/* ------------------------------------------------------------------------ */
/* a10: ProductQuantizer (given codebooks) — internal/quantization/pq.go     */
/* ------------------------------------------------------------------------ */
/* BuildDistanceTable :452-491 (table stride = K) */
VGO_API void vgo_pq_build_table(const float *query, int64_t dim, int64_t m, int64_t k, const int8_t *codebooks,
                                const float *scales, const float *offsets, float *table) {
    int64_t ds = dim / m;
    for (int64_t s = 0; s < m; s++)
        vgo_build_distance_table_int8(query + s * ds, codebooks + s * k * ds, ds, scales[s], offsets[s], k,
                                      table + s * k);
}
/* Encode :147-182 */
VGO_API void vgo_pq_encode(const float *vec, int64_t dim, int64_t m, int64_t k, const int8_t *codebooks,
                           const float *scales, const float *offsets, uint8_t *codes) {
    int64_t ds = dim / m;
    for (int64_t s = 0; s < m; s++)
        codes[s] = (uint8_t)vgo_find_nearest_centroid_int8(vec + s * ds, codebooks + s * k * ds, ds, k, scales[s],
                                                           offsets[s]);
}
/* Decode :185-229 */
VGO_API void vgo_pq_decode(const uint8_t *codes, int64_t dim, int64_t m, int64_t k, const int8_t *codebooks,
                           const float *scales, const float *offsets, float *out) {
    int64_t ds = dim / m;
    for (int64_t s = 0; s < m; s++) {
        const int8_t *c = codebooks + (s * k + codes[s]) * ds;
        for (int64_t i = 0; i < ds; i++) {
            float p = (float)c[i] * scales[s];
            out[s * ds + i] = p + offsets[s];
        }
    }
}
/* ComputeAsymmetricDistance :234-260 */
VGO_API float vgo_pq_asymmetric(const float *query, const uint8_t *codes, int64_t dim, int64_t m, int64_t k,
                                const int8_t *codebooks, const float *scales, const float *offsets) {
    int64_t ds = dim / m;
    float dist = 0;
    for (int64_t s = 0; s < m; s++) {
        const int8_t *c = codebooks + (s * k + codes[s]) * ds;
        dist = dist + vgo_sql2_int8_dequant(query + s * ds, c, ds, scales[s], offsets[s]);
    }
    return dist;
}
/* Train :98-136 — float32 centroids of ONE subspace → int8 codebook + scale/offset */
VGO_API void vgo_pq_quantize_centroids(const float *cent, int64_t count, int8_t *out, float *scale_out,
                                       float *offset_out) {
    float mn = 3.40282346638528859811704183484516925440e+38f, mx = -mn;
    for (int64_t i = 0; i < count; i++) {
        if (cent[i] < mn) mn = cent[i];
        if (cent[i] > mx) mx = cent[i];
    }
    if (mx == mn) mx = mn + 1e-6f;
    float range = mx - mn;
    float scale = range / 255.0f;
    float s128 = 128.0f * scale;
    float offset = mn + s128;
    for (int64_t i = 0; i < count; i++) {
        float t = cent[i] - mn;
        float r = t / scale;
        int val = (int)round((double)r);
        if (val < 0) val = 0;
        if (val > 255) val = 255;
        out[i] = (int8_t)(val - 128);
    }
    *scale_out = scale;
    *offset_out = offset;
}

/* ------------------------------------------------------------------------ */
/* a15: OptimizedProductQuantizer — internal/quantization/opq.go, svd.go      */
/* Go float32 arithmetic: every operation individually rounded, never fused.  */
/* ------------------------------------------------------------------------ */
/* NewOptimizedProductQuantizer block size :41-58 */
VGO_API int64_t vgo_opq_block_size(int64_t dim, int64_t m) {
    int64_t sub = dim / m, bs = dim;
    if (dim > 64) {
        int64_t best = 1000;
        for (int64_t b = sub; b <= dim; b += sub) {
            if (dim % b == 0) {
                int64_t diff = b > 32 ? b - 32 : 32 - b;
                if (diff < best) {
                    best = diff;
                    bs = b;
                }
            }
        }
    }
    return bs;
}
/* rotateVector :196-214: dst[b*bs+i] = simd.Dot(R_b[i], src_b)  (rot = [blocks][bs][bs]) */
VGO_API void vgo_opq_rotate(const float *src, int64_t dim, int64_t bs, const float *rot, float *dst) {
    for (int64_t b = 0; b < dim / bs; b++)
        for (int64_t i = 0; i < bs; i++) dst[b * bs + i] = vgo_dot_a512(rot + (b * bs + i) * bs, src + b * bs, bs);
}
/* Decode's inverse rotation :244-262: dst[i] = sum_j R[j][i]*src[j], sequential, unfused */
VGO_API void vgo_opq_unrotate(const float *src, int64_t dim, int64_t bs, const float *rot, float *dst) {
    for (int64_t b = 0; b < dim / bs; b++)
        for (int64_t i = 0; i < bs; i++) {
            float sum = 0;
            for (int64_t j = 0; j < bs; j++) {
                float p = rot[(b * bs + j) * bs + i] * src[b * bs + j];
                sum = sum + p;
            }
            dst[b * bs + i] = sum;
        }
}
/* Train step 3 accumulation :139-178: M_b[r][c] += x_b[r] * yhat_b[c], samples in order */
VGO_API void vgo_opq_accumulate_m(const float *x, const float *yhat, int64_t n, int64_t dim, int64_t bs, float *M) {
    int64_t blocks = dim / bs;
    memset(M, 0, sizeof(float) * (size_t)(blocks * bs * bs));
    for (int64_t i = 0; i < n; i++)
        for (int64_t b = 0; b < blocks; b++)
            for (int64_t r = 0; r < bs; r++) {
                float xr = x[i * dim + b * bs + r];
                float *row = M + (b * bs + r) * bs;
                for (int64_t c = 0; c < bs; c++) {
                    float p = xr * yhat[i * dim + b * bs + c];
                    row[c] = row[c] + p;
                }
            }
}
/* svd.go:13-127 one-sided Jacobi; a [n][n] becomes U, v [n][n], sigma [n] */
static void opq_svd(float *u, float *v, float *sigma, int64_t n) {
    const double tol = 1e-5;
    memset(v, 0, sizeof(float) * (size_t)(n * n));
    for (int64_t i = 0; i < n; i++) v[i * n + i] = 1.0f;
    for (int iter = 0; iter < 100; iter++) {
        int changed = 0;
        for (int64_t i = 0; i < n - 1; i++)
            for (int64_t j = i + 1; j < n; j++) {
                float alpha = 0, beta = 0, gamma = 0;
                for (int64_t k = 0; k < n; k++) {
                    float a = u[k * n + i] * u[k * n + i];
                    alpha = alpha + a;
                    float b = u[k * n + j] * u[k * n + j];
                    beta = beta + b;
                    float g = u[k * n + i] * u[k * n + j];
                    gamma = gamma + g;
                }
                if (alpha < 1e-12f || beta < 1e-12f) continue;
                float ab = alpha * beta;
                if (fabs((double)gamma) < tol * sqrt((double)ab)) continue;
                changed = 1;
                float ba = beta - alpha;
                float g2 = 2 * gamma;
                float zeta = ba / g2;
                float zz = zeta * zeta;
                float opz = 1 + zz;
                float rt = (float)sqrt((double)opz);
                float t;
                if (zeta > 0) {
                    float den = zeta + rt;
                    t = 1 / den;
                } else {
                    float den = -zeta + rt;
                    t = -1 / den;
                }
                float tt = t * t;
                float opt = 1 + tt;
                float c = 1 / (float)sqrt((double)opt);
                float sn = c * t;
                for (int64_t k = 0; k < n; k++) {
                    float t1 = u[k * n + i], t2 = u[k * n + j];
                    float a1 = c * t1, a2 = sn * t2, b1 = sn * t1, b2 = c * t2;
                    u[k * n + i] = a1 - a2;
                    u[k * n + j] = b1 + b2;
                }
                for (int64_t k = 0; k < n; k++) {
                    float t1 = v[k * n + i], t2 = v[k * n + j];
                    float a1 = c * t1, a2 = sn * t2, b1 = sn * t1, b2 = c * t2;
                    v[k * n + i] = a1 - a2;
                    v[k * n + j] = b1 + b2;
                }
            }
        if (!changed) break;
    }
    for (int64_t j = 0; j < n; j++) {
        float sum = 0;
        for (int64_t i = 0; i < n; i++) {
            float p = u[i * n + j] * u[i * n + j];
            sum = sum + p;
        }
        sigma[j] = (float)sqrt((double)sum);
        if (sigma[j] > 1e-10f) {
            float inv = 1.0f / sigma[j]; /* Go: `inv := 1.0 / sigma[j]` — untyped constant / float32 is a float32 division */
            for (int64_t i = 0; i < n; i++) u[i * n + j] = u[i * n + j] * inv;
        }
    }
}
/* svd.go:184-216 determinant (partial pivoting; |.| compared in float64) */
static float opq_det(const float *m, int64_t n) {
    float *t = (float *)malloc(sizeof(float) * (size_t)(n * n));
    memcpy(t, m, sizeof(float) * (size_t)(n * n));
    float det = 1.0f;
    for (int64_t i = 0; i < n; i++) {
        int64_t pivot = i;
        for (int64_t j = i + 1; j < n; j++)
            if (fabs((double)t[j * n + i]) > fabs((double)t[pivot * n + i])) pivot = j;
        if (pivot != i) {
            for (int64_t k = 0; k < n; k++) {
                float tmp = t[i * n + k];
                t[i * n + k] = t[pivot * n + k];
                t[pivot * n + k] = tmp;
            }
            det = det * -1;
        }
        if (t[i * n + i] == 0) {
            free(t);
            return 0;
        }
        det = det * t[i * n + i];
        for (int64_t j = i + 1; j < n; j++) {
            float factor = t[j * n + i] / t[i * n + i];
            for (int64_t k = i + 1; k < n; k++) {
                float p = factor * t[i * n + k];
                t[j * n + k] = t[j * n + k] - p;
            }
        }
    }
    free(t);
    return det;
}
/* computeProcrustesRotation svd.go:129-182: R = U V^T, reflection fixed on the smallest singular value.
 * M [n][n] is consumed (becomes U, as in the reference).  Optional outputs u/v/sigma for the SVD property tests. */
VGO_API void vgo_opq_procrustes(float *M, int64_t n, float *R, float *sigma_out, float *v_out) {
    float *v = (float *)malloc(sizeof(float) * (size_t)(n * n));
    float *sigma = (float *)malloc(sizeof(float) * (size_t)n);
    opq_svd(M, v, sigma, n);
    if (sigma_out) memcpy(sigma_out, sigma, sizeof(float) * (size_t)n);
    if (v_out) memcpy(v_out, v, sizeof(float) * (size_t)(n * n));
    int64_t mi = 0;
    float ms = sigma[0];
    for (int64_t i = 1; i < n; i++)
        if (sigma[i] < ms) {
            ms = sigma[i];
            mi = i;
        }
    for (int pass = 0; pass < 2; pass++) {
        for (int64_t i = 0; i < n; i++)
            for (int64_t j = 0; j < n; j++) {
                float sum = 0;
                for (int64_t k = 0; k < n; k++) {
                    float p = M[i * n + k] * v[j * n + k];
                    sum = sum + p;
                }
                R[i * n + j] = sum;
            }
        if (pass == 1) break;
        if (!(opq_det(R, n) < 0)) break;
        for (int64_t i = 0; i < n; i++) M[i * n + mi] = M[i * n + mi] * -1;
    }
    free(v);
    free(sigma);
}

/* ------------------------------------------------------------------------ */
/* Deterministic stand-in for Go's unseeded global math/rand (pq.go:294,     */
/* 308,314,409; kmeans.go:25,131): the reference's training is not           */
/* reproducible even against itself, so both this oracle and the CUDA path   */
/* take an explicit seed and use the same counter-based generator.           */
/* ------------------------------------------------------------------------ */
VGO_API uint64_t vgo_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline uint64_t rng_u64(uint64_t seed, uint64_t a, uint64_t b) {
    return vgo_splitmix64(vgo_splitmix64(seed ^ (a * 0xD6E8FEB86659FD93ull)) ^ b);
}
static inline int64_t rng_intn(uint64_t seed, uint64_t a, uint64_t b, int64_t n) {
    return (int64_t)(rng_u64(seed, a, b) % (uint64_t)n);
}
static inline float rng_f32(uint64_t seed, uint64_t a, uint64_t b) {
    return (float)(rng_u64(seed, a, b) >> 40) * (1.0f / 16777216.0f); /* [0,1) */
}
VGO_API int64_t vgo_rng_intn(uint64_t seed, uint64_t a, uint64_t b, int64_t n) { return rng_intn(seed, a, b, n); }
VGO_API float vgo_rng_f32(uint64_t seed, uint64_t a, uint64_t b) { return rng_f32(seed, a, b); }

/* ------------------------------------------------------------------------ */
/* a14: PQ private k-means — pq.go:275-433. Operates on ONE subspace of      */
/* row-major vectors [n × dim] (columns [start, start+ds)).                  */
/* Distances: simd.SquaredL2 (A.2 order).                                    */
/* ------------------------------------------------------------------------ */
/* initializeCentroids :281-345 (k-means++; RNG streams: tag = subspace) */
VGO_API void vgo_pq_kmeanspp_init(const float *vecs, int64_t n, int64_t dim, int64_t start, int64_t ds, int64_t k,
                                  uint64_t seed, uint64_t tag, float *cent) {
    if (n < k) {
        for (int64_t i = 0; i < k; i++) memcpy(cent + i * ds, vecs + (i % n) * dim + start, sizeof(float) * (size_t)ds);
        return;
    }
    int64_t first = rng_intn(seed, tag, 0, n);
    memcpy(cent, vecs + first * dim + start, sizeof(float) * (size_t)ds);
    float *mind = (float *)malloc(sizeof(float) * (size_t)n);
    float sum = 0;
    for (int64_t i = 0; i < n; i++) {
        float d = vgo_sql2_a512(vecs + i * dim + start, cent, ds);
        mind[i] = d;
        sum = sum + d;
    }
    for (int64_t c = 1; c < k; c++) {
        if (sum == 0) {
            int64_t idx = rng_intn(seed, tag, (uint64_t)c, n);
            memcpy(cent + c * ds, vecs + idx * dim + start, sizeof(float) * (size_t)ds);
            continue;
        }
        float target = rng_f32(seed, tag, (uint64_t)c) * sum;
        float cum = 0;
        int64_t chosen = 0;
        for (int64_t i = 0; i < n; i++) {
            cum = cum + mind[i];
            if (cum >= target) {
                chosen = i;
                break;
            }
        }
        memcpy(cent + c * ds, vecs + chosen * dim + start, sizeof(float) * (size_t)ds);
        sum = 0;
        for (int64_t i = 0; i < n; i++) {
            float d = vgo_sql2_a512(vecs + i * dim + start, cent + c * ds, ds);
            if (d < mind[i]) mind[i] = d;
            sum = sum + mind[i];
        }
    }
    free(mind);
}
/* findNearestCentroid :416-433 — strict <, from MaxFloat32, first wins */
VGO_API int64_t vgo_pq_find_nearest(const float *sub, const float *cent, int64_t k, int64_t ds) {
    float md = 3.40282346638528859811704183484516925440e+38f;
    int64_t best = 0;
    for (int64_t c = 0; c < k; c++) {
        float d = vgo_sql2_a512(sub, cent + c * ds, ds);
        if (d < md) {
            md = d;
            best = c;
        }
    }
    return best;
}
/* runKMeansIterations :347-433.  assignments start at 0; loop stops when an
 * assignment pass changes nothing.  Empty cluster → re-seed from the
 * deterministic RNG (stream tag, iteration, cluster).  Returns iterations
 * that executed an update step. */
VGO_API int64_t vgo_pq_lloyd(const float *vecs, int64_t n, int64_t dim, int64_t start, int64_t ds, int64_t k,
                             int64_t max_iters, uint64_t seed, uint64_t tag, float *cent, int32_t *assign) {
    float *sums = (float *)malloc(sizeof(float) * (size_t)(k * ds));
    int64_t *counts = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
    memset(assign, 0, sizeof(int32_t) * (size_t)n);
    int64_t it = 0;
    for (; it < max_iters; it++) {
        int changed = 0;
        for (int64_t i = 0; i < n; i++) {
            int32_t c = (int32_t)vgo_pq_find_nearest(vecs + i * dim + start, cent, k, ds);
            if (assign[i] != c) {
                assign[i] = c;
                changed = 1;
            }
        }
        if (!changed) break;
        memset(sums, 0, sizeof(float) * (size_t)(k * ds));
        memset(counts, 0, sizeof(int64_t) * (size_t)k);
        for (int64_t i = 0; i < n; i++) {
            int32_t c = assign[i];
            counts[c]++;
            for (int64_t j = 0; j < ds; j++) sums[c * ds + j] = sums[c * ds + j] + vecs[i * dim + start + j];
        }
        for (int64_t c = 0; c < k; c++) {
            if (counts[c] > 0) {
                for (int64_t j = 0; j < ds; j++) cent[c * ds + j] = sums[c * ds + j] / (float)counts[c];
            } else {
                int64_t idx = rng_intn(seed, tag ^ 0xE0E0E0E0ull, (uint64_t)(it * k + c), n);
                memcpy(cent + c * ds, vecs + idx * dim + start, sizeof(float) * (size_t)ds);
            }
        }
    }
    free(sums);
    free(counts);
    return it;
}

/* ------------------------------------------------------------------------ */
/* a13: internal/kmeans/kmeans.go                                            */
/* ------------------------------------------------------------------------ */
/* AssignPartition :142-196 (metric: 0 L2, 1 cosine, 2 dot) */
VGO_API int64_t vgo_kmeans_assign(const float *vec, const float *cent, int64_t dim, int64_t k, int metric) {
    float *d = (float *)malloc(sizeof(float) * (size_t)k);
    int64_t best = 0;
    if (metric == 0) {
        vgo_sql2_batch_a512(vec, cent, dim, k, d);
        float m = d[0];
        for (int64_t j = 1; j < k; j++)
            if (d[j] < m) {
                m = d[j];
                best = j;
            }
    } else {
        vgo_dot_batch_a512(vec, cent, dim, k, d);
        float m = d[0];
        for (int64_t j = 1; j < k; j++)
            if (d[j] > m) {
                m = d[j];
                best = j;
            }
    }
    free(d);
    return best;
}
/* TrainKMeans :16-138 — init = caller-provided permutation prefix (Go: rand.Perm) */
VGO_API int64_t vgo_kmeans_train(const float *vecs, int64_t n, int64_t dim, int64_t k, int metric, int64_t max_iter,
                                 const int64_t *init_rows, uint64_t seed, float *cent, int32_t *assign) {
    if (n < k) return -1;
    for (int64_t i = 0; i < k; i++) memcpy(cent + i * dim, vecs + init_rows[i] * dim, sizeof(float) * (size_t)dim);
    float *sums = (float *)malloc(sizeof(float) * (size_t)(k * dim));
    int64_t *counts = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
    memset(assign, 0, sizeof(int32_t) * (size_t)n);
    int64_t it = 0;
    for (; it < max_iter; it++) {
        int changed = 0;
        for (int64_t i = 0; i < n; i++) {
            int32_t c = (int32_t)vgo_kmeans_assign(vecs + i * dim, cent, dim, k, metric);
            if (assign[i] != c) {
                assign[i] = c;
                changed = 1;
            }
        }
        if (!changed) break;
        memset(sums, 0, sizeof(float) * (size_t)(k * dim));
        memset(counts, 0, sizeof(int64_t) * (size_t)k);
        for (int64_t i = 0; i < n; i++) {
            int32_t c = assign[i];
            for (int64_t d = 0; d < dim; d++) sums[c * dim + d] = sums[c * dim + d] + vecs[i * dim + d];
            counts[c]++;
        }
        for (int64_t j = 0; j < k; j++) {
            if (counts[j] > 0) {
                float scale = 1.0f / (float)counts[j];
                for (int64_t d = 0; d < dim; d++) cent[j * dim + d] = sums[j * dim + d] * scale;
            } else {
                int64_t idx = rng_intn(seed, 0xE0E0E0E0ull, (uint64_t)(it * k + j), n);
                memcpy(cent + j * dim, vecs + idx * dim, sizeof(float) * (size_t)dim);
            }
        }
    }
    free(sums);
    free(counts);
    return it;
}
/* FindClosestCentroids :217-280.  Selection for (n <= k/4 && n < 16), else a
 * full sort by distance.  Go's slices.SortFunc is unstable pdqsort; ties in
 * f32 centroid distance are measure-zero on real data, and we define the
 * oracle (and the CUDA path) as stable-by-id on ties. */
typedef struct {
    int64_t id;
    float dist;
} cdist_t;
static int cdist_cmp(const void *a, const void *b) {
    const cdist_t *x = a, *y = b;
    if (x->dist < y->dist) return -1;
    if (x->dist > y->dist) return 1;
    return (x->id > y->id) - (x->id < y->id);
}
VGO_API int64_t vgo_find_closest_centroids(const float *q, const float *cent, int64_t dim, int64_t k, int64_t n,
                                           int metric, int64_t *out) {
    if (n > k) n = k;
    cdist_t *d = (cdist_t *)malloc(sizeof(cdist_t) * (size_t)k);
    float *v = (float *)malloc(sizeof(float) * (size_t)k);
    if (metric == 0) vgo_sql2_batch_a512(q, cent, dim, k, v);
    else vgo_dot_batch_a512(q, cent, dim, k, v);
    for (int64_t i = 0; i < k; i++) {
        d[i].id = i;
        d[i].dist = (metric == 0) ? v[i] : -v[i];
    }
    if (n <= k / 4 && n < 16) {
        for (int64_t i = 0; i < n; i++) {
            int64_t mi = i;
            for (int64_t j = i + 1; j < k; j++)
                if (d[j].dist < d[mi].dist) mi = j;
            cdist_t t = d[i];
            d[i] = d[mi];
            d[mi] = t;
            out[i] = d[i].id;
        }
    } else {
        qsort(d, (size_t)k, sizeof(cdist_t), cdist_cmp);
        for (int64_t i = 0; i < n; i++) out[i] = d[i].id;
    }
    free(d);
    free(v);
    return n;
}

/* ------------------------------------------------------------------------ */
/* a5: searcher.CandidateHeap — internal/searcher/candidate_queue.go         */
/* 4-ary worst-on-top heap, total order (score, SegmentID, RowID).           */
/* ------------------------------------------------------------------------ */
typedef struct {
    uint32_t seg;
    uint32_t row;
    float score;
    uint32_t approx;
} vgo_cand_t;

static inline int cand_better(vgo_cand_t a, vgo_cand_t b, int desc) { /* :12-23 */
    if (a.score != b.score) return desc ? (a.score > b.score) : (a.score < b.score);
    if (a.seg != b.seg) return a.seg < b.seg;
    return a.row < b.row;
}
static inline int cand_worse(vgo_cand_t a, vgo_cand_t b, int desc) { /* :27-38 */
    if (a.score != b.score) return desc ? (a.score < b.score) : (a.score > b.score);
    if (a.seg != b.seg) return a.seg > b.seg;
    return a.row > b.row;
}
typedef struct {
    vgo_cand_t *c;
    int64_t len;
    int desc;
} heap_t;
static void heap_up(heap_t *h, int64_t j) { /* :138-149 */
    vgo_cand_t item = h->c[j];
    while (j > 0) {
        int64_t i = (j - 1) / 4;
        if (!cand_worse(item, h->c[i], h->desc)) break;
        h->c[j] = h->c[i];
        j = i;
    }
    h->c[j] = item;
}
static void heap_down(heap_t *h, int64_t i0, int64_t n) { /* :153-179 */
    int64_t i = i0;
    vgo_cand_t item = h->c[i];
    for (;;) {
        int64_t fc = 4 * i + 1;
        if (fc >= n) break;
        int64_t best = fc, lc = fc + 4;
        if (lc > n) lc = n;
        for (int64_t c = fc + 1; c < lc; c++)
            if (cand_worse(h->c[c], h->c[best], h->desc)) best = c;
        if (!cand_worse(h->c[best], item, h->desc)) break;
        h->c[i] = h->c[best];
        i = best;
    }
    h->c[i] = item;
}
static inline void heap_offer(heap_t *h, vgo_cand_t x, int64_t k) { /* TryPushBounded :120-134 */
    if (h->len < k) {
        h->c[h->len++] = x;
        heap_up(h, h->len - 1);
    } else if (k > 0 && cand_better(x, h->c[0], h->desc)) {
        h->c[0] = x;
        heap_down(h, 0, h->len);
    }
}
static int cand_sort_cmp_asc(const void *a, const void *b) {
    vgo_cand_t x = *(const vgo_cand_t *)a, y = *(const vgo_cand_t *)b;
    if (cand_better(x, y, 0)) return -1;
    if (cand_better(y, x, 0)) return 1;
    return 0;
}
static int cand_sort_cmp_desc(const void *a, const void *b) {
    vgo_cand_t x = *(const vgo_cand_t *)a, y = *(const vgo_cand_t *)b;
    if (cand_better(x, y, 1)) return -1;
    if (cand_better(y, x, 1)) return 1;
    return 0;
}
/* Test hook: push a stream through the heap, return best-first (SortedResults :190-213). */
VGO_API int64_t vgo_heap_topk(const vgo_cand_t *in, int64_t n, int64_t k, int desc, vgo_cand_t *out) {
    heap_t h = {out, 0, desc};
    for (int64_t i = 0; i < n; i++) heap_offer(&h, in[i], k);
    qsort(out, (size_t)h.len, sizeof(vgo_cand_t), desc ? cand_sort_cmp_desc : cand_sort_cmp_asc);
    return h.len;
}

/* ------------------------------------------------------------------------ */
/* a4: flat.(*Segment).Search — internal/segment/flat/segment.go:447-752     */
/* One query, one segment, optional row mask (segment.Filter.Matches) and    */
/* optional IVF partitions.  Kernels are injected as function pointers so    */
/* the same scan loop drives either this file's restatements or the real     */
/* reference C kernels from oracle/_ref.                                     */
/* ------------------------------------------------------------------------ */
typedef void (*sq8_batch_fn)(const float *, const uint8_t *, const float *, const float *, int64_t, int64_t, float *);
typedef float (*pair_fn)(const float *, const float *, int64_t);
typedef float (*adc_fn)(const float *, const uint8_t *, int64_t);
/* calling convention of the reference's C kernels (result via out-pointer;
 * pqAdcLookup takes the offsets[i]=i*256 table, kernels_amd64.go:38-44) */
typedef void (*ref_pair_fn)(float *, float *, int64_t, float *);
typedef void (*ref_adc_fn)(float *, uint8_t *, int64_t, float *, const void *);

typedef struct {
    /* segment */
    int64_t rows, dim;
    uint32_t segment_id;
    int metric;      /* distance.Metric: 0 L2, 1 cosine, 2 dot */
    int quant;       /* flat numbering: 0 none, 1 SQ8, 2 PQ (format.go:22-26) */
    const float *vectors;
    const uint8_t *codes;
    const float *mins, *inv; /* SQ8 */
    int64_t pq_m, pq_k;      /* PQ */
    const int8_t *pq_codebooks;
    const float *pq_scales, *pq_offsets;
    int64_t num_partitions;
    const float *centroids;
    const uint32_t *partition_offsets;
    /* kernels */
    sq8_batch_fn sq8_batch;
    pair_fn l2, dot;
    adc_fn adc;
    /* optional: the real reference kernels from oracle/_ref (used when non-NULL) */
    ref_pair_fn l2_ref, dot_ref;
    ref_adc_fn adc_ref;
    const void *adc_offsets;
} vgo_flat_t;

static inline float call_l2(const vgo_flat_t *s, const float *q, const float *v, int64_t n) {
    if (s->l2_ref) {
        float r;
        s->l2_ref((float *)q, (float *)v, n, &r);
        return r;
    }
    return s->l2(q, v, n);
}
static inline float call_dot(const vgo_flat_t *s, const float *q, const float *v, int64_t n) {
    if (s->dot_ref) {
        float r;
        s->dot_ref((float *)q, (float *)v, n, &r);
        return r;
    }
    return s->dot(q, v, n);
}
static inline float call_adc(const vgo_flat_t *s, const float *t, const uint8_t *c, int64_t m) {
    if (s->adc_ref) {
        float r;
        s->adc_ref((float *)t, (uint8_t *)c, m, &r, s->adc_offsets);
        return r;
    }
    return s->adc(t, c, m);
}

static void flat_scan(const vgo_flat_t *s, const float *q, int64_t k, const uint8_t *mask, const float *pq_table,
                      int64_t start, int64_t end, heap_t *h) {
    int64_t dim = s->dim;
    int desc = s->metric != 0;
    if (s->quant == 1 && s->metric == 0) { /* :518-604: batches of 256 */
        float scores[256];
        for (int64_t i = start; i < end;) {
            int64_t limit = i + 256 < end ? i + 256 : end;
            int64_t count = limit - i;
            s->sq8_batch(q, s->codes + i * dim, s->mins, s->inv, dim, count, scores);
            for (int64_t j = 0; j < count; j++) {
                int64_t idx = i + j;
                if (mask && !((mask[idx >> 3] >> (idx & 7)) & 1)) continue;
                vgo_cand_t c = {s->segment_id, (uint32_t)idx, scores[j], 1};
                heap_offer(h, c, k);
            }
            i += count;
        }
        return;
    }
    for (int64_t i = start; i < end; i++) { /* :606-724 */
        if (mask && !((mask[i >> 3] >> (i & 7)) & 1)) continue;
        float dist;
        uint32_t approx;
        if (s->quant == 1) {
            dist = (s->metric == 0) ? vgo_sq8_l2_go(q, s->codes + i * dim, dim, s->mins, s->inv)
                                    : vgo_sq8_dot_go(q, s->codes + i * dim, dim, s->mins, s->inv);
            approx = 1;
        } else if (s->quant == 2) {
            dist = call_adc(s, pq_table, s->codes + i * s->pq_m, s->pq_m);
            approx = 1;
        } else {
            dist = (s->metric == 0) ? call_l2(s, q, s->vectors + i * dim, dim) : call_dot(s, q, s->vectors + i * dim, dim);
            approx = 0;
        }
        vgo_cand_t c = {s->segment_id, (uint32_t)i, dist, approx};
        heap_offer(h, c, k);
    }
}

/* Search one query; writes ≤k candidates best-first; returns count. */
VGO_API int64_t vgo_flat_search(const vgo_flat_t *s, const float *q, int64_t k, int64_t nprobes, const uint8_t *mask,
                                vgo_cand_t *out) {
    heap_t h = {out, 0, s->metric != 0};
    float *table = NULL;
    if (s->quant == 2) {
        table = (float *)malloc(sizeof(float) * (size_t)(s->pq_m * s->pq_k));
        vgo_pq_build_table(q, s->dim, s->pq_m, s->pq_k, s->pq_codebooks, s->pq_scales, s->pq_offsets, table);
    }
    if (s->num_partitions > 1) { /* :726-745 */
        if (nprobes <= 0) nprobes = 1;
        int64_t *parts = (int64_t *)malloc(sizeof(int64_t) * (size_t)s->num_partitions);
        int64_t np = vgo_find_closest_centroids(q, s->centroids, s->dim, s->num_partitions, nprobes, s->metric, parts);
        for (int64_t p = 0; p < np; p++)
            flat_scan(s, q, k, mask, table, s->partition_offsets[parts[p]], s->partition_offsets[parts[p] + 1], &h);
        free(parts);
    } else {
        flat_scan(s, q, k, mask, table, 0, s->rows, &h);
    }
    free(table);
    qsort(out, (size_t)h.len, sizeof(vgo_cand_t), h.desc ? cand_sort_cmp_desc : cand_sort_cmp_asc);
    return h.len;
}

/* a6: flat.(*Segment).Rerank — segment.go:754-781: exact score per candidate row,
 * input order preserved, rows >= RowCount skipped (marked by NaN here). */
VGO_API void vgo_flat_rerank(const vgo_flat_t *s, const float *q, const uint32_t *rows, int64_t n, float *scores) {
    for (int64_t i = 0; i < n; i++) {
        if ((int64_t)rows[i] >= s->rows) {
            scores[i] = NAN;
            continue;
        }
        const float *v = s->vectors + (int64_t)rows[i] * s->dim;
        scores[i] = (s->metric == 0) ? call_l2(s, q, v, s->dim) : call_dot(s, q, v, s->dim);
    }
}

/* Engine.BatchSearch (internal/engine/engine.go:1305-1365): one worker per
 * query.  Threads pull queries from a shared counter. */
typedef struct {
    const vgo_flat_t *s;
    const float *queries;
    int64_t nq, k, nprobes;
    const uint8_t *mask;
    vgo_cand_t *out;
    int64_t *counts;
    volatile int64_t *next;
} batch_job_t;
static void *batch_worker(void *arg) {
    batch_job_t *j = (batch_job_t *)arg;
    for (;;) {
        int64_t qi = __sync_fetch_and_add(j->next, 1);
        if (qi >= j->nq) break;
        j->counts[qi] = vgo_flat_search(j->s, j->queries + qi * j->s->dim, j->k, j->nprobes, j->mask, j->out + qi * j->k);
    }
    return NULL;
}
VGO_API void vgo_flat_search_batch(const vgo_flat_t *s, const float *queries, int64_t nq, int64_t k, int64_t nprobes,
                                   const uint8_t *mask, int threads, vgo_cand_t *out, int64_t *counts) {
    volatile int64_t next = 0;
    batch_job_t job = {s, queries, nq, k, nprobes, mask, out, counts, &next};
    if (threads <= 1) {
        batch_worker(&job);
        return;
    }
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; i++) pthread_create(&t[i], NULL, batch_worker, &job);
    for (int i = 0; i < threads; i++) pthread_join(t[i], NULL);
    free(t);
}

/* ------------------------------------------------------------------------ */
/* Generic top-k scans over codes that have no flat-segment caller in the    */
/* reference (INT4, RaBitQ, BQ: DiskANN per-node distances,                  */
/* segment/diskann/segment.go:511-588) — the north star scans them linearly  */
/* with the same heap order.                                                 */
/* ------------------------------------------------------------------------ */
typedef void (*int4_batch_fn)(const float *, const uint8_t *, int64_t, int64_t, const float *, const float *, float *);
typedef long long (*hamming_fn)(const unsigned char *, const unsigned char *, int64_t);

VGO_API int64_t vgo_int4_search(const float *q, const uint8_t *codes, int64_t rows, int64_t dim, const float *minv,
                                const float *diff, int64_t k, int4_batch_fn fn, vgo_cand_t *out) {
    heap_t h = {out, 0, 0};
    float scores[256];
    int64_t cs = (dim + 1) / 2;
    for (int64_t i = 0; i < rows; i += 256) {
        int64_t cnt = rows - i < 256 ? rows - i : 256;
        fn(q, codes + i * cs, dim, cnt, minv, diff, scores);
        for (int64_t j = 0; j < cnt; j++) {
            vgo_cand_t c = {0, (uint32_t)(i + j), scores[j], 1};
            heap_offer(&h, c, k);
        }
    }
    qsort(out, (size_t)h.len, sizeof(vgo_cand_t), cand_sort_cmp_asc);
    return h.len;
}
static long long hamming_default(const unsigned char *a, const unsigned char *b, int64_t n) {
    return vgo_hamming(a, b, n);
}
/* RaBitQ linear scan: rq.Distance per row (rabitq.go:119-176) */
VGO_API int64_t vgo_rabitq_search(const float *q, const uint8_t *codes, int64_t rows, int64_t dim, int64_t k,
                                  hamming_fn hfn, vgo_cand_t *out, int32_t *hamming_out) {
    if (!hfn) hfn = hamming_default;
    heap_t h = {out, 0, 0};
    int64_t nbytes = ((dim + 63) / 64) * 8, stride = nbytes + 4;
    uint8_t *qb = (uint8_t *)malloc((size_t)nbytes);
    vgo_bq_encode(q, dim, 0.0f, qb);
    float qn = vgo_rabitq_qnorm(q, dim);
    for (int64_t i = 0; i < rows; i++) {
        const uint8_t *code = codes + i * stride;
        float yn;
        memcpy(&yn, code + nbytes, 4);
        int64_t hm = hfn(qb, code, nbytes);
        if (hamming_out) hamming_out[i] = (int32_t)hm;
        vgo_cand_t c = {0, (uint32_t)i, vgo_rabitq_estimate(qn, yn, dim, hm), 1};
        heap_offer(&h, c, k);
    }
    free(qb);
    qsort(out, (size_t)h.len, sizeof(vgo_cand_t), cand_sort_cmp_asc);
    return h.len;
}
/* BQ linear scan: Hamming distance as the score (distance.Hamming → float32) */
VGO_API int64_t vgo_bq_search(const uint8_t *qcode, const uint8_t *codes, int64_t rows, int64_t nbytes, int64_t k,
                              hamming_fn hfn, vgo_cand_t *out) {
    if (!hfn) hfn = hamming_default;
    heap_t h = {out, 0, 0};
    for (int64_t i = 0; i < rows; i++) {
        vgo_cand_t c = {0, (uint32_t)i, (float)hfn(qcode, codes + i * nbytes, nbytes), 1};
        heap_offer(&h, c, k);
    }
    qsort(out, (size_t)h.len, sizeof(vgo_cand_t), cand_sort_cmp_asc);
    return h.len;
}

/* Threaded wrappers for the CPU-baseline arm (one query per worker). */
typedef struct {
    int kind; /* 0 int4, 1 rabitq */
    const float *queries;
    int64_t nq, rows, dim, k;
    const uint8_t *codes;
    const float *minv, *diff;
    int4_batch_fn i4;
    hamming_fn hf;
    vgo_cand_t *out;
    int64_t *counts;
    volatile int64_t *next;
} scan_job_t;
static void *scan_worker(void *arg) {
    scan_job_t *j = (scan_job_t *)arg;
    for (;;) {
        int64_t qi = __sync_fetch_and_add(j->next, 1);
        if (qi >= j->nq) break;
        const float *q = j->queries + qi * j->dim;
        if (j->kind == 0)
            j->counts[qi] = vgo_int4_search(q, j->codes, j->rows, j->dim, j->minv, j->diff, j->k, j->i4, j->out + qi * j->k);
        else
            j->counts[qi] = vgo_rabitq_search(q, j->codes, j->rows, j->dim, j->k, j->hf, j->out + qi * j->k, NULL);
    }
    return NULL;
}
static void run_scan_job(scan_job_t *job, int threads) {
    if (threads <= 1) {
        scan_worker(job);
        return;
    }
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 0; i < threads; i++) pthread_create(&t[i], NULL, scan_worker, job);
    for (int i = 0; i < threads; i++) pthread_join(t[i], NULL);
    free(t);
}
VGO_API void vgo_int4_search_batch(const float *queries, int64_t nq, const uint8_t *codes, int64_t rows, int64_t dim,
                                   const float *minv, const float *diff, int64_t k, int4_batch_fn fn, int threads,
                                   vgo_cand_t *out, int64_t *counts) {
    volatile int64_t next = 0;
    scan_job_t job = {0, queries, nq, rows, dim, k, codes, minv, diff, fn, NULL, out, counts, &next};
    run_scan_job(&job, threads);
}
VGO_API void vgo_rabitq_search_batch(const float *queries, int64_t nq, const uint8_t *codes, int64_t rows, int64_t dim,
                                     int64_t k, hamming_fn hfn, int threads, vgo_cand_t *out, int64_t *counts) {
    volatile int64_t next = 0;
    scan_job_t job = {1, queries, nq, rows, dim, k, codes, NULL, NULL, NULL, hfn, out, counts, &next};
    run_scan_job(&job, threads);
}

/* ------------------------------------------------------------------------ */
/* CRC32C (Castagnoli) — internal/hash: checksum of the flat file body       */
/* (flat/segment.go:170-181).                                                */
/* ------------------------------------------------------------------------ */
VGO_API uint32_t vgo_crc32c(const uint8_t *data, int64_t n) {
    static uint32_t tbl[256];
    static int init = 0;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int j = 0; j < 8; j++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            tbl[i] = c;
        }
        init = 1;
    }
    uint32_t crc = 0xFFFFFFFFu;
    for (int64_t i = 0; i < n; i++) crc = tbl[(crc ^ data[i]) & 0xFF] ^ (crc >> 8);
    return crc ^ 0xFFFFFFFFu;
}
