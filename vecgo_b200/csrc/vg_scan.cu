// vg_scan.cu — the scan kernels: exact Flat L2/dot, SQ8, INT4, PQ-ADC, BQ and
// RaBitQ distance evaluation fused with the shared-memory bounded top-k.
//
// Thread model.  The reference's AVX-512 kernels keep 16 float lanes per
// accumulator and finish with _mm512_reduce_add_ps.  Here a HALF-WARP is one
// zmm register: lane l of the half-warp owns AVX lane l, walks the dimensions
// d = 16*t + l in the same order, and the 16 partial sums are combined with
// reduce16() in the same tree order, so every float32 distance is
// bit-identical to the reference's.  Each half-warp
// register-tiles R rows x QH queries, so a code byte / vector element fetched
// once is reused for all queries of the CTA's query tile.
//
// Grid.  blockIdx.x = query tile (QT queries whose top-k state lives in this
// CTA's shared memory for the whole sweep), blockIdx.y = row split (only when
// there are too few query tiles to fill 148 SMs; partials are merged by
// merge_keys_kernel).  All CTAs sweep rows in the same direction, so code
// tiles are served from the 126 MB L2 after the first CTA touched them.
#include <type_traits>

#include "vg_scan.cuh"

namespace vg {

// ============================================================ sinks
struct SinkTopK {
    TopK tk;
    const uint8_t *mask;
    const int32_t *probe;
    const uint32_t *part_off;
    int nprobe, trigger, q0;
    uint32_t row_base;
    bool descending;
    __device__ __forceinline__ void operator()(int slot, int64_t row, float score) const {
        if (mask && !((mask[row >> 3] >> (row & 7)) & 1)) return;
        if (probe) {
            const int32_t *pp = probe + (int64_t)(q0 + slot) * nprobe;
            bool ok = false;
            for (int j = 0; j < nprobe; j++) {
                const int p = pp[j];
                if (p >= 0 && row >= part_off[p] && row < part_off[p + 1]) ok = true;
            }
            if (!ok) return;
        }
        topk_offer(tk, slot, make_key(score, row_base + (uint32_t)row, descending), trigger);
    }
};
struct SinkDense {
    float *out;
    int64_t n;
    int q0;
    __device__ __forceinline__ void operator()(int slot, int64_t row, float score) const {
        out[(int64_t)(q0 + slot) * n + row] = score;
    }
};

// ============================================================ codec: F32
// simd.SquaredL2 / simd.Dot (floats_avx512.c:12-129) and the Batch variants
// (batch_avx512.c:19-143).
template <bool BATCH>
struct CodecF32 {
    static constexpr int QT = 4, R = 2, THREADS = 256, RB = 16 * R, MINB = 2;
    static size_t qsmem(const CodecParams &P) { return (size_t)QT * P.dim * 4; }
    __device__ static void stage(const CodecParams &P, const float *queries, int64_t qs_, int q0, int nqv, unsigned char *sm, int tid) {
        float *qs = reinterpret_cast<float *>(sm);
        const int64_t dim = P.dim;
        for (int64_t i = tid; i < (int64_t)QT * dim; i += THREADS) {
            int q = (int)(i / dim);
            int64_t d = i - (int64_t)q * dim;
            qs[i] = (q < nqv) ? queries[(int64_t)(q0 + q) * qs_ + d] : 0.0f;
        }
    }
    template <class Sink>
    __device__ static void tile(const CodecParams &P, bool is_dot, const unsigned char *sm, int64_t base, int64_t row_end,
                                int nqv, int tid, const Sink &sink) {
        const float *qs = reinterpret_cast<const float *>(sm);
        const int hw = tid >> 4, lane = tid & 15;
        const int64_t dim = P.dim;
        const int64_t r0 = base + (int64_t)hw * R;
        const float *x[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            int64_t rr = r0 + r;
            if (rr > row_end - 1) rr = row_end - 1;
            x[r] = P.vectors + rr * dim;
        }
        float acc[R][QT][4];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int q = 0; q < QT; q++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[r][q][j] = 0.0f;
        const int64_t epochs = dim >> 6;
        for (int64_t e = 0; e < epochs; e++) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int64_t d = e * 64 + j * 16 + lane;
                float xv[R];
#pragma unroll
                for (int r = 0; r < R; r++) xv[r] = __ldg(x[r] + d);
#pragma unroll
                for (int q = 0; q < QT; q++) {
                    const float qv = qs[(int64_t)q * dim + d];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        if (is_dot) {
                            acc[r][q][j] = __fmaf_rn(qv, xv[r], acc[r][q][j]);
                        } else {
                            const float df = __fsub_rn(qv, xv[r]);
                            acc[r][q][j] = __fmaf_rn(df, df, acc[r][q][j]);
                        }
                    }
                }
            }
        }
        int64_t done = epochs * 64;
        float c[R][QT];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int q = 0; q < QT; q++)
                c[r][q] = __fadd_rn(__fadd_rn(acc[r][q][0], acc[r][q][1]), __fadd_rn(acc[r][q][2], acc[r][q][3]));
        if (BATCH) {
            for (; done + 16 <= dim; done += 16) {
                const int64_t d = done + lane;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const float xv = __ldg(x[r] + d);
#pragma unroll
                    for (int q = 0; q < QT; q++) {
                        const float qv = qs[(int64_t)q * dim + d];
                        if (is_dot) {
                            c[r][q] = __fmaf_rn(qv, xv, c[r][q]);
                        } else {
                            const float df = __fsub_rn(qv, xv);
                            c[r][q] = __fmaf_rn(df, df, c[r][q]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int q = 0; q < QT; q++) c[r][q] = reduce16(c[r][q]);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (r0 + r >= row_end) continue;
#pragma unroll
                for (int q = 0; q < QT; q++) {
                    if (q >= nqv) continue;
                    float tot = c[r][q];
                    for (int64_t d = done; d < dim; d++) {  // scalar tail, FMA-contracted in the shipped asm
                        const float qv = qs[(int64_t)q * dim + d], xv = __ldg(x[r] + d);
                        if (is_dot) {
                            tot = __fmaf_rn(qv, xv, tot);
                        } else {
                            const float df = __fsub_rn(qv, xv);
                            tot = __fmaf_rn(df, df, tot);
                        }
                    }
                    sink(q, r0 + r, tot);
                }
            }
        }
    }
};

// ============================================================ codec: SQ8
// simd.Sq8uL2BatchPerDimension (sq8_avx512.c:59-104): one 16-lane accumulator,
// rec = fma(float(c), invScale, min); diff = q - rec; acc = fma(diff, diff, acc).
// Natural (row-major) device layout; any dim.
struct CodecSQ8 {
    static constexpr int QT = 8, R = 4, THREADS = 256, RB = 16 * R, MINB = 2;
    static size_t qsmem(const CodecParams &P) { return (size_t)(QT + 2) * P.dim * 4; }
    __device__ static void stage(const CodecParams &P, const float *queries, int64_t qs_, int q0, int nqv, unsigned char *sm, int tid) {
        float *qs = reinterpret_cast<float *>(sm);
        const int64_t dim = P.dim;
        for (int64_t i = tid; i < (int64_t)QT * dim; i += THREADS) {
            int q = (int)(i / dim);
            int64_t d = i - (int64_t)q * dim;
            qs[i] = (q < nqv) ? queries[(int64_t)(q0 + q) * qs_ + d] : 0.0f;
        }
        float *mn = qs + (int64_t)QT * dim, *iv = mn + dim;
        for (int64_t d = tid; d < dim; d += THREADS) {
            mn[d] = P.p0[d];
            iv[d] = P.p1[d];
        }
    }
    template <class Sink>
    __device__ static void tile(const CodecParams &P, bool, const unsigned char *sm, int64_t base, int64_t row_end, int nqv,
                                int tid, const Sink &sink) {
        const float *qs = reinterpret_cast<const float *>(sm);
        const int64_t dim = P.dim;
        const float *mn = qs + (int64_t)QT * dim, *iv = mn + dim;
        const int hw = tid >> 4, lane = tid & 15;
        const int64_t r0 = base + (int64_t)hw * R;
        const uint8_t *code[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            int64_t rr = r0 + r;
            if (rr > row_end - 1) rr = row_end - 1;
            code[r] = P.codes + rr * P.row_bytes;
        }
        float acc[R][QT];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int q = 0; q < QT; q++) acc[r][q] = 0.0f;
        int64_t j = 0;
        for (; j + 16 <= dim; j += 16) {
            const int64_t d = j + lane;
            const float m = mn[d], s = iv[d];
            float rec[R];
#pragma unroll
            for (int r = 0; r < R; r++) rec[r] = __fmaf_rn(u8_to_f32(__ldg(code[r] + d)), s, m);
#pragma unroll
            for (int q = 0; q < QT; q++) {
                const float qv = qs[(int64_t)q * dim + d];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const float df = __fsub_rn(qv, rec[r]);
                    acc[r][q] = __fmaf_rn(df, df, acc[r][q]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int q = 0; q < QT; q++) acc[r][q] = reduce16(acc[r][q]);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (r0 + r >= row_end) continue;
#pragma unroll
                for (int q = 0; q < QT; q++) {
                    if (q >= nqv) continue;
                    float tot = acc[r][q];
                    for (int64_t d = j; d < dim; d++) {
                        const float rec = __fmaf_rn(u8_to_f32(__ldg(code[r] + d)), iv[d], mn[d]);
                        const float df = __fsub_rn(qs[(int64_t)q * dim + d], rec);
                        tot = __fmaf_rn(df, df, tot);
                    }
                    sink(q, r0 + r, tot);
                }
            }
        }
    }
};

// ---- packed float32x2 helpers (sm_100a FADD2 / FFMA2).  Each half is an
// IEEE round-to-nearest op, so results are bit-identical to the scalar
// intrinsics; one instruction issue carries two lanes of work.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// Transposed _mm512_reduce_add_ps for NA accumulators per lane of a half-warp
// (NA = 32): the same pairs are added at every level of the reference tree
// (i+8, i+4, i+2, i+1; floats_avx512.s:46-53) — float addition is commutative,
// so only WHERE each partial sum lives changes.  At the level with lane
// distance D the lanes with (lane & D) == 0 keep the lower half of the
// accumulators still alive and their partners keep the upper half, so after
// four levels lane l holds the finished totals of accumulators 2l and 2l+1.
// 30 shuffles + 30 adds instead of 128 + 128, and the 32 results end up spread
// over the 16 lanes, which lets every lane run its own top-k offers.
__device__ __forceinline__ void reduce16_transposed32(float (&v)[32], int lane, float &out0, float &out1) {
    float a16[16], a8[8], a4[4];
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const float send = up ? v[i] : v[16 + i];
            const float keep = up ? v[16 + i] : v[i];
            a16[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 8, 16));
        }
    }
    {
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float send = up ? a16[i] : a16[8 + i];
            const float keep = up ? a16[8 + i] : a16[i];
            a8[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 4, 16));
        }
    }
    {
        const bool up = (lane & 2) != 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float send = up ? a8[i] : a8[4 + i];
            const float keep = up ? a8[4 + i] : a8[i];
            a4[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 2, 16));
        }
    }
    {
        const bool up = (lane & 1) != 0;
        const float s0 = up ? a4[0] : a4[2], k0 = up ? a4[2] : a4[0];
        const float s1 = up ? a4[1] : a4[3], k1 = up ? a4[3] : a4[1];
        out0 = __fadd_rn(k0, __shfl_xor_sync(0xffffffffu, s0, 1, 16));
        out1 = __fadd_rn(k1, __shfl_xor_sync(0xffffffffu, s1, 1, 16));
    }
}

// Same for 16 accumulators per lane: lane l ends with the total of accumulator l.
__device__ __forceinline__ float reduce16_transposed16(float (&v)[16], int lane) {
    float a8[8], a4[4], a2[2];
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float send = up ? v[i] : v[8 + i];
            const float keep = up ? v[8 + i] : v[i];
            a8[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 8, 16));
        }
    }
    {
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float send = up ? a8[i] : a8[4 + i];
            const float keep = up ? a8[4 + i] : a8[i];
            a4[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 4, 16));
        }
    }
    {
        const bool up = (lane & 2) != 0;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float send = up ? a4[i] : a4[2 + i];
            const float keep = up ? a4[2 + i] : a4[i];
            a2[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 2, 16));
        }
    }
    const bool up = (lane & 1) != 0;
    const float send = up ? a2[0] : a2[1], keep = up ? a2[1] : a2[0];
    return __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 1, 16));
}

// SQ8 fast path: dim % (16*VB) == 0 and codes stored lane-transposed on device
// (upload kernel permute_sq8): inside every block of 16*VB dims the byte of
// (step s, lane l) sits at l*VB + s, so lane l fetches VB consecutive steps
// with one VB-byte load and the half-warp's 16 loads cover 16*VB contiguous
// bytes.
//
// Arithmetic (sq8_avx512.c:59-104) per (query, row, dim):
//     rec = fma(c, inv, min); e = q - rec; acc = fma(e, e, acc)
// evaluated here as nrec = fma(c, -inv, -min) (= -rec exactly), e = q + nrec,
// with the e / acc updates of two QUERIES packed into one FADD2 / FFMA2.
// Shared-memory layout: -mins / -invScales lane-major [16][SP]; queries
// [16 lanes][steps][8 queries] with a 16-byte pad per lane so the two
// LDS.128 that fetch the 8 query values of one step are conflict-free.
template <int VB>
struct CodecSQ8Perm {
    static constexpr int QT = 8, R = 4, THREADS = 256, RB = 16 * R, MINB = 2;
    __host__ __device__ static int steps_padded(int64_t dim) {
        int s = (int)(dim / 16);
        int sp = (s + 3) & ~3;
        if (((sp >> 2) & 1) == 0) sp += 4;
        return sp;
    }
    __host__ __device__ static int lane_stride(int64_t dim) { return (int)(dim / 16) * QT + 4; }  // floats
    static size_t qsmem(const CodecParams &P) { return ((size_t)16 * lane_stride(P.dim) + (size_t)2 * 16 * steps_padded(P.dim)) * 4; }
    __device__ static void stage(const CodecParams &P, const float *queries, int64_t qs_, int q0, int nqv, unsigned char *sm, int tid) {
        float *qs = reinterpret_cast<float *>(sm);
        const int64_t dim = P.dim;
        const int SP = steps_padded(dim), LS = lane_stride(dim);
        for (int64_t i = tid; i < (int64_t)QT * dim; i += THREADS) {
            int q = (int)(i / dim);
            int d = (int)(i - (int64_t)q * dim);
            qs[(int64_t)(d & 15) * LS + (d >> 4) * QT + q] = (q < nqv) ? queries[(int64_t)(q0 + q) * qs_ + d] : 0.0f;
        }
        float *mn = qs + (int64_t)16 * LS, *iv = mn + 16 * SP;
        for (int d = tid; d < dim; d += THREADS) {
            mn[(d & 15) * SP + (d >> 4)] = -P.p0[d];
            iv[(d & 15) * SP + (d >> 4)] = -P.p1[d];
        }
    }
    template <class Sink>
    __device__ static void tile(const CodecParams &P, bool, const unsigned char *sm, int64_t base, int64_t row_end, int nqv,
                                int tid, const Sink &sink) {
        const int64_t dim = P.dim;
        const int SP = steps_padded(dim), LS = lane_stride(dim);
        const float *qs = reinterpret_cast<const float *>(sm);
        const float *mn = qs + (int64_t)16 * LS, *iv = mn + 16 * SP;
        const int hw = tid >> 4, lane = tid & 15;
        const int64_t r0 = base + (int64_t)hw * R;
        const uint8_t *code[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            int64_t rr = r0 + r;
            if (rr > row_end - 1) rr = row_end - 1;
            code[r] = P.codes + rr * P.row_bytes + lane * VB;
        }
        f32x2 acc[R][QT / 2];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int q = 0; q < QT / 2; q++) acc[r][q] = 0ull;
        const float *mnl = mn + lane * SP, *ivl = iv + lane * SP;
        const float *ql = qs + (int64_t)lane * LS;
        const int nblk = (int)(dim / (16 * VB));
        for (int b = 0; b < nblk; b++) {
            uint32_t w[R][VB / 4];
#pragma unroll
            for (int r = 0; r < R; r++) {
                if constexpr (VB == 16) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(code[r] + (int64_t)b * 16 * VB));
                    w[r][0] = v.x;
                    w[r][1] = v.y;
                    w[r][2] = v.z;
                    w[r][3] = v.w;
                } else {
                    w[r][0] = __ldg(reinterpret_cast<const uint32_t *>(code[r] + (int64_t)b * 16 * VB));
                }
            }
#pragma unroll
            for (int s4 = 0; s4 < VB / 4; s4++) {
                const int t0 = b * VB + s4 * 4;
                const float4 m4 = *reinterpret_cast<const float4 *>(mnl + t0);
                const float4 i4 = *reinterpret_cast<const float4 *>(ivl + t0);
                const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
                const float ii[4] = {i4.x, i4.y, i4.z, i4.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const ulonglong2 qa = *reinterpret_cast<const ulonglong2 *>(ql + (t0 + i) * QT);      // queries 0..3
                    const ulonglong2 qb = *reinterpret_cast<const ulonglong2 *>(ql + (t0 + i) * QT + 4);  // queries 4..7
                    const f32x2 qq[4] = {qa.x, qa.y, qb.x, qb.y};
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        // byte -> float via 0x4B000000|b = 8388608+b (exact), minus 8388608 (exact)
                        const float f = __fsub_rn(__uint_as_float(__byte_perm(w[r][s4], 0x4B000000u, 0x7650 + i)), 8388608.0f);
                        const float nrec = __fmaf_rn(f, ii[i], mm[i]);
                        const f32x2 n2 = pk2(nrec, nrec);
#pragma unroll
                        for (int q = 0; q < QT / 2; q++) {
                            const f32x2 e = add2(qq[q], n2);
                            acc[r][q] = fma2(e, e, acc[r][q]);
                        }
                    }
                }
            }
        }
        float v[32];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int q = 0; q < QT / 2; q++) unpk2(acc[r][q], v[r * QT + 2 * q], v[r * QT + 2 * q + 1]);
        float o0, o1;
        reduce16_transposed32(v, lane, o0, o1);
        // lane l holds accumulators 2l, 2l+1 = (row l/4, queries 2(l%4), 2(l%4)+1)
        const int64_t row = r0 + (lane >> 2);
        const int qq0 = 2 * (lane & 3);
        if (row < row_end) {
            if (qq0 < nqv) sink(qq0, row, o0);
            if (qq0 + 1 < nqv) sink(qq0 + 1, row, o1);
        }
    }
};

// SQ8 through the scalar Go loops (quantizer.go:78-91,109-119): sequential,
// unfused.  flat.Search uses these for SQ8 segments whose metric is not L2.
// One thread per row (the sum is one dependent chain).
struct CodecSQ8Go {
    static constexpr int QT = 4, R = 1, THREADS = 256, RB = THREADS, MINB = 2;
    static size_t qsmem(const CodecParams &P) { return CodecSQ8::qsmem(P); }  // shares CodecSQ8's 8-slot layout
    __device__ static void stage(const CodecParams &P, const float *queries, int64_t qs_, int q0, int nqv, unsigned char *sm, int tid) {
        CodecSQ8::stage(P, queries, qs_, q0, nqv, sm, tid);
    }
    template <class Sink>
    __device__ static void tile(const CodecParams &P, bool is_dot, const unsigned char *sm, int64_t base, int64_t row_end,
                                int nqv, int tid, const Sink &sink) {
        const float *qs = reinterpret_cast<const float *>(sm);
        const int64_t dim = P.dim;
        const float *mn = qs + (int64_t)8 * dim, *iv = mn + dim;  // CodecSQ8::stage layout (QT = 8 slots)
        const int64_t row = base + tid;
        if (row >= row_end) return;
        const uint8_t *code = P.codes + row * P.row_bytes;
        float acc[QT];
#pragma unroll
        for (int q = 0; q < QT; q++) acc[q] = 0.0f;
        for (int64_t d = 0; d < dim; d++) {
            const float val = __fadd_rn(mn[d], __fmul_rn(u8_to_f32(__ldg(code + d)), iv[d]));
#pragma unroll
            for (int q = 0; q < QT; q++) {
                const float qv = qs[(int64_t)q * dim + d];
                if (is_dot) {
                    acc[q] = __fadd_rn(acc[q], __fmul_rn(qv, val));
                } else {
                    const float df = __fsub_rn(qv, val);
                    acc[q] = __fadd_rn(acc[q], __fmul_rn(df, df));
                }
            }
        }
#pragma unroll
        for (int q = 0; q < QT; q++)
            if (q < nqv) sink(q, row, acc[q]);
    }
};

// ============================================================ codec: INT4
// simd.Int4L2DistanceBatch (int4_avx512.c:193-299).  High nibble = even dim.
__device__ __forceinline__ float int4_nibble(const uint8_t *code, int64_t d) {
    const uint32_t b = __ldg(code + (d >> 1));
    return u8_to_f32((d & 1) ? (b & 0x0Fu) : (b >> 4));
}
struct CodecINT4 {
    static constexpr int QT = 8, R = 2, THREADS = 256, RB = 16 * R, MINB = 2;
    static size_t qsmem(const CodecParams &P) { return (size_t)(QT + 2) * P.dim * 4; }
    __device__ static void stage(const CodecParams &P, const float *queries, int64_t qs_, int q0, int nqv, unsigned char *sm, int tid) {
        CodecSQ8::stage(P, queries, qs_, q0, nqv, sm, tid);  // same layout: queries | p0 (min) | p1 (diff)
    }
    template <class Sink>
    __device__ static void tile(const CodecParams &P, bool, const unsigned char *sm, int64_t base, int64_t row_end, int nqv,
                                int tid, const Sink &sink) {
        const float k15 = __uint_as_float(0x3d888889u);
        const float *qs = reinterpret_cast<const float *>(sm);
        const int64_t dim = P.dim;
        const float *mn = qs + (int64_t)QT * dim, *df_ = mn + dim;
        const int hw = tid >> 4, lane = tid & 15;
        const int64_t r0 = base + (int64_t)hw * R;
        const uint8_t *code[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            int64_t rr = r0 + r;
            if (rr > row_end - 1) rr = row_end - 1;
            code[r] = P.codes + rr * P.row_bytes;
        }
        float s1[R][QT], s2[R][QT];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int q = 0; q < QT; q++) s1[r][q] = s2[r][q] = 0.0f;
        int64_t i = 0;
        for (; i + 64 <= dim; i += 64) {
#pragma unroll
            for (int blk = 0; blk < 4; blk++) {
                const int64_t d = i + blk * 16 + lane;
                const float m = mn[d], dd = df_[d];
                float deq[R];
#pragma unroll
                for (int r = 0; r < R; r++) deq[r] = __fmaf_rn(__fmul_rn(int4_nibble(code[r], d), k15), dd, m);
#pragma unroll
                for (int q = 0; q < QT; q++) {
                    const float qv = qs[(int64_t)q * dim + d];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const float e = __fsub_rn(qv, deq[r]);
                        if (blk < 2) s1[r][q] = __fmaf_rn(e, e, s1[r][q]);
                        else s2[r][q] = __fmaf_rn(e, e, s2[r][q]);
                    }
                }
            }
        }
        for (; i + 32 <= dim; i += 32) {
#pragma unroll
            for (int blk = 0; blk < 2; blk++) {
                const int64_t d = i + blk * 16 + lane;
                const float m = mn[d], dd = df_[d];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const float deq = __fmaf_rn(__fmul_rn(int4_nibble(code[r], d), k15), dd, m);
#pragma unroll
                    for (int q = 0; q < QT; q++) {
                        const float e = __fsub_rn(qs[(int64_t)q * dim + d], deq);
                        s1[r][q] = __fmaf_rn(e, e, s1[r][q]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int q = 0; q < QT; q++) s1[r][q] = reduce16(__fadd_rn(s1[r][q], s2[r][q]));
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (r0 + r >= row_end) continue;
#pragma unroll
                for (int q = 0; q < QT; q++) {
                    if (q >= nqv) continue;
                    float tot = s1[r][q];
                    for (int64_t d = i; d < dim; d++) {
                        const float deq = __fmaf_rn(__fmul_rn(int4_nibble(code[r], d), k15), df_[d], mn[d]);
                        const float e = __fsub_rn(qs[(int64_t)q * dim + d], deq);
                        tot = __fmaf_rn(e, e, tot);
                    }
                    sink(q, r0 + r, tot);
                }
            }
        }
    }
};

// INT4 fast path: dim % 256 == 0, codes lane-transposed on device
// (permute_int4): inside each 128-byte block (256 dims = 4 epochs of 64) the
// byte holding dims (64e + 16b + 2p, +1) of lane pair p sits at 16p + 4e + b,
// so lanes 2p and 2p+1 read the same 16 bytes (one broadcast request) and get
// their nibbles for 4 epochs x 4 blocks.  Shared-memory layout and the packed
// two-queries-per-instruction arithmetic are CodecSQ8Perm's: ndeq =
// fma(g, -diff, -min) = -deq exactly, e = q + ndeq, S = fma(e, e, S).
struct CodecINT4Perm {
    static constexpr int QT = 8, R = 2, THREADS = 256, RB = 16 * R, MINB = 2;
    static size_t qsmem(const CodecParams &P) { return CodecSQ8Perm<16>::qsmem(P); }
    __device__ static void stage(const CodecParams &P, const float *queries, int64_t qs_, int q0, int nqv, unsigned char *sm, int tid) {
        CodecSQ8Perm<16>::stage(P, queries, qs_, q0, nqv, sm, tid);  // queries | -min | -diff
    }
    template <class Sink>
    __device__ static void tile(const CodecParams &P, bool, const unsigned char *sm, int64_t base, int64_t row_end, int nqv,
                                int tid, const Sink &sink) {
        const float k15 = __uint_as_float(0x3d888889u);
        const int64_t dim = P.dim;
        const int SP = CodecSQ8Perm<16>::steps_padded(dim), LS = CodecSQ8Perm<16>::lane_stride(dim);
        const float *qs = reinterpret_cast<const float *>(sm);
        const float *mn = qs + (int64_t)16 * LS, *dfp = mn + 16 * SP;
        const int hw = tid >> 4, lane = tid & 15;
        const int64_t r0 = base + (int64_t)hw * R;
        const uint8_t *code[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            int64_t rr = r0 + r;
            if (rr > row_end - 1) rr = row_end - 1;
            code[r] = P.codes + rr * P.row_bytes + (lane >> 1) * 16;
        }
        const int shift = (lane & 1) ? 0 : 4;  // even lane (even dim) = high nibble
        f32x2 s1[R][QT / 2], s2[R][QT / 2];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int q = 0; q < QT / 2; q++) s1[r][q] = s2[r][q] = 0ull;
        const float *mnl = mn + lane * SP, *dfl = dfp + lane * SP;
        const float *ql = qs + (int64_t)lane * LS;
        const int nblk = (int)(dim / 256);
        for (int b = 0; b < nblk; b++) {
            uint32_t w[R][4];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(code[r] + (int64_t)b * 128));
                w[r][0] = v.x;
                w[r][1] = v.y;
                w[r][2] = v.z;
                w[r][3] = v.w;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {  // epoch e of this block: steps t0..t0+3 are blk 0..3
                const int t0 = b * 16 + e * 4;
                const float4 m4 = *reinterpret_cast<const float4 *>(mnl + t0);
                const float4 d4 = *reinterpret_cast<const float4 *>(dfl + t0);
                const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
                const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int blk = 0; blk < 4; blk++) {
                    const ulonglong2 qa = *reinterpret_cast<const ulonglong2 *>(ql + (t0 + blk) * QT);
                    const ulonglong2 qb = *reinterpret_cast<const ulonglong2 *>(ql + (t0 + blk) * QT + 4);
                    const f32x2 qq[4] = {qa.x, qa.y, qb.x, qb.y};
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const uint32_t nib = (w[r][e] >> (8 * blk + shift)) & 0xFu;
                        const float f = __fsub_rn(__uint_as_float(0x4B000000u | nib), 8388608.0f);
                        const float ndeq = __fmaf_rn(__fmul_rn(f, k15), dd[blk], mm[blk]);
                        const f32x2 n2 = pk2(ndeq, ndeq);
#pragma unroll
                        for (int q = 0; q < QT / 2; q++) {
                            const f32x2 ee = add2(qq[q], n2);
                            if (blk < 2) s1[r][q] = fma2(ee, ee, s1[r][q]);
                            else s2[r][q] = fma2(ee, ee, s2[r][q]);
                        }
                    }
                }
            }
        }
        float v[16];
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
            for (int q = 0; q < QT / 2; q++) {
                const f32x2 t = add2(s1[r][q], s2[r][q]);
                unpk2(t, v[r * QT + 2 * q], v[r * QT + 2 * q + 1]);
            }
        const float tot = reduce16_transposed16(v, lane);
        // lane l holds accumulator l = (row l/8, query l%8)
        const int64_t row = r0 + (lane >> 3);
        const int q = lane & 7;
        if (row < row_end && q < nqv) sink(q, row, tot);
    }
};

// ============================================================ codec: PQ ADC
// simd.PqAdcLookup (floats_avx512.c:135-167): lane l sums table[(16t+l)*256 +
// code[16t+l]] over t in order, reduce, sequential tail.  The per-query table
// is simd.BuildDistanceTableInt8 (kernels.go:354-374, the live generic path:
// sequential, unfused) built INSIDE the kernel from the int8 codebooks — or
// copied from caller-provided tables for the simd mirror.
//
// One CTA owns TWO queries; half-warp 0 of every warp serves query A, half-warp
// 1 query B.  Table layout: word (t*256 + c)*32 + 16*half + lane, so the 32
// lanes of any lookup instruction hit 32 distinct banks whatever the codes are.
struct CodecPQ {
    static constexpr int QT = 2, R = 4, THREADS = 256, RB = (THREADS / 32) * R, MINB = 1;
    static size_t qsmem(const CodecParams &P) {
        const int t16 = P.pq_m / 16, tail = P.pq_m % 16;
        return ((size_t)t16 * 256 * 32 + (size_t)QT * tail * 256) * 4;
    }
    __device__ static float entry(const CodecParams &P, const float *query, int m, int c) {
        const int ds = P.pq_dsub;
        const int8_t *cb = P.pq_codebooks + ((int64_t)m * P.pq_k + c) * ds;
        const float scale = P.pq_scales[m], offset = P.pq_offsets[m];
        const float *qv = query + (int64_t)m * ds;
        float sum = 0.0f;
        for (int i = 0; i < ds; i++) {
            const float v = __fadd_rn(__fmul_rn((float)cb[i], scale), offset);
            const float d = __fsub_rn(qv[i], v);
            sum = __fadd_rn(sum, __fmul_rn(d, d));
        }
        return sum;
    }
    // `queries`: float vectors [nq][dim] (tables built here) or, when P.pq_tables is set, ignored.
    __device__ static void stage(const CodecParams &P, const float *queries, int64_t qs_, int q0, int nqv, unsigned char *sm, int tid) {
        float *lut = reinterpret_cast<float *>(sm);
        const int M = P.pq_m, t16 = M / 16, tail = M % 16;
        float *tl = lut + (int64_t)t16 * 256 * 32;
        for (int idx = tid; idx < M * 256; idx += THREADS) {
            const int m = idx >> 8, c = idx & 255;
#pragma unroll
            for (int h = 0; h < QT; h++) {
                float v = 0.0f;
                if (h < nqv) {
                    if (P.pq_tables) v = P.pq_tables[(int64_t)(q0 + h) * M * 256 + idx];
                    else v = entry(P, queries + (int64_t)(q0 + h) * qs_, m, c);
                }
                if (m < t16 * 16) lut[((int64_t)(m >> 4) * 256 + c) * 32 + 16 * h + (m & 15)] = v;
                else tl[((int64_t)h * tail + (m - t16 * 16)) * 256 + c] = v;
            }
        }
    }
    template <class Sink>
    __device__ static void tile(const CodecParams &P, bool, const unsigned char *sm, int64_t base, int64_t row_end, int nqv,
                                int tid, const Sink &sink) {
        const float *lut = reinterpret_cast<const float *>(sm);
        const int M = P.pq_m, t16 = M / 16, tail = M % 16;
        const float *tl = lut + (int64_t)t16 * 256 * 32;
        const int warp = tid >> 5, half = (tid >> 4) & 1, lane = tid & 15;
        const int64_t r0 = base + (int64_t)warp * R;
        const uint8_t *code[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            int64_t rr = r0 + r;
            if (rr > row_end - 1) rr = row_end - 1;
            code[r] = P.codes + rr * P.row_bytes;
        }
        float s[R];
#pragma unroll
        for (int r = 0; r < R; r++) s[r] = 0.0f;
        const float *lane_lut = lut + 16 * half + lane;
        for (int t = 0; t < t16; t++) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const uint32_t c = __ldg(code[r] + t * 16 + lane);
                s[r] = __fadd_rn(s[r], lane_lut[((int64_t)t * 256 + c) * 32]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) s[r] = reduce16(s[r]);
        if (lane == 0 && half < nqv) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (r0 + r >= row_end) continue;
                float tot = s[r];
                for (int m = 0; m < tail; m++)
                    tot = __fadd_rn(tot, tl[((int64_t)half * tail + m) * 256 + __ldg(code[r] + t16 * 16 + m)]);
                sink(half, r0 + r, tot);
            }
        }
    }
};

// PQ ADC fast path (K = 256, codes tiled by permute_pq): ONE THREAD PER ROW, two
// queries per CTA whose tables are interleaved as float2 (T_q0[m][c], T_q1[m][c]),
// so a single 8-byte shared-memory lookup + one FADD2 advances both queries.
// The thread keeps the reference's 16 lane accumulators itself (lane l sums
// table[(16t+l)*256 + code[16t+l]] for t = 0,1,...), then applies the
// _mm512_reduce_add_ps tree and the sequential tail (floats_avx512.c:135-167) —
// the same additions in the same order as the half-warp version above.
// Bound: random 8-byte LDS = 7.6 lookups/clk/SM measured (profiles/r01_ubench.log).
struct CodecPQ2 {
    static constexpr int QT = 2, R = 1, THREADS = 512, RB = THREADS, MINB = 1;
    static constexpr bool PIPELINED = true;
    static constexpr int MAXCH = 6;  // code chunks of 16 subspaces held in registers: M <= 96
    // The codes of the NEXT tile are requested before the current tile's lookups start, so the L2/HBM latency
    // overlaps ~100 shared-memory lookups instead of being exposed after every block-wide barrier.
    struct State {
        uint4 w4[MAXCH];
    };
    static size_t qsmem(const CodecParams &P) { return (size_t)P.pq_m * 256 * 8; }
    __device__ static void stage(const CodecParams &P, const float *queries, int64_t qs_, int q0, int nqv, unsigned char *sm, int tid) {
        float2 *lut = reinterpret_cast<float2 *>(sm);
        const int M = P.pq_m;
        for (int idx = tid; idx < M * 256; idx += THREADS) {
            const int m = idx >> 8, c = idx & 255;
            float v[2] = {0.0f, 0.0f};
#pragma unroll
            for (int h = 0; h < 2; h++)
                if (h < nqv) {
                    if (P.pq_tables) v[h] = P.pq_tables[(int64_t)(q0 + h) * M * 256 + idx];
                    else v[h] = CodecPQ::entry(P, queries + (int64_t)(q0 + h) * qs_, m, c);
                }
            lut[idx] = make_float2(v[0], v[1]);
        }
    }
    __device__ static void prefetch(const CodecParams &P, State &st, int64_t base, int64_t row_end, int tid) {
        const int chunks = (P.pq_m + 15) >> 4;
        int64_t rr = base + tid;
        if (rr > row_end - 1) rr = row_end - 1;
        if (rr < 0) rr = 0;
        const uint8_t *cbase = P.codes + (rr >> 5) * (32 * P.row_bytes) + (rr & 31) * 16;
#pragma unroll
        for (int t = 0; t < MAXCH; t++)
            if (t < chunks) st.w4[t] = __ldg(reinterpret_cast<const uint4 *>(cbase + (int64_t)t * 512));
    }
    // `next_base`: first row of the tile this thread block will process after this one (prefetched here).
    template <class Sink>
    __device__ static void tile_p(const CodecParams &P, bool, const unsigned char *sm, int64_t base, int64_t next_base,
                                  int64_t row_end, int nqv, int tid, const Sink &sink, State &st) {
        const int M = P.pq_m, t16 = M >> 4, tail = M & 15;
        const int64_t row = base + tid;
        const unsigned char *lut = sm;
        uint4 w4[MAXCH];
#pragma unroll
        for (int t = 0; t < MAXCH; t++) w4[t] = st.w4[t];
        if (next_base < row_end) prefetch(P, st, next_base, row_end, tid);
        f32x2 acc[16];
#pragma unroll
        for (int j = 0; j < 16; j++) acc[j] = 0ull;
#pragma unroll
        for (int t = 0; t < MAXCH; t++) {
            if (t < t16) {
                const uint32_t w[4] = {w4[t].x, w4[t].y, w4[t].z, w4[t].w};
                const unsigned char *lt = lut + (size_t)t * 16 * 2048;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint32_t off = (j & 3) == 0 ? (w[j >> 2] << 3) & 0x7F8u : (w[j >> 2] >> (8 * (j & 3) - 3)) & 0x7F8u;
                    acc[j] = add2(acc[j], *reinterpret_cast<const f32x2 *>(lt + j * 2048 + off));
                }
            }
        }
        // _mm512_reduce_add_ps: i+8, i+4, i+2, i+1
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = add2(acc[j], acc[j + 8]);
#pragma unroll
        for (int j = 0; j < 4; j++) acc[j] = add2(acc[j], acc[j + 4]);
        acc[0] = add2(acc[0], acc[2]);
        acc[1] = add2(acc[1], acc[3]);
        f32x2 tot = add2(acc[0], acc[1]);
        if (tail) {
#pragma unroll
            for (int t = 0; t < MAXCH; t++) {
                if (t == t16) {
                    const uint32_t w[4] = {w4[t].x, w4[t].y, w4[t].z, w4[t].w};
                    const unsigned char *lt = lut + (size_t)t * 16 * 2048;
#pragma unroll
                    for (int j = 0; j < 15; j++) {
                        if (j < tail) {
                            const uint32_t c = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                            tot = add2(tot, *reinterpret_cast<const f32x2 *>(lt + j * 2048 + c * 8));
                        }
                    }
                }
            }
        }
        if (row < row_end) {
            float s0, s1;
            unpk2(tot, s0, s1);
            sink(0, row, s0);
            if (nqv > 1) sink(1, row, s1);
        }
    }
};

// ============================================================ codec: sign bits
// BQ: score = float32(Hamming) (distance.Hamming).  RaBitQ: estimator of
// rabitq.go:119-176, unfused Go arithmetic.  Integer popcounts are exact in any
// order, so one thread owns one row and loops over the CTA's QT queries.
// Query side = prepared sign words P.q_words [nq][words32] (+ P.q_norms for
// RaBitQ), see prep_sign_queries(); the float `queries` pointer is unused.
template <bool RABITQ>
struct CodecSign {
    static constexpr int QT = 8, R = 1, THREADS = 256, RB = THREADS, MINB = 2;
    static size_t qsmem(const CodecParams &P) { return (size_t)QT * (P.words32 + 4) * 4; }
    __device__ static void stage(const CodecParams &P, const float *queries, int64_t qs_, int q0, int nqv, unsigned char *sm, int tid) {
        (void)queries;
        (void)qs_;
        uint32_t *qw = reinterpret_cast<uint32_t *>(sm);
        const int W = P.words32;
        for (int i = tid; i < QT * W; i += THREADS) {
            const int q = i / W, j = i - q * W;
            qw[i] = (q < nqv) ? P.q_words[(int64_t)(q0 + q) * W + j] : 0u;
        }
        float *qn = reinterpret_cast<float *>(qw + QT * W);
        if (tid < QT) qn[tid] = (RABITQ && tid < nqv) ? P.q_norms[q0 + tid] : 0.0f;
    }
    template <class Sink>
    __device__ static void tile(const CodecParams &P, bool, const unsigned char *sm, int64_t base, int64_t row_end, int nqv,
                                int tid, const Sink &sink) {
        const uint32_t *qw = reinterpret_cast<const uint32_t *>(sm);
        const int W = P.words32;
        const float *qn = reinterpret_cast<const float *>(qw + QT * W);
        const int64_t row = base + tid;
        if (row >= row_end) return;
        const uint32_t *code = reinterpret_cast<const uint32_t *>(P.codes + row * P.row_bytes);
        int h[QT];
#pragma unroll
        for (int q = 0; q < QT; q++) h[q] = 0;
        int j = 0;
        for (; j + 4 <= W; j += 4) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(code + j));
#pragma unroll
            for (int q = 0; q < QT; q++) {
                const uint4 u = *reinterpret_cast<const uint4 *>(qw + q * W + j);
                h[q] += __popc(v.x ^ u.x) + __popc(v.y ^ u.y) + __popc(v.z ^ u.z) + __popc(v.w ^ u.w);
            }
        }
        for (; j < W; j++) {
            const uint32_t v = __ldg(code + j);
#pragma unroll
            for (int q = 0; q < QT; q++) h[q] += __popc(v ^ qw[q * W + j]);
        }
        float yn = 0.0f;
        if (RABITQ) yn = __ldg(P.norms + row);
        const float fdim = (float)P.dim;
#pragma unroll
        for (int q = 0; q < QT; q++) {
            if (q >= nqv) continue;
            float score;
            if (RABITQ) {
                const float hm = (float)h[q];
                const float t1 = __fsub_rn(qn[q], yn);
                const float t1sq = __fmul_rn(t1, t1);
                float a = __fmul_rn(4.0f, qn[q]);
                a = __fmul_rn(a, yn);
                a = __fdiv_rn(a, fdim);
                score = __fadd_rn(t1sq, __fmul_rn(a, hm));
            } else {
                score = (float)h[q];
            }
            sink(q, row, score);
        }
    }
};

// ============================================================ kernels
template <class T, class = void>
struct is_pipelined : std::false_type {};
template <class T>
struct is_pipelined<T, std::enable_if_t<T::PIPELINED>> : std::true_type {};

template <class Codec>
__global__ void __launch_bounds__(Codec::THREADS, Codec::MINB) scan_topk_kernel(CodecParams P, ScanArgs A, size_t qbytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * Codec::QT;
    const int nqv = (A.nq - q0 < Codec::QT) ? (int)(A.nq - q0) : Codec::QT;
    const int split = blockIdx.y;
    SinkTopK sink;
    sink.tk = topk_carve(smem + qbytes, Codec::QT, A.C, A.k);
    sink.mask = A.mask;
    sink.probe = A.probe;
    sink.part_off = A.part_off;
    sink.nprobe = A.nprobe;
    sink.trigger = A.trigger;
    sink.q0 = q0;
    sink.row_base = A.row_base;
    sink.descending = A.descending != 0;
    Codec::stage(P, A.queries, A.q_stride ? A.q_stride : P.dim, q0, nqv, smem, tid);
    topk_init(sink.tk, Codec::QT, tid, Codec::THREADS);
    __syncthreads();
    int64_t row_begin = (int64_t)split * A.rows_per_split;
    int64_t row_end = row_begin + A.rows_per_split;
    if (row_end > A.rows) row_end = A.rows;
    if (A.by_partition) {
        // the block's rows: from the first to the last partition of its (sorted) virtual queries — one partition, or two
        // neighbouring ones where the block straddles a boundary of the sorted list; the sink keeps each slot to its own
        int64_t lo = A.rows, hi = 0;
        for (int s = 0; s < nqv; s++) {
            const int p = A.probe[q0 + s];
            if (p < 0) continue;
            const int64_t b = A.part_off[p], e = A.part_off[p + 1];
            if (e > b) {
                lo = b < lo ? b : lo;
                hi = e > hi ? e : hi;
            }
        }
        row_begin = lo / Codec::RB * Codec::RB;
        row_end = hi < A.rows ? hi : A.rows;
        if (row_end < row_begin) row_end = row_begin;
    }
    if constexpr (is_pipelined<Codec>::value) {
        typename Codec::State stt;
        Codec::prefetch(P, stt, row_begin, row_end, tid);
        for (int64_t base = row_begin; base < row_end; base += Codec::RB) {
            Codec::tile_p(P, A.is_dot != 0, smem, base, base + Codec::RB, row_end, nqv, tid, sink, stt);
            __syncthreads();
            topk_block_maintain(sink.tk, Codec::QT, tid, Codec::THREADS);
        }
    } else {
        for (int64_t base = row_begin; base < row_end; base += Codec::RB) {
            Codec::tile(P, A.is_dot != 0, smem, base, row_end, nqv, tid, sink);
            __syncthreads();
            topk_block_maintain(sink.tk, Codec::QT, tid, Codec::THREADS);
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int s = warp; s < nqv; s += Codec::THREADS / 32) {
        const int64_t q = q0 + s;
        if (A.by_partition)
            topk_emit_keys_warp(sink.tk, s, lane, A.partial + (int64_t)A.emit_index[q] * A.k, A.k);
        else if (A.splits == 1)
            topk_emit_warp(sink.tk, s, lane, A.descending != 0, A.out_rows + q * A.k, A.out_scores + q * A.k, A.out_counts + q,
                           A.k);
        else
            topk_emit_keys_warp(sink.tk, s, lane, A.partial + (q * A.splits + split) * A.k, A.k);
    }
}

template <class Codec>
__global__ void __launch_bounds__(Codec::THREADS, Codec::MINB)
scan_dense_kernel(CodecParams P, const float *queries, int64_t nq, int64_t n, int is_dot, float *out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * Codec::QT;
    const int nqv = (nq - q0 < Codec::QT) ? (int)(nq - q0) : Codec::QT;
    SinkDense sink{out, n, q0};
    Codec::stage(P, queries, P.dim, q0, nqv, smem, tid);
    __syncthreads();
    if constexpr (is_pipelined<Codec>::value) {
        typename Codec::State stt;
        const int64_t step = (int64_t)gridDim.y * Codec::RB;
        Codec::prefetch(P, stt, (int64_t)blockIdx.y * Codec::RB, n, tid);
        for (int64_t base = (int64_t)blockIdx.y * Codec::RB; base < n; base += step)
            Codec::tile_p(P, is_dot != 0, smem, base, base + step, n, nqv, tid, sink, stt);
    } else {
        for (int64_t base = (int64_t)blockIdx.y * Codec::RB; base < n; base += (int64_t)gridDim.y * Codec::RB)
            Codec::tile(P, is_dot != 0, smem, base, n, nqv, tid, sink);
    }
}

// One warp per query merges `lists` sorted key lists.
__global__ void __launch_bounds__(256) merge_keys_kernel(const unsigned long long *keys, int64_t lists, int64_t nq, int64_t k_in,
                                                         int64_t list_stride, int64_t query_stride, int descending, int k_out,
                                                         int C, uint32_t *out_rows, float *out_scores, int32_t *out_counts) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    TopK tk = topk_carve(smem, nw, C, k_out);
    topk_init(tk, nw, threadIdx.x, blockDim.x);
    __syncthreads();
    const int64_t q = (int64_t)blockIdx.x * nw + warp;
    if (q >= nq) return;
    const int trigger = C - 32;
    for (int64_t l = 0; l < lists; l++) {
        const unsigned long long *src = keys + l * list_stride + q * query_stride;
        for (int64_t i0 = 0; i0 < k_in; i0 += 32) {
            const int64_t i = i0 + lane;
            if (i < k_in) {
                const unsigned long long key = src[i];
                if (key != VG_KEY_EMPTY && key < tk.tau[warp]) {
                    int pos = atomicAdd(&tk.cnt[warp], 1);
                    if (pos < C) tk.keys[(size_t)warp * C + pos] = key;
                }
            }
            __syncwarp();
            if (tk.cnt[warp] > trigger) topk_compact_warp(tk, warp, lane, false);
            __syncwarp();
        }
    }
    topk_emit_warp(tk, warp, lane, descending != 0, out_rows + q * k_out, out_scores + q * k_out, out_counts + q, k_out);
}

__global__ void pairs_to_keys_kernel(const uint32_t *rows, const float *scores, int64_t n, int descending,
                                     unsigned long long *keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (rows[i] == 0xFFFFFFFFu) ? VG_KEY_EMPTY : make_key(scores[i], rows[i], descending != 0);
}

vg_status launch_merge_keys(const unsigned long long *d_keys, int64_t lists, int64_t nq, int64_t k_in, int64_t list_stride,
                            int64_t query_stride, bool descending, int64_t k_out, uint32_t *d_rows, float *d_scores,
                            int32_t *d_counts, cudaStream_t st) {
    if (nq <= 0) return VG_OK;
    const int C = topk_capacity((int)k_out, 32);
    int nw = 8;
    while (nw > 1 && topk_smem_bytes(nw, C) > 200 * 1024) nw >>= 1;
    const size_t sm = topk_smem_bytes(nw, C);
    if (sm > 220 * 1024) return fail(VG_ERR_UNSUPPORTED, "k too large for the shared-memory top-k");
    VG_CUDA(cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int64_t blocks = (nq + nw - 1) / nw;
    merge_keys_kernel<<<(unsigned)blocks, nw * 32, sm, st>>>(d_keys, lists, nq, k_in, list_stride, query_stride, descending ? 1 : 0,
                                                             (int)k_out, C, d_rows, d_scores, d_counts);
    VG_LAUNCHED();
    return VG_OK;
}

vg_status launch_merge_pairs(const uint32_t *d_rows_in, const float *d_scores_in, int64_t lists, int64_t nq, int64_t k_in,
                             bool descending, int64_t k_out, uint32_t *d_rows, float *d_scores, int32_t *d_counts,
                             cudaStream_t st) {
    const int64_t n = lists * nq * k_in;
    if (n <= 0 || nq <= 0) return VG_OK;
    DevBuf keys;
    VG_TRY(keys.alloc((size_t)n * 8));
    pairs_to_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_rows_in, d_scores_in, n, descending ? 1 : 0,
                                                                      keys.as<unsigned long long>());
    VG_LAUNCHED();
    // `keys` goes back to the stream-ordered pool when this returns: freed in stream order, no host wait
    return launch_merge_keys(keys.as<unsigned long long>(), lists, nq, k_in, nq * k_in, k_in, descending, k_out, d_rows, d_scores,
                             d_counts, st);
}

// Exchange format of the sharded search: one 8-byte sortable key per candidate (what the all-gather moves and the merge
// reads directly): packing is one pass over the shard's [nq][k] result, merging needs no conversion pass.
vg_status launch_pack_keys(const uint32_t *d_rows_in, const float *d_scores_in, int64_t n, bool descending, unsigned long long *d_keys,
                           cudaStream_t st) {
    if (n <= 0) return VG_OK;
    pairs_to_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_rows_in, d_scores_in, n, descending ? 1 : 0, d_keys);
    VG_LAUNCHED();
    return VG_OK;
}

// ------------------------------------------------------------ dispatch
template <class Codec>
static vg_status run_topk(const CodecParams &cp, ScanArgs a, cudaStream_t st) {
    const size_t qb = (Codec::qsmem(cp) + 15) & ~(size_t)15;
    a.C = topk_capacity(a.k, Codec::RB);
    a.trigger = a.C - Codec::RB;
    const size_t sm = qb + topk_smem_bytes(Codec::QT, a.C);
    if (sm > 227 * 1024) return fail(VG_ERR_UNSUPPORTED, "query tile + top-k state exceed 227 KB of shared memory (dim or k too large)");
    const int64_t qtiles = (a.nq + Codec::QT - 1) / Codec::QT;
    if (a.by_partition) {  // one block per four virtual queries, no row splits; the caller owns `partial` and merges
        a.splits = 1;
        a.rows_per_split = a.rows;
        VG_CUDA(cudaFuncSetAttribute(scan_topk_kernel<Codec>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        scan_topk_kernel<Codec><<<dim3((unsigned)qtiles, 1), Codec::THREADS, sm, st>>>(cp, a, qb);
        VG_LAUNCHED();
        return VG_OK;
    }
    // Row splits only when the query tiles alone cannot fill the machine.
    const int64_t target = (int64_t)sm_count() * Codec::MINB;
    int64_t splits = 1;
    if (qtiles < target) {
        splits = (target + qtiles - 1) / qtiles;
        const int64_t min_rows = (int64_t)Codec::RB * 8;
        const int64_t max_splits = (a.rows + min_rows - 1) / min_rows;
        if (splits > max_splits) splits = max_splits;
        if (splits > 1024) splits = 1024;
        if (splits < 1) splits = 1;
    }
    int64_t rps = (a.rows + splits - 1) / splits;
    rps = (rps + Codec::RB - 1) / Codec::RB * Codec::RB;
    splits = (a.rows + rps - 1) / rps;
    if (splits < 1) splits = 1;
    a.splits = (int)splits;
    a.rows_per_split = rps;
    DevBuf partial;
    if (splits > 1) {
        VG_TRY(partial.alloc((size_t)a.nq * splits * a.k * 8));
        a.partial = partial.as<unsigned long long>();
    }
    VG_CUDA(cudaFuncSetAttribute(scan_topk_kernel<Codec>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid((unsigned)qtiles, (unsigned)splits);
    scan_topk_kernel<Codec><<<grid, Codec::THREADS, sm, st>>>(cp, a, qb);
    VG_LAUNCHED();
    if (splits > 1) {
        VG_TRY(launch_merge_keys(a.partial, splits, a.nq, a.k, a.k, splits * a.k, a.descending != 0, a.k, a.out_rows, a.out_scores,
                                 a.out_counts, st));
        VG_CUDA(cudaStreamSynchronize(st));  // partial is freed on return
    }
    return VG_OK;
}

template <class Codec>
static vg_status run_dense(const CodecParams &cp, const float *q, int64_t nq, int64_t n, int is_dot, float *out,
                           cudaStream_t st) {
    const size_t sm = (Codec::qsmem(cp) + 15) & ~(size_t)15;
    if (sm > 227 * 1024) return fail(VG_ERR_UNSUPPORTED, "query tile exceeds 227 KB of shared memory");
    if (nq <= 0 || n <= 0) return VG_OK;
    VG_CUDA(cudaFuncSetAttribute(scan_dense_kernel<Codec>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int64_t qtiles = (nq + Codec::QT - 1) / Codec::QT;
    int64_t gy = (n + Codec::RB - 1) / Codec::RB;
    const int64_t want = (4LL * sm_count() + qtiles - 1) / qtiles;
    if (gy > want) gy = want;
    if (gy > 65535) gy = 65535;
    dim3 grid((unsigned)qtiles, (unsigned)gy);
    scan_dense_kernel<Codec><<<grid, Codec::THREADS, sm, st>>>(cp, q, nq, n, is_dot, out);
    VG_LAUNCHED();
    return VG_OK;
}

#define VG_DISPATCH(FN, ...)                                                                                     \
    switch (cp.codec) {                                                                                          \
        case VG_CODEC_F32:                                                                                       \
            return (cp.variant & VG_VAR_BATCH) ? FN<CodecF32<true>>(__VA_ARGS__) : FN<CodecF32<false>>(__VA_ARGS__); \
        case VG_CODEC_SQ8:                                                                                       \
            if (cp.variant & VG_VAR_GO_SCALAR) return FN<CodecSQ8Go>(__VA_ARGS__);                               \
            if (cp.variant & VG_VAR_PERM)                                                                        \
                return (cp.dim % 256 == 0) ? FN<CodecSQ8Perm<16>>(__VA_ARGS__) : FN<CodecSQ8Perm<4>>(__VA_ARGS__); \
            return FN<CodecSQ8>(__VA_ARGS__);                                                                    \
        case VG_CODEC_INT4:                                                                                      \
            return (cp.variant & VG_VAR_PERM) ? FN<CodecINT4Perm>(__VA_ARGS__) : FN<CodecINT4>(__VA_ARGS__);     \
        case VG_CODEC_PQ:                                                                                        \
        case VG_CODEC_OPQ:                                                                                       \
            return (cp.variant & VG_VAR_PERM) ? FN<CodecPQ2>(__VA_ARGS__) : FN<CodecPQ>(__VA_ARGS__);            \
        default:                                                                                                 \
            break;                                                                                               \
    }

vg_status scan_topk(const CodecParams &cp, ScanArgs a, cudaStream_t st) {
    if (a.nq <= 0) return VG_OK;
    if (a.k <= 0) return fail(VG_ERR_INVALID, "k must be positive");
    if (a.rows <= 0) {
        // empty index: counts = 0, rows = 0xFFFFFFFF
        VG_CUDA(cudaMemsetAsync(a.out_counts, 0, (size_t)a.nq * 4, st));
        VG_CUDA(cudaMemsetAsync(a.out_rows, 0xFF, (size_t)a.nq * a.k * 4, st));
        VG_CUDA(cudaMemsetAsync(a.out_scores, 0xFF, (size_t)a.nq * a.k * 4, st));
        return VG_OK;
    }
    if ((cp.codec == VG_CODEC_PQ || cp.codec == VG_CODEC_OPQ) && cp.pq_k != 256)
        return fail(VG_ERR_UNSUPPORTED, "PQ ADC requires K=256 (simd.PqAdcLookup hard-wires the table stride, kernels.go:249)");
    VG_DISPATCH(run_topk, cp, a, st)
    if (cp.codec == VG_CODEC_BQ) return run_topk<CodecSign<false>>(cp, a, st);
    if (cp.codec == VG_CODEC_RABITQ) return run_topk<CodecSign<true>>(cp, a, st);
    return fail(VG_ERR_UNSUPPORTED, "unknown codec");
}

__global__ void gather_rows_kernel(const float *src, int64_t stride, const int32_t *idx, int64_t n, int64_t dim, float *dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * dim) return;
    const int64_t r = i / dim, c = i - r * dim;
    dst[i] = src[(int64_t)idx[r] * stride + c];
}
__global__ void scatter_results_kernel(const uint32_t *rows, const float *scores, const int32_t *counts, const int32_t *idx, int64_t n,
                                       int64_t k, uint32_t *out_rows, float *out_scores, int32_t *out_counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * k) return;
    const int64_t r = i / k, c = i - r * k;
    out_rows[(int64_t)idx[r] * k + c] = rows[i];
    out_scores[(int64_t)idx[r] * k + c] = scores[i];
    if (c == 0) out_counts[idx[r]] = counts[r];
}
vg_status dev_gather_rows(const float *d_src, int64_t stride, const int32_t *d_idx, int64_t n, int64_t dim, float *d_dst, cudaStream_t st) {
    if (n <= 0) return VG_OK;
    gather_rows_kernel<<<(unsigned)((n * dim + 255) / 256), 256, 0, st>>>(d_src, stride, d_idx, n, dim, d_dst);
    VG_LAUNCHED();
    return VG_OK;
}
vg_status dev_scatter_results(const uint32_t *d_rows, const float *d_scores, const int32_t *d_counts, const int32_t *d_idx, int64_t n,
                              int64_t k, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts, cudaStream_t st) {
    if (n <= 0) return VG_OK;
    scatter_results_kernel<<<(unsigned)((n * k + 255) / 256), 256, 0, st>>>(d_rows, d_scores, d_counts, d_idx, n, k, d_out_rows, d_out_scores,
                                                                           d_out_counts);
    VG_LAUNCHED();
    return VG_OK;
}
// ------------------------------------------------------------ partition-grouped scan
__global__ void __launch_bounds__(256) vq_count_kernel(const int32_t *probe, int64_t nv, int P, int *hist) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const int p = probe[i];
    atomicAdd(&hist[(p >= 0 && p < P) ? p : P], 1);
}
// exclusive scan of hist[0 .. n) into cursor (one block)
__global__ void __launch_bounds__(1024) vq_scan_kernel(const int *hist, int n, int *cursor) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + tid;
        const int v = i < n ? hist[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; w++) before += wsum[w];
        const int carry = carry_s;
        if (i < n) cursor[i] = carry + before + x - v;
        __syncthreads();
        if (tid == 1023) carry_s = carry + before + x;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) vq_scatter_kernel(const int32_t *probe, int64_t nv, int P, int np, int *cursor, int32_t *order,
                                                         int32_t *probe_v, int32_t *qidx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const int p = probe[i];
    const bool ok = p >= 0 && p < P;
    const int pos = atomicAdd(&cursor[ok ? p : P], 1);
    order[pos] = (int32_t)i;
    probe_v[pos] = ok ? p : -1;
    qidx[pos] = (int32_t)(i / np);
}
vg_status scan_topk_partitioned(const CodecParams &cp, ScanArgs a, cudaStream_t st) {
    if (a.nq <= 0) return VG_OK;
    if (a.k <= 0) return fail(VG_ERR_INVALID, "k must be positive");
    if (!a.probe || a.nprobe < 1 || !a.part_off || a.num_parts < 1) return fail(VG_ERR_INVALID, "partitioned scan without probes");
    if (a.rows <= 0) return scan_topk(cp, a, st);
    const int np = a.nprobe, P = a.num_parts;
    const int64_t nv = a.nq * np, dim = cp.dim, k = a.k;
    if (nv >= (1ll << 31)) return fail(VG_ERR_UNSUPPORTED, "too many (query, partition) pairs in one batch");
    DevBuf hist, cursor, order, probe_v, qidx, vq, partial, vqw, vqn;
    VG_TRY(hist.alloc((size_t)(P + 1) * 4));
    VG_TRY(cursor.alloc((size_t)(P + 1) * 4));
    VG_TRY(order.alloc((size_t)nv * 4));
    VG_TRY(probe_v.alloc((size_t)nv * 4));
    VG_TRY(qidx.alloc((size_t)nv * 4));
    VG_TRY(vq.alloc((size_t)nv * dim * 4));
    VG_TRY(partial.alloc((size_t)nv * k * 8));
    VG_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)(P + 1) * 4, st));
    vq_count_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(a.probe, nv, P, hist.as<int>());
    VG_LAUNCHED();
    vq_scan_kernel<<<1, 1024, 0, st>>>(hist.as<int>(), P + 1, cursor.as<int>());
    VG_LAUNCHED();
    vq_scatter_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(a.probe, nv, P, np, cursor.as<int>(), order.as<int32_t>(), probe_v.as<int32_t>(),
                                                                   qidx.as<int32_t>());
    VG_LAUNCHED();
    gather_rows_kernel<<<(unsigned)((nv * dim + 255) / 256), 256, 0, st>>>(a.queries, a.q_stride ? a.q_stride : dim, qidx.as<int32_t>(), nv, dim,
                                                                          vq.as<float>());
    VG_LAUNCHED();
    CodecParams cps = cp;
    if (cp.q_words) {  // sign codecs scan prepared per-query sign words (+ norms)
        VG_TRY(vqw.alloc((size_t)nv * cp.words32 * 4));
        gather_rows_kernel<<<(unsigned)((nv * cp.words32 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float *>(cp.q_words), cp.words32,
                                                                                    qidx.as<int32_t>(), nv, cp.words32, vqw.as<float>());
        VG_LAUNCHED();
        cps.q_words = vqw.as<uint32_t>();
    }
    if (cp.q_norms) {
        VG_TRY(vqn.alloc((size_t)nv * 4));
        gather_rows_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(cp.q_norms, 1, qidx.as<int32_t>(), nv, 1, vqn.as<float>());
        VG_LAUNCHED();
        cps.q_norms = vqn.as<float>();
    }
    ScanArgs av = a;
    av.queries = vq.as<float>();
    av.q_stride = 0;
    av.nq = nv;
    av.probe = probe_v.as<int32_t>();
    av.nprobe = 1;
    av.by_partition = 1;
    av.emit_index = order.as<int32_t>();
    av.partial = partial.as<unsigned long long>();
    VG_TRY(scan_topk(cps, av, st));
    // the nprobe sorted lists of query q sit at partial[(q * nprobe + j) * k]
    return launch_merge_keys(av.partial, np, a.nq, k, k, (int64_t)np * k, a.descending != 0, (int)k, a.out_rows, a.out_scores, a.out_counts, st);
}

vg_status scan_topk_subset(const CodecParams &cp, ScanArgs a, const std::vector<int32_t> &which, cudaStream_t st) {
    const int64_t nb = (int64_t)which.size();
    if (nb == 0) return VG_OK;
    const int64_t dim = cp.dim, k = a.k;
    DevBuf bidx, bq, brow, bsc, bcnt;
    VG_TRY(bidx.alloc((size_t)nb * 4));
    VG_TRY(bq.alloc((size_t)nb * dim * 4));
    VG_TRY(brow.alloc((size_t)nb * k * 4));
    VG_TRY(bsc.alloc((size_t)nb * k * 4));
    VG_TRY(bcnt.alloc((size_t)nb * 4));
    VG_CUDA(cudaMemcpyAsync(bidx.p, which.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, st));
    gather_rows_kernel<<<(unsigned)((nb * dim + 255) / 256), 256, 0, st>>>(a.queries, a.q_stride ? a.q_stride : dim, bidx.as<int32_t>(), nb, dim,
                                                                          bq.as<float>());
    VG_LAUNCHED();
    // sign codecs (BQ / RaBitQ) scan prepared per-query sign words (+ norms): gather those rows as well
    CodecParams cps = cp;
    DevBuf bqw, bqn;
    if (cp.q_words) {
        VG_TRY(bqw.alloc((size_t)nb * cp.words32 * 4));
        gather_rows_kernel<<<(unsigned)((nb * cp.words32 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float *>(cp.q_words), cp.words32,
                                                                                    bidx.as<int32_t>(), nb, cp.words32, bqw.as<float>());
        VG_LAUNCHED();
        cps.q_words = bqw.as<uint32_t>();
    }
    if (cp.q_norms) {
        VG_TRY(bqn.alloc((size_t)nb * 4));
        gather_rows_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(cp.q_norms, 1, bidx.as<int32_t>(), nb, 1, bqn.as<float>());
        VG_LAUNCHED();
        cps.q_norms = bqn.as<float>();
    }
    uint32_t *out_rows = a.out_rows;
    float *out_scores = a.out_scores;
    int32_t *out_counts = a.out_counts;
    a.queries = bq.as<float>();
    a.q_stride = 0;
    a.nq = nb;
    a.out_rows = brow.as<uint32_t>();
    a.out_scores = bsc.as<float>();
    a.out_counts = bcnt.as<int32_t>();
    VG_TRY(scan_topk(cps, a, st));
    scatter_results_kernel<<<(unsigned)((nb * k + 255) / 256), 256, 0, st>>>(brow.as<uint32_t>(), bsc.as<float>(), bcnt.as<int32_t>(),
                                                                            bidx.as<int32_t>(), nb, k, out_rows, out_scores, out_counts);
    VG_LAUNCHED();
    VG_CUDA(cudaStreamSynchronize(st));  // `which` (host) and the temporaries are released on return
    return VG_OK;
}

vg_status scan_dense(const CodecParams &cp, const float *d_queries, int64_t nq, int64_t n, int is_dot, float *d_out,
                     cudaStream_t st) {
    if ((cp.codec == VG_CODEC_PQ || cp.codec == VG_CODEC_OPQ) && cp.pq_k != 256)
        return fail(VG_ERR_UNSUPPORTED, "PQ ADC requires K=256");
    VG_DISPATCH(run_dense, cp, d_queries, nq, n, is_dot, d_out, st)
    if (cp.codec == VG_CODEC_BQ) return run_dense<CodecSign<false>>(cp, d_queries, nq, n, is_dot, d_out, st);
    if (cp.codec == VG_CODEC_RABITQ) return run_dense<CodecSign<true>>(cp, d_queries, nq, n, is_dot, d_out, st);
    return fail(VG_ERR_UNSUPPORTED, "unknown codec");
}

// ------------------------------------------------------------ rerank (gather)
// flat.Rerank (segment.go:754-781): half-warp per (query, candidate) pair.
__global__ void __launch_bounds__(256) rerank_kernel(const float *vectors, int64_t nrows, int64_t dim, const float *queries,
                                                     int64_t nq, const uint32_t *rows, int64_t r, int is_dot, float *out) {
    const int64_t pair = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const int64_t total = nq * r;
    const bool live = pair < total;
    const int64_t p = live ? pair : total - 1;
    const int64_t q = p / r;
    const uint32_t row = rows[p];
    const bool valid = (int64_t)row < nrows;
    const float *x = vectors + (valid ? (int64_t)row : 0) * dim;
    const float *qv = queries + q * dim;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t epochs = dim >> 6;
    for (int64_t e = 0; e < epochs; e++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t d = e * 64 + j * 16 + lane;
            if (is_dot) {
                a[j] = __fmaf_rn(qv[d], __ldg(x + d), a[j]);
            } else {
                const float df = __fsub_rn(qv[d], __ldg(x + d));
                a[j] = __fmaf_rn(df, df, a[j]);
            }
        }
    float tot = reduce16(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])));
    if (lane == 0 && live) {
        for (int64_t d = epochs * 64; d < dim; d++) {
            if (is_dot) {
                tot = __fmaf_rn(qv[d], __ldg(x + d), tot);
            } else {
                const float df = __fsub_rn(qv[d], __ldg(x + d));
                tot = __fmaf_rn(df, df, tot);
            }
        }
        out[p] = valid ? tot : __uint_as_float(0x7fc00000u);
    }
}

vg_status rerank_gather(const float *d_vectors, int64_t nrows, int64_t dim, const float *d_queries, int64_t nq,
                        const uint32_t *d_rows, int64_t r, int is_dot, float *d_out, cudaStream_t st) {
    const int64_t total = nq * r;
    if (total <= 0) return VG_OK;
    if (!d_vectors) return fail(VG_ERR_STATE, "index holds no float32 vectors to rerank against");
    const int64_t threads = total * 16;
    rerank_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_vectors, nrows, dim, d_queries, nq, d_rows, r, is_dot, d_out);
    VG_LAUNCHED();
    return VG_OK;
}

// Rerank against float32 rows in mapped HOST memory (vg_index_set_host_vectors).  Scalar 64-byte half-warp reads over the
// host link reach about a third of what it can carry, so the candidate rows are first pulled into an HBM staging buffer by
// a copy-only kernel (one warp per pair, 16-byte loads, every load of a row in flight before the first store), chunk by
// chunk, and the rerank kernel then reads that buffer: same arithmetic, same order, same bits.
__global__ void __launch_bounds__(256) host_rows_gather_kernel(const float *vectors, int64_t nrows, int64_t dim, const uint32_t *rows,
                                                               int64_t pair0, int64_t npairs, float *stage) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= npairs) return;
    const uint32_t row = rows[pair0 + w];
    if ((int64_t)row >= nrows) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(vectors + (int64_t)row * dim);
    uint4 *dst = reinterpret_cast<uint4 *>(stage + w * dim);
    const int n16 = (int)(dim >> 2);
    for (int i = lane; i < n16; i += 32 * 8) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (i + u * 32 < n16) v[u] = __ldcs(src + i + u * 32);
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (i + u * 32 < n16) dst[i + u * 32] = v[u];
    }
}
// rerank_kernel over a staged chunk: pair p = pair0 + local reads row `local` of the staging buffer
__global__ void __launch_bounds__(256) rerank_staged_kernel(const float *stage, int64_t nrows, int64_t dim, const float *queries,
                                                            const uint32_t *rows, int64_t r, int64_t pair0, int64_t npairs, int is_dot, float *out) {
    const int64_t local = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const bool live = local < npairs;
    const int64_t l = live ? local : npairs - 1;
    const int64_t p = pair0 + l;
    const int64_t q = p / r;
    const bool valid = (int64_t)rows[p] < nrows;
    const float *x = stage + (valid ? l : 0) * dim;
    const float *qv = queries + q * dim;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t epochs = dim >> 6;
    for (int64_t e = 0; e < epochs; e++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t d = e * 64 + j * 16 + lane;
            if (is_dot) {
                a[j] = __fmaf_rn(qv[d], x[d], a[j]);
            } else {
                const float df = __fsub_rn(qv[d], x[d]);
                a[j] = __fmaf_rn(df, df, a[j]);
            }
        }
    float tot = reduce16(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])));
    if (lane == 0 && live) {
        for (int64_t d = epochs * 64; d < dim; d++) {
            if (is_dot) {
                tot = __fmaf_rn(qv[d], x[d], tot);
            } else {
                const float df = __fsub_rn(qv[d], x[d]);
                tot = __fmaf_rn(df, df, tot);
            }
        }
        out[p] = valid ? tot : __uint_as_float(0x7fc00000u);
    }
}
vg_status rerank_gather_host(const float *d_host_alias, int64_t nrows, int64_t dim, const float *d_queries, int64_t nq,
                             const uint32_t *d_rows, int64_t r, int is_dot, float *d_out, cudaStream_t st) {
    const int64_t total = nq * r;
    if (total <= 0) return VG_OK;
    if (dim % 4 != 0)   // rows are not 16-byte aligned: read them in place
        return rerank_gather(d_host_alias, nrows, dim, d_queries, nq, d_rows, r, is_dot, d_out, st);
    const int64_t chunk = std::max<int64_t>(1024, std::min<int64_t>(total, ((int64_t)768 << 20) / (dim * 4)));
    DevBuf stage;
    VG_TRY(stage.alloc((size_t)chunk * dim * 4));
    for (int64_t p0 = 0; p0 < total; p0 += chunk) {
        const int64_t np = std::min(chunk, total - p0);
        host_rows_gather_kernel<<<(unsigned)((np * 32 + 255) / 256), 256, 0, st>>>(d_host_alias, nrows, dim, d_rows, p0, np, stage.as<float>());
        VG_LAUNCHED();
        rerank_staged_kernel<<<(unsigned)((np * 16 + 255) / 256), 256, 0, st>>>(stage.as<float>(), nrows, dim, d_queries, d_rows, r, p0, np, is_dot, d_out);
        VG_LAUNCHED();
    }
    return VG_OK;
}

// ------------------------------------------------------------ bounded L2 (gather)
// simd.SquaredL2Bounded (kernels.go:163-176, AVX-512 path registered at kernels_amd64.go:269: bounded_l2_avx512.c:19-107)
// for every (query, candidate row) pair: four 16-lane FMA accumulators over 64-dim blocks; after EVERY block the running
// total (s1+s2)+(s3+s4), summed across the lanes in hsum512's order — i+8, i+4, then the two vhaddps: (u0+u1)+(u2+u3) —
// is compared with the bound and the pair exits with (that partial total, exceeded = 1) when it is larger.  The rest goes
// in 8-wide steps (unfused diff*diff, lanes i+4, (w0+w1)+(w2+w3), added to the total) and a fused scalar tail
// (bounded_l2_avx512.s:94-114).  NaN totals never exit early (ucomiss / jbe) and end with exceeded = 0 (seta).
__device__ __forceinline__ float hsum512_order(float v) {
    v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, 8, 16));
    v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, 4, 16));
    v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, 1, 16));   // (u0+u1) in lane 0, (u2+u3) in lane 2
    v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, 2, 16));
    return v;
}
__global__ void __launch_bounds__(256) bounded_l2_kernel(const float *vectors, int64_t nrows, int64_t dim, const float *queries, int64_t nq,
                                                         const uint32_t *rows, int64_t r, const float *bounds, int64_t bound_stride, float *out,
                                                         uint8_t *exceeded) {
    const int64_t pair = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const int64_t total = nq * r;
    const bool live = pair < total;
    const int64_t p = live ? pair : total - 1;
    const int64_t q = p / r;
    const uint32_t row = rows[p];
    const bool valid = (int64_t)row < nrows;
    const float *x = vectors + (valid ? (int64_t)row : 0) * dim;
    const float *qv = queries + q * dim;
    const float bound = bounds[bound_stride ? p : q];
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    float tot = 0.0f;
    int64_t i = 0;
    bool done = false;
    for (; i + 64 <= dim; i += 64) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t d = i + j * 16 + lane;
            const float df = __fsub_rn(qv[d], __ldg(x + d));
            a[j] = __fmaf_rn(df, df, a[j]);
        }
        tot = hsum512_order(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])));
        tot = __shfl_sync(0xffffffffu, tot, 0, 16);
        if (tot > bound) {  // uniform across the half-warp
            done = true;
            break;
        }
    }
    uint8_t ex = 1;
    if (!done) {
        for (; i + 8 <= dim; i += 8) {
            float sq = 0.0f;
            if (lane < 8) {
                const float df = __fsub_rn(qv[i + lane], __ldg(x + i + lane));
                sq = __fmul_rn(df, df);
            }
            sq = __fadd_rn(sq, __shfl_down_sync(0xffffffffu, sq, 4, 16));  // w[i] = sq[i] + sq[i+4]
            sq = __fadd_rn(sq, __shfl_down_sync(0xffffffffu, sq, 1, 16));  // w0+w1 (lane 0), w2+w3 (lane 2)
            sq = __fadd_rn(sq, __shfl_down_sync(0xffffffffu, sq, 2, 16));
            sq = __shfl_sync(0xffffffffu, sq, 0, 16);
            tot = __fadd_rn(tot, sq);
        }
        for (; i < dim; i++) {
            const float df = __fsub_rn(qv[i], __ldg(x + i));
            tot = __fmaf_rn(df, df, tot);
        }
        ex = tot > bound ? 1 : 0;
    }
    if (lane == 0 && live) {
        out[p] = valid ? tot : __uint_as_float(0x7fc00000u);
        exceeded[p] = valid ? ex : 0;
    }
}
vg_status bounded_l2_gather(const float *d_vectors, int64_t nrows, int64_t dim, const float *d_queries, int64_t nq, const uint32_t *d_rows,
                            int64_t r, const float *d_bounds, int per_pair_bounds, float *d_out, uint8_t *d_exceeded, cudaStream_t st) {
    const int64_t total = nq * r;
    if (total <= 0) return VG_OK;
    if (!d_vectors) return fail(VG_ERR_STATE, "index holds no float32 vectors");
    const int64_t threads = total * 16;
    bounded_l2_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_vectors, nrows, dim, d_queries, nq, d_rows, r, d_bounds,
                                                                        per_pair_bounds ? 1 : 0, d_out, d_exceeded);
    VG_LAUNCHED();
    return VG_OK;
}

// ------------------------------------------------------------ hamming matrix
__global__ void __launch_bounds__(256) hamming_kernel(const uint8_t *q, int64_t nq, const uint8_t *codes, int64_t n,
                                                      int64_t nbytes, int32_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * n) return;
    const int64_t qi = i / n, r = i - qi * n;
    const uint8_t *a = q + qi * nbytes, *b = codes + r * nbytes;
    int h = 0;
    for (int64_t j = 0; j < nbytes; j++) h += __popc((uint32_t)(a[j] ^ b[j]));
    out[i] = h;
}
vg_status hamming_dense(const uint8_t *d_q, int64_t nq, const uint8_t *d_codes, int64_t n, int64_t nbytes, int32_t *d_out,
                        cudaStream_t st) {
    const int64_t total = nq * n;
    if (total <= 0) return VG_OK;
    hamming_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_q, nq, d_codes, n, nbytes, d_out);
    VG_LAUNCHED();
    return VG_OK;
}

// ------------------------------------------------------------ sign-query prep
// Query sign words (q[i] >= threshold → bit i, LSB first: rabitq.go:143-151,
// binary.go:138-152) and ||q|| = Sqrt(simd.Dot(q,q)) in AVX-512 order.
__global__ void __launch_bounds__(256) prep_sign_kernel(const float *queries, int64_t nq, int64_t dim, float threshold,
                                                        int words32, uint32_t *words, float *norms) {
    const int64_t hwid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const bool live = hwid < nq;
    const int64_t q = live ? hwid : nq - 1;
    const float *v = queries + q * dim;
    if (live)
        for (int w = lane; w < words32; w += 16) {
            uint32_t bits = 0;
            for (int b = 0; b < 32; b++) {
                const int64_t d = (int64_t)w * 32 + b;
                if (d < dim && v[d] >= threshold) bits |= 1u << b;
            }
            words[q * words32 + w] = bits;
        }
    if (norms) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        const int64_t epochs = dim >> 6;
        for (int64_t e = 0; e < epochs; e++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float x = v[e * 64 + j * 16 + lane];
                a[j] = __fmaf_rn(x, x, a[j]);
            }
        float tot = reduce16(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])));
        if (lane == 0 && live) {
            for (int64_t d = epochs * 64; d < dim; d++) tot = __fmaf_rn(v[d], v[d], tot);
            norms[q] = (float)sqrt((double)tot);  // simd.Sqrt: float32(math.Sqrt(float64(x)))
        }
    }
}
vg_status prep_sign_queries(const float *d_queries, int64_t nq, int64_t dim, float threshold, uint32_t *d_words,
                            float *d_norms, cudaStream_t st) {
    if (nq <= 0) return VG_OK;
    const int words32 = (int)(((dim + 63) / 64) * 2);
    const int64_t threads = nq * 16;
    prep_sign_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_queries, nq, dim, threshold, words32, d_words, d_norms);
    VG_LAUNCHED();
    return VG_OK;
}

}  // namespace vg
