// vg_quant_tc.cu — the quantized scans (SQ8, INT4 decode-and-scan, PQ/OPQ ADC) as a tcgen05 decode-GEMM *filter*
// followed by an exact re-check in the reference's own arithmetic.
//
// Why.  flat.(*Segment).Search scores every (query,row) pair of a quantized segment with
// simd.Sq8uL2BatchPerDimension / simd.Int4L2DistanceBatch / simd.PqAdcLookup
// (internal/segment/flat/segment.go:543-552,603-611; sq8_avx512.c:59-104, int4_avx512.c:193-299,
// floats_avx512.c:135-167).  All three are  ||q - decode(code)||^2  up to float32 rounding, i.e. a dense
// Q x N x d contraction  ||x^||^2 - 2 q.x^  once the row is decoded.  A CUDA-core scan that replays the reference's
// FMA order is bound by the FP32 pipe (SQ8/INT4) or by shared-memory table lookups (PQ) far below what a batch of
// 10k queries allows.  Here the codes are decoded ONCE per 256-query tile, inside the kernel, into fp16 B tiles in
// shared memory (never materialised in HBM: the database stays at 1 / 0.5 / 0.125 bytes per dimension), contracted on
// the tensor cores, and the epilogue keeps per-group minima exactly like the Flat filter (vg_flat_tc.cu).  The few
// hundred surviving rows per query are then scored in the reference's exact order (AVX-512 lane order, the same
// fma/rounding sequence, PQ table built by the generic Go loop) and a certificate proves that no other row could
// have entered the top-k; queries without a proof are re-run on the exact CUDA-core scan.  Result: ids and float32
// scores bit-identical to the reference path.
//
// Exact integer B operand.  All three codecs decode as  x^_d = mid_d + w_d * b_d  with a small signed integer b_d
// (SQ8: code - 128, INT4: nibble - 8, PQ: the int8 codebook entry) and per-dimension constants (SQ8: w = invScale,
// mid = min + 128 w; INT4: w = diff / 15, mid = min + 8 w; PQ: w = scale_m, mid = offset_m).  So
//   q.x^ = q.mid + sum_d (q_d w_d) b_d :
// the B tile holds the integers b_d as fp16 — EXACT, and one PRMT + one HSUB2 per two elements to produce — the
// weights move to the query side (a_d = fp16(q_d w_d 2^e), e chosen per query so that max|a_d| is in [2^11, 2^12)),
// and q.mid is a per-query constant c_q.  kind::f16 runs at twice the TF32 rate; only the A operand is rounded.
//
//   s'(q,x)  = ||x^||^2 - 2 sum_d (q_d w_d) b_d          (what the GEMM epilogue ranks; s = s' + c_q, c_q = -2 q.mid)
//   |s'_tc - s'| <= E = c1 ||q|| max||x^ - mid|| + c2 (||q||^2 + max(||x^||^2, ||x^ - mid||^2))
//                       + (2^-22 + G 2^-23) max|s'| + 2^-21 ||q|| (||mid|| + max||x^ - mid||)
//   c1 = 2^-10 * 1.125 (A rounded to 2^-11 relative, factor 2 of the L2 form, slack for the subnormal tail),
//   c2 = 2^-14 + d 2^-23 (norms, fp32 accumulation with truncation), the third term pays for the log2 G mantissa
//   bits that carry the row index and the epilogue's own rounding, the last for the few-ulp difference between
//   mid + w b and the reference's float32 decode (fma(c, inv, min); fma(nib * (1/15), diff, min); c * scale + offset).
// The reference's own float32 evaluation differs from the real ||q - x^||^2 by at most (d + 64) 2^-24 relative
// (non-negative terms), which is added on the exact side of the comparison.
//
// Kernel (one CTA per SM = 256 queries x a contiguous row range, 576 threads):
//   warp 0       TMA producer of the fp16 query k-blocks (256 x 64 halves, 128-byte swizzle)
//   warp 1       tcgen05.mma.cta_group::1.kind::f16 issuer, M=128 x N=128 x K=16, two M halves per B tile,
//                fp32 accumulators in TMEM (2 stages x 2 halves x 128 columns)
//   warps 2-9    epilogue: one thread = one query; s = fma(f_q, acc, ||x^||^2), group (min, second min)
//   warps 10-17  decode producers, two groups of 128 threads that alternate k-blocks; thread = row: code bytes ->
//                fp16 integers (PRMT into 0x64xx = 1024 + byte, HSUB2) -> 16-byte stores at the swizzled position of
//                the B tile, then fence.proxy.async + mbarrier arrive.  Loads for the group's next k-block are in
//                flight meanwhile.
// The code layout on the device is whatever the CUDA-core scan uses (lane-transposed SQ8/INT4, tiled PQ): the
// contraction does not care about the order of the dimensions, so the QUERY tile is permuted into storage order
// instead and the producer converts bytes in the order they are stored.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <mutex>
#include <vector>

#include "vg_flat_tc.cuh"
#include <type_traits>

#include "vg_quant_tc.cuh"
#include "vg_tiles.cuh"
#include "vg_tc_ptx.cuh"
#include "vg_topk.cuh"

namespace vg {
namespace qtc {

using namespace vg::tc;

constexpr int BM = 128;    // UMMA M
constexpr int BMQ = 256;   // queries per CTA
constexpr int BN = 128;    // rows per tile (UMMA N)
constexpr int BK = 64;     // halves per k-block: one 128-byte swizzle atom
constexpr int STAGES = 4;
constexpr int A_KB_BYTES = BMQ * BK * 2;  // 32 KB
constexpr int B_KB_BYTES = BN * BK * 2;   // 16 KB
constexpr int STAGE_BYTES = A_KB_BYTES + B_KB_BYTES;
constexpr int PROD_WARP0 = 10;            // first decode-producer warp
constexpr int NTHREADS = (PROD_WARP0 + 8) * 32;
constexpr size_t OFF_XN = (size_t)STAGES * STAGE_BYTES;
constexpr size_t OFF_BAR = OFF_XN + (size_t)2 * BN * 4;
constexpr size_t SMEM_BYTES = OFF_BAR + (size_t)(2 * STAGES + 4) * 8 + 16 + 1024;  // + slack for the 1024-byte alignment
constexpr int Q_SQ8 = 0, Q_INT4 = 1, Q_PQ = 2, Q_RABITQ = 3, Q_BQ = 4;
// GEMM-only variant of SQ8 (qtc2_kernel): kind::i8 — the raw code bytes ARE the B operand (unsigned 8-bit, loaded by TMA,
// no decode warps), the query tile is quantised to signed 8-bit per query; everything after the GEMM is Q_SQ8's.
constexpr int Q_SQ8I = 5;
// kind::i8 variants that still decode in the kernel (streamed query k-blocks, decode warps write BYTES): RaBitQ / BQ sign
// bits as exact +-1 signed bytes (128 dims per 128-byte k-block instead of 64)
constexpr int Q_RABITQI = 6, Q_BQI = 7;
// INT4 nibbles as unsigned bytes 0..15 (64 stored bytes = 128 dims per k-block), queries quantised as for Q_SQ8I
constexpr int Q_INT4I = 8;
// PQ (dsub = 8): the int8 codebook entries ARE signed bytes — 16 gathers of 8 bytes per row and k-block (128 dims), stored as they are
constexpr int Q_PQI = 9;
__host__ __device__ constexpr bool i8_codec(int c) { return c == Q_SQ8I || c == Q_RABITQI || c == Q_BQI || c == Q_INT4I || c == Q_PQI; }
__host__ __device__ constexpr int base_codec(int c) {
    return c == Q_SQ8I ? Q_SQ8 : c == Q_RABITQI ? Q_RABITQ : c == Q_BQI ? Q_BQ : c == Q_INT4I ? Q_INT4 : c == Q_PQI ? Q_PQ : c;
}
// sign-bit codes (RaBitQ, BQ): the B tile is +-1, the GEMM is exact (acc = D - 2 Hamming)
__host__ __device__ constexpr bool sign_codec(int c) { return c == Q_RABITQ || c == Q_BQ; }
constexpr int LIST_CAP = 8192;            // candidate rows per query in the exact stage (work bound; beyond it the query goes to the exact scan)

// kind::f16 instruction descriptor: D = f32 (bits 4-5 = 1), A = B = F16 (0 at bits 7-9 / 10-12), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// two fp16 integers from two bytes: 0x6400 | u is the half 1024 + u (exact), minus `bias` (packed halves) is exact too
__device__ __forceinline__ uint32_t hsub2_bits(uint32_t h2, uint32_t bias2) {
    uint32_t r;
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(h2), "r"(bias2));
    return r;
}
constexpr uint32_t H2_1152 = 0x64806480u;  // (1152, 1152): byte - 128
constexpr uint32_t H2_1032 = 0x64086408u;  // (1032, 1032): nibble - 8
// bytes 0,1 / 2,3 of w as (1024 + byte) halves
__device__ __forceinline__ uint32_t bytes01_h2(uint32_t w) { return __byte_perm(w, 0x64646464u, 0x4140); }
__device__ __forceinline__ uint32_t bytes23_h2(uint32_t w) { return __byte_perm(w, 0x64646464u, 0x4342); }

struct KArgs {
    const float *xn;        // [rows] ||x^||^2
    const uint32_t *mask;   // optional row bitmap as 32-bit words
    // tile skipping (pair kernel, with a row bitmap): the 256-row tiles that hold at least one allowed row, in ascending
    // order, and their number; the splits share the list evenly.  nullptr = every tile of the split's row range.
    const int32_t *tile_list;
    const int32_t *tile_count;
    const float *fq;        // [nq] -2 / query scale
    const int32_t *asum;    // kind::i8 sign codecs: [nq] sum of the query's +-1 bytes (B holds the bits 0 / 1), else nullptr
    int64_t nq, rows, rows_per_split;
    float half_dim;           // BQ: D / 2 (Hamming = D / 2 - acc / 2)
    int kb;                 // k-blocks = dimp / 64
    int cpg;                // 32-row chunks per group
    float2 *mins;           // [nq_pad][groups]
    int64_t groups;
    uint32_t idx_mask;
    uint32_t keep_hi;       // ~31u (passed in so that (bits & keep_hi) | j compiles to one LOP3 with an immediate j)
    // decode
    const uint8_t *codes;
    int64_t row_bytes;
    const int8_t *codebooks;
    int dsub_shift;         // PQ: log2(dsub)
    int tiled;              // PQ: codes stored in 32-row tiles (permute_pq)
    // threshold-collect epilogue (THRESH kernels): rows with s' < Ts[q] go to the (query, split, column half) list
    const float *Ts;        // [nq] thresholds in s'-space
    uint2 *cand;            // [nq][slots][cap] (s' bits, local row)
    int *ccnt;              // [nq][slots]
    int *ovf;               // [nq] set when a list overflowed
    int cap, slots;
};

// ------------------------------------------------------------------ decode producers
template <int CODEC>
struct Producer;
template <int CODEC>
struct ProducerBytes;

// SQ8 / INT4: BPR stored bytes of every row per k-block (64 / 32).  A warp fills a slab of 32 tile rows; its global
// loads are arranged so that LPR = BPR/16 neighbouring lanes read the contiguous BPR bytes of one row and one 16-byte
// load instruction covers 32/LPR rows (8 / 16 cache lines per instruction instead of 32): thread (lane) holds, for
// j < LPR, the 16-byte piece (lane % LPR) of slab row  j * (32/LPR) + lane / LPR.
template <int CODEC>
struct ProducerBytes {
    static constexpr int BPR = (CODEC == Q_SQ8 || CODEC == Q_INT4I) ? 64 : 32;
    static constexpr int LPR = BPR / 16;
    static constexpr int RPI = 32 / LPR;  // rows per load instruction
    uint4 w[LPR];
    // Row pointers of the thread's LPR pieces for one tile (rows past the end are clamped; masked later by xn = BIG).
    struct Rows {
        const uint4 *p[LPR];
        __device__ __forceinline__ void set(const KArgs &A, int64_t slab_row0, int lane) {
#pragma unroll
            for (int j = 0; j < LPR; j++) {
                int64_t row = slab_row0 + j * RPI + lane / LPR;
                row = row < A.rows ? row : A.rows - 1;
                p[j] = reinterpret_cast<const uint4 *>(A.codes + row * A.row_bytes) + (lane % LPR);
            }
        }
    };
    __device__ __forceinline__ void fetch(const Rows &R, int kb) {
#pragma unroll
        for (int j = 0; j < LPR; j++) w[j] = __ldg(R.p[j] + kb * (BPR / 16));
    }
    // b_tile: shared address of the 128 x 128-byte B tile; slab: first tile row of this warp's slab
    __device__ __forceinline__ void convert(uint32_t b_tile, int slab, int lane) const {
        const int piece = lane % LPR;
#pragma unroll
        for (int j = 0; j < LPR; j++) {
            const int r = slab + j * RPI + lane / LPR;
            const uint32_t dst = b_tile + (uint32_t)r * 128u;
            const int swz = r & 7;
            const uint32_t ww[4] = {w[j].x, w[j].y, w[j].z, w[j].w};
            if constexpr (CODEC == Q_INT4I) {
                // kind::i8: 16 stored bytes = 32 nibbles -> 32 unsigned bytes = chunks 2*piece, 2*piece + 1; word h of the piece
                // gives the two words (low nibbles of its four bytes, high nibbles of its four bytes)
#pragma unroll
                for (int h = 0; h < 2; h++)
                    sts128(dst + (uint32_t)(((2 * piece + h) ^ swz) << 4), ww[2 * h] & 0x0F0F0F0Fu, (ww[2 * h] >> 4) & 0x0F0F0F0Fu,
                           ww[2 * h + 1] & 0x0F0F0F0Fu, (ww[2 * h + 1] >> 4) & 0x0F0F0F0Fu);
            } else if constexpr (CODEC == Q_SQ8) {
                // 16 bytes = chunks 2*piece, 2*piece + 1; code - 128
#pragma unroll
                for (int h = 0; h < 2; h++)
                    sts128(dst + (uint32_t)(((2 * piece + h) ^ swz) << 4), hsub2_bits(bytes01_h2(ww[2 * h]), H2_1152),
                           hsub2_bits(bytes23_h2(ww[2 * h]), H2_1152), hsub2_bits(bytes01_h2(ww[2 * h + 1]), H2_1152),
                           hsub2_bits(bytes23_h2(ww[2 * h + 1]), H2_1152));
            } else {
                // 16 bytes = 4 words = chunks 4*piece .. 4*piece + 3; one word = one chunk, element order
                // b0.lo b2.lo b0.hi b2.hi b1.lo b3.lo b1.hi b3.hi (bytes b0..b3 of the word); nibble - 8
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const uint32_t v = ww[h];
                    sts128(dst + (uint32_t)(((4 * piece + h) ^ swz) << 4), hsub2_bits((v & 0x000F000Fu) | 0x64006400u, H2_1032),
                           hsub2_bits(((v >> 4) & 0x000F000Fu) | 0x64006400u, H2_1032),
                           hsub2_bits(((v >> 8) & 0x000F000Fu) | 0x64006400u, H2_1032),
                           hsub2_bits(((v >> 12) & 0x000F000Fu) | 0x64006400u, H2_1032));
                }
            }
        }
    }
};

// PQ: per 16-byte chunk (8 dims) one 8-byte gather from the int8 codebook of its subspace -> the int8 value.
template <>
struct Producer<Q_PQ> {
    uint2 g[8];
    __device__ __forceinline__ static uint2 load_codes(const KArgs &A, int64_t row, int kb) {
        const int mb = ((kb * 64) >> A.dsub_shift) & ~7;  // 8 codes that cover this k-block's subspaces
        const uint8_t *p = A.tiled ? A.codes + (row >> 5) * (32 * A.row_bytes) + (int64_t)(mb >> 4) * 512 + (row & 31) * 16 + (mb & 15)
                                   : A.codes + row * A.row_bytes + mb;
        return __ldg(reinterpret_cast<const uint2 *>(p));
    }
    __device__ __forceinline__ void gather(const KArgs &A, uint2 c8, int kb) {
        const unsigned long long cw = ((unsigned long long)c8.y << 32) | c8.x;
        const int mb = ((kb * 64) >> A.dsub_shift) & ~7;
        const int dsub = 1 << A.dsub_shift;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int d0 = kb * 64 + 8 * c;
            const int m = d0 >> A.dsub_shift, o = d0 & (dsub - 1);
            const uint32_t code = (uint32_t)(cw >> (8 * (m - mb))) & 0xFFu;
            g[c] = __ldg(reinterpret_cast<const uint2 *>(A.codebooks + (((int64_t)m * 256 + code) << A.dsub_shift) + o));
        }
    }
    // same gather from a 16 KB codebook slice in shared memory (the subspaces of k-block kb start at byte 0 of the slice)
    __device__ __forceinline__ void gather_smem(const KArgs &A, uint2 c8, int kb, uint32_t slice) {
        const unsigned long long cw = ((unsigned long long)c8.y << 32) | c8.x;
        const int m0 = (kb * 64) >> A.dsub_shift;
        const int mb = m0 & ~7;
        const int dsub = 1 << A.dsub_shift;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int d0 = kb * 64 + 8 * c;
            const int m = d0 >> A.dsub_shift, o = d0 & (dsub - 1);
            const uint32_t code = (uint32_t)(cw >> (8 * (m - mb))) & 0xFFu;
            const uint32_t addr = slice + ((((uint32_t)(m - m0) << 8) + code) << A.dsub_shift) + (uint32_t)o;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(g[c].x), "=r"(g[c].y) : "r"(addr));
        }
    }
    __device__ __forceinline__ void convert(const KArgs &, int, uint32_t dst_row, int swz) const {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const uint32_t w0 = g[c].x ^ 0x80808080u, w1 = g[c].y ^ 0x80808080u;  // int8 + 128 as unsigned bytes
            sts128(dst_row + (uint32_t)((c ^ swz) << 4), hsub2_bits(bytes01_h2(w0), H2_1152), hsub2_bits(bytes23_h2(w0), H2_1152),
                   hsub2_bits(bytes01_h2(w1), H2_1152), hsub2_bits(bytes23_h2(w1), H2_1152));
        }
    }
};

// RaBitQ sign bits: 8 stored bytes (64 dims) per k-block -> +1 / -1 as fp16, exact.  One 32-bit word = four 16-byte
// chunks; half2 p (0..15) of a word holds bit p (low half) and bit p + 16 (high half): (~w << (15 - p)) puts the two
// inverted bits on the sign positions, OR 0x3C003C00 makes them -1.0 / +1.0 (bit set = component >= 0 = +1).
template <>
struct Producer<Q_RABITQ> {
    uint2 w;
    __device__ __forceinline__ void fetch(const KArgs &A, int64_t row, int kb) {
        w = __ldg(reinterpret_cast<const uint2 *>(A.codes + row * A.row_bytes + (int64_t)kb * 8));
    }
    __device__ __forceinline__ void convert(const KArgs &, int, uint32_t dst_row, int swz) const {
        const uint32_t y[2] = {~w.x, ~w.y};
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const uint32_t v = y[c >> 2];
            const int p0 = (c & 3) * 4;
            uint32_t h[4];
#pragma unroll
            for (int j = 0; j < 4; j++) h[j] = ((v << (15 - (p0 + j))) & 0x80008000u) | 0x3C003C00u;
            sts128(dst_row + (uint32_t)((c ^ swz) << 4), h[0], h[1], h[2], h[3]);
        }
    }
};

// The same for kind::i8: 16 stored bytes (128 dims) per k-block -> the bits themselves as UNSIGNED bytes 0 / 1 in natural
// dimension order.  Four bits n -> four bytes: (n * 0x00204081) & 0x01010101 puts bit i into byte i (the four shifted copies
// of a 4-bit n do not overlap, so nothing carries).  With a = +-1 on the query side, sum a s_x = 2 sum a bit - sum a: the
// epilogue forms 2 acc - asum with one IMAD (FMA pipe) — cheaper than mapping the bytes to +-1 here (two more ALU operations
// per four bits; the decode warps were two thirds of a 70 % busy ALU pipe).
struct ProducerSignI8 {
    uint4 w;
    __device__ __forceinline__ void fetch(const KArgs &A, int64_t row, int kb) {
        w = __ldg(reinterpret_cast<const uint4 *>(A.codes + row * A.row_bytes + (int64_t)kb * 16));
    }
    __device__ __forceinline__ void convert(const KArgs &, int, uint32_t dst_row, int swz) const {
        const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const uint32_t v = wv[c >> 1] >> ((c & 1) * 16);
            uint32_t h[4];
#pragma unroll
            for (int j = 0; j < 4; j++) h[j] = (((v >> (4 * j)) & 0xFu) * 0x00204081u) & 0x01010101u;
            sts128(dst_row + (uint32_t)((c ^ swz) << 4), h[0], h[1], h[2], h[3]);
        }
    }
};

// PQ for kind::i8 (dsub = 8): the 16 codes of a k-block (subspaces 16 kb .. 16 kb + 15 = 128 dims) in one 16-byte load, 16
// gathers of one centroid (8 signed bytes) from the 32 KB codebook slice in shared memory, stored unchanged: chunk c of
// the B row = centroids of subspaces 2c, 2c + 1 — natural dimension order.
struct ProducerPQI8 {
    using Codes = uint4;
    uint2 g[16];
    __device__ __forceinline__ static Codes zero() { return make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ static Codes load_codes(const KArgs &A, int64_t row, int kb) {
        const uint8_t *p = A.tiled ? A.codes + (row >> 5) * (32 * A.row_bytes) + (int64_t)kb * 512 + (row & 31) * 16
                                   : A.codes + row * A.row_bytes + (int64_t)kb * 16;
        return __ldg(reinterpret_cast<const uint4 *>(p));
    }
    __device__ __forceinline__ void gather_smem(const KArgs &, Codes c16, int, uint32_t slice) {
        const uint32_t cw[4] = {c16.x, c16.y, c16.z, c16.w};
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const uint32_t code = (cw[c >> 2] >> (8 * (c & 3))) & 0xFFu;
            const uint32_t addr = slice + ((((uint32_t)c << 8) + code) << 3);
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(g[c].x), "=r"(g[c].y) : "r"(addr));
        }
    }
    __device__ __forceinline__ void convert(const KArgs &, int, uint32_t dst_row, int swz) const {
#pragma unroll
        for (int c = 0; c < 8; c++) sts128(dst_row + (uint32_t)((c ^ swz) << 4), g[2 * c].x, g[2 * c].y, g[2 * c + 1].x, g[2 * c + 1].y);
    }
};
struct ProducerPQF16 : Producer<Q_PQ> {
    using Codes = uint2;
    __device__ __forceinline__ static Codes zero() { return make_uint2(0u, 0u); }
};

// ------------------------------------------------------------------ GEMM + group minima
template <int CODEC>
__global__ void __launch_bounds__(NTHREADS, 1) qtc_kernel(const __grid_constant__ CUtensorMap map_q, KArgs A) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * BMQ;
    const int split = blockIdx.y;
    const int64_t row_begin = (int64_t)split * A.rows_per_split;
    int64_t row_end = row_begin + A.rows_per_split;
    if (row_end > A.rows) row_end = A.rows;
    const int ntiles = row_end > row_begin ? (int)((row_end - row_begin + BN - 1) / BN) : 0;

    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar0 = s_base + (uint32_t)OFF_BAR;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_bar(s), 1 + 4);   // TMA (query k-block) + the four warps of one decode group
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 256);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const int total_it = ntiles * A.kb;

    if (warp == 0) {
        // ===================== TMA producer: query k-blocks =====================
        if (lane == 0) {
            for (int it = 0; it < total_it; it++) {
                const int st = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int kb = it % A.kb;
                mbar_wait(empty_bar(st), ph ^ 1);
                mbar_expect_tx(full_bar(st), A_KB_BYTES);
                tma_load_2d(s_base + st * STAGE_BYTES, &map_q, kb * BK, q0, full_bar(st));
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(BN);
            int it = 0;
            for (int t = 0; t < ntiles; t++) {
                const int as = t & 1;
                const uint32_t aph = (t >> 1) & 1;
                mbar_wait(tempty_bar(as), aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * 2 * BN);
                for (int kb = 0; kb < A.kb; kb++, it++) {
                    const int st = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(full_bar(st), ph);
                    tc_fence_after();
                    const uint32_t sa = s_base + st * STAGE_BYTES;
                    const uint64_t bdesc = make_sdesc(sa + A_KB_BYTES);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint64_t adesc = make_sdesc(sa + h * (BM * BK * 2));
#pragma unroll
                        for (int k = 0; k < BK / 16; k++)  // 16 halves = 32 bytes per UMMA
                            umma_f16(d_tmem + (uint32_t)(h * BN), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                     (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar(st));
                }
                umma_commit(tfull_bar(as));
            }
        }
    } else if (warp < PROD_WARP0) {
        // ===================== epilogue: warps 2..9, one thread per query =====================
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int slot = half * BM + quad * 32 + lane;
        const int et = (warp - 2) * 32 + lane;
        const int64_t q = (int64_t)q0 + slot;
        float *xs = reinterpret_cast<float *>(smem + OFF_XN);
        const float BIG = 3.0e38f;
        const float fq = q < A.nq ? __ldg(A.fq + q) : 0.0f;
        float g1 = BIG, g2 = BIG;
        int cc = 0;
        const uint32_t keep_hi = A.keep_hi;  // 0xFFFFFFE0 as a run-time value: stays a register operand of the LOP3s below
        for (int t = 0; t < ntiles; t++) {
            const int as = t & 1;
            const uint32_t aph = (t >> 1) & 1;
            const int64_t n0 = row_begin + (int64_t)t * BN;
            float *xt = xs + as * BN;
            if (et < BN) {
                const int64_t row = n0 + et;
                xt[et] = (row < row_end) ? __ldg(A.xn + row) : BIG;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 2 * BN + half * BN);
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)(c * 32), v);
                uint32_t mw = 0xFFFFFFFFu;
                if (A.mask) mw = (n0 + c * 32 < A.rows) ? __ldg(A.mask + ((n0 + c * 32) >> 5)) : 0u;
                tmem_ld_wait();
                float s[32];
                const float4 *x4 = reinterpret_cast<const float4 *>(xt + c * 32);
#pragma unroll
                for (int j4 = 0; j4 < 8; j4++) {
                    const float4 xv = x4[j4];
                    const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) s[j4 * 4 + i] = __fmaf_rn(fq, __uint_as_float(v[j4 * 4 + i]), xx[i]);
                }
                if (mw != 0xFFFFFFFFu) {
#pragma unroll
                    for (int j = 0; j < 32; j++) s[j] = (mw >> j) & 1u ? s[j] : BIG;
                }
                // the row's position inside the 32-row chunk goes into the low 5 mantissa bits (one LOP3 with an immediate);
                // the chunk's position inside the group is added to the chunk minimum afterwards (group sizes > 32)
                float a1[4] = {BIG, BIG, BIG, BIG}, a2[4] = {BIG, BIG, BIG, BIG};
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    uint32_t vb;  // (bits & ~31) | j as ONE LOP3 (register mask, immediate j)
                    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(vb) : "r"(__float_as_uint(s[j])), "r"(keep_hi), "r"((uint32_t)j));
                    const float v1 = __uint_as_float(vb);
                    a2[j & 3] = fminf(a2[j & 3], fmaxf(a1[j & 3], v1));
                    a1[j & 3] = fminf(a1[j & 3], v1);
                }
                const float p1 = fminf(a1[0], a1[1]), p2 = fminf(fmaxf(a1[0], a1[1]), fminf(a2[0], a2[1]));
                const float r1 = fminf(a1[2], a1[3]), r2 = fminf(fmaxf(a1[2], a1[3]), fminf(a2[2], a2[3]));
                const uint32_t cidx = (uint32_t)(n0 + c * 32) & A.idx_mask & ~31u;
                const float c1 = __uint_as_float((__float_as_uint(fminf(p1, r1)) & ~A.idx_mask) | cidx | (__float_as_uint(fminf(p1, r1)) & 31u));
                const float c2 = fminf(fmaxf(p1, r1), fminf(p2, r2));
                g2 = fminf(fmaxf(g1, c1), fminf(g2, c2));
                g1 = fminf(g1, c1);
                if (++cc == A.cpg) {
                    const int64_t gid = (n0 + c * 32) / (32 * (int64_t)A.cpg);
                    if (gid < A.groups) A.mins[q * A.groups + gid] = make_float2(g1, g2);
                    g1 = BIG;
                    g2 = BIG;
                    cc = 0;
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(as));
        }
        if (cc > 0 && ntiles > 0) {
            const int64_t last_chunk_row = row_begin + (int64_t)ntiles * BN - 32;
            const int64_t gid = last_chunk_row / (32 * (int64_t)A.cpg);
            if (gid < A.groups) A.mins[q * A.groups + gid] = make_float2(g1, g2);
        }
    } else {
        // ===================== decode producers: warps 10..17, two groups alternating k-blocks =====================
        const int grp = (warp - PROD_WARP0) >> 2;
        // (tile, k-block) cursor of iteration `it`, advanced by two iterations at a time without divisions
        struct Cursor {
            int t, kb;
            __device__ __forceinline__ void init(int it, int KB) {
                t = it / KB;
                kb = it - t * KB;
            }
            __device__ __forceinline__ bool advance2(int KB) {  // returns true when the tile changed
                kb += 2;
                bool moved = false;
                while (kb >= KB) {
                    kb -= KB;
                    t++;
                    moved = true;
                }
                return moved;
            }
        };
        if constexpr (CODEC == Q_PQ) {
            const int r = ((warp - PROD_WARP0) & 3) * 32 + lane;  // row of the B tile this thread fills
            const int swz = r & 7;
            auto row_of = [&](int t) {
                const int64_t row = row_begin + (int64_t)t * BN + r;
                return row < A.rows ? row : A.rows - 1;  // padding rows of the last tile: any valid row (masked by xn = BIG)
            };
            Producer<Q_PQ> cur, nxt;
            uint2 cnn = make_uint2(0u, 0u);
            Cursor c0, c2, c4;  // iterations it, it + 2, it + 4
            c0.init(grp, A.kb);
            c2 = c0;
            c2.advance2(A.kb);
            c4 = c2;
            c4.advance2(A.kb);
            if (grp < total_it) nxt.gather(A, Producer<Q_PQ>::load_codes(A, row_of(c0.t), c0.kb), c0.kb);
            if (grp + 2 < total_it) cnn = Producer<Q_PQ>::load_codes(A, row_of(c2.t), c2.kb);
            for (int it = grp; it < total_it; it += 2) {
                cur = nxt;
                if (it + 2 < total_it) nxt.gather(A, cnn, c2.kb);
                if (it + 4 < total_it) cnn = Producer<Q_PQ>::load_codes(A, row_of(c4.t), c4.kb);
                const int st = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(empty_bar(st), ph ^ 1);
                cur.convert(A, c0.kb, s_base + st * STAGE_BYTES + A_KB_BYTES + r * 128, swz);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(st));
                c0 = c2;
                c2 = c4;
                c4.advance2(A.kb);
            }
        } else {
            const int slab = ((warp - PROD_WARP0) & 3) * 32;  // this warp fills tile rows slab .. slab + 31
            ProducerBytes<CODEC> cur, nxt;
            typename ProducerBytes<CODEC>::Rows rows;
            Cursor c2;  // iteration it + 2 (the one being prefetched)
            c2.init(grp, A.kb);
            rows.set(A, row_begin + (int64_t)c2.t * BN + slab, lane);
            if (grp < total_it) nxt.fetch(rows, c2.kb);
            for (int it = grp; it < total_it; it += 2) {
                cur = nxt;
                if (c2.advance2(A.kb)) rows.set(A, row_begin + (int64_t)c2.t * BN + slab, lane);
                if (it + 2 < total_it) nxt.fetch(rows, c2.kb);
                const int st = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(empty_bar(st), ph ^ 1);
                cur.convert(s_base + st * STAGE_BYTES + A_KB_BYTES, slab, lane);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(st));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// A cluster of two CTAs (one TPC) shares every MMA: tcgen05.mma.cta_group::2 with M = 256 (CTA r owns queries
// q0 + 128 r .. + 127 and their accumulators: 128 TMEM lanes x 256 columns) and N = 256 (CTA r decodes rows
// n0 + 128 r .. + 127 of the 256-row tile into ITS shared memory; the hardware exchanges the B halves).  Per SM the
// operand traffic of one instruction is 4 KB of A + 4 KB of B for twice the FLOPs of the single-CTA M=128 x N=128
// instruction, which is what lifts the shared-memory bound of the single-CTA kernel, and each CTA fetches only its
// 128-query half of the query tile from L2.  Stage = 16 KB A + 16 KB B, six stages.
//   full[s]   (leader CTA): leader's arrive.expect_tx (both CTAs' TMA bytes land here) + 4 producer warps of each CTA
//   empty[s]  (both CTAs) : tcgen05.commit multicast
//   tfull[a]  (both CTAs) : tcgen05.commit multicast;  tempty[a] (leader): 8 epilogue warps of each CTA
// Row groups are at most 128 rows so that a group lives in one thread (thread = query x 128-column half).
namespace pair {
using namespace vg::tc::pair;  // cluster / cta_group::2 PTX helpers (vg_tc_ptx.cuh)
constexpr int STAGES2 = 6;
constexpr int A2_BYTES = BM * BK * 2;   // 16 KB: this CTA's 128 queries x 64 halves
constexpr int B2_BYTES = BN * BK * 2;   // 16 KB: this CTA's 128 rows x 64 halves
constexpr int STAGE2_BYTES = A2_BYTES + B2_BYTES;
constexpr int TILE_ROWS = 2 * BN;       // 256 rows per pair tile
constexpr size_t OFF_XN2 = (size_t)STAGES2 * STAGE2_BYTES;
constexpr size_t OFF_BAR2 = OFF_XN2 + (size_t)4 * TILE_ROWS * 4;   // row-norm staging: two buffers per epilogue group
constexpr size_t SMEM2_BYTES = OFF_BAR2 + (size_t)40 * 8 + 16 + 1024;   // 40 mbarrier slots (kind::i8 uses 20..36)

}  // namespace pair

// THRESH = false: the epilogue keeps (min, second min) of every row group and writes the [queries][groups] plane the
// selection kernel reads (the normal filter).  THRESH = true: the epilogue compares the minimum of every 32-row chunk with
// the query's threshold T and appends the rows below T to a candidate list — every row with s' < T is then scored
// exactly, so with T = (a known upper bound of the k-th best score) + E the result needs no certificate.  This is the
// second pass for queries whose certificate failed (tightly clustered data: thousands of rows within E of the k-th
// best), where "a few more candidate groups" cannot help.
template <int CODEC, bool THRESH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
qtc2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, KArgs A) {
    using namespace pair;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int q0 = (int)(blockIdx.x >> 1) * BMQ + (int)rank * BM;  // this CTA's 128 queries
    const int split = blockIdx.y;
    const int64_t row_begin = (int64_t)split * A.rows_per_split;
    int64_t row_end = row_begin + A.rows_per_split;
    if (row_end > A.rows) row_end = A.rows;
    int ntiles = row_end > row_begin ? (int)((row_end - row_begin + TILE_ROWS - 1) / TILE_ROWS) : 0;
    const int32_t *tiles = nullptr;  // this split's slice of the active-tile list (tile skipping)
    if (A.tile_list) {
        const int n_act = __ldg(A.tile_count);
        const int per = (n_act + (int)gridDim.y - 1) / (int)gridDim.y;
        const int first = split * per;
        ntiles = n_act - first < per ? n_act - first : per;
        if (ntiles < 0) ntiles = 0;
        tiles = A.tile_list + first;
        row_end = A.rows;
    }
    // first row of the split's t-th tile
    auto tile_row0 = [&](int t) -> int64_t {
        return tiles ? (int64_t)__ldg(tiles + t) * TILE_ROWS : row_begin + (int64_t)t * TILE_ROWS;
    };

    // PQ trades one pipeline stage for a two-slot ring of 16 KB codebook slices (the int8 centroids of the subspaces of one
    // k-block, bulk-copied by the TMA thread): the decode warps gather from shared memory instead of from L2.
    // (kind::i8 PQ: k-blocks of 128 dims = 32 KB slices, four stages)
    constexpr bool IS_PQ = CODEC == Q_PQ || CODEC == Q_PQI;
    constexpr int NST = CODEC == Q_PQ ? STAGES2 - 1 : CODEC == Q_PQI ? STAGES2 - 2 : STAGES2;
    constexpr uint32_t SLICE_BYTES = CODEC == Q_PQI ? 32768 : 16384;  // 64 (128) dims x 256 centroids x 1 byte, whatever dsub is
    constexpr uint32_t OFF_SLICE = (uint32_t)NST * STAGE2_BYTES;
    constexpr uint32_t OFF_XN = OFF_SLICE + (IS_PQ ? 2 * SLICE_BYTES : 0);
    constexpr uint32_t OFF_BAR = OFF_XN + 4 * TILE_ROWS * 4;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar0 = s_base + OFF_BAR;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (NST + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * NST + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * NST + 2 + s); };
    auto sfull_bar = [&](int s) { return bar0 + 8u * (2 * NST + 4 + s); };
    auto sempty_bar = [&](int s) { return bar0 + 8u * (2 * NST + 6 + s); };
    constexpr uint32_t TMEM_COLS = 512;  // 2 accumulator stages x 256 columns
    // kind::i8 (SQ8I): the CTA's 128 queries x dim bytes stay RESIDENT in shared memory (kb x 16 KB at offset 0, loaded once)
    // and only the code tiles stream: nstb stages of 16 KB behind them, inside the 192 KB the other codecs use for their
    // six 32 KB stages.  Reloading the query k-blocks with every row tile doubled the L2 -> SM traffic (ncu: 10.8 TB/s,
    // tensor pipe 69 % active).
    const int nstb = 12 - A.kb < 8 ? 12 - A.kb : 8;
    auto bfull_bar = [&](int s) { return bar0 + 8u * (20 + s); };
    auto bempty_bar = [&](int s) { return bar0 + 8u * (28 + s); };
    const uint32_t afull_bar = bar0 + 8u * 36;
    const uint32_t b_base = s_base + (uint32_t)A.kb * A2_BYTES;
    // INT4I keeps the query tile resident as well when it fits (dim <= 1024): its decode warps then fill the same B-only stages
    // (streamed query k-blocks + decoded B + MMA operand reads saturated the shared-memory port: C2b 75 ms against SQ8's 65)
    const bool res_a = CODEC == Q_SQ8I || (CODEC == Q_INT4I && A.kb <= 8);

    constexpr int BASE = base_codec(CODEC);
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < NST; s++) {
            mbar_init(full_bar(s), 1 + 4 + 4);  // leader's expect_tx arrive + one decode group (4 warps) of each CTA
            mbar_init(empty_bar(s), 1);
        }
        if constexpr (CODEC == Q_SQ8I || CODEC == Q_INT4I) {
            for (int s = 0; s < 8; s++) {
                // SQ8I: the leader's expect_tx arrive (both CTAs' TMA bytes land here); INT4I: one decode group of each CTA
                mbar_init(bfull_bar(s), CODEC == Q_SQ8I ? 1 : 4 + 4);
                mbar_init(bempty_bar(s), 1);
            }
            mbar_init(afull_bar, 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 8 + 8);    // epilogue warps of both CTAs
            mbar_init(sfull_bar(s), 1);         // codebook slice landed (PQ)
            mbar_init(sempty_bar(s), 4);        // the four warps of decode group s are done gathering from it
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers are initialised before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const int total_it = ntiles * A.kb;

    if (warp == 0 && res_a) {
        // ===================== TMA producer (kind::i8): the query tile once, then this CTA's half of every code tile =========
        if (lane == 0) {
            if (total_it > 0) {
                if (leader) mbar_expect_tx(afull_bar, 2u * A2_BYTES * (uint32_t)A.kb);
                for (int kb = 0; kb < A.kb; kb++) tma_load_2d_pair(s_base + kb * A2_BYTES, &map_q, kb * 128, q0, afull_bar);
            }
            for (int it = 0; CODEC == Q_SQ8I && it < total_it; it++) {   // INT4I: the decode warps fill the B stages
                const int st = it % nstb;
                const uint32_t ph = (it / nstb) & 1;
                const int t = it / A.kb, kb = it - t * A.kb;
                mbar_wait(bempty_bar(st), ph ^ 1);
                if (leader) mbar_expect_tx(bfull_bar(st), 2 * B2_BYTES);
                const int n0 = (int)(tile_row0(t) + (int64_t)rank * BN);
                tma_load_2d_pair(b_base + st * B2_BYTES, &map_x, kb * 128, n0, bfull_bar(st));
            }
        }
    } else if (warp == 0) {
        // ===================== TMA producer: this CTA's half of the query k-block =====================
        if (lane == 0) {
            for (int it = 0; it < total_it; it++) {
                const int st = it % NST;
                const uint32_t ph = (it / NST) & 1;
                const int kb = it % A.kb;
                if constexpr (IS_PQ) {  // slice of iteration `it` into slot it % 2 (= the decode group that handles it)
                    const int slot = it & 1;
                    mbar_wait(sempty_bar(slot), ((uint32_t)(it >> 1) & 1) ^ 1);
                    mbar_expect_tx(sfull_bar(slot), SLICE_BYTES);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     s_base + OFF_SLICE + (uint32_t)slot * SLICE_BYTES),
                                 "l"(A.codebooks + (int64_t)kb * SLICE_BYTES), "r"(SLICE_BYTES), "r"(sfull_bar(slot))
                                 : "memory");
                }
                mbar_wait(empty_bar(st), ph ^ 1);
                if (leader) mbar_expect_tx(full_bar(st), 2 * A2_BYTES);  // both CTAs' loads are counted on the leader's barrier
                tma_load_2d_pair(s_base + st * STAGE2_BYTES, &map_q, kb * (i8_codec(CODEC) ? 128 : BK), q0, full_bar(st));
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            // kind::i8: signed A (query side); B = unsigned code bytes / nibbles / sign bits, or the signed int8 centroids (PQI: b_format bit 10)
            constexpr uint32_t idesc = CODEC == Q_PQI    ? (make_idesc_i8_pair() | (1u << 10))   // signed B: the int8 centroids
                                       : i8_codec(CODEC) ? make_idesc_i8_pair()                 // unsigned B: code bytes, nibbles, sign bits
                                                         : make_idesc_f16_pair();
            int it = 0;
            if (res_a) {
                if (ntiles > 0) {
                    mbar_wait_cluster(afull_bar, 0);
                    tc_fence_after();
                }
                for (int t = 0; t < ntiles; t++) {
                    const int as = t & 1;
                    const uint32_t aph = (t >> 1) & 1;
                    mbar_wait_cluster(tempty_bar(as), aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(as * TILE_ROWS);
                    for (int kb = 0; kb < A.kb; kb++, it++) {
                        const int st = it % nstb;
                        const uint32_t ph = (it / nstb) & 1;
                        mbar_wait_cluster(bfull_bar(st), ph);
                        tc_fence_after();
                        const uint64_t adesc = make_sdesc(s_base + kb * A2_BYTES), bdesc = make_sdesc(b_base + st * B2_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; k++)  // 32 bytes of K per instruction
                            umma_i8_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_pair(bempty_bar(st));
                    }
                    umma_commit_pair(tfull_bar(as));
                }
            } else
            for (int t = 0; t < ntiles; t++) {
                const int as = t & 1;
                const uint32_t aph = (t >> 1) & 1;
                mbar_wait_cluster(tempty_bar(as), aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * TILE_ROWS);
                for (int kb = 0; kb < A.kb; kb++, it++) {
                    const int st = it % NST;
                    const uint32_t ph = (it / NST) & 1;
                    mbar_wait_cluster(full_bar(st), ph);
                    tc_fence_after();
                    const uint32_t sa = s_base + st * STAGE2_BYTES;
                    const uint64_t adesc = make_sdesc(sa), bdesc = make_sdesc(sa + A2_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; k++) {  // 32 bytes of K per instruction: 16 halves or 32 bytes
                        if constexpr (i8_codec(CODEC)) umma_i8_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                        else umma_f16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_pair(empty_bar(st));
                }
                umma_commit_pair(tfull_bar(as));
            }
        }
    } else if (warp < PROD_WARP0 || CODEC == Q_SQ8I) {
        // ===================== epilogue: warps 2..9; thread = one query x one 128-column half of the tile =====================
        // kind::i8 has no decode warps and an MMA twice as fast: warps 10..17 are a SECOND epilogue group; group g takes the
        // tiles t = g, g + 2, ... (= accumulator stage g), so each group has two tile times per tile.
        constexpr int EG = CODEC == Q_SQ8I ? 2 : 1;
        const int eg = EG == 2 ? (warp - 2) >> 3 : 0;
        const int wl = (warp - 2) & 7;          // warp inside its group
        const int quad = warp & 3;              // TMEM lane quarter this warp may read
        const int colhalf = wl >> 2;
        const int et = wl * 32 + lane;
        const int64_t q = (int64_t)q0 + quad * 32 + lane;
        float *xs = reinterpret_cast<float *>(smem + OFF_XN);
        const float BIG = 3.0e38f;
        const float fq = q < A.nq ? __ldg(A.fq + q) : 0.0f;
        const int nasum = ((CODEC == Q_RABITQI || CODEC == Q_BQI) && q < A.nq) ? -__ldg(A.asum + q) : 0;
        const uint32_t keep_hi = A.keep_hi;
        float g1 = BIG, g2 = BIG;
        int cc = 0;
        // threshold mode: this thread's list and threshold
        const bool qlive = q < A.nq;
        const float T = (THRESH && qlive) ? fminf(__ldg(A.Ts + q), 2.9e38f) : -BIG;   // masked / padding rows carry BIG: never listed
        const int myslot = (split * EG + eg) * 2 + colhalf;
        uint2 *mylist = THRESH ? A.cand + ((size_t)(qlive ? q : 0) * A.slots + myslot) * A.cap : nullptr;
        int ncand = 0;
        // row norms of a tile are fetched one tile ahead (global-load latency out of the per-tile critical path): the value
        // for tile t + 1 is loaded at the top of tile t and stored after tile t's columns are reduced
        auto xn_of = [&](int t_) {
            const int64_t n0 = tile_row0(t_);
            const int64_t row = n0 + et;
            if constexpr (BASE == Q_BQ) return (row < row_end) ? A.half_dim : BIG;  // s = D / 2 - acc / 2 = Hamming, exact
            else return (row < row_end) ? __ldg(A.xn + row) : (BASE == Q_RABITQ ? 1.0e19f : BIG);
        };
        // staging buffer of tile t: two per epilogue group (a group's next tile is written while slower warps of the group may
        // still read the current one)
        auto xbuf = [&](int t_) { return EG == 2 ? eg * 2 + ((t_ >> 1) & 1) : (t_ & 1); };
        if (eg < ntiles) xs[xbuf(eg) * TILE_ROWS + et] = xn_of(eg);
        for (int t = eg; t < ntiles; t += EG) {
            const int as = t & 1;
            const uint32_t aph = (t >> 1) & 1;
            const int64_t n0 = tile_row0(t);
            float *xt = xs + xbuf(t) * TILE_ROWS;
            if (EG == 2 && eg == 1) asm volatile("bar.sync 2, 256;" ::: "memory");
            else asm volatile("bar.sync 1, 256;" ::: "memory");
            const float xn_next = (t + EG < ntiles) ? xn_of(t + EG) : 0.0f;
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * TILE_ROWS + colhalf * BN);
            const int64_t nh = n0 + colhalf * BN;  // first row of this thread's column half
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)(c * 32), v);
                uint32_t mw = 0xFFFFFFFFu;
                if (A.mask) mw = (nh + c * 32 < A.rows) ? __ldg(A.mask + ((nh + c * 32) >> 5)) : 0u;
                tmem_ld_wait();
                float s[32];
                const float4 *x4 = reinterpret_cast<const float4 *>(xt + colhalf * BN + c * 32);
#pragma unroll
                for (int j4 = 0; j4 < 8; j4++) {
                    const float4 xv = x4[j4];
                    const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        // kind::i8 accumulates in int32 (|acc| <= 768 * 127 * 255 < 2^25: the conversion is exact or within 2^-24)
                        float accf;
                        if constexpr (CODEC == Q_RABITQI || CODEC == Q_BQI) {   // sum a s_x = 2 sum a bit - sum a
                            int t;
                            asm("mad.lo.s32 %0, %1, 2, %2;" : "=r"(t) : "r"((int)v[j4 * 4 + i]), "r"(nasum));
                            accf = __int2float_rn(t);
                        } else accf = i8_codec(CODEC) ? __int2float_rn((int)v[j4 * 4 + i]) : __uint_as_float(v[j4 * 4 + i]);
                        const float t = __fmaf_rn(fq, accf, xx[i]);
                        s[j4 * 4 + i] = BASE == Q_RABITQ ? __fmul_rn(t, xx[i]) : t;  // RaBitQ: yn^2 - (2 qn / D) yn acc
                    }
                }
                if (mw != 0xFFFFFFFFu) {
#pragma unroll
                    for (int j = 0; j < 32; j++) s[j] = (mw >> j) & 1u ? s[j] : BIG;
                }
                if constexpr (THRESH) {
                    // one FFMA per element above, a 3-input minimum tree here, one compare per chunk; a chunk that beats the
                    // threshold is opened one 8-row group at a time
                    float gm[4];
#pragma unroll
                    for (int g = 0; g < 4; g++)
                        gm[g] = fminf(fminf(fminf(s[g * 8], s[g * 8 + 1]), s[g * 8 + 2]),
                                      fminf(fminf(fminf(s[g * 8 + 3], s[g * 8 + 4]), s[g * 8 + 5]), fminf(s[g * 8 + 6], s[g * 8 + 7])));
                    if (fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3])) < T) {
                        const uint32_t r0 = (uint32_t)(nh + c * 32);
#pragma unroll
                        for (int g = 0; g < 4; g++) {
                            if (gm[g] < T) {
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    if (s[g * 8 + j] < T) {
                                        if (ncand < A.cap) mylist[ncand] = make_uint2(__float_as_uint(s[g * 8 + j]), r0 + (uint32_t)(g * 8 + j));
                                        ncand++;
                                    }
                                }
                            }
                        }
                    }
                } else {
                float a1[4] = {BIG, BIG, BIG, BIG}, a2[4] = {BIG, BIG, BIG, BIG};
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    uint32_t vb;
                    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(vb) : "r"(__float_as_uint(s[j])), "r"(keep_hi), "r"((uint32_t)j));
                    const float v1 = __uint_as_float(vb);
                    a2[j & 3] = fminf(a2[j & 3], fmaxf(a1[j & 3], v1));
                    a1[j & 3] = fminf(a1[j & 3], v1);
                }
                const float p1 = fminf(a1[0], a1[1]), p2 = fminf(fmaxf(a1[0], a1[1]), fminf(a2[0], a2[1]));
                const float r1 = fminf(a1[2], a1[3]), r2 = fminf(fmaxf(a1[2], a1[3]), fminf(a2[2], a2[3]));
                const uint32_t cidx = (uint32_t)(nh + c * 32) & A.idx_mask & ~31u;
                const float c1 = __uint_as_float((__float_as_uint(fminf(p1, r1)) & ~A.idx_mask) | cidx | (__float_as_uint(fminf(p1, r1)) & 31u));
                const float c2 = fminf(fmaxf(p1, r1), fminf(p2, r2));
                g2 = fminf(fmaxf(g1, c1), fminf(g2, c2));
                g1 = fminf(g1, c1);
                if (++cc == A.cpg) {  // groups are <= 128 rows and aligned: they never leave this thread's column half
                    const int64_t gid = (nh + c * 32) / (32 * (int64_t)A.cpg);
                    if (gid < A.groups) A.mins[q * A.groups + gid] = make_float2(g1, g2);
                    g1 = BIG;
                    g2 = BIG;
                    cc = 0;
                }
                }
            }
            if (t + EG < ntiles) xs[xbuf(t + EG) * TILE_ROWS + et] = xn_next;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_bar(as), 0);
        }
        if constexpr (THRESH) {
            if (qlive) {
                A.ccnt[(size_t)q * A.slots + myslot] = ncand < A.cap ? ncand : A.cap;
                if (ncand > A.cap) A.ovf[q] = 1;
            }
        }
    } else {
        // ===================== decode producers: this CTA's 128 rows of the 256-row tile =====================
        const int grp = (warp - PROD_WARP0) >> 2;
        struct Cursor {
            int t, kb;
            __device__ __forceinline__ void init(int it, int KB) {
                t = it / KB;
                kb = it - t * KB;
            }
            __device__ __forceinline__ bool advance2(int KB) {
                kb += 2;
                bool moved = false;
                while (kb >= KB) {
                    kb -= KB;
                    t++;
                    moved = true;
                }
                return moved;
            }
        };
        const int64_t half_off = (int64_t)rank * BN;   // this CTA's half of every tile
        if constexpr (CODEC == Q_SQ8I) {
            (void)grp;   // kind::i8: the code bytes are the operand, TMA delivers them — these warps have nothing to decode
            (void)half_off;
        } else if constexpr (IS_PQ) {
            using PQProducer = std::conditional_t<CODEC == Q_PQI, ProducerPQI8, ProducerPQF16>;
            using PQCodes = typename PQProducer::Codes;
            const int r = ((warp - PROD_WARP0) & 3) * 32 + lane;
            const int swz = r & 7;
            auto row_of = [&](int t) {
                const int64_t row = tile_row0(t) + half_off + r;
                return row < A.rows ? row : A.rows - 1;
            };
            // codes of iterations it, it + 2, it + 4 in flight (8 bytes each); the centroid bytes come from the slice ring
            PQProducer cur;
            Cursor c0, c2, c4;
            c0.init(grp, A.kb);
            c2 = c0;
            c2.advance2(A.kb);
            c4 = c2;
            c4.advance2(A.kb);
            PQCodes k0 = PQProducer::zero(), k2 = k0, k4 = k0;
            if (grp < total_it) k0 = PQProducer::load_codes(A, row_of(c0.t), c0.kb);
            if (grp + 2 < total_it) k2 = PQProducer::load_codes(A, row_of(c2.t), c2.kb);
            const uint32_t slice = s_base + OFF_SLICE + (uint32_t)grp * SLICE_BYTES;
            for (int it = grp; it < total_it; it += 2) {
                if (it + 4 < total_it) k4 = PQProducer::load_codes(A, row_of(c4.t), c4.kb);
                mbar_wait(sfull_bar(grp), (uint32_t)(it >> 1) & 1);
                cur.gather_smem(A, k0, c0.kb, slice);
                const int st = it % NST;
                const uint32_t ph = (it / NST) & 1;
                mbar_wait(empty_bar(st), ph ^ 1);
                cur.convert(A, c0.kb, s_base + st * STAGE2_BYTES + A2_BYTES + r * 128, swz);
                fence_proxy_async_smem();
                __syncwarp();
                // The slot is released only AFTER convert() has consumed the gathered registers: released right behind the
                // ld.shared instructions, the next slice (an async-proxy bulk copy) could land while gathers were still in
                // flight — 1 batch in ~14 came back with a wrong neighbour (found by the 8-GPU merged-parity check).
                if (lane == 0) {
                    mbar_arrive(sempty_bar(grp));  // the slot may take the slice of it + 2
                    mbar_arrive_cluster(full_bar(st), 0);
                }
                k0 = k2;
                k2 = k4;
                c0 = c2;
                c2 = c4;
                c4.advance2(A.kb);
            }
        } else if constexpr (sign_codec(BASE)) {
            using SignProducer = std::conditional_t<i8_codec(CODEC), ProducerSignI8, Producer<Q_RABITQ>>;
            const int r = ((warp - PROD_WARP0) & 3) * 32 + lane;
            const int swz = r & 7;
            auto row_of = [&](int t) {
                const int64_t row = tile_row0(t) + half_off + r;
                return row < A.rows ? row : A.rows - 1;
            };
            // three buffers in rotation (see the byte producers below): loads of it + 2 and it + 4 in flight
            SignProducer b0, b1, b2;
            Cursor cf;
            cf.init(grp, A.kb);
            int it = grp;
            auto fetch_next = [&](SignProducer &buf, int it_f) {
                if (it_f < total_it) buf.fetch(A, row_of(cf.t), cf.kb);
                cf.advance2(A.kb);
            };
            int kb_cur = cf.kb;  // k-block of iteration `it` (convert() does not need it, kept for symmetry)
            auto step = [&](const SignProducer &cur, SignProducer &far) {
                fetch_next(far, it + 4);
                const int st = it % NST;
                const uint32_t ph = (it / NST) & 1;
                mbar_wait(empty_bar(st), ph ^ 1);
                cur.convert(A, kb_cur, s_base + st * STAGE2_BYTES + A2_BYTES + r * 128, swz);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(full_bar(st), 0);
                it += 2;
            };
            fetch_next(b0, it);
            fetch_next(b1, it + 2);
            while (it < total_it) {
                step(b0, b2);
                if (it >= total_it) break;
                step(b1, b0);
                if (it >= total_it) break;
                step(b2, b1);
            }
        } else {
            // Three register buffers in rotation: while iteration `it` is converted, the loads of it + 2 and it + 4 (one and
            // two steps of this group ahead) are in flight — one step (two k-blocks of MMA time) is shorter than the
            // L2 latency under load, which left the MMA waiting for B (ncu: the producers' top stall was the load result).
            const int slab = ((warp - PROD_WARP0) & 3) * 32;
            ProducerBytes<CODEC> b0, b1, b2;
            typename ProducerBytes<CODEC>::Rows rows;
            Cursor cf;  // cursor of the iteration whose loads are issued next
            cf.init(grp, A.kb);
            rows.set(A, tile_row0(cf.t < ntiles ? cf.t : 0) + half_off + slab, lane);
            int it = grp;
            auto fetch_next = [&](ProducerBytes<CODEC> &buf, int it_f) {
                if (it_f < total_it) buf.fetch(rows, cf.kb);
                if (cf.advance2(A.kb)) rows.set(A, tile_row0(cf.t < ntiles ? cf.t : 0) + half_off + slab, lane);
            };
            auto step = [&](const ProducerBytes<CODEC> &cur, ProducerBytes<CODEC> &far) {
                fetch_next(far, it + 4);
                if (CODEC == Q_INT4I && res_a) {   // B-only stages behind the resident query tile
                    const int st = it % nstb;
                    const uint32_t ph = (it / nstb) & 1;
                    mbar_wait(bempty_bar(st), ph ^ 1);
                    cur.convert(b_base + st * B2_BYTES, slab, lane);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(bfull_bar(st), 0);
                } else {
                    const int st = it % NST;
                    const uint32_t ph = (it / NST) & 1;
                    mbar_wait(empty_bar(st), ph ^ 1);
                    cur.convert(s_base + st * STAGE2_BYTES + A2_BYTES, slab, lane);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(full_bar(st), 0);
                }
                it += 2;
            };
            fetch_next(b0, it);
            fetch_next(b1, it + 2);
            while (it < total_it) {
                step(b0, b2);
                if (it >= total_it) break;
                step(b1, b0);
                if (it >= total_it) break;
                step(b2, b1);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still touch its barriers / shared memory
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ query preparation
// One warp per query.  a_p = fp16(q[perm[p]] * w[p] * 2^e) in STORAGE order (perm[p] = dimension at storage position p,
// -1 = padding; w[p] = decode weight of that dimension), e such that max|a_p| lands in [2^11, 2^12);
// f_q = -2 / 2^e;  c_q = -2 q.mid (float64 accumulation).
__global__ void __launch_bounds__(256) prep_queries_kernel(const float *queries, int64_t nq, int64_t q_stride, int dimp, const int32_t *perm,
                                                           const float *wq, const float *midp, __half *a16, float *fq, float *cq) {
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const float *qv = queries + q * q_stride;
    float mx = 0.0f;
    double cm = 0.0;
    for (int p = lane; p < dimp; p += 32) {
        const int d = perm[p];
        if (d >= 0) {
            mx = fmaxf(mx, fabsf(__fmul_rn(qv[d], wq[p])));
            cm += (double)qv[d] * (double)midp[p];
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        cm += __shfl_xor_sync(0xffffffffu, cm, o);
    }
    int e = 0;
    if (mx > 0.0f && mx < __int_as_float(0x7f800000)) {
        int ex;
        frexpf(mx, &ex);  // mx = m * 2^ex, m in [0.5, 1)
        e = 12 - ex;
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
    }
    const float sq = ldexpf(1.0f, e);
    for (int p = lane; p < dimp; p += 32) {
        const int d = perm[p];
        a16[q * dimp + p] = __float2half_rn(d >= 0 ? __fmul_rn(__fmul_rn(qv[d], wq[p]), sq) : 0.0f);
    }
    if (lane == 0) {
        fq[q] = -ldexpf(1.0f, 1 - e);
        cq[q] = (float)(-2.0 * cm);
    }
}

// kind::i8 form of the query tile (SQ8).  a_p = q[perm[p]] * w[p] as above, quantised per query to signed 8-bit:
// Delta = max|a_p| / 127, ah_p = rint(a_p / Delta), e_p = a_p - Delta ah_p.  With the RAW code bytes c_p = b_p + 128 as the
// B operand the GEMM gives acc = sum ah_p c_p exactly, and
//     sum a_p b_p = Delta acc - 128 Delta sum ah_p + sum e_p b_p
// so the epilogue's s' = ||x^||^2 - 2 Delta acc differs from the fp16 filter's target by the per-query constant
// 256 Delta sum ah_p (2 x offset x Delta sum ah_p in general; folded into c_q) and by 2 sum e_p b_p, which Cauchy-Schwarz bounds by
// 2 ||e / w|| ||x^ - mid|| = ea[q] * max ||x^ - mid||: the term that replaces c1 ||q|| max||x^ - mid|| in the certificate.
__global__ void __launch_bounds__(256) prep_queries_i8_kernel(const float *queries, int64_t nq, int64_t q_stride, int dimp, const int32_t *perm,
                                                              const float *wq, const float *midp, int8_t *a8, float *fq, float *cq, float *ea,
                                                              float code_offset /* c = b + offset: 128 (SQ8 bytes), 8 (INT4 nibbles) */) {
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const float *qv = queries + q * q_stride;
    float mx = 0.0f;
    double cm = 0.0;
    for (int p = lane; p < dimp; p += 32) {
        const int d = perm[p];
        if (d >= 0) {
            mx = fmaxf(mx, fabsf(__fmul_rn(qv[d], wq[p])));
            cm += (double)qv[d] * (double)midp[p];
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        cm += __shfl_xor_sync(0xffffffffu, cm, o);
    }
    const bool ok = mx > 0.0f && mx < __int_as_float(0x7f800000);
    const float delta = ok ? __fdiv_rn(mx, 127.0f) : 0.0f;
    const float inv = ok ? __fdiv_rn(127.0f, mx) : 0.0f;
    double err2 = 0.0;
    int suma = 0;
    bool bad = false;
    for (int p = lane; p < dimp; p += 32) {
        const int d = perm[p];
        int ah = 0;
        if (d >= 0) {
            const float a = __fmul_rn(qv[d], wq[p]);
            if (!(fabsf(a) <= mx)) bad = true;   // NaN
            ah = __float2int_rn(__fmul_rn(a, inv));
            ah = ah > 127 ? 127 : (ah < -127 ? -127 : ah);
            const double e = (double)a - (double)delta * (double)ah;
            if (wq[p] != 0.0f) {
                const double ew = e / (double)wq[p];
                err2 += ew * ew;
            }
            suma += ah;
        }
        a8[q * dimp + p] = (int8_t)ah;
    }
    for (int o = 16; o > 0; o >>= 1) {
        err2 += __shfl_xor_sync(0xffffffffu, err2, o);
        suma += __shfl_xor_sync(0xffffffffu, suma, o);
        bad = __shfl_xor_sync(0xffffffffu, bad ? 1 : 0, o) || bad;
    }
    if (lane == 0) {
        fq[q] = __fmul_rn(-2.0f, delta);
        // s'_true = s'_filter + 256 Delta sum ah - 2 sum e b;  reference ~ s'_true + c_q + ||q||^2 with c_q = -2 q.mid
        cq[q] = (float)(-2.0 * cm + 2.0 * (double)code_offset * (double)delta * (double)suma);
        const bool finite_q = (mx == 0.0f) || ok;
        ea[q] = (bad || !finite_q) ? __int_as_float(0x7f800000) : (float)(2.0 * sqrt(err2) * 1.000001 + 1.0e-30);
    }
}

// RaBitQ: the query side of the estimator (rabitq.go:119-176) is sign(q) and ||q||, both already prepared for the exact
// scan (prep_sign_queries): a_p = +-1 in storage order, f_q = -2 ||q|| / D, c_q = ||q||^2, so that
// dist = (qn - yn)^2 + (4 qn yn / D) h = c_q + yn^2 + f_q yn acc   with acc = sum s_q s_x = D - 2 h  (exact in fp16 x fp16 -> fp32).
__global__ void __launch_bounds__(256) prep_queries_sign_kernel(const uint32_t *q_words, const float *q_norms, int64_t nq, int words32, int dim,
                                                                int dimp, const int32_t *perm, __half *a16, float *fq, float *cq) {
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const uint32_t *qw = q_words + q * words32;
    for (int p = lane; p < dimp; p += 32) {
        const int d = perm[p];
        float v = 0.0f;
        if (d >= 0) v = ((qw[d >> 5] >> (d & 31)) & 1u) ? 1.0f : -1.0f;
        a16[q * dimp + p] = __float2half_rn(v);
    }
    if (lane == 0) {
        if (q_norms) {
            const float qn = q_norms[q];
            fq[q] = __fdiv_rn(__fmul_rn(-2.0f, qn), (float)dim);
            cq[q] = __fmul_rn(qn, qn);
        } else {  // BQ: s = D / 2 - acc / 2
            fq[q] = -0.5f;
            cq[q] = 0.0f;
        }
    }
}
// kind::i8 form: sign(q) as +-1 signed bytes in natural dimension order (dim % 128 == 0: no padding), same f_q / c_q
__global__ void __launch_bounds__(256) prep_queries_sign_i8_kernel(const uint32_t *q_words, const float *q_norms, int64_t nq, int words32, int dim,
                                                                   int8_t *a8, float *fq, float *cq, int32_t *asum) {
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const uint32_t *qw = q_words + q * words32;
    int sa = 0;
    for (int d = lane; d < dim; d += 32) {
        const int v = ((qw[d >> 5] >> (d & 31)) & 1u) ? 1 : -1;
        a8[q * dim + d] = (int8_t)v;
        sa += v;
    }
    for (int o = 16; o > 0; o >>= 1) sa += __shfl_xor_sync(0xffffffffu, sa, o);
    if (lane == 0) {
        asum[q] = sa;
        if (q_norms) {
            const float qn = q_norms[q];
            fq[q] = __fdiv_rn(__fmul_rn(-2.0f, qn), (float)dim);
            cq[q] = __fmul_rn(qn, qn);
        } else {  // BQ: s = D / 2 - acc / 2
            fq[q] = -0.5f;
            cq[q] = 0.0f;
        }
    }
}
__global__ void __launch_bounds__(256) norm_sq_max_kernel(const float *norms, int64_t rows, unsigned int *max_bits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const float y = norms[i];
    atomicMax(max_bits, __float_as_uint(__fmul_rn(y, y)));
}

// ------------------------------------------------------------------ exact decode helpers (reference arithmetic)
struct EArgs {
    const uint8_t *codes;
    int64_t row_bytes;
    int layout;               // SQ8: 0 row-major | VB of the lane-transposed layout; INT4: 0 | 1 permuted; PQ: 0 | 1 tiled
    int int4_lut;             // INT4 gather scoring: 1 = simd.Int4L2DistancePrecomputed over BuildInt4LookupTable values
    const float *p0, *p1;     // SQ8 mins, invScales | INT4 min, diff (natural dimension order)
    const int8_t *codebooks;  // PQ
    const float *pq_scales, *pq_offsets;
    int pq_m, pq_dsub;
    const float *norms;       // RaBitQ: stored row norms
    const uint32_t *q_words;  // RaBitQ: prepared query sign words [nq][words32]
    const float *q_norms;     // RaBitQ: prepared query norms
    int words32;
    int64_t dim, rows;
    const float *queries;
    int64_t q_stride;
    const uint32_t *cand;
    const int32_t *gcnt;
    int kc, G;
    const float *tau, *qn, *cq;
    const float *ea;                // kind::i8 filter: per-query operand-quantisation term (replaces c1 ||q||), else nullptr
    const unsigned int *xmax_bits;  // [0] max ||x^||^2, [1] max ||x^ - mid||^2 (float bits)
    float mid_norm;                 // ||mid||
    const uint8_t *mask;
    int k, C;
    uint32_t row_base;
    uint32_t *out_rows;
    float *out_scores;
    int32_t *out_counts, *fail_flags;
};
__device__ __forceinline__ int64_t sq8_off(int64_t d, int vb) {
    if (vb == 0) return d;
    const int blk = 16 * vb;
    const int64_t b = d / blk;
    const int o = (int)(d - b * blk);
    return b * blk + (o & 15) * vb + (o >> 4);
}
__device__ __forceinline__ int64_t int4_off(int64_t d, int perm) {
    const int64_t o = d >> 1;
    if (!perm) return o;
    const int w = (int)(o & 127);
    return (o >> 7) * 128 + 16 * (w & 7) + 4 * (w >> 5) + ((w >> 3) & 3);
}
__device__ __forceinline__ int64_t pq_off(int64_t row, int m, int64_t row_bytes, int tiled) {
    return tiled ? (row >> 5) * (32 * row_bytes) + (int64_t)(m >> 4) * 512 + (row & 31) * 16 + (m & 15) : row * row_bytes + m;
}
__device__ __forceinline__ float sq8_value(const EArgs &E, const uint8_t *code, int64_t d) {
    return __fmaf_rn(u8_to_f32(__ldg(code + sq8_off(d, E.layout))), __ldg(E.p1 + d), __ldg(E.p0 + d));
}
__device__ __forceinline__ float int4_value(const EArgs &E, const uint8_t *code, int64_t d) {
    const uint32_t b = __ldg(code + int4_off(d, E.layout));
    const float nib = u8_to_f32((d & 1) ? (b & 0x0Fu) : (b >> 4));
    return __fmaf_rn(__fmul_rn(nib, __uint_as_float(0x3d888889u)), __ldg(E.p1 + d), __ldg(E.p0 + d));
}

// Half-warp score of one row in the reference's order; valid in lane 0.
template <int CODEC>
__device__ __forceinline__ float exact_score(const EArgs &E, const float *qs, const float *table, int64_t row, int lane) {
    const int64_t dim = E.dim;
    if constexpr (CODEC == Q_SQ8) {
        // sq8_avx512.c:59-104: one 16-lane accumulator, rec = fma(c, inv, min); diff = q - rec; acc = fma(diff, diff, acc).
        // `table` holds mins | invScales staged in shared memory.  In the lane-transposed layouts lane l's codes of VB
        // consecutive steps are contiguous: one 16-byte (4-byte) load feeds 16 (4) steps.
        const uint8_t *code = E.codes + row * E.row_bytes;
        const float *mn = table, *iv = table + dim;
        float acc = 0.0f;
        int64_t j = 0;
        if (E.layout == 16) {
            for (; j + 256 <= dim; j += 256) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(code + j + lane * 16));
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int s_ = 0; s_ < 16; s_++) {
                    const int64_t d = j + 16 * s_ + lane;
                    const float c = u8_to_f32((w[s_ >> 2] >> (8 * (s_ & 3))) & 0xFFu);
                    const float df = __fsub_rn(qs[d], __fmaf_rn(c, iv[d], mn[d]));
                    acc = __fmaf_rn(df, df, acc);
                }
            }
        } else if (E.layout == 4) {
            for (; j + 64 <= dim; j += 64) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(code + j + lane * 4));
#pragma unroll
                for (int s_ = 0; s_ < 4; s_++) {
                    const int64_t d = j + 16 * s_ + lane;
                    const float c = u8_to_f32((w >> (8 * s_)) & 0xFFu);
                    const float df = __fsub_rn(qs[d], __fmaf_rn(c, iv[d], mn[d]));
                    acc = __fmaf_rn(df, df, acc);
                }
            }
        }
        for (; j + 16 <= dim; j += 16) {
            const int64_t d = j + lane;
            const float df = __fsub_rn(qs[d], __fmaf_rn(u8_to_f32(__ldg(code + sq8_off(d, E.layout))), iv[d], mn[d]));
            acc = __fmaf_rn(df, df, acc);
        }
        float tot = reduce16(acc);
        if (lane == 0)
            for (int64_t d = j; d < dim; d++) {
                const float df = __fsub_rn(qs[d], __fmaf_rn(u8_to_f32(__ldg(code + sq8_off(d, E.layout))), iv[d], mn[d]));
                tot = __fmaf_rn(df, df, tot);
            }
        return tot;
    } else if constexpr (CODEC == Q_INT4) {
        // int4_avx512.c:193-299: S1 takes the first two 16-dim blocks of every 64, S2 the last two; 32-blocks into S1.
        // `table` holds min | diff staged in shared memory.  Permuted layout: lanes 2p, 2p+1 share the 16 bytes at 16 p of
        // every 128-byte block, byte 4 e + b of them holds dims 64 e + 16 b + 2 p (+1).
        const uint8_t *code = E.codes + row * E.row_bytes;
        const float *mn = table, *df_ = table + dim;
        if (E.int4_lut) {
            // Int4Quantizer.L2Distance after Train / UnmarshalBinary (int4.go:62,141-144,216): simd.Int4L2DistancePrecomputed
            // (int4_avx512.c:141-189: ONE 16-lane accumulator, e = q - lut[d*16 + nib], acc = fma(e, e, acc), lane tree,
            // FMA tail) over the table BuildInt4LookupTable fills with unfused Go arithmetic (kernels.go:94-103:
            // (float32(nib) / 15) * diff + min).  The table entry is recomputed here — same three roundings, same bits.
            auto lutv = [&](int64_t d, float nib) { return __fadd_rn(__fmul_rn(__fdiv_rn(nib, 15.0f), df_[d]), mn[d]); };
            auto nib_of = [&](int64_t d) {
                const uint32_t b = __ldg(code + int4_off(d, E.layout));
                return u8_to_f32((d & 1) ? (b & 0x0Fu) : (b >> 4));
            };
            float acc = 0.0f;
            int64_t i = 0;
            for (; i + 16 <= dim; i += 16) {
                const int64_t d = i + lane;
                const float e = __fsub_rn(qs[d], lutv(d, nib_of(d)));
                acc = __fmaf_rn(e, e, acc);
            }
            float tot = reduce16(acc);
            if (lane == 0)
                for (int64_t d = i; d < dim; d++) {
                    const float e = __fsub_rn(qs[d], lutv(d, nib_of(d)));
                    tot = __fmaf_rn(e, e, tot);
                }
            return tot;
        }
        const float k15 = __uint_as_float(0x3d888889u);
        auto val = [&](int64_t d, float nib) { return __fmaf_rn(__fmul_rn(nib, k15), df_[d], mn[d]); };
        float s1 = 0.0f, s2 = 0.0f;
        int64_t i = 0;
        if (E.layout) {
            const int sh = (lane & 1) ? 0 : 4;  // even lane (even dim) = high nibble
            for (; i + 256 <= dim; i += 256) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(code + (i >> 1) + (lane >> 1) * 16));
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; e++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int64_t d = i + 64 * e + 16 * b + lane;
                        const float nib = u8_to_f32((w[e] >> (8 * b + sh)) & 0xFu);
                        const float e_ = __fsub_rn(qs[d], val(d, nib));
                        if (b < 2) s1 = __fmaf_rn(e_, e_, s1);
                        else s2 = __fmaf_rn(e_, e_, s2);
                    }
            }
        }
        auto nibble = [&](int64_t d) {
            const uint32_t b = __ldg(code + int4_off(d, E.layout));
            return u8_to_f32((d & 1) ? (b & 0x0Fu) : (b >> 4));
        };
        for (; i + 64 <= dim; i += 64) {
#pragma unroll
            for (int blk = 0; blk < 4; blk++) {
                const int64_t d = i + blk * 16 + lane;
                const float e = __fsub_rn(qs[d], val(d, nibble(d)));
                if (blk < 2) s1 = __fmaf_rn(e, e, s1);
                else s2 = __fmaf_rn(e, e, s2);
            }
        }
        for (; i + 32 <= dim; i += 32) {
#pragma unroll
            for (int blk = 0; blk < 2; blk++) {
                const int64_t d = i + blk * 16 + lane;
                const float e = __fsub_rn(qs[d], val(d, nibble(d)));
                s1 = __fmaf_rn(e, e, s1);
            }
        }
        float tot = reduce16(__fadd_rn(s1, s2));
        if (lane == 0)
            for (int64_t d = i; d < dim; d++) {
                const float e = __fsub_rn(qs[d], val(d, nibble(d)));
                tot = __fmaf_rn(e, e, tot);
            }
        return tot;
    } else if constexpr (sign_codec(CODEC)) {
        // rabitq.go:119-176: exact popcount (popcount_avx512.c:25-46), then the unfused Go estimator
        const uint32_t *code = reinterpret_cast<const uint32_t *>(E.codes + row * E.row_bytes);
        const uint32_t *qw = reinterpret_cast<const uint32_t *>(table);  // the query's sign words staged in shared memory
        int h = 0;
        for (int w = lane; w < E.words32; w += 16) h += __popc(__ldg(code + w) ^ qw[w]);
        h += __shfl_down_sync(0xffffffffu, h, 8, 16);
        h += __shfl_down_sync(0xffffffffu, h, 4, 16);
        h += __shfl_down_sync(0xffffffffu, h, 2, 16);
        h += __shfl_down_sync(0xffffffffu, h, 1, 16);
        if constexpr (CODEC == Q_BQ) return (float)h;  // distance.Hamming as float32 (binary.go: score = float32(popcount))
        const float qn = qs[0], yn = __ldg(E.norms + row);
        const float t1 = __fsub_rn(qn, yn);
        float a = __fmul_rn(4.0f, qn);
        a = __fmul_rn(a, yn);
        a = __fdiv_rn(a, (float)E.dim);
        return __fadd_rn(__fmul_rn(t1, t1), __fmul_rn(a, (float)h));
    } else {
        // floats_avx512.c:135-167: lane l sums table[(16t+l)*256 + code[16t+l]] over t, reduce, sequential tail
        const int M = E.pq_m, t16 = M >> 4, tail = M & 15;
        float s = 0.0f;
        for (int t = 0; t < t16; t++) {
            const int m = 16 * t + lane;
            const uint32_t c = __ldg(E.codes + pq_off(row, m, E.row_bytes, E.layout));
            s = __fadd_rn(s, table[m * 256 + c]);
        }
        float tot = reduce16(s);
        if (lane == 0)
            for (int m = t16 * 16; m < t16 * 16 + tail; m++)
                tot = __fadd_rn(tot, table[m * 256 + __ldg(E.codes + pq_off(row, m, E.row_bytes, E.layout))]);
        return tot;
    }
}

// Exact stage: one CTA per query (see tc_exact_kernel in vg_flat_tc.cu; same candidate list format).
template <int CODEC>
__global__ void __launch_bounds__(128) qtc_exact_kernel(EArgs E) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x, hw = tid >> 4, lane = tid & 15;
    float *qs = reinterpret_cast<float *>(smem);
    const size_t qbytes = ((size_t)E.dim * 4 + 15) & ~(size_t)15;
    TopK tk = topk_carve(smem + qbytes, 1, E.C, E.k);
    // candidate lists: rows named by a group's minimum, and crowded groups (all G rows are scored) — the r-th candidate
    // row is found from these two short lists, so the CTA's shared memory does not depend on the row budget
    uint32_t *singles = reinterpret_cast<uint32_t *>(smem + qbytes + topk_smem_bytes(1, E.C));
    uint32_t *crowded = singles + E.kc;
    float *table = reinterpret_cast<float *>(crowded + E.kc);
    if constexpr (sign_codec(CODEC)) {
        if (tid == 0) qs[0] = CODEC == Q_RABITQ ? E.q_norms[q] : 0.0f;
        uint32_t *qw = reinterpret_cast<uint32_t *>(table);
        for (int w = tid; w < E.words32; w += 128) qw[w] = E.q_words[q * E.words32 + w];
    } else {
        for (int64_t d = tid; d < E.dim; d += 128) qs[d] = E.queries[q * E.q_stride + d];
        if constexpr (CODEC == Q_SQ8 || CODEC == Q_INT4)
            for (int64_t d = tid; d < E.dim; d += 128) {  // decode parameters next to the query
                table[d] = E.p0[d];
                table[E.dim + d] = E.p1[d];
            }
    }
    topk_init(tk, 1, tid, 128);
    __syncthreads();
    if constexpr (CODEC == Q_PQ) {
        // simd.BuildDistanceTableInt8, live generic path (kernels.go:354-374): sequential, unfused
        const int ds = E.pq_dsub;
        for (int idx = tid; idx < E.pq_m * 256; idx += 128) {
            const int m = idx >> 8;
            const int8_t *cb = E.codebooks + (int64_t)idx * ds;
            const float scale = E.pq_scales[m], offset = E.pq_offsets[m];
            const float *qv = qs + (int64_t)m * ds;
            float sum = 0.0f;
            for (int i = 0; i < ds; i++) {
                const float v = __fadd_rn(__fmul_rn((float)cb[i], scale), offset);
                const float d = __fsub_rn(qv[i], v);
                sum = __fadd_rn(sum, __fmul_rn(d, d));
            }
            table[idx] = sum;
        }
    }
    const int ng = E.gcnt[q];
    const int trigger = E.C - 16;
    __shared__ int s_total, s_ns;
    if (tid < 32) {
        int ns = 0, nc = 0;
        for (int g0 = 0; g0 < ng; g0 += 32) {
            const int gi = g0 + tid;
            const uint32_t c = gi < ng ? E.cand[q * E.kc + gi] : 0u;
            const bool is_c = gi < ng && (c & 0x80000000u) != 0, is_s = gi < ng && !is_c;
            const unsigned bs = __ballot_sync(0xffffffffu, is_s), bc = __ballot_sync(0xffffffffu, is_c);
            const unsigned lt = (1u << tid) - 1u;
            if (is_s) singles[ns + __popc(bs & lt)] = c;
            if (is_c) crowded[nc + __popc(bc & lt)] = c & 0x7FFFFFFFu;
            ns += __popc(bs);
            nc += __popc(bc);
        }
        if (tid == 0) {
            s_ns = ns;
            s_total = ns + nc * E.G;
        }
    }
    __syncthreads();
    const bool overflow = s_total > LIST_CAP;
    if (!overflow) {
        const int total = s_total, ns = s_ns;
        for (int r0 = 0; r0 < total; r0 += 16) {
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int r = r0 + hw * 2 + u;
                int64_t row = -1;
                if (r < ns) {
                    row = (int64_t)singles[r];
                } else if (r < total) {
                    const int idx = r - ns;
                    row = (int64_t)crowded[idx / E.G] * E.G + (idx % E.G);
                }
                if (row >= E.rows) row = -1;
                if (row >= 0 && E.mask && !((E.mask[row >> 3] >> (row & 7)) & 1)) row = -1;
                const bool valid = row >= 0;
                const float tot = exact_score<CODEC>(E, qs, table, valid ? row : 0, lane);
                if (lane == 0 && valid) topk_offer(tk, 0, make_key(tot, E.row_base + (uint32_t)row, false), trigger);
            }
            __syncthreads();
            topk_block_maintain(tk, 1, tid, 128);
        }
    }
    __syncthreads();
    if (tid < 32) {
        topk_emit_warp(tk, 0, tid, false, E.out_rows + q * E.k, E.out_scores + q * E.k, E.out_counts + q, E.k);
        __syncwarp();
        if (tid == 0) {
            const int m = tk.cnt[0];
            int fail = overflow ? 1 : 0;
            const float t = E.tau[q];
            if (!overflow && t < __int_as_float(0x7f800000)) {
                if (m < E.k) {
                    fail = 1;
                } else if (CODEC == Q_BQ) {
                    // integer scores: tau carries the Hamming distance in its high bits (index bits below; exact for
                    // D < 65536).  Every unscored row has Hamming >= that, so the k best are certain when the k-th
                    // exact distance is strictly smaller (a tie could hide a smaller row id in an unscored group).
                    const float th = __uint_as_float(__float_as_uint(t) & ~(uint32_t)(E.G - 1));
                    if (!(E.out_scores[q * E.k + (E.k - 1)] < th)) fail = 1;
                } else if (CODEC == Q_RABITQ) {
                    // the GEMM is exact (acc = D - 2 Hamming); only the epilogue's float32 arithmetic, the index bits and
                    // the reference estimator's own roundings separate s' + c_q from the reference score
                    const double qq = (double)E.cq[q], xx = (double)__uint_as_float(E.xmax_bits[0]);
                    const double qn_ = sqrt(qq), yn = sqrt(xx);
                    const double smax = xx + 2.0 * qn_ * yn;
                    const double Eb = smax * (1.0 / 2097152.0 + (double)E.G / 8388608.0);
                    const double eref = (qn_ + yn) * (qn_ + yn) / 2097152.0;
                    const double ex = (double)E.out_scores[q * E.k + (E.k - 1)];
                    if (!(ex < (double)t - Eb + qq - eref)) fail = 1;
                } else {
                    const double qq = (double)E.qn[q], xx = (double)__uint_as_float(E.xmax_bits[0]), bb = (double)__uint_as_float(E.xmax_bits[1]);
                    const double qn_ = sqrt(qq), bn = sqrt(bb);
                    const double c1 = 1.125 / 1024.0, c2 = 1.0 / 16384.0 + (double)E.dim / 8388608.0;
                    const double smax = xx + 2.0 * qn_ * bn;  // |s'| of any row
                    const double ca = E.ea ? (double)E.ea[q] : c1 * qn_;   // operand rounding: int8 (measured per query) or fp16 (2^-11 relative)
                    const double Eb = ca * bn + c2 * (qq + fmax(xx, bb)) + smax * (1.0 / 4194304.0 + (double)E.G / 8388608.0) +
                                      qn_ * ((double)E.mid_norm + bn) / 2097152.0;
                    const double eref = ((double)E.dim + 64.0) / 16777216.0;  // the reference's own float32 summation
                    const double ex = (double)E.out_scores[q * E.k + (E.k - 1)];
                    // an unscored row has s' >= tau, i.e. a true distance > tau + c_q - Eb + ||q||^2, and a reference
                    // score >= (1 - eref) of that
                    const double lower = (double)t + (double)E.cq[q] - Eb + qq;
                    if (!(ex < lower * (1.0 - eref) - eref * qq)) fail = 1;
                }
            }
            E.fail_flags[q] = fail;
        }
    }
}

// ------------------------------------------------------------------ threshold pass (second chance)
// Thresholds of the second pass.  U = the k-th best EXACT score the first pass found for the query (scores of real
// rows: an upper bound of the true k-th best).  The pass lists every row with s' < T; T is chosen so that a row that is
// NOT listed (s' >= T) has a reference score strictly above U — the same inequalities as the certificate of
// qtc_exact_kernel, solved for the threshold instead of checked — so the k best of the listed rows ARE the k best of
// the segment, ties included.  U = +inf (fewer than k rows found) lists everything: the lists overflow and the query
// goes to the exact scan.
template <int CODEC>
__global__ void __launch_bounds__(256) qtc_thresh_kernel(const float *kth, const float *qn, const float *cq, const unsigned int *xmax_bits,
                                                         float mid_norm, int dim, int64_t nq, float *Ts, const float *ea) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const double U = (double)kth[q];
    double Td;
    if (CODEC == Q_BQ) {
        Td = U + 0.5;   // s' is the Hamming distance itself: everything at distance <= U is scored (ties decide by row id)
    } else if (CODEC == Q_RABITQ) {
        const double qq = (double)cq[q], xx = (double)__uint_as_float(xmax_bits[0]);
        const double qn_ = sqrt(qq), yn = sqrt(xx);
        const double smax = xx + 2.0 * qn_ * yn;
        const double Eb = smax / 2097152.0;
        const double eref = (qn_ + yn) * (qn_ + yn) / 2097152.0;
        Td = U + Eb - qq + eref;   // unlisted: reference >= T - Eb + qq - eref
    } else {
        const double qq = (double)qn[q], xx = (double)__uint_as_float(xmax_bits[0]), bb = (double)__uint_as_float(xmax_bits[1]);
        const double qn_ = sqrt(qq), bn = sqrt(bb);
        const double c1 = 1.125 / 1024.0, c2 = 1.0 / 16384.0 + (double)dim / 8388608.0;
        const double smax = xx + 2.0 * qn_ * bn;
        const double ca = ea ? (double)ea[q] : c1 * qn_;
        const double Eb = ca * bn + c2 * (qq + fmax(xx, bb)) + smax / 4194304.0 + qn_ * ((double)mid_norm + bn) / 2097152.0;
        const double eref = ((double)dim + 64.0) / 16777216.0;
        // unlisted: reference >= (T + c_q - Eb + ||q||^2)(1 - eref) - eref ||q||^2
        Td = (U + eref * qq) / (1.0 - eref) - (double)cq[q] - qq + Eb;
    }
    Td += fabs(Td) * 1.0e-6 + 1.0e-30;   // strictly above: a tie with the k-th best would decide by row id
    Ts[q] = (Td < 3.0e38) ? __double2float_ru(Td) : __int_as_float(0x7f800000);
}

// kth[i] = the k-th best score of query idx[i] in a [nq][k] result, +inf when the query holds fewer than k rows
__global__ void __launch_bounds__(256) gather_kth_kernel(const float *scores, const int32_t *counts, const int32_t *idx, int64_t n, int k,
                                                         float *kth) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t q = idx ? idx[i] : i;
    kth[i] = counts[q] >= k ? scores[q * k + (k - 1)] : __int_as_float(0x7f800000);
}

// Exact stage of the threshold pass: one CTA per query scores EVERY listed row in the codec's reference order (half-warp
// per row, exact_score) and keeps the best k under (score, row).  No certificate: the lists are complete by construction;
// a query whose list overflowed is flagged.
template <int CODEC>
__global__ void __launch_bounds__(128) qtc_exact_list_kernel(EArgs E, const uint2 *cand, const int *ccnt, const int *ovf, int slots, int cap) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x, hw = tid >> 4, lane = tid & 15;
    float *qs = reinterpret_cast<float *>(smem);
    const size_t qbytes = ((size_t)E.dim * 4 + 15) & ~(size_t)15;
    TopK tk = topk_carve(smem + qbytes, 1, E.C, E.k);
    int *pre = reinterpret_cast<int *>(smem + qbytes + topk_smem_bytes(1, E.C));   // prefix of the list lengths, slots + 1 entries
    float *table = reinterpret_cast<float *>(pre + ((slots + 1 + 3) & ~3));
    if constexpr (sign_codec(CODEC)) {
        if (tid == 0) qs[0] = CODEC == Q_RABITQ ? E.q_norms[q] : 0.0f;
        uint32_t *qw = reinterpret_cast<uint32_t *>(table);
        for (int w = tid; w < E.words32; w += 128) qw[w] = E.q_words[q * E.words32 + w];
    } else {
        for (int64_t d = tid; d < E.dim; d += 128) qs[d] = E.queries[q * E.q_stride + d];
        if constexpr (CODEC == Q_SQ8 || CODEC == Q_INT4)
            for (int64_t d = tid; d < E.dim; d += 128) {
                table[d] = E.p0[d];
                table[E.dim + d] = E.p1[d];
            }
    }
    topk_init(tk, 1, tid, 128);
    if (tid == 0) {
        int acc = 0;
        for (int sl = 0; sl < slots; sl++) {
            pre[sl] = acc;
            acc += ccnt[(size_t)q * slots + sl];
        }
        pre[slots] = acc;
    }
    __syncthreads();
    if constexpr (CODEC == Q_PQ) {
        // simd.BuildDistanceTableInt8, live generic path (kernels.go:354-374): sequential, unfused
        const int ds = E.pq_dsub;
        for (int idx = tid; idx < E.pq_m * 256; idx += 128) {
            const int m = idx >> 8;
            const int8_t *cb = E.codebooks + (int64_t)idx * ds;
            const float scale = E.pq_scales[m], offset = E.pq_offsets[m];
            const float *qv = qs + (int64_t)m * ds;
            float sum = 0.0f;
            for (int i = 0; i < ds; i++) {
                const float v = __fadd_rn(__fmul_rn((float)cb[i], scale), offset);
                const float d = __fsub_rn(qv[i], v);
                sum = __fadd_rn(sum, __fmul_rn(d, d));
            }
            table[idx] = sum;
        }
        __syncthreads();
    }
    const int total = pre[slots];
    const int trigger = E.C - 16;
    for (int r0 = 0; r0 < total; r0 += 16) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int r = r0 + hw * 2 + u;
            int64_t row = -1;
            if (r < total) {
                int lo = 0, hi = slots;   // the list that holds candidate r: last slot with pre[slot] <= r
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (pre[mid] <= r) lo = mid;
                    else hi = mid;
                }
                row = (int64_t)cand[((size_t)q * slots + lo) * cap + (r - pre[lo])].y;
            }
            if (row >= E.rows) row = -1;
            const bool valid = row >= 0;
            const float tot = exact_score<CODEC>(E, qs, table, valid ? row : 0, lane);
            if (lane == 0 && valid) topk_offer(tk, 0, make_key(tot, E.row_base + (uint32_t)row, false), trigger);
        }
        __syncthreads();
        topk_block_maintain(tk, 1, tid, 128);
    }
    __syncthreads();
    if (tid < 32) {
        topk_emit_warp(tk, 0, tid, false, E.out_rows + q * E.k, E.out_scores + q * E.k, E.out_counts + q, E.k);
        if (tid == 0) E.fail_flags[q] = ovf[q] ? 1 : 0;
    }
}

// ------------------------------------------------------------------ gather scoring
// Quantized distance of every query to ITS r candidate rows, in the reference's arithmetic (the neighbour-list scoring of
// the DiskANN traversal: internal/segment/diskann/segment.go:511-588 — pq.AdcDistance / int4.L2Distance /
// rabitq.Distance per neighbour).  One CTA per query stages the query state once (query vector, decode parameters,
// the PQ distance table built by the generic Go loop, RaBitQ sign words); a half-warp scores one row with exact_score.
template <int CODEC>
__global__ void __launch_bounds__(128) qtc_score_kernel(EArgs E, const uint32_t *rows, int r, float *out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x, hw = tid >> 4, lane = tid & 15;
    float *qs = reinterpret_cast<float *>(smem);
    const size_t qbytes = ((size_t)E.dim * 4 + 15) & ~(size_t)15;
    float *table = reinterpret_cast<float *>(smem + qbytes);
    if constexpr (sign_codec(CODEC)) {
        if (tid == 0) qs[0] = CODEC == Q_RABITQ ? E.q_norms[q] : 0.0f;
        uint32_t *qw = reinterpret_cast<uint32_t *>(table);
        for (int w = tid; w < E.words32; w += 128) qw[w] = E.q_words[q * E.words32 + w];
    } else {
        for (int64_t d = tid; d < E.dim; d += 128) qs[d] = E.queries[q * E.q_stride + d];
        if constexpr (CODEC == Q_SQ8 || CODEC == Q_INT4)
            for (int64_t d = tid; d < E.dim; d += 128) {
                table[d] = E.p0[d];
                table[E.dim + d] = E.p1[d];
            }
    }
    __syncthreads();
    if constexpr (CODEC == Q_PQ) {
        const int ds = E.pq_dsub;
        for (int idx = tid; idx < E.pq_m * 256; idx += 128) {
            const int m = idx >> 8;
            const int8_t *cb = E.codebooks + (int64_t)idx * ds;
            const float scale = E.pq_scales[m], offset = E.pq_offsets[m];
            const float *qv = qs + (int64_t)m * ds;
            float sum = 0.0f;
            for (int i = 0; i < ds; i++) {
                const float v = __fadd_rn(__fmul_rn((float)cb[i], scale), offset);
                const float d = __fsub_rn(qv[i], v);
                sum = __fadd_rn(sum, __fmul_rn(d, d));
            }
            table[idx] = sum;
        }
        __syncthreads();
    }
    for (int j0 = 0; j0 < r; j0 += 8) {
        const int j = j0 + hw;
        const uint32_t row = j < r ? rows[q * r + j] : 0xFFFFFFFFu;
        const bool valid = (int64_t)row < E.rows;
        const float tot = exact_score<CODEC>(E, qs, table, valid ? (int64_t)row : 0, lane);
        if (lane == 0 && j < r) out[q * r + j] = valid ? tot : __uint_as_float(0x7fc00000u);
    }
}

// ------------------------------------------------------------------ ||decode(row)||^2 and ||decode(row) - mid||^2
// One thread per row, eight interleaved float32 accumulators (the values only feed the filter and its bound).
template <int CODEC>
__global__ void __launch_bounds__(128) code_norms_kernel(EArgs E, float *xn, unsigned int *max_bits) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= E.rows) return;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, bcc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if constexpr (CODEC == Q_PQ) {
        const int ds = E.pq_dsub;
        for (int m = 0; m < E.pq_m; m++) {
            const uint32_t c = __ldg(E.codes + pq_off(row, m, E.row_bytes, E.layout));
            const int8_t *cb = E.codebooks + ((int64_t)m * 256 + c) * ds;
            const float scale = __ldg(E.pq_scales + m), offset = __ldg(E.pq_offsets + m);
            for (int i = 0; i < ds; i++) {
                const float bw = __fmul_rn((float)cb[i], scale);
                const float v = __fadd_rn(bw, offset);
                acc[i & 7] = __fmaf_rn(v, v, acc[i & 7]);
                bcc[i & 7] = __fmaf_rn(bw, bw, bcc[i & 7]);
            }
        }
    } else {
        const uint8_t *code = E.codes + row * E.row_bytes;
        for (int64_t d = 0; d < E.dim; d++) {
            float v, bw;
            if constexpr (CODEC == Q_SQ8) {
                v = sq8_value(E, code, d);
                bw = __fmul_rn(u8_to_f32(__ldg(code + sq8_off(d, E.layout))) - 128.0f, __ldg(E.p1 + d));
            } else {
                v = int4_value(E, code, d);
                const uint32_t b = __ldg(code + int4_off(d, E.layout));
                const float nib = u8_to_f32((d & 1) ? (b & 0x0Fu) : (b >> 4));
                bw = __fmul_rn(nib - 8.0f, __fmul_rn(__ldg(E.p1 + d), __uint_as_float(0x3d888889u)));
            }
            acc[d & 7] = __fmaf_rn(v, v, acc[d & 7]);
            bcc[d & 7] = __fmaf_rn(bw, bw, bcc[d & 7]);
        }
    }
    const float a = __fadd_rn(__fadd_rn(__fadd_rn(acc[0], acc[1]), __fadd_rn(acc[2], acc[3])),
                              __fadd_rn(__fadd_rn(acc[4], acc[5]), __fadd_rn(acc[6], acc[7])));
    const float b = __fadd_rn(__fadd_rn(__fadd_rn(bcc[0], bcc[1]), __fadd_rn(bcc[2], bcc[3])),
                              __fadd_rn(__fadd_rn(bcc[4], bcc[5]), __fadd_rn(bcc[6], bcc[7])));
    xn[row] = a;
    atomicMax(max_bits, __float_as_uint(a));
    atomicMax(max_bits + 1, __float_as_uint(b));
}

// ------------------------------------------------------------------ host
static int q_codec(const CodecParams &cp) {
    switch (cp.codec) {
        case VG_CODEC_SQ8: return Q_SQ8;
        case VG_CODEC_INT4: return Q_INT4;
        case VG_CODEC_PQ:
        case VG_CODEC_OPQ: return Q_PQ;
        case VG_CODEC_RABITQ: return Q_RABITQ;
        case VG_CODEC_BQ: return Q_BQ;
        default: return -1;
    }
}
static int layout_of(const CodecParams &cp) {
    const bool perm = (cp.variant & VG_VAR_PERM) != 0;
    switch (cp.codec) {
        case VG_CODEC_SQ8: return perm ? (cp.dim % 256 == 0 ? 16 : 4) : 0;
        default: return perm ? 1 : 0;
    }
}

static std::atomic<int> g_enabled{-1};
static std::atomic<uint64_t> g_queries{0}, g_fallbacks{0};
static bool enabled() {
    int v = g_enabled.load();
    if (v < 0) {
        const char *e = getenv("VECGO_QUANT_TC");
        v = (e && e[0] == '0') ? 0 : 1;
        g_enabled.store(v);
    }
    return v != 0 && tc::enabled();
}
void stats(uint64_t *queries, uint64_t *fallbacks) {
    if (queries) *queries = g_queries.load();
    if (fallbacks) *fallbacks = g_fallbacks.load();
}
// Optional CUDA-event timing of the GEMM kernel on its own stream (bench.py's roofline line).
static std::atomic<int> g_prof{0};
static std::mutex g_prof_mu;
static double g_gemm_ms = 0.0;
static uint64_t g_gemm_launches = 0;
// event pairs recorded around GEMM launches and not read yet: the search call does NOT wait for them (a synchronisation per
// launch cost the 8-GPU step 4 %); profile() settles them when the counters are read
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pending;
static void settle_profile_events() {   // g_prof_mu held
    for (auto &pr : g_prof_pending) {
        float ms = 0.0f;
        if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
            g_gemm_ms += ms;
            g_gemm_launches++;
        }
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    g_prof_pending.clear();
    cudaGetLastError();
}
void profile(int enable, double *gemm_ms, uint64_t *gemm_launches) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    settle_profile_events();
    if (gemm_ms) *gemm_ms = g_gemm_ms;
    if (gemm_launches) *gemm_launches = g_gemm_launches;
    if (enable >= 0) {
        g_gemm_ms = 0.0;
        g_gemm_launches = 0;
        g_prof.store(enable ? 1 : 0);
    }
}

// CTA-pair kernel (cta_group::2) unless VECGO_QTC_PAIR=0
static bool use_pair() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("VECGO_QTC_PAIR");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
bool supported(const CodecParams &cp, int metric, int64_t rows, int64_t nq, int64_t k, int64_t num_partitions) {
    if (!enabled() || num_partitions > 1) return false;
    if (rows < 8192 || rows >= (1ll << 31) || nq < 16 || k < 1) return false;
    if (k > 1024) return false;                                             // the engine's refine depth (k * RefineFactor) fits
    if (rows / 32 < 2 * (int64_t)(k <= 16 ? 32 : 2 * k)) return false;      // candidate groups must exist (2k groups of >= 32 rows, twice over)
    if (cp.dim % 64 != 0 || cp.dim < 64 || cp.dim > 2048) return false;
    if ((reinterpret_cast<uintptr_t>(cp.codes) & 15) != 0) return false;
    switch (cp.codec) {
        case VG_CODEC_SQ8:
            return metric == VG_METRIC_L2 && !(cp.variant & VG_VAR_GO_SCALAR);
        case VG_CODEC_INT4:
            return true;  // Int4 scores are L2 distances whatever the segment metric
        case VG_CODEC_RABITQ:
            // the estimator is a distance whatever the segment metric; CTA-pair kernel only; candidate groups must exist
            return use_pair() && cp.norms != nullptr && cp.q_words != nullptr && cp.q_norms != nullptr && rows / 32 >= 2 * (k <= 16 ? 32 : 2 * k);
        case VG_CODEC_BQ:
            // Hamming distance whatever the segment metric; same kernel as RaBitQ without the row norms
            return use_pair() && cp.q_words != nullptr && rows / 32 >= 2 * (k <= 16 ? 32 : 2 * k);
        case VG_CODEC_PQ:
        case VG_CODEC_OPQ: {
            if (metric != VG_METRIC_L2 || cp.pq_k != 256 || cp.pq_tables) return false;
            const int ds = cp.pq_dsub;
            if (ds != 8 && ds != 16 && ds != 32 && ds != 64) return false;
            return (cp.variant & VG_VAR_PERM) || cp.pq_m % 8 == 0;
        }
        default:
            return false;
    }
}

static EArgs eargs_of(const CodecParams &cp, int64_t rows) {
    EArgs e{};
    e.codes = cp.codes;
    e.row_bytes = cp.row_bytes;
    e.layout = layout_of(cp);
    e.p0 = cp.p0;
    e.p1 = cp.p1;
    e.codebooks = cp.pq_codebooks;
    e.pq_scales = cp.pq_scales;
    e.pq_offsets = cp.pq_offsets;
    e.pq_m = cp.pq_m;
    e.pq_dsub = cp.pq_dsub;
    e.norms = cp.norms;
    e.q_words = cp.q_words;
    e.q_norms = cp.q_norms;
    e.words32 = cp.words32;
    e.dim = cp.dim;
    e.rows = rows;
    return e;
}

vg_status prepare(const CodecParams &cp, int64_t rows, const float *h_p0, const float *h_p1, Prepared &pp, cudaStream_t st) {
    const int qc = q_codec(cp);
    if (qc < 0) return fail(VG_ERR_UNSUPPORTED, "codec has no tensor-core filter");
    const int dim = (int)cp.dim, dimp = (dim + 63) / 64 * 64;
    const int layout = layout_of(cp);
    // storage position (= position along K of the B tile the producer writes) -> dimension
    std::vector<int32_t> perm((size_t)dimp, -1);
    if (sign_codec(qc)) {
        // Producer<Q_RABITQ>::convert: position 64 kb + 32 word + 8 c + 2 j + hi holds bit 64 kb + 32 word + 4 c + j + 16 hi
        for (int p = 0; p < dimp; p++) {
            const int base = p & ~31, in = p & 31, c = in >> 3, j = (in >> 1) & 3, hi = in & 1;
            const int d = base + 4 * c + j + 16 * hi;
            perm[(size_t)p] = d < dim ? d : -1;
        }
        VG_TRY(pp.perm.alloc_persistent((size_t)dimp * 4));
        VG_TRY(pp.xmax.alloc_persistent(16));
        VG_CUDA(cudaMemcpyAsync(pp.perm.p, perm.data(), (size_t)dimp * 4, cudaMemcpyHostToDevice, st));
        VG_CUDA(cudaMemsetAsync(pp.xmax.p, 0, 16, st));
        if (rows > 0 && qc == Q_RABITQ) {
            norm_sq_max_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(cp.norms, rows, pp.xmax.as<unsigned int>());
            VG_LAUNCHED();
        }
        VG_CUDA(cudaStreamSynchronize(st));
        pp.dimp = dimp;
        pp.mid_norm = 0.0f;
        pp.ready = true;
        return VG_OK;
    }
    if (qc == Q_SQ8) {
        for (int d = 0; d < dim; d++) {
            int p = d;
            if (layout) {
                const int blk = 16 * layout, b = d / blk, o = d % blk;
                p = b * blk + (o & 15) * layout + (o >> 4);
            }
            perm[(size_t)p] = d;
        }
    } else if (qc == Q_INT4) {
        // device byte B holds natural byte o (dims 2o high nibble, 2o+1 low nibble); the producer turns the word of
        // bytes b0..b3 into the chunk  b0.lo b2.lo b0.hi b2.hi b1.lo b3.lo b1.hi b3.hi  (Producer<Q_INT4>::convert)
        static const int slot_byte[8] = {0, 2, 0, 2, 1, 3, 1, 3}, slot_hi[8] = {0, 0, 1, 1, 0, 0, 1, 1};
        std::vector<int> nat((size_t)(dim / 2));
        for (int o = 0; o < dim / 2; o++) {
            int B = o;
            if (layout) {
                const int w = o & 127;
                B = (o >> 7) * 128 + 16 * (w & 7) + 4 * (w >> 5) + ((w >> 3) & 3);
            }
            nat[(size_t)B] = o;
        }
        for (int p = 0; p < dim; p++) {
            const int word = p >> 3, e = p & 7;
            const int o = nat[(size_t)(4 * word + slot_byte[e])];
            perm[(size_t)p] = 2 * o + (slot_hi[e] ? 0 : 1);  // high nibble = even dimension
        }
    } else {
        for (int d = 0; d < dim; d++) perm[(size_t)d] = d;
    }
    // kind::i8 decode of INT4 (ProducerBytes<Q_INT4I>): position 128 kb + 32 piece + 16 (h >> 1) + 8 (h & 1) + 4 hi + i of the
    // K axis holds the (hi ? high : low) nibble of device byte 64 kb + 16 piece + 4 h + i
    std::vector<int32_t> perm8;
    if (qc == Q_INT4 && dim % 128 == 0) {
        std::vector<int> nat((size_t)(dim / 2));
        for (int o = 0; o < dim / 2; o++) {
            int B = o;
            if (layout) {
                const int w = o & 127;
                B = (o >> 7) * 128 + 16 * (w & 7) + 4 * (w >> 5) + ((w >> 3) & 3);
            }
            nat[(size_t)B] = o;
        }
        perm8.assign((size_t)dim, -1);
        for (int p = 0; p < dim; p++) {
            const int kb = p >> 7, r = p & 127, pc = r >> 5, h = 2 * ((r >> 4) & 1) + ((r >> 3) & 1), hi = (r >> 2) & 1, i = r & 3;
            const int o = nat[(size_t)(64 * kb + 16 * pc + 4 * h + i)];
            perm8[(size_t)p] = 2 * o + (hi ? 0 : 1);  // high nibble = even dimension
        }
    }
    // x^_d = mid_d + w_d * b_d with the integer b_d the producer emits
    std::vector<float> w((size_t)dim), mid((size_t)dim);
    const float k15 = 1.0f / 15.0f;  // 0x3d888889, the constant of int4_avx512.c
    for (int d = 0; d < dim; d++) {
        if (qc == Q_SQ8) {
            w[(size_t)d] = h_p1[d];
            mid[(size_t)d] = (float)((double)h_p0[d] + 128.0 * (double)h_p1[d]);
        } else if (qc == Q_INT4) {
            w[(size_t)d] = h_p1[d] * k15;
            mid[(size_t)d] = (float)((double)h_p0[d] + 8.0 * (double)w[(size_t)d]);
        } else {
            const int m = d / cp.pq_dsub;
            w[(size_t)d] = h_p0[m];
            mid[(size_t)d] = h_p1[m];
        }
    }
    std::vector<float> wq((size_t)dimp, 0.0f), midp((size_t)dimp, 0.0f);
    double mm = 0.0;
    for (int p = 0; p < dimp; p++) {
        const int d = perm[(size_t)p];
        if (d < 0) continue;
        wq[(size_t)p] = w[(size_t)d];
        midp[(size_t)p] = mid[(size_t)d];
        mm += (double)mid[(size_t)d] * (double)mid[(size_t)d];
    }
    VG_TRY(pp.perm.alloc_persistent((size_t)dimp * 4));
    VG_TRY(pp.wq.alloc_persistent((size_t)dimp * 4));
    VG_TRY(pp.midp.alloc_persistent((size_t)dimp * 4));
    VG_TRY(pp.xn.alloc_persistent((size_t)std::max<int64_t>(rows, 1) * 4));
    VG_TRY(pp.xmax.alloc_persistent(16));
    VG_CUDA(cudaMemcpyAsync(pp.perm.p, perm.data(), (size_t)dimp * 4, cudaMemcpyHostToDevice, st));
    VG_CUDA(cudaMemcpyAsync(pp.wq.p, wq.data(), (size_t)dimp * 4, cudaMemcpyHostToDevice, st));
    VG_CUDA(cudaMemcpyAsync(pp.midp.p, midp.data(), (size_t)dimp * 4, cudaMemcpyHostToDevice, st));
    VG_CUDA(cudaMemsetAsync(pp.xmax.p, 0, 16, st));
    std::vector<float> wq8, midp8;
    if (!perm8.empty()) {
        wq8.resize((size_t)dim);
        midp8.resize((size_t)dim);
        for (int p = 0; p < dim; p++) {
            wq8[(size_t)p] = w[(size_t)perm8[(size_t)p]];
            midp8[(size_t)p] = mid[(size_t)perm8[(size_t)p]];
        }
        VG_TRY(pp.perm8.alloc_persistent((size_t)dim * 4));
        VG_TRY(pp.wq8.alloc_persistent((size_t)dim * 4));
        VG_TRY(pp.midp8.alloc_persistent((size_t)dim * 4));
        VG_CUDA(cudaMemcpyAsync(pp.perm8.p, perm8.data(), (size_t)dim * 4, cudaMemcpyHostToDevice, st));
        VG_CUDA(cudaMemcpyAsync(pp.wq8.p, wq8.data(), (size_t)dim * 4, cudaMemcpyHostToDevice, st));
        VG_CUDA(cudaMemcpyAsync(pp.midp8.p, midp8.data(), (size_t)dim * 4, cudaMemcpyHostToDevice, st));
    }
    EArgs e = eargs_of(cp, rows);
    const unsigned blocks = (unsigned)((rows + 127) / 128);
    if (rows > 0) {
        if (qc == Q_SQ8) code_norms_kernel<Q_SQ8><<<blocks, 128, 0, st>>>(e, pp.xn.as<float>(), pp.xmax.as<unsigned int>());
        else if (qc == Q_INT4) code_norms_kernel<Q_INT4><<<blocks, 128, 0, st>>>(e, pp.xn.as<float>(), pp.xmax.as<unsigned int>());
        else code_norms_kernel<Q_PQ><<<blocks, 128, 0, st>>>(e, pp.xn.as<float>(), pp.xmax.as<unsigned int>());
        VG_LAUNCHED();
    }
    VG_CUDA(cudaStreamSynchronize(st));  // the host vectors above are read by the async copies
    pp.dimp = dimp;
    pp.mid_norm = (float)std::sqrt(mm) * 1.0001f;
    pp.ready = true;
    return VG_OK;
}

template <int CODEC>
static vg_status launch_score(const EArgs &e, int64_t nq, const uint32_t *d_rows, int r, float *d_out, cudaStream_t st) {
    const size_t sm = (((size_t)e.dim * 4 + 15) & ~(size_t)15) +
                      (CODEC == Q_PQ ? (size_t)e.pq_m * 256 * 4 : sign_codec(CODEC) ? (size_t)e.words32 * 4 : (size_t)e.dim * 8);
    if (sm > 200 * 1024) return fail(VG_ERR_UNSUPPORTED, "dimension too large for the gather-scoring kernel");
    VG_CUDA(cudaFuncSetAttribute(qtc_score_kernel<CODEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    qtc_score_kernel<CODEC><<<(unsigned)nq, 128, sm, st>>>(e, d_rows, r, d_out);
    VG_LAUNCHED();
    return VG_OK;
}
vg_status score_rows(const CodecParams &cp, int64_t rows, const float *d_queries, int64_t q_stride, int64_t nq, const uint32_t *d_rows,
                     int64_t r, float *d_out, int int4_lut, cudaStream_t st) {
    if (nq <= 0 || r <= 0) return VG_OK;
    const int qc = q_codec(cp);
    if (qc < 0) return fail(VG_ERR_UNSUPPORTED, "codec has no gather-scoring kernel");
    if (qc == Q_PQ && cp.pq_k != 256) return fail(VG_ERR_UNSUPPORTED, "PQ ADC requires K=256 (simd.PqAdcLookup hard-wires the table stride)");
    if (qc == Q_SQ8 && (cp.variant & VG_VAR_GO_SCALAR)) return fail(VG_ERR_UNSUPPORTED, "SQ8 gather scoring is the L2 kernel (Sq8uL2BatchPerDimension)");
    EArgs e = eargs_of(cp, rows);
    e.queries = d_queries;
    e.q_stride = q_stride ? q_stride : cp.dim;
    e.int4_lut = int4_lut;
    if (qc == Q_SQ8) return launch_score<Q_SQ8>(e, nq, d_rows, (int)r, d_out, st);
    if (qc == Q_INT4) return launch_score<Q_INT4>(e, nq, d_rows, (int)r, d_out, st);
    if (qc == Q_RABITQ) return launch_score<Q_RABITQ>(e, nq, d_rows, (int)r, d_out, st);
    if (qc == Q_BQ) return launch_score<Q_BQ>(e, nq, d_rows, (int)r, d_out, st);
    return launch_score<Q_PQ>(e, nq, d_rows, (int)r, d_out, st);
}

static int candidates_for(int64_t k) { return k <= 16 ? 32 : (int)(2 * k); }

// SQ8 through kind::i8 (twice the kind::f16 rate, no decode): on unless VECGO_QTC_I8=0 / set_i8(false).  The 8-bit query
// tile makes the certificate margin ~13x the fp16 one, so the filter keeps 3x the candidate groups (6k instead of 2k:
// measured on the headline shape, 5k left 2 of 10 000 queries to the second chance, 4k 49, 3k 2330);
// shapes whose candidate count would not fit stay on the fp16 kernel.
static std::atomic<int> g_i8{-1};
static bool i8_on() {
    int v = g_i8.load();
    if (v < 0) {
        const char *e = getenv("VECGO_QTC_I8");
        v = (e && e[0] == '0') ? 0 : 1;
        g_i8.store(v);
    }
    return v != 0;
}
void set_i8(bool on) { g_i8.store(on ? 1 : 0); }
bool i8_state() { return i8_on() && use_pair(); }
static bool pq_i8_off() {   // VECGO_QTC_PQ_I8=0 keeps PQ on the fp16 kernel (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("VECGO_QTC_PQ_I8");
        v = (e && e[0] == '0') ? 1 : 0;
    }
    return v != 0;
}
static int candidates_i8(int64_t k) {
    static int tenths = -1;   // candidate groups per k, in tenths (VECGO_QTC_I8_KC, tuning / measurement)
    if (tenths < 0) {
        const char *e = getenv("VECGO_QTC_I8_KC");
        const int v = e ? atoi(e) : 0;
        tenths = v >= 20 && v <= 200 ? v : 60;
    }
    return k <= 16 ? 64 : (int)(k * tenths / 10);
}
static bool use_i8(const CodecParams &cp, int64_t rows, int64_t k) {
    const int qc = q_codec(cp);
    if (!i8_on() || !use_pair() || (qc != Q_SQ8 && qc != Q_INT4 && qc != Q_PQ)) return false;
    if (cp.dim % 128 != 0) return false;                                              // 128-byte k-blocks
    if (qc == Q_PQ && (cp.pq_dsub != 8 || cp.pq_k != 256 || pq_i8_off())) return false;   // 16 whole subspaces per k-block
    if (qc == Q_SQ8 && (cp.dim > 1024 || cp.row_bytes != cp.dim)) return false;      // resident query tile <= 128 KB, TMA row stride % 16
    if (qc == Q_INT4 && cp.row_bytes != cp.dim / 2) return false;
    const int kc = candidates_i8(k);
    return kc <= 2048 && rows / 32 >= 4 * (int64_t)kc;
}

template <int CODEC>
static vg_status launch_gemm(const CUtensorMap &mq, const KArgs &a, int64_t qtiles, int splits, cudaStream_t st) {
    const size_t sm = SMEM_BYTES;
    VG_CUDA(cudaFuncSetAttribute(qtc_kernel<CODEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid((unsigned)qtiles, (unsigned)splits);
    qtc_kernel<CODEC><<<grid, NTHREADS, sm, st>>>(mq, a);
    VG_LAUNCHED();
    return VG_OK;
}
template <int CODEC, bool THRESH = false>
static vg_status launch_gemm_pair(const CUtensorMap &mq, const KArgs &a, int64_t qtiles, int splits, cudaStream_t st, const CUtensorMap *mx = nullptr) {
    const size_t sm = pair::SMEM2_BYTES;
    VG_CUDA(cudaFuncSetAttribute(qtc2_kernel<CODEC, THRESH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid((unsigned)(2 * qtiles), (unsigned)splits);  // clusters of two CTAs along x (__cluster_dims__)
    qtc2_kernel<CODEC, THRESH><<<grid, NTHREADS, sm, st>>>(mq, mx ? *mx : mq, a);
    VG_LAUNCHED();
    return VG_OK;
}
// Rows per minimum group: as the Flat filter, but at most 128 for the CTA-pair kernel (a group stays inside one thread).
static int64_t qtc_group_rows(int64_t rows, int kc) {
    const int64_t G = tc::group_rows(rows, kc);
    return use_pair() ? std::min<int64_t>(G, 128) : G;
}
template <int CODEC>
static vg_status launch_exact(const EArgs &e, int64_t nq, cudaStream_t st) {
    const size_t sm = (((size_t)e.dim * 4 + 15) & ~(size_t)15) + topk_smem_bytes(1, e.C) + (size_t)e.kc * 8 +
                      (CODEC == Q_PQ ? (size_t)e.pq_m * 256 * 4 : sign_codec(CODEC) ? (size_t)e.words32 * 4 : (size_t)e.dim * 8);
    if (sm > 200 * 1024) return fail(VG_ERR_UNSUPPORTED, "dimension too large for the exact stage");
    VG_CUDA(cudaFuncSetAttribute(qtc_exact_kernel<CODEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    qtc_exact_kernel<CODEC><<<(unsigned)nq, 128, sm, st>>>(e);
    VG_LAUNCHED();
    return VG_OK;
}

template <int CODEC>
static vg_status launch_exact_list(const EArgs &e, int64_t nq, const uint2 *cand, const int *ccnt, const int *ovf, int slots, int cap,
                                   cudaStream_t st) {
    const size_t sm = (((size_t)e.dim * 4 + 15) & ~(size_t)15) + topk_smem_bytes(1, e.C) + (size_t)((slots + 1 + 3) & ~3) * 4 +
                      (CODEC == Q_PQ ? (size_t)e.pq_m * 256 * 4 : sign_codec(CODEC) ? (size_t)e.words32 * 4 : (size_t)e.dim * 8);
    if (sm > 200 * 1024) return fail(VG_ERR_UNSUPPORTED, "dimension too large for the exact stage");
    VG_CUDA(cudaFuncSetAttribute(qtc_exact_list_kernel<CODEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    qtc_exact_list_kernel<CODEC><<<(unsigned)nq, 128, sm, st>>>(e, cand, ccnt, ovf, slots, cap);
    VG_LAUNCHED();
    return VG_OK;
}
template <int CODEC>
static vg_status launch_thresh(const float *kth, const float *qn, const float *cq, const unsigned int *xmax_bits, float mid_norm, int dim,
                               int64_t nq, float *Ts, cudaStream_t st, const float *ea = nullptr) {
    qtc_thresh_kernel<CODEC><<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(kth, qn, cq, xmax_bits, mid_norm, dim, nq, Ts, ea);
    VG_LAUNCHED();
    return VG_OK;
}

// One chunk of queries through filter, select and exact stage (d_kth == nullptr), or through the threshold pass:
// thresholds from d_kth (the k-th best exact score known per query), threshold-collect GEMM, exact stage over the lists.
static vg_status search_chunk(const CodecParams &cp, const Prepared &pp, const SearchIO &io, int kc, int32_t *d_fail, cudaStream_t st,
                              const float *d_kth = nullptr) {
    const int qc = q_codec(cp);
    const int64_t nq = io.nq, rows = io.rows;
    const int64_t q_stride = io.q_stride ? io.q_stride : cp.dim;
    const int64_t qtiles = (nq + BMQ - 1) / BMQ, nq_pad = qtiles * BMQ;
    const bool pair_mode = use_pair();
    const int64_t G = qtc_group_rows(rows, kc);
    const int64_t groups = (rows + G - 1) / G;
    const bool thresh = d_kth != nullptr;
    if (thresh && !pair_mode) return fail(VG_ERR_UNSUPPORTED, "the threshold pass needs the CTA-pair kernel");
    const bool i8 = pair_mode && use_i8(cp, rows, io.k);
    const bool i8s = pair_mode && sign_codec(qc) && i8_on() && cp.dim % 128 == 0 && cp.row_bytes % 16 == 0;   // sign bits as +-1 bytes
    DevBuf a16, fq, cq, qn, mins, gids, gcnt, tau, eab;
    VG_TRY(a16.alloc((size_t)nq * pp.dimp * ((i8 || i8s) ? 1 : 2)));
    if (i8 || i8s) VG_TRY(eab.alloc((size_t)nq * 4));   // i8: certificate term (float); i8s: sum of the query's +-1 bytes (int32)
    const float *ea_p = i8 ? eab.as<float>() : nullptr;
    VG_TRY(fq.alloc((size_t)nq * 4));
    VG_TRY(cq.alloc((size_t)nq * 4));
    VG_TRY(qn.alloc((size_t)nq * 4));
    if (!thresh) {
        VG_TRY(mins.alloc((size_t)groups * nq_pad * 8));
        VG_TRY(gids.alloc((size_t)nq * kc * 4));
        VG_TRY(gcnt.alloc((size_t)nq * 4));
        VG_TRY(tau.alloc((size_t)nq * 4));
    }
    if (i8s) {
        prep_queries_sign_i8_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, st>>>(cp.q_words + io.q_index0 * cp.words32,
                                                                                      qc == Q_RABITQ ? cp.q_norms + io.q_index0 : nullptr, nq, cp.words32,
                                                                                      (int)cp.dim, a16.as<int8_t>(), fq.as<float>(), cq.as<float>(),
                                                                                      eab.as<int32_t>());
        VG_LAUNCHED();
    } else if (sign_codec(qc)) {
        prep_queries_sign_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, st>>>(cp.q_words + io.q_index0 * cp.words32,
                                                                                   qc == Q_RABITQ ? cp.q_norms + io.q_index0 : nullptr, nq,
                                                                                   cp.words32, (int)cp.dim, pp.dimp, pp.perm.as<int32_t>(),
                                                                                   a16.as<__half>(), fq.as<float>(), cq.as<float>());
        VG_LAUNCHED();
    } else {
    VG_TRY(tc::sqnorms(io.d_queries, nq, cp.dim, q_stride, qn.as<float>(), nullptr, st));
    if (i8 && qc == Q_INT4) {
        if (!pp.perm8.p) return fail(VG_ERR_STATE, "kind::i8 decode tables of the INT4 index were not prepared");
        prep_queries_i8_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, st>>>(io.d_queries, nq, q_stride, pp.dimp, pp.perm8.as<int32_t>(),
                                                                                 pp.wq8.as<float>(), pp.midp8.as<float>(), a16.as<int8_t>(), fq.as<float>(),
                                                                                 cq.as<float>(), eab.as<float>(), 8.0f);
    } else if (i8)
        prep_queries_i8_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, st>>>(io.d_queries, nq, q_stride, pp.dimp, pp.perm.as<int32_t>(),
                                                                                 pp.wq.as<float>(), pp.midp.as<float>(), a16.as<int8_t>(), fq.as<float>(),
                                                                                 cq.as<float>(), eab.as<float>(), qc == Q_PQ ? 0.0f : 128.0f);
    else
        prep_queries_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, st>>>(io.d_queries, nq, q_stride, pp.dimp, pp.perm.as<int32_t>(), pp.wq.as<float>(),
                                                                              pp.midp.as<float>(), a16.as<__half>(), fq.as<float>(), cq.as<float>());
    VG_LAUNCHED();
    }
    CUtensorMap mq, mx;
    if (i8 && qc == Q_SQ8) {
        VG_TRY(tc::tensor_map_2d_u8(&mq, a16.p, nq, pp.dimp, pp.dimp, 128, BM));
        VG_TRY(tc::tensor_map_2d_u8(&mx, cp.codes, rows, cp.dim, cp.row_bytes, 128, BN));
    } else if (i8) {
        VG_TRY(tc::tensor_map_2d_u8(&mq, a16.p, nq, pp.dimp, pp.dimp, 128, BM));
    } else if (i8s) {
        VG_TRY(tc::tensor_map_2d_u8(&mq, a16.p, nq, cp.dim, cp.dim, 128, BM));
    } else
    VG_TRY(tc::tensor_map_2d(&mq, true, a16.p, nq, pp.dimp, pp.dimp, BK, pair_mode ? BM : BMQ));
    // row splits: one CTA (pair) per SM (pair), whole waves
    const int64_t unit = pair_mode ? pair::TILE_ROWS : std::max<int64_t>(BN, G);
    const int64_t sms = pair_mode ? sm_count() / 2 : sm_count();
    const int64_t max_splits = std::max<int64_t>(1, rows / (4 * unit));
    int64_t splits = 1;
    double best = 0.0;
    for (int64_t s_ = 1; s_ <= sms && s_ <= max_splits; s_++) {
        const int64_t ctas = qtiles * s_, waves = (ctas + sms - 1) / sms;
        const double eff = (double)ctas / (double)(waves * sms);
        if (eff > best + 0.02) {
            best = eff;
            splits = s_;
        }
        if (eff >= 0.97) break;
    }
    int64_t rps = (rows + splits - 1) / splits;
    rps = (rps + unit - 1) / unit * unit;
    splits = (rows + rps - 1) / rps;
    KArgs a{};
    a.xn = qc == Q_RABITQ ? cp.norms : pp.xn.as<float>();  // BQ: unused
    a.half_dim = 0.5f * (float)cp.dim;
    a.mask = reinterpret_cast<const uint32_t *>(io.d_mask);
    a.fq = fq.as<float>();
    a.asum = i8s ? eab.as<int32_t>() : nullptr;
    a.nq = nq;
    a.rows = rows;
    a.rows_per_split = rps;
    a.kb = i8 ? pp.dimp / 128 : i8s ? (int)(cp.dim / 128) : pp.dimp / BK;
    a.cpg = (int)(G / 32);
    a.mins = mins.as<float2>();
    a.groups = groups;
    a.idx_mask = (uint32_t)(G - 1);
    a.keep_hi = ~31u;
    a.codes = cp.codes;
    a.row_bytes = cp.row_bytes;
    a.codebooks = cp.pq_codebooks;
    a.dsub_shift = 0;
    while ((1 << a.dsub_shift) < cp.pq_dsub) a.dsub_shift++;
    a.tiled = (qc == Q_PQ && (cp.variant & VG_VAR_PERM)) ? 1 : 0;
    tiles::Lists tl;  // active / skipped 256-row tiles (tile skipping: vg_tiles.cuh)
    if (io.d_mask && pair_mode && tiles::enabled()) {
        VG_TRY(tiles::build(a.mask, rows, tl, st));
        if (!thresh) VG_TRY(tiles::fill_skipped_groups(tl, (int)(pair::TILE_ROWS / G), groups, nq, a.mins, st));
        a.tile_list = tl.list;
        a.tile_count = tl.count;
    }
    if (thresh) {
        // ---- threshold pass: lists instead of the minima plane, no selection, no certificate
        const int slots = (int)splits * ((i8 && qc == Q_SQ8) ? 4 : 2);   // (split, epilogue group [SQ8I], column half)
        int64_t cap = ((int64_t)4 << 30) / std::max<int64_t>(1, nq * slots * 8);   // <= 4 GiB of lists per chunk
        cap = std::max<int64_t>(64, std::min<int64_t>(8192, cap));
        DevBuf Ts, cand, ccnt, ovf;
        VG_TRY(Ts.alloc((size_t)nq * 4));
        VG_TRY(cand.alloc((size_t)nq * slots * cap * 8));
        VG_TRY(ccnt.alloc((size_t)nq * slots * 4));
        VG_TRY(ovf.alloc((size_t)nq * 4));
        VG_CUDA(cudaMemsetAsync(ovf.p, 0, (size_t)nq * 4, st));
        const float *qn_p = qn.as<float>(), *cq_p = cq.as<float>();
        const unsigned int *xm = pp.xmax.as<unsigned int>();
        if (qc == Q_SQ8) VG_TRY(launch_thresh<Q_SQ8>(d_kth, qn_p, cq_p, xm, pp.mid_norm, (int)cp.dim, nq, Ts.as<float>(), st, ea_p));
        else if (qc == Q_INT4) VG_TRY(launch_thresh<Q_INT4>(d_kth, qn_p, cq_p, xm, pp.mid_norm, (int)cp.dim, nq, Ts.as<float>(), st, ea_p));
        else if (qc == Q_RABITQ) VG_TRY(launch_thresh<Q_RABITQ>(d_kth, qn_p, cq_p, xm, pp.mid_norm, (int)cp.dim, nq, Ts.as<float>(), st));
        else if (qc == Q_BQ) VG_TRY(launch_thresh<Q_BQ>(d_kth, qn_p, cq_p, xm, pp.mid_norm, (int)cp.dim, nq, Ts.as<float>(), st));
        else VG_TRY(launch_thresh<Q_PQ>(d_kth, qn_p, cq_p, xm, pp.mid_norm, (int)cp.dim, nq, Ts.as<float>(), st, ea_p));
        a.mins = nullptr;
        a.Ts = Ts.as<float>();
        a.cand = cand.as<uint2>();
        a.ccnt = ccnt.as<int>();
        a.ovf = ovf.as<int>();
        a.cap = (int)cap;
        a.slots = slots;
        if (i8 && qc == Q_INT4) VG_TRY((launch_gemm_pair<Q_INT4I, true>(mq, a, qtiles, (int)splits, st)));
        else if (i8 && qc == Q_PQ) VG_TRY((launch_gemm_pair<Q_PQI, true>(mq, a, qtiles, (int)splits, st)));
        else if (i8) VG_TRY((launch_gemm_pair<Q_SQ8I, true>(mq, a, qtiles, (int)splits, st, &mx)));
        else if (qc == Q_SQ8) VG_TRY((launch_gemm_pair<Q_SQ8, true>(mq, a, qtiles, (int)splits, st)));
        else if (qc == Q_INT4) VG_TRY((launch_gemm_pair<Q_INT4, true>(mq, a, qtiles, (int)splits, st)));
        else if (qc == Q_RABITQ && i8s) VG_TRY((launch_gemm_pair<Q_RABITQI, true>(mq, a, qtiles, (int)splits, st)));
        else if (qc == Q_BQ && i8s) VG_TRY((launch_gemm_pair<Q_BQI, true>(mq, a, qtiles, (int)splits, st)));
        else if (qc == Q_RABITQ) VG_TRY((launch_gemm_pair<Q_RABITQ, true>(mq, a, qtiles, (int)splits, st)));
        else if (qc == Q_BQ) VG_TRY((launch_gemm_pair<Q_BQ, true>(mq, a, qtiles, (int)splits, st)));
        else VG_TRY((launch_gemm_pair<Q_PQ, true>(mq, a, qtiles, (int)splits, st)));
        EArgs e = eargs_of(cp, rows);
        e.queries = io.d_queries;
        e.q_stride = q_stride;
        e.mask = io.d_mask;
        e.k = io.k;
        e.C = topk_capacity(io.k, 16);
        e.row_base = io.row_base;
        e.out_rows = io.d_rows;
        e.out_scores = io.d_scores;
        e.out_counts = io.d_counts;
        e.fail_flags = d_fail;
        if (sign_codec(qc)) {
            e.q_words = cp.q_words + io.q_index0 * cp.words32;
            e.q_norms = qc == Q_RABITQ ? cp.q_norms + io.q_index0 : nullptr;
            if (qc == Q_RABITQ) return launch_exact_list<Q_RABITQ>(e, nq, a.cand, a.ccnt, a.ovf, slots, (int)cap, st);
            return launch_exact_list<Q_BQ>(e, nq, a.cand, a.ccnt, a.ovf, slots, (int)cap, st);
        }
        if (qc == Q_SQ8) return launch_exact_list<Q_SQ8>(e, nq, a.cand, a.ccnt, a.ovf, slots, (int)cap, st);
        if (qc == Q_INT4) return launch_exact_list<Q_INT4>(e, nq, a.cand, a.ccnt, a.ovf, slots, (int)cap, st);
        return launch_exact_list<Q_PQ>(e, nq, a.cand, a.ccnt, a.ovf, slots, (int)cap, st);
    }
    const bool prof = g_prof.load() != 0;
    cudaEvent_t g_ev[2] = {nullptr, nullptr};  // per call: the event pair lives on this call's device and stream
    if (prof) {
        VG_CUDA(cudaEventCreate(&g_ev[0]));
        VG_CUDA(cudaEventCreate(&g_ev[1]));
        VG_CUDA(cudaEventRecord(g_ev[0], st));
    }
    if (pair_mode) {
        if (i8 && qc == Q_INT4) VG_TRY((launch_gemm_pair<Q_INT4I>(mq, a, qtiles, (int)splits, st)));
        else if (i8 && qc == Q_PQ) VG_TRY((launch_gemm_pair<Q_PQI>(mq, a, qtiles, (int)splits, st)));
        else if (i8) VG_TRY((launch_gemm_pair<Q_SQ8I>(mq, a, qtiles, (int)splits, st, &mx)));
        else if (qc == Q_SQ8) VG_TRY(launch_gemm_pair<Q_SQ8>(mq, a, qtiles, (int)splits, st));
        else if (qc == Q_INT4) VG_TRY(launch_gemm_pair<Q_INT4>(mq, a, qtiles, (int)splits, st));
        else if (qc == Q_RABITQ && i8s) VG_TRY(launch_gemm_pair<Q_RABITQI>(mq, a, qtiles, (int)splits, st));
        else if (qc == Q_BQ && i8s) VG_TRY(launch_gemm_pair<Q_BQI>(mq, a, qtiles, (int)splits, st));
        else if (qc == Q_RABITQ) VG_TRY(launch_gemm_pair<Q_RABITQ>(mq, a, qtiles, (int)splits, st));
        else if (qc == Q_BQ) VG_TRY(launch_gemm_pair<Q_BQ>(mq, a, qtiles, (int)splits, st));
        else VG_TRY(launch_gemm_pair<Q_PQ>(mq, a, qtiles, (int)splits, st));
    } else {
        if (sign_codec(qc)) return fail(VG_ERR_UNSUPPORTED, "the RaBitQ / BQ filter needs the CTA-pair kernel");
        if (qc == Q_SQ8) VG_TRY(launch_gemm<Q_SQ8>(mq, a, qtiles, (int)splits, st));
        else if (qc == Q_INT4) VG_TRY(launch_gemm<Q_INT4>(mq, a, qtiles, (int)splits, st));
        else VG_TRY(launch_gemm<Q_PQ>(mq, a, qtiles, (int)splits, st));
    }
    if (prof) VG_CUDA(cudaEventRecord(g_ev[1], st));
    VG_TRY(tc::select_groups(a.mins, groups, nq, kc, G, tau.as<float>(), gids.as<uint32_t>(), gcnt.as<int32_t>(), st));
    EArgs e = eargs_of(cp, rows);
    e.queries = io.d_queries;
    e.q_stride = q_stride;
    e.cand = gids.as<uint32_t>();
    e.gcnt = gcnt.as<int32_t>();
    e.kc = kc;
    e.G = (int)G;
    e.tau = tau.as<float>();
    e.qn = qn.as<float>();
    e.cq = cq.as<float>();
    e.ea = ea_p;
    e.mid_norm = pp.mid_norm;
    e.xmax_bits = pp.xmax.as<unsigned int>();
    e.mask = io.d_mask;
    e.k = io.k;
    e.C = topk_capacity(io.k, 16);
    e.row_base = io.row_base;
    e.out_rows = io.d_rows;
    e.out_scores = io.d_scores;
    e.out_counts = io.d_counts;
    e.fail_flags = d_fail;
    if (sign_codec(qc)) {
        e.q_words = cp.q_words + io.q_index0 * cp.words32;
        e.q_norms = qc == Q_RABITQ ? cp.q_norms + io.q_index0 : nullptr;
        if (qc == Q_RABITQ) VG_TRY(launch_exact<Q_RABITQ>(e, nq, st));
        else VG_TRY(launch_exact<Q_BQ>(e, nq, st));
    } else if (qc == Q_SQ8) VG_TRY(launch_exact<Q_SQ8>(e, nq, st));
    else if (qc == Q_INT4) VG_TRY(launch_exact<Q_INT4>(e, nq, st));
    else VG_TRY(launch_exact<Q_PQ>(e, nq, st));
    // the temporaries above go back to the stream-ordered pool (freed in stream order): no synchronisation needed for them
    if (prof) {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof_pending.emplace_back(g_ev[0], g_ev[1]);
        if (g_prof_pending.size() > 4096) settle_profile_events();   // bounded: a forgotten profile switch must not leak events
    }
    return VG_OK;
}

// Candidate groups per query of the first pass and of the second chance (twice as many: a wider gap between the k-th
// best and tau); 0 when the segment is too short for a second pass to differ.
static int kc_for(const CodecParams &cp, int64_t rows, int k, int kc_scale) {
    const int kc = use_i8(cp, rows, k) ? candidates_i8(k) : candidates_for(k);
    if (kc_scale <= 1) return kc;
    const int kc2 = kc * kc_scale;
    if (kc2 > 4096 || rows / 32 < 2 * (int64_t)kc2) return 0;
    return kc2;
}
bool second_chance_possible(const CodecParams &cp, int64_t rows, int64_t k) { return kc_for(cp, rows, (int)k, 2) > 0; }

// Enqueue only: filter, select, exact stage, certificate flags (d_fail[q] = 1 where the proof did not hold).  Nothing
// here waits for the device unless the GEMM profile is on.
vg_status enqueue(const CodecParams &cp, const Prepared &pp, const SearchIO &io, int kc_scale, int32_t *d_fail, cudaStream_t st) {
    if (!pp.ready) return fail(VG_ERR_STATE, "decode-GEMM filter state was not prepared");
    const int kc = kc_for(cp, io.rows, io.k, kc_scale);
    if (kc <= 0) return fail(VG_ERR_UNSUPPORTED, "segment too short for a second filter pass");
    const int64_t G = qtc_group_rows(io.rows, kc);
    const int64_t groups = (io.rows + G - 1) / G;
    // the [queries][groups] minima buffer is kept under 8 GiB: longer batches go through in chunks of whole query tiles
    // (the headline shape — 10k queries x 78k groups = 6.4 GB — is one launch: one tail instead of two)
    int64_t chunk = std::max<int64_t>(BMQ, ((8ll << 30) / (groups * 8)) / BMQ * BMQ);
    for (int64_t q0 = 0; q0 < io.nq; q0 += chunk) {
        SearchIO part = io;
        part.nq = std::min(chunk, io.nq - q0);
        part.q_stride = io.q_stride ? io.q_stride : cp.dim;
        part.d_queries = io.d_queries + q0 * part.q_stride;
        part.q_index0 = io.q_index0 + q0;
        part.d_rows = io.d_rows + q0 * io.k;
        part.d_scores = io.d_scores + q0 * io.k;
        part.d_counts = io.d_counts + q0;
        VG_TRY(search_chunk(cp, pp, part, kc, d_fail + q0, st));
    }
    if (kc_scale <= 1) g_queries.fetch_add((uint64_t)io.nq);
    return VG_OK;
}
void count_fallbacks(uint64_t n) { g_fallbacks.fetch_add(n); }

// Second pass for queries whose certificate failed: d_kth[q] = the k-th best exact score the first pass found (an upper
// bound of the true k-th best, +inf if it found fewer than k rows).  Lists every row whose filter score could still beat
// it, scores all of them exactly: the result needs no certificate; d_fail[q] = 1 only when a list overflowed.
vg_status enqueue_threshold(const CodecParams &cp, const Prepared &pp, const SearchIO &io, const float *d_kth, int32_t *d_fail,
                            cudaStream_t st) {
    if (!pp.ready) return fail(VG_ERR_STATE, "decode-GEMM filter state was not prepared");
    const int kc = kc_for(cp, io.rows, io.k, 1);
    // lists are sized per chunk (<= 4 GiB): chunks of at most 64 query tiles
    const int64_t chunk = 64 * BMQ;
    for (int64_t q0 = 0; q0 < io.nq; q0 += chunk) {
        SearchIO part = io;
        part.nq = std::min(chunk, io.nq - q0);
        part.q_stride = io.q_stride ? io.q_stride : cp.dim;
        part.d_queries = io.d_queries + q0 * part.q_stride;
        part.q_index0 = io.q_index0 + q0;
        part.d_rows = io.d_rows + q0 * io.k;
        part.d_scores = io.d_scores + q0 * io.k;
        part.d_counts = io.d_counts + q0;
        VG_TRY(search_chunk(cp, pp, part, kc, d_fail + q0, st, d_kth + q0));
    }
    return VG_OK;
}
bool threshold_pass_possible(const CodecParams &cp, int64_t rows, int64_t nq) { return use_pair() && rows >= 8192 && nq >= 1 && q_codec(cp) >= 0; }
vg_status gather_kth(const float *d_scores, const int32_t *d_counts, const int32_t *d_idx, int64_t n, int k, float *d_kth, cudaStream_t st) {
    if (n <= 0) return VG_OK;
    gather_kth_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_scores, d_counts, d_idx, n, k, d_kth);
    VG_LAUNCHED();
    return VG_OK;
}

// Host-synchronous form: enqueue and read the certificate flags back.  `failed` lists the queries without a proof.
vg_status search(const CodecParams &cp, const Prepared &pp, const SearchIO &io, std::vector<int32_t> &failed, cudaStream_t st) {
    failed.clear();
    DevBuf failb;
    VG_TRY(failb.alloc((size_t)io.nq * 4));
    VG_TRY(enqueue(cp, pp, io, 1, failb.as<int32_t>(), st));
    std::vector<int32_t> h_fail((size_t)io.nq);
    VG_CUDA(cudaMemcpyAsync(h_fail.data(), failb.p, (size_t)io.nq * 4, cudaMemcpyDeviceToHost, st));
    VG_CUDA(cudaStreamSynchronize(st));
    for (int64_t q = 0; q < io.nq; q++)
        if (h_fail[(size_t)q]) failed.push_back((int32_t)q);
    g_fallbacks.fetch_add((uint64_t)failed.size());
    return VG_OK;
}

}  // namespace qtc
}  // namespace vg
