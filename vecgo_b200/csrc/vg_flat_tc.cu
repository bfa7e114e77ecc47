// vg_flat_tc.cu — exact Flat L2 / dot search as a tcgen05 (TF32) GEMM *filter* fed by TMA,
// followed by an exact re-check in the reference's own summation order.
//
// Why a filter.  flat.(*Segment).Search (internal/segment/flat/segment.go:690-697) calls
// simd.SquaredL2 / simd.Dot per (query,row); BASELINE's north_star wants that dense
// Q x N x d contraction on the 5th-gen tensor cores AND top-k ids identical to the SIMD path.
// Tensor cores give q.x only to TF32 accuracy, so they are used to shrink N rows to a few
// dozen candidates per query; the survivors are scored by the exact AVX-512-order kernel
// (same arithmetic as Segment.Rerank) and a certificate proves that nothing outside the
// candidate set could have entered the top-k.  Queries whose certificate fails are re-run on
// the exact CUDA-core scan.
//
//   s(q,x) = ||x||^2 - 2 q.x   (L2; the per-query constant ||q||^2 is dropped)
//   s(q,x) = -q.x              (dot / cosine, descending in the reference)
//
// One GEMM pass with a branch-free epilogue (no per-query heap inside the GEMM), then a tiny exact stage:
//   gemm    every thread (= one query) reduces s over groups of G consecutive rows to (m1, m2): the smallest
//           value — with the row's index inside the group written into its low mantissa bits, so the
//           minimum also names its row — and the second smallest.  FFMA + LOP3 + 3 FMNMX per element.
//   select  tau(q) = the kc-th smallest m1 and the kc groups that reach it.  Those kc groups each hold
//           a row with s <= tau, so at least kc rows lie at or below tau; with many more groups than kc
//           the bound is nearly tight.  A selected group whose m2 is also <= tau is "crowded".
//   scan    the arg-min row of every selected group (and ALL rows of a crowded group) is scored EXACTLY
//           in simd.SquaredL2 / simd.Dot order; the heap order (score, row) keeps the best k.
//
// Certificate (per query).  Every row that was not scored is either in an unselected group (s >= m1 >= tau)
// or a non-minimal row of an uncrowded selected group (s >= m2 > tau).  With E >= |s_exact - s_approx|
// for every row (bound below) such a row has s_exact > tau - E.  If the exact k-th best scored row
// satisfies s_exact(e_k) < tau - E, no other row can tie or beat it: the result IS the exact top-k of
// the segment (ties by row id included).
//
// Error bound.  kind::tf32 keeps 10 explicit mantissa bits of each fp32 operand:
// |fl_tf32(a) - a| <= 2^-10 |a|, so |q.x - (q.x)_tc| <= (2^-9 + 2^-20) sum|q_i x_i|
// <= 2^-9 (1 + 2^-11) ||q|| ||x||, plus fp32 accumulation (<= d 2^-23 ||q|| ||x||).
// With the factor 2 of the L2 form:  E = c1 ||q|| max||x|| + c2 (||q||^2 + max||x||^2),
// c1 = 2^-8 * 1.125 (L2) or 2^-9 * 1.125 (dot), c2 = 2^-14 (norm rounding, accumulation
// slack), plus 2^-(23-b) max|s| for the b = log2(G) mantissa bits that carry the row index.
// tests/test_gpu_flat_tc.py measures the realised error against E.
//
// GEMM kernel (one CTA = 256 queries x a contiguous row range, 320 threads):
//   warp 0     TMA producer (cp.async.bulk.tensor.2d, 128B swizzle).  dim <= 128: the 256 x dim
//              query tile is loaded ONCE and stays resident (128 KB), only 128-row B tiles stream
//              through a 4-stage mbarrier ring; larger dims stream A and B k-blocks together.
//   warp 1     MMA issuer: tcgen05.mma.cta_group::1.kind::tf32, M=128 x N=128 x K=8, two M halves
//              per B tile, fp32 accumulators in TMEM (2 stages x 2 halves x 128 columns = 512)
//   warps 2-9  epilogue: one thread = one query (= one TMEM lane); tcgen05.ld 32 columns at a
//              time, 32 independent FFMA, four (m1, m2) accumulators, one 8-byte store per group.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <vector>

#include "vg_flat_tc.cuh"
#include "vg_tiles.cuh"
#include "vg_scan.cuh"
#include "vg_tc_ptx.cuh"
#include "vg_topk.cuh"

namespace vg {
namespace tc {

constexpr int BM = 128;    // UMMA M (= TMEM lanes)
constexpr int BMQ = 256;   // queries per CTA: two UMMA M halves that share every B tile
constexpr int BK = 32;     // floats per k-block: one 128-byte swizzle atom
constexpr int NTHREADS = 320;

// kind::tf32 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=TF32 (2 at bits 7-9 and
// 10-12), both K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------ kernel
constexpr int BN = 128;        // database rows per tile (UMMA N)
constexpr int STAGES = 4;
constexpr int A_KB_BYTES = BMQ * BK * 4;   // 32 KB: one k-block of the 256-query tile (two UMMA M=128 halves)
constexpr int B_KB_BYTES = BN * BK * 4;    // 16 KB
constexpr int MAX_RES_KB = 4;              // query tile kept resident in shared memory when dim <= 128

struct Args {
    const float *xn;        // [rows] ||x||^2 (L2) or nullptr (dot)
    const uint32_t *mask;   // optional row bitmap, read as 32-bit words (bit = 1 keeps the row)
    int64_t nq, nq_pad, rows, rows_per_split;
    int kb;                 // k-blocks = ceil(dim / 32)
    int cpg;                // 32-row chunks per minimum group (G / 32)
    uint32_t row_base;
    float2 *mins;           // out: [nq_pad][groups] (m1 with the row index in its low log2(G) bits, m2)
    int64_t groups;
    uint32_t idx_mask;      // G - 1
};

template <bool RESIDENT>
struct Smem {
    // RESIDENT: [A: kb x 32 KB][B ring: STAGES x 16 KB];  streaming: [ring: STAGES x (32 KB A + 16 KB B)]
    static constexpr int STAGE_BYTES = RESIDENT ? B_KB_BYTES : A_KB_BYTES + B_KB_BYTES;
    static constexpr size_t OFF_RING = RESIDENT ? (size_t)MAX_RES_KB * A_KB_BYTES : 0;
    static constexpr size_t OFF_XN = OFF_RING + (size_t)STAGES * STAGE_BYTES;
    static constexpr size_t OFF_BAR = OFF_XN + (size_t)2 * BN * 4;
    static constexpr size_t TOTAL = OFF_BAR + (size_t)(2 * STAGES + 5) * 8 + 16;
};

// GEMM + group minima.
template <bool RESIDENT, bool IS_DOT>
__global__ void __launch_bounds__(NTHREADS, 1)
flat_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, Args A) {
    using S = Smem<RESIDENT>;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    // SWIZZLE_128B tiles need 1024-byte aligned bases; the dynamic segment only guarantees 16
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * BMQ;
    const int split = blockIdx.y;
    const int64_t row_begin = (int64_t)split * A.rows_per_split;
    int64_t row_end = row_begin + A.rows_per_split;
    if (row_end > A.rows) row_end = A.rows;
    const int ntiles = row_end > row_begin ? (int)((row_end - row_begin + BN - 1) / BN) : 0;

    const uint32_t s_base = smem_u32(smem);
    const uint32_t ring = s_base + (uint32_t)S::OFF_RING;
    const uint32_t bar0 = s_base + (uint32_t)S::OFF_BAR;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
    const uint32_t afull_bar = bar0 + 8u * (2 * STAGES + 4);
    constexpr uint32_t TMEM_COLS = 512;  // 2 accumulator stages x 2 query halves x 128 columns

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 256);
        }
        mbar_init(afull_bar, 1);
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

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0 && ntiles > 0) {
            if (RESIDENT) {
                mbar_expect_tx(afull_bar, (uint32_t)A.kb * A_KB_BYTES);
                for (int kb = 0; kb < A.kb; kb++) tma_load_2d(s_base + kb * A_KB_BYTES, &map_q, kb * BK, q0, afull_bar);
            }
            uint32_t it = 0;
            for (int t = 0; t < ntiles; t++) {
                const int n0 = (int)(row_begin + (int64_t)t * BN);
                for (int kb = 0; kb < A.kb; kb++, it++) {
                    const int st = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(empty_bar(st), ph ^ 1);
                    mbar_expect_tx(full_bar(st), S::STAGE_BYTES);
                    const uint32_t sa = ring + st * S::STAGE_BYTES;
                    if (!RESIDENT) {
                        tma_load_2d(sa, &map_q, kb * BK, q0, full_bar(st));
                        tma_load_2d(sa + A_KB_BYTES, &map_x, kb * BK, n0, full_bar(st));
                    } else {
                        tma_load_2d(sa, &map_x, kb * BK, n0, full_bar(st));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && ntiles > 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            if (RESIDENT) {
                mbar_wait(afull_bar, 0);
                tc_fence_after();
            }
            uint32_t it = 0;
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
                    const uint32_t sa = ring + st * S::STAGE_BYTES;
                    const uint32_t a_addr = RESIDENT ? s_base + kb * A_KB_BYTES : sa;
                    const uint32_t b_addr = RESIDENT ? sa : sa + A_KB_BYTES;
                    const uint64_t bdesc = make_sdesc(b_addr);
#pragma unroll
                    for (int h = 0; h < 2; h++) {  // the two 128-query halves share the B tile
                        const uint64_t adesc = make_sdesc(a_addr + h * (BM * BK * 4));
#pragma unroll
                        for (int k = 0; k < BK / 8; k++)  // 8 tf32 = 32 bytes per UMMA: advance inside the swizzle atom
                            umma_tf32(d_tmem + (uint32_t)(h * BN), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                      (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar(st));
                }
                umma_commit(tfull_bar(as));
            }
        }
    } else {
        // ===================== epilogue: warps 2..9, one thread per query =====================
        const int quad = warp & 3;            // TMEM lanes 32*quad .. +31 are the ones this warp may read
        const int half = (warp - 2) >> 2;     // which 128-query half (= accumulator column block)
        const int slot = half * BM + quad * 32 + lane;
        const int et = (warp - 2) * 32 + lane;
        const int64_t q = (int64_t)q0 + slot;
        float *xs = reinterpret_cast<float *>(smem + S::OFF_XN);
        const float INF = __int_as_float(0x7f800000);
        const float BIG = 3.0e38f;            // padding / masked rows: finite, so index bits can be written into it
        float g1 = BIG, g2 = BIG;             // running (smallest, second smallest) of the current group
        int cc = 0;
        for (int t = 0; t < ntiles; t++) {
            const int as = t & 1;
            const uint32_t aph = (t >> 1) & 1;
            const int64_t n0 = row_begin + (int64_t)t * BN;
            float *xt = xs + as * BN;
            if (et < BN) {
                const int64_t row = n0 + et;
                xt[et] = (row < row_end) ? (IS_DOT ? 0.0f : __ldg(A.xn + row)) : BIG;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 2 * BN + half * BN);
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)(c * 32), v);
                // rows n0 + 32c .. +31 start on a multiple of 32: one mask word covers the chunk
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
                    for (int i = 0; i < 4; i++) {
                        const float dot = __uint_as_float(v[j4 * 4 + i]);
                        s[j4 * 4 + i] = IS_DOT ? __fsub_rn(xx[i], dot) : __fmaf_rn(-2.0f, dot, xx[i]);  // dot: xx = 0 | BIG (padding)
                    }
                }
                if (mw != 0xFFFFFFFFu) {
#pragma unroll
                    for (int j = 0; j < 32; j++) s[j] = (mw >> j) & 1u ? s[j] : BIG;
                }
                {
                    // index of the chunk's first row inside its group (rows are group-aligned per CTA range)
                    const uint32_t cidx = (uint32_t)(n0 + c * 32) & A.idx_mask;
                    float a1[4] = {BIG, BIG, BIG, BIG}, a2[4] = {BIG, BIG, BIG, BIG};
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float v1 = __uint_as_float((__float_as_uint(s[j]) & ~A.idx_mask) | (cidx + j));
                        a2[j & 3] = fminf(a2[j & 3], fmaxf(a1[j & 3], v1));
                        a1[j & 3] = fminf(a1[j & 3], v1);
                    }
                    // merge (m1, m2) pairs: m2 = min(max(a1, b1), a2, b2)
                    const float p1 = fminf(a1[0], a1[1]), p2 = fminf(fmaxf(a1[0], a1[1]), fminf(a2[0], a2[1]));
                    const float r1 = fminf(a1[2], a1[3]), r2 = fminf(fmaxf(a1[2], a1[3]), fminf(a2[2], a2[3]));
                    const float c1 = fminf(p1, r1), c2 = fminf(fmaxf(p1, r1), fminf(p2, r2));
                    g2 = fminf(fmaxf(g1, c1), fminf(g2, c2));
                    g1 = fminf(g1, c1);
                    if (++cc == A.cpg) {
                        const int64_t gid = (n0 + c * 32) / (32 * (int64_t)A.cpg);
                        if (gid < A.groups) A.mins[q * A.groups + gid] = make_float2(g1, g2);  // chunks past the last row: no group
                        g1 = BIG;
                        g2 = BIG;
                        cc = 0;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(as));
        }
        if (cc > 0 && ntiles > 0) {  // partial last group of this row range
            const int64_t last_chunk_row = row_begin + (int64_t)ntiles * BN - 32;
            const int64_t gid = last_chunk_row / (32 * (int64_t)A.cpg);
            if (gid < A.groups) A.mins[q * A.groups + gid] = make_float2(g1, g2);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ CTA-pair fp16 variant
// Same filter on a cluster of two CTAs with tcgen05.mma.cta_group::2.kind::f16, M = 256 x N = 256 (see vg_quant_tc.cu
// for the protocol): CTA r owns queries q0 + 128 r .. and TMA-loads rows n0 + 128 r .. of an fp16 SHADOW of the
// vectors (x * 2^sx, built once per index; the float32 rows stay the source of the exact stage).  kind::f16 runs at
// twice the TF32 rate, rounds to nearest (2^-11 relative per operand instead of TF32's 2^-10 truncation) and moves
// half the operand bytes; queries are scaled per query so that max|q_i| lands in [2^11, 2^12).
namespace pairf {
using namespace vg::tc::pair;
constexpr int BKH = 64;                  // halves per k-block (128 bytes)
constexpr int STAGES2 = 6;
constexpr int A2_BYTES = BM * BKH * 2;   // 16 KB
constexpr int B2_BYTES = BN * BKH * 2;   // 16 KB
constexpr int STAGE2_BYTES = A2_BYTES + B2_BYTES;
constexpr int TILE_ROWS = 2 * BN;
constexpr size_t OFF_XN2 = (size_t)STAGES2 * STAGE2_BYTES;
constexpr size_t OFF_BAR2 = OFF_XN2 + (size_t)2 * TILE_ROWS * 4;
constexpr size_t SMEM2_BYTES = OFF_BAR2 + (size_t)(2 * STAGES2 + 4) * 8 + 16 + 1024;
}  // namespace pairf

struct Args2 {
    const float *xn;        // [rows] ||x||^2 (L2) or nullptr (dot)
    const uint32_t *mask;
    const int32_t *tile_list;   // tile skipping (vg_tiles.cuh): active 256-row tiles and their number, or nullptr
    const int32_t *tile_count;
    const float *fq;        // [nq] -2 (L2) or -1 (dot) / (query scale x database scale)
    int64_t nq, rows, rows_per_split;
    int kb;                 // k-blocks = dimp / 64
    int cpg;                // 32-row chunks per group (<= 4)
    float2 *mins;
    int64_t groups;
    uint32_t idx_mask, keep_hi;
};

// EPW = epilogue warps per CTA: 8 (thread = query x 128 columns of the tile) or 16 (query x 64 columns).  With few k-blocks
// per tile (d <= 256) the epilogue sets the pace and runs at the latency of its dependent min chains; four warps per
// scheduler instead of two hide that latency (groups are then at most 64 rows).
template <bool IS_DOT, int EPW, bool MINONLY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((2 + EPW) * 32, 1)
flat2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, Args2 A) {
    using namespace pairf;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int q0 = (int)(blockIdx.x >> 1) * BMQ + (int)rank * BM;
    const int split = blockIdx.y;
    const int64_t row_begin = (int64_t)split * A.rows_per_split;
    int64_t row_end = row_begin + A.rows_per_split;
    if (row_end > A.rows) row_end = A.rows;
    int ntiles = row_end > row_begin ? (int)((row_end - row_begin + TILE_ROWS - 1) / TILE_ROWS) : 0;
    const int32_t *tiles = nullptr;  // this split's slice of the active-tile list
    if (A.tile_list) {
        const int n_act = __ldg(A.tile_count);
        const int per = (n_act + (int)gridDim.y - 1) / (int)gridDim.y;
        const int first = split * per;
        ntiles = n_act - first < per ? n_act - first : per;
        if (ntiles < 0) ntiles = 0;
        tiles = A.tile_list + first;
        row_end = A.rows;
    }
    auto tile_row0 = [&](int t) -> int64_t {
        return tiles ? (int64_t)__ldg(tiles + t) * TILE_ROWS : row_begin + (int64_t)t * TILE_ROWS;
    };

    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar0 = s_base + (uint32_t)OFF_BAR2;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES2 + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES2 + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES2 + 2 + s); };
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES2; s++) {
            mbar_init(full_bar(s), 1);   // the leader's expect_tx arrive; all four TMA loads of the pair count their bytes here
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), EPW + EPW);
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
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer: this CTA's query half and row half of every k-block =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = 0; t < ntiles; t++) {
                const int n0 = (int)tile_row0(t) + (int)rank * BN;
                for (int kb = 0; kb < A.kb; kb++, it++) {
                    const int st = it % STAGES2;
                    const uint32_t ph = (it / STAGES2) & 1;
                    mbar_wait(empty_bar(st), ph ^ 1);
                    if (leader) mbar_expect_tx(full_bar(st), 2 * STAGE2_BYTES);
                    const uint32_t sa = s_base + st * STAGE2_BYTES;
                    tma_load_2d_pair(sa, &map_q, kb * BKH, q0, full_bar(st));
                    tma_load_2d_pair(sa + A2_BYTES, &map_x, kb * BKH, n0, full_bar(st));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16_pair();
            uint32_t it = 0;
            for (int t = 0; t < ntiles; t++) {
                const int as = t & 1;
                const uint32_t aph = (t >> 1) & 1;
                mbar_wait_cluster(tempty_bar(as), aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * TILE_ROWS);
                for (int kb = 0; kb < A.kb; kb++, it++) {
                    const int st = it % STAGES2;
                    const uint32_t ph = (it / STAGES2) & 1;
                    mbar_wait_cluster(full_bar(st), ph);
                    tc_fence_after();
                    const uint32_t sa = s_base + st * STAGE2_BYTES;
                    const uint64_t adesc = make_sdesc(sa), bdesc = make_sdesc(sa + A2_BYTES);
#pragma unroll
                    for (int k = 0; k < BKH / 16; k++)
                        umma_f16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit_pair(empty_bar(st));
                }
                umma_commit_pair(tfull_bar(as));
            }
        }
    } else {
        // ===================== epilogue: warps 2 .. 2+EPW-1; thread = one query x CPT columns of the tile =====================
        constexpr int CPT = TILE_ROWS / (EPW / 4);  // columns (rows of the tile) per thread: 128 or 64
        const int quad = warp & 3;
        const int colhalf = (warp - 2) >> 2;        // which CPT-column part
        const int et = (warp - 2) * 32 + lane;
        const int64_t q = (int64_t)q0 + quad * 32 + lane;
        float *xs = reinterpret_cast<float *>(smem + OFF_XN2);
        const float BIG = 3.0e38f;
        const float fq = q < A.nq ? __ldg(A.fq + q) : 0.0f;
        const uint32_t keep_hi = A.keep_hi;
        float g1 = BIG, g2 = BIG;
        int cc = 0;
        // row norms of a tile are fetched one tile ahead (global-load latency out of the per-tile critical path): the value
        // for tile t + 1 is loaded at the top of tile t and stored after tile t's columns are reduced
        auto xn_of = [&](int t_) {
            const int64_t n0 = tile_row0(t_);
            const int64_t row = n0 + et;
            return (row < row_end) ? (IS_DOT ? 0.0f : __ldg(A.xn + row)) : BIG;
        };
        if (ntiles > 0 && et < TILE_ROWS) xs[et] = xn_of(0);
        for (int t = 0; t < ntiles; t++) {
            const int as = t & 1;
            const uint32_t aph = (t >> 1) & 1;
            const int64_t n0 = tile_row0(t);
            float *xt = xs + as * TILE_ROWS;
            asm volatile("bar.sync 1, %0;" ::"n"(EPW * 32) : "memory");
            const float xn_next = (t + 1 < ntiles && et < TILE_ROWS) ? xn_of(t + 1) : 0.0f;
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * TILE_ROWS + colhalf * CPT);
            const int64_t nh = n0 + colhalf * CPT;
            // the TMEM load of chunk c + 1 is issued before chunk c is reduced (tcgen05.wait::ld at the top of the next
            // iteration): with few k-blocks per tile (d = 128) the epilogue, not the MMA, sets the pace
            uint32_t vn[32];
            tmem_ld32(taddr, vn);
#pragma unroll 1
            for (int c = 0; c < CPT / 32; c++) {
                uint32_t v[32];
                uint32_t mw = 0xFFFFFFFFu;
                if (A.mask) mw = (nh + c * 32 < A.rows) ? __ldg(A.mask + ((nh + c * 32) >> 5)) : 0u;
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = vn[j];
                if (c + 1 < CPT / 32) tmem_ld32(taddr + (uint32_t)((c + 1) * 32), vn);
                float s[32];
                const float4 *x4 = reinterpret_cast<const float4 *>(xt + colhalf * CPT + c * 32);
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
                if constexpr (MINONLY) {
                    // Short contractions (d <= 256) over long segments: the epilogue is bound by the ALU pipe (LOP3 and FMNMX issue every other
                    // cycle per scheduler), so only the group MINIMUM is kept — one FFMA + one FMNMX per element — and stored
                    // as (m, m): the selection then treats every selected group as crowded and the exact stage scores all of
                    // its (<= 64) rows.  Unscored rows are exactly the rows of unselected groups, s >= m >= tau.
                    float a1[4] = {BIG, BIG, BIG, BIG};
#pragma unroll
                    for (int j = 0; j < 32; j++) a1[j & 3] = fminf(a1[j & 3], s[j]);
                    const float c1 = fminf(fminf(a1[0], a1[1]), fminf(a1[2], a1[3]));
                    g1 = fminf(g1, c1);
                    g2 = g1;
                } else {
                float a1[4] = {BIG, BIG, BIG, BIG}, a2[4] = {BIG, BIG, BIG, BIG};
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    uint32_t vb;  // (bits & ~31) | j as one LOP3
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
                }
                if (++cc == A.cpg) {
                    const int64_t gid = (nh + c * 32) / (32 * (int64_t)A.cpg);
                    if (gid < A.groups) A.mins[q * A.groups + gid] = make_float2(g1, g2);
                    g1 = BIG;
                    g2 = BIG;
                    cc = 0;
                }
            }
            if (t + 1 < ntiles && et < TILE_ROWS) xs[((t + 1) & 1) * TILE_ROWS + et] = xn_next;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_bar(as), 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// fp16 shadow of the vectors: x16[r][p] = half(x[r][p] * 2^sx) for p < dim, 0 for the padding up to dimp.
__global__ void __launch_bounds__(256) shadow16_kernel(const float *x, int64_t rows, int64_t dim, int dimp, int sx_exp, __half *x16) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * dimp) return;
    const int64_t r = i / dimp;
    const int p = (int)(i - r * dimp);
    x16[i] = __float2half_rn(p < dim ? __fmul_rn(x[r * dim + p], ldexpf(1.0f, sx_exp)) : 0.0f);
}
// One warp per query: a16 = half(q * 2^e), e such that max|q_i| lands in [2^11, 2^12); f_q = -(2 | 1) / (2^e 2^sx).
__global__ void __launch_bounds__(256) prep_queries16_kernel(const float *queries, int64_t nq, int64_t q_stride, int64_t dim, int dimp, int sx_exp,
                                                             int is_dot, __half *a16, float *fq) {
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const float *qv = queries + q * q_stride;
    float mx = 0.0f;
    for (int p = lane; p < dim; p += 32) mx = fmaxf(mx, fabsf(qv[p]));
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int e = 0;
    if (mx > 0.0f && mx < __int_as_float(0x7f800000)) {
        int ex;
        frexpf(mx, &ex);
        e = 12 - ex;
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
    }
    const float sq = ldexpf(1.0f, e);
    for (int p = lane; p < dimp; p += 32) a16[q * dimp + p] = __float2half_rn(p < dim ? __fmul_rn(qv[p], sq) : 0.0f);
    if (lane == 0) fq[q] = -ldexpf(1.0f, (is_dot ? 0 : 1) - e - sx_exp);
}

// Compaction of a streaming selection WITHOUT a sort: any superset of the k smallest keys may stay, so it is enough to
// find a pivot with at least k keys below it and drop the rest.  32 sampled keys are sorted across the warp with
// shuffles, the smallest sample with >= k keys below it becomes the pivot (binary search, one counting pass each), and
// the survivors are packed in place with ballots.  tau = pivot is a valid filter (the k-th smallest is below it).
// ~10x fewer shared-memory operations than the bitonic sort of the whole buffer; the one exact sort happens at the end.
__device__ __forceinline__ void select_compact_approx(const TopK &t, int slot, int lane) {
    unsigned long long *a = t.keys + (size_t)slot * t.C;
    int n = t.cnt[slot];
    if (n > t.C) n = t.C;
    __syncwarp();
    if (n <= t.k) return;
    unsigned long long s = a[(int)(((long long)lane * n) >> 5)];
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, s, stride);
            const bool take_min = ((lane & size) == 0) == ((lane & stride) == 0);
            s = take_min ? (s < o ? s : o) : (s > o ? s : o);
        }
    int lo = 0, hi = 31, best = -1, kept = 0;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const unsigned long long pv = __shfl_sync(0xffffffffu, s, mid);
        int c = 0;
        for (int i = lane; i < n; i += 32) c += a[i] < pv ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= t.k) {
            best = mid;
            kept = c;
            hi = mid - 1;
        } else {
            lo = mid + 1;
        }
    }
    if (best < 0 || kept > t.k + (t.C - t.k) / 2) {  // no usable pivot among the samples: exact compaction
        topk_compact_warp(t, slot, lane, false);
        return;
    }
    const unsigned long long pivot = __shfl_sync(0xffffffffu, s, best);
    __syncwarp();  // the counting passes above read the whole buffer: ordered before the in-place packing below
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const unsigned long long v = i < n ? a[i] : VG_KEY_EMPTY;
        const bool keep = i < n && v < pivot;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        __syncwarp();  // every lane has read its key of this chunk before any lane writes (memory ordering, not only convergence)
        if (keep) a[base + __popc(b & ((1u << lane) - 1u))] = v;  // target index <= i: inside chunks already read
        base += __popc(b);
    }
    __syncwarp();
    if (lane == 0) {
        t.cnt[slot] = base;
        t.tau[slot] = pivot;
    }
    __syncwarp();
}

// tau(q) = kc-th smallest group minimum (m1) of query q and the kc groups that reach it: one warp per query streams
// mins[q][*] (coalesced) through the shared-memory bounded top-k (threshold filter + bitonic compaction).  Output per
// selected group: the row its m1 names (gid * G + index bits) and a "crowded" flag (bit 31) when m2 <= tau as well.
// tau = +inf when there are fewer than kc groups (then every group is listed as crowded: the scan covers everything).
#define VG_TC_CROWDED 0x80000000u
constexpr int TC_LIST_CAP = 4096;   // candidate rows per query in the exact stage
__global__ void __launch_bounds__(256) tc_select_kernel(const float2 *mins, int64_t groups, int64_t nq, int kc, int C, uint32_t idx_mask,
                                                        int g_shift, float *tau, uint32_t *cand, int32_t *gcnt) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    TopK tk = topk_carve(smem, nw, C, kc);
    topk_init(tk, nw, threadIdx.x, blockDim.x);
    __syncthreads();
    const int64_t q = (int64_t)blockIdx.x * nw + warp;
    if (q >= nq) return;
    const float INF = __int_as_float(0x7f800000);
    const float2 *src = mins + q * groups;
    const int trigger = C - 32;
    // 8 independent coalesced loads per lane per step (the loop is otherwise bound by L2 latency).  The kernel is
    // instruction-bound (ncu: 60 % issue slots, 0.65 TB/s): once tau is tight almost nothing passes, so a float compare
    // against tau's score and one ballot per 32 groups skip the 64-bit key path for the whole warp.
    float tau_f = INF;  // score of the current threshold key: key < tau implies value <= tau_f
    for (int64_t g0 = 0; g0 < groups; g0 += 256) {
        float v[8];
        if (g0 + 256 <= groups) {
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = __ldcg(&src[g0 + u * 32 + lane].x);
        } else {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int64_t g = g0 + u * 32 + lane;
                v[u] = g < groups ? __ldcg(&src[g].x) : INF;
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int64_t g = g0 + u * 32 + lane;
            const bool maybe = v[u] <= tau_f && g < groups;
            if (__ballot_sync(0xffffffffu, maybe) == 0u) continue;
            if (maybe) {
                const unsigned long long key = ((unsigned long long)f32_orderable(v[u]) << 32) | (unsigned long long)(uint32_t)g;
                if (key < tk.tau[warp]) {
                    const int pos = atomicAdd(&tk.cnt[warp], 1);
                    if (pos < C) tk.keys[(size_t)warp * C + pos] = key;
                }
            }
            __syncwarp();
            if (tk.cnt[warp] > trigger) {
                select_compact_approx(tk, warp, lane);
                const unsigned long long tkey = tk.tau[warp];
                tau_f = tkey == VG_KEY_EMPTY ? INF : f32_from_orderable((uint32_t)(tkey >> 32));
            }
            __syncwarp();
        }
    }
    topk_compact_warp(tk, warp, lane, true);
    const int n = tk.cnt[warp];
    const unsigned long long *a = tk.keys + (size_t)warp * C;
    const float t = (n >= kc) ? f32_from_orderable((uint32_t)(a[kc - 1] >> 32)) : INF;
    for (int i = lane; i < kc; i += 32) {
        uint32_t out = 0xFFFFFFFFu;
        if (i < n) {
            const uint32_t g = (uint32_t)a[i];
            const float m1 = f32_from_orderable((uint32_t)(a[i] >> 32));
            const float m2 = __ldcg(&src[g].y);
            const uint32_t row = (g << g_shift) | (__float_as_uint(m1) & idx_mask);
            out = (m2 <= t) ? (VG_TC_CROWDED | g) : row;   // crowded: scan the whole group g
        }
        cand[q * kc + i] = out;
    }
    if (lane == 0) {
        gcnt[q] = n;
        tau[q] = t;
    }
}

// Same selection with a whole CTA per query (small batches: a warp per query leaves most of the machine idle and walks
// the groups at L2 latency): 256 threads offer their groups to one shared top-k, compaction between rounds.
__global__ void __launch_bounds__(256) tc_select_block_kernel(const float2 *mins, int64_t groups, int64_t nq, int kc, int C, uint32_t idx_mask,
                                                              int g_shift, float *tau, uint32_t *cand, int32_t *gcnt) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    TopK tk = topk_carve(smem, 1, C, kc);
    topk_init(tk, 1, tid, 256);
    __syncthreads();
    const float INF = __int_as_float(0x7f800000);
    const float2 *src = mins + q * groups;
    const int trigger = C - 512;  // two rounds of 256 offers fit above the trigger
    // 2048 groups per step: eight loads per thread in flight, then four rounds of (512 offers, maintenance).  One dependent
    // load round trip per 512 groups made the kernel latency-bound (1250 queries x 78k groups: 1.40 ms for 0.8 GB).
    for (int64_t g0 = 0; g0 < groups; g0 += 2048) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int64_t g = g0 + u * 256 + tid;
            v[u] = g < groups ? __ldcg(&src[g].x) : INF;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (g0 + r * 512 >= groups) break;   // uniform
            // float compare against the threshold's score first: after the first few thousand groups almost nothing passes, and
            // the 64-bit key is only built for what does (a score equal to the threshold's may still win on the group index)
            const unsigned long long tkey = tk.tau[0];
            const float tau_f = tkey == VG_KEY_EMPTY ? INF : f32_from_orderable((uint32_t)(tkey >> 32));
#pragma unroll
            for (int u = 2 * r; u < 2 * r + 2; u++) {
                const int64_t g = g0 + u * 256 + tid;
                if (g < groups && v[u] <= tau_f)
                    topk_offer(tk, 0, ((unsigned long long)f32_orderable(v[u]) << 32) | (unsigned long long)(uint32_t)g, trigger);
            }
            __syncthreads();
            topk_block_maintain_single(tk, tid, 256);
        }
    }
    __syncthreads();
    topk_block_maintain_single(tk, tid, 256, true);   // final sort, by the whole block
    {
        const int n = tk.cnt[0];
        const unsigned long long *a = tk.keys;
        const float t = (n >= kc) ? f32_from_orderable((uint32_t)(a[kc - 1] >> 32)) : INF;
        const int lane = tid;
        for (int i = tid; i < kc; i += 256) {
            uint32_t out = 0xFFFFFFFFu;
            if (i < n) {
                const uint32_t g = (uint32_t)a[i];
                const float m1 = f32_from_orderable((uint32_t)(a[i] >> 32));
                const float m2 = __ldcg(&src[g].y);
                const uint32_t row = (g << g_shift) | (__float_as_uint(m1) & idx_mask);
                out = (m2 <= t) ? (VG_TC_CROWDED | g) : row;
            }
            cand[q * kc + i] = out;
        }
        if (lane == 0) {
            gcnt[q] = n;
            tau[q] = t;
        }
    }
}

// Exact stage: one CTA per query.  The candidate rows (arg-min row of every selected group, all rows of crowded
// groups) are scored half-warp per row in simd.SquaredL2 / simd.Dot order (floats_avx512.c:12-129: 4 x 16-lane FMA
// accumulators, (A1+A2)+(A3+A4), lane tree, FMA scalar tail), bounded top-k under the heap order (score, row), then
// the certificate in double precision.
__global__ void __launch_bounds__(128) tc_exact_kernel(const float *vectors, int64_t dim, int64_t rows, const float *queries, int64_t q_stride,
                                                       const uint32_t *cand, const int32_t *gcnt, int kc, int G, const float *tau,
                                                       const float *qn, const unsigned int *xmax_bits, const uint8_t *mask, int k,
                                                       int C, int is_dot, int fp16, uint32_t row_base, uint32_t *out_rows, float *out_scores,
                                                       int32_t *out_counts, int32_t *fail_flags) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x, hw = tid >> 4, lane = tid & 15;
    float *qs = reinterpret_cast<float *>(smem);
    const size_t qbytes = ((size_t)dim * 4 + 15) & ~(size_t)15;
    TopK tk = topk_carve(smem + qbytes, 1, C, k);
    for (int64_t d = tid; d < dim; d += 128) qs[d] = queries[q * q_stride + d];
    topk_init(tk, 1, tid, 128);
    __syncthreads();
    const int ng = gcnt[q];
    const int64_t epochs = dim >> 6;
    const int trigger = C - 16;
    // expand the candidate entries (1 row, or G rows when crowded) into a flat row list in shared memory
    int32_t *rowlist = reinterpret_cast<int32_t *>(smem + qbytes + topk_smem_bytes(1, C));
    __shared__ int s_total;
    if (tid == 0) {
        int n = 0;
        for (int gi = 0; gi < ng; gi++) {
            const uint32_t c = cand[q * kc + gi];
            n += (c & VG_TC_CROWDED) ? G : 1;
        }
        s_total = n;
    }
    __syncthreads();
    const bool overflow = s_total > TC_LIST_CAP;   // pathological (most selected groups crowded): exact re-run
    if (!overflow) {
        if (tid < 32) {  // one warp writes the list: entries in order, crowded groups expanded by the lanes
            int off = 0;
            for (int gi = 0; gi < ng; gi++) {
                const uint32_t c = cand[q * kc + gi];
                if (c & VG_TC_CROWDED) {
                    const int64_t first = (int64_t)(c & ~VG_TC_CROWDED) * G;
                    for (int r = tid; r < G; r += 32) rowlist[off + r] = (first + r < rows) ? (int32_t)(first + r) : -1;
                    off += G;
                } else {
                    if (tid == 0) rowlist[off] = ((int64_t)c < rows) ? (int32_t)c : -1;
                    off += 1;
                }
            }
        }
        __syncthreads();
        const int total = s_total;
        for (int r0 = 0; r0 < total; r0 += 16) {  // 8 half-warps x 2 rows
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int r = r0 + hw * 2 + u;
                int64_t row = (r < total) ? (int64_t)rowlist[r] : -1;
                if (row >= 0 && mask && !((mask[row >> 3] >> (row & 7)) & 1)) row = -1;
                const bool valid = row >= 0;  // dead rows are computed (full-mask shuffles in reduce16) but not offered
                const float *x = vectors + (valid ? row : 0) * dim;
                float a[4] = {0.f, 0.f, 0.f, 0.f};
                for (int64_t e = 0; e < epochs; e++)
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) {
                        const int64_t d = e * 64 + jj * 16 + lane;
                        if (is_dot) {
                            a[jj] = __fmaf_rn(qs[d], __ldg(x + d), a[jj]);
                        } else {
                            const float df = __fsub_rn(qs[d], __ldg(x + d));
                            a[jj] = __fmaf_rn(df, df, a[jj]);
                        }
                    }
                float tot = reduce16(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])));
                if (lane == 0 && valid) {
                    for (int64_t d = epochs * 64; d < dim; d++) {
                        if (is_dot) {
                            tot = __fmaf_rn(qs[d], __ldg(x + d), tot);
                        } else {
                            const float df = __fsub_rn(qs[d], __ldg(x + d));
                            tot = __fmaf_rn(df, df, tot);
                        }
                    }
                    topk_offer(tk, 0, make_key(tot, row_base + (uint32_t)row, is_dot != 0), trigger);
                }
            }
            __syncthreads();
            topk_block_maintain(tk, 1, tid, 128);
        }
    }
    __syncthreads();
    if (tid < 32) {
        topk_emit_warp(tk, 0, tid, is_dot != 0, out_rows + q * k, out_scores + q * k, out_counts + q, k);
        __syncwarp();
        if (tid == 0) {
            const int m = tk.cnt[0];
            int fail = overflow ? 1 : 0;
            const float t = tau[q];
            if (!overflow && t < __int_as_float(0x7f800000)) {  // finite threshold: rows that were not scored exist
                if (m < k) {
                    fail = 1;
                } else {
                    const double qq = (double)qn[q], xx = (double)__uint_as_float(*xmax_bits);
                    // TF32 truncates each operand to 2^-10 relative; fp16 (pair kernel) rounds to 2^-11 and adds fp32 accumulation slack
                    const double c1 = (is_dot ? 1.0 / 512.0 : 1.0 / 256.0) * 1.125 * (fp16 ? 0.5 : 1.0);
                    const double c2 = 1.0 / 16384.0 + (fp16 ? (double)dim / 8388608.0 : 0.0);
                    const double smax = is_dot ? sqrt(qq * xx) : xx + 2.0 * sqrt(qq * xx);   // |s| of any row
                    const double E = c1 * sqrt(qq * xx) + c2 * (qq + xx) + smax * (double)G / 8388608.0;  // index bits: 2^-(23-log2 G)
                    const double ex = (double)out_scores[q * k + (k - 1)];
                    const double s_exact = is_dot ? -ex : ex - qq;
                    if (!(s_exact < (double)t - E)) fail = 1;
                }
            }
            fail_flags[q] = fail;
        }
    }
}

// ------------------------------------------------------------------ norms
// ||v||^2 per row, half-warp per row, 4x16-lane FMA accumulators (any accurate fp32
// order would do: the value only feeds the filter and its error bound).
__global__ void __launch_bounds__(256) sqnorm_kernel(const float *v, int64_t n, int64_t dim, int64_t stride, float *out,
                                                     unsigned int *max_bits) {
    const int64_t hw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const bool live = hw < n;
    const float *x = v + (live ? hw : n - 1) * stride;
    float a = 0.0f;
    for (int64_t d = lane; d < dim; d += 16) a = __fmaf_rn(x[d], x[d], a);
    a = reduce16(a);
    if (lane == 0 && live) {
        out[hw] = a;
        if (max_bits) atomicMax(max_bits, __float_as_uint(a));  // a >= 0: uint order = float order
    }
}

// ------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static vg_status get_encode() {
    if (g_encode) return VG_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(VG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return VG_OK;
}

static vg_status make_map(CUtensorMap *map, const float *base, int64_t rows, int64_t dim, int64_t stride, int box_rows) {
    VG_TRY(get_encode());
    const cuuint64_t dims[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)stride * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VG_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return VG_OK;
}

bool supported(int64_t dim, int64_t rows, int64_t nq, int64_t k) {
    return dim >= 16 && dim % 4 == 0 && rows >= 8192 && rows < (1ll << 31) && nq >= 16 && k >= 1 && k <= 32;
}
int candidates_for(int64_t k, int64_t dim) {
    const int kc = k <= 16 ? 32 : (int)(2 * k);
    return dim > 256 ? std::max(kc, 64) : kc;  // the error bound grows with ||q|| ||x||: keep more slack on long vectors
}
// Nearest-centroid assignment (k = 1 against a small table): a few groups are enough.
bool supported_assign(int64_t dim, int64_t rows, int64_t nq, int64_t q_stride) {
    return dim >= 16 && dim % 4 == 0 && q_stride % 4 == 0 && rows >= 128 && rows < (1ll << 31) && nq >= 1024;
}
int candidates_for_assign(int64_t rows) {
    const int64_t groups = (rows + 31) / 32;
    return (int)std::max<int64_t>(2, std::min<int64_t>(8, groups / 2));
}

vg_status sqnorms(const float *d_v, int64_t n, int64_t dim, int64_t stride, float *d_out, unsigned int *d_max_bits, cudaStream_t st) {
    if (n <= 0) return VG_OK;
    const int64_t threads = n * 16;
    sqnorm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_v, n, dim, stride, d_out, d_max_bits);
    VG_LAUNCHED();
    return VG_OK;
}

template <bool RESIDENT, bool IS_DOT>
static vg_status launch(const CUtensorMap &mq, const CUtensorMap &mx, const Args &a, int64_t qtiles, int splits, cudaStream_t st) {
    using S = Smem<RESIDENT>;
    const size_t sm = S::TOTAL + 1024;  // slack for the 1024-byte alignment of the dynamic segment
    VG_CUDA(cudaFuncSetAttribute(flat_tc_kernel<RESIDENT, IS_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid((unsigned)qtiles, (unsigned)splits);
    flat_tc_kernel<RESIDENT, IS_DOT><<<grid, NTHREADS, sm, st>>>(mq, mx, a);
    VG_LAUNCHED();
    return VG_OK;
}

vg_status select_groups(const float2 *d_mins, int64_t groups, int64_t nq, int kc, int64_t G, float *d_tau, uint32_t *d_cand,
                        int32_t *d_gcnt, cudaStream_t st) {
    int g_shift0 = 0;
    while ((1ll << g_shift0) < G) g_shift0++;
    if (nq <= 2048 && groups >= 65536) {  // few queries over very long group lists (measured: C4 33.5 -> 29.6 ms; short lists are faster per warp)
        const int Cb = topk_capacity(kc, 512);
        const size_t smb = topk_smem_bytes(1, Cb);
        if (smb <= 200 * 1024) {
            if (smb > 48 * 1024) VG_CUDA(cudaFuncSetAttribute(tc_select_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
            tc_select_block_kernel<<<(unsigned)nq, 256, smb, st>>>(d_mins, groups, nq, kc, Cb, (uint32_t)(G - 1), g_shift0, d_tau, d_cand, d_gcnt);
            VG_LAUNCHED();
            return VG_OK;
        }
    }
    const int C = topk_capacity(kc, 32);
    int nw = 8;
    while (nw > 1 && topk_smem_bytes(nw, C) > 96 * 1024) nw >>= 1;
    const size_t sm = topk_smem_bytes(nw, C);
    if (sm > 48 * 1024) VG_CUDA(cudaFuncSetAttribute(tc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int g_shift = 0;
    while ((1ll << g_shift) < G) g_shift++;
    tc_select_kernel<<<(unsigned)((nq + nw - 1) / nw), nw * 32, sm, st>>>(d_mins, groups, nq, kc, C, (uint32_t)(G - 1), g_shift, d_tau, d_cand,
                                                                         d_gcnt);
    VG_LAUNCHED();
    return VG_OK;
}

vg_status tensor_map_2d(void *map, bool f16, const void *base, int64_t rows, int64_t cols, int64_t stride_elems, int box_cols, int box_rows) {
    VG_TRY(get_encode());
    const int es = f16 ? 2 : 4;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)stride_elems * es};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(reinterpret_cast<CUtensorMap *>(map), f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VG_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return VG_OK;
}

// the same for byte matrices (kind::i8 operands): box_cols bytes x box_rows rows, 128-byte swizzle
vg_status tensor_map_2d_u8(void *map, const void *base, int64_t rows, int64_t cols, int64_t stride_bytes, int box_cols, int box_rows) {
    VG_TRY(get_encode());
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)stride_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(reinterpret_cast<CUtensorMap *>(map), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VG_ERR_CUDA, "cuTensorMapEncodeTiled (uint8) failed (code " + std::to_string((int)r) + ")");
    return VG_OK;
}

int64_t group_rows(int64_t rows, int kc) {
    // minimum groups of G rows: about 128*kc groups make tau tight (two of the best kc rows rarely share a group)
    // while the exact scan of kc*G rows per query stays ~1% of the GEMM's work
    static int per_kc = -1;   // groups per candidate; VECGO_TC_GROUPS_PER_KC overrides (tuning / measurement)
    if (per_kc < 0) {
        const char *e = getenv("VECGO_TC_GROUPS_PER_KC");
        const int v = e ? atoi(e) : 0;
        per_kc = v >= 4 && v <= 4096 ? v : 128;
    }
    // short segments (a shard of a multi-GPU database): half as many groups — the selection reads half the plane and the
    // exact stage barely grows (measured on a 1.25M-row SQ8 shard, 10k queries: 16.4 -> 15.6 ms)
    const int64_t want = (getenv("VECGO_TC_GROUPS_PER_KC") || rows > (4ll << 20)) ? per_kc : 64;
    int64_t G = 32;
    while (G < 1024 && rows / (G * 2) >= want * kc) G *= 2;  // <= 10 mantissa bits carry the row index
    return G;
}

bool pair_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("VECGO_FLAT_PAIR");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
bool uses_pair(const FilterArgs &f) {
    return f.d_x16 != nullptr && pair_enabled() && f.rows >= 8192 && (f.q_stride == 0 || f.q_stride == f.dim) &&
           (reinterpret_cast<uintptr_t>(f.d_x16) & 15) == 0;
}
int64_t filter_group_rows(const FilterArgs &f) {
    const int64_t G = group_rows(f.rows, f.kc);
    if (!uses_pair(f)) return G;
    return std::min<int64_t>(G, f.dim <= 256 ? 64 : 128);  // pair kernel: a group stays inside one thread's 128 (64) columns
}
vg_status make_shadow16(const float *d_x, int64_t rows, int64_t dim, int dimp, int sx_exp, void *d_x16, cudaStream_t st) {
    const int64_t total = rows * dimp;
    if (total <= 0) return VG_OK;
    shadow16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_x, rows, dim, dimp, sx_exp, reinterpret_cast<__half *>(d_x16));
    VG_LAUNCHED();
    return VG_OK;
}

template <bool IS_DOT, int EPW, bool MINONLY>
static vg_status launch_pair(const CUtensorMap &mq, const CUtensorMap &mx, const Args2 &a, int64_t qtiles, int splits, cudaStream_t st) {
    const size_t sm = pairf::SMEM2_BYTES;
    VG_CUDA(cudaFuncSetAttribute(flat2_kernel<IS_DOT, EPW, MINONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid((unsigned)(2 * qtiles), (unsigned)splits);
    flat2_kernel<IS_DOT, EPW, MINONLY><<<grid, (2 + EPW) * 32, sm, st>>>(mq, mx, a);
    VG_LAUNCHED();
    return VG_OK;
}
// sixteen epilogue warps when a tile has at most four k-blocks of MMA work (d <= 256)
static bool pair_wide_epilogue(const FilterArgs &f) { return f.dim <= 256; }

// The filter on the CTA-pair fp16 kernel (f.d_x16 set).
static vg_status filter_pair(const FilterArgs &f, cudaStream_t st) {
    const int dimp = (int)((f.dim + 63) / 64 * 64);
    const int64_t qtiles = (f.nq + BMQ - 1) / BMQ, nq_pad = qtiles * BMQ;
    const int64_t G = filter_group_rows(f);
    const int64_t groups = (f.rows + G - 1) / G;
    DevBuf a16, fq, mins;
    VG_TRY(a16.alloc((size_t)f.nq * dimp * 2));
    VG_TRY(fq.alloc((size_t)f.nq * 4));
    VG_TRY(mins.alloc((size_t)groups * nq_pad * 8));
    prep_queries16_kernel<<<(unsigned)((f.nq * 32 + 255) / 256), 256, 0, st>>>(f.d_queries, f.nq, f.q_stride ? f.q_stride : f.dim, f.dim, dimp, f.x16_exp,
                                                                              f.is_dot, a16.as<__half>(), fq.as<float>());
    VG_LAUNCHED();
    CUtensorMap mq, mx;
    VG_TRY(tensor_map_2d(&mq, true, a16.p, f.nq, dimp, dimp, pairf::BKH, BM));
    VG_TRY(tensor_map_2d(&mx, true, f.d_x16, f.rows, dimp, dimp, pairf::BKH, BN));
    const int64_t unit = pairf::TILE_ROWS;
    const int64_t sms = sm_count() / 2;
    const int64_t max_splits = std::max<int64_t>(1, f.rows / (4 * unit));
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
    int64_t rps = (f.rows + splits - 1) / splits;
    rps = (rps + unit - 1) / unit * unit;
    splits = (f.rows + rps - 1) / rps;
    Args2 a;
    a.xn = f.d_xn;
    a.mask = reinterpret_cast<const uint32_t *>(f.d_mask);
    a.fq = fq.as<float>();
    a.nq = f.nq;
    a.rows = f.rows;
    a.rows_per_split = rps;
    a.kb = dimp / pairf::BKH;
    a.cpg = (int)(G / 32);
    a.mins = mins.as<float2>();
    a.groups = groups;
    a.idx_mask = (uint32_t)(G - 1);
    a.keep_hi = ~31u;
    a.tile_list = nullptr;
    a.tile_count = nullptr;
    tiles::Lists tl;
    if (f.d_mask && tiles::enabled()) {
        VG_TRY(tiles::build(a.mask, f.rows, tl, st));
        VG_TRY(tiles::fill_skipped_groups(tl, (int)(pairf::TILE_ROWS / G), groups, f.nq, a.mins, st));
        a.tile_list = tl.list;
        a.tile_count = tl.count;
    }
    if (pair_wide_epilogue(f)) {
        // minimum-only epilogue (every selected group scored whole, <= kc * 64 rows per query) when the GEMM dominates:
        // on a 100k-row segment the larger exact stage costs more than the epilogue saves (measured 0.19 -> 0.28 ms)
        const bool minonly = f.rows >= (1ll << 20);
        if (f.is_dot) VG_TRY(minonly ? (launch_pair<true, 16, true>(mq, mx, a, qtiles, (int)splits, st)) : (launch_pair<true, 16, false>(mq, mx, a, qtiles, (int)splits, st)));
        else VG_TRY(minonly ? (launch_pair<false, 16, true>(mq, mx, a, qtiles, (int)splits, st)) : (launch_pair<false, 16, false>(mq, mx, a, qtiles, (int)splits, st)));
    } else {
        if (f.is_dot) VG_TRY((launch_pair<true, 8, false>(mq, mx, a, qtiles, (int)splits, st)));
        else VG_TRY((launch_pair<false, 8, false>(mq, mx, a, qtiles, (int)splits, st)));
    }
    VG_TRY(select_groups(a.mins, groups, f.nq, f.kc, G, f.d_tau, f.d_gids, f.d_gcnt, st));
    return VG_OK;  // a16 / fq / mins are returned to the stream-ordered pool (freed in stream order); tensor maps were copied at launch
}

vg_status filter(const FilterArgs &f, cudaStream_t st) {
    if (f.dim < 16 || f.dim % 4 != 0 || f.rows < 128 || f.rows >= (1ll << 31) || f.kc < 1 || f.kc > 64 || (f.rows + 31) / 32 < f.kc)
        return fail(VG_ERR_UNSUPPORTED, "shape not supported by the tensor-core filter");
    if (uses_pair(f)) return filter_pair(f, st);
    CUtensorMap mq, mx;
    VG_TRY(make_map(&mq, f.d_queries, f.nq, f.dim, f.q_stride ? f.q_stride : f.dim, BMQ));
    VG_TRY(make_map(&mx, f.d_vectors, f.rows, f.dim, f.dim, BN));
    const int64_t qtiles = (f.nq + BMQ - 1) / BMQ;
    const int64_t nq_pad = qtiles * BMQ;
    const int64_t G = group_rows(f.rows, f.kc);
    const int64_t groups = (f.rows + G - 1) / G;
    const int64_t unit = std::max<int64_t>(BN, G);  // row ranges are whole tiles and whole groups
    // one CTA per SM (shared memory): pick the row-split count whose CTA total fills whole waves best
    const int64_t sms = sm_count();
    const int64_t max_splits = std::max<int64_t>(1, f.rows / (4 * unit));
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
    int64_t rps = (f.rows + splits - 1) / splits;
    rps = (rps + unit - 1) / unit * unit;
    splits = (f.rows + rps - 1) / rps;
    Args a;
    a.xn = f.d_xn;
    a.mask = reinterpret_cast<const uint32_t *>(f.d_mask);
    a.nq = f.nq;
    a.nq_pad = nq_pad;
    a.rows = f.rows;
    a.rows_per_split = rps;
    a.kb = (int)((f.dim + BK - 1) / BK);
    a.cpg = (int)(G / 32);
    a.row_base = f.row_base;
    a.groups = groups;
    DevBuf mins;
    VG_TRY(mins.alloc((size_t)groups * nq_pad * 8));
    a.mins = mins.as<float2>();
    a.idx_mask = (uint32_t)(G - 1);
    const bool resident = a.kb <= MAX_RES_KB;
    if (resident) {
        if (f.is_dot) VG_TRY((launch<true, true>(mq, mx, a, qtiles, (int)splits, st)));
        else VG_TRY((launch<true, false>(mq, mx, a, qtiles, (int)splits, st)));
    } else {
        if (f.is_dot) VG_TRY((launch<false, true>(mq, mx, a, qtiles, (int)splits, st)));
        else VG_TRY((launch<false, false>(mq, mx, a, qtiles, (int)splits, st)));
    }
    VG_TRY(select_groups(a.mins, groups, f.nq, f.kc, G, f.d_tau, f.d_gids, f.d_gcnt, st));
    return VG_OK;  // mins is returned to the stream-ordered pool (freed in stream order)
}

vg_status finalize(const FilterArgs &f, int k, const float *d_qn, const unsigned int *d_xmax_bits, uint32_t *d_rows, float *d_scores,
                   int32_t *d_counts, int32_t *d_fail, cudaStream_t st) {
    const int C = topk_capacity(k, 16);
    const size_t sm = (((size_t)f.dim * 4 + 15) & ~(size_t)15) + topk_smem_bytes(1, C) + (size_t)TC_LIST_CAP * 4;
    if (sm > 200 * 1024) return fail(VG_ERR_UNSUPPORTED, "dimension too large for the exact stage");
    VG_CUDA(cudaFuncSetAttribute(tc_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    tc_exact_kernel<<<(unsigned)f.nq, 128, sm, st>>>(f.d_vectors, f.dim, f.rows, f.d_queries, f.q_stride ? f.q_stride : f.dim, f.d_gids, f.d_gcnt, f.kc,
                                                    (int)filter_group_rows(f), f.d_tau, d_qn, d_xmax_bits, f.d_mask, k, C, f.is_dot,
                                                    uses_pair(f) ? 1 : 0, f.row_base, d_rows, d_scores, d_counts, d_fail);
    VG_LAUNCHED();
    return VG_OK;
}

// ------------------------------------------------------------------ switch + statistics
static std::atomic<int> g_enabled{-1};
static std::atomic<uint64_t> g_queries{0}, g_fallbacks{0};
bool enabled() {
    int v = g_enabled.load();
    if (v < 0) {
        const char *e = getenv("VECGO_FLAT_TC");
        v = (e && e[0] == '0') ? 0 : 1;
        g_enabled.store(v);
    }
    return v != 0;
}
void set_enabled(bool on) { g_enabled.store(on ? 1 : 0); }
void stats(uint64_t *queries, uint64_t *fallbacks) {
    if (queries) *queries = g_queries.load();
    if (fallbacks) *fallbacks = g_fallbacks.load();
}

// Filter + exact stage + certificate for one batch (norms of the queries included): enqueue only.  d_fail[q] = 1
// where the certificate did not hold.  Nothing here waits for the device.
static vg_status enqueue_once(const SearchIO &io, int kc, int32_t *d_fail, cudaStream_t st) {
    FilterArgs f;
    f.d_queries = io.d_queries;
    f.q_stride = io.q_stride;
    f.d_vectors = io.d_vectors;
    f.d_xn = io.d_xn;
    f.d_x16 = io.d_x16;
    f.x16_exp = io.x16_exp;
    f.d_mask = io.d_mask;
    f.nq = io.nq;
    f.rows = io.rows;
    f.dim = io.dim;
    f.kc = kc;
    f.is_dot = io.is_dot;
    f.row_base = io.row_base;
    // the query tiles of one launch share a [queries][groups] minima buffer: keep it under 8 GiB by chunking the batch
    // (whole query tiles), as the quantized filter does
    const int64_t G = filter_group_rows(f);
    const int64_t groups = (io.rows + G - 1) / G;
    const int64_t chunk = std::max<int64_t>(BMQ, ((8ll << 30) / (groups * 8)) / BMQ * BMQ);
    const int64_t q_stride = io.q_stride ? io.q_stride : io.dim;
    for (int64_t q0 = 0; q0 < io.nq; q0 += chunk) {
        const int64_t nq = std::min(chunk, io.nq - q0);
        DevBuf gids, gcnt, tau, qn;
        VG_TRY(gids.alloc((size_t)nq * kc * 4));
        VG_TRY(gcnt.alloc((size_t)nq * 4));
        VG_TRY(tau.alloc((size_t)nq * 4));
        VG_TRY(qn.alloc((size_t)nq * 4));
        f.d_queries = io.d_queries + q0 * q_stride;
        f.nq = nq;
        f.d_gids = gids.as<uint32_t>();
        f.d_gcnt = gcnt.as<int32_t>();
        f.d_tau = tau.as<float>();
        VG_TRY(sqnorms(f.d_queries, nq, io.dim, q_stride, qn.as<float>(), nullptr, st));
        VG_TRY(filter(f, st));
        VG_TRY(finalize(f, io.k, qn.as<float>(), io.d_xmax_bits, io.d_rows + q0 * io.k, io.d_scores + q0 * io.k, io.d_counts + q0,
                        d_fail + q0, st));
    }
    return VG_OK;  // the temporaries go back to the stream-ordered pool (freed in stream order)
}

void count_queries(uint64_t n) { g_queries.fetch_add(n); }

vg_status enqueue(const SearchIO &io, int kc, int32_t *d_fail, cudaStream_t st) {
    if (io.d_x16 && (io.q_stride == 0 || io.q_stride == io.dim) && fs::single_supported(io.dim, io.rows, io.nq, io.k)) {
        VG_TRY(fs::single_enqueue(io, d_fail, st));
        g_queries.fetch_add((uint64_t)io.nq);
        return VG_OK;
    }
    VG_TRY(enqueue_once(io, kc, d_fail, st));
    g_queries.fetch_add((uint64_t)io.nq);
    return VG_OK;
}

// Queries whose certificate failed get a second chance with twice the number of groups (a wider gap between the k-th
// best and tau); the ones that fail again are returned in `failed` for the exact re-run.
vg_status retry(const SearchIO &io, int kc, std::vector<int32_t> &failed, cudaStream_t st) {
    const int64_t groups = (io.rows + group_rows(io.rows, kc) - 1) / group_rows(io.rows, kc);
    const int kc2 = (int)std::min<int64_t>(64, std::min<int64_t>(2 * (int64_t)kc, groups / 2));
    if (!failed.empty() && kc2 > kc) {
        const int64_t nb = (int64_t)failed.size();
        DevBuf bidx, bq, brow, bsc, bcnt, bfail;
        VG_TRY(bidx.alloc((size_t)nb * 4));
        VG_TRY(bq.alloc((size_t)nb * io.dim * 4));
        VG_TRY(brow.alloc((size_t)nb * io.k * 4));
        VG_TRY(bsc.alloc((size_t)nb * io.k * 4));
        VG_TRY(bcnt.alloc((size_t)nb * 4));
        VG_TRY(bfail.alloc((size_t)nb * 4));
        VG_CUDA(cudaMemcpyAsync(bidx.p, failed.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, st));
        VG_TRY(dev_gather_rows(io.d_queries, io.q_stride ? io.q_stride : io.dim, bidx.as<int32_t>(), nb, io.dim, bq.as<float>(), st));
        SearchIO io2 = io;
        io2.d_queries = bq.as<float>();
        io2.q_stride = 0;
        io2.nq = nb;
        io2.d_rows = brow.as<uint32_t>();
        io2.d_scores = bsc.as<float>();
        io2.d_counts = bcnt.as<int32_t>();
        VG_TRY(enqueue_once(io2, kc2, bfail.as<int32_t>(), st));
        // results of the retried queries (also of those that failed again: the caller overwrites them)
        VG_TRY(dev_scatter_results(brow.as<uint32_t>(), bsc.as<float>(), bcnt.as<int32_t>(), bidx.as<int32_t>(), nb, io.k, io.d_rows,
                                   io.d_scores, io.d_counts, st));
        std::vector<int32_t> h_fail((size_t)nb);
        VG_CUDA(cudaMemcpyAsync(h_fail.data(), bfail.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
        VG_CUDA(cudaStreamSynchronize(st));
        std::vector<int32_t> still;
        for (int64_t j = 0; j < nb; j++)
            if (h_fail[(size_t)j]) still.push_back(failed[(size_t)j]);
        failed.swap(still);
    }
    g_fallbacks.fetch_add((uint64_t)failed.size());
    return VG_OK;
}

// One batch end to end, host-synchronous: enqueue, read the certificate flags back, second chance.
vg_status search(const SearchIO &io, int kc, std::vector<int32_t> &failed, cudaStream_t st) {
    failed.clear();
    DevBuf failb;
    VG_TRY(failb.alloc((size_t)io.nq * 4));
    VG_TRY(enqueue(io, kc, failb.as<int32_t>(), st));
    std::vector<int32_t> h_fail((size_t)io.nq);
    VG_CUDA(cudaMemcpyAsync(h_fail.data(), failb.p, (size_t)io.nq * 4, cudaMemcpyDeviceToHost, st));
    VG_CUDA(cudaStreamSynchronize(st));
    for (int64_t q = 0; q < io.nq; q++)
        if (h_fail[(size_t)q]) failed.push_back((int32_t)q);
    return retry(io, kc, failed, st);
}

}  // namespace tc
}  // namespace vg
