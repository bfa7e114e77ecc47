// vg_flat_tc.cu — exact Flat L2 / dot search as a tcgen05 (TF32) GEMM *filter*
// fed by TMA, with a register/shared-memory top-k' in the epilogue, followed by
// an exact re-check in the reference's own summation order.
//
// Why a filter.  flat.(*Segment).Search (internal/segment/flat/segment.go:690-697)
// calls simd.SquaredL2 / simd.Dot per (query,row); BASELINE's north_star wants that
// dense Q x N x d contraction on the 5th-gen tensor cores, AND top-k ids identical
// to the SIMD path.  Tensor cores give q.x only to TF32 accuracy, so they are used
// to shrink N rows to k' >= k candidates per query; the survivors are then scored
// by the exact AVX-512-order kernel (same code path as Segment.Rerank) and a
// certificate proves nothing outside the candidate set could have entered the
// top-k.  Queries whose certificate fails are re-run on the exact CUDA-core scan.
//
//   s(q,x) = ||x||^2 - 2 q.x   (L2; the per-query constant ||q||^2 is dropped)
//   s(q,x) = -q.x              (dot / cosine, descending in the reference)
//
// Certificate (per query).  Let T = the k'-th smallest approximate s, E >= |s_exact
// - s_approx| for every row (bound below), e_k = exact k-th best.  Every row outside
// the candidate set has s_approx >= T, hence s_exact >= T - E.  If s_exact(e_k) <
// T - E no outside row can tie or beat the k-th candidate: the exact top-k of the
// candidate set IS the exact top-k of the segment (ties by row id included).
//
// Error bound.  kind::tf32 keeps 10 explicit mantissa bits of each fp32 operand:
// |fl_tf32(a) - a| <= 2^-10 |a|, so |q.x - (q.x)_tc| <= (2^-9 + 2^-20) sum|q_i x_i|
// <= 2^-9 (1 + 2^-11) ||q|| ||x||, plus fp32 accumulation (<= d 2^-23 ||q|| ||x||).
// With the factor 2 of the L2 form:  E = c1 ||q|| max||x|| + c2 (||q||^2 + max||x||^2),
// c1 = 2^-8 * 1.125 (L2) or 2^-9 * 1.125 (dot), c2 = 2^-14 (norm rounding, accumulation
// slack).  tests/test_gpu_flat_tc.py measures the realised error against E.
//
// Kernel (one CTA = 128 queries x a contiguous row range, 192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor.2d of the 128 x 32-float query
//              k-block (A) and the BN x 32-float row k-block (B), 128B swizzle,
//              STAGES-deep mbarrier ring
//   warp 1     MMA issuer: tcgen05.mma.cta_group::1.kind::tf32, M=128, N=BN, K=8,
//              fp32 accumulators in TMEM, double-buffered (2 x BN columns)
//   warps 2-5  epilogue: tcgen05.ld 32 columns at a time; thread = one query
//              (TMEM lane), threshold in a register, candidate buffer private to
//              the thread in shared memory (no atomics); warp-cooperative bitonic
//              compaction when a buffer fills
#include <cuda.h>

#include "vg_flat_tc.cuh"
#include "vg_topk.cuh"

namespace vg {
namespace tc {

constexpr int BM = 128;   // queries per tile (UMMA M, = TMEM lanes)
constexpr int BK = 32;    // floats per k-block: one 128-byte swizzle atom
constexpr int NTHREADS = 192;

// ------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory written by TMA with SWIZZLE_128B: row r
// at r*128 bytes, 8-row groups 1024 bytes apart (SBO); LBO unused for swizzled
// K-major; descriptor version 1 (sm_100); layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=TF32 (2 at bits 7-9 and
// 10-12), both K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------ kernel
struct Args {
    const float *xn;        // [rows] ||x||^2 (L2) or nullptr (dot)
    const uint8_t *mask;    // optional row bitmap
    int64_t nq, rows, rows_per_split;
    int kb;                 // k-blocks = ceil(dim / 32)
    int kc;                 // candidates kept per query (k')
    int is_dot;
    uint32_t row_base;
    unsigned long long *partial;  // [nq][splits][kc] ascending keys, VG_KEY_EMPTY padded
};

template <int BN, int STAGES, int C>
struct Smem {
    static constexpr int A_BYTES = BM * BK * 4;        // 16 KB
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int KEY_STRIDE = C + 1;            // 8-byte words per query slot (+1: pushes of a warp spread over banks)
    static constexpr size_t OFF_KEYS = (size_t)STAGES * STAGE_BYTES;
    static constexpr size_t OFF_XN = OFF_KEYS + (size_t)BM * KEY_STRIDE * 8;
    static constexpr size_t OFF_BAR = OFF_XN + (size_t)2 * BN * 4;
    static constexpr size_t TOTAL = OFF_BAR + (size_t)(2 * STAGES + 4) * 8 + 16;
};

// Warp-cooperative: sort `n` keys of one slot ascending (bitonic over the power-of-two
// prefix), keep the best kc.  Returns the new count; *tau_out = kc-th key or EMPTY.
__device__ __forceinline__ int compact_slot(unsigned long long *a, int n, int kc, int cap, int lane, unsigned long long *tau_out) {
    int len = 32;
    while (len < n) len <<= 1;
    if (len > cap) len = cap;
    for (int i = n + lane; i < len; i += 32) a[i] = VG_KEY_EMPTY;
    __syncwarp();
    for (int size = 2; size <= len; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = lane; i < (len >> 1); i += 32) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long x = a[lo], y = a[hi];
                if ((x > y) == up) {
                    a[lo] = y;
                    a[hi] = x;
                }
            }
            __syncwarp();
        }
    }
    const int m = n < kc ? n : kc;
    *tau_out = (n >= kc) ? a[kc - 1] : VG_KEY_EMPTY;
    return m;
}

template <int BN, int STAGES, int C>
__global__ void __launch_bounds__(NTHREADS, 1)
flat_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, Args A) {
    using S = Smem<BN, STAGES, C>;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    // SWIZZLE_128B tiles need 1024-byte aligned bases; the dynamic segment only guarantees 16
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * BM;
    const int split = blockIdx.y, splits = gridDim.y;
    const int64_t row_begin = (int64_t)split * A.rows_per_split;
    int64_t row_end = row_begin + A.rows_per_split;
    if (row_end > A.rows) row_end = A.rows;
    const int ntiles = row_end > row_begin ? (int)((row_end - row_begin + BN - 1) / BN) : 0;

    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar0 = s_base + (uint32_t)S::OFF_BAR;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
    constexpr uint32_t TMEM_COLS = 2 * BN;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 128);
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

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = 0; t < ntiles; t++) {
                const int n0 = (int)(row_begin + (int64_t)t * BN);
                for (int kb = 0; kb < A.kb; kb++, it++) {
                    const int st = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(empty_bar(st), ph ^ 1);
                    mbar_expect_tx(full_bar(st), S::STAGE_BYTES);
                    const uint32_t sa = s_base + st * S::STAGE_BYTES;
                    tma_load_2d(sa, &map_q, kb * BK, q0, full_bar(st));
                    tma_load_2d(sa + S::A_BYTES, &map_x, kb * BK, n0, full_bar(st));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            uint32_t it = 0;
            for (int t = 0; t < ntiles; t++) {
                const int as = t & 1;
                const uint32_t aph = (t >> 1) & 1;
                mbar_wait(tempty_bar(as), aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < A.kb; kb++, it++) {
                    const int st = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(full_bar(st), ph);
                    tc_fence_after();
                    const uint32_t sa = s_base + st * S::STAGE_BYTES;
                    const uint64_t adesc = make_sdesc(sa), bdesc = make_sdesc(sa + S::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; k++)  // 8 tf32 = 32 bytes per UMMA: advance the start address inside the swizzle atom
                        umma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty_bar(st));
                }
                umma_commit(tfull_bar(as));
            }
        }
    } else {
        // ===================== epilogue: warps 2..5 =====================
        const int quad = warp & 3;            // TMEM lanes 32*quad .. 32*quad+31 are the ones this warp may read
        const int slot = quad * 32 + lane;    // query within the tile = TMEM lane
        const int et = (warp - 2) * 32 + lane;
        unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem + S::OFF_KEYS) + (size_t)slot * S::KEY_STRIDE;
        float *xs = reinterpret_cast<float *>(smem + S::OFF_XN);
        const bool q_live = (int64_t)(q0 + slot) < A.nq;
        float tau_f = q_live ? __int_as_float(0x7f800000) : -__int_as_float(0x7f800000);  // dead query rows accept nothing
        int cnt = 0;
        const float INF = __int_as_float(0x7f800000);
        for (int t = 0; t < ntiles; t++) {
            const int as = t & 1;
            const uint32_t aph = (t >> 1) & 1;
            const int64_t n0 = row_begin + (int64_t)t * BN;
            float *xt = xs + as * BN;
            for (int i = et; i < BN; i += 128) {
                const int64_t row = n0 + i;
                xt[i] = (row < row_end) ? (A.is_dot ? 0.0f : __ldg(A.xn + row)) : INF;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)(c * 32), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const float dot = __uint_as_float(v[j]);
                    const float xv = xt[c * 32 + j];
                    const float s = A.is_dot ? __fsub_rn(xv, dot) : __fmaf_rn(-2.0f, dot, xv);  // dot: xv = 0 (live) or +inf (padding)
                    if (s <= tau_f) {
                        const int64_t row = n0 + c * 32 + j;
                        bool ok = row < row_end;
                        if (ok && A.mask) ok = (A.mask[row >> 3] >> (row & 7)) & 1;
                        if (ok && cnt < C) {
                            keys[cnt] = ((unsigned long long)f32_orderable(s) << 32) | (unsigned long long)(A.row_base + (uint32_t)row);
                            cnt++;
                        }
                    }
                }
                // a buffer that could overflow during the next 32 columns is compacted now (warp-cooperative)
                unsigned need = __ballot_sync(0xffffffffu, cnt > C - 32);
                if (need) __syncwarp();  // candidate stores of the owning lanes become visible to the warp
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const int n_src = __shfl_sync(0xffffffffu, cnt, src);
                    unsigned long long tau_key;
                    unsigned long long *a = reinterpret_cast<unsigned long long *>(smem + S::OFF_KEYS) + (size_t)(quad * 32 + src) * S::KEY_STRIDE;
                    const int m = compact_slot(a, n_src, A.kc, C, lane, &tau_key);
                    if (lane == src) {
                        cnt = m;
                        tau_f = (tau_key == VG_KEY_EMPTY) ? INF : f32_from_orderable((uint32_t)(tau_key >> 32));
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(as));
        }
        // final: every slot sorted, best kc emitted
        __syncwarp();
        for (int src = 0; src < 32; src++) {
            const int n_src = __shfl_sync(0xffffffffu, cnt, src);
            unsigned long long tau_key;
            unsigned long long *a = reinterpret_cast<unsigned long long *>(smem + S::OFF_KEYS) + (size_t)(quad * 32 + src) * S::KEY_STRIDE;
            const int m = compact_slot(a, n_src, A.kc, C, lane, &tau_key);
            const int64_t q = (int64_t)q0 + quad * 32 + src;
            if (q < A.nq) {
                unsigned long long *out = A.partial + ((size_t)q * splits + split) * A.kc;
                for (int i = lane; i < A.kc; i += 32) out[i] = (i < m) ? a[i] : VG_KEY_EMPTY;
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ norms
// ||v||^2 per row, half-warp per row, 4x16-lane FMA accumulators (any accurate fp32
// order would do: the value only feeds the filter and its error bound).
__global__ void __launch_bounds__(256) sqnorm_kernel(const float *v, int64_t n, int64_t dim, float *out, unsigned int *max_bits) {
    const int64_t hw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane = threadIdx.x & 15;
    const bool live = hw < n;
    const float *x = v + (live ? hw : n - 1) * dim;
    float a = 0.0f;
    for (int64_t d = lane; d < dim; d += 16) a = __fmaf_rn(x[d], x[d], a);
    a = reduce16(a);
    if (lane == 0 && live) {
        out[hw] = a;
        if (max_bits) atomicMax(max_bits, __float_as_uint(a));  // a >= 0: uint order = float order
    }
}

// ------------------------------------------------------------------ finalize
// One CTA per query: exact scores of the k' candidates in simd.SquaredL2 / simd.Dot
// order (floats_avx512.c:12-129: 4 x 16-lane FMA accumulators, (A1+A2)+(A3+A4), lane
// tree, FMA scalar tail), then the heap order (score, row) picks the top k and the
// certificate is evaluated in double precision.
__global__ void __launch_bounds__(128) flat_tc_finalize_kernel(const float *vectors, int64_t dim, const float *queries, int64_t nq,
                                                               const uint32_t *cand_rows, const float *cand_s, const int32_t *cand_cnt,
                                                               int kc, int k, int is_dot, uint32_t row_base, const float *qn,
                                                               const unsigned int *xmax_bits, uint32_t *out_rows, float *out_scores,
                                                               int32_t *out_counts, int32_t *fail_flags) {
    __shared__ unsigned long long ek[128];
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x, hw = tid >> 4, lane = tid & 15;
    const int n = cand_cnt[q];
    const float *qv = queries + q * dim;
    for (int j0 = 0; j0 < 128; j0 += 8) {
        const int j = j0 + hw;
        if (j0 >= kc) break;
        const bool live = j < n;
        const uint32_t row = live ? cand_rows[q * kc + j] : row_base;
        const float *x = vectors + (int64_t)(row - row_base) * dim;
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        const int64_t epochs = dim >> 6;
        for (int64_t e = 0; e < epochs; e++)
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int64_t d = e * 64 + jj * 16 + lane;
                if (is_dot) {
                    a[jj] = __fmaf_rn(qv[d], __ldg(x + d), a[jj]);
                } else {
                    const float df = __fsub_rn(qv[d], __ldg(x + d));
                    a[jj] = __fmaf_rn(df, df, a[jj]);
                }
            }
        float tot = reduce16(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])));
        if (lane == 0 && j < 128) {
            if (live) {
                for (int64_t d = epochs * 64; d < dim; d++) {
                    if (is_dot) {
                        tot = __fmaf_rn(qv[d], __ldg(x + d), tot);
                    } else {
                        const float df = __fsub_rn(qv[d], __ldg(x + d));
                        tot = __fmaf_rn(df, df, tot);
                    }
                }
                ek[j] = make_key(tot, row, is_dot != 0);
            } else {
                ek[j] = VG_KEY_EMPTY;
            }
        }
    }
    __syncthreads();
    if (tid < 32) {
        unsigned long long tau;
        // sort all 128 slots (EMPTY beyond kc were written above only up to the 8-aligned bound; fill the rest)
        const int filled = ((kc + 7) / 8) * 8;
        for (int i = filled + tid; i < 128; i += 32) ek[i] = VG_KEY_EMPTY;
        __syncwarp();
        compact_slot(ek, 128, 128, 128, tid, &tau);
        const int m = n < k ? n : k;
        for (int i = tid; i < k; i += 32) {
            if (i < m) {
                out_rows[q * k + i] = key_row(ek[i]);
                out_scores[q * k + i] = key_score(ek[i], is_dot != 0);
            } else {
                out_rows[q * k + i] = 0xFFFFFFFFu;
                out_scores[q * k + i] = __uint_as_float(0x7fc00000u);
            }
        }
        if (tid == 0) {
            out_counts[q] = m;
            int fail = 0;
            if (n >= kc && m > 0) {  // candidate list was full: rows outside it exist (or may exist)
                const double T = (double)cand_s[q * kc + kc - 1];
                const double qq = (double)qn[q], xx = (double)__uint_as_float(*xmax_bits);
                const double c1 = (is_dot ? 1.0 / 512.0 : 1.0 / 256.0) * 1.125, c2 = 1.0 / 16384.0;
                const double E = c1 * sqrt(qq * xx) + c2 * (qq + xx);
                const double ex = (double)key_score(ek[m - 1], is_dot != 0);
                const double s_exact = is_dot ? -ex : ex - qq;
                if (!(s_exact < T - E)) fail = 1;
                if (m < k) fail = 1;  // fewer than k candidates although the list was full: cannot happen (kc >= k)
            }
            fail_flags[q] = fail;
        }
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

static vg_status make_map(CUtensorMap *map, const float *base, int64_t rows, int64_t dim, int box_rows) {
    VG_TRY(get_encode());
    const cuuint64_t dims[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)dim * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VG_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return VG_OK;
}

bool supported(int64_t dim, int64_t rows, int64_t nq, int64_t k) {
    return dim >= 16 && dim % 4 == 0 && rows >= 1024 && rows < (1ll << 31) && nq >= 16 && k >= 1 && k <= 64;
}
int candidates_for(int64_t k) { return k <= 10 ? 32 : (int)(k + 32); }

vg_status sqnorms(const float *d_v, int64_t n, int64_t dim, float *d_out, unsigned int *d_max_bits, cudaStream_t st) {
    if (n <= 0) return VG_OK;
    const int64_t threads = n * 16;
    sqnorm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_v, n, dim, d_out, d_max_bits);
    VG_LAUNCHED();
    return VG_OK;
}

template <int BN, int STAGES, int C>
static vg_status launch(const CUtensorMap &mq, const CUtensorMap &mx, const Args &a, int64_t qtiles, int splits, cudaStream_t st) {
    using S = Smem<BN, STAGES, C>;
    const size_t sm = S::TOTAL + 1024;  // slack for the 1024-byte alignment of the dynamic segment
    VG_CUDA(cudaFuncSetAttribute(flat_tc_kernel<BN, STAGES, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid((unsigned)qtiles, (unsigned)splits);
    flat_tc_kernel<BN, STAGES, C><<<grid, NTHREADS, sm, st>>>(mq, mx, a);
    VG_LAUNCHED();
    return VG_OK;
}

vg_status filter(const FilterArgs &f, cudaStream_t st) {
    if (!supported(f.dim, f.rows, f.nq, f.kc > 32 ? f.kc - 32 : 1)) return fail(VG_ERR_UNSUPPORTED, "shape not supported by the tensor-core filter");
    CUtensorMap mq, mx;
    const bool wide = f.kc <= 32;  // BN=256, C=64  |  BN=128, C=128
    VG_TRY(make_map(&mq, f.d_queries, f.nq, f.dim, BM));
    VG_TRY(make_map(&mx, f.d_vectors, f.rows, f.dim, wide ? 256 : 128));
    const int bn = wide ? 256 : 128;
    const int64_t qtiles = (f.nq + BM - 1) / BM;
    // one CTA per SM (shared memory): pick the row-split count whose CTA total fills whole waves best
    const int64_t sms = sm_count();
    int64_t splits = 1;
    double best = 0.0;
    for (int64_t s_ = 1; s_ <= sms; s_++) {
        const int64_t ctas = qtiles * s_, waves = (ctas + sms - 1) / sms;
        const double eff = (double)ctas / (double)(waves * sms);
        if (eff > best + 0.02) {
            best = eff;
            splits = s_;
        }
        if (eff >= 0.97) break;
    }
    const int64_t max_splits = (f.rows + 4 * bn - 1) / (4 * bn);
    if (splits > max_splits) splits = max_splits;
    int64_t rps = (f.rows + splits - 1) / splits;
    rps = (rps + bn - 1) / bn * bn;
    splits = (f.rows + rps - 1) / rps;
    Args a;
    a.xn = f.d_xn;
    a.mask = f.d_mask;
    a.nq = f.nq;
    a.rows = f.rows;
    a.rows_per_split = rps;
    a.kb = (int)((f.dim + BK - 1) / BK);
    a.kc = f.kc;
    a.is_dot = f.is_dot;
    a.row_base = f.row_base;
    DevBuf partial;
    VG_TRY(partial.alloc((size_t)f.nq * splits * f.kc * 8));
    a.partial = partial.as<unsigned long long>();
    if (wide) VG_TRY((launch<256, 3, 64>(mq, mx, a, qtiles, (int)splits, st)));
    else VG_TRY((launch<128, 2, 128>(mq, mx, a, qtiles, (int)splits, st)));
    // merge the per-split lists: keys are already in s-space (ascending)
    VG_TRY(launch_merge_keys(a.partial, splits, f.nq, f.kc, f.kc, splits * f.kc, false, f.kc, f.d_cand_rows, f.d_cand_s, f.d_cand_cnt, st));
    VG_CUDA(cudaStreamSynchronize(st));  // partial is freed on return
    return VG_OK;
}

vg_status finalize(const FilterArgs &f, int k, const float *d_qn, const unsigned int *d_xmax_bits, uint32_t *d_rows, float *d_scores,
                   int32_t *d_counts, int32_t *d_fail, cudaStream_t st) {
    flat_tc_finalize_kernel<<<(unsigned)f.nq, 128, 0, st>>>(f.d_vectors, f.dim, f.d_queries, f.nq, f.d_cand_rows, f.d_cand_s, f.d_cand_cnt,
                                                            f.kc, k, f.is_dot, f.row_base, d_qn, d_xmax_bits, d_rows, d_scores, d_counts,
                                                            d_fail);
    VG_LAUNCHED();
    return VG_OK;
}

}  // namespace tc
}  // namespace vg
