// vg_flat_single.cu — flat.(*Segment).Search (internal/segment/flat/segment.go:447-752, distance at :690-697) for short
// vectors (dim <= 256) as ONE kernel launch with no host round trip: the BASELINE configs[0] shape (100k x 128 float32,
// 1000 queries, k = 10) spent two thirds of its 0.18 ms between six launches and a host read of the certificate flags.
//
// Same contract as vg_flat_tc.cu: the tensor cores only FILTER; every returned row is scored in simd.SquaredL2 /
// simd.Dot order (floats_avx512.c:12-129) and a certificate proves that the result is the exact scan's, ties by row id
// included; queries without a proof are flagged (d_fail) and re-run by the caller.
//
// What differs is the shape of the filter.  vg_flat_tc.cu keeps (min, second min) of every group of rows and writes a
// [queries][groups] plane that a second kernel selects from; its epilogue costs 9 ALU-pipe cycles per 32 elements and
// is the bound for short vectors.  Here the epilogue is a THRESHOLD test: one FFMA per element (s = f_q acc + ||x||^2),
// a 3-input minimum tree per 32-row chunk, and one compare of the chunk minimum with the query's threshold T; only
// chunks that beat T (a few per thousand) are opened and their rows appended to a private candidate list.
//
//   phase 0  query preparation inside the kernel: each CTA converts its 128 float32 queries to fp16 (per-query
//            power-of-two scale) straight into the 128-byte-swizzled A tile, which stays resident for all row tiles.
//   phase S  the first S row tiles of every row split are contracted and reduced to chunk minima only (no threshold
//            yet): G_s sample minima per query, written to global memory.
//   barrier  grid-wide (all CTAs are co-resident: at most one CTA per SM, see launch_single).
//   T        the sample chunks of a query are dealt into B = 96 buckets; T(q) = the r'-th smallest bucket minimum,
//            r' = k + 2.  r' different buckets hold a row with s <= T, so at least r' rows of the segment lie at or
//            below T — deterministically — while the expected number of rows below T stays at a few dozen for any
//            segment size (the sample is ~25 % of the rows; expected count = rows * ln(B / (B - r')) / rows_per_bucket).
//   phase M  all row tiles: rows with s < T go to the (query, thread) candidate list.
//   barrier
//   phase X  one warp per query: gathers the query's lists, keeps the kc = 32 candidates with the smallest filter
//            score (tau = the largest of them), scores those rows exactly, sorts by (score, row) and checks the
//            certificate: a row that was NOT scored has s >= tau (not selected) or s >= T >= tau (never listed), so its
//            exact score exceeds tau - E; if the exact k-th best is below that, nothing else can tie or beat it.
//
// CTA pair = cluster of two CTAs, tcgen05.mma.cta_group::2.kind::f16 (M = 256 queries x N = 256 rows), B tiles by TMA
// from the fp16 shadow of the vectors (see vg_flat_tc.cu), accumulators in TMEM (2 stages x 256 columns).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "vg_flat_tc.cuh"
#include "vg_tc_ptx.cuh"

namespace vg {
namespace tc {
namespace fs {
using namespace vg::tc::pair;

constexpr int BM = 128;            // queries per CTA (UMMA M half)
constexpr int BN = 128;            // rows per CTA per tile (UMMA N half)
constexpr int TILE_ROWS = 256;
constexpr int BKH = 64;            // halves per k-block
constexpr int EPW = 16;            // epilogue warps: two groups of eight, one per accumulator stage
constexpr int EPG = 8;             // warps per epilogue group: thread = one query x 128 columns of the tile
constexpr int MAX_SAMP = 16;       // sample chunks a thread keeps in registers (4 per sample tile of its group)
constexpr int MAX_PL = 10;         // bucket minima per lane in the threshold search (Gs <= 320)
constexpr int NT = (2 + EPW) * 32;
constexpr int STAGES = 6;
constexpr int MAX_KB = 4;          // dim <= 256
constexpr int KB_BYTES = BM * BKH * 2;   // 16 KB: one k-block of this CTA's query half / of its row half
constexpr int KC = 32;             // candidates scored exactly per query (k <= 16: twice k)
constexpr int SEL_CAP = 64;        // ... plus ties of the kc-th filter score
constexpr int LIST_CAP = 1024;     // candidates gathered per query in phase X
constexpr size_t OFF_A = 0;
constexpr size_t OFF_B = OFF_A + (size_t)MAX_KB * KB_BYTES;
constexpr size_t OFF_XN = OFF_B + (size_t)STAGES * KB_BYTES;
constexpr size_t OFF_FQ = OFF_XN + (size_t)4 * TILE_ROWS * 4;   // two ||x||^2 buffers per epilogue group
constexpr size_t OFF_T = OFF_FQ + (size_t)BM * 4;
constexpr size_t OFF_BAR = OFF_T + (size_t)BM * 4;
constexpr size_t SMEM_BYTES = OFF_BAR + (size_t)(2 * STAGES + 5) * 8 + 16 + 1024;
// phase X reuses the operand region: per warp LIST_CAP keys (8 B) + SEL_CAP exact keys
constexpr size_t X_WARP_BYTES = (size_t)LIST_CAP * 8 + (size_t)SEL_CAP * 8;
static_assert((size_t)(2 + EPW) * X_WARP_BYTES <= OFF_XN, "phase X scratch must fit in the operand region");

struct Args {
    const float *queries;      // [nq][dim] float32
    const float *vectors;      // [rows][dim] float32 (exact stage)
    const float *xn;           // [rows] ||x||^2 (L2) — unused for dot
    const uint32_t *mask;      // optional row bitmap
    const unsigned int *xmax_bits;
    int64_t nq, rows, rows_per_split;
    int dim, dimp, kb, x16_exp, is_dot, k;
    int splits, S, Gs, rprime, cap, slots;   // sample tiles per split, sample chunks per query, buckets, list capacity, lists per query
    uint32_t row_base;
    // scratch (global)
    float *smin;               // [nq_pad][Gs]
    float *Tq;                 // [nq]
    float *qn;                 // [nq]
    uint2 *cand;               // [nq_pad][slots][cap] (filter score bits, local row)
    int *ccnt;                 // [nq_pad][slots]
    float4 *samp;              // [nq_pad][slots][MAX_SAMP] sample chunk results (m1, m2, m3, -), thread-private
    unsigned int *sync;        // [4] grid barrier counters (zeroed before the launch) followed by
    int *ovf;                  // [nq_pad] overflow flags (zeroed before the launch)
    // outputs
    uint32_t *out_rows;
    float *out_scores;
    int32_t *out_counts;
    int32_t *fail;
    int nohit;                 // measurement aid (VECGO_FS_NOHIT=1): thresholds at -inf, every query ends flagged
    unsigned long long *dbg;   // optional [16] phase timestamps of CTA 0 (VECGO_FS_DEBUG=1)
};
__device__ __forceinline__ void stamp(const Args &A, int i) {
    if (A.dbg && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        A.dbg[i] = t;
        if (i == 1 || i == 5) A.dbg[10 + (i == 5)] = (unsigned long long)clock64();
    }
}

__device__ __forceinline__ void grid_barrier(unsigned int *ctr, unsigned int nctas) {
    // one thread per CTA; the caller brackets this with CTA-level barriers
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int v;
    do {
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v < nctas) __nanosleep(64);
    } while (v < nctas);
    __threadfence();
}

// minimum of 8 consecutive elements (4 FMNMX3 / FMNMX) — the chunk minimum is the minimum of four of these, and a
// chunk that beats the threshold is opened one 8-element group at a time
__device__ __forceinline__ float min8(const float (&s)[32], int o) {
    return fminf(fminf(fminf(s[o], s[o + 1]), s[o + 2]), fminf(fminf(fminf(s[o + 3], s[o + 4]), s[o + 5]), fminf(s[o + 6], s[o + 7])));
}

// T of this CTA's queries r = ew, ew + 16, ... (see the call site); PL = bucket minima per lane.
template <int PL, int NQ>
__device__ __forceinline__ void threshold_search(const Args &A, int64_t q0, int ew, int lane, float *T_s) {
    for (int r = ew; r < BM; r += NQ * EPW) {
        uint32_t ob[NQ][PL];
#pragma unroll
        for (int u = 0; u < NQ; u++) {
            const float *src = A.smin + (size_t)(q0 + r + u * EPW) * A.Gs;
#pragma unroll
            for (int j = 0; j < PL; j++) {
                const int g = j * 32 + lane;
                ob[u][j] = (g < A.Gs && r + u * EPW < BM) ? f32_orderable(__ldcg(src + g)) : 0xFFFFFFFFu;
            }
        }
        uint32_t prefix[NQ];
#pragma unroll
        for (int u = 0; u < NQ; u++) prefix[u] = 0u;
        for (int bit = 31; bit >= 14; bit--) {
            int cnt[NQ];
#pragma unroll
            for (int u = 0; u < NQ; u++) {
                const uint32_t trial = prefix[u] | (1u << bit);
                cnt[u] = 0;
#pragma unroll
                for (int j = 0; j < PL; j++) cnt[u] += ob[u][j] < trial ? 1 : 0;
            }
#pragma unroll
            for (int u = 0; u < NQ; u++) cnt[u] = __reduce_add_sync(0xffffffffu, cnt[u]);
#pragma unroll
            for (int u = 0; u < NQ; u++)
                if (cnt[u] < A.rprime) prefix[u] |= 1u << bit;   // the r'-th smallest is >= trial
        }
#pragma unroll
        for (int u = 0; u < NQ; u++)
            if (lane == 0 && r + u * EPW < BM) T_s[r + u * EPW] = f32_from_orderable(prefix[u] | 0x3FFFu);
    }
}

template <bool IS_DOT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) flat_single_kernel(const __grid_constant__ CUtensorMap map_x, Args A) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (A.dbg && tid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(A.dbg + 12, t);
    }
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair_id = (int)(blockIdx.x >> 1);
    const int qtile = pair_id / A.splits, split = pair_id - qtile * A.splits;
    const int64_t q0 = (int64_t)qtile * 256 + (int64_t)rank * BM;   // first query of this CTA
    const int64_t row_begin = (int64_t)split * A.rows_per_split;
    int64_t row_end = row_begin + A.rows_per_split;
    if (row_end > A.rows) row_end = A.rows;
    const int ntiles = row_end > row_begin ? (int)((row_end - row_begin + TILE_ROWS - 1) / TILE_ROWS) : 0;
    const int S = A.S < ntiles ? A.S : ntiles;   // sample tiles = the first S tiles of the split; they are NOT contracted twice
    const unsigned int nctas = gridDim.x;

    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar0 = s_base + (uint32_t)OFF_BAR;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
    const uint32_t aready_bar = bar0 + 8u * (2 * STAGES + 4);
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), EPG + EPG);   // the eight warps of one epilogue group in each CTA of the pair
        }
        mbar_init(aready_bar, EPW + EPW);          // every epilogue warp of both CTAs once its share of the A tile is written
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (tid == 0) stamp(A, 0);
    float *fq_s = reinterpret_cast<float *>(smem + OFF_FQ);
    float *T_s = reinterpret_cast<float *>(smem + OFF_T);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // barriers and the TMEM allocation of both CTAs exist: TMA may start while the queries are converted
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer: this CTA's row half of every k-block =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = 0; t < ntiles; t++) {
                const int n0 = (int)(row_begin + (int64_t)t * TILE_ROWS) + (int)rank * BN;
                for (int kb = 0; kb < A.kb; kb++, it++) {
                    const int st = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(empty_bar(st), ph ^ 1);
                    if (leader) mbar_expect_tx(full_bar(st), 2 * KB_BYTES);
                    tma_load_2d_pair(s_base + (uint32_t)OFF_B + st * KB_BYTES, &map_x, kb * BKH, n0, full_bar(st));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16_pair();
            mbar_wait_cluster(aready_bar, 0);   // both halves of the query tile are in shared memory
            tc_fence_after();
            uint32_t it = 0;
            for (int t = 0; t < ntiles; t++) {
                const int as = t & 1;
                const uint32_t aph = (t >> 1) & 1;
                mbar_wait_cluster(tempty_bar(as), aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * TILE_ROWS);
                for (int kb = 0; kb < A.kb; kb++, it++) {
                    const int st = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait_cluster(full_bar(st), ph);
                    tc_fence_after();
                    const uint64_t adesc = make_sdesc(s_base + (uint32_t)OFF_A + (uint32_t)kb * KB_BYTES);
                    const uint64_t bdesc = make_sdesc(s_base + (uint32_t)OFF_B + st * KB_BYTES);
#pragma unroll
                    for (int k = 0; k < BKH / 16; k++)
                        umma_f16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit_pair(empty_bar(st));
                }
                umma_commit_pair(tfull_bar(as));
            }
        }
    } else {
        // ===================== epilogue warps =====================
        // Two groups of eight warps; group g serves accumulator stage g (tiles t = g, g + 2, ...), so two tiles are in the
        // epilogue at once and the per-tile latencies (barrier, tcgen05.ld round trips) overlap.  Inside a group:
        // thread = one query (TMEM lane) x one 128-column half of the tile = four 32-column chunks.
        const int ew = warp - 2;
        const int grp = ew >> 3, gw = ew & 7;
        const int quad = warp & 3;                // TMEM lanes 32*quad.. are the ones this warp may read
        const int colh = gw >> 2;                 // which 128-column half
        const int gt = gw * 32 + lane;            // thread inside the group, 0..255
        const int et = ew * 32 + lane;            // 0..511
        const int qslot = quad * 32 + lane;       // query inside this CTA
        const int64_t q = q0 + qslot;             // scratch arrays are padded to whole query tiles: q indexes them directly
        const float BIG = 3.0e38f;

        // ---- phase 0: this CTA's 128 queries -> fp16, 128-byte-swizzled A tile (resident for every row tile).
        // a16 = half(q * 2^e), e such that max|q_i| lands in [2^11, 2^12); f_q = -(2 | 1) / (2^e 2^sx).  Row r of k-block kb
        // sits at kb*16K + (r/8)*1024 + (r%8)*128 with its 16-byte chunk c stored at chunk (c ^ (r%8)) — what TMA writes for
        // SWIZZLE_128B and what the UMMA descriptor (make_sdesc) expects.  Warp ew converts queries ew, ew + 16, ...: four at
        // a time so that four 1 KB loads are in flight.
        for (int r0 = ew; r0 < BM; r0 += 4 * EPW) {
            float v[4][8];   // lane covers dims [8*lane, 8*lane+8)
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int r = r0 + u * EPW;
                const int64_t qq_ = q0 + r;
                const bool live = r < BM && qq_ < A.nq;
                const float *qv = A.queries + (live ? qq_ : 0) * (int64_t)A.dim;
                const bool in = live && lane * 8 < A.dim;
                if (in && lane * 8 + 8 <= A.dim && (A.dim & 3) == 0) {
                    const float4 a = __ldg(reinterpret_cast<const float4 *>(qv + lane * 8));
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(qv + lane * 8 + 4));
                    v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w;
                    v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[u][i] = (in && lane * 8 + i < A.dim) ? __ldg(qv + lane * 8 + i) : 0.0f;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int r = r0 + u * EPW;
                if (r >= BM) continue;
                const int64_t qq_ = q0 + r;
                const bool live = qq_ < A.nq;
                float mx = 0.0f, nrm = 0.0f;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    mx = fmaxf(mx, fabsf(v[u][i]));
                    nrm = __fmaf_rn(v[u][i], v[u][i], nrm);
                }
                for (int o = 16; o > 0; o >>= 1) {
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    nrm = __fadd_rn(nrm, __shfl_xor_sync(0xffffffffu, nrm, o));
                }
                int e = 0;
                if (mx > 0.0f && mx < __int_as_float(0x7f800000)) {
                    int ex;
                    frexpf(mx, &ex);
                    e = 12 - ex;
                    e = e > 100 ? 100 : (e < -100 ? -100 : e);
                }
                const float sq = ldexpf(1.0f, e);
                if (lane * 8 < A.dimp) {
                    uint32_t h[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const __half2 p = __floats2half2_rn(__fmul_rn(v[u][2 * i], sq), __fmul_rn(v[u][2 * i + 1], sq));
                        h[i] = *reinterpret_cast<const uint32_t *>(&p);
                    }
                    const int kb = lane >> 3, c = lane & 7;   // 8 lanes (64 halves) per k-block, lane = 16-byte chunk c
                    const uint32_t addr = s_base + (uint32_t)OFF_A + (uint32_t)kb * KB_BYTES + (uint32_t)(r >> 3) * 1024u +
                                          (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
                }
                if (lane == 0) {
                    fq_s[r] = live ? -ldexpf(1.0f, (IS_DOT ? 0 : 1) - e - A.x16_exp) : 0.0f;
                    if (live && split == 0) A.qn[qq_] = nrm;
                }
            }
        }
        fence_proxy_async_smem();   // generic-proxy writes of the A tile -> visible to the tensor core's operand reads
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(aready_bar, 0);
        asm volatile("bar.sync 3, %0;" ::"n"(EPW * 32) : "memory");   // fq_s is complete
        if (et == 0) stamp(A, 1);
        const float fq = fq_s[qslot];

        float *xs = reinterpret_cast<float *>(smem + OFF_XN) + grp * (2 * TILE_ROWS);   // this group's two ||x||^2 buffers
        auto xn_of = [&](int t_) {
            const int64_t row = row_begin + (int64_t)t_ * TILE_ROWS + gt;
            return (t_ < ntiles && row < row_end) ? (IS_DOT ? 0.0f : __ldg(A.xn + row)) : BIG;
        };
        const int myslot = split * 4 + grp * 2 + colh;
        uint2 *mylist = A.cand + ((size_t)q * A.slots + myslot) * A.cap;
        int ncand = 0;
        float T = 0.0f;
        // the three smallest values (each with its row's position in the low 5 bits) of this thread's sample chunks: parked
        // in a private stretch of global scratch until the threshold is known (48 registers otherwise)
        float4 *samp = A.samp + ((size_t)q * A.slots + myslot) * MAX_SAMP;
        int nsamp = 0;
        float bucket = BIG;      // this thread's bucket minimum = the minimum over its sample chunks
        float xn_cur = xn_of(grp);
        int iter = 0;
        bool past_sample = false;
        auto boundary = [&]() {
            // ---- every sample minimum is written: grid barrier, thresholds of this CTA's queries, then list the sample rows
            A.smin[(size_t)q * A.Gs + myslot] = bucket;
            asm volatile("bar.sync 3, %0;" ::"n"(EPW * 32) : "memory");
            if (et == 0) stamp(A, 2);
            if (et == 0) grid_barrier(A.sync + 0, nctas);
            asm volatile("bar.sync 3, %0;" ::"n"(EPW * 32) : "memory");
            if (et == 0) stamp(A, 3);
            // Warp ew computes T of queries ew, ew + 16, ... several at a time (independent count chains): lane l holds bucket
            // minima l, l + 32, ...; T = the r'-th smallest, found by a bitwise search on the top 18 bits of the orderable score
            // (one warp-wide count per bit) and rounded UP — any T with at least r' buckets at or below it is valid.
            if (A.Gs <= 96) threshold_search<3, 8>(A, q0, ew, lane, T_s);   // all eight queries of the warp in one round
            else threshold_search<MAX_PL, 2>(A, q0, ew, lane, T_s);
            asm volatile("bar.sync 3, %0;" ::"n"(EPW * 32) : "memory");
            T = A.nohit ? -BIG : fminf(T_s[qslot], 2.9e38f);   // masked and padding rows carry s = BIG: never listed, whatever the threshold
            if (et == 0) stamp(A, 4);
            if (grp == 0 && colh == 0 && split == 0 && q < A.nq) A.Tq[q] = T;
            // sample rows below T: the chunk's arg-min row (named by the low 5 bits of m1); a chunk whose second minimum beats
            // T as well lists its other 31 rows with m2 as a LOWER bound of their filter score (exact scoring sorts it out)
            for (int i = 0; i < nsamp; i++) {
                const float4 sp = samp[i];
                if (sp.x < T) {
                    const int t_ = grp + 2 * (i >> 2), c = i & 3;
                    const uint32_t r0 = (uint32_t)(row_begin + (int64_t)t_ * TILE_ROWS + colh * 128 + c * 32);
                    const uint32_t pos1 = __float_as_uint(sp.x) & 31u;
                    if (ncand < A.cap) mylist[ncand] = make_uint2(__float_as_uint(sp.x), r0 + pos1);
                    ncand++;
                    if (sp.y < T) {
                        const uint32_t pos2 = __float_as_uint(sp.y) & 31u;
                        if (ncand < A.cap) mylist[ncand] = make_uint2(__float_as_uint(sp.y), r0 + pos2);
                        ncand++;
                        if (sp.z < T) {
                            // three rows of one sample chunk below T (a few per thousand queries): the other thirty are listed
                            // with the third minimum as a LOWER bound of their filter score; exact scoring sorts it out
                            const uint32_t mw = A.mask ? ((int64_t)r0 < A.rows ? __ldg(A.mask + (r0 >> 5)) : 0u) : 0xFFFFFFFFu;
                            for (uint32_t j = 0; j < 32; j++) {
                                if (j == pos1 || j == pos2 || !((mw >> j) & 1u) || (int64_t)(r0 + j) >= row_end) continue;
                                if (ncand < A.cap) mylist[ncand] = make_uint2(__float_as_uint(sp.z), r0 + j);
                                ncand++;
                            }
                        }
                    }
                }
            }
        };
        for (int t = grp; t < ntiles; t += 2, iter++) {
            const bool sample = t < S;
            if (!sample && !past_sample) {
                boundary();
                past_sample = true;
            }
            const uint32_t aph = (t >> 1) & 1;
            float *xt = xs + (iter & 1) * TILE_ROWS;
            xt[gt] = xn_cur;
            asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(EPG * 32) : "memory");   // also: everyone is done with the buffer of tile t-2... (t-4's reads ended two barriers ago)
            xn_cur = xn_of(t + 2);
            mbar_wait(tfull_bar(grp), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * TILE_ROWS + colh * 128);
            const int64_t nh = row_begin + (int64_t)t * TILE_ROWS + colh * 128;
            // 18 warps leave 96 registers per thread (five warps share one register-file partition): one accumulator buffer,
            // no software prefetch — the other warps of the scheduler cover the tcgen05.ld round trip
#pragma unroll 1
            for (int c = 0; c < 4; c++) {
                uint32_t vn[32];
                tmem_ld32(taddr + (uint32_t)(c * 32), vn);
                uint32_t mw = 0xFFFFFFFFu;
                if (A.mask) mw = (nh + c * 32 < A.rows) ? __ldg(A.mask + ((nh + c * 32) >> 5)) : 0u;
                tmem_ld_wait();
                float s[32];
                const float4 *x4 = reinterpret_cast<const float4 *>(xt + colh * 128 + c * 32);
#pragma unroll
                for (int j4 = 0; j4 < 8; j4++) {
                    const float4 xv = x4[j4];
                    s[j4 * 4 + 0] = __fmaf_rn(fq, __uint_as_float(vn[j4 * 4 + 0]), xv.x);
                    s[j4 * 4 + 1] = __fmaf_rn(fq, __uint_as_float(vn[j4 * 4 + 1]), xv.y);
                    s[j4 * 4 + 2] = __fmaf_rn(fq, __uint_as_float(vn[j4 * 4 + 2]), xv.z);
                    s[j4 * 4 + 3] = __fmaf_rn(fq, __uint_as_float(vn[j4 * 4 + 3]), xv.w);
                }
                if (mw != 0xFFFFFFFFu) {
#pragma unroll
                    for (int j = 0; j < 32; j++) s[j] = (mw >> j) & 1u ? s[j] : BIG;
                }
                if (sample) {
                    // (min carrying the row's position in its low 5 bits, second min) — the filter epilogue of vg_flat_tc.cu;
                    // the chunk's pair stays in registers until the threshold is known
                    // three smallest values of the chunk, every value carrying its row's position in its low 5 bits: the two
                    // smallest name their rows; the third is a lower bound for the other thirty
                    float a1[2] = {BIG, BIG}, a2[2] = {BIG, BIG}, a3[2] = {BIG, BIG};
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        uint32_t vb;  // (bits & ~31) | j as one LOP3
                        asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(vb) : "r"(__float_as_uint(s[j])), "r"(0xFFFFFFE0u), "r"((uint32_t)j));
                        const float v1 = __uint_as_float(vb);
                        const int h = j & 1;
                        a3[h] = fminf(a3[h], fmaxf(a2[h], v1));
                        a2[h] = fminf(a2[h], fmaxf(a1[h], v1));
                        a1[h] = fminf(a1[h], v1);
                    }
                    // merge the two sorted triples (a1 <= a2 <= a3 each) into the three smallest of the six
                    const float m1 = fminf(a1[0], a1[1]);
                    const float hi1 = fmaxf(a1[0], a1[1]);
                    const float lo2 = fminf(a2[0], a2[1]);
                    const float m2 = fminf(hi1, lo2);
                    const float m3 = fminf(fmaxf(hi1, lo2), fminf(fmaxf(a2[0], a2[1]), fminf(a3[0], a3[1])));
                    samp[(iter << 2) + c] = make_float4(m1, m2, m3, 0.0f);   // iter < MAX_SAMP / 4 for sample tiles (make_plan)
                    nsamp = (iter << 2) + c + 1;
                    bucket = fminf(bucket, m1);
                } else {
                    const float g0 = min8(s, 0), g1 = min8(s, 8), g2 = min8(s, 16), g3 = min8(s, 24);
                    const float m = fminf(fminf(g0, g1), fminf(g2, g3));
                    if (m < T) {
                        // a chunk that beats the threshold (a few per thousand): list every row below T, one 8-row group at a time
                        const uint32_t r0 = (uint32_t)(nh + c * 32);
                        const float gm[4] = {g0, g1, g2, g3};
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
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_bar(grp), 0);
        }
        if (!past_sample) boundary();   // a group whose tiles were all sample tiles (or that had none)
        if (et == 0) stamp(A, 5);
        A.ccnt[(size_t)q * A.slots + myslot] = ncand < A.cap ? ncand : A.cap;
        if (ncand > A.cap) A.ovf[q] = 1;
    }
    // ===================== every list is written: second grid barrier, then phase X =====================
    tc_fence_before();
    __syncthreads();
    if (tid == 0) grid_barrier(A.sync + 1, nctas);
    __syncthreads();
    if (tid == 0) stamp(A, 6);
    cluster_sync_all();   // the pair's MMAs are complete (every tile was consumed) before TMEM goes away
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
    {
        // Queries are dealt round-robin over all warps of the grid; the CTA works in rounds of one query per warp:
        //   step 1 (warp per query)   gather the query's lists, tau = kc-th smallest filter score, selected rows
        //   step 2 (whole CTA)        every half-warp scores (query, row) items of ALL the round's queries exactly —
        //                             two items in flight per half-warp: the row gathers are latency-bound
        //   step 3 (warp per query)   rank by (score, row), emit the best k, certificate
        constexpr int NW = NT / 32;
        unsigned char *xs_base = smem;   // operand region (A tile, B ring): free now
        unsigned long long *keys = reinterpret_cast<unsigned long long *>(xs_base + (size_t)warp * X_WARP_BYTES);
        auto ekeys_of = [&](int w) { return reinterpret_cast<unsigned long long *>(xs_base + (size_t)w * X_WARP_BYTES) + LIST_CAP; };
        int *meta = reinterpret_cast<int *>(smem + OFF_XN);   // [NW][4]: nsel, fail, tau bits, valid
        const int hw_id = tid >> 4, hl = tid & 15;           // 36 half-warps
        const int64_t epochs = A.dim >> 6;
        const int64_t per_cta = (A.nq + nctas - 1) / nctas;
        const int rounds = (int)((per_cta + NW - 1) / NW);
        for (int rd = 0; rd < rounds; rd++) {
            const int64_t q = (int64_t)blockIdx.x + ((int64_t)rd * NW + warp) * (int64_t)nctas;
            const bool qvalid = q < A.nq;
            int fail = 0, nsel = 0;
            uint32_t tau_o = 0;
            if (qvalid) {
                // counts of all lists first (MAX_PL rounds of 32 lists), then their entries: two dependent round trips in all
                int cs[MAX_PL], offs[MAX_PL];
                const int ovf_q = __ldcg(A.ovf + q);
                const float T_q = __ldcg(A.Tq + q);
#pragma unroll
                for (int j = 0; j < MAX_PL; j++) {
                    const int sl = j * 32 + lane;
                    cs[j] = sl < A.slots ? __ldcg(A.ccnt + (size_t)q * A.slots + sl) : 0;
                }
                fail = ovf_q ? 1 : 0;   // nonzero = no proof; the value says why (1 list overflow, 2 too many candidates, 3 ties, 4 < k rows, 5 certificate)
                int n = 0;
#pragma unroll
                for (int j = 0; j < MAX_PL; j++) {
                    if (j * 32 >= A.slots) {   // warp-uniform: no lists in this round
                        offs[j] = n;
                        continue;
                    }
                    int incl = cs[j];
                    for (int o = 1; o < 32; o <<= 1) {
                        const int y = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += y;
                    }
                    offs[j] = n + incl - cs[j];
                    n += __shfl_sync(0xffffffffu, incl, 31);
                }
#pragma unroll
                for (int j = 0; j < MAX_PL; j++) {
                    if (j * 32 >= A.slots) continue;
                    const int sl = j * 32 + lane;
                    const uint2 *src = A.cand + ((size_t)q * A.slots + (sl < A.slots ? sl : 0)) * A.cap;
                    for (int i = 0; i < cs[j]; i += 4) {
                        uint2 e[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) e[u] = (i + u < cs[j]) ? __ldcg(src + i + u) : make_uint2(0u, 0u);
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (i + u < cs[j] && offs[j] + i + u < LIST_CAP)
                                keys[offs[j] + i + u] = ((unsigned long long)f32_orderable(__uint_as_float(e[u].x)) << 32) | e[u].y;
                    }
                }
                if (n > LIST_CAP) {
                    fail = 2;
                    n = LIST_CAP;
                }
                __syncwarp();
                // tau = (an upper bound of) the kc-th smallest filter score: bitwise search on the top 24 bits of the orderable
                // score, rounded up; every candidate at or below it is selected
                tau_o = f32_orderable(T_q);   // fewer than kc candidates: all are scored, the bound is T itself
                nsel = n;
                if (n > KC) {
                    uint32_t prefix = 0;
                    for (int bit = 31; bit >= 8; bit--) {
                        const uint32_t trial = prefix | (1u << bit);
                        int cnt = 0;
                        for (int i = lane; i < n; i += 32) cnt += ((uint32_t)(keys[i] >> 32) < trial) ? 1 : 0;
                        cnt = __reduce_add_sync(0xffffffffu, cnt);
                        if (cnt < KC) prefix = trial;   // the kc-th smallest score is >= trial
                    }
                    tau_o = prefix | 0xFFu;
                    int cnt = 0;
                    for (int i = lane; i < n; i += 32) cnt += ((uint32_t)(keys[i] >> 32) <= tau_o) ? 1 : 0;
                    nsel = __reduce_add_sync(0xffffffffu, cnt);
                    if (nsel > SEL_CAP && !fail) fail = 3;   // massive ties at the kc-th filter score
                }
                // selected rows -> front of this warp's ekeys (row ids for now)
                unsigned long long *ek = ekeys_of(warp);
                int base = 0;
                for (int i0 = 0; i0 < n; i0 += 32) {
                    const int i = i0 + lane;
                    const bool keep = i < n && (n <= KC || (uint32_t)(keys[i] >> 32) <= tau_o);
                    const unsigned b = __ballot_sync(0xffffffffu, keep);
                    const int pos = base + __popc(b & ((1u << lane) - 1u));
                    if (keep && pos < SEL_CAP) ek[pos] = keys[i] & 0xFFFFFFFFull;
                    base += __popc(b);
                }
                if (nsel > SEL_CAP) nsel = SEL_CAP;
            }
            if (tid == 0) stamp(A, 7);
            if (lane == 0) {
                meta[warp * 4 + 0] = nsel;
                meta[warp * 4 + 1] = fail;
                meta[warp * 4 + 2] = (int)tau_o;
                meta[warp * 4 + 3] = qvalid ? 1 : 0;
            }
            __syncthreads();
            // ---- step 2: exact scores in simd.SquaredL2 / simd.Dot order (floats_avx512.c:12-129): half-warp per row, 4 x 16-lane
            // FMA accumulators over 64-dim epochs, (A1+A2)+(A3+A4), lane tree, FMA scalar tail
            {
                int total = 0;
                for (int w = 0; w < NW; w++) total += meta[w * 4];
                // every thread runs the same number of iterations (the lane tree below shuffles across the whole warp)
                for (int tb = 0; tb < total; tb += 2 * (NT / 16)) {
                    int wq[2], rr[2];
                    bool val[2];
                    const float *xp[2], *qp[2];
                    int64_t rowi[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int t = tb + hw_id + u * (NT / 16);
                        val[u] = t < total;
                        int w = 0, acc = 0;
                        if (val[u]) {
                            while (acc + meta[w * 4] <= t) {
                                acc += meta[w * 4];
                                w++;
                            }
                        }
                        wq[u] = w;
                        rr[u] = val[u] ? t - acc : 0;
                        const int64_t qq_ = (int64_t)blockIdx.x + ((int64_t)rd * NW + w) * (int64_t)nctas;
                        rowi[u] = val[u] ? (int64_t)(uint32_t)ekeys_of(w)[rr[u]] : 0;
                        xp[u] = A.vectors + rowi[u] * A.dim;
                        qp[u] = A.queries + (val[u] ? qq_ : 0) * (int64_t)A.dim;
                    }
                    float a[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                    for (int64_t e = 0; e < epochs; e++) {
                        float xv[2][4], qv[2][4];
#pragma unroll
                        for (int u = 0; u < 2; u++)
#pragma unroll
                            for (int jj = 0; jj < 4; jj++) {
                                xv[u][jj] = __ldg(xp[u] + e * 64 + jj * 16 + hl);
                                qv[u][jj] = __ldg(qp[u] + e * 64 + jj * 16 + hl);
                            }
#pragma unroll
                        for (int u = 0; u < 2; u++)
#pragma unroll
                            for (int jj = 0; jj < 4; jj++) {
                                if (IS_DOT) {
                                    a[u][jj] = __fmaf_rn(qv[u][jj], xv[u][jj], a[u][jj]);
                                } else {
                                    const float df = __fsub_rn(qv[u][jj], xv[u][jj]);
                                    a[u][jj] = __fmaf_rn(df, df, a[u][jj]);
                                }
                            }
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        float tot = reduce16(__fadd_rn(__fadd_rn(a[u][0], a[u][1]), __fadd_rn(a[u][2], a[u][3])));
                        if (hl == 0 && val[u]) {
                            for (int64_t d = epochs * 64; d < A.dim; d++) {
                                const float qd = __ldg(qp[u] + d), xd = __ldg(xp[u] + d);
                                if (IS_DOT) {
                                    tot = __fmaf_rn(qd, xd, tot);
                                } else {
                                    const float df = __fsub_rn(qd, xd);
                                    tot = __fmaf_rn(df, df, tot);
                                }
                            }
                            ekeys_of(wq[u])[rr[u]] = make_key(tot, A.row_base + (uint32_t)rowi[u], IS_DOT);
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) stamp(A, 8);
            // ---- step 3: rank by counting (<= 64 distinct keys), emit the best k, certificate
            if (qvalid) {
                const unsigned long long *ek = ekeys_of(warp);
                const int kk = A.k;
                for (int i = lane; i < kk; i += 32) {
                    A.out_rows[q * kk + i] = 0xFFFFFFFFu;
                    A.out_scores[q * kk + i] = __uint_as_float(0x7fc00000u);
                }
                __syncwarp();
                unsigned long long kth = VG_KEY_EMPTY;
                for (int i = lane; i < nsel; i += 32) {
                    const unsigned long long me = ek[i];
                    int rk = 0;
                    for (int j = 0; j < nsel; j++) rk += ek[j] < me ? 1 : 0;
                    if (rk < kk) {
                        A.out_rows[q * kk + rk] = key_row(me);
                        A.out_scores[q * kk + rk] = key_score(me, IS_DOT);
                    }
                    if (rk == kk - 1) kth = me;
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long y = __shfl_xor_sync(0xffffffffu, kth, o);
                    kth = y < kth ? y : kth;
                }
                if (lane == 0) {
                    const int m = nsel < kk ? nsel : kk;
                    A.out_counts[q] = m;
                    if (!fail) {
                        if (m < kk) {
                            fail = 4;   // fewer than k rows below the threshold (row masks can do it)
                        } else {
                            const float tau = f32_from_orderable(tau_o);
                            const double qq = (double)__ldcg(A.qn + q), xx = (double)__uint_as_float(*A.xmax_bits);
                            // fp16 operands round to 2^-11 relative each; fp32 accumulation over dim terms (vg_flat_tc.cu, pair kernel)
                            const double c1 = (IS_DOT ? 1.0 / 512.0 : 1.0 / 256.0) * 1.125 * 0.5;
                            const double c2 = 1.0 / 16384.0 + (double)A.dim / 8388608.0;
                            const double smax = IS_DOT ? sqrt(qq * xx) : xx + 2.0 * sqrt(qq * xx);   // |s| of any row
                            const double E = c1 * sqrt(qq * xx) + c2 * (qq + xx) + smax * 32.0 / 8388608.0;   // last: the 5 index bits of sample minima
                            const double ex = (double)key_score(kth, IS_DOT);
                            const double s_exact = IS_DOT ? -ex : ex - qq;
                            if (!(s_exact < (double)tau - E)) fail = 5;
                        }
                    }
                    A.fail[q] = fail ? (A.dbg ? fail : 1) : 0;   // the reason code only under VECGO_FS_DEBUG
                }
            }
            __syncthreads();
            if (tid == 0) stamp(A, 9);
        }
    }
    if (A.dbg && tid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(A.dbg + 13, t);
    }
}

// ------------------------------------------------------------------ host
static std::atomic<int> g_single{-1};
bool single_enabled() {
    int v = g_single.load();
    if (v < 0) {
        const char *e = getenv("VECGO_FLAT_SINGLE");
        v = (e && e[0] == '0') ? 0 : 1;
        g_single.store(v);
    }
    return v != 0;
}

struct Plan {
    int64_t qtiles, splits, rps;
    int S, Gs, rprime, cap, slots;
};
static bool make_plan(int64_t nq, int64_t rows, int k, Plan &p) {
    const int64_t pairs = sm_count() / 2;
    p.qtiles = (nq + 255) / 256;
    if (p.qtiles < 1 || p.qtiles > pairs / 2) return false;
    p.splits = pairs / p.qtiles;
    const int64_t tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
    if (p.splits > tiles / 4) p.splits = tiles / 4;
    if (p.splits < 1) return false;
    int64_t tps = (tiles + p.splits - 1) / p.splits;   // tiles per split
    p.rps = tps * TILE_ROWS;
    p.splits = (rows + p.rps - 1) / p.rps;
    // sample: ~25 % of the tiles of every split (at most MAX_SAMP / 4 per epilogue group).  Every (query, split, group,
    // column half) thread keeps ONE bucket minimum over its sample chunks: Gs = slots buckets per query.
    int S = (int)std::max<int64_t>(2, (tps + 2) / 4);
    if (S > 2 * (MAX_SAMP / 4)) S = 2 * (MAX_SAMP / 4);
    if (S > tps) S = (int)tps;
    p.S = S;
    p.slots = (int)(p.splits * 4);
    p.Gs = p.slots;
    p.rprime = k + 12;   // at least r' rows lie below T: a few more than k, so that tau - E clears the k-th best
    if (p.Gs < 2 * p.rprime || p.Gs > 32 * MAX_PL) return false;
    // expected rows below T = rows * ln(B / (B - r')) / rows_per_bucket (header); lists hold 6x the per-thread mean + 48
    // (+ room for one fully listed sample chunk)
    const double rows_per_bucket = (double)S * TILE_ROWS / 4.0;
    const double per_query = (double)rows * std::log((double)p.Gs / (double)(p.Gs - p.rprime)) / rows_per_bucket;
    p.cap = 2 * (int)std::min<double>(2048.0, 40.0 + 3.0 * per_query / p.slots);   // even: the scratch behind the lists stays 16-byte aligned
    return per_query <= 0.6 * LIST_CAP;
}

bool single_supported(int64_t dim, int64_t rows, int64_t nq, int64_t k) {
    if (!single_enabled() || !pair_enabled()) return false;
    if (dim < 16 || dim > 256 || dim % 4 != 0 || rows < 8192 || rows >= (1ll << 31) || nq < 16 || k < 1 || k > 16) return false;
    Plan p;
    return make_plan(nq, rows, (int)k, p);
}

// The kernel has grid-wide barriers, so all its CTAs must become resident.  It is launched as an ordinary kernel (a
// cooperative launch costs ~40 us of extra launch latency — measured — which is the whole budget of this path): the grid
// never exceeds the SM count with one CTA per SM, kernels of OTHER kinds sharing the GPU finish on their own and free
// their SMs, and two kernels of THIS kind never overlap — each launch waits (on the device, cudaStreamWaitEvent) for the
// previous one on the same GPU, whichever stream it ran on.
static std::mutex g_chain_mu[64];
static cudaEvent_t g_chain_ev[64] = {};
template <bool IS_DOT>
static vg_status launch_single(const CUtensorMap &mx, const Args &a, int64_t ctas, cudaStream_t st) {
    VG_CUDA(cudaFuncSetAttribute(flat_single_kernel<IS_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    int dev = 0;
    VG_CUDA(cudaGetDevice(&dev));
    dev &= 63;
    std::lock_guard<std::mutex> lk(g_chain_mu[dev]);
    if (!g_chain_ev[dev]) VG_CUDA(cudaEventCreateWithFlags(&g_chain_ev[dev], cudaEventDisableTiming));
    else VG_CUDA(cudaStreamWaitEvent(st, g_chain_ev[dev], 0));
    flat_single_kernel<IS_DOT><<<dim3((unsigned)ctas), dim3(NT), SMEM_BYTES, st>>>(mx, a);
    VG_LAUNCHED();
    VG_CUDA(cudaEventRecord(g_chain_ev[dev], st));
    return VG_OK;
}

// One launch (plus one 4 KB memset): filter, selection, exact stage and certificate of a whole batch.
vg_status single_enqueue(const SearchIO &io, int32_t *d_fail, cudaStream_t st) {
    Plan p;
    if (!make_plan(io.nq, io.rows, io.k, p)) return fail(VG_ERR_UNSUPPORTED, "shape not supported by the single-launch flat search");
    const int dimp = (int)((io.dim + 63) / 64 * 64);
    const int64_t nq_pad = p.qtiles * 256;
    // one scratch block: [sync 16 B | ovf nq_pad ints] (zeroed) | smin | Tq | qn | ccnt | cand
    const size_t o_ovf = 16, o_smin = (o_ovf + (size_t)nq_pad * 4 + 255) & ~(size_t)255;
    const size_t o_T = o_smin + (size_t)nq_pad * p.Gs * 4, o_qn = o_T + (size_t)nq_pad * 4, o_cnt = o_qn + (size_t)nq_pad * 4;
    const size_t o_cand = (o_cnt + (size_t)nq_pad * p.slots * 4 + 255) & ~(size_t)255;   // o_samp below stays 16-byte aligned (cap * 8 per list)
    const size_t o_samp = o_cand + (size_t)nq_pad * p.slots * p.cap * 8;
    const size_t total = o_samp + (size_t)nq_pad * p.slots * MAX_SAMP * 16;
    DevBuf scratch;
    VG_TRY(scratch.alloc(total));
    unsigned char *base = scratch.as<unsigned char>();
    VG_CUDA(cudaMemsetAsync(base, 0, o_smin, st));
    CUtensorMap mx;
    VG_TRY(tensor_map_2d(&mx, true, io.d_x16, io.rows, dimp, dimp, BKH, BN));
    Args a{};
    a.queries = io.d_queries;
    a.vectors = io.d_vectors;
    a.xn = io.d_xn;
    a.mask = reinterpret_cast<const uint32_t *>(io.d_mask);
    a.xmax_bits = io.d_xmax_bits;
    a.nq = io.nq;
    a.rows = io.rows;
    a.rows_per_split = p.rps;
    a.dim = (int)io.dim;
    a.dimp = dimp;
    a.kb = dimp / BKH;
    a.x16_exp = io.x16_exp;
    a.is_dot = io.is_dot;
    a.k = io.k;
    a.splits = (int)p.splits;
    a.S = p.S;
    a.Gs = p.Gs;
    a.rprime = p.rprime;
    a.cap = p.cap;
    a.slots = p.slots;
    a.row_base = io.row_base;
    a.sync = reinterpret_cast<unsigned int *>(base);
    a.ovf = reinterpret_cast<int *>(base + o_ovf);
    a.smin = reinterpret_cast<float *>(base + o_smin);
    a.Tq = reinterpret_cast<float *>(base + o_T);
    a.qn = reinterpret_cast<float *>(base + o_qn);
    a.ccnt = reinterpret_cast<int *>(base + o_cnt);
    a.cand = reinterpret_cast<uint2 *>(base + o_cand);
    a.samp = reinterpret_cast<float4 *>(base + o_samp);
    a.out_rows = io.d_rows;
    a.out_scores = io.d_scores;
    a.out_counts = io.d_counts;
    a.fail = d_fail;
    const int64_t ctas = 2 * p.qtiles * p.splits;
    static const bool debug = getenv("VECGO_FS_DEBUG") != nullptr;
    static const bool nohit = getenv("VECGO_FS_NOHIT") != nullptr;
    a.nohit = nohit ? 1 : 0;
    DevBuf dbg;
    if (debug) {
        VG_TRY(dbg.alloc(16 * 8));
        VG_CUDA(cudaMemsetAsync(dbg.p, 0, 16 * 8, st));
        VG_CUDA(cudaMemsetAsync(dbg.as<unsigned long long>() + 12, 0xFF, 8, st));
        a.dbg = dbg.as<unsigned long long>();
    }
    VG_TRY(io.is_dot ? launch_single<true>(mx, a, ctas, st) : launch_single<false>(mx, a, ctas, st));
    if (debug) {
        unsigned long long h[16];
        VG_CUDA(cudaMemcpyAsync(h, dbg.p, sizeof h, cudaMemcpyDeviceToHost, st));
        VG_CUDA(cudaStreamSynchronize(st));
        static unsigned long long last_end = 0;
        fprintf(stderr, "[flat_single] first CTA start -> stamp0 %.1f us, stamp9 -> last CTA end %.1f us, whole %.1f us, SM clock %.0f MHz | ",
                (h[0] - h[12]) * 1e-3, (h[13] - h[9]) * 1e-3, (h[13] - h[12]) * 1e-3, (double)(h[11] - h[10]) / ((h[5] - h[1]) * 1e-3));
        last_end = h[9];
        fprintf(stderr, "[flat_single] ctas=%lld splits=%lld S=%d Gs=%d cap=%d | phase0 %.1f sample %.1f barrier %.1f T %.1f main %.1f barrier %.1f X1 %.1f X2 %.1f X3 %.1f us\n",
                (long long)ctas, (long long)p.splits, p.S, p.Gs, p.cap, (h[1] - h[0]) * 1e-3, (h[2] - h[1]) * 1e-3, (h[3] - h[2]) * 1e-3,
                (h[4] - h[3]) * 1e-3, (h[5] - h[4]) * 1e-3, (h[6] - h[5]) * 1e-3, (h[7] - h[6]) * 1e-3, (h[8] - h[7]) * 1e-3, (h[9] - h[8]) * 1e-3);
    }
    return VG_OK;
}

}  // namespace fs
}  // namespace tc
}  // namespace vg
