// vg_pq_assign_tc.cu — PQ-training assignment (pq.go:347-386 assignClusters / findNearestCentroid) for 8-dim subspaces
// and 256 centroids on the tcgen05 tensor cores, made exact by a gap certificate.
//
// The reference assigns every sample's subspace to argmin_k simd.SquaredL2(x_m, c_{m,k}) (strict <, first wins; for
// dsub = 8 the AVX-512 kernel is its sequential FMA tail).  argmin_k ||x - c_k||^2 = argmax_k (x.c_k - ||c_k||^2 / 2), a
// [samples x 8] . [8 x 256] contraction per subspace.  K = 8 would waste the MMA's 16-wide k-step, so the spare slots
// carry a hi/lo fp16 split of both operands and the -||c||^2/2 term:
//
//      A (sample)   : x_hi(8) | x_lo(8) | x_hi(8) | 1, 1, 1, 0...      (32 halves = two k-steps of kind::f16)
//      B (centroid) : c_hi(8) | c_hi(8) | c_lo(8) | h, m, l, 0...      (h + m + l = -||c||^2 / 2)
//      acc = x_hi.c_hi + x_lo.c_hi + x_hi.c_lo + h + m + l = x.c - ||c||^2/2 - (x_lo.c_lo, ~2^-22)
//
// i.e. a float32-grade score from fp16 tensor-core instructions, with no epilogue arithmetic at all: the epilogue only
// finds the row maximum and counts the columns within the certificate margin of it.  Samples are scaled by a power of two so that max|x| lies in
// [8, 16): fp16 keeps hi/lo in range and ||c||^2/2 <= 1024.  |acc_tc - acc| <= 2^-20 B with B = max||x_m||^2 +
// max||c_m||^2 (dropped lo.lo, lo rounding, ||c||^2 in float32, fp32 accumulation of two instructions); the reference's
// own float32 sum is within 2^-21 of the true distance.  A (sample, subspace) with exactly one score within 2^-16 B of its best
// therefore has a certain argmin — identical to the reference's, no tie possible; the others (~0.5 % on
// Gaussian data) are listed and re-evaluated by the exact sequential-FMA loop over all 256 centroids.
//
// Layout: one CTA (576 threads) per 128 samples walks the 48 subspace pairs; a 4-stage ring of (16 KB sample tile + 32 KB centroid
// tile) filled by TMA (128-byte swizzle); both accumulator halves of TMEM (2 x 256 columns) alternate between the two
// subspaces of a pair, so the MMA of the next subspace overlaps the epilogue of the current one.
#include <cuda.h>
#include <cuda_fp16.h>

#include <atomic>
#include <cstdlib>

#include "vg_flat_tc.cuh"
#include "vg_pq_assign_tc.cuh"
#include "vg_quant.cuh"
#include "vg_tc_ptx.cuh"

namespace vg {
namespace pqa {
using namespace tc;

constexpr int TM = 128;                    // samples per CTA (UMMA M)
constexpr int TN = 256;                    // centroids (UMMA N)
constexpr int KH = 64;                     // halves per k-block = two subspaces x 32 slots
constexpr int STAGES = 4;
constexpr int A_BYTES = TM * KH * 2;       // 16 KB
constexpr int B_BYTES = TN * KH * 2;       // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NQ = 4;                      // column quarters: epilogue thread = one sample x 64 columns
constexpr size_t OFF_COMB = (size_t)STAGES * STAGE_BYTES;          // [2][NQ][128] x (float, int, float)
constexpr size_t OFF_BAR = OFF_COMB + (size_t)2 * NQ * TM * 12;
constexpr size_t SMEM_BYTES = OFF_BAR + (size_t)(2 * STAGES + 4) * 8 + 1024 + 16;
constexpr int NTHREADS = 64 + NQ * 128;     // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue

__host__ __device__ constexpr uint32_t make_idesc() { return (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24); }

// ------------------------------------------------------------------ operand preparation
__device__ __forceinline__ void split16(float v, __half &hi, __half &lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(__fsub_rn(v, __half2float(hi)));
}
// One thread per (sample, subspace): 32 halves of the A row + the subspace's running max ||x_m||^2 (scaled).
__global__ void __launch_bounds__(256) shadow_kernel(const float *vecs, int64_t n, int64_t dim, int G, float scale, __half *x16,
                                                     unsigned int *xn_max_bits) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * G) return;
    const int64_t r = t / G;
    const int g = (int)(t - r * G);
    const float4 a = *reinterpret_cast<const float4 *>(vecs + r * dim + (int64_t)g * 8);
    const float4 b = *reinterpret_cast<const float4 *>(vecs + r * dim + (int64_t)g * 8 + 4);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __align__(16) __half row[32];
    float nn = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float v = __fmul_rn(x[i], scale);
        nn = __fmaf_rn(v, v, nn);
        __half hi, lo;
        split16(v, hi, lo);
        row[i] = hi;
        row[8 + i] = lo;
        row[16 + i] = hi;
        row[24 + i] = __float2half_rn(i < 3 ? 1.0f : 0.0f);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(x16 + t * 32);
    const uint4 *src = reinterpret_cast<const uint4 *>(row);
#pragma unroll
    for (int i = 0; i < 4; i++) dst[i] = src[i];
    atomicMax(xn_max_bits + g, __float_as_uint(nn * 1.0001f));
}
// One thread per (subspace, centroid): its B row inside the pair tile + the subspace's max ||c||^2 (scaled).
__global__ void __launch_bounds__(256) centroid_kernel(const float *cent /*[G][256][8]*/, int G, float scale, __half *c16 /*[G/2][256][64]*/,
                                                       unsigned int *cn_max_bits) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * TN) return;
    const int g = t / TN, k = t - g * TN;
    const float *c = cent + (int64_t)t * 8;
    __align__(16) __half row[32];
    float cn = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float v = __fmul_rn(c[i], scale);
        cn = __fmaf_rn(v, v, cn);
        __half hi, lo;
        split16(v, hi, lo);
        row[i] = hi;
        row[8 + i] = hi;
        row[16 + i] = lo;
        row[24 + i] = __float2half_rn(0.0f);
    }
    const float v = __fmul_rn(-0.5f, cn);
    const __half h = __float2half_rn(v);
    const float r1 = __fsub_rn(v, __half2float(h));
    const __half m = __float2half_rn(r1);
    const __half l = __float2half_rn(__fsub_rn(r1, __half2float(m)));
    row[24] = h;
    row[25] = m;
    row[26] = l;
    uint4 *dst = reinterpret_cast<uint4 *>(c16 + ((int64_t)(g >> 1) * TN + k) * KH + (g & 1) * 32);
    const uint4 *src = reinterpret_cast<const uint4 *>(row);
#pragma unroll
    for (int i = 0; i < 4; i++) dst[i] = src[i];
    atomicMax(cn_max_bits + g, __float_as_uint(cn * 1.0001f));
}

// ------------------------------------------------------------------ GEMM + (max, argmax, second max) + certificate
struct KArgs {
    int64_t n;
    int G;
    const unsigned int *xn_max_bits, *cn_max_bits;
    uint32_t *assign;        // [G][n]
    uint32_t *list;          // uncertified (sample * G + subspace)
    unsigned int *list_count;
};

__global__ void __launch_bounds__(NTHREADS, 1)
pqa_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_c, KArgs A) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * TM;
    const int pairs = A.G >> 1, units = A.G;

    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar0 = s_base + (uint32_t)OFF_BAR;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), NQ * 4);   // one arrival per epilogue warp: a unit is short, 512 arrivals would serialise on the barrier
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
        if (lane == 0) {
            for (int p = 0; p < pairs; p++) {
                const int st = p % STAGES;
                const uint32_t ph = (p / STAGES) & 1;
                mbar_wait(empty_bar(st), ph ^ 1);
                mbar_expect_tx(full_bar(st), STAGE_BYTES);
                tma_load_2d(s_base + st * STAGE_BYTES, &map_x, p * KH, (int)row0, full_bar(st));
                tma_load_2d(s_base + st * STAGE_BYTES + A_BYTES, &map_c, 0, p * TN, full_bar(st));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc();
            for (int p = 0; p < pairs; p++) {
                const int st = p % STAGES;
                const uint32_t ph = (p / STAGES) & 1;
                mbar_wait(full_bar(st), ph);
                tc_fence_after();
                const uint32_t sa = s_base + st * STAGE_BYTES;
                const uint64_t adesc = make_sdesc(sa), bdesc = make_sdesc(sa + A_BYTES);
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int u = 2 * p + j;
                    const int as = u & 1;   // = j
                    const uint32_t aph = (u >> 1) & 1;
                    mbar_wait(tempty_bar(as), aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(as * TN);
                    umma_f16(d_tmem, adesc + (uint64_t)((2 * j) * 2), bdesc + (uint64_t)((2 * j) * 2), idesc, 0u);
                    umma_f16(d_tmem, adesc + (uint64_t)((2 * j + 1) * 2), bdesc + (uint64_t)((2 * j + 1) * 2), idesc, 1u);
                    umma_commit(tfull_bar(as));
                }
                umma_commit(empty_bar(st));
            }
        }
    } else {
        // ===================== epilogue: warps 2..17; thread = one sample x one 64-column quarter =====================
        const int quad = warp & 3;
        const int colq = (warp - 2) >> 2;
        const int r = quad * 32 + lane;
        const int64_t i = row0 + r;
        float *comb_m1 = reinterpret_cast<float *>(smem + OFF_COMB);        // [2][NQ][128]
        float *comb_cnt = comb_m1 + 2 * NQ * TM;
        float *comb_idx = comb_cnt + 2 * NQ * TM;
        const float NEG = -3.0e38f;
        for (int u = 0; u < units; u++) {
            const int as = u & 1;
            const uint32_t aph = (u >> 1) & 1;
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * TN + colq * 64);
            uint32_t v0[32], v1[32];
            tmem_ld32(taddr, v0);
            tmem_ld32(taddr + 32u, v1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(as));   // the accumulator is in registers: the next subspace's MMA may overwrite it
            // pass 1: this quarter's maximum (FMNMX3, ALU pipe).  pass 2 on the FMA pipe — ncu showed the compare / select /
            // integer-add form ALU-pipe-bound (IPC 2.46, FSETP + SEL + VIADD + IADD3 = 57 % of the instructions): the
            // indicator [v > lim] is sat((v - lim) * 2^60) from one FFMA.SAT (0 or 1 for every float32 v except within
            // 2^-60 of the limit, where it is fractional), cnt = sum of the indicators and idx = sum of indicator * column.
            // "Certain" = cnt is exactly 1: the row maximum always contributes exactly 1, every indicator is >= 0, so a
            // total of 1.0 means every other one is exactly 0 — and then idx is the arg-max column.
            const float thr = (__uint_as_float(__ldg(A.xn_max_bits + u)) + __uint_as_float(__ldg(A.cn_max_bits + u))) * (1.0f / 65536.0f);
            float mx = NEG;
#pragma unroll
            for (int j = 0; j < 32; j++) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
            const float H = 1152921504606846976.0f;  // 2^60
            const float nlimH = __fmul_rn(__fsub_rn(thr, mx), H);
            float cnt[2] = {0.0f, 0.0f}, idx[2] = {0.0f, 0.0f};
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const float c0 = __saturatef(__fmaf_rn(__uint_as_float(v0[j]), H, nlimH));
                const float c1 = __saturatef(__fmaf_rn(__uint_as_float(v1[j]), H, nlimH));
                cnt[0] = __fadd_rn(cnt[0], c0);
                cnt[1] = __fadd_rn(cnt[1], c1);
                idx[0] = __fmaf_rn(c0, (float)j, idx[0]);          // column inside the quarter: compile-time constants
                idx[1] = __fmaf_rn(c1, (float)(32 + j), idx[1]);
            }
            idx[0] = __fmaf_rn(__fadd_rn(cnt[0], cnt[1]), (float)(colq * 64), idx[0]);   // exact whenever the count is 1
            const int slot = (as * NQ + colq) * TM + r;
            comb_m1[slot] = mx;
            comb_cnt[slot] = __fadd_rn(cnt[0], cnt[1]);
            comb_idx[slot] = __fadd_rn(idx[0], idx[1]);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (colq == 0) {
                float qm[NQ], qc[NQ], qi[NQ];
                float M1 = NEG;
#pragma unroll
                for (int qq = 0; qq < NQ; qq++) {
                    qm[qq] = comb_m1[(as * NQ + qq) * TM + r];
                    qc[qq] = comb_cnt[(as * NQ + qq) * TM + r];
                    qi[qq] = comb_idx[(as * NQ + qq) * TM + r];
                    M1 = fmaxf(M1, qm[qq]);
                }
                // a quarter whose own maximum is within thr of the row maximum counts its candidates against its own (lower
                // or equal) limit: never fewer than the true number, and exact when it is the only such quarter
                float totalf = 0.0f, If = 0.0f;
#pragma unroll
                for (int qq = 0; qq < NQ; qq++)
                    if (qm[qq] >= M1 - thr) {
                        totalf = __fadd_rn(totalf, qc[qq]);
                        If = qi[qq];
                    }
                const int total = totalf == 1.0f ? 1 : 2;
                const int I1 = (int)If;
                const bool live = i < A.n;
                const bool unsure = live && total != 1;
                if (live) A.assign[(int64_t)u * A.n + i] = (uint32_t)(total == 1 ? I1 : 0);
                const unsigned um = __ballot_sync(0xffffffffu, unsure);
                if (um) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(A.list_count, (unsigned)__popc(um));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (unsure) A.list[base + __popc(um & ((1u << lane) - 1u))] = (uint32_t)(i * A.G + u);
                }
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

// Exact re-evaluation of the listed (sample, subspace) pairs: the reference loop (sequential FMA distance, strict <,
// first wins) over all 256 centroids.
__global__ void __launch_bounds__(256) fallback_kernel(const float *vecs, int64_t n, int64_t dim, int G, const float *cent,
                                                       const uint32_t *list, const unsigned int *list_count, uint32_t *assign) {
    const unsigned total = *list_count;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const uint32_t id = list[e];
        const int64_t i = id / (uint32_t)G;
        const int g = (int)(id - (uint32_t)i * (uint32_t)G);
        const float4 a = *reinterpret_cast<const float4 *>(vecs + i * dim + (int64_t)g * 8);
        const float4 b = *reinterpret_cast<const float4 *>(vecs + i * dim + (int64_t)g * 8 + 4);
        const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const float4 *c4 = reinterpret_cast<const float4 *>(cent + (int64_t)g * TN * 8);
        float best = 3.402823466e+38f;
        int idx = 0;
        for (int c = 0; c < TN; c++) {
            const float4 p = __ldg(c4 + 2 * c), q = __ldg(c4 + 2 * c + 1);
            const float y[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
            float tot = 0.0f;
#pragma unroll
            for (int d = 0; d < 8; d++) {
                const float df = __fsub_rn(x[d], y[d]);
                tot = __fmaf_rn(df, df, tot);
            }
            if (tot < best) {
                best = tot;
                idx = c;
            }
        }
        assign[(int64_t)g * n + i] = (uint32_t)idx;
    }
}

__global__ void absmax_kernel(const float *mins, const float *maxs, int64_t dim, float *out /*[2]: max|x|, non-finite flag*/) {
    float m = 0.0f;
    bool bad = false;
    for (int64_t d = threadIdx.x; d < dim; d += blockDim.x) {
        const float a = mins[d], b = maxs[d];
        if (!(fabsf(a) < 3.0e38f) || !(fabsf(b) < 3.0e38f)) bad = true;
        m = fmaxf(m, fmaxf(fabsf(a), fabsf(b)));
    }
    __shared__ float sm[256];
    __shared__ int sb;
    if (threadIdx.x == 0) sb = 0;
    __syncthreads();
    sm[threadIdx.x] = m;
    if (bad) sb = 1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] = fmaxf(sm[threadIdx.x], sm[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = sm[0];
        out[1] = sb ? 1.0f : 0.0f;
    }
}

// ------------------------------------------------------------------ host
static bool enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("VECGO_PQ_ASSIGN_TC");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0 && tc::enabled();
}

bool Assigner::supported(int64_t n, int64_t dim, int G, int K, int ds) {
    return enabled() && ds == 8 && K == TN && (G & 1) == 0 && dim >= (int64_t)G * 8 && dim % 4 == 0 && n >= 4096 && n * (int64_t)G < (1ll << 32) &&
           n * (int64_t)G * 64 <= (32ll << 30);  // the fp16 shadow is twice the training set
}

vg_status Assigner::prepare(const float *d_vecs, int64_t n_, int64_t dim_, int G_, cudaStream_t st) {
    ready = false;
    n = n_;
    dim = dim_;
    G = G_;
    vecs = d_vecs;
    if ((reinterpret_cast<uintptr_t>(d_vecs) & 15) != 0) return VG_OK;
    DevBuf mm, am;
    const int64_t cols = (int64_t)G * 8;  // the G subspaces trained here: a column slice when the subspaces are split across GPUs
    VG_TRY(mm.alloc((size_t)cols * 8));
    VG_TRY(am.alloc(8));
    VG_TRY(dev_minmax_strided(d_vecs, n, cols, dim, mm.as<float>(), mm.as<float>() + cols, st));
    absmax_kernel<<<1, 256, 0, st>>>(mm.as<float>(), mm.as<float>() + cols, cols, am.as<float>());
    VG_LAUNCHED();
    float h[2] = {0.f, 0.f};
    VG_CUDA(cudaMemcpyAsync(h, am.p, 8, cudaMemcpyDeviceToHost, st));
    VG_CUDA(cudaStreamSynchronize(st));
    if (!(h[0] > 0.0f) || h[1] != 0.0f || !(h[0] < 1.0e30f) || h[0] < 1.0e-30f) return VG_OK;  // constant / non-finite / extreme data: exact path
    int ex = 0;
    frexpf(h[0], &ex);           // h[0] = f * 2^ex, f in [0.5, 1)
    scale = ldexpf(1.0f, 4 - ex);  // max|x| * scale in [8, 16)
    // from the stream-ordered pool: a cold 6 GB cudaMalloc cost 2.3 s (measured), the pool 0.04 s
    VG_TRY(x16.alloc((size_t)n * G * 32 * 2));
    VG_TRY(c16.alloc((size_t)(G / 2) * TN * KH * 2));
    VG_TRY(maxbits.alloc((size_t)2 * G * 4 + 16));
    VG_TRY(list.alloc((size_t)n * G * 4));
    VG_CUDA(cudaMemsetAsync(maxbits.p, 0, maxbits.bytes, st));
    const int64_t t = n * G;
    shadow_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(d_vecs, n, dim, G, scale, x16.as<__half>(), maxbits.as<unsigned int>());
    VG_LAUNCHED();
    ready = true;
    return VG_OK;
}

vg_status Assigner::assign(const float *d_cent, uint32_t *d_assign, cudaStream_t st) {
    unsigned int *xn_bits = maxbits.as<unsigned int>(), *cn_bits = xn_bits + G, *count = cn_bits + G;
    VG_CUDA(cudaMemsetAsync(cn_bits, 0, (size_t)G * 4 + 16, st));
    centroid_kernel<<<(unsigned)((G * TN + 255) / 256), 256, 0, st>>>(d_cent, G, scale, c16.as<__half>(), cn_bits);
    VG_LAUNCHED();
    CUtensorMap mx, mc;
    VG_TRY(tc::tensor_map_2d(&mx, true, x16.p, n, (int64_t)G * 32, (int64_t)G * 32, KH, TM));
    VG_TRY(tc::tensor_map_2d(&mc, true, c16.p, (int64_t)(G / 2) * TN, KH, KH, KH, TN));
    KArgs a{};
    a.n = n;
    a.G = G;
    a.xn_max_bits = xn_bits;
    a.cn_max_bits = cn_bits;
    a.assign = d_assign;
    a.list = list.as<uint32_t>();
    a.list_count = count;
    VG_CUDA(cudaFuncSetAttribute(pqa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    pqa_kernel<<<(unsigned)((n + TM - 1) / TM), NTHREADS, SMEM_BYTES, st>>>(mx, mc, a);
    VG_LAUNCHED();
    fallback_kernel<<<sm_count() * 8, 256, 0, st>>>(vecs, n, dim, G, d_cent, list.as<uint32_t>(), count, d_assign);
    VG_LAUNCHED();
    return VG_OK;
}

static std::atomic<uint64_t> g_pairs{0}, g_fallback_pairs{0};
void stats(uint64_t *pairs, uint64_t *fallback_pairs) {
    if (pairs) *pairs = g_pairs.load();
    if (fallback_pairs) *fallback_pairs = g_fallback_pairs.load();
}
vg_status Assigner::account(cudaStream_t st) {
    uint64_t h = 0;
    VG_TRY(last_fallbacks(&h, st));
    g_pairs.fetch_add((uint64_t)n * (uint64_t)G);
    g_fallback_pairs.fetch_add(h);
    return VG_OK;
}
vg_status Assigner::last_fallbacks(uint64_t *pairs, cudaStream_t st) {
    unsigned int h = 0;
    VG_CUDA(cudaMemcpyAsync(&h, maxbits.as<unsigned int>() + 2 * G, 4, cudaMemcpyDeviceToHost, st));
    VG_CUDA(cudaStreamSynchronize(st));
    *pairs = h;
    return VG_OK;
}

}  // namespace pqa
}  // namespace vg
