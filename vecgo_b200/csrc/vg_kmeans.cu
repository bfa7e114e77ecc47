// vg_kmeans.cu — Lloyd k-means (internal/kmeans/kmeans.go) and the
// ProductQuantizer's private per-subspace k-means + k-means++ init + int8
// codebook quantisation (internal/quantization/pq.go:68-143,275-433).
//
// Exactness plan.  Assignment is an argmin with "strict <, first wins", i.e.
// the smallest (distance, centroid id) — distances are evaluated in the
// reference's summation order.  The centroid update is a SEQUENTIAL float32
// sum in sample order per (cluster, dim); we keep that order by building the
// per-cluster member lists with a stable partition and letting one thread walk
// each (cluster, dim) chain — thousands of chains run in parallel, each chain
// is order-exact.  k-means++ draws from an explicit seed (the reference uses
// Go's unseeded global RNG and is not reproducible even against itself); its
// float32 running sums are sequential chains too and are kept sequential.
#include "vg_flat_tc.cuh"
#include "vg_kmeans.cuh"
#include "vg_pq_assign_tc.cuh"

#include <cstdlib>
#include <vector>

#include "vg_scan.cuh"

namespace vg {

// ---------------------------------------------------------------- RNG
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline uint64_t rng_u64(uint64_t seed, uint64_t a, uint64_t b) {
    return splitmix64(splitmix64(seed ^ (a * 0xD6E8FEB86659FD93ull)) ^ b);
}
__host__ __device__ inline int64_t rng_intn(uint64_t seed, uint64_t a, uint64_t b, int64_t n) {
    return (int64_t)(rng_u64(seed, a, b) % (uint64_t)n);
}
__host__ __device__ inline float rng_f32(uint64_t seed, uint64_t a, uint64_t b) {
    return (float)(rng_u64(seed, a, b) >> 40) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------- helpers
// simd.SquaredL2 (floats_avx512.c:69-129) evaluated by ONE thread, same order.
__device__ float sql2_pair_thread(const float *a, const float *b, int n) {
    float tot = 0.0f;
    const int epochs = n >> 6;
    if (epochs > 0) {
        float acc[64];
        for (int i = 0; i < 64; i++) acc[i] = 0.0f;
        for (int e = 0; e < epochs; e++)
            for (int i = 0; i < 64; i++) {
                const float d = __fsub_rn(a[e * 64 + i], b[e * 64 + i]);
                acc[i] = __fmaf_rn(d, d, acc[i]);
            }
        float c[16];
        for (int l = 0; l < 16; l++) c[l] = __fadd_rn(__fadd_rn(acc[l], acc[16 + l]), __fadd_rn(acc[32 + l], acc[48 + l]));
        for (int l = 0; l < 8; l++) c[l] = __fadd_rn(c[l], c[l + 8]);
        for (int l = 0; l < 4; l++) c[l] = __fadd_rn(c[l], c[l + 4]);
        tot = __fadd_rn(__fadd_rn(c[0], c[2]), __fadd_rn(c[1], c[3]));
    }
    for (int i = epochs * 64; i < n; i++) {
        const float d = __fsub_rn(a[i], b[i]);
        tot = __fmaf_rn(d, d, tot);
    }
    return tot;
}

// ---------------------------------------------------------------- assignment
// Generic: top-1 scan over the centroid table (any dim, pair or batch order).
static vg_status assign_generic(const float *d_vecs, int64_t n, int64_t stride, int64_t dim, const float *d_cent, int64_t k,
                                int variant, int is_dot, uint32_t *d_assign, float *d_score, int32_t *d_cnt, cudaStream_t st) {
    CodecParams cp;
    cp.codec = VG_CODEC_F32;
    cp.variant = variant;
    cp.dim = dim;
    cp.vectors = d_cent;
    ScanArgs a;
    a.queries = d_vecs;
    a.q_stride = stride;
    a.nq = n;
    a.rows = k;
    a.k = 1;
    a.descending = is_dot;
    a.is_dot = is_dot;
    a.out_rows = d_assign;
    a.out_scores = d_score;
    a.out_counts = d_cnt;
    // Nearest-centroid assignment is a dense [n x dim] . [dim x k] contraction: run it through the tcgen05 filter
    // (vg_flat_tc.cu: TF32 GEMM -> arg-min rows of the best groups -> exact re-check in simd order -> certificate)
    // whenever the Batch kernel's order coincides with the pair order the exact stage uses (dim % 64 < 16).
    const bool same_order = !(variant & VG_VAR_BATCH) || (dim % 64) < 16;
    if (tc::enabled() && same_order && tc::supported_assign(dim, k, n, stride) && (reinterpret_cast<uintptr_t>(d_vecs) & 15) == 0) {
        DevBuf xn, xmax;
        VG_TRY(xn.alloc((size_t)k * 4));
        VG_TRY(xmax.alloc(16));
        VG_CUDA(cudaMemsetAsync(xmax.p, 0, 16, st));
        VG_TRY(tc::sqnorms(d_cent, k, dim, dim, xn.as<float>(), xmax.as<unsigned int>(), st));
        tc::SearchIO io;
        io.d_queries = d_vecs;
        io.q_stride = stride;
        io.nq = n;
        io.d_vectors = d_cent;
        io.rows = k;
        io.dim = dim;
        io.d_xn = xn.as<float>();
        io.d_xmax_bits = xmax.as<unsigned int>();
        io.k = 1;
        io.is_dot = is_dot;
        io.d_rows = d_assign;
        io.d_scores = d_score;
        io.d_counts = d_cnt;
        std::vector<int32_t> bad;
        VG_TRY(tc::search(io, tc::candidates_for_assign(k), bad, st));
        return scan_topk_subset(cp, a, bad, st);  // samples whose certificate failed: exact scan
    }
    return scan_topk(cp, a, st);
}

// PQ subspaces with ds < 64: simd.SquaredL2 runs entirely in its FMA scalar tail
// (sum = fma(d, d, sum), floats_avx512.s:316-323).  One thread per (sample, subspace);
// the subspace's K x DS centroid table sits in shared memory (broadcast reads).
template <int DS>
__global__ void __launch_bounds__(256) pq_assign_small_kernel(const float *vecs, int64_t n, int64_t dim, int K,
                                                              const float *cent /*[G][K][DS]*/, uint32_t *assign /*[G][n]*/) {
    extern __shared__ __align__(16) float sc[];
    const int g = blockIdx.y;
    for (int i = threadIdx.x; i < K * DS; i += blockDim.x) sc[i] = cent[(int64_t)g * K * DS + i];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float x[DS];
#pragma unroll
    for (int i = 0; i < DS; i++) x[i] = vecs[r * dim + (int64_t)g * DS + i];
    float best = 3.402823466e+38f;
    int idx = 0;
    for (int c = 0; c < K; c++) {
        float tot = 0.0f;
#pragma unroll
        for (int i = 0; i < DS; i++) {
            const float d = __fsub_rn(x[i], sc[c * DS + i]);
            tot = __fmaf_rn(d, d, tot);
        }
        if (tot < best) {
            best = tot;
            idx = c;
        }
    }
    assign[(int64_t)g * n + r] = (uint32_t)idx;
}

// Same assignment with two CENTROIDS per packed instruction (sm_100a FADD2 / FFMA2, each half an IEEE round-to-nearest
// op): the table is staged as (-c_{2j,i}, -c_{2j+1,i}) pairs, d = x + (-c) equals x - c exactly, and the two sums come
// out bit-identical to the scalar chain.  K must be even.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t km_pk2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void km_unpk2(f32x2_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t km_add2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t km_fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
template <int DS>
__global__ void __launch_bounds__(256) pq_assign_small2_kernel(const float *vecs, int64_t n, int64_t dim, int K,
                                                               const float *cent /*[G][K][DS]*/, uint32_t *assign /*[G][n]*/) {
    extern __shared__ __align__(16) float sc[];  // [K/2][DS] pairs (-c_even, -c_odd)
    const int g = blockIdx.y;
    f32x2_t *sp = reinterpret_cast<f32x2_t *>(sc);
    for (int i = threadIdx.x; i < (K / 2) * DS; i += blockDim.x) {
        const int j = i / DS, d = i - j * DS;
        const float *c0 = cent + ((int64_t)g * K + 2 * j) * DS;
        sp[i] = km_pk2(-c0[d], -c0[DS + d]);
    }
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    f32x2_t x2[DS];
#pragma unroll
    for (int i = 0; i < DS; i++) {
        const float x = vecs[r * dim + (int64_t)g * DS + i];
        x2[i] = km_pk2(x, x);
    }
    float best = 3.402823466e+38f;
    int idx = 0;
    for (int j = 0; j < K / 2; j++) {
        f32x2_t tot = 0ull;
#pragma unroll
        for (int i = 0; i < DS; i++) {
            const f32x2_t d = km_add2(x2[i], sp[j * DS + i]);
            tot = km_fma2(d, d, tot);
        }
        float t0, t1;
        km_unpk2(tot, t0, t1);
        if (t0 < best) {
            best = t0;
            idx = 2 * j;
        }
        if (t1 < best) {
            best = t1;
            idx = 2 * j + 1;
        }
    }
    assign[(int64_t)g * n + r] = (uint32_t)idx;
}

// changed[g] |= (new != old); old = new  (assignments start at zero, pq.go:349)
__global__ void __launch_bounds__(256) diff_assign_kernel(const uint32_t *newa, int32_t *olda, int64_t n, int G, const int *active,
                                                          int *changed) {
    const int g = blockIdx.y;
    if (!active[g]) return;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool diff = false;
    if ((n & 3) == 0) {  // four samples per thread, 16-byte accesses (rows of n samples stay 16-byte aligned)
        if (4 * t < n) {
            const int4 v = *reinterpret_cast<const int4 *>(newa + (int64_t)g * n + 4 * t);
            int4 *o = reinterpret_cast<int4 *>(olda + (int64_t)g * n + 4 * t);
            const int4 w = *o;
            diff = v.x != w.x || v.y != w.y || v.z != w.z || v.w != w.w;
            if (diff) *o = v;
        }
    } else {
        for (int64_t i = 4 * t; i < n && i < 4 * t + 4; i++) {
            const int32_t v = (int32_t)newa[(int64_t)g * n + i];
            if (olda[(int64_t)g * n + i] != v) {
                olda[(int64_t)g * n + i] = v;
                diff = true;
            }
        }
    }
    if (__any_sync(0xffffffffu, diff) && (threadIdx.x & 31) == 0) changed[g] = 1;
}
// after the assignment pass: a group whose pass changed nothing stops (break before update)
__global__ void settle_kernel(int *active, int *changed, int *iters, int G) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    if (active[g]) {
        if (!changed[g]) active[g] = 0;
        else iters[g] += 1;
    }
    changed[g] = 0;
}

// ---------------------------------------------------------------- stable partition
static const int kSB = 1024;  // samples per partition block
__global__ void __launch_bounds__(256) part_count_kernel(const int32_t *assign, int64_t n, int K, int64_t B, const int *active,
                                                         int *counts /*[G][K][B]*/) {
    extern __shared__ int hist[];
    const int g = blockIdx.y;
    if (!active[g]) return;
    const int64_t b = blockIdx.x;
    for (int c = threadIdx.x; c < K; c += blockDim.x) hist[c] = 0;
    __syncthreads();
    const int64_t i0 = b * kSB;
    for (int i = threadIdx.x; i < kSB; i += blockDim.x)
        if (i0 + i < n) atomicAdd(&hist[assign[(int64_t)g * n + i0 + i]], 1);
    __syncthreads();
    for (int c = threadIdx.x; c < K; c += blockDim.x) counts[((int64_t)g * K + c) * B + b] = hist[c];
}
__global__ void __launch_bounds__(256) part_scan_kernel(int *counts, int K, int64_t B, int G, const int *active, int *totals) {
    // one warp per (group, cluster): exclusive scan of its B block counts, 32 at a time (coalesced)
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= (int64_t)G * K) return;
    if (!active[t / K]) return;
    int run = 0;
    int *p = counts + t * B;
    for (int64_t b0 = 0; b0 < B; b0 += 32) {
        const int64_t b = b0 + lane;
        const int v = b < B ? p[b] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (b < B) p[b] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) totals[t] = run;
}
__global__ void part_base_kernel(const int *totals, int K, int G, const int *active, int64_t n, int64_t *base) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G || !active[g]) return;
    int64_t run = (int64_t)g * n;
    for (int c = 0; c < K; c++) {
        base[(int64_t)g * K + c] = run;
        run += totals[(int64_t)g * K + c];
    }
}
// Stable scatter of one block of kSB samples: each of the 8 warps owns 128 consecutive samples; per-warp histograms
// turn into per-(warp, cluster) start positions, and inside a warp the rank of a sample among the earlier samples of
// the same cluster comes from __match_any_sync — O(samples) work, members of a cluster stay in sample order.
constexpr int kScatterMaxK = 2048;  // 8 warps x K counters in shared memory
__global__ void __launch_bounds__(256) part_scatter_kernel(const int32_t *assign, int64_t n, int K, int64_t B, const int *active,
                                                           const int *counts, const int64_t *base, uint32_t *members) {
    extern __shared__ int sc_smem[];
    int *a = sc_smem;         // [kSB]
    int *hw = sc_smem + kSB;  // [8][K]
    const int g = blockIdx.y;
    if (!active[g]) return;
    const int64_t b = blockIdx.x;
    const int64_t i0 = b * kSB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cnt = (int)((n - i0 < kSB) ? (n - i0) : kSB);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) a[i] = assign[(int64_t)g * n + i0 + i];
    for (int i = threadIdx.x; i < 8 * K; i += blockDim.x) hw[i] = 0;
    __syncthreads();
    constexpr int PER_WARP = kSB / 8;
    for (int s_ = 0; s_ < PER_WARP; s_ += 32) {
        const int i = warp * PER_WARP + s_ + lane;
        if (i < cnt) atomicAdd(&hw[warp * K + a[i]], 1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < K; c += blockDim.x) {
        // position inside the group's member list (fits an int: n < 2^31)
        int run = (int)(base[(int64_t)g * K + c] - (int64_t)g * n) + counts[((int64_t)g * K + c) * B + b];
        for (int w = 0; w < 8; w++) {
            const int t = hw[w * K + c];
            hw[w * K + c] = run;
            run += t;
        }
    }
    __syncthreads();
    uint32_t *mem = members + (int64_t)g * n;
    for (int s_ = 0; s_ < PER_WARP; s_ += 32) {
        const int i = warp * PER_WARP + s_ + lane;
        const bool valid = i < cnt;
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int key = a[i];
            const unsigned same = __match_any_sync(vm, key);
            const int rank = __popc(same & ((1u << lane) - 1u));
            const int pos = hw[warp * K + key] + rank;
            mem[pos] = (uint32_t)(i0 + i);
            __syncwarp(vm);
            if (rank == 0) hw[warp * K + key] += __popc(same);
        }
        __syncwarp();
    }
}
// Any K: one thread per cluster walks the block (O(K x samples)).
__global__ void __launch_bounds__(256) part_scatter_anyk_kernel(const int32_t *assign, int64_t n, int K, int64_t B, const int *active,
                                                                const int *counts, const int64_t *base, uint32_t *members) {
    __shared__ int a[kSB];
    const int g = blockIdx.y;
    if (!active[g]) return;
    const int64_t b = blockIdx.x;
    const int64_t i0 = b * kSB;
    int cnt = (int)((n - i0 < kSB) ? (n - i0) : kSB);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) a[i] = assign[(int64_t)g * n + i0 + i];
    __syncthreads();
    for (int c = threadIdx.x; c < K; c += blockDim.x) {
        int64_t pos = base[(int64_t)g * K + c] + counts[((int64_t)g * K + c) * B + b];
        for (int i = 0; i < cnt; i++)
            if (a[i] == c) members[pos++] = (uint32_t)(i0 + i);
    }
}
// One thread per (group, cluster, dim): sequential float32 sum over the members in
// sample order, then  sum/count (pq.go:388-413)  or  sum*(1/count) (kmeans.go:110-134).
__global__ void __launch_bounds__(256) centroid_update_kernel(const float *vecs, int64_t n, int64_t stride, int ds, int K, int G,
                                                              const int *active, const int *totals, const int64_t *base,
                                                              const uint32_t *members, int reciprocal, uint64_t seed,
                                                              uint64_t tag_xor, int tag_is_group, int g_base, const int *iters,
                                                              float *cent /*[G][K][ds]*/) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)G * K * ds) return;
    const int j = (int)(t % ds);
    const int64_t gc = t / ds;
    const int g = (int)(gc / K), c = (int)(gc % K);
    if (!active[g]) return;
    const int cntc = totals[gc];
    const float *col = vecs + (int64_t)g * ds + j;
    if (cntc > 0) {
        const uint32_t *mem = members + base[gc];
        float sum = 0.0f;
        int i = 0;
        for (; i + 4 <= cntc; i += 4) {
            const float v0 = col[(int64_t)mem[i] * stride], v1 = col[(int64_t)mem[i + 1] * stride];
            const float v2 = col[(int64_t)mem[i + 2] * stride], v3 = col[(int64_t)mem[i + 3] * stride];
            sum = __fadd_rn(sum, v0);
            sum = __fadd_rn(sum, v1);
            sum = __fadd_rn(sum, v2);
            sum = __fadd_rn(sum, v3);
        }
        for (; i < cntc; i++) sum = __fadd_rn(sum, col[(int64_t)mem[i] * stride]);
        if (reciprocal) cent[t] = __fmul_rn(sum, __fdiv_rn(1.0f, (float)cntc));
        else cent[t] = __fdiv_rn(sum, (float)cntc);
    } else {
        // empty cluster: re-seed from a random sample (explicit RNG; iteration = completed updates)
        const uint64_t tag = (tag_is_group ? (uint64_t)(g_base + g) : 0ull) ^ tag_xor;
        const int64_t idx = rng_intn(seed, tag, (uint64_t)((int64_t)(iters[g] - 1) * K + c), n);
        cent[t] = col[idx * stride];
    }
}

struct Lloyd {
    int G, K, ds;
    int g_base = 0;   // index of the first subspace trained here inside the whole quantizer (selects the RNG streams)
    int64_t n, stride, B;
    DevBuf assign_new, assign_old, counts, totals, base, members, active, changed, iters, score, cnt;
    vg_status init(int G_, int K_, int ds_, int64_t n_, int64_t stride_, cudaStream_t st) {
        G = G_;
        K = K_;
        ds = ds_;
        n = n_;
        stride = stride_;
        B = (n + kSB - 1) / kSB;
        VG_TRY(assign_new.alloc((size_t)G * n * 4));
        VG_TRY(assign_old.alloc((size_t)G * n * 4));
        VG_TRY(counts.alloc((size_t)G * K * B * 4));
        VG_TRY(totals.alloc((size_t)G * K * 4));
        VG_TRY(base.alloc((size_t)G * K * 8));
        VG_TRY(members.alloc((size_t)G * n * 4));
        VG_TRY(active.alloc((size_t)G * 4));
        VG_TRY(changed.alloc((size_t)G * 4));
        VG_TRY(iters.alloc((size_t)G * 4));
        VG_CUDA(cudaMemsetAsync(assign_old.p, 0, assign_old.bytes, st));
        VG_CUDA(cudaMemsetAsync(changed.p, 0, changed.bytes, st));
        VG_CUDA(cudaMemsetAsync(iters.p, 0, iters.bytes, st));
        std::vector<int> ones((size_t)G, 1);
        VG_CUDA(cudaMemcpyAsync(active.p, ones.data(), (size_t)G * 4, cudaMemcpyHostToDevice, st));
        VG_CUDA(cudaStreamSynchronize(st));
        return VG_OK;
    }
    // assign_new must hold this pass's assignments.  Returns (via host) whether any group is still active.
    vg_status step(const float *d_vecs, float *d_cent, int reciprocal, uint64_t seed, uint64_t tag_xor, int tag_is_group,
                   bool *any_active, cudaStream_t st) {
        dim3 gd((unsigned)(((n + 3) / 4 + 255) / 256), (unsigned)G);
        diff_assign_kernel<<<gd, 256, 0, st>>>(assign_new.as<uint32_t>(), assign_old.as<int32_t>(), n, G, active.as<int>(),
                                               changed.as<int>());
        VG_LAUNCHED();
        settle_kernel<<<(G + 63) / 64, 64, 0, st>>>(active.as<int>(), changed.as<int>(), iters.as<int>(), G);
        VG_LAUNCHED();
        std::vector<int> h((size_t)G);
        VG_CUDA(cudaMemcpyAsync(h.data(), active.p, (size_t)G * 4, cudaMemcpyDeviceToHost, st));
        VG_CUDA(cudaStreamSynchronize(st));
        bool any = false;
        for (int v : h) any = any || v;
        *any_active = any;
        if (!any) return VG_OK;
        dim3 gb((unsigned)B, (unsigned)G);
        part_count_kernel<<<gb, 256, (size_t)K * 4, st>>>(assign_old.as<int32_t>(), n, K, B, active.as<int>(), counts.as<int>());
        VG_LAUNCHED();
        part_scan_kernel<<<(unsigned)(((int64_t)G * K * 32 + 255) / 256), 256, 0, st>>>(counts.as<int>(), K, B, G, active.as<int>(),
                                                                                  totals.as<int>());
        VG_LAUNCHED();
        part_base_kernel<<<(G + 63) / 64, 64, 0, st>>>(totals.as<int>(), K, G, active.as<int>(), n, base.as<int64_t>());
        VG_LAUNCHED();
        if (K <= kScatterMaxK) {
            const size_t ssm = ((size_t)kSB + 8 * (size_t)K) * 4;
            if (ssm > 48 * 1024) VG_CUDA(cudaFuncSetAttribute(part_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
            part_scatter_kernel<<<gb, 256, ssm, st>>>(assign_old.as<int32_t>(), n, K, B, active.as<int>(), counts.as<int>(),
                                                      base.as<int64_t>(), members.as<uint32_t>());
        } else {
            part_scatter_anyk_kernel<<<gb, 256, 0, st>>>(assign_old.as<int32_t>(), n, K, B, active.as<int>(), counts.as<int>(),
                                                         base.as<int64_t>(), members.as<uint32_t>());
        }
        VG_LAUNCHED();
        const int64_t chains = (int64_t)G * K * ds;
        centroid_update_kernel<<<(unsigned)((chains + 255) / 256), 256, 0, st>>>(
            d_vecs, n, stride, ds, K, G, active.as<int>(), totals.as<int>(), base.as<int64_t>(), members.as<uint32_t>(), reciprocal,
            seed, tag_xor, tag_is_group, g_base, iters.as<int>(), d_cent);
        VG_LAUNCHED();
        return VG_OK;
    }
};

// ---------------------------------------------------------------- public device API
vg_status dev_find_closest(const float *d_queries, int64_t nq, int64_t dim, const float *d_centroids, int64_t k, int64_t np,
                           int metric, int32_t *d_out, cudaStream_t st) {
    if (np > k) np = k;
    DevBuf sc, cnt;
    VG_TRY(sc.alloc((size_t)nq * np * 4));
    VG_TRY(cnt.alloc((size_t)nq * 4));
    if (np == 1) {  // AssignPartition: nearest centroid (tensor-core filter for large batches)
        VG_TRY(assign_generic(d_queries, nq, dim, dim, d_centroids, k, VG_VAR_BATCH, metric != VG_METRIC_L2,
                              reinterpret_cast<uint32_t *>(d_out), sc.as<float>(), cnt.as<int32_t>(), st));
        VG_CUDA(cudaStreamSynchronize(st));
        return VG_OK;
    }
    CodecParams cp;
    cp.codec = VG_CODEC_F32;
    cp.variant = VG_VAR_BATCH;
    cp.dim = dim;
    cp.vectors = d_centroids;
    ScanArgs a;
    a.queries = d_queries;
    a.nq = nq;
    a.rows = k;
    a.k = (int)np;
    a.descending = metric != VG_METRIC_L2;
    a.is_dot = metric != VG_METRIC_L2;
    a.out_rows = reinterpret_cast<uint32_t *>(d_out);
    a.out_scores = sc.as<float>();
    a.out_counts = cnt.as<int32_t>();
    VG_TRY(scan_topk(cp, a, st));
    VG_CUDA(cudaStreamSynchronize(st));
    return VG_OK;
}

// ---------------------------------------------------------------- k-means++ (pq.go:281-345)
__global__ void __launch_bounds__(256) gather_centroid_kernel(const float *vecs, int64_t stride, int ds, int K, int G, int c,
                                                              const int64_t *chosen, float *cent) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G * ds) return;
    const int g = t / ds, j = t % ds;
    cent[((int64_t)g * K + c) * ds + j] = vecs[chosen[g] * stride + (int64_t)g * ds + j];
}
// minDistSq update (pq.go:321-330).  One CTA = PP_ROWS samples x all subspaces: consecutive threads read consecutive
// subspaces of the same sample (whole rows, coalesced), the distances go through a shared-memory transpose and the
// min-update of mind[g][i0 .. i0+PP_ROWS) is written as contiguous segments.
constexpr int PP_ROWS = 64;
// simd.SquaredL2 for a subspace shorter than 64 dims = the kernel's FMA scalar tail (floats_avx512.s:316-323):
// sum = fma(d, d, sum) in order.  DS known at compile time: 16-byte loads, fully unrolled chain.
template <int DS>
__device__ __forceinline__ float sql2_tail_fixed(const float *a, const float *b) {
    float tot = 0.0f;
    if constexpr (DS % 4 == 0) {
#pragma unroll
        for (int i = 0; i < DS; i += 4) {
            const float4 x = *reinterpret_cast<const float4 *>(a + i), y = __ldg(reinterpret_cast<const float4 *>(b + i));
            float d = __fsub_rn(x.x, y.x);
            tot = __fmaf_rn(d, d, tot);
            d = __fsub_rn(x.y, y.y);
            tot = __fmaf_rn(d, d, tot);
            d = __fsub_rn(x.z, y.z);
            tot = __fmaf_rn(d, d, tot);
            d = __fsub_rn(x.w, y.w);
            tot = __fmaf_rn(d, d, tot);
        }
    } else {
#pragma unroll
        for (int i = 0; i < DS; i++) {
            const float d = __fsub_rn(a[i], b[i]);
            tot = __fmaf_rn(d, d, tot);
        }
    }
    return tot;
}
template <int DS>  // DS = 0: any subspace length (sql2_pair_thread)
__global__ void __launch_bounds__(256) pp_dist_kernel(const float *vecs, int64_t n, int64_t stride, int ds, int K, int G, int c,
                                                      const float *cent, const int *zero, float *mind /*[G][n]*/) {
    extern __shared__ float sd[];  // [G][PP_ROWS + 1]
    const int64_t i0 = (int64_t)blockIdx.x * PP_ROWS;
    const int pairs = PP_ROWS * G;
    if constexpr (DS >= 8) {
        // DS / 4 neighbouring lanes share one (sample, subspace) pair, one 16-byte load each: a warp reads 512 contiguous
        // bytes of a sample row per instruction (a lane per pair read 32-byte-strided halves of every sector: ncu showed
        // 16 of 32 bytes per sector used and the L1/L2 path 80 % busy).  The FMA chain keeps its order: lane j continues
        // from the running sum of lane j - 1 (shuffle).
        constexpr int L = DS / 4;
        const int lane = threadIdx.x & 31, sub = lane & (L - 1);
        const int items = pairs * L;
        for (int p0 = 0; p0 < items; p0 += blockDim.x) {
            const int it = p0 + threadIdx.x;
            const int p = it / L;
            const int r = p / G, g = p - r * G;
            const int64_t i = i0 + r;
            const bool live = it < items && i < n;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
            if (live) {
                x = *reinterpret_cast<const float4 *>(vecs + i * stride + (int64_t)g * DS + sub * 4);
                y = __ldg(reinterpret_cast<const float4 *>(cent + ((int64_t)g * K + c) * DS + sub * 4));
            }
            float tot = 0.0f;
#pragma unroll
            for (int s_ = 0; s_ < L; s_++) {
                if (s_ > 0) tot = __shfl_sync(0xffffffffu, tot, (lane & ~(L - 1)) + s_ - 1);
                if (sub == s_) {
                    float d = __fsub_rn(x.x, y.x);
                    tot = __fmaf_rn(d, d, tot);
                    d = __fsub_rn(x.y, y.y);
                    tot = __fmaf_rn(d, d, tot);
                    d = __fsub_rn(x.z, y.z);
                    tot = __fmaf_rn(d, d, tot);
                    d = __fsub_rn(x.w, y.w);
                    tot = __fmaf_rn(d, d, tot);
                }
            }
            if (it < items && sub == L - 1) sd[g * (PP_ROWS + 1) + r] = live ? tot : 0.0f;
        }
    } else {
        for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
            const int r = p / G, g = p - r * G;
            const int64_t i = i0 + r;
            float d = 0.0f;
            if (i < n) {
                const float *a = vecs + i * stride + (int64_t)g * ds, *b = cent + ((int64_t)g * K + c) * ds;
                if constexpr (DS > 0) d = sql2_tail_fixed<DS>(a, b);
                else d = sql2_pair_thread(a, b, ds);
            }
            sd[g * (PP_ROWS + 1) + r] = d;
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
        const int g = p / PP_ROWS, r = p - g * PP_ROWS;
        const int64_t i = i0 + r;
        if (i >= n) continue;
        if (c > 0 && zero[g]) continue;  // `sum == 0` branch: no distance update (pq.go:306-310)
        const float d = sd[g * (PP_ROWS + 1) + r];
        float *m = mind + (int64_t)g * n + i;
        if (c == 0 || d < *m) *m = d;
    }
}
// One warp per group.  The float32 running sum of pq.go:299-303,327-329 is a sequential chain; the cumulative search
// of pq.go:314-320 walks the SAME chain, so one pass is enough: the running sum at the start of every PP_TILE-element
// tile is kept as a checkpoint, the target's tile is the last one whose checkpoint is below the target (the chain is
// non-decreasing: fl(s + d) >= s for d >= 0) and only that tile is walked again.  All 32 lanes evaluate the identical
// chain from broadcast 16-byte shared-memory reads while cp.async streams the next tile in, so the cost per element is
// the dependent-FADD latency.
constexpr int PP_TILE = 2048;
__device__ __forceinline__ void pp_load_tile(const float *m, int64_t n, int64_t t, float *dst, int lane) {
    const int64_t b = t * PP_TILE;
    for (int j = lane; j < PP_TILE; j += 32) {
        if (b + j < n) {
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst + j);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(m + b + j) : "memory");
        } else {
            dst[j] = 0.0f;  // fl(s + 0) = s: padding does not change the chain
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__global__ void __launch_bounds__(32) pp_pick_kernel(const float *mind, int64_t n, int c, uint64_t seed, int g_base, int *zero,
                                                     int64_t *chosen, float *ckpt /*[G][tiles]*/) {
    __shared__ __align__(16) float buf[2][PP_TILE];
    const int g = blockIdx.x, lane = threadIdx.x;
    const float *m = mind + (int64_t)g * n;
    const int64_t tiles = (n + PP_TILE - 1) / PP_TILE;
    float *ck = ckpt + (int64_t)g * tiles;
    float sum = 0.0f;
    pp_load_tile(m, n, 0, buf[0], lane);
    for (int64_t t = 0; t < tiles; t++) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (t + 1 < tiles) pp_load_tile(m, n, t + 1, buf[(t + 1) & 1], lane);
        if (lane == 0) ck[t] = sum;
        // the chain is one dependent FADD per element (4 cycles); the shared-memory reads of the NEXT 32 elements are issued
        // before the current 32 are added, so their latency stays out of the chain
        const float4 *b4 = reinterpret_cast<const float4 *>(buf[t & 1]);
        float4 cur[8], nxt[8];
#pragma unroll
        for (int u = 0; u < 8; u++) cur[u] = b4[u];
#pragma unroll 1
        for (int j = 8; j <= PP_TILE / 4; j += 8) {
            if (j < PP_TILE / 4) {
#pragma unroll
                for (int u = 0; u < 8; u++) nxt[u] = b4[j + u];
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                sum = __fadd_rn(sum, cur[u].x);
                sum = __fadd_rn(sum, cur[u].y);
                sum = __fadd_rn(sum, cur[u].z);
                sum = __fadd_rn(sum, cur[u].w);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) cur[u] = nxt[u];
        }
        __syncwarp();
    }
    if (sum == 0.0f) {
        if (lane == 0) {
            zero[g] = 1;
            chosen[g] = rng_intn(seed, (uint64_t)(g_base + g), (uint64_t)c, n);
        }
        return;
    }
    const float target = __fmul_rn(rng_f32(seed, (uint64_t)(g_base + g), (uint64_t)c), sum);
    __syncwarp();
    long long bt = 0;
    for (int64_t t = lane; t < tiles; t += 32)
        if (ck[t] < target) bt = t;  // t increases per lane: the last hit is the lane's largest
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, bt, o);
        bt = other > bt ? other : bt;
    }
    const int64_t b = (int64_t)bt * PP_TILE;
    for (int j = lane; j < PP_TILE; j += 32) buf[0][j] = (b + j < n) ? m[b + j] : 0.0f;
    __syncwarp();
    if (lane == 0) {
        const int cnt = (int)((n - b < PP_TILE) ? (n - b) : PP_TILE);
        float cum = ck[bt];
        int64_t pick = 0;
        for (int j = 0; j < cnt; j++) {
            cum = __fadd_rn(cum, buf[0][j]);
            if (cum >= target) {
                pick = b + j;
                break;
            }
        }
        zero[g] = 0;
        chosen[g] = pick;
    }
}
// ---------------------------------------------------------------- exact parallel form of the float32 prefix chain
// s_{i+1} = fl(s_i + x_i), x_i >= 0, is sequential, but inside one binade of s it is INTEGER arithmetic: with
// u = ulp(s) and s = S u (S in [2^23, 2^24)), fl(s + x) = (S + round(x / u)) u where round() is to nearest and an
// exact half goes to the side that makes the result even.  So every element is a function of S of the form
// S -> S + m (no tie) or S -> S + floor + [(S + floor) odd] (tie; the result is even, so every later tie of the block
// is decided), a block of elements composes to  F(S) = S + K + t [(S + c) odd]  and these (K, c, t, parity) summaries
// form a monoid: a block scan gives every thread the exact state at the start of its 32 elements.  The binade
// changes ~25 times over a million elements; the chunk in which S would reach 2^24 is walked with real FADDs by its
// thread and the scan restarts behind it with the new ulp.  The result — every 32-element checkpoint and the total —
// is bit-identical to the sequential loop of pq.go:299-303,327-329 (tests/test_gpu_parity.py: PQ / OPQ training
// against the sequential chain; tools/prefix_proto.py is the Python model of the algorithm with its fuzz test).
constexpr int PX_T = 1024;          // threads per group (one CTA per group: the whole SM works on one chain)
constexpr int PX_CH = 32;           // elements per thread per round
constexpr int PX_HUGE = 1 << 29;    // "certainly leaves the binade"
constexpr size_t PX_SMEM = (size_t)PX_T * (PX_CH + 1) * 4;
struct PxSum {
    int K;       // increment (saturating at PX_HUGE)
    int f;       // bit 0: c, bit 1: t (tie seen), bit 2: parity of the output once a tie was seen
};
__device__ __forceinline__ PxSum px_compose(PxSum a, PxSum b) {  // b after a
    PxSum r;
    if (a.K >= PX_HUGE || b.K >= PX_HUGE) {
        r.K = PX_HUGE;
        r.f = 0;
        return r;
    }
    const int ca = a.f & 1, ta = (a.f >> 1) & 1, pa = (a.f >> 2) & 1;
    const int cb = b.f & 1, tb = (b.f >> 1) & 1, pb = (b.f >> 2) & 1;
    if (ta) {
        const int extra = tb & ((pa + cb) & 1);
        r.K = a.K + b.K + extra;
        const int pi = tb ? pb : ((pa + b.K) & 1);
        r.f = ca | 2 | (pi << 2);
    } else if (tb) {
        r.K = a.K + b.K;
        r.f = ((a.K + cb) & 1) | 2 | (pb << 2);
    } else {
        r.K = a.K + b.K;
        r.f = 0;
    }
    if (r.K >= PX_HUGE) {
        r.K = PX_HUGE;
        r.f = 0;
    }
    return r;
}
__device__ __forceinline__ int px_apply(PxSum a, int S) { return S + a.K + (((a.f >> 1) & 1) & ((S + (a.f & 1)) & 1)); }
// summary of up to PX_CH elements (shared-memory row) under ulp 2^Eu
__device__ __forceinline__ PxSum px_summarize(const float *x, int cnt, int Eu) {
    int K = 0, c = 0, t = 0, pi = 0;
    for (int i = 0; i < cnt; i++) {
        const uint32_t b = __float_as_uint(x[i]);
        const int ex = (int)((b >> 23) & 0xFF);
        const uint32_t fr = b & 0x7FFFFFu;
        const uint32_t M = ex ? (fr | 0x800000u) : fr;
        if (M == 0) continue;
        const int E = ex ? ex - 150 : -149;
        const int d = Eu - E;
        if (d <= 0 || ex == 255) {
            K = PX_HUGE;
            break;
        }
        if (d >= 25) continue;
        const int fl = (int)(M >> d);
        const uint32_t rem = M & ((1u << d) - 1u), half = 1u << (d - 1);
        if (rem != half) {
            const int m = fl + (rem > half ? 1 : 0);
            K += m;
            if (t) pi = (pi + m) & 1;
        } else if (!t) {
            c = (K + fl) & 1;
            K += fl;
            t = 1;
            pi = 0;
        } else {
            K += fl + ((pi + fl) & 1);
            pi = 0;
        }
        if (K >= PX_HUGE) {
            K = PX_HUGE;
            break;
        }
    }
    PxSum r;
    r.K = K;
    r.f = K >= PX_HUGE ? 0 : (c | (t << 1) | (pi << 2));
    return r;
}
// One CTA per group: checkpoints ck[g][chunk] = running sum before element 32 * chunk, the total, then the k-means++
// pick (pq.go:305-320) by binary search over the checkpoints (the chain is non-decreasing) + a 32-element walk.
__global__ void __launch_bounds__(PX_T) pp_pick_parallel_kernel(const float *mind, int64_t n, int c, uint64_t seed, int g_base, int *zero,
                                                                int64_t *chosen, float *ckpt /*[G][chunks]*/) {
    extern __shared__ float tile[];  // [PX_T][PX_CH + 1]
    __shared__ PxSum wsum[PX_T / 32], wpre[PX_T / 32];
    __shared__ float s_state;
    __shared__ int s_cross;
    __shared__ int s_cnt[PX_T / 32];
    const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *m = mind + (int64_t)g * n;
    const int64_t chunks = (n + PX_CH - 1) / PX_CH;
    float *ck = ckpt + (int64_t)g * chunks;
    const bool vec_ok = (n & 3) == 0 && (reinterpret_cast<uintptr_t>(mind) & 15) == 0;
    if (tid == 0) s_state = 0.0f;
    __syncthreads();
    int64_t p = 0;
    while (p < n) {
        // coalesced load of the next PX_T * PX_CH elements, one padded shared-memory row per thread
        if (vec_ok && p + (int64_t)PX_T * PX_CH <= n) {
            float4 v[PX_CH / 4];
#pragma unroll
            for (int j = 0; j < PX_CH / 4; j++) v[j] = __ldg(reinterpret_cast<const float4 *>(m + p) + j * PX_T + tid);
#pragma unroll
            for (int j = 0; j < PX_CH / 4; j++) {
                const int i = (j * PX_T + tid) * 4;  // element index inside the round: 4 consecutive elements of one row
                float *dst = tile + (i / PX_CH) * (PX_CH + 1) + (i % PX_CH);
                dst[0] = v[j].x;
                dst[1] = v[j].y;
                dst[2] = v[j].z;
                dst[3] = v[j].w;
            }
        } else {
            for (int i = tid; i < PX_T * PX_CH; i += PX_T) {
                const int64_t e = p + i;
                tile[(i / PX_CH) * (PX_CH + 1) + (i % PX_CH)] = e < n ? m[e] : 0.0f;
            }
        }
        if (tid == 0) s_cross = PX_T;
        __syncthreads();
        const float S = s_state;
        const float *row = tile + tid * (PX_CH + 1);
        const int64_t my0 = p + (int64_t)tid * PX_CH;
        const int cnt = my0 >= n ? 0 : (int)((n - my0 < PX_CH) ? (n - my0) : PX_CH);
        const uint32_t sb = __float_as_uint(S);
        const int sex = (int)((sb >> 23) & 0xFF);
        if (sex == 0 || sex == 255) {
            // zero / denormal / non-finite state: skip whole chunks of zeros, then one thread walks its chunk
            bool nz = false;
            if (S == 0.0f)
                for (int i = 0; i < cnt; i++) nz |= row[i] != 0.0f;
            else nz = tid == 0;
            if (nz) atomicMin(&s_cross, tid);
            __syncthreads();
            const int f = s_cross;
            if (cnt > 0 && tid <= f) ck[my0 / PX_CH] = S;   // chunks before the first non-zero one start at S (= 0) too
            if (tid == f) {
                float acc = S;
                for (int i = 0; i < cnt; i++) acc = __fadd_rn(acc, row[i]);
                s_state = acc;
            }
            __syncthreads();
            p += (int64_t)(f < PX_T ? f + 1 : PX_T) * PX_CH;
            continue;
        }
        const int Eu = sex - 150;
        const int S0 = (int)((sb & 0x7FFFFFu) | 0x800000u);
        PxSum mine = px_summarize(row, cnt, Eu);
        // inclusive scan of the summaries (composition is associative, not commutative: left operand = earlier elements)
        PxSum inc = mine;
        for (int o = 1; o < 32; o <<= 1) {
            PxSum prev;
            prev.K = __shfl_up_sync(0xffffffffu, inc.K, o);
            prev.f = __shfl_up_sync(0xffffffffu, inc.f, o);
            if (lane >= o) inc = px_compose(prev, inc);
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the warp summaries
            PxSum wi = wsum[lane];
            for (int o = 1; o < 32; o <<= 1) {
                PxSum prev;
                prev.K = __shfl_up_sync(0xffffffffu, wi.K, o);
                prev.f = __shfl_up_sync(0xffffffffu, wi.f, o);
                if (lane >= o) wi = px_compose(prev, wi);
            }
            PxSum ex;
            ex.K = __shfl_up_sync(0xffffffffu, wi.K, 1);
            ex.f = __shfl_up_sync(0xffffffffu, wi.f, 1);
            if (lane == 0) {
                ex.K = 0;
                ex.f = 0;
            }
            wpre[lane] = ex;
        }
        __syncthreads();
        PxSum before = wpre[warp];  // composition of everything before this thread
        {
            PxSum prev;
            prev.K = __shfl_up_sync(0xffffffffu, inc.K, 1);
            prev.f = __shfl_up_sync(0xffffffffu, inc.f, 1);
            if (lane > 0) before = px_compose(before, prev);
        }
        const bool prefix_ok = before.K < PX_HUGE;
        const int s_in = prefix_ok ? px_apply(before, S0) : (1 << 24);
        const bool crosses = !prefix_ok || s_in >= (1 << 24) || mine.K >= PX_HUGE || s_in + mine.K + 1 >= (1 << 24);
        if (crosses && cnt > 0) atomicMin(&s_cross, tid);
        __syncthreads();
        const int cross = s_cross;
        if (cnt > 0 && tid <= cross) ck[my0 / PX_CH] = ldexpf((float)s_in, Eu);  // exact: s_in < 2^24
        if (tid == cross) {
            float acc = ldexpf((float)s_in, Eu);
            for (int i = 0; i < cnt; i++) acc = __fadd_rn(acc, row[i]);
            s_state = acc;
        } else if (cross == PX_T && tid == PX_T - 1) {
            s_state = ldexpf((float)px_apply(mine, s_in), Eu);  // state after the last element of the round
        }
        __syncthreads();
        p += (int64_t)(cross < PX_T ? cross + 1 : PX_T) * PX_CH;
    }
    __syncthreads();
    const float sum = s_state;
    if (sum == 0.0f) {
        if (tid == 0) {
            zero[g] = 1;
            chosen[g] = rng_intn(seed, (uint64_t)(g_base + g), (uint64_t)c, n);
        }
        return;
    }
    const float target = __fmul_rn(rng_f32(seed, (uint64_t)(g_base + g), (uint64_t)c), sum);
    // last chunk whose start is below the target (checkpoints are non-decreasing): count them, all threads; chunk 0 when none is
    int below = 0;
    for (int64_t j = tid; j < chunks; j += PX_T) below += ck[j] < target ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if (lane == 0) s_cnt[warp] = below;
    __syncthreads();
    if (tid != 0) return;
    int64_t lo = 0;
    for (int w = 0; w < PX_T / 32; w++) lo += s_cnt[w];
    lo = lo > 0 ? lo - 1 : 0;
    int64_t pick = 0;
    bool found = false;
    float cum = ck[lo];
    for (int64_t i = lo * PX_CH; i < n && !found; i++) {
        cum = __fadd_rn(cum, m[i]);
        if (cum >= target) {
            pick = i;
            found = true;
        }
    }
    zero[g] = 0;
    chosen[g] = pick;
}

__global__ void pp_first_kernel(int64_t n, uint64_t seed, int G, int g_base, int64_t *chosen, int *zero) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    chosen[g] = rng_intn(seed, (uint64_t)(g_base + g), 0, n);
    zero[g] = 0;
}
__global__ void __launch_bounds__(256) pp_small_init_kernel(const float *vecs, int64_t n, int64_t stride, int ds, int K, int G,
                                                            float *cent) {
    // len(vectors) < k: centroid i = vectors[i % n] (pq.go:285-290)
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)G * K * ds) return;
    const int j = (int)(t % ds);
    const int64_t gc = t / ds;
    const int g = (int)(gc / K), c = (int)(gc % K);
    cent[t] = vecs[(int64_t)(c % n) * stride + (int64_t)g * ds + j];
}

// Train (pq.go:98-136): float32 centroids of one subspace → int8 + scale/offset. One CTA per subspace.
__global__ void __launch_bounds__(256) pq_quantize_kernel(const float *cent, int count, int8_t *out, float *scales, float *offsets) {
    __shared__ float smin[256], smax[256];
    const int g = blockIdx.x;
    const float *c = cent + (int64_t)g * count;
    float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const float v = c[i];
        if (v < mn) mn = v;
        if (v > mx) mx = v;
    }
    smin[threadIdx.x] = mn;
    smax[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            if (smin[threadIdx.x + o] < smin[threadIdx.x]) smin[threadIdx.x] = smin[threadIdx.x + o];
            if (smax[threadIdx.x + o] > smax[threadIdx.x]) smax[threadIdx.x] = smax[threadIdx.x + o];
        }
        __syncthreads();
    }
    mn = smin[0];
    mx = smax[0];
    if (mx == mn) mx = __fadd_rn(mn, 1e-6f);
    const float scale = __fdiv_rn(__fsub_rn(mx, mn), 255.0f);
    const float offset = __fadd_rn(mn, __fmul_rn(128.0f, scale));
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const float r = __fdiv_rn(__fsub_rn(c[i], mn), scale);
        int val = (int)round((double)r);
        if (val < 0) val = 0;
        if (val > 255) val = 255;
        out[(int64_t)g * count + i] = (int8_t)(val - 128);
    }
    if (threadIdx.x == 0) {
        scales[g] = scale;
        offsets[g] = offset;
    }
}

template <int DS>
static vg_status launch_small_assign(const float *d_vecs, int64_t n, int64_t dim, int G, int K, const float *d_cent,
                                     uint32_t *d_assign, cudaStream_t st) {
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)G);
    if (K % 2 == 0 && DS <= 16) pq_assign_small2_kernel<DS><<<grid, 256, (size_t)K * DS * 4, st>>>(d_vecs, n, dim, K, d_cent, d_assign);
    else pq_assign_small_kernel<DS><<<grid, 256, (size_t)K * DS * 4, st>>>(d_vecs, n, dim, K, d_cent, d_assign);
    VG_LAUNCHED();
    return VG_OK;
}

static vg_status pq_assign_all(const float *d_vecs, int64_t n, int64_t dim, int G, int K, int ds, const float *d_cent,
                               uint32_t *d_assign, DevBuf &score, DevBuf &cnt, cudaStream_t st) {
    switch (ds) {
        case 2: return launch_small_assign<2>(d_vecs, n, dim, G, K, d_cent, d_assign, st);
        case 4: return launch_small_assign<4>(d_vecs, n, dim, G, K, d_cent, d_assign, st);
        case 8: return launch_small_assign<8>(d_vecs, n, dim, G, K, d_cent, d_assign, st);
        case 16: return launch_small_assign<16>(d_vecs, n, dim, G, K, d_cent, d_assign, st);
        case 32: return launch_small_assign<32>(d_vecs, n, dim, G, K, d_cent, d_assign, st);
        default: break;
    }
    if (!score.p) VG_TRY(score.alloc((size_t)n * 4));
    if (!cnt.p) VG_TRY(cnt.alloc((size_t)n * 4));
    for (int g = 0; g < G; g++)
        VG_TRY(assign_generic(d_vecs + (int64_t)g * ds, n, dim, ds, d_cent + (int64_t)g * K * ds, K, VG_VAR_PAIR, 0,
                              d_assign + (int64_t)g * n, score.as<float>(), cnt.as<int32_t>(), st));
    return VG_OK;
}

}  // namespace vg

using namespace vg;

namespace vg {
// VECGO_KMEANSPP_SEQUENTIAL=1 keeps the one-lane sequential prefix chain (A/B check of the parallel exact form)
static bool pp_sequential_pick() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("VECGO_KMEANSPP_SEQUENTIAL");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v != 0;
}
// ProductQuantizer.Train on device-resident vectors (pq.go:68-143,275-433); outputs stay on the device.
vg_status dev_pq_train(const float *d_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t iters, uint64_t seed, DevBuf &cent,
                       DevBuf &cb, DevBuf &sc, DevBuf &of, cudaStream_t st) {
    return dev_pq_train_range(d_vecs, n, dim, m, k, iters, seed, 0, m, cent, cb, sc, of, st);
}
// Subspaces [g0, g1) of the quantizer only (pq.go:79-140 trains every subspace in its own goroutine: they are independent,
// so a subset trained elsewhere — another GPU — gives the same centroids bit for bit).  Outputs hold g1 - g0 subspaces.
vg_status dev_pq_train_range(const float *d_vecs_all, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t iters, uint64_t seed, int64_t g0,
                             int64_t g1, DevBuf &cent, DevBuf &cb, DevBuf &sc, DevBuf &of, cudaStream_t st) {
    const int G = (int)(g1 - g0), K = (int)k, ds = (int)(dim / m), g_base = (int)g0;
    const float *d_vecs = d_vecs_all + g0 * ds;   // column slice: same row stride `dim`
    DevBuf mind, zero, chosen, score, cnt, ckpt;
    VG_TRY(cent.alloc((size_t)G * K * ds * 4));
    // ---- initializeCentroids
    if (n < k) {
        const int64_t t = (int64_t)G * K * ds;
        pp_small_init_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(d_vecs, n, dim, ds, K, G, cent.as<float>());
        VG_LAUNCHED();
    } else {
        VG_TRY(mind.alloc((size_t)G * n * 4));
        VG_TRY(zero.alloc((size_t)G * 4));
        VG_TRY(chosen.alloc((size_t)G * 8));
        VG_TRY(ckpt.alloc((size_t)G * ((n + PX_CH - 1) / PX_CH) * 4));
        pp_first_kernel<<<(G + 63) / 64, 64, 0, st>>>(n, seed, G, g_base, chosen.as<int64_t>(), zero.as<int>());
        VG_LAUNCHED();
        const size_t pp_sm = (size_t)G * (PP_ROWS + 1) * 4;
        // subspace lengths with a compile-time kernel (16-byte loads, unrolled chain); the row stride must keep them aligned
        const int dsk = (dim % 4 == 0 && (ds == 4 || ds == 8 || ds == 16 || ds == 32)) ? ds : 0;
        auto pp_dist = dsk == 4 ? pp_dist_kernel<4> : dsk == 8 ? pp_dist_kernel<8> : dsk == 16 ? pp_dist_kernel<16> : dsk == 32 ? pp_dist_kernel<32> : pp_dist_kernel<0>;
        if (pp_sm > 48 * 1024) VG_CUDA(cudaFuncSetAttribute(pp_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp_sm));
        VG_CUDA(cudaFuncSetAttribute(pp_pick_parallel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PX_SMEM));
        for (int c = 0; c < K; c++) {
            if (c > 0) {
                if (pp_sequential_pick())
                    pp_pick_kernel<<<G, 32, 0, st>>>(mind.as<float>(), n, c, seed, g_base, zero.as<int>(), chosen.as<int64_t>(), ckpt.as<float>());
                else
                    pp_pick_parallel_kernel<<<G, PX_T, PX_SMEM, st>>>(mind.as<float>(), n, c, seed, g_base, zero.as<int>(), chosen.as<int64_t>(),
                                                                ckpt.as<float>());
                VG_LAUNCHED();
            }
            gather_centroid_kernel<<<(G * ds + 255) / 256, 256, 0, st>>>(d_vecs, dim, ds, K, G, c, chosen.as<int64_t>(),
                                                                       cent.as<float>());
            VG_LAUNCHED();
            if (c + 1 < K) {
                pp_dist<<<(unsigned)((n + PP_ROWS - 1) / PP_ROWS), 256, pp_sm, st>>>(d_vecs, n, dim, ds, K, G, c, cent.as<float>(),
                                                                                            zero.as<int>(), mind.as<float>());
                VG_LAUNCHED();
            }
        }
    }
    // ---- runKMeansIterations
    Lloyd L;
    VG_TRY(L.init(G, K, ds, n, dim, st));
    L.g_base = g_base;
    // 8-dim subspaces x 256 centroids: assignment on the tensor cores with an exactness certificate (vg_pq_assign_tc.cu)
    pqa::Assigner tca;
    if (iters > 0 && pqa::Assigner::supported(n, dim, G, K, ds)) VG_TRY(tca.prepare(d_vecs, n, dim, G, st));
    for (int64_t it = 0; it < iters; it++) {
        if (tca.ready) VG_TRY(tca.assign(cent.as<float>(), L.assign_new.as<uint32_t>(), st));
        else VG_TRY(pq_assign_all(d_vecs, n, dim, G, K, ds, cent.as<float>(), L.assign_new.as<uint32_t>(), score, cnt, st));
        bool any = false;
        VG_TRY(L.step(d_vecs, cent.as<float>(), 0, seed, 0xE0E0E0E0ull, 1, &any, st));
        if (tca.ready) VG_TRY(tca.account(st));
        if (!any) break;
    }
    // ---- int8 codebooks
    VG_TRY(cb.alloc((size_t)G * K * ds));
    VG_TRY(sc.alloc((size_t)G * 4));
    VG_TRY(of.alloc((size_t)G * 4));
    pq_quantize_kernel<<<G, 256, 0, st>>>(cent.as<float>(), K * ds, cb.as<int8_t>(), sc.as<float>(), of.as<float>());
    VG_LAUNCHED();
    return VG_OK;
}
}  // namespace vg

extern "C" {

vg_status vg_kmeans_find_closest(const float *h_queries, int64_t nq, int64_t dim, const float *h_centroids, int64_t k,
                                 int64_t nprobe, int32_t metric, int32_t *h_out) {
    VG_ENTER();
    if (nq <= 0) return VG_OK;
    if (k <= 0 || dim <= 0 || nprobe <= 0) return fail(VG_ERR_INVALID, "bad shape");
    if (nprobe > k) nprobe = k;
    DevBuf q, c, out;
    VG_TRY(q.alloc((size_t)nq * dim * 4));
    VG_TRY(staged_h2d(q.p, h_queries, (size_t)nq * dim * 4));
    VG_TRY(c.alloc((size_t)k * dim * 4));
    VG_TRY(staged_h2d(c.p, h_centroids, (size_t)k * dim * 4));
    VG_TRY(out.alloc((size_t)nq * nprobe * 4));
    VG_TRY(dev_find_closest(q.as<float>(), nq, dim, c.as<float>(), k, nprobe, metric, out.as<int32_t>(), stream()));
    return staged_d2h(h_out, out.p, (size_t)nq * nprobe * 4);
}

vg_status vg_kmeans_assign(const float *h_vecs, int64_t n, int64_t dim, const float *h_centroids, int64_t k, int32_t metric,
                           int32_t *h_assign) {
    return vg_kmeans_find_closest(h_vecs, n, dim, h_centroids, k, 1, metric, h_assign);
}

vg_status vg_kmeans_train(const float *h_vecs, int64_t n, int64_t dim, int64_t k, int32_t metric, int64_t max_iter,
                          const int64_t *h_init_rows, uint64_t seed, float *h_centroids, int32_t *h_assign, int64_t *iters_run) {
    VG_ENTER();
    if (dim <= 0 || k <= 0) return fail(VG_ERR_INVALID, "bad shape");
    if (n < k) return fail(VG_ERR_INVALID, "not enough vectors to cluster (n < k)");  // Go returns (nil, nil)
    if (metric != VG_METRIC_L2 && metric != VG_METRIC_COSINE && metric != VG_METRIC_DOT)
        return fail(VG_ERR_UNSUPPORTED, "unsupported metric for float32");
    cudaStream_t st = stream();
    DevBuf v, cent, score, cnt;
    VG_TRY(v.alloc((size_t)n * dim * 4));
    VG_TRY(staged_h2d(v.p, h_vecs, (size_t)n * dim * 4));
    VG_TRY(cent.alloc((size_t)k * dim * 4));
    for (int64_t i = 0; i < k; i++) {
        if (h_init_rows[i] < 0 || h_init_rows[i] >= n) return fail(VG_ERR_INVALID, "init row out of range");
        VG_CUDA(cudaMemcpyAsync(cent.as<float>() + i * dim, v.as<float>() + h_init_rows[i] * dim, (size_t)dim * 4,
                                cudaMemcpyDeviceToDevice, st));
    }
    VG_TRY(score.alloc((size_t)n * 4));
    VG_TRY(cnt.alloc((size_t)n * 4));
    Lloyd L;
    VG_TRY(L.init(1, (int)k, (int)dim, n, dim, st));
    const int is_dot = metric != VG_METRIC_L2;
    int64_t it = 0;
    for (; it < max_iter; it++) {
        VG_TRY(assign_generic(v.as<float>(), n, dim, dim, cent.as<float>(), k, VG_VAR_BATCH, is_dot, L.assign_new.as<uint32_t>(),
                              score.as<float>(), cnt.as<int32_t>(), st));
        bool any = false;
        VG_TRY(L.step(v.as<float>(), cent.as<float>(), 1, seed, 0xE0E0E0E0ull, 0, &any, st));
        if (!any) break;
    }
    VG_CUDA(cudaStreamSynchronize(st));
    if (iters_run) *iters_run = it;
    VG_TRY(staged_d2h(h_centroids, cent.p, (size_t)k * dim * 4));
    if (h_assign) VG_TRY(staged_d2h(h_assign, L.assign_old.p, (size_t)n * 4));
    return VG_OK;
}

vg_status vg_pq_train(const float *h_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t iters, uint64_t seed,
                      int8_t *h_codebooks, float *h_scales, float *h_offsets, float *h_centroids_f32) {
    VG_ENTER();
    if (n <= 0) return fail(VG_ERR_INVALID, "no vectors provided for training");
    if (m <= 0 || dim <= 0 || dim % m != 0) return fail(VG_ERR_INVALID, "dimension must be divisible by numSubvectors");
    if (k <= 0 || k > 256) return fail(VG_ERR_INVALID, "numCentroids must be <= 256 for uint8 encoding");
    cudaStream_t st = stream();
    const int G = (int)m, K = (int)k, ds = (int)(dim / m);
    DevBuf v, cent, cb, sc, of;
    VG_TRY(v.alloc((size_t)n * dim * 4));
    VG_TRY(staged_h2d(v.p, h_vecs, (size_t)n * dim * 4));
    VG_TRY(dev_pq_train(v.as<float>(), n, dim, m, k, iters, seed, cent, cb, sc, of, st));
    VG_CUDA(cudaStreamSynchronize(st));
    VG_TRY(staged_d2h(h_codebooks, cb.p, (size_t)G * K * ds));
    VG_TRY(staged_d2h(h_scales, sc.p, (size_t)G * 4));
    VG_TRY(staged_d2h(h_offsets, of.p, (size_t)G * 4));
    if (h_centroids_f32) VG_TRY(staged_d2h(h_centroids_f32, cent.p, (size_t)G * K * ds * 4));
    return VG_OK;
}

// Same with the training set already resident on the device (the writer staged the segment's vectors, or they were
// generated there): no host copy of the vectors inside the call.
vg_status vg_pq_train_dev(const float *d_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t iters, uint64_t seed,
                          int8_t *h_codebooks, float *h_scales, float *h_offsets, float *h_centroids_f32) {
    VG_ENTER();
    if (n <= 0) return fail(VG_ERR_INVALID, "no vectors provided for training");
    if (m <= 0 || dim <= 0 || dim % m != 0) return fail(VG_ERR_INVALID, "dimension must be divisible by numSubvectors");
    if (k <= 0 || k > 256) return fail(VG_ERR_INVALID, "numCentroids must be <= 256 for uint8 encoding");
    cudaStream_t st = stream();
    const int G = (int)m, K = (int)k, ds = (int)(dim / m);
    DevBuf cent, cb, sc, of;
    VG_TRY(dev_pq_train(d_vecs, n, dim, m, k, iters, seed, cent, cb, sc, of, st));
    VG_CUDA(cudaStreamSynchronize(st));
    VG_TRY(staged_d2h(h_codebooks, cb.p, (size_t)G * K * ds));
    VG_TRY(staged_d2h(h_scales, sc.p, (size_t)G * 4));
    VG_TRY(staged_d2h(h_offsets, of.p, (size_t)G * 4));
    if (h_centroids_f32) VG_TRY(staged_d2h(h_centroids_f32, cent.p, (size_t)G * K * ds * 4));
    return VG_OK;
}

vg_status vg_pq_train_range_dev(const float *d_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, int64_t iters, uint64_t seed,
                                int64_t subspace_lo, int64_t subspace_hi, int8_t *h_codebooks, float *h_scales, float *h_offsets,
                                float *h_centroids_f32) {
    VG_ENTER();
    if (n <= 0) return fail(VG_ERR_INVALID, "no vectors provided for training");
    if (m <= 0 || dim <= 0 || dim % m != 0) return fail(VG_ERR_INVALID, "dimension must be divisible by numSubvectors");
    if (k <= 0 || k > 256) return fail(VG_ERR_INVALID, "numCentroids must be <= 256 for uint8 encoding");
    if (subspace_lo < 0 || subspace_hi > m || subspace_lo > subspace_hi) return fail(VG_ERR_INVALID, "subspace range outside the quantizer");
    if (subspace_lo == subspace_hi) return VG_OK;
    cudaStream_t st = stream();
    const int G = (int)(subspace_hi - subspace_lo), K = (int)k, ds = (int)(dim / m);
    DevBuf cent, cb, sc, of;
    VG_TRY(dev_pq_train_range(d_vecs, n, dim, m, k, iters, seed, subspace_lo, subspace_hi, cent, cb, sc, of, st));
    VG_CUDA(cudaStreamSynchronize(st));
    VG_TRY(staged_d2h(h_codebooks, cb.p, (size_t)G * K * ds));
    VG_TRY(staged_d2h(h_scales, sc.p, (size_t)G * 4));
    VG_TRY(staged_d2h(h_offsets, of.p, (size_t)G * 4));
    if (h_centroids_f32) VG_TRY(staged_d2h(h_centroids_f32, cent.p, (size_t)G * K * ds * 4));
    return VG_OK;
}

}  // extern "C"
