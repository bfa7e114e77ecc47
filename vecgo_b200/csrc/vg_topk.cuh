// vg_topk.cuh — shared-memory bounded top-k used by every scan kernel.
//
// Semantics = searcher.CandidateHeap + TryPushBounded
// (/root/reference/internal/searcher/candidate_queue.go:12-38,120-134): keep the
// k smallest candidates under the total order (score asc | desc, then row asc).
// Because the order is total and a candidate only replaces the worst when
// strictly better, the surviving SET is independent of scan order, so a
// parallel threshold filter + periodic compaction returns exactly what the
// reference's sequential 4-ary heap returns.
//
// Mechanics: candidates are 64-bit sortable keys (vg_common.cuh::make_key).
// Each query owns keys[C] in shared memory, a count and a threshold tau (the
// current k-th best key, or EMPTY while fewer than k are held).  Producers
// append keys < tau with one shared-memory atomicAdd; when a buffer may
// overflow during the next tile, a warp bitonic-sorts it, keeps the best k and
// tightens tau.  After the first few tiles only ~k/row_index of the rows pass
// the filter, so the cost per (query,row) pair is one compare.
#pragma once
#include "vg_common.cuh"

namespace vg {

struct TopK {
    unsigned long long *keys;  // [nq_slots][C]
    unsigned long long *tau;   // [nq_slots]
    int *cnt;                  // [nq_slots]
    int *flag;                 // set when some buffer passed the trigger.  TWO flags in turn (flag / flag_next swap at every
    int *flag_next;            // topk_block_maintain): a thread that has already read "no compaction" and runs ahead into
                               // the next tile's offers must not raise the flag a slower thread is still about to read —
                               // that thread would enter the compaction (and its barriers) alone (racecheck found this)
    int C;                     // capacity per query, power of two
    int k;
};

__host__ __device__ inline size_t topk_smem_bytes(int slots, int C) {
    return (size_t)slots * C * 8 + (size_t)slots * 8 + (size_t)slots * 4 + 16;
}
// Capacity: power of two >= k + 2*max pushes per query between compaction checks.
inline int topk_capacity(int k, int burst) {
    int need = k + 2 * burst;
    int c = 64;
    while (c < need) c <<= 1;
    return c;
}

#ifdef __CUDACC__
__device__ __forceinline__ TopK topk_carve(unsigned char *smem, int slots, int C, int k) {
    TopK t;
    t.keys = reinterpret_cast<unsigned long long *>(smem);
    t.tau = t.keys + (size_t)slots * C;
    t.cnt = reinterpret_cast<int *>(t.tau + slots);
    t.flag = t.cnt + slots;
    t.flag_next = t.flag + 1;
    t.C = C;
    t.k = k;
    return t;
}

__device__ __forceinline__ void topk_init(const TopK &t, int slots, int tid, int nthreads) {
    for (int i = tid; i < slots; i += nthreads) {
        t.tau[i] = VG_KEY_EMPTY;
        t.cnt[i] = 0;
    }
    if (tid == 0) {
        t.flag[0] = 0;
        t.flag[1] = 0;
    }
}

// Offer one candidate.  `trigger`: request a compaction once count exceeds it.
__device__ __forceinline__ void topk_offer(const TopK &t, int slot, unsigned long long key, int trigger) {
    if (key < t.tau[slot]) {
        int pos = atomicAdd(&t.cnt[slot], 1);
        if (pos < t.C) t.keys[(size_t)slot * t.C + pos] = key;
        if (pos >= trigger) *t.flag = 1;
    }
}

// One warp: sort slot's buffer ascending, keep the best k, refresh tau.
__device__ __forceinline__ void topk_compact_warp(const TopK &t, int slot, int lane, bool force_sort) {
    unsigned long long *a = t.keys + (size_t)slot * t.C;
    int n = t.cnt[slot];
    if (n > t.C) n = t.C;
    __syncwarp();
    if (n >= t.k || force_sort) {
        for (int i = n + lane; i < t.C; i += 32) a[i] = VG_KEY_EMPTY;
        __syncwarp();
        // sort only the power-of-two prefix that contains all n live keys
        int len = 32;
        while (len < n) len <<= 1;
        for (int size = 2; size <= len; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = lane; i < (len >> 1); i += 32) {
                    int lo = 2 * i - (i & (stride - 1));
                    int hi = lo + stride;
                    bool up = ((lo & size) == 0);
                    unsigned long long x = a[lo], y = a[hi];
                    if ((x > y) == up) {
                        a[lo] = y;
                        a[hi] = x;
                    }
                }
                __syncwarp();
            }
        }
        if (n > t.k) n = t.k;
    }
    if (lane == 0) {
        t.cnt[slot] = n;
        t.tau[slot] = (n >= t.k && t.k > 0) ? a[t.k - 1] : VG_KEY_EMPTY;
        if (t.k == 0) t.tau[slot] = 0;  // k = 0: accept nothing
    }
    __syncwarp();
}

// Block-wide: after a __syncthreads(), compact every slot if any producer asked.
// Must be called by all threads of the block, the same number of times (the two flags alternate per call).
__device__ __forceinline__ void topk_block_maintain(TopK &t, int slots, int tid, int nthreads) {
    int *cur = t.flag;
    t.flag = t.flag_next;   // offers of the next tile raise the OTHER flag
    t.flag_next = cur;
    if (*cur) {
        int warp = tid >> 5, lane = tid & 31, nw = nthreads >> 5;
        for (int s = warp; s < slots; s += nw) topk_compact_warp(t, s, lane, false);
        __syncthreads();
        if (tid == 0) *cur = 0;   // not raised again before the barrier that ends the tile after next
    }
}

// The same for ONE slot owned by the whole block: all threads sort (a single warp sorting 2048 keys while seven wait
// was 0.2 ms per compaction in the block selection kernel).  Called like topk_block_maintain.
__device__ __forceinline__ void topk_block_maintain_single(TopK &t, int tid, int nthreads, bool force = false) {
    int *cur = t.flag;
    t.flag = t.flag_next;
    t.flag_next = cur;
    if (*cur || force) {   // uniform: read behind the caller's barrier
        unsigned long long *a = t.keys;
        int n = t.cnt[0];
        if (n > t.C) n = t.C;
        if (n >= t.k || force) {
            for (int i = n + tid; i < t.C; i += nthreads) a[i] = VG_KEY_EMPTY;
            int len = 32;
            while (len < n) len <<= 1;
            __syncthreads();
            for (int size = 2; size <= len; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int i = tid; i < (len >> 1); i += nthreads) {
                        const int lo = 2 * i - (i & (stride - 1));
                        const int hi = lo + stride;
                        const bool up = ((lo & size) == 0);
                        const unsigned long long x = a[lo], y = a[hi];
                        if ((x > y) == up) {
                            a[lo] = y;
                            a[hi] = x;
                        }
                    }
                    __syncthreads();
                }
            }
            if (n > t.k) n = t.k;
        }
        if (tid == 0) {
            t.cnt[0] = n;
            t.tau[0] = (n >= t.k && t.k > 0) ? a[t.k - 1] : VG_KEY_EMPTY;
            if (t.k == 0) t.tau[0] = 0;
            *cur = 0;
        }
        __syncthreads();
    }
}

// Final: sorted best-first emission of one slot by one warp.
__device__ __forceinline__ void topk_emit_warp(const TopK &t, int slot, int lane, bool descending, uint32_t *out_rows,
                                               float *out_scores, int32_t *out_count, int64_t k_stride) {
    topk_compact_warp(t, slot, lane, true);
    const unsigned long long *a = t.keys + (size_t)slot * t.C;
    int n = t.cnt[slot];
    for (int i = lane; i < (int)k_stride; i += 32) {
        if (i < n) {
            out_rows[i] = key_row(a[i]);
            out_scores[i] = key_score(a[i], descending);
        } else {
            out_rows[i] = 0xFFFFFFFFu;
            out_scores[i] = __uint_as_float(0x7fc00000u);
        }
    }
    if (lane == 0) *out_count = n;
}
// Final, keys only (row-split partials: [slot][k] sorted, EMPTY padded).
__device__ __forceinline__ void topk_emit_keys_warp(const TopK &t, int slot, int lane, unsigned long long *out,
                                                    int64_t k_stride) {
    topk_compact_warp(t, slot, lane, true);
    const unsigned long long *a = t.keys + (size_t)slot * t.C;
    int n = t.cnt[slot];
    for (int i = lane; i < (int)k_stride; i += 32) out[i] = (i < n) ? a[i] : VG_KEY_EMPTY;
}
#endif

// Merge `lists` key lists per query into the best k (one warp per query).
vg_status launch_merge_keys(const unsigned long long *d_keys, int64_t lists, int64_t nq, int64_t k_in, int64_t list_stride,
                            int64_t query_stride, bool descending, int64_t k_out, uint32_t *d_rows, float *d_scores,
                            int32_t *d_counts, cudaStream_t st);
vg_status launch_merge_pairs(const uint32_t *d_rows_in, const float *d_scores_in, int64_t lists, int64_t nq, int64_t k_in,
                             bool descending, int64_t k_out, uint32_t *d_rows, float *d_scores, int32_t *d_counts,
                             cudaStream_t st);

// (rows, scores) of a result -> sortable keys (row 0xFFFFFFFF -> empty key).
vg_status launch_pack_keys(const uint32_t *d_rows_in, const float *d_scores_in, int64_t n, bool descending, unsigned long long *d_keys,
                           cudaStream_t st);

}  // namespace vg
