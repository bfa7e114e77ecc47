// vg_common.cuh — shared device/host helpers for libvecgo_cuda (sm_100a only).
//
// Arithmetic contract: every float32 operation on a distance path is written
// with an explicit round-to-nearest intrinsic (__fmaf_rn / __fadd_rn / ...)
// and the library is compiled with -fmad=false, so the instruction stream
// fuses exactly where the reference's AVX-512 kernels fuse and nowhere else
// (/root/reference/internal/simd/src/*_avx512.c).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/vecgo_cuda.h"

namespace vg {

// ------------------------------------------------------------------ errors
void set_error(const std::string &msg);
vg_status fail(vg_status code, const std::string &msg);
vg_status cuda_fail(cudaError_t e, const char *what);
extern std::atomic<uint64_t> g_launches;
// The stream of the API call running on this thread (see Call below).
cudaStream_t stream();

// One API call.  Constructed at the top of every extern "C" entry point: binds the calling thread to a device context
// (the handle's device, else the device this thread selected with vg_init, else the process default) and to ONE stream
// for the duration of the call: the handle's stream (vg_index_set_stream), else the calling thread's stream
// (vg_set_stream), else a stream leased from the device context's pool — so concurrent callers (one goroutine per
// query in the reference, engine.go:1320-1360) run on different streams and never share scratch.  Nested entry points
// (vg_index_rerank -> vg_index_rerank_dev) reuse the outer call's binding.
struct Call {
    vg_status status = VG_OK;
    bool outer = false;
    explicit Call(int device = -1, bool handle_stream_set = false, cudaStream_t handle_stream = nullptr);
    ~Call();
    Call(const Call &) = delete;
    Call &operator=(const Call &) = delete;
    bool leased() const;      // the call runs on a library-owned stream (nobody else can order work after it)
    // Results written by a call on a leased stream must be complete when the entry point returns; on a caller-provided
    // stream the call stays stream-ordered.
    vg_status finish();
};
#define VG_ENTER(...)                       \
    ::vg::Call _vg_call{__VA_ARGS__};       \
    do {                                    \
        if (_vg_call.status != VG_OK) return _vg_call.status; \
    } while (0)

#define VG_CUDA(expr)                                              \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return ::vg::cuda_fail(_e, #expr);  \
    } while (0)
#define VG_TRY(expr)                    \
    do {                                \
        vg_status _s = (expr);          \
        if (_s != VG_OK) return _s;     \
    } while (0)
#define VG_LAUNCHED()                                              \
    do {                                                           \
        ::vg::g_launches.fetch_add(1, std::memory_order_relaxed);  \
        VG_CUDA(cudaGetLastError());                               \
    } while (0)

// RAII device buffer for temporaries inside one API call.
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    bool pooled = false;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    vg_status alloc(size_t n);              // per-call scratch: stream-ordered pool (cudaMallocAsync)
    vg_status alloc_persistent(size_t n);   // long-lived index sections: cudaMalloc, never part of the scratch pool
    void release();
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};
// Host→device through the pinned staging ring (chunked, double buffered).
vg_status staged_h2d(void *d_dst, const void *h_src, size_t bytes);
vg_status staged_d2h(void *h_dst, const void *d_src, size_t bytes);
int sm_count();

// ------------------------------------------------------------ device math
#ifdef __CUDACC__
// _mm512_reduce_add_ps over the 16 lanes of a half-warp, in the order the
// reference's assembly uses (i+8, i+4, i+2, i+1; floats_avx512.s:46-53).
// Result is valid in lane 0 of each aligned 16-lane group.
__device__ __forceinline__ float reduce16(float v) {
    v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, 8, 16));
    v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, 4, 16));
    v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, 2, 16));
    v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, 1, 16));
    return v;
}

// uint8 → float32, exact.
__device__ __forceinline__ float u8_to_f32(uint32_t b) { return __uint2float_rn(b); }

// Sortable 64-bit key: (score under the heap order, row).  Smaller key = better
// candidate.  -0.0 is canonicalised to +0.0 so that equal scores tie on the row
// like Go's `a.Score != b.Score` does (candidate_queue.go:12-23).
__device__ __forceinline__ uint32_t f32_orderable(float f) {
    uint32_t u = __float_as_uint(__fadd_rn(f, 0.0f));
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t row, bool descending) {
    float s = descending ? -score : score;
    return ((unsigned long long)f32_orderable(s) << 32) | row;
}
__device__ __forceinline__ float key_score(unsigned long long key, bool descending) {
    float s = f32_from_orderable((uint32_t)(key >> 32));
    // a zero score is stored as +0.0 (make_key canonicalises); negating it back must not turn it into -0.0: the
    // reference's accumulators start at +0.0 and return +0.0 for orthogonal / zero vectors
    return descending ? __fadd_rn(-s, 0.0f) : s;
}
__device__ __forceinline__ uint32_t key_row(unsigned long long key) { return (uint32_t)key; }
#define VG_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#endif

}  // namespace vg
