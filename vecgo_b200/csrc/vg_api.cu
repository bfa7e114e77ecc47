// vg_api.cu — the C ABI (include/vecgo_cuda.h): runtime, pinned staging,
// device-resident scan indexes, flat-segment opening, simd/quantizer mirrors.
// Host-side logic mirrors the Go callers it replaces; all arithmetic that
// decides a result runs in the CUDA kernels of vg_scan.cu / vg_quant.cu /
// vg_kmeans.cu.  There is no CPU compute path in this file.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "vg_kmeans.cuh"
#include "vg_quant.cuh"
#include "vg_scan.cuh"
#include "vg_flat_tc.cuh"
#include "vg_quant_tc.cuh"
#include "vg_tiles.cuh"
#include "vg_pq_assign_tc.cuh"

namespace vg {

// ------------------------------------------------------------------ runtime
// State is per DEVICE (streams, pinned staging ring) and per THREAD (the device selected with vg_init, the stream set
// with vg_set_stream, the binding of the call in flight); nothing about a call lives in a process-wide variable, so a
// host may drive several GPUs from one process and many threads may search one handle at once (Threading).
static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};
static const int kMaxDevices = 64;
static const size_t kStageBytes = 32u << 20;

struct DeviceCtx {
    int device = -1;
    int sms = 0;
    std::mutex mu;                           // stream leases
    std::vector<cudaStream_t> free_streams;  // non-blocking streams owned by the library, handed out one per call
    // Pinned staging ring: two 32 MiB page-locked buffers.  The CPU fills one while the copy engine drains the other
    // (mmap'd segment pages are pageable, so this is where they become DMA-able).
    std::mutex stage_mu;
    void *stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
};
static DeviceCtx *g_ctx[kMaxDevices] = {};
static std::mutex g_mu;                      // contexts and the handle table
static std::atomic<int> g_default_device{-1};

struct ThreadState {
    int device = -1;                 // vg_init on this thread (-1: process default)
    bool user_stream_set = false;    // vg_set_stream on this thread
    cudaStream_t user_stream = nullptr;
    // the call in flight
    int depth = 0;
    DeviceCtx *ctx = nullptr;
    cudaStream_t st = nullptr;
    bool leased = false;
};
static thread_local ThreadState t_ts;

void set_error(const std::string &msg) { t_error = msg; }
vg_status fail(vg_status code, const std::string &msg) {
    t_error = msg;
    return code;
}
vg_status cuda_fail(cudaError_t e, const char *what) {
    t_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    cudaGetLastError();
    return VG_ERR_CUDA;
}
cudaStream_t stream() { return t_ts.st; }
int sm_count() { return t_ts.ctx && t_ts.ctx->sms > 0 ? t_ts.ctx->sms : 148; }

static vg_status get_ctx(int device, DeviceCtx **out) {
    if (device < 0 || device >= kMaxDevices) return fail(VG_ERR_INVALID, "device ordinal out of range");
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx[device]) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            if (e != cudaSuccess) cudaGetLastError();
            return fail(VG_ERR_CUDA, "no CUDA device: libvecgo_cuda has no CPU path");
        }
        if (device >= n) return fail(VG_ERR_INVALID, "device ordinal out of range");
        cudaDeviceProp prop;
        VG_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) return fail(VG_ERR_CUDA, "libvecgo_cuda is built for sm_100a (Blackwell) only");
        VG_CUDA(cudaSetDevice(device));
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;  // scratch stays cached in the pool (never trimmed at synchronisation points): its size is
                                    // bounded by what the calls in flight need (<= ~8.5 GiB each for the largest group-minima buffer)
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        DeviceCtx *c = new DeviceCtx();
        c->device = device;
        c->sms = prop.multiProcessorCount;
        g_ctx[device] = c;
        int none = -1;
        g_default_device.compare_exchange_strong(none, device);
    }
    *out = g_ctx[device];
    return VG_OK;
}

Call::Call(int device, bool handle_stream_set, cudaStream_t handle_stream) {
    ThreadState &ts = t_ts;
    if (ts.depth > 0) {  // nested entry point: keep the outer call's device and stream
        if (device >= 0 && ts.ctx && ts.ctx->device != device) {
            status = fail(VG_ERR_INVALID, "nested call on a different device");
            return;
        }
        ts.depth++;
        return;
    }
    if (device < 0) device = ts.device >= 0 ? ts.device : (g_default_device.load() >= 0 ? g_default_device.load() : 0);
    DeviceCtx *ctx = nullptr;
    status = get_ctx(device, &ctx);
    if (status != VG_OK) return;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        status = cuda_fail(e, "cudaSetDevice");
        return;
    }
    ts.ctx = ctx;
    ts.leased = false;
    if (handle_stream_set) {
        ts.st = handle_stream;
    } else if (ts.user_stream_set) {
        ts.st = ts.user_stream;
    } else {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (!ctx->free_streams.empty()) {
            ts.st = ctx->free_streams.back();
            ctx->free_streams.pop_back();
        } else {
            e = cudaStreamCreateWithFlags(&ts.st, cudaStreamNonBlocking);
            if (e != cudaSuccess) {
                status = cuda_fail(e, "cudaStreamCreateWithFlags");
                ts.ctx = nullptr;
                return;
            }
        }
        ts.leased = true;
    }
    ts.depth = 1;
    outer = true;
}
Call::~Call() {
    ThreadState &ts = t_ts;
    if (status != VG_OK && !outer) return;  // a failed nested constructor did not change the depth
    if (ts.depth > 0) ts.depth--;
    if (outer && ts.depth == 0) {
        if (ts.leased && ts.ctx) {
            // Work that is still queued on a leased stream (stream-ordered frees of scratch) stays ordered: the next
            // lease of this stream simply queues behind it.
            std::lock_guard<std::mutex> lk(ts.ctx->mu);
            ts.ctx->free_streams.push_back(ts.st);
        }
        ts.ctx = nullptr;
        ts.st = nullptr;
        ts.leased = false;
    }
}
bool Call::leased() const { return t_ts.leased; }
vg_status Call::finish() {
    if (t_ts.leased && t_ts.depth == 1) VG_CUDA(cudaStreamSynchronize(t_ts.st));
    return VG_OK;
}

// Scratch buffers (per-call temporaries: query copies, partial top-k lists, result staging) come from the device's
// stream-ordered memory pool — after warm-up an allocation is a pointer bump instead of a 100+ us cudaMalloc/cudaFree
// pair, which matters for small batches.  Large, long-lived buffers (code / vector sections) use cudaMalloc.
static const size_t kPoolMaxBytes = (size_t)8704 << 20;  // includes the <= 8 GiB group-minima buffers of the tensor-core filters
vg_status DevBuf::alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    cudaError_t e;
    if (n <= kPoolMaxBytes) {
        e = cudaMallocAsync(&p, n, stream());
        pooled = true;
    } else {
        e = cudaMalloc(&p, n);
        pooled = false;
    }
    if (e != cudaSuccess) {
        p = nullptr;
        return cuda_fail(e, "cudaMalloc");
    }
    bytes = n;
    return VG_OK;
}
vg_status DevBuf::alloc_persistent(size_t n) {
    release();
    if (n == 0) n = 16;
    const cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {
        p = nullptr;
        return cuda_fail(e, "cudaMalloc");
    }
    pooled = false;
    bytes = n;
    return VG_OK;
}
void DevBuf::release() {
    if (p) {
        if (pooled) cudaFreeAsync(p, stream());
        else cudaFree(p);
    }
    p = nullptr;
    bytes = 0;
}

static vg_status ensure_stage(DeviceCtx *c) {
    if (c->stage[0]) return VG_OK;
    for (int i = 0; i < 2; i++) {
        VG_CUDA(cudaHostAlloc(&c->stage[i], kStageBytes, cudaHostAllocDefault));
        VG_CUDA(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
    }
    return VG_OK;
}
// Pageable -> pinned staging copy: large chunks are split over a few host threads (one thread moves ~10 GB/s, the
// PCIe link takes ~50 GB/s), so the staging copy of the next chunk keeps up with the DMA of the current one.
static void stage_copy(void *dst, const void *src, size_t n) {
    static const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned want = (unsigned)std::min<size_t>(std::min(8u, std::max(1u, hw / 2)), n / (4u << 20));
    if (want <= 1) {
        memcpy(dst, src, n);
        return;
    }
    const size_t slice = ((n + want - 1) / want + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    th.reserve(want);
    for (size_t off = slice; off < n; off += slice) {
        const size_t len = std::min(slice, n - off);
        th.emplace_back([=] { memcpy((char *)dst + off, (const char *)src + off, len); });
    }
    memcpy(dst, src, std::min(slice, n));
    for (auto &t : th) t.join();
}
// Page-locked host memory (cudaHostAlloc / cudaHostRegister, e.g. the caller's own staging of mmap'd segments) is DMA'd
// directly; pageable memory goes through the library's pinned ring.
static bool host_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}
// Small transfers from pageable memory (a query, a few result rows): cudaMemcpyAsync stages them through the driver's
// own pinned buffer and returns once the source was read, so concurrent callers do not queue on the ring's mutex.
static const size_t kSmallCopy = 64u << 10;
vg_status staged_h2d(void *d_dst, const void *h_src, size_t bytes) {
    if (bytes == 0) return VG_OK;
    cudaStream_t st = stream();
    if (bytes < kSmallCopy || host_pinned(h_src)) {
        VG_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        VG_CUDA(cudaStreamSynchronize(st));
        return VG_OK;
    }
    DeviceCtx *c = t_ts.ctx;
    std::lock_guard<std::mutex> lk(c->stage_mu);
    VG_TRY(ensure_stage(c));
    size_t off = 0;
    int i = 0;
    while (off < bytes) {
        const size_t n = bytes - off < kStageBytes ? bytes - off : kStageBytes;
        VG_CUDA(cudaEventSynchronize(c->stage_ev[i]));
        stage_copy(c->stage[i], (const char *)h_src + off, n);
        VG_CUDA(cudaMemcpyAsync((char *)d_dst + off, c->stage[i], n, cudaMemcpyHostToDevice, st));
        VG_CUDA(cudaEventRecord(c->stage_ev[i], st));
        off += n;
        i ^= 1;
    }
    VG_CUDA(cudaEventSynchronize(c->stage_ev[0]));
    VG_CUDA(cudaEventSynchronize(c->stage_ev[1]));
    return VG_OK;
}
vg_status staged_d2h(void *h_dst, const void *d_src, size_t bytes) {
    if (bytes == 0) return VG_OK;
    cudaStream_t st = stream();
    if (bytes < kSmallCopy || host_pinned(h_dst)) {
        VG_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
        VG_CUDA(cudaStreamSynchronize(st));
        return VG_OK;
    }
    DeviceCtx *c = t_ts.ctx;
    std::lock_guard<std::mutex> lk(c->stage_mu);
    VG_TRY(ensure_stage(c));
    size_t off = 0;
    while (off < bytes) {
        const size_t n = bytes - off < kStageBytes ? bytes - off : kStageBytes;
        VG_CUDA(cudaMemcpyAsync(c->stage[0], (const char *)d_src + off, n, cudaMemcpyDeviceToHost, st));
        VG_CUDA(cudaStreamSynchronize(st));
        memcpy((char *)h_dst + off, c->stage[0], n);
        off += n;
    }
    return VG_OK;
}

// Upload helper: host array → fresh device buffer.
template <class T>
static vg_status to_device(DevBuf &buf, const T *h, size_t count) {
    VG_TRY(buf.alloc(count * sizeof(T)));
    return staged_h2d(buf.p, h, count * sizeof(T));
}

// ------------------------------------------------------------------ index
struct Index {
    vg_index_desc d{};
    int device = 0;                  // the GPU that holds this index: every call on the handle runs there
    bool stream_set = false;         // vg_index_set_stream: calls on this handle run on the caller's stream
    cudaStream_t user_stream = nullptr;
    int variant = 0;
    int64_t code_row_bytes = 0;      // host/file layout
    int64_t dev_row_bytes = 0;       // device layout
    DevBuf codes, vectors, p0, p1, norms, ids;
    DevBuf pq_cb, pq_scales, pq_offsets, centroids, part_off, rotation;
    std::vector<uint32_t> h_part_off;
    int words32 = 0;
    const float *host_vectors = nullptr;   // vg_index_set_host_vectors: float32 rows stay in (page-locked) host memory, rerank reads them over PCIe
    const float *host_vectors_base = nullptr;   // the caller's host address of that region
    bool host_registered = false;          // the library page-locked the region itself (undone at close)
    bool int4_direct = false;        // vg_index_score on INT4: simd.Int4L2Distance instead of the precomputed-LUT path
    bool has_vectors = false, has_codes = false, has_ids = false;
    // tensor-core Flat filter state (vg_flat_tc.cu): squared row norms + their maximum, rebuilt after uploads
    DevBuf xn, xmax;
    bool xn_dirty = true;
    DevBuf x16;           // fp16 shadow of the vectors (x * 2^x16_exp, rows padded to a multiple of 64 dims) for the CTA-pair filter
    int x16_exp = 0;
    // decode-GEMM filter state of the quantized scans (vg_quant_tc.cu), rebuilt after code uploads
    qtc::Prepared qtc;
    bool qtc_dirty = true;
    // Set when a quarter or more of a batch failed the filter's certificate (tightly clustered data: thousands of rows
    // within the error bound of the k-th best).  Later batches then skip the doomed full first pass: the filter runs on a
    // 1/16 row prefix only, to get an upper bound of every query's k-th best score, and the threshold pass does the rest.
    std::atomic<int> qtc_hard{0};
    std::mutex prep_mu;   // lazy filter state (row norms, fp16 shadow, decode tables) is built once even if searches race
    size_t device_bytes() const {
        return codes.bytes + vectors.bytes + p0.bytes + p1.bytes + norms.bytes + ids.bytes + pq_cb.bytes + centroids.bytes;
    }
};
static std::unordered_map<uint64_t, Index *> g_indexes;
static uint64_t g_next_handle = 1;
#define VG_ENTER_IX(ix) VG_ENTER((ix)->device, (ix)->stream_set, (ix)->user_stream)

static Index *lookup(vg_index_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_indexes.find(h);
    return it == g_indexes.end() ? nullptr : it->second;
}

// float32 rows a rerank gathers from: the device copy, or the device alias of a page-locked host region
static const float *rerank_source(const Index *ix) { return ix->host_vectors ? ix->host_vectors : ix->vectors.as<float>(); }
static vg_status rerank_rows(const Index *ix, const float *d_queries, int64_t nq, const uint32_t *d_rows, int64_t r, int is_dot,
                             float *d_out, cudaStream_t st) {
    if (ix->host_vectors) return rerank_gather_host(ix->host_vectors, ix->d.rows, ix->d.dim, d_queries, nq, d_rows, r, is_dot, d_out, st);
    return rerank_gather(ix->vectors.as<float>(), ix->d.rows, ix->d.dim, d_queries, nq, d_rows, r, is_dot, d_out, st);
}

static int64_t host_code_bytes(const vg_index_desc &d) {
    switch (d.codec) {
        case VG_CODEC_SQ8: return d.dim;
        case VG_CODEC_INT4: return (d.dim + 1) / 2;
        case VG_CODEC_PQ:
        case VG_CODEC_OPQ: return d.pq_m;
        case VG_CODEC_BQ: return ((d.dim + 63) / 64) * 8;
        case VG_CODEC_RABITQ: return ((d.dim + 63) / 64) * 8 + 4;
        default: return 0;
    }
}

static CodecParams params_of(const Index &ix) {
    CodecParams cp;
    cp.codec = ix.d.codec;
    cp.variant = ix.variant;
    cp.dim = ix.d.dim;
    cp.row_bytes = ix.dev_row_bytes;
    cp.codes = ix.codes.as<uint8_t>();
    cp.vectors = ix.vectors.as<float>();
    cp.p0 = ix.p0.as<float>();
    cp.p1 = ix.p1.as<float>();
    cp.pq_m = (int)ix.d.pq_m;
    cp.pq_k = (int)ix.d.pq_k;
    cp.pq_dsub = ix.d.pq_m > 0 ? (int)(ix.d.dim / ix.d.pq_m) : 0;
    cp.pq_codebooks = ix.pq_cb.as<int8_t>();
    cp.pq_scales = ix.pq_scales.as<float>();
    cp.pq_offsets = ix.pq_offsets.as<float>();
    cp.norms = ix.norms.as<float>();
    cp.words32 = ix.words32;
    return cp;
}

}  // namespace vg

using namespace vg;

// =================================================================== C ABI
extern "C" {

const char *vg_last_error(void) { return t_error.c_str(); }
const char *vg_version(void) { return "vecgo_b200 0.1 (sm_100a)"; }
uint64_t vg_launch_count(void) { return g_launches.load(); }

vg_status vg_device_count(int32_t *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return cuda_fail(e, "cudaGetDeviceCount");
    }
    *count = n;
    return VG_OK;
}

vg_status vg_init(int32_t device) {
    // Selects `device` for the calling thread: calls without a handle (quantizer training, simd mirrors, vg_index_create)
    // made by this thread run there.  The first device any thread initialises is also the process default, which
    // threads that never call vg_init use.  Handles remember their device: calls on a handle run on its GPU whatever
    // the calling thread selected, so one process can serve several GPUs.
    DeviceCtx *ctx = nullptr;
    VG_TRY(get_ctx(device, &ctx));
    VG_CUDA(cudaSetDevice(device));
    t_ts.device = device;
    return VG_OK;
}
vg_status vg_synchronize(void) {
    // every stream the library owns on the calling thread's device, and the thread's own stream if it set one
    VG_ENTER();
    DeviceCtx *c = t_ts.ctx;
    std::vector<cudaStream_t> all;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        all = c->free_streams;
    }
    for (cudaStream_t s_ : all) VG_CUDA(cudaStreamSynchronize(s_));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return VG_OK;
}
vg_status vg_set_stream(uint64_t cuda_stream) {
    // 0 is a real stream handle (the legacy default stream, what torch.cuda.current_stream().cuda_stream returns for
    // PyTorch's default stream): the library must run ON it, otherwise its own non-blocking streams race with the
    // caller's work on stream 0.  ~0 switches back to library-owned streams.  The setting belongs to the calling THREAD.
    t_ts.user_stream_set = cuda_stream != ~0ull;
    t_ts.user_stream = t_ts.user_stream_set ? reinterpret_cast<cudaStream_t>(cuda_stream) : nullptr;
    return VG_OK;
}
vg_status vg_index_set_stream(vg_index_t idx, uint64_t cuda_stream) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    ix->stream_set = cuda_stream != ~0ull;
    ix->user_stream = ix->stream_set ? reinterpret_cast<cudaStream_t>(cuda_stream) : nullptr;
    return VG_OK;
}
vg_status vg_index_device(vg_index_t idx, int32_t *device) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    if (device) *device = ix->device;
    return VG_OK;
}
vg_status vg_dev_alloc(void **d_ptr, size_t bytes) {
    VG_ENTER();
    VG_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 16));
    return VG_OK;
}
vg_status vg_dev_free(void *d_ptr) {
    if (d_ptr) VG_CUDA(cudaFree(d_ptr));
    return VG_OK;
}
vg_status vg_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes) {
    VG_ENTER();
    return staged_h2d(d_dst, h_src, bytes);
}
vg_status vg_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes) {
    VG_ENTER();
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_dst, d_src, bytes);
}

// ----------------------------------------------------------- index lifecycle
vg_status vg_index_create_on(int32_t device, const vg_index_desc *desc, vg_index_t *out) {
    VG_ENTER(device);
    if (!desc || !out) return fail(VG_ERR_INVALID, "null argument");
    const vg_index_desc &d = *desc;
    if (d.dim <= 0 || d.rows < 0) return fail(VG_ERR_INVALID, "dim must be positive and rows non-negative");
    if (d.rows > 0xFFFFFFFEll) return fail(VG_ERR_INVALID, "rows exceed uint32 RowID space");
    // global ids are uint32 (searcher.InternalCandidate.RowID) and 0xFFFFFFFF is the empty sentinel of the merges
    if (d.row_base > 0xFFFFFFFEull || d.row_base + (uint64_t)d.rows > 0xFFFFFFFEull)
        return fail(VG_ERR_INVALID, "row_base + rows exceeds the uint32 RowID space");
    std::unique_ptr<Index> ix(new Index());
    ix->d = d;
    ix->device = t_ts.ctx->device;
    ix->code_row_bytes = host_code_bytes(d);
    ix->dev_row_bytes = ix->code_row_bytes;
    const int64_t rows = d.rows > 0 ? d.rows : 1;
    switch (d.codec) {
        case VG_CODEC_F32:
            break;
        case VG_CODEC_SQ8:
            if (!d.sq8_mins || !d.sq8_inv_scales) return fail(VG_ERR_STATE, "ScalarQuantizer not trained");
            VG_TRY(to_device(ix->p0, d.sq8_mins, (size_t)d.dim));
            VG_TRY(to_device(ix->p1, d.sq8_inv_scales, (size_t)d.dim));
            if (d.metric != VG_METRIC_L2) ix->variant = VG_VAR_GO_SCALAR;  // flat.Search: sq.DotProduct per row
            else if (d.dim % 64 == 0) ix->variant = VG_VAR_PERM;
            break;
        case VG_CODEC_INT4:
            if (!d.int4_min || !d.int4_diff) return fail(VG_ERR_STATE, "Int4Quantizer not trained");
            VG_TRY(to_device(ix->p0, d.int4_min, (size_t)d.dim));
            VG_TRY(to_device(ix->p1, d.int4_diff, (size_t)d.dim));
            if (d.dim % 256 == 0) ix->variant = VG_VAR_PERM;
            break;
        case VG_CODEC_PQ:
        case VG_CODEC_OPQ: {
            if (!d.pq_codebooks || !d.pq_scales || !d.pq_offsets) return fail(VG_ERR_STATE, "ProductQuantizer not trained");
            if (d.pq_m <= 0 || d.dim % d.pq_m != 0) return fail(VG_ERR_INVALID, "dimension must be divisible by numSubvectors");
            if (d.pq_k <= 0 || d.pq_k > 256) return fail(VG_ERR_INVALID, "numCentroids must be in 1..256");
            VG_TRY(to_device(ix->pq_cb, d.pq_codebooks, (size_t)(d.pq_m * d.pq_k * (d.dim / d.pq_m))));
            VG_TRY(to_device(ix->pq_scales, d.pq_scales, (size_t)d.pq_m));
            VG_TRY(to_device(ix->pq_offsets, d.pq_offsets, (size_t)d.pq_m));
            // thread-per-row ADC scan with a two-query float2 table (M*256*8 B of shared memory): K=256 only
            if (d.pq_k == 256 && d.pq_m <= 96) {  // 192 KB of tables + 32 KB of top-k state
                ix->variant = VG_VAR_PERM;
                ix->dev_row_bytes = (d.pq_m + 15) / 16 * 16;
            }
            if (d.codec == VG_CODEC_OPQ) {
                if (!d.opq_rotation || d.opq_block <= 0 || d.dim % d.opq_block != 0)
                    return fail(VG_ERR_INVALID, "OPQ needs block rotations with dim % block == 0");
                VG_TRY(to_device(ix->rotation, d.opq_rotation, (size_t)(d.dim * d.opq_block)));
            }
            break;
        }
        case VG_CODEC_BQ:
        case VG_CODEC_RABITQ: {
            const int64_t nbytes = ((d.dim + 63) / 64) * 8;
            ix->dev_row_bytes = (nbytes + 15) / 16 * 16;
            ix->words32 = (int)(ix->dev_row_bytes / 4);
            if (d.codec == VG_CODEC_RABITQ) VG_TRY(ix->norms.alloc_persistent((size_t)rows * 4));
            break;
        }
        default:
            return fail(VG_ERR_INVALID, "unknown codec");
    }
    if (ix->code_row_bytes > 0) {
        const bool pq_tiled = (d.codec == VG_CODEC_PQ || d.codec == VG_CODEC_OPQ) && (ix->variant & VG_VAR_PERM);
        const int64_t alloc_rows = pq_tiled ? (rows + 31) / 32 * 32 : rows;  // PQ tiles hold 32 rows
        VG_TRY(ix->codes.alloc_persistent((size_t)alloc_rows * ix->dev_row_bytes));
        if (pq_tiled) VG_CUDA(cudaMemsetAsync(ix->codes.p, 0, ix->codes.bytes, stream()));
    }
    if (d.num_partitions > 1) {
        if (!d.centroids || !d.partition_offsets) return fail(VG_ERR_INVALID, "partitioned index needs centroids and offsets");
        VG_TRY(to_device(ix->centroids, d.centroids, (size_t)(d.num_partitions * d.dim)));
        VG_TRY(to_device(ix->part_off, d.partition_offsets, (size_t)(d.num_partitions + 1)));
        ix->h_part_off.assign(d.partition_offsets, d.partition_offsets + d.num_partitions + 1);
    }
    // the descriptor's host pointers are not retained
    ix->d.sq8_mins = ix->d.sq8_inv_scales = ix->d.int4_min = ix->d.int4_diff = nullptr;
    ix->d.pq_codebooks = nullptr;
    ix->d.pq_scales = ix->d.pq_offsets = ix->d.opq_rotation = ix->d.centroids = nullptr;
    ix->d.partition_offsets = nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    const uint64_t h = g_next_handle++;
    g_indexes[h] = ix.release();
    *out = h;
    return VG_OK;
}

vg_status vg_index_create(const vg_index_desc *desc, vg_index_t *out) { return vg_index_create_on(-1, desc, out); }

vg_status vg_index_close(vg_index_t idx) {
    Index *ix = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_indexes.find(idx);
        if (it == g_indexes.end()) return fail(VG_ERR_STATE, "unknown or closed index handle");
        ix = it->second;
        g_indexes.erase(it);
    }
    VG_ENTER(ix->device);
    cudaDeviceSynchronize();  // searches still in flight on other streams of this device finish before the sections are freed
    if (ix->host_registered) cudaHostUnregister(const_cast<float *>(ix->host_vectors_base));
    delete ix;
    return VG_OK;
}

vg_status vg_index_info(vg_index_t idx, int64_t *rows, int64_t *dim, int64_t *code_bytes_per_row, int64_t *device_bytes) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    if (rows) *rows = ix->d.rows;
    if (dim) *dim = ix->d.dim;
    if (code_bytes_per_row) *code_bytes_per_row = ix->code_row_bytes;
    if (device_bytes) *device_bytes = (int64_t)ix->device_bytes();
    return VG_OK;
}

// Place rows [row0,row0+n) of codes (device source, host layout) into the index's device layout.
static vg_status place_codes(Index *ix, int64_t row0, int64_t n, const uint8_t *d_src) {
    cudaStream_t st = stream();
    uint8_t *dst = ix->codes.as<uint8_t>() + row0 * ix->dev_row_bytes;
    switch (ix->d.codec) {
        case VG_CODEC_SQ8:
            if (ix->variant & VG_VAR_PERM) return dev_permute_sq8(d_src, dst, n, ix->d.dim, ix->d.dim % 256 == 0 ? 16 : 4, st);
            break;
        case VG_CODEC_INT4:
            if (ix->variant & VG_VAR_PERM) return dev_permute_int4(d_src, dst, n, ix->code_row_bytes, st);
            break;
        case VG_CODEC_PQ:
        case VG_CODEC_OPQ:
            if (ix->variant & VG_VAR_PERM)
                return dev_permute_pq(d_src, ix->codes.as<uint8_t>(), row0, n, (int)ix->d.pq_m, (int)ix->dev_row_bytes, st);
            break;
        case VG_CODEC_BQ:
            return dev_split_sign(d_src, n, ix->code_row_bytes, ix->code_row_bytes, ix->dev_row_bytes, dst, nullptr, st);
        case VG_CODEC_RABITQ:
            return dev_split_sign(d_src, n, ix->code_row_bytes - 4, ix->code_row_bytes, ix->dev_row_bytes, dst,
                                  ix->norms.as<float>() + row0, st);
        default:
            break;
    }
    VG_CUDA(cudaMemcpyAsync(dst, d_src, (size_t)n * ix->code_row_bytes, cudaMemcpyDeviceToDevice, st));
    return VG_OK;
}

static vg_status ensure_vectors(Index *ix) {
    if (!ix->vectors.p) VG_TRY(ix->vectors.alloc_persistent((size_t)(ix->d.rows > 0 ? ix->d.rows : 1) * ix->d.dim * 4));
    return VG_OK;
}

vg_status vg_index_upload_dev(vg_index_t idx, int64_t row0, int64_t n, const void *d_codes, const float *d_vectors) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (row0 < 0 || n < 0 || row0 + n > ix->d.rows) return fail(VG_ERR_INVALID, "row range outside the index");
    if (n == 0) return VG_OK;
    if (d_codes) {
        if (ix->code_row_bytes == 0) return fail(VG_ERR_INVALID, "this index has no code section");
        VG_TRY(place_codes(ix, row0, n, (const uint8_t *)d_codes));
        ix->has_codes = true;
        ix->qtc_dirty = true;
    }
    if (d_vectors) {
        VG_TRY(ensure_vectors(ix));
        VG_CUDA(cudaMemcpyAsync(ix->vectors.as<float>() + row0 * ix->d.dim, d_vectors, (size_t)n * ix->d.dim * 4,
                                cudaMemcpyDeviceToDevice, stream()));
        ix->has_vectors = true;
        ix->xn_dirty = true;
    }
    return VG_OK;
}

vg_status vg_index_upload(vg_index_t idx, int64_t row0, int64_t n, const void *h_codes, const float *h_vectors) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (row0 < 0 || n < 0 || row0 + n > ix->d.rows) return fail(VG_ERR_INVALID, "row range outside the index");
    if (n == 0) return VG_OK;
    if (h_codes) {
        if (ix->code_row_bytes == 0) return fail(VG_ERR_INVALID, "this index has no code section");
        // chunks of <= 64 MiB go host → pinned → device scratch → (re-tiled) final place
        const int64_t chunk_rows = std::max<int64_t>(1, (64ll << 20) / ix->code_row_bytes);
        DevBuf tmp;
        VG_TRY(tmp.alloc((size_t)std::min(chunk_rows, n) * ix->code_row_bytes));
        for (int64_t r = 0; r < n; r += chunk_rows) {
            const int64_t c = std::min(chunk_rows, n - r);
            VG_TRY(staged_h2d(tmp.p, (const uint8_t *)h_codes + r * ix->code_row_bytes, (size_t)c * ix->code_row_bytes));
            VG_TRY(place_codes(ix, row0 + r, c, tmp.as<uint8_t>()));
            VG_CUDA(cudaStreamSynchronize(stream()));
        }
        ix->has_codes = true;
        ix->qtc_dirty = true;
    }
    if (h_vectors) {
        VG_TRY(ensure_vectors(ix));
        VG_TRY(staged_h2d(ix->vectors.as<float>() + row0 * ix->d.dim, h_vectors, (size_t)n * ix->d.dim * 4));
        ix->has_vectors = true;
        ix->xn_dirty = true;
    }
    return VG_OK;
}

// --------------------------------------------------------------- search
// Per-call statistics of the last search on the calling thread (the reference's searcher.FilterGateStats /
// model.QueryStats counters a Go caller fills from them, flat/segment.go:448-471,553-591).
static thread_local vg_search_stats t_stats;
static thread_local uint64_t t_rows_skipped = 0;   // rows of the blocks the current vg_index_search_blocks* call jumps over

// Lazy per-index filter state.  Built once even if searches race (prep_mu) and COMPLETE on the device before the lock is
// released: another thread's search runs on another stream and must not start before these kernels have finished.
static vg_status ensure_row_norms(Index *ix, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(ix->prep_mu);
    if (!ix->xn_dirty && ix->xn.p) return VG_OK;
    const int64_t rows = ix->d.rows;
    if (!ix->xn.p) VG_TRY(ix->xn.alloc_persistent((size_t)rows * 4));
    if (!ix->xmax.p) VG_TRY(ix->xmax.alloc_persistent(16));
    VG_CUDA(cudaMemsetAsync(ix->xmax.p, 0, 16, st));
    VG_TRY(tc::sqnorms(ix->vectors.as<float>(), rows, ix->d.dim, ix->d.dim, ix->xn.as<float>(), ix->xmax.as<unsigned int>(), st));
    if (rows >= 8192 && tc::pair_enabled()) {
        // fp16 shadow for the CTA-pair filter: |x_i| <= sqrt(max ||x||^2) is scaled below 2^12
        float xmax = 0.0f;
        VG_CUDA(cudaMemcpyAsync(&xmax, ix->xmax.p, 4, cudaMemcpyDeviceToHost, st));
        VG_CUDA(cudaStreamSynchronize(st));
        int e = 0;
        if (xmax > 0.0f && std::isfinite(xmax)) {
            int ex;
            std::frexp(std::sqrt((double)xmax), &ex);
            e = std::min(60, std::max(-60, 12 - ex));
        }
        const int dimp = (int)((ix->d.dim + 63) / 64 * 64);
        if (!ix->x16.p) VG_TRY(ix->x16.alloc_persistent((size_t)rows * dimp * 2));
        VG_TRY(tc::make_shadow16(ix->vectors.as<float>(), rows, ix->d.dim, dimp, e, ix->x16.p, st));
        ix->x16_exp = e;
    }
    VG_CUDA(cudaStreamSynchronize(st));
    ix->xn_dirty = false;
    return VG_OK;
}
static vg_status ensure_qtc(Index *ix, const CodecParams &cp, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(ix->prep_mu);
    if (!ix->qtc_dirty && ix->qtc.ready) return VG_OK;
    const vg_index_desc &d = ix->d;
    const bool pq = d.codec == VG_CODEC_PQ || d.codec == VG_CODEC_OPQ;
    const size_t np = pq ? (size_t)d.pq_m : (size_t)d.dim;
    std::vector<float> h0(np), h1(np);
    if (d.codec != VG_CODEC_RABITQ && d.codec != VG_CODEC_BQ) {  // the sign codes have no decode parameters
        VG_CUDA(cudaMemcpyAsync(h0.data(), pq ? ix->pq_scales.p : ix->p0.p, np * 4, cudaMemcpyDeviceToHost, st));
        VG_CUDA(cudaMemcpyAsync(h1.data(), pq ? ix->pq_offsets.p : ix->p1.p, np * 4, cudaMemcpyDeviceToHost, st));
        VG_CUDA(cudaStreamSynchronize(st));
    }
    VG_TRY(qtc::prepare(cp, d.rows, h0.data(), h1.data(), ix->qtc, st));
    VG_CUDA(cudaStreamSynchronize(st));
    ix->qtc_dirty = false;
    return VG_OK;
}

// Everything one search call derives from its queries: codec view, scan arguments, prepared sign words (BQ / RaBitQ),
// rotated queries (OPQ), probed partitions (IVF).  The buffers come from the stream-ordered pool: they are released in
// stream order when the struct dies, after the kernels that read them.
struct SearchTemps {
    CodecParams cp;
    ScanArgs a;
    DevBuf qwords, qnorms, probe, rotated, tight;
};
// BQ / RaBitQ: sign words of the queries (+ norms in simd.Dot order) in rows of words32 (zero padded).
static vg_status prep_sign(Index *ix, const float *d_queries, int64_t nq, DevBuf &qwords, DevBuf &qnorms, DevBuf &tight, cudaStream_t st) {
    const vg_index_desc &d = ix->d;
    VG_TRY(qwords.alloc((size_t)nq * ix->words32 * 4));
    if (d.codec == VG_CODEC_RABITQ) VG_TRY(qnorms.alloc((size_t)nq * 4));
    const float thr = d.codec == VG_CODEC_BQ ? d.bq_threshold : 0.0f;
    // prep writes ceil(dim/64)*2 words per query
    const int w_live = (int)(((d.dim + 63) / 64) * 2);
    if (w_live == ix->words32) return prep_sign_queries(d_queries, nq, d.dim, thr, qwords.as<uint32_t>(), qnorms.as<float>(), st);
    VG_CUDA(cudaMemsetAsync(qwords.p, 0, qwords.bytes, st));
    VG_TRY(tight.alloc((size_t)nq * w_live * 4));
    VG_TRY(prep_sign_queries(d_queries, nq, d.dim, thr, tight.as<uint32_t>(), qnorms.as<float>(), st));
    VG_CUDA(cudaMemcpy2DAsync(qwords.p, (size_t)ix->words32 * 4, tight.p, (size_t)w_live * 4, (size_t)w_live * 4, (size_t)nq,
                              cudaMemcpyDeviceToDevice, st));
    return VG_OK;
}
static vg_status make_temps(Index *ix, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes, const uint8_t *d_mask, uint32_t *d_rows,
                            float *d_scores, int32_t *d_counts, SearchTemps &t, cudaStream_t st) {
    const vg_index_desc &d = ix->d;
    t.cp = params_of(*ix);
    ScanArgs &a = t.a;
    a.queries = d_queries;
    a.nq = nq;
    a.rows = d.rows;
    a.k = (int)k;
    a.descending = d.metric != VG_METRIC_L2;
    a.is_dot = d.metric != VG_METRIC_L2;
    a.row_base = (uint32_t)d.row_base;
    a.mask = d_mask;
    a.out_rows = d_rows;
    a.out_scores = d_scores;
    a.out_counts = d_counts;
    if (d.codec == VG_CODEC_INT4 || d.codec == VG_CODEC_BQ || d.codec == VG_CODEC_RABITQ) {
        a.descending = 0;  // these scores are distances whatever the segment metric
        a.is_dot = 0;
    }
    if (d.codec == VG_CODEC_BQ || d.codec == VG_CODEC_RABITQ) {
        VG_TRY(prep_sign(ix, d_queries, nq, t.qwords, t.qnorms, t.tight, st));
        t.cp.q_words = t.qwords.as<uint32_t>();
        t.cp.q_norms = t.qnorms.as<float>();
    }
    if (d.codec == VG_CODEC_OPQ) {
        VG_TRY(t.rotated.alloc((size_t)nq * d.dim * 4));
        VG_TRY(dev_opq_rotate(d_queries, nq, d.dim, (int)d.opq_block, ix->rotation.as<float>(), t.rotated.as<float>(), st));  // opq.go:196-214
        a.queries = t.rotated.as<float>();
    }
    if (d.num_partitions > 1) {
        // kmeans.FindClosestCentroids per query (flat/segment.go:726-745)
        int64_t np = nprobes <= 0 ? 1 : nprobes;
        if (np > d.num_partitions) np = d.num_partitions;
        VG_TRY(t.probe.alloc((size_t)nq * np * 4));
        VG_TRY(dev_find_closest(a.queries, nq, d.dim, ix->centroids.as<float>(), d.num_partitions, np, d.metric, t.probe.as<int32_t>(), st));
        a.probe = t.probe.as<int32_t>();
        a.nprobe = (int)np;
        a.part_off = ix->part_off.as<uint32_t>();
        a.num_parts = (int)d.num_partitions;
    }
    return VG_OK;
}

static tc::SearchIO flat_io(Index *ix, const float *d_queries, int64_t nq, int64_t k, const uint8_t *d_mask, uint32_t *d_rows,
                            float *d_scores, int32_t *d_counts) {
    const vg_index_desc &d = ix->d;
    tc::SearchIO io;
    io.d_queries = d_queries;
    io.nq = nq;
    io.d_vectors = ix->vectors.as<float>();
    io.rows = d.rows;
    io.dim = d.dim;
    io.d_xn = ix->xn.as<float>();
    io.d_xmax_bits = ix->xmax.as<unsigned int>();
    io.d_x16 = ix->x16.p;
    io.x16_exp = ix->x16_exp;
    io.d_mask = d_mask;
    io.k = (int)k;
    io.is_dot = d.metric != VG_METRIC_L2;
    io.row_base = (uint32_t)d.row_base;
    io.d_rows = d_rows;
    io.d_scores = d_scores;
    io.d_counts = d_counts;
    return io;
}
static qtc::SearchIO quant_io(Index *ix, const ScanArgs &a) {
    qtc::SearchIO io;
    io.d_queries = a.queries;
    io.q_stride = a.q_stride;
    io.nq = a.nq;
    io.rows = ix->d.rows;
    io.d_mask = a.mask;
    io.k = a.k;
    io.row_base = a.row_base;
    io.d_rows = a.out_rows;
    io.d_scores = a.out_scores;
    io.d_counts = a.out_counts;
    return io;
}

enum { SEARCH_EXACT = 0, SEARCH_FLAT_TC = 1, SEARCH_QUANT_TC = 2 };

// IVF-partitioned segments scan only the probed partitions (scan_topk_partitioned); VECGO_IVF_GROUPED=0 selects the
// full scan with a per-row partition test (same results; kept for A/B measurements).
static std::atomic<int> g_ivf_grouped{-1};
static bool partition_grouping() {
    int v = g_ivf_grouped.load();
    if (v < 0) {
        const char *e = getenv("VECGO_IVF_GROUPED");
        v = (e && e[0] == '0') ? 0 : 1;
        g_ivf_grouped.store(v);
    }
    return v != 0;
}

// Which path a (shape, pointer alignment) takes; same answer in enqueue and resolve.
static int search_mode(Index *ix, const float *d_queries, int64_t nq, int64_t k, const uint8_t *d_mask, const CodecParams *cp_ready) {
    const vg_index_desc &d = ix->d;
    if (d.codec == VG_CODEC_F32) {
        if (!tc::enabled() || d.num_partitions > 1 || !ix->has_vectors) return SEARCH_EXACT;
        if (!tc::supported(d.dim, d.rows, nq, k)) return SEARCH_EXACT;
        if ((reinterpret_cast<uintptr_t>(d_queries) & 15) != 0) return SEARCH_EXACT;  // TMA needs 16-byte aligned bases
        if ((reinterpret_cast<uintptr_t>(d_mask) & 3) != 0) return SEARCH_EXACT;      // the filter reads the row bitmap as 32-bit words
        return SEARCH_FLAT_TC;
    }
    if (!cp_ready || !ix->has_codes) return SEARCH_EXACT;
    if (!qtc::supported(*cp_ready, d.metric, d.rows, nq, k, d.num_partitions)) return SEARCH_EXACT;
    if ((reinterpret_cast<uintptr_t>(d_mask) & 3) != 0) return SEARCH_EXACT;
    return SEARCH_QUANT_TC;
}

// Launches the whole search of one batch and returns without waiting for the device.  d_fail[q] = 1 marks a query whose
// tensor-core filter result carries no proof yet (search_resolve re-runs those); the exact scan leaves all flags 0.
static vg_status search_enqueue(Index *ix, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes, const uint8_t *d_mask,
                                uint32_t *d_rows, float *d_scores, int32_t *d_counts, int32_t *d_fail, int *mode_out) {
    const vg_index_desc &d = ix->d;
    if (d.codec == VG_CODEC_F32 ? !ix->has_vectors : !ix->has_codes) {
        if (d.rows > 0) return fail(VG_ERR_STATE, "index rows were never uploaded");
    }
    cudaStream_t st = stream();
    t_stats.queries += (uint64_t)nq;
    t_stats.distance_computations += (uint64_t)nq * ((uint64_t)d.rows - t_rows_skipped);
    if (d.codec == VG_CODEC_F32) {
        const int mode = search_mode(ix, d_queries, nq, k, d_mask, nullptr);
        *mode_out = mode;
        if (mode == SEARCH_FLAT_TC) {
            VG_TRY(ensure_row_norms(ix, st));
            t_stats.filter_queries += (uint64_t)nq;
            return tc::enqueue(flat_io(ix, d_queries, nq, k, d_mask, d_rows, d_scores, d_counts), tc::candidates_for(k, d.dim), d_fail, st);
        }
    }
    SearchTemps t;
    VG_TRY(make_temps(ix, d_queries, nq, k, nprobes, d_mask, d_rows, d_scores, d_counts, t, st));
    const int mode = d.codec == VG_CODEC_F32 ? SEARCH_EXACT : search_mode(ix, d_queries, nq, k, d_mask, &t.cp);
    *mode_out = mode;
    if (mode == SEARCH_QUANT_TC) {
        VG_TRY(ensure_qtc(ix, t.cp, st));
        t_stats.filter_queries += (uint64_t)nq;
        const int64_t rows_p = (d.rows / 16 + 255) / 256 * 256;
        if (ix->qtc_hard.load() && qtc::threshold_pass_possible(t.cp, d.rows, nq) &&
            qtc::supported(t.cp, d.metric, rows_p, nq, k, d.num_partitions)) {
            // hard data: filter a row prefix (any k real rows bound the k-th best from above), then ONE threshold pass over
            // all rows for the whole batch — everything enqueued, no host wait
            qtc::SearchIO io = quant_io(ix, t.a);
            DevBuf f1, kth;
            VG_TRY(f1.alloc((size_t)nq * 4));
            VG_TRY(kth.alloc((size_t)nq * 4));
            io.rows = rows_p;
            VG_TRY(qtc::enqueue(t.cp, ix->qtc, io, 1, f1.as<int32_t>(), st));
            VG_TRY(qtc::gather_kth(d_scores, d_counts, nullptr, nq, (int)k, kth.as<float>(), st));
            io.rows = d.rows;
            t_stats.threshold_pass_queries += (uint64_t)nq;
            return qtc::enqueue_threshold(t.cp, ix->qtc, io, kth.as<float>(), d_fail, st);
        }
        return qtc::enqueue(t.cp, ix->qtc, quant_io(ix, t.a), 1, d_fail, st);
    }
    if (d_fail) VG_CUDA(cudaMemsetAsync(d_fail, 0, (size_t)nq * 4, st));
    if (t.a.probe && partition_grouping()) return scan_topk_partitioned(t.cp, t.a, st);
    return scan_topk(t.cp, t.a, st);
}

// Second half: `bad` = the queries whose flag was set (read back by the caller at a point where it synchronises anyway).
// They get the filter's second chance (twice the candidate groups) and, failing that too, the exact CUDA-core scan.
static vg_status search_resolve(Index *ix, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes, const uint8_t *d_mask,
                                uint32_t *d_rows, float *d_scores, int32_t *d_counts, std::vector<int32_t> &bad, int mode) {
    if (bad.empty() || mode == SEARCH_EXACT) return VG_OK;
    const vg_index_desc &d = ix->d;
    cudaStream_t st = stream();
    SearchTemps t;
    if (mode == SEARCH_FLAT_TC) {
        t_stats.second_chance_queries += (uint64_t)bad.size();
        VG_TRY(tc::retry(flat_io(ix, d_queries, nq, k, d_mask, d_rows, d_scores, d_counts), tc::candidates_for(k, d.dim), bad, st));
        if (bad.empty()) return VG_OK;
        VG_TRY(make_temps(ix, d_queries, nq, k, nprobes, d_mask, d_rows, d_scores, d_counts, t, st));
    } else {
        VG_TRY(make_temps(ix, d_queries, nq, k, nprobes, d_mask, d_rows, d_scores, d_counts, t, st));
        if (nq >= 64 && (int64_t)bad.size() * 4 >= nq) ix->qtc_hard.store(1);   // later batches go straight to the threshold pass
        if (qtc::threshold_pass_possible(t.cp, d.rows, (int64_t)bad.size())) {
            t_stats.threshold_pass_queries += (uint64_t)bad.size();
            // Second chance = a THRESHOLD pass over the failed queries: the k-th best exact score the first pass found bounds
            // the true k-th best from above; every row whose filter score could still beat it is listed and scored exactly
            // (vg_quant_tc.cu).  Gather the failed queries (and their prepared sign words / norms) and that bound.
            const int64_t nb = (int64_t)bad.size();
            t_stats.second_chance_queries += (uint64_t)nb;
            DevBuf bidx, bq, bqw, bqn, brow, bsc, bcnt, bfail, bkth;
            VG_TRY(bidx.alloc((size_t)nb * 4));
            VG_TRY(bq.alloc((size_t)nb * d.dim * 4));
            VG_TRY(brow.alloc((size_t)nb * k * 4));
            VG_TRY(bsc.alloc((size_t)nb * k * 4));
            VG_TRY(bcnt.alloc((size_t)nb * 4));
            VG_TRY(bfail.alloc((size_t)nb * 4));
            VG_TRY(bkth.alloc((size_t)nb * 4));
            VG_CUDA(cudaMemcpyAsync(bidx.p, bad.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, st));
            VG_TRY(dev_gather_rows(t.a.queries, d.dim, bidx.as<int32_t>(), nb, d.dim, bq.as<float>(), st));
            VG_TRY(qtc::gather_kth(d_scores, d_counts, bidx.as<int32_t>(), nb, (int)k, bkth.as<float>(), st));
            CodecParams cps = t.cp;
            if (t.cp.q_words) {
                VG_TRY(bqw.alloc((size_t)nb * t.cp.words32 * 4));
                VG_TRY(dev_gather_rows(reinterpret_cast<const float *>(t.cp.q_words), t.cp.words32, bidx.as<int32_t>(), nb, t.cp.words32,
                                       bqw.as<float>(), st));
                cps.q_words = bqw.as<uint32_t>();
            }
            if (t.cp.q_norms) {
                VG_TRY(bqn.alloc((size_t)nb * 4));
                VG_TRY(dev_gather_rows(t.cp.q_norms, 1, bidx.as<int32_t>(), nb, 1, bqn.as<float>(), st));
                cps.q_norms = bqn.as<float>();
            }
            qtc::SearchIO io = quant_io(ix, t.a);
            io.d_queries = bq.as<float>();
            io.q_stride = 0;
            io.nq = nb;
            io.d_rows = brow.as<uint32_t>();
            io.d_scores = bsc.as<float>();
            io.d_counts = bcnt.as<int32_t>();
            VG_TRY(qtc::enqueue_threshold(cps, ix->qtc, io, bkth.as<float>(), bfail.as<int32_t>(), st));
            VG_TRY(dev_scatter_results(brow.as<uint32_t>(), bsc.as<float>(), bcnt.as<int32_t>(), bidx.as<int32_t>(), nb, k, d_rows, d_scores,
                                       d_counts, st));
            std::vector<int32_t> h_fail((size_t)nb);
            VG_CUDA(cudaMemcpyAsync(h_fail.data(), bfail.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
            VG_CUDA(cudaStreamSynchronize(st));
            std::vector<int32_t> still;
            for (int64_t j = 0; j < nb; j++)
                if (h_fail[(size_t)j]) still.push_back(bad[(size_t)j]);
            bad.swap(still);
        }
        qtc::count_fallbacks((uint64_t)bad.size());
        if (bad.empty()) return VG_OK;
    }
    // exact CUDA-core scan for the queries no certificate could clear
    t_stats.exact_rerun_queries += (uint64_t)bad.size();
    t_stats.distance_computations += (uint64_t)bad.size() * ((uint64_t)d.rows - t_rows_skipped);
    return scan_topk_subset(t.cp, t.a, bad, st);
}

static vg_status read_flags(const int32_t *d_fail, int64_t nq, std::vector<int32_t> &bad, cudaStream_t st) {
    std::vector<int32_t> h((size_t)nq);
    VG_CUDA(cudaMemcpyAsync(h.data(), d_fail, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    VG_CUDA(cudaStreamSynchronize(st));
    bad.clear();
    for (int64_t q = 0; q < nq; q++)
        if (h[(size_t)q]) bad.push_back((int32_t)q);
    return VG_OK;
}

// Host-synchronous search of device-resident queries: enqueue, read the flags, resolve.
static vg_status search_dev_impl(Index *ix, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes,
                                 const uint8_t *d_mask, uint32_t *d_rows, float *d_scores, int32_t *d_counts) {
    if (nq < 0 || k <= 0) return fail(VG_ERR_INVALID, "nq must be >= 0 and k > 0");
    if (nq == 0) return VG_OK;
    DevBuf failb;
    VG_TRY(failb.alloc((size_t)nq * 4));
    int mode = SEARCH_EXACT;
    VG_TRY(search_enqueue(ix, d_queries, nq, k, nprobes, d_mask, d_rows, d_scores, d_counts, failb.as<int32_t>(), &mode));
    if (mode == SEARCH_EXACT) return VG_OK;  // stream-ordered: nothing to resolve
    std::vector<int32_t> bad;
    VG_TRY(read_flags(failb.as<int32_t>(), nq, bad, stream()));
    return search_resolve(ix, d_queries, nq, k, nprobes, d_mask, d_rows, d_scores, d_counts, bad, mode);
}

vg_status vg_index_search_dev(vg_index_t idx, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes,
                              const uint8_t *d_row_mask, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    t_stats = vg_search_stats{};
    VG_TRY(search_dev_impl(ix, d_queries, nq, k, nprobes, d_row_mask, d_out_rows, d_out_scores, d_out_counts));
    return _vg_call.finish();
}

// ---------------------------------------------------------------- block-stat skipping
// flat.Segment.Search jumps over a whole BlockSize = 1024-row block when the block's field statistics cannot match the
// filter (flat/segment.go:524-541,613-630; filter.MatchesBlock / matchesFilterSet are host-side metadata logic).  The
// caller hands the verdicts over as a bitmap (bit b set = scan block b); only blocks that lie wholly inside the segment
// are ever skipped (i + BlockSize <= end).  On the device the verdicts are folded into the row bitmap; the tensor-core
// filters then drop every 256-row tile without an allowed row before it is fetched (vg_quant_tc.cu: tile skipping).
constexpr int64_t kBlockRows = 1024;
__global__ void __launch_bounds__(256) apply_block_mask_kernel(const uint8_t *row_mask, const uint8_t *block_keep, int64_t rows, int64_t words,
                                                               uint32_t *eff) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    const int64_t blk = (w * 32) / kBlockRows;
    const bool full = (blk + 1) * kBlockRows <= rows;
    const bool keep = !full || ((block_keep[blk >> 3] >> (blk & 7)) & 1);
    uint32_t v = 0;
    if (keep) {
        if (!row_mask) v = 0xFFFFFFFFu;
        else {
            const int64_t nbytes = (rows + 7) / 8;
            for (int b = 0; b < 4; b++)
                if (w * 4 + b < nbytes) v |= (uint32_t)row_mask[w * 4 + b] << (8 * b);
        }
    }
    eff[w] = v;
}
static vg_status effective_mask(Index *ix, const uint8_t *d_row_mask, const uint8_t *h_block_keep, DevBuf &eff, uint64_t *rows_skipped) {
    const int64_t rows = ix->d.rows;
    const int64_t full_blocks = rows / kBlockRows, words = (rows + 31) / 32;
    uint64_t skipped = 0;
    for (int64_t b = 0; b < full_blocks; b++)
        if (!((h_block_keep[b >> 3] >> (b & 7)) & 1)) skipped++;
    *rows_skipped = skipped * (uint64_t)kBlockRows;
    DevBuf keep;
    const size_t kb = (size_t)((rows / kBlockRows + 1 + 7) / 8);
    std::vector<uint8_t> h_keep(kb, 0xFF);   // bits past the caller's bitmap (the ragged last block) read as "scan"
    memcpy(h_keep.data(), h_block_keep, (size_t)((full_blocks + 7) / 8));
    VG_TRY(to_device(keep, h_keep.data(), kb));
    VG_TRY(eff.alloc((size_t)words * 4 + 16));
    apply_block_mask_kernel<<<(unsigned)((words + 255) / 256), 256, 0, stream()>>>(d_row_mask, keep.as<uint8_t>(), rows, words, eff.as<uint32_t>());
    VG_LAUNCHED();
    return VG_OK;
}

vg_status vg_index_search_blocks_dev(vg_index_t idx, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes, const uint8_t *d_row_mask,
                                     const uint8_t *h_block_keep, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts) {
    if (!h_block_keep) return vg_index_search_dev(idx, d_queries, nq, k, nprobes, d_row_mask, d_out_rows, d_out_scores, d_out_counts);
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    t_stats = vg_search_stats{};
    DevBuf eff;
    uint64_t skipped = 0;
    VG_TRY(effective_mask(ix, d_row_mask, h_block_keep, eff, &skipped));
    t_rows_skipped = skipped;
    const vg_status rc = search_dev_impl(ix, d_queries, nq, k, nprobes, eff.as<uint8_t>(), d_out_rows, d_out_scores, d_out_counts);
    t_rows_skipped = 0;
    if (rc != VG_OK) return rc;
    return _vg_call.finish();   // the folded bitmap goes back to the stream-ordered pool behind the kernels that read it
}

vg_status vg_index_search_blocks(vg_index_t idx, const float *h_queries, int64_t nq, int64_t k, int64_t nprobes, const uint8_t *h_row_mask,
                                 const uint8_t *h_block_keep, uint32_t *h_out_rows, float *h_out_scores, int32_t *h_out_counts) {
    if (!h_block_keep) return vg_index_search(idx, h_queries, nq, k, nprobes, h_row_mask, h_out_rows, h_out_scores, h_out_counts);
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    if (nq < 0 || k <= 0) return fail(VG_ERR_INVALID, "nq must be >= 0 and k > 0");
    if (nq == 0) return VG_OK;
    // fold the verdicts into the row bitmap on the host (rows / 8 bytes) and take the ordinary path
    const int64_t rows = ix->d.rows, nbytes = (rows + 7) / 8, full_blocks = rows / kBlockRows;
    std::vector<uint8_t> m((size_t)nbytes + 4, 0);
    if (h_row_mask) memcpy(m.data(), h_row_mask, (size_t)nbytes);
    else memset(m.data(), 0xFF, (size_t)nbytes);
    uint64_t skipped = 0;
    for (int64_t b = 0; b < full_blocks; b++)
        if (!((h_block_keep[b >> 3] >> (b & 7)) & 1)) {
            memset(m.data() + b * (kBlockRows / 8), 0, (size_t)(kBlockRows / 8));
            skipped++;
        }
    t_rows_skipped = skipped * (uint64_t)kBlockRows;
    const vg_status rc = vg_index_search(idx, h_queries, nq, k, nprobes, m.data(), h_out_rows, h_out_scores, h_out_counts);
    t_rows_skipped = 0;
    return rc;
}

vg_status vg_ivf_grouped_enable(int32_t on) {
    g_ivf_grouped.store(on != 0 ? 1 : 0);
    return VG_OK;
}

vg_status vg_quant_tc_i8_enable(int32_t on) {
    qtc::set_i8(on != 0);
    return VG_OK;
}
vg_status vg_quant_tc_i8_state(int32_t *on) {
    if (!on) return fail(VG_ERR_INVALID, "null argument");
    *on = qtc::i8_state() ? 1 : 0;
    return VG_OK;
}

vg_status vg_tile_skip_enable(int32_t on) {
    tiles::set_enabled(on != 0);
    return VG_OK;
}

vg_status vg_index_search_dev_async(vg_index_t idx, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes,
                                    const uint8_t *d_row_mask, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts,
                                    int32_t *d_unproven) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (_vg_call.leased()) return fail(VG_ERR_STATE, "asynchronous search needs a caller stream (vg_set_stream / vg_index_set_stream)");
    if (nq < 0 || k <= 0 || !d_unproven) return fail(VG_ERR_INVALID, "nq must be >= 0, k > 0 and d_unproven non-null");
    if (nq == 0) return VG_OK;
    t_stats = vg_search_stats{};
    int mode = SEARCH_EXACT;
    return search_enqueue(ix, d_queries, nq, k, nprobes, d_row_mask, d_out_rows, d_out_scores, d_out_counts, d_unproven, &mode);
}

vg_status vg_index_search_resolve(vg_index_t idx, const float *d_queries, int64_t nq, int64_t k, int64_t nprobes,
                                  const uint8_t *d_row_mask, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts,
                                  const int32_t *d_unproven, int64_t *n_resolved) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (n_resolved) *n_resolved = 0;
    if (nq <= 0) return VG_OK;
    if (k <= 0 || !d_unproven) return fail(VG_ERR_INVALID, "k must be > 0 and d_unproven non-null");
    t_stats = vg_search_stats{};
    std::vector<int32_t> bad;
    VG_TRY(read_flags(d_unproven, nq, bad, stream()));
    if (n_resolved) *n_resolved = (int64_t)bad.size();
    if (bad.empty()) return VG_OK;
    // the mode is a function of the shape and the pointers: recompute what enqueue chose
    int mode = SEARCH_EXACT;
    if (ix->d.codec == VG_CODEC_F32) {
        mode = search_mode(ix, d_queries, nq, k, d_row_mask, nullptr);
    } else {
        CodecParams cp = params_of(*ix);
        static const uint32_t dummy_words = 0;
        static const float dummy_norm = 0.0f;
        if (ix->d.codec == VG_CODEC_BQ || ix->d.codec == VG_CODEC_RABITQ) {  // supported() only checks that the query state exists
            cp.q_words = &dummy_words;
            cp.q_norms = &dummy_norm;
        }
        mode = search_mode(ix, d_queries, nq, k, d_row_mask, &cp);
    }
    VG_TRY(search_resolve(ix, d_queries, nq, k, nprobes, d_row_mask, d_out_rows, d_out_scores, d_out_counts, bad, mode));
    return _vg_call.finish();
}

vg_status vg_last_search_stats(vg_search_stats *out) {
    if (!out) return fail(VG_ERR_INVALID, "null argument");
    *out = t_stats;
    return VG_OK;
}

vg_status vg_quant_tc_stats(uint64_t *queries, uint64_t *fallbacks) {
    qtc::stats(queries, fallbacks);
    return VG_OK;
}

vg_status vg_pq_assign_tc_stats(uint64_t *pairs, uint64_t *fallback_pairs) {
    pqa::stats(pairs, fallback_pairs);
    return VG_OK;
}

vg_status vg_quant_tc_profile(int32_t enable, double *gemm_ms, uint64_t *gemm_launches) {
    qtc::profile(enable, gemm_ms, gemm_launches);
    return VG_OK;
}

vg_status vg_flat_tc_enable(int32_t on) {
    tc::set_enabled(on != 0);
    return VG_OK;
}
vg_status vg_flat_tc_stats(uint64_t *queries, uint64_t *fallbacks) {
    tc::stats(queries, fallbacks);
    return VG_OK;
}
vg_status vg_flat_tc_candidates(vg_index_t idx, const float *h_queries, int64_t nq, int64_t kc, uint32_t *h_groups, int32_t *h_counts,
                                float *h_tau, int64_t *group_rows) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (ix->d.codec != VG_CODEC_F32 || !ix->has_vectors) return fail(VG_ERR_STATE, "not a float32 index");
    if (kc < 1 || kc > 64) return fail(VG_ERR_INVALID, "kc must be in 1..64");
    if (!tc::supported(ix->d.dim, ix->d.rows, nq, 1)) return fail(VG_ERR_UNSUPPORTED, "shape not supported by the tensor-core filter");
    cudaStream_t st = stream();
    VG_TRY(ensure_row_norms(ix, st));
    DevBuf q, gids, gcnt, tau;
    VG_TRY(to_device(q, h_queries, (size_t)nq * ix->d.dim));
    VG_TRY(gids.alloc((size_t)nq * kc * 4));
    VG_TRY(gcnt.alloc((size_t)nq * 4));
    VG_TRY(tau.alloc((size_t)nq * 4));
    tc::FilterArgs f;
    f.d_queries = q.as<float>();
    f.d_vectors = ix->vectors.as<float>();
    f.d_xn = ix->xn.as<float>();
    f.nq = nq;
    f.rows = ix->d.rows;
    f.dim = ix->d.dim;
    f.kc = (int)kc;
    f.is_dot = ix->d.metric != VG_METRIC_L2;
    f.row_base = (uint32_t)ix->d.row_base;
    f.d_gids = gids.as<uint32_t>();
    f.d_gcnt = gcnt.as<int32_t>();
    f.d_tau = tau.as<float>();
    VG_TRY(tc::filter(f, st));
    VG_CUDA(cudaStreamSynchronize(st));
    VG_TRY(staged_d2h(h_groups, gids.p, (size_t)nq * kc * 4));
    VG_TRY(staged_d2h(h_counts, gcnt.p, (size_t)nq * 4));
    VG_TRY(staged_d2h(h_tau, tau.p, (size_t)nq * 4));
    if (group_rows) *group_rows = tc::group_rows(ix->d.rows, (int)kc);
    return VG_OK;
}

// Device -> host without a host wait when the destination is page-locked (or small: the driver stages it itself);
// everything else goes through the pinned ring, which synchronises.
static vg_status d2h_enqueue(void *h_dst, const void *d_src, size_t bytes) {
    if (bytes == 0) return VG_OK;
    if (bytes < kSmallCopy || host_pinned(h_dst)) {
        VG_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, stream()));
        return VG_OK;
    }
    return staged_d2h(h_dst, d_src, bytes);
}

vg_status vg_index_search(vg_index_t idx, const float *h_queries, int64_t nq, int64_t k, int64_t nprobes,
                          const uint8_t *h_row_mask, uint32_t *h_out_rows, float *h_out_scores, int32_t *h_out_counts) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (nq < 0 || k <= 0) return fail(VG_ERR_INVALID, "nq must be >= 0 and k > 0");
    if (nq == 0) return VG_OK;
    t_stats = vg_search_stats{};
    cudaStream_t st = stream();
    DevBuf q, mask, rows, scores, counts, failb;
    VG_TRY(to_device(q, h_queries, (size_t)nq * ix->d.dim));
    if (h_row_mask) VG_TRY(to_device(mask, h_row_mask, (size_t)((ix->d.rows + 7) / 8)));
    VG_TRY(rows.alloc((size_t)nq * k * 4));
    VG_TRY(scores.alloc((size_t)nq * k * 4));
    VG_TRY(counts.alloc((size_t)nq * 4));
    VG_TRY(failb.alloc((size_t)nq * 4));
    const uint8_t *d_mask = h_row_mask ? mask.as<uint8_t>() : nullptr;
    int mode = SEARCH_EXACT;
    VG_TRY(search_enqueue(ix, q.as<float>(), nq, k, nprobes, d_mask, rows.as<uint32_t>(), scores.as<float>(), counts.as<int32_t>(),
                          failb.as<int32_t>(), &mode));
    // results and certificate flags come back behind ONE synchronisation; the (rare) unproven queries are then
    // resolved and the results copied again
    std::vector<int32_t> h_fail;
    if (mode != SEARCH_EXACT) {
        h_fail.resize((size_t)nq);
        VG_CUDA(cudaMemcpyAsync(h_fail.data(), failb.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    }
    for (int pass = 0; pass < 2; pass++) {
        VG_TRY(d2h_enqueue(h_out_rows, rows.p, (size_t)nq * k * 4));
        VG_TRY(d2h_enqueue(h_out_scores, scores.p, (size_t)nq * k * 4));
        VG_TRY(d2h_enqueue(h_out_counts, counts.p, (size_t)nq * 4));
        VG_CUDA(cudaStreamSynchronize(st));
        if (pass == 1 || mode == SEARCH_EXACT) break;
        std::vector<int32_t> bad;
        for (int64_t i = 0; i < nq; i++)
            if (h_fail[(size_t)i]) bad.push_back((int32_t)i);
        if (bad.empty()) break;
        VG_TRY(search_resolve(ix, q.as<float>(), nq, k, nprobes, d_mask, rows.as<uint32_t>(), scores.as<float>(), counts.as<int32_t>(), bad,
                              mode));
    }
    return VG_OK;
}

vg_status vg_index_rerank_dev(vg_index_t idx, const float *d_queries, int64_t nq, const uint32_t *d_rows, int64_t r,
                              float *d_scores) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (!ix->has_vectors) return fail(VG_ERR_STATE, "index holds no float32 vectors to rerank against");
    VG_TRY(rerank_rows(ix, d_queries, nq, d_rows, r, ix->d.metric != VG_METRIC_L2, d_scores, stream()));
    return _vg_call.finish();  // stream-ordered on a caller stream; complete on return when the library chose the stream
}

vg_status vg_index_rerank(vg_index_t idx, const float *h_queries, int64_t nq, const uint32_t *h_rows, int64_t r, float *h_scores) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (nq <= 0 || r <= 0) return VG_OK;
    DevBuf q, rows, scores;
    VG_TRY(to_device(q, h_queries, (size_t)nq * ix->d.dim));
    VG_TRY(to_device(rows, h_rows, (size_t)nq * r));
    VG_TRY(scores.alloc((size_t)nq * r * 4));
    VG_TRY(vg_index_rerank_dev(idx, q.as<float>(), nq, rows.as<uint32_t>(), r, scores.as<float>()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_scores, scores.p, (size_t)nq * r * 4);
}

// Quantized gather scoring (DiskANN neighbour-list scoring, diskann/segment.go:511-588): the codec's own distance of every
// query to its r candidate rows, in the reference's arithmetic.  A float32 index scores exactly (Segment.Rerank).
vg_status vg_index_score_dev(vg_index_t idx, const float *d_queries, int64_t nq, const uint32_t *d_rows, int64_t r, float *d_scores) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (nq <= 0 || r <= 0) return VG_OK;
    const vg_index_desc &d = ix->d;
    if (d.codec == VG_CODEC_F32) return vg_index_rerank_dev(idx, d_queries, nq, d_rows, r, d_scores);
    if (!ix->has_codes) return fail(VG_ERR_STATE, "index rows were never uploaded");
    cudaStream_t st = stream();
    CodecParams cp = params_of(*ix);
    DevBuf qwords, qnorms, rotated, tight;
    const float *queries = d_queries;
    if (d.codec == VG_CODEC_RABITQ || d.codec == VG_CODEC_BQ) {
        VG_TRY(prep_sign(ix, d_queries, nq, qwords, qnorms, tight, st));
        cp.q_words = qwords.as<uint32_t>();
        cp.q_norms = qnorms.as<float>();
    }
    if (d.codec == VG_CODEC_OPQ) {
        VG_TRY(rotated.alloc((size_t)nq * d.dim * 4));
        VG_TRY(dev_opq_rotate(d_queries, nq, d.dim, (int)d.opq_block, ix->rotation.as<float>(), rotated.as<float>(), st));
        queries = rotated.as<float>();
    }
    VG_TRY(qtc::score_rows(cp, d.rows, queries, 0, nq, d_rows, r, d_scores, ix->int4_direct ? 0 : 1, st));
    return _vg_call.finish();  // the temporaries above are released in stream order
}

__global__ void iota_kernel(uint32_t *p, int64_t n);
// simd.SquaredL2Bounded per (query, candidate) pair against the index's float32 rows (the HNSW / DiskANN traversal's
// early-exit distance, distance/distance.go:24-31).
vg_status vg_index_l2_bounded_dev(vg_index_t idx, const float *d_queries, int64_t nq, const uint32_t *d_rows, int64_t r,
                                  const float *d_bounds, int32_t per_pair_bounds, float *d_scores, uint8_t *d_exceeded) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (nq <= 0 || r <= 0) return VG_OK;
    if (!ix->has_vectors) return fail(VG_ERR_STATE, "index holds no float32 vectors");
    VG_TRY(bounded_l2_gather(rerank_source(ix), ix->d.rows, ix->d.dim, d_queries, nq, d_rows, r, d_bounds, per_pair_bounds, d_scores,
                             d_exceeded, stream()));
    return _vg_call.finish();
}
vg_status vg_index_l2_bounded(vg_index_t idx, const float *h_queries, int64_t nq, const uint32_t *h_rows, int64_t r, const float *h_bounds,
                              int32_t per_pair_bounds, float *h_scores, uint8_t *h_exceeded) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (nq <= 0 || r <= 0) return VG_OK;
    DevBuf q, rows, bounds, scores, ex;
    VG_TRY(to_device(q, h_queries, (size_t)nq * ix->d.dim));
    VG_TRY(to_device(rows, h_rows, (size_t)nq * r));
    VG_TRY(to_device(bounds, h_bounds, (size_t)(per_pair_bounds ? nq * r : nq)));
    VG_TRY(scores.alloc((size_t)nq * r * 4));
    VG_TRY(ex.alloc((size_t)nq * r));
    VG_TRY(vg_index_l2_bounded_dev(idx, q.as<float>(), nq, rows.as<uint32_t>(), r, bounds.as<float>(), per_pair_bounds, scores.as<float>(),
                                   ex.as<uint8_t>()));
    VG_TRY(staged_d2h(h_scores, scores.p, (size_t)nq * r * 4));
    return staged_d2h(h_exceeded, ex.p, (size_t)nq * r);
}
// simd.SquaredL2Bounded mirror for explicit pairs (a[i], b[i]).
vg_status vg_simd_squared_l2_bounded(const float *h_a, const float *h_b, int64_t n_pairs, int64_t dim, const float *h_bounds, float *h_out,
                                     uint8_t *h_exceeded) {
    VG_ENTER();
    if (n_pairs <= 0) return VG_OK;
    if (dim <= 0) {
        for (int64_t i = 0; i < n_pairs; i++) {
            h_out[i] = 0.0f;
            h_exceeded[i] = 0.0f > h_bounds[i] ? 1 : 0;
        }
        return VG_OK;
    }
    DevBuf a, b, rows, bounds, out, ex;
    VG_TRY(to_device(a, h_a, (size_t)n_pairs * dim));
    VG_TRY(to_device(b, h_b, (size_t)n_pairs * dim));
    VG_TRY(to_device(bounds, h_bounds, (size_t)n_pairs));
    VG_TRY(rows.alloc((size_t)n_pairs * 4));
    VG_TRY(out.alloc((size_t)n_pairs * 4));
    VG_TRY(ex.alloc((size_t)n_pairs));
    iota_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, stream()>>>(rows.as<uint32_t>(), n_pairs);
    VG_LAUNCHED();
    VG_TRY(bounded_l2_gather(b.as<float>(), n_pairs, dim, a.as<float>(), n_pairs, rows.as<uint32_t>(), 1, bounds.as<float>(), 0, out.as<float>(),
                             ex.as<uint8_t>(), stream()));
    VG_TRY(staged_d2h(h_out, out.p, (size_t)n_pairs * 4));
    return staged_d2h(h_exceeded, ex.p, (size_t)n_pairs);
}
// simd.BuildInt4LookupTable (kernels.go:94-103): table[d*16 + q] = (float32(q) / 15) * diff[d] + min[d], each operation
// rounded to float32 on its own (Go never fuses) — parameter algebra, like vg_sq8_set_bounds.
vg_status vg_int4_build_lookup_table(const float *h_min, const float *h_diff, int64_t dim, float *h_table) {
    if (dim <= 0 || !h_min || !h_diff || !h_table) return fail(VG_ERR_INVALID, "dimension mismatch");
    for (int64_t d = 0; d < dim; d++)
        for (int q = 0; q < 16; q++) {
            volatile float n = (float)q / 15.0f;
            volatile float v = n * h_diff[d];
            volatile float t = v + h_min[d];
            h_table[d * 16 + q] = t;
        }
    return VG_OK;
}

// Float32 rows that stay in HOST memory (the mmap'd vector section of a segment whose rows do not fit next to the codes:
// 100M x 1536 float32 = 614 GB against 180 GB of HBM, SURVEY hard part 6).  The region is page-locked (unless the caller
// already did) and mapped into the device's address space; Segment.Rerank (flat/segment.go:754-781) then gathers the R
// candidate rows of every query over the host link instead of from HBM — R * dim * 4 bytes per query, same arithmetic,
// same bits.  The caller keeps the region valid and unchanged until vg_index_close (an mmap lives until Segment.Close).
vg_status vg_index_set_host_vectors(vg_index_t idx, const float *h_vectors, int64_t rows) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (!h_vectors || rows != ix->d.rows) return fail(VG_ERR_INVALID, "host vector region must cover every row of the index");
    if (ix->d.codec == VG_CODEC_F32) return fail(VG_ERR_UNSUPPORTED, "a float32 index scans its rows: they must be device-resident");
    if (ix->has_vectors) return fail(VG_ERR_STATE, "the index already holds float32 rows");
    const size_t bytes = (size_t)rows * (size_t)ix->d.dim * 4;
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, h_vectors) == cudaSuccess && at.type == cudaMemoryTypeHost;
    if (!pinned) {
        cudaGetLastError();
        cudaError_t e = cudaHostRegister(const_cast<float *>(h_vectors), bytes, cudaHostRegisterMapped | cudaHostRegisterReadOnly);
        if (e == cudaErrorNotSupported || e == cudaErrorInvalidValue) {
            cudaGetLastError();
            e = cudaHostRegister(const_cast<float *>(h_vectors), bytes, cudaHostRegisterMapped);
        }
        if (e != cudaSuccess) return cuda_fail(e, "cudaHostRegister");
        ix->host_registered = true;
    }
    void *dptr = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&dptr, const_cast<float *>(h_vectors), 0);
    if (e != cudaSuccess) {
        if (ix->host_registered) cudaHostUnregister(const_cast<float *>(h_vectors));
        ix->host_registered = false;
        return cuda_fail(e, "cudaHostGetDevicePointer");
    }
    ix->host_vectors = reinterpret_cast<const float *>(dptr);
    ix->host_vectors_base = h_vectors;
    ix->has_vectors = true;
    return VG_OK;
}

vg_status vg_index_set_int4_score_mode(vg_index_t idx, int32_t mode) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    if (mode != VG_INT4_SCORE_LUT && mode != VG_INT4_SCORE_DIRECT) return fail(VG_ERR_INVALID, "unknown INT4 score mode");
    ix->int4_direct = mode == VG_INT4_SCORE_DIRECT;
    return VG_OK;
}

vg_status vg_index_score(vg_index_t idx, const float *h_queries, int64_t nq, const uint32_t *h_rows, int64_t r, float *h_scores) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (nq <= 0 || r <= 0) return VG_OK;
    DevBuf q, rows, scores;
    VG_TRY(to_device(q, h_queries, (size_t)nq * ix->d.dim));
    VG_TRY(to_device(rows, h_rows, (size_t)nq * r));
    VG_TRY(scores.alloc((size_t)nq * r * 4));
    VG_TRY(vg_index_score_dev(idx, q.as<float>(), nq, rows.as<uint32_t>(), r, scores.as<float>()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_scores, scores.p, (size_t)nq * r * 4);
}

// exact scores → keys → per-query top-k (one list per query of length r)
__global__ void __launch_bounds__(256) rerank_keys_kernel(const uint32_t *rows, const float *scores, int64_t n, uint32_t row_base,
                                                          int descending, unsigned long long *keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = rows[i];
    keys[i] = (g == 0xFFFFFFFFu) ? VG_KEY_EMPTY : make_key(scores[i], g, descending != 0);
    (void)row_base;
}
__global__ void __launch_bounds__(256) local_rows_kernel(const uint32_t *rows, int64_t n, uint32_t row_base, uint32_t *local) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    local[i] = (rows[i] == 0xFFFFFFFFu) ? 0xFFFFFFFFu : rows[i] - row_base;
}

vg_status vg_index_search_rerank(vg_index_t idx, const float *h_queries, int64_t nq, int64_t r, int64_t k, uint32_t *h_out_rows,
                                 float *h_out_scores, int32_t *h_out_counts) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (nq <= 0) return VG_OK;
    if (r < k) r = k;
    if (!ix->has_vectors) return fail(VG_ERR_STATE, "index holds no float32 vectors to rerank against");
    cudaStream_t st = stream();
    DevBuf q, rows, local, scores, counts, exact, keys, orows, oscores, ocounts;
    VG_TRY(to_device(q, h_queries, (size_t)nq * ix->d.dim));
    VG_TRY(rows.alloc((size_t)nq * r * 4));
    VG_TRY(local.alloc((size_t)nq * r * 4));
    VG_TRY(scores.alloc((size_t)nq * r * 4));
    VG_TRY(counts.alloc((size_t)nq * 4));
    VG_TRY(exact.alloc((size_t)nq * r * 4));
    VG_TRY(keys.alloc((size_t)nq * r * 8));
    VG_TRY(orows.alloc((size_t)nq * k * 4));
    VG_TRY(oscores.alloc((size_t)nq * k * 4));
    VG_TRY(ocounts.alloc((size_t)nq * 4));
    VG_TRY(search_dev_impl(ix, q.as<float>(), nq, r, 0, nullptr, rows.as<uint32_t>(), scores.as<float>(), counts.as<int32_t>()));
    const int64_t n = nq * r;
    local_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rows.as<uint32_t>(), n, (uint32_t)ix->d.row_base,
                                                                   local.as<uint32_t>());
    VG_LAUNCHED();
    const int desc = ix->d.metric != VG_METRIC_L2;
    VG_TRY(rerank_rows(ix, q.as<float>(), nq, local.as<uint32_t>(), r, desc, exact.as<float>(), st));
    rerank_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rows.as<uint32_t>(), exact.as<float>(), n,
                                                                    (uint32_t)ix->d.row_base, desc, keys.as<unsigned long long>());
    VG_LAUNCHED();
    VG_TRY(launch_merge_keys(keys.as<unsigned long long>(), 1, nq, r, 0, r, desc != 0, k, orows.as<uint32_t>(), oscores.as<float>(),
                             ocounts.as<int32_t>(), st));
    VG_CUDA(cudaStreamSynchronize(st));
    VG_TRY(staged_d2h(h_out_rows, orows.p, (size_t)nq * k * 4));
    VG_TRY(staged_d2h(h_out_scores, oscores.p, (size_t)nq * k * 4));
    VG_TRY(staged_d2h(h_out_counts, ocounts.p, (size_t)nq * 4));
    return VG_OK;
}

// --------------------------------------------------------------- top-k merge
vg_status vg_topk_merge_dev(const uint32_t *d_rows, const float *d_scores, int64_t lists, int64_t nq, int64_t k_in,
                            int32_t descending, int64_t k_out, uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts) {
    VG_ENTER();
    if (lists <= 0 || nq < 0 || k_in <= 0 || k_out <= 0) return fail(VG_ERR_INVALID, "bad merge shape");
    VG_TRY(launch_merge_pairs(d_rows, d_scores, lists, nq, k_in, descending != 0, k_out, d_out_rows, d_out_scores, d_out_counts,
                              stream()));
    return _vg_call.finish();
}
// Packed exchange format of the sharded search (one 8-byte key per candidate): pack this shard's result, all-gather the
// keys (ONE collective instead of one for rows and one for scores), merge the gathered [lists][nq][k_in] keys directly.
vg_status vg_topk_pack_dev(const uint32_t *d_rows, const float *d_scores, int64_t n, int32_t descending, uint64_t *d_keys) {
    VG_ENTER();
    if (n < 0 || !d_keys) return fail(VG_ERR_INVALID, "bad pack shape");
    VG_TRY(launch_pack_keys(d_rows, d_scores, n, descending != 0, reinterpret_cast<unsigned long long *>(d_keys), stream()));
    return _vg_call.finish();
}
vg_status vg_topk_merge_keys_dev(const uint64_t *d_keys, int64_t lists, int64_t nq, int64_t k_in, int32_t descending, int64_t k_out,
                                 uint32_t *d_out_rows, float *d_out_scores, int32_t *d_out_counts) {
    VG_ENTER();
    if (lists <= 0 || nq < 0 || k_in <= 0 || k_out <= 0) return fail(VG_ERR_INVALID, "bad merge shape");
    VG_TRY(launch_merge_keys(reinterpret_cast<const unsigned long long *>(d_keys), lists, nq, k_in, nq * k_in, k_in, descending != 0, k_out,
                             d_out_rows, d_out_scores, d_out_counts, stream()));
    return _vg_call.finish();
}
vg_status vg_topk_merge(const uint32_t *h_rows, const float *h_scores, int64_t lists, int64_t nq, int64_t k_in,
                        int32_t descending, int64_t k_out, uint32_t *h_out_rows, float *h_out_scores, int32_t *h_out_counts) {
    VG_ENTER();
    if (lists <= 0 || nq < 0 || k_in <= 0 || k_out <= 0) return fail(VG_ERR_INVALID, "bad merge shape");
    if (nq == 0) return VG_OK;
    DevBuf r, s, orow, osc, ocnt;
    VG_TRY(to_device(r, h_rows, (size_t)(lists * nq * k_in)));
    VG_TRY(to_device(s, h_scores, (size_t)(lists * nq * k_in)));
    VG_TRY(orow.alloc((size_t)nq * k_out * 4));
    VG_TRY(osc.alloc((size_t)nq * k_out * 4));
    VG_TRY(ocnt.alloc((size_t)nq * 4));
    VG_TRY(launch_merge_pairs(r.as<uint32_t>(), s.as<float>(), lists, nq, k_in, descending != 0, k_out, orow.as<uint32_t>(),
                              osc.as<float>(), ocnt.as<int32_t>(), stream()));
    VG_TRY(staged_d2h(h_out_rows, orow.p, (size_t)nq * k_out * 4));
    VG_TRY(staged_d2h(h_out_scores, osc.p, (size_t)nq * k_out * 4));
    return staged_d2h(h_out_counts, ocnt.p, (size_t)nq * 4);
}

// ------------------------------------------------------------ simd mirrors
static vg_status dense_host(CodecParams cp, const float *h_q, int64_t nq, int64_t qfloats, int64_t n, int is_dot, float *h_out) {
    DevBuf q, out;
    if (h_q) VG_TRY(to_device(q, h_q, (size_t)nq * qfloats));
    VG_TRY(out.alloc((size_t)nq * n * 4));
    VG_TRY(scan_dense(cp, q.as<float>(), nq, n, is_dot, out.as<float>(), stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_out, out.p, (size_t)nq * n * 4);
}

static vg_status f32_dense(const float *h_q, int64_t nq, const float *h_t, int64_t n, int64_t dim, int is_dot, int variant,
                           float *h_out) {
    VG_ENTER();
    if (dim < 0 || nq < 0 || n < 0) return fail(VG_ERR_INVALID, "negative size");
    if (nq == 0 || n == 0) return VG_OK;
    if (dim == 0) {
        memset(h_out, 0, (size_t)nq * n * 4);
        return VG_OK;
    }
    DevBuf t;
    VG_TRY(to_device(t, h_t, (size_t)n * dim));
    CodecParams cp;
    cp.codec = VG_CODEC_F32;
    cp.variant = variant;
    cp.dim = dim;
    cp.vectors = t.as<float>();
    return dense_host(cp, h_q, nq, dim, n, is_dot, h_out);
}
vg_status vg_simd_dot_batch(const float *h_q, int64_t nq, const float *h_t, int64_t n, int64_t dim, float *h_out) {
    return f32_dense(h_q, nq, h_t, n, dim, 1, VG_VAR_BATCH, h_out);
}
vg_status vg_simd_squared_l2_batch(const float *h_q, int64_t nq, const float *h_t, int64_t n, int64_t dim, float *h_out) {
    return f32_dense(h_q, nq, h_t, n, dim, 0, VG_VAR_BATCH, h_out);
}

// per-pair Dot / SquaredL2: rerank kernel with rows[i] = i, one "query" per pair
__global__ void iota_kernel(uint32_t *p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}
static vg_status pair_host(const float *h_a, const float *h_b, int64_t n_pairs, int64_t dim, int is_dot, float *h_out) {
    VG_ENTER();
    if (n_pairs <= 0) return VG_OK;
    if (dim <= 0) {
        memset(h_out, 0, (size_t)n_pairs * 4);
        return VG_OK;
    }
    DevBuf a, b, rows, out;
    VG_TRY(to_device(a, h_a, (size_t)n_pairs * dim));
    VG_TRY(to_device(b, h_b, (size_t)n_pairs * dim));
    VG_TRY(rows.alloc((size_t)n_pairs * 4));
    VG_TRY(out.alloc((size_t)n_pairs * 4));
    iota_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, stream()>>>(rows.as<uint32_t>(), n_pairs);
    VG_LAUNCHED();
    VG_TRY(rerank_gather(b.as<float>(), n_pairs, dim, a.as<float>(), n_pairs, rows.as<uint32_t>(), 1, is_dot, out.as<float>(),
                         stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_out, out.p, (size_t)n_pairs * 4);
}
vg_status vg_simd_dot(const float *h_a, const float *h_b, int64_t n_pairs, int64_t dim, float *h_out) {
    return pair_host(h_a, h_b, n_pairs, dim, 1, h_out);
}
vg_status vg_simd_squared_l2(const float *h_a, const float *h_b, int64_t n_pairs, int64_t dim, float *h_out) {
    return pair_host(h_a, h_b, n_pairs, dim, 0, h_out);
}

vg_status vg_simd_sq8u_l2_batch(const float *h_q, int64_t nq, const uint8_t *h_codes, int64_t n, int64_t dim, const float *h_mins,
                                const float *h_inv, float *h_out) {
    VG_ENTER();
    if (nq <= 0 || n <= 0) return VG_OK;
    if (dim <= 0) {
        memset(h_out, 0, (size_t)nq * n * 4);
        return VG_OK;
    }
    DevBuf c, mn, iv;
    VG_TRY(to_device(c, h_codes, (size_t)n * dim));
    VG_TRY(to_device(mn, h_mins, (size_t)dim));
    VG_TRY(to_device(iv, h_inv, (size_t)dim));
    CodecParams cp;
    cp.codec = VG_CODEC_SQ8;
    cp.dim = dim;
    cp.row_bytes = dim;
    cp.codes = c.as<uint8_t>();
    cp.p0 = mn.as<float>();
    cp.p1 = iv.as<float>();
    return dense_host(cp, h_q, nq, dim, n, 0, h_out);
}
vg_status vg_simd_int4_l2_batch(const float *h_q, int64_t nq, const uint8_t *h_codes, int64_t n, int64_t dim, const float *h_min,
                                const float *h_diff, float *h_out) {
    VG_ENTER();
    if (nq <= 0 || n <= 0) return VG_OK;
    if (dim <= 0) {
        memset(h_out, 0, (size_t)nq * n * 4);
        return VG_OK;
    }
    const int64_t cs = (dim + 1) / 2;
    DevBuf c, mn, df;
    VG_TRY(to_device(c, h_codes, (size_t)n * cs));
    VG_TRY(to_device(mn, h_min, (size_t)dim));
    VG_TRY(to_device(df, h_diff, (size_t)dim));
    CodecParams cp;
    cp.codec = VG_CODEC_INT4;
    cp.dim = dim;
    cp.row_bytes = cs;
    cp.codes = c.as<uint8_t>();
    cp.p0 = mn.as<float>();
    cp.p1 = df.as<float>();
    return dense_host(cp, h_q, nq, dim, n, 0, h_out);
}
vg_status vg_simd_pq_adc_lookup(const float *h_tables, int64_t nq, const uint8_t *h_codes, int64_t n, int64_t m, float *h_out) {
    VG_ENTER();
    if (nq <= 0 || n <= 0) return VG_OK;
    if (m <= 0) {
        memset(h_out, 0, (size_t)nq * n * 4);
        return VG_OK;
    }
    DevBuf c, t;
    VG_TRY(to_device(c, h_codes, (size_t)n * m));
    VG_TRY(to_device(t, h_tables, (size_t)nq * m * 256));
    CodecParams cp;
    cp.codec = VG_CODEC_PQ;
    cp.dim = m;
    cp.row_bytes = m;
    cp.codes = c.as<uint8_t>();
    cp.pq_m = (int)m;
    cp.pq_k = 256;
    cp.pq_tables = t.as<float>();
    return dense_host(cp, nullptr, nq, 0, n, 0, h_out);
}
vg_status vg_simd_hamming(const uint8_t *h_q, int64_t nq, const uint8_t *h_codes, int64_t n, int64_t nbytes, int32_t *h_out) {
    VG_ENTER();
    if (nq <= 0 || n <= 0) return VG_OK;
    if (nbytes <= 0) {
        memset(h_out, 0, (size_t)nq * n * 4);
        return VG_OK;
    }
    DevBuf q, c, out;
    VG_TRY(to_device(q, h_q, (size_t)nq * nbytes));
    VG_TRY(to_device(c, h_codes, (size_t)n * nbytes));
    VG_TRY(out.alloc((size_t)nq * n * 4));
    VG_TRY(hamming_dense(q.as<uint8_t>(), nq, c.as<uint8_t>(), n, nbytes, out.as<int32_t>(), stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_out, out.p, (size_t)nq * n * 4);
}
vg_status vg_simd_scale(float *h_a, int64_t n, float scalar) {
    VG_ENTER();
    if (n <= 0) return VG_OK;
    DevBuf a;
    VG_TRY(to_device(a, h_a, (size_t)n));
    VG_TRY(dev_scale(a.as<float>(), n, scalar, stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_a, a.p, (size_t)n * 4);
}
vg_status vg_normalize_l2(float *h_vecs, int64_t n, int64_t dim, uint8_t *h_ok) {
    VG_ENTER();
    if (n <= 0) return VG_OK;
    if (dim <= 0) {
        if (h_ok) memset(h_ok, 0, (size_t)n);
        return VG_OK;
    }
    DevBuf v, ok;
    VG_TRY(to_device(v, h_vecs, (size_t)n * dim));
    VG_TRY(ok.alloc((size_t)n));
    VG_TRY(dev_normalize(v.as<float>(), n, dim, ok.as<uint8_t>(), stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    VG_TRY(staged_d2h(h_vecs, v.p, (size_t)n * dim * 4));
    if (h_ok) VG_TRY(staged_d2h(h_ok, ok.p, (size_t)n));
    return VG_OK;
}

// ------------------------------------------------------------ quantizers
vg_status vg_sq8_set_bounds(const float *h_mins, const float *h_maxs, int64_t dim, float *h_scales, float *h_inv) {
    // quantizer.go:51-75 — parameter algebra only (single rounded float32 ops)
    if (dim <= 0) return fail(VG_ERR_INVALID, "dimension mismatch");
    for (int64_t i = 0; i < dim; i++) {
        volatile float diff = h_maxs[i] - h_mins[i];
        if (diff < 1e-9f) {
            h_scales[i] = 0;
            h_inv[i] = 0;
        } else {
            volatile float s = 255.0f / diff, v = diff / 255.0f;
            h_scales[i] = s;
            h_inv[i] = v;
        }
    }
    return VG_OK;
}
vg_status vg_sq8_train(const float *h_vecs, int64_t n, int64_t dim, float *h_mins, float *h_maxs, float *h_scales, float *h_inv) {
    VG_ENTER();
    if (n <= 0) return fail(VG_ERR_INVALID, "no vectors provided for training");
    if (dim <= 0) return fail(VG_ERR_INVALID, "vector dimension mismatch");
    DevBuf v, mm;
    VG_TRY(to_device(v, h_vecs, (size_t)n * dim));
    VG_TRY(mm.alloc((size_t)dim * 8));
    VG_TRY(dev_minmax(v.as<float>(), n, dim, mm.as<float>(), mm.as<float>() + dim, stream()));
    VG_TRY(staged_d2h(h_mins, mm.p, (size_t)dim * 4));
    VG_TRY(staged_d2h(h_maxs, mm.as<float>() + dim, (size_t)dim * 4));
    for (int64_t i = 0; i < dim; i++) {  // quantizer.go:166-176
        if (h_mins[i] == h_maxs[i]) {
            volatile float t = h_mins[i] + 1e-6f;
            h_maxs[i] = t;
        }
        volatile float range = h_maxs[i] - h_mins[i];
        volatile float s = 255.0f / range, iv = range / 255.0f;
        h_scales[i] = s;
        h_inv[i] = iv;
    }
    return VG_OK;
}
vg_status vg_sq8_encode(const float *h_vecs, int64_t n, int64_t dim, const float *h_mins, const float *h_maxs, const float *h_scales,
                        uint8_t *h_codes) {
    VG_ENTER();
    if (!h_mins || !h_maxs || !h_scales) return fail(VG_ERR_STATE, "ScalarQuantizer not trained");
    if (n <= 0) return VG_OK;
    DevBuf v, mn, mx, sc, out;
    VG_TRY(to_device(v, h_vecs, (size_t)n * dim));
    VG_TRY(to_device(mn, h_mins, (size_t)dim));
    VG_TRY(to_device(mx, h_maxs, (size_t)dim));
    VG_TRY(to_device(sc, h_scales, (size_t)dim));
    VG_TRY(out.alloc((size_t)n * dim));
    VG_TRY(dev_sq8_encode(v.as<float>(), n, dim, mn.as<float>(), mx.as<float>(), sc.as<float>(), out.as<uint8_t>(), stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_codes, out.p, (size_t)n * dim);
}
vg_status vg_sq8_decode(const uint8_t *h_codes, int64_t n, int64_t dim, const float *h_mins, const float *h_inv, float *h_vecs) {
    VG_ENTER();
    if (!h_mins || !h_inv) return fail(VG_ERR_STATE, "ScalarQuantizer not trained");
    if (n <= 0) return VG_OK;
    DevBuf c, mn, iv, out;
    VG_TRY(to_device(c, h_codes, (size_t)n * dim));
    VG_TRY(to_device(mn, h_mins, (size_t)dim));
    VG_TRY(to_device(iv, h_inv, (size_t)dim));
    VG_TRY(out.alloc((size_t)n * dim * 4));
    VG_TRY(dev_sq8_decode(c.as<uint8_t>(), n, dim, mn.as<float>(), iv.as<float>(), out.as<float>(), stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_vecs, out.p, (size_t)n * dim * 4);
}
vg_status vg_int4_train(const float *h_vecs, int64_t n, int64_t dim, float *h_min, float *h_diff) {
    VG_ENTER();
    if (n <= 0) return fail(VG_ERR_INVALID, "no vectors provided for training");
    DevBuf v, mm;
    VG_TRY(to_device(v, h_vecs, (size_t)n * dim));
    VG_TRY(mm.alloc((size_t)dim * 8));
    VG_TRY(dev_minmax(v.as<float>(), n, dim, mm.as<float>(), mm.as<float>() + dim, stream()));
    std::vector<float> mx((size_t)dim);
    VG_TRY(staged_d2h(h_min, mm.p, (size_t)dim * 4));
    VG_TRY(staged_d2h(mx.data(), mm.as<float>() + dim, (size_t)dim * 4));
    for (int64_t i = 0; i < dim; i++) {  // int4.go:53-59
        volatile float df = mx[(size_t)i] - h_min[i];
        h_diff[i] = (df == 0) ? 1.0f : (float)df;
    }
    return VG_OK;
}
vg_status vg_int4_encode(const float *h_vecs, int64_t n, int64_t dim, const float *h_min, const float *h_diff, uint8_t *h_codes) {
    VG_ENTER();
    if (n <= 0) return VG_OK;
    const int64_t cs = (dim + 1) / 2;
    DevBuf v, mn, df, out;
    VG_TRY(to_device(v, h_vecs, (size_t)n * dim));
    VG_TRY(to_device(mn, h_min, (size_t)dim));
    VG_TRY(to_device(df, h_diff, (size_t)dim));
    VG_TRY(out.alloc((size_t)n * cs));
    VG_TRY(dev_int4_encode(v.as<float>(), n, dim, mn.as<float>(), df.as<float>(), out.as<uint8_t>(), stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_codes, out.p, (size_t)n * cs);
}
vg_status vg_int4_decode(const uint8_t *h_codes, int64_t n, int64_t dim, const float *h_min, const float *h_diff, float *h_vecs) {
    VG_ENTER();
    if (n <= 0) return VG_OK;
    const int64_t cs = (dim + 1) / 2;
    DevBuf c, mn, df, out;
    VG_TRY(to_device(c, h_codes, (size_t)n * cs));
    VG_TRY(to_device(mn, h_min, (size_t)dim));
    VG_TRY(to_device(df, h_diff, (size_t)dim));
    VG_TRY(out.alloc((size_t)n * dim * 4));
    VG_TRY(dev_int4_decode(c.as<uint8_t>(), n, dim, mn.as<float>(), df.as<float>(), out.as<float>(), stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_vecs, out.p, (size_t)n * dim * 4);
}
vg_status vg_bq_train(const float *h_vecs, int64_t n, int64_t dim, float *h_threshold) {
    VG_ENTER();
    if (n <= 0) return fail(VG_ERR_INVALID, "no vectors provided for training");
    DevBuf v;
    VG_TRY(to_device(v, h_vecs, (size_t)n * dim));
    double sum = 0;
    VG_TRY(dev_mean_f64(v.as<float>(), n * dim, &sum, stream()));
    *h_threshold = (float)(sum / (double)(n * dim));
    return VG_OK;
}
static vg_status sign_encode_host(const float *h_vecs, int64_t n, int64_t dim, float thr, bool with_norm, uint8_t *h_codes) {
    VG_ENTER();
    if (n <= 0) return VG_OK;
    const int64_t stride = ((dim + 63) / 64) * 8 + (with_norm ? 4 : 0);
    DevBuf v, out;
    VG_TRY(to_device(v, h_vecs, (size_t)n * dim));
    VG_TRY(out.alloc((size_t)n * stride));
    VG_TRY(dev_sign_encode(v.as<float>(), n, dim, thr, with_norm, out.as<uint8_t>(), stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_codes, out.p, (size_t)n * stride);
}
vg_status vg_bq_encode(const float *h_vecs, int64_t n, int64_t dim, float threshold, uint8_t *h_codes) {
    return sign_encode_host(h_vecs, n, dim, threshold, false, h_codes);
}
vg_status vg_rabitq_encode(const float *h_vecs, int64_t n, int64_t dim, uint8_t *h_codes) {
    return sign_encode_host(h_vecs, n, dim, 0.0f, true, h_codes);
}

struct PqDev {
    DevBuf cb, sc, of;
};
static vg_status pq_params(PqDev &p, int64_t dim, int64_t m, int64_t k, const int8_t *h_cb, const float *h_sc, const float *h_of) {
    if (m <= 0 || dim <= 0 || dim % m != 0) return fail(VG_ERR_INVALID, "dimension must be divisible by numSubvectors");
    if (k <= 0 || k > 256) return fail(VG_ERR_INVALID, "numCentroids must be <= 256 for uint8 encoding");
    if (!h_cb || !h_sc || !h_of) return fail(VG_ERR_STATE, "ProductQuantizer not trained");
    VG_TRY(to_device(p.cb, h_cb, (size_t)(m * k * (dim / m))));
    VG_TRY(to_device(p.sc, h_sc, (size_t)m));
    return to_device(p.of, h_of, (size_t)m);
}
vg_status vg_pq_encode(const float *h_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, const int8_t *h_cb, const float *h_sc,
                       const float *h_of, uint8_t *h_codes) {
    VG_ENTER();
    PqDev p;
    VG_TRY(pq_params(p, dim, m, k, h_cb, h_sc, h_of));
    if (n <= 0) return VG_OK;
    DevBuf v, out;
    VG_TRY(to_device(v, h_vecs, (size_t)n * dim));
    VG_TRY(out.alloc((size_t)n * m));
    VG_TRY(dev_pq_encode(v.as<float>(), n, dim, (int)m, (int)k, p.cb.as<int8_t>(), p.sc.as<float>(), p.of.as<float>(),
                         out.as<uint8_t>(), stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_codes, out.p, (size_t)n * m);
}
vg_status vg_pq_decode(const uint8_t *h_codes, int64_t n, int64_t dim, int64_t m, int64_t k, const int8_t *h_cb, const float *h_sc,
                       const float *h_of, float *h_vecs) {
    VG_ENTER();
    PqDev p;
    VG_TRY(pq_params(p, dim, m, k, h_cb, h_sc, h_of));
    if (n <= 0) return VG_OK;
    DevBuf c, out;
    VG_TRY(to_device(c, h_codes, (size_t)n * m));
    VG_TRY(out.alloc((size_t)n * dim * 4));
    VG_TRY(dev_pq_decode(c.as<uint8_t>(), n, dim, (int)m, (int)k, p.cb.as<int8_t>(), p.sc.as<float>(), p.of.as<float>(),
                         out.as<float>(), stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_vecs, out.p, (size_t)n * dim * 4);
}
vg_status vg_pq_build_distance_table(const float *h_q, int64_t nq, int64_t dim, int64_t m, int64_t k, const int8_t *h_cb,
                                     const float *h_sc, const float *h_of, float *h_tables) {
    VG_ENTER();
    PqDev p;
    VG_TRY(pq_params(p, dim, m, k, h_cb, h_sc, h_of));
    if (nq <= 0) return VG_OK;
    DevBuf q, out;
    VG_TRY(to_device(q, h_q, (size_t)nq * dim));
    VG_TRY(out.alloc((size_t)nq * m * k * 4));
    VG_TRY(dev_pq_tables(q.as<float>(), nq, dim, (int)m, (int)k, p.cb.as<int8_t>(), p.sc.as<float>(), p.of.as<float>(),
                         out.as<float>(), stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_tables, out.p, (size_t)nq * m * k * 4);
}

// device-resident variants
vg_status vg_minmax_dev(const float *d_vecs, int64_t n, int64_t dim, float *h_mins, float *h_maxs) {
    VG_ENTER();
    if (n <= 0 || dim <= 0) return fail(VG_ERR_INVALID, "no vectors provided for training");
    DevBuf mm;
    VG_TRY(mm.alloc((size_t)dim * 8));
    VG_TRY(dev_minmax(d_vecs, n, dim, mm.as<float>(), mm.as<float>() + dim, stream()));
    VG_TRY(staged_d2h(h_mins, mm.p, (size_t)dim * 4));
    return staged_d2h(h_maxs, mm.as<float>() + dim, (size_t)dim * 4);
}
vg_status vg_sq8_encode_dev(const float *d_vecs, int64_t n, int64_t dim, const float *h_mins, const float *h_maxs,
                            const float *h_scales, uint8_t *d_codes) {
    VG_ENTER();
    if (!h_mins || !h_maxs || !h_scales) return fail(VG_ERR_STATE, "ScalarQuantizer not trained");
    if (n <= 0) return VG_OK;
    DevBuf mn, mx, sc;
    VG_TRY(to_device(mn, h_mins, (size_t)dim));
    VG_TRY(to_device(mx, h_maxs, (size_t)dim));
    VG_TRY(to_device(sc, h_scales, (size_t)dim));
    VG_TRY(dev_sq8_encode(d_vecs, n, dim, mn.as<float>(), mx.as<float>(), sc.as<float>(), d_codes, stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return VG_OK;
}
vg_status vg_int4_encode_dev(const float *d_vecs, int64_t n, int64_t dim, const float *h_min, const float *h_diff, uint8_t *d_codes) {
    VG_ENTER();
    if (n <= 0) return VG_OK;
    DevBuf mn, df;
    VG_TRY(to_device(mn, h_min, (size_t)dim));
    VG_TRY(to_device(df, h_diff, (size_t)dim));
    VG_TRY(dev_int4_encode(d_vecs, n, dim, mn.as<float>(), df.as<float>(), d_codes, stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return VG_OK;
}
vg_status vg_rabitq_encode_dev(const float *d_vecs, int64_t n, int64_t dim, uint8_t *d_codes) {
    VG_ENTER();
    if (n <= 0) return VG_OK;
    VG_TRY(dev_sign_encode(d_vecs, n, dim, 0.0f, true, d_codes, stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return VG_OK;
}
vg_status vg_pq_encode_dev(const float *d_vecs, int64_t n, int64_t dim, int64_t m, int64_t k, const int8_t *h_cb, const float *h_sc,
                           const float *h_of, uint8_t *d_codes) {
    VG_ENTER();
    PqDev p;
    VG_TRY(pq_params(p, dim, m, k, h_cb, h_sc, h_of));
    if (n <= 0) return VG_OK;
    VG_TRY(dev_pq_encode(d_vecs, n, dim, (int)m, (int)k, p.cb.as<int8_t>(), p.sc.as<float>(), p.of.as<float>(), d_codes, stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    return VG_OK;
}

// ------------------------------------------------------------ flat segment file
// format.go:110-165 / segment.go:105-342
static uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }
static const size_t kFlatHeaderSize = 152;

vg_status vg_flat_decode_header(const uint8_t *f, size_t len, vg_flat_header *out) {
    if (!f || len < kFlatHeaderSize) return fail(VG_ERR_FORMAT, "buffer too small for header");
    if (rd32(f) != 0x56454331u) return fail(VG_ERR_FORMAT, "invalid magic number");
    if (rd32(f + 4) != 1) return fail(VG_ERR_FORMAT, "unsupported version");
    out->segment_id = rd64(f + 8);
    out->row_count = rd32(f + 16);
    out->dim = rd32(f + 20);
    out->metric = f[24];
    out->num_partitions = rd32(f + 28);
    out->quantization_type = f[32];
    out->checksum = rd32(f + 104);
    return VG_OK;
}

// CRC32C of the body, computed on the device copy of the file (one pass, 4 KiB
// strides combined on the host would need GF(2) matrices; the body is hashed
// by a single sequential slicing-by-1 kernel thread per 1 MiB chunk and the
// chunk CRCs are combined with the standard zero-extension operator).
__global__ void crc32c_chunks_kernel(const uint8_t *data, size_t n, size_t chunk, uint32_t *out) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t b = c * chunk;
    if (b >= n) return;
    size_t e = b + chunk;
    if (e > n) e = n;
    uint32_t crc = 0;  // raw register value, no pre/post inversion
    for (size_t i = b; i < e; i++) {
        crc ^= data[i];
        for (int k = 0; k < 8; k++) crc = (crc >> 1) ^ (0x82F63B78u & (0u - (crc & 1u)));
    }
    out[c] = crc;
}
static uint32_t gf2_times(const uint32_t *mat, uint32_t vec) {
    uint32_t s = 0;
    while (vec) {
        if (vec & 1) s ^= *mat;
        vec >>= 1;
        mat++;
    }
    return s;
}
static void gf2_square(uint32_t *sq, const uint32_t *mat) {
    for (int n = 0; n < 32; n++) sq[n] = gf2_times(mat, mat[n]);
}
// advance raw crc register over `len` zero bytes
static uint32_t crc_shift(uint32_t crc, size_t len) {
    uint32_t even[32], odd[32];
    odd[0] = 0x82F63B78u;
    uint32_t row = 1;
    for (int n = 1; n < 32; n++) {
        odd[n] = row;
        row <<= 1;
    }
    gf2_square(even, odd);
    gf2_square(odd, even);
    do {
        gf2_square(even, odd);
        if (len & 1) crc = gf2_times(even, crc);
        len >>= 1;
        if (!len) break;
        gf2_square(odd, even);
        if (len & 1) crc = gf2_times(odd, crc);
        len >>= 1;
    } while (len);
    return crc;
}
static vg_status device_crc32c(const uint8_t *d_data, size_t n, uint32_t *crc_out) {
    const size_t chunk = 4096;
    const size_t chunks = (n + chunk - 1) / chunk;
    if (n == 0) {
        *crc_out = 0;
        return VG_OK;
    }
    DevBuf out;
    VG_TRY(out.alloc(chunks * 4));
    crc32c_chunks_kernel<<<(unsigned)((chunks + 127) / 128), 128, 0, stream()>>>(d_data, n, chunk, out.as<uint32_t>());
    VG_LAUNCHED();
    std::vector<uint32_t> h(chunks);
    VG_CUDA(cudaMemcpyAsync(h.data(), out.p, chunks * 4, cudaMemcpyDeviceToHost, stream()));
    VG_CUDA(cudaStreamSynchronize(stream()));
    // crc(A‖B) with init I: raw(A‖B) = shift(raw_I(A), |B|) ^ raw_0(B); fold left to right.
    uint32_t reg = 0xFFFFFFFFu;
    for (size_t c = 0; c < chunks; c++) {
        const size_t len = (c + 1 == chunks) ? n - c * chunk : chunk;
        reg = crc_shift(reg, len) ^ h[c];
    }
    *crc_out = reg ^ 0xFFFFFFFFu;
    return VG_OK;
}

// A section [off, off + size) lies inside a file of `len` bytes — written so that neither side can wrap: a crafted
// header with an offset near 2^64 must fail here, not read out of bounds (the Go reference slices with bounds checks).
static bool section_ok(uint64_t len, uint64_t off, uint64_t size) { return off <= len && size <= len - off; }
static bool mul_ok(uint64_t a, uint64_t b, uint64_t *out) {
    if (a != 0 && b > UINT64_MAX / a) return false;
    *out = a * b;
    return true;
}

vg_status vg_flat_open_on(int32_t device, const uint8_t *f, size_t len, int32_t verify_checksum, vg_index_t *out) {
    VG_ENTER(device);
    vg_flat_header h;
    VG_TRY(vg_flat_decode_header(f, len, &h));
    // the checksum covers everything behind the header: verify it before any section is trusted
    if (verify_checksum && h.checksum != 0 && len > kFlatHeaderSize) {
        DevBuf body;
        VG_TRY(to_device(body, f + kFlatHeaderSize, len - kFlatHeaderSize));
        uint32_t crc = 0;
        VG_TRY(device_crc32c(body.as<uint8_t>(), len - kFlatHeaderSize, &crc));
        if (crc != h.checksum) {
            char msg[96];
            snprintf(msg, sizeof msg, "checksum mismatch: expected %x, got %x", h.checksum, crc);
            return fail(VG_ERR_FORMAT, msg);
        }
    }
    const uint64_t off_centroid = rd64(f + 40), off_part = rd64(f + 48), off_quant = rd64(f + 56), off_codes = rd64(f + 64),
                   off_vec = rd64(f + 72), off_pk = rd64(f + 80), off_meta = rd64(f + 88);
    const uint64_t rows = h.row_count, dim = h.dim, L = (uint64_t)len;
    if (dim == 0) return fail(VG_ERR_FORMAT, "zero dimension");
    uint64_t row_floats = 0, vec_bytes = 0;
    if (!mul_ok(rows, dim, &row_floats) || !mul_ok(row_floats, 4, &vec_bytes)) return fail(VG_ERR_FORMAT, "rows x dim overflows");
    vg_index_desc d;
    memset(&d, 0, sizeof d);
    d.dim = (int64_t)dim;
    d.rows = (int64_t)rows;
    d.metric = (int32_t)h.metric;
    d.segment_id = (uint32_t)h.segment_id;
    std::vector<float> mins, maxs, scales, inv, cent;
    std::vector<uint32_t> poff;
    std::vector<float> pq_sc, pq_of;
    const uint8_t *codes = nullptr;
    if (h.num_partitions > 0) {
        uint64_t cb = 0;
        if (!mul_ok((uint64_t)h.num_partitions, dim * 4, &cb) || !section_ok(L, off_centroid, cb))
            return fail(VG_ERR_FORMAT, "file too short for centroids");
        if (!section_ok(L, off_part, ((uint64_t)h.num_partitions + 1) * 4)) return fail(VG_ERR_FORMAT, "file too short for partition offsets");
        cent.resize((size_t)h.num_partitions * dim);
        memcpy(cent.data(), f + off_centroid, cb);
        poff.resize((size_t)h.num_partitions + 1);
        memcpy(poff.data(), f + off_part, poff.size() * 4);
        for (size_t i = 0; i + 1 < poff.size(); i++)
            if (poff[i] > poff[i + 1] || poff[i + 1] > rows) return fail(VG_ERR_FORMAT, "partition offsets are not a monotone cover of the rows");
        d.num_partitions = h.num_partitions;
        d.centroids = cent.data();
        d.partition_offsets = poff.data();
    }
    if (h.quantization_type == 1) {  // QuantizationSQ8 → SetBounds(mins, maxs)
        if (!section_ok(L, off_quant, dim * 8)) return fail(VG_ERR_FORMAT, "file too short for quantization metadata");
        mins.resize(dim);
        maxs.resize(dim);
        scales.resize(dim);
        inv.resize(dim);
        memcpy(mins.data(), f + off_quant, dim * 4);
        memcpy(maxs.data(), f + off_quant + dim * 4, dim * 4);
        VG_TRY(vg_sq8_set_bounds(mins.data(), maxs.data(), (int64_t)dim, scales.data(), inv.data()));
        if (!section_ok(L, off_codes, row_floats)) return fail(VG_ERR_FORMAT, "file too short for codes");
        d.codec = VG_CODEC_SQ8;
        d.sq8_mins = mins.data();
        d.sq8_inv_scales = inv.data();
        codes = f + off_codes;
    } else if (h.quantization_type == 2) {  // QuantizationPQ
        if (!section_ok(L, off_quant, 8)) return fail(VG_ERR_FORMAT, "file too short for PQ metadata");
        const uint64_t m = rd32(f + off_quant), k = rd32(f + off_quant + 4);
        if (m == 0 || dim % m != 0) return fail(VG_ERR_FORMAT, "dimension must be divisible by numSubvectors");
        if (k == 0 || k > 256) return fail(VG_ERR_FORMAT, "numCentroids must be in 1..256");
        const uint64_t dsub = dim / m, cbsize = m * k * dsub, meta = 8 + m * 8 + cbsize;  // m <= dim < 2^32, k <= 256: no overflow
        if (!section_ok(L, off_quant, meta)) return fail(VG_ERR_FORMAT, "file too short for PQ metadata");
        pq_sc.resize(m);
        pq_of.resize(m);
        memcpy(pq_sc.data(), f + off_quant + 8, m * 4);
        memcpy(pq_of.data(), f + off_quant + 8 + m * 4, m * 4);
        uint64_t code_bytes = 0;
        if (!mul_ok(rows, m, &code_bytes) || !section_ok(L, off_codes, code_bytes)) return fail(VG_ERR_FORMAT, "file too short for codes");
        d.codec = VG_CODEC_PQ;
        d.pq_m = (int64_t)m;
        d.pq_k = (int64_t)k;
        d.pq_codebooks = reinterpret_cast<const int8_t *>(f + off_quant + 8 + m * 8);
        d.pq_scales = pq_sc.data();
        d.pq_offsets = pq_of.data();
        codes = f + off_codes;
    } else if (h.quantization_type == 0) {
        d.codec = VG_CODEC_F32;
    } else {
        return fail(VG_ERR_FORMAT, "unknown quantization type");
    }
    if (!section_ok(L, off_vec, vec_bytes)) return fail(VG_ERR_FORMAT, "file too short for vectors");
    if (off_meta < off_pk || !section_ok(L, off_pk, off_meta - off_pk)) return fail(VG_ERR_FORMAT, "file too short for IDs");
    if (rows > 0 && (off_meta - off_pk) / 8 < rows) return fail(VG_ERR_FORMAT, "id section too small");
    vg_index_t idx = 0;
    VG_TRY(vg_index_create_on(t_ts.ctx->device, &d, &idx));
    vg_status s = VG_OK;
    if (rows > 0) {
        // sections are only 4-byte aligned in the file (HeaderSize = 152); staging re-aligns them
        std::vector<float> tmp;
        const float *vec = reinterpret_cast<const float *>(f + off_vec);
        if ((reinterpret_cast<uintptr_t>(vec) & 3) != 0) {
            tmp.resize(rows * dim);
            memcpy(tmp.data(), f + off_vec, rows * dim * 4);
            vec = tmp.data();
        }
        s = vg_index_upload(idx, 0, (int64_t)rows, codes, vec);
        if (s == VG_OK) {
            Index *ix = lookup(idx);
            s = ix->ids.alloc_persistent(rows * 8);
            if (s == VG_OK) s = staged_h2d(ix->ids.p, f + off_pk, rows * 8);
            if (s == VG_OK) ix->has_ids = true;
        }
    }
    if (s != VG_OK) {
        std::string keep = t_error;
        vg_index_close(idx);
        t_error = keep;
        return s;
    }
    *out = idx;
    return VG_OK;
}
vg_status vg_flat_open(const uint8_t *f, size_t len, int32_t verify_checksum, vg_index_t *out) {
    return vg_flat_open_on(-1, f, len, verify_checksum, out);
}

__global__ void gather_u64_kernel(const uint64_t *src, const uint32_t *rows, int64_t n, int64_t nrows, uint64_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = ((int64_t)rows[i] < nrows) ? src[rows[i]] : 0ull;
}
vg_status vg_index_fetch_ids(vg_index_t idx, const uint32_t *h_rows, int64_t n, uint64_t *h_ids) {
    Index *ix = lookup(idx);
    if (!ix) return fail(VG_ERR_STATE, "unknown or closed index handle");
    VG_ENTER_IX(ix);
    if (!ix->has_ids) return fail(VG_ERR_STATE, "index has no id column");
    if (n <= 0) return VG_OK;
    DevBuf r, o;
    VG_TRY(to_device(r, h_rows, (size_t)n));
    VG_TRY(o.alloc((size_t)n * 8));
    gather_u64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>(ix->ids.as<uint64_t>(), r.as<uint32_t>(), n, ix->d.rows,
                                                                         o.as<uint64_t>());
    VG_LAUNCHED();
    VG_CUDA(cudaStreamSynchronize(stream()));
    return staged_d2h(h_ids, o.p, (size_t)n * 8);
}

// ------------------------------------------------------------ shard groups (row shards + NCCL merge inside the library)
// SURVEY 8(b) "Ownership": the library owns the NCCL communicators; 8(e): every GPU scans its row shard for the whole
// query batch and ONE all-gather of the per-shard top-k (8-byte sortable keys) feeds a device merge on every GPU —
// engine/search.go:903-908 with the segments living on different GPUs.  NCCL is resolved at run time (dlopen of
// libnccl.so.2: the copy a host process already loaded, else the system library), so the shared library carries no
// link-time dependency a Go host could not satisfy.
//   single process, W GPUs (what a Go host uses):  vg_shard_group_create(devices, W) — ncclCommInitAll; calls take the W
//     shard handles and fan out over W host threads (one per GPU), results come from member 0
//   one process per GPU (torchrun):  vg_nccl_unique_id on rank 0, shipped to the others by the launcher's own means,
//     then vg_shard_group_create_rank(id, rank, world, device); calls take this process's ONE shard handle
namespace {
typedef void *nccl_comm_t;
struct NcclId {
    char internal[128];
};
struct NcclApi {
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, NcclId, int) = nullptr;
    int (*CommInitAll)(nccl_comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;
void load_nccl() {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy the host process already uses (e.g. PyTorch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))dlsym(h, "ncclCommInitAll");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommInitAll && g_nccl.CommDestroy && g_nccl.AllGather;
}
vg_status need_nccl() {
    std::call_once(g_nccl_once, load_nccl);
    return g_nccl.ok ? VG_OK : fail(VG_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
}
vg_status nccl_fail(int r, const char *what) {
    return fail(VG_ERR_CUDA, std::string("NCCL error: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?") + " in " + what);
}
const int kNcclUint64 = 5;   // ncclUint64 (nccl.h ncclDataType_t)

struct ShardGroup {
    int world = 1;                       // shards in the group (= GPUs)
    int members = 1;                     // shards driven by THIS process: world (single process) or 1
    int rank0 = 0;                       // global rank of member 0
    std::vector<int> devices;            // [members]
    std::vector<nccl_comm_t> comms;      // [members]
    std::vector<cudaStream_t> streams;   // [members] the group's own stream on each device (searches of one group are serialised)
    std::mutex mu;
};
std::unordered_map<uint64_t, ShardGroup *> g_groups;
ShardGroup *lookup_group(uint64_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_groups.find(h);
    return it == g_groups.end() ? nullptr : it->second;
}
uint64_t register_group(ShardGroup *g) {
    std::lock_guard<std::mutex> lk(g_mu);
    const uint64_t h = g_next_handle++;
    g_groups[h] = g;
    return h;
}

// global rows a shard owns -> local row ids (others 0xFFFFFFFF), and the global id kept only where owned
__global__ void __launch_bounds__(256) owned_rows_kernel(const uint32_t *rows, int64_t n, uint32_t row_base, uint32_t nrows, uint32_t *local,
                                                         uint32_t *mine) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = rows[i];
    const bool own = g != 0xFFFFFFFFu && g >= row_base && g - row_base < nrows;
    local[i] = own ? g - row_base : 0xFFFFFFFFu;
    mine[i] = own ? g : 0xFFFFFFFFu;
}

// One member's part of a sharded search on ITS device and stream: local scan, pack, all-gather, merge (and for the
// rerank form: global approximate top-r, exact scores of the owned rows, second exchange, final top-k).
vg_status member_search(ShardGroup *g, int m, Index *ix, const float *d_queries, int64_t nq, int64_t r, int64_t k, bool rerank, uint32_t *d_rows,
                        float *d_scores, int32_t *d_counts) {
    VG_ENTER(g->devices[(size_t)m], true, g->streams[(size_t)m]);
    cudaStream_t st = stream();
    const int W = g->world;
    const int64_t kin = rerank ? r : k;
    const bool seg_desc = ix->d.metric != VG_METRIC_L2;
    const bool approx_desc = (ix->d.codec == VG_CODEC_F32 || ix->d.codec == VG_CODEC_SQ8 || ix->d.codec == VG_CODEC_PQ || ix->d.codec == VG_CODEC_OPQ)
                                 ? seg_desc : false;   // INT4 / BQ / RaBitQ scores are distances whatever the segment metric
    DevBuf lr, ls, lc, keys, allk;
    VG_TRY(lr.alloc((size_t)nq * kin * 4));
    VG_TRY(ls.alloc((size_t)nq * kin * 4));
    VG_TRY(lc.alloc((size_t)nq * 4));
    VG_TRY(keys.alloc((size_t)nq * kin * 8));
    VG_TRY(allk.alloc((size_t)W * nq * kin * 8));
    VG_TRY(search_dev_impl(ix, d_queries, nq, kin, 0, nullptr, lr.as<uint32_t>(), ls.as<float>(), lc.as<int32_t>()));
    VG_TRY(launch_pack_keys(lr.as<uint32_t>(), ls.as<float>(), nq * kin, approx_desc, keys.as<unsigned long long>(), st));
    int rc = g_nccl.AllGather(keys.p, allk.p, (size_t)(nq * kin), kNcclUint64, g->comms[(size_t)m], st);
    if (rc != 0) return nccl_fail(rc, "ncclAllGather");
    if (!rerank) {
        VG_TRY(launch_merge_keys(allk.as<unsigned long long>(), W, nq, kin, nq * kin, kin, approx_desc, k, d_rows, d_scores, d_counts, st));
        VG_CUDA(cudaStreamSynchronize(st));
        return VG_OK;
    }
    // rerank form (engine/search.go:188-192,913-973): the GLOBAL approximate top-r is reranked, so ids do not depend on W
    if (!ix->has_vectors) return fail(VG_ERR_STATE, "index holds no float32 vectors to rerank against");
    DevBuf gr, gs, gc, local, mine, exact;
    VG_TRY(gr.alloc((size_t)nq * r * 4));
    VG_TRY(gs.alloc((size_t)nq * r * 4));
    VG_TRY(gc.alloc((size_t)nq * 4));
    VG_TRY(local.alloc((size_t)nq * r * 4));
    VG_TRY(mine.alloc((size_t)nq * r * 4));
    VG_TRY(exact.alloc((size_t)nq * r * 4));
    VG_TRY(launch_merge_keys(allk.as<unsigned long long>(), W, nq, r, nq * r, r, approx_desc, r, gr.as<uint32_t>(), gs.as<float>(), gc.as<int32_t>(),
                             st));
    const int64_t n = nq * r;
    owned_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gr.as<uint32_t>(), n, (uint32_t)ix->d.row_base, (uint32_t)ix->d.rows,
                                                                   local.as<uint32_t>(), mine.as<uint32_t>());
    VG_LAUNCHED();
    VG_TRY(rerank_rows(ix, d_queries, nq, local.as<uint32_t>(), r, seg_desc, exact.as<float>(), st));
    VG_TRY(launch_pack_keys(mine.as<uint32_t>(), exact.as<float>(), n, seg_desc, keys.as<unsigned long long>(), st));
    rc = g_nccl.AllGather(keys.p, allk.p, (size_t)n, kNcclUint64, g->comms[(size_t)m], st);
    if (rc != 0) return nccl_fail(rc, "ncclAllGather");
    VG_TRY(launch_merge_keys(allk.as<unsigned long long>(), W, nq, r, nq * r, r, seg_desc, k, d_rows, d_scores, d_counts, st));
    VG_CUDA(cudaStreamSynchronize(st));
    return VG_OK;
}

// Fan one call out over the group's members (one host thread per GPU beyond the first), host queries in, member 0's
// result out.  d_queries_per_member non-null: device-resident form (queries and outputs per member).
vg_status group_search(uint64_t gh, const vg_index_t *shards, const float *h_queries, const float *const *d_q, int64_t nq, int64_t r, int64_t k,
                       bool rerank, uint32_t *h_rows, float *h_scores, int32_t *h_counts, uint32_t *const *d_rows, float *const *d_scores,
                       int32_t *const *d_counts) {
    VG_TRY(need_nccl());
    ShardGroup *g = lookup_group(gh);
    if (!g) return fail(VG_ERR_STATE, "unknown or closed shard group");
    if (!shards || nq < 0 || k <= 0 || (rerank && r < k)) return fail(VG_ERR_INVALID, "bad sharded search arguments");
    if (nq == 0) return VG_OK;
    std::lock_guard<std::mutex> lk(g->mu);   // collectives of one communicator must be issued in the same order everywhere
    const int M = g->members;
    std::vector<Index *> ix((size_t)M);
    for (int m = 0; m < M; m++) {
        ix[(size_t)m] = lookup(shards[m]);
        if (!ix[(size_t)m]) return fail(VG_ERR_STATE, "unknown or closed index handle");
        if (ix[(size_t)m]->device != g->devices[(size_t)m]) return fail(VG_ERR_INVALID, "shard handle lives on another GPU than its group member");
    }
    std::vector<vg_status> status((size_t)M, VG_OK);
    std::vector<std::string> errs((size_t)M);
    auto work = [&](int m) {
        auto run = [&]() -> vg_status {
            VG_ENTER(g->devices[(size_t)m], true, g->streams[(size_t)m]);
            DevBuf q, rows, scores, counts;
            const float *dq = d_q ? d_q[m] : nullptr;
            uint32_t *orow = d_rows ? d_rows[m] : nullptr;
            float *osc = d_scores ? d_scores[m] : nullptr;
            int32_t *ocnt = d_counts ? d_counts[m] : nullptr;
            if (!dq) {
                VG_TRY(to_device(q, h_queries, (size_t)nq * ix[(size_t)m]->d.dim));
                dq = q.as<float>();
            }
            if (!orow) {
                VG_TRY(rows.alloc((size_t)nq * k * 4));
                VG_TRY(scores.alloc((size_t)nq * k * 4));
                VG_TRY(counts.alloc((size_t)nq * 4));
                orow = rows.as<uint32_t>();
                osc = scores.as<float>();
                ocnt = counts.as<int32_t>();
            }
            VG_TRY(member_search(g, m, ix[(size_t)m], dq, nq, r, k, rerank, orow, osc, ocnt));
            if (m == 0 && h_rows) {
                VG_TRY(staged_d2h(h_rows, orow, (size_t)nq * k * 4));
                VG_TRY(staged_d2h(h_scores, osc, (size_t)nq * k * 4));
                VG_TRY(staged_d2h(h_counts, ocnt, (size_t)nq * 4));
            }
            return VG_OK;
        };
        status[(size_t)m] = run();
        if (status[(size_t)m] != VG_OK) errs[(size_t)m] = t_error;
    };
    std::vector<std::thread> th;
    for (int m = 1; m < M; m++) th.emplace_back(work, m);
    work(0);
    for (auto &t : th) t.join();
    for (int m = 0; m < M; m++)
        if (status[(size_t)m] != VG_OK) return fail(status[(size_t)m], errs[(size_t)m]);
    return VG_OK;
}
}  // namespace

vg_status vg_nccl_unique_id(uint8_t *id128) {
    VG_TRY(need_nccl());
    if (!id128) return fail(VG_ERR_INVALID, "null argument");
    NcclId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return nccl_fail(rc, "ncclGetUniqueId");
    memcpy(id128, id.internal, 128);
    return VG_OK;
}

vg_status vg_shard_group_create(const int32_t *devices, int32_t n, vg_shard_group_t *out) {
    VG_TRY(need_nccl());
    if (!devices || n < 1 || n > kMaxDevices || !out) return fail(VG_ERR_INVALID, "bad device list");
    std::unique_ptr<ShardGroup> g(new ShardGroup());
    g->world = g->members = n;
    g->rank0 = 0;
    g->devices.assign(devices, devices + n);
    g->comms.assign((size_t)n, nullptr);
    g->streams.assign((size_t)n, nullptr);
    for (int m = 0; m < n; m++) {
        VG_ENTER(devices[m]);   // creates the device context (checks the ordinal)
        VG_CUDA(cudaStreamCreateWithFlags(&g->streams[(size_t)m], cudaStreamNonBlocking));
    }
    std::vector<int> devs(devices, devices + n);
    const int rc = g_nccl.CommInitAll(g->comms.data(), n, devs.data());
    if (rc != 0) return nccl_fail(rc, "ncclCommInitAll");
    *out = register_group(g.release());
    return VG_OK;
}

vg_status vg_shard_group_create_rank(const uint8_t *id128, int32_t rank, int32_t world, int32_t device, vg_shard_group_t *out) {
    VG_TRY(need_nccl());
    if (!id128 || world < 1 || rank < 0 || rank >= world || !out) return fail(VG_ERR_INVALID, "bad rank / world");
    VG_ENTER(device);
    std::unique_ptr<ShardGroup> g(new ShardGroup());
    g->world = world;
    g->members = 1;
    g->rank0 = rank;
    g->devices.assign(1, t_ts.ctx->device);
    g->comms.assign(1, nullptr);
    g->streams.assign(1, nullptr);
    VG_CUDA(cudaStreamCreateWithFlags(&g->streams[0], cudaStreamNonBlocking));
    NcclId id;
    memcpy(id.internal, id128, 128);
    const int rc = g_nccl.CommInitRank(&g->comms[0], world, id, rank);
    if (rc != 0) return nccl_fail(rc, "ncclCommInitRank");
    *out = register_group(g.release());
    return VG_OK;
}

vg_status vg_shard_group_destroy(vg_shard_group_t gh) {
    ShardGroup *g = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_groups.find(gh);
        if (it == g_groups.end()) return fail(VG_ERR_STATE, "unknown or closed shard group");
        g = it->second;
        g_groups.erase(it);
    }
    for (int m = 0; m < g->members; m++) {
        cudaSetDevice(g->devices[(size_t)m]);
        if (g->streams[(size_t)m]) {
            cudaStreamSynchronize(g->streams[(size_t)m]);
            cudaStreamDestroy(g->streams[(size_t)m]);
        }
        if (g->comms[(size_t)m] && g_nccl.CommDestroy) g_nccl.CommDestroy(g->comms[(size_t)m]);
    }
    delete g;
    return VG_OK;
}

vg_status vg_shard_group_info(vg_shard_group_t gh, int32_t *world, int32_t *members, int32_t *first_rank) {
    ShardGroup *g = lookup_group(gh);
    if (!g) return fail(VG_ERR_STATE, "unknown or closed shard group");
    if (world) *world = g->world;
    if (members) *members = g->members;
    if (first_rank) *first_rank = g->rank0;
    return VG_OK;
}

vg_status vg_shard_group_search(vg_shard_group_t gh, const vg_index_t *shards, const float *h_queries, int64_t nq, int64_t k, uint32_t *h_out_rows,
                                float *h_out_scores, int32_t *h_out_counts) {
    return group_search(gh, shards, h_queries, nullptr, nq, k, k, false, h_out_rows, h_out_scores, h_out_counts, nullptr, nullptr, nullptr);
}
vg_status vg_shard_group_search_rerank(vg_shard_group_t gh, const vg_index_t *shards, const float *h_queries, int64_t nq, int64_t r, int64_t k,
                                       uint32_t *h_out_rows, float *h_out_scores, int32_t *h_out_counts) {
    return group_search(gh, shards, h_queries, nullptr, nq, r, k, true, h_out_rows, h_out_scores, h_out_counts, nullptr, nullptr, nullptr);
}
vg_status vg_shard_group_search_dev(vg_shard_group_t gh, const vg_index_t *shards, const float *const *d_queries, int64_t nq, int64_t k,
                                    uint32_t *const *d_out_rows, float *const *d_out_scores, int32_t *const *d_out_counts) {
    if (!d_queries || !d_out_rows || !d_out_scores || !d_out_counts) return fail(VG_ERR_INVALID, "null argument");
    return group_search(gh, shards, nullptr, d_queries, nq, k, k, false, nullptr, nullptr, nullptr, d_out_rows, d_out_scores, d_out_counts);
}
vg_status vg_shard_group_search_rerank_dev(vg_shard_group_t gh, const vg_index_t *shards, const float *const *d_queries, int64_t nq, int64_t r,
                                           int64_t k, uint32_t *const *d_out_rows, float *const *d_out_scores, int32_t *const *d_out_counts) {
    if (!d_queries || !d_out_rows || !d_out_scores || !d_out_counts) return fail(VG_ERR_INVALID, "null argument");
    return group_search(gh, shards, nullptr, d_queries, nq, r, k, true, nullptr, nullptr, nullptr, d_out_rows, d_out_scores, d_out_counts);
}

}  // extern "C"
