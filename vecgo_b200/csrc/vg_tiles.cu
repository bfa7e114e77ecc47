// Tile lists for the CTA-pair filters: see vg_tiles.cuh.
#include <atomic>
#include <cstdlib>

#include "vg_tiles.cuh"

namespace vg {
namespace tiles {

static std::atomic<int> g_on{-1};
bool enabled() {
    int v = g_on.load();
    if (v < 0) {
        const char *e = getenv("VECGO_TILE_SKIP");
        v = (e && e[0] == '0') ? 0 : 1;
        g_on.store(v);
    }
    return v != 0;
}
void set_enabled(bool on) { g_on.store(on ? 1 : 0); }

//   list[0 .. *count)  active tiles, ascending (one block, ordered compaction: the work split is reproducible)
//   skip[0 .. *nskip)  the others (their row groups get the "nothing here" entry the select kernel expects)
__global__ void __launch_bounds__(1024) build_tile_list_kernel(const uint32_t *mask, int64_t rows, int ntiles, int32_t *list, int32_t *count,
                                                               int32_t *skip, int32_t *nskip) {
    __shared__ int wsum[32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < ntiles; t0 += 1024) {
        const int t = t0 + tid;
        bool act = false;
        if (t < ntiles) {
            const int64_t n0 = (int64_t)t * TILE_ROWS;
            if (n0 + TILE_ROWS > rows) act = true;   // the ragged last tile is always scanned (as i + BlockSize <= end)
            else {
                uint32_t any = 0;
#pragma unroll
                for (int w = 0; w < TILE_ROWS / 32; w++) any |= __ldg(mask + (n0 >> 5) + w);
                act = any != 0;
            }
        }
        const uint32_t b = __ballot_sync(0xffffffffu, act);
        if (lane == 0) wsum[warp] = __popc(b);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; w++) {
            const int c = wsum[w];
            if (w < warp) before += c;
            total += c;
        }
        const int base = base_s;
        if (t < ntiles) {
            const int rank_act = before + __popc(b & ((1u << lane) - 1u));
            if (act) list[base + rank_act] = t;
            else skip[(t0 - base) + (tid - rank_act)] = t;   // skipped tiles before t0 = t0 - base
        }
        __syncthreads();
        if (tid == 0) base_s = base + total;
        __syncthreads();
    }
    if (tid == 0) {
        *count = base_s;
        *nskip = ntiles - base_s;
    }
}
// the minima-plane entries of the skipped tiles' row groups: (BIG, BIG) — what the epilogue writes for a group whose
// rows are all masked (the exact stage drops any row it decodes from such an entry: its bitmap bit is clear)
__global__ void __launch_bounds__(256) fill_skipped_groups_kernel(const int32_t *skip, const int32_t *nskip, int gpt, int64_t groups, int64_t nq,
                                                                  float2 *mins) {
    const int n = __ldg(nskip);
    const float BIG = 3.0e38f;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int64_t g0 = (int64_t)__ldg(skip + i) * gpt;
        for (int64_t q = threadIdx.x; q < nq; q += blockDim.x)
            for (int j = 0; j < gpt; j++)
                if (g0 + j < groups) mins[q * groups + g0 + j] = make_float2(BIG, BIG);
    }
}


vg_status build(const uint32_t *mask, int64_t rows, Lists &out, cudaStream_t st) {
    const int64_t nt = (rows + TILE_ROWS - 1) / TILE_ROWS;
    const size_t list_ints = (size_t)nt + 1024;   // slack: a split's empty slice may start a few entries past the end
    VG_TRY(out.buf.alloc((2 * list_ints + 2) * 4));
    VG_CUDA(cudaMemsetAsync(out.buf.p, 0, (2 * list_ints + 2) * 4, st));
    out.list = out.buf.as<int32_t>();
    out.skip = out.list + list_ints;
    out.count = out.skip + list_ints;
    out.nskip = out.count + 1;
    build_tile_list_kernel<<<1, 1024, 0, st>>>(mask, rows, (int)nt, out.list, out.count, out.skip, out.nskip);
    VG_LAUNCHED();
    return VG_OK;
}
vg_status fill_skipped_groups(const Lists &l, int gpt, int64_t groups, int64_t nq, float2 *mins, cudaStream_t st) {
    fill_skipped_groups_kernel<<<(unsigned)(sm_count() * 4), 256, 0, st>>>(l.skip, l.nskip, gpt, groups, nq, mins);
    VG_LAUNCHED();
    return VG_OK;
}

}  // namespace tiles
}  // namespace vg
