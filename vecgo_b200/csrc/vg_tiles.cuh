// Tile skipping of the CTA-pair filters (vg_quant_tc.cu qtc2_kernel, vg_flat_tc.cu flat2_kernel).
//
// A 256-row tile that lies wholly inside the segment and whose 256 bitmap bits are all clear cannot contribute a row:
// the pair kernels then never fetch, decode or multiply it.  This is where the reference's block-stat skipping lands on
// the device (flat/segment.go:524-541,613-630: a BlockSize = 1024-row block whose statistics cannot match the filter is
// jumped over — the host clears those blocks' bits, vg_index_search_blocks*), and the same test also drops tiles emptied
// by tombstones or a selective metadata filter.
#pragma once
#include "vg_common.cuh"

namespace vg {
namespace tiles {
constexpr int TILE_ROWS = 256;
struct Lists {
    DevBuf buf;
    int32_t *list = nullptr;    // active tiles, ascending; the splits of a launch share it evenly
    int32_t *count = nullptr;   // their number (device)
    int32_t *skip = nullptr;    // the other tiles
    int32_t *nskip = nullptr;
};
bool enabled();                 // on unless VECGO_TILE_SKIP=0 / vg_tile_skip_enable(0)
void set_enabled(bool on);
// mask: the row bitmap as 32-bit words.  Launch only (stream-ordered scratch).
vg_status build(const uint32_t *mask, int64_t rows, Lists &out, cudaStream_t st);
// (BIG, BIG) into the minima-plane entries [q][g] of the skipped tiles' row groups (gpt = groups per tile): what the
// epilogue writes for a group whose rows are all masked.
vg_status fill_skipped_groups(const Lists &l, int gpt, int64_t groups, int64_t nq, float2 *mins, cudaStream_t st);
}  // namespace tiles
}  // namespace vg
