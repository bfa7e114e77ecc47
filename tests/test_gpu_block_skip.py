"""Block-stat skipping (flat/segment.go:524-541,613-630) through the C ABI: vg_index_search_blocks[_dev].

The caller's verdict bitmap over the 1024-row blocks is folded into the row bitmap; the tensor-core filters then never
fetch a 256-row tile without an allowed row.  Whatever is skipped, the result must be what flat.(*Segment).Search
returns with the same blocks jumped over — the oracle's scan under the equivalent row mask — bit for bit, and
distance_computations must count only the rows of the blocks that were scanned.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as o

pytestmark = pytest.mark.gpu
F = np.float32
BLOCK = 1024


def bits(x):
    return np.ascontiguousarray(x, F).view(np.uint32)


@pytest.fixture(scope="module")
def vg():
    import vecgo_b200

    return vecgo_b200


def equivalent_row_mask(n, keep, row_bits=None):
    """Row mask the reference's loop is equivalent to: rows of skipped FULL blocks cleared, the ragged tail always scanned."""
    m = np.ones(n, bool) if row_bits is None else row_bits.copy()
    for b in range(n // BLOCK):
        if not keep[b]:
            m[b * BLOCK:(b + 1) * BLOCK] = False
    return m


def verdicts(rng, n, kind):
    full = n // BLOCK
    if kind == "random":
        return rng.random(full) < 0.4
    if kind == "clustered":      # a time-ordered field: one contiguous range of blocks can match
        k = np.zeros(full, bool)
        k[full // 3: full // 3 + max(1, full // 5)] = True
        return k
    if kind == "none":           # nothing but the ragged tail
        return np.zeros(full, bool)
    return np.ones(full, bool)


def make(vg, codec, rng, n, dim):
    L = vg._lib
    x = rng.standard_normal((n, dim)).astype(F)
    if codec == "f32":
        codes = None
        mk = lambda: vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n)
        okw = dict(dim=dim, metric=0, vectors=x)
    elif codec == "sq8":
        sq = vg.quantization.ScalarQuantizer(dim)
        sq.Train(x)
        codes = sq.EncodeBatch(x)
        mk = lambda: vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales))
        okw = dict(dim=dim, metric=0, quant=1, codes=codes, mins=sq.mins, inv=sq.invScales)
    elif codec == "int4":
        iq = vg.quantization.Int4Quantizer(dim)
        iq.Train(x)
        codes = iq.EncodeBatch(x)
        mk = lambda: vg.index.DeviceIndex(codec=L.CODEC_INT4, metric=0, dim=dim, rows=n, int4=(iq.min, iq.diff))
        okw = None
    else:
        codes = vg.quantization.RaBitQuantizer(dim).EncodeBatch(x)
        mk = lambda: vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=n)
        okw = None
    return x, codes, mk, okw


@pytest.mark.parametrize("codec,n,dim", [("sq8", 70_000 + 333, 128), ("int4", 40_000 + 77, 128), ("rabitq", 50_000 + 1, 256),
                                         ("f32", 60_000 + 99, 128)])
@pytest.mark.parametrize("kind", ["random", "clustered", "none", "all"])
def test_block_skipping_matches_masked_scan(vg, codec, n, dim, kind):
    L = vg._lib
    rng = np.random.default_rng(len(codec) * 100 + len(kind))
    x, codes, mk, okw = make(vg, codec, rng, n, dim)
    nq, k = 40, (20 if codec == "f32" else 10)    # k > 16: the Flat CTA-pair filter (the single-launch kernel serves k <= 16)
    q = rng.standard_normal((nq, dim)).astype(F)
    keep = verdicts(rng, n, kind)
    tomb = rng.random(n) < 0.9                     # tombstones on top of the block verdicts
    for row_bits in (None, tomb):
        eq = equivalent_row_mask(n, keep, row_bits)
        eq_mask = np.packbits(eq, bitorder="little")
        rmask = None if row_bits is None else np.packbits(row_bits, bitorder="little")
        with mk() as ix:
            if codec == "f32":
                ix.upload(vectors=x)
            else:
                ix.upload(codes=codes)
            got = ix.search(q, k, row_mask=rmask, block_keep=keep)
            st = L.last_search_stats()
            L.call("vg_tile_skip_enable", 0)
            try:
                noskip = ix.search(q, k, row_mask=rmask, block_keep=keep)
            finally:
                L.call("vg_tile_skip_enable", 1)
            L.call("vg_flat_tc_enable", 0)
            try:
                exact = ix.search(q, k, row_mask=eq_mask)   # the exact CUDA-core scan under the equivalent row mask
            finally:
                L.call("vg_flat_tc_enable", 1)
        skipped_rows = int((~keep).sum()) * BLOCK
        assert st["distance_computations"] == nq * (n - skipped_rows), st
        assert st["filter_queries"] == nq and st["exact_rerun_queries"] == 0, st
        for a, b in ((got, exact), (got, noskip)):
            assert np.array_equal(a[2], b[2])
            assert np.array_equal(a[0], b[0])
            assert np.array_equal(bits(a[1]), bits(b[1]))
        live = got[0][got[0] != 0xFFFFFFFF]
        assert eq[live].all()
        if kind == "none":
            assert (live >= (n // BLOCK) * BLOCK).all()     # only the ragged last block was scanned
        if okw is not None:
            seg = o.FlatOracle(**okw)
            out, cnt = seg.search_batch(q, k, threads=8, mask=eq_mask)
            for i in range(nq):
                c = int(cnt[i])
                assert int(got[2][i]) == c
                assert np.array_equal(got[0][i, :c], out[i, :c]["row"])
                assert np.array_equal(bits(got[1][i, :c]), bits(out[i, :c]["score"]))


def test_block_skipping_device_entry_and_threshold_pass(vg):
    """vg_index_search_blocks_dev with a device row bitmap, on data whose certificates fail (15000 identical rows): the
    threshold pass walks the same tile list."""
    import torch

    L = vg._lib
    n, dim, nq, k = 60_000, 128, 32, 10
    rng = np.random.default_rng(5)
    x = rng.standard_normal((n, dim)).astype(F)
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.Train(x)
    codes = sq.EncodeBatch(x)
    codes[5000:20000] = codes[5000]    # ~470 groups of 32 tied rows: more than any candidate budget (64 groups with kind::i8)
    q = (x[5000] + 0.01 * rng.standard_normal((nq, dim))).astype(F)
    keep = np.ones(n // BLOCK, bool)
    keep[5:7] = False              # rows 5120 .. 7167 of the tied range are jumped over
    keep[20:50] = False
    tomb = rng.random(n) < 0.95
    eq = equivalent_row_mask(n, keep, tomb)
    dev = torch.device("cuda:0")
    L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
    try:
        with vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales)) as ix:
            ix.upload(codes=codes)
            dq = torch.from_numpy(q).to(dev)
            dm = torch.from_numpy(np.packbits(tomb, bitorder="little")).to(dev)
            r = torch.empty((nq, k), dtype=torch.int32, device=dev)
            s = torch.empty((nq, k), dtype=torch.float32, device=dev)
            c = torch.empty((nq,), dtype=torch.int32, device=dev)
            ix.search_blocks_dev(dq.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr(), keep, d_mask=dm.data_ptr())
            torch.cuda.synchronize()
            st = L.last_search_stats()
    finally:
        L.call("vg_set_stream", 0xFFFFFFFFFFFFFFFF)
    assert st["second_chance_queries"] > 0 and st["exact_rerun_queries"] == 0, st
    seg = o.FlatOracle(dim=dim, metric=0, quant=1, codes=codes, mins=sq.mins, inv=sq.invScales)
    out, cnt = seg.search_batch(q, k, threads=8, mask=np.packbits(eq, bitorder="little"))
    rows = r.cpu().numpy().view(np.uint32)
    for i in range(nq):
        assert int(c[i]) == int(cnt[i]) == k
        assert np.array_equal(rows[i], out[i]["row"])
        assert np.array_equal(bits(s[i].cpu().numpy()), bits(out[i]["score"]))
