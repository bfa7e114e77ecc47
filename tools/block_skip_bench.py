"""What block-stat skipping buys on the device (flat/segment.go:524-541): the C2a shape (SQ8, 10M x 768 rows,
10 000 queries, k = 100) searched with block verdicts that keep 100 % / 50 % / 10 % of the 1024-row blocks, clustered (one
contiguous range: a time-ordered field) and random, with tile skipping on and off (off = the row bitmap only masks scores in
the epilogue; every tile is still fetched, decoded and multiplied).  One JSON line per case; results are compared between
the two modes (ids and score bits).

    python tools/block_skip_bench.py [--small]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vecgo_b200 as vg

L = vg._lib
F = np.float32
SMALL = "--small" in sys.argv
CHUNK = 1 << 18
dev = torch.device("cuda:0")


def main():
    n, dim, nq, k = (1_000_000, 768, 2048, 100) if SMALL else (10_000_000, 768, 10_000, 100)
    L.call("vg_init", 0)
    L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device=dev).manual_seed(1)
    x0 = torch.randn((CHUNK, dim), device=dev, generator=g)
    mins, maxs = np.zeros(dim, F), np.zeros(dim, F)
    L.call("vg_minmax_dev", x0.data_ptr(), CHUNK, dim, L.ptr(mins, L.f32p), L.ptr(maxs, L.f32p))
    sq = vg.quantization.ScalarQuantizer(dim)
    sq.SetBounds(mins, maxs)
    ix = vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales))
    codes = torch.empty((CHUNK, dim), dtype=torch.uint8, device=dev)
    for c in range((n + CHUNK - 1) // CHUNK):
        m = min(CHUNK, n - c * CHUNK)
        x = x0 if c == 0 else torch.randn((CHUNK, dim), device=dev, generator=g)
        L.call("vg_sq8_encode_dev", x.data_ptr(), m, dim, L.ptr(sq.mins, L.f32p), L.ptr(sq.maxs, L.f32p), L.ptr(sq.scales, L.f32p), codes.data_ptr())
        ix.upload_dev(m, d_codes=codes.data_ptr(), row0=c * CHUNK)
    q = torch.randn((nq, dim), device=dev, generator=g)
    r = torch.empty((nq, k), dtype=torch.int32, device=dev)
    s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    c_ = torch.empty((nq,), dtype=torch.int32, device=dev)
    full = n // 1024
    rng = np.random.default_rng(3)

    def timed(fn, steps=4):
        ts = []
        for i in range(steps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i:
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    base = timed(lambda: ix.search_dev(q.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c_.data_ptr()))
    print(json.dumps({"case": "no filter", "rows": n, "queries": nq, "k": k, "ms": base, "queries_per_s": nq / base * 1e3}), flush=True)
    for frac in (1.0, 0.5, 0.1):
        for kind in ("clustered", "random"):
            keep = np.zeros(full, bool)
            nk = int(full * frac)
            if kind == "clustered":
                keep[full // 4: full // 4 + nk] = True
                if nk > full - full // 4:
                    keep[:] = True
            else:
                keep[rng.permutation(full)[:nk]] = True
            if frac == 1.0 and kind == "random":
                continue
            out = {}
            res = {}
            for mode in (1, 0):
                L.call("vg_tile_skip_enable", mode)
                ms = timed(lambda: ix.search_blocks_dev(q.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c_.data_ptr(), keep))
                st = L.last_search_stats()
                out["tile_skip_on" if mode else "tile_skip_off"] = {"ms": ms, "queries_per_s": nq / ms * 1e3}
                res[mode] = (r.clone(), s.clone())
            L.call("vg_tile_skip_enable", 1)
            same = bool(torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1].view(torch.int32), res[1][1].view(torch.int32)))
            print(json.dumps({"case": f"{int(frac * 100)} % of the blocks kept, {kind}", "blocks": full, "kept": int(keep.sum()),
                              **out, "speedup": out["tile_skip_off"]["ms"] / out["tile_skip_on"]["ms"],
                              "vs_no_filter": out["tile_skip_on"]["ms"] / base, "identical": same,
                              "distance_computations_per_query": st["distance_computations"] // nq,
                              "second_chance": st["second_chance_queries"], "exact_rerun": st["exact_rerun_queries"]}), flush=True)
    ix.close()


if __name__ == "__main__":
    main()
