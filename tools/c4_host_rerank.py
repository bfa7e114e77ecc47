"""C4 shard (RaBitQ top-1000 over 12.5M x 1536-d rows, exact float32 rerank to the final top-100) with the float32 rows
in HOST memory instead of HBM (SURVEY hard part 6: 100M x 1536 float32 = 614 GB does not fit 8 x 180 GB
next to anything else, and a single-GPU deployment keeps only the 196-byte codes on the device).

The GPU box has 196 GB of host RAM, so the full 614 GB region cannot exist here; this run holds ONE shard's rows
(12.5M x 1536 x 4 = 76.8 GB) page-locked on the host and compares, on the same codes and queries,

    device-resident rerank   vg_index_upload_dev(d_vectors=...)      rows gathered from HBM
    host-resident rerank     vg_index_set_host_vectors(h_vectors)    rows gathered over the host link (zero-copy)

Both must return the same ids and score bits; the JSON line reports the step times and the host-link rate the gather
reaches (R x dim x 4 bytes per query).

    python tools/c4_host_rerank.py [--rows N] [--queries Q]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vecgo_b200 as vg

L = vg._lib


def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    n, dim, nq, r_top, k = arg("--rows", 12_500_000), 1536, arg("--queries", 1000), 1000, 100
    dev = torch.device("cuda:0")
    L.call("vg_init", 0)
    L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
    t0 = time.time()
    host = torch.empty((n, dim), dtype=torch.float32, pin_memory=True)
    pin_s = time.time() - t0
    code_bytes = dim // 8 + 4
    a = vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=n)   # rows in HBM
    b = vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=n)   # rows on the host
    chunk = 1 << 18
    codes = torch.empty((chunk, code_bytes), dtype=torch.uint8, device=dev)
    t0 = time.time()
    for r0 in range(0, n, chunk):
        mm = min(chunk, n - r0)
        g = torch.Generator(device=dev).manual_seed(4242 + r0 // chunk)
        x = torch.randn((mm, dim), dtype=torch.float32, device=dev, generator=g)
        L.call("vg_rabitq_encode_dev", x.data_ptr(), mm, dim, codes.data_ptr())
        a.upload_dev(mm, d_codes=codes.data_ptr(), d_vectors=x.data_ptr(), row0=r0)
        b.upload_dev(mm, d_codes=codes.data_ptr(), row0=r0)
        host[r0:r0 + mm].copy_(x)
        del x
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    L.call("vg_index_set_host_vectors", b.handle, host.data_ptr(), n)
    g = torch.Generator(device=dev).manual_seed(7)
    q = torch.randn((nq, dim), dtype=torch.float32, device=dev, generator=g)
    out = {}
    res = {}
    rr = torch.empty((nq, r_top), dtype=torch.int32, device=dev)
    ss = torch.empty((nq, r_top), dtype=torch.float32, device=dev)
    cc = torch.empty((nq,), dtype=torch.int32, device=dev)
    ex = torch.empty((nq, r_top), dtype=torch.float32, device=dev)

    def step(ix):
        # Segment.Search(top-R) -> Segment.Rerank -> final top-k by (exact score, row); non-negative L2 scores order as integers
        ix.search_dev(q.data_ptr(), nq, r_top, rr.data_ptr(), ss.data_ptr(), cc.data_ptr())
        L.call("vg_index_rerank_dev", ix.handle, q.data_ptr(), nq, rr.data_ptr(), r_top, ex.data_ptr())
        key = (ex.view(torch.int32).to(torch.int64) << 32) | (rr.to(torch.int64) & 0xFFFFFFFF)
        best = torch.topk(key, k, dim=1, largest=False, sorted=True).values
        return (best & 0xFFFFFFFF).to(torch.int32), (best >> 32).to(torch.int32)

    for name, ix in (("device_resident", a), ("host_resident", b)):
        times = []
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rows, sc = step(ix)
            e1.record()
            torch.cuda.synchronize()
            if i:
                times.append(e0.elapsed_time(e1))
        out[name] = {"ms_per_step": float(np.median(times)), "queries_per_s": nq / (float(np.median(times)) * 1e-3)}
        res[name] = (rows, sc)
    # scan only (no rerank), to separate the gather's share
    times = []
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.search_dev(q.data_ptr(), nq, r_top, rr.data_ptr(), ss.data_ptr(), cc.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        if i:
            times.append(e0.elapsed_time(e1))
    scan_ms = float(np.median(times))
    same = bool(torch.equal(res["device_resident"][0], res["host_resident"][0]) and
                torch.equal(res["device_resident"][1], res["host_resident"][1]))
    gather_bytes = nq * r_top * dim * 4
    host_ms = out["host_resident"]["ms_per_step"] - scan_ms
    print(json.dumps({"workload": f"C4 shard: RaBitQ top-{r_top} over {n} x {dim}-d rows + exact float32 rerank to top-{k}, {nq} queries/step",
                      "device_resident": out["device_resident"], "host_resident": out["host_resident"], "scan_only_ms": scan_ms,
                      "results_identical": same, "host_gather_bytes_per_step": gather_bytes,
                      "host_gather_ms": host_ms, "host_link_GBps": gather_bytes / (host_ms * 1e-3) / 1e9 if host_ms > 0 else None,
                      "host_region_GB": n * dim * 4 / 1e9, "pin_s": pin_s, "build_s": gen_s,
                      "note": "614 GB for 100M rows exceeds this box's 196 GB of host RAM; one shard's rows (1/8) are held instead"}), flush=True)
    a.close()
    b.close()
    scan_rows = arg("--scan-rows", 0)
    if scan_rows:
        # the whole 100M-row code section (19.6 GB) on ONE GPU, scan only: what a single-GPU deployment with host-resident
        # float32 rows spends before the rerank gather (whose cost per query does not depend on the region's size)
        del host
        torch.cuda.empty_cache()
        c = vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=scan_rows)
        for r0 in range(0, scan_rows, chunk):
            mm = min(chunk, scan_rows - r0)
            g = torch.Generator(device=dev).manual_seed(4242 + r0 // chunk)
            x = torch.randn((mm, dim), dtype=torch.float32, device=dev, generator=g)
            L.call("vg_rabitq_encode_dev", x.data_ptr(), mm, dim, codes.data_ptr())
            c.upload_dev(mm, d_codes=codes.data_ptr(), row0=r0)
            del x
        times = []
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            c.search_dev(q.data_ptr(), nq, r_top, rr.data_ptr(), ss.data_ptr(), cc.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            if i:
                times.append(e0.elapsed_time(e1))
        st = L.last_search_stats()
        ms = float(np.median(times))
        print(json.dumps({"workload": f"RaBitQ top-{r_top} scan only, {scan_rows} x {dim}-d codes ({scan_rows * code_bytes / 1e9:.1f} GB) on one GPU, {nq} queries/step",
                          "ms_per_step": ms, "queries_per_s": nq / ms * 1e3, "exact_rerun_queries": st["exact_rerun_queries"],
                          "with_host_gather_projected": {"ms_per_step": ms + host_ms, "queries_per_s": nq / (ms + host_ms) * 1e3,
                                                         "note": "scan time at 100M rows + the host-link gather measured above on a 76.8 GB region; arithmetic, "
                                                                 "not a measurement: 614 GB of float32 rows do not fit this box's host memory"}}), flush=True)
        c.close()


if __name__ == "__main__":
    main()
