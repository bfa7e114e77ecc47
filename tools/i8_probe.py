"""SQ8 filter through tcgen05 kind::i8 vs the fp16 decode-GEMM: same results, step times, certificate statistics.
python tools/i8_probe.py [rows] [queries] [k] [dim]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vecgo_b200 as vg

L = vg._lib
F = np.float32
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dim = int(sys.argv[4]) if len(sys.argv) > 4 else 768
CHUNK = 1 << 18
dev = torch.device("cuda:0")
L.call("vg_init", 0)
L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device=dev).manual_seed(1)
x0 = torch.randn((min(CHUNK, n), dim), device=dev, generator=g)
mins, maxs = np.zeros(dim, F), np.zeros(dim, F)
L.call("vg_minmax_dev", x0.data_ptr(), x0.shape[0], dim, L.ptr(mins, L.f32p), L.ptr(maxs, L.f32p))
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
res = {}
for mode in (0, 1, -1):
    if mode >= 0:
        L.call("vg_quant_tc_i8_enable", mode)
    else:
        L.call("vg_flat_tc_enable", 0)
    nqm = nq if mode >= 0 else min(nq, 64)
    r = torch.empty((nqm, k), dtype=torch.int32, device=dev)
    s = torch.empty((nqm, k), dtype=torch.float32, device=dev)
    c_ = torch.empty((nqm,), dtype=torch.int32, device=dev)
    ts = []
    for i in range(4 if mode >= 0 else 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ix.search_dev(q.data_ptr(), nqm, k, r.data_ptr(), s.data_ptr(), c_.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    st = L.last_search_stats()
    res[mode] = (r, s)
    name = {0: "fp16 decode-GEMM", 1: "kind::i8", -1: "exact scan (64 queries)"}[mode]
    print(json.dumps({"mode": name, "rows": n, "queries": nqm, "k": k, "ms": float(np.median(ts[1:] or ts)), "ms_first": ts[0],
                      "second_chance": st["second_chance_queries"], "exact_rerun": st["exact_rerun_queries"],
                      "threshold_pass": st["threshold_pass_queries"]}), flush=True)
L.call("vg_flat_tc_enable", 1)
same01 = bool(torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1].view(torch.int32), res[1][1].view(torch.int32)))
m = res[-1][0].shape[0]
same1x = bool(torch.equal(res[1][0][:m], res[-1][0]) and torch.equal(res[1][1][:m].view(torch.int32), res[-1][1].view(torch.int32)))
print(json.dumps({"i8_identical_to_fp16": same01, "i8_identical_to_exact_scan": same1x}))
ix.close()
