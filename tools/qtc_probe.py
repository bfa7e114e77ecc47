"""Per-call timing of the tensor-core quantized scans (diagnostic): python tools/qtc_probe.py sq8|int4|pq [rows] [queries] [k] [calls]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vecgo_b200 as vg

L = vg._lib
codec = sys.argv[1] if len(sys.argv) > 1 else "sq8"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
k = int(sys.argv[4]) if len(sys.argv) > 4 else 100
calls = int(sys.argv[5]) if len(sys.argv) > 5 else 6
dim, m = 768, 96
L.call("vg_init", 0)
L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
if codec == "sq8":
    ix = vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(np.full(dim, -4, np.float32), np.full(dim, 8 / 255, np.float32)))
    cb = dim
elif codec == "int4":
    ix = vg.index.DeviceIndex(codec=L.CODEC_INT4, metric=0, dim=dim, rows=n, int4=(np.full(dim, -4, np.float32), np.full(dim, 8, np.float32)))
    cb = dim // 2
else:
    rng = np.random.default_rng(0)
    ix = vg.index.DeviceIndex(codec=L.CODEC_PQ, metric=0, dim=dim, rows=n,
                              pq=(rng.integers(-128, 128, m * 256 * (dim // m), dtype=np.int8), np.full(m, 0.01, np.float32), np.zeros(m, np.float32), m, 256))
    cb = m
chunk = 1 << 20
for r0 in range(0, n, chunk):
    mm = min(chunk, n - r0)
    codes = torch.randint(0, 256, (mm, cb), dtype=torch.uint8, device=dev, generator=g)
    ix.upload_dev(mm, d_codes=codes.data_ptr(), row0=r0)
q = torch.randn((nq, dim), dtype=torch.float32, device=dev, generator=g)
r = torch.empty((nq, k), dtype=torch.int32, device=dev)
s = torch.empty((nq, k), dtype=torch.float32, device=dev)
c = torch.empty((nq,), dtype=torch.int32, device=dev)
ms, nl = C.c_double(), C.c_uint64()
for i in range(calls):
    L.call("vg_quant_tc_profile", 1, C.byref(ms), C.byref(nl))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ix.search_dev(q.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr())
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    L.call("vg_quant_tc_profile", 0, C.byref(ms), C.byref(nl))
    print(f"{codec} call {i}: wall {1e3 * (t1 - t0):8.2f} ms, gemm {ms.value:8.2f} ms in {nl.value} launch(es)", flush=True)
ix.close()
