"""C1 (Flat 100k x 128, 1000 queries, k = 10) timing probe: launch-only search loop with CUDA events, hot (L2-resident)
and cold (L2 flushed before every step), result checked against the exact CUDA-core scan."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vecgo_b200 as vg

L = vg._lib
n, dim, nq, k = [int(x) for x in (sys.argv[1:5] + [100000, 128, 1000, 10][len(sys.argv[1:5]):])]
rng = np.random.default_rng(42)
x = rng.random((n, dim), dtype=np.float32)
q = np.random.default_rng(43).random((nq, dim), dtype=np.float32)
dev = torch.device("cuda:0")
L.call("vg_init", 0)
L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
ix = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n)
ix.upload(vectors=x)
dq = torch.from_numpy(q).to(dev)
r = torch.empty((nq, k), dtype=torch.int32, device=dev)
s = torch.empty((nq, k), dtype=torch.float32, device=dev)
c = torch.empty((nq,), dtype=torch.int32, device=dev)
f = torch.zeros((nq,), dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step():
    ix.search_dev_async(dq.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr(), f.data_ptr())


for _ in range(5):
    step()
torch.cuda.synchronize()
l0 = vg.launch_count()
step()
torch.cuda.synchronize()
print("launches per step:", vg.launch_count() - l0, "flags set:", int(f.sum().item()))
for mode in ("hot", "cold"):
    ts = []
    for _ in range(30):
        if mode == "cold":
            flush.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{mode}: median {np.median(ts):.1f} us, min {np.min(ts):.1f} us  ({2.0 * nq * n * dim / np.median(ts) / 1e6:.1f} TFLOP/s)")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    step()
e1.record()
torch.cuda.synchronize()
print(f"back to back: {e0.elapsed_time(e1) * 1e3 / 50:.1f} us per step")
r1, s1 = r.clone(), s.clone()
L.call("vg_flat_tc_enable", 0)
ix.search_dev(dq.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr())
L.call("vg_flat_tc_enable", 1)
print("identical to the exact scan:", bool(torch.equal(r1, r) and torch.equal(s1.view(torch.int32), s.view(torch.int32))))
