"""Times ProductQuantizer.Train from host vectors and vg_pq_train_dev twice each (cold / warm) at the C5 shape."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vecgo_b200 as vg  # noqa: E402
from vecgo_b200 import _lib as L  # noqa: E402

n, dim, m, iters = 1_000_000, 768, 96, 25
L.call("vg_init", 0)
L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
x = np.random.default_rng(42).standard_normal((n, dim), dtype=np.float32)
dx = torch.from_numpy(x).cuda()
ds = dim // m
cb, sc, of = np.zeros(m * 256 * ds, np.int8), np.zeros(m, np.float32), np.zeros(m, np.float32)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    L.call("vg_pq_train_dev", dx.data_ptr(), n, dim, m, 256, iters, 1, L.ptr(cb, L.i8p), L.ptr(sc, L.f32p), L.ptr(of, L.f32p), None)
    torch.cuda.synchronize()
    print(f"device-resident run {rep}: {time.time() - t0:.3f} s", flush=True)
for rep in range(2):
    pq_ = vg.quantization.ProductQuantizer(dim, m, 256)
    torch.cuda.synchronize()
    t0 = time.time()
    pq_.Train(x, iters=iters, seed=1)
    torch.cuda.synchronize()
    print(f"host-vector run {rep}: {time.time() - t0:.3f} s", flush=True)
