"""C5 (PQ codebook training, 1M x 768-d, 96 x 256 x 8, k-means++ + 25 Lloyd iterations) for a subspace range, once:
the command an ncu launch list is taken of to see which kernel the time goes to (python tools/c5_breakdown.py LO HI)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vecgo_b200 as vg

L = vg._lib
lo, hi = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n, dim, m, k = 1_000_000, 768, 96, 256
dev = torch.device("cuda:0")
L.call("vg_init", 0)
g = torch.Generator(device=dev).manual_seed(42)
x = torch.randn((n, dim), device=dev, generator=g)
G = hi - lo
cb, sc, of = np.zeros(G * k * (dim // m), np.int8), np.zeros(G, np.float32), np.zeros(G, np.float32)
for r in range(reps):
    torch.cuda.synchronize()
    t0 = time.time()
    L.call("vg_pq_train_range_dev", x.data_ptr(), n, dim, m, k, 25, 7, lo, hi, L.ptr(cb, L.i8p), L.ptr(sc, L.f32p), L.ptr(of, L.f32p), None)
    torch.cuda.synchronize()
    print(f"subspaces [{lo},{hi}): {time.time() - t0:.4f} s", flush=True)
