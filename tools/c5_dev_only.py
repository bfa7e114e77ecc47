"""One device-resident C5 training (vg_pq_train_dev, 1M x 768, 96 x 256, 25 iterations) — the command the per-kernel
launch list of the training is taken from (ncu --metrics gpu__time_duration.sum)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vecgo_b200 as vg  # noqa: E402
from vecgo_b200 import _lib as L  # noqa: E402

n, dim, m, iters = (int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000), 768, 96, 25
L.call("vg_init", 0)
L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
x = torch.randn((n, dim), dtype=torch.float32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(42))
ds = dim // m
cb, sc, of = np.zeros(m * 256 * ds, np.int8), np.zeros(m, np.float32), np.zeros(m, np.float32)
torch.cuda.synchronize()
L.call("vg_pq_train_dev", x.data_ptr(), n, dim, m, 256, iters, 1, L.ptr(cb, L.i8p), L.ptr(sc, L.f32p), L.ptr(of, L.f32p), None)
torch.cuda.synchronize()
print("done", int(np.abs(cb.astype(np.int32)).sum()))
