"""Quick device-resident timing probe (not the bench): python tools/probe.py [codec ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import vecgo_b200 as vg

L = vg._lib


def timed(fn, warm=1, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    which = sys.argv[1:] or ["sq8", "int4", "f32", "pq", "rabitq"]
    dev = torch.device("cuda:0")
    L.call("vg_init", 0)
    L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device=dev).manual_seed(1)
    for name in which:
        if name == "sq8":
            n, dim, nq, k = 1_000_000, 768, 2048, 100
            codes = torch.randint(0, 256, (n, dim), dtype=torch.uint8, device=dev, generator=g)
            ix = vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n,
                                      sq8=(np.full(dim, -4, np.float32), np.full(dim, 8 / 255, np.float32)))
            ix.upload_dev(n, d_codes=codes.data_ptr())
            bytes_per_row = dim
        elif name == "int4":
            n, dim, nq, k = 1_000_000, 768, 2048, 100
            codes = torch.randint(0, 256, (n, dim // 2), dtype=torch.uint8, device=dev, generator=g)
            ix = vg.index.DeviceIndex(codec=L.CODEC_INT4, metric=0, dim=dim, rows=n,
                                      int4=(np.full(dim, -4, np.float32), np.full(dim, 8, np.float32)))
            ix.upload_dev(n, d_codes=codes.data_ptr())
            bytes_per_row = dim // 2
        elif name == "f32":
            n, dim, nq, k = 100_000, 128, 1000, 10
            codes = torch.rand((n, dim), dtype=torch.float32, device=dev, generator=g)
            ix = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=0, dim=dim, rows=n)
            ix.upload_dev(n, d_vectors=codes.data_ptr())
            bytes_per_row = dim * 4
        elif name == "pq":
            n, dim, m, nq, k = 4_000_000, 768, 96, 512, 100
            codes = torch.randint(0, 256, (n, m), dtype=torch.uint8, device=dev, generator=g)
            rng = np.random.default_rng(0)
            ix = vg.index.DeviceIndex(codec=L.CODEC_PQ, metric=0, dim=dim, rows=n,
                                      pq=(rng.integers(-128, 128, m * 256 * 8, dtype=np.int8), np.full(m, 0.01, np.float32),
                                          np.zeros(m, np.float32), m, 256))
            ix.upload_dev(n, d_codes=codes.data_ptr())
            bytes_per_row = m
        elif name == "rabitq":
            n, dim, nq, k = 4_000_000, 1536, 512, 1000
            codes = torch.randint(0, 256, (n, 196), dtype=torch.uint8, device=dev, generator=g)
            codes[:, 192:] = torch.tensor([0, 0, 0x80, 0x3F], dtype=torch.uint8, device=dev)  # norm = 1.0
            ix = vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=n)
            ix.upload_dev(n, d_codes=codes.data_ptr())
            bytes_per_row = 196
        else:
            continue
        q = torch.randn((nq, dim), dtype=torch.float32, device=dev, generator=g)
        rows = torch.empty((nq, k), dtype=torch.int32, device=dev)
        scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
        counts = torch.empty((nq,), dtype=torch.int32, device=dev)
        t0 = time.time()
        ms = timed(lambda: ix.search_dev(q.data_ptr(), nq, k, rows.data_ptr(), scores.data_ptr(), counts.data_ptr()))
        pairs = n * nq
        print(f"{name}: n={n} dim={dim} nq={nq} k={k}: {ms:.2f} ms  {pairs / ms / 1e6:.1f} Gpairs/s  "
              f"algorithmic {pairs * bytes_per_row / ms / 1e6:.0f} GB/s  qps@n={nq / ms * 1e3:.0f}  wall {time.time() - t0:.1f}s",
              flush=True)
        ix.close()
        del codes


if __name__ == "__main__":
    main()
