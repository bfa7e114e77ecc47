"""GPU diagnostic for the tcgen05 Flat filter: approximate-score error vs the certificate bound,
candidate-set correctness, end-to-end identity with the exact CUDA-core scan, timing.
    python tools/tc_check.py [n] [dim] [nq] [k] [metric]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vecgo_b200 as vg

L = vg._lib


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    metric = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    rng = np.random.default_rng(42)
    x = rng.random((n, dim), dtype=np.float32)
    q = np.random.default_rng(43).random((nq, dim), dtype=np.float32)
    if metric != 0:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    L.call("vg_init", 0)
    ix = vg.index.DeviceIndex(codec=L.CODEC_F32, metric=metric, dim=dim, rows=n)
    ix.upload(vectors=x)
    kc = 32 if k <= 16 else 2 * k
    if dim > 256:
        kc = max(kc, 64)
    CAP = 128
    rows = np.zeros((nq, CAP), np.uint32)
    s = np.zeros((nq, CAP), np.float32)
    cnt = np.zeros(nq, np.int32)
    tau = np.zeros(nq, np.float32)
    t0 = time.time()
    L.call("vg_flat_tc_candidates", ix.handle, L.ptr(q, L.f32p), nq, kc, L.ptr(rows, L.u32p), L.ptr(s, L.f32p), L.ptr(cnt, L.i32p),
           L.ptr(tau, L.f32p))
    print(f"candidates call: {time.time() - t0:.3f}s  survivors per query min/mean/max {cnt.min()}/{cnt.mean():.1f}/{cnt.max()} (kc={kc})", flush=True)
    # exact s in float64
    x64, q64 = x.astype(np.float64), q.astype(np.float64)
    nchk = min(nq, 64)
    dots = q64[:nchk] @ x64.T
    s_true = (np.sum(x64 * x64, 1)[None, :] - 2 * dots) if metric == 0 else -dots
    qn, xn = np.sum(q64 * q64, 1), np.sum(x64 * x64, 1)
    c1 = (1 / 512 if metric != 0 else 1 / 256) * 1.125
    E = c1 * np.sqrt(qn[:nchk] * xn.max()) + (qn[:nchk] + xn.max()) / 16384
    worst = 0.0
    inside = 0.0
    true_top = np.argsort(s_true, axis=1, kind="stable")[:, :k]
    for i in range(nchk):
        c = min(int(cnt[i]), CAP)
        err = np.abs(s_true[i, rows[i, :c].astype(np.int64)] - s[i, :c])
        worst = max(worst, float(err.max() / E[i]))
        inside += len(set(true_top[i]) & set(rows[i, :c].tolist())) / k
        # the list must be exactly the rows with approximate s <= tau: check with the true s and the error bound
        must = np.where(s_true[i] <= tau[i] - E[i])[0]
        assert set(must.tolist()) <= set(rows[i, :c].tolist()), "a row far below tau is missing from the candidate list"
    print(f"max err/E = {worst:.4f}; true top-{k} contained in candidates: {inside / nchk:.4f}")
    # end-to-end identity with the exact scan
    qa, fb = C.c_uint64(), C.c_uint64()
    L.call("vg_flat_tc_enable", 1)
    r1, s1, c1_ = ix.search(q, k)
    L.call("vg_flat_tc_stats", C.byref(qa), C.byref(fb))
    L.call("vg_flat_tc_enable", 0)
    r0, s0, c0 = ix.search(q, k)
    same_rows = np.array_equal(r0, r1)
    same_scores = np.array_equal(s0.view(np.uint32), s1.view(np.uint32))
    print(f"tc vs exact scan: rows identical={same_rows} scores bit-identical={same_scores} counts={np.array_equal(c0, c1_)}; "
          f"tc queries={qa.value} fallbacks={fb.value}")
    if not same_rows:
        bad = np.where((r0 != r1).any(1))[0]
        print("mismatching queries:", bad[:10], "of", len(bad))
        i = bad[0]
        print(r0[i], r1[i], s0[i], s1[i])
    # timing (device-resident)
    dq = torch.from_numpy(q).cuda()
    dr = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    ds = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    dc = torch.empty((nq,), dtype=torch.int32, device="cuda")
    L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
    for on in (1, 0):
        L.call("vg_flat_tc_enable", on)
        for _ in range(2):
            ix.search_dev(dq.data_ptr(), nq, k, dr.data_ptr(), ds.data_ptr(), dc.data_ptr())
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ix.search_dev(dq.data_ptr(), nq, k, dr.data_ptr(), ds.data_ptr(), dc.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        print(f"{'tcgen05 filter' if on else 'exact CUDA-core'}: {ms:.3f} ms  {nq / ms * 1e3:.0f} QPS  {2.0 * nq * n * dim / ms / 1e9:.1f} TFLOP/s (2QNd)")
    ix.close()


if __name__ == "__main__":
    main()
