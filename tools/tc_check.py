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
    gids = np.zeros((nq, kc), np.uint32)
    cnt = np.zeros(nq, np.int32)
    tau = np.zeros(nq, np.float32)
    G = C.c_int64()
    t0 = time.time()
    L.call("vg_flat_tc_candidates", ix.handle, L.ptr(q, L.f32p), nq, kc, L.ptr(gids, L.u32p), L.ptr(cnt, L.i32p), L.ptr(tau, L.f32p),
           C.byref(G))
    G = int(G.value)
    print(f"filter call: {time.time() - t0:.3f}s  groups of {G} rows, {-(-n // G)} groups, kc={kc}, listed min/max {cnt.min()}/{cnt.max()}", flush=True)
    # exact s in float64: group minima vs tau and the error bound
    x64, q64 = x.astype(np.float64), q.astype(np.float64)
    nchk = min(nq, 64)
    dots = q64[:nchk] @ x64.T
    s_true = (np.sum(x64 * x64, 1)[None, :] - 2 * dots) if metric == 0 else -dots
    qn, xn = np.sum(q64 * q64, 1), np.sum(x64 * x64, 1)
    c1 = (1 / 512 if metric != 0 else 1 / 256) * 1.125
    E = c1 * np.sqrt(qn[:nchk] * xn.max()) + (qn[:nchk] + xn.max()) / 16384
    ng = -(-n // G)
    pad = ng * G - n
    true_top = np.argsort(s_true, axis=1, kind="stable")[:, :k]
    inside, worst = 0.0, 0.0
    for i in range(nchk):
        gm = np.pad(s_true[i], (0, pad), constant_values=np.inf).reshape(ng, G).min(1)   # true group minima
        ent = gids[i, : cnt[i]].astype(np.int64)
        crowded = (ent & 0x80000000) != 0
        sel = np.where(crowded, ent & 0x7FFFFFFF, ent // G)
        rows_named = ent[~crowded]
        # the row a group minimum names must be (within the error bound) the true minimum of its group
        assert np.all(s_true[i, rows_named] <= gm[rows_named // G] + 2 * E[i]), "arg-min row is not the group minimum"
        # every group whose true minimum is below tau - E must be selected; no selected group may be above tau + E
        assert set(np.where(gm <= tau[i] - E[i])[0].tolist()) <= set(sel.tolist()), "a group far below tau was not selected"
        assert np.all(gm[sel] <= tau[i] + E[i]), "a selected group is far above tau"
        worst = max(worst, float(np.abs(np.sort(gm)[kc - 1] - tau[i]) / E[i]))
        inside += len(set((true_top[i] // G).tolist()) - set(sel.tolist())) == 0
    print(f"|tau - true kc-th group minimum| / E max = {worst:.4f}; true top-{k} rows inside the selected groups: {inside / nchk:.4f}")
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
