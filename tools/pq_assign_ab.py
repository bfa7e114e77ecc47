"""A/B check of the tensor-core PQ-training assignment (vg_pq_assign_tc.cu) against the exact CUDA-core assignment
(VECGO_PQ_ASSIGN_TC=0): float32 centroids and int8 codebooks of the whole training must be bit-identical.
   python tools/pq_assign_ab.py [n] [dim] [m] [iters] [kind]     kind: gauss | ties | scaled"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def train(n, dim, m, iters, kind):
    import torch

    import vecgo_b200 as vg
    from vecgo_b200 import _lib as L

    L.call("vg_init", 0)
    L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((n, dim), dtype=torch.float32, device="cuda", generator=g)
    if kind == "ties":      # coarse grid: many exactly equal distances (first-wins ties) and duplicate samples
        x = torch.round(x * 2) / 2
        x[n // 2: n // 2 + 5000] = x[0]
    elif kind == "scaled":  # large dynamic range between subspaces and a far offset
        x = x * torch.logspace(-3, 3, dim, device="cuda")[None, :] + 100.0
    ds = dim // m
    cb, sc, of = np.zeros(m * 256 * ds, np.int8), np.zeros(m, np.float32), np.zeros(m, np.float32)
    cen = np.zeros(m * 256 * ds, np.float32)
    torch.cuda.synchronize()
    t0 = time.time()
    L.call("vg_pq_train_dev", x.data_ptr(), n, dim, m, 256, iters, 11, L.ptr(cb, L.i8p), L.ptr(sc, L.f32p), L.ptr(of, L.f32p), L.ptr(cen, L.f32p))
    torch.cuda.synchronize()
    secs = time.time() - t0
    import ctypes as C

    a, b = C.c_uint64(), C.c_uint64()
    L.call("vg_pq_assign_tc_stats", C.byref(a), C.byref(b))
    if os.environ.get("VECGO_PQ_ASSIGN_TC") != "0":
        assert a.value > 0, "the tensor-core assignment did not run"
        print(f"tensor-core assignment: {a.value} pairs, {b.value} re-evaluated exactly ({100.0 * b.value / a.value:.2f} %)")
    else:
        assert a.value == 0
    return cb, sc, of, cen, secs


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    m = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    kind = sys.argv[5] if len(sys.argv) > 5 else "gauss"
    child = os.environ.get("PQ_ASSIGN_AB_CHILD")
    cb, sc, of, cen, secs = train(n, dim, m, iters, kind)
    if child:
        np.savez(child, cb=cb, sc=sc, of=of, cen=cen, secs=secs)
        sys.exit(0)
    out = "/tmp/pq_assign_ab_child.npz"
    env = dict(os.environ, VECGO_PQ_ASSIGN_TC="0", PQ_ASSIGN_AB_CHILD=out)
    subprocess.check_call([sys.executable, os.path.abspath(__file__), str(n), str(dim), str(m), str(iters), kind], env=env)
    z = np.load(out)
    same = (np.array_equal(cen.view(np.uint32), z["cen"].view(np.uint32)) and np.array_equal(cb, z["cb"])
            and np.array_equal(sc.view(np.uint32), z["sc"].view(np.uint32)) and np.array_equal(of.view(np.uint32), z["of"].view(np.uint32)))
    print(f"{kind} n={n} dim={dim} m={m} iters={iters}: tensor-core assignment == exact assignment: {same}   "
          f"({secs:.3f} s vs {float(z['secs']):.3f} s, first call includes context start-up)")
    sys.exit(0 if same else 1)
