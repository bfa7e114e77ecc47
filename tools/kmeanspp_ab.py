"""A/B check of the exact parallel k-means++ prefix chain against the one-lane sequential chain (VECGO_KMEANSPP_SEQUENTIAL=1):
both must give bit-identical PQ codebooks.   python tools/kmeanspp_ab.py [n] [dim] [m]"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def train(n, dim, m):
    import vecgo_b200 as vg

    rng = np.random.default_rng(5)
    x = rng.standard_normal((n, dim)).astype(np.float32)
    x[: n // 10] = np.round(x[: n // 10] * 4) / 4      # many exact ties / powers of two in the distances
    x[n // 2: n // 2 + 500] = x[0]                      # zero distances once row 0's neighbourhood is chosen
    pq = vg.quantization.ProductQuantizer(dim, m, 256)
    pq.Train(x, iters=3, seed=11)
    return pq.codebooks.copy(), pq.scales.copy(), pq.offsets.copy()


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    m = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    if os.environ.get("KMEANSPP_AB_CHILD"):
        cb, sc, of = train(n, dim, m)
        np.savez(os.environ["KMEANSPP_AB_CHILD"], cb=cb, sc=sc, of=of)
        sys.exit(0)
    cb, sc, of = train(n, dim, m)
    out = "/tmp/kmeanspp_ab_child.npz"
    env = dict(os.environ, VECGO_KMEANSPP_SEQUENTIAL="1", KMEANSPP_AB_CHILD=out)
    subprocess.check_call([sys.executable, os.path.abspath(__file__), str(n), str(dim), str(m)], env=env)
    z = np.load(out)
    same = np.array_equal(cb, z["cb"]) and np.array_equal(sc.view(np.uint32), z["sc"].view(np.uint32)) and np.array_equal(of.view(np.uint32), z["of"].view(np.uint32))
    print("parallel exact prefix == sequential chain:", same)
    sys.exit(0 if same else 1)
