"""Certificate robustness of the tensor-core filters on non-benign data: re-run rates and p50 / p99 step
times of SQ8 / INT4 (C2 shape), PQ (C3 shard shape) and RaBitQ (C4 shard shape, top-1000) on

    gaussian    N(0,1) rows and queries (the bench's data: the baseline every other row is compared with)
    clustered   1000 Gaussian blobs, sigma = 0.05 around N(0,1) centres; queries are fresh points of the same blobs
    unit        unit-normalised Gaussian rows and queries (testutil.UnitVectors)
    duplicates  N(0,1) with 1 % of the rows exact copies of other rows; 10 % of the queries ARE database rows
    lowdim      rank-16 data: x = z A, z ~ N(0, I_16), A a fixed 16 x dim matrix

(integration_test/quantization_recall_test.go:17-117 data kinds, testutil/testutil.go:69-175).  Every row of the table
is produced by encoding real vectors with a quantizer trained on that data (not random codes).  One JSON line per
(codec, data) pair; a query whose certificate fails gets the second chance (2x candidate groups) and then the exact
CUDA-core scan — the counters come from vg_last_search_stats.

    python tools/certificate_robustness.py [sq8 int4 pq rabitq] [--small] [--kinds gaussian,clustered,...]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vecgo_b200 as vg

L = vg._lib
F = np.float32
SMALL = "--small" in sys.argv
KINDS = ["gaussian", "clustered", "unit", "duplicates", "lowdim"]
for i, a_ in enumerate(sys.argv):
    if a_ == "--kinds":
        KINDS = sys.argv[i + 1].split(",")
CHUNK = 1 << 18
dev = torch.device("cuda:0")


class Data:
    """Row chunks and queries of one data kind, generated on the device from per-chunk seeds."""

    def __init__(self, kind, dim):
        self.kind, self.dim = kind, dim
        g = torch.Generator(device=dev).manual_seed(1234)
        if kind == "clustered":
            self.centres = torch.randn((1000, dim), device=dev, generator=g)
        if kind == "lowdim":
            self.A = torch.randn((16, dim), device=dev, generator=g) / 4.0

    def _draw(self, n, g):
        dim = self.dim
        if self.kind == "clustered":
            c = torch.randint(0, 1000, (n,), device=dev, generator=g)
            return self.centres[c] + 0.05 * torch.randn((n, dim), device=dev, generator=g)
        if self.kind == "lowdim":
            return torch.randn((n, 16), device=dev, generator=g) @ self.A
        x = torch.randn((n, dim), device=dev, generator=g)
        if self.kind == "unit":
            x = x / x.norm(dim=1, keepdim=True)
        return x

    def chunk(self, c, n):
        g = torch.Generator(device=dev).manual_seed(10_000 + c)
        x = self._draw(n, g)
        if self.kind == "duplicates":
            nd = n // 100
            src = torch.randint(0, n, (nd,), device=dev, generator=g)
            dst = torch.randint(0, n, (nd,), device=dev, generator=g)
            x[dst] = x[src]
        return x.contiguous()

    def queries(self, nq):
        g = torch.Generator(device=dev).manual_seed(77)
        q = self._draw(nq, g)
        if self.kind == "duplicates":
            rows = self.chunk(0, CHUNK)
            q[: nq // 10] = rows[: nq // 10]
        return q.contiguous()


def build(codec, data, n, dim):
    x0 = data.chunk(0, CHUNK)
    mins, maxs = np.zeros(dim, F), np.zeros(dim, F)
    L.call("vg_minmax_dev", x0.data_ptr(), CHUNK, dim, L.ptr(mins, L.f32p), L.ptr(maxs, L.f32p))
    enc = None
    if codec == "sq8":
        sq = vg.quantization.ScalarQuantizer(dim)
        sq.SetBounds(mins, maxs)
        ix = vg.index.DeviceIndex(codec=L.CODEC_SQ8, metric=0, dim=dim, rows=n, sq8=(sq.mins, sq.invScales))
        cb = dim
        enc = lambda x, m, out: L.call("vg_sq8_encode_dev", x.data_ptr(), m, dim, L.ptr(sq.mins, L.f32p), L.ptr(sq.maxs, L.f32p), L.ptr(sq.scales, L.f32p), out.data_ptr())
    elif codec == "int4":
        diff = (maxs - mins).astype(F)
        diff[diff == 0] = 1.0
        ix = vg.index.DeviceIndex(codec=L.CODEC_INT4, metric=0, dim=dim, rows=n, int4=(mins, diff))
        cb = dim // 2
        enc = lambda x, m, out: L.call("vg_int4_encode_dev", x.data_ptr(), m, dim, L.ptr(mins, L.f32p), L.ptr(diff, L.f32p), out.data_ptr())
    elif codec == "pq":
        m_ = 96
        ds = dim // m_
        cbk, sc, of = np.zeros(m_ * 256 * ds, np.int8), np.zeros(m_, F), np.zeros(m_, F)
        L.call("vg_pq_train_dev", x0.data_ptr(), min(CHUNK, 131072), dim, m_, 256, 8, 1, L.ptr(cbk, L.i8p), L.ptr(sc, L.f32p), L.ptr(of, L.f32p), None)
        ix = vg.index.DeviceIndex(codec=L.CODEC_PQ, metric=0, dim=dim, rows=n, pq=(cbk, sc, of, m_, 256))
        cb = m_
        enc = lambda x, m, out: L.call("vg_pq_encode_dev", x.data_ptr(), m, dim, m_, 256, L.ptr(cbk, L.i8p), L.ptr(sc, L.f32p), L.ptr(of, L.f32p), out.data_ptr())
    else:
        ix = vg.index.DeviceIndex(codec=L.CODEC_RABITQ, metric=0, dim=dim, rows=n)
        cb = dim // 8 + 4
        enc = lambda x, m, out: L.call("vg_rabitq_encode_dev", x.data_ptr(), m, dim, out.data_ptr())
    codes = torch.empty((CHUNK, cb), dtype=torch.uint8, device=dev)
    for c in range((n + CHUNK - 1) // CHUNK):
        m = min(CHUNK, n - c * CHUNK)
        x = x0 if c == 0 else data.chunk(c, CHUNK)
        enc(x, m, codes)
        ix.upload_dev(m, d_codes=codes.data_ptr(), row0=c * CHUNK)
        del x
    torch.cuda.synchronize()
    return ix


def run(codec, kind):
    if codec in ("sq8", "int4"):
        n, dim, nq, k = (1_000_000, 768, 2048, 100) if SMALL else (10_000_000, 768, 10_000, 100)
    elif codec == "pq":
        n, dim, nq, k = (2_000_000, 768, 2048, 100) if SMALL else (25_000_000, 768, 10_000, 100)
    else:
        n, dim, nq, k = (1_000_000, 1536, 512, 1000) if SMALL else (12_500_000, 1536, 1000, 1000)
    data = Data(kind, dim)
    t0 = time.time()
    ix = build(codec, data, n, dim)
    build_s = time.time() - t0
    q = data.queries(nq)
    r = torch.empty((nq, k), dtype=torch.int32, device=dev)
    s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    c = torch.empty((nq,), dtype=torch.int32, device=dev)
    times, second, exact = [], [], []
    steps = 3 if SMALL else 8
    for i in range(steps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ix.search_dev(q.data_ptr(), nq, k, r.data_ptr(), s.data_ptr(), c.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        st = L.last_search_stats()
        if i > 0:
            times.append(e0.elapsed_time(e1))
            second.append(st["second_chance_queries"])
            exact.append(st["exact_rerun_queries"])
    # the certified result must equal the exact scan's on a sample of the queries (ids and score bits)
    nchk = min(64, nq)
    r2, s2, c2 = torch.empty_like(r[:nchk]), torch.empty_like(s[:nchk]), torch.empty_like(c[:nchk])
    L.call("vg_flat_tc_enable", 0)
    ix.search_dev(q.data_ptr(), nchk, k, r2.data_ptr(), s2.data_ptr(), c2.data_ptr())
    L.call("vg_flat_tc_enable", 1)
    same = bool(torch.equal(r[:nchk], r2) and torch.equal(s[:nchk].view(torch.int32), s2.view(torch.int32)))
    ix.close()
    torch.cuda.empty_cache()
    print(json.dumps({"codec": codec, "data": kind, "rows": n, "dim": dim, "queries": nq, "k": k, "steps": steps,
                      "ms_p50": float(np.median(times)), "ms_p99": float(np.max(times)), "ms_min": float(np.min(times)),
                      "second_chance_queries_per_step": float(np.mean(second)), "exact_rerun_queries_per_step": float(np.mean(exact)),
                      "rerun_rate": float(np.mean(exact)) / nq, "identical_to_exact_scan_on_64_queries": same, "build_s": build_s}), flush=True)


def main():
    codecs = [a_ for a_ in sys.argv[1:] if a_ in ("sq8", "int4", "pq", "rabitq")] or ["sq8", "int4", "pq", "rabitq"]
    L.call("vg_init", 0)
    L.call("vg_set_stream", torch.cuda.current_stream().cuda_stream)
    for codec in codecs:
        for kind in KINDS:
            try:
                run(codec, kind)
            except Exception as ex:  # noqa: BLE001
                print(json.dumps({"codec": codec, "data": kind, "error": repr(ex)}), flush=True)
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
